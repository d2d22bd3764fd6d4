// ppp_oracle.cpp — CPU ORACLE. TEST INFRASTRUCTURE ONLY.
//
// A dependency-free C++17 restatement of the PCL 1.10 / FLANN 1.9.1 / Eigen 3.3 CPU path that
// tsai0507/PolishPathPlanning runs for: per-point radius / kNN search + PCA normals
// (estimate_normal), the PassThrough x-band (rangedX_index), left/right classification,
// pairing (gen-2 brute force "variant A", SectPath kd-tree "variant B"), interpolation onto the
// plane and the std::map ascending-y ordering (insert_point / path_track / OnePath).
//
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
// load this library, and only as the checker / the reported CPU baseline. The product
// (libppp_gpu.so) never links, loads or falls back to it.
//
// PARITY UNPINNED against the reference itself: PCL, FLANN, Eigen and GSL are not vendored in
// /root/reference, are not installed in the build image, and the reference ships no tests, no
// fixtures and no sample cloud.  The third-party semantics restated below are
// recalled from upstream PCL 1.10.0 / FLANN 1.9.1 / Eigen 3.3 and marked [upstream].  What pins
// this file instead: (1) tests/golden/*.npz produced by OpenCV's fork of the same FLANN
// KDTREE_SINGLE index (cv2.flann, float32) and scipy.spatial.cKDTree
// (tests/golden/make_golden.py), (2) a brute-force O(N^2) path inside this file that the
// kd-tree path must equal, (3) float64 numpy eigen-decomposition sanity bounds for normals.
//
// Reference call sites restated (relative to /root/reference):
//   estimate_normal      src/Path_Generation.cpp:323-333, src/contour_alg.cpp:142-151,
//                        src/slicing_method.cpp:204-214 (r = 2.5 / 3, NormalEstimation)
//   rangedX_index        src/Path_Generation.cpp:94-104, src/contour_alg.cpp:153-163
//   insert_point (A)     src/Path_Generation.cpp:107-206, src/slicing_method.cpp:112-202
//   insert_point (B)     src/contour_alg.cpp:165-237, src/Path_Alg/path_slicing_alg.cpp:164-237
//   path_track / OnePath src/Path_Generation.cpp:659-687, src/contour_alg.cpp:240-264
//   plane sweeps         src/Path_Generation.cpp:282-321,689-755, src/contour_alg.cpp:287-339
//
// Build: see oracle/Makefile (g++ -O2 -ffp-contract=off -fno-fast-math; FMA contraction pinned
// off because the reference's PCL binaries are built without FMA and float32 covariance sums are
// order- and rounding-sensitive at the 1e-3 level).

#include <algorithm>
#include <array>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <map>
#include <memory>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

namespace {

struct Pts {
  const float* p;
  int64_t n;
  int64_t sf;  // stride in floats (8 for pcl::PointXYZRGB, 3 or 4 for packed arrays)
  const float* at(int64_t i) const { return p + i * sf; }
};

inline bool finite3(const float* q) {
  return std::isfinite(q[0]) && std::isfinite(q[1]) && std::isfinite(q[2]);
}

// [upstream] flann::L2_Simple<float>::operator(): result = 0; for d: diff = a-b; result += diff*diff
inline float d2_flann(const float* a, const float* b) {
  float r = 0.0f, d;
  d = a[0] - b[0]; r += d * d;
  d = a[1] - b[1]; r += d * d;
  d = a[2] - b[2]; r += d * d;
  return r;
}

struct Key {
  float d2;
  int32_t idx;
};
inline bool key_less(const Key& a, const Key& b) {
  // [upstream] flann DistIndex::operator<; also the (d2, idx) total order that replaces
  // KNNSimpleResultSet's traversal-order ties.
  return (a.d2 < b.d2) || (a.d2 == b.d2 && a.idx < b.idx);
}

// ---------------------------------------------------------------------------------------------
// Exact kd-tree (stands in for flann::KDTreeSingleIndex, leaf_max_size 15, exact search).
// Any exact index returns the same answer under the (d2, idx) total order; this one exists so the
// timed CPU baseline has the same algorithmic shape as the reference's FLANN path.
// ---------------------------------------------------------------------------------------------
struct KdTree {
  struct Node {
    int32_t lo, hi;      // range in vind
    int32_t left, right; // children (-1 for leaf)
    float bmin[3], bmax[3];
  };
  std::vector<int32_t> vind;  // original indices, reordered
  std::vector<float> data;    // reordered xyz, 3 floats per point (FLANN reorder=true)
  std::vector<Node> nodes;
  static constexpr int kLeaf = 15;

  void build(const Pts& P, const int32_t* subset, int64_t m) {
    // [upstream] pcl::KdTreeFLANN::convertCloudToArray drops non-finite points.
    vind.clear();
    if (subset) {
      for (int64_t i = 0; i < m; i++)
        if (finite3(P.at(subset[i]))) vind.push_back(subset[i]);
    } else {
      for (int64_t i = 0; i < P.n; i++)
        if (finite3(P.at(i))) vind.push_back((int32_t)i);
    }
    nodes.clear();
    nodes.reserve(vind.size() / 4 + 16);
    if (!vind.empty()) build_rec(P, 0, (int32_t)vind.size());
    data.resize(vind.size() * 3);
    for (size_t i = 0; i < vind.size(); i++) {
      const float* q = P.at(vind[i]);
      data[3 * i] = q[0]; data[3 * i + 1] = q[1]; data[3 * i + 2] = q[2];
    }
  }

  int32_t build_rec(const Pts& P, int32_t lo, int32_t hi) {
    int32_t id = (int32_t)nodes.size();
    nodes.push_back(Node());
    Node nd;
    nd.lo = lo; nd.hi = hi; nd.left = nd.right = -1;
    for (int d = 0; d < 3; d++) { nd.bmin[d] = INFINITY; nd.bmax[d] = -INFINITY; }
    for (int32_t i = lo; i < hi; i++) {
      const float* q = P.at(vind[i]);
      for (int d = 0; d < 3; d++) { nd.bmin[d] = std::min(nd.bmin[d], q[d]); nd.bmax[d] = std::max(nd.bmax[d], q[d]); }
    }
    if (hi - lo > kLeaf) {
      int dim = 0; float span = nd.bmax[0] - nd.bmin[0];
      for (int d = 1; d < 3; d++) if (nd.bmax[d] - nd.bmin[d] > span) { span = nd.bmax[d] - nd.bmin[d]; dim = d; }
      int32_t mid = lo + (hi - lo) / 2;
      std::nth_element(vind.begin() + lo, vind.begin() + mid, vind.begin() + hi,
                       [&](int32_t a, int32_t b) {
                         float xa = P.at(a)[dim], xb = P.at(b)[dim];
                         return xa < xb || (xa == xb && a < b);
                       });
      int32_t l = build_rec(P, lo, mid);
      int32_t r = build_rec(P, mid, hi);
      nd.left = l; nd.right = r;
    }
    nodes[id] = nd;
    return id;
  }

  static inline double box_lb(const Node& nd, const float* q) {
    double s = 0;
    for (int d = 0; d < 3; d++) {
      double g = 0;
      if (q[d] < nd.bmin[d]) g = (double)nd.bmin[d] - q[d];
      else if (q[d] > nd.bmax[d]) g = (double)q[d] - nd.bmax[d];
      s += g * g;
    }
    // float d2 of any point in the box is >= s*(1-3*2^-24); shrink so pruning can never drop a
    // candidate that ties or beats the current worst under the (d2, idx) order.
    return s * (1.0 - 1e-6);
  }

  // k smallest keys, ascending, into out (size k_eff = min(k, size)). Returns k_eff.
  int knn(const float* q, int k, Key* out) const {
    int n = (int)vind.size();
    if (k > n) k = n;  // [upstream] pcl::KdTreeFLANN::nearestKSearch: k = min(k, total_nr_points_)
    if (k <= 0) return 0;
    std::vector<Key> heap; heap.reserve(k);
    knn_rec(0, q, k, heap);
    std::sort_heap(heap.begin(), heap.end(), key_less);
    for (int i = 0; i < k; i++) out[i] = heap[i];
    return k;
  }
  void knn_rec(int32_t id, const float* q, int k, std::vector<Key>& heap) const {
    const Node& nd = nodes[id];
    if (nd.left < 0) {
      for (int32_t i = nd.lo; i < nd.hi; i++) {
        Key c{d2_flann(q, &data[3 * (size_t)i]), vind[i]};
        if ((int)heap.size() < k) { heap.push_back(c); std::push_heap(heap.begin(), heap.end(), key_less); }
        else if (key_less(c, heap.front())) {
          std::pop_heap(heap.begin(), heap.end(), key_less); heap.back() = c;
          std::push_heap(heap.begin(), heap.end(), key_less);
        }
      }
      return;
    }
    double ll = box_lb(nodes[nd.left], q), lr = box_lb(nodes[nd.right], q);
    int32_t a = nd.left, b = nd.right;
    if (lr < ll) { std::swap(a, b); std::swap(ll, lr); }
    if ((int)heap.size() < k || ll <= (double)heap.front().d2) knn_rec(a, q, k, heap);
    if ((int)heap.size() < k || lr <= (double)heap.front().d2) knn_rec(b, q, k, heap);
  }

  // [upstream] pcl::KdTreeFLANN::radiusSearch: r2 = (float)(radius*radius); flann
  // RadiusResultSet::addPoint keeps dist < radius (strict); sorted by DistIndex::operator<.
  void radius(const float* q, float r2, std::vector<Key>& out) const {
    out.clear();
    if (!nodes.empty()) radius_rec(0, q, r2, out);
    std::sort(out.begin(), out.end(), key_less);
  }
  void radius_rec(int32_t id, const float* q, float r2, std::vector<Key>& out) const {
    const Node& nd = nodes[id];
    if (box_lb(nd, q) >= (double)r2) return;
    if (nd.left < 0) {
      for (int32_t i = nd.lo; i < nd.hi; i++) {
        float d2 = d2_flann(q, &data[3 * (size_t)i]);
        if (d2 < r2) out.push_back(Key{d2, vind[i]});
      }
      return;
    }
    radius_rec(nd.left, q, r2, out);
    radius_rec(nd.right, q, r2, out);
  }
};

// ---------------------------------------------------------------------------------------------
// PCL 1.10 normal estimation, float32, no FMA    [upstream]
// ---------------------------------------------------------------------------------------------
inline void compute_roots2(float b, float c, float roots[3]) {
  // [upstream] pcl::computeRoots2: Scalar d = Scalar (b * b - 4.0 * c);
  roots[0] = 0.0f;
  float bb = b * b;
  float d = (float)((double)bb - 4.0 * (double)c);
  if (d < 0.0f) d = 0.0f;
  float sd = std::sqrt(d);
  roots[2] = 0.5f * (b + sd);
  roots[1] = 0.5f * (b - sd);
}

inline void compute_roots(const float m[6] /*00 01 02 11 12 22*/, float roots[3]) {
  const float m00 = m[0], m01 = m[1], m02 = m[2], m11 = m[3], m12 = m[4], m22 = m[5];
  float c0 = m00 * m11 * m22 + 2.0f * m01 * m02 * m12 - m00 * m12 * m12 - m11 * m02 * m02 - m22 * m01 * m01;
  float c1 = m00 * m11 - m01 * m01 + m00 * m22 - m02 * m02 + m11 * m22 - m12 * m12;
  float c2 = m00 + m11 + m22;
  if (std::fabs(c0) < std::numeric_limits<float>::epsilon()) {
    compute_roots2(c2, c1, roots);
  } else {
    const float s_inv3 = (float)(1.0 / 3.0);
    const float s_sqrt3 = std::sqrt(3.0f);
    float c2_over_3 = c2 * s_inv3;
    float a_over_3 = (c1 - c2 * c2_over_3) * s_inv3;
    if (a_over_3 > 0.0f) a_over_3 = 0.0f;
    float half_b = 0.5f * (c0 + c2_over_3 * (2.0f * c2_over_3 * c2_over_3 - c1));
    float q = half_b * half_b + a_over_3 * a_over_3 * a_over_3;
    if (q > 0.0f) q = 0.0f;
    float rho = std::sqrt(-a_over_3);
    float theta = std::atan2(std::sqrt(-q), half_b) * s_inv3;
    float cos_theta = std::cos(theta);
    float sin_theta = std::sin(theta);
    roots[0] = c2_over_3 + 2.0f * rho * cos_theta;
    roots[1] = c2_over_3 - rho * (cos_theta + s_sqrt3 * sin_theta);
    roots[2] = c2_over_3 - rho * (cos_theta - s_sqrt3 * sin_theta);
    if (roots[0] >= roots[1]) std::swap(roots[0], roots[1]);
    if (roots[1] >= roots[2]) {
      std::swap(roots[1], roots[2]);
      if (roots[0] >= roots[1]) std::swap(roots[0], roots[1]);
    }
    if (roots[0] <= 0.0f) compute_roots2(c2, c1, roots);
  }
}

inline void cross3(const float a[3], const float b[3], float o[3]) {
  // [upstream] Eigen MatrixBase::cross
  o[0] = a[1] * b[2] - a[2] * b[1];
  o[1] = a[2] * b[0] - a[0] * b[2];
  o[2] = a[0] * b[1] - a[1] * b[0];
}
inline float sqnorm3(const float v[3]) {
  // [upstream] Eigen redux_novec_unroller<.,.,0,3>: x*x + (y*y + z*z)
  return v[0] * v[0] + (v[1] * v[1] + v[2] * v[2]);
}

// cov: 00 01 02 11 12 22.  out: nx ny nz curvature  ([upstream] pcl::solvePlaneParameters + eigen33)
inline void solve_plane(const float cov[6], float out[4]) {
  float scale = 0.0f;
  for (int i = 0; i < 6; i++) scale = std::max(scale, std::fabs(cov[i]));
  if (scale <= std::numeric_limits<float>::min()) scale = 1.0f;
  float s[6];
  for (int i = 0; i < 6; i++) s[i] = cov[i] / scale;
  float roots[3];
  compute_roots(s, roots);
  float eigenvalue = roots[0] * scale;
  float d0 = s[0] - roots[0], d1 = s[3] - roots[0], d2 = s[5] - roots[0];
  float r0[3] = {d0, s[1], s[2]}, r1[3] = {s[1], d1, s[4]}, r2[3] = {s[2], s[4], d2};
  float v1[3], v2[3], v3[3];
  cross3(r0, r1, v1); cross3(r0, r2, v2); cross3(r1, r2, v3);
  float l1 = sqnorm3(v1), l2 = sqnorm3(v2), l3 = sqnorm3(v3);
  const float* v; float l;
  if (l1 >= l2 && l1 >= l3) { v = v1; l = l1; }
  else if (l2 >= l1 && l2 >= l3) { v = v2; l = l2; }
  else { v = v3; l = l3; }
  float sl = std::sqrt(l);
  out[0] = v[0] / sl; out[1] = v[1] / sl; out[2] = v[2] / sl;
  float eig_sum = cov[0] + cov[3] + cov[5];
  out[3] = (eig_sum != 0.0f) ? std::fabs(eigenvalue / eig_sum) : 0.0f;
}

// cov_variant 0: PCL 1.10.0 single-pass E[xx]-E[x]E[x]; 1: later PCL (first neighbour subtracted).
inline void normal_from_neighbors(const Pts& P, const Key* nb, int m, const float* q, const float vp[3],
                                  int cov_variant, float out[4]) {
  const float nan = std::numeric_limits<float>::quiet_NaN();
  if (m < 3) { out[0] = out[1] = out[2] = out[3] = nan; return; }  // [upstream] computePointNormal
  float acc[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
  float K[3] = {0, 0, 0};
  if (cov_variant == 1) { const float* f = P.at(nb[0].idx); K[0] = f[0]; K[1] = f[1]; K[2] = f[2]; }
  for (int j = 0; j < m; j++) {
    const float* a = P.at(nb[j].idx);
    float x = a[0], y = a[1], z = a[2];
    if (cov_variant == 1) { x = x - K[0]; y = y - K[1]; z = z - K[2]; }
    acc[0] += x * x; acc[1] += x * y; acc[2] += x * z;
    acc[3] += y * y; acc[4] += y * z; acc[5] += z * z;
    acc[6] += x; acc[7] += y; acc[8] += z;
  }
  float cnt = (float)m;
  for (int i = 0; i < 9; i++) acc[i] = acc[i] / cnt;
  float cov[6];
  cov[0] = acc[0] - acc[6] * acc[6];
  cov[1] = acc[1] - acc[6] * acc[7];
  cov[2] = acc[2] - acc[6] * acc[8];
  cov[3] = acc[3] - acc[7] * acc[7];
  cov[4] = acc[4] - acc[7] * acc[8];
  cov[5] = acc[5] - acc[8] * acc[8];
  solve_plane(cov, out);
  // [upstream] pcl::flipNormalTowardsViewpoint
  float vx = vp[0] - q[0], vy = vp[1] - q[1], vz = vp[2] - q[2];
  float c = vx * out[0] + vy * out[1] + vz * out[2];
  if (c < 0.0f) { out[0] *= -1.0f; out[1] *= -1.0f; out[2] *= -1.0f; }
}

struct Cloud {
  Pts P;
  KdTree tree;  // the persistent "kdtree" member (Set_kdtree / kdtree.setInputCloud)
};

inline int clamp_threads(int t) {
#ifdef _OPENMP
  if (t <= 0) t = omp_get_max_threads();
  return t;
#else
  (void)t; return 1;
#endif
}

// PassThrough("x") restatement. [upstream] pcl::PassThrough<PointT>::applyFilterIndices
inline void passthrough_x(const Pts& P, float lo, float hi, std::vector<int32_t>& out) {
  out.clear();
  for (int64_t i = 0; i < P.n; i++) {
    const float* q = P.at(i);
    if (!finite3(q)) continue;
    if (q[0] < lo || q[0] > hi) continue;
    out.push_back((int32_t)i);
  }
}

typedef std::map<double, std::array<double, 2>> NodeMap;

inline void classify(const Pts& P, const int32_t* ind, int64_t m, float px, std::vector<int32_t>& El,
                     std::vector<int32_t>& Er) {
  // (point - PlanePoint).dot((1,0,0)) = dx*1 + (dy*0 + dz*0); for finite points its sign is sign(x - px)
  El.clear(); Er.clear();
  for (int64_t t = 0; t < m; t++) {
    int32_t i = ind[t];
    const float* q = P.at(i);
    float dx = q[0] - px, dy = q[1] - 0.0f, dz = q[2] - 0.0f;
    float dist = dx * 1.0f + (dy * 0.0f + dz * 0.0f);
    if (dist > 0) El.push_back(i);
    else if (dist < 0) Er.push_back(i);
  }
}

inline void interpolate_nodes(const Pts& P, const std::vector<int32_t>& left_pair,
                              const std::vector<int32_t>& right_pair, float px, NodeMap& node) {
  // src/Path_Generation.cpp:189-202 / src/contour_alg.cpp:220-234
  for (size_t i = 0; i < left_pair.size(); i++) {
    const float* r = P.at(right_pair[i]);
    const float* l = P.at(left_pair[i]);
    float t = (px - r[0]) / (l[0] - r[0]);
    float x = px;
    float y = r[1] + t * (l[1] - r[1]);
    float z = r[2] + t * (l[2] - r[2]);
    node[(double)y] = {(double)x, (double)z};
  }
}

// Variant A: gen-2 brute force + greedy flags. src/Path_Generation.cpp:129-179
inline void pair_variant_a(const Pts& P, const std::vector<int32_t>& El, const std::vector<int32_t>& Er,
                           std::vector<int32_t>& left_pair, std::vector<int32_t>& right_pair) {
  left_pair.clear(); right_pair.clear();
  if (El.empty() || Er.empty()) return;  // reference dereferences an empty map here (UB); define: no pairs
  std::vector<char> Elf(El.size(), 0), Erf(Er.size(), 0);
  auto norm_eigen = [&](const float* a, const float* b) {
    float v[3] = {a[0] - b[0], a[1] - b[1], a[2] - b[2]};
    return std::sqrt(sqnorm3(v));
  };
  // std::map<float,int> compare; compare[norm] = j  => equal keys keep the last j; begin() = min key
  auto argmin_last = [&](const float* a, const std::vector<int32_t>& set) {
    int best = -1; float bk = 0;
    for (int j = 0; j < (int)set.size(); j++) {
      float k = norm_eigen(a, P.at(set[j]));
      if (best < 0 || k <= bk) { best = j; bk = k; }
    }
    return best;
  };
  for (int i = 0; i < (int)El.size(); i++) {
    if (Elf[i]) continue;
    int j = argmin_last(P.at(El[i]), Er);
    if (Erf[j]) continue;
    int32_t rp = Er[j];
    right_pair.push_back(rp);
    Erf[j] = 1;
    int i2 = argmin_last(P.at(rp), El);
    if (!Elf[i2]) { left_pair.push_back(El[i2]); Elf[i2] = 1; }
  }
}

// Variant B: SectPath kd-tree pairing. src/contour_alg.cpp:185-211
inline void pair_variant_b(const Cloud& C, const std::vector<int32_t>& El, const std::vector<int32_t>& Er,
                           std::vector<int32_t>& left_pair, std::vector<int32_t>& right_pair) {
  left_pair.clear(); right_pair.clear();
  if (El.empty() || Er.empty()) return;  // nearestKSearch on an empty tree: define as no pairs
  const Pts& P = C.P;
  KdTree tl, tr;
  tl.build(P, El.data(), (int64_t)El.size());
  tr.build(P, Er.data(), (int64_t)Er.size());
  Key k;
  for (size_t i = 0; i < El.size(); i++) {
    const float* pl = P.at(El[i]);
    tr.knn(pl, 1, &k);                 // treeEr.nearestKSearch(pl, 1, idr, dis)
    const float* pr = P.at(k.idx);
    C.tree.knn(pr, 1, &k);             // kdtree.nearestKSearch(pr, 1, idsave, dis)
    right_pair.push_back(k.idx);
    tl.knn(pr, 1, &k);                 // treeEl.nearestKSearch(pr, 1, idl, dis)
    const float* pl2 = P.at(k.idx);
    C.tree.knn(pl2, 1, &k);            // kdtree.nearestKSearch(pl, 1, idsave, dis)
    left_pair.push_back(k.idx);
  }
}

inline void band_limits(float plane_x, float half_width, int truncate_center, float& lo, float& hi) {
  if (truncate_center) {
    // rangedX_index(int position): the float plane x is truncated by the call; limits are
    // (float)(-2 + position), (float)(2 + position) (int arithmetic first).
    int position = (int)plane_x;
    int hw = (int)half_width;
    lo = (float)(-hw + position);
    hi = (float)(hw + position);
  } else {
    lo = plane_x - half_width;
    hi = plane_x + half_width;
  }
}

}  // namespace

extern "C" {

struct ppo_cloud;  // opaque = Cloud

int ppo_num_threads() { return clamp_threads(0); }

void* ppo_cloud_create(const float* pts, int64_t n, int64_t stride_floats) {
  Cloud* c = new Cloud();
  c->P = Pts{pts, n, stride_floats};
  c->tree.build(c->P, nullptr, 0);
  return c;
}
void ppo_cloud_destroy(void* h) { delete (Cloud*)h; }

// getMinMax3D [upstream]: per-axis min/max over finite points.
void ppo_minmax(const float* pts, int64_t n, int64_t sf, float mn[3], float mx[3]) {
  Pts P{pts, n, sf};
  for (int d = 0; d < 3; d++) { mn[d] = std::numeric_limits<float>::max(); mx[d] = -std::numeric_limits<float>::max(); }
  for (int64_t i = 0; i < n; i++) {
    const float* q = P.at(i);
    if (!finite3(q)) continue;
    for (int d = 0; d < 3; d++) { mn[d] = std::min(mn[d], q[d]); mx[d] = std::max(mx[d], q[d]); }
  }
}

// kNN for nq queries (q == NULL: the cloud's own points). idx_out nq*k (-1 padded), d2_out nullable.
int ppo_knn(void* h, const float* q, int64_t nq, int64_t qsf, int k, int32_t* idx_out, float* d2_out, int threads) {
  Cloud* C = (Cloud*)h;
  Pts Q = q ? Pts{q, nq, qsf} : C->P;
  threads = clamp_threads(threads);
#pragma omp parallel num_threads(threads)
  {
    std::vector<Key> buf(std::max(k, 1));
#pragma omp for schedule(dynamic, 1024)
    for (int64_t i = 0; i < Q.n; i++) {
      int m = 0;
      if (finite3(Q.at(i))) m = C->tree.knn(Q.at(i), k, buf.data());
      for (int j = 0; j < k; j++) {
        idx_out[i * k + j] = j < m ? buf[j].idx : -1;
        if (d2_out) d2_out[i * k + j] = j < m ? buf[j].d2 : std::numeric_limits<float>::infinity();
      }
    }
  }
  return 0;
}

// Brute-force O(N*nq) kNN: verifier for the kd-tree path.
int ppo_knn_brute(const float* pts, int64_t n, int64_t sf, const float* q, int64_t nq, int64_t qsf, int k,
                  int32_t* idx_out, float* d2_out) {
  Pts P{pts, n, sf};
  Pts Q = q ? Pts{q, nq, qsf} : P;
  std::vector<Key> all;
  for (int64_t i = 0; i < Q.n; i++) {
    all.clear();
    if (finite3(Q.at(i)))
      for (int64_t j = 0; j < n; j++)
        if (finite3(P.at(j))) all.push_back(Key{d2_flann(Q.at(i), P.at(j)), (int32_t)j});
    int m = (int)std::min<int64_t>(k, (int64_t)all.size());
    std::partial_sort(all.begin(), all.begin() + m, all.end(), key_less);
    for (int j = 0; j < k; j++) {
      idx_out[i * k + j] = j < m ? all[j].idx : -1;
      if (d2_out) d2_out[i * k + j] = j < m ? all[j].d2 : std::numeric_limits<float>::infinity();
    }
  }
  return 0;
}

// Radius search, two-call sizing: counts always written; if idx_out != NULL, offsets (nq+1) must
// hold the exclusive scan of counts and lists are written sorted by (d2, idx).
int ppo_radius(void* h, const float* q, int64_t nq, int64_t qsf, double radius, int32_t* counts,
               const int64_t* offsets, int32_t* idx_out, float* d2_out, int threads) {
  Cloud* C = (Cloud*)h;
  Pts Q = q ? Pts{q, nq, qsf} : C->P;
  float r2 = (float)(radius * radius);
  threads = clamp_threads(threads);
#pragma omp parallel num_threads(threads)
  {
    std::vector<Key> buf;
#pragma omp for schedule(dynamic, 1024)
    for (int64_t i = 0; i < Q.n; i++) {
      buf.clear();
      if (finite3(Q.at(i))) C->tree.radius(Q.at(i), r2, buf);
      counts[i] = (int32_t)buf.size();
      if (idx_out) {
        int64_t o = offsets[i];
        for (size_t j = 0; j < buf.size(); j++) {
          idx_out[o + j] = buf[j].idx;
          if (d2_out) d2_out[o + j] = buf[j].d2;
        }
      }
    }
  }
  return 0;
}

// NormalEstimation::compute restatement. out: n x 4 floats (nx, ny, nz, curvature).
// mode 0: radius search (radius), mode 1: k search (k).
int ppo_normals(void* h, int mode, double radius, int k, const float vp[3], int cov_variant, float* out,
                int32_t* nn_count_out, int threads) {
  Cloud* C = (Cloud*)h;
  const Pts& P = C->P;
  float r2 = (float)(radius * radius);
  threads = clamp_threads(threads);
  const float nan = std::numeric_limits<float>::quiet_NaN();
#pragma omp parallel num_threads(threads)
  {
    std::vector<Key> buf;
    if (mode == 1) buf.resize(std::max(k, 1));
#pragma omp for schedule(dynamic, 1024)
    for (int64_t i = 0; i < P.n; i++) {
      float* o = out + 4 * i;
      const float* q = P.at(i);
      if (!finite3(q)) { o[0] = o[1] = o[2] = o[3] = nan; if (nn_count_out) nn_count_out[i] = 0; continue; }
      int m;
      if (mode == 0) { C->tree.radius(q, r2, buf); m = (int)buf.size(); }
      else m = C->tree.knn(q, k, buf.data());
      if (nn_count_out) nn_count_out[i] = m;
      normal_from_neighbors(P, buf.data(), m, q, vp, cov_variant, o);
    }
  }
  return 0;
}

// Normal from an explicit neighbour list (unit-test hook for the covariance/eigen restatement).
void ppo_normal_from_list(const float* pts, int64_t n, int64_t sf, const int32_t* nb, int m, const float q[3],
                          const float vp[3], int cov_variant, float out[4]) {
  Pts P{pts, n, sf};
  std::vector<Key> keys(m);
  for (int j = 0; j < m; j++) keys[j] = Key{0.f, nb[j]};
  normal_from_neighbors(P, keys.data(), m, q, vp, cov_variant, out);
}

// rangedX_index for one plane. Returns count; idx_out may be NULL (sizing call).
int64_t ppo_band(const float* pts, int64_t n, int64_t sf, float plane_x, float half_width, int truncate_center,
                 int32_t* idx_out) {
  Pts P{pts, n, sf};
  float lo, hi; band_limits(plane_x, half_width, truncate_center, lo, hi);
  std::vector<int32_t> v; passthrough_x(P, lo, hi, v);
  if (idx_out) std::memcpy(idx_out, v.data(), v.size() * sizeof(int32_t));
  return (int64_t)v.size();
}

// All S slices, reference-style (one full PassThrough scan per slice: O(N*S)). offsets: S+1.
int ppo_slice_bands(const float* pts, int64_t n, int64_t sf, const float* plane_x, int S, float half_width,
                    int truncate_center, int64_t* offsets, int32_t* idx_out, int threads) {
  Pts P{pts, n, sf};
  threads = clamp_threads(threads);
  std::vector<std::vector<int32_t>> bands(S);
#pragma omp parallel for schedule(dynamic, 1) num_threads(threads)
  for (int s = 0; s < S; s++) {
    float lo, hi; band_limits(plane_x[s], half_width, truncate_center, lo, hi);
    passthrough_x(P, lo, hi, bands[s]);
  }
  offsets[0] = 0;
  for (int s = 0; s < S; s++) offsets[s + 1] = offsets[s] + (int64_t)bands[s].size();
  if (idx_out)
    for (int s = 0; s < S; s++) std::memcpy(idx_out + offsets[s], bands[s].data(), bands[s].size() * sizeof(int32_t));
  return 0;
}

// insert_point + path_track/OnePath flattening for one slice given its band indices.
// mode 0 = variant A (gen-2), 1 = variant B (SectPath). Outputs up to cap nodes (ascending y);
// returns node count (may exceed cap: caller re-calls). pair outputs nullable (cap_pairs each).
int64_t ppo_insert_point(void* h, const int32_t* indices, int64_t m, float plane_x, int mode, double* y, double* x,
                         double* z, int64_t cap, int32_t* left_pair_out, int32_t* right_pair_out,
                         int64_t* n_left, int64_t* n_right) {
  Cloud* C = (Cloud*)h;
  std::vector<int32_t> El, Er, lp, rp;
  classify(C->P, indices, m, plane_x, El, Er);
  if (mode == 0) pair_variant_a(C->P, El, Er, lp, rp);
  else pair_variant_b(*C, El, Er, lp, rp);
  if (n_left) *n_left = (int64_t)lp.size();
  if (n_right) *n_right = (int64_t)rp.size();
  if (left_pair_out) std::memcpy(left_pair_out, lp.data(), lp.size() * sizeof(int32_t));
  if (right_pair_out) std::memcpy(right_pair_out, rp.data(), rp.size() * sizeof(int32_t));
  NodeMap node;
  interpolate_nodes(C->P, lp, rp, plane_x, node);
  int64_t i = 0;
  for (auto& kv : node) {
    if (i < cap) { y[i] = kv.first; x[i] = kv.second[0]; z[i] = kv.second[1]; }
    i++;
  }
  return (int64_t)node.size();
}

// Whole sweep: for each plane, rangedX_index -> insert_point -> flatten. node_offsets: S+1.
// y/x/z hold up to cap nodes in total; returns total node count (re-call with a larger cap if bigger).
int64_t ppo_slice_contours(void* h, const float* plane_x, int S, float half_width, int truncate_center, int mode,
                           int64_t* node_offsets, double* y, double* x, double* z, int64_t cap, int threads) {
  Cloud* C = (Cloud*)h;
  threads = clamp_threads(threads);
  std::vector<std::vector<double>> ys(S), xs(S), zs(S);
#pragma omp parallel for schedule(dynamic, 1) num_threads(threads)
  for (int s = 0; s < S; s++) {
    float lo, hi; band_limits(plane_x[s], half_width, truncate_center, lo, hi);
    std::vector<int32_t> band, El, Er, lp, rp;
    passthrough_x(C->P, lo, hi, band);
    classify(C->P, band.data(), (int64_t)band.size(), plane_x[s], El, Er);
    if (mode == 0) pair_variant_a(C->P, El, Er, lp, rp);
    else pair_variant_b(*C, El, Er, lp, rp);
    NodeMap node;
    interpolate_nodes(C->P, lp, rp, plane_x[s], node);
    for (auto& kv : node) { ys[s].push_back(kv.first); xs[s].push_back(kv.second[0]); zs[s].push_back(kv.second[1]); }
  }
  node_offsets[0] = 0;
  for (int s = 0; s < S; s++) node_offsets[s + 1] = node_offsets[s] + (int64_t)ys[s].size();
  int64_t total = node_offsets[S];
  if (total <= cap && y)
    for (int s = 0; s < S; s++) {
      std::memcpy(y + node_offsets[s], ys[s].data(), ys[s].size() * sizeof(double));
      std::memcpy(x + node_offsets[s], xs[s].data(), xs[s].size() * sizeof(double));
      std::memcpy(z + node_offsets[s], zs[s].data(), zs[s].size() * sizeof(double));
    }
  return total;
}

// Plane generators. variant: 0 gen-2 Contact_Path_Generation
// (src/Path_Generation.cpp:711-723), 1 gen-2 slicing_method (:295-303), 2 gen-1 free
// slicing_method step 40 (src/slicing_method.cpp:617-626), 3 SectPath::GenPath centre-out
// (src/contour_alg.cpp:305-327; returned in Path_set order: front sweep reversed, then back sweep).
int ppo_planes(int variant, float min_x, float max_x, double tool_radius, float* out, int cap) {
  std::vector<float> v;
  int step = (int)(tool_radius * 2);
  if (variant == 0) {
    float loc = min_x + tool_radius;  // float + double -> double -> float, as the reference's initialiser
    while (loc < max_x) { v.push_back(loc); loc += step; if (step <= 0) break; }
  } else if (variant == 1) {
    float m = min_x; m += step / 2;
    while (m < max_x) { v.push_back(m); m += step; if (step <= 0) break; }
  } else if (variant == 2) {
    float m = min_x; m += 40;
    while (m < max_x) { v.push_back(m); m += 40; }
  } else if (variant == 4) {
    // gen-3 two-thread sweep (src/Path_Alg/path_dynamic_alg.cpp:308-366): centre path at the float centre, the two
    // workers count in INT from the int-truncated bounding box (include/Path_Generate_Algorithm.h:110-117)
    const int mn = (int)min_x, mx = (int)max_x;
    std::vector<float> front, back;
    int loc = (mx + mn) / 2 - step;
    while (mx > loc && loc > mn) { front.insert(front.begin(), (float)loc); loc -= step; if (step <= 0) break; }
    loc = (mx + mn) / 2 + step;
    while (mx > loc && loc > mn) { back.push_back((float)loc); loc += step; if (step <= 0) break; }
    v = front; v.push_back((min_x + max_x) / 2); v.insert(v.end(), back.begin(), back.end());
  } else if (variant == 5) {
    // gen-3 single-direction sweep (src/Path_Alg/dynamic_alg_sdir.cpp:349-374): int loc = min_pt.x + toolRadius
    int loc = (int)(min_x + tool_radius);
    v.push_back((float)loc);
    loc += step;
    while (loc < max_x) { v.push_back((float)loc); loc += step; if (step <= 0) break; }
  } else {
    std::vector<float> front, back;
    float loc = (min_x + max_x) / 2 - step;
    while (loc > min_x) { front.insert(front.begin(), loc); loc -= step; if (step <= 0) break; }
    loc = (min_x + max_x) / 2;
    while (loc < max_x) { back.push_back(loc); loc += step; if (step <= 0) break; }
    v = front; v.insert(v.end(), back.begin(), back.end());
  }
  int n = (int)v.size();
  for (int i = 0; i < n && i < cap; i++) out[i] = v[i];
  return n;
}

// ---------------------------------------------------------------------------------------------
// "next" rows — restated for the host-side consumers of the hot path's output.
// ---------------------------------------------------------------------------------------------

// gsl_interp_steffen (GSL >= 2.0, interpolation/steffen.c) [upstream, recalled; GSL is not in the
// image]: monotone cubic Hermite spline of Steffen (1990).  include/Spline.h:10-29 builds two of
// them, x(y) and z(y), over the ordered contour nodes.  Evaluates nq abscissae; returns -1 if
// n < 3 or the abscissae are not strictly increasing (GSL's error handler would abort), and
// writes NaN for queries outside [xa[0], xa[n-1]] (gsl_spline_eval domain error).
int ppo_steffen_eval(const double* xa, const double* ya, int64_t n, const double* xq, int64_t nq, double* out) {
  if (n < 3) return -1;
  for (int64_t i = 1; i < n; i++) if (!(xa[i] > xa[i - 1])) return -1;
  std::vector<double> yp(n), a(n - 1), b(n - 1), c(n - 1), d(n - 1);
  auto cps = [](double v) { return std::copysign(1.0, v); };
  double h0 = xa[1] - xa[0];
  yp[0] = (ya[1] - ya[0]) / h0;
  for (int64_t i = 1; i < n - 1; i++) {
    double hi = xa[i + 1] - xa[i], him1 = xa[i] - xa[i - 1];
    double si = (ya[i + 1] - ya[i]) / hi, sim1 = (ya[i] - ya[i - 1]) / him1;
    double pi = (sim1 * hi + si * him1) / (him1 + hi);
    yp[i] = (cps(sim1) + cps(si)) * std::min(std::fabs(sim1), std::min(std::fabs(si), 0.5 * std::fabs(pi)));
  }
  yp[n - 1] = (ya[n - 1] - ya[n - 2]) / (xa[n - 1] - xa[n - 2]);
  for (int64_t i = 0; i < n - 1; i++) {
    double hi = xa[i + 1] - xa[i], si = (ya[i + 1] - ya[i]) / hi;
    a[i] = (yp[i] + yp[i + 1] - 2 * si) / hi / hi;
    b[i] = (3 * si - 2 * yp[i] - yp[i + 1]) / hi;
    c[i] = yp[i];
    d[i] = ya[i];
  }
  for (int64_t q = 0; q < nq; q++) {
    double x = xq[q];
    if (!(x >= xa[0] && x <= xa[n - 1])) { out[q] = std::numeric_limits<double>::quiet_NaN(); continue; }
    int64_t i = std::upper_bound(xa, xa + n, x) - xa - 1;  // gsl_interp_bsearch: xa[i] <= x < xa[i+1]
    if (i > n - 2) i = n - 2;
    double dx = x - xa[i];
    out[q] = d[i] + dx * (c[i] + dx * (b[i] + dx * a[i]));
  }
  return 0;
}

// compute_coverage (src/Path_Generation.cpp:483-496): kdtree.radiusSearch(node, radius) and
// coverage_flag[i] = 1 for every neighbour; flags is in/out (N bytes).
int ppo_coverage_mark(void* h, const float* q, int64_t nq, int64_t qsf, double radius, unsigned char* flags) {
  Cloud* C = (Cloud*)h;
  float r2 = (float)(radius * radius);
  std::vector<Key> buf;
  for (int64_t i = 0; i < nq; i++) {
    const float* p = q + i * qsf;
    if (!finite3(p)) continue;
    C->tree.radius(p, r2, buf);
    for (auto& k : buf) flags[k.idx] = 1;
  }
  return 0;
}

// pcl::PrincipalCurvaturesEstimation::computePointPrincipalCurvatures [upstream, recalled, PCL 1.10
// features/impl/principal_curvatures.hpp] as called by compute_transform
// (src/Path_Generation.cpp:362-400 with k = 10, src/Path_Alg/path_dynamic_alg.cpp:77-110 with k = 50):
// kNN of the query, normals of the neighbours projected into the tangent plane of neighbour [0],
// 3x3 covariance of the projections, eigen33 (values) + computeCorrespondingEigenVector for the
// largest eigenvalue.  out per query: pcx, pcy, pcz, pc1, pc2; nn0 = nearest point index.
// normals: n x nsf floats (nx, ny, nz at 0..2).
namespace {
inline void eigen33_values(const float m[9], float evals[3]) {
  float scale = 0.0f;
  for (int i = 0; i < 9; i++) scale = std::max(scale, std::fabs(m[i]));
  if (scale <= std::numeric_limits<float>::min()) scale = 1.0f;
  float s[6] = {m[0] / scale, m[1] / scale, m[2] / scale, m[4] / scale, m[5] / scale, m[8] / scale};
  compute_roots(s, evals);
  for (int i = 0; i < 3; i++) evals[i] *= scale;
}
inline void corresponding_eigenvector(const float m[9], float eigenvalue, float v[3]) {
  float scale = 0.0f;
  for (int i = 0; i < 9; i++) scale = std::max(scale, std::fabs(m[i]));
  if (scale <= std::numeric_limits<float>::min()) scale = 1.0f;
  float a[9];
  for (int i = 0; i < 9; i++) a[i] = m[i] / scale;
  float ev = eigenvalue / scale;
  a[0] -= ev; a[4] -= ev; a[8] -= ev;
  float v1[3], v2[3], v3[3];
  cross3(&a[0], &a[3], v1); cross3(&a[0], &a[6], v2); cross3(&a[3], &a[6], v3);
  float l1 = sqnorm3(v1), l2 = sqnorm3(v2), l3 = sqnorm3(v3);
  const float* w; float l;
  if (l1 >= l2 && l1 >= l3) { w = v1; l = l1; }
  else if (l2 >= l1 && l2 >= l3) { w = v2; l = l2; }
  else { w = v3; l = l3; }
  float sl = std::sqrt(l);
  v[0] = w[0] / sl; v[1] = w[1] / sl; v[2] = w[2] / sl;
}
}  // namespace

int ppo_principal_curvatures(void* h, const float* normals, int64_t nsf, const float* q, int64_t nq, int64_t qsf, int k,
                             float* out /* nq x 5 */, int32_t* nn0) {
  Cloud* C = (Cloud*)h;
  std::vector<Key> nb(std::max(k, 1));
  const float nan = std::numeric_limits<float>::quiet_NaN();
  for (int64_t t = 0; t < nq; t++) {
    const float* p = q + t * qsf;
    float* o = out + 5 * t;
    int m = finite3(p) ? C->tree.knn(p, k, nb.data()) : 0;
    if (m == 0) { for (int i = 0; i < 5; i++) o[i] = nan; if (nn0) nn0[t] = -1; continue; }
    if (nn0) nn0[t] = nb[0].idx;
    const float* n0 = normals + (int64_t)nb[0].idx * nsf;
    // M = I - n n^T  (row-major 3x3)
    float M[9];
    for (int i = 0; i < 3; i++)
      for (int j = 0; j < 3; j++) M[3 * i + j] = (i == j ? 1.0f : 0.0f) - n0[i] * n0[j];
    std::vector<std::array<float, 3>> proj(m);
    float cen[3] = {0, 0, 0};
    for (int j = 0; j < m; j++) {
      const float* nj = normals + (int64_t)nb[j].idx * nsf;
      for (int i = 0; i < 3; i++) proj[j][i] = M[3 * i] * nj[0] + (M[3 * i + 1] * nj[1] + M[3 * i + 2] * nj[2]);
      for (int i = 0; i < 3; i++) cen[i] += proj[j][i];
    }
    for (int i = 0; i < 3; i++) cen[i] /= (float)m;
    float cov[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    for (int j = 0; j < m; j++) {
      float d[3] = {proj[j][0] - cen[0], proj[j][1] - cen[1], proj[j][2] - cen[2]};
      float xy = d[0] * d[1], xz = d[0] * d[2], yz = d[1] * d[2];
      cov[0] += d[0] * d[0]; cov[1] += xy; cov[2] += xz;
      cov[3] += xy; cov[4] += d[1] * d[1]; cov[5] += yz;
      cov[6] += xz; cov[7] += yz; cov[8] += d[2] * d[2];
    }
    float ev[3], vec[3];
    eigen33_values(cov, ev);
    corresponding_eigenvector(cov, ev[2], vec);
    float inv = 1.0f / (float)m;
    o[0] = vec[0]; o[1] = vec[1]; o[2] = vec[2];
    o[3] = ev[2] * inv;
    o[4] = ev[1] * inv;
  }
  return 0;
}

// pcl::StatisticalOutlierRemoval<PointT>::applyFilterIndices [upstream, recalled, PCL 1.10
// filters/impl/statistical_outlier_removal.hpp] as run by SectPath::remove_outlier
// (src/contour_alg.cpp:101-108: setMeanK(50), setStddevMulThresh(1.0); twins
// src/Path_Alg/path_slicing_alg.cpp:101-108, commented out at src/Path_Generation.cpp:17-22).
// First pass: per point, nearestKSearch(point, mean_k + 1); dist_sum (double) += sqrt(nn_dists[j]) for
// j = 1..mean_k (entry 0 is the query itself); distances[i] = (float)(dist_sum / mean_k).  Non-finite
// points get 0 and are not counted in n_valid.  `sqrt (nn_dists[k])` is an unqualified call on a float:
// whether it resolves to the float overload depends on the headers of the translation unit, so both
// are provided (sqrt_float = 0: double sqrt of the widened value; 1: float sqrt).
int ppo_sor_mean_distances(void* h, int mean_k, int sqrt_float, float* dist, int64_t* n_valid, int threads) {
  Cloud* C = (Cloud*)h;
  const int k = mean_k + 1;
  if (mean_k < 1 || C->P.n < k) return -1;   // the reference reads past the result list here
  threads = clamp_threads(threads);
  int64_t valid = 0;
#pragma omp parallel num_threads(threads) reduction(+ : valid)
  {
    std::vector<Key> buf(k);
#pragma omp for schedule(dynamic, 1024)
    for (int64_t i = 0; i < C->P.n; i++) {
      dist[i] = 0.0f;
      if (!finite3(C->P.at(i))) continue;
      int m = C->tree.knn(C->P.at(i), k, buf.data());
      if (m < k) continue;                   // fewer finite points than k: treated as a failed search
      double dist_sum = 0.0;
      for (int j = 1; j < k; j++) dist_sum += sqrt_float ? (double)std::sqrt(buf[j].d2) : std::sqrt((double)buf[j].d2);
      dist[i] = (float)(dist_sum / mean_k);
      valid++;
    }
  }
  *n_valid = valid;
  return 0;
}

// Second half of applyFilterIndices: sequential double sums over ALL distances (zeros of the invalid
// points included), mean / variance over n_valid, threshold = mean + std_mul * stddev, inliers are
// distances[i] <= threshold (negative = 0).  kept: ascending indices; returns their count.
int64_t ppo_sor_select(const float* dist, int64_t n, int64_t n_valid, double std_mul, int negative, int32_t* kept,
                       double* threshold_out) {
  double sum = 0, sq_sum = 0;
  for (int64_t i = 0; i < n; i++) {
    sum += dist[i];
    sq_sum += dist[i] * dist[i];             // float product, widened on accumulation
  }
  double mean = sum / (double)n_valid;
  double variance = (sq_sum - sum * sum / (double)n_valid) / ((double)n_valid - 1);
  double stddev = std::sqrt(variance);
  double thr = mean + std_mul * stddev;
  if (threshold_out) *threshold_out = thr;
  int64_t o = 0;
  for (int64_t i = 0; i < n; i++) {
    bool outlier = (!negative && dist[i] > thr) || (negative && dist[i] <= thr);
    if (!outlier) kept[o++] = (int32_t)i;
  }
  return o;
}

}  // extern "C"
