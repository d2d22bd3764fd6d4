// Cloud ingest (pack + bounding box = getMinMax3D) and the column-grid spatial index that
// replaces pcl::KdTreeFLANN::setInputCloud (src/Path_Generation.cpp:335-338,695;
// src/contour_alg.cpp:292) and NormalEstimation's private search::KdTree (:326-327).
//
// Index = counting sort of the finite points by cell id (v*nu + u) over a dense cell table:
// histogram (atomics) -> exclusive scan -> scatter.  Rows of cells are contiguous in the sorted
// array, so a query's candidates are (2R+1) contiguous ranges that consecutive (sorted) queries
// share: coalesced float4 loads, L1/L2 reuse, no tree walk.
#include <stdlib.h>

#include <algorithm>
#include <cmath>

#include "ppp_device.cuh"

namespace {

__device__ __forceinline__ uint32_t f2ord(float f) {
  uint32_t b = __float_as_uint(f);
  return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
inline float ord2f(uint32_t u) {
  uint32_t b = (u & 0x80000000u) ? (u & 0x7FFFFFFFu) : ~u;
  float f;
  memcpy(&f, &b, 4);
  return f;
}

struct BBoxAcc {
  uint32_t mn[3];
  uint32_t mx[3];
  unsigned long long n_finite;
};

__global__ void k_bbox_init(BBoxAcc* acc) {
  for (int d = 0; d < 3; d++) { acc->mn[d] = 0xFFFFFFFFu; acc->mx[d] = 0u; }
  acc->n_finite = 0ull;
}

// Pack stride-`sf` records to float4 (original order) and reduce the bounding box of the finite
// points.  One pass over the raw cloud: 12 useful bytes of each record in, 16 out.
__global__ void __launch_bounds__(256) k_pack_bbox(const float* __restrict__ raw, int64_t n, int sf, int vec_ok,
                                                   float4* __restrict__ xyz4, BBoxAcc* __restrict__ acc) {
  float mn[3] = {CUDART_INF_F, CUDART_INF_F, CUDART_INF_F};
  float mx[3] = {-CUDART_INF_F, -CUDART_INF_F, -CUDART_INF_F};
  unsigned cnt = 0;
  constexpr int PPT = 4;  // points per thread per sweep: PPT independent loads in flight
  for (int64_t base = (int64_t)blockIdx.x * blockDim.x * PPT; base < n; base += (int64_t)gridDim.x * blockDim.x * PPT) {
    float x[PPT], y[PPT], z[PPT];
#pragma unroll
    for (int j = 0; j < PPT; j++) {
      int64_t i = base + (int64_t)j * blockDim.x + threadIdx.x;
      x[j] = y[j] = z[j] = 0.f;
      if (i < n) {
        if (vec_ok) {
          float4 p = __ldg(reinterpret_cast<const float4*>(raw + i * sf));
          x[j] = p.x; y[j] = p.y; z[j] = p.z;
        } else {
          const float* p = raw + i * sf;
          x[j] = __ldg(p); y[j] = __ldg(p + 1); z[j] = __ldg(p + 2);
        }
      }
    }
#pragma unroll
    for (int j = 0; j < PPT; j++) {
      int64_t i = base + (int64_t)j * blockDim.x + threadIdx.x;
      if (i >= n) continue;
      bool fin = finite3(x[j], y[j], z[j]);
      xyz4[i] = make_float4(x[j], y[j], z[j], fin ? 0.0f : CUDART_NAN_F);
      if (fin) {
        cnt++;
        mn[0] = fminf(mn[0], x[j]); mx[0] = fmaxf(mx[0], x[j]);
        mn[1] = fminf(mn[1], y[j]); mx[1] = fmaxf(mx[1], y[j]);
        mn[2] = fminf(mn[2], z[j]); mx[2] = fmaxf(mx[2], z[j]);
      }
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
    for (int d = 0; d < 3; d++) {
      mn[d] = fminf(mn[d], __shfl_xor_sync(0xffffffffu, mn[d], o));
      mx[d] = fmaxf(mx[d], __shfl_xor_sync(0xffffffffu, mx[d], o));
    }
    cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
  }
  // block-level combine, then ONE set of global atomics per block (per-warp atomics on the same
  // seven addresses serialise in L2 and dominated this kernel)
  __shared__ float s_mn[8][3], s_mx[8][3];
  __shared__ unsigned s_cnt[8];
  const int w = threadIdx.x >> 5;
  if ((threadIdx.x & 31) == 0) {
#pragma unroll
    for (int d = 0; d < 3; d++) { s_mn[w][d] = mn[d]; s_mx[w][d] = mx[d]; }
    s_cnt[w] = cnt;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned tot = 0;
    for (int i = 0; i < (int)(blockDim.x >> 5); i++) {
      tot += s_cnt[i];
#pragma unroll
      for (int d = 0; d < 3; d++) { mn[d] = fminf(mn[d], s_mn[i][d]); mx[d] = fmaxf(mx[d], s_mx[i][d]); }
    }
    if (tot) {
#pragma unroll
      for (int d = 0; d < 3; d++) {
        atomicMin(&acc->mn[d], f2ord(mn[d]));
        atomicMax(&acc->mx[d], f2ord(mx[d]));
      }
      atomicAdd(&acc->n_finite, (unsigned long long)tot);
    }
  }
}

__device__ __forceinline__ int cell_of(const GridView& g, float x, float y, float z) {
  int cu = clampi(cell_coord_raw(axis_of(x, y, z, g.au), g.min_u, g.inv_h), 0, g.nu - 1);
  int cv = clampi(cell_coord_raw(axis_of(x, y, z, g.av), g.min_v, g.inv_h), 0, g.nv - 1);
  return cv * g.nu + cu;
}

__global__ void __launch_bounds__(256) k_cell_count(GridView g, const float4* __restrict__ xyz4, int64_t n,
                                                    int32_t* __restrict__ counts) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float4 p = __ldg(xyz4 + i);
  if (p.w != p.w) return;  // non-finite point: not indexed
  atomicAdd(counts + cell_of(g, p.x, p.y, p.z), 1);
}

// counts[] still holds the histogram; each point claims slot start[c] + (--counts[c]).
__global__ void __launch_bounds__(256) k_cell_scatter(GridView g, const float4* __restrict__ xyz4, int64_t n,
                                                      int32_t* __restrict__ counts, const int32_t* __restrict__ start,
                                                      float4* __restrict__ sorted, int32_t* __restrict__ order) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float4 p = __ldg(xyz4 + i);
  if (p.w != p.w) return;
  int c = cell_of(g, p.x, p.y, p.z);
  int slot = __ldg(start + c) + atomicSub(counts + c, 1) - 1;
  sorted[slot] = make_float4(p.x, p.y, p.z, __int_as_float((int)i));
  order[slot] = (int)i;
}

}  // namespace

int cloud_ingest(ppp_cloud* c, const void* pts_dev, size_t stride_bytes) {
  ppp_ctx* ctx = c->ctx;
  int sf = (int)(stride_bytes / 4);
  int vec_ok = (stride_bytes % 16 == 0) && (((uintptr_t)pts_dev) % 16 == 0);
  PPP_TRY(dev_alloc_keep(ctx, &c->xyz4, (size_t)c->n));
  BBoxAcc* acc = nullptr;
  PPP_TRY(dev_alloc(ctx, &acc, 1));
  PPP_LAUNCH(ctx, "bbox_init", k_bbox_init, 1, 1, 0, acc);
  PPP_CHECK_LAUNCH();
  if (c->n > 0) {
    int blocks = (int)std::max<int64_t>(1, std::min<int64_t>((c->n + 1023) / 1024, (int64_t)ctx->sm_count * 16));
    PPP_LAUNCH(ctx, "pack_bbox", k_pack_bbox, blocks, 256, 0, (const float*)pts_dev, c->n, sf, vec_ok, c->xyz4, acc);
    PPP_CHECK_LAUNCH();
  }
  BBoxAcc h;
  PPP_TRY(fetch_small(ctx, acc, sizeof(h), &h));
  dev_free(ctx, acc);
  c->n_finite = (int64_t)h.n_finite;
  for (int d = 0; d < 3; d++) {
    // [upstream] getMinMax3D starts from +/-FLT_MAX; an all-non-finite cloud keeps those.
    c->bmin[d] = c->n_finite ? ord2f(h.mn[d]) : 3.402823466e+38f;
    c->bmax[d] = c->n_finite ? ord2f(h.mx[d]) : -3.402823466e+38f;
  }
  // (u, v) = the two axes with the largest extent; ties keep x, y.
  double ext[3];
  for (int d = 0; d < 3; d++) ext[d] = c->n_finite ? (double)c->bmax[d] - (double)c->bmin[d] : 0.0;
  int drop = 2;  // the axis with the smallest extent is not gridded (ties: drop z, then y)
  if (ext[1] < ext[drop]) drop = 1;
  if (ext[0] < ext[drop]) drop = 0;
  c->au = drop == 0 ? 1 : 0;
  c->av = drop == 2 ? 1 : 2;
  double area = std::max(ext[c->au], 1e-9) * std::max(ext[c->av], 1e-9);
  c->density = c->n_finite > 0 ? (double)c->n_finite / area : 0.0;
  return PPP_OK;
}

// Rings of cells in the fixed candidate block of the fast k-nearest kernels (2: a 5 x 5 block).
int knn_block_rings() {
  static int r0 = [] {
    int v = 2;
    if (const char* e = getenv("PPP_KNN_R0")) { int t = atoi(e); if (t >= 1 && t <= 4) v = t; }  // tuning aid
    return v;
  }();
  return r0;
}

// Cell size for a k-search: 2 rings of cells should cover the expected k-th neighbour distance
// with ~35% head room (queries that need more simply expand further rings).
double cloud_cell_for_k(const ppp_cloud* c, int k) {
  if (c->cell_hint > 0) return c->cell_hint;
  double rho = c->density > 0 ? c->density : 1.0;
  double rk = std::sqrt((double)std::max(k, 1) / (3.14159265358979 * rho));
  double f = 1.35;
  if (const char* e = getenv("PPP_CELL_FACTOR")) { double v = atof(e); if (v > 0.5 && v < 4.0) f = v; }  // tuning aid
  return f * rk / (double)knn_block_rings();
}
// Cell size for a radius search: R = 2 rings cover r exactly (plus rounding slack).
double cloud_cell_for_radius(const ppp_cloud* c, double r) {
  (void)c;
  return r * 0.5 * (1.0 + 1e-3);
}

int cloud_get_grid(ppp_cloud* c, double h, GridStore** out) {
  ppp_ctx* ctx = c->ctx;
  // cap the dense cell table (nu*nv <= 2^28) by growing h if needed
  double eu = c->n_finite ? (double)c->bmax[c->au] - (double)c->bmin[c->au] : 0.0;
  double ev = c->n_finite ? (double)c->bmax[c->av] - (double)c->bmin[c->av] : 0.0;
  if (!(h > 0)) h = 1.0;
  while ((std::floor(eu / h) + 2) * (std::floor(ev / h) + 2) > 268435456.0) h *= 1.5;
  for (auto& g : c->grids)
    if (std::fabs(g.h - h) <= 1e-9 * h) { *out = &g; return PPP_OK; }
  GridStore gs;
  gs.h = h;
  GridView& v = gs.v;
  v.au = c->au; v.av = c->av;
  v.h = (float)h;
  v.inv_h = 1.0f / v.h;
  v.min_u = c->n_finite ? c->bmin[c->au] : 0.0f;
  v.min_v = c->n_finite ? c->bmin[c->av] : 0.0f;
  v.nu = (int)std::floor(eu * (double)v.inv_h) + 2;
  v.nv = (int)std::floor(ev * (double)v.inv_h) + 2;
  // rounding slack: cell coordinates are exact to a few ulps of (a-min)*inv_h
  v.slack = (float)(h * (double)(v.nu + v.nv) * 9.5367431640625e-07 + 1e-30);
  v.n_sorted = (int)c->n_finite;
  int64_t ncells = (int64_t)v.nu * v.nv;
  PPP_TRY(dev_alloc_keep(ctx, &gs.sorted, (size_t)std::max<int64_t>(c->n_finite, 1)));
  PPP_TRY(dev_alloc_keep(ctx, &gs.order, (size_t)std::max<int64_t>(c->n_finite, 1)));
  PPP_TRY(dev_alloc_keep(ctx, &gs.cell_start, (size_t)ncells + 1));
  int32_t* counts = nullptr;
  PPP_TRY(dev_alloc(ctx, &counts, (size_t)ncells));
  PPP_CUDA(cudaMemsetAsync(counts, 0, (size_t)ncells * sizeof(int32_t), ctx->stream));
  v.sorted = gs.sorted;
  v.cell_start = gs.cell_start;
  if (c->n > 0) {
    unsigned blocks = (unsigned)((c->n + 255) / 256);
    PPP_LAUNCH(ctx, "cell_count", k_cell_count, blocks, 256, 0, v, (const float4*)c->xyz4, c->n, counts);
    PPP_CHECK_LAUNCH();
  }
  PPP_TRY(scan_exclusive_i32(ctx, counts, gs.cell_start, ncells));
  if (c->n > 0) {
    unsigned blocks = (unsigned)((c->n + 255) / 256);
    PPP_LAUNCH(ctx, "cell_scatter", k_cell_scatter, blocks, 256, 0, v, (const float4*)c->xyz4, c->n, counts,
               (const int32_t*)gs.cell_start, gs.sorted, gs.order);
    PPP_CHECK_LAUNCH();
  }
  dev_free(ctx, counts);
  PPP_CUDA(cudaEventCreateWithFlags(&gs.ready, cudaEventDisableTiming));
  PPP_CUDA(cudaEventRecord(gs.ready, ctx->stream));
  c->grids.push_back(gs);
  *out = &c->grids.back();
  return PPP_OK;
}

// ---------------------------------------------------------------------------------------------
__global__ void k_fetch_small(const unsigned* __restrict__ src, unsigned* __restrict__ dst, int words) {
  for (int i = threadIdx.x; i < words; i += blockDim.x) dst[i] = src[i];
}

int fetch_small(ppp_ctx* ctx, const void* dev_src, size_t bytes, void* host_dst) {
  if (bytes == 0) { PPP_CUDA(cudaStreamSynchronize(ctx->stream)); return PPP_OK; }
  if (bytes > FETCH_BYTES || (bytes & 3) || ((uintptr_t)dev_src & 3) || !ctx->fetch_host) {
    PPP_CUDA(cudaMemcpyAsync(host_dst, dev_src, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    PPP_CUDA(cudaStreamSynchronize(ctx->stream));
    return PPP_OK;
  }
  PPP_LAUNCH(ctx, "fetch_small", k_fetch_small, 1, 256, 0, (const unsigned*)dev_src, (unsigned*)ctx->fetch_host, (int)(bytes / 4));
  PPP_CHECK_LAUNCH();
  PPP_CUDA(cudaStreamSynchronize(ctx->stream));
  memcpy(host_dst, ctx->fetch_host, bytes);
  return PPP_OK;
}
