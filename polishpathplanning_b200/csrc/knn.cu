// Exact k-nearest-neighbour / radius search on the column grid, with the PCL normal estimation
// fused behind it.  Replaces pcl::KdTreeFLANN::nearestKSearch / radiusSearch and
// pcl::NormalEstimation::compute (src/Path_Generation.cpp:323-333, src/contour_alg.cpp:142-151).
//
// One thread per query, queries taken in sorted (cell-major) order so a warp's 32 queries share
// the same few cell rows: candidate loads are float4, contiguous per row and hit L1/L2.  The
// per-query result list lives in shared memory (column layout: entry j of thread t at
// [j*blockDim + t], bank = t, conflict-free) as 64-bit keys (d2 bits << 32 | index), so the
// reference's (d2, index) order is a single unsigned compare and the list is already in the
// order the PCL covariance must be accumulated in.
#include <algorithm>
#include <cmath>

#include "ppp_device.cuh"

namespace {

struct SearchParams {
  GridView g;
  const float4* xyz4;   // original order (neighbour gather for the covariance)
  const float* q;       // external queries (nullptr: the cloud's own points in sorted order)
  int q_sf;
  int64_t nq;
  int64_t first;        // self mode: first sorted position
  int cap;              // list capacity (k in k-mode)
  int kk;               // k-mode: min(k, indexed points)
  int mode;             // 0 = k nearest, 1 = radius (d2 < r2, fixed block radius R0)
  int R0;
  float r2;
  int32_t* idx_out;     // k-mode: row*cap ; radius mode: offsets[row]
  float* d2_out;
  const int64_t* offsets;
  float* normals;
  int nsf;
  float vpx, vpy, vpz;
  unsigned flags;
};

__device__ __forceinline__ void list_insert(u64* L, int BD, int& cnt, int cap, u64 key) {
  int j = cnt < cap ? cnt++ : cap - 1;
  while (j > 0) {
    u64 prev = L[(j - 1) * BD];
    if (prev < key) break;
    L[j * BD] = prev;
    j--;
  }
  L[j * BD] = key;
}

__global__ void __launch_bounds__(128) k_search(SearchParams P) {
  extern __shared__ u64 s_keys[];
  const int BD = blockDim.x;
  int64_t t = (int64_t)blockIdx.x * BD + threadIdx.x;
  if (t >= P.nq) return;
  const GridView& g = P.g;
  float qx, qy, qz;
  int64_t row;
  if (P.q) {
    const float* qp = P.q + t * P.q_sf;
    qx = __ldg(qp); qy = __ldg(qp + 1); qz = __ldg(qp + 2);
    row = t;
  } else {
    float4 p = __ldg(g.sorted + P.first + t);
    qx = p.x; qy = p.y; qz = p.z;
    row = __float_as_int(p.w);
  }
  u64* L = s_keys + threadIdx.x;
  int cnt = 0;
  const bool fin = finite3(qx, qy, qz);
  if (fin && g.n_sorted > 0) {
    int cu = cell_coord_raw(axis_of(qx, qy, qz, g.au), g.min_u, g.inv_h);
    int cv = cell_coord_raw(axis_of(qx, qy, qz, g.av), g.min_v, g.inv_h);
    if (P.mode == 1) {
      u64 tau = make_key(P.r2, 0);
      visit_annulus(g, cu, cv, -1, P.R0, [&](float4 c) {
        u64 key = make_key(d2_flann(qx, qy, qz, c.x, c.y, c.z), __float_as_int(c.w));
        if (key < tau) {
          list_insert(L, BD, cnt, P.cap, key);
          if (cnt == P.cap) tau = min(tau, L[(P.cap - 1) * BD] + 1);  // full: keep the cap smallest
        }
      });
    } else {
      u64 tau = PPP_KEY_INF;
      // queries outside the gridded rectangle: skip the rings that cannot contain any cell
      int du = max(max(-cu, cu - (g.nu - 1)), 0), dv = max(max(-cv, cv - (g.nv - 1)), 0);
      int R = max(P.R0, max(du, dv));
      int R_prev = -1;
      while (true) {
        visit_annulus(g, cu, cv, R_prev, R, [&](float4 c) {
          u64 key = make_key(d2_flann(qx, qy, qz, c.x, c.y, c.z), __float_as_int(c.w));
          if (key < tau) {
            list_insert(L, BD, cnt, P.cap, key);
            if (cnt == P.cap) tau = L[(P.cap - 1) * BD];
          }
        });
        if (cnt >= P.kk && (P.kk == 0 || key_d2(L[(P.kk - 1) * BD]) < ring_bound2(g, R, cu, cv))) break;
        if (block_covers_grid(g, cu, cv, R)) break;
        R_prev = R;
        R++;
      }
    }
  }
  // ---- outputs ----
  if (P.idx_out) {
    if (P.mode == 0) {
      int32_t* io = P.idx_out + row * (int64_t)P.cap;
      float* dout = P.d2_out ? P.d2_out + row * (int64_t)P.cap : nullptr;
      for (int j = 0; j < P.cap; j++) {
        u64 key = j < cnt ? L[j * BD] : 0;
        io[j] = j < cnt ? key_idx(key) : -1;
        if (dout) dout[j] = j < cnt ? key_d2(key) : CUDART_INF_F;
      }
    } else {
      int64_t o = P.offsets[row];
      for (int j = 0; j < cnt; j++) {
        u64 key = L[j * BD];
        P.idx_out[o + j] = key_idx(key);
        if (P.d2_out) P.d2_out[o + j] = key_d2(key);
      }
    }
  }
  if (P.normals) {
    float o[4];
    if (!fin || cnt < 3) {
      o[0] = o[1] = o[2] = o[3] = CUDART_NAN_F;
    } else {
      float acc[9] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
      float kx = 0.f, ky = 0.f, kz = 0.f;
      const bool shifted = (P.flags & PPP_COV_SHIFTED) != 0;
      if (shifted) {
        float4 f = __ldg(P.xyz4 + key_idx(L[0]));
        kx = f.x; ky = f.y; kz = f.z;
      }
      for (int j = 0; j < cnt; j++) {
        float4 a = __ldg(P.xyz4 + key_idx(L[j * BD]));
        float x = a.x, y = a.y, z = a.z;
        if (shifted) { x = __fsub_rn(x, kx); y = __fsub_rn(y, ky); z = __fsub_rn(z, kz); }
        accumulate_point(acc, x, y, z);
      }
      normal_from_accumulators(acc, cnt, qx, qy, qz, P.vpx, P.vpy, P.vpz, o);
    }
    store_normal(P.normals, row, P.nsf, o);
  }
}

__global__ void __launch_bounds__(128) k_radius_count(SearchParams P, int32_t* __restrict__ counts, int32_t* __restrict__ max_count) {
  int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int cnt = 0;
  if (t < P.nq) {
    const GridView& g = P.g;
    float qx, qy, qz;
    int64_t row;
    if (P.q) {
      const float* qp = P.q + t * P.q_sf;
      qx = __ldg(qp); qy = __ldg(qp + 1); qz = __ldg(qp + 2);
      row = t;
    } else {
      float4 p = __ldg(g.sorted + P.first + t);
      qx = p.x; qy = p.y; qz = p.z;
      row = __float_as_int(p.w);
    }
    if (finite3(qx, qy, qz) && g.n_sorted > 0) {
      int cu = cell_coord_raw(axis_of(qx, qy, qz, g.au), g.min_u, g.inv_h);
      int cv = cell_coord_raw(axis_of(qx, qy, qz, g.av), g.min_v, g.inv_h);
      const float r2 = P.r2;
      visit_annulus(g, cu, cv, -1, P.R0, [&](float4 c) {
        cnt += d2_flann(qx, qy, qz, c.x, c.y, c.z) < r2 ? 1 : 0;
      });
    }
    if (counts) counts[row] = cnt;
  }
  int m = cnt;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0 && m > 0) atomicMax(max_count, m);
}

// counts for non-finite self points that are not in the sorted array: zero-fill first
__global__ void k_fill_i32(int32_t* p, int64_t n, int32_t v) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = v;
}

// Self-mode outputs are written by original index of the *indexed* points only; rows of
// non-finite points get their defaults here (idx -1 / d2 inf / NaN normals).
__global__ void k_default_rows(const float4* __restrict__ xyz4, int64_t n, int cap, int32_t* idx_out, float* d2_out,
                               float* normals, int nsf) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float4 p = __ldg(xyz4 + i);
  if (p.w == p.w) return;
  if (idx_out)
    for (int j = 0; j < cap; j++) {
      idx_out[i * (int64_t)cap + j] = -1;
      if (d2_out) d2_out[i * (int64_t)cap + j] = CUDART_INF_F;
    }
  if (normals) {
    float o[4] = {CUDART_NAN_F, CUDART_NAN_F, CUDART_NAN_F, CUDART_NAN_F};
    store_normal(normals, i, nsf, o);
  }
}

int pick_block(ppp_ctx* ctx, int cap, int* block, size_t* smem) {
  // largest block in {128, 64, 32} whose list storage fits in opt-in shared memory
  size_t lim = ctx->smem_optin ? ctx->smem_optin : 48 * 1024;
  for (int b = 128; b >= 32; b >>= 1) {
    size_t need = (size_t)cap * 8 * b;
    if (need <= lim) { *block = b; *smem = need; return PPP_OK; }
  }
  ppp_set_error("neighbour list capacity %d exceeds the shared-memory path (max %zu entries per query)", cap,
                lim / (8 * 32));
  return PPP_ERR_UNSUPPORTED;
}

int radius_rings(const GridView& g, double r) {
  // need R*h - slack >= r
  // r comes from sqrt((float)(r*r)); a float32 d2 below r2 can belong to a point up to ~1e-7 beyond it
  r *= 1.0 + 1e-6;
  int R = (int)std::ceil((r + (double)g.slack) / (double)g.h - 1e-12);
  while ((double)R * (double)g.h - (double)g.slack < r) R++;
  return std::max(R, 0);
}

}  // namespace

static int launch_search(ppp_cloud* c, SearchParams& P) {
  ppp_ctx* ctx = c->ctx;
  int block; size_t smem;
  PPP_TRY(pick_block(ctx, std::max(P.cap, 1), &block, &smem));
  if (smem > 48 * 1024)
    PPP_CUDA(cudaFuncSetAttribute(k_search, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  if (P.nq > 0) {
    unsigned blocks = (unsigned)((P.nq + block - 1) / block);
    PPP_LAUNCH(ctx, P.mode == 0 ? (P.normals ? "knn_normals" : "knn") : (P.normals ? "radius_normals" : "radius_fill"),
               k_search, blocks, block, smem, P);
    PPP_CHECK_LAUNCH();
  }
  return PPP_OK;
}

int knn_launch(ppp_cloud* c, const GridStore& gs, const float* q_dev, int64_t nq, int q_stride_f, int64_t first, int k,
               int32_t* idx_dev, float* d2_dev, bool with_normals, const float vp[3], unsigned flags,
               float* normals_dev, int normal_stride_f) {
  ppp_ctx* ctx = c->ctx;
  SearchParams P{};
  P.g = gs.v; P.xyz4 = c->xyz4; P.q = q_dev; P.q_sf = q_stride_f; P.nq = nq; P.first = first;
  P.cap = k; P.kk = (int)std::min<int64_t>(k, c->n_finite); P.mode = 0; P.R0 = 2; P.r2 = 0;
  P.idx_out = idx_dev; P.d2_out = d2_dev; P.offsets = nullptr;
  P.normals = with_normals ? normals_dev : nullptr; P.nsf = normal_stride_f;
  P.vpx = vp ? vp[0] : 0; P.vpy = vp ? vp[1] : 0; P.vpz = vp ? vp[2] : 0; P.flags = flags;
  if (!q_dev && c->n_finite < c->n && first == 0) {
    unsigned blocks = (unsigned)((c->n + 255) / 256);
    PPP_LAUNCH(ctx, "default_rows", k_default_rows, blocks, 256, 0, (const float4*)c->xyz4, c->n, k, idx_dev, d2_dev,
               P.normals, normal_stride_f);
    PPP_CHECK_LAUNCH();
  }
  return launch_search(c, P);
}

int radius_count_launch(ppp_cloud* c, const GridStore& gs, const float* q_dev, int64_t nq, int q_stride_f,
                        int64_t first, float r2, int32_t* counts_dev) {
  ppp_ctx* ctx = c->ctx;
  SearchParams P{};
  P.g = gs.v; P.xyz4 = c->xyz4; P.q = q_dev; P.q_sf = q_stride_f; P.nq = nq; P.first = first;
  P.mode = 1; P.r2 = r2; P.R0 = radius_rings(gs.v, std::sqrt((double)r2));
  int32_t* mx = nullptr;
  PPP_TRY(dev_alloc(ctx, &mx, 1));
  PPP_CUDA(cudaMemsetAsync(mx, 0, 4, ctx->stream));
  if (!q_dev && c->n_finite < c->n && counts_dev) {
    unsigned blocks = (unsigned)((c->n + 255) / 256);
    PPP_LAUNCH(ctx, "fill_i32", k_fill_i32, blocks, 256, 0, counts_dev, c->n, 0);
    PPP_CHECK_LAUNCH();
  }
  if (nq > 0) {
    unsigned blocks = (unsigned)((nq + 127) / 128);
    PPP_LAUNCH(ctx, "radius_count", k_radius_count, blocks, 128, 0, P, counts_dev, mx);
    PPP_CHECK_LAUNCH();
  }
  int hmx = 0;
  PPP_CUDA(cudaMemcpyAsync(&hmx, mx, 4, cudaMemcpyDeviceToHost, ctx->stream));
  PPP_CUDA(cudaStreamSynchronize(ctx->stream));
  dev_free(ctx, mx);
  return hmx;  // >= 0: maximum neighbour count (used to size the fill pass)
}

int radius_fill_launch(ppp_cloud* c, const GridStore& gs, const float* q_dev, int64_t nq, int q_stride_f,
                       int64_t first, float r2, const int64_t* offsets_dev, int32_t* idx_dev, float* d2_dev) {
  // the caller has run radius_count_launch and passes the maximum count through d2-independent path:
  // recompute it here (cheap) so this entry stays self-contained.
  int mx = radius_count_launch(c, gs, q_dev, nq, q_stride_f, first, r2, nullptr);
  if (mx < 0) return mx;
  SearchParams P{};
  P.g = gs.v; P.xyz4 = c->xyz4; P.q = q_dev; P.q_sf = q_stride_f; P.nq = nq; P.first = first;
  P.cap = std::max(mx, 1); P.mode = 1; P.r2 = r2; P.R0 = radius_rings(gs.v, std::sqrt((double)r2));
  P.idx_out = idx_dev; P.d2_out = d2_dev; P.offsets = offsets_dev;
  return launch_search(c, P);
}

int normals_radius_launch(ppp_cloud* c, const GridStore& gs, int64_t first, int64_t count, float r2,
                          const float vp[3], unsigned flags, float* normals_dev, int normal_stride_f) {
  ppp_ctx* ctx = c->ctx;
  int mx = radius_count_launch(c, gs, nullptr, count, 0, first, r2, nullptr);
  if (mx < 0) return mx;
  SearchParams P{};
  P.g = gs.v; P.xyz4 = c->xyz4; P.q = nullptr; P.nq = count; P.first = first;
  P.cap = std::max(mx, 1); P.mode = 1; P.r2 = r2; P.R0 = radius_rings(gs.v, std::sqrt((double)r2));
  P.normals = normals_dev; P.nsf = normal_stride_f;
  P.vpx = vp ? vp[0] : 0; P.vpy = vp ? vp[1] : 0; P.vpz = vp ? vp[2] : 0; P.flags = flags;
  if (c->n_finite < c->n && first == 0) {
    unsigned blocks = (unsigned)((c->n + 255) / 256);
    PPP_LAUNCH(ctx, "default_rows", k_default_rows, blocks, 256, 0, (const float4*)c->xyz4, c->n, 0, (int32_t*)nullptr,
               (float*)nullptr, normals_dev, normal_stride_f);
    PPP_CHECK_LAUNCH();
  }
  return launch_search(c, P);
}
