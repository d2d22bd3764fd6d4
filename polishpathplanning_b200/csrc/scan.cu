// Exclusive prefix sums over int32 counts (cell histograms, per-query neighbour counts,
// per-slice band sizes).  Three-phase block scan: per-tile reduce -> scan of tile sums
// (recursive) -> per-tile scan with carried base.  n+1 outputs: out[n] is the total.
#include "ppp_internal.cuh"

namespace {

constexpr int SCAN_THREADS = 512;
constexpr int SCAN_ITEMS = 8;
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;  // 4096

template <typename T>
__device__ __forceinline__ T warp_incl_scan(T v) {
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    T t = __shfl_up_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) >= o) v += t;
  }
  return v;
}

// Block-wide exclusive scan of one value per thread; returns exclusive prefix, total in *total.
template <typename T>
__device__ __forceinline__ T block_excl_scan(T v, T* total) {
  __shared__ T warp_sums[SCAN_THREADS / 32];
  __shared__ T block_total;
  int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  T inc = warp_incl_scan(v);
  if (lane == 31) warp_sums[w] = inc;
  __syncthreads();
  if (w == 0) {
    T s = lane < SCAN_THREADS / 32 ? warp_sums[lane] : T(0);
    T si = warp_incl_scan(s);
    if (lane < SCAN_THREADS / 32) warp_sums[lane] = si - s;
    if (lane == SCAN_THREADS / 32 - 1) block_total = si;
  }
  __syncthreads();
  T r = inc - v + warp_sums[w];
  *total = block_total;
  __syncthreads();
  return r;
}

template <typename TI, typename TO>
__global__ void __launch_bounds__(SCAN_THREADS) k_scan_tile_sums(const TI* __restrict__ in, int64_t n, TO* __restrict__ tile_sums) {
  int64_t base = (int64_t)blockIdx.x * SCAN_TILE;
  TO s = 0;
#pragma unroll
  for (int i = 0; i < SCAN_ITEMS; i++) {
    int64_t j = base + (int64_t)i * SCAN_THREADS + threadIdx.x;
    if (j < n) s += (TO)in[j];
  }
  TO tot;
  block_excl_scan<TO>(s, &tot);
  if (threadIdx.x == 0) tile_sums[blockIdx.x] = tot;
}

template <typename TI, typename TO>
__global__ void __launch_bounds__(SCAN_THREADS) k_scan_tiles(const TI* __restrict__ in, int64_t n, const TO* __restrict__ tile_base,
                                                             TO* __restrict__ out, int write_total) {
  int64_t base = (int64_t)blockIdx.x * SCAN_TILE + (int64_t)threadIdx.x * SCAN_ITEMS;
  TO v[SCAN_ITEMS];
  TO s = 0;
#pragma unroll
  for (int i = 0; i < SCAN_ITEMS; i++) {
    int64_t j = base + i;
    v[i] = j < n ? (TO)in[j] : TO(0);
    s += v[i];
  }
  TO tot;
  TO ex = block_excl_scan<TO>(s, &tot);
  TO run = ex + (tile_base ? tile_base[blockIdx.x] : TO(0));
#pragma unroll
  for (int i = 0; i < SCAN_ITEMS; i++) {
    int64_t j = base + i;
    if (j < n) out[j] = run;
    run += v[i];
  }
  // the thread holding the last valid element also writes the grand total at out[n]
  if (write_total) {
    int64_t last = n - 1;
    if (last >= base && last < base + SCAN_ITEMS) out[n] = run;
  }
}

template <typename T>
__global__ void k_scan_zero_total(T* out) { out[0] = 0; }

template <typename TI, typename TO>
int scan_impl(ppp_ctx* ctx, const TI* in, TO* out, int64_t n, int write_total) {
  if (n <= 0) {
    if (write_total) {
      auto kz = k_scan_zero_total<TO>;
      PPP_LAUNCH(ctx, "scan_zero_total", kz, 1, 1, 0, out);
      PPP_CHECK_LAUNCH();
    }
    return PPP_OK;
  }
  int64_t tiles = (n + SCAN_TILE - 1) / SCAN_TILE;
  auto k_tiles = k_scan_tiles<TI, TO>;
  auto k_sums = k_scan_tile_sums<TI, TO>;
  if (tiles == 1) {
    PPP_LAUNCH(ctx, "scan_tiles", k_tiles, 1, SCAN_THREADS, 0, in, n, (const TO*)nullptr, out, write_total);
    PPP_CHECK_LAUNCH();
    return PPP_OK;
  }
  TO* sums = nullptr;
  TO* bases = nullptr;
  PPP_TRY(dev_alloc(ctx, &sums, (size_t)tiles));
  PPP_TRY(dev_alloc(ctx, &bases, (size_t)tiles + 1));
  PPP_LAUNCH(ctx, "scan_tile_sums", k_sums, (unsigned)tiles, SCAN_THREADS, 0, in, n, sums);
  PPP_CHECK_LAUNCH();
  PPP_TRY((scan_impl<TO, TO>(ctx, sums, bases, tiles, 0)));
  PPP_LAUNCH(ctx, "scan_tiles", k_tiles, (unsigned)tiles, SCAN_THREADS, 0, in, n, (const TO*)bases, out, write_total);
  PPP_CHECK_LAUNCH();
  dev_free(ctx, sums);
  dev_free(ctx, bases);
  return PPP_OK;
}

}  // namespace

int scan_exclusive_i32(ppp_ctx* ctx, const int32_t* in, int32_t* out, int64_t n) {
  return scan_impl<int32_t, int32_t>(ctx, in, out, n, 1);
}
int scan_exclusive_i32_to_i64(ppp_ctx* ctx, const int32_t* in, int64_t* out, int64_t n) {
  return scan_impl<int32_t, long long>(ctx, in, (long long*)out, n, 1);
}
