// Exclusive prefix sums over int32 counts (cell histograms, per-query neighbour counts,
// per-slice band sizes).  Three-phase block scan: per-tile reduce -> scan of tile sums
// (recursive) -> per-tile scan with carried base.  n+1 outputs: out[n] is the total.
#include <stdlib.h>

#include <algorithm>

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
  pdl_prologue();
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
  pdl_prologue();
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
__global__ void k_scan_zero_total(T* out) {
  pdl_prologue(); out[0] = 0; }

// Single-pass exclusive scan (chained tiles with decoupled look-back).  A tile takes its number from a ticket
// counter -- so every tile it may wait for is already running --, scans its 4096 items, publishes
// {1: tile sum | 2: inclusive prefix} as ONE 64-bit word (flag << 32 | value) and walks back over its
// predecessors' words until it meets an inclusive prefix.  state[0] is the ticket, state[1 + t] tile t's word.
// The tile with the last ticket waits until every word is an inclusive prefix (a tile's last access), then zeroes the
// words and the ticket for the next call: no memset between calls.
__global__ void __launch_bounds__(SCAN_THREADS) k_scan_chained(const int32_t* __restrict__ in, int64_t n, int32_t* __restrict__ out,
                                                               unsigned long long* __restrict__ state, int tiles) {
  pdl_prologue();
  __shared__ int s_tile;
  __shared__ int32_t s_prefix;
  if (threadIdx.x == 0) s_tile = (int)atomicAdd(state, 1ull);
  __syncthreads();
  const int tile = s_tile;
  PPP_DEV_ASSERT(tile >= 0 && tile < tiles);
  volatile unsigned long long* words = state + 1;
  const int64_t base = (int64_t)tile * SCAN_TILE + (int64_t)threadIdx.x * SCAN_ITEMS;
  int32_t v[SCAN_ITEMS];
  int32_t s = 0;
  if (base + SCAN_ITEMS <= n && (((uintptr_t)in & 15) == 0)) {
    const int4 a = __ldg(reinterpret_cast<const int4*>(in + base)), b = __ldg(reinterpret_cast<const int4*>(in + base) + 1);
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; i++) s += v[i];
  } else {
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; i++) {
      const int64_t j = base + i;
      v[i] = j < n ? in[j] : 0;
      s += v[i];
    }
  }
  int32_t tot;
  const int32_t ex = block_excl_scan<int32_t>(s, &tot);
  if (threadIdx.x < 32) {
    // the first warp walks back over the predecessors' words 32 at a time: lane l looks at tile (base - l), the walk
    // ends at the nearest inclusive prefix (a tile before the first counts as inclusive 0)
    const int lane = threadIdx.x;
    int32_t prefix = 0;
    if (tile > 0) {
      if (lane == 0) words[tile] = (1ull << 32) | (unsigned long long)(uint32_t)tot;
      for (int base = tile - 1;; base -= 32) {
        const int p = base - lane;
        unsigned long long w = 2ull << 32;
        if (p >= 0) { do { w = words[p]; } while ((w >> 32) == 0ull); }
        const unsigned incl = __ballot_sync(0xffffffffu, (w >> 32) == 2ull);
        const int stop = incl ? __ffs(incl) - 1 : 31;            // nearest inclusive prefix in this window
        prefix += __reduce_add_sync(0xffffffffu, lane <= stop ? (int32_t)(uint32_t)w : 0);
        if (incl) break;
      }
    }
    if (lane == 0) {
      words[tile] = (2ull << 32) | (unsigned long long)(uint32_t)(prefix + tot);
      s_prefix = prefix;
    }
  }
  __syncthreads();
  int32_t run = ex + s_prefix;
#pragma unroll
  for (int i = 0; i < SCAN_ITEMS; i++) {
    const int64_t j = base + i;
    if (j < n) out[j] = run;
    run += v[i];
  }
  const int64_t last = n - 1;
  if (last >= base && last < base + SCAN_ITEMS) out[n] = run;
  if (tile == tiles - 1) {
    // a tile's last access to the words is the store of its inclusive prefix: once all are there nobody reads any more
    for (int i = threadIdx.x; i < tiles; i += SCAN_THREADS)
      while ((words[i] >> 32) != 2ull) {}
    __syncthreads();
    for (int i = threadIdx.x; i <= tiles; i += SCAN_THREADS) state[i] = 0ull;
  }
}

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
  const int64_t tiles = (n + SCAN_TILE - 1) / SCAN_TILE;
  static const bool three_phase = getenv("PPP_SCAN_3PHASE") != nullptr;   // comparison aid
  if (tiles <= 1 || tiles > (1 << 20) || three_phase) return scan_impl<int32_t, int32_t>(ctx, in, out, n, 1);
  // one state array per working stream (the two streams may scan at the same time); grown on demand, zeroed once
  const int si = ctx->stream == ctx->aux_stream ? 1 : 0;
  if (ctx->scan_state_cap[si] < tiles + 1) {
    if (ctx->scan_state[si]) PPP_CUDA(cudaFreeAsync(ctx->scan_state[si], ctx->stream));
    const int64_t cap = std::max<int64_t>(2 * (tiles + 1), 4096);
    PPP_CUDA(cudaMallocAsync((void**)&ctx->scan_state[si], (size_t)cap * 8, ctx->stream));
    PPP_CUDA(cudaMemsetAsync(ctx->scan_state[si], 0, (size_t)cap * 8, ctx->stream));
    ctx->scan_state_cap[si] = cap;
  }
  PPP_LAUNCH(ctx, "scan_chained", k_scan_chained, (unsigned)tiles, SCAN_THREADS, 0, in, n, out, ctx->scan_state[si], (int)tiles);
  PPP_CHECK_LAUNCH();
  return PPP_OK;
}
int scan_exclusive_i32_to_i64(ppp_ctx* ctx, const int32_t* in, int64_t* out, int64_t n) {
  return scan_impl<int32_t, long long>(ctx, in, (long long*)out, n, 1);
}

bool ppp_pdl_enabled() {
  static const bool on = [] { const char* e = getenv("PPP_PDL"); return !(e && e[0] == '0'); }();
  return on;
}
