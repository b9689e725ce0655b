// Hand-written stable LSD radix sort of (key, lookup position) pairs for the backward's sorted list
// (uint32 keys of up to 32 bits, values = 0..n-1 implied by the input order), sized for this workload: a few
// million pairs that live in L2, a key space of 24-30 bits.
//
// Per 8-bit pass three small kernels instead of cub's onesweep (whose 290 large tiles of 8 832 pairs are
// latency-bound at this size: 26 us per pass for 27 MB of L2-resident traffic, profiles/r02_sharded_1gpu_ncu.txt):
//   count    tile histogram of the pass's digit, digit-major [256][tiles]
//   scan     one CTA per digit: exclusive prefix of its row over the tiles, and the row total
//   scatter  the tile's pairs re-read, ranked stably inside the tile and written to their places
// A tile is kRadixTile consecutive pairs, walked warp by warp in order, 32 consecutive pairs at a time:
// __match_any_sync gives every pair its rank among equal digits of the same 32, a per-warp running count the
// pairs of earlier rounds, a per-digit prefix over the warps the pairs of earlier warps.  Stable by construction
// (no atomics decide an order), so equal rows stay in lookup order and the gradient sums stay deterministic.
// The first pass reads the caller's keys and takes the position from the index (no iota array).
#pragma once
#include "common.cuh"

namespace dir {

constexpr int kRadixTile = 4096;   // pairs per CTA
constexpr int kRadixWarps = 8;     // 256 threads: 16 pairs per thread
constexpr int kRadixRounds = kRadixTile / (kRadixWarps * 32);

struct RadixTemp {
  uint32_t* alt_keys;  // [n]
  uint32_t* alt_vals;  // [n]
  uint32_t* hist;      // [256][tiles] digit-major tile counts, then [256] row totals
  size_t total;
};
inline int64_t radix_tiles(int64_t n) { return (n + kRadixTile - 1) / kRadixTile; }
inline size_t radix_hist_bytes(int64_t n) { return align_up((size_t)(256 * radix_tiles(n) + 256) * 4, 256); }

__global__ void __launch_bounds__(256)
radix_count_kernel(const uint32_t* __restrict__ keys, int64_t n, int shift, int64_t tiles,
                   uint32_t* __restrict__ hist, uint32_t* zero_a, unsigned long long* zero_b) {
  __shared__ uint32_t sh[256];
  if (blockIdx.x == 0 && threadIdx.x == 0) {  // counters of the consumer of the sorted list start at zero
    if (zero_a) *zero_a = 0u;
    if (zero_b) *zero_b = 0ull;
  }
  sh[threadIdx.x] = 0u;
  __syncthreads();
  const int64_t base = (int64_t)blockIdx.x * kRadixTile;
#pragma unroll
  for (int r = 0; r < kRadixRounds; ++r) {
    const int64_t i = base + r * 256 + threadIdx.x;
    if (i < n) atomicAdd(&sh[(__ldg(keys + i) >> shift) & 255u], 1u);  // integer: the count is order-free
  }
  __syncthreads();
  hist[(int64_t)threadIdx.x * tiles + blockIdx.x] = sh[threadIdx.x];
}

// block-wide exclusive scan of one value per thread (256 threads); returns the exclusive prefix, *total = the sum
__device__ __forceinline__ uint32_t block_excl_scan256(uint32_t v, uint32_t* s_warp /*[8]*/, uint32_t* total) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  uint32_t inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += t;
  }
  if (lane == 31) s_warp[w] = inc;
  __syncthreads();
  uint32_t base = 0, tot = 0;
#pragma unroll
  for (int ww = 0; ww < 8; ++ww) {
    const uint32_t c = s_warp[ww];
    if (ww < w) base += c;
    tot += c;
  }
  __syncthreads();  // s_warp may be reused by the caller's next chunk
  *total = tot;
  return base + inc - v;
}

// one CTA per digit: the row hist[d][0 .. tiles) becomes its exclusive prefix over the tiles, rowtot[d] its sum
__global__ void __launch_bounds__(256)
radix_rowscan_kernel(uint32_t* __restrict__ hist, int64_t tiles, uint32_t* __restrict__ rowtot) {
  __shared__ uint32_t s_warp[8];
  uint32_t* row = hist + (int64_t)blockIdx.x * tiles;
  uint32_t carry = 0;
  for (int64_t c0 = 0; c0 < tiles; c0 += 256) {
    const int64_t i = c0 + threadIdx.x;
    const uint32_t v = i < tiles ? row[i] : 0u;
    uint32_t tot;
    const uint32_t ex = block_excl_scan256(v, s_warp, &tot);
    if (i < tiles) row[i] = carry + ex;
    carry += tot;
  }
  if (threadIdx.x == 0) rowtot[blockIdx.x] = carry;
}

// FIRST: the values are the indices themselves (vin unused)
template <bool FIRST>
__global__ void __launch_bounds__(256)
radix_scatter_kernel(const uint32_t* __restrict__ kin, const uint32_t* __restrict__ vin, int64_t n, int shift,
                     int64_t tiles, const uint32_t* __restrict__ hist, uint32_t* __restrict__ kout,
                     uint32_t* __restrict__ vout) {
  __shared__ uint32_t wcount[kRadixWarps][256];  // pairs of digit d in warp w's part of the tile, then the prefix
  __shared__ uint32_t gbase[256];                // where digit d of this tile starts in the output
  __shared__ uint32_t s_warp[8];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  for (int d = lane; d < 256; d += 32) wcount[w][d] = 0u;
  {  // digit d starts after all pairs of smaller digits (scan of the 256 row totals) + this digit's earlier tiles
    uint32_t tot;
    const uint32_t digit_base = block_excl_scan256(__ldg(hist + 256 * tiles + threadIdx.x), s_warp, &tot);
    gbase[threadIdx.x] = digit_base + __ldg(hist + (int64_t)threadIdx.x * tiles + blockIdx.x);
  }
  __syncwarp();
  // warp w owns pairs [w * R * 32, (w + 1) * R * 32) of the tile, R = kRadixRounds, round r = 32 consecutive pairs
  const int64_t wbase = (int64_t)blockIdx.x * kRadixTile + (int64_t)w * kRadixRounds * 32;
  uint32_t key[kRadixRounds], val[kRadixRounds], rank[kRadixRounds];
#pragma unroll
  for (int r = 0; r < kRadixRounds; ++r) {
    const int64_t i = wbase + r * 32 + lane;
    key[r] = i < n ? __ldg(kin + i) : 0xffffffffu;
    val[r] = FIRST ? (uint32_t)i : (i < n ? __ldg(vin + i) : 0u);
  }
#pragma unroll
  for (int r = 0; r < kRadixRounds; ++r) {
    const int64_t i = wbase + r * 32 + lane;
    const bool live = i < n;
    const uint32_t d = (key[r] >> shift) & 255u;
    // lanes past the end must not match live ones: give them a digit of their own (256 + lane cannot collide)
    const unsigned peers = __match_any_sync(0xffffffffu, live ? d : 256u + (uint32_t)lane);
    const uint32_t before = live ? wcount[w][d] : 0u;  // pairs of digit d in this warp's earlier rounds
    __syncwarp();
    rank[r] = before + (uint32_t)__popc(peers & ((1u << lane) - 1u));
    if (live && (peers & ((1u << lane) - 1u)) == 0u) wcount[w][d] = before + (uint32_t)__popc(peers);  // one writer
    __syncwarp();
  }
  __syncthreads();
  {  // per digit: exclusive prefix over the warps (thread d), in warp order
    const int d = threadIdx.x;
    uint32_t run = 0;
#pragma unroll
    for (int ww = 0; ww < kRadixWarps; ++ww) {
      const uint32_t c = wcount[ww][d];
      wcount[ww][d] = run;
      run += c;
    }
  }
  __syncthreads();
#pragma unroll
  for (int r = 0; r < kRadixRounds; ++r) {
    const int64_t i = wbase + r * 32 + lane;
    if (i < n) {
      const uint32_t d = (key[r] >> shift) & 255u;
      const uint32_t dst = gbase[d] + wcount[w][d] + rank[r];
      kout[dst] = key[r];
      vout[dst] = val[r];
    }
  }
}

// Sorts (keys[i], i) by the low `end_bit` bits of the key, stable.  The result lands in (kout, vout); (alt_keys,
// alt_vals) is the other half of the ping-pong.  3 launches per 8-bit pass.
inline int radix_sort_pairs(const uint32_t* keys, int64_t n, int end_bit, uint32_t* kout, uint32_t* vout,
                            uint32_t* alt_keys, uint32_t* alt_vals, uint32_t* hist, uint32_t* zero_a,
                            unsigned long long* zero_b, cudaStream_t st) {
  const int passes = (end_bit + 7) / 8;
  const int64_t tiles = radix_tiles(n);
  const uint32_t* kin = keys;
  const uint32_t* vin = nullptr;
  for (int p = 0; p < passes; ++p) {
    const bool to_out = ((passes - 1 - p) & 1) == 0;  // the last pass writes (kout, vout)
    uint32_t* kd = to_out ? kout : alt_keys;
    uint32_t* vd = to_out ? vout : alt_vals;
    radix_count_kernel<<<(unsigned)tiles, 256, 0, st>>>(kin, n, p * 8, tiles, hist, p == 0 ? zero_a : nullptr,
                                                        p == 0 ? zero_b : nullptr);
    radix_rowscan_kernel<<<256, 256, 0, st>>>(hist, tiles, hist + 256 * tiles);
    if (p == 0) radix_scatter_kernel<true><<<(unsigned)tiles, 256, 0, st>>>(kin, vin, n, p * 8, tiles, hist, kd, vd);
    else radix_scatter_kernel<false><<<(unsigned)tiles, 256, 0, st>>>(kin, vin, n, p * 8, tiles, hist, kd, vd);
    kin = kd;
    vin = vd;
  }
  return passes * 3;
}

}  // namespace dir
