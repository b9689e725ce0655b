// Hand-written stable LSD radix sort of (key, lookup position) pairs for the backward's sorted list
// (uint32 keys of up to 32 bits, values = 0..n-1 implied by the input order), sized for this workload: a few
// million pairs that live in L2, a key space of 24-30 bits.
//
// Per 8-bit pass two small kernels instead of cub's onesweep (whose 290 large tiles of 8 832 pairs are
// latency-bound at this size: 26 us per pass for 27 MB of L2-resident traffic, profiles/r02_sharded_1gpu_ncu.txt):
//   row scan  one CTA per digit: exclusive prefix of the digit's tile counts over the tiles, and the row total
//   scatter   a tile's pairs ranked stably inside the tile, staged in shared memory in sorted order and written out
//             in runs (consecutive threads, consecutive addresses inside a digit's run); while writing, the pair's
//             NEXT digit is counted into the next pass's tile histogram (integer atomics: counts are order-free),
//             so only the first pass needs a count of its own -- radix_count_kernel, or the kernel that formed the
//             keys (keys_count_kernel in embed_bwd.cu: dir_shard_keys_sort).  Every global load of a tile is issued
//             before the first barrier: the kernel runs as one wave, so a tile's latency is the pass's time
// A tile is kRadixTile consecutive pairs, walked warp by warp in order, 32 consecutive pairs at a time: eight
// ballots (measured faster here than one __match_any_sync: 27 vs 30 us per pass) give every pair its rank among
// equal digits of the same 32, a per-warp running count the pairs of
// earlier rounds, a per-digit prefix over the warps the pairs of earlier warps.  Stable by construction (no atomic
// decides an order), so equal rows stay in lookup order and the gradient sums stay deterministic.
// The first pass reads the caller's keys and takes the position from the index (no iota array).
#pragma once
#include "common.cuh"

namespace dir {

constexpr int kRadixTile = 2048;   // pairs per CTA
constexpr int kRadixWarps = 8;     // 256 threads: 8 pairs per thread
constexpr int kRadixRounds = kRadixTile / (kRadixWarps * 32);

inline int64_t radix_tiles(int64_t n) { return (n + kRadixTile - 1) / kRadixTile; }
// two histograms ([256][tiles] digit-major tile counts + [256] row totals each): a pass scans one while its
// scatter fills the other for the next pass
inline size_t radix_hist_words(int64_t n) { return (size_t)(256 * radix_tiles(n) + 256); }
inline size_t radix_skew_words() { return 256; }  // per low digit: does it hold more than twice its uniform share?
inline size_t radix_hist_bytes(int64_t n) { return align_up((2 * radix_hist_words(n) + radix_skew_words()) * 4, 256); }

// first pass only: tile histogram of the low digit
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
    const bool live = i < n;
    const unsigned act = __ballot_sync(0xffffffffu, live);
    if (live) {  // one atomic per distinct digit of the warp: a Zipf-hot row would otherwise serialise 32 lanes
      const uint32_t d = (__ldg(keys + i) >> shift) & 255u;
      const unsigned peers = __match_any_sync(act, d);
      if ((peers & ((1u << (threadIdx.x & 31)) - 1u)) == 0u) atomicAdd(&sh[d], (uint32_t)__popc(peers));
    }
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

// one CTA per digit: the row hist[d][0 .. tiles) becomes its exclusive prefix over the tiles, rowtot[d] its sum;
// the same row of the OTHER histogram is zeroed for the scatter that follows to count into.  A thread owns up to
// kRowChunk consecutive tiles, so a row of up to 256 * kRowChunk tiles (8.4 M pairs) takes ONE block scan.
constexpr int kRowChunk = 16;
__global__ void __launch_bounds__(256)
radix_rowscan_kernel(uint32_t* __restrict__ hist, int64_t tiles, uint32_t* __restrict__ rowtot,
                     uint32_t* __restrict__ next_hist, uint32_t* __restrict__ skew, uint32_t skew_above) {
  __shared__ uint32_t s_warp[8];
  uint32_t* row = hist + (int64_t)blockIdx.x * tiles;
  uint32_t* nrow = next_hist ? next_hist + (int64_t)blockIdx.x * tiles : nullptr;
  const int64_t per = (tiles + 255) / 256;
  const int C = (int)(per < kRowChunk ? per : kRowChunk);  // tiles per thread and round
  uint32_t carry = 0;
  for (int64_t c0 = 0; c0 < tiles; c0 += (int64_t)256 * C) {
    const int64_t i0 = c0 + (int64_t)threadIdx.x * C;
    uint32_t v[kRowChunk];
    uint32_t sum = 0;
#pragma unroll
    for (int e = 0; e < kRowChunk; ++e) {
      v[e] = (e < C && i0 + e < tiles) ? row[i0 + e] : 0u;
      sum += v[e];
    }
    uint32_t tot;
    uint32_t run = carry + block_excl_scan256(sum, s_warp, &tot);
#pragma unroll
    for (int e = 0; e < kRowChunk; ++e) {
      if (e < C && i0 + e < tiles) {
        row[i0 + e] = run;
        run += v[e];
        if (nrow) nrow[i0 + e] = 0u;
      }
    }
    carry += tot;
  }
  if (threadIdx.x == 0) {
    rowtot[blockIdx.x] = carry;
    if (skew) skew[blockIdx.x] = carry > skew_above ? 1u : 0u;  // first pass: a hot digit means hot rows (Zipf)
  }
}

// FIRST: the values are the indices themselves (vin unused).  next_hist != NULL: count digit (shift + 8) of every
// pair into the tile it lands in.
template <bool FIRST>
__global__ void __launch_bounds__(256, 4)
radix_scatter_kernel(const uint32_t* __restrict__ kin, const uint32_t* __restrict__ vin, int64_t n, int shift,
                     int64_t tiles, const uint32_t* __restrict__ hist, uint32_t* __restrict__ kout,
                     uint32_t* __restrict__ vout, uint32_t* __restrict__ next_hist,
                     const uint32_t* __restrict__ skew) {
  __shared__ uint32_t wcount[kRadixWarps][256];  // pairs of digit d in warp w's part of the tile, then the prefix
  __shared__ uint32_t gbase[256];                // where digit d of this tile starts in the output
  __shared__ uint32_t tstart[256];               // where digit d starts inside the sorted tile
  __shared__ uint32_t s_warp[8];
  __shared__ uint32_t s_key[kRadixTile], s_val[kRadixTile];
  const unsigned FULL = 0xffffffffu;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  // warp w owns pairs [w * R * 32, (w + 1) * R * 32) of the tile, R = kRadixRounds, round r = 32 consecutive pairs.
  // Every global load of the tile is issued up front (one round trip to L2 instead of three behind barriers).
  const int64_t tbase = (int64_t)blockIdx.x * kRadixTile;
  const int64_t wbase = tbase + (int64_t)w * kRadixRounds * 32;
  uint32_t key[kRadixRounds], val[kRadixRounds], rank[kRadixRounds];
#pragma unroll
  for (int r = 0; r < kRadixRounds; ++r) {
    const int64_t i = wbase + r * 32 + lane;
    key[r] = i < n ? __ldg(kin + i) : 0u;
    val[r] = FIRST ? (uint32_t)i : (i < n ? __ldg(vin + i) : 0u);
  }
  const uint32_t row_total = __ldg(hist + 256 * tiles + threadIdx.x);
  const uint32_t tiles_before = __ldg(hist + (int64_t)threadIdx.x * tiles + blockIdx.x);
  const bool hot = next_hist != nullptr && __ldg(skew + threadIdx.x) != 0u;
  for (int d = lane; d < 256; d += 32) wcount[w][d] = 0u;
  // hot rows in this batch?  Then equal keys sit next to each other below and their counter updates are combined.
  const bool skewed = __syncthreads_or(hot) != 0;
  {  // digit d starts after all pairs of smaller digits (scan of the 256 row totals) + this digit's earlier tiles
    uint32_t tot;
    const uint32_t digit_base = block_excl_scan256(row_total, s_warp, &tot);
    gbase[threadIdx.x] = digit_base + tiles_before;
  }
#pragma unroll
  for (int r = 0; r < kRadixRounds; ++r) {
    const int64_t i = wbase + r * 32 + lane;
    const bool live = i < n;
    const uint32_t d = (key[r] >> shift) & 255u;
    unsigned peers = __ballot_sync(FULL, live);  // lanes of this round with my digit
#pragma unroll
    for (int b = 0; b < 8; ++b) {
      const unsigned m = __ballot_sync(FULL, (d >> b) & 1u);
      peers &= ((d >> b) & 1u) ? m : ~m;
    }
    const uint32_t before = live ? wcount[w][d] : 0u;  // pairs of digit d in this warp's earlier rounds
    __syncwarp();
    const unsigned lower = peers & ((1u << lane) - 1u);
    rank[r] = before + (uint32_t)__popc(lower);
    if (live && lower == 0u) wcount[w][d] = before + (uint32_t)__popc(peers);  // one writer per digit
    __syncwarp();
  }
  __syncthreads();
  {  // per digit (thread d): exclusive prefix over the warps, then over the digits: the sorted tile's layout
    const int d = threadIdx.x;
    uint32_t run = 0;
#pragma unroll
    for (int ww = 0; ww < kRadixWarps; ++ww) {
      const uint32_t c = wcount[ww][d];
      wcount[ww][d] = run;
      run += c;
    }
    uint32_t tot;
    tstart[d] = block_excl_scan256(run, s_warp, &tot);
  }
  __syncthreads();
#pragma unroll
  for (int r = 0; r < kRadixRounds; ++r) {
    const int64_t i = wbase + r * 32 + lane;
    if (i < n) {
      const uint32_t d = (key[r] >> shift) & 255u;
      const uint32_t pos = tstart[d] + wcount[w][d] + rank[r];
      s_key[pos] = key[r];
      s_val[pos] = val[r];
    }
  }
  __syncthreads();
  const int cnt = (int)(n - tbase < kRadixTile ? n - tbase : kRadixTile);
  for (int j = threadIdx.x; j < cnt; j += 256) {  // sorted order: runs of equal digits go to consecutive addresses
    const uint32_t k = s_key[j];
    const uint32_t d = (k >> shift) & 255u;
    const uint32_t dst = gbase[d] + ((uint32_t)j - tstart[d]);
    kout[dst] = k;
    vout[dst] = s_val[j];
    if (next_hist) {
      // integer atomics (counts are order-free).  With hot rows (first pass saw a digit holding more than twice its uniform share of the keys) one
      // per distinct counter of the warp: a Zipf-hot row gives runs of equal keys, 32 atomics on one counter each
      // (cfg5: 552 us per sort without this, 297 with); uniform ids skip the MATCH (cfg2: 99 us instead of 112)
      const uint32_t c = ((k >> (shift + 8)) & 255u) * (uint32_t)tiles + dst / kRadixTile;
      if (skewed) {
        const unsigned peers = __match_any_sync(__activemask(), c);
        if ((peers & ((1u << lane) - 1u)) == 0u) atomicAdd(next_hist + c, (uint32_t)__popc(peers));
      } else {
        atomicAdd(next_hist + c, 1u);
      }
    }
  }
}

// Sorts (keys[i], i) by the low `end_bit` bits of the key, stable.  The result lands in (kout, vout); (alt_keys,
// alt_vals) is the other half of the ping-pong.  2 launches per 8-bit pass + 1.
inline int radix_sort_pairs(const uint32_t* keys, int64_t n, int end_bit, uint32_t* kout, uint32_t* vout,
                            uint32_t* alt_keys, uint32_t* alt_vals, uint32_t* hist, uint32_t* zero_a,
                            unsigned long long* zero_b, cudaStream_t st, bool counted = false) {
  const int passes = (end_bit + 7) / 8;
  const int64_t tiles = radix_tiles(n);
  uint32_t* h[2] = {hist, hist + radix_hist_words(n)};
  uint32_t* skew = hist + 2 * radix_hist_words(n);
  const uint32_t* kin = keys;
  const uint32_t* vin = nullptr;
  // counted: the kernel that produced the keys already left the first pass's tile histogram in h[0] (and zeroed
  // the consumer's counters)
  if (!counted) radix_count_kernel<<<(unsigned)tiles, 256, 0, st>>>(kin, n, 0, tiles, h[0], zero_a, zero_b);
  for (int p = 0; p < passes; ++p) {
    const bool to_out = ((passes - 1 - p) & 1) == 0;  // the last pass writes (kout, vout)
    uint32_t* kd = to_out ? kout : alt_keys;
    uint32_t* vd = to_out ? vout : alt_vals;
    uint32_t* cur = h[p & 1];
    uint32_t* nxt = p + 1 < passes ? h[(p + 1) & 1] : nullptr;
    radix_rowscan_kernel<<<256, 256, 0, st>>>(cur, tiles, cur + 256 * tiles, nxt, p == 0 ? skew : nullptr,
                                              (uint32_t)(n / 128));
    if (p == 0) radix_scatter_kernel<true><<<(unsigned)tiles, 256, 0, st>>>(kin, vin, n, p * 8, tiles, cur, kd, vd, nxt, skew);
    else radix_scatter_kernel<false><<<(unsigned)tiles, 256, 0, st>>>(kin, vin, n, p * 8, tiles, cur, kd, vd, nxt, skew);
    kin = kd;
    vin = vd;
  }
  return passes * 2 + (counted ? 0 : 1);
}

}  // namespace dir
