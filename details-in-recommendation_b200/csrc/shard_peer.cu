// Row-sharded tables, device-driven exchange over NVLink peer memory (no NCCL, no host read on the step).
//
// Every rank owns one exchange buffer per parity with the same layout (dir_peer_layout); peers store
// straight into it through peer-mapped addresses.  Sizes that depend on the data travel as headers and are
// read on the device, so every launch of a step is sized by capacities known up front and the whole step --
// id phase included -- can be replayed from a CUDA graph:
//
//   requester q                                         owner o
//   dir_shard_keys / sort / dir_shard_unique  (ids only)
//   dir_shard_ids_push    hdr[q] = (count, base_u), ids[q][0..count)  -->  o's buffer
//   ---- barrier (ids) ----
//                                                       dir_shard_slots      slot[row][q] = i + 1
//                                                       dir_shard_gather_send  T[row] -> q's rows[base_u + i],
//                                                                              w[row] -> q's w[base_u + i]
//   ---- barrier (rows) ----
//   dir_embed_fm_fwd on rows / w, indexed by inv
//   dir_embed_bwd_reduce_emit_to   per-distinct-row sums  -->  o's g[q][i]       (embed_bwd.cu)
//   dir_shard_g1_push              first-order sums       -->  o's g1[q][i]
//   dir_shard_dense_emit           one-row fields' sums   -->  everybody's dense[q][j]
//   ---- barrier (grads) ----
//                                                       dir_shard_owner_update  per local row: the requesters'
//                                                         sums added in rank order, fused row update
//                                                       dir_shard_slots(clear)
//   dir_shard_dense_apply   replicated one-row fields: ranks' sums added in rank order, same update everywhere
//
// The owner never sorts: a requester sends each row at most once, so slot[row][q] (one cell per local row and
// requester, written without atomics) tells an arrival whether an earlier rank asked for the same row; the first
// one merges.  Deterministic: no floating-point atomics, sums in rank order.  A cell is (epoch << 24) | (i + 1) and
// counts only while its epoch is the buffer's current one, so nothing has to be cleared after a step; the epoch
// (1..255, advanced on the device by the id push) wraps once per 255 uses, when the whole map is zeroed.
//
// Reference: the partitioner hook around the embedding variables, models/DeepFM/deepFM.py:163-175 (under a TF
// parameter-server cluster the variables are sharded by row and ids / IndexedSlices travel over gRPC).
#include "common.cuh"
#include "update.cuh"

namespace dir {

constexpr int kMaxG = 64;
constexpr int kPeerCtas = kSMs * 8;  // grid-stride launches: the real counts are known on the device only

__device__ __forceinline__ char* peer_buf(const dir_peer_layout& L, int q) {
  return reinterpret_cast<char*>(__ldg(L.peer_base + q));
}
__device__ __forceinline__ int seg_of(const int64_t* s, int G, int64_t i) {  // s[G+1] ascending, in shared memory
  int q = 0;
  while (q + 1 < G && i >= s[q + 1]) ++q;
  return q;
}

// requester: header + distinct local rows to every owner
__global__ void __launch_bounds__(256)
peer_ids_push_kernel(const dir_peer_layout L, const int32_t* __restrict__ ulocal,
                     const int64_t* __restrict__ owner_off, int64_t n_cap, int* err, uint32_t* epoch) {
  __shared__ int64_t s_off[kMaxG + 1];
  if (threadIdx.x <= L.G) s_off[threadIdx.x] = owner_off[threadIdx.x];
  if (epoch != nullptr && blockIdx.x == 0 && threadIdx.x == 0) {
    // this buffer's slot map enters a new epoch: epoch[0] = current (1..255), epoch[1] = "zero the map first"
    const uint32_t e = epoch[0];
    epoch[1] = e >= 255u ? 1u : 0u;
    epoch[0] = e >= 255u ? 1u : e + 1u;
  }
  __syncthreads();
  if (blockIdx.x == 0 && threadIdx.x < L.G) {
    const int o = threadIdx.x;
    int64_t cnt = s_off[o + 1] - s_off[o];
    if (cnt > L.seg_cap || s_off[L.G] > L.u_cap) {
      *err = 1;  // more distinct rows than the buffers were sized for: nothing is written out of bounds
      cnt = cnt > L.seg_cap ? L.seg_cap : cnt;
    }
    int64_t* hdr = reinterpret_cast<int64_t*>(peer_buf(L, o) + L.off_hdr) + (int64_t)L.rank * 4;
    hdr[0] = cnt;
    hdr[1] = s_off[o];
  }
  int64_t total = s_off[L.G];
  total = total < n_cap ? total : n_cap;
  const int64_t step = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += step) {
    const int o = seg_of(s_off, L.G, i);
    const int64_t j = i - s_off[o];
    if (j < L.seg_cap)
      reinterpret_cast<int32_t*>(peer_buf(L, o) + L.off_ids)[(int64_t)L.rank * L.seg_cap + j] = __ldg(ulocal + i);
  }
}

// owner: what arrived.  pre[q] = arrivals from requesters < q; base[q] = where q keeps my rows.
struct Arrivals {
  int64_t pre[kMaxG + 1];
  int64_t base[kMaxG];
};
__device__ __forceinline__ void load_arrivals(const dir_peer_layout& L, Arrivals& s) {
  if (threadIdx.x == 0) {
    const int64_t* hdr = reinterpret_cast<const int64_t*>(L.local + L.off_hdr);
    int64_t pre = 0;
    for (int q = 0; q < L.G; ++q) {
      s.pre[q] = pre;
      int64_t c = hdr[q * 4];
      c = c < 0 ? 0 : (c > L.seg_cap ? L.seg_cap : c);
      pre += c;
      s.base[q] = hdr[q * 4 + 1];
    }
    s.pre[L.G] = pre;
  }
  __syncthreads();
}

// owner: the slot map is zeroed when its epoch wraps (epoch[1] set by the id push): a launch that exits at once
// 254 times out of 255
__global__ void __launch_bounds__(256)
peer_slots_reset_kernel(uint32_t* __restrict__ slot, int64_t n_cells, const uint32_t* __restrict__ epoch) {
  if (epoch[1] == 0u) return;
  const int64_t step = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_cells; i += step) slot[i] = 0u;
}

// owner: slot[row * G + q] = (epoch << 24) | (i + 1) for arrival i of requester q
__global__ void __launch_bounds__(256)
peer_slots_kernel(const dir_peer_layout L, uint32_t* __restrict__ slot, int64_t n_local, int* err,
                  int64_t* zero_counter, const uint32_t* __restrict__ epoch) {
  __shared__ Arrivals s;
  if (zero_counter != nullptr && blockIdx.x == 0 && threadIdx.x == 0) *zero_counter = 0;
  load_arrivals(L, s);
  const uint32_t tag = epoch[0] << 24;
  const int32_t* ids = reinterpret_cast<const int32_t*>(L.local + L.off_ids);
  const int64_t total = s.pre[L.G];
  const int64_t step = (int64_t)gridDim.x * blockDim.x;
  for (int64_t a = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; a < total; a += step) {
    const int q = seg_of(s.pre, L.G, a);
    const int64_t i = a - s.pre[q];
    const int64_t r = __ldg(ids + (int64_t)q * L.seg_cap + i);
    if (r < 0 || r >= n_local) {
      *err = 2;
      continue;
    }
    slot[r * L.G + q] = tag | (uint32_t)(i + 1);
  }
}

// owner: fused gather + send.  The arrivals are cut into tiles of 256 consecutive arrivals of one requester
// (16 KB of contiguous destination rows); a CTA takes tiles t = blockIdx.x, + gridDim.x, ... where tile t belongs to
// requester (rank + 1 + t) mod G: every destination is being written all the time, and no two ranks start on the
// same one (processing requester 0's rows first on every rank makes rank 0's NVLink ingress the bottleneck for all
// of them).  LPR lanes carry one row (16 bytes each); a thread has up to four row loads in flight before its first
// NVLink store.  The first-order weights of the tile go out lane-per-arrival (128-byte contiguous stores per warp).
// CTA 0 also refreshes the replicated one-row fields' rows behind this rank's own rows.
constexpr int kTileArr = 256;
template <int LPR>
__global__ void __launch_bounds__(256)
peer_gather_send_kernel(const dir_peer_layout L, const float* __restrict__ table, int64_t row_stride,
                        const float* __restrict__ lin, int64_t lin_stride,
                        const float* __restrict__ dense_table, int64_t dense_stride,
                        const float* __restrict__ dense_lin) {
  constexpr int K = LPR * 4;
  constexpr int UN = LPR < 4 ? LPR : 4;  // row loads in flight per thread
  __shared__ Arrivals s;
  __shared__ char* s_peer[kMaxG];
  __shared__ int64_t s_tiles;
  if (threadIdx.x < L.G) s_peer[threadIdx.x] = peer_buf(L, threadIdx.x);
  load_arrivals(L, s);
  if (threadIdx.x == 0) {
    int64_t m = 0;
    for (int q = 0; q < L.G; ++q) {
      const int64_t c = (s.pre[q + 1] - s.pre[q] + kTileArr - 1) / kTileArr;
      m = c > m ? c : m;
    }
    s_tiles = m;
  }
  __syncthreads();
  const int32_t* ids = reinterpret_cast<const int32_t*>(L.local + L.off_ids);
  const int G = L.G;
  const int64_t T = s_tiles * G;
  for (int64_t t = blockIdx.x; t < T; t += gridDim.x) {
    const int q = (int)((L.rank + 1 + t % G) % G);
    const int64_t i0 = (t / G) * kTileArr;
    const int64_t cnt = s.pre[q + 1] - s.pre[q];
    if (i0 >= cnt) continue;
    const int n = (int)(cnt - i0 < kTileArr ? cnt - i0 : kTileArr);
    const int32_t* tid_ids = ids + (int64_t)q * L.seg_cap + i0;
    const int64_t d0 = s.base[q] + i0;  // first destination row of the tile
    if (d0 < 0 || d0 + n > L.u_cap) continue;
    float4* drows = reinterpret_cast<float4*>(s_peer[q] + L.off_rows) + d0 * LPR;
#pragma unroll
    for (int k0 = 0; k0 < LPR; k0 += UN) {
      int64_t r[UN];
      int e[UN];
      float4 v[UN];
#pragma unroll
      for (int k = 0; k < UN; ++k) {
        e[k] = threadIdx.x + 256 * (k0 + k);  // float4 of the tile: arrival e / LPR, part e % LPR
        r[k] = e[k] / LPR < n ? (int64_t)__ldg(tid_ids + e[k] / LPR) : -1;
      }
#pragma unroll
      for (int k = 0; k < UN; ++k)
        if (r[k] >= 0) v[k] = __ldg(reinterpret_cast<const float4*>(table + r[k] * row_stride) + e[k] % LPR);
#pragma unroll
      for (int k = 0; k < UN; ++k)
        if (r[k] >= 0) drows[e[k]] = v[k];
    }
    if (lin != nullptr && (int)threadIdx.x < n)
      reinterpret_cast<float*>(s_peer[q] + L.off_w)[d0 + threadIdx.x] =
          __ldg(lin + (int64_t)__ldg(tid_ids + threadIdx.x) * lin_stride);
  }
  if (blockIdx.x == 0 && L.n_dense > 0 && dense_table != nullptr) {
    float* rows = reinterpret_cast<float*>(L.local + L.off_rows) + L.u_cap * K;
    float* w = reinterpret_cast<float*>(L.local + L.off_w) + L.u_cap;
    for (int e = threadIdx.x; e < L.n_dense * K; e += blockDim.x)
      rows[e] = dense_table[(int64_t)(e / K) * dense_stride + (e % K)];
    for (int j = threadIdx.x; j < L.n_dense; j += blockDim.x) w[j] = dense_lin ? dense_lin[j] : 0.f;
  }
}

// requester: first-order gradient sums of the distinct rows (g1_local[u], grouped by owner) -> owners
__global__ void __launch_bounds__(256)
peer_g1_push_kernel(const dir_peer_layout L, const float* __restrict__ g1_local,
                    const int64_t* __restrict__ owner_off, int64_t n_cap) {
  __shared__ int64_t s_off[kMaxG + 1];
  if (threadIdx.x <= L.G) s_off[threadIdx.x] = owner_off[threadIdx.x];
  __syncthreads();
  int64_t total = s_off[L.G];
  total = total < n_cap ? total : n_cap;
  const int64_t step = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += step) {
    const int o = seg_of(s_off, L.G, i);
    const int64_t j = i - s_off[o];
    if (j < L.seg_cap)
      reinterpret_cast<float*>(peer_buf(L, o) + L.off_g1)[(int64_t)L.rank * L.seg_cap + j] = __ldg(g1_local + i);
  }
}

// owner: merge + fused update.  A warp takes 32 arrivals at a time.  Lane-per-arrival stage: local row, then the
// row's slot cells -- does an earlier requester merge this row (then this arrival has nothing to do), which later
// ones contribute.  Vector stage over the arrivals that lead, compacted: LPR lanes per row, the row / accumulator /
// gradient loads of PB passes in flight together; the other requesters' sums are added in a fixed order, then the
// fused update (row and accumulator share a 128-byte line).
// NBUF = 2: the batch was exchanged as two micro-batches, each through an exchange buffer (and slot map) of its own;
// half s of rank q counts as virtual requester s * G + q, and the merge runs over both buffers in that order.
struct OwnerSrc {
  const char* local[2];      // the exchange buffers (same layout)
  const uint32_t* slot[2];
  const uint32_t* epoch[2];
};

template <int LPR, int NBUF>
__global__ void __launch_bounds__(256, LPR <= 4 ? 3 : 1)
peer_owner_update_kernel(const dir_peer_layout L, const OwnerSrc src, float* table, float* accum,
                         int64_t row_stride, float* lin, float* lin_accum, int64_t lin_stride, const LinOpt lo,
                         const RowRule rule, int64_t n_local, unsigned long long* n_unique) {
  constexpr int SLOTS = 32 / LPR;
  constexpr int PB = LPR <= 4 ? 2 : 1;
  constexpr unsigned FULL = 0xffffffffu;
  __shared__ Arrivals s[NBUF];
  uint32_t tag[NBUF];
  int64_t tot[NBUF];
#pragma unroll
  for (int b = 0; b < NBUF; ++b) {
    dir_peer_layout Lb = L;
    Lb.local = const_cast<char*>(src.local[b]);
    load_arrivals(Lb, s[b]);
    tag[b] = src.epoch[b][0];  // a cell counts only while it carries its buffer's current epoch
    tot[b] = s[b].pre[L.G];
  }
  const int64_t total = tot[0] + (NBUF > 1 ? tot[NBUF - 1] : 0);
  const bool adagrad = rule.opt != DIR_OPT_SGD;  // the rule keeps an accumulator
  const int lane = threadIdx.x & 31;
  const int sub = lane % LPR, grp = lane / LPR;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const uint64_t pol_row = policy_evict_first();
  const int G = L.G;
  unsigned long long leaders = 0;
  for (int64_t a0 = warp * 32; a0 < total; a0 += nwarps * 32) {  // warp-uniform trip count
    // ---- lane-per-arrival stage
    int64_t a = a0 + lane;
    bool lead = false;
    int64_t e = 0, r = 0;        // e: this arrival's entry of gbuf / g1buf, + (buffer << 40)
    unsigned long long more = 0; // virtual requesters after this one that asked for the same row
    if (a < total) {
      const int b = (NBUF > 1 && a >= tot[0]) ? 1 : 0;
      if (b) a -= tot[0];
      const int q = seg_of(s[b].pre, G, a);
      const int64_t i = a - s[b].pre[q];
      e = (int64_t)q * L.seg_cap + i;
      r = __ldg(reinterpret_cast<const int32_t*>(src.local[b] + L.off_ids) + e);
      e |= (int64_t)b << 40;
      if (r >= 0 && r < n_local) {
        bool earlier = false;
#pragma unroll
        for (int bb = 0; bb < NBUF; ++bb) {
          const uint32_t* sl = src.slot[bb] + r * G;
          uint32_t sv[8];
#pragma unroll
          for (int p = 0; p < 8; ++p) sv[p] = p < G ? __ldg(sl + p) : 0u;  // independent loads, one sector
#pragma unroll
          for (int p = 0; p < 8; ++p) {
            const bool on = (sv[p] >> 24) == tag[bb];
            const int v = bb * G + p, me = b * G + q;
            earlier |= on && v < me;
            if (on && v > me) more |= 1ull << v;
          }
          for (int p = 8; p < G; ++p) {
            const bool on = (__ldg(sl + p) >> 24) == tag[bb];
            const int v = bb * G + p, me = b * G + q;
            earlier |= on && v < me;
            if (on && v > me) more |= 1ull << v;
          }
        }
        lead = !earlier;  // else an earlier (virtual) requester merges this row
      }
    }
    const unsigned lm = __ballot_sync(FULL, lead);
    const int nl = __popc(lm);
    leaders += (unsigned long long)nl;
    // ---- vector stage over the leading arrivals
    for (int k0 = 0; k0 < nl; k0 += SLOTS * PB) {
      int64_t rr[PB], ee[PB];
      unsigned long long mm[PB];
      bool on[PB];
      float4 T[PB], A[PB], g[PB];
      float g1[PB], w1[PB], n1[PB], z1[PB];
#pragma unroll
      for (int j = 0; j < PB; ++j) {
        const int n = k0 + j * SLOTS + grp;
        on[j] = n < nl;
        const int from = on[j] ? (int)__fns(lm, 0, n + 1) : 0;
        rr[j] = __shfl_sync(FULL, r, from);
        ee[j] = __shfl_sync(FULL, e, from);
        mm[j] = __shfl_sync(FULL, more, from);
        T[j] = A[j] = g[j] = make_float4(0.f, 0.f, 0.f, 0.f);
        g1[j] = w1[j] = n1[j] = z1[j] = 0.f;
        if (on[j]) {
          const int64_t ro = rr[j] * row_stride;
          T[j] = ld_hint(table + ro + sub * 4, pol_row);
          if (adagrad) A[j] = ld_hint(accum + ro + sub * 4, pol_row);
          const char* base = src.local[NBUF > 1 ? (int)(ee[j] >> 40) : 0];
          const int64_t ent = ee[j] & ((1ll << 40) - 1);
          g[j] = __ldg(reinterpret_cast<const float4*>(base + L.off_g) + ent * LPR + sub);
          if (sub == 0 && lin != nullptr) {
            g1[j] = __ldg(reinterpret_cast<const float*>(base + L.off_g1) + ent);
            w1[j] = lin[rr[j] * lin_stride];
            lin_load(lo, lin_accum, rr[j] * lin_stride, n1[j], z1[j]);
          }
        }
      }
#pragma unroll
      for (int j = 0; j < PB; ++j) {
        if (!on[j]) continue;
        unsigned long long m = mm[j];
        while (m) {  // fixed order: buffer, then rank
          const int v = __ffsll((long long)m) - 1;
          m &= m - 1;
          const int bb = NBUF > 1 ? v / G : 0, p = NBUF > 1 ? v - bb * G : v;
          const uint32_t sp = __ldg(src.slot[bb] + rr[j] * G + p) & 0xffffffu;
          const int64_t e2 = (int64_t)p * L.seg_cap + (sp - 1u);
          const float4 o = __ldg(reinterpret_cast<const float4*>(src.local[bb] + L.off_g) + e2 * LPR + sub);
          g[j].x = __fadd_rn(g[j].x, o.x);
          g[j].y = __fadd_rn(g[j].y, o.y);
          g[j].z = __fadd_rn(g[j].z, o.z);
          g[j].w = __fadd_rn(g[j].w, o.w);
          if (sub == 0 && lin != nullptr)
            g1[j] = __fadd_rn(g1[j], __ldg(reinterpret_cast<const float*>(src.local[bb] + L.off_g1) + e2));
        }
        const int64_t ro = rr[j] * row_stride;
        T[j].x = upd_rule(T[j].x, g[j].x, A[j].x, rule);
        T[j].y = upd_rule(T[j].y, g[j].y, A[j].y, rule);
        T[j].z = upd_rule(T[j].z, g[j].z, A[j].z, rule);
        T[j].w = upd_rule(T[j].w, g[j].w, A[j].w, rule);
        *(reinterpret_cast<float4*>(table + ro) + sub) = T[j];
        if (adagrad) *(reinterpret_cast<float4*>(accum + ro) + sub) = A[j];
        if (lin != nullptr && sub == 0) {
          const int64_t off = rr[j] * lin_stride;
          lin_apply(lo, lin + off, lin_accum + off, lo.z + off, w1[j], n1[j], z1[j], g1[j]);
        }
      }
    }
  }
  if (lane == 0 && leaders && n_unique) atomicAdd(n_unique, leaders);
}

// replicated one-row fields: the G ranks' sums added in rank order, the same update applied to every replica;
// shard_row[j] >= 0 names the row of the sharded table to mirror into (on the rank that owns it)
struct DenseApplyArgs {
  float* table;  // replica: [n_dense] rows, row_stride apart (row | accumulator when Adagrad)
  float* accum;
  int64_t row_stride;
  float* lin;
  float* lin_accum;
  LinOpt lo;  // lo.z: [n_dense]
  RowRule rr;
  float* s_table;  // the sharded table's copy of the row
  float* s_accum;
  int64_t s_row_stride;
  float* s_lin;
  float* s_lin_accum;
  float* s_lin_z;
  int64_t s_lin_stride;
  const int64_t* shard_row;
  unsigned long long* n_unique;
};

template <int LPR>
__global__ void __launch_bounds__(128)
peer_dense_apply_kernel(const dir_peer_layout L, const char* local_b, const DenseApplyArgs a) {
  constexpr int K = LPR * 4;
  const int j = blockIdx.x, c = threadIdx.x;
  if (c > K) return;
  float g = 0.f, touched = 0.f;
  for (int b = 0; b < (local_b ? 2 : 1); ++b) {  // buffer (micro-batch), then rank: the same sum on every rank
    const float* gbuf = reinterpret_cast<const float*>((b ? local_b : L.local) + L.off_dense);
    for (int q = 0; q < L.G; ++q) {
      const float* src = gbuf + ((int64_t)q * L.n_dense + j) * (K + 4);
      g = (b == 0 && q == 0) ? src[c] : __fadd_rn(g, src[c]);
      touched += src[K + 1];
    }
  }
  if (touched == 0.f) return;  // no rank had a surviving lookup: the row is not touched
  const int64_t sr = __ldg(a.shard_row + j);
  const bool adagrad = a.rr.opt != DIR_OPT_SGD;
  if (c < K) {
    float* tp = a.table + (int64_t)j * a.row_stride + c;
    float acc = 0.f;
    if (adagrad) acc = a.accum[(int64_t)j * a.row_stride + c];
    const float t = upd_rule(*tp, g, acc, a.rr);
    *tp = t;
    if (adagrad) a.accum[(int64_t)j * a.row_stride + c] = acc;
    if (sr >= 0) {
      a.s_table[sr * a.s_row_stride + c] = t;
      if (adagrad) a.s_accum[sr * a.s_row_stride + c] = acc;
    }
  } else {
    if (a.lin != nullptr) {
      float n1, z1;
      lin_load(a.lo, a.lin_accum, j, n1, z1);
      lin_apply(a.lo, a.lin + j, a.lin_accum + j, a.lo.z + j, a.lin[j], n1, z1, g);
      if (sr >= 0) {
        a.s_lin[sr * a.s_lin_stride] = a.lin[j];
        if (a.lo.opt != DIR_OPT_SGD) a.s_lin_accum[sr * a.s_lin_stride] = a.lin_accum[j];
        if (a.lo.opt == DIR_OPT_FTRL) a.s_lin_z[sr * a.s_lin_stride] = a.lo.z[j];
      }
    }
    if (a.n_unique) atomicAdd(a.n_unique, 1ull);
  }
}

// inv[b, f] of the replicated one-row fields: their rows sit behind the exchanged ones (u_cap + j); an id other
// than 0 or a value <= 0 prunes the lookup ([TF] _safe_embedding_lookup_sparse)
__global__ void __launch_bounds__(256)
dense_inv_kernel(const int64_t* __restrict__ idx, const float* __restrict__ val,
                 const int32_t* __restrict__ fields, int n_fields, int64_t B, int F, int64_t tail,
                 int64_t* __restrict__ inv, int* oob_flag) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= B * n_fields) return;
  const int64_t b = t / n_fields;
  const int j = (int)(t % n_fields);
  const int64_t p = b * F + __ldg(fields + j);
  const int64_t id = __ldg(idx + p);
  const float v = val ? __ldg(val + p) : 1.f;
  if (id > 0 && v > 0.f && oob_flag) *oob_flag = 1;
  inv[p] = (id == 0 && v > 0.f) ? tail + j : -1;
}

static int check_layout(const char* what, const dir_peer_layout* L) {
  if (!L) return fail(DIR_EINVAL, "%s: layout is required", what);
  if (L->G <= 0 || L->G > kMaxG || L->rank < 0 || L->rank >= L->G)
    return fail(DIR_EINVAL, "%s: need 0 <= rank < G <= 64", what);
  if (L->K != 4 && L->K != 8 && L->K != 16 && L->K != 32 && L->K != 64)
    return fail(DIR_EINVAL, "%s: K must be one of 4, 8, 16, 32, 64", what);
  if (L->seg_cap <= 0 || L->u_cap <= 0 || L->n_dense < 0 || L->n_dense > 64)
    return fail(DIR_EINVAL, "%s: seg_cap, u_cap > 0 and 0 <= n_dense <= 64 required", what);
  if (!L->peer_base || !L->local) return fail(DIR_EINVAL, "%s: peer_base and local are required", what);
  if ((reinterpret_cast<uintptr_t>(L->local) & 255u) != 0)
    return fail(DIR_EINVAL, "%s: the exchange buffer must be 256-byte aligned", what);
  return 0;
}

}  // namespace dir

extern "C" int dir_peer_layout_init(int G, int rank, int K, int n_dense, int64_t seg_cap, int64_t u_cap,
                                    dir_peer_layout* out) {
  using namespace dir;
  if (!out) return fail(DIR_EINVAL, "peer_layout_init: out is required");
  if (G <= 0 || G > kMaxG || rank < 0 || rank >= G) return fail(DIR_EINVAL, "peer_layout_init: need 0 <= rank < G <= 64");
  if (K != 4 && K != 8 && K != 16 && K != 32 && K != 64)
    return fail(DIR_EINVAL, "peer_layout_init: K must be one of 4, 8, 16, 32, 64");
  if (seg_cap <= 0 || u_cap <= 0 || n_dense < 0 || n_dense > 64)
    return fail(DIR_EINVAL, "peer_layout_init: seg_cap, u_cap > 0 and 0 <= n_dense <= 64 required");
  out->G = G;
  out->rank = rank;
  out->K = K;
  out->n_dense = n_dense;
  out->seg_cap = seg_cap;
  out->u_cap = u_cap;
  size_t off = 0;
  auto take = [&](size_t bytes) {
    const size_t at = off;
    off += align_up(bytes, 256);
    return (int64_t)at;
  };
  out->off_hdr = take((size_t)G * 4 * 8);
  out->off_ids = take((size_t)G * seg_cap * 4);
  out->off_rows = take((size_t)(u_cap + n_dense) * K * 4);
  out->off_w = take((size_t)(u_cap + n_dense) * 4);
  out->off_g = take((size_t)G * seg_cap * K * 4);
  out->off_g1 = take((size_t)G * seg_cap * 4);
  out->off_dense = take((size_t)G * (n_dense > 0 ? n_dense : 1) * (K + 4) * 4);
  out->total_bytes = (int64_t)off;
  out->peer_base = nullptr;
  out->local = nullptr;
  return 0;
}

extern "C" int dir_shard_dense_inv(const int64_t* feature_index, const float* feature_value,
                                   const int32_t* onerow_fields, int n_onerow, int64_t B, int F, int64_t tail_row,
                                   int64_t* inv, int* oob_flag, dir_stream_t stream) {
  using namespace dir;
  if (B < 0 || F <= 0 || n_onerow < 0 || n_onerow > F) return fail(DIR_EINVAL, "shard_dense_inv: bad sizes");
  if (B == 0 || n_onerow == 0) return 0;
  if (!feature_index || !onerow_fields || !inv) return fail(DIR_EINVAL, "shard_dense_inv: null pointer");
  const int64_t n = B * n_onerow;
  dense_inv_kernel<<<(unsigned)((n + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      feature_index, feature_value, onerow_fields, n_onerow, B, F, tail_row, inv, oob_flag);
  return launched("shard_dense_inv");
}

extern "C" int dir_shard_ids_push(const dir_peer_layout* layout, const int32_t* unique_local_rows,
                                  const int64_t* owner_off, int64_t n_capacity, int* err_flag,
                                  uint32_t* slot_epoch, dir_stream_t stream) {
  using namespace dir;
  if (int rc = check_layout("shard_ids_push", layout)) return rc;
  if (!owner_off || !err_flag || n_capacity < 0 || (n_capacity > 0 && !unique_local_rows))
    return fail(DIR_EINVAL, "shard_ids_push: owner_off, err_flag and the id list are required");
  const int64_t want = (n_capacity + 255) / 256;
  const unsigned grid = (unsigned)(want < 1 ? 1 : (want < kPeerCtas ? want : kPeerCtas));
  peer_ids_push_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(*layout, unique_local_rows, owner_off,
                                                                           n_capacity, err_flag, slot_epoch);
  return launched("shard_ids_push");
}

extern "C" int dir_shard_slots(const dir_peer_layout* layout, uint32_t* slot, int64_t n_local_rows,
                               const uint32_t* slot_epoch, int* err_flag, int64_t* zero_counter, dir_stream_t stream) {
  using namespace dir;
  if (int rc = check_layout("shard_slots", layout)) return rc;
  if (!slot || !err_flag || !slot_epoch || n_local_rows <= 0)
    return fail(DIR_EINVAL, "shard_slots: slot, slot_epoch, err_flag, n_local_rows > 0 required");
  if (layout->seg_cap >= (1 << 24)) return fail(DIR_EINVAL, "shard_slots: seg_cap must be < 2^24");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  peer_slots_reset_kernel<<<kPeerCtas, 256, 0, st>>>(slot, n_local_rows * layout->G, slot_epoch);
  peer_slots_kernel<<<kPeerCtas, 256, 0, st>>>(*layout, slot, n_local_rows, err_flag, zero_counter, slot_epoch);
  return launched("shard_slots", 2);
}

extern "C" int dir_shard_gather_send(const dir_peer_layout* layout, const float* table, int64_t row_stride,
                                     const float* lin, int64_t lin_stride, const float* dense_table,
                                     int64_t dense_row_stride, const float* dense_lin, int ctas_per_sm,
                                     dir_stream_t stream) {
  using namespace dir;
  if (int rc = check_layout("shard_gather_send", layout)) return rc;
  const int K = layout->K;
  if (!table || row_stride < K || (row_stride & 3) || !aligned16(table))
    return fail(DIR_EINVAL, "shard_gather_send: table (16-byte aligned), row_stride >= K and a multiple of 4 required");
  if (layout->n_dense > 0 && (!dense_table || dense_row_stride < K))
    return fail(DIR_EINVAL, "shard_gather_send: the replicated rows are required when n_dense > 0");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const unsigned gs_grid = (unsigned)(kSMs * (ctas_per_sm >= 1 && ctas_per_sm <= 8 ? ctas_per_sm : 8));
#define DIR_GS(LP) \
  peer_gather_send_kernel<LP><<<gs_grid, 256, 0, st>>>(*layout, table, row_stride, lin, lin_stride, dense_table, \
                                                       dense_row_stride, dense_lin)
  switch (K / 4) {
    case 1: DIR_GS(1); break;
    case 2: DIR_GS(2); break;
    case 4: DIR_GS(4); break;
    case 8: DIR_GS(8); break;
    default: DIR_GS(16); break;
  }
#undef DIR_GS
  return launched("shard_gather_send");
}

extern "C" int dir_shard_g1_push(const dir_peer_layout* layout, const float* g1_local, const int64_t* owner_off,
                                 int64_t n_capacity, dir_stream_t stream) {
  using namespace dir;
  if (int rc = check_layout("shard_g1_push", layout)) return rc;
  if (!owner_off || n_capacity < 0 || (n_capacity > 0 && !g1_local))
    return fail(DIR_EINVAL, "shard_g1_push: owner_off and g1_local are required");
  const int64_t want = (n_capacity + 255) / 256;
  const unsigned grid = (unsigned)(want < 1 ? 1 : (want < kPeerCtas ? want : kPeerCtas));
  peer_g1_push_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(*layout, g1_local, owner_off, n_capacity);
  return launched("shard_g1_push");
}

extern "C" int dir_shard_owner_update(const dir_peer_layout* layout, const uint32_t* slot, float* table, float* accum,
                                      int64_t row_stride, float* lin, float* lin_accum, int64_t lin_stride,
                                      int64_t n_local_rows, const uint32_t* slot_epoch, int optimizer, float lr,
                                      const dir_table_opt* table_opt, const dir_linear_opt* linear_opt,
                                      const dir_peer_layout* layout_b,
                                      const uint32_t* slot_b, const uint32_t* slot_epoch_b, int64_t* n_unique_out,
                                      dir_stream_t stream) {
  using namespace dir;
  if (int rc = check_layout("shard_owner_update", layout)) return rc;
  if (layout_b != nullptr) {
    if (int rc = check_layout("shard_owner_update", layout_b)) return rc;
    if (!slot_b || !slot_epoch_b || layout_b->G != layout->G || layout_b->seg_cap != layout->seg_cap ||
        layout_b->off_g != layout->off_g || layout->G > 32)
      return fail(DIR_EINVAL, "shard_owner_update: the second buffer needs its slot map and epoch, the same layout, G <= 32");
  }
  const int K = layout->K;
  if (optimizer != DIR_OPT_SGD && optimizer != DIR_OPT_ADAGRAD && optimizer != DIR_OPT_PROXIMAL_ADAGRAD)
    return fail(DIR_EINVAL, "shard_owner_update: unknown optimizer");
  const RowRule rr{optimizer, lr, table_opt ? table_opt->l1 : 0.f, table_opt ? table_opt->l2 : 0.f};
  if (rr.l1 < 0.f || rr.l2 < 0.f) return fail(DIR_EINVAL, "shard_owner_update: l1, l2 must be >= 0");
  if (!slot || !slot_epoch || !table || n_local_rows <= 0)
    return fail(DIR_EINVAL, "shard_owner_update: slot, slot_epoch, table, n_local_rows > 0 required");
  if (optimizer != DIR_OPT_SGD && !accum) return fail(DIR_EINVAL, "shard_owner_update: Adagrad needs accum");
  if (row_stride < K || (row_stride & 3) || !aligned16(table) || !aligned16(accum))
    return fail(DIR_EINVAL, "shard_owner_update: rows must be 16-byte aligned, row_stride >= K and a multiple of 4");
  LinOpt lo;
  if (int rc = resolve_lin("shard_owner_update", linear_opt, optimizer, lr, lin, lin_accum, lo)) return rc;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  unsigned long long* nu = reinterpret_cast<unsigned long long*>(n_unique_out);  // += (zeroed by dir_shard_slots)
  OwnerSrc src{{layout->local, layout_b ? layout_b->local : nullptr}, {slot, slot_b}, {slot_epoch, slot_epoch_b}};
#define DIR_OU(LP)                                                                                                  \
  if (layout_b)                                                                                                     \
    peer_owner_update_kernel<LP, 2><<<kPeerCtas, 256, 0, st>>>(*layout, src, table, accum, row_stride, lin, lin_accum, \
                                                               lin_stride, lo, rr, n_local_rows, nu);                \
  else                                                                                                              \
    peer_owner_update_kernel<LP, 1><<<kPeerCtas, 256, 0, st>>>(*layout, src, table, accum, row_stride, lin, lin_accum, \
                                                               lin_stride, lo, rr, n_local_rows, nu)
  switch (K / 4) {
    case 1: DIR_OU(1); break;
    case 2: DIR_OU(2); break;
    case 4: DIR_OU(4); break;
    case 8: DIR_OU(8); break;
    default: DIR_OU(16); break;
  }
#undef DIR_OU
  return launched("shard_owner_update");
}

extern "C" int dir_shard_dense_apply(const dir_peer_layout* layout, float* dense_table, float* dense_accum,
                                     int64_t row_stride, float* dense_lin, float* dense_lin_accum, int optimizer,
                                     float lr, const dir_table_opt* table_opt, const dir_linear_opt* linear_opt,
                                     float* shard_table,
                                     float* shard_accum, int64_t shard_row_stride, float* shard_lin,
                                     float* shard_lin_accum, float* shard_lin_z, int64_t shard_lin_stride,
                                     const int64_t* shard_row, const dir_peer_layout* layout_b,
                                     int64_t* n_unique_inout, dir_stream_t stream) {
  using namespace dir;
  if (int rc = check_layout("shard_dense_apply", layout)) return rc;
  if (layout_b != nullptr) {
    if (int rc = check_layout("shard_dense_apply", layout_b)) return rc;
    if (layout_b->off_dense != layout->off_dense || layout_b->G != layout->G)
      return fail(DIR_EINVAL, "shard_dense_apply: the second buffer must have the same layout");
  }
  const char* local_b = layout_b ? layout_b->local : nullptr;
  if (layout->n_dense == 0) return 0;
  if (optimizer != DIR_OPT_SGD && optimizer != DIR_OPT_ADAGRAD && optimizer != DIR_OPT_PROXIMAL_ADAGRAD)
    return fail(DIR_EINVAL, "shard_dense_apply: unknown optimizer");
  const RowRule rr{optimizer, lr, table_opt ? table_opt->l1 : 0.f, table_opt ? table_opt->l2 : 0.f};
  if (!dense_table || !shard_row || !shard_table) return fail(DIR_EINVAL, "shard_dense_apply: null pointer");
  if (optimizer != DIR_OPT_SGD && (!dense_accum || !shard_accum))
    return fail(DIR_EINVAL, "shard_dense_apply: Adagrad needs the accumulators");
  LinOpt lo;
  if (int rc = resolve_lin("shard_dense_apply", linear_opt, optimizer, lr, dense_lin, dense_lin_accum, lo)) return rc;
  if (dense_lin && (!shard_lin || (lo.opt != DIR_OPT_SGD && !shard_lin_accum) || (lo.opt == DIR_OPT_FTRL && !shard_lin_z)))
    return fail(DIR_EINVAL, "shard_dense_apply: the sharded copies of the linear state are required");
  DenseApplyArgs a{dense_table, dense_accum, row_stride, dense_lin, dense_lin_accum, lo, rr,
                   shard_table, shard_accum, shard_row_stride, shard_lin, shard_lin_accum, shard_lin_z,
                   shard_lin_stride, shard_row, reinterpret_cast<unsigned long long*>(n_unique_inout)};
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int n = layout->n_dense;
  switch (layout->K / 4) {
    case 1: peer_dense_apply_kernel<1><<<n, 128, 0, st>>>(*layout, local_b, a); break;
    case 2: peer_dense_apply_kernel<2><<<n, 128, 0, st>>>(*layout, local_b, a); break;
    case 4: peer_dense_apply_kernel<4><<<n, 128, 0, st>>>(*layout, local_b, a); break;
    case 8: peer_dense_apply_kernel<8><<<n, 128, 0, st>>>(*layout, local_b, a); break;
    default: peer_dense_apply_kernel<16><<<n, 128, 0, st>>>(*layout, local_b, a); break;
  }
  return launched("shard_dense_apply");
}
