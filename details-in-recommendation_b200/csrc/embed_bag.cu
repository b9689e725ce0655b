// Multi-hot / weighted bags: every (sample, field) holds a variable-length list of (id, weight) instead of
// one id -- what `myself_input_layer` exists for (models/DeepFM/deepFM.py:53, 77: "support multi-hot
// features"; the weighted column of dataset/SequenceTensorFlowDataset/test4.py:50-55, 113-116).
// [TF] _safe_embedding_lookup_sparse: entries with id < 0 or weight <= 0 are pruned, then per bag
//     sum    e = sum_i w_i T[id_i]
//     mean   e = sum_i w_i T[id_i] / sum_i w_i
//     sqrtn  e = sum_i w_i T[id_i] / sqrt(sum_i w_i^2)
// (an empty bag is the zero vector); the first-order term always combines with 'sum'
// (linear_model(sparse_combiner='sum'), deepFM.py:255-263).  FM, S and the logits are formed from the
// combined e exactly as in embed_fwd.cu.
//
// One warp per sample; LPR = K/4 lanes own one field's bag and walk it in order (deterministic sums),
// so a warp covers 32/LPR fields at a time.  Besides the forward outputs the kernel leaves what the
// backward needs per ENTRY: its sort key (global row, n_rows when pruned), the slot b*F+f it belongs
// to and its effective scale x = w / norm.  HBM/latency-bound gather: no tensor cores.
#include "common.cuh"

namespace dir {

struct BagFwdArgs {
  const float* table;
  int64_t row_stride;
  const float* lin;
  int64_t lin_stride;
  const float* bias;
  const int64_t* bag_offsets;  // [B*F + 1]
  const int64_t* bag_index;    // [nnz]
  const float* bag_weight;     // [nnz] or NULL
  const int64_t* field_offset;
  const int64_t* field_rows;
  int64_t n_rows;
  int64_t B;
  int F;
  int combiner;  // 0 sum, 1 mean, 2 sqrtn
  float* emb;
  float* S;
  float* first;
  float* fm;
  uint32_t* sort_keys;   // [nnz]
  uint32_t* entry_slot;  // [nnz]
  float* entry_x;        // [nnz]
  int* oob_flag;
};

template <int LPR>
__global__ void __launch_bounds__(256) embed_bag_fm_fwd_kernel(const BagFwdArgs a) {
  constexpr int K = LPR * 4;
  constexpr int RPW = 32 / LPR;  // bags a warp walks side by side
  const int lane = threadIdx.x & 31;
  const int slot = lane / LPR;
  const int sub = lane % LPR;
  const int64_t b = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (b >= a.B) return;  // whole warp leaves together
  const int F = a.F;
  float4 Sv = make_float4(0.f, 0.f, 0.f, 0.f);
  float4 Qv = make_float4(0.f, 0.f, 0.f, 0.f);
  float fo = 0.f;

  for (int f0 = 0; f0 < F; f0 += RPW) {
    const int f = f0 + slot;
    if (f < F) {
      const int64_t s = b * F + f;
      const int64_t j0 = __ldg(a.bag_offsets + s), j1 = __ldg(a.bag_offsets + s + 1);
      const int64_t lo = __ldg(a.field_offset + f);
      const int64_t nf = a.field_rows ? __ldg(a.field_rows + f) : a.n_rows - lo;
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
      float lin_acc = 0.f, wsum = 0.f, wsq = 0.f;
      for (int64_t j = j0; j < j1; ++j) {  // entries in order: deterministic sums
        const int64_t id = __ldg(a.bag_index + j);
        const float w = a.bag_weight ? __ldg(a.bag_weight + j) : 1.f;
        bool keep = id >= 0 && w > 0.f;
        if (keep && id >= nf) {  // TF's CPU Gather raises here; prune and report
          keep = false;
          if (a.oob_flag) *a.oob_flag = 1;
        }
        if (!keep) continue;
        const int64_t row = lo + id;
        const float4 t = __ldg(reinterpret_cast<const float4*>(a.table + row * a.row_stride) + sub);
        acc.x = __fadd_rn(acc.x, __fmul_rn(w, t.x));
        acc.y = __fadd_rn(acc.y, __fmul_rn(w, t.y));
        acc.z = __fadd_rn(acc.z, __fmul_rn(w, t.z));
        acc.w = __fadd_rn(acc.w, __fmul_rn(w, t.w));
        wsum = __fadd_rn(wsum, w);
        wsq = __fadd_rn(wsq, __fmul_rn(w, w));
        if (a.lin != nullptr && sub == 0)
          lin_acc = __fadd_rn(lin_acc, __fmul_rn(w, __ldg(a.lin + row * a.lin_stride)));
      }
      float norm = 1.f;
      if (a.combiner == 1) norm = wsum;
      if (a.combiner == 2) norm = __fsqrt_rn(wsq);
      float4 e = acc;
      if (a.combiner != 0 && wsum > 0.f) {  // an empty bag stays the zero vector
        e.x = __fdiv_rn(acc.x, norm);
        e.y = __fdiv_rn(acc.y, norm);
        e.z = __fdiv_rn(acc.z, norm);
        e.w = __fdiv_rn(acc.w, norm);
      }
      Sv.x += e.x; Sv.y += e.y; Sv.z += e.z; Sv.w += e.w;
      Qv.x = fmaf(e.x, e.x, Qv.x); Qv.y = fmaf(e.y, e.y, Qv.y);
      Qv.z = fmaf(e.z, e.z, Qv.z); Qv.w = fmaf(e.w, e.w, Qv.w);
      fo += lin_acc;
      stg_stream(a.emb + s * K + sub * 4, e);
      // what the backward needs per entry (one lane of the group writes)
      if (sub == 0 && a.sort_keys != nullptr) {
        for (int64_t j = j0; j < j1; ++j) {
          const int64_t id = __ldg(a.bag_index + j);
          const float w = a.bag_weight ? __ldg(a.bag_weight + j) : 1.f;
          const bool keep = id >= 0 && w > 0.f && id < nf;
          a.sort_keys[j] = keep ? (uint32_t)(lo + id) : (uint32_t)a.n_rows;
          a.entry_slot[j] = (uint32_t)s;
          a.entry_x[j] = keep ? (a.combiner == 0 ? w : __fdiv_rn(w, norm)) : 0.f;
        }
      }
    }
  }

  // sum the RPW bag slots: lanes with equal `sub` hold the same 4 embedding components
#pragma unroll
  for (int o = LPR; o < 32; o <<= 1) {
    Sv.x += __shfl_xor_sync(0xffffffffu, Sv.x, o);
    Sv.y += __shfl_xor_sync(0xffffffffu, Sv.y, o);
    Sv.z += __shfl_xor_sync(0xffffffffu, Sv.z, o);
    Sv.w += __shfl_xor_sync(0xffffffffu, Sv.w, o);
    Qv.x += __shfl_xor_sync(0xffffffffu, Qv.x, o);
    Qv.y += __shfl_xor_sync(0xffffffffu, Qv.y, o);
    Qv.z += __shfl_xor_sync(0xffffffffu, Qv.z, o);
    Qv.w += __shfl_xor_sync(0xffffffffu, Qv.w, o);
  }
  // 0.5 * sum_k ((sum_f e)^2 - sum_f e^2), deepFM.py:331-333
  float fmv = (Sv.x * Sv.x - Qv.x) + (Sv.y * Sv.y - Qv.y) + (Sv.z * Sv.z - Qv.z) +
              (Sv.w * Sv.w - Qv.w);
#pragma unroll
  for (int o = 1; o < LPR; o <<= 1) fmv += __shfl_xor_sync(0xffffffffu, fmv, o);
  fo = warp_sum(fo);
  if (a.S && slot == 0) *reinterpret_cast<float4*>(a.S + b * K + sub * 4) = Sv;
  if (lane == 0) {
    a.fm[b] = 0.5f * fmv;
    if (a.first) a.first[b] = fo + (a.bias ? __ldg(a.bias) : 0.f);
  }
}

}  // namespace dir

extern "C" int dir_embed_bag_fm_fwd(const float* table, int64_t row_stride, const float* lin,
                                    int64_t lin_stride, const float* bias, const int64_t* bag_offsets,
                                    const int64_t* bag_index, const float* bag_weight, int64_t nnz,
                                    const int64_t* field_offset, const int64_t* field_rows,
                                    int64_t n_rows, int64_t B, int F, int K, int combiner, float* emb,
                                    float* S, float* first, float* fm, uint32_t* sort_keys,
                                    uint32_t* entry_slot, float* entry_x, int* oob_flag,
                                    dir_stream_t stream) {
  using namespace dir;
  if (B < 0 || F <= 0 || nnz < 0) return fail(DIR_EINVAL, "embed_bag_fm_fwd: B >= 0, F > 0, nnz >= 0 required");
  if (K != 4 && K != 8 && K != 16 && K != 32 && K != 64)
    return fail(DIR_EINVAL, "embed_bag_fm_fwd: K must be one of 4, 8, 16, 32, 64");
  if (combiner < 0 || combiner > 2)
    return fail(DIR_EINVAL, "embed_bag_fm_fwd: combiner must be DIR_COMBINER_SUM, _MEAN or _SQRTN");
  if (B == 0) return 0;
  if (!table || !bag_offsets || !field_offset || !emb || !fm || (nnz > 0 && !bag_index))
    return fail(DIR_EINVAL, "embed_bag_fm_fwd: table, bag_offsets, bag_index, field_offset, emb, fm are required");
  if (lin && !first) return fail(DIR_EINVAL, "embed_bag_fm_fwd: `first` is required when `lin` is given");
  if (sort_keys && (!entry_slot || !entry_x))
    return fail(DIR_EINVAL, "embed_bag_fm_fwd: sort_keys needs entry_slot and entry_x");
  if (row_stride < K || (row_stride & 3))
    return fail(DIR_EINVAL, "embed_bag_fm_fwd: row_stride must be >= K and a multiple of 4");
  if (!aligned16(table) || !aligned16(emb) || !aligned16(S))
    return fail(DIR_EINVAL, "embed_bag_fm_fwd: table, emb and S must be 16-byte aligned");
  if (n_rows <= 0 || n_rows >= 0xffffffffLL || B * F >= 0x7fffffffLL || nnz >= 0x7fffffffLL)
    return fail(DIR_EINVAL, "embed_bag_fm_fwd: 0 < n_rows < 2^32-1, B*F < 2^31, nnz < 2^31 required");
  BagFwdArgs a{table, row_stride, lin, lin_stride, bias, bag_offsets, bag_index, bag_weight, field_offset,
               field_rows, n_rows, B, F, combiner, emb, S, first, fm, sort_keys, entry_slot, entry_x, oob_flag};
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const dim3 grid((unsigned)((B + 7) / 8));
  switch (K) {
    case 4: embed_bag_fm_fwd_kernel<1><<<grid, 256, 0, st>>>(a); break;
    case 8: embed_bag_fm_fwd_kernel<2><<<grid, 256, 0, st>>>(a); break;
    case 16: embed_bag_fm_fwd_kernel<4><<<grid, 256, 0, st>>>(a); break;
    case 32: embed_bag_fm_fwd_kernel<8><<<grid, 256, 0, st>>>(a); break;
    default: embed_bag_fm_fwd_kernel<16><<<grid, 256, 0, st>>>(a); break;
  }
  return launched("embed_bag_fm_fwd");
}
