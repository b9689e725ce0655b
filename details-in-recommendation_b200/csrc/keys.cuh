// Sort key of one lookup: the global row of (field, id), or -- row-sharded over G ranks -- the composite
// (owner, local row) = (row mod G) * cap + row div G; the pruned key G * cap (= n_rows when G == 1 and cap = n_rows)
// for lookups the forward prunes (id < 0, value <= 0) and for ids beyond their field (which also raise oob_flag).
// Shared by dir_shard_keys (shard.cu) and the key kernel fused with the sort's first histogram (embed_bwd.cu).
#pragma once
#include "common.cuh"

namespace dir {

struct KeyArgs {
  const int64_t* idx;           // [B, F]
  const float* val;             // [B, F] or NULL
  const int64_t* field_offset;  // [F]
  const int64_t* field_rows;    // [F] or NULL
  int64_t n_rows;
  int64_t n;  // entries of the [B, n_sel] key list
  int F, G;
  int64_t cap;
  const int32_t* field_sel;  // [n_sel] or NULL (all fields)
  int n_sel;
  int* oob_flag;
};

// entry o = b * n_sel + j of the key list, with b and j already split
__device__ __forceinline__ uint32_t make_key_bj(const KeyArgs& a, int64_t o, uint32_t b, int j);

__device__ __forceinline__ uint32_t make_key(const KeyArgs& a, int64_t o) {
  const uint32_t b = (uint32_t)o / (uint32_t)a.n_sel;  // n < 2^31
  return make_key_bj(a, o, b, (int)((uint32_t)o - b * (uint32_t)a.n_sel));
}

__device__ __forceinline__ uint32_t make_key_bj(const KeyArgs& a, int64_t o, uint32_t b, int j) {
  int f = j;
  int64_t i = o;
  if (a.field_sel != nullptr) {
    f = __ldg(a.field_sel + j);
    i = (int64_t)b * a.F + f;
  }
  const int64_t id = __ldg(a.idx + i);
  const float v = a.val ? __ldg(a.val + i) : 1.f;
  const int64_t lo = __ldg(a.field_offset + f);
  const int64_t nf = a.field_rows ? __ldg(a.field_rows + f) : a.n_rows - lo;
  bool keep = id >= 0 && v > 0.f;
  if (keep && id >= nf) {
    keep = false;
    if (a.oob_flag) *a.oob_flag = 1;
  }
  const uint32_t row = (uint32_t)(lo + id);  // n_rows < 2^32
  uint32_t key = row;                        // one rank: the key is the global row
  if (a.G > 1) key = (row % (uint32_t)a.G) * (uint32_t)a.cap + row / (uint32_t)a.G;
  return keep ? key : (uint32_t)(a.G * a.cap);
}

// argument checks shared by the two entry points; fills `a` (without keys) and returns 0, or a negative code
inline int key_args(const char* what, const int64_t* feature_index, const float* feature_value,
                    const int64_t* field_offset, const int64_t* field_rows, int64_t n_rows, int64_t B, int F, int G,
                    const int32_t* field_sel, int n_sel, int* oob_flag, KeyArgs& a) {
  if (B < 0 || F <= 0 || G <= 0 || n_rows <= 0) return fail(DIR_EINVAL, "%s: B >= 0, F > 0, G > 0, n_rows > 0 required", what);
  const int64_t cap = (n_rows + G - 1) / G;
  if ((uint64_t)cap * (uint64_t)G >= 0xffffffffULL)
    return fail(DIR_EINVAL, "%s: ceil(n_rows / G) * G must be < 2^32-1", what);
  if (field_sel == nullptr) n_sel = F;
  if (n_sel < 0 || n_sel > F) return fail(DIR_EINVAL, "%s: 0 <= n_sel <= F required", what);
  if (B * F >= 0x7fffffffLL) return fail(DIR_EINVAL, "%s: B*F must be < 2^31", what);
  const int64_t n = B * n_sel;
  if (n > 0 && (!feature_index || !field_offset)) return fail(DIR_EINVAL, "%s: null pointer", what);
  a = KeyArgs{feature_index, feature_value, field_offset, field_rows, n_rows, n, F, G, cap, field_sel, n_sel, oob_flag};
  return 0;
}

}  // namespace dir
