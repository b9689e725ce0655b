// Column feed -> the layer's [B, F] inputs.
//
// The reference's input_fn hands the graph ONE TENSOR PER COLUMN: ids for the categorical columns,
// floats for the numeric ones (models/DeepCrossNetwork/train.py:127-156 decodes the csv into that
// dict; :57-100 builds the columns).  The dense [B,F] feature_index / feature_value pair the kernels
// read is this repo's resolved form of it: a categorical field is (id, 1.0), a numeric field is
// (row 0 of its one-row table, x).  Two thirds of that pair are constants, so the host ships the
// columns only -- ids[B, n_sparse] (int32 or int64) and x[B, n_dense] -- and this kernel widens
// them on the device, on the copy stream, right behind the H2D copy.  Pure data movement,
// HBM-bound: 12 B written per lookup, coalesced.
#include "common.cuh"

namespace dir {

template <typename IdT>
__global__ void __launch_bounds__(256)
expand_features_kernel(const IdT* __restrict__ sparse_index, const float* __restrict__ dense_value,
                       const int32_t* __restrict__ field_src, int64_t n, int F, int n_sparse,
                       int n_dense, int64_t* __restrict__ feature_index,
                       float* __restrict__ feature_value) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t o = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; o < n; o += stride) {
    const uint32_t b = (uint32_t)(o / F);  // n < 2^31 * F is checked by the caller
    const int f = (int)(o - (int64_t)b * F);
    const int src = __ldg(field_src + f);
    int64_t id = 0;
    float v = 1.f;
    if (src >= 0) {
      id = (int64_t)__ldg(sparse_index + (int64_t)b * n_sparse + src);
    } else {
      v = __ldg(dense_value + (int64_t)b * n_dense + (-src - 1));
    }
    feature_index[o] = id;
    feature_value[o] = v;
  }
}

}  // namespace dir

extern "C" int dir_expand_features(const void* sparse_index, int index_bytes,
                                   const float* dense_value, const int32_t* field_src, int64_t B,
                                   int F, int n_sparse, int n_dense, int64_t* feature_index,
                                   float* feature_value, dir_stream_t stream) {
  using namespace dir;
  if (B < 0 || F <= 0 || n_sparse < 0 || n_dense < 0 || n_sparse + n_dense > F)
    return fail(DIR_EINVAL, "expand_features: B >= 0, F > 0, n_sparse + n_dense <= F required");
  if (index_bytes != 4 && index_bytes != 8)
    return fail(DIR_EINVAL, "expand_features: index_bytes must be 4 (int32) or 8 (int64)");
  if (B == 0) return 0;
  if (!field_src || !feature_index || !feature_value || (n_sparse > 0 && !sparse_index) ||
      (n_dense > 0 && !dense_value))
    return fail(DIR_EINVAL, "expand_features: null pointer");
  if (B > 0x7fffffffLL) return fail(DIR_EINVAL, "expand_features: B too large");
  const int64_t n = B * F;
  const unsigned grid = (unsigned)((n + 255) / 256 < (int64_t)kSMs * 16 ? (n + 255) / 256 : kSMs * 16);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (index_bytes == 4)
    expand_features_kernel<int32_t><<<grid, 256, 0, st>>>(
        static_cast<const int32_t*>(sparse_index), dense_value, field_src, n, F, n_sparse, n_dense,
        feature_index, feature_value);
  else
    expand_features_kernel<int64_t><<<grid, 256, 0, st>>>(
        static_cast<const int64_t*>(sparse_index), dense_value, field_src, n, F, n_sparse, n_dense,
        feature_index, feature_value);
  return launched("expand_features");
}
