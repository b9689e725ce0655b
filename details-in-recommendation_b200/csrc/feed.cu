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

// ------------------------------------------------------------------------------------------------
// tf.feature_column.input_layer in the reference's own DCN convention
// (models/DeepCrossNetwork/DeepCrossNetwork.py:126 over the columns of train.py:88-100): every dense
// column side by side, SORTED BY COLUMN NAME -- numeric columns pass through 1-wide, indicator columns
// are one-hot vectors of their id, embedding columns are the K-wide rows dir_embed_fm_fwd gathered
// (census: d = 5 + 9 + 16 + 7 + 6 + 8 = 51).  A column map built on the host says where each of the d
// output columns comes from; one thread per output element, coalesced stores.  The backward hands the
// embedding columns' slice of dL/dx0 back as the [B, F*K] upstream gradient of the fused backward
// (numeric / indicator columns carry no parameters).
namespace dir {

__global__ void __launch_bounds__(256)
input_layer_fwd_kernel(const float* __restrict__ numeric, int n_numeric,
                       const int64_t* __restrict__ ind_ids, int n_indicator,
                       const float* __restrict__ emb, int emb_width,
                       const int32_t* __restrict__ col_kind, const int32_t* __restrict__ col_src,
                       const int32_t* __restrict__ col_arg, int64_t n, int d, float* __restrict__ x0) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t o = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; o < n; o += stride) {
    const int64_t b = o / d;
    const int c = (int)(o - b * d);
    const int kind = __ldg(col_kind + c), src = __ldg(col_src + c);
    float v;
    if (kind == 0) {
      v = __ldg(numeric + b * n_numeric + src);
    } else if (kind == 1) {
      v = __ldg(ind_ids + b * n_indicator + src) == (int64_t)__ldg(col_arg + c) ? 1.f : 0.f;
    } else {
      v = __ldg(emb + b * emb_width + src);
    }
    x0[o] = v;
  }
}

__global__ void __launch_bounds__(256)
input_layer_bwd_kernel(const float* __restrict__ dx0, const int32_t* __restrict__ emb_col, int64_t n,
                       int d, int emb_width, float* __restrict__ u) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t o = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; o < n; o += stride) {
    const int64_t b = o / emb_width;
    const int c = __ldg(emb_col + (int)(o - b * emb_width));
    u[o] = c >= 0 ? __ldg(dx0 + b * d + c) : 0.f;
  }
}

static unsigned stream_grid(int64_t n) {
  const int64_t want = (n + 255) / 256;
  return (unsigned)(want < (int64_t)kSMs * 16 ? want : (int64_t)kSMs * 16);
}

}  // namespace dir

extern "C" int dir_input_layer_fwd(const float* numeric, int n_numeric, const int64_t* indicator_ids,
                                   int n_indicator, const float* emb, int emb_width,
                                   const int32_t* col_kind, const int32_t* col_src,
                                   const int32_t* col_arg, int64_t B, int d, float* x0,
                                   dir_stream_t stream) {
  using namespace dir;
  if (B < 0 || d <= 0 || n_numeric < 0 || n_indicator < 0 || emb_width < 0)
    return fail(DIR_EINVAL, "input_layer_fwd: B >= 0, d > 0 and non-negative widths required");
  if (B == 0) return 0;
  if (!col_kind || !col_src || !col_arg || !x0 || (n_numeric > 0 && !numeric) ||
      (n_indicator > 0 && !indicator_ids) || (emb_width > 0 && !emb))
    return fail(DIR_EINVAL, "input_layer_fwd: null pointer");
  if (B * (int64_t)d >= ((int64_t)1 << 40)) return fail(DIR_EINVAL, "input_layer_fwd: B*d too large");
  const int64_t n = B * d;
  input_layer_fwd_kernel<<<stream_grid(n), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      numeric, n_numeric, indicator_ids, n_indicator, emb, emb_width, col_kind, col_src, col_arg, n, d, x0);
  return launched("input_layer_fwd");
}

extern "C" int dir_input_layer_bwd(const float* dx0, const int32_t* emb_col, int64_t B, int d,
                                   int emb_width, float* u, dir_stream_t stream) {
  using namespace dir;
  if (B < 0 || d <= 0 || emb_width < 0) return fail(DIR_EINVAL, "input_layer_bwd: B >= 0, d > 0, emb_width >= 0 required");
  if (B == 0 || emb_width == 0) return 0;
  if (!dx0 || !emb_col || !u) return fail(DIR_EINVAL, "input_layer_bwd: null pointer");
  const int64_t n = B * emb_width;
  input_layer_bwd_kernel<<<stream_grid(n), 256, 0, static_cast<cudaStream_t>(stream)>>>(dx0, emb_col, n, d,
                                                                                        emb_width, u);
  return launched("input_layer_bwd");
}
