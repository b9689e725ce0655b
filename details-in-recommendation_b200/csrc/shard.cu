// Row-sharded tables: routing kernels either side of the NCCL all-to-all.
//
// Tables are sharded by row over G ranks (owner = global row mod G, local row = global row div G:
// modulo, so hot low-numbered rows spread evenly).  The batch stays data-parallel.  Per step a rank
//   1. dir_shard_keys      forms one composite key per lookup, (owner, local row), owner-major
//   2. dir_embed_bwd_sort  sorts (key, lookup position)                      [embed_bwd.cu]
//   3. dir_shard_unique    numbers the distinct keys: only those cross NVLink
//   4. all-to-all of the distinct local rows; the owner answers with dir_rows_gather
//   5. dir_embed_fm_fwd    runs on the received unique-row buffer, indexed by `inv`  [embed_fwd.cu]
//   6. dir_embed_bwd_reduce_emit sums the gradients of each distinct row locally   [embed_bwd.cu]
//   7. all-to-all of those sums; the owner runs dir_embed_bwd_sort + dir_rows_reduce_update.
// The reference has no counterpart: its only hook is the partitioner wrapped around the embedding
// variables (models/DeepFM/deepFM.py:163-175), which under a TF parameter-server cluster shards
// variables by row and ships ids / IndexedSlices over gRPC.
#include <cub/device/device_scan.cuh>
#include <cub/iterator/counting_input_iterator.cuh>
#include <cub/iterator/transform_input_iterator.cuh>

#include "common.cuh"

namespace dir {

__global__ void __launch_bounds__(256)
shard_keys_kernel(const int64_t* __restrict__ idx, const float* __restrict__ val,
                  const int64_t* __restrict__ field_offset, const int64_t* __restrict__ field_rows,
                  int64_t n_rows, int64_t n, int F, int G, int64_t cap,
                  const int32_t* __restrict__ field_sel, int n_sel, uint32_t* __restrict__ keys,
                  int* oob_flag) {
  const int64_t o = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;  // entry of the [B, n_sel] key list
  if (o >= n) return;
  int f = (int)((uint32_t)o % (uint32_t)n_sel);  // n < 2^31
  int64_t i = o;
  if (field_sel != nullptr) {
    const uint32_t b = (uint32_t)o / (uint32_t)n_sel;
    f = __ldg(field_sel + f);
    i = (int64_t)b * F + f;
  }
  const int64_t id = __ldg(idx + i);
  const float v = val ? __ldg(val + i) : 1.f;
  const int64_t lo = __ldg(field_offset + f);
  const int64_t nf = field_rows ? __ldg(field_rows + f) : n_rows - lo;
  bool keep = id >= 0 && v > 0.f;
  if (keep && id >= nf) {
    keep = false;
    if (oob_flag) *oob_flag = 1;
  }
  const uint32_t row = (uint32_t)(lo + id);  // n_rows < 2^32
  uint32_t key = row;                        // one rank: the key is the global row
  if (G > 1) key = (row % (uint32_t)G) * (uint32_t)cap + row / (uint32_t)G;
  keys[o] = keep ? key : (uint32_t)(G * cap);
}

struct HeadFlag {
  const uint32_t* keys;
  uint32_t pruned;
  __host__ __device__ uint32_t operator()(int i) const {
    const uint32_t k = keys[i];
    return (k != pruned && (i == 0 || keys[i - 1] != k)) ? 1u : 0u;
  }
};

__global__ void __launch_bounds__(256)
shard_number_kernel(const uint32_t* __restrict__ keys, const uint32_t* __restrict__ pos,
                    const uint32_t* __restrict__ incl, int64_t n, uint32_t pruned, uint32_t cap,
                    uint32_t* __restrict__ uidx, int32_t* __restrict__ ulocal,
                    int64_t* __restrict__ inv) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint32_t k = __ldg(keys + i);
  const uint32_t p = __ldg(pos + i);
  if (k == pruned) {
    uidx[i] = 0u;
    inv[p] = -1;  // pruned by the forward kernel (id < 0)
    return;
  }
  const uint32_t u = __ldg(incl + i) - 1u;
  uidx[i] = u;
  inv[p] = (int64_t)u;
  if (i == 0 || __ldg(keys + i - 1) != k) ulocal[u] = (int32_t)(k % cap);
}

// owner_off[g] = number of distinct keys below g*cap, g = 0..G  (owner_off[G] = all of them)
__global__ void shard_bounds_kernel(const uint32_t* __restrict__ keys,
                                    const uint32_t* __restrict__ incl, int64_t n, int G,
                                    uint32_t cap, int64_t* __restrict__ owner_off) {
  const int g = threadIdx.x;
  if (g > G) return;
  const uint64_t target = (uint64_t)g * cap;
  int64_t lo = 0, hi = n;  // lower bound of target
  while (lo < hi) {
    const int64_t mid = (lo + hi) >> 1;
    if ((uint64_t)keys[mid] < target) lo = mid + 1; else hi = mid;
  }
  owner_off[g] = lo > 0 ? (int64_t)incl[lo - 1] : 0;
}

template <int LPR>
__global__ void __launch_bounds__(256)
rows_gather_kernel(const float* __restrict__ table, int64_t row_stride, const float* __restrict__ lin,
                   int64_t lin_stride, const int32_t* __restrict__ ids, int64_t n,
                   float* __restrict__ out, int64_t out_stride) {
  constexpr int K = LPR * 4;
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t i = t / LPR;
  const int sub = (int)(t % LPR);
  if (i >= n) return;
  const int64_t r = __ldg(ids + i);
  const float4 v = __ldg(reinterpret_cast<const float4*>(table + r * row_stride) + sub);
  stg_stream(out + i * out_stride + sub * 4, v);
  if (sub == 0) out[i * out_stride + K] = lin ? __ldg(lin + r * lin_stride) : 0.f;
}

// ---- exchange over NVLink peer memory (no NCCL on the payload path) ---------------------------------
// The n rows are grouped into G segments (seg_start[G+1]); segment q goes to rank q's buffer
// (peer_ptrs[q], a peer-mapped device pointer from symmetric memory) starting at row dst_row_off[q].
// Stores to a peer pointer travel over NVLink; the caller runs a cross-rank barrier afterwards.
__device__ __forceinline__ int segment_of(const int64_t* __restrict__ seg_start, int G, int64_t i) {
  int q = 0;
  while (q + 1 < G && i >= __ldg(seg_start + q + 1)) ++q;
  return q;
}

// owner side, fused gather + send: LPR lanes carry the row, one more lane carries (w, 0, 0, 0).
// A thread handles kRowsPerThread rows a grid-stride apart, loads first, so several 128-byte lines
// per thread are in flight before the first NVLink store.
constexpr int kRowsPerThread = 4;

template <int LPR>
__global__ void __launch_bounds__(256)
rows_gather_to_kernel(const float* __restrict__ table, int64_t row_stride, const float* __restrict__ lin,
                      int64_t lin_stride, const int32_t* __restrict__ ids, int64_t n, int G,
                      const int64_t* __restrict__ seg_start, const int64_t* __restrict__ peer_ptrs,
                      const int64_t* __restrict__ dst_row_off, int64_t out_stride) {
  constexpr int TPR = LPR + 1;  // threads per row
  {  // `n` bounds the launch; the rows really there are seg_start[G] (known on the device alone)
    const int64_t total = __ldg(seg_start + G);
    n = total < n ? total : n;
  }
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t i0 = t / TPR;
  const int sub = (int)(t % TPR);
  const int64_t step = ((int64_t)gridDim.x * blockDim.x) / TPR;
  float4 v[kRowsPerThread];
  int64_t r[kRowsPerThread];
#pragma unroll
  for (int k = 0; k < kRowsPerThread; ++k) {
    const int64_t i = i0 + k * step;
    r[k] = i < n ? (int64_t)__ldg(ids + i) : -1;
  }
#pragma unroll
  for (int k = 0; k < kRowsPerThread; ++k) {
    v[k] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (r[k] >= 0) {
      if (sub < LPR)
        v[k] = __ldg(reinterpret_cast<const float4*>(table + r[k] * row_stride) + sub);
      else if (lin)
        v[k].x = __ldg(lin + r[k] * lin_stride);
    }
  }
#pragma unroll
  for (int k = 0; k < kRowsPerThread; ++k) {
    const int64_t i = i0 + k * step;
    if (r[k] < 0) continue;
    const int q = segment_of(seg_start, G, i);
    float* dst = reinterpret_cast<float*>(__ldg(peer_ptrs + q)) +
                 (__ldg(dst_row_off + q) + i - __ldg(seg_start + q)) * out_stride;
    *(reinterpret_cast<float4*>(dst) + sub) = v[k];
  }
}

// requester side: ship rows [n, stride] to their owners' buffers
__global__ void __launch_bounds__(256)
rows_push_kernel(const float* __restrict__ src, int64_t n, int stride4, int G,
                 const int64_t* __restrict__ seg_start, const int64_t* __restrict__ peer_ptrs,
                 const int64_t* __restrict__ dst_row_off) {
  const int64_t t0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t step = (int64_t)gridDim.x * blockDim.x;
  const int64_t total = n * stride4;
  float4 v[kRowsPerThread];
#pragma unroll
  for (int k = 0; k < kRowsPerThread; ++k) {
    const int64_t t = t0 + k * step;
    v[k] = t < total ? ldg_stream(src + t * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
  }
#pragma unroll
  for (int k = 0; k < kRowsPerThread; ++k) {
    const int64_t t = t0 + k * step;
    if (t >= total) continue;
    const int64_t i = t / stride4;
    const int c = (int)(t % stride4);
    const int q = segment_of(seg_start, G, i);
    float4* dst = reinterpret_cast<float4*>(__ldg(peer_ptrs + q)) +
                  (__ldg(dst_row_off + q) + i - __ldg(seg_start + q)) * stride4 + c;
    *dst = v[k];
  }
}

// requester side, ids: ship the distinct local rows [n] (int32) to their owners' landing buffers.  Same segment
// arithmetic as rows_push_kernel with 4-byte elements; n bounds the launch, seg_start[G] is the real count.
__global__ void __launch_bounds__(256)
ids_push_kernel(const int32_t* __restrict__ src, int64_t n, int G, const int64_t* __restrict__ seg_start,
                const int64_t* __restrict__ peer_ptrs, const int64_t* __restrict__ dst_off) {
  const int64_t total = __ldg(seg_start + G);
  n = total < n ? total : n;
  const int64_t step = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += step) {
    const int q = segment_of(seg_start, G, i);
    int32_t* dst = reinterpret_cast<int32_t*>(__ldg(peer_ptrs + q)) + (__ldg(dst_off + q) + i - __ldg(seg_start + q));
    *dst = __ldg(src + i);
  }
}

struct UniqueWs {
  uint32_t* incl;
  void* cub_temp;
  size_t cub_bytes;
  size_t total;
};

static UniqueWs unique_carve(void* base, int64_t n) {
  UniqueWs w;
  char* p = static_cast<char*>(base);
  size_t off = 0;
  auto take = [&](size_t bytes) {
    char* r = p ? p + off : nullptr;
    off += align_up(bytes, 256);
    return r;
  };
  w.incl = reinterpret_cast<uint32_t*>(take((size_t)n * 4));
  size_t bytes = 0;
  cub::CountingInputIterator<int> cnt(0);
  cub::TransformInputIterator<uint32_t, HeadFlag, cub::CountingInputIterator<int>> it(cnt, HeadFlag{nullptr, 0});
  cub::DeviceScan::InclusiveSum(nullptr, bytes, it, (uint32_t*)nullptr, (int)n, (cudaStream_t)0);
  cudaGetLastError();
  const size_t floor_bytes = (size_t)n / 64 + (1u << 16);
  w.cub_bytes = bytes > floor_bytes ? bytes : floor_bytes;
  w.cub_temp = take(w.cub_bytes);
  w.total = off;
  return w;
}

}  // namespace dir

extern "C" int dir_shard_keys(const int64_t* feature_index, const float* feature_value,
                              const int64_t* field_offset, const int64_t* field_rows,
                              int64_t n_rows, int64_t B, int F, int G, const int32_t* field_sel,
                              int n_sel, uint32_t* keys, int* oob_flag, dir_stream_t stream) {
  using namespace dir;
  if (B < 0 || F <= 0 || G <= 0 || n_rows <= 0)
    return fail(DIR_EINVAL, "shard_keys: B >= 0, F > 0, G > 0, n_rows > 0 required");
  const int64_t cap = (n_rows + G - 1) / G;
  if ((uint64_t)cap * (uint64_t)G >= 0xffffffffULL)
    return fail(DIR_EINVAL, "shard_keys: ceil(n_rows / G) * G must be < 2^32-1");
  if (field_sel == nullptr) n_sel = F;
  if (n_sel < 0 || n_sel > F) return fail(DIR_EINVAL, "shard_keys: 0 <= n_sel <= F required");
  if (B * F >= 0x7fffffffLL) return fail(DIR_EINVAL, "shard_keys: B*F must be < 2^31");
  const int64_t n = B * n_sel;
  if (n == 0) return 0;
  if (!feature_index || !field_offset || !keys) return fail(DIR_EINVAL, "shard_keys: null pointer");
  shard_keys_kernel<<<(unsigned)((n + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      feature_index, feature_value, field_offset, field_rows, n_rows, n, F, G, cap, field_sel, n_sel, keys,
      oob_flag);
  return launched("shard_keys");
}

extern "C" size_t dir_shard_unique_workspace_bytes(int64_t n_lookups) {
  if (n_lookups <= 0) return 0;
  return dir::unique_carve(nullptr, n_lookups).total;
}

extern "C" int dir_shard_unique(const uint32_t* sorted_keys, const uint32_t* sorted_pos,
                                int64_t n_lookups, int64_t n_rows, int G, uint32_t* uidx,
                                int32_t* unique_local_rows, int64_t* inv, int64_t* owner_off,
                                void* workspace, size_t workspace_bytes, dir_stream_t stream) {
  using namespace dir;
  if (n_lookups < 0 || n_lookups >= 0x7fffffffLL || G <= 0 || G > 1023 || n_rows <= 0)
    return fail(DIR_EINVAL, "shard_unique: 0 <= n_lookups < 2^31, 0 < G <= 1023, n_rows > 0 required");
  if (!owner_off) return fail(DIR_EINVAL, "shard_unique: owner_off is required");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (n_lookups == 0) {
    cudaMemsetAsync(owner_off, 0, (size_t)(G + 1) * 8, st);
    return 0;
  }
  if (!sorted_keys || !sorted_pos || !uidx || !unique_local_rows || !inv || !workspace)
    return fail(DIR_EINVAL, "shard_unique: null pointer");
  const int64_t cap = (n_rows + G - 1) / G;
  const uint32_t pruned = (uint32_t)(cap * G);
  UniqueWs w = unique_carve(workspace, n_lookups);
  if (workspace_bytes < w.total) return fail(DIR_ENOMEM, "shard_unique: workspace too small");
  cub::CountingInputIterator<int> cnt(0);
  cub::TransformInputIterator<uint32_t, HeadFlag, cub::CountingInputIterator<int>> flags(
      cnt, HeadFlag{sorted_keys, pruned});
  size_t bytes = w.cub_bytes;
  cudaError_t e = cub::DeviceScan::InclusiveSum(w.cub_temp, bytes, flags, w.incl, (int)n_lookups, st);
  if (e != cudaSuccess) return fail(DIR_EIO, "shard_unique: %s", cudaGetErrorString(e));
  shard_number_kernel<<<(unsigned)((n_lookups + 255) / 256), 256, 0, st>>>(
      sorted_keys, sorted_pos, w.incl, n_lookups, pruned, (uint32_t)cap, uidx, unique_local_rows, inv);
  shard_bounds_kernel<<<1, 1024, 0, st>>>(sorted_keys, w.incl, n_lookups, G, (uint32_t)cap, owner_off);
  return launched("shard_unique", 4);
}

extern "C" int dir_rows_gather(const float* table, int64_t row_stride, const float* lin,
                               int64_t lin_stride, const int32_t* local_rows, int64_t n, int K,
                               float* out, int64_t out_stride, dir_stream_t stream) {
  using namespace dir;
  if (n < 0) return fail(DIR_EINVAL, "rows_gather: n >= 0 required");
  if (K != 4 && K != 8 && K != 16 && K != 32 && K != 64)
    return fail(DIR_EINVAL, "rows_gather: K must be one of 4, 8, 16, 32, 64");
  if (n == 0) return 0;
  if (!table || !local_rows || !out) return fail(DIR_EINVAL, "rows_gather: null pointer");
  if (row_stride < K || (row_stride & 3) || out_stride < K + 1 || (out_stride & 3))
    return fail(DIR_EINVAL, "rows_gather: strides must be multiples of 4, >= K (rows), >= K+1 (out)");
  if (!aligned16(table) || !aligned16(out)) return fail(DIR_EINVAL, "rows_gather: 16-byte alignment required");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int lpr = K / 4;
  const unsigned grid = (unsigned)((n * lpr + 255) / 256);
  switch (lpr) {
    case 1: rows_gather_kernel<1><<<grid, 256, 0, st>>>(table, row_stride, lin, lin_stride, local_rows, n, out, out_stride); break;
    case 2: rows_gather_kernel<2><<<grid, 256, 0, st>>>(table, row_stride, lin, lin_stride, local_rows, n, out, out_stride); break;
    case 4: rows_gather_kernel<4><<<grid, 256, 0, st>>>(table, row_stride, lin, lin_stride, local_rows, n, out, out_stride); break;
    case 8: rows_gather_kernel<8><<<grid, 256, 0, st>>>(table, row_stride, lin, lin_stride, local_rows, n, out, out_stride); break;
    default: rows_gather_kernel<16><<<grid, 256, 0, st>>>(table, row_stride, lin, lin_stride, local_rows, n, out, out_stride); break;
  }
  return launched("rows_gather");
}

extern "C" int dir_rows_gather_to(const float* table, int64_t row_stride, const float* lin,
                                  int64_t lin_stride, const int32_t* local_rows, int64_t n, int K, int G,
                                  const int64_t* seg_start, const int64_t* peer_ptrs,
                                  const int64_t* dst_row_off, int64_t out_stride, dir_stream_t stream) {
  using namespace dir;
  if (n < 0 || G <= 0) return fail(DIR_EINVAL, "rows_gather_to: n >= 0, G > 0 required");
  if (K != 4 && K != 8 && K != 16 && K != 32 && K != 64)
    return fail(DIR_EINVAL, "rows_gather_to: K must be one of 4, 8, 16, 32, 64");
  if (n == 0) return 0;
  if (!table || !local_rows || !seg_start || !peer_ptrs || !dst_row_off)
    return fail(DIR_EINVAL, "rows_gather_to: null pointer");
  if (row_stride < K || (row_stride & 3) || out_stride < K + 4 || (out_stride & 3))
    return fail(DIR_EINVAL, "rows_gather_to: strides must be multiples of 4, >= K (rows), >= K+4 (out)");
  if (!aligned16(table)) return fail(DIR_EINVAL, "rows_gather_to: 16-byte alignment required");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int lpr = K / 4;
  const unsigned grid = (unsigned)((n * (lpr + 1) + 256 * kRowsPerThread - 1) / (256 * kRowsPerThread));
#define DIR_GT(L) rows_gather_to_kernel<L><<<grid, 256, 0, st>>>(table, row_stride, lin, lin_stride, local_rows, n, G, seg_start, peer_ptrs, dst_row_off, out_stride)
  switch (lpr) {
    case 1: DIR_GT(1); break;
    case 2: DIR_GT(2); break;
    case 4: DIR_GT(4); break;
    case 8: DIR_GT(8); break;
    default: DIR_GT(16); break;
  }
#undef DIR_GT
  return launched("rows_gather_to");
}

extern "C" int dir_rows_push(const float* src, int64_t n, int64_t stride, int G, const int64_t* seg_start,
                             const int64_t* peer_ptrs, const int64_t* dst_row_off, dir_stream_t stream) {
  using namespace dir;
  if (n < 0 || G <= 0 || stride <= 0 || (stride & 3))
    return fail(DIR_EINVAL, "rows_push: n >= 0, G > 0, stride a positive multiple of 4 required");
  if (n == 0) return 0;
  if (!src || !seg_start || !peer_ptrs || !dst_row_off) return fail(DIR_EINVAL, "rows_push: null pointer");
  if (!aligned16(src)) return fail(DIR_EINVAL, "rows_push: 16-byte alignment required");
  const int stride4 = (int)(stride / 4);
  const unsigned grid = (unsigned)((n * stride4 + 256 * kRowsPerThread - 1) / (256 * kRowsPerThread));
  rows_push_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(src, n, stride4, G, seg_start, peer_ptrs,
                                                                         dst_row_off);
  return launched("rows_push");
}

extern "C" int dir_ids_push(const int32_t* src, int64_t n, int G, const int64_t* seg_start,
                            const int64_t* peer_ptrs, const int64_t* dst_off, dir_stream_t stream) {
  using namespace dir;
  if (n < 0 || G <= 0) return fail(DIR_EINVAL, "ids_push: n >= 0, G > 0 required");
  if (n == 0) return 0;
  if (!src || !seg_start || !peer_ptrs || !dst_off) return fail(DIR_EINVAL, "ids_push: null pointer");
  const int64_t want = (n + 255) / 256;
  const unsigned grid = (unsigned)(want < (int64_t)kSMs * 8 ? want : (int64_t)kSMs * 8);
  ids_push_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(src, n, G, seg_start, peer_ptrs, dst_off);
  return launched("ids_push");
}
