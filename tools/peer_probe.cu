// tools/peer_probe.cu -- how fast can an SM kernel push rows into a PEER's memory over NVLink?
// Variants of the owner-side "gather + send" of the sharded step (csrc/shard_peer.cu), timed by tools/peer_probe.py:
//   0  streaming copy, 16-byte loads / 16-byte stores per lane (LSU both ways)
//   1  random 64-byte rows gathered from a [n_src, 32]-float table, 16-byte LSU stores (what round 2 v1 did)
//   2  streaming copy staged through shared memory, one cp.async.bulk (TMA) store per 8 KB tile
//   3  random 64-byte rows gathered with cp.async (LDGSTS) into shared memory, one cp.async.bulk store per tile
// Built on its own: nvcc -gencode arch=compute_100a,code=sm_100a -shared -Xcompiler -fPIC -o libpeer_probe.so
#include <cuda_runtime.h>
#include <stdint.h>

namespace {

constexpr int kTileRows = 128;              // 128 rows x 64 B = 8 KB per bulk store
constexpr int kStages = 3;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void bulk_store(void* gdst, const void* ssrc, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(smem_u32(ssrc)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void cp_async16(void* sdst, const void* gsrc) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(sdst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

// variants 0 / 1: rows of 16 floats; ids == nullptr -> row i of src (stride 16), else src[ids[i]] (stride 32)
__global__ void __launch_bounds__(256) lsu_kernel(const float* __restrict__ src, const int* __restrict__ ids,
                                                  float* dst, int64_t n) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int sub = (int)(t & 3);
  const int64_t step = ((int64_t)gridDim.x * blockDim.x) >> 2;
  for (int64_t i0 = t >> 2; i0 < n; i0 += step * 4) {
    float4 v[4];
    int64_t r[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int64_t i = i0 + k * step;
      r[k] = i < n ? (ids ? (int64_t)__ldg(ids + i) * 32 : i * 16) : -1;
    }
#pragma unroll
    for (int k = 0; k < 4; ++k)
      if (r[k] >= 0) v[k] = __ldg(reinterpret_cast<const float4*>(src + r[k]) + sub);
#pragma unroll
    for (int k = 0; k < 4; ++k)
      if (r[k] >= 0) *(reinterpret_cast<float4*>(dst + (i0 + k * step) * 16) + sub) = v[k];
  }
}

// variants 2 / 3: persistent CTAs, tiles of kTileRows rows staged in shared memory, one bulk store per tile
__global__ void __launch_bounds__(256) bulk_kernel(const float* __restrict__ src, const int* __restrict__ ids,
                                                   float* dst, int64_t n) {
  __shared__ __align__(128) float tile[kStages][kTileRows * 16];
  const int64_t tiles = (n + kTileRows - 1) / kTileRows;
  int stage = 0;
  for (int64_t tl = blockIdx.x; tl < tiles; tl += gridDim.x) {
    const int64_t i0 = tl * kTileRows;
    const int rows = (int)((n - i0) < kTileRows ? (n - i0) : kTileRows);
    if (threadIdx.x == 0) bulk_wait_read<kStages - 1>();  // the store that last read this stage has drained
    __syncthreads();
    for (int e = threadIdx.x; e < rows * 4; e += blockDim.x) {
      const int r = e >> 2, sub = e & 3;
      const int64_t row = ids ? (int64_t)__ldg(ids + i0 + r) * 32 : (i0 + r) * 16;
      cp_async16(&tile[stage][r * 16 + sub * 4], src + row + sub * 4);
    }
    cp_async_commit();
    cp_async_wait<0>();
    fence_async();
    __syncthreads();
    if (threadIdx.x == 0) {
      bulk_store(dst + i0 * 16, tile[stage], (uint32_t)rows * 64u);
      bulk_commit();
    }
    stage = stage + 1 == kStages ? 0 : stage + 1;
  }
  if (threadIdx.x == 0) bulk_wait_all();
}

}  // namespace

extern "C" int probe_launch(int variant, const float* src, const int* ids, float* dst, int64_t n, int ctas,
                            void* stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (variant == 0) lsu_kernel<<<ctas, 256, 0, st>>>(src, nullptr, dst, n);
  else if (variant == 1) lsu_kernel<<<ctas, 256, 0, st>>>(src, ids, dst, n);
  else if (variant == 2) bulk_kernel<<<ctas, 256, 0, st>>>(src, nullptr, dst, n);
  else bulk_kernel<<<ctas, 256, 0, st>>>(src, ids, dst, n);
  return (int)cudaGetLastError();
}
