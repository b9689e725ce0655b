// Probe: what does a random 64-B / 4-B gather cost in DRAM traffic on B200, as a function of row
// stride and cudaLimitMaxL2FetchGranularity?  Run under ncu for dram__bytes_read.sum per kernel.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o gather_probe gather_probe.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

__device__ __forceinline__ uint32_t hash32(uint32_t x) {
  x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
  return x;
}

// 4 lanes per row (float4 each), 8 rows per warp-load, UNR loads in flight
// load flavours of the 16-byte row loads: 2..4 = ld.global.nc with an L2 prefetch-size hint
template <int MODE>
__device__ __forceinline__ float4 ld_mode(const float4* p) {
  float4 r;
  if (MODE == 2)
    asm volatile("ld.global.nc.L2::64B.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
  else if (MODE == 3)
    asm volatile("ld.global.nc.L2::128B.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
  else if (MODE == 4)
    asm volatile("ld.global.nc.L2::256B.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
  else if (MODE == 5)
    asm volatile("ld.global.nc.L1::no_allocate.L2::64B.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
  else
    r = __ldg(p);
  return r;
}

template <int UNR, int MODE>
__global__ void gather64m(const float* __restrict__ table, int64_t stride, uint32_t n_rows, int64_t n, float* out) {
  const int lane = threadIdx.x & 31, sub = lane & 3, slot = lane >> 2;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  float4 acc = make_float4(0, 0, 0, 0);
  const int64_t base = warp * 8 * UNR;
  if (base >= n) return;
  float4 t[UNR];
#pragma unroll
  for (int j = 0; j < UNR; ++j) {
    const uint32_t r = hash32((uint32_t)(base + j * 8 + slot)) % n_rows;
    t[j] = ld_mode<MODE>(reinterpret_cast<const float4*>(table + (int64_t)r * stride) + sub);
  }
#pragma unroll
  for (int j = 0; j < UNR; ++j) { acc.x += t[j].x; acc.y += t[j].y; acc.z += t[j].z; acc.w += t[j].w; }
  if (acc.x == 123.456f) out[0] = acc.y + acc.z + acc.w;
}

template <int UNR, bool NC>
__global__ void gather64(const float* __restrict__ table, int64_t stride, uint32_t n_rows, int64_t n, float* out) {
  const int lane = threadIdx.x & 31, sub = lane & 3, slot = lane >> 2;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  float4 acc = make_float4(0, 0, 0, 0);
  const int64_t base = warp * 8 * UNR;
  if (base >= n) return;
  float4 t[UNR];
#pragma unroll
  for (int j = 0; j < UNR; ++j) {
    const uint32_t r = hash32((uint32_t)(base + j * 8 + slot)) % n_rows;
    const float4* p = reinterpret_cast<const float4*>(table + (int64_t)r * stride) + sub;
    if (NC) t[j] = __ldg(p); else t[j] = *p;
  }
#pragma unroll
  for (int j = 0; j < UNR; ++j) { acc.x += t[j].x; acc.y += t[j].y; acc.z += t[j].z; acc.w += t[j].w; }
  if (acc.x == 123.456f) out[0] = acc.y + acc.z + acc.w;
}

// one lane per 4-B element
template <int UNR>
__global__ void gather4(const float* __restrict__ arr, int64_t stride, uint32_t n_rows, int64_t n, float* out) {
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t base = tid * UNR;
  if (base >= n) return;
  float t[UNR];
#pragma unroll
  for (int j = 0; j < UNR; ++j) {
    const uint32_t r = hash32((uint32_t)(base + j)) % n_rows;
    t[j] = __ldg(arr + (int64_t)r * stride);
  }
  float acc = 0;
#pragma unroll
  for (int j = 0; j < UNR; ++j) acc += t[j];
  if (acc == 123.456f) out[0] = acc;
}

// read-modify-write of a full 128-B line per row (8 lanes x float4), rows distinct
__global__ void rmw128(float* table, uint32_t n_rows, int64_t n, uint32_t step) {
  const int lane = threadIdx.x & 31, sub = lane & 7, slot = lane >> 3;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t i = warp * 4 + slot;
  if (i >= n) return;
  const uint32_t r = (uint32_t)((i * step) % n_rows);
  float4* p = reinterpret_cast<float4*>(table + (int64_t)r * 32) + sub;
  float4 v = *p;
  v.x += 1.f; v.y += 1.f; v.z += 1.f; v.w += 1.f;
  *p = v;
}

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1);} } while (0)

template <typename F>
float timeit(F f, int iters = 5) {
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  f(); cudaDeviceSynchronize();
  cudaEventRecord(a);
  for (int i = 0; i < iters; ++i) f();
  cudaEventRecord(b); cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms, a, b);
  return ms / iters * 1e3f;
}

int main(int argc, char** argv) {
  int gran = argc > 1 ? atoi(argv[1]) : 0;
  size_t cur = 0;
  CK(cudaDeviceGetLimit(&cur, cudaLimitMaxL2FetchGranularity));
  printf("default MaxL2FetchGranularity = %zu\n", cur);
  if (gran) { CK(cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, gran)); CK(cudaDeviceGetLimit(&cur, cudaLimitMaxL2FetchGranularity)); printf("set to %zu\n", cur); }
  const uint32_t N = 10000000;
  const int64_t n = 1 << 22;  // 4 M lookups
  float *t, *out;
  CK(cudaMalloc(&t, (size_t)N * 32 * 4));
  CK(cudaMemset(t, 0, (size_t)N * 32 * 4));
  CK(cudaMalloc(&out, 256));
  const int64_t warps = n / (8 * 5);
  const unsigned grid = (unsigned)((warps * 32 + 255) / 256);
  float us;
  us = timeit([&] { gather64<5, true><<<grid, 256>>>(t, 16, N, n, out); });
  printf("gather 64B rows, stride 64B  (ldg) : %8.1f us  %7.1f GB/s algorithmic\n", us, n * 64.0 / us / 1e3);
  us = timeit([&] { gather64<5, true><<<grid, 256>>>(t, 32, N, n, out); });
  printf("gather 64B rows, stride 128B (ldg) : %8.1f us  %7.1f GB/s algorithmic\n", us, n * 64.0 / us / 1e3);
  us = timeit([&] { gather64<5, false><<<grid, 256>>>(t, 32, N, n, out); });
  printf("gather 64B rows, stride 128B (ld)  : %8.1f us  %7.1f GB/s algorithmic\n", us, n * 64.0 / us / 1e3);
  us = timeit([&] { gather64m<5, 2><<<grid, 256>>>(t, 32, N, n, out); });
  printf("gather 64B rows, stride 128B (ld.nc.L2::64B)  : %8.1f us  %7.1f GB/s algorithmic\n", us, n * 64.0 / us / 1e3);
  us = timeit([&] { gather64m<5, 3><<<grid, 256>>>(t, 32, N, n, out); });
  printf("gather 64B rows, stride 128B (ld.nc.L2::128B) : %8.1f us  %7.1f GB/s algorithmic\n", us, n * 64.0 / us / 1e3);
  us = timeit([&] { gather64m<5, 4><<<grid, 256>>>(t, 32, N, n, out); });
  printf("gather 64B rows, stride 128B (ld.nc.L2::256B) : %8.1f us  %7.1f GB/s algorithmic\n", us, n * 64.0 / us / 1e3);
  us = timeit([&] { gather64m<5, 5><<<grid, 256>>>(t, 32, N, n, out); });
  printf("gather 64B rows, stride 128B (ld.nc.L1::no_allocate.L2::64B) : %8.1f us  %7.1f GB/s algorithmic\n", us, n * 64.0 / us / 1e3);
  const unsigned g4 = (unsigned)((n / 8 + 255) / 256);
  us = timeit([&] { gather4<8><<<g4, 256>>>(t, 1, N, n, out); });
  printf("gather 4B, stride 4B               : %8.1f us  %7.1f GB/s algorithmic\n", us, n * 4.0 / us / 1e3);
  us = timeit([&] { gather4<8><<<g4, 256>>>(t, 2, N, n, out); });
  printf("gather 4B, stride 8B               : %8.1f us  %7.1f GB/s algorithmic\n", us, n * 4.0 / us / 1e3);
  us = timeit([&] { gather4<8><<<g4, 256>>>(t, 32, 8 * N / 8, n, out); });
  printf("gather 4B, stride 128B             : %8.1f us  %7.1f GB/s algorithmic\n", us, n * 4.0 / us / 1e3);
  const int64_t nr = 1 << 21;
  const unsigned gr = (unsigned)((nr / 4 * 32 + 255) / 256);
  us = timeit([&] { rmw128<<<gr, 256>>>(t, N, nr, 4); });
  printf("rmw 128B lines, every 4th row      : %8.1f us  %7.1f GB/s algorithmic (r+w)\n", us, nr * 256.0 / us / 1e3);
  CK(cudaDeviceSynchronize());
  return 0;
}
