// gvv_microbench.cu -- L2 atomic throughput probe for the roofline denominators:
// R_min64 = red.global.min.u64 over a P-word buffer, R_add32 = red.global.add.f32 over a 3N-float
// buffer, one operation per thread, pseudo-random (hashed) addresses.  Device-timed.
#include <cstdio>
#include "gvv_internal.h"

namespace gvv {

__device__ __forceinline__ unsigned hash32(unsigned x) {
  x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16;
  return x;
}

__global__ void atomic_min64_kernel(unsigned long long* buf, long long nAddr, long long nOps, unsigned seed) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nOps) return;
  const unsigned h = hash32((unsigned)i ^ seed);
  atomicMin(buf + (h % (unsigned long long)nAddr), ((unsigned long long)hash32(h) << 32) | (unsigned)i);
}

__global__ void atomic_add32_kernel(float* buf, long long nAddr, long long nOps, unsigned seed) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nOps) return;
  const unsigned h = hash32((unsigned)i ^ seed);
  atomicAdd(buf + (h % (unsigned long long)nAddr), 1.0f);
}

}  // namespace gvv

extern "C" int gvv_bench_atomics(int32_t device, int32_t kind, int64_t nAddr, int64_t nOps, int32_t iters, double* opsPerS) {
  using namespace gvv;
  if (!opsPerS || nAddr <= 0 || nOps <= 0 || iters <= 0 || (kind != 0 && kind != 1)) return GVV_EINVAL;
  if (cudaSetDevice(device) != cudaSuccess) return GVV_ECUDA;
  void* buf = nullptr;
  const size_t bytes = (size_t)nAddr * (kind == 0 ? 8 : 4);
  if (cudaMalloc(&buf, bytes) != cudaSuccess) return GVV_ENOMEM;
  cudaMemset(buf, kind == 0 ? 0xff : 0, bytes);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int threads = 256;
  const unsigned blocks = (unsigned)((nOps + threads - 1) / threads);
  auto run = [&](unsigned seed) {
    if (kind == 0) atomic_min64_kernel<<<blocks, threads>>>((unsigned long long*)buf, nAddr, nOps, seed);
    else atomic_add32_kernel<<<blocks, threads>>>((float*)buf, nAddr, nOps, seed);
  };
  for (int i = 0; i < 3; ++i) run(1000u + i);
  cudaEventRecord(e0);
  for (int i = 0; i < iters; ++i) run(17u * i + 1u);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms = 0.f;
  cudaEventElapsedTime(&ms, e0, e1);
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  cudaFree(buf);
  if (cudaGetLastError() != cudaSuccess || ms <= 0.f) return GVV_ECUDA;
  *opsPerS = (double)nOps * iters / (ms * 1e-3);
  return GVV_OK;
}
