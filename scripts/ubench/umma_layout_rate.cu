// Micro-benchmark: rate of tcgen05.mma.kind::tf32 (M=128, K=8) as a function of the OPERAND LAYOUTS.
// umma_rate.cu measured K-major operands (SWIZZLE_128B): 128 B/cycle of operand fetch per SM.  The
// weight-gradient kernels use MN-major operands (SWIZZLE_128B_BASE32B, the only MN-major layout tf32 has);
// this measures what they cost: A and/or B MN-major, N in {32, 64, 128, 256}.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I behavenet_b200/csrc -o umma_layout_rate scripts/ubench/umma_layout_rate.cu
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdint.h>
#include "tc_common.cuh"
using namespace bn_tc;

__device__ __forceinline__ uint64_t desc_k(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
__device__ __forceinline__ uint64_t desc_mn(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)1 << 61;
  return d;
}

template <int N, int AMN, int BMN>
__global__ void __launch_bounds__(128) rate_kernel(int iters, long long* out) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_ptr;
  const int tid = threadIdx.x;
  for (int i = tid; i < (16384 + N * 128) / 4; i += 128) reinterpret_cast<float*>(smem)[i] = 1.0f;
  if (tid == 0) { mbar_init(smem_u32(&bar), 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  if (tid < 32) tmem_alloc<256>(smem_u32(&tmem_ptr));
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = tmem_ptr;
  if (tid == 0) {
    const uint32_t sa = smem_u32(smem), sb = sa + 16384;
    const uint32_t idesc = make_idesc(128, N) | (AMN ? (1u << 15) : 0u) | (BMN ? (1u << 16) : 0u);
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const uint64_t ad = AMN ? desc_mn(sa + k * 1024, 4096, 512) : desc_k(sa + k * 32);
        const uint64_t bd = BMN ? desc_mn(sb + k * 1024, 4096, 512) : desc_k(sb + k * 32);
        umma_tf32(tm, ad, bd, idesc, 1u);
      }
    }
    umma_commit(smem_u32(&bar));
    mbar_wait(smem_u32(&bar), 0);
    long long t1 = clock64();
    if (blockIdx.x == 0) out[0] = t1 - t0;
  }
  __syncthreads();
  if (tid < 32) { tc_fence_after(); tmem_dealloc<256>(tm); }
}

template <int N, int AMN, int BMN>
void run(long long* d_out) {
  const int iters = 2000;
  size_t smem = 16384 + N * 128;
  cudaFuncSetAttribute(rate_kernel<N, AMN, BMN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  rate_kernel<N, AMN, BMN><<<148, 128, smem>>>(iters, d_out);
  cudaDeviceSynchronize();
  rate_kernel<N, AMN, BMN><<<148, 128, smem>>>(iters, d_out);
  cudaError_t e = cudaDeviceSynchronize();
  long long cyc = 0;
  cudaMemcpy(&cyc, d_out, 8, cudaMemcpyDeviceToHost);
  double per = (double)cyc / (iters * 4.0);
  printf("N=%3d A %s B %s : %6.1f cycles per MMA (tensor floor %3.0f, operand bytes %5d -> %5.1f B/cycle) %s\n", N,
         AMN ? "MN-major" : "K-major ", BMN ? "MN-major" : "K-major ", per, 128.0 * N / 256.0, (128 + N) * 32,
         (128 + N) * 32 / per, e == cudaSuccess ? "" : cudaGetErrorString(e));
}

template <int N>
void run_all(long long* d_out) {
  run<N, 0, 0>(d_out);
  run<N, 0, 1>(d_out);
  run<N, 1, 0>(d_out);
  run<N, 1, 1>(d_out);
}

int main() {
  long long* d_out;
  cudaMalloc(&d_out, 8);
  run_all<32>(d_out);
  run_all<64>(d_out);
  run_all<128>(d_out);
  run_all<256>(d_out);
  return 0;
}
