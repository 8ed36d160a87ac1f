// Micro-benchmark: rate of tcgen05.mma.kind::tf32 (M=128, K=8, N=192 -- the ARHMM emission GEMM) for K-major operands
// in the four shared-memory layouts: no swizzle (8 x 16 B core matrices, what arhmm_tc.cu used), SWIZZLE_32B,
// SWIZZLE_64B, SWIZZLE_128B.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I behavenet_b200/csrc -o scripts/ubench/umma_nosw_rate.bin scripts/ubench/umma_nosw_rate.cu
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdint.h>
#include "tc_common.cuh"
using namespace bn_tc;

__device__ __forceinline__ uint64_t desc(uint32_t saddr, uint32_t lbo, uint32_t sbo, uint32_t layout) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)layout << 61;
  return d;
}

// MODE 0: no swizzle, K = 40 columns -> 10 k-chunks of 16 B, LBO = rows * 16, SBO = 128
// MODE 1: SWIZZLE_32B  rows of 32 B, SBO = 256, k-step = whole new tile (tile stride rows * 32)
// MODE 2: SWIZZLE_64B  rows of 64 B, SBO = 512, k-step 32 B inside the row
// MODE 3: SWIZZLE_128B rows of 128 B, SBO = 1024, k-step 32 B inside the row
template <int N, int MODE>
__global__ void __launch_bounds__(128) rate_kernel(int iters, long long* out) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_ptr;
  const int tid = threadIdx.x;
  for (int i = tid; i < (128 + N) * 256 / 4; i += 128) reinterpret_cast<float*>(smem)[i] = 1.0f;
  if (tid == 0) { mbar_init(smem_u32(&bar), 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  if (tid < 32) tmem_alloc<256>(smem_u32(&tmem_ptr));
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = tmem_ptr;
  if (tid == 0) {
    const uint32_t sa = smem_u32(smem), sb = sa + 128 * 256;
    const uint32_t idesc = make_idesc(128, N);
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        uint64_t ad, bd;
        if (MODE == 0) { ad = desc(sa + k * 2 * 128 * 16, 128 * 16, 128, 0); bd = desc(sb + k * 2 * N * 16, N * 16, 128, 0); }
        else if (MODE == 1) { ad = desc(sa + k * 128 * 32, 1, 256, 6); bd = desc(sb + k * N * 32, 1, 256, 6); }
        else if (MODE == 2) { ad = desc(sa + (k & 1) * 32 + (k >> 1) * 128 * 64, 1, 512, 4); bd = desc(sb + (k & 1) * 32 + (k >> 1) * N * 64, 1, 512, 4); }
        else { ad = desc(sa + k * 32, 1, 1024, 2); bd = desc(sb + k * 32, 1, 1024, 2); }
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

template <int N, int MODE>
void run(long long* d_out) {
  const int iters = 2000;
  size_t smem = (128 + N) * 256;
  cudaFuncSetAttribute(rate_kernel<N, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  rate_kernel<N, MODE><<<148, 128, smem>>>(iters, d_out);
  cudaDeviceSynchronize();
  rate_kernel<N, MODE><<<148, 128, smem>>>(iters, d_out);
  cudaError_t e = cudaDeviceSynchronize();
  long long cyc = 0;
  cudaMemcpy(&cyc, d_out, 8, cudaMemcpyDeviceToHost);
  double per = (double)cyc / (iters * 4.0);
  const char* names[4] = {"no swizzle  ", "SWIZZLE_32B ", "SWIZZLE_64B ", "SWIZZLE_128B"};
  printf("N=%3d K-major %s : %6.1f cycles per MMA (tensor floor %3.0f, operand bytes %5d -> %5.1f B/cycle) %s\n", N, names[MODE],
         per, 128.0 * N / 256.0, (128 + N) * 32, (128 + N) * 32 / per, e == cudaSuccess ? "" : cudaGetErrorString(e));
}

int main() {
  long long* d_out;
  cudaMalloc(&d_out, 8);
  run<192, 0>(d_out); run<192, 1>(d_out); run<192, 2>(d_out); run<192, 3>(d_out);
  run<64, 0>(d_out); run<64, 1>(d_out); run<64, 2>(d_out); run<64, 3>(d_out);
  return 0;
}
