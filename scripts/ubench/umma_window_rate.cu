// Micro-benchmark: rate of tcgen05.mma.kind::tf32 (M=128, N=64, K=8) when the A operand is a shifted WINDOW of a
// halo buffer, as in the halo kernels: SWIZZLE_128B rows of 128 B, 8-row groups SBO bytes apart (SBO = halo width
// x 128 instead of 1024), start address offset by `row0` rows (not a multiple of the 1024-byte swizzle atom).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I behavenet_b200/csrc -o scripts/ubench/umma_window_rate.bin scripts/ubench/umma_window_rate.cu
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdint.h>
#include "tc_common.cuh"
using namespace bn_tc;

__device__ __forceinline__ uint64_t desc(uint32_t saddr, uint32_t sbo) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

template <int N>
__global__ void __launch_bounds__(128) rate_kernel(int iters, int sbo, int row0, int walk, long long* out) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_ptr;
  const int tid = threadIdx.x;
  for (int i = tid; i < (65536 + N * 128) / 4; i += 128) reinterpret_cast<float*>(smem)[i] = 1.0f;
  if (tid == 0) { mbar_init(smem_u32(&bar), 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  if (tid < 32) tmem_alloc<256>(smem_u32(&tmem_ptr));
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = tmem_ptr;
  if (tid == 0) {
    const uint32_t sa = smem_u32(smem), sb = sa + 65536;
    const uint32_t idesc = make_idesc(128, N);
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
      const uint32_t a0 = sa + (row0 + (walk ? (i % 9) * 3 : 0)) * 128;     // walk: a different window per "tap"
#pragma unroll
      for (int k = 0; k < 4; ++k) umma_tf32(tm, desc(a0 + k * 32, sbo), desc(sb + k * 32, 1024), idesc, 1u);
    }
    umma_commit(smem_u32(&bar));
    mbar_wait(smem_u32(&bar), 0);
    long long t1 = clock64();
    if (blockIdx.x == 0) out[0] = t1 - t0;
  }
  __syncthreads();
  if (tid < 32) { tc_fence_after(); tmem_dealloc<256>(tm); }
}

// same MMAs issued by an elected lane of a CONVERGED warp, descriptors built as (constant high word, low word += 2
// per k-step) so that nothing but two integer adds sits between two MMAs
template <int N>
__global__ void __launch_bounds__(128) rate_kernel_elect(int iters, int sbo, int row0, int walk, long long* out) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_ptr;
  const int tid = threadIdx.x;
  for (int i = tid; i < (65536 + N * 128) / 4; i += 128) reinterpret_cast<float*>(smem)[i] = 1.0f;
  if (tid == 0) { mbar_init(smem_u32(&bar), 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  if (tid < 32) tmem_alloc<256>(smem_u32(&tmem_ptr));
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = tmem_ptr;
  if (tid < 32) {
    const uint32_t sa = smem_u32(smem), sb = sa + 65536;
    const uint32_t idesc = make_idesc(128, N);
    const uint32_t a_hi = ((uint32_t)(sbo >> 4) & 0x3FFFu) | (1u << 14) | (2u << 29);
    const uint32_t b_hi = (1024u >> 4) | (1u << 14) | (2u << 29);
    const uint32_t b_lo = ((sb >> 4) & 0x3FFFu) | (1u << 16);
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
      const uint32_t a_lo = (((sa + (row0 + (walk ? (i % 9) * 3 : 0)) * 128) >> 4) & 0x3FFFu) | (1u << 16);
#pragma unroll
      for (int k = 0; k < 4; ++k)
        umma_tf32_elect(tm, ((uint64_t)a_hi << 32) | (a_lo + 2 * k), ((uint64_t)b_hi << 32) | (b_lo + 2 * k), idesc, 1u);
    }
    umma_commit_elect(smem_u32(&bar));
    mbar_wait(smem_u32(&bar), 0);
    long long t1 = clock64();
    if (blockIdx.x == 0 && tid == 0) out[0] = t1 - t0;
  }
  __syncthreads();
  if (tid < 32) { tc_fence_after(); tmem_dealloc<256>(tm); }
}

template <int N>
void run_elect(long long* d_out, int sbo, int row0, int walk) {
  const int iters = 2000;
  size_t smem = 65536 + N * 128;
  cudaFuncSetAttribute(rate_kernel_elect<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  rate_kernel_elect<N><<<148, 128, smem>>>(iters, sbo, row0, walk, d_out);
  cudaDeviceSynchronize();
  rate_kernel_elect<N><<<148, 128, smem>>>(iters, sbo, row0, walk, d_out);
  cudaError_t e = cudaDeviceSynchronize();
  long long cyc = 0;
  cudaMemcpy(&cyc, d_out, 8, cudaMemcpyDeviceToHost);
  printf("N=%3d SBO=%4d row0=%2d walk=%d elected lane, incremental descriptors : %6.1f cycles per MMA %s\n", N, sbo, row0, walk,
         (double)cyc / (iters * 4.0), e == cudaSuccess ? "" : cudaGetErrorString(e));
}

template <int N>
void run(long long* d_out, int sbo, int row0, int walk) {
  const int iters = 2000;
  size_t smem = 65536 + N * 128;
  cudaFuncSetAttribute(rate_kernel<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  rate_kernel<N><<<148, 128, smem>>>(iters, sbo, row0, walk, d_out);
  cudaDeviceSynchronize();
  rate_kernel<N><<<148, 128, smem>>>(iters, sbo, row0, walk, d_out);
  cudaError_t e = cudaDeviceSynchronize();
  long long cyc = 0;
  cudaMemcpy(&cyc, d_out, 8, cudaMemcpyDeviceToHost);
  printf("N=%3d SBO=%4d row0=%2d walk=%d : %6.1f cycles per MMA %s\n", N, sbo, row0, walk, (double)cyc / (iters * 4.0),
         e == cudaSuccess ? "" : cudaGetErrorString(e));
}

int main() {
  long long* d_out;
  cudaMalloc(&d_out, 8);
  run<64>(d_out, 1024, 0, 0);
  run<64>(d_out, 1280, 0, 0);
  run<64>(d_out, 1280, 1, 0);
  run<64>(d_out, 1280, 11, 0);
  run<64>(d_out, 1280, 11, 1);
  run<64>(d_out, 1024, 1, 0);
  run<64>(d_out, 2048, 0, 0);
  run<128>(d_out, 1024, 0, 0);
  run<128>(d_out, 1280, 11, 1);
  run_elect<64>(d_out, 1024, 0, 0);
  run_elect<64>(d_out, 1280, 11, 1);
  run_elect<32>(d_out, 1280, 11, 1);
  run_elect<128>(d_out, 1280, 11, 1);
  return 0;
}
