// Micro-benchmark: issue rate of tcgen05.mma.kind::tf32 (M=128, K=8) with both operands in shared
// memory, as a function of N, operand layout and CTAs per SM.  Answers the design question of
// is a small-N MMA bound by the tensor pipe (128*N/256 cycles) or by fetching
// its (128 + N) x 32-byte operands from shared memory?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I behavenet_b200/csrc -o umma_rate scripts/ubench/umma_rate.cu
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdint.h>
#include "tc_common.cuh"
using namespace bn_tc;

__device__ __forceinline__ uint64_t desc_sw128(uint32_t saddr, uint32_t sbo) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

template <int N>
__global__ void __launch_bounds__(128) rate_kernel(int iters, int same_acc, long long* out) {
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
    const uint32_t idesc = make_idesc(128, N);
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
      for (int k = 0; k < 4; ++k)
        umma_tf32(tm + (same_acc ? 0 : (k & 1) * N), desc_sw128(sa + k * 32, 1024), desc_sw128(sb + k * 32, 1024), idesc, 1u);
    }
    umma_commit(smem_u32(&bar));
    mbar_wait(smem_u32(&bar), 0);
    long long t1 = clock64();
    if (blockIdx.x == 0) out[0] = t1 - t0;
  }
  __syncthreads();
  if (tid < 32) { tc_fence_after(); tmem_dealloc<256>(tm); }
}

template <int N>
void run(int ctas, long long* d_out) {
  const int iters = 2000;
  size_t smem = 16384 + N * 128;
  cudaFuncSetAttribute(rate_kernel<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  for (int same = 0; same < 2; ++same) {
    rate_kernel<N><<<ctas, 128, smem>>>(iters, same, d_out);
    cudaDeviceSynchronize();
    rate_kernel<N><<<ctas, 128, smem>>>(iters, same, d_out);
    cudaError_t e = cudaDeviceSynchronize();
    long long cyc = 0;
    cudaMemcpy(&cyc, d_out, 8, cudaMemcpyDeviceToHost);
    double per = (double)cyc / (iters * 4.0);
    printf("N=%3d ctas/SM=%d same_acc=%d : %.1f cycles per MMA (tensor floor %.0f, operand bytes %d -> %.1f B/cycle) %s\n", N,
           ctas / 148, same, per, 128.0 * N / 256.0, (128 + N) * 32, (128 + N) * 32 / per, e == cudaSuccess ? "" : cudaGetErrorString(e));
  }
}

int main() {
  long long* d_out;
  cudaMalloc(&d_out, 8);
  for (int ctas : {148, 296}) {
    run<32>(ctas, d_out);
    run<64>(ctas, d_out);
    run<128>(ctas, d_out);
    if (ctas == 148) run<256>(ctas, d_out);
  }
  return 0;
}
