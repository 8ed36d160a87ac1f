// Micro-benchmark: tcgen05.ld (32x32b.x32) throughput of one SM, alone and while a fifth warp keeps the tensor pipe
// busy with M=128 N=192 K=8 tf32 MMAs into a second accumulator (the ARHMM emission kernel's situation).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I behavenet_b200/csrc -o scripts/ubench/tmem_ld_rate.bin scripts/ubench/tmem_ld_rate.cu
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdint.h>
#include "tc_common.cuh"
using namespace bn_tc;

__device__ __forceinline__ uint64_t desc(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

// LDW = warps that load (1..4), WITH_MMA = 0/1, NLD = x32 loads per wait
template <int NLD>
__global__ void __launch_bounds__(160) k(int iters, int ldw, int with_mma, long long* out) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_ptr;
  __shared__ int stop;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < (128 + 192) * 128 / 4; i += 160) reinterpret_cast<float*>(smem)[i] = 1.0f;
  if (tid == 0) { mbar_init(smem_u32(&bar), 1); stop = 0; asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  if (warp == 4) tmem_alloc<512>(smem_u32(&tmem_ptr));
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = tmem_ptr;
  if (warp == 4) {
    if (with_mma) {
      const uint32_t sa = smem_u32(smem), sb = sa + 16384;
      const uint32_t idesc = make_idesc(128, 192);
      long long t0 = clock64();
      int n = 0;
      while (*(volatile int*)&stop < ldw) {
        for (int q = 0; q < 16; ++q) {
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) umma_tf32_elect(tm + 256, desc(sa + kk * 32), desc(sb + kk * 32), idesc, 1u);
        }
        n += 64;
        umma_commit_elect(smem_u32(&bar));
        mbar_wait(smem_u32(&bar), (n / 64 - 1) & 1);
      }
      long long t1 = clock64();
      if (blockIdx.x == 0 && (tid & 31) == 0) { out[2] = t1 - t0; out[3] = n; }
    }
  } else if (warp < ldw) {
    uint32_t r[32];
    unsigned acc = 0;
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
      for (int j = 0; j < NLD; ++j) {
        tmem_ld32(tm + ((uint32_t)(warp * 32) << 16) + j * 32, r);
        acc += r[0];
      }
      tmem_ld_wait();
    }
    long long t1 = clock64();
    if (blockIdx.x == 0 && tid == 0) { out[0] = t1 - t0; out[1] = acc; }
    __syncwarp();
    if ((tid & 31) == 0) atomicAdd(&stop, 1);
  }
  __syncthreads();
  if (warp == 4) { tc_fence_after(); tmem_dealloc<512>(tm); }
}

int main() {
  long long* d_out;
  cudaMalloc(&d_out, 64);
  const int iters = 2000;
  cudaFuncSetAttribute(k<6>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
  for (int with_mma = 0; with_mma < 2; ++with_mma)
    for (int ldw = 1; ldw <= 4; ldw *= 2) {
      cudaMemset(d_out, 0, 64);
      k<6><<<148, 160, 65536>>>(iters, ldw, with_mma, d_out);
      cudaError_t e = cudaDeviceSynchronize();
      long long h[4];
      cudaMemcpy(h, d_out, 32, cudaMemcpyDeviceToHost);
      double per = (double)h[0] / (iters * 6.0);
      printf("%d loading warp(s), MMAs %s: %6.1f cycles per tcgen05.ld.32x32b.x32 per warp (4 KB) -> %6.1f B/cycle per SM", ldw,
             with_mma ? "running" : "off    ", per, ldw * 4096.0 / per);
      if (with_mma) printf("; MMA N=192: %6.1f cycles each (alone: 96)", (double)h[2] / (double)h[3]);
      printf(" %s\n", e == cudaSuccess ? "" : cudaGetErrorString(e));
    }
  return 0;
}
