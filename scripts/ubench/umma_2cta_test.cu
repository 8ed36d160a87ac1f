// Correctness probe of the CTA-pair MMA (tcgen05.mma.cta_group::2, kind::tf32) with hand-written PTX:
//   D[256 x N] = A[256 x 32] * B[N x 32]^T,  rows 0..127 of A (and of D) live in CTA 0, rows 128..255 in CTA 1,
//   B is SPLIT: CTA 0 holds rows 0..N/2-1, CTA 1 rows N/2..N-1 (each at the same shared-memory offset).
// Checks the semantics the resident-weight kernels rely on: tcgen05.alloc.cta_group::2 issued by the same warp
// of both CTAs, MMA issued by the leader only, tcgen05.commit multicast to both CTAs' mbarriers, each CTA
// reading its own 128 accumulator rows with tcgen05.ld.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I behavenet_b200/csrc -o umma_2cta_test scripts/ubench/umma_2cta_test.cu
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
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
__device__ __forceinline__ uint32_t cta_rank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n barrier.cluster.wait.acquire.aligned;" ::: "memory");
}

template <int N>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128) pair_kernel(const float* __restrict__ A, const float* __restrict__ B,
                                                                             float* __restrict__ D) {
  extern __shared__ __align__(1024) unsigned char smem[];          // A tile 16 KB | B half (N/2 rows x 128 B)
  __shared__ uint64_t done_bar;
  __shared__ uint32_t tmem_ptr;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t rank = cta_rank();
  // A: this CTA's 128 rows; K-major SWIZZLE_128B (row r at r*128, 16-byte chunk c at position c ^ (r & 7))
  for (int i = tid; i < 128 * 8; i += 128) {
    const int r = i >> 3, c = i & 7;
    const float4 v = *reinterpret_cast<const float4*>(A + (size_t)(rank * 128 + r) * 32 + c * 4);
    *reinterpret_cast<float4*>(smem + r * 128 + ((c ^ (r & 7)) << 4)) = v;
  }
  for (int i = tid; i < (N / 2) * 8; i += 128) {
    const int r = i >> 3, c = i & 7;
    const float4 v = *reinterpret_cast<const float4*>(B + (size_t)(rank * (N / 2) + r) * 32 + c * 4);
    *reinterpret_cast<float4*>(smem + 16384 + r * 128 + ((c ^ (r & 7)) << 4)) = v;
  }
  if (tid == 0) {
    mbar_init(smem_u32(&done_bar), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_ptr)), "n"(256) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  cluster_sync();                       // both CTAs' operands and barriers are in place
  tc_fence_after();
  const uint32_t tm = tmem_ptr;
  if (rank == 0 && tid == 0) {
    const uint32_t idesc = make_idesc(256, N);
    const uint32_t sa = smem_u32(smem), sb = sa + 16384;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const uint32_t z = 0;
      asm volatile(
          "{\n .reg .pred p;\n setp.ne.b32 p, %4, 0;\n"
          " tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, {%5, %5, %5, %5, %5, %5, %5, %5}, p;\n}"
          ::"r"(tm), "l"(desc_k(sa + k * 32)), "l"(desc_k(sb + k * 32)), "r"(idesc), "r"(k ? 1u : 0u), "r"(z) : "memory");
    }
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(smem_u32(&done_bar)), "h"((uint16_t)3) : "memory");
  }
  mbar_wait(smem_u32(&done_bar), 0);
  tc_fence_after();
  for (int j = 0; j < N / 32; ++j) {
    uint32_t r[32];
    tmem_ld32(tm + ((uint32_t)(warp * 32) << 16) + j * 32, r);
    tmem_ld_wait();
    float* out = D + (size_t)(rank * 128 + warp * 32 + lane) * N + j * 32;
    for (int q = 0; q < 32; ++q) out[q] = __uint_as_float(r[q]);
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync();
  if (warp == 0) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tm), "n"(256) : "memory");
  }
}

// issue rate of the pair MMA: 2000 x 4 back-to-back MMAs from the leader, all 74 CTA pairs of the chip busy
template <int N>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128) pair_rate_kernel(int iters, long long* out) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ uint64_t done_bar;
  __shared__ uint32_t tmem_ptr;
  const int tid = threadIdx.x, warp = tid >> 5;
  const uint32_t rank = cta_rank();
  for (int i = tid; i < (16384 + (N / 2) * 128) / 4; i += 128) reinterpret_cast<float*>(smem)[i] = 1.0f;
  if (tid == 0) {
    mbar_init(smem_u32(&done_bar), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_ptr)), "n"(256) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  cluster_sync();
  tc_fence_after();
  const uint32_t tm = tmem_ptr;
  if (rank == 0 && tid == 0) {
    const uint32_t idesc = make_idesc(256, N);
    const uint32_t sa = smem_u32(smem), sb = sa + 16384;
    const long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const uint32_t z = 0;
        asm volatile(
            "{\n .reg .pred p;\n setp.ne.b32 p, %4, 0;\n"
            " tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, {%5, %5, %5, %5, %5, %5, %5, %5}, p;\n}"
            ::"r"(tm), "l"(desc_k(sa + k * 32)), "l"(desc_k(sb + k * 32)), "r"(idesc), "r"(1u), "r"(z) : "memory");
      }
    }
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(smem_u32(&done_bar)), "h"((uint16_t)3) : "memory");
    mbar_wait(smem_u32(&done_bar), 0);
    const long long t1 = clock64();
    if (blockIdx.x == 0) out[0] = t1 - t0;
  } else {
    mbar_wait(smem_u32(&done_bar), 0);
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync();
  if (warp == 0) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tm), "n"(256) : "memory");
  }
}

template <int N>
void rate() {
  long long* d_out;
  cudaMalloc(&d_out, 8);
  const size_t smem = 16384 + (N / 2) * 128;
  cudaFuncSetAttribute(pair_rate_kernel<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  const int iters = 2000;
  pair_rate_kernel<N><<<148, 128, smem>>>(iters, d_out);
  cudaDeviceSynchronize();
  pair_rate_kernel<N><<<148, 128, smem>>>(iters, d_out);
  cudaError_t e = cudaDeviceSynchronize();
  long long cyc = 0;
  cudaMemcpy(&cyc, d_out, 8, cudaMemcpyDeviceToHost);
  const double per = (double)cyc / (iters * 4.0);
  printf("N=%3d: cta_group::2 M=256 K=8: %6.1f cycles per MMA (per SM: 128 x %d x 8; tensor floor %3.0f; one-CTA M=128 form measured %s) %s\n",
         N, per, N, 128.0 * N / 256.0, N == 32 ? "44.6" : N == 64 ? "48.0" : N == 128 ? "64.0" : "128.1",
         e == cudaSuccess ? "" : cudaGetErrorString(e));
  cudaFree(d_out);
}

template <int N>
int run() {
  float *hA = (float*)malloc(256 * 32 * 4), *hB = (float*)malloc(N * 32 * 4), *hD = (float*)malloc(256 * N * 4);
  for (int i = 0; i < 256 * 32; ++i) hA[i] = (float)((i * 7 + 3) % 17 - 8) / 16.f;          // TF32-exact values
  for (int i = 0; i < N * 32; ++i) hB[i] = (float)((i * 5 + 1) % 13 - 6) / 8.f;
  float *dA, *dB, *dD;
  cudaMalloc(&dA, 256 * 32 * 4); cudaMalloc(&dB, N * 32 * 4); cudaMalloc(&dD, 256 * N * 4);
  cudaMemcpy(dA, hA, 256 * 32 * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, hB, N * 32 * 4, cudaMemcpyHostToDevice);
  cudaMemset(dD, 0xff, 256 * N * 4);
  const size_t smem = 16384 + (N / 2) * 128;
  cudaFuncSetAttribute(pair_kernel<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  pair_kernel<N><<<2, 128, smem>>>(dA, dB, dD);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("N=%d: CUDA error %s\n", N, cudaGetErrorString(e)); return 1; }
  cudaMemcpy(hD, dD, 256 * N * 4, cudaMemcpyDeviceToHost);
  double maxerr = 0;
  int bad = 0;
  for (int m = 0; m < 256; ++m)
    for (int n = 0; n < N; ++n) {
      double s = 0;
      for (int k = 0; k < 32; ++k) s += (double)hA[m * 32 + k] * hB[n * 32 + k];
      const double err = fabs(s - hD[m * N + n]);
      if (!(err < 1e-4)) ++bad;
      if (err > maxerr || err != err) maxerr = err;
    }
  printf("N=%3d: cta_group::2 M=256 K=32 -> max |err| %.3g, %d of %d entries wrong\n", N, maxerr, bad, 256 * N);
  return bad != 0;
}

int main() {
  int rc = 0;
  rc |= run<32>();
  rc |= run<64>();
  rc |= run<128>();
  rc |= run<256>();
  rate<32>();
  rate<64>();
  rate<128>();
  rate<256>();
  return rc;
}
