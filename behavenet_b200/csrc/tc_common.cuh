// tcgen05 / TMEM / mbarrier / cp.async inline-PTX helpers shared by the tensor-core kernels (sm_100a).
#pragma once
#include <stdint.h>

namespace bn_tc {

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
      : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  return ok != 0;
}
// bounded spin: a protocol bug traps (launch failure) instead of hanging the GPU
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  for (uint32_t it = 0; !mbar_try_wait(bar, parity); ++it)
    if (it > (1u << 28)) __trap();
}
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, uint32_t src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

template <int NCOLS>
__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "n"(NCOLS) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
template <int NCOLS>
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(NCOLS) : "memory");
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n .reg .pred p;\n setp.ne.b32 p, %4, 0;\n"
      " tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
// Same instruction issued by ONE elected lane of a converged warp.  With the whole warp running the issue loop the
// operands stay warp-uniform, so ptxas keeps descriptors in uniform registers; issuing from inside `if (lane == 0)`
// makes it wrap every MMA in an ELECT / R2UR.BROADCAST / BRA.U.ANY waterfall (~12 extra instructions per MMA).
__device__ __forceinline__ void umma_tf32_elect(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                                uint32_t accumulate) {
  asm volatile(
      "{\n .reg .pred p, e;\n elect.sync _|e, 0xffffffff;\n setp.ne.b32 p, %4, 0;\n"
      " @e tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit_elect(uint32_t bar) {
  asm volatile(
      "{\n .reg .pred e;\n elect.sync _|e, 0xffffffff;\n"
      " @e tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n}" ::"r"(bar) : "memory");
}
// high / low words of a K-major SWIZZLE_128B descriptor: successive k-steps (32 bytes) add 2 to the low word
__device__ __forceinline__ constexpr uint32_t desc_hi_sw128(uint32_t sbo_bytes) {
  return ((sbo_bytes >> 4) & 0x3FFFu) | (1u << 14) | (2u << 29);
}
__device__ __forceinline__ uint32_t desc_lo_sw128(uint32_t saddr) { return ((saddr >> 4) & 0x3FFFu) | (1u << 16); }
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
// instruction descriptor: D = F32, A = B = TF32, both K-major, dense
__host__ __device__ constexpr uint32_t make_idesc(int M, int N) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// Epilogue of a 32-row x 32-column fp32 accumulator tile held one row per lane (the tcgen05.ld 32x32b
// register layout).  Writing it straight from that layout makes every STG.128 touch 32 different
// 128-byte lines, and applying bias / activation there costs 32 dependent scalar loads per lane.
// Instead the warp transposes the RAW accumulators through a 4 KB shared-memory tile (XOR-swizzled,
// conflict-free both ways); afterwards a lane owns 4 fixed columns of 8 rows, so the bias is one
// float4 per lane, all eight LDS / mask loads are issued before they are consumed, and each store
// instruction writes four whole 128-byte rows.
//   idx   : element offset of this lane's row segment in `out` (and `dact`), or < 0 for an invalid row
//   bias  : 32 per-column biases of this tile (16-byte aligned) or nullptr
//   act   : 0 none, 1 LeakyReLU(leak), 2 sigmoid
//   dact  : optional activation whose sign selects the LeakyReLU derivative (1 or leak) per element
//   colacc : optional per-lane accumulator of the stored values of this lane's 4 columns (bias gradient
//            = column sums of a gradient image, fused here instead of re-reading the image);
//            flush with warp_flush_colsum
// core of the transposing store: ridx[i] = element offset of row (4 i + (lane >> 3)) of this warp's 32-row tile in
// `out` / `dact` (or < 0 for a row that is not stored), as the CALLER knows it -- kernels whose rows are pixels of a
// fixed 16 x 8 block compute it arithmetically; warp_store_rows32 below gathers it from the row-per-lane value
// with shuffles.
__device__ __forceinline__ void warp_store_rows32_at(float* __restrict__ out, const float* __restrict__ dact, float leak,
                                                     const long long (&ridx)[8], const uint32_t (&r)[32],
                                                     const float* __restrict__ bias, int act, float* tile, int lane,
                                                     float4& colacc, bool do_colacc) {
#pragma unroll
  for (int c = 0; c < 8; ++c)
    *reinterpret_cast<uint4*>(tile + lane * 32 + ((c ^ (lane & 7)) << 2)) =
        make_uint4(r[4 * c], r[4 * c + 1], r[4 * c + 2], r[4 * c + 3]);
  __syncwarp();
  const int sub = lane >> 3, chunk = lane & 7;
  const float4 b4 = bias ? __ldg(reinterpret_cast<const float4*>(bias) + chunk) : make_float4(0.f, 0.f, 0.f, 0.f);
  float4 x[8], d[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int row = i * 4 + sub;
    x[i] = *reinterpret_cast<const float4*>(tile + row * 32 + ((chunk ^ (row & 7)) << 2));
  }
  if (dact) {
#pragma unroll
    for (int i = 0; i < 8; ++i)
      d[i] = ridx[i] >= 0 ? __ldg(reinterpret_cast<const float4*>(dact + ridx[i]) + chunk) : make_float4(1.f, 1.f, 1.f, 1.f);
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    float v[4] = {x[i].x + b4.x, x[i].y + b4.y, x[i].z + b4.z, x[i].w + b4.w};
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      if (act == 1) v[e] = v[e] > 0.f ? v[e] : leak * v[e];
      else if (act == 2) v[e] = 1.f / (1.f + expf(-v[e]));
    }
    if (dact) {
      v[0] *= d[i].x > 0.f ? 1.f : leak;
      v[1] *= d[i].y > 0.f ? 1.f : leak;
      v[2] *= d[i].z > 0.f ? 1.f : leak;
      v[3] *= d[i].w > 0.f ? 1.f : leak;
    }
    if (ridx[i] >= 0) {
      *(reinterpret_cast<float4*>(out + ridx[i]) + chunk) = make_float4(v[0], v[1], v[2], v[3]);
      if (do_colacc) { colacc.x += v[0]; colacc.y += v[1]; colacc.z += v[2]; colacc.w += v[3]; }
    }
  }
  __syncwarp();
}

__device__ __forceinline__ void warp_store_rows32(float* __restrict__ out, const float* __restrict__ dact, float leak,
                                                  long long idx, const uint32_t (&r)[32],
                                                  const float* __restrict__ bias, int act, float* tile, int lane,
                                                  float4& colacc, bool do_colacc) {
  const int sub = lane >> 3;
  const int lo = (int)(unsigned long long)idx, hi = (int)((unsigned long long)idx >> 32);
  long long ridx[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int row = i * 4 + sub;
    const unsigned rlo = (unsigned)__shfl_sync(0xffffffffu, lo, row);
    const int rhi = __shfl_sync(0xffffffffu, hi, row);
    ridx[i] = (long long)(((unsigned long long)(unsigned)rhi << 32) | rlo);
  }
  warp_store_rows32_at(out, dact, leak, ridx, r, bias, act, tile, lane, colacc, do_colacc);
}

__device__ __forceinline__ void warp_store_rows32(float* __restrict__ out, const float* __restrict__ dact, float leak,
                                                  long long idx, const uint32_t (&r)[32],
                                                  const float* __restrict__ bias, int act, float* tile, int lane) {
  float4 unused = make_float4(0.f, 0.f, 0.f, 0.f);
  warp_store_rows32(out, dact, leak, idx, r, bias, act, tile, lane, unused, false);
}

// colsum[4 * chunk + e] += sum over the warp's lanes with that chunk (lane & 7) of acc: two shuffle
// steps over the row-group bits, then 8 lanes x 4 atomics per warp
__device__ __forceinline__ void warp_flush_colsum(float* __restrict__ colsum, float4 acc, int lane) {
#pragma unroll
  for (int o = 8; o <= 16; o <<= 1) {
    acc.x += __shfl_xor_sync(0xffffffffu, acc.x, o);
    acc.y += __shfl_xor_sync(0xffffffffu, acc.y, o);
    acc.z += __shfl_xor_sync(0xffffffffu, acc.z, o);
    acc.w += __shfl_xor_sync(0xffffffffu, acc.w, o);
  }
  if (lane < 8) {
    atomicAdd(colsum + 4 * lane + 0, acc.x);
    atomicAdd(colsum + 4 * lane + 1, acc.y);
    atomicAdd(colsum + 4 * lane + 2, acc.z);
    atomicAdd(colsum + 4 * lane + 3, acc.w);
  }
}

// Per-CTA column sums without atomics: every epilogue warp folds the per-lane accumulator of one 32-column
// block into its own row of a shared-memory table ([warps][ncols], zeroed by the warp beforehand); after the
// warps have met, cta_store_colpart writes the CTA's sums as ONE row of a [ctas][C] partial table that the
// batched reduction kernel (cae_simt.cu, job kind 1) adds into the bias gradient.
__device__ __forceinline__ void warp_fold_colsum(float* __restrict__ row32, float4 acc, int lane) {
#pragma unroll
  for (int o = 8; o <= 16; o <<= 1) {
    acc.x += __shfl_xor_sync(0xffffffffu, acc.x, o);
    acc.y += __shfl_xor_sync(0xffffffffu, acc.y, o);
    acc.z += __shfl_xor_sync(0xffffffffu, acc.z, o);
    acc.w += __shfl_xor_sync(0xffffffffu, acc.w, o);
  }
  if (lane < 8) {
    float4* p = reinterpret_cast<float4*>(row32) + lane;
    float4 v = *p;
    v.x += acc.x; v.y += acc.y; v.z += acc.z; v.w += acc.w;
    *p = v;
  }
}

}  // namespace bn_tc
