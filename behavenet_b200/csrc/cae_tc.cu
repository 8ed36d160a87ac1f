// tcgen05 (5th-gen tensor core) implicit-GEMM kernels for the fat CAE layers, TF32 operands with
// fp32 accumulation in TMEM.
//
//   D[128 x BN] (TMEM) += A[128 x 32] (smem, K-major) * B[BN x 32]^T (smem, K-major)     per k-chunk
//
// A is the im2col view of an NHWC activation tensor: row m = output pixel, 32 consecutive k =
// 32 channels of one filter tap.  Four producer warps gather it with 16-byte cp.async copies
// (zero-fill outside the image: this is where ZeroPad2d / the transposed-conv border live) straight
// into the canonical SWIZZLE_128B K-major UMMA layout (what TMA would write), a fifth warp's elected
// lane issues tcgen05.mma and signals stage reuse with tcgen05.commit -> mbarrier; the producer
// warps then read the accumulator back with tcgen05.ld and apply bias / LeakyReLU / sigmoid / the
// activation-derivative mask of the backward pass before storing NHWC.
//
// Weights are pre-packed K-major ([c_out][(tap, c_in)]) and pre-rounded to TF32 (cvt.rna) by
// bn_cae_pack_params; activations are consumed as stored (the tensor core ignores the low 13
// mantissa bits).
#include <cuda.h>

#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <mutex>
#include <vector>

#include "cae_kernels.cuh"
#include "tc_common.cuh"

namespace {

using namespace bn_tc;

constexpr int BM = 128;         // rows per CTA tile = UMMA M
constexpr int BK = 32;          // floats per k-chunk = 128 bytes = 8 core-matrix columns
constexpr int NPROD = 128;      // producer / epilogue threads (warps 0-3)
constexpr int NTHREADS = 160;   // + warp 4: TMEM allocator and MMA issuer


__device__ __forceinline__ uint64_t make_desc_sw128_sbo(uint32_t saddr, uint32_t sbo_bytes);
// K-major SWIZZLE_128B shared-memory matrix descriptor (the layout TMA would write):
//   a tile row is 128 contiguous bytes (32 tf32), 8 rows form a 1024-byte swizzle atom in which the
//   16-byte chunk c of row r is stored at chunk position c ^ (r & 7).  SBO = 1024 bytes between
//   8-row groups; one MMA consumes 32 bytes of K, so successive MMAs advance the start address by 32.
//   Tile bases must be 1024-byte aligned.
__device__ __forceinline__ uint64_t make_desc_sw128(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)1 << 16;                    // LBO (unused for swizzled K-major), canonical value 1
  d |= (uint64_t)(1024 >> 4) << 32;          // SBO
  d |= (uint64_t)1 << 46;                    // descriptor version 1 (sm_100)
  d |= (uint64_t)2 << 61;                    // layout type 2 = SWIZZLE_128B
  return d;
}
// same layout with an explicit stride between 8-row groups (rows of an image tile whose pitch is not 8 pixels)
__device__ __forceinline__ uint64_t make_desc_sw128_sbo(uint32_t saddr, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

struct TcArgs {
  const float* in;
  int Hi, Wi, Ci;
  const float* wt;        // K-major packed weights [Co][KK * Ci]
  int wrow;               // KK * Ci (row length of wt)
  const float* bias;
  float* out;
  int Ho, Wo, Co;
  const float* dact;
  const TapClass* classes;
  int gs, os, n, act;
  int ksplit;             // > 1: single-class op whose k-chunks are split over grid.z
  float* split_out;       // [ksplit][M][Co] raw partial sums (bias / activation applied by the reducer)
  float* colsum;          // optional [Co]: += column sums of the stored output (fused bias gradient; persistent halo kernel)
  float* colpart;         // optional [ctas][Co]: row (blockIdx.z * gridDim.x + blockIdx.x) = this CTA's column sums of
                          // what it stored (no atomics; summed by the batched reduction kernel)
};

template <int BN, int STAGES>
struct TcSmem {
  static constexpr int A_BYTES = BM * BK * 4;
  static constexpr int B_BYTES = BN * BK * 4;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int TOTAL = STAGES * STAGE_BYTES + 1024;   // + barriers / class table / tmem ptr
};

template <int BN, int STAGES>
__global__ void __launch_bounds__(NTHREADS) igemm_tc_kernel(const TcArgs a) {
  bn_pdl_trigger();
  bn_pdl_wait();
  extern __shared__ __align__(1024) unsigned char smem[];
  using S = TcSmem<BN, STAGES>;
  constexpr int NCOLS = BN < 32 ? 32 : BN;
  unsigned char* tail = smem + STAGES * S::STAGE_BYTES;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(tail);              // [STAGES]
  uint64_t* empty_bar = full_bar + STAGES;                             // [STAGES]
  uint64_t* accum_bar = empty_bar + STAGES;                            // [1]
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(accum_bar + 1);     // [1] (+pad)
  TapClass* cls = reinterpret_cast<TapClass*>(tmem_ptr + 2);

  const int tid = threadIdx.x;
  const int warp = tid >> 5;
  {
    const int* src = reinterpret_cast<const int*>(a.classes + (a.ksplit > 1 ? 0 : blockIdx.z));
    int* dst = reinterpret_cast<int*>(cls);
    for (int i = tid; i < (int)(sizeof(TapClass) / 4); i += NTHREADS) dst[i] = src[i];
  }
  if (tid == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(smem_u32(full_bar + s), NPROD);
      mbar_init(smem_u32(empty_bar + s), 1);
    }
    mbar_init(smem_u32(accum_bar), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const int HmWm = cls->Hm * cls->Wm;
  const long long M = (long long)a.n * HmWm;
  const long long m0 = (long long)blockIdx.x * BM;
  if (m0 >= M) {                               // uniform per CTA, before any TMEM allocation
    if (a.colpart != nullptr && a.ksplit <= 1)
      for (int c = tid; c < BN; c += NTHREADS)
        a.colpart[((long long)blockIdx.z * gridDim.x + blockIdx.x) * a.Co + blockIdx.y * BN + c] = 0.f;
    return;
  }
  if (warp == 4) tmem_alloc<NCOLS>(smem_u32(tmem_ptr));
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  const int n0 = blockIdx.y * BN;
  const int Ci = a.Ci;
  const int cpt = Ci / BK;
  // k-chunk range of this CTA (all of them, or one split-K slice)
  int cbeg = 0, nchunks = cls->ntaps * cpt;
  if (a.ksplit > 1) {
    const int per = (nchunks + a.ksplit - 1) / a.ksplit;
    cbeg = blockIdx.z * per;
    const int cend = cbeg + per < nchunks ? cbeg + per : nchunks;
    nchunks = cend > cbeg ? cend - cbeg : 0;
  }
  const uint32_t smem_base = smem_u32(smem);

  if (warp < 4) {
    // ======================= producers ===========================================================
    // Thread t owns the im2col bookkeeping of tile row t.  Copies are issued so that 8 consecutive
    // lanes move the 8 16-byte chunks of ONE 128-byte row (one L1/L2 line per 8 lanes, 4 lines per
    // warp instruction) and land them conflict-free in the swizzled tile; the row pointer is
    // fetched from its owner lane with shuffles.
    const long long m = m0 + tid;
    const bool rvalid = m < M;
    int ybase = 0, xbase = 0;
    long long foff = 0;
    if (rvalid) {
      int f = (int)(m / HmWm);
      int rem = (int)(m - (long long)f * HmWm);
      int ym = rem / cls->Wm;
      int xm = rem - ym * cls->Wm;
      ybase = ym * a.gs;
      xbase = xm * a.gs;
      foff = (long long)f * a.Hi * a.Wi * Ci;
    }
    const int lane = tid & 31;
    const int kc = lane & 7;           // 16-byte chunk of the 128-byte row
    const int rsub = lane >> 3;        // row within the group of 4 rows moved per instruction

    auto issue = [&](int c) {
      const int stage = c % STAGES;
      const int tap = (cbeg + c) / cpt;
      const int c0 = (cbeg + c - tap * cpt) * BK;
      const uint32_t sa = smem_base + stage * S::STAGE_BYTES;
      const uint32_t sb = sa + S::A_BYTES;
      {
        int y = ybase + cls->dy[tap], x = xbase + cls->dx[tap];
        bool ok = rvalid && (unsigned)y < (unsigned)a.Hi && (unsigned)x < (unsigned)a.Wi;
        // own row's source (nullptr = zero fill)
        const float* own = ok ? a.in + foff + ((long long)y * a.Wi + x) * Ci + c0 : nullptr;
        unsigned long long own_u = (unsigned long long)own;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int rl = 4 * i + rsub;                       // row within this warp's 32 rows
          unsigned long long pu = __shfl_sync(0xffffffffu, own_u, rl);
          const int row = warp * 32 + rl;
          const uint32_t dst = sa + row * 128 + ((kc ^ (row & 7)) << 4);
          const float* src = pu ? reinterpret_cast<const float*>(pu) + kc * 4 : a.in;
          cp_async16(dst, src, pu ? 16u : 0u);
        }
      }
      const float* wsrc = a.wt + (long long)cls->wt[tap] * Ci + c0 + kc * 4;
#pragma unroll
      for (int r = warp * 4 + rsub; r < BN; r += 16) {
        const uint32_t dst = sb + r * 128 + ((kc ^ (r & 7)) << 4);
        cp_async16(dst, wsrc + (long long)(n0 + r) * a.wrow, 16u);
      }
    };

    for (int c = 0; c < STAGES - 1; ++c) {
      if (c < nchunks) issue(c);
      cp_async_commit();
    }
    for (int c = 0; c < nchunks; ++c) {
      const int cn = c + STAGES - 1;
      if (cn < nchunks) {
        if (cn >= STAGES) mbar_wait(smem_u32(empty_bar + cn % STAGES), ((cn / STAGES) - 1) & 1);
        issue(cn);
      }
      cp_async_commit();
      cp_async_wait<STAGES - 1>();      // chunk c has landed (this thread's part)
      fence_proxy_async();              // make it visible to the tensor core (async proxy)
      mbar_arrive(smem_u32(full_bar + c % STAGES));
    }

    // ======================= epilogue: TMEM -> registers -> global ============================
    if (nchunks > 0) {
      mbar_wait(smem_u32(accum_bar), 0);
      tc_fence_after();
    }
    long long obase = 0;
    if (rvalid) {
      int f = (int)(m / HmWm);
      int rem = (int)(m - (long long)f * HmWm);
      int ym = rem / cls->Wm;
      int xm = rem - ym * cls->Wm;
      int oy = cls->oy0 + a.os * ym, ox = cls->ox0 + a.os * xm;
      obase = (((long long)f * a.Ho + oy) * a.Wo + ox) * a.Co + n0;
    }
    // all MMAs have retired, so the pipeline stages are free: reuse them as the transposition tiles
    float* tile = reinterpret_cast<float*>(smem) + warp * 1024;
    float* colred = reinterpret_cast<float*>(smem) + 4 * 1024;      // [4 warps][BN] per-CTA column sums
    const int elane = tid & 31;
    const bool want_col = a.colpart != nullptr && a.ksplit <= 1;
    if (want_col)
      for (int c = elane; c < BN; c += 32) colred[warp * BN + c] = 0.f;
#pragma unroll 1
    for (int j = 0; j < BN / 32; ++j) {
      uint32_t r[32];
      if (nchunks > 0) {
        tmem_ld32(tmem_base + ((uint32_t)(warp * 32) << 16) + j * 32, r);
        tmem_ld_wait();
      } else {
#pragma unroll
        for (int q = 0; q < 32; ++q) r[q] = 0u;
      }
      if (a.ksplit > 1) {
        const long long pidx = rvalid ? ((long long)blockIdx.z * M + m) * a.Co + n0 + j * 32 : -1;
        warp_store_rows32(a.split_out, nullptr, BN_LEAK, pidx, r, nullptr, BN_ACT_NONE, tile, elane);
      } else {
        float4 cacc = make_float4(0.f, 0.f, 0.f, 0.f);
        warp_store_rows32(a.out, a.dact, BN_LEAK, rvalid ? obase + j * 32 : -1, r,
                          a.bias ? a.bias + n0 + j * 32 : nullptr, a.act, tile, elane, cacc, want_col);
        if (want_col) warp_fold_colsum(colred + warp * BN + j * 32, cacc, elane);
      }
    }
    if (want_col) {
      asm volatile("bar.sync 1, 128;" ::: "memory");               // the four epilogue warps
      float* dst = a.colpart + ((long long)blockIdx.z * gridDim.x + blockIdx.x) * a.Co + n0;
      for (int c = tid; c < BN; c += NPROD)
        dst[c] = colred[c] + colred[BN + c] + colred[2 * BN + c] + colred[3 * BN + c];
    }
    tc_fence_before();
  } else {
    // ======================= MMA issuer (warp 4, one elected lane) =============================
    if ((tid & 31) == 0) {
      constexpr uint32_t idesc = make_idesc(BM, BN);
      for (int c = 0; c < nchunks; ++c) {
        const int stage = c % STAGES;
        mbar_wait(smem_u32(full_bar + stage), (c / STAGES) & 1);
        tc_fence_after();
        const uint32_t sa = smem_base + stage * S::STAGE_BYTES;
        const uint32_t sb = sa + S::A_BYTES;
#pragma unroll
        for (int k = 0; k < BK / 8; ++k) {
          uint64_t ad = make_desc_sw128(sa + k * 32);
          uint64_t bd = make_desc_sw128(sb + k * 32);
          umma_tf32(tmem_base, ad, bd, idesc, (c | k) != 0 ? 1u : 0u);
        }
        umma_commit(smem_u32(empty_bar + stage));     // stage reusable once these MMAs retire
      }
      if (nchunks > 0) umma_commit(smem_u32(accum_bar));
    }
    __syncwarp();
  }
  __syncthreads();
  if (warp == 4) {
    tc_fence_after();
    tmem_dealloc<NCOLS>(tmem_base);
  }
}

template <int BN, int STAGES>
int launch_tc(const TcArgs& a, int nclasses, int maxM, cudaStream_t st) {
  using S = TcSmem<BN, STAGES>;
  auto kern = igemm_tc_kernel<BN, STAGES>;
  static bool configured = false;
  if (!configured) {
    BN_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, S::TOTAL));
    configured = true;
  }
  dim3 grid(bn_cdiv((long long)a.n * maxM, BM), a.Co / BN, a.ksplit > 1 ? a.ksplit : nclasses);
  BN_CUDA(bn_launch(kern, grid, NTHREADS, S::TOTAL, st, a));
  BN_LAUNCHED();
  return 0;
}


// ------------------------------------------------------------------------------------------------
// TMA-fed variant of the implicit GEMM (fprop form and dgrad form with <= 4 residue classes).
// The whole producer side collapses to ONE thread issuing two bulk-tensor copies per k-chunk:
//   * A: cp.async.bulk.tensor.4d ... im2col -- the TMA unit walks 128 consecutive output pixels of
//     the NHWC activation (wrapping over rows and frames inside the bounding box that encodes the
//     padding), fetches 32 channels at the filter-tap offset, zero-fills out-of-image taps and
//     writes the SWIZZLE_128B tile the MMA descriptors expect;
//   * B: a plain 2-D tile of the K-major weights.
// Completion is signalled on the stage's mbarrier by transaction bytes (complete_tx).
// ------------------------------------------------------------------------------------------------
struct alignas(64) TmaSet {
  CUtensorMap a[4];     // one im2col map per residue class
  CUtensorMap b;        // weights [Co][KK*Ci]
  int lw[4], lh[4];     // bounding-box lower corner per class (tensor coordinates of base pixel 0)
};

__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_im2col_4d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c, int w,
                                              int h, int n, uint16_t offw, uint16_t offh) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.im2col.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5, %6}], [%2], {%7, %8};"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c), "r"(w), "r"(h), "r"(n), "h"(offw), "h"(offh)
      : "memory");
}
__device__ __forceinline__ void tma_tile_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int x, int y) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(x), "r"(y)
      : "memory");
}

// MT = M-tiles (128 output pixels each) per CTA.  The kernel is bound by L2 -> SM bandwidth (measured:
// l1tex__m_xbar2l1tex_read_bytes / time = 10.7-14 TB/s on every im2col kernel), and every 128-pixel tile
// re-streams the layer's whole weight slice (205 KB .. 3.3 MB) through that path; with MT = 2 one weight tile
// feeds two accumulators, which halves the weight traffic per output pixel.
template <int BN, int STAGES, int MT>
struct TmaSmem {
  static constexpr int A_BYTES = MT * BM * BK * 4;
  static constexpr int B_BYTES = BN * BK * 4;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int TOTAL = STAGES * STAGE_BYTES + 1024;
};

template <int BN, int STAGES, int MT>
__global__ void __launch_bounds__(NTHREADS) igemm_tma_kernel(const __grid_constant__ TmaSet tm, const TcArgs a) {
  bn_pdl_trigger();
  extern __shared__ __align__(1024) unsigned char smem[];
  using S = TmaSmem<BN, STAGES, MT>;
  constexpr int NCOLS = MT * BN < 32 ? 32 : MT * BN;
  static_assert(NCOLS <= 512 && (NCOLS & (NCOLS - 1)) == 0, "TMEM columns");
  unsigned char* tail = smem + STAGES * S::STAGE_BYTES;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(tail);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* accum_bar = empty_bar + STAGES;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(accum_bar + 1);
  TapClass* cls = reinterpret_cast<TapClass*>(tmem_ptr + 2);

  const int tid = threadIdx.x;
  const int warp = tid >> 5;
  const int cls_idx = a.ksplit > 1 ? 0 : blockIdx.z;
  {
    const int* src = reinterpret_cast<const int*>(a.classes + cls_idx);
    int* dst = reinterpret_cast<int*>(cls);
    for (int i = tid; i < (int)(sizeof(TapClass) / 4); i += NTHREADS) dst[i] = src[i];
  }
  if (tid == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(smem_u32(full_bar + s), 1);
      mbar_init(smem_u32(empty_bar + s), 1);
    }
    mbar_init(smem_u32(accum_bar), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const int HmWm = cls->Hm * cls->Wm;
  const long long M = (long long)a.n * HmWm;
  const long long m0 = (long long)blockIdx.x * (MT * BM);
  if (m0 >= M) {
    if (a.colpart != nullptr && a.ksplit <= 1) {
      bn_pdl_wait();
      for (int c = tid; c < BN; c += NTHREADS)
        a.colpart[((long long)blockIdx.z * gridDim.x + blockIdx.x) * a.Co + blockIdx.y * BN + c] = 0.f;
    }
    return;
  }
  if (warp == 4) tmem_alloc<NCOLS>(smem_u32(tmem_ptr));
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  bn_pdl_wait();     // set-up above touched only shared memory, TMEM and constant plan tables

  const int n0 = blockIdx.y * BN;
  const int Ci = a.Ci;
  const int cpt = Ci / BK;
  int cbeg = 0, nchunks = cls->ntaps * cpt;
  if (a.ksplit > 1) {
    const int per = (nchunks + a.ksplit - 1) / a.ksplit;
    cbeg = blockIdx.z * per;
    const int cend = cbeg + per < nchunks ? cbeg + per : nchunks;
    nchunks = cend > cbeg ? cend - cbeg : 0;
  }
  const uint32_t smem_base = smem_u32(smem);
  // number of M-tiles of this CTA that hold at least one output pixel
  const int mt_live = (M - m0 + BM - 1) / BM < MT ? (int)((M - m0 + BM - 1) / BM) : MT;

  if (warp < 4) {
    if (warp == 0) {
      // ======================= TMA producer (warp 0: lanes 0 .. MT-1 the im2col tiles, lane MT the weights) ====
      const int lw = tm.lw[cls_idx], lh = tm.lh[cls_idx];
      int w0 = 0, h0 = 0, f0 = 0;
      if (tid < MT) {
        const long long mm = m0 + (long long)tid * BM;
        f0 = (int)(mm / HmWm);
        const int rem0 = (int)(mm - (long long)f0 * HmWm);
        const int ym0 = rem0 / cls->Wm, xm0 = rem0 - ym0 * cls->Wm;
        w0 = lw + xm0 * a.gs;
        h0 = lh + ym0 * a.gs;
      }
      const uint32_t tx_bytes = (uint32_t)(mt_live * BM * BK * 4 + S::B_BYTES);
      for (int c = 0; c < nchunks; ++c) {
        const int stage = c % STAGES;
        if (c >= STAGES) mbar_wait(smem_u32(empty_bar + stage), ((c / STAGES) - 1) & 1);
        const int tap = (cbeg + c) / cpt;
        const int c0 = (cbeg + c - tap * cpt) * BK;
        const uint32_t sa = smem_base + stage * S::STAGE_BYTES;
        const uint32_t sb = sa + S::A_BYTES;
        const uint32_t bar = smem_u32(full_bar + stage);
        if (tid == 0) mbar_expect_tx(bar, tx_bytes);
        __syncwarp();
        if (tid < mt_live)
          tma_im2col_4d(sa + tid * (BM * BK * 4), &tm.a[cls_idx], bar, c0, w0, h0, f0, (uint16_t)(cls->dx[tap] - lw),
                        (uint16_t)(cls->dy[tap] - lh));
        else if (tid == MT)
          tma_tile_2d(sb, &tm.b, bar, cls->wt[tap] * Ci + c0, n0);
      }
    }
    __syncwarp();
    // ======================= epilogue ============================================================
    if (nchunks > 0) {
      mbar_wait(smem_u32(accum_bar), 0);
      tc_fence_after();
    }
    // all MMAs have retired, so the pipeline stages are free: reuse them as the transposition tiles
    float* tile = reinterpret_cast<float*>(smem) + warp * 1024;
    float* colred = reinterpret_cast<float*>(smem) + 4 * 1024;      // [4 warps][BN] per-CTA column sums
    const int elane = tid & 31;
    const bool want_col = a.colpart != nullptr && a.ksplit <= 1;
    if (want_col)
      for (int c = elane; c < BN; c += 32) colred[warp * BN + c] = 0.f;
#pragma unroll 1
    for (int mt = 0; mt < mt_live; ++mt) {
      const long long m = m0 + (long long)mt * BM + tid;
      const bool rvalid = m < M;
      long long obase = 0;
      if (rvalid) {
        int f = (int)(m / HmWm);
        int rem = (int)(m - (long long)f * HmWm);
        int ym = rem / cls->Wm;
        int xm = rem - ym * cls->Wm;
        int oy = cls->oy0 + a.os * ym, ox = cls->ox0 + a.os * xm;
        obase = (((long long)f * a.Ho + oy) * a.Wo + ox) * a.Co + n0;
      }
#pragma unroll 1
      for (int j = 0; j < BN / 32; ++j) {
        uint32_t r[32];
        if (nchunks > 0) {
          tmem_ld32(tmem_base + ((uint32_t)(warp * 32) << 16) + mt * BN + j * 32, r);
          tmem_ld_wait();
        } else {
#pragma unroll
          for (int q = 0; q < 32; ++q) r[q] = 0u;
        }
        if (a.ksplit > 1) {
          const long long pidx = rvalid ? ((long long)blockIdx.z * M + m) * a.Co + n0 + j * 32 : -1;
          warp_store_rows32(a.split_out, nullptr, BN_LEAK, pidx, r, nullptr, BN_ACT_NONE, tile, elane);
        } else {
          float4 cacc = make_float4(0.f, 0.f, 0.f, 0.f);
          warp_store_rows32(a.out, a.dact, BN_LEAK, rvalid ? obase + j * 32 : -1, r,
                            a.bias ? a.bias + n0 + j * 32 : nullptr, a.act, tile, elane, cacc, want_col);
          if (want_col) warp_fold_colsum(colred + warp * BN + j * 32, cacc, elane);
        }
      }
    }
    if (want_col) {
      asm volatile("bar.sync 1, 128;" ::: "memory");               // the four epilogue warps
      float* dst = a.colpart + ((long long)blockIdx.z * gridDim.x + blockIdx.x) * a.Co + n0;
      for (int c = tid; c < BN; c += NPROD)
        dst[c] = colred[c] + colred[BN + c] + colred[2 * BN + c] + colred[3 * BN + c];
    }
    tc_fence_before();
  } else {
    {
      // the whole warp runs the issue loop converged and one elected lane issues: descriptors stay cheap (constant
      // high word, low word += 2 per k-step) instead of ~20 instructions of address arithmetic + ELECT / R2UR
      // waterfall per MMA, which at >= 66 cycles per MMA out-lasted every MMA with N <= 128 (umma_window_rate.cu)
      constexpr uint32_t idesc = make_idesc(BM, BN);
      constexpr uint32_t d_hi = desc_hi_sw128(1024);
      uint32_t accflag = 0u;
      for (int c = 0; c < nchunks; ++c) {
        const int stage = c % STAGES;
        mbar_wait(smem_u32(full_bar + stage), (c / STAGES) & 1);
        tc_fence_after();
        const uint32_t sa = smem_base + stage * S::STAGE_BYTES;
        const uint32_t a_lo = desc_lo_sw128(sa), b_lo = desc_lo_sw128(sa + S::A_BYTES);
#pragma unroll
        for (int k = 0; k < BK / 8; ++k) {
#pragma unroll
          for (int mt = 0; mt < MT; ++mt)
            if (mt < mt_live)
              umma_tf32_elect(tmem_base + mt * BN, ((uint64_t)d_hi << 32) | (a_lo + mt * (BM * BK * 4 / 16) + 2 * k),
                              ((uint64_t)d_hi << 32) | (b_lo + 2 * k), idesc, accflag);
          accflag = 1u;
        }
        umma_commit_elect(smem_u32(empty_bar + stage));
      }
      if (nchunks > 0) umma_commit_elect(smem_u32(accum_bar));
    }
    __syncwarp();
  }
  __syncthreads();
  if (warp == 4) {
    tc_fence_after();
    tmem_dealloc<NCOLS>(tmem_base);
  }
}

// ---- host side: tensor-map encoding through the driver entry points (no link-time libcuda) ------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
typedef CUresult (*EncodeIm2colFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                   const cuuint64_t*, const int*, const int*, cuuint32_t, cuuint32_t,
                                   const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                   CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn g_encode_tiled = nullptr;
EncodeIm2colFn g_encode_im2col = nullptr;
int g_tma_state = 0;      // 0 unknown, 1 available, -1 unavailable
// BN_TMA=0 in the environment selects the cp.async producers instead (A/B tests)
std::mutex g_tma_mutex;

bool tma_available() {
  std::lock_guard<std::mutex> lk(g_tma_mutex);
  if (g_tma_state == 0) {
    const char* env = getenv("BN_TMA");
    if (env && env[0] == '0') {
      g_tma_state = -1;
      return false;
    }
    void* f1 = nullptr;
    void* f2 = nullptr;
    cudaDriverEntryPointQueryResult q1, q2;
    cudaError_t e1 = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f1, cudaEnableDefault, &q1);
    cudaError_t e2 = cudaGetDriverEntryPoint("cuTensorMapEncodeIm2col", &f2, cudaEnableDefault, &q2);
    if (e1 == cudaSuccess && e2 == cudaSuccess && f1 && f2 && q1 == cudaDriverEntryPointSuccess &&
        q2 == cudaDriverEntryPointSuccess) {
      g_encode_tiled = (EncodeTiledFn)f1;
      g_encode_im2col = (EncodeIm2colFn)f2;
      g_tma_state = 1;
    } else {
      g_tma_state = -1;
      cudaGetLastError();
    }
  }
  return g_tma_state == 1;
}

// im2col map over an NHWC fp32 image: `count_w x count_h` base pixels per frame starting at
// (lw, lh) with traversal stride `ts`; box = 32 channels x `pixels` base pixels
bool encode_im2col(CUtensorMap* map, const float* p, int N, int H, int W, int C, int lw, int lh, int count_w,
                   int count_h, int ts, int pixels, CUtensorMapSwizzle swz) {
  cuuint64_t gdim[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
  cuuint64_t gstr[3] = {(cuuint64_t)C * 4, (cuuint64_t)W * C * 4, (cuuint64_t)H * W * C * 4};
  // number of base pixels along w: (W + upper - lower - 1) / ts + 1 == count_w
  int lower[2] = {lw, lh};
  int upper[2] = {(count_w - 1) * ts + lw + 1 - W, (count_h - 1) * ts + lh + 1 - H};
  for (int i = 0; i < 2; ++i)
    if (lower[i] < -128 || lower[i] > 127 || upper[i] < -128 || upper[i] > 127) return false;
  cuuint32_t estr[4] = {1, (cuuint32_t)ts, (cuuint32_t)ts, 1};
  CUresult r = g_encode_im2col(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, (void*)p, gdim, gstr, lower, upper, 32,
                               (cuuint32_t)pixels, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swz,
                               CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS;
}

bool encode_tiled_2d(CUtensorMap* map, const float* p, long long rows, long long cols, int box_cols, int box_rows,
                     CUtensorMapSwizzle swz) {
  cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t gstr[1] = {(cuuint64_t)cols * 4};
  cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = g_encode_tiled(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)p, gdim, gstr, box, estr,
                              CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS;
}

template <int BN, int STAGES, int MT>
int launch_tma(const TmaSet& tm, const TcArgs& a, int nclasses, int maxM, cudaStream_t st) {
  using S = TmaSmem<BN, STAGES, MT>;
  static_assert(S::TOTAL <= 227 * 1024, "shared memory per CTA");
  auto kern = igemm_tma_kernel<BN, STAGES, MT>;
  static bool configured = false;
  if (!configured) {
    BN_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, S::TOTAL));
    configured = true;
  }
  dim3 grid(bn_cdiv((long long)a.n * maxM, MT * BM), a.Co / BN, a.ksplit > 1 ? a.ksplit : nclasses);
  BN_CUDA(bn_launch(kern, grid, NTHREADS, S::TOTAL, st, tm, a));
  BN_LAUNCHED();
  return 0;
}

// out[i] = act(bias[i % Co] + sum_z part[z][i]) * lrelu'(dact[i])
__global__ void splitk_reduce_kernel(const float* __restrict__ part, int splits, long long total, int Co,
                                     const float* __restrict__ bias, const float* __restrict__ dact, int act,
                                     float* __restrict__ out) {
  bn_pdl_trigger();
  bn_pdl_wait();
  long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (i >= total) return;
  float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int z = 0; z < splits; ++z) {
    float4 v = __ldg(reinterpret_cast<const float4*>(part + (long long)z * total + i));
    s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
  }
  float v[4] = {s.x, s.y, s.z, s.w};
  const int c = (int)(i % Co);
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    float x = v[e] + (bias ? __ldg(bias + c + e) : 0.f);
    if (act == BN_ACT_LEAKY) x = x > 0.f ? x : BN_LEAK * x;
    else if (act == BN_ACT_SIGMOID) x = 1.f / (1.f + expf(-x));
    if (dact) x *= __ldg(dact + i + e) > 0.f ? 1.f : BN_LEAK;
    v[e] = x;
  }
  *reinterpret_cast<float4*>(out + i) = make_float4(v[0], v[1], v[2], v[3]);
}


// ------------------------------------------------------------------------------------------------
// Weight gradient on tensor cores.
//   D[kk (128 rows), cs (BN)] += sum over 32-pixel chunks of  A(m, kk) * S(m, cs)
// with A(m, kk=(tap, cb)) = big[pix(m) + off(tap), cb] (im2col of the layer's big image) and
// S = the small image's gradient / activation.  The reduction index m is the GEMM K dimension and
// both operands are contiguous along their M/N index in memory (channels), so both are staged in
// the MN-major canonical layout that tf32 supports, SWIZZLE_128B_BASE32B: an atom is 4 k-rows x
// 128 bytes (32 channels) whose 32-byte granule g of k-row r is stored at granule g ^ (r & 3);
// atoms adjacent along M/N are LBO apart, adjacent along K are SBO apart.  One tcgen05.mma
// (K = 8) consumes two k-atoms.
// Each CTA reduces a slice of the pixels (split-K over grid.z) into partial[z][kk][cs]; the
// fp32 reduce kernel of cae_simt.cu sums the slices into the torch-layout gradient.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint64_t make_desc_mn_sw128(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)1 << 61;                    // layout type 1 = SWIZZLE_128B_BASE32B
  return d;
}
__host__ __device__ constexpr uint32_t make_idesc_mn(int M, int N) {
  return make_idesc(M, N) | (1u << 15) | (1u << 16);      // A and B MN-major
}

struct WgTcArgs {
  const float* big;
  int Hb, Wb, Cb;
  const float* small;
  int Hs, Ws, Cs;
  const TapClass* cls;     // fprop class (all k*k taps)
  int gs, n, Ktot;
  long long rows_per_split;
  float* partial;
};

template <int BN, int STAGES>
__global__ void __launch_bounds__(NTHREADS) wgrad_tc_kernel(const WgTcArgs a) {
  bn_pdl_trigger();
  bn_pdl_wait();
  extern __shared__ __align__(1024) unsigned char smem[];
  using S = TcSmem<BN, STAGES>;
  constexpr int NCOLS = BN < 32 ? 32 : BN;
  unsigned char* tail = smem + STAGES * S::STAGE_BYTES;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(tail);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* accum_bar = empty_bar + STAGES;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(accum_bar + 1);
  TapClass* cls = reinterpret_cast<TapClass*>(tmem_ptr + 2);

  const int tid = threadIdx.x;
  const int warp = tid >> 5;
  {
    const int* src = reinterpret_cast<const int*>(a.cls);
    int* dst = reinterpret_cast<int*>(cls);
    for (int i = tid; i < (int)(sizeof(TapClass) / 4); i += NTHREADS) dst[i] = src[i];
  }
  if (tid == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(smem_u32(full_bar + s), NPROD);
      mbar_init(smem_u32(empty_bar + s), 1);
    }
    mbar_init(smem_u32(accum_bar), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 4) tmem_alloc<NCOLS>(smem_u32(tmem_ptr));
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  const int HsWs = a.Hs * a.Ws;
  const long long M = (long long)a.n * HsWs;
  const long long mbeg = (long long)blockIdx.z * a.rows_per_split;
  const long long mend = mbeg + a.rows_per_split < M ? mbeg + a.rows_per_split : M;
  const int nchunks = mbeg < mend ? (int)((mend - mbeg + BK - 1) / BK) : 0;
  const int kk0 = blockIdx.x * BM;
  const int n0 = blockIdx.y * BN;
  const int Cb = a.Cb, Cs = a.Cs;
  const uint32_t smem_base = smem_u32(smem);
  constexpr uint32_t A_LBO = 512, A_SBO = (BM / 32) * 512;     // 4-row k-atoms of 512 bytes
  constexpr uint32_t B_LBO = 512, B_SBO = (BN / 32) * 512;

  if (warp < 4) {
    // producers.  Warp w moves the 32-channel slab kk0 + 32w .. +31 of the A tile: a fixed
    // (tap, channel offset) for the whole kernel; lane l decodes pixel l of each 32-pixel chunk.
    const int lane = tid & 31;
    const int kc = lane & 7;
    const int rsub = lane >> 3;
    const int akk = kk0 + 32 * warp;
    const bool kvalid = akk < a.Ktot;
    const int tap = kvalid ? akk / Cb : 0;
    const int cb0 = akk - tap * Cb;
    const int dy = cls->dy[tap], dx = cls->dx[tap];

    auto issue = [&](int c) {
      const int stage = c % STAGES;
      const uint32_t sa = smem_base + stage * S::STAGE_BYTES;
      const uint32_t sb = sa + S::A_BYTES;
      const long long mc = mbeg + (long long)c * BK;
      // lane l: source of pixel mc + l for this warp's channel slab
      unsigned long long own_u = 0;
      {
        long long m = mc + lane;
        if (kvalid && m < mend) {
          int f = (int)(m / HsWs);
          int rem = (int)(m - (long long)f * HsWs);
          int ym = rem / a.Ws;
          int xm = rem - ym * a.Ws;
          int y = ym * a.gs + dy, x = xm * a.gs + dx;
          if ((unsigned)y < (unsigned)a.Hb && (unsigned)x < (unsigned)a.Wb)
            own_u = (unsigned long long)(a.big + (((long long)f * a.Hb + y) * a.Wb + x) * Cb + cb0);
        }
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int px = 4 * i + rsub;                            // pixel (k row) within the chunk
        unsigned long long pu = __shfl_sync(0xffffffffu, own_u, px);
        const uint32_t dst = sa + (px >> 2) * A_SBO + warp * A_LBO + (px & 3) * 128 +
                             ((((kc >> 1) ^ (px & 3)) << 5) | ((kc & 1) << 4));
        const float* src = pu ? reinterpret_cast<const float*>(pu) + kc * 4 : a.big;
        cp_async16(dst, src, pu ? 16u : 0u);
      }
      // S tile: rows (pixel, 32-channel slab j)
#pragma unroll
      for (int q = warp * 4 + rsub; q < 32 * (BN / 32); q += 16) {
        const int px = q & 31, j = q >> 5;
        const long long m = mc + px;
        const bool ok = m < mend;
        const uint32_t dst = sb + (px >> 2) * B_SBO + j * B_LBO + (px & 3) * 128 +
                             ((((kc >> 1) ^ (px & 3)) << 5) | ((kc & 1) << 4));
        const float* src = ok ? a.small + m * Cs + n0 + 32 * j + kc * 4 : a.small;
        cp_async16(dst, src, ok ? 16u : 0u);
      }
    };

    for (int c = 0; c < STAGES - 1; ++c) {
      if (c < nchunks) issue(c);
      cp_async_commit();
    }
    for (int c = 0; c < nchunks; ++c) {
      const int cn = c + STAGES - 1;
      if (cn < nchunks) {
        if (cn >= STAGES) mbar_wait(smem_u32(empty_bar + cn % STAGES), ((cn / STAGES) - 1) & 1);
        issue(cn);
      }
      cp_async_commit();
      cp_async_wait<STAGES - 1>();
      fence_proxy_async();
      mbar_arrive(smem_u32(full_bar + c % STAGES));
    }
    if (nchunks > 0) {
      mbar_wait(smem_u32(accum_bar), 0);
      tc_fence_after();
    }
    // epilogue: row kk of the partial slice, 32 columns at a time
    const int kk = kk0 + tid;
    const long long prow_idx = ((long long)blockIdx.z * a.Ktot + kk) * Cs + n0;
    float* tile = reinterpret_cast<float*>(smem) + warp * 1024;      // pipeline stages are idle now
#pragma unroll 1
    for (int j = 0; j < BN / 32; ++j) {
      uint32_t r[32];
      if (nchunks > 0) {
        tmem_ld32(tmem_base + ((uint32_t)(warp * 32) << 16) + j * 32, r);
        tmem_ld_wait();
      } else {
#pragma unroll
        for (int q = 0; q < 32; ++q) r[q] = 0u;
      }
      warp_store_rows32(a.partial, nullptr, BN_LEAK, kk < a.Ktot ? prow_idx + j * 32 : -1, r, nullptr, BN_ACT_NONE, tile,
                        tid & 31);
    }
    tc_fence_before();
  } else {
    if ((tid & 31) == 0) {
      constexpr uint32_t idesc = make_idesc_mn(BM, BN);
      for (int c = 0; c < nchunks; ++c) {
        const int stage = c % STAGES;
        mbar_wait(smem_u32(full_bar + stage), (c / STAGES) & 1);
        tc_fence_after();
        const uint32_t sa = smem_base + stage * S::STAGE_BYTES;
        const uint32_t sb = sa + S::A_BYTES;
#pragma unroll
        for (int k = 0; k < BK / 8; ++k) {
          uint64_t ad = make_desc_mn_sw128(sa + k * 2 * A_SBO, A_LBO, A_SBO);
          uint64_t bd = make_desc_mn_sw128(sb + k * 2 * B_SBO, B_LBO, B_SBO);
          umma_tf32(tmem_base, ad, bd, idesc, (c | k) != 0 ? 1u : 0u);
        }
        umma_commit(smem_u32(empty_bar + stage));
      }
      if (nchunks > 0) umma_commit(smem_u32(accum_bar));
    }
    __syncwarp();
  }
  __syncthreads();
  if (warp == 4) {
    tc_fence_after();
    tmem_dealloc<NCOLS>(tmem_base);
  }
}

// TMA-fed weight gradient: one thread issues, per 32-pixel chunk, one im2col copy per 32-channel
// slab of the A tile (each slab = one filter tap x 32 channels) and one 2-D copy per 32-channel slab
// of the small image; both land as 32 rows x 128 bytes with the 32-byte-atom 128B swizzle, i.e. the
// MN-major canonical layout with LBO = 4096 (next slab) and SBO = 512 (next 4 pixels).
struct alignas(64) WgTmaSet {
  CUtensorMap a;      // im2col over the big image, 32 pixels x 32 channels per copy
  CUtensorMap s;      // small image as a [M][Cs] matrix, 32 rows x 32 channels per copy
  int lw, lh;
};

// MT = 128-row tiles of the (tap, channel) axis per CTA: the small-image tile (the B operand) of a pixel chunk is
// fetched once and multiplied into MT accumulators, which divides its L2 -> SM traffic by MT (this kernel, too,
// runs at the L2 -> SM bandwidth: 1.31 GB in 122 us for encoder conv1).
template <int BN, int STAGES, int MT>
__global__ void __launch_bounds__(NTHREADS) wgrad_tma_kernel(const __grid_constant__ WgTmaSet tm, const WgTcArgs a) {
  bn_pdl_trigger();
  extern __shared__ __align__(1024) unsigned char smem[];
  using S = TmaSmem<BN, STAGES, MT>;
  constexpr int NCOLS = MT * BN < 32 ? 32 : MT * BN;
  static_assert(NCOLS <= 512 && (NCOLS & (NCOLS - 1)) == 0, "TMEM columns");
  unsigned char* tail = smem + STAGES * S::STAGE_BYTES;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(tail);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* accum_bar = empty_bar + STAGES;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(accum_bar + 1);
  TapClass* cls = reinterpret_cast<TapClass*>(tmem_ptr + 2);
  const int tid = threadIdx.x;
  const int warp = tid >> 5;
  {
    const int* src = reinterpret_cast<const int*>(a.cls);
    int* dst = reinterpret_cast<int*>(cls);
    for (int i = tid; i < (int)(sizeof(TapClass) / 4); i += NTHREADS) dst[i] = src[i];
  }
  if (tid == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(smem_u32(full_bar + s), 1);
      mbar_init(smem_u32(empty_bar + s), 1);
    }
    mbar_init(smem_u32(accum_bar), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 4) tmem_alloc<NCOLS>(smem_u32(tmem_ptr));
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  bn_pdl_wait();     // set-up above touched only shared memory, TMEM and constant plan tables

  const int HsWs = a.Hs * a.Ws;
  const long long M = (long long)a.n * HsWs;
  const long long mbeg = (long long)blockIdx.z * a.rows_per_split;
  const long long mend = mbeg + a.rows_per_split < M ? mbeg + a.rows_per_split : M;
  const int nchunks = mbeg < mend ? (int)((mend - mbeg + BK - 1) / BK) : 0;
  const int kk0 = blockIdx.x * (MT * BM);
  const int n0 = blockIdx.y * BN;
  const int Cb = a.Cb, Cs = a.Cs;
  const uint32_t smem_base = smem_u32(smem);
  constexpr uint32_t SLAB = 32 * 128;                 // 32 pixels x 32 channels
  constexpr uint32_t LBO = SLAB, SBO = 512;
  // valid 32-channel slabs of this CTA's MT tiles (the last tile of Ktot = 25*Cb may be short) and live tiles
  int nslab = (a.Ktot - kk0 + 31) / 32;
  nslab = nslab > MT * (BM / 32) ? MT * (BM / 32) : nslab;
  const int mt_live = (nslab + BM / 32 - 1) / (BM / 32);

  if (warp < 4) {
    if (warp == 0) {
      // Producer warp.  A single thread issuing the copies of a chunk back to back (each with its own coordinate
      // arithmetic) was the bottleneck of this kernel, so the copies of a chunk are issued by different lanes:
      // lane j < nslab the im2col slab j, the next BN/32 lanes the tiles of the small image.  Lane 0 posts the
      // transaction count first.
      const uint32_t bytes = (uint32_t)(nslab * SLAB + (BN / 32) * SLAB);
      const int lane = tid;
      int my_cb0 = 0;
      uint16_t my_dx = 0, my_dy = 0;
      if (lane < nslab) {
        const int akk = kk0 + 32 * lane;
        const int tap = akk / Cb;
        my_cb0 = akk - tap * Cb;
        my_dx = (uint16_t)(cls->dx[tap] - tm.lw);
        my_dy = (uint16_t)(cls->dy[tap] - tm.lh);
      }
      for (int c = 0; c < nchunks; ++c) {
        const int stage = c % STAGES;
        if (c >= STAGES) mbar_wait(smem_u32(empty_bar + stage), ((c / STAGES) - 1) & 1);
        const uint32_t sa = smem_base + stage * S::STAGE_BYTES;
        const uint32_t sb = sa + S::A_BYTES;
        const uint32_t bar = smem_u32(full_bar + stage);
        const long long mc = mbeg + (long long)c * BK;
        if (lane == 0) mbar_expect_tx(bar, bytes);
        __syncwarp();
        if (lane < nslab) {
          const int f = (int)(mc / HsWs);
          const int rem = (int)(mc - (long long)f * HsWs);
          const int ym = rem / a.Ws, xm = rem - ym * a.Ws;
          const int w0 = tm.lw + xm * a.gs, h0 = tm.lh + ym * a.gs;
          tma_im2col_4d(sa + lane * SLAB, &tm.a, bar, my_cb0, w0, h0, f, my_dx, my_dy);
        } else if (lane < nslab + BN / 32) {
          const int j = lane - nslab;
          tma_tile_2d(sb + j * SLAB, &tm.s, bar, n0 + 32 * j, (int)mc);
        }
      }
    }
    __syncwarp();
    if (nchunks > 0) {
      mbar_wait(smem_u32(accum_bar), 0);
      tc_fence_after();
    }
    float* tile = reinterpret_cast<float*>(smem) + warp * 1024;      // pipeline stages are idle now
#pragma unroll 1
    for (int mt = 0; mt < mt_live; ++mt) {
      const int kk = kk0 + mt * BM + tid;
      const long long prow_idx = ((long long)blockIdx.z * a.Ktot + kk) * Cs + n0;
#pragma unroll 1
      for (int j = 0; j < BN / 32; ++j) {
        uint32_t r[32];
        if (nchunks > 0) {
          tmem_ld32(tmem_base + ((uint32_t)(warp * 32) << 16) + mt * BN + j * 32, r);
          tmem_ld_wait();
        } else {
#pragma unroll
          for (int q = 0; q < 32; ++q) r[q] = 0u;
        }
        warp_store_rows32(a.partial, nullptr, BN_LEAK, kk < a.Ktot ? prow_idx + j * 32 : -1, r, nullptr, BN_ACT_NONE, tile,
                          tid & 31);
      }
    }
    tc_fence_before();
  } else {
    {
      // converged warp, one elected lane issues (see igemm_tma_kernel)
      constexpr uint32_t idesc = make_idesc_mn(BM, BN);
      constexpr uint32_t d_hi = (((uint32_t)SBO >> 4) & 0x3FFFu) | (1u << 14) | (1u << 29);      // SWIZZLE_128B_BASE32B
      constexpr uint32_t lbo_w = (((uint32_t)LBO >> 4) & 0x3FFFu) << 16;
      uint32_t accflag = 0u;
      for (int c = 0; c < nchunks; ++c) {
        const int stage = c % STAGES;
        mbar_wait(smem_u32(full_bar + stage), (c / STAGES) & 1);
        tc_fence_after();
        const uint32_t sa = smem_base + stage * S::STAGE_BYTES;
        const uint32_t a_lo = ((sa >> 4) & 0x3FFFu) | lbo_w, b_lo = (((sa + S::A_BYTES) >> 4) & 0x3FFFu) | lbo_w;
#pragma unroll
        for (int k = 0; k < BK / 8; ++k) {
#pragma unroll
          for (int mt = 0; mt < MT; ++mt)
            if (mt < mt_live)
              umma_tf32_elect(tmem_base + mt * BN, ((uint64_t)d_hi << 32) | (a_lo + (mt * (BM / 32) * SLAB + k * 2 * SBO) / 16),
                              ((uint64_t)d_hi << 32) | (b_lo + (k * 2 * SBO) / 16), idesc, accflag);
          accflag = 1u;
        }
        umma_commit_elect(smem_u32(empty_bar + stage));
      }
      if (nchunks > 0) umma_commit_elect(smem_u32(accum_bar));
    }
    __syncwarp();
  }
  __syncthreads();
  if (warp == 4) {
    tc_fence_after();
    tmem_dealloc<NCOLS>(tmem_base);
  }
}

template <int BN, int STAGES, int MT>
int launch_wgrad_tma(const WgTmaSet& tm, const WgTcArgs& a, int splits, cudaStream_t st) {
  using S = TmaSmem<BN, STAGES, MT>;
  static_assert(S::TOTAL <= 227 * 1024, "shared memory per CTA");
  auto kern = wgrad_tma_kernel<BN, STAGES, MT>;
  static bool configured = false;
  if (!configured) {
    BN_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, S::TOTAL));
    configured = true;
  }
  dim3 grid(bn_cdiv(a.Ktot, MT * BM), a.Cs / BN, splits);
  BN_CUDA(bn_launch(kern, grid, NTHREADS, S::TOTAL, st, tm, a));
  BN_LAUNCHED();
  return 0;
}

template <int BN, int STAGES>
int launch_wgrad_tc(const WgTcArgs& a, int splits, cudaStream_t st) {
  using S = TcSmem<BN, STAGES>;
  auto kern = wgrad_tc_kernel<BN, STAGES>;
  static bool configured = false;
  if (!configured) {
    BN_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, S::TOTAL));
    configured = true;
  }
  dim3 grid(bn_cdiv(a.Ktot, BM), a.Cs / BN, splits);
  BN_CUDA(bn_launch(kern, grid, NTHREADS, S::TOTAL, st, a));
  BN_LAUNCHED();
  return 0;
}

}  // namespace

namespace {

struct TmaKey {
  const void* in;
  const void* wt;
  const void* cls;
  int n, H, W, C, gs, nclasses, Co, wrow;
  bool operator==(const TmaKey& o) const {
    return in == o.in && wt == o.wt && cls == o.cls && n == o.n && H == o.H && W == o.W && C == o.C &&
           gs == o.gs && nclasses == o.nclasses && Co == o.Co && wrow == o.wrow;
  }
};
std::vector<std::pair<TmaKey, TmaSet>> g_tma_cache;

// Tensor maps of one op (cached: activations live at stable workspace addresses across steps)
const TmaSet* get_tma_set(const ImgView& in, const float* wt, int wrow, int Co, int bn, const TapClass* d_classes,
                          const TapClass* h_classes, int nclasses, int gs, int n) {
  if (nclasses > 4 || !tma_available()) return nullptr;
  TmaKey key{in.p, wt, d_classes, n, in.H, in.W, in.C, gs, nclasses, Co, wrow};
  std::lock_guard<std::mutex> lk(g_tma_mutex);
  for (auto& kv : g_tma_cache)
    if (kv.first == key) return &kv.second;
  TmaSet tm;
  memset(&tm, 0, sizeof(tm));
  for (int c = 0; c < nclasses; ++c) {
    const TapClass& k = h_classes[c];
    if (k.ntaps == 0 || k.Hm == 0 || k.Wm == 0) continue;
    int lw = 127, lh = 127;
    for (int t = 0; t < k.ntaps; ++t) {
      lw = k.dx[t] < lw ? k.dx[t] : lw;
      lh = k.dy[t] < lh ? k.dy[t] : lh;
    }
    tm.lw[c] = lw; tm.lh[c] = lh;
    if (!encode_im2col(&tm.a[c], in.p, n, in.H, in.W, in.C, lw, lh, k.Wm, k.Hm, gs, BM, CU_TENSOR_MAP_SWIZZLE_128B))
      return nullptr;
  }
  if (!encode_tiled_2d(&tm.b, wt, Co, wrow, BK, bn, CU_TENSOR_MAP_SWIZZLE_128B)) return nullptr;
  if (g_tma_cache.size() >= 256) g_tma_cache.clear();
  g_tma_cache.emplace_back(key, tm);
  return &g_tma_cache.back().second;
}

// ------------------------------------------------------------------------------------------------
// Halo-resident transposed convolution (dgrad form, stride 2: ConvTranspose2d forward and Conv2d
// backward-data of the fat layers).
//
// The im2col kernels above are bound by L2 -> SM bandwidth, not by the tensor pipe: every small-image
// pixel is fetched once per filter tap (25x for the four stride-residue classes together).  Here a CTA
// owns a 16 x 8 block of small-image pixels of one frame and ALL FOUR residue classes of the 32 x 16
// output pixels they produce.  The (16+2)-row x 16-pixel input halo of a 32-channel chunk is staged
// ONCE by a single tiled TMA copy (zero fill outside the image = the transposed-conv border) as
// 128-byte pixel rows in the SWIZZLE_128B K-major layout, so the A operand of filter tap (dy, dx) is
// the same buffer at row offset (dy - lo_y) * 16 + (dx - lo_x): 8 consecutive pixels of an image row
// are one swizzle atom, the next tile row is SBO = 16 rows = 2048 bytes further.  No data moves per
// tap; 25 taps x 4 k-steps of tcgen05.mma read the halo in place into four TMEM accumulators (one
// per class).  Only the weights stream: one TMA tile (NB x 32 floats) per (tap, chunk).
// A-side traffic drops from 25 fetches per small pixel to 2.25.
//
// The shifted window starts on a 128-byte row that is not a multiple of 8 rows: the descriptor's
// base-offset field carries (row & 7) so that the 128B-swizzle phase matches what TMA wrote.
// ------------------------------------------------------------------------------------------------
constexpr int HALO_W = 10, HALO_H = 18;                 // box: (8 + 2) pixels x (16 + 2) rows
constexpr int HALO_ABYTES = HALO_W * HALO_H * 128;      // one 32-channel chunk: 23040 bytes
constexpr int HALO_ASTRIDE = (HALO_ABYTES + 1023) / 1024 * 1024;
constexpr int HALO_THREADS = 192;                       // warps 0-3 epilogue, 4 MMA, 5 TMA producer
constexpr int HALO_MAXG = 12;

__device__ __forceinline__ void tma_tile_4d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c, int w, int h,
                                            int n) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c), "r"(w), "r"(h), "r"(n)
      : "memory");
}

// ---- thread-block-cluster helpers: weight tiles are identical for every tile of a layer, so the
// CTAs of a cluster each fetch 1/csz of them and TMA-multicast into all members' shared memory ----
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tma_tile_2d_mc(uint32_t dst, const CUtensorMap* map, uint32_t bar, int x, int y,
                                               uint16_t mask) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster"
      " [%0], [%1, {%3, %4}], [%2], %5;"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(x), "r"(y), "h"(mask)
      : "memory");
}
__device__ __forceinline__ void umma_commit_mc_elect(uint32_t bar, uint16_t mask) {
  asm volatile(
      "{\n .reg .pred e;\n elect.sync _|e, 0xffffffff;\n"
      " @e tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;\n}"
      ::"r"(bar), "h"(mask) : "memory");
}
__device__ __forceinline__ void umma_commit_mc(uint32_t bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(bar), "h"(mask) : "memory");
}

// per-role phase timestamps of CTA 0 (debug aid, BN_HALO_DBG=1; read back with bn_debug_halo_times)
__device__ long long g_halo_dbg[8 * 8];
__device__ __forceinline__ long long dbg_clock() {
  long long v;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(v));
  return v;
}
#define HALO_STAMP(tile, slot)                                                                      \
  do {                                                                                              \
    if (h.dbg && blockIdx.x == 0 && (tile) < 8) g_halo_dbg[(tile) * 8 + (slot)] = dbg_clock();      \
  } while (0)

// One MMA group = one halo window (tap offset) shared by `ncls` residue classes whose accumulators
// sit in adjacent TMEM column blocks col0 .. col0 + ncls - 1: a single tcgen05.mma of N = ncls * NB
// against the stacked weight tiles of those classes.  Small-N MMAs are bound by the 4 KB A-operand
// fetch, not by the tensor pipe, so sharing the window between classes is what makes them efficient.
struct HaloGroup {
  unsigned short row_off;    // window offset in halo rows (pixels)
  unsigned char col0, ncls;
  unsigned char wt[4];       // weight tap index per stacked class
};

struct HaloArgs {
  TcArgs a;
  int lo_y, lo_x;            // smallest tap offset over all classes
  int tiles_x;               // 8-pixel column blocks per frame
  int ngroups;
  int tiles_per_frame;
  long long total_tiles;     // tiles_per_frame * frames
  int csz;                   // cluster size (1, 2 or 4): CTAs that share the weight-tile stream
  int dbg;
  int pos[4];                // TMEM column block of class c
  int Hm[4], Wm[4], oy0[4], ox0[4];
  HaloGroup g[HALO_MAXG];
};

struct alignas(64) HaloMaps {
  CUtensorMap a;       // NHWC input, box 32 ch x 10 px x 18 rows x 1 frame
  CUtensorMap b;       // K-major weights, box 32 x NB
};

template <int NB, int NST, bool PERSIST>
struct HaloSmem {
  static constexpr int B_TILE = NB * 128;
  static constexpr int B_BYTES = 4 * B_TILE;                          // up to four stacked class tiles
  static constexpr int OFF_A = NST * B_BYTES;                         // multiple of 1024
  // 4 warps x 4 KB transposition tiles: their own region when the weight ring keeps streaming under
  // the epilogue (persistent), aliased onto the (then idle) ring when a CTA owns a single tile
  static constexpr int OFF_EPI = PERSIST ? OFF_A + 2 * HALO_ASTRIDE : 0;
  static constexpr int OFF_BAR = OFF_A + 2 * HALO_ASTRIDE + (PERSIST ? 4 * 4096 : 0);
  static constexpr int TOTAL = OFF_BAR + 8 * (2 * NST + 8) + 32;
};

// PERSIST: a CTA walks tiles blockIdx.x, +gridDim.x, .. (tile = 16 x 8 small pixels of one frame).
// Three roles run the same tile sequence decoupled by mbarriers: the TMA thread streams halo chunks
// (2 buffers) and weight-tile groups (NST-deep ring) without ever draining between tiles, the MMA
// thread accumulates tile i into TMEM buffer i & 1, and the four epilogue warps drain buffer i & 1
// while the MMAs of tile i + 1 fill the other one.  Barrier / TMEM setup is paid once per CTA.
// !PERSIST (NB = 64: two accumulator sets would take all 512 TMEM columns and leave one CTA per SM,
// which measured slower): one tile per CTA, one accumulator set, two CTAs per SM overlap each other.
template <int NB, int NST, bool PERSIST>
__global__ void __launch_bounds__(HALO_THREADS) dgrad_halo_kernel(const __grid_constant__ HaloMaps maps,
                                                                  const __grid_constant__ HaloArgs h) {
  bn_pdl_trigger();
  extern __shared__ __align__(1024) unsigned char smem[];
  using S = HaloSmem<NB, NST, PERSIST>;
  constexpr int ACC_COLS = 4 * NB;                   // one accumulator per class
  constexpr int NCOLS = (PERSIST ? 2 : 1) * ACC_COLS;
  const TcArgs& a = h.a;
  uint64_t* b_full = reinterpret_cast<uint64_t*>(smem + S::OFF_BAR);
  uint64_t* b_empty = b_full + NST;
  uint64_t* a_full = b_empty + NST;       // [2]
  uint64_t* a_empty = a_full + 2;         // [2]
  uint64_t* acc_full = a_empty + 2;       // [2]
  uint64_t* acc_empty = acc_full + 2;     // [2]
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(acc_empty + 2);

  const int tid = threadIdx.x;
  const int warp = tid >> 5;
  if (tid == 0) {
    for (int s = 0; s < NST; ++s) {
      mbar_init(smem_u32(b_full + s), 1);
      mbar_init(smem_u32(b_empty + s), h.csz);      // released by the MMAs of every CTA of the cluster
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(smem_u32(a_full + s), 1);
      mbar_init(smem_u32(a_empty + s), 1);
      mbar_init(smem_u32(acc_full + s), 1);
      mbar_init(smem_u32(acc_empty + s), 4);        // one arrival per epilogue warp
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 4) tmem_alloc<NCOLS>(smem_u32(tmem_ptr));
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const int csz = h.csz;
  const uint32_t crank = csz > 1 ? cluster_ctarank() : 0;
  const uint16_t cmask = (uint16_t)((1u << csz) - 1u);
  if (csz > 1) cluster_sync_all();      // every member's barriers exist before any remote arrive / copy
  const uint32_t tmem_base = *tmem_ptr;
  bn_pdl_wait();     // set-up above touched only shared memory, TMEM and constant plan tables
  const int Ci = a.Ci;
  const int nchunk = Ci / BK;
  const uint32_t smem_base = smem_u32(smem);
  const long long total = h.total_tiles;
  // tile sequence: clusters take groups of csz consecutive tiles, member r the r-th of each group, so
  // all members run the same number of iterations (total_tiles is a multiple of csz)
  const long long t_first = (long long)(blockIdx.x / csz) * csz + crank;
  const long long t_step = (long long)(gridDim.x / csz) * csz;

  if (warp < 4) {
    // ======================= epilogue ============================================================
    float* tile = reinterpret_cast<float*>(smem + S::OFF_EPI) + warp * 1024;
    const int elane = tid & 31;
    // fused bias gradient (persistent variant only, where NB = 32: one flush per CTA; a flush per tile costs
    // more in same-address atomics than the separate column-sum pass): this lane's 4 columns
    static_assert(!PERSIST || NB == 32, "the fused column sums keep one accumulator per lane");
    float4 cacc = make_float4(0.f, 0.f, 0.f, 0.f);
    // one-tile-per-CTA variant: per-CTA column sums go to row blockIdx.x of the partial table (no atomics)
    float* colred = reinterpret_cast<float*>(smem + S::OFF_EPI) + 4 * 1024;      // [4 warps][NB]
    const bool want_part = !PERSIST && a.colpart != nullptr;
    int ti = 0;
    for (long long T = t_first; T < total; T += t_step, ++ti) {
      const int f = (int)(T / h.tiles_per_frame);
      const int blk = (int)(T - (long long)f * h.tiles_per_frame);
      const int by = blk / h.tiles_x, bx = blk - by * h.tiles_x;
      const int ym = by * 16 + (tid >> 3), xm = bx * 8 + (tid & 7);
      const int buf = PERSIST ? (ti & 1) : 0;
      if (tid == 0) HALO_STAMP(ti, 3);
      mbar_wait(smem_u32(acc_full + buf), (ti >> 1) & 1);
      tc_fence_after();
      if (tid == 0) HALO_STAMP(ti, 4);
      if (want_part)      // (the table aliases the weight ring: only touch it once every MMA has retired)
        for (int c = elane; c < NB; c += 32) colred[warp * NB + c] = 0.f;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const bool rvalid = ym < h.Hm[c] && xm < h.Wm[c];
        const int oy = h.oy0[c] + a.os * ym, ox = h.ox0[c] + a.os * xm;
        const long long obase = (((long long)f * a.Ho + oy) * a.Wo + ox) * a.Co;
#pragma unroll 1
        for (int j = 0; j < NB / 32; ++j) {
          uint32_t r[32];
          if (!(h.dbg & 8)) {
            tmem_ld32(tmem_base + ((uint32_t)(warp * 32) << 16) + buf * ACC_COLS + h.pos[c] * NB + j * 32, r);
            tmem_ld_wait();
          } else {
#pragma unroll
            for (int q = 0; q < 32; ++q) r[q] = 0u;
          }
          if (want_part) {
            float4 pacc = make_float4(0.f, 0.f, 0.f, 0.f);
            warp_store_rows32(a.out, a.dact, BN_LEAK, (rvalid && !(h.dbg & 4)) ? obase + j * 32 : -1, r,
                              a.bias ? a.bias + j * 32 : nullptr, a.act, tile, elane, pacc, true);
            warp_fold_colsum(colred + warp * NB + j * 32, pacc, elane);
          } else {
            warp_store_rows32(a.out, a.dact, BN_LEAK, (rvalid && !(h.dbg & 4)) ? obase + j * 32 : -1, r,
                              a.bias ? a.bias + j * 32 : nullptr, a.act, tile, elane, cacc, PERSIST && a.colsum != nullptr);
          }
        }
      }
      tc_fence_before();                       // this thread's TMEM reads of the buffer are complete
      __syncwarp();
      if (elane == 0) mbar_arrive(smem_u32(acc_empty + buf));
      if (tid == 0) HALO_STAMP(ti, 5);
    }
    if (PERSIST && a.colsum) warp_flush_colsum(a.colsum, cacc, elane);
    if (want_part) {
      asm volatile("bar.sync 1, 128;" ::: "memory");               // the four epilogue warps
      float* dst = a.colpart + (long long)blockIdx.x * a.Co;
      for (int c = tid; c < NB; c += 128)
        dst[c] = colred[c] + colred[NB + c] + colred[2 * NB + c] + colred[3 * NB + c];
    }
  } else if (warp == 4) {
    // ======================= MMA issuer ==========================================================
    {
      // whole warp converged, one elected lane issues (see igemm_tma_kernel): constant high descriptor words, the low
      // word of a window = halo base + row_off * 8, += 2 per k-step
      constexpr uint32_t a_hi = desc_hi_sw128(HALO_W * 128), b_hi = desc_hi_sw128(1024);
      const bool lane0 = (tid & 31) == 0;
      int ai = 0, bi = 0, ti = 0;
      for (long long T = t_first; T < total; T += t_step, ++ti) {
        const int buf = PERSIST ? (ti & 1) : 0;
        if (lane0) HALO_STAMP(ti, 0);
        if (ti >= 2) {
          mbar_wait(smem_u32(acc_empty + buf), ((ti >> 1) - 1) & 1);
          tc_fence_after();
        }
        if (lane0) HALO_STAMP(ti, 1);
        uint32_t accflag = 0u;
        for (int c = 0; c < nchunk; ++c, ++ai) {
          const int slot = ai & 1;
          mbar_wait(smem_u32(a_full + slot), (ai >> 1) & 1);
          tc_fence_after();
          const uint32_t a_lo0 = desc_lo_sw128(smem_base + S::OFF_A + slot * HALO_ASTRIDE);
          for (int gi = 0; gi < h.ngroups; ++gi, ++bi) {
            const int stage = bi % NST;
            mbar_wait(smem_u32(b_full + stage), (bi / NST) & 1);
            tc_fence_after();
            const uint32_t b_lo = desc_lo_sw128(smem_base + stage * S::B_BYTES);
            const HaloGroup g = h.g[gi];
            const uint32_t idesc = make_idesc(BM, g.ncls * NB);
            const uint32_t acc = tmem_base + buf * ACC_COLS + g.col0 * NB;
            // the 128B swizzle is a function of the shared-memory ADDRESS bits, so a window that
            // starts on any 128-byte row of the TMA-written halo needs no descriptor base offset
            const uint32_t a_lo = a_lo0 + (uint32_t)g.row_off * 8u;
            if (!(h.dbg & 2)) {
#pragma unroll
              for (int k = 0; k < BK / 8; ++k) {
                umma_tf32_elect(acc, ((uint64_t)a_hi << 32) | (a_lo + 2 * k), ((uint64_t)b_hi << 32) | (b_lo + 2 * k), idesc, accflag);
                accflag = 1u;
              }
            }
            if (csz > 1) umma_commit_mc_elect(smem_u32(b_empty + stage), cmask);
            else umma_commit_elect(smem_u32(b_empty + stage));
          }
          umma_commit_elect(smem_u32(a_empty + slot));
        }
        umma_commit_elect(smem_u32(acc_full + buf));
        if (lane0) HALO_STAMP(ti, 2);
      }
    }
    __syncwarp();
  } else {
    // ======================= TMA producer: halo chunks and weight-tile groups ====================
    // whole warp runs the loop; lane 0 requests the halos, lane j < ncls the j-th weight tile of a group
    {
      const int lane = tid & 31;
      int ai = 0, bi = 0, bt = 0;
      // chunk sequence number q -> (tile, chunk); the halo of chunk q + 1 is requested before the
      // weight tiles of chunk q so that it lands while chunk q is being multiplied
      auto load_halo = [&](long long T, int c, int q) {
        const int slot = q & 1;
        if (q >= 2) mbar_wait(smem_u32(a_empty + slot), ((q >> 1) - 1) & 1);
        if (lane == 0) {
          const int f = (int)(T / h.tiles_per_frame);
          const int blk = (int)(T - (long long)f * h.tiles_per_frame);
          const int by = blk / h.tiles_x, bx = blk - by * h.tiles_x;
          const uint32_t bar = smem_u32(a_full + slot);
          mbar_expect_tx(bar, (uint32_t)HALO_ABYTES);
          tma_tile_4d(smem_base + S::OFF_A + slot * HALO_ASTRIDE, &maps.a, bar, c * BK, bx * 8 + h.lo_x, by * 16 + h.lo_y, f);
        }
      };
      if (t_first < total) load_halo(t_first, 0, 0);
      for (long long T = t_first; T < total; T += t_step) {
        for (int c = 0; c < nchunk; ++c, ++ai) {
          // next chunk in sequence
          if (c + 1 < nchunk) load_halo(T, c + 1, ai + 1);
          else if (T + t_step < total) load_halo(T + t_step, 0, ai + 1);
          for (int gi = 0; gi < h.ngroups; ++gi, ++bi) {
            const int stage = bi % NST;
            if (bi >= NST) mbar_wait(smem_u32(b_empty + stage), ((bi / NST) - 1) & 1);
            const uint32_t bar = smem_u32(b_full + stage);
            const HaloGroup g = h.g[gi];
            if (lane == 0) mbar_expect_tx(bar, (uint32_t)(g.ncls * S::B_TILE));
            __syncwarp();
            if (lane < g.ncls) {
              const uint32_t dst = smem_base + stage * S::B_BYTES + lane * S::B_TILE;
              if (csz == 1) tma_tile_2d(dst, &maps.b, bar, g.wt[lane] * Ci + c * BK, 0);
              else if (((bt + lane) % csz) == (int)crank) tma_tile_2d_mc(dst, &maps.b, bar, g.wt[lane] * Ci + c * BK, 0, cmask);
            }
            bt += g.ncls;
          }
        }
      }
    }
    __syncwarp();
  }
  tc_fence_before();
  __syncthreads();
  if (csz > 1) cluster_sync_all();      // no member leaves while peers may still multicast into it
  if (warp == 4) {
    tc_fence_after();
    tmem_dealloc<NCOLS>(tmem_base);
  }
}

// ------------------------------------------------------------------------------------------------
// CTA-pair form of the halo kernel for C_out = 32 (ConvTranspose2d forward into / Conv2d backward-data out of
// the 32-channel 64 x 64 maps): two CTAs of a cluster (the two SMs of a TPC) each own a 16 x 8 tile and issue
// ONE tcgen05.mma.cta_group::2 per window: M = 256 (128 rows per CTA), N = ncls * 32, with the weight operand
// SPLIT between the two CTAs (each holds N / 2 rows at the same shared-memory offset).
//
// Why: dbg timestamps show the one-CTA kernel waiting for data, not for the tensor pipe -- 7.2 us per tile go by
// with every MMA switched off -- and 200 of the 246 KB a tile pulls through L2 are the layer's WEIGHTS (205 KB in
// total), re-streamed by each of the 2048 tiles.  Split over a CTA pair the whole layer's weights are 102 KB per
// SM: they are loaded ONCE per CTA and stay resident, the ring of weight tiles disappears, and what is left to
// stream is the 46 KB of input halo per tile.
//
// Roles per CTA: warps 0-7 epilogue (own 128 accumulator rows, two classes each), warp 8 MMA issue (leader CTA only), warp 9 TMA.
// Cross-CTA signalling: both CTAs' TMA copies complete on the LEADER's "full" barriers (cp.async.bulk.tensor
// .cta_group::2 with the peer bit of the mbarrier address cleared), the leader's tcgen05.commit multicasts to both
// CTAs' "empty" / "accumulator ready" barriers, and the peer's epilogue warps release an accumulator with a remote
// mbarrier.arrive on the leader's barrier.
// ------------------------------------------------------------------------------------------------
constexpr uint32_t PEER_BIT_MASK = 0xFEFFFFFFu;        // shared::cluster address of the even CTA of a pair

__device__ __forceinline__ void tma_tile_4d_pair(uint32_t dst, const CUtensorMap* map, uint32_t leader_bar, int c, int w, int h,
                                                 int n) {
  asm volatile(
      "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(leader_bar), "r"(c), "r"(w), "r"(h), "r"(n)
      : "memory");
}
__device__ __forceinline__ void tma_tile_2d_pair(uint32_t dst, const CUtensorMap* map, uint32_t leader_bar, int x, int y) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(leader_bar), "r"(x), "r"(y)
      : "memory");
}
__device__ __forceinline__ void umma_tf32_pair(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  const uint32_t z = 0;
  asm volatile(
      "{\n .reg .pred p;\n setp.ne.b32 p, %4, 0;\n"
      " tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, {%5, %5, %5, %5, %5, %5, %5, %5}, p;\n}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(z) : "memory");
}
__device__ __forceinline__ void umma_commit_pair(uint32_t bar) {      // arrives on the barrier at this offset in BOTH CTAs
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(bar), "h"((uint16_t)3) : "memory");
}
// elected-lane forms for a CONVERGED issuing warp (see umma_tf32_elect in tc_common.cuh): issuing from inside
// `if (lane == 0)` costs >= 66 cycles per MMA in address arithmetic + ELECT / R2UR waterfall (scripts/ubench/
// umma_window_rate.cu), more than the 44.6 / 48 cycles an N = 32 / 64 MMA occupies the tensor pipe
__device__ __forceinline__ void umma_tf32_pair_elect(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                                     uint32_t accumulate) {
  const uint32_t z = 0;
  asm volatile(
      "{\n .reg .pred p, e;\n elect.sync _|e, 0xffffffff;\n setp.ne.b32 p, %4, 0;\n"
      " @e tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, {%5, %5, %5, %5, %5, %5, %5, %5}, p;\n}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(z) : "memory");
}
__device__ __forceinline__ void umma_commit_pair_elect(uint32_t bar) {
  asm volatile(
      "{\n .reg .pred e;\n elect.sync _|e, 0xffffffff;\n"
      " @e tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;\n}"
      ::"r"(bar), "h"((uint16_t)3) : "memory");
}
// Cross-CTA signalling of the pair kernels guards tensor-memory and async-proxy (TMA, tcgen05) traffic only: what a
// barrier phase has to order is covered by tcgen05.fence::before / after_thread_sync and by the complete_tx / commit
// mechanisms themselves.  The .release.cluster / .acquire.cluster forms used at first compile to MEMBAR.ALL + ERRBAR
// before every remote arrive (the epilogue warps then wait for their own earlier global STOREs to drain: 10 % of the
// stall samples) and to an L1 invalidate (CCTL.IVALL) after every wait; the default .cta-scope forms are what the
// 2-SM GEMM pipelines of CUTLASS use for the same hand-offs.  BN_PAIR_FENCES=1 at build time restores the heavy forms.
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
#ifdef BN_PAIR_FENCES
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
#else
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
#endif
}
__device__ __forceinline__ bool mbar_try_wait_cluster(uint32_t bar, uint32_t parity) {
  uint32_t ok;
#ifdef BN_PAIR_FENCES
  asm volatile(
      "{\n .reg .pred p;\n mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
      : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
#else
  asm volatile(
      "{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
      : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
#endif
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait_cluster(uint32_t bar, uint32_t parity) {
  for (uint32_t it = 0; !mbar_try_wait_cluster(bar, parity); ++it)
    if (it > (1u << 28)) __trap();
}

constexpr int PAIR_MAXCHUNK = 2;                      // C_in <= 64: half of the layer's weights fits one SM
struct HaloPairArgs {
  HaloArgs h;
  unsigned int w_off[PAIR_MAXCHUNK][HALO_MAXG];       // byte offset of (chunk, group)'s half weight tile in the W region
  unsigned int w_bytes;                               // resident bytes per CTA
};
struct alignas(64) HaloPairMaps {
  CUtensorMap a;       // NHWC input, box 32 ch x 10 px x 18 rows x 1 frame
  CUtensorMap b16;     // K-major weights, box 32 x 16 rows
  CUtensorMap b32;     // K-major weights, box 32 x 32 rows
};
constexpr int PAIR_EPI_WARPS = 8;                     // + warp 8 MMA issue, warp 9 TMA
constexpr int PAIR_THREADS = (PAIR_EPI_WARPS + 2) * 32;
constexpr int PAIR_SLOTS = 4;                         // halo buffers in flight: TMA latency / MMA time per chunk ~ 2-3
struct HaloPairSmem {
  static constexpr int W_MAX = 100 * 1024;            // 25 class-taps x 64 input channels x 16 rows x 4 B
  static constexpr int OFF_A = W_MAX;
  static constexpr int OFF_EPI = OFF_A + PAIR_SLOTS * HALO_ASTRIDE;
  static constexpr int OFF_BAR = OFF_EPI + PAIR_EPI_WARPS * 4096;
  static constexpr int TOTAL = OFF_BAR + 8 * (2 * PAIR_SLOTS + 6) + 32;
};
static_assert(HaloPairSmem::TOTAL <= 227 * 1024, "shared memory per CTA");

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(PAIR_THREADS, 1)
dgrad_halo_pair_kernel(const __grid_constant__ HaloPairMaps maps, const __grid_constant__ HaloPairArgs pa) {
  bn_pdl_trigger();
  constexpr int NB = 32, ACC_COLS = 4 * NB, NCOLS = 2 * ACC_COLS;
  using S = HaloPairSmem;
  extern __shared__ __align__(1024) unsigned char smem[];
  const HaloArgs& h = pa.h;
  const TcArgs& a = h.a;
  uint64_t* w_full = reinterpret_cast<uint64_t*>(smem + S::OFF_BAR);   // [1]  (leader's is the one waited on)
  uint64_t* a_full = w_full + 1;          // [PAIR_SLOTS]  leader: both CTAs' halo bytes
  uint64_t* a_empty = a_full + PAIR_SLOTS;   // [PAIR_SLOTS]  multicast commit
  uint64_t* acc_full = a_empty + PAIR_SLOTS; // [2]  multicast commit
  uint64_t* acc_empty = acc_full + 2;     // [2]  leader: 8 arrivals (4 epilogue warps x 2 CTAs)
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(acc_empty + 2);
  const int tid = threadIdx.x, warp = tid >> 5;
  const uint32_t rank = cluster_ctarank();
  if (tid == 0) {
    mbar_init(smem_u32(w_full), 1);
    for (int s = 0; s < PAIR_SLOTS; ++s) {
      mbar_init(smem_u32(a_full + s), 1);
      mbar_init(smem_u32(a_empty + s), 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(smem_u32(acc_full + s), 1);
      mbar_init(smem_u32(acc_empty + s), 2 * PAIR_EPI_WARPS);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == PAIR_EPI_WARPS) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr)), "n"(NCOLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                  // both CTAs' barriers exist before any remote signal / copy
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  bn_pdl_wait();
  const int Ci = a.Ci;
  const int nchunk = Ci / BK;
  const uint32_t smem_base = smem_u32(smem);
  const long long total = h.total_tiles;
  const long long npairs = (total + 1) / 2;
  const long long p_first = blockIdx.x >> 1, p_step = gridDim.x >> 1;

  if (warp < PAIR_EPI_WARPS) {
    // ======================= epilogue (own tile, own 128 accumulator rows) =======================
    // eight warps: warp w drains TMEM lanes 32 (w & 3) .. (its 32 pixels) of two of the four classes
    float* tile = reinterpret_cast<float*>(smem + S::OFF_EPI) + warp * 1024;
    const int elane = tid & 31;
    const int q = warp & 3, chalf = warp >> 2;
    const int pix = q * 32 + elane;
    float4 cacc = make_float4(0.f, 0.f, 0.f, 0.f);
    uint32_t leader_acc_empty[2];
#pragma unroll
    for (int s = 0; s < 2; ++s)
      asm volatile("mapa.shared::cluster.u32 %0, %1, 0;" : "=r"(leader_acc_empty[s]) : "r"(smem_u32(acc_empty + s)));
    int ti = 0;
    for (long long P = p_first; P < npairs; P += p_step, ++ti) {
      const long long T = 2 * P + rank;
      const bool tvalid = T < total;
      const int f = (int)(T / h.tiles_per_frame);
      const int blk = (int)(T - (long long)f * h.tiles_per_frame);
      const int by = blk / h.tiles_x, bx = blk - by * h.tiles_x;
      const int ym = by * 16 + (pix >> 3), xm = bx * 8 + (pix & 7);
      const int buf = ti & 1;
      mbar_wait_cluster(smem_u32(acc_full + buf), (ti >> 1) & 1);
      tc_fence_after();
      // (computing the eight row offsets per lane arithmetically instead of gathering them with shuffles, and
      // fetching both classes' accumulators before the first store, measured 8 % SLOWER: 82 vs 76 us per launch)
#pragma unroll
      for (int cc = 0; cc < 2; ++cc) {
        const int c = 2 * chalf + cc;
        const bool rvalid = tvalid && ym < h.Hm[c] && xm < h.Wm[c];
        const int oy = h.oy0[c] + a.os * ym, ox = h.ox0[c] + a.os * xm;
        const long long obase = (((long long)f * a.Ho + oy) * a.Wo + ox) * a.Co;
        uint32_t r[32];
        tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + buf * ACC_COLS + h.pos[c] * NB, r);
        tmem_ld_wait();
        warp_store_rows32(a.out, a.dact, BN_LEAK, rvalid ? obase : -1, r, a.bias, a.act, tile, elane, cacc, a.colsum != nullptr);
      }
      tc_fence_before();                       // this thread's TMEM reads of the buffer are complete
      __syncwarp();
      if (elane == 0) mbar_arrive_cluster(leader_acc_empty[buf]);
    }
    if (a.colsum) warp_flush_colsum(a.colsum, cacc, elane);
  } else if (warp == PAIR_EPI_WARPS) {
    // ======================= MMA issuer: leader CTA only =========================================
    if (rank == 0) {                           // whole warp converged, one elected lane issues (see igemm_tma_kernel)
      mbar_wait_cluster(smem_u32(w_full), 0);  // both CTAs' resident weights have landed
      tc_fence_after();
      constexpr uint32_t a_hi = desc_hi_sw128(HALO_W * 128), b_hi = desc_hi_sw128(1024);
      int ai = 0, ti = 0;
      for (long long P = p_first; P < npairs; P += p_step, ++ti) {
        const int buf = ti & 1;
        if (ti >= 2) {
          mbar_wait_cluster(smem_u32(acc_empty + buf), ((ti >> 1) - 1) & 1);
          tc_fence_after();
        }
        uint32_t accflag = 0u;
        for (int c = 0; c < nchunk; ++c, ++ai) {
          const int slot = ai % PAIR_SLOTS;
          mbar_wait_cluster(smem_u32(a_full + slot), (ai / PAIR_SLOTS) & 1);
          tc_fence_after();
          const uint32_t a_lo0 = desc_lo_sw128(smem_base + S::OFF_A + slot * HALO_ASTRIDE);
          for (int gi = 0; gi < h.ngroups; ++gi) {
            const HaloGroup g = h.g[gi];
            const uint32_t idesc = make_idesc(256, g.ncls * NB);
            const uint32_t acc = tmem_base + buf * ACC_COLS + g.col0 * NB;
            const uint32_t a_lo = a_lo0 + (uint32_t)g.row_off * 8u;
            const uint32_t b_lo = desc_lo_sw128(smem_base + pa.w_off[c][gi]);
#pragma unroll
            for (int k = 0; k < BK / 8; ++k) {
              umma_tf32_pair_elect(acc, ((uint64_t)a_hi << 32) | (a_lo + 2 * k), ((uint64_t)b_hi << 32) | (b_lo + 2 * k), idesc,
                                   accflag);
              accflag = 1u;
            }
          }
          umma_commit_pair_elect(smem_u32(a_empty + slot));
        }
        umma_commit_pair_elect(smem_u32(acc_full + buf));
      }
    }
    __syncwarp();
  } else {
    // ======================= TMA: resident half weights once, then this CTA's halos ===============
    const int lane = tid & 31;
    const uint32_t leader_w_full = smem_u32(w_full) & PEER_BIT_MASK;
    if (lane == 0) {
      if (rank == 0) mbar_expect_tx(smem_u32(w_full), 2u * pa.w_bytes);
    }
    __syncwarp();
    // one lane per (chunk, group): this CTA's half of the stacked class tiles
    for (int i = lane; i < nchunk * h.ngroups; i += 32) {
      const int c = i / h.ngroups, gi = i - c * h.ngroups;
      const HaloGroup g = h.g[gi];
      const uint32_t dst = smem_base + pa.w_off[c][gi];
      if (g.ncls == 1) {
        tma_tile_2d_pair(dst, &maps.b16, leader_w_full, g.wt[0] * Ci + c * BK, (int)rank * 16);
      } else {
        const int half = g.ncls / 2;                // whole class tiles per CTA
        for (int j = 0; j < half; ++j)
          tma_tile_2d_pair(dst + j * (NB * 128), &maps.b32, leader_w_full, g.wt[rank * half + j] * Ci + c * BK, 0);
      }
    }
    __syncwarp();
    if (lane == 0) {
      const uint32_t leader_a_full0 = smem_u32(a_full) & PEER_BIT_MASK;
      int q = 0;
      for (long long P = p_first; P < npairs; P += p_step) {
        const long long T = 2 * P + rank;
        // a tile past the end (odd tile count) loads frame n: out of bounds, zero fill, nothing stored
        const int f = (int)(T / h.tiles_per_frame);
        const int blk = (int)(T - (long long)f * h.tiles_per_frame);
        const int by = blk / h.tiles_x, bx = blk - by * h.tiles_x;
        for (int c = 0; c < nchunk; ++c, ++q) {
          const int slot = q % PAIR_SLOTS;
          if (q >= PAIR_SLOTS) mbar_wait_cluster(smem_u32(a_empty + slot), ((q / PAIR_SLOTS) - 1) & 1);
          if (rank == 0) mbar_expect_tx(smem_u32(a_full + slot), 2u * (uint32_t)HALO_ABYTES);
          tma_tile_4d_pair(smem_base + S::OFF_A + slot * HALO_ASTRIDE, &maps.a, leader_a_full0 + slot * 8, c * BK, bx * 8 + h.lo_x,
                           by * 16 + h.lo_y, f);
        }
      }
    }
    __syncwarp();
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                  // no CTA leaves (or frees TMEM) while its peer may still signal / read it
  if (warp == PAIR_EPI_WARPS) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(NCOLS) : "memory");
  }
}

// ------------------------------------------------------------------------------------------------
// CTA-pair halo kernel for C_out = 64 (ConvTranspose2d forward into / Conv2d backward-data out of the 64-channel
// 32 x 32 maps).  The layer's weights (819 KB at 128 input channels) cannot stay resident, but split over the pair
// every SM streams only HALF of each stacked class tile for TWO tiles' worth of MMAs (M = 256): a quarter of the
// one-CTA kernel's weight traffic per output pixel, which is what bounds that kernel (5.7 TB/s of L2 -> SM reads,
// 95 us for 26.8 GFLOP).  Same roles and cross-CTA signalling as dgrad_halo_pair_kernel, plus a ring of weight
// stages whose copies (both CTAs') complete on the leader's barrier.
// ------------------------------------------------------------------------------------------------
constexpr int P64_WST = 6;                               // weight stages: one MMA group each, <= 16 KB per CTA
constexpr int P64_WBYTES = 2 * 64 * 128;                 // two whole 64-row class tiles
struct HaloPair64Smem {
  static constexpr int OFF_A = P64_WST * P64_WBYTES;
  static constexpr int OFF_EPI = OFF_A + PAIR_SLOTS * HALO_ASTRIDE;
  static constexpr int OFF_BAR = OFF_EPI + PAIR_EPI_WARPS * 4096;
  static constexpr int TOTAL = OFF_BAR + 8 * (2 * PAIR_SLOTS + 2 * P64_WST + 4) + 32;
};
static_assert(HaloPair64Smem::TOTAL <= 227 * 1024, "shared memory per CTA");
struct alignas(64) HaloPair64Maps {
  CUtensorMap a;       // NHWC input, box 32 ch x 10 px x 18 rows x 1 frame
  CUtensorMap b32;     // K-major weights, box 32 x 32 rows
  CUtensorMap b64;     // K-major weights, box 32 x 64 rows
};

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(PAIR_THREADS, 1)
dgrad_halo_pair64_kernel(const __grid_constant__ HaloPair64Maps maps, const __grid_constant__ HaloArgs h) {
  bn_pdl_trigger();
  constexpr int NB = 64, ACC_COLS = 4 * NB, NCOLS = 2 * ACC_COLS;
  using S = HaloPair64Smem;
  extern __shared__ __align__(1024) unsigned char smem[];
  const TcArgs& a = h.a;
  uint64_t* a_full = reinterpret_cast<uint64_t*>(smem + S::OFF_BAR);   // [PAIR_SLOTS]  leader: both CTAs' halo bytes
  uint64_t* a_empty = a_full + PAIR_SLOTS;     // multicast commit
  uint64_t* b_full = a_empty + PAIR_SLOTS;     // [P64_WST]  leader: both CTAs' weight bytes
  uint64_t* b_empty = b_full + P64_WST;        // multicast commit
  uint64_t* acc_full = b_empty + P64_WST;      // [2]
  uint64_t* acc_empty = acc_full + 2;          // [2]  leader: 16 arrivals
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(acc_empty + 2);
  const int tid = threadIdx.x, warp = tid >> 5;
  const uint32_t rank = cluster_ctarank();
  if (tid == 0) {
    for (int s = 0; s < PAIR_SLOTS; ++s) { mbar_init(smem_u32(a_full + s), 1); mbar_init(smem_u32(a_empty + s), 1); }
    for (int s = 0; s < P64_WST; ++s) { mbar_init(smem_u32(b_full + s), 1); mbar_init(smem_u32(b_empty + s), 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(smem_u32(acc_full + s), 1); mbar_init(smem_u32(acc_empty + s), 2 * PAIR_EPI_WARPS); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == PAIR_EPI_WARPS) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr)), "n"(NCOLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  bn_pdl_wait();
  const int Ci = a.Ci;
  const int nchunk = Ci / BK;
  const uint32_t smem_base = smem_u32(smem);
  const long long total = h.total_tiles;
  const long long npairs = (total + 1) / 2;
  const long long p_first = blockIdx.x >> 1, p_step = gridDim.x >> 1;

  if (warp < PAIR_EPI_WARPS) {
    // ======================= epilogue: warp w drains lanes 32 (w & 3) .. of classes 2 (w >> 2), 2 (w >> 2) + 1 ==
    float* tile = reinterpret_cast<float*>(smem + S::OFF_EPI) + warp * 1024;
    const int elane = tid & 31;
    const int q = warp & 3, chalf = warp >> 2;
    const int pix = q * 32 + elane;
    float4 cacc[2] = {make_float4(0.f, 0.f, 0.f, 0.f), make_float4(0.f, 0.f, 0.f, 0.f)};
    uint32_t leader_acc_empty[2];
#pragma unroll
    for (int s = 0; s < 2; ++s)
      asm volatile("mapa.shared::cluster.u32 %0, %1, 0;" : "=r"(leader_acc_empty[s]) : "r"(smem_u32(acc_empty + s)));
    int ti = 0;
    for (long long P = p_first; P < npairs; P += p_step, ++ti) {
      const long long T = 2 * P + rank;
      const bool tvalid = T < total;
      const int f = (int)(T / h.tiles_per_frame);
      const int blk = (int)(T - (long long)f * h.tiles_per_frame);
      const int by = blk / h.tiles_x, bx = blk - by * h.tiles_x;
      const int ym = by * 16 + (pix >> 3), xm = bx * 8 + (pix & 7);
      const int buf = ti & 1;
      mbar_wait_cluster(smem_u32(acc_full + buf), (ti >> 1) & 1);
      tc_fence_after();
#pragma unroll
      for (int cc = 0; cc < 2; ++cc) {
        const int c = 2 * chalf + cc;
        const bool rvalid = tvalid && ym < h.Hm[c] && xm < h.Wm[c];
        const int oy = h.oy0[c] + a.os * ym, ox = h.ox0[c] + a.os * xm;
        const long long obase = (((long long)f * a.Ho + oy) * a.Wo + ox) * a.Co;
#pragma unroll
        for (int cb = 0; cb < 2; ++cb) {
          uint32_t r[32];
          tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + buf * ACC_COLS + h.pos[c] * NB + cb * 32, r);
          tmem_ld_wait();
          if (cc == 1 && cb == 1) {
            tc_fence_before();                 // the last block of this tile is in registers: release the accumulator
            __syncwarp();
            if (elane == 0) mbar_arrive_cluster(leader_acc_empty[buf]);
          }
          warp_store_rows32(a.out, a.dact, BN_LEAK, rvalid ? obase + cb * 32 : -1, r, a.bias ? a.bias + cb * 32 : nullptr, a.act,
                            tile, elane, cacc[cb], a.colsum != nullptr);
        }
      }
    }
    if (a.colsum) {
      warp_flush_colsum(a.colsum, cacc[0], elane);
      warp_flush_colsum(a.colsum + 32, cacc[1], elane);
    }
  } else if (warp == PAIR_EPI_WARPS) {
    // ======================= MMA issuer: leader CTA only, converged warp, one elected lane ========
    if (rank == 0) {
      constexpr uint32_t a_hi = desc_hi_sw128(HALO_W * 128), b_hi = desc_hi_sw128(1024);
      int ai = 0, bi = 0, ti = 0;
      for (long long P = p_first; P < npairs; P += p_step, ++ti) {
        const int buf = ti & 1;
        if (ti >= 2) {
          mbar_wait_cluster(smem_u32(acc_empty + buf), ((ti >> 1) - 1) & 1);
          tc_fence_after();
        }
        uint32_t accflag = 0u;
        for (int c = 0; c < nchunk; ++c, ++ai) {
          const int slot = ai % PAIR_SLOTS;
          mbar_wait_cluster(smem_u32(a_full + slot), (ai / PAIR_SLOTS) & 1);
          tc_fence_after();
          const uint32_t a_lo0 = desc_lo_sw128(smem_base + S::OFF_A + slot * HALO_ASTRIDE);
          for (int gi = 0; gi < h.ngroups; ++gi, ++bi) {
            const int stage = bi % P64_WST;
            mbar_wait_cluster(smem_u32(b_full + stage), (bi / P64_WST) & 1);
            tc_fence_after();
            const HaloGroup g = h.g[gi];
            const uint32_t idesc = make_idesc(256, g.ncls * NB);
            const uint32_t acc = tmem_base + buf * ACC_COLS + g.col0 * NB;
            const uint32_t a_lo = a_lo0 + (uint32_t)g.row_off * 8u;
            const uint32_t b_lo = desc_lo_sw128(smem_base + stage * P64_WBYTES);
#pragma unroll
            for (int k = 0; k < BK / 8; ++k) {
              umma_tf32_pair_elect(acc, ((uint64_t)a_hi << 32) | (a_lo + 2 * k), ((uint64_t)b_hi << 32) | (b_lo + 2 * k), idesc,
                                   accflag);
              accflag = 1u;
            }
            umma_commit_pair_elect(smem_u32(b_empty + stage));
          }
          umma_commit_pair_elect(smem_u32(a_empty + slot));
        }
        umma_commit_pair_elect(smem_u32(acc_full + buf));
      }
    }
    __syncwarp();
  } else {
    // ======================= TMA: this CTA's halos and its half of every stacked weight group =====
    const int lane = tid & 31;
    const uint32_t leader_a_full0 = smem_u32(a_full) & PEER_BIT_MASK;
    const uint32_t leader_b_full0 = smem_u32(b_full) & PEER_BIT_MASK;
    const long long my_tiles = p_first < npairs ? (npairs - p_first + p_step - 1) / p_step : 0;
    const long long nq = my_tiles * nchunk;            // chunk sequence of this CTA
    auto load_halo = [&](long long q) {                // whole warp calls; lane 0 issues
      const int slot = (int)(q % PAIR_SLOTS);
      if (q >= PAIR_SLOTS) mbar_wait_cluster(smem_u32(a_empty + slot), (uint32_t)((q / PAIR_SLOTS) - 1) & 1u);
      if (lane == 0) {
        const long long P = p_first + (q / nchunk) * p_step;
        const int c = (int)(q % nchunk);
        const long long T = 2 * P + rank;
        // a tile past the end (odd tile count) loads frame n: out of bounds, zero fill, nothing stored
        const int f = (int)(T / h.tiles_per_frame);
        const int blk = (int)(T - (long long)f * h.tiles_per_frame);
        const int by = blk / h.tiles_x, bx = blk - by * h.tiles_x;
        if (rank == 0) mbar_expect_tx(smem_u32(a_full + slot), 2u * (uint32_t)HALO_ABYTES);
        tma_tile_4d_pair(smem_base + S::OFF_A + slot * HALO_ASTRIDE, &maps.a, leader_a_full0 + slot * 8, c * BK, bx * 8 + h.lo_x,
                         by * 16 + h.lo_y, f);
      }
    };
    constexpr int AHEAD = 2;                           // halo chunks requested ahead of the weight stream
    for (long long q = 0; q < AHEAD && q < nq; ++q) load_halo(q);
    long long bi = 0;
    for (long long q = 0; q < nq; ++q) {
      if (q + AHEAD < nq) load_halo(q + AHEAD);
      const int c = (int)(q % nchunk);
      for (int gi = 0; gi < h.ngroups; ++gi, ++bi) {
        const int stage = (int)(bi % P64_WST);
        if (bi >= P64_WST) mbar_wait_cluster(smem_u32(b_empty + stage), (uint32_t)((bi / P64_WST) - 1) & 1u);
        const HaloGroup g = h.g[gi];
        const uint32_t bar = leader_b_full0 + stage * 8;
        if (lane == 0 && rank == 0) mbar_expect_tx(smem_u32(b_full + stage), 2u * (uint32_t)g.ncls * (NB * 128 / 2));
        __syncwarp();
        const uint32_t dst = smem_base + stage * P64_WBYTES;
        if (g.ncls == 1) {
          if (lane == 0) tma_tile_2d_pair(dst, &maps.b32, bar, g.wt[0] * Ci + c * BK, (int)rank * 32);
        } else {
          const int half = g.ncls / 2;                 // whole class tiles per CTA
          if (lane < half) tma_tile_2d_pair(dst + lane * (NB * 128), &maps.b64, bar, g.wt[rank * half + lane] * Ci + c * BK, 0);
        }
      }
    }
    __syncwarp();
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == PAIR_EPI_WARPS) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(NCOLS) : "memory");
  }
}

struct HaloKey {
  const void* in;
  const void* wt;
  int n, H, W, C, Co, wrow;
  bool operator==(const HaloKey& o) const {
    return in == o.in && wt == o.wt && n == o.n && H == o.H && W == o.W && C == o.C && Co == o.Co && wrow == o.wrow;
  }
};
std::vector<std::pair<HaloKey, HaloMaps>> g_halo_cache;

bool encode_tiled_4d(CUtensorMap* map, const float* p, int N, int H, int W, int C, int bc, int bw, int bh) {
  cuuint64_t gdim[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
  cuuint64_t gstr[3] = {(cuuint64_t)C * 4, (cuuint64_t)W * C * 4, (cuuint64_t)H * W * C * 4};
  cuuint32_t box[4] = {(cuuint32_t)bc, (cuuint32_t)bw, (cuuint32_t)bh, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = g_encode_tiled(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, (void*)p, gdim, gstr, box, estr,
                              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                              CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS;
}

template <int NB, int NST, bool PERSIST>
int launch_halo(const HaloMaps& maps, HaloArgs& h, cudaStream_t st) {
  using S = HaloSmem<NB, NST, PERSIST>;
  auto kern = dgrad_halo_kernel<NB, NST, PERSIST>;
  static bool configured = false;
  if (!configured) {
    BN_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, S::TOTAL));
    configured = true;
  }
  constexpr int CTAS_PER_SM = (2 * (S::TOTAL + 1024) <= 227 * 1024 && 2 * (PERSIST ? 2 : 1) * 4 * NB <= 512) ? 2 : 1;
  static_assert(S::TOTAL + 1024 <= 227 * 1024, "shared memory per CTA");
  // Cluster multicast of the weight tiles is implemented and parity-tested (BN_HALO_CLUSTER=2|4) but
  // measured SLOWER on B200 (C_out=32: 128 / 135 / 213 us at cluster size 1 / 2 / 4): the weight
  // stream is not what bounds this kernel, and the cluster couples the members' pipelines.  Default off.
  static const int max_csz = [] { const char* e = getenv("BN_HALO_CLUSTER"); return e ? atoi(e) : 1; }();
  int csz = 1;
  for (int c = 4; c >= 2; c >>= 1)
    if (c <= max_csz && h.total_tiles % c == 0) { csz = c; break; }
  long long grid = PERSIST ? CTAS_PER_SM * 148 : h.total_tiles;
  if (grid > h.total_tiles) grid = h.total_tiles;
  grid -= grid % csz;
  if (grid <= 0 || grid > 0x7fffffffLL) return 1;
  h.csz = csz;
  static const int dbg = [] { const char* e = getenv("BN_HALO_DBG"); return e ? atoi(e) : 0; }();
  h.dbg = dbg;
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3((unsigned)grid);
  cfg.blockDim = dim3(HALO_THREADS);
  cfg.dynamicSmemBytes = S::TOTAL;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = csz;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = bn_pdl_enabled() ? 2 : 1;
  BN_CUDA(cudaLaunchKernelEx(&cfg, kern, maps, (const HaloArgs&)h));
  BN_LAUNCHED();
  return 0;
}

struct HaloPairKey {
  const void* in;
  const void* wt;
  int n, H, W, C, wrow;
  bool operator==(const HaloPairKey& o) const {
    return in == o.in && wt == o.wt && n == o.n && H == o.H && W == o.W && C == o.C && wrow == o.wrow;
  }
};
std::vector<std::pair<HaloPairKey, HaloPairMaps>> g_halo_pair_cache;

// returns 1 when not applicable
int launch_halo_pair(const TcArgs& a, const HaloArgs& h, cudaStream_t st) {
  HaloPairArgs pa;
  memset(&pa, 0, sizeof(pa));
  pa.h = h;
  pa.h.csz = 2;
  pa.h.dbg = 0;
  const int nchunk = a.Ci / BK;
  unsigned int off = 0;
  for (int c = 0; c < nchunk; ++c)
    for (int gi = 0; gi < h.ngroups; ++gi) {
      pa.w_off[c][gi] = off;
      off += (unsigned int)h.g[gi].ncls * 16 * 128;            // half of ncls stacked 32 x 32 tiles
    }
  pa.w_bytes = off;
  if (off > (unsigned int)HaloPairSmem::W_MAX) return 1;
  const HaloPairMaps* pm = nullptr;
  {
    HaloPairKey key{a.in, a.wt, a.n, a.Hi, a.Wi, a.Ci, a.wrow};
    std::lock_guard<std::mutex> lk(g_tma_mutex);
    for (auto& kv : g_halo_pair_cache)
      if (kv.first == key) pm = &kv.second;
    if (!pm) {
      HaloPairMaps m;
      memset(&m, 0, sizeof(m));
      if (!encode_tiled_4d(&m.a, a.in, a.n, a.Hi, a.Wi, a.Ci, BK, HALO_W, HALO_H)) return 1;
      if (!encode_tiled_2d(&m.b16, a.wt, a.Co, a.wrow, BK, 16, CU_TENSOR_MAP_SWIZZLE_128B)) return 1;
      if (!encode_tiled_2d(&m.b32, a.wt, a.Co, a.wrow, BK, 32, CU_TENSOR_MAP_SWIZZLE_128B)) return 1;
      if (g_halo_pair_cache.size() >= 256) g_halo_pair_cache.clear();
      g_halo_pair_cache.emplace_back(key, m);
      pm = &g_halo_pair_cache.back().second;
    }
  }
  HaloPairMaps local = *pm;
  auto kern = dgrad_halo_pair_kernel;
  static bool configured = false;
  if (!configured) {
    BN_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, HaloPairSmem::TOTAL));
    configured = true;
  }
  const long long npairs = (h.total_tiles + 1) / 2;
  long long grid = 2 * (npairs < 74 ? npairs : 74);             // one CTA per SM, whole pairs
  BN_CUDA(bn_launch(kern, dim3((unsigned)grid), PAIR_THREADS, HaloPairSmem::TOTAL, st, local, (const HaloPairArgs&)pa));
  BN_LAUNCHED();
  return 0;
}

// ------------------------------------------------------------------------------------------------
// CTA-pair halo kernel for the FPROP-form stride-2 layers with 32 input and 64 output channels (encoder conv1
// forward, decoder convtranspose3 backward-data: the two slowest launches of a C2 step).
//
// igemm_tma_kernel<64> fetches every input pixel once per filter tap that touches it (25 / 4 times) plus the
// layer's weights per tile: 1.05 GB through L2 for 208 MB of HBM traffic, and it runs at exactly that L2 -> SM
// rate (~10 TB/s, 109 us).  A stride-2 convolution is four stride-1 convolutions over the PARITY PLANES of its
// input (in[2y'+py][2x'+px]): tap (dy, dx) reads plane (dy & 1, dx & 1) at offset ((dy - py) / 2, (dx - px) / 2),
// which is in {-1, 0, 1}.  So a CTA owns 16 x 8 output pixels, stages the 18 x 10 halo of each parity plane with
// ONE tiled TMA copy (the plane is a strided view of the NHWC tensor: its own tensor map), and every tap is that
// buffer at a row offset -- 1.4 fetches per input pixel instead of 6.25.  As in dgrad_halo_pair_kernel two CTAs
// issue one tcgen05.mma.cta_group::2 (M = 256, N = 64) with the weight operand split between them, so all 25
// weight tiles (100 KB per SM) are loaded once and stay resident.
// ------------------------------------------------------------------------------------------------
constexpr int FP_MAXT = 25;
struct FpropTap {
  unsigned short row_off;     // window offset inside the plane halo, in pixels
  unsigned char plane, wt;
};
struct FpropPairArgs {
  TcArgs a;
  int tiles_x, tiles_per_frame;
  long long total_tiles;
  int Hm, Wm, oy0, ox0;
  int ay_min[4], ax_min[4];   // per plane: plane-space offset of the halo origin relative to the tile origin
  int ntaps_plane[4];         // taps are sorted by plane
  int dbg;
  FpropTap t[FP_MAXT];
};
struct alignas(64) FpropPairMaps {
  CUtensorMap a[4];    // parity planes of the NHWC input, box 32 ch x 10 px x 18 rows x 1 frame
  CUtensorMap b32;     // K-major weights, box 32 x 32 rows
};

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(PAIR_THREADS, 1)
fprop_halo_pair_kernel(const __grid_constant__ FpropPairMaps maps, const __grid_constant__ FpropPairArgs fa) {
  bn_pdl_trigger();
  constexpr int NB = 64, ACC_COLS = NB, NCOLS = 2 * ACC_COLS;
  using S = HaloPairSmem;
  extern __shared__ __align__(1024) unsigned char smem[];
  const TcArgs& a = fa.a;
  uint64_t* w_full = reinterpret_cast<uint64_t*>(smem + S::OFF_BAR);
  uint64_t* a_full = w_full + 1;
  uint64_t* a_empty = a_full + PAIR_SLOTS;
  uint64_t* acc_full = a_empty + PAIR_SLOTS;
  uint64_t* acc_empty = acc_full + 2;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(acc_empty + 2);
  const int tid = threadIdx.x, warp = tid >> 5;
  const uint32_t rank = cluster_ctarank();
  if (tid == 0) {
    mbar_init(smem_u32(w_full), 1);
    for (int s = 0; s < PAIR_SLOTS; ++s) {
      mbar_init(smem_u32(a_full + s), 1);
      mbar_init(smem_u32(a_empty + s), 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(smem_u32(acc_full + s), 1);
      mbar_init(smem_u32(acc_empty + s), 2 * PAIR_EPI_WARPS);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == PAIR_EPI_WARPS) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr)), "n"(NCOLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  bn_pdl_wait();
  const uint32_t smem_base = smem_u32(smem);
  const long long total = fa.total_tiles;
  const long long npairs = (total + 1) / 2;
  const long long p_first = blockIdx.x >> 1, p_step = gridDim.x >> 1;
#define FP_STAMP(tile, slot)                                                                        \
  do {                                                                                              \
    if (fa.dbg && blockIdx.x == 0 && (tile) < 8) g_halo_dbg[(tile) * 8 + (slot)] = dbg_clock();     \
  } while (0)
  const int ntaps = fa.ntaps_plane[0] + fa.ntaps_plane[1] + fa.ntaps_plane[2] + fa.ntaps_plane[3];

  if (warp < PAIR_EPI_WARPS) {
    // ======================= epilogue: warp w drains lanes 32 (w & 3) .. of column block w >> 2 ===
    float* tile = reinterpret_cast<float*>(smem + S::OFF_EPI) + warp * 1024;
    const int elane = tid & 31;
    const int q = warp & 3, chalf = warp >> 2;
    const int pix = q * 32 + elane;
    float4 cacc = make_float4(0.f, 0.f, 0.f, 0.f);
    uint32_t leader_acc_empty[2];
#pragma unroll
    for (int s = 0; s < 2; ++s)
      asm volatile("mapa.shared::cluster.u32 %0, %1, 0;" : "=r"(leader_acc_empty[s]) : "r"(smem_u32(acc_empty + s)));
    const float* bias = a.bias ? a.bias + chalf * 32 : nullptr;
    int ti = 0;
    for (long long P = p_first; P < npairs; P += p_step, ++ti) {
      const long long T = 2 * P + rank;
      const bool tvalid = T < total;
      const int f = (int)(T / fa.tiles_per_frame);
      const int blk = (int)(T - (long long)f * fa.tiles_per_frame);
      const int by = blk / fa.tiles_x, bx = blk - by * fa.tiles_x;
      const int ym = by * 16 + (pix >> 3), xm = bx * 8 + (pix & 7);
      const int buf = ti & 1;
      mbar_wait_cluster(smem_u32(acc_full + buf), (ti >> 1) & 1);
      tc_fence_after();
      if (tid == 0) FP_STAMP(ti, 6);
      const bool rvalid = tvalid && ym < fa.Hm && xm < fa.Wm;
      const int oy = fa.oy0 + a.os * ym, ox = fa.ox0 + a.os * xm;
      const long long obase = (((long long)f * a.Ho + oy) * a.Wo + ox) * a.Co + chalf * 32;
      uint32_t r[32];
      tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + buf * ACC_COLS + chalf * 32, r);
      tmem_ld_wait();
      tc_fence_before();                       // the accumulator block is in registers: release it before the stores
      __syncwarp();
      if (elane == 0) mbar_arrive_cluster(leader_acc_empty[buf]);
      warp_store_rows32(a.out, a.dact, BN_LEAK, rvalid ? obase : -1, r, bias, a.act, tile, elane, cacc, a.colsum != nullptr);
      if (tid == 0) FP_STAMP(ti, 7);
    }
    if (a.colsum) warp_flush_colsum(a.colsum + chalf * 32, cacc, elane);
  } else if (warp == PAIR_EPI_WARPS) {
    // ======================= MMA issuer: leader CTA only =========================================
    if (rank == 0) {                           // whole warp, converged; one elected lane issues
      mbar_wait_cluster(smem_u32(w_full), 0);
      tc_fence_after();
      constexpr uint32_t idesc = make_idesc(256, NB);
      constexpr uint32_t a_hi = desc_hi_sw128(HALO_W * 128), b_hi = desc_hi_sw128(1024);
      const uint32_t b_lo0 = desc_lo_sw128(smem_base);
      int ai = 0, ti = 0;
      for (long long P = p_first; P < npairs; P += p_step, ++ti) {
        const int buf = ti & 1;
        if (ti >= 2) {
          mbar_wait_cluster(smem_u32(acc_empty + buf), ((ti >> 1) - 1) & 1);
          tc_fence_after();
        }
        const uint32_t acc = tmem_base + buf * ACC_COLS;
        if ((tid & 31) == 0) FP_STAMP(ti, 0);
        int i = 0;
        uint32_t accflag = 0u;
#pragma unroll 1
        for (int pl = 0; pl < 4; ++pl, ++ai) {
          const int slot = ai % PAIR_SLOTS;
          mbar_wait_cluster(smem_u32(a_full + slot), (ai / PAIR_SLOTS) & 1);
          tc_fence_after();
          if ((tid & 31) == 0) FP_STAMP(ti, 1 + pl);
          const uint32_t a_lo0 = desc_lo_sw128(smem_base + S::OFF_A + slot * HALO_ASTRIDE);
          const int jend = i + fa.ntaps_plane[pl];
#pragma unroll 1
          for (; i < jend; ++i) {
            const uint32_t a_lo = a_lo0 + (uint32_t)fa.t[i].row_off * 8u;       // row_off * 128 bytes >> 4
            const uint32_t b_lo = b_lo0 + (uint32_t)i * (32 * 128 / 16);
#pragma unroll
            for (int kk = 0; kk < BK / 8; ++kk) {
              umma_tf32_pair_elect(acc, ((uint64_t)a_hi << 32) | (a_lo + 2 * kk), ((uint64_t)b_hi << 32) | (b_lo + 2 * kk), idesc,
                                   accflag);
              accflag = 1u;
            }
          }
          umma_commit_pair_elect(smem_u32(a_empty + slot));
        }
        umma_commit_pair_elect(smem_u32(acc_full + buf));
        if ((tid & 31) == 0) FP_STAMP(ti, 5);
      }
    }
    __syncwarp();
  } else {
    // ======================= TMA: resident half weights once, then this CTA's plane halos ==========
    const int lane = tid & 31;
    const uint32_t leader_w_full = smem_u32(w_full) & PEER_BIT_MASK;
    if (lane == 0 && rank == 0) mbar_expect_tx(smem_u32(w_full), 2u * (uint32_t)ntaps * (32 * 128));
    __syncwarp();
    for (int i = lane; i < ntaps; i += 32)
      tma_tile_2d_pair(smem_base + i * (32 * 128), &maps.b32, leader_w_full, fa.t[i].wt * a.Ci, (int)rank * 32);
    __syncwarp();
    if (lane == 0) {
      const uint32_t leader_a_full0 = smem_u32(a_full) & PEER_BIT_MASK;
      int q = 0;
      for (long long P = p_first; P < npairs; P += p_step) {
        const long long T = 2 * P + rank;
        // a tile past the end (odd tile count) loads frame n: out of bounds, zero fill, nothing stored
        const int f = (int)(T / fa.tiles_per_frame);
        const int blk = (int)(T - (long long)f * fa.tiles_per_frame);
        const int by = blk / fa.tiles_x, bx = blk - by * fa.tiles_x;
#pragma unroll 1
        for (int pl = 0; pl < 4; ++pl, ++q) {
          const int slot = q % PAIR_SLOTS;
          if (q >= PAIR_SLOTS) mbar_wait_cluster(smem_u32(a_empty + slot), ((q / PAIR_SLOTS) - 1) & 1);
          if (rank == 0) mbar_expect_tx(smem_u32(a_full + slot), 2u * (uint32_t)HALO_ABYTES);
          tma_tile_4d_pair(smem_base + S::OFF_A + slot * HALO_ASTRIDE, &maps.a[pl], leader_a_full0 + slot * 8, 0,
                           bx * 8 + fa.ax_min[pl], by * 16 + fa.ay_min[pl], f);
        }
      }
    }
    __syncwarp();
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == PAIR_EPI_WARPS) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(NCOLS) : "memory");
  }
}

struct FpropPairKey {
  const void* in;
  const void* wt;
  int n, H, W, C, wrow;
  bool operator==(const FpropPairKey& o) const {
    return in == o.in && wt == o.wt && n == o.n && H == o.H && W == o.W && C == o.C && wrow == o.wrow;
  }
};
std::vector<std::pair<FpropPairKey, FpropPairMaps>> g_fprop_pair_cache;

// parity plane (py, px) of an NHWC image as a 4-D tensor map: pixel (y', x') = image pixel (2 y' + py, 2 x' + px)
bool encode_plane_4d(CUtensorMap* map, const float* p, int N, int H, int W, int C, int py, int px, int bc, int bw, int bh) {
  const int Hp = (H - py + 1) / 2, Wp = (W - px + 1) / 2;
  if (Hp < 1 || Wp < 1) return false;
  cuuint64_t gdim[4] = {(cuuint64_t)C, (cuuint64_t)Wp, (cuuint64_t)Hp, (cuuint64_t)N};
  cuuint64_t gstr[3] = {(cuuint64_t)2 * C * 4, (cuuint64_t)2 * W * C * 4, (cuuint64_t)H * W * C * 4};
  cuuint32_t box[4] = {(cuuint32_t)bc, (cuuint32_t)bw, (cuuint32_t)bh, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = g_encode_tiled(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, (void*)(p + ((long long)py * W + px) * C), gdim, gstr,
                              box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                              CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS;
}

struct HaloPair64Key {
  const void* in;
  const void* wt;
  int n, H, W, C, wrow;
  bool operator==(const HaloPair64Key& o) const {
    return in == o.in && wt == o.wt && n == o.n && H == o.H && W == o.W && C == o.C && wrow == o.wrow;
  }
};
std::vector<std::pair<HaloPair64Key, HaloPair64Maps>> g_halo_pair64_cache;

// returns 1 when not applicable
int launch_halo_pair64(const TcArgs& a, const HaloArgs& h, cudaStream_t st) {
  for (int gi = 0; gi < h.ngroups; ++gi)
    if (h.g[gi].ncls != 1 && h.g[gi].ncls != 2 && h.g[gi].ncls != 4) return 1;
  HaloArgs hh = h;
  hh.csz = 2;
  hh.dbg = 0;
  const HaloPair64Maps* pm = nullptr;
  {
    HaloPair64Key key{a.in, a.wt, a.n, a.Hi, a.Wi, a.Ci, a.wrow};
    std::lock_guard<std::mutex> lk(g_tma_mutex);
    for (auto& kv : g_halo_pair64_cache)
      if (kv.first == key) pm = &kv.second;
    if (!pm) {
      HaloPair64Maps m;
      memset(&m, 0, sizeof(m));
      if (!encode_tiled_4d(&m.a, a.in, a.n, a.Hi, a.Wi, a.Ci, BK, HALO_W, HALO_H)) return 1;
      if (!encode_tiled_2d(&m.b32, a.wt, a.Co, a.wrow, BK, 32, CU_TENSOR_MAP_SWIZZLE_128B)) return 1;
      if (!encode_tiled_2d(&m.b64, a.wt, a.Co, a.wrow, BK, 64, CU_TENSOR_MAP_SWIZZLE_128B)) return 1;
      if (g_halo_pair64_cache.size() >= 256) g_halo_pair64_cache.clear();
      g_halo_pair64_cache.emplace_back(key, m);
      pm = &g_halo_pair64_cache.back().second;
    }
  }
  HaloPair64Maps local = *pm;
  auto kern = dgrad_halo_pair64_kernel;
  static bool configured = false;
  if (!configured) {
    BN_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, HaloPair64Smem::TOTAL));
    configured = true;
  }
  const long long npairs = (h.total_tiles + 1) / 2;
  long long grid = 2 * (npairs < 74 ? npairs : 74);             // one CTA per SM, whole pairs
  BN_CUDA(bn_launch(kern, dim3((unsigned)grid), PAIR_THREADS, HaloPair64Smem::TOTAL, st, local, (const HaloArgs&)hh));
  BN_LAUNCHED();
  return 0;
}

// returns 1 when the op is not a stride-2 fprop-form layer with 32 input / 64 output channels
int try_fprop_halo_pair(const TcArgs& a, const TapClass* hc, int nclasses, cudaStream_t st) {
  static const bool off = [] { const char* e = getenv("BN_FPROP_HALO"); return e && e[0] == '0'; }();
  if (off || nclasses != 1 || a.gs != 2 || a.os != 1 || a.ksplit > 1) return 1;
  if (a.Ci != BK || a.Co != 64 || !tma_available()) return 1;
  const TapClass& k = hc[0];
  if (k.ntaps < 4 || k.ntaps > FP_MAXT || k.Hm < 16 || k.Wm < 8) return 1;
  FpropPairArgs fa;
  memset(&fa, 0, sizeof(fa));
  fa.a = a;
  fa.Hm = k.Hm; fa.Wm = k.Wm; fa.oy0 = k.oy0; fa.ox0 = k.ox0;
  int lo_y[4], lo_x[4], hi_y[4], hi_x[4];
  for (int p = 0; p < 4; ++p) { lo_y[p] = lo_x[p] = 127; hi_y[p] = hi_x[p] = -128; }
  auto plane_of = [](int dy, int dx) { return (dy & 1) * 2 + (dx & 1); };
  for (int t = 0; t < k.ntaps; ++t) {
    const int p = plane_of(k.dy[t], k.dx[t]);
    const int ay = (k.dy[t] - (k.dy[t] & 1)) >> 1, ax = (k.dx[t] - (k.dx[t] & 1)) >> 1;
    lo_y[p] = ay < lo_y[p] ? ay : lo_y[p]; hi_y[p] = ay > hi_y[p] ? ay : hi_y[p];
    lo_x[p] = ax < lo_x[p] ? ax : lo_x[p]; hi_x[p] = ax > hi_x[p] ? ax : hi_x[p];
    ++fa.ntaps_plane[p];
  }
  // slot order: planes with fewer taps first, so their halo buffers are released (and refilled for the next tile)
  // as early as possible -- with one tile's planes in flight the kernel is bound by the latency of those refills
  int cnt[4], order[4] = {0, 1, 2, 3};
  for (int p = 0; p < 4; ++p) { cnt[p] = fa.ntaps_plane[p]; if (cnt[p] == 0) return 1; }
  for (int i = 0; i < 4; ++i)
    for (int j = i + 1; j < 4; ++j)
      if (cnt[order[j]] < cnt[order[i]]) { const int tmp = order[i]; order[i] = order[j]; order[j] = tmp; }
  int n = 0;
  for (int s = 0; s < 4; ++s) {
    const int p = order[s];
    if (hi_y[p] - lo_y[p] > HALO_H - 16 || hi_x[p] - lo_x[p] > HALO_W - 8) return 1;
    fa.ntaps_plane[s] = cnt[p];
    fa.ay_min[s] = lo_y[p]; fa.ax_min[s] = lo_x[p];
    for (int t = 0; t < k.ntaps; ++t)
      if (plane_of(k.dy[t], k.dx[t]) == p) {
        const int ay = (k.dy[t] - (k.dy[t] & 1)) >> 1, ax = (k.dx[t] - (k.dx[t] & 1)) >> 1;
        fa.t[n].row_off = (unsigned short)((ay - lo_y[p]) * HALO_W + (ax - lo_x[p]));
        fa.t[n].plane = (unsigned char)s;
        fa.t[n].wt = k.wt[t];
        ++n;
      }
  }
  if ((size_t)n * 32 * 128 > (size_t)HaloPairSmem::W_MAX) return 1;
  fa.tiles_x = bn_cdiv(k.Wm, 8);
  fa.tiles_per_frame = fa.tiles_x * bn_cdiv(k.Hm, 16);
  fa.total_tiles = (long long)fa.tiles_per_frame * a.n;
  if (fa.total_tiles < 2) return 1;
  static const int dbg = [] { const char* e = getenv("BN_HALO_DBG"); return e ? atoi(e) : 0; }();
  fa.dbg = dbg;
  const FpropPairMaps* pm = nullptr;
  {
    FpropPairKey key{a.in, a.wt, a.n, a.Hi, a.Wi, a.Ci, a.wrow};
    std::lock_guard<std::mutex> lk(g_tma_mutex);
    for (auto& kv : g_fprop_pair_cache)
      if (kv.first == key) pm = &kv.second;
    if (!pm) {
      FpropPairMaps m;
      memset(&m, 0, sizeof(m));
      for (int s = 0; s < 4; ++s)
        if (!encode_plane_4d(&m.a[s], a.in, a.n, a.Hi, a.Wi, a.Ci, order[s] >> 1, order[s] & 1, BK, HALO_W, HALO_H)) return 1;
      if (!encode_tiled_2d(&m.b32, a.wt, a.Co, a.wrow, BK, 32, CU_TENSOR_MAP_SWIZZLE_128B)) return 1;
      if (g_fprop_pair_cache.size() >= 256) g_fprop_pair_cache.clear();
      g_fprop_pair_cache.emplace_back(key, m);
      pm = &g_fprop_pair_cache.back().second;
    }
  }
  FpropPairMaps local = *pm;
  auto kern = fprop_halo_pair_kernel;
  static bool configured = false;
  if (!configured) {
    BN_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, HaloPairSmem::TOTAL));
    configured = true;
  }
  const long long npairs = (fa.total_tiles + 1) / 2;
  long long grid = 2 * (npairs < 74 ? npairs : 74);             // one CTA per SM, whole pairs
  BN_CUDA(bn_launch(kern, dim3((unsigned)grid), PAIR_THREADS, HaloPairSmem::TOTAL, st, local, (const FpropPairArgs&)fa));
  BN_LAUNCHED();
  return 0;
}

// returns 1 when the op does not have the stride-2 four-class shape this kernel covers
int try_dgrad_halo(const TcArgs& a, const TapClass* hc, int nclasses, int* colsum_fused, float* colsum_out,
                   float* colpart, size_t colpart_floats, cudaStream_t st) {
  static const bool off = [] { const char* e = getenv("BN_HALO"); return e && e[0] == '0'; }();
  if (off || nclasses != 4 || a.gs != 1 || a.os != 2 || a.ksplit > 1) return 1;
  if (a.Co != 32 && a.Co != 64) return 1;
  if (!tma_available()) return 1;
  int lo_y = 127, lo_x = 127, hi_y = -128, hi_x = -128, Hm = 0, Wm = 0, ntaps = 0;
  for (int c = 0; c < 4; ++c) {
    const TapClass& k = hc[c];
    if (k.ntaps < 1) return 1;
    ntaps += k.ntaps;
    Hm = k.Hm > Hm ? k.Hm : Hm;
    Wm = k.Wm > Wm ? k.Wm : Wm;
    for (int t = 0; t < k.ntaps; ++t) {
      lo_y = k.dy[t] < lo_y ? k.dy[t] : lo_y; hi_y = k.dy[t] > hi_y ? k.dy[t] : hi_y;
      lo_x = k.dx[t] < lo_x ? k.dx[t] : lo_x; hi_x = k.dx[t] > hi_x ? k.dx[t] : hi_x;
    }
  }
  if (hi_y - lo_y > HALO_H - 16 || hi_x - lo_x > HALO_W - 8) return 1;
  // small images would leave most of the 16 x 8 block empty: the im2col kernel packs frames instead
  if (Hm < 16 || Wm < 8) return 1;

  HaloArgs h;
  memset(&h, 0, sizeof(h));
  h.a = a; h.lo_y = lo_y; h.lo_x = lo_x; h.tiles_x = bn_cdiv(Wm, 8);
  for (int c = 0; c < 4; ++c) { h.Hm[c] = hc[c].Hm; h.Wm[c] = hc[c].Wm; h.oy0[c] = hc[c].oy0; h.ox0[c] = hc[c].ox0; }
  // windows: tap offsets -> the set of classes that use them (at most one tap per class and window)
  const int nwy = hi_y - lo_y + 1, nwx = hi_x - lo_x + 1;
  int wmask[9] = {0}, wtap[9][4];
  for (int c = 0; c < 4; ++c)
    for (int t = 0; t < hc[c].ntaps; ++t) {
      const int w = (hc[c].dy[t] - lo_y) * nwx + (hc[c].dx[t] - lo_x);
      if (wmask[w] & (1 << c)) return 1;
      wmask[w] |= 1 << c;
      wtap[w][c] = hc[c].wt[t];
    }
  // TMEM order of the classes: a permutation that makes every window's class set contiguous, and
  // puts a window that covers all four classes first (its MMA initialises every accumulator)
  int perm[4] = {0, 1, 2, 3}, best[4] = {-1, -1, -1, -1};
  const int perms[24][4] = {{0,1,2,3},{0,1,3,2},{0,2,1,3},{0,2,3,1},{0,3,1,2},{0,3,2,1},{1,0,2,3},{1,0,3,2},
                            {1,2,0,3},{1,2,3,0},{1,3,0,2},{1,3,2,0},{2,0,1,3},{2,0,3,1},{2,1,0,3},{2,1,3,0},
                            {2,3,0,1},{2,3,1,0},{3,0,1,2},{3,0,2,1},{3,1,0,2},{3,1,2,0},{3,2,0,1},{3,2,1,0}};
  (void)perm;
  int full = -1;
  for (int w = 0; w < nwy * nwx; ++w)
    if (wmask[w] == 15) { full = w; break; }
  if (full < 0) return 1;
  for (int pi = 0; pi < 24 && best[0] < 0; ++pi) {
    bool ok = true;
    for (int w = 0; w < nwy * nwx && ok; ++w) {
      if (!wmask[w]) continue;
      int lo = 4, hi = -1, cnt = 0;
      for (int c = 0; c < 4; ++c)
        if (wmask[w] & (1 << c)) { lo = perms[pi][c] < lo ? perms[pi][c] : lo; hi = perms[pi][c] > hi ? perms[pi][c] : hi; ++cnt; }
      ok = hi - lo + 1 == cnt;
    }
    if (ok) for (int c = 0; c < 4; ++c) best[c] = perms[pi][c];
  }
  if (best[0] < 0) return 1;
  for (int c = 0; c < 4; ++c) h.pos[c] = best[c];
  int ng = 0;
  auto add_group = [&](int w) {
    HaloGroup g;
    memset(&g, 0, sizeof(g));
    g.row_off = (unsigned short)((w / nwx) * HALO_W + (w % nwx));
    int lo = 4, cnt = 0;
    for (int c = 0; c < 4; ++c)
      if (wmask[w] & (1 << c)) { lo = best[c] < lo ? best[c] : lo; ++cnt; }
    g.col0 = (unsigned char)lo; g.ncls = (unsigned char)cnt;
    for (int c = 0; c < 4; ++c)
      if (wmask[w] & (1 << c)) g.wt[best[c] - lo] = (unsigned char)wtap[w][c];
    h.g[ng++] = g;
  };
  add_group(full);
  for (int w = 0; w < nwy * nwx; ++w)
    if (wmask[w] && w != full) add_group(w);
  h.ngroups = ng;
  (void)ntaps;


  const HaloMaps* hm = nullptr;
  {
    HaloKey key{a.in, a.wt, a.n, a.Hi, a.Wi, a.Ci, a.Co, a.wrow};
    std::lock_guard<std::mutex> lk(g_tma_mutex);
    for (auto& kv : g_halo_cache)
      if (kv.first == key) hm = &kv.second;
    if (!hm) {
      HaloMaps m;
      memset(&m, 0, sizeof(m));
      if (!encode_tiled_4d(&m.a, a.in, a.n, a.Hi, a.Wi, a.Ci, BK, HALO_W, HALO_H)) return 1;
      if (!encode_tiled_2d(&m.b, a.wt, a.Co, a.wrow, BK, a.Co, CU_TENSOR_MAP_SWIZZLE_128B)) return 1;
      if (g_halo_cache.size() >= 256) g_halo_cache.clear();
      g_halo_cache.emplace_back(key, m);
      hm = &g_halo_cache.back().second;
    }
  }
  h.tiles_per_frame = h.tiles_x * bn_cdiv(Hm, 16);
  h.total_tiles = (long long)h.tiles_per_frame * a.n;
  // C_out = 32 with <= 64 input channels: CTA-pair kernel, the layer's weights resident (half per SM)
  static const bool pair_off = [] { const char* e = getenv("BN_HALO_PAIR"); return e && e[0] == '0'; }();
  if (a.Co == 32 && a.Ci <= 32 * PAIR_MAXCHUNK && !pair_off && h.total_tiles >= 2) {
    int r = launch_halo_pair(a, h, st);
    if (r <= 0) return r;
  }
  // C_out = 64: CTA-pair kernel with the weight stream split between the two SMs (BN_HALO_PAIR64=0: one-CTA kernel)
  static const bool pair64_off = [] { const char* e = getenv("BN_HALO_PAIR64"); return e && e[0] == '0'; }();
  if (a.Co == 64 && !pair64_off && h.total_tiles >= 2) {
    int r = launch_halo_pair64(a, h, st);
    if (r <= 0) return r;
  }
  HaloMaps local = *hm;
  static const int deep = [] { const char* e = getenv("BN_HALO_DEEP"); return e ? atoi(e) : 0; }();
  if (a.Co == 32 && deep == 1) return launch_halo<32, 10, true>(local, h, st);      // one CTA per SM, 160 KB weight ring
  if (a.Co == 32) return launch_halo<32, 3, true>(local, h, st);
  // one tile per CTA: no atomics; the CTA's column sums become one row of the partial table
  h.a.colsum = nullptr;
  if (colsum_fused) *colsum_fused = 0;
  h.a.colpart = nullptr;
  if (colsum_out && colpart && bn_reduce_deferring() && (size_t)h.total_tiles * a.Co <= colpart_floats)
    h.a.colpart = colpart;
  int r = launch_halo<64, 2, false>(local, h, st);
  if (r == 0 && h.a.colpart) {
    BN_TRY(bn_colsum_reduce_defer(colpart, (int)h.total_tiles, a.Co, colsum_out));
    if (colsum_fused) *colsum_fused = 1;
  }
  return r;
}

}  // namespace

namespace {
// tile configuration of the TMA kernels: "stages,mtiles" from BN_WG<bn> / BN_IG<bn>, else the defaults
struct TileCfg { int stages, mt; };
TileCfg env_cfg(const char* name, TileCfg dflt) {
  const char* e = getenv(name);
  if (!e) return dflt;
  TileCfg c = dflt;
  if (sscanf(e, "%d,%d", &c.stages, &c.mt) != 2) return dflt;
  return c;
}
}  // namespace

int bn_launch_igemm_tc(const ImgView& in, const float* wt, int wrow, const float* bias, float* out,
                       int Ho, int Wo, int Co, const float* dact, const TapClass* d_classes,
                       const TapClass* h_classes, int nclasses, int maxM, int maxtaps, int gs, int os, int n,
                       int act, float* split_buf, size_t split_floats, float* colsum, int* colsum_fused,
                       float* colpart, size_t colpart_floats, cudaStream_t st) {
  if (colsum_fused) *colsum_fused = 0;
  // shapes this kernel covers: NHWC-dense input with C % 32 == 0, C_out in {32, 64, 128, 256, 512},
  // enough rows to fill the machine (the stride-5 layers with a few hundred rows stay on the
  // CUDA-core kernel until the split-K variant lands)
  if (wt == nullptr) return 1;
  if (in.sc != 1 || in.C % BK != 0 || in.sx != in.C || in.sy != (long long)in.W * in.C ||
      in.sn != (long long)in.H * in.W * in.C)
    return 1;
  if (((uintptr_t)in.p & 15) || ((uintptr_t)out & 15) || ((uintptr_t)wt & 15) || (dact && ((uintptr_t)dact & 15)) ||
      (bias && ((uintptr_t)bias & 15)))
    return 1;
  const int bn = Co >= 256 ? 256 : Co;
  if (Co != 32 && Co != 64 && Co != 128 && Co != 256 && Co != 512) return 1;
  const long long M = (long long)n * maxM;
  const long long ctas = (long long)bn_cdiv(M, BM) * (Co / bn) * nclasses;
  const int nchunks = maxtaps * (in.C / BK);
  int ksplit = 1;
  if (ctas < 64 && nclasses == 1 && os == 1 && split_buf != nullptr && nchunks >= 16) {
    // few output tiles but a long reduction (the stride-5 layers): split the taps over grid.z
    ksplit = (int)((148 + ctas - 1) / ctas);
    if (ksplit > nchunks / 4) ksplit = nchunks / 4;
    if (ksplit > 32) ksplit = 32;
    while (ksplit > 1 && (size_t)ksplit * M * Co > split_floats) --ksplit;
  }
  if (ctas * ksplit < 24) return 1;            // too little work to fill the machine: CUDA-core kernel
  TcArgs a;
  a.ksplit = ksplit;
  a.split_out = split_buf;
  a.in = in.p; a.Hi = in.H; a.Wi = in.W; a.Ci = in.C; a.wt = wt; a.wrow = wrow; a.bias = bias; a.out = out;
  a.Ho = Ho; a.Wo = Wo; a.Co = Co; a.dact = dact; a.classes = d_classes; a.gs = gs; a.os = os;
  a.n = n; a.act = act;
  // the epilogue can accumulate the column sums of what it stores (split-K leaves that to the reducer)
  a.colsum = (ksplit == 1 && !((uintptr_t)colsum & 15)) ? colsum : nullptr;
  a.colpart = nullptr;
  if (colsum_fused) *colsum_fused = a.colsum != nullptr;
  int r = try_dgrad_halo(a, h_classes, nclasses, colsum_fused, a.colsum, colpart, colpart_floats, st);
  if (r <= 0) return r;
  r = try_fprop_halo_pair(a, h_classes, nclasses, st);
  if (r <= 0) return r;
  // one tile per CTA below: a flush per tile makes ~10^4 same-address atomics per column, which costs
  // more than the separate column-sum pass saves (measured).  Instead every CTA writes its column sums
  // as one row of a partial table and the batched reduction kernel of the backward call adds them up.
  // two M-tiles per CTA (shared weight tiles) when that still fills the machine; BN_IG<co>="stages,mtiles" overrides
  static const TileCfg ig32 = env_cfg("BN_IG32", {3, 2}), ig64 = env_cfg("BN_IG64", {2, 2}), ig128 = env_cfg("BN_IG128", {2, 2});
  const TileCfg ig = Co == 32 ? ig32 : Co == 64 ? ig64 : Co == 128 ? ig128 : TileCfg{4, 1};
  const TmaSet* tm = get_tma_set(in, wt, wrow, Co, bn, d_classes, h_classes, nclasses, gs, n);
  const int MTs = (tm && ig.mt == 2 && ksplit == 1 && Co <= 128 && (long long)bn_cdiv(M, 2 * BM) * (Co / bn) * nclasses >= 148)
                      ? 2 : 1;
  const long long col_rows = (long long)bn_cdiv(M, MTs * BM) * nclasses;
  const bool part = a.colsum != nullptr && colpart != nullptr && !((uintptr_t)colpart & 15) && bn_reduce_deferring() &&
                    (size_t)col_rows * Co <= colpart_floats;
  float* colsum_out = a.colsum;
  a.colsum = nullptr;
  a.colpart = part ? colpart : nullptr;
  if (colsum_fused) *colsum_fused = 0;
  if (tm) {
    TmaSet local = *tm;      // copied into the kernel parameter space (__grid_constant__)
    const int code = Co * 100 + MTs * 10 + ig.stages;
    switch (MTs == ig.mt ? code : Co * 100 + 10 + (Co == 128 ? 3 : 4)) {
      case 3200 + 14: r = launch_tma<32, 4, 1>(local, a, nclasses, maxM, st); break;
      case 3200 + 23: r = launch_tma<32, 3, 2>(local, a, nclasses, maxM, st); break;
      case 6400 + 14: r = launch_tma<64, 4, 1>(local, a, nclasses, maxM, st); break;
      case 6400 + 22: r = launch_tma<64, 2, 2>(local, a, nclasses, maxM, st); break;
      case 6400 + 23: r = launch_tma<64, 3, 2>(local, a, nclasses, maxM, st); break;
      case 6400 + 24: r = launch_tma<64, 4, 2>(local, a, nclasses, maxM, st); break;
      case 12800 + 13: r = launch_tma<128, 3, 1>(local, a, nclasses, maxM, st); break;
      case 12800 + 22: r = launch_tma<128, 2, 2>(local, a, nclasses, maxM, st); break;
      case 12800 + 23: r = launch_tma<128, 3, 2>(local, a, nclasses, maxM, st); break;
      case 12800 + 24: r = launch_tma<128, 4, 2>(local, a, nclasses, maxM, st); break;
      default:
        if (Co >= 256) r = launch_tma<256, 4, 1>(local, a, nclasses, maxM, st);
        else BN_FAIL("igemm: no kernel instance for Co=%d stages/M-tiles code %d", Co, code);
    }
  } else {
    switch (Co) {
      case 32: r = launch_tc<32, 4>(a, nclasses, maxM, st); break;
      case 64: r = launch_tc<64, 4>(a, nclasses, maxM, st); break;
      case 128: r = launch_tc<128, 3>(a, nclasses, maxM, st); break;
      default: r = launch_tc<256, 4>(a, nclasses, maxM, st); break;
    }
  }
  if (r == 0 && part) {
    BN_TRY(bn_colsum_reduce_defer(colpart, (int)col_rows, Co, colsum_out));
    if (colsum_fused) *colsum_fused = 1;
  }
  if (r || ksplit == 1) return r;
  const long long total = M * Co;          // fprop-form output is linear in (m, co)
  BN_CUDA(bn_launch(splitk_reduce_kernel, dim3(bn_cdiv(total / 4, 256)), 256, 0, st, split_buf, ksplit, total, Co, bias, dact, act, out));
  BN_LAUNCHED();
  return 0;
}

namespace {

struct WgKey {
  const void* big;
  const void* small;
  int n, H, W, C, Hs, Ws, Cs, k, s, pt, pl;
  bool operator==(const WgKey& o) const {
    return big == o.big && small == o.small && n == o.n && H == o.H && W == o.W && C == o.C && Hs == o.Hs &&
           Ws == o.Ws && Cs == o.Cs && k == o.k && s == o.s && pt == o.pt && pl == o.pl;
  }
};
std::vector<std::pair<WgKey, WgTmaSet>> g_wg_cache;

const WgTmaSet* get_wgrad_tma_set(const ImgView& big, const float* small, const ConvGeom& g, int n, long long M) {
  if (!tma_available()) return nullptr;
  WgKey key{big.p, small, n, big.H, big.W, big.C, g.Hs, g.Ws, g.Cs, g.k, g.s, g.pt, g.pl};
  std::lock_guard<std::mutex> lk(g_tma_mutex);
  for (auto& kv : g_wg_cache)
    if (kv.first == key) return &kv.second;
  WgTmaSet tm;
  memset(&tm, 0, sizeof(tm));
  tm.lw = -g.pl; tm.lh = -g.pt;
  if (!encode_im2col(&tm.a, big.p, n, big.H, big.W, big.C, tm.lw, tm.lh, g.Ws, g.Hs, g.s, BK,
                     CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B))
    return nullptr;
  if (!encode_tiled_2d(&tm.s, small, M, g.Cs, 32, BK, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B)) return nullptr;
  if (g_wg_cache.size() >= 256) g_wg_cache.clear();
  g_wg_cache.emplace_back(key, tm);
  return &g_wg_cache.back().second;
}

}  // namespace

namespace {
TileCfg wg_cfg(int bn) {
  static const TileCfg c32 = env_cfg("BN_WG32", {3, 2}), c64 = env_cfg("BN_WG64", {2, 2}), c128 = env_cfg("BN_WG128", {3, 1}),
                       c256 = env_cfg("BN_WG256", {3, 2});
  return bn == 32 ? c32 : bn == 64 ? c64 : bn == 128 ? c128 : c256;
}
int wg_mtiles(int bn) { return wg_cfg(bn).mt; }
int wg_stages(int bn, int mt) {
  const TileCfg c = wg_cfg(bn);
  if (c.mt == mt) return c.stages;
  return bn == 128 ? 3 : 4;          // fell back to one M-tile (short Ktot): the single-tile defaults
}
int wg_ctas_per_sm(int bn, int mt) {
  const size_t stage = (size_t)mt * 16384 + (size_t)bn * 128;
  const size_t total = wg_stages(bn, mt) * stage + 2048;
  int c = (int)((227 * 1024) / total);
  const int tmem = 512 / (mt * bn < 32 ? 32 : mt * bn);
  c = c < tmem ? c : tmem;
  return c < 1 ? 1 : (c > 2 ? 2 : c);
}
}  // namespace

int bn_launch_wgrad_tc(const ImgView& big, const float* small, const ConvGeom& g, int n, float* partial,
                       size_t partial_floats, float* grad, cudaStream_t st) {
  if (n <= 0 || grad == nullptr) return 0;
  if (big.sc != 1 || big.C % 32 != 0 || big.sx != big.C || big.sy != (long long)big.W * big.C ||
      big.sn != (long long)big.H * big.W * big.C)
    return 1;
  if (((uintptr_t)big.p & 15) || ((uintptr_t)small & 15) || ((uintptr_t)partial & 15)) return 1;
  const int Cs = g.Cs;
  int bn;
  switch (Cs) {
    case 32: bn = 32; break;
    case 64: bn = 64; break;
    case 128: bn = 128; break;
    case 256: case 512: bn = 256; break;
    default: return 1;
  }
  const int Ktot = g.k * g.k * g.Cb;
  const long long M = (long long)n * g.Hs * g.Ws;
  if (M < 256) return 1;
  // several (tap, channel) tiles per CTA share every small-image tile when the TMA path is available
  const WgTmaSet* tm = get_wgrad_tma_set(big, small, g, n, M);
  int MTs = tm ? wg_mtiles(bn) : 1;
  while (MTs > 1 && Ktot <= (MTs / 2) * BM) MTs /= 2;
  long long tiles = (long long)bn_cdiv(Ktot, MTs * BM) * (Cs / bn);
  // one full wave and no more: the CTAs are long-running (each walks its whole pixel slice), so a
  // grid that exceeds the resident slots by a handful of CTAs doubles the kernel time
  const long long slots = 148LL * wg_ctas_per_sm(bn, MTs);
  long long splits = slots / tiles;
  if (splits < 1) splits = 1;
  long long maxs = (M + 127) / 128;
  if (splits > maxs) splits = maxs;
  long long cap = (long long)(partial_floats / ((size_t)Ktot * Cs));
  if (splits > cap) splits = cap;
  if (splits < 1) return 1;
  long long rps = (M + splits - 1) / splits;
  rps = (rps + BK - 1) / BK * BK;
  splits = (M + rps - 1) / rps;
  WgTcArgs a;
  a.big = big.p; a.Hb = big.H; a.Wb = big.W; a.Cb = big.C; a.small = small; a.Hs = g.Hs; a.Ws = g.Ws;
  a.Cs = Cs; a.cls = g.d_fprop; a.gs = g.s; a.n = n; a.Ktot = Ktot; a.rows_per_split = rps;
  a.partial = partial;
  int r;
  if (tm) {
    WgTmaSet local = *tm;
    // (stages, M-tiles) per BN: defaults are the measured-best ones (profiles/r02_tile_sweep.txt); BN_WG<bn>="s,m" overrides
    const int cfg = MTs * 10 + wg_stages(bn, MTs);
    switch (bn * 100 + cfg) {
      case 3200 + 14: r = launch_wgrad_tma<32, 4, 1>(local, a, (int)splits, st); break;
      case 3200 + 23: r = launch_wgrad_tma<32, 3, 2>(local, a, (int)splits, st); break;
      case 6400 + 14: r = launch_wgrad_tma<64, 4, 1>(local, a, (int)splits, st); break;
      case 6400 + 22: r = launch_wgrad_tma<64, 2, 2>(local, a, (int)splits, st); break;
      case 6400 + 23: r = launch_wgrad_tma<64, 3, 2>(local, a, (int)splits, st); break;
      case 6400 + 42: r = launch_wgrad_tma<64, 2, 4>(local, a, (int)splits, st); break;
      case 6400 + 43: r = launch_wgrad_tma<64, 3, 4>(local, a, (int)splits, st); break;
      case 12800 + 13: r = launch_wgrad_tma<128, 3, 1>(local, a, (int)splits, st); break;
      case 12800 + 22: r = launch_wgrad_tma<128, 2, 2>(local, a, (int)splits, st); break;
      case 12800 + 23: r = launch_wgrad_tma<128, 3, 2>(local, a, (int)splits, st); break;
      case 12800 + 42: r = launch_wgrad_tma<128, 2, 4>(local, a, (int)splits, st); break;
      case 25600 + 14: r = launch_wgrad_tma<256, 4, 1>(local, a, (int)splits, st); break;
      case 25600 + 23: r = launch_wgrad_tma<256, 3, 2>(local, a, (int)splits, st); break;
      default: BN_FAIL("wgrad: no kernel instance for BN=%d stages/M-tiles code %d", bn, cfg);
    }
    if (r) return r;
    return bn_launch_wgrad_reduce(partial, (int)splits, Ktot, Cs, g.Cb, g.k * g.k, g.d_fprop, grad, st);
  }
  switch (bn) {
    case 32: r = launch_wgrad_tc<32, 4>(a, (int)splits, st); break;
    case 64: r = launch_wgrad_tc<64, 4>(a, (int)splits, st); break;
    case 128: r = launch_wgrad_tc<128, 3>(a, (int)splits, st); break;
    default: r = launch_wgrad_tc<256, 4>(a, (int)splits, st); break;
  }
  if (r) return r;
  return bn_launch_wgrad_reduce(partial, (int)splits, Ktot, Cs, g.Cb, g.k * g.k, g.d_fprop, grad, st);
}


// debug aid: phase timestamps (ns, %globaltimer) of CTA 0's first 8 tiles in the last halo-kernel launch
// run with BN_HALO_DBG=1: per tile [mma loop top, acc buffer free, mmas issued, epi wait, acc ready, epi done]
extern "C" int bn_debug_halo_times(long long* h_out) {
  BN_CUDA(cudaMemcpyFromSymbol(h_out, g_halo_dbg, sizeof(long long) * 64));
  return 0;
}
