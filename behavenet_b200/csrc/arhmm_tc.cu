// ARHMM emission log-likelihoods on the 5th-gen tensor cores (tcgen05), fp32-accurate via the
// 3xTF32 split.
//
// The AR Gaussian log-density of every state is one whitened affine map (arhmm.cu):
//     y_k = W_k psi_t,   psi_t = [x_t, x_{t-1}, .., x_{t-L}, 1],   ll_t(k) = c_k - |y_k|^2 / 2,
// i.e. a (timesteps x KT) x (KT x K*D) GEMM followed by per-state sums of squares.  On the FP32
// pipe that GEMM is ~13.6 kFLOP per timestep and bounds the whole E-step; here a CTA turns 128
// consecutive timesteps of a trial into one UMMA tile:
//     D[128 x NT] (TMEM, fp32) = psi_lo * W_hi^T + psi_hi * W_lo^T + psi_hi * W_hi^T
// with psi split on the fly (cvt.rna.tf32 and the exact remainder) and W split on the host in
// fp64.  Each dropped term (lo * lo) is below 2^-22 relative, so the result carries fp32-level error
// and the posteriors keep the 1e-5 parity bound of the CUDA-core kernel.
//
// Operands are K-major in the no-swizzle canonical layout: 8-row x 16-byte core matrices, cores
// adjacent along M/N 128 bytes apart (SBO), cores adjacent along K one whole column of cores apart
// (LBO) -- thread r writes the four floats of k-chunk c of its row at c*LBO + r*16, a fully
// coalesced 2 KB store per chunk.  Warps 0-3 build psi, wait for the accumulator, read their TMEM
// lane quarter back with tcgen05.ld and reduce per state; warp 4 allocates TMEM and issues the
// MMAs.  Two CTAs per SM overlap one tile's MMAs with the other's epilogue.
#include <stdlib.h>

#include "arhmm_common.cuh"
#include "tc_common.cuh"

namespace {

using namespace bn_tc;

constexpr int TM = 128;          // timesteps per tile = UMMA M
constexpr int NTHREADS = 160;

// K-major, SWIZZLE_NONE shared-memory matrix descriptor
__device__ __forceinline__ uint64_t make_desc_k_nosw(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;        // descriptor version 1 (sm_100)
  return d;                      // layout type 0 = no swizzle
}

__device__ __forceinline__ float tf32_rna(float x) {
  uint32_t u;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(x));
  return __uint_as_float(u);
}

__device__ __forceinline__ float2 ffma2_tc(float2 a, float2 b, float2 c) {
  unsigned long long ua = *reinterpret_cast<unsigned long long*>(&a), ub = *reinterpret_cast<unsigned long long*>(&b),
                     uc = *reinterpret_cast<unsigned long long*>(&c), ud;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(ud) : "l"(ua), "l"(ub), "l"(uc));
  return *reinterpret_cast<float2*>(&ud);
}

__device__ __forceinline__ void epi_bar() { asm volatile("bar.sync 1, 128;" ::: "memory"); }

__host__ __device__ constexpr int gcd_c(int a, int b) { return b == 0 ? a : gcd_c(b, a % b); }

// phase timestamps of CTA 0 (debug aid, BN_EMIT_DBG=1; bn_debug_emit_times): per tile
// [loop top, psi posted, x prefetch issued, accumulator ready, epilogue done]
__device__ long long g_emit_dbg[32 * 8];
__device__ __forceinline__ long long emit_clock() {
  long long v;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(v) :: "memory");    // clobber: a stamp must not let stores slide past it
  return v;
}
#define EMIT_STAMP(slot)                                                                              \
  do {                                                                                                \
    if (a.dbg && blockIdx.x == 0 && blockIdx.y == 0 && tid == 0 && ntile < 16) g_emit_dbg[ntile * 8 + (slot)] = emit_clock(); \
  } while (0)

struct EmitTcArgs {
  int dbg;
  const unsigned char* blob;
  const float* x;
  const long long* offsets;
  int K, D, lags, n_trials;
  float* Bsc;
  float* mx;
};

template <int DP>
__global__ void __launch_bounds__(NTHREADS) emission_tc_kernel(const EmitTcArgs a) {
  constexpr int SC = DP * 32 / gcd_c(DP, 32);      // columns per epilogue super-chunk (whole states, whole 32-col loads)
  constexpr int SPC = SC / DP;                     // states per super-chunk
  extern __shared__ __align__(1024) unsigned char smem[];
  const BlobHeader* hd = reinterpret_cast<const BlobHeader*>(a.blob);
  const int K = a.K, D = a.D, L = a.lags;
  const int NT = hd->NT, KT = hd->KT;
  const uint32_t A_LBO = TM * 16, B_LBO = (uint32_t)NT * 16;
  // smem: W_hi | W_lo | psi_hi | psi_lo | c | barriers
  float* Bhi = reinterpret_cast<float*>(smem);
  float* Blo = Bhi + (size_t)NT * KT;
  float* Ahi = Blo + (size_t)NT * KT;
  float* Alo = Ahi + (size_t)TM * KT;
  float* csm = Alo + (size_t)TM * KT;                        // K + 1
  uint64_t* bars = reinterpret_cast<uint64_t*>(csm + ((K + 1 + 3) & ~3));   // full, accum
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 2);
  const uint32_t full_bar = smem_u32(bars), accum_bar = smem_u32(bars + 1);

  const int tid = threadIdx.x;
  const int warp = tid >> 5;
  {
    const int4* s0 = reinterpret_cast<const int4*>(a.blob + hd->off_W_hi);
    const int4* s1 = reinterpret_cast<const int4*>(a.blob + hd->off_W_lo);
    int4* d0 = reinterpret_cast<int4*>(Bhi);
    int4* d1 = reinterpret_cast<int4*>(Blo);
    const int n16 = NT * KT / 4;
    for (int i = tid; i < n16; i += NTHREADS) { d0[i] = __ldg(s0 + i); d1[i] = __ldg(s1 + i); }
    const float* cg = reinterpret_cast<const float*>(a.blob + hd->off_c_f);
    for (int i = tid; i < K + 1; i += NTHREADS) csm[i] = cg[i];
  }
  if (tid == 0) {
    mbar_init(full_bar, 4 + (L > 0 ? 1 : 0));   // one arrival per producer warp (+ warp 4 for the halo rows)
    mbar_init(accum_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 4) tmem_alloc<256>(smem_u32(tmem_ptr));
  fence_proxy_async();             // W tiles (generic-proxy stores) -> visible to the tensor core
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  uint32_t phase = 0;

  // Row r of the psi tile is timestep t0 + r; x_t feeds row r (block 0), row r+1 (block 1), .. so every
  // x row is split into hi / lo ONCE by its owner thread and scattered to the L+1 rows that use it.
  // The first L rows of a tile also need x_{t0-1} .. x_{t0-L}: lanes 0..L-1 of warp 4 own those.
  auto load_x = [&](long long beg, int T, int ts, float4 (&v)[DP / 4]) {
    const bool ok = ts >= 0 && ts < T;
    const float* xr = a.x + (beg + (ok ? ts : 0)) * D;
#pragma unroll
    for (int c = 0; c < DP / 4; ++c) {
      v[c] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (ok) {
        if (D == DP) {
          v[c] = __ldg(reinterpret_cast<const float4*>(xr) + c);
        } else {
          float e[4];
#pragma unroll
          for (int q = 0; q < 4; ++q) e[q] = 4 * c + q < D ? __ldg(xr + 4 * c + q) : 0.f;
          v[c] = make_float4(e[0], e[1], e[2], e[3]);
        }
      }
    }
  };
  // scatter x_{t0 + r0} (hi / lo) to rows r0 + b, b = 0..L, that exist in this tile
  auto scatter_x = [&](int r0, const float4 (&v)[DP / 4]) {
#pragma unroll
    for (int c = 0; c < DP / 4; ++c) {
      const float4 h = make_float4(tf32_rna(v[c].x), tf32_rna(v[c].y), tf32_rna(v[c].z), tf32_rna(v[c].w));
      const float4 l = make_float4(tf32_rna(v[c].x - h.x), tf32_rna(v[c].y - h.y), tf32_rna(v[c].z - h.z),
                                   tf32_rna(v[c].w - h.w));
      for (int b = 0; b <= L; ++b) {
        const int row = r0 + b;
        if (row < 0 || row >= TM) continue;
        const size_t off = (size_t)(b * (DP / 4) + c) * A_LBO + row * 16;
        *reinterpret_cast<float4*>(reinterpret_cast<unsigned char*>(Ahi) + off) = h;
        *reinterpret_cast<float4*>(reinterpret_cast<unsigned char*>(Alo) + off) = l;
      }
    }
  };
  const bool halo_lane = warp == 4 && (tid & 31) < L;      // owns x_{t0 - 1 - lane}
  const int my_r0 = warp < 4 ? tid : -1 - (tid & 31);

  // work list of this CTA: (trial, t0) pairs in a fixed order known to every warp
  int trial = blockIdx.x;
  int t0 = blockIdx.y * TM;
  long long beg = 0;
  int T = 0;
  auto advance = [&]() {             // move to the next non-empty tile; false when the list is exhausted
    while (trial < a.n_trials) {
      beg = a.offsets[trial];
      T = (int)(a.offsets[trial + 1] - beg);
      if (t0 < T) return true;
      trial += gridDim.x;
      t0 = blockIdx.y * TM;
    }
    return false;
  };
  auto step_tile = [&]() { t0 += gridDim.y * TM; };

  bool have = advance();
  float4 xv[DP / 4];
  if (have && (warp < 4 || halo_lane)) load_x(beg, T, t0 + my_r0, xv);
  int ntile = 0;
  while (have) {
    EMIT_STAMP(0);
    const long long cbeg = beg;
    const int cT = T, ct0 = t0;
    if (warp < 4 || halo_lane) scatter_x(my_r0, xv);
    if (warp < 4) {
      for (int kc = (L + 1) * (DP / 4); kc < KT / 4; ++kc) {
        const float one = kc == (L + 1) * (DP / 4) ? 1.f : 0.f;      // bias column, then zero padding
        *reinterpret_cast<float4*>(reinterpret_cast<unsigned char*>(Ahi) + (size_t)kc * A_LBO + tid * 16) =
            make_float4(one, 0.f, 0.f, 0.f);
        *reinterpret_cast<float4*>(reinterpret_cast<unsigned char*>(Alo) + (size_t)kc * A_LBO + tid * 16) =
            make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
    fence_proxy_async();
    tc_fence_before();               // this thread's TMEM reads of the previous tile are complete
    // one mbarrier arrival per warp (128 single-thread arrivals on one shared word serialise):
    // __syncwarp orders the lanes' writes before lane 0's releasing arrive
    __syncwarp();
    if ((tid & 31) == 0 && (warp < 4 || L > 0)) mbar_arrive(full_bar);
    EMIT_STAMP(1);
    // next tile's x rows are fetched while the tensor core works on this one
    step_tile();
    have = advance();
    float qinit = 0.f;
    if (warp < 4 && ct0 + tid < L) {
      // initial segment: N(0, I) for every state (ssm mu_init = 0, Sigma_init = I)
#pragma unroll
      for (int c = 0; c < DP / 4; ++c)
        qinit += xv[c].x * xv[c].x + xv[c].y * xv[c].y + xv[c].z * xv[c].z + xv[c].w * xv[c].w;
    }
    if (have && (warp < 4 || halo_lane)) load_x(beg, T, t0 + my_r0, xv);
    EMIT_STAMP(2);

    if (warp < 4) {
      // ---------------- epilogue
      const int t = ct0 + tid;
      mbar_wait(accum_bar, phase);
      tc_fence_after();
      EMIT_STAMP(3);
      float vmax = -INFINITY;
      float vals[32];                       // K <= 32 log-likelihoods of this row
#pragma unroll
      for (int k = 0; k < 32; ++k) vals[k] = -INFINITY;
#pragma unroll
      for (int sc = 0; sc < (32 + SPC - 1) / SPC; ++sc) {
        if (sc * SPC < K) {
          uint32_t r[SC / 32][32];
#pragma unroll
          for (int c32 = 0; c32 < SC / 32; ++c32)
            tmem_ld32(tmem_base + ((uint32_t)(warp * 32) << 16) + sc * SC + c32 * 32, r[c32]);
          tmem_ld_wait();
          float q[SPC];
#pragma unroll
          for (int s = 0; s < SPC; ++s) q[s] = 0.f;
#pragma unroll
          for (int c32 = 0; c32 < SC / 32; ++c32)
#pragma unroll
            for (int e = 0; e < 32; ++e) {
              const float y = __uint_as_float(r[c32][e]);
              q[(c32 * 32 + e) / DP] = fmaf(y, y, q[(c32 * 32 + e) / DP]);
            }
#pragma unroll
          for (int s = 0; s < SPC; ++s) {
            const int k = sc * SPC + s;
            if (k < 32 && k < K) {
              vals[k] = csm[k] - 0.5f * q[s];
              vmax = fmaxf(vmax, vals[k]);
            }
          }
        }
      }
      if (t < cT) {
        const bool init = t < L;
        float* og = a.Bsc + (cbeg + t) * K;
        a.mx[cbeg + t] = init ? csm[K] - 0.5f * qinit : vmax;
        if ((K & 3) == 0) {
#pragma unroll
          for (int k = 0; k < 32; k += 4)
            if (k < K)
              *reinterpret_cast<float4*>(og + k) =
                  init ? make_float4(1.f, 1.f, 1.f, 1.f)
                       : make_float4(__expf(vals[k] - vmax), __expf(vals[k + 1] - vmax), __expf(vals[k + 2] - vmax),
                                     __expf(vals[k + 3] - vmax));
        } else {
#pragma unroll
          for (int k = 0; k < 32; ++k)
            if (k < K) og[k] = init ? 1.f : __expf(vals[k] - vmax);
        }
      }
      EMIT_STAMP(4);
    } else if ((tid & 31) == 0) {
      // ---------------- MMA issuer
      mbar_wait(full_bar, phase);
      tc_fence_after();
      const uint32_t idesc = make_idesc(TM, NT);
      const uint32_t ahi = smem_u32(Ahi), alo = smem_u32(Alo), bhi = smem_u32(Bhi), blo = smem_u32(Blo);
      const int ksteps = KT / 8;
      for (int pass = 0; pass < 3; ++pass) {
        const uint32_t pa = pass == 0 ? alo : ahi;       // small terms first
        const uint32_t pb = pass == 1 ? blo : bhi;
        for (int s = 0; s < ksteps; ++s) {
          const uint64_t ad = make_desc_k_nosw(pa + s * 2 * A_LBO, A_LBO, 128);
          const uint64_t bd = make_desc_k_nosw(pb + s * 2 * B_LBO, B_LBO, 128);
          umma_tf32(tmem_base, ad, bd, idesc, (pass | s) != 0 ? 1u : 0u);
        }
      }
      umma_commit(accum_bar);
    }
    if (warp == 4) {
      // the halo lanes rewrite rows of the psi tile in the next iteration: not before this tile's
      // MMAs have retired (and never more than one tile ahead of the row owners)
      __syncwarp();
      mbar_wait(accum_bar, phase);
    }
    phase ^= 1;
    ++ntile;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 4) {
    tc_fence_after();
    tmem_dealloc<256>(tmem_base);
  }
}

// ------------------------------------------------------------------------------------------------
// Warp-specialised persistent form (default; D % 4 == 0).  What the measurements on the kernel above said (C4):
//   * ncu: 31 % of all stall samples sit on the accumulator mbarrier -- the four worker warps of a CTA build psi, WAIT
//     for the tile's 15 MMAs and only then run the epilogue; the tensor pipe is 39 % busy, the spin loops are 15 % of
//     the executed instructions;
//   * %globaltimer stamps in a first decoupled version: with the roles separated the BUILDERS set the pace, 1.4-2 us
//     per tile although their scatter is 64 ns of work -- every register-prefetched global load of x sits on the same
//     few scoreboards as the loads issued for later tiles, so "4 tiles ahead" degenerates into waiting for the newest
//     request (deeper register prefetch made the kernel SLOWER, 240 -> 355 us);
//   * micro-benchmarks (scripts/ubench): tcgen05.ld sustains 130 B/cycle per warp with or without MMAs in flight and
//     the no-swizzle operand layout feeds an N = 192 MMA at the tensor rate (96 cycles) -- neither is the limit.
// So x never goes through registers here.  One CTA per SM:
//   warp  13    copy producer: one cp.async.bulk per tile (its TM + lags rows of x are contiguous) into an NXS-deep
//               staging ring, completion on an mbarrier
//   warps 0-3   builders: thread r reads rows r .. r+lags of the staged tile, splits hi / lo and writes ITS psi row
//               (no scatter to neighbouring rows, no halo special case) into an NS-deep ring of psi tiles
//   warp  12    MMA issue (whole warp converged, one elected lane issues: descriptors stay in uniform registers):
//               tile i -> accumulator i & 1 (2 x NT TMEM columns), tcgen05.commit releases the psi slot
//   warps 4-11  two epilogue groups, group g drains accumulator g: tcgen05.ld, frees the accumulator as soon as its
//               last column block is in registers, then per-state sums of squares, max, exp, stores
// ------------------------------------------------------------------------------------------------
constexpr int NTHREADS2 = 14 * 32;
constexpr int ACC_STRIDE = 256;          // TMEM columns between the two accumulators
constexpr int NXS = 4;                   // staged x tiles

__device__ __forceinline__ void mbar_expect_tx2(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

template <int DP>
__global__ void __launch_bounds__(NTHREADS2, 1) emission_tc2_kernel(const EmitTcArgs a, const int NS, const int NT, const int KT) {
  constexpr int SC = DP * 32 / gcd_c(DP, 32);
  constexpr int SPC = SC / DP;
  extern __shared__ __align__(1024) unsigned char smem[];
  const BlobHeader* hd = reinterpret_cast<const BlobHeader*>(a.blob);
  const int K = a.K, L = a.lags;               // D == DP
  const uint32_t A_LBO = TM * 16, B_LBO = (uint32_t)NT * 16;
  const size_t psi_floats = (size_t)TM * KT;
  const uint32_t xs_bytes = (uint32_t)(((TM + L) * DP * 4 + 127) & ~127);
  // smem: W_hi | W_lo | NS x (psi_hi | psi_lo) | NXS x staged x | c | barriers
  float* Bhi = reinterpret_cast<float*>(smem);
  float* Blo = Bhi + (size_t)NT * KT;
  float* psi0 = Blo + (size_t)NT * KT;
  unsigned char* xs0 = reinterpret_cast<unsigned char*>(psi0 + 2 * psi_floats * NS);
  float* csm = reinterpret_cast<float*>(xs0 + (size_t)NXS * xs_bytes);      // K + 1
  uint64_t* bars = reinterpret_cast<uint64_t*>(csm + ((K + 1 + 3) & ~3));
  // barriers: psi_full[4] | psi_empty[4] | acc_full[2] | acc_empty[2] | x_full[4] | x_empty[4]
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 20);
  const uint32_t bar0 = smem_u32(bars);
  auto psi_full = [&](int s) { return bar0 + 8u * s; };
  auto psi_empty = [&](int s) { return bar0 + 8u * (4 + s); };
  auto acc_full = [&](int g) { return bar0 + 8u * (8 + g); };
  auto acc_empty = [&](int g) { return bar0 + 8u * (10 + g); };
  auto x_full = [&](int s) { return bar0 + 8u * (12 + s); };
  auto x_empty = [&](int s) { return bar0 + 8u * (16 + s); };

  const int tid = threadIdx.x;
  const int warp = tid >> 5;
  const int lane = tid & 31;
#define EMIT2_STAMP(tile, slot)                                                                          \
  do {                                                                                                   \
    if (a.dbg && blockIdx.x == 0 && blockIdx.y == 0 && lane == 0 && (warp & 3) == 0 && (tile) < 32)      \
      g_emit_dbg[(tile) * 8 + (slot)] = emit_clock();                                                    \
  } while (0)
  {
    const int4* s0 = reinterpret_cast<const int4*>(a.blob + hd->off_W_hi);
    const int4* s1 = reinterpret_cast<const int4*>(a.blob + hd->off_W_lo);
    int4* d0 = reinterpret_cast<int4*>(Bhi);
    int4* d1 = reinterpret_cast<int4*>(Blo);
    const int n16 = NT * KT / 4;
    for (int i = tid; i < n16; i += NTHREADS2) { d0[i] = __ldg(s0 + i); d1[i] = __ldg(s1 + i); }
    const float* cg = reinterpret_cast<const float*>(a.blob + hd->off_c_f);
    for (int i = tid; i < K + 1; i += NTHREADS2) csm[i] = cg[i];
    // staged-x rows that a copy never touches (before a trial's first row) must hold finite numbers
    for (int i = tid; i < (int)(NXS * xs_bytes / 16); i += NTHREADS2) reinterpret_cast<int4*>(xs0)[i] = make_int4(0, 0, 0, 0);
  }
  if (tid == 0) {
    for (int s = 0; s < 4; ++s) {
      mbar_init(psi_full(s), 4); mbar_init(psi_empty(s), 1);
      mbar_init(x_full(s), 1); mbar_init(x_empty(s), 4);
    }
    for (int g = 0; g < 2; ++g) { mbar_init(acc_full(g), 1); mbar_init(acc_empty(g), 4); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 12) tmem_alloc<512>(smem_u32(tmem_ptr));
  fence_proxy_async();             // W tiles, zeroed staging (generic-proxy stores) -> visible to the async proxy
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  // work list of this CTA: (trial, t0) pairs in a fixed order every role walks on its own
  int trial = blockIdx.x;
  int t0 = blockIdx.y * TM;
  long long beg = 0;
  int T = 0;
  int cur_trial = -1;                // the offsets are re-read only when the walk moves to another trial
  auto advance = [&]() {
    while (trial < a.n_trials) {
      if (trial != cur_trial) {
        beg = a.offsets[trial];
        T = (int)(a.offsets[trial + 1] - beg);
        cur_trial = trial;
      }
      if (t0 < T) return true;
      trial += gridDim.x;
      t0 = blockIdx.y * TM;
    }
    return false;
  };
  auto step_tile = [&]() { t0 += gridDim.y * TM; };

  if (warp == 13) {
    // ------------------------------------------------------------ copy producer
    if (lane == 0) {
      bool have = advance();
      for (int i = 0; have; ++i) {
        const int sx = i % NXS;
        mbar_wait(x_empty(sx), (((uint32_t)(i / NXS)) & 1u) ^ 1u);
        int start = t0 - L, dst_row = 0;               // staged row j holds x_{t0 - L + j}
        if (start < 0) { dst_row = -start; start = 0; }
        const int end = t0 + TM < T ? t0 + TM : T;
        const uint32_t bytes = (uint32_t)(end - start) * (DP * 4);
        mbar_expect_tx2(x_full(sx), bytes);
        bulk_g2s(smem_u32(xs0 + (size_t)sx * xs_bytes) + dst_row * (DP * 4), a.x + (beg + start) * DP, bytes, x_full(sx));
        step_tile();
        have = advance();
      }
    }
  } else if (warp < 4) {
    // ------------------------------------------------------------ builders: thread r writes psi row r
    bool have = advance();
    for (int i = 0; have; ++i) {
      const int sx = i % NXS, s = i % NS;
      EMIT2_STAMP(i, 0);
      mbar_wait(x_full(sx), ((uint32_t)(i / NXS)) & 1u);
      mbar_wait(psi_empty(s), (((uint32_t)(i / NS)) & 1u) ^ 1u);
      EMIT2_STAMP(i, 1);
      // thread r builds ITS psi row from staged rows r + L - b (x_{t - b}), b = 0 .. L: no scatter to neighbouring rows,
      // no halo special case (splitting every x row once and scattering it to the L + 1 rows that use it executes fewer
      // conversions but measured slower, 1.28 vs 0.96 us per tile).  The loads of lag b + 1 are issued before lag b is
      // converted, so the LDS latency is off the chain.
      const float4* xrow = reinterpret_cast<const float4*>(xs0 + (size_t)sx * xs_bytes) + (size_t)(tid + L) * (DP / 4);
      unsigned char* Ahi = reinterpret_cast<unsigned char*>(psi0 + 2 * psi_floats * s) + tid * 16;
      unsigned char* Alo = Ahi + 4 * psi_floats;
      float4 v[DP / 4], vn[DP / 4];
#pragma unroll
      for (int c = 0; c < DP / 4; ++c) v[c] = xrow[c];
      for (int b = 0; b <= L; ++b) {
        if (b < L) {
#pragma unroll
          for (int c = 0; c < DP / 4; ++c) vn[c] = xrow[c - (b + 1) * (DP / 4)];
        }
#pragma unroll
        for (int c = 0; c < DP / 4; ++c) {
          const float4 h = make_float4(tf32_rna(v[c].x), tf32_rna(v[c].y), tf32_rna(v[c].z), tf32_rna(v[c].w));
          const float4 l = make_float4(tf32_rna(v[c].x - h.x), tf32_rna(v[c].y - h.y), tf32_rna(v[c].z - h.z), tf32_rna(v[c].w - h.w));
          const size_t off = (size_t)(b * (DP / 4) + c) * A_LBO;
          *reinterpret_cast<float4*>(Ahi + off) = h;
          *reinterpret_cast<float4*>(Alo + off) = l;
        }
#pragma unroll
        for (int c = 0; c < DP / 4; ++c) v[c] = vn[c];
      }
      for (int kc = (L + 1) * (DP / 4); kc < KT / 4; ++kc) {
        const float one = kc == (L + 1) * (DP / 4) ? 1.f : 0.f;      // bias column, then zero padding
        *reinterpret_cast<float4*>(Ahi + (size_t)kc * A_LBO) = make_float4(one, 0.f, 0.f, 0.f);
        *reinterpret_cast<float4*>(Alo + (size_t)kc * A_LBO) = make_float4(0.f, 0.f, 0.f, 0.f);
      }
      EMIT2_STAMP(i, 7);
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(x_empty(sx));          // this warp's rows of the staged tile are in registers / psi
        mbar_arrive(psi_full(s));
      }
      EMIT2_STAMP(i, 2);
      step_tile();
      have = advance();
    }
  } else if (warp == 12) {
    // ------------------------------------------------------------ MMA issuer
    // Descriptors differ between MMAs only in the 14-bit start-address field of their low word, so the inner loop is
    // two integer adds and the MMA; NT / KT arrive as kernel parameters and the whole warp runs the loop converged
    // (one elected lane issues), so everything stays in uniform registers.
    {
      const uint32_t idesc = make_idesc(TM, NT);
      const int ksteps = KT / 8;
      const uint32_t a_hiword = (128u >> 4) | (1u << 14), b_hiword = a_hiword;      // SBO = 128, descriptor version 1
      const uint32_t a_lbo = ((A_LBO >> 4) & 0x3FFFu) << 16, b_lbo = ((B_LBO >> 4) & 0x3FFFu) << 16;
      const uint32_t a_step = (2 * A_LBO) >> 4, b_step = (2 * B_LBO) >> 4;
      const uint32_t bhi0 = ((smem_u32(Bhi) >> 4) & 0x3FFFu) | b_lbo, blo0 = ((smem_u32(Blo) >> 4) & 0x3FFFu) | b_lbo;
      const uint32_t psi_stage = (uint32_t)(8 * psi_floats) >> 4, psi_half = (uint32_t)(4 * psi_floats) >> 4;
      const uint32_t a00 = ((smem_u32(psi0) >> 4) & 0x3FFFu) | a_lbo;
      bool have = advance();
      for (int i = 0; have; ++i) {
        const int s = i % NS, g = i & 1;
        mbar_wait(psi_full(s), ((uint32_t)(i / NS)) & 1u);
        mbar_wait(acc_empty(g), (((uint32_t)(i >> 1)) & 1u) ^ 1u);
        tc_fence_after();
        EMIT2_STAMP(i, 3);
        const uint32_t ahi0 = a00 + (uint32_t)s * psi_stage, alo0 = ahi0 + psi_half;
        const uint32_t dcol = tmem_base + (uint32_t)(g * ACC_STRIDE);
        uint32_t acc = 0u;
#pragma unroll 1
        for (int pass = 0; pass < 3; ++pass) {
          uint32_t pa = pass == 0 ? alo0 : ahi0;           // small terms first
          uint32_t pb = pass == 1 ? blo0 : bhi0;
#pragma unroll 1
          for (int k = 0; k < ksteps; ++k) {
            umma_tf32_elect(dcol, ((uint64_t)a_hiword << 32) | pa, ((uint64_t)b_hiword << 32) | pb, idesc, acc);
            acc = 1u;
            pa += a_step;
            pb += b_step;
          }
        }
        umma_commit_elect(psi_empty(s));
        umma_commit_elect(acc_full(g));
        EMIT2_STAMP(i, 4);
        step_tile();
        have = advance();
      }
    }
  } else {
    // ------------------------------------------------------------ epilogue groups
    const int g = (warp - 4) >> 2;
    const int row = ((warp & 3) << 5) + lane;
    const uint32_t tlane = (uint32_t)((warp & 3) * 32) << 16;
    bool have = advance();
    for (int i = 0; have; ++i) {
      if ((i & 1) == g) {
        const int t = t0 + row;
        float qinit = 0.f;
        if (t < L && t < T) {
          // initial segment: N(0, I) for every state (ssm mu_init = 0, Sigma_init = I)
          const float* xr = a.x + (beg + t) * DP;
          for (int d = 0; d < DP; ++d) { const float xv = __ldg(xr + d); qinit = fmaf(xv, xv, qinit); }
        }
        mbar_wait(acc_full(g), ((uint32_t)(i >> 1)) & 1u);
        tc_fence_after();
        EMIT2_STAMP(i, 5);
        const uint32_t tacc = tmem_base + tlane + (uint32_t)(g * ACC_STRIDE);
        float vmax = -INFINITY;
        float vals[32];
#pragma unroll
        for (int k = 0; k < 32; ++k) vals[k] = -INFINITY;
        constexpr int NSC = (32 + SPC - 1) / SPC;
#pragma unroll
        for (int sc = 0; sc < NSC; ++sc) {
          if (sc * SPC < K) {
            uint32_t r[SC / 32][32];
#pragma unroll
            for (int c32 = 0; c32 < SC / 32; ++c32) tmem_ld32(tacc + sc * SC + c32 * 32, r[c32]);
            tmem_ld_wait();
            if ((sc + 1) * SPC >= K) {
              // last column block of this tile is in registers: hand the accumulator back to the MMA warp
              tc_fence_before();
              __syncwarp();
              if (lane == 0) mbar_arrive(acc_empty(g));
            }
            // per-state sums of squares with packed fp32 FMAs (DP is a multiple of 4: column pairs never straddle states)
            float2 q2[SPC];
#pragma unroll
            for (int s2 = 0; s2 < SPC; ++s2) q2[s2] = make_float2(0.f, 0.f);
#pragma unroll
            for (int c32 = 0; c32 < SC / 32; ++c32)
#pragma unroll
              for (int e = 0; e < 32; e += 2) {
                const float2 y = make_float2(__uint_as_float(r[c32][e]), __uint_as_float(r[c32][e + 1]));
                q2[(c32 * 32 + e) / DP] = ffma2_tc(y, y, q2[(c32 * 32 + e) / DP]);
              }
#pragma unroll
            for (int s2 = 0; s2 < SPC; ++s2) {
              const int k = sc * SPC + s2;
              if (k < 32 && k < K) {
                vals[k] = csm[k] - 0.5f * (q2[s2].x + q2[s2].y);
                vmax = fmaxf(vmax, vals[k]);
              }
            }
          }
        }
        if (t < T) {
          const bool init = t < L;
          float* og = a.Bsc + (beg + t) * K;
          a.mx[beg + t] = init ? csm[K] - 0.5f * qinit : vmax;
          if ((K & 3) == 0) {
#pragma unroll
            for (int k = 0; k < 32; k += 4)
              if (k < K)
                *reinterpret_cast<float4*>(og + k) =
                    init ? make_float4(1.f, 1.f, 1.f, 1.f)
                         : make_float4(__expf(vals[k] - vmax), __expf(vals[k + 1] - vmax), __expf(vals[k + 2] - vmax),
                                       __expf(vals[k + 3] - vmax));
          } else {
#pragma unroll
            for (int k = 0; k < 32; ++k)
              if (k < K) og[k] = init ? 1.f : __expf(vals[k] - vmax);
          }
        }
        EMIT2_STAMP(i, 6);
      }
      step_tile();
      have = advance();
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 12) {
    tc_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}

template <int DP>
int launch2(const EmitTcArgs& a, int NT, int KT, int max_T, cudaStream_t st) {
  constexpr int SC = DP * 32 / gcd_c(DP, 32);
  constexpr int SPC = SC / DP;
  if (a.D != DP) return 1;                                        // bulk copies need 16-byte rows
  if (((a.K + SPC - 1) / SPC) * SC > ACC_STRIDE) return 1;       // epilogue super-chunks must stay inside one accumulator
  if (SC > 96) return 1;                                         // 128 registers per thread at 14 warps: DP = 20 / 28 column blocks spill
  const size_t xs_bytes = (((size_t)(TM + a.lags) * DP * 4 + 127) & ~(size_t)127);
  const size_t fixed = 4 * (2 * (size_t)NT * KT + ((a.K + 1 + 3) & ~3)) + NXS * xs_bytes + 20 * 8 + 64;
  const size_t per_stage = 4 * 2 * (size_t)TM * KT;
  int NS = 4;
  while (NS >= 2 && fixed + NS * per_stage > 227 * 1024) --NS;
  if (NS < 2) return 1;
  size_t smem = fixed + NS * per_stage;
  if (smem < 120 * 1024) smem = 120 * 1024;                // one CTA per SM (it allocates all 512 TMEM columns)
  auto kern = emission_tc2_kernel<DP>;
  static bool configured = false;
  if (!configured) {
    BN_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    configured = true;
  }
  const int slots = 148;
  int gx = a.n_trials < slots ? a.n_trials : slots;
  int tiles = bn_cdiv(max_T, TM);
  int gy = a.n_trials >= slots ? 1 : bn_cdiv(slots, a.n_trials);
  if (gy > tiles) gy = tiles;
  kern<<<dim3(gx, gy), NTHREADS2, smem, st>>>(a, NS, NT, KT);
  BN_LAUNCHED();
  return 0;
}

template <int DP>
int launch(const EmitTcArgs& a, int NT, int KT, int max_T, cudaStream_t st) {
  constexpr int SC = DP * 32 / gcd_c(DP, 32);
  constexpr int SPC = SC / DP;
  if (((a.K + SPC - 1) / SPC) * SC > 256) return 1;       // epilogue super-chunks must stay inside the allocation
  size_t smem = 4 * (2 * (size_t)NT * KT + 2 * (size_t)TM * KT + ((a.K + 1 + 3) & ~3)) + 64;
  if (smem > 113 * 1024) return 1;                        // two CTAs per SM
  auto kern = emission_tc_kernel<DP>;
  static bool configured = false;
  if (!configured) {
    BN_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 113 * 1024));
    configured = true;
  }
  const int slots = 2 * 148;
  int gx = a.n_trials < slots ? a.n_trials : slots;
  int tiles = bn_cdiv(max_T, TM);
  int gy = a.n_trials >= slots ? 1 : bn_cdiv(slots, a.n_trials);
  if (gy > tiles) gy = tiles;
  kern<<<dim3(gx, gy), NTHREADS, smem, st>>>(a);
  BN_LAUNCHED();
  return 0;
}

}  // namespace

int bn_launch_emission_tc(const unsigned char* d_blob, const float* d_x, const long long* d_offsets, int K, int D,
                          int lags, int n_trials, int max_T, float* d_Bsc, float* d_mx, cudaStream_t st) {
  const BlobHeader h = bn_blob_layout(K, D, lags);
  if (h.NT > 256 || h.NT < 16) return 1;
  if (((uintptr_t)d_x & 15) != 0) return 1;
  EmitTcArgs a;
  a.blob = d_blob; a.x = d_x; a.offsets = d_offsets; a.K = K; a.D = D; a.lags = lags; a.n_trials = n_trials;
  a.Bsc = d_Bsc; a.mx = d_mx;
  static const int dbg = [] { const char* e = getenv("BN_EMIT_DBG"); return e ? atoi(e) : 0; }();
  a.dbg = dbg;
  static const bool one_stage = [] { const char* e = getenv("BN_EMIT"); return e && e[0] == '1'; }();
  if (!one_stage) {
    int r = 1;
    switch (h.DP) {
      case 4: r = launch2<4>(a, h.NT, h.KT, max_T, st); break;
      case 8: r = launch2<8>(a, h.NT, h.KT, max_T, st); break;
      case 12: r = launch2<12>(a, h.NT, h.KT, max_T, st); break;
      case 16: r = launch2<16>(a, h.NT, h.KT, max_T, st); break;
      case 20: r = launch2<20>(a, h.NT, h.KT, max_T, st); break;
      case 24: r = launch2<24>(a, h.NT, h.KT, max_T, st); break;
      case 28: r = launch2<28>(a, h.NT, h.KT, max_T, st); break;
      case 32: r = launch2<32>(a, h.NT, h.KT, max_T, st); break;
      default: return 1;
    }
    if (r <= 0) return r;
  }
  switch (h.DP) {
    case 4: return launch<4>(a, h.NT, h.KT, max_T, st);
    case 8: return launch<8>(a, h.NT, h.KT, max_T, st);
    case 12: return launch<12>(a, h.NT, h.KT, max_T, st);
    case 16: return launch<16>(a, h.NT, h.KT, max_T, st);
    case 20: return launch<20>(a, h.NT, h.KT, max_T, st);
    case 24: return launch<24>(a, h.NT, h.KT, max_T, st);
    case 28: return launch<28>(a, h.NT, h.KT, max_T, st);
    case 32: return launch<32>(a, h.NT, h.KT, max_T, st);
    default: return 1;
  }
}

extern "C" int bn_debug_emit_times(long long* h_out) {
  BN_CUDA(cudaMemcpyFromSymbol(h_out, g_emit_dbg, sizeof(long long) * 256));
  return 0;
}
