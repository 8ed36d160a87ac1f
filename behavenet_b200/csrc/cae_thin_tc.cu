// tcgen05 weight gradient of the thin first / last layers (kernel 5, stride 2, 1-4 image channels <-> a
// 32k-channel feature map), TF32 operands with fp32 accumulation in TMEM.
//
//   dW[(tap, cb), cs] = sum over small pixels p of  thin[2p + tap - pad, cb] * fat[p, cs]
//
// is a GEMM whose reduction index is the pixel: D[128 x 32] (TMEM) += A[128 x 8] * B[8 x 32] per MMA with
//   A (K-major, SWIZZLE_128B): row r = (tap, cb) (25*CB <= 100 rows are real), k = small pixel.  The CUDA
//     cores build it from the staged thin-image patch: a warp writes 32 consecutive pixels of one row,
//     i.e. one whole 128-byte swizzled line, conflict-free.
//   B (MN-major, SWIZZLE_128B_BASE32B): k = small pixel, n = channel -- exactly the NHWC pixel rows of the
//     fat image, so ONE tiled TMA copy per 4 x 32-pixel tile stages it with no data movement by threads.
// The fat image (134 MB for config C2) is read once at HBM rate and never passes through registers;
// the FP32 kernel it replaces (thin_wgrad_kernel, cae_thin.cu) needs 25*CB FMAs per pixel and channel and
// is issue-bound at a quarter of that rate.  Persistent CTAs accumulate their whole tile sequence into one
// TMEM accumulator and write a single [(tap, cb)][cs] slice for the batched reduction kernel.
#include <cuda.h>

#include <stdlib.h>
#include <string.h>

#include <mutex>
#include <vector>

#include "cae_kernels.cuh"
#include "tc_common.cuh"

namespace {

using namespace bn_tc;

constexpr int WT_H = 4, WT_W = 32;                 // small-pixel tile: 4 rows x 32 columns = 4 k-chunks of 32 pixels
constexpr int WP_ROWS = 2 * (WT_H - 1) + 5;        // 11 patch rows
constexpr int WP_COLS = 72;                        // 2*(WT_W-1)+5 = 67 used


struct ThinWgArgs {
  ImgView thin;            // thin image (C <= 4), any strides
  int Hs, Ws, Cs;          // fat image (NHWC dense)
  int pt, pl, n;
  int tiles_x, tiles_per_frame;
  long long total_tiles;
  float* partial;          // [gridDim.x][25*CB][Cs]
};

template <int CB>
struct ThinWgSmem {
  static constexpr int ROWS = 25 * CB;
  static constexpr int A_CHUNK = ((ROWS + 7) / 8) * 1024;            // one 32-pixel k-chunk: 8-row swizzle atoms
  static constexpr int A_BUF = WT_H * A_CHUNK;
  static constexpr int B_BUF = WT_H * WT_W * 128;                    // 4 x 32 pixels x 32 channels fp32 = 16 KB
  static constexpr int PATCH = CB * WP_ROWS * WP_COLS * 4;
  static constexpr int OFF_B = 2 * A_BUF;                            // multiple of 1024
  static constexpr int OFF_PATCH = OFF_B + 3 * B_BUF;
  static constexpr int OFF_BAR = OFF_PATCH + ((2 * PATCH + 15) & ~15);
  // the M = 128 MMA reads 16 KB per chunk whatever ROWS is: rows >= ROWS only feed accumulator rows that
  // are never read, but their addresses must stay inside the CTA's shared-memory window
  static constexpr int TOTAL_MIN = OFF_BAR + 96;
  static constexpr int A_OVERREAD_END = A_BUF + (WT_H - 1) * A_CHUNK + 16384;
  static constexpr int TOTAL = TOTAL_MIN > A_OVERREAD_END ? TOTAL_MIN : A_OVERREAD_END;
};

__device__ __forceinline__ void cp_async4z(float* dst, const float* src, bool ok) {
  const uint32_t d = (uint32_t)__cvta_generic_to_shared(dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(d), "l"(src), "r"(ok ? 4u : 0u) : "memory");
}
__device__ __forceinline__ uint64_t desc_k_sw128(uint32_t saddr) {            // K-major SWIZZLE_128B, SBO = 1024
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
__device__ __forceinline__ uint64_t desc_mn_sw128(uint32_t saddr, uint32_t lbo, uint32_t sbo) {   // MN-major, 32 B atoms
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)1 << 61;
  return d;
}
__device__ __forceinline__ void mbar_expect_tx2(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_tile_4d2(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c, int w, int h,
                                             int n) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c), "r"(w), "r"(h), "r"(n)
      : "memory");
}

// Builder warps (0-7) and one control warp (8) run the tile sequence decoupled by mbarriers: the builders stage
// the next thin patch (cp.async), gather the im2col rows of the current one into A[buf] and arrive on
// a_full[buf]; lane 0 of the control warp streams the fat tiles (TMA, 3 stages), issues the 16 MMAs of a tile as
// soon as its A and B have landed and commits them to ab_free, which hands the A buffer back to the builders
// and the B stage back to the TMA.  No thread ever waits for the single-thread MMA issue inside a barrier.
constexpr int WBUILD = 256;                       // builder threads
constexpr int WB_STAGES = 3;

template <int CB, int T0, int T1>
__device__ __forceinline__ void build_rows(unsigned char* __restrict__ abase, const float* __restrict__ prow0,
                                           const uint32_t (&swz)[8], int pcol) {
  // rows (tap, cb), tap in [T0, T1): every offset below is a compile-time constant except the swizzled column
  float v[(T1 - T0) * CB];
#pragma unroll
  for (int tap = T0; tap < T1; ++tap)
#pragma unroll
    for (int cb = 0; cb < CB; ++cb)
      v[(tap - T0) * CB + cb] = prow0[(cb * WP_ROWS + tap / 5) * WP_COLS + tap % 5];
#pragma unroll
  for (int tap = T0; tap < T1; ++tap)
#pragma unroll
    for (int cb = 0; cb < CB; ++cb) {
      const int r = tap * CB + cb;
      *reinterpret_cast<float*>(abase + (r >> 3) * 1024 + (r & 7) * 128 + swz[r & 7]) = v[(tap - T0) * CB + cb];
    }
  (void)pcol;
}

template <int CB>
__global__ void __launch_bounds__(WBUILD + 32, 2) thin_wgrad_tc_kernel(const __grid_constant__ CUtensorMap fat_map,
                                                                      const ThinWgArgs a) {
  bn_pdl_trigger();
  using S = ThinWgSmem<CB>;
  constexpr int ROWS = S::ROWS;
  extern __shared__ __align__(1024) unsigned char sm[];
  float* patch[2] = {reinterpret_cast<float*>(sm + S::OFF_PATCH), reinterpret_cast<float*>(sm + S::OFF_PATCH + S::PATCH)};
  uint64_t* b_full = reinterpret_cast<uint64_t*>(sm + S::OFF_BAR);    // [WB_STAGES] TMA landed
  uint64_t* a_full = b_full + WB_STAGES;                              // [2] builders done (one arrival per warp)
  uint64_t* ab_free = a_full + 2;                                     // [2] MMAs of the tile that used A[s] retired
  uint64_t* done_bar = ab_free + 2;                                   // [1]
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(done_bar + 1);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    for (int s = 0; s < WB_STAGES; ++s) mbar_init(smem_u32(b_full + s), 1);
    for (int s = 0; s < 2; ++s) {
      mbar_init(smem_u32(a_full + s), WBUILD / 32);
      mbar_init(smem_u32(ab_free + s), 1);
    }
    mbar_init(smem_u32(done_bar), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 8) tmem_alloc<32>(smem_u32(tmem_ptr));
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  const uint32_t sbase = smem_u32(sm);
  const int c0 = blockIdx.y * 32;                  // channel group of the fat image
  bn_pdl_wait();

  auto decode = [&](long long t, int& f, int& y0, int& x0) {
    f = (int)(t / a.tiles_per_frame);
    const int tt = (int)(t - (long long)f * a.tiles_per_frame);
    const int ty = tt / a.tiles_x;
    y0 = ty * WT_H;
    x0 = (tt - ty * a.tiles_x) * WT_W;
  };
  const long long total = a.total_tiles;
  int n_my = 0;
  for (long long t = blockIdx.x; t < total; t += gridDim.x) ++n_my;

  if (warp < 8) {
    // ======================= builders ============================================================
    auto issue_patch = [&](float* dst, long long t) {
      int f, y0, x0;
      decode(t, f, y0, x0);
      for (int i = tid; i < CB * WP_ROWS * WP_COLS; i += WBUILD) {
        const int c = i / (WP_ROWS * WP_COLS);
        const int rem = i - c * WP_ROWS * WP_COLS;
        const int r = rem / WP_COLS, col = rem - r * WP_COLS;
        const int y = 2 * y0 - a.pt + r, x = 2 * x0 - a.pl + col;
        const bool ok = (unsigned)y < (unsigned)a.thin.H && (unsigned)x < (unsigned)a.thin.W;
        const long long off = ok ? (long long)f * a.thin.sn + (long long)y * a.thin.sy + (long long)x * a.thin.sx +
                                       (long long)c * a.thin.sc
                                 : 0;
        cp_async4z(dst + i, a.thin.p + off, ok);
      }
    };
    const int prow = (tid & 127) >> 5, pcol = lane;              // this thread's pixel of the 4 x 32 tile
    const int half = tid >> 7;                                   // taps 0..12 / 13..24
    uint32_t swz[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) swz[q] = (uint32_t)((((pcol >> 2) ^ q) << 4) + (pcol & 3) * 4);
    long long t = blockIdx.x;
    if (t < total) issue_patch(patch[0], t);
    asm volatile("cp.async.commit_group;" ::: "memory");
    for (int it = 0; it < n_my; ++it, t += gridDim.x) {
      const int buf = it & 1;
      asm volatile("cp.async.wait_group 0;" ::: "memory");
      asm volatile("bar.sync 1, 256;" ::: "memory");             // patch[buf] complete; everyone left patch[buf ^ 1]
      if (it + 1 < n_my) issue_patch(patch[buf ^ 1], t + gridDim.x);
      asm volatile("cp.async.commit_group;" ::: "memory");
      if (it >= 2) mbar_wait(smem_u32(ab_free + buf), ((it - 2) >> 1) & 1);     // MMAs of tile it - 2 retired
      unsigned char* abase = sm + buf * S::A_BUF + prow * S::A_CHUNK;
      const float* prow0 = patch[buf] + (2 * prow) * WP_COLS + 2 * pcol;
      if (half == 0) build_rows<CB, 0, 13>(abase, prow0, swz, pcol);
      else build_rows<CB, 13, 25>(abase, prow0, swz, pcol);
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(a_full + buf));
    }
  } else {
    // ======================= control warp: fat-tile TMA + MMA issue ==============================
    if (lane == 0) {
      auto issue_fat = [&](int j) {                              // j-th tile of this CTA
        int f, y0, x0;
        decode(blockIdx.x + (long long)j * gridDim.x, f, y0, x0);
        const int st = j % WB_STAGES;
        const uint32_t bar = smem_u32(b_full + st);
        mbar_expect_tx2(bar, (uint32_t)S::B_BUF);
        tma_tile_4d2(sbase + S::OFF_B + st * S::B_BUF, &fat_map, bar, c0, x0, y0, f);
      };
      for (int j = 0; j < WB_STAGES - 1 && j < n_my; ++j) issue_fat(j);
      constexpr uint32_t idesc = make_idesc(128, 32) | (1u << 16);          // A K-major, B MN-major
      const uint64_t ad0 = desc_k_sw128(sbase), bd0 = desc_mn_sw128(sbase + S::OFF_B, 4096, 512);
      for (int it = 0; it < n_my; ++it) {
        const int buf = it & 1, st = it % WB_STAGES;
        // stage (it + 2) % 3 was read by the MMAs of tile it - 1
        if (it + WB_STAGES - 1 < n_my) {
          if (it >= 1) mbar_wait(smem_u32(ab_free + (buf ^ 1)), ((it - 1) >> 1) & 1);
          issue_fat(it + WB_STAGES - 1);
        }
        mbar_wait(smem_u32(a_full + buf), (it >> 1) & 1);
        mbar_wait(smem_u32(b_full + st), (it / WB_STAGES) & 1);
        tc_fence_after();
        const uint64_t ad = ad0 + (uint64_t)((buf * S::A_BUF) >> 4), bd = bd0 + (uint64_t)((st * S::B_BUF) >> 4);
#pragma unroll
        for (int kc = 0; kc < WT_H; ++kc)
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_tf32(tmem_base, ad + (uint64_t)((kc * S::A_CHUNK + k * 32) >> 4), bd + (uint64_t)((kc * 4096 + k * 1024) >> 4),
                      idesc, (it | kc | k) != 0 ? 1u : 0u);
        umma_commit(smem_u32(ab_free + buf));
      }
      umma_commit(smem_u32(done_bar));
    }
    __syncwarp();
  }
  // ---- epilogue: this CTA's slice [(tap, cb)][cs] of the partial buffer
  __syncthreads();
  if (n_my > 0) {
    mbar_wait(smem_u32(done_bar), 0);
    tc_fence_after();
  }
  if (warp < 4) {
    uint32_t r[32];
    if (n_my > 0) {
      tmem_ld32(tmem_base + ((uint32_t)(warp * 32) << 16), r);
      tmem_ld_wait();
    } else {
#pragma unroll
      for (int q = 0; q < 32; ++q) r[q] = 0u;
    }
    const int row = warp * 32 + lane;
    const long long idx = row < ROWS ? ((long long)blockIdx.x * ROWS + row) * a.Cs + c0 : -1;
    warp_store_rows32(a.partial, nullptr, BN_LEAK, idx, r, nullptr, BN_ACT_NONE, reinterpret_cast<float*>(sm) + warp * 1024,
                      lane);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 8) {
    tc_fence_after();
    tmem_dealloc<32>(tmem_base);
  }
}

// ---- host side --------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn g_enc = nullptr;
int g_state = 0;
std::mutex g_mu;

bool have_tma() {
  std::lock_guard<std::mutex> lk(g_mu);
  if (g_state == 0) {
    const char* env = getenv("BN_TMA");
    void* f = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (!(env && env[0] == '0') &&
        cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) == cudaSuccess && f &&
        q == cudaDriverEntryPointSuccess) {
      g_enc = (EncodeTiledFn)f;
      g_state = 1;
    } else {
      g_state = -1;
      cudaGetLastError();
    }
  }
  return g_state == 1;
}

struct MapKey {
  const void* p;
  int n, H, W, C;
  bool operator==(const MapKey& o) const { return p == o.p && n == o.n && H == o.H && W == o.W && C == o.C; }
};
std::vector<std::pair<MapKey, CUtensorMap>> g_maps;

// NHWC fp32 image as a 4-D tensor (C, W, H, N); box = 32 channels x 32 pixels x 4 rows, 128-byte swizzle with
// 32-byte atoms (the MN-major canonical layout of a tf32 operand)
bool fat_map(const float* p, int n, int H, int W, int C, CUtensorMap* out) {
  MapKey key{p, n, H, W, C};
  std::lock_guard<std::mutex> lk(g_mu);
  for (auto& kv : g_maps)
    if (kv.first == key) { *out = kv.second; return true; }
  cuuint64_t gdim[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)n};
  cuuint64_t gstr[3] = {(cuuint64_t)C * 4, (cuuint64_t)W * C * 4, (cuuint64_t)H * W * C * 4};
  cuuint32_t box[4] = {32, (cuuint32_t)WT_W, (cuuint32_t)WT_H, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUtensorMap m;
  if (g_enc(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, (void*)p, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
    return false;
  if (g_maps.size() >= 64) g_maps.clear();
  g_maps.emplace_back(key, m);
  *out = m;
  return true;
}

template <int CB>
int launch(const CUtensorMap& map, const ThinWgArgs& a, int blocks, int cgroups, cudaStream_t st) {
  using S = ThinWgSmem<CB>;
  auto kern = thin_wgrad_tc_kernel<CB>;
  static bool configured = false;
  if (!configured) {
    BN_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, S::TOTAL));
    configured = true;
  }
  BN_CUDA(bn_launch(kern, dim3(blocks, cgroups), WBUILD + 32, S::TOTAL, st, map, a));
  BN_LAUNCHED();
  return 0;
}

template <int CB>
int resident_ctas() {
  const int per = (227 * 1024) / (ThinWgSmem<CB>::TOTAL + 1024);
  return per > 3 ? 3 : (per < 1 ? 1 : per);
}


// ------------------------------------------------------------------------------------------------
// thin image -> fat image (first encoder layer forward; last decoder layer backward-data), second form.
// Same arithmetic as thin_fprop_tc_kernel (cae_thin.cu): every output pixel's 25*CB patch values are one row of
// a K-major A tile, the TF32 weights sit in a resident B tile, an 8 x 32-pixel tile is two M = 128 MMAs per
// k-chunk.  There the 256 threads of a CTA gather, wait for the MMAs and run the epilogue in turn; here the
// three jobs belong to different warps and overlap: eight builder warps (patch staging + im2col gather into a
// double-buffered A), one control warp (MMA issue into a double-buffered accumulator) and eight epilogue warps
// (TMEM -> bias / LeakyReLU / derivative mask -> 128-byte row stores, fused column sums).
// ------------------------------------------------------------------------------------------------
constexpr int FT_H = 8, FT_W = 32;                       // small-pixel tile
constexpr int FP_ROWS = 2 * (FT_H - 1) + 5;              // 19 patch rows
constexpr int FP_COLS = 72;
constexpr int FW_THREADS = 17 * 32;                      // warps 0-7 builders, 8-15 epilogue, 16 control

struct ThinFwArgs {
  ImgView thin;
  const unsigned char* thin8;      // raw 0..255 video with the strides of `thin`, or NULL
  int Hs, Ws, Cs, pt, pl, n;
  const float* wft;                // [c_small][(tap, c_big)] TF32-rounded
  const float* bias;
  float* out;
  const float* dact;
  int act;
  float* colsum;
  int tiles_x, tiles_per_frame;
  long long total_tiles;
};

template <int CB>
struct ThinFwSmem {
  static constexpr int NKC = (25 * CB + 31) / 32;
  static constexpr int NBUF = NKC <= 2 ? 2 : 1;                       // A buffers that fit next to the rest
  static constexpr int A_TILE = 2 * NKC * 16384;
  static constexpr int OFF_B = NBUF * A_TILE;
  static constexpr int B_BYTES = NKC * 4096;
  static constexpr int OFF_PATCH = OFF_B + B_BYTES;
  static constexpr int PATCH = CB * FP_ROWS * FP_COLS * 4;
  static constexpr int OFF_EPI = OFF_PATCH + ((2 * PATCH + 1023) & ~1023);
  static constexpr int OFF_BAR = OFF_EPI + 8 * 4096;
  static constexpr int TOTAL = OFF_BAR + 128;
};

template <int CB>
__global__ void __launch_bounds__(FW_THREADS, 1) thin_fprop_tc2_kernel(const ThinFwArgs a) {
  bn_pdl_trigger();
  using S = ThinFwSmem<CB>;
  constexpr int NKC = S::NKC, KTOT = 25 * CB, NBUF = S::NBUF;
  extern __shared__ __align__(1024) unsigned char sm[];
  float* patch[2] = {reinterpret_cast<float*>(sm + S::OFF_PATCH), reinterpret_cast<float*>(sm + S::OFF_PATCH + S::PATCH)};
  uint64_t* a_full = reinterpret_cast<uint64_t*>(sm + S::OFF_BAR);     // [2] builders done (8 warp arrivals)
  uint64_t* a_free = a_full + 2;                                       // [2] MMAs that read A[s] retired
  uint64_t* acc_full = a_free + 2;                                     // [2]
  uint64_t* acc_free = acc_full + 2;                                   // [2] 8 epilogue-warp arrivals
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(acc_free + 2);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int c0 = blockIdx.y * 32;                                      // channel group of this CTA
  if (tid == 0) {
    for (int s = 0; s < 2; ++s) {
      mbar_init(smem_u32(a_full + s), 8);
      mbar_init(smem_u32(a_free + s), 1);
      mbar_init(smem_u32(acc_full + s), 1);
      mbar_init(smem_u32(acc_free + s), 8);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 16) tmem_alloc<128>(smem_u32(tmem_ptr));
  bn_pdl_wait();
  // resident B tile: rows n = channel, k-chunk kc, 16-byte chunk c at position c ^ (n & 7)
  for (int i = tid; i < NKC * 32 * 8; i += FW_THREADS) {
    const int kc = i / 256, rem = i - kc * 256;
    const int n = rem >> 3, c = rem & 7;
    float v[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int k = kc * 32 + c * 4 + e;
      v[e] = k < KTOT ? __ldg(a.wft + (long long)(c0 + n) * KTOT + k) : 0.f;
    }
    *reinterpret_cast<float4*>(sm + S::OFF_B + kc * 4096 + n * 128 + ((c ^ (n & 7)) << 4)) = make_float4(v[0], v[1], v[2], v[3]);
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  const uint32_t sbase = smem_u32(sm);
  const long long total = a.total_tiles;
  auto decode = [&](long long t, int& f, int& y0, int& x0) {
    f = (int)(t / a.tiles_per_frame);
    const int tt = (int)(t - (long long)f * a.tiles_per_frame);
    const int ty = tt / a.tiles_x;
    y0 = ty * FT_H;
    x0 = (tt - ty * a.tiles_x) * FT_W;
  };
  int n_my = 0;
  for (long long t = blockIdx.x; t < total; t += gridDim.x) ++n_my;

  if (warp < 8) {
    // ======================= builders ============================================================
    auto issue_patch = [&](float* dst, long long t) {
      int f, y0, x0;
      decode(t, f, y0, x0);
      for (int i = tid; i < CB * FP_ROWS * FP_COLS; i += 256) {
        const int c = i / (FP_ROWS * FP_COLS);
        const int rem = i - c * FP_ROWS * FP_COLS;
        const int r = rem / FP_COLS, col = rem - r * FP_COLS;
        const int y = 2 * y0 - a.pt + r, x = 2 * x0 - a.pl + col;
        const bool ok = (unsigned)y < (unsigned)a.thin.H && (unsigned)x < (unsigned)a.thin.W;
        const long long off = ok ? (long long)f * a.thin.sn + (long long)y * a.thin.sy + (long long)x * a.thin.sx +
                                       (long long)c * a.thin.sc
                                 : 0;
        if (a.thin8 != nullptr) dst[i] = ok ? __fdiv_rn((float)__ldg(a.thin8 + off), 255.f) : 0.f;   // numpy's float32(u8) / 255
        else cp_async4z(dst + i, a.thin.p + off, ok);
      }
    };
    const int trow = tid >> 5, tcol = lane;                  // this thread's output pixel of the 8 x 32 tile
    const int mt = trow >> 2, arow = (trow & 3) * 32 + tcol; // M-tile and row inside it
    long long t = blockIdx.x;
    if (n_my > 0) issue_patch(patch[0], t);
    asm volatile("cp.async.commit_group;" ::: "memory");
    for (int it = 0; it < n_my; ++it, t += gridDim.x) {
      const int pbuf = it & 1, abuf = NBUF == 2 ? (it & 1) : 0;
      asm volatile("cp.async.wait_group 0;" ::: "memory");
      asm volatile("bar.sync 1, 256;" ::: "memory");          // patch[pbuf] complete; everyone left patch[pbuf ^ 1]
      if (it + 1 < n_my) issue_patch(patch[pbuf ^ 1], t + gridDim.x);
      asm volatile("cp.async.commit_group;" ::: "memory");
      if (it >= NBUF) mbar_wait(smem_u32(a_free + abuf), ((it - NBUF) / NBUF) & 1);
      // this thread's im2col row: k = (ky*5 + kx)*CB + cb  <-  patch[cb][2*trow + ky][2*tcol + kx]
      const float* p0 = patch[pbuf] + (2 * trow) * FP_COLS + 2 * tcol;
      unsigned char* abase = sm + abuf * S::A_TILE + (mt * NKC) * 16384 + arow * 128;
#pragma unroll
      for (int kc = 0; kc < NKC; ++kc) {
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          float v[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int k = kc * 32 + c * 4 + e;                // compile-time
            if (k < KTOT) {
              const int tap = k / CB, cb = k - tap * CB;
              v[e] = p0[(cb * FP_ROWS + tap / 5) * FP_COLS + tap % 5];
            } else {
              v[e] = 0.f;
            }
          }
          *reinterpret_cast<float4*>(abase + kc * 16384 + ((c ^ (arow & 7)) << 4)) = make_float4(v[0], v[1], v[2], v[3]);
        }
      }
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(a_full + abuf));
    }
  } else if (warp == 16) {
    // ======================= control warp: MMA issue =============================================
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc(128, 32);
      const uint64_t ad0 = desc_k_sw128(sbase), bd0 = desc_k_sw128(sbase + S::OFF_B);
      for (int it = 0; it < n_my; ++it) {
        const int abuf = NBUF == 2 ? (it & 1) : 0, cbuf = it & 1;
        if (it >= 2) {
          mbar_wait(smem_u32(acc_free + cbuf), ((it >> 1) - 1) & 1);
          tc_fence_after();
        }
        mbar_wait(smem_u32(a_full + abuf), (it / NBUF) & 1);
        tc_fence_after();
#pragma unroll
        for (int m = 0; m < 2; ++m)
#pragma unroll
          for (int kc = 0; kc < NKC; ++kc)
#pragma unroll
            for (int k = 0; k < 4; ++k)
              umma_tf32(tmem_base + cbuf * 64 + m * 32,
                        ad0 + (uint64_t)((abuf * S::A_TILE + (m * NKC + kc) * 16384 + k * 32) >> 4),
                        bd0 + (uint64_t)((kc * 4096 + k * 32) >> 4), idesc, (kc | k) != 0 ? 1u : 0u);
        umma_commit(smem_u32(a_free + abuf));
        umma_commit(smem_u32(acc_full + cbuf));
      }
    }
    __syncwarp();
  } else {
    // ======================= epilogue warps ======================================================
    const int e = warp - 8, q = e & 3, m = e >> 2;           // TMEM lane quarter (= warp % 4) and M-tile
    float* etile = reinterpret_cast<float*>(sm + S::OFF_EPI) + e * 1024;
    float4 cacc = make_float4(0.f, 0.f, 0.f, 0.f);
    long long t = blockIdx.x;
    for (int it = 0; it < n_my; ++it, t += gridDim.x) {
      const int cbuf = it & 1;
      int f, y0, x0;
      decode(t, f, y0, x0);
      mbar_wait(smem_u32(acc_full + cbuf), (it >> 1) & 1);
      tc_fence_after();
      uint32_t r[32];
      tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + cbuf * 64 + m * 32, r);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(acc_free + cbuf));
      const int oy = y0 + m * 4 + q, ox = x0 + lane;
      const bool valid = oy < a.Hs && ox < a.Ws;
      const long long idx = valid ? (((long long)f * a.Hs + oy) * a.Ws + ox) * a.Cs + c0 : -1;
      warp_store_rows32(a.out, a.dact, BN_LEAK, idx, r, a.bias ? a.bias + c0 : nullptr, a.act, etile, lane, cacc,
                        a.colsum != nullptr);
    }
    if (a.colsum != nullptr) {
      // 8 warps x 32 columns -> one atomic per column and CTA
      float* red = reinterpret_cast<float*>(sm + S::OFF_EPI);
#pragma unroll
      for (int o = 8; o <= 16; o <<= 1) {
        cacc.x += __shfl_xor_sync(0xffffffffu, cacc.x, o);
        cacc.y += __shfl_xor_sync(0xffffffffu, cacc.y, o);
        cacc.z += __shfl_xor_sync(0xffffffffu, cacc.z, o);
        cacc.w += __shfl_xor_sync(0xffffffffu, cacc.w, o);
      }
      asm volatile("bar.sync 2, 256;" ::: "memory");          // every epilogue warp is done with its tile
      if (lane < 8) *(reinterpret_cast<float4*>(red + e * 32) + lane) = cacc;
      asm volatile("bar.sync 2, 256;" ::: "memory");
      if (e == 0) {
        float sum = 0.f;
#pragma unroll
        for (int w8 = 0; w8 < 8; ++w8) sum += red[w8 * 32 + lane];
        atomicAdd(a.colsum + c0 + lane, sum);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 16) {
    tc_fence_after();
    tmem_dealloc<128>(tmem_base);
  }
}

template <int CB>
int launch_fw(const ThinFwArgs& a, int cgroups, cudaStream_t st) {
  using S = ThinFwSmem<CB>;
  static_assert(S::TOTAL <= 227 * 1024, "shared memory per CTA");
  auto kern = thin_fprop_tc2_kernel<CB>;
  static bool configured = false;
  if (!configured) {
    BN_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, S::TOTAL));
    configured = true;
  }
  long long blocks = 148 / cgroups;
  if (blocks < 1) blocks = 1;
  if (blocks > a.total_tiles) blocks = a.total_tiles;
  BN_CUDA(bn_launch(kern, dim3((unsigned)blocks, cgroups), FW_THREADS, S::TOTAL, st, a));
  BN_LAUNCHED();
  return 0;
}

// ------------------------------------------------------------------------------------------------
// Last decoder layer forward (ConvTranspose2d k5 s2 into 1-4 image channels + sigmoid + fused reconstruction
// loss, aes.py:463-470 + losses.py:36-96) as ONE small GEMM plus a col2im epilogue.
//
// With so few output channels the transposed convolution is cheapest the other way round: first
//     P[(cb, tap), p] = sum_ci  W[ci][cb][tap] * small[p, ci]          (p = small pixel)
// for ALL 25 taps at once -- D[128 x 256] (TMEM) = A[128 x 32] * B[256 x 32]^T with the 25*CB weight rows as
// the M operand and 14 x 18 small pixels (one tiled TMA copy, zero fill outside the image, K-major rows of 32
// channels) as the N operand, 4 MMAs per tile, every small pixel read from shared memory ONCE -- and then
//     out[Y, X, cb] = sum over the <= 9 taps with the right parities of  P[(cb, tap), ((Y + pt - ky) / 2, (X + pl - kx) / 2)]
// in the epilogue (768 output pixels per tile, ~6 shared-memory reads each), followed by bias, sigmoid, the
// per-chunk sum of squared errors and dL/d(pre-sigmoid).  The FP32 kernel it replaces (thin_dgrad5_kernel)
// spends 800*CB FMAs plus 34 LDS.128 per 100 FMAs per output quad and runs at a fifth of the HBM rate.
// Used by the fused training pass only (bn_cae_decode with a target and no x_hat output): the reconstruction a
// caller ASKS for is still computed by the fp32 kernel, so the 1e-4 reconstruction contract is untouched.
// ------------------------------------------------------------------------------------------------
constexpr int OT_H = 24, OT_W = 32;                 // output tile
constexpr int ST_H = OT_H / 2 + 2, ST_W = OT_W / 2 + 2;   // 14 x 18 small pixels (one-pixel halo each side)
constexpr int ST_PIX = ST_H * ST_W;                 // 252 (MMA N = 256)
constexpr int P_STRIDE = 260;                       // floats per P row: conflict-free float4 stores per lane
constexpr int DG_EPI_WARPS = 8;                     // epilogue warps 0-7; warp 8 TMA producer, warp 9 MMA issuer
constexpr int DG_THREADS = (DG_EPI_WARPS + 2) * 32;
constexpr int DG_STAGES = 3;
constexpr int DG_BSTAGE = 256 * 128;                // 32 KB: 256 rows x 32 channels (rows >= 252 never written)

struct ThinDgArgs {
  int Hs, Ws, Hb, Wb, pt, pl, n;
  const float* wdt;        // [cb][tap][32] TF32-rounded
  const float* bias;
  float* xhat_ws;          // (n, CB, Hb, Wb)
  const float* target;     // (n, CB, Hb, Wb)
  const float* mask;       // same or NULL
  int chunk_size, frame_offset, n_total;
  float coef;
  double* sse;
  float* dpre;             // (n, Hb, Wb, CB)
  int tiles_x, tiles_per_frame;
  long long total_tiles;
};

template <int CB>
struct ThinDgSmem {
  static constexpr int OFF_B = 16384;                                   // A: 128 rows x 128 B
  static constexpr int OFF_P = OFF_B + DG_STAGES * DG_BSTAGE;
  static constexpr int P_BYTES = 25 * CB * P_STRIDE * 4;
  static constexpr int OFF_BAR = OFF_P + P_BYTES;
  static constexpr int TOTAL = OFF_BAR + 8 * (2 * DG_STAGES + 4) + 16;
};

// PTO / PLO = parities of the crop offsets: they decide at compile time which taps feed which pixel of a
// thread's 2 x 2 output quad, so the col2im sum is 25 shared-memory reads at constant offsets.
template <int CB, int PTO, int PLO>
__global__ void __launch_bounds__(DG_THREADS, 1) thin_dgrad_tc_kernel(const __grid_constant__ CUtensorMap small_map,
                                                                       const ThinDgArgs a) {
  bn_pdl_trigger();
  using S = ThinDgSmem<CB>;
  extern __shared__ __align__(1024) unsigned char sm[];
  float* P = reinterpret_cast<float*>(sm + S::OFF_P);
  uint64_t* b_full = reinterpret_cast<uint64_t*>(sm + S::OFF_BAR);      // [DG_STAGES]
  uint64_t* b_empty = b_full + DG_STAGES;                               // [DG_STAGES]
  uint64_t* acc_full = b_empty + DG_STAGES;                             // [2]
  uint64_t* acc_empty = acc_full + 2;                                   // [2]
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(acc_empty + 2);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    for (int s = 0; s < DG_STAGES; ++s) {
      mbar_init(smem_u32(b_full + s), 1);
      mbar_init(smem_u32(b_empty + s), 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(smem_u32(acc_full + s), 1);
      mbar_init(smem_u32(acc_empty + s), DG_EPI_WARPS);     // one arrival per epilogue warp
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == DG_EPI_WARPS + 1) tmem_alloc<512>(smem_u32(tmem_ptr));
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  const uint32_t sbase = smem_u32(sm);
  bn_pdl_wait();
  // A tile: weight row (cb, tap) sits at row (tap / 7) * 32 + cb * 7 + tap % 7, so that the four TMEM lane
  // quarters (one per epilogue warp pair) each hold a quarter of the taps; all other rows are zero
  for (int i = tid; i < 128 * 8; i += DG_THREADS) {
    const int row = i >> 3, c = i & 7;
    const int w4 = row >> 5, l = row & 31;
    const int cb = l / 7, j = l - cb * 7;
    const int tap = 7 * w4 + j;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (cb < CB && tap < 25) v = __ldg(reinterpret_cast<const float4*>(a.wdt + ((long long)cb * 25 + tap) * 32) + c);
    *reinterpret_cast<float4*>(sm + (row >> 3) * 1024 + (row & 7) * 128 + ((c ^ (row & 7)) << 4)) = v;
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();

  const long long total = a.total_tiles;
  auto decode = [&](long long T, int& f, int& Y0, int& X0) {
    f = (int)(T / a.tiles_per_frame);
    const int tt = (int)(T - (long long)f * a.tiles_per_frame);
    const int ty = tt / a.tiles_x;
    Y0 = ty * OT_H;
    X0 = (tt - ty * a.tiles_x) * OT_W;
  };

  if (warp == DG_EPI_WARPS) {
    // ======================= TMA producer ========================================================
    if (lane == 0) {
      int bi = 0;
      for (long long T = blockIdx.x; T < total; T += gridDim.x, ++bi) {
        const int stage = bi % DG_STAGES;
        if (bi >= DG_STAGES) mbar_wait(smem_u32(b_empty + stage), ((bi / DG_STAGES) - 1) & 1);
        int f, Y0, X0;
        decode(T, f, Y0, X0);
        const int sy_lo = (Y0 + a.pt - 3) >> 1, sx_lo = (X0 + a.pl - 3) >> 1;      // ceil((v - 4) / 2)
        const uint32_t bar = smem_u32(b_full + stage);
        mbar_expect_tx2(bar, (uint32_t)(ST_PIX * 128));
        tma_tile_4d2(sbase + S::OFF_B + stage * DG_BSTAGE, &small_map, bar, 0, sx_lo, sy_lo, f);
      }
    }
    __syncwarp();
  } else if (warp == DG_EPI_WARPS + 1) {
    // ======================= MMA issuer ==========================================================
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc(128, 256);
      const uint64_t ad0 = desc_k_sw128(sbase), bd0 = desc_k_sw128(sbase + S::OFF_B);
      int bi = 0;
      for (long long T = blockIdx.x; T < total; T += gridDim.x, ++bi) {
        const int stage = bi % DG_STAGES, buf = bi & 1;
        if (bi >= 2) {
          mbar_wait(smem_u32(acc_empty + buf), ((bi >> 1) - 1) & 1);
          tc_fence_after();
        }
        mbar_wait(smem_u32(b_full + stage), (bi / DG_STAGES) & 1);
        tc_fence_after();
        const uint64_t bd = bd0 + (uint64_t)((stage * DG_BSTAGE) >> 4);
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_tf32(tmem_base + buf * 256, ad0 + (uint64_t)(k * 2), bd + (uint64_t)(k * 2), idesc, k ? 1u : 0u);
        umma_commit(smem_u32(b_empty + stage));
        umma_commit(smem_u32(acc_full + buf));
      }
    }
    __syncwarp();
  } else {
    // ======================= epilogue: TMEM -> P (shared) -> col2im + sigmoid + loss ==============
    // copy role: warp w drains TMEM lanes 32 (w & 3) .., column blocks 4 (w >> 2) .. + 3
    const int w4 = warp & 3, chalf = warp >> 2;
    const int cbl = lane / 7, jl = lane - cbl * 7;
    const int my_tap = 7 * w4 + jl;
    const bool my_valid = cbl < CB && my_tap < 25;
    float* prow = P + (cbl * 25 + my_tap) * P_STRIDE + chalf * 128;
    // col2im role: thread q < 192 owns the 2 x 2 output quad (qy, qx) of the 24 x 32 tile
    const int qy = tid >> 4, qx = tid & 15;
    const bool has_quad = tid < (OT_H / 2) * (OT_W / 2);
    // small pixel of tap (ky, kx) for output (2 qy + py, 2 qx + px):  qy + (py + PTO - ky) / 2 + cy  (same in x), with
    // cy = (pt - PTO) / 2 - floor((pt - 3) / 2): the distance of the quad grid from the staged tile's origin
    const int cy = ((a.pt - PTO) >> 1) - ((a.pt - 3) >> 1), cx = ((a.pl - PLO) >> 1) - ((a.pl - 3) >> 1);
    const float* pq = P + (qy + cy) * ST_W + (qx + cx);
    float bias[CB];
#pragma unroll
    for (int c = 0; c < CB; ++c) bias[c] = a.bias ? __ldg(a.bias + c) : 0.f;
    double my_sse = 0.0;
    int cur_chunk = -1;
    auto flush = [&]() {
      if (a.target == nullptr || cur_chunk < 0) return;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) my_sse += __shfl_xor_sync(0xffffffffu, my_sse, o);
      if (lane == 0 && my_sse != 0.0) atomicAdd(a.sse + cur_chunk, my_sse);
      my_sse = 0.0;
    };
    int bi = 0;
    for (long long T = blockIdx.x; T < total; T += gridDim.x, ++bi) {
      const int buf = bi & 1;
      int f, Y0, X0;
      decode(T, f, Y0, X0);
      const int chunk = (f + a.frame_offset) / a.chunk_size;
      if (chunk != cur_chunk) {
        flush();
        cur_chunk = chunk;
      }
      const int len = min(a.chunk_size, a.n_total - chunk * a.chunk_size);
      const float gsc = a.coef / (float)len;
      // this thread's targets / masks, fetched before waiting for the accumulator
      const int Yq = Y0 + 2 * qy, Xq = X0 + 2 * qx;
      float tg[2][2][CB], mk[2][2][CB];
#pragma unroll
      for (int py = 0; py < 2; ++py)
#pragma unroll
        for (int px = 0; px < 2; ++px) {
          const bool ok = has_quad && Yq + py < a.Hb && Xq + px < a.Wb;
#pragma unroll
          for (int c = 0; c < CB; ++c) {
            const long long inchw = (((long long)f * CB + c) * a.Hb + Yq + py) * a.Wb + Xq + px;
            tg[py][px][c] = (ok && a.target) ? __ldg(a.target + inchw) : 0.f;
            mk[py][px][c] = (ok && a.mask) ? __ldg(a.mask + inchw) : 1.f;
          }
        }
      mbar_wait(smem_u32(acc_full + buf), (bi >> 1) & 1);
      tc_fence_after();
#pragma unroll
      for (int jb = 0; jb < 4; ++jb) {
        uint32_t r[32];
        tmem_ld32(tmem_base + ((uint32_t)(w4 * 32) << 16) + buf * 256 + chalf * 128 + jb * 32, r);
        tmem_ld_wait();
        if (my_valid) {
#pragma unroll
          for (int q = 0; q < 8; ++q)
            *reinterpret_cast<uint4*>(prow + jb * 32 + 4 * q) = make_uint4(r[4 * q], r[4 * q + 1], r[4 * q + 2], r[4 * q + 3]);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(acc_empty + buf));
      asm volatile("bar.sync 1, 256;" ::: "memory");
      if (has_quad) {
        float acc[2][2][CB];
#pragma unroll
        for (int py = 0; py < 2; ++py)
#pragma unroll
          for (int px = 0; px < 2; ++px)
#pragma unroll
            for (int c = 0; c < CB; ++c) acc[py][px][c] = bias[c];
#pragma unroll
        for (int ky = 0; ky < 5; ++ky) {
          constexpr int dummy = 0;
          (void)dummy;
          const int py = (ky + PTO) & 1;                       // the quad row this tap row feeds
          const int dy = (py + PTO - ky) / 2;                  // exact: same parity (C++ division of an even number)
#pragma unroll
          for (int kx = 0; kx < 5; ++kx) {
            const int px = (kx + PLO) & 1;
            const int dx = (px + PLO - kx) / 2;
#pragma unroll
            for (int c = 0; c < CB; ++c)
              acc[py][px][c] += pq[((c * 25 + ky * 5 + kx) * P_STRIDE) + dy * ST_W + dx];
          }
        }
#pragma unroll
        for (int py = 0; py < 2; ++py)
#pragma unroll
          for (int px = 0; px < 2; ++px) {
            const int Y = Yq + py, X = Xq + px;
            if (Y < a.Hb && X < a.Wb) {
#pragma unroll
              for (int c = 0; c < CB; ++c) {
                const float v = 1.f / (1.f + expf(-acc[py][px][c]));
                const long long inchw = (((long long)f * CB + c) * a.Hb + Y) * a.Wb + X;
                a.xhat_ws[inchw] = v;
                if (a.target) {
                  const float d = v - tg[py][px][c];
                  my_sse += (double)(d * d * mk[py][px][c]);
                  a.dpre[(((long long)f * a.Hb + Y) * a.Wb + X) * CB + c] = gsc * d * mk[py][px][c] * v * (1.f - v);
                }
              }
            }
          }
      }
      asm volatile("bar.sync 1, 256;" ::: "memory");        // P is free for the next tile
    }
    flush();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == DG_EPI_WARPS + 1) {
    tc_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}

std::vector<std::pair<MapKey, CUtensorMap>> g_dg_maps;

// NHWC fp32 image (32 channels) as (C, W, H, N); box = 32 channels x 18 pixels x 14 rows, SWIZZLE_128B (K-major rows)
bool small_map(const float* p, int n, int H, int W, int C, CUtensorMap* out) {
  MapKey key{p, n, H, W, C};
  std::lock_guard<std::mutex> lk(g_mu);
  for (auto& kv : g_dg_maps)
    if (kv.first == key) { *out = kv.second; return true; }
  cuuint64_t gdim[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)n};
  cuuint64_t gstr[3] = {(cuuint64_t)C * 4, (cuuint64_t)W * C * 4, (cuuint64_t)H * W * C * 4};
  cuuint32_t box[4] = {32, (cuuint32_t)ST_W, (cuuint32_t)ST_H, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUtensorMap m;
  if (g_enc(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, (void*)p, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
    return false;
  if (g_dg_maps.size() >= 64) g_dg_maps.clear();
  g_dg_maps.emplace_back(key, m);
  *out = m;
  return true;
}

template <int CB, int PTO, int PLO>
int launch_dg2(const CUtensorMap& map, const ThinDgArgs& a, cudaStream_t st) {
  using S = ThinDgSmem<CB>;
  static_assert(S::TOTAL <= 227 * 1024, "shared memory per CTA");
  auto kern = thin_dgrad_tc_kernel<CB, PTO, PLO>;
  static bool configured = false;
  if (!configured) {
    BN_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, S::TOTAL));
    configured = true;
  }
  const long long blocks = a.total_tiles < 148 ? a.total_tiles : 148;
  BN_CUDA(bn_launch(kern, dim3((unsigned)blocks), DG_THREADS, S::TOTAL, st, map, a));
  BN_LAUNCHED();
  return 0;
}

template <int CB>
int launch_dg(const CUtensorMap& map, const ThinDgArgs& a, cudaStream_t st) {
  switch ((a.pt & 1) * 2 + (a.pl & 1)) {
    case 0: return launch_dg2<CB, 0, 0>(map, a, st);
    case 1: return launch_dg2<CB, 0, 1>(map, a, st);
    case 2: return launch_dg2<CB, 1, 0>(map, a, st);
    default: return launch_dg2<CB, 1, 1>(map, a, st);
  }
}

}  // namespace

// returns 1 when the geometry is not covered (caller falls back to the FP32 kernel)
int bn_launch_thin_wgrad_tc(const ImgView& big, const float* small, const ConvGeom& g, int n, float* partial,
                            size_t partial_floats, float* grad, cudaStream_t st) {
  if (grad == nullptr || n <= 0) return grad == nullptr ? 0 : 1;
  if (!(g.k == 5 && g.s == 2 && g.Cb >= 1 && g.Cb <= 4 && g.Cs % 32 == 0)) return 1;
  if (((uintptr_t)small & 127) || ((uintptr_t)partial & 15) || !have_tma()) return 1;
  static const bool off = [] { const char* e = getenv("BN_THIN_WGRAD_TC"); return e && e[0] == '0'; }();
  if (off) return 1;
  CUtensorMap map;
  if (!fat_map(small, n, g.Hs, g.Ws, g.Cs, &map)) return 1;
  ThinWgArgs a;
  a.thin = big; a.Hs = g.Hs; a.Ws = g.Ws; a.Cs = g.Cs; a.pt = g.pt; a.pl = g.pl; a.n = n;
  a.tiles_x = bn_cdiv(g.Ws, WT_W);
  a.tiles_per_frame = a.tiles_x * bn_cdiv(g.Hs, WT_H);
  a.total_tiles = (long long)a.tiles_per_frame * n;
  a.partial = partial;
  const int Ktot = 25 * g.Cb, cgroups = g.Cs / 32;
  int per_sm;
  switch (g.Cb) {
    case 1: per_sm = resident_ctas<1>(); break;
    case 2: per_sm = resident_ctas<2>(); break;
    case 3: per_sm = resident_ctas<3>(); break;
    default: per_sm = resident_ctas<4>(); break;
  }
  long long blocks = 148LL * per_sm / cgroups;
  if (blocks < 1) blocks = 1;
  if (blocks > a.total_tiles) blocks = a.total_tiles;
  const long long cap = (long long)(partial_floats / ((size_t)Ktot * g.Cs));
  if (blocks > cap) blocks = cap;
  if (blocks < 1) return 1;
  int r;
  switch (g.Cb) {
    case 1: r = launch<1>(map, a, (int)blocks, cgroups, st); break;
    case 2: r = launch<2>(map, a, (int)blocks, cgroups, st); break;
    case 3: r = launch<3>(map, a, (int)blocks, cgroups, st); break;
    default: r = launch<4>(map, a, (int)blocks, cgroups, st); break;
  }
  if (r) return r;
  return bn_launch_wgrad_reduce(partial, (int)blocks, Ktot, g.Cs, g.Cb, 25, g.d_fprop, grad, st);
}

// Last decoder layer forward with the fused reconstruction loss, tensor-core form (training pass only: needs a
// target, writes x_hat to the workspace copy).  Returns 1 when the geometry is not covered.
int bn_launch_thin_dgrad_tc(const float* small, const ConvGeom& g, const float* wdt, const float* bias, int n,
                            float* xhat_ws, const float* target, const float* mask, int chunk_size, int frame_offset,
                            int n_total, float grad_coef, double* sse, float* dpre, cudaStream_t st) {
  if (n <= 0 || target == nullptr) return 1;
  if (!(g.k == 5 && g.s == 2 && g.Cb >= 1 && g.Cb <= 4 && g.Cs == 32)) return 1;
  if (g.pt < 0 || g.pl < 0 || g.pt > 4 || g.pl > 4) return 1;
  if (((uintptr_t)small & 127) || ((uintptr_t)wdt & 15) || !have_tma()) return 1;
  static const bool off = [] { const char* e = getenv("BN_THIN_DGRAD_TC"); return e && e[0] == '0'; }();
  if (off) return 1;
  CUtensorMap map;
  if (!small_map(small, n, g.Hs, g.Ws, g.Cs, &map)) return 1;
  ThinDgArgs a;
  a.Hs = g.Hs; a.Ws = g.Ws; a.Hb = g.Hb; a.Wb = g.Wb; a.pt = g.pt; a.pl = g.pl; a.n = n;
  a.wdt = wdt; a.bias = bias; a.xhat_ws = xhat_ws; a.target = target; a.mask = mask;
  a.n_total = n_total > 0 ? n_total : n; a.frame_offset = frame_offset;
  a.chunk_size = chunk_size > 0 ? chunk_size : a.n_total; a.coef = grad_coef; a.sse = sse; a.dpre = dpre;
  a.tiles_x = bn_cdiv(g.Wb, OT_W);
  a.tiles_per_frame = a.tiles_x * bn_cdiv(g.Hb, OT_H);
  a.total_tiles = (long long)a.tiles_per_frame * n;
  switch (g.Cb) {
    case 1: return launch_dg<1>(map, a, st);
    case 2: return launch_dg<2>(map, a, st);
    case 3: return launch_dg<3>(map, a, st);
    default: return launch_dg<4>(map, a, st);
  }
}

// thin image -> fat image, decoupled-role tensor-core form (see thin_fprop_tc2_kernel).  Returns 1 when not
// applicable (the caller then uses thin_fprop_tc_kernel / the fp32 kernel of cae_thin.cu).
int bn_launch_thin_fprop_tc2(const ImgView& big, const unsigned char* big_u8, const ConvGeom& g, const float* wft,
                             const float* bias, float* out, const float* dact, int act, int n, float* colsum,
                             cudaStream_t st) {
  // Measured on B200 (C2 step, same box): 2.346 ms with the single-role kernel of cae_thin.cu (two CTAs per SM
  // overlap each other) vs 2.367 ms with this one-CTA-per-SM role split -- parity-tested, kept behind
  // BN_THIN_FPROP_V2=1 for the record, off by default.
  static const bool on = [] { const char* e = getenv("BN_THIN_FPROP_V2"); return e && e[0] == '1'; }();
  if (!on || n <= 0) return 1;
  if (!(g.k == 5 && g.s == 2 && g.Cb >= 1 && g.Cb <= 4 && g.Cs % 32 == 0)) return 1;
  if (((uintptr_t)out & 15) || (dact && ((uintptr_t)dact & 15)) || (bias && ((uintptr_t)bias & 15))) return 1;
  ThinFwArgs a;
  a.thin = big; a.thin8 = big_u8; a.Hs = g.Hs; a.Ws = g.Ws; a.Cs = g.Cs; a.pt = g.pt; a.pl = g.pl; a.n = n;
  a.wft = wft; a.bias = bias; a.out = out; a.dact = dact; a.act = act; a.colsum = colsum;
  a.tiles_x = bn_cdiv(g.Ws, FT_W);
  a.tiles_per_frame = a.tiles_x * bn_cdiv(g.Hs, FT_H);
  a.total_tiles = (long long)a.tiles_per_frame * n;
  switch (g.Cb) {
    case 1: return launch_fw<1>(a, g.Cs / 32, st);
    case 2: return launch_fw<2>(a, g.Cs / 32, st);
    case 3: return launch_fw<3>(a, g.Cs / 32, st);
    default: return launch_fw<4>(a, g.Cs / 32, st);
  }
}
