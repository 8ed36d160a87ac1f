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
constexpr int WTHREADS = 256;

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
  static constexpr int OFF_PATCH = OFF_B + 2 * B_BUF;
  static constexpr int OFF_BAR = OFF_PATCH + ((2 * PATCH + 15) & ~15);
  // the M = 128 MMA reads 16 KB per chunk whatever ROWS is: rows >= ROWS only feed accumulator rows that
  // are never read, but their addresses must stay inside the CTA's shared-memory window
  static constexpr int TOTAL_MIN = OFF_BAR + 64;
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

template <int CB>
__global__ void __launch_bounds__(WTHREADS, 2) thin_wgrad_tc_kernel(const __grid_constant__ CUtensorMap fat_map,
                                                                 const ThinWgArgs a) {
  bn_pdl_trigger();
  using S = ThinWgSmem<CB>;
  constexpr int ROWS = S::ROWS;
  extern __shared__ __align__(1024) unsigned char sm[];
  typedef float (*Patch)[WP_ROWS][WP_COLS];
  Patch patch[2] = {reinterpret_cast<Patch>(sm + S::OFF_PATCH), reinterpret_cast<Patch>(sm + S::OFF_PATCH + S::PATCH)};
  uint64_t* b_full = reinterpret_cast<uint64_t*>(sm + S::OFF_BAR);    // [2] TMA landed
  uint64_t* ab_free = b_full + 2;                                     // [2] MMAs that read A / B buffer s have retired
  uint64_t* done_bar = ab_free + 2;                                   // [1]
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(done_bar + 1);
  const int tid = threadIdx.x, warp = tid >> 5;
  if (tid == 0) {
    for (int s = 0; s < 2; ++s) {
      mbar_init(smem_u32(b_full + s), 1);
      mbar_init(smem_u32(ab_free + s), 1);
    }
    mbar_init(smem_u32(done_bar), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) tmem_alloc<32>(smem_u32(tmem_ptr));
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
  auto issue_patch = [&](Patch dst, int f, int y0, int x0) {
    for (int i = tid; i < CB * WP_ROWS * WP_COLS; i += WTHREADS) {
      const int c = i / (WP_ROWS * WP_COLS);
      const int rem = i - c * WP_ROWS * WP_COLS;
      const int r = rem / WP_COLS, col = rem - r * WP_COLS;
      const int y = 2 * y0 - a.pt + r, x = 2 * x0 - a.pl + col;
      const bool ok = (unsigned)y < (unsigned)a.thin.H && (unsigned)x < (unsigned)a.thin.W;
      const long long off = ok ? (long long)f * a.thin.sn + (long long)y * a.thin.sy + (long long)x * a.thin.sx +
                                     (long long)c * a.thin.sc
                               : 0;
      cp_async4z(&dst[c][r][col], a.thin.p + off, ok);
    }
  };
  auto issue_fat = [&](int buf, int f, int y0, int x0) {       // one thread
    const uint32_t bar = smem_u32(b_full + buf);
    mbar_expect_tx2(bar, (uint32_t)S::B_BUF);
    tma_tile_4d2(sbase + S::OFF_B + buf * S::B_BUF, &fat_map, bar, c0, x0, y0, f);
  };

  long long t = blockIdx.x;
  int f, y0, x0;
  if (t < a.total_tiles) {
    decode(t, f, y0, x0);
    issue_patch(patch[0], f, y0, x0);
    if (tid == 0) issue_fat(0, f, y0, x0);
  }
  asm volatile("cp.async.commit_group;" ::: "memory");
  const int prow = (tid & 127) >> 5, pcol = tid & 31;          // this thread's pixel of the 4 x 32 tile
  const int half = tid >> 7;                                   // taps 0..12 / 13..24
  int it = 0;
  for (; t < a.total_tiles; t += gridDim.x, ++it) {
    const int buf = it & 1;
    const long long tn = t + gridDim.x;
    if (tn < a.total_tiles) {
      int fn, yn, xn;
      decode(tn, fn, yn, xn);
      issue_patch(patch[buf ^ 1], fn, yn, xn);                 // its last readers finished before the previous barrier
      if (tid == 0) {
        // fat buffer buf^1 and A buffer buf^1 were last read by the MMAs of tile it-1
        if (it >= 1) mbar_wait(smem_u32(ab_free + (buf ^ 1)), ((it - 1) >> 1) & 1);
        issue_fat(buf ^ 1, fn, yn, xn);
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
    asm volatile("cp.async.wait_group 1;" ::: "memory");
    __syncthreads();                                           // patch[buf] is complete and visible
    // A buffer buf was last read by the MMAs of tile it-2
    if (it >= 2) mbar_wait(smem_u32(ab_free + buf), ((it - 2) >> 1) & 1);
    {
      unsigned char* abase = sm + buf * S::A_BUF + prow * S::A_CHUNK;
      const int tap0 = half ? 13 : 0, tap1 = half ? 25 : 13;
#pragma unroll 1
      for (int tap = tap0; tap < tap1; ++tap) {
        const int ky = tap / 5, kx = tap - ky * 5;
#pragma unroll
        for (int cb = 0; cb < CB; ++cb) {
          const int r = tap * CB + cb;
          const float v = patch[buf][cb][2 * prow + ky][2 * pcol + kx];
          *reinterpret_cast<float*>(abase + (r >> 3) * 1024 + (r & 7) * 128 + ((((pcol >> 2) ^ (r & 7))) << 4) +
                                    (pcol & 3) * 4) = v;
        }
      }
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    if (tid == 0) {
      tc_fence_after();
      mbar_wait(smem_u32(b_full + buf), (it >> 1) & 1);
      tc_fence_after();
      constexpr uint32_t idesc = make_idesc(128, 32) | (1u << 16);          // A K-major, B MN-major
      const uint32_t sa = sbase + buf * S::A_BUF, sb = sbase + S::OFF_B + buf * S::B_BUF;
#pragma unroll
      for (int kc = 0; kc < WT_H; ++kc)
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_tf32(tmem_base, desc_k_sw128(sa + kc * S::A_CHUNK + k * 32),
                    desc_mn_sw128(sb + kc * 4096 + k * 1024, 4096, 512), idesc, (it | kc | k) != 0 ? 1u : 0u);
      umma_commit(smem_u32(ab_free + buf));
    }
  }
  // ---- epilogue: this CTA's slice [(tap, cb)][cs] of the partial buffer
  if (tid == 0) umma_commit(smem_u32(done_bar));
  __syncthreads();
  if (it > 0) {
    mbar_wait(smem_u32(done_bar), 0);
    tc_fence_after();
  }
  if (warp < 4) {
    const int lane = tid & 31;
    uint32_t r[32];
    if (it > 0) {
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
  if (warp == 0) {
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
  BN_CUDA(bn_launch(kern, dim3(blocks, cgroups), WTHREADS, S::TOTAL, st, map, a));
  BN_LAUNCHED();
  return 0;
}

template <int CB>
int resident_ctas() {
  const int per = (227 * 1024) / (ThinWgSmem<CB>::TOTAL + 1024);
  return per > 3 ? 3 : (per < 1 ? 1 : per);
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
