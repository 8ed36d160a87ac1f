// Shared-memory-tiled kernels for the thin first / last layers of the default architecture
// (kernel 5, stride 2, 1-4 image channels <-> 32k feature channels).  These layers carry < 2 % of
// the FLOPs but move the largest activations (64x64x32 per frame); they are HBM-bound, so the
// im2col duplication of the implicit-GEMM kernels (6.25x re-reads through L2) is what has to go:
// every input element is staged ONCE per tile in shared memory and the 25 taps read it from there.
//
//   thin_fprop_kernel  : thin image -> fat image   (encoder conv0 forward; last convT backward-data)
//   thin_wgrad_kernel  : weight gradient of the same geometry (conv0.weight, convtranspose{last}.weight)
//   thin_dgrad5_kernel : fat image -> thin image + sigmoid + fused reconstruction loss
//                        (last ConvTranspose2d forward, aes.py:463-470 + losses.py:36-96)
// Lane = feature channel in the first two (coalesced 128-byte rows per pixel, taps read by
// broadcast LDS.128); lane = output pixel of one stride-residue class in the third.
#include <stdlib.h>

#include "cae_kernels.cuh"
#include "tc_common.cuh"

namespace {

constexpr int TH = 8, TW = 32;             // small-image tile per block (fprop / wgrad)
constexpr int PROWS = 2 * (TH - 1) + 5;    // 19 patch rows
constexpr int PCOLS = 72;                  // 2*(TW-1)+5 = 67 used, padded for 16-byte LDS

struct ThinGeo {
  ImgView big;          // thin image (C <= 4), any strides
  int Hs, Ws, Cs;       // fat image
  int pt, pl, n;
  // raw uint8 video (0..255) with the strides of `big`; the patch loader divides by 255 on the way
  // to shared memory (the reference does that on the host, data_generator.py:258-263)
  const unsigned char* big8 = nullptr;
};

__device__ __forceinline__ void cp_async4(float* dst, const float* src, bool ok) {
  const uint32_t d = (uint32_t)__cvta_generic_to_shared(dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(d), "l"(src), "r"(ok ? 4u : 0u) : "memory");
}
__device__ __forceinline__ void cp_async16(float* dst, const float* src, bool ok) {
  const uint32_t d = (uint32_t)__cvta_generic_to_shared(dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(src), "r"(ok ? 16u : 0u) : "memory");
}
__device__ __forceinline__ void cp_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// asynchronous (cp.async, zero-filled outside the image) copy of one tile's input patch
template <int CB>
__device__ __forceinline__ void issue_patch(float (*patch)[PROWS][PCOLS], const ThinGeo& g, int f, int y0,
                                            int x0, int tid) {
  for (int i = tid; i < CB * PROWS * PCOLS; i += 256) {
    int c = i / (PROWS * PCOLS);
    int rem = i - c * PROWS * PCOLS;
    int r = rem / PCOLS, col = rem - r * PCOLS;
    int y = 2 * y0 - g.pt + r, x = 2 * x0 - g.pl + col;
    const bool ok = (unsigned)y < (unsigned)g.big.H && (unsigned)x < (unsigned)g.big.W;
    const long long off = ok ? (long long)f * g.big.sn + (long long)y * g.big.sy + (long long)x * g.big.sx +
                                   (long long)c * g.big.sc
                             : 0;
    if (g.big8 != nullptr) {
      // same rounding as numpy's float32(u8) / 255
      patch[c][r][col] = ok ? __fdiv_rn((float)__ldg(g.big8 + off), 255.f) : 0.f;
    } else {
      cp_async4(&patch[c][r][col], g.big.p + off, ok);
    }
  }
}

struct TileIter {
  int tiles_x, tiles_per_frame;
  long long total;
  __device__ __forceinline__ void decode(long long t, int& f, int& y0, int& x0) const {
    f = (int)(t / tiles_per_frame);
    const int tt = (int)(t - (long long)f * tiles_per_frame);
    const int ty = tt / tiles_x;
    y0 = ty * TH;
    x0 = (tt - ty * tiles_x) * TW;
  }
};

// Persistent blocks: the 25*CB weights of this lane's channel stay in registers for the whole kernel,
// the next tile's patch streams in with cp.async while the current one is computed, and the
// activation-derivative mask of a half row is fetched before the FMAs that precede its use.
template <int CB>
__global__ void __launch_bounds__(256) thin_fprop_kernel(const ThinGeo g, const float* __restrict__ w,
                                                         const float* __restrict__ bias,
                                                         float* __restrict__ out,
                                                         const float* __restrict__ dact, int act,
                                                         const TileIter it) {
  bn_pdl_trigger();
  bn_pdl_wait();
  extern __shared__ __align__(16) float thin_sm[];
  typedef float (*Patch)[PROWS][PCOLS];
  Patch patch[2] = {reinterpret_cast<Patch>(thin_sm), reinterpret_cast<Patch>(thin_sm + CB * PROWS * PCOLS)};
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int c = blockIdx.y * 32 + lane;
  float wr[25 * CB];
#pragma unroll
  for (int i = 0; i < 25 * CB; ++i) wr[i] = __ldg(w + (long long)i * g.Cs + c);    // [(tap, cb)][cs]
  const float b = bias ? __ldg(bias + c) : 0.f;
  long long t = blockIdx.x;
  int f, y0, x0;
  // activation-derivative mask of half a row (16 pixels of this lane's channel), fetched half a tile
  // ahead of its use: dq[0] / dq[1] alternate between "being loaded" and "being consumed"
  float dq[2][TW / 2];
  auto load_mask = [&](float (&d)[TW / 2], int ff, int yy0, int xx0, int half) {
    const int oy = yy0 + warp;
    const long long rb = (((long long)ff * g.Hs + oy) * g.Ws + xx0) * g.Cs + c;
#pragma unroll
    for (int e = 0; e < TW / 2; ++e)
      d[e] = (oy < g.Hs && xx0 + half * (TW / 2) + e < g.Ws) ? __ldg(dact + rb + (long long)(half * (TW / 2) + e) * g.Cs)
                                                            : 1.f;
  };
  if (t < it.total) {
    it.decode(t, f, y0, x0);
    issue_patch<CB>(patch[0], g, f, y0, x0, tid);
    if (dact) load_mask(dq[0], f, y0, x0, 0);
  }
  cp_commit();
  for (int buf = 0; t < it.total; t += gridDim.x, buf ^= 1) {
    it.decode(t, f, y0, x0);
    const long long tn = t + gridDim.x;
    int fn = 0, yn = 0, xn = 0;
    if (tn < it.total) {
      it.decode(tn, fn, yn, xn);
      issue_patch<CB>(patch[buf ^ 1], g, fn, yn, xn, tid);
    }
    cp_commit();
    cp_wait<1>();
    __syncthreads();
    const int oy = y0 + warp;
    const long long rowbase = (((long long)f * g.Hs + oy) * g.Ws + x0) * g.Cs + c;
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      if (dact) {
        if (half == 0) load_mask(dq[1], f, y0, x0, 1);
        else if (tn < it.total) load_mask(dq[0], fn, yn, xn, 0);
      }
      if (oy < g.Hs) {
#pragma unroll
        for (int jj = 0; jj < TW / 4; ++jj) {
          const int j = half * (TW / 4) + jj;
          float a0 = 0.f, a1 = 0.f;
#pragma unroll
          for (int ky = 0; ky < 5; ++ky) {
#pragma unroll
            for (int cb = 0; cb < CB; ++cb) {
              const float4 v0 = *reinterpret_cast<const float4*>(&patch[buf][cb][2 * warp + ky][4 * j]);
              const float4 v1 = *reinterpret_cast<const float4*>(&patch[buf][cb][2 * warp + ky][4 * j + 4]);
              const float v[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
#pragma unroll
              for (int kx = 0; kx < 5; ++kx) {
                const float wv = wr[(ky * 5 + kx) * CB + cb];
                a0 = fmaf(v[kx], wv, a0);
                a1 = fmaf(v[kx + 2], wv, a1);
              }
            }
          }
          float r[2] = {a0 + b, a1 + b};
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            const int xo = 2 * j + e;
            if (x0 + xo >= g.Ws) continue;
            float x = r[e];
            if (act == BN_ACT_LEAKY) x = x > 0.f ? x : BN_LEAK * x;
            if (dact) x *= dq[half][2 * jj + e] > 0.f ? 1.f : BN_LEAK;
            out[rowbase + (long long)xo * g.Cs] = x;
          }
        }
      }
    }
    __syncthreads();       // patch[buf] is refilled by the next iteration's prefetch
  }
}

template <int CB>
__global__ void __launch_bounds__(256) thin_wgrad_kernel(const ThinGeo g, const float* __restrict__ small,
                                                         float* __restrict__ partial, const TileIter it) {
  bn_pdl_trigger();
  bn_pdl_wait();
  // persistent: each block walks tiles blockIdx.x, +gridDim.x, ...; lane = small-image channel
  extern __shared__ __align__(16) float thin_sm[];
  typedef float (*Patch)[PROWS][PCOLS];
  Patch patch[2] = {reinterpret_cast<Patch>(thin_sm), reinterpret_cast<Patch>(thin_sm + CB * PROWS * PCOLS)};
  float* red = thin_sm + 2 * CB * PROWS * PCOLS;          // [8][25*CB][32]
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int c = blockIdx.y * 32 + lane;
  float acc[25 * CB];
#pragma unroll
  for (int i = 0; i < 25 * CB; ++i) acc[i] = 0.f;
  long long t = blockIdx.x;
  int f, y0, x0;
  if (t < it.total) {
    it.decode(t, f, y0, x0);
    issue_patch<CB>(patch[0], g, f, y0, x0, tid);
  }
  cp_commit();
  for (int buf = 0; t < it.total; t += gridDim.x, buf ^= 1) {
    it.decode(t, f, y0, x0);
    const long long tn = t + gridDim.x;
    if (tn < it.total) {
      int fn, yn, xn;
      it.decode(tn, fn, yn, xn);
      issue_patch<CB>(patch[buf ^ 1], g, fn, yn, xn, tid);
    }
    cp_commit();
    const int oy = y0 + warp;
    // this warp's row of the small image: fetched before waiting for the patch
    float s[TW];
    if (oy < g.Hs) {
      const float* srow = small + (((long long)f * g.Hs + oy) * g.Ws + x0) * g.Cs + c;
#pragma unroll
      for (int e = 0; e < TW; ++e) s[e] = x0 + e < g.Ws ? __ldg(srow + (long long)e * g.Cs) : 0.f;
    }
    cp_wait<1>();
    __syncthreads();
    if (oy < g.Hs) {
#pragma unroll
      for (int j = 0; j < TW / 2; ++j) {
#pragma unroll
        for (int ky = 0; ky < 5; ++ky) {
#pragma unroll
          for (int cb = 0; cb < CB; ++cb) {
            const float4 v0 = *reinterpret_cast<const float4*>(&patch[buf][cb][2 * warp + ky][4 * j]);
            const float4 v1 = *reinterpret_cast<const float4*>(&patch[buf][cb][2 * warp + ky][4 * j + 4]);
            const float v[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
#pragma unroll
            for (int kx = 0; kx < 5; ++kx) {
              float& a = acc[(ky * 5 + kx) * CB + cb];
              a = fmaf(v[kx], s[2 * j], a);
              a = fmaf(v[kx + 2], s[2 * j + 1], a);
            }
          }
        }
      }
    }
    __syncthreads();
  }
  // cross-warp reduction, then this block's slice of the partial buffer [(tap, cb)][Cs]
#pragma unroll
  for (int i = 0; i < 25 * CB; ++i) red[(warp * 25 * CB + i) * 32 + lane] = acc[i];
  __syncthreads();
  for (int i = warp; i < 25 * CB; i += 8) {
    float sum = 0.f;
#pragma unroll
    for (int q = 0; q < 8; ++q) sum += red[(q * 25 * CB + i) * 32 + lane];
    partial[((long long)blockIdx.x * 25 * CB + i) * g.Cs + c] = sum;
  }
}

// ------------------------------------------------------------------------------------------------
// last decoder layer, kernel 5 / stride 2 fast path
// ------------------------------------------------------------------------------------------------
struct Dg5Args {
  const float* small;
  int Hs, Ws, Cs, Hb, Wb, pt, pl, n;
  const float* wd;        // [(tap, cs)][cb]
  const float* bias;
  float* xhat_ws;
  float* xhat_user;
  const float* target;
  const float* mask;
  int chunk_size, frame_offset, n_total;
  float coef;
  double* sse;
  float* dpre;
};

constexpr int DT = 16;      // small-image tile edge: a thread owns one small pixel = a 2x2 output quad
constexpr int DP = DT + 2;  // input patch edge (3x3 neighbourhood)

// Thread = one small-image pixel and its 2x2 quad of output pixels (the four stride-residue classes).
// The quad's 25 taps read only the 3x3 small-pixel neighbourhood, so every neighbour vector is
// fetched from shared memory ONCE per 4 channels and reused by all taps / classes that touch it
// (9 lane-strided + 25 broadcast LDS.128 per 100*CB FMAs: FMA-bound instead of LDS-bound).
// PTO / PLO = parity of the crop offsets: they fix which taps belong to which class at compile time.
template <int CB, int PTO, int PLO>
__global__ void __launch_bounds__(256) thin_dgrad5_kernel(const Dg5Args a, int tiles_x) {
  bn_pdl_trigger();
  bn_pdl_wait();
  extern __shared__ __align__(16) float sm[];
  const int Cs = a.Cs, PS = Cs + 4;                    // padded pixel stride: conflict-free LDS.128
  float* patch = sm;                                   // [DP*DP][PS]
  float* wsm = sm + DP * DP * PS;                      // [25][CB][Cs]
  __shared__ double sse_sm;
  const int tid = threadIdx.x, lane = tid & 31;
  const int f = blockIdx.y;
  const int ty = blockIdx.x / tiles_x, tx = blockIdx.x - ty * tiles_x;
  const int sy0 = ty * DT, sx0 = tx * DT;              // small-pixel origin of the tile
  // output pixel (2*ys + py, 2*xs + px) gathers small pixel (ys + (py + pt - ky) / 2, ..) for the
  // taps with ky = (py + pt) mod 2 (+2, +4): offsets span [(pt - PTO) / 2 - 1 - .., ..]; with
  // dy(py, ky) = (py + PTO - ky) / 2 in {-2..1} relative to the shifted origin ys + (pt - PTO) / 2
  const int shy = (a.pt - PTO) >> 1, shx = (a.pl - PLO) >> 1;
  // neighbour offsets: odd crop -> dy in {-1, 0, 1}, even crop -> dy in {-2, -1, 0}; patch row = dy + OY
  constexpr int OY = PTO ? 1 : 2, OX = PLO ? 1 : 2;
  const int iy0 = sy0 + shy - OY, ix0 = sx0 + shx - OX;
  for (int i = tid; i < DP * DP * (Cs / 4); i += 256) {
    const int p = i / (Cs / 4), q = i - p * (Cs / 4);
    const int pr = p / DP, pc = p - pr * DP;
    const int iy = iy0 + pr, ix = ix0 + pc;
    const bool ok = (unsigned)iy < (unsigned)a.Hs && (unsigned)ix < (unsigned)a.Ws;
    cp_async16(patch + p * PS + q * 4, ok ? a.small + (((long long)f * a.Hs + iy) * a.Ws + ix) * Cs + q * 4 : a.small, ok);
  }
  cp_commit();
  for (int i = tid; i < 25 * CB * Cs; i += 256) {
    const int tap = i / (CB * Cs);
    const int rem = i - tap * CB * Cs;
    const int cb = rem / Cs, ci = rem - cb * Cs;
    wsm[i] = __ldg(a.wd + ((long long)tap * Cs + ci) * CB + cb);
  }
  if (tid == 0) sse_sm = 0.0;
  cp_wait<0>();
  __syncthreads();
  const int yy = tid >> 4, xx = tid & 15;
  float acc[2][2][CB];
#pragma unroll
  for (int py = 0; py < 2; ++py)
#pragma unroll
    for (int px = 0; px < 2; ++px)
#pragma unroll
      for (int c = 0; c < CB; ++c) acc[py][px][c] = 0.f;
  for (int q = 0; q < Cs; q += 4) {
#pragma unroll
    for (int ny = 0; ny < 3; ++ny) {
#pragma unroll
      for (int nx = 0; nx < 3; ++nx) {
        const float4 v = *reinterpret_cast<const float4*>(patch + ((yy + ny) * DP + xx + nx) * PS + q);
#pragma unroll
        for (int py = 0; py < 2; ++py) {
#pragma unroll
          for (int ky = (py + PTO) & 1; ky < 5; ky += 2) {
            if ((py + PTO - ky) / 2 + OY != ny) continue;         // compile-time after unrolling
#pragma unroll
            for (int px = 0; px < 2; ++px) {
#pragma unroll
              for (int kx = (px + PLO) & 1; kx < 5; kx += 2) {
                if ((px + PLO - kx) / 2 + OX != nx) continue;
                const float* wp = wsm + (ky * 5 + kx) * CB * Cs + q;
#pragma unroll
                for (int c = 0; c < CB; ++c) {
                  const float4 wv = *reinterpret_cast<const float4*>(wp + c * Cs);
                  float s = acc[py][px][c];
                  s = fmaf(v.x, wv.x, s);
                  s = fmaf(v.y, wv.y, s);
                  s = fmaf(v.z, wv.z, s);
                  s = fmaf(v.w, wv.w, s);
                  acc[py][px][c] = s;
                }
              }
            }
          }
        }
      }
    }
  }
  double my_sse = 0.0;
  const int chunk = (f + a.frame_offset) / a.chunk_size;
  const int len = min(a.chunk_size, a.n_total - chunk * a.chunk_size);
  const float gsc = a.coef / (float)len;
#pragma unroll
  for (int py = 0; py < 2; ++py) {
    const int y = 2 * (sy0 + yy) + py;
#pragma unroll
    for (int px = 0; px < 2; ++px) {
      const int x = 2 * (sx0 + xx) + px;
      if (y >= a.Hb || x >= a.Wb) continue;
#pragma unroll
      for (int c = 0; c < CB; ++c) {
        float v = acc[py][px][c] + (a.bias ? __ldg(a.bias + c) : 0.f);
        v = 1.f / (1.f + expf(-v));
        const long long inchw = (((long long)f * CB + c) * a.Hb + y) * a.Wb + x;
        a.xhat_ws[inchw] = v;
        if (a.xhat_user) a.xhat_user[inchw] = v;
        if (a.target) {
          const float d = v - __ldg(a.target + inchw);
          const float m = a.mask ? __ldg(a.mask + inchw) : 1.f;
          my_sse += (double)(d * d * m);
          a.dpre[(((long long)f * a.Hb + y) * a.Wb + x) * CB + c] = gsc * d * m * v * (1.f - v);
        }
      }
    }
  }
  if (a.target) {
    // a block lies inside one frame, hence inside one chunk
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) my_sse += __shfl_xor_sync(0xffffffffu, my_sse, o);
    if (lane == 0 && my_sse != 0.0) atomicAdd(&sse_sm, my_sse);
    __syncthreads();
    if (tid == 0 && sse_sm != 0.0) atomicAdd(a.sse + chunk, sse_sm);
  }
}

bool fast_geom(const ConvGeom& g) { return g.k == 5 && g.s == 2 && g.Cb >= 1 && g.Cb <= 4 && g.Cs % 32 == 0; }

}  // namespace

// ------------------------------------------------------------------------------------------------
// Tensor-core form of thin_fprop (mode 1).  The FP32 kernel above is issue-bound: 25*CB FMAs per
// output pixel and channel.  Here the 25*CB patch values of an output pixel become one row of a
// K-major SWIZZLE_128B A tile (each thread gathers its own row from the staged patch with scalar LDS
// and writes it as 16-byte chunks), the TF32 weights [c_small][(tap, c_big)] sit in a B tile for the
// whole kernel, and an 8 x 32 output tile is two tcgen05.mma M-tiles (K padded to a multiple of 32).
// Epilogue = the shared transposing store (bias, LeakyReLU, derivative mask).
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint64_t thin_desc_sw128(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

template <int CB>
struct ThinTcSmem {
  static constexpr int NKC = (25 * CB + 31) / 32;           // 32-float k-chunks
  static constexpr int A_BYTES = 2 * NKC * 128 * 128;       // two M-tiles (8 rows x 32 pixels)
  static constexpr int B_BYTES = NKC * 32 * 128;            // N = 32 channels
  static constexpr int OFF_B = A_BYTES;
  static constexpr int OFF_PATCH = OFF_B + B_BYTES;
  static constexpr int PATCH_BYTES = 2 * CB * PROWS * PCOLS * 4;
  static constexpr int OFF_BAR = OFF_PATCH + ((PATCH_BYTES + 15) & ~15);
  static constexpr int TOTAL = OFF_BAR + 32;
};

template <int CB>
__global__ void __launch_bounds__(256) thin_fprop_tc_kernel(const ThinGeo g, const float* __restrict__ wft,
                                                            const float* __restrict__ bias, float* __restrict__ out,
                                                            const float* __restrict__ dact, int act, const TileIter it,
                                                            float* __restrict__ colsum) {
  bn_pdl_trigger();
  bn_pdl_wait();
  using namespace bn_tc;
  using S = ThinTcSmem<CB>;
  constexpr int NKC = S::NKC, KTOT = 25 * CB;
  extern __shared__ __align__(1024) unsigned char tsm[];
  typedef float (*Patch)[PROWS][PCOLS];
  Patch patch[2] = {reinterpret_cast<Patch>(tsm + S::OFF_PATCH),
                    reinterpret_cast<Patch>(tsm + S::OFF_PATCH + CB * PROWS * PCOLS * 4)};
  uint64_t* accum_bar = reinterpret_cast<uint64_t*>(tsm + S::OFF_BAR);
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(accum_bar + 1);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int c0 = blockIdx.y * 32;                          // channel group of this block
  // B tile: rows n = channel, k-chunk kc, 16-byte chunk c at position c ^ (n & 7)
  for (int i = tid; i < NKC * 32 * 8; i += 256) {
    const int kc = i / 256, rem = i - kc * 256;
    const int n = rem >> 3, c = rem & 7;
    float v[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int k = kc * 32 + c * 4 + e;
      v[e] = k < KTOT ? __ldg(wft + (long long)(c0 + n) * KTOT + k) : 0.f;
    }
    *reinterpret_cast<float4*>(tsm + S::OFF_B + kc * 4096 + n * 128 + ((c ^ (n & 7)) << 4)) =
        make_float4(v[0], v[1], v[2], v[3]);
  }
  if (tid == 0) {
    mbar_init(smem_u32(accum_bar), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) tmem_alloc<64>(smem_u32(tmem_ptr));
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  const uint32_t sa = smem_u32(tsm), sb = sa + S::OFF_B;
  const int trow = tid >> 5, tcol = tid & 31;              // output pixel of this thread inside the 8 x 32 tile
  const int mt = trow >> 2, arow = (trow & 3) * 32 + tcol; // M-tile and row inside it

  long long t = blockIdx.x;
  int f, y0, x0;
  if (t < it.total) {
    it.decode(t, f, y0, x0);
    issue_patch<CB>(patch[0], g, f, y0, x0, tid);
  }
  cp_commit();
  // fused column sums of the stored image (the bias gradient of the layer below when `out` is a gradient
  // image): per-lane accumulators over all tiles of this persistent CTA, one flush at the end
  float4 cacc = make_float4(0.f, 0.f, 0.f, 0.f);
  uint32_t phase = 0;
  for (int buf = 0; t < it.total; t += gridDim.x, buf ^= 1, phase ^= 1) {
    it.decode(t, f, y0, x0);
    const long long tn = t + gridDim.x;
    if (tn < it.total) {
      int fn, yn, xn;
      it.decode(tn, fn, yn, xn);
      issue_patch<CB>(patch[buf ^ 1], g, fn, yn, xn, tid);
    }
    cp_commit();
    cp_wait<1>();
    __syncthreads();                   // patch[buf] landed; previous tile's epilogue tiles (aliasing A) are idle
    // ---- this thread's im2col row: k = (ky*5 + kx)*CB + cb  <-  patch[cb][2*trow + ky][2*tcol + kx]
    {
      unsigned char* abase = tsm + (mt * NKC) * 16384 + arow * 128;
#pragma unroll
      for (int kc = 0; kc < NKC; ++kc) {
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          float v[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int k = kc * 32 + c * 4 + e;            // compile-time
            if (k < KTOT) {
              const int tap = k / CB, cb = k - tap * CB;
              v[e] = patch[buf][cb][2 * trow + tap / 5][2 * tcol + tap % 5];
            } else {
              v[e] = 0.f;
            }
          }
          *reinterpret_cast<float4*>(abase + kc * 16384 + ((c ^ (arow & 7)) << 4)) = make_float4(v[0], v[1], v[2], v[3]);
        }
      }
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    if (tid == 0) {
      tc_fence_after();
      constexpr uint32_t idesc = make_idesc(128, 32);
#pragma unroll
      for (int m = 0; m < 2; ++m)
#pragma unroll
        for (int kc = 0; kc < NKC; ++kc)
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_tf32(tmem_base + m * 32, thin_desc_sw128(sa + (m * NKC + kc) * 16384 + k * 32),
                      thin_desc_sw128(sb + kc * 4096 + k * 32), idesc, (kc | k) != 0 ? 1u : 0u);
      umma_commit(smem_u32(accum_bar));
    }
    mbar_wait(smem_u32(accum_bar), phase);
    tc_fence_after();
    // ---- epilogue: warps 0-3 drain M-tile 0, warps 4-7 M-tile 1; the A tiles are free again
    {
      const int q = warp & 3, m = warp >> 2;
      uint32_t r[32];
      tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + m * 32, r);
      tmem_ld_wait();
      const int oy = y0 + m * 4 + q, ox = x0 + lane;
      const bool valid = oy < g.Hs && ox < g.Ws;
      const long long idx = valid ? (((long long)f * g.Hs + oy) * g.Ws + ox) * g.Cs + c0 : -1;
      warp_store_rows32(out, dact, BN_LEAK, idx, r, bias ? bias + c0 : nullptr, act,
                        reinterpret_cast<float*>(tsm) + warp * 1024, lane, cacc, colsum != nullptr);
    }
    tc_fence_before();
  }
  __syncthreads();
  if (colsum != nullptr) {
    float* red = reinterpret_cast<float*>(tsm);           // [8 warps][32 columns]; the A tiles are idle
#pragma unroll
    for (int o = 8; o <= 16; o <<= 1) {
      cacc.x += __shfl_xor_sync(0xffffffffu, cacc.x, o);
      cacc.y += __shfl_xor_sync(0xffffffffu, cacc.y, o);
      cacc.z += __shfl_xor_sync(0xffffffffu, cacc.z, o);
      cacc.w += __shfl_xor_sync(0xffffffffu, cacc.w, o);
    }
    if (lane < 8) *(reinterpret_cast<float4*>(red + warp * 32) + lane) = cacc;
    __syncthreads();
    if (tid < 32) {
      float s = 0.f;
#pragma unroll
      for (int w8 = 0; w8 < 8; ++w8) s += red[w8 * 32 + tid];
      atomicAdd(colsum + c0 + tid, s);
    }
  }
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc<64>(tmem_base);
  }
}

template <int CB>
static int launch_thin_fprop_tc(const ThinGeo& t, const float* wft, const float* bias, float* out, const float* dact,
                                int act, const TileIter& it, int cgroups, cudaStream_t st, float* colsum) {
  using S = ThinTcSmem<CB>;
  auto kern = thin_fprop_tc_kernel<CB>;
  static bool configured = false;
  if (!configured) {
    BN_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, S::TOTAL));
    configured = true;
  }
  // residency: 64 TMEM columns and S::TOTAL bytes per CTA
  int per_sm = (227 * 1024) / (S::TOTAL + 1024);
  per_sm = per_sm > 2 ? 2 : (per_sm < 1 ? 1 : per_sm);       // ~120 registers x 256 threads: two CTAs per SM
  long long blocks = it.total < 148LL * per_sm ? it.total : 148LL * per_sm;
  BN_CUDA(bn_launch(kern, dim3((unsigned)blocks, cgroups), 256, S::TOTAL, st, t, wft, bias, out, dact, act, it, colsum));
  BN_LAUNCHED();
  return 0;
}

template <int CB>
static int launch_thin_fprop(const ThinGeo& t, const float* wf, const float* bias, float* out, const float* dact,
                             int act, const TileIter& it, int cgroups, cudaStream_t st) {
  const size_t smem = (size_t)2 * CB * PROWS * PCOLS * sizeof(float);
  auto kern = thin_fprop_kernel<CB>;
  static bool configured = false;
  if (!configured) {
    BN_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = true;
  }
  long long blocks = it.total < 2 * 148 ? it.total : 2 * 148;      // two resident blocks per SM (128 registers x 256 threads)
  BN_CUDA(bn_launch(kern, dim3((unsigned)blocks, cgroups), 256, smem, st, t, wf, bias, out, dact, act, it));
  BN_LAUNCHED();
  return 0;
}

int bn_launch_thin_fprop(const ImgView& big, const ConvGeom& g, const float* wf, const float* wft, const float* bias,
                         float* out, const float* dact, int act, int n, cudaStream_t st,
                         const unsigned char* big_u8, float* colsum, int* colsum_fused) {
  if (colsum_fused) *colsum_fused = 0;
  if (!fast_geom(g) || n <= 0) return 1;
  ThinGeo t;
  t.big8 = big_u8;
  t.big = big; t.Hs = g.Hs; t.Ws = g.Ws; t.Cs = g.Cs; t.pt = g.pt; t.pl = g.pl; t.n = n;
  TileIter it;
  it.tiles_x = bn_cdiv(g.Ws, TW);
  it.tiles_per_frame = it.tiles_x * bn_cdiv(g.Hs, TH);
  it.total = (long long)it.tiles_per_frame * n;
  static const bool tc_off = [] { const char* e = getenv("BN_THIN_TC"); return e && e[0] == '0'; }();
  if (wft != nullptr && !tc_off && !((uintptr_t)out & 15) && !(dact && ((uintptr_t)dact & 15)) &&
      !(bias && ((uintptr_t)bias & 15))) {
    // tensor-core form (TF32): wft = K-major weights [c_small][(tap, c_big)]
    if (colsum_fused) *colsum_fused = colsum != nullptr;
    {
      const int r2 = bn_launch_thin_fprop_tc2(big, big_u8, g, wft, bias, out, dact, act, n, colsum, st);
      if (r2 <= 0) return r2;
    }
    switch (g.Cb) {
      case 1: return launch_thin_fprop_tc<1>(t, wft, bias, out, dact, act, it, g.Cs / 32, st, colsum);
      case 2: return launch_thin_fprop_tc<2>(t, wft, bias, out, dact, act, it, g.Cs / 32, st, colsum);
      case 3: return launch_thin_fprop_tc<3>(t, wft, bias, out, dact, act, it, g.Cs / 32, st, colsum);
      default: return launch_thin_fprop_tc<4>(t, wft, bias, out, dact, act, it, g.Cs / 32, st, colsum);
    }
  }
  switch (g.Cb) {
    case 1: return launch_thin_fprop<1>(t, wf, bias, out, dact, act, it, g.Cs / 32, st);
    case 2: return launch_thin_fprop<2>(t, wf, bias, out, dact, act, it, g.Cs / 32, st);
    case 3: return launch_thin_fprop<3>(t, wf, bias, out, dact, act, it, g.Cs / 32, st);
    default: return launch_thin_fprop<4>(t, wf, bias, out, dact, act, it, g.Cs / 32, st);
  }
}

template <int CB>
static int launch_thin_wgrad(const ThinGeo& t, const float* small, float* partial, int blocks, const TileIter& it,
                             int cgroups, cudaStream_t st) {
  const size_t smem = ((size_t)2 * CB * PROWS * PCOLS + (size_t)8 * 25 * CB * 32) * sizeof(float);
  auto kern = thin_wgrad_kernel<CB>;
  static bool configured = false;
  if (!configured) {
    BN_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = true;
  }
  BN_CUDA(bn_launch(kern, dim3(blocks, cgroups), 256, smem, st, t, small, partial, it));
  BN_LAUNCHED();
  return 0;
}

int bn_launch_thin_wgrad(const ImgView& big, const float* small, const ConvGeom& g, int n, float* partial,
                         size_t partial_floats, float* grad, cudaStream_t st) {
  if (!fast_geom(g) || n <= 0 || grad == nullptr) return grad == nullptr ? 0 : 1;
  ThinGeo t;
  t.big = big; t.Hs = g.Hs; t.Ws = g.Ws; t.Cs = g.Cs; t.pt = g.pt; t.pl = g.pl; t.n = n;
  TileIter it;
  it.tiles_x = bn_cdiv(g.Ws, TW);
  it.tiles_per_frame = it.tiles_x * bn_cdiv(g.Hs, TH);
  it.total = (long long)it.tiles_per_frame * n;
  const int Ktot = 25 * g.Cb;
  long long blocks = it.total < 2 * 148 ? it.total : 2 * 148;
  long long cap = (long long)(partial_floats / ((size_t)Ktot * g.Cs));
  if (blocks > cap) blocks = cap;
  if (blocks < 1) return 1;
  int r;
  switch (g.Cb) {
    case 1: r = launch_thin_wgrad<1>(t, small, partial, (int)blocks, it, g.Cs / 32, st); break;
    case 2: r = launch_thin_wgrad<2>(t, small, partial, (int)blocks, it, g.Cs / 32, st); break;
    case 3: r = launch_thin_wgrad<3>(t, small, partial, (int)blocks, it, g.Cs / 32, st); break;
    default: r = launch_thin_wgrad<4>(t, small, partial, (int)blocks, it, g.Cs / 32, st); break;
  }
  if (r) return r;
  return bn_launch_wgrad_reduce(partial, (int)blocks, Ktot, g.Cs, g.Cb, 25, g.d_fprop, grad, st);
}

template <int CB, int PTO, int PLO>
static int launch_dgrad5(const Dg5Args& a, dim3 grid, int tiles_x, size_t smem, cudaStream_t st) {
  auto kern = thin_dgrad5_kernel<CB, PTO, PLO>;
  static bool configured = false;
  if (!configured) {
    BN_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
    configured = true;
  }
  BN_CUDA(bn_launch(kern, grid, 256, smem, st, a, tiles_x));
  BN_LAUNCHED();
  return 0;
}

template <int CB>
static int launch_dgrad5_p(const Dg5Args& a, dim3 grid, int tiles_x, size_t smem, cudaStream_t st) {
  switch ((a.pt & 1) * 2 + (a.pl & 1)) {
    case 0: return launch_dgrad5<CB, 0, 0>(a, grid, tiles_x, smem, st);
    case 1: return launch_dgrad5<CB, 0, 1>(a, grid, tiles_x, smem, st);
    case 2: return launch_dgrad5<CB, 1, 0>(a, grid, tiles_x, smem, st);
    default: return launch_dgrad5<CB, 1, 1>(a, grid, tiles_x, smem, st);
  }
}

int bn_launch_thin_dgrad5(const float* small, const ConvGeom& g, const float* wd, const float* bias, int n,
                          float* xhat_ws, float* xhat_user, const float* target, const float* mask,
                          int chunk_size, int frame_offset, int n_total, float grad_coef, double* sse,
                          float* dpre, cudaStream_t st) {
  if (!(g.k == 5 && g.s == 2 && g.Cb >= 1 && g.Cb <= 4 && g.Cs % 4 == 0) || n <= 0 || n > 65535) return 1;
  if (g.pt < 0 || g.pl < 0 || g.pt > 4 || g.pl > 4) return 1;
  // every output pixel must be covered by a small pixel of the tile grid
  if (g.Hb > 2 * g.Hs + 2 || g.Wb > 2 * g.Ws + 2) return 1;
  size_t smem = ((size_t)DP * DP * (g.Cs + 4) + (size_t)25 * g.Cb * g.Cs) * sizeof(float);
  if (smem > 100 * 1024) return 1;
  Dg5Args a;
  a.small = small; a.Hs = g.Hs; a.Ws = g.Ws; a.Cs = g.Cs; a.Hb = g.Hb; a.Wb = g.Wb; a.pt = g.pt; a.pl = g.pl;
  a.n = n; a.wd = wd; a.bias = bias; a.xhat_ws = xhat_ws; a.xhat_user = xhat_user; a.target = target;
  a.mask = mask; a.n_total = n_total > 0 ? n_total : n; a.frame_offset = frame_offset;
  a.chunk_size = chunk_size > 0 ? chunk_size : a.n_total; a.coef = grad_coef; a.sse = sse; a.dpre = dpre;
  // tiles of DT x DT small pixels = 2*DT x 2*DT output pixels
  const int tiles_x = bn_cdiv(bn_cdiv(g.Wb, 2), DT), tiles_y = bn_cdiv(bn_cdiv(g.Hb, 2), DT);
  dim3 grid(tiles_x * tiles_y, n);
  switch (g.Cb) {
    case 1: return launch_dgrad5_p<1>(a, grid, tiles_x, smem, st);
    case 2: return launch_dgrad5_p<2>(a, grid, tiles_x, smem, st);
    case 3: return launch_dgrad5_p<3>(a, grid, tiles_x, smem, st);
    default: return launch_dgrad5_p<4>(a, grid, tiles_x, smem, st);
  }
}
