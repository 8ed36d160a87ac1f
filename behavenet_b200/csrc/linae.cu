// Linear autoencoder (reference behavenet/models/aes.py:491-613, model_type = 'linear'): z = x W^T + b,
// x_hat = z W + c with the decoder using the TRANSPOSED encoder weights plus its own bias (aes.py:583-586,
// 603-606, the form AE.build_model constructs, aes.py:684-687), trained on the masked MSE of aes.py:751-769.
//
// The whole model is a rank-L factorisation of a (frames x pixels) matrix with L ~ 10: 2 * L flops per byte of
// frame data, i.e. HBM-bound on streaming the frames.  Three launches per training call, each reading the
// frames once:
//   linae_encode_kernel      : block per 1 / 2 / 4 frames, z = x W^T + b           (reads x)
//   linae_decode_loss_kernel : thread per pixel, frames in a loop: x_hat, masked squared error per reference
//                              chunk, g = dL/dx_hat (stored), dc += g, dW += z^T g  (reads x [, mask], writes g)
//   linae_encode_kernel on g : dz = g W^T                                          (reads g)
//   linae_encode_bwd_kernel  : thread per pixel: dW += dz^T x, db = colsum(dz)     (reads x)
// i.e. four launches per training call that stream 5 x (frames x pixels) floats.  W (L x P floats, < 1 MB) is
// re-read through L1 / L2.  Gradients are ACCUMULATED into the caller's tensors.  A first version reduced dz
// across the pixels of a block inside the decode kernel's frame loop (L warp reductions and two barriers per
// frame): 0.24 ms for 256 frames of 128 x 128, latency-bound at 0.03 of the HBM rate.
#include <string.h>

#include "../../include/behavenet_b200.h"
#include "bn_common.cuh"

namespace {

__device__ __forceinline__ float warp_sum_f(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// z[f][l] = b[l] + sum_p x[f][p] W[l][p] for the F frames of a block, 16 latents per pass over the pixels (one
// pass for L <= 16): every W value fetched (through L1 / L2: the matrix is re-read by every block) is used for F
// frames, and the loads of two pixel steps are issued before their FMAs (the first version had one step's loads
// outstanding per thread: 350 us for 2048 frames, latency-bound).  Also run on the gradient image (x = g, no
// bias) for dz = g W^T.
template <int F>
__global__ void __launch_bounds__(256) linae_encode_kernel(const float* __restrict__ x, const float* __restrict__ W,
                                                           const float* __restrict__ b, float* __restrict__ z,
                                                           int n, int P, int L) {
  bn_pdl_trigger();
  bn_pdl_wait();
  __shared__ float red[8][F * 16];
  const int f0 = blockIdx.x * F, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const float* xf[F];
#pragma unroll
  for (int q = 0; q < F; ++q) xf[q] = x + (long long)min(f0 + q, n - 1) * P;
  for (int l0 = 0; l0 < L; l0 += 16) {
    const float* wr[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) wr[j] = W + (long long)min(l0 + j, L - 1) * P;
    float acc[F][16];
#pragma unroll
    for (int q = 0; q < F; ++q)
#pragma unroll
      for (int j = 0; j < 16; ++j) acc[q][j] = 0.f;
    int p = tid;
    for (; p + 256 < P; p += 512) {
      float xa[F], xb[F], wa[16], wb[16];
#pragma unroll
      for (int q = 0; q < F; ++q) { xa[q] = __ldg(xf[q] + p); xb[q] = __ldg(xf[q] + p + 256); }
#pragma unroll
      for (int j = 0; j < 16; ++j) { wa[j] = __ldg(wr[j] + p); wb[j] = __ldg(wr[j] + p + 256); }
#pragma unroll
      for (int j = 0; j < 16; ++j)
#pragma unroll
        for (int q = 0; q < F; ++q) acc[q][j] = fmaf(xb[q], wb[j], fmaf(xa[q], wa[j], acc[q][j]));
    }
    for (; p < P; p += 256) {
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const float wv = __ldg(wr[j] + p);
#pragma unroll
        for (int q = 0; q < F; ++q) acc[q][j] = fmaf(__ldg(xf[q] + p), wv, acc[q][j]);
      }
    }
#pragma unroll
    for (int q = 0; q < F; ++q)
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const float s = warp_sum_f(acc[q][j]);
        if (lane == 0) red[warp][q * 16 + j] = s;
      }
    __syncthreads();
    if (tid < F * 16) {
      const int q = tid >> 4, j = tid & 15;
      if (f0 + q < n && l0 + j < L) {
        float s = b ? b[l0 + j] : 0.f;
#pragma unroll
        for (int w = 0; w < 8; ++w) s += red[w][tid];
        z[(long long)(f0 + q) * L + l0 + j] = s;
      }
    }
    __syncthreads();
  }
}

struct LinDecArgs {
  const float* x;        // (n, P) targets (= the frames)
  const float* mask;     // (n, P) or NULL
  const float* W;        // (L, P)
  const float* c;        // (P) decoder bias
  const float* z;        // (n, L)
  float* xhat;           // (n, P) or NULL
  float* g;              // (n, P) dL/dx_hat, or NULL (forward / loss only)
  float* gW;             // (L, P) accumulated, or NULL
  float* gc;             // (P) accumulated, or NULL
  double* sse;           // per reference chunk of the WHOLE batch, accumulated, or NULL
  int n, P, L, chunk_size, frame_offset, n_total, fsplit;
};

#define LIN_FB 32      // frames whose latents are staged in shared memory per barrier (a multiple of 4)

// thread = pixel (its W column in registers), block = 256 pixels x the frames of one frame split.  Nothing in
// the frame loop crosses threads: the squared error is kept per thread and reduced when the reference chunk
// changes, g goes to memory for the dz pass (the encode kernel run on g).  Per-frame constants (chunk index,
// gradient coefficient) are staged next to the latents; the frames are walked four at a time with their x / mask
// loads issued up front.
template <int LMAX>
__global__ void __launch_bounds__(256) linae_decode_loss_kernel(const LinDecArgs a) {
  bn_pdl_trigger();
  bn_pdl_wait();
  __shared__ __align__(16) float zs[LIN_FB][LMAX];
  __shared__ float coef[LIN_FB];
  __shared__ int chunk_of[LIN_FB];
  __shared__ float red[8];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int p = blockIdx.x * 256 + tid;
  const bool in = p < a.P;
  const int pc = in ? p : 0;
  const int per = (a.n + a.fsplit - 1) / a.fsplit;
  const int f0 = blockIdx.y * per, f1 = min(a.n, f0 + per);
  const int L = a.L;
  float w[LMAX], gw[LMAX];
#pragma unroll
  for (int l = 0; l < LMAX; ++l) {
    w[l] = (in && l < L) ? __ldg(a.W + (long long)l * a.P + p) : 0.f;
    gw[l] = 0.f;
  }
  const float cb = in ? __ldg(a.c + p) : 0.f;
  float gcb = 0.f, e2sum = 0.f;
  const bool train = a.g != nullptr;
  const bool want_loss = a.sse != nullptr || train;
  int cur_chunk = -1;

  auto flush_chunk = [&]() {      // block-wide: every thread takes the same branch (chunks depend on f only)
    const float s = warp_sum_f(e2sum);
    if (lane == 0) red[warp] = s;
    __syncthreads();
    if (tid == 0 && a.sse && cur_chunk >= 0) {
      double t = 0.0;
#pragma unroll
      for (int q = 0; q < 8; ++q) t += (double)red[q];
      atomicAdd(a.sse + cur_chunk, t);
    }
    __syncthreads();
    e2sum = 0.f;
  };

  for (int fb = f0; fb < f1; fb += LIN_FB) {
    const int nf = min(LIN_FB, f1 - fb);
    __syncthreads();
    for (int i = tid; i < LIN_FB * LMAX; i += 256) {
      const int q = i / LMAX, l = i - q * LMAX;
      zs[q][l] = (q < nf && l < L) ? a.z[(long long)(fb + q) * L + l] : 0.f;
    }
    if (tid < LIN_FB) {
      const int gf = a.frame_offset + fb + min(tid, nf - 1);
      const int chunk = a.chunk_size > 0 ? gf / a.chunk_size : 0;
      const int clen = a.chunk_size > 0 ? min(a.chunk_size, a.n_total - chunk * a.chunk_size) : a.n_total;
      chunk_of[tid] = chunk;
      // dL/dx_hat of the chunk's mean squared error (losses.mse: mean over every element, masked or not)
      coef[tid] = 2.f / ((float)a.P * (float)clen);
    }
    __syncthreads();
    for (int q0 = 0; q0 < nf; q0 += 4) {
      float xv[4], mv[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const long long row = (long long)(fb + min(q0 + u, nf - 1)) * a.P + pc;
        xv[u] = want_loss ? __ldg(a.x + row) : 0.f;
        mv[u] = (want_loss && a.mask) ? __ldg(a.mask + row) : 1.f;
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int q = q0 + u;
        if (q >= nf) break;
        float zr[LMAX];
#pragma unroll
        for (int l = 0; l < LMAX; l += 4) {
          const float4 t = *reinterpret_cast<const float4*>(&zs[q][l]);
          zr[l] = t.x; zr[l + 1] = t.y; zr[l + 2] = t.z; zr[l + 3] = t.w;
        }
        float xh = cb;
#pragma unroll
        for (int l = 0; l < LMAX; ++l) xh = fmaf(zr[l], w[l], xh);
        const long long row = (long long)(fb + q) * a.P + pc;
        if (a.xhat && in) a.xhat[row] = xh;
        if (!want_loss) continue;
        if (chunk_of[q] != cur_chunk) {
          if (cur_chunk >= 0) flush_chunk();
          cur_chunk = chunk_of[q];
        }
        const float m = in ? mv[u] : 0.f;
        const float err = xh - xv[u];
        e2sum = fmaf(err * err, m, e2sum);
        if (train) {
          const float g = coef[q] * err * m;
          if (in) a.g[row] = g;
          gcb += g;
#pragma unroll
          for (int l = 0; l < LMAX; ++l) gw[l] = fmaf(zr[l], g, gw[l]);
        }
      }
    }
  }
  if (want_loss && cur_chunk >= 0) flush_chunk();
  if (train && in) {
    if (a.gc) atomicAdd(a.gc + p, gcb);
    if (a.gW) {
#pragma unroll
      for (int l = 0; l < LMAX; ++l)
        if (l < L) atomicAdd(a.gW + (long long)l * a.P + p, gw[l]);
    }
  }
}

// gW[l][p] += sum_f dz[f][l] x[f][p]; gb[l] += sum_f dz[f][l] (blocks of pixel tile 0)
template <int LMAX>
__global__ void __launch_bounds__(256) linae_encode_bwd_kernel(const float* __restrict__ x, const float* __restrict__ dz,
                                                               float* __restrict__ gW, float* __restrict__ gb, int n, int P,
                                                               int L, int fsplit) {
  bn_pdl_trigger();
  bn_pdl_wait();
  __shared__ __align__(16) float ds[LIN_FB][LMAX];
  const int tid = threadIdx.x;
  const int p = blockIdx.x * 256 + tid;
  const bool in = p < P;
  const int pc = in ? p : 0;
  const int per = (n + fsplit - 1) / fsplit;
  const int f0 = blockIdx.y * per, f1 = min(n, f0 + per);
  float acc[LMAX];
#pragma unroll
  for (int l = 0; l < LMAX; ++l) acc[l] = 0.f;
  float bsum = 0.f;
  for (int fb = f0; fb < f1; fb += LIN_FB) {
    const int nf = min(LIN_FB, f1 - fb);
    __syncthreads();
    for (int i = tid; i < LIN_FB * LMAX; i += 256) {
      const int q = i / LMAX, l = i - q * LMAX;
      ds[q][l] = (q < nf && l < L) ? dz[(long long)(fb + q) * L + l] : 0.f;
    }
    __syncthreads();
    for (int q0 = 0; q0 < LIN_FB; q0 += 4) {       // rows beyond nf hold zeros
      if (q0 >= nf) break;
      float xv[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) xv[u] = __ldg(x + (long long)(fb + min(q0 + u, nf - 1)) * P + pc);
#pragma unroll
      for (int u = 0; u < 4; ++u) {
#pragma unroll
        for (int l = 0; l < LMAX; l += 4) {
          const float4 t = *reinterpret_cast<const float4*>(&ds[q0 + u][l]);
          acc[l] = fmaf(t.x, xv[u], acc[l]);
          acc[l + 1] = fmaf(t.y, xv[u], acc[l + 1]);
          acc[l + 2] = fmaf(t.z, xv[u], acc[l + 2]);
          acc[l + 3] = fmaf(t.w, xv[u], acc[l + 3]);
        }
        if (blockIdx.x == 0 && tid < L) bsum += ds[q0 + u][tid];
      }
    }
  }
  if (in && gW) {
#pragma unroll
    for (int l = 0; l < LMAX; ++l)
      if (l < L) atomicAdd(gW + (long long)l * P + p, acc[l]);
  }
  if (gb && blockIdx.x == 0 && tid < L) atomicAdd(gb + tid, bsum);
}

int frame_splits(int n, int P) {
  // 256-pixel tiles x frame splits: ~4 blocks per SM's worth of resident work, at least 32 frames per block
  const int tiles = bn_cdiv(P, 256);
  int fs = bn_cdiv(6 * 148, tiles);
  if (fs > bn_cdiv(n, LIN_FB)) fs = bn_cdiv(n, LIN_FB);
  if (fs > 128) fs = 128;
  return fs < 1 ? 1 : fs;
}

int launch_encode(const float* x, const float* W, const float* b, float* z, int n, int P, int L, cudaStream_t st) {
  // two frames per block (W fetches shared) once that still leaves two waves of blocks
  if (n >= 4 * 148) {
    BN_CUDA(bn_launch(linae_encode_kernel<2>, dim3(bn_cdiv(n, 2)), 256, 0, st, x, W, b, z, n, P, L));
  } else {
    BN_CUDA(bn_launch(linae_encode_kernel<1>, dim3(n), 256, 0, st, x, W, b, z, n, P, L));
  }
  BN_LAUNCHED();
  return 0;
}

template <int LMAX>
int launch_decode(const LinDecArgs& a, cudaStream_t st) {
  BN_CUDA(bn_launch(linae_decode_loss_kernel<LMAX>, dim3(bn_cdiv(a.P, 256), a.fsplit), 256, 0, st, a));
  BN_LAUNCHED();
  return 0;
}

template <int LMAX>
int launch_encode_bwd(const float* x, const float* dz, float* gW, float* gb, int n, int P, int L, int fs,
                      cudaStream_t st) {
  BN_CUDA(bn_launch(linae_encode_bwd_kernel<LMAX>, dim3(bn_cdiv(P, 256), fs), 256, 0, st, x, dz, gW, gb, n, P, L, fs));
  BN_LAUNCHED();
  return 0;
}

}  // namespace

extern "C" size_t bn_linae_workspace_bytes(int n, int P, int L) {
  return n > 0 && L > 0 && P > 0 ? sizeof(float) * ((size_t)2 * n * L + (size_t)n * P) : 0;
}

extern "C" int bn_linae_forward(int n, int P, int L, const float* d_x, const float* d_W, const float* d_b,
                                const float* d_c, float* d_z, float* d_xhat, void* stream) {
  if (!d_x || !d_W || !d_z) BN_FAIL("bn_linae_forward: null argument");
  if (P < 1 || L < 1 || L > 64) BN_FAIL("bn_linae_forward: P=%d L=%d (1 <= L <= 64)", P, L);
  if (d_xhat && !d_c) BN_FAIL("bn_linae_forward: the decoder needs its bias");
  if (n <= 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  BN_TRY(launch_encode(d_x, d_W, d_b, d_z, n, P, L, st));
  if (!d_xhat) return 0;
  LinDecArgs a;
  memset(&a, 0, sizeof(a));
  a.x = d_x; a.W = d_W; a.c = d_c; a.z = d_z; a.xhat = d_xhat;
  a.n = n; a.P = P; a.L = L; a.n_total = n; a.fsplit = frame_splits(n, P);
  return L <= 16 ? launch_decode<16>(a, st) : (L <= 32 ? launch_decode<32>(a, st) : launch_decode<64>(a, st));
}

extern "C" int bn_linae_decode(int n, int P, int L, const float* d_z, const float* d_W, const float* d_c,
                               float* d_xhat, void* stream) {
  if (!d_z || !d_W || !d_c || !d_xhat) BN_FAIL("bn_linae_decode: null argument");
  if (P < 1 || L < 1 || L > 64) BN_FAIL("bn_linae_decode: P=%d L=%d (1 <= L <= 64)", P, L);
  if (n <= 0) return 0;
  LinDecArgs a;
  memset(&a, 0, sizeof(a));
  a.W = d_W; a.c = d_c; a.z = d_z; a.xhat = d_xhat;
  a.n = n; a.P = P; a.L = L; a.n_total = n; a.fsplit = frame_splits(n, P);
  cudaStream_t st = (cudaStream_t)stream;
  return L <= 16 ? launch_decode<16>(a, st) : (L <= 32 ? launch_decode<32>(a, st) : launch_decode<64>(a, st));
}

extern "C" int bn_linae_loss(int n, int P, int L, const float* d_x, const float* d_mask, const float* d_W,
                             const float* d_b, const float* d_c, int chunk_size, int frame_offset, int n_total,
                             void* d_ws, double* d_sse, float* d_gW, float* d_gb, float* d_gc, void* stream) {
  if (!d_x || !d_W || !d_c || !d_ws || !d_sse) BN_FAIL("bn_linae_loss: null argument");
  if (P < 1 || L < 1 || L > 64) BN_FAIL("bn_linae_loss: P=%d L=%d (1 <= L <= 64)", P, L);
  if (n_total <= 0) n_total = n;
  if (frame_offset < 0 || frame_offset + n > n_total) BN_FAIL("bn_linae_loss: frames [%d, %d) outside the batch of %d", frame_offset, frame_offset + n, n_total);
  if (n <= 0) return 0;
  const bool train = d_gW || d_gb || d_gc;
  cudaStream_t st = (cudaStream_t)stream;
  float* z = (float*)d_ws;
  float* dz = z + (size_t)n * L;
  float* g = dz + (size_t)n * L;
  BN_TRY(launch_encode(d_x, d_W, d_b, z, n, P, L, st));
  LinDecArgs a;
  memset(&a, 0, sizeof(a));
  a.x = d_x; a.mask = d_mask; a.W = d_W; a.c = d_c; a.z = z; a.g = train ? g : nullptr;
  a.gW = d_gW; a.gc = d_gc; a.sse = d_sse;
  a.n = n; a.P = P; a.L = L; a.chunk_size = chunk_size; a.frame_offset = frame_offset; a.n_total = n_total;
  a.fsplit = frame_splits(n, P);
  BN_TRY(L <= 16 ? launch_decode<16>(a, st) : (L <= 32 ? launch_decode<32>(a, st) : launch_decode<64>(a, st)));
  if (!train) return 0;
  BN_TRY(launch_encode(g, d_W, nullptr, dz, n, P, L, st));          // dz = g W^T
  const int fs = a.fsplit;
  return L <= 16 ? launch_encode_bwd<16>(d_x, dz, d_gW, d_gb, n, P, L, fs, st)
                 : (L <= 32 ? launch_encode_bwd<32>(d_x, dz, d_gW, d_gb, n, P, L, fs, st)
                            : launch_encode_bwd<64>(d_x, dz, d_gW, d_gb, n, P, L, fs, st));
}
