// Linear autoencoder (reference behavenet/models/aes.py:491-613, model_type = 'linear'): z = x W^T + b,
// x_hat = z W + c with the decoder using the TRANSPOSED encoder weights plus its own bias (aes.py:583-586,
// 603-606, the form AE.build_model constructs, aes.py:684-687), trained on the masked MSE of aes.py:751-769.
//
// The whole model is a rank-L factorisation of a (frames x pixels) matrix with L ~ 10: 2 * L flops per byte of
// frame data, i.e. HBM-bound on streaming the frames.  Three launches per training call, each reading the
// frames once:
//   linae_encode_kernel      : block per frame, z = x W^T + b                      (reads x)
//   linae_decode_loss_kernel : thread per pixel, frames in a loop: x_hat, masked squared error per reference
//                              chunk, g = dL/dx_hat, dc += g, dW += z^T g, dz = g W^T  (reads x [, mask])
//   linae_encode_bwd_kernel  : thread per pixel: dW += dz^T x, db = colsum(dz)     (reads x)
// W (L x P floats, < 1 MB) is re-read through L1 / L2.  Gradients are ACCUMULATED into the caller's tensors.
#include <string.h>

#include "../../include/behavenet_b200.h"
#include "bn_common.cuh"

namespace {

__device__ __forceinline__ float warp_sum_f(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// z[f][l] = b[l] + sum_p x[f][p] W[l][p]; one block per frame, latents in groups of 8
__global__ void __launch_bounds__(256) linae_encode_kernel(const float* __restrict__ x, const float* __restrict__ W,
                                                           const float* __restrict__ b, float* __restrict__ z,
                                                           int P, int L) {
  bn_pdl_trigger();
  bn_pdl_wait();
  __shared__ float red[8][8];
  const int f = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const float* xf = x + (long long)f * P;
  for (int l0 = 0; l0 < L; l0 += 8) {
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.f;
    for (int p = tid; p < P; p += 256) {
      const float xv = __ldg(xf + p);
#pragma unroll
      for (int j = 0; j < 8; ++j)
        if (l0 + j < L) acc[j] = fmaf(xv, __ldg(W + (long long)(l0 + j) * P + p), acc[j]);
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float s = warp_sum_f(acc[j]);
      if (lane == 0) red[warp][j] = s;
    }
    __syncthreads();
    if (tid < 8 && l0 + tid < L) {
      float s = b ? b[l0 + tid] : 0.f;
#pragma unroll
      for (int w = 0; w < 8; ++w) s += red[w][tid];
      z[(long long)f * L + l0 + tid] = s;
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
  float* dz;             // (n, L), zero-initialised, or NULL (forward only)
  float* gW;             // (L, P) accumulated, or NULL
  float* gc;             // (P) accumulated, or NULL
  double* sse;           // per reference chunk of the WHOLE batch, accumulated, or NULL
  int n, P, L, chunk_size, frame_offset, n_total, fsplit;
};

// thread = pixel, block = 256 pixels x the frames of one frame split
template <int LMAX>
__global__ void __launch_bounds__(256) linae_decode_loss_kernel(const LinDecArgs a) {
  bn_pdl_trigger();
  bn_pdl_wait();
  __shared__ float zs[LMAX];
  __shared__ float red[8][LMAX + 1];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int p = blockIdx.x * 256 + tid;
  const bool in = p < a.P;
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
  float gcb = 0.f;
  const bool train = a.dz != nullptr;
  double sse_acc = 0.0;         // thread 0: squared error of the current chunk seen by this block
  int cur_chunk = -1;
  for (int f = f0; f < f1; ++f) {
    if (tid < L) zs[tid] = a.z[(long long)f * L + tid];
    __syncthreads();
    float xh = cb;
#pragma unroll
    for (int l = 0; l < LMAX; ++l) xh = fmaf(zs[l < L ? l : 0], w[l], xh);
    if (a.xhat && in) a.xhat[(long long)f * a.P + p] = xh;
    if (a.sse || train) {
      const int gf = a.frame_offset + f;
      const int chunk = a.chunk_size > 0 ? gf / a.chunk_size : 0;
      const int clen = a.chunk_size > 0 ? min(a.chunk_size, a.n_total - chunk * a.chunk_size) : a.n_total;
      float err = 0.f, m = 0.f;
      if (in) {
        m = a.mask ? __ldg(a.mask + (long long)f * a.P + p) : 1.f;
        err = xh - __ldg(a.x + (long long)f * a.P + p);
      }
      const float e2 = err * err * m;
      // dL/dx_hat of the chunk's mean squared error (losses.mse: mean over every element, masked or not)
      const float g = 2.f * err * m / ((float)a.P * (float)clen);
      float s = warp_sum_f(e2);
      if (lane == 0) red[warp][LMAX] = s;
      if (train) {
        gcb += g;
#pragma unroll
        for (int l = 0; l < LMAX; ++l) {
          gw[l] = fmaf(zs[l < L ? l : 0], g, gw[l]);
          const float d = warp_sum_f(g * w[l]);
          if (lane == 0) red[warp][l] = d;
        }
      }
      __syncthreads();
      if (train && tid < L) {
        float d = 0.f;
#pragma unroll
        for (int q = 0; q < 8; ++q) d += red[q][tid];
        atomicAdd(a.dz + (long long)f * L + tid, d);
      }
      if (tid == 0 && a.sse) {
        if (chunk != cur_chunk) {
          if (cur_chunk >= 0) atomicAdd(a.sse + cur_chunk, sse_acc);
          cur_chunk = chunk;
          sse_acc = 0.0;
        }
        float t = 0.f;
#pragma unroll
        for (int q = 0; q < 8; ++q) t += red[q][LMAX];
        sse_acc += (double)t;
      }
    }
    __syncthreads();
  }
  if (tid == 0 && a.sse && cur_chunk >= 0) atomicAdd(a.sse + cur_chunk, sse_acc);
  if (train && in) {
    if (a.gc) atomicAdd(a.gc + p, gcb);
    if (a.gW) {
#pragma unroll
      for (int l = 0; l < LMAX; ++l)
        if (l < L) atomicAdd(a.gW + (long long)l * a.P + p, gw[l]);
    }
  }
}

// gW[l][p] += sum_f dz[f][l] x[f][p]; gb[l] += sum_f dz[f][l] (block (0, 0))
template <int LMAX>
__global__ void __launch_bounds__(256) linae_encode_bwd_kernel(const float* __restrict__ x, const float* __restrict__ dz,
                                                               float* __restrict__ gW, float* __restrict__ gb, int n, int P,
                                                               int L, int fsplit) {
  bn_pdl_trigger();
  bn_pdl_wait();
  __shared__ float ds[LMAX];
  const int tid = threadIdx.x;
  const int p = blockIdx.x * 256 + tid;
  const bool in = p < P;
  const int per = (n + fsplit - 1) / fsplit;
  const int f0 = blockIdx.y * per, f1 = min(n, f0 + per);
  float acc[LMAX];
#pragma unroll
  for (int l = 0; l < LMAX; ++l) acc[l] = 0.f;
  float bsum = 0.f;
  for (int f = f0; f < f1; ++f) {
    if (tid < L) ds[tid] = dz[(long long)f * L + tid];
    __syncthreads();
    const float xv = in ? __ldg(x + (long long)f * P + p) : 0.f;
#pragma unroll
    for (int l = 0; l < LMAX; ++l) acc[l] = fmaf(ds[l < L ? l : 0], xv, acc[l]);
    if (blockIdx.x == 0 && tid < L) bsum += ds[tid];
    __syncthreads();
  }
  if (in && gW) {
#pragma unroll
    for (int l = 0; l < LMAX; ++l)
      if (l < L) atomicAdd(gW + (long long)l * P + p, acc[l]);
  }
  if (gb && blockIdx.x == 0 && tid < L) atomicAdd(gb + tid, bsum);
}

int frame_splits(int n, int P) {
  const int tiles = bn_cdiv(P, 256);
  int fs = bn_cdiv(2 * 148, tiles);
  if (fs > n) fs = n;
  if (fs > 64) fs = 64;
  return fs < 1 ? 1 : fs;
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

extern "C" size_t bn_linae_workspace_bytes(int n, int L) {
  return n > 0 && L > 0 ? 2 * sizeof(float) * (size_t)n * L : 0;
}

extern "C" int bn_linae_forward(int n, int P, int L, const float* d_x, const float* d_W, const float* d_b,
                                const float* d_c, float* d_z, float* d_xhat, void* stream) {
  if (!d_x || !d_W || !d_z) BN_FAIL("bn_linae_forward: null argument");
  if (P < 1 || L < 1 || L > 64) BN_FAIL("bn_linae_forward: P=%d L=%d (1 <= L <= 64)", P, L);
  if (d_xhat && !d_c) BN_FAIL("bn_linae_forward: the decoder needs its bias");
  if (n <= 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  BN_CUDA(bn_launch(linae_encode_kernel, dim3(n), 256, 0, st, d_x, d_W, d_b, d_z, P, L));
  BN_LAUNCHED();
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
  BN_CUDA(bn_launch(linae_encode_kernel, dim3(n), 256, 0, st, d_x, d_W, d_b, z, P, L));
  BN_LAUNCHED();
  if (train) BN_CUDA(cudaMemsetAsync(dz, 0, sizeof(float) * (size_t)n * L, st));
  LinDecArgs a;
  memset(&a, 0, sizeof(a));
  a.x = d_x; a.mask = d_mask; a.W = d_W; a.c = d_c; a.z = z; a.dz = train ? dz : nullptr;
  a.gW = d_gW; a.gc = d_gc; a.sse = d_sse;
  a.n = n; a.P = P; a.L = L; a.chunk_size = chunk_size; a.frame_offset = frame_offset; a.n_total = n_total;
  a.fsplit = frame_splits(n, P);
  BN_TRY(L <= 16 ? launch_decode<16>(a, st) : (L <= 32 ? launch_decode<32>(a, st) : launch_decode<64>(a, st)));
  if (!train) return 0;
  const int fs = a.fsplit;
  return L <= 16 ? launch_encode_bwd<16>(d_x, dz, d_gW, d_gb, n, P, L, fs, st)
                 : (L <= 32 ? launch_encode_bwd<32>(d_x, dz, d_gW, d_gb, n, P, L, fs, st)
                            : launch_encode_bwd<64>(d_x, dz, d_gW, d_gb, n, P, L, fs, st));
}
