// Bandwidth-bound kernels of the CAE hot path: thin last decoder layer with the fused
// reconstruction-loss epilogue, linear heads, weight packing, bias-gradient column sums.
#include "cae_kernels.cuh"

namespace {

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// ------------------------------------------------------------------------------------------------
// Last decoder layer: ConvTranspose2d -> crop -> Sigmoid (aes.py:326-330, 463-470) with
// losses.mse / losses.gaussian_ll (losses.py:36-96) fused into the epilogue.
// One thread per output pixel, all C_big (<= 4) channels.
// ------------------------------------------------------------------------------------------------
struct ThinArgs {
  const float* small;
  int Hs, Ws, Cs, Hb, Wb, Cb, k, s, pt, pl, n;
  const float* wd;     // [(tap, cs)][cb]
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

constexpr int THIN_SSE_SLOTS = 8;

__global__ void __launch_bounds__(256) thin_dgrad_kernel(const ThinArgs a) {
  bn_pdl_trigger();
  bn_pdl_wait();
  extern __shared__ float wsm[];
  __shared__ double sse_sm[THIN_SSE_SLOTS];
  const int t = threadIdx.x;
  const int nw = a.k * a.k * a.Cs * a.Cb;
  for (int i = t; i < nw; i += 256) wsm[i] = a.wd[i];
  if (t < THIN_SSE_SLOTS) sse_sm[t] = 0.0;
  __syncthreads();
  const long long npix = (long long)a.n * a.Hb * a.Wb;
  const long long p = (long long)blockIdx.x * 256 + t;
  const bool valid = p < npix;
  const long long p0 = (long long)blockIdx.x * 256;
  const int f_first = (int)(p0 / ((long long)a.Hb * a.Wb));
  const int chunk_first = (f_first + a.frame_offset) / a.chunk_size;
  int f = 0, y = 0, x = 0;
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  if (valid) {
    f = (int)(p / ((long long)a.Hb * a.Wb));
    int rem = (int)(p - (long long)f * a.Hb * a.Wb);
    y = rem / a.Wb;
    x = rem - y * a.Wb;
    const bool vec = (a.Cs & 3) == 0;
    for (int ky = 0; ky < a.k; ++ky) {
      int ty = y + a.pt - ky;
      if (ty < 0 || ty % a.s != 0) continue;
      int iy = ty / a.s;
      if (iy >= a.Hs) continue;
      for (int kx = 0; kx < a.k; ++kx) {
        int tx = x + a.pl - kx;
        if (tx < 0 || tx % a.s != 0) continue;
        int ix = tx / a.s;
        if (ix >= a.Ws) continue;
        const float* ip = a.small + (((long long)f * a.Hs + iy) * a.Ws + ix) * a.Cs;
        const float* wp = wsm + (ky * a.k + kx) * a.Cs * a.Cb;
        if (vec) {
          for (int c = 0; c < a.Cs; c += 4) {
            float4 v = __ldg(reinterpret_cast<const float4*>(ip + c));
            float vv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int q = 0; q < 4; ++q)
#pragma unroll
              for (int cb = 0; cb < 4; ++cb)
                if (cb < a.Cb) acc[cb] = fmaf(vv[q], wp[(c + q) * a.Cb + cb], acc[cb]);
          }
        } else {
          for (int c = 0; c < a.Cs; ++c) {
            float v = __ldg(ip + c);
#pragma unroll
            for (int cb = 0; cb < 4; ++cb)
              if (cb < a.Cb) acc[cb] = fmaf(v, wp[c * a.Cb + cb], acc[cb]);
          }
        }
      }
    }
  }
  double my_sse = 0.0;
  int my_chunk = 0;
  if (valid) {
    int len = 1;
    if (a.target) {
      my_chunk = (f + a.frame_offset) / a.chunk_size;
      int cbeg = my_chunk * a.chunk_size;
      len = min(a.chunk_size, a.n_total - cbeg);
    }
    const float gsc = a.coef / (float)len;
#pragma unroll
    for (int cb = 0; cb < 4; ++cb) {
      if (cb >= a.Cb) break;
      float v = acc[cb] + (a.bias ? __ldg(a.bias + cb) : 0.f);
      v = 1.f / (1.f + expf(-v));
      long long inchw = (((long long)f * a.Cb + cb) * a.Hb + y) * a.Wb + x;
      a.xhat_ws[inchw] = v;
      if (a.xhat_user) a.xhat_user[inchw] = v;
      if (a.target) {
        float d = v - __ldg(a.target + inchw);
        float m = a.mask ? __ldg(a.mask + inchw) : 1.f;
        my_sse += (double)(d * d * m);
        a.dpre[p * a.Cb + cb] = gsc * d * m * v * (1.f - v);
      }
    }
  }
  if (a.target) {
    // block-level reduction per chunk slot (a block almost always lies inside one chunk)
    int slot = my_chunk - chunk_first;
    bool uniform = __all_sync(0xffffffffu, slot == __shfl_sync(0xffffffffu, slot, 0));
    if (uniform && slot < THIN_SSE_SLOTS) {
      double s = warp_sum_d(my_sse);
      if ((t & 31) == 0 && s != 0.0) atomicAdd(&sse_sm[slot], s);
    } else if (valid && my_sse != 0.0) {
      atomicAdd(a.sse + my_chunk, my_sse);
    }
    __syncthreads();
    if (t < THIN_SSE_SLOTS && sse_sm[t] != 0.0) atomicAdd(a.sse + chunk_first + t, sse_sm[t]);
  }
}

__global__ void sigmoid_bwd_kernel(const float* __restrict__ dxhat, const float* __restrict__ xhat,
                                   float* __restrict__ dpre, int n, int C, int H, int W) {
  bn_pdl_trigger();
  bn_pdl_wait();
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;   // NHWC index
  long long tot = (long long)n * C * H * W;
  if (i >= tot) return;
  int c = (int)(i % C);
  long long pix = i / C;
  int x = (int)(pix % W);
  long long r = pix / W;
  int y = (int)(r % H);
  int f = (int)(r / H);
  long long j = (((long long)f * C + c) * H + y) * W + x;
  float v = xhat[j];
  dpre[i] = dxhat[j] * v * (1.f - v);
}

// ------------------------------------------------------------------------------------------------
// column sums (bias gradients)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) colsum_wide_kernel(const float* __restrict__ x, long long M,
                                                          int C, long long rows_per_block,
                                                          float* __restrict__ out) {
  bn_pdl_trigger();
  bn_pdl_wait();
  __shared__ float red[8][33];
  int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  int c = blockIdx.x * 32 + tx;
  long long r0 = (long long)blockIdx.y * rows_per_block;
  long long r1 = min(M, r0 + rows_per_block);
  float s = 0.f;
  if (c < C)
    for (long long r = r0 + ty; r < r1; r += 8) s += __ldg(x + r * C + c);
  red[ty][tx] = s;
  __syncthreads();
  if (ty == 0) {
#pragma unroll
    for (int i = 1; i < 8; ++i) s += red[i][tx];
    if (c < C) atomicAdd(out + c, s);
  }
}

__global__ void __launch_bounds__(256) colsum_thin_kernel(const float* __restrict__ x,
                                                          long long total, int C,
                                                          float* __restrict__ out) {
  bn_pdl_trigger();
  bn_pdl_wait();
  // C <= 4: flat grid-stride walk, per-thread accumulators per column.  C in {1, 2, 4} with a 16-byte aligned
  // image: float4 loads, element j of a vector always belongs to column j % C (no per-element 64-bit modulo,
  // which cost ~50 instructions per load in the scalar loop: 16 us for 17 MB)
  __shared__ float red[8][4];
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  const long long stride = (long long)gridDim.x * blockDim.x;
  const long long first = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if ((C == 1 || C == 2 || C == 4) && (total & 3) == 0 && ((uintptr_t)x & 15) == 0) {
    const float4* __restrict__ x4 = reinterpret_cast<const float4*>(x);
    const long long n4 = total >> 2;
    long long i = first;
    for (; i + 3 * stride < n4; i += 4 * stride) {
      const float4 a = __ldg(x4 + i), b = __ldg(x4 + i + stride), c = __ldg(x4 + i + 2 * stride),
                   d = __ldg(x4 + i + 3 * stride);
      acc[0] += (a.x + b.x) + (c.x + d.x);
      acc[1] += (a.y + b.y) + (c.y + d.y);
      acc[2] += (a.z + b.z) + (c.z + d.z);
      acc[3] += (a.w + b.w) + (c.w + d.w);
    }
    for (; i < n4; i += stride) {
      const float4 a = __ldg(x4 + i);
      acc[0] += a.x; acc[1] += a.y; acc[2] += a.z; acc[3] += a.w;
    }
    if (C == 1) { acc[0] = (acc[0] + acc[1]) + (acc[2] + acc[3]); acc[1] = acc[2] = acc[3] = 0.f; }
    if (C == 2) { acc[0] += acc[2]; acc[1] += acc[3]; acc[2] = acc[3] = 0.f; }
  } else {
    int c = (int)(first % C);
    const int cstep = (int)(stride % C);
    for (long long i = first; i < total; i += stride) {
      const float v = __ldg(x + i);
#pragma unroll
      for (int q = 0; q < 4; ++q) acc[q] += (q == c) ? v : 0.f;
      c += cstep;
      if (c >= C) c -= C;
    }
  }
#pragma unroll
  for (int q = 0; q < 4; ++q) acc[q] = warp_sum(acc[q]);
  int w = threadIdx.x >> 5;
  if ((threadIdx.x & 31) == 0)
    for (int q = 0; q < 4; ++q) red[w][q] = acc[q];
  __syncthreads();
  if (threadIdx.x < C) {
    float s = 0.f;
    for (int i = 0; i < 8; ++i) s += red[i][threadIdx.x];
    atomicAdd(out + threadIdx.x, s);
  }
}

// ------------------------------------------------------------------------------------------------
// packing
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float round_tf32(float x) {
  uint32_t u;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(x));
  return __uint_as_float(u);
}

// Weight packing, src [cs][cb][tap] (torch layout, tap fastest) -> four GEMM orders.  A one-thread-
// per-element version writes four scattered 4-byte stores per element (a 32-byte sector each); here
// two block bodies stage bricks in shared memory so that every global access is a contiguous run:
//   pack_rows_block : one block per cs: row (cb, tap) -> wft [cs][tap][cb], wd [tap][cs][cb]
//   pack_cols_block : block = 32 cs x 8 cb x all taps -> wf [tap][cb][cs], wdt [cb][tap][cs]
// All conv layers of a plan are packed by ONE launch (pack_all_kernel): after every optimizer step
// the step used to open with 20 small launches (32..512 blocks each, ~100 us in total).
__device__ __forceinline__ void pack_rows_block(const PackJob& jb, int cs, float* prow) {
  const int Cs = jb.Cs, Cb = jb.Cb, kk = jb.kk;   // prow: [cb][kk + 1]
  const int n = Cb * kk;
  const float* s = jb.src + (long long)cs * n;
  for (int i = threadIdx.x; i < n; i += 256) {
    const int cb = i / kk, tap = i - cb * kk;
    prow[cb * (kk + 1) + tap] = s[i];
  }
  __syncthreads();
  for (int i = threadIdx.x; i < n; i += 256) {
    const int tap = i / Cb, cb = i - tap * Cb;
    const float v = prow[cb * (kk + 1) + tap];
    jb.wd[((long long)tap * Cs + cs) * Cb + cb] = v;
    jb.wft[((long long)cs * kk + tap) * Cb + cb] = round_tf32(v);
  }
}

constexpr int PCS = 32, PCB = 8;
__device__ __forceinline__ void pack_cols_block(const PackJob& jb, int bx, int by, float* brick) {
  const int Cs = jb.Cs, Cb = jb.Cb, kk = jb.kk;   // brick: [cs 32][PCB * kk + 1]
  const int cs0 = bx * PCS, cb0 = by * PCB;
  const int ncb = min(PCB, Cb - cb0), ncs = min(PCS, Cs - cs0);
  const int run = ncb * kk, pitch = PCB * kk + 1;
  for (int i = threadIdx.x; i < ncs * run; i += 256) {
    const int c = i / run, r = i - c * run;
    brick[c * pitch + r] = jb.src[((long long)(cs0 + c) * Cb + cb0) * kk + r];    // runs of ncb*kk contiguous floats
  }
  __syncthreads();
  for (int i = threadIdx.x; i < run * PCS; i += 256) {
    const int r = i / PCS, c = i - r * PCS;             // c = cs fastest
    if (c >= ncs) continue;
    const int cbl = r / kk, tap = r - cbl * kk;
    const float v = brick[c * pitch + r];
    jb.wf[((long long)tap * Cb + cb0 + cbl) * Cs + cs0 + c] = v;
    jb.wdt[((long long)(cb0 + cbl) * kk + tap) * Cs + cs0 + c] = round_tf32(v);
  }
}

__global__ void __launch_bounds__(256) pack_all_kernel(const __grid_constant__ PackJobs jobs) {
  bn_pdl_trigger();
  bn_pdl_wait();
  extern __shared__ float pack_smem[];
  const int b = blockIdx.x;
  for (int q = 0; q < jobs.n; ++q) {
    const PackJob& jb = jobs.j[q];
    if (b >= jb.block0 && b < jb.block0 + jb.nrows) {
      pack_rows_block(jb, b - jb.block0, pack_smem);
      return;
    }
    const int c = b - jb.block0 - jb.nrows;
    if (c >= 0 && c < jb.gx * jb.gy) {
      pack_cols_block(jb, c % jb.gx, c / jb.gx, pack_smem);
      return;
    }
  }
}

__global__ void pack_heads_kernel(const float* __restrict__ w0, const float* __restrict__ w1, int L,
                                  int C, int H, int W, float* __restrict__ wcat) {
  bn_pdl_trigger();
  bn_pdl_wait();
  long long F = (long long)C * H * W;
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  long long tot = (w1 ? 2 : 1) * L * F;
  if (i >= tot) return;
  long long inhwc = i % F;
  int hj = (int)(i / F);
  int c = (int)(inhwc % C);
  long long pix = inhwc / C;
  int x = (int)(pix % W);
  int y = (int)(pix / W);
  long long ichw = ((long long)c * H + y) * W + x;
  const float* src = hj < L ? w0 : w1;
  int j = hj < L ? hj : hj - L;
  wcat[i] = src[(long long)j * F + ichw];
}

// ------------------------------------------------------------------------------------------------
// encoder heads (FF / logvar Linear layers, aes.py:118-125, 214-218)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) heads_fwd_kernel(const float* __restrict__ feat,
                                                        const float* __restrict__ wcat,
                                                        const float* __restrict__ b0,
                                                        const float* __restrict__ b1, int n, int L,
                                                        int nheads, int F, float* __restrict__ mu,
                                                        float* __restrict__ logvar) {
  bn_pdl_trigger();
  bn_pdl_wait();
  int warp = (blockIdx.x * 256 + threadIdx.x) >> 5;
  int lane = threadIdx.x & 31;
  int HJ = nheads * L;
  if (warp >= n * HJ) return;
  int f = warp / HJ, hj = warp - f * HJ;
  const float* fp = feat + (long long)f * F;
  const float* wp = wcat + (long long)hj * F;
  float s = 0.f;
  if ((F & 3) == 0) {
    for (int i = lane * 4; i < F; i += 128) {
      float4 a = __ldg(reinterpret_cast<const float4*>(fp + i));
      float4 b = __ldg(reinterpret_cast<const float4*>(wp + i));
      s = fmaf(a.x, b.x, s); s = fmaf(a.y, b.y, s); s = fmaf(a.z, b.z, s); s = fmaf(a.w, b.w, s);
    }
  } else {
    for (int i = lane; i < F; i += 32) s = fmaf(__ldg(fp + i), __ldg(wp + i), s);
  }
  s = warp_sum(s);
  if (lane == 0) {
    if (hj < L) mu[(long long)f * L + hj] = s + b0[hj];
    else logvar[(long long)f * L + hj - L] = s + b1[hj - L];
  }
}

__device__ __forceinline__ float head_grad(const float* dmu, const float* dlv, int f, int hj, int L) {
  if (hj < L) return dmu ? __ldg(dmu + (long long)f * L + hj) : 0.f;
  return dlv ? __ldg(dlv + (long long)f * L + hj - L) : 0.f;
}

__global__ void heads_bwd_data_kernel(const float* __restrict__ feat, const float* __restrict__ wcat,
                                      const float* __restrict__ dmu, const float* __restrict__ dlv,
                                      int n, int L, int HJ, int F, float* __restrict__ dpre) {
  bn_pdl_trigger();
  bn_pdl_wait();
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)n * F) return;
  int f = (int)(i / F);
  int k = (int)(i - (long long)f * F);
  float s = 0.f;
  for (int hj = 0; hj < HJ; ++hj) s = fmaf(head_grad(dmu, dlv, f, hj, L), __ldg(wcat + (long long)hj * F + k), s);
  dpre[i] = s * (feat[i] > 0.f ? 1.f : BN_LEAK);
}

// dW_head[j][i_chw] += sum_f dhead[f][j] * feat[f][i_nhwc].  The frame reduction is split over
// grid.y (one atomicAdd per split) and unrolled into four independent chains: the old one-thread
// 256-long dependent chain was pure latency.
__global__ void __launch_bounds__(256) heads_bwd_w_kernel(const float* __restrict__ feat, const float* __restrict__ dmu,
                                                          const float* __restrict__ dlv, int n, int L, int HJ, int C, int H,
                                                          int W, int fper, float* __restrict__ gw0,
                                                          float* __restrict__ gw1) {
  bn_pdl_trigger();
  bn_pdl_wait();
  long long F = (long long)C * H * W;
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)HJ * F) return;
  int hj = (int)(i / F);
  long long inhwc = i - (long long)hj * F;
  float* g = hj < L ? gw0 : gw1;
  if (!g) return;
  if ((hj < L && !dmu) || (hj >= L && !dlv)) return;
  const int f0 = blockIdx.y * fper, f1 = min(n, f0 + fper);
  float s[4] = {0.f, 0.f, 0.f, 0.f};
  int f = f0;
  for (; f + 4 <= f1; f += 4) {
#pragma unroll
    for (int q = 0; q < 4; ++q)
      s[q] = fmaf(head_grad(dmu, dlv, f + q, hj, L), __ldg(feat + (long long)(f + q) * F + inhwc), s[q]);
  }
  for (; f < f1; ++f) s[0] = fmaf(head_grad(dmu, dlv, f, hj, L), __ldg(feat + (long long)f * F + inhwc), s[0]);
  int c = (int)(inhwc % C);
  long long pix = inhwc / C;
  int x = (int)(pix % W);
  int y = (int)(pix / W);
  int j = hj < L ? hj : hj - L;
  atomicAdd(g + (long long)j * F + ((long long)c * H + y) * W + x, (s[0] + s[1]) + (s[2] + s[3]));
}

// db_head[j] += sum_f dhead[f][j]: one block per output, frames strided over the threads
__global__ void __launch_bounds__(128) heads_bwd_b_kernel(const float* __restrict__ dmu, const float* __restrict__ dlv,
                                                          int n, int L, int HJ, float* __restrict__ gb0,
                                                          float* __restrict__ gb1) {
  bn_pdl_trigger();
  bn_pdl_wait();
  __shared__ float red[4];
  const int hj = blockIdx.x;
  float* g = hj < L ? gb0 : gb1;
  if (!g) return;
  if ((hj < L && !dmu) || (hj >= L && !dlv)) return;
  float s = 0.f;
  for (int f = threadIdx.x; f < n; f += 128) s += head_grad(dmu, dlv, f, hj, L);
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) g[hj < L ? hj : hj - L] += (red[0] + red[1]) + (red[2] + red[3]);
}

// ------------------------------------------------------------------------------------------------
// decoder FF (aes.py:263-266, 453-458): Linear then view(C0,H0,W0); we emit NHWC directly
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ long long nhwc_to_chw(long long inhwc, int C, int H, int W) {
  int c = (int)(inhwc % C);
  long long pix = inhwc / C;
  int x = (int)(pix % W);
  int y = (int)(pix / W);
  return ((long long)c * H + y) * W + x;
}

__global__ void decff_fwd_kernel(const float* __restrict__ z, const float* __restrict__ w,
                                 const float* __restrict__ b, int n, int L, int C, int H, int W,
                                 float* __restrict__ h0) {
  bn_pdl_trigger();
  bn_pdl_wait();
  long long F = (long long)C * H * W;
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)n * F) return;
  int f = (int)(i / F);
  long long ichw = nhwc_to_chw(i - (long long)f * F, C, H, W);
  const float* wp = w + ichw * L;
  const float* zp = z + (long long)f * L;
  float s = __ldg(b + ichw);
  for (int j = 0; j < L; ++j) s = fmaf(__ldg(zp + j), __ldg(wp + j), s);
  h0[i] = s;
}

// dz[f][j] = sum_i dh0[f][i_nhwc] * W[i_chw][j]: one block per frame, thread t owns features t, t+256, ..
// and all L outputs (W rows are contiguous in j), then a block reduction per output.
template <int LMAX>
__global__ void __launch_bounds__(256) decff_bwd_z_kernel(const float* __restrict__ w,
                                                          const float* __restrict__ dh0, int n, int L,
                                                          int C, int H, int W, int j0, float* __restrict__ dz) {
  bn_pdl_trigger();
  bn_pdl_wait();
  __shared__ float red[8][LMAX];
  const int f = blockIdx.x;
  const long long F = (long long)C * H * W;
  float acc[LMAX];
#pragma unroll
  for (int j = 0; j < LMAX; ++j) acc[j] = 0.f;
  for (long long i = threadIdx.x; i < F; i += 256) {
    const float d = __ldg(dh0 + (long long)f * F + i);
    const float* wp = w + nhwc_to_chw(i, C, H, W) * L + j0;
#pragma unroll
    for (int j = 0; j < LMAX; ++j)
      if (j0 + j < L) acc[j] = fmaf(d, __ldg(wp + j), acc[j]);
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
  for (int j = 0; j < LMAX; ++j) {
    const float s = warp_sum(acc[j]);
    if (lane == 0) red[warp][j] = s;
  }
  __syncthreads();
  if (threadIdx.x < LMAX && j0 + threadIdx.x < L) {
    float s = 0.f;
#pragma unroll
    for (int q = 0; q < 8; ++q) s += red[q][threadIdx.x];
    dz[(long long)f * L + j0 + threadIdx.x] = s;
  }
}

// dW[i_chw][j] += sum_f dh0[f][i_nhwc] z[f][j], db[i_chw] += sum_f dh0[f][i_nhwc]; frames split over
// grid.y (atomic accumulate), four independent chains per thread
__global__ void __launch_bounds__(256) decff_bwd_w_kernel(const float* __restrict__ z, const float* __restrict__ dh0, int n,
                                                          int L, int C, int H, int W, int fper, float* __restrict__ gw,
                                                          float* __restrict__ gb) {
  bn_pdl_trigger();
  bn_pdl_wait();
  long long F = (long long)C * H * W;
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;   // over F * (L + 1)
  if (i >= F * (L + 1)) return;
  long long inhwc = i / (L + 1);
  int j = (int)(i - inhwc * (L + 1));
  long long ichw = nhwc_to_chw(inhwc, C, H, W);
  const int f0 = blockIdx.y * fper, f1 = min(n, f0 + fper);
  float s[4] = {0.f, 0.f, 0.f, 0.f};
  int f = f0;
  if (j < L) {
    if (!gw) return;
    for (; f + 4 <= f1; f += 4) {
#pragma unroll
      for (int q = 0; q < 4; ++q)
        s[q] = fmaf(__ldg(dh0 + (long long)(f + q) * F + inhwc), __ldg(z + (long long)(f + q) * L + j), s[q]);
    }
    for (; f < f1; ++f) s[0] = fmaf(__ldg(dh0 + (long long)f * F + inhwc), __ldg(z + (long long)f * L + j), s[0]);
    atomicAdd(gw + ichw * L + j, (s[0] + s[1]) + (s[2] + s[3]));
  } else {
    if (!gb) return;
    for (; f + 4 <= f1; f += 4) {
#pragma unroll
      for (int q = 0; q < 4; ++q) s[q] += __ldg(dh0 + (long long)(f + q) * F + inhwc);
    }
    for (; f < f1; ++f) s[0] += __ldg(dh0 + (long long)f * F + inhwc);
    atomicAdd(gb + ichw, (s[0] + s[1]) + (s[2] + s[3]));
  }
}

}  // namespace

int bn_launch_thin_dgrad(const float* small, const ConvGeom& g, const float* wd, const float* bias,
                         int n, float* xhat_ws, float* xhat_user, const float* target,
                         const float* mask, int chunk_size, int frame_offset, int n_total,
                         float grad_coef, double* sse, float* dpre, cudaStream_t st) {
  if (n <= 0) return 0;
  if (g.Cb > 4) BN_FAIL("thin_dgrad: C_big=%d > 4", g.Cb);
  ThinArgs a;
  a.small = small; a.Hs = g.Hs; a.Ws = g.Ws; a.Cs = g.Cs; a.Hb = g.Hb; a.Wb = g.Wb; a.Cb = g.Cb;
  a.k = g.k; a.s = g.s; a.pt = g.pt; a.pl = g.pl; a.n = n; a.wd = wd; a.bias = bias;
  a.xhat_ws = xhat_ws; a.xhat_user = xhat_user; a.target = target; a.mask = mask;
  a.n_total = n_total > 0 ? n_total : n;
  a.frame_offset = frame_offset;
  a.chunk_size = chunk_size > 0 ? chunk_size : a.n_total;
  a.coef = grad_coef; a.sse = sse; a.dpre = dpre;
  size_t smem = (size_t)g.k * g.k * g.Cs * g.Cb * sizeof(float);
  if (smem > 48 * 1024) {
    BN_CUDA(cudaFuncSetAttribute(thin_dgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  }
  long long npix = (long long)n * g.Hb * g.Wb;
  BN_CUDA(bn_launch(thin_dgrad_kernel, dim3(bn_cdiv(npix, 256)), 256, smem, st, a));
  BN_LAUNCHED();
  return 0;
}

int bn_launch_sigmoid_bwd(const float* dxhat, const float* xhat, float* dpre, int n, int C, int H,
                          int W, cudaStream_t st) {
  long long tot = (long long)n * C * H * W;
  if (tot == 0) return 0;
  BN_CUDA(bn_launch(sigmoid_bwd_kernel, dim3(bn_cdiv(tot, 256)), 256, 0, st, dxhat, xhat, dpre, n, C, H, W));
  BN_LAUNCHED();
  return 0;
}

int bn_launch_colsum(const float* x, long long M, int C, float* out, cudaStream_t st) {
  if (M <= 0 || out == nullptr) return 0;
  if (C <= 4) {
    long long total = M * C;
    int blocks = (int)min((long long)148 * 8, (total + 255) / 256);
    BN_CUDA(bn_launch(colsum_thin_kernel, dim3(blocks), 256, 0, st, x, total, C, out));
  } else {
    int gx = bn_cdiv(C, 32);
    int gy = (int)min((long long)bn_cdiv(148 * 8, gx), (M + 63) / 64);
    if (gy < 1) gy = 1;
    long long rpb = (M + gy - 1) / gy;
    BN_CUDA(bn_launch(colsum_wide_kernel, dim3(gx, gy), 256, 0, st, x, M, C, rpb, out));
  }
  BN_LAUNCHED();
  return 0;
}

int bn_launch_pack_all(PackJobs& jobs, cudaStream_t st) {
  if (jobs.n <= 0) return 0;
  size_t smem = 0;
  int blocks = 0;
  for (int q = 0; q < jobs.n; ++q) {
    PackJob& jb = jobs.j[q];
    const size_t smem_rows = (size_t)jb.Cb * (jb.kk + 1) * sizeof(float);
    const size_t smem_cols = (size_t)PCS * (PCB * jb.kk + 1) * sizeof(float);
    if (smem_rows > 200 * 1024 || smem_cols > 200 * 1024) BN_FAIL("pack: layer too wide (C_big=%d, k*k=%d)", jb.Cb, jb.kk);
    smem = smem_rows > smem ? smem_rows : smem;
    smem = smem_cols > smem ? smem_cols : smem;
    jb.block0 = blocks;
    jb.nrows = jb.Cs;
    jb.gx = bn_cdiv(jb.Cs, PCS);
    jb.gy = bn_cdiv(jb.Cb, PCB);
    blocks += jb.nrows + jb.gx * jb.gy;
  }
  static size_t cfg = 0;
  if (smem > cfg) {
    BN_CUDA(cudaFuncSetAttribute(pack_all_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cfg = smem;
  }
  BN_CUDA(bn_launch(pack_all_kernel, dim3(blocks), 256, smem, st, jobs));
  BN_LAUNCHED();
  return 0;
}

int bn_launch_pack_heads(const float* w0, const float* w1, int L, int C, int H, int W, float* wcat,
                         cudaStream_t st) {
  long long tot = (long long)(w1 ? 2 : 1) * L * C * H * W;
  BN_CUDA(bn_launch(pack_heads_kernel, dim3(bn_cdiv(tot, 256)), 256, 0, st, w0, w1, L, C, H, W, wcat));
  BN_LAUNCHED();
  return 0;
}

int bn_launch_heads_fwd(const float* feat, const float* wcat, const float* b0, const float* b1,
                        int n, int L, int nheads, int F, float* mu, float* logvar, cudaStream_t st) {
  long long warps = (long long)n * nheads * L;
  if (warps == 0) return 0;
  BN_CUDA(bn_launch(heads_fwd_kernel, dim3(bn_cdiv(warps * 32, 256)), 256, 0, st, feat, wcat, b0, b1, n, L, nheads, F, mu, logvar));
  BN_LAUNCHED();
  return 0;
}

int bn_launch_heads_bwd(const float* feat, const float* wcat, const float* dmu, const float* dlogvar,
                        int n, int L, int C, int H, int W, float* dpre_feat, float* gw0, float* gb0,
                        float* gw1, float* gb1, cudaStream_t st) {
  if (n <= 0) return 0;
  int F = C * H * W;
  int HJ = dlogvar ? 2 * L : L;
  BN_CUDA(bn_launch(heads_bwd_data_kernel, dim3(bn_cdiv((long long)n * F, 256)), 256, 0, st, feat, wcat, dmu, dlogvar, n, L, HJ, F, dpre_feat));
  BN_LAUNCHED();
  const int fsplit = n >= 64 ? 8 : 1, fper = bn_cdiv(n, fsplit);
  BN_CUDA(bn_launch(heads_bwd_w_kernel, dim3(bn_cdiv((long long)HJ * F, 256), fsplit), 256, 0, st, feat, dmu, dlogvar, n, L, HJ, C, H, W,
                                                                                   fper, gw0, gw1));
  BN_LAUNCHED();
  BN_CUDA(bn_launch(heads_bwd_b_kernel, dim3(HJ), 128, 0, st, dmu, dlogvar, n, L, HJ, gb0, gb1));
  BN_LAUNCHED();
  return 0;
}

int bn_launch_decff_fwd(const float* z, const float* w, const float* b, int n, int L, int C, int H,
                        int W, float* h0, cudaStream_t st) {
  long long tot = (long long)n * C * H * W;
  if (tot == 0) return 0;
  BN_CUDA(bn_launch(decff_fwd_kernel, dim3(bn_cdiv(tot, 256)), 256, 0, st, z, w, b, n, L, C, H, W, h0));
  BN_LAUNCHED();
  return 0;
}

int bn_launch_decff_bwd(const float* z, const float* w, const float* dh0, int n, int L, int C, int H,
                        int W, float* dz, float* gw, float* gb, cudaStream_t st) {
  if (n <= 0) return 0;
  if (dz) {
    if (L <= 16) {
      BN_CUDA(bn_launch(decff_bwd_z_kernel<16>, dim3(n), 256, 0, st, w, dh0, n, L, C, H, W, 0, dz));
      BN_LAUNCHED();
    } else {
      for (int j0 = 0; j0 < L; j0 += 32) {
        BN_CUDA(bn_launch(decff_bwd_z_kernel<32>, dim3(n), 256, 0, st, w, dh0, n, L, C, H, W, j0, dz));
        BN_LAUNCHED();
      }
    }
  }
  long long F = (long long)C * H * W;
  const int fsplit = n >= 64 ? 8 : 1, fper = bn_cdiv(n, fsplit);
  BN_CUDA(bn_launch(decff_bwd_w_kernel, dim3(bn_cdiv(F * (L + 1), 256), fsplit), 256, 0, st, z, dh0, n, L, C, H, W, fper, gw, gb));
  BN_LAUNCHED();
  return 0;
}
