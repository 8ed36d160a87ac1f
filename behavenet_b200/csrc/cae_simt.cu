// fp32 CUDA-core implicit-GEMM kernels for the convolutional autoencoder hot path.
//
// These are the general-shape kernels: they handle every layer geometry the reference's
// architecture generator can emit (thin first/last layers, non-square frames, any stride/kernel
// with k*k <= 49) and serve as the on-device cross-check for the tcgen05 kernels in cae_tc.cu.
//
//   igemm_fwd_kernel : C[m, co] = act(bias + sum_{tap,ci} img[pix(m) + off(tap), ci] * W[(tap,ci), co])
//                      fprop form  -> Conv2d forward (aes.py:203-212) and ConvTranspose2d backward-data
//                      dgrad form  -> ConvTranspose2d forward (aes.py:463-470) and Conv2d backward-data
//                      Zero padding (ZeroPad2d, aes.py:149-155) and the decoder crop
//                      (F.pad(-pad), aes.py:467-470) are folded into the gather / store indexing.
//   wgrad_kernel     : dW[(tap,cb), cs] = sum_m big[pix(m) + off(tap), cb] * small[m, cs]
#include <string.h>

#include "cae_kernels.cuh"

namespace {

constexpr int FBM = 128, FBN = 64, FBK = 16;

struct FwdArgs {
  ImgView in;
  const float* w;
  const float* bias;
  float* out;
  int Ho, Wo, Co;
  const float* dact;
  const TapClass* classes;
  int gs, os, n, act;
};

__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }

template <bool VEC>
__global__ void __launch_bounds__(256) igemm_fwd_kernel(const FwdArgs a) {
  __shared__ TapClass cls;
  __shared__ __align__(16) float As[FBK][FBM + 4];
  __shared__ __align__(16) float Bs[FBK][FBN + 4];
  const int t = threadIdx.x;
  {
    const int* src = reinterpret_cast<const int*>(a.classes + blockIdx.z);
    int* dst = reinterpret_cast<int*>(&cls);
    for (int i = t; i < (int)(sizeof(TapClass) / 4); i += 256) dst[i] = src[i];
  }
  __syncthreads();
  const int HmWm = cls.Hm * cls.Wm;
  const long long M = (long long)a.n * HmWm;
  const long long m0 = (long long)blockIdx.x * FBM;
  if (m0 >= M) return;
  const int n0 = blockIdx.y * FBN;
  const int Ci = a.in.C;
  const int Co = a.Co;
  const int Ktot = cls.ntaps * Ci;
  const int cpt = Ci / FBK;  // chunks per tap (VEC)
  const int nchunks = VEC ? cls.ntaps * cpt : (Ktot + FBK - 1) / FBK;

  // ---- A-load bookkeeping
  constexpr int NR = VEC ? 2 : 1;
  int ybase[NR], xbase[NR];
  long long foff[NR];
  bool rvalid[NR];
  int arow[NR];
  const int kq = t & 3;              // VEC: which float4 of the 16 channels
  const int kl0 = (t >> 7) * 8;      // scalar: first k of this thread's 8
#pragma unroll
  for (int i = 0; i < NR; ++i) {
    arow[i] = VEC ? (t >> 2) + 64 * i : (t & 127);
    long long m = m0 + arow[i];
    rvalid[i] = m < M;
    long long mm = rvalid[i] ? m : 0;
    int f = (int)(mm / HmWm);
    int rem = (int)(mm - (long long)f * HmWm);
    int ym = rem / cls.Wm;
    int xm = rem - ym * cls.Wm;
    ybase[i] = ym * a.gs;
    xbase[i] = xm * a.gs;
    foff[i] = (long long)f * a.in.sn;
  }
  // ---- B-load bookkeeping
  const int bkl = t >> 4;       // 0..15
  const int bcol = n0 + (t & 15) * 4;
  const bool vecB = (Co & 3) == 0;

  float4 ra4[2];
  float ra[8];
  float4 rb;

  auto load_chunk = [&](int c) {
    int wrow;
    bool bvalid;
    if (VEC) {
      int tap = c / cpt;
      int c0 = (c - tap * cpt) * FBK;
      int dy = cls.dy[tap], dx = cls.dx[tap];
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        int y = ybase[i] + dy, x = xbase[i] + dx;
        bool ok = rvalid[i] && (unsigned)y < (unsigned)a.in.H && (unsigned)x < (unsigned)a.in.W;
        ra4[i] = ok ? ldg4(a.in.p + foff[i] + (long long)y * a.in.sy + (long long)x * a.in.sx + c0 + kq * 4)
                    : make_float4(0.f, 0.f, 0.f, 0.f);
      }
      wrow = cls.wt[tap] * Ci + c0 + bkl;
      bvalid = true;
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        int kg = c * FBK + kl0 + j;
        float v = 0.f;
        if (kg < Ktot && rvalid[0]) {
          int tap = kg / Ci;
          int ci = kg - tap * Ci;
          int y = ybase[0] + cls.dy[tap], x = xbase[0] + cls.dx[tap];
          if ((unsigned)y < (unsigned)a.in.H && (unsigned)x < (unsigned)a.in.W)
            v = __ldg(a.in.p + foff[0] + (long long)y * a.in.sy + (long long)x * a.in.sx + (long long)ci * a.in.sc);
        }
        ra[j] = v;
      }
      int kg = c * FBK + bkl;
      bvalid = kg < Ktot;
      int tap = bvalid ? kg / Ci : 0;
      int ci = kg - tap * Ci;
      wrow = cls.wt[tap] * Ci + ci;
    }
    rb = make_float4(0.f, 0.f, 0.f, 0.f);
    if (bvalid) {
      const float* wp = a.w + (long long)wrow * Co + bcol;
      if (vecB && bcol + 3 < Co) {
        rb = ldg4(wp);
      } else {
        if (bcol + 0 < Co) rb.x = __ldg(wp + 0);
        if (bcol + 1 < Co) rb.y = __ldg(wp + 1);
        if (bcol + 2 < Co) rb.z = __ldg(wp + 2);
        if (bcol + 3 < Co) rb.w = __ldg(wp + 3);
      }
    }
  };
  auto store_chunk = [&]() {
    if (VEC) {
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        As[kq * 4 + 0][arow[i]] = ra4[i].x;
        As[kq * 4 + 1][arow[i]] = ra4[i].y;
        As[kq * 4 + 2][arow[i]] = ra4[i].z;
        As[kq * 4 + 3][arow[i]] = ra4[i].w;
      }
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) As[kl0 + j][arow[0]] = ra[j];
    }
    *reinterpret_cast<float4*>(&Bs[bkl][(t & 15) * 4]) = rb;
  };

  const int tx = t & 15, ty = t >> 4;
  float acc[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  load_chunk(0);
  for (int c = 0; c < nchunks; ++c) {
    store_chunk();
    __syncthreads();
    if (c + 1 < nchunks) load_chunk(c + 1);
#pragma unroll
    for (int k = 0; k < FBK; ++k) {
      float4 a0 = *reinterpret_cast<const float4*>(&As[k][ty * 8]);
      float4 a1 = *reinterpret_cast<const float4*>(&As[k][ty * 8 + 4]);
      float4 b = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
      float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      float bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }

  // ---- epilogue
  const int col0 = n0 + tx * 4;
  float bv[4] = {0.f, 0.f, 0.f, 0.f};
  if (a.bias) {
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if (col0 + j < Co) bv[j] = __ldg(a.bias + col0 + j);
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    long long m = m0 + ty * 8 + i;
    if (m >= M) continue;
    int f = (int)(m / HmWm);
    int rem = (int)(m - (long long)f * HmWm);
    int ym = rem / cls.Wm;
    int xm = rem - ym * cls.Wm;
    int oy = cls.oy0 + a.os * ym, ox = cls.ox0 + a.os * xm;
    long long idx = (((long long)f * a.Ho + oy) * a.Wo + ox) * Co + col0;
    float v[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float x = acc[i][j] + bv[j];
      if (a.act == BN_ACT_LEAKY) x = x > 0.f ? x : BN_LEAK * x;
      else if (a.act == BN_ACT_SIGMOID) x = 1.f / (1.f + expf(-x));
      v[j] = x;
    }
    if (vecB && col0 + 3 < Co) {
      if (a.dact) {
        float4 d = ldg4(a.dact + idx);
        v[0] *= d.x > 0.f ? 1.f : BN_LEAK;
        v[1] *= d.y > 0.f ? 1.f : BN_LEAK;
        v[2] *= d.z > 0.f ? 1.f : BN_LEAK;
        v[3] *= d.w > 0.f ? 1.f : BN_LEAK;
      }
      *reinterpret_cast<float4*>(a.out + idx) = make_float4(v[0], v[1], v[2], v[3]);
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        if (col0 + j < Co) {
          float x = v[j];
          if (a.dact) x *= __ldg(a.dact + idx + j) > 0.f ? 1.f : BN_LEAK;
          a.out[idx + j] = x;
        }
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// weight gradient
// ------------------------------------------------------------------------------------------------
constexpr int WBK = 128, WBN = 64, WBM = 16;

struct WgArgs {
  ImgView big;
  const float* small;
  int Hs, Ws, Cs;
  const TapClass* cls;
  int gs, n, Ktot;
  long long rows_per_split;
  float* partial;
};

template <bool VEC>
__global__ void __launch_bounds__(256) wgrad_kernel(const WgArgs a) {
  __shared__ TapClass cls;
  __shared__ __align__(16) float As[WBM][WBK + 4];
  __shared__ __align__(16) float Ss[WBM][WBN + 4];
  const int t = threadIdx.x;
  {
    const int* src = reinterpret_cast<const int*>(a.cls);
    int* dst = reinterpret_cast<int*>(&cls);
    for (int i = t; i < (int)(sizeof(TapClass) / 4); i += 256) dst[i] = src[i];
  }
  __syncthreads();
  const int HsWs = a.Hs * a.Ws;
  const long long M = (long long)a.n * HsWs;
  const long long mbeg = (long long)blockIdx.z * a.rows_per_split;
  const long long mend = min(M, mbeg + a.rows_per_split);
  const int kk0 = blockIdx.x * WBK;
  const int cs0 = blockIdx.y * WBN;
  const int Cb = a.big.C;
  const int Cs = a.Cs;

  // A-load: fixed (tap, cb) per thread
  const int akk = VEC ? kk0 + (t & 31) * 4 : kk0 + (t & 127);
  const bool kvalid = akk < a.Ktot;
  int tap = kvalid ? akk / Cb : 0;
  const int cb = akk - tap * Cb;
  const int dy = cls.dy[tap], dx = cls.dx[tap];
  const int aml0 = VEC ? (t >> 5) : (t >> 7);   // first reduction row of this thread
  // S-load
  const int sml = t >> 4;
  const int scol = cs0 + (t & 15) * 4;
  const bool vecS = (Cs & 3) == 0;

  float4 ra4[2];
  float ra[8];
  float4 rs;

  auto gather = [&](long long m) -> const float* {
    int f = (int)(m / HsWs);
    int rem = (int)(m - (long long)f * HsWs);
    int ym = rem / a.Ws;
    int xm = rem - ym * a.Ws;
    int y = ym * a.gs + dy, x = xm * a.gs + dx;
    if ((unsigned)y >= (unsigned)a.big.H || (unsigned)x >= (unsigned)a.big.W) return nullptr;
    return a.big.p + (long long)f * a.big.sn + (long long)y * a.big.sy + (long long)x * a.big.sx + (long long)cb * a.big.sc;
  };
  auto load_chunk = [&](long long mc) {
    if (VEC) {
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        long long m = mc + aml0 + 8 * i;
        const float* p = (kvalid && m < mend) ? gather(m) : nullptr;
        ra4[i] = p ? ldg4(p) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
    } else {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        long long m = mc + aml0 + 2 * i;
        const float* p = (kvalid && m < mend) ? gather(m) : nullptr;
        ra[i] = p ? __ldg(p) : 0.f;
      }
    }
    rs = make_float4(0.f, 0.f, 0.f, 0.f);
    long long m = mc + sml;
    if (m < mend) {
      const float* sp = a.small + m * Cs + scol;
      if (vecS && scol + 3 < Cs) {
        rs = ldg4(sp);
      } else {
        if (scol + 0 < Cs) rs.x = __ldg(sp + 0);
        if (scol + 1 < Cs) rs.y = __ldg(sp + 1);
        if (scol + 2 < Cs) rs.z = __ldg(sp + 2);
        if (scol + 3 < Cs) rs.w = __ldg(sp + 3);
      }
    }
  };
  auto store_chunk = [&]() {
    if (VEC) {
#pragma unroll
      for (int i = 0; i < 2; ++i) *reinterpret_cast<float4*>(&As[aml0 + 8 * i][(t & 31) * 4]) = ra4[i];
    } else {
#pragma unroll
      for (int i = 0; i < 8; ++i) As[aml0 + 2 * i][t & 127] = ra[i];
    }
    *reinterpret_cast<float4*>(&Ss[sml][(t & 15) * 4]) = rs;
  };

  const int tx = t & 15, ty = t >> 4;
  float acc[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  if (mbeg < mend) {
    load_chunk(mbeg);
    for (long long mc = mbeg; mc < mend; mc += WBM) {
      store_chunk();
      __syncthreads();
      if (mc + WBM < mend) load_chunk(mc + WBM);
#pragma unroll
      for (int k = 0; k < WBM; ++k) {
        float4 a0 = *reinterpret_cast<const float4*>(&As[k][ty * 8]);
        float4 a1 = *reinterpret_cast<const float4*>(&As[k][ty * 8 + 4]);
        float4 b = *reinterpret_cast<const float4*>(&Ss[k][tx * 4]);
        float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
        float bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
      }
      __syncthreads();
    }
  }
  float* pz = a.partial + (long long)blockIdx.z * a.Ktot * Cs;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    int kk = kk0 + ty * 8 + i;
    if (kk >= a.Ktot) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int c = cs0 + tx * 4 + j;
      if (c < Cs) pz[(long long)kk * Cs + c] = acc[i][j];
    }
  }
}

// grad[cs][cb][wt[tap]] += sum_z partial[z][(tap, cb)][cs].  The partial slices are cs-fastest, the
// torch gradient is tap-fastest: a block owns 32 cs x 32 consecutive (cb, tap) output positions, its
// warps read the matching partial rows (128 contiguous bytes of cs each, 8 loads in flight per lane)
// and the brick is transposed through shared memory so the read-modify-write runs are 128 bytes too.
// Few-output layers (the thin first / last layer has 800 outputs but ~300 slices) spread the slices
// over grid.z and finish with atomicAdd; everything else is a deterministic +=.
struct ReduceJob {
  const float* partial;
  float* grad;
  const TapClass* cls;
  int splits, Ktot, Cs, Cb, KK;
  int gx, gy, gz, block0;
  int kind;                   // 0: weight-gradient slices -> torch layout; 1: column sums of a [splits][Cs] table
};
struct ReduceJobs {
  int n;
  ReduceJob j[32];            // 2 * BN_MAX_LAYERS weight gradients + as many bias gradients
};

// out[c] += sum_r part[r][c]: block (bx, bz) owns 32 columns and the rows bz, bz + gz, .. (8 warps interleaved)
__device__ __forceinline__ void colsum_reduce_block(const ReduceJob& jb, int bx, int bz, float (*tile)[33]) {
  const float* __restrict__ part = jb.partial;
  const int rows = jb.splits, C = jb.Cs, gz = jb.gz;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int c = bx * 32 + tx;
  float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
  if (c < C) {
    const int step = 8 * gz;
    int r = bz * 8 + ty;
    for (; r + 3 * step < rows; r += 4 * step) {
      s0 += __ldg(part + (long long)r * C + c);
      s1 += __ldg(part + (long long)(r + step) * C + c);
      s2 += __ldg(part + (long long)(r + 2 * step) * C + c);
      s3 += __ldg(part + (long long)(r + 3 * step) * C + c);
    }
    for (; r < rows; r += step) s0 += __ldg(part + (long long)r * C + c);
  }
  tile[ty][tx] = (s0 + s1) + (s2 + s3);
  __syncthreads();
  if (ty == 0 && c < C) {
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += tile[i][tx];
    if (gz > 1) atomicAdd(jb.grad + c, s);
    else jb.grad[c] += s;
  }
}

__device__ __forceinline__ void wgrad_reduce_block(const ReduceJob& jb, int bx, int by, int bz, float (*tile)[33],
                                                   unsigned char* inv) {
  const float* __restrict__ partial = jb.partial;
  const int splits = jb.splits, Ktot = jb.Ktot, Cs = jb.Cs, Cb = jb.Cb, KK = jb.KK, gz = jb.gz;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid < KK) inv[jb.cls->wt[tid]] = (unsigned char)tid;
  __syncthreads();
  const int cs = bx * 32 + lane;
  const int o0 = by * 32;
  const int nout = Cb * KK;
  const long long tot = (long long)Ktot * Cs;
  const int zper = (splits + gz - 1) / gz;
  const int z0 = bz * zper, z1 = min(splits, z0 + zper);
  float s[4] = {0.f, 0.f, 0.f, 0.f};
  const float* p[4];
  bool ok[4];
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const int o = o0 + warp * 4 + r;
    ok[r] = o < nout && cs < Cs;
    const int cb = ok[r] ? o / KK : 0;
    const int tap = ok[r] ? inv[o - cb * KK] : 0;
    p[r] = partial + ((long long)tap * Cb + cb) * Cs + (cs < Cs ? cs : 0);
  }
  int z = z0;
  for (; z + 2 <= z1; z += 2) {
    float v[8];
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      v[2 * r] = ok[r] ? __ldg(p[r] + (long long)z * tot) : 0.f;
      v[2 * r + 1] = ok[r] ? __ldg(p[r] + (long long)(z + 1) * tot) : 0.f;
    }
#pragma unroll
    for (int r = 0; r < 4; ++r) s[r] += v[2 * r] + v[2 * r + 1];
  }
  for (; z < z1; ++z)
#pragma unroll
    for (int r = 0; r < 4; ++r) s[r] += ok[r] ? __ldg(p[r] + (long long)z * tot) : 0.f;
#pragma unroll
  for (int r = 0; r < 4; ++r) tile[lane][warp * 4 + r] = s[r];
  __syncthreads();
  // 8 warps x 4 cs rows; lane = output position
  const int o = o0 + lane;
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const int c = bx * 32 + warp * 4 + r;
    if (c < Cs && o < nout) {
      float* g = jb.grad + (long long)c * nout + o;
      if (gz > 1) atomicAdd(g, tile[warp * 4 + r][lane]);
      else *g += tile[warp * 4 + r][lane];
    }
  }
}

// every queued reduction of one backward call in ONE launch (block ranges per job)
__global__ void __launch_bounds__(256) wgrad_reduce_kernel(const __grid_constant__ ReduceJobs jobs) {
  bn_pdl_trigger();
  bn_pdl_wait();
  __shared__ float tile[32][33];
  __shared__ unsigned char inv[BN_MAX_TAPS];
  const int b = blockIdx.x;
  for (int q = 0; q < jobs.n; ++q) {
    const ReduceJob& jb = jobs.j[q];
    const int l = b - jb.block0;
    if (l >= 0 && l < jb.gx * jb.gy * jb.gz) {
      const int bx = l % jb.gx, by = (l / jb.gx) % jb.gy, bz = l / (jb.gx * jb.gy);
      if (jb.kind == 1) colsum_reduce_block(jb, bx, bz, tile);
      else wgrad_reduce_block(jb, bx, by, bz, tile, inv);
      return;
    }
  }
}

thread_local ReduceJobs t_reduce_jobs;
thread_local bool t_reduce_defer = false;

int launch_reduce_jobs(ReduceJobs& jobs, cudaStream_t st) {
  if (jobs.n == 0) return 0;
  int blocks = 0;
  for (int q = 0; q < jobs.n; ++q) {
    jobs.j[q].block0 = blocks;
    blocks += jobs.j[q].gx * jobs.j[q].gy * jobs.j[q].gz;
  }
  BN_CUDA(bn_launch(wgrad_reduce_kernel, dim3(blocks), 256, 0, st, jobs));
  BN_LAUNCHED();
  jobs.n = 0;
  return 0;
}

}  // namespace

int bn_launch_igemm(const ImgView& in, const float* w, const float* bias, float* out, int Ho, int Wo,
                    int Co, const float* dact, const TapClass* d_classes, int nclasses, int maxM,
                    int gs, int os, int n, int act, cudaStream_t st) {
  if (n <= 0 || maxM <= 0) return 0;
  FwdArgs a;
  a.in = in; a.w = w; a.bias = bias; a.out = out; a.Ho = Ho; a.Wo = Wo; a.Co = Co; a.dact = dact;
  a.classes = d_classes; a.gs = gs; a.os = os; a.n = n; a.act = act;
  dim3 grid(bn_cdiv((long long)n * maxM, FBM), bn_cdiv(Co, FBN), nclasses);
  bool vec = (in.C % FBK == 0) && in.sc == 1 && (in.sx % 4 == 0) && (in.sy % 4 == 0) && (in.sn % 4 == 0) &&
             ((uintptr_t)in.p % 16 == 0);
  if (vec) igemm_fwd_kernel<true><<<grid, 256, 0, st>>>(a);
  else igemm_fwd_kernel<false><<<grid, 256, 0, st>>>(a);
  BN_LAUNCHED();
  return 0;
}

static int wgrad_splits(const ConvGeom& g, int n) {
  long long M = (long long)n * g.Hs * g.Ws;
  int Ktot = g.k * g.k * g.Cb;
  long long tiles = (long long)bn_cdiv(Ktot, WBK) * bn_cdiv(g.Cs, WBN);
  long long want = (4 * 148 + tiles - 1) / tiles;
  long long maxs = (M + 255) / 256;
  if (want > maxs) want = maxs;
  if (want < 1) want = 1;
  if (want > 256) want = 256;
  return (int)want;
}

size_t bn_wgrad_partial_floats(const ConvGeom& g, int n) {
  // room for the CUDA-core split count and for the tensor-core kernel's (up to ~2 waves of CTAs)
  long long M = (long long)n * g.Hs * g.Ws;
  long long Ktot = (long long)g.k * g.k * g.Cb;
  long long tiles_tc = ((Ktot + 127) / 128) * ((g.Cs + 255) / 256);
  long long s_tc = (2 * 148 + tiles_tc - 1) / tiles_tc;
  long long maxs = (M + 127) / 128;
  if (s_tc > maxs) s_tc = maxs;
  if (s_tc < 1) s_tc = 1;
  long long s = wgrad_splits(g, n);
  if (s_tc > s) s = s_tc;
  return (size_t)s * Ktot * g.Cs;
}

int bn_launch_wgrad_reduce(const float* partial, int splits, int Ktot, int Cs, int Cb, int KK,
                           const TapClass* cls, float* grad, cudaStream_t st) {
  const int gx = bn_cdiv(Cs, 32), gy = bn_cdiv((long long)Cb * KK, 32);
  int gz = 2 * 148 / (gx * gy);
  if (gz > splits / 4) gz = splits / 4;
  if (gz < 1) gz = 1;
  ReduceJob jb;
  jb.partial = partial; jb.grad = grad; jb.cls = cls;
  jb.splits = splits; jb.Ktot = Ktot; jb.Cs = Cs; jb.Cb = Cb; jb.KK = KK;
  jb.gx = gx; jb.gy = gy; jb.gz = gz; jb.block0 = 0; jb.kind = 0;
  if (t_reduce_defer && t_reduce_jobs.n < (int)(sizeof(t_reduce_jobs.j) / sizeof(t_reduce_jobs.j[0]))) {
    t_reduce_jobs.j[t_reduce_jobs.n++] = jb;       // launched by bn_wgrad_reduce_flush
    return 0;
  }
  ReduceJobs one;
  one.n = 1;
  one.j[0] = jb;
  return launch_reduce_jobs(one, st);
}

// Backward entry points give every layer its own partial-sum region and queue the reductions
// (begin), then run them all in one launch (flush): ten small grids become two large ones.
void bn_wgrad_reduce_defer_begin() {
  t_reduce_jobs.n = 0;
  t_reduce_defer = true;
}

bool bn_reduce_deferring() {
  return t_reduce_defer && t_reduce_jobs.n < (int)(sizeof(t_reduce_jobs.j) / sizeof(t_reduce_jobs.j[0]));
}

int bn_colsum_reduce_defer(const float* part, int rows, int C, float* out) {
  if (!bn_reduce_deferring()) BN_FAIL("bn_colsum_reduce_defer: no batched reduction is open");
  ReduceJob jb;
  memset(&jb, 0, sizeof(jb));
  jb.partial = part; jb.grad = out; jb.cls = nullptr;
  jb.splits = rows; jb.Cs = C; jb.kind = 1;
  jb.gx = bn_cdiv(C, 32); jb.gy = 1;
  jb.gz = rows >= 1024 ? 16 : (rows >= 128 ? 4 : 1);
  t_reduce_jobs.j[t_reduce_jobs.n++] = jb;
  return 0;
}

int bn_wgrad_reduce_flush(cudaStream_t st) {
  t_reduce_defer = false;
  return launch_reduce_jobs(t_reduce_jobs, st);
}

int bn_launch_wgrad(const ImgView& big, const float* small, const ConvGeom& g, int n, float* partial,
                    size_t partial_floats, float* grad, cudaStream_t st) {
  if (n <= 0 || grad == nullptr) return 0;
  int splits = wgrad_splits(g, n);
  int Ktot = g.k * g.k * g.Cb;
  if ((size_t)splits * Ktot * g.Cs > partial_floats) BN_FAIL("wgrad partial buffer too small");
  long long M = (long long)n * g.Hs * g.Ws;
  long long rps = (M + splits - 1) / splits;
  rps = (rps + WBM - 1) / WBM * WBM;
  WgArgs a;
  a.big = big; a.small = small; a.Hs = g.Hs; a.Ws = g.Ws; a.Cs = g.Cs; a.cls = g.d_fprop; a.gs = g.s;
  a.n = n; a.Ktot = Ktot; a.rows_per_split = rps; a.partial = partial;
  dim3 grid(bn_cdiv(Ktot, WBK), bn_cdiv(g.Cs, WBN), splits);
  bool vec = (big.C % 4 == 0) && big.sc == 1 && (big.sx % 4 == 0) && (big.sy % 4 == 0) && (big.sn % 4 == 0) &&
             ((uintptr_t)big.p % 16 == 0);
  if (vec) wgrad_kernel<true><<<grid, 256, 0, st>>>(a);
  else wgrad_kernel<false><<<grid, 256, 0, st>>>(a);
  BN_LAUNCHED();
  return bn_launch_wgrad_reduce(partial, splits, Ktot, g.Cs, g.Cb, g.k * g.k, g.d_fprop, grad, st);
}
