// Hot path 2: ARHMM E-step (forward-backward), log-likelihood and Viterbi on sm_100a.
//
// What the reference reaches through ssm (fitting/arhmm_grid_search.py:170-204, fitting/eval.py:167):
//   hmm.fit(method='em') E-step  -> bn_arhmm_estep   (gamma, sum_t xi_t, log normaliser per trial)
//   hmm.log_likelihood           -> bn_arhmm_estep with NULL Ez/Ezz (forward pass only)
//   hmm.most_likely_states       -> bn_arhmm_viterbi
//
// Two kernels per call, by design (hard parts 1 and 7):
//   1. emission kernel: embarrassingly parallel over timesteps.  The AR Gaussian log-density is
//      folded into one whitened affine map per state, y = W_k [x_t, x_{t-1..t-L}, 1] with
//      W_k = chol(Sigma_k)^-1 [I, -A_k, -b_k] (built on the host in fp64), ll = c_k - |y|^2 / 2.
//      Each thread owns TS timesteps, W_k^T is broadcast from shared memory as float4, the x tile
//      is staged once per block with coalesced loads.  Output: per-step max m_t and scaled
//      likelihoods exp(ll - m_t) in (0, 1]  (fp32) -- or raw fp64 ll for Viterbi.
//   2. scan kernel: KP lanes per trial (KP = K rounded up to a power of two, so a warp carries
//      32/KP trials); lane k keeps column/row k of the transition matrix in registers; the
//      matrix-vector step is KP shuffles + FMAs, the normaliser a log2(KP)-step shuffle
//      reduction.  Messages are per-step NORMALISED probabilities in fp32 (fp32 log-space
//      messages miss the 1e-5 posterior tolerance, scaled ones meet it); the log normaliser is
//      accumulated in fp64.  alpha_hat is written into the Ez buffer by the forward sweep and
//      overwritten in place with gamma by the backward sweep (same thread, same address).
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <vector>

#include "../../include/behavenet_b200.h"
#include "arhmm_common.cuh"

namespace {

constexpr double LN2PI = 1.8378770664093453;

// ------------------------------------------------------------------------------------------------
// emission kernel
// ------------------------------------------------------------------------------------------------
template <typename real>
struct EmitArgs {
  const unsigned char* blob;
  const float* x;
  const long long* offsets;
  int K, D, lags, J;
  float* Bsc;      // (total_T, K) scaled likelihoods   [real == float]
  float* mx;       // (total_T) per-step max            [real == float]
  double* ll;      // (total_T, K) raw log-likelihoods  [real == double]
};

template <typename real, int DP, int TS>
__global__ void __launch_bounds__(128) emission_kernel(const EmitArgs<real> a) {
  constexpr int TILE = 128 * TS;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const BlobHeader* hd = reinterpret_cast<const BlobHeader*>(a.blob);
  const int K = a.K, D = a.D, L = a.lags, J = a.J;
  const int XS = D | 1;                         // odd row stride: conflict-free column reads
  real* Wsm = reinterpret_cast<real*>(smem_raw);                    // K * J * DP
  real* csm = Wsm + (size_t)K * J * DP;                             // K + 1 (+pad)
  float* xs = reinterpret_cast<float*>(csm + ((K + 1 + 3) & ~3));   // (TILE + L) * XS
  float* outs = xs + (size_t)(TILE + L) * XS;                       // TILE * (K + 1)   [float path]

  const int trial = blockIdx.x;
  const long long beg = a.offsets[trial], end = a.offsets[trial + 1];
  const int T = (int)(end - beg);
  if ((int)blockIdx.y * TILE >= T) return;
  const int tid = threadIdx.x;
  {
    // the whitened parameters are staged ONCE per block and reused for all of its time tiles
    const real* Wg = reinterpret_cast<const real*>(a.blob + (sizeof(real) == 4 ? hd->off_W_f : hd->off_W_d));
    const real* cg = reinterpret_cast<const real*>(a.blob + (sizeof(real) == 4 ? hd->off_c_f : hd->off_c_d));
    if ((K * J * DP * (int)sizeof(real)) % 16 == 0) {
      const int4* src = reinterpret_cast<const int4*>(Wg);
      int4* dst = reinterpret_cast<int4*>(Wsm);
      for (int i = tid; i < K * J * DP * (int)sizeof(real) / 16; i += 128) dst[i] = __ldg(src + i);
    } else {
      for (int i = tid; i < K * J * DP; i += 128) Wsm[i] = Wg[i];
    }
    for (int i = tid; i < K + 1; i += 128) csm[i] = cg[i];
  }
  for (int t0 = blockIdx.y * TILE; t0 < T; t0 += gridDim.y * TILE) {
  __syncthreads();      // previous tile's x / out staging is no longer read
  {
    // x rows t0-L .. t0+TILE-1 of this trial (rows before the trial start are never used)
    const int nrows = min(TILE, T - t0) + L;
    const float* xg = a.x + (beg + t0 - L) * D;
    for (int i = tid; i < nrows * D; i += 128) {
      int r = i / D, c = i - r * D;
      xs[r * XS + c] = (t0 - L + r >= 0) ? __ldg(xg + i) : 0.f;
    }
  }
  __syncthreads();

  // Register tile: KG states x TS timesteps x DP whitened residual components per thread.  Every
  // W row (DP values, one broadcast LDS.128 per 4) is reused for TS timesteps and every psi value
  // (one conflict-free LDS) for KG states, so the loop is FMA-bound instead of LDS-bound.
  constexpr int REGW = (int)(sizeof(real) / 4);
  constexpr int KG0 = 96 / (TS * DP * REGW);
  constexpr int KG = KG0 < 1 ? 1 : (KG0 > 4 ? 4 : KG0);
  constexpr int VEC = 16 / (int)sizeof(real);          // elements per 16-byte shared load
  struct alignas(16) Vec { real v[VEC]; };
  real vmax[TS];
#pragma unroll
  for (int s = 0; s < TS; ++s) vmax[s] = (real)-INFINITY;
  for (int kg = 0; kg < K; kg += KG) {
    real acc[KG][TS][DP];
#pragma unroll
    for (int g = 0; g < KG; ++g)
#pragma unroll
      for (int s = 0; s < TS; ++s)
#pragma unroll
        for (int i = 0; i < DP; ++i) acc[g][s][i] = 0;
    const real* Wg[KG];
#pragma unroll
    for (int g = 0; g < KG; ++g) Wg[g] = Wsm + (size_t)min(kg + g, K - 1) * J * DP;
    int l = 0, d = 0;
    for (int j = 0; j < J - 1; ++j) {
      real p[TS];
#pragma unroll
      for (int s = 0; s < TS; ++s) p[s] = (real)xs[(size_t)(tid + 128 * s + L - l) * XS + d];
#pragma unroll
      for (int g = 0; g < KG; ++g) {
        const Vec* wr = reinterpret_cast<const Vec*>(Wg[g] + (size_t)j * DP);
#pragma unroll
        for (int i4 = 0; i4 < DP / VEC; ++i4) {
          const Vec w = wr[i4];
#pragma unroll
          for (int e = 0; e < VEC; ++e)
#pragma unroll
            for (int s = 0; s < TS; ++s) acc[g][s][i4 * VEC + e] = fma(p[s], w.v[e], acc[g][s][i4 * VEC + e]);
        }
      }
      if (++d == D) { d = 0; ++l; }
    }
#pragma unroll
    for (int g = 0; g < KG; ++g) {
      const int k = kg + g;
      const real* wb = Wg[g] + (size_t)(J - 1) * DP;
#pragma unroll
      for (int s = 0; s < TS; ++s) {
        real q = 0;
#pragma unroll
        for (int i = 0; i < DP; ++i) {
          const real y = acc[g][s][i] + wb[i];
          q = fma(y, y, q);
        }
        const real v = csm[min(k, K - 1)] - (real)0.5 * q;
        const int tl = tid + 128 * s;
        const int t = t0 + tl;
        if (k < K && t < T && t >= L) {
          if (sizeof(real) == 4) {
            outs[tl * (K + 1) + k] = (float)v;
            vmax[s] = v > vmax[s] ? v : vmax[s];
          } else {
            a.ll[(beg + t) * K + k] = (double)v;
          }
        }
      }
    }
  }
#pragma unroll
  for (int s = 0; s < TS; ++s) {
    const int tl = tid + 128 * s;
    const int t = t0 + tl;
    if (t >= T) continue;
    if (t < L) {
      // initial segment: N(0, I) for every state (ssm mu_init = 0, Sigma_init = I)
      const float* xr = xs + (size_t)(tl + L) * XS;
      real q = 0;
      for (int dd = 0; dd < D; ++dd) q += (real)xr[dd] * (real)xr[dd];
      const real v = csm[K] - (real)0.5 * q;
      if (sizeof(real) == 4) {
        for (int k = 0; k < K; ++k) outs[tl * (K + 1) + k] = 1.f;
        a.mx[beg + t] = (float)v;
      } else {
        for (int k = 0; k < K; ++k) a.ll[(beg + t) * K + k] = (double)v;
      }
    } else if (sizeof(real) == 4) {
      for (int k = 0; k < K; ++k) outs[tl * (K + 1) + k] = expf(outs[tl * (K + 1) + k] - (float)vmax[s]);
      a.mx[beg + t] = (float)vmax[s];
    }
  }
  if (sizeof(real) == 4) {
    __syncthreads();
    const int nrows = min(TILE, T - t0);
    float* og = a.Bsc + (beg + t0) * K;
    for (int i = tid; i < nrows * K; i += 128) {
      int r = i / K, c = i - r * K;
      og[i] = outs[r * (K + 1) + c];
    }
  }
  }   // time-tile loop
}

// ------------------------------------------------------------------------------------------------
// forward-backward scan
// ------------------------------------------------------------------------------------------------
template <int KP>
__device__ __forceinline__ float group_sum(float v) {
#pragma unroll
  for (int o = KP / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o, KP);
  return v;
}

struct ScanArgs {
  const unsigned char* blob;
  const float* Bsc;
  const float* mx;
  const long long* offsets;
  int n_trials, K;
  float* Ez;      // nullable
  float* Ezz;     // nullable
  double* logZ;   // nullable
  float* cinv;    // (total_T) workspace: 1 / c_t of the forward sweep, reused by the backward sweep
};

template <int KP>
__global__ void __launch_bounds__(128) scan_kernel(const ScanArgs a) {
  constexpr int GPW = 32 / KP;                     // trials per warp
  const BlobHeader* hd = reinterpret_cast<const BlobHeader*>(a.blob);
  const float* Pg = reinterpret_cast<const float*>(a.blob + hd->off_P_f);     // KP x KP, zero padded
  const float* pi0g = reinterpret_cast<const float*>(a.blob + hd->off_pi0_f);
  const int lane = threadIdx.x & 31;
  const int k = lane % KP;
  const int warp_global = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int trial = warp_global * GPW + lane / KP;
  const bool active = trial < a.n_trials;
  const int K = a.K;
  const bool kvalid = k < K;
  const long long beg = active ? a.offsets[trial] : 0;
  const int T = active ? (int)(a.offsets[trial + 1] - beg) : 0;
  // all lanes of a warp iterate to the longest trial in the warp so shuffles stay convergent
  int Tmax = T;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) Tmax = max(Tmax, __shfl_xor_sync(0xffffffffu, Tmax, o));
  if (Tmax == 0) return;

  float Pcol[KP], Prow[KP];
#pragma unroll
  for (int j = 0; j < KP; ++j) {
    Pcol[j] = Pg[j * KP + k];
    Prow[j] = Pg[k * KP + j];
  }
  const bool want_post = a.Ez != nullptr;
  const float* Bp = a.Bsc + beg * K + k;
  const float* mp = a.mx + beg;
  float* Ep = want_post ? a.Ez + beg * K + k : nullptr;

  // ---------------- forward
  // Lazily normalised recursion.  The state a_t carried on the serial chain is
  //     a_t(k) = [sum_j a_{t-1}(j) P(j,k)] * b_t(k) / S_{t-1},     S_{t-1} = sum_j a_{t-1}(j),
  // so sum_k a_t(k) = c_t (the per-step normaliser) and alpha_hat_t = a_t / S_t.  Each step the KP
  // lanes of a trial exchange their state through a 2 x 32-float shared-memory slot (one STS, four
  // broadcast LDS.128 per lane -- cheaper and shorter than KP shuffles), so every lane has all
  // a_{t-1}(j) and S_{t-1} is a LOCAL sum: no cross-lane reduction sits between two steps.
  // 1 / c_t is kept (a.cinv) so that the backward sweep needs no reduction at all.
  __shared__ __align__(16) float xch[4][2][32];
  const int wib = threadIdx.x >> 5;
  const int gbase = (lane / KP) * KP;
  constexpr int PF = 8;
  double logZ = 0.0;
  float acur = 0.f;                 // a_t(k)
  float m0 = 0.f;
  if (Tmax > 0) {
    const float b0 = (T > 0 && kvalid) ? __ldg(Bp) : 0.f;
    acur = pi0g[k] * b0;
    if (T > 0) m0 = __ldg(mp);
  }
  logZ += (double)m0;
  float* cip = a.cinv + beg;
  int par = 0;
  auto exchange = [&](float mine, float (&v)[KP]) {
    xch[wib][par][lane] = mine;
    __syncwarp();
    if (KP >= 4) {
#pragma unroll
      for (int j = 0; j < KP; j += 4) {
        const float4 q = *reinterpret_cast<const float4*>(&xch[wib][par][gbase + j]);
        v[j] = q.x; v[j + 1] = q.y; v[j + 2] = q.z; v[j + 3] = q.w;
      }
    } else {
#pragma unroll
      for (int j = 0; j < KP; ++j) v[j] = xch[wib][par][gbase + j];
    }
    par ^= 1;
  };
  // steps every trial of this warp has (no per-step conditionals, prefetched inputs) ...
  int Tmin = active ? T : 0;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) Tmin = min(Tmin, __shfl_xor_sync(0xffffffffu, Tmin, o));
  // one step of the forward recursion: consumes a_{t-1}, finalises step t-1, produces a_t
  auto fwd_step = [&](int t, float b, float m, bool has_prev, bool has_cur, float& lsum) {
    float v[KP];
    exchange(acur, v);
    float s0 = 0.f, s1 = 0.f, q0 = 0.f, q1 = 0.f;
#pragma unroll
    for (int j = 0; j < KP; j += 2) {
      s0 = fmaf(v[j], Pcol[j], s0);
      s1 = fmaf(v[j + 1], Pcol[j + 1], s1);
      q0 += v[j];
      q1 += v[j + 1];
    }
    const float S = q0 + q1;                       // = c_{t-1}
    const float inv = S > 0.f ? __fdividef(1.f, S) : 0.f;
    if (has_prev) {
      lsum += __logf(S);
      if (want_post && kvalid) Ep[(long long)(t - 1) * K] = acur * inv;
      if (want_post && k == 0) cip[t - 1] = inv;
    }
    if (has_cur) {
      acur = (s0 + s1) * b * inv;
      lsum += m;
    }
  };
  int t = 1;
  if (Tmin > 1) {
    float bq[PF], mq[PF];
#pragma unroll
    for (int i = 0; i < PF; ++i) {
      const int tt = min(1 + i, Tmin - 1);
      bq[i] = kvalid ? __ldg(Bp + (long long)tt * K) : 0.f;
      mq[i] = __ldg(mp + tt);
    }
    for (; t + PF <= Tmin; t += PF) {
      float nb[PF], nm[PF];
#pragma unroll
      for (int i = 0; i < PF; ++i) {
        const int tt = min(t + PF + i, Tmin - 1);
        nb[i] = kvalid ? __ldg(Bp + (long long)tt * K) : 0.f;
        nm[i] = __ldg(mp + tt);
      }
      float lsum = 0.f;
#pragma unroll
      for (int i = 0; i < PF; ++i) fwd_step(t + i, bq[i], mq[i], true, true, lsum);
      logZ += (double)lsum;
#pragma unroll
      for (int i = 0; i < PF; ++i) { bq[i] = nb[i]; mq[i] = nm[i]; }
    }
  }
  // ... and the ragged remainder (trial ends inside the warp's range), one checked step at a time
  for (; t <= Tmax; ++t) {
    const bool cur = t < T;
    const float b = (cur && kvalid) ? __ldg(Bp + (long long)t * K) : 0.f;
    const float m = cur ? __ldg(mp + t) : 0.f;
    float lsum = 0.f;
    fwd_step(t, b, m, t - 1 < T, cur, lsum);
    logZ += (double)lsum;
  }
  if (active && k == 0 && a.logZ) a.logZ[trial] = logZ;
  if (!want_post && a.Ezz == nullptr) return;

  // ---------------- backward (lane j = k owns row j of P)
  // Scaled with the forward normalisers: beta_hat_t(j) = sum_k P(j,k) b_{t+1}(k) beta_hat_{t+1}(k) / c_{t+1},
  // which makes gamma_t = alpha_hat_t * beta_hat_t and xi_t(j,k) = alpha_hat_t(j) P(j,k) v_k exactly
  // normalised -- no reduction and no division in this sweep.
  float X[KP];
#pragma unroll
  for (int j = 0; j < KP; ++j) X[j] = 0.f;
  float beta = 1.f;       // beta_hat_{t+1}(k)
  auto bwd_step = [&](int t, float bc, float w, bool on) {
    float vk[KP];
    exchange(bc * beta, vk);                       // v_k = b_{t+1}(k) beta_hat_{t+1}(k) / c_{t+1}
    float u0 = 0.f, u1 = 0.f;
#pragma unroll
    for (int j = 0; j < KP; j += 2) {
      u0 = fmaf(Prow[j], vk[j], u0);
      u1 = fmaf(Prow[j + 1], vk[j + 1], u1);
    }
    if (on) {
      beta = u0 + u1;
#pragma unroll
      for (int j = 0; j < KP; ++j) X[j] = fmaf(w, vk[j], X[j]);
      if (kvalid) Ep[(long long)t * K] = w * beta;
    }
  };
  // ragged head (some trials of the warp are shorter): checked steps
  int tb = Tmax - 2;
  for (; tb >= 0 && tb + 1 >= Tmin; --tb) {
    const bool on = tb + 1 < T;
    const float bc = (on && kvalid) ? __ldg(Bp + (long long)(tb + 1) * K) * cip[tb + 1] : 0.f;
    const float w = (on && kvalid) ? Ep[(long long)tb * K] : 0.f;
    bwd_step(tb, bc, w, on);
  }
  // uniform part: every trial has steps tb and tb+1
  if (tb >= 0) {
    float bq[PF], wq[PF];
    auto fetch = [&](int t0, float (&bb)[PF], float (&ww)[PF]) {
#pragma unroll
      for (int i = 0; i < PF; ++i) {
        const int tt = max(t0 - i, 0);
        bb[i] = kvalid ? __ldg(Bp + (long long)(tt + 1) * K) * cip[tt + 1] : 0.f;
        ww[i] = kvalid ? Ep[(long long)tt * K] : 0.f;
      }
    };
    fetch(tb, bq, wq);
    for (; tb - PF + 1 >= 0; tb -= PF) {
      float nb[PF], nw[PF];
      fetch(tb - PF, nb, nw);
#pragma unroll
      for (int i = 0; i < PF; ++i) bwd_step(tb - i, bq[i], wq[i], true);
#pragma unroll
      for (int i = 0; i < PF; ++i) { bq[i] = nb[i]; wq[i] = nw[i]; }
    }
    for (; tb >= 0; --tb) {
      const float bc = kvalid ? __ldg(Bp + (long long)(tb + 1) * K) * cip[tb + 1] : 0.f;
      const float w = kvalid ? Ep[(long long)tb * K] : 0.f;
      bwd_step(tb, bc, w, true);
    }
  }
  if (active && kvalid && a.Ezz) {
    float* out = a.Ezz + ((long long)trial * K + k) * K;
#pragma unroll
    for (int j = 0; j < KP; ++j)
      if (j < K) out[j] = X[j] * Prow[j];        // compile-time indices keep X / Prow in registers
  }
}

// ------------------------------------------------------------------------------------------------
// forward-backward scan, meet-in-the-middle form.
//
// The serial chain of the E-step is 2T dependent steps when the backward sweep waits for the forward
// one.  Here a trial is owned by TWO lane groups in two different warps of the same block: the
// forward group sweeps t = 0 .. T-1, the backward group t = T-1 .. 0, concurrently, each with its own
// lazy normalisation.  Until they cross at h = T/2 each stores its message (alpha_hat into the Ez
// buffer, beta_tilde into the workspace); after one block-level barrier each finds the other's message
// already in memory and finishes the posteriors on the fly:
//     gamma_t    = alpha_hat_t * beta_t / G_t,                  G_t = sum_k alpha_hat_t(k) beta_t(k)
//     xi_t(j,k)  = alpha_hat_t(j) P(j,k) b_{t+1}(k) beta_{t+1}(k) / (G_t D_t)        (backward group, t < h)
//                = alpha_hat_t(j) P(j,k) b_{t+1}(k) beta_{t+1}(k) / (c_{t+1} G_{t+1})  (forward group, t >= h)
// (D_t, c_t = the local normalisers).  G_t is a KP-lane shuffle reduction that does not feed the
// recursion, so it stays off the dependent chain; the chain itself is T steps of
// STS -> LDS.128 x KP/4 -> KP FMAs.  No third pass, no extra HBM traffic beyond one message array.
// ------------------------------------------------------------------------------------------------
struct Scan2Args {
  const unsigned char* blob;
  const float* Bsc;
  const float* mx;
  const long long* offsets;
  int n_trials, K;
  float* Ez;      // nullable (then forward only: log normalisers)
  float* Ezz;     // nullable
  double* logZ;   // nullable
  float* beta;    // (total_T, K) workspace
};

template <int PF, class In, class Load, class Step>
__device__ __forceinline__ void run_range(int t0, int n, int dir, Load load, Step step) {
  int done = 0;
  if (n >= PF) {
    In q[PF];
#pragma unroll
    for (int i = 0; i < PF; ++i) q[i] = load(t0 + dir * i);
    for (; done + 2 * PF <= n; done += PF) {
      In nq[PF];
#pragma unroll
      for (int i = 0; i < PF; ++i) nq[i] = load(t0 + dir * (done + PF + i));
#pragma unroll
      for (int i = 0; i < PF; ++i) step(t0 + dir * (done + i), q[i]);
#pragma unroll
      for (int i = 0; i < PF; ++i) q[i] = nq[i];
    }
#pragma unroll
    for (int i = 0; i < PF; ++i) step(t0 + dir * (done + i), q[i]);
    done += PF;
  }
  for (; done < n; ++done) step(t0 + dir * done, load(t0 + dir * done));
}

struct ScanIn { float b, m, x; };

// (Measured and dropped: warp-uniform trip counts with full-mask __syncwarp() in place of the sub-warp mask, which
// compiles to MATCH.ANY + REDUX + VOTE per barrier -- the early-out of dead steps costs more: 461 vs 443 us per E-step.)
// packed fp32 pairs (sm_100 FFMA2 / FADD2): the scan is issue-bound, so two lanes per instruction
__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) {
  unsigned long long ua = *reinterpret_cast<unsigned long long*>(&a), ub = *reinterpret_cast<unsigned long long*>(&b),
                     uc = *reinterpret_cast<unsigned long long*>(&c), ud;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(ud) : "l"(ua), "l"(ub), "l"(uc));
  return *reinterpret_cast<float2*>(&ud);
}
__device__ __forceinline__ float2 fadd2(float2 a, float2 b) {
  unsigned long long ua = *reinterpret_cast<unsigned long long*>(&a), ub = *reinterpret_cast<unsigned long long*>(&b), ud;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(ud) : "l"(ua), "l"(ub));
  return *reinterpret_cast<float2*>(&ud);
}

// 1 / x for a positive normal x, 0 otherwise: one MUFU + a select (the guarded __fdividef spent six instructions of
// every step on denormal handling the scaled messages never need)
__device__ __forceinline__ float rcp_pos(float x) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return x > 1.1754944e-38f ? r : 0.f;
}

// log2 of a positive normal x in one MUFU (the normalisers S_t are sums of scaled probabilities, >= 1e-30)
__device__ __forceinline__ float lg2_pos(float x) {
  float r;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}

template <int KP, bool POST>
__global__ void __launch_bounds__(128, KP == 32 ? 2 : 4) scan2_kernel(const Scan2Args a) {
  constexpr int GPW = 32 / KP;                     // trials per warp
  constexpr int H2 = KP / 2;                       // packed pairs per message
  constexpr int PF = 8;                            // prefetch depth of the first halves
  constexpr int PF2 = 4;                           // second halves carry more live state
  const BlobHeader* hd = reinterpret_cast<const BlobHeader*>(a.blob);
  const float* Pg = reinterpret_cast<const float*>(a.blob + hd->off_P_f);     // KP x KP, zero padded
  const float* pi0g = reinterpret_cast<const float*>(a.blob + hd->off_pi0_f);
  const int lane = threadIdx.x & 31;
  const int wib = threadIdx.x >> 5;
  const int k = lane % KP;
  const int grp = lane / KP;
  // POST: warps 0,1 run the forward sweeps of trial sets 0,1 of this block, warps 2,3 the backward
  // sweeps of the same sets.  !POST (log-likelihood only): four forward warps.
  const bool is_bwd = POST && wib >= 2;
  const int set = POST ? (wib & 1) : wib;
  const int trial = (blockIdx.x * (POST ? 2 : 4) + set) * GPW + grp;
  const bool active = trial < a.n_trials;
  const int K = a.K;
  const bool kvalid = k < K && active;
  const long long beg = active ? a.offsets[trial] : 0;
  const int T = active ? (int)(a.offsets[trial + 1] - beg) : 0;
  const int h = T / 2;
  const unsigned gmask = KP == 32 ? 0xffffffffu : (((1u << KP) - 1u) << (grp * KP));

  __shared__ __align__(16) float xch[4][2][32];
  __shared__ __align__(16) float xg[4][2][32];
  __shared__ float xz[2][GPW][KP][KP + 1];
  const int gbase = grp * KP;
  int par = 0;
  // all-gather of one float per lane inside the KP-lane group (double-buffered by parity: a lane is
  // at most one step ahead of its group)
  auto exchange = [&](float mine, float2 (&v)[H2]) {
    xch[wib][par][lane] = mine;
    __syncwarp(gmask);
    if (KP >= 4) {
#pragma unroll
      for (int j = 0; j < KP; j += 4) {
        const float4 q = *reinterpret_cast<const float4*>(&xch[wib][par][gbase + j]);
        v[j / 2] = make_float2(q.x, q.y);
        v[j / 2 + 1] = make_float2(q.z, q.w);
      }
    } else {
      v[0] = *reinterpret_cast<const float2*>(&xch[wib][par][gbase]);
    }
    par ^= 1;
  };
  // same, plus the group SUM of a second float: the posterior normaliser G travels with the message
  // instead of through a 4-deep shuffle butterfly, so it never lengthens the in-order critical path
  auto exchange2 = [&](float mine, float gval, float2 (&v)[H2], float& G) {
    xch[wib][par][lane] = mine;
    xg[wib][par][lane] = gval;
    __syncwarp(gmask);
    float2 acc = make_float2(0.f, 0.f);
    if (KP >= 4) {
#pragma unroll
      for (int j = 0; j < KP; j += 4) {
        const float4 q = *reinterpret_cast<const float4*>(&xch[wib][par][gbase + j]);
        v[j / 2] = make_float2(q.x, q.y);
        v[j / 2 + 1] = make_float2(q.z, q.w);
        const float4 e = *reinterpret_cast<const float4*>(&xg[wib][par][gbase + j]);
        acc = fadd2(acc, make_float2(e.x, e.y));
        acc = fadd2(acc, make_float2(e.z, e.w));
      }
    } else {
      v[0] = *reinterpret_cast<const float2*>(&xch[wib][par][gbase]);
      acc = *reinterpret_cast<const float2*>(&xg[wib][par][gbase]);
    }
    G = acc.x + acc.y;
    par ^= 1;
  };
  auto pair_barrier = [&]() {
    __syncwarp();
    if (POST) asm volatile("bar.sync %0, 64;" ::"r"(1 + set) : "memory");
  };
  // dot(v, p) and sum(v) of a packed message
  auto dot_sum = [&](const float2 (&v)[H2], const float2 (&p)[H2], float& dot, float& sum) {
    float2 s = make_float2(0.f, 0.f), q = make_float2(0.f, 0.f);
#pragma unroll
    for (int j = 0; j < H2; ++j) {
      s = ffma2(v[j], p[j], s);
      q = fadd2(q, v[j]);
    }
    dot = s.x + s.y;
    sum = q.x + q.y;
  };

  const float* Bp = a.Bsc + beg * K + k;
  const float* mp = a.mx + beg;
  float* Ep = POST ? a.Ez + beg * K + k : nullptr;
  float* Bt = POST ? a.beta + beg * K + k : nullptr;

  if (!is_bwd) {
    // =============================== forward group (lane k owns column k of P)
    float2 Pcol[H2];
#pragma unroll
    for (int j = 0; j < H2; ++j) Pcol[j] = make_float2(Pg[(2 * j) * KP + k], Pg[(2 * j + 1) * KP + k]);
    // The message on the serial chain is a_t = (P^T a_{t-1}) b_t / S_{t-1} with S_{t-1} = sum_j a_{t-1}(j)
    // taken from the SAME exchange (a local sum: every lane holds the whole vector), so sum_k a_t(k)
    // = c_t stays in [min P, 1].  (Scaling by the older S_{t-2} would take the reciprocal off the
    // chain but obeys L_t = log c_t + L_{t-1} - L_{t-2}, a marginally stable recurrence whose
    // random-walk growth overflows fp32 within ~1000 steps.)  gamma and xi are normalised explicitly
    // by G, and log Z = sum_t (log S_t + m_t).
    double logZ = 0.0;
    float acur = 0.f, bprev = 0.f;
    float rcur = 1.f;                 // scale that produced acur
    if (T > 0) {
      bprev = kvalid ? __ldg(Bp) : 0.f;
      acur = pi0g[k] * bprev;
      logZ = (double)__ldg(mp);
    }
    // first half: t = 1 .. h; stores the (unnormalised) a_{t-1}, produces a_t
    // run_range calls load() and step() for consecutive t, so the streams are walked with running pointers
    // (recomputing base + t * K in 64 bits was ~15 of the 53 instructions of a step)
    const float* bl = Bp + K;                        // b_t of the next load (t = 1 ..)
    const float* ml = mp + 1;
    float* es = Ep;                                  // a_{t-1} / gamma_{t-1} of the next step
    auto load1 = [&](int) {
      ScanIn in;
      in.b = kvalid ? __ldg(bl) : 0.f;
      in.m = __ldg(ml);
      in.x = 0.f;
      bl += K;
      ++ml;
      return in;
    };
    auto step1 = [&](int, const ScanIn& in) {
      float2 v[H2];
      exchange(acur, v);
      float dot, S;                                  // S = S_{t-1}
      dot_sum(v, Pcol, dot, S);
      if (POST) {
        if (kvalid) *es = acur;
        es += K;
      }
      const float inv = rcp_pos(S);
      logZ += (double)fmaf(lg2_pos(S), 0.69314718f, in.m);            // fp64 accumulation, off the dependent chain
      acur = dot * in.b * inv;
      rcur = inv;
      bprev = in.b;
    };
    const int n1 = POST ? h : (T > 0 ? T - 1 : 0);
    run_range<PF, ScanIn>(1, n1, 1, load1, step1);
    if (!POST) {
      // finalise the last step: log S_{T-1}
      if (T > 0) {
        float2 v[H2];
        exchange(acur, v);
        float dot, S;
        dot_sum(v, Pcol, dot, S);
        logZ += (double)(lg2_pos(S) * 0.69314718f);
      }
      if (active && k == 0 && a.logZ) a.logZ[trial] = logZ;
      return;
    }
    pair_barrier();
    // second half: t = h+1 .. T; finalises step t-1 into gamma_{t-1} and xi_{t-2}; t == T has no a_t
    float2 X[H2], vprev[H2];
#pragma unroll
    for (int j = 0; j < H2; ++j) { X[j] = make_float2(0.f, 0.f); vprev[j] = make_float2(0.f, 0.f); }
    const float* xl = Bt + (long long)h * K;         // beta_tilde_{t-1} of the next load (t = h+1 ..)
    auto load2 = [&](int t) {
      ScanIn in;
      const bool cur = t < T;
      in.b = (cur && kvalid) ? __ldg(bl) : 0.f;
      in.m = cur ? __ldg(ml) : 0.f;
      in.x = kvalid ? *xl : 0.f;                                  // beta_tilde_{t-1}(k)
      bl += K;
      ++ml;
      xl += K;
      return in;
    };
    auto step2 = [&](int, const ScanIn& in) {
      float2 v[H2];
      const float ab = acur * in.x;
      float G;                                       // sum_k a_{t-1}(k) beta_{t-1}(k)
      exchange2(acur, ab, v, G);
      float dot, S;
      dot_sum(v, Pcol, dot, S);
      const float rG = rcp_pos(G);
      if (kvalid) *es = ab * rG;
      es += K;
      const float w = bprev * in.x * rcur * rG;      // xi_{t-2}: vprev is all-zero on the first step
      const float2 w2 = make_float2(w, w);
#pragma unroll
      for (int j = 0; j < H2; ++j) {
        X[j] = ffma2(vprev[j], w2, X[j]);
        vprev[j] = v[j];
      }
      const float inv = rcp_pos(S);
      logZ += (double)fmaf(lg2_pos(S), 0.69314718f, in.m);            // in.m = 0 on the last step (t == T)
      acur = dot * in.b * inv;
      rcur = inv;
      bprev = in.b;
    };
    run_range<PF2, ScanIn>(h + 1, T - h, 1, load2, step2);
    if (active && k == 0 && a.logZ) a.logZ[trial] = logZ;
    pair_barrier();                                  // backward group's xi rows are in xz
    if (kvalid && a.Ezz) {
      float* out = a.Ezz + (long long)trial * K * K + k;
#pragma unroll
      for (int j = 0; j < KP; ++j)
        if (j < K) {
          const float xj = (j & 1) ? X[j / 2].y : X[j / 2].x;
          const float pj = (j & 1) ? Pcol[j / 2].y : Pcol[j / 2].x;
          out[(long long)j * K] = pj * (xj + xz[set][grp][j][k]);
        }
    }
  } else {
    // =============================== backward group (lane j = k owns row j of P)
    float2 Prow[H2];
#pragma unroll
    for (int j = 0; j < H2; ++j) Prow[j] = make_float2(Pg[k * KP + 2 * j], Pg[k * KP + 2 * j + 1]);
    float beta = 1.f;                                // beta_tilde_{t+1}(j)
    if (T > 0 && kvalid) Bt[(T - 1) * K] = 1.f;
    const float* bl = Bp + (long long)(T - 1) * K;   // b_{t+1} of the next load (t = T-2 .. 0)
    float* bs = Bt + (long long)(T - 2) * K;         // beta_tilde_t of the next first-half step
    auto load1 = [&](int) {
      ScanIn in;
      in.b = kvalid ? __ldg(bl) : 0.f;
      in.m = 0.f;
      in.x = 0.f;
      bl -= K;
      return in;
    };
    auto step1 = [&](int, const ScanIn& in) {
      float2 vk[H2];
      exchange(in.b * beta, vk);
      float u, Dn;
      dot_sum(vk, Prow, u, Dn);
      beta = u * rcp_pos(Dn);
      if (kvalid) *bs = beta;
      bs -= K;
    };
    run_range<PF, ScanIn>(T - 2, max(T - 1 - h, 0), -1, load1, step1);      // t = T-2 .. h
    pair_barrier();
    // second half: t = h-1 .. 0.  gamma_t / xi_t need G_t = sum_j a_t(j) beta_t(j), which exists only
    // at the END of step t: it rides along with the next step's message, so step t is finalised
    // one iteration later (and t = 0 by a trailing exchange).
    float2 X[H2], vkp[H2];
#pragma unroll
    for (int j = 0; j < H2; ++j) { X[j] = make_float2(0.f, 0.f); vkp[j] = make_float2(0.f, 0.f); }
    float al_p = 0.f, ab_p = 0.f, rho_p = 0.f;
    const float* xl = Ep + (long long)(h - 1) * K;   // a_t of the next load (t = h-1 .. 0)
    auto load2 = [&](int) {
      ScanIn in;
      in.b = kvalid ? __ldg(bl) : 0.f;
      in.m = 0.f;
      in.x = kvalid ? *xl : 0.f;                                  // a_t(j) stored by the forward group
      bl -= K;
      xl -= K;
      return in;
    };
    float* gs = Ep + (long long)(h - 1) * K;         // gamma_tp of the next finalise (tp = h-1 .. 0)
    auto finalise = [&](int, float G) {
      const float rG = rcp_pos(G);
      if (kvalid) *gs = ab_p * rG;
      gs -= K;
      const float w = al_p * rho_p * rG;
      const float2 w2 = make_float2(w, w);
#pragma unroll
      for (int j = 0; j < H2; ++j) X[j] = ffma2(vkp[j], w2, X[j]);
    };
    auto step2 = [&](int t, const ScanIn& in) {
      float2 vk[H2];
      float G;
      exchange2(in.b * beta, ab_p, vk, G);
      if (t + 1 < h) finalise(t + 1, G);
      float u, Dn;
      dot_sum(vk, Prow, u, Dn);
      rho_p = rcp_pos(Dn);
      beta = u * rho_p;
      al_p = in.x;
      ab_p = in.x * beta;
#pragma unroll
      for (int j = 0; j < H2; ++j) vkp[j] = vk[j];
    };
    run_range<PF2, ScanIn>(h - 1, h, -1, load2, step2);                       // t = h-1 .. 0
    if (h > 0) {
      float2 vk[H2];
      float G;
      exchange2(0.f, ab_p, vk, G);
      finalise(0, G);
    }
#pragma unroll
    for (int j = 0; j < H2; ++j) {
      xz[set][grp][k][2 * j] = X[j].x;
      xz[set][grp][k][2 * j + 1] = X[j].y;
    }
    pair_barrier();
  }
}

// ------------------------------------------------------------------------------------------------
// forward-backward scan, meet-in-the-middle form, SPL states per lane.
//
// scan2_kernel gives every state its own lane: a C4 step is 84 warp instructions for TWO trials (the
// all-gather, the normaliser and the address arithmetic are replicated in all 16 lanes of a trial), and ncu shows
// the kernel issue-bound (56 % issue utilisation at 3.5 warps per scheduler, 173 M warp instructions).  Here a lane
// owns SPL consecutive states (K = 16: 4 lanes per trial, 8 trials per warp): the per-step overhead is shared by
// SPL states, the mat-vec of a lane is SPL independent chains of KP/2 packed FMAs on the all-gathered message,
// and the normaliser sums travel with the message as per-lane partial sums.  The mathematics (scaling, lazy
// finalisation of gamma / xi one step late, fp64 log normaliser) is scan2_kernel's, statement for statement.
// Needs K % SPL == 0 (a lane's states are all real or all padding; vector loads stay aligned).
// Block = 64 threads: warp 0 forward, warp 1 backward sweeps of the same 32 / (KP / SPL) trials.
// ------------------------------------------------------------------------------------------------
template <int SPL>
struct ScanInV {
  float b[SPL];
  float x[SPL];
  float m;
};

template <int SPL>
__device__ __forceinline__ void ldv_nc(const float* p, float (&v)[SPL]) {
  if (SPL == 4) {
    const float4 q = __ldg(reinterpret_cast<const float4*>(p));
    v[0] = q.x; v[1] = q.y; v[SPL > 2 ? 2 : 0] = q.z; v[SPL > 2 ? 3 : 1] = q.w;
  } else {
    const float2 q = __ldg(reinterpret_cast<const float2*>(p));
    v[0] = q.x; v[1] = q.y;
  }
}
// coherent loads: messages written by the partner warp earlier in this kernel
template <int SPL>
__device__ __forceinline__ void ldv_cg(const float* p, float (&v)[SPL]) {
  if (SPL == 4) {
    const float4 q = __ldcg(reinterpret_cast<const float4*>(p));
    v[0] = q.x; v[1] = q.y; v[SPL > 2 ? 2 : 0] = q.z; v[SPL > 2 ? 3 : 1] = q.w;
  } else {
    const float2 q = __ldcg(reinterpret_cast<const float2*>(p));
    v[0] = q.x; v[1] = q.y;
  }
}
template <int SPL>
__device__ __forceinline__ void stv(float* p, const float (&v)[SPL]) {
  if (SPL == 4) *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[SPL > 2 ? 2 : 0], v[SPL > 2 ? 3 : 1]);
  else *reinterpret_cast<float2*>(p) = make_float2(v[0], v[1]);
}

// ring prefetch: PF inputs in flight, slot i reloaded right after it is consumed.  The trip count nmax is the
// warp-wide maximum of the lanes' own counts n: every lane runs every iteration (steps past its own range are
// `live = false` no-ops), so the exchange inside a step can use a plain full-mask __syncwarp().  A sub-warp mask
// compiles to MATCH.ANY + REDUX + VOTE in front of every barrier -- measured at ~300 cycles of a 540-cycle step
// when one warp per scheduler has nothing to hide them behind.
template <int PF, class In, class Load, class Step>
__device__ __forceinline__ void run_ring(int t0, int n, int dir, Load load, Step step) {
  const int nmax = __reduce_max_sync(0xffffffffu, n);
  In q[PF];
#pragma unroll
  for (int i = 0; i < PF; ++i) q[i] = load(t0 + dir * i, i < n);
  for (int done = 0; done < nmax; done += PF) {
#pragma unroll
    for (int i = 0; i < PF; ++i) {
      const int idx = done + i;
      if (idx < nmax) {
        const In cur = q[i];
        q[i] = load(t0 + dir * (idx + PF), idx + PF < n);
        step(t0 + dir * idx, cur, idx < n);
      }
    }
  }
}

template <int KP, int SPL, bool POST>
__global__ void __launch_bounds__(64) scan4_kernel(const Scan2Args a) {
  constexpr int LPT = KP / SPL;                    // lanes per trial
  constexpr int GPW = 32 / LPT;                    // trials per warp
  constexpr int H2 = KP / 2;                       // packed pairs per message
  constexpr int STRIDE = KP + 2 * LPT + 4;         // message + two partial-sum arrays; the pad keeps LDS.128 of the groups conflict-free
  constexpr int PF = 6;                            // prefetch depth of the first halves
  constexpr int PF2 = 2;                           // second halves: registers go to the xi accumulators, so the
  constexpr int PFL2 = 8;                          //   likelihood stream is pulled into L2 PFL2 steps ahead instead
  // Few warps carry the whole problem (C4: 512), so the loads in flight, not the latency of one, bound the likelihood
  // stream (Little: 1 TB/s x 1 us = 1 MB; PF register slots x 64 B x 4096 chains = 1.5 MB at best).  Every lane
  // therefore also pulls its row PFFAR steps ahead into L2; the register ring then only has to cover L2 latency.
  constexpr int PFFAR = 32;
  static_assert(LPT >= 4 && LPT % 4 == 0, "partial sums are read as float4");
  const BlobHeader* hd = reinterpret_cast<const BlobHeader*>(a.blob);
  const float* Pg = reinterpret_cast<const float*>(a.blob + hd->off_P_f);     // KP x KP, zero padded
  const float* pi0g = reinterpret_cast<const float*>(a.blob + hd->off_pi0_f);
  const int lane = threadIdx.x & 31;
  const int wib = threadIdx.x >> 5;
  const int lig = lane % LPT;
  const int grp = lane / LPT;
  const int j0 = lig * SPL;                        // first state of this lane
  const bool is_bwd = POST && wib == 1;
  const int trial = (POST ? blockIdx.x : blockIdx.x * 2 + wib) * GPW + grp;
  const bool active = trial < a.n_trials;
  const int K = a.K;
  const bool kvalid = active && j0 < K;
  const long long beg = active ? a.offsets[trial] : 0;
  const int T = active ? (int)(a.offsets[trial + 1] - beg) : 0;
  const int h = T / 2;

  __shared__ __align__(16) float xch[2][2][GPW * STRIDE];
  __shared__ float xz[POST ? GPW : 1][POST ? KP : 1][POST ? KP + 1 : 1];
  float* xme = &xch[wib][0][grp * STRIDE];
  int par = 0;
  // all-gather of the lanes' SPL-float message slices inside the trial's lane group, plus the group sums of up to
  // two scalars (per-lane partials p0, p1); double-buffered by parity (a lane is at most one step ahead of its group)
  auto exchange = [&](const float (&mine)[SPL], float p0, float p1, bool two, float2 (&v)[H2], float& s0, float& s1) {
    float* x = xme + par * (GPW * STRIDE);
    stv<SPL>(x + j0, mine);
    x[KP + lig] = p0;
    if (two) x[KP + LPT + lig] = p1;
    __syncwarp();
#pragma unroll
    for (int j = 0; j < KP; j += 4) {
      const float4 q = *reinterpret_cast<const float4*>(x + j);
      v[j / 2] = make_float2(q.x, q.y);
      v[j / 2 + 1] = make_float2(q.z, q.w);
    }
    float2 acc = make_float2(0.f, 0.f), acc1 = make_float2(0.f, 0.f);
#pragma unroll
    for (int j = 0; j < LPT; j += 4) {
      const float4 e = *reinterpret_cast<const float4*>(x + KP + j);
      acc = fadd2(acc, fadd2(make_float2(e.x, e.y), make_float2(e.z, e.w)));
      if (two) {
        const float4 f = *reinterpret_cast<const float4*>(x + KP + LPT + j);
        acc1 = fadd2(acc1, fadd2(make_float2(f.x, f.y), make_float2(f.z, f.w)));
      }
    }
    s0 = acc.x + acc.y;
    s1 = acc1.x + acc1.y;
    par ^= 1;
  };
  // out[s] = sum_i v(i) * M[s](i): SPL independent chains of packed FMAs
  auto dots = [&](const float2 (&v)[H2], const float2 (&M)[SPL][H2], float (&out)[SPL]) {
    float2 acc[SPL];
#pragma unroll
    for (int s = 0; s < SPL; ++s) acc[s] = make_float2(0.f, 0.f);
#pragma unroll
    for (int j = 0; j < H2; ++j)
#pragma unroll
      for (int s = 0; s < SPL; ++s) acc[s] = ffma2(v[j], M[s][j], acc[s]);
#pragma unroll
    for (int s = 0; s < SPL; ++s) out[s] = acc[s].x + acc[s].y;
  };
  auto lsum = [&](const float (&v)[SPL]) {
    float r = v[0];
#pragma unroll
    for (int s = 1; s < SPL; ++s) r += v[s];
    return r;
  };

  const float* Bp = a.Bsc + beg * K + j0;
  const float* mp = a.mx + beg;
  float* Ep = POST ? a.Ez + beg * K + j0 : nullptr;
  float* Bt = POST ? a.beta + beg * K + j0 : nullptr;
  typedef ScanInV<SPL> In;

  if (!is_bwd) {
    // =============================== forward group (the lane owns columns j0 .. j0+SPL-1 of P)
    float2 Pc[SPL][H2];
#pragma unroll
    for (int s = 0; s < SPL; ++s)
#pragma unroll
      for (int j = 0; j < H2; ++j) Pc[s][j] = make_float2(Pg[(2 * j) * KP + j0 + s], Pg[(2 * j + 1) * KP + j0 + s]);
    double logZ = 0.0;
    float acur[SPL], bprev[SPL];
    float rcur = 1.f;
#pragma unroll
    for (int s = 0; s < SPL; ++s) { acur[s] = 0.f; bprev[s] = 0.f; }
    if (T > 0) {
      if (kvalid) ldv_nc<SPL>(Bp, bprev);
#pragma unroll
      for (int s = 0; s < SPL; ++s) acur[s] = pi0g[j0 + s] * bprev[s];
      logZ = (double)__ldg(mp);
    }
    auto load1 = [&](int t, bool ok) {
      In in;
#pragma unroll
      for (int s = 0; s < SPL; ++s) { in.b[s] = 0.f; in.x[s] = 0.f; }
      if (kvalid && ok) ldv_nc<SPL>(Bp + (long long)t * K, in.b);
      if (kvalid && t + PFFAR < T) asm volatile("prefetch.global.L2 [%0];" ::"l"(Bp + (long long)(t + PFFAR) * K));
      in.m = ok ? __ldg(mp + t) : 0.f;
      return in;
    };
    auto step1 = [&](int t, const In& in, bool live) {
      float2 v[H2];
      float S, unused;                               // S = S_{t-1} = sum_j a_{t-1}(j)
      exchange(acur, lsum(acur), 0.f, false, v, S, unused);
      if (!live) return;
      float dot[SPL];
      dots(v, Pc, dot);
      if (POST && kvalid) stv<SPL>(Ep + (long long)(t - 1) * K, acur);
      const float inv = S > 0.f ? __fdividef(1.f, S) : 0.f;
      logZ += (double)(__logf(S) + in.m);            // fp64 accumulation, off the dependent chain
#pragma unroll
      for (int s = 0; s < SPL; ++s) { acur[s] = dot[s] * in.b[s] * inv; bprev[s] = in.b[s]; }
      rcur = inv;
    };
    const int n1 = POST ? h : (T > 0 ? T - 1 : 0);
    run_ring<PF, In>(1, n1, 1, load1, step1);
    if (!POST) {
      float2 v[H2];
      float S, unused;
      exchange(acur, lsum(acur), 0.f, false, v, S, unused);
      if (T > 0) logZ += (double)__logf(S);
      if (active && lig == 0 && a.logZ) a.logZ[trial] = logZ;
      return;
    }
    __syncthreads();
    // second half: t = h+1 .. T; finalises step t-1 into gamma_{t-1} and xi_{t-2}; t == T has no a_t
    float2 X[SPL][H2], vprev[H2];
#pragma unroll
    for (int j = 0; j < H2; ++j) {
      vprev[j] = make_float2(0.f, 0.f);
#pragma unroll
      for (int s = 0; s < SPL; ++s) X[s][j] = make_float2(0.f, 0.f);
    }
    auto load2 = [&](int t, bool ok) {
      In in;
      const bool cur = ok && t < T;
#pragma unroll
      for (int s = 0; s < SPL; ++s) { in.b[s] = 0.f; in.x[s] = 0.f; }
      if (cur && kvalid) ldv_nc<SPL>(Bp + (long long)t * K, in.b);
      if (kvalid && t + PFL2 < T) asm volatile("prefetch.global.L2 [%0];" ::"l"(Bp + (long long)(t + PFL2) * K));
      in.m = cur ? __ldg(mp + t) : 0.f;
      if (kvalid && ok) ldv_cg<SPL>(Bt + (long long)(t - 1) * K, in.x);      // beta_tilde_{t-1}
      return in;
    };
    auto step2 = [&](int t, const In& in, bool live) {
      float2 v[H2];
      float ab[SPL];
#pragma unroll
      for (int s = 0; s < SPL; ++s) ab[s] = acur[s] * in.x[s];
      float S, G;                                    // G = sum_k a_{t-1}(k) beta_{t-1}(k)
      exchange(acur, lsum(acur), lsum(ab), true, v, S, G);
      if (!live) return;
      float dot[SPL];
      dots(v, Pc, dot);
      const float rG = G > 0.f ? __fdividef(1.f, G) : 0.f;
      float gam[SPL];
#pragma unroll
      for (int s = 0; s < SPL; ++s) gam[s] = ab[s] * rG;
      if (kvalid) stv<SPL>(Ep + (long long)(t - 1) * K, gam);
      const float rr = rcur * rG;                    // xi_{t-2}: vprev is all-zero on the first step
#pragma unroll
      for (int s = 0; s < SPL; ++s) {
        const float w = bprev[s] * in.x[s] * rr;
        const float2 w2 = make_float2(w, w);
#pragma unroll
        for (int j = 0; j < H2; ++j) X[s][j] = ffma2(vprev[j], w2, X[s][j]);
      }
#pragma unroll
      for (int j = 0; j < H2; ++j) vprev[j] = v[j];
      const float inv = S > 0.f ? __fdividef(1.f, S) : 0.f;
      logZ += (double)(__logf(S) + in.m);            // in.m = 0 on the last step (t == T)
#pragma unroll
      for (int s = 0; s < SPL; ++s) { acur[s] = dot[s] * in.b[s] * inv; bprev[s] = in.b[s]; }
      rcur = inv;
    };
    run_ring<PF2, In>(h + 1, T - h, 1, load2, step2);
    if (active && lig == 0 && a.logZ) a.logZ[trial] = logZ;
    __syncthreads();                                 // backward group's xi rows are in xz
    if (kvalid && a.Ezz) {
      float* out = a.Ezz + (long long)trial * K * K + j0;
#pragma unroll
      for (int i = 0; i < KP; ++i)
        if (i < K) {
          float o[SPL];
#pragma unroll
          for (int s = 0; s < SPL; ++s) {
            const float xi = (i & 1) ? X[s][i / 2].y : X[s][i / 2].x;
            const float pi = (i & 1) ? Pc[s][i / 2].y : Pc[s][i / 2].x;
            o[s] = pi * (xi + xz[grp][i][j0 + s]);
          }
          stv<SPL>(out + (long long)i * K, o);
        }
    }
  } else {
    // =============================== backward group (the lane owns rows j0 .. j0+SPL-1 of P)
    float2 Pr[SPL][H2];
#pragma unroll
    for (int s = 0; s < SPL; ++s)
#pragma unroll
      for (int j = 0; j < H2; ++j) Pr[s][j] = make_float2(Pg[(j0 + s) * KP + 2 * j], Pg[(j0 + s) * KP + 2 * j + 1]);
    float beta[SPL];                                 // beta_tilde_{t+1}
#pragma unroll
    for (int s = 0; s < SPL; ++s) beta[s] = 1.f;
    if (T > 0 && kvalid) stv<SPL>(Bt + (long long)(T - 1) * K, beta);
    auto load1 = [&](int t, bool ok) {
      In in;
#pragma unroll
      for (int s = 0; s < SPL; ++s) { in.b[s] = 0.f; in.x[s] = 0.f; }
      if (kvalid && ok) ldv_nc<SPL>(Bp + (long long)(t + 1) * K, in.b);
      if (kvalid && t + 1 >= PFFAR) asm volatile("prefetch.global.L2 [%0];" ::"l"(Bp + (long long)(t + 1 - PFFAR) * K));
      in.m = 0.f;
      return in;
    };
    auto step1 = [&](int t, const In& in, bool live) {
      float2 vk[H2];
      float msg[SPL];
#pragma unroll
      for (int s = 0; s < SPL; ++s) msg[s] = in.b[s] * beta[s];
      float Dn, unused;
      exchange(msg, lsum(msg), 0.f, false, vk, Dn, unused);
      if (!live) return;
      float u[SPL];
      dots(vk, Pr, u);
      const float rho = Dn > 0.f ? __fdividef(1.f, Dn) : 0.f;
#pragma unroll
      for (int s = 0; s < SPL; ++s) beta[s] = u[s] * rho;
      if (kvalid) stv<SPL>(Bt + (long long)t * K, beta);
    };
    run_ring<PF, In>(T - 2, max(T - 1 - h, 0), -1, load1, step1);            // t = T-2 .. h
    __syncthreads();
    // second half: t = h-1 .. 0.  gamma_t / xi_t need G_t = sum_j a_t(j) beta_t(j), which exists only at the END of
    // step t: it rides along with the next step's message, so step t is finalised one iteration later (t = 0 by a
    // trailing exchange).
    float2 X[SPL][H2], vkp[H2];
#pragma unroll
    for (int j = 0; j < H2; ++j) {
      vkp[j] = make_float2(0.f, 0.f);
#pragma unroll
      for (int s = 0; s < SPL; ++s) X[s][j] = make_float2(0.f, 0.f);
    }
    float al_p[SPL], ab_p[SPL];
    float rho_p = 0.f;
#pragma unroll
    for (int s = 0; s < SPL; ++s) { al_p[s] = 0.f; ab_p[s] = 0.f; }
    auto load2 = [&](int t, bool ok) {
      In in;
#pragma unroll
      for (int s = 0; s < SPL; ++s) { in.b[s] = 0.f; in.x[s] = 0.f; }
      if (kvalid && ok) {
        ldv_nc<SPL>(Bp + (long long)(t + 1) * K, in.b);
        ldv_cg<SPL>(Ep + (long long)t * K, in.x);                           // a_t stored by the forward group
        if (t + 1 >= PFL2) asm volatile("prefetch.global.L2 [%0];" ::"l"(Bp + (long long)(t + 1 - PFL2) * K));
      }
      in.m = 0.f;
      return in;
    };
    auto finalise = [&](int tp, float G) {
      const float rG = G > 0.f ? __fdividef(1.f, G) : 0.f;
      float gam[SPL];
#pragma unroll
      for (int s = 0; s < SPL; ++s) gam[s] = ab_p[s] * rG;
      if (kvalid) stv<SPL>(Ep + (long long)tp * K, gam);
      const float rr = rho_p * rG;
#pragma unroll
      for (int s = 0; s < SPL; ++s) {
        const float w = al_p[s] * rr;
        const float2 w2 = make_float2(w, w);
#pragma unroll
        for (int j = 0; j < H2; ++j) X[s][j] = ffma2(vkp[j], w2, X[s][j]);
      }
    };
    auto step2 = [&](int t, const In& in, bool live) {
      float2 vk[H2];
      float msg[SPL];
#pragma unroll
      for (int s = 0; s < SPL; ++s) msg[s] = in.b[s] * beta[s];
      float Dn, G;
      exchange(msg, lsum(msg), lsum(ab_p), true, vk, Dn, G);
      if (!live) return;
      if (t + 1 < h) finalise(t + 1, G);
      float u[SPL];
      dots(vk, Pr, u);
      rho_p = Dn > 0.f ? __fdividef(1.f, Dn) : 0.f;
#pragma unroll
      for (int s = 0; s < SPL; ++s) {
        beta[s] = u[s] * rho_p;
        al_p[s] = in.x[s];
        ab_p[s] = in.x[s] * beta[s];
      }
#pragma unroll
      for (int j = 0; j < H2; ++j) vkp[j] = vk[j];
    };
    run_ring<PF2, In>(h - 1, h, -1, load2, step2);                           // t = h-1 .. 0
    {
      float2 vk[H2];
      float zero[SPL];
#pragma unroll
      for (int s = 0; s < SPL; ++s) zero[s] = 0.f;
      float Dn, G;
      exchange(zero, 0.f, lsum(ab_p), true, vk, Dn, G);
      if (h > 0) finalise(0, G);
    }
#pragma unroll
    for (int s = 0; s < SPL; ++s)
#pragma unroll
      for (int j = 0; j < H2; ++j) {
        xz[grp][j0 + s][2 * j] = X[s][j].x;
        xz[grp][j0 + s][2 * j + 1] = X[s][j].y;
      }
    __syncthreads();
  }
}

// ------------------------------------------------------------------------------------------------
// Viterbi (fp64 max-sum, backward recursion like ssm.messages.viterbi)
// ------------------------------------------------------------------------------------------------
struct VitArgs {
  const unsigned char* blob;
  const double* ll;
  const long long* offsets;
  int n_trials, K;
  unsigned char* args;   // (total_T, K)
  int* z;
};

__device__ __forceinline__ double shfl_d(double v, int src, int width) {
  int lo = __double2loint(v), hi = __double2hiint(v);
  lo = __shfl_sync(0xffffffffu, lo, src, width);
  hi = __shfl_sync(0xffffffffu, hi, src, width);
  return __hiloint2double(hi, lo);
}

template <int KP>
__global__ void __launch_bounds__(128) viterbi_kernel(const VitArgs a) {
  constexpr int GPW = 32 / KP;
  const BlobHeader* hd = reinterpret_cast<const BlobHeader*>(a.blob);
  const double* lPg = reinterpret_cast<const double*>(a.blob + hd->off_logP_d);   // KP x KP, -inf padded
  const double* lpi = reinterpret_cast<const double*>(a.blob + hd->off_logpi0_d);
  const int lane = threadIdx.x & 31;
  const int k = lane % KP;
  const int warp_global = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int trial = warp_global * GPW + lane / KP;
  const bool active = trial < a.n_trials;
  const int K = a.K;
  const bool kvalid = k < K;
  const long long beg = active ? a.offsets[trial] : 0;
  const int T = active ? (int)(a.offsets[trial + 1] - beg) : 0;
  int Tmax = T;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) Tmax = max(Tmax, __shfl_xor_sync(0xffffffffu, Tmax, o));
  if (Tmax == 0) return;
  double lrow[KP];
#pragma unroll
  for (int j = 0; j < KP; ++j) lrow[j] = lPg[k * KP + j];
  const double* llp = a.ll + beg * K + k;
  unsigned char* ap = a.args + beg * K + k;
  double score = 0.0;          // scores[T-1] = 0
  for (int t = Tmax - 2; t >= 0; --t) {
    const bool on = t + 1 < T;
    double s = (on && kvalid) ? score + llp[(long long)(t + 1) * K] : -INFINITY;
    double best = -INFINITY;
    int arg = 0;
#pragma unroll
    for (int j = 0; j < KP; ++j) {
      double v = lrow[j] + shfl_d(s, j, KP);
      if (v > best) { best = v; arg = j; }
    }
    if (on) {
      score = best;
      if (kvalid) ap[(long long)(t + 1) * K] = (unsigned char)arg;
    }
  }
  // z_0 = argmax_j scores_0(j) + log pi0_j + ll_0(j), first index on ties
  double s0 = (T > 0 && kvalid) ? score + lpi[k] + llp[0] : -INFINITY;
  double best = -INFINITY;
  int arg = 0;
#pragma unroll
  for (int j = 0; j < KP; ++j) {
    double v = shfl_d(s0, j, KP);
    if (v > best) { best = v; arg = j; }
  }
  __syncwarp();
  if (active && k == 0 && T > 0) {
    int zt = arg;
    a.z[beg] = zt;
    const unsigned char* ag = a.args + beg * K;
    for (int t = 1; t < T; ++t) {
      zt = ag[(long long)t * K + zt];
      a.z[beg + t] = zt;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// AR sufficient statistics for the M-step (fp64 accumulation)
// ------------------------------------------------------------------------------------------------
struct StatArgs {
  const float* x;
  const long long* offsets;
  const float* Ez;
  int n_trials, K, D, lags, Q;   // Q = D*lags + 1 + D
  double* stats;                 // (K, Q, Q)
  double* counts;                // (K)
};

__global__ void __launch_bounds__(256) ar_stats_kernel(const StatArgs a) {
  // each block owns a strided subset of trials and a private (K, Q, Q) fp64 accumulator in
  // shared memory; one atomic flush per block at the end.
  extern __shared__ __align__(16) unsigned char smem_raw[];
  constexpr int TT = 64;
  const int K = a.K, D = a.D, L = a.lags, Q = a.Q;
  const int PS = Q | 1;
  double* acc = reinterpret_cast<double*>(smem_raw);              // K*Q*Q
  double* cnt = acc + (size_t)K * Q * Q;                          // K
  float* phi = reinterpret_cast<float*>(cnt + ((K + 1) & ~1));    // TT * PS
  float* gam = phi + (size_t)TT * PS;                             // TT * K
  const int tid = threadIdx.x;
  const int items = K * Q * Q;
  for (int i = tid; i < items + K; i += 256) acc[i] = 0.0;
  __syncthreads();
  for (int trial = blockIdx.x; trial < a.n_trials; trial += gridDim.x) {
    const long long beg = a.offsets[trial];
    const int T = (int)(a.offsets[trial + 1] - beg);
    for (int t0 = L; t0 < T; t0 += TT) {
      const int nt = min(TT, T - t0);
      for (int i = tid; i < nt * Q; i += 256) {
        int r = i / Q, q = i - r * Q;
        int t = t0 + r;
        float v;
        if (q < D * L) { int l = q / D, d = q - l * D; v = __ldg(a.x + (beg + t - l - 1) * D + d); }
        else if (q == D * L) v = 1.f;
        else v = __ldg(a.x + (beg + t) * D + (q - D * L - 1));
        phi[r * PS + q] = v;
      }
      for (int i = tid; i < nt * K; i += 256) gam[i] = __ldg(a.Ez + (beg + t0) * K + i);
      __syncthreads();
      for (int it = tid; it < items; it += 256) {
        int k = it / (Q * Q);
        int rem = it - k * Q * Q;
        int p = rem / Q, q = rem - p * Q;
        if (q < p) continue;                       // upper triangle only; mirrored on the host
        float s = 0.f;
        for (int r = 0; r < nt; ++r) s = fmaf(gam[r * K + k] * phi[r * PS + p], phi[r * PS + q], s);
        acc[it] += (double)s;
      }
      if (tid < K) {
        float s = 0.f;
        for (int r = 0; r < nt; ++r) s += gam[r * K + tid];
        cnt[tid] += (double)s;
      }
      __syncthreads();
    }
  }
  for (int it = tid; it < items; it += 256) {
    int rem = it % (Q * Q);
    int p = rem / Q, q = rem - p * Q;
    if (q >= p && acc[it] != 0.0) atomicAdd(a.stats + it, acc[it]);
  }
  if (tid < K && cnt[tid] != 0.0) atomicAdd(a.counts + tid, cnt[tid]);
}

struct WsLayoutH {
  size_t Bsc, mx, beta, ll, args, total;
};
WsLayoutH hmm_ws(int K, long long total_T, int fp64) {
  WsLayoutH w;
  size_t o = 0;
  auto take = [&](size_t b) { size_t r = o; o += (b + 255) & ~(size_t)255; return r; };
  w.Bsc = take((size_t)total_T * K * 4);
  w.mx = take((size_t)total_T * 4);
  w.beta = take((size_t)total_T * K * 4);      // backward messages of the second half of every trial
  w.ll = fp64 ? take((size_t)total_T * K * 8) : 0;
  w.args = fp64 ? take((size_t)total_T * K) : 0;
  w.total = o;
  return w;
}

template <typename real, int DP, int TS>
int launch_emission_t(const EmitArgs<real>& a, int n_trials, int max_T, cudaStream_t st) {
  constexpr int TILE = 128 * TS;
  const int XS = a.D | 1;
  size_t smem = sizeof(real) * ((size_t)a.K * a.J * DP + ((a.K + 1 + 3) & ~3)) +
                4 * ((size_t)(TILE + a.lags) * XS + (sizeof(real) == 4 ? (size_t)TILE * (a.K + 1) : 0));
  if (smem > 227 * 1024) BN_FAIL("arhmm emission: K=%d D=%d lags=%d needs %zu B of shared memory", a.K, a.D, a.lags, smem);
  auto kern = emission_kernel<real, DP, TS>;
  BN_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  // one block per trial walks all of its time tiles (parameters staged once) when there are enough
  // trials to fill the machine; otherwise the time tiles are spread over grid.y as well
  int tiles = bn_cdiv(max_T, TILE);
  int gy = n_trials >= 2 * 148 ? 1 : bn_cdiv(2 * 148, n_trials);
  if (gy > tiles) gy = tiles;
  dim3 grid(n_trials, gy);
  kern<<<grid, 128, smem, st>>>(a);
  BN_LAUNCHED();
  return 0;
}

template <typename real, int TS>
int launch_emission(const EmitArgs<real>& a, int n_trials, int max_T, cudaStream_t st) {
  switch (bn_round_dp(a.D)) {
    case 4: return launch_emission_t<real, 4, TS>(a, n_trials, max_T, st);
    case 8: return launch_emission_t<real, 8, TS>(a, n_trials, max_T, st);
    case 12: return launch_emission_t<real, 12, TS>(a, n_trials, max_T, st);
    case 16: return launch_emission_t<real, 16, TS>(a, n_trials, max_T, st);
    case 20: return launch_emission_t<real, 20, TS>(a, n_trials, max_T, st);
    case 24: return launch_emission_t<real, 24, TS>(a, n_trials, max_T, st);
    case 28: return launch_emission_t<real, 28, TS>(a, n_trials, max_T, st);
    case 32: return launch_emission_t<real, 32, TS>(a, n_trials, max_T, st);
    default: BN_FAIL("arhmm: observation dim D=%d > 32 unsupported", a.D);
  }
}

int check_dims(int K, int D, int lags) {
  if (K < 1 || K > 32) BN_FAIL("arhmm: K=%d outside [1, 32]", K);
  if (D < 1 || D > 32) BN_FAIL("arhmm: D=%d outside [1, 32]", D);
  if (lags < 0 || lags > 8) BN_FAIL("arhmm: lags=%d outside [0, 8]", lags);
  return 0;
}

}  // namespace

extern "C" size_t bn_arhmm_params_bytes(int K, int D, int lags) {
  if (check_dims(K, D, lags)) return 0;
  return (size_t)bn_blob_layout(K, D, lags).total;
}

extern "C" int bn_arhmm_pack_params(int K, int D, int lags, const double* log_pi0, const double* log_Ps,
                                    const double* As, const double* bs, const double* Sigmas,
                                    void* h_blob) {
  BN_TRY(check_dims(K, D, lags));
  if (!log_pi0 || !log_Ps || !As || !bs || !Sigmas || !h_blob) BN_FAIL("bn_arhmm_pack_params: null argument");
  BlobHeader h = bn_blob_layout(K, D, lags);
  unsigned char* blob = (unsigned char*)h_blob;
  memset(blob, 0, (size_t)h.total);
  memcpy(blob, &h, sizeof(h));
  const int KP = h.KP, DP = h.DP, J = h.J;
  float* pi0f = (float*)(blob + h.off_pi0_f);
  float* Pf = (float*)(blob + h.off_P_f);
  float* Wf = (float*)(blob + h.off_W_f);
  float* cf = (float*)(blob + h.off_c_f);
  double* lpid = (double*)(blob + h.off_logpi0_d);
  double* lPd = (double*)(blob + h.off_logP_d);
  double* Wd = (double*)(blob + h.off_W_d);
  double* cd = (double*)(blob + h.off_c_d);
  for (int j = 0; j < KP; ++j) {
    lpid[j] = j < K ? log_pi0[j] : -INFINITY;
    pi0f[j] = j < K ? (float)exp(log_pi0[j]) : 0.f;
    for (int k = 0; k < KP; ++k) {
      bool ok = j < K && k < K;
      lPd[j * KP + k] = ok ? log_Ps[j * K + k] : -INFINITY;
      Pf[j * KP + k] = ok ? (float)exp(log_Ps[j * K + k]) : 0.f;
    }
  }
  std::vector<double> Lc(D * D), Li(D * D), M(D * (size_t)J);
  for (int k = 0; k < K; ++k) {
    const double* S = Sigmas + (size_t)k * D * D;
    // Cholesky S = Lc Lc^T
    std::fill(Lc.begin(), Lc.end(), 0.0);
    for (int i = 0; i < D; ++i)
      for (int j = 0; j <= i; ++j) {
        double s = S[i * D + j];
        for (int q = 0; q < j; ++q) s -= Lc[i * D + q] * Lc[j * D + q];
        if (i == j) {
          if (!(s > 0.0)) BN_FAIL("arhmm: Sigma[%d] is not positive definite", k);
          Lc[i * D + i] = sqrt(s);
        } else {
          Lc[i * D + j] = s / Lc[j * D + j];
        }
      }
    // Li = Lc^-1 (lower triangular)
    std::fill(Li.begin(), Li.end(), 0.0);
    for (int c = 0; c < D; ++c)
      for (int i = c; i < D; ++i) {
        double s = (i == c) ? 1.0 : 0.0;
        for (int q = c; q < i; ++q) s -= Lc[i * D + q] * Li[q * D + c];
        Li[i * D + c] = s / Lc[i * D + i];
      }
    // M = [I, -A_k, -b_k] (D x J);  W = Li M
    const double* A = As + (size_t)k * D * D * lags;
    const double* b = bs + (size_t)k * D;
    std::fill(M.begin(), M.end(), 0.0);
    for (int i = 0; i < D; ++i) {
      M[(size_t)i * J + i] = 1.0;
      for (int c = 0; c < D * lags; ++c) M[(size_t)i * J + D + c] = -A[(size_t)i * D * lags + c];
      M[(size_t)i * J + J - 1] = -b[i];
    }
    double logdet = 0.0;
    for (int i = 0; i < D; ++i) logdet += log(Lc[i * D + i]);
    cd[k] = -0.5 * D * LN2PI - logdet;
    cf[k] = (float)cd[k];
    for (int j = 0; j < J; ++j)
      for (int i = 0; i < D; ++i) {
        double s = 0.0;
        for (int q = 0; q <= i; ++q) s += Li[i * D + q] * M[(size_t)q * J + j];
        Wd[((size_t)k * J + j) * DP + i] = s;
        Wf[((size_t)k * J + j) * DP + i] = (float)s;
      }
  }
  cd[K] = -0.5 * D * LN2PI;
  cf[K] = (float)cd[K];
  // TF32 hi / lo split of W for the tensor-core emission kernel: W = hi + lo to ~2^-22 relative, both
  // exactly representable in TF32.  Row n = k*DP + i, column kt = b*DP + d for x_{t-b}[d] and
  // (lags+1)*DP for the bias; stored as 8x(16-byte) core matrices, K-adjacent cores NT/8*128 bytes apart.
  {
    float* Whi = (float*)(blob + h.off_W_hi);
    float* Wlo = (float*)(blob + h.off_W_lo);
    auto tf32 = [](double v) {
      float f = (float)v;
      uint32_t u;
      memcpy(&u, &f, 4);
      u = (u + 0x1000u) & 0xFFFFE000u;        // round to nearest (ties away), 10 explicit mantissa bits
      memcpy(&f, &u, 4);
      return f;
    };
    for (int k = 0; k < K; ++k)
      for (int j = 0; j < J; ++j) {
        int kt;
        if (j == J - 1) kt = (lags + 1) * DP;
        else { int b = j / D, d = j - b * D; kt = b * DP + d; }
        for (int i = 0; i < D; ++i) {
          const double w = Wd[((size_t)k * J + j) * DP + i];
          const int n = k * DP + i;
          const size_t o = (size_t)(kt / 4) * (h.NT / 8 * 32) + (size_t)(n / 8) * 32 + (n % 8) * 4 + kt % 4;
          const float hi = tf32(w);
          Whi[o] = hi;
          Wlo[o] = tf32(w - (double)hi);
        }
      }
  }
  return 0;
}

extern "C" size_t bn_arhmm_workspace_bytes(int K, int D, int lags, int64_t total_T, int n_trials, int fp64) {
  (void)D; (void)lags; (void)n_trials;
  return hmm_ws(K, total_T, fp64).total;
}

template <int KP>
static int launch_scan(const ScanArgs& a, cudaStream_t st) {
  constexpr int GPW = 32 / KP;
  int warps = bn_cdiv(a.n_trials, GPW);
  scan_kernel<KP><<<bn_cdiv(warps, 4), 128, 0, st>>>(a);
  BN_LAUNCHED();
  return 0;
}

template <int KP>
static int launch_scan2(const Scan2Args& a, cudaStream_t st) {
  constexpr int GPW = 32 / KP;
  if (a.Ez) {
    scan2_kernel<KP, true><<<bn_cdiv(a.n_trials, 2 * GPW), 128, 0, st>>>(a);
  } else {
    scan2_kernel<KP, false><<<bn_cdiv(a.n_trials, 4 * GPW), 128, 0, st>>>(a);
  }
  BN_LAUNCHED();
  return 0;
}

// SPL-states-per-lane scan: K = 8 .. 32 with K % SPL == 0 (BN_SCAN=2 keeps the lane-per-state kernel)
template <int KP, int SPL>
static int launch_scan4(const Scan2Args& a, cudaStream_t st) {
  constexpr int GPW = 32 / (KP / SPL);
  if (a.Ez) {
    scan4_kernel<KP, SPL, true><<<bn_cdiv(a.n_trials, GPW), 64, 0, st>>>(a);
  } else {
    scan4_kernel<KP, SPL, false><<<bn_cdiv(a.n_trials, 2 * GPW), 64, 0, st>>>(a);
  }
  BN_LAUNCHED();
  return 0;
}

template <int KP>
static int launch_vit(const VitArgs& a, cudaStream_t st) {
  constexpr int GPW = 32 / KP;
  int warps = bn_cdiv(a.n_trials, GPW);
  viterbi_kernel<KP><<<bn_cdiv(warps, 4), 128, 0, st>>>(a);
  BN_LAUNCHED();
  return 0;
}

// side stream + events of the emission / scan overlap inside bn_arhmm_estep (per host thread and device)
struct OverlapStreams {
  int device;
  cudaStream_t side;
  cudaEvent_t start, done[8];
};
static OverlapStreams* overlap_streams() {
  thread_local OverlapStreams cache[16];
  thread_local int n_cached = 0;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return nullptr;
  for (int i = 0; i < n_cached; ++i)
    if (cache[i].device == dev) return &cache[i];
  if (n_cached >= 16) return nullptr;
  OverlapStreams o;
  o.device = dev;
  if (cudaStreamCreateWithFlags(&o.side, cudaStreamNonBlocking) != cudaSuccess) return nullptr;
  if (cudaEventCreateWithFlags(&o.start, cudaEventDisableTiming) != cudaSuccess) return nullptr;
  for (int i = 0; i < 8; ++i)
    if (cudaEventCreateWithFlags(&o.done[i], cudaEventDisableTiming) != cudaSuccess) return nullptr;
  cache[n_cached] = o;
  return &cache[n_cached++];
}

extern "C" int bn_arhmm_estep(int K, int D, int lags, const void* d_blob, const float* d_x,
                              const int64_t* d_offsets, int n_trials, int64_t total_T, int max_T,
                              void* d_ws, float* d_Ez, float* d_Ezz, double* d_logZ, void* stream) {
  BN_TRY(check_dims(K, D, lags));
  if (!d_blob || !d_x || !d_offsets || !d_ws) BN_FAIL("bn_arhmm_estep: null argument");
  if (d_Ezz && !d_Ez) BN_FAIL("bn_arhmm_estep: d_Ezz requires d_Ez (alpha_hat is staged there)");
  if (n_trials <= 0 || total_T <= 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  WsLayoutH w = hmm_ws(K, total_T, 0);
  unsigned char* ws = (unsigned char*)d_ws;
  EmitArgs<float> e;
  e.blob = (const unsigned char*)d_blob; e.x = d_x; e.offsets = (const long long*)d_offsets;
  e.K = K; e.D = D; e.lags = lags; e.J = D * (lags + 1) + 1;
  e.Bsc = (float*)(ws + w.Bsc); e.mx = (float*)(ws + w.mx); e.ll = nullptr;
  static const bool legacy_scan = [] { const char* e = getenv("BN_SCAN"); return e && e[0] == '1'; }();
  // Optional emission / scan overlap (BN_ESTEP_GROUPS=g > 1): the call is cut into g trial groups, the emission
  // kernels of all groups run back to back on a side stream and the scan of group i starts on the caller's
  // stream as soon as ITS likelihoods are written.  MEASURED SLOWER on B200 (C4: 555 / 772 / 979 / 1571 us at
  // g = 1 / 2 / 4 / 8, profiles/r02_estep_groups.txt): the scan is a T-step dependent chain whose duration
  // hardly depends on how many trials are in flight, so g scans cost ~g chains while the emission kernels they
  // were meant to hide take less than one.  Default off; the switch stays for the record.
  static const int max_groups = [] { const char* e = getenv("BN_ESTEP_GROUPS"); return e ? atoi(e) : 1; }();
  int groups = 1;
  if (!legacy_scan && max_groups > 1 && n_trials >= 64 * max_groups) groups = max_groups > 8 ? 8 : max_groups;
  OverlapStreams* ov = groups > 1 ? overlap_streams() : nullptr;
  if (groups > 1 && !ov) groups = 1;
  if (groups > 1) {
    BN_CUDA(cudaEventRecord(ov->start, st));
    BN_CUDA(cudaStreamWaitEvent(ov->side, ov->start, 0));        // inputs (and the workspace's last readers) are done
  }
  const int per = (n_trials + groups - 1) / groups;
  for (int g = 0; g < groups; ++g) {
    const int g0 = g * per, gn = std::min(per, n_trials - g0);
    if (gn <= 0) break;
    cudaStream_t est = groups > 1 ? ov->side : st;
    EmitArgs<float> eg = e;
    eg.offsets = e.offsets + g0;
    int tc = 1;
    if (bn_get_tensor_core_mode())
      tc = bn_launch_emission_tc(eg.blob, d_x, eg.offsets, K, D, lags, gn, max_T, eg.Bsc, eg.mx, est);
    if (tc < 0) return tc;
    if (tc > 0) BN_TRY((launch_emission<float, 2>(eg, gn, max_T, est)));
    if (groups > 1) {
      BN_CUDA(cudaEventRecord(ov->done[g], ov->side));
      BN_CUDA(cudaStreamWaitEvent(st, ov->done[g], 0));
    }
    if (legacy_scan) {
      // sequential forward-then-backward kernel (kept for A/B timing: BN_SCAN=1)
      ScanArgs s;
      s.blob = e.blob; s.Bsc = e.Bsc; s.mx = e.mx; s.offsets = e.offsets; s.n_trials = n_trials; s.K = K;
      s.Ez = d_Ez; s.Ezz = d_Ezz; s.logZ = d_logZ; s.cinv = (float*)(ws + w.beta);
      switch (bn_round_kp(K)) {
        case 2: return launch_scan<2>(s, st);
        case 4: return launch_scan<4>(s, st);
        case 8: return launch_scan<8>(s, st);
        case 16: return launch_scan<16>(s, st);
        default: return launch_scan<32>(s, st);
      }
    }
    Scan2Args s;
    s.blob = e.blob; s.Bsc = e.Bsc; s.mx = e.mx; s.offsets = eg.offsets; s.n_trials = gn; s.K = K;
    s.Ez = d_Ez; s.Ezz = d_Ezz ? d_Ezz + (size_t)g0 * K * K : nullptr; s.logZ = d_logZ ? d_logZ + g0 : nullptr;
    s.beta = (float*)(ws + w.beta);
    int r;
    // Measured on B200 (C4, profiles/r02_scan4.txt): the SPL-per-lane kernel executes 93 M warp instructions against
    // scan2's 173 M but takes 484 us against 314 us -- with 0.86 warps per scheduler nothing hides the ~5 cycles per
    // dependent instruction and the global-load latency, so issue utilisation drops from 56 % to 20 %.  The lane-per-
    // state kernel stays the default; BN_SCAN=4 selects this one.
    static const bool lane_per_state = [] { const char* e = getenv("BN_SCAN"); return !(e && e[0] == '4'); }();
    const int kp = bn_round_kp(K);
    static const int spl16 = [] { const char* e = getenv("BN_SCAN4_SPL"); return e ? atoi(e) : 4; }();
    if (!lane_per_state && kp == 16 && K % 4 == 0 && spl16 == 4) {
      r = launch_scan4<16, 4>(s, st);
    } else if (!lane_per_state && kp == 16 && K % 2 == 0) {
      r = launch_scan4<16, 2>(s, st);
    } else if (!lane_per_state && kp == 32 && K % 2 == 0) {
      r = launch_scan4<32, 2>(s, st);
    } else if (!lane_per_state && kp == 8 && K % 2 == 0) {
      r = launch_scan4<8, 2>(s, st);
    } else
    switch (kp) {
      case 2: r = launch_scan2<2>(s, st); break;
      case 4: r = launch_scan2<4>(s, st); break;
      case 8: r = launch_scan2<8>(s, st); break;
      case 16: r = launch_scan2<16>(s, st); break;
      default: r = launch_scan2<32>(s, st); break;
    }
    if (r) return r;
  }
  return 0;
}

extern "C" int bn_arhmm_viterbi(int K, int D, int lags, const void* d_blob, const float* d_x,
                                const int64_t* d_offsets, int n_trials, int64_t total_T, int max_T,
                                void* d_ws, int32_t* d_z, void* stream) {
  BN_TRY(check_dims(K, D, lags));
  if (!d_blob || !d_x || !d_offsets || !d_ws || !d_z) BN_FAIL("bn_arhmm_viterbi: null argument");
  if (n_trials <= 0 || total_T <= 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  WsLayoutH w = hmm_ws(K, total_T, 1);
  unsigned char* ws = (unsigned char*)d_ws;
  EmitArgs<double> e;
  e.blob = (const unsigned char*)d_blob; e.x = d_x; e.offsets = (const long long*)d_offsets;
  e.K = K; e.D = D; e.lags = lags; e.J = D * (lags + 1) + 1;
  e.Bsc = nullptr; e.mx = nullptr; e.ll = (double*)(ws + w.ll);
  BN_TRY((launch_emission<double, 1>(e, n_trials, max_T, st)));
  VitArgs v;
  v.blob = e.blob; v.ll = e.ll; v.offsets = e.offsets; v.n_trials = n_trials; v.K = K;
  v.args = ws + w.args; v.z = d_z;
  switch (bn_round_kp(K)) {
    case 2: return launch_vit<2>(v, st);
    case 4: return launch_vit<4>(v, st);
    case 8: return launch_vit<8>(v, st);
    case 16: return launch_vit<16>(v, st);
    default: return launch_vit<32>(v, st);
  }
}

extern "C" int bn_arhmm_ar_stats(int K, int D, int lags, const float* d_x, const int64_t* d_offsets,
                                 int n_trials, int64_t total_T, const float* d_Ez, double* d_stats,
                                 double* d_counts, void* stream) {
  BN_TRY(check_dims(K, D, lags));
  if (!d_x || !d_offsets || !d_Ez || !d_stats || !d_counts) BN_FAIL("bn_arhmm_ar_stats: null argument");
  if (n_trials <= 0 || total_T <= 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  StatArgs a;
  a.x = d_x; a.offsets = (const long long*)d_offsets; a.Ez = d_Ez; a.n_trials = n_trials;
  a.K = K; a.D = D; a.lags = lags; a.Q = D * lags + 1 + D; a.stats = d_stats; a.counts = d_counts;
  const int PS = a.Q | 1;
  size_t smem = 8 * ((size_t)K * a.Q * a.Q + ((K + 1) & ~1)) + 4 * ((size_t)64 * PS + (size_t)64 * K);
  if (smem > 227 * 1024) BN_FAIL("arhmm ar_stats: K=%d D=%d lags=%d needs %zu B of shared memory", K, D, lags, smem);
  BN_CUDA(cudaFuncSetAttribute(ar_stats_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int grid = n_trials < 296 ? n_trials : 296;
  ar_stats_kernel<<<grid, 256, smem, st>>>(a);
  BN_LAUNCHED();
  return 0;
}


// ------------------------------------------------------------------------------------------------
// Host-side staging helper: the reference hands ssm a python list of per-trial (T_i, D) arrays
// (arhmm_grid_search.py:170); concatenating ~10^3 of them into the pinned staging buffer with one
// thread costs more than the E-step they feed.  Rows are copied (fp64 sources converted) by a few
// threads, each taking a contiguous share of the bytes.
// ------------------------------------------------------------------------------------------------
#include <thread>

extern "C" int bn_host_gather_rows(const void* const* h_src, const int64_t* h_rows, const int32_t* h_is_f64,
                                   int n, int D, float* h_dst, int threads) {
  if (n < 0 || D <= 0 || (n > 0 && (!h_src || !h_rows || !h_dst))) BN_FAIL("bn_host_gather_rows: bad argument");
  std::vector<int64_t> off(n + 1, 0);
  for (int i = 0; i < n; ++i) {
    if (h_rows[i] < 0 || (h_rows[i] > 0 && !h_src[i])) BN_FAIL("bn_host_gather_rows: bad trial %d", i);
    off[i + 1] = off[i] + h_rows[i] * D;
  }
  const int64_t total = off[n];
  if (total == 0) return 0;
  int nt = threads < 1 ? 1 : (threads > 64 ? 64 : threads);
  if (total < (int64_t)1 << 18) nt = 1;
  auto work = [&](int64_t lo, int64_t hi) {          // element range [lo, hi) of the destination
    int i = (int)(std::upper_bound(off.begin(), off.end(), lo) - off.begin()) - 1;
    for (; i < n && off[i] < hi; ++i) {
      const int64_t a = lo > off[i] ? lo : off[i], b = hi < off[i + 1] ? hi : off[i + 1];
      if (b <= a) continue;
      if (h_is_f64 && h_is_f64[i]) {
        const double* s = (const double*)h_src[i] + (a - off[i]);
        for (int64_t k = 0; k < b - a; ++k) h_dst[a + k] = (float)s[k];
      } else {
        memcpy(h_dst + a, (const float*)h_src[i] + (a - off[i]), (size_t)(b - a) * sizeof(float));
      }
    }
  };
  if (nt == 1) {
    work(0, total);
    return 0;
  }
  std::vector<std::thread> pool;
  const int64_t per = (total + nt - 1) / nt;
  for (int t = 0; t < nt; ++t) {
    const int64_t lo = t * per, hi = lo + per < total ? lo + per : total;
    if (lo < hi) pool.emplace_back(work, lo, hi);
  }
  for (auto& th : pool) th.join();
  return 0;
}
