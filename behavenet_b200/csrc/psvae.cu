// PS-VAE latent block: subspace projections, reparameterisation, label / KL / decomposed-KL
// loss terms and their gradients, for one reference chunk (<= 200 frames in every caller).
//
// Reference: models/vaes.py:571-601 (forward), 669-696 (loss terms), 17-35 (reparameterize);
// fitting/losses.py:62-96 (gaussian_ll), 130-147 (kl_div_to_std_normal), 284-372 (decomposed_kl).
//
// The tensors are tiny (n x 16); the kernels favour determinism and simplicity over speed: every
// output element is produced by exactly one thread, in a fixed summation order.
#include <math.h>

#include "../../include/behavenet_b200.h"
#include "bn_common.cuh"

namespace {

constexpr float LN2PI_F = 1.8378770664093453f;

struct LatArgs {
  int n, L, nl;
  const float *pre, *logvar, *A, *B, *Dw, *Db, *eps, *labels, *lmask;
  float alpha, beta, klw, kls;
  float *mu, *z, *yhat;
  double* terms;
  float *gmu, *glv, *gz, *gDw, *gDb;
  float* log_qz;   // (n)
  float* lse;      // (n, du)
  float* P;        // (n, n): exp(joint(j, i) - max_i joint(j, i)), written by dkl_rows_kernel
  float* rse;      // (n): 1 / sum_i P[j][i], so that P[j][i] * rse[j] = q(z_j | x_i) / (n q(z_j))
};

// K0: projections + reparameterisation + supervised terms.  One thread per (frame, latent dim).
__global__ void latent_fwd_kernel(const LatArgs a) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= a.n * a.L) return;
  int f = i / a.L, d = i - f * a.L;
  const float* row = d < a.nl ? a.A + (size_t)d * a.L : a.B + (size_t)(d - a.nl) * a.L;
  const float* p = a.pre + (size_t)f * a.L;
  float m = 0.f;
  for (int c = 0; c < a.L; ++c) m = fmaf(__ldg(row + c), __ldg(p + c), m);
  float lv = a.logvar[i];
  float z = a.eps ? fmaf(a.eps[i], expf(lv), m) : m;
  a.mu[i] = m;
  a.z[i] = z;
  if (d < a.nl) {
    float yh = fmaf(m, a.Dw[d], a.Db[d]);
    a.yhat[(size_t)f * a.nl + d] = yh;
    float gy = 0.f;
    if (a.labels) {
      float diff = yh - a.labels[(size_t)f * a.nl + d];
      float mk = a.lmask ? a.lmask[(size_t)f * a.nl + d] : 1.f;
      atomicAdd(a.terms + 0, (double)(diff * diff * mk));
      gy = a.alpha * diff * mk / (float)a.n;           // d(-alpha * label_ll)/d yhat
      if (a.gDw) atomicAdd(a.gDw + d, gy * m);
      if (a.gDb) atomicAdd(a.gDb + d, gy);
    }
    // supervised-latent KL to N(0,1): 0.5 * (exp(lv) - lv + mu^2 - 1)
    atomicAdd(a.terms + 1, (double)(0.5f * (expf(lv) - lv + m * m - 1.f)));
    if (a.gmu) {
      a.gmu[i] = gy * a.Dw[d] + a.kls * m / (float)a.n;
      a.glv[i] = a.kls * 0.5f * (expf(lv) - 1.f) / (float)a.n;
      a.gz[i] = 0.f;
    }
  }
}

__device__ __forceinline__ float lq_elem(float z, float mu, float lv) {
  float d = z - mu;
  return -0.5f * (expf(-lv) * d * d + lv + LN2PI_F);
}

// K1: one block per sample j: log q(z_j) and per-dimension log-sum-exps over the chunk.
__global__ void __launch_bounds__(128) dkl_rows_kernel(const LatArgs a) {
  __shared__ float red[128];
  __shared__ float zj[64];
  const int j = blockIdx.x, tid = threadIdx.x;
  const int du = a.L - a.nl;
  if (tid < du) zj[tid] = a.z[(size_t)j * a.L + a.nl + tid];
  __syncthreads();
  auto block_max = [&](float v) {
    red[tid] = v;
    __syncthreads();
    for (int o = 64; o > 0; o >>= 1) { if (tid < o) red[tid] = fmaxf(red[tid], red[tid + o]); __syncthreads(); }
    float r = red[0];
    __syncthreads();
    return r;
  };
  auto block_sum = [&](float v) {
    red[tid] = v;
    __syncthreads();
    for (int o = 64; o > 0; o >>= 1) { if (tid < o) red[tid] += red[tid + o]; __syncthreads(); }
    float r = red[0];
    __syncthreads();
    return r;
  };
  // joint over dims
  float mx = -INFINITY;
  for (int i = tid; i < a.n; i += 128) {
    float s = 0.f;
    for (int l = 0; l < du; ++l)
      s += lq_elem(zj[l], a.mu[(size_t)i * a.L + a.nl + l], a.logvar[(size_t)i * a.L + a.nl + l]);
    mx = fmaxf(mx, s);
  }
  mx = block_max(mx);
  float se = 0.f, diag = 0.f;
  for (int i = tid; i < a.n; i += 128) {
    float s = 0.f;
    for (int l = 0; l < du; ++l)
      s += lq_elem(zj[l], a.mu[(size_t)i * a.L + a.nl + l], a.logvar[(size_t)i * a.L + a.nl + l]);
    const float e = expf(s - mx);
    if (a.P) a.P[(size_t)j * a.n + i] = e;
    se += e;
    if (i == j) diag = s;
  }
  se = block_sum(se);
  diag = block_sum(diag);
  const float log_qz = mx + logf(se);
  float prod = 0.f;
  for (int l = 0; l < du; ++l) {
    float m2 = -INFINITY;
    for (int i = tid; i < a.n; i += 128)
      m2 = fmaxf(m2, lq_elem(zj[l], a.mu[(size_t)i * a.L + a.nl + l], a.logvar[(size_t)i * a.L + a.nl + l]));
    m2 = block_max(m2);
    float s2 = 0.f;
    for (int i = tid; i < a.n; i += 128)
      s2 += expf(lq_elem(zj[l], a.mu[(size_t)i * a.L + a.nl + l], a.logvar[(size_t)i * a.L + a.nl + l]) - m2);
    s2 = block_sum(s2);
    float v = m2 + logf(s2);
    if (tid == 0) a.lse[(size_t)j * du + l] = v;
    prod += v;
  }
  if (tid == 0) {
    a.log_qz[j] = log_qz;
    if (a.rse) a.rse[j] = 1.f / se;
    float lpz = 0.f;
    for (int l = 0; l < du; ++l) lpz += -0.5f * (zj[l] * zj[l] + LN2PI_F);
    atomicAdd(a.terms + 2, (double)(diag - log_qz));
    atomicAdd(a.terms + 3, (double)(log_qz - prod));
    atomicAdd(a.terms + 4, (double)(prod - lpz));
  }
}

// K2: gradients.  role 0: thread per (j, l) -> gz ; role 1: thread per (i, l) -> gmu, glv.
// Coefficient on lq[j,i,l]:  G = (a * [i == j] + c * p_ji - c * q_jil) / n,
//   a = kl_w, c = beta - kl_w, p_ji = softmax_i(joint[j, :]), q_jil = softmax_i(lq[j, :, l]).
// K2: gradients of the decomposed KL.  One WARP per (sample `me`, unsupervised dim l, role): the lanes stride over
// the other sample of the pair and the partial sums meet in a fixed shuffle tree (deterministic).  The joint
// densities come from the matrix dkl_rows_kernel stored -- the first version recomputed the du-term joint in every
// thread for every pair, one thread per (me, l) walking all n partners serially: 176 us per 200-frame chunk,
// 10 % of a PS-VAE training step.
//   role 0: me = j (the sample z_j), sums over i -> d/dz_j       role 1: me = i, sums over j -> d/dmu_i, d/dlogvar_i
__global__ void __launch_bounds__(128) dkl_grad_kernel(const LatArgs a) {
  const int du = a.L - a.nl;
  const int wg = (blockIdx.x * 128 + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (wg >= a.n * du) return;
  const int role = blockIdx.y;
  const int me = wg / du, l = wg - me * du;
  const float ca = a.klw, cc = a.beta - a.klw, invn = 1.f / (float)a.n;
  const size_t o = (size_t)me * a.L + a.nl + l;
  float g0 = 0.f, g1 = 0.f;
  if (role == 0) {
    const float zl = a.z[o], lse = a.lse[(size_t)me * du + l], rs = a.rse[me];
    const float* Pr = a.P + (size_t)me * a.n;
    for (int i = lane; i < a.n; i += 32) {
      const size_t oi = (size_t)i * a.L + a.nl + l;
      const float mu = a.mu[oi], lv = a.logvar[oi];
      const float p = Pr[i] * rs;
      const float qq = expf(lq_elem(zl, mu, lv) - lse);
      const float G = ((i == me ? ca : 0.f) + cc * p - cc * qq) * invn;
      g0 = fmaf(G, -expf(-lv) * (zl - mu), g0);
    }
  } else {
    const float mu = a.mu[o], lv = a.logvar[o], w = expf(-lv);
    for (int j = lane; j < a.n; j += 32) {
      const float zl = a.z[(size_t)j * a.L + a.nl + l];
      const float p = a.P[(size_t)j * a.n + me] * a.rse[j];
      const float qq = expf(lq_elem(zl, mu, lv) - a.lse[(size_t)j * du + l]);
      const float G = ((j == me ? ca : 0.f) + cc * p - cc * qq) * invn;
      const float dlt = zl - mu;
      g0 = fmaf(G, w * dlt, g0);
      g1 = fmaf(G, 0.5f * (w * dlt * dlt - 1.f), g1);
    }
  }
#pragma unroll
  for (int s = 16; s > 0; s >>= 1) {
    g0 += __shfl_xor_sync(0xffffffffu, g0, s);
    g1 += __shfl_xor_sync(0xffffffffu, g1, s);
  }
  if (lane == 0) {
    if (role == 0) {
      a.gz[o] = g0 + ca * invn * a.z[o];      // - a * log p(z): d/dz = + a z
    } else {
      a.gmu[o] = g0;
      a.glv[o] = g1;
    }
  }
}

__global__ void latent_bwd_kernel(int n, int L, int nl, const float* A, const float* B,
                                  const float* eps, const float* logvar, const float* gz_dec,
                                  const float* gmu_p, const float* glv_p, const float* gz_p,
                                  float* gpre, float* glogvar) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n * L) return;
  int f = i / L, c = i - f * L;
  // glogvar for (f, c)
  {
    float gz = (gz_p ? gz_p[i] : 0.f) + (gz_dec ? gz_dec[i] : 0.f);
    float e = eps ? eps[i] * expf(logvar[i]) : 0.f;
    glogvar[i] = (glv_p ? glv_p[i] : 0.f) + gz * e;
  }
  // gpre[f, c] = sum_d M[d, c] * (gmu_part[f, d] + gz[f, d])
  float s = 0.f;
  for (int d = 0; d < L; ++d) {
    size_t o = (size_t)f * L + d;
    float g = (gmu_p ? gmu_p[o] : 0.f) + (gz_p ? gz_p[o] : 0.f) + (gz_dec ? gz_dec[o] : 0.f);
    float m = d < nl ? A[(size_t)d * L + c] : B[(size_t)(d - nl) * L + c];
    s = fmaf(m, g, s);
  }
  gpre[i] = s;
}

}  // namespace

extern "C" size_t bn_psvae_latent_workspace_bytes(int n, int n_latents) {
  if (n <= 0 || n_latents <= 0) return 0;
  return (size_t)n * ((size_t)n_latents + 2 + (size_t)n) * sizeof(float);      // log q(z), per-dim lse, 1 / sum, pair matrix
}

extern "C" int bn_psvae_latent(int n, int L, int nl, const float* d_pre, const float* d_logvar,
                               const float* d_A, const float* d_B, const float* d_Dw, const float* d_Db,
                               const float* d_eps, const float* d_labels, const float* d_labels_mask,
                               float alpha, float beta, float kl_w, float kl_s_w, void* d_ws, float* d_mu,
                               float* d_z,
                               float* d_yhat, double* d_terms, float* d_gmu_part, float* d_glogvar_part,
                               float* d_gz_part, float* d_gDw, float* d_gDb, void* stream) {
  if (n <= 0) return 0;
  if (L < 1 || nl < 0 || nl > L) BN_FAIL("bn_psvae_latent: n_latents=%d n_labels=%d", L, nl);
  if (L - nl > 64) BN_FAIL("bn_psvae_latent: more than 64 unsupervised latents");
  if (!d_pre || !d_logvar || (nl > 0 && (!d_A || !d_Dw || !d_Db || !d_yhat)) || (!d_B && L > nl) || !d_mu || !d_z ||
      !d_terms)
    BN_FAIL("bn_psvae_latent: null argument");
  const bool grads = d_gmu_part && d_glogvar_part && d_gz_part;
  if (grads && !d_ws) BN_FAIL("bn_psvae_latent: workspace required");
  cudaStream_t st = (cudaStream_t)stream;
  LatArgs a;
  a.n = n; a.L = L; a.nl = nl; a.pre = d_pre; a.logvar = d_logvar; a.A = d_A; a.B = d_B; a.Dw = d_Dw;
  a.Db = d_Db; a.eps = d_eps; a.labels = d_labels; a.lmask = d_labels_mask; a.alpha = alpha;
  a.beta = beta; a.klw = kl_w; a.kls = kl_s_w; a.mu = d_mu; a.z = d_z; a.yhat = d_yhat; a.terms = d_terms;
  a.gmu = grads ? d_gmu_part : nullptr; a.glv = d_glogvar_part; a.gz = d_gz_part; a.gDw = d_gDw; a.gDb = d_gDb;
  a.log_qz = (float*)d_ws;
  a.lse = d_ws ? (float*)d_ws + n : nullptr;
  a.rse = (d_ws && grads) ? (float*)d_ws + (size_t)n * (L + 1) : nullptr;
  a.P = (d_ws && grads) ? (float*)d_ws + (size_t)n * (L + 2) : nullptr;
  latent_fwd_kernel<<<bn_cdiv((long long)n * L, 128), 128, 0, st>>>(a);
  BN_LAUNCHED();
  const int du = L - nl;
  if (du > 0 && d_ws) {
    dkl_rows_kernel<<<n, 128, 0, st>>>(a);
    BN_LAUNCHED();
    if (grads) {
      dkl_grad_kernel<<<dim3(bn_cdiv((long long)n * du * 32, 128), 2), 128, 0, st>>>(a);
      BN_LAUNCHED();
    }
  }
  return 0;
}

extern "C" int bn_psvae_latent_bwd(int n, int L, int nl, const float* d_A, const float* d_B,
                                   const float* d_eps, const float* d_logvar, const float* d_gz_dec,
                                   const float* d_gmu_part, const float* d_glogvar_part,
                                   const float* d_gz_part, float* d_gpre, float* d_glogvar, void* stream) {
  if (n <= 0) return 0;
  if ((!d_A && nl > 0) || (!d_B && L > nl) || !d_logvar || !d_gpre || !d_glogvar)
    BN_FAIL("bn_psvae_latent_bwd: null argument");
  latent_bwd_kernel<<<bn_cdiv((long long)n * L, 128), 128, 0, (cudaStream_t)stream>>>(
      n, L, nl, d_A, d_B, d_eps, d_logvar, d_gz_dec, d_gmu_part, d_glogvar_part, d_gz_part, d_gpre, d_glogvar);
  BN_LAUNCHED();
  return 0;
}
