// Shared between arhmm.cu (CUDA-core kernels, packing, C ABI) and arhmm_tc.cu (tcgen05 emission kernel).
#pragma once
#include <string.h>

#include "bn_common.cuh"

struct BlobHeader {
  int K, D, lags, DP, J, KP;
  int NT, KT;             // tensor-core operand: NT = K*DP rounded up to 16 rows, KT = (lags+1)*DP + 4 rounded up to 8
  // offsets in bytes from blob start
  long long off_pi0_f, off_P_f, off_W_f, off_c_f;
  long long off_logpi0_d, off_logP_d, off_W_d, off_c_d;
  long long off_W_hi, off_W_lo;      // TF32 hi / lo split of W in the UMMA no-swizzle K-major layout
  long long total;
};

static inline int bn_round_dp(int D) { return (D + 3) & ~3; }
static inline int bn_round_kp(int K) { int kp = 2; while (kp < K) kp <<= 1; return kp; }

static inline BlobHeader bn_blob_layout(int K, int D, int lags) {
  BlobHeader h;
  memset(&h, 0, sizeof(h));
  h.K = K; h.D = D; h.lags = lags; h.DP = bn_round_dp(D); h.J = D * (lags + 1) + 1; h.KP = bn_round_kp(K);
  h.NT = (K * h.DP + 15) & ~15;
  h.KT = ((lags + 1) * h.DP + 4 + 7) & ~7;
  long long o = 256;
  auto take = [&](long long bytes) { long long r = o; o += (bytes + 255) & ~255LL; return r; };
  h.off_pi0_f = take(4LL * h.KP);
  h.off_P_f = take(4LL * h.KP * h.KP);
  h.off_W_f = take(4LL * K * h.J * h.DP);       // [k][j][i] (i fastest, padded to DP)
  h.off_c_f = take(4LL * (K + 1));               // c_k, then c_init
  h.off_logpi0_d = take(8LL * h.KP);
  h.off_logP_d = take(8LL * h.KP * h.KP);
  h.off_W_d = take(8LL * K * h.J * h.DP);
  h.off_c_d = take(8LL * (K + 1));
  h.off_W_hi = take(4LL * h.NT * h.KT);
  h.off_W_lo = take(4LL * h.NT * h.KT);
  h.total = o;
  return h;
}

// tcgen05 emission kernel (arhmm_tc.cu): returns 1 when the shape is not covered, 0 on success
int bn_launch_emission_tc(const unsigned char* d_blob, const float* d_x, const long long* d_offsets, int K, int D,
                          int lags, int n_trials, int max_T, float* d_Bsc, float* d_mx, cudaStream_t st);
