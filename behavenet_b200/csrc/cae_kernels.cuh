// Internal launch API of the CAE kernels (implemented in cae_simt.cu / cae_tc.cu / cae_misc.cu).
#pragma once
#include "bn_common.cuh"

enum { BN_ACT_NONE = 0, BN_ACT_LEAKY = 1, BN_ACT_SIGMOID = 2 };

// Geometry of one conv layer seen as "big image <-> small image":
//   encoder Conv2d:            big = layer input,  small = layer output
//   decoder ConvTranspose2d:   big = layer output, small = layer input
// small[y, x] <-> big[y*s - pt + ky, x*s - pl + kx].  Torch weight layout for both layer kinds is
// [c_small][c_big][ky][kx] (Conv2d: O,I,kh,kw; ConvTranspose2d: I,O,kh,kw).
struct ConvGeom {
  int Hb, Wb, Cb;   // big image
  int Hs, Ws, Cs;   // small image
  int k, s, pt, pl;
  // device-side class tables (owned by the plan)
  TapClass* d_fprop;    // 1 class
  TapClass* d_dgrad;    // n_dgrad classes
  int n_dgrad;
  int dgrad_maxM;       // max Hm*Wm over dgrad classes
  int dgrad_maxtaps;
  // packed weights (offsets in floats into the packed cache)
  size_t off_wf;        // fprop weights  [(tap, c_big)][c_small]
  size_t off_wd;        // dgrad weights  [(tap, c_small)][c_big]
  size_t off_wft;       // fprop weights, K-major + TF32-rounded  [c_small][(tap, c_big)]
  size_t off_wdt;       // dgrad weights, K-major + TF32-rounded  [c_big][(tap, c_small)]
  // torch parameter table indices
  int p_w, p_b;
};

// C[m, co] = act(bias[co] + sum_k A(m,k) W[k, co]) (* lrelu'(dact)) over the classes of an op.
//   in   : gathered image (NHWC for the vector path, any strides for the scalar path)
//   w    : packed weights, rows (wt[tap], ci), Co contiguous
//   out  : NHWC dense (Ho, Wo, Co); pixel (oy0 + os*ym, ox0 + os*xm)
int bn_launch_igemm(const ImgView& in, const float* w, const float* bias, float* out, int Ho, int Wo,
                    int Co, const float* dact, const TapClass* d_classes, int nclasses, int maxM,
                    int gs, int os, int n, int act, cudaStream_t st);

// Weight gradient: partial[z][(tap, cb)][cs] = sum_{m in split z} big(m, tap, cb) * small[m, cs],
// then grad[((cs*Cb + cb)*k*k + wt[tap])] += sum_z partial.
int bn_launch_wgrad(const ImgView& big, const float* small, const ConvGeom& g, int n, float* partial,
                    size_t partial_floats, float* grad, cudaStream_t st);
size_t bn_wgrad_partial_floats(const ConvGeom& g, int n);
// grad[((cs*Cb + cb)*KK + wt[tap])] += sum_z partial[z][(tap, cb)][cs]
int bn_launch_wgrad_reduce(const float* partial, int splits, int Ktot, int Cs, int Cb, int KK,
                           const TapClass* cls, float* grad, cudaStream_t st);
void bn_wgrad_reduce_defer_begin();
int bn_wgrad_reduce_flush(cudaStream_t st);
// true between bn_wgrad_reduce_defer_begin() and bn_wgrad_reduce_flush() on this thread
bool bn_reduce_deferring();
// queue out[c] += sum_r part[r * C + c] (r < rows) for the batched reduction launch; only while deferring
int bn_colsum_reduce_defer(const float* part, int rows, int C, float* out);

// out[c] += sum_m x[m*C + c]
int bn_launch_colsum(const float* x, long long M, int C, float* out, cudaStream_t st);

// Last decoder layer (thin output, C_big <= 4): transposed conv + sigmoid + fused recon loss.
//   small: NHWC (Hs, Ws, Cs) input; writes xhat (NCHW) to up to two destinations, dpre (NHWC).
int bn_launch_thin_dgrad(const float* small, const ConvGeom& g, const float* wd, const float* bias,
                         int n, float* xhat_ws, float* xhat_user, const float* target,
                         const float* mask, int chunk_size, int frame_offset, int n_total,
                         float grad_coef, double* sse, float* dpre, cudaStream_t st);

// Shared-memory-tiled fast paths for kernel-5 / stride-2 thin layers (cae_thin.cu).  Each returns 1
// when the geometry is not covered (caller falls back to the general kernels).
// wft != NULL selects the tcgen05 (TF32) form: K-major weights [c_small][(tap, c_big)]
int bn_launch_thin_fprop(const ImgView& big, const ConvGeom& g, const float* wf, const float* wft, const float* bias,
                         float* out, const float* dact, int act, int n, cudaStream_t st,
                         const unsigned char* big_u8 = nullptr,    // big_u8: raw 0..255 video with big's strides
                         float* colsum = nullptr, int* colsum_fused = nullptr);   // optional fused column sums of `out`
int bn_launch_thin_wgrad(const ImgView& big, const float* small, const ConvGeom& g, int n, float* partial,
                         size_t partial_floats, float* grad, cudaStream_t st);
// decoupled-role tensor-core form of bn_launch_thin_fprop (cae_thin_tc.cu); returns 1 when not applicable
int bn_launch_thin_fprop_tc2(const ImgView& big, const unsigned char* big_u8, const ConvGeom& g, const float* wft,
                             const float* bias, float* out, const float* dact, int act, int n, float* colsum,
                             cudaStream_t st);
// tcgen05 (TF32) form of bn_launch_thin_wgrad (cae_thin_tc.cu); returns 1 when not applicable
int bn_launch_thin_wgrad_tc(const ImgView& big, const float* small, const ConvGeom& g, int n, float* partial,
                            size_t partial_floats, float* grad, cudaStream_t st);
int bn_launch_thin_dgrad5(const float* small, const ConvGeom& g, const float* wd, const float* bias, int n,
                          float* xhat_ws, float* xhat_user, const float* target, const float* mask,
                          int chunk_size, int frame_offset, int n_total, float grad_coef, double* sse,
                          float* dpre, cudaStream_t st);

// tcgen05 (TF32) form of the fused last-layer forward + loss (cae_thin_tc.cu): GEMM over all taps + col2im
// epilogue; needs a target, writes x_hat only to the workspace copy; returns 1 when not applicable
int bn_launch_thin_dgrad_tc(const float* small, const ConvGeom& g, const float* wdt, const float* bias, int n,
                            float* xhat_ws, const float* target, const float* mask, int chunk_size, int frame_offset,
                            int n_total, float grad_coef, double* sse, float* dpre, cudaStream_t st);

// dpre[n,y,x,c] = dxhat[n,c,y,x] * xhat * (1 - xhat)
int bn_launch_sigmoid_bwd(const float* dxhat, const float* xhat, float* dpre, int n, int C, int H,
                          int W, cudaStream_t st);

// packing: src [cs][cb][kk] -> wf [(tap,cb)][cs], wd [(tap,cs)][cb] (fp32, exact) and the K-major
// TF32-rounded copies wft [cs][(tap,cb)], wdt [cb][(tap,cs)] for the tensor-core kernels
// one conv layer's weights: src [cs][cb][tap] (torch layout) -> wf / wd (fp32) and wft / wdt (TF32-rounded)
struct PackJob {
  const float* src;
  float *wf, *wd, *wft, *wdt;
  int Cs, Cb, kk;
  int block0, nrows, gx, gy;      // filled by bn_launch_pack_all: block range of this job
};
struct PackJobs {
  int n;
  PackJob j[16];               // 2 * BN_MAX_LAYERS (include/behavenet_b200.h)
};
int bn_launch_pack_all(PackJobs& jobs, cudaStream_t st);
int bn_launch_pack_heads(const float* w0, const float* w1, int L, int C, int H, int W, float* wcat,
                         cudaStream_t st);

// encoder heads forward: out_h[f, j] = b_h[j] + sum_i feat[f, i] * wcat[(h, j), i]
int bn_launch_heads_fwd(const float* feat, const float* wcat, const float* b0, const float* b1,
                        int n, int L, int nheads, int F, float* mu, float* logvar, cudaStream_t st);
// encoder heads backward: dfeat (with lrelu' of feat applied), dW (torch layout), db
int bn_launch_heads_bwd(const float* feat, const float* wcat, const float* dmu, const float* dlogvar,
                        int n, int L, int C, int H, int W, float* dpre_feat, float* gw0, float* gb0,
                        float* gw1, float* gb1, cudaStream_t st);
// decoder FF forward: h0[f, i_nhwc] = b[i_chw] + sum_j z[f, j] * W[i_chw, j]
int bn_launch_decff_fwd(const float* z, const float* w, const float* b, int n, int L, int C, int H,
                        int W, float* h0, cudaStream_t st);
int bn_launch_decff_bwd(const float* z, const float* w, const float* dh0, int n, int L, int C, int H,
                        int W, float* dz, float* gw, float* gb, cudaStream_t st);

// tcgen05 TF32 tensor-core implicit GEMM (cae_tc.cu).  Returns 1 if the shape is not supported
// (caller then uses the CUDA-core kernel), 0 on success, <0 on error.
//   wt : K-major packed weights [Co][wrow], wrow = k*k*Ci, TF32-rounded
//   split_buf : scratch for the split-K variant (few output tiles, long reduction); may be NULL
int bn_launch_igemm_tc(const ImgView& in, const float* wt, int wrow, const float* bias, float* out,
                       int Ho, int Wo, int Co, const float* dact, const TapClass* d_classes,
                       const TapClass* h_classes, int nclasses, int maxM, int maxtaps, int gs, int os, int n,
                       int act, float* split_buf, size_t split_floats, float* colsum, int* colsum_fused,
                       float* colpart, size_t colpart_floats, cudaStream_t st);
int bn_launch_wgrad_tc(const ImgView& big, const float* small, const ConvGeom& g, int n,
                       float* partial, size_t partial_floats, float* grad, cudaStream_t st);
