/*
 * behavenet_b200 -- C ABI of the B200 (sm_100a) hot-path library  (libbehavenet_b200.so)
 *
 * The reference (themattinthehatt/behavenet) is pure Python and has no FFI of its own: its hot
 * paths are reached through Python object protocols.  This header is the
 * boundary a maintainer binds with ctypes (INTEGRATION.md shows the stub); every entry point cites
 * the reference interface it stands behind.  Conventions:
 *
 *   - extern "C"; plain pointers and sizes; no torch / C++ types.
 *   - every function returns int: 0 = ok, < 0 = error; bn_last_error() gives the message of the
 *     last failure on the calling thread.
 *   - pointers named d_* are DEVICE pointers (HBM), h_* are HOST pointers.
 *   - the library never allocates or frees caller memory: activations / scratch live in a caller
 *     workspace sized by the *_workspace_bytes queries.  Plans own only a few KB of immutable
 *     device-side geometry tables.
 *   - stream is a cudaStream_t passed as void* (0 = legacy default stream).  All work is enqueued
 *     asynchronously on it; nothing synchronises the device.
 *   - fp32 tensors unless stated; frames are NCHW at the boundary (the reference's layout,
 *     data/data_generator.py:258-263), activations NHWC inside the workspace.
 */
#ifndef BEHAVENET_B200_H_
#define BEHAVENET_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BN_MAX_LAYERS 8
#define BN_ABI_VERSION 3

int bn_abi_version(void);
const char* bn_last_error(void);
/* number of kernels this library has launched so far in this process (bench.py "gpu_launches") */
int64_t bn_launch_count(void);
/* compute path selector: 0 = fp32 CUDA-core implicit GEMM everywhere; 1 (default) = tcgen05
 * TF32 tensor-core implicit GEMM for the layers it supports. */
int bn_set_tensor_core_mode(int mode);
int bn_get_tensor_core_mode(void);

/* ------------------------------------------------------------------------------------------------
 * Hot path 1: convolutional autoencoder   (reference: behavenet/models/aes.py)
 * ----------------------------------------------------------------------------------------------*/

/* Layer geometry = the per-layer lists of the reference's hparams dict
 * (aes.py:25-36, 229-242; built by ae_model_architecture_generator.py:482-592).
 * Encoder layer i: Conv2d(k, stride s) with explicit zero padding (top,bottom,left,right) and
 * LeakyReLU(0.05) (aes.py:71-114, 127-163).  Decoder layer i: ConvTranspose2d(k, s), output
 * cropped by (top,bottom,left,right), LeakyReLU(0.05) except Sigmoid after the last (aes.py:283-341,
 * 361-430, 467-470).  A NEGATIVE bottom / right crop (> -s) is the `output_padding` of the reference's
 * 'valid' padding mode (aes.py:382-405): the appended rows / columns lie outside the reach of every tap
 * and hold the bias only.  With per-session input / output layers (`fit_sess_io_layers`, aes.py:69-80,
 * 298-312) the caller puts the chosen session's tensors into the parameter table. */
typedef struct bn_cae_desc {
  int32_t n_layers;              /* conv layers per side, 1..BN_MAX_LAYERS */
  int32_t in_c, in_h, in_w;      /* hparams['ae_input_dim'] */
  int32_t n_latents;             /* hparams['n_ae_latents'] (= hidden_layer_size) */
  int32_t n_heads;               /* 1: FF only; 2: FF + logvar head (hparams['variational']) */
  int32_t enc_c[BN_MAX_LAYERS], enc_k[BN_MAX_LAYERS], enc_s[BN_MAX_LAYERS];
  int32_t enc_h[BN_MAX_LAYERS], enc_w[BN_MAX_LAYERS];           /* output dims per layer */
  int32_t enc_pt[BN_MAX_LAYERS], enc_pb[BN_MAX_LAYERS], enc_pl[BN_MAX_LAYERS], enc_pr[BN_MAX_LAYERS];
  int32_t dec_c0, dec_h0, dec_w0;                               /* ae_decoding_starting_dim */
  int32_t dec_c[BN_MAX_LAYERS], dec_k[BN_MAX_LAYERS], dec_s[BN_MAX_LAYERS];
  int32_t dec_h[BN_MAX_LAYERS], dec_w[BN_MAX_LAYERS];           /* output dims per layer (after crop) */
  int32_t dec_pt[BN_MAX_LAYERS], dec_pb[BN_MAX_LAYERS], dec_pl[BN_MAX_LAYERS], dec_pr[BN_MAX_LAYERS];
} bn_cae_desc;

typedef struct bn_cae_plan bn_cae_plan;

/* Parameter pointer table (device pointers to the torch-layout fp32 parameters, i.e. exactly the
 * reference's state_dict tensors).  Index helpers for n = n_layers:
 *   2*i, 2*i+1          encoding.encoder.conv{i}.weight (O,I,kh,kw) / .bias
 *   2n, 2n+1            encoding.FF.weight (L, C*H*W) / .bias
 *   2n+2, 2n+3          encoding.logvar.weight / .bias          (NULL when n_heads == 1)
 *   2n+4, 2n+5          decoding.FF.weight (C0*H0*W0, L) / .bias
 *   2n+6+2i, 2n+7+2i    decoding.decoder.convtranspose{i}.weight (I,O,kh,kw) / .bias
 * The gradient table uses the same indexing; kernels ACCUMULATE (+=) into it, matching autograd's
 * .grad semantics that AE.loss relies on across chunks (aes.py:754-769).  NULL entries are skipped. */
#define BN_CAE_N_PARAMS(n) (4 * (n) + 6)

int bn_cae_plan_create(const bn_cae_desc* desc, bn_cae_plan** out);
void bn_cae_plan_destroy(bn_cae_plan* plan);
/* bytes of the packed-weight cache (GEMM-ordered copies of the conv / FF weights) */
size_t bn_cae_packed_bytes(const bn_cae_plan* plan);
/* bytes of activation + gradient scratch for a batch of n frames */
size_t bn_cae_workspace_bytes(const bn_cae_plan* plan, int n);

/* Re-pack the weights after an optimizer step (or load_state_dict).  Layers whose weight pointer
 * is NULL are skipped (an encoder-only / decoder-only caller packs just its side). */
int bn_cae_pack_params(bn_cae_plan* plan, const float* const* d_params, void* d_packed,
                       void* stream);

/* ConvAEEncoder.forward (aes.py:181-218).  d_x: (n, C, H, W).  Writes d_mu (n, L) and, when
 * n_heads == 2, d_logvar (n, L) (= self.logvar(x), aes.py:215-216; PS-VAE vaes.py:1363).
 * Layer outputs are kept in the workspace for bn_cae_encode_bwd. */
int bn_cae_encode(bn_cae_plan* plan, int n, const float* d_x, const float* const* d_params,
                  const void* d_packed, void* d_ws, float* d_mu, float* d_logvar, void* stream);

/* The same forward pass over raw uint8 video, d_x: (n, C, H, W) bytes in 0..255 as stored in the
 * reference's HDF5 files; the first layer's loader converts with float32(v) / 255, i.e. the host-side
 * scaling of data/data_generator.py:258-263, so the frames cross PCIe and HBM once, at one byte per
 * pixel (the encode-only export path, fitting/eval.py:6-118).  Inference only: the workspace it
 * leaves is not valid for bn_cae_encode_bwd.  The first layer must have <= 4 input channels,
 * kernel 5, stride 2 and a multiple of 32 output channels (else an error is returned). */
int bn_cae_encode_u8(bn_cae_plan* plan, int n, const uint8_t* d_x, const float* const* d_params,
                     const void* d_packed, void* d_ws, float* d_mu, float* d_logvar, void* stream);

/* ConvAEDecoder.forward (aes.py:432-488) with the reconstruction loss fused into the last layer's
 * epilogue.  d_z: (n, L).  d_xhat: (n, C, H, W) or NULL.
 * When d_target != NULL the epilogue also computes, per reference chunk c = frame / chunk_size
 * (aes.py:751-763, vaes.py:655-672):
 *     d_sse[c] += sum_{frames in c} sum_pixels (x - xhat)^2 * mask          (double, accumulates)
 * and stores dL/d(pre-sigmoid) = grad_coef / len(chunk) * (xhat - x) * mask * xhat * (1 - xhat)
 * in the workspace for bn_cae_decode_bwd.  grad_coef = 2 / (C*H*W) reproduces losses.mse
 * (losses.py:36-59); grad_coef = 1 reproduces -gaussian_ll (losses.py:62-96).
 * frame_offset / n_total place this call's n frames inside the whole (possibly rank-sharded)
 * batch so that chunk membership and len(chunk) are those of the unsharded reference batch:
 * frame f belongs to chunk (frame_offset + f) / chunk_size; pass 0 and n when not sharded. */
int bn_cae_decode(bn_cae_plan* plan, int n, const float* d_z, const float* const* d_params,
                  const void* d_packed, void* d_ws, float* d_xhat, const float* d_target,
                  const float* d_mask, int chunk_size, int frame_offset, int n_total,
                  float grad_coef, double* d_sse, void* stream);

/* Backward of the decoder.  d_dxhat: (n, C, H, W) upstream gradient, or NULL to use the fused
 * loss gradient left in the workspace by bn_cae_decode.  Accumulates parameter gradients into
 * d_grads and writes d_dz (n, L). */
int bn_cae_decode_bwd(bn_cae_plan* plan, int n, const float* d_dxhat,
                      const float* const* d_params, const void* d_packed, void* d_ws,
                      float* const* d_grads, float* d_dz, void* stream);

/* Backward of the encoder.  d_dmu / d_dlogvar: (n, L) upstream gradients (d_dlogvar may be NULL).
 * Accumulates parameter gradients into d_grads. */
int bn_cae_encode_bwd(bn_cae_plan* plan, int n, const float* d_x, const float* d_dmu,
                      const float* d_dlogvar, const float* const* d_params, const void* d_packed,
                      void* d_ws, float* const* d_grads, void* stream);

/* The same backward pass in two calls, for data-parallel callers (gradient buckets
 * all-reduced underneath the remaining backward kernels; the reference has no counterpart, its
 * loss.backward() at aes.py:766 is one autograd sweep).  phase 1: heads and top conv layer -- on return the
 * gradients of encoding.FF / encoding.logvar and of the last encoding.encoder.conv layer are final on `stream`;
 * phase 2: everything below (call it after phase 1 with the same arguments and an untouched workspace);
 * phase 0: both, identical to bn_cae_encode_bwd. */
int bn_cae_encode_bwd_phase(bn_cae_plan* plan, int n, const float* d_x, const float* d_dmu,
                            const float* d_dlogvar, const float* const* d_params, const void* d_packed,
                            void* d_ws, float* const* d_grads, void* stream, int phase);

/* Run ONE layer operation of the plan on caller tensors (NHWC fp32): kernel-level parity tests of
 * the tensor-core kernels against the CUDA-core kernels, and per-kernel timing for bench.py.
 *   side: 0 = encoder layer `layer` (Conv2d), 1 = decoder layer `layer` (ConvTranspose2d)
 *   op:   0 = forward (bias + the layer's activation), 1 = backward-data (no activation mask),
 *         2 = weight gradient (d_in = big image, d_in2 = small image, d_out = torch-layout gradient,
 *             ACCUMULATES)
 *         3 = last decoder layer forward with the fused loss (side 1, layer n_layers - 1; reference chunks of
 *             200 frames, -gaussian_ll coefficient): d_in = small image, d_in2 = target (n, C, H, W);
 *             d_out = [x_hat (n C H W floats) | dL/d(pre-sigmoid) (n H W C floats) | per-chunk sums of squared
 *             errors (ceil(n / 200) doubles)]
 * "big"/"small" are the conv-input-side / conv-output-side images of the layer. */
int bn_cae_layer_op(bn_cae_plan* plan, int side, int layer, int op, int n, const float* d_in,
                    const float* d_in2, float* d_out, const float* const* d_params,
                    const void* d_packed, void* d_ws, void* stream);

/* Linear autoencoder (model_type 'linear': LinearAEEncoder / LinearAEDecoder aes.py:491-613 as AE.build_model
 * ties them, aes.py:684-687): z = x W^T + b, x_hat = z W + c.  d_x (n, P) frames flattened as the reference's
 * x.view(n, -1), d_W (L, P) = encoding.encoder.weight, d_b (L) = encoding.encoder.bias, d_c (P) = decoding.bias;
 * 1 <= L <= 64.
 *   bn_linae_forward: AE.forward (aes.py:714-716): d_z (n, L) and, unless NULL, d_xhat (n, P).
 *   bn_linae_decode:  LinearAEDecoder.forward (aes.py:588-613) on caller latents d_z (n, L) -> d_xhat (n, P).
 *   bn_linae_loss:    AE.loss (aes.py:722-773) with the reference's chunk rule: frames [frame_offset,
 *     frame_offset + n) of a batch of n_total frames; d_sse[chunk] += sum of mask * (x_hat - x)^2 (doubles, one per
 *     reference chunk of the WHOLE batch); gradients of sum_chunks mean(chunk) ACCUMULATE into d_gW / d_gb / d_gc
 *     (all NULL = loss only).  d_mask (n, P) may be NULL.  d_ws: bn_linae_workspace_bytes(n, P, L). */
size_t bn_linae_workspace_bytes(int n, int n_pixels, int n_latents);
int bn_linae_forward(int n, int n_pixels, int n_latents, const float* d_x, const float* d_W, const float* d_b,
                     const float* d_c, float* d_z, float* d_xhat, void* stream);
int bn_linae_decode(int n, int n_pixels, int n_latents, const float* d_z, const float* d_W, const float* d_c,
                    float* d_xhat, void* stream);
int bn_linae_loss(int n, int n_pixels, int n_latents, const float* d_x, const float* d_mask, const float* d_W,
                  const float* d_b, const float* d_c, int chunk_size, int frame_offset, int n_total, void* d_ws,
                  double* d_sse, float* d_gW, float* d_gb, float* d_gc, void* stream);

/* PS-VAE latent block (vaes.py:571-601, 669-696; losses.py:130-147, 284-372), one reference chunk
 * of n frames at a time (the MI/TC/DWKL estimators are pairwise over the chunk).
 * Inputs: d_pre (n, L) = FF output, d_logvar (n, L), frozen orthogonal d_A (n_labels, L) /
 * d_B (L - n_labels, L), diagonal d_Dw / d_Db (n_labels), d_eps (n, L) reparameterisation noise
 * (reparameterize, vaes.py:17-35: z = eps * exp(logvar) + mu), d_labels / d_labels_mask
 * (n, n_labels; mask may be NULL).
 * Outputs: d_mu, d_z (n, L), d_yhat (n, n_labels); d_terms[5] doubles that ACCUMULATE sums over
 * the chunk's frames of {sum_d (y - yhat)^2 mask, zs_kl, index-code MI, total correlation,
 * dimension-wise KL}; and the partial gradients of
 *     -alpha * label_ll + kl_s_w * zs_kl + kl_w * MI + beta * TC + kl_w * DWKL        (chunk means)
 * with respect to mu / logvar / z taken as independent variables: d_gmu_part, d_glogvar_part,
 * d_gz_part (n, L).  D's gradients accumulate (+=) into d_gDw / d_gDb (may be NULL). */
size_t bn_psvae_latent_workspace_bytes(int n, int n_latents);
int bn_psvae_latent(int n, int n_latents, int n_labels, const float* d_pre, const float* d_logvar,
                    const float* d_A, const float* d_B, const float* d_Dw, const float* d_Db,
                    const float* d_eps, const float* d_labels, const float* d_labels_mask,
                    float alpha, float beta, float kl_w, float kl_s_w, void* d_ws, float* d_mu, float* d_z,
                    float* d_yhat, double* d_terms, float* d_gmu_part, float* d_glogvar_part,
                    float* d_gz_part, float* d_gDw, float* d_gDb, void* stream);

/* Chain the decoder's dL/dz (d_gz_dec, may be NULL) and the partial gradients above back to the
 * encoder heads: d_gpre (n, L) = [A;B]^T (gmu_part + gz), d_glogvar (n, L) = glogvar_part +
 * gz * eps * exp(logvar), with gz = gz_part + gz_dec. */
int bn_psvae_latent_bwd(int n, int n_latents, int n_labels, const float* d_A, const float* d_B,
                        const float* d_eps, const float* d_logvar, const float* d_gz_dec,
                        const float* d_gmu_part, const float* d_glogvar_part,
                        const float* d_gz_part, float* d_gpre, float* d_glogvar, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Hot path 2: ARHMM E-step / log-likelihood / Viterbi
 * (reference call sites: fitting/arhmm_grid_search.py:170-204, fitting/eval.py:167; the arithmetic
 *  is ssm's HMM.expected_states / log_likelihood / most_likely_states for observations='ar')
 * ----------------------------------------------------------------------------------------------*/

/* Model parameters in ssm's parameterisation, HOST fp64:
 *   h_log_pi0 (K), h_log_Ps (K,K), h_As (K, D, D*lags) [block l multiplies x_{t-l-1}],
 *   h_bs (K, D), h_Sigmas (K, D, D).  The first `lags` steps of a trial use N(0, I).
 * bn_arhmm_params_bytes() sizes the DEVICE blob that bn_arhmm_pack_params fills (Cholesky
 * whitening is done on the host in fp64; both fp32 and fp64 copies are stored). */
size_t bn_arhmm_params_bytes(int K, int D, int lags);
int bn_arhmm_pack_params(int K, int D, int lags, const double* h_log_pi0, const double* h_log_Ps,
                         const double* h_As, const double* h_bs, const double* h_Sigmas,
                         void* h_blob);

/* Trials are concatenated: d_x is (total_T, D) fp32, d_offsets (n_trials + 1) int64 row offsets
 * (ragged trials, as the lists of (T_i, D) arrays the reference passes to ssm). */
size_t bn_arhmm_workspace_bytes(int K, int D, int lags, int64_t total_T, int n_trials, int fp64);

/* E-step: d_Ez (total_T, K) posteriors gamma, d_Ezz (n_trials, K, K) = sum_t xi_t,
 * d_logZ (n_trials) fp64 log normalisers.  Any output may be NULL (then only the forward pass is
 * run when d_Ez and d_Ezz are both NULL: that is hmm.log_likelihood). */
int bn_arhmm_estep(int K, int D, int lags, const void* d_params_blob, const float* d_x,
                   const int64_t* d_offsets, int n_trials, int64_t total_T, int max_T, void* d_ws,
                   float* d_Ez, float* d_Ezz, double* d_logZ, void* stream);

/* Viterbi: d_z (total_T) int32 most likely states; fp64 scores, first-index tie-break. */
int bn_arhmm_viterbi(int K, int D, int lags, const void* d_params_blob, const float* d_x,
                     const int64_t* d_offsets, int n_trials, int64_t total_T, int max_T,
                     void* d_ws, int32_t* d_z, void* stream);

/* M-step sufficient statistics for observations='ar' (weighted regression of x_t on
 * [x_{t-1..t-lags}, 1] with weights gamma_t(k), t >= lags), fp64 accumulation:
 *   d_stats (K, P+D, P+D) with P = D*lags + 1: the Gram matrix of [phi_t ; x_t], and d_counts (K).
 * Outputs accumulate (+=). */
int bn_arhmm_ar_stats(int K, int D, int lags, const float* d_x, const int64_t* d_offsets,
                      int n_trials, int64_t total_T, const float* d_Ez, double* d_stats,
                      double* d_counts, void* stream);

/* Host-side staging of the list of per-trial (T_i, D) arrays the reference passes to ssm
 * (arhmm_grid_search.py:170) into one (sum T_i, D) fp32 buffer (normally pinned memory that is then
 * copied to d_x): h_src[i] points at trial i's C-contiguous rows, fp32 or (h_is_f64[i] != 0) fp64;
 * h_is_f64 may be NULL.  Copies with `threads` host threads.  No device work. */
int bn_host_gather_rows(const void* const* h_src, const int64_t* h_rows, const int32_t* h_is_f64,
                        int n, int D, float* h_dst, int threads);

#ifdef __cplusplus
}
#endif
#endif  /* BEHAVENET_B200_H_ */
