// Shared helpers for the behavenet_b200 CUDA library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <atomic>

#define BN_LEAK 0.05f

extern thread_local char g_bn_err[512];
extern std::atomic<long long> g_bn_launches;

#define BN_FAIL(...)                                   \
  do {                                                 \
    snprintf(g_bn_err, sizeof(g_bn_err), __VA_ARGS__); \
    return -1;                                         \
  } while (0)

#define BN_CUDA(expr)                                                                          \
  do {                                                                                         \
    cudaError_t e__ = (expr);                                                                  \
    if (e__ != cudaSuccess) BN_FAIL("%s:%d %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(e__)); \
  } while (0)

// call after every kernel launch: counts the launch and surfaces launch-configuration errors
#define BN_LAUNCHED()                                                                        \
  do {                                                                                       \
    g_bn_launches.fetch_add(1, std::memory_order_relaxed);                                   \
    cudaError_t e__ = cudaGetLastError();                                                    \
    if (e__ != cudaSuccess) BN_FAIL("%s:%d launch -> %s", __FILE__, __LINE__, cudaGetErrorString(e__)); \
  } while (0)

#define BN_TRY(expr)        \
  do {                      \
    int r__ = (expr);       \
    if (r__ != 0) return r__; \
  } while (0)

static inline int bn_cdiv(long long a, long long b) { return (int)((a + b - 1) / b); }

// ------------------------------------------------------------------------------------------------
// Programmatic dependent launch.  A training step is ~70 dependent launches of 10-150 us kernels; with
// plain stream order every boundary costs the drain of one grid plus the launch and per-CTA set-up
// (barrier init, tcgen05.alloc, descriptor prefetch: ~2 us) of the next.  Kernels launched through
// bn_launch() carry cudaLaunchAttributeProgrammaticStreamSerialization: each calls bn_pdl_trigger() first
// (so the NEXT grid may be scheduled as soon as every CTA of this one has started and SM resources free
// up), does its set-up, and calls bn_pdl_wait() before its first global-memory access (which returns once
// the PREVIOUS grid has completed and flushed).  Rule: every kernel launched through bn_launch() executes
// bn_pdl_wait() in all threads, so completion of a grid implies completion of everything before it.
// Kernels launched with <<<>>> keep plain stream order on both sides.  BN_PDL=0 drops the attribute.
// ------------------------------------------------------------------------------------------------
#ifdef __CUDACC__
__device__ __forceinline__ void bn_pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void bn_pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
#endif

bool bn_pdl_enabled();

template <typename... KArgs, typename... Args>
static inline cudaError_t bn_launch(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                                    Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = bn_pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, static_cast<Args&&>(args)...);
}

// One tile class of an implicit GEMM: a logical pixel grid (Hm x Wm per frame) whose every pixel
// uses the same tap list.  fprop-form ops have one class; dgrad-form (transposed) ops have one
// class per output residue (stride^2 classes).
#define BN_MAX_TAPS 49
struct TapClass {
  int Hm, Wm;        // logical pixels per frame
  int oy0, ox0;      // output pixel of logical (0,0)
  int ntaps;
  signed char dy[BN_MAX_TAPS], dx[BN_MAX_TAPS];   // input offset per tap
  unsigned char wt[BN_MAX_TAPS];                  // weight tap index ky*k + kx
  unsigned char pad_[1];
};

// strided view of an image (strides in floats); NHWC-dense views have sc == 1
struct ImgView {
  const float* p;
  int H, W, C;
  long long sn, sy, sx, sc;
};

static inline ImgView nhwc_view(const float* p, int H, int W, int C) {
  ImgView v;
  v.p = p; v.H = H; v.W = W; v.C = C;
  v.sn = (long long)H * W * C; v.sy = (long long)W * C; v.sx = C; v.sc = 1;
  return v;
}
static inline ImgView nchw_view(const float* p, int H, int W, int C) {
  ImgView v;
  v.p = p; v.H = H; v.W = W; v.C = C;
  v.sn = (long long)H * W * C; v.sy = W; v.sx = 1; v.sc = (long long)H * W;
  return v;
}
