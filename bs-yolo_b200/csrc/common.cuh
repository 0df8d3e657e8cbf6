// Internal helpers shared by the libb2y kernels (sm_100a only).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/b2y.h"

typedef __nv_bfloat16 bf16;

void b2y_set_error(const char *fmt, ...);

#define B2Y_CUDA_CHECK(expr)                                                                      \
  do {                                                                                            \
    cudaError_t _e = (expr);                                                                      \
    if (_e != cudaSuccess) {                                                                      \
      b2y_set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
      return B2Y_ERR_CUDA;                                                                        \
    }                                                                                             \
  } while (0)

#define B2Y_LAUNCH_CHECK()                                                                          \
  do {                                                                                              \
    cudaError_t _e = cudaPeekAtLastError();                                                         \
    if (_e != cudaSuccess) {                                                                        \
      b2y_set_error("kernel launch failed: %s (%s:%d)", cudaGetErrorString(_e), __FILE__, __LINE__); \
      return B2Y_ERR_CUDA;                                                                          \
    }                                                                                               \
  } while (0)

#define B2Y_REQUIRE(cond, code, ...) \
  do {                               \
    if (!(cond)) {                   \
      b2y_set_error(__VA_ARGS__);    \
      return (code);                 \
    }                                \
  } while (0)

static inline int ceil_div(int a, int b) { return (a + b - 1) / b; }

// ---- device guard ------------------------------------------------------------------------------------------------
// The C entry points may be called while another device is current (one thread driving several GPUs, Engine(device=1)):
// every entry point makes the device that owns its memory current for the duration of the call and restores the caller's.
struct B2YDeviceGuard {
  int prev = -1, dev = -1;
  explicit B2YDeviceGuard(int device) : dev(device) {
    if (dev < 0) return;
    if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
    if (prev != dev) cudaSetDevice(dev);
  }
  ~B2YDeviceGuard() {
    if (dev >= 0 && prev >= 0 && prev != dev) cudaSetDevice(prev);
  }
};
// device that owns a device pointer (-1 if it is not device memory known to the runtime)
static inline int b2y_device_of(const void *ptr) {
  cudaPointerAttributes at;
  if (ptr == nullptr || cudaPointerGetAttributes(&at, ptr) != cudaSuccess) {
    (void)cudaGetLastError();
    return -1;
  }
  return (at.type == cudaMemoryTypeDevice || at.type == cudaMemoryTypeManaged) ? at.device : -1;
}

// ---- programmatic dependent launch (PDL) -------------------------------------------------------------------------
// Every kernel of the library is launched through b2y_launch with programmatic stream serialization: kernel N+1 may
// become resident (and run its prologue: barrier init, TMEM alloc, weight loads) while kernel N is still draining.
// Contract for every kernel launched this way:
//   * pdl_launch_dependents() as early as the kernel's own resources are secured (after tcgen05.alloc);
//   * pdl_wait() before the first access to memory another kernel may have written or may still read.  It returns only
//     when the preceding grid has completed and flushed; because every grid executes it, completion is transitive.
// B2Y_PDL=0 in the environment turns the launch attribute off (the device-side instructions are then no-ops).
bool b2y_pdl_enabled();

#ifdef __CUDACC__
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

template <typename... KArgs, typename... Args>
static inline cudaError_t b2y_launch(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                                     Args &&...args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = b2y_pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}
#endif

// A strided NHWC view: element (b,h,w,c) lives at base[((b*H + h)*W + w)*cs + c].
struct View {
  void *base;   // already offset to the view's first channel
  int cs;       // channels of the underlying buffer (pixel stride, elements)
  int c;        // channels of the view
  int dtype;    // b2y_dtype
};

__device__ __forceinline__ float silu_f(float x) { return __fdividef(x, 1.0f + __expf(-x)); }
__device__ __forceinline__ float sigmoid_f(float x) { return __fdividef(1.0f, 1.0f + __expf(-x)); }

__device__ __forceinline__ void unpack8(const uint4 &v, float *f) {
  const __nv_bfloat162 *p = reinterpret_cast<const __nv_bfloat162 *>(&v);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float2 t = __bfloat1622float2(p[i]);
    f[2 * i] = t.x;
    f[2 * i + 1] = t.y;
  }
}
__device__ __forceinline__ uint4 pack8(const float *f) {
  uint4 v;
  __nv_bfloat162 *p = reinterpret_cast<__nv_bfloat162 *>(&v);
#pragma unroll
  for (int i = 0; i < 4; ++i) p[i] = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
  return v;
}
__device__ __forceinline__ void unpack4(const uint2 &v, float *f) {
  const __nv_bfloat162 *p = reinterpret_cast<const __nv_bfloat162 *>(&v);
  float2 a = __bfloat1622float2(p[0]), b = __bfloat1622float2(p[1]);
  f[0] = a.x; f[1] = a.y; f[2] = b.x; f[3] = b.y;
}
__device__ __forceinline__ uint2 pack4(const float *f) {
  uint2 v;
  __nv_bfloat162 *p = reinterpret_cast<__nv_bfloat162 *>(&v);
  p[0] = __floats2bfloat162_rn(f[0], f[1]);
  p[1] = __floats2bfloat162_rn(f[2], f[3]);
  return v;
}

// ---- parameters of the convolution family (dense / depthwise / stem) -------------------------------------------
struct ConvParams {
  const bf16 *in; int in_cs;
  void *out; int out_cs; int out_f32;
  const bf16 *res; int res_cs; int res_mode;
  const void *w;        // dense: bf16 [cout_pad][kh*kw*cin]; depthwise: fp32 [kh*kw][c]; stem: fp32 [cout][27]
  const float *bias;    // fp32 [cout_pad]
  int B, H, W, Ho, Wo, cin, cout, cout_pad;
  int kh, kw, sh, sw, ph, pw, act;
  int asym;                          // 1: padding only before (top/left) -- the 2x2 space-to-depth stem
  int grouped;                       // 1: block-diagonal weights in groups of 16 channels (depthwise on the tensor cores)
  int dw_chunk, dw_in_stride;        // depthwise only: output channel j reads input channel (j / dw_chunk) * dw_in_stride +
                                     // j % dw_chunk of the view (0: contiguous) -- Attention.pe over the v slices of qkv
  int epi_mode, a_off, A, nc;        // fused Detect decode (0 none, 1 DFL box, 2 class sigmoid) -> y (B,4+nc,A)
  float det_stride;
};

// launchers (one per translation unit); all enqueue on `st` and return a b2y error code
int launch_conv_simt(const ConvParams &p, cudaStream_t st);
int launch_dwconv(const ConvParams &p, cudaStream_t st);
bool dw3_tile_supported(const ConvParams &p);          // dw3_tile.cu: shared-memory tiled depthwise 3x3 (stride 1 | 2)
int launch_dw3_tile(const ConvParams &p, cudaStream_t st);

struct ConvTcPlan;  // opaque: tensor maps + geometry for one tcgen05 conv
struct TcConfig {
  int mb;      // 128-row accumulator blocks per tile (0: heuristic)
  int per_sm;  // CTAs meant to co-reside per SM: 1..3 (0: heuristic = 1)
};
bool conv_tc_supported(const ConvParams &p, char *why, int why_len);
int conv_tc_plan_create(const ConvParams &p, int device, const TcConfig &cfg, ConvTcPlan **out);
void conv_tc_plan_destroy(ConvTcPlan *p);
int launch_conv_tc(const ConvTcPlan *p, cudaStream_t st);
int conv_tc_plan_trace(ConvTcPlan *p, cudaStream_t st);  // debug: one traced launch, timeline printed to stderr

// fused Bottleneck (two chained 3x3 convs + residual) on the tensor cores, conv_pair.cu
struct PairParams {
  const bf16 *in; int in_cs;
  void *out; int out_cs;
  const void *w1, *w2;        // bf16 [chid][9*cin], [cout][9*chid]
  const float *b1, *b2;       // fp32 [chid], [cout]
  int B, H, W, cin, chid, cout, add;
};
struct ConvPairPlan;
bool conv_pair_supported(const PairParams &p, char *why, int why_len);
int conv_pair_plan_create(const PairParams &p, int device, ConvPairPlan **out);
void conv_pair_plan_destroy(ConvPairPlan *p);
int launch_conv_pair(const ConvPairPlan *p, cudaStream_t st);

// fused depthwise tail of a PMSFA block, pmsfa_dw.cu
struct PmsfaParams {
  const bf16 *in; int in_cs;
  bf16 *out; int out_cs;
  const float *w2, *b2, *w3, *b3;  // conv2 [25][c/2], [c/2]; conv3 [49][c/4], [c/4]
  int B, H, W, c;
};
int launch_pmsfa_dw(const PmsfaParams &p, cudaStream_t st);

// MSCAAttention up to the four branch maps + their global average pools, msca.cu
struct MscaParams {
  const bf16 *in; int in_cs;
  bf16 *out; int out_cs;          // four branch maps: channel slices [jC, (j+1)C) of `out`
  float *gap;                     // (B, 4, C) fp32
  const float *params;            // packed, see plan.msca
  int B, H, W, C;
};
int launch_msca_branch(const MscaParams &p, cudaStream_t st);

int launch_sppf_pool(const bf16 *in, int in_cs, bf16 *out, int out_cs, int c, int B, int H, int W, cudaStream_t st);
int launch_resample(const bf16 *in, int in_cs, bf16 *out, int out_cs, int c, int B, int H, int W, int scale,
                    cudaStream_t st);
int launch_d2s(const bf16 *in, int in_cs, bf16 *out, int out_cs, int c, int B, int Hi, int Wi, cudaStream_t st);  // seg.cu
int launch_attention(const bf16 *qkv, int qkv_cs, bf16 *out, int out_cs, int B, int N, int nh, int kd, int hd,
                     cudaStream_t st);
int launch_rowcol_mean(const bf16 *x, int cs, int c, int B, int H, int W, int nbands, float *rowm, float *colm, cudaStream_t st);
int launch_ela_gate(const float *rowm, const float *colm, const float *params, int c, int B, int H, int W, int ks,
                    int groups, float eps, int nbands, float *hgate, float *wgate, float *chgate, cudaStream_t st);
int launch_ela_apply(const bf16 *x, int cs, bf16 *out, int out_cs, int c, int B, int H, int W, const float *hgate,
                     const float *wgate, const float *chgate, float ch_c, float sp_c, float res_c, cudaStream_t st);
int launch_gap4(const bf16 *const *br, const int *br_cs, int c, int B, int HW, float *gap, cudaStream_t st);
int launch_msca_gate(const float *gap, const float *se_w, const float *se_b, int c, int B, float *gate,
                     cudaStream_t st);
int launch_wsum4(const bf16 *const *br, const int *br_cs, const float *gate, bf16 *out, int out_cs, int c, int B,
                 int HW, cudaStream_t st);
