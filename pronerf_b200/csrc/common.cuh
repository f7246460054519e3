// Shared declarations for the pronerf_b200 CUDA library (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/pronerf_b200.h"

namespace pn {

void set_error(const char* fmt, ...);
int cuda_fail(cudaError_t e, const char* what);

#define PN_CUDA_OK(expr)                                          \
  do {                                                            \
    cudaError_t _e = (expr);                                      \
    if (_e != cudaSuccess) return ::pn::cuda_fail(_e, #expr);     \
  } while (0)

#define PN_LAUNCH_OK(what)                                        \
  do {                                                            \
    cudaError_t _e = cudaGetLastError();                          \
    if (_e != cudaSuccess) return ::pn::cuda_fail(_e, what);      \
  } while (0)

#define PN_REQUIRE(cond, ...)                                     \
  do {                                                            \
    if (!(cond)) { ::pn::set_error(__VA_ARGS__); return PN_EINVAL; } \
  } while (0)

inline cudaStream_t as_stream(pn_stream_t s) { return reinterpret_cast<cudaStream_t>(s); }

// ---------------------------------------------------------------------------------------------
// Programmatic dependent launch for the kernels of the composed path (sampler MLP -> refine input -> refine MLP -> interval
// refinement -> view-direction term -> NeRF MLP -> composite).  Launched with the programmatic-stream-serialization attribute, a
// kernel's CTAs may become resident while the previous kernel of the stream is still draining: barrier init, TMEM allocation,
// the bias table and the first weight blocks of an MLP kernel then overlap the predecessor's last wave instead of following it.
// Contract: every such kernel executes pdl_wait() at its top (before ANY global access that a predecessor may have written or may
// still be reading, and before any early return), which blocks until the preceding grid has completed and flushed; pdl_launch()
// lets the successor's CTAs start.  MEASURED SLOWER on B200 (3-view step 3.570 -> 3.650 ms device-timed, 3.562 -> 3.806 ms end to
// end: early-resident CTAs of the small kernels sit next to the persistent MLP CTAs and the MLP kernels' early-resident clusters
// stream weights and poll barriers through the predecessor's last wave), so the attribute is OFF unless PN_PDL=1 is set in the
// environment; without it the device-side calls are no-ops and the launches are ordinary stream-ordered ones.
bool pdl_enabled();
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

template <typename... KArgs, typename... Args>
inline cudaError_t launch_chain(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}

constexpr int kHidden = 256;          // width of every trunk layer (netwidth / mmnetwidth = 256)
constexpr int kMaxLayers = 8;
constexpr int kMaxViews = 16;        // views per multi-view batch (pn_frame_t.n_views)
constexpr int kOutPad = 96;           // padded width of an output layer (4S+3 = 67 at S = 16)

// ---------------------------------------------------------------------------------------------
// One network's packed weights, owned by pn_ctx.
struct NetF32 {
  int n_layers = 0;                   // trunk layers + 1 output layer
  int in_dim[kMaxLayers] = {0};       // as given (nn.Linear in_features)
  int out_dim[kMaxLayers] = {0};
  int k_pad[kMaxLayers] = {0};        // in_dim rounded up to a multiple of 16 (trunk) / 4 (output)
  int k_off[kMaxLayers] = {0};        // first row of a trunk layer inside `trunk`
  int trunk_rows = 0;                 // sum of k_pad over trunk layers
  float* trunk = nullptr;             // device, all trunk layers back to back, k-major [trunk_rows][256]
  float* wout = nullptr;              // device, output layer k-major [k_pad[last]][kOutPad], zero padded
  float* bias = nullptr;              // device [n_layers][256] (output layer uses the first 64)
  bool loaded = false;
};

// How the first-layer operand rows are produced.
enum InputMode : int {
  IN_LOAD = 0,        // rows come from a dense [M, K0] tensor
  IN_LOAD2 = 1,       // DoNeRFTRT.forward: [M,63] embedded + [M,27] embedded_dirs (joins at the last layer)
  IN_ENCODE = 2,      // run_network: [M,3] points + per-ray [M/S,3] view dirs, encoded in-kernel
  IN_PLUECKER = 3,    // sampler: rays [N, stride] -> 6P Pluecker features generated in-kernel
  IN_CLASSIC = 5,     // classic NeRF (helpers.py:792-847): [M,3] points + per-ray view dirs, both encoded in-kernel (tensor-core tier)
  IN_LOAD16 = 4,      // rows come from a dense fp16 [M, K0] tensor (K0 % 8 == 0, 16-byte aligned): refine_input from pn_refine_input_f16
};

// Output-head activations, applied per column range of the last layer.
enum HeadAct : int { HEAD_NONE = 0, HEAD_SIGMOID = 1, HEAD_TANH = 2 };

struct MlpLaunch {
  const NetF32* net;
  int act;                 // 0 = ReLU (NeRF), 1 = ELU (sampler / refine)
  int input_mode;
  const float* in0;        // IN_LOAD: x; IN_LOAD2: embedded; IN_ENCODE: pts; IN_PLUECKER: rays; IN_LOAD16: fp16 x (cast)
  const float* in1;        // IN_LOAD2: embedded_dirs; IN_ENCODE: viewdirs
  int in_stride;           // row stride of in0 in floats
  int in1_stride;          // row stride of in1 in floats (IN_ENCODE: viewdirs; 3 dense, 11 inside a ray batch)
  int S;                   // samples per ray (IN_ENCODE: rows per viewdir)
  int P;                   // IN_PLUECKER: points per ray
  int64_t M;               // rows
  float* out;              // [M, out_dim]
  int head_lo[4];          // column ranges [head_lo[i], head_lo[i+1]) get head_act[i]
  int head_act[3];
  const float* dirterm_ready;   // tensor-core NeRF (IN_ENCODE): the per-ray view-direction term [M/S,4] has already been computed
                                // (by the interval-refinement kernel of the composed path): skip the pre-pass launch
};

int launch_mlp_f32(const MlpLaunch& L, cudaStream_t stream);
int pack_layer_f32(const float* W, const float* b, int out_dim, int in_dim, int k_pad, int n_pad, float* wt,
                   float* bias, int bias_pad, cudaStream_t stream);

// ---------------------------------------------------------------------------------------------
// fp32 layer-program tier (mlp_prog.cu): topologies outside the fused kernels, e.g. the classic NeRF of stage-2 checkpoints
constexpr int kMaxProgSteps = 16;
struct ProgStep {
  int in0_off, in0_len;      // first input segment: region offset inside the activation row, length (K rows of W)
  int in1_off, in1_len;      // optional second segment (skip / view-direction concatenations); len 0 = none
  int out_off, n_out;        // output region and width (<= 256)
  int act;                   // 1 = ReLU, 0 = linear
  int w_off, b_off;          // float offsets of the k-major weights [in0_len + in1_len][n_out] and the bias in the blob
};
struct NetProg {
  int n_steps = 0;
  ProgStep st[kMaxProgSteps];
  float* blob = nullptr;
  bool loaded = false;
};
void prog_free(NetProg& n);
int prog_load_nerf_classic(NetProg& net, const int* in_dims, const int* out_dims, const float* const* W, const float* const* b,
                           cudaStream_t stream);
int prog_launch(const NetProg& net, int input_mode, const float* in0, const float* in1, int in1_stride, int S, int64_t M,
                float* out, cudaStream_t stream);

// fused sort/lift + Pluecker + project/gather -> fp16 refine input (gather.cu); multi-view form of pn_refine_input_f16
int launch_refine_input_f16(const float* heads, int head_stride, const float* rays, const float* or_rays, int ray_stride,
                            const float* texels, const int* tex_index_host, int n_views, int64_t rays_per_view, int NN, int H,
                            int W, const float* project_mat, int64_t N, int S, float* depth, float* add, float* mul,
                            void* refine_in_f16, int32_t* x0y0, cudaStream_t stream, int64_t ray_base = 0);

// raw2outputs (elementwise.cu) with output rows placed for banded multi-view batches (pn_frame_t.out_view_stride)
int composite_mapped(const float* raw, const float* z, const float* rays, int ray_stride, int ray_d_col, const float* add,
                     const float* mul, float raw_clamp, int64_t N, int S, float* rgb, float* depth, float* disp, float* acc,
                     float* weights, int64_t rays_per_view, int64_t out_view_stride, int64_t ray_base, cudaStream_t st,
                     const float* dnorm = nullptr);
// pn_interval_refine that also writes ||d_ndc|| per ray (dnorm [N]) for composite_mapped: the compositing kernel then reads 4 B
// per ray instead of pulling a 44-byte ray row through 32-byte sectors for 12 of its bytes
int interval_refine_dnorm(const float* rays, int ray_stride, const float* depth, const float* refine_out, int refine_stride, int64_t N,
                          int S, float* z, float* query, float* dnorm, cudaStream_t st);

// View-direction term of DoNeRFTRT's last layer for one ray, fp32: o[k] = sum_j W7[k][256 + j] * gamma_4(v)[j]
// (helpers.py:666-671 with L = 4: [v, sin(2^l v), cos(2^l v)]_{l<4}; s_w = the 4 x 27 weights)
__device__ __forceinline__ float4 dirterm_of(const float* __restrict__ v, const float* __restrict__ s_w) {
  const float d[3] = {v[0], v[1], v[2]};
  float g[27];
  g[0] = d[0]; g[1] = d[1]; g[2] = d[2];
#pragma unroll
  for (int l = 0; l < 4; ++l)
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      float s, co;
      sincosf(d[c] * (float)(1 << l), &s, &co);
      g[3 + 6 * l + c] = s;
      g[6 + 6 * l + c] = co;
    }
  float o[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    float a = 0.f;
#pragma unroll
    for (int j = 0; j < 27; ++j) a = fmaf(s_w[k * 27 + j], g[j], a);
    o[k] = a;
  }
  return make_float4(o[0], o[1], o[2], o[3]);
}

// ---------------------------------------------------------------------------------------------
// Accurate fp32 helpers shared by kernels.  Nothing here may be compiled with --use_fast_math.
__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + expf(-x)); }
__device__ __forceinline__ float eluf_(float x) { return x > 0.f ? x : (expf(x) - 1.f); }

// [normalize(d), o x normalize(d)]   (F.normalize eps = 1e-12; helpers.py:629-632)
__device__ __forceinline__ void pluecker6(float ox, float oy, float oz, float dx, float dy, float dz, float* out6) {
  // torch's vector norm accumulates x*x with an FMA chain (closest match measured on CPU: 99.4% bit-equal)
  float n = sqrtf(__fmaf_rn(dz, dz, __fmaf_rn(dy, dy, __fmul_rn(dx, dx))));
  n = fmaxf(n, 1e-12f);
  float ux = __fdiv_rn(dx, n), uy = __fdiv_rn(dy, n), uz = __fdiv_rn(dz, n);
  out6[0] = ux; out6[1] = uy; out6[2] = uz;
  out6[3] = __fsub_rn(__fmul_rn(oy, uz), __fmul_rn(oz, uy));
  out6[4] = __fsub_rn(__fmul_rn(oz, ux), __fmul_rn(ox, uz));
  out6[5] = __fsub_rn(__fmul_rn(ox, uy), __fmul_rn(oy, ux));
}

}  // namespace pn
