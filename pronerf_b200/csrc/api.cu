// C ABI glue: error state, the context (packed weights + scratch), and the composed render path.
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "common.cuh"
#include "tc.cuh"

namespace pn {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

bool pdl_enabled() {
  static const bool on = getenv("PN_PDL") && atoi(getenv("PN_PDL")) != 0;      // opt-in: measured slower
  return on;
}

int cuda_fail(cudaError_t e, const char* what) {
  set_error("CUDA error %d (%s) in %s", (int)e, cudaGetErrorString(e), what);
  return PN_ECUDA;
}

}  // namespace pn

using namespace pn;

struct pn_ctx {
  int device = 0;
  NetF32 f32[3];
  NetTC tc[3];
  NetProg nerf_classic;             // classic NeRF topology (stage-2 checkpoints), fp32 layer-program tier
  // scratch for pn_render_rays, grown on demand (never shrunk)
  float* scratch = nullptr;
  size_t scratch_floats = 0;
  // per-view buffers for pn_render_view_host
  float* view_buf = nullptr;
  size_t view_floats = 0;
  float* pm_dev = nullptr;         // projection matrices of the host-buffer entry points: [kMaxViews][8][12]
  // host-buffer entry points: results go home on this stream while the next chunk / the next call renders
  cudaStream_t d2h_stream = nullptr;
  cudaEvent_t chunk_done = nullptr;
  // device frames of pn_render_views_host_async, double-buffered: call k + 2 may only composite into slot k % 2 once call k's
  // download has finished (waited for on the device, not on the host)
  struct HostSlot {
    float* out = nullptr;            // rgb [n,3] then depth [n]
    size_t out_floats = 0;
    cudaEvent_t rendered = nullptr, downloaded = nullptr;
    bool used = false;
    int64_t ticket = -1;             // the latest call whose download `downloaded` marks
  } hs[2];
  int64_t next_ticket = 0;
  int sm_count = 0;
  // stage timing ring (pn_ctx_profile)
  bool profile = false;
  std::vector<cudaEvent_t> ev;      // PN_PROFILE_RING * (PN_N_STAGES + 1), created lazily
  int prof_head = 0, prof_count = 0;
};

static int ensure(float** buf, size_t* have, size_t want) {
  if (*have >= want) return PN_OK;
  if (*buf) cudaFree(*buf);
  *buf = nullptr;
  *have = 0;
  cudaError_t e = cudaMalloc((void**)buf, want * sizeof(float));
  if (e != cudaSuccess) { set_error("cudaMalloc of %zu bytes failed: %s", want * sizeof(float), cudaGetErrorString(e)); return PN_ENOMEM; }
  *have = want;
  return PN_OK;
}

static void free_net(NetF32& n) {
  if (n.trunk) cudaFree(n.trunk);
  if (n.wout) cudaFree(n.wout);
  if (n.bias) cudaFree(n.bias);
  n = NetF32();
}

extern "C" {

int pn_version(void) { return PN_VERSION; }
const char* pn_last_error(void) { return g_err; }
int pn_has_bf16_tier(void) { return tc_available() ? 1 : 0; }
int pn_debug_tc_timeline(void* dev_buf_208_i64) { tc_set_timeline(reinterpret_cast<long long*>(dev_buf_208_i64)); return PN_OK; }
int pn_debug_tc_clock(void* dev_buf_12_i64) { tc_set_clock(reinterpret_cast<long long*>(dev_buf_12_i64)); return PN_OK; }

int pn_device_check(int device) {
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count == 0) { set_error("no CUDA device: %s", cudaGetErrorString(e)); return PN_ENODEVICE; }
  if (device < 0 || device >= count) { set_error("device %d out of range (%d present)", device, count); return PN_ENODEVICE; }
  cudaDeviceProp prop;
  PN_CUDA_OK(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10) {
    set_error("device %d is sm_%d%d; this library is built for sm_100a (B200) only and has no fallback", device, prop.major, prop.minor);
    return PN_ENODEVICE;
  }
  return PN_OK;
}

int pn_ctx_create(int device, pn_ctx_t** out) {
  PN_REQUIRE(out, "pn_ctx_create: out is NULL");
  int rc = pn_device_check(device);
  if (rc != PN_OK) return rc;
  PN_CUDA_OK(cudaSetDevice(device));
  pn_ctx* c = new pn_ctx();
  c->device = device;
  cudaError_t e = cudaMalloc((void**)&c->pm_dev, (size_t)kMaxViews * 8 * 12 * sizeof(float));
  if (e != cudaSuccess) { delete c; return cuda_fail(e, "cudaMalloc(pm_dev)"); }
  *out = c;
  return PN_OK;
}

void pn_ctx_destroy(pn_ctx_t* c) {
  if (!c) return;
  cudaSetDevice(c->device);
  for (int i = 0; i < 3; ++i) { free_net(c->f32[i]); tc_free_net(c->tc[i]); }
  prog_free(c->nerf_classic);
  if (c->scratch) cudaFree(c->scratch);
  if (c->view_buf) cudaFree(c->view_buf);
  if (c->pm_dev) cudaFree(c->pm_dev);
  if (c->chunk_done) cudaEventDestroy(c->chunk_done);
  for (auto& h : c->hs) {
    if (h.out) cudaFree(h.out);
    if (h.rendered) cudaEventDestroy(h.rendered);
    if (h.downloaded) cudaEventDestroy(h.downloaded);
  }
  if (c->d2h_stream) cudaStreamDestroy(c->d2h_stream);
  for (cudaEvent_t e : c->ev) cudaEventDestroy(e);
  delete c;
}

int pn_ctx_profile(pn_ctx_t* c, int enable) {
  PN_REQUIRE(c, "pn_ctx_profile: ctx is NULL");
  PN_CUDA_OK(cudaSetDevice(c->device));
  if (enable && c->ev.empty()) {
    c->ev.resize((size_t)PN_PROFILE_RING * (PN_N_STAGES + 1));
    for (auto& e : c->ev) PN_CUDA_OK(cudaEventCreate(&e));
  }
  c->profile = enable != 0;
  c->prof_head = 0;
  c->prof_count = 0;
  return PN_OK;
}

int pn_ctx_profile_read(pn_ctx_t* c, float* ms, int max_frames) {
  if (!c || !ms) { set_error("pn_ctx_profile_read: null pointer"); return PN_EINVAL; }
  if (c->prof_count == 0) return 0;
  PN_CUDA_OK(cudaSetDevice(c->device));
  int n = c->prof_count < max_frames ? c->prof_count : max_frames;
  int first = (c->prof_head - c->prof_count + 2 * PN_PROFILE_RING) % PN_PROFILE_RING;
  int last = (c->prof_head - 1 + PN_PROFILE_RING) % PN_PROFILE_RING;
  PN_CUDA_OK(cudaEventSynchronize(c->ev[(size_t)last * (PN_N_STAGES + 1) + PN_N_STAGES]));
  for (int i = 0; i < n; ++i) {
    int slot = (first + i) % PN_PROFILE_RING;
    cudaEvent_t* e = &c->ev[(size_t)slot * (PN_N_STAGES + 1)];
    for (int k = 0; k < PN_N_STAGES; ++k) PN_CUDA_OK(cudaEventElapsedTime(&ms[i * PN_N_STAGES + k], e[k], e[k + 1]));
  }
  c->prof_head = 0;
  c->prof_count = 0;
  return n;
}

int pn_ctx_load_net(pn_ctx_t* c, int net, int n_layers, const int* in_dims, const int* out_dims,
                    const float* const* W, const float* const* b, pn_stream_t stream) {
  PN_REQUIRE(c && in_dims && out_dims && W && b, "pn_ctx_load_net: null pointer");
  PN_REQUIRE(net >= 0 && net < 3, "pn_ctx_load_net: unknown net id %d", net);
  PN_REQUIRE(n_layers >= 2 && n_layers <= kMaxLayers, "pn_ctx_load_net: n_layers=%d unsupported (2..%d)", n_layers, kMaxLayers);
  for (int l = 0; l < n_layers - 1; ++l) {
    PN_REQUIRE(out_dims[l] == kHidden, "pn_ctx_load_net: layer %d has width %d; only 256-wide trunks are built", l, out_dims[l]);
    PN_REQUIRE(l == 0 || in_dims[l] == kHidden, "pn_ctx_load_net: layer %d has %d inputs; skip connections inside the trunk are not built", l, in_dims[l]);
  }
  PN_REQUIRE(in_dims[0] >= 1 && in_dims[0] <= 288, "pn_ctx_load_net: first layer has %d inputs (max 288)", in_dims[0]);
  const int last = n_layers - 1;
  PN_REQUIRE(out_dims[last] >= 1 && out_dims[last] <= kOutPad, "pn_ctx_load_net: output width %d unsupported (1..%d)", out_dims[last], kOutPad);
  PN_REQUIRE(in_dims[last] >= kHidden && in_dims[last] <= 288, "pn_ctx_load_net: output layer has %d inputs (256..288)", in_dims[last]);
  PN_CUDA_OK(cudaSetDevice(c->device));
  cudaStream_t st = as_stream(stream);
  NetF32& n = c->f32[net];
  free_net(n);
  if (net == PN_NET_NERF) prog_free(c->nerf_classic);      // one shading network per context
  n.n_layers = n_layers;
  int rows = 0;
  for (int l = 0; l < n_layers; ++l) {
    n.in_dim[l] = in_dims[l];
    n.out_dim[l] = out_dims[l];
    if (l < last) {
      n.k_pad[l] = (in_dims[l] + 15) / 16 * 16;
      n.k_off[l] = rows;
      rows += n.k_pad[l];
    } else {
      n.k_pad[l] = (in_dims[l] + 3) / 4 * 4;
    }
  }
  n.trunk_rows = rows;
  PN_CUDA_OK(cudaMalloc((void**)&n.trunk, (size_t)rows * kHidden * sizeof(float)));
  PN_CUDA_OK(cudaMalloc((void**)&n.wout, (size_t)n.k_pad[last] * kOutPad * sizeof(float)));
  PN_CUDA_OK(cudaMalloc((void**)&n.bias, (size_t)n_layers * kHidden * sizeof(float)));
  for (int l = 0; l < n_layers; ++l) {
    PN_REQUIRE(W[l] && b[l], "pn_ctx_load_net: layer %d weight/bias pointer is NULL", l);
    int rc;
    if (l < last)
      rc = pack_layer_f32(W[l], b[l], out_dims[l], in_dims[l], n.k_pad[l], kHidden, n.trunk + (size_t)n.k_off[l] * kHidden,
                          n.bias + (size_t)l * kHidden, kHidden, st);
    else
      rc = pack_layer_f32(W[l], b[l], out_dims[l], in_dims[l], n.k_pad[l], kOutPad, n.wout, n.bias + (size_t)l * kHidden, kHidden, st);
    if (rc != PN_OK) return rc;
  }
  n.loaded = true;
  int rc = tc_load_net(c->tc[net], net, n_layers, in_dims, out_dims, W, b, st);
  if (rc != PN_OK) return rc;
  // the kernels of the composed path read their (static) weights BEFORE the programmatic-dependency wait: packing must be
  // complete, not merely ordered, when the first of them is launched
  PN_CUDA_OK(cudaStreamSynchronize(st));
  return PN_OK;
}

int pn_ctx_load_nerf_classic(pn_ctx_t* c, const int* in_dims, const int* out_dims, const float* const* W, const float* const* b,
                             pn_stream_t stream) {
  PN_REQUIRE(c && in_dims && out_dims && W && b, "pn_ctx_load_nerf_classic: null pointer");
  for (int l = 0; l < 12; ++l) PN_REQUIRE(W[l] && b[l], "pn_ctx_load_nerf_classic: tensor %d is NULL", l);
  PN_CUDA_OK(cudaSetDevice(c->device));
  free_net(c->f32[PN_NET_NERF]);                            // one shading network per context
  int rc = prog_load_nerf_classic(c->nerf_classic, in_dims, out_dims, W, b, as_stream(stream));
  if (rc != PN_OK) return rc;
  return tc_load_nerf_classic(c->tc[PN_NET_NERF], in_dims, out_dims, W, b, as_stream(stream));
}

// classic topology: fp32 only in this build, and said loudly
static int classic_precision_ok(int precision, const char* who) {
  if (precision == PN_PREC_FP32) return PN_OK;
  set_error("%s: this form of the classic NeRF topology runs in PN_PREC_FP32 only (the tensor-core tier covers it through "
            "pn_run_network / pn_render_rays, where both encodings are generated in-kernel)", who);
  return PN_ESTATE;
}

// ------------------------------------------------------------------------------------------------ MLP entry points
static int run_network_impl(pn_ctx_t* c, const float* pts, const float* viewdirs, int viewdir_stride, int64_t N, int S, float* raw,
                            int precision, pn_stream_t stream, const float* dirterm_ready);

static void heads_sampler(MlpLaunch& L, int S) {
  L.head_lo[0] = 0; L.head_lo[1] = S; L.head_lo[2] = 3 * S; L.head_lo[3] = 3 * S + 3;
  L.head_act[0] = HEAD_SIGMOID; L.head_act[1] = HEAD_NONE; L.head_act[2] = HEAD_SIGMOID;
}
static void heads_refine(MlpLaunch& L, int S) {
  L.head_lo[0] = 0; L.head_lo[1] = S; L.head_lo[2] = 4 * S; L.head_lo[3] = 4 * S + 3;
  L.head_act[0] = HEAD_SIGMOID; L.head_act[1] = HEAD_TANH; L.head_act[2] = HEAD_SIGMOID;
}
static void heads_none(MlpLaunch& L) {
  for (int i = 0; i < 4; ++i) L.head_lo[i] = 0;
  for (int i = 0; i < 3; ++i) L.head_act[i] = HEAD_NONE;
}

static int run_mlp(pn_ctx_t* c, int net, MlpLaunch& L, int precision, cudaStream_t st) {
  L.net = &c->f32[net];
  if (!c->f32[net].loaded) { set_error("network %d not loaded (pn_ctx_load_net)", net); return PN_ESTATE; }
  if (precision == PN_PREC_FP32) return launch_mlp_f32(L, st);
  if (precision == PN_PREC_F16) return tc_launch_mlp(c->tc[net], L, st);
  set_error("unknown precision %d", precision);
  return PN_EINVAL;
}

int pn_sampler_forward(pn_ctx_t* c, const float* x, int64_t N, int S, float* out, int precision, pn_stream_t stream) {
  if (N == 0) return PN_OK;            // empty batch
  PN_REQUIRE(c && x && out && N >= 0 && S >= 1, "pn_sampler_forward: bad arguments");
  const NetF32& n = c->f32[PN_NET_SAMPLER];
  PN_REQUIRE(!n.loaded || n.out_dim[n.n_layers - 1] == 3 * S + 3, "pn_sampler_forward: net output width %d != 3S+3 (S=%d)",
             n.out_dim[n.n_layers - 1], S);
  MlpLaunch L{};
  L.act = 1; L.input_mode = IN_LOAD; L.in0 = x; L.in1 = nullptr; L.in_stride = n.in_dim[0]; L.S = S; L.P = 0; L.M = N; L.out = out;
  heads_sampler(L, S);
  return run_mlp(c, PN_NET_SAMPLER, L, precision, as_stream(stream));
}

int pn_sampler_forward_rays(pn_ctx_t* c, const float* rays, int ray_stride, int64_t N, int S, int P, float* out, int precision,
                            pn_stream_t stream) {
  if (N == 0) return PN_OK;            // empty batch
  PN_REQUIRE(c && rays && out && N >= 0 && S >= 1 && P >= 1 && ray_stride >= 6, "pn_sampler_forward_rays: bad arguments");
  const NetF32& n = c->f32[PN_NET_SAMPLER];
  PN_REQUIRE(!n.loaded || (n.out_dim[n.n_layers - 1] == 3 * S + 3 && n.in_dim[0] == 6 * P),
             "pn_sampler_forward_rays: net shape does not match S=%d P=%d", S, P);
  MlpLaunch L{};
  L.act = 1; L.input_mode = IN_PLUECKER; L.in0 = rays; L.in1 = nullptr; L.in_stride = ray_stride; L.S = S; L.P = P; L.M = N; L.out = out;
  heads_sampler(L, S);
  return run_mlp(c, PN_NET_SAMPLER, L, precision, as_stream(stream));
}

int pn_refine_forward(pn_ctx_t* c, const float* x, int64_t N, int S, float* out, int precision, pn_stream_t stream) {
  if (N == 0) return PN_OK;            // empty batch
  PN_REQUIRE(c && x && out && N >= 0 && S >= 1, "pn_refine_forward: bad arguments");
  const NetF32& n = c->f32[PN_NET_REFINE];
  PN_REQUIRE(!n.loaded || n.out_dim[n.n_layers - 1] == 4 * S + 3, "pn_refine_forward: net output width %d != 4S+3 (S=%d)",
             n.out_dim[n.n_layers - 1], S);
  MlpLaunch L{};
  L.act = 1; L.input_mode = IN_LOAD; L.in0 = x; L.in1 = nullptr; L.in_stride = n.in_dim[0]; L.S = S; L.P = 0; L.M = N; L.out = out;
  heads_refine(L, S);
  return run_mlp(c, PN_NET_REFINE, L, precision, as_stream(stream));
}

int pn_refine_forward_f16(pn_ctx_t* c, const void* x_f16, int64_t N, int S, float* out, pn_stream_t stream) {
  if (N == 0) return PN_OK;            // empty batch
  PN_REQUIRE(c && x_f16 && out && N >= 0 && S >= 1, "pn_refine_forward_f16: bad arguments");
  const NetF32& n = c->f32[PN_NET_REFINE];
  PN_REQUIRE(!n.loaded || n.out_dim[n.n_layers - 1] == 4 * S + 3, "pn_refine_forward_f16: net output width %d != 4S+3 (S=%d)",
             n.out_dim[n.n_layers - 1], S);
  MlpLaunch L{};
  L.act = 1; L.input_mode = IN_LOAD16; L.in0 = reinterpret_cast<const float*>(x_f16); L.in1 = nullptr; L.in_stride = n.in_dim[0];
  L.S = S; L.P = 0; L.M = N; L.out = out;
  heads_refine(L, S);
  return run_mlp(c, PN_NET_REFINE, L, PN_PREC_F16, as_stream(stream));
}

static int check_nerf(const NetF32& n) {
  if (!n.loaded) return PN_OK;   // run_mlp reports it
  PN_REQUIRE(n.in_dim[0] == 63 && n.in_dim[n.n_layers - 1] == kHidden + 27 && n.out_dim[n.n_layers - 1] == 4,
             "NeRF net must be 63 -> 256.. -> (256+27) -> 4 (DoNeRFTRT with multires 10 / 4)");
  return PN_OK;
}

int pn_nerf_forward(pn_ctx_t* c, const float* embedded, const float* embedded_dirs, int64_t M, float* raw, int precision,
                    pn_stream_t stream) {
  if (M == 0) return PN_OK;            // empty batch
  PN_REQUIRE(c && embedded && embedded_dirs && raw && M >= 0, "pn_nerf_forward: bad arguments");
  if (c->nerf_classic.loaded) {
    int rcp = classic_precision_ok(precision, "pn_nerf_forward");
    if (rcp != PN_OK) return rcp;
    PN_CUDA_OK(cudaSetDevice(c->device));
    return prog_launch(c->nerf_classic, IN_LOAD2, embedded, embedded_dirs, 27, 1, M, raw, as_stream(stream));
  }
  int rc = check_nerf(c->f32[PN_NET_NERF]);
  if (rc != PN_OK) return rc;
  MlpLaunch L{};
  L.act = 0; L.input_mode = IN_LOAD2; L.in0 = embedded; L.in1 = embedded_dirs; L.in_stride = 63; L.S = 1; L.P = 0; L.M = M; L.out = raw;
  heads_none(L);
  return run_mlp(c, PN_NET_NERF, L, precision, as_stream(stream));
}

int pn_run_network(pn_ctx_t* c, const float* pts, const float* viewdirs, int viewdir_stride, int64_t N, int S, float* raw,
                   int precision, pn_stream_t stream) {
  return run_network_impl(c, pts, viewdirs, viewdir_stride, N, S, raw, precision, stream, nullptr);
}
}  // extern "C"

// dirterm_ready: the per-ray view-direction term of DoNeRFTRT's last layer, already on the device (composed path), or NULL
static int run_network_impl(pn_ctx_t* c, const float* pts, const float* viewdirs, int viewdir_stride, int64_t N, int S, float* raw,
                            int precision, pn_stream_t stream, const float* dirterm_ready) {
  if (N == 0) return PN_OK;            // empty batch
  PN_REQUIRE(c && pts && viewdirs && raw && N >= 0 && S >= 1 && viewdir_stride >= 3, "pn_run_network: bad arguments");
  if (c->nerf_classic.loaded) {
    PN_CUDA_OK(cudaSetDevice(c->device));
    if (precision == PN_PREC_F16)
      return tc_launch_nerf_classic(c->tc[PN_NET_NERF], pts, viewdirs, viewdir_stride, S, N * S, raw, as_stream(stream));
    int rcp = classic_precision_ok(precision, "pn_run_network");
    if (rcp != PN_OK) return rcp;
    return prog_launch(c->nerf_classic, IN_ENCODE, pts, viewdirs, viewdir_stride, S, N * S, raw, as_stream(stream));
  }
  int rc = check_nerf(c->f32[PN_NET_NERF]);
  if (rc != PN_OK) return rc;
  MlpLaunch L{};
  L.act = 0; L.input_mode = IN_ENCODE; L.in0 = pts; L.in1 = viewdirs; L.in_stride = 3; L.in1_stride = viewdir_stride; L.S = S; L.P = 0; L.M = N * S; L.out = raw;
  L.dirterm_ready = dirterm_ready;
  heads_none(L);
  return run_mlp(c, PN_NET_NERF, L, precision, as_stream(stream));
}

extern "C" {

// ------------------------------------------------------------------------------------------------ the whole path
// Scratch layout per ray (floats): heads 3S+3 | depth S | add S | mul S | depth3d S | refine_in 6S+3NN*S |
// refine_out 4S+3 | z S | query 3S | raw 4S
}  // extern "C"

// rows [ray_base, ray_base + f->N) of a multi-view batch whose views hold f->rays_per_view rays each: every pointer of `f`
// already points at the chunk's first row; ray_base only selects the view (matrices, neighbour ordering) of a row.
static int render_rays_chunk(pn_ctx_t* c, const pn_frame_t* f, int64_t ray_base, pn_stream_t stream) {
  if (f && f->N == 0) return PN_OK;    // empty batch: nothing to validate or launch
  PN_REQUIRE(c && f, "pn_render_rays: null pointer");
  PN_REQUIRE(f->rays && f->or_rays && f->texels && f->project_mat && f->rgb && f->depth, "pn_render_rays: frame has a NULL buffer");
  PN_REQUIRE(f->N >= 0 && f->S >= 1 && f->S <= 64 && f->NN >= 1 && f->NN <= 8 && f->P >= 1 && f->H >= 2 && f->W >= 2,
             "pn_render_rays: bad frame shape");
  const int64_t N = f->N;
  const int S = f->S, NN = f->NN, P = f->P;
  const int nv = f->n_views > 1 ? f->n_views : 1;
  PN_REQUIRE(nv <= kMaxViews, "pn_render_rays: n_views=%d exceeds PN_MAX_VIEWS=%d", nv, kMaxViews);
  PN_REQUIRE(nv == 1 || (f->rays_per_view >= 1 && ray_base >= 0 && ray_base + N <= f->rays_per_view * nv),
             "pn_render_rays: rows [%lld, +%lld) are outside n_views=%d x rays_per_view=%lld", (long long)ray_base, (long long)N, nv,
             (long long)f->rays_per_view);
  const int64_t rpv = nv > 1 ? f->rays_per_view : N;
  PN_REQUIRE(f->out_view_stride == 0 || f->out_view_stride >= rpv, "pn_render_rays: out_view_stride=%lld is smaller than rays_per_view=%lld",
             (long long)f->out_view_stride, (long long)rpv);
  int tix[kMaxViews * 8];
  for (int v = 0; v < nv; ++v)
    for (int k = 0; k < NN; ++k) tix[v * NN + k] = (nv > 1 && f->tex_index_views) ? f->tex_index_views[v * NN + k] : f->tex_index[k];
  PN_CUDA_OK(cudaSetDevice(c->device));
  cudaStream_t st = as_stream(stream);
  const NetF32& ns = c->f32[PN_NET_SAMPLER];
  const NetF32& nr = c->f32[PN_NET_REFINE];
  PN_REQUIRE(ns.loaded && nr.loaded && (c->f32[PN_NET_NERF].loaded || c->nerf_classic.loaded), "pn_render_rays: load all three networks first");

  PN_REQUIRE(ns.in_dim[0] == 6 * P && ns.out_dim[ns.n_layers - 1] == 3 * S + 3, "pn_render_rays: sampler net shape does not match P=%d S=%d", P, S);
  PN_REQUIRE(nr.in_dim[0] == 6 * S + 3 * NN * S && nr.out_dim[nr.n_layers - 1] == 4 * S + 3, "pn_render_rays: refine net shape does not match S=%d NN=%d", S, NN);
  int rc = check_nerf(c->f32[PN_NET_NERF]);
  if (rc != PN_OK) return rc;

  const int hs = 3 * S + 3, ri = 6 * S + 3 * NN * S, ro = 4 * S + 3;
  // carve the arena; every sub-buffer starts on a 256-byte boundary (float4 / bulk accesses need 16)
  auto al = [](size_t n) { return (n + 63) & ~(size_t)63; };
  const size_t nN = (size_t)N;
  const size_t total = al(nN * hs) + 4 * al(nN * S) + al(nN * ri) + al(nN * ro) + al(nN * S) + al(nN * 3 * S) + al(nN * 4 * S) + al(nN);
  rc = ensure(&c->scratch, &c->scratch_floats, total);
  if (rc != PN_OK) return rc;
  float* p = c->scratch;
  float* heads = p;      p += al(nN * hs);
  float* depth = p;      p += al(nN * S);
  float* add = p;        p += al(nN * S);
  float* mul = p;        p += al(nN * S);
  float* depth3d = p;    p += al(nN * S);
  float* rin = p;        p += al(nN * ri);
  float* rout = p;       p += al(nN * ro);
  float* z = p;          p += al(nN * S);
  float* query = p;      p += al(nN * 3 * S);
  float* raw = p;        p += al(nN * 4 * S);
  float* dnorm = p;

  cudaEvent_t* pe = nullptr;
  if (c->profile) {
    pe = &c->ev[(size_t)c->prof_head * (PN_N_STAGES + 1)];
    c->prof_head = (c->prof_head + 1) % PN_PROFILE_RING;
    if (c->prof_count < PN_PROFILE_RING) ++c->prof_count;
  }
#define PN_STAGE_MARK(i) do { if (pe) PN_CUDA_OK(cudaEventRecord(pe[i], st)); } while (0)
  PN_STAGE_MARK(0);
  // (1) sampler MLP  trt.py:628
  {
    MlpLaunch L{};
    L.act = 1; L.S = S; L.P = P; L.M = N; L.out = heads;
    if (f->mm_input) { L.input_mode = IN_LOAD; L.in0 = f->mm_input; L.in_stride = 6 * P; }
    else             { L.input_mode = IN_PLUECKER; L.in0 = f->rays; L.in_stride = 11; }
    heads_sampler(L, S);
    rc = run_mlp(c, PN_NET_SAMPLER, L, f->precision, st);
    if (rc != PN_OK) return rc;
  }
  PN_STAGE_MARK(1);
  // fp16 tier: sort/lift + Pluecker + project/gather run as ONE kernel that writes the refine input as fp16 rows,
  // which the refine MLP's first-layer operand loads chunk for chunk (timed as the project_gather stage)
  if (f->texels_ready) PN_CUDA_OK(cudaStreamWaitEvent(st, reinterpret_cast<cudaEvent_t>(f->texels_ready), 0));
  // (rows of the fp16 refine input must be whole 16-byte chunks for the refine kernel's loads: otherwise the per-stage route)
  const bool fused_input = f->precision == PN_PREC_F16 && (S == 4 || S == 8 || S == 16) && c->tc[PN_NET_REFINE].supported && ri % 8 == 0;
  if (fused_input) {
    PN_STAGE_MARK(2);
    PN_STAGE_MARK(3);
    rc = launch_refine_input_f16(heads, hs, f->rays, f->or_rays, 11, f->texels, tix, nv, rpv, NN, f->H, f->W, f->project_mat, N, S,
                                 depth, add, mul, rin, nullptr, st, ray_base);
    if (rc != PN_OK) return rc;
    if (f->texels_done) PN_CUDA_OK(cudaEventRecord(reinterpret_cast<cudaEvent_t>(f->texels_done), st));
  } else {
  PN_REQUIRE(ray_base == 0, "pn_render_rays: chunked passes need the fused refine-input kernel (tensor-core tier, S in 4/8/16)");
  // (2) sort + lift  trt.py:631-637
  rc = pn_sort_lift(heads, hs, f->rays, 11, N, S, depth, add, mul, nullptr, depth3d, stream);
  if (rc != PN_OK) return rc;
  PN_STAGE_MARK(2);
  // (3) refine input: Pluecker part + projected colours written in place  trt.py:649-661
  rc = pn_refine_pluecker(f->rays, 11, depth, N, S, rin, ri, stream);
  if (rc != PN_OK) return rc;
  PN_STAGE_MARK(3);
  for (int v = 0; v < nv; ++v) {       // one gather per view: each has its own matrices and neighbour ordering
    const int64_t o = v * rpv;
    rc = pn_project_gather(f->texels, tix + v * NN, NN, f->H, f->W, f->project_mat + (size_t)v * NN * 12, f->or_rays + o * 11,
                           f->or_rays + o * 11 + 3, 11, depth3d + o * S, rpv, S, rin + o * ri, ri, 6 * S, nullptr, stream);
    if (rc != PN_OK) return rc;
  }
  if (f->texels_done) PN_CUDA_OK(cudaEventRecord(reinterpret_cast<cudaEvent_t>(f->texels_done), st));
  }
  PN_STAGE_MARK(4);
  // (4) refine MLP  trt.py:668
  {
    MlpLaunch L{};
    L.act = 1; L.input_mode = fused_input ? IN_LOAD16 : IN_LOAD; L.in0 = rin; L.in_stride = ri; L.S = S; L.P = 0; L.M = N; L.out = rout;
    heads_refine(L, S);
    rc = run_mlp(c, PN_NET_REFINE, L, f->precision, st);
    if (rc != PN_OK) return rc;
  }
  PN_STAGE_MARK(5);
  // (5) interval refinement + offsets  trt.py:671-681
  // (the same kernel leaves ||d_ndc|| for the compositing kernel.  Measured and dropped: also computing the NeRF last layer's
  //  per-ray view-direction term here -- one lane of a ray's S doing 12 sincosf + 108 FMAs while the others idle made this
  //  kernel 0.088 ms longer per 3-view step to save a 0.030 ms launch)
  rc = interval_refine_dnorm(f->rays, 11, depth, rout, ro, N, S, z, query, dnorm, st);
  if (rc != PN_OK) return rc;
  PN_STAGE_MARK(6);
  // (6) encode + NeRF MLP  trt.py:691 ; viewdirs = rays[:, 8:11]
  rc = run_network_impl(c, query, f->rays + 8, 11, N, S, raw, f->precision, stream, nullptr);
  if (rc != PN_OK) return rc;
  PN_STAGE_MARK(7);
  // (7) composite  trt.py:694
  rc = composite_mapped(raw, z, f->rays, 11, 3, add, mul, 0.f, N, S, f->rgb, f->depth, nullptr, nullptr, nullptr, rpv,
                        f->out_view_stride, ray_base, st, dnorm);
  if (rc != PN_OK) return rc;
  PN_STAGE_MARK(8);
#undef PN_STAGE_MARK
  return PN_OK;
}

static int render_views_host_impl(pn_ctx_t* c, int H, int W, double fx, double fy, double cx, double cy, int n_views,
                                  const float* c2w_host, const float* texels, const int* tex_index_host,
                                  const float* project_mat_host, int NN, int S, int P, int precision, int row0, int nrows,
                                  float* rgb_host, float* depth_host, int64_t host_view_stride, void* texels_ready_event,
                                  void* texels_done_event, pn_stream_t stream, int64_t* ticket, bool two_chunks);

extern "C" {

int pn_render_rays(pn_ctx_t* c, const pn_frame_t* f, pn_stream_t stream) {
  if (f && f->N == 0) return PN_OK;
  PN_REQUIRE(c && f, "pn_render_rays: null pointer");
  const int nv = f->n_views > 1 ? f->n_views : 1;
  PN_REQUIRE(nv == 1 || (f->rays_per_view >= 1 && f->rays_per_view * nv == f->N),
             "pn_render_rays: N=%lld is not n_views=%d x rays_per_view=%lld", (long long)f->N, nv, (long long)f->rays_per_view);
  return render_rays_chunk(c, f, 0, stream);
}


// ------------------------------------------------------------------------------------------------ host-buffer flavour
int pn_render_view_host(pn_ctx_t* c, int H, int W, double fx, double fy, double cx, double cy, const float* c2w_host,
                        const float* texels, const int* tex_index_host, const float* project_mat_host, int NN, int S,
                        int P, int precision, int row0, int nrows, float* rgb_host, float* depth_host,
                        pn_stream_t stream) {
  PN_REQUIRE(c && c2w_host && texels && project_mat_host && rgb_host && depth_host, "pn_render_view_host: null pointer");
  PN_REQUIRE(NN >= 1 && NN <= 8 && row0 >= 0 && nrows >= 0 && row0 + nrows <= H, "pn_render_view_host: bad shape");
  if (nrows == 0) return PN_OK;
  PN_CUDA_OK(cudaSetDevice(c->device));
  cudaStream_t st = as_stream(stream);
  const int64_t n = (int64_t)nrows * W;
  int rc = ensure(&c->view_buf, &c->view_floats, (size_t)n * (11 + 11 + 3 + 1));
  if (rc != PN_OK) return rc;
  float* rays = c->view_buf;
  float* or_rays = rays + n * 11;
  float* rgb = or_rays + n * 11;
  float* depth = rgb + n * 3;
  PN_CUDA_OK(cudaMemcpyAsync(c->pm_dev, project_mat_host, (size_t)NN * 12 * sizeof(float), cudaMemcpyHostToDevice, st));
  rc = pn_raygen(H, W, fx, fy, cx, cy, c2w_host, 0.f, 1.f, 1.f, 10.f, row0, nrows, rays, or_rays, stream);
  if (rc != PN_OK) return rc;
  pn_frame_t f;
  memset(&f, 0, sizeof(f));
  f.rays = rays; f.or_rays = or_rays; f.mm_input = nullptr; f.texels = texels; f.project_mat = c->pm_dev;
  for (int k = 0; k < 8; ++k) f.tex_index[k] = (tex_index_host && k < NN) ? tex_index_host[k] : k;
  f.N = n; f.S = S; f.NN = NN; f.P = P; f.H = H; f.W = W; f.precision = precision; f.rgb = rgb; f.depth = depth;
  rc = pn_render_rays(c, &f, stream);
  if (rc != PN_OK) return rc;
  PN_CUDA_OK(cudaMemcpyAsync(rgb_host, rgb, (size_t)n * 3 * sizeof(float), cudaMemcpyDeviceToHost, st));
  PN_CUDA_OK(cudaMemcpyAsync(depth_host, depth, (size_t)n * sizeof(float), cudaMemcpyDeviceToHost, st));
  PN_CUDA_OK(cudaStreamSynchronize(st));
  return PN_OK;
}

int pn_render_views_host_async(pn_ctx_t* c, int H, int W, double fx, double fy, double cx, double cy, int n_views,
                               const float* c2w_host, const float* texels, const int* tex_index_host,
                               const float* project_mat_host, int NN, int S, int P, int precision, int row0, int nrows,
                               float* rgb_host, float* depth_host, int64_t host_view_stride, void* texels_ready_event,
                               void* texels_done_event, pn_stream_t stream, int64_t* ticket) {
  return render_views_host_impl(c, H, W, fx, fy, cx, cy, n_views, c2w_host, texels, tex_index_host, project_mat_host, NN, S, P, precision,
                                row0, nrows, rgb_host, depth_host, host_view_stride, texels_ready_event, texels_done_event, stream, ticket,
                                false);
}
}  // extern "C"

// two_chunks: split the pass so that the first chunk's frames travel while the second renders -- lowers the LATENCY of a lone
// synchronous call; a pipelined caller overlaps the whole download with the next call instead and renders in one pass.
static int render_views_host_impl(pn_ctx_t* c, int H, int W, double fx, double fy, double cx, double cy, int n_views,
                                  const float* c2w_host, const float* texels, const int* tex_index_host,
                                  const float* project_mat_host, int NN, int S, int P, int precision, int row0, int nrows,
                                  float* rgb_host, float* depth_host, int64_t host_view_stride, void* texels_ready_event,
                                  void* texels_done_event, pn_stream_t stream, int64_t* ticket, bool two_chunks) {
  PN_REQUIRE(c && c2w_host && texels && project_mat_host && rgb_host && depth_host && ticket, "pn_render_views_host: null pointer");
  PN_REQUIRE(NN >= 1 && NN <= 8 && n_views >= 0 && n_views <= kMaxViews && H >= 2 && W >= 2,
             "pn_render_views_host: bad shape (n_views=%d, at most %d per batch)", n_views, kMaxViews);
  PN_REQUIRE(row0 >= 0 && nrows >= 0 && row0 + nrows <= H, "pn_render_views_host: rows [%d, +%d) outside H=%d", row0, nrows, H);
  const int64_t npv = (int64_t)nrows * W, n = npv * n_views;
  const int64_t hvs = host_view_stride > 0 ? host_view_stride : npv;
  PN_REQUIRE(hvs >= npv, "pn_render_views_host: host_view_stride=%lld is smaller than the band (%lld rays)", (long long)hvs, (long long)npv);
  PN_CUDA_OK(cudaSetDevice(c->device));
  cudaStream_t st = as_stream(stream);
  if (!c->d2h_stream) {
    PN_CUDA_OK(cudaStreamCreateWithFlags(&c->d2h_stream, cudaStreamNonBlocking));
    PN_CUDA_OK(cudaEventCreateWithFlags(&c->chunk_done, cudaEventDisableTiming));
    for (auto& h : c->hs) {
      PN_CUDA_OK(cudaEventCreateWithFlags(&h.rendered, cudaEventDisableTiming));
      PN_CUDA_OK(cudaEventCreateWithFlags(&h.downloaded, cudaEventDisableTiming));
    }
  }
  const int64_t tk = c->next_ticket++;
  *ticket = tk;
  pn_ctx::HostSlot& hs = c->hs[tk & 1];
  if (n == 0) return PN_OK;            // empty batch: nothing of its own to wait for (pn_wait still covers every earlier call)
  if (hs.out_floats < (size_t)n * 4) {
    if (hs.used) PN_CUDA_OK(cudaEventSynchronize(hs.downloaded));      // growing a frame that may still be on its way home
    int rc = ensure(&hs.out, &hs.out_floats, (size_t)n * 4);
    if (rc != PN_OK) return rc;
  }
  int rc = ensure(&c->view_buf, &c->view_floats, (size_t)n * 22);
  if (rc != PN_OK) return rc;
  float* rays = c->view_buf;
  float* or_rays = rays + n * 11;
  float* rgb = hs.out;
  float* depth = rgb + n * 3;
  PN_CUDA_OK(cudaMemcpyAsync(c->pm_dev, project_mat_host, (size_t)n_views * NN * 12 * sizeof(float), cudaMemcpyHostToDevice, st));
  for (int v = 0; v < n_views; ++v) {
    rc = pn_raygen(H, W, fx, fy, cx, cy, c2w_host + 12 * v, 0.f, 1.f, 1.f, 10.f, row0, nrows, rays + v * npv * 11, or_rays + v * npv * 11, stream);
    if (rc != PN_OK) return rc;
  }
  if (hs.used) PN_CUDA_OK(cudaStreamWaitEvent(st, hs.downloaded, 0));  // the call two back has left this device frame
  pn_frame_t f;
  memset(&f, 0, sizeof(f));
  f.rays = rays; f.or_rays = or_rays; f.mm_input = nullptr; f.texels = texels; f.project_mat = c->pm_dev;
  for (int k = 0; k < 8; ++k) f.tex_index[k] = (tex_index_host && k < NN) ? tex_index_host[k] : k;
  f.N = n; f.S = S; f.NN = NN; f.P = P; f.H = H; f.W = W; f.precision = precision; f.rgb = rgb; f.depth = depth;
  f.n_views = n_views; f.rays_per_view = npv; f.tex_index_views = tex_index_host; f.texels_ready = texels_ready_event;
  f.texels_done = texels_done_event;
  // rows [a, b) of the batch -> the host frame set [n_views][hvs], view by view
  auto download = [&](int64_t a, int64_t b) -> int {
    for (int v = 0; v < n_views; ++v) {
      const int64_t lo = a > v * npv ? a : v * npv, hi = b < (v + 1) * npv ? b : (v + 1) * npv;
      if (lo >= hi) continue;
      const int64_t h0 = v * hvs + (lo - v * npv);
      PN_CUDA_OK(cudaMemcpyAsync(rgb_host + h0 * 3, rgb + lo * 3, (size_t)(hi - lo) * 3 * sizeof(float), cudaMemcpyDeviceToHost, c->d2h_stream));
      PN_CUDA_OK(cudaMemcpyAsync(depth_host + h0, depth + lo, (size_t)(hi - lo) * sizeof(float), cudaMemcpyDeviceToHost, c->d2h_stream));
    }
    return PN_OK;
  };
  // Two chunks on the tensor-core tier: the first is a whole number of waves of every persistent MLP kernel (one wave =
  // a 512-ray unit per CTA pair), so splitting adds no tail; its rgb / depth travel to the host on the download stream while
  // the second chunk renders.  Every ray is independent: the chunks' results are bit-identical to the single pass.
  int64_t n_a = 0;
  static const int env_chunks = getenv("PN_HOST_CHUNKS") ? atoi(getenv("PN_HOST_CHUNKS")) : -1;     // tuning aid: 0 / 1 forces
  if (env_chunks >= 0) two_chunks = env_chunks != 0;
  if (two_chunks && precision == PN_PREC_F16 && (S == 4 || S == 8 || S == 16) && c->tc[PN_NET_REFINE].supported && (6 * S + 3 * NN * S) % 8 == 0) {
    if (c->sm_count == 0) PN_CUDA_OK(cudaDeviceGetAttribute(&c->sm_count, cudaDevAttrMultiProcessorCount, c->device));
    const int64_t wave = (int64_t)(c->sm_count / 2) * 512;
    n_a = wave > 0 ? (n * 7 / 8) / wave * wave : 0;      // the rest still renders longer than the first chunk's frames travel
    if (n_a >= n) n_a = 0;
  }
  if (n_a > 0) {
    pn_frame_t fa = f, fb = f;
    fa.texels_done = nullptr;                            // the second chunk holds the pass's last texel reader
    fa.N = n_a;
    fb.N = n - n_a;
    fb.rays = rays + n_a * 11; fb.or_rays = or_rays + n_a * 11; fb.rgb = rgb + n_a * 3; fb.depth = depth + n_a;
    rc = render_rays_chunk(c, &fa, 0, stream);
    if (rc != PN_OK) return rc;
    PN_CUDA_OK(cudaEventRecord(c->chunk_done, st));
    PN_CUDA_OK(cudaStreamWaitEvent(c->d2h_stream, c->chunk_done, 0));
    rc = download(0, n_a);
    if (rc != PN_OK) return rc;
    rc = render_rays_chunk(c, &fb, n_a, stream);
    if (rc != PN_OK) return rc;
  } else {
    rc = pn_render_rays(c, &f, stream);
    if (rc != PN_OK) return rc;
  }
  PN_CUDA_OK(cudaEventRecord(hs.rendered, st));
  PN_CUDA_OK(cudaStreamWaitEvent(c->d2h_stream, hs.rendered, 0));
  rc = download(n_a, n);
  if (rc != PN_OK) return rc;
  PN_CUDA_OK(cudaEventRecord(hs.downloaded, c->d2h_stream));
  hs.used = true;
  hs.ticket = tk;
  return PN_OK;
}

extern "C" {

int pn_wait(pn_ctx_t* c, int64_t ticket) {
  PN_REQUIRE(c && ticket >= 0 && ticket < c->next_ticket, "pn_wait: unknown ticket %lld", (long long)ticket);
  PN_CUDA_OK(cudaSetDevice(c->device));
  // The download stream is in order.  A slot whose latest download belongs to this call or an earlier one is waited for; the
  // other slot may already carry a LATER call, which a pipelined caller does not want to wait for -- unless it is this
  // ticket's own slot, reused since (then its event is the only handle left on that download).
  for (int s = 0; s < 2; ++s) {
    pn_ctx::HostSlot& hs = c->hs[s];
    if (hs.used && (hs.ticket <= ticket || s == (int)(ticket & 1))) PN_CUDA_OK(cudaEventSynchronize(hs.downloaded));
  }
  return PN_OK;
}

int pn_render_views_host(pn_ctx_t* c, int H, int W, double fx, double fy, double cx, double cy, int n_views,
                         const float* c2w_host, const float* texels, const int* tex_index_host,
                         const float* project_mat_host, int NN, int S, int P, int precision, float* rgb_host,
                         float* depth_host, void* texels_ready_event, pn_stream_t stream) {
  int64_t ticket = 0;
  int rc = render_views_host_impl(c, H, W, fx, fy, cx, cy, n_views, c2w_host, texels, tex_index_host, project_mat_host, NN, S, P,
                                  precision, 0, H, rgb_host, depth_host, 0, texels_ready_event, nullptr, stream, &ticket, true);
  if (rc != PN_OK) return rc;
  rc = pn_wait(c, ticket);
  if (rc != PN_OK) return rc;
  PN_CUDA_OK(cudaStreamSynchronize(as_stream(stream)));      // synchronous flavour: the caller's stream is idle on return, as before
  return PN_OK;
}

// ------------------------------------------------------------------------------------------------ peer frame buffers
// The tile gather of a sharded frame as direct stores over NVLink: the destination rank allocates the frame
// buffer here (cudaMalloc, so that the IPC handle covers exactly this allocation), the other ranks map it into their
// address space and hand `frame + band offset` to pn_render_rays as rgb / depth -- the compositing kernel's stores then land
// in the destination GPU's memory while the band is still being rendered, and no collective follows.
int pn_peer_alloc(int device, size_t bytes, void** dev_ptr, unsigned char* handle64) {
  PN_REQUIRE(dev_ptr && handle64 && bytes > 0, "pn_peer_alloc: bad arguments");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
  PN_CUDA_OK(cudaSetDevice(device));
  void* p = nullptr;
  PN_CUDA_OK(cudaMalloc(&p, bytes));
  cudaIpcMemHandle_t h;
  cudaError_t e = cudaIpcGetMemHandle(&h, p);
  if (e != cudaSuccess) { cudaFree(p); return cuda_fail(e, "cudaIpcGetMemHandle"); }
  memcpy(handle64, &h, 64);
  *dev_ptr = p;
  return PN_OK;
}

int pn_peer_open(int device, const unsigned char* handle64, void** dev_ptr) {
  PN_REQUIRE(dev_ptr && handle64, "pn_peer_open: bad arguments");
  PN_CUDA_OK(cudaSetDevice(device));
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, 64);
  PN_CUDA_OK(cudaIpcOpenMemHandle(dev_ptr, h, cudaIpcMemLazyEnablePeerAccess));
  return PN_OK;
}

}  // extern "C"

namespace pn {
__global__ void peer_signal_kernel(int* flag, int step, int* counter) {
  // step <= 0: the step number lives on the device (counter, advanced here), so that a captured CUDA graph can be replayed
  if (step <= 0) step = ++(*counter);
  // the compositing kernel's stores into the peer frame are complete (stream order); publish the step at system scope
  __threadfence_system();
  asm volatile("st.release.sys.global.s32 [%0], %1;" ::"l"(flag), "r"(step) : "memory");
}

__global__ void peer_wait_kernel(const int* flags, int n, int step, unsigned long long timeout_ns, int* status, int* counter) {
  const int i = threadIdx.x;
  if (step <= 0) {                                   // device-resident step number (graph replay): every lane reads, lane 0 advances
    step = *counter + 1;
    __syncwarp();
    if (i == 0) *counter = step;
  }
  if (i >= n) return;
  unsigned long long t0;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
  for (;;) {
    int v;
    asm volatile("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(flags + i) : "memory");
    if (v >= step) break;
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    if (t - t0 > timeout_ns) {                       // a rank is late or gone: report instead of hanging the GPU
      if (status) atomicExch(status, 1 + i);
      break;
    }
    __nanosleep(100);
  }
}
}  // namespace pn

extern "C" {

int pn_peer_signal(int* flag_dev, int step, int* counter_dev, pn_stream_t stream) {
  PN_REQUIRE(flag_dev && (step > 0 || counter_dev), "pn_peer_signal: flag is NULL, or step <= 0 without a device counter");
  peer_signal_kernel<<<1, 1, 0, as_stream(stream)>>>(flag_dev, step, counter_dev);
  PN_LAUNCH_OK("pn_peer_signal");
  return PN_OK;
}

int pn_peer_wait(const int* flags_dev, int n_flags, int step, int timeout_ms, int* status_dev, int* counter_dev, pn_stream_t stream) {
  PN_REQUIRE(flags_dev && n_flags >= 1 && n_flags <= 32 && timeout_ms > 0 && (step > 0 || counter_dev),
             "pn_peer_wait: bad arguments (n_flags=%d)", n_flags);
  peer_wait_kernel<<<1, 32, 0, as_stream(stream)>>>(flags_dev, n_flags, step, (unsigned long long)timeout_ms * 1000000ull, status_dev,
                                                    counter_dev);
  PN_LAUNCH_OK("pn_peer_wait");
  return PN_OK;
}

int pn_peer_close(void* dev_ptr) {
  if (!dev_ptr) return PN_OK;
  PN_CUDA_OK(cudaIpcCloseMemHandle(dev_ptr));
  return PN_OK;
}

int pn_peer_free(void* dev_ptr) {
  if (!dev_ptr) return PN_OK;
  PN_CUDA_OK(cudaFree(dev_ptr));
  return PN_OK;
}

}  // extern "C"
