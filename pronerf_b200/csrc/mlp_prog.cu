// fp32 "layer program" MLP: the parity tier of network topologies the fused kernels (mlp_f32.cu / mlp_tc.cu) do not
// cover.  First user: the classic NeRF of stage-2 checkpoints (run_nerf_helpers.py:792-847),
//
//   63 -> 256 x5 (ReLU) -> cat(pts_enc, h) -> 256 x3 (ReLU) -> { alpha 256->1 ; feature 256->256 (linear) }
//   -> cat(feature, dirs_enc) -> 128 (ReLU) -> rgb 3            out = [rgb, alpha]
//
// which the reference's own infer script cannot load (it builds DoNeRFTRT and strict-loads a NeRF state_dict: defect Q7).
//
// One CTA owns 32 rows.  Its activation tile lives in shared memory as named regions -- encoded points, encoded view
// directions, two ping-pong 256-wide buffers, the 4 outputs -- and the network is a list of steps
//   out_region[r][n] = act(b[n] + sum_k in_seg0[r][k] W[k][n] + sum_k in_seg1[r][k] W[K0 + k][n])
// with up to two input segments (the skip / view concatenations become a second segment instead of a copy).  Thread n owns
// output column n for all 32 rows (32 accumulators); weights are k-major [K][N] in global memory (coalesced across
// threads, L2-resident), activations are shared-memory broadcasts.  FMA order over k is sequential, like the fused fp32
// tier.  Speed is not the point of this tier (no tensor cores); the tcgen05 form of this topology is described in
// DESIGN.md section 4.3.
#include <vector>

#include "common.cuh"

namespace pn {

constexpr int PG_TM = 32;                 // rows per CTA
constexpr int PG_NT = 256;                // threads per CTA = widest layer
// activation regions (float offsets inside one row)
constexpr int PG_PTS = 0;                 // 64: encoded point (63 + zero pad)
constexpr int PG_DIR = 64;                // 32: encoded view direction (27 + zero pad)
constexpr int PG_A = 96;                  // 256: ping
constexpr int PG_B = 352;                 // 256: pong
constexpr int PG_OUT = 608;               // 4: final outputs
constexpr int PG_LD = 612;
constexpr size_t kProgSmem = (size_t)PG_TM * PG_LD * sizeof(float);

struct ProgParams {
  int n_steps;
  ProgStep st[kMaxProgSteps];
  const float* blob;
  int input_mode;                         // IN_ENCODE (pts + per-ray view dirs) or IN_LOAD2 (embedded, embedded_dirs)
  const float* in0;
  const float* in1;
  int in1_stride, S;
  long long M;
  float* out;                             // [M, 4]
};

__global__ void __launch_bounds__(PG_NT) mlp_prog_kernel(ProgParams p) {
  extern __shared__ __align__(16) float tile[];
  const int tid = threadIdx.x;
  const long long row0 = (long long)blockIdx.x * PG_TM;
  // ---- inputs: encoded point -> PG_PTS, encoded view direction -> PG_DIR (helpers.py:666-671) ----
  if (p.input_mode == IN_ENCODE) {
    for (int idx = tid; idx < PG_TM * 3; idx += PG_NT) {
      const int r = idx / 3, c = idx - r * 3;
      const long long row = row0 + r;
      float* a = tile + r * PG_LD;
      float x = 0.f, v = 0.f;
      if (row < p.M) { x = p.in0[row * 3 + c]; v = p.in1[(row / p.S) * p.in1_stride + c]; }
      a[PG_PTS + c] = x;
      float f = 1.f;
      for (int l = 0; l < 10; ++l) {
        float s, co;
        sincosf(__fmul_rn(x, f), &s, &co);
        a[PG_PTS + 3 + 6 * l + c] = s;
        a[PG_PTS + 6 + 6 * l + c] = co;
        f *= 2.f;
      }
      a[PG_DIR + c] = v;
      f = 1.f;
      for (int l = 0; l < 4; ++l) {
        float s, co;
        sincosf(__fmul_rn(v, f), &s, &co);
        a[PG_DIR + 3 + 6 * l + c] = s;
        a[PG_DIR + 6 + 6 * l + c] = co;
        f *= 2.f;
      }
      if (c == 0) {
        a[PG_PTS + 63] = 0.f;
        for (int j = 27; j < 32; ++j) a[PG_DIR + j] = 0.f;
      }
    }
  } else {
    for (int idx = tid; idx < PG_TM * 64; idx += PG_NT) {
      const int r = idx >> 6, c = idx & 63;
      const long long row = row0 + r;
      tile[r * PG_LD + PG_PTS + c] = (row < p.M && c < 63) ? p.in0[row * 63 + c] : 0.f;
    }
    for (int idx = tid; idx < PG_TM * 32; idx += PG_NT) {
      const int r = idx >> 5, c = idx & 31;
      const long long row = row0 + r;
      tile[r * PG_LD + PG_DIR + c] = (row < p.M && c < 27) ? p.in1[row * 27 + c] : 0.f;
    }
  }
  __syncthreads();
  // ---- the program ----
  for (int s = 0; s < p.n_steps; ++s) {
    const ProgStep st = p.st[s];
    if (tid < st.n_out) {
      float acc[PG_TM];
#pragma unroll
      for (int r = 0; r < PG_TM; ++r) acc[r] = 0.f;
      const float* W = p.blob + st.w_off + tid;
      for (int seg = 0; seg < 2; ++seg) {
        const int off = seg ? st.in1_off : st.in0_off, len = seg ? st.in1_len : st.in0_len;
        const float* a = tile + off;
        for (int k = 0; k < len; ++k) {
          const float w = __ldg(W);
          W += st.n_out;
#pragma unroll
          for (int r = 0; r < PG_TM; ++r) acc[r] = fmaf(a[r * PG_LD + k], w, acc[r]);
        }
      }
      const float b = p.blob[st.b_off + tid];
#pragma unroll
      for (int r = 0; r < PG_TM; ++r) {
        float y = acc[r] + b;
        if (st.act == 1) y = fmaxf(y, 0.f);
        tile[r * PG_LD + st.out_off + tid] = y;
      }
    }
    __syncthreads();
  }
  // ---- outputs: [rgb, alpha] ----
  for (int idx = tid; idx < PG_TM * 4; idx += PG_NT) {
    const int r = idx >> 2, c = idx & 3;
    const long long row = row0 + r;
    if (row < p.M) p.out[row * 4 + c] = tile[r * PG_LD + PG_OUT + c];
  }
}

// W [out][in] -> k-major [k_len][out] rows k_dst0.. taken from input columns k_src0.. (zero rows beyond k_used)
__global__ void pack_prog_kernel(const float* __restrict__ W, int out_dim, int in_dim, int k_src0, int k_used, int k_len,
                                 float* __restrict__ dst) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= k_len * out_dim) return;
  const int k = idx / out_dim, n = idx - k * out_dim;
  dst[idx] = k < k_used ? W[(size_t)n * in_dim + k_src0 + k] : 0.f;
}

void prog_free(NetProg& n) {
  if (n.blob) cudaFree(n.blob);
  n = NetProg();
}

// The classic NeRF (helpers.py:792-847) with D = 8, W = 256, skips = [4], use_viewdirs, input_ch = 63, input_ch_views = 27.
// W/b order: pts_linears.0..7, alpha_linear, feature_linear, views_linears.0, rgb_linear (the checkpoint's key order).
int prog_load_nerf_classic(NetProg& net, const int* in_dims, const int* out_dims, const float* const* W, const float* const* b,
                           cudaStream_t stream) {
  prog_free(net);
  const int Wd = kHidden;
  const int want_in[12] = {63, Wd, Wd, Wd, Wd, Wd + 63, Wd, Wd, Wd, Wd, Wd + 27, Wd / 2};
  const int want_out[12] = {Wd, Wd, Wd, Wd, Wd, Wd, Wd, Wd, 1, Wd, Wd / 2, 3};
  for (int l = 0; l < 12; ++l)
    PN_REQUIRE(in_dims[l] == want_in[l] && out_dims[l] == want_out[l],
               "classic NeRF: tensor %d is [%d, %d]; this build covers D=8, W=256, skips=[4], use_viewdirs, multires 10/4 "
               "(expected [%d, %d])", l, out_dims[l], in_dims[l], want_out[l], want_in[l]);
  // steps: (layer, segment 0 region/len/source column, segment 1 region/len/source column, out region, act)
  struct Seg { int region, len, src0, used; };
  struct Def { int layer; Seg s0, s1; int out_off, act; };
  const Seg none{0, 0, 0, 0};
  const Def defs[12] = {
      {0, {PG_PTS, 64, 0, 63}, none, PG_A, 1},
      {1, {PG_A, Wd, 0, Wd}, none, PG_B, 1},
      {2, {PG_B, Wd, 0, Wd}, none, PG_A, 1},
      {3, {PG_A, Wd, 0, Wd}, none, PG_B, 1},
      {4, {PG_B, Wd, 0, Wd}, none, PG_A, 1},
      {5, {PG_PTS, 64, 0, 63}, {PG_A, Wd, 63, Wd}, PG_B, 1},          // h = cat([input_pts, h])  (helpers.py:833-834)
      {6, {PG_B, Wd, 0, Wd}, none, PG_A, 1},
      {7, {PG_A, Wd, 0, Wd}, none, PG_B, 1},
      {8, {PG_B, Wd, 0, Wd}, none, PG_OUT + 3, 0},                    // alpha_linear
      {9, {PG_B, Wd, 0, Wd}, none, PG_A, 0},                          // feature_linear (no activation)
      {10, {PG_A, Wd, 0, Wd}, {PG_DIR, 32, Wd, 27}, PG_B, 1},         // cat([feature, input_views]) -> views_linears.0
      {11, {PG_B, Wd / 2, 0, Wd / 2}, none, PG_OUT, 0},               // rgb_linear
  };
  size_t floats = 0;
  for (int s = 0; s < 12; ++s) floats += (size_t)(defs[s].s0.len + defs[s].s1.len) * out_dims[defs[s].layer] + out_dims[defs[s].layer];
  PN_CUDA_OK(cudaMalloc((void**)&net.blob, floats * sizeof(float)));
  size_t off = 0;
  for (int s = 0; s < 12; ++s) {
    const Def& d = defs[s];
    const int l = d.layer, n_out = out_dims[l];
    ProgStep& st = net.st[s];
    st.in0_off = d.s0.region; st.in0_len = d.s0.len; st.in1_off = d.s1.region; st.in1_len = d.s1.len;
    st.out_off = d.out_off; st.n_out = n_out; st.act = d.act; st.w_off = (int)off;
    const Seg* segs[2] = {&d.s0, &d.s1};
    for (int g = 0; g < 2; ++g) {
      const Seg& sg = *segs[g];
      if (sg.len == 0) continue;
      const int total = sg.len * n_out;
      pack_prog_kernel<<<(total + 255) / 256, 256, 0, stream>>>(W[l], n_out, in_dims[l], sg.src0, sg.used, sg.len, net.blob + off);
      PN_LAUNCH_OK("pack_prog_kernel");
      off += (size_t)total;
    }
    st.b_off = (int)off;
    PN_CUDA_OK(cudaMemcpyAsync(net.blob + off, b[l], (size_t)n_out * sizeof(float), cudaMemcpyDeviceToDevice, stream));
    off += (size_t)n_out;
  }
  net.n_steps = 12;
  net.loaded = true;
  return PN_OK;
}

int prog_launch(const NetProg& net, int input_mode, const float* in0, const float* in1, int in1_stride, int S, int64_t M,
                float* out, cudaStream_t stream) {
  if (!net.loaded) { set_error("classic NeRF weights not loaded (pn_ctx_load_nerf_classic)"); return PN_ESTATE; }
  if (M == 0) return PN_OK;
  ProgParams p;
  p.n_steps = net.n_steps;
  for (int s = 0; s < net.n_steps; ++s) p.st[s] = net.st[s];
  p.blob = net.blob; p.input_mode = input_mode; p.in0 = in0; p.in1 = in1; p.in1_stride = in1_stride; p.S = S > 0 ? S : 1;
  p.M = M; p.out = out;
  PN_CUDA_OK(cudaFuncSetAttribute(mlp_prog_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kProgSmem));
  const long long tiles = (M + PG_TM - 1) / PG_TM;
  if (tiles > 2147483647LL) { set_error("mlp_prog: too many rows"); return PN_EINVAL; }
  mlp_prog_kernel<<<(unsigned)tiles, PG_NT, kProgSmem, stream>>>(p);
  PN_LAUNCH_OK("mlp_prog_kernel");
  return PN_OK;
}

}  // namespace pn
