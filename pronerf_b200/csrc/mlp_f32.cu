// Fused fp32 MLP (the <=1e-3 parity tier of the three networks).
//
//   sampler  MinMaxRaySamplerTRT_Net     helpers.py:1473-1507   288 -> 256 x6 (ELU) -> 27
//   refine   MinMaxRayEpiSamplerTRT_Net  helpers.py:1509-1540   144 -> 256 x6 (ELU) -> 35
//   NeRF     DoNeRFTRT                   helpers.py:1186-1343   63 -> 256 x7 (ReLU) -> [256 ++ 27] -> 4
//
// One CTA owns a tile of 64 rows for the whole network: the activations live in shared memory
// ([64][292] fp32, updated in place) and never touch HBM between layers; weights stream from L2 through
// a double-buffered cp.async pipeline in chunks of 16 k-rows x 256 outputs that runs straight across
// layer boundaries.  256 threads = 8 warps; warp w owns rows 8w..8w+7 (so activation reads are smem
// broadcasts and the in-place update needs no CTA barrier), lane l owns output columns 4l..4l+3 and
// 128+4l..128+4l+3 (conflict-free 128-bit weight reads and activation writes).  Each thread keeps an
// 8x8 accumulator tile: 64 FFMA per 2 weight LDS.128 + 2 activation LDS.128 (amortised).
//
// The first-layer operand is either loaded (drop-in module forwards) or generated in the kernel:
// positional encodings (helpers.py:666-671) for run_network, Pluecker ray features (trt.py:274-278)
// for the sampler -- neither ever exists in HBM on the fused path.
//
// This tier runs on the fp32 FMA pipes (~75 TFLOP/s peak on B200); the throughput tier is the bf16
// tcgen05 kernel in mlp_tc.cu.
#include "common.cuh"

namespace pn {

constexpr int TM = 64;            // rows per CTA
constexpr int LD = 292;           // activation row stride in floats (288 + 4; 16-byte aligned rows)
constexpr int KC = 16;            // k-rows per weight chunk
constexpr int NT = 256;           // threads per CTA
constexpr int OUT_PAD = kOutPad;       // padded width of the output layer (4S+3 = 67 at S = 16)
constexpr size_t kSmemBytes = (size_t)TM * LD * 4 + 2 * (size_t)KC * kHidden * 4;

struct MlpParams {
  const float* trunk;             // [trunk_rows][256]
  const float* wout;              // [k_out_pad][96]
  const float* bias;              // [n_layers][256]
  int n_trunk;
  int k_pad[kMaxLayers];
  int k0;                         // true input width
  int k_out_pad;                  // padded K of the output layer
  int n_out;                      // true output width
  int total_chunks;
  int act;
  int input_mode;
  const float* in0;
  const float* in1;
  int in_stride, in1_stride;
  int S, P;
  long long M;
  float* out;
  int head_lo[4];
  int head_act[3];
};

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
  unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

__device__ __forceinline__ void prefetch_chunk(float* wb, const float* trunk, int chunk, int tid) {
  const float4* src = reinterpret_cast<const float4*>(trunk + (size_t)chunk * KC * kHidden);
  float4* dst = reinterpret_cast<float4*>(wb);
#pragma unroll
  for (int i = 0; i < (KC * kHidden / 4) / NT; ++i) cp_async16(dst + tid + i * NT, src + tid + i * NT);
}

__device__ __forceinline__ float linspace01_(int i, int P) {
  float step = __fdiv_rn(1.f, (float)(P - 1));
  return (i < P / 2) ? __fmul_rn(step, (float)i) : __fmaf_rn(-step, (float)(P - 1 - i), 1.f);
}

__device__ __forceinline__ float head_apply(float v, int kind) {
  if (kind == HEAD_SIGMOID) return sigmoidf_(v);
  if (kind == HEAD_TANH) return tanhf(v);
  return v;
}

template <int ACT>
__global__ void __launch_bounds__(NT, 2) mlp_f32_kernel(MlpParams p) {
  extern __shared__ __align__(16) float smem[];
  float* act = smem;                       // [TM][LD]
  float* wbuf = smem + TM * LD;            // [2][KC][256]
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const long long row0 = (long long)blockIdx.x * TM;

  // start the weight stream before building the input tile
  prefetch_chunk(wbuf, p.trunk, 0, tid);
  cp_async_commit();

  // ---------------- first-layer operand ----------------
  const int k0p = p.k_pad[0];
  if (p.input_mode == IN_LOAD) {
    for (int idx = tid; idx < TM * k0p; idx += NT) {
      int r = idx / k0p, c = idx - r * k0p;
      long long row = row0 + r;
      act[r * LD + c] = (row < p.M && c < p.k0) ? p.in0[row * p.in_stride + c] : 0.f;
    }
  } else if (p.input_mode == IN_LOAD2) {
    for (int idx = tid; idx < TM * 64; idx += NT) {
      int r = idx >> 6, c = idx & 63;
      long long row = row0 + r;
      act[r * LD + c] = (row < p.M && c < 63) ? p.in0[row * p.in_stride + c] : 0.f;
    }
    for (int idx = tid; idx < TM * 32; idx += NT) {
      int r = idx >> 5, c = idx & 31;
      long long row = row0 + r;
      act[r * LD + 256 + c] = (row < p.M && c < 27) ? p.in1[row * 27 + c] : 0.f;
    }
  } else if (p.input_mode == IN_ENCODE) {
    // gamma_10(point) -> cols 0..62 (col 63 = 0); gamma_4(viewdir) -> cols 256..282 (283..287 = 0)
    for (int idx = tid; idx < TM * 3; idx += NT) {
      int r = idx / 3, c = idx - r * 3;
      long long row = row0 + r;
      float* a = act + r * LD;
      float x = 0.f, v = 0.f;
      if (row < p.M) { x = p.in0[row * 3 + c]; v = p.in1[(row / p.S) * p.in1_stride + c]; }
      a[c] = x;
      float f = 1.f;
#pragma unroll
      for (int l = 0; l < 10; ++l) {
        float s, co;
        sincosf(__fmul_rn(x, f), &s, &co);
        a[3 + 6 * l + c] = s;
        a[6 + 6 * l + c] = co;
        f *= 2.f;
      }
      a[256 + c] = v;
      f = 1.f;
#pragma unroll
      for (int l = 0; l < 4; ++l) {
        float s, co;
        sincosf(__fmul_rn(v, f), &s, &co);
        a[259 + 6 * l + c] = s;
        a[262 + 6 * l + c] = co;
        f *= 2.f;
      }
      if (c == 0) { a[63] = 0.f; a[283] = 0.f; a[284] = 0.f; a[285] = 0.f; a[286] = 0.f; a[287] = 0.f; }
    }
  } else {  // IN_PLUECKER: rays [N, stride] -> P x [normalize(d), (o + d t_p) x normalize(d)]
    const int P = p.P;
    for (int idx = tid; idx < TM * P; idx += NT) {
      int r = idx / P, pt = idx - r * P;
      long long row = row0 + r;
      float f6[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
      if (row < p.M) {
        const float* ray = p.in0 + row * p.in_stride;
        float tt = (P > 1) ? linspace01_(pt, P) : 0.f;
        float ox = __fadd_rn(ray[0], __fmul_rn(ray[3], tt));
        float oy = __fadd_rn(ray[1], __fmul_rn(ray[4], tt));
        float oz = __fadd_rn(ray[2], __fmul_rn(ray[5], tt));
        pluecker6(ox, oy, oz, ray[3], ray[4], ray[5], f6);
      }
#pragma unroll
      for (int i = 0; i < 6; ++i) act[r * LD + 6 * pt + i] = f6[i];
    }
    for (int idx = tid; idx < TM * (k0p - 6 * P); idx += NT) {      // zero the k padding
      int r = idx / (k0p - 6 * P), c = idx - r * (k0p - 6 * P);
      act[r * LD + 6 * P + c] = 0.f;
    }
  }
  // (first __syncthreads of the chunk loop publishes the tile)

  // ---------------- trunk: 256-wide layers ----------------
  float acc[8][8];
  float* arow = act + (warp * 8) * LD;
  int chunk = 0;
  for (int l = 0; l < p.n_trunk; ++l) {
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
    const int nch = p.k_pad[l] / KC;
    for (int kc = 0; kc < nch; ++kc, ++chunk) {
      cp_async_wait<0>();
      __syncthreads();                                   // chunk landed for everyone; previous buffer is free
      if (chunk + 1 < p.total_chunks) {
        prefetch_chunk(wbuf + ((chunk + 1) & 1) * KC * kHidden, p.trunk, chunk + 1, tid);
        cp_async_commit();
      }
      const float* wb = wbuf + (chunk & 1) * KC * kHidden;
      const float* ak = arow + kc * KC;
#pragma unroll
      for (int kk = 0; kk < KC; kk += 4) {
        float4 a[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) a[i] = *reinterpret_cast<const float4*>(ak + i * LD + kk);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          float4 w0 = *reinterpret_cast<const float4*>(wb + (kk + j) * kHidden + 4 * lane);
          float4 w1 = *reinterpret_cast<const float4*>(wb + (kk + j) * kHidden + 128 + 4 * lane);
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            float av = (j == 0) ? a[i].x : (j == 1) ? a[i].y : (j == 2) ? a[i].z : a[i].w;
            acc[i][0] = fmaf(av, w0.x, acc[i][0]);
            acc[i][1] = fmaf(av, w0.y, acc[i][1]);
            acc[i][2] = fmaf(av, w0.z, acc[i][2]);
            acc[i][3] = fmaf(av, w0.w, acc[i][3]);
            acc[i][4] = fmaf(av, w1.x, acc[i][4]);
            acc[i][5] = fmaf(av, w1.y, acc[i][5]);
            acc[i][6] = fmaf(av, w1.z, acc[i][6]);
            acc[i][7] = fmaf(av, w1.w, acc[i][7]);
          }
        }
      }
    }
    // bias + activation, in place (rows are private to this warp)
    __syncwarp();
    const float4 b0 = *reinterpret_cast<const float4*>(p.bias + l * kHidden + 4 * lane);
    const float4 b1 = *reinterpret_cast<const float4*>(p.bias + l * kHidden + 128 + 4 * lane);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      float4 o0 = make_float4(acc[i][0] + b0.x, acc[i][1] + b0.y, acc[i][2] + b0.z, acc[i][3] + b0.w);
      float4 o1 = make_float4(acc[i][4] + b1.x, acc[i][5] + b1.y, acc[i][6] + b1.z, acc[i][7] + b1.w);
      if (ACT == 0) {
        o0.x = fmaxf(o0.x, 0.f); o0.y = fmaxf(o0.y, 0.f); o0.z = fmaxf(o0.z, 0.f); o0.w = fmaxf(o0.w, 0.f);
        o1.x = fmaxf(o1.x, 0.f); o1.y = fmaxf(o1.y, 0.f); o1.z = fmaxf(o1.z, 0.f); o1.w = fmaxf(o1.w, 0.f);
      } else {
        o0.x = eluf_(o0.x); o0.y = eluf_(o0.y); o0.z = eluf_(o0.z); o0.w = eluf_(o0.w);
        o1.x = eluf_(o1.x); o1.y = eluf_(o1.y); o1.z = eluf_(o1.z); o1.w = eluf_(o1.w);
      }
      *reinterpret_cast<float4*>(arow + i * LD + 4 * lane) = o0;
      *reinterpret_cast<float4*>(arow + i * LD + 128 + 4 * lane) = o1;
    }
    __syncwarp();
  }

  // ---------------- output layer (<= 96 wide), weights straight from L1/L2 ----------------
  {
    constexpr int NH = OUT_PAD / 32;
    float o[8][NH];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int h = 0; h < NH; ++h) o[i][h] = 0.f;
    const int nh = (p.n_out + 31) / 32;
    const float* wo = p.wout;
    for (int k = 0; k < p.k_out_pad; k += 4) {
      float4 a[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) a[i] = *reinterpret_cast<const float4*>(arow + i * LD + k);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float w[NH];
#pragma unroll
        for (int h = 0; h < NH; ++h) w[h] = (h < nh) ? __ldg(wo + (k + j) * OUT_PAD + 32 * h + lane) : 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          float av = (j == 0) ? a[i].x : (j == 1) ? a[i].y : (j == 2) ? a[i].z : a[i].w;
#pragma unroll
          for (int h = 0; h < NH; ++h) o[i][h] = fmaf(av, w[h], o[i][h]);
        }
      }
    }
    const float* bo = p.bias + p.n_trunk * kHidden;
#pragma unroll
    for (int h = 0; h < NH; ++h) {
      int col = lane + 32 * h;
      if (col >= p.n_out) continue;
      int kind = HEAD_NONE;
#pragma unroll
      for (int g = 0; g < 3; ++g)
        if (col >= p.head_lo[g] && col < p.head_lo[g + 1]) kind = p.head_act[g];
      float b = bo[col];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        long long row = row0 + warp * 8 + i;
        if (row < p.M) p.out[row * p.n_out + col] = head_apply(o[i][h] + b, kind);
      }
    }
  }
}

// W [out][in] (nn.Linear) -> k-major [k_pad][n_pad], zero padded; bias -> [bias_pad]
__global__ void pack_f32_kernel(const float* __restrict__ W, const float* __restrict__ b, int out_dim, int in_dim,
                                int k_pad, int n_pad, float* __restrict__ wt, float* __restrict__ bias, int bias_pad) {
  int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx < k_pad * n_pad) {
    int k = idx / n_pad, o = idx - k * n_pad;
    wt[idx] = (k < in_dim && o < out_dim) ? W[(size_t)o * in_dim + k] : 0.f;
  }
  if (idx < bias_pad) bias[idx] = idx < out_dim ? b[idx] : 0.f;
}

int pack_layer_f32(const float* W, const float* b, int out_dim, int in_dim, int k_pad, int n_pad, float* wt,
                   float* bias, int bias_pad, cudaStream_t stream) {
  int total = k_pad * n_pad;
  if (total < bias_pad) total = bias_pad;
  pack_f32_kernel<<<(total + 255) / 256, 256, 0, stream>>>(W, b, out_dim, in_dim, k_pad, n_pad, wt, bias, bias_pad);
  PN_LAUNCH_OK("pack_f32_kernel");
  return PN_OK;
}

int launch_mlp_f32(const MlpLaunch& L, cudaStream_t stream) {
  const NetF32& n = *L.net;
  if (!n.loaded) { set_error("network weights not loaded (call pn_ctx_load_net first)"); return PN_ESTATE; }
  if (L.M == 0) return PN_OK;
  MlpParams p;
  p.trunk = n.trunk; p.wout = n.wout; p.bias = n.bias;
  p.n_trunk = n.n_layers - 1;
  for (int i = 0; i < kMaxLayers; ++i) p.k_pad[i] = n.k_pad[i];
  p.k0 = n.in_dim[0];
  p.k_out_pad = n.k_pad[n.n_layers - 1];
  p.n_out = n.out_dim[n.n_layers - 1];
  p.total_chunks = n.trunk_rows / KC;
  p.act = L.act; p.input_mode = L.input_mode; p.in0 = L.in0; p.in1 = L.in1; p.in_stride = L.in_stride; p.in1_stride = L.in1_stride;
  p.S = L.S; p.P = L.P; p.M = L.M; p.out = L.out;
  for (int i = 0; i < 4; ++i) p.head_lo[i] = L.head_lo[i];
  for (int i = 0; i < 3; ++i) p.head_act[i] = L.head_act[i];
  PN_CUDA_OK(cudaFuncSetAttribute(mlp_f32_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBytes));
  PN_CUDA_OK(cudaFuncSetAttribute(mlp_f32_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBytes));
  long long tiles = (L.M + TM - 1) / TM;
  if (tiles > 2147483647LL) { set_error("mlp_f32: too many rows"); return PN_EINVAL; }
  if (L.act == 0)
    mlp_f32_kernel<0><<<(unsigned)tiles, NT, kSmemBytes, stream>>>(p);
  else
    mlp_f32_kernel<1><<<(unsigned)tiles, NT, kSmemBytes, stream>>>(p);
  PN_LAUNCH_OK("mlp_f32_kernel");
  return PN_OK;
}

}  // namespace pn
