// Projection-aware colour gather (inverse_warp.py:584-619 + trt.py:649-655).
//
// For every ray, sample s and neighbour view k:  w = ro + rd * depth3d  ->  p = M_k w (3x4)  ->
// X = p.x/p.z, Y = p.y/p.z  ->  normalise / un-normalise round trip of grid_sample(align_corners=True)
// ->  bilinear fetch with zero padding.  This is the only data-dependent gather of the path; it is
// bound by memory (L2/HBM), not arithmetic.
//
// Bit-exactness contract: floor(ix), floor(iy) must equal the reference's.  The reference's arithmetic
// as PyTorch executes it on CPU is restated op for op, every rounding pinned with __f*_rn intrinsics:
//   w_c   = fl(ro_c + fl(rd_c * depth))                          (separate mul and add, iw.py:600)
//   p_r   = fma(m_r3, w_3, fma(m_r2, w_2, fma(m_r1, w_1, fl(m_r0 * w_0))))   (MKL sgemm k-loop, iw.py:601)
//   X     = fl(p_0 / p_2),  Y = fl(p_1 / p_2)                    (iw.py:603)
//   Xn    = fl(fl(fl(2 X) / (W-1)) - 1)                          (true division, iw.py:607-608)
//   ix    = fl(fl(Xn + 1) * ((W-1)/2))                           (grid_sample un-normalise)
//   x0    = floor(ix);  w = ix - x0;  e = 1 - w   (same for y: n, s)
//   out_c = nw*I[y0][x0] + ne*I[y0][x0+1] + sw*I[y0+1][x0] + se*I[y0+1][x0+1],  taps outside -> 0
//
// Data layout: the fast path reads RGBA fp32 texels [NN][H][W] (16 B each, from pn_pack_images), so one
// tap is one 128-bit read-only load and the two taps of a row are 32 contiguous bytes; the four
// reference views (12 MB at 504x378) stay L2-resident.  One thread handles one (ray, sample) and loops
// over the NN views so the ray and depth are loaded once; the 3*NN results of a thread are written
// straight into the refine network's input row.
#include <cuda_fp16.h>

#include "common.cuh"

namespace pn {

struct TexIndex { int v[8]; };
struct TexIndexViews { int v[kMaxViews][8]; };

struct Tap {
  float ix, iy;
  float x0f, y0f;
};

__device__ __forceinline__ Tap project_point(const float* __restrict__ M, float w0, float w1, float w2, float w3,
                                             float wm1, float hm1, float wh, float hh) {
  float p0 = __fmaf_rn(M[3], w3, __fmaf_rn(M[2], w2, __fmaf_rn(M[1], w1, __fmul_rn(M[0], w0))));
  float p1 = __fmaf_rn(M[7], w3, __fmaf_rn(M[6], w2, __fmaf_rn(M[5], w1, __fmul_rn(M[4], w0))));
  float p2 = __fmaf_rn(M[11], w3, __fmaf_rn(M[10], w2, __fmaf_rn(M[9], w1, __fmul_rn(M[8], w0))));
  float X = __fdiv_rn(p0, p2);
  float Y = __fdiv_rn(p1, p2);
  float Xn = __fsub_rn(__fdiv_rn(__fmul_rn(2.f, X), wm1), 1.f);
  float Yn = __fsub_rn(__fdiv_rn(__fmul_rn(2.f, Y), hm1), 1.f);
  Tap t;
  t.ix = __fmul_rn(__fadd_rn(Xn, 1.f), wh);
  t.iy = __fmul_rn(__fadd_rn(Yn, 1.f), hh);
  t.x0f = floorf(t.ix);
  t.y0f = floorf(t.iy);
  return t;
}

__device__ __forceinline__ int32_t clamp_index(float f) {
  const float big = 1073741824.f;   // 2^30
  if (!(fabsf(f) <= 3.0e38f)) return -1073741824;        // NaN / inf
  return (int32_t)fminf(fmaxf(f, -big), big);
}

// Bilinear weights and validity of the four taps (float comparisons: NaN fails them all).
struct Bilin {
  float nw, ne, sw, se;
  bool vx0, vx1, vy0, vy1;
  int x0, y0;
};

__device__ __forceinline__ Bilin bilinear_setup(const Tap& t, int W, int H) {
  Bilin b;
  float w = __fsub_rn(t.ix, t.x0f), e = __fsub_rn(1.f, w);
  float n = __fsub_rn(t.iy, t.y0f), s = __fsub_rn(1.f, n);
  b.nw = __fmul_rn(s, e); b.ne = __fmul_rn(s, w); b.sw = __fmul_rn(n, e); b.se = __fmul_rn(n, w);
  float x1f = t.x0f + 1.f, y1f = t.y0f + 1.f;
  b.vx0 = (t.x0f >= 0.f) && (t.x0f <= (float)(W - 1));
  b.vx1 = (x1f >= 0.f) && (x1f <= (float)(W - 1));
  b.vy0 = (t.y0f >= 0.f) && (t.y0f <= (float)(H - 1));
  b.vy1 = (y1f >= 0.f) && (y1f <= (float)(H - 1));
  // only dereferenced when valid, so the clamp just keeps the conversion defined
  b.x0 = (int)fminf(fmaxf(t.x0f, -2.f), (float)W);
  b.y0 = (int)fminf(fmaxf(t.y0f, -2.f), (float)H);
  return b;
}

// ------------------------------------------------------------------------------------------------
// Fast path: RGBA texels, one thread per (ray, sample), loop over neighbours.
template <int NN_T>
__global__ void __launch_bounds__(256)
project_gather_kernel(const float4* __restrict__ texels, TexIndex tex, int NNr, int H, int W, const float* __restrict__ pm,
                      const float* __restrict__ ro_w, const float* __restrict__ rd_w, int rs, const float* __restrict__ depth3d,
                      int64_t N, int S, float* __restrict__ epi, int epi_stride, int epi_col0,
                      int32_t* __restrict__ x0y0) {
  __shared__ float sM[8 * 12];
  const int NN = NN_T > 0 ? NN_T : NNr;
  for (int i = threadIdx.x; i < NN * 12; i += blockDim.x) sM[i] = pm[i];
  __syncthreads();
  int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= N * S) return;
  int64_t r = t / S;
  int s = (int)(t - r * S);
  float d = depth3d[t];
  float w0 = __fadd_rn(ro_w[rs * r], __fmul_rn(rd_w[rs * r], d));
  float w1 = __fadd_rn(ro_w[rs * r + 1], __fmul_rn(rd_w[rs * r + 1], d));
  float w2 = __fadd_rn(ro_w[rs * r + 2], __fmul_rn(rd_w[rs * r + 2], d));
  float w3 = __fadd_rn(1.f, __fmul_rn(0.f, d));            // ro1[3] = 1, rd1[3] = 0   (trt.py:258-260)
  const float wm1 = (float)(W - 1), hm1 = (float)(H - 1);
  const float wh = wm1 / 2.f, hh = hm1 / 2.f;
  float* orow = epi + r * epi_stride + epi_col0;
#pragma unroll
  for (int k = 0; k < (NN_T > 0 ? NN_T : 8); ++k) {
    if (k >= NN) break;
    Tap tp = project_point(sM + 12 * k, w0, w1, w2, w3, wm1, hm1, wh, hh);
    Bilin b = bilinear_setup(tp, W, H);
    const float4* img = texels + (int64_t)tex.v[k] * H * W;
    float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
    const float4* row0 = img + (int64_t)b.y0 * W + b.x0;
    const float4* row1 = row0 + W;
    float4 c00 = (b.vx0 && b.vy0) ? __ldg(row0) : z4;
    float4 c01 = (b.vx1 && b.vy0) ? __ldg(row0 + 1) : z4;
    float4 c10 = (b.vx0 && b.vy1) ? __ldg(row1) : z4;
    float4 c11 = (b.vx1 && b.vy1) ? __ldg(row1 + 1) : z4;
    float nw = (b.vx0 && b.vy0) ? b.nw : 0.f, ne = (b.vx1 && b.vy0) ? b.ne : 0.f;
    float sw = (b.vx0 && b.vy1) ? b.sw : 0.f, se = (b.vx1 && b.vy1) ? b.se : 0.f;
    float* o = orow + (k * S + s) * 3;
    o[0] = c00.x * nw + c01.x * ne + c10.x * sw + c11.x * se;
    o[1] = c00.y * nw + c01.y * ne + c10.y * sw + c11.y * se;
    o[2] = c00.z * nw + c01.z * ne + c10.z * sw + c11.z * se;
    if (x0y0) {
      int64_t b_idx = ((int64_t)(k * S + s) * N + r) * 2;
      x0y0[b_idx] = clamp_index(tp.x0f);
      x0y0[b_idx + 1] = clamp_index(tp.y0f);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Fused form for the fp16 tensor-core tier: trt.py:631-661 in ONE pass over the sampler heads.
//   scale + stable sort + gather add/mul + depth lift (631-637)  ->  per-sample Pluecker features (656-658)
//   ->  projection + bilinear fetch (649-655)  ->  refine_input row [6S + 3*NN*S] written as fp16.
// One thread per (ray, sample); the S lanes of a ray rank their depths with shuffles (rank = position in torch's
// stable ascending sort, NaN last), so thread (ray, i) simply becomes sorted sample rho(i).  The refine-input rows
// of a block (256/S rays, contiguous in memory) are staged in shared memory and leave as 16-byte coalesced stores;
// the fp16 rows are what the refine MLP's first-layer operand consumes (mlp_tc.cu, IN_LOAD16), chunk for chunk.
// Same projection arithmetic as project_gather_kernel: floor(ix), floor(iy) stay bit-exact.
__device__ __forceinline__ bool sort_after(float a, float b) { return (a > b) || (isnan(a) && !isnan(b)); }

template <int S, int NN_T>
__global__ void __launch_bounds__(256)
refine_input_kernel(const float* __restrict__ heads, int head_stride, const float* __restrict__ rays,
                    const float* __restrict__ or_rays, int rs, const float4* __restrict__ texels, TexIndexViews tex, int NNr,
                    int H, int W, const float* __restrict__ pm, int n_views, int64_t rays_per_view, int64_t N,
                    float* __restrict__ depth, float* __restrict__ add, float* __restrict__ mul, __half* __restrict__ rin,
                    int32_t* __restrict__ x0y0, int64_t ray_base) {
  constexpr int RPB = 256 / S;                              // rays per block
  constexpr int KMAX = 6 * S + 3 * 8 * S;
  __shared__ float sM[kMaxViews * 8 * 12];
  __shared__ __align__(16) __half stage[RPB * KMAX];
  const int NN = NN_T > 0 ? NN_T : NNr;
  const int K0 = 6 * S + 3 * NN * S;
  pdl_wait();                                               // the sampler MLP's heads (and, first in a pass, the previous pass) are done
  pdl_launch();
  for (int i = threadIdx.x; i < n_views * NN * 12; i += blockDim.x) sM[i] = pm[i];
  __syncthreads();
  const int64_t ray0 = (int64_t)blockIdx.x * RPB;
  const int rl = threadIdx.x / S, i = threadIdx.x % S;
  const int64_t r_raw = ray0 + rl;
  const bool live = r_raw < N;
  const int64_t r = live ? r_raw : N - 1;
  const float* h = heads + r * head_stride;
  const float* ray = rays + r * rs;
  // multi-view batches: rays of view `view` occupy rows [view * rays_per_view, +rays_per_view); each view has its own
  // neighbour ordering (texel indices) and projection matrices (trt.py:281-294)
  int view = 0;
  // (ray_base: this launch covers rows [ray_base, ray_base + N) of the batch -- chunked passes of pn_render_views_host)
  if (n_views > 1) { view = (int)((r + ray_base) / rays_per_view); view = view < n_views ? view : n_views - 1; }
  const float* vM = sM + view * NN * 12;
  const float near_ = ray[6], far_ = ray[7];
  const float v = __fadd_rn(__fmul_rn(h[i], __fsub_rn(far_, near_)), near_);   // depth * (far - near) + near   trt.py:631
  int rho = 0;
#pragma unroll
  for (int j = 0; j < S; ++j) {
    const float vj = __shfl_sync(0xffffffffu, v, j, S);
    rho += (sort_after(v, vj) || (!sort_after(vj, v) && j < i)) ? 1 : 0;
  }
  if (live) {
    depth[r * S + rho] = v;
    add[r * S + rho] = h[S + i];
    mul[r * S + rho] = h[2 * S + i];
  }
  __half* srow = stage + rl * K0;
  {
    // Pluecker features of (o + d * depth, d)   trt.py:656-658
    float f[6];
    pluecker6(__fadd_rn(ray[0], __fmul_rn(ray[3], v)), __fadd_rn(ray[1], __fmul_rn(ray[4], v)),
              __fadd_rn(ray[2], __fmul_rn(ray[5], v)), ray[3], ray[4], ray[5], f);
#pragma unroll
    for (int j = 0; j < 6; ++j) srow[6 * rho + j] = __float2half_rn(f[j]);
  }
  // lift (trt.py:637) and project into the NN neighbour views
  const float d3 = __fdiv_rn(1.f, __fsub_rn(__fsub_rn(1.f, v), 1e-5f));
  const float* orr = or_rays + r * rs;
  const float w0 = __fadd_rn(orr[0], __fmul_rn(orr[3], d3));
  const float w1 = __fadd_rn(orr[1], __fmul_rn(orr[4], d3));
  const float w2 = __fadd_rn(orr[2], __fmul_rn(orr[5], d3));
  const float w3 = __fadd_rn(1.f, __fmul_rn(0.f, d3));
  const float wm1 = (float)(W - 1), hm1 = (float)(H - 1);
  const float wh = wm1 / 2.f, hh = hm1 / 2.f;
#pragma unroll
  for (int k = 0; k < (NN_T > 0 ? NN_T : 8); ++k) {
    if (k >= NN) break;
    Tap tp = project_point(vM + 12 * k, w0, w1, w2, w3, wm1, hm1, wh, hh);
    Bilin b = bilinear_setup(tp, W, H);
    const float4* img = texels + (int64_t)tex.v[view][k] * H * W;
    const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
    const float4* row0 = img + (int64_t)b.y0 * W + b.x0;
    const float4* row1 = row0 + W;
    const bool v00 = b.vx0 && b.vy0, v01 = b.vx1 && b.vy0, v10 = b.vx0 && b.vy1, v11 = b.vx1 && b.vy1;
    float4 c00 = v00 ? __ldg(row0) : z4;
    float4 c01 = v01 ? __ldg(row0 + 1) : z4;
    float4 c10 = v10 ? __ldg(row1) : z4;
    float4 c11 = v11 ? __ldg(row1 + 1) : z4;
    float nw = v00 ? b.nw : 0.f, ne = v01 ? b.ne : 0.f, sw = v10 ? b.sw : 0.f, se = v11 ? b.se : 0.f;
    __half* o = srow + 6 * S + (k * S + rho) * 3;
    o[0] = __float2half_rn(c00.x * nw + c01.x * ne + c10.x * sw + c11.x * se);
    o[1] = __float2half_rn(c00.y * nw + c01.y * ne + c10.y * sw + c11.y * se);
    o[2] = __float2half_rn(c00.z * nw + c01.z * ne + c10.z * sw + c11.z * se);
    if (x0y0 && live) {
      int64_t b_idx = ((int64_t)(k * S + rho) * N + r) * 2;
      x0y0[b_idx] = clamp_index(tp.x0f);
      x0y0[b_idx + 1] = clamp_index(tp.y0f);
    }
  }
  __syncthreads();
  // coalesced copy-out of the block's rows (contiguous in global memory; K0 * 2 bytes per row is a multiple of 16)
  const int64_t live_rays = (N - ray0) < RPB ? (N - ray0) : RPB;
  const int halves = (int)(live_rays * K0);
  const int chunks = halves / 8;
  const uint4* s4 = reinterpret_cast<const uint4*>(stage);
  uint4* g4 = reinterpret_cast<uint4*>(rin + ray0 * K0);
  for (int c = threadIdx.x; c < chunks; c += blockDim.x) g4[c] = s4[c];
  // rows of K0 % 8 != 0 halves (S = 4 with an odd neighbour count): the last block's tail is not a whole 16-byte chunk
  for (int e = chunks * 8 + threadIdx.x; e < halves; e += blockDim.x) rin[ray0 * K0 + e] = stage[e];
}

// ------------------------------------------------------------------------------------------------
// Drop-in form: planar img [B,C,H,W], per-batch depth [B,N], rays [*,4,N] with a batch stride, w2c [B,3,4].
__global__ void __launch_bounds__(256)
warp_kernel(const float* __restrict__ img, int B, int C, int H, int W, const float* __restrict__ depth,
            const float* __restrict__ ro1, const float* __restrict__ rd1, int64_t ro_bstride,
            const float* __restrict__ w2c, int64_t N, float* __restrict__ out, int32_t* __restrict__ x0y0) {
  int b = blockIdx.y;
  int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  const float* ro = ro1 + b * ro_bstride;
  const float* rd = rd1 + b * ro_bstride;
  float d = depth[(int64_t)b * N + n];
  float w0 = __fadd_rn(ro[n], __fmul_rn(rd[n], d));
  float w1 = __fadd_rn(ro[N + n], __fmul_rn(rd[N + n], d));
  float w2 = __fadd_rn(ro[2 * N + n], __fmul_rn(rd[2 * N + n], d));
  float w3 = __fadd_rn(ro[3 * N + n], __fmul_rn(rd[3 * N + n], d));
  float M[12];
#pragma unroll
  for (int i = 0; i < 12; ++i) M[i] = w2c[b * 12 + i];
  const float wm1 = (float)(W - 1), hm1 = (float)(H - 1);
  Tap tp = project_point(M, w0, w1, w2, w3, wm1, hm1, wm1 / 2.f, hm1 / 2.f);
  Bilin bl = bilinear_setup(tp, W, H);
  bool v00 = bl.vx0 && bl.vy0, v01 = bl.vx1 && bl.vy0, v10 = bl.vx0 && bl.vy1, v11 = bl.vx1 && bl.vy1;
  for (int c = 0; c < C; ++c) {
    const float* pl = img + ((int64_t)b * C + c) * H * W + (int64_t)bl.y0 * W + bl.x0;
    float a = 0.f;
    a += v00 ? __ldg(pl) * bl.nw : 0.f;
    a += v01 ? __ldg(pl + 1) * bl.ne : 0.f;
    a += v10 ? __ldg(pl + W) * bl.sw : 0.f;
    a += v11 ? __ldg(pl + W + 1) * bl.se : 0.f;
    out[((int64_t)b * C + c) * N + n] = a;
  }
  if (x0y0) {
    x0y0[((int64_t)b * N + n) * 2] = clamp_index(tp.x0f);
    x0y0[((int64_t)b * N + n) * 2 + 1] = clamp_index(tp.y0f);
  }
}

// ------------------------------------------------------------------------------------------------
// Stage-2 TRAINING warp (inverse_warp.py:515-581, inverse_warp_rod1_rt2_coords; scale = 1, zeros padding): the source pose is
// inverted here (R' w - R' t), the projection divides by |z| + 1e-8 and flips y, pixels whose normalised coordinate leaves
// [-1, 1] are pushed out of the image before grid_sample.  Op order as PyTorch executes it on CPU (oracle.warp_train):
//   t_r = -((R[0][r] t0 + R[1][r] t1) + R[2][r] t2)              (plain multiply-adds: a [3,3]x[3,1] product)
//   c_r = fma(R[2][r], w2, fma(R[1][r], w1, R[0][r] w0)) + t_r   (FMA chain over k)
//   c_  = c / (|c_z| + 1e-8), c__z = 1, c__y = -c__y;  p = K c_  (FMA chain);  X = p_0, Y = p_1
__global__ void __launch_bounds__(256)
warp_train_kernel(const float* __restrict__ img, int B, int C, int H, int W, const float* __restrict__ depth,
                  const float* __restrict__ ro1, const float* __restrict__ rd1, int64_t ro_bstride, const float* __restrict__ c2w,
                  const float* __restrict__ Kmat, int64_t N, float* __restrict__ out, int32_t* __restrict__ x0y0) {
  int b = blockIdx.y;
  int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  const float* ro = ro1 + b * ro_bstride;
  const float* rd = rd1 + b * ro_bstride;
  const float* P = c2w + b * 12;      // [3][4]: R = P[r][0..2], t = P[r][3]
  const float* Kb = Kmat + b * 9;
  float d = depth[(int64_t)b * N + n];
  float w[3], c[3];
#pragma unroll
  for (int i = 0; i < 3; ++i) w[i] = __fadd_rn(ro[i * N + n], __fmul_rn(rd[i * N + n], d));
#pragma unroll
  for (int r = 0; r < 3; ++r) {
    float t = -__fadd_rn(__fadd_rn(__fmul_rn(P[0 * 4 + r], P[0 * 4 + 3]), __fmul_rn(P[1 * 4 + r], P[1 * 4 + 3])), __fmul_rn(P[2 * 4 + r], P[2 * 4 + 3]));
    float acc = __fmaf_rn(P[2 * 4 + r], w[2], __fmaf_rn(P[1 * 4 + r], w[1], __fmul_rn(P[0 * 4 + r], w[0])));
    c[r] = __fadd_rn(acc, t);
  }
  float z = __fadd_rn(fabsf(c[2]), 1e-8f);
  float cx = __fdiv_rn(c[0], z), cy = -__fdiv_rn(c[1], z), cz = 1.f;
  float X = __fmaf_rn(Kb[2], cz, __fmaf_rn(Kb[1], cy, __fmul_rn(Kb[0], cx)));
  float Y = __fmaf_rn(Kb[5], cz, __fmaf_rn(Kb[4], cy, __fmul_rn(Kb[3], cx)));
  const float wm1 = (float)(W - 1), hm1 = (float)(H - 1);
  float Xn = __fsub_rn(__fdiv_rn(__fmul_rn(2.f, X), wm1), 1.f);
  float Yn = __fsub_rn(__fdiv_rn(__fmul_rn(2.f, Y), hm1), 1.f);
  if (Xn > 1.f || Xn < -1.f) Xn = 2.f;            // iw.py:561-565 (NaN compares false and stays NaN, like the reference)
  if (Yn > 1.f || Yn < -1.f) Yn = 2.f;
  Tap tp;
  tp.ix = __fmul_rn(__fadd_rn(Xn, 1.f), wm1 / 2.f);
  tp.iy = __fmul_rn(__fadd_rn(Yn, 1.f), hm1 / 2.f);
  tp.x0f = floorf(tp.ix);
  tp.y0f = floorf(tp.iy);
  Bilin bl = bilinear_setup(tp, W, H);
  bool v00 = bl.vx0 && bl.vy0, v01 = bl.vx1 && bl.vy0, v10 = bl.vx0 && bl.vy1, v11 = bl.vx1 && bl.vy1;
  for (int ch = 0; ch < C; ++ch) {
    const float* pl = img + ((int64_t)b * C + ch) * H * W + (int64_t)bl.y0 * W + bl.x0;
    float a = 0.f;
    a += v00 ? __ldg(pl) * bl.nw : 0.f;
    a += v01 ? __ldg(pl + 1) * bl.ne : 0.f;
    a += v10 ? __ldg(pl + W) * bl.sw : 0.f;
    a += v11 ? __ldg(pl + W + 1) * bl.se : 0.f;
    out[((int64_t)b * C + ch) * N + n] = a;
  }
  if (x0y0) {
    x0y0[((int64_t)b * N + n) * 2] = clamp_index(tp.x0f);
    x0y0[((int64_t)b * N + n) * 2 + 1] = clamp_index(tp.y0f);
  }
}

// refine2.py:616-626: per ray gather its NN source views out of the k_ref warped ones, replace warps that fell outside their
// source image (channel sum <= 0) by the mean over the ray's valid views, write epi_features [N, 3*S*NN].
// One thread per (ray, sample).  warps [k_ref*S][3][N]; ref_nos [N][NN] int32.
__global__ void epi_features_train_kernel(const float* __restrict__ warps, const int32_t* __restrict__ ref_nos, int k_ref, int NN,
                                          int S, int64_t N, int sample_major, float* __restrict__ epi) {
  int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= N * S) return;
  int64_t n = t / S;
  int s = (int)(t - n * S);
  float sum[3] = {0.f, 0.f, 0.f}, cnt = 0.f;
  float v[8][3];
  bool ok[8];
  for (int k = 0; k < NN; ++k) {
    int id = ref_nos[n * NN + k];
    id = id < 0 ? 0 : (id >= k_ref ? k_ref - 1 : id);
    const float* src = warps + ((int64_t)(id * S + s) * 3) * N + n;
    v[k][0] = src[0]; v[k][1] = src[N]; v[k][2] = src[2 * N];
    ok[k] = (v[k][0] + v[k][1]) + v[k][2] > 0.f;                     // torch.sum over the channel dim
    if (ok[k]) { sum[0] += v[k][0]; sum[1] += v[k][1]; sum[2] += v[k][2]; cnt += 1.f; }
  }
  const float den = cnt + 1e-6f;
  float* o = epi + n * (3 * S * NN);
  for (int k = 0; k < NN; ++k)
    for (int c = 0; c < 3; ++c)      // stage 2 / infer: (k*S + s)*3 + c (refine2.py:626);  stage 1: s*(NN*3) + k*3 + c (base.py:664-665)
      o[sample_major ? (s * NN + k) * 3 + c : (k * S + s) * 3 + c] = ok[k] ? v[k][c] : sum[c] / den;
}

// tex_index_host: [n_views][NN] ints (NULL = identity for every view); project_mat: device [n_views][NN][12]
int launch_refine_input_f16(const float* heads, int head_stride, const float* rays, const float* or_rays, int ray_stride,
                            const float* texels, const int* tex_index_host, int n_views, int64_t rays_per_view, int NN, int H,
                            int W, const float* project_mat, int64_t N, int S, float* depth, float* add, float* mul,
                            void* refine_in_f16, int32_t* x0y0, cudaStream_t st, int64_t ray_base) {
  if (N == 0) return PN_OK;            // empty batch
  PN_REQUIRE(heads && rays && or_rays && texels && project_mat && depth && add && mul && refine_in_f16,
             "pn_refine_input_f16: null pointer");
  PN_REQUIRE(NN >= 1 && NN <= 8 && H >= 2 && W >= 2 && N >= 0 && ray_stride >= 8 && head_stride >= 3 * S,
             "pn_refine_input_f16: bad shape (NN=%d H=%d W=%d S=%d)", NN, H, W, S);
  PN_REQUIRE(n_views >= 1 && n_views <= kMaxViews && (n_views == 1 || rays_per_view >= 1),
             "pn_refine_input_f16: n_views=%d unsupported (1..%d)", n_views, kMaxViews);
  PN_REQUIRE(S == 4 || S == 8 || S == 16, "pn_refine_input_f16: S=%d unsupported (4, 8, 16); use the per-stage entry points", S);
  PN_REQUIRE((reinterpret_cast<uintptr_t>(refine_in_f16) & 15) == 0, "pn_refine_input_f16: refine_in_f16 must be 16-byte aligned");
  TexIndexViews ti;
  for (int v = 0; v < kMaxViews; ++v)
    for (int k = 0; k < 8; ++k) ti.v[v][k] = (tex_index_host && v < n_views && k < NN) ? tex_index_host[v * NN + k] : k;
  for (int v = 0; v < n_views; ++v)
    for (int k = 0; k < NN; ++k) PN_REQUIRE(ti.v[v][k] >= 0, "pn_refine_input_f16: negative texel image index");
  const float4* tx = reinterpret_cast<const float4*>(texels);
  __half* rin = reinterpret_cast<__half*>(refine_in_f16);
#define PN_RI(SS, NT)                                                                                                    \
  PN_CUDA_OK(launch_chain(refine_input_kernel<SS, NT>, dim3((unsigned)((N + (256 / SS) - 1) / (256 / SS))), dim3(256), 0, st, heads, \
                          head_stride, rays, or_rays, ray_stride, tx, ti, NN, H, W, project_mat, n_views, rays_per_view, N, depth, add, \
                          mul, rin, x0y0, ray_base))
  if (NN == 4) {
    if (S == 4) PN_RI(4, 4); else if (S == 8) PN_RI(8, 4); else PN_RI(16, 4);
  } else {
    if (S == 4) PN_RI(4, 0); else if (S == 8) PN_RI(8, 0); else PN_RI(16, 0);
  }
#undef PN_RI
  PN_LAUNCH_OK("pn_refine_input_f16");
  return PN_OK;
}

}  // namespace pn

using namespace pn;

extern "C" {

int pn_project_gather(const float* texels, const int* tex_index_host, int NN, int H, int W, const float* project_mat,
                      const float* ro_w, const float* rd_w, int ray_stride, const float* depth3d, int64_t N, int S,
                      float* epi, int epi_stride, int epi_col0, int32_t* x0y0, pn_stream_t stream) {
  if (N == 0) return PN_OK;            // empty batch: nothing to validate or launch
  PN_REQUIRE(texels && project_mat && ro_w && rd_w && depth3d && epi, "pn_project_gather: null pointer");
  PN_REQUIRE(NN >= 1 && NN <= 8 && H >= 2 && W >= 2 && S >= 1 && N >= 0 && epi_col0 >= 0 && ray_stride >= 3 &&
                 epi_stride >= epi_col0 + 3 * NN * S,
             "pn_project_gather: bad shape (NN=%d H=%d W=%d S=%d epi_stride=%d epi_col0=%d)", NN, H, W, S, epi_stride,
             epi_col0);
  int64_t total = N * S;
  unsigned nb = (unsigned)((total + 255) / 256);
  const float4* tx = reinterpret_cast<const float4*>(texels);
  TexIndex ti;
  for (int k = 0; k < 8; ++k) ti.v[k] = (tex_index_host && k < NN) ? tex_index_host[k] : k;
  for (int k = 0; k < NN; ++k) PN_REQUIRE(ti.v[k] >= 0, "pn_project_gather: negative texel image index");
  if (NN == 4)
    project_gather_kernel<4><<<nb, 256, 0, as_stream(stream)>>>(tx, ti, NN, H, W, project_mat, ro_w, rd_w, ray_stride, depth3d, N, S, epi,
                                                               epi_stride, epi_col0, x0y0);
  else
    project_gather_kernel<0><<<nb, 256, 0, as_stream(stream)>>>(tx, ti, NN, H, W, project_mat, ro_w, rd_w, ray_stride, depth3d, N, S, epi,
                                                               epi_stride, epi_col0, x0y0);
  PN_LAUNCH_OK("pn_project_gather");
  return PN_OK;
}

int pn_refine_input_f16(const float* heads, int head_stride, const float* rays, const float* or_rays, int ray_stride,
                        const float* texels, const int* tex_index_host, int NN, int H, int W, const float* project_mat,
                        int64_t N, int S, float* depth, float* add, float* mul, void* refine_in_f16, int32_t* x0y0,
                        pn_stream_t stream) {
  return launch_refine_input_f16(heads, head_stride, rays, or_rays, ray_stride, texels, tex_index_host, 1, N, NN, H, W,
                                 project_mat, N, S, depth, add, mul, refine_in_f16, x0y0, as_stream(stream));
}

int pn_warp_train(const float* img, int B, int C, int H, int W, const float* depth, const float* ro1, const float* rd1,
                  int64_t ro_bstride, const float* c2w2, const float* intrinsics, int64_t N, float* out, int32_t* x0y0,
                  pn_stream_t stream) {
  if (N == 0) return PN_OK;            // empty batch
  PN_REQUIRE(img && depth && ro1 && rd1 && c2w2 && intrinsics && out, "pn_warp_train: null pointer");
  PN_REQUIRE(B >= 1 && B <= 65535 && C >= 1 && H >= 2 && W >= 2 && N >= 0 && ro_bstride >= 0,
             "pn_warp_train: bad shape (B=%d C=%d H=%d W=%d)", B, C, H, W);
  dim3 grid((unsigned)((N + 255) / 256), (unsigned)B);
  warp_train_kernel<<<grid, 256, 0, as_stream(stream)>>>(img, B, C, H, W, depth, ro1, rd1, ro_bstride, c2w2, intrinsics, N, out, x0y0);
  PN_LAUNCH_OK("pn_warp_train");
  return PN_OK;
}

int pn_epi_features_train(const float* warps, const int32_t* ref_nos, int k_ref, int NN, int S, int64_t N, int sample_major,
                          float* epi, pn_stream_t stream) {
  if (N == 0) return PN_OK;            // empty batch
  PN_REQUIRE(warps && ref_nos && epi && k_ref >= 1 && NN >= 1 && NN <= 8 && S >= 1 && N >= 0, "pn_epi_features_train: bad arguments");
  epi_features_train_kernel<<<(unsigned)((N * S + 255) / 256), 256, 0, as_stream(stream)>>>(warps, ref_nos, k_ref, NN, S, N, sample_major, epi);
  PN_LAUNCH_OK("pn_epi_features_train");
  return PN_OK;
}

int pn_warp(const float* img, int B, int C, int H, int W, const float* depth, const float* ro1, const float* rd1,
            int64_t ro_bstride, const float* w2c, int64_t N, float* out, int32_t* x0y0, pn_stream_t stream) {
  if (N == 0) return PN_OK;            // empty batch: nothing to validate or launch
  PN_REQUIRE(img && depth && ro1 && rd1 && w2c && out, "pn_warp: null pointer");
  PN_REQUIRE(B >= 1 && B <= 65535 && C >= 1 && H >= 2 && W >= 2 && N >= 0 && ro_bstride >= 0,
             "pn_warp: bad shape (B=%d C=%d H=%d W=%d)", B, C, H, W);
  dim3 grid((unsigned)((N + 255) / 256), (unsigned)B);
  warp_kernel<<<grid, 256, 0, as_stream(stream)>>>(img, B, C, H, W, depth, ro1, rd1, ro_bstride, w2c, N, out, x0y0);
  PN_LAUNCH_OK("pn_warp");
  return PN_OK;
}

}  // extern "C"
