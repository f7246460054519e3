// Per-ray elementwise / scan stages of the ProNeRF render hot path (fp32, HBM-bound).
//
//   pn_embed            helpers.py:654-692   frequency positional encoding (stand-alone form)
//   pn_pluecker         helpers.py:629-632
//   pn_sampler_input    trt.py:274-278       48-point Pluecker ray encoding
//   pn_sort_lift        trt.py:631-637       scale, stable 8-sort, gather add/mul, depth lift
//   pn_refine_pluecker  trt.py:656-658
//   pn_interval_refine  trt.py:671-681
//   pn_composite        trt.py:564-597       alpha compositing as a warp-level transmittance scan
//   pn_raygen           trt.py:245-271; helpers.py:2705-2714, 2776-2793
//   pn_pack_images      trt.py:286, 296-298  (replaces the x8 replicated planar copy by one RGBA texel array)
//
// Arithmetic follows the op order of the reference as executed by PyTorch (see oracle/pronerf_oracle.py);
// where an integer result depends on it (sort permutation, depth lift feeding the projection) the
// rounding of every operation is pinned with __f*_rn intrinsics so the compiler cannot contract it.
#include "common.cuh"

namespace pn {

constexpr int kThreads = 256;
static inline unsigned blocks_for(int64_t n, int per_block = kThreads) {
  int64_t b = (n + per_block - 1) / per_block;
  return (unsigned)(b < 1 ? 1 : b);
}

// ------------------------------------------------------------------------------------------------ embed
// One thread per (row, component); writes x, then sin/cos for L octaves.  Output row is 3+6L floats.
__global__ void embed_kernel(const float* __restrict__ x, int64_t M, int L, float* __restrict__ out) {
  int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= M * 3) return;
  int64_t row = t / 3;
  int c = (int)(t - row * 3);
  float v = x[t];
  float* o = out + row * (3 + 6 * L);
  o[c] = v;
  float f = 1.f;
  for (int l = 0; l < L; ++l) {
    float s, co;
    sincosf(__fmul_rn(v, f), &s, &co);        // argument = fp32 product x * 2^l (exact), accurate range reduction
    o[3 + 6 * l + c] = s;
    o[3 + 6 * l + 3 + c] = co;
    f *= 2.f;
  }
}

// ------------------------------------------------------------------------------------------------ pluecker
__global__ void pluecker_kernel(const float* __restrict__ o, const float* __restrict__ d, int64_t M,
                                float* __restrict__ out) {
  int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= M) return;
  float f[6];
  pluecker6(o[3 * r], o[3 * r + 1], o[3 * r + 2], d[3 * r], d[3 * r + 1], d[3 * r + 2], f);
#pragma unroll
  for (int i = 0; i < 6; ++i) out[6 * r + i] = f[i];
}

// torch.linspace(0, 1, P)[i] as the CPU kernel computes it: step = 1/(P-1);
// first half  fl(step * i), second half fma(-step, P-1-i, 1).
__device__ __forceinline__ float linspace01(int i, int P) {
  float step = __fdiv_rn(1.f, (float)(P - 1));
  return (i < P / 2) ? __fmul_rn(step, (float)i) : __fmaf_rn(-step, (float)(P - 1 - i), 1.f);
}

// torch.linspace(0, end, n)[i] as the CPU kernel computes it: step = end/(n-1); first half fl(step * i), second half end - step*(n-1-i)
__device__ __forceinline__ float linspace_step(int i, int n, float end) {
  if (n <= 1) return 0.f;
  float step = __fdiv_rn(end, (float)(n - 1));
  return (i < n / 2) ? __fmul_rn(step, (float)i) : __fsub_rn(end, __fmul_rn(step, (float)(n - 1 - i)));
}

// One thread per (ray, point): 6 outputs.  (All P blocks are equal up to rounding, but the reference feeds
// the rounded ones to the sampler, so they are reproduced op for op.)
__global__ void sampler_input_kernel(const float* __restrict__ rays, int ray_stride, int64_t N, int P,
                                     float* __restrict__ mm) {
  int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= N * P) return;
  int64_t r = t / P;
  int p = (int)(t - r * P);
  const float* ray = rays + r * ray_stride;
  float tt = (P > 1) ? linspace01(p, P) : 0.f;
  float ox = __fadd_rn(ray[0], __fmul_rn(ray[3], tt));
  float oy = __fadd_rn(ray[1], __fmul_rn(ray[4], tt));
  float oz = __fadd_rn(ray[2], __fmul_rn(ray[5], tt));
  float f[6];
  pluecker6(ox, oy, oz, ray[3], ray[4], ray[5], f);
  float* o = mm + t * 6;
#pragma unroll
  for (int i = 0; i < 6; ++i) o[i] = f[i];
}

// ------------------------------------------------------------------------------------------------ sort + lift
// torch.sort order: ascending, stable, NaN last.
__device__ __forceinline__ bool sort_gt(float a, float b) { return (a > b) || (isnan(a) && !isnan(b)); }

template <int S>
__global__ void sort_lift_kernel(const float* __restrict__ heads, int head_stride, const float* __restrict__ rays,
                                 int ray_stride, int64_t N, float* __restrict__ depth, float* __restrict__ add,
                                 float* __restrict__ mul, int32_t* __restrict__ perm, float* __restrict__ depth3d) {
  int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= N) return;
  const float* h = heads + r * head_stride;
  float near_ = rays[r * ray_stride + 6], far_ = rays[r * ray_stride + 7];
  float span = __fsub_rn(far_, near_);
  float v[S];
  int idx[S];
#pragma unroll
  for (int i = 0; i < S; ++i) {
    v[i] = __fadd_rn(__fmul_rn(h[i], span), near_);     // depth * (far - near) + near   trt.py:631
    idx[i] = i;
  }
  // stable insertion sort, fully unrolled (S is 4/8/16): registers only
#pragma unroll
  for (int i = 1; i < S; ++i) {
#pragma unroll
    for (int j = i; j > 0; --j) {
      bool sw = sort_gt(v[j - 1], v[j]);
      float a = v[j - 1], b = v[j];
      int ia = idx[j - 1], ib = idx[j];
      v[j - 1] = sw ? b : a;  v[j] = sw ? a : b;
      idx[j - 1] = sw ? ib : ia;  idx[j] = sw ? ia : ib;
    }
  }
#pragma unroll
  for (int i = 0; i < S; ++i) {
    if (depth) depth[r * S + i] = v[i];
    if (perm) perm[r * S + i] = idx[i];
    if (add) add[r * S + i] = h[S + idx[i]];
    if (mul) mul[r * S + i] = h[2 * S + idx[i]];
    // 1/(1 - depth - 1e-5): two subtractions, then a correctly rounded reciprocal   trt.py:637
    if (depth3d) depth3d[r * S + i] = __fdiv_rn(1.f, __fsub_rn(__fsub_rn(1.f, v[i]), 1e-5f));
  }
}

// generic S (<= 64): same algorithm, arrays in local memory
__global__ void sort_lift_generic_kernel(const float* __restrict__ heads, int head_stride,
                                         const float* __restrict__ rays, int ray_stride, int64_t N, int S,
                                         float* __restrict__ depth, float* __restrict__ add, float* __restrict__ mul,
                                         int32_t* __restrict__ perm, float* __restrict__ depth3d) {
  int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= N) return;
  const float* h = heads + r * head_stride;
  float near_ = rays[r * ray_stride + 6], far_ = rays[r * ray_stride + 7];
  float span = __fsub_rn(far_, near_);
  float v[64];
  int idx[64];
  for (int i = 0; i < S; ++i) {
    float x = __fadd_rn(__fmul_rn(h[i], span), near_);
    int j = i;
    while (j > 0 && sort_gt(v[j - 1], x)) { v[j] = v[j - 1]; idx[j] = idx[j - 1]; --j; }
    v[j] = x; idx[j] = i;
  }
  for (int i = 0; i < S; ++i) {
    if (depth) depth[r * S + i] = v[i];
    if (perm) perm[r * S + i] = idx[i];
    if (add) add[r * S + i] = h[S + idx[i]];
    if (mul) mul[r * S + i] = h[2 * S + idx[i]];
    if (depth3d) depth3d[r * S + i] = __fdiv_rn(1.f, __fsub_rn(__fsub_rn(1.f, v[i]), 1e-5f));
  }
}

// ------------------------------------------------------------------------------------------------ refine input
__global__ void refine_pluecker_kernel(const float* __restrict__ rays, int ray_stride, const float* __restrict__ depth,
                                       int64_t N, int S, float* __restrict__ out, int out_stride) {
  int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= N * S) return;
  int64_t r = t / S;
  int s = (int)(t - r * S);
  const float* ray = rays + r * ray_stride;
  float z = depth[t];
  float ox = __fadd_rn(ray[0], __fmul_rn(ray[3], z));
  float oy = __fadd_rn(ray[1], __fmul_rn(ray[4], z));
  float oz = __fadd_rn(ray[2], __fmul_rn(ray[5], z));
  float f[6];
  pluecker6(ox, oy, oz, ray[3], ray[4], ray[5], f);
  float* o = out + r * out_stride + 6 * s;
#pragma unroll
  for (int i = 0; i < 6; ++i) o[i] = f[i];
}

// ------------------------------------------------------------------------------------------------ interval refine
// ||d|| of the NDC direction, the factor of raw2outputs' distances (trt.py:578): ONE definition for the kernel that can precompute
// it (interval refinement reads the ray row anyway) and the compositing kernel, so that both forms give the same bits
__device__ __forceinline__ float ray_dir_norm(const float* __restrict__ rd) {
  return sqrtf(rd[0] * rd[0] + rd[1] * rd[1] + rd[2] * rd[2]);
}

__global__ void interval_refine_kernel(const float* __restrict__ rays, int ray_stride, const float* __restrict__ depth,
                                       const float* __restrict__ ro, int ro_stride, int64_t N, int S,
                                       float* __restrict__ z, float* __restrict__ q, float* __restrict__ dnorm) {
  pdl_wait();
  pdl_launch();
  int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= N * S) return;
  int64_t r = t / S;
  int s = (int)(t - r * S);
  const float* ray = rays + r * ray_stride;
  const float* d = depth + r * S;
  float near_ = ray[6], far_ = ray[7];
  float dc = d[s];
  float upper = (s + 1 < S) ? __fmul_rn(0.5f, __fadd_rn(d[s + 1], dc)) : __fmul_rn(0.5f, __fadd_rn(far_, dc));
  float lower = (s > 0) ? __fmul_rn(0.5f, __fadd_rn(dc, d[s - 1])) : __fmul_rn(0.5f, __fadd_rn(near_, dc));
  float frac = ro[r * ro_stride + s];
  float zz = __fadd_rn(lower, __fmul_rn(__fsub_rn(upper, lower), frac));
  if (z) z[t] = zz;
  if (dnorm && s == 0) dnorm[r] = ray_dir_norm(ray + 3);     // for the compositing kernel: 4 B/ray instead of a 44-byte ray row
  const float* off = ro + r * ro_stride + S + 3 * s;
#pragma unroll
  for (int c = 0; c < 3; ++c)
    q[t * 3 + c] = __fadd_rn(__fadd_rn(ray[c], __fmul_rn(ray[3 + c], zz)), __fmul_rn(1e-2f, off[c]));
}

// ------------------------------------------------------------------------------------------------ stage-1 exploration sampling
// base.py:689-707, 730 (the NeRF-only training step of stage 1), deterministic "forward" variant: every predicted sample s
// spawns n_mult samples z = d_s + (m / n_mult) * |d_s - d_{s+1}| (d_S := far), m = 0..n_mult-1 -- already ascending, so the
// reference's torch.sort (":not really needed?") is the identity -- and the query points o + dir * z.  The reference draws
// n_mult, the direction and an extra |N(0, 0.2)| jitter at random; a benchmark fixes n_mult and drops the jitter.
__global__ void explore_samples_kernel(const float* __restrict__ rays, int ray_stride, const float* __restrict__ depth, int64_t N,
                                       int S, int n_mult, float mult_end, float* __restrict__ z, float* __restrict__ q) {
  const int So = S * n_mult;
  int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= N * So) return;
  int64_t r = t / So;
  int j = (int)(t - r * So);
  int s = j / n_mult, m = j - s * n_mult;
  const float* ray = rays + r * ray_stride;
  const float* d = depth + r * S;
  float dc = d[s];
  float dn = (s + 1 < S) ? d[s + 1] : ray[7];                                   // far * ones   (base.py:698)
  float mult = (n_mult > 1) ? linspace_step(m, n_mult, mult_end) : 0.f;         // torch.linspace(0, 1 - 1/n_mult, n_mult)[m]
  float zz = __fadd_rn(dc, __fmul_rn(mult, fabsf(__fsub_rn(dc, dn))));
  z[t] = zz;
#pragma unroll
  for (int c = 0; c < 3; ++c) q[t * 3 + c] = __fadd_rn(ray[c], __fmul_rn(ray[3 + c], zz));
}

// The TRAINING-time form of the same step (base.py:689-729, randomize=True, train_sampler=False) with the reference's random draws
// as INPUTS: n_mult = random.randint(1, 64/S); dir1 = random.random() > 0.5 (spread towards the next sample / far, else back
// towards the previous sample / near); noise [N, S*n_mult] = the |N(0,1)|/5 draw clamped at 0.99 (base.py:716-718; NULL = none);
// dir2 = the second random.random() > 0.5.  One warp per ray:
//   z1[s*n+m] = d_s +- mult_m * |d_s - d_(s+-1)|   (base.py:694-707),  sorted ascending (torch.sort, base.py:709),
//   z [j]     = z1[j] +- noise[j] * |z1[j] - z1[j+-1]|   with far / near beyond the ends (base.py:719-728; NOT re-sorted),
//   q [j]     = o + dir * z[j]                           (base.py:730).
// The sort is a rank sort through shared memory (<= 64 values): in the forward case the values are ascending by construction
// except where a rounded gap lifts the last replica of a sample one ulp above the next sample; torch.sort orders those too.
__global__ void __launch_bounds__(256)
explore_samples_rand_kernel(const float* __restrict__ rays, int ray_stride, const float* __restrict__ depth, int64_t N, int S,
                            int n_mult, float mult_end, int dir1_fwd, const float* __restrict__ noise, int dir2_fwd,
                            float* __restrict__ z, float* __restrict__ q) {
  __shared__ float s_a[8][64], s_b[8][64];
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t r = (int64_t)blockIdx.x * 8 + w;
  if (r >= N) return;
  const int So = S * n_mult;
  const float* ray = rays + r * ray_stride;
  const float* d = depth + r * S;
  const float near_ = ray[6], far_ = ray[7];
  for (int j = lane; j < So; j += 32) {
    const int s = j / n_mult, m = j - s * n_mult;
    const float dc = d[s];
    float v = dc;
    if (n_mult > 1) {
      const float mult = linspace_step(m, n_mult, mult_end);
      if (dir1_fwd) {
        const float dn = (s + 1 < S) ? d[s + 1] : far_;
        v = __fadd_rn(dc, __fmul_rn(mult, fabsf(__fsub_rn(dc, dn))));
      } else {
        const float dp = (s > 0) ? d[s - 1] : near_;
        v = __fadd_rn(dc, __fmul_rn(-mult, fabsf(__fsub_rn(dc, dp))));
      }
    }
    s_a[w][j] = v;
  }
  __syncwarp();
  for (int j = lane; j < So; j += 32) {                      // stable ascending rank (NaN last, like torch.sort)
    const float v = s_a[w][j];
    int rank = 0;
    for (int i = 0; i < So; ++i) {
      const float u = s_a[w][i];
      const bool before = (u < v) || (!(u != u) && (v != v)) || ((u == v || ((u != u) && (v != v))) && i < j);
      rank += before ? 1 : 0;
    }
    s_b[w][rank] = v;
  }
  __syncwarp();
  for (int j = lane; j < So; j += 32) {
    const float zc = (n_mult > 1) ? s_b[w][j] : s_a[w][j];
    float zz = zc;
    if (noise) {
      const float* zs = (n_mult > 1) ? s_b[w] : s_a[w];
      const float nz = noise[r * So + j];
      if (dir2_fwd) {
        const float zn = (j + 1 < So) ? zs[j + 1] : far_;
        zz = __fadd_rn(zc, __fmul_rn(nz, fabsf(__fsub_rn(zc, zn))));
      } else {
        const float zp = (j > 0) ? zs[j - 1] : near_;
        zz = __fadd_rn(zc, __fmul_rn(-nz, fabsf(__fsub_rn(zc, zp))));
      }
    }
    z[r * So + j] = zz;
#pragma unroll
    for (int c = 0; c < 3; ++c) q[(r * So + j) * 3 + c] = __fadd_rn(ray[c], __fmul_rn(ray[3 + c], zz));
  }
}

// ------------------------------------------------------------------------------------------------ compositing
// S lanes per ray (S a power of two <= 32): lane = (ray-in-warp, sample).  Every global access is
// coalesced (consecutive lanes read consecutive float4 / float); the transmittance T_s = prod_{j<s}(1-a_j+1e-10)
// is an exclusive product scan over the S-lane segment with __shfl_up_sync, the weighted sums are
// segment reductions with __shfl_xor_sync.
// 1 / max(1e-10, depth / acc); torch.max propagates the NaN of 0/0, fmaxf would not.
__device__ __forceinline__ float disp_of(float wz, float ws) {
  float q = wz / ws;
  return 1.f / ((q != q) ? q : fmaxf(1e-10f, q));
}

// stage 1 clamps the network output before compositing (torch.clamp(raw, -10, 10), base.py:523); the infer path does not (Q4)
__device__ __forceinline__ float4 clamp_raw(float4 r, float c) {
  if (c > 0.f) { r.x = fminf(fmaxf(r.x, -c), c); r.y = fminf(fmaxf(r.y, -c), c); r.z = fminf(fmaxf(r.z, -c), c); r.w = fminf(fmaxf(r.w, -c), c); }
  return r;
}

// Where ray r of the batch lands in the output frame.  Dense (view_stride = 0): row r.  Banded multi-view batches (a rank's
// band of every view written straight into a frame set [n_views][view_stride rays], pn_frame_t.out_view_stride): ray
// g = ray_base + r belongs to view v = g / rays_per_view and goes to row v * view_stride + (g - v * rays_per_view).
struct OutMap {
  int64_t rays_per_view, view_stride, ray_base;
  __device__ __forceinline__ int64_t row(int64_t r) const {
    if (view_stride == 0) return r;
    const int64_t g = r + ray_base, v = g / rays_per_view;
    return v * view_stride + (g - v * rays_per_view);
  }
};

template <int S>
__global__ void composite_scan_kernel(const float* __restrict__ raw, const float* __restrict__ z,
                                      const float* __restrict__ rays, int ray_stride, int ray_d_col,
                                      const float* __restrict__ add, const float* __restrict__ mul, int64_t N,
                                      float* __restrict__ rgb, float* __restrict__ depth, float* __restrict__ disp,
                                      float* __restrict__ acc, float* __restrict__ weights, float raw_clamp, OutMap om,
                                      const float* __restrict__ dnorm) {
  pdl_wait();
  pdl_launch();
  int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;      // one (ray, sample) per thread
  int64_t r = t / S;
  int s = (int)(t % S);
  bool live = r < N;
  int64_t rr = live ? r : (N - 1);
  int64_t tt = rr * S + s;
  float4 rw = clamp_raw(reinterpret_cast<const float4*>(raw)[tt], raw_clamp);
  float zc = z[tt];
  float zn = __shfl_down_sync(0xffffffffu, zc, 1);                  // z_{s+1} (same segment when s < S-1)
  float dn = dnorm ? dnorm[rr] : ray_dir_norm(rays + rr * ray_stride + ray_d_col);
  float dist = (s == S - 1) ? 1e10f : (zn - zc);
  dist *= dn;
  float sig = fmaxf(rw.w + (add ? add[tt] : 0.f), 0.f);
  float alpha = (1.f - expf(-sig * dist)) * (mul ? fmaxf(mul[tt], 0.f) : 1.f);
  // exclusive product scan of (1 - alpha + 1e-10) over the segment
  float f = 1.f - alpha + 1e-10f;
  float incl = f;
#pragma unroll
  for (int o = 1; o < S; o <<= 1) {
    float up = __shfl_up_sync(0xffffffffu, incl, o);
    if (s >= o) incl *= up;
  }
  float T = __shfl_up_sync(0xffffffffu, incl, 1);
  if (s == 0) T = 1.f;
  float w = alpha * T;
  float cr = w * sigmoidf_(rw.x), cg = w * sigmoidf_(rw.y), cb = w * sigmoidf_(rw.z);
  float wz = w * zc, ws = w;
#pragma unroll
  for (int o = S / 2; o > 0; o >>= 1) {
    cr += __shfl_xor_sync(0xffffffffu, cr, o);
    cg += __shfl_xor_sync(0xffffffffu, cg, o);
    cb += __shfl_xor_sync(0xffffffffu, cb, o);
    wz += __shfl_xor_sync(0xffffffffu, wz, o);
    ws += __shfl_xor_sync(0xffffffffu, ws, o);
  }
  if (!live) return;
  if (weights) weights[tt] = w;
  if (s == 0) {
    const int64_t ro = om.row(r);
    rgb[3 * ro] = cr; rgb[3 * ro + 1] = cg; rgb[3 * ro + 2] = cb;
    depth[ro] = wz;
    if (acc) acc[r] = ws;
    if (disp) disp[r] = disp_of(wz, ws);
  }
}

// generic S: one thread per ray, sequential (exactly the reference's cumprod order)
__global__ void composite_seq_kernel(const float* __restrict__ raw, const float* __restrict__ z,
                                     const float* __restrict__ rays, int ray_stride, int ray_d_col,
                                     const float* __restrict__ add, const float* __restrict__ mul, int64_t N, int S,
                                     float* __restrict__ rgb, float* __restrict__ depth, float* __restrict__ disp,
                                     float* __restrict__ acc, float* __restrict__ weights, float raw_clamp, OutMap om,
                                     const float* __restrict__ dnorm) {
  pdl_wait();
  pdl_launch();
  int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= N) return;
  float dn = dnorm ? dnorm[r] : ray_dir_norm(rays + r * ray_stride + ray_d_col);
  float T = 1.f, cr = 0.f, cg = 0.f, cb = 0.f, wz = 0.f, ws = 0.f;
  for (int s = 0; s < S; ++s) {
    int64_t tt = r * S + s;
    float4 rw = clamp_raw(reinterpret_cast<const float4*>(raw)[tt], raw_clamp);
    float zc = z[tt];
    float dist = (s == S - 1) ? 1e10f : (z[tt + 1] - zc);
    dist *= dn;
    float alpha = (1.f - expf(-fmaxf(rw.w + (add ? add[tt] : 0.f), 0.f) * dist)) * (mul ? fmaxf(mul[tt], 0.f) : 1.f);
    float w = alpha * T;
    T *= (1.f - alpha + 1e-10f);
    cr += w * sigmoidf_(rw.x); cg += w * sigmoidf_(rw.y); cb += w * sigmoidf_(rw.z);
    wz += w * zc; ws += w;
    if (weights) weights[tt] = w;
  }
  const int64_t ro = om.row(r);
  rgb[3 * ro] = cr; rgb[3 * ro + 1] = cg; rgb[3 * ro + 2] = cb;
  depth[ro] = wz;
  if (acc) acc[r] = ws;
  if (disp) disp[r] = disp_of(wz, ws);
}

// ------------------------------------------------------------------------------------------------ ray generation
struct RaygenParams {
  int H, W, row0, nrows;
  float fx, fy, cx, cy;          // fp32 images of the float64 intrinsics (torch folds python scalars to fp32)
  float a, b;                    // -1/(W/(2f)), -1/(H/(2f)) evaluated in double, rounded once
  float c2w[12];
  float near_, far_, or_near, or_far;
};

// One ray per thread; a block's rows ([kThreads][11] floats, contiguous in global memory) are staged in shared memory and
// leave as 16-byte coalesced stores -- 22 four-byte stores at a 44-byte stride per thread reached 0.73 TB/s (ncu launch list).
__global__ void __launch_bounds__(kThreads) raygen_kernel(RaygenParams p, float* __restrict__ rays, float* __restrict__ or_rays) {
  __shared__ __align__(16) float s_ndc[kThreads * 11];
  __shared__ __align__(16) float s_or[kThreads * 11];
  const int64_t t0 = (int64_t)blockIdx.x * blockDim.x;
  const int64_t t = t0 + threadIdx.x;
  const int64_t n = (int64_t)p.nrows * p.W;
  if (t < n) {
    int j = p.row0 + (int)(t / p.W);
    int i = (int)(t % p.W);
    // get_rays  helpers.py:2705-2714
    float dx = __fdiv_rn(__fsub_rn((float)i, p.cx), p.fx);
    float dy = __fdiv_rn(-__fsub_rn((float)j, p.cy), p.fy);
    float dz = -1.f;
    float d[3], o[3];
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      d[r] = __fadd_rn(__fadd_rn(__fmul_rn(dx, p.c2w[4 * r]), __fmul_rn(dy, p.c2w[4 * r + 1])), __fmul_rn(dz, p.c2w[4 * r + 2]));
      o[r] = p.c2w[4 * r + 3];
    }
    float nrm = sqrtf(__fmaf_rn(d[2], d[2], __fmaf_rn(d[1], d[1], __fmul_rn(d[0], d[0]))));
    float v[3] = {__fdiv_rn(d[0], nrm), __fdiv_rn(d[1], nrm), __fdiv_rn(d[2], nrm)};
    {
      float* q = s_or + threadIdx.x * 11;
      q[0] = o[0]; q[1] = o[1]; q[2] = o[2]; q[3] = d[0]; q[4] = d[1]; q[5] = d[2];
      q[6] = p.or_near; q[7] = p.or_far; q[8] = v[0]; q[9] = v[1]; q[10] = v[2];
    }
    // ndc_rays with near = 1   helpers.py:2776-2793
    float tn = __fdiv_rn(-__fadd_rn(1.f, o[2]), d[2]);
    float ox = __fadd_rn(o[0], __fmul_rn(tn, d[0]));
    float oy = __fadd_rn(o[1], __fmul_rn(tn, d[1]));
    float oz = __fadd_rn(o[2], __fmul_rn(tn, d[2]));
    float rz = __fdiv_rn(1.f, oz);
    float o0 = __fdiv_rn(__fmul_rn(p.a, ox), oz);
    float o1 = __fdiv_rn(__fmul_rn(p.b, oy), oz);
    float o2 = __fadd_rn(1.f, __fmul_rn(rz, 2.f));
    float d0 = __fmul_rn(p.a, __fsub_rn(__fdiv_rn(d[0], d[2]), __fdiv_rn(ox, oz)));
    float d1 = __fmul_rn(p.b, __fsub_rn(__fdiv_rn(d[1], d[2]), __fdiv_rn(oy, oz)));
    float d2 = __fmul_rn(rz, -2.f);
    float* q = s_ndc + threadIdx.x * 11;
    q[0] = o0; q[1] = o1; q[2] = o2; q[3] = d0; q[4] = d1; q[5] = d2;
    q[6] = p.near_; q[7] = p.far_; q[8] = v[0]; q[9] = v[1]; q[10] = v[2];
  }
  __syncthreads();
  const int64_t left = n - t0;
  const int nfl = (int)(left < (int64_t)blockDim.x ? left : (int64_t)blockDim.x) * 11;
  auto flush = [&](const float* src, float* dst) {
    if ((reinterpret_cast<uintptr_t>(dst) & 15) == 0) {
      for (int i = threadIdx.x * 4; i < nfl; i += blockDim.x * 4) {
        if (i + 4 <= nfl) *reinterpret_cast<float4*>(dst + i) = *reinterpret_cast<const float4*>(src + i);
        else for (int k = i; k < nfl; ++k) dst[k] = src[k];
      }
    } else {
      for (int i = threadIdx.x; i < nfl; i += blockDim.x) dst[i] = src[i];
    }
  };
  flush(s_ndc, rays + t0 * 11);
  if (or_rays) flush(s_or, or_rays + t0 * 11);
}

// ------------------------------------------------------------------------------------------------ image packing
__global__ void pack_images_kernel(const float* __restrict__ hwc, int64_t npix, float4* __restrict__ texels) {
  int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= npix) return;
  texels[t] = make_float4(hwc[3 * t], hwc[3 * t + 1], hwc[3 * t + 2], 0.f);
}

}  // namespace pn

using namespace pn;

extern "C" {

int pn_embed(const float* x, int64_t M, int L, float* out, pn_stream_t stream) {
  if (M == 0) return PN_OK;            // empty batch: nothing to validate or launch
  PN_REQUIRE(x && out && M >= 0 && L >= 0 && L <= 16, "pn_embed: bad arguments (M=%lld L=%d)", (long long)M, L);
  embed_kernel<<<blocks_for(M * 3), kThreads, 0, as_stream(stream)>>>(x, M, L, out);
  PN_LAUNCH_OK("pn_embed");
  return PN_OK;
}

int pn_pluecker(const float* o, const float* d, int64_t M, float* out, pn_stream_t stream) {
  if (M == 0) return PN_OK;            // empty batch: nothing to validate or launch
  PN_REQUIRE(o && d && out && M >= 0, "pn_pluecker: bad arguments");
  pluecker_kernel<<<blocks_for(M), kThreads, 0, as_stream(stream)>>>(o, d, M, out);
  PN_LAUNCH_OK("pn_pluecker");
  return PN_OK;
}

int pn_sampler_input(const float* rays, int ray_stride, int64_t N, int P, float* mm_input, pn_stream_t stream) {
  if (N == 0) return PN_OK;            // empty batch: nothing to validate or launch
  PN_REQUIRE(rays && mm_input && N >= 0 && P >= 1 && ray_stride >= 6, "pn_sampler_input: bad arguments");
  sampler_input_kernel<<<blocks_for(N * P), kThreads, 0, as_stream(stream)>>>(rays, ray_stride, N, P, mm_input);
  PN_LAUNCH_OK("pn_sampler_input");
  return PN_OK;
}

int pn_sort_lift(const float* heads, int head_stride, const float* rays, int ray_stride, int64_t N, int S,
                 float* depth, float* add, float* mul, int32_t* perm, float* depth3d, pn_stream_t stream) {
  if (N == 0) return PN_OK;            // empty batch: nothing to validate or launch
  PN_REQUIRE(heads && rays && N >= 0 && S >= 1 && S <= 64 && head_stride >= 3 * S && ray_stride >= 8,
             "pn_sort_lift: bad arguments (S=%d head_stride=%d ray_stride=%d)", S, head_stride, ray_stride);
  cudaStream_t st = as_stream(stream);
  unsigned nb = blocks_for(N, 128);
  switch (S) {
    case 4:  sort_lift_kernel<4><<<nb, 128, 0, st>>>(heads, head_stride, rays, ray_stride, N, depth, add, mul, perm, depth3d); break;
    case 8:  sort_lift_kernel<8><<<nb, 128, 0, st>>>(heads, head_stride, rays, ray_stride, N, depth, add, mul, perm, depth3d); break;
    case 16: sort_lift_kernel<16><<<nb, 128, 0, st>>>(heads, head_stride, rays, ray_stride, N, depth, add, mul, perm, depth3d); break;
    default: sort_lift_generic_kernel<<<nb, 128, 0, st>>>(heads, head_stride, rays, ray_stride, N, S, depth, add, mul, perm, depth3d);
  }
  PN_LAUNCH_OK("pn_sort_lift");
  return PN_OK;
}

int pn_refine_pluecker(const float* rays, int ray_stride, const float* depth, int64_t N, int S, float* out,
                       int out_stride, pn_stream_t stream) {
  if (N == 0) return PN_OK;            // empty batch: nothing to validate or launch
  PN_REQUIRE(rays && depth && out && N >= 0 && S >= 1 && ray_stride >= 6 && out_stride >= 6 * S,
             "pn_refine_pluecker: bad arguments");
  refine_pluecker_kernel<<<blocks_for(N * S), kThreads, 0, as_stream(stream)>>>(rays, ray_stride, depth, N, S, out, out_stride);
  PN_LAUNCH_OK("pn_refine_pluecker");
  return PN_OK;
}

int pn_interval_refine(const float* rays, int ray_stride, const float* depth, const float* refine_out,
                       int refine_stride, int64_t N, int S, float* z, float* query, pn_stream_t stream) {
  if (N == 0) return PN_OK;            // empty batch: nothing to validate or launch
  PN_REQUIRE(rays && depth && refine_out && query && N >= 0 && S >= 1 && ray_stride >= 8 && refine_stride >= 4 * S,
             "pn_interval_refine: bad arguments");
  PN_CUDA_OK(launch_chain(interval_refine_kernel, dim3(blocks_for(N * S)), dim3(kThreads), 0, as_stream(stream), rays, ray_stride, depth,
                          refine_out, refine_stride, N, S, z, query, (float*)nullptr));
  PN_LAUNCH_OK("pn_interval_refine");
  return PN_OK;
}

}  // extern "C"

namespace pn {
// pn_interval_refine for the composed path: also leaves ||d_ndc|| per ray for the compositing kernel
int interval_refine_dnorm(const float* rays, int ray_stride, const float* depth, const float* refine_out, int refine_stride, int64_t N,
                          int S, float* z, float* query, float* dnorm, cudaStream_t st) {
  if (N == 0) return PN_OK;
  PN_REQUIRE(rays && depth && refine_out && query && dnorm && S >= 1 && ray_stride >= 8 && refine_stride >= 4 * S, "interval_refine: bad arguments");
  PN_CUDA_OK(launch_chain(interval_refine_kernel, dim3(blocks_for(N * S)), dim3(kThreads), 0, st, rays, ray_stride, depth, refine_out,
                          refine_stride, N, S, z, query, dnorm));
  PN_LAUNCH_OK("pn_interval_refine");
  return PN_OK;
}
}  // namespace pn

extern "C" {

int pn_explore_samples(const float* rays, int ray_stride, const float* depth, int64_t N, int S, int n_mult, float* z, float* query,
                       pn_stream_t stream) {
  if (N == 0) return PN_OK;            // empty batch
  PN_REQUIRE(rays && depth && z && query && N >= 0 && S >= 1 && n_mult >= 1 && S * n_mult <= 1024 && ray_stride >= 8,
             "pn_explore_samples: bad arguments (S=%d n_mult=%d)", S, n_mult);
  const float mult_end = (float)(1.0 - 1.0 / (double)n_mult);   // the Python double `1 - 1/n_mult`, rounded once like torch does
  explore_samples_kernel<<<blocks_for(N * S * n_mult), kThreads, 0, as_stream(stream)>>>(rays, ray_stride, depth, N, S, n_mult, mult_end,
                                                                                       z, query);
  PN_LAUNCH_OK("pn_explore_samples");
  return PN_OK;
}

int pn_explore_samples_rand(const float* rays, int ray_stride, const float* depth, int64_t N, int S, int n_mult, int dir1_forward,
                            const float* noise, int dir2_forward, float* z, float* query, pn_stream_t stream) {
  if (N == 0) return PN_OK;            // empty batch
  PN_REQUIRE(rays && depth && z && query && N >= 0 && S >= 1 && n_mult >= 1 && S * n_mult <= 64 && ray_stride >= 8,
             "pn_explore_samples_rand: bad arguments (S=%d n_mult=%d; S * n_mult <= 64 as in the reference, base.py:690)", S, n_mult);
  const float mult_end = (float)(1.0 - 1.0 / (double)n_mult);
  explore_samples_rand_kernel<<<(unsigned)((N + 7) / 8), 256, 0, as_stream(stream)>>>(rays, ray_stride, depth, N, S, n_mult, mult_end,
                                                                                    dir1_forward ? 1 : 0, noise, dir2_forward ? 1 : 0, z, query);
  PN_LAUNCH_OK("pn_explore_samples_rand");
  return PN_OK;
}

}  // extern "C"

namespace pn {
// raw2outputs with the output rows placed by (rays_per_view, out_view_stride, ray_base): see OutMap
int composite_mapped(const float* raw, const float* z, const float* rays, int ray_stride, int ray_d_col, const float* add,
                     const float* mul, float raw_clamp, int64_t N, int S, float* rgb, float* depth, float* disp, float* acc,
                     float* weights, int64_t rays_per_view, int64_t out_view_stride, int64_t ray_base, cudaStream_t st,
                     const float* dnorm) {
  if (N == 0) return PN_OK;            // empty batch: nothing to validate or launch
  PN_REQUIRE(raw && z && rays && rgb && depth && N >= 0 && S >= 1 && ray_stride >= ray_d_col + 3 && ((add == nullptr) == (mul == nullptr)),
             "pn_composite: bad arguments");
  PN_REQUIRE(out_view_stride == 0 || (rays_per_view >= 1 && out_view_stride >= rays_per_view && !disp && !acc),
             "pn_composite: out_view_stride=%lld needs rays_per_view in [1, out_view_stride]", (long long)out_view_stride);
  OutMap om{rays_per_view, out_view_stride, ray_base};
#define PN_COMP(SS)                                                                                                         \
  PN_CUDA_OK(launch_chain(composite_scan_kernel<SS>, dim3(blocks_for(N * SS)), dim3(kThreads), 0, st, raw, z, rays, ray_stride, \
                          ray_d_col, add, mul, N, rgb, depth, disp, acc, weights, raw_clamp, om, dnorm))
  switch (S) {
    case 2: PN_COMP(2); break;
    case 4: PN_COMP(4); break;
    case 8: PN_COMP(8); break;
    case 16: PN_COMP(16); break;
    case 32: PN_COMP(32); break;
    default:
      PN_CUDA_OK(launch_chain(composite_seq_kernel, dim3(blocks_for(N)), dim3(kThreads), 0, st, raw, z, rays, ray_stride, ray_d_col, add, mul,
                              N, S, rgb, depth, disp, acc, weights, raw_clamp, om, dnorm));
  }
#undef PN_COMP
  PN_LAUNCH_OK("pn_composite");
  return PN_OK;
}
}  // namespace pn

extern "C" {

int pn_composite_stage1(const float* raw, const float* z, const float* rays, int ray_stride, int ray_d_col, const float* add,
                        const float* mul, float raw_clamp, int64_t N, int S, float* rgb, float* depth, float* disp, float* acc,
                        float* weights, pn_stream_t stream) {
  return pn::composite_mapped(raw, z, rays, ray_stride, ray_d_col, add, mul, raw_clamp, N, S, rgb, depth, disp, acc, weights, 0, 0, 0,
                              as_stream(stream), nullptr);
}

int pn_composite(const float* raw, const float* z, const float* rays, int ray_stride, int ray_d_col, const float* add,
                 const float* mul, int64_t N, int S, float* rgb, float* depth, float* disp, float* acc, float* weights,
                 pn_stream_t stream) {
  PN_REQUIRE(N == 0 || (add && mul), "pn_composite: bad arguments");
  return pn_composite_stage1(raw, z, rays, ray_stride, ray_d_col, add, mul, 0.f, N, S, rgb, depth, disp, acc, weights, stream);
}

int pn_raygen(int H, int W, double fx, double fy, double cx, double cy, const float* c2w_host, float near_, float far_,
              float or_near, float or_far, int row0, int nrows, float* rays, float* or_rays, pn_stream_t stream) {
  PN_REQUIRE(H > 0 && W > 0 && c2w_host && rays && row0 >= 0 && nrows >= 0 && row0 + nrows <= H && fx != 0 && fy != 0,
             "pn_raygen: bad arguments (H=%d W=%d row0=%d nrows=%d)", H, W, row0, nrows);
  if (nrows == 0) return PN_OK;
  RaygenParams p;
  p.H = H; p.W = W; p.row0 = row0; p.nrows = nrows;
  p.fx = (float)fx; p.fy = (float)fy; p.cx = (float)cx; p.cy = (float)cy;
  p.a = (float)(-1. / (W / (2. * fx)));
  p.b = (float)(-1. / (H / (2. * fx)));        // the reference passes K[0][0] as "focal" for both axes (trt.py:265)
  for (int i = 0; i < 12; ++i) p.c2w[i] = c2w_host[i];
  p.near_ = near_; p.far_ = far_; p.or_near = or_near; p.or_far = or_far;
  raygen_kernel<<<blocks_for((int64_t)nrows * W), kThreads, 0, as_stream(stream)>>>(p, rays, or_rays);
  PN_LAUNCH_OK("pn_raygen");
  return PN_OK;
}

int pn_pack_images(const float* images_hwc, int NN, int H, int W, float* texels, pn_stream_t stream) {
  PN_REQUIRE(images_hwc && texels && NN >= 1 && H >= 1 && W >= 1, "pn_pack_images: bad arguments");
  int64_t npix = (int64_t)NN * H * W;
  pack_images_kernel<<<blocks_for(npix), kThreads, 0, as_stream(stream)>>>(images_hwc, npix, reinterpret_cast<float4*>(texels));
  PN_LAUNCH_OK("pn_pack_images");
  return PN_OK;
}

}  // extern "C"
