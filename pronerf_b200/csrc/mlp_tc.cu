// bf16 tensor-core tier of the three MLPs: a persistent, warp-specialised, fully fused kernel on
// tcgen05.mma with fp32 accumulators in TMEM and bulk-async (TMA engine) weight streaming.
//
//   sampler  MinMaxRaySamplerTRT_Net     helpers.py:1473-1507   288 -> 256 x6 (ELU) -> 27
//   refine   MinMaxRayEpiSamplerTRT_Net  helpers.py:1509-1540   144 -> 256 x6 (ELU) -> 35
//   NeRF     DoNeRFTRT                   helpers.py:1186-1343   63 -> 256 x7 (ReLU) -> [256 ++ 27] -> 4
//
// One CTA per SM (persistent over 128-row tiles).  Per tile and layer the GEMM is
//   D[128 x N] (fp32, TMEM) = A[128 x K] (bf16, smem, K-major, 128B swizzle) * W^T (bf16, smem, K-major)
// issued as M128 x N256 x K16 tcgen05.mma instructions by one thread.  Roles:
//
//   warp 0      weight producer: streams every layer's weights, in consumption order, through a 4-slot ring
//               (32 KB = one 64-wide K block of a 256-row layer) with cp.async.bulk + mbarrier complete_tx;
//               the global image is pre-swizzled into the UMMA canonical layout so a linear copy lands it.
//   warp 1      MMA issuer: waits for (activation K-block ready, weight slot full), issues 4 MMAs per K block,
//               tcgen05.commit frees the slot; after a layer's last block commits "accumulator full".
//   warps 2-5   epilogue / operand producers (thread = row): tcgen05.ld the accumulator 64 columns at a time,
//               bias + ReLU/ELU, convert to bf16 and store straight into the next layer's A operand (the
//               swizzle makes the row-per-thread 16-byte stores bank-conflict free), fence.proxy.async, and
//               signal that K block -- so layer l+1 starts while layer l's epilogue is still draining.
//               Two 256-column accumulators ping-pong in TMEM (512 columns) to make that overlap legal.
//
// The first-layer operand is generated in the kernel (frequency encoding, Pluecker features) or loaded;
// the NeRF view-direction term (27 inputs of the last layer, identical for a ray's samples) is added on the
// CUDA cores in the output epilogue, so the tensor-core part of the last layer is a clean K = 256, N = 16 GEMM.
#include <cuda_bf16.h>

#include "tc.cuh"

namespace pn {
namespace tc {

constexpr int TILE_M = 128;
constexpr int KBLK = 64;                                // bf16 elements per 128-byte swizzle row
constexpr int A_BLOCK_BYTES = TILE_M * 128;             // one K block of the activation tile: 16 KB
constexpr int MAX_KB = 5;                               // first layer up to 320 inputs
constexpr int A_BYTES = MAX_KB * A_BLOCK_BYTES;         // 80 KB
constexpr int SLOT_BYTES = kHidden * 128;               // 32 KB
constexpr int N_SLOTS = 4;
constexpr int BIAS_FLOATS = kMaxLayers * kHidden;       // 8 KB
constexpr int WDIR_FLOATS = 4 * 28;
constexpr int OFF_A = 0;
constexpr int OFF_RING = OFF_A + A_BYTES;
constexpr int OFF_BIAS = OFF_RING + N_SLOTS * SLOT_BYTES;
constexpr int OFF_WDIR = OFF_BIAS + BIAS_FLOATS * 4;
constexpr int OFF_BAR = OFF_WDIR + WDIR_FLOATS * 4;
// barriers: full[4], empty[4], a_ready[8] (32-column halves), in_ready, acc_full[2]  (8 bytes each) + tmem pointer
constexpr int N_BARS = N_SLOTS * 2 + 8 + 1 + 2;
constexpr int OFF_TMEM = OFF_BAR + N_BARS * 8;
constexpr int SMEM_BYTES = OFF_TMEM + 16;
constexpr int SMEM_ALLOC = SMEM_BYTES + 1024;           // slack to align the base to 1024 B
constexpr int NTHREADS = 320;                           // producer, MMA issuer, 2 x 4 epilogue warps
constexpr int TMEM_COLS = 512;

struct Params {
  // network
  int n_layers;
  int kblocks[kMaxLayers];
  int n_pad[kMaxLayers];
  const uint8_t* wimg;          // chunk stream
  const float* bias;            // [n_layers][256]
  const float* wdir;            // [4][27] or nullptr
  int k0;                       // true width of the first layer
  int n_out;
  int act;                      // 0 ReLU, 1 ELU
  // io
  int input_mode;
  const float* in0;
  const float* in1;
  int in_stride, in1_stride;
  int S, P;
  long long M;
  float* out;
  int head_lo[4];
  int head_act[3];
  int* error_flag;
  long long* timeline;          // debug: CTA 0 writes clock64() stamps (see TL_* below), nullptr in production
};

// ------------------------------------------------------------------------------------------------ PTX helpers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
// Spin with a watchdog: a protocol bug must not hang the GPU (it traps and reports instead).
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity, int* error_flag, int code) {
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if (++spins > (1u << 22)) {
      if (error_flag) atomicExch(error_flag, code);
      __threadfence_system();
      __trap();
    }
  }
}
__device__ __forceinline__ void bulk_copy_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t cols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(cols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t cols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void umma_bf16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// 32 lanes x 32 columns of fp32: thread t of the warp gets lane (warp%4)*32 + t, columns c..c+31
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float* v) {
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float* v) {
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// UMMA shared-memory descriptor: K-major, 128-byte swizzle, 8-row groups 1024 B apart (dense tile).
__device__ __forceinline__ uint64_t umma_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFFu);          // start address, bits [0,14)
  d |= (uint64_t)1 << 16;                               // leading byte offset (unused for swizzled K-major)
  d |= (uint64_t)(1024u >> 4) << 32;                    // stride byte offset, bits [32,46)
  d |= (uint64_t)1 << 46;                               // descriptor version 1 (sm_100)
  d |= (uint64_t)2 << 61;                               // layout type: SWIZZLE_128B
  return d;
}
// Instruction descriptor, kind::f16: D = f32, A = B = bf16, both K-major, M = 128, N = n.
__device__ __forceinline__ uint32_t umma_idesc(int n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(TILE_M >> 4) << 24);
}

__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
  __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&h);
}
// byte offset of 16-byte chunk c (8 bf16) of row r inside one K block (swizzle: chunk ^= row % 8)
__device__ __forceinline__ uint32_t a_chunk_off(int r, int c) {
  return (uint32_t)((r >> 3) * 1024 + (r & 7) * 128 + ((c ^ (r & 7)) << 4));
}
__device__ __forceinline__ void st_shared_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ void st_shared_b32(uint32_t addr, uint32_t a) {
  asm volatile("st.shared.b32 [%0], %1;" ::"r"(addr), "r"(a) : "memory");
}

// sin / cos of x * 2^l for l = 0..L-1 with one range reduction: t = x / 2pi in turns, scaled exactly by powers
// of two; the fractional turn goes to the SFU.  (bf16 tier: the operand is rounded to 8 bits anyway.)
__device__ __forceinline__ void sincos_octaves(float x, int l, float* s, float* c) {
  float t = x * 0.15915494309189535f * (float)(1 << l);
  float f = t - rintf(t);
  float a = f * 6.283185307179586f;
  *s = __sinf(a);
  *c = __cosf(a);
}

__device__ __forceinline__ float act_apply(float v, int act) {
  if (act == 0) return fmaxf(v, 0.f);
  return v > 0.f ? v : (__expf(v) - 1.f);
}
__device__ __forceinline__ float head_apply_fast(float v, int kind) {
  if (kind == HEAD_SIGMOID) return 1.f / (1.f + __expf(-v));
  if (kind == HEAD_TANH) return tanhf(v);
  return v;
}

// ------------------------------------------------------------------------------------------------ the kernel
// packed fp32x2 add (Blackwell FADD2): {o0,o1} = {x0,x1} + {b0,b1}
__device__ __forceinline__ void add2(float x0, float x1, float b0, float b1, float& o0, float& o1) {
  asm("{\n\t.reg .b64 a, b, c;\n\t"
      "mov.b64 a, {%2, %3};\n\t"
      "mov.b64 b, {%4, %5};\n\t"
      "add.rn.f32x2 c, a, b;\n\t"
      "mov.b64 {%0, %1}, c;\n\t}"
      : "=f"(o0), "=f"(o1)
      : "f"(x0), "f"(x1), "f"(b0), "f"(b1));
}
// two floats -> packed bf16x2 with ReLU folded into the conversion (lo in the low half)
__device__ __forceinline__ uint32_t pack_bf16_relu(float lo, float hi) {
  uint32_t d;
  asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi), "f"(lo));
  return d;
}
__device__ __forceinline__ float elu_fast(float v) {
  // max(v,0) + (exp(min(v,0)) - 1): branch-free, one MUFU.EX2
  float n = fminf(v, 0.f);
  return fmaxf(v, 0.f) + (exp2f(n * 1.4426950408889634f) - 1.f);
}

// One 32-column half of a K block: accumulator columns -> bias + activation -> bf16 -> A operand (4 x 16 B per row).
template <int ACT>
__device__ __forceinline__ void epilogue_half(const float* v, const float4* bias4, uint32_t dst_block, int r, int c0) {
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    const float4 ba = bias4[2 * c], bb = bias4[2 * c + 1];
    const float* x = v + 8 * c;
    float y[8];
    add2(x[0], x[1], ba.x, ba.y, y[0], y[1]);
    add2(x[2], x[3], ba.z, ba.w, y[2], y[3]);
    add2(x[4], x[5], bb.x, bb.y, y[4], y[5]);
    add2(x[6], x[7], bb.z, bb.w, y[6], y[7]);
    uint32_t w[4];
    if (ACT == 0) {
#pragma unroll
      for (int u = 0; u < 4; ++u) w[u] = pack_bf16_relu(y[2 * u], y[2 * u + 1]);
    } else {
#pragma unroll
      for (int u = 0; u < 4; ++u) w[u] = pack_bf16(elu_fast(y[2 * u]), elu_fast(y[2 * u + 1]));
    }
    st_shared_v4(dst_block + a_chunk_off(r, c0 + c), w[0], w[1], w[2], w[3]);
  }
}

// debug timeline slots (second tile of CTA 0): [0..7]*8 layers MMA half issue, epilogue events
constexpr int TL_MMA = 0;        // + l*8 + kb*2 + half          (64)
constexpr int TL_ACC = 64;       // + grp*8 + l                  (16)
constexpr int TL_ARR = 80;       // + grp*64 + l*8 + kb*2 + half (128)
constexpr int TL_N = 208;
__device__ __forceinline__ void tl_mark(long long* tl, bool on, int slot) {
  if (tl && on) tl[slot] = clock64();
}

constexpr int STAGE_BLOCK = 4;          // A block used as the layer-0 operand of 1-block first layers (NeRF), and as
                                        // the fifth K block of the 288-wide sampler input

// ACT: 0 ReLU / 1 ELU.  MODE: InputMode.
template <int ACT, int MODE>
__global__ void __launch_bounds__(NTHREADS, 1) mlp_tc_kernel(Params p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  const uint32_t base = (raw_addr + 1023u) & ~1023u;              // 1024-byte aligned (swizzle atom)
  uint8_t* sm = smem_raw + (base - raw_addr);
  float* s_bias = reinterpret_cast<float*>(sm + OFF_BIAS);
  float* s_wdir = reinterpret_cast<float*>(sm + OFF_WDIR);
  const uint32_t bar0 = base + OFF_BAR;
  auto bar_full = [&](int s) { return bar0 + 8u * s; };
  auto bar_empty = [&](int s) { return bar0 + 8u * (N_SLOTS + s); };
  auto bar_aready = [&](int h) { return bar0 + 8u * (2 * N_SLOTS + h); };            // h = 32-column half block, 0..7
  auto bar_inready = [&]() { return bar0 + 8u * (2 * N_SLOTS + 8); };                // staging block (STAGE_BLOCK)
  auto bar_accfull = [&](int b) { return bar0 + 8u * (2 * N_SLOTS + 9 + b); };
  volatile uint32_t* s_tmem = reinterpret_cast<volatile uint32_t*>(sm + OFF_TMEM);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long n_tiles = (p.M + TILE_M - 1) / TILE_M;
  constexpr bool kStaged = (MODE == IN_ENCODE || MODE == IN_LOAD2);   // 1-block first layer living in STAGE_BLOCK
  const int kb0 = p.kblocks[0];

  // ---- one-time setup ----
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < N_SLOTS; ++s) { mbar_init(bar_full(s), 1); mbar_init(bar_empty(s), 1); }
    for (int h = 0; h < 8; ++h) mbar_init(bar_aready(h), 4);
    mbar_init(bar_inready(), 4);
    mbar_init(bar_accfull(0), 1);
    mbar_init(bar_accfull(1), 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc(smem_u32((const void*)s_tmem), TMEM_COLS);
  if (warp >= 2) {
    const int t = threadIdx.x - 64;
    for (int i = t; i < p.n_layers * kHidden; i += NTHREADS - 64) s_bias[i] = p.bias[i];
    if (p.wdir) for (int i = t; i < 4 * 27; i += NTHREADS - 64) s_wdir[i] = p.wdir[i];
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *s_tmem;
  const int last = p.n_layers - 1;

  if (warp == 0) {
    // =============================== weight producer ===============================
    if (lane == 0) {
      uint32_t g = 0;
      for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const uint8_t* src = p.wimg;
        for (int l = 0; l < p.n_layers; ++l) {
          const uint32_t bytes = (uint32_t)p.n_pad[l] * 128u;
          for (int kb = 0; kb < p.kblocks[l]; ++kb, ++g) {
            const int slot = g % N_SLOTS;
            mbar_wait(bar_empty(slot), ((g / N_SLOTS) & 1) ^ 1, p.error_flag, 1);
            mbar_arrive_expect_tx(bar_full(slot), bytes);
            bulk_copy_g2s(base + OFF_RING + slot * SLOT_BYTES, src, bytes, bar_full(slot));
            src += bytes;
          }
        }
      }
    }
  } else if (warp == 1) {
    // =============================== MMA issuer ===============================
    uint32_t g = 0, layer_ctr = 0;
    uint32_t a_phase = 0;                                  // bit h = parity to wait for on a_ready[h]; bit 8 = in_ready
    for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
      for (int l = 0; l < p.n_layers; ++l, ++layer_ctr) {
        const uint32_t d_tmem = tmem_base + (layer_ctr & 1u) * kHidden;
        const uint32_t idesc = umma_idesc(p.n_pad[l]);
        const int nkb = p.kblocks[l];
        for (int kb = 0; kb < nkb; ++kb, ++g) {
          const int slot = g % N_SLOTS;
          // which A block holds this K block, and which barriers publish it
          const bool staged = (l == 0) && (kStaged || kb == STAGE_BLOCK);
          const int blk = staged ? STAGE_BLOCK : kb;
          mbar_wait(bar_full(slot), (g / N_SLOTS) & 1, p.error_flag, 3);
          const uint64_t a_desc = umma_desc(base + OFF_A + blk * A_BLOCK_BYTES);
          const uint64_t b_desc = umma_desc(base + OFF_RING + slot * SLOT_BYTES);
#pragma unroll
          for (int half = 0; half < 2; ++half) {
            if (staged) {
              if (half == 0) { mbar_wait(bar_inready(), (a_phase >> 8) & 1u, p.error_flag, 2); a_phase ^= 1u << 8; }
            } else {
              const int h = 2 * kb + half;
              mbar_wait(bar_aready(h), (a_phase >> h) & 1u, p.error_flag, 2);
              a_phase ^= 1u << h;
            }
            tc_fence_after();
            if (lane == 0) {
              tl_mark(p.timeline, blockIdx.x == 0 && tile == (long long)gridDim.x, TL_MMA + l * 8 + kb * 2 + half);
#pragma unroll
              for (int s = 2 * half; s < 2 * half + 2; ++s)   // K = 16 per instruction: +32 bytes = +2 in desc units
                umma_bf16(d_tmem, a_desc + 2u * s, b_desc + 2u * s, idesc, (kb | s) != 0 ? 1u : 0u);
            }
            __syncwarp();
          }
          if (lane == 0) {
            umma_commit(bar_empty(slot));                  // slot is free once these MMAs have read it
            if (kb == nkb - 1) umma_commit(bar_accfull(layer_ctr & 1u));
          }
          __syncwarp();
        }
      }
    }
  } else {
    // =============================== epilogue / operand producers ===============================
    // Two groups of 4 warps; group g drains K blocks kb % 2 == g of every trunk layer.  thread = row within a group.
    const int grp = (warp - 2) >> 2;
    const int q = warp & 3;                                // TMEM lane quadrant this warp may access
    const int r = q * 32 + lane;                           // row of the tile owned by this thread
    uint32_t acc_par = 0, layer_ctr = 0;
    const uint32_t a_base = base + OFF_A;

    // layer-0 operand of tile `t` (see MODE); `which` selects the K blocks this group fills
    auto produce_input = [&](long long t) {
      const long long row = t * TILE_M + r;
      const bool live = row < p.M;
      if (MODE == IN_ENCODE) {
        // group 1 only: gamma_10(point) -> STAGE_BLOCK
        float x[3] = {0.f, 0.f, 0.f};
        if (live) { x[0] = p.in0[row * 3]; x[1] = p.in0[row * 3 + 1]; x[2] = p.in0[row * 3 + 2]; }
        float e[64];
        e[0] = x[0]; e[1] = x[1]; e[2] = x[2]; e[63] = 0.f;
#pragma unroll
        for (int l = 0; l < 10; ++l)
#pragma unroll
          for (int c = 0; c < 3; ++c) sincos_octaves(x[c], l, &e[3 + 6 * l + c], &e[6 + 6 * l + c]);
#pragma unroll
        for (int c = 0; c < 8; ++c)
          st_shared_v4(a_base + STAGE_BLOCK * A_BLOCK_BYTES + a_chunk_off(r, c), pack_bf16(e[8 * c], e[8 * c + 1]),
                       pack_bf16(e[8 * c + 2], e[8 * c + 3]), pack_bf16(e[8 * c + 4], e[8 * c + 5]),
                       pack_bf16(e[8 * c + 6], e[8 * c + 7]));
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_inready());
      } else if (MODE == IN_PLUECKER) {
        // 6 Pluecker features of the ray, replicated P times (the P copies agree to 2.4e-7, far below bf16 resolution)
        float f6[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        if (live) {
          const float* ray = p.in0 + row * p.in_stride;
          pluecker6(ray[0], ray[1], ray[2], ray[3], ray[4], ray[5], f6);
        }
        uint32_t pk[3] = {pack_bf16(f6[0], f6[1]), pack_bf16(f6[2], f6[3]), pack_bf16(f6[4], f6[5])};
        const int kmax = 6 * p.P;
        for (int kb = grp; kb < kb0; kb += 2) {
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            uint32_t w[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              int k = kb * 64 + c * 8 + 2 * j;             // even element index; pairs never straddle a feature pair
              w[j] = (k < kmax) ? pk[(k % 6) >> 1] : 0u;
            }
            st_shared_v4(a_base + kb * A_BLOCK_BYTES + a_chunk_off(r, c), w[0], w[1], w[2], w[3]);
          }
          fence_proxy_async();
          __syncwarp();
          if (lane == 0) {
            if (kb == STAGE_BLOCK) mbar_arrive(bar_inready());
            else { mbar_arrive(bar_aready(2 * kb)); mbar_arrive(bar_aready(2 * kb + 1)); }
          }
        }
      } else {
        // IN_LOAD / IN_LOAD2: warp-cooperative coalesced row loads; lane owns elements (2*lane, 2*lane+1) of a K block
        const int k0 = p.k0;
        for (int kb = (kStaged ? 0 : grp); kb < kb0; kb += 2) {
          const int blk = kStaged ? STAGE_BLOCK : kb;
          for (int rr = 0; rr < 32; ++rr) {
            const int trow = q * 32 + rr;
            const long long grow = t * TILE_M + trow;
            const int k = kb * 64 + 2 * lane;
            float v0 = 0.f, v1 = 0.f;
            if (grow < p.M) {
              const float* src = p.in0 + grow * p.in_stride;
              if (k < k0) v0 = __ldg(src + k);
              if (k + 1 < k0) v1 = __ldg(src + k + 1);
            }
            st_shared_b32(a_base + blk * A_BLOCK_BYTES + a_chunk_off(trow, lane >> 2) + (lane & 3) * 4, pack_bf16(v0, v1));
          }
          fence_proxy_async();
          __syncwarp();
          if (lane == 0) {
            if (blk == STAGE_BLOCK) mbar_arrive(bar_inready());
            else { mbar_arrive(bar_aready(2 * kb)); mbar_arrive(bar_aready(2 * kb + 1)); }
          }
        }
      }
    };

    // prologue: first tile's operand (staged modes: group 1 owns the staging block)
    if (!kStaged || grp == 1) produce_input(blockIdx.x);

    for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
      const long long row = tile * TILE_M + r;
      const bool live = row < p.M;
      for (int l = 0; l < p.n_layers; ++l, ++layer_ctr) {
        const uint32_t b = layer_ctr & 1u;
        mbar_wait(bar_accfull(b), (acc_par >> b) & 1u, p.error_flag, 4);
        acc_par ^= 1u << b;
        tc_fence_after();
        const bool tl_on = blockIdx.x == 0 && tile == (long long)gridDim.x && q == 0 && lane == 0;
        tl_mark(p.timeline, tl_on, TL_ACC + grp * 8 + l);
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + b * kHidden;
        if (l < last) {
          const float4* bl4 = reinterpret_cast<const float4*>(s_bias + l * kHidden);
#pragma unroll 1
          for (int kb = grp; kb < 4; kb += 2) {            // this group's K blocks of the next layer's operand
            float v[64];
            tmem_ld32(taddr + kb * 64, v);
            tmem_ld32(taddr + kb * 64 + 32, v + 32);
            float4 bias4[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) bias4[i] = bl4[kb * 16 + i];       // overlaps the TMEM load latency
            tmem_wait_ld();
            const uint32_t dst = a_base + kb * A_BLOCK_BYTES;
#pragma unroll
            for (int half = 0; half < 2; ++half) {
              epilogue_half<ACT>(v + 32 * half, bias4 + 8 * half, dst, r, 4 * half);
              fence_proxy_async();
              tc_fence_before();
              __syncwarp();
              if (lane == 0) mbar_arrive(bar_aready(2 * kb + half));
              tl_mark(p.timeline, tl_on, TL_ARR + grp * 64 + l * 8 + kb * 2 + half);
            }
          }
          // staged first layers: group 1 builds the NEXT tile's layer-0 operand while the MMAs of this tile run
          // (the staging block is free once layer 0's accumulator has been observed complete)
          if (kStaged && grp == 1 && l == 1 && tile + gridDim.x < n_tiles) produce_input(tile + gridDim.x);
        } else if (grp == 0) {
          // output layer: n_out <= 48 columns of the accumulator
          float v[48];
          const int npad = p.n_pad[last];
          tmem_ld16(taddr, v);
          if (npad > 16) tmem_ld16(taddr + 16, v + 16);
          if (npad > 32) tmem_ld16(taddr + 32, v + 32);
          tmem_wait_ld();
          tc_fence_before();
          const float* bo = s_bias + last * kHidden;
          if (kStaged) {
            // view-direction term of DoNeRFTRT's last layer: W7[:, 256:283] . gamma_4(viewdir)
            float g[27];
            if (MODE == IN_LOAD2) {
#pragma unroll
              for (int i = 0; i < 27; ++i) g[i] = live ? p.in1[row * 27 + i] : 0.f;
            } else {
              float d[3] = {0.f, 0.f, 0.f};
              if (live) {
                const float* vd = p.in1 + (row / p.S) * p.in1_stride;
                d[0] = vd[0]; d[1] = vd[1]; d[2] = vd[2];
              }
              g[0] = d[0]; g[1] = d[1]; g[2] = d[2];
#pragma unroll
              for (int lv = 0; lv < 4; ++lv)
#pragma unroll
                for (int c = 0; c < 3; ++c) sincos_octaves(d[c], lv, &g[3 + 6 * lv + c], &g[6 + 6 * lv + c]);
            }
#pragma unroll
            for (int o = 0; o < 4; ++o) {
              float a = v[o] + bo[o];
#pragma unroll
              for (int i = 0; i < 27; ++i) a = fmaf(s_wdir[o * 27 + i], g[i], a);
              v[o] = a;
            }
            if (live) *reinterpret_cast<float4*>(p.out + row * 4) = make_float4(v[0], v[1], v[2], v[3]);
          } else if (live) {
            float* orow = p.out + row * p.n_out;
#pragma unroll
            for (int o = 0; o < 48; ++o) {
              if (o < p.n_out) {
                int kind = HEAD_NONE;
#pragma unroll
                for (int gq = 0; gq < 3; ++gq)
                  if (o >= p.head_lo[gq] && o < p.head_lo[gq + 1]) kind = p.head_act[gq];
                orow[o] = head_apply_fast(v[o] + bo[o], kind);
              }
            }
          }
        }
      }
      // non-staged modes: next tile's operand goes into blocks the output layer has just finished reading
      if (!kStaged && tile + gridDim.x < n_tiles) produce_input(tile + gridDim.x);
    }
  }

  // ---- teardown ----
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, TMEM_COLS);
}

// ------------------------------------------------------------------------------------------------ weight packing
// W [out][in] fp32 -> stream of K-block chunks, each [n_pad rows][64 k] bf16 in the UMMA K-major 128B-swizzle layout
__global__ void pack_tc_kernel(const float* __restrict__ W, int out_dim, int in_dim, int k_used, int n_pad, int kblocks,
                               uint8_t* __restrict__ dst) {
  const int total = kblocks * n_pad * 64;
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
    const int kb = idx / (n_pad * 64);
    const int rem = idx - kb * n_pad * 64;
    const int n = rem >> 6, k = rem & 63;
    const int ks = kb * 64 + k;
    const float v = (n < out_dim && ks < k_used) ? W[(size_t)n * in_dim + ks] : 0.f;
    const size_t off = (size_t)kb * n_pad * 128 + (size_t)(n >> 3) * 1024 + (n & 7) * 128 + (((k >> 3) ^ (n & 7)) << 4) + (k & 7) * 2;
    *reinterpret_cast<__nv_bfloat16*>(dst + off) = __float2bfloat16_rn(v);
  }
}

__global__ void pack_tc_bias_kernel(const float* __restrict__ b, int out_dim, float* __restrict__ dst) {
  int i = threadIdx.x;
  if (i < kHidden) dst[i] = i < out_dim ? b[i] : 0.f;
}

__global__ void pack_tc_wdir_kernel(const float* __restrict__ W, int in_dim, float* __restrict__ dst) {
  int i = threadIdx.x;                                     // 4 x 27
  if (i < 4 * 27) dst[i] = W[(size_t)(i / 27) * in_dim + kHidden + (i % 27)];
}

}  // namespace tc

// ------------------------------------------------------------------------------------------------ host side
static long long* g_tc_timeline = nullptr;
void tc_set_timeline(long long* dev_buf) { g_tc_timeline = dev_buf; }

struct TcLayout {
  int kblocks[kMaxLayers];
  int n_pad[kMaxLayers];
  size_t chunk_off[kMaxLayers];
  size_t img_bytes, bias_off, wdir_off, total;
};

static TcLayout tc_layout(int net_id, int n_layers, const int* in_dims, const int* out_dims) {
  TcLayout L{};
  size_t off = 0;
  for (int l = 0; l < n_layers; ++l) {
    const bool last = l == n_layers - 1;
    int k_used = in_dims[l];
    if (last && net_id == PN_NET_NERF) k_used = kHidden;    // the 27 view-direction inputs go through wdir
    L.kblocks[l] = (k_used + 63) / 64;
    L.n_pad[l] = last ? (out_dims[l] + 15) / 16 * 16 : kHidden;
    L.chunk_off[l] = off;
    off += (size_t)L.kblocks[l] * L.n_pad[l] * 128;
  }
  L.img_bytes = off;
  L.bias_off = (off + 255) & ~(size_t)255;
  L.wdir_off = L.bias_off + (size_t)n_layers * kHidden * 4;
  L.total = L.wdir_off + 4 * 28 * 4;
  return L;
}

void tc_free_net(NetTC& n) {
  if (n.blob) cudaFree(n.blob);
  if (n.error_flag) cudaFree(n.error_flag);
  n = NetTC();
}

bool tc_available() { return true; }

int tc_load_net(NetTC& n, int net_id, int n_layers, const int* in_dims, const int* out_dims, const float* const* W,
                const float* const* b, cudaStream_t stream) {
  tc_free_net(n);
  n.net_id = net_id;
  n.n_layers = n_layers;
  for (int l = 0; l < n_layers; ++l) { n.in_dim[l] = in_dims[l]; n.out_dim[l] = out_dims[l]; }
  n.supported = in_dims[0] <= tc::MAX_KB * 64 && out_dims[n_layers - 1] <= 48;
  if (!n.supported) return PN_OK;                           // fp32 tier still works; bf16 launch reports it
  TcLayout L = tc_layout(net_id, n_layers, in_dims, out_dims);
  PN_CUDA_OK(cudaMalloc(&n.blob, L.total));
  PN_CUDA_OK(cudaMalloc((void**)&n.error_flag, sizeof(int)));
  PN_CUDA_OK(cudaMemsetAsync(n.error_flag, 0, sizeof(int), stream));
  n.blob_bytes = L.total;
  uint8_t* blob = reinterpret_cast<uint8_t*>(n.blob);
  for (int l = 0; l < n_layers; ++l) {
    const bool last = l == n_layers - 1;
    int k_used = (last && net_id == PN_NET_NERF) ? kHidden : in_dims[l];
    int total = L.kblocks[l] * L.n_pad[l] * 64;
    tc::pack_tc_kernel<<<(total + 255) / 256, 256, 0, stream>>>(W[l], out_dims[l], in_dims[l], k_used, L.n_pad[l], L.kblocks[l],
                                                                 blob + L.chunk_off[l]);
    PN_LAUNCH_OK("pack_tc_kernel");
    tc::pack_tc_bias_kernel<<<1, kHidden, 0, stream>>>(b[l], out_dims[l], reinterpret_cast<float*>(blob + L.bias_off) + (size_t)l * kHidden);
    PN_LAUNCH_OK("pack_tc_bias_kernel");
  }
  if (net_id == PN_NET_NERF) {
    tc::pack_tc_wdir_kernel<<<1, 128, 0, stream>>>(W[n_layers - 1], in_dims[n_layers - 1], reinterpret_cast<float*>(blob + L.wdir_off));
    PN_LAUNCH_OK("pack_tc_wdir_kernel");
  }
  n.loaded = true;
  return PN_OK;
}

int tc_launch_mlp(const NetTC& n, const MlpLaunch& Lc, cudaStream_t stream) {
  if (!n.loaded) {
    set_error(n.supported ? "PN_PREC_BF16: network weights not loaded" : "PN_PREC_BF16: this network shape is outside the tensor-core "
              "kernel's limits (first layer <= 320 inputs, output <= 48); use PN_PREC_FP32");
    return PN_ESTATE;
  }
  if (Lc.M == 0) return PN_OK;
  TcLayout L = tc_layout(n.net_id, n.n_layers, n.in_dim, n.out_dim);
  tc::Params p{};
  p.n_layers = n.n_layers;
  for (int l = 0; l < n.n_layers; ++l) { p.kblocks[l] = L.kblocks[l]; p.n_pad[l] = L.n_pad[l]; }
  const uint8_t* blob = reinterpret_cast<const uint8_t*>(n.blob);
  p.wimg = blob;
  p.bias = reinterpret_cast<const float*>(blob + L.bias_off);
  p.wdir = n.net_id == PN_NET_NERF ? reinterpret_cast<const float*>(blob + L.wdir_off) : nullptr;
  p.k0 = n.in_dim[0];
  p.n_out = n.out_dim[n.n_layers - 1];
  p.act = Lc.act;
  p.input_mode = Lc.input_mode;
  p.in0 = Lc.in0; p.in1 = Lc.in1; p.in_stride = Lc.in_stride; p.in1_stride = Lc.in1_stride;
  p.S = Lc.S > 0 ? Lc.S : 1; p.P = Lc.P; p.M = Lc.M; p.out = Lc.out;
  for (int i = 0; i < 4; ++i) p.head_lo[i] = Lc.head_lo[i];
  for (int i = 0; i < 3; ++i) p.head_act[i] = Lc.head_act[i];
  p.error_flag = n.error_flag;
  p.timeline = g_tc_timeline;
  if (Lc.input_mode == IN_PLUECKER && 6 * Lc.P != n.in_dim[0]) { set_error("tc sampler: 6P != first-layer width"); return PN_EINVAL; }
  int dev = 0, sms = 0;
  PN_CUDA_OK(cudaGetDevice(&dev));
  PN_CUDA_OK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  long long tiles = (Lc.M + tc::TILE_M - 1) / tc::TILE_M;
  unsigned grid = (unsigned)(tiles < sms ? tiles : sms);
#define PN_TC_LAUNCH(ACT, MODE)                                                                                            \
  do {                                                                                                                     \
    PN_CUDA_OK(cudaFuncSetAttribute(tc::mlp_tc_kernel<ACT, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::SMEM_ALLOC)); \
    tc::mlp_tc_kernel<ACT, MODE><<<grid, tc::NTHREADS, tc::SMEM_ALLOC, stream>>>(p);                                      \
  } while (0)
  if (Lc.act == 0 && Lc.input_mode == IN_ENCODE) PN_TC_LAUNCH(0, IN_ENCODE);
  else if (Lc.act == 0 && Lc.input_mode == IN_LOAD2) PN_TC_LAUNCH(0, IN_LOAD2);
  else if (Lc.act == 1 && Lc.input_mode == IN_PLUECKER) PN_TC_LAUNCH(1, IN_PLUECKER);
  else if (Lc.act == 1 && Lc.input_mode == IN_LOAD) PN_TC_LAUNCH(1, IN_LOAD);
  else { set_error("tc: unsupported (activation, input mode) = (%d, %d)", Lc.act, Lc.input_mode); return PN_EINVAL; }
#undef PN_TC_LAUNCH
  PN_LAUNCH_OK("mlp_tc_kernel");
  return PN_OK;
}

}  // namespace pn
