// Tensor-core tier of the three MLPs: a persistent, warp-specialised, fully fused kernel on tcgen05.mma
// (fp16 operands, fp32 accumulators in TMEM) running on CTA PAIRS (cta_group::2) with two row tiles in flight.
//
//   sampler  MinMaxRaySamplerTRT_Net     helpers.py:1473-1507   288 -> 256 x6 (ELU) -> 27
//   refine   MinMaxRayEpiSamplerTRT_Net  helpers.py:1509-1540   144 -> 256 x6 (ELU) -> 35
//   NeRF     DoNeRFTRT                   helpers.py:1186-1343   63 -> 256 x7 (ReLU) -> [256 ++ 27] -> 4
//
// A cluster of two CTAs (one per SM of a TPC) walks 512-row units.  Each CTA owns two 128-row tiles ("slots" X and
// Y); one tcgen05.mma.cta_group::2 instruction multiplies the 256 rows of a slot pair (128 from each CTA) by a
// [256 x 16] weight step whose rows are split between the two CTAs' shared memories, so every SM reads half of the
// weights and the instruction stream is issued once per pair (by the leader CTA's MMA thread).
//
//   D_slot[256 x N] (fp32, TMEM, 128 lanes per CTA) = A_slot[256 x K] (fp16, smem, K-major, 128B swizzle) * W^T
//
// Pipeline (per slot the layers are strictly sequential; the two slots run in step, so the tensor pipe runs slot Y's layer
// while the CUDA cores drain slot X's accumulator).  Two CTA shapes (use_handover()):
//
//   ELU kernels (sampler, refine) -- 16 warps, hand-over warps:
//   warp 12     weight producer (both CTAs): streams its half of every K block (16 KB) of every layer, in
//               consumption order, through a 5-slot ring with cp.async.bulk + mbarrier complete_tx; the global image
//               is pre-swizzled into the UMMA canonical layout so a linear copy lands a ready B operand.
//   warp 13     leader: MMA issuer -- per (layer, slot): wait "operand ready", then per K block wait "weights
//               landed in both CTAs", issue 4 MMAs (M256 N256 K16), tcgen05.commit frees the ring slot in both CTAs;
//               after the last block commit "accumulator full" (hidden layers) or "output full" to both CTAs.
//               follower: relay -- forwards "my half of the weights landed" to the leader's full barrier.
//   warps 0-7   hidden epilogues (8 warps; warp = TMEM lane quadrant x 128-column half; thread = row):
//               tcgen05.ld the accumulator, bias + ELU, convert to fp16 and store straight into the slot's
//               A operand for the next layer (the swizzle makes row-per-thread 16-byte stores conflict free), then
//               fence.proxy.async and arrive on the leader's "operand ready" barrier (remote arrive from the follower).
//   warps 8-11  tile hand-over (one warp per TMEM lane quadrant; thread = row): everything that is NOT a hidden epilogue --
//               the output layer's accumulator, the head activations, the stores to global memory, and the NEXT tile's
//               first-layer operand (generated in registers while the slot's last layers run, or fetched by TMA: three
//               cp.async.bulk.tensor boxes per tile), published a few hundred cycles after "output full".  Both slots'
//               hand-overs come first, the head outputs afterwards.  The hidden-epilogue warps never see a tile boundary:
//               they go from the last hidden layer of one tile straight to the first hidden layer of the next.
//   warps 14-15 exist to lend their registers (setmaxnreg works on whole warpgroups): the CTA launches with 128 registers
//               per thread, warps 12-15 drop to 48, warps 8-11 keep 128, warps 0-7 grow to 168.
//
//   ReLU kernels (DoNeRFTRT, classic NeRF) -- 10 warps at 168 registers: warps 0-7 as above plus the output phase (16 output
//               columns, one float4 per row, the next tile's frequency encoding), warp 8 weight producer, warp 9 MMA issuer.
//
// The first-layer operand is generated in the kernel (frequency encoding, Pluecker features) or loaded; the NeRF
// view-direction term (27 inputs of the last layer, identical for a ray's samples) is a per-ray fp32 pre-pass added in
// the output epilogue, so the tensor-core part of the last layer is a clean K = 256, N = 16 GEMM.
#include <cuda.h>          // CUtensorMap (types only: the encoder is fetched through cudaGetDriverEntryPoint, libcuda is not linked)
#include <cuda_fp16.h>

#include <cstdlib>
#include <type_traits>

#include "tc.cuh"

namespace pn {
namespace tc {

constexpr int TILE_M = 128;                             // rows per CTA per slot
constexpr int PAIR_M = 2 * TILE_M;                      // rows per MMA (cta_group::2)
constexpr int UNIT_M = 2 * PAIR_M;                      // rows per cluster iteration (two slots)
constexpr int A_BLOCK_BYTES = TILE_M * 128;             // one 64-wide K block of a slot's operand: 16 KB
constexpr int A_SLOT_BYTES = 4 * A_BLOCK_BYTES;         // 64 KB
constexpr int RING_SLOT_BYTES = (kHidden / 2) * 128;    // this CTA's half of a 256-row weight K block: 16 KB
constexpr int N_RING = 5;
constexpr int BIAS_ROWS = 14;                           // 8 layer biases (+ classic NeRF: 3 more, the alpha weights, 2 rows of alpha partial sums)
constexpr int BIAS_FLOATS = BIAS_ROWS * kHidden;        // 14 KB
// Output staging (sampler / refine heads): a warp's 32 rows x n_out floats are one contiguous global range, but thread = row,
// so direct stores scatter 4 bytes over 32 sectors per instruction.  Each ch = 0 warp transposes through a private 2304-byte
// window (rows_per_pass x n_out floats) and writes 16 bytes per lane.  The window overlays bias rows 8.. (classic NeRF only,
// whose float4 output needs no staging) and extends past the bias table.
constexpr int STAGE_BIAS_ROW0 = 8;
constexpr int STAGE_FLOATS = 576;                       // per warp: 16 rows x 36 floats
constexpr int STAGE_BYTES = 4 * STAGE_FLOATS * 4;       // 9216
constexpr int BIAS_REGION = (STAGE_BIAS_ROW0 * kHidden * 4 + STAGE_BYTES > BIAS_FLOATS * 4) ? STAGE_BIAS_ROW0 * kHidden * 4 + STAGE_BYTES : BIAS_FLOATS * 4;
constexpr int OFF_A = 0;
constexpr int OFF_RING = OFF_A + 2 * A_SLOT_BYTES;
constexpr int OFF_BIAS = OFF_RING + N_RING * RING_SLOT_BYTES;
constexpr int OFF_STAGE = OFF_BIAS + STAGE_BIAS_ROW0 * kHidden * 4;
constexpr int OFF_BAR = OFF_BIAS + BIAS_REGION;
// barriers (8 bytes each): full[N_RING], empty[N_RING], a_ready[2 slots][2 halves], acc_full[2], out_full[2], in_full[2]
constexpr int N_BARS = 2 * N_RING + 10;
constexpr int OFF_TMEM = OFF_BAR + N_BARS * 8;
constexpr int SMEM_BYTES = OFF_TMEM + 16;
constexpr int SMEM_ALLOC = SMEM_BYTES + 1024;           // slack to align the base to 1024 B (swizzle atom)
static_assert(SMEM_ALLOC <= 232448, "over the 227 KB per-CTA shared-memory limit of sm_100");
constexpr int N_EPI_WARPS = 8;                          // hidden epilogues: lane quadrant x column half (warpgroups 0, 1)
constexpr int N_OUT_WARPS = 4;                          // tile hand-over: one warp per lane quadrant (warpgroup 2)
constexpr int W_OUT0 = N_EPI_WARPS;                     // warps 8..11
// Two shapes of the CTA.  The ELU kernels (sampler, refine) run with HAND-OVER WARPS: 16 warps, the register split below.  The ReLU
// kernels (DoNeRFTRT, classic NeRF) keep 10 warps at 168 registers with the output phase on the epilogue warps: their hand-over is
// a 16-column accumulator read and one float4 per row, the kernel runs into the board's power limit (a train of launches draws
// 990 W of 1000 and the SM clock sits at 1.45-1.5 GHz), and the 16-warp shape measured 1.5-1.9 % MORE time per launch there for 1 %
// fewer cycles (profiles/r02_handover).
__host__ __device__ constexpr bool use_handover(int act) { return act == 1; }
__host__ __device__ constexpr int n_threads(int act) { return use_handover(act) ? 16 * 32 : (N_EPI_WARPS + 2) * 32; }
// Register budget.  The register file is per SM sub-partition (16 K registers = 512 per lane), a sub-partition holds one warp of
// every warpgroup, and setmaxnreg moves registers between warpgroups INSIDE the launch allocation (512 threads x 128 = all of
// it): 2 x 168 (hidden epilogues: 128 live accumulator columns) + 128 (hand-over: both slots' 48 output columns live) + 48 (roles) = 512.
// ptxas budgets a region by the setmaxnreg that dominates it, and fails outright (C7600) when a role region cannot fit.
constexpr int kEpiRegs = 168;
constexpr int kOutRegs = 128;
constexpr int kRoleRegs = 48;
static_assert(2 * kEpiRegs + kOutRegs + kRoleRegs <= 512, "register budget per SM sub-partition lane");
// The two single-thread roles get the HIGHEST active warp ids: the SM's warp arbiter favours higher ids, and a starved MMA
// issuer (or weight producer) stalls the whole pair (measured: 2x slower issue as warp 1 behind four epilogue warps).
__host__ __device__ constexpr int w_producer(int act) { return use_handover(act) ? W_OUT0 + N_OUT_WARPS : N_EPI_WARPS; }   // warp 12 / 8 (scheduler 0)
constexpr int TMEM_COLS = 512;
constexpr int MAX_PHASES = 14;

enum EpiKind : int { EPI_HIDDEN = 0, EPI_MORE = 1, EPI_OUT = 2 };
enum MoreSrc : int { MORE_LOAD = 0, MORE_PTS = 1, MORE_DIRS = 2 };   // what an EPI_MORE phase writes into K block 0

// One (layer, K range) step of a slot: MMAs over nkb K blocks, then an epilogue.
struct Phase {
  int layer;        // bias row / weight layer
  int nkb;          // K blocks consumed
  int k16_last;     // K=16 steps in the last block (1..4)
  int n_pad;        // MMA N (multiple of 16)
  int acc;          // 1: accumulate onto what the previous phase left in TMEM
  int epi;          // EpiKind
  int merged;       // 1: all nkb K blocks of this (narrow) layer travel as ONE ring chunk [rank][kb][n_pad/2 rows x 128 B]
  uint32_t w_off;   // byte offset of the first chunk in the weight image
  int act_none;     // hidden epilogue without activation (classic NeRF: feature_linear)
  int more_src;     // EPI_MORE: MoreSrc
  int side;         // hidden epilogue also accumulates the alpha head's dot product (classic NeRF: pts_linears.7)
};

struct Params {
  int n_phases;
  Phase ph[MAX_PHASES];
  const uint8_t* wimg;
  const float* bias;            // [n_bias_rows][256]
  int n_layers;
  int n_bias_rows;
  int n_out;
  int k0;                       // width of the loaded first-layer operand (IN_LOAD / IN_LOAD2)
  const float* in0;
  int in_stride;
  const float* dirterm;         // NeRF: [M / dir_div][4] fp32 view-direction term of the last layer (no bias)
  int dir_div;
  const float* vdir;            // classic NeRF: per-ray view directions [M / dir_div][vdir_stride], encoded in-kernel
  int vdir_stride;
  int alpha_row;                // classic NeRF: bias-table row holding alpha_linear's 256 weights; rows +1, +2 = partial sums
  float alpha_bias;
  long long M;
  float* out;
  float4 head_tab[96];          // output column c: y = x * .w + (.y / (1 + 2^(x * .x)) + .z), x = accumulator + bias (head_coeffs)
  uint32_t head_linear;         // bit b: output columns [8b, 8b + 8) are all linear (or padding): y = x
  alignas(64) CUtensorMap tmap_in;   // IN_LOAD16: the [M, K0] fp16 input as a 2-D tensor, box = 64 columns x 128 rows, 128-byte swizzle
  int share;                    // 1: a layer's weight blocks are streamed ONCE per unit and multiplied into both slots (see "weight sharing" below)
  int split;                    // hidden epilogues publish their first two K blocks early (two-step operand hand-over): bit 0 = all, bit 1 = first layer only
  int out_rpp;                  // rows per staging pass of the head output (multiple of 4)
  int* error_flag;
  long long* timeline;          // debug: leader CTA of cluster 0 stamps clock64() of its second iteration
  long long* clk;               // measurement aid (pn_debug_tc_clock): CTA 0 writes {clock64, %globaltimer ns} after setup and before teardown
};

// ------------------------------------------------------------------------------------------------ PTX helpers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// arrive on the barrier at the same shared-memory offset in CTA `rank` of the cluster.  Default semantics (release at
// CTA scope), as CUTLASS's ClusterBarrier does: a cluster-scope release would cost a MEMBAR.GPU per arrive, and
// everything published through these barriers lives in the arriving CTA's own shared memory (performed locally,
// made visible to the tensor-core proxy by fence.proxy.async beforehand).
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t bar, uint32_t rank) {
  asm volatile(
      "{\n\t.reg .b32 ra;\n\t"
      "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
      "mbarrier.arrive.shared::cluster.b64 _, [ra];\n\t}" ::"r"(bar), "r"(rank)
      : "memory");
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
// Spin with a watchdog: a protocol bug must not hang the GPU (it traps and reports instead).  Inlined: once the warpgroups'
// register budgets differ (setmaxnreg) ptxas cannot allocate across a real call (C7600).
__device__ __forceinline__ void mbar_wait_slow(uint32_t bar, uint32_t parity, int* error_flag, int code) {
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if (++spins > (1u << 22)) {
      if (error_flag) atomicExch(error_flag, code);
      __threadfence_system();
      __trap();
    }
  }
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity, int* error_flag, int code) {
  if (!mbar_try_wait(bar, parity)) mbar_wait_slow(bar, parity, error_flag, code);
}
// The hand-over warps wait for most of a tile's lifetime: they poll with a back-off (32 ns; 500 ns - 4 us measured the same).
__device__ __forceinline__ void mbar_wait_polite(uint32_t bar, uint32_t parity, int* error_flag, int code) {
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    __nanosleep(32);
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
// One [128 rows x 64 halves] box of a row-major fp16 tensor -> one K block of an A operand.  The tensor map's 128-byte swizzle is the
// UMMA K-major SWIZZLE_128B layout (16-byte chunk index ^ row % 8), columns / rows outside the tensor arrive as zeros, and the whole
// box (16 KB) counts towards the barrier's transaction bytes.
__device__ __forceinline__ void tma_load_box(uint32_t dst, const CUtensorMap* tmap, int col, int row, uint32_t bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(dst),
               "l"(reinterpret_cast<uint64_t>(tmap)), "r"(col), "r"(row), "r"(bar)
               : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t cols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(cols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t cols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
}
// arrive (count 1) on the barrier at this offset in BOTH CTAs of the pair once all prior MMAs have completed
__device__ __forceinline__ void umma_commit_pair(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
               "h"((uint16_t)3)
               : "memory");
}
__device__ __forceinline__ void umma_f16_pair(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
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
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
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
// Instruction descriptor, kind::f16: D = f32 (bit 4), A = B = f16 (format 0), both K-major, M = 256 (pair), N = n.
__device__ __forceinline__ uint32_t umma_idesc(int n) {
  return (1u << 4) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(PAIR_M >> 4) << 24);
}

// byte offset of 16-byte chunk c (8 halves) of row r inside one K block (swizzle: chunk ^= row % 8)
__device__ __forceinline__ uint32_t a_chunk_off(int r, int c) {
  return (uint32_t)((r >> 3) * 1024 + (r & 7) * 128 + ((c ^ (r & 7)) << 4));
}
// Operand-buffer stores.  No "memory" clobber on purpose: these only ever write the A operand buffers, which no C++-level
// access touches, and a clobber would pin every bias load (plain shared-memory reads) behind the previous store -- one
// exposed LDS latency per 16-byte chunk.  Ordering against the tensor-core proxy comes from fence_proxy_async() (a
// volatile asm WITH a clobber) in publish().
__device__ __forceinline__ void st_shared_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d));
}
__device__ __forceinline__ void st_shared_b32(uint32_t addr, uint32_t a) {
  asm volatile("st.shared.b32 [%0], %1;" ::"r"(addr), "r"(a));
}

// two floats -> packed f16x2 (lo in the low half), saturating to the largest finite half
__device__ __forceinline__ uint32_t pack_h2(float lo, float hi) {
  uint32_t d;
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi), "f"(lo));
  return d;
}
__device__ __forceinline__ uint32_t pack_h2_relu(float lo, float hi) {
  uint32_t d;
  asm("cvt.rn.relu.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi), "f"(lo));
  return d;
}
// ELU on a packed pair, in half precision: max(h, 2^(min(h,0) * log2 e) - 1)   (for h > 0 the right side is 0 < h;
// for h <= 0, e^h - 1 >= h).  One MUFU.EX2 per element; everything else is packed half2 arithmetic.
__device__ __forceinline__ uint32_t pack_h2_elu(float lo, float hi) {
  const uint32_t h = pack_h2(lo, hi);
  uint32_t m, t, e, n, r;
  asm("min.f16x2 %0, %1, %2;" : "=r"(m) : "r"(h), "r"(0u));
  asm("mul.rn.f16x2 %0, %1, %2;" : "=r"(t) : "r"(m), "r"(0x3DC53DC5u));      // log2(e) = 1.4427 -> 0x3DC5
  asm("ex2.approx.f16x2 %0, %1;" : "=r"(e) : "r"(t));
  asm("add.rn.f16x2 %0, %1, %2;" : "=r"(n) : "r"(e), "r"(0xBC00BC00u));       // - 1.0
  asm("max.f16x2 %0, %1, %2;" : "=r"(r) : "r"(h), "r"(n));
  return r;
}
// ELU in fp32: max(y, 2^(-|y| log2 e) - 1).  For y <= 0 that is e^y - 1 (>= y); for y > 0 the right side is negative.  The
// -|.| rides on the MUFU operand modifiers, and FMUL / FADD / FMNMX issue at twice the rate of the packed-half ops
// (measured: scripts/ubench/alu_rate.cu), so this beats the f16x2 form (whose ex2 is two MUFU ops anyway) and is exact to fp32.
__device__ __forceinline__ float elu_f32(float y) {
  float e;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(-fabsf(y * 1.4426950408889634f)));
  return fmaxf(y, e - 1.f);
}
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

// Head activations of the tensor-core tier as one branch-free form: y = c x + (a / (1 + 2^(s x)) + b), per-column
// coefficients in the kernel parameters (constant-bank operands, no loads):
//   none: (0, 0, 0, 1)    sigmoid x = 1 / (1 + 2^(-x log2 e)): (-log2 e, 1, 0, 0)    tanh x = 2 sigmoid(2x) - 1: (-2 log2 e, 2, -1, 0)
// on the SFU (ex2 + rcp, ~1e-6 relative; 2^inf = inf gives 0).  With a run-time `kind` test per column every column was
// its own basic block (constant loads, compares, branches) and the lone output warp of a scheduler walked ~30 dependent
// chains one after the other: 8 K cycles per 128-row tile for 27 columns (timeline stamps TL_OUT).
__device__ __forceinline__ float head_apply_tab(float x, const float4& h) {
  float e, r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(x * h.x));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(1.f + e));
  return fmaf(h.w, x, fmaf(h.y, r, h.z));
}
static float4 head_coeffs(int kind) {
  const float l2e = 1.4426950408889634f;
  if (kind == HEAD_SIGMOID) return make_float4(-l2e, 1.f, 0.f, 0.f);
  if (kind == HEAD_TANH) return make_float4(-2.f * l2e, 2.f, -1.f, 0.f);
  return make_float4(0.f, 0.f, 0.f, 1.f);
}

// The padded frequency encoding gamma_10(x) = [x, sin(2^l x), cos(2^l x)]_{l<10}, 0 (helpers.py:666-671), 64 elements per
// point; element k >= 3 is (l, sin|cos, coordinate) = ((k-3)/6, ((k-3)%6)/3, (k-3)%3).  Thread (row, CH) produces elements
// [32 CH, 32 CH + 32): octaves 0..4 (CH = 0) or 4..9 (CH = 1).  Per coordinate ONE exact range reduction in turns
// (x * 2^l / 2pi; the power-of-two scaling is exact) and one SFU sin/cos at the first octave, then the double-angle
// recurrence sin 2a = 2 sin a cos a, cos 2a = 1 - 2 sin^2 a: the error doubles per octave (<= 2^5 * 4e-7), far below
// fp16 resolution, and the SFU work drops from 60 to 6 operations per point.
template <int CH>
__device__ __forceinline__ void encode32(const float* x, uint32_t* w) {
  constexpr int L0 = CH ? 4 : 0, NL = CH ? 6 : 5;
  float sn[3][NL], cs[3][NL];
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    float t = x[c] * 0.15915494309189535f * (float)(1 << L0);
    t -= rintf(t);
    sn[c][0] = __sinf(t * 6.283185307179586f);
    cs[c][0] = __cosf(t * 6.283185307179586f);
#pragma unroll
    for (int i = 1; i < NL; ++i) {
      sn[c][i] = 2.f * sn[c][i - 1] * cs[c][i - 1];
      cs[c][i] = fmaf(-2.f * sn[c][i - 1], sn[c][i - 1], 1.f);
    }
  }
  float e[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) {
    const int k = 32 * CH + i;
    if (k < 3) e[i] = x[k];
    else if (k >= 63) e[i] = 0.f;
    else {
      const int j = k - 3, l = j / 6, rem = j % 6, c = rem % 3;
      e[i] = rem >= 3 ? cs[c][l - L0] : sn[c][l - L0];
    }
  }
#pragma unroll
  for (int i = 0; i < 16; ++i) w[i] = pack_h2(e[2 * i], e[2 * i + 1]);
}

#ifndef PN_TC_TIMELINE
#define PN_TC_TIMELINE 0          // build with -DPN_TC_TIMELINE=1 (PN_TC_TIMELINE=1 python -m pronerf_b200.build) to compile the stamps in
#endif
constexpr bool kTimeline = PN_TC_TIMELINE != 0;
__device__ __forceinline__ float4 ld_shared_v4f(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
  return v;
}
// 64 accumulator columns (already in registers) -> bias + activation -> fp16 -> one K block of the next layer's operand
// (8 x 16 B of this thread's row).  row_base = block + row offset, xr = (r & 7) << 4 (the 128B swizzle).
// ptxas cannot tell the bias reads from the operand stores apart (both shared memory), so it never hoists a bias load
// above an earlier store: the loads are issued by hand two chunks ahead (volatile asm keeps the order).
// ACT: 0 ReLU, 1 ELU, 2 none.  SIDE: also return sum_c relu(y_c) * wa[c] over the 64 columns (wa = shared-memory address of
// 64 fp32 weights), evaluated on the fp32 activations before they are rounded to fp16 (classic NeRF's alpha head).
template <int ACT, bool SIDE = false>
__device__ __forceinline__ float epilogue_store64(const float* v, uint32_t bias_addr, uint32_t row_base, uint32_t xr, uint32_t wa_addr = 0u) {
  float4 b[20];
  float side = 0.f;
#pragma unroll
  for (int i = 0; i < 4; ++i) b[i] = ld_shared_v4f(bias_addr + 16u * i);
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    if (c + 2 < 8) {
      b[2 * c + 4] = ld_shared_v4f(bias_addr + 16u * (2 * c + 4));
      b[2 * c + 5] = ld_shared_v4f(bias_addr + 16u * (2 * c + 5));
    }
    const float4 ba = b[2 * c], bb = b[2 * c + 1];
    const float* x = v + 8 * c;
    float y[8];
    add2(x[0], x[1], ba.x, ba.y, y[0], y[1]);
    add2(x[2], x[3], ba.z, ba.w, y[2], y[3]);
    add2(x[4], x[5], bb.x, bb.y, y[4], y[5]);
    add2(x[6], x[7], bb.z, bb.w, y[6], y[7]);
    if (SIDE) {
      const float4 w0 = ld_shared_v4f(wa_addr + 32u * c), w1 = ld_shared_v4f(wa_addr + 32u * c + 16u);
      side = fmaf(fmaxf(y[0], 0.f), w0.x, side); side = fmaf(fmaxf(y[1], 0.f), w0.y, side);
      side = fmaf(fmaxf(y[2], 0.f), w0.z, side); side = fmaf(fmaxf(y[3], 0.f), w0.w, side);
      side = fmaf(fmaxf(y[4], 0.f), w1.x, side); side = fmaf(fmaxf(y[5], 0.f), w1.y, side);
      side = fmaf(fmaxf(y[6], 0.f), w1.z, side); side = fmaf(fmaxf(y[7], 0.f), w1.w, side);
    }
    uint32_t w[4];
#pragma unroll
    for (int u = 0; u < 4; ++u)
      w[u] = (ACT == 0) ? pack_h2_relu(y[2 * u], y[2 * u + 1]) : (ACT == 1) ? pack_h2(elu_f32(y[2 * u]), elu_f32(y[2 * u + 1])) : pack_h2(y[2 * u], y[2 * u + 1]);
    st_shared_v4(row_base + ((uint32_t)(c << 4) ^ xr), w[0], w[1], w[2], w[3]);
  }
  return side;
}

// debug timeline slots (leader CTA of cluster 0, its second iteration): index = base + phase * 2 + slot
constexpr int TL_MMA0 = 0;       // MMA thread: operand-ready wait satisfied
constexpr int TL_MMA1 = 20;      // MMA thread: last MMA of the phase issued
constexpr int TL_ACC = 40;       // epilogue warp 0: accumulator-full observed
constexpr int TL_ARR = 60;       // epilogue warp 0: arrived on operand-ready
constexpr int TL_ARRL = 80;      // epilogue warp 7: arrived on operand-ready
constexpr int TL_EPI = 100;      // epilogue warp 0, phase 2 slot 0: [0] first 64 columns stored, [1] second 64 columns stored
constexpr int TL_FACC = 150;     // follower CTA, epilogue warp 0: accumulator-full observed
constexpr int TL_FARR = 170;     // follower CTA, epilogue warp 0: arrived on operand-ready
constexpr int TL_SYNC = 140;    // clock64() right after the setup cluster barrier: [0] leader, [1] follower (per-SM clock offset)
constexpr int TL_OUT = 110;     // epilogue warp 0, output phase, slot t: [6t + 0] body entry, [1] before the accumulator wait, [2] accumulator in registers, [3] next operand stored, [4] outputs stored
constexpr int TL_CLK = 200;     // CTA 0, thread 0: [0] clock64 / [1] %globaltimer (ns) after setup, [2] / [3] the same before teardown -> effective SM clock
constexpr int TL_N = 208;
__device__ __forceinline__ void tl_mark(long long* tl, bool on, int slot) {
  if (kTimeline && tl && on) tl[slot] = clock64();
}

// ACT: 0 ReLU / 1 ELU.  MODE: InputMode.
template <int ACT, int MODE>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(n_threads(ACT), 1) mlp_tc_kernel(const __grid_constant__ Params p) {
  constexpr bool kHandover = use_handover(ACT);
  constexpr int W_PRODUCER = w_producer(ACT), W_MMA = W_PRODUCER + 1;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  const uint32_t base = (raw_addr + 1023u) & ~1023u;              // identical in both CTAs of the pair
  uint8_t* sm = smem_raw + (base - raw_addr);
  float* s_bias = reinterpret_cast<float*>(sm + OFF_BIAS);
  const uint32_t bar0 = base + OFF_BAR;
  auto bar_full = [&](int s) { return bar0 + 8u * s; };
  auto bar_empty = [&](int s) { return bar0 + 8u * (N_RING + s); };
  auto bar_aready = [&](int t, int h) { return bar0 + 8u * (2 * N_RING + 2 * t + h); };   // h = 0: K blocks {0,1} + accumulator drained; 1: K blocks {2,3}
  auto bar_accfull = [&](int t) { return bar0 + 8u * (2 * N_RING + 4 + t); };      // a hidden layer's accumulator is complete
  auto bar_outfull = [&](int t) { return bar0 + 8u * (2 * N_RING + 6 + t); };      // the output layer's accumulator is complete
  auto bar_infull = [&](int t) { return bar0 + 8u * (2 * N_RING + 8 + t); };       // IN_LOAD16: the next tile's rows have landed (TMA)
  volatile uint32_t* s_tmem = reinterpret_cast<volatile uint32_t*>(sm + OFF_TMEM);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const long long n_clusters = gridDim.x >> 1, cluster_id = blockIdx.x >> 1;
  const long long n_pairs = (p.M + PAIR_M - 1) / PAIR_M;          // 256-row MMA tiles
  // iteration `it` of this cluster covers pair tiles T0 = (it * n_clusters + cluster_id) * 2 (slot 0) and T0 + 1 (slot 1)

  // ---- one-time setup ----
  if (warp == W_MMA && lane == 0) {
    for (int s = 0; s < N_RING; ++s) { mbar_init(bar_full(s), rank == 0 ? 2 : 1); mbar_init(bar_empty(s), 1); }
    for (int t = 0; t < 2; ++t) {
      // operand-ready: 16 arrivals per phase -- one per hidden-epilogue warp of the pair, or two per hand-over warp
      mbar_init(bar_aready(t, 0), 2 * N_EPI_WARPS); mbar_init(bar_aready(t, 1), 2 * N_EPI_WARPS);
      mbar_init(bar_accfull(t), 1); mbar_init(bar_outfull(t), 1); mbar_init(bar_infull(t), 1);
    }
    fence_barrier_init();
  }
  if (warp == W_PRODUCER) tmem_alloc(smem_u32((const void*)s_tmem), TMEM_COLS);
  if (warp < N_EPI_WARPS) {
    for (int i = threadIdx.x; i < p.n_bias_rows * kHidden; i += N_EPI_WARPS * 32) s_bias[i] = p.bias[i];
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *s_tmem;
  if (threadIdx.x == 0) pdl_launch();                      // the successor's CTAs may take over SMs as this grid's CTAs exit
  if (p.clk && blockIdx.x == 0 && threadIdx.x == 0) {           // SM cycles against wall time: the clock this launch really runs at
    unsigned long long ns;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(ns));
    p.clk[0] = clock64(); p.clk[1] = (long long)ns;
  }
  if (kTimeline && p.timeline && blockIdx.x < 2 && threadIdx.x == 0) p.timeline[TL_SYNC + blockIdx.x] = clock64();
  if (kTimeline && p.timeline && blockIdx.x == 0 && threadIdx.x == 0) {
    unsigned long long ns;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(ns));
    p.timeline[TL_CLK + 0] = clock64(); p.timeline[TL_CLK + 1] = (long long)ns;
  }

  // The single-thread roles walk the same (slot, phase) sequence: the two slots alternate phase by phase.  (Round 2 also measured
  // slot 0's first layer of its next tile issued BEFORE slot 1's output layer, which the hand-over warps make possible: +-0 in a
  // train of launches, and it breaks the adjacency the weight sharing below needs.)
  struct Cursor {
    long long n_pairs, stride, tile0, tile1;
    int np, ph0, ph1;
    __device__ __forceinline__ long long tile(int t) const { return t ? tile1 : tile0; }
    __device__ __forceinline__ int ph(int t) const { return t ? ph1 : ph0; }
    __device__ __forceinline__ bool live(int t) const { return tile(t) < n_pairs; }
    __device__ __forceinline__ void advance(int t) {
      if (t) { if (++ph1 == np) { ph1 = 0; tile1 += stride; } }
      else   { if (++ph0 == np) { ph0 = 0; tile0 += stride; } }
    }
  };
  Cursor cur;
  cur.n_pairs = n_pairs; cur.stride = 2 * n_clusters; cur.np = p.n_phases;
  cur.ph0 = cur.ph1 = 0;
  cur.tile0 = 2 * cluster_id; cur.tile1 = 2 * cluster_id + 1;
  // Weight sharing: in the plain alternating order slot 1's phase follows slot 0's SAME phase, so the layer's blocks need to come
  // through the ring only once per unit: slot 0 multiplies them without releasing the ring slots, slot 1 multiplies them again
  // and releases.  A layer's blocks (<= 4) fit the 5-slot ring, and the prefetch distance stays what it was (a block's slot is
  // refilled 2.2 K cycles before the next layer needs it).  Halves the L2 -> shared-memory weight traffic (7 GB per 571 536-ray
  // NeRF launch, 4.5 % of that kernel's energy: profiles/r02_handover/sustained_power_nerf_without_weight_stream.txt).
  // shared_step(): called in body(0) BEFORE the cursor advances -- both slots live and at the same phase.
  const bool walk_share = p.share != 0;
  auto shared_step = [&]() { return walk_share && cur.live(1) && cur.ph0 == cur.ph1; };
  const long long tl_tile0 = 2 * cluster_id + cur.stride;           // timeline: the second tile of slot 0 / slot 1
  // run body(0), body(1) alternately until both slots are out of tiles.  ONE copy of the body (the slot index is a
  // run-time value): two inlined copies overflow the instruction cache.
#define PN_WALK(body)                                                        \
  for (;;) {                                                                 \
    bool any = false;                                                        \
    _Pragma("unroll 1")                                                      \
    for (int t_ = 0; t_ < 2; ++t_) {                                         \
      if (!cur.live(t_)) continue;                                           \
      any = true;                                                            \
      body(t_);                                                              \
      cur.advance(t_);                                                       \
    }                                                                        \
    if (!any) break;                                                         \
  }

  // Register hand-over between the warpgroups (one static instruction per warpgroup, warpgroup-aligned): the role warpgroup and
  // the hand-over warps release, the hidden-epilogue warpgroups take (setmaxnreg.inc waits until the registers are free).  Each
  // instruction sits at the top of the code it governs: ptxas budgets a region by the setmaxnreg that DOMINATES it (after a join
  // of differently budgeted paths it assumes the smallest).
  if (kHandover && warp >= W_PRODUCER) asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(kRoleRegs));
  if (warp == W_PRODUCER) {
    // =============================== weight producer (both CTAs) ===============================
    if (lane == 0) {
      uint32_t slot = 0, ring_par = 1;                       // empty barriers: the first pass over the ring is free
      bool shared = false;
      auto body = [&](int t) {
        if (t == 0) shared = shared_step();
        else if (shared) return;                             // slot 1 multiplies the blocks slot 0's step has brought
        const int ph = cur.ph(t);
        const bool merged = p.ph[ph].merged != 0;
        const int nchunks = merged ? 1 : p.ph[ph].nkb;
        const uint32_t bytes = (uint32_t)p.ph[ph].n_pad * 64u * (merged ? (uint32_t)p.ph[ph].nkb : 1u);   // this CTA's share of a chunk
        const uint8_t* src = p.wimg + p.ph[ph].w_off + rank * bytes;
        for (int c = 0; c < nchunks; ++c) {
          mbar_wait(bar_empty(slot), ring_par, p.error_flag, 1);
          mbar_arrive_expect_tx(bar_full(slot), bytes);
          bulk_copy_g2s(base + OFF_RING + slot * RING_SLOT_BYTES, src, bytes, bar_full(slot));
          src += 2 * bytes;
          if (++slot == N_RING) { slot = 0; ring_par ^= 1u; }
        }
      };
      PN_WALK(body)
    }
  } else if (warp == W_MMA) {
    if (lane == 0 && rank != 0) {
      // =============================== follower: weights-landed relay ===============================
      uint32_t slot = 0, ring_par = 0;
      bool shared = false;
      auto body = [&](int t) {
        if (t == 0) shared = shared_step();
        else if (shared) return;
        const int ph = cur.ph(t);
        const int n = p.ph[ph].merged ? 1 : p.ph[ph].nkb;
        for (int i = 0; i < n; ++i) {
          mbar_wait(bar_full(slot), ring_par, p.error_flag, 5);
          mbar_arrive_cluster(bar_full(slot), 0);
          if (++slot == N_RING) { slot = 0; ring_par ^= 1u; }
        }
      };
      PN_WALK(body)
    } else if (rank == 0) {
      // =============================== leader: MMA issuer ===============================
      // The whole warp walks the loop convergently (waits included) and one elected lane issues: with warp-uniform
      // control flow every tcgen05 operand lives in uniform registers.  A single thread retires one dependent
      // instruction every ~5 cycles, so this loop is kept to a few dozen instructions per K block (4 MMAs = 512 cycles).
      uint32_t slot = 0, ring_par = 0;
      uint32_t ar_par = 0;                                   // bit 2t+h = parity to wait for on a_ready[t][h]
      constexpr uint32_t kDescHi = 0x40004040u;              // SBO 1024 B | descriptor version 1 | SWIZZLE_128B
      const uint32_t a_lo0 = (((base + OFF_A) >> 4) & 0x3FFFu) | (1u << 16);
      const uint32_t b_lo0 = (((base + OFF_RING) >> 4) & 0x3FFFu) | (1u << 16);
      // 4 MMAs over one 64-wide K block: a_lo / b_lo are the low descriptor words of the block's first K = 16 step
      auto issue_block = [&](uint32_t d_tmem, uint32_t a_lo, uint32_t b_lo, uint32_t idesc, uint32_t acc_first) {
        const uint64_t a_desc = ((uint64_t)kDescHi << 32) | a_lo;
        const uint64_t b_desc = ((uint64_t)kDescHi << 32) | b_lo;
        umma_f16_pair(d_tmem, a_desc, b_desc, idesc, acc_first);       // K = 16 per instruction: +32 bytes = +2 in descriptor units
        umma_f16_pair(d_tmem, a_desc + 2, b_desc + 2, idesc, 1u);
        umma_f16_pair(d_tmem, a_desc + 4, b_desc + 4, idesc, 1u);
        umma_f16_pair(d_tmem, a_desc + 6, b_desc + 6, idesc, 1u);
      };
      auto ring_next = [&]() { if (++slot == N_RING) { slot = 0; ring_par ^= 1u; } };
      bool shared = false;
      uint32_t shared_slot = 0, shared_par = 0;
      auto body = [&](int t) {
        const int ph = cur.ph(t);
        // weight sharing: slot 0's step leaves the layer's blocks in the ring, slot 1's step revisits them and releases
        bool do_wait = true, do_release = true;
        if (t == 0) {
          shared = shared_step();
          if (shared) { do_release = false; shared_slot = slot; shared_par = ring_par; }
        } else if (shared) {
          do_wait = false; slot = shared_slot; ring_par = shared_par;
        }
        const bool tl_on = kTimeline && blockIdx.x == 0 && cur.tile(t) == tl_tile0 + t && lane == 0;
        const int nkb = p.ph[ph].nkb, k16_last = p.ph[ph].k16_last;
        const uint32_t acc0 = (uint32_t)p.ph[ph].acc;
        const uint32_t idesc = umma_idesc(p.ph[ph].n_pad);
        const bool merged = p.ph[ph].merged != 0;
        const uint32_t par0 = (ar_par >> (2 * t)) & 1u, par1 = (ar_par >> (2 * t + 1)) & 1u;
        ar_par ^= 3u << (2 * t);
        const uint32_t d_tmem = tmem_base + (uint32_t)t * kHidden;
        const uint32_t bar_done = (kHandover && p.ph[ph].epi == EPI_OUT) ? bar_outfull(t) : bar_accfull(t);   // who drains this accumulator
        const uint32_t a_lo_t = a_lo0 + (uint32_t)t * (A_SLOT_BYTES >> 4);
        constexpr uint32_t kBlk = A_BLOCK_BYTES >> 4, kRing = RING_SLOT_BYTES >> 4;
        // operand halves: [0] = K blocks {0,1} written and the accumulator drained, [1] = K blocks {2,3} written
        mbar_wait(bar_aready(t, 0), par0, p.error_flag, 2);
        tc_fence_after();
        tl_mark(p.timeline, tl_on, TL_MMA0 + ph * 2 + t);
        if (nkb == 4 && k16_last == 4 && !merged) {
          // the standard full-width layer: straight-line issue, 2 K blocks per operand half
          if (do_wait) mbar_wait(bar_full(slot), ring_par, p.error_flag, 3);
          tc_fence_after();
          if (elect_one()) { issue_block(d_tmem, a_lo_t, b_lo0 + slot * kRing, idesc, acc0); if (do_release) umma_commit_pair(bar_empty(slot)); }
          __syncwarp();
          ring_next();
          if (do_wait) mbar_wait(bar_full(slot), ring_par, p.error_flag, 3);
          tc_fence_after();
          if (elect_one()) { issue_block(d_tmem, a_lo_t + kBlk, b_lo0 + slot * kRing, idesc, 1u); if (do_release) umma_commit_pair(bar_empty(slot)); }
          __syncwarp();
          ring_next();
          mbar_wait(bar_aready(t, 1), par1, p.error_flag, 2);
          if (do_wait) mbar_wait(bar_full(slot), ring_par, p.error_flag, 3);
          tc_fence_after();
          if (elect_one()) { issue_block(d_tmem, a_lo_t + 2 * kBlk, b_lo0 + slot * kRing, idesc, 1u); if (do_release) umma_commit_pair(bar_empty(slot)); }
          __syncwarp();
          ring_next();
          if (do_wait) mbar_wait(bar_full(slot), ring_par, p.error_flag, 3);
          tc_fence_after();
          if (elect_one()) {
            issue_block(d_tmem, a_lo_t + 3 * kBlk, b_lo0 + slot * kRing, idesc, 1u);
            if (do_release) umma_commit_pair(bar_empty(slot));
            umma_commit_pair(bar_done);
          }
          __syncwarp();
          ring_next();
        } else {
          // first layer (1..4 K blocks, last one possibly short) and the narrow output layer (one merged ring chunk)
          mbar_wait(bar_aready(t, 1), par1, p.error_flag, 2);
          const uint32_t half16 = (uint32_t)p.ph[ph].n_pad * 4u;      // one K block of this CTA's weight half, in 16-byte units
          uint32_t a_lo = a_lo_t;
          if (merged) {
            if (do_wait) mbar_wait(bar_full(slot), ring_par, p.error_flag, 3);
            tc_fence_after();
            if (elect_one()) {
              uint32_t b_lo = b_lo0 + slot * kRing;
              for (int kb = 0; kb < nkb; ++kb) {
                issue_block(d_tmem, a_lo, b_lo, idesc, acc0 | (uint32_t)kb);
                a_lo += kBlk;
                b_lo += half16;
              }
              if (do_release) umma_commit_pair(bar_empty(slot));
            }
            __syncwarp();
            ring_next();
          } else {
            for (int kb = 0; kb < nkb; ++kb) {
              if (do_wait) mbar_wait(bar_full(slot), ring_par, p.error_flag, 3);
              tc_fence_after();
              if (elect_one()) {
                const uint32_t b_lo = b_lo0 + slot * kRing;
                if (kb + 1 < nkb || k16_last == 4) {
                  issue_block(d_tmem, a_lo, b_lo, idesc, acc0 | (uint32_t)kb);
                } else {
                  const uint64_t a_desc = ((uint64_t)kDescHi << 32) | a_lo, b_desc = ((uint64_t)kDescHi << 32) | b_lo;
                  for (int s2 = 0; s2 < k16_last; ++s2) umma_f16_pair(d_tmem, a_desc + 2u * s2, b_desc + 2u * s2, idesc, acc0 | (uint32_t)(kb | s2));
                }
                if (do_release) umma_commit_pair(bar_empty(slot));           // ring slot is free (in both CTAs) once these MMAs have read it
              }
              __syncwarp();
              a_lo += kBlk;
              ring_next();
            }
          }
          if (elect_one()) umma_commit_pair(bar_done);
          __syncwarp();
        }
        tl_mark(p.timeline, tl_on, TL_MMA1 + ph * 2 + t);
      };
      PN_WALK(body)
    }
  } else if (warp < W_PRODUCER) {
    // =============================== hidden-epilogue warps (0-7) and hand-over warps (8-11), both CTAs ===============================
    const bool is_out = kHandover && warp >= W_OUT0;
    const int ew = warp;
    const int ch = (ew >> 2) & 1;                          // hidden epilogues: this warp drains columns [64ch,+64) and [128+64ch,+64)
    const int q = warp & 3;                                // TMEM lane quadrant this warp may access
    const int r = q * 32 + lane;                           // row of the tile owned by this thread
    const uint32_t a_base = base + OFF_A;
    const uint32_t row_off = (uint32_t)((r >> 3) * 1024 + (r & 7) * 128), xr = (uint32_t)(r & 7) << 4;
    constexpr bool kClassic = (MODE == IN_CLASSIC);          // classic NeRF: extra phase kinds, see DESIGN.md 4.3
    constexpr bool kCompute = (MODE == IN_ENCODE || MODE == IN_PLUECKER || kClassic);
    constexpr bool kNerf = (MODE == IN_ENCODE || MODE == IN_LOAD2);
    constexpr int kXin = (MODE == IN_ENCODE || kClassic) ? 3 : 6;
    const int kb_first = p.ph[0].nkb;                      // K blocks of the first phase (<= 4)
    const int np = p.n_phases;
    const long long stride = cur.stride;

    // global row of this thread in pair tile `tile`
    auto row_of = [&](long long tile) { return tile * PAIR_M + (long long)rank * TILE_M + r; };

    // --- first-layer operand, "compute" modes: 64 operand elements per row as two halves of 32 (16 packed words) ---
    auto fetch_input = [&](long long row, float* xin) {
#pragma unroll
      for (int i = 0; i < kXin; ++i) xin[i] = 0.f;
      if (row < p.M) {
        const float* src = p.in0 + row * ((MODE == IN_ENCODE || kClassic) ? 3 : p.in_stride);
#pragma unroll
        for (int i = 0; i < kXin; ++i) xin[i] = __ldg(src + i);
      }
    };
    auto precompute_half = [&](const float* xin, bool row_live, uint32_t* pre, int half) {
      if (MODE == IN_ENCODE || kClassic) {
        if (half == 0) encode32<0>(xin, pre);
        else encode32<1>(xin, pre);
      } else if (MODE == IN_PLUECKER) {
        // the 6 Pluecker features of the ray; the sampler's P replicated copies are folded into the weights
        // (W_eff = sum over copies, tc_load_net), so the operand is 6 wide: one K = 16 step, all of it in half 0
#pragma unroll
        for (int i = 0; i < 16; ++i) pre[i] = 0u;
        if (half == 0 && row_live) {
          float f6[6];
          pluecker6(xin[0], xin[1], xin[2], xin[3], xin[4], xin[5], f6);
          pre[0] = pack_h2(f6[0], f6[1]); pre[1] = pack_h2(f6[2], f6[3]); pre[2] = pack_h2(f6[4], f6[5]);
        }
      }
    };
    auto store_half = [&](int t, const uint32_t* pre, int half) {
      const uint32_t dst = a_base + t * A_SLOT_BYTES + row_off;      // block 0
      if (MODE == IN_PLUECKER) {
        if (half == 0) {
          st_shared_v4(dst + (0u ^ xr), pre[0], pre[1], pre[2], pre[3]);
          st_shared_v4(dst + (16u ^ xr), 0u, 0u, 0u, 0u);
        }
      } else {
#pragma unroll
        for (int c = 0; c < 4; ++c)
          st_shared_v4(dst + ((uint32_t)((4 * half + c) << 4) ^ xr), pre[4 * c], pre[4 * c + 1], pre[4 * c + 2], pre[4 * c + 3]);
      }
    };
    // --- first-layer operand, "load" modes (fp32 rows): the calling warp loads rows trow0 .. trow0 + 15 of K blocks [kb_lo, kb_lo+nblk) ---
    auto load_input = [&](long long tile, int t, int kb_lo, int nblk, int trow0) {
      const long long row_base = tile * PAIR_M + (long long)rank * TILE_M;
      const int k0 = p.k0;
      for (int kb = 0; kb < nblk; ++kb) {
        const uint32_t dst = a_base + t * A_SLOT_BYTES + kb * A_BLOCK_BYTES;
        const int k = (kb_lo + kb) * 64 + 2 * lane;
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const int trow = trow0 + i;
          const long long grow = row_base + trow;
          float v0 = 0.f, v1 = 0.f;
          if (grow < p.M) {
            const float* src = p.in0 + grow * p.in_stride;
            if (k < k0) v0 = __ldg(src + k);
            if (k + 1 < k0) v1 = __ldg(src + k + 1);
          }
          st_shared_b32(dst + a_chunk_off(trow, lane >> 2) + (lane & 3) * 4, pack_h2(v0, v1));
        }
      }
    };
    // --- first-layer operand, fp16 "load" mode: the tile's rows are one contiguous range of 16-byte chunks (8 halves);
    //     chunk g of the range -> (row, chunk-in-row) -> the swizzled operand position.  `tid` of `nt` threads share a range. ---
    const int cpr = p.k0 >> 3;                             // real chunks per row
    struct Range16 { int c_lo, w, total; uint32_t magic; };   // g / w == (g * magic) >> 20 for g < 128 * w, 2 <= w <= 32
    auto range16 = [&](int kb_lo, int nblk) {
      Range16 R;
      R.c_lo = kb_lo * 8;
      int c_hi = (kb_lo + nblk) * 8;
      if (c_hi >= cpr) c_hi = (cpr + 1) & ~1;              // last block: pad to a whole K = 16 step with zero chunks
      R.w = c_hi - R.c_lo;
      R.total = TILE_M * R.w;
      R.magic = ((1u << 20) + (uint32_t)R.w - 1u) / (uint32_t)R.w;
      return R;
    };
    // a tile's rows: first chunk in global memory and how many of its 128 rows exist (32-bit math per chunk from here on)
    struct Tile16 { const uint4* src; int rows; };
    auto tile16 = [&](long long tile) {
      Tile16 T;
      const long long row0 = tile * PAIR_M + (long long)rank * TILE_M;
      const long long left = p.M - row0;
      T.rows = left >= TILE_M ? TILE_M : (left > 0 ? (int)left : 0);
      T.src = reinterpret_cast<const uint4*>(p.in0) + row0 * cpr;
      return T;
    };
    auto fetch16 = [&](const Tile16& T, const Range16& R, int g, uint4& v) {
      v = make_uint4(0u, 0u, 0u, 0u);
      if (g < R.total) {
        const int row = (int)(((uint32_t)g * R.magic) >> 20), c = R.c_lo + (g - row * R.w);
        if (row < T.rows && c < cpr) v = __ldg(T.src + (row * cpr + c));
      }
    };
    auto store16 = [&](int t, const Range16& R, int g, const uint4& v) {
      if (g < R.total) {
        const int row = (int)(((uint32_t)g * R.magic) >> 20), cc = g - row * R.w;       // chunk within the range: block cc >> 3, chunk cc & 7
        st_shared_v4(a_base + t * A_SLOT_BYTES + (cc >> 3) * A_BLOCK_BYTES + a_chunk_off(row, cc & 7), v.x, v.y, v.z, v.w);
      }
    };
    // synchronous form (prologue, and the second part of a first layer wider than 256): load 4 chunks, store 4 chunks
    auto load_input16 = [&](long long tile, int t, int kb_lo, int nblk, int tid, int nt) {
      const Range16 R = range16(kb_lo, nblk);
      const Tile16 T = tile16(tile);
      for (int g0 = 0; g0 < R.total; g0 += 4 * nt) {
        uint4 v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) fetch16(T, R, g0 + tid + u * nt, v[u]);
#pragma unroll
        for (int u = 0; u < 4; ++u) store16(t, R, g0 + tid + u * nt, v[u]);
      }
    };
    // publish this warp's share of slot t's operand: half 0 (its first K block; also "my accumulator reads are done"),
    // half 1 (its second K block), or both at once (first-layer operands).  narr arrivals per warp (the barriers count 16
    // per phase: 8 hidden-epilogue warps x 2 CTAs x 1, or 4 hand-over warps x 2 CTAs x 2).
    auto publish = [&](int t, int halves, int narr) {
      fence_proxy_async();
      tc_fence_before();
      __syncwarp();
      if (lane < narr) {
        if (halves & 1) mbar_arrive_cluster(bar_aready(t, 0), 0);
        if (halves & 2) mbar_arrive_cluster(bar_aready(t, 1), 0);
      }
    };
    auto fetch_dterm = [&](long long row) {
      float4 d = make_float4(0.f, 0.f, 0.f, 0.f);
      if (row < p.M) {
        const uint32_t idx = (uint32_t)row / (uint32_t)p.dir_div;       // M < 2^32 (checked on the host): a 64-bit division is a real call
        d = __ldg(reinterpret_cast<const float4*>(p.dirterm) + idx);
      }
      return d;
    };

    // Everything above -- barrier init, TMEM allocation, the bias table -- and the weight producer's first blocks touch only
    // this network's static weights: with programmatic dependent launch they run while the previous kernel of the stream drains.
    // The inputs (and every output buffer) belong to the dependency chain: wait for the predecessor's completion here.
    pdl_wait();

    if (kHandover && is_out) {
      asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(kOutRegs));
      // =============================== hand-over warps: output layer + next tile's first operand ===============================
      // Per slot and tile: fetch the next tile's raw inputs and turn them into the packed first-layer operand IN REGISTERS (or
      // prefetch its rows towards L2), then wait for "output full".  From there the slot's critical path is: operand stores (or
      // three TMA boxes) -> tcgen05.ld of the output columns -> operand landed -> proxy fence -> arrive; the activations of the
      // head outputs and every global store come after BOTH slots have been handed back to the tensor pipe.
      const int htid = (int)threadIdx.x - W_OUT0 * 32;       // 0..127
      uint32_t out_par = 0, in_par = 0;
      // prologue: first operands of both slots.  The kernel starts cold (the rows come from HBM, ~1.5 us per dependent round trip), so
      // both slots' loads are in flight before either is waited for: at 8 GPUs a rank's whole kernel is two iterations, ~70 us.
      if (MODE == IN_LOAD16) {
        if (htid == 0) {
#pragma unroll 1
          for (int t = 0; t < 2; ++t) {
            if (!cur.live(t)) continue;
            const int row0 = (int)(row_of(cur.tile(t)) - r);
            mbar_arrive_expect_tx(bar_infull(t), (uint32_t)kb_first * A_BLOCK_BYTES);
            for (int kb = 0; kb < kb_first; ++kb)
              tma_load_box(a_base + t * A_SLOT_BYTES + kb * A_BLOCK_BYTES, &p.tmap_in, kb * 64, row0, bar_infull(t));
          }
        }
#pragma unroll 1
        for (int t = 0; t < 2; ++t) {
          if (!cur.live(t)) continue;
          mbar_wait_polite(bar_infull(t), 0u, p.error_flag, 7);
          in_par ^= 1u << t;
          publish(t, 3, 2);
        }
      } else if (kCompute) {
        float xa[kXin], xb[kXin];
        const long long rowa = row_of(cur.tile(0)), rowb = row_of(cur.tile(1));
        fetch_input(rowa, xa);
        fetch_input(cur.live(1) ? rowb : p.M, xb);
#pragma unroll
        for (int t = 0; t < 2; ++t) {
          if (!cur.live(t)) continue;
          uint32_t pre[32];
          const bool row_live = (t ? rowb : rowa) < p.M;
          precompute_half(t ? xb : xa, row_live, pre, 0);
          if (MODE != IN_PLUECKER) precompute_half(t ? xb : xa, row_live, pre + 16, 1);
          store_half(t, pre, 0);
          if (MODE != IN_PLUECKER) store_half(t, pre + 16, 1);
          publish(t, 3, 2);
        }
      } else {
#pragma unroll 1
        for (int t = 0; t < 2; ++t) {
          if (!cur.live(t)) continue;
          load_input(cur.tile(t), t, 0, kb_first, q * 32);
          load_input(cur.tile(t), t, 0, kb_first, q * 32 + 16);
          publish(t, 3, 2);
        }
      }
      const int layer_out = p.ph[np - 1].layer, n_pad_out = p.ph[np - 1].n_pad;
      const int n_out = p.n_out;
      const uint32_t tmem_q = tmem_base + ((uint32_t)(q * 32) << 16);
      float* stg = reinterpret_cast<float*>(sm + OFF_STAGE) + q * STAGE_FLOATS;

      // ---- the slot's critical path: "output full" -> next operand on its way -> accumulator columns [c0, c0 + 48) in registers ->
      //      (last chunk) operand landed -> publish.  `pre` = the next tile's packed first-layer operand (compute modes). ----
      auto wait_outfull = [&](int t) {
        mbar_wait_polite(bar_outfull(t), (out_par >> t) & 1u, p.error_flag, 6);
        out_par ^= 1u << t;
        tc_fence_after();
      };
      auto start_next_operand = [&](long long tile, int t, const uint32_t* pre) {
        // the slot's operand buffer is free (the output layer has read it)
        if (kCompute) {
          store_half(t, pre, 0);
          if (MODE != IN_PLUECKER) store_half(t, pre + 16, 1);
        } else if (MODE == IN_LOAD16) {
          // three instructions instead of 2 304 16-byte copies with their address arithmetic (2 K cycles on the slot's critical path)
          if (htid == 0) {
            const int row0 = (int)(row_of(tile + stride) - r);
            mbar_arrive_expect_tx(bar_infull(t), (uint32_t)kb_first * A_BLOCK_BYTES);
            for (int kb = 0; kb < kb_first; ++kb)
              tma_load_box(a_base + t * A_SLOT_BYTES + kb * A_BLOCK_BYTES, &p.tmap_in, kb * 64, row0, bar_infull(t));
          }
        } else {
          load_input(tile + stride, t, 0, kb_first, q * 32);
          load_input(tile + stride, t, 0, kb_first, q * 32 + 16);
        }
      };
      auto hand_back = [&](int t, bool has_next) {
        if (MODE == IN_LOAD16 && has_next) {                       // the next tile's rows have landed
          mbar_wait_polite(bar_infull(t), (in_par >> t) & 1u, p.error_flag, 7);
          in_par ^= 1u << t;
        }
        publish(t, 3, 2);
      };
      // what the critical path prepares before "output full": the next tile's operand in registers, an L2 prefetch of its rows
      auto prepare_next = [&](long long tile, bool has_next, uint32_t* pre) {
        if (kCompute && has_next) {
          float xin[kXin];
          const long long nrow = row_of(tile + stride);
          fetch_input(nrow, xin);
          precompute_half(xin, nrow < p.M, pre, 0);
          if (MODE != IN_PLUECKER) precompute_half(xin, nrow < p.M, pre + 16, 1);
        }
        if (MODE == IN_LOAD16 && htid == 0 && has_next) {
          // pull the next tile's rows (one contiguous range) towards L2
          const long long row0 = row_of(tile + stride) - r;
          long long nrow = p.M - row0;
          if (nrow > TILE_M) nrow = TILE_M;
          if (nrow > 0) {
            const uint32_t bytes = (uint32_t)(nrow * (long long)p.k0 * 2);
            asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(reinterpret_cast<const __half*>(p.in0) + row0 * p.k0), "r"(bytes) : "memory");
          }
        }
      };
      // ---- off the critical path: head activations of output columns [c0, c0 + 48) and the stores.  A warp's 32 rows x n_out
      //      floats are ONE contiguous global range, but thread = row: direct stores would scatter 4 bytes over 32 sectors per
      //      instruction.  Activation in registers (branch-free coefficient form), then rows_per_pass rows at a time through the
      //      lane quadrant's staging window and out as coalesced stores. ----
      auto emit_heads = [&](float* v, int c0, long long row) {
        const float* bo = s_bias + layer_out * kHidden + c0;
        // blocks of 8 columns: a block of padding / linear outputs (p.head_linear) skips the SFU form -- the SFU is the unit
        // the hidden epilogues of these kernels saturate, and the sampler's direction outputs (16 of 27 columns) are linear
        auto activate = [&](auto c0_tag) {
          constexpr int C0 = decltype(c0_tag)::value;
#pragma unroll
          for (int g = 0; g < 6; ++g) {
            if (C0 + 8 * g >= n_out) continue;
            if ((p.head_linear >> (C0 / 8 + g)) & 1u) {
#pragma unroll
              for (int i = 0; i < 8; ++i) v[8 * g + i] += bo[8 * g + i];
            } else {
#pragma unroll
              for (int i = 0; i < 8; ++i) v[8 * g + i] = head_apply_tab(v[8 * g + i] + bo[8 * g + i], p.head_tab[C0 + 8 * g + i]);
            }
          }
        };
        if (c0 == 0) activate(std::integral_constant<int, 0>());
        else activate(std::integral_constant<int, 48>());
        int w = n_out - c0;                                  // real columns of this chunk
        if (w > 48) w = 48;
        if (w <= 0) return;
        int rpp = (STAGE_FLOATS / w) & ~3;                   // rows per staging pass (a multiple of 4: 16-byte aligned passes)
        if (rpp > 32) rpp = 32;
        const bool whole_rows = w == n_out;                  // the usual case: the staged rows are one contiguous global range
        const long long grow_w = row - lane;                 // global row of the quadrant's first thread (a multiple of 32)
#pragma unroll 1
        for (int r0 = 0; r0 < 32; r0 += rpp) {
          const int l = lane - r0;
          if (l >= 0 && l < rpp) {
            float* d = stg + l * w;
#pragma unroll
            for (int o = 0; o < 48; ++o)
              if (o < w) d[o] = v[o];
          }
          __syncwarp();
          const long long nl = p.M - (grow_w + r0);
          int nrows = 32 - r0 < rpp ? 32 - r0 : rpp;
          if (nl < nrows) nrows = nl > 0 ? (int)nl : 0;
          const int nfl = nrows * w;
          float* dst = p.out + (grow_w + r0) * n_out;
          if (whole_rows && (reinterpret_cast<uintptr_t>(dst) & 15) == 0) {
            for (int i = lane * 4; i < nfl; i += 128) {
              if (i + 4 <= nfl) {
                *reinterpret_cast<float4*>(dst + i) = *reinterpret_cast<const float4*>(stg + i);
              } else {
                for (int k = i; k < nfl; ++k) dst[k] = stg[k];
              }
            }
          } else if (whole_rows) {
            for (int i = lane; i < nfl; i += 32) dst[i] = stg[i];
          } else {
            for (int i = lane; i < nfl; i += 32) {
              const int rr = i / w, cc = i - rr * w;
              dst[(long long)rr * n_out + c0 + cc] = stg[i];
            }
          }
          __syncwarp();
        }
      };
      auto ld48 = [&](uint32_t taddr, int c0, float* v) {
        tmem_ld16(taddr + c0, v);
        if (n_pad_out > c0 + 16) tmem_ld16(taddr + c0 + 16, v + 16);
        if (n_pad_out > c0 + 32) tmem_ld16(taddr + c0 + 32, v + 32);
        tmem_wait_ld();
      };
      // One shape for any number of head columns: first the critical paths of both slots (each slot's accumulator chunks are
      // drained into a per-thread scratch array -- local memory on purpose: two slots x up to two 48-column chunks do not fit the
      // hand-over warps' registers next to the activation temporaries, and an indexed array keeps ONE copy of the output code), then
      // the head activations and stores of everything.  Slot 1's "output full" arrives one hidden epilogue after slot 0's, and
      // behind slot 0's head activations (which queue at the SFU the hidden epilogues saturate: ~5 K cycles) its hand-over would
      // come 2-3 K cycles late, with the hidden-epilogue warps waiting.
      constexpr int kPre = (kCompute && MODE != IN_PLUECKER) ? 32 : 16;
      float vbuf[2][2][48];
      int it_idx = 0;
      for (long long T0 = 2 * cluster_id; T0 < n_pairs; T0 += stride, ++it_idx) {
        const int nslots = (T0 + 1 < n_pairs) ? 2 : 1;
        const bool tl_o = kTimeline && blockIdx.x == 0 && it_idx == 1 && warp == W_OUT0 && lane == 0;
#pragma unroll 1
        for (int t = 0; t < nslots; ++t) {
          const long long tile = T0 + t;
          const bool has_next = tile + stride < n_pairs;
          tl_mark(p.timeline, tl_o, TL_OUT + t * 6 + 0);
          uint32_t pre[kPre];
          prepare_next(tile, has_next, pre);
          tl_mark(p.timeline, tl_o, TL_OUT + t * 6 + 1);
          wait_outfull(t);
          if (has_next) start_next_operand(tile, t, pre);
#pragma unroll 1
          for (int c0 = 0, ci = 0; c0 < n_pad_out; c0 += 48, ++ci) {
            float v[48];
            ld48(tmem_q + (uint32_t)t * kHidden, c0, v);
            if (c0 + 48 >= n_pad_out) {                            // the slot's last chunk is in registers: hand the slot back
              tl_mark(p.timeline, tl_o, TL_OUT + t * 6 + 2);
              hand_back(t, has_next);
              tl_mark(p.timeline, tl_o, TL_OUT + t * 6 + 3);
            }
#pragma unroll
            for (int o = 0; o < 48; ++o)
              if (o < 16 || n_pad_out > c0 + (o & ~15)) vbuf[t][ci][o] = v[o];
          }
        }
#pragma unroll 1
        for (int t = 0; t < nslots; ++t) {
#pragma unroll 1
          for (int c0 = 0, ci = 0; c0 < n_pad_out; c0 += 48, ++ci) {
            float v[48];
#pragma unroll
            for (int o = 0; o < 48; ++o) v[o] = (o < 16 || n_pad_out > c0 + (o & ~15)) ? vbuf[t][ci][o] : 0.f;
            emit_heads(v, c0, row_of(T0 + t));
          }
          tl_mark(p.timeline, tl_o, TL_OUT + t * 6 + 4);
        }
      }
    } else {
    if (kHandover) asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(kEpiRegs));
    // =============================== hidden-epilogue warps ===============================
    // These are the busiest warps of the kernel and a lone warp retires one DEPENDENT instruction every ~5 cycles, so their
    // control flow is spelled out as plain nested loops (iteration -> phase -> slot; the two slots run in step) with everything
    // per-phase hoisted, instead of the generic cursor walk of the single-thread roles.
    uint32_t acc_par = 0;
    // ReLU kernels (no hand-over warps): these warps also run the output phase and build the next tile's first operand, warp
    // (q, ch) producing half `ch` of it.  The raw inputs of each slot's NEXT tile are fetched one phase ahead.
    float xin0[kXin], xin1[kXin];
    float4 dterm0 = make_float4(0.f, 0.f, 0.f, 0.f), dterm1 = dterm0;
    if (!kHandover) {
#pragma unroll
      for (int i = 0; i < kXin; ++i) { xin0[i] = 0.f; xin1[i] = 0.f; }
#pragma unroll 1
      for (int t = 0; t < 2; ++t) {
        if (!cur.live(t)) continue;
        if (kCompute) {
          float xin[kXin];
          uint32_t pre[16];
          const long long row = row_of(cur.tile(t));
          fetch_input(row, xin);
          precompute_half(xin, row < p.M, pre, ch);
          store_half(t, pre, ch);
        } else {
          load_input(cur.tile(t), t, 0, kb_first, q * 32 + ch * 16);
        }
        publish(t, 3, 1);
      }
    }
    const bool more = p.ph[0].epi == EPI_MORE;
    const uint32_t bias_base = base + OFF_BIAS + (uint32_t)(ch * 64) * 4u;
    const uint32_t tmem_lane = tmem_base + ((uint32_t)(q * 32) << 16);
    const uint32_t a_row = a_base + row_off;
    int it_idx = 0;
    for (long long T0 = 2 * cluster_id; T0 < n_pairs; T0 += stride, ++it_idx) {
      const int nslots = (T0 + 1 < n_pairs) ? 2 : 1;
      const bool tl_it = kTimeline && blockIdx.x == 0 && it_idx == 1 && lane == 0;
      const bool tl_f = kTimeline && blockIdx.x == 1 && it_idx == 1 && threadIdx.x == 0;
      int ph = 0;
      if (more) {
        // first layer wider than 256: the remaining K blocks replace the ones just consumed
#pragma unroll 1
        for (int t = 0; t < nslots; ++t) {
          mbar_wait(bar_accfull(t), (acc_par >> t) & 1u, p.error_flag, 4);
          acc_par ^= 1u << t;
          tc_fence_after();
          if (MODE == IN_LOAD16) load_input16(T0 + t, t, kb_first, p.ph[1].nkb, (int)threadIdx.x, N_EPI_WARPS * 32);
          else load_input(T0 + t, t, kb_first, p.ph[1].nkb, q * 32 + ch * 16);
          publish(t, 3, 1);
        }
        ph = 1;
      }
#pragma unroll 1
      for (; ph < np - 1; ++ph) {
        const uint32_t bias_addr = bias_base + (uint32_t)p.ph[ph].layer * (uint32_t)(kHidden * 4);
        if (kClassic && p.ph[ph].epi == EPI_MORE) {
          // A layer whose input is a concatenation (helpers.py:833-834, 839): the K = 256 part has just been multiplied;
          // the other part -- the encoded points (skip layer) or the encoded view direction (view layer) -- is written
          // into K block 0 and a short accumulating phase follows.  The raw values are fetched before the accumulator wait.
          const int src = p.ph[ph].more_src;
#pragma unroll 1
          for (int t = 0; t < nslots; ++t) {
            const long long row = row_of(T0 + t);
            float x3[3] = {0.f, 0.f, 0.f};
            if (row < p.M) {
              const float* g = (src == MORE_PTS) ? p.in0 + row * 3
                                                 : p.vdir + (long long)((uint32_t)row / (uint32_t)p.dir_div) * p.vdir_stride;
              x3[0] = __ldg(g); x3[1] = __ldg(g + 1); x3[2] = __ldg(g + 2);
            }
            mbar_wait(bar_accfull(t), (acc_par >> t) & 1u, p.error_flag, 4);
            acc_par ^= 1u << t;
            tc_fence_after();
            uint32_t pre[16];
            if (src == MORE_PTS) {
              precompute_half(x3, row < p.M, pre, ch);
              store_half(t, pre, ch);
            } else if (ch == 0) {
              // gamma_4(viewdir): [v, sin(2^l v), cos(2^l v)]_{l<4}, 27 values + 5 zeros = K block 0, chunks 0..3
              float e[32];
#pragma unroll
              for (int i = 0; i < 32; ++i) e[i] = 0.f;
#pragma unroll
              for (int c = 0; c < 3; ++c) {
                e[c] = x3[c];
#pragma unroll
                for (int l = 0; l < 4; ++l) {
                  float sn, cs;
                  __sincosf(x3[c] * (float)(1 << l), &sn, &cs);
                  e[3 + 6 * l + c] = sn;
                  e[6 + 6 * l + c] = cs;
                }
              }
              const uint32_t dst = a_base + t * A_SLOT_BYTES + row_off;
#pragma unroll
              for (int c = 0; c < 4; ++c)
                st_shared_v4(dst + ((uint32_t)(c << 4) ^ xr), pack_h2(e[8 * c], e[8 * c + 1]), pack_h2(e[8 * c + 2], e[8 * c + 3]),
                             pack_h2(e[8 * c + 4], e[8 * c + 5]), pack_h2(e[8 * c + 6], e[8 * c + 7]));
            }
            publish(t, 3, 1);
          }
          continue;
        }
        const bool act_none = kClassic && p.ph[ph].act_none != 0;
        const bool side = kClassic && p.ph[ph].side != 0;
        const bool narrow = kClassic && p.ph[ph].n_pad <= 128;           // 128-wide layer: only the first column group exists
#pragma unroll 1
        for (int t = 0; t < nslots; ++t) {
          const bool tl_e = tl_it && ew == 0 && ph == 2 && t == 0;      // fine-grained stamps of one hidden epilogue
          tl_mark(p.timeline, tl_e, TL_EPI + 0);
          if (!kHandover && ph == np - 2) {
            // one phase before the output layer: start the global loads the output epilogue will need.  Each slot's
            // registers are written directly (no select on the loaded values: that would wait for the loads here).
            const long long tile = T0 + t;
            if (kCompute && tile + stride < n_pairs) {
              if (t == 0) fetch_input(row_of(tile + stride), xin0);
              else fetch_input(row_of(tile + stride), xin1);
            }
            if (kNerf && ch == 0) {
              if (t == 0) dterm0 = fetch_dterm(row_of(tile));
              else dterm1 = fetch_dterm(row_of(tile));
            }
          }
          tl_mark(p.timeline, tl_e, TL_EPI + 1);
          mbar_wait(bar_accfull(t), (acc_par >> t) & 1u, p.error_flag, 4);
          acc_par ^= 1u << t;
          tc_fence_after();
          tl_mark(p.timeline, tl_it && ew == 0, TL_ACC + ph * 2 + t);
          tl_mark(p.timeline, tl_f, TL_FACC + ph * 2 + t);
          // 128 columns per thread, in two 64-column groups: the second group's tcgen05.ld is in flight while the first
          // group is converted and stored.  This thread's columns: [64 ch, +64) -> K block ch, then [128 + 64 ch, +64)
          // -> K block 2 + ch, so the warps' first groups together complete K blocks {0,1} (operand half 0) and their
          // second groups {2,3} (half 1)
          const uint32_t taddr = tmem_lane + (uint32_t)t * kHidden + ch * 64;
          const uint32_t row_base = a_row + t * A_SLOT_BYTES + ch * A_BLOCK_BYTES;
          float v[128];
          tmem_ld32(taddr, v);
          tmem_ld32(taddr + 32, v + 32);
          tmem_wait_ld();
          tl_mark(p.timeline, tl_e, TL_EPI + 2);
          if (kClassic && (act_none || side || narrow)) {
            // classic NeRF's special layers: linear feature layer, the layer feeding the alpha head, the 128-wide view layer
            if (!narrow) {
              tmem_ld32(taddr + 128, v + 64);
              tmem_ld32(taddr + 128 + 32, v + 96);
            }
            const uint32_t wa = base + OFF_BIAS + (uint32_t)(p.alpha_row * kHidden + ch * 64) * 4u;
            float dot = 0.f;
            if (act_none) epilogue_store64<2>(v, bias_addr, row_base, xr);
            else if (side) dot = epilogue_store64<0, true>(v, bias_addr, row_base, xr, wa);
            else epilogue_store64<0>(v, bias_addr, row_base, xr);
            tmem_wait_ld();
            if (!narrow) {
              if (act_none) epilogue_store64<2>(v + 64, bias_addr + 512u, row_base + 2 * A_BLOCK_BYTES, xr);
              else if (side) dot += epilogue_store64<0, true>(v + 64, bias_addr + 512u, row_base + 2 * A_BLOCK_BYTES, xr, wa + 512u);
              else epilogue_store64<0>(v + 64, bias_addr + 512u, row_base + 2 * A_BLOCK_BYTES, xr);
            }
            if (side)      // this thread's half of alpha_linear(h) for its row; the hand-over warps add the two halves
              s_bias[(p.alpha_row + 1 + t) * kHidden + ch * TILE_M + r] = dot;
            publish(t, 3, 1);
          } else {
          tmem_ld32(taddr + 128, v + 64);
          tmem_ld32(taddr + 128 + 32, v + 96);
          epilogue_store64<ACT>(v, bias_addr, row_base, xr);
          tl_mark(p.timeline, tl_e, TL_EPI + 3);
          tmem_wait_ld();                                      // every accumulator column of this thread is in registers
          tl_mark(p.timeline, tl_e, TL_EPI + 4);
          const bool split = (p.split & 1) || ((p.split & 2) && ph == 0);
          if (split) publish(t, 1, 1);                         // K blocks {0,1} are ready: the next layer's MMAs may start
          epilogue_store64<ACT>(v + 64, bias_addr + 512u, row_base + 2 * A_BLOCK_BYTES, xr);
          tl_mark(p.timeline, tl_e, TL_EPI + 5);
          publish(t, split ? 2 : 3, 1);
          }
          tl_mark(p.timeline, tl_e, TL_EPI + 6);
          tl_mark(p.timeline, tl_it && ew == 0, TL_ARR + ph * 2 + t);
          tl_mark(p.timeline, tl_it && ew == N_EPI_WARPS - 1, TL_ARRL + ph * 2 + t);
          tl_mark(p.timeline, tl_f, TL_FARR + ph * 2 + t);
        }
      }
      if (!kHandover) {
        // ---------------- output phase (ReLU kernels) ----------------
        // Order matters: the accumulator is pulled into registers, the next tile's first-layer operand is written and the
        // slot is PUBLISHED before anything goes to global memory -- the proxy fence in publish() is a CTA-wide memory
        // barrier, and behind a global store it would wait for the store's L2 round trip.
        const int layer_out = p.ph[np - 1].layer;
#pragma unroll 1
        for (int t = 0; t < nslots; ++t) {
          const long long tile = T0 + t;
          const bool has_next = tile + stride < n_pairs;
          const bool tl_o = tl_it && ew == 0;
          tl_mark(p.timeline, tl_o, TL_OUT + t * 6 + 0);
          uint32_t pre[16];
          if (kCompute && has_next) {                            // overlaps the wait below
            float xin[kXin];
#pragma unroll
            for (int i = 0; i < kXin; ++i) xin[i] = t ? xin1[i] : xin0[i];
            precompute_half(xin, row_of(tile + stride) < p.M, pre, ch);
          }
          tl_mark(p.timeline, tl_o, TL_OUT + t * 6 + 1);
          mbar_wait(bar_accfull(t), (acc_par >> t) & 1u, p.error_flag, 4);
          acc_par ^= 1u << t;
          tc_fence_after();
          tl_mark(p.timeline, tl_it && ew == 0, TL_ACC + ph * 2 + t);
          tl_mark(p.timeline, tl_f, TL_FACC + ph * 2 + t);
          const long long row = row_of(tile);
          float v[16];
          if (ch == 0) {                                         // 4 real output columns: the ch = 1 warps have none
            tmem_ld16(tmem_lane + (uint32_t)t * kHidden, v);
            tmem_wait_ld();
          }
          tl_mark(p.timeline, tl_o, TL_OUT + t * 6 + 2);
          if (has_next) {
            if (kCompute) store_half(t, pre, ch);
            else load_input(tile + stride, t, 0, kb_first, q * 32 + ch * 16);
          }
          tl_mark(p.timeline, tl_o, TL_OUT + t * 6 + 3);
          publish(t, 3, 1);
          tl_mark(p.timeline, tl_it && ew == 0, TL_ARR + ph * 2 + t);
          tl_mark(p.timeline, tl_it && ew == N_EPI_WARPS - 1, TL_ARRL + ph * 2 + t);
          tl_mark(p.timeline, tl_f, TL_FARR + ph * 2 + t);
          if (ch == 0 && row < p.M) {
            const float* bo = s_bias + layer_out * kHidden;
            if (kClassic) {
              // [rgb_linear(h), alpha_linear(h7)] (helpers.py:843-844): alpha = the two half dot products of the pts_linears.7 epilogue
              const float* ap = s_bias + (p.alpha_row + 1 + t) * kHidden;
              *reinterpret_cast<float4*>(p.out + row * 4) =
                  make_float4(v[0] + bo[0], v[1] + bo[1], v[2] + bo[2], ap[r] + ap[TILE_M + r] + p.alpha_bias);
            } else {
              // DoNeRFTRT's last layer: hidden part from the tensor cores + W7[:, 256:283] . gamma_4(viewdir) (pre-pass)
              const float4 d = t ? dterm1 : dterm0;
              *reinterpret_cast<float4*>(p.out + row * 4) = make_float4(v[0] + bo[0] + d.x, v[1] + bo[1] + d.y, v[2] + bo[2] + d.z, v[3] + bo[3] + d.w);
            }
          }
          tl_mark(p.timeline, tl_o, TL_OUT + t * 6 + 4);
        }
      }
    }
    }
  }
#undef PN_WALK

  // ---- teardown: nobody may leave while the peer can still touch this CTA's shared memory or barriers ----
  tc_fence_before();
  __syncthreads();
  if (p.clk && blockIdx.x == 0 && threadIdx.x == 0) {
    unsigned long long ns;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(ns));
    p.clk[2] = clock64(); p.clk[3] = (long long)ns;
  }
  if (kTimeline && p.timeline && blockIdx.x == 0 && threadIdx.x == 0) {
    unsigned long long ns;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(ns));
    p.timeline[TL_CLK + 2] = clock64(); p.timeline[TL_CLK + 3] = (long long)ns;
  }
  cluster_sync_all();
  if (warp == W_PRODUCER) tmem_dealloc(tmem_base, TMEM_COLS);
}

// ------------------------------------------------------------------------------------------------ weight packing
// W [out][in] fp32 -> stream of K-block chunks.  A chunk holds n_pad rows x 64 k fp16 as [CTA-0 half][CTA-1 half],
// each half (n_pad/2 rows) in the UMMA K-major 128B-swizzle layout (narrow output layers: one chunk per layer, see below).  fold > 1: input column k stands for the sum of
// columns k, k + fold_stride, ... (the sampler's P replicated Pluecker blocks).
__global__ void pack_tc_kernel(const float* __restrict__ W, int out_dim, int in_dim, int k_used, int fold, int fold_stride, int n_pad,
                               int kblocks, int merged, uint8_t* __restrict__ dst, int k_src0 = 0) {
  const int total = kblocks * n_pad * 64;
  const int half_rows = n_pad / 2;
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
    const int kb = idx / (n_pad * 64);
    const int rem = idx - kb * n_pad * 64;
    const int n = rem >> 6, k = rem & 63;
    const int ks = kb * 64 + k;
    float v = 0.f;
    if (n < out_dim && ks < k_used)
      for (int f = 0; f < fold; ++f) v += W[(size_t)n * in_dim + k_src0 + ks + f * fold_stride];
    const int h = n / half_rows, rr = n - h * half_rows;
    // plain: [kb][rank][rows];  merged (narrow layers, one chunk per layer): [rank][kb][rows]
    const size_t blk = merged ? ((size_t)h * kblocks + kb) : ((size_t)kb * 2 + h);
    const size_t off = blk * half_rows * 128 + (size_t)(rr >> 3) * 1024 + (rr & 7) * 128 + (((k >> 3) ^ (rr & 7)) << 4) + (k & 7) * 2;
    *reinterpret_cast<__half*>(dst + off) = __float2half_rn(v);
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

// View-direction term of DoNeRFTRT's last layer, fp32: out[i][o] = sum_j W7[o][256 + j] * g[j].
// mode 0: g = gamma_4(viewdir i) (helpers.py:666-671; one per ray);  mode 1: g = row i of embedded_dirs [n, 27].
__global__ void dirterm_kernel(const float* __restrict__ in, int stride, int mode, const float* __restrict__ wdir, long long n,
                               float* __restrict__ out) {
  __shared__ float s_w[4 * 27];
  if (threadIdx.x < 4 * 27) s_w[threadIdx.x] = wdir[threadIdx.x];      // static weights: before the dependency wait
  pdl_wait();
  pdl_launch();
  __syncthreads();
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    if (mode == 0) {
      *reinterpret_cast<float4*>(out + i * 4) = dirterm_of(in + i * stride, s_w);
      continue;
    }
    float g[27];
#pragma unroll
    for (int j = 0; j < 27; ++j) g[j] = in[i * stride + j];
    float o[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      float a = 0.f;
#pragma unroll
      for (int j = 0; j < 27; ++j) a = fmaf(s_w[k * 27 + j], g[j], a);
      o[k] = a;
    }
    *reinterpret_cast<float4*>(out + i * 4) = make_float4(o[0], o[1], o[2], o[3]);
  }
}

}  // namespace tc

// ------------------------------------------------------------------------------------------------ host side
static long long* g_tc_timeline = nullptr;
void tc_set_timeline(long long* dev_buf) { g_tc_timeline = dev_buf; }
static long long* g_tc_clock = nullptr;                  // [3 networks][4] int64 (pn_debug_tc_clock)
void tc_set_clock(long long* dev_buf) { g_tc_clock = dev_buf; }

struct TcLayout {
  int kblocks[kMaxLayers];
  int n_pad[kMaxLayers];
  int k_used[kMaxLayers];
  size_t chunk_off[kMaxLayers];
  bool merged[kMaxLayers];       // narrow layer whose K blocks all fit one ring slot: streamed as a single chunk
  bool has_fold;                 // sampler: a second, folded image of layer 0 (6 inputs) for the in-kernel Pluecker operand
  size_t fold_off;
  size_t img_bytes, bias_off, wdir_off, total;
};

static TcLayout tc_layout(int net_id, int n_layers, const int* in_dims, const int* out_dims) {
  TcLayout L{};
  size_t off = 0;
  for (int l = 0; l < n_layers; ++l) {
    const bool last = l == n_layers - 1;
    int k_used = in_dims[l];
    if (last && net_id == PN_NET_NERF) k_used = kHidden;    // the 27 view-direction inputs go through the dirterm pre-pass
    L.k_used[l] = k_used;
    L.kblocks[l] = (k_used + 63) / 64;
    L.n_pad[l] = last ? (out_dims[l] + 15) / 16 * 16 : kHidden;
    L.chunk_off[l] = off;
    L.merged[l] = last && (size_t)L.kblocks[l] * L.n_pad[l] * 64 <= (size_t)tc::RING_SLOT_BYTES;
    off += (size_t)L.kblocks[l] * L.n_pad[l] * 128;
  }
  L.has_fold = net_id == PN_NET_SAMPLER && in_dims[0] % 6 == 0;
  L.fold_off = off;
  if (L.has_fold) off += (size_t)kHidden * 128;
  L.img_bytes = off;
  L.bias_off = (off + 255) & ~(size_t)255;
  L.wdir_off = L.bias_off + (size_t)n_layers * kHidden * 4;
  L.total = L.wdir_off + 4 * 28 * 4;
  return L;
}

const float* tc_wdir(const NetTC& n) {
  if (!n.loaded || n.classic || n.net_id != PN_NET_NERF || !n.blob) return nullptr;
  const TcLayout L = tc_layout(n.net_id, n.n_layers, n.in_dim, n.out_dim);
  return reinterpret_cast<const float*>(reinterpret_cast<const uint8_t*>(n.blob) + L.wdir_off);
}

void tc_free_net(NetTC& n) {
  if (n.blob) cudaFree(n.blob);
  if (n.error_flag) cudaFree(n.error_flag);
  if (n.dirterm) cudaFree(n.dirterm);
  n = NetTC();
}

bool tc_available() { return true; }

int tc_load_net(NetTC& n, int net_id, int n_layers, const int* in_dims, const int* out_dims, const float* const* W,
                const float* const* b, cudaStream_t stream) {
  tc_free_net(n);
  n.net_id = net_id;
  n.n_layers = n_layers;
  for (int l = 0; l < n_layers; ++l) { n.in_dim[l] = in_dims[l]; n.out_dim[l] = out_dims[l]; }
  n.supported = in_dims[0] <= 8 * 64 && out_dims[n_layers - 1] <= 96;
  if (!n.supported) return PN_OK;                           // fp32 tier still works; the fp16 launch reports it
  TcLayout L = tc_layout(net_id, n_layers, in_dims, out_dims);
  PN_CUDA_OK(cudaMalloc(&n.blob, L.total));
  PN_CUDA_OK(cudaMalloc((void**)&n.error_flag, sizeof(int)));
  PN_CUDA_OK(cudaMemsetAsync(n.error_flag, 0, sizeof(int), stream));
  n.blob_bytes = L.total;
  uint8_t* blob = reinterpret_cast<uint8_t*>(n.blob);
  for (int l = 0; l < n_layers; ++l) {
    int total = L.kblocks[l] * L.n_pad[l] * 64;
    tc::pack_tc_kernel<<<(total + 255) / 256, 256, 0, stream>>>(W[l], out_dims[l], in_dims[l], L.k_used[l], 1, 0, L.n_pad[l], L.kblocks[l],
                                                                 L.merged[l] ? 1 : 0, blob + L.chunk_off[l]);
    PN_LAUNCH_OK("pack_tc_kernel");
    tc::pack_tc_bias_kernel<<<1, kHidden, 0, stream>>>(b[l], out_dims[l], reinterpret_cast<float*>(blob + L.bias_off) + (size_t)l * kHidden);
    PN_LAUNCH_OK("pack_tc_bias_kernel");
  }
  if (L.has_fold) {
    tc::pack_tc_kernel<<<(kHidden * 64 + 255) / 256, 256, 0, stream>>>(W[0], out_dims[0], in_dims[0], 6, in_dims[0] / 6, 6, kHidden, 1, 0,
                                                                        blob + L.fold_off);
    PN_LAUNCH_OK("pack_tc_kernel(fold)");
  }
  if (net_id == PN_NET_NERF) {
    tc::pack_tc_wdir_kernel<<<1, 128, 0, stream>>>(W[n_layers - 1], in_dims[n_layers - 1], reinterpret_cast<float*>(blob + L.wdir_off));
    PN_LAUNCH_OK("pack_tc_wdir_kernel");
  }
  n.loaded = true;
  return PN_OK;
}

// ------------------------------------------------------------------------------------------------ classic NeRF
// helpers.py:792-847 on the phase machinery: 13 phases.  Tensors in checkpoint order: pts_linears.0..7 (0..7), alpha_linear (8),
// feature_linear (9), views_linears.0 (10), rgb_linear (11).
struct ClassicPhase { int tensor, k_src0, k_used, n_out, bias_row, epi, act_none, more_src, side; };
static const ClassicPhase kClassicPhases[13] = {
    {0, 0, 63, 256, 0, tc::EPI_HIDDEN, 0, 0, 0},
    {1, 0, 256, 256, 1, tc::EPI_HIDDEN, 0, 0, 0},
    {2, 0, 256, 256, 2, tc::EPI_HIDDEN, 0, 0, 0},
    {3, 0, 256, 256, 3, tc::EPI_HIDDEN, 0, 0, 0},
    {4, 0, 256, 256, 4, tc::EPI_HIDDEN, 0, 0, 0},
    {5, 63, 256, 256, 5, tc::EPI_MORE, 0, tc::MORE_PTS, 0},       // W5[:, 63:319] . h   (h = cat([input_pts, h]), helpers.py:833-834)
    {5, 0, 63, 256, 5, tc::EPI_HIDDEN, 0, 0, 0},                  // + W5[:, 0:63] . gamma(pts)
    {6, 0, 256, 256, 6, tc::EPI_HIDDEN, 0, 0, 0},
    {7, 0, 256, 256, 7, tc::EPI_HIDDEN, 0, 0, 1},                 // its epilogue also evaluates alpha_linear on the fp32 activations
    {9, 0, 256, 256, 8, tc::EPI_HIDDEN, 1, 0, 0},                 // feature_linear: no activation
    {10, 0, 256, 128, 9, tc::EPI_MORE, 0, tc::MORE_DIRS, 0},      // Wv[:, 0:256] . feature
    {10, 256, 27, 128, 9, tc::EPI_HIDDEN, 0, 0, 0},               // + Wv[:, 256:283] . gamma(viewdir), ReLU
    {11, 0, 128, 3, 10, tc::EPI_OUT, 0, 0, 0},                    // rgb_linear
};
constexpr int kClassicAlphaRow = 11;
constexpr int kClassicBiasRows = 12;

struct ClassicLayout { int kblocks[13], n_pad[13]; bool merged[13]; size_t w_off[13]; size_t img_bytes, bias_off, total; };
static ClassicLayout classic_layout() {
  ClassicLayout L{};
  size_t off = 0;
  for (int i = 0; i < 13; ++i) {
    const ClassicPhase& c = kClassicPhases[i];
    L.kblocks[i] = (c.k_used + 63) / 64;
    L.n_pad[i] = c.epi == tc::EPI_OUT ? 16 : c.n_out;
    L.merged[i] = c.epi == tc::EPI_OUT;
    L.w_off[i] = off;
    off += (size_t)L.kblocks[i] * L.n_pad[i] * 128;
  }
  L.img_bytes = off;
  L.bias_off = (off + 255) & ~(size_t)255;
  L.total = L.bias_off + (size_t)kClassicBiasRows * kHidden * 4;
  return L;
}

__global__ void pack_classic_alpha_kernel(const float* __restrict__ w_alpha, float* __restrict__ dst) {
  if (threadIdx.x < kHidden) dst[threadIdx.x] = w_alpha[threadIdx.x];
}

int tc_load_nerf_classic(NetTC& n, const int* in_dims, const int* out_dims, const float* const* W, const float* const* b,
                         cudaStream_t stream) {
  tc_free_net(n);
  n.net_id = PN_NET_NERF;
  n.classic = true;
  n.supported = true;
  const ClassicLayout L = classic_layout();
  PN_CUDA_OK(cudaMalloc(&n.blob, L.total));
  PN_CUDA_OK(cudaMemsetAsync(n.blob, 0, L.total, stream));
  PN_CUDA_OK(cudaMalloc((void**)&n.error_flag, sizeof(int)));
  PN_CUDA_OK(cudaMemsetAsync(n.error_flag, 0, sizeof(int), stream));
  n.blob_bytes = L.total;
  uint8_t* blob = reinterpret_cast<uint8_t*>(n.blob);
  float* bias = reinterpret_cast<float*>(blob + L.bias_off);
  for (int i = 0; i < 13; ++i) {
    const ClassicPhase& c = kClassicPhases[i];
    const int total = L.kblocks[i] * L.n_pad[i] * 64;
    tc::pack_tc_kernel<<<(total + 255) / 256, 256, 0, stream>>>(W[c.tensor], out_dims[c.tensor], in_dims[c.tensor], c.k_used, 1, 0, L.n_pad[i],
                                                                 L.kblocks[i], L.merged[i] ? 1 : 0, blob + L.w_off[i], c.k_src0);
    PN_LAUNCH_OK("pack_tc_kernel(classic)");
    tc::pack_tc_bias_kernel<<<1, kHidden, 0, stream>>>(b[c.tensor], out_dims[c.tensor], bias + (size_t)c.bias_row * kHidden);
    PN_LAUNCH_OK("pack_tc_bias_kernel(classic)");
  }
  pack_classic_alpha_kernel<<<1, kHidden, 0, stream>>>(W[8], bias + (size_t)kClassicAlphaRow * kHidden);
  PN_LAUNCH_OK("pack_classic_alpha_kernel");
  PN_CUDA_OK(cudaMemcpyAsync(&n.alpha_bias, b[8], sizeof(float), cudaMemcpyDeviceToHost, stream));
  PN_CUDA_OK(cudaStreamSynchronize(stream));
  n.loaded = true;
  return PN_OK;
}

static int tc_max_clusters(const void* func, int nthreads, int* out);
constexpr int kMaxDevices = 64;

// The [M, K0] fp16 first-layer input as a 2-D tensor map (box = one K block of a 128-row tile, 128-byte swizzle, zero fill outside).
// cuTensorMapEncodeTiled is a driver entry point; it is looked up at run time so that the library needs no -lcuda.
static int tc_encode_input_map(CUtensorMap* map, const void* base, long long M, int k0) {
  typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                               const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  static EncodeFn encode = nullptr;
  if (!encode) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult q = cudaDriverEntryPointSymbolNotFound;
    PN_CUDA_OK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
    if (q != cudaDriverEntryPointSuccess || !fn) { set_error("tc: the driver does not export cuTensorMapEncodeTiled"); return PN_ECUDA; }
    encode = reinterpret_cast<EncodeFn>(fn);
  }
  const cuuint64_t gdim[2] = {(cuuint64_t)k0, (cuuint64_t)M};
  const cuuint64_t gstride[1] = {(cuuint64_t)k0 * 2};               // bytes between rows (a multiple of 16: K0 % 8 == 0)
  const cuuint32_t box[2] = {64, (cuuint32_t)tc::TILE_M};
  const cuuint32_t estride[2] = {1, 1};
  const CUresult rc = encode(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(base), gdim, gstride, box, estride, CU_TENSOR_MAP_INTERLEAVE_NONE,
                             CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (rc != CUDA_SUCCESS) { set_error("tc: cuTensorMapEncodeTiled failed (%d) for a [%lld, %d] fp16 input", (int)rc, M, k0); return PN_ECUDA; }
  return PN_OK;
}

// run_network with the classic NeRF: pts [M,3] (M = N*S rows), per-ray view directions, both encoded in-kernel -> raw [M,4]
int tc_launch_nerf_classic(NetTC& n, const float* pts, const float* viewdirs, int viewdir_stride, int S, int64_t M, float* raw,
                           cudaStream_t stream) {
  if (!n.loaded || !n.classic) { set_error("classic NeRF weights not loaded on the tensor-core tier"); return PN_ESTATE; }
  if (M == 0) return PN_OK;
  PN_REQUIRE(M < (1LL << 32), "tc: %lld rows in one launch (limit 2^32 - 1)", (long long)M);
  const ClassicLayout L = classic_layout();
  tc::Params p{};
  const uint8_t* blob = reinterpret_cast<const uint8_t*>(n.blob);
  p.wimg = blob;
  p.bias = reinterpret_cast<const float*>(blob + L.bias_off);
  p.n_layers = 13; p.n_bias_rows = kClassicBiasRows; p.n_out = 4; p.k0 = 63;
  p.in0 = pts; p.in_stride = 3; p.M = M; p.out = raw;
  p.vdir = viewdirs; p.vdir_stride = viewdir_stride; p.dir_div = S > 0 ? S : 1;
  p.alpha_row = kClassicAlphaRow; p.alpha_bias = n.alpha_bias;
  p.error_flag = n.error_flag; p.timeline = nullptr; p.clk = g_tc_clock ? g_tc_clock + 8 : nullptr; p.split = 0; p.share = 1;
  p.n_phases = 13;
  for (int i = 0; i < 13; ++i) {
    const ClassicPhase& c = kClassicPhases[i];
    tc::Phase& P = p.ph[i];
    const int rem = c.k_used - (L.kblocks[i] - 1) * 64;
    P.layer = c.bias_row; P.nkb = L.kblocks[i]; P.k16_last = (rem + 15) / 16; P.n_pad = L.n_pad[i];
    P.acc = (i > 0 && kClassicPhases[i - 1].epi == tc::EPI_MORE) ? 1 : 0;
    P.epi = c.epi; P.merged = L.merged[i] ? 1 : 0; P.w_off = (uint32_t)L.w_off[i];
    P.act_none = c.act_none; P.more_src = c.more_src; P.side = c.side;
  }
  auto kern = tc::mlp_tc_kernel<0, IN_CLASSIC>;
  static int max_clusters_dev[kMaxDevices] = {0};          // per device: attribute + occupancy are set / queried on the current one
  int dev_id = 0;
  PN_CUDA_OK(cudaGetDevice(&dev_id));
  PN_REQUIRE(dev_id >= 0 && dev_id < kMaxDevices, "tc: device index %d out of range", dev_id);
  int& max_clusters = max_clusters_dev[dev_id];
  if (max_clusters == 0) {
    PN_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::SMEM_ALLOC));
    int rc = tc_max_clusters((const void*)kern, tc::n_threads(0), &max_clusters);
    if (rc != PN_OK) return rc;
    if (max_clusters <= 0) { set_error("tc: no co-resident CTA pair fits on this device"); return PN_ECUDA; }
  }
  const long long units = (M + tc::UNIT_M - 1) / tc::UNIT_M;
  const unsigned clusters = (unsigned)(units < max_clusters ? units : max_clusters);
  PN_CUDA_OK(launch_chain(kern, dim3(2 * clusters), dim3(tc::n_threads(0)), tc::SMEM_ALLOC, stream, p));
  PN_LAUNCH_OK("mlp_tc_kernel(classic)");
  return PN_OK;
}

static int tc_max_clusters(const void* func, int nthreads, int* out) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(2 * 148);
  cfg.blockDim = dim3(nthreads);
  cfg.dynamicSmemBytes = tc::SMEM_ALLOC;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  int n = 0;
  PN_CUDA_OK(cudaOccupancyMaxActiveClusters(&n, func, &cfg));
  *out = n;
  return PN_OK;
}

int tc_launch_mlp(NetTC& n, const MlpLaunch& Lc, cudaStream_t stream) {
  if (!n.loaded) {
    set_error(n.supported ? "PN_PREC_FP16: network weights not loaded" : "PN_PREC_FP16: this network shape is outside the tensor-core "
              "kernel's limits (first layer <= 512 inputs, output <= 96); use PN_PREC_FP32");
    return PN_ESTATE;
  }
  if (Lc.M == 0) return PN_OK;
  PN_REQUIRE(Lc.M < (1LL << 32), "tc: %lld rows in one launch (limit 2^32 - 1)", (long long)Lc.M);
  TcLayout L = tc_layout(n.net_id, n.n_layers, n.in_dim, n.out_dim);
  tc::Params p{};
  const uint8_t* blob = reinterpret_cast<const uint8_t*>(n.blob);
  p.wimg = blob;
  p.bias = reinterpret_cast<const float*>(blob + L.bias_off);
  p.n_layers = n.n_layers;
  p.n_bias_rows = n.n_layers;
  p.n_out = n.out_dim[n.n_layers - 1];
  p.k0 = n.in_dim[0];
  p.in0 = Lc.in0; p.in_stride = Lc.in_stride;
  p.M = Lc.M; p.out = Lc.out;
  uint32_t nonlinear_blocks = 0;
  for (int c = 0; c < 96; ++c) {
    int kind = HEAD_NONE;
    for (int g = 0; g < 3; ++g)
      if (c >= Lc.head_lo[g] && c < Lc.head_lo[g + 1]) kind = Lc.head_act[g];
    p.head_tab[c] = tc::head_coeffs(kind);
    if (kind != HEAD_NONE && c < p.n_out) nonlinear_blocks |= 1u << (c / 8);
  }
  p.head_linear = ~nonlinear_blocks;
  p.error_flag = n.error_flag;
  p.timeline = g_tc_timeline;
  p.clk = g_tc_clock ? g_tc_clock + 4 * (n.net_id == PN_NET_SAMPLER ? 0 : n.net_id == PN_NET_REFINE ? 1 : 2) : nullptr;
  {
    // schedule knobs (defaults = the measured best; the environment overrides are a tuning aid)
    static const int env_split = getenv("PN_TC_SPLIT") ? atoi(getenv("PN_TC_SPLIT")) : -1;
    p.split = env_split >= 0 ? env_split : 0;            // bit 0: every hidden epilogue, bit 1: the first layer's only
    static const int env_share = getenv("PN_TC_SHARE") ? atoi(getenv("PN_TC_SHARE")) : 1;
    p.share = env_share != 0;
    p.out_rpp = 0;         // (the hand-over warps derive the rows per staging pass from the chunk width)
  }
  if (Lc.input_mode == IN_LOAD16) {
    if (p.k0 % 8 != 0 || Lc.in_stride != p.k0 || (reinterpret_cast<uintptr_t>(Lc.in0) & 15) != 0) {
      set_error("tc fp16 input: needs a dense, 16-byte aligned [M, K0] fp16 tensor with K0 %% 8 == 0 (K0 = %d, stride %d)", p.k0, Lc.in_stride);
      return PN_EINVAL;
    }
    PN_REQUIRE(Lc.M < (1LL << 31), "tc fp16 input: %lld rows in one launch (tensor-map coordinates are 32-bit)", (long long)Lc.M);
    const int rc = tc_encode_input_map(&p.tmap_in, Lc.in0, Lc.M, p.k0);
    if (rc != PN_OK) return rc;
  }
  if (Lc.input_mode == IN_PLUECKER) {
    if (6 * Lc.P != n.in_dim[0] || !L.has_fold) { set_error("tc sampler: 6P != first-layer width"); return PN_EINVAL; }
  }
  // phase table: layer 0 (possibly split when wider than 256 inputs, or folded), hidden layers, output layer
  int np = 0;
  auto add_phase = [&](int layer, int nkb, int k16_last, int n_pad, int acc, int epi, size_t w_off, bool merged = false) {
    tc::Phase& P = p.ph[np++];
    P.layer = layer; P.nkb = nkb; P.k16_last = k16_last; P.n_pad = n_pad; P.acc = acc; P.epi = epi; P.merged = merged ? 1 : 0; P.w_off = (uint32_t)w_off;
  };
  const int last = n.n_layers - 1;
  if (Lc.input_mode == IN_PLUECKER) {
    add_phase(0, 1, 1, kHidden, 0, tc::EPI_HIDDEN, L.fold_off);                 // 6 inputs: one K = 16 step
  } else {
    const int kb0 = L.kblocks[0];
    auto k16 = [&](int kb_count, int k_used) { int rem = k_used - (kb_count - 1) * 64; return (rem + 15) / 16; };
    if (kb0 <= 4) {
      add_phase(0, kb0, k16(kb0, L.k_used[0]), kHidden, 0, tc::EPI_HIDDEN, L.chunk_off[0]);
    } else {
      add_phase(0, 4, 4, kHidden, 0, tc::EPI_MORE, L.chunk_off[0]);
      add_phase(0, kb0 - 4, k16(kb0 - 4, L.k_used[0] - 256), kHidden, 1, tc::EPI_HIDDEN, L.chunk_off[0] + (size_t)4 * kHidden * 128);
    }
  }
  for (int l = 1; l < last; ++l) add_phase(l, 4, 4, kHidden, 0, tc::EPI_HIDDEN, L.chunk_off[l]);
  add_phase(last, 4, 4, L.n_pad[last], 0, tc::EPI_OUT, L.chunk_off[last], L.merged[last]);
  p.n_phases = np;

  // NeRF: view-direction term of the last layer, fp32, one row per ray (run_network) or per sample (forward)
  if (n.net_id == PN_NET_NERF && Lc.input_mode == IN_ENCODE && Lc.dirterm_ready) {
    p.dirterm = Lc.dirterm_ready;                           // computed upstream (interval_refine_dnorm)
    p.dir_div = Lc.S > 0 ? Lc.S : 1;
  } else if (n.net_id == PN_NET_NERF) {
    const long long rows = Lc.input_mode == IN_ENCODE ? Lc.M / (Lc.S > 0 ? Lc.S : 1) : Lc.M;
    if ((size_t)rows > n.dirterm_rows) {
      if (n.dirterm) cudaFree(n.dirterm);
      n.dirterm = nullptr; n.dirterm_rows = 0;
      PN_CUDA_OK(cudaMalloc((void**)&n.dirterm, (size_t)rows * 4 * sizeof(float)));
      n.dirterm_rows = (size_t)rows;
    }
    const float* wdir = reinterpret_cast<const float*>(blob + L.wdir_off);
    const unsigned blocks = (unsigned)((rows + 255) / 256 < 148 * 8 ? (rows + 255) / 256 : 148 * 8);
    if (Lc.input_mode == IN_ENCODE) PN_CUDA_OK(launch_chain(tc::dirterm_kernel, dim3(blocks), dim3(256), 0, stream, Lc.in1, Lc.in1_stride, 0, wdir, rows, n.dirterm));
    else PN_CUDA_OK(launch_chain(tc::dirterm_kernel, dim3(blocks), dim3(256), 0, stream, Lc.in1, 27, 1, wdir, rows, n.dirterm));
    PN_LAUNCH_OK("dirterm_kernel");
    p.dirterm = n.dirterm;
    p.dir_div = Lc.input_mode == IN_ENCODE ? (Lc.S > 0 ? Lc.S : 1) : 1;
  }

  const long long units = (Lc.M + tc::UNIT_M - 1) / tc::UNIT_M;
  int dev_id = 0;
  PN_CUDA_OK(cudaGetDevice(&dev_id));
  PN_REQUIRE(dev_id >= 0 && dev_id < kMaxDevices, "tc: device index %d out of range", dev_id);
#define PN_TC_LAUNCH(ACT, MODE)                                                                                            \
  do {                                                                                                                     \
    auto kern = tc::mlp_tc_kernel<ACT, MODE>;                                                                              \
    static int max_clusters_dev[kMaxDevices] = {0};                                                                        \
    int& max_clusters = max_clusters_dev[dev_id];                                                                          \
    if (max_clusters == 0) {                                                                                               \
      PN_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::SMEM_ALLOC));                 \
      int rc = tc_max_clusters((const void*)kern, tc::n_threads(ACT), &max_clusters);                                      \
      if (rc != PN_OK) return rc;                                                                                          \
      if (max_clusters <= 0) { set_error("tc: no co-resident CTA pair fits on this device"); return PN_ECUDA; }            \
    }                                                                                                                      \
    const unsigned clusters = (unsigned)(units < max_clusters ? units : max_clusters);                                     \
    PN_CUDA_OK(launch_chain(kern, dim3(2 * clusters), dim3(tc::n_threads(ACT)), tc::SMEM_ALLOC, stream, p));               \
  } while (0)
  if (Lc.act == 0 && Lc.input_mode == IN_ENCODE) PN_TC_LAUNCH(0, IN_ENCODE);
  else if (Lc.act == 0 && Lc.input_mode == IN_LOAD2) PN_TC_LAUNCH(0, IN_LOAD2);
  else if (Lc.act == 1 && Lc.input_mode == IN_PLUECKER) PN_TC_LAUNCH(1, IN_PLUECKER);
  else if (Lc.act == 1 && Lc.input_mode == IN_LOAD) PN_TC_LAUNCH(1, IN_LOAD);
  else if (Lc.act == 1 && Lc.input_mode == IN_LOAD16) PN_TC_LAUNCH(1, IN_LOAD16);
  else { set_error("tc: unsupported (activation, input mode) = (%d, %d)", Lc.act, Lc.input_mode); return PN_EINVAL; }
#undef PN_TC_LAUNCH
  PN_LAUNCH_OK("mlp_tc_kernel");
  return PN_OK;
}

}  // namespace pn
