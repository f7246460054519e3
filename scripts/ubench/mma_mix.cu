// Micro-benchmark (debug aid): tcgen05.mma rate (cta_group 1 and 2) with concurrent bulk-copy traffic into shared memory
// and concurrent epilogue-style LDTM + STS traffic.  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_mix mma_mix.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
               : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  return ok != 0;
}
__device__ __forceinline__ uint64_t umma_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFFu);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024u >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
__device__ __forceinline__ uint32_t umma_idesc(int m, int n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}
template <int CG>
__device__ __forceinline__ void umma_bf16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t acc) {
  if (CG == 1)
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
                 "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(acc) : "memory");
  else
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
                 "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr) : "memory");
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

// smem map (from the 1024-aligned base): A 64 KB | B 2 x 32 KB (CG=2: each CTA's half lives in the first 16 KB of a slot) |
// ring 2 x 32 KB | epilogue store target = A region
// stream_chunk > 0: warp 1 streams one chunk every stream_period cycles;  epi_warps > 0: warps 2.. run LDTM + STS loops
template <int CG>
__global__ void __launch_bounds__(320, 1) k(const uint8_t* img, int n_mma, uint32_t stream_chunk, int stream_period, int epi_warps,
                                            int epi_period, long long* out, int commit_every, int wait_every, int n_cols) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  __shared__ uint32_t s_tmem;
  __shared__ __align__(8) uint64_t s_bar[4];
  __shared__ volatile int s_stop;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint32_t cta_rank = 0;
  if (CG == 2) asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(cta_rank));
  if (threadIdx.x == 0) {
    for (int i = 0; i < 4; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&s_bar[i])) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    s_stop = 0;
  }
  if (warp == 0) {
    if (CG == 1) {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)), "r"(512) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    } else {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)), "r"(512) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
  }
  for (uint32_t i = threadIdx.x; i < (192u * 1024u) / 16u; i += blockDim.x)
    asm volatile("st.shared.v4.b32 [%0], {%1,%1,%1,%1};" ::"r"(base + i * 16u), "r"(0u) : "memory");
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (CG == 2) cluster_sync_all();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = s_tmem;
  const long long t0 = clock64();
  if (warp == 0) {
    if (lane == 0 && cta_rank == 0) {
      const uint32_t idesc = umma_idesc(128 * CG, n_cols);
      if (commit_every == 0 && wait_every == 0) {
        // lean issue: descriptors hoisted, 8 MMAs per trip (the tensor pipe, not this thread, sets the pace)
        const uint64_t a0 = umma_desc(base), b0 = umma_desc(base + 65536u);
        for (int i = 0; i < n_mma; i += 8) {
          const uint32_t d = tmem + ((i >> 4) & 1) * 256;
          umma_bf16<CG>(d, a0, b0, idesc, (i & 15) ? 1u : 0u);
          umma_bf16<CG>(d, a0 + 2, b0 + 2, idesc, 1u);
          umma_bf16<CG>(d, a0 + 4, b0 + 4, idesc, 1u);
          umma_bf16<CG>(d, a0 + 6, b0 + 6, idesc, 1u);
          umma_bf16<CG>(d, a0 + 1024, b0 + 2048, idesc, 1u);
          umma_bf16<CG>(d, a0 + 1026, b0 + 2050, idesc, 1u);
          umma_bf16<CG>(d, a0 + 1028, b0 + 2052, idesc, 1u);
          umma_bf16<CG>(d, a0 + 1030, b0 + 2054, idesc, 1u);
        }
      } else
      for (int i = 0; i < n_mma; ++i) {
        const uint32_t a = base + (i & 3) * 16384u + ((i >> 2) & 3) * 32u;
        const uint32_t b = base + 65536u + ((i >> 2) & 1) * 32768u + (i & 3) * 32u;
        umma_bf16<CG>(tmem + ((i >> 4) & 1) * 256, umma_desc(a), umma_desc(b), idesc, (i & 15) ? 1u : 0u);
        if (commit_every && (i % commit_every) == commit_every - 1) {
          if (CG == 1) asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&s_bar[3])) : "memory");
          else asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(&s_bar[3])), "h"((uint16_t)3) : "memory");
        }
        if (wait_every && (i % wait_every) == wait_every - 1) { volatile bool w = mbar_try_wait(smem_u32(&s_bar[0]), 1); (void)w; }
      }
      if (CG == 1) asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&s_bar[0])) : "memory");
      else asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(&s_bar[0])), "h"((uint16_t)3) : "memory");
    }
    if (lane == 0) {
      while (!mbar_try_wait(smem_u32(&s_bar[0]), 0)) {}
      out[blockIdx.x * 4 + 0] = clock64() - t0;
      s_stop = 1;
    }
  } else if (warp == 1) {
    if (lane == 0 && stream_chunk) {
      long long next = clock64();
      uint32_t n = 0;
      while (!s_stop) {
        const int s = n & 1;
        if (n >= 2) while (!mbar_try_wait(smem_u32(&s_bar[1 + s]), ((n >> 1) - 1) & 1)) {}
        while (clock64() < next) {}
        next += stream_period;
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&s_bar[1 + s])), "r"(stream_chunk) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(base + 131072u + s * 32768u),
                     "l"(img + (size_t)(n & 15) * 32768u), "r"(stream_chunk), "r"(smem_u32(&s_bar[1 + s])) : "memory");
        ++n;
      }
      // drain
      for (uint32_t m = (n >= 2 ? n - 2 : 0); m < n; ++m) while (!mbar_try_wait(smem_u32(&s_bar[1 + (m & 1)]), (m >> 1) & 1)) {}
      out[blockIdx.x * 4 + 1] = n;
    }
  } else if (warp - 2 < epi_warps) {
    const int q = warp & 3;
    const int r = q * 32 + lane;
    uint32_t it = 0;
    long long next = clock64();
    while (!s_stop) {
      uint32_t v[64];
      const uint32_t col = (uint32_t)(((it * 2 + ((warp - 2) >> 2)) * 64) & 511);
      tmem_ld32(tmem + ((uint32_t)(q * 32) << 16) + col, v);
      tmem_ld32(tmem + ((uint32_t)(q * 32) << 16) + col + 32, v + 32);
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        const uint32_t off = (uint32_t)((r >> 3) * 1024 + (r & 7) * 128 + ((c ^ (r & 7)) << 4)) + ((it & 3) * 16384u);
        asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(base + off), "r"(v[8 * c] ^ v[8 * c + 1]),
                     "r"(v[8 * c + 2] ^ v[8 * c + 3]), "r"(v[8 * c + 4] ^ v[8 * c + 5]), "r"(v[8 * c + 6] ^ v[8 * c + 7]) : "memory");
      }
      ++it;
      if (epi_period) { next += epi_period; while (clock64() < next && !s_stop) {} }
    }
    if (lane == 0 && warp == 2) out[blockIdx.x * 4 + 2] = it;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (CG == 2) cluster_sync_all();
  if (warp == 0) {
    if (CG == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
    else asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
  }
}

int main() {
  uint8_t* d_img; long long* d_out;
  cudaMalloc(&d_img, 1 << 20); cudaMemset(d_img, 0, 1 << 20);
  cudaMalloc(&d_out, 148 * 4 * 8);
  const int smem = 200 * 1024;
  cudaFuncSetAttribute(k<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaFuncSetAttribute(k<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  long long h[148 * 4];
  auto run = [&](int cg, int grid, uint32_t chunk, int period, int epi_warps, int epi_period, int commit_every = 0, int wait_every = 0, int n_cols = 256) {
    const int n_mma = 2048;
    cudaMemset(d_out, 0, sizeof(h));
    for (int rep = 0; rep < 2; ++rep) {
      if (cg == 1) k<1><<<grid, 320, smem>>>(d_img, n_mma, chunk, period, epi_warps, epi_period, d_out, commit_every, wait_every, n_cols);
      else {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(grid); cfg.blockDim = dim3(320); cfg.dynamicSmemBytes = smem;
        cudaLaunchAttribute at[1]; at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        cfg.attrs = at; cfg.numAttrs = 1;
        cudaLaunchKernelEx(&cfg, k<2>, (const uint8_t*)d_img, n_mma, chunk, period, epi_warps, epi_period, d_out, commit_every, wait_every, n_cols);
      }
    }
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("cg %d grid %d: CUDA error %s\n", cg, grid, cudaGetErrorString(e)); return; }
    cudaMemcpy(h, d_out, sizeof(long long) * grid * 4, cudaMemcpyDeviceToHost);
    long long mx = 0;
    for (int i = 0; i < grid; ++i) if (h[i * 4] > mx) mx = h[i * 4];
    printf("N %3d ce %d we %d cta_group %d grid %3d | stream %5u B / %4d cyc | epi warps %d period %4d : %6.1f cyc/MMA   (chunks streamed %lld = %.1f B/cyc, epi iters %lld = %.1f elem-rows)\n",
           n_cols, commit_every, wait_every, cg, grid, chunk, period, epi_warps, epi_period, (double)mx / n_mma, h[1], (double)h[1] * chunk / (double)h[0], h[2], 0.0);
  };
  for (int cg : {1, 2})
    for (int n : {16, 32, 48, 64, 128, 256}) run(cg, 148, 0, 0, 0, 0, 0, 0, n);
  return 0;
}
