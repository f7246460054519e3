// Micro-benchmark (debug aid, not product): tcgen05.ld throughput per SM and tcgen05.mma issue rate, alone and together.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tmem_bw tmem_bw.cu && ./tmem_bw
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t* r) {
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
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
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
__device__ __forceinline__ void umma_bf16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
               "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
               : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  return ok != 0;
}

// mode bit0: epilogue warps run LDTM loops; bit1: warp 0 issues MMAs; bit2: epilogue warps also write 16 B / 8 columns to smem
// nw_epi epilogue warps (warps 1..nw_epi); iters x (2 x ld32 + wait)
__global__ void __launch_bounds__(288, 1) bench_kernel(int mode, int nw_epi, int iters, int n_mma, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  __shared__ uint32_t s_tmem;
  __shared__ __align__(8) uint64_t s_bar;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    if (lane == 0) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&s_bar)) : "memory");
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
  }
  // zero the operand area (A 64 KB at base, B 32 KB x 2 after it)
  for (uint32_t i = threadIdx.x; i < (192u * 1024u) / 16u; i += blockDim.x)
    asm volatile("st.shared.v4.b32 [%0], {%1,%1,%1,%1};" ::"r"(base + i * 16u), "r"(0u) : "memory");
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = s_tmem;
  long long t0 = clock64();
  if (warp == 0) {
    if ((mode & 2) && lane == 0) {
      const uint32_t idesc = umma_idesc(128, 256);
      for (int i = 0; i < n_mma; ++i) {
        const uint32_t a = base + (i & 3) * 16384u + ((i >> 2) & 3) * 32u;    // 4 K blocks of A, 4 K steps each
        const uint32_t b = base + 65536u + ((i >> 2) & 1) * 32768u + (i & 3) * 32u;
        umma_bf16(tmem + ((i >> 4) & 1) * 256, umma_desc(a), umma_desc(b), idesc, (i & 15) ? 1u : 0u);
      }
      umma_commit(smem_u32(&s_bar));
      while (!mbar_try_wait(smem_u32(&s_bar), 0)) {}
      long long t1 = clock64();
      out[blockIdx.x * 16 + 0] = t1 - t0;
    }
  } else if (warp <= nw_epi && (mode & 1)) {
    const int q = warp & 3;
    uint32_t acc = 0;
    const int half = ((warp - 1) >> 2) & 1;
    for (int it = 0; it < iters; ++it) {
      uint32_t v[64];
      const uint32_t col = (uint32_t)(((it * 2 + half) * 64) & 511);
      tmem_ld32(tmem + ((uint32_t)(q * 32) << 16) + col, v);
      tmem_ld32(tmem + ((uint32_t)(q * 32) << 16) + col + 32, v + 32);
      tmem_wait_ld();
      if (mode & 4) {
        const int r = q * 32 + lane;
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          const uint32_t off = (uint32_t)((r >> 3) * 1024 + (r & 7) * 128 + ((c ^ (r & 7)) << 4)) + ((it & 3) * 16384u);
          asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(base + 131072u + (off & 0xFFFFu)), "r"(v[8 * c] ^ v[8 * c + 1]),
                       "r"(v[8 * c + 2] ^ v[8 * c + 3]), "r"(v[8 * c + 4] ^ v[8 * c + 5]), "r"(v[8 * c + 6] ^ v[8 * c + 7]) : "memory");
        }
      } else {
#pragma unroll
        for (int i = 0; i < 64; ++i) acc ^= v[i];
      }
    }
    long long t1 = clock64();
    if (lane == 0) out[blockIdx.x * 16 + warp] = t1 - t0;
    if (acc == 0x12345678u) out[0] = 0;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
}

int main() {
  long long* d_out;
  cudaMalloc(&d_out, 148 * 16 * sizeof(long long));
  const int smem = 200 * 1024;
  cudaFuncSetAttribute(bench_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  long long h[148 * 16];
  auto run = [&](const char* name, int grid, int mode, int nw, int iters, int n_mma) {
    cudaMemset(d_out, 0, sizeof(h));
    for (int rep = 0; rep < 2; ++rep) bench_kernel<<<grid, 288, smem>>>(mode, nw, iters, n_mma, d_out);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("%s: CUDA error %s\n", name, cudaGetErrorString(e)); return; }
    cudaMemcpy(h, d_out, sizeof(h), cudaMemcpyDeviceToHost);
    long long epi_max = 0;
    for (int w = 1; w <= nw; ++w) if (h[w] > epi_max) epi_max = h[w];
    double ld_bytes = (double)nw * iters * 2 * 4096;
    printf("%-44s grid %3d: ", name, grid);
    if (mode & 1) printf("LDTM %8lld cyc -> %7.1f B/cyc/SM (%d warps)   ", epi_max, ld_bytes / (double)epi_max, nw);
    if (mode & 2) printf("MMA %8lld cyc -> %6.1f cyc/MMA", h[0], (double)h[0] / n_mma);
    printf("\n");
  };
  for (int grid : {1, 148}) {
    run("ldtm only, 1 warp", grid, 1, 1, 512, 0);
    run("ldtm only, 4 warps (one per quadrant)", grid, 1, 4, 512, 0);
    run("ldtm only, 8 warps", grid, 1, 8, 512, 0);
    run("ldtm + sts.128, 4 warps", grid, 5, 4, 512, 0);
    run("ldtm + sts.128, 8 warps", grid, 5, 8, 512, 0);
    run("mma only (M128 N256 K16 SS)", grid, 2, 0, 0, 1024);
    run("mma + ldtm 4 warps", grid, 3, 4, 512, 1024);
    run("mma + ldtm 8 warps", grid, 3, 8, 512, 1024);
    run("mma + ldtm + sts 4 warps", grid, 7, 4, 512, 1024);
    run("mma + ldtm + sts 8 warps", grid, 7, 8, 512, 1024);
  }
  return 0;
}
