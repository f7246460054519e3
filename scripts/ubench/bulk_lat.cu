// Micro-benchmark (debug aid): where does the time go in a bulk-copy ring?  Per-iteration clock stamps.
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
__device__ __forceinline__ bool mbar_test_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
               : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  return ok != 0;
}
// variant 0: try_wait spin; 1: test_wait spin; 2: issue-only latency probe: one copy at a time (issue, wait, repeat)
// 3: like 0 but copies issued from `nlanes` different warps round-robin (each warp owns slots s % nwarps)
template <int SLOTS>
__global__ void __launch_bounds__(256, 1) k(const uint8_t* img, uint32_t chunk, int variant, int iters, long long* out, long long* stamps) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  __shared__ __align__(8) uint64_t bars[SLOTS];
  if (threadIdx.x == 0) {
    for (int s = 0; s < SLOTS; ++s) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bars[s])) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (variant <= 2 && threadIdx.x == 0) {
    long long t0 = clock64();
    if (variant == 2) {
      for (int g = 0; g < iters; ++g) {
        long long a = clock64();
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bars[0])), "r"(chunk) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(base),
                     "l"(img + (size_t)(g & 15) * chunk), "r"(chunk), "r"(smem_u32(&bars[0])) : "memory");
        long long b = clock64();
        while (!mbar_try_wait(smem_u32(&bars[0]), g & 1)) {}
        long long c = clock64();
        if (blockIdx.x == 0 && g < 64) { stamps[g * 3] = a - t0; stamps[g * 3 + 1] = b - t0; stamps[g * 3 + 2] = c - t0; }
      }
    } else {
      for (int g = 0; g < iters + SLOTS; ++g) {
        const int s = g % SLOTS;
        long long a = clock64();
        if (g >= SLOTS) {
          const uint32_t par = ((g / SLOTS) - 1) & 1;
          if (variant == 0) while (!mbar_try_wait(smem_u32(&bars[s]), par)) {}
          else while (!mbar_test_wait(smem_u32(&bars[s]), par)) {}
        }
        long long b = clock64();
        if (g < iters) {
          asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bars[s])), "r"(chunk) : "memory");
          asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(base + s * chunk),
                       "l"(img + (size_t)(g & 15) * chunk), "r"(chunk), "r"(smem_u32(&bars[s])) : "memory");
        }
        long long c = clock64();
        if (blockIdx.x == 0 && g < 64) { stamps[g * 3] = a - t0; stamps[g * 3 + 1] = b - t0; stamps[g * 3 + 2] = c - t0; }
      }
    }
    out[blockIdx.x] = clock64() - t0;
  }
  if (variant == 3 && warp < SLOTS && lane == 0) {
    // each warp owns one slot and streams independently
    long long t0 = clock64();
    const int n = iters / SLOTS;
    for (int g = 0; g < n; ++g) {
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bars[warp])), "r"(chunk) : "memory");
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(base + warp * chunk),
                   "l"(img + (size_t)((g * SLOTS + warp) & 15) * chunk), "r"(chunk), "r"(smem_u32(&bars[warp])) : "memory");
      while (!mbar_try_wait(smem_u32(&bars[warp]), g & 1)) {}
    }
    if (warp == 0) out[blockIdx.x] = clock64() - t0;
  }
}
int main() {
  uint8_t* d_img; long long *d_out, *d_st;
  cudaMalloc(&d_img, 1 << 20); cudaMemset(d_img, 1, 1 << 20);
  cudaMalloc(&d_out, 148 * 8); cudaMalloc(&d_st, 64 * 3 * 8);
  const int smem = 200 * 1024;
  cudaFuncSetAttribute(k<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaFuncSetAttribute(k<6>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  long long h[148], st[192];
  auto run = [&](int slots, int grid, uint32_t chunk, int variant, bool dump) {
    const int iters = 2048;
    cudaMemset(d_st, 0, sizeof(st));
    for (int rep = 0; rep < 2; ++rep) {
      if (slots == 4) k<4><<<grid, 256, smem>>>(d_img, chunk, variant, iters, d_out, d_st);
      else k<6><<<grid, 256, smem>>>(d_img, chunk, variant, iters, d_out, d_st);
    }
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("CUDA error %s\n", cudaGetErrorString(e)); return; }
    cudaMemcpy(h, d_out, 8 * grid, cudaMemcpyDeviceToHost);
    cudaMemcpy(st, d_st, sizeof(st), cudaMemcpyDeviceToHost);
    long long mx = 0;
    for (int i = 0; i < grid; ++i) if (h[i] > mx) mx = h[i];
    printf("variant %d slots %d grid %3d chunk %5u: %7.1f cyc/chunk -> %6.1f B/cyc/SM\n", variant, slots, grid, chunk, (double)mx / iters,
           (double)iters * chunk / mx);
    if (dump) for (int g = 0; g < 24; ++g) printf("   it %2d: start %6lld  wait_done %6lld (+%lld)  issued %6lld (+%lld)\n", g, st[g * 3], st[g * 3 + 1],
                                                    st[g * 3 + 1] - st[g * 3], st[g * 3 + 2], st[g * 3 + 2] - st[g * 3 + 1]);
  };
  run(4, 1, 32768, 0, true);
  run(4, 1, 32768, 1, true);
  run(4, 1, 32768, 2, true);
  run(4, 1, 8192, 2, true);
  run(4, 1, 1024, 2, false);
  run(4, 1, 32768, 3, false);
  run(6, 1, 32768, 3, false);
  run(6, 1, 16384, 3, false);
  run(6, 1, 16384, 0, false);
  run(4, 148, 32768, 3, false);
  run(6, 148, 32768, 3, false);
  run(6, 148, 16384, 3, false);
  return 0;
}
