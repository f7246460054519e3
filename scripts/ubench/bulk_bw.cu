// Micro-benchmark (debug aid, not product): cp.async.bulk global->shared streaming rate of an L2-resident weight image
// through an mbarrier ring, per SM and chip-wide.  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o bulk_bw bulk_bw.cu
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
// one thread streams `total_chunks` chunks of `chunk` bytes, `slots` in flight, reading a `img_bytes` image cyclically
// starting at a per-CTA offset (stagger) -- `split` = number of cp.async.bulk instructions per chunk
__global__ void __launch_bounds__(128, 1) stream_kernel(const uint8_t* img, uint32_t img_bytes, uint32_t chunk, int slots, int total_chunks,
                                                         int stagger, int split, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  __shared__ __align__(8) uint64_t bars[8];
  if (threadIdx.x == 0) {
    for (int s = 0; s < slots; ++s) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bars[s])) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    const uint32_t n_img_chunks = img_bytes / chunk;
    uint32_t src_idx = stagger ? (blockIdx.x * 7u) % n_img_chunks : 0u;
    long long t0 = clock64();
    const uint32_t part = chunk / split;
    for (int g = 0; g < total_chunks + slots; ++g) {
      const int s = g % slots;
      if (g >= slots) { while (!mbar_try_wait(smem_u32(&bars[s]), ((g / slots) - 1) & 1)) {} }
      if (g < total_chunks) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bars[s])), "r"(chunk) : "memory");
        for (int q = 0; q < split; ++q)
          asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(base + s * chunk + q * part),
                       "l"(img + (size_t)src_idx * chunk + q * part), "r"(part), "r"(smem_u32(&bars[s])) : "memory");
        src_idx = (src_idx + 1) % n_img_chunks;
      }
    }
    out[blockIdx.x] = clock64() - t0;
  }
}
int main() {
  const uint32_t img_bytes = 832 * 1024;   // ~ one NeRF weight image
  uint8_t* d_img; long long* d_out;
  cudaMalloc(&d_img, img_bytes); cudaMemset(d_img, 1, img_bytes);
  cudaMalloc(&d_out, 148 * 8);
  const int smem = 200 * 1024;
  cudaFuncSetAttribute(stream_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  long long h[148];
  auto run = [&](int grid, uint32_t chunk, int slots, int stagger, int split) {
    const int total = 4096 * 1024 / chunk * 4;   // 16 MB per CTA
    for (int rep = 0; rep < 2; ++rep) stream_kernel<<<grid, 128, smem>>>(d_img, img_bytes, chunk, slots, total, stagger, split, d_out);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("CUDA error %s\n", cudaGetErrorString(e)); return; }
    cudaMemcpy(h, d_out, sizeof(long long) * grid, cudaMemcpyDeviceToHost);
    long long mx = 0, mn = 1LL << 60;
    for (int i = 0; i < grid; ++i) { if (h[i] > mx) mx = h[i]; if (h[i] < mn) mn = h[i]; }
    double bytes = (double)total * chunk;
    printf("grid %3d chunk %5u B slots %d stagger %d split %d: %6.1f B/cyc/SM (slowest) %6.1f (fastest) -> chip %7.0f B/cyc; %6.0f cyc per chunk-slot\n", grid, chunk,
           slots, stagger, split, bytes / mx, bytes / mn, bytes / mx * grid, (double)mx / total * slots);
  };
  for (int grid : {1, 16, 74, 148}) {
    run(grid, 32768, 4, 0, 1);
    run(grid, 32768, 4, 1, 1);
    run(grid, 32768, 4, 0, 4);
    run(grid, 16384, 4, 0, 1);
    run(grid, 16384, 8, 0, 1);
    run(grid, 16384, 8, 1, 1);
    run(grid, 8192, 8, 0, 1);
  }
  return 0;
}
