// Micro-benchmark (debug aid): issue rate of the epilogue's candidate instructions, in cycles per warp-instruction per SM
// sub-partition (8 warps per SM, 2 per scheduler, independent chains).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define REP16(X) X X X X X X X X X X X X X X X X
template <int OP>
__global__ void __launch_bounds__(256, 1) k(uint32_t* out, int iters, float seed) {
  uint32_t a[16]; float f[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) { a[i] = threadIdx.x * 17 + i; f[i] = seed * (float)(threadIdx.x + i); }
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      if (OP == 0) asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(a[i]) : "f"(__uint_as_float(a[i])), "f"(__uint_as_float(a[(i + 1) & 15])));
      if (OP == 1) asm volatile("cvt.rn.relu.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(a[i]) : "f"(__uint_as_float(a[i])), "f"(__uint_as_float(a[(i + 1) & 15])));
      if (OP == 2) asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(a[i]) : "f"(__uint_as_float(a[i])), "f"(__uint_as_float(a[(i + 1) & 15])));
      if (OP == 3) asm volatile("{.reg .b64 x, y, z; mov.b64 x, {%0, %1}; mov.b64 y, {%2, %2}; add.rn.f32x2 z, x, y; mov.b64 {%0, %1}, z;}" : "+f"(f[i]), "+f"(f[(i + 8) & 15]) : "f"(seed));
      if (OP == 4) asm volatile("prmt.b32 %0, %0, %1, 0x7632;" : "+r"(a[i]) : "r"(a[(i + 1) & 15]));
      if (OP == 5) asm volatile("shf.r.clamp.b32 %0, %0, %1, 13;" : "+r"(a[i]) : "r"(a[(i + 1) & 15]));
      if (OP == 6) asm volatile("max.s32 %0, %0, %1;" : "+r"(a[i]) : "r"(a[(i + 1) & 15]));
      if (OP == 7) asm volatile("fma.rn.f16x2 %0, %0, %1, %1;" : "+r"(a[i]) : "r"(a[(i + 1) & 15]));
      if (OP == 8) asm volatile("max.f16x2 %0, %0, %1;" : "+r"(a[i]) : "r"(a[(i + 1) & 15]));
      if (OP == 9) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(f[i]));
      if (OP == 10) asm volatile("fma.rn.f32 %0, %0, %1, %1;" : "+f"(f[i]) : "f"(seed));
      if (OP == 11) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a[i]) : "r"(a[(i + 1) & 15]), "r"(a[(i + 2) & 15]));
      if (OP == 12) asm volatile("max.f32 %0, %0, %1;" : "+f"(f[i]) : "f"(seed));
      if (OP == 13) asm volatile("{.reg .b64 x, y, z; mov.b64 x, {%0, %1}; mov.b64 y, {%2, %2}; fma.rn.f32x2 z, x, y, y; mov.b64 {%0, %1}, z;}" : "+f"(f[i]), "+f"(f[(i + 8) & 15]) : "f"(seed));
      if (OP == 14) { unsigned short hh; asm volatile("cvt.rn.f16.f32 %0, %1;" : "=h"(hh) : "f"(__uint_as_float(a[i]))); a[i] = hh * 65537u; }
      if (OP == 15) asm volatile("mad.lo.u32 %0, %0, %1, %1;" : "+r"(a[i]) : "r"(a[(i + 1) & 15]));
      if (OP == 16) asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(a[i]));
      if (OP == 17) asm volatile("ex2.approx.ftz.bf16x2 %0, %0;" : "+r"(a[i]));
      if (OP == 18) asm volatile("tanh.approx.f32 %0, %0;" : "+f"(f[i]));
      if (OP == 19) asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(f[i]) : "f"(seed));
      if (OP == 20) asm volatile("mul.rn.f32 %0, %0, %1;" : "+f"(f[i]) : "f"(seed));
      if (OP == 21) asm volatile("add.rn.f16x2 %0, %0, %1;" : "+r"(a[i]) : "r"(a[(i + 1) & 15]));
      if (OP == 22) asm volatile("{.reg .pred p; setp.gt.f32 p, %0, %1; selp.f32 %0, %0, %1, p;}" : "+f"(f[i]) : "f"(seed));
    }
  }
  long long t1 = clock64();
  uint32_t acc = 0;
#pragma unroll
  for (int i = 0; i < 16; ++i) acc ^= a[i] ^ __float_as_uint(f[i]);
  out[blockIdx.x * 256 + threadIdx.x] = acc;
  if (threadIdx.x == 0) out[148 * 256 + blockIdx.x] = (uint32_t)(t1 - t0);
}
int main() {
  uint32_t* d; cudaMalloc(&d, (148 * 256 + 148) * 4);
  const char* names[] = {"cvt.rn.f16x2.f32 (F2FP)", "cvt.rn.relu.satfinite.f16x2.f32", "cvt.rn.bf16x2.f32", "add.f32x2 (FADD2)", "prmt", "shf.r", "max.s32 (IMNMX)",
                         "fma.f16x2 (HFMA2)", "max.f16x2 (HMNMX2)", "ex2.approx.f32 (MUFU)", "fma.f32 (FFMA)", "lop3", "max.f32 (FMNMX)", "fma.f32x2 (FFMA2)", "cvt.rn.f16.f32 (F2F)", "mad.lo.u32 (IMAD)", "ex2.approx.f16x2 (2 MUFU.F16 + 2 PRMT)", "ex2.approx.ftz.bf16x2", "tanh.approx.f32", "add.f32 (FADD)", "mul.f32 (FMUL)", "add.f16x2 (HADD2)", "setp+selp f32"};
  const int iters = 2000;
  uint32_t h[148];
#define RUN(OP) { k<OP><<<148, 256>>>(d, iters, 1.0001f); cudaDeviceSynchronize(); k<OP><<<148, 256>>>(d, iters, 1.0001f); cudaError_t e = cudaDeviceSynchronize(); \
    cudaMemcpy(h, d + 148 * 256, 148 * 4, cudaMemcpyDeviceToHost); double cyc = h[0]; \
    printf("%-36s %s: %.2f cycles per warp-instruction per scheduler  (%.1f lanes/clk/SM)\n", names[OP], e == cudaSuccess ? "ok" : cudaGetErrorString(e), cyc / (iters * 16.0 * 2.0), 32.0 * 4 / (cyc / (iters * 16.0 * 2.0))); }
  RUN(0) RUN(1) RUN(2) RUN(3) RUN(4) RUN(5) RUN(6) RUN(7) RUN(8) RUN(9) RUN(10) RUN(11) RUN(12) RUN(13) RUN(14) RUN(15) RUN(16) RUN(17) RUN(18) RUN(19) RUN(20) RUN(21) RUN(22)
  return 0;
}
