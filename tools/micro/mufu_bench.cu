// Microbenchmark: MUFU ex2 throughput f32 vs f16x2 vs bf16x2, and packed fp32x2 FMA, on sm_100a.
#include <cstdio>
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <cuda_bf16.h>
#include <stdint.h>

template <int MODE>
__global__ void k(float* out, int iters) {
  float a0 = threadIdx.x * 1e-3f, a1 = a0 + 0.1f, a2 = a0 + 0.2f, a3 = a0 + 0.3f;
  uint32_t h0 = 0x3c003800u + threadIdx.x, h1 = h0 + 1, h2 = h0 + 2, h3 = h0 + 3;
  for (int i = 0; i < iters; ++i) {
    if (MODE == 0) {
      asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a0));
      asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a1));
      asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a2));
      asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a3));
    } else if (MODE == 1) {
      asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(h0));
      asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(h1));
      asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(h2));
      asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(h3));
    } else if (MODE == 2) {
      asm volatile("ex2.approx.ftz.bf16x2 %0, %0;" : "+r"(h0));
      asm volatile("ex2.approx.ftz.bf16x2 %0, %0;" : "+r"(h1));
      asm volatile("ex2.approx.ftz.bf16x2 %0, %0;" : "+r"(h2));
      asm volatile("ex2.approx.ftz.bf16x2 %0, %0;" : "+r"(h3));
    } else if (MODE == 3) {  // plain FFMA x4
      a0 = fmaf(a0, 1.0001f, 0.5f); a1 = fmaf(a1, 1.0001f, 0.5f); a2 = fmaf(a2, 1.0001f, 0.5f); a3 = fmaf(a3, 1.0001f, 0.5f);
    } else if (MODE == 4) {  // packed fp32x2 FMA x2 (same element count as MODE 3)
      unsigned long long p0, p1, m = 0x3f8003473f800347ull, c = 0x3f0000003f000000ull;
      asm volatile("mov.b64 %0, {%1, %2};" : "=l"(p0) : "f"(a0), "f"(a1));
      asm volatile("mov.b64 %0, {%1, %2};" : "=l"(p1) : "f"(a2), "f"(a3));
      asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p0) : "l"(m), "l"(c));
      asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p1) : "l"(m), "l"(c));
      asm volatile("mov.b64 {%0, %1}, %2;" : "=f"(a0), "=f"(a1) : "l"(p0));
      asm volatile("mov.b64 {%0, %1}, %2;" : "=f"(a2), "=f"(a3) : "l"(p1));
    }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + __uint_as_float(h0 ^ h1 ^ h2 ^ h3);
}

template <int MODE>
void run(const char* name, int elems_per_instr) {
  float* d;
  cudaMalloc(&d, 148 * 8 * 1024 * 4);
  int iters = 20000;
  k<MODE><<<148 * 4, 512>>>(d, 100);
  cudaDeviceSynchronize();
  cudaEvent_t s, e;
  cudaEventCreate(&s); cudaEventCreate(&e);
  cudaEventRecord(s);
  k<MODE><<<148 * 4, 512>>>(d, iters);
  cudaEventRecord(e);
  cudaEventSynchronize(e);
  float ms;
  cudaEventElapsedTime(&ms, s, e);
  double instr = 148.0 * 4 * 512 * iters * 4;
  int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  printf("%-28s %8.3f ms  %.2f G thread-instr/s  %.1f elems/clk/SM (at %d MHz nominal) err=%s\n", name, ms, instr / ms / 1e6,
         instr * elems_per_instr / (ms * 1e-3) / 148 / (clk * 1e3), clk / 1000, cudaGetErrorString(cudaGetLastError()));
  cudaFree(d);
}

int main() {
  run<0>("ex2.approx.ftz.f32", 1);
  run<1>("ex2.approx.f16x2", 2);
  run<2>("ex2.approx.ftz.bf16x2", 2);
  run<3>("ffma f32 (x4)", 1);
  run<4>("fma.rn.f32x2 (x2 packed)", 1);
  return 0;
}
