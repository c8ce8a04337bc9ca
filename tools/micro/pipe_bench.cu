// Microbenchmark (sm_100a): cycles per warp-instruction on one SM sub-partition for the instruction types the
// attention softmax is made of, at 1 / 2 / 4 warps per sub-partition.  One block per SM; clock64 around the loop.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipe_bench pipe_bench.cu && ./pipe_bench
#include <cstdio>
#include <cuda_runtime.h>
#include <stdint.h>

#define F2(op, d0, d1, a0, a1, b0, b1)                                                                        \
  asm volatile("{\n\t.reg .b64 va, vb;\n\tmov.b64 va, {%2,%3};\n\tmov.b64 vb, {%4,%5};\n\t" op                 \
               " va, va, vb;\n\tmov.b64 {%0,%1}, va;\n\t}"                                                     \
               : "=f"(d0), "=f"(d1)                                                                            \
               : "f"(a0), "f"(a1), "f"(b0), "f"(b1))
#define FMA2(d0, d1, a0, a1, b0, b1, c0, c1)                                                                  \
  asm volatile("{\n\t.reg .b64 va, vb, vc;\n\tmov.b64 va, {%2,%3};\n\tmov.b64 vb, {%4,%5};\n\tmov.b64 vc, {%6,%7};\n\t" \
               "fma.rn.f32x2 va, va, vb, vc;\n\tmov.b64 {%0,%1}, va;\n\t}"                                     \
               : "=f"(d0), "=f"(d1)                                                                            \
               : "f"(a0), "f"(a1), "f"(b0), "f"(b1), "f"(c0), "f"(c1))

__device__ __forceinline__ float ex2(float x) {
  float y;
  asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ uint32_t pack(float a, float b) {
  uint32_t r;
  asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a));
  return r;
}
__device__ __forceinline__ float max3(float a, float b, float c) {
  float d;
  asm volatile("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
  return d;
}

template <int MODE>
__global__ void k(float* out, long long* cyc, int iters, float c, float d) {
  float a[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) a[i] = threadIdx.x * 1e-3f + i * 0.01f;
  uint32_t u[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) u[i] = threadIdx.x + i;
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    const float dd = d + float(it) * 1e-7f;   // loop-variant input so nothing can be hoisted
    if (MODE == 0) {          // 16 MUFU.EX2
#pragma unroll
      for (int i = 0; i < 16; ++i) a[i] = ex2(a[i]);
    } else if (MODE == 1) {   // 16 ex2.bf16x2
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        asm volatile("ex2.approx.ftz.bf16x2 %0, %0;" : "+r"(u[i]));
        asm volatile("ex2.approx.ftz.bf16x2 %0, %0;" : "+r"(u[i]));
      }
    } else if (MODE == 2) {   // 16 FFMA, 3 register operands
#pragma unroll
      for (int i = 0; i < 16; ++i) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a[i]) : "f"(c), "f"(d));
    } else if (MODE == 3) {   // 8 FFMA2, three 64-bit register operands
#pragma unroll
      for (int i = 0; i < 16; i += 2) FMA2(a[i], a[i + 1], a[i], a[i + 1], a[(i + 2) & 15], a[(i + 3) & 15], c, d);
    } else if (MODE == 4) {   // 8 FFMA2, packed a, broadcast scalars (the softmax scale step)
#pragma unroll
      for (int i = 0; i < 16; i += 2) FMA2(a[i], a[i + 1], a[i], a[i + 1], c, c, d, d);
    } else if (MODE == 5) {   // 8 FADD2
#pragma unroll
      for (int i = 0; i < 16; i += 2) F2("add.rn.f32x2", a[i], a[i + 1], a[i], a[i + 1], c, d);
    } else if (MODE == 6) {   // 8 FMNMX3
#pragma unroll
      for (int i = 0; i < 8; ++i) a[i] = max3(a[i], a[i + 8], c);
    } else if (MODE == 7) {   // 16 FMNMX
#pragma unroll
      for (int i = 0; i < 16; ++i) asm volatile("max.f32 %0, %0, %1;" : "+f"(a[i]) : "f"(c));
    } else if (MODE == 8) {   // 8 F2FP
#pragma unroll
      for (int i = 0; i < 8; ++i) u[i] = pack(a[2 * i], __uint_as_float(u[i]));
    } else if (MODE == 9) {   // 8 LEA (shift-add)
#pragma unroll
      for (int i = 0; i < 8; ++i) asm volatile("{\n\t.reg .b32 t;\n\tshl.b32 t, %1, 23;\n\tadd.s32 %0, %0, t;\n\t}" : "+r"(u[i]) : "r"(u[(i + 1) & 7]));
    } else if (MODE == 10) {  // softmax MUFU path for 16 elements: 8 FFMA2(scale) 16 MUFU 8 FADD2 8 F2FP
      float l0 = 0, l1 = 0;
#pragma unroll
      for (int i = 0; i < 16; i += 2) {
        float x0, x1;
        FMA2(x0, x1, a[i], a[i + 1], c, c, dd, dd);
        x0 = ex2(x0);
        x1 = ex2(x1);
        F2("add.rn.f32x2", l0, l1, l0, l1, x0, x1);
        u[i / 2] ^= pack(x0, x1);
      }
      a[0] += l0 * 1e-30f;
      a[1] += l1 * 1e-30f;
    } else if (MODE == 11) {  // softmax polynomial path for 16 elements
      float l0 = 0, l1 = 0;
#pragma unroll
      for (int i = 0; i < 16; i += 2) {
        float x0, x1, t0, t1, n0, n1, f0, f1, p0, p1;
        FMA2(x0, x1, a[i], a[i + 1], c, c, dd, dd);
        x0 = fmaxf(x0, -126.f);
        x1 = fmaxf(x1, -126.f);
        F2("add.rm.ftz.f32x2", t0, t1, x0, x1, 12582912.f, 12582912.f);
        F2("add.rn.ftz.f32x2", n0, n1, t0, t1, -12582912.f, -12582912.f);
        F2("sub.rn.ftz.f32x2", f0, f1, x0, x1, n0, n1);
        FMA2(p0, p1, f0, f1, 0.0771190897f, 0.0771190897f, 0.2275643945f, 0.2275643945f);
        FMA2(p0, p1, p0, p1, f0, f1, 0.6951461434f, 0.6951461434f);
        FMA2(p0, p1, p0, p1, f0, f1, 1.0f, 1.0f);
        x0 = __uint_as_float(__float_as_uint(p0) + (__float_as_uint(t0) << 23));
        x1 = __uint_as_float(__float_as_uint(p1) + (__float_as_uint(t1) << 23));
        F2("add.rn.f32x2", l0, l1, l0, l1, x0, x1);
        u[i / 2] ^= pack(x0, x1);
      }
      a[0] += l0 * 1e-30f;
      a[1] += l1 * 1e-30f;
    } else if (MODE == 12) {  // 16 elements: 12 MUFU + 4 polynomial (POLY16 = 4)
      float l0 = 0, l1 = 0;
#pragma unroll
      for (int i = 0; i < 16; i += 2) {
        float x0, x1;
        FMA2(x0, x1, a[i], a[i + 1], c, c, dd, dd);
        if (i < 4) {
          float t0, t1, n0, n1, f0, f1, p0, p1;
          x0 = fmaxf(x0, -126.f);
          x1 = fmaxf(x1, -126.f);
          F2("add.rm.ftz.f32x2", t0, t1, x0, x1, 12582912.f, 12582912.f);
          F2("add.rn.ftz.f32x2", n0, n1, t0, t1, -12582912.f, -12582912.f);
          F2("sub.rn.ftz.f32x2", f0, f1, x0, x1, n0, n1);
          FMA2(p0, p1, f0, f1, 0.0771190897f, 0.0771190897f, 0.2275643945f, 0.2275643945f);
          FMA2(p0, p1, p0, p1, f0, f1, 0.6951461434f, 0.6951461434f);
          FMA2(p0, p1, p0, p1, f0, f1, 1.0f, 1.0f);
          x0 = __uint_as_float(__float_as_uint(p0) + (__float_as_uint(t0) << 23));
          x1 = __uint_as_float(__float_as_uint(p1) + (__float_as_uint(t1) << 23));
        } else {
          x0 = ex2(x0);
          x1 = ex2(x1);
        }
        F2("add.rn.f32x2", l0, l1, l0, l1, x0, x1);
        u[i / 2] ^= pack(x0, x1);
      }
      a[0] += l0 * 1e-30f;
      a[1] += l1 * 1e-30f;
    } else if (MODE == 13) {  // 16 FMUL by scalar
#pragma unroll
      for (int i = 0; i < 16; ++i) asm volatile("mul.rn.f32 %0, %0, %1;" : "+f"(a[i]) : "f"(c));
    } else if (MODE == 14) {  // 16 FADD scalar
#pragma unroll
      for (int i = 0; i < 16; ++i) asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(a[i]) : "f"(c));
    }
  }
  long long t1 = clock64();
  float s = 0;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += a[i];
#pragma unroll
  for (int i = 0; i < 8; ++i) s += __uint_as_float(u[i]);
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int MODE>
void run(const char* name, int instr_per_iter, int elems_per_iter) {
  float* d;
  long long* c;
  cudaMalloc(&d, 148 * 1024 * 4);
  cudaMalloc(&c, 148 * 8);
  const int iters = 4000;
  printf("%-44s", name);
  for (int threads : {128, 256, 512}) {
    k<MODE><<<148, threads>>>(d, c, 10, 0.999f, -0.5f);
    k<MODE><<<148, threads>>>(d, c, iters, 0.999f, -0.5f);
    long long h[148];
    cudaMemcpy(h, c, sizeof(h), cudaMemcpyDeviceToHost);
    double avg = 0;
    for (int i = 0; i < 148; ++i) avg += h[i];
    avg /= 148;
    const int wps = threads / 128;   // warps per sub-partition
    // cycles of sub-partition time per warp-instruction, and per element
    printf("  w%d: %6.2f cyc/instr %6.2f cyc/elem", wps, avg / (double(iters) * instr_per_iter * wps),
           avg / (double(iters) * elems_per_iter * wps));
  }
  printf("  %s\n", cudaGetErrorString(cudaGetLastError()));
  cudaFree(d);
  cudaFree(c);
}

int main() {
  run<0>("MUFU.EX2 f32", 16, 16);
  run<1>("MUFU.EX2 bf16x2", 16, 32);
  run<2>("FFMA (3 reg)", 16, 16);
  run<3>("FFMA2 (3 x 64-bit reg)", 8, 16);
  run<4>("FFMA2 (pair, scalar, scalar)", 8, 16);
  run<5>("FADD2", 8, 16);
  run<6>("FMNMX3", 8, 16);
  run<7>("FMNMX", 16, 16);
  run<8>("F2FP.BF16 pack", 8, 16);
  run<9>("LEA shift-add", 8, 8);
  run<13>("FMUL", 16, 16);
  run<14>("FADD", 16, 16);
  run<10>("softmax MUFU path (40 instr / 16 elem)", 40, 16);
  run<11>("softmax poly path (120 instr / 16 elem)", 120, 16);
  run<12>("softmax 12 MUFU + 4 poly / 16 elem", 56, 16);
  return 0;
}
