// What does one FFMA2 cost on sm_100a?  The fused warp kernel issues ~68 packed FMAs per row pair and sits at ~57 %
// issue-slot utilisation whatever the number of warps or the ILP per warp (profiles/r1_tile_experiments.txt), so the
// question for round 2 is whether FFMA2 with three distinct register-pair operands issues every cycle or every other
// cycle, and whether the scalar-broadcast operand forms (Rx.F32) or plain FADD2 / FMUL2 are cheaper.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/ubench_ffma2 tools/ubench_ffma2.cu && tools/ubench_ffma2
//
// Every mode runs N = 12 independent chains per thread (no dependency stalls at >= 2 warps per scheduler) and reports
// cycles per warp-instruction per SM sub-partition for 1, 2, 4 and 8 warps per sub-partition.
#include <cstdio>
#include <cuda_runtime.h>

typedef unsigned long long u64;
__device__ __forceinline__ u64 pk(float2 a) { return *reinterpret_cast<u64*>(&a); }
__device__ __forceinline__ float2 up(u64 a) { return *reinterpret_cast<float2*>(&a); }
__device__ __forceinline__ float2 fma2(float2 a, float2 b, float2 c) {
  u64 r;
  asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(pk(a)), "l"(pk(b)), "l"(pk(c)));
  return up(r);
}
__device__ __forceinline__ float2 add2(float2 a, float2 b) {
  u64 r;
  asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(pk(a)), "l"(pk(b)));
  return up(r);
}
__device__ __forceinline__ float2 mul2(float2 a, float2 b) {
  u64 r;
  asm volatile("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(pk(a)), "l"(pk(b)));
  return up(r);
}
__device__ __forceinline__ float fma1(float a, float b, float c) {
  float r;
  asm volatile("fma.rn.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
  return r;
}

constexpr int N = 12;
// MODE 0: FFMA2 a = fma2(a, b_i, c_i)      three distinct register pairs per instruction
//      1: FFMA2 a = fma2(a, splat(k), c_i) one scalar-broadcast operand (the kernel's MUL2 / ADD2 identities)
//      2: FFMA2 a = fma2(a, splat(k), splat(m))  two scalar-broadcast operands
//      3: FADD2 a = add2(a, b_i)
//      4: FMUL2 a = mul2(a, b_i)
//      5: scalar FFMA x 2 (a.x, a.y) with distinct operands
//      6: FFMA2 a = fma2(a, b_0, c_0)      every instruction reuses the same two operand pairs (reuse cache)
template <int MODE>
__global__ void k(float* out, int iters, float k0, float m0, long long* cycles) {
  float2 a[N], b[N], c[N];
#pragma unroll
  for (int i = 0; i < N; ++i) {
    a[i] = make_float2(threadIdx.x * 0.001f + i, threadIdx.x * 0.002f + i);
    // run-time values: immediates or folded constants would not read the register file
    b[i] = make_float2(k0 + m0 * (float)(i + threadIdx.x), k0 - m0 * (float)(i + 2 * threadIdx.x));
    c[i] = make_float2(m0 * (float)(i + 1) + m0 * threadIdx.x, -m0 * (float)(i + 1) - m0 * threadIdx.x);
  }
  const float2 ks = make_float2(k0, k0), ms = make_float2(m0, m0);
  const long long t0 = clock64();
#pragma unroll 8
  for (int it = 0; it < iters; ++it) {   // 96 instructions per trip: loop overhead below 3 %
#pragma unroll
    for (int i = 0; i < N; ++i) {
      if (MODE == 0) a[i] = fma2(a[i], b[i], c[i]);
      if (MODE == 1) a[i] = fma2(a[i], ks, c[i]);
      if (MODE == 2) a[i] = fma2(a[i], ks, ms);
      if (MODE == 3) a[i] = add2(a[i], b[i]);
      if (MODE == 4) a[i] = mul2(a[i], b[i]);
      if (MODE == 5) {
        a[i].x = fma1(a[i].x, b[i].x, c[i].x);
        a[i].y = fma1(a[i].y, b[i].y, c[i].y);
      }
      if (MODE == 6) a[i] = fma2(a[i], b[0], c[0]);
    }
  }
  const long long t1 = clock64();
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < N; ++i) s += a[i].x + a[i].y;
  if (s == 1.2345f) out[threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cycles = t1 - t0;
}

template <int MODE>
void run(const char* name, float* o, long long* cyc) {
  const int iters = 20000;
  printf("%-44s", name);
  for (int wps = 1; wps <= 8; wps *= 2) {      // warps per sub-partition: one CTA per SM of 4 * wps warps (8: only the modes with few registers launch)
    k<MODE><<<148, 128 * wps>>>(o, iters, 1.0000001f, 1e-6f, cyc);
    if (cudaDeviceSynchronize() != cudaSuccess || cudaGetLastError() != cudaSuccess) { printf("  %dw: launch failed", wps); continue; }
    long long c;
    cudaMemcpy(&c, cyc, sizeof(c), cudaMemcpyDeviceToHost);
    const double instr_per_smsp = (double)iters * N * wps * ((MODE == 5) ? 2 : 1);
    printf("  %dw: %.2f", wps, (double)c / instr_per_smsp);
  }
  printf("   cycles / warp-instruction / SMSP\n");
}

int main() {
  float* o;
  long long* cyc;
  cudaMalloc(&o, 1 << 16);
  cudaMalloc(&cyc, 8);
  run<0>("FFMA2 three distinct register pairs", o, cyc);
  run<1>("FFMA2 one scalar-broadcast operand", o, cyc);
  run<2>("FFMA2 two scalar-broadcast operands", o, cyc);
  run<6>("FFMA2 same two operand pairs (reuse)", o, cyc);
  run<3>("FADD2", o, cyc);
  run<4>("FMUL2", o, cyc);
  run<5>("scalar FFMA (per instruction, 2 per pair)", o, cyc);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) printf("CUDA error: %s\n", cudaGetErrorString(e));
  return 0;
}
