// Does packing two fp32 ops into FADD2/FMUL2/FFMA2 (sm_100) free issue slots?  Mixed FP + INT stream per thread.
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE>  // 0: scalar fp only, 1: packed fp only, 2: scalar fp + int, 3: packed fp + int
__global__ void k(float* out, int iters, float c0, int ci) {
  float2 a[4];
  int z[8];
  for (int i = 0; i < 4; ++i) a[i] = make_float2(threadIdx.x * 0.5f + i, threadIdx.x * 0.25f + i);
  for (int i = 0; i < 8; ++i) z[i] = threadIdx.x + i;
  const float2 c = make_float2(c0, c0 * 1.5f);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int r = 0; r < 4; ++r) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        if (MODE == 0 || MODE == 2) {
          a[i].x = __fadd_rn(__fmul_rn(a[i].x, c.x), c.y);
          a[i].y = __fadd_rn(__fmul_rn(a[i].y, c.y), c.x);
        } else {
          a[i] = __fadd2_rn(__fmul2_rn(a[i], c), make_float2(c.y, c.x));
        }
      }
      if (MODE >= 2) {
#pragma unroll
        for (int i = 0; i < 8; ++i) z[i] = (z[i] ^ ci) + (z[(i + 1) & 7] >> 1);
      }
    }
  }
  float s = 0.f;
  for (int i = 0; i < 4; ++i) s += a[i].x + a[i].y;
  int zz = 0;
  for (int i = 0; i < 8; ++i) zz += z[i];
  if (s == 1.2345f || zz == 12345) out[threadIdx.x] = s + zz;
}
int main() {
  float* o; cudaMalloc(&o, 4096);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int iters = 4000;
  const char* names[4] = {"scalar fp (32 ops/iter)", "packed fp (16 ops/iter)", "scalar fp + 64 int", "packed fp + 64 int"};
  for (int m = 0; m < 4; ++m) {
    for (int rep = 0; rep < 2; ++rep) {
      cudaEventRecord(e0);
      if (m == 0) k<0><<<148 * 4, 256>>>(o, iters, 1.0001f, 3);
      if (m == 1) k<1><<<148 * 4, 256>>>(o, iters, 1.0001f, 3);
      if (m == 2) k<2><<<148 * 4, 256>>>(o, iters, 1.0001f, 3);
      if (m == 3) k<3><<<148 * 4, 256>>>(o, iters, 1.0001f, 3);
      cudaEventRecord(e1); cudaDeviceSynchronize();
    }
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    // per SM: 4 CTAs x 8 warps = 32 warps; cycles per warp-iteration per SMSP
    printf("%-28s %8.3f ms   %.1f cycles/iter/SMSP (8 warps per SMSP)\n", names[m], ms, ms * 1e-3 * 1.9e9 / iters);
  }
  return 0;
}
