// Micro-benchmarks that decide the backward-scatter design (run on the B200 box):
//   REDG fp32 scalar / v2 / v4, coalesced rows with a 1-element shift (the bilinear tap pattern),
//   shared-memory float atomicAdd (CAS loop), shared int atomicAdd, plain shared RMW, plain stores.
#include <cstdio>
#include <cuda_runtime.h>
#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("ERR %s line %d\n", cudaGetErrorString(e), __LINE__); return 1;}}while(0)

// each thread: one pixel; T taps = REDG to (row*W + x + shift_t)
template <int TAPS>
__global__ void redg_scalar(float* g, int W, int H, long long plane_stride) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  const int y = blockIdx.y;
  float* p = g + (long long)blockIdx.z * plane_stride + (long long)y * W + x;
  if (x >= W - 2 || y >= H - 2) return;
  const float v = 1e-3f * (x & 7);
  if (TAPS >= 1) atomicAdd(p, v);
  if (TAPS >= 2) atomicAdd(p + 1, v);
  if (TAPS >= 3) atomicAdd(p + W, v);
  if (TAPS >= 4) atomicAdd(p + W + 1, v);
  if (TAPS >= 5) atomicAdd(p + 2 * W, v);
}
__global__ void redg_v4(float* g, int W, int H, long long plane_stride) {
  const int x = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
  const int y = blockIdx.y;
  float* p = g + (long long)blockIdx.z * plane_stride + (long long)y * W + x;
  if (x >= W || y >= H) return;
  const float v = 1e-3f * (x & 7);
  asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v), "f"(v), "f"(v), "f"(v) : "memory");
}
__global__ void redg_v2(float* g, int W, int H, long long plane_stride) {
  const int x = (blockIdx.x * blockDim.x + threadIdx.x) * 2;
  const int y = blockIdx.y;
  float* p = g + (long long)blockIdx.z * plane_stride + (long long)y * W + x;
  if (x >= W || y >= H) return;
  const float v = 1e-3f * (x & 7);
  asm volatile("red.global.add.v2.f32 [%0], {%1,%2};" ::"l"(p), "f"(v), "f"(v) : "memory");
}
__global__ void st_scalar(float* g, int W, int H, long long plane_stride) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  const int y = blockIdx.y;
  float* p = g + (long long)blockIdx.z * plane_stride + (long long)y * W + x;
  if (x >= W || y >= H) return;
  *p = 1e-3f * (x & 7);
}
__global__ void st_v4(float* g, int W, int H, long long plane_stride) {
  const int x = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
  const int y = blockIdx.y;
  float* p = g + (long long)blockIdx.z * plane_stride + (long long)y * W + x;
  if (x >= W || y >= H) return;
  const float v = 1e-3f * (x & 7);
  *reinterpret_cast<float4*>(p) = make_float4(v, v, v, v);
}

// shared memory: MODE 0 = float atomicAdd (CAS), 1 = int atomicAdd, 2 = plain RMW, 3 = red.shared via PTX
template <int MODE>
__global__ void smem_acc(float* out, int iters) {
  __shared__ float s[66 * 34];
  for (int i = threadIdx.x; i < 66 * 34; i += blockDim.x) s[i] = 0.f;
  __syncthreads();
  const int lane = threadIdx.x & 31, row = threadIdx.x >> 5;  // 256 threads: 8 rows
  for (int it = 0; it < iters; ++it) {
    const int yy = (row + 8 * (it & 3));
    const int a = yy * 66 + lane + ((it >> 2) & 1) * 32;
    const float v = 1e-3f * lane;
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      const int ad = a + (t & 1) + (t >> 1) * 66;
      if (MODE == 0) atomicAdd(&s[ad], v);
      else if (MODE == 1) atomicAdd(reinterpret_cast<int*>(&s[ad]), (int)(v * 1e6f));
      else if (MODE == 2) s[ad] += v;
      else { unsigned sa = (unsigned)__cvta_generic_to_shared(&s[ad]); asm volatile("red.shared.add.f32 [%0], %1;" ::"r"(sa), "f"(v) : "memory"); }
    }
  }
  __syncthreads();
  float acc = 0.f;
  for (int i = threadIdx.x; i < 66 * 34; i += blockDim.x) acc += s[i];
  if (acc == 123.456f) out[blockIdx.x] = acc;
}

int main() {
  const int W = 576, H = 320, B = 128;  // 94 MB buffer
  const long long plane = (long long)W * H;
  float* g;
  CK(cudaMalloc(&g, sizeof(float) * plane * B));
  CK(cudaMemset(g, 0, sizeof(float) * plane * B));
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  float ms;
  const double npx = (double)plane * B;
#define TIME(name, per_px, ...) do { for (int i = 0; i < 3; ++i) { __VA_ARGS__; } CK(cudaDeviceSynchronize()); cudaEventRecord(e0); for (int i = 0; i < 10; ++i) { __VA_ARGS__; } cudaEventRecord(e1); CK(cudaDeviceSynchronize()); cudaEventElapsedTime(&ms, e0, e1); ms /= 10; printf("%-28s %8.1f us  %7.1f Gpx/s  %7.1f G-elem-atomics/s\n", name, ms * 1e3, npx / ms / 1e6, npx * per_px / ms / 1e6); } while (0)
  dim3 g1((W + 127) / 128, H, B), b1(128);
  TIME("st scalar", 1, (st_scalar<<<g1, b1>>>(g, W, H, plane)));
  TIME("st v4", 1, (st_v4<<<dim3((W / 4 + 127) / 128, H, B), b1>>>(g, W, H, plane)));
  TIME("memset", 1, cudaMemsetAsync(g, 0, sizeof(float) * plane * B));
  TIME("redg scalar x1", 1, (redg_scalar<1><<<g1, b1>>>(g, W, H, plane)));
  TIME("redg scalar x2 (x,x+1)", 2, (redg_scalar<2><<<g1, b1>>>(g, W, H, plane)));
  TIME("redg scalar x4 (2x2)", 4, (redg_scalar<4><<<g1, b1>>>(g, W, H, plane)));
  TIME("redg scalar x5", 5, (redg_scalar<5><<<g1, b1>>>(g, W, H, plane)));
  TIME("redg v2", 1, (redg_v2<<<dim3((W / 2 + 127) / 128, H, B), b1>>>(g, W, H, plane)));
  TIME("redg v4", 1, (redg_v4<<<dim3((W / 4 + 127) / 128, H, B), b1>>>(g, W, H, plane)));
  float* o; CK(cudaMalloc(&o, 4 * 148 * 8));
  const int iters = 2000;
  const double nops = 148.0 * 8 * 256 * iters * 4;
#define TIMES(name, ...) do { __VA_ARGS__; CK(cudaDeviceSynchronize()); cudaEventRecord(e0); __VA_ARGS__; cudaEventRecord(e1); CK(cudaDeviceSynchronize()); cudaEventElapsedTime(&ms, e0, e1); printf("%-28s %8.1f us  %7.1f G-elem-ops/s (%.2f cyc/warp-op/SM @1.9GHz)\n", name, ms * 1e3, nops / ms / 1e6, ms * 1e-3 * 1.9e9 / (nops / 32 / 148)); } while (0)
  TIMES("smem float atomicAdd (CAS)", (smem_acc<0><<<148 * 8, 256>>>(o, iters)));
  TIMES("smem int atomicAdd", (smem_acc<1><<<148 * 8, 256>>>(o, iters)));
  TIMES("smem plain RMW", (smem_acc<2><<<148 * 8, 256>>>(o, iters)));
  TIMES("smem red.shared.add.f32", (smem_acc<3><<<148 * 8, 256>>>(o, iters)));
  return 0;
}
