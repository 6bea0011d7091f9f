// REDG / store throughput with a persistent grid (no CTA-launch bound): every thread walks rows.
#include <cstdio>
#include <cuda_runtime.h>
#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("ERR %s line %d\n", cudaGetErrorString(e), __LINE__); return 1;}}while(0)

// MODE 0: plain store 1/px; 1: REDG 1/px; 2: REDG 2x2 taps (4/px) shifted by (dx,dy);
// 3: REDG 4 taps + 1 dense (5/px); 4: v4 REDG 1 per 4px; 5: load only (4 taps)
template <int MODE>
__global__ void k(float* __restrict__ g, const float* __restrict__ src, int W, int H, int B, int dx, int dy, float* sink) {
  const long long plane = (long long)W * H;
  const int rows = H * B;
  float acc = 0.f;
  for (int r = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); r < rows; r += gridDim.x * (blockDim.x >> 5)) {
    const int b = r / H, y = r - b * H;
    float* gp = g + b * plane;
    const float* sp = src + b * plane;
    for (int x0 = (threadIdx.x & 31); x0 < W; x0 += 32) {
      const int x = x0;
      const float v = 1e-3f * (x & 7);
      int xs = min(max(x + dx, 0), W - 2), ys = min(max(y + dy, 0), H - 2);
      if (MODE == 0) gp[y * W + x] = v;
      if (MODE == 1) atomicAdd(gp + y * W + x, v);
      if (MODE == 2 || MODE == 3) {
        float* q = gp + ys * W + xs;
        atomicAdd(q, v); atomicAdd(q + 1, v); atomicAdd(q + W, v); atomicAdd(q + W + 1, v);
        if (MODE == 3) atomicAdd(gp + y * W + x, v);
      }
      if (MODE == 5) {
        const float* q = sp + ys * W + xs;
        acc += __ldg(q) + __ldg(q + 1) + __ldg(q + W) + __ldg(q + W + 1);
      }
    }
    if (MODE == 4) {
      for (int x = (threadIdx.x & 31) * 4; x < W; x += 128) {
        float* q = gp + y * W + x;
        const float v = 1e-3f;
        asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(q), "f"(v), "f"(v), "f"(v), "f"(v) : "memory");
      }
    }
  }
  if (acc == 123.f) *sink = acc;
}

int main() {
  const int W = 576, H = 320;
  for (int B : {32, 128, 512}) {
    const long long plane = (long long)W * H;
    float *g, *s, *sink;
    CK(cudaMalloc(&g, sizeof(float) * plane * B));
    CK(cudaMalloc(&s, sizeof(float) * plane * B));
    CK(cudaMalloc(&sink, 4));
    CK(cudaMemset(g, 0, sizeof(float) * plane * B));
    CK(cudaMemset(s, 0, sizeof(float) * plane * B));
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    float ms;
    const double npx = (double)plane * B;
    printf("---- B=%d (%.0f MB)\n", B, npx * 4 / 1e6);
#define TIME(name, per_px, ...) do { for (int i = 0; i < 3; ++i) { __VA_ARGS__; } CK(cudaDeviceSynchronize()); cudaEventRecord(e0); for (int i = 0; i < 10; ++i) { __VA_ARGS__; } cudaEventRecord(e1); CK(cudaDeviceSynchronize()); cudaEventElapsedTime(&ms, e0, e1); ms /= 10; printf("%-34s %8.1f us  %7.1f Gpx/s  %7.1f G-elem/s\n", name, ms * 1e3, npx / ms / 1e6, npx * per_px / ms / 1e6); } while (0)
    const int grid = 148 * 8, blk = 256;
    TIME("memset", 1, cudaMemsetAsync(g, 0, sizeof(float) * plane * B));
    TIME("store 1/px", 1, (k<0><<<grid, blk>>>(g, s, W, H, B, 0, 0, sink)));
    TIME("ld 4 taps", 4, (k<5><<<grid, blk>>>(g, s, W, H, B, 3, 2, sink)));
    TIME("redg 1/px", 1, (k<1><<<grid, blk>>>(g, s, W, H, B, 0, 0, sink)));
    TIME("redg 2x2 taps", 4, (k<2><<<grid, blk>>>(g, s, W, H, B, 3, 2, sink)));
    TIME("redg 2x2 + dense (5/px)", 5, (k<3><<<grid, blk>>>(g, s, W, H, B, 3, 2, sink)));
    TIME("redg v4 (1 elem/px)", 1, (k<4><<<grid, blk>>>(g, s, W, H, B, 0, 0, sink)));
    TIME("memset+redg 5/px", 5, (cudaMemsetAsync(g, 0, sizeof(float) * plane * B), k<3><<<grid, blk>>>(g, s, W, H, B, 3, 2, sink)));
    cudaFree(g); cudaFree(s); cudaFree(sink);
  }
  return 0;
}
