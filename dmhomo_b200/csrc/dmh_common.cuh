// Shared host/device helpers for libdmhomo (sm_100a).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>
#include <cstdarg>
#include <cstdio>

#include "../../include/dmhomo.h"

namespace dmh {

// ---- host side -----------------------------------------------------------------------
extern thread_local char g_last_error[512];
extern std::atomic<uint64_t> g_launches;
extern thread_local const char* g_last_kernel;

inline int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_last_error, sizeof(g_last_error), fmt, ap);
  va_end(ap);
  return code;
}

// Call right after a <<<>>> launch: counts it and turns launch errors into DMH_ECUDA.
inline int launched(const char* what) {
  g_launches.fetch_add(1, std::memory_order_relaxed);
  g_last_kernel = what;
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail(DMH_ECUDA, "%s: %s", what, cudaGetErrorString(e));
  return DMH_OK;
}

#define DMH_REQUIRE(cond, ...) \
  do {                         \
    if (!(cond)) return ::dmh::fail(DMH_EINVAL, __VA_ARGS__); \
  } while (0)

// Development knobs (dmh_set_tuning): plain ints read at launch time.  Defaults are the measured best; none of them
// changes a result.  Not meant to be raced against launches from other threads.
struct Tuning {
  int tile = 3;           // 0: scalar kernels only; 1: TMA tile kernel for the dense C = 1 launches; 2: also the gradient-free C = 3 launches; 3: also the C = 3 training launch
  int tile_interior = 3;  // bit 0: interior-tile body, bit 1: mixed (per-row-pair vote) body
  int tile_flow = 1;      // explicit-flow C = 1 launches on the tile kernel (0: scalar kernels)
  int channels = 1;       // pixel-per-thread forward kernel for explicit-flow warps of feature maps with C other than 1 / 3 (0: channel groups)
  int tile_wide = 0;      // (-DDMH_TILE_WIDE variant builds only) gradient-free C = 1 launches on the 24-consumer-warp geometry
  int tile_pair_major = -1; // two-term launches walk (sample, term, tile) instead of (term, sample, tile); -1: at C = 3
  int tile_dyn = -1;      // share (percent) of the tile list handed out dynamically (guided: long runs first, single tiles last); -1: per launch kind
  int tile_chunk = -1;    // longest run of tiles per dynamic claim (1 .. 8); -1: per launch kind
};
Tuning& tuning();

inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }

constexpr int kNumSMs = 148;  // B200

// Elementwise launches over (B, h, w): blockIdx.y walks the samples, blockIdx.x * threads walks one plane, both
// grid-stride, all index arithmetic in 32 bits (a flat 64-bit index costs two 64-bit divisions per pixel, which
// is what these kernels then spend their time on: profiles/r1_kernels.txt).  Needs h * w < 2^31.
inline dim3 plane_grid(long long plane, int B, int threads = 256) {
  long long gx = (plane + threads - 1) / threads;
  if (gx > 4096) gx = 4096;
  if (gx < 1) gx = 1;
  long long gy = ((long long)kNumSMs * 16 + gx - 1) / gx;
  if (gy > B) gy = B;
  if (gy < 1) gy = 1;
  if (gy > 65535) gy = 65535;
  return dim3((unsigned)gx, (unsigned)gy, 1);
}
#define DMH_PLANE_LOOP(b, p, B_, plane_)                 \
  for (int b = blockIdx.y; b < (B_); b += gridDim.y)     \
    for (unsigned p = blockIdx.x * blockDim.x + threadIdx.x; p < (unsigned)(plane_); p += gridDim.x * blockDim.x)

// ---- device side ----------------------------------------------------------------------
// Separately rounded fp32 arithmetic: the reference is a chain of individually rounded
// ATen elementwise ops, so bit-exact coordinates / indices / masks need "no FMA
// contraction" spelled out.
__device__ __forceinline__ float mul_rn(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float add_rn(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float sub_rn(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ float div_rn(float a, float b) { return __fdiv_rn(a, b); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__device__ __forceinline__ float sign_of(float v) { return (v > 0.f) ? 1.f : ((v < 0.f) ? -1.f : 0.f); }

// fire-and-forget fp32 add (SASS: REDG.E.ADD.F32)
__device__ __forceinline__ void red_add(float* p, float v) { atomicAdd(p, v); }

}  // namespace dmh
