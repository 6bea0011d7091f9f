// Elementwise / reduction kernels around the warp: homography -> flow (fp32 and fp64 variants),
// basis-flow combine, validity masks, plain L1, loss finish, flow -> RGB, evaluation point error.
#include "dmh_common.cuh"

namespace dmh {

constexpr int kThreads = 256;

static inline unsigned blocks_for(long long n, int per_block = kThreads) {
  long long b = (n + per_block - 1) / per_block;
  const long long cap = (long long)kNumSMs * 32;  // grid-stride beyond a few waves
  return (unsigned)(b < cap ? (b > 0 ? b : 1) : cap);
}

// ---- A5: get_flow (HEM/model/utils.py:400-440) -------------------------------------------------
__device__ __forceinline__ void h2flow(const float* hm, float gx, float gy, float& fx, float& fy, float& qX,
                                       float& qY, float& qT) {
  qX = add_rn(add_rn(mul_rn(hm[0], gx), mul_rn(hm[1], gy)), hm[2]);
  qY = add_rn(add_rn(mul_rn(hm[3], gx), mul_rn(hm[4], gy)), hm[5]);
  qT = add_rn(add_rn(mul_rn(hm[6], gx), mul_rn(hm[7], gy)), hm[8]);
  if (!(fabsf(qT) >= 1e-7f)) qT = add_rn(qT, 1e-6f);
  fx = sub_rn(div_rn(qX, qT), gx);
  fy = sub_rn(div_rn(qY, qT), gy);
}

// V pixels per thread (V = 4 when w % 4 == 0: the four pixels share a row, stores are 128-bit).
template <int V>
__global__ void __launch_bounds__(kThreads) h2flow_fwd_kernel(const float* __restrict__ H, float* __restrict__ flow,
                                                              int B, int h, int w, int dv, float sx, float sy,
                                                              const float* __restrict__ start) {
  const long long plane = (long long)h * w;
  const unsigned groups = (unsigned)(plane / V);
  for (int b = blockIdx.y; b < B; b += gridDim.y) {
    float hm[9];
    if (dv == 1) {   // one homography per sample: loaded once per sample, not once per pixel
#pragma unroll
      for (int k = 0; k < 9; ++k) hm[k] = __ldg(H + (size_t)b * 9 + k);
    }
    const float ox = start ? __ldg(start + 2 * b) : sx, oy = start ? __ldg(start + 2 * b + 1) : sy;
    for (unsigned q = blockIdx.x * blockDim.x + threadIdx.x; q < groups; q += gridDim.x * blockDim.x) {
      const unsigned p = q * V;
      const int y = (int)(p / (unsigned)w), x0 = (int)(p - (unsigned)y * (unsigned)w);
      float fxv[V], fyv[V];
#pragma unroll
      for (int v = 0; v < V; ++v) {
        const int x = x0 + v;
        if (dv != 1) {
          const int cell = min(y / (h / dv), dv - 1) * dv + min(x / (w / dv), dv - 1);
          const float* hp = H + ((size_t)b * dv * dv + cell) * 9;
#pragma unroll
          for (int k = 0; k < 9; ++k) hm[k] = __ldg(hp + k);
        }
        float qX, qY, qT;
        h2flow(hm, add_rn((float)x, ox), add_rn((float)y, oy), fxv[v], fyv[v], qX, qY, qT);
      }
      float* o = flow + ((size_t)b * 2) * plane + p;
      if (V == 4) {
        *reinterpret_cast<float4*>(o) = make_float4(fxv[0], fxv[1], fxv[2], fxv[3]);
        *reinterpret_cast<float4*>(o + plane) = make_float4(fyv[0], fyv[1], fyv[2], fyv[3]);
      } else {
        o[0] = fxv[0];
        o[plane] = fyv[0];
      }
    }
  }
}

// grad_H[b,cell] += sum_px (gX*(gx,gy,1), gY*(gx,gy,1), gT*(gx,gy,1)); one CTA = 8 rows x 256 px of a cell row
__global__ void __launch_bounds__(kThreads) h2flow_bwd_kernel(const float* __restrict__ H,
                                                              const float* __restrict__ gflow, float* __restrict__ gH,
                                                              int B, int h, int w, int dv, float sx, float sy,
                                                              const float* __restrict__ start) {
  // grid: (chunks of a cell's pixels, cell, b)
  const int b = blockIdx.z, cell = blockIdx.y;
  if (start) {
    sx = __ldg(start + 2 * b);
    sy = __ldg(start + 2 * b + 1);
  }
  const int ch = h / dv, cw = w / dv;
  const int cy0 = (cell / dv) * ch, cx0 = (cell % dv) * cw;
  const int cell_h = (cell / dv == dv - 1) ? h - cy0 : ch, cell_w = (cell % dv == dv - 1) ? w - cx0 : cw;
  const long long npx = (long long)cell_h * cell_w, plane = (long long)h * w;
  const float* hp = H + ((size_t)b * dv * dv + cell) * 9;
  float hm[9];
#pragma unroll
  for (int k = 0; k < 9; ++k) hm[k] = __ldg(hp + k);
  float acc[9];
#pragma unroll
  for (int k = 0; k < 9; ++k) acc[k] = 0.f;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < npx; i += (long long)gridDim.x * blockDim.x) {
    const int yy = (int)(i / cell_w), xx = (int)(i - (long long)yy * cell_w);
    const int y = cy0 + yy, x = cx0 + xx;
    const float gx = add_rn((float)x, sx), gy = add_rn((float)y, sy);
    float fx, fy, qX, qY, qT;
    h2flow(hm, gx, gy, fx, fy, qX, qY, qT);
    const long long p = (long long)y * w + x;
    const float gfx = __ldg(gflow + ((size_t)b * 2) * plane + p), gfy = __ldg(gflow + ((size_t)b * 2 + 1) * plane + p);
    const float rT = 1.f / qT;
    const float gX = gfx * rT, gY = gfy * rT, gT = -(gfx * qX + gfy * qY) * rT * rT;
    acc[0] += gX * gx; acc[1] += gX * gy; acc[2] += gX;
    acc[3] += gY * gx; acc[4] += gY * gy; acc[5] += gY;
    acc[6] += gT * gx; acc[7] += gT * gy; acc[8] += gT;
  }
  __shared__ float red[kThreads / 32][9];
  const int lane = threadIdx.x & 31, wrp = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < 9; ++k) {
    const float v = warp_sum(acc[k]);
    if (lane == 0) red[wrp][k] = v;
  }
  __syncthreads();
  if (threadIdx.x < 9) {
    float v = 0.f;
#pragma unroll
    for (int q = 0; q < kThreads / 32; ++q) v += red[q][threadIdx.x];
    red_add(gH + ((size_t)b * dv * dv + cell) * 9 + threadIdx.x, v);
  }
}

// ---- A15: numpy fp64 homography -> flow / mapping (ddpm.py:913-975; ...operations.py:454-484) ----
template <int V>
__global__ void __launch_bounds__(kThreads) h2flow_f64_kernel(const double* __restrict__ H, float* __restrict__ out,
                                                              int B, int h, int w, double eps, int channels_last,
                                                              int as_mapping) {
  const long long plane = (long long)h * w;
  const unsigned groups = (unsigned)(plane / V);
  for (int b = blockIdx.y; b < B; b += gridDim.y) {
    double hm[9];
#pragma unroll
    for (int k = 0; k < 9; ++k) hm[k] = H[(size_t)b * 9 + k];
    for (unsigned q = blockIdx.x * blockDim.x + threadIdx.x; q < groups; q += gridDim.x * blockDim.x) {
      const unsigned p = q * V;
      const int y = (int)(p / (unsigned)w), x0 = (int)(p - (unsigned)y * (unsigned)w);
      float oxv[V], oyv[V];
#pragma unroll
      for (int v = 0; v < V; ++v) {
        const int x = x0 + v;
        const double dx = (double)x, dy = (double)y;
        const double X = __dadd_rn(__dadd_rn(__dmul_rn(hm[0], dx), __dmul_rn(hm[1], dy)), hm[2]);
        const double Y = __dadd_rn(__dadd_rn(__dmul_rn(hm[3], dx), __dmul_rn(hm[4], dy)), hm[5]);
        const double T = __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(hm[6], dx), __dmul_rn(hm[7], dy)), hm[8]), eps);
        double ox = __ddiv_rn(X, T), oy = __ddiv_rn(Y, T);
        if (as_mapping == 0) {
          ox = __dsub_rn(ox, dx);
          oy = __dsub_rn(oy, dy);
        } else if (as_mapping == 2) {
          // homo_convert_to_flow (HEM/dataset/data_loader.py:42-52): the mapping is rounded to fp32 first
          // (map_x.astype(np.float32)), then convert_mapping_to_flow subtracts the fp32 grid in fp32
          ox = (double)__fsub_rn((float)ox, (float)x);
          oy = (double)__fsub_rn((float)oy, (float)y);
        }
        oxv[v] = (float)ox;
        oyv[v] = (float)oy;
      }
      if (channels_last) {
        float2* o = reinterpret_cast<float2*>(out) + (size_t)b * plane + p;
        if (V == 4) {
          *reinterpret_cast<float4*>(o) = make_float4(oxv[0], oyv[0], oxv[1], oyv[1]);
          *reinterpret_cast<float4*>(o + 2) = make_float4(oxv[V - 2], oyv[V - 2], oxv[V - 1], oyv[V - 1]);
        } else {
          o[0] = make_float2(oxv[0], oyv[0]);
        }
      } else {
        float* o = out + ((size_t)b * 2) * plane + p;
        if (V == 4) {
          *reinterpret_cast<float4*>(o) = make_float4(oxv[0], oxv[1], oxv[V - 2], oxv[V - 1]);
          *reinterpret_cast<float4*>(o + plane) = make_float4(oyv[0], oyv[1], oyv[V - 2], oyv[V - 1]);
        } else {
          o[0] = oxv[0];
          o[plane] = oyv[0];
        }
      }
    }
  }
}

static inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

// ---- A12: basis combine (HEM/model/net.py:808-815) ------------------------------------------------
// One thread = one pixel; the 16 basis values stay in registers while the thread loops over the
// batch, so the shared basis tensor is read once per launch instead of once per sample.
__global__ void __launch_bounds__(kThreads) basis_combine_kernel(const float* __restrict__ basis,
                                                                 const float* __restrict__ weight,
                                                                 float* __restrict__ flow, int B, int h, int w,
                                                                 int b_per_block) {
  const long long plane = (long long)h * w;
  const long long p = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (p >= plane) return;
  float bx[8], by[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    bx[k] = __ldg(basis + (size_t)(2 * k) * plane + p);
    by[k] = __ldg(basis + (size_t)(2 * k + 1) * plane + p);
  }
  const int b0 = blockIdx.y * b_per_block, b1 = min(B, b0 + b_per_block);
  for (int b = b0; b < b1; ++b) {
    const float* wp = weight + (size_t)b * 8;
    float fx = mul_rn(bx[0], __ldg(wp)), fy = mul_rn(by[0], __ldg(wp));
#pragma unroll
    for (int k = 1; k < 8; ++k) {
      const float wk = __ldg(wp + k);
      fx = add_rn(fx, mul_rn(bx[k], wk));
      fy = add_rn(fy, mul_rn(by[k], wk));
    }
    flow[((size_t)b * 2) * plane + p] = fx;
    flow[((size_t)b * 2 + 1) * plane + p] = fy;
  }
}

// Four pixels per thread (plane % 4 == 0): 128-bit loads and stores, same per-pixel operation order.
__global__ void __launch_bounds__(kThreads, 3) basis_combine4_kernel(const float* __restrict__ basis,
                                                                     const float* __restrict__ weight,
                                                                     float* __restrict__ flow, int B, long long plane4,
                                                                     int b_per_block) {
  const long long q = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (q >= plane4) return;
  const float4* bs = reinterpret_cast<const float4*>(basis);
  float4* out = reinterpret_cast<float4*>(flow);
  float4 bv[16];
#pragma unroll
  for (int k = 0; k < 16; ++k) bv[k] = __ldg(bs + (size_t)k * plane4 + q);
  const int b0 = blockIdx.y * b_per_block, b1 = min(B, b0 + b_per_block);
  for (int b = b0; b < b1; ++b) {
    const float* wp = weight + (size_t)b * 8;
    const float w0 = __ldg(wp);
    float4 fx = make_float4(mul_rn(bv[0].x, w0), mul_rn(bv[0].y, w0), mul_rn(bv[0].z, w0), mul_rn(bv[0].w, w0));
    float4 fy = make_float4(mul_rn(bv[1].x, w0), mul_rn(bv[1].y, w0), mul_rn(bv[1].z, w0), mul_rn(bv[1].w, w0));
#pragma unroll
    for (int k = 1; k < 8; ++k) {
      const float wk = __ldg(wp + k);
      const float4 cx = bv[2 * k], cy = bv[2 * k + 1];
      fx.x = add_rn(fx.x, mul_rn(cx.x, wk)); fx.y = add_rn(fx.y, mul_rn(cx.y, wk));
      fx.z = add_rn(fx.z, mul_rn(cx.z, wk)); fx.w = add_rn(fx.w, mul_rn(cx.w, wk));
      fy.x = add_rn(fy.x, mul_rn(cy.x, wk)); fy.y = add_rn(fy.y, mul_rn(cy.y, wk));
      fy.z = add_rn(fy.z, mul_rn(cy.z, wk)); fy.w = add_rn(fy.w, mul_rn(cy.w, wk));
    }
    out[((size_t)b * 2) * plane4 + q] = fx;
    out[((size_t)b * 2 + 1) * plane4 + q] = fy;
  }
}

// dL/dw[b,k] = sum_px ( gflow[b,0,px] * basis[k,0,px] + gflow[b,1,px] * basis[k,1,px] ): a skinny (B x 2hw) x (2hw x 8) product.
// General form (any plane size): one sample per CTA row, the basis re-read per sample from the L2.
__global__ void __launch_bounds__(kThreads) basis_combine_bwd_kernel(const float* __restrict__ basis,
                                                                     const float* __restrict__ gflow,
                                                                     float* __restrict__ gw, int B, int h, int w) {
  const int b = blockIdx.y;
  const long long plane = (long long)h * w;
  float acc[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) acc[k] = 0.f;
  for (long long p = blockIdx.x * (long long)blockDim.x + threadIdx.x; p < plane; p += (long long)gridDim.x * blockDim.x) {
    const float gx = __ldg(gflow + ((size_t)b * 2) * plane + p), gy = __ldg(gflow + ((size_t)b * 2 + 1) * plane + p);
#pragma unroll
    for (int k = 0; k < 8; ++k)
      acc[k] += gx * __ldg(basis + (size_t)(2 * k) * plane + p) + gy * __ldg(basis + (size_t)(2 * k + 1) * plane + p);
  }
  __shared__ float red[kThreads / 32][8];
  const int lane = threadIdx.x & 31, wrp = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const float v = warp_sum(acc[k]);
    if (lane == 0) red[wrp][k] = v;
  }
  __syncthreads();
  if (threadIdx.x < 8) {
    float v = 0.f;
#pragma unroll
    for (int q = 0; q < kThreads / 32; ++q) v += red[q][threadIdx.x];
    red_add(gw + (size_t)b * 8 + threadIdx.x, v);
  }
}

// plane % 4 == 0: a thread keeps the 16 basis values of its four pixels in registers (64) and walks kBwdSamples
// samples, two 128-bit loads of the upstream gradient each; the eight partial sums of a sample are reduced across the
// warp at once by a transposing butterfly (4 + 2 + 1 + 2 shuffles, lane l ends up with the total of one k) instead of
// eight separate warp sums. The basis is read B / kBwdSamples times from the L2 (per-sample kernel above: B times, 755
// MB of L2 reads and 82 us at cfg2's size; holding 8 samples x 8 sums per thread instead: 168 registers, 186 us).
constexpr int kBwdSamples = 8;
__device__ __forceinline__ float dot4(const float4 a, const float4 b, float acc) {
  return fmaf(a.w, b.w, fmaf(a.z, b.z, fmaf(a.y, b.y, fmaf(a.x, b.x, acc))));
}
__global__ void __launch_bounds__(kThreads, 2) basis_combine_bwd4_kernel(const float* __restrict__ basis,
                                                                         const float* __restrict__ gflow,
                                                                         float* __restrict__ gw, int B, long long plane4) {
  const long long q = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  const bool valid = q < plane4;
  const float4* bs = reinterpret_cast<const float4*>(basis);
  const float4* gs = reinterpret_cast<const float4*>(gflow);
  const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
  float4 bv[16];
#pragma unroll
  for (int k = 0; k < 16; ++k) bv[k] = valid ? __ldg(bs + (size_t)k * plane4 + q) : zero;
  const int b0 = blockIdx.y * kBwdSamples;
  const int lane = threadIdx.x & 31, wrp = threadIdx.x >> 5;
  const bool o1 = lane & 1, o2 = lane & 2, o4 = lane & 4;
  float tot[kBwdSamples];
#pragma unroll
  for (int s = 0; s < kBwdSamples; ++s) {
    const bool have = valid && b0 + s < B;
    const float4 gx = have ? __ldg(gs + ((size_t)(b0 + s) * 2) * plane4 + q) : zero;
    const float4 gy = have ? __ldg(gs + ((size_t)(b0 + s) * 2 + 1) * plane4 + q) : zero;
    float v[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) v[k] = dot4(gy, bv[2 * k + 1], dot4(gx, bv[2 * k], 0.f));
    float r4[4], r2[2];
#pragma unroll
    for (int i = 0; i < 4; ++i) r4[i] = (o1 ? v[i + 4] : v[i]) + __shfl_xor_sync(0xffffffffu, o1 ? v[i] : v[i + 4], 1);
#pragma unroll
    for (int i = 0; i < 2; ++i) r2[i] = (o2 ? r4[i + 2] : r4[i]) + __shfl_xor_sync(0xffffffffu, o2 ? r4[i] : r4[i + 2], 2);
    float r = (o4 ? r2[1] : r2[0]) + __shfl_xor_sync(0xffffffffu, o4 ? r2[0] : r2[1], 4);
    r += __shfl_xor_sync(0xffffffffu, r, 8);
    r += __shfl_xor_sync(0xffffffffu, r, 16);
    tot[s] = r;   // the warp's sum for k = 4*bit0 + 2*bit1 + bit2 of the lane
  }
  __shared__ float red[kThreads / 32][kBwdSamples * 8];
  if (lane < 8) {
    const int k = ((lane & 1) << 2) | (lane & 2) | ((lane >> 2) & 1);
#pragma unroll
    for (int s = 0; s < kBwdSamples; ++s) red[wrp][s * 8 + k] = tot[s];
  }
  __syncthreads();
  if (threadIdx.x < kBwdSamples * 8 && b0 + (int)(threadIdx.x >> 3) < B) {
    float v = 0.f;
#pragma unroll
    for (int w8 = 0; w8 < kThreads / 32; ++w8) v += red[w8][threadIdx.x];
    red_add(gw + (size_t)(b0 + (threadIdx.x >> 3)) * 8 + (threadIdx.x & 7), v);
  }
}

// ---- A10/A11: masks (HEM/utils_operations/flow_and_mapping_operations.py:6-71) -----------------------
template <int V>
__global__ void __launch_bounds__(kThreads) border_mask_kernel(const float* __restrict__ flow, uint8_t* __restrict__ m8,
                                                               float* __restrict__ mf, int B, int h, int w) {
  const long long plane = (long long)h * w;
  const unsigned groups = (unsigned)(plane / V);
  DMH_PLANE_LOOP(b, q, B, groups) {
    const unsigned p = q * V;
    const long long i = (long long)b * plane + p;
    const int y = (int)(p / (unsigned)w), x0 = (int)(p - (unsigned)y * (unsigned)w);
    float fx[V], fy[V];
    const float* f = flow + ((size_t)b * 2) * plane + p;
    if (V == 4) {
      const float4 a4 = __ldg(reinterpret_cast<const float4*>(f)), b4 = __ldg(reinterpret_cast<const float4*>(f + plane));
      fx[0] = a4.x; fx[1] = a4.y; fx[V - 2] = a4.z; fx[V - 1] = a4.w;
      fy[0] = b4.x; fy[1] = b4.y; fy[V - 2] = b4.z; fy[V - 1] = b4.w;
    } else {
      fx[0] = __ldg(f);
      fy[0] = __ldg(f + plane);
    }
    uint8_t ok[V];
#pragma unroll
    for (int v = 0; v < V; ++v) {
      const float mx = add_rn(fx[v], (float)(x0 + v)), my = add_rn(fy[v], (float)y);
      ok[v] = ((mx >= 0.f) && (mx <= (float)w) && (my >= 0.f) && (my <= (float)h)) ? 1 : 0;
    }
    if (V == 4) {
      if (m8) *reinterpret_cast<uchar4*>(m8 + i) = make_uchar4(ok[0], ok[1], ok[V - 2], ok[V - 1]);
      if (mf) *reinterpret_cast<float4*>(mf + i) = make_float4((float)ok[0], (float)ok[1], (float)ok[V - 2], (float)ok[V - 1]);
    } else {
      if (m8) m8[i] = ok[0];
      if (mf) mf[i] = (float)ok[0];
    }
  }
}

__global__ void __launch_bounds__(kThreads) zero_border_mask_kernel(const float* __restrict__ img,
                                                                    uint8_t* __restrict__ mask, int B, int h, int w,
                                                                    float eps) {
  const long long plane = (long long)h * w;
  DMH_PLANE_LOOP(b, p, B, plane) {
    const long long i = (long long)b * plane + p;   // flat pixel index (channels-last / per-sample-plane layouts)
    const float* ip = img + (size_t)b * 3 * plane + p;
    const bool occ = (__ldg(ip) <= eps) && (__ldg(ip + plane) <= eps) && (__ldg(ip + 2 * plane) <= eps);
    mask[i] = occ ? 0 : 1;
  }
}

// ---- A13: LossL1 (HEM/loss/losses.py:10-17) ---------------------------------------------------------
// 128-bit loads (n4 = number of float4 groups when both pointers are 16-byte aligned, else 0) + scalar tail
__global__ void __launch_bounds__(kThreads) l1_sum_kernel(const float* __restrict__ a, const float* __restrict__ b,
                                                          long long n, long long n4, double* __restrict__ acc) {
  float s = 0.f;
  double sd = 0.0;
  int cnt = 0;
  const long long stride = (long long)gridDim.x * blockDim.x, t0 = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  for (long long i = t0; i < n4; i += stride) {
    const float4 x = __ldg(reinterpret_cast<const float4*>(a) + i), y = __ldg(reinterpret_cast<const float4*>(b) + i);
    s += (fabsf(x.x - y.x) + fabsf(x.y - y.y)) + (fabsf(x.z - y.z) + fabsf(x.w - y.w));
    if (++cnt == 16) {  // bound the fp32 partial (64 terms)
      sd += (double)s;
      s = 0.f;
      cnt = 0;
    }
  }
  for (long long i = 4 * n4 + t0; i < n; i += stride) s += fabsf(__ldg(a + i) - __ldg(b + i));
  sd += (double)s;
  sd = warp_sum(sd);
  __shared__ double red[kThreads / 32];
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = sd;
  __syncthreads();
  if (threadIdx.x == 0) {
    double v = 0.0;
#pragma unroll
    for (int q = 0; q < kThreads / 32; ++q) v += red[q];
    atomicAdd(acc, v);
  }
}

__global__ void __launch_bounds__(kThreads) l1_bwd_kernel(const float* __restrict__ a, const float* __restrict__ b,
                                                          long long n, long long n4, const float* __restrict__ g, float scale,
                                                          float* __restrict__ ga, float* __restrict__ gb) {
  const float gs = (g ? __ldg(g) : 1.f) * scale;
  const long long stride = (long long)gridDim.x * blockDim.x, t0 = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  for (long long i = t0; i < n4; i += stride) {
    const float4 x = __ldg(reinterpret_cast<const float4*>(a) + i), y = __ldg(reinterpret_cast<const float4*>(b) + i);
    const float4 v = make_float4(gs * sign_of(x.x - y.x), gs * sign_of(x.y - y.y), gs * sign_of(x.z - y.z), gs * sign_of(x.w - y.w));
    if (ga) reinterpret_cast<float4*>(ga)[i] = v;
    if (gb) reinterpret_cast<float4*>(gb)[i] = make_float4(-v.x, -v.y, -v.z, -v.w);
  }
  for (long long i = 4 * n4 + t0; i < n; i += stride) {
    const float v = gs * sign_of(__ldg(a + i) - __ldg(b + i));
    if (ga) ga[i] = v;
    if (gb) gb[i] = -v;
  }
}

struct FinishArgs {
  const double* acc[8];
  const float* sw[8];
};

__global__ void loss_finish_kernel(FinishArgs args, int n_acc, int B, float scale, float* __restrict__ loss) {
  double s = 0.0;
  for (int i = threadIdx.x; i < n_acc * B; i += blockDim.x) {
    const int a = i / B, b = i - a * B;
    const double wgt = args.sw[a] ? (double)__ldg(args.sw[a] + b) : 1.0;
    s += wgt * args.acc[a][b];
  }
  s = warp_sum(s);
  __shared__ double red[kThreads / 32];
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    double v = 0.0;
    for (int q = 0; q < (int)(blockDim.x >> 5); ++q) v += red[q];
    loss[0] = (float)(v * (double)scale);
  }
}

__global__ void __launch_bounds__(kThreads) scale_inplace_kernel(float* __restrict__ x, long long n,
                                                                 const float* __restrict__ g) {
  const float gs = __ldg(g);
  if (gs == 1.0f) return;  // unit upstream gradient: the forward-computed gradient is already final
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    x[i] *= gs;
}

// ---- A16: flow_to_image (ddpm.py:1471-1486) + hsv_to_rgb (App. A.6) ---------------------------------
__global__ void __launch_bounds__(kThreads) flow_to_rgb_kernel(const float* __restrict__ flow, float* __restrict__ rgb,
                                                               int B, int h, int w, float max_flow, int in_cl,
                                                               int out_cl) {
  const long long plane = (long long)h * w;
  const float two_pi = 6.2831855f;  // float32(2*np.pi)
  DMH_PLANE_LOOP(b, p, B, plane) {
    const long long i = (long long)b * plane + p;   // flat pixel index (channels-last / per-sample-plane layouts)
    float u, v;
    if (in_cl) {
      const float2 f = __ldg(reinterpret_cast<const float2*>(flow) + i);
      u = f.x;
      v = f.y;
    } else {
      u = __ldg(flow + ((size_t)b * 2) * plane + p);
      v = __ldg(flow + ((size_t)b * 2 + 1) * plane + p);
    }
    const float mag = __fsqrt_rn(add_rn(mul_rn(u, u), mul_rn(v, v)));
    const float ang = atan2f(v, u);
    float hh = add_rn(div_rn(ang, two_pi), 1.0f);
    hh = hh - floorf(hh);                        // np.mod(., 1) for a value in [0.5, 1.5]
    const float s = fminf(fmaxf(div_rn(mul_rn(mag, 8.0f), max_flow), 0.f), 1.f);
    const float val = fminf(fmaxf(sub_rn(8.0f, s), 0.f), 1.f);
    // hsv_to_rgb
    const float h6 = mul_rn(hh, 6.0f);
    const int sec = (int)h6;
    const float f = sub_rn(h6, (float)sec);
    const float pp = mul_rn(val, sub_rn(1.f, s));
    const float qq = mul_rn(val, sub_rn(1.f, mul_rn(s, f)));
    const float tt = mul_rn(val, sub_rn(1.f, mul_rn(s, sub_rn(1.f, f))));
    float r, g, bl;
    switch (sec % 6) {
      case 0: r = val; g = tt; bl = pp; break;
      case 1: r = qq; g = val; bl = pp; break;
      case 2: r = pp; g = val; bl = tt; break;
      case 3: r = pp; g = qq; bl = val; break;
      case 4: r = tt; g = pp; bl = val; break;
      default: r = val; g = pp; bl = qq; break;
    }
    if (s == 0.f) r = g = bl = val;
    if (out_cl) {
      float* o = rgb + (size_t)i * 3;
      o[0] = r; o[1] = g; o[2] = bl;
    } else {
      float* o = rgb + (size_t)b * 3 * plane + p;
      o[0] = r; o[plane] = g; o[2 * plane] = bl;
    }
  }
}

// ---- A18: ComputeErrFlow / compute_eval_results (HEM/loss/losses.py:208-211, 263-296) ------------------
__global__ void eval_point_error_kernel(const float* __restrict__ pts, const float* __restrict__ ff,
                                        const float* __restrict__ fb, float* __restrict__ err, int B, int P, int h,
                                        int w) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  float tot = 0.f;
  for (int j = 0; j < P; ++j) {
    const float* q = pts + ((size_t)b * P + j) * 4;
    const float p1x = q[0], p1y = q[1], p2x = q[2], p2y = q[3];
    float e[2];
#pragma unroll
    for (int dir = 0; dir < 2; ++dir) {
      if (dir == 1 && fb == nullptr) {  // single-direction query (ComputeErrFlow)
        e[1] = e[0];
        break;
      }
      const float sx_ = dir ? p2x : p1x, sy_ = dir ? p2y : p1y, dx_ = dir ? p1x : p2x, dy_ = dir ? p1y : p2y;
      const float* fl = dir ? fb : ff;
      int iy = (int)sy_, ix = (int)sx_;          // int() truncation (losses.py:209)
      iy = iy < 0 ? iy + h : iy;                 // python negative indexing
      ix = ix < 0 ? ix + w : ix;
      iy = min(max(iy, 0), h - 1);
      ix = min(max(ix, 0), w - 1);
      const float* f = fl + (((size_t)b * h + iy) * w + ix) * 2;
      const float rx = sub_rn(dx_, add_rn(sx_, f[0])), ry = sub_rn(dy_, add_rn(sy_, f[1]));
      e[dir] = __fsqrt_rn(add_rn(mul_rn(rx, rx), mul_rn(ry, ry)));
    }
    tot = add_rn(tot, fminf(e[0], e[1]));
  }
  err[b] = div_rn(tot, (float)P);
}

}  // namespace dmh

using namespace dmh;

extern "C" int dmh_homography_to_flow(const float* H, float* flow, int B, int h, int w, int divide, float start_x,
                                      float start_y, const float* start, void* stream) {
  DMH_REQUIRE(H && flow, "homography_to_flow: null pointer");
  DMH_REQUIRE(B > 0 && h > 0 && w > 0 && divide >= 1, "homography_to_flow: bad size");
  DMH_REQUIRE(h % divide == 0 && w % divide == 0, "homography_to_flow: h,w must be divisible by divide");
  if ((w & 3) == 0 && (reinterpret_cast<uintptr_t>(flow) & 15) == 0)
    h2flow_fwd_kernel<4><<<plane_grid((long long)h * w / 4, B), kThreads, 0, as_stream(stream)>>>(H, flow, B, h, w, divide,
                                                                                                start_x, start_y, start);
  else
    h2flow_fwd_kernel<1><<<plane_grid((long long)h * w, B), kThreads, 0, as_stream(stream)>>>(H, flow, B, h, w, divide,
                                                                                            start_x, start_y, start);
  return launched("h2flow_fwd_kernel");
}

extern "C" int dmh_homography_to_flow_backward(const float* H, const float* grad_flow, float* grad_H, int B, int h,
                                               int w, int divide, float start_x, float start_y, const float* start,
                                               void* stream) {
  DMH_REQUIRE(H && grad_flow && grad_H, "homography_to_flow_backward: null pointer");
  DMH_REQUIRE(B > 0 && h > 0 && w > 0 && divide >= 1, "homography_to_flow_backward: bad size");
  DMH_REQUIRE(h % divide == 0 && w % divide == 0, "homography_to_flow_backward: h,w must be divisible by divide");
  DMH_REQUIRE(B <= 65535 && divide * divide <= 65535, "homography_to_flow_backward: B or divide^2 > 65535");
  const long long npx = (long long)(h / divide) * (w / divide);
  long long chunks = (npx + kThreads * 8 - 1) / (kThreads * 8);
  if (chunks < 1) chunks = 1;
  dim3 grid((unsigned)chunks, (unsigned)(divide * divide), (unsigned)B);
  h2flow_bwd_kernel<<<grid, kThreads, 0, as_stream(stream)>>>(H, grad_flow, grad_H, B, h, w, divide, start_x, start_y,
                                                              start);
  return launched("h2flow_bwd_kernel");
}

extern "C" int dmh_homography_to_flow_f64(const double* H, float* out, int B, int h, int w, double eps,
                                          int channels_last, int as_mapping, void* stream) {
  DMH_REQUIRE(H && out, "homography_to_flow_f64: null pointer");
  DMH_REQUIRE(B > 0 && h > 0 && w > 0, "homography_to_flow_f64: bad size");
  if ((w & 3) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0)
    h2flow_f64_kernel<4><<<plane_grid((long long)h * w / 4, B), kThreads, 0, as_stream(stream)>>>(H, out, B, h, w, eps,
                                                                                                channels_last, as_mapping);
  else
    h2flow_f64_kernel<1><<<plane_grid((long long)h * w, B), kThreads, 0, as_stream(stream)>>>(H, out, B, h, w, eps,
                                                                                            channels_last, as_mapping);
  return launched("h2flow_f64_kernel");
}

extern "C" int dmh_basis_combine(const float* basis, const float* weight, float* flow, int B, int h, int w,
                                 void* stream) {
  DMH_REQUIRE(basis && weight && flow, "basis_combine: null pointer");
  DMH_REQUIRE(B > 0 && h > 0 && w > 0, "basis_combine: bad size");
  const long long plane = (long long)h * w;
  const bool vec = plane % 4 == 0 && aligned16(basis) && aligned16(flow);
  const long long items = vec ? plane / 4 : plane;
  const unsigned bx = (unsigned)((items + kThreads - 1) / kThreads);
  // enough CTAs for a few waves, but as many samples per CTA as possible (basis reuse)
  int by = (int)(((vec ? 4LL : 32LL) * kNumSMs + bx - 1) / bx);
  by = by < 1 ? 1 : (by > B ? B : by);
  const int b_per_block = (B + by - 1) / by;
  by = (B + b_per_block - 1) / b_per_block;
  if (vec)
    basis_combine4_kernel<<<dim3(bx, (unsigned)by), kThreads, 0, as_stream(stream)>>>(basis, weight, flow, B, items,
                                                                                      b_per_block);
  else
    basis_combine_kernel<<<dim3(bx, (unsigned)by), kThreads, 0, as_stream(stream)>>>(basis, weight, flow, B, h, w,
                                                                                     b_per_block);
  return launched("basis_combine_kernel");
}

extern "C" int dmh_basis_combine_backward(const float* basis, const float* grad_flow, float* grad_weight, int B, int h,
                                          int w, void* stream) {
  DMH_REQUIRE(basis && grad_flow && grad_weight, "basis_combine_backward: null pointer");
  DMH_REQUIRE(B > 0 && B <= 65535 && h > 0 && w > 0, "basis_combine_backward: bad size");
  const long long plane = (long long)h * w;
  if (plane % 4 == 0 && aligned16(basis) && aligned16(grad_flow)) {
    const long long plane4 = plane / 4;
    const unsigned gx = (unsigned)((plane4 + kThreads - 1) / kThreads);
    const unsigned gy = (unsigned)((B + kBwdSamples - 1) / kBwdSamples);
    basis_combine_bwd4_kernel<<<dim3(gx, gy), kThreads, 0, as_stream(stream)>>>(basis, grad_flow, grad_weight, B, plane4);
  } else {
    long long chunks = (plane + kThreads * 8 - 1) / (kThreads * 8);
    basis_combine_bwd_kernel<<<dim3((unsigned)chunks, (unsigned)B), kThreads, 0, as_stream(stream)>>>(
        basis, grad_flow, grad_weight, B, h, w);
  }
  return launched("basis_combine_bwd_kernel");
}

extern "C" int dmh_border_mask(const float* flow, uint8_t* mask_u8, float* mask_f32, int B, int h, int w,
                               void* stream) {
  DMH_REQUIRE(flow && (mask_u8 || mask_f32), "border_mask: null pointer");
  DMH_REQUIRE(B > 0 && h > 0 && w > 0, "border_mask: bad size");
  const bool vec = (w & 3) == 0 && (reinterpret_cast<uintptr_t>(flow) & 15) == 0 && (reinterpret_cast<uintptr_t>(mask_u8) & 3) == 0 &&
                   (reinterpret_cast<uintptr_t>(mask_f32) & 15) == 0;
  if (vec)
    border_mask_kernel<4><<<plane_grid((long long)h * w / 4, B), kThreads, 0, as_stream(stream)>>>(flow, mask_u8, mask_f32, B, h, w);
  else
    border_mask_kernel<1><<<plane_grid((long long)h * w, B), kThreads, 0, as_stream(stream)>>>(flow, mask_u8, mask_f32, B, h, w);
  return launched("border_mask_kernel");
}

extern "C" int dmh_zero_border_mask(const float* image, uint8_t* mask, int B, int h, int w, float eps, void* stream) {
  DMH_REQUIRE(image && mask, "zero_border_mask: null pointer");
  DMH_REQUIRE(B > 0 && h > 0 && w > 0, "zero_border_mask: bad size");
  zero_border_mask_kernel<<<plane_grid((long long)h * w, B), kThreads, 0, as_stream(stream)>>>(image, mask, B, h, w,
                                                                                               eps);
  return launched("zero_border_mask_kernel");
}

extern "C" int dmh_l1_sum(const float* a, const float* b, int64_t n, double* acc, void* stream) {
  DMH_REQUIRE(a && b && acc, "l1_sum: null pointer");
  DMH_REQUIRE(n > 0, "l1_sum: n must be positive");
  const bool vec = ((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(b)) & 15) == 0;
  l1_sum_kernel<<<blocks_for(n, kThreads * 16), kThreads, 0, as_stream(stream)>>>(a, b, n, vec ? n / 4 : 0, acc);
  return launched("l1_sum_kernel");
}

extern "C" int dmh_l1_backward(const float* a, const float* b, int64_t n, const float* g, float scale, float* ga,
                               float* gb, void* stream) {
  DMH_REQUIRE(a && b && (ga || gb), "l1_backward: null pointer");
  DMH_REQUIRE(n > 0, "l1_backward: n must be positive");
  const bool vec = ((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(b) | reinterpret_cast<uintptr_t>(ga) | reinterpret_cast<uintptr_t>(gb)) & 15) == 0;
  l1_bwd_kernel<<<blocks_for(n, kThreads * 4), kThreads, 0, as_stream(stream)>>>(a, b, n, vec ? n / 4 : 0, g, scale, ga, gb);
  return launched("l1_bwd_kernel");
}

extern "C" int dmh_loss_finish(const double* const* acc, const float* const* sample_weight, int n_acc, int B,
                               float scale, float* loss, void* stream) {
  DMH_REQUIRE(acc && loss, "loss_finish: null pointer");
  DMH_REQUIRE(n_acc > 0 && n_acc <= 8 && B > 0, "loss_finish: n_acc must be in 1..8, B positive");
  FinishArgs args;
  for (int i = 0; i < 8; ++i) {
    args.acc[i] = (i < n_acc) ? acc[i] : nullptr;
    args.sw[i] = (i < n_acc && sample_weight) ? sample_weight[i] : nullptr;
    if (i < n_acc) DMH_REQUIRE(acc[i] != nullptr, "loss_finish: acc[%d] is null", i);
  }
  loss_finish_kernel<<<1, kThreads, 0, as_stream(stream)>>>(args, n_acc, B, scale, loss);
  return launched("loss_finish_kernel");
}

extern "C" int dmh_scale_inplace(float* x, int64_t n, const float* g, void* stream) {
  DMH_REQUIRE(x && g, "scale_inplace: null pointer");
  DMH_REQUIRE(n > 0, "scale_inplace: n must be positive");
  // grid-stride over a bounded grid: with a unit upstream gradient (the usual case) every block returns at once,
  // and a grid of n / threads blocks would spend microseconds just being scheduled
  const long long nb = blocks_for(n);
  scale_inplace_kernel<<<(unsigned)(nb < 8 * kNumSMs ? nb : 8 * kNumSMs), kThreads, 0, as_stream(stream)>>>(x, n, g);
  return launched("scale_inplace_kernel");
}

extern "C" int dmh_flow_to_rgb(const float* flow, float* rgb, int B, int h, int w, float max_flow,
                               int in_channels_last, int out_channels_last, void* stream) {
  DMH_REQUIRE(flow && rgb, "flow_to_rgb: null pointer");
  DMH_REQUIRE(B > 0 && h > 0 && w > 0, "flow_to_rgb: bad size");
  DMH_REQUIRE(max_flow > 0.f, "flow_to_rgb: max_flow must be positive");
  flow_to_rgb_kernel<<<plane_grid((long long)h * w, B), kThreads, 0, as_stream(stream)>>>(
      flow, rgb, B, h, w, max_flow, in_channels_last, out_channels_last);
  return launched("flow_to_rgb_kernel");
}

extern "C" int dmh_eval_point_error(const float* pts, const float* flow_f, const float* flow_b, float* err, int B,
                                    int P, int h, int w, void* stream) {
  DMH_REQUIRE(pts && flow_f && err, "eval_point_error: null pointer");
  DMH_REQUIRE(B > 0 && P > 0 && h > 0 && w > 0, "eval_point_error: bad size");
  eval_point_error_kernel<<<(B + 63) / 64, 64, 0, as_stream(stream)>>>(pts, flow_f, flow_b, err, B, P, h, w);
  return launched("eval_point_error_kernel");
}
