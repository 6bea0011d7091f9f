// 4-point DLT: one warp per 8x8 system, the augmented matrix lives in registers
// (lane r < 8 owns row r), partial pivoting and back-substitution through warp shuffles.
//
// Replaces DLT.forward(method='Axb') / WarpMat / DLT_solve:
//   HEM/model/utils.py:55-101, 19-43, 360-397; HEM/model/net.py:24-92.
// The reference forms inverse(A) @ b in fp32 (LU); its own error against the exact solution
// of its fp32 system is ~1e-7 (Frobenius-relative).  We build A with the reference's fp32
// roundings (entries -(u*x) rounded once), then eliminate in fp64 so that the only
// difference to the reference is the reference's own rounding noise.
#include "dmh_common.cuh"

namespace dmh {

// Solves M z = rhs for the 8x8 system whose row r = lane & 7 is m[0..7] | m[8] (the four 8-lane groups of the warp
// hold identical copies and run in lockstep).  Returns z_k broadcast in sol[k] on every lane.
// Gauss-Jordan with partial pivoting among the rows not yet used as a pivot: every other row is eliminated at each
// step, so there is no serial back-substitution - the latency chain is 8 x (3 shuffle rounds + one fp64 division +
// one FMA) and a final division, which is what this launch costs (it is latency-bound: 128 systems, 9 KB of data).
__device__ __forceinline__ void solve8_warp(double (&m)[9], int lane, double (&sol)[8]) {
  const int r = lane & 7;
  bool used = false;
  int mycol = 0;
  int piv[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    // pivot search: largest |m[k]| among unused rows, lowest row wins ties
    double best = used ? -1.0 : fabs(m[k]);
    int who = r;
#pragma unroll
    for (int o = 4; o > 0; o >>= 1) {
      const double ob = __shfl_xor_sync(0xffffffffu, best, o, 8);
      const int ow = __shfl_xor_sync(0xffffffffu, who, o, 8);
      if (ob > best || (ob == best && ow < who)) {
        best = ob;
        who = ow;
      }
    }
    piv[k] = who;
    const double pk = __shfl_sync(0xffffffffu, m[k], who, 8);
    const bool is_piv = (r == who);
    const double f = is_piv ? 0.0 : m[k] / pk;
#pragma unroll
    for (int j = k + 1; j < 9; ++j) {
      const double pj = __shfl_sync(0xffffffffu, m[j], who, 8);
      m[j] = fma(-f, pj, m[j]);
    }
    if (is_piv) {
      used = true;
      mycol = k;
    }
  }
  // row r pivoted column mycol: x_mycol = m[8] / m[mycol] (static register indexing)
  double d = m[0];
#pragma unroll
  for (int q = 1; q < 8; ++q)
    if (mycol == q) d = m[q];
  const double xr = m[8] / d;
#pragma unroll
  for (int k = 0; k < 8; ++k) sol[k] = __shfl_sync(0xffffffffu, xr, piv[k], 8);
}

// Row `r` (0..7) of the DLT system for point i = r/2 (App. A.1).
__device__ __forceinline__ void dlt_row(int r, float x, float y, float u, float v, double (&m)[9]) {
  const bool top = (r & 1) == 0;
  const float t = top ? u : v;
  m[0] = top ? x : 0.f;
  m[1] = top ? y : 0.f;
  m[2] = top ? 1.f : 0.f;
  m[3] = top ? 0.f : x;
  m[4] = top ? 0.f : y;
  m[5] = top ? 0.f : 1.f;
  m[6] = (double)(-mul_rn(t, x));
  m[7] = (double)(-mul_rn(t, y));
  m[8] = t;
}

// Element (row j, column c) of the same system with a run-time column (the adjoint solves need column r of A as
// their row): selects instead of an indexed local array.
__device__ __forceinline__ double dlt_elem(int j, int c, float x, float y, float u, float v) {
  const bool top = (j & 1) == 0;
  const float t = top ? u : v;
  const int cc = top ? c : c - 3;
  float val = 0.f;
  if (cc == 0) val = x;
  if (cc == 1) val = y;
  if (cc == 2) val = 1.f;
  if (c == 6) val = -mul_rn(t, x);
  if (c == 7) val = -mul_rn(t, y);
  return (double)val;
}

__global__ void __launch_bounds__(128) dlt4_fwd_kernel(const float* __restrict__ src, const float* __restrict__ dst,
                                                       float* __restrict__ H, int N) {
  const int n = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (n >= N) return;  // warp-uniform
  const int lane = threadIdx.x & 31;
  const int r = lane & 7, i = r >> 1;
  const float x = __ldg(src + (size_t)n * 8 + 2 * i), y = __ldg(src + (size_t)n * 8 + 2 * i + 1);
  const float u = __ldg(dst + (size_t)n * 8 + 2 * i), v = __ldg(dst + (size_t)n * 8 + 2 * i + 1);
  double m[9], sol[8];
  dlt_row(r, x, y, u, v, m);
  solve8_warp(m, lane, sol);
  float hv = 1.0f;                       // lane 8 writes h22 = 1
#pragma unroll
  for (int q = 0; q < 8; ++q)            // static register indexing
    if (lane == q) hv = (float)sol[q];
  if (lane < 9) H[(size_t)n * 9 + lane] = hv;
}

// h = A^-1 b.  With z = A^-T g_h:
//   g_u_i = z[2i] (1 + x_i h6 + y_i h7),  g_v_i = z[2i+1] (1 + x_i h6 + y_i h7)
//   g_x_i = -z[2i] (h0 - u_i h6) - z[2i+1] (h3 - v_i h6),  g_y_i = -z[2i] (h1 - u_i h7) - z[2i+1] (h4 - v_i h7)
__global__ void __launch_bounds__(128) dlt4_bwd_kernel(const float* __restrict__ src, const float* __restrict__ dst,
                                                       const float* __restrict__ H, const float* __restrict__ gH,
                                                       float* __restrict__ g_dst, float* __restrict__ g_src, int N) {
  const int n = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (n >= N) return;
  const int lane = threadIdx.x & 31;
  const int r = lane & 7;
  // row r of A^T = column r of A
  double m[9];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int i = j >> 1;
    const float x = __ldg(src + (size_t)n * 8 + 2 * i), y = __ldg(src + (size_t)n * 8 + 2 * i + 1);
    const float u = __ldg(dst + (size_t)n * 8 + 2 * i), v = __ldg(dst + (size_t)n * 8 + 2 * i + 1);
    m[j] = dlt_elem(j, r, x, y, u, v);
  }
  m[8] = (double)__ldg(gH + (size_t)n * 9 + r);
  double z[8];
  solve8_warp(m, lane, z);
  if (lane < 4) {
    const int i = lane;
    const double x = __ldg(src + (size_t)n * 8 + 2 * i), y = __ldg(src + (size_t)n * 8 + 2 * i + 1);
    const double u = __ldg(dst + (size_t)n * 8 + 2 * i), v = __ldg(dst + (size_t)n * 8 + 2 * i + 1);
    const double h0 = __ldg(H + (size_t)n * 9 + 0), h1 = __ldg(H + (size_t)n * 9 + 1);
    const double h3 = __ldg(H + (size_t)n * 9 + 3), h4 = __ldg(H + (size_t)n * 9 + 4);
    const double h6 = __ldg(H + (size_t)n * 9 + 6), h7 = __ldg(H + (size_t)n * 9 + 7);
    double zt = 0.0, zb = 0.0;
#pragma unroll
    for (int q = 0; q < 4; ++q) {  // static register indexing
      if (q == i) {
        zt = z[2 * q];
        zb = z[2 * q + 1];
      }
    }
    const double s = 1.0 + x * h6 + y * h7;
    g_dst[(size_t)n * 8 + 2 * i] = (float)(zt * s);
    g_dst[(size_t)n * 8 + 2 * i + 1] = (float)(zb * s);
    if (g_src) {
      g_src[(size_t)n * 8 + 2 * i] = (float)(-zt * (h0 - u * h6) - zb * (h3 - v * h6));
      g_src[(size_t)n * 8 + 2 * i + 1] = (float)(-zt * (h1 - u * h7) - zb * (h4 - v * h7));
    }
  }
}

// ---- basis flow at the four image corners (cfg 2 "DLT step") ---------------------------------
// offsets[b, corner, xy] = sum_k basis[k, xy, corner] * w[b,k], sequential, separately rounded.
__global__ void basis_corner_fwd_kernel(const float* __restrict__ basis, const float* __restrict__ weight,
                                        float* __restrict__ off, int B, int h, int w) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= B * 8) return;
  const int b = t >> 3, c = (t >> 1) & 3, xy = t & 1;
  const int py = (c >> 1) ? h - 1 : 0, px = (c & 1) ? w - 1 : 0;
  const size_t plane = (size_t)h * w, po = (size_t)py * w + px;
  float acc = mul_rn(__ldg(basis + (size_t)xy * plane + po), __ldg(weight + b * 8));
  for (int k = 1; k < 8; ++k)
    acc = add_rn(acc, mul_rn(__ldg(basis + (size_t)(2 * k + xy) * plane + po), __ldg(weight + b * 8 + k)));
  off[t] = acc;
}

__global__ void basis_corner_bwd_kernel(const float* __restrict__ basis, const float* __restrict__ g_off,
                                        float* __restrict__ g_w, int B, int h, int w) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= B * 8) return;
  const int b = t >> 3, k = t & 7;
  const size_t plane = (size_t)h * w;
  float acc = 0.f;
  for (int c = 0; c < 4; ++c) {
    const int py = (c >> 1) ? h - 1 : 0, px = (c & 1) ? w - 1 : 0;
    const size_t po = (size_t)py * w + px;
    acc += __ldg(g_off + b * 8 + c * 2) * __ldg(basis + (size_t)(2 * k) * plane + po);
    acc += __ldg(g_off + b * 8 + c * 2 + 1) * __ldg(basis + (size_t)(2 * k + 1) * plane + po);
  }
  red_add(g_w + t, acc);
}


// ---- fused: 8 basis weights -> corner offsets -> 4-point DLT (cfg 2 prologue), one warp per sample ----
// Lane j < 8 owns offset j = (corner j/2, xy j&1) exactly as basis_corner_fwd_kernel computes it (sequential,
// separately rounded), then the warp assembles and solves the 8x8 system like dlt4_fwd_kernel.  Up to four
// weight sets (e.g. forward / backward direction) share one launch.
struct BasisHArgs {
  const float* weight[4];
  const float* grad_H[4];
  float* H[4];
  float* grad_weight[4];
};

__device__ __forceinline__ void corner_xy(int c, int h, int w, float& x, float& y) {
  x = (c & 1) ? (float)(w - 1) : 0.f;
  y = (c >> 1) ? (float)(h - 1) : 0.f;
}

__global__ void __launch_bounds__(128) basis_h_fwd_kernel(const float* __restrict__ basis, const __grid_constant__ BasisHArgs args, int n_sets,
                                                          int B, int h, int w) {
  const int n = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (n >= n_sets * B) return;  // warp-uniform
  const int set = n / B, b = n - set * B;
  const int lane = threadIdx.x & 31, j = lane & 7;
  const float* wp = args.weight[set] + (size_t)b * 8;
  const int c = j >> 1, xy = j & 1;
  const int py = (c >> 1) ? h - 1 : 0, px = (c & 1) ? w - 1 : 0;
  const size_t plane = (size_t)h * w, po = (size_t)py * w + px;
  float bv[8], wv[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) {      // all 16 loads in flight at once
    bv[k] = __ldg(basis + (size_t)(2 * k + xy) * plane + po);
    wv[k] = __ldg(wp + k);
  }
  float off = mul_rn(bv[0], wv[0]);
#pragma unroll
  for (int k = 1; k < 8; ++k) off = add_rn(off, mul_rn(bv[k], wv[k]));
  // row r = lane & 7 of the system needs (x, y, u, v) of point i = r / 2: u, v = corner + offset
  const int i = j >> 1;
  float x, y;
  corner_xy(i, h, w, x, y);
  const float ox = __shfl_sync(0xffffffffu, off, 2 * i), oy = __shfl_sync(0xffffffffu, off, 2 * i + 1);
  const float u = add_rn(x, ox), v = add_rn(y, oy);
  double m[9], sol[8];
  dlt_row(j, x, y, u, v, m);
  solve8_warp(m, lane, sol);
  float* H = args.H[set] + (size_t)b * 9;
  float hv = 1.0f;                       // lane 8 writes h22 = 1
#pragma unroll
  for (int q = 0; q < 8; ++q)            // static register indexing
    if (lane == q) hv = (float)sol[q];
  if (lane < 9) H[lane] = hv;
}

// grad_H -> grad_weight (written): adjoint DLT solve (as dlt4_bwd_kernel), then the transpose of the corner
// sampling: g_w[k] = sum_{corner, xy} g_off[corner, xy] * basis[k, xy, corner].
__global__ void __launch_bounds__(128) basis_h_bwd_kernel(const float* __restrict__ basis, const __grid_constant__ BasisHArgs args, int n_sets,
                                                          int B, int h, int w) {
  const int n = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (n >= n_sets * B) return;
  const int set = n / B, b = n - set * B;
  const int lane = threadIdx.x & 31, r = lane & 7;
  const float* H = args.H[set] + (size_t)b * 9;
  // destination points u, v = corner + offset are recovered from H itself: dst_i = H * src_i (exact up to the
  // solve's rounding, far below the gradient tolerance)
  float xs[4], ys[4], us[4], vs[4];
  const float h0 = __ldg(H), h1 = __ldg(H + 1), h2 = __ldg(H + 2), h3 = __ldg(H + 3), h4 = __ldg(H + 4), h5 = __ldg(H + 5),
              h6 = __ldg(H + 6), h7 = __ldg(H + 7);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    corner_xy(i, h, w, xs[i], ys[i]);
    const float T = h6 * xs[i] + h7 * ys[i] + 1.0f;
    us[i] = (h0 * xs[i] + h1 * ys[i] + h2) / T;
    vs[i] = (h3 * xs[i] + h4 * ys[i] + h5) / T;
  }
  double m[9];
#pragma unroll
  for (int jj = 0; jj < 8; ++jj) {
    m[jj] = dlt_elem(jj, r, xs[jj >> 1], ys[jj >> 1], us[jj >> 1], vs[jj >> 1]);
  }
  m[8] = (double)__ldg(args.grad_H[set] + (size_t)b * 9 + r);
  double z[8];
  solve8_warp(m, lane, z);
  // g_off[2i] = z[2i] * s_i, g_off[2i+1] = z[2i+1] * s_i, s_i = 1 + x_i h6 + y_i h7
  float goff[8];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const double s = 1.0 + (double)xs[i] * h6 + (double)ys[i] * h7;
    goff[2 * i] = (float)(z[2 * i] * s);
    goff[2 * i + 1] = (float)(z[2 * i + 1] * s);
  }
  if (lane < 8) {
    const int k = lane;
    const size_t plane = (size_t)h * w;
    float acc = 0.f;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const int py = (c >> 1) ? h - 1 : 0, px = (c & 1) ? w - 1 : 0;
      const size_t po = (size_t)py * w + px;
      acc += goff[2 * c] * __ldg(basis + (size_t)(2 * k) * plane + po);
      acc += goff[2 * c + 1] * __ldg(basis + (size_t)(2 * k + 1) * plane + po);
    }
    args.grad_weight[set][(size_t)b * 8 + k] = acc;
  }
}

}  // namespace dmh

extern "C" int dmh_dlt4_forward(const float* src, const float* dst, float* H, int N, void* stream) {
  DMH_REQUIRE(src && dst && H, "dlt4_forward: null pointer");
  DMH_REQUIRE(N > 0, "dlt4_forward: N must be positive");
  dmh::dlt4_fwd_kernel<<<(N + 3) / 4, 128, 0, dmh::as_stream(stream)>>>(src, dst, H, N);
  return dmh::launched("dlt4_fwd_kernel");
}

extern "C" int dmh_dlt4_backward(const float* src, const float* dst, const float* H, const float* grad_H,
                                 float* grad_dst, float* grad_src, int N, void* stream) {
  DMH_REQUIRE(src && dst && H && grad_H && grad_dst, "dlt4_backward: null pointer");
  DMH_REQUIRE(N > 0, "dlt4_backward: N must be positive");
  dmh::dlt4_bwd_kernel<<<(N + 3) / 4, 128, 0, dmh::as_stream(stream)>>>(src, dst, H, grad_H, grad_dst, grad_src, N);
  return dmh::launched("dlt4_bwd_kernel");
}

extern "C" int dmh_basis_corner_offsets(const float* basis, const float* weight, float* offsets, int B, int h, int w,
                                        void* stream) {
  DMH_REQUIRE(basis && weight && offsets, "basis_corner_offsets: null pointer");
  DMH_REQUIRE(B > 0 && h > 0 && w > 0, "basis_corner_offsets: non-positive size");
  dmh::basis_corner_fwd_kernel<<<(B * 8 + 127) / 128, 128, 0, dmh::as_stream(stream)>>>(basis, weight, offsets, B, h, w);
  return dmh::launched("basis_corner_fwd_kernel");
}

extern "C" int dmh_basis_corner_offsets_backward(const float* basis, const float* grad_offsets, float* grad_weight,
                                                 int B, int h, int w, void* stream) {
  DMH_REQUIRE(basis && grad_offsets && grad_weight, "basis_corner_offsets_backward: null pointer");
  DMH_REQUIRE(B > 0 && h > 0 && w > 0, "basis_corner_offsets_backward: non-positive size");
  dmh::basis_corner_bwd_kernel<<<(B * 8 + 127) / 128, 128, 0, dmh::as_stream(stream)>>>(basis, grad_offsets,
                                                                                      grad_weight, B, h, w);
  return dmh::launched("basis_corner_bwd_kernel");
}

extern "C" int dmh_basis_homography_forward(const float* basis, const float* const* weights, float* const* H, int n_sets,
                                            int B, int h, int w, void* stream) {
  DMH_REQUIRE(basis && weights && H, "basis_homography_forward: null pointer");
  DMH_REQUIRE(n_sets >= 1 && n_sets <= 4 && B > 0 && h > 1 && w > 1, "basis_homography_forward: bad size");
  dmh::BasisHArgs a = {};
  for (int i = 0; i < n_sets; ++i) {
    DMH_REQUIRE(weights[i] && H[i], "basis_homography_forward: null pointer in set %d", i);
    a.weight[i] = weights[i];
    a.H[i] = H[i];
  }
  const int N = n_sets * B;
  dmh::basis_h_fwd_kernel<<<(N + 3) / 4, 128, 0, dmh::as_stream(stream)>>>(basis, a, n_sets, B, h, w);
  return dmh::launched("basis_h_fwd_kernel");
}

extern "C" int dmh_basis_homography_backward(const float* basis, const float* const* H, const float* const* grad_H,
                                             float* const* grad_weight, int n_sets, int B, int h, int w, void* stream) {
  DMH_REQUIRE(basis && H && grad_H && grad_weight, "basis_homography_backward: null pointer");
  DMH_REQUIRE(n_sets >= 1 && n_sets <= 4 && B > 0 && h > 1 && w > 1, "basis_homography_backward: bad size");
  dmh::BasisHArgs a = {};
  for (int i = 0; i < n_sets; ++i) {
    DMH_REQUIRE(H[i] && grad_H[i] && grad_weight[i], "basis_homography_backward: null pointer in set %d", i);
    a.H[i] = const_cast<float*>(H[i]);
    a.grad_H[i] = grad_H[i];
    a.grad_weight[i] = grad_weight[i];
  }
  const int N = n_sets * B;
  dmh::basis_h_bwd_kernel<<<(N + 3) / 4, 128, 0, dmh::as_stream(stream)>>>(basis, a, n_sets, B, h, w);
  return dmh::launched("basis_h_bwd_kernel");
}
