// DGM-side kernels: cv2.warpPerspective-compatible warp (S4) and least-squares flow -> homography.
#include "dmh_common.cuh"

namespace dmh {

// cv::invert on a 3x3 double matrix (closed form: cofactors times 1/det) - the inverse
// cv2.warpPerspective takes of H before mapping destination pixels.
__device__ __forceinline__ bool invert3(const double* s, double* t) {
  const double c0 = s[4] * s[8] - s[5] * s[7];
  const double c1 = s[3] * s[8] - s[5] * s[6];
  const double c2 = s[3] * s[7] - s[4] * s[6];
  double d = s[0] * c0 - s[1] * c1 + s[2] * c2;
  if (d == 0.0) return false;
  d = 1.0 / d;
  t[0] = c0 * d;
  t[1] = (s[2] * s[7] - s[1] * s[8]) * d;
  t[2] = (s[1] * s[5] - s[2] * s[4]) * d;
  t[3] = (s[5] * s[6] - s[3] * s[8]) * d;
  t[4] = (s[0] * s[8] - s[2] * s[6]) * d;
  t[5] = (s[2] * s[3] - s[0] * s[5]) * d;
  t[6] = c2 * d;
  t[7] = (s[1] * s[6] - s[0] * s[7]) * d;
  t[8] = (s[0] * s[4] - s[1] * s[3]) * d;
  return true;
}

// One thread per destination pixel; all channels.  Coordinates follow OpenCV's
// warpPerspective: per 64-wide block X0 = M0*bx + M1*y + M2, then (X0 + M0*x1) * (32 / W),
// rounded half-to-even to 1/32 pixel; constant-zero border per tap.  ddpm.py:1520-1529, data_loader.py:151,
// generate_nyps_to_single_case.py:15.
//   float32 images: bilinear taps weighted by the float table entries (1-fy)(1-fx) ..
//   uint8 images:   OpenCV's fixed-point remap - the same table scaled to 2^15 (exact integers at 1/32 steps),
//                   (sum w_i p_i + 2^14) >> 15; bit-identical to cv2 on uint8 (tests/test_oracle_golden.py).
// The inverse of H is taken once per CTA (one fp64 division instead of one per thread).
// MAP: cv2.remap(src, map_x, map_y, INTER_LINEAR, BORDER_CONSTANT) instead - the source coordinates come from a
// (B,2,h,w) fp32 tensor `map` (absolute coordinates; disp != 0: displacements, the pixel grid is added in fp32, which is
// the reference's fp64 sum rounded to fp32) and are fixed to 1/32 px by cvRound(v * 32) (convertMaps / remap, round
// half to even, out of range -> INT_MIN as cvtss2si gives it), the integer part saturated to int16 as OpenCV stores it.
template <typename T, bool MAP>
__global__ void __launch_bounds__(256) warp_persp_kernel(const T* __restrict__ src, const double* __restrict__ H,
                                                         const float* __restrict__ map, int disp, T* __restrict__ dst, int B,
                                                         int C, int Hs, int Ws, int h, int w, int cl, int bw0) {
  constexpr bool kU8 = sizeof(T) == 1;
  const int b = blockIdx.z;
  const int x = blockIdx.x * 64 + (threadIdx.x & 63);
  const int y = blockIdx.y * 4 + (threadIdx.x >> 6);
  __shared__ double Ms[9];
  if (!MAP && threadIdx.x == 0) {
    double Hm[9], Mi[9];
#pragma unroll
    for (int k = 0; k < 9; ++k) Hm[k] = __ldg(H + (size_t)b * 9 + k);
    if (!invert3(Hm, Mi)) {
#pragma unroll
      for (int k = 0; k < 9; ++k) Mi[k] = 0.0;
    }
#pragma unroll
    for (int k = 0; k < 9; ++k) Ms[k] = Mi[k];
  }
  if (!MAP) __syncthreads();
  if (x >= w || y >= h) return;
  int X, Y, sx, sy;
  if (MAP) {
    const size_t po = (size_t)b * 2 * h * w + (size_t)y * w + x;
    float mx = __ldg(map + po), my = __ldg(map + po + (size_t)h * w);
    if (disp) {
      mx = add_rn((float)x, mx);
      my = add_rn((float)y, my);
    }
    const float fx32 = mul_rn(mx, 32.0f), fy32 = mul_rn(my, 32.0f);
    X = (fabsf(fx32) < 2147483648.0f) ? __float2int_rn(fx32) : (int)0x80000000;    // (NaN compares false)
    Y = (fabsf(fy32) < 2147483648.0f) ? __float2int_rn(fy32) : (int)0x80000000;
    sx = min(max(X >> 5, -32768), 32767);
    sy = min(max(Y >> 5, -32768), 32767);
  } else {
  double M[9];
#pragma unroll
  for (int k = 0; k < 9; ++k) M[k] = Ms[k];
  const int bx = (x / bw0) * bw0, x1 = x - bx;
  const double dbx = (double)bx, dy = (double)y, dx1 = (double)x1;
  const double X0 = __dadd_rn(__dadd_rn(__dmul_rn(M[0], dbx), __dmul_rn(M[1], dy)), M[2]);
  const double Y0 = __dadd_rn(__dadd_rn(__dmul_rn(M[3], dbx), __dmul_rn(M[4], dy)), M[5]);
  const double W0 = __dadd_rn(__dadd_rn(__dmul_rn(M[6], dbx), __dmul_rn(M[7], dy)), M[8]);
  double Wd = __dadd_rn(W0, __dmul_rn(M[6], dx1));
  Wd = (Wd != 0.0) ? __ddiv_rn(32.0, Wd) : 0.0;
  const double lo = -2147483648.0, hi = 2147483647.0;
  const double fX = fmax(lo, fmin(hi, __dmul_rn(__dadd_rn(X0, __dmul_rn(M[0], dx1)), Wd)));
  const double fY = fmax(lo, fmin(hi, __dmul_rn(__dadd_rn(Y0, __dmul_rn(M[3], dx1)), Wd)));
  X = __double2int_rn(fX);
  Y = __double2int_rn(fY);
  sx = X >> 5;
  sy = Y >> 5;
  }
  const int iax = X & 31, iay = Y & 31;
  const float ax = mul_rn((float)iax, 0.03125f), ay = mul_rn((float)iay, 0.03125f);   // i / 32, exact either way
  const float bx0 = sub_rn(1.0f, ax), by0 = sub_rn(1.0f, ay);
  const float w00 = mul_rn(by0, bx0), w01 = mul_rn(by0, ax), w10 = mul_rn(ay, bx0), w11 = mul_rn(ay, ax);
  // (1-fy)(1-fx) * 2^15 with fx, fy multiples of 1/32: (32-iay)(32-iax) * 32, exact
  const int i00 = (32 - iay) * (32 - iax) * 32, i01 = (32 - iay) * iax * 32, i10 = iay * (32 - iax) * 32, i11 = iay * iax * 32;
  const bool x0in = (sx >= 0 && sx < Ws), x1in = (sx + 1 >= 0 && sx + 1 < Ws);
  const bool y0in = (sy >= 0 && sy < Hs), y1in = (sy + 1 >= 0 && sy + 1 < Hs);
  const int cx0 = min(max(sx, 0), Ws - 1), cx1 = min(max(sx + 1, 0), Ws - 1);
  const int cy0 = min(max(sy, 0), Hs - 1), cy1 = min(max(sy + 1, 0), Hs - 1);
  // tap offsets inside one plane (32-bit; the host checks Hs * Ws * C < 2^31) and their border flags, once for all channels
  const int o00 = cy0 * Ws + cx0, o01 = cy0 * Ws + cx1, o10 = cy1 * Ws + cx0, o11 = cy1 * Ws + cx1;
  const bool in00 = x0in && y0in, in01 = x1in && y0in, in10 = x0in && y1in, in11 = x1in && y1in;
  const int plane_s = Hs * Ws, plane_d = h * w;
  const T* sp = src + (size_t)b * C * plane_s;
  T* dp = dst + (size_t)b * C * plane_d + (cl ? (y * w + x) * C : y * w + x);
  const int tap_mul = cl ? C : 1, chan_s = cl ? 1 : plane_s, chan_d = cl ? 1 : plane_d;
#pragma unroll 3
  for (int c = 0; c < C; ++c) {
    const T* spc = sp + c * chan_s;
    T v00 = __ldg(spc + o00 * tap_mul), v01 = __ldg(spc + o01 * tap_mul), v10 = __ldg(spc + o10 * tap_mul), v11 = __ldg(spc + o11 * tap_mul);
    v00 = in00 ? v00 : (T)0;
    v01 = in01 ? v01 : (T)0;
    v10 = in10 ? v10 : (T)0;
    v11 = in11 ? v11 : (T)0;
    T o;
    if (kU8) {
      const int acc = (int)v00 * i00 + (int)v01 * i01 + (int)v10 * i10 + (int)v11 * i11;
      o = (T)((acc + (1 << 14)) >> 15);
    } else {
      o = (T)add_rn(add_rn(add_rn(mul_rn((float)v00, w00), mul_rn((float)v01, w01)), mul_rn((float)v10, w10)), mul_rn((float)v11, w11));
    }
    dp[c * chan_d] = o;
  }
}

// ---- least-squares DLT over all pixels (homo_gen, ddpm.py:1577-1661) ---------------------------------
// Normal equations of the reference's system A h = b (rows [x y 1 0 0 0 -ux -uy], [0 0 0 x y 1 -vx -vy]),
// with the columns scaled (a diagonal change of variables: same minimiser) so that the 8x8 normal matrix
// is well conditioned in fp64.  29 unique sums per sample.
constexpr int kLsAcc = 29;
constexpr int kLsStride = 45;

__global__ void __launch_bounds__(256) ls_accumulate_kernel(const float* __restrict__ flow, double* __restrict__ ws,
                                                            int B, int h, int w) {
  const int b = blockIdx.y;
  const long long plane = (long long)h * w;
  const double isw = 1.0 / (double)max(w - 1, 1), ish = 1.0 / (double)max(h - 1, 1);
  const double iS = fmin(isw, ish);
  double a[kLsAcc];
#pragma unroll
  for (int k = 0; k < kLsAcc; ++k) a[k] = 0.0;
  for (long long p = blockIdx.x * (long long)blockDim.x + threadIdx.x; p < plane; p += (long long)gridDim.x * blockDim.x) {
    const int yi = (int)(p / w), xi = (int)(p - (long long)yi * w);
    const double x = (double)xi, y = (double)yi;
    const double u = x + (double)__ldg(flow + ((size_t)b * 2) * plane + p);
    const double v = y + (double)__ldg(flow + ((size_t)b * 2 + 1) * plane + p);
    const double px = x * isw, py = y * ish;
    const double c6 = -(u * x) * isw * iS, c7 = -(u * y) * ish * iS;
    const double d6 = -(v * x) * isw * iS, d7 = -(v * y) * ish * iS;
    // p p^T
    a[0] += px * px; a[1] += px * py; a[2] += px; a[3] += py * py; a[4] += py; a[5] += 1.0;
    // p c^T, p d^T
    a[6] += px * c6; a[7] += px * c7; a[8] += py * c6; a[9] += py * c7; a[10] += c6; a[11] += c7;
    a[12] += px * d6; a[13] += px * d7; a[14] += py * d6; a[15] += py * d7; a[16] += d6; a[17] += d7;
    // c c^T + d d^T
    a[18] += c6 * c6 + d6 * d6; a[19] += c6 * c7 + d6 * d7; a[20] += c7 * c7 + d7 * d7;
    // rhs
    a[21] += px * u; a[22] += py * u; a[23] += u;
    a[24] += px * v; a[25] += py * v; a[26] += v;
    a[27] += c6 * u + d6 * v; a[28] += c7 * u + d7 * v;
  }
  __shared__ double red[8][kLsAcc];
  const int lane = threadIdx.x & 31, wrp = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < kLsAcc; ++k) {
    const double v = warp_sum(a[k]);
    if (lane == 0) red[wrp][k] = v;
  }
  __syncthreads();
  if (threadIdx.x < kLsAcc) {
    double v = 0.0;
#pragma unroll
    for (int q = 0; q < 8; ++q) v += red[q][threadIdx.x];
    atomicAdd(ws + (size_t)b * kLsStride + threadIdx.x, v);
  }
}

__device__ __forceinline__ void solve8_warp_ls(double (&m)[9], int lane, double (&sol)[8]);

__global__ void __launch_bounds__(32) ls_solve_kernel(const double* __restrict__ ws, double* __restrict__ H, int B,
                                                      int h, int w) {
  const int b = blockIdx.x, lane = threadIdx.x;
  const double* a = ws + (size_t)b * kLsStride;
  // assemble the symmetric 8x8 normal matrix N and rhs r
  double N[8][8], r[8];
  const double P[3][3] = {{a[0], a[1], a[2]}, {a[1], a[3], a[4]}, {a[2], a[4], a[5]}};
  for (int i = 0; i < 8; ++i)
    for (int j = 0; j < 8; ++j) N[i][j] = 0.0;
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) {
      N[i][j] = P[i][j];
      N[3 + i][3 + j] = P[i][j];
    }
  for (int i = 0; i < 3; ++i) {
    N[i][6] = N[6][i] = a[6 + 2 * i];
    N[i][7] = N[7][i] = a[7 + 2 * i];
    N[3 + i][6] = N[6][3 + i] = a[12 + 2 * i];
    N[3 + i][7] = N[7][3 + i] = a[13 + 2 * i];
  }
  N[6][6] = a[18]; N[6][7] = N[7][6] = a[19]; N[7][7] = a[20];
  for (int i = 0; i < 8; ++i) r[i] = a[21 + i];
  double m[9], sol[8];
  const int row = lane & 7;
  for (int j = 0; j < 8; ++j) m[j] = N[row][j];
  m[8] = r[row];
  solve8_warp_ls(m, lane, sol);
  if (lane == 0) {
    const double isw = 1.0 / (double)max(w - 1, 1), ish = 1.0 / (double)max(h - 1, 1);
    const double iS = fmin(isw, ish);
    double* o = H + (size_t)b * 9;
    o[0] = sol[0] * isw; o[1] = sol[1] * ish; o[2] = sol[2];
    o[3] = sol[3] * isw; o[4] = sol[4] * ish; o[5] = sol[5];
    o[6] = sol[6] * isw * iS; o[7] = sol[7] * ish * iS; o[8] = 1.0;
  }
}

// partial-pivot elimination over warp shuffles (same scheme as the 4-point DLT)
__device__ __forceinline__ void solve8_warp_ls(double (&m)[9], int lane, double (&sol)[8]) {
  bool used = false;
  int piv[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    double best = (lane < 8 && !used) ? fabs(m[k]) : -1.0;
    int who = lane;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const double ob = __shfl_xor_sync(0xffffffffu, best, o);
      const int ow = __shfl_xor_sync(0xffffffffu, who, o);
      if (ob > best || (ob == best && ow < who)) {
        best = ob;
        who = ow;
      }
    }
    piv[k] = who;
    const double pk = __shfl_sync(0xffffffffu, m[k], who);
    const bool elim = (lane < 8) && !used && (lane != who);
    const double f = elim ? m[k] / pk : 0.0;
#pragma unroll
    for (int j = k + 1; j < 9; ++j) {
      const double pj = __shfl_sync(0xffffffffu, m[j], who);
      if (elim) m[j] = fma(-f, pj, m[j]);
    }
    if (lane == who) used = true;
  }
#pragma unroll
  for (int k = 7; k >= 0; --k) {
    double acc = m[8];
#pragma unroll
    for (int j = k + 1; j < 8; ++j) acc = fma(-m[j], sol[j], acc);
    const double xk = acc / m[k];
    sol[k] = __shfl_sync(0xffffffffu, xk, piv[k]);
  }
}

}  // namespace dmh

using namespace dmh;

namespace {
template <typename T>
int warp_perspective_launch(const T* src, const double* H, T* dst, int B, int C, int Hs, int Ws, int h, int w, int channels_last,
                            void* stream) {
  DMH_REQUIRE(src && H && dst, "warp_perspective: null pointer");
  DMH_REQUIRE(B > 0 && B <= 65535 && C > 0 && Hs > 0 && Ws > 0 && h > 0 && w > 0, "warp_perspective: bad size");
  DMH_REQUIRE((long long)Hs * Ws * C < 2147483647LL && (long long)h * w * C < 2147483647LL, "warp_perspective: image too large");
  // OpenCV's block geometry (BLOCK_SZ = 32): bh0 = min(16, h); bw0 = min(1024 / bh0, w)
  const int bh0 = h < 16 ? h : 16;
  int bw0 = 1024 / bh0;
  bw0 = bw0 < w ? bw0 : w;
  dim3 grid((unsigned)((w + 63) / 64), (unsigned)((h + 3) / 4), (unsigned)B);
  warp_persp_kernel<T, false><<<grid, 256, 0, as_stream(stream)>>>(src, H, nullptr, 0, dst, B, C, Hs, Ws, h, w, channels_last, bw0);
  return launched("warp_persp_kernel");
}

template <typename T>
int remap_launch(const T* src, const float* map, T* dst, int B, int C, int Hs, int Ws, int h, int w, int channels_last,
                 int displacement, void* stream) {
  DMH_REQUIRE(src && map && dst, "remap: null pointer");
  DMH_REQUIRE(B > 0 && B <= 65535 && C > 0 && Hs > 0 && Ws > 0 && h > 0 && w > 0, "remap: bad size");
  DMH_REQUIRE((long long)Hs * Ws * C < 2147483647LL && (long long)h * w * C < 2147483647LL, "remap: image too large");
  dim3 grid((unsigned)((w + 63) / 64), (unsigned)((h + 3) / 4), (unsigned)B);
  warp_persp_kernel<T, true><<<grid, 256, 0, as_stream(stream)>>>(src, nullptr, map, displacement, dst, B, C, Hs, Ws, h, w,
                                                                  channels_last, 1);
  return launched("remap_kernel");
}
}  // namespace

extern "C" int dmh_remap(const float* src, const float* map, float* dst, int B, int C, int Hs, int Ws, int h, int w,
                         int channels_last, int displacement, void* stream) {
  return remap_launch<float>(src, map, dst, B, C, Hs, Ws, h, w, channels_last, displacement, stream);
}

extern "C" int dmh_remap_u8(const uint8_t* src, const float* map, uint8_t* dst, int B, int C, int Hs, int Ws, int h, int w,
                            int channels_last, int displacement, void* stream) {
  return remap_launch<uint8_t>(src, map, dst, B, C, Hs, Ws, h, w, channels_last, displacement, stream);
}

extern "C" int dmh_warp_perspective(const float* src, const double* H, float* dst, int B, int C, int Hs, int Ws, int h,
                                    int w, int channels_last, void* stream) {
  return warp_perspective_launch<float>(src, H, dst, B, C, Hs, Ws, h, w, channels_last, stream);
}

extern "C" int dmh_warp_perspective_u8(const uint8_t* src, const double* H, uint8_t* dst, int B, int C, int Hs, int Ws, int h,
                                       int w, int channels_last, void* stream) {
  return warp_perspective_launch<uint8_t>(src, H, dst, B, C, Hs, Ws, h, w, channels_last, stream);
}

extern "C" int dmh_flow_to_homography_ls(const float* flow, double* H, double* workspace, int B, int h, int w,
                                         void* stream) {
  DMH_REQUIRE(flow && H && workspace, "flow_to_homography_ls: null pointer");
  DMH_REQUIRE(B > 0 && B <= 65535 && h > 0 && w > 0, "flow_to_homography_ls: bad size");
  cudaError_t e = cudaMemsetAsync(workspace, 0, sizeof(double) * kLsStride * (size_t)B, as_stream(stream));
  if (e != cudaSuccess) return fail(DMH_ECUDA, "flow_to_homography_ls: memset: %s", cudaGetErrorString(e));
  const long long plane = (long long)h * w;
  long long chunks = (plane + 256 * 8 - 1) / (256 * 8);
  ls_accumulate_kernel<<<dim3((unsigned)chunks, (unsigned)B), 256, 0, as_stream(stream)>>>(flow, workspace, B, h, w);
  int rc = launched("ls_accumulate_kernel");
  if (rc) return rc;
  ls_solve_kernel<<<B, 32, 0, as_stream(stream)>>>(workspace, H, B, h, w);
  return launched("ls_solve_kernel");
}
