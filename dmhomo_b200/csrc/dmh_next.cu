// The data formats either side of the warp path:
//
//  * pairs_u8_kernel      the on-disk pair format {"img12": (6,H,W) uint8} -> what the HEM loader hands the
//                         network: normalised grey full images, the crop patch, and the /255 RGB pair
//                         (HEM/dataset/data_loader.py:121-146, 217-255; generate_nyps_to_single_case.py:29-47);
//  * flow_upsample_*      upsample2d_flow_as(flow, target, if_rate) - bilinear, align_corners, optional rate
//                         scaling - and its adjoint, the step between the basis flow and the pyramid-level feature
//                         warp inside the backbone (HEM/model/utils.py:556-572, swin_multi.py:161-166, 1175-1182).
//
// Both are one-pass, coalesced, HBM-bound elementwise kernels (grid-stride over a bounded grid).
#include "dmh_common.cuh"

namespace dmh {

namespace {

constexpr int kThreads = 256;

struct PairNorm {
  double mean[3], std[3];
};

// grey = float32( ((n0 + n1) + n2) / 3 ), n_c = (u8_c - mean_c) / std_c in fp64: numpy's `(img - mean_I) / std_I`
// followed by `np.mean(axis=2)` (left-associated add.reduce, then a true division by the count) and torch.Tensor().
// n_c takes 256 values per channel: every CTA tabulates them once in shared memory (768 fp64 divisions per CTA
// instead of six per pixel), and float32(u8) / 255 likewise.
// One thread = four consecutive pixels of one sample (W % 4 == 0): six uchar4 loads, float4 stores.
__global__ void __launch_bounds__(kThreads) pairs_u8_kernel(const uint8_t* __restrict__ img12, const int* __restrict__ start,
                                                            float* __restrict__ grey_full, float* __restrict__ grey_patch,
                                                            float* __restrict__ rgb_full, const __grid_constant__ PairNorm nm,
                                                            int B, int H, int W, int ph, int pw, int planar) {
  __shared__ double lut[3][256];
  __shared__ float lut255[256];
  for (int t = threadIdx.x; t < 768; t += blockDim.x) {
    const int c = t >> 8, v = t & 255;
    lut[c][v] = __ddiv_rn(__dsub_rn((double)v, nm.mean[c]), nm.std[c]);
  }
  for (int t = threadIdx.x; t < 256; t += blockDim.x) lut255[t] = __fdiv_rn((float)t, 255.f);
  __syncthreads();
  auto grey_of = [&](unsigned r, unsigned g, unsigned bb) -> float {
    return (float)__ddiv_rn(__dadd_rn(__dadd_rn(lut[0][r], lut[1][g]), lut[2][bb]), 3.0);
  };
  const long long plane = (long long)H * W;
  const unsigned quads = (unsigned)(plane / 4);
  DMH_PLANE_LOOP(b, q, B, quads) {
    const unsigned p = q * 4u;
    const int y = (int)(p / (unsigned)W), x = (int)(p - (unsigned)y * (unsigned)W);
    const uint8_t* src = img12 + (size_t)b * 6 * plane + p;
    uchar4 ch[6];
#pragma unroll
    for (int c = 0; c < 6; ++c) ch[c] = __ldg(reinterpret_cast<const uchar4*>(src + (size_t)c * plane));
    float4 g1, g2;
    g1.x = grey_of(ch[0].x, ch[1].x, ch[2].x); g1.y = grey_of(ch[0].y, ch[1].y, ch[2].y);
    g1.z = grey_of(ch[0].z, ch[1].z, ch[2].z); g1.w = grey_of(ch[0].w, ch[1].w, ch[2].w);
    g2.x = grey_of(ch[3].x, ch[4].x, ch[5].x); g2.y = grey_of(ch[3].y, ch[4].y, ch[5].y);
    g2.z = grey_of(ch[3].z, ch[4].z, ch[5].z); g2.w = grey_of(ch[3].w, ch[4].w, ch[5].w);
    if (grey_full) {
      float* o = grey_full + (size_t)b * 2 * plane + p;
      *reinterpret_cast<float4*>(o) = g1;
      *reinterpret_cast<float4*>(o + plane) = g2;
    }
    if (rgb_full) {   // torch.Tensor(uint8 image).float() / 255.
      float* o = rgb_full + (size_t)b * 6 * plane + p;
#pragma unroll
      for (int c = 0; c < 6; ++c)
        *reinterpret_cast<float4*>(o + (size_t)c * plane) = make_float4(lut255[ch[c].x], lut255[ch[c].y], lut255[ch[c].z], lut255[ch[c].w]);
    }
    if (grey_patch) {   // img[y0 : y0 + ph, x0 : x0 + pw] of the grey images (random_crop_tt)
      const int x0 = __ldg(start + 2 * b), y0 = __ldg(start + 2 * b + 1);
      const int py = y - y0;
      if (py >= 0 && py < ph) {
        // (B,2,ph,pw), or planar (2,B,ph,pw): the two images of a pair as two dense batches
        const size_t pp = (size_t)ph * pw, second = planar ? (size_t)B * pp : pp;
        float* o = grey_patch + (size_t)b * (planar ? pp : 2 * pp) + (size_t)py * pw;
        const float v1[4] = {g1.x, g1.y, g1.z, g1.w}, v2[4] = {g2.x, g2.y, g2.z, g2.w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const int px = x + k - x0;
          if (px >= 0 && px < pw) {
            o[px] = v1[k];
            o[second + px] = v2[k];
          }
        }
      }
    }
  }
}

// torch's upsample_bilinear2d source index (align_corners: scale * dst with scale = (in - 1) / (out - 1) in fp32;
// otherwise scale * (dst + 0.5) - 0.5 clamped at 0 with scale = in / out), index clamp and lambda.
__device__ __forceinline__ void src_index(int dst, int in_size, float scale, bool align, int& i0, int& i1, float& l0, float& l1) {
  float real = align ? scale * (float)dst : fmaxf(scale * ((float)dst + 0.5f) - 0.5f, 0.f);
  i0 = min((int)real, in_size - 1);
  i1 = i0 + ((i0 < in_size - 1) ? 1 : 0);
  l1 = fminf(fmaxf(real - (float)i0, 0.f), 1.f);
  l0 = 1.f - l1;
}

// out[b, c, y, x] = rate_c * bilinear(in[b, c]), c = 0: horizontal flow (rate_x), c = 1: vertical flow (rate_y).
// V output pixels of one row per thread (V = 4 when wo % 4 == 0: 128-bit stores, the row indices shared).
template <int V>
__global__ void __launch_bounds__(kThreads) flow_upsample_fwd_kernel(const float* __restrict__ in, float* __restrict__ out, int B,
                                                                     int hi, int wi, int ho, int wo, float sy, float sx,
                                                                     float rate_x, float rate_y, int align) {
  const long long plane = (long long)ho * wo;
  const unsigned groups = (unsigned)(plane / V);
  DMH_PLANE_LOOP(b, q, B, groups) {
    const unsigned p = q * V;
    const int y = (int)(p / (unsigned)wo), x0 = (int)(p - (unsigned)y * (unsigned)wo);
    int y0, y1;
    float ly0, ly1;
    src_index(y, hi, sy, align != 0, y0, y1, ly0, ly1);
    float o0[V], o1[V];
#pragma unroll
    for (int v = 0; v < V; ++v) {
      int xa, xb;
      float lx0, lx1;
      src_index(x0 + v, wi, sx, align != 0, xa, xb, lx0, lx1);
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        const float* s = in + ((size_t)b * 2 + c) * hi * wi;
        const float r = c ? rate_y : rate_x;
        // the in-place `inputs[:, c] *= rate` of the reference happens before the interpolation
        const float a = mul_rn(__ldg(s + (size_t)y0 * wi + xa), r), bq = mul_rn(__ldg(s + (size_t)y0 * wi + xb), r);
        const float cq = mul_rn(__ldg(s + (size_t)y1 * wi + xa), r), d = mul_rn(__ldg(s + (size_t)y1 * wi + xb), r);
        const float val = ly0 * (lx0 * a + lx1 * bq) + ly1 * (lx0 * cq + lx1 * d);
        if (c) o1[v] = val; else o0[v] = val;
      }
    }
    float* o = out + ((size_t)b * 2) * plane + p;
    if (V == 4) {
      *reinterpret_cast<float4*>(o) = make_float4(o0[0], o0[1], o0[V - 2], o0[V - 1]);
      *reinterpret_cast<float4*>(o + plane) = make_float4(o1[0], o1[1], o1[V - 2], o1[V - 1]);
    } else {
      o[0] = o0[0];
      o[plane] = o1[0];
    }
  }
}

// adjoint: grad_in[b, c, yi, xi] = rate_c * sum over the output pixels whose footprint contains (yi, xi).
// Gather form (no atomics): one thread per INPUT pixel walks the output rows / columns that map near it; with
// align_corners the outputs touching input row yi lie in (yi - 1, yi + 1) / scale.
__global__ void __launch_bounds__(kThreads) flow_upsample_bwd_kernel(const float* __restrict__ gout, float* __restrict__ gin, int B,
                                                                     int hi, int wi, int ho, int wo, float sy, float sx,
                                                                     float rate_x, float rate_y, int align) {
  const long long plane_i = (long long)hi * wi;
  const long long plane_o = (long long)ho * wo;
  DMH_PLANE_LOOP(b, p, B, plane_i) {
    const int yi = (int)(p / (unsigned)wi), xi = (int)(p - (unsigned)yi * (unsigned)wi);
    // candidate output range: the outputs whose source coordinate lies in (yi - 1, yi + 1), two pixels of slack, the
    // clamped ends widened to the border; every candidate is verified through src_index itself
    const float isy = (sy > 0.f) ? 1.f / sy : (float)ho, isx = (sx > 0.f) ? 1.f / sx : (float)wo;
    const float oy = align ? 0.f : 0.5f, ox = align ? 0.f : 0.5f;   // real = s * (dst + o) - o
    int ya = (int)floorf(((float)yi - 1.f + oy) * isy - oy) - 2, yb = (int)ceilf(((float)yi + 1.f + oy) * isy - oy) + 2;
    int xa = (int)floorf(((float)xi - 1.f + ox) * isx - ox) - 2, xb = (int)ceilf(((float)xi + 1.f + ox) * isx - ox) + 2;
    if (yi == 0) ya = 0;
    if (yi == hi - 1) yb = ho - 1;
    if (xi == 0) xa = 0;
    if (xi == wi - 1) xb = wo - 1;
    ya = max(ya, 0); yb = min(yb, ho - 1); xa = max(xa, 0); xb = min(xb, wo - 1);
    float acc0 = 0.f, acc1 = 0.f;
    for (int y = ya; y <= yb; ++y) {
      int y0, y1;
      float ly0, ly1;
      src_index(y, hi, sy, align != 0, y0, y1, ly0, ly1);
      const float wy = ((y0 == yi) ? ly0 : 0.f) + ((y1 == yi) ? ly1 : 0.f);
      if (wy == 0.f) continue;
      for (int x = xa; x <= xb; ++x) {
        int x0, x1;
        float lx0, lx1;
        src_index(x, wi, sx, align != 0, x0, x1, lx0, lx1);
        const float wx = ((x0 == xi) ? lx0 : 0.f) + ((x1 == xi) ? lx1 : 0.f);
        if (wx == 0.f) continue;
        const float wgt = wy * wx;
        acc0 = fmaf(wgt, __ldg(gout + ((size_t)b * 2) * plane_o + (size_t)y * wo + x), acc0);
        acc1 = fmaf(wgt, __ldg(gout + ((size_t)b * 2 + 1) * plane_o + (size_t)y * wo + x), acc1);
      }
    }
    gin[((size_t)b * 2) * plane_i + p] = acc0 * rate_x;
    gin[((size_t)b * 2 + 1) * plane_i + p] = acc1 * rate_y;
  }
}

float scale_of(int in, int out, bool align) {
  if (align) return (out > 1) ? (float)(in - 1) / (float)(out - 1) : 0.f;
  return (float)in / (float)out;
}

// dst = fl(fl(float(u8) * scale) + bias): 16 bytes in, 64 bytes out per thread and step
__global__ void __launch_bounds__(kThreads) u8_to_f32_kernel(const uint8_t* __restrict__ src, float* __restrict__ dst, long long n16,
                                                             long long n, float scale, float bias) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n16; i += stride) {
    const uint4 v = __ldg(reinterpret_cast<const uint4*>(src) + i);
    const unsigned wd[4] = {v.x, v.y, v.z, v.w};
    float4* o = reinterpret_cast<float4*>(dst) + 4 * i;
#pragma unroll
    for (int k = 0; k < 4; ++k)
      o[k] = make_float4(add_rn(mul_rn((float)(wd[k] & 255u), scale), bias), add_rn(mul_rn((float)((wd[k] >> 8) & 255u), scale), bias),
                         add_rn(mul_rn((float)((wd[k] >> 16) & 255u), scale), bias), add_rn(mul_rn((float)(wd[k] >> 24), scale), bias));
  }
  for (long long i = n16 * 16 + blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += stride)
    dst[i] = add_rn(mul_rn((float)src[i], scale), bias);
}

// normalize / unnormalize / unormalise_and_convert_mapping_to_flow (flow_and_mapping_operations.py:227-451), channel-first
// (B,2,H,W): channel 0 against W, channel 1 against H, every operation rounded separately in the reference's order:
//   mode 0  normalize    2 * t / (S - 1) - 1
//   mode 1  unnormalize  (t + 1) * (S - 1) / 2
//   mode 2  unnormalize, then minus the pixel grid (mapping -> flow)
__global__ void __launch_bounds__(kThreads) grid_normalize_kernel(const float* __restrict__ src, float* __restrict__ dst, int B, int H,
                                                                  int W, int mode) {
  const unsigned plane = (unsigned)(H * W);
  const float sw = (float)(W - 1), sh = (float)(H - 1);
  DMH_PLANE_LOOP(b, p, B, plane) {
    const int y = (int)(p / (unsigned)W), x = (int)(p - (unsigned)y * (unsigned)W);
    const size_t o = (size_t)b * 2 * plane + p;
    const float u = src[o], v = src[o + plane];
    float ru, rv;
    if (mode == 0) {
      ru = sub_rn(div_rn(mul_rn(2.f, u), sw), 1.f);
      rv = sub_rn(div_rn(mul_rn(2.f, v), sh), 1.f);
    } else {
      ru = div_rn(mul_rn(add_rn(u, 1.f), sw), 2.f);
      rv = div_rn(mul_rn(add_rn(v, 1.f), sh), 2.f);
      if (mode == 2) {
        ru = sub_rn(ru, (float)x);
        rv = sub_rn(rv, (float)y);
      }
    }
    dst[o] = ru;
    dst[o + plane] = rv;
  }
}

}  // namespace
}  // namespace dmh

extern "C" int dmh_grid_normalize(const float* src, float* dst, int B, int H, int W, int mode, void* stream) {
  DMH_REQUIRE(src && dst, "grid_normalize: null pointer");
  DMH_REQUIRE(B > 0 && H > 0 && W > 0 && (long long)H * W < 2147483647LL, "grid_normalize: bad size");
  DMH_REQUIRE(mode >= 0 && mode <= 2, "grid_normalize: bad mode %d", mode);
  dmh::grid_normalize_kernel<<<dmh::plane_grid((long long)H * W, B), dmh::kThreads, 0, dmh::as_stream(stream)>>>(src, dst, B, H, W, mode);
  return dmh::launched("grid_normalize_kernel");
}

extern "C" int dmh_u8_to_f32(const uint8_t* src, float* dst, int64_t n, float scale, float bias, void* stream) {
  DMH_REQUIRE(src && dst && n > 0, "u8_to_f32: null pointer or empty");
  const bool vec = ((reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(dst)) & 15) == 0;
  const long long n16 = vec ? n / 16 : 0;
  long long blocks = ((vec ? n16 : n) + dmh::kThreads - 1) / dmh::kThreads;
  if (blocks > dmh::kNumSMs * 16) blocks = dmh::kNumSMs * 16;
  if (blocks < 1) blocks = 1;
  dmh::u8_to_f32_kernel<<<(unsigned)blocks, dmh::kThreads, 0, dmh::as_stream(stream)>>>(src, dst, n16, n, scale, bias);
  return dmh::launched("u8_to_f32_kernel");
}

extern "C" int dmh_pairs_u8_to_gray(const uint8_t* img12, const int* start, float* gray_full, float* gray_patch, float* rgb_full,
                                    const double* mean3, const double* std3, int B, int H, int W, int patch_h, int patch_w,
                                    int patch_planar, void* stream) {
  DMH_REQUIRE(img12 && mean3 && std3, "pairs_u8_to_gray: null pointer");
  DMH_REQUIRE(gray_full || gray_patch || rgb_full, "pairs_u8_to_gray: no output requested");
  DMH_REQUIRE(B > 0 && H > 0 && W > 0 && (long long)H * W < 2147483647LL, "pairs_u8_to_gray: bad size");
  DMH_REQUIRE((W & 3) == 0, "pairs_u8_to_gray: W must be a multiple of 4 (got %d)", W);
  DMH_REQUIRE(!gray_patch || (start && patch_h > 0 && patch_w > 0), "pairs_u8_to_gray: a patch needs start (B,2) and a size");
  dmh::PairNorm nm;
  for (int c = 0; c < 3; ++c) {
    nm.mean[c] = mean3[c];
    nm.std[c] = std3[c];
    DMH_REQUIRE(std3[c] != 0.0, "pairs_u8_to_gray: std[%d] is zero", c);
  }
  dmh::pairs_u8_kernel<<<dmh::plane_grid((long long)H * W / 4, B), dmh::kThreads, 0, dmh::as_stream(stream)>>>(img12, start, gray_full, gray_patch,
                                                                                        rgb_full, nm, B, H, W, patch_h, patch_w, patch_planar);
  return dmh::launched("pairs_u8_kernel");
}

extern "C" int dmh_flow_upsample(const float* flow, float* out, int B, int hi, int wi, int ho, int wo, int if_rate,
                                 int align_corners, void* stream) {
  DMH_REQUIRE(flow && out, "flow_upsample: null pointer");
  DMH_REQUIRE(B > 0 && hi > 0 && wi > 0 && ho > 0 && wo > 0, "flow_upsample: non-positive size");
  const bool al = align_corners != 0;
  // Python float division w / w_ is fp64; the in-place multiply of an fp32 tensor rounds the scalar to fp32 first
  const float rx = if_rate ? (float)((double)wo / (double)wi) : 1.f, ry = if_rate ? (float)((double)ho / (double)hi) : 1.f;
  if ((wo & 3) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0)
    dmh::flow_upsample_fwd_kernel<4><<<dmh::plane_grid((long long)ho * wo / 4, B), dmh::kThreads, 0, dmh::as_stream(stream)>>>(
        flow, out, B, hi, wi, ho, wo, dmh::scale_of(hi, ho, al), dmh::scale_of(wi, wo, al), rx, ry, al ? 1 : 0);
  else
    dmh::flow_upsample_fwd_kernel<1><<<dmh::plane_grid((long long)ho * wo, B), dmh::kThreads, 0, dmh::as_stream(stream)>>>(
        flow, out, B, hi, wi, ho, wo, dmh::scale_of(hi, ho, al), dmh::scale_of(wi, wo, al), rx, ry, al ? 1 : 0);
  return dmh::launched("flow_upsample_fwd_kernel");
}

extern "C" int dmh_flow_upsample_backward(const float* grad_out, float* grad_flow, int B, int hi, int wi, int ho, int wo,
                                          int if_rate, int align_corners, void* stream) {
  DMH_REQUIRE(grad_out && grad_flow, "flow_upsample_backward: null pointer");
  DMH_REQUIRE(B > 0 && hi > 0 && wi > 0 && ho > 0 && wo > 0, "flow_upsample_backward: non-positive size");
  const bool al = align_corners != 0;
  const float rx = if_rate ? (float)((double)wo / (double)wi) : 1.f, ry = if_rate ? (float)((double)ho / (double)hi) : 1.f;
  dmh::flow_upsample_bwd_kernel<<<dmh::plane_grid((long long)hi * wi, B), dmh::kThreads, 0, dmh::as_stream(stream)>>>(
      grad_out, grad_flow, B, hi, wi, ho, wo, dmh::scale_of(hi, ho, al), dmh::scale_of(wi, wo, al), rx, ry, al ? 1 : 0);
  return dmh::launched("flow_upsample_bwd_kernel");
}
