// Fused flow-generate + bilinear-sample + validity-mask + masked-L1 (+ gradients) kernel.
//
// Replaces, per output pixel, the reference's chain
//   get_grid -> get_flow / basis combine -> get_warp_flow/transformer (or warp / flow_warp)
//   -> create_border_mask -> LossL1(mask*a, mask*b)            (and its autograd backward)
// HEM/model/utils.py:400-553, HEM/utils_operations/pixel_wise_mapping.py:55-113,
// HEM/utils_operations/flow_and_mapping_operations.py:40-71, HEM/loss/losses.py:10-17,142-146,
// DGM/denoising_diffusion_models/classifier_free_guidance.py:784-806.
//
// Work decomposition (B200: 148 SMs, HBM-bound gather/stencil, no tensor cores):
//   * one CTA = a 64 x 64 tile of one sample of one term (256 threads = 2 x 4 warps);
//   * a warp owns 32 consecutive columns, so target / output / mask traffic is one fully
//     coalesced 128-byte line per row and the four source taps of a warp fall on 1-2 lines;
//   * a thread walks 16 consecutive rows of ITS column.  Everything that only depends on the
//     column (h0*x, h3*x, h6*x) is hoisted; per-sample reductions (loss, dL/dH, dL/dw) stay in
//     registers for 16 pixels and then go warp shuffle -> shared memory -> ONE atomic per CTA
//     and value;
//   * backward scatter: the bottom taps of row r and the top taps of row r+1 of a column hit
//     the same two source addresses under any near-rigid warp, so they are merged in registers
//     and each column issues ~2 (not 4) fire-and-forget REDG per pixel and channel; lanes are
//     consecutive addresses, so a warp-wide REDG is one coalesced 128-byte reduction at L2.
#include "dmh_common.cuh"
#include "dmh_sampler.cuh"
#include "dmh_warp_fast.h"

namespace dmh {

constexpr int NT = 256;
constexpr int WX = 2;          // warps along x in a CTA
constexpr int WY = 4;          // warps along y in a CTA
constexpr int RPT = 16;        // rows per thread
constexpr int TW = 32 * WX;    // tile width  (64)
constexpr int TH = WY * RPT;   // tile height (64)
constexpr int kMaxBatch = 4;

enum { PASS_FWD = 0, PASS_BWD = 1, PASS_FUSED = 2 };

struct WarpBatch {
  dmh_warp_desc d[kMaxBatch];
};

template <int SAMPLER, int PARAM, int PASS, int CT>
__global__ void __launch_bounds__(NT, (CT == 1 && PARAM != DMH_PARAM_BASIS8) ? 3 : 2)
    warp_kernel(const __grid_constant__ WarpBatch batch) {
  const dmh_warp_desc& d = batch.d[blockIdx.y];
  constexpr bool kGrad = (PASS != PASS_FWD);
  constexpr bool kOut = (PASS != PASS_BWD);
  constexpr bool kCombine = kGrad && (CT > 0);  // merge vertically adjacent taps before REDG
  constexpr int CC = (CT > 0) ? CT : 1;
  const int C = CT ? CT : d.C;
  const int h = d.h, w = d.w, Hs = d.Hs, Ws = d.Ws;
  const int tiles_x = (w + TW - 1) / TW, tiles_y = (h + TH - 1) / TH;
  int t = blockIdx.x;
  const int b = t / (tiles_x * tiles_y);
  t -= b * tiles_x * tiles_y;
  const int tyi = t / tiles_x, txi = t - tyi * tiles_x;
  const int lane = threadIdx.x & 31, wrp = threadIdx.x >> 5;
  const int x = txi * TW + (wrp % WX) * 32 + lane;
  const int y_begin = tyi * TH + (wrp / WX) * RPT;
  const int y_end = min(y_begin + RPT, h);
  const bool col_live = x < w;

  const int plane_o = h * w;      // validated < 2^31 on the host
  const int plane_s = Hs * Ws;
  const size_t img_o = (size_t)b * C * plane_o;   // first output-shaped plane of sample b
  const size_t img_s = (size_t)b * C * plane_s;

  float sx = d.start_x, sy = d.start_y;
  if (d.start) {
    sx = __ldg(d.start + 2 * b);
    sy = __ldg(d.start + 2 * b + 1);
  }
  const float gx = add_rn((float)x, sx);

  // per-sample parameters
  float hm[9];
  float bw[8];
  if (PARAM == DMH_PARAM_HOMOGRAPHY && d.divide == 1) {
#pragma unroll
    for (int k = 0; k < 9; ++k) hm[k] = __ldg(d.param + (size_t)b * 9 + k);
  }
  if (PARAM == DMH_PARAM_BASIS8) {
#pragma unroll
    for (int k = 0; k < 8; ++k) bw[k] = __ldg(d.param + (size_t)b * 8 + k);
  }
  const bool single_h = (PARAM == DMH_PARAM_HOMOGRAPHY) && (d.divide == 1);
  // WarpImages (S1B) applies H to the un-offset grid and adds `start` afterwards
  // (HEM/model/utils.py:171-192); get_flow applies it to grid + start (utils.py:400-440).
  const float hx = (SAMPLER == DMH_S1B) ? (float)x : gx;

  const bool want_mask = (d.valid != nullptr) || d.use_border_mask;
  const bool has_loss = (d.loss_form != DMH_LOSS_NONE) && (d.target != nullptr);
  const bool masked_diff = (d.loss_form == DMH_LOSS_MASKED_DIFF);
  const bool want_gsrc = kGrad && (d.grad_src != nullptr);
  const bool want_gpar = kGrad && (d.grad_param != nullptr);

  float gscale = 0.f;
  if (kGrad && has_loss) {
    gscale = d.grad_loss_scale;
    if (PASS == PASS_BWD && d.grad_loss) gscale *= __ldg(d.grad_loss);
    if (d.sample_weight) gscale *= __ldg(d.sample_weight + b);
  }

  float lsum = 0.f;
  // HOMOGRAPHY: column-wise partial sums  sa = sum a, say = sum a*hy  (a = gcx/T, b = gcy/T,
  // c = -(a*qx + b*qy)); the hx factor is constant per thread and applied once at the end.
  // BASIS8: gacc[k] = sum gcx*bx_k + gcy*by_k.
  float gacc[9];
#pragma unroll
  for (int k = 0; k < 9; ++k) gacc[k] = 0.f;

  // vertically merged scatter state (kCombine): bottom taps of the previous row
  int p_ib = -1, p_id = -1;
  float pB[CC], pD[CC];
#pragma unroll
  for (int c = 0; c < CC; ++c) pB[c] = pD[c] = 0.f;

  if (col_live) {
    for (int y = y_begin; y < y_end; ++y) {
      const int po = y * w + x;  // offset in an output plane
      const float gy = add_rn((float)y, sy);

      // ---- sampling coordinate ------------------------------------------------------
      float fx = 0.f, fy = 0.f, cx, cy;
      float qx = 0.f, qy = 0.f, qT = 1.f;  // X/T', Y/T', T' (for the homography gradient)
      float hy = gy;
      int cell = 0;
      if (PARAM == DMH_PARAM_FLOW) {
        fx = __ldg(d.param + ((size_t)b * 2) * plane_o + po);
        fy = __ldg(d.param + ((size_t)b * 2 + 1) * plane_o + po);
        cx = add_rn(gx, fx);
        cy = add_rn(gy, fy);
      } else if (PARAM == DMH_PARAM_COORDS) {
        cx = __ldg(d.param + ((size_t)b * 2) * plane_o + po);
        cy = __ldg(d.param + ((size_t)b * 2 + 1) * plane_o + po);
        fx = sub_rn(cx, gx);
        fy = sub_rn(cy, gy);
      } else if (PARAM == DMH_PARAM_HOMOGRAPHY) {
        if (!single_h) {
          const int dv = d.divide;
          cell = min(y / (h / dv), dv - 1) * dv + min(x / (w / dv), dv - 1);
          const float* Hp = d.param + ((size_t)b * dv * dv + cell) * 9;
#pragma unroll
          for (int k = 0; k < 9; ++k) hm[k] = __ldg(Hp + k);
        }
        hy = (SAMPLER == DMH_S1B) ? (float)y : gy;
        float qX, qY;
        if (SAMPLER == DMH_S1B) {
          // torch.bmm(H, goal) on the CPU (MKL sgemm, k = 3): acc = h0*x; acc = fma(h1, y, acc);
          // acc = acc + h2   (pinned against the reference in tests/test_oracle_vs_reference.py)
          qX = add_rn(__fmaf_rn(hm[1], hy, mul_rn(hm[0], hx)), hm[2]);
          qY = add_rn(__fmaf_rn(hm[4], hy, mul_rn(hm[3], hx)), hm[5]);
          qT = add_rn(__fmaf_rn(hm[7], hy, mul_rn(hm[6], hx)), hm[8]);
        } else {
          // (h0*x + h1*y) + h2, every product and sum rounded separately (App. A.2)
          qX = add_rn(add_rn(mul_rn(hm[0], hx), mul_rn(hm[1], hy)), hm[2]);
          qY = add_rn(add_rn(mul_rn(hm[3], hx), mul_rn(hm[4], hy)), hm[5]);
          qT = add_rn(add_rn(mul_rn(hm[6], hx), mul_rn(hm[7], hy)), hm[8]);
        }
        if (!(fabsf(qT) >= 1e-7f)) qT = add_rn(qT, 1e-6f);
        qx = div_rn(qX, qT);
        qy = div_rn(qY, qT);
        fx = sub_rn(qx, hx);
        fy = sub_rn(qy, hy);
        cx = add_rn(gx, fx);
        cy = add_rn(gy, fy);
      } else {  // BASIS8: acc = b0*w0; acc += bk*wk, products rounded separately (A12)
        const float* bp = d.basis + po;
        fx = mul_rn(__ldg(bp), bw[0]);
        fy = mul_rn(__ldg(bp + plane_o), bw[0]);
#pragma unroll
        for (int k = 1; k < 8; ++k) {
          fx = add_rn(fx, mul_rn(__ldg(bp + (size_t)(2 * k) * plane_o), bw[k]));
          fy = add_rn(fy, mul_rn(__ldg(bp + (size_t)(2 * k + 1) * plane_o), bw[k]));
        }
        cx = add_rn(gx, fx);
        cy = add_rn(gy, fy);
      }

      // ---- M1 validity mask on fl(flow + grid) (no start), inclusive bounds w, h ----------
      float m = 1.f;
      bool m1 = true;
      if (want_mask) {
        const float mx = (PARAM == DMH_PARAM_COORDS) ? cx : add_rn(fx, (float)x);
        const float my = (PARAM == DMH_PARAM_COORDS) ? cy : add_rn(fy, (float)y);
        m1 = (mx >= 0.f) && (mx <= (float)w) && (my >= 0.f) && (my <= (float)h);
        if (kOut && d.valid) d.valid[(size_t)b * plane_o + po] = m1 ? 1 : 0;
        if (d.use_border_mask) m = m1 ? 1.f : 0.f;
      }
      if (d.soft_mask) m = mul_rn(m, __ldg(d.soft_mask + (size_t)b * plane_o + po));
      if (kOut && d.flow_out) {
        d.flow_out[((size_t)b * 2) * plane_o + po] = fx;
        d.flow_out[((size_t)b * 2 + 1) * plane_o + po] = fy;
      }

      // ---- taps ------------------------------------------------------------------------------
      Taps tp;
      int x0, y0, x1, y1;
      make_taps<SAMPLER>(cx, cy, Hs, Ws, tp, x0, y0, x1, y1);
      if (kOut && d.indices) {
        const size_t n = (size_t)d.B * plane_o, o = (size_t)b * plane_o + po;
        d.indices[o] = x0;
        d.indices[n + o] = y0;
        d.indices[2 * n + o] = x1;
        d.indices[3 * n + o] = y1;
      }

      float gcx = 0.f, gcy = 0.f, gmask = 0.f;
      float cA[CC], cB[CC], cC[CC], cD[CC];
      bool any_go = false;
#pragma unroll(CT ? CT : 1)
      for (int c = 0; c < C; ++c) {
        const float* sp = d.src + img_s + (size_t)c * plane_s;
        const size_t oo = img_o + (size_t)c * plane_o + po;
        float Ia = __ldg(sp + tp.ia), Ib = __ldg(sp + tp.ib), Ic = __ldg(sp + tp.ic), Id = __ldg(sp + tp.id);
        if (SAMPLER == DMH_S2_ZEROS) {
          Ia = tp.va ? Ia : 0.f; Ib = tp.vb ? Ib : 0.f; Ic = tp.vc ? Ic : 0.f; Id = tp.vd ? Id : 0.f;
        }
        const float wv = blend<SAMPLER>(tp, Ia, Ib, Ic, Id);
        if (kOut && d.out) d.out[oo] = wv;
        float go = 0.f;  // dL/d(out)
        if (PASS == PASS_BWD && d.grad_out) go = __ldg(d.grad_out + oo);
        if (has_loss) {
          const float tv = __ldg(d.target + oo);
          float s, gm_c;
          if (masked_diff) {
            const float u = sub_rn(mul_rn(m, tv), mul_rn(m, wv));
            if (kOut) lsum += fabsf(u);
            s = sign_of(u);                 // d|u|/d(tv) = m*s ; d/d(wv) = -m*s ; d/dm = s*(tv-wv)
            gm_c = s * (tv - wv);
            s = m * s;
          } else {
            const float u = sub_rn(wv, tv);
            if (kOut) lsum += m * fabsf(u);
            gm_c = fabsf(u);
            s = -m * sign_of(u);            // d/d(tv) = -m*sign ; d/d(wv) = +m*sign
          }
          if (kGrad) {
            const float gt = gscale * s;
            go -= gt;
            if (d.grad_target && gt != 0.f) red_add(d.grad_target + oo, gt);
            gmask += gscale * gm_c;
          }
        }
        if (kGrad) {
          if (want_gsrc) {
            const float a_ = (SAMPLER != DMH_S2_ZEROS || tp.va) ? tp.wa * go : 0.f;
            const float b_ = (SAMPLER != DMH_S2_ZEROS || tp.vb) ? tp.wb * go : 0.f;
            const float c_ = (SAMPLER != DMH_S2_ZEROS || tp.vc) ? tp.wc * go : 0.f;
            const float d_ = (SAMPLER != DMH_S2_ZEROS || tp.vd) ? tp.wd * go : 0.f;
            if (kCombine) {
              cA[c] = a_; cB[c] = b_; cC[c] = c_; cD[c] = d_;
              any_go = any_go || (go != 0.f);
            } else if (go != 0.f) {
              float* gp = d.grad_src + img_s + (size_t)c * plane_s;
              red_add(gp + tp.ia, a_);
              red_add(gp + tp.ib, b_);
              red_add(gp + tp.ic, c_);
              red_add(gp + tp.id, d_);
            }
          }
          // d out / d cx = ay1*(Ic-Ia) + ay0*(Id-Ib);  d out / d cy = ax1*(Ib-Ia) + ax0*(Id-Ic)
          gcx += go * (tp.ay1 * (Ic - Ia) + tp.ay0 * (Id - Ib));
          gcy += go * (tp.ax1 * (Ib - Ia) + tp.ax0 * (Id - Ic));
        }
      }

      if (kCombine && want_gsrc && any_go) {
        float* gp = d.grad_src + img_s;
        if (p_ib >= 0) {
          if (p_ib == tp.ia && p_id == tp.ic) {  // previous bottom taps == this row's top taps
#pragma unroll
            for (int c = 0; c < CC; ++c) {
              cA[c] += pB[c];
              cC[c] += pD[c];
            }
          } else {
#pragma unroll
            for (int c = 0; c < CC; ++c) {
              red_add(gp + (size_t)c * plane_s + p_ib, pB[c]);
              red_add(gp + (size_t)c * plane_s + p_id, pD[c]);
            }
          }
        }
#pragma unroll
        for (int c = 0; c < CC; ++c) {
          red_add(gp + (size_t)c * plane_s + tp.ia, cA[c]);
          red_add(gp + (size_t)c * plane_s + tp.ic, cC[c]);
          pB[c] = cB[c];
          pD[c] = cD[c];
        }
        p_ib = tp.ib;
        p_id = tp.id;
      }

      if (kGrad) {
        gcx *= tp.gate_x;
        gcy *= tp.gate_y;
        if (d.grad_soft_mask)
          d.grad_soft_mask[(size_t)b * plane_o + po] = (d.use_border_mask && !m1) ? 0.f : gmask;
        if (want_gpar) {
          if (PARAM == DMH_PARAM_FLOW || PARAM == DMH_PARAM_COORDS) {
            d.grad_param[((size_t)b * 2) * plane_o + po] = gcx;
            d.grad_param[((size_t)b * 2 + 1) * plane_o + po] = gcy;
          } else if (PARAM == DMH_PARAM_HOMOGRAPHY) {
            // flow = q/T' - g  =>  dL/dX = gcx/T', dL/dY = gcy/T', dL/dT = -(gcx*X + gcy*Y)/T'^2
            const float rT = __frcp_rn(qT);
            const float ga = gcx * rT, gb = gcy * rT;
            const float gc = -(ga * qx + gb * qy);
            if (single_h) {
              gacc[0] += ga; gacc[1] = fmaf(ga, hy, gacc[1]);
              gacc[2] += gb; gacc[3] = fmaf(gb, hy, gacc[3]);
              gacc[4] += gc; gacc[5] = fmaf(gc, hy, gacc[5]);
            } else {  // mesh of homographies: rare path, straight to global
              float* gp = d.grad_param + ((size_t)b * d.divide * d.divide + cell) * 9;
              red_add(gp + 0, ga * hx); red_add(gp + 1, ga * hy); red_add(gp + 2, ga);
              red_add(gp + 3, gb * hx); red_add(gp + 4, gb * hy); red_add(gp + 5, gb);
              red_add(gp + 6, gc * hx); red_add(gp + 7, gc * hy); red_add(gp + 8, gc);
            }
          } else {
            const float* bp = d.basis + po;
#pragma unroll
            for (int k = 0; k < 8; ++k)
              gacc[k] += gcx * __ldg(bp + (size_t)(2 * k) * plane_o) + gcy * __ldg(bp + (size_t)(2 * k + 1) * plane_o);
          }
        }
      }
    }
    if (kCombine && p_ib >= 0) {
      float* gp = d.grad_src + img_s;
#pragma unroll
      for (int c = 0; c < CC; ++c) {
        red_add(gp + (size_t)c * plane_s + p_ib, pB[c]);
        red_add(gp + (size_t)c * plane_s + p_id, pD[c]);
      }
    }
  }

  // ---- per-CTA reductions: warp shuffle -> shared -> one atomic per value ---------------------
  const bool reduce_param = want_gpar && (single_h || PARAM == DMH_PARAM_BASIS8);
  const bool reduce_loss = kOut && has_loss && d.loss_acc;
  if (!reduce_param && !reduce_loss) return;
  __shared__ float red[NT / 32][10];
  float v[10];
  if (PARAM == DMH_PARAM_HOMOGRAPHY) {
    // (sum a, sum a*hy) -> (hx*sum a, sum a*hy, sum a) for the three rows of H
    v[0] = gacc[0] * hx; v[1] = gacc[1]; v[2] = gacc[0];
    v[3] = gacc[2] * hx; v[4] = gacc[3]; v[5] = gacc[2];
    v[6] = gacc[4] * hx; v[7] = gacc[5]; v[8] = gacc[4];
  } else {
#pragma unroll
    for (int k = 0; k < 9; ++k) v[k] = gacc[k];
  }
  v[9] = lsum;
#pragma unroll
  for (int k = 0; k < 10; ++k) {
    if ((k == 9) ? reduce_loss : reduce_param) {
      const float s = warp_sum(v[k]);
      if (lane == 0) red[wrp][k] = s;
    }
  }
  __syncthreads();
  if (threadIdx.x < 10) {
    const int k = threadIdx.x;
    if ((k == 9 && reduce_loss) || (k < 9 && reduce_param)) {
      float s = 0.f;
#pragma unroll
      for (int q = 0; q < NT / 32; ++q) s += red[q][k];
      if (k == 9) {
        atomicAdd(d.loss_acc + b, (double)s);
      } else if (PARAM == DMH_PARAM_HOMOGRAPHY) {
        red_add(d.grad_param + (size_t)b * 9 + k, s);
      } else if (k < 8) {
        red_add(d.grad_param + (size_t)b * 8 + k, s);
      }
    }
  }
}

// ---- host dispatch -------------------------------------------------------------------------

template <int SAMPLER, int PARAM, int PASS, int CT>
static int launch(const WarpBatch& batch, int n, cudaStream_t stream) {
  const dmh_warp_desc& d = batch.d[0];
  const long long tiles = (long long)((d.w + TW - 1) / TW) * ((d.h + TH - 1) / TH) * d.B;
  if (tiles > 2147483647LL) return fail(DMH_EUNSUPPORTED, "warp: too many tiles (%lld)", tiles);
  dim3 grid((unsigned)tiles, (unsigned)n, 1);
  warp_kernel<SAMPLER, PARAM, PASS, CT><<<grid, NT, 0, stream>>>(batch);
  return launched("warp_kernel");
}

template <int SAMPLER, int PARAM, int PASS>
static int launch_c(const WarpBatch& batch, int n, cudaStream_t stream) {
  const int C = batch.d[0].C;
  if (SAMPLER == DMH_S1 || SAMPLER == DMH_S3_BORDER) {
    if (C == 1) return launch<SAMPLER, PARAM, PASS, 1>(batch, n, stream);
    if (C == 3) return launch<SAMPLER, PARAM, PASS, 3>(batch, n, stream);
  }
  return launch<SAMPLER, PARAM, PASS, 0>(batch, n, stream);
}

template <int SAMPLER, int PASS>
static int launch_p(const WarpBatch& batch, int n, cudaStream_t stream) {
  switch (batch.d[0].param_kind) {
    case DMH_PARAM_FLOW: return launch_c<SAMPLER, DMH_PARAM_FLOW, PASS>(batch, n, stream);
    case DMH_PARAM_COORDS: return launch_c<SAMPLER, DMH_PARAM_COORDS, PASS>(batch, n, stream);
    case DMH_PARAM_HOMOGRAPHY: return launch_c<SAMPLER, DMH_PARAM_HOMOGRAPHY, PASS>(batch, n, stream);
    case DMH_PARAM_BASIS8: return launch_c<SAMPLER, DMH_PARAM_BASIS8, PASS>(batch, n, stream);
  }
  return fail(DMH_EINVAL, "warp: bad param_kind %d", batch.d[0].param_kind);
}

template <int PASS>
static int launch_s(const WarpBatch& batch, int n, cudaStream_t stream) {
  switch (batch.d[0].sampler) {
    case DMH_S1: return launch_p<DMH_S1, PASS>(batch, n, stream);
    case DMH_S1B: return launch_p<DMH_S1B, PASS>(batch, n, stream);
    case DMH_S2_ZEROS: return launch_p<DMH_S2_ZEROS, PASS>(batch, n, stream);
    case DMH_S3_BORDER: return launch_p<DMH_S3_BORDER, PASS>(batch, n, stream);
  }
  return fail(DMH_EINVAL, "warp: bad sampler %d", batch.d[0].sampler);
}

static int validate(const dmh_warp_desc& d, bool backward) {
  DMH_REQUIRE(d.struct_size == sizeof(dmh_warp_desc), "warp: struct_size %u != %zu (ABI mismatch)", d.struct_size,
              sizeof(dmh_warp_desc));
  DMH_REQUIRE(d.B > 0 && d.C > 0 && d.Hs > 0 && d.Ws > 0 && d.h > 0 && d.w > 0, "warp: non-positive size");
  DMH_REQUIRE((long long)d.Hs * d.Ws < 2147483647LL && (long long)d.h * d.w < 2147483647LL, "warp: plane too large");
  DMH_REQUIRE(d.src != nullptr, "warp: src is null");
  DMH_REQUIRE(d.param != nullptr, "warp: param is null");
  DMH_REQUIRE(d.param_kind != DMH_PARAM_BASIS8 || d.basis != nullptr, "warp: BASIS8 needs basis");
  DMH_REQUIRE(d.loss_form >= DMH_LOSS_NONE && d.loss_form <= DMH_LOSS_DIFF_MASKED, "warp: bad loss_form");
  if (d.param_kind == DMH_PARAM_HOMOGRAPHY) {
    DMH_REQUIRE(d.divide >= 1, "warp: divide must be >= 1");
    DMH_REQUIRE(d.h % d.divide == 0 && d.w % d.divide == 0, "warp: h,w must be divisible by divide");
  }
  if (d.loss_form != DMH_LOSS_NONE) DMH_REQUIRE(d.target != nullptr, "warp: loss needs target");
  if (backward)
    DMH_REQUIRE(d.grad_out != nullptr || d.loss_form != DMH_LOSS_NONE, "warp backward: no upstream gradient");
  return DMH_OK;
}

static bool same_config(const dmh_warp_desc& a, const dmh_warp_desc& b) {
  return a.sampler == b.sampler && a.param_kind == b.param_kind && a.B == b.B && a.C == b.C && a.Hs == b.Hs &&
         a.Ws == b.Ws && a.h == b.h && a.w == b.w;
}

static int run(const dmh_warp_desc* descs, int n, void* stream, bool backward) {
  DMH_REQUIRE(descs != nullptr && n > 0, "warp: no descriptors");
  for (int i = 0; i < n; ++i) {
    int rc = validate(descs[i], backward);
    if (rc) return rc;
  }
  int i = 0;
  while (i < n) {
    WarpBatch batch;
    int m = 0;
    const bool fused0 = !backward && descs[i].compute_grads;
    while (i + m < n && m < kMaxBatch && same_config(descs[i], descs[i + m]) &&
           (!backward && descs[i + m].compute_grads) == fused0) {
      batch.d[m] = descs[i + m];
      ++m;
    }
    const int pass = backward ? PASS_BWD : (fused0 ? PASS_FUSED : PASS_FWD);
    int done = 0, rc = DMH_OK;
    for (; done < m; done += 2) {  // lean specialisations take up to two terms per launch
      rc = warp_fast_try(descs + i + done, (m - done) < 2 ? (m - done) : 2, pass, as_stream(stream));
      if (rc != DMH_OK) break;
    }
    if (rc < 0) return rc;
    if (done >= m) {
      i += m;
      continue;
    }
    // general kernel for whatever the lean path did not take
    for (int k = done; k < m; ++k) batch.d[k - done] = descs[i + k];
    for (int k = m - done; k < kMaxBatch; ++k) batch.d[k] = batch.d[0];
    const int left = m - done;
    rc = backward ? launch_s<PASS_BWD>(batch, left, as_stream(stream))
                  : (fused0 ? launch_s<PASS_FUSED>(batch, left, as_stream(stream))
                            : launch_s<PASS_FWD>(batch, left, as_stream(stream)));
    if (rc) return rc;
    i += m;
  }
  return DMH_OK;
}

}  // namespace dmh

extern "C" int dmh_warp_forward(const dmh_warp_desc* descs, int n, void* stream) {
  return dmh::run(descs, n, stream, false);
}
extern "C" int dmh_warp_backward(const dmh_warp_desc* descs, int n, void* stream) {
  return dmh::run(descs, n, stream, true);
}
