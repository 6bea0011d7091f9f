// Lean specialisations of the fused warp kernel for the combinations the reference's training and
// evaluation loops actually run (everything else goes through the general kernel in dmh_warp.cu):
//
//   sampler  S1 (get_warp_flow / transformer, HEM/model/utils.py:443-553) or
//            S3 (flow_warp, ddpm.py:1262-1280 / data_loader.py:84-94)
//   param    one homography per sample (get_flow, utils.py:400-440) or an explicit flow
//   pass     forward | backward | forward + gradients in one pass
//   loss     none | |m*t - m*w| (losses.py:142-146) | m*|w - t| (classifier_free_guidance.py:799-806)
//   C        1 or 3
//
// Same decomposition as the general kernel (64x64 tile per CTA, a warp = 32 consecutive columns, a
// thread = 16 consecutive rows of one column, vertical tap merging before REDG), but every optional
// feature is a template parameter, the per-term descriptor is copied to registers once per CTA
// (no constant-bank indexing in the row loop), all global addresses are one 64-bit base plus a
// 32-bit element offset, and the row loop is unrolled by two so that the taps / target of row r+1
// are in flight while row r is blended, reduced and scattered.
#include "dmh_common.cuh"
#include "dmh_sampler.cuh"
#include "dmh_warp_fast.h"

namespace dmh {

namespace {

constexpr int NT = 256;
constexpr int WX = 2, WY = 4, RPT = 16;
constexpr int TW = 32 * WX, TH = WY * RPT;

enum { PASS_FWD = 0, PASS_BWD = 1, PASS_FUSED = 2 };

// default tuning (tools/tune.sh sweep, gpurun_out/tune*.txt): no software pipeline, L1 prefetch 3 rows
// ahead, register budget for 4 resident CTAs/SM (C = 1) - the kernel is issue-bound, so occupancy wins
constexpr int kTuneDense = 0 | 3 << 4 | 4 << 8;
constexpr int kTuneGeneral = 0 | 3 << 4 | 2 << 8;
constexpr int kTuneDenseC3 = 0 | 3 << 4 | 3 << 8;   // cfg4 sweep: 80 registers, 3 CTAs/SM

// Explicit global-space accesses (the pinned bases below are opaque to the compiler, which would
// otherwise fall back to generic-address atomics): ld.global.nc, st.global, red.global.add.
__device__ __forceinline__ float ldg_f(const float* base, unsigned off) {
  float v;
  asm("ld.global.nc.f32 %0, [%1];" : "=f"(v) : "l"(base + off));  // not volatile: free to hoist
  return v;
}
__device__ __forceinline__ void stg_f(float* base, unsigned off, float v) {
  asm volatile("st.global.f32 [%0], %1;" ::"l"(base + off), "f"(v) : "memory");
}
__device__ __forceinline__ void stg_u8(uint8_t* base, unsigned off, int v) {
  asm volatile("st.global.u8 [%0], %1;" ::"l"(base + off), "r"(v) : "memory");
}
__device__ __forceinline__ void red_f(float* base, unsigned off, float v) {
  asm volatile("red.global.add.f32 [%0], %1;" ::"l"(base + off), "f"(v) : "memory");
}

// Pin a per-CTA base pointer into a register pair: every later access is then ONE
// IMAD.WIDE.U32 (base + 4 * offset) instead of a re-derived 64-bit sum of uniform parts.
template <typename T>
__device__ __forceinline__ T* pin(T* p) {
  asm volatile("" : "+l"(p));
  return p;
}
__device__ __forceinline__ void prefetch_l1(const void* p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }
__device__ __forceinline__ float rcp_fast(float x) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}

// PROFILE 0: optional pointers are tested at run time.
// PROFILE 1 ("dense"): the caller guarantees the layout of the headline paths, so the tests fold away:
//   forward            out + valid written, no soft mask
//   backward / fused   border mask on, no soft mask, gradients to src, target and param all wanted,
//                      loss accumulated (fused); no grad_soft_mask, no grad_out.
// TUNE = PIPE | PF << 4 | MINB << 8: software-pipeline depth-2 on/off, L1 prefetch distance in rows (0 = off),
// minimum resident CTAs per SM for the register allocator.
template <int SAMPLER, int PARAM, int PASS, int CT, int LOSS, int PROFILE, int TUNE>
__global__ void __launch_bounds__(NT, (TUNE >> 8) & 15) warp_fast_kernel(const __grid_constant__ FastArgs a) {
  constexpr bool kPipe = (TUNE & 1) != 0;
  constexpr int kPF = (TUNE >> 4) & 15;
  constexpr bool kGrad = (PASS != PASS_FWD);
  constexpr bool kOut = (PASS != PASS_BWD);
  constexpr bool kLoss = (LOSS != DMH_LOSS_NONE);
  constexpr bool kDense = (PROFILE == 1);
  // ---- per-CTA setup: the term's fields live in registers from here on ----------------------
  const FastTerm tm = (blockIdx.y == 0) ? a.t[0] : a.t[1];
  const int h = a.h, w = a.w, Hs = a.Hs, Ws = a.Ws;
  int t = blockIdx.x;
  const int per = a.tiles_x * a.tiles_y;
  // a launch over C = groups x CT channels walks (sample, channel group) pairs: the planes of a group are contiguous
  // ((b * C + g * CT) = bv * CT), everything per sample (flow, masks, H, accumulators) is indexed by b
  const int bv = t / per;
  t -= bv * per;
  const int b = (a.groups > 1) ? bv / a.groups : bv;
  const bool first_group = (a.groups <= 1) || (bv == b * a.groups);
  const int tyi = t / a.tiles_x, txi = t - tyi * a.tiles_x;
  const int lane = threadIdx.x & 31, wrp = threadIdx.x >> 5;
  const int x = txi * TW + (wrp % WX) * 32 + lane;
  const int y_begin = tyi * TH + (wrp / WX) * RPT;
  const int y_end = min(y_begin + RPT, h);
  const bool col_live = x < w;
  const unsigned plane_o = (unsigned)(h * w), plane_s = (unsigned)(Hs * Ws);

  // sample-b bases; every access below is base[32-bit offset]
  const float* __restrict__ src = pin(tm.src + (size_t)bv * CT * plane_s);
  const float* __restrict__ tgt = kLoss ? pin(tm.target + (size_t)bv * CT * plane_o) : nullptr;
  const float* __restrict__ gout = (!kDense && PASS == PASS_BWD && tm.grad_out) ? tm.grad_out + (size_t)bv * CT * plane_o : nullptr;
  const float* __restrict__ soft = (!kDense && tm.soft_mask) ? tm.soft_mask + (size_t)b * plane_o : nullptr;
  const float* __restrict__ flow = (PARAM == DMH_PARAM_FLOW) ? tm.param + (size_t)b * 2 * plane_o : nullptr;
  const bool has_out = kOut && (kDense ? (PASS == PASS_FWD) : (tm.out != nullptr));
  const bool has_valid = kOut && (kDense ? (PASS == PASS_FWD) : (tm.valid != nullptr)) && first_group;
  const bool has_gsrc = kGrad && (kDense || tm.grad_src != nullptr);
  const bool has_gtgt = kGrad && kLoss && (kDense || tm.grad_target != nullptr);
  const bool has_gsoft = kGrad && !kDense && (tm.grad_soft_mask != nullptr);
  const bool has_gflow = kGrad && (PARAM == DMH_PARAM_FLOW) && (kDense || tm.grad_param != nullptr);
  float* __restrict__ out = has_out ? pin(tm.out + (size_t)bv * CT * plane_o) : nullptr;
  uint8_t* __restrict__ valid = has_valid ? tm.valid + (size_t)b * plane_o : nullptr;
  float* __restrict__ gsrc = has_gsrc ? pin(tm.grad_src + (size_t)bv * CT * plane_s) : nullptr;
  float* __restrict__ gtgt = has_gtgt ? pin(tm.grad_target + (size_t)bv * CT * plane_o) : nullptr;
  float* __restrict__ gsoft = has_gsoft ? tm.grad_soft_mask + (size_t)b * plane_o : nullptr;
  float* __restrict__ gflow = has_gflow ? tm.grad_param + (size_t)b * 2 * plane_o : nullptr;
  const bool want_gH = kGrad && (PARAM == DMH_PARAM_HOMOGRAPHY) && (kDense || tm.grad_param != nullptr);
  const bool use_border = kDense ? kGrad : (tm.use_border_mask != 0);
  const bool want_mask = use_border || has_valid;

  const float sx = a.sx, sy = a.sy;
  const float xf = (float)x;
  const float gx = add_rn(xf, sx);
  const float wf = (float)w, hf = (float)h;

  float hm[9];
  float h0x = 0.f, h3x = 0.f, h6x = 0.f;
  if (PARAM == DMH_PARAM_HOMOGRAPHY) {
#pragma unroll
    for (int k = 0; k < 9; ++k) hm[k] = __ldg(tm.param + (size_t)b * 9 + k);
    h0x = mul_rn(hm[0], gx);
    h3x = mul_rn(hm[3], gx);
    h6x = mul_rn(hm[6], gx);
  }

  float gscale = 0.f;
  if (kGrad && kLoss) {
    gscale = tm.grad_loss_scale;
    if (PASS == PASS_BWD && tm.grad_loss) gscale *= __ldg(tm.grad_loss);
    if (tm.sample_weight) gscale *= __ldg(tm.sample_weight + b);
  }

  float lsum = 0.f;
  float sa = 0.f, say = 0.f, sb = 0.f, sby = 0.f, sc = 0.f, scy = 0.f;  // dL/dH column sums
  int p_ib = -1, p_id = -1;                                              // merged scatter state
  float pB[CT], pD[CT];
#pragma unroll
  for (int c = 0; c < CT; ++c) pB[c] = pD[c] = 0.f;

  // One row of this thread's column, split in two stages so that the loads of row r+1 are in
  // flight while row r is blended, reduced and scattered (software pipeline, depth 2).
  struct Row {
    unsigned po;
    int ia, ib, ic, id;
    float ax0, ax1, ay0, ay1, gate_x, gate_y;
    float m, soft, gy, qx, qy, qT;
    bool m1;
    float I[CT][4], tv[CT], go[CT];
  };

  // stage 1: coordinate, mask, taps, issue every load of the row
  auto issue = [&](int y, Row& r) {
    const float yf = (float)y;
    const float gy = add_rn(yf, sy);
    const unsigned po = (unsigned)(y * w + x);
    r.po = po;
    r.gy = gy;
    float fx, fy;
    r.qx = r.qy = 0.f;
    r.qT = 1.f;
    if (PARAM == DMH_PARAM_FLOW) {
      fx = ldg_f(flow, po);
      fy = ldg_f(flow, po + plane_o);
    } else {
      // (h0*x + h1*y) + h2, every product and sum rounded separately (App. A.2)
      const float qX = add_rn(add_rn(h0x, mul_rn(hm[1], gy)), hm[2]);
      const float qY = add_rn(add_rn(h3x, mul_rn(hm[4], gy)), hm[5]);
      float qT = add_rn(add_rn(h6x, mul_rn(hm[7], gy)), hm[8]);
      if (!(fabsf(qT) >= 1e-7f)) qT = add_rn(qT, 1e-6f);
      r.qx = div_rn(qX, qT);
      r.qy = div_rn(qY, qT);
      r.qT = qT;
      fx = sub_rn(r.qx, gx);
      fy = sub_rn(r.qy, gy);
    }
    const float cx = add_rn(gx, fx), cy = add_rn(gy, fy);
    // M1 validity mask on fl(flow + grid) (no start), inclusive bounds w, h
    r.m = 1.f;
    r.m1 = true;
    if (want_mask) {
      const float mx = add_rn(fx, xf), my = add_rn(fy, yf);
      r.m1 = (mx >= 0.f) && (mx <= wf) && (my >= 0.f) && (my <= hf);
      if (has_valid) stg_u8(valid, po, r.m1 ? 1 : 0);
      if (use_border) r.m = r.m1 ? 1.f : 0.f;
    }
    r.soft = (!kDense && soft) ? ldg_f(soft, po) : 1.f;
    Taps tp;
    int x0, y0, x1, y1;
    make_taps<SAMPLER>(cx, cy, Hs, Ws, tp, x0, y0, x1, y1);
    r.ia = tp.ia; r.ib = tp.ib; r.ic = tp.ic; r.id = tp.id;
    r.ax0 = tp.ax0; r.ax1 = tp.ax1; r.ay0 = tp.ay0; r.ay1 = tp.ay1;
    r.gate_x = tp.gate_x; r.gate_y = tp.gate_y;
    if (kPF > 0) {
      // pull the lines rows kPF ahead will need into L1 now: the target row and (assuming the
      // column keeps marching down the source) the bottom tap row
      const unsigned pt = min(po + (unsigned)(kPF * w), plane_o - 1);
      const unsigned ps = min((unsigned)(tp.ib + kPF * Ws), plane_s - 1);
#pragma unroll
      for (int c = 0; c < CT; ++c) {
        if (kLoss) prefetch_l1(tgt + ((unsigned)c * plane_o + pt));
        prefetch_l1(src + ((unsigned)c * plane_s + ps));
      }
    }
#pragma unroll
    for (int c = 0; c < CT; ++c) {
      const unsigned cs = (unsigned)c * plane_s, oo = po + (unsigned)c * plane_o;
      r.I[c][0] = ldg_f(src, cs + tp.ia);
      r.I[c][1] = ldg_f(src, cs + tp.ib);
      r.I[c][2] = ldg_f(src, cs + tp.ic);
      r.I[c][3] = ldg_f(src, cs + tp.id);
      r.tv[c] = kLoss ? ldg_f(tgt, oo) : 0.f;
      r.go[c] = (!kDense && PASS == PASS_BWD && gout) ? ldg_f(gout, oo) : 0.f;
    }
  };

  // stage 2: blend, loss, gradients, scatter
  auto consume = [&](const Row& r) {
    const unsigned po = r.po;
    const float m = (!kDense && soft) ? mul_rn(r.m, r.soft) : r.m;
    Taps tp;
    tp.ax0 = r.ax0; tp.ax1 = r.ax1; tp.ay0 = r.ay0; tp.ay1 = r.ay1;
    // S1: wa=(x1f-x)(y1f-y) wb=(x1f-x)(y-y0f) wc=(x-x0f)(y1f-y) wd=(x-x0f)(y-y0f); S3: the same products
    tp.wa = mul_rn(r.ax1, r.ay1);
    tp.wb = mul_rn(r.ax1, r.ay0);
    tp.wc = mul_rn(r.ax0, r.ay1);
    tp.wd = mul_rn(r.ax0, r.ay0);
    float gcx = 0.f, gcy = 0.f, gmask = 0.f;
    float cA[CT], cB[CT], cC[CT], cD[CT];
#pragma unroll
    for (int c = 0; c < CT; ++c) {
      const unsigned oo = po + (unsigned)c * plane_o;
      const float Ia = r.I[c][0], Ib = r.I[c][1], Ic = r.I[c][2], Id = r.I[c][3];
      const float wv = blend<SAMPLER>(tp, Ia, Ib, Ic, Id);
      if (has_out) stg_f(out, oo, wv);
      float go = r.go[c];  // dL/d(out)
      if (kLoss) {
        const float tv = r.tv[c];
        float u;
        if (LOSS == DMH_LOSS_MASKED_DIFF) {
          u = sub_rn(mul_rn(m, tv), mul_rn(m, wv));   // |m*t - m*w|
          if (kOut) lsum += fabsf(u);
        } else {
          u = sub_rn(tv, wv);                         // m*|w - t|  (sign convention: u = t - w)
          if (kOut) lsum += m * fabsf(u);
        }
        if (kGrad) {
          // d/dt = +gm*sign(u), d/dw = -gm*sign(u): flip gm's sign bit with u's, zero where u == 0
          const float gm = gscale * m;
          float gt = __int_as_float(__float_as_int(gm) ^ (__float_as_int(u) & 0x80000000));
          gt = (u == 0.f) ? 0.f : gt;
          go -= gt;
          if (has_gtgt) red_f(gtgt, oo, gt);
          if (has_gsoft) {
            const float sg = __int_as_float(__float_as_int(gscale) ^ (__float_as_int(u) & 0x80000000));
            gmask += (u == 0.f) ? 0.f : ((LOSS == DMH_LOSS_MASKED_DIFF) ? sg * (tv - wv) : gscale * fabsf(u));
          }
        }
      }
      if (kGrad) {
        cA[c] = tp.wa * go; cB[c] = tp.wb * go; cC[c] = tp.wc * go; cD[c] = tp.wd * go;
        // d out / d cx = ay1*(Ic-Ia) + ay0*(Id-Ib);  d out / d cy = ax1*(Ib-Ia) + ax0*(Id-Ic)
        gcx = fmaf(go, fmaf(r.ay1, Ic - Ia, r.ay0 * (Id - Ib)), gcx);
        gcy = fmaf(go, fmaf(r.ax1, Ib - Ia, r.ax0 * (Id - Ic)), gcy);
      }
    }

    if (kGrad && has_gsrc) {
      // The previous row's bottom taps (pB at p_ib, pD at p_id) either coincide with this row's top
      // taps (any near-rigid warp: almost always) and ride along in registers, or are flushed now.
      // Rows whose upstream gradient is exactly zero (masked out) still post their (zero) top taps:
      // one uniform code path is cheaper than the divergence bookkeeping that would skip them.
      const bool same = (p_ib == r.ia) && (p_id == r.ic);
      if (!same && p_ib >= 0) {
#pragma unroll
        for (int c = 0; c < CT; ++c) {
          red_f(gsrc, (unsigned)c * plane_s + (unsigned)p_ib, pB[c]);
          red_f(gsrc, (unsigned)c * plane_s + (unsigned)p_id, pD[c]);
        }
      }
#pragma unroll
      for (int c = 0; c < CT; ++c) {
        const unsigned cs = (unsigned)c * plane_s;
        red_f(gsrc, cs + (unsigned)r.ia, cA[c] + (same ? pB[c] : 0.f));
        red_f(gsrc, cs + (unsigned)r.ic, cC[c] + (same ? pD[c] : 0.f));
        pB[c] = cB[c];
        pD[c] = cD[c];
      }
      p_ib = r.ib;
      p_id = r.id;
    }

    if (kGrad) {
      gcx *= r.gate_x;
      gcy *= r.gate_y;
      if (has_gsoft) stg_f(gsoft, po, (use_border && !r.m1) ? 0.f : gmask);
      if (has_gflow) {
        if (a.groups > 1) {                 // the channel groups of a sample add up (the launcher zeroed the buffer)
          red_f(gflow, po, gcx);
          red_f(gflow, po + plane_o, gcy);
        } else {
          stg_f(gflow, po, gcx);
          stg_f(gflow, po + plane_o, gcy);
        }
      }
      if (want_gH) {
        // flow = q/T' - g  =>  dL/dX = gcx/T', dL/dY = gcy/T', dL/dT = -(gcx*X + gcy*Y)/T'^2
        const float rT = rcp_fast(r.qT);
        const float ga = gcx * rT, gb = gcy * rT;
        const float gc = -fmaf(ga, r.qx, gb * r.qy);
        sa += ga; say = fmaf(ga, r.gy, say);
        sb += gb; sby = fmaf(gb, r.gy, sby);
        sc += gc; scy = fmaf(gc, r.gy, scy);
      }
    }
  };

  if (col_live && y_begin < y_end) {
    if (kPipe) {
      Row r0, r1;
      issue(y_begin, r0);
      int y = y_begin;
      // depth-2 pipeline, unrolled by two so that the row contexts never move between registers
      for (; y + 2 < y_end; y += 2) {
        issue(y + 1, r1);
        consume(r0);
        issue(y + 2, r0);
        consume(r1);
      }
      if (y + 1 < y_end) {
        issue(y + 1, r1);
        consume(r0);
        consume(r1);
      } else {
        consume(r0);
      }
    } else {
      for (int y = y_begin; y < y_end; ++y) {
        Row r;
        issue(y, r);
        consume(r);
      }
    }
    if (kGrad && has_gsrc && p_ib >= 0) {
#pragma unroll
      for (int c = 0; c < CT; ++c) {
        red_f(gsrc, (unsigned)c * plane_s + p_ib, pB[c]);
        red_f(gsrc, (unsigned)c * plane_s + p_id, pD[c]);
      }
    }
  }

  // ---- per-CTA reductions: warp shuffle -> shared -> one atomic per value ---------------------
  const bool reduce_loss = kOut && kLoss && (kDense || tm.loss_acc != nullptr);
  if (!want_gH && !reduce_loss) return;
  __shared__ float red[NT / 32][10];
  float v[10];
  v[0] = sa * gx; v[1] = say; v[2] = sa;
  v[3] = sb * gx; v[4] = sby; v[5] = sb;
  v[6] = sc * gx; v[7] = scy; v[8] = sc;
  v[9] = lsum;
#pragma unroll
  for (int k = 0; k < 10; ++k) {
    if ((k == 9) ? reduce_loss : want_gH) {
      const float s = warp_sum(v[k]);
      if (lane == 0) red[wrp][k] = s;
    }
  }
  __syncthreads();
  if (threadIdx.x < 10) {
    const int k = threadIdx.x;
    if ((k == 9) ? reduce_loss : want_gH) {
      float s = 0.f;
#pragma unroll
      for (int q = 0; q < NT / 32; ++q) s += red[q][k];
      if (k == 9)
        atomicAdd(tm.loss_acc + b, (double)s);
      else
        red_add(tm.grad_param + (size_t)b * 9 + k, s);
    }
  }
}

// Forward warp of a feature map with many channels by an explicit flow (the Swin pyramid levels, C = 12 / 24 at
// 80 x 144 and below, HEM/model/swin_multi.py:161-166): one thread per output pixel - coordinate, mask and taps once,
// then the channels in a loop (four coalesced loads, the sampler's own blend, one store each).  The tiled kernel above
// walks 16 rows per thread on 64 x 64 tiles, which leaves half the threads of such small planes idle and pays the
// coordinate arithmetic once per group of three channels: 68 us 46 us vs 26 us here on 64 x 12 x 80 x 144 under a basis flow.
// (The adjoint stays on the tiled kernel, channel group by channel group: its vertical tap merging halves the REDs -
// a pixel-per-thread adjoint with four REDs per channel and pixel measured 122 us against 69 us.)
template <int SAMPLER>
__global__ void __launch_bounds__(NT) warp_flow_channels_fwd_kernel(const __grid_constant__ FastArgs a, int C) {
  const FastTerm tm = (blockIdx.z == 0) ? a.t[0] : a.t[1];
  const int h = a.h, w = a.w, Hs = a.Hs, Ws = a.Ws;
  const unsigned plane_o = (unsigned)(h * w), plane_s = (unsigned)(Hs * Ws);
  const float wf = (float)w, hf = (float)h;
  DMH_PLANE_LOOP(b, p, a.B, plane_o) {
    const int y = (int)(p / (unsigned)w), x = (int)(p - (unsigned)y * (unsigned)w);
    const float xf = (float)x, yf = (float)y;
    const float gx = add_rn(xf, a.sx), gy = add_rn(yf, a.sy);
    const float* fl = tm.param + (size_t)b * 2 * plane_o;
    const float fx = __ldg(fl + p), fy = __ldg(fl + plane_o + p);
    const float cx = add_rn(gx, fx), cy = add_rn(gy, fy);
    if (tm.valid) {
      const float mx = add_rn(fx, xf), my = add_rn(fy, yf);
      tm.valid[(size_t)b * plane_o + p] = ((mx >= 0.f) && (mx <= wf) && (my >= 0.f) && (my <= hf)) ? 1 : 0;
    }
    Taps tp;
    int x0, y0, x1, y1;
    make_taps<SAMPLER>(cx, cy, Hs, Ws, tp, x0, y0, x1, y1);
    tp.wa = mul_rn(tp.ax1, tp.ay1);
    tp.wb = mul_rn(tp.ax1, tp.ay0);
    tp.wc = mul_rn(tp.ax0, tp.ay1);
    tp.wd = mul_rn(tp.ax0, tp.ay0);
    const float* __restrict__ sp = tm.src + (size_t)b * C * plane_s;
    float* __restrict__ op = tm.out + (size_t)b * C * plane_o + p;
    // four channels at a time: their sixteen loads are issued before the first blend (a store between two channels
    // would otherwise order the next channel's loads behind it)
    int c = 0;
    for (; c + 4 <= C; c += 4) {
      float I[4][4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float* sc = sp + (size_t)(c + k) * plane_s;
        I[k][0] = __ldg(sc + tp.ia); I[k][1] = __ldg(sc + tp.ib); I[k][2] = __ldg(sc + tp.ic); I[k][3] = __ldg(sc + tp.id);
      }
#pragma unroll
      for (int k = 0; k < 4; ++k) op[(size_t)(c + k) * plane_o] = blend<SAMPLER>(tp, I[k][0], I[k][1], I[k][2], I[k][3]);
    }
    for (; c < C; ++c) {
      const float* sc = sp + (size_t)c * plane_s;
      const float Ia = __ldg(sc + tp.ia), Ib = __ldg(sc + tp.ib), Ic = __ldg(sc + tp.ic), Id = __ldg(sc + tp.id);
      op[(size_t)c * plane_o] = blend<SAMPLER>(tp, Ia, Ib, Ic, Id);
    }
  }
}

template <int SAMPLER, int PARAM, int PASS, int CT, int LOSS>
int launch(const FastArgs& a, int n, long long tiles, int flags, cudaStream_t stream) {
  dim3 grid((unsigned)tiles, (unsigned)n, 1);
  // the dense profile exists for the S1 headline paths only (homography or flow parameterised)
  constexpr bool kHasDense = (SAMPLER == DMH_S1) && (PASS == PASS_FWD ? LOSS == DMH_LOSS_NONE : LOSS == DMH_LOSS_MASKED_DIFF);
  const bool dense = (flags & 1) != 0;
  if constexpr (kHasDense) {
    if (dense) {
      warp_fast_kernel<SAMPLER, PARAM, PASS, CT, LOSS, 1, (CT == 1) ? kTuneDense : kTuneDenseC3><<<grid, NT, 0, stream>>>(a);
      return launched("warp_fast_kernel");
    }
  }
  warp_fast_kernel<SAMPLER, PARAM, PASS, CT, LOSS, 0, kTuneGeneral><<<grid, NT, 0, stream>>>(a);
  return launched("warp_fast_kernel");
}

template <int SAMPLER, int PARAM, int PASS, int CT>
int launch_l(const FastArgs& a, int n, long long tiles, int loss, int flags, cudaStream_t stream) {
  switch (loss) {
    case DMH_LOSS_NONE:
      if (PASS == PASS_FUSED) break;  // a fused pass without a loss has nothing to differentiate
      return launch<SAMPLER, PARAM, PASS, CT, DMH_LOSS_NONE>(a, n, tiles, flags, stream);
    case DMH_LOSS_MASKED_DIFF: return launch<SAMPLER, PARAM, PASS, CT, DMH_LOSS_MASKED_DIFF>(a, n, tiles, flags, stream);
    case DMH_LOSS_DIFF_MASKED: return launch<SAMPLER, PARAM, PASS, CT, DMH_LOSS_DIFF_MASKED>(a, n, tiles, flags, stream);
  }
  return 1;
}

template <int SAMPLER, int PARAM, int PASS>
int launch_c(const FastArgs& a, int n, long long tiles, int C, int loss, int flags, cudaStream_t stream) {
  if (C == 1) return launch_l<SAMPLER, PARAM, PASS, 1>(a, n, tiles, loss, flags, stream);
  if (C == 3) return launch_l<SAMPLER, PARAM, PASS, 3>(a, n, tiles, loss, flags, stream);
  return 1;
}

template <int SAMPLER, int PARAM>
int launch_pass(const FastArgs& a, int n, long long tiles, int pass, int C, int loss, int flags, cudaStream_t stream) {
  switch (pass) {
    case PASS_FWD: return launch_c<SAMPLER, PARAM, PASS_FWD>(a, n, tiles, C, loss, flags, stream);
    case PASS_BWD: return launch_c<SAMPLER, PARAM, PASS_BWD>(a, n, tiles, C, loss, flags, stream);
    case PASS_FUSED: return launch_c<SAMPLER, PARAM, PASS_FUSED>(a, n, tiles, C, loss, flags, stream);
  }
  return 1;
}

}  // namespace

// Returns DMH_OK / DMH_ECUDA when it launched, 1 when the request is outside the fast path.
int warp_fast_try(const dmh_warp_desc* d, int n, int pass, cudaStream_t stream) {
  const dmh_warp_desc& d0 = d[0];
  if (n > 2) return 1;
  if (d0.C < 1) return 1;
  // C = 1 and C = 3 are the compiled channel counts; other C run as channel groups of 3 (C % 3 == 0: the Swin pyramid
  // levels, C = 12 / 24, HEM/model/swin_multi.py:161-166) or of 1
  const int CT = (d0.C % 3 == 0) ? 3 : 1;
  const int groups = d0.C / CT;
  if (d0.sampler != DMH_S1 && d0.sampler != DMH_S3_BORDER) return 1;
  if (d0.param_kind != DMH_PARAM_HOMOGRAPHY && d0.param_kind != DMH_PARAM_FLOW) return 1;
  if ((long long)CT * d0.Hs * d0.Ws >= 2147483647LL || (long long)CT * d0.h * d0.w >= 2147483647LL) return 1;
  FastArgs a;
  a.groups = groups;
  const int loss = (d0.target != nullptr) ? d0.loss_form : DMH_LOSS_NONE;
  bool dense = true;
  for (int i = 0; i < n; ++i) {
    const dmh_warp_desc& s = d[i];
    if (pass == PASS_FWD)
      dense = dense && s.out && s.valid && !s.soft_mask && loss == DMH_LOSS_NONE;
    else
      dense = dense && s.use_border_mask && !s.soft_mask && s.grad_src && s.grad_target && s.grad_param &&
              !s.grad_soft_mask && !s.grad_out && (pass == PASS_BWD || s.loss_acc) && loss == DMH_LOSS_MASKED_DIFF;
    if (s.start || s.flow_out || s.indices) return 1;
    if (groups > 1 && (s.grad_soft_mask || s.C != d0.C)) return 1;   // (dL/dsoft_mask is a plain store per sample)
    if (s.param_kind == DMH_PARAM_HOMOGRAPHY && s.divide != 1) return 1;
    if (s.loss_form != d0.loss_form || s.start_x != d0.start_x || s.start_y != d0.start_y) return 1;
    if (pass == PASS_FWD && !s.out && !s.valid && !(s.loss_form != DMH_LOSS_NONE && s.loss_acc)) return 1;
    FastTerm& t = a.t[i];
    t.src = s.src; t.param = s.param; t.target = s.target; t.soft_mask = s.soft_mask;
    t.grad_out = s.grad_out; t.grad_loss = s.grad_loss; t.sample_weight = s.sample_weight;
    t.out = s.out; t.valid = s.valid; t.loss_acc = s.loss_acc;
    t.grad_src = s.grad_src; t.grad_target = s.grad_target; t.grad_param = s.grad_param;
    t.grad_soft_mask = s.grad_soft_mask;
    t.grad_loss_scale = s.grad_loss_scale;
    t.use_border_mask = s.use_border_mask;
  }
  if (n == 1) a.t[1] = a.t[0];
  a.B = d0.B; a.Hs = d0.Hs; a.Ws = d0.Ws; a.h = d0.h; a.w = d0.w;
  a.sx = d0.start_x; a.sy = d0.start_y;
  a.tiles_x = (d0.w + TW - 1) / TW;
  a.tiles_y = (d0.h + TH - 1) / TH;
  const long long tiles = (long long)a.tiles_x * a.tiles_y * d0.B * groups;
  if (tiles > 2147483647LL) return 1;
  // ---- the persistent TMA tile kernel takes the dense S1 homography launches (dmh_warp_tile.cu) ----
  const int tile_mode = tuning().tile;
  // tuning "tile": 0 never | 1 C = 1 | 2 also the gradient-free C = 3 launches | 3 also the C = 3 training launch (default:
  // with the pair-major tile order it beats the scalar kernel below, 0.95 vs 1.13 ms on 128 pairs 3x512x512)
  if (d0.sampler == DMH_S1 && d0.param_kind == DMH_PARAM_HOMOGRAPHY && pass != PASS_BWD && tile_mode > 0 &&
      (d0.C == 1 || tile_mode > (pass == PASS_FUSED ? 2 : 1))) {
    int mode = -1;
    for (int i = 0; i < n; ++i) {
      const dmh_warp_desc& s = d[i];
      int m = 0;
      const bool aligned = ((reinterpret_cast<uintptr_t>(s.src) | reinterpret_cast<uintptr_t>(s.target) |
                             reinterpret_cast<uintptr_t>(s.out) | reinterpret_cast<uintptr_t>(s.grad_target)) & 15) == 0;
      const bool l1 = loss == DMH_LOSS_MASKED_DIFF && s.target && s.loss_acc && s.use_border_mask;
      if (!aligned || s.soft_mask || s.grad_out || (loss != DMH_LOSS_NONE && !l1)) {
        m = 0;
      } else if (pass == PASS_FUSED) {
        m = (l1 && s.grad_src && s.grad_target && s.grad_param && !s.grad_soft_mask) ? 6 : 0;
      } else if (s.out && s.valid) {
        m = 1 | (l1 ? 2 : 0);
      } else if (!s.out && !s.valid && l1) {
        m = 2;
      }
      mode = (i == 0 || m == mode) ? m : 0;
      if (mode == 0) break;
    }
    if (mode > 0) {
      FastArgs at = a;
      const int rc = warp_tile_launch(at, n, mode, d0.C, false, stream);
      if (rc != 1) return rc;
    }
  }
  // ... and the C = 1 launches that warp by an explicit flow tensor (get_warp_flow(img, flow) and its backward, the fused
  // loss on the reference's own basis flow): the flow tile is staged next to the source window
  if (d0.sampler == DMH_S1 && d0.param_kind == DMH_PARAM_FLOW && d0.C == 1 && tile_mode > 0 && tuning().tile_flow > 0) {
    int mode = -1;
    for (int i = 0; i < n; ++i) {
      const dmh_warp_desc& s = d[i];
      int m = 0;
      const bool aligned = ((reinterpret_cast<uintptr_t>(s.src) | reinterpret_cast<uintptr_t>(s.param) | reinterpret_cast<uintptr_t>(s.target) |
                             reinterpret_cast<uintptr_t>(s.out) | reinterpret_cast<uintptr_t>(s.grad_target) |
                             reinterpret_cast<uintptr_t>(s.grad_param) | reinterpret_cast<uintptr_t>(s.grad_out)) & 15) == 0;
      const bool l1 = loss == DMH_LOSS_MASKED_DIFF && s.target && s.loss_acc && s.use_border_mask;
      if (!aligned || s.soft_mask || s.grad_soft_mask) {
        m = 0;
      } else if (pass == PASS_FUSED) {
        m = (l1 && s.grad_src && s.grad_target && s.grad_param && !s.grad_out) ? 6 : 0;
      } else if (pass == PASS_FWD) {
        m = (s.out && loss == DMH_LOSS_NONE) ? 1 : 0;
      } else {
        m = (s.grad_out && loss == DMH_LOSS_NONE && !s.grad_loss && (s.grad_src || s.grad_param)) ? 12 : 0;
      }
      mode = (i == 0 || m == mode) ? m : 0;
      if (mode == 0) break;
    }
    if (mode > 0) {
      FastArgs at = a;
      const int rc = warp_tile_launch(at, n, mode, d0.C, true, stream);
      if (rc != 1) return rc;
    }
  }
  if (groups > 1 && pass == PASS_FWD && d0.param_kind == DMH_PARAM_FLOW && loss == DMH_LOSS_NONE && (tuning().channels & 1)) {
    bool plain = true;
    for (int i = 0; i < n; ++i) plain = plain && d[i].out && !d[i].soft_mask;
    if (plain) {
      const dim3 g = plane_grid((long long)d0.h * d0.w, d0.B, NT);
      const dim3 grid(g.x, g.y, (unsigned)n);
      if (d0.sampler == DMH_S1)
        warp_flow_channels_fwd_kernel<DMH_S1><<<grid, NT, 0, stream>>>(a, d0.C);
      else
        warp_flow_channels_fwd_kernel<DMH_S3_BORDER><<<grid, NT, 0, stream>>>(a, d0.C);
      return launched("warp_flow_channels_fwd_kernel");
    }
  }
  if (groups > 1 && pass != PASS_FWD && d0.param_kind == DMH_PARAM_FLOW) {
    // dL/dflow of a sample is the sum over its channel groups: accumulated with REDs into a zeroed buffer
    for (int i = 0; i < n; ++i)
      if (d[i].grad_param &&
          cudaMemsetAsync(d[i].grad_param, 0, (size_t)d0.B * 2 * d0.h * d0.w * sizeof(float), stream) != cudaSuccess)
        return fail(DMH_ECUDA, "warp: cudaMemsetAsync(grad_param) failed");
  }
  if (d0.sampler == DMH_S1) {
    if (d0.param_kind == DMH_PARAM_HOMOGRAPHY)
      return launch_pass<DMH_S1, DMH_PARAM_HOMOGRAPHY>(a, n, tiles, pass, CT, loss, dense ? 1 : 0, stream);
    return launch_pass<DMH_S1, DMH_PARAM_FLOW>(a, n, tiles, pass, CT, loss, dense ? 1 : 0, stream);
  }
  if (d0.param_kind == DMH_PARAM_FLOW)
    return launch_pass<DMH_S3_BORDER, DMH_PARAM_FLOW>(a, n, tiles, pass, CT, loss, dense ? 1 : 0, stream);
  return 1;
}

}  // namespace dmh
