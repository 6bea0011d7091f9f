// Two-rows-per-thread ("paired") form of the dense S1 warp kernels: the B200-specific step.
//
// sm_100 adds packed fp32 arithmetic (PTX fma.rn.f32x2 -> SASS FFMA2): one issue slot, two IEEE
// fp32 results.  The warp path is issue-bound at C = 1 (profiles/r1_ncu_v5_summary.txt: IPC 3.1,
// 78 % issue slots busy, ~200 instructions per 25 algorithmic bytes), and ~40 % of those
// instructions are separately rounded fp32 mul / add chains the reference's semantics force
// (SURVEY.md App. A.2-A.3).  Here a thread owns rows (y, y+1) of its column and every such chain
// runs once on a float2 (lo = row y, hi = row y+1).
//
// Bit-exactness: ptxas contracts mul.rn.f32x2 + add.rn.f32x2 into one FFMA2 and even folds
// fma(a, 1, c) / fma(a, b, -0) chains (observed with CUDA 12.9, also under -fmad=false), which
// would change the reference's rounding.  Every exactly-rounded packed op is therefore an explicit
// fma.rn.f32x2 whose identity operand comes from a kernel parameter the compiler cannot see through:
//   a + b = fma(a, ONE, b)    a * b = fma(a, b, NEG_ZERO)    a - b = fma(b, MINUS_ONE, a)
// each of which is one correctly rounded IEEE operation, bit-identical to __fadd_rn / __fmul_rn /
// __fsub_rn (tests/test_gpu_parity.py and tests/test_gpu_golden.py demand bit-equality against the
// CPU oracle and against the general kernel).
//
// Covered: sampler S1, one homography per sample or an explicit flow, C in {1, 3}, the "dense"
// profile of dmh_warp_fast.cu - forward (out + validity mask) and forward+gradients (border mask,
// |m*t - m*w|, gradients to source, target and H / flow).  Same tile / column-strip decomposition,
// vertical tap merging (now also inside the pair) and per-CTA reductions as dmh_warp_fast.cu.
#include "dmh_common.cuh"
#include "dmh_warp_fast.h"

#include <cstdlib>

namespace dmh {

namespace {

constexpr int NT = 256;
constexpr int WX = 2, WY = 4, RPT = 16;
constexpr int TW = 32 * WX, TH = WY * RPT;
enum { PASS_FWD = 0, PASS_BWD = 1, PASS_FUSED = 2 };

typedef unsigned long long u64;
__device__ __forceinline__ u64 pk(float2 a) { return *reinterpret_cast<u64*>(&a); }
__device__ __forceinline__ float2 up(u64 a) { return *reinterpret_cast<float2*>(&a); }
__device__ __forceinline__ float2 fma2(float2 a, float2 b, float2 c) {
  u64 r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(pk(a)), "l"(pk(b)), "l"(pk(c)));
  return up(r);
}
__device__ __forceinline__ float2 splat(float v) { return make_float2(v, v); }

__device__ __forceinline__ float ldg_f(const float* base, unsigned off) {
  float v;
  asm("ld.global.nc.f32 %0, [%1];" : "=f"(v) : "l"(base + off));
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
// sign(u) * g, 0 where u == 0
__device__ __forceinline__ float signed_by(float g, float u) {
  const float s = __int_as_float(__float_as_int(g) ^ (__float_as_int(u) & 0x80000000));
  return (u == 0.f) ? 0.f : s;
}

// S1 taps of one coordinate (HEM/model/utils.py:463-490): floor, +1, clamp both to the source.
struct TapIdx {
  int ia, ib, ic, id;
  float x0f, x1f, y0f, y1f;
};
__device__ __forceinline__ TapIdx s1_taps(float cx, float cy, int Wm1, int Hm1, int Ws) {
  const int xt = max(min(__float2int_rd(cx), Wm1), -1);
  const int yt = max(min(__float2int_rd(cy), Hm1), -1);
  const int x0 = max(xt, 0), x1 = min(xt + 1, Wm1);
  const int y0 = max(yt, 0), y1 = min(yt + 1, Hm1);
  TapIdx t;
  const int r0 = y0 * Ws, r1 = y1 * Ws;
  t.ia = r0 + x0; t.ib = r1 + x0; t.ic = r0 + x1; t.id = r1 + x1;
  t.x0f = (float)x0; t.x1f = (float)x1; t.y0f = (float)y0; t.y1f = (float)y1;
  return t;
}

template <int PARAM, int PASS, int CT, int MINB, int PF = 3>
__global__ void __launch_bounds__(NT, MINB) warp_pair_kernel(const __grid_constant__ FastArgs a) {
  constexpr bool kGrad = (PASS == PASS_FUSED);
  constexpr bool kFwd = (PASS == PASS_FWD);
  // opaque identities (see the header comment)
  const float2 K1 = splat(a.one), KN0 = splat(a.neg_zero), KM1 = splat(a.minus_one);
#define ADD2(p, q) fma2((p), K1, (q))
#define MUL2(p, q) fma2((p), (q), KN0)
#define SUB2(p, q) fma2((q), KM1, (p))

  const FastTerm tm = (blockIdx.y == 0) ? a.t[0] : a.t[1];
  const int h = a.h, w = a.w, Hs = a.Hs, Ws = a.Ws;
  int t = blockIdx.x;
  const int per = a.tiles_x * a.tiles_y;
  const int b = t / per;
  t -= b * per;
  const int tyi = t / a.tiles_x, txi = t - tyi * a.tiles_x;
  const int lane = threadIdx.x & 31, wrp = threadIdx.x >> 5;
  const int x = txi * TW + (wrp % WX) * 32 + lane;
  const int y_begin = tyi * TH + (wrp / WX) * RPT;
  const int y_end = min(y_begin + RPT, h);
  const bool col_live = x < w;
  const unsigned plane_o = (unsigned)(h * w), plane_s = (unsigned)(Hs * Ws);
  const int Wm1 = Ws - 1, Hm1 = Hs - 1;

  const float* src = pin(tm.src + (size_t)b * CT * plane_s);
  const float* tgt = kGrad ? pin(tm.target + (size_t)b * CT * plane_o) : nullptr;
  const float* flow = (PARAM == DMH_PARAM_FLOW) ? pin(tm.param + (size_t)b * 2 * plane_o) : nullptr;
  float* out = kFwd ? pin(tm.out + (size_t)b * CT * plane_o) : nullptr;
  uint8_t* valid = kFwd ? pin(tm.valid + (size_t)b * plane_o) : nullptr;
  float* gsrc = kGrad ? pin(tm.grad_src + (size_t)b * CT * plane_s) : nullptr;
  float* gtgt = kGrad ? pin(tm.grad_target + (size_t)b * CT * plane_o) : nullptr;
  float* gflow = (kGrad && PARAM == DMH_PARAM_FLOW) ? pin(tm.grad_param + (size_t)b * 2 * plane_o) : nullptr;

  const float sx = a.sx, sy = a.sy;
  const float xf = (float)x;
  const float gx = add_rn(xf, sx);
  const float wf = (float)w, hf = (float)h;
  const float2 gx2 = splat(gx), xf2 = splat(xf), sy2 = splat(sy);

  float hm[9];
  float2 h0x2 = splat(0.f), h3x2 = splat(0.f), h6x2 = splat(0.f);
  if (PARAM == DMH_PARAM_HOMOGRAPHY) {
#pragma unroll
    for (int k = 0; k < 9; ++k) hm[k] = __ldg(tm.param + (size_t)b * 9 + k);
    h0x2 = splat(mul_rn(hm[0], gx));
    h3x2 = splat(mul_rn(hm[3], gx));
    h6x2 = splat(mul_rn(hm[6], gx));
  }

  float gscale = 0.f;
  if (kGrad) {
    gscale = tm.grad_loss_scale;
    if (tm.sample_weight) gscale *= __ldg(tm.sample_weight + b);
  }

  float lsum = 0.f;
  float2 sa = splat(0.f), say = splat(0.f), sb = splat(0.f), sby = splat(0.f), sc = splat(0.f), scy = splat(0.f);
  int p_ib = -1, p_id = -1;
  float pB[CT], pD[CT];
#pragma unroll
  for (int c = 0; c < CT; ++c) pB[c] = pD[c] = 0.f;

  if (col_live) {
    for (int y = y_begin; y < y_end; y += 2) {
      const bool live_b = (y + 1 < y_end);          // false only on the last row of an odd-height strip
      const int yb = live_b ? y + 1 : y;            // the dead lane recomputes row y and is discarded
      const unsigned po_a = (unsigned)(y * w + x), po_b = (unsigned)(yb * w + x);
      const float2 yf2 = make_float2((float)y, (float)yb);
      const float2 gy2 = ADD2(yf2, sy2);

      // ---- sampling coordinates of both rows ------------------------------------------------
      float2 fx2, fy2, qx2 = splat(0.f), qy2 = splat(0.f), qT2 = splat(1.f);
      if (PARAM == DMH_PARAM_FLOW) {
        fx2 = make_float2(ldg_f(flow, po_a), ldg_f(flow, po_b));
        fy2 = make_float2(ldg_f(flow, po_a + plane_o), ldg_f(flow, po_b + plane_o));
      } else {
        // (h0*x + h1*y) + h2, every product and sum rounded separately (App. A.2)
        const float2 qX2 = ADD2(ADD2(h0x2, MUL2(splat(hm[1]), gy2)), splat(hm[2]));
        const float2 qY2 = ADD2(ADD2(h3x2, MUL2(splat(hm[4]), gy2)), splat(hm[5]));
        qT2 = ADD2(ADD2(h6x2, MUL2(splat(hm[7]), gy2)), splat(hm[8]));
        if (!(fabsf(qT2.x) >= 1e-7f)) qT2.x = add_rn(qT2.x, 1e-6f);
        if (!(fabsf(qT2.y) >= 1e-7f)) qT2.y = add_rn(qT2.y, 1e-6f);
        qx2 = make_float2(div_rn(qX2.x, qT2.x), div_rn(qX2.y, qT2.y));
        qy2 = make_float2(div_rn(qY2.x, qT2.x), div_rn(qY2.y, qT2.y));
        fx2 = SUB2(qx2, gx2);
        fy2 = SUB2(qy2, gy2);
      }
      const float2 cx2 = ADD2(gx2, fx2), cy2 = ADD2(gy2, fy2);

      // ---- M1 validity mask on fl(flow + grid) (no start), inclusive bounds w, h --------------
      const float2 mx2 = ADD2(fx2, xf2), my2 = ADD2(fy2, yf2);
      const bool m1a = (mx2.x >= 0.f) && (mx2.x <= wf) && (my2.x >= 0.f) && (my2.x <= hf);
      const bool m1b = (mx2.y >= 0.f) && (mx2.y <= wf) && (my2.y >= 0.f) && (my2.y <= hf);
      if (kFwd) {
        stg_u8(valid, po_a, m1a ? 1 : 0);
        if (live_b) stg_u8(valid, po_b, m1b ? 1 : 0);
      }
      const float2 m2 = make_float2(m1a ? 1.f : 0.f, (m1b && live_b) ? 1.f : 0.f);

      // ---- taps -------------------------------------------------------------------------------
      const TapIdx ta = s1_taps(cx2.x, cy2.x, Wm1, Hm1, Ws), tb = s1_taps(cx2.y, cy2.y, Wm1, Hm1, Ws);
      const float2 ax1 = SUB2(make_float2(ta.x1f, tb.x1f), cx2), ax0 = SUB2(cx2, make_float2(ta.x0f, tb.x0f));
      const float2 ay1 = SUB2(make_float2(ta.y1f, tb.y1f), cy2), ay0 = SUB2(cy2, make_float2(ta.y0f, tb.y0f));
      const float2 wa = MUL2(ax1, ay1), wb = MUL2(ax1, ay0), wc = MUL2(ax0, ay1), wd = MUL2(ax0, ay0);

      {  // pull the lines the pair two iterations ahead will need into L1
        const unsigned pt = min(po_b + (unsigned)(PF * w), plane_o - 1), pt2 = min(po_b + (unsigned)((PF + 1) * w), plane_o - 1);
        const unsigned ps = min((unsigned)(tb.ib + PF * Ws), plane_s - 1), ps2 = min((unsigned)(tb.ib + (PF + 1) * Ws), plane_s - 1);
#pragma unroll
        for (int c = 0; c < CT; ++c) {
          if (kGrad) {
            prefetch_l1(tgt + ((unsigned)c * plane_o + pt));
            prefetch_l1(tgt + ((unsigned)c * plane_o + pt2));
          }
          prefetch_l1(src + ((unsigned)c * plane_s + ps));
          prefetch_l1(src + ((unsigned)c * plane_s + ps2));
        }
      }

      float2 gcx = splat(0.f), gcy = splat(0.f);
      float2 cA[CT], cB[CT], cC[CT], cD[CT];
#pragma unroll
      for (int c = 0; c < CT; ++c) {
        const unsigned cs = (unsigned)c * plane_s, co = (unsigned)c * plane_o;
        const float2 Ia = make_float2(ldg_f(src, cs + ta.ia), ldg_f(src, cs + tb.ia));
        const float2 Ib = make_float2(ldg_f(src, cs + ta.ib), ldg_f(src, cs + tb.ib));
        const float2 Ic = make_float2(ldg_f(src, cs + ta.ic), ldg_f(src, cs + tb.ic));
        const float2 Id = make_float2(ldg_f(src, cs + ta.id), ldg_f(src, cs + tb.id));
        // output = wa*Ia + wb*Ib + wc*Ic + wd*Id, left to right, no FMA (utils.py:523)
        const float2 wv = ADD2(ADD2(ADD2(MUL2(wa, Ia), MUL2(wb, Ib)), MUL2(wc, Ic)), MUL2(wd, Id));
        if (kFwd) {
          stg_f(out, co + po_a, wv.x);
          if (live_b) stg_f(out, co + po_b, wv.y);
        }
        if (kGrad) {
          const float2 tv = make_float2(ldg_f(tgt, co + po_a), ldg_f(tgt, co + po_b));
          const float2 u = SUB2(MUL2(m2, tv), MUL2(m2, wv));      // |m*t - m*w| (losses.py:142-146)
          lsum += fabsf(u.x) + fabsf(u.y);
          // d/dt = +gm*sign(u), d/dw = -gm*sign(u)
          const float2 gt = make_float2(signed_by(gscale * m2.x, u.x), signed_by(gscale * m2.y, u.y));
          red_f(gtgt, co + po_a, gt.x);
          if (live_b) red_f(gtgt, co + po_b, gt.y);
          const float2 go = make_float2(-gt.x, -gt.y);
          cA[c] = fma2(wa, go, KN0); cB[c] = fma2(wb, go, KN0); cC[c] = fma2(wc, go, KN0); cD[c] = fma2(wd, go, KN0);
          // d out / d cx = ay1*(Ic-Ia) + ay0*(Id-Ib);  d out / d cy = ax1*(Ib-Ia) + ax0*(Id-Ic)
          const float2 dca = SUB2(Ic, Ia), ddb = SUB2(Id, Ib), dba = SUB2(Ib, Ia), ddc = SUB2(Id, Ic);
          gcx = fma2(go, fma2(ay1, dca, fma2(ay0, ddb, KN0)), gcx);
          gcy = fma2(go, fma2(ax1, dba, fma2(ax0, ddc, KN0)), gcy);
        }
      }

      if (kGrad) {
        // ---- scatter with vertical merging: pending(prev pair, row b) | row a | row b ----------
        const bool same_p = (p_ib == ta.ia) && (p_id == ta.ic);
        if (!same_p && p_ib >= 0) {
#pragma unroll
          for (int c = 0; c < CT; ++c) {
            red_f(gsrc, (unsigned)c * plane_s + (unsigned)p_ib, pB[c]);
            red_f(gsrc, (unsigned)c * plane_s + (unsigned)p_id, pD[c]);
          }
        }
        const bool same_m = (ta.ib == tb.ia) && (ta.id == tb.ic);
        if (!same_m) {
#pragma unroll
          for (int c = 0; c < CT; ++c) {
            red_f(gsrc, (unsigned)c * plane_s + (unsigned)ta.ib, cB[c].x);
            red_f(gsrc, (unsigned)c * plane_s + (unsigned)ta.id, cD[c].x);
          }
        }
#pragma unroll
        for (int c = 0; c < CT; ++c) {
          const unsigned cs = (unsigned)c * plane_s;
          red_f(gsrc, cs + (unsigned)ta.ia, cA[c].x + (same_p ? pB[c] : 0.f));
          red_f(gsrc, cs + (unsigned)ta.ic, cC[c].x + (same_p ? pD[c] : 0.f));
          red_f(gsrc, cs + (unsigned)tb.ia, cA[c].y + (same_m ? cB[c].x : 0.f));
          red_f(gsrc, cs + (unsigned)tb.ic, cC[c].y + (same_m ? cD[c].x : 0.f));
          pB[c] = cB[c].y;
          pD[c] = cD[c].y;
        }
        p_ib = tb.ib;
        p_id = tb.id;

        if (PARAM == DMH_PARAM_FLOW) {
          stg_f(gflow, po_a, gcx.x);
          stg_f(gflow, po_a + plane_o, gcy.x);
          if (live_b) {
            stg_f(gflow, po_b, gcx.y);
            stg_f(gflow, po_b + plane_o, gcy.y);
          }
        } else {
          // flow = q/T' - g  =>  dL/dX = gcx/T', dL/dY = gcy/T', dL/dT = -(gcx*X + gcy*Y)/T'^2
          const float2 rT = make_float2(rcp_fast(qT2.x), rcp_fast(qT2.y));
          const float2 ga = fma2(gcx, rT, KN0), gb = fma2(gcy, rT, KN0);
          const float2 gcn = fma2(ga, qx2, fma2(gb, qy2, KN0));   // = -dL/dT; the sign is applied once at the end
          sa = fma2(ga, K1, sa); say = fma2(ga, gy2, say);
          sb = fma2(gb, K1, sb); sby = fma2(gb, gy2, sby);
          sc = fma2(gcn, K1, sc); scy = fma2(gcn, gy2, scy);
        }
      }
    }
    if (kGrad && p_ib >= 0) {
#pragma unroll
      for (int c = 0; c < CT; ++c) {
        red_f(gsrc, (unsigned)c * plane_s + p_ib, pB[c]);
        red_f(gsrc, (unsigned)c * plane_s + p_id, pD[c]);
      }
    }
  }
#undef ADD2
#undef MUL2
#undef SUB2

  // ---- per-CTA reductions: warp shuffle -> shared -> one atomic per value ---------------------
  if (!kGrad) return;
  constexpr bool kH = (PARAM == DMH_PARAM_HOMOGRAPHY);
  __shared__ float red[NT / 32][10];
  float v[10];
  const float s_a = sa.x + sa.y, s_b = sb.x + sb.y, s_c = -(sc.x + sc.y);
  v[0] = s_a * gx; v[1] = say.x + say.y; v[2] = s_a;
  v[3] = s_b * gx; v[4] = sby.x + sby.y; v[5] = s_b;
  v[6] = s_c * gx; v[7] = -(scy.x + scy.y); v[8] = s_c;
  v[9] = lsum;
#pragma unroll
  for (int k = 0; k < 10; ++k) {
    if (k == 9 || kH) {
      const float s = warp_sum(v[k]);
      if (lane == 0) red[wrp][k] = s;
    }
  }
  __syncthreads();
  if (threadIdx.x < 10) {
    const int k = threadIdx.x;
    if (k == 9 || kH) {
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

template <int PARAM, int PASS>
int launch_pair_c(const FastArgs& a, int n, long long tiles, int C, cudaStream_t stream) {
  dim3 grid((unsigned)tiles, (unsigned)n, 1);
  if (C == 1) {
#ifdef DMH_TUNE_BUILD
    static const int minb = getenv("DMH_PAIR_MINB") ? atoi(getenv("DMH_PAIR_MINB")) : 2;
    if (minb == 3) { warp_pair_kernel<PARAM, PASS, 1, 3><<<grid, NT, 0, stream>>>(a); return launched("warp_pair_kernel"); }
    if (minb == 35) { warp_pair_kernel<PARAM, PASS, 1, 3, 5><<<grid, NT, 0, stream>>>(a); return launched("warp_pair_kernel"); }
    if (minb == 37) { warp_pair_kernel<PARAM, PASS, 1, 3, 7><<<grid, NT, 0, stream>>>(a); return launched("warp_pair_kernel"); }
    if (minb == 25) { warp_pair_kernel<PARAM, PASS, 1, 2, 5><<<grid, NT, 0, stream>>>(a); return launched("warp_pair_kernel"); }
    if (minb == 27) { warp_pair_kernel<PARAM, PASS, 1, 2, 7><<<grid, NT, 0, stream>>>(a); return launched("warp_pair_kernel"); }
#endif
    warp_pair_kernel<PARAM, PASS, 1, 3><<<grid, NT, 0, stream>>>(a);
  } else
    warp_pair_kernel<PARAM, PASS, 3, 2><<<grid, NT, 0, stream>>>(a);
  return launched("warp_pair_kernel");
}

}  // namespace

// Dense S1 forward / fused launches in paired form.  `a` is fully populated by warp_fast_try.
int warp_pair_launch(FastArgs& a, int n, long long tiles, int param_kind, int pass, int C, cudaStream_t stream) {
  a.one = 1.0f;
  a.neg_zero = -0.0f;
  a.minus_one = -1.0f;
  if (param_kind == DMH_PARAM_HOMOGRAPHY) {
    if (pass == PASS_FWD) return launch_pair_c<DMH_PARAM_HOMOGRAPHY, PASS_FWD>(a, n, tiles, C, stream);
    return launch_pair_c<DMH_PARAM_HOMOGRAPHY, PASS_FUSED>(a, n, tiles, C, stream);
  }
  if (pass == PASS_FWD) return launch_pair_c<DMH_PARAM_FLOW, PASS_FWD>(a, n, tiles, C, stream);
  return launch_pair_c<DMH_PARAM_FLOW, PASS_FUSED>(a, n, tiles, C, stream);
}

}  // namespace dmh
