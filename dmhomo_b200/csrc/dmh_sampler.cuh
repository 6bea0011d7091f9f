// Sampling semantics shared by the warp kernels: the four taps, their
// weights and the clamp gates of S1 / S1B (HEM/model/utils.py:443-545, 104-164) and of the
// grid_sample based S2 / S3 (pixel_wise_mapping.py:55-113, data_loader.py:84-94).
#pragma once
#include "dmh_common.cuh"

namespace dmh {

struct Taps {
  int ia, ib, ic, id;      // offsets inside one source plane: (y0,x0) (y1,x0) (y0,x1) (y1,x1)
  float wa, wb, wc, wd;    // bilinear weights in the same order
  float ax0, ax1, ay0, ay1;  // S1: x-x0f, x1f-x, y-y0f, y1f-y.  S2/S3: w, e, n, s
  float gate_x, gate_y;    // d(sample coord)/d(cx): 0 where a clamp is active
  bool va, vb, vc, vd;     // tap contributes (S2 zeros padding)
};

template <int SAMPLER>
__device__ __forceinline__ void make_taps(float cx, float cy, int Hs, int Ws, Taps& t, int& x0o, int& y0o,
                                          int& x1o, int& y1o) {
  t.gate_x = 1.f;
  t.gate_y = 1.f;
  t.va = t.vb = t.vc = t.vd = true;
  if (SAMPLER == DMH_S1 || SAMPLER == DMH_S1B) {
    if (SAMPLER == DMH_S1B) {
      // torch.clamp(coord, 0, size-1) before sampling (HEM/model/utils.py:108-109)
      const float mx = (float)(Ws - 1), my = (float)(Hs - 1);
      if (cx < 0.f || cx > mx) t.gate_x = 0.f;
      if (cy < 0.f || cy > my) t.gate_y = 0.f;
      cx = fminf(fmaxf(cx, 0.f), mx);
      cy = fminf(fmaxf(cy, 0.f), my);
    }
    // floor -> int32 -> +1 -> clamp both to [0, size-1] (utils.py:463-471).  The saturating
    // float->int conversion plus a first clamp to [-1, size-1] keeps `+ 1` from overflowing.
    const int xt = max(min(__float2int_rd(cx), Ws - 1), -1);
    const int yt = max(min(__float2int_rd(cy), Hs - 1), -1);
    const int x0 = max(xt, 0), x1 = min(xt + 1, Ws - 1);
    const int y0 = max(yt, 0), y1 = min(yt + 1, Hs - 1);
    const float x0f = (float)x0, x1f = (float)x1, y0f = (float)y0, y1f = (float)y1;
    t.ax1 = sub_rn(x1f, cx);
    t.ax0 = sub_rn(cx, x0f);
    t.ay1 = sub_rn(y1f, cy);
    t.ay0 = sub_rn(cy, y0f);
    t.wa = mul_rn(t.ax1, t.ay1);
    t.wb = mul_rn(t.ax1, t.ay0);
    t.wc = mul_rn(t.ax0, t.ay1);
    t.wd = mul_rn(t.ax0, t.ay0);
    const int r0 = y0 * Ws, r1 = y1 * Ws;
    t.ia = r0 + x0;
    t.ib = r1 + x0;
    t.ic = r0 + x1;
    t.id = r1 + x1;
    x0o = x0; y0o = y0; x1o = x1; y1o = y1;
  } else {
    // normalise to [-1,1] exactly as the reference does, then ATen's align_corners=True
    // un-normalisation (pixel_wise_mapping.py:79-80 / data_loader.py:80-81; App. A.4)
    const float dw = (SAMPLER == DMH_S2_ZEROS) ? (float)max(Ws - 1, 1) : (float)(Ws - 1);
    const float dh = (SAMPLER == DMH_S2_ZEROS) ? (float)max(Hs - 1, 1) : (float)(Hs - 1);
    const float nx = sub_rn(div_rn(mul_rn(2.0f, cx), dw), 1.0f);
    const float ny = sub_rn(div_rn(mul_rn(2.0f, cy), dh), 1.0f);
    const float sxf = (float)(Ws - 1) * 0.5f, syf = (float)(Hs - 1) * 0.5f;
    float ix = mul_rn(add_rn(nx, 1.0f), sxf);
    float iy = mul_rn(add_rn(ny, 1.0f), syf);
    t.gate_x = mul_rn(sxf, div_rn(2.0f, dw));
    t.gate_y = mul_rn(syf, div_rn(2.0f, dh));
    if (SAMPLER == DMH_S3_BORDER) {
      const float mx = (float)(Ws - 1), my = (float)(Hs - 1);
      if (!(ix > 0.f && ix < mx)) t.gate_x = 0.f;
      if (!(iy > 0.f && iy < my)) t.gate_y = 0.f;
      ix = fminf(fmaxf(ix, 0.f), mx);
      iy = fminf(fmaxf(iy, 0.f), my);
    }
    const float fxl = fminf(fmaxf(floorf(ix), -1.0e9f), 1.0e9f);
    const float fyl = fminf(fmaxf(floorf(iy), -1.0e9f), 1.0e9f);
    const int x0 = (int)fxl, y0 = (int)fyl, x1 = x0 + 1, y1 = y0 + 1;
    const float wx = sub_rn(ix, fxl), wy = sub_rn(iy, fyl);  // dist to west / north
    const float ex = sub_rn(1.0f, wx), sy = sub_rn(1.0f, wy);
    t.ax0 = wx; t.ax1 = ex; t.ay0 = wy; t.ay1 = sy;
    t.wa = mul_rn(sy, ex);  // nw
    t.wc = mul_rn(sy, wx);  // ne
    t.wb = mul_rn(wy, ex);  // sw
    t.wd = mul_rn(wy, wx);  // se
    const bool x0in = (x0 >= 0 && x0 <= Ws - 1), x1in = (x1 >= 0 && x1 <= Ws - 1);
    const bool y0in = (y0 >= 0 && y0 <= Hs - 1), y1in = (y1 >= 0 && y1 <= Hs - 1);
    t.va = x0in && y0in; t.vb = x0in && y1in; t.vc = x1in && y0in; t.vd = x1in && y1in;
    const int x0c = min(max(x0, 0), Ws - 1), x1c = min(max(x1, 0), Ws - 1);
    const int y0c = min(max(y0, 0), Hs - 1), y1c = min(max(y1, 0), Hs - 1);
    t.ia = y0c * Ws + x0c; t.ib = y1c * Ws + x0c; t.ic = y0c * Ws + x1c; t.id = y1c * Ws + x1c;
    x0o = x0c; y0o = y0c; x1o = x1c; y1o = y1c;
  }
}

template <int SAMPLER>
__device__ __forceinline__ float blend(const Taps& t, float Ia, float Ib, float Ic, float Id) {
  if (SAMPLER == DMH_S1 || SAMPLER == DMH_S1B) {
    // output = wa*Ia + wb*Ib + wc*Ic + wd*Id, left to right, no FMA (utils.py:523)
    return add_rn(add_rn(add_rn(mul_rn(t.wa, Ia), mul_rn(t.wb, Ib)), mul_rn(t.wc, Ic)), mul_rn(t.wd, Id));
  } else {
    // ATen: nw*nw_val + ne*ne_val + sw*sw_val + se*se_val
    return add_rn(add_rn(add_rn(mul_rn(Ia, t.wa), mul_rn(Ic, t.wc)), mul_rn(Ib, t.wb)), mul_rn(Id, t.wd));
  }
}

}  // namespace dmh
