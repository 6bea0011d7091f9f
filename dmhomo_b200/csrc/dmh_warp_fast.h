// Internal interface between the general warp dispatcher (dmh_warp.cu) and the lean
// specialisations (dmh_warp_fast.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/dmhomo.h"

namespace dmh {

// The fields of dmh_warp_desc the lean kernels read, per term.
struct FastTerm {
  const float* src;
  const float* param;
  const float* target;
  const float* soft_mask;
  const float* grad_out;
  const float* grad_loss;
  const float* sample_weight;
  float* out;
  uint8_t* valid;
  double* loss_acc;
  float* grad_src;
  float* grad_target;
  float* grad_param;
  float* grad_soft_mask;
  float grad_loss_scale;
  int use_border_mask;
};

struct FastArgs {
  FastTerm t[2];
  int B, Hs, Ws, h, w;
  int tiles_x, tiles_y;
  float sx, sy;
  // opaque identities for the packed (f32x2) kernel, see dmh_warp_tile.cu
  float one, neg_zero, minus_one;
  // tiled persistent kernel (dmh_warp_tile.cu): length of the tile list, "start offsets are benign" flag
  int n_tiles, start_sane;
  int n_static, dyn_chunk, counter_slot;   // schedule: statically split prefix of the tile list, tiles per dynamic claim, counter slot
  int pair_major;               // tile list ordered (sample, term, tile) instead of (term, sample, tile)
  int interior_ok;              // bit 0: interior-tile body, bit 1: mixed (per-row-pair vote) body (dmh_set_tuning "tile_interior")
  int groups;                   // lean scalar kernel: channel groups of CT channels per sample (C = groups * CT; 1 = the whole sample)
};

// pass: 0 forward, 1 backward, 2 forward + gradients.  Returns DMH_OK / DMH_ECUDA when it
// launched, 1 when the request is outside the lean path (caller falls back to the general kernel).
int warp_fast_try(const dmh_warp_desc* descs, int n, int pass, cudaStream_t stream);

// Persistent, TMA staged, packed-fp32 form of the dense S1 launches.  mode = bits 1 (warped output + validity mask) |
// 2 (masked L1 against the target) | 4 (gradients in the same pass) | 8 (upstream gradient instead of a loss: the
// backward of a plain warp); flow_param: explicit flow tensor instead of one homography per sample.  Returns 1 when
// the shape is outside what it takes.
int warp_tile_launch(FastArgs& a, int n, int mode, int C, bool flow_param, cudaStream_t stream);

}  // namespace dmh
