/*
 * dmhomo.h - C ABI of libdmhomo.so: B200 (sm_100a) kernels for DMHomo's batched
 * homography-warp hot path.
 *
 * The reference (lhaippp/DMHomo) is pure Python and has no FFI layer; its boundary for
 * this path is a set of Python functions on torch tensors.  Each entry point below
 * replaces the chain of ATen / OpenCV / numpy calls behind one or more of those
 * functions; the citation beside each declaration is the reference interface it
 * replaces (paths relative to the reference root).  dmhomo_b200/_lib.py binds this
 * header with ctypes; INTEGRATION.md shows the reference-side stub.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer borrowed for the duration of the enqueue,
 *     except dmh_warp_desc* itself (host memory, copied at launch);
 *   - `stream` is a cudaStream_t passed as void*; work is only enqueued - no entry
 *     point synchronises, allocates device memory or touches the default stream;
 *   - tensors are dense, row-major, fp32 NCHW unless stated;
 *   - return value: DMH_OK (0) or a negative dmh_status; dmh_last_error_string()
 *     returns a thread-local description of the last failure;
 *   - buffers documented "accumulated" must be zeroed by the caller;
 *   - re-entrant; the only process-wide state is a launch counter and the explicit
 *     dmh_set_tuning() knobs (no environment variables are read); lazy CUDA
 *     initialisation (nothing happens at dlopen time, so the library may be loaded
 *     in forked data-loader workers).
 */
#ifndef DMHOMO_H_
#define DMHOMO_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DMH_ABI_VERSION 2

#if defined(__GNUC__)
#define DMH_API __attribute__((visibility("default")))
#else
#define DMH_API
#endif

typedef enum dmh_status {
  DMH_OK = 0,
  DMH_EINVAL = -1,       /* bad argument (null pointer, non-positive size, bad enum) */
  DMH_EUNSUPPORTED = -2, /* valid request this build has no kernel for */
  DMH_ECUDA = -3         /* CUDA runtime error at launch */
} dmh_status;

/* Sampling semantics. */
typedef enum dmh_sampler {
  DMH_S1 = 0,        /* transformer(): floor, clamp the 4 corner indices, weights from clamped
                        corners and raw coordinate.  HEM/model/utils.py:443-545 */
  DMH_S1B = 1,       /* WarpImages(): coordinate clamped to the source first, then S1.
                        HEM/model/utils.py:104-164 */
  DMH_S2_ZEROS = 2,  /* warp(): normalise by max(W-1,1), grid_sample(zeros, align_corners).
                        HEM/utils_operations/pixel_wise_mapping.py:55-88 */
  DMH_S3_BORDER = 3  /* flow_warp(): normalise by (W-1), grid_sample(border, align_corners).
                        HEM/dataset/data_loader.py:84-94; ddpm.py:1262-1280 */
} dmh_sampler;

/* How the sampling coordinate of output pixel (x,y) of sample b is parameterised. */
typedef enum dmh_param_kind {
  DMH_PARAM_FLOW = 0,       /* param = flow (B,2,h,w); coordinate = (grid + start) + flow.
                               get_warp_flow(), HEM/model/utils.py:548-553 */
  DMH_PARAM_COORDS = 1,     /* param = absolute coordinates (B,2,h,w).  transformer(I, vgrid),
                               warp_with_mapping(x, vgrid) */
  DMH_PARAM_HOMOGRAPHY = 2, /* param = H (B,divide*divide,3,3); flow generated per pixel as in
                               get_flow(), HEM/model/utils.py:400-440 (App. A.2), then as FLOW */
  DMH_PARAM_BASIS8 = 3      /* param = weights (B,8), `basis` = (8,2,h,w); flow = sum_k w_k basis_k
                               in the reference's sequential order.  HEM/model/net.py:808-815 */
} dmh_param_kind;

typedef enum dmh_loss_form {
  DMH_LOSS_NONE = 0,
  DMH_LOSS_MASKED_DIFF = 1, /* |m*t - m*w|  (LossL1 on mask*a, mask*b; HEM/loss/losses.py:142-146) */
  DMH_LOSS_DIFF_MASKED = 2  /* m*|w - t|    (DGM photo loss; classifier_free_guidance.py:799-802) */
} dmh_loss_form;

/*
 * One warp problem: `out[b,c,y,x] = sample(src[b,c], coord(b,x,y))`, optionally fused with
 * the validity mask, the masked L1 term against `target`, and its gradients.
 *
 * Forward products (each optional, written when non-null):
 *   out (B,C,h,w); valid (B,h,w) uint8 = M1 mask 0<=x'<=w && 0<=y'<=h on x' = fl(flow+grid)
 *   (get_gt_correspondence_mask, HEM/utils_operations/flow_and_mapping_operations.py:45-71);
 *   flow_out (B,2,h,w); indices (4,B,h,w) int32 = clamped x0,y0,x1,y1 (S1/S1B, parity aid);
 *   loss_acc (B) double, ACCUMULATED: sum over c,y,x of the loss term, unweighted.
 * Loss term: m = (use_border_mask ? M1 : 1) * (soft_mask ? soft_mask[b,0,y,x] : 1).
 * Gradients (dmh_warp_backward, or dmh_warp_forward with compute_grads=1 which assumes an
 * upstream loss gradient of 1):
 *   upstream = grad_out (B,C,h,w, may be null)  +  d(loss)/d(out) where
 *   loss = grad_loss_scale * (grad_loss ? *grad_loss : 1) * sum_b sample_weight[b] * loss_acc[b];
 *   grad_src (B,C,Hs,Ws) ACCUMULATED; grad_target (B,C,h,w) ACCUMULATED;
 *   grad_soft_mask (B,1,h,w) written; grad_param: FLOW/COORDS (B,2,h,w) written,
 *   HOMOGRAPHY (B,divide^2,3,3) ACCUMULATED, BASIS8 (B,8) ACCUMULATED.
 */
typedef struct dmh_warp_desc {
  uint32_t struct_size; /* sizeof(dmh_warp_desc), ABI guard */
  int32_t sampler;      /* dmh_sampler */
  int32_t param_kind;   /* dmh_param_kind */
  int32_t loss_form;    /* dmh_loss_form */
  int32_t B, C, Hs, Ws, h, w;
  int32_t divide;           /* HOMOGRAPHY: mesh cells per side (>=1); h,w divisible by it */
  int32_t use_border_mask;  /* multiply the loss mask by M1 */
  int32_t compute_grads;    /* forward only: also produce gradients of the loss (upstream 1) */
  int32_t reserved0;
  float start_x, start_y;   /* scalar `start` of get_grid() */
  float grad_loss_scale;    /* weight / (B*C*h*w): folds `mean` and the loss weight */
  float reserved1;
  const float* src;           /* (B,C,Hs,Ws) */
  const float* param;         /* see dmh_param_kind */
  const float* basis;         /* BASIS8 only: (8,2,h,w) */
  const float* start;         /* optional (B,2) per-sample start (overrides start_x/y) */
  const float* target;        /* (B,C,h,w) or null */
  const float* soft_mask;     /* (B,1,h,w) or null */
  const float* sample_weight; /* (B) or null (=1) */
  const float* grad_out;      /* (B,C,h,w) or null */
  const float* grad_loss;     /* device scalar or null (=1) */
  float* out;
  uint8_t* valid;
  float* flow_out;
  int32_t* indices;
  double* loss_acc;
  float* grad_src;
  float* grad_target;
  float* grad_param;
  float* grad_soft_mask;
} dmh_warp_desc;

/* Library / diagnostics. */
DMH_API int dmh_version(void);
DMH_API const char* dmh_last_error_string(void);
/* Number of kernels this library has launched in this process (bench.py's gpu_launches). */
DMH_API uint64_t dmh_launch_count(void);
/* Name of the kernel the calling thread launched last through this library ("" before the first launch);
 * bench.py labels its roofline object with it. */
DMH_API const char* dmh_last_kernel_name(void);
/* Development knobs, process-wide, read at launch time (defaults = the measured best; none changes a result):
 *   "tile"          0 scalar kernels only | 1 TMA tile kernel for the dense C = 1 launches | 2 also the gradient-free
 *                   C = 3 launches | 3 also the C = 3 training launch (default)
 *   "tile_interior" bit 0 interior-tile body, bit 1 mixed-tile body of the tile kernel (default 3)
 *   "tile_flow"     1 (default): the C = 1 launches that warp by an explicit flow go to the tile kernel, 0: scalar kernels
 *   "tile_pair_major" 1: two-term tile launches walk (sample, term, tile), 0: (term, sample, tile); default -1: 1 at C = 3
 *   "tile_dyn"      percent of a tile launch's tile list handed out dynamically, the rest is split statically
 *                   (default -1: 0 for the C = 1 training launch, 100 otherwise)
 *   "tile_chunk"    longest run of tiles per dynamic claim, 1 .. 8; runs shrink to single tiles at the end
 *                   (default -1: 1 for the C = 1 training launch, 8 otherwise)
 *   "channels"      1 (default): the pixel-per-thread forward kernel for explicit-flow warps of feature maps with C other
 *                   than 1 / 3; 0: channel groups on the tiled scalar kernel (which the backward always uses)
 *   "tile_wide"     accepted, without effect in the product library (a 24-consumer-warp geometry of experiment builds)
 * The library reads no environment variables.  Unknown key: DMH_EINVAL. */
DMH_API int dmh_set_tuning(const char* key, int value);
DMH_API int dmh_get_tuning(const char* key, int* value);

/* --- warp (A6-A9, A12-A14) -------------------------------------------------------------
 * dmh_warp_forward: get_warp_flow / transformer / WarpImages / warp / warp_with_mapping /
 * flow_warp, plus create_border_mask and the masked-L1 terms when requested.
 * `n` descriptors with identical sampler/param_kind/C/loss configuration run as ONE launch
 * (e.g. the forward and backward direction of a pair); otherwise one launch each. */
DMH_API int dmh_warp_forward(const dmh_warp_desc* descs, int n, void* stream);
DMH_API int dmh_warp_backward(const dmh_warp_desc* descs, int n, void* stream);

/* loss[0] = scale * sum_i sum_b w_i[b] * acc_i[b]  over `n_acc` accumulators of length B
 * (finishes LossL1 'mean' / the DGM per-sample weighting on the device). */
DMH_API int dmh_loss_finish(const double* const* acc, const float* const* sample_weight, int n_acc, int B,
                    float scale, float* loss, void* stream);
/* x[i] *= *g for i < n unless *g == 1 (backward of a forward-computed unit gradient). */
DMH_API int dmh_scale_inplace(float* x, int64_t n, const float* g, void* stream);

/* --- DLT (A1-A3) ------------------------------------------------------------------------
 * DLT.forward(src_pt, dst_pt, 'Axb') / WarpMat / DLT_solve: N independent 4-point 8x8
 * systems, one warp per system.  HEM/model/utils.py:55-101, 360-397; HEM/model/net.py:24-92.
 * src, dst: (N,4,2); H: (N,3,3) with H[2][2] = 1. */
DMH_API int dmh_dlt4_forward(const float* src, const float* dst, float* H, int N, void* stream);
/* grad_H (N,3,3) -> grad_dst (N,4,2) and optional grad_src (N,4,2) (written). */
DMH_API int dmh_dlt4_backward(const float* src, const float* dst, const float* H, const float* grad_H,
                      float* grad_dst, float* grad_src, int N, void* stream);

/* --- homography -> flow (A4, A5, A15) ----------------------------------------------------
 * get_flow(): fp32, reference rounding order, eps rule.  HEM/model/utils.py:400-440. */
DMH_API int dmh_homography_to_flow(const float* H, float* flow, int B, int h, int w, int divide,
                           float start_x, float start_y, const float* start /* optional (B,2) */,
                           void* stream);
/* grad_flow (B,2,h,w) -> grad_H (B,divide^2,3,3) ACCUMULATED. */
DMH_API int dmh_homography_to_flow_backward(const float* H, const float* grad_flow, float* grad_H, int B,
                                    int h, int w, int divide, float start_x, float start_y,
                                    const float* start /* optional (B,2) */, void* stream);
/* homo_to_flow()/get_flow_np() and from_homography_to_pixel_wise_mapping(): fp64 arithmetic,
 * T + eps always, rounded to fp32 at the end.  ddpm.py:913-975;
 * HEM/utils_operations/flow_and_mapping_operations.py:454-484.
 * H: (B,3,3) double.  out: channels_last ? (B,h,w,2) : (B,2,h,w).
 * as_mapping: 0 = flow q/T - grid rounded once; 1 = mapping q/T; 2 = the loaders' ground-truth flow
 * homo_convert_to_flow(H, size) = fl32(fl32(q/T) - grid) (HEM/dataset/data_loader.py:42-52, eps 1e-8). */
DMH_API int dmh_homography_to_flow_f64(const double* H, float* out, int B, int h, int w, double eps,
                               int channels_last, int as_mapping, void* stream);

/* --- data formats either side of the path ------------------------ */
/* The on-disk pair format {"img12": (6,H,W) uint8} batched as (B,6,H,W) -> what DGMTrainData.__getitem__ /
 * data_aug hand the network (HEM/dataset/data_loader.py:121-146, 217-255): gray_full (B,2,H,W) =
 * float32(mean_c((u8 - mean_c) / std_c)) in fp64 (numpy), gray_patch (B,2,ph,pw) = its crop at start[b] = (x, y)
 * (int32, (B,2)), rgb_full (B,6,H,W) = float32(u8) / 255.  Any output may be null.  W % 4 == 0.
 * patch_planar != 0: gray_patch is laid out (2,B,ph,pw) - image 1 and image 2 of every pair as two dense batches,
 * the form the warp entry points take.  mean3 / std3 are HOST pointers to three doubles. */
DMH_API int dmh_pairs_u8_to_gray(const uint8_t* img12, const int* start, float* gray_full, float* gray_patch,
                         float* rgb_full, const double* mean3, const double* std3, int B, int H, int W,
                         int patch_h, int patch_w, int patch_planar, void* stream);
/* dst[i] = fl(fl(float(src[i]) * scale) + bias), i < n: uint8 frames / grey patches shipped over PCIe at one byte
 * per pixel and expanded in HBM (the loaders' torch.Tensor(img).float() / 255, HEM/dataset/data_loader.py:139-146). */
DMH_API int dmh_u8_to_f32(const uint8_t* src, float* dst, int64_t n, float scale, float bias, void* stream);
/* normalize(tensor) / unnormalize(tensor) / unormalise_and_convert_mapping_to_flow(map)
 * (HEM/utils_operations/flow_and_mapping_operations.py:419-451, 384-416, 227-315; torch branches): src, dst (B,2,H,W);
 * mode 0: 2*t/(S-1) - 1; mode 1: (t+1)*(S-1)/2; mode 2: mode 1 minus the pixel grid; S = W for channel 0, H for channel 1. */
DMH_API int dmh_grid_normalize(const float* src, float* dst, int B, int H, int W, int mode, void* stream);
/* upsample2d_flow_as(inputs, target_as, mode="bilinear", if_rate, align_corners)
 * (HEM/model/utils.py:556-572; swin_multi.py:1175-1182): flow (B,2,hi,wi) -> out (B,2,ho,wo), bilinear with
 * torch's index / lambda rules; if_rate multiplies channel 0 by wo/wi and channel 1 by ho/hi first (the
 * reference does that in place on its input - the compat layer reproduces the side effect). */
DMH_API int dmh_flow_upsample(const float* flow, float* out, int B, int hi, int wi, int ho, int wo, int if_rate,
                      int align_corners, void* stream);
/* Adjoint of dmh_flow_upsample: grad_out (B,2,ho,wo) -> grad_flow (B,2,hi,wi) WRITTEN (gather form, no atomics). */
DMH_API int dmh_flow_upsample_backward(const float* grad_out, float* grad_flow, int B, int hi, int wi, int ho,
                               int wo, int if_rate, int align_corners, void* stream);

/* --- basis flows (A12) ------------------------------------------------------------------ */
/* flow = sum_k w_k basis_k.  basis (8,2,h,w), weight (B,8) -> flow (B,2,h,w). */
DMH_API int dmh_basis_combine(const float* basis, const float* weight, float* flow, int B, int h, int w,
                      void* stream);
/* grad_flow (B,2,h,w) -> grad_weight (B,8) ACCUMULATED. */
DMH_API int dmh_basis_combine_backward(const float* basis, const float* grad_flow, float* grad_weight, int B,
                               int h, int w, void* stream);
/* The basis flow at the 4 image corners as 4-pt offsets (B,4,2), order TL,TR,BL,BR. */
DMH_API int dmh_basis_corner_offsets(const float* basis, const float* weight, float* offsets, int B, int h,
                             int w, void* stream);
DMH_API int dmh_basis_corner_offsets_backward(const float* basis, const float* grad_offsets,
                                      float* grad_weight, int B, int h, int w, void* stream);

/* Fused cfg-2 prologue: for each of `n_sets` (<= 4) weight sets, weights[s] (B,8) -> basis flow at the 4
 * image corners -> 4-point DLT -> H[s] (B,3,3); one launch, one warp per sample.  Replaces
 * net.py:808-815 sampled at the corners + utils.py:55-101.  Backward writes grad_weight[s] (B,8). */
DMH_API int dmh_basis_homography_forward(const float* basis, const float* const* weights, float* const* H, int n_sets,
                                 int B, int h, int w, void* stream);
DMH_API int dmh_basis_homography_backward(const float* basis, const float* const* H, const float* const* grad_H,
                                  float* const* grad_weight, int n_sets, int B, int h, int w, void* stream);

/* --- masks (A10, A11) -------------------------------------------------------------------- */
/* get_gt_correspondence_mask / create_border_mask.  flow (B,2,h,w); either output may be null. */
DMH_API int dmh_border_mask(const float* flow, uint8_t* mask_u8, float* mask_f32, int B, int h, int w,
                    void* stream);
/* define_mask_zero_borders: ~(c0<=eps & c1<=eps & c2<=eps).  image (B,3,h,w). */
DMH_API int dmh_zero_border_mask(const float* image, uint8_t* mask, int B, int h, int w, float eps,
                         void* stream);

/* --- plain L1 (A13: LossL1) --------------------------------------------------------------- */
/* acc[0] += sum |a-b| (double, ACCUMULATED). */
DMH_API int dmh_l1_sum(const float* a, const float* b, int64_t n, double* acc, void* stream);
/* ga = g*scale*sign(a-b), gb = -ga (either may be null). */
DMH_API int dmh_l1_backward(const float* a, const float* b, int64_t n, const float* g, float scale, float* ga,
                    float* gb, void* stream);

/* --- DGM condition rendering (A16, A17) --------------------------------------------------- */
/* flow_to_image()/visulize_flow(): HSV wheel.  ddpm.py:1471-1502.
 * flow: in_channels_last ? (B,h,w,2) : (B,2,h,w); rgb: out_channels_last ? (B,h,w,3) : (B,3,h,w). */
DMH_API int dmh_flow_to_rgb(const float* flow, float* rgb, int B, int h, int w, float max_flow,
                    int in_channels_last, int out_channels_last, void* stream);
/* cv2.warpPerspective(img, H, (w,h)) with defaults (INTER_LINEAR, BORDER_CONSTANT 0): inverse of H
 * in fp64, 1/32-pixel fixed-point coordinates.  ddpm.py:1520-1529; data_loader.py:151.
 * src: channels_last ? (B,Hs,Ws,C) : (B,C,Hs,Ws); dst likewise with (h,w); H (B,3,3) double. */
DMH_API int dmh_warp_perspective(const float* src, const double* H, float* dst, int B, int C, int Hs, int Ws,
                         int h, int w, int channels_last, void* stream);
/* The same call on uint8 images (generate_nyps_to_single_case.py:15 warps the uint8 halves of the {"imgs","homos"}
 * sample batches): OpenCV's fixed-point remap, weights (1-fy)(1-fx) * 2^15, (sum + 2^14) >> 15; bit-identical to cv2. */
DMH_API int dmh_warp_perspective_u8(const uint8_t* src, const double* H, uint8_t* dst, int B, int C, int Hs, int Ws,
                            int h, int w, int channels_last, void* stream);
/* cv2.remap(img, map_x, map_y, INTER_LINEAR, BORDER_CONSTANT 0), batched: remap_using_correspondence_map
 * (HEM/utils_operations/pixel_wise_mapping.py:35-52; displacement = 0, map = absolute coordinates) and
 * remap_using_flow_fields (pixel_wise_mapping.py:7-32; displacement = 1, map = flow, the pixel grid is added).
 * map: (B,2,h,w) fp32 (x then y); src / dst as in dmh_warp_perspective.  Coordinates fixed to 1/32 px as OpenCV does
 * (cvRound(v * 32), integer part saturated to int16); bit-identical to cv2 for fp32 and uint8 images. */
DMH_API int dmh_remap(const float* src, const float* map, float* dst, int B, int C, int Hs, int Ws, int h, int w,
              int channels_last, int displacement, void* stream);
DMH_API int dmh_remap_u8(const uint8_t* src, const float* map, uint8_t* dst, int B, int C, int Hs, int Ws, int h, int w,
                 int channels_last, int displacement, void* stream);

/* --- evaluation metric (A18) --------------------------------------------------------------- */
/* compute_eval_results(): per sample mean over P points of min(err(p1->p2, flow_f), err(p2->p1, flow_b)),
 * err = ||dst - (src + flow[int(y), int(x)])||.  HEM/loss/losses.py:208-211, 263-296.
 * pts (B,P,2,2); flows (B,h,w,2); err (B).  flow_b may be null: forward direction only (ComputeErrFlow). */
DMH_API int dmh_eval_point_error(const float* pts, const float* flow_f, const float* flow_b, float* err, int B,
                         int P, int h, int w, void* stream);

/* --- "next" row 1: least-squares flow -> homography ----------------------------------------
 * homo_gen()/DLT_solve(pinv): all h*w pixels as correspondences.  ddpm.py:1577-1661.
 * flow (B,2,h,w) fp32 -> H (B,3,3) double.  workspace: B*45 doubles (zeroed by the call). */
DMH_API int dmh_flow_to_homography_ls(const float* flow, double* H, double* workspace, int B, int h, int w,
                              void* stream);

#ifdef __cplusplus
}
#endif
#endif /* DMHOMO_H_ */
