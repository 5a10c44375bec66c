/*
 * scnerf_b200 — C ABI of the B200-native SCNeRF volumetric-rendering hot path.
 *
 * The reference (POSTECH-CVLab/SCNeRF) has no FFI for this path: it is eager PyTorch behind plain
 * Python functions.  This header is therefore the boundary a maintainer would bind
 * from those functions (ctypes stubs: INTEGRATION.md).  Each entry point names the reference
 * interface it replaces (file:line relative to the reference root).
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless the name ends in _host;
 *   - float tensors are fp32, row-major, contiguous; indices are int64 (the reference's dtypes);
 *   - no allocation and no implicit synchronisation inside; scratch comes from the caller
 *     (`*_workspace_bytes`); all work is enqueued on `stream` (a cudaStream_t passed as void*);
 *   - return 0 on success, <0 on error; `scnerf_last_error()` gives the message (thread-local);
 *   - gradient outputs ACCUMULATE (+=) into caller-zeroed buffers, like autograd's .grad.
 */
#ifndef SCNERF_B200_H
#define SCNERF_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SCNERF_OK 0
#define SCNERF_ERR_ARG (-1)
#define SCNERF_ERR_UNSUPPORTED (-2)
#define SCNERF_ERR_CUDA (-3)
#define SCNERF_ERR_WORKSPACE (-4)

#define SCNERF_MAX_DEPTH 16

const char* scnerf_last_error(void);
/* ABI version; bumped on any signature change. */
int scnerf_abi_version(void);
/* Compute capability this build targets (100 for sm_100a) and what the current device reports. */
int scnerf_built_for_sm(void);
int scnerf_device_sm(void);

/* ------------------------------------------------------------------------------------------------
 * searchsorted — replaces torch.searchsorted at NeRF/render.py:444 and the vendored
 * NeRF/torchsearchsorted/src/cuda/searchsorted_cuda_kernel.cu:83-142 (same broadcasting: a row count
 * of 1 broadcasts).  out[r,c] = #elements of a[r,:] that are < v (left) or <= v (right).  Bit exact.
 * ---------------------------------------------------------------------------------------------- */
int scnerf_searchsorted_f32(const float* a, const float* v, int64_t* out, int64_t nrow_a,
                            int64_t nrow_v, int64_t ncol_a, int64_t ncol_v, int right, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Learnable camera — model/camera_model.py:120-206 (PinholeModelRotNoiseLearning10kRayoRayd) and
 * :209-312 (…Distortion).  NULL noise pointers mean "attribute absent / frozen at zero".
 * ---------------------------------------------------------------------------------------------- */
typedef struct scnerf_camera {
  const float* intrinsics_initial; /* [4]  fx fy cx cy */
  const float* intrinsics_noise;   /* [4] */
  const float* extrinsics_initial; /* [n_cams,9]  rot6d(6) + translation(3) */
  const float* extrinsics_noise;   /* [n_cams,9] */
  const float* ray_o_noise;        /* [gh,gw,3] */
  const float* ray_d_noise;        /* [gh,gw,3] */
  float intrinsics_noise_scale, extrinsics_noise_scale, ray_o_noise_scale, ray_d_noise_scale;
  int32_t multiplicative_noise;
  int32_t n_cams, H, W, gh, gw;
} scnerf_camera;

typedef struct scnerf_camera_grads { /* any member may be NULL (requires_grad False) */
  float* intrinsics_noise; /* [4] */
  float* extrinsics_noise; /* [n_cams,9] */
  float* ray_o_noise;      /* [gh,gw,3] */
  float* ray_d_noise;      /* [gh,gw,3] */
} scnerf_camera_grads;

/* K[4,4] and c2w[n_cams,4,4] — CameraModel.get_intrinsic / get_extrinsic (camera_model.py:166-190). */
int scnerf_camera_matrices(const scnerf_camera* cam, float* K_out, float* E_out, void* stream);

/* The camera quantities the PRD loss projects with (model/ray_dist_loss.py:51-65,113-126), for the image pair (i0, i1):
 * K4[4] = [fx_sign * fx, fy, cx, cy] (fx_sign = -1 for method "NeRF", :116-118) and E2[2,3,4] = the two camera-to-world
 * blocks; the backward accumulates d(K4), d(E2) into the camera gradients (intrinsics_noise / extrinsics_noise). */
int scnerf_camera_pair_fwd(const scnerf_camera* cam, int64_t i0, int64_t i1, float fx_sign, float* K4, float* E2,
                           void* stream);
int scnerf_camera_pair_bwd(const scnerf_camera* cam, int64_t i0, int64_t i1, float fx_sign, const float* d_K4,
                           const float* d_E2, const scnerf_camera_grads* g, void* stream);

/* Pixel -> world ray.  Replaces NeRF/get_rays.py:
 *   :93-148  get_rays_kps_use_camera        (cam != NULL, kps != NULL)
 *   :26-72   get_rays_full_image_use_camera (cam != NULL, kps == NULL: ray i is pixel (i%W, i/W))
 *   :75-90   get_rays_kps_no_camera         (cam == NULL, focal given)
 *   :5-23    get_rays_full_image_no_camera  (cam == NULL, kps == NULL)
 * Pose source (exactly one): idx[N] (per-ray camera index), idx_scalar >= 0, or `extrinsic`
 * ([4,4], or [N,4,4] when extrinsic_per_ray).  kps is [N,2] int64 (x, y); sub-pixel keypoints (SIFT / SuperGlue
 * matches fed to the PRD loss) come as kps_f32 [N,2] float instead: the direction uses the float value
 * (get_rays.py:112-123 `kps_list_expand.float()`), the ray_o / ray_d residual lookup its truncation (`.long()`, :134,140).
 * A per-ray idx outside [0, n_cams) yields NaN rays (forward) and contributes no gradient (backward). */
typedef struct scnerf_raygen_args {
  const scnerf_camera* cam; /* NULL = fixed pinhole with `focal`, H, W below */
  float focal;
  int32_t H, W;
  const int64_t* kps;
  const int64_t* idx;
  int64_t idx_scalar; /* used when idx == NULL and extrinsic == NULL */
  const float* extrinsic;
  int32_t extrinsic_per_ray;
  int64_t N;
  const float* kps_f32; /* [N,2] float (x, y); used when kps == NULL */
} scnerf_raygen_args;

int scnerf_raygen_fwd(const scnerf_raygen_args* a, float* rays_o, float* rays_d, void* stream);
/* Backward of the cam != NULL variants: d(rays_o), d(rays_d) [N,3] -> camera parameter grads
 * (autograd graph of get_rays.py:119-146 through torch.inverse, ortho2rotation, F.interpolate). */
int scnerf_raygen_bwd(const scnerf_raygen_args* a, const float* d_rays_o, const float* d_rays_d,
                      const scnerf_camera_grads* g, void* stream);

/* Per-step ray batch — NeRF/run_nerf.py:368-398 (batching mode with a camera model):
 * shuffled_ray_idx[N] are global ray ids in [0, n_train*H*W); outputs kps[N,2] = (x, y), image_idx[N] =
 * id / (H*W) (index into the train cameras) and target[N,3] = images[i_train[image_idx], y, x].
 * images: [n_images, H, W, 3] fp32 on the device. */
int scnerf_ray_batch(const int64_t* shuffled_ray_idx, int64_t N, const float* images, const int64_t* i_train,
                     int64_t n_train, int32_t H, int32_t W, int64_t* kps, int64_t* image_idx, float* target,
                     void* stream);

/* render()'s ray packing — NeRF/render.py:105-130 (+ ndc_rays :357-374 / ndc_rays_camera :376-396):
 * viewdirs = d/|d|, optional NDC (near plane 1), rays[N, 8|11] = [o, d, near, far, viewdirs].
 * fx,fy come from `cam` (learnable, differentiable) when cam != NULL, else from `focal`. */
typedef struct scnerf_rayprep_args {
  const scnerf_camera* cam;
  float focal;
  int32_t H, W;
  int32_t ndc, use_viewdirs;
  float near_, far_;
  int64_t N;
} scnerf_rayprep_args;
int scnerf_rayprep_fwd(const scnerf_rayprep_args* a, const float* rays_o, const float* rays_d,
                       float* rays /* [N, 8|11] */, void* stream);
int scnerf_rayprep_bwd(const scnerf_rayprep_args* a, const float* rays_o, const float* rays_d,
                       const float* d_rays, float* d_rays_o, float* d_rays_d /* overwrite */,
                       float* d_intrinsics_noise /* [4] += or NULL */, void* stream);

/* ------------------------------------------------------------------------------------------------
 * The field: NeRF MLP of NeRF/run_nerf_helpers.py:76-128 in nn.Linear layout (weight [out,in]).
 * ---------------------------------------------------------------------------------------------- */
typedef struct scnerf_mlp {
  int32_t D, W, input_ch, input_ch_views, skip, use_viewdirs, output_ch;
  int32_t L_pos, L_dir; /* PE frequencies: input_ch = 3+6*L_pos, input_ch_views = 3+6*L_dir */
  float* pts_w[SCNERF_MAX_DEPTH];
  float* pts_b[SCNERF_MAX_DEPTH];
  float *views_w, *views_b;     /* [W/2, W+input_ch_views] */
  float *feature_w, *feature_b; /* [W, W] */
  float *alpha_w, *alpha_b;     /* [1, W] */
  float *rgb_w, *rgb_b;         /* [3, W/2] */
  float *output_w, *output_b;   /* [output_ch, W] (use_viewdirs == 0) */
  int32_t pts_dim;              /* 0|3: xyz points; 4: NeRF++ background (x,y,z,1/r), input_ch = 4*(1+2*L_pos) */
} scnerf_mlp; /* the same struct carries parameters (read) or their gradients (+=) */

/* ------------------------------------------------------------------------------------------------
 * render_rays — NeRF/render.py:186-300: stratified sampling -> PE + coarse MLP -> raw2outputs
 * -> sample_pdf -> sort-merge -> PE + fine MLP -> raw2outputs.
 * ---------------------------------------------------------------------------------------------- */
#define SCNERF_PRECISION_FP32 0   /* fp32 CUDA-core path (exact-order reference numerics) */
#define SCNERF_PRECISION_BF16X3 1 /* tcgen05, split-bf16 (hi/lo) 3-pass: ~fp32 accuracy */
#define SCNERF_PRECISION_BF16 2   /* tcgen05, single-pass bf16, fp32 accumulate */

typedef struct scnerf_render_cfg {
  int32_t N_samples, N_importance;
  int32_t ray_cols; /* 8 or 11 */
  int32_t lindisp, white_bkgd;
  int32_t perturb;     /* >0: stratified jitter */
  float raw_noise_std; /* >0: additive sigma noise */
  int32_t retraw;
  int32_t training;    /* keep what backward needs in the workspace */
  int32_t precision;   /* SCNERF_PRECISION_* */
  uint64_t seed;       /* Philox key for draws not supplied below */
} scnerf_render_cfg;

typedef struct scnerf_render_rand { /* injected randomness (parity tests / pytest=True); NULL = draw */
  const float* t_rand; /* [N, N_samples]              U[0,1)   render.py:249 */
  const float* u;      /* [N, N_importance]           U[0,1)   render.py:429 (NULL+perturb==0: linspace) */
  const float* noise0; /* [N, N_samples]              unit noise for the coarse pass, render.py:330 */
  const float* noise1; /* [N, N_samples+N_importance] unit noise for the fine pass */
} scnerf_render_rand;

typedef struct scnerf_render_out {
  float *rgb_map, *disp_map, *acc_map; /* [N,3] [N] [N]  (fine if N_importance>0) */
  float *rgb0, *disp0, *acc0, *z_std;  /* coarse outputs + std of the importance samples (or NULL) */
  float* raw;                          /* [N, S_last, 4] if cfg.retraw */
  float* z_vals;                       /* [N, S_last] optional (NULL ok) */
  float* weights;                      /* [N, S_last] optional */
  float* depth_map;                    /* [N] optional */
} scnerf_render_out;

size_t scnerf_render_workspace_bytes(const scnerf_render_cfg* cfg, const scnerf_mlp* coarse,
                                     int64_t N);
int scnerf_render_rays_fwd(const scnerf_render_cfg* cfg, const float* rays, int64_t N,
                           const scnerf_mlp* coarse, const scnerf_mlp* fine /* NULL: reuse coarse */,
                           const scnerf_render_rand* rnd, const scnerf_render_out* out,
                           void* workspace, size_t workspace_bytes, void* stream);

typedef struct scnerf_render_grads_in { /* dL/d(outputs); NULL = zero */
  const float *rgb_map, *disp_map, *acc_map, *rgb0, *disp0, *acc0;
} scnerf_render_grads_in;

/* Backward of render_rays using the workspace a training-mode forward left behind (same cfg,
 * same `rnd`: injected noise is re-read, Philox noise is re-generated from cfg.seed).
 * d_rays [N, ray_cols] is overwritten (near/far columns get 0).  MLP grads accumulate.
 * Graph = the reference's autograd graph: no gradient through z_vals / sample_pdf
 * (render.py:274 detaches), gradient through pts -> (o, d), viewdirs, and dists*|d|. */
int scnerf_render_rays_bwd(const scnerf_render_cfg* cfg, const float* rays, int64_t N,
                           const scnerf_mlp* coarse, const scnerf_mlp* fine,
                           const scnerf_render_rand* rnd,
                           const scnerf_render_grads_in* gin, const scnerf_mlp* g_coarse,
                           const scnerf_mlp* g_fine, float* d_rays, void* workspace,
                           size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Stage-level entry points (the reference's own functions; used by the Python mirror and tests).
 * ---------------------------------------------------------------------------------------------- */
/* Embedder.embed, run_nerf_helpers.py:24-72: x[P,3] -> [P, 3+6L]. */
int scnerf_posenc_fwd(const float* x, int64_t P, int32_t L, float* out, void* stream);
/* run_network + NeRF.forward, create_nerf.py:18-32 + run_nerf_helpers.py:105-128:
 * pts[N,S,3], viewdirs[N,3]|NULL -> raw[N,S,4|output_ch]. */
size_t scnerf_field_workspace_bytes(const scnerf_mlp* m, int64_t P, int32_t training);
/* Workspace of scnerf_field_fwd for a given precision: the tensor-core precisions keep the activations in TMEM and need
 * only the packed weight image (a few MB, independent of P); fp32 needs the [P, W] activation buffers. */
size_t scnerf_field_infer_workspace_bytes(const scnerf_mlp* m, int64_t P, int32_t precision);
int scnerf_field_fwd(const scnerf_mlp* m, const float* pts, const float* viewdirs, int64_t N,
                     int64_t S, float* raw, int32_t precision, void* workspace,
                     size_t workspace_bytes, void* stream);
/* raw2outputs, render.py:302-355.  noise: already-scaled additive sigma noise or NULL. */
int scnerf_raw2outputs_fwd(const float* raw, int32_t raw_cols, const float* z, const float* rays_d,
                           int32_t rays_d_stride, const float* noise, int32_t white_bkgd, int64_t N,
                           int64_t S, float* rgb_map, float* disp_map, float* acc_map,
                           float* weights, float* depth_map, void* stream);
/* sample_pdf, render.py:417-460.  bins[N,M], weights[N,M-1], u[N,Nf] (NULL: linspace(0,1,Nf)).
 * inds (optional) receives the searchsorted(right=True) indices. */
int scnerf_sample_pdf_fwd(const float* bins, const float* weights, const float* u, int64_t N,
                          int64_t M, int64_t Nf, float* samples, int64_t* inds, void* stream);
/* sort(cat([a, b], -1)), render.py:276.  a[N,Na], b[N,Nb] -> out[N,Na+Nb] ascending. */
int scnerf_sort_merge(const float* a, const float* b, int64_t N, int64_t Na, int64_t Nb, float* out,
                      void* stream);

/* ------------------------------------------------------------------------------------------------
 * Whole training-step forward+backward with HOST inputs (the end-to-end call bench.py times):
 * pixels -> rays -> render -> loss = mse(rgb)+mse(rgb0) -> all parameter gradients.
 * Mirrors NeRF/run_nerf.py:385-506,600 (camera branch).  kps/idx/target are pinned host buffers;
 * loss_host receives the scalar after the call's stream work completes (caller synchronises).
 * ---------------------------------------------------------------------------------------------- */
typedef struct scnerf_step_io {
  const int64_t* kps_host;  /* [N,2] */
  const int64_t* idx_host;  /* [N] */
  const float* target_host; /* [N,3] */
  float* loss_host;         /* [1] */
  /* device staging the caller owns: */
  int64_t* kps_dev;
  int64_t* idx_dev;
  float* target_dev;
  float* loss_dev;
} scnerf_step_io;

size_t scnerf_train_step_workspace_bytes(const scnerf_render_cfg* cfg, const scnerf_mlp* coarse,
                                         int64_t N);
int scnerf_train_step(const scnerf_camera* cam, const scnerf_camera_grads* g_cam,
                      const scnerf_render_cfg* cfg, int32_t ndc, float near_, float far_,
                      const scnerf_mlp* coarse,
                      const scnerf_mlp* fine, const scnerf_mlp* g_coarse, const scnerf_mlp* g_fine,
                      const scnerf_step_io* io, int32_t inputs_on_host, int64_t N, void* workspace,
                      size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Optimiser step: the reference's CustomAdamOptimizer — NeRF/create_nerf.py:199-336,
 * nerfplusplus/custom_optim.py:11-147 — as ONE multi-tensor launch per <= 40 tensors.  `tensors_host` is
 * a HOST array; weight_decay applies only to tensors with decay != 0 (the reference decays by POSITION:
 * the last ray_o / ray_d / distortion parameters, create_nerf.py:222-230; the host mirror computes that).
 * step is the tensor's 1-based update count (bias correction).  vmax != NULL selects amsgrad.
 * ---------------------------------------------------------------------------------------------- */
typedef struct scnerf_adam_tensor {
  float* param; const float* grad; float* exp_avg; float* exp_avg_sq; float* max_exp_avg_sq;
  int64_t numel;
  int32_t step, decay;
} scnerf_adam_tensor;
int scnerf_adam_step(const scnerf_adam_tensor* tensors_host, int32_t n_tensors, float lr, float beta1,
                     float beta2, float eps, float weight_decay, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Projected ray distance loss — model/ray_dist_loss.py:22-246 `proj_ray_dist_loss_single`
 * for N matched keypoint pairs: rays of image 0 / image 1 ([N,3] each, un-normalised directions), matched
 * keypoints kps0 / kps1 ([N,2] float (x, y)), K4 = [fx, fy, cx, cy] (fx NEGATED by the caller for method
 * "NeRF", :116-118), E2 = [2,3,4] camera-to-world of the two images.  acc5 is 5 floats of scratch that the
 * backward re-reads; loss[1] and n_match[1] are device scalars.  train != 0: mean over errors that are finite
 * and below `threshold`; train == 0: errors clamped to `threshold`, mean over the chirality-valid matches.
 * Backward (train): ray gradients overwrite, d_K4[4] and d_E2[2,3,4] accumulate (either may be NULL).
 * ---------------------------------------------------------------------------------------------- */
int scnerf_prd_loss_fwd(const float* rays0_o, const float* rays0_d, const float* rays1_o, const float* rays1_d,
                        const float* kps0, const float* kps1, const float* K4, const float* E2, float eps,
                        float threshold, int32_t train, int64_t N, float* acc5, float* loss, float* n_match,
                        void* stream);
int scnerf_prd_loss_bwd(const float* rays0_o, const float* rays0_d, const float* rays1_o, const float* rays1_d,
                        const float* kps0, const float* kps1, const float* K4, const float* E2, float eps,
                        float threshold, int64_t N, const float* acc5, const float* d_loss, float* d_rays0_o,
                        float* d_rays0_d, float* d_rays1_o, float* d_rays1_d, float* d_K4, float* d_E2, void* stream);

/* Hardware self-test of the tcgen05 building blocks (descriptor encodings, TMEM, bulk copy):
 * D[128,N] = bf16(A[128,K]) * bf16(B[N,K])^T with fp32 accumulation, one CTA.
 * variant bit0 swaps the LBO/SBO descriptor fields (diagnostic), bit1 stages through cp.async.bulk. */
int scnerf_tc_selftest(const float* A, const float* B, float* D, int32_t N, int32_t K, int32_t variant,
                       void* workspace, size_t workspace_bytes, void* stream);

/* Diagnostics: have the next fused-forward launches record clock64() stamps of CTA 0 for the first
 * `tiles` tiles into dev_buf (NULL disables): the pipelined forward writes [tiles][10 stages][16]
 * (tools/timeline_pipe.py), the serial one [tiles][10][4] (tools/timeline.py).  Not part of the product path. */
int scnerf_debug_timeline(long long* dev_buf, int32_t tiles);

/* Diagnostics: raw tcgen05 throughput probe.  mode 0..3: cycles for iters*16 MMAs of 128x256x16 per CTA
 * (0 SS K-major, 1 TS, 2 SS MN-major/MN-major, 3 SS MN-major A + K-major B); mode = 256 + (N/8 << 4) + pattern:
 * iters*24 MMAs of 128xNx16, N in {256,128,64,32}, pattern 0 TS, 1 TS alternating accumulators, 2 SS,
 * 3 TS/SS/TS, 4 SS/TS (tools/mma_bench.py). */
int scnerf_debug_mma_bench(int32_t mode, int32_t iters, long long* dev_out, int32_t nblocks, void* stream);

/* Diagnostics (host only, no GPU needed): the compile-time slab tables the tensor-core kernels unroll.
 * which: 0 serial forward, 1 pipelined forward (3-D points), 2 pipelined forward (4-D points), 3 serial dgrad,
 * 4 pipelined dgrad.  index < 0: out = {n_slabs, n_stages, split-bf16 weight-image bytes, 0...};
 * else out = {n, acc_col, a_off, a_lo_delta, a_kind, flags, stage, pad, img_off} of slab `index`.
 * Returns 0, or SCNERF_ERR_ARG.  Used by tests/test_host_abi.py to check the plans' hazard invariants. */
int scnerf_debug_slab_plan(int32_t which, int32_t index, int64_t* out9);

/* Kernel-launch counter (bench.py's gpu_launches): number of kernels this library has launched
 * in this process since the last reset. */
int64_t scnerf_launch_count(int32_t reset);

/* Per-launch kernel times (measurement aid; not on the product path).  scnerf_kernel_timing(1) clears the log and
 * makes every launch of this library record a CUDA event before and after it on the launching stream;
 * scnerf_kernel_timing(0) stops.  scnerf_kernel_timing_report writes one line per recorded launch, in launch order —
 * "<launch site>\t<grid size>\t<milliseconds>\n" — into buf (NUL-terminated, truncated to cap) after waiting for the
 * events, and returns the number of bytes the full report needs. */
int scnerf_kernel_timing(int32_t enable);
int64_t scnerf_kernel_timing_report(char* buf, int64_t cap);

#ifdef __cplusplus
}
#endif
#endif /* SCNERF_B200_H */
