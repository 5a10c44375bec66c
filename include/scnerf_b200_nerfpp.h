/*
 * scnerf_b200 — C ABI, NeRF++ (inverted-sphere) rows of the hot path: SURVEY.md §8 a6, a14, a15.
 * Same conventions as scnerf_b200.h (device pointers, fp32 row-major, int64 indices, caller's stream,
 * no allocation, gradient outputs accumulate unless marked "overwrite").
 *
 * The reference composes these stages in Python (nerfplusplus/ddp_train_nerf.py:421-488); the host
 * mirror in scnerf_b200/nerfplusplus/ does the same with one autograd.Function per stage.
 */
#ifndef SCNERF_B200_NERFPP_H
#define SCNERF_B200_NERFPP_H

#include "scnerf_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

/* ------------------------------------------------------------------------------------------------
 * a6 — render_ray_from_camera, nerfplusplus/nerf_sample_ray_split.py:196-258.
 * Pixel centres (+0.5), optional 2-coefficient radial distortion (:227-232), analytic K^-1
 * (:234-243), OpenCV camera (no axis flip), ray_o/ray_d residuals looked up by flat pixel index,
 * d normalised when the camera has ray_d_noise (:252-254).  depth[N] = c2w.T[2,3] (:256).
 * ---------------------------------------------------------------------------------------------- */
typedef struct scnerf_pp_raygen_args {
  const scnerf_camera* cam;
  const float* distortion_initial; /* [2] or NULL: camera class without distortion_noise */
  const float* distortion_noise;   /* [2] */
  float distortion_noise_scale;
  const int64_t* select_inds;      /* [N] flat pixel index y*W+x */
  int64_t camera_idx;              /* >= 0: camera_model(camera_idx); < 0: use `extrinsic` */
  const float* extrinsic;          /* [4,4] camera-to-world (test-time path, :210-212) */
  int64_t N;
} scnerf_pp_raygen_args;
int scnerf_pp_raygen_fwd(const scnerf_pp_raygen_args* a, float* rays_o, float* rays_d,
                         float* depth /* [N] or NULL */, void* stream);
int scnerf_pp_raygen_bwd(const scnerf_pp_raygen_args* a, const float* d_rays_o, const float* d_rays_d,
                         const scnerf_camera_grads* g, float* d_distortion_noise /* [2] or NULL */,
                         void* stream);

/* ------------------------------------------------------------------------------------------------
 * a14 — intersect_sphere, nerfplusplus/ddp_train_nerf.py:50-68.  far[N] = depth of the unit-sphere
 * exit.  *miss_count (device int, caller-zeroed) counts rays that never enter the sphere; the
 * reference raises for those, and so does the host mirror after reading the counter.
 * ---------------------------------------------------------------------------------------------- */
int scnerf_pp_intersect_sphere_fwd(const float* rays_o, const float* rays_d, int64_t N, float* far_,
                                   int32_t* miss_count, void* stream);
int scnerf_pp_intersect_sphere_bwd(const float* rays_o, const float* rays_d, const float* d_far,
                                   int64_t N, float* d_rays_o, float* d_rays_d, void* stream);

/* Level-0 depths, ddp_train_nerf.py:437-449 (+ perturb_samples :71-80 when t_* != NULL):
 *   fg[N,S] between `near_` (min_depth) and far[N]; coef[N,S] = d(fg)/d(far); bg[N,S] on linspace(0,1).
 * t_fg / t_bg: the U[0,1) draws of perturb_samples ([N,S]) or NULL for no jitter. */
int scnerf_pp_level0_depths(const float* far_, float near_, int64_t N, int64_t S, const float* t_fg,
                            const float* t_bg, float* fg, float* coef, float* bg, void* stream);

/* Same with a per-ray near depth: near_rays[N] = ray_batch['min_depth'] (ddp_train_nerf.py:438; datasets that ship
 * per-pixel min-depth maps, nerf_sample_ray_split.py:166-171). */
int scnerf_pp_level0_depths_rays(const float* far_, const float* near_rays, int64_t N, int64_t S, const float* t_fg,
                                 const float* t_bg, float* fg, float* coef, float* bg, void* stream);

/* Cascade level >= 1, ddp_train_nerf.py:451-467: sample_pdf (:83-132) on the depth mid-points with
 * weights[..., 1:-1], then sort(cat(depth, samples)).  u[N,Nf] = the torch.rand draw, NULL = det.
 * coef / merged_coef (optional) carry d(depth)/d(far) through the lerp and the sort.
 * samples[N,Nf] and above[N,Nf] (the count-based bin index, :110) are optional outputs. */
int scnerf_pp_sample_pdf(const float* depth, const float* coef, const float* weights, const float* u,
                         int64_t N, int64_t S, int64_t Nf, float* samples, int64_t* above,
                         float* merged, float* merged_coef, void* stream);
/* The reference's own signatures, for callers that compose the cascade themselves:
 *   sample_pdf(bins[N,M1], weights[N,M1-1], N_samples, det) -> samples[N,Nf]   (ddp_train_nerf.py:83-132)
 *     above[N,Nf] / tfrac[N,Nf] (optional) are what the backward w.r.t. `bins` needs (d_bins +=);
 *     there is no gradient w.r.t. weights (the reference's only caller detaches them, :452,461).
 *   perturb_samples(z_vals[N,S]) with the torch.rand_like draw passed in                 (:71-80) */
int scnerf_pp_sample_pdf_bins(const float* bins, const float* weights, const float* u, int64_t N, int64_t M1,
                              int64_t Nf, float* samples, int64_t* above, float* tfrac, void* stream);
int scnerf_pp_sample_pdf_bins_bwd(const float* d_samples, const int64_t* above, const float* tfrac, int64_t N,
                                  int64_t M1, int64_t Nf, float* d_bins, void* stream);
int scnerf_pp_perturb_samples_fwd(const float* z_vals, const float* t_rand, int64_t N, int64_t S, float* out,
                                  void* stream);
int scnerf_pp_perturb_samples_bwd(const float* d_out, const float* t_rand, int64_t N, int64_t S, float* d_z_vals,
                                  void* stream);
/* d_far[N] += sum_s d_depth[N,S] * coef[N,S] */
int scnerf_pp_depth_bwd(const float* d_depth, const float* coef, int64_t N, int64_t S, float* d_far,
                        void* stream);

/* ------------------------------------------------------------------------------------------------
 * a15 — the field.  MLPNet (nerfplusplus/nerf_network.py:68-142) has the topology of the NeRF/ MLP:
 * scnerf_mlp carries it with base_layers -> pts_*, sigma_layers -> alpha_*, base_remap_layers ->
 * feature_*, rgb_layers.0 -> views_*, rgb_layers.2 -> rgb_*; raw[.,0:3] is rgb BEFORE the sigmoid and
 * raw[.,3] sigma BEFORE the abs (both applied in scnerf_pp_composite_*).  The background network sets
 * pts_dim = 4, input_ch = 84 (its own tensor-core slab plans: 6 PE slabs, 96-wide d(PE) stages).
 *
 * Generic field evaluation that keeps what the backward needs in the caller's workspace.
 *   points: (rays[N,ray_cols], z[N,S]) -> pts = o + z d, viewdirs = rays[:,8:11]     (pts == NULL)
 *       or  explicit pts[N,S,pts_dim] + viewdirs[N,3]                                (rays == NULL)
 * precision: SCNERF_PRECISION_*; the tensor-core paths take the (rays, z) form for 3-D points and the explicit
 * form for 4-D points (the two network shapes of the reference).
 * Backward outputs (any may be NULL): d_rays[N,ray_cols] += (columns o, d, viewdirs), d_z[N,S]
 * overwrite (d_pts . d), d_pts[N,S,pts_dim] overwrite, d_viewdirs[N,3] +=.
 * ---------------------------------------------------------------------------------------------- */
/* No-grad twin of scnerf_field_train_fwd (same arguments, same raw[N,S,4]): no tile images or activations are kept, so the
 * workspace is the packed weight image only for the tensor-core precisions (render_single_image, ddp_train_nerf.py:135-256). */
size_t scnerf_field_infer_rays_workspace_bytes(const scnerf_mlp* m, int64_t N, int64_t S, int32_t precision);
int scnerf_field_infer_fwd(const scnerf_mlp* m, const float* rays, int32_t ray_cols, const float* z,
                           const float* pts, const float* viewdirs, int64_t N, int64_t S, float* raw,
                           int32_t precision, void* workspace, size_t workspace_bytes, void* stream);
size_t scnerf_field_train_workspace_bytes(const scnerf_mlp* m, int64_t N, int64_t S, int32_t precision);
int scnerf_field_train_fwd(const scnerf_mlp* m, const float* rays, int32_t ray_cols, const float* z,
                           const float* pts, const float* viewdirs, int64_t N, int64_t S, float* raw,
                           int32_t precision, void* workspace, size_t workspace_bytes, void* stream);
int scnerf_field_train_bwd(const scnerf_mlp* m, const scnerf_mlp* g_m, const float* rays,
                           int32_t ray_cols, const float* z, const float* pts, const float* viewdirs,
                           int64_t N, int64_t S, const float* d_raw, float* d_rays, float* d_z,
                           float* d_pts, float* d_viewdirs, int32_t precision, void* workspace,
                           size_t workspace_bytes, void* stream);

/* rays[N,11] = [o, d, 0, 0, d/|d|] (ddp_model.py:83-84) and its backward into d_rays_o/d_rays_d (+=). */
int scnerf_pp_pack_rays(const float* rays_o, const float* rays_d, int64_t N, float* rays, void* stream);
int scnerf_pp_pack_rays_bwd(const float* rays_d, const float* d_rays, int64_t N, float* d_rays_o,
                            float* d_rays_d, void* stream);

/* depth2pts_outside, nerfplusplus/ddp_model.py:16-45, in the flipped sample order the background MLP
 * is fed (:120): pts4[r, s] = (unit-sphere point, 1/r) for inverse depth bg_z[r, S-1-s]. */
int scnerf_pp_bg_points_fwd(const float* rays_o, const float* rays_d, const float* bg_z, int64_t N,
                            int64_t S, float* pts4, float* depth_real /* [N,S] or NULL (:44, no backward) */,
                            void* stream);
int scnerf_pp_bg_points_bwd(const float* rays_o, const float* rays_d, const float* bg_z,
                            const float* d_pts4, int64_t N, int64_t S, float* d_rays_o, float* d_rays_d,
                            void* stream);

/* Alpha compositing of NerfNet.forward, ddp_model.py:98-142.
 * fg: weights[N,S], fg_rgb[N,3], fg_depth[N], bg_lambda[N] from raw_fg, fg_z, fg_z_max, |rays_d|.
 * bg: raw_bg is in flipped order, bg_z is the UNflipped inverse depth; outputs bg_weights[N,S] (flipped
 *     order, as the reference returns them), bg_rgb[N,3] and bg_depth[N] already scaled by bg_lambda,
 *     and rgb[N,3] = fg_rgb + bg_rgb. */
int scnerf_pp_composite_fg_fwd(const float* raw, const float* fg_z, const float* fg_z_max,
                               const float* rays_d, int64_t N, int64_t S, float* weights, float* fg_rgb,
                               float* fg_depth, float* bg_lambda, void* stream);
int scnerf_pp_composite_fg_bwd(const float* raw, const float* fg_z, const float* fg_z_max,
                               const float* rays_d, int64_t N, int64_t S, const float* d_fg_rgb,
                               const float* d_bg_lambda /* or NULL */, float* d_raw /* overwrite */,
                               float* d_fg_z /* overwrite */, float* d_fg_z_max /* += */,
                               float* d_rays_d /* += */, void* stream);
int scnerf_pp_composite_bg_fwd(const float* raw, const float* bg_z, const float* bg_lambda,
                               const float* fg_rgb, int64_t N, int64_t S, float* weights, float* bg_rgb,
                               float* bg_depth, float* rgb, void* stream);
int scnerf_pp_composite_bg_bwd(const float* raw, const float* bg_z, const float* bg_lambda, int64_t N,
                               int64_t S, const float* d_rgb, float* d_raw /* overwrite */,
                               float* d_bg_lambda /* overwrite */, void* stream);

/* ------------------------------------------------------------------------------------------------
 * One whole optimisation step's forward + backward of nerfplusplus/ddp_train_nerf.py:421-488,552 in ONE call:
 * pixel indices -> render_ray_from_camera -> intersect_sphere -> for each cascade level { depths (level 0:
 * :437-449; level >= 1: sample_pdf + sort, :451-467) -> NerfNet (fg field, fg composite, depth2pts_outside,
 * bg field, bg composite) -> img2mse } -> loss = sum of the levels' losses -> every gradient: both networks of
 * every level (written into g_nets, accumulate), the camera (g_cam / g_distortion, accumulate).
 * No auto-exposure (optim_autoexpo = False), no PRD term (add it with scnerf_prd_loss_*; the camera gradients
 * accumulate).  Rays are generated once: the reference regenerates identical rays per level (:423-428).
 * ---------------------------------------------------------------------------------------------- */
typedef struct scnerf_pp_step_cfg {
  int32_t cascade_level;       /* 1 or 2 */
  int32_t cascade_samples[2];  /* samples ADDED at each level: [64, 128] -> 64 at level 0, 192 at level 1 */
  int32_t precision;           /* SCNERF_PRECISION_* */
  int32_t perturb;             /* 1: perturb_samples + random u (training); 0: deterministic (evaluation) */
  float min_depth;             /* fg near depth when io->min_depth_dev == NULL (1e-4) */
  uint64_t seed;               /* counter RNG seed for the draws that are not injected */
} scnerf_pp_step_cfg;
typedef struct scnerf_pp_step_rand { /* injected draws (parity tests); NULL = draw from `seed` */
  const float* t_fg; /* [N, cascade_samples[0]] perturb_samples(fg) */
  const float* t_bg; /* [N, cascade_samples[0]] perturb_samples(bg) */
  const float* u_fg; /* [N, cascade_samples[1]] sample_pdf(fg) */
  const float* u_bg; /* [N, cascade_samples[1]] sample_pdf(bg) */
} scnerf_pp_step_rand;
typedef struct scnerf_pp_nets { /* net_m = NerfNet of cascade level m: foreground (3-D points) and background (4-D) */
  const scnerf_mlp* fg[2];
  const scnerf_mlp* bg[2];
} scnerf_pp_nets;
typedef struct scnerf_pp_step_io {
  const int64_t* select_inds_host; /* [N] pinned (inputs_on_host) */
  const float* target_host;        /* [N,3] pinned */
  float* loss_host;                /* [1] pinned */
  int64_t* select_inds_dev;        /* [N] staging / device-resident inputs */
  float* target_dev;               /* [N,3] */
  float* loss_dev;                 /* [1] (overwritten) */
  const float* min_depth_dev;      /* [N] per-ray near depth or NULL */
  int32_t* miss_dev;               /* [1] rays that miss the unit sphere (the reference raises, :61-65); overwritten */
  float* rgb_dev;                  /* [cascade_level, N, 3] rendered colours per level, or NULL */
} scnerf_pp_step_io;
size_t scnerf_pp_train_step_workspace_bytes(const scnerf_pp_step_cfg* cfg, const scnerf_mlp* fg, const scnerf_mlp* bg,
                                            int64_t N);
int scnerf_pp_train_step(const scnerf_pp_raygen_args* cam /* select_inds / N ignored */, const scnerf_camera_grads* g_cam,
                         float* g_distortion /* [2] or NULL */, const scnerf_pp_step_cfg* cfg,
                         const scnerf_pp_nets* nets, const scnerf_pp_nets* g_nets, const scnerf_pp_step_rand* rnd,
                         const scnerf_pp_step_io* io, int32_t inputs_on_host, int64_t N, void* workspace,
                         size_t workspace_bytes, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SCNERF_B200_NERFPP_H */
