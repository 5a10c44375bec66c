// C-ABI entry points of libscnerf_b200.so (see include/scnerf_b200.h for the contract and the
// reference file:line each replaces).  Host code here only sequences kernels on the caller's
// stream; it never allocates or synchronises.
#include <algorithm>
#include <cstring>
#include <cmath>

#include "common.cuh"
#include "raygen.cuh"
#include "sampling.cuh"
#include "composite.cuh"
#include "gemm_simt.cuh"
#include "field_simt.cuh"
#include "field_tc.cuh"
#include "nerfpp.cuh"
#include "adam.cuh"
#include "prd_loss.cuh"
#include "../../include/scnerf_b200_nerfpp.h"

using namespace scnerf;

extern "C" {

const char* scnerf_last_error(void) { return err_buf(); }
int scnerf_abi_version(void) { return 1; }
int scnerf_built_for_sm(void) { return 100; }
int scnerf_device_sm(void) {
  int dev = 0, maj = 0, min = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return -1;
  cudaDeviceGetAttribute(&maj, cudaDevAttrComputeCapabilityMajor, dev);
  cudaDeviceGetAttribute(&min, cudaDevAttrComputeCapabilityMinor, dev);
  return maj * 10 + min;
}
int64_t scnerf_launch_count(int32_t reset) {
  int64_t v = launch_counter().load();
  if (reset) launch_counter().store(0);
  return v;
}

int scnerf_kernel_timing(int32_t enable) {
  KernelTimes& k = kernel_times();
  std::lock_guard<std::mutex> g(k.mu);
  for (auto& r : k.recs) { k.pool.push_back(r.e0); k.pool.push_back(r.e1); }
  k.recs.clear();
  k.on.store(enable != 0);
  return 0;
}
int64_t scnerf_kernel_timing_report(char* buf, int64_t cap) {
  KernelTimes& k = kernel_times();
  std::lock_guard<std::mutex> g(k.mu);
  int64_t need = 0;
  for (auto& r : k.recs) {
    float ms = -1.f;
    if (cudaEventSynchronize(r.e1) == cudaSuccess) cudaEventElapsedTime(&ms, r.e0, r.e1);
    char line[256];
    int n = snprintf(line, sizeof(line), "%s\t%u\t%.6f\n", r.name, r.grid, ms);
    if (buf && need + n < cap) memcpy(buf + need, line, n);
    need += n;
  }
  if (buf && cap > 0) buf[need < cap ? need : cap - 1] = 0;
  return need + 1;
}

int scnerf_debug_mma_bench(int32_t mode, int32_t iters, long long* dev_out, int32_t nblocks, void* stream) {
  return tc_mma_bench(mode, iters, dev_out, nblocks, stream);
}

int scnerf_debug_slab_plan(int32_t which, int32_t index, int64_t* out9) {
  SCNERF_CHECK_ARG(out9 != nullptr, "slab_plan: null output");
  static const eng::Plan plans[5] = {fpipe::make_plan<3, 4>(), fpipe::make_plan<3, 4>(), fpipe::make_plan<3, 6>(),
                                     dpipe::make_plan<96>(), dpipe::make_plan<64>()};  // (0 = 1 since the serial forward was retired; 3 was the serial 4-D dgrad)
  SCNERF_CHECK_ARG(which >= 0 && which < 5, "slab_plan: unknown plan %d", which);
  const eng::Plan& P = plans[which];
  for (int i = 0; i < 9; ++i) out9[i] = 0;
  if (index < 0) {
    out9[0] = P.n_slabs; out9[1] = P.n_stages; out9[2] = (int64_t)fused::plan_image_bytes(P, 3);
    return 0;
  }
  SCNERF_CHECK_ARG(index < P.n_slabs, "slab_plan: slab %d of %d", index, P.n_slabs);
  const eng::SlabDef& d = P.slab[index];
  out9[0] = d.n; out9[1] = d.acc_col; out9[2] = d.a_off; out9[3] = d.a_lo_delta; out9[4] = d.a_kind;
  out9[5] = d.flags; out9[6] = d.stage; out9[7] = d.pad; out9[8] = d.img_off;
  return 0;
}

int scnerf_debug_timeline(long long* dev_buf, int32_t tiles) {
  tc_dbg_ptr() = dev_buf;
  tc_dbg_tiles() = tiles;
  return 0;
}

int scnerf_tc_selftest(const float* A, const float* B, float* D, int32_t N, int32_t K, int32_t variant,
                       void* workspace, size_t workspace_bytes, void* stream) {
  SCNERF_CHECK_ARG(A && B && D, "selftest: null pointer");
  return tc_selftest(A, B, D, N, K, variant, workspace, workspace_bytes, stream);
}

int scnerf_searchsorted_f32(const float* a, const float* v, int64_t* out, int64_t nrow_a,
                            int64_t nrow_v, int64_t ncol_a, int64_t ncol_v, int right, void* stream) {
  SCNERF_CHECK_ARG(a && v && out, "searchsorted: null pointer");
  SCNERF_CHECK_ARG(nrow_a == nrow_v || nrow_a == 1 || nrow_v == 1,
                   "searchsorted: row counts %lld vs %lld do not broadcast", (long long)nrow_a,
                   (long long)nrow_v);
  int64_t nrow = std::max(nrow_a, nrow_v);
  int64_t tot = nrow * ncol_v;
  if (tot == 0) return 0;
  SCNERF_LAUNCH(searchsorted_kernel, (unsigned)cdiv(tot, 256), 256, 0, stream, a, v, out, nrow, nrow_a,
                nrow_v, ncol_a, ncol_v, right);
  return 0;
}

int scnerf_camera_matrices(const scnerf_camera* cam, float* K_out, float* E_out, void* stream) {
  SCNERF_CHECK_ARG(cam && cam->intrinsics_initial && cam->extrinsics_initial, "camera: null params");
  SCNERF_LAUNCH(camera_matrices_kernel, (unsigned)cdiv(std::max(cam->n_cams, 1), 64), 64, 0, stream,
                *cam, K_out, E_out);
  return 0;
}

static int camera_pair_check(const scnerf_camera* cam, int64_t i0, int64_t i1) {
  SCNERF_CHECK_ARG(cam && cam->intrinsics_initial && cam->extrinsics_initial, "camera_pair: needs the learnable camera");
  SCNERF_CHECK_ARG(i0 >= 0 && i0 < cam->n_cams && i1 >= 0 && i1 < cam->n_cams, "camera_pair: camera index out of range");
  return 0;
}
int scnerf_camera_pair_fwd(const scnerf_camera* cam, int64_t i0, int64_t i1, float fx_sign, float* K4, float* E2,
                           void* stream) {
  int rc = camera_pair_check(cam, i0, i1);
  if (rc) return rc;
  SCNERF_CHECK_ARG(K4 && E2, "camera_pair_fwd: null outputs");
  SCNERF_LAUNCH(camera_pair_fwd_kernel, 1, 32, 0, stream, *cam, i0, i1, fx_sign, K4, E2);
  return 0;
}
int scnerf_camera_pair_bwd(const scnerf_camera* cam, int64_t i0, int64_t i1, float fx_sign, const float* d_K4,
                           const float* d_E2, const scnerf_camera_grads* g, void* stream) {
  int rc = camera_pair_check(cam, i0, i1);
  if (rc) return rc;
  SCNERF_CHECK_ARG(g, "camera_pair_bwd: null grads");
  SCNERF_LAUNCH(camera_pair_bwd_kernel, 1, 32, 0, stream, *cam, i0, i1, fx_sign, d_K4, d_E2, *g);
  return 0;
}

static int make_raygen_dev(const scnerf_raygen_args* a, RaygenDev& d) {
  SCNERF_CHECK_ARG(a, "raygen: null args");
  std::memset(&d, 0, sizeof(d));
  d.has_cam = a->cam != nullptr;
  if (a->cam) {
    d.cam = *a->cam;
    d.H = a->cam->H; d.W = a->cam->W;
    SCNERF_CHECK_ARG(d.cam.intrinsics_initial, "raygen: camera without intrinsics");
  } else {
    d.H = a->H; d.W = a->W; d.focal = a->focal;
    SCNERF_CHECK_ARG(a->extrinsic, "raygen: fixed-pinhole variant needs an extrinsic matrix");
  }
  SCNERF_CHECK_ARG(d.H > 0 && d.W > 0, "raygen: bad image size %dx%d", d.H, d.W);
  d.kps = a->kps; d.kps_f = a->kps ? nullptr : a->kps_f32; d.idx = a->idx; d.idx_scalar = a->idx_scalar;
  d.extrinsic = a->extrinsic; d.extrinsic_per_ray = a->extrinsic_per_ray; d.N = a->N;
  if (!a->extrinsic) {
    // exactly one pose source (NeRF/get_rays.py:107-110)
    SCNERF_CHECK_ARG(a->cam && a->cam->extrinsics_initial, "raygen: idx given but camera has no extrinsics");
    SCNERF_CHECK_ARG(a->idx || (a->idx_scalar >= 0 && a->idx_scalar < a->cam->n_cams),
                     "raygen: camera index %lld out of range", (long long)a->idx_scalar);
  } else {
    SCNERF_CHECK_ARG(!a->idx, "raygen: give idx or extrinsic, not both");
  }
  return 0;
}

int scnerf_raygen_fwd(const scnerf_raygen_args* a, float* rays_o, float* rays_d, void* stream) {
  RaygenDev d;
  int rc = make_raygen_dev(a, d);
  if (rc) return rc;
  if (d.N == 0) return 0;
  SCNERF_LAUNCH(raygen_fwd_kernel, (unsigned)cdiv(d.N, 128), 128, 0, stream, d, rays_o, rays_d);
  return 0;
}
int scnerf_raygen_bwd(const scnerf_raygen_args* a, const float* d_rays_o, const float* d_rays_d,
                      const scnerf_camera_grads* g, void* stream) {
  RaygenDev d;
  int rc = make_raygen_dev(a, d);
  if (rc) return rc;
  SCNERF_CHECK_ARG(d.has_cam && g, "raygen_bwd: needs the learnable camera and a grads struct");
  if (d.N == 0) return 0;
  SCNERF_LAUNCH(raygen_bwd_kernel, (unsigned)cdiv(d.N, 128), 128, 0, stream, d, d_rays_o, d_rays_d, *g);
  return 0;
}

int scnerf_ray_batch(const int64_t* shuffled_ray_idx, int64_t N, const float* images, const int64_t* i_train,
                     int64_t n_train, int32_t H, int32_t W, int64_t* kps, int64_t* image_idx, float* target,
                     void* stream) {
  SCNERF_CHECK_ARG(shuffled_ray_idx && images && i_train && kps && image_idx && target, "ray_batch: null pointer");
  SCNERF_CHECK_ARG(H > 0 && W > 0 && n_train > 0, "ray_batch: bad sizes");
  if (N == 0) return 0;
  SCNERF_LAUNCH(ray_batch_kernel, (unsigned)cdiv(N, 256), 256, 0, stream, shuffled_ray_idx, N, images, i_train, H, W,
                kps, image_idx, target);
  return 0;
}

static int make_rayprep_dev(const scnerf_rayprep_args* a, RayprepDev& d) {
  SCNERF_CHECK_ARG(a, "rayprep: null args");
  std::memset(&d, 0, sizeof(d));
  d.has_cam = a->cam != nullptr;
  if (a->cam) { d.cam = *a->cam; d.H = a->cam->H; d.W = a->cam->W; }
  else { d.H = a->H; d.W = a->W; }
  d.focal = a->focal; d.ndc = a->ndc; d.use_viewdirs = a->use_viewdirs;
  d.near_ = a->near_; d.far_ = a->far_; d.N = a->N;
  SCNERF_CHECK_ARG(!d.ndc || d.has_cam || d.focal > 0.f, "rayprep: ndc needs a focal length");
  return 0;
}
int scnerf_rayprep_fwd(const scnerf_rayprep_args* a, const float* rays_o, const float* rays_d,
                       float* rays, void* stream) {
  RayprepDev d;
  int rc = make_rayprep_dev(a, d);
  if (rc) return rc;
  if (d.N == 0) return 0;
  SCNERF_LAUNCH(rayprep_fwd_kernel, (unsigned)cdiv(d.N, 128), 128, 0, stream, d, rays_o, rays_d, rays);
  return 0;
}
int scnerf_rayprep_bwd(const scnerf_rayprep_args* a, const float* rays_o, const float* rays_d,
                       const float* d_rays, float* d_rays_o, float* d_rays_d,
                       float* d_intrinsics_noise, void* stream) {
  RayprepDev d;
  int rc = make_rayprep_dev(a, d);
  if (rc) return rc;
  if (d.N == 0) return 0;
  SCNERF_LAUNCH(rayprep_bwd_kernel, (unsigned)cdiv(d.N, 128), 128, 0, stream, d, rays_o, rays_d, d_rays,
                d_rays_o, d_rays_d, d_intrinsics_noise);
  return 0;
}

int scnerf_posenc_fwd(const float* x, int64_t P, int32_t L, float* out, void* stream) {
  SCNERF_CHECK_ARG(x && out && L >= 0 && L <= 16, "posenc: bad arguments");
  if (P == 0) return 0;
  SCNERF_LAUNCH(pe_points_kernel, (unsigned)cdiv(P * (L + 1), 256), 256, 0, stream, nullptr, 0, nullptr,
                x, P, 1, L, out, (int64_t)(3 + 6 * L));
  return 0;
}

size_t scnerf_field_workspace_bytes(const scnerf_mlp* m, int64_t P, int32_t training) {
  if (!m) return 0;
  Arena ar(nullptr, 0);
  FieldBufs B;
  field_bufs_alloc(ar, *m, P, training != 0, B);
  return ar.off + 256;
}
size_t scnerf_field_infer_workspace_bytes(const scnerf_mlp* m, int64_t P, int32_t precision) {
  if (!m) return 0;
  Arena ar(nullptr, 0);
  FieldBufs B;
  field_bufs_alloc(ar, *m, P, false, B, precision != SCNERF_PRECISION_FP32);
  return ar.off + 256;
}

static int field_fwd_dispatch(const scnerf_mlp& m, int precision, const float* rays, int ray_cols,
                              const float* z, const float* pts, const float* viewdirs, int64_t N,
                              int S, const FieldBufs& B, float* raw, void* stream,
                              const TcFwdImages* imgs = nullptr, const CompositeArgs* comp = nullptr,
                              bool* comp_done = nullptr, bool raw_needed = true) {
  if (comp_done) *comp_done = false;
  if (precision == SCNERF_PRECISION_FP32)
    return field_simt_fwd(m, rays, ray_cols, z, pts, viewdirs, N, S, B, raw, stream);
  return field_tc_fwd(m, precision, rays, ray_cols, z, pts, viewdirs, N, S, B, raw, stream, imgs, comp, comp_done, raw_needed);
}

int scnerf_field_fwd(const scnerf_mlp* m, const float* pts, const float* viewdirs, int64_t N, int64_t S,
                     float* raw, int32_t precision, void* workspace, size_t workspace_bytes,
                     void* stream) {
  SCNERF_CHECK_ARG(m && pts && raw, "field_fwd: null pointer");
  int rc = field_check(*m);
  if (rc) return rc;
  SCNERF_CHECK_ARG(!m->use_viewdirs || viewdirs, "field_fwd: use_viewdirs needs viewdirs");
  if (N * S == 0) return 0;
  Arena ar(workspace, workspace_bytes);
  FieldBufs B;
  // tensor-core precisions keep every activation in TMEM: only the packed weight image and constants live here
  field_bufs_alloc(ar, *m, N * S, false, B, precision != SCNERF_PRECISION_FP32);
  if (!workspace || !ar.ok())
    return fail(SCNERF_ERR_WORKSPACE, "field_fwd: workspace %zu < %zu bytes", workspace_bytes, ar.off);
  return field_fwd_dispatch(*m, precision, nullptr, 0, nullptr, pts, viewdirs, N, (int)S, B, raw, stream);
}

int scnerf_raw2outputs_fwd(const float* raw, int32_t raw_cols, const float* z, const float* rays_d,
                           int32_t rays_d_stride, const float* noise, int32_t white_bkgd, int64_t N,
                           int64_t S, float* rgb_map, float* disp_map, float* acc_map, float* weights,
                           float* depth_map, void* stream) {
  SCNERF_CHECK_ARG(raw && z && rays_d && rgb_map && disp_map && acc_map, "raw2outputs: null pointer");
  SCNERF_CHECK_ARG(raw_cols >= 4, "raw2outputs: raw needs >= 4 channels");
  if (N == 0) return 0;
  CompositeArgs a{};
  a.raw = raw; a.raw_cols = raw_cols; a.z = z; a.rays_d = rays_d; a.d_stride = rays_d_stride;
  a.noise = noise; a.noise_std = noise ? 1.f : 0.f; a.white_bkgd = white_bkgd; a.N = N; a.S = (int)S;
  a.rgb_map = rgb_map; a.disp_map = disp_map; a.acc_map = acc_map; a.weights = weights;
  a.depth_map = depth_map;
  SCNERF_LAUNCH(composite_fwd_kernel, (unsigned)cdiv(N, 4), 128, 0, stream, a);
  return 0;
}

int scnerf_sample_pdf_fwd(const float* bins, const float* weights, const float* u, int64_t N, int64_t M,
                          int64_t Nf, float* samples, int64_t* inds, void* stream) {
  SCNERF_CHECK_ARG(bins && weights && samples, "sample_pdf: null pointer");
  SCNERF_CHECK_ARG(M >= 2 && M <= 4096, "sample_pdf: %lld bins unsupported", (long long)M);
  if (N == 0 || Nf == 0) return 0;
  SamplePdfArgs a{};
  a.bins = bins; a.M = (int)M; a.w = weights; a.w_stride = M - 1; a.w_off = 0;
  a.u = u; a.det = (u == nullptr); a.Nf = (int)Nf; a.samples = samples; a.inds = inds; a.N = N;
  a.sort_n = 0;
  size_t smem = sizeof(float) * (2 * M);
  SCNERF_LAUNCH(sample_pdf_kernel, (unsigned)N, 128, smem, stream, a);
  return 0;
}

int scnerf_sort_merge(const float* a, const float* b, int64_t N, int64_t Na, int64_t Nb, float* out,
                      void* stream) {
  SCNERF_CHECK_ARG(a && b && out, "sort_merge: null pointer");
  SCNERF_CHECK_ARG(Na + Nb <= 8192, "sort_merge: %lld keys per row unsupported", (long long)(Na + Nb));
  if (N == 0) return 0;
  int sn = next_pow2((int)(Na + Nb));
  SCNERF_LAUNCH(sort_merge_kernel, (unsigned)N, 128, sizeof(float) * sn, stream, a, b, (int)Na, (int)Nb,
                sn, out);
  return 0;
}

// ---------------------------------------------------------------------------------------------------
// render_rays
// ---------------------------------------------------------------------------------------------------
struct RenderWS {
  float *z_c, *raw_c, *w_c, *acc_c, *depth_c;
  float *z_f, *raw_f, *w_f, *acc_f, *depth_f;
  float* g_raw;
  FieldBufs fb_c, fb_f;
  FieldGradBufs gb;
  bool tc_train;            // tensor-core training path: bf16 tile images instead of fp32 activations
  TcFwdImages im_c, im_f;
  TcBwdBufs tb;
};

static void render_ws_layout(Arena& ar, const scnerf_render_cfg& cfg, const scnerf_mlp& m, int64_t N,
                             RenderWS& w) {
  const int Nc = cfg.N_samples, Nf = cfg.N_importance, St = Nc + Nf, rc = field_raw_cols(m);
  w.z_c = ar.get<float>(N * Nc);
  w.raw_c = ar.get<float>(N * Nc * rc);
  w.w_c = ar.get<float>(N * Nc);
  w.acc_c = ar.get<float>(N);
  w.depth_c = ar.get<float>(N);
  w.z_f = w.raw_f = w.w_f = w.acc_f = w.depth_f = nullptr;
  if (Nf > 0) {
    w.z_f = ar.get<float>(N * St);
    w.raw_f = ar.get<float>(N * St * rc);
    w.w_f = ar.get<float>(N * St);
    w.acc_f = ar.get<float>(N);
    w.depth_f = ar.get<float>(N);
  }
  w.tc_train = cfg.training && cfg.precision != SCNERF_PRECISION_FP32;
  if (w.tc_train) {
    const int ns = cfg.precision == SCNERF_PRECISION_BF16X3 ? 3 : 1;
    w.g_raw = ar.get<float>(N * std::max(Nc, St) * 4);
    field_bufs_alloc(ar, m, N * Nc, true, w.fb_c, true);
    tc_fwd_images_alloc(ar, N * Nc, ns, w.im_c);
    if (Nf > 0) {
      field_bufs_alloc(ar, m, N * St, true, w.fb_f, true);
      tc_fwd_images_alloc(ar, N * St, ns, w.im_f);
    }
    tc_bwd_bufs_alloc(ar, N * std::max(Nc, St), ns, w.tb);
  } else if (cfg.training) {
    w.g_raw = ar.get<float>(N * std::max(Nc, St) * 4);
    field_bufs_alloc(ar, m, N * Nc, true, w.fb_c);
    if (Nf > 0) field_bufs_alloc(ar, m, N * St, true, w.fb_f);
    field_grad_bufs_alloc(ar, m, N * std::max(Nc, St), w.gb);
  } else {
    w.g_raw = nullptr;
    field_bufs_alloc(ar, m, N * std::max(Nc, St), false, w.fb_c, cfg.precision != SCNERF_PRECISION_FP32);
    w.fb_f = w.fb_c;
  }
}

static int render_cfg_check(const scnerf_render_cfg* cfg, const scnerf_mlp* coarse, const scnerf_mlp* fine) {
  SCNERF_CHECK_ARG(cfg && coarse, "render: null cfg/mlp");
  SCNERF_CHECK_ARG(cfg->N_samples >= 2 && cfg->N_samples <= 2048, "render: N_samples %d unsupported", cfg->N_samples);
  SCNERF_CHECK_ARG(cfg->N_importance >= 0 && cfg->N_samples + cfg->N_importance <= 4096, "render: N_importance %d unsupported", cfg->N_importance);
  SCNERF_CHECK_ARG(cfg->ray_cols == 8 || cfg->ray_cols == 11, "render: ray batch must have 8 or 11 columns");
  SCNERF_CHECK_ARG(cfg->precision >= 0 && cfg->precision <= 2, "render: unknown precision %d", cfg->precision);
  int rc = field_check(*coarse);
  if (rc) return rc;
  SCNERF_CHECK_ARG(!coarse->use_viewdirs || cfg->ray_cols == 11, "render: use_viewdirs needs 11-column rays");
  if (fine) {
    rc = field_check(*fine);
    if (rc) return rc;
    SCNERF_CHECK_ARG(fine->D == coarse->D && fine->W == coarse->W && fine->use_viewdirs == coarse->use_viewdirs &&
                     fine->input_ch == coarse->input_ch && fine->skip == coarse->skip &&
                     fine->output_ch == coarse->output_ch,
                     "render: coarse and fine networks must share an architecture");
  }
  return 0;
}

size_t scnerf_render_workspace_bytes(const scnerf_render_cfg* cfg, const scnerf_mlp* coarse, int64_t N) {
  if (!cfg || !coarse) return 0;
  Arena ar(nullptr, 0);
  RenderWS w;
  render_ws_layout(ar, *cfg, *coarse, N, w);
  return ar.off + 256;
}

static CompositeArgs composite_args(const scnerf_render_cfg& cfg, const float* raw, int rc, const float* z,
                                    const float* rays, const float* noise, uint32_t stream_id, int64_t N, int S,
                                    float* rgb, float* disp, float* acc, float* weights, float* depth) {
  CompositeArgs a{};
  a.raw = raw; a.raw_cols = rc; a.z = z; a.rays_d = rays + 3; a.d_stride = cfg.ray_cols;
  a.noise = noise; a.noise_std = cfg.raw_noise_std; a.seed = cfg.seed; a.rng_stream = stream_id;
  a.white_bkgd = cfg.white_bkgd; a.N = N; a.S = S;
  a.rgb_map = rgb; a.disp_map = disp; a.acc_map = acc; a.weights = weights; a.depth_map = depth;
  return a;
}
static int composite_launch(const CompositeArgs& a, void* stream) {
  SCNERF_LAUNCH(composite_fwd_kernel, (unsigned)cdiv(a.N, 4), 128, 0, stream, a);
  return 0;
}

int scnerf_render_rays_fwd(const scnerf_render_cfg* cfg, const float* rays, int64_t N,
                           const scnerf_mlp* coarse, const scnerf_mlp* fine,
                           const scnerf_render_rand* rnd, const scnerf_render_out* out, void* workspace,
                           size_t workspace_bytes, void* stream) {
  int rc = render_cfg_check(cfg, coarse, fine);
  if (rc) return rc;
  SCNERF_CHECK_ARG(rays && out && out->rgb_map && out->disp_map && out->acc_map, "render: null outputs");
  if (N == 0) return 0;
  SCNERF_CHECK_ARG(N * (int64_t)(cfg->N_samples + cfg->N_importance) <= (int64_t)65535 * 128,
                   "render: %lld rays per call too many (chunk the batch)", (long long)N);
  const scnerf_mlp& mc = *coarse;
  const scnerf_mlp& mf = fine ? *fine : *coarse;
  Arena ar(workspace, workspace_bytes);
  RenderWS w;
  render_ws_layout(ar, *cfg, mc, N, w);
  if (!workspace || !ar.ok())
    return fail(SCNERF_ERR_WORKSPACE, "render: workspace %zu < %zu bytes", workspace_bytes, ar.off);
  const int Nc = cfg->N_samples, Nf = cfg->N_importance, St = Nc + Nf, rcn = field_raw_cols(mc);
  scnerf_render_rand none{};
  if (!rnd) rnd = &none;

  SCNERF_LAUNCH(stratified_kernel, (unsigned)cdiv(N * Nc, 256), 256, 0, stream, rays, cfg->ray_cols, N, Nc,
                cfg->lindisp, cfg->perturb > 0, rnd->t_rand, cfg->seed, w.z_c);
  // Inference (no backward to feed): the alpha-composite runs inside the field kernel's epilogue and `raw` leaves the
  // kernel only when the caller asked for it (retraw, last level).  Training keeps raw for composite_bwd.
  const bool two = Nf > 0;
  const bool fuse = !cfg->training && rcn == 4;
  bool fused_c = false;
  CompositeArgs ca = composite_args(*cfg, w.raw_c, rcn, w.z_c, rays, rnd->noise0, RNG_NOISE0, N, Nc,
                                    two ? out->rgb0 : out->rgb_map, two ? out->disp0 : out->disp_map,
                                    two ? out->acc0 : out->acc_map, w.w_c, w.depth_c);
  const bool keep_raw_c = !two && cfg->retraw && out->raw;
  rc = field_fwd_dispatch(mc, cfg->precision, rays, cfg->ray_cols, w.z_c, nullptr, nullptr, N, Nc, w.fb_c,
                          w.raw_c, stream, w.tc_train ? &w.im_c : nullptr, fuse ? &ca : nullptr, &fused_c, keep_raw_c);
  if (rc) return rc;
  if (!fused_c) {
    rc = composite_launch(ca, stream);
    if (rc) return rc;
  }
  if (two) SCNERF_CHECK_ARG(out->rgb0 && out->disp0 && out->acc0, "render: N_importance>0 needs rgb0/disp0/acc0");
  const float* z_last = w.z_c; const float* raw_last = w.raw_c; const float* w_last = w.w_c;
  const float* depth_last = w.depth_c;
  int S_last = Nc;
  if (two) {
    SamplePdfArgs a{};
    a.bins = nullptr; a.z_c = w.z_c; a.Nc = Nc; a.M = Nc - 1;
    a.w = w.w_c; a.w_stride = Nc; a.w_off = 1;   // weights[..., 1:-1]  (render.py:271)
    a.u = rnd->u; a.det = (cfg->perturb == 0); a.seed = cfg->seed; a.Nf = Nf;
    a.samples = nullptr; a.inds = nullptr; a.merged = w.z_f; a.z_std = out->z_std; a.N = N;
    a.sort_n = next_pow2(St);
    size_t smem = sizeof(float) * (2 * (Nc - 1) + a.sort_n);
    SCNERF_LAUNCH(sample_pdf_kernel, (unsigned)N, 128, smem, stream, a);
    bool fused_f = false;
    CompositeArgs cf = composite_args(*cfg, w.raw_f, rcn, w.z_f, rays, rnd->noise1, RNG_NOISE1, N, St, out->rgb_map,
                                      out->disp_map, out->acc_map, w.w_f, w.depth_f);
    rc = field_fwd_dispatch(mf, cfg->precision, rays, cfg->ray_cols, w.z_f, nullptr, nullptr, N, St, w.fb_f,
                            w.raw_f, stream, w.tc_train ? &w.im_f : nullptr, fuse ? &cf : nullptr, &fused_f,
                            cfg->retraw && out->raw);
    if (rc) return rc;
    if (!fused_f) {
      rc = composite_launch(cf, stream);
      if (rc) return rc;
    }
    z_last = w.z_f; raw_last = w.raw_f; w_last = w.w_f; depth_last = w.depth_f; S_last = St;
  }
  // forward copies of acc for the backward (disp gradient) live in the user outputs; keep our own
  cudaStream_t st = (cudaStream_t)stream;
  if (cfg->training) {
    SCNERF_CUDA(cudaMemcpyAsync(w.acc_c, two ? out->acc0 : out->acc_map, sizeof(float) * N,
                                cudaMemcpyDeviceToDevice, st));
    if (two) SCNERF_CUDA(cudaMemcpyAsync(w.acc_f, out->acc_map, sizeof(float) * N, cudaMemcpyDeviceToDevice, st));
  }
  if (cfg->retraw && out->raw)
    SCNERF_CUDA(cudaMemcpyAsync(out->raw, raw_last, sizeof(float) * N * S_last * rcn, cudaMemcpyDeviceToDevice, st));
  if (out->z_vals)
    SCNERF_CUDA(cudaMemcpyAsync(out->z_vals, z_last, sizeof(float) * N * S_last, cudaMemcpyDeviceToDevice, st));
  if (out->weights)
    SCNERF_CUDA(cudaMemcpyAsync(out->weights, w_last, sizeof(float) * N * S_last, cudaMemcpyDeviceToDevice, st));
  if (out->depth_map)
    SCNERF_CUDA(cudaMemcpyAsync(out->depth_map, depth_last, sizeof(float) * N, cudaMemcpyDeviceToDevice, st));
  return 0;
}

static int composite_bwd_launch(const scnerf_render_cfg& cfg, const float* raw, int rc, const float* z,
                                const float* rays, const float* noise, uint32_t stream_id, int64_t N,
                                int S, const float* g_rgb, const float* g_disp, const float* g_acc,
                                const float* acc_saved, const float* depth_saved, float* g_raw,
                                float* d_rays, void* stream) {
  CompositeBwdArgs b{};
  b.f.raw = raw; b.f.raw_cols = rc; b.f.z = z; b.f.rays_d = rays + 3; b.f.d_stride = cfg.ray_cols;
  b.f.noise = noise; b.f.noise_std = cfg.raw_noise_std; b.f.seed = cfg.seed; b.f.rng_stream = stream_id;
  b.f.white_bkgd = cfg.white_bkgd; b.f.N = N; b.f.S = S;
  b.g_rgb = g_rgb; b.g_disp = g_disp; b.g_acc = g_acc; b.acc_saved = acc_saved; b.depth_saved = depth_saved;
  b.g_raw = g_raw; b.g_rays = d_rays; b.g_rays_cols = cfg.ray_cols;
  size_t smem = sizeof(float) * 2 * S * 4;
  if (smem > 48 * 1024) {    // S > 1536: opt in to the large dynamic shared-memory carve-out (S <= 4096 -> 128 KB)
    static bool opted = false;
    if (!opted) {
      SCNERF_CUDA(cudaFuncSetAttribute(composite_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * 4096 * 4 * 4));
      opted = true;
    }
  }
  SCNERF_LAUNCH(composite_bwd_kernel, (unsigned)cdiv(N, 4), 128, smem, stream, b);
  return 0;
}

int scnerf_render_rays_bwd(const scnerf_render_cfg* cfg, const float* rays, int64_t N,
                           const scnerf_mlp* coarse, const scnerf_mlp* fine,
                           const scnerf_render_rand* rnd, const scnerf_render_grads_in* gin,
                           const scnerf_mlp* g_coarse, const scnerf_mlp* g_fine, float* d_rays,
                           void* workspace, size_t workspace_bytes, void* stream) {
  int rc = render_cfg_check(cfg, coarse, fine);
  if (rc) return rc;
  SCNERF_CHECK_ARG(cfg->training, "render_bwd: the forward must have run with cfg.training=1");
  SCNERF_CHECK_ARG(rays && gin && g_coarse && d_rays, "render_bwd: null pointer");
  SCNERF_CHECK_ARG(!fine || g_fine, "render_bwd: fine network given without a gradient struct");
  if (N == 0) return 0;
  const scnerf_mlp& mc = *coarse;
  const scnerf_mlp& mf = fine ? *fine : *coarse;
  const scnerf_mlp& gc = *g_coarse;
  const scnerf_mlp& gf = fine ? *g_fine : *g_coarse;
  Arena ar(workspace, workspace_bytes);
  RenderWS w;
  render_ws_layout(ar, *cfg, mc, N, w);
  if (!workspace || !ar.ok())
    return fail(SCNERF_ERR_WORKSPACE, "render_bwd: workspace %zu < %zu bytes", workspace_bytes, ar.off);
  const int Nc = cfg->N_samples, Nf = cfg->N_importance, St = Nc + Nf, rcn = field_raw_cols(mc);
  const bool two = Nf > 0;
  scnerf_render_rand none{};
  if (!rnd) rnd = &none;
  SCNERF_CUDA(cudaMemsetAsync(d_rays, 0, sizeof(float) * N * cfg->ray_cols, (cudaStream_t)stream));
  if (two && (gin->rgb_map || gin->disp_map || gin->acc_map)) {
    rc = composite_bwd_launch(*cfg, w.raw_f, rcn, w.z_f, rays, rnd->noise1, RNG_NOISE1, N, St, gin->rgb_map,
                              gin->disp_map, gin->acc_map, w.acc_f, w.depth_f, w.g_raw, d_rays, stream);
    if (rc) return rc;
    rc = w.tc_train ? field_tc_bwd(mf, gf, cfg->precision, rays, cfg->ray_cols, w.z_f, N, St, w.fb_f, w.im_f, w.tb,
                                   w.g_raw, d_rays, stream)
                    : field_simt_bwd(mf, gf, rays, cfg->ray_cols, w.z_f, N, St, w.fb_f, w.gb, w.g_raw, d_rays, stream);
    if (rc) return rc;
  }
  const float* g_rgb = two ? gin->rgb0 : gin->rgb_map;
  const float* g_disp = two ? gin->disp0 : gin->disp_map;
  const float* g_acc = two ? gin->acc0 : gin->acc_map;
  if (g_rgb || g_disp || g_acc) {
    rc = composite_bwd_launch(*cfg, w.raw_c, rcn, w.z_c, rays, rnd->noise0, RNG_NOISE0, N, Nc, g_rgb, g_disp,
                              g_acc, w.acc_c, w.depth_c, w.g_raw, d_rays, stream);
    if (rc) return rc;
    rc = w.tc_train ? field_tc_bwd(mc, gc, cfg->precision, rays, cfg->ray_cols, w.z_c, N, Nc, w.fb_c, w.im_c, w.tb,
                                   w.g_raw, d_rays, stream)
                    : field_simt_bwd(mc, gc, rays, cfg->ray_cols, w.z_c, N, Nc, w.fb_c, w.gb, w.g_raw, d_rays, stream);
    if (rc) return rc;
  }
  return 0;
}

// ---------------------------------------------------------------------------------------------------
// whole training step (NeRF/run_nerf.py:385-506,600)
// ---------------------------------------------------------------------------------------------------
struct StepWS {
  float *ro, *rd, *rays, *d_rays, *d_ro, *d_rd;
  float *rgb, *disp, *acc, *rgb0, *disp0, *acc0, *z_std, *g_rgb, *g_rgb0;
  void* render_ws;
  size_t render_bytes;
};
static void step_ws_layout(Arena& ar, const scnerf_render_cfg& cfg, const scnerf_mlp& m, int64_t N, StepWS& s) {
  s.ro = ar.get<float>(N * 3); s.rd = ar.get<float>(N * 3);
  s.rays = ar.get<float>(N * 11); s.d_rays = ar.get<float>(N * 11);
  s.d_ro = ar.get<float>(N * 3); s.d_rd = ar.get<float>(N * 3);
  s.rgb = ar.get<float>(N * 3); s.disp = ar.get<float>(N); s.acc = ar.get<float>(N);
  s.rgb0 = ar.get<float>(N * 3); s.disp0 = ar.get<float>(N); s.acc0 = ar.get<float>(N);
  s.z_std = ar.get<float>(N); s.g_rgb = ar.get<float>(N * 3); s.g_rgb0 = ar.get<float>(N * 3);
  Arena sub(nullptr, 0);
  RenderWS w;
  render_ws_layout(sub, cfg, m, N, w);
  s.render_bytes = sub.off + 256;
  s.render_ws = ar.get<char>(s.render_bytes);
}
size_t scnerf_train_step_workspace_bytes(const scnerf_render_cfg* cfg, const scnerf_mlp* coarse, int64_t N) {
  if (!cfg || !coarse) return 0;
  scnerf_render_cfg c = *cfg;
  c.training = 1;
  Arena ar(nullptr, 0);
  StepWS s;
  step_ws_layout(ar, c, *coarse, N, s);
  return ar.off + 256;
}

int scnerf_train_step(const scnerf_camera* cam, const scnerf_camera_grads* g_cam,
                      const scnerf_render_cfg* cfg_in, int32_t ndc, float near_, float far_,
                      const scnerf_mlp* coarse, const scnerf_mlp* fine, const scnerf_mlp* g_coarse,
                      const scnerf_mlp* g_fine, const scnerf_step_io* io, int32_t inputs_on_host,
                      int64_t N, void* workspace, size_t workspace_bytes, void* stream) {
  SCNERF_CHECK_ARG(cam && g_cam && cfg_in && coarse && g_coarse && io, "train_step: null pointer");
  SCNERF_CHECK_ARG(io->kps_dev && io->idx_dev && io->target_dev && io->loss_dev, "train_step: device staging buffers missing");
  scnerf_render_cfg cfg = *cfg_in;
  cfg.training = 1;
  cfg.ray_cols = coarse->use_viewdirs ? 11 : 8;
  cfg.retraw = 0;
  Arena ar(workspace, workspace_bytes);
  StepWS s;
  step_ws_layout(ar, cfg, *coarse, N, s);
  if (!workspace || !ar.ok())
    return fail(SCNERF_ERR_WORKSPACE, "train_step: workspace %zu < %zu bytes", workspace_bytes, ar.off);
  cudaStream_t st = (cudaStream_t)stream;
  if (inputs_on_host) {
    SCNERF_CHECK_ARG(io->kps_host && io->idx_host && io->target_host && io->loss_host, "train_step: host buffers missing");
    SCNERF_CUDA(cudaMemcpyAsync(io->kps_dev, io->kps_host, sizeof(int64_t) * 2 * N, cudaMemcpyHostToDevice, st));
    SCNERF_CUDA(cudaMemcpyAsync(io->idx_dev, io->idx_host, sizeof(int64_t) * N, cudaMemcpyHostToDevice, st));
    SCNERF_CUDA(cudaMemcpyAsync(io->target_dev, io->target_host, sizeof(float) * 3 * N, cudaMemcpyHostToDevice, st));
  }
  SCNERF_CUDA(cudaMemsetAsync(io->loss_dev, 0, sizeof(float), st));
  scnerf_raygen_args ra{};
  ra.cam = cam; ra.kps = io->kps_dev; ra.idx = io->idx_dev; ra.idx_scalar = -1; ra.N = N;
  int rc = scnerf_raygen_fwd(&ra, s.ro, s.rd, stream);
  if (rc) return rc;
  scnerf_rayprep_args pa{};
  pa.cam = cam; pa.ndc = ndc; pa.use_viewdirs = coarse->use_viewdirs; pa.near_ = near_; pa.far_ = far_; pa.N = N;
  rc = scnerf_rayprep_fwd(&pa, s.ro, s.rd, s.rays, stream);
  if (rc) return rc;
  scnerf_render_out out{};
  out.rgb_map = s.rgb; out.disp_map = s.disp; out.acc_map = s.acc;
  out.rgb0 = s.rgb0; out.disp0 = s.disp0; out.acc0 = s.acc0; out.z_std = s.z_std;
  rc = scnerf_render_rays_fwd(&cfg, s.rays, N, coarse, fine, nullptr, &out, s.render_ws, s.render_bytes, stream);
  if (rc) return rc;
  const bool two = cfg.N_importance > 0;
  SCNERF_LAUNCH(mse_loss_kernel, (unsigned)cdiv(N * 3, 256), 256, 0, stream, s.rgb, two ? s.rgb0 : nullptr,
                io->target_dev, N * 3, s.g_rgb, s.g_rgb0, io->loss_dev);
  scnerf_render_grads_in gin{};
  gin.rgb_map = s.g_rgb;
  gin.rgb0 = two ? s.g_rgb0 : nullptr;
  rc = scnerf_render_rays_bwd(&cfg, s.rays, N, coarse, fine, nullptr, &gin, g_coarse, g_fine, s.d_rays,
                              s.render_ws, s.render_bytes, stream);
  if (rc) return rc;
  rc = scnerf_rayprep_bwd(&pa, s.ro, s.rd, s.d_rays, s.d_ro, s.d_rd, g_cam->intrinsics_noise, stream);
  if (rc) return rc;
  rc = scnerf_raygen_bwd(&ra, s.d_ro, s.d_rd, g_cam, stream);
  if (rc) return rc;
  if (inputs_on_host)
    SCNERF_CUDA(cudaMemcpyAsync(io->loss_host, io->loss_dev, sizeof(float), cudaMemcpyDeviceToHost, st));
  return 0;
}

}  // extern "C"

extern "C" int scnerf_adam_step(const scnerf_adam_tensor* tensors_host, int32_t n_tensors, float lr, float beta1,
                                float beta2, float eps, float weight_decay, void* stream) {
  SCNERF_CHECK_ARG(tensors_host || n_tensors == 0, "adam_step: null tensor table");
  SCNERF_CHECK_ARG(beta1 >= 0.f && beta1 < 1.f && beta2 >= 0.f && beta2 < 1.f && eps >= 0.f && lr >= 0.f,
                   "adam_step: bad hyper-parameters");
  int i = 0;
  while (i < n_tensors) {
    AdamTable tab{};
    tab.beta1 = beta1; tab.beta2 = beta2; tab.eps = eps;
    int64_t blocks = 0;
    for (; i < n_tensors && tab.n < ADAM_MAX_TENSORS; ++i) {
      const scnerf_adam_tensor& s = tensors_host[i];
      if (s.numel == 0) continue;
      SCNERF_CHECK_ARG(s.param && s.grad && s.exp_avg && s.exp_avg_sq && s.step >= 1, "adam_step: tensor %d incomplete", i);
      AdamTensor& t = tab.t[tab.n++];
      t.p = s.param; t.g = s.grad; t.m = s.exp_avg; t.v = s.exp_avg_sq; t.vmax = s.max_exp_avg_sq; t.n = s.numel;
      t.first_block = (int32_t)blocks;
      const double bc1 = 1.0 - std::pow((double)beta1, (double)s.step), bc2 = 1.0 - std::pow((double)beta2, (double)s.step);
      t.step_size = (float)((double)lr / bc1);
      t.inv_sqrt_bc2 = (float)(1.0 / std::sqrt(bc2));
      t.weight_decay = s.decay ? weight_decay : 0.f;
      blocks += cdiv(s.numel, ADAM_BLOCK * ADAM_ILP);
    }
    if (tab.n == 0) break;
    SCNERF_LAUNCH(adam_multi_kernel, (unsigned)blocks, ADAM_BLOCK, 0, stream, tab);
  }
  return 0;
}

extern "C" {
static int make_prd_args(const float* o0, const float* d0, const float* o1, const float* d1, const float* kps0,
                         const float* kps1, const float* K4, const float* E2, float eps, float threshold, int train,
                         int64_t N, prd::Args& a) {
  SCNERF_CHECK_ARG(o0 && d0 && o1 && d1 && kps0 && kps1 && K4 && E2, "prd_loss: null pointer");
  a.o0 = o0; a.d0 = d0; a.o1 = o1; a.d1 = d1; a.kps0 = kps0; a.kps1 = kps1; a.K = K4; a.E = E2;
  a.eps = eps; a.threshold = threshold; a.train = train; a.N = N;
  return 0;
}
int scnerf_prd_loss_fwd(const float* rays0_o, const float* rays0_d, const float* rays1_o, const float* rays1_d,
                        const float* kps0, const float* kps1, const float* K4, const float* E2, float eps,
                        float threshold, int32_t train, int64_t N, float* acc5, float* loss, float* n_match,
                        void* stream) {
  prd::Args a;
  int rc = make_prd_args(rays0_o, rays0_d, rays1_o, rays1_d, kps0, kps1, K4, E2, eps, threshold, train, N, a);
  if (rc) return rc;
  SCNERF_CHECK_ARG(acc5 && loss, "prd_loss_fwd: null outputs");
  SCNERF_CUDA(cudaMemsetAsync(acc5, 0, 5 * sizeof(float), (cudaStream_t)stream));
  if (N > 0) SCNERF_LAUNCH(prd::fwd_kernel, (unsigned)cdiv(N, 128), 128, 0, stream, a, acc5);
  SCNERF_LAUNCH(prd::finalize_kernel, 1, 32, 0, stream, acc5, loss, n_match);
  return 0;
}
int scnerf_prd_loss_bwd(const float* rays0_o, const float* rays0_d, const float* rays1_o, const float* rays1_d,
                        const float* kps0, const float* kps1, const float* K4, const float* E2, float eps,
                        float threshold, int64_t N, const float* acc5, const float* d_loss, float* d_rays0_o,
                        float* d_rays0_d, float* d_rays1_o, float* d_rays1_d, float* d_K4, float* d_E2, void* stream) {
  prd::Args a;
  int rc = make_prd_args(rays0_o, rays0_d, rays1_o, rays1_d, kps0, kps1, K4, E2, eps, threshold, 1, N, a);
  if (rc) return rc;
  SCNERF_CHECK_ARG(acc5 && d_loss && d_rays0_o && d_rays0_d && d_rays1_o && d_rays1_d, "prd_loss_bwd: null pointer");
  if (N == 0) return 0;
  SCNERF_LAUNCH(prd::bwd_kernel, (unsigned)cdiv(N, 128), 128, 0, stream, a, acc5, d_loss, d_rays0_o, d_rays0_d,
                d_rays1_o, d_rays1_d, d_K4, d_E2);
  return 0;
}
}  // extern "C"

#include "api_pp.inc"
