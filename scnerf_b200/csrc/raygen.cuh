// Differentiable ray generation from the learnable camera, and render()'s ray packing (NDC).
// One thread per ray; everything a ray needs (K^-1, 6-D rotation, bilinear residual taps) is
// recomputed in registers — no [H*W,3] upsampled field, no [n,4,4] pose tensor is materialised
// (the reference builds both every call: model/camera_model.py:24-46, :179-190).
//
// Algorithmic HBM bytes per ray: 24 B in (kps 16 + idx 8) + 24 B out (o, d);
// camera tables (17x9 + 4 + 2x37x50x3 floats = 45 KB) stay L1/L2 resident.
#pragma once
#include "common.cuh"

namespace scnerf {

struct Intr { float fx, fy, cx, cy; };

__device__ __forceinline__ Intr load_intrinsics(const scnerf_camera& c) {
  // model/camera_model.py:166-177: init + noise*scale*init (multiplicative) | init + noise*scale
  float p[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float init = c.intrinsics_initial[i];
    float n = c.intrinsics_noise ? c.intrinsics_noise[i] : 0.f;
    float ns = __fmul_rn(n, c.intrinsics_noise_scale);
    p[i] = c.multiplicative_noise ? __fadd_rn(init, __fmul_rn(ns, init)) : __fadd_rn(init, ns);
  }
  return {p[0], p[1], p[2], p[3]};
}

__device__ __forceinline__ float dot3(const float* a, const float* b) {
  return a[0] * b[0] + a[1] * b[1] + a[2] * b[2];
}

// unit(v) = v / (max(|v|,1e-8) + 1e-10)   model/camera_utils.py:88-95
__device__ __forceinline__ void unit_fwd(const float* v, float* out, float& mag) {
  mag = sqrtf(dot3(v, v));
  float inv = 1.f / (fmaxf(mag, 1e-8f) + 1e-10f);
#pragma unroll
  for (int i = 0; i < 3; ++i) out[i] = v[i] * inv;
}
__device__ __forceinline__ void unit_bwd(const float* v, float mag, const float* g_out, float* g_v) {
  float den = fmaxf(mag, 1e-8f) + 1e-10f;
  float inv = 1.f / den;
  float gv = dot3(g_out, v);
  float k = (mag > 1e-8f) ? gv * inv * inv / mag : 0.f;
#pragma unroll
  for (int i = 0; i < 3; ++i) g_v[i] = g_out[i] * inv - k * v[i];
}

struct Pose {          // camera-to-world
  float x[3], y[3], z[3], t[3];  // rotation columns, translation
  // Gram-Schmidt intermediates (kept for the backward)
  float a[3], b[3], yp[3], mag_a, mag_y, coef, xx, xb;
};

// model/camera_model.py:179-190 + model/camera_utils.py:78-133 for camera `ci`
__device__ __forceinline__ void pose_from_params(const scnerf_camera& c, int64_t ci, Pose& P) {
  const float* e0 = c.extrinsics_initial + ci * 9;
  const float* en = c.extrinsics_noise ? c.extrinsics_noise + ci * 9 : nullptr;
  float p[9];
#pragma unroll
  for (int i = 0; i < 9; ++i)
    p[i] = en ? __fadd_rn(e0[i], __fmul_rn(c.extrinsics_noise_scale, en[i])) : e0[i];
#pragma unroll
  for (int i = 0; i < 3; ++i) { P.a[i] = p[i]; P.b[i] = p[3 + i]; P.t[i] = p[6 + i]; }
  unit_fwd(P.a, P.x, P.mag_a);
  P.xb = dot3(P.x, P.b);
  P.xx = dot3(P.x, P.x);
  P.coef = P.xb / (fmaxf(P.xx, 1e-8f) + 1e-10f);
#pragma unroll
  for (int i = 0; i < 3; ++i) P.yp[i] = P.b[i] - P.coef * P.x[i];
  unit_fwd(P.yp, P.y, P.mag_y);
  P.z[0] = P.x[1] * P.y[2] - P.x[2] * P.y[1];
  P.z[1] = P.x[2] * P.y[0] - P.x[0] * P.y[2];
  P.z[2] = P.x[0] * P.y[1] - P.x[1] * P.y[0];
}
__device__ __forceinline__ void pose_from_matrix(const float* E, Pose& P) {
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    P.x[j] = E[j * 4 + 0]; P.y[j] = E[j * 4 + 1]; P.z[j] = E[j * 4 + 2]; P.t[j] = E[j * 4 + 3];
  }
}

// d(loss)/d(columns x,y,z) -> d(loss)/d(a,b) through z = x × y and the Gram-Schmidt steps.
__device__ __forceinline__ void pose_bwd(const Pose& P, float* gx, float* gy, const float* gz,
                                         float* g_a, float* g_b) {
  // z = x × y :  gx += y × gz ; gy += gz × x
  gx[0] += P.y[1] * gz[2] - P.y[2] * gz[1];
  gx[1] += P.y[2] * gz[0] - P.y[0] * gz[2];
  gx[2] += P.y[0] * gz[1] - P.y[1] * gz[0];
  gy[0] += gz[1] * P.x[2] - gz[2] * P.x[1];
  gy[1] += gz[2] * P.x[0] - gz[0] * P.x[2];
  gy[2] += gz[0] * P.x[1] - gz[1] * P.x[0];
  float g_yp[3];
  unit_bwd(P.yp, P.mag_y, gy, g_yp);
  // yp = b - coef*x ; coef = xb / (max(xx,1e-8)+1e-10)
  float den = fmaxf(P.xx, 1e-8f) + 1e-10f;
  float g_coef = -dot3(g_yp, P.x);
  float g_xb = g_coef / den;
  float g_xx = (P.xx > 1e-8f) ? -g_coef * P.xb / (den * den) : 0.f;
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    g_b[i] = g_yp[i] + g_xb * P.x[i];
    gx[i] += -P.coef * g_yp[i] + g_xb * P.b[i] + 2.f * g_xx * P.x[i];
  }
  unit_bwd(P.a, P.mag_a, gx, g_a);
}

// F.interpolate(bilinear, align_corners=False) source taps for destination index p
// (ATen area_pixel_compute_source_index: max(scale*(p+0.5)-0.5, 0), scale = in/out in fp32).
struct Tap { int i0, i1; float w0, w1; };
__device__ __forceinline__ Tap bilinear_tap(int p, int n_in, int n_out) {
  float scale = (float)n_in / (float)n_out;
  float f = fmaxf(__fadd_rn(__fmul_rn(scale, (float)p + 0.5f), -0.5f), 0.f);
  int i0 = min((int)f, n_in - 1);
  int i1 = i0 + (i0 < n_in - 1 ? 1 : 0);
  float l1 = f - (float)i0;
  return {i0, i1, 1.f - l1, l1};
}
__device__ __forceinline__ void grid_lookup(const float* grid, int gw, const Tap& ty, const Tap& tx,
                                            float scale, float* out) {
  const float* r0 = grid + ((int64_t)ty.i0 * gw) * 3;
  const float* r1 = grid + ((int64_t)ty.i1 * gw) * 3;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    float top = tx.w0 * r0[tx.i0 * 3 + c] + tx.w1 * r0[tx.i1 * 3 + c];
    float bot = tx.w0 * r1[tx.i0 * 3 + c] + tx.w1 * r1[tx.i1 * 3 + c];
    out[c] = (ty.w0 * top + ty.w1 * bot) * scale;
  }
}
__device__ __forceinline__ void grid_scatter(float* ggrid, int gw, const Tap& ty, const Tap& tx,
                                             float scale, const float* g) {
  float* r0 = ggrid + ((int64_t)ty.i0 * gw) * 3;
  float* r1 = ggrid + ((int64_t)ty.i1 * gw) * 3;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    float v = g[c] * scale;
    atomicAdd(r0 + tx.i0 * 3 + c, v * ty.w0 * tx.w0);
    atomicAdd(r0 + tx.i1 * 3 + c, v * ty.w0 * tx.w1);
    atomicAdd(r1 + tx.i0 * 3 + c, v * ty.w1 * tx.w0);
    atomicAdd(r1 + tx.i1 * 3 + c, v * ty.w1 * tx.w1);
  }
}

struct RaygenDev {  // by-value kernel argument (camera struct copied, pointers are device)
  scnerf_camera cam;
  int has_cam;
  float focal;
  int H, W;
  const int64_t* kps;
  const float* kps_f;      // sub-pixel keypoints (used when kps == nullptr)
  const int64_t* idx;
  int64_t idx_scalar;
  const float* extrinsic;
  int extrinsic_per_ray;
  int64_t N;
};

// (u, v): the pixel coordinate the direction is computed from (float for sub-pixel keypoints, get_rays.py:112-123);
// (px, py): its truncation, which indexes the ray_o / ray_d residual fields (`.long()`, get_rays.py:134,140)
__device__ __forceinline__ void ray_pixel(const RaygenDev& a, int64_t i, int& px, int& py, float& u, float& v) {
  if (a.kps) { px = (int)a.kps[2 * i]; py = (int)a.kps[2 * i + 1]; u = (float)px; v = (float)py; }
  else if (a.kps_f) { u = a.kps_f[2 * i]; v = a.kps_f[2 * i + 1]; px = (int)u; py = (int)v; }
  else { px = (int)(i % a.W); py = (int)(i / a.W); u = (float)px; v = (float)py; }
}
// returns whether the pose is a learnable camera's; `ok` = false for a per-ray camera index out of range
__device__ __forceinline__ bool ray_pose(const RaygenDev& a, int64_t i, Pose& P, bool& ok) {
  ok = true;
  if (a.extrinsic) {
    pose_from_matrix(a.extrinsic + (a.extrinsic_per_ray ? i * 16 : 0), P);
    return false;
  }
  int64_t ci = a.idx ? a.idx[i] : a.idx_scalar;
  if ((uint64_t)ci >= (uint64_t)a.cam.n_cams) { ok = false; ci = 0; }
  pose_from_params(a.cam, ci, P);
  return true;
}

// NeRF/get_rays.py:93-148 (camera) / :75-90 (pinhole)
__global__ void __launch_bounds__(128) raygen_fwd_kernel(RaygenDev a, float* __restrict__ rays_o,
                                                         float* __restrict__ rays_d) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= a.N) return;
  int px, py;
  float u, v;
  ray_pixel(a, i, px, py, u, v);
  Pose P;
  bool pose_ok;
  ray_pose(a, i, P, pose_ok);
  float dc[3];
  if (a.has_cam) {
    Intr K = load_intrinsics(a.cam);
    // torch.inverse of the upper-triangular K: [1/fx, 0, -cx/fx; 0, 1/fy, -cy/fy; 0 0 1]
    float i00 = 1.f / K.fx, i02 = -K.cx / K.fx, i11 = 1.f / K.fy, i12 = -K.cy / K.fy;
    dc[0] = u * i00 + i02;
    dc[1] = -(v * i11 + i12);
    dc[2] = -1.f;
  } else {
    dc[0] = (u - a.W * .5f) / a.focal;
    dc[1] = -(v - a.H * .5f) / a.focal;
    dc[2] = -1.f;
  }
  float o[3], d[3];
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    d[j] = dc[0] * P.x[j] + dc[1] * P.y[j] + dc[2] * P.z[j];
    o[j] = P.t[j];
  }
  if (a.has_cam) {
    Tap ty = bilinear_tap(py, a.cam.gh, a.cam.H), tx = bilinear_tap(px, a.cam.gw, a.cam.W);
    if (a.cam.ray_o_noise) {
      float r[3];
      grid_lookup(a.cam.ray_o_noise, a.cam.gw, ty, tx, a.cam.ray_o_noise_scale, r);
#pragma unroll
      for (int j = 0; j < 3; ++j) o[j] += r[j];
    }
    if (a.cam.ray_d_noise) {
      float r[3];
      grid_lookup(a.cam.ray_d_noise, a.cam.gw, ty, tx, a.cam.ray_d_noise_scale, r);
#pragma unroll
      for (int j = 0; j < 3; ++j) d[j] += r[j];
      float inv = 1.f / (sqrtf(dot3(d, d)) + 1e-10f);  // get_rays.py:146
#pragma unroll
      for (int j = 0; j < 3; ++j) d[j] *= inv;
    }
  }
  if (!pose_ok) {   // camera index out of range: poison the ray instead of reading out of bounds
#pragma unroll
    for (int j = 0; j < 3; ++j) { o[j] = __int_as_float(0x7fc00000); d[j] = o[j]; }
  }
#pragma unroll
  for (int j = 0; j < 3; ++j) { rays_o[i * 3 + j] = o[j]; rays_d[i * 3 + j] = d[j]; }
}

__global__ void __launch_bounds__(128) raygen_bwd_kernel(RaygenDev a, const float* __restrict__ g_o,
                                                         const float* __restrict__ g_d,
                                                         scnerf_camera_grads G) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  float g_intr[4] = {0.f, 0.f, 0.f, 0.f};
  bool pose_ok = true;
  int px = 0, py = 0;
  float u = 0.f, v = 0.f;
  Pose P;
  bool learn_pose = false;
  if (i < a.N) {
    ray_pixel(a, i, px, py, u, v);
    learn_pose = ray_pose(a, i, P, pose_ok);
  }
  if (i < a.N && pose_ok) {
    Intr K = load_intrinsics(a.cam);
    float dc[3] = {u / K.fx - K.cx / K.fx, -(v / K.fy - K.cy / K.fy), -1.f};
    float go[3] = {g_o[i * 3], g_o[i * 3 + 1], g_o[i * 3 + 2]};
    float gd[3] = {g_d[i * 3], g_d[i * 3 + 1], g_d[i * 3 + 2]};
    Tap ty = bilinear_tap(py, a.cam.gh, a.cam.H), tx = bilinear_tap(px, a.cam.gw, a.cam.W);
    if (a.cam.ray_d_noise) {
      // recompute d' = R dc + residual, then back through d = d'/(|d'|+1e-10)
      float dp[3], r[3];
      grid_lookup(a.cam.ray_d_noise, a.cam.gw, ty, tx, a.cam.ray_d_noise_scale, r);
#pragma unroll
      for (int j = 0; j < 3; ++j) dp[j] = dc[0] * P.x[j] + dc[1] * P.y[j] + dc[2] * P.z[j] + r[j];
      float n = sqrtf(dot3(dp, dp)), den = n + 1e-10f;
      float k = dot3(gd, dp) / (n * den * den);
#pragma unroll
      for (int j = 0; j < 3; ++j) gd[j] = gd[j] / den - k * dp[j];
      if (G.ray_d_noise) grid_scatter(G.ray_d_noise, a.cam.gw, ty, tx, a.cam.ray_d_noise_scale, gd);
    }
    if (a.cam.ray_o_noise && G.ray_o_noise)
      grid_scatter(G.ray_o_noise, a.cam.gw, ty, tx, a.cam.ray_o_noise_scale, go);
    // d_world = x dc0 + y dc1 + z dc2
    float g_dc0 = dot3(gd, P.x), g_dc1 = dot3(gd, P.y);
    if (learn_pose && G.extrinsics_noise) {
      float gx[3], gy[3], gz[3], g_a[3], g_b[3];
#pragma unroll
      for (int j = 0; j < 3; ++j) { gx[j] = gd[j] * dc[0]; gy[j] = gd[j] * dc[1]; gz[j] = gd[j] * dc[2]; }
      pose_bwd(P, gx, gy, gz, g_a, g_b);
      int64_t ci = a.idx ? a.idx[i] : a.idx_scalar;
      float* ge = G.extrinsics_noise + ci * 9;
      float s = a.cam.extrinsics_noise_scale;
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        atomicAdd(ge + j, s * g_a[j]);
        atomicAdd(ge + 3 + j, s * g_b[j]);
        atomicAdd(ge + 6 + j, s * go[j]);
      }
    }
    // dc0 = (px - cx)/fx ; dc1 = -(py - cy)/fy
    g_intr[0] = -g_dc0 * (u - K.cx) / (K.fx * K.fx);
    g_intr[2] = -g_dc0 / K.fx;
    g_intr[1] = g_dc1 * (v - K.cy) / (K.fy * K.fy);
    g_intr[3] = g_dc1 / K.fy;
  }
  if (G.intrinsics_noise) {   // block reduction -> 4 atomics per block
    __shared__ float red[4][4];
    int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      float v = warp_sum(g_intr[k]);
      if (lane == 0) red[w][k] = v;
    }
    __syncthreads();
    if (threadIdx.x < 4) {
      int k = threadIdx.x;
      float v = red[0][k] + red[1][k] + red[2][k] + red[3][k];
      float init = a.cam.intrinsics_initial[k];
      float s = a.cam.intrinsics_noise_scale * (a.cam.multiplicative_noise ? init : 1.f);
      atomicAdd(G.intrinsics_noise + k, v * s);
    }
  }
}

// CameraModel.get_intrinsic / get_extrinsic as dense matrices (API parity; not on the hot path).
__global__ void camera_matrices_kernel(scnerf_camera c, float* K_out, float* E_out) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i == 0 && K_out) {
    Intr K = load_intrinsics(c);
    for (int k = 0; k < 16; ++k) K_out[k] = (k % 5 == 0) ? 1.f : 0.f;
    K_out[0] = K.fx; K_out[5] = K.fy; K_out[2] = K.cx; K_out[6] = K.cy;
  }
  if (i < c.n_cams && E_out) {
    Pose P;
    pose_from_params(c, i, P);
    float* E = E_out + (int64_t)i * 16;
    for (int j = 0; j < 3; ++j) {
      E[j * 4 + 0] = P.x[j]; E[j * 4 + 1] = P.y[j]; E[j * 4 + 2] = P.z[j]; E[j * 4 + 3] = P.t[j];
    }
    E[12] = 0.f; E[13] = 0.f; E[14] = 0.f; E[15] = 1.f;
  }
}

// K4 = [fx_sign*fx, fy, cx, cy] and the two camera-to-world [3,4] blocks the PRD loss projects with
// (model/ray_dist_loss.py:51-65,113-126), straight from the learnable parameters — and their backward into the camera
// gradients.  The reference (and the first version here) builds get_intrinsic() / get_extrinsic() for ALL cameras with
// ~60 eager torch kernels per call and back-propagates through them; this is one tiny launch each way.
__global__ void camera_pair_fwd_kernel(scnerf_camera c, int64_t i0, int64_t i1, float fx_sign, float* __restrict__ K4,
                                       float* __restrict__ E2) {
  const int t = threadIdx.x;
  if (t == 2) {
    Intr K = load_intrinsics(c);
    K4[0] = fx_sign * K.fx; K4[1] = K.fy; K4[2] = K.cx; K4[3] = K.cy;
  }
  if (t < 2) {
    Pose P;
    pose_from_params(c, t == 0 ? i0 : i1, P);
    float* E = E2 + t * 12;
#pragma unroll
    for (int j = 0; j < 3; ++j) { E[j * 4 + 0] = P.x[j]; E[j * 4 + 1] = P.y[j]; E[j * 4 + 2] = P.z[j]; E[j * 4 + 3] = P.t[j]; }
  }
}
__global__ void camera_pair_bwd_kernel(scnerf_camera c, int64_t i0, int64_t i1, float fx_sign, const float* __restrict__ gK4,
                                       const float* __restrict__ gE2, scnerf_camera_grads G) {
  const int t = threadIdx.x;
  if (t == 2 && G.intrinsics_noise && gK4) {
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      float init = c.intrinsics_initial[k];
      float s = c.intrinsics_noise_scale * (c.multiplicative_noise ? init : 1.f);
      atomicAdd(G.intrinsics_noise + k, gK4[k] * (k == 0 ? fx_sign : 1.f) * s);
    }
  }
  if (t < 2 && G.extrinsics_noise && gE2) {
    const int64_t ci = t == 0 ? i0 : i1;
    Pose P;
    pose_from_params(c, ci, P);
    const float* g = gE2 + t * 12;
    float gx[3], gy[3], gz[3], gt[3], g_a[3], g_b[3];
#pragma unroll
    for (int j = 0; j < 3; ++j) { gx[j] = g[j * 4 + 0]; gy[j] = g[j * 4 + 1]; gz[j] = g[j * 4 + 2]; gt[j] = g[j * 4 + 3]; }
    pose_bwd(P, gx, gy, gz, g_a, g_b);
    float* ge = G.extrinsics_noise + ci * 9;
    const float s = c.extrinsics_noise_scale;
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      atomicAdd(ge + j, s * g_a[j]);
      atomicAdd(ge + 3 + j, s * g_b[j]);
      atomicAdd(ge + 6 + j, s * gt[j]);
    }
  }
}

// ---- per-step ray batch: NeRF/run_nerf.py:368-398 — from a slice of the shuffled global
// ray indices to pixel coordinates (x, y), per-ray train-image index and target colours, in one pass.
// The reference does this with numpy on the host and three H2D copies every step.
__global__ void __launch_bounds__(256) ray_batch_kernel(const int64_t* __restrict__ shuffled, int64_t N,
                                                        const float* __restrict__ images,
                                                        const int64_t* __restrict__ i_train, int H, int W,
                                                        int64_t* __restrict__ kps, int64_t* __restrict__ img_idx,
                                                        float* __restrict__ target) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  const int64_t hw = (int64_t)H * W, g = shuffled[i];
  const int64_t t = g / hw, rem = g % hw, y = rem / W, x = rem % W;      // :372-374
  kps[2 * i] = x; kps[2 * i + 1] = y;
  img_idx[i] = t;                                                        // index into the TRAIN cameras
  const float* px = images + ((i_train[t] * H + y) * W + x) * 3;         // images[i_train[t], y, x]  (:392-395)
  target[3 * i] = px[0]; target[3 * i + 1] = px[1]; target[3 * i + 2] = px[2];
}

// ---- render()'s ray packing: viewdirs + NDC + [o d near far viewdirs] (render.py:105-130) -------
struct RayprepDev {
  scnerf_camera cam;
  int has_cam;
  float focal;
  int H, W, ndc, use_viewdirs;
  float near_, far_;
  int64_t N;
};

__global__ void __launch_bounds__(128) rayprep_fwd_kernel(RayprepDev a, const float* __restrict__ ro,
                                                          const float* __restrict__ rd,
                                                          float* __restrict__ rays) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= a.N) return;
  const int C = a.use_viewdirs ? 11 : 8;
  float o[3] = {ro[i * 3], ro[i * 3 + 1], ro[i * 3 + 2]};
  float d[3] = {rd[i * 3], rd[i * 3 + 1], rd[i * 3 + 2]};
  float* out = rays + i * C;
  if (a.use_viewdirs) {
    float inv = 1.f / sqrtf(dot3(d, d));  // render.py:107-108 (no epsilon)
    out[8] = d[0] * inv; out[9] = d[1] * inv; out[10] = d[2] * inv;
  }
  if (a.ndc) {
    float fx = a.focal, fy = a.focal;
    if (a.has_cam) { Intr K = load_intrinsics(a.cam); fx = K.fx; fy = K.fy; }
    const float near = 1.f;
    float t = -(near + o[2]) / d[2];
    float p[3] = {o[0] + t * d[0], o[1] + t * d[1], o[2] + t * d[2]};
    float sx = -1.f / (a.W / (2.f * fx)), sy = -1.f / (a.H / (2.f * fy));
    float px_pz = p[0] / p[2], py_pz = p[1] / p[2];
    o[0] = sx * px_pz; o[1] = sy * py_pz; o[2] = 1.f + 2.f * near / p[2];
    float dn0 = sx * (d[0] / d[2] - px_pz), dn1 = sy * (d[1] / d[2] - py_pz), dn2 = -2.f * near / p[2];
    d[0] = dn0; d[1] = dn1; d[2] = dn2;
  }
  out[0] = o[0]; out[1] = o[1]; out[2] = o[2];
  out[3] = d[0]; out[4] = d[1]; out[5] = d[2];
  out[6] = a.near_; out[7] = a.far_;
}

__global__ void __launch_bounds__(128) rayprep_bwd_kernel(RayprepDev a, const float* __restrict__ ro,
                                                          const float* __restrict__ rd,
                                                          const float* __restrict__ g_rays,
                                                          float* __restrict__ g_ro,
                                                          float* __restrict__ g_rd,
                                                          float* g_intr_noise) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  float g_fx = 0.f, g_fy = 0.f;
  Intr K = {a.focal, a.focal, 0.f, 0.f};
  if (a.has_cam) K = load_intrinsics(a.cam);
  if (i < a.N) {
    const int C = a.use_viewdirs ? 11 : 8;
    float o[3] = {ro[i * 3], ro[i * 3 + 1], ro[i * 3 + 2]};
    float d[3] = {rd[i * 3], rd[i * 3 + 1], rd[i * 3 + 2]};
    const float* g = g_rays + i * C;
    float go[3] = {0.f, 0.f, 0.f}, gd[3] = {0.f, 0.f, 0.f};
    if (a.use_viewdirs) {  // v = d/|d|
      float n2 = dot3(d, d), inv = rsqrtf(n2);
      float gv[3] = {g[8], g[9], g[10]};
      float k = dot3(gv, d) * inv / n2;
#pragma unroll
      for (int j = 0; j < 3; ++j) gd[j] += gv[j] * inv - k * d[j];
    }
    float gO[3] = {g[0], g[1], g[2]}, gD[3] = {g[3], g[4], g[5]};
    if (a.ndc) {
      const float near = 1.f;
      float t = -(near + o[2]) / d[2];
      float p[3] = {o[0] + t * d[0], o[1] + t * d[1], o[2] + t * d[2]};
      float sx = -2.f * K.fx / a.W, sy = -2.f * K.fy / a.H;   // = -1/(W/(2fx))
      float ipz = 1.f / p[2], idz = 1.f / d[2];
      float qx = p[0] * ipz, qy = p[1] * ipz, ex = d[0] * idz, ey = d[1] * idz;
      // outputs: O0 = sx qx, O1 = sy qy, O2 = 1 + 2/pz ; D0 = sx (ex - qx), D1 = sy (ey - qy), D2 = -2/pz
      float g_qx = sx * (gO[0] - gD[0]), g_qy = sy * (gO[1] - gD[1]);
      float g_ex = sx * gD[0], g_ey = sy * gD[1];
      float g_ipz = 2.f * near * (gO[2] - gD[2]) + g_qx * p[0] + g_qy * p[1];
      g_fx = (gO[0] * qx + gD[0] * (ex - qx)) * (-2.f / a.W);
      g_fy = (gO[1] * qy + gD[1] * (ey - qy)) * (-2.f / a.H);
      float gp[3] = {g_qx * ipz, g_qy * ipz, -g_ipz * ipz * ipz};
      float g_idz = g_ex * d[0] + g_ey * d[1];
      gd[0] += g_ex * idz; gd[1] += g_ey * idz; gd[2] += -g_idz * idz * idz;
      // p = o + t d ; t = -(near + oz)/dz
      float g_t = dot3(gp, d);
#pragma unroll
      for (int j = 0; j < 3; ++j) { go[j] += gp[j]; gd[j] += gp[j] * t; }
      go[2] += -g_t * idz;
      gd[2] += g_t * (near + o[2]) * idz * idz;
    } else {
#pragma unroll
      for (int j = 0; j < 3; ++j) { go[j] += gO[j]; gd[j] += gD[j]; }
    }
#pragma unroll
    for (int j = 0; j < 3; ++j) { g_ro[i * 3 + j] = go[j]; g_rd[i * 3 + j] = gd[j]; }
  }
  if (g_intr_noise && a.has_cam && a.ndc) {
    __shared__ float red[4][2];
    int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    float v0 = warp_sum(g_fx), v1 = warp_sum(g_fy);
    if (lane == 0) { red[w][0] = v0; red[w][1] = v1; }
    __syncthreads();
    if (threadIdx.x < 2) {
      int k = threadIdx.x;
      float v = red[0][k] + red[1][k] + red[2][k] + red[3][k];
      float init = a.cam.intrinsics_initial[k];
      float s = a.cam.intrinsics_noise_scale * (a.cam.multiplicative_noise ? init : 1.f);
      atomicAdd(g_intr_noise + k, v * s);
    }
  }
}

}  // namespace scnerf
