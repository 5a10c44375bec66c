// Projected ray distance loss: model/ray_dist_loss.py:22-246 fused into one forward
// and one backward kernel over the matches (<= a few thousand), instead of ~60 tiny einsum / elementwise /
// boolean-mask launches with device syncs.
//
// Per match: closest points p0, p1 of the two (normalised) rays (:136-165), p0 projected into image 1 and
// p1 into image 0 through the inverse extrinsics and K (:167-183), chirality mask t0 > 0 && t1 > 0
// (:188-190), squared pixel error against the matched keypoints (:207-212), then
//   train: mean over errors that are finite and below the threshold, 0.5 (loss0 + loss1)   (:214-231)
//   val/test: errors clamped to the threshold, mean over the chirality-valid matches         (:233-246)
// Latency-bound: 2 x 24 B of rays + 2 x 8 B of keypoints per match.
#pragma once
#include "common.cuh"

namespace scnerf {
namespace prd {

struct Args {
  const float *o0, *d0, *o1, *d1;   // [N,3]
  const float *kps0, *kps1;         // [N,2] (x, y) as float
  const float* K;                   // [4] fx, fy, cx, cy  (fx already negated for method "NeRF", :116-118)
  const float* E;                   // [2,3,4] camera-to-world of image 0 and image 1
  float eps, threshold;
  int train;
  int64_t N;
};
struct Match {   // everything the backward needs, recomputed in registers
  float d0[3], d1[3], n0, n1, c, w[3], a0, a1, den, t0, t1, p0[3], p1[3];
  float q0[3], q1[3], u0, v0, u1, v1, zz0, zz1, x0, y0, x1, y1, loss0, loss1;
  bool valid;
};
__device__ __forceinline__ float dot3v(const float* a, const float* b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }

__device__ __forceinline__ void eval(const Args& a, int64_t i, Match& M) {
  const float* o0 = a.o0 + i * 3; const float* o1 = a.o1 + i * 3;
  const float* r0 = a.d0 + i * 3; const float* r1 = a.d1 + i * 3;
  M.n0 = sqrtf(dot3v(r0, r0)); M.n1 = sqrtf(dot3v(r1, r1));
#pragma unroll
  for (int j = 0; j < 3; ++j) { M.d0[j] = r0[j] / (M.n0 + a.eps); M.d1[j] = r1[j] / (M.n1 + a.eps); M.w[j] = o0[j] - o1[j]; }
  M.c = dot3v(M.d0, M.d1);
  M.a0 = dot3v(M.d0, M.w); M.a1 = dot3v(M.d1, M.w);
  M.den = M.c * M.c - 1.f + a.eps;
  M.t0 = (M.a0 - M.c * M.a1) / M.den;
  M.t1 = (-M.a1 + M.c * M.a0) / M.den;            // d1.(o1-o0) - c d0.(o1-o0)
#pragma unroll
  for (int j = 0; j < 3; ++j) { M.p0[j] = M.t0 * M.d0[j] + o0[j]; M.p1[j] = M.t1 * M.d1[j] + o1[j]; }
  const float* E0 = a.E; const float* E1 = a.E + 12;
  // q = R^T (p - T): p0 into camera 1, p1 into camera 0
  float s0[3], s1[3];
#pragma unroll
  for (int j = 0; j < 3; ++j) { s0[j] = M.p0[j] - E1[j * 4 + 3]; s1[j] = M.p1[j] - E0[j * 4 + 3]; }
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    M.q0[j] = E1[0 * 4 + j] * s0[0] + E1[1 * 4 + j] * s0[1] + E1[2 * 4 + j] * s0[2];
    M.q1[j] = E0[0 * 4 + j] * s1[0] + E0[1 * 4 + j] * s1[1] + E0[2 * 4 + j] * s1[2];
  }
  const float fx = a.K[0], fy = a.K[1], cx = a.K[2], cy = a.K[3];
  M.x0 = fx * M.q0[0] + cx * M.q0[2]; M.y0 = fy * M.q0[1] + cy * M.q0[2]; M.zz0 = M.q0[2] + a.eps;
  M.x1 = fx * M.q1[0] + cx * M.q1[2]; M.y1 = fy * M.q1[1] + cy * M.q1[2]; M.zz1 = M.q1[2] + a.eps;
  M.u0 = M.x0 / M.zz0; M.v0 = M.y0 / M.zz0; M.u1 = M.x1 / M.zz1; M.v1 = M.y1 / M.zz1;
  M.valid = M.t0 > 0.f && M.t1 > 0.f;
  float e;
  e = M.u1 - a.kps0[i * 2]; M.loss0 = e * e; e = M.v1 - a.kps0[i * 2 + 1]; M.loss0 += e * e;   // p1 in image 0
  e = M.u0 - a.kps1[i * 2]; M.loss1 = e * e; e = M.v0 - a.kps1[i * 2 + 1]; M.loss1 += e * e;   // p0 in image 1
}
__device__ __forceinline__ bool keep(const Args& a, float loss) { return loss < a.threshold && isfinite(loss); }

// acc[0..1] = sum of kept loss0 / loss1, acc[2..3] = their counts, acc[4] = matches kept by both (train) or the
// chirality-valid count (val)
__global__ void __launch_bounds__(128) fwd_kernel(Args a, float* __restrict__ acc) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  float v[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
  if (i < a.N) {
    Match M;
    eval(a, i, M);
    if (M.valid) {
      if (a.train) {
        bool k0 = keep(a, M.loss0), k1 = keep(a, M.loss1);
        if (k0) { v[0] = M.loss0; v[2] = 1.f; }
        if (k1) { v[1] = M.loss1; v[3] = 1.f; }
        v[4] = (k0 && k1) ? 1.f : 0.f;
      } else {
        v[0] = keep(a, M.loss0) && !(M.loss0 > a.threshold) ? M.loss0 : a.threshold;
        v[1] = keep(a, M.loss1) && !(M.loss1 > a.threshold) ? M.loss1 : a.threshold;
        v[2] = 1.f; v[3] = 1.f; v[4] = 1.f;
      }
    }
  }
  __shared__ float red[4][5];
  int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < 5; ++k) {
    float s = warp_sum(v[k]);
    if (lane == 0) red[w][k] = s;
  }
  __syncthreads();
  if (threadIdx.x < 5) atomicAdd(acc + threadIdx.x, red[0][threadIdx.x] + red[1][threadIdx.x] + red[2][threadIdx.x] + red[3][threadIdx.x]);
}
// loss = 0.5 (acc0/acc2 + acc1/acc3): empty selections give nan like torch's mean of an empty tensor
__global__ void finalize_kernel(const float* __restrict__ acc, float* __restrict__ loss, float* __restrict__ n_match) {
  if (threadIdx.x == 0) { loss[0] = 0.5f * (acc[0] / acc[2] + acc[1] / acc[3]); if (n_match) n_match[0] = acc[4]; }
}

// backward (train mode): g_out = d(total)/d(loss).  Ray gradients are written (overwrite); K and E gradients
// accumulate with block reductions + atomics (gK[4], gE[2,3,4]).
__global__ void __launch_bounds__(128) bwd_kernel(Args a, const float* __restrict__ acc, const float* __restrict__ g_out,
                                                  float* __restrict__ g_o0, float* __restrict__ g_d0,
                                                  float* __restrict__ g_o1, float* __restrict__ g_d1,
                                                  float* __restrict__ gK, float* __restrict__ gE) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  float gk[4] = {0.f, 0.f, 0.f, 0.f}, ge[24];
#pragma unroll
  for (int k = 0; k < 24; ++k) ge[k] = 0.f;
  if (i < a.N) {
    Match M;
    eval(a, i, M);
    float go0[3] = {0.f, 0.f, 0.f}, go1[3] = {0.f, 0.f, 0.f}, gd0[3] = {0.f, 0.f, 0.f}, gd1[3] = {0.f, 0.f, 0.f};
    const float gl0 = (M.valid && keep(a, M.loss0)) ? 0.5f * g_out[0] / acc[2] : 0.f;   // loss0: p1 -> image 0
    const float gl1 = (M.valid && keep(a, M.loss1)) ? 0.5f * g_out[0] / acc[3] : 0.f;   // loss1: p0 -> image 1
    if (gl0 != 0.f || gl1 != 0.f) {
      const float fx = a.K[0], fy = a.K[1], cx = a.K[2], cy = a.K[3];
      const float* E0 = a.E; const float* E1 = a.E + 12;
      float gp0[3] = {0.f, 0.f, 0.f}, gp1[3] = {0.f, 0.f, 0.f};
      auto project_bwd = [&](float gl, float u, float v, float x, float y, float zz, const float* q, const float* p,
                             const float* kp, const float* Ecam, float* gEcam, float* gp) {
        if (gl == 0.f) return;
        const float gu = 2.f * (u - kp[0]) * gl, gv = 2.f * (v - kp[1]) * gl;
        const float gx = gu / zz, gy = gv / zz, gz = -(gu * x + gv * y) / (zz * zz);
        gk[0] += gx * q[0]; gk[2] += gx * q[2]; gk[1] += gy * q[1]; gk[3] += gy * q[2];
        const float gq[3] = {gx * fx, gy * fy, gx * cx + gy * cy + gz};
        // q = R^T (p - T)
#pragma unroll
        for (int r = 0; r < 3; ++r) {
          const float gpr = Ecam[r * 4 + 0] * gq[0] + Ecam[r * 4 + 1] * gq[1] + Ecam[r * 4 + 2] * gq[2];   // (R gq)_r
          gp[r] += gpr;
          gEcam[r * 4 + 3] -= gpr;
          const float s = p[r] - Ecam[r * 4 + 3];
#pragma unroll
          for (int c = 0; c < 3; ++c) gEcam[r * 4 + c] += s * gq[c];
        }
      };
      project_bwd(gl1, M.u0, M.v0, M.x0, M.y0, M.zz0, M.q0, M.p0, a.kps1 + i * 2, E1, ge + 12, gp0);
      project_bwd(gl0, M.u1, M.v1, M.x1, M.y1, M.zz1, M.q1, M.p1, a.kps0 + i * 2, E0, ge, gp1);
      // p0 = t0 d0 + o0 ; p1 = t1 d1 + o1
      const float gt0 = dot3v(gp0, M.d0), gt1 = dot3v(gp1, M.d1);
#pragma unroll
      for (int j = 0; j < 3; ++j) { gd0[j] += M.t0 * gp0[j]; gd1[j] += M.t1 * gp1[j]; go0[j] += gp0[j]; go1[j] += gp1[j]; }
      // t0 = (a0 - c a1)/den ; t1 = (c a0 - a1)/den ; den = c^2 - 1 + eps
      const float ga0 = (gt0 + gt1 * M.c) / M.den, ga1 = -(gt0 * M.c + gt1) / M.den;
      const float gc = (-gt0 * M.a1 + gt1 * M.a0) / M.den - (gt0 * M.t0 + gt1 * M.t1) * 2.f * M.c / M.den;
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        const float gw = ga0 * M.d0[j] + ga1 * M.d1[j];
        gd0[j] += ga0 * M.w[j] + gc * M.d1[j];
        gd1[j] += ga1 * M.w[j] + gc * M.d0[j];
        go0[j] += gw; go1[j] -= gw;
      }
      // d = r / (|r| + eps)
      const float* r0 = a.d0 + i * 3; const float* r1 = a.d1 + i * 3;
      const float k0 = dot3v(gd0, r0) / (M.n0 * (M.n0 + a.eps) * (M.n0 + a.eps));
      const float k1 = dot3v(gd1, r1) / (M.n1 * (M.n1 + a.eps) * (M.n1 + a.eps));
#pragma unroll
      for (int j = 0; j < 3; ++j) { gd0[j] = gd0[j] / (M.n0 + a.eps) - k0 * r0[j]; gd1[j] = gd1[j] / (M.n1 + a.eps) - k1 * r1[j]; }
    }
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      g_o0[i * 3 + j] = go0[j]; g_d0[i * 3 + j] = gd0[j]; g_o1[i * 3 + j] = go1[j]; g_d1[i * 3 + j] = gd1[j];
    }
  }
  __shared__ float red[4][28];
  int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < 4; ++k) { float s = warp_sum(gk[k]); if (lane == 0) red[w][k] = s; }
#pragma unroll
  for (int k = 0; k < 24; ++k) { float s = warp_sum(ge[k]); if (lane == 0) red[w][4 + k] = s; }
  __syncthreads();
  if (threadIdx.x < 28) {
    const int k = threadIdx.x;
    const float s = red[0][k] + red[1][k] + red[2][k] + red[3][k];
    if (k < 4) { if (gK) atomicAdd(gK + k, s); }
    else if (gE) atomicAdd(gE + (k - 4), s);
  }
}

}  // namespace prd
}  // namespace scnerf
