// Field backward, data-gradient chain: the kernel arguments and the per-ray reductions that follow the chain.
//
// The chain itself (d(raw)[4] -> rgb/alpha heads -> views layer -> feature layer -> trunk 7..0 -> d(PE) -> d(pts), d(viewdirs),
// every dZ left as a bf16 (hi[,lo]) TILE IMAGE for the wgrad kernel) is field_tc_dgrad_pipe.cuh; the serial-schedule kernel
// that used to live here was retired once the pipelined one served the 96-wide d(PE) of 4-D points as well (round 2).
//
// Graph = NeRF.forward's autograd graph (NeRF/run_nerf_helpers.py:105-128) + Embedder (:24-72).
#pragma once
#include "common.cuh"
#include "tc_prims.cuh"
#include "tc_engine.cuh"
#include "field_tc_fused.cuh"

namespace scnerf {
namespace dgrad {

using eng::TILE_M;
using fused::PlanSrc;
using fused::SrcDef;
struct Args {
  const float* rays; int ray_cols; const float* z; int64_t P; int S; int num_tiles;
  const float* pts; const float* viewdirs;   // explicit 4-D points [P,4] + per-ray directions [N,3] (XN = 96)
  const float* g_raw;                  // [P,4]
  const uint8_t* wimg; const float* cbuf;
  const uint4* relu_bits;              // forward ReLU masks, 1 bit per activation (fused::Args::relu_bits)
  eng::ImgDump out_dz[8], out_dfeat, out_dzv;   // produced for the wgrad kernel
  float* g_pts;                        // [P,3] d(loss)/d(point)   ([P,4] for 4-D points)
  float* g_vd;                         // [P,3] d(loss)/d(viewdir)
};

// d_rays[r, 0:3] += sum_s g_pts ; d_rays[r, 3:6] += sum_s z*g_pts ; d_rays[r, 8:11] += sum_s g_vd
__global__ void __launch_bounds__(128) reduce_pts_grad_kernel(const float* __restrict__ g_pts,
                                                              const float* __restrict__ g_vd,
                                                              const float* __restrict__ z, int64_t N, int S,
                                                              int ray_cols, float* __restrict__ d_rays) {
  const int lane = threadIdx.x & 31;
  const int64_t r = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (r >= N) return;
  float acc[9] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  for (int s = lane; s < S; s += 32) {
    const int64_t p = r * S + s;
    const float zz = z[p];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float g = g_pts[p * 3 + c];
      acc[c] += g; acc[3 + c] += zz * g; acc[6 + c] += g_vd[p * 3 + c];
    }
  }
#pragma unroll
  for (int k = 0; k < 9; ++k) acc[k] = warp_sum(acc[k]);
  if (lane == 0) {
#pragma unroll
    for (int k = 0; k < 9; ++k) {
      const int col = k < 6 ? k : 8 + (k - 6);
      if (col < ray_cols) d_rays[r * ray_cols + col] += acc[k];
    }
  }
}

// g_viewdirs[r, 0:3] += sum_s g_vd[r*S + s]   (explicit-point form: d(pts) stays per sample)
__global__ void __launch_bounds__(128) reduce_vd_grad_kernel(const float* __restrict__ g_vd, int64_t N, int S,
                                                             float* __restrict__ g_viewdirs) {
  const int lane = threadIdx.x & 31;
  const int64_t r = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (r >= N) return;
  float acc[3] = {0.f, 0.f, 0.f};
  for (int s = lane; s < S; s += 32) {
    const int64_t p = r * S + s;
    acc[0] += g_vd[p * 3]; acc[1] += g_vd[p * 3 + 1]; acc[2] += g_vd[p * 3 + 2];
  }
#pragma unroll
  for (int k = 0; k < 3; ++k) acc[k] = warp_sum(acc[k]);
  if (lane == 0) { g_viewdirs[r * 3] += acc[0]; g_viewdirs[r * 3 + 1] += acc[1]; g_viewdirs[r * 3 + 2] += acc[2]; }
}

}  // namespace dgrad
}  // namespace scnerf
