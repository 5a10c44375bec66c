// raw2outputs (NeRF/render.py:302-355) forward and backward: one warp per ray, chunked warp
// scans for the exclusive cumulative product.  Stand-alone this is HBM-bound
// (S*(16+4) B in per ray).
#pragma once
#include "common.cuh"

namespace scnerf {

struct CompositeArgs {
  const float* raw; int raw_cols;       // [N,S,raw_cols], channels 0..2 rgb, 3 sigma
  const float* z;                       // [N,S]
  const float* rays_d; int d_stride;    // direction of ray r at rays_d + r*d_stride
  const float* noise; float noise_std;  // unit noise [N,S] (NULL + std>0 -> Philox normal)
  uint64_t seed; uint32_t rng_stream;
  int white_bkgd;
  int64_t N; int S;
  float *rgb_map, *disp_map, *acc_map, *weights, *depth_map;  // outputs (weights/depth optional)
};

__device__ __forceinline__ float warp_incl_prod(float v, int lane) {
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    float n = __shfl_up_sync(0xffffffffu, v, o);
    if (lane >= o) v *= n;
  }
  return v;
}

__device__ __forceinline__ float sigma_noise(const CompositeArgs& a, int64_t g) {
  if (a.noise_std <= 0.f) return 0.f;
  float n = a.noise ? a.noise[g] : Philox::normal(a.seed, a.rng_stream, (uint64_t)g);
  return n * a.noise_std;
}

// One ray, one warp: `raw_ray` points at the ray's [S, raw_cols] raw values — global memory for the stand-alone kernel,
// shared memory when the fused field forward composites a ray group from its own epilogue (fpipe::field_fwd_pipe_kernel).
__device__ __forceinline__ void composite_ray(const CompositeArgs& a, int64_t r, int lane, const float* raw_ray, int raw_cols) {
  const float* d = a.rays_d + r * a.d_stride;
  const float dn = sqrtf(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
  float carry = 1.f;  // prod_{j<chunk start} (1 - alpha_j + 1e-10)
  float acc_r = 0.f, acc_g = 0.f, acc_b = 0.f, acc_w = 0.f, acc_z = 0.f;
  for (int s0 = 0; s0 < a.S; s0 += 32) {
    int s = s0 + lane;
    float alpha = 0.f, zs = 0.f, cr = 0.f, cg = 0.f, cb = 0.f;
    if (s < a.S) {
      int64_t g = r * a.S + s;
      zs = a.z[g];
      float dist = (s + 1 < a.S) ? a.z[g + 1] - zs : 1e10f;
      dist *= dn;
      const float* rw = raw_ray + (int64_t)s * raw_cols;
      float sg = rw[3] + sigma_noise(a, g);
      alpha = 1.f - expf(-fmaxf(sg, 0.f) * dist);
      cr = 1.f / (1.f + expf(-rw[0]));
      cg = 1.f / (1.f + expf(-rw[1]));
      cb = 1.f / (1.f + expf(-rw[2]));
    }
    float f = (s < a.S) ? (1.f - alpha + 1e-10f) : 1.f;
    float incl = warp_incl_prod(f, lane);
    float excl = __shfl_up_sync(0xffffffffu, incl, 1);
    if (lane == 0) excl = 1.f;
    float T = carry * excl;
    float w = alpha * T;
    carry *= __shfl_sync(0xffffffffu, incl, 31);
    if (s < a.S) {
      if (a.weights) a.weights[r * a.S + s] = w;
      acc_r += w * cr; acc_g += w * cg; acc_b += w * cb; acc_w += w; acc_z += w * zs;
    }
  }
  acc_r = warp_sum(acc_r); acc_g = warp_sum(acc_g); acc_b = warp_sum(acc_b);
  acc_w = warp_sum(acc_w); acc_z = warp_sum(acc_z);
  if (lane == 0) {
    if (a.white_bkgd) { acc_r += 1.f - acc_w; acc_g += 1.f - acc_w; acc_b += 1.f - acc_w; }
    a.rgb_map[r * 3] = acc_r; a.rgb_map[r * 3 + 1] = acc_g; a.rgb_map[r * 3 + 2] = acc_b;
    a.disp_map[r] = 1.f / fmaxf(1e-10f, acc_z / (acc_w + 1e-10f));
    a.acc_map[r] = acc_w;
    if (a.depth_map) a.depth_map[r] = acc_z;
  }
}

__global__ void __launch_bounds__(128) composite_fwd_kernel(CompositeArgs a) {
  const int lane = threadIdx.x & 31;
  const int64_t r = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (r >= a.N) return;
  composite_ray(a, r, lane, a.raw + r * a.S * a.raw_cols, a.raw_cols);
}

struct CompositeBwdArgs {
  CompositeArgs f;            // forward inputs (outputs unused except weights == saved weights)
  const float *g_rgb, *g_disp, *g_acc;  // [N,3] [N] [N] (NULL = 0)
  const float* acc_saved;     // [N] forward acc_map
  const float* depth_saved;   // [N] forward depth_map
  float* g_raw;               // [N,S,4] overwrite
  float* g_rays;  int g_rays_cols;  // d(loss)/d(rays)[r, 3:6] += via |d|   (or NULL)
};

// Backward: per ray recompute alpha and T (two chunked scans: forward product, reverse sum).
__global__ void __launch_bounds__(128) composite_bwd_kernel(CompositeBwdArgs b) {
  extern __shared__ float smem[];
  const CompositeArgs& a = b.f;
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int64_t r = (int64_t)blockIdx.x * (blockDim.x >> 5) + wib;
  if (r >= a.N) return;
  float* T_s = smem + (size_t)wib * 2 * a.S;  // transmittance
  float* A_s = T_s + a.S;                      // alpha
  const float* d = a.rays_d + r * a.d_stride;
  const float dn = sqrtf(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
  float carry = 1.f;
  for (int s0 = 0; s0 < a.S; s0 += 32) {
    int s = s0 + lane;
    float alpha = 0.f;
    if (s < a.S) {
      int64_t g = r * a.S + s;
      float dist = ((s + 1 < a.S) ? a.z[g + 1] - a.z[g] : 1e10f) * dn;
      float sg = a.raw[g * a.raw_cols + 3] + sigma_noise(a, g);
      alpha = 1.f - expf(-fmaxf(sg, 0.f) * dist);
    }
    float f = (s < a.S) ? (1.f - alpha + 1e-10f) : 1.f;
    float incl = warp_incl_prod(f, lane);
    float excl = __shfl_up_sync(0xffffffffu, incl, 1);
    if (lane == 0) excl = 1.f;
    if (s < a.S) { T_s[s] = carry * excl; A_s[s] = alpha; }
    carry *= __shfl_sync(0xffffffffu, incl, 31);
  }
  __syncwarp();
  // upstream gradients of the per-ray reductions
  float gr = b.g_rgb ? b.g_rgb[r * 3] : 0.f, gg = b.g_rgb ? b.g_rgb[r * 3 + 1] : 0.f,
        gb = b.g_rgb ? b.g_rgb[r * 3 + 2] : 0.f;
  float g_acc = b.g_acc ? b.g_acc[r] : 0.f, g_depth = 0.f;
  if (a.white_bkgd) g_acc -= (gr + gg + gb);
  if (b.g_disp) {
    float acc = b.acc_saved[r], depth = b.depth_saved[r];
    float q = depth / (acc + 1e-10f);
    if (q > 1e-10f) {
      float g_q = -b.g_disp[r] / (q * q);
      g_depth = g_q / (acc + 1e-10f);
      g_acc += -g_q * depth / ((acc + 1e-10f) * (acc + 1e-10f));
    }
  }
  // reverse pass: suffix = sum_{j>s} g_w_j * w_j
  float suffix_carry = 0.f, g_dn = 0.f;
  const int nchunk = (a.S + 31) / 32;
  for (int c = nchunk - 1; c >= 0; --c) {
    int s = c * 32 + lane;
    float gw = 0.f, w = 0.f, alpha = 0.f, T = 0.f, zs = 0.f, cr = 0.f, cg = 0.f, cb = 0.f;
    int64_t g = r * a.S + s;
    if (s < a.S) {
      alpha = A_s[s]; T = T_s[s]; w = alpha * T; zs = a.z[g];
      const float* rw = a.raw + g * a.raw_cols;
      cr = 1.f / (1.f + expf(-rw[0])); cg = 1.f / (1.f + expf(-rw[1])); cb = 1.f / (1.f + expf(-rw[2]));
      gw = gr * cr + gg * cg + gb * cb + g_depth * zs + g_acc;
    }
    // inclusive reverse scan of gw*w within the chunk
    float v = gw * w, incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      float n = __shfl_down_sync(0xffffffffu, incl, o);
      if (lane + o < 32) incl += n;
    }
    float suffix = suffix_carry + (incl - v);  // strictly after s
    suffix_carry += __shfl_sync(0xffffffffu, incl, 0);
    if (s < a.S) {
      float g_alpha = gw * T - suffix / (1.f - alpha + 1e-10f);
      float dist = ((s + 1 < a.S) ? a.z[g + 1] - zs : 1e10f);
      float sg = a.raw[g * a.raw_cols + 3] + sigma_noise(a, g);
      float rs = fmaxf(sg, 0.f);
      float e = expf(-rs * dist * dn);           // = 1 - alpha
      float g_sigma = (sg > 0.f) ? g_alpha * dist * dn * e : 0.f;
      g_dn += g_alpha * rs * dist * e;
      float* go = b.g_raw + g * 4;
      go[0] = w * gr * cr * (1.f - cr);
      go[1] = w * gg * cg * (1.f - cg);
      go[2] = w * gb * cb * (1.f - cb);
      go[3] = g_sigma;
    }
  }
  if (b.g_rays) {
    g_dn = warp_sum(g_dn);
    if (lane < 3 && dn > 0.f) b.g_rays[r * b.g_rays_cols + 3 + lane] += g_dn * d[lane] / dn;
  }
}

// loss = mean((min(rgb,1)-t)^2) [+ same for rgb0]; writes d(loss)/d(rgb) with the reference's
// in-place saturation semantics (render.py:404-406: channels >= 1 are overwritten, so no gradient).
__global__ void __launch_bounds__(256) mse_loss_kernel(const float* __restrict__ rgb,
                                                       const float* __restrict__ rgb0,
                                                       const float* __restrict__ target, int64_t n,
                                                       float* __restrict__ g_rgb,
                                                       float* __restrict__ g_rgb0, float* loss) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  float l = 0.f;
  if (i < n) {
    float t = target[i], inv = 1.f / (float)n;
    float v = rgb[i];
    if (v >= 1.f) { l += (1.f - t) * (1.f - t); g_rgb[i] = 0.f; }
    else { l += (v - t) * (v - t); g_rgb[i] = 2.f * (v - t) * inv; }
    if (rgb0) {
      v = rgb0[i];
      if (v >= 1.f) { l += (1.f - t) * (1.f - t); g_rgb0[i] = 0.f; }
      else { l += (v - t) * (v - t); g_rgb0[i] = 2.f * (v - t) * inv; }
    }
    l *= inv;
  }
  __shared__ float red[8];
  l = warp_sum(l);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = l;
  __syncthreads();
  if (threadIdx.x == 0) {
    float s = 0.f;
    for (int k = 0; k < 8; ++k) s += red[k];
    atomicAdd(loss, s);
  }
}

}  // namespace scnerf
