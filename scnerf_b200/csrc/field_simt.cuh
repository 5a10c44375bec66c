// SCNERF_PRECISION_FP32 field: positional encoding + NeRF MLP forward/backward, layer by layer on
// CUDA cores with activations in the caller's workspace.  This is the exact-fp32 parity anchor
// (and the backward engine until the fused tcgen05 backward replaces it); the production forward
// is the fused tcgen05 kernel in field_tc.cuh.
//
// Reference: run_network NeRF/create_nerf.py:18-32, Embedder NeRF/run_nerf_helpers.py:24-72,
// NeRF.forward NeRF/run_nerf_helpers.py:105-128.
#pragma once
#include <algorithm>
#include "common.cuh"
#include "gemm_simt.cuh"

namespace scnerf {

// ---- positional encoding ------------------------------------------------------------------------
// points are formed on the fly from (rays, z): pts = o + d*z  (render.py:259); one thread per
// (point, band) where band 0 is the identity and band f+1 is [sin(2^f x), cos(2^f x)].
__global__ void __launch_bounds__(256) pe_points_kernel(const float* __restrict__ rays, int ray_cols,
                                                        const float* __restrict__ z,
                                                        const float* __restrict__ pts_in,
                                                        int64_t P, int S, int L,
                                                        float* __restrict__ out, int64_t ldo, int dim = 3) {
  int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int nb = L + 1;
  if (g >= P * nb) return;
  int64_t p = g / nb;
  int band = (int)(g % nb);
  if (dim == 4) {   // NeRF++ background points (x, y, z, 1/r): nerfplusplus/nerf_network.py:42-60
    const float* xi = pts_in + p * 4;
    float* o = out + p * ldo;
    if (band == 0) { o[0] = xi[0]; o[1] = xi[1]; o[2] = xi[2]; o[3] = xi[3]; }
    else {
      float f = (float)(1 << (band - 1));
      o += 4 + 8 * (band - 1);
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        float sn, cs;
        sincosf(xi[c] * f, &sn, &cs);
        o[c] = sn; o[4 + c] = cs;
      }
    }
    return;
  }
  float x[3];
  if (pts_in) {
    x[0] = pts_in[p * 3]; x[1] = pts_in[p * 3 + 1]; x[2] = pts_in[p * 3 + 2];
  } else {
    int64_t r = p / S;
    const float* ry = rays + r * ray_cols;
    float zz = z[p];
#pragma unroll
    for (int c = 0; c < 3; ++c) x[c] = __fadd_rn(ry[c], __fmul_rn(ry[3 + c], zz));
  }
  float* o = out + p * ldo;
  if (band == 0) {
    o[0] = x[0]; o[1] = x[1]; o[2] = x[2];
  } else {
    float f = (float)(1 << (band - 1));
    o += 3 + 6 * (band - 1);
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      float s, cs;
      sincosf(x[c] * f, &s, &cs);
      o[c] = s; o[3 + c] = cs;
    }
  }
}

// PE of the per-ray view direction, broadcast over the ray's samples (create_nerf.py:24-28)
__global__ void __launch_bounds__(256) pe_dirs_kernel(const float* __restrict__ dirs, int64_t dir_stride,
                                                      int64_t P, int S, int L, float* __restrict__ out,
                                                      int64_t ldo) {
  int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int nb = L + 1;
  if (g >= P * nb) return;
  int64_t p = g / nb;
  int band = (int)(g % nb);
  const float* v = dirs + (p / S) * dir_stride;
  float* o = out + p * ldo;
  if (band == 0) {
    o[0] = v[0]; o[1] = v[1]; o[2] = v[2];
  } else {
    float f = (float)(1 << (band - 1));
    o += 3 + 6 * (band - 1);
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      float s, cs;
      sincosf(v[c] * f, &s, &cs);
      o[c] = s; o[3 + c] = cs;
    }
  }
}

// d(PE)/d(x) contraction for one 3-vector: returns dL/dx given dL/d(PE row)
__device__ __forceinline__ void pe_bwd_point(const float* x, const float* g, int L, float* gx) {
#pragma unroll
  for (int c = 0; c < 3; ++c) gx[c] = g[c];
  for (int f = 0; f < L; ++f) {
    float fr = (float)(1 << f);
    const float* gs = g + 3 + 6 * f;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      float s, cs;
      sincosf(x[c] * fr, &s, &cs);
      gx[c] += fr * (cs * gs[c] - s * gs[3 + c]);
    }
  }
}

// Backward of both encodings, reduced per ray: one CTA per ray, one thread per sample.
//   d_rays[r, 0:3] += sum_s dpts ; d_rays[r, 3:6] += sum_s z*dpts ; d_rays[r, 8:11] += d(viewdirs)
__global__ void __launch_bounds__(256) pe_bwd_kernel(const float* __restrict__ rays, int ray_cols,
                                                     const float* __restrict__ z, int S, int L_pos,
                                                     int L_dir, const float* __restrict__ g_pe,
                                                     int64_t ld_gpe, const float* __restrict__ g_ped,
                                                     int64_t ld_gped, float* __restrict__ d_rays,
                                                     float* __restrict__ g_z = nullptr) {
  const int64_t r = blockIdx.x;
  const float* ry = rays + r * ray_cols;
  float acc[9] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  for (int s = threadIdx.x; s < S; s += blockDim.x) {
    int64_t p = r * S + s;
    float zz = z[p], x[3], gx[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) x[c] = __fadd_rn(ry[c], __fmul_rn(ry[3 + c], zz));
    pe_bwd_point(x, g_pe + p * ld_gpe, L_pos, gx);
#pragma unroll
    for (int c = 0; c < 3; ++c) { acc[c] += gx[c]; acc[3 + c] += zz * gx[c]; }
    if (g_z) g_z[p] = gx[0] * ry[3] + gx[1] * ry[4] + gx[2] * ry[5];   // pts = o + z d
    if (g_ped) {
      // sum the raw PE-dir gradients over samples first is equivalent (PE is per ray), but the
      // contraction is linear so do it per sample and reduce the 3-vector
      pe_bwd_point(ry + 8, g_ped + p * ld_gped, L_dir, gx);
#pragma unroll
      for (int c = 0; c < 3; ++c) acc[6 + c] += gx[c];
    }
  }
  __shared__ float red[8][9];
  int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < 9; ++k) {
    float v = warp_sum(acc[k]);
    if (lane == 0) red[w][k] = v;
  }
  __syncthreads();
  if (threadIdx.x < 9) {
    int k = threadIdx.x, nw = blockDim.x >> 5;
    float v = 0.f;
    for (int i = 0; i < nw; ++i) v += red[i][k];
    int col = k < 6 ? k : 8 + (k - 6);
    if (col < ray_cols) d_rays[r * ray_cols + col] += v;
  }
}

// Backward of the encodings for EXPLICIT points [P,dim] and per-ray view directions [N,3]:
// g_pts[p, 0:dim] (overwrite), g_viewdirs[r, 0:3] (+=).  One CTA per ray.
__global__ void __launch_bounds__(128) pe_bwd_pts_kernel(const float* __restrict__ pts, int dim,
                                                         const float* __restrict__ viewdirs, int S,
                                                         int L_pos, int L_dir,
                                                         const float* __restrict__ g_pe, int64_t ld_gpe,
                                                         const float* __restrict__ g_ped, int64_t ld_gped,
                                                         float* __restrict__ g_pts,
                                                         float* __restrict__ g_viewdirs) {
  const int64_t r = blockIdx.x;
  float acc[3] = {0.f, 0.f, 0.f};
  for (int s = threadIdx.x; s < S; s += blockDim.x) {
    const int64_t p = r * S + s;
    if (g_pts) {
      const float* x = pts + p * dim;
      const float* g = g_pe + p * ld_gpe;
      for (int c = 0; c < dim; ++c) {
        float gx = g[c];
        for (int f = 0; f < L_pos; ++f) {
          float fr = (float)(1 << f), sn, cs;
          sincosf(x[c] * fr, &sn, &cs);
          const float* gs = g + dim + 2 * dim * f;
          gx += fr * (cs * gs[c] - sn * gs[dim + c]);
        }
        g_pts[p * dim + c] = gx;
      }
    }
    if (g_viewdirs && g_ped) {
      float gx[3];
      pe_bwd_point(viewdirs + r * 3, g_ped + p * ld_gped, L_dir, gx);
      acc[0] += gx[0]; acc[1] += gx[1]; acc[2] += gx[2];
    }
  }
  if (g_viewdirs && g_ped) {
    __shared__ float red[4][3];
    int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      float v = warp_sum(acc[k]);
      if (lane == 0) red[w][k] = v;
    }
    __syncthreads();
    if (threadIdx.x < 3) {
      int k = threadIdx.x;
      g_viewdirs[r * 3 + k] += red[0][k] + red[1][k] + red[2][k] + red[3][k];
    }
  }
}

// ---- narrow heads (alpha: 1, rgb: 3, output_linear: 4|5 outputs) ------------------------------
constexpr int HEAD_MAX = 8;
// out[p, col0+o] = x[p,:] . W[o,:] + b[o];  8 lanes per row
__global__ void __launch_bounds__(256) head_fwd_kernel(const float* __restrict__ x, int64_t ldx,
                                                       const float* __restrict__ W,
                                                       const float* __restrict__ b, int64_t P, int K,
                                                       int n_out, float* __restrict__ out,
                                                       int64_t ldo, int col0) {
  int64_t p = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 3;
  int sub = threadIdx.x & 7;
  float acc[HEAD_MAX];
#pragma unroll
  for (int o = 0; o < HEAD_MAX; ++o) acc[o] = 0.f;
  if (p < P) {
    const float* xr = x + p * ldx;
    for (int k = sub; k < K; k += 8) {
      float xv = xr[k];
#pragma unroll
      for (int o = 0; o < HEAD_MAX; ++o)
        if (o < n_out) acc[o] = fmaf(xv, W[o * K + k], acc[o]);
    }
  }
#pragma unroll
  for (int o = 0; o < HEAD_MAX; ++o) {
    acc[o] += __shfl_xor_sync(0xffffffffu, acc[o], 1);
    acc[o] += __shfl_xor_sync(0xffffffffu, acc[o], 2);
    acc[o] += __shfl_xor_sync(0xffffffffu, acc[o], 4);
  }
  if (p < P && sub == 0)
    for (int o = 0; o < n_out; ++o) out[p * ldo + col0 + o] = acc[o] + b[o];
}

// dx[p,k] (+)= sum_o g[p, gcol0+o] W[o,k], then optional relu mask (mask[p,k] > 0)
__global__ void __launch_bounds__(256) head_dgrad_kernel(const float* __restrict__ g, int64_t ldg,
                                                         int gcol0, const float* __restrict__ W,
                                                         int64_t P, int K, int n_out,
                                                         float* __restrict__ dx, int64_t lddx,
                                                         int accumulate,
                                                         const float* __restrict__ mask,
                                                         int64_t ldmask) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= P * K) return;
  int64_t p = i / K;
  int k = (int)(i % K);
  float v = accumulate ? dx[p * lddx + k] : 0.f;
  for (int o = 0; o < n_out; ++o) v = fmaf(g[p * ldg + gcol0 + o], W[o * K + k], v);
  if (mask && !(mask[p * ldmask + k] > 0.f)) v = 0.f;
  dx[p * lddx + k] = v;
}

// dW[o,k] += sum_p g[p,gcol0+o] x[p,k] ; db[o] += sum_p g[p,gcol0+o].   thread <-> k
__global__ void __launch_bounds__(256) head_wgrad_kernel(const float* __restrict__ g, int64_t ldg,
                                                         int gcol0, const float* __restrict__ x,
                                                         int64_t ldx, int64_t P, int K, int n_out,
                                                         int64_t rows_per_block,
                                                         float* __restrict__ dW,
                                                         float* __restrict__ db) {
  int64_t p0 = (int64_t)blockIdx.x * rows_per_block, p1 = min(P, p0 + rows_per_block);
  float acc[HEAD_MAX];
  for (int k = threadIdx.x; k < K; k += blockDim.x) {
#pragma unroll
    for (int o = 0; o < HEAD_MAX; ++o) acc[o] = 0.f;
    for (int64_t p = p0; p < p1; ++p) {
      float xv = x[p * ldx + k];
#pragma unroll
      for (int o = 0; o < HEAD_MAX; ++o)
        if (o < n_out) acc[o] = fmaf(g[p * ldg + gcol0 + o], xv, acc[o]);
    }
    for (int o = 0; o < n_out; ++o) atomicAdd(dW + o * K + k, acc[o]);
  }
  if (threadIdx.x < n_out) {
    float s = 0.f;
    for (int64_t p = p0; p < p1; ++p) s += g[p * ldg + gcol0 + threadIdx.x];
    atomicAdd(db + threadIdx.x, s);
  }
}

// ---- buffers --------------------------------------------------------------------------------------
struct FieldBufs {
  float* X5 = nullptr;                 // [P, ldx5]  = [PE(pts) | h_skip]
  float* H[SCNERF_MAX_DEPTH] = {};     // h_i [P, W]   (H[skip] unused: lives in X5)
  float* F = nullptr;                  // [P, W + icv] = [feature | PE(dir)]
  float* HV = nullptr;                 // [P, W/2]
  int64_t ldx5 = 0, ldf = 0;
  bool keep_all = false;               // training layout (every H[i] distinct)
  uint8_t* tc_img = nullptr;           // packed bf16 weight image for the tcgen05 path
  float* tc_cbuf = nullptr;            // packed biases / head weights
};
constexpr size_t TC_IMG_BYTES = 2560 * 1024;   // >= fused::weight_image_bytes<3>() (checked there)
constexpr size_t TC_CBUF_FLOATS = 4096;
struct FieldGradBufs {
  float *Ga = nullptr, *Gb = nullptr;  // [P, W] ping-pong
  float* Gx5 = nullptr;                // [P, ldx5]
  float* Gf = nullptr;                 // [P, W + icv]
  float* Ghv = nullptr;                // [P, W/2]
};

inline void field_bufs_alloc(Arena& ar, const scnerf_mlp& m, int64_t P, bool keep_all, FieldBufs& B,
                             bool tc_only = false) {
  B.ldx5 = m.input_ch + m.W;
  B.ldf = m.W + m.input_ch_views;
  B.keep_all = keep_all;
  B.tc_img = ar.get<uint8_t>(TC_IMG_BYTES);
  B.tc_cbuf = ar.get<float>(TC_CBUF_FLOATS);
  if (tc_only) return;     // tensor-core training path keeps bf16 tile images instead (TcTrainBufs)
  B.X5 = ar.get<float>(P * B.ldx5);
  if (keep_all) {
    for (int i = 0; i < m.D; ++i)
      if (i != m.skip) B.H[i] = ar.get<float>(P * m.W);
  } else {  // inference: two rotating buffers
    float* r0 = ar.get<float>(P * m.W);
    float* r1 = ar.get<float>(P * m.W);
    int k = 0;
    for (int i = 0; i < m.D; ++i)
      if (i != m.skip) B.H[i] = (k++ & 1) ? r1 : r0;
  }
  if (m.use_viewdirs) {
    B.F = ar.get<float>(P * B.ldf);
    B.HV = ar.get<float>(P * (m.W / 2));
  }
}
inline void field_grad_bufs_alloc(Arena& ar, const scnerf_mlp& m, int64_t P, FieldGradBufs& G) {
  G.Ga = ar.get<float>(P * m.W);
  G.Gb = ar.get<float>(P * m.W);
  G.Gx5 = ar.get<float>(P * (m.input_ch + m.W));
  if (m.use_viewdirs) {
    G.Gf = ar.get<float>(P * (m.W + m.input_ch_views));
    G.Ghv = ar.get<float>(P * (m.W / 2));
  }
}

inline int field_check(const scnerf_mlp& m) {
  SCNERF_CHECK_ARG(m.D >= 2 && m.D <= SCNERF_MAX_DEPTH, "mlp depth %d unsupported", m.D);
  SCNERF_CHECK_ARG(m.skip < m.D - 1, "skip connection after the last trunk layer unsupported");
  const int dim = m.pts_dim == 4 ? 4 : 3;
  SCNERF_CHECK_ARG(m.pts_dim == 0 || m.pts_dim == 3 || m.pts_dim == 4, "pts_dim %d unsupported", m.pts_dim);
  SCNERF_CHECK_ARG(m.input_ch == dim * (1 + 2 * m.L_pos), "input_ch %d != %d*(1+2*%d)", m.input_ch, dim, m.L_pos);
  if (m.use_viewdirs)
    SCNERF_CHECK_ARG(m.input_ch_views == 3 + 6 * m.L_dir, "input_ch_views mismatch");
  else
    SCNERF_CHECK_ARG(m.output_ch >= 4 && m.output_ch <= HEAD_MAX, "output_ch %d unsupported", m.output_ch);
  return 0;
}
inline int field_raw_cols(const scnerf_mlp& m) { return m.use_viewdirs ? 4 : m.output_ch; }

// Forward over P = N*S points.  Either (rays,z) or explicit pts/viewdirs.
inline int field_simt_fwd(const scnerf_mlp& m, const float* rays, int ray_cols, const float* z,
                          const float* pts, const float* viewdirs, int64_t N, int S,
                          const FieldBufs& B, float* raw, void* stream) {
  const int64_t P = N * S;
  const int W = m.W;
  {
    int64_t tot = P * (m.L_pos + 1);
    SCNERF_LAUNCH(pe_points_kernel, (unsigned)cdiv(tot, 256), 256, 0, stream, rays, ray_cols, z, pts,
                  P, S, m.L_pos, B.X5, B.ldx5, m.pts_dim == 4 ? 4 : 3);
  }
  const float* in = B.X5;
  int64_t ld_in = B.ldx5;
  int K = m.input_ch;
  for (int i = 0; i < m.D; ++i) {
    float* out = (i == m.skip) ? B.X5 + m.input_ch : B.H[i];
    int64_t ld_out = (i == m.skip) ? B.ldx5 : W;
    int rc = linear_fwd(in, ld_in, m.pts_w[i], m.pts_b[i], out, ld_out, P, W, K, true, stream);
    if (rc) return rc;
    if (i == m.skip) { in = B.X5; ld_in = B.ldx5; K = m.input_ch + W; }
    else { in = out; ld_in = W; K = W; }
  }
  const int rc_cols = field_raw_cols(m);
  if (m.use_viewdirs) {
    const float* vd = viewdirs ? viewdirs : rays + 8;
    int64_t vstride = viewdirs ? 3 : ray_cols;
    int64_t tot = P * (m.L_dir + 1);
    SCNERF_LAUNCH(pe_dirs_kernel, (unsigned)cdiv(tot, 256), 256, 0, stream, vd, vstride, P, S,
                  m.L_dir, B.F + W, B.ldf);
    SCNERF_LAUNCH(head_fwd_kernel, (unsigned)cdiv(P * 8, 256), 256, 0, stream, in, ld_in, m.alpha_w,
                  m.alpha_b, P, K, 1, raw, (int64_t)rc_cols, 3);
    int rc = linear_fwd(in, ld_in, m.feature_w, m.feature_b, B.F, B.ldf, P, W, K, false, stream);
    if (rc) return rc;
    rc = linear_fwd(B.F, B.ldf, m.views_w, m.views_b, B.HV, W / 2, P, W / 2, W + m.input_ch_views,
                    true, stream);
    if (rc) return rc;
    SCNERF_LAUNCH(head_fwd_kernel, (unsigned)cdiv(P * 8, 256), 256, 0, stream, B.HV, (int64_t)(W / 2),
                  m.rgb_w, m.rgb_b, P, W / 2, 3, raw, (int64_t)rc_cols, 0);
  } else {
    SCNERF_LAUNCH(head_fwd_kernel, (unsigned)cdiv(P * 8, 256), 256, 0, stream, in, ld_in, m.output_w,
                  m.output_b, P, K, m.output_ch, raw, (int64_t)rc_cols, 0);
  }
  return 0;
}

// Backward: g_raw [P,4] -> parameter grads (+=) and d_rays (+=).  B holds the forward activations.
inline int field_simt_bwd(const scnerf_mlp& m, const scnerf_mlp& g, const float* rays, int ray_cols,
                          const float* z, int64_t N, int S, const FieldBufs& B,
                          const FieldGradBufs& G, const float* g_raw, float* d_rays, void* stream,
                          float* g_z = nullptr, const float* pts = nullptr, const float* viewdirs = nullptr,
                          float* g_pts = nullptr, float* g_viewdirs = nullptr) {
  const int64_t P = N * S;
  const int W = m.W, Hh = m.W / 2;
  const int64_t rpb = 2048;
  // final trunk activation h and its layout
  const float* h = (m.D - 1 == m.skip) ? B.X5 : B.H[m.D - 1];
  int64_t ld_h = W;
  int Kh = W;
  float* dz = G.Ga;  // gradient w.r.t. the pre-activation of trunk layer D-1 (masked)
  if (m.use_viewdirs) {
    SCNERF_LAUNCH(head_wgrad_kernel, (unsigned)cdiv(P, rpb), 256, 0, stream, g_raw, (int64_t)4, 0,
                  B.HV, (int64_t)Hh, P, Hh, 3, rpb, g.rgb_w, g.rgb_b);
    SCNERF_LAUNCH(head_dgrad_kernel, (unsigned)cdiv(P * Hh, 256), 256, 0, stream, g_raw, (int64_t)4, 0,
                  m.rgb_w, P, Hh, 3, G.Ghv, (int64_t)Hh, 0, B.HV, (int64_t)Hh);
    int rc = linear_wgrad(G.Ghv, Hh, B.F, B.ldf, g.views_w, W + m.input_ch_views, P, Hh,
                          W + m.input_ch_views, stream);
    if (rc) return rc;
    rc = bias_grad(G.Ghv, Hh, P, Hh, g.views_b, stream);
    if (rc) return rc;
    rc = linear_dgrad(G.Ghv, Hh, m.views_w, W + m.input_ch_views, G.Gf, B.ldf, P, Hh,
                      W + m.input_ch_views, nullptr, 0, 0, false, stream);
    if (rc) return rc;
    rc = linear_wgrad(G.Gf, B.ldf, h, ld_h, g.feature_w, Kh, P, W, Kh, stream);
    if (rc) return rc;
    rc = bias_grad(G.Gf, B.ldf, P, W, g.feature_b, stream);
    if (rc) return rc;
    rc = linear_dgrad(G.Gf, B.ldf, m.feature_w, Kh, dz, W, P, W, Kh, nullptr, 0, 0, false, stream);
    if (rc) return rc;
    SCNERF_LAUNCH(head_wgrad_kernel, (unsigned)cdiv(P, rpb), 256, 0, stream, g_raw, (int64_t)4, 3, h,
                  ld_h, P, Kh, 1, rpb, g.alpha_w, g.alpha_b);
    SCNERF_LAUNCH(head_dgrad_kernel, (unsigned)cdiv(P * Kh, 256), 256, 0, stream, g_raw, (int64_t)4, 3,
                  m.alpha_w, P, Kh, 1, dz, (int64_t)W, 1, h, ld_h);
  } else {
    int no = std::min(m.output_ch, 4);  // channels beyond 3 are never consumed (render.py:316-327)
    SCNERF_LAUNCH(head_wgrad_kernel, (unsigned)cdiv(P, rpb), 256, 0, stream, g_raw, (int64_t)4, 0, h,
                  ld_h, P, Kh, no, rpb, g.output_w, g.output_b);
    SCNERF_LAUNCH(head_dgrad_kernel, (unsigned)cdiv(P * Kh, 256), 256, 0, stream, g_raw, (int64_t)4, 0,
                  m.output_w, P, Kh, no, dz, (int64_t)W, 0, h, ld_h);
  }
  int64_t ld_dz = W;
  for (int i = m.D - 1; i >= 0; --i) {
    // input of layer i
    const float* in;
    int64_t ld_in;
    int K;
    if (i == 0) { in = B.X5; ld_in = B.ldx5; K = m.input_ch; }
    else if (i - 1 == m.skip) { in = B.X5; ld_in = B.ldx5; K = m.input_ch + W; }
    else { in = B.H[i - 1]; ld_in = W; K = W; }
    int rc = linear_wgrad(dz, ld_dz, in, ld_in, g.pts_w[i], K, P, W, K, stream);
    if (rc) return rc;
    rc = bias_grad(dz, ld_dz, P, W, g.pts_b[i], stream);
    if (rc) return rc;
    if (i == 0) {
      // d(PE) from layer 0; adds to the skip branch's share if there is one
      rc = linear_dgrad(dz, ld_dz, m.pts_w[0], K, G.Gx5, B.ldx5, P, W, K, nullptr, 0, 0,
                        m.skip >= 0, stream);
      if (rc) return rc;
    } else if (i - 1 == m.skip) {
      rc = linear_dgrad(dz, ld_dz, m.pts_w[i], K, G.Gx5, B.ldx5, P, W, K, B.X5, B.ldx5, m.input_ch,
                        false, stream);
      if (rc) return rc;
      dz = G.Gx5 + m.input_ch;
      ld_dz = B.ldx5;
    } else {
      float* nxt = (dz == G.Ga) ? G.Gb : G.Ga;
      rc = linear_dgrad(dz, ld_dz, m.pts_w[i], K, nxt, W, P, W, K, B.H[i - 1], W, 0, false, stream);
      if (rc) return rc;
      dz = nxt;
      ld_dz = W;
    }
  }
  if (d_rays) {
    const float* g_ped = m.use_viewdirs && ray_cols > 8 ? G.Gf + W : nullptr;
    int threads = S >= 192 ? 256 : (S > 64 ? 128 : 64);
    SCNERF_LAUNCH(pe_bwd_kernel, (unsigned)N, threads, 0, stream, rays, ray_cols, z, S, m.L_pos,
                  m.L_dir, G.Gx5, B.ldx5, g_ped, B.ldf, d_rays, g_z);
  }
  if (pts && (g_pts || g_viewdirs)) {   // explicit-point mode (NeRF++ background network)
    const float* g_ped = m.use_viewdirs ? G.Gf + W : nullptr;
    SCNERF_LAUNCH(pe_bwd_pts_kernel, (unsigned)N, 128, 0, stream, pts, m.pts_dim == 4 ? 4 : 3, viewdirs, S,
                  m.L_pos, m.L_dir, G.Gx5, B.ldx5, g_ped, B.ldf, g_pts, g_viewdirs);
  }
  return 0;
}

}  // namespace scnerf
