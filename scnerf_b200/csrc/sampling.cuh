// Per-ray samplers: stratified depths, inverse-CDF importance sampling fused with the sort-merge,
// and the searchsorted primitive.  HBM-bound: ~1.0 KB/ray for sample_pdf, ~1.5 KB/ray for the
// merge — fused here so the [N,Nf] samples never round-trip before the sort.
#pragma once
#include "common.cuh"

namespace scnerf {

// torch.linspace(0,1,steps)[i] in fp32 (ATen: symmetric fill from both ends)
__device__ __forceinline__ float linspace01(int i, int steps) {
  if (steps == 1) return 0.f;
  float step = 1.f / (float)(steps - 1);
  return (i < steps / 2) ? __fmul_rn(step, (float)i) : __fadd_rn(1.f, -__fmul_rn(step, (float)(steps - i - 1)));
}

__device__ __forceinline__ float depth_at(float near, float far, int s, int S, int lindisp) {
  float t = linspace01(s, S);
  if (lindisp) return 1.f / (1.f / near * (1.f - t) + 1.f / far * t);
  return __fadd_rn(__fmul_rn(near, 1.f - t), __fmul_rn(far, t));
}

// NeRF/render.py:235-257.  One thread per (ray, sample).
__global__ void __launch_bounds__(256) stratified_kernel(const float* __restrict__ rays, int ray_cols,
                                                         int64_t N, int S, int lindisp, int perturb,
                                                         const float* __restrict__ t_rand,
                                                         uint64_t seed, float* __restrict__ z) {
  int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= N * S) return;
  int64_t r = g / S;
  int s = (int)(g % S);
  float near = rays[r * ray_cols + 6], far = rays[r * ray_cols + 7];
  float zc = depth_at(near, far, s, S, lindisp);
  if (perturb) {
    float zl = s > 0 ? depth_at(near, far, s - 1, S, lindisp) : zc;
    float zu = s < S - 1 ? depth_at(near, far, s + 1, S, lindisp) : zc;
    float lower = s > 0 ? .5f * (zc + zl) : zc;
    float upper = s < S - 1 ? .5f * (zu + zc) : zc;
    float tr = t_rand ? t_rand[g] : Philox::uniform(seed, RNG_T_RAND, (uint64_t)g);
    zc = __fadd_rn(lower, __fmul_rn(upper - lower, tr));
  }
  z[g] = zc;
}

__device__ __forceinline__ int upper_bound(const float* a, int n, float v) {  // #elements <= v
  int lo = 0, hi = n;
  while (lo < hi) {
    int mid = (lo + hi) >> 1;
    if (a[mid] <= v) lo = mid + 1; else hi = mid;
  }
  return lo;
}
__device__ __forceinline__ int lower_bound(const float* a, int n, float v) {  // #elements < v
  int lo = 0, hi = n;
  while (lo < hi) {
    int mid = (lo + hi) >> 1;
    if (a[mid] < v) lo = mid + 1; else hi = mid;
  }
  return lo;
}

// torch.searchsorted / torchsearchsorted semantics (searchsorted_cuda_kernel.cu:83-107)
__global__ void __launch_bounds__(256) searchsorted_kernel(const float* __restrict__ a,
                                                           const float* __restrict__ v,
                                                           int64_t* __restrict__ out, int64_t nrow,
                                                           int64_t nrow_a, int64_t nrow_v,
                                                           int64_t ncol_a, int64_t ncol_v, int right) {
  int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= nrow * ncol_v) return;
  int64_t r = g / ncol_v, c = g % ncol_v;
  const float* row = a + (nrow_a == 1 ? 0 : r) * ncol_a;
  float val = v[(nrow_v == 1 ? 0 : r) * ncol_v + c];
  out[g] = right ? upper_bound(row, (int)ncol_a, val) : lower_bound(row, (int)ncol_a, val);
}

// sample_pdf (NeRF/render.py:417-460) + `.detach(); sort(cat([z, z_samples]))` (:274-276) +
// z_std (:294).  One CTA (128 threads) per ray.
//   bins   : [N,M] explicit bin edges, or NULL -> mids of z_c (M = Nc-1)
//   w      : weights row pointer base, row stride w_stride, first used element w_off, count M-1
//   u      : [N,Nf] or NULL (det -> linspace(0,1,Nf); else Philox)
//   merged : [N, Nc+Nf] sorted union with z_c, or NULL
struct SamplePdfArgs {
  const float* bins; const float* z_c; int Nc;
  const float* w; int64_t w_stride; int w_off; int M;
  const float* u; int det; uint64_t seed; int Nf;
  float* samples; int64_t* inds; float* merged; float* z_std;
  int64_t N; int sort_n;  // power of two >= Nc+Nf
};

__global__ void __launch_bounds__(128) sample_pdf_kernel(SamplePdfArgs a) {
  extern __shared__ float sm[];
  float* cdf = sm;                 // [M]
  float* bins = sm + a.M;          // [M]
  float* srt = sm + 2 * a.M;       // [sort_n]
  __shared__ float red[8];
  const int64_t r = blockIdx.x;
  const int tid = threadIdx.x;
  const int M = a.M;
  // bins
  for (int i = tid; i < M; i += blockDim.x)
    bins[i] = a.bins ? a.bins[r * M + i]
                     : .5f * (a.z_c[r * a.Nc + i + 1] + a.z_c[r * a.Nc + i]);
  // pdf/cdf: sequential fp32 sum and cumulative sum in index order (62 terms; matches
  // torch.cumsum's order on CPU; torch.sum's internal order is not reproducible anyway)
  const float* w = a.w + r * a.w_stride + a.w_off;
  for (int i = tid; i < M - 1; i += blockDim.x) cdf[i + 1] = w[i] + 1e-5f;
  __syncthreads();
  if (tid == 0) {
    float tot = 0.f;
    for (int i = 1; i < M; ++i) tot += cdf[i];
    float run = 0.f;
    cdf[0] = 0.f;
    for (int i = 1; i < M; ++i) { run += cdf[i] / tot; cdf[i] = run; }
  }
  __syncthreads();
  float lsum = 0.f;
  for (int j = tid; j < a.Nf; j += blockDim.x) {
    float u = a.u ? a.u[r * a.Nf + j]
                  : (a.det ? linspace01(j, a.Nf)
                           : Philox::uniform(a.seed, RNG_U, (uint64_t)(r * a.Nf + j)));
    int ind = upper_bound(cdf, M, u);
    int below = max(ind - 1, 0), above = min(ind, M - 1);
    float c0 = cdf[below], c1 = cdf[above], b0 = bins[below], b1 = bins[above];
    float denom = c1 - c0;
    if (denom < 1e-5f) denom = 1.f;
    float t = (u - c0) / denom;
    float s = __fadd_rn(b0, __fmul_rn(t, b1 - b0));
    if (a.samples) a.samples[r * a.Nf + j] = s;
    if (a.inds) a.inds[r * a.Nf + j] = ind;
    if (a.merged) srt[a.Nc + j] = s;
    lsum += s;
  }
  if (a.z_std) {  // torch.std(unbiased=False): two-pass
    float v = warp_sum(lsum);
    if ((tid & 31) == 0) red[tid >> 5] = v;
    __syncthreads();
    float mean = (red[0] + red[1] + red[2] + red[3]) / (float)a.Nf;
    __syncthreads();
    float lsq = 0.f;
    for (int j = tid; j < a.Nf; j += blockDim.x) {
      float s = a.merged ? srt[a.Nc + j] : a.samples[r * a.Nf + j];
      lsq += (s - mean) * (s - mean);
    }
    v = warp_sum(lsq);
    if ((tid & 31) == 0) red[4 + (tid >> 5)] = v;
    __syncthreads();
    if (tid == 0) a.z_std[r] = sqrtf((red[4] + red[5] + red[6] + red[7]) / (float)a.Nf);
  }
  if (!a.merged) return;
  const int tot = a.Nc + a.Nf;
  for (int i = tid; i < a.Nc; i += blockDim.x) srt[i] = a.z_c[r * a.Nc + i];
  for (int i = tot + tid; i < a.sort_n; i += blockDim.x) srt[i] = __int_as_float(0x7f800000);
  __syncthreads();
  // bitonic sort of sort_n keys in shared memory
  for (int k = 2; k <= a.sort_n; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = tid; i < a.sort_n; i += blockDim.x) {
        int p = i ^ j;
        if (p > i) {
          float x = srt[i], y = srt[p];
          bool up = (i & k) == 0;
          if ((x > y) == up) { srt[i] = y; srt[p] = x; }
        }
      }
      __syncthreads();
    }
  }
  for (int i = tid; i < tot; i += blockDim.x) a.merged[r * tot + i] = srt[i];
}

// Stand-alone sort(cat([a,b])) — render.py:276
__global__ void __launch_bounds__(128) sort_merge_kernel(const float* __restrict__ a,
                                                         const float* __restrict__ b, int Na, int Nb,
                                                         int sort_n, float* __restrict__ out) {
  extern __shared__ float srt[];
  const int64_t r = blockIdx.x;
  const int tid = threadIdx.x, tot = Na + Nb;
  for (int i = tid; i < Na; i += blockDim.x) srt[i] = a[r * Na + i];
  for (int i = tid; i < Nb; i += blockDim.x) srt[Na + i] = b[r * Nb + i];
  for (int i = tot + tid; i < sort_n; i += blockDim.x) srt[i] = __int_as_float(0x7f800000);
  __syncthreads();
  for (int k = 2; k <= sort_n; k <<= 1)
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = tid; i < sort_n; i += blockDim.x) {
        int p = i ^ j;
        if (p > i) {
          float x = srt[i], y = srt[p];
          bool up = (i & k) == 0;
          if ((x > y) == up) { srt[i] = y; srt[p] = x; }
        }
      }
      __syncthreads();
    }
  for (int i = tid; i < tot; i += blockDim.x) out[r * tot + i] = srt[i];
}

inline int next_pow2(int v) { int p = 1; while (p < v) p <<= 1; return p; }

}  // namespace scnerf
