// Fused field forward on tcgen05 (sm_100a): positional encoding -> 8x256 trunk (skip at 4) ->
// alpha head + feature layer -> view branch -> rgb head, for one 128-sample tile at a time, with
// NO intermediate activation ever leaving the SM:
//
//   HBM in : 44 B/ray-sample of geometry (rays, z; L2-resident) ; HBM out: raw[4] = 16 B/sample
//   weights: bf16 (hi[,lo]) K=16 slabs streamed L2 -> smem by 1-D bulk TMA through an mbarrier ring
//   A operand (activations): TMEM (TS-mode MMA) for the 256-wide hidden state, smem for the
//                            PE(pts) (K=64) and PE(dir) (K=32) slabs and the constant ONES slab
//   accumulator: TMEM, 128 lanes x 256 fp32 columns; biases are folded into the GEMM as one extra
//                K=16 slab (A = ONES, B = [bias | 0]) so the epilogue has no bias traffic
//
// Warp roles (320 threads):  warp 0 = TMA producer, warp 1 = MMA issuer (+TMEM alloc),
// warps 2..9 = epilogue, two per TMEM lane quadrant (warp & 3), each owning half of the output
// columns: ReLU + bf16 (hi/lo) split with cvt.rn.relu.bf16x2, tcgen05.st back into TMEM as the next
// layer's A operand; alpha / rgb heads are fp32 dot products in the epilogue registers.
//
// NSPLIT = 1: single-pass bf16 (fp32 accumulate).   NSPLIT = 3: split-bf16, A_hi*B_hi + A_lo*B_hi +
// A_hi*B_lo — ~16 mantissa bits per operand, which is what the 1e-4 parity gate needs.
//
// Training mode additionally writes every layer input as bf16 (hi[,lo]) TILE IMAGES in the operand
// layout of the wgrad kernel (tc_engine.cuh: ImgDump), 16-byte coalesced stores.
//
// Algorithmic work: 593,408 MAC/sample; tensor pipe executes
// (593,920 + 10 bias slabs) MAC/sample x NSPLIT.   Reference: NeRF/run_nerf_helpers.py:24-72,105-128,
// NeRF/create_nerf.py:18-32.
#pragma once
#include "common.cuh"
#include "tc_prims.cuh"
#include "tc_engine.cuh"
#include "composite.cuh"

namespace scnerf {
namespace fused {

using eng::TILE_M;
constexpr int NSTAGE = 10;
// per stage: output width N, k16 slabs taken from X (PE pts, smem), H (hidden, TMEM), V (PE dir, smem)
struct StageDef { int N, kx, kh, kv, relu; };
// XS = k16 slabs of PE(pts): 4 (63 channels, 3-D points) or 6 (84 channels, NeRF++ background (x,y,z,1/r))
template <int XS = 4>
__host__ __device__ constexpr StageDef stage_def(int s) {
  return s == 0 ? StageDef{256, XS, 0, 0, 1}
       : s == 5 ? StageDef{256, XS, 16, 0, 1}
       : s == 8 ? StageDef{256, 0, 16, 0, 0}      // feature_linear: no activation
       : s == 9 ? StageDef{128, 0, 16, 2, 1}      // views_linears[0]
                : StageDef{256, 0, 16, 0, 1};
}

// smem A area (byte offsets from its base); the lo images exist only in the split-bf16 build
template <int NSPLIT, int XS = 4> struct ALay {
  static constexpr int XHI = 0;
  static constexpr int XLO = XS * 4096;
  static constexpr int VHI = (NSPLIT == 3 ? 2 : 1) * XS * 4096;
  static constexpr int VLO = VHI + 8192;
  static constexpr int ONES = VHI + (NSPLIT == 3 ? 16384 : 8192);
  static constexpr int BYTES = ONES + 4096;
};

// source of each slab for the pack kernel (parallel to the plan)
struct SrcDef {
  uint8_t wsel;      // index into PackSrc.w / .b
  uint8_t kind;      // 0: B[n][k] = W[n][col0+k]   1: B[n][k] = W[row0+k][col0+n]   2: B[n][0] = b[n]
  uint16_t row0, col0, valid_k, valid_n;
  uint16_t pad;      // kinds 0 and 2: first output row (N-half plans), 0 otherwise
};
struct PlanSrc { SrcDef s[eng::MAX_SLABS]; };
struct PackSrc {
  const float* w[12]; int ld[12];
  const float* b[12];
  const float *alpha_w, *alpha_b, *rgb_w, *rgb_b;
};

inline size_t plan_image_bytes(const eng::Plan& P, int nsplit) {
  size_t b = 0;
  for (int i = 0; i < P.n_slabs; ++i) b += (size_t)P.slab[i].n * 32u * (nsplit == 3 ? 2 : 1);
  return b;
}

// one thread per 16-byte chunk of a slab: blockIdx.y = slab, thread -> (row n, k-chunk)
template <int NSPLIT>
__device__ __forceinline__ void pack_slab_impl(const eng::SlabDef& d, const SrcDef& q, const PackSrc& src,
                                               uint8_t* __restrict__ img) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= d.n * 2) return;
  const int chunk = t / d.n, row = t % d.n;
  float v[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    const int k = chunk * 8 + e;
    float x = 0.f;
    if (q.kind == 0) { if (k < q.valid_k && row < q.valid_n) x = src.w[q.wsel][(int64_t)(q.pad + row) * src.ld[q.wsel] + q.col0 + k]; }
    else if (q.kind == 1) { if (k < q.valid_k && row < q.valid_n) x = src.w[q.wsel][(int64_t)(q.row0 + k) * src.ld[q.wsel] + q.col0 + row]; }
    else if (q.kind == 2) { if (k == 0 && row < q.valid_n) x = src.b[q.wsel][q.pad + row]; }
    v[e] = x;
  }
  uint32_t h[4], l[4];
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    h[e] = eng::cvt_bf16x2(v[2 * e], v[2 * e + 1]);
    l[e] = eng::cvt_bf16x2(v[2 * e] - eng::bf16lo_f(h[e]), v[2 * e + 1] - eng::bf16hi_f(h[e]));
  }
  const size_t sb = (size_t)d.n * 32;
  const size_t in_slab = (size_t)chunk * d.n * 16 + (row >> 3) * 128 + (row & 7) * 16;
  uint8_t* dst = img + (size_t)d.img_off * (NSPLIT == 3 ? 2 : 1);
  *reinterpret_cast<uint4*>(dst + in_slab) = make_uint4(h[0], h[1], h[2], h[3]);
  if (NSPLIT == 3) *reinterpret_cast<uint4*>(dst + sb + in_slab) = make_uint4(l[0], l[1], l[2], l[3]);
}

// constants block (fp32) kept in shared memory
constexpr int C_WALPHA = 0;               // 256
constexpr int C_WRGB = 256;               // 3 x 128
constexpr int C_SCAL = 640;               // b_alpha, b_rgb[3]
constexpr int C_TOTAL = 648;
__global__ void pack_consts_kernel(PackSrc src, float* __restrict__ cbuf) {
  int g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= C_TOTAL) return;
  float v = 0.f;
  if (g < C_WRGB) v = src.alpha_w[g];
  else if (g < C_SCAL) v = src.rgb_w[g - C_WRGB];
  else if (g == C_SCAL) v = src.alpha_b[0];
  else if (g < C_SCAL + 4) v = src.rgb_b[g - C_SCAL - 1];
  cbuf[g] = v;
}

struct Args {
  const float* rays; int ray_cols;     // [N, ray_cols] (o, d, near, far, viewdirs) or NULL
  const float* z;                      // [N, S] depths (with rays)
  const float* pts;                    // [N, S, 3] explicit points (rays == NULL)
  const float* viewdirs;               // [N, 3]    explicit directions (rays == NULL)
  int64_t P; int S;
  const uint8_t* wimg;                 // packed weight image
  const float* cbuf;                   // packed head constants, C_TOTAL floats
  float* raw;                          // [P, 4]
  int num_tiles;
  // training (tensor-core backward): bf16 tile images of every layer input
  eng::ImgDump img_x, img_v, img_out[NSTAGE];
  // ReLU masks for the tensor-core dgrad: 1 bit per activation, [tile][9 layers][2 halves][128 rows] x 16 B
  // (layers 0..7 = trunk, 8 = view layer; a warp's 32 rows store 512 contiguous bytes)
  uint4* relu_bits;
  // training (fp32 CUDA-core backward): fp32 row-major copies
  float* dump[NSTAGE]; int dump_ld[NSTAGE];
  float* dump_pe; int dump_pe_ld;
  float* dump_ped; int dump_ped_ld;
  long long* dbg; int dbg_tiles;       // optional in-kernel timeline (see eng::Ctx)
  // fused alpha-composite (raw2outputs, NeRF/render.py:302-355) in the pipelined forward's epilogue: a CTA takes
  // CONSECUTIVE tiles in groups of comp_G tiles that hold whole rays (comp_G * 128 is a multiple of S), keeps the
  // group's raw values in shared memory and composites its rays there — `raw` (may be NULL) never has to exist in HBM
  int comp_on, comp_G;
  CompositeArgs comp;
};
// which tiles a CTA works on: strided (default) or, with the fused composite, a contiguous run of whole ray groups
__device__ __forceinline__ void cta_tiles(const Args& a, int& first, int& count, int& stride) {
  if (a.comp_on) {
    const int NG = (a.num_tiles + a.comp_G - 1) / a.comp_G;
    const int g0 = (int)((int64_t)blockIdx.x * NG / gridDim.x), g1 = (int)((int64_t)(blockIdx.x + 1) * NG / gridDim.x);
    first = g0 * a.comp_G;
    const int end = g1 * a.comp_G < a.num_tiles ? g1 * a.comp_G : a.num_tiles;
    count = end > first ? end - first : 0;
    stride = 1;
  } else {
    first = blockIdx.x; stride = gridDim.x;
    count = (int)blockIdx.x < a.num_tiles ? (a.num_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
  }
}


// sin/cos with 3-term Cody-Waite reduction by pi/2 and the usual degree-7/8 minimax kernels: <= 7e-8
// absolute error for |a| < 3000 (checked against float64; arguments here are <= 2^9 * |coordinate|).
// Straight-line FMA code (no slow-path branch, no calls), so independent evaluations interleave — the
// libm sincosf version cost ~550 latency-bound cycles per call and 20 % of the tile.
__device__ __forceinline__ void sincos_cw(float a, float& sv, float& cv) {
  const int qi = __float2int_rn(a * 0.636619772f);
  const float q = (float)qi;
  float r = fmaf(q, -1.57079601e+00f, a);
  r = fmaf(q, -3.13916473e-07f, r);
  r = fmaf(q, -5.39030253e-15f, r);
  const float s2 = r * r;
  float ps = fmaf(2.86567956e-6f, s2, -1.98559923e-4f);
  ps = fmaf(ps, s2, 8.33338592e-3f);
  ps = fmaf(ps, s2, -1.66666672e-1f);
  const float sn = fmaf(ps, r * s2, r);
  float pc = fmaf(2.44677067e-5f, s2, -1.38877297e-3f);
  pc = fmaf(pc, s2, 4.16666567e-2f);
  pc = fmaf(pc, s2, -0.5f);
  const float cs = fmaf(pc, s2, 1.0f);
  float so = (qi & 1) ? cs : sn, co = (qi & 1) ? sn : cs;
  sv = (qi & 2) ? -so : so;
  cv = ((qi + 1) & 2) ? -co : co;
}

// PE columns [LO, LO+32) of a 3-vector with L frequencies ([x, sin(2^0 x), cos(2^0 x), ...], zero padded):
// one sin/cos evaluation per (frequency, component) pair that touches the range; indices are compile-time.
template <int L, int LO, int DIM = 3>
__device__ __forceinline__ void pe_fill32(const float (&x)[DIM], bool valid, float (&e)[32]) {
#pragma unroll
  for (int i = 0; i < 32; ++i) e[i] = (valid && LO + i < DIM) ? x[(LO + i) % DIM] : 0.f;
  if (!valid) return;
#pragma unroll
  for (int f = 0; f < L; ++f) {
#pragma unroll
    for (int c = 0; c < DIM; ++c) {
      const int is = DIM + 2 * DIM * f + c, ic = is + DIM;
      const bool need_s = is >= LO && is < LO + 32, need_c = ic >= LO && ic < LO + 32;
      if (need_s || need_c) {
        float sv, cv;
        sincos_cw(x[c] * (float)(1 << f), sv, cv);
        if (need_s) e[is - LO] = sv;
        if (need_c) e[ic - LO] = cv;
      }
    }
  }
}
// store 32 PE columns [LO, LO+32) of row `row`: canonical K-major smem image (+ tile image, + fp32 dump)
template <bool SPLIT>
__device__ __forceinline__ void pe_store32(const float (&e)[32], int lo_col, int ncols_valid, uint8_t* hi_img,
                                           uint8_t* lo_img, int row, const eng::ImgDump& img, int tile,
                                           float* dump, bool valid, uint64_t pol) {
  uint32_t hi[16], lo[16];
  eng::split32<SPLIT, false>(e, hi, lo);
#pragma unroll
  for (int g = 0; g < 4; ++g) {
    const uint32_t off = tc::canon_off(row, lo_col + 8 * g, TILE_M);
    *reinterpret_cast<uint4*>(hi_img + off) = make_uint4(hi[4 * g], hi[4 * g + 1], hi[4 * g + 2], hi[4 * g + 3]);
    if (SPLIT) *reinterpret_cast<uint4*>(lo_img + off) = make_uint4(lo[4 * g], lo[4 * g + 1], lo[4 * g + 2], lo[4 * g + 3]);
  }
  if (img.base) eng::dump32<SPLIT>(img, tile, row, lo_col, hi, lo, pol);
  if (dump && valid) {
#pragma unroll
    for (int i = 0; i < 32; ++i)
      if (lo_col + i < ncols_valid) dump[lo_col + i] = e[i];
  }
}

}  // namespace fused
}  // namespace scnerf
