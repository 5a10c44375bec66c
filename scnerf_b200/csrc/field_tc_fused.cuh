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

// real slabs: XS = 4 -> 164, padded to 168; XS = 6 -> 168 exactly (both = 8 x 21 = 2 x 12 x 7)
__host__ __device__ constexpr int n_pad_slabs(int xs) { return xs == 4 ? 4 : 0; }
template <int NSPLIT, int XS = 4>
__host__ __device__ constexpr eng::Plan make_fwd_plan() {
  constexpr int A_XHI = ALay<NSPLIT, XS>::XHI, A_XLO = ALay<NSPLIT, XS>::XLO, A_VHI = ALay<NSPLIT, XS>::VHI,
                A_VLO = ALay<NSPLIT, XS>::VLO, A_ONES = ALay<NSPLIT, XS>::ONES;
  constexpr int N_PAD_SLABS = n_pad_slabs(XS);
  eng::Plan P{};
  int n = 0;
  uint32_t off = 0;
  for (int s = 0; s < NSTAGE; ++s) {
    const StageDef d = stage_def<XS>(s);
    const int nk = d.kx + d.kh + d.kv;
    for (int j = 0; j <= nk; ++j, ++n) {
      eng::SlabDef e{};
      e.n = (uint16_t)d.N; e.acc_col = 0; e.stage = (uint8_t)s; e.img_off = off;
      e.flags = (uint8_t)((j == 0 ? (eng::F_ZERO_ACC | eng::F_STAGE_BEGIN) : 0) | (j == nk ? eng::F_STAGE_END : 0));
      if (j == nk) {                       // bias slab: A = ONES
        e.a_kind = eng::A_SMEM; e.a_off = A_ONES / 16; e.flags = (uint8_t)(e.flags | eng::F_HI_ONLY_A);
      } else if (j < d.kx) {               // PE(pts) columns
        e.a_kind = eng::A_SMEM; e.a_off = (uint16_t)((A_XHI + j * 4096) / 16); e.a_lo_delta = (A_XLO - A_XHI) / 16;
      } else if (j < d.kx + d.kh) {        // hidden state
        e.a_kind = eng::A_TMEM; e.a_off = (uint16_t)((j - d.kx) * 8);
      } else {                             // PE(dir) columns
        e.a_kind = eng::A_SMEM; e.a_off = (uint16_t)((A_VHI + (j - d.kx - d.kh) * 4096) / 16);
        e.a_lo_delta = (A_VLO - A_VHI) / 16;
      }
      if (N_PAD_SLABS > 0 && s == NSTAGE - 1 && j == nk) e.flags = (uint8_t)(e.flags & ~eng::F_STAGE_END);   // padding follows
      P.slab[n] = e;
      off += (uint32_t)d.N * 32u;
    }
  }
  // 164 real slabs: pad to 168 (= 8 x 21) with zero-weight N=16 slabs (A = ONES, B = 0: adds 0 to 16
  // accumulator columns) so the ring size divides the slab count and slot / parity are compile-time
  for (int k = 0; k < N_PAD_SLABS; ++k, ++n) {
    eng::SlabDef e{};
    e.n = 16; e.acc_col = 0; e.stage = (uint8_t)(NSTAGE - 1); e.img_off = off;
    e.a_kind = eng::A_SMEM; e.a_off = A_ONES / 16;
    e.flags = (uint8_t)(eng::F_HI_ONLY_A | (k == N_PAD_SLABS - 1 ? eng::F_STAGE_END : 0));
    P.slab[n] = e;
    off += 16u * 32u;
  }
  P.n_slabs = n; P.n_stages = NSTAGE;
  return P;
}
template <int XS = 4>
inline void build_fwd_plansrc(PlanSrc& S) {
  constexpr int N_PAD_SLABS = n_pad_slabs(XS);
  constexpr int IN_CH = XS == 4 ? 63 : 84;
  int n = 0;
  for (int s = 0; s < NSTAGE; ++s) {
    const StageDef d = stage_def<XS>(s);
    const int nk = d.kx + d.kh + d.kv;
    for (int j = 0; j <= nk; ++j, ++n) {
      SrcDef q{};
      q.wsel = (uint8_t)s; q.valid_n = (uint16_t)d.N;
      if (j == nk) q.kind = 2;
      else if (j < d.kx) { q.col0 = (uint16_t)(16 * j); q.valid_k = (uint16_t)std::max(0, std::min(16, IN_CH - 16 * j)); }
      else if (j < d.kx + d.kh) { q.col0 = (uint16_t)((d.kx ? IN_CH : 0) + 16 * (j - d.kx)); q.valid_k = 16; }
      else { const int jv = j - d.kx - d.kh; q.col0 = (uint16_t)(256 + 16 * jv); q.valid_k = (uint16_t)std::min(16, 27 - 16 * jv); }
      S.s[n] = q;
    }
  }
  for (int k = 0; k < N_PAD_SLABS; ++k, ++n) { SrcDef q{}; q.kind = 3; S.s[n] = q; }   // zeros
}
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

template <int NSPLIT_, int XS_ = 4> struct Cfg {
  static constexpr int NSPLIT = NSPLIT_;
  static constexpr int XS = XS_;
  static constexpr eng::Plan PLAN = make_fwd_plan<NSPLIT_, XS_>();
  static constexpr int GROUP = NSPLIT_ == 1 ? 2 : 1;         // slabs per ring slot
  static constexpr int NSLOT = NSPLIT_ == 1 ? (XS_ == 4 ? 12 : 7) : 8;   // 168 slabs per tile = 84 pairs = 12*7 = 8*21
                                                                         // (XS = 6, bf16: 7 slots, the X slabs need the room)
  static constexpr int SLOT_BYTES = 16384;
  static_assert((PLAN.n_slabs / GROUP) % NSLOT == 0 && PLAN.n_slabs % GROUP == 0, "ring size must divide the slab-group count");
  static constexpr int OFF_RING = 0;
  static constexpr int OFF_A = NSLOT * SLOT_BYTES;
  static constexpr int OFF_C = OFF_A + ALay<NSPLIT_, XS_>::BYTES;
  static constexpr int OFF_OUT = OFF_C + ((C_TOTAL * 4 + 127) / 128) * 128;   // [128][4] head partial sums
  static constexpr int OFF_BAR = OFF_OUT + 128 * 4 * 4;
  static constexpr int SMEM_BYTES = OFF_BAR + (2 * NSLOT + 2) * 8 + 16;
  static_assert(SMEM_BYTES <= 232448, "exceeds the 227 KB shared-memory limit");
};

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
};

__device__ eng::Plan d_plan_fwd;       // runtime copy of the constexpr plan, for the pack kernel
__device__ PlanSrc d_plansrc_fwd;
__device__ eng::Plan d_plan_fwd6;      // ... of the XS = 6 (4-D points) plan
__device__ PlanSrc d_plansrc_fwd6;
template <int NSPLIT, int XS = 4>
__global__ void __launch_bounds__(256) pack_fwd_kernel(PackSrc src, uint8_t* __restrict__ img) {
  const int i = blockIdx.y;
  const eng::Plan& P = XS == 4 ? d_plan_fwd : d_plan_fwd6;
  const PlanSrc& S = XS == 4 ? d_plansrc_fwd : d_plansrc_fwd6;
  if (i < P.n_slabs) pack_slab_impl<NSPLIT>(P.slab[i], S.s[i], src, img);
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

// One epilogue stage for this warp's half of the columns (compile-time stage parameters).
template <int NSPLIT, int S>
__device__ __forceinline__ void epi_stage(const Args& a, const float* cst, uint32_t T_ACC, uint32_t T_AHI,
                                          uint32_t T_ALO, uint32_t lane_base, int half, int row, int tile,
                                          int64_t p, bool valid, uint32_t acc_full_addr, uint32_t a_ready_addr,
                                          float& alpha, float (&rgb)[3], const eng::Ctx& ctx, int tile_iter) {
  constexpr bool SPLIT = NSPLIT == 3;
  constexpr StageDef d = stage_def(S);
  constexpr int nchunk = d.N / 64;           // 32-column chunks owned by this warp
  const int cbase = half * (d.N / 2);
  if constexpr (S >= 1) {
    // training: the previous stage's output (this stage's A operand, in TMEM) goes to its tile image now,
    // under this stage's MMA phase
    if (a.img_out[S - 1].base != nullptr)
      eng::dump_from_tmem<SPLIT, 4>(a.img_out[S - 1], tile, row, T_AHI, T_ALO, lane_base, half * 128, ctx.pol_stream);
  }
  eng::mbar_wait_a(acc_full_addr, (uint32_t)(S & 1));    // 10 stages per tile (even): parity = S & 1
  tc::tc_fence_after();
  if (threadIdx.x == 64) eng::dbg_stamp(ctx, tile_iter, S, NSTAGE, 2);
  // all of this warp's TMEM loads for the stage are issued back to back, then one wait
  uint32_t v[nchunk][32];
#pragma unroll
  for (int cc = 0; cc < nchunk; ++cc) tc::tmem_ld32(T_ACC + lane_base + cbase + cc * 32, v[cc]);
  tc::tmem_ld_wait();
#pragma unroll
  for (int cc = 0; cc < nchunk; ++cc) {
    const int cu = cbase + cc * 32;
    float f[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[cc][j]);
    if (a.dump[S] != nullptr && valid) {
      float* dp = a.dump[S] + p * a.dump_ld[S] + cu;
#pragma unroll
      for (int j = 0; j < 32; ++j) dp[j] = d.relu ? fmaxf(f[j], 0.f) : f[j];
    }
    if constexpr (S == 7) {
#pragma unroll
      for (int j = 0; j < 32; ++j) alpha = fmaf(fmaxf(f[j], 0.f), cst[C_WALPHA + cu + j], alpha);
    }
    if constexpr (S == 9) {
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        const float t = fmaxf(f[j], 0.f);
        rgb[0] = fmaf(t, cst[C_WRGB + cu + j], rgb[0]);
        rgb[1] = fmaf(t, cst[C_WRGB + 128 + cu + j], rgb[1]);
        rgb[2] = fmaf(t, cst[C_WRGB + 256 + cu + j], rgb[2]);
      }
    }
    if (S != 9 || a.img_out[9].base != nullptr) {
      uint32_t hi[16], lo[16];
      eng::split32<SPLIT, d.relu != 0>(f, hi, lo);
      if constexpr (S != 9) {
        tc::tmem_st16(T_AHI + lane_base + (uint32_t)(cu >> 1), hi);
        if constexpr (SPLIT) tc::tmem_st16(T_ALO + lane_base + (uint32_t)(cu >> 1), lo);
      }
      if constexpr (S == 9) {
        if (a.img_out[9].base != nullptr) eng::dump32<SPLIT>(a.img_out[9], tile, row, cu, hi, lo, ctx.pol_stream);
      }
    }
  }
  if constexpr (S < 9) {
    tc::tmem_st_wait();
    tc::tc_fence_before();
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(a_ready_addr) : "memory");
  }
  if constexpr (d.relu != 0) {
    if (a.relu_bits != nullptr) {      // after the hand-off: off the MMA's critical path
      constexpr int L = S == 9 ? 8 : S;
      uint4 mb = make_uint4(0u, 0u, 0u, 0u);
      uint32_t* mw = reinterpret_cast<uint32_t*>(&mb);
#pragma unroll
      for (int cc = 0; cc < nchunk; ++cc) {
        uint32_t bits = 0;
#pragma unroll
        for (int j = 0; j < 32; ++j) bits |= (__uint_as_float(v[cc][j]) > 0.f ? 1u : 0u) << j;
        mw[cc] = bits;
      }
      tc::st_v4_hint(a.relu_bits + ((size_t)(tile * 9 + L) * 2 + half) * 128 + row, mb, ctx.pol_stream);
    }
  }
  if (threadIdx.x == 64) eng::dbg_stamp(ctx, tile_iter, S, NSTAGE, 3);
}
template <int NSPLIT, size_t... Ss>
__device__ __forceinline__ void epi_tile(const Args& a, const float* cst, uint32_t T_ACC, uint32_t T_AHI,
                                         uint32_t T_ALO, uint32_t lane_base, int half, int row, int tile, int64_t p,
                                         bool valid, uint32_t acc_full_addr, uint32_t a_ready_addr, float& alpha,
                                         float (&rgb)[3], const eng::Ctx& ctx, int tile_iter,
                                         std::index_sequence<Ss...>) {
  (epi_stage<NSPLIT, (int)Ss>(a, cst, T_ACC, T_AHI, T_ALO, lane_base, half, row, tile, p, valid, acc_full_addr,
                              a_ready_addr, alpha, rgb, ctx, tile_iter), ...);
}

template <int NSPLIT, int XS = 4>
__global__ void __launch_bounds__(320, 1) field_fused_fwd_kernel(const __grid_constant__ Args a) {
  using C = Cfg<NSPLIT, XS>;
  constexpr bool SPLIT = NSPLIT == 3;
  constexpr int A_XHI = ALay<NSPLIT, XS>::XHI, A_XLO = ALay<NSPLIT, XS>::XLO, A_VHI = ALay<NSPLIT, XS>::VHI,
                A_VLO = ALay<NSPLIT, XS>::VLO, A_ONES = ALay<NSPLIT, XS>::ONES;
  extern __shared__ __align__(128) uint8_t fsm[];
  uint8_t* ringp = fsm + C::OFF_RING;
  uint8_t* areg = fsm + C::OFF_A;
  float* cst = reinterpret_cast<float*>(fsm + C::OFF_C);
  float* out_s = reinterpret_cast<float*>(fsm + C::OFF_OUT);
  uint64_t* full = reinterpret_cast<uint64_t*>(fsm + C::OFF_BAR);
  uint64_t* empty = full + C::NSLOT;
  uint64_t* acc_full = empty + C::NSLOT;
  uint64_t* a_ready = acc_full + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(a_ready + 1);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    for (int i = 0; i < C::NSLOT; ++i) { tc::mbar_init(&full[i], 1); tc::mbar_init(&empty[i], 1); }
    tc::mbar_init(acc_full, 1);
    tc::mbar_init(a_ready, 256);
    tc::fence_mbar_init();
  }
  for (int i = tid; i < C_TOTAL; i += blockDim.x) cst[i] = a.cbuf[i];
  // constant ONES slab: [128 rows x 16 k], k == 0 -> 1.0 (bf16 0x3F80), else 0
  for (int i = tid; i < 4096 / 16; i += blockDim.x) {
    const bool k0chunk = i < 128;          // first k-chunk (k 0..7): 128 rows x 16 B
    *reinterpret_cast<uint4*>(areg + A_ONES + i * 16) = make_uint4(k0chunk ? 0x00003F80u : 0u, 0u, 0u, 0u);
  }
  tc::fence_proxy_async();
  __syncthreads();
  if (warp == 1) tc::tmem_alloc(tmem_slot, 512);
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const uint32_t T_ACC = tmem, T_AHI = tmem + 256, T_ALO = tmem + 384;
  eng::Ctx ctx;
  ctx.ring_addr = tc::smem_u32(ringp); ctx.full_addr = tc::smem_u32(full); ctx.empty_addr = tc::smem_u32(empty);
  ctx.acc_full_addr = tc::smem_u32(acc_full); ctx.a_ready_addr = tc::smem_u32(a_ready);
  ctx.tmem_acc = T_ACC; ctx.tmem_ahi = T_AHI; ctx.tmem_alo = T_ALO; ctx.smem_a = tc::smem_u32(areg);
  ctx.dbg = a.dbg; ctx.dbg_tiles = a.dbg_tiles;
  ctx.pol_keep = tc::policy_evict_last(); ctx.pol_stream = tc::policy_evict_first();

  if (warp == 0) {
    if (lane == 0) eng::producer_loop<C>(ctx, a.wimg, a.num_tiles);
  } else if (warp == 1) {
    if (lane == 0) eng::mma_loop<C>(ctx, a.num_tiles);
  } else {
    // ===================== epilogue: 8 warps, 2 per TMEM lane quadrant =============================
    const int quad = warp & 3, half = (warp - 2) >> 2;
    const int row = quad * 32 + lane;
    const uint32_t lane_base = (uint32_t)(quad * 32) << 16;
    int tile_iter = 0;
    for (int tile = blockIdx.x; tile < a.num_tiles; tile += gridDim.x, ++tile_iter) {
      const int64_t p = (int64_t)tile * TILE_M + row;
      const bool valid = p < a.P;
      // ---- positional encodings -> smem A slabs (half 0: X chunks 0-3; half 1: X 4-7 and V 0-3)
      if constexpr (XS == 6) {
        // NeRF++ background: explicit 4-D points (x, y, z, 1/r), 84-channel encoding in 6 slabs
        // (half 0: X columns 0..31 + PE(dir); half 1: X columns 32..95)
        float x[4] = {0.f, 0.f, 0.f, 0.f}, vd[3] = {0.f, 0.f, 0.f};
        if (valid) {
          const int64_t r = p / a.S;
          const float4 q = *reinterpret_cast<const float4*>(a.pts + p * 4);
          x[0] = q.x; x[1] = q.y; x[2] = q.z; x[3] = q.w;
#pragma unroll
          for (int c = 0; c < 3; ++c) vd[c] = a.viewdirs[r * 3 + c];
        }
        float* dpe = a.dump_pe ? a.dump_pe + p * a.dump_pe_ld : nullptr;
        float* dped = a.dump_ped ? a.dump_ped + p * a.dump_ped_ld : nullptr;
        float e[32];
        if (half == 0) {
          pe_fill32<10, 0, 4>(x, valid, e);
          pe_store32<SPLIT>(e, 0, 84, areg + A_XHI, areg + A_XLO, row, a.img_x, tile, dpe, valid, ctx.pol_stream);
          pe_fill32<4, 0>(vd, valid, e);
          pe_store32<SPLIT>(e, 0, 27, areg + A_VHI, areg + A_VLO, row, a.img_v, tile, dped, valid, ctx.pol_stream);
          *reinterpret_cast<float4*>(out_s + row * 4) = make_float4(0.f, 0.f, 0.f, 0.f);
        } else {
          pe_fill32<10, 32, 4>(x, valid, e);
          pe_store32<SPLIT>(e, 32, 84, areg + A_XHI, areg + A_XLO, row, a.img_x, tile, dpe, valid, ctx.pol_stream);
          pe_fill32<10, 64, 4>(x, valid, e);
          pe_store32<SPLIT>(e, 64, 84, areg + A_XHI, areg + A_XLO, row, a.img_x, tile, dpe, valid, ctx.pol_stream);
        }
        tc::fence_proxy_async();
        tc::mbar_arrive(a_ready);
      } else {
        float x[3] = {0.f, 0.f, 0.f}, vd[3] = {0.f, 0.f, 0.f};
        if (valid) {
          const int64_t r = p / a.S;
          if (a.rays) {
            const float* ry = a.rays + r * a.ray_cols;
            const float zz = a.z[p];
#pragma unroll
            for (int c = 0; c < 3; ++c) {
              x[c] = __fadd_rn(ry[c], __fmul_rn(ry[3 + c], zz));   // render.py:259
              vd[c] = ry[8 + c];
            }
          } else {
#pragma unroll
            for (int c = 0; c < 3; ++c) { x[c] = a.pts[p * 3 + c]; vd[c] = a.viewdirs[r * 3 + c]; }
          }
        }
        float* dpe = a.dump_pe ? a.dump_pe + p * a.dump_pe_ld : nullptr;
        float* dped = a.dump_ped ? a.dump_ped + p * a.dump_ped_ld : nullptr;
        float e[32];
        if (half == 0) {          // X columns 0..31 (15 sincosf) + PE(dir) (12 sincosf)
          pe_fill32<10, 0>(x, valid, e);
          pe_store32<SPLIT>(e, 0, 63, areg + A_XHI, areg + A_XLO, row, a.img_x, tile, dpe, valid, ctx.pol_stream);
          pe_fill32<4, 0>(vd, valid, e);
          pe_store32<SPLIT>(e, 0, 27, areg + A_VHI, areg + A_VLO, row, a.img_v, tile, dped, valid, ctx.pol_stream);
          *reinterpret_cast<float4*>(out_s + row * 4) = make_float4(0.f, 0.f, 0.f, 0.f);
        } else {                  // X columns 32..63 (16 sincosf)
          pe_fill32<10, 32>(x, valid, e);
          pe_store32<SPLIT>(e, 32, 63, areg + A_XHI, areg + A_XLO, row, a.img_x, tile, dpe, valid, ctx.pol_stream);
        }
        tc::fence_proxy_async();
        tc::mbar_arrive(a_ready);
      }
      float alpha = 0.f, rgb[3] = {0.f, 0.f, 0.f};
      epi_tile<NSPLIT>(a, cst, T_ACC, T_AHI, T_ALO, lane_base, half, row, tile, p, valid, ctx.acc_full_addr,
                       ctx.a_ready_addr, alpha, rgb, ctx, tile_iter, std::make_index_sequence<NSTAGE>{});
      // combine the two column-halves of each row: both add into smem, half 0 finishes
      atomicAdd(out_s + row * 4 + 0, rgb[0]);
      atomicAdd(out_s + row * 4 + 1, rgb[1]);
      atomicAdd(out_s + row * 4 + 2, rgb[2]);
      atomicAdd(out_s + row * 4 + 3, alpha);
      tc::tc_fence_before();
      asm volatile("bar.sync 1, 256;" ::: "memory");
      if (half == 0 && valid) {
        const float4 o = *reinterpret_cast<const float4*>(out_s + row * 4);
        *reinterpret_cast<float4*>(a.raw + p * 4) =
            make_float4(o.x + cst[C_SCAL + 1], o.y + cst[C_SCAL + 2], o.z + cst[C_SCAL + 3], o.w + cst[C_SCAL]);
      }
      asm volatile("bar.sync 1, 256;" ::: "memory");   // out_s is re-zeroed by the next tile's prologue
    }
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 1) tc::tmem_dealloc(tmem, 512);
}

}  // namespace fused
}  // namespace scnerf
