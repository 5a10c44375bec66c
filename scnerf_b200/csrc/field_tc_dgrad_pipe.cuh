// Fused field backward (data-gradient chain), N-half PIPELINED variant (sm_100a, tcgen05).
//
// XN = 64: 3-D points formed from (rays, z), 63-channel encoding; XN = 96: explicit 4-D points (NeRF++ background network),
// 84-channel encoding.  The schedule is the one of field_tc_fwd_pipe.cuh: every transposed-weight GEMM runs as two 128-column N-halves into two
// accumulators, the epilogue hands the next stage's A operand (dZ) over in 64-column quarters, the first A
// half is double-buffered (P0/P1), the second is single (Q), and in split-bf16 the lo operand lives in shared
// memory.  Ten stages per tile:
//
//   T0  [g_feat | g_V]  = dZ_v (K=128) * W_views        (g_V: 32 columns, extra accumulator ACCX, rides the h1 pass)
//   T1  g_h7 = g_feat * W_feature (+ alpha head)        T2, T3  g_h6, g_h5
//   T4  g_h4 = dZ5 * W5[:, 63:]   (+ the 64-wide skip share of d(PE) = dZ5 * W5[:, :63] in ACCX, rides the h1 pass)
//   T5..T8  g_h3 .. g_h0                                T9  d(PE) = dZ0 * W0  (64 columns, one pass)
//
//   TMEM: acc0 [0,128)  acc1 [128,256)  A_hi: P0 [256,320)  P1 [320,384)  Q [384,448)  ACCX [448,512)
//
// XN = 96: the skip share of d(PE) is 96 columns wide and ACCX has 64, so T4 runs it as a SIDE PASS between its two halves:
//   h0 (acc0) | side: dZ5 * W5[:, :84] -> acc1[0,96) , commit side[0] | h1 (acc1), first slab waits side[1]
// and the epilogue parks the side result in shared memory (side[0] -> 48 KB gx_s -> side[1]) BEFORE it converts h0, so the
// h1 pass starts ~0.5 K cycles after the side pass ends and T5's operand quarters are still ready when h1 is done.
//
// Graph = NeRF.forward's autograd graph (NeRF/run_nerf_helpers.py:105-128) + Embedder (:24-72).
#pragma once
#include "common.cuh"
#include "tc_prims.cuh"
#include "tc_engine.cuh"
#include "field_tc_fused.cuh"
#include "field_tc_fwd_pipe.cuh"
#include "field_tc_dgrad.cuh"

namespace scnerf {
namespace dpipe {

using eng::TILE_M;
using fused::PlanSrc;
using fused::SrcDef;
using dgrad::Args;
constexpr int NSTAGE = 10;
constexpr int ACCX_COL = 448;

// f(stage, h, N, acc_col, j, first_in_pass, last_in_pass, wsel, col0, valid_n, nk);  h = 2: side pass (XN = 96, T4)
template <int XN = 64, class F>
__host__ __device__ constexpr void for_each_slab(F&& f) {
  constexpr int IN_CH = XN == 64 ? 63 : 84;
  // T0: views layer, K = 128
  for (int j = 0; j < 8; ++j) f(0, 0, 128, 0, j, j == 0, j == 7, 9, 0, 128, 8);
  for (int j = 0; j < 8; ++j) f(0, 1, 128, 128, j, j == 0, false, 9, 128, 128, 8);
  for (int j = 0; j < 8; ++j) f(0, 1, 32, ACCX_COL, j, j == 0, j == 7, 9, 256, 27, 8);
  const int wsel[10] = {9, 8, 7, 6, 5, 4, 3, 2, 1, 0};
  for (int t = 1; t <= 8; ++t) {
    const int c0 = t == 4 ? IN_CH : 0;
    for (int j = 0; j < 16; ++j) f(t, 0, 128, 0, j, j == 0, j == 15, wsel[t], c0, 128, 16);
    if (t == 4 && XN == 96)
      for (int j = 0; j < 16; ++j) f(t, 2, XN, 128, j, j == 0, j == 15, 5, 0, IN_CH, 16);
    for (int j = 0; j < 16; ++j)
      f(t, 1, 128, 128, j, j == 0, (t == 4 && XN == 64) ? false : j == 15, wsel[t], c0 + 128, 128, 16);
    if (t == 4 && XN == 64)
      for (int j = 0; j < 16; ++j) f(t, 1, XN, ACCX_COL, j, j == 0, j == 15, 5, 0, IN_CH, 16);
  }
  for (int j = 0; j < 16; ++j) f(9, 0, XN, 0, j, j == 0, j == 15, 0, 0, IN_CH, 16);
}
template <int XN>
struct PlanFiller {
  eng::Plan P{};
  int n = 0;
  uint32_t off = 0;
  __host__ __device__ constexpr void operator()(int s, int h, int N, int acc_col, int j, bool first, bool last,
                                                int, int, int, int nk) {
    eng::SlabDef e{};
    e.n = (uint16_t)N; e.acc_col = (uint16_t)acc_col; e.stage = (uint8_t)s; e.pad = (uint8_t)(h & 1); e.img_off = off;
    e.a_kind = eng::A_MIX;
    e.a_off = (uint16_t)fpipe::a_buf_col(s, j); e.a_lo_delta = (uint16_t)(fpipe::a_buf_lo(s, j) / 16);
    uint16_t fl = 0;
    if (first) fl |= eng::F_ZERO_ACC;
    const bool pass0 = h == 0 && acc_col == 0;
    if (pass0 && j == 0) fl |= eng::F_STAGE_BEGIN;
    if (pass0 && (j == 4)) fl |= eng::F_WAIT_Q1;
    if (pass0 && (nk == 8 ? j == 0 : j == 8)) fl |= eng::F_WAIT_Q2;     // K = 128: the unused quarters' phases are
    if (pass0 && (nk == 8 ? j == 0 : j == 12)) fl |= eng::F_WAIT_Q3;    // consumed at the stage start
    if (h == 2) {
      if (last) fl |= eng::F_COMMIT_SIDE;                                // side pass: own barrier, not a stage half
    } else {
      if (last) fl |= eng::F_STAGE_END;
      if (last && s == NSTAGE - 1) fl |= eng::F_COMMIT_BOTH;
      if (XN == 96 && s == 4 && h == 1 && j == 0) fl |= eng::F_WAIT_SIDE;   // acc1 is re-initialised: side result parked
    }
    e.flags = fl;
    P.slab[n++] = e;
    off += (uint32_t)N * 32u;
  }
};
template <int XN = 64>
__host__ __device__ constexpr eng::Plan make_plan() {
  PlanFiller<XN> f{};
  for_each_slab<XN>(f);
  f.P.n_slabs = f.n; f.P.n_stages = NSTAGE;
  return f.P;
}
template <int XN = 64>
inline void build_plansrc(PlanSrc& S) {
  int n = 0;
  for_each_slab<XN>([&](int, int, int, int, int j, bool, bool, int wsel, int col0, int valid_n, int) {
    SrcDef d{};
    d.wsel = (uint8_t)wsel; d.kind = 1; d.row0 = (uint16_t)(16 * j); d.col0 = (uint16_t)col0;
    d.valid_k = 16; d.valid_n = (uint16_t)valid_n;
    S.s[n++] = d;
  });
}

__device__ eng::Plan d_plan_dpipe;
__device__ PlanSrc d_plansrc_dpipe;
__device__ eng::Plan d_plan_dpipe96;      // 4-D points
__device__ PlanSrc d_plansrc_dpipe96;
template <int NSPLIT, int XN = 64>
__global__ void __launch_bounds__(256) pack_dpipe_kernel(fused::PackSrc src, uint8_t* __restrict__ img) {
  const int i = blockIdx.y;
  const eng::Plan& P = XN == 64 ? d_plan_dpipe : d_plan_dpipe96;
  const PlanSrc& S = XN == 64 ? d_plansrc_dpipe : d_plansrc_dpipe96;
  if (i < P.n_slabs) fused::pack_slab_impl<NSPLIT>(P.slab[i], S.s[i], src, img);
}

template <int NSPLIT_, int XN_ = 64> struct Cfg {
  static constexpr int NSPLIT = NSPLIT_;
  static constexpr int XN = XN_;
  static constexpr eng::Plan PLAN = make_plan<XN_>();
  static constexpr int GROUP = NSPLIT_ == 1 ? 4 : 2;          // 312 slabs = 2 x 4 x 39 = 4 x 6 x 13
  static constexpr int NSLOT = NSPLIT_ == 1 ? 6 : 4;
  static constexpr int SLOT_BYTES = 16384;
  static_assert(PLAN.n_slabs % (GROUP * NSLOT) == 0, "ring size must divide the slab-group count");
  static constexpr int LO_BYTES = NSPLIT_ == 3 ? 3 * 32768 : 0;
  static constexpr int OFF_RING = 0;
  static constexpr int OFF_LO = NSLOT * SLOT_BYTES;
  static constexpr int OFF_C = OFF_LO + LO_BYTES;
  static constexpr int OFF_GX = OFF_C + ((fused::C_TOTAL * 4 + 127) / 128) * 128;   // [XN][128] fp32 skip-branch d(PE)
  static constexpr int OFF_OUT = OFF_GX + XN_ * 128 * 4;                              // [128][4]
  static constexpr int OFF_BAR = OFF_OUT + 128 * 4 * 4;
  static constexpr int SMEM_BYTES = OFF_BAR + (2 * NSLOT + 8) * 8 + 16;             // full/empty, accf[2], aq[4], side[2]
  static_assert(SMEM_BYTES <= 232448, "exceeds the 227 KB shared-memory limit");
};

// d(PE columns [C0, C0+32)) -> d(x): total d(PE) = layer-0 share (v) + skip share (gx_s), contracted with dPE/dx
template <int C0>
__device__ __forceinline__ void pe_contract(const uint32_t (&v)[32], const float* gx_s, int row, const float (&x)[3],
                                            float (&gx)[3]) {
#pragma unroll
  for (int j = 0; j < 32; ++j) {
    constexpr int dummy = 0; (void)dummy;
    const int i = C0 + j;
    const float g = __uint_as_float(v[j]) + gx_s[i * TILE_M + row];
    if (i < 3) gx[i] += g;
    else if (i < 63) {
      const int f = (i - 3) / 6, r = (i - 3) % 6, cc = r % 3;
      const float fr = (float)(1 << f);
      float sv, cv;
      fused::sincos_cw(x[cc] * fr, sv, cv);
      gx[cc] += (r < 3) ? fr * cv * g : -fr * sv * g;
    }
  }
}

// 4-D points (84 channels: x(4), then per frequency sin(4), cos(4)): column i of d(PE) -> d(x)
__device__ __forceinline__ void pe_contract4(int i, float g, const float (&x)[4], float (&gx)[4]) {
  if (i < 4) gx[i] += g;
  else if (i < 84) {
    const int f = (i - 4) >> 3, r = (i - 4) & 7, c = r & 3;
    const float fr = (float)(1 << f);
    float sv, cv;
    fused::sincos_cw(x[c] * fr, sv, cv);     // same evaluation as the forward's PE
    gx[c] += (r < 4) ? fr * cv * g : -fr * sv * g;
  }
}
// XN = 96, stage T4: park the side pass (skip share of d(PE), acc1 columns [0,96)) in shared memory and release acc1 for the
// h1 pass.  This warp owns columns [32 half, +32) and [64 + 16 half, +16).
__device__ __forceinline__ void side_drain(const fpipe::PCtx& c, float* gx_s, uint32_t lane_base, int half, int row,
                                           uint32_t tile_par) {
  eng::mbar_wait_a(c.aq_addr + 32, tile_par);
  tc::tc_fence_after();
  uint32_t va[32], vb[16];
  const int ca = half * 32, cb = 64 + half * 16;
  tc::tmem_ld32(c.e.tmem_acc + lane_base + 128 + ca, va);
  tc::tmem_ld16(c.e.tmem_acc + lane_base + 128 + cb, vb);
  tc::tmem_ld_wait();
#pragma unroll
  for (int j = 0; j < 32; ++j) gx_s[(ca + j) * TILE_M + row] = __uint_as_float(va[j]);
#pragma unroll
  for (int j = 0; j < 16; ++j) gx_s[(cb + j) * TILE_M + row] = __uint_as_float(vb[j]);
  tc::tc_fence_before();
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(c.aq_addr + 40) : "memory");
}

// One N-half of one stage for this warp: chunk cc = columns [H*128 + cc*64 + half*32, +32).
template <int NSPLIT, int T, int H, int XN = 64>
__device__ __forceinline__ void epi_half(const Args& a, const float* cst, const fpipe::PCtx& c, uint8_t* lo_area,
                                         float* gx_s, float* out_s, uint32_t lane_base, int half, int row, int tile,
                                         int64_t p, bool valid, const float4& gr) {
  constexpr bool SPLIT = NSPLIT == 3;
  if constexpr (T == 9) {
    eng::mbar_wait_a(c.accf_addr + H * 8, (uint32_t)(T & 1));
    if constexpr (H == 0 && XN == 96) {
      tc::tc_fence_after();
      uint32_t va[32], vb[16];
      const int ca = half * 32, cb = 64 + half * 16;
      tc::tmem_ld32(c.e.tmem_acc + lane_base + ca, va);
      tc::tmem_ld16(c.e.tmem_acc + lane_base + cb, vb);
      tc::tmem_ld_wait();
      float x[4] = {0.f, 0.f, 0.f, 0.f};
      if (valid) {
        const float4 q = *reinterpret_cast<const float4*>(a.pts + p * 4);
        x[0] = q.x; x[1] = q.y; x[2] = q.z; x[3] = q.w;
      }
      float gx[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int j = 0; j < 32; ++j) pe_contract4(ca + j, __uint_as_float(va[j]) + gx_s[(ca + j) * TILE_M + row], x, gx);
#pragma unroll
      for (int j = 0; j < 16; ++j) pe_contract4(cb + j, __uint_as_float(vb[j]) + gx_s[(cb + j) * TILE_M + row], x, gx);
#pragma unroll
      for (int cc = 0; cc < 4; ++cc) atomicAdd(out_s + row * 4 + cc, gx[cc]);
    } else if constexpr (H == 0) {
      // total d(PE row) = layer-0 share + skip share; contract with dPE/dx (Embedder backward)
      tc::tc_fence_after();
      uint32_t v[32];
      const int c0 = half * 32;
      tc::tmem_ld32(c.e.tmem_acc + lane_base + c0, v);
      tc::tmem_ld_wait();
      float x[3] = {0.f, 0.f, 0.f};
      if (valid) {
        const float* ry = a.rays + (p / a.S) * a.ray_cols;
        const float zz = a.z[p];
#pragma unroll
        for (int cc = 0; cc < 3; ++cc) x[cc] = __fadd_rn(ry[cc], __fmul_rn(ry[3 + cc], zz));
      }
      float gx[3] = {0.f, 0.f, 0.f};
      if (half == 0) pe_contract<0>(v, gx_s, row, x, gx);      // compile-time column indices per warp-half
      else pe_contract<32>(v, gx_s, row, x, gx);
      atomicAdd(out_s + row * 4 + 0, gx[0]);
      atomicAdd(out_s + row * 4 + 1, gx[1]);
      atomicAdd(out_s + row * 4 + 2, gx[2]);
    }
  } else {
    constexpr int mlayer = T == 0 ? -1 : 8 - T;      // trunk layer whose ReLU is differentiated here
    // ReLU mask words of this warp's two chunks, fetched before waiting for the accumulator
    uint32_t mk[2] = {0xffffffffu, 0xffffffffu};
    if constexpr (mlayer >= 0) {
      const uint32_t* mw = reinterpret_cast<const uint32_t*>(a.relu_bits + ((size_t)(tile * 9 + mlayer) * 2 + H) * 128 + row);
      mk[0] = __ldg(mw + half);
      mk[1] = __ldg(mw + 2 + half);
    }
    eng::mbar_wait_a(c.accf_addr + H * 8, (uint32_t)(T & 1));
    tc::tc_fence_after();
    uint32_t v[2][32];
#pragma unroll
    for (int cc = 0; cc < 2; ++cc) tc::tmem_ld32(c.e.tmem_acc + lane_base + H * 128 + half * 32 + cc * 64, v[cc]);
    tc::tmem_ld_wait();
#pragma unroll
    for (int cc = 0; cc < 2; ++cc) {
      const int cb = half * 32 + cc * 64;     // column inside the 128-wide half
      const int cu = H * 128 + cb;            // output column
      float f[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[cc][j]);
      if constexpr (T == 1) {                 // alpha head: g_h7 += g_alpha * w_alpha
#pragma unroll
        for (int j = 0; j < 32; ++j) f[j] = fmaf(gr.w, cst[fused::C_WALPHA + cu + j], f[j]);
      }
#pragma unroll
      for (int j = 0; j < 32; ++j) f[j] = eng::relu_bit(mk[cc], j) ? f[j] : 0.f;
      uint32_t hi[16], lo[16];
      eng::split32<SPLIT, false>(f, hi, lo);
      constexpr int buf_col = H == 0 ? ((T + 1) & 1) * 64 : 128;
      constexpr int buf_lo = H == 0 ? ((T + 1) & 1) * 32768 : 65536;
      tc::tmem_st16(c.e.tmem_ahi + lane_base + (uint32_t)(buf_col + (cb >> 1)), hi);
      if constexpr (SPLIT) {
#pragma unroll
        for (int g = 0; g < 4; ++g)
          *reinterpret_cast<uint4*>(lo_area + buf_lo + tc::canon_off(row, cb + 8 * g, TILE_M)) =
              make_uint4(lo[4 * g], lo[4 * g + 1], lo[4 * g + 2], lo[4 * g + 3]);
        tc::fence_proxy_async();
      }
      tc::tmem_st_wait();
      tc::tc_fence_before();
      asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(c.aq_addr + (uint32_t)(H * 2 + cc) * 8u) : "memory");
      // image of this stage's output (the next stage's A operand), consumed by the wgrad kernel
      if constexpr (T == 0) eng::dump32<SPLIT>(a.out_dfeat, tile, row, cu, hi, lo, c.e.pol_stream);
      else eng::dump32<SPLIT>(a.out_dz[8 - T], tile, row, cu, hi, lo, c.e.pol_stream);
    }
    if constexpr (T == 0 && H == 1) {
      // d(PE(dir)) sits in ACCX (committed with this half): warp-half 1 contracts it
      if (half == 1) {
        uint32_t w[32];
        tc::tmem_ld32(c.e.tmem_acc + lane_base + ACCX_COL, w);
        tc::tmem_ld_wait();
        float vd[3] = {0.f, 0.f, 0.f}, gv[3] = {0.f, 0.f, 0.f};
        if (valid) {
          if constexpr (XN == 96) {
            const float* vv = a.viewdirs + (p / a.S) * 3;
            vd[0] = vv[0]; vd[1] = vv[1]; vd[2] = vv[2];
          } else {
            const float* ry = a.rays + (p / a.S) * a.ray_cols;
            vd[0] = ry[8]; vd[1] = ry[9]; vd[2] = ry[10];
          }
        }
#pragma unroll
        for (int i = 0; i < 27; ++i) {
          const float g = __uint_as_float(w[i]);
          if (i < 3) gv[i] += g;
          else {
            const int f = (i - 3) / 6, r = (i - 3) % 6, cc = r % 3;
            const float fr = (float)(1 << f);
            float sv, cv;
            fused::sincos_cw(vd[cc] * fr, sv, cv);
            gv[cc] += (r < 3) ? fr * cv * g : -fr * sv * g;
          }
        }
        if (valid) { a.g_vd[p * 3] = gv[0]; a.g_vd[p * 3 + 1] = gv[1]; a.g_vd[p * 3 + 2] = gv[2]; }
      }
    }
    if constexpr (T == 4 && H == 1 && XN == 64) {
      // skip share of d(PE) (ACCX, 64 columns): parked in shared memory until T9; same thread reads it back
      uint32_t w[32];
      const int c0 = half * 32;
      tc::tmem_ld32(c.e.tmem_acc + lane_base + ACCX_COL + c0, w);
      tc::tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 32; ++j) gx_s[(c0 + j) * TILE_M + row] = __uint_as_float(w[j]);
    }
  }
}
// Trunk stages T2, T3, T5..T8 share one epilogue (mask of layer 8-T, image out_dz[8-T], no head term).  ROLL builds run
// them from one copy of the code with T as a run-time value (see fpipe::epi_half_rt for why).
template <int NSPLIT, int H>
__device__ __forceinline__ void epi_half_rt(const Args& a, const fpipe::PCtx& c, uint8_t* lo_area, uint32_t lane_base,
                                            int half, int row, int tile, int T) {
  constexpr bool SPLIT = NSPLIT == 3;
  const int mlayer = 8 - T;
  const uint32_t* mw = reinterpret_cast<const uint32_t*>(a.relu_bits + ((size_t)(tile * 9 + mlayer) * 2 + H) * 128 + row);
  uint32_t mk[2];
  mk[0] = __ldg(mw + half);
  mk[1] = __ldg(mw + 2 + half);
  eng::mbar_wait_a(c.accf_addr + H * 8, (uint32_t)(T & 1));
  tc::tc_fence_after();
  uint32_t v[2][32];
#pragma unroll
  for (int cc = 0; cc < 2; ++cc) tc::tmem_ld32(c.e.tmem_acc + lane_base + H * 128 + half * 32 + cc * 64, v[cc]);
  tc::tmem_ld_wait();
  const uint32_t par = (uint32_t)((T + 1) & 1);
  const uint32_t buf_col = H == 0 ? par * 64u : 128u;
  const uint32_t buf_lo = H == 0 ? par * 32768u : 65536u;
#pragma unroll
  for (int cc = 0; cc < 2; ++cc) {
    const int cb = half * 32 + cc * 64;     // column inside the 128-wide half
    const int cu = H * 128 + cb;            // output column
    float f[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) f[j] = eng::relu_bit(mk[cc], j) ? __uint_as_float(v[cc][j]) : 0.f;
    uint32_t hi[16], lo[16];
    eng::split32<SPLIT, false>(f, hi, lo);
    tc::tmem_st16(c.e.tmem_ahi + lane_base + buf_col + (uint32_t)(cb >> 1), hi);
    if constexpr (SPLIT) {
#pragma unroll
      for (int g = 0; g < 4; ++g)
        *reinterpret_cast<uint4*>(lo_area + buf_lo + tc::canon_off(row, cb + 8 * g, TILE_M)) =
            make_uint4(lo[4 * g], lo[4 * g + 1], lo[4 * g + 2], lo[4 * g + 3]);
      tc::fence_proxy_async();
    }
    tc::tmem_st_wait();
    tc::tc_fence_before();
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(c.aq_addr + (uint32_t)(H * 2 + cc) * 8u) : "memory");
    eng::dump32<SPLIT>(a.out_dz[8 - T], tile, row, cu, hi, lo, c.e.pol_stream);
  }
}
template <int NSPLIT, int XN = 64>     // trunk stages T2, T3, T5..T8 run from one copy of the epilogue code (epi_half_rt)
__global__ void __launch_bounds__(320, 1) field_dgrad_pipe_kernel(const __grid_constant__ Args a) {
  using C = Cfg<NSPLIT, XN>;
  constexpr bool SPLIT = NSPLIT == 3;
  extern __shared__ __align__(128) uint8_t qsm[];
  uint8_t* lo_area = qsm + C::OFF_LO;
  float* cst = reinterpret_cast<float*>(qsm + C::OFF_C);
  float* gx_s = reinterpret_cast<float*>(qsm + C::OFF_GX);
  float* out_s = reinterpret_cast<float*>(qsm + C::OFF_OUT);
  uint64_t* full = reinterpret_cast<uint64_t*>(qsm + C::OFF_BAR);
  uint64_t* empty = full + C::NSLOT;
  uint64_t* accf = empty + C::NSLOT;      // [2]
  uint64_t* aq = accf + 2;                // [4], then side[2] (XN = 96: side pass committed / side result parked)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(aq + 6);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    for (int i = 0; i < C::NSLOT; ++i) { tc::mbar_init(&full[i], 1); tc::mbar_init(&empty[i], 1); }
    tc::mbar_init(&accf[0], 1); tc::mbar_init(&accf[1], 1);
    for (int i = 0; i < 4; ++i) tc::mbar_init(&aq[i], 256);
    tc::mbar_init(&aq[4], 1); tc::mbar_init(&aq[5], 256);
    tc::fence_mbar_init();
  }
  for (int i = tid; i < fused::C_TOTAL; i += blockDim.x) cst[i] = a.cbuf[i];
  __syncthreads();
  if (warp == 1) tc::tmem_alloc(tmem_slot, 512);
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  fpipe::PCtx ctx;
  ctx.e.ring_addr = tc::smem_u32(qsm + C::OFF_RING); ctx.e.full_addr = tc::smem_u32(full); ctx.e.empty_addr = tc::smem_u32(empty);
  ctx.e.acc_full_addr = 0; ctx.e.a_ready_addr = 0;
  ctx.e.tmem_acc = tmem; ctx.e.tmem_ahi = tmem + 256; ctx.e.tmem_alo = 0; ctx.e.smem_a = 0;
  ctx.e.dbg = nullptr; ctx.e.dbg_tiles = 0;
  ctx.e.pol_keep = tc::policy_evict_last(); ctx.e.pol_stream = tc::policy_evict_first();
  ctx.accf_addr = tc::smem_u32(accf); ctx.aq_addr = tc::smem_u32(aq);
  ctx.smem_lo = tc::smem_u32(lo_area);
  ctx.dbg = nullptr; ctx.dbg_tiles = 0;

  const int tile_count = (int)blockIdx.x < a.num_tiles ? (a.num_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
  if (warp == 0) {
    if (lane == 0) eng::producer_loop_n<C>(ctx.e, a.wimg, tile_count);
  } else if (warp == 1) {
    if (lane == 0) {
      fpipe::mma_loop_r<C>(ctx, (int)blockIdx.x, a.num_tiles, (int)gridDim.x);
    }
  } else {
    const int quad = warp & 3, half = (warp - 2) >> 2;
    const int row = quad * 32 + lane;
    const uint32_t lane_base = (uint32_t)(quad * 32) << 16;
    uint32_t tile_par = 0;                  // parity of the once-per-tile side barriers (XN = 96)
    for (int tile = blockIdx.x; tile < a.num_tiles; tile += gridDim.x) {
      const int64_t p = (int64_t)tile * TILE_M + row;
      const bool valid = p < a.P;
      float4 gr = make_float4(0.f, 0.f, 0.f, 0.f);
      if (valid) gr = *reinterpret_cast<const float4*>(a.g_raw + p * 4);
      // ---- E0: rgb head dgrad (fp32), ReLU mask of the view layer -> dZ_v: chunk cc = columns [cc*64 + half*32, +32)
      if (half == 0) *reinterpret_cast<float4*>(out_s + row * 4) = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int cc = 0; cc < 2; ++cc) {
        const int c0 = cc * 64 + half * 32;
        // view-layer masks: halves of 64 columns, 32-bit word = (column % 64) / 32
        const uint32_t mv = __ldg(reinterpret_cast<const uint32_t*>(a.relu_bits + ((size_t)(tile * 9 + 8) * 2 + cc) * 128 + row) + half);
        float f[32];
#pragma unroll
        for (int j = 0; j < 32; ++j)
          f[j] = gr.x * cst[fused::C_WRGB + c0 + j] + gr.y * cst[fused::C_WRGB + 128 + c0 + j] +
                 gr.z * cst[fused::C_WRGB + 256 + c0 + j];
#pragma unroll
        for (int j = 0; j < 32; ++j) f[j] = eng::relu_bit(mv, j) ? f[j] : 0.f;
        uint32_t hi[16], lo[16];
        eng::split32<SPLIT, false>(f, hi, lo);
        tc::tmem_st16(ctx.e.tmem_ahi + lane_base + (uint32_t)(c0 >> 1), hi);          // T0's A operand: P0
        if constexpr (SPLIT) {
#pragma unroll
          for (int g = 0; g < 4; ++g)
            *reinterpret_cast<uint4*>(lo_area + tc::canon_off(row, c0 + 8 * g, TILE_M)) =
                make_uint4(lo[4 * g], lo[4 * g + 1], lo[4 * g + 2], lo[4 * g + 3]);
        }
        eng::dump32<SPLIT>(a.out_dzv, tile, row, c0, hi, lo, ctx.e.pol_stream);
      }
      if constexpr (SPLIT) tc::fence_proxy_async();
      tc::tmem_st_wait();
      tc::tc_fence_before();
      for (int i = 0; i < 4; ++i) tc::mbar_arrive(&aq[i]);
      epi_half<NSPLIT, 0, 0, XN>(a, cst, ctx, lo_area, gx_s, out_s, lane_base, half, row, tile, p, valid, gr);
      epi_half<NSPLIT, 0, 1, XN>(a, cst, ctx, lo_area, gx_s, out_s, lane_base, half, row, tile, p, valid, gr);
      epi_half<NSPLIT, 1, 0, XN>(a, cst, ctx, lo_area, gx_s, out_s, lane_base, half, row, tile, p, valid, gr);
      epi_half<NSPLIT, 1, 1, XN>(a, cst, ctx, lo_area, gx_s, out_s, lane_base, half, row, tile, p, valid, gr);
#pragma unroll 1
      for (int T = 2; T < 9; ++T) {
        if (T == 4) {
          if constexpr (XN == 96) side_drain(ctx, gx_s, lane_base, half, row, tile_par);
          epi_half<NSPLIT, 4, 0, XN>(a, cst, ctx, lo_area, gx_s, out_s, lane_base, half, row, tile, p, valid, gr);
          epi_half<NSPLIT, 4, 1, XN>(a, cst, ctx, lo_area, gx_s, out_s, lane_base, half, row, tile, p, valid, gr);
        } else {
          epi_half_rt<NSPLIT, 0>(a, ctx, lo_area, lane_base, half, row, tile, T);
          epi_half_rt<NSPLIT, 1>(a, ctx, lo_area, lane_base, half, row, tile, T);
        }
      }
      epi_half<NSPLIT, 9, 0, XN>(a, cst, ctx, lo_area, gx_s, out_s, lane_base, half, row, tile, p, valid, gr);
      epi_half<NSPLIT, 9, 1, XN>(a, cst, ctx, lo_area, gx_s, out_s, lane_base, half, row, tile, p, valid, gr);
      tc::tc_fence_before();
      asm volatile("bar.sync 1, 256;" ::: "memory");
      if (half == 0 && valid) {
        if constexpr (XN == 96) *reinterpret_cast<float4*>(a.g_pts + p * 4) = *reinterpret_cast<const float4*>(out_s + row * 4);
        else { a.g_pts[p * 3] = out_s[row * 4]; a.g_pts[p * 3 + 1] = out_s[row * 4 + 1]; a.g_pts[p * 3 + 2] = out_s[row * 4 + 2]; }
      }
      asm volatile("bar.sync 1, 256;" ::: "memory");
      tile_par ^= 1u;
    }
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 1) tc::tmem_dealloc(tmem, 512);
}

}  // namespace dpipe
}  // namespace scnerf
