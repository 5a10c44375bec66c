// Fused field forward, N-half PIPELINED variant (sm_100a, tcgen05).
//
// Same maths, weights, PE code and outputs as field_tc_fused.cuh; what changes is the schedule inside a
// 128-sample tile.  There, a layer's MMAs and its epilogue are serial (the epilogue produces the next
// layer's A operand and TMEM is full), so the tensor pipe idles ~40 % of the time.  Here every GEMM stage
// runs as TWO N-halves (128 output columns each, 64-cycle N=128 MMAs — measured at the nominal rate,
// tools/mma_bench.py) into two 128-column accumulators:
//
//   tensor pipe :  | L.h0 | L.h1 | L+1.h0 (K 0..127) (K 128..255) | L+1.h1 | ...
//   epilogue    :         | epi(L.h0) | epi(L.h1)  |                | epi(L+1.h0) | ...
//
// epi(L.h0) runs under L.h1's MMAs and writes columns 0..127 of the next layer's A operand, so L+1.h0 can
// start on its first eight K-slabs as soon as L.h1 has been issued; only its last eight K-slabs wait for
// epi(L.h1).  Buffers: the first A half is double-buffered (P0/P1, written while the previous layer's MMAs
// still read theirs), the second half is single (Q: written after the layer that read it has completed).
//
//   TMEM (512 columns): acc0 [0,128)  acc1 [128,256)  A_hi: P0 [256,320)  P1 [320,384)  Q [384,448)
//   split-bf16 only   : A_lo lives in shared memory, canonical K-major images P0/P1/Q x 32 KB, consumed by
//                       SS-mode MMAs (lo x W_hi); hi x W_hi and hi x W_lo stay TS-mode.
//
// Barriers (all count their phases per stage, 10 per tile, so parity = stage & 1):
//   accf[h]  MMA -> epilogue : N-half h of the stage is complete          (tcgen05.commit)
//   aq[0..3] epilogue -> MMA : the accumulator half is drained and A columns [64q, 64q+64) are written
//                              (256 arrivals each).  The epilogue hands the next layer's operand over in
//                              64-column quarters, so the K-slabs that read quarter q start while the
//                              quarters behind it are still being converted.
//
// Reference: NeRF/run_nerf_helpers.py:24-72,105-128, NeRF/create_nerf.py:18-32.
#pragma once
#include "common.cuh"
#include "tc_prims.cuh"
#include "tc_engine.cuh"
#include "field_tc_fused.cuh"

namespace scnerf {
namespace fpipe {

using eng::TILE_M;
using fused::ALay;
using fused::Args;
using fused::NSTAGE;
using fused::PlanSrc;
using fused::SrcDef;
using fused::StageDef;
using fused::stage_def;
using fused::C_SCAL;
using fused::C_TOTAL;
using fused::C_WALPHA;
using fused::C_WRGB;

constexpr int PLAN_MULT = 24;     // slabs per tile padded to a multiple of GROUP x NSLOT of both precisions (8, 24)
// TMEM columns of the A_hi buffers relative to the A base (32-bit columns, two bf16 each)
__host__ __device__ constexpr int a_buf_col(int stage, int j) { return j < 8 ? (stage & 1) * 64 + j * 8 : 128 + (j - 8) * 8; }
// byte offset of the k16 slab j of stage `stage`'s A operand in the shared-memory lo area
__host__ __device__ constexpr int a_buf_lo(int stage, int j) { return j < 8 ? (stage & 1) * 32768 + j * 4096 : 65536 + (j - 8) * 4096; }

// Slab order inside one N-half pass: bias (ONES), PE(pts) slabs, PE(dir) slabs, hidden slabs 0..15 —
// everything that does not depend on the previous epilogue first.
template <int XS, class F>
__host__ __device__ constexpr void for_each_slab(F&& f) {
  // f(stage, h, kind {0 bias, 1 X, 2 V, 3 H}, j, first_in_pass, last_in_pass)
  for (int s = 0; s < NSTAGE; ++s) {
    const StageDef d = stage_def<XS>(s);
    const int nk = 1 + d.kx + d.kv + d.kh;
    for (int h = 0; h < 2; ++h) {
      int i = 0;
      f(s, h, 0, 0, true, nk == 1); ++i;
      for (int j = 0; j < d.kx; ++j, ++i) f(s, h, 1, j, false, i == nk - 1);
      for (int j = 0; j < d.kv; ++j, ++i) f(s, h, 2, j, false, i == nk - 1);
      for (int j = 0; j < d.kh; ++j, ++i) f(s, h, 3, j, false, i == nk - 1);
    }
  }
}

template <int NSPLIT, int XS>
struct PlanFiller {
  eng::Plan P{};
  int n = 0;
  uint32_t off = 0;
  __host__ __device__ constexpr void operator()(int s, int h, int kind, int j, bool first, bool last) {
    using L = ALay<NSPLIT, XS>;
    const StageDef d = stage_def<XS>(s);
    const int nh = d.N / 2;
    eng::SlabDef e{};
    e.n = (uint16_t)nh; e.acc_col = (uint16_t)(h * 128); e.stage = (uint8_t)s; e.pad = (uint8_t)h; e.img_off = off;
    uint16_t fl = 0;
    if (first) fl |= eng::F_ZERO_ACC;
    if (first && h == 0) fl |= eng::F_STAGE_BEGIN;                       // wait aq[0] (acc0 drained)
    if (h == 0 && ((s == 0 && first) || (kind == 3 && j == 4))) fl |= eng::F_WAIT_Q1;
    if (h == 0 && ((s == 0 && first) || (kind == 3 && j == 8))) fl |= eng::F_WAIT_Q2;    // (acc1 drained)
    if (h == 0 && ((s == 0 && first) || (kind == 3 && j == 12))) fl |= eng::F_WAIT_Q3;
    if (last) fl |= eng::F_STAGE_END;                                    // commit accf[h]
    if (kind == 0) { e.a_kind = eng::A_SMEM; e.a_off = L::ONES / 16; fl |= eng::F_HI_ONLY_A; }
    else if (kind == 1) { e.a_kind = eng::A_SMEM; e.a_off = (uint16_t)((L::XHI + j * 4096) / 16); e.a_lo_delta = (L::XLO - L::XHI) / 16; }
    else if (kind == 2) { e.a_kind = eng::A_SMEM; e.a_off = (uint16_t)((L::VHI + j * 4096) / 16); e.a_lo_delta = (L::VLO - L::VHI) / 16; }
    else { e.a_kind = eng::A_MIX; e.a_off = (uint16_t)a_buf_col(s, j); e.a_lo_delta = (uint16_t)(a_buf_lo(s, j) / 16); }
    e.flags = fl;
    P.slab[n++] = e;
    off += (uint32_t)nh * 32u;
  }
};
template <int NSPLIT, int XS>
__host__ __device__ constexpr eng::Plan make_plan() {
  PlanFiller<NSPLIT, XS> f{};
  for_each_slab<XS>(f);
  // pad with zero-weight N=16 slabs (A = ONES, B = 0) in front of the tile's last commit, so that the ring
  // size divides the slab-group count and ring slot / parity stay compile-time
  const int n_pad = (PLAN_MULT - f.n % PLAN_MULT) % PLAN_MULT;
  if (n_pad > 0) {
    f.P.slab[f.n - 1].flags = (uint16_t)(f.P.slab[f.n - 1].flags & ~eng::F_STAGE_END);
    for (int k = 0; k < n_pad; ++k) {
      eng::SlabDef e{};
      e.n = 16; e.acc_col = 128; e.stage = (uint8_t)(NSTAGE - 1); e.pad = 1; e.img_off = f.off;
      e.a_kind = eng::A_SMEM; e.a_off = ALay<NSPLIT, XS>::ONES / 16;
      e.flags = (uint16_t)(eng::F_HI_ONLY_A | (k == n_pad - 1 ? eng::F_STAGE_END : 0));
      f.P.slab[f.n++] = e;
      f.off += 16u * 32u;
    }
  }
  f.P.n_slabs = f.n; f.P.n_stages = NSTAGE;
  return f.P;
}
template <int XS>
inline void build_plansrc(PlanSrc& S) {
  constexpr int IN_CH = XS == 4 ? 63 : 84;
  int n = 0;
  for_each_slab<XS>([&](int s, int h, int kind, int j, bool, bool) {
    const StageDef d = stage_def<XS>(s);
    SrcDef q{};
    q.wsel = (uint8_t)s; q.valid_n = (uint16_t)(d.N / 2); q.pad = (uint16_t)(h * (d.N / 2));   // pad = first output row
    if (kind == 0) q.kind = 2;
    else if (kind == 1) { q.col0 = (uint16_t)(16 * j); q.valid_k = (uint16_t)std::max(0, std::min(16, IN_CH - 16 * j)); }
    else if (kind == 2) { q.col0 = (uint16_t)(256 + 16 * j); q.valid_k = (uint16_t)std::min(16, 27 - 16 * j); }
    else { q.col0 = (uint16_t)((d.kx ? IN_CH : 0) + 16 * j); q.valid_k = 16; }
    S.s[n++] = q;
  });
  for (; n % PLAN_MULT != 0; ++n) { SrcDef q{}; q.kind = 3; S.s[n] = q; }   // zeros
}

template <int NSPLIT_, int XS_ = 4> struct Cfg {
  static constexpr int NSPLIT = NSPLIT_;
  static constexpr int XS = XS_;
  static constexpr eng::Plan PLAN = make_plan<NSPLIT_, XS_>();
  // half-slabs are 8 KB (split) / 4 KB (bf16): 16 KB ring slots keep the issue thread's per-slot work
  // (wait + commit) amortised over 6 / 4 MMAs
  static constexpr int GROUP = NSPLIT_ == 1 ? 4 : 2;
  static constexpr int NSLOT = NSPLIT_ == 1 ? 6 : (XS_ == 4 ? 4 : 3);
  static constexpr int SLOT_BYTES = 16384;
  static_assert(PLAN.n_slabs % (GROUP * NSLOT) == 0, "ring size must divide the slab-group count");
  static constexpr int LO_BYTES = NSPLIT_ == 3 ? 3 * 32768 : 0;
  static constexpr int OFF_RING = 0;
  static constexpr int OFF_A = NSLOT * SLOT_BYTES;
  static constexpr int OFF_LO = OFF_A + ALay<NSPLIT_, XS_>::BYTES;
  static constexpr int OFF_C = OFF_LO + LO_BYTES;
  static constexpr int OFF_OUT = OFF_C + ((C_TOTAL * 4 + 127) / 128) * 128;   // [128][4] head partial sums
  static constexpr int COMP_MAX_G = 3;                            // fused composite: up to 3 tiles (384 samples) of raw values
  static constexpr int OFF_COMP = OFF_OUT + 128 * 4 * 4;
  static constexpr int OFF_BAR = OFF_COMP + COMP_MAX_G * 128 * 4 * 4;
  static constexpr int SMEM_BYTES = OFF_BAR + (2 * NSLOT + 6) * 8 + 16;
  static_assert(SMEM_BYTES <= 232448, "exceeds the 227 KB shared-memory limit");
};

struct PCtx {
  eng::Ctx e;                  // ring / full / empty / tmem_acc / tmem_ahi (= A base) / smem_a / policies
  uint32_t accf_addr;          // accf[0], accf[1] 8 bytes apart
  uint32_t aq_addr;            // aq[0..3] 8 bytes apart
  uint32_t smem_lo;            // shared address of the lo area
  long long* dbg; int dbg_tiles;
};
// timeline: [tile][stage][16] clock64 stamps of CTA 0 (tools/timeline_pipe.py):
//   0 MMA passed aq[0], 1 MMA passed aq[2], 2 MMA committed h0, 3 MMA committed h1,
//   4 epilogue saw accf[0], 5 epilogue finished h0, 6 epilogue saw accf[1], 7 epilogue finished h1,
//   8 + 4h: accumulator loads of half h done, 9 + 4h / 10 + 4h: quarter 2h / 2h+1 handed over
// Compiled in only with -DSCNERF_TIMELINE (build.sh -DSCNERF_TIMELINE; tools/timeline_pipe.py needs such a build): even
// predicated off, the stamps cost the single MMA-issuing thread a spill/reload pair around every clock read once the
// kernel sat at its 168-register cap (round 2: 60 -> 992 bytes of spill code, +1 ms per step on the training forward).
__device__ __forceinline__ void stamp(const PCtx& c, int tile_iter, int stage, int slot) {
#ifdef SCNERF_TIMELINE
  if (c.dbg != nullptr && blockIdx.x == 0 && tile_iter < c.dbg_tiles)
    c.dbg[((size_t)tile_iter * NSTAGE + stage) * 16 + slot] = clock64();
#else
  (void)c; (void)tile_iter; (void)stage; (void)slot;
#endif
}

template <class K, int I>
__device__ __forceinline__ void mma_step(const PCtx& c, uint32_t tp, int tile_iter) {
  constexpr eng::SlabDef d = K::PLAN.slab[I];
  constexpr bool SPLIT = K::NSPLIT == 3;
  constexpr int G = I / K::GROUP;
  constexpr int idx = G % K::NSLOT, wrap = G / K::NSLOT;
  constexpr bool wraps_odd = (((K::PLAN.n_slabs / K::GROUP) / K::NSLOT) & 1) != 0;
  constexpr uint32_t in_slot = eng::group_bytes<K>(G * K::GROUP, I);
  static_assert((K::PLAN.n_stages & 1) == 0, "stage parity assumes an even number of stages per tile");
  if constexpr ((d.flags & eng::F_STAGE_BEGIN) != 0) {
    eng::mbar_wait_a(c.aq_addr, (uint32_t)(d.stage & 1));
    tc::tc_fence_after();
    stamp(c, tile_iter, d.stage, 0);
  }
  if constexpr ((d.flags & eng::F_WAIT_Q1) != 0) {
    eng::mbar_wait_a(c.aq_addr + 8, (uint32_t)(d.stage & 1));
    tc::tc_fence_after();
  }
  if constexpr ((d.flags & eng::F_WAIT_Q2) != 0) {
    eng::mbar_wait_a(c.aq_addr + 16, (uint32_t)(d.stage & 1));
    tc::tc_fence_after();
    stamp(c, tile_iter, d.stage, 1);
  }
  if constexpr ((d.flags & eng::F_WAIT_Q3) != 0) {
    eng::mbar_wait_a(c.aq_addr + 24, (uint32_t)(d.stage & 1));
    tc::tc_fence_after();
  }
  if constexpr ((d.flags & eng::F_WAIT_SIDE) != 0) {     // side[1] (8 bytes after aq[3] + side[0]): side result parked by the epilogue
    eng::mbar_wait_a(c.aq_addr + 40, tp);
    tc::tc_fence_after();
  }
  if constexpr (I % K::GROUP == 0) {
#ifdef SCNERF_TIMELINE
    // slots 11 / 15: cycles the issue thread spent waiting for weight slabs (ring full barriers) in N-half 0 / 1 of this stage
    const long long tw0 = clock64();
#endif
    eng::mbar_wait_a(c.e.full_addr + idx * 8, (uint32_t)(wrap & 1) ^ (wraps_odd ? tp : 0u));
    tc::tc_fence_after();
#ifdef SCNERF_TIMELINE
    if (c.dbg != nullptr && blockIdx.x == 0 && tile_iter < c.dbg_tiles) {
      long long* w = c.dbg + ((size_t)tile_iter * NSTAGE + d.stage) * 16 + 11 + 4 * d.pad;
      const long long dt = clock64() - tw0;
      *w = ((d.flags & eng::F_ZERO_ACC) != 0 ? 0 : *w) + dt;
    }
#endif
  }
  constexpr uint32_t idesc = tc::idesc_bf16_f32(TILE_M, d.n);
  constexpr uint32_t LBO_B = (uint32_t)d.n * 16u;
  const uint32_t slot = c.e.ring_addr + idx * K::SLOT_BYTES + in_slot;
  const uint64_t b_hi = eng::desc_at<LBO_B, 128>(slot);
  const uint64_t b_lo = eng::desc_at<LBO_B, 128>(slot + (uint32_t)d.n * 32u);
  const uint32_t acc = c.e.tmem_acc + d.acc_col;
  constexpr uint32_t first = (d.flags & eng::F_ZERO_ACC) ? 0u : 1u;
  if constexpr (d.a_kind == eng::A_MIX) {
    tc::mma_ts(acc, c.e.tmem_ahi + d.a_off, b_hi, idesc, first);
    if constexpr (SPLIT) {
      tc::mma_ss(acc, eng::desc_at<2048, 128>(c.smem_lo + (uint32_t)d.a_lo_delta * 16u), b_hi, idesc, 1);
      tc::mma_ts(acc, c.e.tmem_ahi + d.a_off, b_lo, idesc, 1);
    }
  } else {
    const uint32_t a_addr = c.e.smem_a + (uint32_t)d.a_off * 16u;
    const uint64_t a_hi = eng::desc_at<2048, 128>(a_addr);
    tc::mma_ss(acc, a_hi, b_hi, idesc, first);
    if constexpr (SPLIT) {
      if constexpr ((d.flags & eng::F_HI_ONLY_A) == 0)
        tc::mma_ss(acc, eng::desc_at<2048, 128>(a_addr + (uint32_t)d.a_lo_delta * 16u), b_hi, idesc, 1);
      tc::mma_ss(acc, a_hi, b_lo, idesc, 1);
    }
  }
  if constexpr (I % K::GROUP == K::GROUP - 1) eng::commit_a(c.e.empty_addr + idx * 8);
  if constexpr ((d.flags & eng::F_COMMIT_SIDE) != 0) eng::commit_a(c.aq_addr + 32);    // side[0]: side pass complete
  if constexpr ((d.flags & eng::F_STAGE_END) != 0) {
    eng::commit_a(c.accf_addr + d.pad * 8);
    if constexpr ((d.flags & eng::F_COMMIT_BOTH) != 0) eng::commit_a(c.accf_addr + (d.pad ^ 1) * 8);   // stage without a second N-half
    stamp(c, tile_iter, d.stage, 2 + d.pad);
  }
}
template <class K, size_t... Is>
__device__ __forceinline__ void mma_tile(const PCtx& c, uint32_t tp, int tile_iter, std::index_sequence<Is...>) {
  (mma_step<K, (int)Is>(c, tp, tile_iter), ...);
}
// This CTA's tiles as a range.  The FORM of this loop matters: the issue thread's 336 unrolled steps leave ptxas no spare
// register at the kernel's 168 cap, and a counted loop (`for it < count`) instead of this compare-against-end form made
// it spill a register around every mbarrier wait of the issue thread (60 -> 976 bytes of spill code, measured +1 ms per
// step on the training forward; bisected with -Xptxas -v, profiles/README.md round 2).
template <class K>
__device__ __forceinline__ void mma_loop_r(const PCtx& c, int first, int end, int stride) {
  uint32_t tp = 0;
  int it = 0;
  for (int tile = first; tile < end; tile += stride, tp ^= 1u, ++it)
    mma_tile<K>(c, tp, it, std::make_index_sequence<K::PLAN.n_slabs>{});
}

// Alpha-composite of the rays of one group from the raw values parked in shared memory (one warp per ray).  Kept OUT
// of line: inlined into the persistent kernel it raised the epilogue's register pressure (72 -> 2 KB of spill code,
// training forward 3.5 -> 4.6 ms per step); it runs once per ray group, a call costs nothing there.
__device__ __noinline__ void composite_group(const CompositeArgs& ca, const float* comp_s, int64_t ray0, int n_rays, int S,
                                             int ewarp, int lane) {
  for (int rr = ewarp; rr < n_rays; rr += 8)
    composite_ray(ca, ray0 + rr, lane, comp_s + (size_t)rr * S * 4, 4);
}

// Epilogue of N-half H of stage S for this warp's columns (compile-time stage parameters).
// 256-wide stages: the N-half is handed over as two 64-column quarters; chunk cc of warp `half` (0/1) is
// columns [H*128 + cc*64 + half*32, +32).  View layer (N-halves of 64): one chunk, [H*64 + half*32, +32).
template <int NSPLIT, int S, int H>
__device__ __forceinline__ void epi_half(const Args& a, const float* cst, const PCtx& c, uint8_t* lo_area,
                                         uint32_t lane_base, int half, int row, int tile, int64_t p, bool valid,
                                         float& alpha, float (&rgb)[3], int tile_iter) {
  constexpr bool SPLIT = NSPLIT == 3;
  constexpr StageDef d = stage_def(S);
  constexpr int NH = d.N / 2, NW = NH / 2, nchunk = NW / 32;   // 128/64/2 (256-wide) or 64/32/1 (view layer)
  constexpr int CSTEP = nchunk == 2 ? 64 : 32;   // column distance between this warp's chunks
  const int cw = half * 32;                 // first column inside the N-half (chunk 0)
  const int cbase = H * NH + cw;            // first layer-output column
  eng::mbar_wait_a(c.accf_addr + H * 8, (uint32_t)(S & 1));
  tc::tc_fence_after();
  if (threadIdx.x == 64) stamp(c, tile_iter, S, 4 + 2 * H);
  uint32_t v[nchunk][32];
#pragma unroll
  for (int cc = 0; cc < nchunk; ++cc) tc::tmem_ld32(c.e.tmem_acc + lane_base + H * 128 + cw + cc * CSTEP, v[cc]);
  tc::tmem_ld_wait();
  if (threadIdx.x == 64) stamp(c, tile_iter, S, 8 + 4 * H);
  uint32_t mbits[nchunk];                   // ReLU-mask words of this warp's chunks (training), see eng::relu_mask16
#pragma unroll
  for (int cc = 0; cc < nchunk; ++cc) {
    const int cu = cbase + cc * CSTEP;      // layer-output column of v[cc][0]
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
    if (S == 9 && a.img_out[9].base == nullptr) mbits[cc] = eng::relu_mask32f(v[cc]);
    if (S != 9 || a.img_out[9].base != nullptr) {
      uint32_t hi[16], lo[16];
      if constexpr (d.relu != 0) eng::split32_relu<SPLIT>(f, hi, lo, mbits[cc]);
      else eng::split32<SPLIT, false>(f, hi, lo);
      if constexpr (S != 9) {
        // next stage's A operand: columns [0,128) -> P[(S+1)&1], [128,256) -> Q
        const int cb = cw + cc * CSTEP;     // column inside the 128-wide buffer
        constexpr int buf_col = H == 0 ? ((S + 1) & 1) * 64 : 128;
        constexpr int buf_lo = H == 0 ? ((S + 1) & 1) * 32768 : 65536;
        tc::tmem_st16(c.e.tmem_ahi + lane_base + (uint32_t)(buf_col + (cb >> 1)), hi);
        if constexpr (SPLIT) {
#pragma unroll
          for (int g = 0; g < 4; ++g)
            *reinterpret_cast<uint4*>(lo_area + buf_lo + tc::canon_off(row, cb + 8 * g, TILE_M)) =
                make_uint4(lo[4 * g], lo[4 * g + 1], lo[4 * g + 2], lo[4 * g + 3]);
        }
      }
      if constexpr (S < 9) {
        // hand this quarter over before converting the next one
        if constexpr (SPLIT) tc::fence_proxy_async();
        tc::tmem_st_wait();
        tc::tc_fence_before();
        asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(c.aq_addr + (uint32_t)(H * 2 + cc) * 8u) : "memory");
        if (threadIdx.x == 64) stamp(c, tile_iter, S, 9 + 4 * H + cc);
      }
      if (a.img_out[S].base != nullptr) eng::dump32<SPLIT>(a.img_out[S], tile, row, cu, hi, lo, c.e.pol_stream);
    }
  }
  if (threadIdx.x == 64) stamp(c, tile_iter, S, 5 + 2 * H);
  if constexpr (d.relu != 0) {
    if (a.relu_bits != nullptr) {      // after the hand-off: off the MMA's critical path
      // layout (fused::Args::relu_bits): [tile][9 layers][2 halves][128 rows] x 16 B, bit = column inside the
      // half; this warp's 32-bit words are number cw/32 .. of half H (view layer: halves are 64 columns)
      constexpr int L = S == 9 ? 8 : S;
      uint32_t* mw = reinterpret_cast<uint32_t*>(a.relu_bits + ((size_t)(tile * 9 + L) * 2 + H) * 128 + row) + (cw >> 5);   // chunk cc -> word (cw + cc * CSTEP) / 32
#pragma unroll
      for (int cc = 0; cc < nchunk; ++cc) mw[cc * (CSTEP / 32)] = mbits[cc];
    }
  }
}
// Stages 0..6 share one epilogue (256-wide, ReLU, no head term).  ROLL builds run them from ONE copy of the code with
// the stage as a run-time value (buffer parity, barrier parity, image / dump pointers indexed in the kernel
// parameters): the fully unrolled epilogue is ~220 KB of straight-line code per tile, executed once, and the ncu stall
// samples of the training forward put 27 % of its stalls on instruction fetch (profiles/README.md section 5).
template <int NSPLIT, int H>
__device__ __forceinline__ void epi_half_rt(const Args& a, const PCtx& c, uint8_t* lo_area, uint32_t lane_base, int half,
                                            int row, int tile, int64_t p, bool valid, int tile_iter, int S) {
  constexpr bool SPLIT = NSPLIT == 3;
  const int cw = half * 32;                 // first column inside the N-half (chunk 0)
  const int cbase = H * 128 + cw;           // first layer-output column
  eng::mbar_wait_a(c.accf_addr + H * 8, (uint32_t)(S & 1));
  tc::tc_fence_after();
  if (threadIdx.x == 64) stamp(c, tile_iter, S, 4 + 2 * H);
  uint32_t v[2][32];
#pragma unroll
  for (int cc = 0; cc < 2; ++cc) tc::tmem_ld32(c.e.tmem_acc + lane_base + H * 128 + cw + cc * 64, v[cc]);
  tc::tmem_ld_wait();
  if (threadIdx.x == 64) stamp(c, tile_iter, S, 8 + 4 * H);
  // next stage's A operand: columns [0,128) -> P[(S+1)&1], [128,256) -> Q
  const uint32_t par = (uint32_t)((S + 1) & 1);
  const uint32_t buf_col = H == 0 ? par * 64u : 128u;
  const uint32_t buf_lo = H == 0 ? par * 32768u : 65536u;
  float* const dump_s = a.dump[S];
  const bool has_img = a.img_out[S].base != nullptr;
  uint32_t mbits[2];                        // ReLU-mask words of this warp's two chunks (eng::relu_mask16)
#pragma unroll
  for (int cc = 0; cc < 2; ++cc) {
    const int cu = cbase + cc * 64;         // layer-output column of v[cc][0]
    float f[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[cc][j]);
    if (dump_s != nullptr && valid) {
      float* dp = dump_s + p * a.dump_ld[S] + cu;
#pragma unroll
      for (int j = 0; j < 32; ++j) dp[j] = fmaxf(f[j], 0.f);
    }
    uint32_t hi[16], lo[16];
    eng::split32_relu<SPLIT>(f, hi, lo, mbits[cc]);
    const int cb = cw + cc * 64;            // column inside the 128-wide buffer
    tc::tmem_st16(c.e.tmem_ahi + lane_base + buf_col + (uint32_t)(cb >> 1), hi);
    if constexpr (SPLIT) {
#pragma unroll
      for (int g = 0; g < 4; ++g)
        *reinterpret_cast<uint4*>(lo_area + buf_lo + tc::canon_off(row, cb + 8 * g, TILE_M)) =
            make_uint4(lo[4 * g], lo[4 * g + 1], lo[4 * g + 2], lo[4 * g + 3]);
      tc::fence_proxy_async();
    }
    // hand this quarter over before converting the next one
    tc::tmem_st_wait();
    tc::tc_fence_before();
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(c.aq_addr + (uint32_t)(H * 2 + cc) * 8u) : "memory");
    if (threadIdx.x == 64) stamp(c, tile_iter, S, 9 + 4 * H + cc);
    if (has_img) eng::dump32<SPLIT>(a.img_out[S], tile, row, cu, hi, lo, c.e.pol_stream);
  }
  if (threadIdx.x == 64) stamp(c, tile_iter, S, 5 + 2 * H);
  if (a.relu_bits != nullptr) {      // after the hand-off: off the MMA's critical path (layout: see epi_half)
    uint32_t* mw = reinterpret_cast<uint32_t*>(a.relu_bits + ((size_t)(tile * 9 + S) * 2 + H) * 128 + row) + (cw >> 5);
    mw[0] = mbits[0];
    mw[2] = mbits[1];
  }
}
template <int NSPLIT, size_t... Ss>
__device__ __forceinline__ void epi_tile(const Args& a, const float* cst, const PCtx& c, uint8_t* lo_area,
                                         uint32_t lane_base, int half, int row, int tile, int64_t p, bool valid,
                                         float& alpha, float (&rgb)[3], int tile_iter, std::index_sequence<Ss...>) {
  ((epi_half<NSPLIT, (int)Ss, 0>(a, cst, c, lo_area, lane_base, half, row, tile, p, valid, alpha, rgb, tile_iter),
    epi_half<NSPLIT, (int)Ss, 1>(a, cst, c, lo_area, lane_base, half, row, tile, p, valid, alpha, rgb, tile_iter)), ...);
}

template <int NSPLIT, int XS = 4, int ROLL = 0>     // ROLL: 0 fully unrolled epilogue (4-D points), 1 stages 0-6 from one copy of the code (3-D points)
__global__ void __launch_bounds__(320, 1) field_fwd_pipe_kernel(const __grid_constant__ Args a) {
  using C = Cfg<NSPLIT, XS>;
  using L = ALay<NSPLIT, XS>;
  constexpr bool SPLIT = NSPLIT == 3;
  extern __shared__ __align__(128) uint8_t psm[];
  uint8_t* areg = psm + C::OFF_A;
  uint8_t* lo_area = psm + C::OFF_LO;
  float* cst = reinterpret_cast<float*>(psm + C::OFF_C);
  float* out_s = reinterpret_cast<float*>(psm + C::OFF_OUT);
  float* comp_s = reinterpret_cast<float*>(psm + C::OFF_COMP);      // [comp_G * 128][4] raw values of the current ray group
  uint64_t* full = reinterpret_cast<uint64_t*>(psm + C::OFF_BAR);
  uint64_t* empty = full + C::NSLOT;
  uint64_t* accf = empty + C::NSLOT;      // [2]
  uint64_t* aq = accf + 2;                // [4]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(aq + 4);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    for (int i = 0; i < C::NSLOT; ++i) { tc::mbar_init(&full[i], 1); tc::mbar_init(&empty[i], 1); }
    tc::mbar_init(&accf[0], 1); tc::mbar_init(&accf[1], 1);
    for (int i = 0; i < 4; ++i) tc::mbar_init(&aq[i], 256);
    tc::fence_mbar_init();
  }
  for (int i = tid; i < C_TOTAL; i += blockDim.x) cst[i] = a.cbuf[i];
  // constant ONES slab: [128 rows x 16 k], k == 0 -> 1.0 (bf16 0x3F80), else 0
  for (int i = tid; i < 4096 / 16; i += blockDim.x) {
    const bool k0chunk = i < 128;
    *reinterpret_cast<uint4*>(areg + L::ONES + i * 16) = make_uint4(k0chunk ? 0x00003F80u : 0u, 0u, 0u, 0u);
  }
  tc::fence_proxy_async();
  __syncthreads();
  if (warp == 1) tc::tmem_alloc(tmem_slot, 512);
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  PCtx ctx;
  ctx.e.ring_addr = tc::smem_u32(psm + C::OFF_RING); ctx.e.full_addr = tc::smem_u32(full); ctx.e.empty_addr = tc::smem_u32(empty);
  ctx.e.acc_full_addr = 0; ctx.e.a_ready_addr = 0;
  ctx.e.tmem_acc = tmem; ctx.e.tmem_ahi = tmem + 256; ctx.e.tmem_alo = 0; ctx.e.smem_a = tc::smem_u32(areg);
  ctx.e.dbg = nullptr; ctx.e.dbg_tiles = 0;
  ctx.e.pol_keep = tc::policy_evict_last(); ctx.e.pol_stream = tc::policy_evict_first();
  ctx.accf_addr = tc::smem_u32(accf); ctx.aq_addr = tc::smem_u32(aq);
  ctx.smem_lo = tc::smem_u32(lo_area);
  ctx.dbg = a.dbg; ctx.dbg_tiles = a.dbg_tiles;

  int tile_first, tile_count, tile_stride;
  fused::cta_tiles(a, tile_first, tile_count, tile_stride);
  // (one loop form for both tile mappings: a second instantiation of the 336-step unrolled issue code made ptxas spill
  //  inside the MMA-issuing thread — 60 -> 1972 bytes — and cost the training forward 1 ms per step)
  if (warp == 0) {
    if (lane == 0) eng::producer_loop_n<C>(ctx.e, a.wimg, tile_count);
  } else if (warp == 1) {
    if (lane == 0) {
      mma_loop_r<C>(ctx, tile_first, tile_first + tile_count * tile_stride, tile_stride);
    }
  } else {
    // ===================== epilogue: 8 warps, 2 per TMEM lane quadrant =============================
    const int quad = warp & 3, half = (warp - 2) >> 2;
    const int row = quad * 32 + lane;
    const uint32_t lane_base = (uint32_t)(quad * 32) << 16;
    for (int tile_iter = 0; tile_iter < tile_count; ++tile_iter) {
      const int tile = tile_first + tile_iter * tile_stride;
      const int64_t p = (int64_t)tile * TILE_M + row;
      const bool valid = p < a.P;
      float* dpe = a.dump_pe ? a.dump_pe + p * a.dump_pe_ld : nullptr;
      float* dped = a.dump_ped ? a.dump_ped + p * a.dump_ped_ld : nullptr;
      float e[32];
      // ---- positional encodings -> smem A slabs (same split of the work between the two warps as the serial kernel)
      if constexpr (XS == 6) {
        float x[4] = {0.f, 0.f, 0.f, 0.f}, vd[3] = {0.f, 0.f, 0.f};
        if (valid) {
          const int64_t r = p / a.S;
          const float4 q = *reinterpret_cast<const float4*>(a.pts + p * 4);
          x[0] = q.x; x[1] = q.y; x[2] = q.z; x[3] = q.w;
#pragma unroll
          for (int cc = 0; cc < 3; ++cc) vd[cc] = a.viewdirs[r * 3 + cc];
        }
        if (half == 0) {
          fused::pe_fill32<10, 0, 4>(x, valid, e);
          fused::pe_store32<SPLIT>(e, 0, 84, areg + L::XHI, areg + L::XLO, row, a.img_x, tile, dpe, valid, ctx.e.pol_stream);
          fused::pe_fill32<4, 0>(vd, valid, e);
          fused::pe_store32<SPLIT>(e, 0, 27, areg + L::VHI, areg + L::VLO, row, a.img_v, tile, dped, valid, ctx.e.pol_stream);
          *reinterpret_cast<float4*>(out_s + row * 4) = make_float4(0.f, 0.f, 0.f, 0.f);
        } else {
          fused::pe_fill32<10, 32, 4>(x, valid, e);
          fused::pe_store32<SPLIT>(e, 32, 84, areg + L::XHI, areg + L::XLO, row, a.img_x, tile, dpe, valid, ctx.e.pol_stream);
          fused::pe_fill32<10, 64, 4>(x, valid, e);
          fused::pe_store32<SPLIT>(e, 64, 84, areg + L::XHI, areg + L::XLO, row, a.img_x, tile, dpe, valid, ctx.e.pol_stream);
        }
      } else {
        float x[3] = {0.f, 0.f, 0.f}, vd[3] = {0.f, 0.f, 0.f};
        if (valid) {
          const int64_t r = p / a.S;
          if (a.rays) {
            const float* ry = a.rays + r * a.ray_cols;
            const float zz = a.z[p];
#pragma unroll
            for (int cc = 0; cc < 3; ++cc) {
              x[cc] = __fadd_rn(ry[cc], __fmul_rn(ry[3 + cc], zz));   // render.py:259
              vd[cc] = ry[8 + cc];
            }
          } else {
#pragma unroll
            for (int cc = 0; cc < 3; ++cc) { x[cc] = a.pts[p * 3 + cc]; vd[cc] = a.viewdirs[r * 3 + cc]; }
          }
        }
        if (half == 0) {
          fused::pe_fill32<10, 0>(x, valid, e);
          fused::pe_store32<SPLIT>(e, 0, 63, areg + L::XHI, areg + L::XLO, row, a.img_x, tile, dpe, valid, ctx.e.pol_stream);
          fused::pe_fill32<4, 0>(vd, valid, e);
          fused::pe_store32<SPLIT>(e, 0, 27, areg + L::VHI, areg + L::VLO, row, a.img_v, tile, dped, valid, ctx.e.pol_stream);
          *reinterpret_cast<float4*>(out_s + row * 4) = make_float4(0.f, 0.f, 0.f, 0.f);
        } else {
          fused::pe_fill32<10, 32>(x, valid, e);
          fused::pe_store32<SPLIT>(e, 32, 63, areg + L::XHI, areg + L::XLO, row, a.img_x, tile, dpe, valid, ctx.e.pol_stream);
        }
      }
      tc::fence_proxy_async();
      tc::tc_fence_before();        // the previous tile's accumulator loads are complete (ordered before the arrive)
      for (int i = 0; i < 4; ++i) tc::mbar_arrive(&aq[i]);
      float alpha = 0.f, rgb[3] = {0.f, 0.f, 0.f};
      if constexpr (ROLL >= 1) {
#pragma unroll 1
        for (int S = 0; S < 7; ++S) {
          epi_half_rt<NSPLIT, 0>(a, ctx, lo_area, lane_base, half, row, tile, p, valid, tile_iter, S);
          epi_half_rt<NSPLIT, 1>(a, ctx, lo_area, lane_base, half, row, tile, p, valid, tile_iter, S);
        }
        epi_half<NSPLIT, 7, 0>(a, cst, ctx, lo_area, lane_base, half, row, tile, p, valid, alpha, rgb, tile_iter);
        epi_half<NSPLIT, 7, 1>(a, cst, ctx, lo_area, lane_base, half, row, tile, p, valid, alpha, rgb, tile_iter);
        epi_half<NSPLIT, 8, 0>(a, cst, ctx, lo_area, lane_base, half, row, tile, p, valid, alpha, rgb, tile_iter);
        epi_half<NSPLIT, 8, 1>(a, cst, ctx, lo_area, lane_base, half, row, tile, p, valid, alpha, rgb, tile_iter);
        epi_half<NSPLIT, 9, 0>(a, cst, ctx, lo_area, lane_base, half, row, tile, p, valid, alpha, rgb, tile_iter);
        epi_half<NSPLIT, 9, 1>(a, cst, ctx, lo_area, lane_base, half, row, tile, p, valid, alpha, rgb, tile_iter);
      } else {
        epi_tile<NSPLIT>(a, cst, ctx, lo_area, lane_base, half, row, tile, p, valid, alpha, rgb, tile_iter,
                         std::make_index_sequence<NSTAGE>{});
      }
      // combine the two column-halves of each row: both add into smem, half 0 finishes
      atomicAdd(out_s + row * 4 + 0, rgb[0]);
      atomicAdd(out_s + row * 4 + 1, rgb[1]);
      atomicAdd(out_s + row * 4 + 2, rgb[2]);
      atomicAdd(out_s + row * 4 + 3, alpha);
      tc::tc_fence_before();
      asm volatile("bar.sync 1, 256;" ::: "memory");
      if (half == 0) {
        const float4 o = *reinterpret_cast<const float4*>(out_s + row * 4);
        const float4 r4 = make_float4(o.x + cst[C_SCAL + 1], o.y + cst[C_SCAL + 2], o.z + cst[C_SCAL + 3], o.w + cst[C_SCAL]);
        if (valid && a.raw != nullptr) *reinterpret_cast<float4*>(a.raw + p * 4) = r4;
        if (a.comp_on) *reinterpret_cast<float4*>(comp_s + ((tile_iter % a.comp_G) * TILE_M + row) * 4) = r4;
      }
      asm volatile("bar.sync 1, 256;" ::: "memory");   // out_s is re-zeroed by the next tile's prologue
      if (a.comp_on && ((tile_iter % a.comp_G) == a.comp_G - 1 || tile_iter == tile_count - 1)) {
        // the group's rays are complete: alpha-composite them from shared memory, one warp per ray (render.py:302-355)
        const int g_first = tile - (tile_iter % a.comp_G);                  // first tile of this group
        const int64_t p0 = (int64_t)g_first * TILE_M;
        const int64_t p1 = (int64_t)(tile + 1) * TILE_M < a.P ? (int64_t)(tile + 1) * TILE_M : a.P;
        const int n_rays = (int)((p1 - p0) / a.S);
        const int64_t ray0 = p0 / a.S;
        composite_group(a.comp, comp_s, ray0, n_rays, a.S, warp - 2, lane);
        asm volatile("bar.sync 1, 256;" ::: "memory");   // comp_s is rewritten by the next group
      }
    }
  }
  tc::tc_fence_before();
  __syncthreads();
#ifdef SCNERF_TIMELINE
  // per-CTA end stamps behind the [tile][stage][16] block: how evenly does the static round-robin tile split end?
  if (a.dbg != nullptr && tid == 0) {
    unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    a.dbg[(size_t)a.dbg_tiles * NSTAGE * 16 + blockIdx.x * 2] = (long long)t;
    a.dbg[(size_t)a.dbg_tiles * NSTAGE * 16 + blockIdx.x * 2 + 1] = tile_count;
  }
#endif
  if (warp == 1) tc::tmem_dealloc(tmem, 512);
}

__device__ eng::Plan d_plan_pipe;        // runtime copies of the constexpr plans, for the pack kernel
__device__ PlanSrc d_plansrc_pipe;
__device__ eng::Plan d_plan_pipe6;
__device__ PlanSrc d_plansrc_pipe6;
template <int NSPLIT, int XS = 4>
__global__ void __launch_bounds__(256) pack_pipe_kernel(fused::PackSrc src, uint8_t* __restrict__ img) {
  const int i = blockIdx.y;
  const eng::Plan& P = XS == 4 ? d_plan_pipe : d_plan_pipe6;
  const PlanSrc& S = XS == 4 ? d_plansrc_pipe : d_plansrc_pipe6;
  if (i < P.n_slabs) fused::pack_slab_impl<NSPLIT>(P.slab[i], S.s[i], src, img);
}

}  // namespace fpipe
}  // namespace scnerf
