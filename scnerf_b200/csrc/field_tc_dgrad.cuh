// Fused field backward, data-gradient chain, on tcgen05 (sm_100a).
//
// For each 128-sample tile:  d(raw)[4] -> rgb/alpha heads (fp32 in registers) -> views layer ->
// feature layer -> trunk 7..0, every  g_in = dZ * W  GEMM on the tensor cores with the SAME slab
// engine as the forward (A = dZ in TMEM, B = transposed bf16 weight slabs streamed by bulk TMA),
// ReLU masks taken from the forward's 1-bit-per-activation mask buffer, and every dZ written back as a bf16 (hi[,lo])
// TILE IMAGE that the wgrad kernel consumes directly.  Also produces d(pts) and d(viewdirs) per
// sample (PE backward in registers), reduced per ray by reduce_pts_grad_kernel.
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
constexpr int NSTAGE = 11;   // S1, S2, S3, S4, S5a, S5b, S6, S7, S8, S9, S10
constexpr int ACC2_COL = 320;

// one k16 slab of a transposed-weight GEMM stage
struct SlabSpec { int N, acc_col, j, first, last, stage_begin, wsel, col0, valid_n, stage; };
// enumerate the slab sequence (shared by the constexpr plan and the host-side pack table)
template <int XN = 64, class F>
__host__ __device__ constexpr void for_each_slab(F&& f) {
  constexpr int IN_CH = XN == 64 ? 63 : 84;
  // S1 (stage 0): [g_feat | g_V] = dZ_v (K=128) * W_views
  for (int j = 0; j < 8; ++j) {
    f(SlabSpec{256, 0, j, j == 0, 0, j == 0, 9, 0, 256, 0});
    f(SlabSpec{32, ACC2_COL, j, j == 0, j == 7, 0, 9, 256, 27, 0});
  }
  // wsel, N, col0, valid_n for stages 1..10
  const int spec[10][4] = {{8, 256, 0, 256},   // S2 : g_h7 = g_feat * W_feature
                           {7, 256, 0, 256},   // S3 : g_h6 = dZ7 * W7
                           {6, 256, 0, 256},   // S4 : g_h5 = dZ6 * W6
                           {5, XN, 0, IN_CH},  // S5a: g_X (skip branch) = dZ5 * W5[:, :63]
                           {5, 256, IN_CH, 256},  // S5b: g_h4 = dZ5 * W5[:, 63:]
                           {4, 256, 0, 256},   // S6 : g_h3
                           {3, 256, 0, 256},   // S7 : g_h2
                           {2, 256, 0, 256},   // S8 : g_h1
                           {1, 256, 0, 256},   // S9 : g_h0
                           {0, XN, 0, IN_CH}}; // S10: g_X (layer 0) = dZ0 * W0
  for (int t = 0; t < 10; ++t)
    for (int j = 0; j < 16; ++j)
      f(SlabSpec{spec[t][1], 0, j, j == 0, j == 15, j == 0, spec[t][0], spec[t][2], spec[t][3], t + 1});
}
struct PlanFiller {
  eng::Plan P{};
  int n = 0;
  uint32_t off = 0;
  __host__ __device__ constexpr void operator()(const SlabSpec& q) {
    eng::SlabDef e{};
    e.n = (uint16_t)q.N; e.acc_col = (uint16_t)q.acc_col; e.a_kind = eng::A_TMEM; e.a_off = (uint16_t)(q.j * 8);
    e.stage = (uint8_t)q.stage; e.img_off = off;
    e.flags = (uint8_t)((q.first ? eng::F_ZERO_ACC : 0) | (q.last ? eng::F_STAGE_END : 0) |
                        (q.stage_begin ? eng::F_STAGE_BEGIN : 0));
    P.slab[n++] = e;
    off += (uint32_t)q.N * 32u;
  }
};
template <int XN = 64>
__host__ __device__ constexpr eng::Plan make_plan() {
  PlanFiller f{};
  for_each_slab<XN>(f);
  f.P.n_slabs = f.n; f.P.n_stages = NSTAGE;
  return f.P;
}
template <int XN = 64>
inline void build_plansrc(PlanSrc& S) {
  int n = 0;
  for_each_slab<XN>([&](const SlabSpec& q) {
    SrcDef d{};
    d.wsel = (uint8_t)q.wsel; d.kind = 1; d.row0 = (uint16_t)(16 * q.j); d.col0 = (uint16_t)q.col0;
    d.valid_k = 16; d.valid_n = (uint16_t)q.valid_n;
    S.s[n++] = d;
  });
}

__device__ eng::Plan d_plan_dgrad;
__device__ PlanSrc d_plansrc_dgrad;
__device__ eng::Plan d_plan_dgrad96;      // 4-D points: 96-wide d(PE) stages
__device__ PlanSrc d_plansrc_dgrad96;
template <int NSPLIT, int XN = 64>
__global__ void __launch_bounds__(256) pack_dgrad_kernel(fused::PackSrc src, uint8_t* __restrict__ img) {
  const int i = blockIdx.y;
  const eng::Plan& P = XN == 64 ? d_plan_dgrad : d_plan_dgrad96;
  const PlanSrc& S = XN == 64 ? d_plansrc_dgrad : d_plansrc_dgrad96;
  if (i < P.n_slabs) fused::pack_slab_impl<NSPLIT>(P.slab[i], S.s[i], src, img);
}

template <int NSPLIT_, int XN_ = 64> struct Cfg {
  static constexpr int NSPLIT = NSPLIT_;
  static constexpr int XN = XN_;
  static constexpr eng::Plan PLAN = make_plan<XN_>();
  static constexpr int GROUP = NSPLIT_ == 1 ? 2 : 1;          // slabs per ring slot
  static constexpr int NSLOT = XN_ == 64 ? 11 : 8;            // 176 slabs per tile = 88 pairs = 11*8 = 11*16 = 8*22
                                                              // (XN = 96: the [96][128] fp32 d(PE) buffer needs the room)
  static constexpr int SLOT_BYTES = 16384;
  static_assert((PLAN.n_slabs / GROUP) % NSLOT == 0 && PLAN.n_slabs % GROUP == 0, "ring size must divide the slab-group count");
  static constexpr int OFF_RING = 0;
  static constexpr int OFF_C = NSLOT * SLOT_BYTES;
  static constexpr int OFF_GX = OFF_C + ((fused::C_TOTAL * 4 + 127) / 128) * 128;   // [64][128] fp32 skip-branch d(PE)
  static constexpr int OFF_OUT = OFF_GX + XN_ * 128 * 4;                              // [128][4]
  static constexpr int OFF_BAR = OFF_OUT + 128 * 4 * 4;
  static constexpr int SMEM_BYTES = OFF_BAR + (2 * NSLOT + 2) * 8 + 16;
  static_assert(SMEM_BYTES <= 232448, "exceeds the 227 KB shared-memory limit");
};

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

template <int NSPLIT, int XN = 64>
__global__ void __launch_bounds__(320, 1) field_fused_dgrad_kernel(const __grid_constant__ Args a) {
  using C = Cfg<NSPLIT, XN>;
  constexpr bool SPLIT = NSPLIT == 3;
  extern __shared__ __align__(128) uint8_t dsm[];
  uint8_t* ringp = dsm + C::OFF_RING;
  float* cst = reinterpret_cast<float*>(dsm + C::OFF_C);
  float* gx_s = reinterpret_cast<float*>(dsm + C::OFF_GX);
  float* out_s = reinterpret_cast<float*>(dsm + C::OFF_OUT);
  uint64_t* full = reinterpret_cast<uint64_t*>(dsm + C::OFF_BAR);
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
  for (int i = tid; i < fused::C_TOTAL; i += blockDim.x) cst[i] = a.cbuf[i];
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
  ctx.tmem_acc = T_ACC; ctx.tmem_ahi = T_AHI; ctx.tmem_alo = T_ALO; ctx.smem_a = 0;
  ctx.dbg = nullptr; ctx.dbg_tiles = 0;
  ctx.pol_keep = tc::policy_evict_last(); ctx.pol_stream = tc::policy_evict_first();

  if (warp == 0) {
    if (lane == 0) eng::producer_loop<C>(ctx, a.wimg, a.num_tiles);
  } else if (warp == 1) {
    if (lane == 0) eng::mma_loop<C>(ctx, a.num_tiles);
  } else {
    const int quad = warp & 3, half = (warp - 2) >> 2;
    const int row = quad * 32 + lane;
    const uint32_t lane_base = (uint32_t)(quad * 32) << 16;
    uint32_t m = 0;
    for (int tile = blockIdx.x; tile < a.num_tiles; tile += gridDim.x) {
      const int64_t p = (int64_t)tile * TILE_M + row;
      const bool valid = p < a.P;
      float4 gr = make_float4(0.f, 0.f, 0.f, 0.f);
      if (valid) gr = *reinterpret_cast<const float4*>(a.g_raw + p * 4);
      // ---- E0: rgb head dgrad (fp32), ReLU mask of the view layer -> dZ_v (this warp's 64 columns)
      {
        if (half == 0) *reinterpret_cast<float4*>(out_s + row * 4) = make_float4(0.f, 0.f, 0.f, 0.f);
        const uint4 mv = tc::ld_v4_hint(a.relu_bits + ((size_t)(tile * 9 + 8) * 2 + half) * 128 + row, ctx.pol_stream);
        const uint32_t mvw[2] = {mv.x, mv.y};
#pragma unroll
        for (int cc = 0; cc < 2; ++cc) {
          const int c0 = half * 64 + cc * 32;
          float f[32];
#pragma unroll
          for (int j = 0; j < 32; ++j)
            f[j] = gr.x * cst[fused::C_WRGB + c0 + j] + gr.y * cst[fused::C_WRGB + 128 + c0 + j] +
                   gr.z * cst[fused::C_WRGB + 256 + c0 + j];
#pragma unroll
          for (int j = 0; j < 32; ++j) f[j] = eng::relu_bit(mvw[cc], j) ? f[j] : 0.f;
          uint32_t hi[16], lo[16];
          eng::split32<SPLIT, false>(f, hi, lo);
          tc::tmem_st16(T_AHI + lane_base + (uint32_t)(c0 >> 1), hi);
          if (SPLIT) tc::tmem_st16(T_ALO + lane_base + (uint32_t)(c0 >> 1), lo);
        }
        tc::tmem_st_wait();
        tc::tc_fence_before();
        tc::mbar_arrive(a_ready);
      }
#pragma unroll 1
      for (int s = 0; s < NSTAGE; ++s, ++m) {
        // stage -> (mask image, output image); by value: taking the address of a kernel parameter would
        // spill the whole struct to local memory
        // stage -> trunk layer whose ReLU is differentiated here (-1: none: feature_linear / d(PE) stages)
        const int mlayer = (s >= 1 && s <= 3) ? 8 - s : ((s >= 5 && s <= 9) ? 9 - s : -1);
        // this stage's A operand (TMEM) is a dZ the wgrad kernel needs: write its tile image now, under
        // the MMA phase (off the critical path)
        switch (s) {
          case 0: eng::dump_from_tmem<SPLIT, 2>(a.out_dzv, tile, row, T_AHI, T_ALO, lane_base, half * 64, ctx.pol_stream); break;
          case 1: eng::dump_from_tmem<SPLIT, 4>(a.out_dfeat, tile, row, T_AHI, T_ALO, lane_base, half * 128, ctx.pol_stream); break;
          case 2: eng::dump_from_tmem<SPLIT, 4>(a.out_dz[7], tile, row, T_AHI, T_ALO, lane_base, half * 128, ctx.pol_stream); break;
          case 3: eng::dump_from_tmem<SPLIT, 4>(a.out_dz[6], tile, row, T_AHI, T_ALO, lane_base, half * 128, ctx.pol_stream); break;
          case 4: eng::dump_from_tmem<SPLIT, 4>(a.out_dz[5], tile, row, T_AHI, T_ALO, lane_base, half * 128, ctx.pol_stream); break;
          case 6: eng::dump_from_tmem<SPLIT, 4>(a.out_dz[4], tile, row, T_AHI, T_ALO, lane_base, half * 128, ctx.pol_stream); break;
          case 7: eng::dump_from_tmem<SPLIT, 4>(a.out_dz[3], tile, row, T_AHI, T_ALO, lane_base, half * 128, ctx.pol_stream); break;
          case 8: eng::dump_from_tmem<SPLIT, 4>(a.out_dz[2], tile, row, T_AHI, T_ALO, lane_base, half * 128, ctx.pol_stream); break;
          case 9: eng::dump_from_tmem<SPLIT, 4>(a.out_dz[1], tile, row, T_AHI, T_ALO, lane_base, half * 128, ctx.pol_stream); break;
          case 10: eng::dump_from_tmem<SPLIT, 4>(a.out_dz[0], tile, row, T_AHI, T_ALO, lane_base, half * 128, ctx.pol_stream); break;
          default: break;   // 5 (S5b): A is still dZ5
        }
        // ReLU masks of this warp's 128 columns are fetched BEFORE waiting for the accumulator, i.e.
        // under the MMA phase of this stage
        uint32_t mk[4] = {0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu};
        if (mlayer >= 0) {
          const uint4 mb = tc::ld_v4_hint(a.relu_bits + ((size_t)(tile * 9 + mlayer) * 2 + half) * 128 + row, ctx.pol_stream);
          mk[0] = mb.x; mk[1] = mb.y; mk[2] = mb.z; mk[3] = mb.w;
        }
        tc::mbar_wait(acc_full, m & 1);
        tc::tc_fence_after();
        if constexpr (XN == 96) {
          if (s == 4 || s == 10) {
            // ---- 96-wide d(PE) stages (4-D points): this warp owns columns [32 half, +32) and [64 + 16 half, +16)
            uint32_t va[32], vb[16];
            const int ca = half * 32, cb = 64 + half * 16;
            tc::tmem_ld32(T_ACC + lane_base + ca, va);
            tc::tmem_ld16(T_ACC + lane_base + cb, vb);
            tc::tmem_ld_wait();
            if (s == 4) {
#pragma unroll
              for (int j = 0; j < 32; ++j) gx_s[(ca + j) * TILE_M + row] = __uint_as_float(va[j]);
#pragma unroll
              for (int j = 0; j < 16; ++j) gx_s[(cb + j) * TILE_M + row] = __uint_as_float(vb[j]);
              tc::tc_fence_before();
              tc::mbar_arrive(a_ready);
            } else {
              float x[4] = {0.f, 0.f, 0.f, 0.f};
              if (valid) {
                const float4 q = *reinterpret_cast<const float4*>(a.pts + p * 4);
                x[0] = q.x; x[1] = q.y; x[2] = q.z; x[3] = q.w;
              }
              float gx[4] = {0.f, 0.f, 0.f, 0.f};
              auto contract = [&](int i, float gacc) {      // PE column i of [x(4), sin(2^f x)(4), cos(2^f x)(4), ...]
                const float g = gacc + gx_s[i * TILE_M + row];
                if (i < 4) gx[i] += g;
                else if (i < 84) {
                  const int f = (i - 4) >> 3, r = (i - 4) & 7, c = r & 3;
                  const float fr = (float)(1 << f);
                  float sv, cv;
                  fused::sincos_cw(x[c] * fr, sv, cv);     // same evaluation as the forward's PE
                  gx[c] += (r < 4) ? fr * cv * g : -fr * sv * g;
                }
              };
#pragma unroll
              for (int j = 0; j < 32; ++j) contract(ca + j, __uint_as_float(va[j]));
#pragma unroll
              for (int j = 0; j < 16; ++j) contract(cb + j, __uint_as_float(vb[j]));
#pragma unroll
              for (int c = 0; c < 4; ++c) atomicAdd(out_s + row * 4 + c, gx[c]);
            }
            continue;
          }
        }
        if (s == 4 || s == 10) {
          // ---- 64-wide d(PE) stages: this warp owns 32 columns
          uint32_t v[32];
          const int c0 = half * 32;
          tc::tmem_ld32(T_ACC + lane_base + c0, v);
          tc::tmem_ld_wait();
          if (s == 4) {
#pragma unroll
            for (int j = 0; j < 32; ++j) gx_s[(c0 + j) * TILE_M + row] = __uint_as_float(v[j]);
            tc::tc_fence_before();
            tc::mbar_arrive(a_ready);
          } else {
            // total d(PE row) = layer-0 share + skip share; contract with dPE/dx (Embedder backward)
            float x[3] = {0.f, 0.f, 0.f};
            if (valid) {
              const float* ry = a.rays + (p / a.S) * a.ray_cols;
              const float zz = a.z[p];
#pragma unroll
              for (int c = 0; c < 3; ++c) x[c] = __fadd_rn(ry[c], __fmul_rn(ry[3 + c], zz));
            }
            float gx[3] = {0.f, 0.f, 0.f};
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              const int i = c0 + j;         // PE column (compile-time after unrolling for each half)
              const float g = __uint_as_float(v[j]) + gx_s[i * TILE_M + row];
              if (i < 3) gx[i] += g;
              else if (i < 63) {
                const int f = (i - 3) / 6, r = (i - 3) % 6, c = r % 3;
                const float fr = (float)(1 << f);
                float sv, cv;
                fused::sincos_cw(x[c] * fr, sv, cv);       // same evaluation as the forward's PE
                gx[c] += (r < 3) ? fr * cv * g : -fr * sv * g;
              }
            }
            atomicAdd(out_s + row * 4 + 0, gx[0]);
            atomicAdd(out_s + row * 4 + 1, gx[1]);
            atomicAdd(out_s + row * 4 + 2, gx[2]);
          }
          continue;
        }
        // ---- 256-wide stages: this warp owns 128 columns = 4 chunks
        if (s == 0 && half == 1) {
          // d(PE(dir)) lives in ACC2: read it before this warp's stores overwrite those TMEM columns
          uint32_t v[32];
          tc::tmem_ld32(T_ACC + lane_base + ACC2_COL, v);
          tc::tmem_ld_wait();
          float vd[3] = {0.f, 0.f, 0.f}, gv[3] = {0.f, 0.f, 0.f};
          if (valid) {
            if (XN == 96) {
              const float* vv = a.viewdirs + (p / a.S) * 3;
              vd[0] = vv[0]; vd[1] = vv[1]; vd[2] = vv[2];
            } else {
              const float* ry = a.rays + (p / a.S) * a.ray_cols;
              vd[0] = ry[8]; vd[1] = ry[9]; vd[2] = ry[10];
            }
          }
#pragma unroll
          for (int i = 0; i < 27; ++i) {
            const float g = __uint_as_float(v[i]);
            if (i < 3) gv[i] += g;
            else {
              const int f = (i - 3) / 6, r = (i - 3) % 6, c = r % 3;
              const float fr = (float)(1 << f);
              float sv, cv;
              fused::sincos_cw(vd[c] * fr, sv, cv);
              gv[c] += (r < 3) ? fr * cv * g : -fr * sv * g;
            }
          }
          if (valid) { a.g_vd[p * 3] = gv[0]; a.g_vd[p * 3 + 1] = gv[1]; a.g_vd[p * 3 + 2] = gv[2]; }
        }
        uint32_t v[4][32];
#pragma unroll
        for (int cc = 0; cc < 4; ++cc) tc::tmem_ld32(T_ACC + lane_base + half * 128 + cc * 32, v[cc]);
        tc::tmem_ld_wait();
#pragma unroll
        for (int cc = 0; cc < 4; ++cc) {
          const int cu = half * 128 + cc * 32;
          float f[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[cc][j]);
          if (s == 1) {   // alpha head: g_h7 += g_alpha * w_alpha
#pragma unroll
            for (int j = 0; j < 32; ++j) f[j] = fmaf(gr.w, cst[fused::C_WALPHA + cu + j], f[j]);
          }
#pragma unroll
          for (int j = 0; j < 32; ++j) f[j] = eng::relu_bit(mk[cc], j) ? f[j] : 0.f;
          uint32_t hi[16], lo[16];
          eng::split32<SPLIT, false>(f, hi, lo);
          tc::tmem_st16(T_AHI + lane_base + (uint32_t)(cu >> 1), hi);
          if (SPLIT) tc::tmem_st16(T_ALO + lane_base + (uint32_t)(cu >> 1), lo);
        }
        tc::tmem_st_wait();
        tc::tc_fence_before();
        tc::mbar_arrive(a_ready);
      }
      tc::tc_fence_before();
      asm volatile("bar.sync 1, 256;" ::: "memory");
      if (half == 0 && valid) {
        if (XN == 96) *reinterpret_cast<float4*>(a.g_pts + p * 4) = *reinterpret_cast<const float4*>(out_s + row * 4);
        else { a.g_pts[p * 3] = out_s[row * 4]; a.g_pts[p * 3 + 1] = out_s[row * 4 + 1]; a.g_pts[p * 3 + 2] = out_s[row * 4 + 2]; }
      }
      asm volatile("bar.sync 1, 256;" ::: "memory");
    }
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 1) tc::tmem_dealloc(tmem, 512);
}

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
