// Fused field forward on tcgen05 (sm_100a): positional encoding -> 8x256 trunk (skip at 4) ->
// alpha head + feature layer -> view branch -> rgb head, for one 128-sample tile at a time, with
// NO intermediate activation ever leaving the SM:
//
//   HBM in : 44 B/ray-sample of geometry (rays, z; L2-resident) ; HBM out: raw[4] = 16 B/sample
//   weights: bf16 (hi[,lo]) slabs streamed L2 -> smem by 1-D bulk TMA through an mbarrier ring
//   A operand (activations): TMEM (TS-mode MMA) for the 256-wide hidden state, smem for the
//                            PE(pts) (K=64) and PE(dir) (K=32) slabs
//   accumulator: TMEM, 128 lanes x 256 fp32 columns
//
// Warp roles (192 threads):  warp 0 = TMA producer, warp 1 = MMA issuer (+TMEM alloc),
// warps 2..5 = epilogue (TMEM lane quadrant = warp & 3): bias, ReLU, bf16 (hi/lo) split, tcgen05.st
// back into TMEM as the next layer's A operand; alpha / rgb heads are fp32 dot products in the
// epilogue registers.
//
// NSPLIT = 1: single-pass bf16 (fp32 accumulate).   NSPLIT = 3: split-bf16, A_hi*B_hi + A_lo*B_hi +
// A_hi*B_lo — ~16 mantissa bits per operand, which is what the 1e-4 parity gate needs.
//
// Algorithmic work: 593,408 MAC/sample; tensor pipe executes 593,920 MAC/sample
// (K padded 63->64, 27->32) x NSPLIT.   Reference: NeRF/run_nerf_helpers.py:24-72,105-128 and
// NeRF/create_nerf.py:18-32.
#pragma once
#include "common.cuh"
#include "tc_prims.cuh"

namespace scnerf {
namespace fused {

constexpr int TILE_M = 128;
constexpr int NSTAGE = 10;
// per stage: output width N, k16 slabs taken from X (PE pts, smem), H (hidden, TMEM), V (PE dir, smem)
struct StageDef { int N, kx, kh, kv, relu; };
__host__ __device__ constexpr StageDef stage_def(int s) {
  return s == 0 ? StageDef{256, 4, 0, 0, 1}
       : s == 5 ? StageDef{256, 4, 16, 0, 1}
       : s == 8 ? StageDef{256, 0, 16, 0, 0}      // feature_linear: no activation
       : s == 9 ? StageDef{128, 0, 16, 2, 1}      // views_linears[0]
                : StageDef{256, 0, 16, 0, 1};
}
__host__ __device__ constexpr int stage_k16(int s) {
  return stage_def(s).kx + stage_def(s).kh + stage_def(s).kv;
}
__host__ __device__ constexpr int slab_bytes(int s) { return stage_def(s).N * 32; }  // one k16 slab, one half

// A ring slot carries 16 KB (N=256) / 8 KB (N=128): NSPLIT==1 -> two consecutive k16 slabs (K=32),
// NSPLIT==3 -> the hi and lo slab of one k16.
template <int NSPLIT> __host__ __device__ constexpr int slots_in_stage(int s) {
  return NSPLIT == 1 ? stage_k16(s) / 2 : stage_k16(s);
}
template <int NSPLIT> __host__ __device__ constexpr int slots_per_tile() {
  int n = 0;
  for (int s = 0; s < NSTAGE; ++s) n += slots_in_stage<NSPLIT>(s);
  return n;
}
template <int NSPLIT> __host__ __device__ constexpr size_t weight_image_bytes() {
  size_t n = 0;
  for (int s = 0; s < NSTAGE; ++s) n += (size_t)slots_in_stage<NSPLIT>(s) * slab_bytes(s) * 2;
  return n;
}

// constants block (fp32) kept in shared memory
constexpr int C_BIAS = 0;                 // 8 x 256 trunk biases
constexpr int C_BFEAT = 8 * 256;          // 256
constexpr int C_BVIEW = C_BFEAT + 256;    // 128
constexpr int C_WALPHA = C_BVIEW + 128;   // 256
constexpr int C_WRGB = C_WALPHA + 256;    // 3 x 128
constexpr int C_SCAL = C_WRGB + 384;      // b_alpha, b_rgb[3]
constexpr int C_TOTAL = C_SCAL + 8;       // 3080 floats

template <int NSPLIT> struct Cfg {
  static constexpr int NSLOT = NSPLIT == 1 ? 10 : 9;
  static constexpr int SLOT_BYTES = 16384;
  static constexpr int X_BYTES = 16384;   // 128 rows x 64 k bf16
  static constexpr int V_BYTES = 8192;    // 128 rows x 32 k bf16
  static constexpr int NHALF = NSPLIT == 1 ? 1 : 2;
  static constexpr int OFF_RING = 0;
  static constexpr int OFF_X = NSLOT * SLOT_BYTES;
  static constexpr int OFF_V = OFF_X + X_BYTES * NHALF;
  static constexpr int OFF_C = OFF_V + V_BYTES * NHALF;
  static constexpr int OFF_BAR = OFF_C + ((C_TOTAL * 4 + 127) / 128) * 128;
  static constexpr int SMEM_BYTES = OFF_BAR + (2 * NSLOT + 2) * 8 + 16 + 1024;  // + alignment slack
};

struct Args {
  const float* rays; int ray_cols;     // [N, ray_cols] (o, d, near, far, viewdirs) or NULL
  const float* z;                      // [N, S] depths (with rays)
  const float* pts;                    // [N, S, 3] explicit points (rays == NULL)
  const float* viewdirs;               // [N, 3]    explicit directions (rays == NULL)
  int64_t P; int S;
  const uint8_t* wimg;                 // packed weight image (see pack_weights_kernel)
  const float* cbuf;                   // packed constants, C_TOTAL floats
  float* raw;                          // [P, 4]
  int num_tiles;
  // training mode: fp32 copies of every layer input for the backward (NULL = inference)
  float* dump[NSTAGE]; int dump_ld[NSTAGE];   // post-activation output of stage s
  float* dump_pe; int dump_pe_ld;             // PE(pts)  [P, 63]
  float* dump_ped; int dump_ped_ld;           // PE(dir)  [P, 27]
};

// ---- weight packing: fp32 nn.Linear tensors -> slab image in MMA-issue order ----------------------
struct PackSrc {
  const float* w[NSTAGE]; int ld[NSTAGE];   // weight of each stage ([N, ld] row-major)
  const float* b[NSTAGE];
  const float *alpha_w, *alpha_b, *rgb_w, *rgb_b;
};

template <int NSPLIT>
__global__ void __launch_bounds__(256) pack_weights_kernel(PackSrc src, uint8_t* __restrict__ img,
                                                           float* __restrict__ cbuf) {
  // one thread per (stage, k16 slab, row, k-chunk of 8)
  int g = blockIdx.x * blockDim.x + threadIdx.x;
  size_t base = 0;
  for (int s = 0; s < NSTAGE; ++s) {
    const StageDef d = stage_def(s);
    const int nk16 = d.kx + d.kh + d.kv;
    const int work = nk16 * d.N * 2;
    if (g < work) {
      int j = g / (d.N * 2), rem = g % (d.N * 2), chunk = rem / d.N, row = rem % d.N;
      // source columns of this k16 slab
      int col0, valid;
      if (j < d.kx) { col0 = 16 * j; valid = min(16, 63 - 16 * j); }
      else if (j < d.kx + d.kh) { col0 = (d.kx ? 63 : 0) + 16 * (j - d.kx); valid = 16; }
      else { col0 = 256 + 16 * (j - d.kx - d.kh); valid = min(16, 27 - 16 * (j - d.kx - d.kh)); }
      float v[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        int k = chunk * 8 + e;
        v[e] = (k < valid) ? src.w[s][(int64_t)row * src.ld[s] + col0 + k] : 0.f;
      }
      uint32_t h[4], l[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        __nv_bfloat16 h0 = __float2bfloat16_rn(v[2 * e]), h1 = __float2bfloat16_rn(v[2 * e + 1]);
        h[e] = (uint32_t)__bfloat16_as_ushort(h0) | ((uint32_t)__bfloat16_as_ushort(h1) << 16);
        l[e] = tc::pack_bf16(v[2 * e] - __bfloat162float(h0), v[2 * e + 1] - __bfloat162float(h1));
      }
      const size_t sb = (size_t)d.N * 32;             // bytes of one k16 slab (one half)
      const size_t in_slab = (size_t)chunk * d.N * 16 + (row >> 3) * 128 + (row & 7) * 16;
      if (NSPLIT == 1) {
        *reinterpret_cast<uint4*>(img + base + (size_t)j * sb + in_slab) = make_uint4(h[0], h[1], h[2], h[3]);
      } else {
        *reinterpret_cast<uint4*>(img + base + (size_t)j * 2 * sb + in_slab) = make_uint4(h[0], h[1], h[2], h[3]);
        *reinterpret_cast<uint4*>(img + base + (size_t)j * 2 * sb + sb + in_slab) = make_uint4(l[0], l[1], l[2], l[3]);
      }
      return;
    }
    g -= work;
    base += (size_t)nk16 * d.N * 32 * (NSPLIT == 1 ? 1 : 2);
  }
  // remaining threads: constants
  if (g < C_TOTAL) {
    float v = 0.f;
    if (g < C_BFEAT) v = src.b[g / 256][g % 256];
    else if (g < C_BVIEW) v = src.b[8][g - C_BFEAT];
    else if (g < C_WALPHA) v = src.b[9][g - C_BVIEW];
    else if (g < C_WRGB) v = src.alpha_w[g - C_WALPHA];
    else if (g < C_SCAL) v = src.rgb_w[g - C_WRGB];
    else if (g == C_SCAL) v = src.alpha_b[0];
    else if (g < C_SCAL + 4) v = src.rgb_b[g - C_SCAL - 1];
    cbuf[g] = v;
  }
}
inline int pack_total_threads() {
  int n = 0;
  for (int s = 0; s < NSTAGE; ++s) n += stage_k16(s) * stage_def(s).N * 2;
  return n + C_TOTAL;
}

// ---- the fused kernel --------------------------------------------------------------------------------
__device__ __forceinline__ void pe_store(uint8_t* img_hi, uint8_t* img_lo, int row, int k0,
                                         const float (&v)[8], bool split) {
  uint32_t h[4], l[4];
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    __nv_bfloat16 h0 = __float2bfloat16_rn(v[2 * e]), h1 = __float2bfloat16_rn(v[2 * e + 1]);
    h[e] = (uint32_t)__bfloat16_as_ushort(h0) | ((uint32_t)__bfloat16_as_ushort(h1) << 16);
    l[e] = tc::pack_bf16(v[2 * e] - __bfloat162float(h0), v[2 * e + 1] - __bfloat162float(h1));
  }
  uint32_t off = tc::canon_off(row, k0, TILE_M);
  *reinterpret_cast<uint4*>(img_hi + off) = make_uint4(h[0], h[1], h[2], h[3]);
  if (split) *reinterpret_cast<uint4*>(img_lo + off) = make_uint4(l[0], l[1], l[2], l[3]);
}

// PE of a 3-vector into `ncols_pad` columns (63 -> 64, 27 -> 32), canonical K-major smem image
template <int L, int NCHUNK>
__device__ __forceinline__ void pe_write(const float (&x)[3], bool valid, uint8_t* hi, uint8_t* lo,
                                         int row, bool split, float* dump) {
  float e[NCHUNK * 8];
#pragma unroll
  for (int i = 0; i < NCHUNK * 8; ++i) e[i] = 0.f;
  if (valid) {
    e[0] = x[0]; e[1] = x[1]; e[2] = x[2];
#pragma unroll
    for (int f = 0; f < L; ++f) {
      float fr = (float)(1 << f);
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        float s, cs;
        sincosf(x[c] * fr, &s, &cs);
        e[3 + 6 * f + c] = s;
        e[3 + 6 * f + 3 + c] = cs;
      }
    }
  }
  if (dump && valid) {
#pragma unroll
    for (int i = 0; i < 3 + 6 * L; ++i) dump[i] = e[i];
  }
#pragma unroll
  for (int ch = 0; ch < NCHUNK; ++ch) {
    float v[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = e[ch * 8 + i];
    pe_store(hi, lo, row, ch * 8, v, split);
  }
}

template <int NSPLIT>
__global__ void __launch_bounds__(192, 1) field_fused_fwd_kernel(Args a) {
  using C = Cfg<NSPLIT>;
  constexpr bool SPLIT = NSPLIT == 3;
  extern __shared__ uint8_t fused_smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(fused_smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* ring = smem + C::OFF_RING;
  uint8_t* Xhi = smem + C::OFF_X;
  uint8_t* Xlo = Xhi + C::X_BYTES;
  uint8_t* Vhi = smem + C::OFF_V;
  uint8_t* Vlo = Vhi + C::V_BYTES;
  float* cst = reinterpret_cast<float*>(smem + C::OFF_C);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + C::OFF_BAR);
  uint64_t* empty = full + C::NSLOT;
  uint64_t* acc_full = empty + C::NSLOT;
  uint64_t* a_ready = acc_full + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(a_ready + 1);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    for (int i = 0; i < C::NSLOT; ++i) { tc::mbar_init(&full[i], 1); tc::mbar_init(&empty[i], 1); }
    tc::mbar_init(acc_full, 1);
    tc::mbar_init(a_ready, 128);
    tc::fence_mbar_init();
  }
  for (int i = tid; i < C_TOTAL; i += blockDim.x) cst[i] = a.cbuf[i];
  __syncthreads();
  if (warp == 1) tc::tmem_alloc(tmem_slot, 512);
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const uint32_t T_ACC = tmem, T_AHI = tmem + 256, T_ALO = tmem + 384;

  if (warp == 0) {
    // ===================== TMA producer: stream the weight image, once per tile ====================
    if (lane == 0) {
      uint32_t n = 0;
      for (int tile = blockIdx.x; tile < a.num_tiles; tile += gridDim.x) {
        const uint8_t* src = a.wimg;
#pragma unroll 1
        for (int s = 0; s < NSTAGE; ++s) {
          const uint32_t bytes = (uint32_t)slab_bytes(s) * 2;
          const int ns = slots_in_stage<NSPLIT>(s);
#pragma unroll 1
          for (int i = 0; i < ns; ++i, ++n) {
            const uint32_t idx = n % C::NSLOT, ph = (n / C::NSLOT) & 1;
            tc::mbar_wait(&empty[idx], ph ^ 1);
            tc::mbar_arrive_expect_tx(&full[idx], bytes);
            tc::bulk_g2s(ring + idx * C::SLOT_BYTES, src, bytes, &full[idx]);
            src += bytes;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (one thread) ==================================================
    if (lane == 0) {
      uint32_t n = 0, q = 0;
      const uint32_t x_addr = tc::smem_u32(Xhi), v_addr = tc::smem_u32(Vhi), ring_addr = tc::smem_u32(ring);
      for (int tile = blockIdx.x; tile < a.num_tiles; tile += gridDim.x) {
#pragma unroll 1
        for (int s = 0; s < NSTAGE; ++s, ++q) {
          const StageDef d = stage_def(s);
          const uint32_t idesc = tc::idesc_bf16_f32(TILE_M, (uint32_t)d.N);
          const uint32_t sb = (uint32_t)d.N * 32, lbo_b = (uint32_t)d.N * 16;
          tc::mbar_wait(a_ready, q & 1);     // this stage's A operand is in place, ACC is free
          tc::tc_fence_after();
          const int nk16 = d.kx + d.kh + d.kv;
          uint32_t acc_flag = 0;
#pragma unroll 1
          for (int j = 0; j < nk16; j += (SPLIT ? 1 : 2), ++n) {
            const uint32_t idx = n % C::NSLOT, ph = (n / C::NSLOT) & 1;
            tc::mbar_wait(&full[idx], ph);
            tc::tc_fence_after();
            const uint32_t slot = ring_addr + idx * C::SLOT_BYTES;
#pragma unroll
            for (int u = 0; u < (SPLIT ? 1 : 2); ++u) {
              const int jj = j + u;
              const uint64_t b_hi = tc::smem_desc(slot + (SPLIT ? 0u : (uint32_t)u * sb), lbo_b, 128);
              const uint64_t b_lo = tc::smem_desc(slot + sb, lbo_b, 128);
              if (jj < d.kx || jj >= d.kx + d.kh) {
                // A slab from shared memory (PE of points / PE of the view direction)
                const bool isx = jj < d.kx;
                const uint32_t abase = (isx ? x_addr : v_addr) + (uint32_t)(isx ? jj : jj - d.kx - d.kh) * 4096u;
                const uint32_t lo_off = isx ? (uint32_t)C::X_BYTES : (uint32_t)C::V_BYTES;
                const uint64_t a_hi = tc::smem_desc(abase, 2048, 128);
                tc::mma_ss(T_ACC, a_hi, b_hi, idesc, acc_flag);
                acc_flag = 1;
                if (SPLIT) {
                  const uint64_t a_lo = tc::smem_desc(abase + lo_off, 2048, 128);
                  tc::mma_ss(T_ACC, a_lo, b_hi, idesc, 1);
                  tc::mma_ss(T_ACC, a_hi, b_lo, idesc, 1);
                }
              } else {
                const uint32_t col = (uint32_t)(jj - d.kx) * 8u;
                tc::mma_ts(T_ACC, T_AHI + col, b_hi, idesc, acc_flag);
                acc_flag = 1;
                if (SPLIT) {
                  tc::mma_ts(T_ACC, T_ALO + col, b_hi, idesc, 1);
                  tc::mma_ts(T_ACC, T_AHI + col, b_lo, idesc, 1);
                }
              }
            }
            tc::tc_commit(&empty[idx]);     // slot reusable once these MMAs retire
          }
          tc::tc_commit(acc_full);          // accumulator of this stage complete
        }
      }
    }
  } else {
    // ===================== epilogue warps (128 threads, thread <-> tile row) ======================
    const int quad = warp & 3;
    const int row = quad * 32 + lane;
    const uint32_t lane_base = (uint32_t)(quad * 32) << 16;
    uint32_t m = 0;
    for (int tile = blockIdx.x; tile < a.num_tiles; tile += gridDim.x) {
      const int64_t p = (int64_t)tile * TILE_M + row;
      const bool valid = p < a.P;
      // ---- positional encodings of this row's point and view direction -> smem A slabs ----------
      {
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
        pe_write<10, 8>(x, valid, Xhi, Xlo, row, SPLIT, a.dump_pe ? a.dump_pe + p * a.dump_pe_ld : nullptr);
        pe_write<4, 4>(vd, valid, Vhi, Vlo, row, SPLIT, a.dump_ped ? a.dump_ped + p * a.dump_ped_ld : nullptr);
        tc::fence_proxy_async();
        tc::mbar_arrive(a_ready);
      }
      float alpha = 0.f, rgb[3] = {0.f, 0.f, 0.f};
#pragma unroll 1
      for (int s = 0; s < NSTAGE; ++s, ++m) {
        const StageDef d = stage_def(s);
        const float* bias = cst + (s < 8 ? C_BIAS + s * 256 : (s == 8 ? C_BFEAT : C_BVIEW));
        tc::mbar_wait(acc_full, m & 1);
        tc::tc_fence_after();
#pragma unroll 1
        for (int c0 = 0; c0 < d.N; c0 += 32) {
          uint32_t v[32];
          tc::tmem_ld32(T_ACC + lane_base + c0, v);
          tc::tmem_ld_wait();
          float f[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            float t = __uint_as_float(v[j]) + bias[c0 + j];
            f[j] = d.relu ? fmaxf(t, 0.f) : t;
          }
          if (a.dump[s] && valid) {
            float* dp = a.dump[s] + p * a.dump_ld[s] + c0;
#pragma unroll
            for (int j = 0; j < 32; ++j) dp[j] = f[j];
          }
          if (s == 7) {
#pragma unroll
            for (int j = 0; j < 32; ++j) alpha = fmaf(f[j], cst[C_WALPHA + c0 + j], alpha);
          }
          if (s == 9) {
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              rgb[0] = fmaf(f[j], cst[C_WRGB + c0 + j], rgb[0]);
              rgb[1] = fmaf(f[j], cst[C_WRGB + 128 + c0 + j], rgb[1]);
              rgb[2] = fmaf(f[j], cst[C_WRGB + 256 + c0 + j], rgb[2]);
            }
          } else {
            uint32_t hi[16], lo[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              __nv_bfloat16 h0 = __float2bfloat16_rn(f[2 * j]), h1 = __float2bfloat16_rn(f[2 * j + 1]);
              hi[j] = (uint32_t)__bfloat16_as_ushort(h0) | ((uint32_t)__bfloat16_as_ushort(h1) << 16);
              if (SPLIT) lo[j] = tc::pack_bf16(f[2 * j] - __bfloat162float(h0), f[2 * j + 1] - __bfloat162float(h1));
            }
            tc::tmem_st16(T_AHI + lane_base + (uint32_t)(c0 >> 1), hi);
            if (SPLIT) tc::tmem_st16(T_ALO + lane_base + (uint32_t)(c0 >> 1), lo);
          }
        }
        if (s < 9) {
          tc::tmem_st_wait();
          tc::tc_fence_before();
          tc::mbar_arrive(a_ready);
        }
      }
      if (valid) {
        float4 o = make_float4(rgb[0] + cst[C_SCAL + 1], rgb[1] + cst[C_SCAL + 2], rgb[2] + cst[C_SCAL + 3],
                               alpha + cst[C_SCAL]);
        *reinterpret_cast<float4*>(a.raw + p * 4) = o;
      }
      // order this tile's last TMEM reads before the next tile's first MMA (issued after a_ready)
      tc::tc_fence_before();
    }
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 1) tc::tmem_dealloc(tmem, 512);
}

}  // namespace fused
}  // namespace scnerf
