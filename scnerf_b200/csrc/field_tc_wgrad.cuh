// Weight-gradient pass on tcgen05 (sm_100a):  dW_l[out,in] = sum_samples dZ_l[s,out] * X_l[s,in].
//
// Both operands are the bf16 (hi[,lo]) TILE IMAGES written by the fused forward (layer inputs) and the
// fused dgrad (dZ): [tile][k16 slab][half][F*32 B], each slab in the UMMA canonical MN-major
// no-swizzle layout, so one 1-D bulk TMA copy per image per K=16 step feeds the MMAs directly
// (A = dZ^T: M = out features, K = samples; B = X^T: N = in features).  A CTA owns ONE job (a layer's
// accumulator, both 128-row halves = 512 TMEM columns) over a contiguous slice of tiles, accumulates
// across the whole slice in TMEM, then adds its partial dW into the fp32 gradient with atomics.
// Helper warps reduce the bias gradients (column sums of dZ) — and d(alpha_linear.weight) — from the
// very same shared-memory slabs while the tensor pipe works.
//
// HBM-bound by construction: per tile and job 2 images x 64 KB x NH in, 48 x NSPLIT/3 MMAs out;
// algorithmic bytes per sample: sum over layers of (F_dZ + F_X) * 2 B * NH.
#pragma once
#include "common.cuh"
#include "tc_prims.cuh"
#include "tc_engine.cuh"

namespace scnerf {
namespace wgrad {

constexpr int MAX_UNITS = 4;
constexpr int NJOBS = 10;

struct Unit {
  int a_img, a_half;      // which A image of the job, which 128-feature half (M block)
  int b_img;              // which B image of the job
  int n;                  // GEMM N (multiple of 16)
  int acc_col;            // TMEM column of this unit's accumulator
  float* out;             // dW + row0*ld + col0
  int ld, rows_valid, cols_valid;
};
struct Job {
  eng::ImgDump a[2]; int na;
  eng::ImgDump b[2]; int nb;
  Unit u[MAX_UNITS]; int nu;
  float* db;              // bias gradient of a[0] (F = a[0].F) or NULL
  const float* g_raw;     // J8 only: d(raw)[P,4] for d(alpha_linear.weight)
  float* dw_alpha;        //          [256]
};
struct Args {
  Job job[NJOBS];
  int num_tiles;
  int cta0[NJOBS + 1];     // job j owns CTAs [cta0[j], cta0[j+1]): its tiles are split evenly over them
  int64_t P;
  int l2_prefetch_slots;   // 0 = off
  long long* dbg; int dbg_slots, dbg_cta;   // optional timeline of CTA 0: [slot][4] = producer got slot free, MMA saw
                                   // slot full, MMA committed, helper warp 2 released the slot
};

template <int NSPLIT> struct Cfg {
  static constexpr int NH = NSPLIT == 3 ? 2 : 1;
  static constexpr int SLOT_BYTES = 38912 / (NSPLIT == 3 ? 1 : 2);   // worst job: 2 x 256 + 96 features (4-D points; 64 for 3-D)
  static constexpr int NSLOT = NSPLIT == 3 ? 5 : 10;
  static constexpr int OFF_BAR = NSLOT * SLOT_BYTES;
  static constexpr int SMEM_BYTES = OFF_BAR + (2 * NSLOT + 1) * 8 + 16;
};

// instruction descriptor: bf16 x bf16 -> f32, A and B both MN-major
__host__ __device__ constexpr uint32_t idesc_mn(uint32_t M, uint32_t N) {
  return (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((N >> 3) << 17) | ((M >> 4) << 24);
}

// sum of the 8 bf16 features of a 16-byte chunk (hi [+ lo]) into acc[8], optionally scaled
__device__ __forceinline__ void add_chunk(const uint4& c, float s, float (&acc)[8]) {
  const uint32_t w[4] = {c.x, c.y, c.z, c.w};
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    acc[2 * e] = fmaf(s, eng::bf16lo_f(w[e]), acc[2 * e]);
    acc[2 * e + 1] = fmaf(s, eng::bf16hi_f(w[e]), acc[2 * e + 1]);
  }
}

template <int NSPLIT>
__global__ void __launch_bounds__(320, 1) field_wgrad_kernel(const __grid_constant__ Args a_) {
  const Args* ap = &a_;
  using C = Cfg<NSPLIT>;
  constexpr int NH = C::NH;
  constexpr bool SPLIT = NSPLIT == 3;
  extern __shared__ __align__(128) uint8_t wsm[];
  uint64_t* full = reinterpret_cast<uint64_t*>(wsm + C::OFF_BAR);
  uint64_t* empty = full + C::NSLOT;
  uint64_t* acc_done = empty + C::NSLOT;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_done + 1);
  __shared__ Job job;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  int jid = 0;
#pragma unroll
  for (int j = 1; j < NJOBS; ++j) jid += (int)blockIdx.x >= ap->cta0[j];
  const int slice = (int)blockIdx.x - ap->cta0[jid];
  if (tid == 0) job = ap->job[jid];
  if (ap->dbg && tid == 0) {
    unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    ap->dbg[ap->dbg_slots * 4 + blockIdx.x * 4 + 0] = (long long)t;
    ap->dbg[ap->dbg_slots * 4 + blockIdx.x * 4 + 3] = clock64();
  }
  if (tid == 32) {
    // full: 1 (producer expect_tx for the TMA part) + 128 (cp.async loader threads); empty: MMA commit + 4 helper warps
    for (int i = 0; i < C::NSLOT; ++i) { tc::mbar_init(&full[i], 1 + 128); tc::mbar_init(&empty[i], 5); }
    tc::mbar_init(acc_done, 1);
    tc::fence_mbar_init();
  }
  __syncthreads();
  if (warp == 1) tc::tmem_alloc(tmem_slot, 512);
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const int nslices = ap->cta0[jid + 1] - ap->cta0[jid];
  const int per = (ap->num_tiles + nslices - 1) / nslices;
  const int t0 = min(slice * per, ap->num_tiles), t1 = min(t0 + per, ap->num_tiles);
  // byte offsets of each image's slab inside a slot
  uint32_t offA[2], offB[2], bytesA[2], bytesB[2], slot_bytes = 0;
  for (int i = 0; i < job.na; ++i) { offA[i] = slot_bytes; bytesA[i] = job.a[i].F * 32u * NH; slot_bytes += bytesA[i]; }
  for (int i = 0; i < job.nb; ++i) { offB[i] = slot_bytes; bytesB[i] = job.b[i].F * 32u * NH; slot_bytes += bytesB[i]; }

  if (warp == 0) {
    if (lane == 0) {
      // a tile's 8 slabs are contiguous in each image: one running pointer per image
      const uint8_t* pa0 = job.a[0].base + (size_t)t0 * job.a[0].tile_bytes();
      const uint8_t* pa1 = job.na > 1 ? job.a[1].base + (size_t)t0 * job.a[1].tile_bytes() : nullptr;
      const uint8_t* pb0 = job.b[0].base + (size_t)t0 * job.b[0].tile_bytes();
      const uint8_t* pb1 = job.nb > 1 ? job.b[1].base + (size_t)t0 * job.b[1].tile_bytes() : nullptr;
      uint32_t idx = 0, ph = 0;
      const int nslots = (t1 - t0) * 8;
      const uint64_t pol = tc::policy_evict_first();   // read-once streams: do not displace the weights in L2
      // The TMA unit keeps only ~16 KB of requests outstanding per SM, so copies that miss L2 stream at
      // ~11 B/clk/SM (2.7 TB/s chip-wide, measured).  L2 prefetches are fire-and-forget: run them PF
      // slots ahead so the bulk copies themselves hit L2.
      const int PF = ap->l2_prefetch_slots;
      for (int it = 0; it < min(PF, nslots); ++it) {
        tc::bulk_prefetch_l2(pa0 + (size_t)it * bytesA[0], bytesA[0]);
        if (pa1) tc::bulk_prefetch_l2(pa1 + (size_t)it * bytesA[1], bytesA[1]);
        tc::bulk_prefetch_l2(pb0 + (size_t)it * bytesB[0], bytesB[0]);
        if (pb1) tc::bulk_prefetch_l2(pb1 + (size_t)it * bytesB[1], bytesB[1]);
      }
#pragma unroll 1
      for (int it = 0; it < nslots; ++it) {
        if (PF > 0 && it + PF < nslots) {
          tc::bulk_prefetch_l2(pa0 + (size_t)PF * bytesA[0], bytesA[0]);
          if (pa1) tc::bulk_prefetch_l2(pa1 + (size_t)PF * bytesA[1], bytesA[1]);
          tc::bulk_prefetch_l2(pb0 + (size_t)PF * bytesB[0], bytesB[0]);
          if (pb1) tc::bulk_prefetch_l2(pb1 + (size_t)PF * bytesB[1], bytesB[1]);
        }
        tc::mbar_wait(&empty[idx], ph ^ 1);
        if (ap->dbg && blockIdx.x == ap->dbg_cta && (it & 31) == 0 && (it >> 5) < ap->dbg_slots) ap->dbg[(it >> 5) * 4 + 0] = clock64();
        tc::mbar_arrive_expect_tx(&full[idx], bytesA[0] + (pa1 ? bytesA[1] : 0u));   // A images ride the TMA
        uint8_t* dst = wsm + idx * C::SLOT_BYTES;
        const uint32_t d32 = tc::smem_u32(dst), fb = tc::smem_u32(&full[idx]);
        tc::bulk_g2s_hint(d32 + offA[0], pa0, bytesA[0], fb, pol); pa0 += bytesA[0];
        if (pa1) { tc::bulk_g2s_hint(d32 + offA[1], pa1, bytesA[1], fb, pol); pa1 += bytesA[1]; }
        if (++idx == C::NSLOT) { idx = 0; ph ^= 1u; }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // everything that does not depend on the ring slot is hoisted: per unit the operand offsets inside
      // a slot, the instruction descriptor and the accumulator address; per slot only the 14-bit start
      // address field of each descriptor changes (one add)
      const uint32_t base = tc::smem_u32(wsm);
      const uint32_t full0 = tc::smem_u32(full), empty0 = tc::smem_u32(empty);
      uint32_t ua[MAX_UNITS], ual[MAX_UNITS], ub[MAX_UNITS], ubl[MAX_UNITS], uid[MAX_UNITS], uacc[MAX_UNITS];
#pragma unroll
      for (int ui = 0; ui < MAX_UNITS; ++ui) {
        const Unit& u = job.u[ui < job.nu ? ui : 0];
        ua[ui] = offA[u.a_img] + (uint32_t)u.a_half * 4096u;      // 16 mn-groups x 256 B
        ual[ui] = ua[ui] + job.a[u.a_img].F * 32u;
        ub[ui] = offB[u.b_img];
        ubl[ui] = ub[ui] + job.b[u.b_img].F * 32u;
        uid[ui] = idesc_mn(128, (uint32_t)u.n);
        uacc[ui] = tmem + (uint32_t)u.acc_col;
      }
      const int nu = job.nu;
      uint32_t idx = 0, ph = 0, first = 0;
      const int nslots = (t1 - t0) * 8;
#pragma unroll 1
      for (int it = 0; it < nslots; ++it) {
        eng::mbar_wait_a(full0 + idx * 8, ph);
        if (ap->dbg && blockIdx.x == ap->dbg_cta && (it & 31) == 0 && (it >> 5) < ap->dbg_slots) ap->dbg[(it >> 5) * 4 + 1] = clock64();
        tc::fence_proxy_async();   // B slabs arrive through cp.async (generic proxy); the MMA reads via the async proxy
        tc::tc_fence_after();
        const uint32_t slot = base + idx * C::SLOT_BYTES;
#pragma unroll
        for (int ui = 0; ui < MAX_UNITS; ++ui) {
          if (ui < nu) {
            const uint64_t dah = eng::desc_at<128, 256>(slot + ua[ui]), dbh = eng::desc_at<128, 256>(slot + ub[ui]);
            tc::mma_ss(uacc[ui], dah, dbh, uid[ui], first);
            if (SPLIT) {
              tc::mma_ss(uacc[ui], eng::desc_at<128, 256>(slot + ual[ui]), dbh, uid[ui], 1);
              tc::mma_ss(uacc[ui], dah, eng::desc_at<128, 256>(slot + ubl[ui]), uid[ui], 1);
            }
          }
        }
        first = 1;
        eng::commit_a(empty0 + idx * 8);
        if (ap->dbg && blockIdx.x == ap->dbg_cta && (it & 31) == 0 && (it >> 5) < ap->dbg_slots) ap->dbg[(it >> 5) * 4 + 2] = clock64();
        if (++idx == C::NSLOT) { idx = 0; ph ^= 1u; }
      }
      tc::tc_commit(acc_done);
      if (ap->dbg) {
        unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        ap->dbg[ap->dbg_slots * 4 + blockIdx.x * 4 + 1] = (long long)t;
      }
    }
  } else if (warp >= 6) {
    // ===================== loader warps 6..9: B images through the LSU path (cp.async) ==============
    // The TMA unit alone sustains ~11 B/clk/SM against HBM latency; 128 threads x 16-byte cp.async keep
    // far more requests in flight, so the two paths together roughly double the stream.
    const int lt = tid - 192;                // 0..127
    const uint8_t* pb0 = job.b[0].base + (size_t)t0 * job.b[0].tile_bytes();
    const uint8_t* pb1 = job.nb > 1 ? job.b[1].base + (size_t)t0 * job.b[1].tile_bytes() : nullptr;
    uint32_t idx = 0, ph = 0;
    const int nslots = (t1 - t0) * 8;
#pragma unroll 1
    for (int it = 0; it < nslots; ++it) {
      tc::mbar_wait(&empty[idx], ph ^ 1);
      const uint32_t d32 = tc::smem_u32(wsm + idx * C::SLOT_BYTES);
      for (uint32_t o = (uint32_t)lt * 16u; o < bytesB[0]; o += 128u * 16u)
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d32 + offB[0] + o), "l"(pb0 + o) : "memory");
      pb0 += bytesB[0];
      if (pb1) {
        for (uint32_t o = (uint32_t)lt * 16u; o < bytesB[1]; o += 128u * 16u)
          asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d32 + offB[1] + o), "l"(pb1 + o) : "memory");
        pb1 += bytesB[1];
      }
      // arrives on full[idx] when this thread's copies have landed (counted in the barrier's 1+128)
      asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(tc::smem_u32(&full[idx])) : "memory");
      if (++idx == C::NSLOT) { idx = 0; ph ^= 1u; }
    }
  } else {
    // ===================== helper warps 2..5: bias / alpha-weight gradients, then the dW epilogue ===
    // Lane -> (sample k, mn-group parity): the 16 samples of one mn-group are 256 contiguous bytes of the
    // slab, so a warp's 32 x 16-byte reads cover 512 contiguous bytes: conflict-free.  (The earlier
    // group-major mapping was 4-way bank conflicted and, with the MMA operand reads and the ring fills,
    // saturated shared-memory bandwidth: profiles/README.md, r1h.)
    const int hw = warp - 2;                 // 0..3
    const int k = lane & 15, gs = lane >> 4; // sample within the slab, mn-group parity
    const int FA = (int)job.a[0].F;          // 256 or 128 features
    const bool do_bias = job.db != nullptr;
    const bool do_alpha = job.dw_alpha != nullptr;
    float accb[4][8], acca[4][8];            // mn-groups hw*8 + 2j + gs, j = 0..3
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
      for (int e = 0; e < 8; ++e) { accb[j][e] = 0.f; acca[j][e] = 0.f; }
    uint32_t idx = 0, ph = 0;
    int hit = 0;
    // d(raw)[:, 3] of the NEXT tile is fetched while the current one is consumed: a dependent global load
    // inside the slot loop would hold the ring slot for a full HBM round trip (measured: 2.2x on job 8)
    float gcur[8], gnext[8];
    auto load_alpha = [&](int tile, float (&dst)[8]) {
#pragma unroll
      for (int ks = 0; ks < 8; ++ks) {
        const int64_t p = (int64_t)tile * 128 + ks * 16 + k;
        dst[ks] = p < ap->P ? __ldg(job.g_raw + p * 4 + 3) : 0.f;
      }
    };
    if (do_alpha && t0 < t1) load_alpha(t0, gnext);
    for (int tile = t0; tile < t1; ++tile) {
      if (do_alpha) {
#pragma unroll
        for (int ks = 0; ks < 8; ++ks) gcur[ks] = gnext[ks];
        if (tile + 1 < t1) load_alpha(tile + 1, gnext);
      }
#pragma unroll
      for (int ks = 0; ks < 8; ++ks, ++hit) {
        tc::mbar_wait(&full[idx], ph);
        const uint8_t* slot = wsm + idx * C::SLOT_BYTES;
        if (do_bias || do_alpha) {
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int g = hw * 8 + j * 2 + gs;
            const uint32_t coff = (uint32_t)g * 256u + (uint32_t)k * 16u;
            if (do_bias && g * 8 < FA) {
              add_chunk(*reinterpret_cast<const uint4*>(slot + offA[0] + coff), 1.f, accb[j]);
              if (SPLIT) add_chunk(*reinterpret_cast<const uint4*>(slot + offA[0] + (uint32_t)FA * 32u + coff), 1.f, accb[j]);
            }
            if (do_alpha) {
              const float ga = gcur[ks];
              add_chunk(*reinterpret_cast<const uint4*>(slot + offB[0] + coff), ga, acca[j]);
              if (SPLIT) add_chunk(*reinterpret_cast<const uint4*>(slot + offB[0] + job.b[0].F * 32u + coff), ga, acca[j]);
            }
          }
        }
        __syncwarp();
        if (lane == 0) tc::mbar_arrive(&empty[idx]);
        if (ap->dbg && blockIdx.x == ap->dbg_cta && tid == 64 && (hit & 31) == 0 && (hit >> 5) < ap->dbg_slots) ap->dbg[(hit >> 5) * 4 + 3] = clock64();
        if (++idx == C::NSLOT) { idx = 0; ph ^= 1u; }
      }
    }
    // reduce over the 16 samples (lanes with equal parity) and publish
    if (t1 > t0) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int g = hw * 8 + j * 2 + gs;
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          float vb = accb[j][e], va = acca[j][e];
#pragma unroll
          for (int m = 1; m < 16; m <<= 1) {
            vb += __shfl_xor_sync(0xffffffffu, vb, m);
            va += __shfl_xor_sync(0xffffffffu, va, m);
          }
          if (k == 0) {
            if (do_bias && g * 8 < FA) atomicAdd(job.db + g * 8 + e, vb);
            if (do_alpha) atomicAdd(job.dw_alpha + g * 8 + e, va);
          }
        }
      }
    }
    // ---- dW epilogue: TMEM -> atomicAdd into the fp32 gradient -------------------------------------
    tc::mbar_wait(acc_done, 0);
    tc::tc_fence_after();
    if (t1 > t0) {
      const int quad = warp & 3;
      const uint32_t lane_base = (uint32_t)(quad * 32) << 16;
      const int r = quad * 32 + lane;
      for (int ui = 0; ui < job.nu; ++ui) {
        const Unit& u = job.u[ui];
        for (int c0 = 0; c0 < u.n; c0 += 32) {
          uint32_t v[32];
          if (u.n - c0 >= 32) tc::tmem_ld32(tmem + lane_base + (uint32_t)(u.acc_col + c0), v);
          else {
            uint32_t w[16];
            asm volatile(
                "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3]), "=r"(w[4]), "=r"(w[5]), "=r"(w[6]), "=r"(w[7]),
                  "=r"(w[8]), "=r"(w[9]), "=r"(w[10]), "=r"(w[11]), "=r"(w[12]), "=r"(w[13]), "=r"(w[14]), "=r"(w[15])
                : "r"(tmem + lane_base + (uint32_t)(u.acc_col + c0)) : "memory");
#pragma unroll
            for (int j = 0; j < 16; ++j) { v[j] = w[j]; v[16 + j] = 0u; }
          }
          tc::tmem_ld_wait();
          if (r < u.rows_valid) {
            float* orow = u.out + (int64_t)r * u.ld + c0;
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (c0 + j < u.cols_valid) atomicAdd(orow + j, __uint_as_float(v[j]));
          }
        }
      }
    }
  }
  tc::tc_fence_before();
  __syncthreads();
  if (ap->dbg && tid == 0) {
    unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    ap->dbg[ap->dbg_slots * 4 + blockIdx.x * 4 + 2] = (long long)t;
    ap->dbg[ap->dbg_slots * 4 + blockIdx.x * 4 + 3] = clock64() - ap->dbg[ap->dbg_slots * 4 + blockIdx.x * 4 + 3];
  }
  if (warp == 1) tc::tmem_dealloc(tmem, 512);
}

// d(rgb_linear.weight)[3,128], d(rgb_linear.bias)[3], d(alpha_linear.bias)[1] from d(raw) and the HV image
template <int NH>
__global__ void __launch_bounds__(128) head_wgrad_img_kernel(eng::ImgDump hv, const float* __restrict__ g_raw,
                                                             int64_t P, int num_tiles, float* __restrict__ dw_rgb,
                                                             float* __restrict__ db_rgb, float* __restrict__ db_alpha) {
  const int t = threadIdx.x, g = t >> 3, q = t & 7;     // 16 mn-groups x 8 sample sub-slices (2 samples)
  float acc[3][8];
#pragma unroll
  for (int c = 0; c < 3; ++c)
#pragma unroll
    for (int e = 0; e < 8; ++e) acc[c][e] = 0.f;
  float sb[4] = {0.f, 0.f, 0.f, 0.f};
  for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
    // 16 samples per thread and tile, in batches of 4 whose loads (d(raw) row + hi/lo chunk: 48 B) are all issued before
    // the first use: the loop is load-latency bound (0.28 -> see profiles/README.md round 2)
#pragma unroll 1
    for (int b = 0; b < 4; ++b) {
      float4 gr[4];
      uint4 ch[4], cl[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int s16 = b * 4 + u;                       // 0..15 -> (ks, i)
        const int kk = (s16 >> 1) * 16 + q * 2 + (s16 & 1);
        const int64_t p = (int64_t)tile * 128 + kk;
        const bool ok = p < P;
        gr[u] = ok ? *reinterpret_cast<const float4*>(g_raw + p * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
        ch[u] = ok ? *reinterpret_cast<const uint4*>(hv.chunk(tile, kk, g * 8, 0)) : make_uint4(0u, 0u, 0u, 0u);
        if (NH == 2) cl[u] = ok ? *reinterpret_cast<const uint4*>(hv.chunk(tile, kk, g * 8, 1)) : make_uint4(0u, 0u, 0u, 0u);
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        float h[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        add_chunk(ch[u], 1.f, h);
        if (NH == 2) add_chunk(cl[u], 1.f, h);
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          acc[0][e] = fmaf(gr[u].x, h[e], acc[0][e]);
          acc[1][e] = fmaf(gr[u].y, h[e], acc[1][e]);
          acc[2][e] = fmaf(gr[u].z, h[e], acc[2][e]);
        }
        if (g == 0) { sb[0] += gr[u].x; sb[1] += gr[u].y; sb[2] += gr[u].z; sb[3] += gr[u].w; }
      }
    }
  }
#pragma unroll
  for (int c = 0; c < 3; ++c)
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      float v = acc[c][e];
      v += __shfl_xor_sync(0xffffffffu, v, 1); v += __shfl_xor_sync(0xffffffffu, v, 2);
      v += __shfl_xor_sync(0xffffffffu, v, 4);
      if (q == 0) atomicAdd(dw_rgb + c * 128 + g * 8 + e, v);
    }
  if (g == 0) {
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      float v = sb[c];
      v += __shfl_xor_sync(0x000000ffu, v, 1); v += __shfl_xor_sync(0x000000ffu, v, 2);
      v += __shfl_xor_sync(0x000000ffu, v, 4);
      if (q == 0) atomicAdd(c < 3 ? db_rgb + c : db_alpha, v);
    }
  }
}

}  // namespace wgrad
}  // namespace scnerf
