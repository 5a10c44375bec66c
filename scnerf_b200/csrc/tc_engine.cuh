// Compile-time tcgen05 "slab engine" shared by the fused field kernels (forward, dgrad).
//
// A tile (128 samples) is processed as a fixed sequence of SLABS.  One slab = one K=16 step of one
// GEMM stage: B operand = an [N x 16] bf16 weight slab (hi [, lo]) that the producer warp streams
// from the packed weight image with one 1-D bulk TMA copy into a ring slot; A operand = 16 columns of
// the activation, either in TMEM (TS-mode MMA) or in a shared-memory canonical image (SS-mode).
//
// The slab sequence is a constexpr table.  The producer and the single MMA-issuing thread run a FULLY
// UNROLLED instantiation of it (one template instance per slab): ring slot, mbarrier address, phase
// parity, operand offsets and instruction descriptor are all immediates, so the issue thread spends a
// handful of uniform-datapath instructions per MMA instead of a table walk (first profile: ~450 issue
// cycles per slab against 128-384 cycles of tensor work).
#pragma once
#include <utility>
#include "common.cuh"
#include "tc_prims.cuh"

namespace scnerf {
namespace eng {

constexpr int TILE_M = 128;
constexpr int MAX_SLABS = 352;   // the N-half pipelined forward issues 336 half-slabs per tile

enum : uint8_t { A_TMEM = 0, A_SMEM = 1, A_MIX = 2 };   // A_MIX: hi in TMEM (a_off), lo in shared memory (a_lo_delta)
enum : uint16_t { F_ZERO_ACC = 1, F_STAGE_END = 2, F_HI_ONLY_A = 4, F_STAGE_BEGIN = 8,
                  F_WAIT_Q1 = 16, F_WAIT_Q2 = 32, F_WAIT_Q3 = 64, F_COMMIT_BOTH = 128,   // pipelined plans: A-ready barriers of column quarters 1..3 (F_STAGE_BEGIN = quarter 0)
                  // side pass that borrows accumulator half 1 (dgrad of 4-D points, 96-wide skip share of d(PE)): its last slab
                  // commits to the side barrier INSTEAD of accf[pad]; the pass that re-initialises the half waits until the
                  // epilogue has parked the side result (one use per tile: parity = tile parity)
                  F_COMMIT_SIDE = 256, F_WAIT_SIDE = 512 };

struct SlabDef {
  uint16_t n;          // rows of the B slab (= GEMM N), multiple of 16
  uint16_t acc_col;    // accumulator column offset in TMEM
  uint16_t a_off;      // A_TMEM: column offset inside the A_hi / A_lo regions (8 per k16)
                       // A_SMEM: byte offset / 16 of the hi slab inside the smem A area
  uint16_t a_lo_delta; // A_SMEM: byte distance / 16 from the hi to the lo image
  uint16_t flags;
  uint8_t a_kind;
  uint8_t stage;
  uint8_t pad;         // pipelined plans: which accumulator-half barrier F_STAGE_END commits to
  uint32_t img_off;    // byte offset of the slab in the NSPLIT==1 weight image (x2 for NSPLIT==3)
};

struct Plan {
  SlabDef slab[MAX_SLABS];
  int n_slabs;
  int n_stages;
};

// Kernel traits K must provide:  static constexpr Plan PLAN;  NSPLIT, GROUP, NSLOT, SLOT_BYTES.
// Requirement: (PLAN.n_slabs / GROUP) % NSLOT == 0 (ring position of a slab group is fixed in every tile).

struct Ctx {
  uint32_t ring_addr;      // shared address of ring slot 0
  uint32_t full_addr;      // shared address of full[0]   (8 bytes apart)
  uint32_t empty_addr;     // shared address of empty[0]
  uint32_t acc_full_addr, a_ready_addr;   // (unused since the serial schedules were retired; the pipelined kernels keep their
  uint32_t tmem_acc, tmem_ahi, tmem_alo;  //  barriers in fpipe::PCtx)
  uint32_t smem_a;         // shared address of the smem A area
  uint64_t pol_keep;       // L2 evict_last  (weight slabs: re-read by every SM for every tile)
  uint64_t pol_stream;     // L2 evict_first (write-once / read-once tile images)
  long long* dbg;          // (unused: the timeline stamps are fpipe::stamp)
  int dbg_tiles;
};
__device__ __forceinline__ void mbar_wait_a(uint32_t addr, uint32_t parity) {
  uint32_t ok;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok) : "r"(addr), "r"(parity) : "memory");
  } while (!ok);
}
__device__ __forceinline__ void commit_a(uint32_t addr) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(addr) : "memory");
}
// descriptor with a compile-time hi word: only the 14-bit start address varies
template <uint32_t LBO, uint32_t SBO>
__device__ __forceinline__ uint64_t desc_at(uint32_t saddr) {
  constexpr uint32_t hi = ((SBO >> 4) & 0x3FFF) | (1u << 14);
  const uint32_t lo = ((saddr >> 4) & 0x3FFF) | (((LBO >> 4) & 0x3FFF) << 16);
  return ((uint64_t)hi << 32) | lo;
}

// bytes of slabs [I0, I1) in the weight image
template <class K>
__host__ __device__ constexpr uint32_t group_bytes(int i0, int i1) {
  uint32_t b = 0;
  for (int i = i0; i < i1; ++i) b += (uint32_t)K::PLAN.slab[i].n * 32u * (K::NSPLIT == 3 ? 2u : 1u);
  return b;
}
// K::GROUP consecutive slabs share one ring slot (one expect_tx + one bulk copy, one commit): keeps the
// per-slot issue overhead (~150 cycles) below the MMA time of the slot also in single-pass bf16
template <class K, int I>
__device__ __forceinline__ void producer_step(const Ctx& c, const uint8_t* __restrict__ wimg, uint32_t tp) {
  if constexpr (I % K::GROUP == 0) {
    constexpr SlabDef d = K::PLAN.slab[I];
    constexpr int G = I / K::GROUP;
    constexpr int idx = G % K::NSLOT, wrap = G / K::NSLOT;
    constexpr bool wraps_odd = (((K::PLAN.n_slabs / K::GROUP) / K::NSLOT) & 1) != 0;
    constexpr uint32_t bytes = group_bytes<K>(I, I + K::GROUP);
    static_assert(bytes <= (uint32_t)K::SLOT_BYTES, "slab group does not fit a ring slot");
    constexpr uint32_t src_off = d.img_off * (K::NSPLIT == 3 ? 2u : 1u);
    const uint32_t ph = (uint32_t)(wrap & 1) ^ (wraps_odd ? tp : 0u);
    mbar_wait_a(c.empty_addr + idx * 8, ph ^ 1u);
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(c.full_addr + idx * 8), "r"(bytes) : "memory");
    tc::bulk_g2s_hint(c.ring_addr + idx * K::SLOT_BYTES, wimg + src_off, bytes, c.full_addr + idx * 8, c.pol_keep);
  }
}
template <class K, size_t... Is>
__device__ __forceinline__ void producer_tile(const Ctx& c, const uint8_t* __restrict__ wimg, uint32_t tp,
                                              std::index_sequence<Is...>) {
  (producer_step<K, (int)Is>(c, wimg, tp), ...);
}
template <class K>
__device__ __forceinline__ void producer_loop_n(const Ctx& c, const uint8_t* __restrict__ wimg, int count) {
  uint32_t tp = 0;
  for (int it = 0; it < count; ++it, tp ^= 1u)
    producer_tile<K>(c, wimg, tp, std::make_index_sequence<K::PLAN.n_slabs>{});
}

// (The MMA issue loop lives with the pipelined schedules: fpipe::mma_step / mma_loop_r in field_tc_fwd_pipe.cuh.)

// ---- epilogue helpers ----------------------------------------------------------------------------------
// pack two fp32 into bf16x2 (first argument -> low half), optionally with ReLU
__device__ __forceinline__ uint32_t cvt_bf16x2(float lo, float hi) {
  uint32_t d;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi), "f"(lo));
  return d;
}
__device__ __forceinline__ uint32_t cvt_relu_bf16x2(float lo, float hi) {
  uint32_t d;
  asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi), "f"(lo));
  return d;
}
__device__ __forceinline__ float bf16lo_f(uint32_t p) { return __uint_as_float(p << 16); }
__device__ __forceinline__ float bf16hi_f(uint32_t p) { return __uint_as_float(p & 0xffff0000u); }

// split 32 fp32 values into packed hi / lo bf16 pairs.  RELU: apply max(x,0) on the fly
// (cvt.rn.relu for the hi half; the residual uses the clamped value).
template <bool SPLIT, bool RELU>
__device__ __forceinline__ void split32(const float (&f)[32], uint32_t (&hi)[16], uint32_t (&lo)[16]) {
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    hi[j] = RELU ? cvt_relu_bf16x2(f[2 * j], f[2 * j + 1]) : cvt_bf16x2(f[2 * j], f[2 * j + 1]);
    if (SPLIT) {
      const float a = RELU ? fmaxf(f[2 * j], 0.f) : f[2 * j], b = RELU ? fmaxf(f[2 * j + 1], 0.f) : f[2 * j + 1];
      lo[j] = cvt_bf16x2(a - bf16lo_f(hi[j]), b - bf16hi_f(hi[j]));
    }
  }
}

// ReLU + split + mask word in one pass (the forward's hidden layers).  hi = bf16(max(x, 0)) by cvt.rn.relu; the packed
// compare m = (hi > 0) (0xffff per true half) then serves twice: it zeroes the lo halves of the clamped columns
// (lo = bf16(x - hi) & m: one LOP3 per PAIR instead of an FMNMX per element on the residual's input) and it is the ReLU
// mask the dgrad reads (bit j = column 2j, bit 16 + j = column 2j + 1).
template <bool SPLIT>
__device__ __forceinline__ void split32_relu(const float (&f)[32], uint32_t (&hi)[16], uint32_t (&lo)[16], uint32_t& bits) {
  bits = 0;
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    hi[j] = cvt_relu_bf16x2(f[2 * j], f[2 * j + 1]);
    uint32_t m;
    asm("set.gt.u32.bf16x2 %0, %1, %2;" : "=r"(m) : "r"(hi[j]), "r"(0u));
    bits |= m & (0x00010001u << j);
    if (SPLIT) lo[j] = cvt_bf16x2(f[2 * j] - bf16lo_f(hi[j]), f[2 * j + 1] - bf16hi_f(hi[j])) & m;
  }
}

// ---- ReLU masks (forward -> dgrad): one 32-bit word per 32 consecutive output columns of a row ----------------------
// Bit layout follows the packed bf16 pairs the epilogue already holds: bit j = column 2j, bit 16 + j = column 2j + 1.
// The forward derives the word from the post-ReLU hi halves with one packed compare per PAIR (set.gt.u32.bf16x2 gives
// 0xffff per true half) and one LOP3 — the per-element FSETP + SEL + add chain of the first version was 28 % of the
// training epilogue's instructions (160 of ~560 per thread and half-stage, cuobjdump).  bf16 keeps fp32's exponent
// range, so "hi > 0" and "x > 0" differ only for |x| < 2^-133.
__device__ __forceinline__ uint32_t relu_mask16(const uint32_t (&hi)[16]) {
  uint32_t bits = 0;
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    uint32_t m;
    asm("set.gt.u32.bf16x2 %0, %1, %2;" : "=r"(m) : "r"(hi[j]), "r"(0u));
    bits |= m & (0x00010001u << j);
  }
  return bits;
}
// the same word from 32 fp32 values (paths that do not form the packed pairs)
__device__ __forceinline__ uint32_t relu_mask32f(const uint32_t (&v)[32]) {
  uint32_t bits = 0;
#pragma unroll
  for (int j = 0; j < 32; ++j) bits |= (__uint_as_float(v[j]) > 0.f ? 1u : 0u) << ((j >> 1) + ((j & 1) << 4));
  return bits;
}
__device__ __forceinline__ bool relu_bit(uint32_t w, int j) { return ((w >> ((j >> 1) + ((j & 1) << 4))) & 1u) != 0u; }

// ---- tile-image dumps (the wgrad kernel's operand format) ---------------------------------------------
// A [128 samples x F features] bf16 tile is stored as 8 K16-slabs (16 samples each); inside a slab the
// layout is the UMMA canonical MN-major no-swizzle image:
//   byte(mn, k) = (mn/8)*256 + ((k/8)&1)*128 + (k&7)*16 + (mn&7)*2        (LBO = 128, SBO = 256)
// global: [tile][slab k/16][half: hi, lo][F*32 bytes]
struct ImgDump {
  uint8_t* base;       // NULL = disabled
  uint32_t F;          // features (multiple of 8)
  uint32_t nhalf;      // 1 (hi) or 2 (hi, lo)
  __host__ __device__ __forceinline__ size_t tile_bytes() const { return (size_t)F * 256u * nhalf; }
  // address of the 16-byte chunk (8 features starting at mn0, sample k of tile `tile`, half h)
  __device__ __forceinline__ uint8_t* chunk(int tile, uint32_t k, uint32_t mn0, uint32_t h) const {
    return base + (size_t)tile * tile_bytes() + (size_t)(k >> 4) * (F * 32u * nhalf) + (size_t)h * (F * 32u) +
           (mn0 >> 3) * 256u + ((k >> 3) & 1u) * 128u + (k & 7u) * 16u;
  }
};
// store 32 consecutive features [c0, c0+32) of sample k (packed pairs) into the image
template <bool SPLIT>
__device__ __forceinline__ void dump32(const ImgDump& d, int tile, uint32_t k, uint32_t c0,
                                       const uint32_t (&hi)[16], const uint32_t (&lo)[16], uint64_t pol) {
  uint8_t* p = d.chunk(tile, k, c0, 0);
#pragma unroll
  for (int g = 0; g < 4; ++g) {
    tc::st_v4_hint(p + g * 256, make_uint4(hi[4 * g], hi[4 * g + 1], hi[4 * g + 2], hi[4 * g + 3]), pol);
    if (SPLIT && d.nhalf == 2)
      tc::st_v4_hint(p + d.F * 32u + g * 256, make_uint4(lo[4 * g], lo[4 * g + 1], lo[4 * g + 2], lo[4 * g + 3]), pol);
  }
}

}  // namespace eng
}  // namespace scnerf
