// Table-driven tcgen05 "slab engine" shared by the fused field kernels (forward, dgrad).
//
// A tile (128 samples) is processed as a fixed sequence of SLABS.  One slab = one K=16 step of one
// GEMM stage: B operand = an [N x 16] bf16 weight slab (hi [, lo]) that the producer warp streams
// from the packed weight image with one 1-D bulk TMA copy into a ring slot; A operand = 16 columns of
// the activation, either in TMEM (TS-mode MMA) or in a shared-memory canonical image (SS-mode).
// The same table drives the producer (how many bytes per slot) and the single MMA-issuing thread.
#pragma once
#include "common.cuh"
#include "tc_prims.cuh"

namespace scnerf {
namespace eng {

constexpr int TILE_M = 128;
constexpr int MAX_SLABS = 256;

enum : uint8_t { A_TMEM = 0, A_SMEM = 1 };
enum : uint8_t { F_ZERO_ACC = 1, F_STAGE_END = 2, F_HI_ONLY_A = 4 };

struct SlabDef {
  uint16_t n;          // rows of the B slab (= GEMM N), multiple of 16
  uint16_t acc_col;    // accumulator column offset in TMEM
  uint16_t a_off;      // A_TMEM: column offset of this k16 inside the A_hi / A_lo regions (8 per k16)
                       // A_SMEM: byte offset / 16 of the hi image slab inside the smem A area
  uint16_t a_lo_delta; // A_SMEM: byte distance / 16 from the hi to the lo image
  uint8_t a_kind;
  uint8_t flags;
  uint16_t pad;
};
static_assert(sizeof(SlabDef) == 12, "SlabDef layout");

struct Plan {
  SlabDef slab[MAX_SLABS];
  int n_slabs;
  int n_stages;
};

// bytes one slab occupies in the weight image / ring slot
template <int NSPLIT> __host__ __device__ __forceinline__ uint32_t slab_bytes(const SlabDef& d) {
  return (uint32_t)d.n * 32u * (NSPLIT == 3 ? 2u : 1u);
}

struct Ring {
  uint8_t* base;       // NSLOT x SLOT_BYTES
  uint64_t* full;
  uint64_t* empty;
};

// ---- producer: one elected thread streams the weight image once per tile --------------------------
template <int NSPLIT, int NSLOT, int SLOT_BYTES>
__device__ __forceinline__ void producer_loop(const Plan& plan, const uint8_t* __restrict__ wimg,
                                              const Ring& ring, int num_tiles) {
  uint32_t n = 0;
  for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
    const uint8_t* src = wimg;
#pragma unroll 1
    for (int i = 0; i < plan.n_slabs; ++i, ++n) {
      const uint32_t bytes = slab_bytes<NSPLIT>(plan.slab[i]);
      const uint32_t idx = n % NSLOT, ph = (n / NSLOT) & 1;
      tc::mbar_wait(&ring.empty[idx], ph ^ 1);
      tc::mbar_arrive_expect_tx(&ring.full[idx], bytes);
      tc::bulk_g2s(ring.base + idx * SLOT_BYTES, src, bytes, &ring.full[idx]);
      src += bytes;
    }
  }
}

// ---- MMA issuer: one thread -------------------------------------------------------------------------
// tmem_acc: TMEM base of the accumulators; tmem_ahi / tmem_alo: TMEM bases of the A operand halves;
// smem_a: shared address (u32) of the smem A area.
template <int NSPLIT, int NSLOT, int SLOT_BYTES>
__device__ __forceinline__ void mma_loop(const Plan& plan, const Ring& ring, uint64_t* a_ready,
                                         uint64_t* acc_full, uint32_t tmem_acc, uint32_t tmem_ahi,
                                         uint32_t tmem_alo, uint32_t smem_a, int num_tiles) {
  constexpr bool SPLIT = NSPLIT == 3;
  const uint32_t ring_addr = tc::smem_u32(ring.base);
  uint32_t n = 0, q = 0;
  for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
    bool stage_start = true;
#pragma unroll 1
    for (int i = 0; i < plan.n_slabs; ++i, ++n) {
      const SlabDef d = plan.slab[i];
      if (stage_start) {          // A operand of this stage written, accumulator drained
        tc::mbar_wait(a_ready, q & 1);
        tc::tc_fence_after();
        ++q;
        stage_start = false;
      }
      const uint32_t idx = n % NSLOT, ph = (n / NSLOT) & 1;
      tc::mbar_wait(&ring.full[idx], ph);
      tc::tc_fence_after();
      const uint32_t slot = ring_addr + idx * SLOT_BYTES;
      const uint32_t idesc = tc::idesc_bf16_f32(TILE_M, d.n);
      const uint32_t lbo_b = (uint32_t)d.n * 16u;
      const uint64_t b_hi = tc::smem_desc(slot, lbo_b, 128);
      const uint64_t b_lo = tc::smem_desc(slot + (uint32_t)d.n * 32u, lbo_b, 128);
      const uint32_t acc = tmem_acc + d.acc_col;
      const uint32_t zero = (d.flags & F_ZERO_ACC) ? 0u : 1u;
      if (d.a_kind == A_TMEM) {
        tc::mma_ts(acc, tmem_ahi + d.a_off, b_hi, idesc, zero);
        if (SPLIT) {
          if (!(d.flags & F_HI_ONLY_A)) tc::mma_ts(acc, tmem_alo + d.a_off, b_hi, idesc, 1);
          tc::mma_ts(acc, tmem_ahi + d.a_off, b_lo, idesc, 1);
        }
      } else {
        const uint32_t a_addr = smem_a + (uint32_t)d.a_off * 16u;
        const uint64_t a_hi = tc::smem_desc(a_addr, 2048, 128);
        tc::mma_ss(acc, a_hi, b_hi, idesc, zero);
        if (SPLIT) {
          if (!(d.flags & F_HI_ONLY_A)) {
            const uint64_t a_lo = tc::smem_desc(a_addr + (uint32_t)d.a_lo_delta * 16u, 2048, 128);
            tc::mma_ss(acc, a_lo, b_hi, idesc, 1);
          }
          tc::mma_ss(acc, a_hi, b_lo, idesc, 1);
        }
      }
      tc::tc_commit(&ring.empty[idx]);
      if (d.flags & F_STAGE_END) {
        tc::tc_commit(acc_full);
        stage_start = true;
      }
    }
  }
}

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

// split 32 fp32 values (already activated) into packed hi / lo bf16 pairs
template <bool SPLIT>
__device__ __forceinline__ void split32(const float (&f)[32], uint32_t (&hi)[16], uint32_t (&lo)[16]) {
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    hi[j] = cvt_bf16x2(f[2 * j], f[2 * j + 1]);
    if (SPLIT) lo[j] = cvt_bf16x2(f[2 * j] - bf16lo_f(hi[j]), f[2 * j + 1] - bf16hi_f(hi[j]));
  }
}

// ---- tile-image dumps (the wgrad kernel's operand format) ---------------------------------------------
// A [128 samples x F features] bf16 tile is stored as 8 K16-slabs (16 samples each); inside a slab the
// layout is the UMMA canonical MN-major no-swizzle image:
//   byte(mn, k) = (mn/8)*256 + ((k/8)&1)*128 + (k&7)*16 + (mn&7)*2        (LBO = 128, SBO = 256)
// global: [tile][slab k/16][half: hi, lo][F*32 bytes]
struct ImgDump {
  uint8_t* base;       // NULL = disabled
  uint32_t F;          // features (multiple of 8)
  uint32_t nhalf;      // 1 (hi) or 2 (hi, lo)
  __device__ __forceinline__ size_t tile_bytes() const { return (size_t)F * 256u * nhalf; }
  // address of the 16-byte chunk (8 features starting at mn0, sample k of tile `tile`, half h)
  __device__ __forceinline__ uint8_t* chunk(int tile, uint32_t k, uint32_t mn0, uint32_t h) const {
    return base + (size_t)tile * tile_bytes() + (size_t)(k >> 4) * (F * 32u * nhalf) + (size_t)h * (F * 32u) +
           (mn0 >> 3) * 256u + ((k >> 3) & 1u) * 128u + (k & 7u) * 16u;
  }
};
// store 32 consecutive features [c0, c0+32) of sample k (packed pairs) into the image
template <bool SPLIT>
__device__ __forceinline__ void dump32(const ImgDump& d, int tile, uint32_t k, uint32_t c0,
                                       const uint32_t (&hi)[16], const uint32_t (&lo)[16]) {
#pragma unroll
  for (int g = 0; g < 4; ++g) {
    *reinterpret_cast<uint4*>(d.chunk(tile, k, c0 + 8 * g, 0)) =
        make_uint4(hi[4 * g], hi[4 * g + 1], hi[4 * g + 2], hi[4 * g + 3]);
    if (SPLIT && d.nhalf == 2)
      *reinterpret_cast<uint4*>(d.chunk(tile, k, c0 + 8 * g, 1)) =
          make_uint4(lo[4 * g], lo[4 * g + 1], lo[4 * g + 2], lo[4 * g + 3]);
  }
}

}  // namespace eng
}  // namespace scnerf
