// Tensor-core (tcgen05) field path.
//  - pack_canon_kernel: fp32 nn.Linear weights -> bf16 (hi[,lo]) images in the UMMA canonical
//    no-swizzle K-major layout, so a weight K-slab is ONE contiguous byte range that a 1-D bulk
//    async copy (TMA, cp.async.bulk) can drop into shared memory with no tensor map.
//  - tc_selftest_kernel: single-tile GEMM that pins the descriptor encodings on hardware.
//  - field_tc_fwd: fused PE + MLP forward (see field_tc_fused.cuh).
#pragma once
#include "common.cuh"
#include "field_simt.cuh"
#include "tc_prims.cuh"
#include "field_tc_fused.cuh"
#include "field_tc_fwd_pipe.cuh"
#include "field_tc_dgrad.cuh"
#include "field_tc_dgrad_pipe.cuh"
#include "field_tc_wgrad.cuh"

namespace scnerf {

// W[rows, cols] fp32 (row stride ld, first source column col0) -> canonical image with rows_pad
// rows and k_pad columns (zero padded).  lo = bf16(W - float(hi)) for the split-bf16 path.
__global__ void __launch_bounds__(256) pack_canon_kernel(const float* __restrict__ W, int64_t ld,
                                                         int rows, int cols, int col0, int rows_pad,
                                                         int k_pad, uint8_t* __restrict__ hi,
                                                         uint8_t* __restrict__ lo) {
  int g = blockIdx.x * blockDim.x + threadIdx.x;
  int nchunk = k_pad >> 3;
  if (g >= rows_pad * nchunk) return;
  int row = g % rows_pad, kc = g / rows_pad;
  float v[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    int k = kc * 8 + j;
    v[j] = (row < rows && k < cols) ? W[(int64_t)row * ld + col0 + k] : 0.f;
  }
  uint32_t h[4], l[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    __nv_bfloat16 h0 = __float2bfloat16_rn(v[2 * j]), h1 = __float2bfloat16_rn(v[2 * j + 1]);
    h[j] = (uint32_t)__bfloat16_as_ushort(h0) | ((uint32_t)__bfloat16_as_ushort(h1) << 16);
    l[j] = tc::pack_bf16(v[2 * j] - __bfloat162float(h0), v[2 * j + 1] - __bfloat162float(h1));
  }
  uint32_t off = tc::canon_off(row, kc * 8, rows_pad);
  *reinterpret_cast<uint4*>(hi + off) = make_uint4(h[0], h[1], h[2], h[3]);
  if (lo) *reinterpret_cast<uint4*>(lo + off) = make_uint4(l[0], l[1], l[2], l[3]);
}

// D[128,N] = A[128,K] * B[N,K]^T from canonical bf16 images in global memory.
// variant bit0: swap the LBO/SBO descriptor fields (diagnostic); bit1: stage with cp.async.bulk;
// bit2: A operand from TMEM (TS-mode MMA), written there with tcgen05.st from registers.
__global__ void __launch_bounds__(128) tc_selftest_kernel(const uint8_t* __restrict__ A_img,
                                                          const uint8_t* __restrict__ B_img,
                                                          float* __restrict__ D, int N, int K,
                                                          int variant) {
  extern __shared__ __align__(1024) uint8_t tc_smem[];
  __shared__ __align__(8) uint64_t bar_mma, bar_tma;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t a_bytes = 128u * K * 2u, b_bytes = (uint32_t)N * K * 2u;
  uint8_t* sA = tc_smem;
  uint8_t* sB = tc_smem + a_bytes;
  uint32_t ncols = 32;
  while (ncols < (uint32_t)N) ncols <<= 1;
  if (variant & 4) ncols = 512;   // D in [0,256), A in [256, 256+K/2)
  if (tid == 0) {
    tc::mbar_init(&bar_mma, 1);
    tc::mbar_init(&bar_tma, 1);
    tc::fence_mbar_init();
  }
  __syncthreads();
  if (warp == 0) tc::tmem_alloc(&tmem_base_s, ncols);
  if (variant & 2) {
    if (tid == 0) {
      tc::mbar_arrive_expect_tx(&bar_tma, a_bytes + b_bytes);
      for (uint32_t o = 0; o < a_bytes; o += 16384) tc::bulk_g2s(sA + o, A_img + o, min(16384u, a_bytes - o), &bar_tma);
      for (uint32_t o = 0; o < b_bytes; o += 16384) tc::bulk_g2s(sB + o, B_img + o, min(16384u, b_bytes - o), &bar_tma);
    }
    tc::mbar_wait(&bar_tma, 0);
  } else {
    for (uint32_t o = tid * 16; o < a_bytes; o += 128 * 16)
      *reinterpret_cast<uint4*>(sA + o) = *reinterpret_cast<const uint4*>(A_img + o);
    for (uint32_t o = tid * 16; o < b_bytes; o += 128 * 16)
      *reinterpret_cast<uint4*>(sB + o) = *reinterpret_cast<const uint4*>(B_img + o);
    tc::fence_proxy_async();
  }
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem = tmem_base_s;
  if (variant & 4) {
    // row (= TMEM lane) r of A as packed bf16 pairs -> TMEM columns [256, 256+K/2)
    const int r = warp * 32 + lane;
    for (int c0 = 0; c0 < K / 2; c0 += 16) {
      uint32_t v[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        int k = 2 * (c0 + j);
        uint32_t lo16 = *reinterpret_cast<const uint16_t*>(sA + tc::canon_off(r, k, 128));
        uint32_t hi16 = *reinterpret_cast<const uint16_t*>(sA + tc::canon_off(r, k + 1, 128));
        v[j] = lo16 | (hi16 << 16);
      }
      tc::tmem_st16(tmem + ((uint32_t)(warp * 32) << 16) + 256 + c0, v);
    }
    tc::tmem_st_wait();
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
  }
  if (tid == 0) {
    uint32_t lboA = 128 * 16, lboB = (uint32_t)N * 16, sboA = 128, sboB = 128;
    const uint32_t stepA = 2 * lboA, stepB = 2 * lboB;   // one K=16 step = two 8-element chunks
    if (variant & 1) { uint32_t t = lboA; lboA = sboA; sboA = t; t = lboB; lboB = sboB; sboB = t; }
    const uint32_t idesc = tc::idesc_bf16_f32(128, (uint32_t)N);
    for (int k = 0; k < K / 16; ++k) {
      uint64_t ad = tc::smem_desc(tc::smem_u32(sA) + k * stepA, lboA, sboA);
      uint64_t bd = tc::smem_desc(tc::smem_u32(sB) + k * stepB, lboB, sboB);
      if (variant & 4) tc::mma_ts(tmem, tmem + 256 + k * 8, bd, idesc, k > 0);
      else tc::mma_ss(tmem, ad, bd, idesc, k > 0);
    }
    tc::tc_commit(&bar_mma);
  }
  tc::mbar_wait(&bar_mma, 0);
  tc::tc_fence_after();
  for (int c0 = 0; c0 < N; c0 += 32) {
    uint32_t v[32];
    tc::tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + c0, v);
    tc::tmem_ld_wait();
    float* drow = D + (int64_t)(warp * 32 + lane) * N + c0;
#pragma unroll
    for (int j = 0; j < 32; ++j)
      if (c0 + j < N) drow[j] = __uint_as_float(v[j]);
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(tmem, ncols);
}

// Tensor-pipe throughput probe: every CTA issues `iters` x 16 MMAs (M=128, N=256, K=16) on fixed
// (uninitialised) operands and reports cycles per MMA.  mode 0: SS K-major (forward X/V slabs),
// 1: TS (A in TMEM), 2: SS with both operands MN-major in the wgrad slab layout (LBO 128, SBO 256),
// 3: SS MN-major A + K-major B.
__global__ void __launch_bounds__(128) tc_mma_bench_kernel(int mode, int iters, long long* out) {
  extern __shared__ __align__(1024) uint8_t tc_smem[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ __align__(8) uint64_t fbar[8];
  __shared__ uint32_t tmem_base_s;
  __shared__ volatile int done_flag;
  const int tid = threadIdx.x, warp = tid >> 5;
  if (tid == 0) {
    tc::mbar_init(&bar, 1);
    for (int i = 0; i < 8; ++i) tc::mbar_init(&fbar[i], 1);
    done_flag = 0;
    tc::fence_mbar_init();
  }
  for (int i = tid; i < 32768 / 4; i += 128) reinterpret_cast<uint32_t*>(tc_smem)[i] = 0x3c003c00u;
  tc::fence_proxy_async();
  __syncthreads();
  if (warp == 0) tc::tmem_alloc(&tmem_base_s, 512);
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem = tmem_base_s;
  // modes 4..7: N = 128 half-slabs (the N-split pipeline's MMA shape).  4: TS only.  5: the split-bf16
  // triple TS(hi,Bhi) / SS(lo in smem,Bhi) / TS(hi,Blo).  6: 5 + a concurrent bulk-copy refill stream of
  // 8 KB per triple (8 copies in flight) from the global scratch behind out[gridDim.x].  7: 4 + refill 4 KB/MMA.
  if (tid == 32 && (mode == 6 || mode == 7)) {
    const uint8_t* gsrc = reinterpret_cast<const uint8_t*>(out + ((gridDim.x + 1) & ~1u));
    uint32_t ph = 0;
    const uint32_t bytes = 8192;
    for (int i = 0; i < 8; ++i) {
      tc::mbar_arrive_expect_tx(&fbar[i], bytes);
      tc::bulk_g2s(tc_smem + 32768 + i * 8192, gsrc + i * 8192, bytes, &fbar[i]);
    }
    while (!done_flag) {
      for (int i = 0; i < 8; ++i) {
        tc::mbar_wait(&fbar[i], ph);
        tc::mbar_arrive_expect_tx(&fbar[i], bytes);
        tc::bulk_g2s(tc_smem + 32768 + i * 8192, gsrc + i * 8192, bytes, &fbar[i]);
      }
      ph ^= 1u;
    }
    for (int i = 0; i < 8; ++i) tc::mbar_wait(&fbar[i], ph);
  }
  if (tid == 0) {
    const uint32_t sa = tc::smem_u32(tc_smem), sb = sa + 16384;
    const uint32_t idk = tc::idesc_bf16_f32(128, 256);
    const uint32_t idh = tc::idesc_bf16_f32(128, 128);
    const uint32_t idmn = idk | (1u << 15) | (1u << 16), idamn = idk | (1u << 15);
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        if (mode == 0) tc::mma_ss(tmem, eng::desc_at<2048, 128>(sa), eng::desc_at<4096, 128>(sb), idk, 1);
        else if (mode == 1) tc::mma_ts(tmem, tmem + 256 + (j & 7) * 8, eng::desc_at<4096, 128>(sb), idk, 1);
        else if (mode == 2) tc::mma_ss(tmem, eng::desc_at<128, 256>(sa), eng::desc_at<128, 256>(sb), idmn, 1);
        else if (mode == 3) tc::mma_ss(tmem, eng::desc_at<128, 256>(sa), eng::desc_at<4096, 128>(sb), idamn, 1);
        else if (mode == 4 || mode == 7) tc::mma_ts(tmem + (j & 1) * 128, tmem + 256 + (j & 7) * 8, eng::desc_at<2048, 128>(sb), idh, 1);
        else {
          const int r = j % 3;
          if (r == 1) tc::mma_ss(tmem, eng::desc_at<2048, 128>(sa + (j & 3) * 4096), eng::desc_at<2048, 128>(sb), idh, 1);
          else tc::mma_ts(tmem, tmem + 256 + (j & 7) * 8, eng::desc_at<2048, 128>(sb + (r == 2 ? 4096 : 0)), idh, 1);
        }
      }
    }
    tc::tc_commit(&bar);
    tc::mbar_wait(&bar, 0);
    const long long t1 = clock64();
    out[blockIdx.x] = t1 - t0;
    done_flag = 1;
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(tmem, 512);
}

// Compile-time MMA-shape probe (the loop body is nothing but the MMAs): N = B rows, PAT 0: TS same accumulator,
// 1: TS alternating accumulator halves, 2: SS, 3: TS/SS/TS triple (split-bf16 with the lo operand in smem),
// 4: SS/TS alternating, 5: TS/TS/TS triple.
template <int N, int PAT>
__global__ void __launch_bounds__(128) tc_mma_probe_kernel(int iters, long long* out) {
  extern __shared__ __align__(1024) uint8_t tc_smem[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5;
  if (tid == 0) { tc::mbar_init(&bar, 1); tc::fence_mbar_init(); }
  for (int i = tid; i < 32768 / 4; i += 128) reinterpret_cast<uint32_t*>(tc_smem)[i] = 0x3c003c00u;
  tc::fence_proxy_async();
  __syncthreads();
  if (warp == 0) tc::tmem_alloc(&tmem_base_s, 512);
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem = tmem_base_s;
  if (tid == 0) {
    const uint32_t sa = tc::smem_u32(tc_smem), sb = sa + 16384;
    constexpr uint32_t idn = tc::idesc_bf16_f32(128, N);
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int j = 0; j < 24; ++j) {
        const uint64_t bd = eng::desc_at<N * 16, 128>(sb + (j & 1) * 8192);
        const uint64_t ad = eng::desc_at<2048, 128>(sa + (j & 3) * 4096);
        const uint32_t ta = tmem + 256 + (j & 7) * 8;
        const uint32_t acc = tmem + ((PAT == 1 && (j & 1)) ? 128 : 0);
        constexpr bool dummy = false; (void)dummy;
        const bool ss = PAT == 2 || (PAT == 3 && j % 3 == 1) || (PAT == 4 && (j & 1) == 0);
        if (ss) tc::mma_ss(acc, ad, bd, idn, 1);
        else tc::mma_ts(acc, ta, bd, idn, 1);
      }
    }
    tc::tc_commit(&bar);
    tc::mbar_wait(&bar, 0);
    const long long t1 = clock64();
    out[blockIdx.x] = t1 - t0;
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(tmem, 512);
}
template <int N, int PAT>
inline int tc_mma_probe_launch(int iters, long long* out, int nblocks, void* stream) {
  SCNERF_LAUNCH((tc_mma_probe_kernel<N, PAT>), nblocks, 128, 32768, stream, iters, out);
  return 0;
}
template <int N>
inline int tc_mma_probe_n(int pat, int iters, long long* out, int nblocks, void* stream) {
  switch (pat) {
    case 0: return tc_mma_probe_launch<N, 0>(iters, out, nblocks, stream);
    case 1: return tc_mma_probe_launch<N, 1>(iters, out, nblocks, stream);
    case 2: return tc_mma_probe_launch<N, 2>(iters, out, nblocks, stream);
    case 3: return tc_mma_probe_launch<N, 3>(iters, out, nblocks, stream);
    case 4: return tc_mma_probe_launch<N, 4>(iters, out, nblocks, stream);
    default: return tc_mma_probe_launch<N, 5>(iters, out, nblocks, stream);
  }
}

inline int tc_mma_bench(int mode, int iters, long long* out, int nblocks, void* stream) {
  if (mode >= 256) {      // mode = 256 + (N/8 << 4) + pattern  (24 MMAs per iteration)
    const int n = ((mode - 256) >> 4) * 8, pat = mode & 15;
    switch (n) {
      case 256: return tc_mma_probe_n<256>(pat, iters, out, nblocks, stream);
      case 128: return tc_mma_probe_n<128>(pat, iters, out, nblocks, stream);
      case 64: return tc_mma_probe_n<64>(pat, iters, out, nblocks, stream);
      case 32: return tc_mma_probe_n<32>(pat, iters, out, nblocks, stream);
      default: return fail(SCNERF_ERR_ARG, "mma probe: N must be 256/128/64/32");
    }
  }
  SCNERF_CUDA(cudaFuncSetAttribute(tc_mma_bench_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 32768 + 65536));
  SCNERF_LAUNCH(tc_mma_bench_kernel, nblocks, 128, 32768 + 65536, stream, mode, iters, out);
  return 0;
}

// host side of the self-test: A[128,K], B[N,K] fp32 row-major -> D[128,N] fp32
inline int tc_selftest(const float* A, const float* B, float* D, int N, int K, int variant,
                       void* workspace, size_t workspace_bytes, void* stream) {
  SCNERF_CHECK_ARG(N % 16 == 0 && N >= 16 && N <= 256 && K % 16 == 0 && K >= 16, "selftest: bad N/K");
  size_t a_bytes = (size_t)128 * K * 2, b_bytes = (size_t)N * K * 2;
  SCNERF_CHECK_ARG(a_bytes + b_bytes <= 220 * 1024, "selftest: tile does not fit shared memory");
  SCNERF_CHECK_ARG(!(variant & 4) || (K <= 512 && K % 32 == 0), "selftest: TS variant needs K%%32==0, K<=512");
  Arena ar(workspace, workspace_bytes);
  uint8_t* Ai = ar.get<uint8_t>(a_bytes);
  uint8_t* Bi = ar.get<uint8_t>(b_bytes);
  if (!workspace || !ar.ok()) return fail(SCNERF_ERR_WORKSPACE, "selftest: workspace too small");
  SCNERF_LAUNCH(pack_canon_kernel, (unsigned)cdiv(128 * (K / 8), 256), 256, 0, stream, A, (int64_t)K, 128, K,
                0, 128, K, Ai, (uint8_t*)nullptr);
  SCNERF_LAUNCH(pack_canon_kernel, (unsigned)cdiv(N * (K / 8), 256), 256, 0, stream, B, (int64_t)K, N, K, 0, N,
                K, Bi, (uint8_t*)nullptr);
  static bool attr_set = false;
  if (!attr_set) {
    SCNERF_CUDA(cudaFuncSetAttribute(tc_selftest_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
    attr_set = true;
  }
  SCNERF_LAUNCH(tc_selftest_kernel, 1, 128, a_bytes + b_bytes, stream, Ai, Bi, D, N, K, variant);
  return 0;
}

// optional in-kernel timeline buffer for the next fused forward launch (diagnostics only)
inline long long*& tc_dbg_ptr() { static long long* p = nullptr; return p; }
inline int& tc_dbg_tiles() { static int n = 0; return n; }

inline int device_sm_count() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
  }
  return n;
}

// one-time (per device) upload of the slab plans into __device__ symbols
template <int XS = 4>
inline const eng::Plan& pipe_plan_host() {
  static eng::Plan P = fpipe::make_plan<3, XS>();   // (slab order, sizes and image offsets do not depend on NSPLIT)
  return P;
}
inline int pipe_plan_init() {
  static bool done[64] = {};
  int dev = 0;
  SCNERF_CUDA(cudaGetDevice(&dev));
  if (dev < 64 && done[dev]) return 0;
  const eng::Plan& P = pipe_plan_host<4>();
  const eng::Plan& P6 = pipe_plan_host<6>();
  static fused::PlanSrc S, S6;
  fpipe::build_plansrc<4>(S);
  fpipe::build_plansrc<6>(S6);
  if (fused::plan_image_bytes(P, 3) > TC_IMG_BYTES || fused::plan_image_bytes(P6, 3) > TC_IMG_BYTES)
    return fail(SCNERF_ERR_WORKSPACE, "TC_IMG_BYTES too small (pipelined forward)");
  SCNERF_CUDA(cudaMemcpyToSymbol(fpipe::d_plan_pipe, &P, sizeof(P)));
  SCNERF_CUDA(cudaMemcpyToSymbol(fpipe::d_plansrc_pipe, &S, sizeof(S)));
  SCNERF_CUDA(cudaMemcpyToSymbol(fpipe::d_plan_pipe6, &P6, sizeof(P6)));
  SCNERF_CUDA(cudaMemcpyToSymbol(fpipe::d_plansrc_pipe6, &S6, sizeof(S6)));
  // 3-D points: the rolled-epilogue build (ROLL = 1: stages 0-6 from one copy of the code, -0.6 ms per step, profiles/r2a_*);
  // 4-D points (NeRF++ background): the unrolled build
  SCNERF_CUDA(cudaFuncSetAttribute((fpipe::field_fwd_pipe_kernel<1, 4, 1>), cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   fpipe::Cfg<1, 4>::SMEM_BYTES));
  SCNERF_CUDA(cudaFuncSetAttribute((fpipe::field_fwd_pipe_kernel<3, 4, 1>), cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   fpipe::Cfg<3, 4>::SMEM_BYTES));
  SCNERF_CUDA(cudaFuncSetAttribute((fpipe::field_fwd_pipe_kernel<1, 6, 0>), cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   fpipe::Cfg<1, 6>::SMEM_BYTES));
  SCNERF_CUDA(cudaFuncSetAttribute((fpipe::field_fwd_pipe_kernel<3, 6, 0>), cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   fpipe::Cfg<3, 6>::SMEM_BYTES));
  if (dev < 64) done[dev] = true;
  return 0;
}
inline fused::PackSrc make_pack_src(const scnerf_mlp& m) {
  fused::PackSrc src{};
  for (int i = 0; i < 8; ++i) {
    src.w[i] = m.pts_w[i]; src.b[i] = m.pts_b[i];
    src.ld[i] = (i == 0) ? m.input_ch : (i == 5 ? m.input_ch + 256 : 256);      // 63 / 319, or 84 / 340 (4-D points)
  }
  src.w[8] = m.feature_w; src.b[8] = m.feature_b; src.ld[8] = 256;
  src.w[9] = m.views_w; src.b[9] = m.views_b; src.ld[9] = 283;
  src.alpha_w = m.alpha_w; src.alpha_b = m.alpha_b; src.rgb_w = m.rgb_w; src.rgb_b = m.rgb_b;
  return src;
}

// bf16 tile images kept by the tensor-core training path (one set of forward images per pass;
// the dgrad images and per-sample geometry gradients are shared between passes)
struct TcFwdImages { eng::ImgDump x, v, h[8], feat, hv; uint4* relu_bits = nullptr; };
struct TcBwdBufs {
  eng::ImgDump dz[8], dfeat, dzv;
  float *g_pts = nullptr, *g_vd = nullptr;
  uint8_t* wimg = nullptr;      // transposed weight slabs for the dgrad kernel
};
inline eng::ImgDump img_alloc(Arena& ar, int64_t tiles, uint32_t F, uint32_t nh) {
  eng::ImgDump d;
  d.F = F; d.nhalf = nh;
  d.base = ar.get<uint8_t>((size_t)tiles * F * 256u * nh);
  return d;
}
inline void tc_fwd_images_alloc(Arena& ar, int64_t P, int nsplit, TcFwdImages& I, int pts_dim = 3) {
  const int64_t T = cdiv(P, 128);
  const uint32_t nh = nsplit == 3 ? 2 : 1;
  I.x = img_alloc(ar, T, pts_dim == 4 ? 96 : 64, nh); I.v = img_alloc(ar, T, 32, nh);
  for (int i = 0; i < 8; ++i) I.h[i] = img_alloc(ar, T, 256, nh);
  I.feat = img_alloc(ar, T, 256, nh); I.hv = img_alloc(ar, T, 128, nh);
  I.relu_bits = ar.get<uint4>((size_t)T * 9 * 2 * 128);
}
inline void tc_bwd_bufs_alloc(Arena& ar, int64_t P, int nsplit, TcBwdBufs& G) {
  const int64_t T = cdiv(P, 128);
  const uint32_t nh = nsplit == 3 ? 2 : 1;
  for (int i = 0; i < 8; ++i) G.dz[i] = img_alloc(ar, T, 256, nh);
  G.dfeat = img_alloc(ar, T, 256, nh); G.dzv = img_alloc(ar, T, 128, nh);
  G.g_pts = ar.get<float>(P * 4); G.g_vd = ar.get<float>(P * 3);     // g_pts: [P,3], or [P,4] for 4-D points
  G.wimg = ar.get<uint8_t>(TC_IMG_BYTES);
}

// fused composite: a group of G consecutive 128-sample tiles must hold whole rays of S samples (G <= 3: 6 KB of smem)
inline int comp_group_tiles(int S) {
  int g = S, b = 128;
  while (b) { int t = g % b; g = b; b = t; }      // gcd(S, 128)
  return S / g;
}
inline bool comp_fusable(int S) {
  static int on = -1;
  if (on < 0) { const char* e = getenv("SCNERF_FUSE_COMPOSITE"); on = e ? (atoi(e) != 0) : 1; }
  return on && S >= 1 && comp_group_tiles(S) <= fpipe::Cfg<3, 4>::COMP_MAX_G;
}
template <int NSPLIT, int XS = 4>
inline int field_tc_fwd_impl(const scnerf_mlp& m, const float* rays, int ray_cols, const float* z,
                             const float* pts, const float* viewdirs, int64_t N, int S,
                             const FieldBufs& B, float* raw, void* stream, const TcFwdImages* imgs = nullptr,
                             const CompositeArgs* comp = nullptr, bool* comp_done = nullptr, bool raw_needed = true) {
  using namespace fused;
  static_assert(C_TOTAL <= (int)TC_CBUF_FLOATS, "TC_CBUF_FLOATS too small");
  int rc = pipe_plan_init();
  if (rc) return rc;
  PackSrc src = make_pack_src(m);
  SCNERF_LAUNCH((fpipe::pack_pipe_kernel<NSPLIT, XS>), dim3(1, (unsigned)pipe_plan_host<XS>().n_slabs), 256, 0, stream, src, B.tc_img);
  SCNERF_LAUNCH(pack_consts_kernel, (unsigned)cdiv(C_TOTAL, 256), 256, 0, stream, src, B.tc_cbuf);
  Args a{};
  a.rays = rays; a.ray_cols = ray_cols; a.z = z; a.pts = pts; a.viewdirs = viewdirs;
  a.P = N * S; a.S = S; a.wimg = B.tc_img; a.cbuf = B.tc_cbuf; a.raw = raw;
  a.num_tiles = (int)cdiv(a.P, TILE_M);
  if (imgs) {                 // training with the tensor-core backward: bf16 tile images
    a.img_x = imgs->x; a.img_v = imgs->v;
    for (int s = 0; s < 8; ++s) a.img_out[s] = imgs->h[s];
    a.img_out[8] = imgs->feat; a.img_out[9] = imgs->hv;
    a.relu_bits = imgs->relu_bits;
  }
  if (XS == 4 && B.keep_all && B.X5) {   // training with the fp32 CUDA-core backward: fp32 row-major layer inputs
    for (int s = 0; s < 8; ++s) {
      a.dump[s] = (s == 4) ? B.X5 + 63 : B.H[s];
      a.dump_ld[s] = (s == 4) ? (int)B.ldx5 : 256;
    }
    a.dump[8] = B.F; a.dump_ld[8] = (int)B.ldf;
    a.dump[9] = B.HV; a.dump_ld[9] = 128;
    a.dump_pe = B.X5; a.dump_pe_ld = (int)B.ldx5;
    a.dump_ped = B.F + 256; a.dump_ped_ld = (int)B.ldf;
  }
  a.dbg = tc_dbg_ptr(); a.dbg_tiles = tc_dbg_tiles();
  if (comp_done) *comp_done = false;
  if (comp && XS == 4 && comp_fusable(S)) {   // alpha-composite in the epilogue (inference): raw may stay out of HBM
    a.comp_on = 1; a.comp_G = comp_group_tiles(S); a.comp = *comp;
    if (!raw_needed) a.raw = nullptr;                 // nobody reads raw: it never reaches HBM
    if (comp_done) *comp_done = true;
  }
  if (!raw) return fail(SCNERF_ERR_ARG, "field forward: null raw output");
  int grid = std::min(device_sm_count(), a.num_tiles);
  SCNERF_LAUNCH((fpipe::field_fwd_pipe_kernel<NSPLIT, XS, (XS == 4 ? 1 : 0)>), grid, 320, (fpipe::Cfg<NSPLIT, XS>::SMEM_BYTES),
                stream, a);
  return 0;
}

inline int field_tc_fwd(const scnerf_mlp& m, int precision, const float* rays, int ray_cols,
                        const float* z, const float* pts, const float* viewdirs, int64_t N, int S,
                        const FieldBufs& B, float* raw, void* stream, const TcFwdImages* imgs = nullptr,
                        const CompositeArgs* comp = nullptr, bool* comp_done = nullptr, bool raw_needed = true) {
  if (comp_done) *comp_done = false;
  if (!(m.D == 8 && m.W == 256 && m.skip == 4 && m.use_viewdirs && m.L_pos == 10 && m.L_dir == 4))
    return fail(SCNERF_ERR_UNSUPPORTED,
                "tensor-core field path is specialised for the 8x256, skip-4, use_viewdirs network "
                "(multires 10/4); use precision fp32 for other shapes");
  if (rays && ray_cols != 11) return fail(SCNERF_ERR_ARG, "tensor-core field path needs 11-column rays");
  if (m.pts_dim == 4) {    // NeRF++ background network: explicit (x, y, z, 1/r) points, 84-channel encoding
    if (!pts || !viewdirs) return fail(SCNERF_ERR_ARG, "tensor-core field path: 4-D points must be given explicitly");
    if (precision == SCNERF_PRECISION_BF16X3)
      return field_tc_fwd_impl<3, 6>(m, nullptr, 0, nullptr, pts, viewdirs, N, S, B, raw, stream, imgs);
    return field_tc_fwd_impl<1, 6>(m, nullptr, 0, nullptr, pts, viewdirs, N, S, B, raw, stream, imgs);
  }
  if (precision == SCNERF_PRECISION_BF16X3)
    return field_tc_fwd_impl<3>(m, rays, ray_cols, z, pts, viewdirs, N, S, B, raw, stream, imgs, comp, comp_done, raw_needed);
  return field_tc_fwd_impl<1>(m, rays, ray_cols, z, pts, viewdirs, N, S, B, raw, stream, imgs, comp, comp_done, raw_needed);
}

// ---- tensor-core backward: dgrad chain -> per-ray reduce -> wgrad pass -> head gradients -----------
inline int bwd_plan_init() {
  static bool done[64] = {};
  int dev = 0;
  SCNERF_CUDA(cudaGetDevice(&dev));
  if (dev < 64 && done[dev]) return 0;
  SCNERF_CUDA(cudaFuncSetAttribute(wgrad::field_wgrad_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   wgrad::Cfg<1>::SMEM_BYTES));
  SCNERF_CUDA(cudaFuncSetAttribute(wgrad::field_wgrad_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   wgrad::Cfg<3>::SMEM_BYTES));
  if (dev < 64) done[dev] = true;
  return 0;
}
// N-half pipelined dgrad chain: XN = 64 (3-D points) and XN = 96 (4-D points of the NeRF++ background network)
template <int XN>
inline const eng::Plan& dpipe_plan_host() {
  static eng::Plan P = dpipe::make_plan<XN>();
  return P;
}
template <int XN>
inline int dpipe_plan_init() {
  static bool done[64] = {};
  int dev = 0;
  SCNERF_CUDA(cudaGetDevice(&dev));
  if (dev < 64 && done[dev]) return 0;
  const eng::Plan& P = dpipe_plan_host<XN>();
  static fused::PlanSrc S;
  dpipe::build_plansrc<XN>(S);
  if (fused::plan_image_bytes(P, 3) > TC_IMG_BYTES) return fail(SCNERF_ERR_WORKSPACE, "TC_IMG_BYTES too small (pipelined dgrad)");
  if (XN == 64) {
    SCNERF_CUDA(cudaMemcpyToSymbol(dpipe::d_plan_dpipe, &P, sizeof(P)));
    SCNERF_CUDA(cudaMemcpyToSymbol(dpipe::d_plansrc_dpipe, &S, sizeof(S)));
  } else {
    SCNERF_CUDA(cudaMemcpyToSymbol(dpipe::d_plan_dpipe96, &P, sizeof(P)));
    SCNERF_CUDA(cudaMemcpyToSymbol(dpipe::d_plansrc_dpipe96, &S, sizeof(S)));
  }
  SCNERF_CUDA(cudaFuncSetAttribute((dpipe::field_dgrad_pipe_kernel<1, XN>), cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   (dpipe::Cfg<1, XN>::SMEM_BYTES)));
  SCNERF_CUDA(cudaFuncSetAttribute((dpipe::field_dgrad_pipe_kernel<3, XN>), cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   (dpipe::Cfg<3, XN>::SMEM_BYTES)));
  if (dev < 64) done[dev] = true;
  return 0;
}

inline void wgrad_unit(wgrad::Unit& u, int a_img, int a_half, int b_img, int n, int acc_col, float* out, int ld,
                       int rows_valid, int cols_valid) {
  u.a_img = a_img; u.a_half = a_half; u.b_img = b_img; u.n = n; u.acc_col = acc_col; u.out = out; u.ld = ld;
  u.rows_valid = rows_valid; u.cols_valid = cols_valid;
}

// XN = 64: 3-D points formed from (rays, z), d_rays reduced per ray.
// XN = 96: explicit 4-D points (NeRF++ background): d_pts = G.g_pts [P,4] stays per sample, d_viewdirs[N,3] +=.
template <int NSPLIT, int XN = 64>
inline int field_tc_bwd_impl(const scnerf_mlp& m, const scnerf_mlp& g, const float* rays, int ray_cols,
                             const float* z, int64_t N, int S, const FieldBufs& B, const TcFwdImages& I,
                             const TcBwdBufs& G, const float* g_raw, float* d_rays, void* stream,
                             const float* pts = nullptr, const float* viewdirs = nullptr,
                             float* d_viewdirs = nullptr) {
  constexpr int IN_CH = XN == 64 ? 63 : 84;
  int rc = bwd_plan_init();
  if (rc) return rc;
  const int64_t P = N * S;
  const int T = (int)cdiv(P, 128);
  fused::PackSrc src = make_pack_src(m);
  rc = dpipe_plan_init<XN>();
  if (rc) return rc;
  SCNERF_LAUNCH((dpipe::pack_dpipe_kernel<NSPLIT, XN>), dim3(1, (unsigned)dpipe_plan_host<XN>().n_slabs), 256, 0, stream, src, G.wimg);
  SCNERF_LAUNCH(fused::pack_consts_kernel, (unsigned)cdiv(fused::C_TOTAL, 256), 256, 0, stream, src, B.tc_cbuf);
  dgrad::Args a{};
  a.rays = rays; a.ray_cols = ray_cols; a.z = z; a.P = P; a.S = S; a.num_tiles = T;
  a.pts = pts; a.viewdirs = viewdirs;
  a.g_raw = g_raw; a.wimg = G.wimg; a.cbuf = B.tc_cbuf;
  for (int i = 0; i < 8; ++i) a.out_dz[i] = G.dz[i];
  a.relu_bits = I.relu_bits; a.out_dfeat = G.dfeat; a.out_dzv = G.dzv; a.g_pts = G.g_pts; a.g_vd = G.g_vd;
  SCNERF_LAUNCH((dpipe::field_dgrad_pipe_kernel<NSPLIT, XN>), std::min(device_sm_count(), T), 320,
                (dpipe::Cfg<NSPLIT, XN>::SMEM_BYTES), stream, a);
  if (XN == 96) {
    if (d_viewdirs)
      SCNERF_LAUNCH(dgrad::reduce_vd_grad_kernel, (unsigned)cdiv(N, 4), 128, 0, stream, G.g_vd, N, S, d_viewdirs);
  } else if (d_rays)
    SCNERF_LAUNCH(dgrad::reduce_pts_grad_kernel, (unsigned)cdiv(N, 4), 128, 0, stream, G.g_pts, G.g_vd, z, N, S,
                  ray_cols, d_rays);
  // ---- wgrad jobs --------------------------------------------------------------------------------------
  wgrad::Args w{};
  w.num_tiles = T; w.P = P;
  {
    // CTAs per job.  Round 1 set them in proportion to the measured cycles per K=16 slot of each job; round 2 re-balanced
    // them on the per-job end times of the kernel's own timeline (tools/timeline_wgrad.py): on 148 SMs
    // {16, 14 x 7, 20, 14} ends every job within 5 % (the feature job J8 also reduces d(alpha_linear.weight) and has the
    // widest spread; J0 issues 12 N=64 MMAs per slot): 3.62 -> 3.38 ms per step (profiles/r2zz_wgrad_alloc*.txt).
    static const int weight[wgrad::NJOBS] = {16, 14, 14, 14, 14, 14, 14, 14, 20, 14};
    static int ncta_env = -1;     // experiment knob: wgrad on fewer SMs (is it still HBM-bound?  profiles/README.md, round 2)
    if (ncta_env < 0) { const char* e = getenv("SCNERF_WGRAD_CTAS"); ncta_env = e ? atoi(e) : 0; }
    const int ncta = std::max(ncta_env > 0 ? std::min(ncta_env, device_sm_count()) : device_sm_count(), wgrad::NJOBS);
    int wsum = 0, used = 0, n[wgrad::NJOBS];
    for (int j = 0; j < wgrad::NJOBS; ++j) wsum += weight[j];
    for (int j = 0; j < wgrad::NJOBS; ++j) { n[j] = std::max(1, ncta * weight[j] / wsum); used += n[j]; }
    static const int order[wgrad::NJOBS] = {8, 0, 1, 2, 3, 4, 5, 6, 7, 9};     // leftovers (other SM counts): heaviest jobs first
    for (int k = 0; used < ncta; k = (k + 1) % wgrad::NJOBS) { ++n[order[k]]; ++used; }
    {   // experiment knob: explicit CTAs per job, "a,b,...,j" (10 numbers summing to at most the SM count)
      static int alloc_env[wgrad::NJOBS] = {-1};
      if (alloc_env[0] == -1) {
        alloc_env[0] = 0;
        if (const char* e = getenv("SCNERF_WGRAD_ALLOC")) {
          int v[wgrad::NJOBS], k = 0, sum = 0;
          for (const char* q = e; *q && k < wgrad::NJOBS; ++k) { v[k] = atoi(q); sum += v[k]; while (*q && *q != ',') ++q; if (*q == ',') ++q; }
          if (k == wgrad::NJOBS && sum <= device_sm_count()) for (int j = 0; j < wgrad::NJOBS; ++j) alloc_env[j] = std::max(1, v[j]);
        }
      }
      if (alloc_env[0] > 0 && ncta_env <= 0) for (int j = 0; j < wgrad::NJOBS; ++j) n[j] = alloc_env[j];
    }
    w.cta0[0] = 0;
    for (int j = 0; j < wgrad::NJOBS; ++j) w.cta0[j + 1] = w.cta0[j] + n[j];
  }
  {
    static int pf = -1;   // tuning knob (default off: measured slower on B200, profiles/README.md)
    if (pf < 0) { const char* e = getenv("SCNERF_WGRAD_L2_PREFETCH"); pf = e ? atoi(e) : 0; }
    w.l2_prefetch_slots = pf;
  }
  {
    static int min_tiles = -1;   // debug timeline: only launches with at least this many tiles write it
    if (min_tiles < 0) { const char* e = getenv("SCNERF_DBG_MIN_TILES"); min_tiles = e ? atoi(e) : 0; }
    static int dbg_cta = -1;
    if (dbg_cta < 0) { const char* e = getenv("SCNERF_DBG_CTA"); dbg_cta = e ? atoi(e) : 0; }
    w.dbg = T >= min_tiles ? tc_dbg_ptr() : nullptr; w.dbg_slots = tc_dbg_tiles() * 8; w.dbg_cta = dbg_cta;
  }
  auto std_job = [&](wgrad::Job& J, const eng::ImgDump& A, const eng::ImgDump& Bi, float* dW, int ld, int col0,
                     int cols_valid, float* db) {
    J.a[0] = A; J.na = 1; J.b[0] = Bi; J.nb = 1; J.nu = 2; J.db = db;
    wgrad_unit(J.u[0], 0, 0, 0, (int)Bi.F, 0, dW + col0, ld, 128, cols_valid);
    wgrad_unit(J.u[1], 0, 1, 0, (int)Bi.F, 256, dW + (int64_t)128 * ld + col0, ld, 128, cols_valid);
  };
  {   // J0: layer 0 (dZ0 x X) and the skip columns of layer 5 (dZ5 x X)
    wgrad::Job& J = w.job[0];
    J.a[0] = G.dz[0]; J.a[1] = G.dz[5]; J.na = 2; J.b[0] = I.x; J.nb = 1; J.nu = 4; J.db = g.pts_b[0];
    wgrad_unit(J.u[0], 0, 0, 0, XN, 0, g.pts_w[0], IN_CH, 128, IN_CH);
    wgrad_unit(J.u[1], 0, 1, 0, XN, XN, g.pts_w[0] + 128 * IN_CH, IN_CH, 128, IN_CH);
    wgrad_unit(J.u[2], 1, 0, 0, XN, 2 * XN, g.pts_w[5], IN_CH + 256, 128, IN_CH);
    wgrad_unit(J.u[3], 1, 1, 0, XN, 3 * XN, g.pts_w[5] + 128 * (IN_CH + 256), IN_CH + 256, 128, IN_CH);
  }
  for (int l = 1; l <= 4; ++l) std_job(w.job[l], G.dz[l], I.h[l - 1], g.pts_w[l], 256, 0, 256, g.pts_b[l]);
  std_job(w.job[5], G.dz[5], I.h[4], g.pts_w[5], IN_CH + 256, IN_CH, 256, g.pts_b[5]);
  std_job(w.job[6], G.dz[6], I.h[5], g.pts_w[6], 256, 0, 256, g.pts_b[6]);
  std_job(w.job[7], G.dz[7], I.h[6], g.pts_w[7], 256, 0, 256, g.pts_b[7]);
  std_job(w.job[8], G.dfeat, I.h[7], g.feature_w, 256, 0, 256, g.feature_b);
  w.job[8].g_raw = g_raw; w.job[8].dw_alpha = g.alpha_w;
  {   // J9: views layer: dZ_v (128 rows) x [feature | PE(dir)]
    wgrad::Job& J = w.job[9];
    J.a[0] = G.dzv; J.na = 1; J.b[0] = I.feat; J.b[1] = I.v; J.nb = 2; J.nu = 2; J.db = g.views_b;
    wgrad_unit(J.u[0], 0, 0, 0, 256, 0, g.views_w, 283, 128, 256);
    wgrad_unit(J.u[1], 0, 0, 1, 32, 256, g.views_w + 256, 283, 128, 27);
  }
  SCNERF_LAUNCH((wgrad::field_wgrad_kernel<NSPLIT>), (unsigned)w.cta0[wgrad::NJOBS], 320,
                wgrad::Cfg<NSPLIT>::SMEM_BYTES, stream, w);
  static int head_ctas = -1;    // CTAs per SM of the head-gradient kernel (each CTA ends with 388 atomics onto the same addresses)
  if (head_ctas < 0) { const char* e = getenv("SCNERF_HEAD_WGRAD_CTAS_PER_SM"); head_ctas = e ? std::max(1, atoi(e)) : 4; }   // 4: profiles/r2v_head_wgrad_grid.txt (8: 0.27 ms, 4: 0.20, 2: 0.22, 16: 0.42)
  SCNERF_LAUNCH((wgrad::head_wgrad_img_kernel<(NSPLIT == 3 ? 2 : 1)>), (unsigned)std::min(T, head_ctas * device_sm_count()), 128, 0,
                stream, I.hv, g_raw, P, T, g.rgb_w, g.rgb_b, g.alpha_b);   // the loop is load-latency bound
  return 0;
}
inline int field_tc_bwd(const scnerf_mlp& m, const scnerf_mlp& g, int precision, const float* rays, int ray_cols,
                        const float* z, int64_t N, int S, const FieldBufs& B, const TcFwdImages& I,
                        const TcBwdBufs& G, const float* g_raw, float* d_rays, void* stream,
                        const float* pts = nullptr, const float* viewdirs = nullptr, float* d_viewdirs = nullptr) {
  if (m.pts_dim == 4) {
    if (precision == SCNERF_PRECISION_BF16X3)
      return field_tc_bwd_impl<3, 96>(m, g, nullptr, 0, nullptr, N, S, B, I, G, g_raw, nullptr, stream, pts, viewdirs, d_viewdirs);
    return field_tc_bwd_impl<1, 96>(m, g, nullptr, 0, nullptr, N, S, B, I, G, g_raw, nullptr, stream, pts, viewdirs, d_viewdirs);
  }
  if (precision == SCNERF_PRECISION_BF16X3)
    return field_tc_bwd_impl<3>(m, g, rays, ray_cols, z, N, S, B, I, G, g_raw, d_rays, stream);
  return field_tc_bwd_impl<1>(m, g, rays, ray_cols, z, N, S, B, I, G, g_raw, d_rays, stream);
}

}  // namespace scnerf
