// fp32 CUDA-core GEMM used by the SCNERF_PRECISION_FP32 field path (exact fp32 numerics, the
// parity anchor for the tensor-core path) and by every backward GEMM until the fused tcgen05
// backward lands.  128x128x16 tiles, 256 threads, 8x8 register micro-tile, register-staged
// double buffering.  Bound: fp32 FMA pipe (148 SMs x 128 FMA/clk), not HBM.
//
//   C[M,N] (op)= epilogue( sum_k A(m,k) * B(k,n) )
//   A_KC: A stored [m][k] (k contiguous, lda)   else stored [k][m] (m contiguous, lda)
//   B_KC: B stored [n][k] (k contiguous, ldb)   else stored [k][n] (n contiguous, ldb)
//     forward  y = x W^T      : A_KC=1 (x[P,K]),   B_KC=1 (W[N,K])
//     dgrad    dx = dz W      : A_KC=1 (dz[P,N']), B_KC=0 (W[N',K'] read as [k][n])
//     wgrad    dW = dz^T x    : A_KC=0 (dz[P,N'] read as [k][m]), B_KC=0 (x[P,K'] as [k][n]), split-K
#pragma once
#include "common.cuh"

namespace scnerf {

struct GemmArgs {
  const float* A; int64_t lda;
  const float* B; int64_t ldb;
  float* C; int64_t ldc;
  int M, N, K;
  const float* bias;                 // [N] or NULL
  int relu;                          // C = max(C, 0)
  const float* mask; int64_t ldmask; // C = mask[m, n] > 0 ? C : 0 for n >= mask_col0 (relu backward)
  int mask_col0;
  int accumulate;                    // C += result (non-atomic)
  int atomic;                        // atomicAdd (split-K)
  int k_chunk;                       // K range per blockIdx.z
};

constexpr int GBM = 128, GBN = 128, GBK = 16, GPAD = 4;

template <bool KC>
__device__ __forceinline__ void gemm_load_tile(const float* __restrict__ X, int64_t ld, int row0,
                                               int nrows, int k0, int kend, bool vec_ok, int tid,
                                               float (&regs)[8]) {
  if (KC) {
    // thread -> (row = tid/4 [+64], k = (tid%4)*4 .. +3)
    int kk = k0 + (tid & 3) * 4;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      int row = row0 + (tid >> 2) + h * 64;
      const float* p = X + (int64_t)row * ld + kk;
      if (row < nrows && vec_ok && kk + 3 < kend) {
        float4 v = *reinterpret_cast<const float4*>(p);
        regs[h * 4 + 0] = v.x; regs[h * 4 + 1] = v.y; regs[h * 4 + 2] = v.z; regs[h * 4 + 3] = v.w;
      } else {
#pragma unroll
        for (int c = 0; c < 4; ++c)
          regs[h * 4 + c] = (row < nrows && kk + c < kend) ? p[c] : 0.f;
      }
    }
  } else {
    // thread -> (k = tid/32 [+8], row = (tid%32)*4 .. +3)
    int rr = row0 + (tid & 31) * 4;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      int kk = k0 + (tid >> 5) + h * 8;
      const float* p = X + (int64_t)kk * ld + rr;
      if (kk < kend && vec_ok && rr + 3 < nrows) {
        float4 v = *reinterpret_cast<const float4*>(p);
        regs[h * 4 + 0] = v.x; regs[h * 4 + 1] = v.y; regs[h * 4 + 2] = v.z; regs[h * 4 + 3] = v.w;
      } else {
#pragma unroll
        for (int c = 0; c < 4; ++c)
          regs[h * 4 + c] = (kk < kend && rr + c < nrows) ? p[c] : 0.f;
      }
    }
  }
}

template <bool KC>
__device__ __forceinline__ void gemm_store_tile(float (*S)[GBM + GPAD], int tid, const float (&regs)[8]) {
  if (KC) {
    int kk = (tid & 3) * 4;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      int row = (tid >> 2) + h * 64;
#pragma unroll
      for (int c = 0; c < 4; ++c) S[kk + c][row] = regs[h * 4 + c];
    }
  } else {
    int rr = (tid & 31) * 4;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      int kk = (tid >> 5) + h * 8;
      *reinterpret_cast<float4*>(&S[kk][rr]) =
          make_float4(regs[h * 4], regs[h * 4 + 1], regs[h * 4 + 2], regs[h * 4 + 3]);
    }
  }
}

template <bool A_KC, bool B_KC>
__global__ void __launch_bounds__(256) gemm_simt_kernel(GemmArgs g) {
  __shared__ __align__(16) float As[GBK][GBM + GPAD];
  __shared__ __align__(16) float Bs[GBK][GBN + GPAD];
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const int m0 = blockIdx.y * GBM, n0 = blockIdx.x * GBN;
  const int kbeg = blockIdx.z * g.k_chunk;
  const int kend = min(g.K, kbeg + g.k_chunk);
  const bool a_vec = ((reinterpret_cast<uintptr_t>(g.A) & 15) == 0) && (g.lda % 4 == 0);
  const bool b_vec = ((reinterpret_cast<uintptr_t>(g.B) & 15) == 0) && (g.ldb % 4 == 0);

  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

  float ra[8], rb[8];
  gemm_load_tile<A_KC>(g.A, g.lda, m0, g.M, kbeg, kend, a_vec, tid, ra);
  gemm_load_tile<B_KC>(g.B, g.ldb, n0, g.N, kbeg, kend, b_vec, tid, rb);
  for (int k0 = kbeg; k0 < kend; k0 += GBK) {
    gemm_store_tile<A_KC>(As, tid, ra);
    gemm_store_tile<B_KC>(Bs, tid, rb);
    __syncthreads();
    if (k0 + GBK < kend) {
      gemm_load_tile<A_KC>(g.A, g.lda, m0, g.M, k0 + GBK, kend, a_vec, tid, ra);
      gemm_load_tile<B_KC>(g.B, g.ldb, n0, g.N, k0 + GBK, kend, b_vec, tid, rb);
    }
#pragma unroll
    for (int kk = 0; kk < GBK; ++kk) {
      float4 a0 = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
      float4 a1 = *reinterpret_cast<const float4*>(&As[kk][64 + ty * 4]);
      float4 b0 = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
      float4 b1 = *reinterpret_cast<const float4*>(&Bs[kk][64 + tx * 4]);
      float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      float bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }

  // epilogue
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    int m = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
    if (m >= g.M) continue;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      int n = n0 + (j < 4 ? tx * 4 + j : 64 + tx * 4 + (j - 4));
      if (n >= g.N) continue;
      float v = acc[i][j];
      if (g.bias) v += g.bias[n];
      if (g.relu) v = fmaxf(v, 0.f);
      float* c = g.C + (int64_t)m * g.ldc + n;
      if (g.atomic) { atomicAdd(c, v); continue; }
      if (g.accumulate) v += *c;
      if (g.mask && n >= g.mask_col0 && !(g.mask[(int64_t)m * g.ldmask + n] > 0.f)) v = 0.f;
      *c = v;
    }
  }
}

inline int gemm_launch(const GemmArgs& g, bool a_kc, bool b_kc, int splits, void* stream) {
  dim3 grid((unsigned)cdiv(g.N, GBN), (unsigned)cdiv(g.M, GBM), (unsigned)splits);
  if (a_kc && b_kc) SCNERF_LAUNCH((gemm_simt_kernel<true, true>), grid, 256, 0, stream, g);
  else if (a_kc && !b_kc) SCNERF_LAUNCH((gemm_simt_kernel<true, false>), grid, 256, 0, stream, g);
  else if (!a_kc && !b_kc) SCNERF_LAUNCH((gemm_simt_kernel<false, false>), grid, 256, 0, stream, g);
  else return fail(SCNERF_ERR_UNSUPPORTED, "gemm layout (A_KC=0,B_KC=1) not instantiated");
  return 0;
}

// y = act(x W^T + b):  x[P,K] (ldx), W[N,K], y[P,N] (ldy)
inline int linear_fwd(const float* x, int64_t ldx, const float* W, const float* b, float* y,
                      int64_t ldy, int64_t P, int N, int K, bool relu, void* stream) {
  GemmArgs g{};
  g.A = x; g.lda = ldx; g.B = W; g.ldb = K; g.C = y; g.ldc = ldy;
  g.M = (int)P; g.N = N; g.K = K; g.bias = b; g.relu = relu; g.k_chunk = K;
  return gemm_launch(g, true, true, 1, stream);
}
// dx = dz W  (optionally += and/or masked by mask>0 from column mask_col0 on)
inline int linear_dgrad(const float* dz, int64_t lddz, const float* W, int64_t ldw, float* dx,
                        int64_t lddx, int64_t P, int Nout, int Kin, const float* mask,
                        int64_t ldmask, int mask_col0, bool accumulate, void* stream) {
  GemmArgs g{};
  g.A = dz; g.lda = lddz; g.B = W; g.ldb = ldw; g.C = dx; g.ldc = lddx;
  g.M = (int)P; g.N = Kin; g.K = Nout; g.mask = mask; g.ldmask = ldmask; g.mask_col0 = mask_col0;
  g.accumulate = accumulate; g.k_chunk = Nout;
  return gemm_launch(g, true, false, 1, stream);
}
// dW[Nout,Kin] += dz^T x   (split over P with atomics)
inline int linear_wgrad(const float* dz, int64_t lddz, const float* x, int64_t ldx, float* dW,
                        int64_t lddw, int64_t P, int Nout, int Kin, void* stream) {
  GemmArgs g{};
  g.A = dz; g.lda = lddz; g.B = x; g.ldb = ldx; g.C = dW; g.ldc = lddw;
  g.M = Nout; g.N = Kin; g.K = (int)P; g.atomic = 1;
  int tiles = (int)(cdiv(Nout, GBM) * cdiv(Kin, GBN));
  int splits = (int)std::max<int64_t>(1, std::min<int64_t>(cdiv(P, 512), (148 * 4) / tiles));
  int chunk = (int)cdiv(cdiv(P, splits), GBK) * GBK;
  g.k_chunk = chunk;
  splits = (int)cdiv(P, chunk);
  return gemm_launch(g, false, false, splits, stream);
}

// db[n] += sum_p dz[p, n]
__global__ void __launch_bounds__(256) colsum_kernel(const float* __restrict__ dz, int64_t ld,
                                                     int64_t P, int N, int64_t rows_per_block,
                                                     float* __restrict__ db) {
  int n = blockIdx.x * 32 + (threadIdx.x & 31);
  int64_t p0 = (int64_t)blockIdx.y * rows_per_block;
  int64_t p1 = min(P, p0 + rows_per_block);
  float s = 0.f;
  if (n < N)
    for (int64_t p = p0 + (threadIdx.x >> 5); p < p1; p += 8) s += dz[p * ld + n];
  __shared__ float red[8][33];
  red[threadIdx.x >> 5][threadIdx.x & 31] = s;
  __syncthreads();
  if (threadIdx.x < 32 && n < N) {
    float t = 0.f;
    for (int k = 0; k < 8; ++k) t += red[k][threadIdx.x];
    atomicAdd(db + n, t);
  }
}
inline int bias_grad(const float* dz, int64_t ld, int64_t P, int N, float* db, void* stream) {
  int64_t nb = std::min<int64_t>(cdiv(P, 256), 592);
  int64_t rpb = cdiv(P, nb);
  dim3 grid((unsigned)cdiv(N, 32), (unsigned)cdiv(P, rpb));
  SCNERF_LAUNCH(colsum_kernel, grid, 256, 0, stream, dz, ld, P, N, rpb, db);
  return 0;
}

}  // namespace scnerf
