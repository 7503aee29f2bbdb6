// ba_dense.cu -- hand-written FP64 tensor-core (DMMA) kernels of the dense phase (sm_100a).
//
// The reduced system of the Schur-complement solve (LV/lm_optimizer.h:1246-1369) is formed and
// factorised with the kernels of this file instead of library calls:
//
//   dgemm_nt_kernel     C (+)= alpha * A B^T on 128 x 128 tiles, FP64 `mma.sync.m8n8k4` (SASS DMMA.8x8x4,
//                       the only FP64 tensor shape sm_100a has), operands staged through shared memory
//                       by a 4-stage cp.async pipeline. One kernel, three uses:
//                         * LOWER + plain epilogue: trailing update S22 -= L21 L21^T of the blocked Cholesky
//                           (the reference factors with Eigen's LDLT, LV/lm_optimizer.h:1361);
//                         * plain epilogue, K = N = 128: the panel solve L21 = A21 L11^-T as a product with the
//                           explicitly inverted diagonal tile;
//                         * LOWER + scatter epilogue: the structured Schur contraction S -= W_g^T W_g on the
//                           compact panel of one group of Schur blocks, scattered straight into S through
//                           the group's column list (the reference contracts with Eigen / cublasXtDgemm,
//                           LV/lm_optimizer.h:1328,1371-1430) -- no m_g x m_g temporary, no scatter pass.
//   potrf_tile_kernel   Cholesky of one 128 x 128 diagonal tile in shared memory + its explicit inverse.
//   small helpers       lambda on the diagonal, column-block copies.
//
// All operands of dgemm_nt are "k-strided": element (i, k) of A lives at A[k * lda + i] (i contiguous),
// which is what both a column-major panel of S and a row-major compact panel W_g[k][j] look like.
// C is column-major (element (i, j) at C[j * ldc + i]); for a symmetric result only i >= j is touched.

#include <cstdio>

#include "ba_kernels.h"

namespace b200ba {

namespace {

constexpr int BM = 128, BN = 128, BK = 16, STAGES = 4;
constexpr int LDT = BM + 4;  // shared-memory row pitch: (k * LDT + m) mod 16 is distinct for k, m in 0..3 -> no bank conflicts
constexpr int GEMM_THREADS = 256;
constexpr size_t kGemmSmem = static_cast<size_t>(STAGES) * 2 * BK * LDT * sizeof(double);

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem, int src_bytes) {
  const unsigned s = static_cast<unsigned>(__cvta_generic_to_shared(smem));
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(s), "l"(gmem), "r"(src_bytes));
}
__device__ __forceinline__ void cp_async8(void* smem, const void* gmem, int src_bytes) {
  const unsigned s = static_cast<unsigned>(__cvta_generic_to_shared(smem));
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;\n" ::"r"(s), "l"(gmem), "r"(src_bytes));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;\n" ::"n"(N));
}
__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}

// One BK x 128 operand tile: rows k0 .. k0 + BK of the k-strided matrix X (leading dimension ldx), columns
// i0 .. i0 + 128, zero-filled beyond (rows, K). `aligned` = every 16-byte chunk is 16-byte aligned in
// global memory (even ldx, even i0, 16-byte aligned base); otherwise 8-byte copies.
__device__ __forceinline__ void load_tile(double* dst, const double* __restrict__ X, int64_t ldx, int rows, int K, int i0,
                                          int k0, bool aligned) {
  if (aligned) {
    // 16 rows x 64 chunks of 2 doubles = 1024 chunks, 4 per thread
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int c = threadIdx.x + q * GEMM_THREADS;
      const int kk = c >> 6, ch = c & 63;
      const int i = i0 + 2 * ch, k = k0 + kk;
      int bytes = 0;
      if (k < K) bytes = (i + 1 < rows) ? 16 : ((i < rows) ? 8 : 0);
      cp_async16(dst + kk * LDT + 2 * ch, bytes ? (X + static_cast<int64_t>(k) * ldx + i) : X, bytes);
    }
  } else {
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const int c = threadIdx.x + q * GEMM_THREADS;
      const int kk = c >> 7, ii = c & 127;
      const int i = i0 + ii, k = k0 + kk;
      const bool ok = (k < K) && (i < rows);
      cp_async8(dst + kk * LDT + ii, ok ? (X + static_cast<int64_t>(k) * ldx + i) : X, ok ? 8 : 0);
    }
  }
}

}  // namespace

// C(i, j) at Cbase + col_off(j) + i, where col_off maps a column to its storage offset (see DenseMap).
template <bool LOWER, int EPI>
__global__ void __launch_bounds__(GEMM_THREADS, 1) dgemm_nt_kernel(GemmArgs g) {
  const int tm = blockIdx.x, tn = blockIdx.y;
  if (LOWER && tn > tm) return;
  extern __shared__ __align__(16) double smem_d[];
  double* As = smem_d;
  double* Bs = smem_d + static_cast<size_t>(STAGES) * BK * LDT;
  const int m0 = tm * BM, n0 = tn * BN;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int wm = (warp & 1) * 64, wn = (warp >> 1) * 32;  // warp tile 64 (m) x 32 (n)
  const int lr = lane >> 2, lc = lane & 3;
  const bool same = LOWER && (tm == tn) && (g.A == g.B) && (g.lda == g.ldb);  // diagonal tile of a syrk: one operand

  double acc[8][4][2];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

  const int nk = (g.K + BK - 1) / BK;
  // prologue
#pragma unroll
  for (int s = 0; s < STAGES - 1; ++s) {
    if (s < nk) {
      load_tile(As + s * BK * LDT, g.A, g.lda, g.M, g.K, m0, s * BK, g.a_aligned);
      if (!same) load_tile(Bs + s * BK * LDT, g.B, g.ldb, g.N, g.K, n0, s * BK, g.b_aligned);
    }
    cp_async_commit();
  }
  for (int kt = 0; kt < nk; ++kt) {
    cp_async_wait<STAGES - 2>();
    __syncthreads();
    // prefetch the tile STAGES - 1 ahead into the slot that was consumed in the previous iteration
    {
      const int nt = kt + STAGES - 1;
      if (nt < nk) {
        const int s = nt % STAGES;
        load_tile(As + s * BK * LDT, g.A, g.lda, g.M, g.K, m0, nt * BK, g.a_aligned);
        if (!same) load_tile(Bs + s * BK * LDT, g.B, g.ldb, g.N, g.K, n0, nt * BK, g.b_aligned);
      }
      cp_async_commit();
    }
    const double* a_s = As + (kt % STAGES) * BK * LDT;
    const double* b_s = same ? a_s : (Bs + (kt % STAGES) * BK * LDT);
#pragma unroll
    for (int ks = 0; ks < BK / 4; ++ks) {
      double af[8], bf[4];
      const double* ap = a_s + (ks * 4 + lc) * LDT + wm + lr;
      const double* bp = b_s + (ks * 4 + lc) * LDT + wn + lr;
#pragma unroll
      for (int i = 0; i < 8; ++i) af[i] = ap[8 * i];
#pragma unroll
      for (int j = 0; j < 4; ++j) bf[j] = bp[8 * j];
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) dmma884(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
    }
  }
  cp_async_wait<0>();

  // epilogue: accumulator (i, j) holds rows m0 + wm + 8 i + lr, columns n0 + wn + 8 j + 2 lc + {0, 1}
#pragma unroll
  for (int j = 0; j < 4; ++j) {
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const int n = n0 + wn + 8 * j + 2 * lc + e;
      if (n >= g.N) continue;
      if (EPI == 0) {
        double* ccol = g.C + static_cast<int64_t>(n) * g.ldc;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int m = m0 + wm + 8 * i + lr;
          if (m >= g.M || (LOWER && m < n)) continue;
          const double v = g.alpha * acc[i][j][e];
          ccol[m] = (g.beta == 0.0) ? v : fma(g.beta, ccol[m], v);
        }
      } else {
        // scatter: compact column n is dense column cols[n] of S; rows likewise (cols ascending, m >= n)
        const int cn = g.cols[n];
        double* scol = g.C + g.map.col_offset(cn);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int m = m0 + wm + 8 * i + lr;
          if (m >= g.M || m < n) continue;
          scol[g.cols[m]] += g.alpha * acc[i][j][e];
        }
      }
    }
  }
}

int launch_dgemm_nt(const GemmArgs& g, bool lower, bool scatter, cudaStream_t s) {
  if (g.M <= 0 || g.N <= 0) return 0;
  static bool configured = false;
  if (!configured) {
    cudaFuncSetAttribute(dgemm_nt_kernel<true, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(kGemmSmem));
    cudaFuncSetAttribute(dgemm_nt_kernel<false, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(kGemmSmem));
    cudaFuncSetAttribute(dgemm_nt_kernel<true, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(kGemmSmem));
    configured = true;
  }
  dim3 grid((g.M + BM - 1) / BM, (g.N + BN - 1) / BN);
  if (scatter)
    dgemm_nt_kernel<true, 1><<<grid, GEMM_THREADS, kGemmSmem, s>>>(g);
  else if (lower)
    dgemm_nt_kernel<true, 0><<<grid, GEMM_THREADS, kGemmSmem, s>>>(g);
  else
    dgemm_nt_kernel<false, 0><<<grid, GEMM_THREADS, kGemmSmem, s>>>(g);
  return cudaGetLastError() == cudaSuccess ? 0 : 1;
}

// ------------------------------------------------------------------------------------------
// 128 x 128 diagonal tile: Cholesky in shared memory + explicit inverse of the factor
// ------------------------------------------------------------------------------------------
// A: column-major tile (leading dimension lda), lower triangle read; on exit its lower triangle holds L.
// Linv: 128 x 128 column-major (leading dimension 128), lower triangle = L^-1, strict upper = 0.
// n <= 128 is the live size (last tile of the matrix); the rest is treated as identity.
// info[0] is raised when a pivot is not positive (the LM loop then rejects the attempt).
constexpr int PT = 128;        // tile size
constexpr int PLD = PT + 1;    // odd pitch: column walks and row walks are both conflict-free
constexpr int POTRF_THREADS = 512;

__global__ void __launch_bounds__(POTRF_THREADS, 1)
    potrf_tile_kernel(double* __restrict__ A, int64_t lda, int n, double* __restrict__ Linv, int* __restrict__ info) {
  extern __shared__ __align__(16) double sm[];
  double* L = sm;                 // [PT][PLD] column-major: L(i, j) at L[j * PLD + i]
  double* X = sm + PT * PLD;      // inverse, same layout
  const int tid = threadIdx.x;
  for (int e = tid; e < PT * PT; e += POTRF_THREADS) {
    const int j = e >> 7, i = e & 127;
    double v = (i == j) ? 1.0 : 0.0;
    if (i < n && j < n && i >= j) v = A[static_cast<int64_t>(j) * lda + i];
    L[j * PLD + i] = (i >= j) ? v : 0.0;
  }
  __syncthreads();
  __shared__ int s_bad;
  if (tid == 0) s_bad = 0;
  // right-looking, one column at a time. Fixed ownership for the trailing update: thread t owns row
  // (t & 127) and the columns c with c % 4 == t >> 7, so consecutive threads touch consecutive
  // shared-memory words of one column and L(row, j) is read once per step.
  const int prow = tid & 127, pgrp = tid >> 7;
  for (int j = 0; j < PT; ++j) {
    __syncthreads();
    const double d = L[j * PLD + j];
    if (tid == 0 && !(d > 0.0)) s_bad = 1;
    const double sq = (d > 0.0) ? sqrt(d) : 1.0;
    __syncthreads();
    if (tid < PT - j) {
      const int i = j + tid;
      L[j * PLD + i] = (tid == 0) ? sq : L[j * PLD + i] / sq;
    }
    __syncthreads();
    if (prow > j) {
      const double lrj = L[j * PLD + prow];
      int c = j + 1 + ((pgrp - (j + 1)) & 3);  // first column > j with c % 4 == pgrp
      for (; c <= prow; c += 4) L[c * PLD + prow] -= lrj * L[j * PLD + c];
    }
  }
  __syncthreads();
  if (tid == 0 && s_bad) info[0] = 1;
  // write L back
  for (int e = tid; e < PT * PT; e += POTRF_THREADS) {
    const int j = e >> 7, i = e & 127;
    if (i < n && j < n && i >= j) A[static_cast<int64_t>(j) * lda + i] = L[j * PLD + i];
  }
  // inverse by forward substitution, one column per thread group: column c of X solves L x = e_c.
  // 4 threads share a column (strided over the inner product), 128 columns -> 512 threads.
  {
    const int c = tid >> 2, q = tid & 3;
    for (int i = 0; i < PT; ++i) {
      // x_i = (delta_ic - sum_{k=c}^{i-1} L(i,k) x_k) / L(i,i); rows above c are zero
      double s = 0.0;
      if (i > c)
        for (int k = c + q; k < i; k += 4) s += L[k * PLD + i] * X[c * PLD + k];
      s += __shfl_xor_sync(0xffffffffu, s, 1);
      s += __shfl_xor_sync(0xffffffffu, s, 2);
      if (q == 0) X[c * PLD + i] = (i < c) ? 0.0 : (((i == c) ? 1.0 : 0.0) - s) / L[i * PLD + i];
      __syncwarp();
    }
  }
  __syncthreads();
  for (int e = tid; e < PT * PT; e += POTRF_THREADS) {
    const int j = e >> 7, i = e & 127;
    Linv[j * PT + i] = X[j * PLD + i];
  }
}

int launch_potrf_tile(double* A, int64_t lda, int n, double* Linv, int* info, cudaStream_t s) {
  static bool configured = false;
  const size_t smem = 2 * static_cast<size_t>(PT) * PLD * sizeof(double);
  if (!configured) {
    cudaFuncSetAttribute(potrf_tile_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    configured = true;
  }
  potrf_tile_kernel<<<1, POTRF_THREADS, smem, s>>>(A, lda, n, Linv, info);
  return cudaGetLastError() == cudaSuccess ? 0 : 1;
}

}  // namespace b200ba
