// ba_dense_fast.h -- cache-blocked, OpenMP-threaded dense kernels of the CPU oracle.
//
// TEST INFRASTRUCTURE ONLY (included by ba_oracle.cc; see ba_oracle.h).
//
// The reference spends most of a large joint-optimisation iteration in two Eigen calls:
//   * schur_M.triangularView<Upper>() -= off_diag^T * D^-1 off_diag   (LV/lm_optimizer.h:1328)
//   * schur_M.selfadjointView<Upper>().ldlt().solve(schur_b)          (LV/lm_optimizer.h:1361)
// Eigen's product kernel is cache-blocked and vectorised, so the plain triple loops of the small-
// problem oracle are not a fair stand-in for it at BASELINE config 2 (n_d = 13 080). These kernels
// are what `bench.py --impl reference` and `cpu_baseline` time at full size; they compute the same
// quantities as schur_solve() / ldlt_solve() in ba_oracle.cc (tests/test_oracle_golden.py checks
// the agreement), with the work split over the host threads:
//   * gemm_nt_lower: C[i][j] (j <= i) += alpha * sum_k A[i][k] * B[j][k], all operands row-major with
//     k contiguous -- both uses below are of this "dot product of two rows" shape, so no packing is
//     needed; 96 x 96 x 256 tiles live in L2, a 4x3 register block of 4-wide vectors streams them.
//   * ldlt_factor_blocked: right-looking blocked LDL^T. Eigen's pivoted LDLT picks, at step k, the
//     largest remaining ORIGINAL diagonal entry (ldlt_inplace<Lower>::unblocked is left-looking:
//     the trailing diagonal is untouched when the pivot is searched), so the whole pivot sequence
//     is known before the elimination starts. It is replayed on the diagonal alone, the symmetric
//     permutation is applied up front and the factorisation itself runs without pivot search.
#pragma once

#include <algorithm>
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <vector>

#ifdef _OPENMP
#include <omp.h>
#endif

namespace oracle_fast {

inline int& thread_count() {
  static int n = 1;
  return n;
}

// 4x3 register block over one k-panel: c[r][s] += sum_k a_r[k] * b_s[k]. Twelve 4-wide vector
// accumulators (GCC vector extension: AVX2 registers under -march=x86-64-v3 / native) + 7 loads per
// 12 multiply-adds; the horizontal sums happen once per panel.
typedef double v4d __attribute__((vector_size(32), aligned(8)));
static inline double hsum(v4d v) { return (v[0] + v[1]) + (v[2] + v[3]); }
static inline void micro_4x3(const double* a0, const double* a1, const double* a2, const double* a3,
                             const double* b0, const double* b1, const double* b2, int kc, double c[4][3]) {
  v4d c00 = {0, 0, 0, 0}, c01 = c00, c02 = c00, c10 = c00, c11 = c00, c12 = c00;
  v4d c20 = c00, c21 = c00, c22 = c00, c30 = c00, c31 = c00, c32 = c00;
  int k = 0;
  for (; k + 4 <= kc; k += 4) {
    const v4d y0 = *reinterpret_cast<const v4d*>(b0 + k), y1 = *reinterpret_cast<const v4d*>(b1 + k),
              y2 = *reinterpret_cast<const v4d*>(b2 + k);
    v4d x = *reinterpret_cast<const v4d*>(a0 + k);
    c00 += x * y0; c01 += x * y1; c02 += x * y2;
    x = *reinterpret_cast<const v4d*>(a1 + k);
    c10 += x * y0; c11 += x * y1; c12 += x * y2;
    x = *reinterpret_cast<const v4d*>(a2 + k);
    c20 += x * y0; c21 += x * y1; c22 += x * y2;
    x = *reinterpret_cast<const v4d*>(a3 + k);
    c30 += x * y0; c31 += x * y1; c32 += x * y2;
  }
  c[0][0] = hsum(c00); c[0][1] = hsum(c01); c[0][2] = hsum(c02);
  c[1][0] = hsum(c10); c[1][1] = hsum(c11); c[1][2] = hsum(c12);
  c[2][0] = hsum(c20); c[2][1] = hsum(c21); c[2][2] = hsum(c22);
  c[3][0] = hsum(c30); c[3][1] = hsum(c31); c[3][2] = hsum(c32);
  for (; k < kc; ++k) {
    const double x[4] = {a0[k], a1[k], a2[k], a3[k]}, y[3] = {b0[k], b1[k], b2[k]};
    for (int r = 0; r < 4; ++r)
      for (int t = 0; t < 3; ++t) c[r][t] += x[r] * y[t];
  }
}

// C[i][j] += alpha * sum_{k < K} A[i][k] * B[j][k] for 0 <= j <= i + diag_shift, i < M, j < N.
// (diag_shift = 0 with M == N: the lower triangle incl. the diagonal.) Row-major, leading dimensions
// lda / ldb / ldc. Threads split the (i-tile, j-tile) pairs.
inline void gemm_nt_lower(int M, int N, int K, double alpha, const double* A, size_t lda, const double* B, size_t ldb,
                          double* C, size_t ldc, int diag_shift) {
  constexpr int TI = 96, TJ = 96, KC = 256;  // 2 x 96 x 256 doubles = 384 KB of operands per tile pair
  const int nti = (M + TI - 1) / TI, ntj = (N + TJ - 1) / TJ;
  std::vector<std::pair<int, int>> tiles;
  for (int ti = 0; ti < nti; ++ti)
    for (int tj = 0; tj < ntj; ++tj) {
      const int i_hi = std::min(M, (ti + 1) * TI) - 1;
      if (tj * TJ <= i_hi + diag_shift) tiles.emplace_back(ti, tj);
    }
  const int nt = static_cast<int>(tiles.size());
#pragma omp parallel for schedule(dynamic, 1) num_threads(thread_count())
  for (int t = 0; t < nt; ++t) {
    const int i0 = tiles[t].first * TI, j0 = tiles[t].second * TJ;
    const int i1 = std::min(M, i0 + TI), j1 = std::min(N, j0 + TJ);
    for (int k0 = 0; k0 < K; k0 += KC) {
      const int kc = std::min(KC, K - k0);
      for (int i = i0; i < i1; i += 4) {
        const int ni = std::min(4, i1 - i);
        const double* ar[4];
        for (int r = 0; r < 4; ++r) ar[r] = A + static_cast<size_t>(std::min(i + r, i1 - 1)) * lda + k0;
        const int jmax = std::min(j1 - 1, i + ni - 1 + diag_shift);
        for (int j = j0; j <= jmax; j += 3) {
          const int nj = std::min(3, j1 - j);
          const double* br[3];
          for (int s = 0; s < 3; ++s) br[s] = B + static_cast<size_t>(std::min(j + s, j1 - 1)) * ldb + k0;
          double c[4][3];
          micro_4x3(ar[0], ar[1], ar[2], ar[3], br[0], br[1], br[2], kc, c);
          for (int r = 0; r < ni; ++r)
            for (int s = 0; s < nj; ++s)
              if (j + s <= i + r + diag_shift) C[static_cast<size_t>(i + r) * ldc + (j + s)] += alpha * c[r][s];
        }
      }
    }
  }
}

// Replays the pivot search of the pivoted LDL^T on the diagonal alone (see the header comment).
inline void ldlt_pivot_sequence(int n, std::vector<double> diag, std::vector<int>* transp) {
  transp->resize(n);
  for (int k = 0; k < n; ++k) {
    int piv = k;
    double best = std::fabs(diag[k]);
    for (int i = k + 1; i < n; ++i) {
      const double v = std::fabs(diag[i]);
      if (v > best) {
        best = v;
        piv = i;
      }
    }
    (*transp)[k] = piv;
    if (piv != k) std::swap(diag[k], diag[piv]);
  }
}

// In-place blocked LDL^T of the symmetric matrix whose LOWER triangle is stored row-major in Lm
// (n x n, leading dimension n), no pivoting (apply the permutation first). On exit: unit-lower L
// below the diagonal, D on the diagonal.
inline void ldlt_factor_blocked(int n, double* Lm) {
  constexpr int NB = 192;
  std::vector<double> W;  // (rows below the panel) x nb: L21 * D
  auto at = [&](int i, int j) -> double& { return Lm[static_cast<size_t>(i) * n + j]; };
  for (int k0 = 0; k0 < n; k0 += NB) {
    const int nb = std::min(NB, n - k0);
    const int k1 = k0 + nb;
    // 1) diagonal block: unblocked right-looking LDL^T
    for (int k = k0; k < k1; ++k) {
      const double dk = at(k, k);
      if (std::fabs(dk) > 0) {
        for (int i = k + 1; i < k1; ++i) {
          const double c = at(i, k);
          const double lik = c / dk;
          at(i, k) = lik;
          // row i, columns k+1..i get -= l_ik * (l_jk d_k); column k of rows j <= i is already final
          for (int j = k + 1; j <= i; ++j) at(i, j) -= lik * (j == i ? c : at(j, k) * dk);
        }
      }
    }
    if (k1 >= n) break;
    const int m = n - k1;
    // 2) panel: rows below solve L21 D L11^T = A21  (row by row, independent rows)
#pragma omp parallel for schedule(static) num_threads(thread_count())
    for (int i = k1; i < n; ++i) {
      double* row = &at(i, 0);
      for (int k = k0; k < k1; ++k) {
        double s = row[k];
        for (int j = k0; j < k; ++j) s -= row[j] * at(k, j) * at(j, j);  // row[j] already holds l_ij
        const double dk = at(k, k);
        row[k] = (std::fabs(dk) > 0) ? s / dk : 0.0;
      }
    }
    // 3) trailing update: A22 -= (L21 D) L21^T on the lower triangle
    W.resize(static_cast<size_t>(m) * nb);
#pragma omp parallel for schedule(static) num_threads(thread_count())
    for (int i = 0; i < m; ++i)
      for (int k = 0; k < nb; ++k) W[static_cast<size_t>(i) * nb + k] = at(k1 + i, k0 + k) * at(k0 + k, k0 + k);
    gemm_nt_lower(m, m, nb, -1.0, W.data(), nb, &at(k1, k0), n, &at(k1, k1), n, 0);
  }
}

}  // namespace oracle_fast
