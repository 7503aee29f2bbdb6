// ba_tile.cuh -- the 128 x 128 diagonal-tile step of the blocked Cholesky (ba_dense.cu): factor the tile and
// invert the factor, written as a sequence of barrier-separated PHASES over the 512 threads of a CTA.
//
// Where this sits: every 128 columns, the factorisation's critical path runs  tile Cholesky -> inverse of the
// factor -> panel solve (a product with that inverse). The first version (potrf_tile_kernel / trinv_tile_kernel in
// ba_dense.cu, still selectable) keeps a 32-column register window per thread and issues 32 shared-memory loads
// per thread and column step: 100 k cycles per tile, LSU-bound, and a second launch of 45 k cycles for the inverse.
// Here both are blocked by 16 columns:
//   * Cholesky: per 16-column panel, 16 software-pipelined column steps on a 128 x 16 register panel (4 entries per
//     thread, one barrier per step through a double-buffered published column; only the pivot load, a reciprocal
//     seed with one cubic correction and the update of the ONE entry published next sit between two barriers, see
//     panel_step), then ONE rank-16 update of the trailing lower triangle with a register block of 2 rows x 8
//     columns per thread (at most 420 blocks: a single pass). Measured: 47 us per tile for factor + inverse
//     (first version 50 + 22 us in two launches); the chain is bound by the dependent FP64 latency.
//   * inverse: the launch has 8 CTAs; each repeats the (deterministic) factorisation and then solves L X = E_J
//     for its own 16 columns J by block forward substitution with the explicitly inverted 16 x 16 diagonal
//     blocks. CTA 0 stores the factor. One launch instead of two, no reload of the tile.
//
// The program is a template over an executor: on the device `run(f)` calls f for this thread and then
// __syncthreads(); the host executor of tests/tile_emulation.cc calls f for all 512 threads of the CTA in a
// forward, reversed or shuffled order -- every phase must give the same result in all of them (no thread may read
// what another thread writes in the same phase), which pins the barrier placement and all index arithmetic
// without a GPU.
#pragma once

#include <cmath>
#include <cstdint>

#if defined(__CUDACC__)
#define B200_TILE_HD __host__ __device__ __forceinline__
#else
#define B200_TILE_HD inline
#endif

namespace b200ba {
namespace tile {

constexpr int PT = 128;            // tile size
constexpr int PLD = PT + 2;        // even pitch: the 4-row groups of the register blocks are 16-byte aligned
constexpr int THREADS = 512;
constexpr int PW = 16;             // panel width / diagonal block of the inverse
constexpr int NPANEL = PT / PW;    // 8
constexpr int CTAS = NPANEL;       // CTA J computes columns 16 J .. 16 J + 15 of the inverse
constexpr int XLD = PW + 1;        // pitch of the inverse's work arrays (conflict-free row and column walks)
constexpr int MAX_BLOCKS = 2 * (PT - PW) / 8 * ((PT - PW) / 8 + 1);  // 420 register blocks (2 rows x 8 columns) in the first trailing update

struct alignas(16) D2 {
  double x, y;
};

struct alignas(16) Shared {
  double T[PT * PLD];        // the tile, column-major: (i, j) at T[j * PLD + i]
  double cb[2][PT];          // published (un-scaled) column of the current step, double-buffered
  double lrj[THREADS];       // per thread: a(r, j) / a(j, j) of the last published column (see panel_step)
  double dpiv[PT];           // the pivots a(j, j) at elimination time
  double invdiag[PT];        // 1 / L(j, j)
  double R[PT * XLD];        // inverse: right-hand side rows x 16 columns
  double X[PT * XLD];        // inverse: solution rows x 16 columns
  double Dv[NPANEL][PW][PW]; // Dv[I][c][i] = (L_II^-1)(i, c) for the diagonal 16 x 16 blocks
  unsigned char bc[MAX_BLOCKS + 4], br[MAX_BLOCKS + 4];  // enumeration of the register blocks of the trailing update
};

struct Thread {
  double pa[4];      // this thread's entries of the current 128 x 16 panel: row tid & 127, columns (tid >> 7) + 4 q
  double m_prev[4];  // the last published column at this thread's four columns: the deferred part of its update
  bool bad;          // a pivot was not positive
};

B200_TILE_HD double inv_sqrt(double x) {
#if defined(__CUDA_ARCH__)
  return rsqrt(x);
#else
  return 1.0 / std::sqrt(x);
#endif
}
// ~23-bit reciprocal (MUFU.RCP64H); refined inside panel_step
B200_TILE_HD double reciprocal_seed(double x) {
#if defined(__CUDA_ARCH__)
  double y;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  return y;
#else
  return static_cast<double>(static_cast<float>(1.0 / x));
#endif
}

// One column step of the panel j0 .. j0 + 15: phase S = 0 .. 16, one barrier each. Column j0 + s is PUBLISHED
// (un-scaled, through the double-buffered cb) in phase s by its owner group s & 3 and applied
//   * in phase s + 1 to the entries of column j0 + s + 1 only ("early": that column is published in the same phase),
//   * in phase s + 2 to all other entries to its right ("deferred": multiplier and column values wait in registers).
// So between two barriers the critical path is  load pivot -> reciprocal -> one entry update -> publish , and the
// bulk of the update overlaps the next step's loads. FP64 operations are long-latency here: the reciprocal is the
// hardware seed x0 plus ONE cubic correction 1 / d = x0 (1 + e2), e = 1 - d x0, e2 = e + e^2 (relative error
// e^3 ~ 2^-69), and the early entry is formed as (a - p x0) - (p x0) e2 with p = a(r, jp) a(jn, jp): three
// dependent FMAs after the seed. Square roots are not on the path at all: phase 16 scales the finished panel
// (L(r, c) = a(r, c) / sqrt(a(c, c))), four independent rsqrt per thread, and checks the pivots.
template <int S>
B200_TILE_HD void panel_step(Thread& t, Shared& sh, int j0, int tid) {
  const int r = tid & (PT - 1), cg = tid >> 7;
  const int jp = j0 + S - 1, jn = j0 + S;
  constexpr int QN = (S & (PW - 1)) >> 2;  // register of column jn in its owner group
  const bool publisher = (S < PW) && cg == (S & 3);
  if (S == 0) {
#pragma unroll
    for (int q = 0; q < 4; ++q) t.pa[q] = sh.T[(j0 + cg + 4 * q) * PLD + r];
  }
  // Everything a phase needs comes from shared memory at its start -- including the row multiplier of the previous
  // column, which its phase stored (sh.lrj) instead of keeping in a register: a register-only result may be sunk
  // below the barrier by the scheduler and then sits, as a dependent chain, in front of these loads (in-order issue).
  const double* cbp = sh.cb[(S + 1) & 1];  // = cb[(S - 1) & 1]: column jp
  double d = 1.0, ar = 0.0, ajn = 0.0, lrj_prev = 0.0, m[4] = {0.0, 0.0, 0.0, 0.0};
  if (S >= 1 && S < PW) {
    d = cbp[jp];
    ajn = cbp[jn];
    ar = cbp[r];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int c = j0 + cg + 4 * q;
      if (c > jn) m[q] = cbp[c];  // columns up to jp are finished, jn is taken early (below): zero = nothing to do
    }
  }
  if (S >= 2 && S <= PW) {
    // deferred part of column jp - 1 (m_prev is zero where nothing is to be done: no selects here)
    lrj_prev = sh.lrj[tid];
#pragma unroll
    for (int q = 0; q < 4; ++q) t.pa[q] = fma(-lrj_prev, t.m_prev[q], t.pa[q]);
  }
  if (S >= 1 && S < PW) {
    const double x0 = reciprocal_seed(d);
    const double e = fma(-d, x0, 1.0);
    const double e2 = fma(e, e, e);
    // branch-free (every thread forms the candidate, the publisher keeps it): inside a divergent block the
    // scheduler would issue the independent products only after e and e2, which puts them back on the chain
    const double t1 = (ar * ajn) * x0;
    const double early = fma(-t1, e2, t.pa[QN] - t1);
    t.pa[QN] = publisher ? early : t.pa[QN];
    const double ax0 = ar * x0;
    sh.lrj[tid] = fma(ax0, e2, ax0);  // a(r, jp) / a(jp, jp), for the deferred part in the next phase
#pragma unroll
    for (int q = 0; q < 4; ++q) t.m_prev[q] = m[q];
  }
  if (publisher) {
    sh.cb[S & 1][r] = (r >= jn) ? t.pa[QN] : 0.0;
    if (r == jn) sh.dpiv[jn] = t.pa[QN];
  }
  if (S == PW) {
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int c = j0 + cg + 4 * q;
      const double dv = sh.dpiv[c];
      t.bad |= !(dv > 1e-290);
      const double rs = inv_sqrt(dv);
      if (r >= c) sh.T[c * PLD + r] = t.pa[q] * rs;
      if (r == c) sh.invdiag[c] = rs;
    }
  }
}

template <int S, class Exec>
struct PanelSteps {
  static B200_TILE_HD void go(Exec& ex, Shared& sh, int j0) {
    ex.run([&](Thread& t, int tid) { panel_step<S>(t, sh, j0, tid); });
    PanelSteps<S + 1, Exec>::go(ex, sh, j0);
  }
};
template <class Exec>
struct PanelSteps<PW + 1, Exec> {
  static B200_TILE_HD void go(Exec&, Shared&, int) {}
};

// Ain: column-major tile (leading dimension lda_in), lower triangle read; n <= 128 live rows / columns (the rest
// is treated as identity). CTA 0 stores L into Lout (leading dimension lda_out; must not alias Ain: the other CTAs
// read Ain while CTA 0 may already be storing) and raises info[0] when a pivot is not positive. Every CTA `cta`
// stores columns 16 cta .. 16 cta + 15 of L^-1 into Linv (128 x 128 column-major, strict upper part zero).
template <class Exec>
B200_TILE_HD void potrf_trinv_program(Exec& ex, Shared& sh, const double* Ain, int64_t lda_in, int n, double* Lout,
                                      int64_t lda_out, double* Linv, int cta, int* info) {
  // ---- load (two batches of 16 independent loads per thread) ----------------------------------------------
  ex.run([&](Thread& t, int tid) {
    t.bad = false;
    constexpr int BATCH = 16;
    for (int base = 0; base < PT * PT; base += THREADS * BATCH) {
      double v[BATCH];
#pragma unroll
      for (int u = 0; u < BATCH; ++u) {
        const int e = base + u * THREADS + tid, j = e >> 7, i = e & (PT - 1);
        v[u] = (i == j) ? 1.0 : 0.0;
        if (i < n && j < n && i >= j) v[u] = Ain[static_cast<int64_t>(j) * lda_in + i];
      }
#pragma unroll
      for (int u = 0; u < BATCH; ++u) {
        const int e = base + u * THREADS + tid, j = e >> 7, i = e & (PT - 1);
        sh.T[j * PLD + i] = (i >= j) ? v[u] : 0.0;
      }
    }
    if (tid < MAX_BLOCKS) {
      // block column jj (counted from the right) holds 4 (jj + 1) row pairs: it starts at 2 jj (jj + 1)
      int jj = static_cast<int>((std::sqrt(2.0 * tid + 1.0) - 1.0) * 0.5);
      while (2 * jj * (jj + 1) > tid) --jj;
      while (2 * (jj + 1) * (jj + 2) <= tid) ++jj;
      sh.bc[tid] = static_cast<unsigned char>(jj);
      sh.br[tid] = static_cast<unsigned char>(tid - 2 * jj * (jj + 1));  // row pair, counted from the bottom
    }
  });

  // ---- Cholesky, 16 columns at a time -----------------------------------------------------------------
#if defined(__CUDA_ARCH__)
#pragma unroll 1
#endif
  for (int p = 0; p < NPANEL; ++p) {
    const int j0 = p * PW, j1 = j0 + PW;
    PanelSteps<0, Exec>::go(ex, sh, j0);  // 17 phases: S = 16 only stores the last column
    if (j1 >= PT) break;
    // trailing update A(r, c) -= sum_k L(r, j0 + k) L(c, j0 + k) on the lower triangle c >= j1, r >= c: one block
    // of 2 rows x 8 columns per thread. Block columns are 8 wide and enumerated from the right, the row pairs of
    // a block column from the bottom (the table does not depend on the trailing size): consecutive threads take
    // consecutive row pairs, so the row operand and the read-modify-write of the block are contiguous 16-byte
    // accesses and the column operand is a broadcast.
    ex.run([&](Thread&, int tid) {
      const int m = PT - j1, q8 = m / 8;
      if (tid >= 2 * q8 * (q8 + 1)) return;
      const int c0 = j1 + 8 * (q8 - 1 - sh.bc[tid]), r0 = j1 + m - 2 * (sh.br[tid] + 1);  // r0 >= c0, both even
      double acc[2][8];
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[0][j] = acc[1][j] = 0.0;
#pragma unroll 4
      for (int k = 0; k < PW; ++k) {
        const double* col = sh.T + (j0 + k) * PLD;  // even offsets throughout: 16-byte aligned pairs
        const D2 rr = *reinterpret_cast<const D2*>(col + r0);
        double lc[8];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const D2 cc = *reinterpret_cast<const D2*>(col + c0 + 2 * j);
          lc[2 * j] = cc.x;
          lc[2 * j + 1] = cc.y;
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          acc[0][j] = fma(rr.x, lc[j], acc[0][j]);
          acc[1][j] = fma(rr.y, lc[j], acc[1][j]);
        }
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        D2* p2 = reinterpret_cast<D2*>(sh.T + (c0 + j) * PLD + r0);
        D2 a = *p2;
        a.x -= acc[0][j];
        a.y -= acc[1][j];
        *p2 = a;
      }
    });
  }

  // ---- store L; invert the diagonal 16 x 16 blocks; right-hand side E_J ------------------------------------
  const int J = cta, J0 = cta * PW;
  ex.run([&](Thread& t, int tid) {
    if (cta == 0) {
      for (int e = tid; e < PT * PT; e += THREADS) {
        const int j = e >> 7, i = e & (PT - 1);
        if (i >= j && i < n && j < n) Lout[static_cast<int64_t>(j) * lda_out + i] = sh.T[j * PLD + i];
      }
      if (tid == 0 && t.bad) info[0] = 1;
    }
    if (tid < NPANEL * PW) {
      const int I = tid >> 4, c = tid & (PW - 1);
      if (I >= J) {
        // column c of L_II^-1 by forward substitution (entries above c come out as exact zeros)
        const double* D = sh.T + (I * PW) * PLD + I * PW;  // D[k * PLD + i] = L_II(i, k)
        double x[PW];
#pragma unroll
        for (int i = 0; i < PW; ++i) {
          double s = (i == c) ? 1.0 : 0.0;
#pragma unroll
          for (int k = 0; k < i; ++k) s = fma(-D[k * PLD + i], x[k], s);
          x[i] = s * sh.invdiag[I * PW + i];
        }
#pragma unroll
        for (int i = 0; i < PW; ++i) sh.Dv[I][c][i] = x[i];
      }
    }
    for (int e = tid; e < PT * PW; e += THREADS) {
      const int r = e >> 4, c = e & (PW - 1);
      sh.R[r * XLD + c] = (r == J0 + c) ? 1.0 : 0.0;
      sh.X[r * XLD + c] = 0.0;
    }
  });

  // ---- block forward substitution L X = E_J -----------------------------------------------------------------
#if defined(__CUDA_ARCH__)
#pragma unroll 1
#endif
  for (int I = J; I < NPANEL; ++I) {
    // X_I = L_II^-1 R_I (the entries of L_II^-1 above the diagonal are stored zeros: no bound on k)
    ex.run([&](Thread&, int tid) {
      if (tid >= PW * PW) return;
      const int i = tid & (PW - 1), c = tid >> 4;
      double s[4] = {0.0, 0.0, 0.0, 0.0};  // four independent chains
#pragma unroll
      for (int k = 0; k < PW; ++k) s[k & 3] = fma(sh.Dv[I][k][i], sh.R[(I * PW + k) * XLD + c], s[k & 3]);
      sh.X[(I * PW + i) * XLD + c] = (s[0] + s[1]) + (s[2] + s[3]);
    });
    if (I + 1 >= NPANEL) break;
    // R_below -= L(below, I) X_I
    ex.run([&](Thread&, int tid) {
      const int first = (I + 1) * PW;
      for (int e = tid; e < (PT - first) * PW; e += THREADS) {
        const int c = e & (PW - 1), r = first + (e >> 4);
        double s[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
        for (int k = 0; k < PW; ++k) s[k & 3] = fma(sh.T[(I * PW + k) * PLD + r], sh.X[(I * PW + k) * XLD + c], s[k & 3]);
        sh.R[r * XLD + c] -= (s[0] + s[1]) + (s[2] + s[3]);
      }
    });
  }

  // ---- store the 16 columns of the inverse -------------------------------------------------------------------
  ex.run([&](Thread&, int tid) {
    for (int e = tid; e < PT * PW; e += THREADS) {
      const int r = e & (PT - 1), c = e >> 7;
      Linv[static_cast<int64_t>(J0 + c) * PT + r] = sh.X[r * XLD + c];
    }
  });
}

}  // namespace tile
}  // namespace b200ba
