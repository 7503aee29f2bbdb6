// ba_tile.cuh -- the 128 x 128 diagonal-tile step of the blocked Cholesky (ba_dense.cu): factor the tile and
// invert the factor, written as a sequence of barrier-separated PHASES over the 512 threads of a CTA.
//
// Where this sits: every 128 columns, the factorisation's critical path runs  tile Cholesky -> inverse of the
// factor -> panel solve (a product with that inverse). The first version (potrf_tile_kernel / trinv_tile_kernel in
// ba_dense.cu, still selectable) keeps a 32-column register window per thread and issues 32 shared-memory loads
// per thread and column step: 100 k cycles per tile, LSU-bound, and a second launch of 45 k cycles for the inverse.
// Here both are blocked by 16 columns:
//   * Cholesky: per 16-column panel, 16 column steps on a 128 x 16 register panel (4 entries per thread, one
//     barrier per step through a double-buffered published column), then ONE rank-16 update of the trailing
//     lower triangle with a 4 x 4 register block per thread (at most 406 blocks: a single pass).
//   * inverse: the launch has 8 CTAs; each repeats the (deterministic) factorisation and then solves L X = E_J
//     for its own 16 columns J by block forward substitution with the explicitly inverted 16 x 16 diagonal
//     blocks. CTA 0 stores the factor. One launch instead of two, no reload of the tile.
//
// The program is a template over an executor: on the device `run(f)` calls f for this thread and then
// __syncthreads(); the host executor of tests/tile_emulation.cc calls f for all 512 threads of the CTA in a
// forward or reversed order -- every phase must give the same result in both (no thread may read what another
// thread writes in the same phase), which pins the barrier placement and all index arithmetic without a GPU.
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
constexpr int MAX_BLOCKS = (PT - PW) / 4 * ((PT - PW) / 4 + 1) / 2;  // 406 register blocks in the first trailing update

struct alignas(16) D2 {
  double x, y;
};

struct alignas(16) Shared {
  double T[PT * PLD];        // the tile, column-major: (i, j) at T[j * PLD + i]
  double cb[2][PT + 2];      // published column of the current step (un-scaled) and, at [PT], 1 / sqrt(pivot)
  double R[PT * XLD];        // inverse: right-hand side rows x 16 columns
  double X[PT * XLD];        // inverse: solution rows x 16 columns
  double Dv[NPANEL][PW][PW]; // Dv[I][c][i] = (L_II^-1)(i, c) for the diagonal 16 x 16 blocks
  unsigned char bi[MAX_BLOCKS + 2], bj[MAX_BLOCKS + 2];  // triangular enumeration of the 4 x 4 register blocks
};

struct Thread {
  double pa[4];  // this thread's entries of the current 128 x 16 panel: row tid & 127, columns (tid >> 7) + 4 q
  bool bad;      // a pivot was not positive
};

B200_TILE_HD double inv_sqrt(double x) {
#if defined(__CUDA_ARCH__)
  return rsqrt(x);
#else
  return 1.0 / std::sqrt(x);
#endif
}

// applies column jp (published in cbp) to the thread's panel entries and stores the finished column
B200_TILE_HD void consume_column(Thread& t, Shared& sh, const double* cbp, int jp, int j0, int r, int cg, int owner_group) {
  const double d = cbp[jp], rs = cbp[PT];
  t.bad |= !(d > 0.0);
  const double lrj = cbp[r] * (rs * rs);  // a(r, jp) / a(jp, jp)
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const int c = j0 + cg + 4 * q;
    if (c > jp) t.pa[q] = fma(-lrj, cbp[c], t.pa[q]);
  }
  if (cg == owner_group && r >= jp) sh.T[jp * PLD + r] = cbp[r] * rs;  // L(r, jp)
}

template <int S, class Exec>
struct PanelSteps {
  static B200_TILE_HD void go(Exec& ex, Shared& sh, int j0) {
    ex.run([&](Thread& t, int tid) {
      const int r = tid & (PT - 1), cg = tid >> 7;
      if (S == 0) {
#pragma unroll
        for (int q = 0; q < 4; ++q) t.pa[q] = sh.T[(j0 + cg + 4 * q) * PLD + r];
      } else {
        consume_column(t, sh, sh.cb[(S - 1) & 1], j0 + S - 1, j0, r, cg, (S - 1) & 3);
      }
      // publish column j0 + S (owned by thread group S & 3, register S >> 2)
      if (cg == (S & 3)) {
        const int j = j0 + S;
        double* cbn = sh.cb[S & 1];
        cbn[r] = (r >= j) ? t.pa[S >> 2] : 0.0;
        if (r == j) cbn[PT] = inv_sqrt(t.pa[S >> 2]);
      }
    });
    PanelSteps<S + 1, Exec>::go(ex, sh, j0);
  }
};
template <class Exec>
struct PanelSteps<PW, Exec> {
  static B200_TILE_HD void go(Exec&, Shared&, int) {}
};

// Ain: column-major tile (leading dimension lda_in), lower triangle read; n <= 128 live rows / columns (the rest
// is treated as identity). CTA 0 stores L into Lout (leading dimension lda_out; must not alias Ain: the other CTAs
// read Ain while CTA 0 may already be storing) and raises info[0] when a pivot is not positive. Every CTA `cta`
// stores columns 16 cta .. 16 cta + 15 of L^-1 into Linv (128 x 128 column-major, strict upper part zero).
template <class Exec>
B200_TILE_HD void potrf_trinv_program(Exec& ex, Shared& sh, const double* Ain, int64_t lda_in, int n, double* Lout,
                                      int64_t lda_out, double* Linv, int cta, int* info) {
  // ---- load ------------------------------------------------------------------------------------------
  ex.run([&](Thread& t, int tid) {
    t.bad = false;
    for (int e = tid; e < PT * PT; e += THREADS) {
      const int j = e >> 7, i = e & (PT - 1);
      double v = (i == j) ? 1.0 : 0.0;
      if (i < n && j < n && i >= j) v = Ain[static_cast<int64_t>(j) * lda_in + i];
      sh.T[j * PLD + i] = (i >= j) ? v : 0.0;
    }
    if (tid < MAX_BLOCKS) {
      // largest b with b (b + 1) / 2 <= tid
      int b = static_cast<int>((std::sqrt(8.0 * tid + 1.0) - 1.0) * 0.5);
      while (b * (b + 1) / 2 > tid) --b;
      while ((b + 1) * (b + 2) / 2 <= tid) ++b;
      sh.bi[tid] = static_cast<unsigned char>(b);                        // bi >= bj
      sh.bj[tid] = static_cast<unsigned char>(tid - b * (b + 1) / 2);
    }
  });

  // ---- Cholesky, 16 columns at a time -----------------------------------------------------------------
#if defined(__CUDA_ARCH__)
#pragma unroll 1
#endif
  for (int p = 0; p < NPANEL; ++p) {
    const int j0 = p * PW, j1 = j0 + PW;
    PanelSteps<0, Exec>::go(ex, sh, j0);
    // last column of the panel: nothing left to update inside the panel, store it
    ex.run([&](Thread& t, int tid) {
      consume_column(t, sh, sh.cb[(PW - 1) & 1], j1 - 1, j0, tid & (PT - 1), tid >> 7, (PW - 1) & 3);
    });
    if (j1 >= PT) break;
    // trailing update A(r, c) -= sum_k L(r, j0 + k) L(c, j0 + k) on the lower triangle c >= j1, r >= c:
    // one 4 x 4 block per thread. The blocks are enumerated from the bottom-right corner, column by column
    // (the table is independent of the trailing size): consecutive threads walk down a block column, so the
    // column operand is a broadcast and the row operand / the read-modify-write of the block are 16-byte
    // accesses 32 bytes apart.
    ex.run([&](Thread&, int tid) {
      const int q = (PT - j1) / 4;
      if (tid >= q * (q + 1) / 2) return;
      const int r0 = j1 + 4 * (q - 1 - sh.bj[tid]), c0 = j1 + 4 * (q - 1 - sh.bi[tid]);  // bj <= bi: r0 >= c0
      double acc[4][4];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.0;
#pragma unroll 4
      for (int k = 0; k < PW; ++k) {
        const double* col = sh.T + (j0 + k) * PLD;  // even offsets throughout: 16-byte aligned pairs
        const D2 r01 = *reinterpret_cast<const D2*>(col + r0), r23 = *reinterpret_cast<const D2*>(col + r0 + 2);
        const D2 c01 = *reinterpret_cast<const D2*>(col + c0), c23 = *reinterpret_cast<const D2*>(col + c0 + 2);
        const double lr[4] = {r01.x, r01.y, r23.x, r23.y}, lc[4] = {c01.x, c01.y, c23.x, c23.y};
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[i][j] = fma(lr[i], lc[j], acc[i][j]);
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        D2* p = reinterpret_cast<D2*>(sh.T + (c0 + j) * PLD + r0);
        D2 a = p[0], b = p[1];
        a.x -= acc[0][j];
        a.y -= acc[1][j];
        b.x -= acc[2][j];
        b.y -= acc[3][j];
        p[0] = a;
        p[1] = b;
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
          x[i] = s / D[i * PLD + i];
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
    // X_I = L_II^-1 R_I
    ex.run([&](Thread&, int tid) {
      if (tid >= PW * PW) return;
      const int i = tid & (PW - 1), c = tid >> 4;
      double s = 0.0;
      for (int k = 0; k <= i; ++k) s = fma(sh.Dv[I][k][i], sh.R[(I * PW + k) * XLD + c], s);
      sh.X[(I * PW + i) * XLD + c] = s;
    });
    if (I + 1 >= NPANEL) break;
    // R_below -= L(below, I) X_I
    ex.run([&](Thread&, int tid) {
      const int first = (I + 1) * PW;
      for (int e = tid; e < (PT - first) * PW; e += THREADS) {
        const int c = e & (PW - 1), r = first + (e >> 4);
        double s = 0.0;
#pragma unroll
        for (int k = 0; k < PW; ++k) s = fma(sh.T[(I * PW + k) * PLD + r], sh.X[(I * PW + k) * XLD + c], s);
        sh.R[r * XLD + c] -= s;
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
