// Host emulation of the diagonal-tile program of the blocked Cholesky (camera_calibration_b200/csrc/ba_tile.cuh):
// the same template the CUDA kernel instantiates, run with an executor that calls each phase for all 512 threads of
// a CTA in forward, reversed and shuffled order. Equal results in all orders = no phase reads what another thread
// writes in the same phase (the barrier placement is right); agreement with a plain Cholesky / triangular inverse =
// the index arithmetic is right. Built and run by tests/test_tile_emulation.py (no GPU involved).
#include <algorithm>
#include <cstdio>
#include <cstring>
#include <memory>
#include <numeric>
#include <random>
#include <vector>

#include "../camera_calibration_b200/csrc/ba_tile.cuh"

using namespace b200ba::tile;

struct HostExec {
  std::vector<Thread> t = std::vector<Thread>(THREADS);
  std::vector<int> order = std::vector<int>(THREADS);
  template <class F>
  void run(F f) {
    for (int k = 0; k < THREADS; ++k) f(t[order[k]], order[k]);
  }
};

static void reference(const std::vector<double>& A, int lda, int n, std::vector<double>* L, std::vector<double>* Linv) {
  // identity-padded lower Cholesky and its inverse, plain loops
  std::vector<double> M(PT * PT, 0.0);
  for (int j = 0; j < PT; ++j)
    for (int i = j; i < PT; ++i) M[j * PT + i] = (i < n && j < n) ? A[static_cast<size_t>(j) * lda + i] : (i == j ? 1.0 : 0.0);
  for (int j = 0; j < PT; ++j) {
    double d = M[j * PT + j];
    for (int k = 0; k < j; ++k) d -= M[k * PT + j] * M[k * PT + j];
    d = std::sqrt(d);
    M[j * PT + j] = d;
    for (int i = j + 1; i < PT; ++i) {
      double s = M[j * PT + i];
      for (int k = 0; k < j; ++k) s -= M[k * PT + i] * M[k * PT + j];
      M[j * PT + i] = s / d;
    }
  }
  *L = M;
  Linv->assign(PT * PT, 0.0);
  for (int c = 0; c < PT; ++c)
    for (int i = c; i < PT; ++i) {
      double s = (i == c) ? 1.0 : 0.0;
      for (int k = c; k < i; ++k) s -= M[k * PT + i] * (*Linv)[c * PT + k];
      (*Linv)[c * PT + i] = s / M[i * PT + i];
    }
}

int main() {
  std::mt19937_64 rng(7);
  std::normal_distribution<double> gauss(0.0, 1.0);
  const int lda_in = 140, lda_out = 150;
  std::unique_ptr<Shared> sh(new Shared);
  double worst_l = 0, worst_inv = 0;
  for (int n : {128, 77, 16, 5}) {
    // SPD: G G^T + n I on the live part
    std::vector<double> G(n * n), A(static_cast<size_t>(lda_in) * PT, std::nan(""));
    for (double& g : G) g = gauss(rng);
    for (int j = 0; j < n; ++j)
      for (int i = 0; i < n; ++i) {
        double s = (i == j) ? n : 0.0;
        for (int k = 0; k < n; ++k) s += G[k * n + i] * G[k * n + j];
        A[static_cast<size_t>(j) * lda_in + i] = s;
      }
    std::vector<double> Lref, Iref;
    reference(A, lda_in, n, &Lref, &Iref);
    std::vector<double> L_first, Inv_first;
    for (int mode = 0; mode < 3; ++mode) {
      std::vector<double> Lout(static_cast<size_t>(lda_out) * PT, -7.0), Linv(PT * PT, std::nan(""));
      int info = 0;
      for (int cta = 0; cta < CTAS; ++cta) {
        HostExec ex;
        std::iota(ex.order.begin(), ex.order.end(), 0);
        if (mode == 1) std::reverse(ex.order.begin(), ex.order.end());
        if (mode == 2) std::shuffle(ex.order.begin(), ex.order.end(), rng);
        std::memset(static_cast<void*>(sh.get()), 0xff, sizeof(Shared));  // NaN patterns: nothing may rely on zeroed memory
        potrf_trinv_program(ex, *sh, A.data(), lda_in, n, Lout.data(), lda_out, Linv.data(), cta, &info);
      }
      if (info != 0) { std::printf("info raised on an SPD tile (n %d)\n", n); return 1; }
      for (int j = 0; j < PT; ++j)
        for (int i = 0; i < PT; ++i) {
          const double got = Lout[static_cast<size_t>(j) * lda_out + i];
          if (i >= j && i < n && j < n) {
            worst_l = std::max(worst_l, std::fabs(got - Lref[j * PT + i]) / std::max(1.0, std::fabs(Lref[j * PT + i])));
          } else if (got != -7.0) {
            std::printf("L stored outside the live lower triangle at (%d, %d), n %d\n", i, j, n);
            return 1;
          }
          const double inv = Linv[j * PT + i];
          if (!(std::fabs(inv - Iref[j * PT + i]) <= 1e-11 * std::max(1.0, std::fabs(Iref[j * PT + i])))) {
            std::printf("inverse differs at (%d, %d), n %d: %.17g vs %.17g\n", i, j, n, inv, Iref[j * PT + i]);
            return 1;
          }
          if (i < j && inv != 0.0) { std::printf("inverse not zero above the diagonal\n"); return 1; }
          worst_inv = std::max(worst_inv, std::fabs(inv - Iref[j * PT + i]));
        }
      if (mode == 0) {
        L_first = Lout;
        Inv_first = Linv;
      } else if (std::memcmp(L_first.data(), Lout.data(), Lout.size() * sizeof(double)) != 0 ||
                 std::memcmp(Inv_first.data(), Linv.data(), Linv.size() * sizeof(double)) != 0) {
        std::printf("result depends on the thread order within a phase (mode %d, n %d): missing barrier\n", mode, n);
        return 1;
      }
    }
  }
  if (worst_l > 1e-12) { std::printf("factor differs from the reference: %.3e\n", worst_l); return 1; }
  // an indefinite tile must raise info
  {
    std::vector<double> A(static_cast<size_t>(lda_in) * PT, 0.0);
    for (int j = 0; j < PT; ++j) A[static_cast<size_t>(j) * lda_in + j] = (j == 40) ? -1.0 : 2.0;
    std::vector<double> Lout(static_cast<size_t>(lda_out) * PT), Linv(PT * PT);
    int info = 0;
    HostExec ex;
    std::iota(ex.order.begin(), ex.order.end(), 0);
    potrf_trinv_program(ex, *sh, A.data(), lda_in, PT, Lout.data(), lda_out, Linv.data(), 0, &info);
    if (info != 1) { std::printf("indefinite tile not reported\n"); return 1; }
  }
  std::printf("TILE_EMULATION_OK max rel factor error %.3e, max abs inverse error %.3e, shared bytes %zu\n", worst_l, worst_inv,
              sizeof(Shared));
  return 0;
}
