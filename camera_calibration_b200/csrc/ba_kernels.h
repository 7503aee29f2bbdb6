// ba_kernels.h -- launchers of the kernels in ba_kernels.cu (internal to libb200ba.so).
#pragma once
#include <vector>

#include "ba_common.h"

namespace b200ba {

void launch_prepare_state(const ProblemDev& pb, const Layout& L, const StateDev& st, int64_t n_control_total,
                          cudaStream_t s);
// uniform_model: the model type shared by all cameras, or -1 for a mixed rig (runtime switch)
void launch_residual_jacobian(int uniform_model, bool jac, const ProblemDev& pb, const Layout& L,
                              const StateDev& st, double2* last_projection, const ObsOut& out, double huber,
                              uint32_t* straggler_list, int* straggler_count, cudaStream_t s,
                              cudaEvent_t main_done = nullptr);
void launch_straggler_pass(int uniform_model, bool jac, const ProblemDev& pb, const Layout& L, const StateDev& st,
                           double2* last_projection, const ObsOut& out, double huber, uint32_t* straggler_list,
                           int* straggler_count, cudaStream_t s);
void launch_accumulate_list(const ProblemDev& pb, const Layout& L, const StateDev& st, const ObsOut& out,
                            const SystemDev& sys, double huber, const uint32_t* list, const int* count, cudaStream_t s);
void launch_expand_jacobian(const ProblemDev& pb, const Layout& L, const StateDev& st, const ObsOut& out, double* jac,
                            cudaStream_t s);
// evaluation budget of the main pass before an observation is deferred to the straggler pass
void set_main_eval_budget(int budget);
void launch_accumulate_scatter(const ProblemDev& pb, const Layout& L, const StateDev& st, const ObsOut& out,
                               const SystemDev& sys, double huber, cudaStream_t s);
void launch_accumulate_cells(const ProblemDev& pb, const Layout& L, const StateDev& st, const ObsOut& out,
                             const SystemDev& sys, double huber, cudaStream_t s);
void launch_schur_blocks(int bs, int n_blocks, const double* Dblk, const double* bp, double lambda, double* Linv,
                         double* v, int* fail, cudaStream_t s);
void launch_schur_scale_rows(int bs, int n_blocks, int nd, const double* B, const double* Linv, double* W,
                             cudaStream_t s);
void launch_schur_backsub(int bs, int n_blocks, const double* Linv, const double* y, double* xp, cudaStream_t s);
void launch_group_support(int bs, int nblocks, int nd, const double* B, const int* group_of_block, uint8_t* flags,
                          cudaStream_t s);
void launch_compact_columns(int ngroups, int nd, const uint8_t* flags, int* cols, int* count, cudaStream_t s);
void launch_gather_scale(int bs, int nblocks_in_group, int nd, int m, int ldw, const double* B, const double* Linv,
                         const int* blocks, const int* cols, double* Wc, cudaStream_t s);
void launch_scatter_sub(int nd, int m, const int* cols, const double* P, double* S, cudaStream_t s);
void launch_block_solve_t(int bs, int n_blocks, const double* Linv, const double* v, double* u, cudaStream_t s);
void launch_block_backsub2(int bs, int n_blocks, const double* Linv, const double* u, const double* t, double* xb,
                           cudaStream_t s);
void launch_add_diagonal(int n, double* M, int64_t ld, double lambda, cudaStream_t s);
void launch_trace(int n_blocks, int bs, const double* Dblk, int nd, const double* C, double* out, cudaStream_t s);
void launch_update_state(const ProblemDev& pb, const Layout& L, const StateDev& src, const StateDev& dst,
                         const double* x, int64_t n_control_total, int64_t n_param_total, cudaStream_t s);
void launch_cost_reduce(int64_t n, const double* trial, const double* base, const double* residual, double* partial,
                        double* out, cudaStream_t s);
int cost_reduce_partial_size();
void launch_project_points(const CamDev& c, const double* intr, int64_t n, const double* lp, double* px, int32_t* ok,
                           cudaStream_t s);
void launch_unproject_pixels(const CamDev& c, const double* intr, int64_t n, const double* px, double* dirs,
                             double* origins, int32_t* ok, cudaStream_t s);
void launch_generic_block_inverse(int bs, int nb, int nd, const double* D, const double* B, const double* b1,
                                  double* DinvB, double* Dinvb, cudaStream_t s);
void launch_symmetrize(int n, double* M, cudaStream_t s);
// direction-grid fit (b200ba_fit_directions)
void launch_dirfit_tangents(int G, const double* grid, double* tan, cudaStream_t s);
void launch_dirfit(bool jac, int gw, int64_t n, const double* gp, const double* dirs, const double* grid,
                   const double* tan, double* H, double* b, int dof, double* cost, double* cost_sum, cudaStream_t s);
void launch_dirfit_update(int G, const double* grid, const double* x, double* out, cudaStream_t s);
void launch_permute_double2(int64_t n, const uint32_t* perm, const double2* src, double2* dst, bool scatter,
                            cudaStream_t s);


void launch_mask_fixed(const Layout& L, const SystemDev& sys, const FixedRanges& fr, double unit, cudaStream_t s);
// ChooseNiceCameraOrientation + Rotate + camera_tr_rig update for every camera (rot: [9 * n_cameras] scratch)
void launch_nice_orientation(const ProblemDev& pb, const StateDev& st, int n_cameras, double* rot, cudaStream_t s);

// ---- dense phase (ba_dense.cu) ----------------------------------------------------------------
// Storage map of the reduced system S (n_d x n_d, column-major, lower triangle valid): columns are
// grouped in blocks of `nb`; block j lives in slot (j % ranks) * blocks_per_rank + j / ranks, so that
// the column blocks one rank owns in the block-cyclic distribution are CONTIGUOUS (one chunk of a
// reduce-scatter). With one rank the map is the identity. Rows are never permuted.
struct DenseMap {
  int nb;               // column-block width
  int ranks;            // R
  int blocks_per_rank;  // ceil(n_blocks / R)
  int64_t ld;           // leading dimension (>= n_d, even)
  __host__ __device__ __forceinline__ int slot(int block) const { return (block % ranks) * blocks_per_rank + block / ranks; }
  __host__ __device__ __forceinline__ int64_t col_offset(int c) const {
    const int b = c / nb;
    return (static_cast<int64_t>(slot(b)) * nb + (c - b * nb)) * ld;
  }
};

// C (+)= alpha * A B^T; A(i, k) at A[k * lda + i], B(j, k) at B[k * ldb + j], C(i, j) at C[j * ldc + i].
struct GemmArgs {
  int M, N, K;
  const double* A;
  int64_t lda;
  const double* B;
  int64_t ldb;
  double* C;
  int64_t ldc;
  double alpha, beta;
  bool a_aligned, b_aligned;  // 16-byte copies allowed (set by make_gemm_args)
  // scatter epilogue: C is the base of S, compact row / column i is dense column cols[i] (ascending)
  const int* cols;
  DenseMap map;
  int64_t n_tiles_lower;  // set by launch_dgemm_nt
  // owned-only trailing update of the block-cyclic factorisation: C = base of S, column n of the product is
  // global column col_base + n (rows likewise); tiles of column blocks owned by other ranks are skipped
  bool owned_only;
  int rank, col_base;
  bool in_place;  // C aliases A (N = K <= 128): forces one 128-wide tile per row block
};
inline bool gemm_operand_aligned(const double* p, int64_t ld) {
  return (reinterpret_cast<uintptr_t>(p) % 16 == 0) && (ld % 2 == 0);
}
// lower: only tiles / entries with i >= j (M == N); scatter: the scatter-subtract epilogue (implies lower).
// leave_sms: cap the persistent grid at (SM count - reserve) so that concurrently running panel kernels find
// a free SM at once (set_gemm_sm_reserve; used for the trailing updates of the factorisation).
int launch_dgemm_nt(const GemmArgs& g, bool lower, bool scatter, cudaStream_t s, bool leave_sms = false,
                    bool one_tile_per_cta = false);
void set_gemm_sm_reserve(int n);
// Cholesky of a 128 x 128 (live size n) column-major diagonal tile in place + Linv [128 x 128, ld 128].
int launch_potrf_tile(double* A, int64_t lda, int n, double* Linv, int* info, cudaStream_t s);

// y += alpha B^T u (two stages, fixed order; partial: gemv_t_partial_size doubles) and t = B x for row-major B
int gemv_t_partial_size(int rows, int cols);
void launch_gemv_t(int rows, int cols, int64_t ld, const double* B, const double* u, double alpha, double* y, double* partial,
                   cudaStream_t s);
void launch_gemv_n(int rows, int cols, int64_t ld, const double* B, const double* x, double* t, cudaStream_t s);
void launch_add_diagonal_map(int n, double* S, const DenseMap& map, double lambda, cudaStream_t s);

// Context of the dense factorisation / solve (ba_dense.cu). The caller owns the buffers.
struct DenseCtx {
  int n = 0, NB = 256, nblk = 0, ntiles = 0;
  int rank = 0, ranks = 1;
  DenseMap map{};
  int64_t chunk = 0;                 // doubles per rank of the S allocation (reduce-scatter chunk)
  std::vector<int64_t> panel_off;    // [nblk + 1] offsets of the packed panels in Lpack
  std::vector<int> panel_h;          // [nblk] leading dimension (even) of each packed panel
  double* S = nullptr;               // ranks * chunk doubles
  double* Lpack = nullptr;           // panel_off[nblk] doubles: the factor; each panel is followed by the explicit
                                     // inverses of its NB / 128 diagonal tiles ([128 * 128] each)
  double* tmp = nullptr;             // [n] intermediate of the triangular solves
  int64_t* d_panel_off = nullptr;    // device copies for the backward step
  int* d_panel_h = nullptr;
  int* info = nullptr;               // device flag: non-positive pivot
  cudaStream_t s_main = nullptr, s_panel = nullptr, s_aux = nullptr;
  cudaEvent_t ev_ready[2] = {nullptr, nullptr}, ev_main[2] = {nullptr, nullptr}, ev_half2[2] = {nullptr, nullptr}, ev_misc = nullptr;
  // broadcast of `count` doubles from rank `root` on stream s (multi-GPU only)
  int (*bcast)(double* buf, size_t count, int root, cudaStream_t s, void* user) = nullptr;
  void* user = nullptr;
};
int dense_plan(DenseCtx* d, int n, int nb, int rank, int ranks);
int dense_factor(DenseCtx* d);
int dense_solve(DenseCtx* d, double* b);

}  // namespace b200ba
