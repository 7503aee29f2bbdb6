// ba_kernels.h -- launchers of the kernels in ba_kernels.cu (internal to libb200ba.so).
#pragma once
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
void launch_accumulate_list(const ProblemDev& pb, const Layout& L, const ObsOut& out, const SystemDev& sys,
                            double huber, const uint32_t* list, const int* count, cudaStream_t s);
// evaluation budget of the main pass before an observation is deferred to the straggler pass
void set_main_eval_budget(int budget);
void launch_accumulate_scatter(const ProblemDev& pb, const Layout& L, const ObsOut& out, const SystemDev& sys,
                               double huber, cudaStream_t s);
void launch_accumulate_cells(const ProblemDev& pb, const Layout& L, const ObsOut& out, const SystemDev& sys,
                             double huber, cudaStream_t s);
void launch_schur_blocks(int bs, int n_blocks, const double* Dblk, const double* bp, double lambda, double* Linv,
                         double* v, int* fail, cudaStream_t s);
void launch_schur_scale_rows(int bs, int n_blocks, int nd, const double* B, const double* Linv, double* W,
                             cudaStream_t s);
void launch_schur_backsub(int bs, int n_blocks, const double* Linv, const double* y, double* xp, cudaStream_t s);
void launch_group_support(int bs, int nblocks, int nd, const double* B, const int* group_of_block, uint8_t* flags,
                          cudaStream_t s);
void launch_compact_columns(int ngroups, int nd, const uint8_t* flags, int* cols, int* count, cudaStream_t s);
void launch_gather_scale(int bs, int nblocks_in_group, int nd, int m, const double* B, const double* Linv,
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

}  // namespace b200ba
