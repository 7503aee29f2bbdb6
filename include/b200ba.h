/*
 * b200ba.h -- C ABI of the B200-native joint-optimisation (bundle-adjustment) path.
 *
 * This is the drop-in boundary for ONE hot path of puzzlepaint/camera_calibration:
 *   double OptimizeJointly(Dataset&, BAState*, ...)            (reference:
 *     applications/camera_calibration/src/camera_calibration/bundle_adjustment/joint_optimization.h:53-70,
 *     implementation joint_optimization.cc:757-953)
 * and its GPU twin
 *   OptimizationReport CudaOptimizeJointly(...)               (reference:
 *     bundle_adjustment/cuda_joint_optimization.h:45-59).
 *
 * The reference has no C ABI; its seam is the C++ free function above plus the
 * CameraModel plugin (models/camera_model.h:42-204). A maintainer binds this
 * library by flattening Dataset / BAState into the POD structs below (see
 * INTEGRATION.md and include/b200ba_shim.hpp, which does exactly that behind the
 * reference's own signature).
 *
 * Conventions
 *   - plain pointers and sizes only; no C++ / torch types cross this line.
 *   - all functions return 0 on success, non-zero on error; b200ba_last_error()
 *     gives the message. Nothing here aborts (the reference CHECK()s).
 *   - the caller owns every host array for the duration of the call only; the
 *     handle owns all device memory, streams, cuBLAS/cuSOLVER handles and the
 *     NCCL communicator.
 *   - a handle is single-owner / not re-entrant, like the reference's cost
 *     function (models/central_grid.h:186).
 *   - quaternions are (w, x, y, z) followed by translation (x, y, z): the order
 *     of the reference's CUDA state upload (cuda_joint_optimization.cc:88-112).
 *     Eigen's coeffs() order is x,y,z,w -- convert at the shim.
 *   - pixels use the "pixel-corner" convention (camera_model.h:67-70).
 */
#ifndef B200BA_H_
#define B200BA_H_

#include <stdint.h>

#if defined(__GNUC__)
#define B200BA_API __attribute__((visibility("default")))
#else
#define B200BA_API
#endif

#ifdef __cplusplus
extern "C" {
#endif

/* Values equal CameraModel::Type (models/camera_model.h:44-54). */
enum {
  B200BA_MODEL_CENTRAL_GENERIC = 0,
  B200BA_MODEL_NONCENTRAL_GENERIC = 1,
  B200BA_MODEL_CENTRAL_THIN_PRISM_FISHEYE = 2, /* not on the accelerated path */
  B200BA_MODEL_CENTRAL_OPENCV = 3,
  B200BA_MODEL_CENTRAL_RADIAL = 4 /* not on the accelerated path */
};

/* Values equal SchurMode (libvis lm_optimizer.h). All modes are numerically
 * interchangeable in the reference (test/central_generic_test.cc:77-91); this
 * library always stores the off-diagonal part densely on the device. */
enum {
  B200BA_SCHUR_DENSE = 0,
  B200BA_SCHUR_DENSE_CUDA = 1,
  B200BA_SCHUR_DENSE_ONTHEFLY = 2,
  B200BA_SCHUR_SPARSE = 3,
  B200BA_SCHUR_SPARSE_ONTHEFLY = 4
};

enum {
  B200BA_JACOBIAN_NUMERIC = 0, /* reference behaviour (finite differences); CPU oracle only */
  B200BA_JACOBIAN_ANALYTIC = 1 /* implicit-function-theorem Jacobian; the GPU path */
};

/* Static description of one camera (CameraModel base members,
 * models/camera_model.h:189-203, plus the grid resolution of the generic models). */
typedef struct b200ba_camera {
  int32_t model_type;
  int32_t width, height;
  int32_t calibration_min_x, calibration_min_y;
  int32_t calibration_max_x, calibration_max_y;
  int32_t grid_width, grid_height; /* 0 for parametric models */
} b200ba_camera;

/* Number of doubles in the flat intrinsics array of a camera:
 *   central-generic   : 3*gw*gh               (grid, row-major, xyz; central_grid.h m_grid)
 *   noncentral-generic: 3*gw*gh direction grid followed by 3*gw*gh point grid
 *   central-opencv    : 12  (fx fy cx cy k1..k6 p1 p2; central_opencv.cc:77-88)  */
B200BA_API int64_t b200ba_intrinsics_size(const b200ba_camera* cam);
/* update_parameter_count() of the model: 2*gw*gh / 5*gw*gh / 12. */
B200BA_API int32_t b200ba_update_parameter_count(const b200ba_camera* cam);

/* The constant part of a BA problem: the flattened Dataset (dataset.h:57-212)
 * restricted to used imagesets. Observations are listed in the reference's
 * residual order: imageset-major, then camera, then feature order
 * (joint_optimization.cc:273-290); obs_imageset must be non-decreasing and, within
 * an imageset, obs_camera non-decreasing. */
typedef struct b200ba_problem {
  int32_t n_cameras;
  const b200ba_camera* cameras;
  int32_t n_imagesets; /* used imagesets (sequential indices) */
  int32_t n_points;
  int64_t n_obs;
  const uint32_t* obs_imageset; /* [n_obs] */
  const uint32_t* obs_camera;   /* [n_obs] */
  const uint32_t* obs_point;    /* [n_obs]  PointFeature::index */
  const float* obs_xy;          /* [2*n_obs] PointFeature::xy (float, dataset.h:67) */
} b200ba_problem;

/* The optimised part of BAState (ba_state.h:79-96) + the warm-start cache
 * PointFeature::last_projection (dataset.h:79-83). Used for both input and output. */
typedef struct b200ba_state {
  double* points;            /* [3*n_points] */
  double* rig_tr_global;     /* [7*n_imagesets]  qw qx qy qz tx ty tz */
  double* camera_tr_rig;     /* [7*n_cameras] */
  double* const* intrinsics; /* [n_cameras] -> b200ba_intrinsics_size() doubles each */
  double* last_projection;   /* [2*n_obs]; may be NULL (= all zero on input, dropped on output) */
} b200ba_state;

/* Arguments of OptimizeJointly (joint_optimization.h:53-70) + the constants it
 * hard-codes (joint_optimization.cc:916-923, :346). */
typedef struct b200ba_options {
  int32_t max_iteration_count;
  double init_lambda;           /* < 0: initialise from H (lm_optimizer.h:766-781) */
  double numerical_diff_delta;  /* used by the numeric Jacobian mode only */
  double regularization_weight; /* the reference ignores it with an error log (:299-305); must be 0 */
  int32_t localize_only;
  int32_t eliminate_points;
  int32_t schur_mode;
  int32_t max_lm_attempts;      /* reference: 50 */
  double init_lambda_factor;    /* reference: 1e-5 */
  double huber_parameter;       /* reference: 1.0 */
  int32_t jacobian_mode;        /* B200BA_JACOBIAN_* */
  int32_t print_progress;
  /* the debug_* switches of OptimizeJointly (joint_optimization.h:63-68):
   *   debug_verify_cost: before optimising, the cost is evaluated without and with Jacobians, twice
   *     (LMOptimizer::VerifyCost, lm_optimizer.h:474-490); the call fails with code 5 where the
   *     reference CHECKs |cost1 - cost2| <= 1e-3 (joint_optimization.cc:866-876).
   *   debug_fix_*: the variables of the group are held fixed (LMOptimizer::FixVariable,
   *     joint_optimization.cc:878-903): their update is 0 and they are removed from the linear system
   *     (lm_optimizer.h:1069-1121). The reference then solves the thinned system densely; here the rows /
   *     columns are masked out of H, b and the Schur-complement path is kept (same solution). */
  int32_t debug_verify_cost;
  int32_t debug_fix_points;
  int32_t debug_fix_poses;
  int32_t debug_fix_rig_poses;
  int32_t debug_fix_intrinsics;
} b200ba_options;

B200BA_API void b200ba_default_options(b200ba_options* opt);

#define B200BA_MAX_TRACE 128

/* OptimizationReport (libvis lm_optimizer.h:55-77) + what OptimizeJointly returns
 * through its out-parameters, + a per-iteration trace for parity checks. */
typedef struct b200ba_report {
  double initial_cost;
  double final_cost;
  double final_lambda;
  int32_t num_iterations_performed;
  int32_t performed_an_iteration;
  double cost_and_jacobian_evaluation_time; /* seconds */
  double solve_time;                        /* seconds */
  int64_t n_valid;    /* residuals valid at the final state */
  int64_t n_invalid;
  double rmse;        /* sqrt(sum |pixel - xy|^2 / n_valid) at the final state */
  int32_t trace_len;
  double trace_cost[B200BA_MAX_TRACE];    /* cost after each outer iteration */
  double trace_lambda[B200BA_MAX_TRACE];  /* lambda after each outer iteration */
  int32_t trace_attempts[B200BA_MAX_TRACE]; /* LM attempts used by each outer iteration */
} b200ba_report;

typedef struct b200ba_handle b200ba_handle;

/* ---- lifetime ---------------------------------------------------------- */
/* device < 0: use the current CUDA device. Copies the problem to the device. */
B200BA_API int b200ba_create(const b200ba_problem* problem, int device, b200ba_handle** out);
B200BA_API void b200ba_destroy(b200ba_handle* h);
B200BA_API const char* b200ba_last_error(const b200ba_handle* h); /* h may be NULL: last create error */

/* ---- state transfer ------------------------------------------------------ */
B200BA_API int b200ba_set_state(b200ba_handle* h, const b200ba_state* state);
B200BA_API int b200ba_get_state(b200ba_handle* h, b200ba_state* state);

/* Device-side copy of the optimised state + last_projection held by the handle into a snapshot
 * slot / back (no host transfer). The reference keeps such a copy implicitly: LMOptimizer works
 * on `State test_state = *state` (libvis lm_optimizer.h:868) and the product writes a checkpoint
 * after every iteration (calibration.cc:240-243). */
B200BA_API int b200ba_snapshot_state(b200ba_handle* h);
B200BA_API int b200ba_restore_state(b200ba_handle* h);

/* ---- the hot path -------------------------------------------------------- */
/* Equivalent of OptimizeJointly (joint_optimization.cc:757-953) on the state held
 * by the handle; the state stays resident on the device between calls. */
B200BA_API int b200ba_optimize(b200ba_handle* h, const b200ba_options* opt, b200ba_report* report);

/* Convenience: set_state + optimize + get_state with HOST buffers, i.e. exactly
 * what a call of the reference's OptimizeJointly(Dataset&, BAState*) does. */
B200BA_API int b200ba_optimize_host(b200ba_handle* h, b200ba_state* state, const b200ba_options* opt,
                         b200ba_report* report);

/* RunBundleAdjustment (calibration.cc:187-304) on the state held by the handle: up to
 * max_iteration_count single LM iterations (lambda carried over, starting from opt->init_lambda, the
 * reference passes -1), after each one ChooseNiceCameraOrientation + Rotate + the camera_tr_rig update
 * for every camera (calibration.cc:245-252, models/central_generic.cc:570-621; skipped when
 * localize_only), stop as soon as cost >= last_cost - cost_reduction_threshold (:298-300). The state
 * never leaves the device; on_iteration (nullable) is called after every iteration -- the place for
 * the reference's per-iteration SaveBAState checkpoint via b200ba_get_state (:240-243) -- and stops
 * the loop by returning non-zero (the reference's 'q' key). With several ranks the call is collective. */
typedef struct b200ba_ba_report {
  double initial_cost, final_cost, final_lambda, rmse;
  int64_t n_valid, n_invalid;
  int32_t iterations;    /* LM iterations run (OptimizeJointly calls) */
  int32_t lm_attempts;   /* linear solves over all of them */
  double device_ms;      /* sum of the iterations' device times */
  double costs[B200BA_MAX_TRACE]; /* cost after each iteration */
} b200ba_ba_report;
B200BA_API int b200ba_run_bundle_adjustment(b200ba_handle* h, const b200ba_options* opt, int32_t max_iteration_count,
                                            double cost_reduction_threshold, b200ba_ba_report* report,
                                            int (*on_iteration)(void* user, int32_t iteration, double cost),
                                            void* user);

/* ---- building blocks, exposed for parity tests and profiling -------------- */
/* One pass of JointOptimizationCostFunction::Compute<compute_jacobians>
 * (joint_optimization.cc:240-306) at the current state.
 *   residuals   [2*n_obs] out, nullable   pixel - xy
 *   costs       [n_obs]   out, nullable   Huber cost, -1 for invalid residuals
 *   total_cost            out, nullable
 * Updates last_projection on the device like the reference mutates the Dataset. */
B200BA_API int b200ba_evaluate(b200ba_handle* h, const b200ba_options* opt, int compute_jacobians,
                    double* residuals, double* costs, double* total_cost);

/* Per-observation Jacobians of the last b200ba_evaluate(compute_jacobians=1):
 *   j_point [n_obs*2*3], j_pose [n_obs*2*6], j_rig [n_obs*2*6] (nullable / zero if 1 camera),
 *   j_intr [n_obs*2*K], intr_index [n_obs*K] (global column of each entry), K = max over cameras
 *   of IntrinsicsJacobianSize (32 / 80 / 12). Row-major [obs][row][col]. */
B200BA_API int b200ba_get_jacobians(b200ba_handle* h, double* j_point, double* j_pose, double* j_rig,
                         double* j_intr, int32_t* intr_index, int32_t K);

/* Build H, b at the current state (hot loop 1) and download them as one dense
 * upper-triangular matrix in the reference's variable ordering
 * (joint_optimization.cc:49-59). H [n*n] row-major, b [n]; only for small problems. */
B200BA_API int b200ba_build_system(b200ba_handle* h, const b200ba_options* opt, int32_t n, double* H,
                        double* b, double* cost);
B200BA_API int32_t b200ba_degrees_of_freedom(const b200ba_handle* h, const b200ba_options* opt);

/* Stand-alone Schur-complement solve (libvis lm_optimizer.h:1246-1369) of
 *   [D B; B^T C] x = [b1; b2],  D block-diagonal with n_blocks blocks of block_size (<= 6).
 * Only the upper triangles of D blocks and C are read. Host buffers, row-major:
 *   D [n_blocks*bs*bs], B [(n_blocks*bs) * n_dense], C [n_dense*n_dense]. */
B200BA_API int b200ba_schur_solve(int device, int32_t block_size, int32_t n_blocks, int32_t n_dense,
                       const double* D, const double* B, const double* C, const double* b1,
                       const double* b2, double* x);

/* Stand-alone dense SPD solve A x = b on the in-tree kernels of the dense phase (blocked Cholesky with
 * FP64 tensor-core trailing updates, packed triangular solves): what replaces
 * `schur_M.selfadjointView<Upper>().ldlt().solve()` (libvis lm_optimizer.h:1361) inside b200ba_optimize.
 * A [n*n] symmetric, host; block_width a multiple of 128; the *_ms outputs (nullable) are device times.
 * Returns 4 if A is not positive definite. Tests / profiling. */
B200BA_API int b200ba_dense_cholesky_solve(int device, int32_t n, int32_t block_width, const double* A,
                                           const double* b, double* x, double* factor_ms, double* solve_ms);

/* CameraModel::ProjectWithInitialEstimate / Unproject for n points, on the device.
 * pixels is in/out (initial estimate / result); ok[i] = 1 on success. */
B200BA_API int b200ba_project(int device, const b200ba_camera* cam, const double* intrinsics, int64_t n,
                   const double* local_points, double* pixels, int32_t* ok);
B200BA_API int b200ba_unproject(int device, const b200ba_camera* cam, const double* intrinsics, int64_t n,
                     const double* pixels, double* directions, double* origins, int32_t* ok);

/* ---- model resampling (row f-4): CentralGenericModel::FitToPixelDirectionsImpl --------
 * (APP/models/central_generic.cc:551-568 with the cost function of :152-228 and the state of
 * :40-83). Levenberg-Marquardt over the direction grid -- 2 local DoF per control point in its
 * tangent frame -- so that the normalised B-spline un-projection at n grid points matches n unit
 * directions: LMOptimizer::Optimize(max_iteration_count, max_lm_attempts = 10, init_lambda = -1,
 * init_lambda_factor = 0.001f), quadratic loss (cost = 1/2 sum r^2 over the 3n scalar residuals),
 * dense solve. Called by FitToDenseModel / FitToPixelDirections when a model is resampled to
 * another grid resolution (APP/calibration.cc:373-522).
 *   grid         [3 * grid_width * grid_height] in/out, row-major, unit directions
 *   grid_points  [2 n] grid coordinates (PixelCornerConvToGridPoint of the sample pixels); each
 *                must have its 4x4 support inside the grid (1 <= g < size - 2)
 *   directions   [3 n] */
typedef struct b200ba_fit_report {
  double initial_cost;
  double final_cost;
  double final_lambda;
  int32_t num_iterations_performed;
  int32_t lm_attempts; /* total number of linear solves */
} b200ba_fit_report;
B200BA_API int b200ba_fit_directions(int device, int32_t grid_width, int32_t grid_height, double* grid, int64_t n,
                          const double* grid_points, const double* directions, int32_t max_iteration_count,
                          b200ba_fit_report* report);

/* ---- multi-GPU: imagesets sharded over ranks, one NCCL all-reduce per H/b build --- */
#define B200BA_NCCL_UNIQUE_ID_BYTES 128
B200BA_API int b200ba_nccl_unique_id(uint8_t id[B200BA_NCCL_UNIQUE_ID_BYTES]);
/* Every rank creates its handle from ITS shard of the observations (all ranks use the
 * same n_imagesets / n_points / cameras) and then joins the communicator. After that
 * b200ba_optimize / b200ba_optimize_host / b200ba_build_system are COLLECTIVE: every rank must
 * issue the same sequence of them with the same options. The first call after
 * b200ba_comm_init that needs the device layout (any of the above or b200ba_evaluate) also
 * all-reduces the bookkeeping that makes the ranks group the Schur blocks identically, so it
 * must be made by every rank too; later b200ba_evaluate / b200ba_get_state /
 * b200ba_get_jacobians calls are rank-local. */
B200BA_API int b200ba_comm_init(b200ba_handle* h, const uint8_t id[B200BA_NCCL_UNIQUE_ID_BYTES], int rank,
                     int n_ranks);

/* ---- instrumentation ------------------------------------------------------ */
/* Device-side timings (CUDA events on the handle's stream) of the last b200ba_optimize. */
typedef struct b200ba_timings {
  double jacobian_kernel_ms; /* sum over launches of the MAIN pass of the residual+Jacobian kernel */
  int32_t jacobian_kernel_launches;
  double accumulate_ms;      /* JtJ / Jtr accumulation kernels */
  double schur_ms;           /* D^-1, D^-1 B, B^T D^-1 B contraction */
  double factor_ms;          /* dense SPD factorisation + solves */
  double trial_cost_ms;      /* residual-only passes + comparison */
  double update_ms;          /* state retraction */
  double allreduce_ms;
  double total_ms;
  int64_t kernel_launches;   /* kernels of this library launched */
  double straggler_ms;       /* straggler passes of the residual/Jacobian kernel (observations whose
                                projection needs more than the main pass's evaluation budget) */
  double solve_ms;           /* triangular solves with the dense factor (part of factor_ms) */
  double contraction_flops;  /* FP64 flops of the Schur contraction S = C - W^T W actually issued (structured:
                                sum over groups of m_g^2 k_g; dense: n_d^2 * k), summed over LM attempts */
  double factor_flops;       /* n_d^3 / 3 per factorisation, summed over LM attempts */
  int32_t lm_attempts;       /* linear solves (LM attempts) of the last b200ba_optimize */
  int32_t build_count;       /* H / b builds (outer iterations) of the last b200ba_optimize */
} b200ba_timings;
B200BA_API int b200ba_get_timings(const b200ba_handle* h, b200ba_timings* t);

B200BA_API const char* b200ba_version(void);

/* ---- diagnostics (not part of the reference interface) ---------------------------------
 * Evaluation budget of the main residual/Jacobian pass (default 16 spline evaluations per
 * observation; what exceeds it is redone by the straggler pass). 1 sends every observation of
 * a generic camera through the straggler pass (tests). Process-wide. */
B200BA_API void b200ba_debug_set_eval_budget(int budget);
/* Spline evaluations the projection LM spent per observation in the last pass that wrote
 * Jacobians; counts [n_obs], caller's observation order. */
B200BA_API int b200ba_debug_eval_counts(b200ba_handle* h, uint16_t* counts);

#ifdef __cplusplus
}
#endif
#endif /* B200BA_H_ */
