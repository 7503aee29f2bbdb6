// ba_common.h -- structures shared by the host logic (ba_host.cu) and the kernels
// (ba_kernels.cu) of libb200ba.so. Not part of the public ABI (that is include/b200ba.h).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/b200ba.h"

namespace b200ba {

constexpr int kMaxCameras = 8;
constexpr int kMaxK = 80;  // IntrinsicsJacobianSize of the non-central model

// Device view of one camera (CameraModel base members + derived constants).
struct CamDev {
  int model_type;
  int width, height;
  int min_x, min_y, max_x, max_y;
  int gw, gh;
  // pixel -> grid map  g = 1 + gmul * (x - min)     (central_grid.h:150-154)
  double gmul_x, gmul_y;
  // PixelScaleToGridScaleX/Y: computed in FLOAT in the reference (central_grid.h:156-161)
  double sx, sy;
  // CenterOfCalibratedArea (camera_model.h:154-157)
  double center_x, center_y;
  int K;              // IntrinsicsJacobianSize: 32 / 80 / 12
  int dof_per_point;  // 2 / 5 / 0
  int64_t intr_off;   // offset of this camera's flat intrinsics in the state's intrinsics buffer
  int64_t tan_off;    // offset of its tangent frames (6 doubles per control point)
  int upd_off;        // offset of its update parameters relative to first_intrinsics
  int upd_count;      // update_parameter_count()
};

// Variable layout of JointOptimizationState (joint_optimization.cc:49-59,97-170) and of the
// block-diagonal / dense split of H (lm_optimizer.h:657-685):
//   eliminate_points:  [points 3P | rig_tr_global 6N | camera_tr_rig 6C (iff C > 1) | intrinsics]
//                      block part = points (3x3 blocks)
//   otherwise:         [rig_tr_global 6N | camera_tr_rig 6C (iff C > 1) | points 3P | intrinsics]
//                      block part = imageset poses (6x6 blocks)
// g_* are offsets in that global ordering; a dense index is (global - nbd).
struct Layout {
  int n_points, n_imagesets, n_cameras;
  int rig_in_state;     // n_cameras > 1
  int localize_only;
  int eliminate_points;
  int dof;              // all unknowns
  int bs;               // Schur block size: 3 (points) or 6 (poses)
  int nblocks;          // number of blocks: P or N
  int dsz;              // packed upper-triangle size of a block: bs (bs + 1) / 2
  int nbd;              // bs * nblocks (block-diagonal part)
  int nd;               // dense part
  int g_point, g_pose, g_rig, g_intr;  // global offsets of the variable groups
  int n_jcols;          // Jacobian columns stored per observation: 3 + 6 + (rig ? 6 : 0) + Kmax
  int Kmax;             // max IntrinsicsJacobianSize over cameras (0 when localize_only)
  int jc_point, jc_pose, jc_rig, jc_intr;  // first storage column of each part
};

struct ProblemDev {
  int64_t n_obs;
  const uint32_t* obs_imageset;
  const uint32_t* obs_camera;
  const uint32_t* obs_point;
  const float2* obs_xy;
  CamDev cams[kMaxCameras];
};

// One copy of the optimised state on the device.
struct StateDev {
  double* points;         // [3 * n_points]
  double* rig_tr_global;  // [7 * n_imagesets]
  double* camera_tr_rig;  // [7 * n_cameras]
  double* intrinsics;     // all cameras, CamDev::intr_off
  // derived, recomputed by prepare_state():
  double* image_tr_global;  // [n_imagesets * n_cameras][12]: R row-major (9), t (3)
  double* tangents;         // per control point t1 (3), t2 (3); CamDev::tan_off
};

// Per-observation outputs of the residual / Jacobian kernel.
struct ObsOut {
  double* residual;    // [2 * n_obs] SoA: rx[n_obs], ry[n_obs]
  double* cost;        // [n_obs], -1 = invalid residual
  double* jac;         // [2 * n_jcols][n_obs] SoA: row-x of column c at (2c) * n_obs, row-y at (2c+1) * n_obs
                       //   (NULL in compact mode, except while b200ba_get_jacobians expands it)
  // Compact mode (all cameras central-generic): instead of the 2 x (9 + 6 + 32) Jacobian entries the
  // kernel stores the 14 numbers they are all built from -- P = d pixel / d local_point (2x3), the two
  // rows of Mn = -(A^T A)^-1 A^T / |sum w G| (2x3) and the fractions (fu, fv) of the B-spline support;
  // consumers rebuild a column as  point / pose / rig: chain rule of P with the (L2-resident) poses,
  // intrinsics: w_k(fu, fv) * Mn [t1 t2]_k with the tangent frames. 116 B instead of 656 B per observation.
  double* cjac;        // [14][n_obs] SoA: P00 P01 P02 P10 P11 P12 | Mn00 Mn01 Mn02 Mn10 Mn11 Mn12 | fu fv
  int compact;
  int32_t* cell;       // [n_obs] top-left control point x0 + y0 * gw of the 4x4 support (generic models)
  uint8_t* has_jac;    // [n_obs]
  uint16_t* evals;     // [n_obs] spline evaluations spent by the projection LM (diagnostics; may be NULL)
};

// The normal equations on the device (all FP64). Everything that takes part in the
// cross-GPU reduction lives in ONE allocation so that a single all-reduce covers it.
struct SystemDev {
  double* base;     // the single allocation
  int64_t total;    // doubles
  double* Dblk;     // [nblocks][dsz]  packed upper triangle of each bs x bs block, row-major
  double* bp;       // [nbd]
  double* B;        // [nbd][nd]   row-major
  double* C;        // [nd][nd] row-major, row <= col valid (== column-major lower)
  double* bd;       // [nd]
  double* scalars;  // [8]: cost, n_valid, ...
};

// Up to four ranges [lo, hi) of global unknown indices held fixed (debug_fix_* of OptimizeJointly).
struct FixedRanges {
  int n;
  int lo[4], hi[4];
  __host__ __device__ bool has(int g) const {
    for (int i = 0; i < n; ++i)
      if (g >= lo[i] && g < hi[i]) return true;
    return false;
  }
};

}  // namespace b200ba
