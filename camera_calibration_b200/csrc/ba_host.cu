// ba_host.cu -- host logic of libb200ba.so: the C ABI of include/b200ba.h, device memory
// management, the Levenberg-Marquardt control loop and the dense-algebra plumbing.
//
// What runs where
//   device, own kernels (ba_kernels.cu): residuals, Jacobians, J^T J accumulation, 3x3 block
//     factorisation, L^-1 B, state retraction, cost comparison.
//   device, libraries: the symmetric rank-k update S = C - W^T W (cublasDsyrk, a plain dense
//     FP64 contraction -- the reference uses cublasXtDgemm for it, LV/lm_optimizer.h:1371-1430)
//     and the dense SPD factorisation (cusolverDnDpotrf/potrs; scaffolding).
//   host: the scalar LM decisions (accept / reject, lambda), exactly LV/lm_optimizer.h:628-991
//     as driven by APP/bundle_adjustment/joint_optimization.cc:905-940.
// There is NO CPU fallback: every entry point fails with an error if CUDA is unavailable.

#include <cublas_v2.h>
#include <cusolverDn.h>
#include <dlfcn.h>

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "ba_kernels.h"

using namespace b200ba;

namespace {

std::string g_create_error;

struct NcclUniqueId {  // layout of ncclUniqueId (nccl.h): 128 opaque bytes, passed BY VALUE
  char internal[128];
};
struct NcclApi {
  void* lib = nullptr;
  int (*GetUniqueId)(void*) = nullptr;
  int (*CommInitRank)(void**, int, NcclUniqueId, int) = nullptr;
  int (*AllReduce)(const void*, void*, size_t, int, int, void*, cudaStream_t) = nullptr;
  int (*Broadcast)(const void*, void*, size_t, int, int, void*, cudaStream_t) = nullptr;
  int (*ReduceScatter)(const void*, void*, size_t, int, int, void*, cudaStream_t) = nullptr;
  int (*CommDestroy)(void*) = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
};
// ncclDataType_t / ncclRedOp_t values (nccl.h): ncclFloat64 = 8, ncclSum = 0
constexpr int kNcclDouble = 8;
constexpr int kNcclInt32 = 2;
constexpr int kNcclSum = 0;
constexpr int kNcclMax = 2;

NcclApi g_nccl;
bool load_nccl(std::string* err) {
  if (g_nccl.lib) return true;
  // Reuse the copy already mapped into the process (torch ships its own libnccl.so.2).
  const char* names[] = {"libnccl.so.2", "libnccl.so"};
  void* lib = nullptr;
  for (const char* n : names) {
    lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
    if (lib) break;
  }
  if (!lib) {
    *err = std::string("cannot load libnccl: ") + dlerror();
    return false;
  }
  g_nccl.lib = lib;
  *reinterpret_cast<void**>(&g_nccl.GetUniqueId) = dlsym(lib, "ncclGetUniqueId");
  *reinterpret_cast<void**>(&g_nccl.CommInitRank) = dlsym(lib, "ncclCommInitRank");
  *reinterpret_cast<void**>(&g_nccl.AllReduce) = dlsym(lib, "ncclAllReduce");
  *reinterpret_cast<void**>(&g_nccl.Broadcast) = dlsym(lib, "ncclBroadcast");
  *reinterpret_cast<void**>(&g_nccl.ReduceScatter) = dlsym(lib, "ncclReduceScatter");
  *reinterpret_cast<void**>(&g_nccl.CommDestroy) = dlsym(lib, "ncclCommDestroy");
  *reinterpret_cast<void**>(&g_nccl.GetErrorString) = dlsym(lib, "ncclGetErrorString");
  if (!g_nccl.GetUniqueId || !g_nccl.CommInitRank || !g_nccl.AllReduce || !g_nccl.Broadcast || !g_nccl.ReduceScatter ||
      !g_nccl.CommDestroy) {
    *err = "libnccl lacks a required symbol";
    g_nccl.lib = nullptr;
    return false;
  }
  return true;
}

constexpr int kMaxCholPanels = 1024;

enum Phase { PH_JAC = 0, PH_ACC, PH_SCHUR, PH_FACTOR, PH_TRIAL, PH_UPDATE, PH_ALLREDUCE, PH_STRAGGLER, PH_SOLVE, PH_COUNT };

}  // namespace

struct b200ba_handle {
  int device = 0;
  cudaStream_t stream = nullptr;
  cudaStream_t side_stream = nullptr;  // straggler pass of the Jacobian kernel, overlapped with accumulation
  cudaEvent_t straggler_done = nullptr;
  cublasHandle_t cublas = nullptr;
  cusolverDnHandle_t cusolver = nullptr;
  std::string error;

  // problem
  std::vector<b200ba_camera> cams_host;
  ProblemDev pb{};
  int n_cameras = 0, n_imagesets = 0, n_points = 0;
  int64_t n_obs = 0;
  int uniform_model = -1;
  int64_t n_control_total = 0, n_param_total = 0, intr_total = 0, tan_total = 0;
  uint32_t *d_obs_imageset = nullptr, *d_obs_camera = nullptr, *d_obs_point = nullptr;
  float2* d_obs_xy = nullptr;

  // state (two copies: current / trial)
  StateDev st[2]{};
  int cur = 0;
  double2* d_last_projection = nullptr;
  bool have_state = false;
  // device-side snapshot of (state, last_projection): b200ba_snapshot_state / b200ba_restore_state
  StateDev snap{};
  double2* d_snap_lp = nullptr;
  bool have_snapshot = false;

  // layout-dependent buffers
  Layout L{};
  bool have_layout = false;
  ObsOut out{};        // base evaluation (with Jacobians)
  ObsOut out_trial{};  // residual-only evaluation of the trial state
  SystemDev sys{};
  double *d_W = nullptr, *d_S = nullptr, *d_Linv = nullptr, *d_v = nullptr, *d_y = nullptr, *d_x = nullptr;
  double trace_H = 0;
  int64_t reduce_count = 0;  // doubles of sys.base covered by the per-build all-reduce
  double* d_potrf_work = nullptr;
  int potrf_lwork = 0;
  int chol_mode = -1;                     // B200BA_DIST_CHOL=0|1 overrides: 1 = column-block Cholesky (see factor_dense)
  int chol_nb = 512;                      // its panel width (B200BA_CHOL_NB)
  int *d_info = nullptr, *d_fail = nullptr;
  // Static cell-major processing order: device position -> index in the caller's (reference)
  // observation order. Computed once at create time from the cell of the measured pixel.
  std::vector<uint32_t> perm;
  uint32_t* d_perm = nullptr;
  double2* d_lp_stage = nullptr;  // last_projection in the caller's order (H2D / D2H staging)
  uint32_t* d_straggler_list = nullptr;  // observations deferred by the main pass of the Jacobian kernel
  int* d_straggler_count = nullptr;
  double *d_partial = nullptr, *d_scal = nullptr;
  double* d_rot = nullptr;  // [9 * n_cameras] rotations of ChooseNiceCameraOrientation
  double* h_scal = nullptr;  // pinned [16]
  int* h_flags = nullptr;    // pinned [2]

  // host copies of the observations (caller's order) for layout-time block grouping
  std::vector<uint32_t> h_obs_imageset, h_obs_camera, h_obs_point;
  std::vector<float> h_obs_xy;
  // structured Schur contraction (groups of Schur blocks with compacted column support)
  int group_blocks_n = 0;                 // blocks per group
  int n_groups = 0;
  std::vector<int> group_start;           // [n_groups + 1] into group_blocks
  std::vector<int> group_count;           // m_g of the current build
  int* d_group_of_block = nullptr;
  int* d_group_blocks = nullptr;
  uint8_t* d_flags = nullptr;
  int* d_cols = nullptr;
  int* d_count = nullptr;
  int* h_count = nullptr;                 // pinned
  double* d_Wc = nullptr;                 // 2 panels (double-buffered)
  double* d_P = nullptr;                  // 2 result buffers (double-buffered)
  size_t wc_stride = 0, p_stride = 0;
  cudaEvent_t ev_syrk[2] = {nullptr, nullptr}, ev_scatter[2] = {nullptr, nullptr}, ev_s_ready = nullptr;
  double* d_u = nullptr;
  bool use_grouped = false;
  std::vector<double> grp_sums;           // [sx | sy | count] per Schur block (see build_groups)
  int force_grouped = -1;                 // B200BA_GROUPED=0|1 overrides the cost model
  bool compact_j = true;                  // B200BA_COMPACT_J=0: expanded Jacobian buffer also for central-generic cameras

  // multi-GPU
  void* comm = nullptr;
  int rank = 0, n_ranks = 1;

  // dense phase on the in-tree DMMA kernels (ba_dense.cu); B200BA_DENSE=lib selects the cuBLAS / cuSOLVER path
  bool own_dense = true;
  DenseCtx dn;
  int dense_planned_ranks = 0, dense_planned_n = -1;
  int dense_nb = 0;                     // column-block width of the factorisation (B200BA_DENSE_NB, multiple of 128); 0 = by rank count
  cudaStream_t panel_stream = nullptr;  // panel factorisations + broadcasts (look-ahead)
  cudaStream_t aux_stream = nullptr;    // second look-ahead update of the single-GPU factorisation
  int* d_ident_cols = nullptr;          // 0 .. nd - 1 (scatter epilogue of the dense contraction with several ranks)
  double* d_gemv_partial = nullptr;     // slab sums of the B^T u product

  // timings
  b200ba_timings timings{};
  struct Pending {
    int phase;
    cudaEvent_t a, b;
    bool own_a = true;  // false: `a` is another entry's `b` and is released there
  };
  std::vector<Pending> pending;
  std::vector<cudaEvent_t> event_pool;
};

namespace {

#define CUDA_TRY(h, expr)                                                                      \
  do {                                                                                         \
    cudaError_t _e = (expr);                                                                   \
    if (_e != cudaSuccess) {                                                                   \
      (h)->error = std::string(#expr) + ": " + cudaGetErrorString(_e);                         \
      return 1;                                                                                \
    }                                                                                          \
  } while (0)
#define CUBLAS_TRY(h, expr)                                                                    \
  do {                                                                                         \
    cublasStatus_t _s = (expr);                                                                \
    if (_s != CUBLAS_STATUS_SUCCESS) {                                                         \
      (h)->error = std::string(#expr) + ": cuBLAS status " + std::to_string(static_cast<int>(_s)); \
      return 1;                                                                                \
    }                                                                                          \
  } while (0)
#define CUSOLVER_TRY(h, expr)                                                                  \
  do {                                                                                         \
    cusolverStatus_t _s = (expr);                                                              \
    if (_s != CUSOLVER_STATUS_SUCCESS) {                                                       \
      (h)->error = std::string(#expr) + ": cuSOLVER status " + std::to_string(static_cast<int>(_s)); \
      return 1;                                                                                \
    }                                                                                          \
  } while (0)

template <class T>
int dev_alloc(b200ba_handle* h, T** p, size_t count) {
  if (*p) {
    cudaFree(*p);
    *p = nullptr;
  }
  if (count == 0) count = 1;
  CUDA_TRY(h, cudaMalloc(reinterpret_cast<void**>(p), count * sizeof(T)));
  return 0;
}

int64_t intrinsics_size(const b200ba_camera& c) {
  const int64_t G = static_cast<int64_t>(c.grid_width) * c.grid_height;
  switch (c.model_type) {
    case B200BA_MODEL_CENTRAL_GENERIC: return 3 * G;
    case B200BA_MODEL_NONCENTRAL_GENERIC: return 6 * G;
    default: return 12;
  }
}
int update_parameter_count(const b200ba_camera& c) {
  const int G = c.grid_width * c.grid_height;
  switch (c.model_type) {
    case B200BA_MODEL_CENTRAL_GENERIC: return 2 * G;
    case B200BA_MODEL_NONCENTRAL_GENERIC: return 5 * G;
    default: return 12;
  }
}
int jacobian_size(const b200ba_camera& c) {
  switch (c.model_type) {
    case B200BA_MODEL_CENTRAL_GENERIC: return 32;
    case B200BA_MODEL_NONCENTRAL_GENERIC: return 80;
    default: return 12;
  }
}

void fill_camdev(const b200ba_camera& c, CamDev* d) {
  d->model_type = c.model_type;
  d->width = c.width;
  d->height = c.height;
  d->min_x = c.calibration_min_x;
  d->min_y = c.calibration_min_y;
  d->max_x = c.calibration_max_x;
  d->max_y = c.calibration_max_y;
  d->gw = c.grid_width;
  d->gh = c.grid_height;
  const int aw = c.calibration_max_x + 1 - c.calibration_min_x;
  const int ah = c.calibration_max_y + 1 - c.calibration_min_y;
  d->gmul_x = (c.grid_width > 0) ? static_cast<double>(c.grid_width - 3) / aw : 0.0;
  d->gmul_y = (c.grid_height > 0) ? static_cast<double>(c.grid_height - 3) / ah : 0.0;
  // PixelScaleToGridScaleX/Y divide in float (APP/models/central_grid.h:156-161)
  d->sx = (c.grid_width > 0) ? static_cast<double>((c.grid_width - 3.f) / aw) : 0.0;
  d->sy = (c.grid_height > 0) ? static_cast<double>((c.grid_height - 3.f) / ah) : 0.0;
  d->center_x = 0.5f * static_cast<float>(c.calibration_min_x + c.calibration_max_x + 1);
  d->center_y = 0.5f * static_cast<float>(c.calibration_min_y + c.calibration_max_y + 1);
  d->K = jacobian_size(c);
  d->dof_per_point = c.model_type == B200BA_MODEL_CENTRAL_GENERIC ? 2 : (c.model_type == B200BA_MODEL_NONCENTRAL_GENERIC ? 5 : 0);
  d->upd_count = update_parameter_count(c);
}

int64_t align32(int64_t n) { return (n + 31) / 32 * 32; }

// ---- timing ----------------------------------------------------------------------------
cudaEvent_t get_event(b200ba_handle* h) {
  if (!h->event_pool.empty()) {
    cudaEvent_t e = h->event_pool.back();
    h->event_pool.pop_back();
    return e;
  }
  cudaEvent_t e;
  cudaEventCreate(&e);
  return e;
}
struct ScopedPhase {
  b200ba_handle* h;
  int phase;
  cudaEvent_t a;
  ScopedPhase(b200ba_handle* h_, int p) : h(h_), phase(p) {
    a = get_event(h);
    cudaEventRecord(a, h->stream);
  }
  ~ScopedPhase() {
    cudaEvent_t b = get_event(h);
    cudaEventRecord(b, h->stream);
    h->pending.push_back({phase, a, b, true});
  }
};
void resolve_timings(b200ba_handle* h) {
  for (auto& p : h->pending) {
    float ms = 0;
    if (cudaEventElapsedTime(&ms, p.a, p.b) == cudaSuccess) {
      switch (p.phase) {
        case PH_JAC: h->timings.jacobian_kernel_ms += ms; h->timings.jacobian_kernel_launches++; break;
        case PH_ACC: h->timings.accumulate_ms += ms; break;
        case PH_SCHUR: h->timings.schur_ms += ms; break;
        case PH_FACTOR: h->timings.factor_ms += ms; break;
        case PH_TRIAL: h->timings.trial_cost_ms += ms; break;
        case PH_UPDATE: h->timings.update_ms += ms; break;
        case PH_ALLREDUCE: h->timings.allreduce_ms += ms; break;
        case PH_STRAGGLER: h->timings.straggler_ms += ms; break;
        case PH_SOLVE: h->timings.solve_ms += ms; h->timings.factor_ms += ms; break;
      }
    }
    if (p.own_a) h->event_pool.push_back(p.a);
    h->event_pool.push_back(p.b);
  }
  h->pending.clear();
}
// All-reduces the centroid sums over the ranks (no-op without a communicator).
int reduce_group_sums(b200ba_handle* h) {
  if (h->n_ranks <= 1 || !h->comm || h->grp_sums.empty()) return 0;
  const size_t n = h->grp_sums.size();
  double* d = nullptr;
  CUDA_TRY(h, cudaMalloc(reinterpret_cast<void**>(&d), n * sizeof(double)));
  cudaMemcpy(d, h->grp_sums.data(), n * sizeof(double), cudaMemcpyHostToDevice);
  const int rc = g_nccl.AllReduce(d, d, n, kNcclDouble, kNcclSum, h->comm, h->stream);
  const cudaError_t ce = cudaStreamSynchronize(h->stream);
  if (rc == 0 && ce == cudaSuccess) cudaMemcpy(h->grp_sums.data(), d, n * sizeof(double), cudaMemcpyDeviceToHost);
  cudaFree(d);
  if (rc != 0 || ce != cudaSuccess) {
    h->error = "all-reduce of the group centroids failed";
    return 1;
  }
  return 0;
}

// Groups of Schur blocks for the structured contraction, from the centroid sums in h->grp_sums
// ([sx | sy | count] per block). Must produce the same grouping on every rank: the ranks split
// the GROUPS among themselves (solve_system), so a rank-dependent order would drop / repeat blocks.
int build_groups(b200ba_handle* h, bool allocate) {
  const Layout& L = h->L;
  const int nb = L.nblocks;
  int gb = (L.bs == 3) ? 96 : 48;  // ~288 rows per group: measured optimum on B200 (config 2: 64 -> 12.0 ms, 96 -> 11.6 ms, 160 -> 11.7 ms)
  if (const char* e = getenv("B200BA_GROUP_BLOCKS")) gb = std::max(1, atoi(e));
  h->group_blocks_n = gb;
  std::vector<int> order(nb);
  for (int i = 0; i < nb; ++i) order[i] = i;
  if (L.eliminate_points && nb > 0) {
    // order the pattern points along a Z-curve of the centroid of their measured pixels: points
    // that are neighbours in the image share most of their control-point support
    const double* sx = h->grp_sums.data();
    const double* sy = sx + nb;
    const double* cnt = sy + nb;
    const double qx = std::max(1.0, h->cams_host[0].width / 64.0), qy = std::max(1.0, h->cams_host[0].height / 64.0);
    std::vector<uint32_t> zkey(nb, 0);
    for (int p = 0; p < nb; ++p) {
      if (cnt[p] <= 0) continue;
      uint32_t ux = static_cast<uint32_t>(std::min(1023.0, std::max(0.0, sx[p] / cnt[p] / qx)));
      uint32_t uy = static_cast<uint32_t>(std::min(1023.0, std::max(0.0, sy[p] / cnt[p] / qy)));
      uint32_t z = 0;
      for (int b = 0; b < 10; ++b) z |= ((ux >> b) & 1u) << (2 * b) | ((uy >> b) & 1u) << (2 * b + 1);
      zkey[p] = z;
    }
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return zkey[a] < zkey[b]; });
  }
  h->n_groups = (nb + gb - 1) / gb;
  h->group_start.assign(h->n_groups + 1, 0);
  std::vector<int> gob(std::max(1, nb), 0);
  for (int g = 0; g < h->n_groups; ++g) {
    h->group_start[g] = g * gb;
    for (int i = g * gb; i < std::min(nb, (g + 1) * gb); ++i) gob[order[i]] = g;
  }
  h->group_start[h->n_groups] = nb;
  h->group_count.assign(std::max(1, h->n_groups), 0);
  if (allocate) {
    if (dev_alloc(h, &h->d_group_of_block, std::max(1, nb))) return 1;
    if (dev_alloc(h, &h->d_group_blocks, std::max(1, nb))) return 1;
    if (dev_alloc(h, &h->d_flags, static_cast<size_t>(std::max(1, h->n_groups)) * std::max(1, L.nd))) return 1;
    if (dev_alloc(h, &h->d_cols, static_cast<size_t>(std::max(1, h->n_groups)) * std::max(1, L.nd))) return 1;
    if (dev_alloc(h, &h->d_count, std::max(1, h->n_groups))) return 1;
    if (h->h_count) cudaFreeHost(h->h_count);
    CUDA_TRY(h, cudaMallocHost(reinterpret_cast<void**>(&h->h_count), std::max(1, h->n_groups) * sizeof(int)));
    h->wc_stride = static_cast<size_t>(gb) * L.bs * (std::max(1, L.nd) + 1);
    h->p_stride = static_cast<size_t>(std::max(1, L.nd)) * std::max(1, L.nd);
    if (dev_alloc(h, &h->d_Wc, 2 * h->wc_stride)) return 1;
    if (!h->own_dense && dev_alloc(h, &h->d_P, 2 * h->p_stride)) return 1;
    for (int i = 0; i < 2; ++i) {
      if (!h->ev_syrk[i]) CUDA_TRY(h, cudaEventCreateWithFlags(&h->ev_syrk[i], cudaEventDisableTiming));
      if (!h->ev_scatter[i]) CUDA_TRY(h, cudaEventCreateWithFlags(&h->ev_scatter[i], cudaEventDisableTiming));
    }
    if (!h->ev_s_ready) CUDA_TRY(h, cudaEventCreateWithFlags(&h->ev_s_ready, cudaEventDisableTiming));
    if (dev_alloc(h, &h->d_u, std::max(1, L.nbd))) return 1;
  }
  CUDA_TRY(h, cudaMemcpy(h->d_group_of_block, gob.data(), std::max(1, nb) * sizeof(int), cudaMemcpyHostToDevice));
  if (nb > 0) CUDA_TRY(h, cudaMemcpy(h->d_group_blocks, order.data(), nb * sizeof(int), cudaMemcpyHostToDevice));
  return 0;
}

// ---- dense phase buffers (own kernels) ------------------------------------------------------
int nccl_bcast_cb(double* buf, size_t count, int root, cudaStream_t s, void* user) {
  b200ba_handle* h = static_cast<b200ba_handle*>(user);
  const int rc = g_nccl.Broadcast(buf, buf, count, kNcclDouble, root, h->comm, s);
  if (rc != 0) {
    h->error = std::string("ncclBroadcast: ") + (g_nccl.GetErrorString ? g_nccl.GetErrorString(rc) : "error");
    return 1;
  }
  return 0;
}

// (Re)plans the storage of S / the packed factor for the current (n_d, rank count).
int plan_dense(b200ba_handle* h) {
  if (!h->own_dense || !h->have_layout) return 0;
  const int nd = h->L.nd;
  if (h->dense_planned_n == nd && h->dense_planned_ranks == h->n_ranks) return 0;
  DenseCtx& d = h->dn;
  // one GPU: 512-wide block columns (fewer, longer panels: 28.6 vs 29.5 ms at n_d = 13 080, profiles/r02zc_dense_timing.log);
  // several ranks: 256, so that the block-cyclic distribution has enough blocks per rank and the broadcasts stay short
  const int nb = h->dense_nb > 0 ? h->dense_nb : (h->n_ranks == 1 ? 512 : 256);
  dense_plan(&d, nd, nb, h->rank, h->n_ranks);
  if (dev_alloc(h, &h->d_S, static_cast<size_t>(std::max<int64_t>(1, d.chunk * h->n_ranks)))) return 1;
  d.S = h->d_S;
  if (dev_alloc(h, &d.Lpack, static_cast<size_t>(std::max<int64_t>(1, d.panel_off[d.nblk])))) return 1;
  if (dev_alloc(h, &d.tmp, std::max(1, nd))) return 1;
  if (dev_alloc(h, &d.d_panel_off, d.panel_off.size())) return 1;
  if (dev_alloc(h, &d.d_panel_h, d.panel_h.size())) return 1;
  CUDA_TRY(h, cudaMemcpy(d.d_panel_off, d.panel_off.data(), d.panel_off.size() * sizeof(int64_t), cudaMemcpyHostToDevice));
  CUDA_TRY(h, cudaMemcpy(d.d_panel_h, d.panel_h.data(), d.panel_h.size() * sizeof(int), cudaMemcpyHostToDevice));
  {
    std::vector<int> ident(std::max(1, nd));
    for (int i = 0; i < nd; ++i) ident[i] = i;
    if (dev_alloc(h, &h->d_ident_cols, ident.size())) return 1;
    CUDA_TRY(h, cudaMemcpy(h->d_ident_cols, ident.data(), ident.size() * sizeof(int), cudaMemcpyHostToDevice));
  }
  if (dev_alloc(h, &h->d_gemv_partial, static_cast<size_t>(gemv_t_partial_size(h->L.nbd, std::max(1, nd))))) return 1;
  // the un-owned chunks of S are never written by the contraction of a single rank but are read by nobody either;
  // zero once so that partial sums start clean
  CUDA_TRY(h, cudaMemset(h->d_S, 0, static_cast<size_t>(std::max<int64_t>(1, d.chunk * h->n_ranks)) * sizeof(double)));
  d.info = h->d_info;
  d.s_main = h->stream;
  d.s_panel = h->panel_stream;
  d.s_aux = h->aux_stream;
  for (int i = 0; i < 2; ++i) {
    if (!d.ev_ready[i]) CUDA_TRY(h, cudaEventCreateWithFlags(&d.ev_ready[i], cudaEventDisableTiming));
    if (!d.ev_main[i]) CUDA_TRY(h, cudaEventCreateWithFlags(&d.ev_main[i], cudaEventDisableTiming));
    if (!d.ev_half2[i]) CUDA_TRY(h, cudaEventCreateWithFlags(&d.ev_half2[i], cudaEventDisableTiming));
  }
  if (!d.ev_misc) CUDA_TRY(h, cudaEventCreateWithFlags(&d.ev_misc, cudaEventDisableTiming));
  d.bcast = nccl_bcast_cb;
  d.user = h;
  h->dense_planned_n = nd;
  h->dense_planned_ranks = h->n_ranks;
  return 0;
}

// ---- layout / buffers ----------------------------------------------------------------------
int make_layout(b200ba_handle* h, const b200ba_options* opt) {
  if (opt->regularization_weight != 0) {
    // the reference logs an error and carries on without the term (joint_optimization.cc:299-305)
    static bool warned = false;
    if (!warned) fprintf(stderr, "[b200ba] regularization_weight is ignored (the term is disabled in the reference, joint_optimization.cc:299-305)\n");
    warned = true;
  }
  Layout L{};
  memset(&L, 0, sizeof(L));
  L.n_points = h->n_points;
  L.n_imagesets = h->n_imagesets;
  L.n_cameras = h->n_cameras;
  L.rig_in_state = h->n_cameras > 1;
  L.localize_only = opt->localize_only ? 1 : 0;
  L.eliminate_points = opt->eliminate_points ? 1 : 0;
  int n_intr = 0, kmax = 0;
  for (int c = 0; c < h->n_cameras; ++c) {
    n_intr += update_parameter_count(h->cams_host[c]);
    kmax = std::max(kmax, jacobian_size(h->cams_host[c]));
  }
  if (L.localize_only) {
    n_intr = 0;
    kmax = 0;
  }
  const int rig_dof = L.rig_in_state ? 6 * h->n_cameras : 0;
  // JointOptimizationState offsets (joint_optimization.cc:142-170)
  if (L.eliminate_points) {
    L.bs = 3;
    L.nblocks = h->n_points;
    L.g_point = 0;
    L.g_pose = 3 * h->n_points;
    L.g_rig = L.g_pose + 6 * h->n_imagesets;
    L.g_intr = L.g_rig + rig_dof;
  } else {
    L.bs = 6;
    L.nblocks = h->n_imagesets;
    L.g_pose = 0;
    L.g_rig = 6 * h->n_imagesets;
    L.g_point = L.g_rig + rig_dof;
    L.g_intr = L.g_point + 3 * h->n_points;
  }
  L.dsz = L.bs * (L.bs + 1) / 2;
  L.nbd = L.bs * L.nblocks;
  L.dof = 3 * h->n_points + 6 * h->n_imagesets + rig_dof + n_intr;
  L.nd = L.dof - L.nbd;
  L.Kmax = kmax;
  L.jc_point = 0;
  L.jc_pose = 3;
  L.jc_rig = 9;
  L.jc_intr = 9 + (L.rig_in_state ? 6 : 0);
  L.n_jcols = L.jc_intr + kmax;
  const bool same = h->have_layout && memcmp(&L, &h->L, sizeof(Layout)) == 0;
  if (same) return plan_dense(h);
  h->L = L;
  h->have_layout = true;
  h->dense_planned_n = -1;

  const int64_t n = h->n_obs;
  if (dev_alloc(h, &h->out.residual, 2 * n)) return 1;
  if (dev_alloc(h, &h->out.cost, n)) return 1;
  // compact Jacobian records when every camera is central-generic (B200BA_COMPACT_J=0 keeps the expanded buffer)
  h->out.compact = (h->uniform_model == B200BA_MODEL_CENTRAL_GENERIC && h->compact_j) ? 1 : 0;
  if (h->out.compact) {
    if (h->out.jac) cudaFree(h->out.jac);
    h->out.jac = nullptr;
    if (dev_alloc(h, &h->out.cjac, 14 * static_cast<size_t>(n))) return 1;
  } else {
    if (dev_alloc(h, &h->out.jac, 2 * static_cast<size_t>(L.n_jcols) * n)) return 1;
  }
  if (dev_alloc(h, &h->out.cell, n)) return 1;
  if (dev_alloc(h, &h->out.has_jac, n)) return 1;
  if (dev_alloc(h, &h->out.evals, n)) return 1;
  h->out_trial.evals = nullptr;
  if (dev_alloc(h, &h->out_trial.residual, 2 * n)) return 1;
  if (dev_alloc(h, &h->out_trial.cost, n)) return 1;
  h->out_trial.jac = nullptr;
  h->out_trial.cjac = nullptr;
  h->out_trial.compact = 0;
  h->out_trial.cell = nullptr;
  h->out_trial.has_jac = nullptr;

  // the normal equations: one allocation (one all-reduce)
  SystemDev& s = h->sys;
  // [D | b_p | B | b_d | scalars] are all-reduced after every build; C comes last and is NOT:
  // with several ranks each rank folds its partial C into its partial Schur complement, and the
  // all-reduce of S makes it global (see solve_system).
  const int64_t oD = 0;
  const int64_t obp = oD + align32(static_cast<int64_t>(L.dsz) * L.nblocks);
  const int64_t oB = obp + align32(L.nbd);
  const int64_t obd = oB + align32(static_cast<int64_t>(L.nbd) * L.nd);
  const int64_t osc = obd + align32(L.nd);
  const int64_t oC = osc + 32;
  s.total = oC + align32(static_cast<int64_t>(L.nd) * L.nd);
  h->reduce_count = oC;
  if (dev_alloc(h, &s.base, s.total)) return 1;
  s.Dblk = s.base + oD;
  s.bp = s.base + obp;
  s.B = s.base + oB;
  s.C = s.base + oC;
  s.bd = s.base + obd;
  s.scalars = s.base + osc;
  if (dev_alloc(h, &h->d_W, static_cast<size_t>(L.nbd) * L.nd)) return 1;
  if (!h->own_dense) {
    if (dev_alloc(h, &h->d_S, static_cast<size_t>(L.nd) * L.nd + L.nd)) return 1;  // + partial rhs tail
  } else if (plan_dense(h)) {
    return 1;
  }
  if (dev_alloc(h, &h->d_Linv, static_cast<size_t>(L.dsz) * L.nblocks)) return 1;
  if (dev_alloc(h, &h->d_v, L.nbd)) return 1;
  if (dev_alloc(h, &h->d_y, L.nbd)) return 1;
  if (dev_alloc(h, &h->d_x, L.dof)) return 1;
  int lwork = 0;
  CUSOLVER_TRY(h, cusolverDnDpotrf_bufferSize(h->cusolver, CUBLAS_FILL_MODE_LOWER, L.nd, h->d_S, std::max(1, L.nd), &lwork));
  {
    int lw2 = 0;
    const int nbp = std::max(1, std::min(L.nd, h->chol_nb));
    CUSOLVER_TRY(h, cusolverDnDpotrf_bufferSize(h->cusolver, CUBLAS_FILL_MODE_LOWER, nbp, h->d_S, std::max(1, L.nd), &lw2));
    lwork = std::max(lwork, lw2);
  }
  h->potrf_lwork = lwork;
  if (dev_alloc(h, &h->d_potrf_work, std::max(1, lwork))) return 1;

  // ---- groups of Schur blocks for the structured contraction ------------------------------------
  {
    // centroid sums of the measured pixels of every pattern point (this rank's observations;
    // b200ba_comm_init all-reduces them so that every rank derives the SAME grouping)
    const int nb = L.nblocks;
    h->grp_sums.assign(3 * static_cast<size_t>(std::max(1, nb)), 0.0);
    if (L.eliminate_points) {
      for (int64_t o = 0; o < h->n_obs; ++o) {
        const int p = static_cast<int>(h->h_obs_point[o]);
        h->grp_sums[p] += h->h_obs_xy[2 * o];
        h->grp_sums[nb + p] += h->h_obs_xy[2 * o + 1];
        h->grp_sums[2 * nb + p] += 1.0;
      }
    }
    if (reduce_group_sums(h)) return 1;
    if (build_groups(h, true)) return 1;
  }
  return 0;
}

int sync_stream(b200ba_handle* h) {
  CUDA_TRY(h, cudaStreamSynchronize(h->stream));
  resolve_timings(h);
  return 0;
}

int all_reduce(b200ba_handle* h, double* buf, size_t count) {
  if (h->n_ranks <= 1 || !h->comm) return 0;
  ScopedPhase ph(h, PH_ALLREDUCE);
  int rc = g_nccl.AllReduce(buf, buf, count, kNcclDouble, kNcclSum, h->comm, h->stream);
  if (rc != 0) {
    h->error = std::string("ncclAllReduce: ") + (g_nccl.GetErrorString ? g_nccl.GetErrorString(rc) : "error");
    return 1;
  }
  return 0;
}

// One residual pass on state `which` into `out`; jac selects Compute<true>/<false>.
// overlap_stragglers: launch the straggler pass on the side stream and return without joining
// (the caller joins with join_stragglers() after independent work).
int evaluate_state(b200ba_handle* h, int which, bool jac, const ObsOut& out, double huber, int phase,
                   bool overlap_stragglers = false) {
  launch_prepare_state(h->pb, h->L, h->st[which], h->n_control_total, h->stream);
  h->timings.kernel_launches += 1;
  {
    // [a .. mid]: main pass (what `roofline` is quoted on)
    cudaEvent_t a = get_event(h), mid = get_event(h);
    cudaEventRecord(a, h->stream);
    launch_residual_jacobian(h->uniform_model, jac, h->pb, h->L, h->st[which], h->d_last_projection, out, huber,
                             h->d_straggler_list, h->d_straggler_count, h->stream, mid);
    if (h->n_obs == 0) cudaEventRecord(mid, h->stream);
    h->pending.push_back({phase, a, mid, true});
    cudaStream_t ss = overlap_stragglers ? h->side_stream : h->stream;
    if (overlap_stragglers) CUDA_TRY(h, cudaStreamWaitEvent(h->side_stream, mid, 0));
    cudaEvent_t s0 = get_event(h), s1 = get_event(h);
    cudaEventRecord(s0, ss);
    launch_straggler_pass(h->uniform_model, jac, h->pb, h->L, h->st[which], h->d_last_projection, out, huber,
                          h->d_straggler_list, h->d_straggler_count, ss);
    cudaEventRecord(s1, ss);
    h->pending.push_back({PH_STRAGGLER, s0, s1, true});
    h->straggler_done = s1;  // stays valid until the next resolve_timings()
    h->timings.kernel_launches += 2 * (h->n_obs > 0);
  }
  CUDA_TRY(h, cudaGetLastError());
  return 0;
}

// Hot loop 1: H, b at the current state (LV/lm_optimizer.h:706-716).
int build_system(b200ba_handle* h, double huber, double* cost, double* n_valid, const b200ba_options* fix = nullptr) {
  // The straggler pass (a handful of observations burning the reference's full iteration
  // allowance) runs on the side stream underneath the accumulation of everything else.
  if (evaluate_state(h, h->cur, true, h->out, huber, PH_JAC, /*overlap_stragglers=*/true)) return 1;
  {
    ScopedPhase ph(h, PH_ACC);
    CUDA_TRY(h, cudaMemsetAsync(h->sys.base, 0, h->sys.total * sizeof(double), h->stream));
    launch_accumulate_scatter(h->pb, h->L, h->st[h->cur], h->out, h->sys, huber, h->stream);
    if (!h->L.localize_only || h->L.rig_in_state) {
      launch_accumulate_cells(h->pb, h->L, h->st[h->cur], h->out, h->sys, huber, h->stream);
      h->timings.kernel_launches += 1;
    }
    // join, then fold in the stragglers that succeeded after all
    CUDA_TRY(h, cudaStreamWaitEvent(h->stream, h->straggler_done, 0));
    launch_accumulate_list(h->pb, h->L, h->st[h->cur], h->out, h->sys, huber, h->d_straggler_list, h->d_straggler_count, h->stream);
    h->timings.kernel_launches += 2;
    launch_cost_reduce(h->n_obs, h->out.cost, nullptr, h->out.residual, h->d_partial, h->sys.scalars, h->stream);
    // trace(H) of this rank's partial system, for the lambda initialisation (lm_optimizer.h:766-781)
    launch_trace(h->L.nblocks, h->L.bs, h->sys.Dblk, h->L.nd, h->sys.C, h->sys.scalars + 8, h->stream);
    h->timings.kernel_launches += 3;
  }
  CUDA_TRY(h, cudaGetLastError());
  // ONE all-reduce per build covers D, b_p, B, b_d, the cost scalars and the trace
  if (all_reduce(h, h->sys.base, static_cast<size_t>(h->reduce_count))) return 1;
  if (fix) {
    // FixVariable (joint_optimization.cc:878-903): mask the fixed unknowns out of H, b. The trace for the
    // lambda initialisation was taken from the full H, like the reference does before it thins the system.
    const Layout& L = h->L;
    FixedRanges fr{};
    auto add = [&](int g0, int count) {
      if (count > 0 && fr.n < 4) {
        fr.lo[fr.n] = g0;
        fr.hi[fr.n] = g0 + count;
        fr.n++;
      }
    };
    int n_intr = 0;
    for (int c = 0; c < h->n_cameras; ++c) n_intr += update_parameter_count(h->cams_host[c]);
    if (fix->debug_fix_points) add(L.g_point, 3 * L.n_points);
    if (fix->debug_fix_poses) add(L.g_pose, 6 * L.n_imagesets);
    if (fix->debug_fix_rig_poses && L.rig_in_state) add(L.g_rig, 6 * L.n_cameras);
    if (fix->debug_fix_intrinsics && !L.localize_only) add(L.g_intr, n_intr);
    // with several ranks C is a partial sum: rank 0 alone carries the unit diagonal of the fixed dense unknowns
    launch_mask_fixed(L, h->sys, fr, h->rank == 0 ? 1.0 : 0.0, h->stream);
    h->timings.kernel_launches += 3;
  }
  CUDA_TRY(h, cudaMemcpyAsync(h->h_scal, h->sys.scalars, 9 * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  if (h->n_groups > 0 && h->L.nd > 0 && h->force_grouped != 0) {
    // exact column support of every group of Schur blocks, from the (global) B itself
    ScopedPhase ph(h, PH_SCHUR);
    CUDA_TRY(h, cudaMemsetAsync(h->d_flags, 0, static_cast<size_t>(h->n_groups) * h->L.nd, h->stream));
    launch_group_support(h->L.bs, h->L.nblocks, h->L.nd, h->sys.B, h->d_group_of_block, h->d_flags, h->stream);
    launch_compact_columns(h->n_groups, h->L.nd, h->d_flags, h->d_cols, h->d_count, h->stream);
    CUDA_TRY(h, cudaMemcpyAsync(h->h_count, h->d_count, h->n_groups * sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    h->timings.kernel_launches += 2;
  }
  if (sync_stream(h)) return 1;
  *cost = h->h_scal[3];
  *n_valid = h->h_scal[4];
  h->trace_H = h->h_scal[8];
  // cost model: grouped contraction sum_g k_g m_g^2 (+ scatter) against the dense n_d^2 * nbd
  h->use_grouped = false;
  if (h->n_groups > 0 && h->L.nd > 0 && h->force_grouped != 0) {
    double grouped = 0;
    for (int g = 0; g < h->n_groups; ++g) {
      const double kg = static_cast<double>(h->group_start[g + 1] - h->group_start[g]) * h->L.bs;
      const double mg = h->h_count[g];
      h->group_count[g] = h->h_count[g];
      grouped += mg * mg * (kg + 24.0);  // + ~24 flop-equivalents per scattered entry
    }
    const double dense = static_cast<double>(h->L.nd) * h->L.nd * h->L.nbd;
    h->use_grouped = (h->force_grouped == 1) || grouped < 0.6 * dense;
  }
  return 0;
}

__global__ void fold_panel_info_kernel(const int* __restrict__ panel_info, int n, int* __restrict__ info) {
  int worst = 0;
  for (int i = threadIdx.x; i < n; i += 32) worst = max(worst, panel_info[i] != 0 ? 1 : 0);
  for (int o = 16; o > 0; o >>= 1) worst = max(worst, __shfl_xor_sync(0xffffffffu, worst, o));
  if (threadIdx.x == 0) info[0] = worst;
}

// Cholesky factorisation of the reduced system S (column-major, lower) in place; info[0] != 0
// when a pivot was not positive.
//   default: cusolverDnDpotrf on the whole matrix (replicated on every rank).
//   B200BA_DIST_CHOL=1: right-looking factorisation over column blocks of width nb, dealt round-robin
//   to the ranks. The owner of block k factors its diagonal tile and solves the tile column below
//   it, broadcasts the finished column block (whole columns: contiguous in the column-major S),
//   then every rank applies the rank-nb update to the column blocks IT owns. Each rank thereby
//   does 1/N of the n^3/3 update flops instead of all of them, and ends with the complete L (its
//   own blocks computed, the others received), so the triangular solves stay local.
int factor_dense(b200ba_handle* h) {
  const Layout& L = h->L;
  const int nd = L.nd;
  const int nb = h->chol_nb;
  const int nblk = (nd + nb - 1) / nb;
  // Opt-in (B200BA_DIST_CHOL=1). Measured on config 2 (n_d = 13 080, nb = 512): the factorisation
  // takes 40.5 ms on one B200 and 32.1 ms on two, against 24.7 ms for cusolver's potrf replicated on
  // every rank -- the panel chain (potrf(512) + trsm + broadcast, ~0.75 ms x 26 panels) is serial
  // without look-ahead. Until the look-ahead / faster panel lands the
  // replicated potrf stays the default.
  bool blocked = (h->chol_mode == 1) && nd > nb;
  if (nblk > kMaxCholPanels) blocked = false;
  if (!blocked) {
    CUSOLVER_TRY(h, cusolverDnDpotrf(h->cusolver, CUBLAS_FILL_MODE_LOWER, nd, h->d_S, nd, h->d_potrf_work, h->potrf_lwork,
                                     h->d_info));
    return 0;
  }
  const double one = 1.0, minus_one = -1.0;
  int* panel_info = h->d_info + 2;
  CUDA_TRY(h, cudaMemsetAsync(panel_info, 0, nblk * sizeof(int), h->stream));
  for (int k = 0; k < nblk; ++k) {
    const int owner = k % h->n_ranks;
    const int j0 = k * nb, w = std::min(nb, nd - j0), below = nd - j0 - w;
    double* Akk = h->d_S + static_cast<size_t>(j0) * nd + j0;
    if (h->rank == owner) {
      CUSOLVER_TRY(h, cusolverDnDpotrf(h->cusolver, CUBLAS_FILL_MODE_LOWER, w, Akk, nd, h->d_potrf_work, h->potrf_lwork,
                                       panel_info + k));
      if (below > 0)  // L_below = A_below L_kk^-T
        CUBLAS_TRY(h, cublasDtrsm(h->cublas, CUBLAS_SIDE_RIGHT, CUBLAS_FILL_MODE_LOWER, CUBLAS_OP_T, CUBLAS_DIAG_NON_UNIT, below,
                                  w, &one, Akk, nd, Akk + w, nd));
    }
    if (h->n_ranks > 1 && h->comm) {
      double* col = h->d_S + static_cast<size_t>(j0) * nd;
      const int rc = g_nccl.Broadcast(col, col, static_cast<size_t>(nd) * w, kNcclDouble, owner, h->comm, h->stream);
      if (rc != 0) {
        h->error = std::string("ncclBroadcast: ") + (g_nccl.GetErrorString ? g_nccl.GetErrorString(rc) : "error");
        return 1;
      }
    }
    // trailing update of the column blocks this rank owns
    for (int j = k + 1; j < nblk; ++j) {
      if (j % h->n_ranks != h->rank) continue;
      const int c0 = j * nb, wj = std::min(nb, nd - c0);
      const double* Lj = h->d_S + static_cast<size_t>(j0) * nd + c0;  // rows c0.. of column block k
      CUBLAS_TRY(h, cublasDgemm(h->cublas, CUBLAS_OP_N, CUBLAS_OP_T, nd - c0, wj, w, &minus_one, Lj, nd, Lj, nd, &one,
                                h->d_S + static_cast<size_t>(c0) * nd + c0, nd));
    }
  }
  fold_panel_info_kernel<<<1, 32, 0, h->stream>>>(panel_info, nblk, h->d_info);
  if (h->n_ranks > 1 && h->comm) {
    // every rank must take the same branch of the LM loop
    const int rc = g_nccl.AllReduce(h->d_info, h->d_info, 1, kNcclInt32, kNcclMax, h->comm, h->stream);
    if (rc != 0) {
      h->error = "ncclAllReduce (factorisation status) failed";
      return 1;
    }
  }
  return 0;
}

// Hot loop 2: Schur complement solve for a given lambda (LV/lm_optimizer.h:1246-1369).
// Leaves x = [x_points | x_dense] in d_x. *spd = 0 when a factorisation met a non-positive pivot.
int solve_system_own(b200ba_handle* h, double lambda, int* spd);

int solve_system(b200ba_handle* h, double lambda, int* spd) {
  if (h->own_dense) return solve_system_own(h, lambda, spd);
  const Layout& L = h->L;
  const double one = 1.0, minus_one = -1.0;
  bool grouped_done = false;
  if (h->use_grouped) {
    // ---- structured contraction: per group gather -> compact rank-k update -> scatter ----------
    ScopedPhase ph(h, PH_SCHUR);
    const double zero = 0.0;
    CUDA_TRY(h, cudaMemsetAsync(h->d_fail, 0, sizeof(int), h->stream));
    launch_schur_blocks(L.bs, L.nblocks, h->sys.Dblk, h->sys.bp, lambda, h->d_Linv, h->d_v, h->d_fail, h->stream);
    // S = C_r is copied on the side stream (0.4 ms of pure HBM traffic at config 2) underneath the
    // block factorisations and the first rank-k update; the scatters follow it in stream order.
    CUDA_TRY(h, cudaEventRecord(h->ev_s_ready, h->stream));  // everything that used S before is done
    CUDA_TRY(h, cudaStreamWaitEvent(h->side_stream, h->ev_s_ready, 0));
    CUDA_TRY(h, cudaMemcpyAsync(h->d_S, h->sys.C, static_cast<size_t>(L.nd) * L.nd * sizeof(double),
                                cudaMemcpyDeviceToDevice, h->side_stream));
    CUDA_TRY(h, cudaEventRecord(h->ev_s_ready, h->side_stream));
    launch_block_solve_t(L.bs, L.nblocks, h->d_Linv, h->d_v, h->d_u, h->stream);  // u = D^-1 b_block
    h->timings.kernel_launches += 2;
    // The compact rank-k updates (compute-bound, main stream) and the scatters into S (memory-
    // bound, side stream) are software-pipelined over two P / Wc buffers. All scatters run on the
    // side stream, hence in order: plain read-modify-write, no atomics.
    int it = 0;
    for (int g = h->rank; g < h->n_groups; g += h->n_ranks) {
      const int nblk = h->group_start[g + 1] - h->group_start[g];
      const int kg = nblk * L.bs, mg = h->group_count[g];
      if (mg == 0 || kg == 0) continue;
      h->timings.contraction_flops += static_cast<double>(mg) * mg * kg;
      const int b = it & 1;
      double* Wc = h->d_Wc + b * h->wc_stride;
      double* P = h->d_P + b * h->p_stride;
      const int* cols = h->d_cols + static_cast<size_t>(g) * L.nd;
      if (it >= 2) CUDA_TRY(h, cudaStreamWaitEvent(h->stream, h->ev_scatter[b], 0));  // buffer b is free again
      launch_gather_scale(L.bs, nblk, L.nd, mg, mg, h->sys.B, h->d_Linv, h->d_group_blocks + h->group_start[g], cols, Wc,
                          h->stream);
      // row-major Wc [kg x mg] is the column-major mg x kg panel: P = Wc^T Wc (lower)
      CUBLAS_TRY(h, cublasDsyrk(h->cublas, CUBLAS_FILL_MODE_LOWER, CUBLAS_OP_N, mg, kg, &one, Wc, mg, &zero, P, mg));
      CUDA_TRY(h, cudaEventRecord(h->ev_syrk[b], h->stream));
      CUDA_TRY(h, cudaStreamWaitEvent(h->side_stream, h->ev_syrk[b], 0));
      launch_scatter_sub(L.nd, mg, cols, P, h->d_S, h->side_stream);
      CUDA_TRY(h, cudaEventRecord(h->ev_scatter[b], h->side_stream));
      h->timings.kernel_launches += 2;
      ++it;
    }
    // join: every scatter has landed before S is used
    CUDA_TRY(h, cudaStreamWaitEvent(h->stream, h->ev_s_ready, 0));  // the copy (covers a rank without groups)
    for (int b = 0; b < 2 && b < it; ++b) CUDA_TRY(h, cudaStreamWaitEvent(h->stream, h->ev_scatter[b], 0));
    grouped_done = true;
  }
  if (grouped_done) {
    if (h->n_ranks > 1) {
      if (all_reduce(h, h->d_S, static_cast<size_t>(L.nd) * L.nd)) return 1;
    }
    ScopedPhase ph(h, PH_SCHUR);
    launch_add_diagonal(L.nd, h->d_S, L.nd, lambda, h->stream);
    // x_dense <- b_d - B^T u   (B is global after the per-build all-reduce: no partial sums here)
    CUDA_TRY(h, cudaMemcpyAsync(h->d_x + L.nbd, h->sys.bd, L.nd * sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
    if (L.nbd > 0 && L.nd > 0)
      CUBLAS_TRY(h, cublasDgemv(h->cublas, CUBLAS_OP_N, L.nd, L.nbd, &minus_one, h->sys.B, L.nd, h->d_u, 1, &one,
                                h->d_x + L.nbd, 1));
    h->timings.kernel_launches += 1;
  } else {
    // Point range of this rank for the contraction: S = sum_r (C_r - W_r^T W_r) + lambda I, where
    // C_r is the rank's partial dense block and W_r the rows of W = L^-1 B of its points (B, D are
    // global after the per-build all-reduce). With one rank this is the plain S = C + lambda I - W^T W.
    const int p0 = static_cast<int>(static_cast<int64_t>(L.nblocks) * h->rank / h->n_ranks);
    const int p1 = static_cast<int>(static_cast<int64_t>(L.nblocks) * (h->rank + 1) / h->n_ranks);
    const int k_rows = L.bs * (p1 - p0);
    double* rhs_tail = h->d_S + static_cast<size_t>(L.nd) * L.nd;
    {
      ScopedPhase ph(h, PH_SCHUR);
      CUDA_TRY(h, cudaMemsetAsync(h->d_fail, 0, sizeof(int), h->stream));
      launch_schur_blocks(L.bs, L.nblocks, h->sys.Dblk, h->sys.bp, lambda, h->d_Linv, h->d_v, h->d_fail, h->stream);
      launch_schur_scale_rows(L.bs, L.nblocks, L.nd, h->sys.B, h->d_Linv, h->d_W, h->stream);
      CUDA_TRY(h, cudaMemcpyAsync(h->d_S, h->sys.C, static_cast<size_t>(L.nd) * L.nd * sizeof(double),
                                  cudaMemcpyDeviceToDevice, h->stream));
      CUDA_TRY(h, cudaMemsetAsync(rhs_tail, 0, L.nd * sizeof(double), h->stream));
      h->timings.kernel_launches += 2;
      if (k_rows > 0 && L.nd > 0) {
        h->timings.contraction_flops += static_cast<double>(L.nd) * L.nd * k_rows;
        const double zero = 0.0;
        const double* Wr = h->d_W + static_cast<size_t>(L.bs) * p0 * L.nd;
        // row-major W [nbd x nd] is the column-major nd x nbd matrix W^T: S -= W_r^T W_r (lower)
        CUBLAS_TRY(h, cublasDsyrk(h->cublas, CUBLAS_FILL_MODE_LOWER, CUBLAS_OP_N, L.nd, k_rows, &minus_one, Wr, L.nd,
                                  &one, h->d_S, L.nd));
        // partial reduced right-hand side: -W_r^T v_r
        CUBLAS_TRY(h, cublasDgemv(h->cublas, CUBLAS_OP_N, L.nd, k_rows, &minus_one, Wr, L.nd, h->d_v + L.bs * p0, 1, &zero,
                                  rhs_tail, 1));
      }
    }
    if (h->n_ranks > 1) {
      if (all_reduce(h, h->d_S, static_cast<size_t>(L.nd) * L.nd + L.nd)) return 1;
    }
    {
      ScopedPhase ph(h, PH_SCHUR);
      launch_add_diagonal(L.nd, h->d_S, L.nd, lambda, h->stream);
      // x_dense <- b_d - W^T v
      CUDA_TRY(h, cudaMemcpyAsync(h->d_x + L.nbd, h->sys.bd, L.nd * sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
      CUBLAS_TRY(h, cublasDaxpy(h->cublas, L.nd, &one, rhs_tail, 1, h->d_x + L.nbd, 1));
      h->timings.kernel_launches += 1;
    }
  }
  {
    ScopedPhase ph(h, PH_FACTOR);
    if (factor_dense(h)) return 1;
    h->timings.factor_flops += static_cast<double>(L.nd) * L.nd * L.nd / 3.0;
  }
  {
    ScopedPhase ph(h, PH_SOLVE);
    CUSOLVER_TRY(h, cusolverDnDpotrs(h->cusolver, CUBLAS_FILL_MODE_LOWER, L.nd, 1, h->d_S, L.nd, h->d_x + L.nbd, L.nd,
                                     h->d_info + 1));
  }
  if (grouped_done) {
    ScopedPhase ph(h, PH_SCHUR);
    // t = B x_d ; x_block = u - D^-1 t   (W is never materialised on this path)
    const double zero = 0.0;
    if (L.nbd > 0 && L.nd > 0)
      CUBLAS_TRY(h, cublasDgemv(h->cublas, CUBLAS_OP_T, L.nd, L.nbd, &one, h->sys.B, L.nd, h->d_x + L.nbd, 1, &zero,
                                h->d_y, 1));
    launch_block_backsub2(L.bs, L.nblocks, h->d_Linv, h->d_u, h->d_y, h->d_x, h->stream);
    h->timings.kernel_launches += 1;
  } else {
    ScopedPhase ph(h, PH_SCHUR);
    // y = v - W x_d ; x_p = L^-T y
    CUDA_TRY(h, cudaMemcpyAsync(h->d_y, h->d_v, L.nbd * sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
    if (L.nbd > 0 && L.nd > 0)
      CUBLAS_TRY(h, cublasDgemv(h->cublas, CUBLAS_OP_T, L.nd, L.nbd, &minus_one, h->d_W, L.nd, h->d_x + L.nbd, 1, &one,
                                h->d_y, 1));
    launch_schur_backsub(L.bs, L.nblocks, h->d_Linv, h->d_y, h->d_x, h->stream);
    h->timings.kernel_launches += 1;
  }
  CUDA_TRY(h, cudaMemcpyAsync(h->h_flags, h->d_info, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
  CUDA_TRY(h, cudaMemcpyAsync(h->h_flags + 1, h->d_fail, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
  if (sync_stream(h)) return 1;
  *spd = (h->h_flags[0] == 0 && h->h_flags[1] == 0) ? 1 : 0;
  return 0;
}

// The same solve on the in-tree kernels (ba_dense.cu): the contraction is a DMMA product whose epilogue
// scatters straight into S, the reduced system is factorised by a blocked right-looking Cholesky with
// look-ahead (block columns dealt cyclically to the ranks, panels broadcast over NVLink), the
// triangular solves run on the packed factor. With several ranks S is formed as partial sums
// S_r = C_r - sum_{g of r} W_g^T W_g that ONE reduce-scatter per attempt turns into the block columns each
// rank owns; nobody holds or reduces the whole matrix.
int solve_system_own(b200ba_handle* h, double lambda, int* spd) {
  const Layout& L = h->L;
  DenseCtx& d = h->dn;
  const int nd = L.nd, R = h->n_ranks;
  {
    ScopedPhase ph(h, PH_SCHUR);
    CUDA_TRY(h, cudaMemsetAsync(h->d_fail, 0, sizeof(int), h->stream));
    CUDA_TRY(h, cudaMemsetAsync(h->d_info, 0, sizeof(int), h->stream));
    launch_schur_blocks(L.bs, L.nblocks, h->sys.Dblk, h->sys.bp, lambda, h->d_Linv, h->d_v, h->d_fail, h->stream);
    // S <- C_r (this rank's partial dense block) through the storage map, on the side stream
    // underneath the block factorisations and the first gather
    CUDA_TRY(h, cudaEventRecord(h->ev_s_ready, h->stream));  // everything that used S before is done
    CUDA_TRY(h, cudaStreamWaitEvent(h->side_stream, h->ev_s_ready, 0));
    if (nd > 0) {
      if (R == 1) {
        CUDA_TRY(h, cudaMemcpy2DAsync(h->d_S, d.map.ld * sizeof(double), h->sys.C, static_cast<size_t>(nd) * sizeof(double),
                                      static_cast<size_t>(nd) * sizeof(double), nd, cudaMemcpyDeviceToDevice, h->side_stream));
      } else {
        for (int j = 0; j < d.nblk; ++j) {
          const int c0 = j * d.NB, w = std::min(d.NB, nd - c0);
          CUDA_TRY(h, cudaMemcpy2DAsync(h->d_S + d.map.col_offset(c0), d.map.ld * sizeof(double),
                                        h->sys.C + static_cast<size_t>(c0) * nd, static_cast<size_t>(nd) * sizeof(double),
                                        static_cast<size_t>(nd) * sizeof(double), w, cudaMemcpyDeviceToDevice, h->side_stream));
        }
      }
    }
    CUDA_TRY(h, cudaEventRecord(h->ev_s_ready, h->side_stream));
    launch_block_solve_t(L.bs, L.nblocks, h->d_Linv, h->d_v, h->d_u, h->stream);  // u = D^-1 b_block
    h->timings.kernel_launches += 2;
    bool joined = false;
    auto join_copy = [&]() {
      if (!joined) cudaStreamWaitEvent(h->stream, h->ev_s_ready, 0);
      joined = true;
    };
    if (h->use_grouped) {
      // structured contraction: per group gather -> DMMA rank-k update with the scatter epilogue
      int it = 0;
      for (int g = h->rank; g < h->n_groups; g += R) {
        const int nblk = h->group_start[g + 1] - h->group_start[g];
        const int kg = nblk * L.bs, mg = h->group_count[g];
        if (mg == 0 || kg == 0) continue;
        const int ldw = (mg + 1) / 2 * 2;
        double* Wc = h->d_Wc + (it & 1) * h->wc_stride;
        const int* cols = h->d_cols + static_cast<size_t>(g) * nd;
        launch_gather_scale(L.bs, nblk, nd, mg, ldw, h->sys.B, h->d_Linv, h->d_group_blocks + h->group_start[g], cols, Wc,
                            h->stream);
        join_copy();
        GemmArgs ga{};
        ga.M = ga.N = mg;
        ga.K = kg;
        ga.A = ga.B = Wc;
        ga.lda = ga.ldb = ldw;
        ga.C = h->d_S;
        ga.alpha = -1.0;
        ga.beta = 1.0;
        ga.a_aligned = ga.b_aligned = gemm_operand_aligned(Wc, ldw);
        ga.cols = cols;
        ga.map = d.map;
        if (launch_dgemm_nt(ga, true, true, h->stream)) {
          h->error = "dgemm_nt (contraction) launch failed";
          return 1;
        }
        h->timings.contraction_flops += static_cast<double>(mg) * mg * kg;
        h->timings.kernel_launches += 2;
        ++it;
      }
    } else if (L.nbd > 0 && nd > 0) {
      // dense contraction over this rank's slice of the Schur blocks: S_r -= W_r^T W_r
      const int p0 = static_cast<int>(static_cast<int64_t>(L.nblocks) * h->rank / R);
      const int p1 = static_cast<int>(static_cast<int64_t>(L.nblocks) * (h->rank + 1) / R);
      const int k_rows = L.bs * (p1 - p0);
      launch_schur_scale_rows(L.bs, L.nblocks, nd, h->sys.B, h->d_Linv, h->d_W, h->stream);
      h->timings.kernel_launches += 1;
      if (k_rows > 0) {
        const double* Wr = h->d_W + static_cast<size_t>(L.bs) * p0 * nd;
        join_copy();
        GemmArgs ga{};
        ga.M = ga.N = nd;
        ga.K = k_rows;
        ga.A = ga.B = Wr;
        ga.lda = ga.ldb = nd;
        ga.C = h->d_S;
        ga.ldc = d.map.ld;
        ga.alpha = -1.0;
        ga.beta = 1.0;
        ga.a_aligned = ga.b_aligned = gemm_operand_aligned(Wr, nd);
        ga.cols = h->d_ident_cols;
        ga.map = d.map;
        if (launch_dgemm_nt(ga, true, /*scatter=*/R > 1, h->stream)) {
          h->error = "dgemm_nt (contraction) launch failed";
          return 1;
        }
        h->timings.contraction_flops += static_cast<double>(nd) * nd * k_rows;
        h->timings.kernel_launches += 1;
      }
    }
    join_copy();
    if (h->rank == 0) launch_add_diagonal_map(nd, h->d_S, d.map, lambda, h->stream);
    // x_dense <- b_d - B^T u   (B, D, b are global after the per-build all-reduce)
    CUDA_TRY(h, cudaMemcpyAsync(h->d_x + L.nbd, h->sys.bd, nd * sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
    if (L.nbd > 0 && nd > 0) launch_gemv_t(L.nbd, nd, nd, h->sys.B, h->d_u, -1.0, h->d_x + L.nbd, h->d_gemv_partial, h->stream);
    h->timings.kernel_launches += 4;
  }
  if (R > 1 && nd > 0) {
    // partial sums -> the block columns each rank owns (in place: rank r keeps chunk r)
    ScopedPhase ph(h, PH_ALLREDUCE);
    const int rc = g_nccl.ReduceScatter(h->d_S, h->d_S + static_cast<size_t>(h->rank) * d.chunk, static_cast<size_t>(d.chunk),
                                        kNcclDouble, kNcclSum, h->comm, h->stream);
    if (rc != 0) {
      h->error = std::string("ncclReduceScatter: ") + (g_nccl.GetErrorString ? g_nccl.GetErrorString(rc) : "error");
      return 1;
    }
  }
  {
    ScopedPhase ph(h, PH_FACTOR);
    if (dense_factor(&d)) {
      if (h->error.empty()) h->error = "dense_factor: launch failed";
      return 1;
    }
    h->timings.factor_flops += static_cast<double>(nd) * nd * nd / 3.0;
    if (R > 1) {
      // every rank must take the same branch of the LM loop
      const int rc = g_nccl.AllReduce(h->d_info, h->d_info, 1, kNcclInt32, kNcclMax, h->comm, h->stream);
      if (rc != 0) {
        h->error = "ncclAllReduce (factorisation status) failed";
        return 1;
      }
    }
  }
  {
    ScopedPhase ph(h, PH_SOLVE);
    if (dense_solve(&d, h->d_x + L.nbd)) {
      h->error = "dense_solve: launch failed";
      return 1;
    }
    h->timings.kernel_launches += 4 * d.ntiles;
  }
  {
    ScopedPhase ph(h, PH_SCHUR);
    // t = B x_d ; x_block = u - D^-1 t   (W is never needed for the back-substitution)
    if (L.nbd > 0 && nd > 0)
      launch_gemv_n(L.nbd, nd, nd, h->sys.B, h->d_x + L.nbd, h->d_y, h->stream);
    else if (L.nbd > 0)
      CUDA_TRY(h, cudaMemsetAsync(h->d_y, 0, L.nbd * sizeof(double), h->stream));
    launch_block_backsub2(L.bs, L.nblocks, h->d_Linv, h->d_u, h->d_y, h->d_x, h->stream);
    h->timings.kernel_launches += 1;
  }
  CUDA_TRY(h, cudaMemcpyAsync(h->h_flags, h->d_info, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
  CUDA_TRY(h, cudaMemcpyAsync(h->h_flags + 1, h->d_fail, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
  if (sync_stream(h)) return 1;
  *spd = (h->h_flags[0] == 0 && h->h_flags[1] == 0) ? 1 : 0;
  return 0;
}

int check_ready(b200ba_handle* h, const b200ba_options* opt) {
  if (!h) return 1;
  if (!opt) {
    h->error = "options are NULL";
    return 2;
  }
  if (!h->have_state) {
    h->error = "no state: call b200ba_set_state first";
    return 2;
  }
  CUDA_TRY(h, cudaSetDevice(h->device));
  return make_layout(h, opt);
}

void free_handle_buffers(b200ba_handle* h) {
  auto F = [](auto*& p) {
    if (p) cudaFree(p);
    p = nullptr;
  };
  F(h->d_obs_imageset); F(h->d_obs_camera); F(h->d_obs_point); F(h->d_obs_xy);
  for (int i = 0; i < 2; ++i) {
    F(h->st[i].points); F(h->st[i].rig_tr_global); F(h->st[i].camera_tr_rig); F(h->st[i].intrinsics);
    F(h->st[i].image_tr_global); F(h->st[i].tangents);
  }
  F(h->d_last_projection);
  F(h->snap.points); F(h->snap.rig_tr_global); F(h->snap.camera_tr_rig); F(h->snap.intrinsics); F(h->d_snap_lp);
  h->have_snapshot = false;
  F(h->out.residual); F(h->out.cost); F(h->out.jac); F(h->out.cjac); F(h->out.cell); F(h->out.has_jac); F(h->out.evals);
  F(h->out_trial.residual); F(h->out_trial.cost);
  F(h->sys.base); F(h->d_W); F(h->d_S); F(h->d_Linv); F(h->d_v); F(h->d_y); F(h->d_x); F(h->d_potrf_work);
  F(h->d_info); F(h->d_fail); F(h->d_straggler_list); F(h->d_straggler_count); F(h->d_perm); F(h->d_lp_stage);
  F(h->d_group_of_block); F(h->d_group_blocks); F(h->d_flags); F(h->d_cols); F(h->d_count); F(h->d_Wc); F(h->d_P); F(h->d_u);
  if (h->h_count) cudaFreeHost(h->h_count);
  h->h_count = nullptr;
  F(h->d_partial); F(h->d_scal); F(h->d_rot);
  F(h->dn.Lpack); F(h->dn.tmp); F(h->dn.d_panel_off); F(h->dn.d_panel_h); F(h->d_ident_cols); F(h->d_gemv_partial);
  h->dn.S = nullptr;
  h->dense_planned_n = -1;
  if (h->h_scal) cudaFreeHost(h->h_scal);
  if (h->h_flags) cudaFreeHost(h->h_flags);
  h->h_scal = nullptr;
  h->h_flags = nullptr;
}

}  // namespace

// ==============================================================================================
// C ABI
// ==============================================================================================
extern "C" {

const char* b200ba_version(void) { return "b200ba 0.1.0 (sm_100a, FP64)"; }

int64_t b200ba_intrinsics_size(const b200ba_camera* cam) { return cam ? intrinsics_size(*cam) : 0; }
int32_t b200ba_update_parameter_count(const b200ba_camera* cam) { return cam ? update_parameter_count(*cam) : 0; }

void b200ba_default_options(b200ba_options* o) {
  if (!o) return;
  o->max_iteration_count = 1;
  o->init_lambda = -1.0;
  o->numerical_diff_delta = 1e-4;   // APP/calibration.cc:201
  o->regularization_weight = 0.0;
  o->localize_only = 0;
  o->eliminate_points = 1;
  o->schur_mode = B200BA_SCHUR_DENSE;
  o->max_lm_attempts = 50;          // joint_optimization.cc:920
  o->init_lambda_factor = 1e-5;     // joint_optimization.cc:922
  o->huber_parameter = 1.0;         // joint_optimization.cc:346
  o->jacobian_mode = B200BA_JACOBIAN_ANALYTIC;
  o->print_progress = 0;
  o->debug_verify_cost = 0;
  o->debug_fix_points = o->debug_fix_poses = o->debug_fix_rig_poses = o->debug_fix_intrinsics = 0;
}

const char* b200ba_last_error(const b200ba_handle* h) { return h ? h->error.c_str() : g_create_error.c_str(); }

int b200ba_create(const b200ba_problem* p, int device, b200ba_handle** out) {
  if (!p || !out) {
    g_create_error = "NULL argument";
    return 2;
  }
  *out = nullptr;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    g_create_error = "no CUDA device available (this library has no CPU fallback)";
    return 3;
  }
  if (p->n_cameras < 1 || p->n_cameras > kMaxCameras) {
    g_create_error = "n_cameras out of range (1..8)";
    return 2;
  }
  for (int c = 0; c < p->n_cameras; ++c) {
    const int t = p->cameras[c].model_type;
    if (t != B200BA_MODEL_CENTRAL_GENERIC && t != B200BA_MODEL_NONCENTRAL_GENERIC && t != B200BA_MODEL_CENTRAL_OPENCV) {
      g_create_error = "camera model not on the accelerated path (central-generic, noncentral-generic, central-opencv)";
      return 2;
    }
    if (t != B200BA_MODEL_CENTRAL_OPENCV && (p->cameras[c].grid_width < 4 || p->cameras[c].grid_height < 4)) {
      g_create_error = "generic models need a grid of at least 4x4 control points";
      return 2;
    }
  }
  if (p->n_obs >= (1LL << 31)) {
    g_create_error = "n_obs must be < 2^31";
    return 2;
  }
  for (int64_t o = 0; o < p->n_obs; ++o) {
    if (p->obs_imageset[o] >= static_cast<uint32_t>(p->n_imagesets) || p->obs_point[o] >= static_cast<uint32_t>(p->n_points) ||
        p->obs_camera[o] >= static_cast<uint32_t>(p->n_cameras)) {
      g_create_error = "observation index out of range";
      return 2;
    }
  }
  b200ba_handle* h = new b200ba_handle();
  auto fail = [&](int rc) {
    g_create_error = h->error;
    free_handle_buffers(h);
    if (h->cublas) cublasDestroy(h->cublas);
    if (h->cusolver) cusolverDnDestroy(h->cusolver);
    if (h->stream) cudaStreamDestroy(h->stream);
    if (h->side_stream) cudaStreamDestroy(h->side_stream);
    delete h;
    return rc;
  };
  if (device < 0) {
    if (cudaGetDevice(&device) != cudaSuccess) device = 0;
  }
  h->device = device;
#define TRYC(expr) \
  if ((expr) != 0) return fail(1)
  auto cuda_ok = [&](cudaError_t e, const char* what) {
    if (e != cudaSuccess) {
      h->error = std::string(what) + ": " + cudaGetErrorString(e);
      return 1;
    }
    return 0;
  };
  TRYC(cuda_ok(cudaSetDevice(device), "cudaSetDevice"));
  TRYC(cuda_ok(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking), "cudaStreamCreate"));
  TRYC(cuda_ok(cudaStreamCreateWithFlags(&h->side_stream, cudaStreamNonBlocking), "cudaStreamCreate"));
  if (cublasCreate(&h->cublas) != CUBLAS_STATUS_SUCCESS) {
    h->error = "cublasCreate failed";
    return fail(1);
  }
  cublasSetStream(h->cublas, h->stream);
  if (cusolverDnCreate(&h->cusolver) != CUSOLVER_STATUS_SUCCESS) {
    h->error = "cusolverDnCreate failed";
    return fail(1);
  }
  cusolverDnSetStream(h->cusolver, h->stream);

  h->n_cameras = p->n_cameras;
  h->n_imagesets = p->n_imagesets;
  h->n_points = p->n_points;
  h->n_obs = p->n_obs;
  h->cams_host.assign(p->cameras, p->cameras + p->n_cameras);
  h->uniform_model = h->cams_host[0].model_type;
  int upd = 0;
  for (int c = 0; c < p->n_cameras; ++c) {
    CamDev& d = h->pb.cams[c];
    fill_camdev(h->cams_host[c], &d);
    d.intr_off = h->intr_total;
    d.tan_off = h->tan_total;
    d.upd_off = upd;
    upd += d.upd_count;
    h->intr_total += intrinsics_size(h->cams_host[c]);
    const int64_t G = static_cast<int64_t>(d.gw) * d.gh;
    h->tan_total += 6 * G;
    h->n_control_total += G;
    if (d.gw == 0) h->n_param_total += 12;
    if (h->cams_host[c].model_type != h->uniform_model) h->uniform_model = -1;
  }
  const int64_t n = p->n_obs;
  TRYC(dev_alloc(h, &h->d_obs_imageset, n));
  TRYC(dev_alloc(h, &h->d_obs_camera, n));
  TRYC(dev_alloc(h, &h->d_obs_point, n));
  TRYC(dev_alloc(h, &h->d_obs_xy, n));
  if (n > 0) {
    // Static cell-major order: sort the observations ONCE by (camera, B-spline cell of the
    // measured pixel). Lanes of a warp then gather the same 4x4 control points (broadcast loads
    // instead of 32 scattered L1 wavefronts) and accumulate_cells_kernel sees long runs. The
    // C ABI keeps the reference's residual order: inputs are permuted here, per-observation
    // outputs are un-permuted on the way out.
    std::vector<uint32_t> key(n);
    for (int64_t o = 0; o < n; ++o) {
      const uint32_t cam = p->obs_camera[o];
      const CamDev& cd = h->pb.cams[cam];
      uint32_t cell = 0;
      if (cd.gw > 0) {
        const double gx = 1.0 + cd.gmul_x * (static_cast<double>(p->obs_xy[2 * o]) - cd.min_x);
        const double gy = 1.0 + cd.gmul_y * (static_cast<double>(p->obs_xy[2 * o + 1]) - cd.min_y);
        const int x0 = std::min(std::max(static_cast<int>(std::floor(gx)) - 1, 0), cd.gw - 4);
        const int y0 = std::min(std::max(static_cast<int>(std::floor(gy)) - 1, 0), cd.gh - 4);
        cell = static_cast<uint32_t>(x0 + y0 * cd.gw);
      }
      key[o] = (cam << 24) | cell;
    }
    h->perm.resize(n);
    for (int64_t o = 0; o < n; ++o) h->perm[o] = static_cast<uint32_t>(o);
    std::stable_sort(h->perm.begin(), h->perm.end(), [&](uint32_t a, uint32_t b) { return key[a] < key[b]; });
    std::vector<uint32_t> t32(n);
    std::vector<float> txy(2 * n);
    for (int64_t i = 0; i < n; ++i) t32[i] = p->obs_imageset[h->perm[i]];
    TRYC(cuda_ok(cudaMemcpy(h->d_obs_imageset, t32.data(), n * sizeof(uint32_t), cudaMemcpyHostToDevice), "H2D"));
    for (int64_t i = 0; i < n; ++i) t32[i] = p->obs_camera[h->perm[i]];
    TRYC(cuda_ok(cudaMemcpy(h->d_obs_camera, t32.data(), n * sizeof(uint32_t), cudaMemcpyHostToDevice), "H2D"));
    for (int64_t i = 0; i < n; ++i) t32[i] = p->obs_point[h->perm[i]];
    TRYC(cuda_ok(cudaMemcpy(h->d_obs_point, t32.data(), n * sizeof(uint32_t), cudaMemcpyHostToDevice), "H2D"));
    for (int64_t i = 0; i < n; ++i) {
      txy[2 * i] = p->obs_xy[2 * h->perm[i]];
      txy[2 * i + 1] = p->obs_xy[2 * h->perm[i] + 1];
    }
    TRYC(cuda_ok(cudaMemcpy(h->d_obs_xy, txy.data(), n * sizeof(float2), cudaMemcpyHostToDevice), "H2D"));
    TRYC(dev_alloc(h, &h->d_perm, n));
    TRYC(cuda_ok(cudaMemcpy(h->d_perm, h->perm.data(), n * sizeof(uint32_t), cudaMemcpyHostToDevice), "H2D"));
  }
  TRYC(dev_alloc(h, &h->d_lp_stage, n));
  if (n > 0) {
    h->h_obs_imageset.assign(p->obs_imageset, p->obs_imageset + n);
    h->h_obs_camera.assign(p->obs_camera, p->obs_camera + n);
    h->h_obs_point.assign(p->obs_point, p->obs_point + n);
    h->h_obs_xy.assign(p->obs_xy, p->obs_xy + 2 * n);
  }
  if (const char* e = getenv("B200BA_DENSE")) h->own_dense = !(strcmp(e, "lib") == 0 || strcmp(e, "0") == 0);
  if (const char* e = getenv("B200BA_DENSE_NB")) h->dense_nb = std::max(128, atoi(e) / 128 * 128);
  set_gemm_sm_reserve(getenv("B200BA_PANEL_SMS") ? atoi(getenv("B200BA_PANEL_SMS")) : 8);
  {
    int lo = 0, hi = 0;
    cudaDeviceGetStreamPriorityRange(&lo, &hi);
    TRYC(cuda_ok(cudaStreamCreateWithPriority(&h->panel_stream, cudaStreamNonBlocking, hi), "cudaStreamCreate"));
    TRYC(cuda_ok(cudaStreamCreateWithPriority(&h->aux_stream, cudaStreamNonBlocking, hi), "cudaStreamCreate"));
  }
  if (const char* e = getenv("B200BA_GROUPED")) h->force_grouped = atoi(e);
  if (const char* e = getenv("B200BA_COMPACT_J")) h->compact_j = atoi(e) != 0;
  if (const char* e = getenv("B200BA_DIST_CHOL")) h->chol_mode = atoi(e);
  if (const char* e = getenv("B200BA_CHOL_NB")) h->chol_nb = std::max(32, atoi(e));
  h->pb.n_obs = n;
  h->pb.obs_imageset = h->d_obs_imageset;
  h->pb.obs_camera = h->d_obs_camera;
  h->pb.obs_point = h->d_obs_point;
  h->pb.obs_xy = h->d_obs_xy;
  for (int i = 0; i < 2; ++i) {
    TRYC(dev_alloc(h, &h->st[i].points, 3 * static_cast<size_t>(h->n_points)));
    TRYC(dev_alloc(h, &h->st[i].rig_tr_global, 7 * static_cast<size_t>(h->n_imagesets)));
    TRYC(dev_alloc(h, &h->st[i].camera_tr_rig, 7 * static_cast<size_t>(h->n_cameras)));
    TRYC(dev_alloc(h, &h->st[i].intrinsics, h->intr_total));
    TRYC(dev_alloc(h, &h->st[i].image_tr_global, 12 * static_cast<size_t>(h->n_imagesets) * h->n_cameras));
    TRYC(dev_alloc(h, &h->st[i].tangents, h->tan_total));
  }
  TRYC(dev_alloc(h, &h->d_last_projection, n));
  TRYC(cuda_ok(cudaMemset(h->d_last_projection, 0, std::max<int64_t>(1, n) * sizeof(double2)), "memset"));
  TRYC(dev_alloc(h, &h->d_straggler_list, n));
  TRYC(dev_alloc(h, &h->d_straggler_count, 1));
  TRYC(dev_alloc(h, &h->d_info, 2 + kMaxCholPanels));
  TRYC(dev_alloc(h, &h->d_fail, 1));
  TRYC(dev_alloc(h, &h->d_partial, cost_reduce_partial_size()));
  TRYC(dev_alloc(h, &h->d_scal, 16));
  TRYC(cuda_ok(cudaMallocHost(reinterpret_cast<void**>(&h->h_scal), 16 * sizeof(double)), "cudaMallocHost"));
  TRYC(cuda_ok(cudaMallocHost(reinterpret_cast<void**>(&h->h_flags), 4 * sizeof(int)), "cudaMallocHost"));
#undef TRYC
  *out = h;
  return 0;
}

void b200ba_destroy(b200ba_handle* h) {
  if (!h) return;
  cudaSetDevice(h->device);
  if (h->stream) cudaStreamSynchronize(h->stream);
  if (h->comm && g_nccl.CommDestroy) g_nccl.CommDestroy(h->comm);
  resolve_timings(h);
  for (auto e : h->event_pool) cudaEventDestroy(e);
  for (int i = 0; i < 2; ++i) {
    if (h->ev_syrk[i]) cudaEventDestroy(h->ev_syrk[i]);
    if (h->ev_scatter[i]) cudaEventDestroy(h->ev_scatter[i]);
  }
  if (h->ev_s_ready) cudaEventDestroy(h->ev_s_ready);
  free_handle_buffers(h);
  if (h->cublas) cublasDestroy(h->cublas);
  if (h->cusolver) cusolverDnDestroy(h->cusolver);
  if (h->stream) cudaStreamDestroy(h->stream);
  if (h->side_stream) cudaStreamDestroy(h->side_stream);
  if (h->panel_stream) cudaStreamDestroy(h->panel_stream);
  if (h->aux_stream) cudaStreamDestroy(h->aux_stream);
  for (int i = 0; i < 2; ++i) {
    if (h->dn.ev_ready[i]) cudaEventDestroy(h->dn.ev_ready[i]);
    if (h->dn.ev_main[i]) cudaEventDestroy(h->dn.ev_main[i]);
    if (h->dn.ev_half2[i]) cudaEventDestroy(h->dn.ev_half2[i]);
  }
  if (h->dn.ev_misc) cudaEventDestroy(h->dn.ev_misc);
  delete h;
}

int b200ba_set_state(b200ba_handle* h, const b200ba_state* s) {
  if (!h) return 1;
  if (!s || !s->points || !s->rig_tr_global || !s->camera_tr_rig || !s->intrinsics) {
    h->error = "NULL state array";
    return 2;
  }
  CUDA_TRY(h, cudaSetDevice(h->device));
  StateDev& d = h->st[h->cur];
  CUDA_TRY(h, cudaMemcpyAsync(d.points, s->points, 3 * sizeof(double) * h->n_points, cudaMemcpyHostToDevice, h->stream));
  CUDA_TRY(h, cudaMemcpyAsync(d.rig_tr_global, s->rig_tr_global, 7 * sizeof(double) * h->n_imagesets, cudaMemcpyHostToDevice, h->stream));
  CUDA_TRY(h, cudaMemcpyAsync(d.camera_tr_rig, s->camera_tr_rig, 7 * sizeof(double) * h->n_cameras, cudaMemcpyHostToDevice, h->stream));
  for (int c = 0; c < h->n_cameras; ++c)
    CUDA_TRY(h, cudaMemcpyAsync(d.intrinsics + h->pb.cams[c].intr_off, s->intrinsics[c],
                                sizeof(double) * intrinsics_size(h->cams_host[c]), cudaMemcpyHostToDevice, h->stream));
  if (s->last_projection) {
    // caller's order -> device staging -> cell-major order (gather on the device)
    CUDA_TRY(h, cudaMemcpyAsync(h->d_lp_stage, s->last_projection, 2 * sizeof(double) * h->n_obs, cudaMemcpyHostToDevice, h->stream));
    launch_permute_double2(h->n_obs, h->d_perm, h->d_lp_stage, h->d_last_projection, /*scatter=*/false, h->stream);
  } else
    CUDA_TRY(h, cudaMemsetAsync(h->d_last_projection, 0, std::max<int64_t>(1, h->n_obs) * sizeof(double2), h->stream));
  CUDA_TRY(h, cudaStreamSynchronize(h->stream));
  h->have_state = true;
  return 0;
}

int b200ba_get_state(b200ba_handle* h, b200ba_state* s) {
  if (!h) return 1;
  if (!s || !h->have_state) {
    h->error = "no state";
    return 2;
  }
  CUDA_TRY(h, cudaSetDevice(h->device));
  const StateDev& d = h->st[h->cur];
  if (s->points) CUDA_TRY(h, cudaMemcpyAsync(s->points, d.points, 3 * sizeof(double) * h->n_points, cudaMemcpyDeviceToHost, h->stream));
  if (s->rig_tr_global) CUDA_TRY(h, cudaMemcpyAsync(s->rig_tr_global, d.rig_tr_global, 7 * sizeof(double) * h->n_imagesets, cudaMemcpyDeviceToHost, h->stream));
  if (s->camera_tr_rig) CUDA_TRY(h, cudaMemcpyAsync(s->camera_tr_rig, d.camera_tr_rig, 7 * sizeof(double) * h->n_cameras, cudaMemcpyDeviceToHost, h->stream));
  if (s->intrinsics)
    for (int c = 0; c < h->n_cameras; ++c)
      CUDA_TRY(h, cudaMemcpyAsync(s->intrinsics[c], d.intrinsics + h->pb.cams[c].intr_off,
                                  sizeof(double) * intrinsics_size(h->cams_host[c]), cudaMemcpyDeviceToHost, h->stream));
  if (s->last_projection) {
    launch_permute_double2(h->n_obs, h->d_perm, h->d_last_projection, h->d_lp_stage, /*scatter=*/true, h->stream);
    CUDA_TRY(h, cudaMemcpyAsync(s->last_projection, h->d_lp_stage, 2 * sizeof(double) * h->n_obs, cudaMemcpyDeviceToHost, h->stream));
  }
  CUDA_TRY(h, cudaStreamSynchronize(h->stream));
  return 0;
}

// Device-side copy of the optimised state and the warm-start cache (no host round trip). Used to
// restart a trajectory from the same point (bench.py) and by callers that want to roll back.
static int copy_state_dev(b200ba_handle* h, const StateDev& src, const double2* src_lp, StateDev& dst, double2* dst_lp) {
  CUDA_TRY(h, cudaMemcpyAsync(dst.points, src.points, 3 * sizeof(double) * h->n_points, cudaMemcpyDeviceToDevice, h->stream));
  CUDA_TRY(h, cudaMemcpyAsync(dst.rig_tr_global, src.rig_tr_global, 7 * sizeof(double) * h->n_imagesets, cudaMemcpyDeviceToDevice, h->stream));
  CUDA_TRY(h, cudaMemcpyAsync(dst.camera_tr_rig, src.camera_tr_rig, 7 * sizeof(double) * h->n_cameras, cudaMemcpyDeviceToDevice, h->stream));
  CUDA_TRY(h, cudaMemcpyAsync(dst.intrinsics, src.intrinsics, sizeof(double) * h->intr_total, cudaMemcpyDeviceToDevice, h->stream));
  CUDA_TRY(h, cudaMemcpyAsync(dst_lp, src_lp, sizeof(double2) * std::max<int64_t>(1, h->n_obs), cudaMemcpyDeviceToDevice, h->stream));
  CUDA_TRY(h, cudaStreamSynchronize(h->stream));
  return 0;
}

int b200ba_snapshot_state(b200ba_handle* h) {
  if (!h) return 1;
  if (!h->have_state) {
    h->error = "no state";
    return 2;
  }
  CUDA_TRY(h, cudaSetDevice(h->device));
  if (!h->snap.points) {
    if (dev_alloc(h, &h->snap.points, 3 * static_cast<size_t>(h->n_points))) return 1;
    if (dev_alloc(h, &h->snap.rig_tr_global, 7 * static_cast<size_t>(h->n_imagesets))) return 1;
    if (dev_alloc(h, &h->snap.camera_tr_rig, 7 * static_cast<size_t>(h->n_cameras))) return 1;
    if (dev_alloc(h, &h->snap.intrinsics, h->intr_total)) return 1;
    if (dev_alloc(h, &h->d_snap_lp, h->n_obs)) return 1;
  }
  if (copy_state_dev(h, h->st[h->cur], h->d_last_projection, h->snap, h->d_snap_lp)) return 1;
  h->have_snapshot = true;
  return 0;
}

int b200ba_restore_state(b200ba_handle* h) {
  if (!h) return 1;
  if (!h->have_snapshot) {
    h->error = "no snapshot: call b200ba_snapshot_state first";
    return 2;
  }
  CUDA_TRY(h, cudaSetDevice(h->device));
  return copy_state_dev(h, h->snap, h->d_snap_lp, h->st[h->cur], h->d_last_projection);
}

int32_t b200ba_degrees_of_freedom(const b200ba_handle* h, const b200ba_options* opt) {
  if (!h || !opt) return -1;
  int n_intr = 0;
  for (int c = 0; c < h->n_cameras; ++c) n_intr += update_parameter_count(h->cams_host[c]);
  return 3 * h->n_points + 6 * h->n_imagesets + (h->n_cameras > 1 ? 6 * h->n_cameras : 0) + (opt->localize_only ? 0 : n_intr);
}

// OptimizeJointly (joint_optimization.cc:757-953): a loop of single LM iterations, each being
// LMOptimizer::OptimizeImpl with max_iteration_count = 1 (LV/lm_optimizer.h:628-991).
int b200ba_optimize(b200ba_handle* h, const b200ba_options* opt, b200ba_report* report) {
  if (!h) return 1;
  if (!report) {
    h->error = "report is NULL";
    return 2;
  }
  if (int rc = check_ready(h, opt)) return rc;
  memset(report, 0, sizeof(*report));
  memset(&h->timings, 0, sizeof(h->timings));
  cudaEvent_t ev_total_a = get_event(h), ev_total_b = get_event(h);
  cudaEventRecord(ev_total_a, h->stream);
  const Layout& L = h->L;
  const double huber = opt->huber_parameter;
  double lambda = 0, init_lambda = opt->init_lambda;
  // on-the-fly block processing (pose elimination with an *OnTheFly Schur mode) starts from a fixed
  // lambda when none is handed in (joint_optimization.cc:801-804)
  if (!opt->eliminate_points && init_lambda < 0 &&
      (opt->schur_mode == B200BA_SCHUR_DENSE_ONTHEFLY || opt->schur_mode == B200BA_SCHUR_SPARSE_ONTHEFLY))
    init_lambda = 0.0001f;
  report->final_lambda = init_lambda;  // what optimizer.lambda() holds if no iteration changes it
  double final_cost = -1;
  // n_valid / sum |r|^2 of the CURRENT state, taken from the passes the LM loop makes anyway (the
  // base pass of a build, or the trial pass of an accepted step) -- no extra pass for statistics
  double stat_valid = 0, stat_sumsq = 0;
  bool have_stats = false;
  if (opt->debug_verify_cost) {
    // LMOptimizer::VerifyCost twice (joint_optimization.cc:866-876, lm_optimizer.h:474-490)
    double c[2] = {0, 0};
    for (int rep2 = 0; rep2 < 2; ++rep2) {
      double with_jac = 0;
      for (int jac = 0; jac < 2; ++jac) {
        if (evaluate_state(h, h->cur, jac != 0, jac ? h->out : h->out_trial, huber, PH_TRIAL)) return 1;
        const ObsOut& oo = jac ? h->out : h->out_trial;
        launch_cost_reduce(h->n_obs, oo.cost, nullptr, oo.residual, h->d_partial, h->d_scal, h->stream);
        if (all_reduce(h, h->d_scal, 6)) return 1;
        CUDA_TRY(h, cudaMemcpyAsync(h->h_scal, h->d_scal, 6 * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
        if (sync_stream(h)) return 1;
        (jac ? with_jac : c[rep2]) = h->h_scal[3];
      }
      if (std::fabs(c[rep2] - with_jac) > 1e-3f)
        fprintf(stderr, "[b200ba] Cost differs when computed with or without Jacobians: %.12g vs %.12g\n", c[rep2], with_jac);
    }
    if (!(std::fabs(c[0] - c[1]) <= 1e-3f)) {
      h->error = "debug_verify_cost: two cost evaluations of the same state differ by more than 1e-3";
      return 5;
    }
  }
  const bool any_fixed = opt->debug_fix_points || opt->debug_fix_poses || opt->debug_fix_rig_poses || opt->debug_fix_intrinsics;
  for (int iteration = 0; iteration < opt->max_iteration_count; ++iteration) {
    double cost = 0, n_valid = 0;
    if (build_system(h, huber, &cost, &n_valid, any_fixed ? opt : nullptr)) return 1;
    h->timings.build_count += 1;
    stat_valid = h->h_scal[4];
    stat_sumsq = h->h_scal[5];
    have_stats = true;
    double last_cost = cost;
    if (iteration == 0) report->initial_cost = cost;
    if (cost == 0) {  // "Cost is zero, stopping." (lm_optimizer.h:755-760)
      final_cost = cost;
      break;
    }
    if (init_lambda >= 0) {
      lambda = init_lambda;
    } else {
      // lambda = init_lambda_factor * trace(H) / dof (lm_optimizer.h:766-781)
      lambda = opt->init_lambda_factor * h->trace_H / L.dof;
    }
    bool applied_update = false;
    int attempts = 0;
    for (int lm_iteration = 0; lm_iteration < opt->max_lm_attempts; ++lm_iteration) {
      ++attempts;
      h->timings.lm_attempts += 1;
      int spd = 1;
      if (solve_system(h, lambda, &spd)) return 1;
      if (!spd) {
        // A non-positive pivot: the analogue of the reference's NaN-update rejection
        // (lm_optimizer.h:905-913) -- increase the damping and retry.
        lambda = 2.f * lambda;
        if (opt->print_progress) fprintf(stderr, "[b200ba]   [%d, %d] factorisation failed, new lambda %g\n", iteration + 1, lm_iteration + 1, lambda);
        continue;
      }
      const int trial = 1 - h->cur;
      {
        ScopedPhase ph(h, PH_UPDATE);
        launch_update_state(h->pb, L, h->st[h->cur], h->st[trial], h->d_x, h->n_control_total, h->n_param_total, h->stream);
        h->timings.kernel_launches += 1;
      }
      if (evaluate_state(h, trial, false, h->out_trial, huber, PH_TRIAL)) return 1;
      {
        ScopedPhase ph(h, PH_TRIAL);
        launch_cost_reduce(h->n_obs, h->out_trial.cost, h->out.cost, h->out_trial.residual, h->d_partial, h->d_scal, h->stream);
        h->timings.kernel_launches += 2;
      }
      if (all_reduce(h, h->d_scal, 6)) return 1;
      CUDA_TRY(h, cudaMemcpyAsync(h->h_scal, h->d_scal, 6 * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
      if (sync_stream(h)) return 1;
      const double left = h->h_scal[0], right = h->h_scal[1], count = h->h_scal[2];
      const double trial_cost = h->h_scal[3];
      // CostIsSmallerThan (lm_optimizer.h:993-1011)
      if (count > 0 && left < right) {
        h->cur = trial;
        stat_valid = h->h_scal[4];
        stat_sumsq = h->h_scal[5];
        lambda = 0.5f * lambda;
        applied_update = true;
        report->num_iterations_performed += 1;
        last_cost = trial_cost;
        if (opt->print_progress) fprintf(stderr, "[b200ba] [%d] update accepted, cost %.12g\n", iteration + 1, trial_cost);
        break;
      } else {
        lambda = 2.f * lambda;
        if (opt->print_progress) fprintf(stderr, "[b200ba]   [%d, %d of %d] update rejected (bad cost: %.12g), new lambda: %g\n", iteration + 1, lm_iteration + 1, opt->max_lm_attempts, trial_cost, lambda);
      }
    }
    final_cost = last_cost;
    init_lambda = lambda;
    report->final_lambda = lambda;
    if (report->trace_len < B200BA_MAX_TRACE) {
      report->trace_cost[report->trace_len] = last_cost;
      report->trace_lambda[report->trace_len] = lambda;
      report->trace_attempts[report->trace_len] = attempts;
      report->trace_len++;
    }
    if (!applied_update) break;
    report->performed_an_iteration = 1;
    if (last_cost == 0) break;
  }
  report->final_cost = final_cost;
  if (!have_stats) {
    // max_iteration_count == 0: statistics of the unchanged state need one residual-only pass
    if (evaluate_state(h, h->cur, false, h->out_trial, huber, PH_TRIAL)) return 1;
    launch_cost_reduce(h->n_obs, h->out_trial.cost, nullptr, h->out_trial.residual, h->d_partial, h->d_scal, h->stream);
    h->timings.kernel_launches += 2;
    if (all_reduce(h, h->d_scal, 6)) return 1;
    CUDA_TRY(h, cudaMemcpyAsync(h->h_scal, h->d_scal, 6 * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    if (sync_stream(h)) return 1;
    stat_valid = h->h_scal[4];
    stat_sumsq = h->h_scal[5];
  }
  cudaEventRecord(ev_total_b, h->stream);
  if (sync_stream(h)) return 1;
  {
    float ms = 0;
    cudaEventElapsedTime(&ms, ev_total_a, ev_total_b);
    h->timings.total_ms = ms;
    h->event_pool.push_back(ev_total_a);
    h->event_pool.push_back(ev_total_b);
  }
  report->n_valid = static_cast<int64_t>(stat_valid);
  // with several ranks n_valid is global (all-reduced) while n_obs is this rank's shard
  report->n_invalid = (h->n_ranks > 1) ? -1 : (h->n_obs - report->n_valid);
  report->rmse = report->n_valid > 0 ? std::sqrt(stat_sumsq / stat_valid) : 0.0;
  report->cost_and_jacobian_evaluation_time = 1e-3 * (h->timings.jacobian_kernel_ms + h->timings.accumulate_ms + h->timings.trial_cost_ms);
  report->solve_time = 1e-3 * (h->timings.schur_ms + h->timings.factor_ms);
  return 0;
}

// RunBundleAdjustment (APP/calibration.cc:187-304) with the state resident on the device: single LM
// iterations with the lambda carried over, the re-orientation of every camera after each iteration
// (ChooseNiceCameraOrientation + camera_tr_rig update, calibration.cc:245-252) and the stopping
// criterion `cost >= last_cost - cost_reduction_threshold` (:298-300). Nothing but the scalars the
// stop rule needs crosses the PCIe bus between iterations.
int b200ba_run_bundle_adjustment(b200ba_handle* h, const b200ba_options* opt_in, int32_t max_iteration_count,
                                 double cost_reduction_threshold, b200ba_ba_report* out,
                                 int (*on_iteration)(void* user, int32_t iteration, double cost), void* user) {
  if (!h) return 1;
  if (!opt_in || !out) {
    h->error = "NULL argument";
    return 2;
  }
  memset(out, 0, sizeof(*out));
  b200ba_options opt = *opt_in;
  opt.max_iteration_count = 1;
  double lambda = opt_in->init_lambda;  // the reference starts from -1 (calibration.cc:203)
  double last_cost = INFINITY;
  b200ba_timings acc{};
  for (int iteration = 0; iteration < max_iteration_count; ++iteration) {
    opt.init_lambda = lambda;
    b200ba_report rep;
    if (int rc = b200ba_optimize(h, &opt, &rep)) return rc;
    lambda = rep.final_lambda;
    const double cost = rep.final_cost;
    if (iteration == 0) out->initial_cost = rep.initial_cost;
    out->final_cost = cost;
    out->final_lambda = lambda;
    out->rmse = rep.rmse;
    out->n_valid = rep.n_valid;
    out->n_invalid = rep.n_invalid;
    out->lm_attempts += rep.trace_len ? rep.trace_attempts[0] : 0;
    if (out->iterations < B200BA_MAX_TRACE) out->costs[out->iterations] = cost;
    out->iterations += 1;
    acc.total_ms += h->timings.total_ms;
    if (!opt.localize_only) {
      // beautify all camera orientations (calibration.cc:245-252)
      if (!h->d_rot && dev_alloc(h, &h->d_rot, 9 * static_cast<size_t>(h->n_cameras))) return 1;
      launch_nice_orientation(h->pb, h->st[h->cur], h->n_cameras, h->d_rot, h->stream);
      CUDA_TRY(h, cudaGetLastError());
    }
    if (on_iteration) {
      CUDA_TRY(h, cudaStreamSynchronize(h->stream));
      if (on_iteration(user, iteration, cost) != 0) break;  // the 'q' key of the reference (calibration.cc:293-295)
    }
    if (cost >= last_cost - cost_reduction_threshold) break;  // stopping criterion
    last_cost = cost;
  }
  CUDA_TRY(h, cudaStreamSynchronize(h->stream));
  out->device_ms = acc.total_ms;
  return 0;
}

int b200ba_optimize_host(b200ba_handle* h, b200ba_state* state, const b200ba_options* opt, b200ba_report* report) {
  if (int rc = b200ba_set_state(h, state)) return rc;
  if (int rc = b200ba_optimize(h, opt, report)) return rc;
  return b200ba_get_state(h, state);
}

int b200ba_evaluate(b200ba_handle* h, const b200ba_options* opt, int compute_jacobians, double* residuals,
                    double* costs, double* total_cost) {
  if (int rc = check_ready(h, opt)) return rc;
  const int64_t n = h->n_obs;
  if (evaluate_state(h, h->cur, compute_jacobians != 0, h->out, opt->huber_parameter, PH_JAC)) return 1;
  launch_cost_reduce(n, h->out.cost, nullptr, h->out.residual, h->d_partial, h->d_scal, h->stream);
  CUDA_TRY(h, cudaMemcpyAsync(h->h_scal, h->d_scal, 6 * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  if (sync_stream(h)) return 1;
  if (residuals) {
    std::vector<double> tmp(2 * n);
    CUDA_TRY(h, cudaMemcpy(tmp.data(), h->out.residual, 2 * n * sizeof(double), cudaMemcpyDeviceToHost));
    for (int64_t i = 0; i < n; ++i) {
      residuals[2 * static_cast<int64_t>(h->perm[i])] = tmp[i];
      residuals[2 * static_cast<int64_t>(h->perm[i]) + 1] = tmp[n + i];
    }
  }
  if (costs) {
    std::vector<double> tmp(n);
    CUDA_TRY(h, cudaMemcpy(tmp.data(), h->out.cost, n * sizeof(double), cudaMemcpyDeviceToHost));
    for (int64_t i = 0; i < n; ++i) costs[h->perm[i]] = tmp[i];
  }
  if (total_cost) *total_cost = h->h_scal[3];
  return 0;
}

int b200ba_get_jacobians(b200ba_handle* h, double* j_point, double* j_pose, double* j_rig, double* j_intr,
                         int32_t* intr_index, int32_t K) {
  if (!h || !h->have_layout) return 1;
  CUDA_TRY(h, cudaSetDevice(h->device));
  const Layout& L = h->L;
  const int64_t n = h->n_obs;
  std::vector<double> jac(2 * static_cast<size_t>(L.n_jcols) * n);
  std::vector<int32_t> cell(n);
  std::vector<uint8_t> has(n);
  if (h->out.compact) {
    // rebuild the expanded buffer from the compact records (state of the evaluation = current state)
    double* tmp = nullptr;
    CUDA_TRY(h, cudaMalloc(reinterpret_cast<void**>(&tmp), std::max<size_t>(1, jac.size()) * sizeof(double)));
    launch_expand_jacobian(h->pb, L, h->st[h->cur], h->out, tmp, h->stream);
    cudaError_t e = cudaStreamSynchronize(h->stream);
    if (e == cudaSuccess) e = cudaMemcpy(jac.data(), tmp, jac.size() * sizeof(double), cudaMemcpyDeviceToHost);
    cudaFree(tmp);
    CUDA_TRY(h, e);
  } else
    CUDA_TRY(h, cudaMemcpy(jac.data(), h->out.jac, jac.size() * sizeof(double), cudaMemcpyDeviceToHost));
  CUDA_TRY(h, cudaMemcpy(cell.data(), h->out.cell, n * sizeof(int32_t), cudaMemcpyDeviceToHost));
  CUDA_TRY(h, cudaMemcpy(has.data(), h->out.has_jac, n, cudaMemcpyDeviceToHost));
  std::vector<uint32_t> cams(n);
  CUDA_TRY(h, cudaMemcpy(cams.data(), h->d_obs_camera, n * sizeof(uint32_t), cudaMemcpyDeviceToHost));
  auto J = [&](int col, int r, int64_t i) { return jac[(2 * static_cast<size_t>(col) + r) * n + i]; };
  for (int64_t i = 0; i < n; ++i) {
    const int64_t o = h->perm[i];  // position i on the device holds the caller's observation o
    const bool v = has[i] == 1 || has[i] == 3;  // valid, or a late success of the straggler pass
    for (int r = 0; r < 2; ++r) {
      if (j_point) for (int j = 0; j < 3; ++j) j_point[(o * 2 + r) * 3 + j] = v ? J(L.jc_point + j, r, i) : 0.0;
      if (j_pose) for (int j = 0; j < 6; ++j) j_pose[(o * 2 + r) * 6 + j] = v ? J(L.jc_pose + j, r, i) : 0.0;
      if (j_rig) for (int j = 0; j < 6; ++j) j_rig[(o * 2 + r) * 6 + j] = (v && L.rig_in_state) ? J(L.jc_rig + j, r, i) : 0.0;
    }
    const CamDev& c = h->pb.cams[cams[i]];
    for (int k = 0; k < K; ++k) {
      const bool vk = v && !L.localize_only && k < c.K;
      if (j_intr) {
        j_intr[(o * 2 + 0) * K + k] = vk ? J(L.jc_intr + k, 0, i) : 0.0;
        j_intr[(o * 2 + 1) * K + k] = vk ? J(L.jc_intr + k, 1, i) : 0.0;
      }
      if (intr_index) {
        int idx = -1;
        if (vk) {
          int local;
          if (c.model_type == B200BA_MODEL_CENTRAL_GENERIC) {
            const int cp = k >> 1;
            local = 2 * (cell[i] + (cp & 3) + (cp >> 2) * c.gw) + (k & 1);
          } else if (c.model_type == B200BA_MODEL_NONCENTRAL_GENERIC) {
            const int cp = k / 5;
            local = 5 * (cell[i] + (cp & 3) + (cp >> 2) * c.gw) + (k - 5 * cp);
          } else {
            local = k;
          }
          idx = L.g_intr + c.upd_off + local;
        }
        intr_index[o * K + k] = idx;
      }
    }
  }
  return 0;
}

int b200ba_build_system(b200ba_handle* h, const b200ba_options* opt, int32_t n, double* H, double* b, double* cost) {
  if (int rc = check_ready(h, opt)) return rc;
  const Layout& L = h->L;
  if (n != L.dof) {
    h->error = "n does not equal the number of unknowns";
    return 2;
  }
  double c = 0, nv = 0;
  if (build_system(h, opt->huber_parameter, &c, &nv)) return 1;
  if (cost) *cost = c;
  std::vector<double> D(static_cast<size_t>(L.dsz) * L.nblocks), bp(L.nbd), B(static_cast<size_t>(L.nbd) * L.nd),
      C(static_cast<size_t>(L.nd) * L.nd), bd(L.nd);
  CUDA_TRY(h, cudaMemcpy(D.data(), h->sys.Dblk, D.size() * sizeof(double), cudaMemcpyDeviceToHost));
  CUDA_TRY(h, cudaMemcpy(bp.data(), h->sys.bp, bp.size() * sizeof(double), cudaMemcpyDeviceToHost));
  CUDA_TRY(h, cudaMemcpy(B.data(), h->sys.B, B.size() * sizeof(double), cudaMemcpyDeviceToHost));
  CUDA_TRY(h, cudaMemcpy(C.data(), h->sys.C, C.size() * sizeof(double), cudaMemcpyDeviceToHost));
  CUDA_TRY(h, cudaMemcpy(bd.data(), h->sys.bd, bd.size() * sizeof(double), cudaMemcpyDeviceToHost));
  std::fill(H, H + static_cast<size_t>(n) * n, 0.0);
  for (int p = 0; p < L.nblocks; ++p) {
    const double* d = &D[static_cast<size_t>(L.dsz) * p];
    const int o = L.bs * p;
    for (int a2 = 0; a2 < L.bs; ++a2)
      for (int b2 = a2; b2 < L.bs; ++b2)
        H[static_cast<size_t>(o + a2) * n + o + b2] = d[a2 * L.bs - (a2 * (a2 - 1)) / 2 + (b2 - a2)];
  }
  for (int i = 0; i < L.nbd; ++i)
    for (int k = 0; k < L.nd; ++k) H[static_cast<size_t>(i) * n + L.nbd + k] = B[static_cast<size_t>(i) * L.nd + k];
  for (int i = 0; i < L.nd; ++i)
    for (int k = i; k < L.nd; ++k) H[static_cast<size_t>(L.nbd + i) * n + L.nbd + k] = C[static_cast<size_t>(i) * L.nd + k];
  for (int i = 0; i < L.nbd; ++i) b[i] = bp[i];
  for (int i = 0; i < L.nd; ++i) b[L.nbd + i] = bd[i];
  return 0;
}

// Diagnostics / tests: evaluation budget of the main Jacobian pass (default 16); 1 forces every
// observation of a generic camera through the straggler pass.
B200BA_API void b200ba_debug_set_eval_budget(int budget) { set_main_eval_budget(budget); }

// Diagnostics: spline evaluations spent per observation by the last pass that wrote Jacobians
// (caller's observation order). Not part of the reference interface.
B200BA_API int b200ba_debug_eval_counts(b200ba_handle* h, uint16_t* counts) {
  if (!h || !counts || !h->have_layout) return 1;
  std::vector<uint16_t> tmp(h->n_obs);
  CUDA_TRY(h, cudaMemcpy(tmp.data(), h->out.evals, h->n_obs * sizeof(uint16_t), cudaMemcpyDeviceToHost));
  for (int64_t i = 0; i < h->n_obs; ++i) counts[h->perm[i]] = tmp[i];
  return 0;
}

int b200ba_get_timings(const b200ba_handle* h, b200ba_timings* t) {
  if (!h || !t) return 1;
  *t = h->timings;
  return 0;
}

// ---- stand-alone dense SPD solve on the in-tree kernels (tests / profiling) ---------------------
// A: n x n symmetric (either major order), host. Factorises with the blocked Cholesky of ba_dense.cu
// (block width nb, a multiple of 128) and solves A x = b. Returns 4 when a pivot is not positive.
int b200ba_dense_cholesky_solve(int device, int32_t n, int32_t nb, const double* A, const double* b, double* x,
                                double* factor_ms, double* solve_ms) {
  if (n < 1 || nb < 128 || nb % 128 != 0 || !A || !b || !x) {
    g_create_error = "b200ba_dense_cholesky_solve: bad argument";
    return 2;
  }
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    g_create_error = "no CUDA device available (this library has no CPU fallback)";
    return 3;
  }
  if (device >= 0) cudaSetDevice(device);
  set_gemm_sm_reserve(getenv("B200BA_PANEL_SMS") ? atoi(getenv("B200BA_PANEL_SMS")) : 8);
  DenseCtx d;
  dense_plan(&d, n, nb, 0, 1);
  int rc = 0;
  double* dA = nullptr;
  double* db = nullptr;
  int* dinfo = nullptr;
  cudaEvent_t e0 = nullptr, e1 = nullptr, e2 = nullptr;
  auto ok = [&](cudaError_t e) {
    if (e != cudaSuccess && rc == 0) {
      g_create_error = std::string("b200ba_dense_cholesky_solve: ") + cudaGetErrorString(e);
      rc = 1;
    }
    return e == cudaSuccess;
  };
  int lo = 0, hi = 0;
  cudaDeviceGetStreamPriorityRange(&lo, &hi);
  ok(cudaStreamCreateWithFlags(&d.s_main, cudaStreamNonBlocking));
  ok(cudaStreamCreateWithPriority(&d.s_panel, cudaStreamNonBlocking, hi));
  ok(cudaStreamCreateWithPriority(&d.s_aux, cudaStreamNonBlocking, hi));
  for (int i = 0; i < 2; ++i) {
    ok(cudaEventCreateWithFlags(&d.ev_ready[i], cudaEventDisableTiming));
    ok(cudaEventCreateWithFlags(&d.ev_main[i], cudaEventDisableTiming));
    ok(cudaEventCreateWithFlags(&d.ev_half2[i], cudaEventDisableTiming));
  }
  ok(cudaEventCreateWithFlags(&d.ev_misc, cudaEventDisableTiming));
  ok(cudaEventCreate(&e0));
  ok(cudaEventCreate(&e1));
  ok(cudaEventCreate(&e2));
  ok(cudaMalloc(reinterpret_cast<void**>(&dA), static_cast<size_t>(n) * n * sizeof(double)));
  ok(cudaMalloc(reinterpret_cast<void**>(&d.S), static_cast<size_t>(d.chunk) * sizeof(double)));
  ok(cudaMalloc(reinterpret_cast<void**>(&d.Lpack), static_cast<size_t>(d.panel_off[d.nblk]) * sizeof(double)));
  ok(cudaMalloc(reinterpret_cast<void**>(&d.tmp), static_cast<size_t>(n) * sizeof(double)));
  ok(cudaMalloc(reinterpret_cast<void**>(&d.d_panel_off), d.panel_off.size() * sizeof(int64_t)));
  ok(cudaMalloc(reinterpret_cast<void**>(&d.d_panel_h), d.panel_h.size() * sizeof(int)));
  ok(cudaMalloc(reinterpret_cast<void**>(&db), static_cast<size_t>(n) * sizeof(double)));
  ok(cudaMalloc(reinterpret_cast<void**>(&dinfo), sizeof(int)));
  d.info = dinfo;
  if (rc == 0) {
    ok(cudaMemcpy(dA, A, static_cast<size_t>(n) * n * sizeof(double), cudaMemcpyHostToDevice));
    ok(cudaMemcpy(db, b, static_cast<size_t>(n) * sizeof(double), cudaMemcpyHostToDevice));
    ok(cudaMemcpy(d.d_panel_off, d.panel_off.data(), d.panel_off.size() * sizeof(int64_t), cudaMemcpyHostToDevice));
    ok(cudaMemcpy(d.d_panel_h, d.panel_h.data(), d.panel_h.size() * sizeof(int), cudaMemcpyHostToDevice));
    ok(cudaMemset(d.S, 0, static_cast<size_t>(d.chunk) * sizeof(double)));
    ok(cudaMemset(dinfo, 0, sizeof(int)));
    ok(cudaMemcpy2D(d.S, d.map.ld * sizeof(double), dA, static_cast<size_t>(n) * sizeof(double),
                    static_cast<size_t>(n) * sizeof(double), n, cudaMemcpyDeviceToDevice));
    ok(cudaDeviceSynchronize());  // device-to-device copies are asynchronous; the work below runs on non-blocking streams
  }
  if (rc == 0) {
    cudaEventRecord(e0, d.s_main);
    if (dense_factor(&d)) rc = 1;
    cudaEventRecord(e1, d.s_main);
    if (rc == 0 && dense_solve(&d, db)) rc = 1;
    cudaEventRecord(e2, d.s_main);
    ok(cudaStreamSynchronize(d.s_main));
    ok(cudaStreamSynchronize(d.s_panel));
    if (rc == 1 && g_create_error.empty()) g_create_error = "b200ba_dense_cholesky_solve: kernel launch failed";
  }
  if (rc == 0) {
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    if (factor_ms) *factor_ms = ms;
    cudaEventElapsedTime(&ms, e1, e2);
    if (solve_ms) *solve_ms = ms;
    int info = 0;
    ok(cudaMemcpy(&info, dinfo, sizeof(int), cudaMemcpyDeviceToHost));
    ok(cudaMemcpy(x, db, static_cast<size_t>(n) * sizeof(double), cudaMemcpyDeviceToHost));
    if (rc == 0 && info != 0) {
      g_create_error = "b200ba_dense_cholesky_solve: the matrix is not positive definite";
      rc = 4;
    }
  }
  for (double* p : {dA, d.S, d.Lpack, d.tmp, db})
    if (p) cudaFree(p);
  if (d.d_panel_off) cudaFree(d.d_panel_off);
  if (d.d_panel_h) cudaFree(d.d_panel_h);
  if (dinfo) cudaFree(dinfo);
  for (int i = 0; i < 2; ++i) {
    if (d.ev_ready[i]) cudaEventDestroy(d.ev_ready[i]);
    if (d.ev_main[i]) cudaEventDestroy(d.ev_main[i]);
    if (d.ev_half2[i]) cudaEventDestroy(d.ev_half2[i]);
  }
  for (cudaEvent_t e : {d.ev_misc, e0, e1, e2})
    if (e) cudaEventDestroy(e);
  if (d.s_main) cudaStreamDestroy(d.s_main);
  if (d.s_panel) cudaStreamDestroy(d.s_panel);
  if (d.s_aux) cudaStreamDestroy(d.s_aux);
  return rc;
}

// ---- stand-alone Schur solve (known-answer tests) ---------------------------------------------
int b200ba_schur_solve(int device, int32_t bs, int32_t nb, int32_t nd, const double* D, const double* B,
                       const double* C, const double* b1, const double* b2, double* x) {
  if (bs < 1 || bs > 6 || nb < 0 || nd < 1 || !D || !B || !C || !b1 || !b2 || !x) {
    g_create_error = "b200ba_schur_solve: bad argument";
    return 2;
  }
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    g_create_error = "no CUDA device available (this library has no CPU fallback)";
    return 3;
  }
  if (device >= 0) cudaSetDevice(device);
  const int nbd = bs * nb;
  double *dD = nullptr, *dB = nullptr, *dC = nullptr, *db1 = nullptr, *db2 = nullptr, *dDinvB = nullptr, *dDinvb = nullptr,
         *dwork = nullptr;
  int *dipiv = nullptr, *dinfo = nullptr;
  cublasHandle_t cb = nullptr;
  cusolverDnHandle_t cs = nullptr;
  int rc = 0;
  auto ok = [&](cudaError_t e) {
    if (e != cudaSuccess && rc == 0) {
      g_create_error = cudaGetErrorString(e);
      rc = 1;
    }
  };
  ok(cudaMalloc(&dD, sizeof(double) * std::max(1, nb * bs * bs)));
  ok(cudaMalloc(&dB, sizeof(double) * std::max<size_t>(1, static_cast<size_t>(nbd) * nd)));
  ok(cudaMalloc(&dC, sizeof(double) * static_cast<size_t>(nd) * nd));
  ok(cudaMalloc(&db1, sizeof(double) * std::max(1, nbd)));
  ok(cudaMalloc(&db2, sizeof(double) * nd));
  ok(cudaMalloc(&dDinvB, sizeof(double) * std::max<size_t>(1, static_cast<size_t>(nbd) * nd)));
  ok(cudaMalloc(&dDinvb, sizeof(double) * std::max(1, nbd)));
  ok(cudaMalloc(&dipiv, sizeof(int) * nd));
  ok(cudaMalloc(&dinfo, sizeof(int)));
  if (rc == 0) {
    ok(cudaMemcpy(dD, D, sizeof(double) * nb * bs * bs, cudaMemcpyHostToDevice));
    ok(cudaMemcpy(dB, B, sizeof(double) * static_cast<size_t>(nbd) * nd, cudaMemcpyHostToDevice));
    ok(cudaMemcpy(dC, C, sizeof(double) * static_cast<size_t>(nd) * nd, cudaMemcpyHostToDevice));
    ok(cudaMemcpy(db1, b1, sizeof(double) * nbd, cudaMemcpyHostToDevice));
    ok(cudaMemcpy(db2, b2, sizeof(double) * nd, cudaMemcpyHostToDevice));
  }
  if (rc == 0 && (cublasCreate(&cb) != CUBLAS_STATUS_SUCCESS || cusolverDnCreate(&cs) != CUSOLVER_STATUS_SUCCESS)) {
    g_create_error = "cuBLAS / cuSOLVER initialisation failed";
    rc = 1;
  }
  if (rc == 0) {
    const double one = 1.0, m1 = -1.0;
    launch_symmetrize(nd, dC, 0);
    launch_generic_block_inverse(bs, nb, nd, dD, dB, db1, dDinvB, dDinvb, 0);
    if (nbd > 0) {
      // S = C - B^T D^-1 B ; sb = b2 - B^T D^-1 b1
      cublasDgemm(cb, CUBLAS_OP_N, CUBLAS_OP_T, nd, nd, nbd, &m1, dB, nd, dDinvB, nd, &one, dC, nd);
      cublasDgemv(cb, CUBLAS_OP_N, nd, nbd, &m1, dB, nd, dDinvb, 1, &one, db2, 1);
    }
    int lwork = 0;
    cusolverDnDgetrf_bufferSize(cs, nd, nd, dC, nd, &lwork);
    ok(cudaMalloc(&dwork, sizeof(double) * std::max(1, lwork)));
    cusolverDnDgetrf(cs, nd, nd, dC, nd, dwork, dipiv, dinfo);
    cusolverDnDgetrs(cs, CUBLAS_OP_N, nd, 1, dC, nd, dipiv, db2, nd, dinfo);
    if (nbd > 0) cublasDgemv(cb, CUBLAS_OP_T, nd, nbd, &m1, dDinvB, nd, db2, 1, &one, dDinvb, 1);
    ok(cudaDeviceSynchronize());
    if (nbd > 0) ok(cudaMemcpy(x, dDinvb, sizeof(double) * nbd, cudaMemcpyDeviceToHost));
    ok(cudaMemcpy(x + nbd, db2, sizeof(double) * nd, cudaMemcpyDeviceToHost));
  }
  cudaFree(dD); cudaFree(dB); cudaFree(dC); cudaFree(db1); cudaFree(db2); cudaFree(dDinvB); cudaFree(dDinvb);
  cudaFree(dwork); cudaFree(dipiv); cudaFree(dinfo);
  if (cb) cublasDestroy(cb);
  if (cs) cusolverDnDestroy(cs);
  return rc;
}

// ---- stand-alone model evaluation ------------------------------------------------------------------
static int model_io(int device, const b200ba_camera* cam, const double* intrinsics, int64_t n, const double* in,
                    int in_w, double* io_pixels, double* dirs, double* origins, int32_t* ok_out, bool project) {
  if (!cam || !intrinsics || n < 0) {
    g_create_error = "bad argument";
    return 2;
  }
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    g_create_error = "no CUDA device available (this library has no CPU fallback)";
    return 3;
  }
  if (device >= 0) cudaSetDevice(device);
  CamDev c{};
  fill_camdev(*cam, &c);
  const int64_t ni = intrinsics_size(*cam);
  double *dintr = nullptr, *din = nullptr, *dpx = nullptr, *ddir = nullptr, *dorg = nullptr;
  int32_t* dok = nullptr;
  int rc = 0;
  auto ok = [&](cudaError_t e) {
    if (e != cudaSuccess && rc == 0) {
      g_create_error = cudaGetErrorString(e);
      rc = 1;
    }
  };
  const size_t nn = std::max<int64_t>(1, n);
  ok(cudaMalloc(&dintr, sizeof(double) * ni));
  ok(cudaMalloc(&din, sizeof(double) * in_w * nn));
  ok(cudaMalloc(&dpx, sizeof(double) * 2 * nn));
  ok(cudaMalloc(&ddir, sizeof(double) * 3 * nn));
  ok(cudaMalloc(&dorg, sizeof(double) * 3 * nn));
  ok(cudaMalloc(&dok, sizeof(int32_t) * nn));
  if (rc == 0) {
    ok(cudaMemcpy(dintr, intrinsics, sizeof(double) * ni, cudaMemcpyHostToDevice));
    if (project) {
      ok(cudaMemcpy(din, in, sizeof(double) * 3 * n, cudaMemcpyHostToDevice));
      ok(cudaMemcpy(dpx, io_pixels, sizeof(double) * 2 * n, cudaMemcpyHostToDevice));
      launch_project_points(c, dintr, n, din, dpx, dok, 0);
    } else {
      ok(cudaMemcpy(dpx, in, sizeof(double) * 2 * n, cudaMemcpyHostToDevice));
      launch_unproject_pixels(c, dintr, n, dpx, ddir, dorg, dok, 0);
    }
    ok(cudaDeviceSynchronize());
    if (project) {
      ok(cudaMemcpy(io_pixels, dpx, sizeof(double) * 2 * n, cudaMemcpyDeviceToHost));
    } else {
      if (dirs) ok(cudaMemcpy(dirs, ddir, sizeof(double) * 3 * n, cudaMemcpyDeviceToHost));
      if (origins) ok(cudaMemcpy(origins, dorg, sizeof(double) * 3 * n, cudaMemcpyDeviceToHost));
    }
    if (ok_out) ok(cudaMemcpy(ok_out, dok, sizeof(int32_t) * n, cudaMemcpyDeviceToHost));
  }
  cudaFree(dintr); cudaFree(din); cudaFree(dpx); cudaFree(ddir); cudaFree(dorg); cudaFree(dok);
  return rc;
}

int b200ba_project(int device, const b200ba_camera* cam, const double* intrinsics, int64_t n,
                   const double* local_points, double* pixels, int32_t* ok) {
  return model_io(device, cam, intrinsics, n, local_points, 3, pixels, nullptr, nullptr, ok, true);
}
int b200ba_unproject(int device, const b200ba_camera* cam, const double* intrinsics, int64_t n, const double* pixels,
                     double* directions, double* origins, int32_t* ok) {
  return model_io(device, cam, intrinsics, n, pixels, 2, nullptr, directions, origins, ok, false);
}

// CentralGenericModel::FitToPixelDirectionsImpl (APP/models/central_generic.cc:551-568):
// LMOptimizer<double>::Optimize(max_iteration_count, max_lm_attempts = 10, init_lambda = -1,
// init_lambda_factor = 0.001f) over the direction grid (LV/lm_optimizer.h:628-991 without the
// Schur structure: SolveDensely). H (row-major upper == column-major lower) is built by
// dirfit_kernel<true>; every attempt factors a copy of H + lambda I with cuSOLVER.
int b200ba_fit_directions(int device, int32_t gw, int32_t gh, double* grid, int64_t n, const double* grid_points,
                          const double* directions, int32_t max_iteration_count, b200ba_fit_report* report) {
  if (gw < 4 || gh < 4 || !grid || n < 0 || (n > 0 && (!grid_points || !directions)) || !report) {
    g_create_error = "b200ba_fit_directions: bad argument";
    return 2;
  }
  for (int64_t i = 0; i < n; ++i) {
    const double gx = grid_points[2 * i], gy = grid_points[2 * i + 1];
    if (!(gx >= 1.0 && gx < gw - 2 && gy >= 1.0 && gy < gh - 2)) {
      g_create_error = "b200ba_fit_directions: a grid point has no complete 4x4 support";
      return 2;
    }
  }
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    g_create_error = "no CUDA device available (this library has no CPU fallback)";
    return 3;
  }
  if (device >= 0) cudaSetDevice(device);
  memset(report, 0, sizeof(*report));
  const int G = gw * gh, dof = 2 * G;
  const size_t nn = static_cast<size_t>(std::max<int64_t>(1, n));
  double *dgrid = nullptr, *dtrial = nullptr, *dtan = nullptr, *dgp = nullptr, *ddir = nullptr, *dH = nullptr,
         *dS = nullptr, *db = nullptr, *dx = nullptr, *dcost = nullptr, *dsum = nullptr, *dwork = nullptr;
  int* dinfo = nullptr;
  cublasHandle_t cb = nullptr;
  cusolverDnHandle_t cs = nullptr;
  int rc = 0;
  auto ok = [&](cudaError_t e) {
    if (e != cudaSuccess && rc == 0) {
      g_create_error = cudaGetErrorString(e);
      rc = 1;
    }
  };
  const size_t hbytes = sizeof(double) * static_cast<size_t>(dof) * dof;
  ok(cudaMalloc(&dgrid, sizeof(double) * 3 * G));
  ok(cudaMalloc(&dtrial, sizeof(double) * 3 * G));
  ok(cudaMalloc(&dtan, sizeof(double) * 6 * G));
  ok(cudaMalloc(&dgp, sizeof(double) * 2 * nn));
  ok(cudaMalloc(&ddir, sizeof(double) * 3 * nn));
  ok(cudaMalloc(&dH, hbytes));
  ok(cudaMalloc(&dS, hbytes));
  ok(cudaMalloc(&db, sizeof(double) * dof));
  ok(cudaMalloc(&dx, sizeof(double) * dof));
  ok(cudaMalloc(&dcost, sizeof(double) * nn));
  ok(cudaMalloc(&dsum, sizeof(double) * 2));
  ok(cudaMalloc(&dinfo, sizeof(int) * 2));
  if (rc == 0 && (cublasCreate(&cb) != CUBLAS_STATUS_SUCCESS || cusolverDnCreate(&cs) != CUSOLVER_STATUS_SUCCESS)) {
    g_create_error = "cuBLAS / cuSOLVER initialisation failed";
    rc = 1;
  }
  int lwork = 0;
  if (rc == 0) {
    if (cusolverDnDpotrf_bufferSize(cs, CUBLAS_FILL_MODE_LOWER, dof, dS, dof, &lwork) != CUSOLVER_STATUS_SUCCESS) {
      g_create_error = "cusolverDnDpotrf_bufferSize failed";
      rc = 1;
    }
    ok(cudaMalloc(&dwork, sizeof(double) * std::max(1, lwork)));
  }
  if (rc == 0) {
    ok(cudaMemcpy(dgrid, grid, sizeof(double) * 3 * G, cudaMemcpyHostToDevice));
    if (n > 0) {
      ok(cudaMemcpy(dgp, grid_points, sizeof(double) * 2 * n, cudaMemcpyHostToDevice));
      ok(cudaMemcpy(ddir, directions, sizeof(double) * 3 * n, cudaMemcpyHostToDevice));
    }
  }
  double lambda = 0, last_cost = 0;
  const double init_lambda_factor = static_cast<double>(0.001f);  // the call site passes a float literal
  for (int iteration = 0; rc == 0 && iteration < max_iteration_count; ++iteration) {
    launch_dirfit_tangents(G, dgrid, dtan, 0);
    ok(cudaMemsetAsync(dH, 0, hbytes, 0));
    ok(cudaMemsetAsync(db, 0, sizeof(double) * dof, 0));
    launch_dirfit(true, gw, n, dgp, ddir, dgrid, dtan, dH, db, dof, dcost, dsum, 0);
    ok(cudaMemcpy(&last_cost, dsum, sizeof(double), cudaMemcpyDeviceToHost));
    if (rc) break;
    if (iteration == 0) report->initial_cost = last_cost;
    if (last_cost == 0) break;
    if (iteration == 0) {
      double trace = 0;  // the diagonal of J^T J is non-negative: sum |h_ii| = trace
      if (cublasDasum(cb, dof, dH, dof + 1, &trace) != CUBLAS_STATUS_SUCCESS) {
        g_create_error = "cublasDasum failed";
        rc = 1;
        break;
      }
      lambda = init_lambda_factor * trace / dof;
    }
    bool applied = false;
    for (int attempt = 0; rc == 0 && attempt < 10; ++attempt) {
      report->lm_attempts++;
      ok(cudaMemcpyAsync(dS, dH, hbytes, cudaMemcpyDeviceToDevice, 0));
      launch_add_diagonal(dof, dS, dof, lambda, 0);
      ok(cudaMemcpyAsync(dx, db, sizeof(double) * dof, cudaMemcpyDeviceToDevice, 0));
      int info[2] = {0, 0};
      if (cusolverDnDpotrf(cs, CUBLAS_FILL_MODE_LOWER, dof, dS, dof, dwork, lwork, dinfo) != CUSOLVER_STATUS_SUCCESS ||
          cusolverDnDpotrs(cs, CUBLAS_FILL_MODE_LOWER, dof, 1, dS, dof, dx, dof, dinfo + 1) != CUSOLVER_STATUS_SUCCESS) {
        g_create_error = "cuSOLVER potrf / potrs failed";
        rc = 1;
        break;
      }
      ok(cudaMemcpy(info, dinfo, sizeof(info), cudaMemcpyDeviceToHost));
      if (rc) break;
      if (info[0] != 0) {  // not positive definite: the reference's NaN-update branch
        lambda = 2.f * lambda;
        continue;
      }
      launch_dirfit_update(G, dgrid, dx, dtrial, 0);
      launch_dirfit(false, gw, n, dgp, ddir, dtrial, dtan, nullptr, nullptr, dof, dcost, dsum + 1, 0);
      double test_cost = 0;
      ok(cudaMemcpy(&test_cost, dsum + 1, sizeof(double), cudaMemcpyDeviceToHost));
      if (rc) break;
      if (test_cost < last_cost) {  // CostIsSmallerThan: every residual is valid in both states
        std::swap(dgrid, dtrial);
        lambda = 0.5f * lambda;
        applied = true;
        report->num_iterations_performed += 1;
        last_cost = test_cost;
        break;
      }
      lambda = 2.f * lambda;
    }
    if (!applied || last_cost == 0) break;
  }
  if (rc == 0) {
    ok(cudaGetLastError());
    ok(cudaMemcpy(grid, dgrid, sizeof(double) * 3 * G, cudaMemcpyDeviceToHost));
    report->final_cost = last_cost;
    report->final_lambda = lambda;
  }
  cudaFree(dgrid); cudaFree(dtrial); cudaFree(dtan); cudaFree(dgp); cudaFree(ddir); cudaFree(dH); cudaFree(dS);
  cudaFree(db); cudaFree(dx); cudaFree(dcost); cudaFree(dsum); cudaFree(dinfo); cudaFree(dwork);
  if (cb) cublasDestroy(cb);
  if (cs) cusolverDnDestroy(cs);
  return rc;
}

// ---- multi-GPU -------------------------------------------------------------------------------------
int b200ba_nccl_unique_id(uint8_t id[B200BA_NCCL_UNIQUE_ID_BYTES]) {
  std::string err;
  if (!load_nccl(&err)) {
    g_create_error = err;
    return 1;
  }
  return g_nccl.GetUniqueId(id) == 0 ? 0 : 1;
}

int b200ba_comm_init(b200ba_handle* h, const uint8_t id[B200BA_NCCL_UNIQUE_ID_BYTES], int rank, int n_ranks) {
  if (!h) return 1;
  if (n_ranks <= 1) {
    h->rank = 0;
    h->n_ranks = 1;
    return 0;
  }
  std::string err;
  if (!load_nccl(&err)) {
    h->error = err;
    return 1;
  }
  CUDA_TRY(h, cudaSetDevice(h->device));
  NcclUniqueId uid;
  memcpy(uid.internal, id, 128);
  int rc = g_nccl.CommInitRank(&h->comm, n_ranks, uid, rank);
  if (rc != 0) {
    h->error = std::string("ncclCommInitRank: ") + (g_nccl.GetErrorString ? g_nccl.GetErrorString(rc) : "error");
    return 1;
  }
  h->rank = rank;
  h->n_ranks = n_ranks;
  if (plan_dense(h)) return 1;
  // every rank must derive the same groups of Schur blocks: reduce the centroid sums
  if (h->L.eliminate_points && h->L.nblocks > 0 && !h->grp_sums.empty()) {
    if (reduce_group_sums(h)) return 1;
    if (build_groups(h, false)) return 1;
  }
  return 0;
}

}  // extern "C"
