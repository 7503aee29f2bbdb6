// b200ba_pipeline.hpp -- C++ host logic of the callers either side of the hot path:
// the outlier deletion between bundle-adjustment rounds, the metric rescaling and the pyramid resampling of the generic
// models, over the containers of
// b200ba_shim.hpp. (RunBundleAdjustment itself -- 8f-2 -- is in b200ba_shim.hpp and runs device-resident in the
// library.) The Python mirror is camera_calibration_b200/pipeline.py.
//
// Header-only. Every numerical step on observations (the re-projection of all features of a camera) runs in
// libb200ba.so through b200ba_project; there is no CPU fallback.
#pragma once

#include <algorithm>
#include <cmath>
#include <functional>
#include <map>

#include "b200ba_shim.hpp"

namespace b200ba_shim {

// (model, local points [3 n]) -> pixels [2 n], ok [n]. The default runs CameraModel::Project (start at the centre of
// the calibrated area, models/central_generic.cc:398-422) for all points in one b200ba_project call.
using ProjectMany = std::function<void(CameraModel&, const std::vector<double>&, std::vector<double>*, std::vector<int32_t>*)>;

inline void ProjectManyOnDevice(CameraModel& model, const std::vector<double>& local_points, std::vector<double>* pixels,
                                std::vector<int32_t>* ok) {
  const int64_t n = static_cast<int64_t>(local_points.size() / 3);
  b200ba_camera c{};
  c.model_type = static_cast<int32_t>(model.type());
  c.width = model.width();
  c.height = model.height();
  c.calibration_min_x = model.calibration_min_x();
  c.calibration_min_y = model.calibration_min_y();
  c.calibration_max_x = model.calibration_max_x();
  c.calibration_max_y = model.calibration_max_y();
  int rx = 0, ry = 0;
  if (model.GetGridResolution(&rx, &ry)) { c.grid_width = rx; c.grid_height = ry; }
  const Vec2d centre = model.CenterOfCalibratedArea();
  pixels->resize(2 * n);
  for (int64_t i = 0; i < n; ++i) { (*pixels)[2 * i] = centre.x; (*pixels)[2 * i + 1] = centre.y; }
  ok->assign(n, 0);
  if (n == 0) return;
  if (b200ba_project(-1, &c, model.flat_intrinsics().data(), n, local_points.data(), pixels->data(), ok->data()) != 0)
    throw std::runtime_error(std::string("b200ba_project: ") + b200ba_last_error(nullptr));
}

// calibration.cc:62-184 -- the quartile rule between BA rounds: re-project every feature of one camera, take the
// first and third quartile q1, q3 of the error magnitudes and erase the features that fail to project or whose
// error exceeds q3 + outlier_removal_factor (q3 - q1). Imagesets left with fewer than three features of this
// camera are marked unused. Returns the number of removed features.
inline int DeleteOutlierFeatures(int camera_index, Dataset* dataset, BAState* state, float outlier_removal_factor,
                                 const ProjectMany& project_many = ProjectManyOnDevice) {
  CameraModel& model = *state->intrinsics[camera_index];
  struct Span { int imageset; size_t first, last; };
  std::vector<Span> spans;
  std::vector<double> local_points;
  std::vector<float> observed;
  for (int i = 0; i < dataset->ImagesetCount(); ++i) {
    if (!state->image_used[i]) continue;
    const std::vector<PointFeature>& features = dataset->GetImageset(i)->FeaturesOfCamera(camera_index);
    const SE3d image_tr_global = state->image_tr_global(camera_index, i);
    const size_t first = observed.size() / 2;
    for (const PointFeature& f : features) {
      const Vec3d p = apply(image_tr_global, state->points[f.index]);
      local_points.insert(local_points.end(), {p.x, p.y, p.z});
      observed.push_back(f.xy.x);
      observed.push_back(f.xy.y);
    }
    spans.push_back(Span{i, first, observed.size() / 2});
  }
  const size_t n = observed.size() / 2;
  if (n == 0) return 0;
  std::vector<double> pixels;
  std::vector<int32_t> ok;
  project_many(model, local_points, &pixels, &ok);
  if (pixels.size() != 2 * n || ok.size() != n) throw std::runtime_error("DeleteOutlierFeatures: projector returned a wrong size");
  std::vector<double> error(n), sorted;
  for (size_t k = 0; k < n; ++k) {
    const double dx = pixels[2 * k] - static_cast<double>(observed[2 * k]), dy = pixels[2 * k + 1] - static_cast<double>(observed[2 * k + 1]);
    error[k] = std::sqrt(dx * dx + dy * dy);
    if (ok[k]) sorted.push_back(error[k]);
  }
  if (sorted.size() < 8) return 0;  // too few to detect outliers reliably (calibration.cc:97-100)
  std::sort(sorted.begin(), sorted.end());
  // quartile positions in float arithmetic, truncated (calibration.cc:103-104)
  const double first_quartile = sorted[static_cast<size_t>(0.25f * static_cast<float>(sorted.size()) + 0.5f)];
  const double third_quartile = sorted[static_cast<size_t>(0.75f * static_cast<float>(sorted.size()) + 0.5f)];
  const double threshold = third_quartile + static_cast<double>(outlier_removal_factor) * (third_quartile - first_quartile);
  int removed = 0;
  for (const Span& span : spans) {
    std::vector<PointFeature>& features = dataset->GetImageset(span.imageset)->FeaturesOfCamera(camera_index);
    size_t kept = 0;
    for (size_t k = span.first; k < span.last; ++k) {
      if (!ok[k] || error[k] > threshold) {
        ++removed;
        continue;
      }
      features[kept++] = features[k - span.first];
    }
    features.resize(kept);
    if (kept < 3) state->image_used[span.imageset] = false;
  }
  return removed;
}

// calibration.cc:307-370 -- geometric-mean ratio of the known pattern cell length to the optimised distance of
// neighbouring corners (right and down neighbours), applied with BAState::ScaleState. Returns the factor; throws
// when no neighbouring pair with known geometry exists (the reference divides by zero there).
inline double ScaleToMetric(const Dataset& dataset, BAState* state) {
  double log_sum = 0;
  long long count = 0;
  for (const KnownGeometry& geometry : dataset.known_geometries()) {
    std::map<std::pair<int, int>, int> position_to_index;
    for (const auto& item : geometry.feature_id_to_position) {
      auto it = state->feature_id_to_points_index.find(item.first);
      if (it != state->feature_id_to_points_index.end()) position_to_index[item.second] = it->second;
    }
    if (position_to_index.empty()) continue;
    for (const auto& item : geometry.feature_id_to_position) {
      auto self = position_to_index.find(item.second);
      if (self == position_to_index.end()) continue;
      const std::pair<int, int> neighbours[2] = {{item.second.first + 1, item.second.second}, {item.second.first, item.second.second + 1}};
      for (const auto& position : neighbours) {
        auto other = position_to_index.find(position);
        if (other == position_to_index.end()) continue;
        const Vec3d &a = state->points[self->second], &b = state->points[other->second];
        const double actual = std::sqrt((a.x - b.x) * (a.x - b.x) + (a.y - b.y) * (a.y - b.y) + (a.z - b.z) * (a.z - b.z));
        log_sum += std::log(static_cast<double>(geometry.cell_length_in_meters) / actual);
        ++count;
      }
    }
  }
  if (count == 0) throw std::runtime_error("ScaleToMetric: no neighbouring corners with known geometry");
  const double factor = std::exp(log_sum / static_cast<double>(count));
  state->ScaleState(factor);
  return factor;
}

// ---- pyramid resampling ------------------------------------------------------------------
// (model, pixels [2 n]) -> directions [3 n], ok [n]: CameraModel::Unproject for all pixels in one b200ba_unproject call
using UnprojectMany = std::function<void(CameraModel&, const std::vector<double>&, std::vector<double>*, std::vector<int32_t>*)>;
// (model, grid points [2 n], directions [3 n], max_iteration_count): CentralGenericModel::FitToPixelDirectionsImpl
using FitGridPoints = std::function<void(CentralGenericModel&, const std::vector<double>&, const std::vector<double>&, int)>;

inline void UnprojectManyOnDevice(CameraModel& model, const std::vector<double>& pixels, std::vector<double>* directions,
                                  std::vector<int32_t>* ok) {
  const int64_t n = static_cast<int64_t>(pixels.size() / 2);
  b200ba_camera c{};
  c.model_type = static_cast<int32_t>(model.type());
  c.width = model.width();
  c.height = model.height();
  c.calibration_min_x = model.calibration_min_x();
  c.calibration_min_y = model.calibration_min_y();
  c.calibration_max_x = model.calibration_max_x();
  c.calibration_max_y = model.calibration_max_y();
  int rx = 0, ry = 0;
  if (model.GetGridResolution(&rx, &ry)) { c.grid_width = rx; c.grid_height = ry; }
  directions->assign(3 * n, 0.0);
  ok->assign(n, 0);
  if (n == 0) return;
  std::vector<double> origins(3 * n, 0.0);
  if (b200ba_unproject(-1, &c, model.flat_intrinsics().data(), n, pixels.data(), directions->data(), origins.data(), ok->data()) != 0)
    throw std::runtime_error(std::string("b200ba_unproject: ") + b200ba_last_error(nullptr));
}
inline void FitGridPointsOnDevice(CentralGenericModel& model, const std::vector<double>& grid_points,
                                  const std::vector<double>& directions, int max_iteration_count) {
  b200ba_fit_report rep;
  if (b200ba_fit_directions(-1, model.gw, model.gh, model.grid.data(), static_cast<int64_t>(grid_points.size() / 2),
                            grid_points.data(), directions.data(), max_iteration_count, &rep) != 0)
    throw std::runtime_error(std::string("b200ba_fit_directions: ") + b200ba_last_error(nullptr));
}

// central_grid.h:138-148 -- FLOAT arithmetic, like the reference
inline Vec2d GridPointToPixelCornerConv(float x, float y, int min_x, int min_y, int max_x, int max_y, int grid_width, int grid_height) {
  const float px = static_cast<float>(min_x) + ((x - 1.f) / (static_cast<float>(grid_width) - 3.f)) * static_cast<float>(max_x + 1 - min_x);
  const float py = static_cast<float>(min_y) + ((y - 1.f) / (static_cast<float>(grid_height) - 3.f)) * static_cast<float>(max_y + 1 - min_y);
  return Vec2d{px, py};
}

// calibration.cc:531-540: integer division, + 0.5f, + the exterior cells, truncated
inline void ComputeGridResolution(int calibration_area_width, int calibration_area_height, int exterior_cells_per_side,
                                  int approx_pixels_per_cell, int* resolution_x, int* resolution_y) {
  *resolution_x = static_cast<int>(static_cast<float>(calibration_area_width / approx_pixels_per_cell) + 0.5f + static_cast<float>(2 * exterior_cells_per_side));
  *resolution_y = static_cast<int>(static_cast<float>(calibration_area_height / approx_pixels_per_cell) + 0.5f + static_cast<float>(2 * exterior_cells_per_side));
}
// calibration.cc:566-569
inline void CalcGridResolutionForLevel(int pyramid_level, int full_resolution_x, int full_resolution_y, int* resolution_x,
                                       int* resolution_y) {
  const double factor = std::pow(1.333, -pyramid_level);
  *resolution_x = static_cast<int>(full_resolution_x * factor + 0.5f);
  *resolution_y = static_cast<int>(full_resolution_y * factor + 0.5f);
}
// calibration.cc:615-641: bounding rectangle of the truncated feature positions of one camera (used imagesets)
inline void ComputeIntegerBoundingRectForFeatures(const Dataset& dataset, int camera_index, const std::vector<bool>& image_used,
                                                  int* min_x, int* min_y, int* max_x, int* max_y) {
  *min_x = *min_y = 2147483647;
  *max_x = *max_y = 0;
  for (int i = 0; i < dataset.ImagesetCount(); ++i) {
    if (!image_used[i]) continue;
    for (const PointFeature& f : dataset.GetImageset(i)->FeaturesOfCamera(camera_index)) {
      const int x = static_cast<int>(f.xy.x), y = static_cast<int>(f.xy.y);
      *min_x = std::min(*min_x, x);
      *min_y = std::min(*min_y, y);
      *max_x = std::max(*max_x, x);
      *max_y = std::max(*max_y, y);
    }
  }
}

namespace detail {
inline bool is_nan3(const double* v) { return v[0] != v[0] || v[1] != v[1] || v[2] != v[2]; }
// libvis Image::InterpolateBilinear for 3-vectors (libvis/image.h:152-176): truncation, FLOAT weights
inline void interpolate_bilinear3(const double* image, int width, double x, double y, double* out) {
  const int ix = static_cast<int>(x), iy = static_cast<int>(y);
  const float fx = static_cast<float>(x - ix), fy = static_cast<float>(y - iy);
  const float fx_inv = 1.f - fx, fy_inv = 1.f - fy;
  const double w00 = fx_inv * fy_inv, w10 = fx * fy_inv, w01 = fx_inv * fy, w11 = fx * fy;
  for (int k = 0; k < 3; ++k)
    out[k] = w00 * image[3 * (static_cast<size_t>(iy) * width + ix) + k] + w10 * image[3 * (static_cast<size_t>(iy) * width + ix + 1) + k] +
             w01 * image[3 * (static_cast<size_t>(iy + 1) * width + ix) + k] + w11 * image[3 * (static_cast<size_t>(iy + 1) * width + ix + 1) + k];
}
}  // namespace detail

// central_generic.cc:267-422 -- dense [dh * dw * 3]: one direction per pixel, NaN where the source model is undefined.
// Initialises every control point from the closest valid pixel (search radius < 5), extrapolates the rest linearly from
// their neighbours (in place: the sweep order matters, like the reference), then fits the grid to the sub-sampled
// dense directions (on the device unless `fit` is given).
inline bool FitToDenseModel(CentralGenericModel* model, const std::vector<double>& dense, int dw, int dh, int subsample_step,
                            int max_iteration_count, const FitGridPoints& fit = FitGridPointsOnDevice) {
  const int gw = model->gw, gh = model->gh;
  const double scale_x = dw / static_cast<double>(model->width()), scale_y = dh / static_cast<double>(model->height());
  auto valid = [&](int x, int y) { return dense[3 * (static_cast<size_t>(y) * dw + x)] == dense[3 * (static_cast<size_t>(y) * dw + x)]; };
  const double nan = std::nan("");
  std::vector<double> grid(3 * static_cast<size_t>(gw) * gh, nan);
  auto take = [&](int gx, int gy, int x, int y) {
    for (int k = 0; k < 3; ++k) grid[3 * (static_cast<size_t>(gy) * gw + gx) + k] = dense[3 * (static_cast<size_t>(y) * dw + x) + k];
  };
  bool have_nan = false;
  for (int gy = 0; gy < gh; ++gy)
    for (int gx = 0; gx < gw; ++gx) {
      const Vec2d p = GridPointToPixelCornerConv(static_cast<float>(gx), static_cast<float>(gy), model->calibration_min_x(), model->calibration_min_y(),
                                                 model->calibration_max_x(), model->calibration_max_y(), gw, gh);
      const int cx = static_cast<int>(scale_x * p.x), cy = static_cast<int>(scale_y * p.y);
      if (cx < 0 || cy < 0 || cx >= dw || cy >= dh) { have_nan = true; continue; }
      if (valid(cx, cy)) { take(gx, gy, cx, cy); continue; }
      bool found = false;
      for (int radius = 1; radius < 5 && !found; ++radius) {
        const int min_x = cx - radius, min_y = cy - radius, max_x = cx + radius, max_y = cy + radius;
        for (int x = std::max(0, min_x); x <= std::min(dw - 1, max_x) && !found; ++x) {  // top and bottom
          if (min_y >= 0 && valid(x, min_y)) { take(gx, gy, x, min_y); found = true; }
          else if (max_y < dh && valid(x, max_y)) { take(gx, gy, x, max_y); found = true; }
        }
        for (int y = std::max(0, min_y); y <= std::min(dh - 1, max_y) && !found; ++y) {  // left and right
          if (min_x >= 0 && valid(min_x, y)) { take(gx, gy, min_x, y); found = true; }
          else if (max_x < dw && valid(max_x, y)) { take(gx, gy, max_x, y); found = true; }
        }
      }
      if (!found) have_nan = true;
    }
  for (int iteration = 0; have_nan && iteration < dw + dh; ++iteration) {
    have_nan = false;
    for (int gy = 0; gy < gh; ++gy)
      for (int gx = 0; gx < gw; ++gx) {
        double* g = &grid[3 * (static_cast<size_t>(gy) * gw + gx)];
        if (!detail::is_nan3(g)) continue;
        double total[3] = {0, 0, 0};
        int count = 0;
        const int steps[4][2] = {{0, 1}, {0, -1}, {1, 0}, {-1, 0}};
        for (const auto& st : steps) {
          const int nx1 = gx + st[0], ny1 = gy + st[1], nx2 = gx + 2 * st[0], ny2 = gy + 2 * st[1];
          if (nx2 < 0 || ny2 < 0 || nx2 >= gw || ny2 >= gh) continue;
          const double* v1 = &grid[3 * (static_cast<size_t>(ny1) * gw + nx1)];
          const double* v2 = &grid[3 * (static_cast<size_t>(ny2) * gw + nx2)];
          if (detail::is_nan3(v1) || detail::is_nan3(v2)) continue;
          for (int k = 0; k < 3; ++k) total[k] += v1[k] + (v1[k] - v2[k]);
          ++count;
        }
        if (count > 0) {
          const double norm = std::sqrt(total[0] * total[0] + total[1] * total[1] + total[2] * total[2]);
          for (int k = 0; k < 3; ++k) g[k] = total[k] / norm;
        } else {
          have_nan = true;
        }
      }
  }
  if (have_nan) return false;
  model->grid = grid;
  // samples: every subsample_step-th pixel of the calibrated area with a valid direction
  const double model_to_camera_x = static_cast<double>(model->width()) / dw, model_to_camera_y = static_cast<double>(model->height()) / dh;
  std::vector<double> grid_points, directions;
  for (int y = model->calibration_min_y(); y <= model->calibration_max_y(); y += subsample_step)
    for (int x = model->calibration_min_x(); x <= model->calibration_max_x(); x += subsample_step) {
      const int dx = static_cast<int>(scale_x * x), dy = static_cast<int>(scale_y * y);
      if (!valid(dx, dy)) continue;
      const Vec2d gp = model->PixelCornerConvToGridPoint(model_to_camera_x * (dx + 0.5f), model_to_camera_y * (dy + 0.5f));
      grid_points.push_back(gp.x);
      grid_points.push_back(gp.y);
      for (int k = 0; k < 3; ++k) directions.push_back(dense[3 * (static_cast<size_t>(dy) * dw + dx) + k]);
    }
  fit(*model, grid_points, directions, max_iteration_count);
  return true;
}

// calibration.cc:373-522 for the generic target models (the parametric targets are outside this path). Returns the new
// model, or an empty pointer where the reference returns false. camera_tr_rig is untouched for these targets.
//   * NoncentralGeneric -> NoncentralGeneric: both grids re-sampled bilinearly (:386-424), host only;
//   * central source: a dense direction image of the old model (one Unproject per pixel centre, on the device) is fitted
//     by a CentralGenericModel of the target resolution (FitToDenseModel(dense, step, 3), at most 300 x 300 samples),
//     optionally wrapped into a NoncentralGenericModel with zero origins.
inline std::shared_ptr<CameraModel> ResampleModel(CameraModel& model_to_optimize, int calibration_min_x, int calibration_min_y,
                                                  int calibration_max_x, int calibration_max_y, CameraModel::Type model_type,
                                                  int target_resolution_x, int target_resolution_y,
                                                  const UnprojectMany& unproject_many = UnprojectManyOnDevice,
                                                  const FitGridPoints& fit = FitGridPointsOnDevice) {
  using T = CameraModel::Type;
  const int w = model_to_optimize.width(), h = model_to_optimize.height();
  if (model_to_optimize.type() == T::NoncentralGeneric && model_type == T::NoncentralGeneric) {
    auto& old = static_cast<NoncentralGenericModel&>(model_to_optimize);
    const size_t old_n = static_cast<size_t>(old.gw) * old.gh;
    std::shared_ptr<NoncentralGenericModel> fresh(new NoncentralGenericModel(target_resolution_x, target_resolution_y, calibration_min_x,
                                                                             calibration_min_y, calibration_max_x, calibration_max_y, w, h));
    const size_t new_n = static_cast<size_t>(target_resolution_x) * target_resolution_y;
    for (int y = 0; y < target_resolution_y; ++y)
      for (int x = 0; x < target_resolution_x; ++x) {
        const Vec2d pixel = GridPointToPixelCornerConv(static_cast<float>(x), static_cast<float>(y), calibration_min_x, calibration_min_y,
                                                       calibration_max_x, calibration_max_y, target_resolution_x, target_resolution_y);
        // noncentral_generic.h:167-171 (double arithmetic with the float constant (grid - 3.f))
        double gx = 1.0 + static_cast<double>(old.gw - 3.f) * (pixel.x - old.calibration_min_x()) / (old.calibration_max_x() + 1 - old.calibration_min_x());
        double gy = 1.0 + static_cast<double>(old.gh - 3.f) * (pixel.y - old.calibration_min_y()) / (old.calibration_max_y() + 1 - old.calibration_min_y());
        gx = std::min(std::max(gx, 0.0), old.gw - 1.001);
        gy = std::min(std::max(gy, 0.0), old.gh - 1.001);
        const size_t o = static_cast<size_t>(y) * target_resolution_x + x;
        detail::interpolate_bilinear3(old.grids.data(), old.gw, gx, gy, &fresh->grids[3 * o]);                      // directions
        detail::interpolate_bilinear3(old.grids.data() + 3 * old_n, old.gw, gx, gy, &fresh->grids[3 * (new_n + o)]);  // line origins
      }
    return fresh;
  }
  if (model_to_optimize.type() == T::NoncentralGeneric) return nullptr;  // not implemented in the reference either (:426-429)
  if (model_type != T::CentralGeneric && model_type != T::NoncentralGeneric) return nullptr;
  // dense direction model of the old camera
  std::vector<double> pixels(2 * static_cast<size_t>(w) * h), directions;
  std::vector<int32_t> ok;
  for (int y = 0; y < h; ++y)
    for (int x = 0; x < w; ++x) {
      pixels[2 * (static_cast<size_t>(y) * w + x)] = x + 0.5;
      pixels[2 * (static_cast<size_t>(y) * w + x) + 1] = y + 0.5;
    }
  unproject_many(model_to_optimize, pixels, &directions, &ok);
  if (directions.size() != 3 * ok.size() || ok.size() != static_cast<size_t>(w) * h) throw std::runtime_error("ResampleModel: un-projector returned a wrong size");
  for (size_t i = 0; i < ok.size(); ++i)
    if (!ok[i]) directions[3 * i] = directions[3 * i + 1] = directions[3 * i + 2] = std::nan("");
  const int area_w = calibration_max_x - calibration_min_x + 1, area_h = calibration_max_y - calibration_min_y + 1;
  const int subsample_step = std::max(1, std::min(area_w / 300, area_h / 300));  // std::round(int / int): already integral
  std::shared_ptr<CentralGenericModel> central(new CentralGenericModel(target_resolution_x, target_resolution_y, calibration_min_x,
                                                                       calibration_min_y, calibration_max_x, calibration_max_y, w, h));
  if (!FitToDenseModel(central.get(), directions, w, h, subsample_step, 3, fit)) return nullptr;
  if (model_type == T::NoncentralGeneric) {
    // noncentral_generic.cc:136-146: same directions, all line origins at the optical centre
    std::shared_ptr<NoncentralGenericModel> fresh(new NoncentralGenericModel(target_resolution_x, target_resolution_y, calibration_min_x,
                                                                             calibration_min_y, calibration_max_x, calibration_max_y, w, h));
    std::copy(central->grid.begin(), central->grid.end(), fresh->grids.begin());
    return fresh;
  }
  return central;
}

// calibration.cc:572-612: re-sample every camera whose grid resolution differs from the one wanted on this pyramid level,
// or whose type differs. Returns the number of re-sampled models.
inline int ResampleModelsIfNecessary(const Dataset& dataset, BAState* state, CameraModel::Type model_type, int approx_pixels_per_cell,
                                     int pyramid_level, const UnprojectMany& unproject_many = UnprojectManyOnDevice,
                                     const FitGridPoints& fit = FitGridPointsOnDevice) {
  int count = 0;
  for (int c = 0; c < dataset.num_cameras(); ++c) {
    CameraModel& model = *state->intrinsics[c];
    int loaded_x = 0, loaded_y = 0;
    const bool has_grid = model.GetGridResolution(&loaded_x, &loaded_y);
    const int exterior = has_grid ? 1 : 0;  // central_generic.h / noncentral_generic.h: exterior_cells_per_side() == 1
    int full_x, full_y, want_x, want_y;
    ComputeGridResolution(model.calibration_max_x() - model.calibration_min_x() + 1, model.calibration_max_y() - model.calibration_min_y() + 1,
                          exterior, approx_pixels_per_cell, &full_x, &full_y);
    CalcGridResolutionForLevel(pyramid_level, full_x, full_y, &want_x, &want_y);
    if ((has_grid && (loaded_x != want_x || loaded_y != want_y)) || model.type() != model_type) {
      std::shared_ptr<CameraModel> fresh = ResampleModel(model, model.calibration_min_x(), model.calibration_min_y(), model.calibration_max_x(),
                                                         model.calibration_max_y(), model_type, want_x, want_y, unproject_many, fit);
      if (fresh) {
        state->intrinsics[c] = fresh;
        ++count;
      }
    }
  }
  return count;
}

}  // namespace b200ba_shim
