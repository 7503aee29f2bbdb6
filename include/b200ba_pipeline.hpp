// b200ba_pipeline.hpp -- C++ host logic of the callers either side of the hot path:
// the outlier deletion between bundle-adjustment rounds and the metric rescaling, over the containers of
// b200ba_shim.hpp. (RunBundleAdjustment itself -- 8f-2 -- is in b200ba_shim.hpp and runs device-resident in the
// library.) The Python mirror is camera_calibration_b200/pipeline.py.
//
// Header-only. Every numerical step on observations (the re-projection of all features of a camera) runs in
// libb200ba.so through b200ba_project; there is no CPU fallback.
#pragma once

#include <algorithm>
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

}  // namespace b200ba_shim
