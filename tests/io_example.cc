// Drives include/b200ba_io.hpp and include/b200ba_pipeline.hpp from the command line so that tests/test_cpp_io.py
// can exchange files with the Python mirror (camera_calibration_b200/io.py, pipeline.py). Host logic only.
#include <cstdio>
#include <cstring>
#include <string>

#include "b200ba_io.hpp"
#include "b200ba_pipeline.hpp"

using namespace b200ba_shim;

// a deterministic stand-in for the device projection (same formula in the Python test): pinhole with
// f = 400, c = (320, 240); points behind the camera or outside 640 x 480 fail
static void pinhole(CameraModel&, const std::vector<double>& lp, std::vector<double>* px, std::vector<int32_t>* ok) {
  const size_t n = lp.size() / 3;
  px->assign(2 * n, 0.0);
  ok->assign(n, 0);
  for (size_t i = 0; i < n; ++i) {
    const double x = lp[3 * i], y = lp[3 * i + 1], z = lp[3 * i + 2];
    if (z <= 0) continue;
    const double u = 400.0 * x / z + 320.0, v = 400.0 * y / z + 240.0;
    (*px)[2 * i] = u;
    (*px)[2 * i + 1] = v;
    (*ok)[i] = (u >= 0 && v >= 0 && u < 640 && v < 480) ? 1 : 0;
  }
}

// stand-ins for the device calls of ResampleModel (same formulas in the Python test): un-projection of a pinhole with
// f = 400, c = (320, 240), undefined in a strip on the left and in a small square; the fit leaves the grid as initialised
static void pinhole_unproject(CameraModel&, const std::vector<double>& px, std::vector<double>* dirs, std::vector<int32_t>* ok) {
  const size_t n = px.size() / 2;
  dirs->assign(3 * n, 0.0);
  ok->assign(n, 0);
  for (size_t i = 0; i < n; ++i) {
    const double x = px[2 * i], y = px[2 * i + 1];
    const double dx = (x - 320.0) / 400.0, dy = (y - 240.0) / 400.0;
    const double norm = std::sqrt(dx * dx + dy * dy + 1.0);
    (*dirs)[3 * i] = dx / norm;
    (*dirs)[3 * i + 1] = dy / norm;
    (*dirs)[3 * i + 2] = 1.0 / norm;
    (*ok)[i] = (x < 7.0 || (x > 200 && x < 204 && y > 100 && y < 104)) ? 0 : 1;
  }
}
static void no_fit(CentralGenericModel&, const std::vector<double>& gp, const std::vector<double>& d, int) {
  std::printf("samples %zu %zu\n", gp.size() / 2, d.size() / 3);
}

int main(int argc, char** argv) {
  if (argc < 2) return 2;
  const std::string mode = argv[1];
  try {
    if (mode == "dataset" && argc == 4) {
      std::shared_ptr<Dataset> ds;
      if (!LoadDataset(argv[2], &ds)) { std::printf("load failed\n"); return 1; }
      std::printf("cameras %d imagesets %d geometries %zu\n", ds->num_cameras(), ds->ImagesetCount(), ds->known_geometries().size());
      return SaveDataset(argv[3], *ds) ? 0 : 1;
    }
    if (mode == "state" && (argc == 4 || argc == 5)) {
      std::shared_ptr<Dataset> ds;
      if (argc == 5 && !LoadDataset(argv[4], &ds)) { std::printf("dataset load failed\n"); return 1; }
      BAState st;
      if (!LoadBAState(argv[2], &st, ds.get())) { std::printf("load failed\n"); return 1; }
      std::printf("cameras %d imagesets %zu points %zu\n", st.num_cameras(), st.rig_tr_global.size(), st.points.size());
      if (ds) {
        long long sum = 0;
        for (int i = 0; i < ds->ImagesetCount(); ++i)
          for (int c = 0; c < ds->num_cameras(); ++c)
            for (const PointFeature& f : ds->GetImageset(i)->FeaturesOfCamera(c)) sum += f.index;
        std::printf("index_sum %lld\n", sum);
      }
      return SaveBAState(argv[3], st) ? 0 : 1;
    }
    if (mode == "scale" && argc == 5) {
      std::shared_ptr<Dataset> ds;
      BAState st;
      if (!LoadDataset(argv[2], &ds) || !LoadBAState(argv[3], &st, ds.get())) { std::printf("load failed\n"); return 1; }
      const double factor = ScaleToMetric(*ds, &st);
      std::printf("factor %.17g\n", factor);
      return SaveBAState(argv[4], st) ? 0 : 1;
    }
    if (mode == "outliers" && argc == 7) {
      std::shared_ptr<Dataset> ds;
      BAState st;
      if (!LoadDataset(argv[2], &ds) || !LoadBAState(argv[3], &st, ds.get())) { std::printf("load failed\n"); return 1; }
      const int removed = DeleteOutlierFeatures(std::atoi(argv[4]), ds.get(), &st, static_cast<float>(std::atof(argv[5])), pinhole);
      std::printf("removed %d\nused", removed);
      for (bool u : st.image_used) std::printf(" %d", u ? 1 : 0);
      std::printf("\n");
      return SaveDataset(argv[6], *ds) ? 0 : 1;
    }
    if (mode == "gridres" && argc == 7) {
      int rx, ry, lx, ly;
      ComputeGridResolution(std::atoi(argv[2]), std::atoi(argv[3]), std::atoi(argv[4]), std::atoi(argv[5]), &rx, &ry);
      CalcGridResolutionForLevel(std::atoi(argv[6]), rx, ry, &lx, &ly);
      std::printf("%d %d %d %d\n", rx, ry, lx, ly);
      return 0;
    }
    if (mode == "bounds" && argc == 4) {
      std::shared_ptr<Dataset> ds;
      if (!LoadDataset(argv[2], &ds)) { std::printf("load failed\n"); return 1; }
      int a, b, c, d;
      std::vector<bool> used(ds->ImagesetCount(), true);
      used[1] = false;
      ComputeIntegerBoundingRectForFeatures(*ds, std::atoi(argv[3]), used, &a, &b, &c, &d);
      std::printf("%d %d %d %d\n", a, b, c, d);
      return 0;
    }
    if (mode == "resample" && argc == 8) {
      // resample <model.yaml> <target type: 0 central, 1 non-central> <res x> <res y> <out.yaml> <unused>
      std::shared_ptr<CameraModel> model = LoadCameraModel(argv[2]);
      if (!model) { std::printf("load failed\n"); return 1; }
      const CameraModel::Type target = std::atoi(argv[3]) == 0 ? CameraModel::Type::CentralGeneric : CameraModel::Type::NoncentralGeneric;
      std::shared_ptr<CameraModel> fresh = ResampleModel(*model, model->calibration_min_x(), model->calibration_min_y(), model->calibration_max_x(),
                                                         model->calibration_max_y(), target, std::atoi(argv[4]), std::atoi(argv[5]),
                                                         pinhole_unproject, no_fit);
      if (!fresh) { std::printf("not resampled\n"); return 0; }
      return SaveCameraModel(*fresh, argv[6]) ? 0 : 1;
    }
    if (mode == "malformed" && argc == 3) {
      // every loader must answer false (not crash, not throw) on this file
      std::shared_ptr<Dataset> ds;
      std::vector<bool> used;
      std::vector<SE3d> poses;
      BAState st;
      const bool a = LoadDataset(argv[2], &ds), b = static_cast<bool>(LoadCameraModel(argv[2])), c = LoadPoses(&used, &poses, argv[2]),
                 d = LoadPointsAndIndexMapping(&st, argv[2]);
      std::printf("%d %d %d %d\n", a, b, c, d);
      return 0;
    }
  } catch (const std::exception& e) {
    std::printf("exception: %s\n", e.what());
    return 4;
  }
  return 2;
}
