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
