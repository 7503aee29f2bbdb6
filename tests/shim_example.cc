// Compiles and links the C++ shim (include/b200ba_shim.hpp) against libb200ba.so and runs one
// OptimizeJointly call on a tiny pinhole-like central-generic problem. Without a GPU the call
// must throw (no CPU fallback); with a GPU the cost must decrease.
#include <cstdio>
#include <cstdlib>
#include "b200ba_shim.hpp"
using namespace b200ba_shim;
int main() {
  const int W = 320, H = 240, GW = 8, GH = 7;
  auto model = std::make_shared<CentralGenericModel>(GW, GH, 0, 0, W - 1, H - 1, W, H);
  for (int y = 0; y < GH; ++y)
    for (int x = 0; x < GW; ++x) {
      const double px = ((x - 1.0) / (GW - 3.0)) * W, py = ((y - 1.0) / (GH - 3.0)) * H;
      double d[3] = {(px - W / 2.0) / 200.0, (py - H / 2.0) / 200.0, 1.0};
      const double n = std::sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
      for (int k = 0; k < 3; ++k) model->grid[3 * (x + y * GW) + k] = d[k] / n;
    }
  BAState state;
  state.intrinsics.push_back(model);
  state.camera_tr_rig.push_back(SE3d());
  Dataset dataset(1);
  unsigned s = 12345;
  auto rnd = [&]() { s = s * 1664525u + 1013904223u; return (s >> 8) / 16777216.0 * 2 - 1; };
  for (int p = 0; p < 40; ++p) {
    state.points.push_back(Vec3d{0.5 * rnd(), 0.35 * rnd(), 0.05 * rnd()});
    state.feature_id_to_points_index[p] = p;
  }
  for (int i = 0; i < 12; ++i) {
    SE3d T;
    T.tx = 0.05 * rnd(); T.ty = 0.05 * rnd(); T.tz = 1.0 + 0.1 * rnd();
    state.rig_tr_global.push_back(T);
    state.image_used.push_back(true);
    auto is = dataset.NewImageset();
    for (int p = 0; p < 40; ++p) {
      const Vec3d& q = state.points[p];
      PointFeature f;
      // pinhole observation + a little noise: the B-spline model fits it to a fraction of a pixel
      f.xy = Vec2f{static_cast<float>(200.0 * (q.x + T.tx) / (q.z + T.tz) + W / 2.0 + 0.05 * rnd()),
                   static_cast<float>(200.0 * (q.y + T.ty) / (q.z + T.tz) + H / 2.0 + 0.05 * rnd())};
      f.id = p;
      is->FeaturesOfCamera(0).push_back(f);
    }
  }
  state.ComputeFeatureIdToPointsIndex(&dataset);
  for (auto& p : state.points) { p.x += 0.002 * rnd(); p.y += 0.002 * rnd(); p.z += 0.002 * rnd(); }
  try {
    double lambda = -1;
    bool performed = false;
    double c0 = OptimizeJointly(dataset, &state, 1, lambda, 1e-4, 0, false, true, SchurMode::Dense, &lambda, &performed,
                                false, false, false, false, false, false);
    double c1 = OptimizeJointly(dataset, &state, 3, lambda, 1e-4, 0, false, true, SchurMode::Dense, &lambda, &performed,
                                false, false, false, false, false, false);
    std::printf("shim: cost after 1 iteration %.6g, after 4 iterations %.6g\n", c0, c1);
    // the product's outer loop (calibration.cc:187-304), device-resident
    b200ba_ba_report ba;
    const double c2 = RunBundleAdjustment(SchurMode::Dense, 5, 1e-12, &dataset, &state, 0, false, /*eliminate_points=*/true,
                                          nullptr, &ba);
    std::printf("shim: RunBundleAdjustment: %d iterations, cost %.6g\n", ba.iterations, c2);
    if (!(c2 <= c1 * (1 + 1e-9) + 1e-12)) return 1;
    // model fitting (FitToPixelDirections): pull the grid towards a shifted pinhole camera
    std::vector<Vec2d> pixels;
    std::vector<Vec3d> directions;
    for (int y = 0; y < H; y += 8)
      for (int x = 0; x < W; x += 8) {
        pixels.push_back(Vec2d{x + 0.5, y + 0.5});
        double d[3] = {(x + 0.5 - (W / 2.0 - 4)) / 200.0, (y + 0.5 - (H / 2.0 + 6)) / 200.0, 1.0};
        const double n = std::sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
        directions.push_back(Vec3d{d[0] / n, d[1] / n, d[2] / n});
      }
    model->FitToPixelDirections(pixels, directions, 5);
    return (c1 <= c0) ? 0 : 1;
  } catch (const std::exception& e) {
    std::printf("shim: error: %s\n", e.what());
    return 3;
  }
}
