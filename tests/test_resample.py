"""Row f-4: model resampling -- FitToDenseModel / FitToPixelDirections / ResampleModel
(APP/calibration.cc:373-522, APP/models/central_generic.cc:267-431,551-568). The recipe and the
threshold are those of the reference's own test (APP/test/util.h:40-78,213-271
``TestModelOptimization`` + ``VerifyUnprojections``: 640x480, 8x6 grid, pinhole ground truth,
0.5 |unprojection - direction|^2 < 5e-4 at every pixel)."""
import numpy as np
import pytest

from camera_calibration_b200 import api, pipeline
from tests import helpers

W, H = 640, 480
K_EPSILON = float(np.float32(5e-4))


def pinhole_dense(fx, fy, cx, cy, w=W, h=H):
    xs, ys = np.meshgrid(np.arange(w) + 0.5, np.arange(h) + 0.5)
    d = np.stack([(xs - cx) / fx, (ys - cy) / fy, np.ones_like(xs)], -1)
    return d / np.linalg.norm(d, axis=-1, keepdims=True)


def oracle_fit(gw, gh, grid, gp, d, iterations):
    from oracle import oracle
    return oracle.fit_directions(gw, gh, grid, gp, d, iterations)


def oracle_unproject_many(model, pixels):
    from oracle import oracle
    d, o, ok = oracle.unproject(model.c_camera(), model.flat_intrinsics(), pixels)
    return d, o, ok


def device_unproject_many(model, pixels):
    return model.UnprojectMany(pixels)


def verify_unprojections(model, dense, unproject_many):
    h, w = dense.shape[:2]
    xs, ys = np.meshgrid(np.arange(w) + 0.5, np.arange(h) + 0.5)
    d, _, ok = unproject_many(model, np.stack([xs.ravel(), ys.ravel()], -1))
    cost = 0.5 * ((d - dense.reshape(-1, 3)) ** 2).sum(-1)
    assert ok.sum() > 0.9 * len(ok)
    assert cost[ok].max() < K_EPSILON
    return cost[ok].max()


def model_optimization(fit_fn, unproject_many):
    """TestModelOptimization (test/util.h:213-271)."""
    model = api.CentralGenericModel(8, 6, 0, 0, W - 1, H - 1, W, H)
    dense = pinhole_dense(H / 2, H / 2, W / 2, H / 2)
    assert model.FitToDenseModel(dense, 2, 10, fit_fn=fit_fn)
    assert np.allclose(np.linalg.norm(model.grid(), axis=-1), 1.0, atol=1e-12)
    e1 = verify_unprojections(model, dense, unproject_many)
    # FitToPixelDirections towards a shifted pinhole camera
    shifted = pinhole_dense(H / 2, H / 2, W / 2 - 10, H / 2 + 20)
    ys, xs = np.meshgrid(np.arange(0, H, 10), np.arange(0, W, 10), indexing="ij")
    pixels = np.stack([xs.ravel() + 0.5, ys.ravel() + 0.5], -1)
    rep = model.FitToPixelDirections(pixels, shifted[ys.ravel(), xs.ravel()], 10, fit_fn=fit_fn)
    assert rep.final_cost < rep.initial_cost
    e2 = verify_unprojections(model, shifted, unproject_many)
    return model, e1, e2


def test_model_optimization_host_logic_with_oracle_fit(oracle_lib):
    model_optimization(oracle_fit, oracle_unproject_many)


def test_fit_to_dense_model_fills_invalid_regions(oracle_lib):
    """Control points whose pixel is invalid take the nearest valid pixel (radius < 5) or are
    extrapolated linearly from their neighbours (central_generic.cc:275-372)."""
    dense = pinhole_dense(H / 2, H / 2, W / 2, H / 2)
    dense[:60, :, :] = np.nan      # a band wider than the search radius: forces the extrapolation
    dense[200:203, 300:303, :] = np.nan  # a small hole: nearest-pixel search
    model = api.CentralGenericModel(8, 6, 0, 60, W - 1, H - 1, W, H)
    assert model.FitToDenseModel(dense, 4, 3, fit_fn=oracle_fit)
    assert np.isfinite(model.grid()).all()
    full = pinhole_dense(H / 2, H / 2, W / 2, H / 2)
    xs, ys = np.meshgrid(np.arange(W) + 0.5, np.arange(H) + 0.5)
    d, _, ok = oracle_unproject_many(model, np.stack([xs.ravel(), ys.ravel()], -1))
    cost = 0.5 * ((d - full.reshape(-1, 3)) ** 2).sum(-1)
    assert ok.reshape(H, W)[60:, :].all() and not ok.reshape(H, W)[:60, :].any()
    assert cost[ok].max() < K_EPSILON
    # nothing to initialise from: the fit reports failure
    empty = np.full((H, W, 3), np.nan)
    assert not api.CentralGenericModel(8, 6, 0, 0, W - 1, H - 1, W, H).FitToDenseModel(empty, 4, 3, fit_fn=oracle_fit)


def test_grid_pixel_maps_are_inverse():
    m = api.CentralGenericModel(12, 9, 3, 5, 600, 400, W, H)
    for gx, gy in ((1, 1), (4, 3), (10, 7)):
        p = m.GridPointToPixelCornerConv(gx, gy)
        g = m.PixelCornerConvToGridPoint(p[0], p[1])
        assert np.allclose(g, [gx, gy], atol=1e-4)  # the forward map is evaluated in float
    assert np.allclose(m.GridPointToPixelCornerConv(1, 1), [3, 5])


def test_resample_noncentral_to_noncentral_is_bilinear(oracle_lib):
    rng = np.random.default_rng(5)
    old = api.NoncentralGenericModel(10, 8, 0, 0, 399, 299, 400, 300)
    d = helpers.xy1_grid(10, 8) + 0.01 * rng.standard_normal((8, 10, 3))
    old.SetDirectionGrid(d / np.linalg.norm(d, axis=-1, keepdims=True))
    old.SetPointGrid(0.01 * rng.standard_normal((8, 10, 3)))
    ok, same = pipeline.ResampleModel(old, None, 0, 0, 399, 299, api.CameraModel.Type.NoncentralGeneric, 10, 8)
    assert ok and same.type() == api.CameraModel.Type.NoncentralGeneric
    # same resolution: every new grid point falls (up to float rounding) on an old one; the last
    # row / column is clamped to size - 1.001 (calibration.cc:401)
    assert np.abs(same.direction_grid()[:-1, :-1] - old.direction_grid()[:-1, :-1]).max() < 1e-4
    assert np.abs(same.point_grid()[:-1, :-1] - old.point_grid()[:-1, :-1]).max() < 1e-5
    ok, finer = pipeline.ResampleModel(old, None, 0, 0, 399, 299, api.CameraModel.Type.NoncentralGeneric, 17, 13)
    assert ok and finer.direction_grid().shape == (13, 17, 3) and np.isfinite(finer.point_grid()).all()
    # a non-central source cannot be resampled to anything else (calibration.cc:426-429)
    ok, _ = pipeline.ResampleModel(old, None, 0, 0, 399, 299, api.CameraModel.Type.CentralGeneric, 10, 8)
    assert not ok


def _resample_real_camera(fit_fn, unproject_many):
    cam, grid = helpers.real_camera()
    old = api.CentralGenericModel(cam.grid_width, cam.grid_height, cam.calibration_min_x, cam.calibration_min_y,
                                  cam.calibration_max_x, cam.calibration_max_y, cam.width, cam.height)
    old.SetGrid(grid)
    args = (cam.calibration_min_x, cam.calibration_min_y, cam.calibration_max_x, cam.calibration_max_y)
    ok, new = pipeline.ResampleModel(old, None, *args, api.CameraModel.Type.CentralGeneric, 20, 15, fit_fn=fit_fn,
                                     unproject_many=unproject_many)
    assert ok and new.GetGridResolution() == (20, 15)
    xs, ys = np.meshgrid(np.arange(cam.calibration_min_x + 2, cam.calibration_max_x - 1, 7) + 0.5,
                         np.arange(cam.calibration_min_y + 2, cam.calibration_max_y - 1, 7) + 0.5)
    px = np.stack([xs.ravel(), ys.ravel()], -1)
    d0, _, ok0 = unproject_many(old, px)
    d1, _, ok1 = unproject_many(new, px)
    assert ok0.all() and ok1.all()
    assert (0.5 * ((d0 - d1) ** 2).sum(-1)).max() < K_EPSILON
    ok, nc = pipeline.ResampleModel(old, None, *args, api.CameraModel.Type.NoncentralGeneric, 20, 15, fit_fn=fit_fn,
                                    unproject_many=unproject_many)
    assert ok and nc.type() == api.CameraModel.Type.NoncentralGeneric
    assert np.abs(nc.direction_grid() - new.grid()).max() < 1e-12 and not nc.point_grid().any()
    return new


def test_resample_real_camera_host_logic(oracle_lib):
    _resample_real_camera(oracle_fit, oracle_unproject_many)


# ---- the same through the device (b200ba_fit_directions, b200ba_unproject) ---------------------
@pytest.mark.gpu
def test_fit_directions_matches_oracle(oracle_lib):
    """Same inputs through the CUDA path and the CPU restatement: same accept sequence, costs and
    grid (1e-9)."""
    from oracle import oracle
    rng = np.random.default_rng(11)
    gw, gh = 12, 9
    model = api.CentralGenericModel(gw, gh, 0, 0, W - 1, H - 1, W, H)
    truth = pinhole_dense(300.0, 310.0, 330.0, 250.0)
    ys, xs = np.meshgrid(np.arange(0, H, 3), np.arange(0, W, 3), indexing="ij")
    gp = model.PixelCornerConvToGridPoint(xs.ravel() + 0.5, ys.ravel() + 0.5)
    dirs = truth[ys.ravel(), xs.ravel()]
    g0 = helpers.xy1_grid(gw, gh) * np.array([0.25, 0.25, 1.0]) + np.array([-1.2, -0.9, 0.0])
    g0 = g0 / np.linalg.norm(g0, axis=-1, keepdims=True)
    g0 = g0 + 0.003 * rng.standard_normal(g0.shape)
    g0 = g0 / np.linalg.norm(g0, axis=-1, keepdims=True)
    for iterations in (1, 4):
        ref_grid, ref = oracle.fit_directions(gw, gh, g0, gp, dirs, iterations)
        model.SetGrid(g0.copy())
        rep = model._fit_grid_points(gp, dirs, iterations)
        assert rep.num_iterations_performed == ref.num_iterations_performed
        assert rep.lm_attempts == ref.lm_attempts
        assert abs(rep.initial_cost - ref.initial_cost) < 1e-10 * ref.initial_cost
        assert abs(rep.final_cost - ref.final_cost) < 1e-8 * ref.initial_cost
        assert np.abs(model.grid() - ref_grid).max() < 1e-9
        assert rep.final_cost < rep.initial_cost


@pytest.mark.gpu
def test_model_optimization_on_device():
    model_optimization(None, device_unproject_many)


@pytest.mark.gpu
def test_resample_real_camera_on_device():
    _resample_real_camera(None, device_unproject_many)


def test_grid_resolution_helpers():
    """calibration.cc:531-569,615-641; SURVEY.md 8d: 2050x1450 at 25 px per cell -> 84x60."""
    assert pipeline.ComputeGridResolution(2050, 1450, 1, 25) == (84, 60)
    assert pipeline.ComputeGridResolution(1200, 950, 1, 25) == (50, 40)
    m = api.CentralGenericModel(84, 60, 0, 0, 2049, 1449, 2050, 1450)
    assert pipeline.ComputeGridResolutionForModel(m, 25) == (84, 60)
    assert pipeline.CalcGridResolutionForLevel(0, 84, 60) == (84, 60)
    assert pipeline.CalcGridResolutionForLevel(2, 84, 60) == (int(84 * 1.333 ** -2 + 0.5), int(60 * 1.333 ** -2 + 0.5))
    ds = api.Dataset(1)
    for k, off in enumerate((0.0, 7.9)):
        s = ds.NewImageset()
        s.SetFeaturesOfCamera(0, np.array([[10.7 + off, 20.2], [300.9, 250.99 - off]]), [1, 2])
    assert pipeline.ComputeIntegerBoundingRectForFeatures(ds, 0, [True, True]) == (10, 20, 300, 250)
    assert pipeline.ComputeIntegerBoundingRectForFeatures(ds, 0, [False, True]) == (18, 20, 300, 243)


def test_resample_models_if_necessary(oracle_lib):
    cam, grid = helpers.real_camera()
    old = api.CentralGenericModel(cam.grid_width, cam.grid_height, cam.calibration_min_x, cam.calibration_min_y,
                                  cam.calibration_max_x, cam.calibration_max_y, cam.width, cam.height)
    old.SetGrid(grid)
    st = api.BAState()
    st.intrinsics = [old]
    st.camera_tr_rig = np.array([[1.0, 0, 0, 0, 0, 0, 0]])
    ds = api.Dataset(1)
    want = pipeline.CalcGridResolutionForLevel(1, *pipeline.ComputeGridResolutionForModel(old, 50))
    n = pipeline.ResampleModelsIfNecessary(ds, st, api.CameraModel.Type.CentralGeneric, 50, 1, fit_fn=oracle_fit,
                                           unproject_many=oracle_unproject_many)
    assert n == 1 and st.intrinsics[0].GetGridResolution() == want
    # already at the wanted resolution and type: untouched
    m = st.intrinsics[0]
    assert pipeline.ResampleModelsIfNecessary(ds, st, api.CameraModel.Type.CentralGeneric, 50, 1, fit_fn=oracle_fit,
                                              unproject_many=oracle_unproject_many) == 0 and st.intrinsics[0] is m
