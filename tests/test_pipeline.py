"""The outer loop around the hot path (RunBundleAdjustment, ChooseNiceCameraOrientation)."""
import os

import numpy as np
import pytest

from camera_calibration_b200 import api, cabi, io, pipeline, synthetic
from tests import helpers


def test_rotation_helpers():
    rng = np.random.default_rng(0)
    for _ in range(20):
        a = rng.standard_normal(3)
        b = rng.standard_normal(3)
        R = pipeline._from_two_vectors(a, b)
        assert np.allclose(R @ R.T, np.eye(3), atol=1e-12) and abs(np.linalg.det(R) - 1) < 1e-12
        assert np.allclose(R @ (a / np.linalg.norm(a)), b / np.linalg.norm(b), atol=1e-12)
        q = pipeline._rot_to_quat(R)
        assert np.allclose(synthetic.quat_to_rot(q), R, atol=1e-12)


@pytest.mark.gpu
def test_choose_nice_camera_orientation_keeps_the_cost():
    """Rotating the model and left-multiplying the rotation onto camera_tr_rig leaves every
    residual unchanged (calibration.cc:248-254)."""
    sp = synthetic.make_problem(2, n_imagesets=8, lattice=(10, 8), image_size=(410, 290))
    ds, st = api.dataset_from_flat(sp.problem, sp.init_state)
    opt = cabi.default_options()
    with api.BundleAdjuster(sp.problem) as adj:
        adj.set_state(sp.init_state)
        e0 = adj.evaluate(opt)
        m = st.intrinsics[0]
        # tilt the model first so that the orientation change is not trivial
        tilt = synthetic.quat_to_rot(synthetic.so3_exp(np.array([0.05, -0.08, 0.2])))
        m.m_grid = m.m_grid @ tilt.T
        R = pipeline.ChooseNiceCameraOrientation(m) @ tilt
        ok, fwd, _ = m.Unproject(0.5 * m.width(), 0.5 * m.height())
        assert ok and abs(fwd[0]) < 1e-9 and abs(fwd[1]) < 1e-9 and fwd[2] > 0.999
        rt = np.concatenate([pipeline._rot_to_quat(R), np.zeros(3)])
        fs = sp.init_state.copy()
        fs.intrinsics = [m.flat_intrinsics().copy()]
        fs.camera_tr_rig[0] = synthetic.pose_mul(rt, fs.camera_tr_rig[0])
        adj.set_state(fs)
        e1 = adj.evaluate(opt)
    assert abs(e1["total_cost"] - e0["total_cost"]) < 1e-7 * e0["total_cost"]
    v = (e0["costs"] >= 0) & (e1["costs"] >= 0)  # invalid residuals are NaN
    assert v.sum() >= len(v) - 2
    assert np.abs(e1["residuals"][v] - e0["residuals"][v]).max() < 1e-6


@pytest.mark.gpu
def test_run_bundle_adjustment_converges_and_checkpoints(tmp_path):
    """The product's loop on the reference's BA test problem: converges below the reference's
    threshold (test/util.h:567-568), writes a loadable state directory after every iteration."""
    problem, st, _ = helpers.reference_ba_test_problem(num_cameras=1, n_points=100, n_poses=60)
    ds, state = api.dataset_from_flat(problem, st)
    out = str(tmp_path / "state")
    seen = []
    costs = pipeline.RunBundleAdjustment(False, api.SchurMode.Dense, 40, 1e-9, ds, state, 0, False, out,
                                         on_iteration=lambda i, c: seen.append(os.path.exists(os.path.join(out, "points.yaml"))))
    assert all(seen) and len(costs) >= 3
    assert all(b <= a + 1e-12 for a, b in zip(costs, costs[1:]))
    assert costs[-1] <= 1e-6
    loaded = io.LoadBAState(out, ds)
    assert loaded is not None and len(loaded.points) == len(state.points)
    assert np.abs(loaded.points - state.points).max() < 1e-6
    # resume from the checkpoint: the cost of the loaded state is the converged one
    cost, _, _ = api.OptimizeJointly(ds, loaded, 1, -1, 1e-4, 0, False, False, api.SchurMode.Dense, print_progress=False)
    assert cost <= 1e-5


@pytest.mark.gpu
def test_device_resident_loop_equals_host_loop():
    """b200ba_run_bundle_adjustment (LM iterations + ChooseNiceCameraOrientation kernel + stop rule, state in
    HBM) against the same loop driven from Python with the numpy re-orientation and one host round trip
    per iteration: same costs, same iteration count, same final state (incl. the rotated grid and
    camera_tr_rig). Central-generic rig + single camera."""
    for cfg, kw in ((2, dict(n_imagesets=10, lattice=(10, 8), image_size=(410, 290))),
                    (4, dict(n_imagesets=8, lattice=(10, 8), image_size=(410, 290)))):
        sp = synthetic.make_problem(cfg, **kw)
        out = []
        for dev in (True, False):
            ds, state = api.dataset_from_flat(sp.problem, sp.init_state)
            costs = pipeline.RunBundleAdjustment(False, api.SchurMode.Dense, 6, 1e-9, ds, state, 0, False, None,
                                                 eliminate_points=True, device_resident=dev)
            out.append((costs, state))
        (c0, s0), (c1, s1) = out
        assert len(c0) == len(c1) and np.allclose(c0, c1, rtol=1e-9)
        assert np.abs(np.asarray(s0.camera_tr_rig) - np.asarray(s1.camera_tr_rig)).max() < 1e-9
        assert np.abs(np.asarray(s0.points) - np.asarray(s1.points)).max() < 1e-9
        for a, b in zip(s0.intrinsics, s1.intrinsics):
            assert np.abs(a.flat_intrinsics() - b.flat_intrinsics()).max() < 1e-9


def _write_colmap(tmp, sp):
    """A COLMAP text model of a synthetic single-camera problem (perturbed poses / points)."""
    d = tmp / "colmap"
    d.mkdir()
    p = sp.problem
    with open(d / "images.txt", "w") as f:
        f.write("# Image list with two lines of data per image:\n")
        for i in range(p.n_imagesets):
            q = sp.init_state.rig_tr_global[i]
            f.write(f"{10 + i} {q[0]:.9g} {q[1]:.9g} {q[2]:.9g} {q[3]:.9g} {q[4]:.9g} {q[5]:.9g} {q[6]:.9g} 1 im{i}.png\n")
            sel = np.nonzero(p.obs_imageset == i)[0]
            row = " ".join(f"{p.obs_xy[o, 0]:.6f} {p.obs_xy[o, 1]:.6f} {100 + int(p.obs_point[o])}" for o in sel)
            f.write(row + " 5.0 6.0 -1\n")  # one observation without a 3D point: must be dropped
    with open(d / "points3D.txt", "w") as f:
        f.write("# 3D point list\n")
        for k in reversed(range(p.n_points)):  # unordered on purpose: the tool sorts by id
            x = sp.init_state.points[k]
            f.write(f"{100 + k} {x[0]:.9g} {x[1]:.9g} {x[2]:.9g} 255 0 0 0.5 1 2 3 4\n")
    return str(d)


def test_colmap_reader(tmp_path):
    sp = synthetic.make_problem(2, n_imagesets=4, lattice=(6, 5), image_size=(300, 220))
    d = _write_colmap(tmp_path, sp)
    _, st0 = api.dataset_from_flat(sp.problem, sp.gt_state)
    ds, st = io.LoadColmapProblem(st0.intrinsics[0], d)
    assert ds.ImagesetCount() == 4 and len(st.points) == sp.problem.n_points
    assert st.feature_id_to_points_index[100] == 0 and st.feature_id_to_points_index[100 + sp.problem.n_points - 1] == sp.problem.n_points - 1
    n = sum(len(ds.GetImageset(i).FeaturesOfCamera(0)["id"]) for i in range(4))
    assert n == sp.problem.n_obs  # the id -1 observations were dropped
    f0 = ds.GetImageset(0).FeaturesOfCamera(0)
    assert np.array_equal(f0["index"], f0["id"] - 100)
    assert np.abs(st.points - sp.init_state.points).max() < 1e-6  # parsed as float like the reference
    assert np.abs(st.rig_tr_global - sp.init_state.rig_tr_global).max() < 1e-6
    assert io.LoadColmapProblem(st0.intrinsics[0], str(tmp_path / "missing")) is None


@pytest.mark.gpu
def test_bundle_adjustment_tool(tmp_path):
    sp = synthetic.make_problem(2, n_imagesets=8, lattice=(10, 8), image_size=(410, 290))
    d = _write_colmap(tmp_path, sp)
    _, st0 = api.dataset_from_flat(sp.problem, sp.gt_state)
    sd = str(tmp_path / "state_in")
    io.SaveCameraModel(st0.intrinsics[0], os.path.join(sd, "intrinsics0.yaml"))
    out = str(tmp_path / "state_out")
    assert pipeline.BundleAdjustment(sd, d, out, max_iteration_count=6) == 0
    cost = float(open(os.path.join(out, "cost.txt")).read())
    # intrinsics are the ground truth and stay fixed: poses and points are recovered to the noise level
    n = sp.problem.n_obs
    assert cost < 0.5 * n * 2 * (0.05 ** 2) * 1.3
    st = io.LoadBAState(out)
    assert st is not None and len(st.points) == sp.problem.n_points
    assert pipeline.BundleAdjustment(str(tmp_path / "nope"), d, out) == 1


def _outlier_scene(seed=0):
    """Small central-generic scene at the true state with a few corrupted features."""
    sp = synthetic.make_problem(2, n_imagesets=6, lattice=(9, 7), image_size=(410, 290), seed=seed)
    ds, st = api.dataset_from_flat(sp.problem, sp.gt_state)
    rng = np.random.default_rng(seed)
    corrupted = []
    for i in (1, 3, 4):
        f = ds.GetImageset(i).FeaturesOfCamera(0)
        k = int(rng.integers(0, len(f["id"])))
        f["xy"][k] += np.float32(7.5)
        corrupted.append((i, int(f["id"][k])))
    return sp, ds, st, corrupted


def _reference_outlier_rule(ds, st, project_many, factor):
    """Literal per-feature restatement of calibration.cc:62-184 for the test."""
    errs = []
    per = {}
    for i in range(ds.ImagesetCount()):
        if not st.image_used[i]:
            continue
        f = ds.GetImageset(i).FeaturesOfCamera(0)
        T = st.image_tr_global(0, i)
        R = synthetic.quat_to_rot(T[:4])
        lp = st.points[f["index"]] @ R.T + T[4:7]
        px, ok = project_many(st.intrinsics[0], lp)
        e = np.linalg.norm(px - f["xy"].astype(np.float64), axis=1)
        per[i] = (e, ok)
        errs += list(e[ok])
    errs = sorted(errs)
    q1 = errs[int(np.float32(0.25) * np.float32(len(errs)) + np.float32(0.5))]
    q3 = errs[int(np.float32(0.75) * np.float32(len(errs)) + np.float32(0.5))]
    thr = q3 + np.float32(factor) * (q3 - q1)
    return {i: set(ds.GetImageset(i).FeaturesOfCamera(0)["id"][ok & (e <= thr)].tolist()) for i, (e, ok) in per.items()}


def _check_outlier_deletion(project_many):
    sp, ds, st, corrupted = _outlier_scene()
    expect = _reference_outlier_rule(ds, st, project_many, 1.5)
    before = sum(len(ds.GetImageset(i).FeaturesOfCamera(0)["id"]) for i in range(ds.ImagesetCount()))
    removed = pipeline.DeleteOutlierFeatures(0, ds, st, 1.5, project_many=project_many)
    after = sum(len(ds.GetImageset(i).FeaturesOfCamera(0)["id"]) for i in range(ds.ImagesetCount()))
    assert removed == before - after and removed >= len(corrupted)
    for i in range(ds.ImagesetCount()):
        f = ds.GetImageset(i).FeaturesOfCamera(0)
        assert set(f["id"].tolist()) == expect[i]
        assert len(f["xy"]) == len(f["id"]) == len(f["index"]) == len(f["last_projection"])
    for i, fid in corrupted:
        assert fid not in ds.GetImageset(i).FeaturesOfCamera(0)["id"]
    # an imageset left with < 3 features of the camera is dropped from the state
    sp, ds, st, _ = _outlier_scene(1)
    f = ds.GetImageset(2).FeaturesOfCamera(0)
    for key in list(f.keys()):
        f[key] = f[key][:3]
    f["xy"][0] += np.float32(40)
    pipeline.DeleteOutlierFeatures(0, ds, st, 1.5, project_many=project_many)
    assert st.image_used[2] is False and all(st.image_used[i] for i in (0, 1, 3, 4, 5))


def test_delete_outlier_features_host_logic(oracle_lib):
    """Quartile rule (calibration.cc:62-184) with the projections done by the CPU restatement."""
    from oracle import oracle

    def project_many(model, lp):
        px, ok = oracle.project(model.c_camera(), model.flat_intrinsics(), lp)
        return px, ok.astype(bool)
    _check_outlier_deletion(project_many)


@pytest.mark.gpu
def test_delete_outlier_features_on_device():
    _check_outlier_deletion(lambda model, lp: model.ProjectMany(lp))


def test_scale_to_metric():
    """calibration.cc:307-370 + ba_state.cc:60-76 on a 6x5 pattern shrunk by a known factor."""
    ds = api.Dataset(1)
    st = api.BAState()
    g = io.KnownGeometry()
    g.cell_length_in_meters = 0.02
    rng = np.random.default_rng(3)
    pts = []
    for y in range(5):
        for x in range(6):
            fid = 100 + x + 6 * y
            g.feature_id_to_position[fid] = (x, y)
            if (x, y) != (2, 2):  # one corner of the pattern was never triangulated
                st.feature_id_to_points_index[fid] = len(pts)
                pts.append([0.02 * x, 0.02 * y, 0.0])
    ds.known_geometries = [g]
    true_scale = 3.7
    st.points = np.array(pts) / true_scale + 1e-6 * rng.standard_normal((len(pts), 3))
    st.rig_tr_global = np.tile(np.array([1.0, 0, 0, 0, 0.1, 0.2, 0.3]), (2, 1))
    st.camera_tr_rig = np.array([[1.0, 0, 0, 0, 0.01, 0.0, 0.0]])
    st.image_used = [True, True]
    m = api.NoncentralGenericModel(6, 5, 0, 0, 99, 79, 100, 80)
    m.SetPointGrid(np.ones((5, 6, 3)))
    st.intrinsics = [m]
    p0 = st.points.copy()
    factor = pipeline.ScaleToMetric(ds, st)
    assert abs(factor - true_scale) < 1e-2
    assert np.allclose(st.points, factor * p0)
    assert np.allclose(st.rig_tr_global[:, 4:], factor * np.array([0.1, 0.2, 0.3])) and np.allclose(st.rig_tr_global[:, :4], [1, 0, 0, 0])
    assert np.allclose(st.camera_tr_rig[0, 4:], [0.01 * factor, 0, 0])
    assert np.allclose(m.point_grid(), factor)
    # literal restatement of the averaging
    logs = []
    idx = st.feature_id_to_points_index
    for fid, (x, y) in g.feature_id_to_position.items():
        for dx, dy in ((1, 0), (0, 1)):
            nid = 100 + (x + dx) + 6 * (y + dy)
            if x + dx < 6 and y + dy < 5 and fid in idx and nid in idx:
                logs.append(np.log(0.02 / np.linalg.norm(p0[idx[fid]] - p0[idx[nid]])))
    assert abs(factor - np.exp(np.mean(logs))) < 1e-12
