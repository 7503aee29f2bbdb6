"""Shared builders for the parity tests (recipes of the reference's own tests)."""
import json
import os

import numpy as np

from camera_calibration_b200 import cabi, synthetic
from camera_calibration_b200.cabi import Camera, FlatProblem, FlatState

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def make_camera(model_type, w, h, rect, gw, gh) -> Camera:
    c = Camera()
    c.model_type = model_type
    c.width, c.height = w, h
    c.calibration_min_x, c.calibration_min_y, c.calibration_max_x, c.calibration_max_y = rect
    c.grid_width, c.grid_height = gw, gh
    return c


def xy1_grid(gw, gh):
    """grid(x, y) = normalize(x, y, 1): test/util.h:128-133 and generic_models/src/main.cc:48-52."""
    g = np.zeros((gh, gw, 3))
    for y in range(gh):
        for x in range(gw):
            v = np.array([x, y, 1.0])
            g[y, x] = v / np.linalg.norm(v)
    return g


def real_camera():
    d = json.load(open(os.path.join(GOLDEN, "real_central_17x13.json")))
    cam = make_camera(cabi.MODEL_CENTRAL_GENERIC, d["width"], d["height"],
                      (d["calibration_min_x"], d["calibration_min_y"], d["calibration_max_x"], d["calibration_max_y"]),
                      d["grid_width"], d["grid_height"])
    grid = np.array(d["grid"], dtype=np.float64).reshape(d["grid_height"], d["grid_width"], 3)
    # the reference re-normalises directions on load (calibration_io.cc)
    grid = grid / np.linalg.norm(grid, axis=-1, keepdims=True)
    return cam, grid


def orthographic_noncentral():
    """test/noncentral_generic_test.cc:49-72: 4x4 grids, origins (x, y, 0), directions (0, 0, 1)."""
    cam = make_camera(cabi.MODEL_NONCENTRAL_GENERIC, 100, 100, (0, 0, 99, 99), 4, 4)
    pg = np.zeros((4, 4, 3))
    dg = np.zeros((4, 4, 3))
    for y in range(4):
        for x in range(4):
            pg[y, x] = (x, y, 0)
            dg[y, x] = (0, 0, 1)
    return cam, np.concatenate([dg.reshape(-1), pg.reshape(-1)])


def reference_ba_test_problem(num_cameras=1, seed=0, n_points=150, n_poses=100):
    """TestOptimizeJointly (test/util.h:275-571): 600x400, 5x5 grid from a pinhole, 150 points in
    a 13 x 7 x 2 box 5 m away, 100 poses, noise-free float observations; perturbed points
    (+-0.05), poses (exp(0.04 U)), rig poses, grid (+0.02 U, renormalised)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    U = lambda *s: rng.uniform(-1, 1, size=s)
    W, H = 600, 400
    cams, gt_intr = [], []
    for c in range(num_cameras):
        cam = make_camera(cabi.MODEL_CENTRAL_GENERIC, W, H, (0, 0, W - 1, H - 1), 5, 5)
        fx = H / 2.0 + 2.0 * c
        fy = H / 2.0
        gx, gy = np.meshgrid(np.arange(5.0), np.arange(5.0))
        px, py = synthetic.grid_point_to_pixel(cam, gx, gy)
        d = np.stack([(px - W / 2.0) / fx, (py - H / 2.0) / fy, np.ones_like(px)], -1)
        d /= np.linalg.norm(d, axis=-1, keepdims=True)
        cams.append(cam)
        gt_intr.append(d.reshape(-1))
    ctr = np.tile(synthetic.IDENTITY_POSE, (num_cameras, 1))
    for c in range(1, num_cameras):
        ctr[c] = synthetic.se3_exp(0.05 * U(6))
    pts = U(n_points, 3) * np.array([6.5, 3.5, 1.0])
    rtg = np.zeros((n_poses, 7))
    oi, oc, op, oxy = [], [], [], []
    for i in range(n_poses):
        base = synthetic.IDENTITY_POSE.copy()
        base[4:] = np.array([0, 0, 5.0]) + U(3)
        rtg[i] = synthetic.pose_mul(synthetic.se3_exp(0.05 * U(6)), base)
        for c in range(num_cameras):
            lp = synthetic.pose_apply(synthetic.pose_mul(ctr[c], rtg[i]), pts)
            grid = gt_intr[c].reshape(5, 5, 3)
            z = np.where(lp[:, 2] > 1e-6, lp[:, 2], 1.0)
            init = np.stack([(H / 2.0) * lp[:, 0] / z + W / 2.0, (H / 2.0) * lp[:, 1] / z + H / 2.0], -1)
            xy, ok = synthetic.central_project_np(cams[c], grid, lp, init)
            ok &= lp[:, 2] > 1e-6
            idx = np.nonzero(ok)[0]
            oi.append(np.full(len(idx), i, np.uint32))
            oc.append(np.full(len(idx), c, np.uint32))
            op.append(idx.astype(np.uint32))
            oxy.append(xy[idx].astype(np.float32))
    problem = FlatProblem(cams, n_poses, n_points, np.concatenate(oi), np.concatenate(oc), np.concatenate(op),
                          np.concatenate(oxy))
    gt = FlatState(pts.copy(), rtg.copy(), ctr.copy(), [a.copy() for a in gt_intr], np.zeros((problem.n_obs, 2)))
    st = gt.copy()
    st.points += 0.05 * U(n_points, 3)
    for i in range(n_poses):
        st.rig_tr_global[i] = synthetic.pose_mul(st.rig_tr_global[i], synthetic.se3_exp(0.04 * U(6)))
    if num_cameras > 1:
        for c in range(num_cameras):
            st.camera_tr_rig[c] = synthetic.pose_mul(st.camera_tr_rig[c], synthetic.se3_exp(0.04 * U(6)))
    for c in range(num_cameras):
        g = st.intrinsics[c].reshape(-1, 3) + 0.02 * U(25, 3)
        st.intrinsics[c] = (g / np.linalg.norm(g, axis=-1, keepdims=True)).reshape(-1)
    return problem, st, gt


def reference_noncentral_ba_test_problem(seed=0):
    """NoncentralGenericBSpline.OptimizeJointly (test/noncentral_generic_test.cc:114-258)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    U = lambda *s: rng.uniform(-1, 1, size=s)
    W, H = 640, 480
    cam = make_camera(cabi.MODEL_NONCENTRAL_GENERIC, W, H, (0, 0, W - 1, H - 1), 4, 4)
    pg = np.zeros((4, 4, 3))
    dg = np.zeros((4, 4, 3))
    for y in range(4):
        for x in range(4):
            pg[y, x] = (-1.5 + x, -1.5 + y, 0)
            v = np.array([0, 0.05 * x, 1.0])
            dg[y, x] = v / np.linalg.norm(v)
    intr = np.concatenate([dg.reshape(-1), pg.reshape(-1)])
    n_points, n_poses = 50, 20
    pts = 0.3 * U(n_points, 3)
    rtg = np.zeros((n_poses, 7))
    oi, op, oxy = [], [], []
    for i in range(n_poses):
        base = synthetic.IDENTITY_POSE.copy()
        base[4:] = (0, 0, 1.0)
        rtg[i] = synthetic.pose_mul(synthetic.se3_exp(0.05 * U(6)), base)
        lp = synthetic.pose_apply(rtg[i], pts)
        init = np.tile(np.array([W / 2.0, H / 2.0]), (n_points, 1))
        # orthographic-like camera: start from the affine guess
        init = np.stack([(lp[:, 0] + 0.5) * W, (lp[:, 1] + 0.5) * H], -1)
        xy, ok = synthetic.noncentral_project_np(cam, dg, pg, lp, init, iters=60)
        idx = np.nonzero(ok)[0]
        oi.append(np.full(len(idx), i, np.uint32))
        op.append(idx.astype(np.uint32))
        oxy.append(xy[idx].astype(np.float32))
    n = sum(len(a) for a in oi)
    problem = FlatProblem([cam], n_poses, n_points, np.concatenate(oi), np.zeros(n, np.uint32), np.concatenate(op),
                          np.concatenate(oxy))
    st = FlatState(pts + 0.05 * U(n_points, 3), rtg.copy(), synthetic.IDENTITY_POSE[None].copy(), [intr.copy()],
                   np.zeros((n, 2)))
    for i in range(1, n_poses):
        st.rig_tr_global[i] = synthetic.pose_mul(st.rig_tr_global[i], synthetic.se3_exp(0.04 * U(6)))
    return problem, st
