"""Pins the CPU oracle to the reference through the golden vectors, known answers and
thresholds of the reference's OWN tests. CPU only.

Paths are relative to /root/reference; APP = applications/camera_calibration/src/camera_calibration,
LV = libvis/src/libvis.
"""
import numpy as np
import pytest

from camera_calibration_b200 import cabi, synthetic
from tests import helpers


def test_schur_known_answer(oracle_lib):
    """LV/test/lm_optimizer.cc:471-557: 6x6 system, lower triangles are NaN to prove that only the
    upper triangle is read; expected x within 0.3."""
    nan = np.nan
    D = np.array([[[1, 5], [nan, 6]], [[9, 5], [nan, 4]]], dtype=float)
    B = np.array([[3, 4], [7, 8], [7, 6], [3, 2]], dtype=float)
    Cm = np.array([[1, 4], [nan, 7]], dtype=float)
    x = oracle_lib.schur_solve(2, D, B, Cm, [1, 2, 3, 4], [5, 6])
    expected = np.array([73.667, 171.667, 189.667, -294.333, 465.667, -582.0])
    assert np.all(np.abs(x - expected) < 0.3)
    # and to full precision against numpy on the symmetrised matrix
    H = np.array([[1, 5, 0, 0, 3, 4], [5, 6, 0, 0, 7, 8], [0, 0, 9, 5, 7, 6], [0, 0, 5, 4, 3, 2],
                  [3, 7, 7, 3, 1, 4], [4, 8, 6, 2, 4, 7]], dtype=float)
    assert np.allclose(x, np.linalg.solve(H, np.arange(1.0, 7.0)), rtol=1e-10)


def test_schur_equals_dense_solve(oracle_lib):
    """LV/test/lm_optimizer.cc:370-468: solving with and without the Schur complement agrees."""
    rng = np.random.default_rng(0)
    nb, bs, nd = 20, 2, 18
    n = nb * bs + nd
    J = rng.standard_normal((3 * n, n))
    H = J.T @ J
    for i in range(nb):  # make the top-left part block diagonal
        for j in range(nb):
            if i != j:
                H[i * bs:(i + 1) * bs, j * bs:(j + 1) * bs] = 0
    H += 0.1 * np.eye(n)
    b = rng.standard_normal(n)
    D = np.stack([H[i * bs:(i + 1) * bs, i * bs:(i + 1) * bs] for i in range(nb)])
    x = oracle_lib.schur_solve(bs, D, H[:nb * bs, nb * bs:], np.triu(H[nb * bs:, nb * bs:]), b[:nb * bs], b[nb * bs:])
    xd = oracle_lib.solve_dense(np.triu(H), b)
    ref = np.linalg.solve(H, b)
    assert np.allclose(x, ref, rtol=1e-8, atol=1e-10)
    assert np.allclose(xd, ref, rtol=1e-8, atol=1e-10)


def test_noncentral_orthographic_known_answers(oracle_lib):
    """APP/test/noncentral_generic_test.cc:49-112."""
    cam, intr = helpers.orthographic_noncentral()
    d, o, ok = oracle_lib.unproject(cam, intr, [[50.0, 50.0]])
    assert ok[0]
    assert abs(o[0, 0] - 1.5) < 1e-5 and abs(o[0, 1] - 1.5) < 1e-5
    assert abs(d[0, 0]) < 1e-5 and abs(d[0, 1]) < 1e-5 and abs(abs(d[0, 2]) - 1) < 1e-5
    px, ok = oracle_lib.project(cam, intr, [[1.5, 1.5, 42.12345]])
    assert ok[0] and abs(px[0, 0] - 50) < 1e-5 and abs(px[0, 1] - 50) < 1e-5
    for gx, gy in ((1.1, 1.2), (1.001, 1.999)):
        px, ok = oracle_lib.project(cam, intr, [[gx, gy, 42.12345]])
        ex, ey = synthetic.grid_point_to_pixel(cam, gx, gy)
        assert ok[0]
        assert abs(px[0, 0] - ex) < 1e-5 and abs(px[0, 1] - ey) < 1e-5


def test_central_project_unproject(oracle_lib):
    """APP/test/util.h:112-164 (TestProjectUnproject): 8x6 grid normalize(x, y, 1), 640x480 with
    the calibrated area (10, 20)-(635, 472); 400 pixels: Unproject == UnprojectWithJacobian to 1e-5,
    re-projection to 1e-4 px. Pixels from a seeded PRNG instead of Eigen::Random."""
    cam = helpers.make_camera(cabi.MODEL_CENTRAL_GENERIC, 640, 480, (10, 20, 635, 472), 8, 6)
    grid = helpers.xy1_grid(8, 6)
    rng = np.random.default_rng(0)
    px = np.array([10.0, 20.0]) + np.abs(rng.uniform(-1, 1, (400, 2))) * np.array([640 - 14, 480 - 27])
    d1, _, ok1 = oracle_lib.unproject(cam, grid, px)
    d2, _, J, ok2 = oracle_lib.unproject(cam, grid, px, with_jacobian=True)
    assert ok1.all() and ok2.all()
    assert np.abs(d1 - d2).max() < 1e-5
    rp, ok3 = oracle_lib.project(cam, grid, d1)
    assert ok3.all()
    assert np.abs(rp - px).max() < 1e-4
    # the analytic Jacobian is the derivative of Unproject
    h = 1e-6
    dx, _, _ = oracle_lib.unproject(cam, grid, px + [h, 0])
    dy, _, _ = oracle_lib.unproject(cam, grid, px + [0, h])
    assert np.abs((dx - d1) / h - J[:, :, 0]).max() < 1e-6
    assert np.abs((dy - d1) / h - J[:, :, 1]).max() < 1e-6


def test_generic_models_reprojection_lattice(oracle_lib):
    """generic_models/src/main.cc:38-84: 8x8 grid over the full 640x480 image, 10-px lattice of
    pixel centres; un-project then project returns the pixel to 1e-3."""
    cam = helpers.make_camera(cabi.MODEL_CENTRAL_GENERIC, 640, 480, (0, 0, 639, 479), 8, 8)
    grid = helpers.xy1_grid(8, 8)
    xs, ys = np.meshgrid(np.arange(0, 640, 10) + 0.5, np.arange(0, 480, 10) + 0.5)
    px = np.stack([xs.ravel(), ys.ravel()], -1)
    d, _, ok = oracle_lib.unproject(cam, grid, px)
    assert ok.all()
    rp, ok2 = oracle_lib.project(cam, grid, d)
    assert ok2.all()
    assert np.linalg.norm(rp - px, axis=1).max() < 1e-3


def test_real_camera_roundtrip(oracle_lib):
    """The real 17x13 calibrated grid of generic_models/src/main.cc:87-97 (fixture
    tests/golden/real_central_17x13.json): round trip to 1e-3 px over the calibrated area, and
    the last grid value the reference test checks (main.cc:133)."""
    cam, grid = helpers.real_camera()
    assert abs(grid[-1, -1, 2] - 0.67986719656337) < 1e-3
    xs, ys = np.meshgrid(np.arange(15, 625, 7) + 0.5, np.arange(16, 465, 7) + 0.5)
    px = np.stack([xs.ravel(), ys.ravel()], -1)
    d, _, ok = oracle_lib.unproject(cam, grid, px)
    assert ok.all()
    rp, ok2 = oracle_lib.project(cam, grid, d)
    assert ok2.all()
    assert np.linalg.norm(rp - px, axis=1).max() < 1e-3
    # and the independent numpy implementation used for data generation agrees with the oracle
    dn = synthetic.central_unproject_np(cam, grid, px[:, 0], px[:, 1])
    assert np.abs(dn - d).max() < 1e-12


def test_bspline_slow_fast(oracle_lib):
    """APP/test/b_spline_test.cc:41-58: 4x4 control points, 500 samples along x at y = 1.5, 1e-5."""
    vals = [0, 0, 0, 0, 0, 1, 2, 3, 0, 4, 5, 6, 0, 7, 8, 9]
    grid = np.zeros((4, 4, 3))
    for i, v in enumerate(vals):
        grid[i // 4, i % 4] = (v, v, 0)
    for i in range(500):
        a = oracle_lib.bspline_eval(grid, 1.0 + i / 500.0, 1.5, slow=True)
        b = oracle_lib.bspline_eval(grid, 1.0 + i / 500.0, 1.5, slow=False)
        assert np.abs(a - b).max() < 1e-5


def test_huber_identities(oracle_lib):
    """LV/test/loss_functions.cc (HuberLoss<double>(1.4)): Compute*() == Compute*FromSquaredResidual()."""
    l = oracle_lib.lib()
    for r in (-2.0, -1.0, 0.0, 1.0, 2.0):
        assert abs(l.oracle_huber_cost(1.4, r) - l.oracle_huber_cost_sq(1.4, r * r)) < 1e-8
        assert abs(l.oracle_huber_weight(1.4, r) - l.oracle_huber_weight_sq(1.4, r * r)) < 1e-8
    assert l.oracle_huber_cost_sq(1.0, 0.25) == 0.125
    assert abs(l.oracle_huber_cost_sq(1.0, 4.0) - 1.5) < 1e-15
    assert l.oracle_huber_weight_sq(1.0, 4.0) == 0.5


@pytest.mark.parametrize("num_cameras,eliminate_points", [(1, 0), (1, 1), (2, 0)])
def test_central_generic_ba_reaches_reference_threshold(oracle_lib, num_cameras, eliminate_points):
    """APP/test/central_generic_test.cc:60-66 -> test/util.h:275-571: <= 20*C single-iteration
    OptimizeJointly calls (numeric Jacobian, delta 1e-4) must reach cost <= C * 1e-6. The
    reference runs eliminate_points=false; eliminate_points=true (the north-star mode, which no
    reference test exercises) must reach the same bar."""
    problem, st, _ = helpers.reference_ba_test_problem(num_cameras=num_cameras, n_points=60, n_poses=40)
    opt = cabi.default_options(max_iteration_count=1, eliminate_points=eliminate_points,
                               jacobian_mode=cabi.JACOBIAN_NUMERIC)
    lam = -1.0
    cost = np.inf
    for _ in range(20 * num_cameras):
        opt.init_lambda = lam
        st, rep = oracle_lib.optimize(problem, st, opt)
        lam = rep.final_lambda
        cost = rep.final_cost
        if not rep.performed_an_iteration:
            break
    assert cost <= num_cameras * 1e-6


def test_noncentral_ba_reaches_reference_threshold(oracle_lib):
    """APP/test/noncentral_generic_test.cc:114-258: max_iteration_count 50, delta 1e-3,
    eliminate_points=false: final cost <= 2e-4."""
    problem, st = helpers.reference_noncentral_ba_test_problem()
    opt = cabi.default_options(max_iteration_count=50, numerical_diff_delta=1e-3, eliminate_points=0,
                               jacobian_mode=cabi.JACOBIAN_NUMERIC)
    st, rep = oracle_lib.optimize(problem, st, opt)
    assert rep.final_cost <= 2e-4


def test_numeric_and_analytic_jacobians_agree(oracle_lib):
    """The analytic (implicit-function) Jacobian the GPU path uses equals the reference's numeric
    one up to the finite-difference truncation."""
    for cfg, size, lat in ((2, (410, 290), (10, 8)), (3, (300, 240), (8, 6)), (1, None, (8, 8))):
        sp = synthetic.make_problem(cfg, n_imagesets=6, lattice=lat, image_size=size)
        opt = cabi.default_options()
        opt.jacobian_mode = cabi.JACOBIAN_NUMERIC
        en = oracle_lib.evaluate(sp.problem, sp.init_state, opt, True)
        opt.jacobian_mode = cabi.JACOBIAN_ANALYTIC
        ea = oracle_lib.evaluate(sp.problem, sp.init_state, opt, True)
        both = (en["has_jacobian"] == 1) & (ea["has_jacobian"] == 1)
        assert both.sum() > 0.9 * sp.n_obs
        assert np.array_equal(en["intr_index"][both], ea["intr_index"][both])
        for k in ("j_point", "j_pose", "j_intr"):
            a, b = en[k][both], ea[k][both]
            assert np.abs(a - b).max() < 2e-3 * np.abs(b).max(), (cfg, k)
        assert np.abs(en["residuals"][both] - ea["residuals"][both]).max() < 1e-9


def test_fast_dense_schur_solve_agrees_with_plain(oracle_lib):
    """The blocked, OpenMP-threaded dense kernels used at full size (oracle/ba_dense_fast.h: 4x3
    register-tile product, blocked LDLT with the pivot sequence replayed up front) compute what the plain
    loops compute: random block-arrow SPD systems, a diagonal that forces pivoting, 1 and 4 threads."""
    import ctypes as C
    rng = np.random.default_rng(5)
    bs, nb, nd = 3, 60, 333
    n = bs * nb + nd
    J = rng.standard_normal((2 * n, n))
    H = J.T @ J + 0.1 * np.eye(n)
    for i in range(nb):
        for j in range(nb):
            if i != j:
                H[i * bs:(i + 1) * bs, j * bs:(j + 1) * bs] = 0
    H[bs * nb:, bs * nb:] += np.diag(rng.uniform(0, 500, nd))  # unsorted diagonal: the pivot order is not the identity
    H += 40 * np.eye(n)
    D = np.stack([np.triu(H[i * bs:(i + 1) * bs, i * bs:(i + 1) * bs]) for i in range(nb)])
    B = np.ascontiguousarray(H[:bs * nb, bs * nb:])
    Cm = np.ascontiguousarray(np.triu(H[bs * nb:, bs * nb:]))
    b = rng.standard_normal(n)
    x0 = oracle_lib.schur_solve(bs, D, B, Cm, b[:bs * nb], b[bs * nb:])
    p = lambda a: np.ascontiguousarray(a).ctypes.data_as(C.POINTER(C.c_double))
    for threads in (1, 4):
        oracle_lib.set_threads(threads)
        try:
            x1 = np.zeros(n)
            b1, b2 = np.ascontiguousarray(b[:bs * nb]), np.ascontiguousarray(b[bs * nb:])
            Dc = np.ascontiguousarray(D)
            oracle_lib.lib().oracle_schur_solve_fast(bs, nb, nd, p(Dc), p(B), p(Cm), p(b1), p(b2), p(x1))
        finally:
            oracle_lib.set_threads(1)
        assert np.abs(x1 - x0).max() < 1e-10 * max(1.0, np.abs(x0).max())
    assert np.abs(H @ x0 - b).max() < 1e-8 * np.abs(b).max() * n


def test_threaded_evaluation_is_bitwise_the_sequential_one(oracle_lib):
    """compute_mt (per-observation work in parallel, accumulation replayed in the reference's order): H, b,
    cost and the LM trajectory are bitwise those of the single-threaded pass, NUMERIC and ANALYTIC."""
    from camera_calibration_b200 import cabi, synthetic
    sp = synthetic.make_problem(2, n_imagesets=6, lattice=(8, 6), image_size=(300, 220))
    for mode in (cabi.JACOBIAN_NUMERIC, cabi.JACOBIAN_ANALYTIC):
        opt = cabi.default_options(jacobian_mode=mode, max_iteration_count=2)
        H0, b0, c0 = oracle_lib.build_system(sp.problem, sp.init_state, opt)
        s0, r0 = oracle_lib.optimize(sp.problem, sp.init_state, opt)
        oracle_lib.set_threads(4)
        try:
            H1, b1, c1 = oracle_lib.build_system(sp.problem, sp.init_state, opt)
            s1, r1 = oracle_lib.optimize(sp.problem, sp.init_state, opt)
        finally:
            oracle_lib.set_threads(1)
        assert np.array_equal(H0, H1) and np.array_equal(b0, b1) and c0 == c1
        assert r0.trace() == r1.trace()
        assert np.array_equal(s0.points, s1.points)
