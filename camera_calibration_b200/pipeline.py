"""The callers either side of the hot path: the product's outer bundle-adjustment
loop ``RunBundleAdjustment`` (applications/camera_calibration/src/camera_calibration/calibration.cc:187-304)
with its per-iteration checkpoint and ``ChooseNiceCameraOrientation``
(models/central_generic.cc:570-621), the outlier deletion between BA rounds
(``DeleteOutlierFeatures``, calibration.cc:62-184) and ``ScaleToMetric``
(calibration.cc:307-370; 8f-4) and the pyramid step ``ResampleModel`` (calibration.cc:373-522; 8f-4). Host logic only; every numerical step (un-projection, the LM
iteration) runs in ``libb200ba.so``.
"""
from __future__ import annotations

import math
from typing import Callable, List, Optional

import numpy as np

from . import api, synthetic
from .api import BAState, CameraModel, CentralGenericModel, Dataset, SchurMode


def _from_two_vectors(a: np.ndarray, b: np.ndarray) -> np.ndarray:
    """Eigen::Quaterniond::FromTwoVectors(a, b) as a rotation matrix (rotates a onto b)."""
    v0 = a / np.linalg.norm(a)
    v1 = b / np.linalg.norm(b)
    c = float(v1 @ v0)
    if c < -1.0 + 1e-12:  # opposite vectors: any perpendicular axis (Eigen uses an SVD here)
        axis = np.cross(v0, [1.0, 0, 0])
        if np.linalg.norm(axis) < 1e-6:
            axis = np.cross(v0, [0, 1.0, 0])
        axis /= np.linalg.norm(axis)
        q = np.array([0.0, *axis])
    else:
        axis = np.cross(v0, v1)
        s = math.sqrt((1.0 + c) * 2.0)
        q = np.array([0.5 * s, *(axis / s)])
    return synthetic.quat_to_rot(q)


def _rot_to_quat(R: np.ndarray) -> np.ndarray:
    """Rotation matrix -> unit quaternion (w, x, y, z)."""
    t = np.trace(R)
    if t > 0:
        s = math.sqrt(t + 1.0) * 2
        q = np.array([0.25 * s, (R[2, 1] - R[1, 2]) / s, (R[0, 2] - R[2, 0]) / s, (R[1, 0] - R[0, 1]) / s])
    else:
        i = int(np.argmax(np.diag(R)))
        j, k = (i + 1) % 3, (i + 2) % 3
        s = math.sqrt(R[i, i] - R[j, j] - R[k, k] + 1.0) * 2
        q = np.zeros(4)
        q[0] = (R[k, j] - R[j, k]) / s
        q[1 + i] = 0.25 * s
        q[1 + j] = (R[j, i] + R[i, j]) / s
        q[1 + k] = (R[k, i] + R[i, k]) / s
    return q / np.linalg.norm(q)


def ChooseNiceCameraOrientation(model: CameraModel) -> np.ndarray:
    """Rotates the model such that +z looks forward at the image centre and +x points right;
    returns the applied rotation (to be left-multiplied onto camera_tr_rig). Only the central
    generic model implements it in the reference (central_generic.cc:570-621); the base class and
    the non-central model return the identity (camera_model.h:120-122, noncentral_generic.h:128-132)."""
    if not isinstance(model, CentralGenericModel):
        return np.eye(3)
    w, h = model.width(), model.height()
    ok, forward, _ = model.Unproject(0.5 * w, 0.5 * h)
    if not ok:
        forward = np.array([0.0, 0, 1.0])
    forward_rotation = _from_two_vectors(forward, np.array([0.0, 0, 1.0]))
    right_min_x, right_max_x = min(w - 1, w // 2 + 11), w - 1
    right_min_y, right_max_y = max(0, h // 2 - 10), min(h - 1, h // 2 + 10)
    xs, ys = np.meshgrid(np.arange(right_min_x, right_max_x + 1), np.arange(right_min_y, right_max_y + 1))
    px = np.stack([xs.ravel() + 0.5, ys.ravel() + 0.5], -1)
    right_rotation = np.eye(3)
    if len(px):
        d, _, okm = model.UnprojectMany(px)
        if okm.any():
            frr = forward_rotation @ d[okm].mean(axis=0)
            angle = math.atan2(-frr[1], frr[0])
            c, s = math.cos(angle), math.sin(angle)
            right_rotation = np.array([[c, -s, 0], [s, c, 0], [0, 0, 1.0]])
    rotation = right_rotation @ forward_rotation
    model.m_grid = model.m_grid @ rotation.T  # Rotate(): every direction d -> rotation * d
    return rotation


def RunBundleAdjustment(use_cuda: bool, schur_mode: SchurMode, max_iteration_count: int,
                        cost_reduction_threshold: float, dataset: Dataset, state: BAState,
                        regularization_weight: float, localize_only: bool,
                        state_output_path: Optional[str] = None, eliminate_points: bool = False,
                        on_iteration: Optional[Callable[[int, float], None]] = None,
                        device_resident: bool = True) -> List[float]:
    """calibration.cc:187-304: single LM iterations until the cost stops falling by more than
    ``cost_reduction_threshold``; the state directory is rewritten after every iteration (the
    reference's checkpoint / resume mechanism) and the cameras are re-oriented. Returns the cost
    after each iteration. ``eliminate_points`` defaults to the product's choice (False).

    ``device_resident`` (default): the whole loop -- LM iterations, ChooseNiceCameraOrientation, stop
    rule -- runs inside the library on the state in HBM (``b200ba_run_bundle_adjustment``); the state
    comes back to the host only for the per-iteration checkpoint (when ``state_output_path`` is given)
    and at the end. ``device_resident=False`` is the same loop driven from Python with one host round
    trip per iteration (kept for the equivalence test)."""
    if device_resident:
        def cb(it, cost, sync):
            if state_output_path:
                from . import io
                sync()
                io.SaveBAState(state_output_path, state)
            if on_iteration:
                on_iteration(it, cost)
            return False
        rep = api.RunBundleAdjustmentOnDevice(dataset, state, max_iteration_count, cost_reduction_threshold,
                                              regularization_weight, localize_only, eliminate_points, schur_mode,
                                              cb if (state_output_path or on_iteration) else None)
        return [float(rep.costs[i]) for i in range(rep.iterations)]
    numerical_diff_delta = 1e-4  # calibration.cc:201
    lam = -1.0
    last_cost = math.inf
    costs: List[float] = []
    for iteration in range(max_iteration_count):
        if use_cuda:
            report, lam = api.CudaOptimizeJointly(dataset, state, 1, 50, lam, numerical_diff_delta,
                                                  regularization_weight, print_progress=False)
            cost = report.final_cost
        else:
            cost, lam, _ = api.OptimizeJointly(dataset, state, 1, lam, numerical_diff_delta, regularization_weight,
                                               localize_only, eliminate_points, schur_mode, print_progress=False)
        costs.append(cost)
        if state_output_path:
            from . import io
            io.SaveBAState(state_output_path, state)
        if not localize_only:
            for c in range(state.num_cameras()):
                rotation = ChooseNiceCameraOrientation(state.intrinsics[c])
                rt = np.concatenate([_rot_to_quat(rotation), np.zeros(3)])
                state.camera_tr_rig[c] = synthetic.pose_mul(rt, state.camera_tr_rig[c])
        if on_iteration:
            on_iteration(iteration, cost)
        if cost >= last_cost - cost_reduction_threshold:
            break
        last_cost = cost
    return costs


def DeleteOutlierFeatures(camera_index: int, dataset: Dataset, state: BAState, outlier_removal_factor: float,
                          project_many: Optional[Callable] = None) -> int:
    """calibration.cc:62-184 (the quartile rule between BA rounds): re-project every feature of one
    camera from the centre of the calibrated area (``Project``, no warm start), take the first and
    third quartile q1, q3 of the error magnitudes, and erase the features that fail to project or
    whose error exceeds ``q3 + outlier_removal_factor (q3 - q1)``. Imagesets left with fewer than
    three features of this camera are marked unused. Returns the number of removed features.
    The projections run on the device (``CameraModel.ProjectMany`` -> ``b200ba_project``);
    ``project_many(model, local_points) -> (pixels, ok)`` may replace it (host-logic tests)."""
    model = state.intrinsics[camera_index]
    if project_many is None:
        project_many = lambda m, lp: m.ProjectMany(lp)  # noqa: E731
    spans = []  # (imageset, first, last) into the concatenated feature list
    chunks = []
    n = 0
    for i in range(dataset.ImagesetCount()):
        if not state.image_used[i]:
            continue
        f = dataset.GetImageset(i).FeaturesOfCamera(camera_index)
        T = state.image_tr_global(camera_index, i)
        R = synthetic.quat_to_rot(T[:4])
        chunks.append(state.points[f["index"]] @ R.T + T[4:7])
        spans.append((i, n, n + len(f["index"])))
        n += len(f["index"])
    if n == 0:
        return 0
    pixels, ok = project_many(model, np.concatenate(chunks))
    xy = np.concatenate([dataset.GetImageset(i).FeaturesOfCamera(camera_index)["xy"] for i, _, _ in spans])
    err = np.linalg.norm(pixels - xy.astype(np.float64), axis=1)
    errors = np.sort(err[ok])
    if len(errors) < 8:  # too few to detect outliers reliably (calibration.cc:97-100)
        return 0
    # index = float(0.25f * size + 0.5f) truncated, in float arithmetic like the reference
    first_quartile = errors[int(np.float32(0.25) * np.float32(len(errors)) + np.float32(0.5))]
    third_quartile = errors[int(np.float32(0.75) * np.float32(len(errors)) + np.float32(0.5))]
    threshold = third_quartile + np.float32(outlier_removal_factor) * (third_quartile - first_quartile)
    remove = ~ok | (err > threshold)
    removed = 0
    for i, a, b in spans:
        f = dataset.GetImageset(i).FeaturesOfCamera(camera_index)
        keep = ~remove[a:b]
        if not keep.all():
            for key in list(f.keys()):
                f[key] = f[key][keep]
            removed += int((~keep).sum())
        if int(keep.sum()) < 3:
            state.image_used[i] = False
    if removed:
        dataset._b200_context = None
    return removed


def ScaleToMetric(dataset: Dataset, state: BAState) -> float:
    """calibration.cc:307-370: geometric-mean ratio of the known pattern cell length to the
    optimised distance of neighbouring corners (right and down neighbours), applied with
    ``BAState.ScaleState``. Returns the factor."""
    log_sum, count = 0.0, 0
    for geometry in getattr(dataset, "known_geometries", []):
        position_to_index = {}
        for feature_id, position in geometry.feature_id_to_position.items():
            idx = state.feature_id_to_points_index.get(feature_id)
            if idx is not None:
                position_to_index[tuple(position)] = idx
        if not position_to_index:
            continue
        for feature_id, position in geometry.feature_id_to_position.items():
            index = position_to_index.get(tuple(position))
            if index is None:
                continue
            for dx, dy in ((1, 0), (0, 1)):
                neighbor = position_to_index.get((position[0] + dx, position[1] + dy))
                if neighbor is None:
                    continue
                actual = float(np.linalg.norm(state.points[index] - state.points[neighbor]))
                log_sum += math.log(geometry.cell_length_in_meters / actual)
                count += 1
    if count == 0:
        raise ValueError("ScaleToMetric: no neighbouring corners with known geometry (the reference divides by zero here)")
    factor = math.exp(log_sum / count)
    state.ScaleState(factor)
    return factor


def _interpolate_bilinear(image: np.ndarray, x: float, y: float) -> np.ndarray:
    """libvis Image::InterpolateBilinear for vector pixels (libvis/image.h:152-176): integer part by
    truncation, FLOAT weights, double accumulation."""
    ix, iy = int(x), int(y)
    fx = np.float32(x - ix)
    fy = np.float32(y - iy)
    fx_inv = np.float32(1) - fx
    fy_inv = np.float32(1) - fy
    return (float(fx_inv * fy_inv) * image[iy, ix] + float(fx * fy_inv) * image[iy, ix + 1]
            + float(fx_inv * fy) * image[iy + 1, ix] + float(fx * fy) * image[iy + 1, ix + 1])


def ResampleModel(model_to_optimize: CameraModel, camera_tr_rig: np.ndarray, calibration_min_x: int,
                  calibration_min_y: int, calibration_max_x: int, calibration_max_y: int,
                  model_type: CameraModel.Type, target_resolution_x: int, target_resolution_y: int,
                  fit_fn=None, unproject_many=None):
    """calibration.cc:373-522 for the generic target models (the radial / thin-prism / OpenCV
    targets are outside this path): returns ``(ok, new_model)``; ``camera_tr_rig`` is untouched for
    these targets (only the parametric fits rotate it).
      * NoncentralGeneric -> NoncentralGeneric: both grids re-sampled bilinearly (:386-424);
      * otherwise a dense direction image of the old model (one ``Unproject`` per pixel centre, on
        the device) is fitted by a CentralGenericModel of the target resolution
        (``FitToDenseModel(dense, step, 3)``; at most 300 x 300 samples), optionally wrapped into a
        NoncentralGenericModel with zero origins."""
    T = CameraModel.Type
    model_type = T(model_type)
    if model_to_optimize.type() == T.NoncentralGeneric and model_type == T.NoncentralGeneric:
        old = model_to_optimize
        ogh, ogw = old.direction_grid().shape[:2]
        new_points = np.zeros((target_resolution_y, target_resolution_x, 3))
        new_dirs = np.zeros((target_resolution_y, target_resolution_x, 3))
        for y in range(target_resolution_y):
            for x in range(target_resolution_x):
                pixel = CentralGenericModel.GridPointToPixelCornerConvStatic(
                    x, y, calibration_min_x, calibration_min_y, calibration_max_x, calibration_max_y,
                    target_resolution_x, target_resolution_y)
                g = old.PixelCornerConvToGridPoint(pixel[0], pixel[1])
                g = np.minimum(np.maximum(g, 0.0), np.array([ogw - 1.001, ogh - 1.001]))
                new_points[y, x] = _interpolate_bilinear(old.point_grid(), g[0], g[1])
                new_dirs[y, x] = _interpolate_bilinear(old.direction_grid(), g[0], g[1])
        new = api.NoncentralGenericModel(target_resolution_x, target_resolution_y, calibration_min_x, calibration_min_y,
                                         calibration_max_x, calibration_max_y, old.width(), old.height())
        new.SetPointGrid(new_points)
        new.SetDirectionGrid(new_dirs)
        return True, new
    if model_to_optimize.type() == T.NoncentralGeneric:
        return False, model_to_optimize  # not implemented in the reference either (:426-429)
    if model_type not in (T.CentralGeneric, T.NoncentralGeneric):
        return False, model_to_optimize  # parametric targets: outside this path
    # dense direction model of the old camera
    w, h = model_to_optimize.width(), model_to_optimize.height()
    xs, ys = np.meshgrid(np.arange(w) + 0.5, np.arange(h) + 0.5)
    pixels = np.stack([xs.ravel(), ys.ravel()], -1)
    if unproject_many is None:
        unproject_many = lambda m, px: m.UnprojectMany(px)  # noqa: E731
    dirs, _, ok = unproject_many(model_to_optimize, pixels)
    dense = np.where(ok[:, None], dirs, np.nan).reshape(h, w, 3)
    area_w = calibration_max_x - calibration_min_x + 1
    area_h = calibration_max_y - calibration_min_y + 1
    # std::round(int / int): the integer quotient is already integral
    subsample_step = max(1, min(area_w // 300, area_h // 300))
    new_central = CentralGenericModel(target_resolution_x, target_resolution_y, calibration_min_x, calibration_min_y,
                                      calibration_max_x, calibration_max_y, w, h)
    if not new_central.FitToDenseModel(dense, subsample_step, 3, fit_fn=fit_fn):
        return False, model_to_optimize
    if model_type == T.NoncentralGeneric:
        new = api.NoncentralGenericModel(target_resolution_x, target_resolution_y, calibration_min_x, calibration_min_y,
                                         calibration_max_x, calibration_max_y, w, h)
        new.InitializeFromCentralGenericModel(new_central)
        return True, new
    return True, new_central


def ComputeGridResolution(calibration_area_width: int, calibration_area_height: int, exterior_cells_per_side: int,
                          approx_pixels_per_cell: int):
    """calibration.cc:531-540: integer division, + 0.5f, + the exterior cells, truncated."""
    rx = int(np.float32(calibration_area_width // approx_pixels_per_cell) + np.float32(0.5) + np.float32(2 * exterior_cells_per_side))
    ry = int(np.float32(calibration_area_height // approx_pixels_per_cell) + np.float32(0.5) + np.float32(2 * exterior_cells_per_side))
    return rx, ry


def ComputeGridResolutionForModel(model: CameraModel, approx_pixels_per_cell: int):
    """calibration.cc:542-560."""
    w = model.calibration_max_x() - model.calibration_min_x() + 1
    h = model.calibration_max_y() - model.calibration_min_y() + 1
    exterior = model.exterior_cells_per_side() if hasattr(model, "exterior_cells_per_side") else 0
    return ComputeGridResolution(w, h, exterior, approx_pixels_per_cell)


def CalcGridResolutionForLevel(pyramid_level: int, full_resolution_x: int, full_resolution_y: int):
    """calibration.cc:566-569."""
    f = math.pow(1.333, -pyramid_level)
    return int(full_resolution_x * f + float(np.float32(0.5))), int(full_resolution_y * f + float(np.float32(0.5)))


def ComputeIntegerBoundingRectForFeatures(dataset: Dataset, camera_index: int, image_used):
    """calibration.cc:615-641: (min_x, min_y, max_x, max_y) of the truncated feature positions."""
    min_x = min_y = np.iinfo(np.int32).max
    max_x = max_y = 0
    for i in range(dataset.ImagesetCount()):
        if not image_used[i]:
            continue
        xy = dataset.GetImageset(i).FeaturesOfCamera(camera_index)["xy"]
        if len(xy) == 0:
            continue
        t = xy.astype(np.int64)  # static_cast<int>: truncation
        min_x, min_y = min(min_x, int(t[:, 0].min())), min(min_y, int(t[:, 1].min()))
        max_x, max_y = max(max_x, int(t[:, 0].max())), max(max_y, int(t[:, 1].max()))
    return min_x, min_y, max_x, max_y


def ResampleModelsIfNecessary(dataset: Dataset, state: BAState, model_type: CameraModel.Type,
                              approx_pixels_per_cell: int, pyramid_level: int, fit_fn=None, unproject_many=None) -> int:
    """calibration.cc:572-612: re-sample every camera whose grid resolution differs from the one
    wanted on this pyramid level, or whose type differs. Returns the number of re-sampled models."""
    count = 0
    for c in range(dataset.num_cameras()):
        model = state.intrinsics[c]
        loaded = model.GetGridResolution()
        fx, fy = ComputeGridResolutionForModel(model, approx_pixels_per_cell)
        dx, dy = CalcGridResolutionForLevel(pyramid_level, fx, fy)
        if (loaded and tuple(loaded) != (dx, dy)) or model.type() != CameraModel.Type(model_type):
            ok, new = ResampleModel(model, state.camera_tr_rig[c], model.calibration_min_x(), model.calibration_min_y(),
                                    model.calibration_max_x(), model.calibration_max_y(), model_type, dx, dy,
                                    fit_fn=fit_fn, unproject_many=unproject_many)
            if ok:
                state.intrinsics[c] = new
                count += 1
    return count


def BundleAdjustment(state_directory: str, model_input_directory: str, model_output_directory: str,
                     max_iteration_count: int = 30) -> int:
    """The ``--bundle_adjustment`` tool (tools/bundle_adjustment.cc:50-220): load
    ``intrinsics0.yaml``, read a COLMAP text model, run <= 30 single LM iterations with
    ``localize_only=true, eliminate_points=true`` and write the state directory + ``cost.txt``
    (14 significant digits) after every iteration. Returns EXIT_SUCCESS / EXIT_FAILURE."""
    import os
    from . import io
    model = io.LoadCameraModel(os.path.join(state_directory, "intrinsics0.yaml"))
    if model is None:
        return 1
    loaded = io.LoadColmapProblem(model, model_input_directory)
    if loaded is None:
        return 1
    dataset, state = loaded
    lam = -1.0
    for _ in range(max_iteration_count):
        cost, lam, _ = api.OptimizeJointly(dataset, state, 1, lam, 1e-4, 0, True, True, SchurMode.Dense, print_progress=False)
        io.SaveBAState(model_output_directory, state)
        with open(os.path.join(model_output_directory, "cost.txt"), "w") as f:
            f.write(f"{cost:.14g}\n")
    return 0
