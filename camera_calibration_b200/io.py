"""Readers / writers of the reference's on-disk formats for the bundle-adjustment path
(SURVEY.md 8f-1): ``dataset.bin`` and the state directory (``intrinsicsN.yaml``,
``rig_tr_global.yaml``, ``camera_tr_rig.yaml``, ``points.yaml``).

Formats follow applications/camera_calibration/src/camera_calibration/io/calibration_io.cc
(``APP/io`` below):

* ``dataset.bin`` (SaveDataset :51-135, LoadDataset :137-246): magic ``calib_data``, u32 version 0,
  u32 camera count, per camera u32 width, height; u32 imageset count, per imageset u32 filename
  length + bytes, per camera u32 n + n x (f32 x, f32 y, i32 id); known geometries: u32 count,
  each f32 cell length, u32 n, n x (i32 id, i32 x, i32 y). Integers are BIG-endian (htonl,
  io_util.h:56-64), floats are written raw in host order (io_util.h:66-69).
* camera model YAML (SaveCameraModel :526-647, LoadCameraModel :649-783): 14 significant digits;
  grids flat row-major x, y, z; directions are re-normalised on load.
* poses YAML (SavePoses :785-839, LoadPoses :841-888): ``pose_count`` + list of
  ``index, tx, ty, tz, qx, qy, qz, qw``; only used images are listed.
* ``points.yaml`` (:890-985): flat ``points`` + ``feature_id_to_point_index`` list.
"""
from __future__ import annotations

import os
import struct
from typing import Dict, List, Optional, Tuple

import numpy as np
import yaml

from .api import (BAState, CameraModel, CentralGenericModel, CentralOpenCVModel, Dataset, Imageset,
                  NoncentralGenericModel)

try:
    _Loader = yaml.CSafeLoader
except AttributeError:  # pragma: no cover
    _Loader = yaml.SafeLoader

_MAGIC = b"calib_data"


class KnownGeometry:
    """APP/dataset.h:45-55."""

    def __init__(self):
        self.cell_length_in_meters = 0.0
        self.feature_id_to_position: Dict[int, Tuple[int, int]] = {}


def _g(v: float) -> str:
    """std::ostream << double with setprecision(14)."""
    return f"{float(v):.14g}"


# ---------------------------------------------------------------------------------------
# dataset.bin
# ---------------------------------------------------------------------------------------
def SaveDataset(path: str, dataset: Dataset, known_geometries: Optional[List[KnownGeometry]] = None) -> bool:
    os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
    geoms = known_geometries if known_geometries is not None else getattr(dataset, "known_geometries", [])
    with open(path, "wb") as f:
        f.write(_MAGIC)
        f.write(struct.pack(">I", 0))
        f.write(struct.pack(">I", dataset.num_cameras()))
        for c in range(dataset.num_cameras()):
            w, h = (int(v) for v in dataset.GetImageSize(c))
            f.write(struct.pack(">II", w, h))
        f.write(struct.pack(">I", dataset.ImagesetCount()))
        for i in range(dataset.ImagesetCount()):
            s = dataset.GetImageset(i)
            name = s.GetFilename().encode()
            f.write(struct.pack(">I", len(name)))
            f.write(name)
            for c in range(dataset.num_cameras()):
                ft = s.FeaturesOfCamera(c)
                n = len(ft["id"])
                f.write(struct.pack(">I", n))
                rec = np.zeros(n, dtype=np.dtype([("x", "<f4"), ("y", "<f4"), ("id", ">i4")]))
                rec["x"] = ft["xy"][:, 0]
                rec["y"] = ft["xy"][:, 1]
                rec["id"] = ft["id"]
                f.write(rec.tobytes())
        f.write(struct.pack(">I", len(geoms)))
        for g in geoms:
            f.write(struct.pack("<f", g.cell_length_in_meters))
            f.write(struct.pack(">I", len(g.feature_id_to_position)))
            for fid, (x, y) in g.feature_id_to_position.items():
                f.write(struct.pack(">iii", fid, x, y))
    return True


def LoadDataset(path: str) -> Optional[Dataset]:
    """Returns the Dataset (with ``known_geometries`` attached) or None on a malformed file."""
    try:
        data = open(path, "rb").read()
    except OSError:
        return None
    if data[:10] != _MAGIC:
        return None
    pos = 10

    def u32():
        nonlocal pos
        (v,) = struct.unpack_from(">I", data, pos)
        pos += 4
        return v

    try:
        if u32() != 0:
            return None
        ncam = u32()
        ds = Dataset(ncam)
        for c in range(ncam):
            w, h = u32(), u32()
            ds.SetImageSize(c, (w, h))
        nset = u32()
        rec_t = np.dtype([("x", "<f4"), ("y", "<f4"), ("id", ">i4")])
        for _ in range(nset):
            ln = u32()
            name = data[pos:pos + ln].decode()
            pos += ln
            s = ds.NewImageset()
            s.SetFilename(name)
            for c in range(ncam):
                n = u32()
                rec = np.frombuffer(data, dtype=rec_t, count=n, offset=pos)
                pos += n * rec_t.itemsize
                s.SetFeaturesOfCamera(c, np.stack([rec["x"], rec["y"]], -1), rec["id"].astype(np.int32))
        ng = u32()
        geoms = []
        for _ in range(ng):
            g = KnownGeometry()
            (g.cell_length_in_meters,) = struct.unpack_from("<f", data, pos)
            pos += 4
            m = u32()
            for _ in range(m):
                fid, x, y = struct.unpack_from(">iii", data, pos)
                pos += 12
                g.feature_id_to_position[fid] = (x, y)
            geoms.append(g)
        ds.known_geometries = geoms
        return ds
    except (struct.error, ValueError, UnicodeDecodeError):
        # truncated file / bad filename bytes: the reference's reader returns false
        return None


# ---------------------------------------------------------------------------------------
# camera models
# ---------------------------------------------------------------------------------------
def _grid_text(grid: np.ndarray) -> str:
    return "[" + ", ".join(_g(v) for v in np.asarray(grid, dtype=np.float64).reshape(-1)) + "]\n"


def SaveCameraModel(model: CameraModel, path: str) -> bool:
    os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
    with open(path, "w") as f:
        if isinstance(model, CentralGenericModel):
            f.write("type : CentralGenericModel\n")
            f.write(f"width : {model.width()}\nheight : {model.height()}\n")
            f.write(f"calibration_min_x : {model.calibration_min_x()}\ncalibration_min_y : {model.calibration_min_y()}\n")
            f.write(f"calibration_max_x : {model.calibration_max_x()}\ncalibration_max_y : {model.calibration_max_y()}\n")
            gw, gh = model.GetGridResolution()
            f.write(f"grid_width : {gw}\ngrid_height : {gh}\n")
            f.write("# The grid is stored in row-major order, top to bottom. Each row is stored left to right. "
                    "Each grid point is stored as x, y, z.\n")
            f.write("grid : " + _grid_text(model.grid()))
        elif isinstance(model, NoncentralGenericModel):
            f.write("type : NoncentralGenericModel\n")
            f.write(f"width : {model.width()}\nheight : {model.height()}\n")
            f.write(f"calibration_min_x : {model.calibration_min_x()}\ncalibration_min_y : {model.calibration_min_y()}\n")
            f.write(f"calibration_max_x : {model.calibration_max_x()}\ncalibration_max_y : {model.calibration_max_y()}\n")
            gw, gh = model.GetGridResolution()
            f.write(f"grid_width : {gw}\ngrid_height : {gh}\n")
            f.write("# The grids are stored in row-major order, top to bottom. Each row is stored left to right. "
                    "Each grid point is stored as x, y, z.\n")
            f.write("point_grid : " + _grid_text(model.point_grid()))
            f.write("direction_grid : " + _grid_text(model.direction_grid()))
        elif isinstance(model, CentralOpenCVModel):
            f.write("type : CentralOpenCVModel\n")
            f.write(f"width : {model.width()}\nheight : {model.height()}\n")
            f.write("parameters : [" + ", ".join(_g(v) for v in model.parameters()) + "]\n")
        else:
            return False
    return True


def LoadCameraModel(path: str) -> Optional[CameraModel]:
    try:
        node = yaml.load(open(path), Loader=_Loader)
    except (OSError, yaml.YAMLError):
        return None
    if not node:
        return None
    width, height = int(node["width"]), int(node["height"])
    if width < 1 or height < 1:
        return None
    t = node["type"]

    def load_grid(key, gw, gh, normalized):
        a = np.array(node[key], dtype=np.float64)
        if a.size != 3 * gw * gh:
            raise ValueError(f"expected {3 * gw * gh} entries in '{key}', got {a.size}")
        a = a.reshape(gh, gw, 3)
        if normalized:  # re-normalise (calibration_io.cc:672-675)
            a = a / np.linalg.norm(a, axis=-1, keepdims=True)
        return a

    try:
        if t == "CentralGenericModel":
            gw, gh = int(node["grid_width"]), int(node["grid_height"])
            m = CentralGenericModel(gw, gh, int(node["calibration_min_x"]), int(node["calibration_min_y"]),
                                    int(node["calibration_max_x"]), int(node["calibration_max_y"]), width, height)
            m.m_grid = load_grid("grid", gw, gh, True)
            return m
        if t == "NoncentralGenericModel":
            gw, gh = int(node["grid_width"]), int(node["grid_height"])
            m = NoncentralGenericModel(gw, gh, int(node["calibration_min_x"]), int(node["calibration_min_y"]),
                                       int(node["calibration_max_x"]), int(node["calibration_max_y"]), width, height)
            m.SetPointGrid(load_grid("point_grid", gw, gh, False))
            m.SetDirectionGrid(load_grid("direction_grid", gw, gh, True))
            return m
        if t == "CentralOpenCVModel":
            p = np.array(node["parameters"], dtype=np.float64)
            if p.size != 12:
                return None
            return CentralOpenCVModel(width, height, p)
    except (KeyError, ValueError):
        return None
    return None  # model type not on the accelerated path


# ---------------------------------------------------------------------------------------
# poses, points, whole state
# ---------------------------------------------------------------------------------------
def SavePoses(image_used, poses: np.ndarray, path: str) -> bool:
    """poses rows are (qw qx qy qz tx ty tz)."""
    if len(image_used) != len(poses):
        raise ValueError("image_used and poses differ in size")  # CHECK_EQ, calibration_io.cc:791
    os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
    with open(path, "w") as f:
        f.write("# Each pose gives the B_tr_A transformation (i.e., A to B with right-multiplication), where the "
                "spaces A and B are defined by the filename. Quaternions are written as used by the Eigen library.\n")
        f.write(f"pose_count: {len(image_used)}\nposes:\n")
        for i, used in enumerate(image_used):
            if not used:
                continue
            qw, qx, qy, qz, tx, ty, tz = (float(v) for v in poses[i])
            f.write(f"  - index: {i}\n    tx: {_g(tx)}\n    ty: {_g(ty)}\n    tz: {_g(tz)}\n"
                    f"    qx: {_g(qx)}\n    qy: {_g(qy)}\n    qz: {_g(qz)}\n    qw: {_g(qw)}\n")
    # the reference also writes <path>.obj with the camera centres (visualisation only)
    with open(path + ".obj", "w") as f:
        for i, used in enumerate(image_used):
            if used:
                q = np.asarray(poses[i][:4], dtype=np.float64)
                t = np.asarray(poses[i][4:], dtype=np.float64)
                w, x, y, z = q
                R = np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y)],
                              [2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x)],
                              [2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)]])
                c = -R.T @ t
                f.write(f"v {_g(c[0])} {_g(c[1])} {_g(c[2])} 1 0 0\n")
    return True


def LoadPoses(path: str):
    """Returns (image_used list, poses [n, 7]) or None (malformed file: the reference returns false)."""
    try:
        node = yaml.load(open(path), Loader=_Loader)
        n = int(node["pose_count"])
        if n < 0:
            return None
        used = [False] * n
        poses = np.tile(np.array([1.0, 0, 0, 0, 0, 0, 0]), (n, 1))
        items = node.get("poses") or []
        if not isinstance(items, list):
            return None
        for it in items:
            i = int(it["index"])
            if i < 0 or i >= n:
                return None
            used[i] = True
            q = np.array([it["qw"], it["qx"], it["qy"], it["qz"]], dtype=np.float64)
            q = q / np.linalg.norm(q)  # SE3::setQuaternion normalises
            poses[i] = np.concatenate([q, [float(it["tx"]), float(it["ty"]), float(it["tz"])]])
        return used, poses
    except (OSError, yaml.YAMLError, TypeError, KeyError, ValueError, AttributeError):
        return None


def SavePointsAndIndexMapping(state: BAState, path: str) -> bool:
    os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
    pts = np.asarray(state.points, dtype=np.float64).reshape(-1)
    with open(path, "w") as f:
        f.write("# Each point is stored as x, y, z.\n")
        f.write("points : [" + ", ".join(_g(v) for v in pts) + "]\n")
        f.write("feature_id_to_point_index:\n")
        for fid, idx in state.feature_id_to_points_index.items():
            f.write(f"  - feature_id: {fid}\n    point_index: {idx}\n")
    with open(path + ".obj", "w") as f:
        for p in np.asarray(state.points, dtype=np.float64).reshape(-1, 3):
            f.write(f"v {_g(p[0])} {_g(p[1])} {_g(p[2])} 0 0 1\n")
    return True


def LoadPointsAndIndexMapping(path: str):
    try:
        node = yaml.load(open(path), Loader=_Loader)
        pts = np.array(node["points"], dtype=np.float64)
        if pts.size % 3 != 0:
            return None
        mapping = {int(it["feature_id"]): int(it["point_index"]) for it in (node.get("feature_id_to_point_index") or [])}
        if any(v < 0 or 3 * v >= max(pts.size, 1) for v in mapping.values()):
            return None
        return pts.reshape(-1, 3), mapping
    except (OSError, yaml.YAMLError, TypeError, KeyError, ValueError, AttributeError):
        return None


def SaveBAState(base_path: str, state: BAState) -> bool:
    """calibration_io.cc:432-464."""
    os.makedirs(base_path, exist_ok=True)
    if not SavePoses(state.image_used, state.rig_tr_global, os.path.join(base_path, "rig_tr_global.yaml")):
        return False
    if not SavePoses([True] * len(state.camera_tr_rig), state.camera_tr_rig, os.path.join(base_path, "camera_tr_rig.yaml")):
        return False
    for c, m in enumerate(state.intrinsics):
        if not SaveCameraModel(m, os.path.join(base_path, f"intrinsics{c}.yaml")):
            return False
    return SavePointsAndIndexMapping(state, os.path.join(base_path, "points.yaml"))


def LoadBAState(base_path: str, dataset: Optional[Dataset] = None) -> Optional[BAState]:
    """calibration_io.cc:466-523."""
    st = BAState()
    r = LoadPoses(os.path.join(base_path, "rig_tr_global.yaml"))
    if r is None:
        return None
    st.image_used, st.rig_tr_global = r
    r = LoadPoses(os.path.join(base_path, "camera_tr_rig.yaml"))
    if r is None:
        return None
    st.camera_tr_rig = r[1]
    c = 0
    while True:
        p = os.path.join(base_path, f"intrinsics{c}.yaml")
        if not os.path.exists(p):
            if c == 0:
                return None
            break
        m = LoadCameraModel(p)
        if m is None:
            return None
        st.intrinsics.append(m)
        c += 1
    r = LoadPointsAndIndexMapping(os.path.join(base_path, "points.yaml"))
    if r is None:
        return None
    st.points, st.feature_id_to_points_index = r
    if dataset is not None:
        st.ComputeFeatureIdToPointsIndex(dataset)
    return st


# ---------------------------------------------------------------------------------------
# COLMAP text model (the --bundle_adjustment entry point)
# ---------------------------------------------------------------------------------------
def ReadColmapImages(images_txt_path: str, read_observations: bool = True):
    """libvis/src/libvis/external_io/colmap_model.cc:96-142: two lines per image,
    ``IMAGE_ID QW QX QY QZ TX TY TZ CAMERA_ID NAME`` then ``X Y POINT3D_ID ...``; values are
    parsed into FLOAT (SE3f / Vector2f) like the reference. Returns {image_id: dict} or None."""
    try:
        lines = open(images_txt_path).read().split("\n")
    except OSError:
        return None
    images = {}
    i = 0
    while i < len(lines):
        line = lines[i]
        i += 1
        if len(line) == 0 or line[0] == "#":
            continue
        f = line.split()
        q = np.array([f[1], f[2], f[3], f[4]], dtype=np.float32)
        t = np.array([f[5], f[6], f[7]], dtype=np.float32)
        obs_line = lines[i] if i < len(lines) else ""
        i += 1
        xy = np.zeros((0, 2), np.float32)
        ids = np.zeros(0, np.int64)
        if read_observations:
            v = obs_line.split()
            n = len(v) // 3
            arr = np.array(v[:3 * n], dtype=np.float64).reshape(n, 3)
            xy = arr[:, :2].astype(np.float32)
            ids = arr[:, 2].astype(np.int64)
        images[int(f[0])] = dict(image_id=int(f[0]), q=q, t=t, camera_id=int(f[8]), file_path=f[9] if len(f) > 9 else "",
                                 xy=xy, point3d_id=ids)
    return images


def ReadColmapPoints3D(points3d_txt_path: str):
    """colmap_model.cc:265-299: ``ID X Y Z R G B ERROR track...`` (positions float, tracks ignored)."""
    try:
        lines = open(points3d_txt_path).read().split("\n")
    except OSError:
        return None
    pts = {}
    for line in lines:
        if len(line) == 0 or line[0] == "#":
            continue
        f = line.split()
        pts[int(f[0])] = np.array(f[1:4], dtype=np.float32)
    return pts


def LoadColmapProblem(model: CameraModel, model_input_directory: str):
    """tools/bundle_adjustment.cc:110-184: COLMAP text model -> (Dataset, BAState); images and
    points are ordered by increasing id; observations without a 3D point (id -1) are dropped; the
    single rig pose is the identity."""
    images = ReadColmapImages(os.path.join(model_input_directory, "images.txt"), True)
    points = ReadColmapPoints3D(os.path.join(model_input_directory, "points3D.txt"))
    if images is None or points is None:
        return None
    ds = Dataset(1)
    ds.SetImageSize(0, (model.width(), model.height()))
    st = BAState()
    st.intrinsics = [model]
    st.camera_tr_rig = np.array([[1.0, 0, 0, 0, 0, 0, 0]])
    ordered = [images[k] for k in sorted(images)]
    st.image_used = [True] * len(ordered)
    st.rig_tr_global = np.zeros((len(ordered), 7))
    for i, im in enumerate(ordered):
        q = im["q"].astype(np.float64)
        st.rig_tr_global[i] = np.concatenate([q / np.linalg.norm(q), im["t"].astype(np.float64)])
        s = ds.NewImageset()
        s.SetFilename(im["file_path"])
        keep = im["point3d_id"] >= 0
        s.SetFeaturesOfCamera(0, im["xy"][keep], im["point3d_id"][keep].astype(np.int32))
    ids = sorted(points)
    st.points = np.array([points[k] for k in ids], dtype=np.float64).reshape(-1, 3)
    st.feature_id_to_points_index = {k: i for i, k in enumerate(ids)}
    st.ComputeFeatureIdToPointsIndex(ds)
    return ds, st
