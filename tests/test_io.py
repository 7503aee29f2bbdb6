"""Round trips of the reference's on-disk formats (recipe of
applications/camera_calibration/src/camera_calibration/test/io_test.cc:87-393: 1e-6 on floats,
exact on integers) plus byte-level checks of what calibration_io.cc writes."""
import os
import struct

import numpy as np

from camera_calibration_b200 import api, io, synthetic


def _dataset():
    ds = api.Dataset(2)
    ds.SetImageSize(0, (640, 480))
    ds.SetImageSize(1, (1280, 720))
    rng = np.random.default_rng(0)
    for i in range(3):
        s = ds.NewImageset()
        s.SetFilename(f"image{i:04d}.png")
        for c in range(2):
            n = int(rng.integers(0, 6))
            s.SetFeaturesOfCamera(c, rng.uniform(0, 400, (n, 2)).astype(np.float32), rng.integers(0, 1000, n))
    g = io.KnownGeometry()
    g.cell_length_in_meters = 0.0119
    g.feature_id_to_position = {5: (1, -2), 9: (3, 4)}
    ds.known_geometries = [g]
    return ds


def test_dataset_bin_roundtrip_and_layout(tmp_path):
    ds = _dataset()
    p = str(tmp_path / "sub" / "dataset.bin")
    assert io.SaveDataset(p, ds)
    raw = open(p, "rb").read()
    assert raw[:10] == b"calib_data"
    assert struct.unpack(">II", raw[10:18]) == (0, 2)                 # version, camera count: big-endian
    assert struct.unpack(">IIII", raw[18:34]) == (640, 480, 1280, 720)
    assert struct.unpack(">I", raw[34:38]) == (3,)
    assert struct.unpack(">I", raw[38:42]) == (len("image0000.png"),)
    ds2 = io.LoadDataset(p)
    assert ds2 is not None and ds2.num_cameras() == 2 and ds2.ImagesetCount() == 3
    for c in range(2):
        assert tuple(ds2.GetImageSize(c)) == tuple(ds.GetImageSize(c))
    for i in range(3):
        assert ds2.GetImageset(i).GetFilename() == ds.GetImageset(i).GetFilename()
        for c in range(2):
            a, b = ds.GetImageset(i).FeaturesOfCamera(c), ds2.GetImageset(i).FeaturesOfCamera(c)
            assert np.array_equal(a["id"], b["id"])
            assert np.array_equal(a["xy"], b["xy"])  # floats are written raw
    g = ds2.known_geometries[0]
    assert abs(g.cell_length_in_meters - 0.0119) < 1e-6
    assert g.feature_id_to_position == {5: (1, -2), 9: (3, 4)}
    # malformed files are rejected
    open(p, "wb").write(b"calib_dat_" + raw[10:])
    assert io.LoadDataset(p) is None
    open(p, "wb").write(raw[:10] + struct.pack(">I", 7) + raw[14:])
    assert io.LoadDataset(p) is None


def test_camera_model_yaml_roundtrip(tmp_path):
    sp = synthetic.make_problem(3, n_imagesets=2, lattice=(5, 4), image_size=(300, 240))
    _, st = api.dataset_from_flat(sp.problem, sp.init_state)
    m = st.intrinsics[0]
    p = str(tmp_path / "intrinsics0.yaml")
    assert io.SaveCameraModel(m, p)
    text = open(p).read()
    assert text.startswith("type : NoncentralGenericModel\nwidth : 300\nheight : 240\ncalibration_min_x : 0\n")
    assert "point_grid : [" in text and "direction_grid : [" in text
    m2 = io.LoadCameraModel(p)
    assert isinstance(m2, api.NoncentralGenericModel)
    assert m2.GetGridResolution() == m.GetGridResolution()
    assert np.abs(m2.point_grid() - m.point_grid()).max() < 1e-6
    assert np.abs(m2.direction_grid() - m.direction_grid()).max() < 1e-6
    assert np.abs(np.linalg.norm(m2.direction_grid(), axis=-1) - 1).max() < 1e-15  # re-normalised on load

    sp = synthetic.make_problem(2, n_imagesets=2, lattice=(5, 4), image_size=(300, 240))
    _, st = api.dataset_from_flat(sp.problem, sp.init_state)
    assert io.SaveCameraModel(st.intrinsics[0], p)
    m3 = io.LoadCameraModel(p)
    assert isinstance(m3, api.CentralGenericModel)
    assert np.abs(m3.grid() - st.intrinsics[0].grid()).max() < 1e-6
    assert (m3.calibration_min_x(), m3.calibration_max_x(), m3.calibration_max_y()) == (0, 299, 239)

    cv = api.CentralOpenCVModel(640, 480, [480, 481, 320, 240, 0.05, -0.01, 1e-3, 0, 0, 0, 1e-4, -2e-4])
    assert io.SaveCameraModel(cv, p)
    assert open(p).read().startswith("type : CentralOpenCVModel\nwidth : 640\nheight : 480\nparameters : [480, 481, 320")
    cv2 = io.LoadCameraModel(p)
    assert isinstance(cv2, api.CentralOpenCVModel) and np.abs(cv2.parameters() - cv.parameters()).max() < 1e-12


def test_ba_state_directory_roundtrip(tmp_path):
    sp = synthetic.make_problem(4, n_imagesets=6, lattice=(5, 4), image_size=(300, 240))
    ds, st = api.dataset_from_flat(sp.problem, sp.init_state)
    st.image_used[2] = False
    st.feature_id_to_points_index = {100 + i: i for i in range(len(st.points))}
    d = str(tmp_path / "state")
    assert io.SaveBAState(d, st)
    for f in ("rig_tr_global.yaml", "camera_tr_rig.yaml", "intrinsics0.yaml", "intrinsics1.yaml", "points.yaml",
              "rig_tr_global.yaml.obj", "points.yaml.obj"):
        assert os.path.exists(os.path.join(d, f)), f
    txt = open(os.path.join(d, "rig_tr_global.yaml")).read()
    assert "pose_count: 6" in txt and "  - index: 2\n" not in txt and "  - index: 3\n" in txt
    st2 = io.LoadBAState(d)
    assert st2 is not None
    assert st2.image_used == st.image_used
    used = np.array(st.image_used)
    assert np.abs(st2.rig_tr_global[used] - st.rig_tr_global[used]).max() < 1e-6
    assert np.abs(st2.camera_tr_rig - st.camera_tr_rig).max() < 1e-6
    assert np.abs(st2.points - st.points).max() < 1e-6
    assert st2.feature_id_to_points_index == st.feature_id_to_points_index
    assert len(st2.intrinsics) == 2
    for a, b in zip(st.intrinsics, st2.intrinsics):
        assert np.abs(a.grid() - b.grid()).max() < 1e-6
    assert io.LoadBAState(str(tmp_path / "missing")) is None
