"""Extracts the golden fixtures this repo's tests pin the oracle to.

Run in the build container (needs /root/reference); the outputs are committed so that the
tests never read /root/reference at run time.

  real_central_17x13.json  the only real calibrated camera in the reference: the 17x13
                           central-generic grid of generic_models/src/main.cc:87-97
                           (640x480, calibrated area 15..624 x 16..464).
"""
import json
import os
import re

REF = "/root/reference/applications/camera_calibration/generic_models/src/main.cc"
HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    src = open(REF).read()
    m = re.search(r'R"yaml\((.*?)\)yaml"', src, re.S)
    text = m.group(1)
    out = {}
    for key in ("width", "height", "calibration_min_x", "calibration_min_y", "calibration_max_x",
                "calibration_max_y", "grid_width", "grid_height"):
        out[key] = int(re.search(rf"^{key} : (\d+)", text, re.M).group(1))
    grid = re.search(r"^grid : \[(.*?)\]", text, re.M | re.S).group(1)
    out["grid"] = [float(v) for v in grid.split(",")]
    assert len(out["grid"]) == 3 * out["grid_width"] * out["grid_height"]
    out["source"] = "applications/camera_calibration/generic_models/src/main.cc:87-97"
    with open(os.path.join(HERE, "real_central_17x13.json"), "w") as f:
        json.dump(out, f)
    print("wrote real_central_17x13.json:", out["grid_width"], "x", out["grid_height"])


if __name__ == "__main__":
    main()
