"""Config 4 (2-camera rig) at full size: in-tree dense kernels vs the cuBLAS / cuSOLVER path, 3 LM iterations."""
import os, sys, numpy as np
sys.path.insert(0, '/root/repo')
from camera_calibration_b200 import api, cabi, synthetic
sp = synthetic.make_problem(4)
out = {}
for mode in ("own", "lib"):
    os.environ["B200BA_DENSE"] = mode
    with api.BundleAdjuster(sp.problem) as adj:
        st = sp.init_state.copy()
        rep = adj.optimize_host(st, cabi.default_options(max_iteration_count=3))
        t = adj.timings()
        out[mode] = (rep.trace(), rep.n_invalid, rep.rmse, st)
        print(mode, "costs", rep.trace()[0], "attempts", rep.trace()[2], "n_invalid", rep.n_invalid, "rmse %.9f" % rep.rmse,
              "factor %.1f schur %.1f solve %.1f" % (t.factor_ms - t.solve_ms, t.schur_ms, t.solve_ms))
a, b = out["own"], out["lib"]
print("max rel cost diff", max(abs(x - y) / y for x, y in zip(a[0][0], b[0][0])), "state diff",
      np.abs(a[3].points - b[3].points).max(), np.abs(a[3].intrinsics[0] - b[3].intrinsics[0]).max())
