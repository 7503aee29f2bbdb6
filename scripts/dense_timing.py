"""Factor / solve time of the in-tree dense kernels at the reduced-system size of BASELINE config 2."""
import sys, time, numpy as np
sys.path.insert(0, '/root/repo')
from camera_calibration_b200 import api
n = int(sys.argv[1]) if len(sys.argv) > 1 else 13080
rng = np.random.default_rng(0)
R = rng.uniform(-1, 1, (n, n))
A = R + R.T + 2.5 * n * np.eye(n)   # diagonally dominant: SPD
b = rng.standard_normal(n)
for nb in (256, 512):
    for rep in range(2):
        x, fm, sm = api.dense_cholesky_solve(A, b, nb)
    r = np.abs(A @ x - b).max()
    print(f"n {n} nb {nb}: factor {fm:.2f} ms ({n**3/3/fm/1e9:.1f} TF/s)  solve {sm:.2f} ms  residual {r:.2e}")
