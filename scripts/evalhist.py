import sys, ctypes as C, numpy as np
sys.path.insert(0,'/root/repo')
from camera_calibration_b200 import api, cabi, synthetic
sp = synthetic.make_problem(2, n_imagesets=int(sys.argv[1]) if len(sys.argv)>1 else 500)
opt = cabi.default_options(max_iteration_count=1)
adj = api.BundleAdjuster(sp.problem)
adj.set_state(sp.init_state)
lib = adj.lib
lib.b200ba_debug_eval_counts.restype=C.c_int
lib.b200ba_debug_eval_counts.argtypes=[C.c_void_p, C.POINTER(C.c_uint16)]
lam=-1.0
for it in range(5):
    e = adj.evaluate(opt, compute_jacobians=False)
    e = adj.evaluate(opt, compute_jacobians=True)
    cnt = np.zeros(sp.n_obs, dtype=np.uint16)
    lib.b200ba_debug_eval_counts(adj._h, cnt.ctypes.data_as(C.POINTER(C.c_uint16)))
    inv = e['costs']<0
    h = np.bincount(np.minimum(cnt,40))
    print(f"iter {it}: invalid {inv.sum()} evals mean {cnt.mean():.2f} max {cnt.max()} p99.9 {np.percentile(cnt,99.9)} >10: {(cnt>10).sum()} >100: {(cnt>100).sum()}  hist {h[:12]}")
    if (cnt>100).sum():
        idx=np.nonzero(cnt>100)[0][:5]
        print('   stragglers xy', sp.problem.obs_xy[idx], 'valid', ~inv[idx], cnt[idx])
    opt.init_lambda=lam
    rep = adj.optimize(opt); lam=rep.final_lambda
