"""Experiment (not product): is an FP32 Cholesky of the reduced system S a usable preconditioner
for an FP64 conjugate-gradient solve at config-2 size? Prints timings and iteration counts."""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from camera_calibration_b200 import api, cabi  # noqa: E402


def tms(fn, n=3):
    fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        r = fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n, r


def reduced(H, b, lam, nbd):
    Hd = torch.from_numpy(H).cuda()
    bd = torch.from_numpy(b).cuda()
    n = Hd.shape[0]
    Hd = torch.triu(Hd) + torch.triu(Hd, 1).T  # whichever triangle is filled: symmetrise from upper
    D = torch.stack([Hd[i:i + 3, i:i + 3] for i in range(0, nbd, 3)]) + lam * torch.eye(3, device="cuda", dtype=torch.float64)
    B = Hd[:nbd, nbd:]
    Cm = Hd[nbd:, nbd:]
    DiB = torch.linalg.solve(D, B.reshape(nbd // 3, 3, -1)).reshape(nbd, -1)
    S = Cm + lam * torch.eye(n - nbd, device="cuda", dtype=torch.float64) - B.T @ DiB
    Dib = torch.linalg.solve(D, bd[:nbd].reshape(-1, 3, 1)).reshape(-1)
    rhs = bd[nbd:] - B.T @ Dib
    return S, rhs


def pcg(S, rhs, apply_M, tol=1e-13, maxit=60):
    x = torch.zeros_like(rhs)
    r = rhs.clone()
    z = apply_M(r)
    p = z.clone()
    rz = torch.dot(r, z)
    r0 = rhs.norm()
    hist = []
    for it in range(maxit):
        Sp = S @ p
        alpha = rz / torch.dot(p, Sp)
        x += alpha * p
        r -= alpha * Sp
        rel = (r.norm() / r0).item()
        hist.append(rel)
        if rel < tol:
            break
        z = apply_M(r)
        rz_new = torch.dot(r, z)
        p = z + (rz_new / rz) * p
        rz = rz_new
    return x, hist


def analyse(tag, S, rhs):
    n = S.shape[0]
    print(f"--- {tag}: n={n}", flush=True)
    t64, L64 = tms(lambda: torch.linalg.cholesky_ex(S)[0])
    x_ref = torch.cholesky_solve(rhs[:, None], L64)[:, 0]
    print(f"fp64 cholesky {t64:.2f} ms; residual {((S @ x_ref - rhs).norm() / rhs.norm()).item():.2e}")
    d = S.diagonal().clone()
    print(f"diag range {d.min().item():.3e} .. {d.max().item():.3e}")
    sc = d.rsqrt()
    for name, scaled in (("plain", False), ("jacobi-scaled", True)):
        A = (S * sc[:, None] * sc[None, :]) if scaled else S
        A32 = A.float()
        t32, (L32, info) = tms(lambda: torch.linalg.cholesky_ex(A32))
        print(f"fp32 cholesky [{name}] {t32:.2f} ms info={int(info)}")
        if int(info) != 0:
            continue
        L32d = L32.double()

        def M32(r):
            rr = (r * sc) if scaled else r
            y = torch.cholesky_solve(rr.float()[:, None], L32)[:, 0].double()
            return (y * sc) if scaled else y

        def M64(r):  # fp32 factor applied in fp64 arithmetic
            rr = (r * sc) if scaled else r
            y = torch.cholesky_solve(rr[:, None], L32d)[:, 0]
            return (y * sc) if scaled else y
        for mname, M in (("solve32", M32), ("solve64", M64)):
            x, hist = pcg(S, rhs, M)
            err = ((x - x_ref).norm() / x_ref.norm()).item()
            print(f"  pcg [{mname}] iters={len(hist)} rel_res={hist[-1]:.2e} x_err_vs_fp64={err:.2e} hist={['%.1e' % v for v in hist[:12]]}")
        tsolve, _ = tms(lambda: M32(rhs), 5)
        tmv, _ = tms(lambda: S @ rhs, 5)
        print(f"  one fp32 two-sided solve {tsolve:.3f} ms, one fp64 symv (torch gemv) {tmv:.3f} ms")
    # eigen extremes by power / inverse iteration through the fp64 factor
    v = torch.randn(n, device="cuda", dtype=torch.float64)
    for _ in range(50):
        v = S @ v
        v /= v.norm()
    lmax = torch.dot(v, S @ v).item()
    w = torch.randn(n, device="cuda", dtype=torch.float64)
    for _ in range(50):
        w = torch.cholesky_solve(w[:, None], L64)[:, 0]
        w /= w.norm()
    lmin = torch.dot(w, S @ w).item()
    print(f"cond(S) ~ {lmax / lmin:.3e}  (lmax {lmax:.3e}, lmin {lmin:.3e})")
    As = S * sc[:, None] * sc[None, :]
    Ls = torch.linalg.cholesky(As)
    v = torch.randn(n, device="cuda", dtype=torch.float64)
    for _ in range(50):
        v = As @ v
        v /= v.norm()
    w = torch.randn(n, device="cuda", dtype=torch.float64)
    for _ in range(50):
        w = torch.cholesky_solve(w[:, None], Ls)[:, 0]
        w /= w.norm()
    print(f"cond(scaled S) ~ {torch.dot(v, As @ v).item() / torch.dot(w, As @ w).item():.3e}")


def main():
    class A:
        config = 2
        imagesets = int(os.environ.get("EXP_IMAGESETS", "0"))
    sp = bench.make_workload(A)
    opt = cabi.default_options(max_iteration_count=1)
    adj = api.BundleAdjuster(sp.problem, 0)
    adj.set_state(sp.init_state.copy())
    nbd = 3 * sp.problem.n_points
    lam = -1.0
    for stage in range(3):
        H, b, cost = adj.build_system(opt)
        dof = H.shape[0]
        if lam <= 0:
            lam = 1e-5 * float(np.trace(H)) / dof
        print(f"stage {stage}: cost {cost:.6e} lambda {lam:.4e} dof {dof}")
        S, rhs = reduced(H, b, lam, nbd)
        del H
        analyse(f"stage{stage}", S, rhs)
        del S
        torch.cuda.empty_cache()
        n_it = 4
        for _ in range(n_it):
            opt.init_lambda = lam
            rep = adj.optimize(opt)
            lam = rep.final_lambda
        print(f"after {n_it} more iterations: cost {rep.final_cost:.6e} lambda {lam:.4e}")


if __name__ == "__main__":
    main()
