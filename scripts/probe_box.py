"""Probe of the GPU box: host cores / memory and measured FP64 dense peaks (cuBLAS dgemm / dsyrk via torch).
Writes gpurun_out/box_probe.json. Diagnostic only (feeds profiles/ and the roofline denominators of the
dense phases)."""
import json, os, subprocess, time
import torch
out = {"nproc": os.cpu_count()}
try:
    out["lscpu"] = subprocess.run("lscpu | egrep 'Model name|Socket|Core|Thread|^CPU\\(s\\)'", shell=True, capture_output=True, text=True).stdout
    out["mem"] = subprocess.run("free -g | head -2", shell=True, capture_output=True, text=True).stdout
except Exception as e:
    out["err"] = str(e)
dev = torch.device("cuda:0")
out["gpu"] = torch.cuda.get_device_name(0)
def timeit(f, n=5):
    f(); torch.cuda.synchronize()
    best = 1e9
    for _ in range(n):
        a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
        a.record(); f(); b.record(); torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b))
    return best
res = {}
for n in (4096, 8192, 13080):
    A = torch.randn(n, n, dtype=torch.float64, device=dev)
    B = torch.randn(n, n, dtype=torch.float64, device=dev)
    ms = timeit(lambda: torch.matmul(A, B))
    res[f"dgemm_{n}"] = {"ms": ms, "tflops": 2 * n**3 / ms / 1e9}
# thin-k shapes like the compact contraction panels and the Cholesky trailing update
for (m, k) in ((4608, 288), (13080, 512), (8192, 512), (8192, 256)):
    A = torch.randn(m, k, dtype=torch.float64, device=dev)
    ms = timeit(lambda: torch.matmul(A, A.t()))
    res[f"dgemm_nt_{m}x{m}x{k}"] = {"ms": ms, "tflops": 2 * m * m * k / ms / 1e9}
n = 13080
S = torch.randn(n, n, dtype=torch.float64, device=dev); S = S @ S.t() + n * torch.eye(n, dtype=torch.float64, device=dev)
ms = timeit(lambda: torch.linalg.cholesky(S), n=3)
res["potrf_13080"] = {"ms": ms, "tflops": n**3 / 3 / ms / 1e9}
out["fp64"] = res
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/box_probe.json", "w"), indent=1)
print(json.dumps(out, indent=1))
