"""Where the end-to-end job of bench.py spends its time (8192x2048 fp64, K=20): each phase with a synchronise after it,
then the whole job as bench.py runs it (no synchronisation inside)."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.getcwd())
import bench
from fingering_dynamics_b200 import pinned_empty

H, W, K = 2048, 8192, int(os.environ.get("K", 20))
job = bench.Job(H, W, 1, 0, 0, "f64", "auto", "peer")
out = {k: pinned_empty((H, W)) for k in ("psi", "rho", "ux", "uy")}


def t(fn):
    job.eng.sync()
    t0 = time.perf_counter()
    fn()
    job.eng.sync()
    return (time.perf_counter() - t0) * 1e3


for rep in range(3):
    a = t(lambda: job.eng.set_geometry(job.solid, job.refl, col0=job.lo))
    b = t(job.init)
    c = t(lambda: job.runner.step(K))
    d = t(lambda: job.runner.get_state(("psi", "rho", "ux", "uy"), out=out))
    d1 = t(lambda: job.runner.get_state(("psi",), out={"psi": out["psi"]}))
    t0 = time.perf_counter()
    job.eng.set_geometry(job.solid, job.refl, col0=job.lo)
    job.init()
    job.runner.step(K)
    job.runner.get_state(("psi", "rho", "ux", "uy"), out=out)
    whole = (time.perf_counter() - t0) * 1e3
    print("rep %d: set_geometry %.2f  init %.2f  %d steps %.2f  get_state(4 fields) %.2f  (psi only %.2f)  sum %.2f | whole job %.2f ms"
          % (rep, a, b, K, c, d, d1, a + b + c + d, whole))
import torch
x = torch.empty(H * W, dtype=torch.float64, device="cuda")
y = torch.empty(H * W, dtype=torch.float64).pin_memory()
torch.cuda.synchronize()
for _ in range(3):
    t0 = time.perf_counter(); y.copy_(x, non_blocking=True); torch.cuda.synchronize(); dt = time.perf_counter() - t0
print("plain pinned D2H of one 134 MB plane: %.2f ms = %.1f GB/s" % (dt * 1e3, x.numel() * 8 / dt / 1e9))
