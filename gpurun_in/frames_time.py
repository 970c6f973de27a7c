"""The reference's default run of config 1 (400 x 400, 4000 iterations, a psi frame every MAX_T // 150 = 26 steps) through the
drop-in driver's run_loop: frames queued (fdlbm_psi_frame_async) against frames read back one by one."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.getcwd())
from fingering_dynamics_b200 import geometry as geo
from fingering_dynamics_b200.lattice_boltzmann import fingering_periodic as FP, _compute
from fingering_dynamics_b200.lattice_boltzmann.create_block import Createblock

H, W = FP.H, FP.W
circles = [((25 + 40 * k, 20 + 40 * m), 10) for k in range(9) for m in range(10)]
bpa, side_list, concave_list, convex_list = Createblock(H, W).setCirleblock(circles)
mask = np.logical_not(bpa == 1)
reflect = geo.reflect_bits_circle(side_list, concave_list, convex_list)
res = {}
for label, budget in (("queued", 1 << 30), ("blocking", 0), ("queued", 1 << 30), ("blocking", 0)):
    _compute._ASYNC_FRAME_BYTES = budget
    cm = FP.Compute(mask)
    t0 = time.perf_counter()
    frames = _compute.run_loop(cm, reflect, FP.MAX_T, frames_every=FP.MAX_T // 150)
    dt = time.perf_counter() - t0
    res.setdefault(label, []).append((dt, frames))
    print("%-8s %d iterations, %d frames: %.3f s (%.1f us per iteration)" % (label, FP.MAX_T, len(frames), dt, dt / FP.MAX_T * 1e6))
a, b = res["queued"][-1][1], res["blocking"][-1][1]
print("frames identical:", len(a) == len(b) and all(np.array_equal(x, y) for x, y in zip(a, b)))

# a large grid: 8192 x 2048, a frame (134 MB) every 26 steps
from fingering_dynamics_b200 import Engine, synthetic as syn
H2, W2 = 2048, 8192
c = syn.fp_constants(H2)
solid, refl = syn.porous_geometry(H2, W2)
for label, budget in (("queued", 1 << 30), ("blocking", 0), ("queued", 1 << 30), ("blocking", 0)):
    _compute._ASYNC_FRAME_BYTES = budget
    e = Engine(H2, W2, tau=c["tau"], gamma=c["gamma"], a=c["a"], kappa=c["kappa"], Eta_n=c["Eta_n"], M=c["M"],
               psi_wall=c["psi_wall"], zou_he="fp", inlet_ux=c["inlet_ux"], outlet_ux=c["outlet_ux"])
    e.set_geometry(solid, refl)
    e.init_state(variant="fp", rho0=c["rho0"])
    e.step(5); e.sync()
    t0 = time.perf_counter()
    fr = _compute.step_with_frames(e, H2, W2, 520, 26)
    e.sync()
    dt = time.perf_counter() - t0
    print("8192x2048 %-8s 520 iterations, %d frames: %.3f s (%.3f ms per iteration; stepping alone 0.75)" % (label, len(fr), dt, dt / 520 * 1e3))
    e.close()
