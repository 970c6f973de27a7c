"""Step time of the reference's own (small) configurations: 400x400 (config 1), 380x380 (config 3), 200x250 (config 2)."""
import os, sys, time, numpy as np
sys.path.insert(0, os.getcwd())
from fingering_dynamics_b200 import Engine, synthetic as syn
for dtype in ("f64", "f32"):
    for (H, W, periodic) in ((400, 400, False), (380, 380, False), (200, 250, True), (1024, 1024, False)):
        c = syn.fp_constants(H)
        solid, refl = syn.porous_geometry(H, W)
        st = syn.fp_initial_state(solid, c)
        kw = dict(tau=c["tau"], gamma=c["gamma"], a=c["a"], kappa=c["kappa"], Eta_n=c["Eta_n"], M=c["M"], psi_wall=c["psi_wall"], dtype=dtype)
        if periodic:
            kw.update(zou_he="none", x_periodic=True, psi_y_wall=True)
        else:
            kw.update(zou_he="fp", inlet_ux=c["inlet_ux"], outlet_ux=c["outlet_ux"])
        e = Engine(H, W, **kw)
        e.set_geometry(solid, refl)
        e.set_state(**st)
        e.step(50); e.sync()
        t0 = time.perf_counter(); e.step(2000); e.sync(); dt = time.perf_counter() - t0
        print("%s %4dx%-4d %7.2f us/step  %8.0f MLUPS" % (dtype, H, W, dt / 2000 * 1e6, H * W * 2000 / dt / 1e6))
        e.close()
