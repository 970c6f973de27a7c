"""Time of the reference's loop BODY (fingering_periodic.py:455-479) driven statement by statement through the twin's
per-operation methods at the shipped size (400 x 400, 90 circles), next to run_loop (whole loop on the device)."""
import copy, os, sys, time
import numpy as np
sys.path.insert(0, os.getcwd())
from fingering_dynamics_b200.lattice_boltzmann import fingering_periodic as FP
from fingering_dynamics_b200.lattice_boltzmann.create_block import Createblock
from fingering_dynamics_b200.lattice_boltzmann.bounce_back import Bounce_back

H, W = FP.H, FP.W
circles = [((25 + 40 * k, 20 + 40 * m), 10) for k in range(9) for m in range(10)]
bpa, side_list, concave_list, convex_list = Createblock(H, W).setCirleblock(circles)
mask = np.logical_not(bpa == 1)
bb = Bounce_back(H, W)
cm = FP.Compute(mask)


def iteration():
    for j in range(9):
        cm.F[j] = cm.getLarge_F(j)
        cm.feq[j] = cm.getfeq(j)
        cm.geq[j] = cm.getgeq(j)
        cm.f[j][mask] = cm.getF(j)
        cm.g[j][mask] = cm.getG(j)
    f_behind = copy.deepcopy(cm.f)
    g_behind = copy.deepcopy(cm.g)
    FP.stream(cm.f, cm.g)
    bb.halfway_bounceback_circle(side_list, concave_list, convex_list, f_behind, g_behind, cm.f, cm.g)
    cm.zou_he_boundary_inlet()
    cm.zou_he_boundary_outlet()
    cm.rho = cm.getRho()
    cm.udpatePsi()
    cm.nabla_psix = cm.getNabla_psix()
    cm.nabla_psiy = cm.getNabla_psiy()
    cm.nabla_psi2 = cm.getNabla_psi2()
    cm.mu = cm.getMu()
    cm.ux = cm.getUx()
    cm.uy = cm.getUy()
    cm.p = cm.getP()
    cm.mix_tau = cm.getMix_tau()


iteration()
n = int(os.environ.get("VL_STEPS", 5))
t0 = time.perf_counter()
for _ in range(n):
    iteration()
dt = (time.perf_counter() - t0) / n
print("verbatim loop body through the per-op twins: %.1f ms per iteration at %dx%d (%.2f MLUPS)" % (dt * 1e3, W, H, H * W / dt / 1e6))
# what the user's own NumPy statements in that body cost (no device work at all)
t0 = time.perf_counter()
for _ in range(n):
    for j in range(9):
        cm.f[j][mask] = cm.f[j][mask]
        cm.g[j][mask] = cm.g[j][mask]
    fb = copy.deepcopy(cm.f); gb = copy.deepcopy(cm.g)
dh = (time.perf_counter() - t0) / n
print("  of which the body's own host statements (18 masked assignments + 2 deepcopies): %.1f ms" % (dh * 1e3))

if os.environ.get("VL_PROFILE"):
    import cProfile, pstats
    pr = cProfile.Profile()
    pr.enable()
    for _ in range(3):
        iteration()
    pr.disable()
    pstats.Stats(pr).sort_stats("cumulative").print_stats(22)
