"""Long bitwise comparison of the fused step against the two-pass kernel at 8192x2048 (fp64) and of two fused runs against
each other (fp32): a timing-dependent hazard in the stage rings would show up as a difference somewhere along the way."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.getcwd())
from fingering_dynamics_b200 import Engine, synthetic as syn

H, W = 2048, 8192
N = int(os.environ.get("N", 1000))
c = syn.fp_constants(H)
solid, refl = syn.porous_geometry(H, W)


def run(kernel, dtype, n):
    e = Engine(H, W, tau=c["tau"], gamma=c["gamma"], a=c["a"], kappa=c["kappa"], Eta_n=c["Eta_n"], M=c["M"],
               psi_wall=c["psi_wall"], zou_he="fp", inlet_ux=c["inlet_ux"], outlet_ux=c["outlet_ux"], kernel=kernel, dtype=dtype)
    e.set_geometry(solid, refl)
    e.init_state(variant="fp", rho0=c["rho0"])
    out = []
    done = 0
    for k in n:
        e.step(k - done)
        done = k
        out.append(e.get_state(("psi", "rho")))
    e.check_finite()
    e.close()
    return out


marks = [N // 10, N // 2, N]
a = run("fused", "f64", marks)
b = run("twopass", "f64", marks)
for m, x, y in zip(marks, a, b):
    print("fp64 fused == twopass after %5d steps: psi %s rho %s  (sum psi %.6f)" % (m, np.array_equal(x["psi"], y["psi"]),
                                                                                 np.array_equal(x["rho"], y["rho"]), x["psi"].sum()))
a = run("fused", "f32", marks)
b = run("fused", "f32", marks)
for m, x, y in zip(marks, a, b):
    print("fp32 fused run 1 == run 2 after %5d steps: psi %s rho %s" % (m, np.array_equal(x["psi"], y["psi"]), np.array_equal(x["rho"], y["rho"])))
