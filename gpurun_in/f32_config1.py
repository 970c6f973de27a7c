"""fp32 vs fp64 on config 1 at the reference's full default length: field errors and interface displacement."""
import os, sys, numpy as np
sys.path.insert(0, os.getcwd())
from tests import helpers as hp
from tests.test_gpu_parity import _fp_full_inputs
from fingering_dynamics_b200 import Engine, geometry as geo, postprocess as pp
d = dict(np.load("tests/golden/fp_full_scalars.npz"))
H, W, mask, cls, P, s0 = _fp_full_inputs(d)
out = {}
for dt in ("f64", "f32"):
    e = Engine(H, W, tau=P.tau, gamma=P.gamma, a=P.a, kappa=P.kappa, Eta_n=P.Eta_n, M=P.M, psi_wall=P.psi_wall,
               zou_he="fp", inlet_ux=d["inlet_ux"], outlet_ux=d["inlet_ux"], dtype=dt)
    e.set_geometry(~mask, geo.reflect_bits_circle(cls[0:4], cls[4:8], cls[8:12]))
    e.set_state(f=s0["f"], g=s0["g"], psi=s0["psi"], rho=s0["rho"], ux=s0["ux"], uy=s0["uy"], p=s0["p"], mu=s0["mu"],
                mix_tau=s0["mix_tau"], nabla_psix=s0["gx"], nabla_psiy=s0["gy"], nabla_psi2=s0["lap"])
    res = {}
    done = 0
    for n in (100, 1000, 4000):
        e.step(n - done); done = n
        res[n] = e.get_state(("psi", "rho", "ux", "uy"))
    out[dt] = res
    e.close()
u0 = float(np.max(d["inlet_ux"]))
for n in (100, 1000, 4000):
    a, b = out["f32"][n], out["f64"][n]
    psi_b = np.where(mask, b["psi"], 0.0); psi_a = np.where(mask, a["psi"], 0.0)
    print(n, "max|dpsi| %.3e  max|drho| %.3e  max|du|/u0 %.3e  interface shift (fluid rows, cells) %.3e  sign mismatches where |psi|>1e-3: %d"
          % (np.abs(a["psi"] - b["psi"]).max(), np.abs(a["rho"] - b["rho"])[mask].max(),
             max(np.abs(a["ux"] - b["ux"])[mask].max(), np.abs(a["uy"] - b["uy"])[mask].max()) / u0,
             pp.interface_shift(np.where(mask, a["psi"], -1.0), np.where(mask, b["psi"], -1.0)),
             int(((np.sign(a["psi"]) != np.sign(b["psi"])) & (np.abs(b["psi"]) > 1e-3) & mask).sum())))
