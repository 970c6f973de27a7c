"""Shared helpers for the parity tests: build oracle runs from the golden fixtures."""
import numpy as np

from oracle import oracle as orc

MACROS = ("psi", "rho", "ux", "uy", "p", "mu", "mix_tau")


def fixture_state(d, tag):
    """arrays of snapshot `tag` in the oracle's field names."""
    a = {k: d["%s_%s" % (tag, k)] for k in ("f", "g") + MACROS if "%s_%s" % (tag, k) in d}
    a["gx"] = d.get(tag + "_nabla_psix")
    a["gy"] = d.get(tag + "_nabla_psiy")
    a["lap"] = d.get(tag + "_nabla_psi2")
    return a


def circ_masks(d, prefix=""):
    return np.stack([d["%sside_%d" % (prefix, k)] for k in range(4)] +
                    [d["%sconcave_%d" % (prefix, k)] for k in range(4)] +
                    [d["%sconvex_%d" % (prefix, k)] for k in range(4)]).astype(np.uint8)


def fp_params(d, H=None, W=None):
    return orc.make_params(int(H or d["H"]), int(W or d["W"]), tau=float(d["c_tau"]), gamma=float(d["c_gamma"]),
                           a=float(d["c_a"]), kappa=float(d["c_kappa"]), Eta_n=float(d["c_Eta_n"]),
                           M=float(d["c_M"]), psi_wall=float(d["c_psi_wall"]), y_wall=0, outlet_f3_coef=2 / 3)


def fg_params(d):
    return orc.make_params(int(d["H"]), int(d["W"]), tau=float(d["c_tau"]), gamma=float(d["c_gamma"]),
                           a=float(d["c_a"]), kappa=float(d["c_kappa"]), Eta_n=float(d["c_Eta_n"]),
                           M=float(d["c_M"]), psi_wall=float(d["c_psi_wall"]), y_wall=1, outlet_f3_coef=1.5)


def va_params(d):
    # validation.py:36,117-118: a>0 and mu = a psi (psi^2-1); the generic form uses -a (exactly equal)
    return orc.make_params(int(d["H"]), int(d["W"]), tau=float(d["c_tau"]), gamma=float(d["c_gamma"]),
                           a=-float(d["c_a"]), kappa=float(d["c_kappa"]), Eta_n=float(d["c_Eta_n"]),
                           M=float(d["c_M"]), psi_wall=float(d["c_psi_wall"]), x_periodic=1, y_wall=1,
                           lap_order=1)


def fp_run(d, tag="s0"):
    H = int(d["H"])
    return orc.Run(fp_params(d), fixture_state(d, tag), mask=d["mask"], circ_masks=circ_masks(d), zou_he=1,
                   inlet_ux=d["inlet_ux"], outlet_ux=d["inlet_ux"])


def fg_run(d, tag="s0"):
    H = int(d["H"])
    u = np.full(H, float(d["c_u0"]))
    return orc.Run(fg_params(d), fixture_state(d, tag), mask=d["mask"], rect_corners=d["corners"],
                   wall_rows=(1, H - 2), zou_he=2, inlet_ux=u, outlet_ux=u)


def va_run(d, tag="s0"):
    a = fixture_state(d, tag)
    if a["gx"] is None:
        a["gx"] = a["gy"] = a["lap"] = np.zeros_like(a["psi"])
    return orc.Run(va_params(d), a, va=orc.va_consts(d["e"], d["w"], d["c_cs"], d["c_a"]))


def rel_err(x, ref):
    ref = np.asarray(ref, dtype=np.float64)
    scale = np.max(np.abs(ref))
    return float(np.max(np.abs(np.asarray(x, dtype=np.float64) - ref)) / (scale if scale > 0 else 1.0))
