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


# ------------------------------------------------------------------------------------------------
# product-side helpers (CUDA engine) -- imported lazily so CPU-only tests never touch the library
# ------------------------------------------------------------------------------------------------
ALL_FIELDS = ("f", "g", "psi", "rho", "ux", "uy", "p", "mu", "mix_tau", "nabla_psix", "nabla_psiy", "nabla_psi2")


def state_for_engine(d, tag):
    a = {k: d["%s_%s" % (tag, k)] for k in ("f", "g") + MACROS}
    for k in ("nabla_psix", "nabla_psiy", "nabla_psi2"):
        a[k] = d["%s_%s" % (tag, k)]
    return a


def fp_engine(d, **kw):
    from fingering_dynamics_b200 import Engine, geometry as geo
    H, W = int(d["H"]), int(d["W"])
    e = Engine(H, W, tau=float(d["c_tau"]), gamma=float(d["c_gamma"]), a=float(d["c_a"]), kappa=float(d["c_kappa"]),
               Eta_n=float(d["c_Eta_n"]), M=float(d["c_M"]), psi_wall=float(d["c_psi_wall"]), zou_he="fp",
               inlet_ux=d["inlet_ux"], outlet_ux=d["inlet_ux"], outlet_f3_coef=2 / 3, **kw)
    side = [d["side_%d" % k] for k in range(4)]
    cave = [d["concave_%d" % k] for k in range(4)]
    vex = [d["convex_%d" % k] for k in range(4)]
    e.set_geometry(~d["mask"], geo.reflect_bits_circle(side, cave, vex))
    return e


def corner_dicts(corners):
    return [{"top_left": (c[0], c[1]), "bottom_left": (c[2], c[3]), "top_right": (c[4], c[5]),
             "bottom_right": (c[6], c[7])} for c in np.asarray(corners).tolist()]


def fg_engine(d, **kw):
    from fingering_dynamics_b200 import Engine, geometry as geo
    H, W = int(d["H"]), int(d["W"])
    u = np.full(H, float(d["c_u0"]))
    e = Engine(H, W, tau=float(d["c_tau"]), gamma=float(d["c_gamma"]), a=float(d["c_a"]), kappa=float(d["c_kappa"]),
               Eta_n=float(d["c_Eta_n"]), M=float(d["c_M"]), psi_wall=float(d["c_psi_wall"]), zou_he="fg",
               psi_y_wall=True, inlet_ux=u, outlet_ux=u, outlet_f3_coef=1.5, **kw)
    refl = geo.reflect_bits_rect(corner_dicts(d["corners"]), H, W) | geo.reflect_bits_wall_rows(H, W, 1, H - 2)
    e.set_geometry(~d["mask"], refl)
    return e


def va_engine(d, **kw):
    from fingering_dynamics_b200 import Engine, geometry as geo
    H, W = int(d["H"]), int(d["W"])
    e = Engine(H, W, tau=float(d["c_tau"]), gamma=float(d["c_gamma"]), a=-float(d["c_a"]), kappa=float(d["c_kappa"]),
               Eta_n=float(d["c_Eta_n"]), M=float(d["c_M"]), psi_wall=float(d["c_psi_wall"]), zou_he="none",
               psi_y_wall=True, x_periodic=True, **kw)
    e.set_geometry(np.zeros((H, W), dtype=np.uint8), geo.reflect_bits_wall_rows(H, W, 0, H - 1))
    return e


ENGINES = {"fp_small": fp_engine, "fg_small": fg_engine, "va_small": va_engine, "va_small_wet": va_engine}
ORACLES = {"fp_small": fp_run, "fg_small": fg_run, "va_small": va_run, "va_small_wet": va_run}
