"""Build-container only: module-level constants and public names of the drop-in twins against the reference modules
(the drivers read H, W, tau, kappa, ... from module globals at call time, so a twin must define the same ones)."""
import contextlib
import importlib
import io
import os
import re

import numpy as np
import pytest

REF = "/root/reference/lattice_boltzmann"
pytestmark = pytest.mark.skipif(not os.path.isdir(REF), reason="the reference tree only exists in the build container")


@pytest.fixture(scope="module")
def mg():
    with contextlib.redirect_stdout(io.StringIO()):
        from tests.golden import make_golden
    return make_golden


def _numeric_globals(ns):
    return {k: v for k, v in ns.items() if not k.startswith("_") and isinstance(v, (int, float, np.integer, np.floating))
            and not isinstance(v, bool)}


@pytest.mark.parametrize("name", ["fingering_periodic", "fingering", "validation"])
def test_module_constants_equal_the_reference(mg, name):
    ref = {"fingering_periodic": mg.FP, "fingering": mg.FG, "validation": mg.VA}[name]
    twin = importlib.import_module("fingering_dynamics_b200.lattice_boltzmann." + name)
    want = _numeric_globals(vars(ref))
    assert len(want) >= 15
    for k, v in want.items():
        assert hasattr(twin, k), "%s.%s is missing in the twin" % (name, k)
        assert getattr(twin, k) == v, (name, k, getattr(twin, k), v)
    for k in ("Compute", "stream", "main"):
        assert callable(getattr(twin, k))


def test_gpu_variant_constants_equal_the_reference_source():
    """fingering_periodic_gpu.py needs CuPy to import; its constant block (plain assignments above `class Compute`) is
    evaluated on its own instead."""
    import math
    src = open(os.path.join(REF, "fingering_periodic_gpu.py")).read()
    head = src.split("class Compute")[0]
    lines = [ln for ln in head.splitlines() if re.match(r"^[A-Za-z_][A-Za-z_0-9]*\s*=", ln)]
    ns = {"math": math, "np": np}
    for ln in lines:
        try:
            exec(ln, ns)
        except Exception:  # noqa: BLE001 -- an assignment that needs cupy: not a numeric constant
            pass
    want = _numeric_globals({k: v for k, v in ns.items() if k not in ("math", "np")})
    twin = importlib.import_module("fingering_dynamics_b200.lattice_boltzmann.fingering_periodic_gpu")
    checked = 0
    for k, v in want.items():
        if hasattr(twin, k):
            assert getattr(twin, k) == v, (k, getattr(twin, k), v)
            checked += 1
    assert checked >= 20 and set(want) <= set(vars(twin)), sorted(set(want) - set(vars(twin)))


def test_createblock_and_bounce_back_surface(mg):
    from fingering_dynamics_b200.lattice_boltzmann import bounce_back, create_block
    for ref, twin, cls in ((mg.CB, create_block, "Createblock"), (mg.BB, bounce_back, "Bounce_back")):
        rc, tc = getattr(ref, cls), getattr(twin, cls)
        for m in [k for k, v in vars(rc).items() if callable(v) and not k.startswith("_")]:
            assert callable(getattr(tc, m, None)), "%s.%s is missing in the twin" % (cls, m)


@pytest.mark.parametrize("name", ["fingering_periodic", "fingering", "validation"])
def test_functions_and_compute_methods_of_the_reference_exist_in_the_twin(mg, name):
    import inspect
    ref = {"fingering_periodic": mg.FP, "fingering": mg.FG, "validation": mg.VA}[name]
    twin = importlib.import_module("fingering_dynamics_b200.lattice_boltzmann." + name)
    ref_methods = {k for k, v in vars(ref.Compute).items() if callable(v) and not k.startswith("__")}
    twin_methods = {k for k in dir(twin.Compute) if callable(getattr(twin.Compute, k))}
    assert not (ref_methods - twin_methods), sorted(ref_methods - twin_methods)
    ref_funcs = {k for k, v in vars(ref).items() if inspect.isfunction(v) and v.__module__ == ref.__name__}
    twin_funcs = {k for k, v in vars(twin).items() if callable(v)}
    assert not (ref_funcs - twin_funcs), sorted(ref_funcs - twin_funcs)


def test_validation_power_law_matches_the_reference(mg):
    from fingering_dynamics_b200.lattice_boltzmann import validation as V
    for t in (0.0, 0.05, 0.2, 0.37):
        assert float(V.power_law(None, t)) == float(mg.VA.power_law(None, t))


def _required(f):
    import inspect
    return [p.name for p in inspect.signature(f).parameters.values()
            if p.default is inspect._empty and p.kind in (p.POSITIONAL_ONLY, p.POSITIONAL_OR_KEYWORD)]


@pytest.mark.parametrize("name", ["fingering_periodic", "fingering", "validation", "create_block", "bounce_back"])
def test_required_arguments_equal_the_reference(mg, name):
    """same positional arguments, by name and order, for every function and method both modules define"""
    import inspect
    ref = {"fingering_periodic": mg.FP, "fingering": mg.FG, "validation": mg.VA, "create_block": mg.CB,
           "bounce_back": mg.BB}[name]
    twin = importlib.import_module("fingering_dynamics_b200.lattice_boltzmann." + name)
    checked = 0
    for cname in ("Compute", "Createblock", "Bounce_back"):
        if hasattr(ref, cname):
            for k, v in vars(getattr(ref, cname)).items():
                if callable(v) and hasattr(getattr(twin, cname), k):
                    assert _required(v) == _required(getattr(getattr(twin, cname), k)), (cname, k)
                    checked += 1
    for k, v in vars(ref).items():
        if inspect.isfunction(v) and v.__module__ == ref.__name__ and hasattr(twin, k):
            assert _required(v) == _required(getattr(twin, k)), k
            checked += 1
    assert checked >= 4
