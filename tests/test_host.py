"""CPU-side checks: the C-ABI library builds, loads and exports what include/fdlbm.h declares; it fails
loudly without a GPU; and the host-side geometry folding (class masks -> reflect bits) agrees with the
reference's class tables as restated by the oracle."""
import ctypes
import os
import re

import numpy as np
import pytest

from oracle import oracle as orc
from tests import helpers as hp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    txt = open(os.path.join(ROOT, "include", "fdlbm.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(fdlbm_[a-z0-9_]+)\s*\(", txt)))


def test_library_exports_every_declared_symbol():
    from fingering_dynamics_b200 import _native as nat
    L = nat.lib()
    syms = _header_symbols()
    assert len(syms) >= 20
    for s in syms:
        assert hasattr(L, s), "libfdlbm.so does not export " + s
        assert s in nat._SIGS, "binding missing for " + s
    assert L.fdlbm_abi_version() == 1


def test_no_cpu_fallback_without_a_device():
    """On a box without CUDA, creating an engine must raise (never silently compute on the CPU)."""
    from fingering_dynamics_b200 import _native as nat
    if nat.lib().fdlbm_device_count() > 0:
        pytest.skip("a CUDA device is present")
    from fingering_dynamics_b200 import Engine
    with pytest.raises(nat.FdlbmError) as ei:
        Engine(16, 16, tau=0.8, gamma=1.0, a=-0.04, kappa=0.09, Eta_n=0.1, M=20.0, psi_wall=0.0, x_periodic=True)
    assert "CUDA" in str(ei.value) or "device" in str(ei.value)
    rc = nat.lib().fdlbm_op_stream(8, 8, np.zeros((9, 8, 8)).ctypes.data, np.zeros((9, 8, 8)).ctypes.data)
    assert rc < 0


def test_product_package_never_touches_the_oracle():
    """The oracle is test infrastructure: no source of the shipped package may import, load or name it."""
    pkg = os.path.join(ROOT, "fingering_dynamics_b200")
    hits = []
    for d, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(d, fn), errors="replace").read()
                if re.search(r"oracle|libfd_oracle", txt):
                    hits.append(os.path.relpath(os.path.join(d, fn), ROOT))
    assert not hits, hits


def test_config_validation_messages():
    from fingering_dynamics_b200 import _native as nat
    cfg = nat.Config()
    h = ctypes.c_void_p()
    rc = nat.lib().fdlbm_create(ctypes.byref(cfg), ctypes.byref(h))
    assert rc == -1 and b"grid too small" in nat.lib().fdlbm_last_error()
    assert nat.lib().fdlbm_step(None, 1) == -1
    assert nat.lib().fdlbm_iterations(None) == -1


def _apply_bits(bits, fb, gb, f, g):
    opp = [0, 3, 4, 1, 2, 7, 8, 5, 6]
    for i in range(1, 9):
        m = (bits >> (i - 1)) & 1 == 1
        f[i][m] = fb[opp[i]][m]
        g[i][m] = gb[opp[i]][m]


def test_reflect_bits_equal_the_class_tables(golden):
    """pull + reflect bits == stream + halfway_bounceback_* (bit-identical), for circles, rectangles,
    wall rows and left_boundary -- against the golden outputs of the reference itself."""
    from fingering_dynamics_b200 import geometry as geo
    d = golden("ops")
    H, W = int(d["H"]), int(d["W"])
    fb, gb = d["f_in"], d["g_in"]

    bits = geo.reflect_bits_circle([d["circ_side_%d" % k] for k in range(4)],
                                   [d["circ_concave_%d" % k] for k in range(4)],
                                   [d["circ_convex_%d" % k] for k in range(4)])
    f, g = d["f_stream"].copy(), d["g_stream"].copy()
    _apply_bits(bits, fb, gb, f, g)
    assert np.array_equal(f, d["f_bb_circle"]) and np.array_equal(g, d["g_bb_circle"])

    bits = geo.reflect_bits_rect(hp.corner_dicts(d["rect_corners"]), H, W)
    f, g = d["f_stream"].copy(), d["g_stream"].copy()
    _apply_bits(bits, fb, gb, f, g)
    assert np.array_equal(f, d["f_bb_rect"]) and np.array_equal(g, d["g_bb_rect"])
    bits |= geo.reflect_bits_wall_rows(H, W, 1, H - 2)
    f, g = d["f_stream"].copy(), d["g_stream"].copy()
    _apply_bits(bits, fb, gb, f, g)
    assert np.array_equal(f, d["f_bb_rect_walls"]) and np.array_equal(g, d["g_bb_rect_walls"])

    bits = geo.reflect_bits_wall_rows(H, W, 0, H - 1)
    f, g = d["f_stream"].copy(), d["g_stream"].copy()
    _apply_bits(bits, fb, gb, f, g)
    assert np.array_equal(f, d["f_bb_va"]) and np.array_equal(g, d["g_bb_va"])

    bits = geo.reflect_bits_left_boundary(H, W, 3)
    f, g = d["f_stream"].copy(), d["g_stream"].copy()
    _apply_bits(bits, fb, gb, f, g)
    assert np.array_equal(f, d["f_left_boundary"]) and np.array_equal(g, d["g_left_boundary"])


def test_solid_flag_ignores_doubly_covered_cells(golden):
    from fingering_dynamics_b200 import geometry as geo
    d = golden("geometry")
    bpa = d["circ_overlap_block_psi_all"]
    assert bpa.max() == 2
    s = geo.solid_from_block_psi(bpa)
    assert s.dtype == np.uint8 and np.array_equal(s == 1, bpa == 1) and s[bpa == 2].sum() == 0


def test_synthetic_generator_windows_and_initial_state():
    """a slab's geometry window equals the same columns of the global geometry, and the product's
    host-side initial state (synthetic.fp_initial_state) equals the oracle's restatement of
    Compute.__init__ (fingering_periodic.py:90-121) bit for bit."""
    from fingering_dynamics_b200 import synthetic as syn
    from fingering_dynamics_b200.slab import slab_bounds
    H, W = 160, 400
    solid, refl = syn.porous_geometry(H, W)
    assert 0.05 < solid.mean() < 0.3 and refl.any()
    cover = 0
    for rank in range(3):
        x0, x1 = slab_bounds(W, 3, rank)
        lo, hi = max(0, x0 - 2), min(W, x1 + 2)
        s, r = syn.porous_geometry(H, W, col0=lo, ncols=hi - lo)
        assert np.array_equal(s, solid[:, lo:hi]) and np.array_equal(r, refl[:, lo:hi])
        cover += x1 - x0
    assert cover == W
    c = syn.fp_constants(H)
    st = syn.fp_initial_state(solid, c)
    P = orc.make_params(H, W, tau=c["tau"], gamma=c["gamma"], a=c["a"], kappa=c["kappa"], Eta_n=c["Eta_n"], M=c["M"],
                        psi_wall=c["psi_wall"])
    ref = orc.fp_initial_state(P, solid == 0)
    for k in ("f", "g", "psi", "rho", "ux", "uy", "p", "mu", "mix_tau"):
        assert np.array_equal(st[k], ref[k]), k
    # a slab's initial state is the same columns
    x0, x1 = slab_bounds(W, 3, 1)
    st1 = syn.fp_initial_state(np.ascontiguousarray(solid[:, x0:x1]), c, col0=x0)
    assert np.array_equal(st1["f"], st["f"][:, :, x0:x1]) and np.array_equal(st1["psi"], st["psi"][:, x0:x1])


def test_fp_constants_match_reference_fixture(golden):
    """the restated constant block of fingering_periodic.py:15-40 gives the reference's floats exactly"""
    from fingering_dynamics_b200 import synthetic as syn
    d = golden("fp_full_scalars")
    c = syn.fp_constants(400)
    for k in ("tau", "gamma", "a", "kappa", "Eta_n", "M", "u0", "psi_wall"):
        assert c[k] == float(d["c_" + k]), k
    assert np.array_equal(c["inlet_ux"], d["inlet_ux"])


def test_contact_angle_estimator_on_analytic_caps():
    from fingering_dynamics_b200 import postprocess as pp
    H, W = 120, 200
    yy, xx = np.mgrid[:H, :W]
    for th in (60.0, 90.0, 120.0):
        R = 50.0
        yc = -0.5 - R * np.cos(np.radians(th))   # the wall is half a cell below row 0
        psi = np.tanh((R - np.sqrt((xx - 100.3) ** 2 + (yy - yc) ** 2)) / 1.5)
        assert abs(pp.droplet_contact_angle(psi) - th) < 0.5
    assert pp.interface_shift(psi, np.roll(psi, 1, axis=1)) == pytest.approx(1.0, abs=1e-6)


def test_frame_file_format_matches_openmovie(tmp_path):
    """openmovie.py:14-15 unpickles `cc` and indexes cc[i] as an (H, W) array"""
    import pickle
    from fingering_dynamics_b200.lattice_boltzmann._compute import save_frames
    frames = [np.full((4, 5), float(k)) for k in range(3)]
    p = tmp_path / "f_list_test.txt"
    save_frames(frames, p)
    cc = pickle.load(open(p, "rb"))
    assert len(cc) == 3 and cc[1].shape == (4, 5) and cc[2][0, 0] == 2.0


def test_step_kernels_keep_their_register_budget():
    """the fused step kernels run three CTAs of 128 threads per SM: at most 168 registers and NO spills (a spilled
    value reloaded inside the marching loop waits behind every global load in flight: measured -12 % with 24 bytes)"""
    import shutil
    import subprocess
    from fingering_dynamics_b200 import _native as nat
    tool = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(tool):
        pytest.skip("cuobjdump not available")
    out = subprocess.run([tool, "--dump-resource-usage", nat.build()], capture_output=True, text=True).stdout
    usage = {m[0]: (int(m[1]), int(m[2])) for m in re.findall(r"Function (\S+):\s*\n\s*REG:(\d+) STACK:(\d+)", out)}
    fused = {k: v for k, v in usage.items() if "k_fused" in k}
    assert any("k_fusedId" in k for k in fused) and any("k_fused_f32p" in k for k in fused), sorted(fused)
    for name, (reg, stack) in fused.items():
        assert reg <= 168, (name, reg)
        assert stack == 0, (name, stack)


def test_ctypes_structures_match_the_header_layout(tmp_path):
    """sizeof / offsetof of every struct of include/fdlbm.h as gcc sees them == the ctypes mirrors in _native.py."""
    import subprocess
    from fingering_dynamics_b200 import _native as nat
    pairs = {"fdlbm_config": nat.Config, "fdlbm_fields": nat.Fields, "fdlbm_halo": nat.Halo,
             "fdlbm_peer_info": nat.PeerInfo, "fdlbm_algebra_out": nat.AlgebraOut}
    lines = ['#include "%s"' % os.path.join(ROOT, "include", "fdlbm.h"), "#include <stdio.h>", "int main(void){"]
    for cname, cls in pairs.items():
        lines.append('printf("%s %%zu", sizeof(%s));' % (cname, cname))
        for fname, *_ in cls._fields_:
            lines.append('printf(" %%zu", offsetof(%s, %s));' % (cname, fname))
        lines.append('printf("\\n");')
    lines.append("return 0;}")
    src = tmp_path / "layout.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "layout"
    subprocess.run(["gcc", "-std=c99", str(src), "-o", str(exe)], check=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.strip().splitlines()
    assert len(out) == len(pairs)
    for line in out:
        name, size, *offs = line.split()
        cls = pairs[name]
        assert ctypes.sizeof(cls) == int(size), name
        assert [getattr(cls, f[0]).offset for f in cls._fields_] == [int(o) for o in offs], name


def test_engine_save_and_load_use_the_exact_path(tmp_path):
    """Engine.save / Engine.load around checkpoint() / restore(), without a device: the blob round-trips through the
    file name given (no '.npy' appended)."""
    from fingering_dynamics_b200 import Engine
    e = object.__new__(Engine)
    e._h = None
    blob = np.arange(1000, dtype=np.uint8)
    got = {}
    e.checkpoint = lambda: blob
    e.restore = lambda b: got.setdefault("blob", np.array(b))
    path = tmp_path / "run.ckpt"
    e.save(str(path))
    assert path.exists() and not (tmp_path / "run.ckpt.npy").exists()
    e.load(str(path))
    assert np.array_equal(got["blob"], blob) and got["blob"].dtype == np.uint8
