"""Slab decomposition (fingering_dynamics_b200/slab.py).

CPU (gloo, world_size 2): the halo-exchange protocol of SlabRunner -- neighbour selection, send/recv pairing,
ring closing -- driven with a tiny NumPy stepper that, like the real step, needs two ghost columns per side.
GPU: two engines on one device exchanging their halo blocks reproduce the single-slab run bit for bit.
"""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_slab_bounds_cover_the_grid():
    from fingering_dynamics_b200.slab import slab_bounds
    for W, n in ((8192, 8), (400, 3), (17, 4), (65536, 8)):
        edges = [slab_bounds(W, n, r) for r in range(n)]
        assert edges[0][0] == 0 and edges[-1][1] == W
        assert all(edges[r][1] == edges[r + 1][0] for r in range(n - 1))
        sizes = [b - a for a, b in edges]
        assert max(sizes) - min(sizes) <= 1


class ToyEngine:
    """columns [x0,x1) of a (H, W) field plus 2 ghost columns per side; one step = a 5-point-wide
    stencil in x (needs BOTH ghost columns) followed by a roll in y."""

    def __init__(self, full, x0, x1):
        import torch
        self.x0, self.x1 = x0, x1
        H = full.shape[0]
        self.a = torch.zeros((x1 - x0 + 4, H), dtype=torch.float64)   # [column][row]: halo blocks are contiguous
        self.a[2:-2] = torch.from_numpy(full[:, x0:x1].T.copy())

    def halo_tensors(self):
        return self.a[2:4], self.a[0:2], self.a[-4:-2], self.a[-2:]

    def step(self, n):
        import torch
        for _ in range(n):
            a = self.a
            new = a.clone()
            new[2:-2] = 0.4 * a[2:-2] + 0.2 * (a[1:-3] + a[3:-1]) + 0.1 * (a[0:-4] + a[4:])
            self.a = torch.roll(new, 1, dims=1)

    def get_state(self, names=None, **kw):
        return self.a[2:-2].numpy().T.copy()


def _toy_reference(full, n, periodic):
    a = full.copy()
    for _ in range(n):
        if periodic:
            p = np.concatenate([a[:, -2:], a, a[:, :2]], axis=1)
        else:
            p = np.pad(a, ((0, 0), (2, 2)))
        a = 0.4 * p[:, 2:-2] + 0.2 * (p[:, 1:-3] + p[:, 3:-1]) + 0.1 * (p[:, 0:-4] + p[:, 4:])
        a = np.roll(a, 1, axis=0)
    return a


def _worker(rank, world, port, periodic, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    from fingering_dynamics_b200.slab import SlabRunner, slab_bounds
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(5)
    full = rng.random((6, 23))
    x0, x1 = slab_bounds(23, world, rank)
    eng = ToyEngine(full, x0, x1)
    run = SlabRunner(eng, rank, world, periodic=periodic)
    run.step(4)
    out = run.get_state()
    ref = _toy_reference(full, 4, periodic)[:, x0:x1]
    q.put((rank, bool(np.array_equal(out, ref)), float(np.abs(out - ref).max())))
    dist.destroy_process_group()


@pytest.mark.parametrize("periodic", [False, True])
def test_halo_exchange_protocol_gloo_world2(periodic):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 500) + (7 if periodic else 0)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, periodic, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    for rank, ok, err in res:
        assert ok, "rank %d differs from the single-domain run by %g" % (rank, err)


@pytest.mark.gpu
@pytest.mark.parametrize("name,nslab", [("fp_small", 2), ("fp_small", 3), ("fg_small", 2), ("va_small", 2)])
def test_slabs_on_one_gpu_are_bit_identical_to_the_single_slab_run(golden, name, nslab):
    """N engines (slabs) on one device, halos moved with device copies: same bits as one engine."""
    import torch
    from tests import helpers as hp
    from fingering_dynamics_b200.slab import slab_bounds, _DevBuf
    d = golden(name)
    W = int(d["W"])
    periodic = name.startswith("va")
    s0 = hp.state_for_engine(d, "s0")
    ref = hp.ENGINES[name](d)
    ref.set_state(**s0)
    ref.step(12)
    want = ref.get_state(("f", "g", "psi", "rho", "ux", "uy"))
    ref.close()

    engs = []
    for r in range(nslab):
        e = hp.ENGINES[name](d, slab=slab_bounds(W, nslab, r), external_halo=True)
        e.set_state(**s0)
        engs.append(e)
    dev = torch.device("cuda", 0)

    def view(ptr, n):
        return torch.as_tensor(_DevBuf(ptr, n), device=dev)

    def exchange():
        for e in engs:
            e.sync()
        hs = [e.halo_regions() for e in engs]
        for r in range(nslab):
            right = r + 1 if r + 1 < nslab else (0 if periodic else None)
            if right is None:
                continue
            view(hs[right].recv_lo, hs[r].bytes).copy_(view(hs[r].send_hi, hs[r].bytes))
            view(hs[r].recv_hi, hs[r].bytes).copy_(view(hs[right].send_lo, hs[r].bytes))
        torch.cuda.synchronize()

    for _ in range(12):
        exchange()
        for e in engs:
            e.step(1)
    exchange()
    for r, e in enumerate(engs):
        x0, x1 = slab_bounds(W, nslab, r)
        got = e.get_state(("f", "g", "psi", "rho", "ux", "uy"))
        for k in got:
            assert np.array_equal(got[k], want[k][..., x0:x1]), (name, r, k)
        e.close()


@pytest.mark.gpu
@pytest.mark.parametrize("name,nslab,kernel", [("fp_small", 2, "fused"), ("fp_small", 3, "fused"), ("fg_small", 2, "twopass"),
                                                ("va_small", 2, "fused"), ("va_small", 3, "fused")])
def test_peer_halo_engines_are_bit_identical_to_the_single_slab_run(golden, name, nslab, kernel):
    """halo exchange fused into the step kernel (stores into the neighbours' ghost columns + stream flags):
    N engines in one process on one device, many steps per call, same bits as one engine."""
    from tests import helpers as hp
    from fingering_dynamics_b200.slab import slab_bounds
    d = golden(name)
    W = int(d["W"])
    periodic = name.startswith("va")
    s0 = hp.state_for_engine(d, "s0")
    ref = hp.ENGINES[name](d)
    ref.set_state(**s0)
    ref.step(12)
    want = ref.get_state(("f", "g", "psi", "rho", "ux", "uy"))
    ref.close()

    engs = [hp.ENGINES[name](d, slab=slab_bounds(W, nslab, r), external_halo=True, kernel=kernel) for r in range(nslab)]
    infos = [e.peer_export() for e in engs]
    for r, e in enumerate(engs):
        left = r - 1 if r > 0 else (nslab - 1 if periodic else None)
        right = r + 1 if r < nslab - 1 else (0 if periodic else None)
        if left is not None:
            e.peer_attach(0, infos[left])
        if right is not None:
            e.peer_attach(1, infos[right])
    for e in engs:
        e.set_state(**s0)
    for e in engs:
        e.sync()
    for n in (1, 4, 7):            # many steps per call, no host involvement in between
        for e in engs:
            e.step(n)
    for r, e in enumerate(engs):
        x0, x1 = slab_bounds(W, nslab, r)
        got = e.get_state(("f", "g", "psi", "rho", "ux", "uy"))
        for k in got:
            assert np.array_equal(got[k], want[k][..., x0:x1]), (name, r, k)
    for e in engs:
        e.close()


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", ["f64", "f32"])
@pytest.mark.parametrize("periodic", [False, True])
def test_peer_halo_slabs_of_a_porous_grid_match_the_single_slab_run(dtype, periodic):
    """256 x 192 synthetic porous medium in three slabs of 64 columns with the halo exchange fused into the step kernel:
    same bits as one engine, for fp64 (k_fused) and for fp32 (the packed two-row kernel: its peer stores, its face CTAs
    on the first / last slab only, marching CTAs that start at a slab edge)."""
    from fingering_dynamics_b200 import Engine, synthetic as syn
    from fingering_dynamics_b200.slab import slab_bounds
    H, W, nslab = 256, 192, 3
    c = syn.fp_constants(H)
    solid, refl = syn.porous_geometry(H, W)
    st = syn.fp_initial_state(solid, c)
    kw = dict(tau=c["tau"], gamma=c["gamma"], a=c["a"], kappa=c["kappa"], Eta_n=c["Eta_n"], M=c["M"], psi_wall=c["psi_wall"],
              dtype=dtype)
    if periodic:
        kw.update(zou_he="none", x_periodic=True)
    else:
        kw.update(zou_he="fp", inlet_ux=c["inlet_ux"], outlet_ux=c["outlet_ux"])
    ref = Engine(H, W, **kw)
    ref.set_geometry(solid, refl)
    ref.set_state(**st)
    ref.step(12)
    want = ref.get_state(("f", "g", "psi", "rho", "ux", "uy"))
    ref.close()

    engs = [Engine(H, W, slab=slab_bounds(W, nslab, r), external_halo=True, **kw) for r in range(nslab)]
    for e in engs:
        e.set_geometry(solid, refl)
    infos = [e.peer_export() for e in engs]
    for r, e in enumerate(engs):
        left = r - 1 if r > 0 else (nslab - 1 if periodic else None)
        right = r + 1 if r < nslab - 1 else (0 if periodic else None)
        if left is not None:
            e.peer_attach(0, infos[left])
        if right is not None:
            e.peer_attach(1, infos[right])
    for e in engs:
        e.set_state(**st)
    for e in engs:
        e.sync()
    for n in (1, 4, 7):
        for e in engs:
            e.step(n)
    for r, e in enumerate(engs):
        x0, x1 = slab_bounds(W, nslab, r)
        got = e.get_state(("f", "g", "psi", "rho", "ux", "uy"))
        for k in got:
            assert np.array_equal(got[k], want[k][..., x0:x1]), (dtype, periodic, r, k)
    for e in engs:
        e.close()


# ---------------------------------------------------------------------------------------------------------------
# the path the multi-GPU bench times: one PROCESS per slab, neighbours' lattices and flag words mapped with CUDA IPC
# (fdlbm_peer_export / fdlbm_peer_attach, the cudaIpcOpenMemHandle branch), steps ordered across processes by
# cuStreamWriteValue32 / cuStreamWaitValue32.  All processes share cuda:0, so this runs on the 1-GPU test box.
# ---------------------------------------------------------------------------------------------------------------
ALL12 = ("f", "g", "psi", "rho", "ux", "uy", "p", "mu", "mix_tau", "nabla_psix", "nabla_psiy", "nabla_psi2")
OUT6 = ("f", "g", "psi", "rho", "ux", "uy")


def _porous_case(dtype, periodic):
    from fingering_dynamics_b200 import synthetic as syn
    H, W = 256, 192
    c = syn.fp_constants(H)
    solid, refl = syn.porous_geometry(H, W)
    st = syn.fp_initial_state(solid, c)
    kw = dict(tau=c["tau"], gamma=c["gamma"], a=c["a"], kappa=c["kappa"], Eta_n=c["Eta_n"], M=c["M"], psi_wall=c["psi_wall"],
              dtype=dtype)
    if periodic:
        kw.update(zou_he="none", x_periodic=True)
    else:
        kw.update(zou_he="fp", inlet_ux=c["inlet_ux"], outlet_ux=c["outlet_ux"])
    return H, W, solid, refl, st, kw


def _ipc_worker(rank, world, port, dtype, periodic, q):
    try:
        os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
        sys.path.insert(0, ROOT)
        import torch.distributed as dist
        from fingering_dynamics_b200 import Engine
        from fingering_dynamics_b200.slab import SlabRunner, slab_bounds
        dist.init_process_group("gloo", rank=rank, world_size=world)
        H, W, solid, refl, st, kw = _porous_case(dtype, periodic)
        x0, x1 = slab_bounds(W, world, rank)
        sl = (Ellipsis, slice(x0, x1))
        bad = []

        def same(got, want, tag):
            for k in got:
                if not np.array_equal(got[k], want[k][sl]):
                    bad.append("%s:%s" % (tag, k))

        # the single-engine run every rank compares its slab with
        ref = Engine(H, W, **kw)
        ref.set_geometry(solid, refl)
        ref.set_state(**st)
        ref.step(12)
        want12 = ref.get_state(ALL12)
        ref.set_state(**want12)          # restart from a mid-run state: populations inside solids are non-zero now
        ref.step(5)
        want17 = ref.get_state(OUT6)
        ref.step(3)
        want20 = ref.get_state(OUT6)
        ref.close()

        eng = Engine(H, W, slab=(x0, x1), external_halo=True, **kw)
        eng.set_geometry(solid, refl)
        run = SlabRunner(eng, rank, world, periodic=periodic, halo="peer")   # IPC handles travel over gloo
        run.set_state(**st)
        for n in (1, 4, 7):              # many steps per call: only the stream flags order the processes
            run.step(n)
        same(run.get_state(OUT6), want12, "step12")
        run.set_state(**want12)          # k_collide_first must push the solid cells of the edge columns too
        run.step(5)
        same(run.get_state(OUT6), want17, "restart17")
        blob = run.checkpoint()          # waits for the neighbours' halo stores of step 17
        run.step(3)
        same(run.get_state(OUT6), want20, "step20")
        run.restore(blob)
        run.step(3)
        same(run.get_state(OUT6), want20, "restored20")
        eng.sync()
        dist.barrier()
        eng.close()
        dist.destroy_process_group()
        q.put((rank, bad))
    except Exception as ex:  # noqa: BLE001
        import traceback
        q.put((rank, ["exception: %s\n%s" % (ex, traceback.format_exc())]))


@pytest.mark.gpu
@pytest.mark.parametrize("world,dtype,periodic", [(2, "f64", False), (3, "f64", False), (2, "f32", False), (3, "f32", False),
                                                   (2, "f64", True), (3, "f32", True)])
def test_peer_halo_across_processes_is_bit_identical_to_one_engine(world, dtype, periodic):
    """one process per slab on cuda:0, IPC-mapped halos: step(1); step(4); step(7), a restart from the step-12 state,
    a checkpoint / restore in peer mode -- every slab bitwise equal to the single-engine run"""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29800 + (os.getpid() % 400) + 3 * world + (1 if periodic else 0) + (50 if dtype == "f32" else 0)
    procs = [ctx.Process(target=_ipc_worker, args=(r, world, port, dtype, periodic, q)) for r in range(world)]
    for p in procs:
        p.start()
    try:
        res = [q.get(timeout=300) for _ in procs]
    finally:
        for p in procs:
            p.join(timeout=30)
            if p.is_alive():
                p.kill()
    for rank, bad in sorted(res):
        assert not bad, "rank %d: %s" % (rank, "; ".join(bad))


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", ["f64", "f32"])
def test_peer_slabs_restart_from_a_mid_run_state(dtype):
    """same-process variant of the restart above: get_state at step 12 -> set_state -> 5 steps on three peer slabs"""
    from fingering_dynamics_b200 import Engine
    from fingering_dynamics_b200.slab import slab_bounds
    H, W, solid, refl, st, kw = _porous_case(dtype, False)
    nslab = 3
    ref = Engine(H, W, **kw)
    ref.set_geometry(solid, refl)
    ref.set_state(**st)
    ref.step(12)
    mid = ref.get_state(ALL12)
    assert np.abs(mid["f"][:, solid != 0]).max() > 0     # the point of the test: solids hold streamed populations
    ref.set_state(**mid)
    ref.step(5)
    want = ref.get_state(OUT6)
    ref.close()
    engs = [Engine(H, W, slab=slab_bounds(W, nslab, r), external_halo=True, **kw) for r in range(nslab)]
    for e in engs:
        e.set_geometry(solid, refl)
    infos = [e.peer_export() for e in engs]
    for r, e in enumerate(engs):
        if r > 0:
            e.peer_attach(0, infos[r - 1])
        if r < nslab - 1:
            e.peer_attach(1, infos[r + 1])
    for e in engs:
        e.set_state(**mid)
    for e in engs:
        e.sync()
    for e in engs:
        e.step(5)
    for r, e in enumerate(engs):
        x0, x1 = slab_bounds(W, nslab, r)
        got = e.get_state(OUT6)
        for k in got:
            assert np.array_equal(got[k], want[k][..., x0:x1]), (dtype, r, k)
        e.close()
