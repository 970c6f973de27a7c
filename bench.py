#!/usr/bin/env python
"""bench.py -- MLUPS of the two-phase D2Q9 step on synthetic porous domains (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--dtype f64|f32] [--impl ours|reference]

N=1 workload: BASELINE.json configs[3], "synthetic random-block porous medium 8192x2048" (W=8192 columns
along the flow, H=2048 rows), fingering_periodic.py's step variant, fp64.  N>1 (torchrun): the same slab per
GPU concatenated along the flow axis (weak scaling, W = 8192*N), 2-column halos pushed into the neighbours' peer
memory by the step kernel itself.  One "step" is one lattice-Boltzmann iteration over the whole grid;
value = H*W*K / t / 1e6 (all cells: solids are streamed too).

Timing (device, CUDA events on the engine stream, max over ranks): after W warm-up steps and a barrier, `reps`
windows are enqueued back to back with NO host synchronisation in between; each window is `align` untimed steps
(they put the ranks, which are only coupled through their neighbours' stream flags, into lock step after the
barrier's exit skew) followed by EXACTLY K timed steps between two events.  ms_per_step is the MEDIAN window
(every window is in `windows_ms`); `sustained` is one further window of >= 2 s, where the GPU sits at its power cap.

Extra legs on the same line: N=1 -- `fp32` (configs[3] says "fp64 and fp32"), `cpu_baseline`; N>1 -- `multi_gpu_check`
(a small grid over the real ranks against one engine, bit for bit) and `strong_c5` (configs[4], 32768x8192 split over
the N GPUs, with the 1-GPU figure of the same job measured on rank 0 in the same run).

--impl reference times the CPU arm: the oracle port (oracle/fd_oracle.c, OpenMP, all host threads) on a bounded
crop of the same workload.  The reference itself is NumPy and does not travel to the GPU box; where /root/reference
is reachable (the build container) the line also carries `numpy_reference_mlups` from the UNMODIFIED reference loop.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

H_DEFAULT, W_PER_GPU = 2048, 8192
C5_H, C5_W = 8192, 32768
B_ALG = {"f64": 288.0, "f32": 144.0}  # algorithmic bytes per lattice update: 18 populations read + written once
FP32_TOLERANCE = ("vs the fp64 engine / reference, relative to the field maximum: psi, rho <= 2e-5, u <= 2e-3 (40 steps of the "
                  "small cases and 6 steps at 8192x2048, tests/test_gpu_parity.py, tests/test_gpu_fullsize.py); over the "
                  "reference's full config-1 run (4000 steps) the psi = 0 interface moves by 1.7e-3 cells")


def peak_hbm():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def copy_bandwidth_here():
    """this box's own copy bandwidth, measured the way MEASURED_PEAKS.json says the pod's was (b.copy_(a) over 1 Gi bf16
    elements, read + write bytes, best of 10, CUDA events): boxes differ from the pod figure by a few per cent"""
    import torch
    try:
        a = torch.empty(1 << 30, dtype=torch.bfloat16, device="cuda")
        b = torch.empty_like(a)
        best = 0.0
        for _ in range(12):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            b.copy_(a)
            e1.record()
            torch.cuda.synchronize()
            best = max(best, 2 * a.numel() * 2 / (e0.elapsed_time(e1) * 1e-3) / 1e9)
        del a, b
        return best
    except Exception:
        return None


class ClockSampler:
    """nvidia-smi SM clocks / throttle reasons sampled during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.rows, self.proc, self.index = [], None, index

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "20"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.thr = threading.Thread(target=self._read, daemon=True)
            self.thr.start()
        except Exception:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [c.strip() for c in line.split(",")]))

    def __exit__(self, *exc):
        if self.proc:
            time.sleep(0.1)
            self.proc.terminate()
            self.thr.join(timeout=2)

    def summary(self, t0, t1):
        """samples taken in [t0, t1] (host clock); a region shorter than the sampling period falls back to the
        samples closest to it and says so"""
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        rows = [r for (ts, r) in self.rows if t0 <= ts <= t1 + 0.05]
        how = "inside the timed region"
        if len(rows) < 3:
            rows = [r for (ts, r) in sorted(self.rows, key=lambda x: abs(x[0] - 0.5 * (t0 + t1)))[:5]]
            how = "the 5 samples closest to a timed region shorter than 3 sampling periods"
        sm, mx, reasons = [], [], set()
        for r in rows:
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
                for n, v in zip(names, r[2:6]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                pass
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons),
                "samples": len(sm), "window": how}


# ------------------------------------------------------------------------------------------------------------
# CPU legs
# ------------------------------------------------------------------------------------------------------------
def cpu_arm(H, W, steps, warmup, threads):
    """The oracle port on a (H, W) instance of the synthetic workload; returns MLUPS."""
    from oracle import oracle as orc
    from fingering_dynamics_b200 import synthetic as syn
    orc.set_threads(threads)
    c = syn.fp_constants(H)
    circles = syn.porous_circles(H, W)
    from fingering_dynamics_b200.lattice_boltzmann.create_block import Createblock
    bpa, side, cave, vex = Createblock(H, W).setCirleblock(circles)
    mask = bpa != 1
    P = orc.make_params(H, W, tau=c["tau"], gamma=c["gamma"], a=c["a"], kappa=c["kappa"], Eta_n=c["Eta_n"], M=c["M"],
                        psi_wall=c["psi_wall"])
    s0 = orc.fp_initial_state(P, mask)
    run = orc.Run(P, s0, mask=mask, circ_masks=np.stack(list(side) + list(cave) + list(vex)).astype(np.uint8), zou_he=1,
                  inlet_ux=c["inlet_ux"], outlet_ux=c["outlet_ux"])
    run.iterate(warmup)
    t0 = time.perf_counter()
    run.iterate(steps)
    dt = time.perf_counter() - t0
    return H * W * steps / dt / 1e6, dt


def numpy_reference_arm(steps=12, warmup=2, ref_dir="/root/reference/lattice_boltzmann"):
    """The UNMODIFIED reference (fingering_periodic.py, config 1: 400x400, 90 circles) for `steps` iterations of its own
    loop body, where the reference is reachable (the build container; it does not travel to the GPU box).  NumPy
    ufuncs are single-threaded: 1 core.  Returns None when the reference is not importable."""
    if not os.path.isdir(ref_dir):
        return None
    gold = os.path.join(ROOT, "tests", "golden")
    try:
        sys.path.insert(0, gold)
        import make_golden as mg   # the harness that imports the reference in place (matplotlib stubbed)
        return mg.time_fp_default(steps, warmup)
    except Exception as ex:  # noqa: BLE001
        return {"error": "%s: %s" % (type(ex).__name__, ex)}
    finally:
        if gold in sys.path:
            sys.path.remove(gold)


def reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    Hs, Ws = H_DEFAULT, 1024  # bounded crop of the 2048x8192 workload (same generator, same H)
    steps = max(1, min(args.steps, 40))
    warm = max(1, min(args.warmup, 5))
    v, dt = cpu_arm(Hs, Ws, steps, warm, cores)
    line = {"impl": "reference", "metric": "MLUPS (two-phase D2Q9)", "value": v, "unit": "MLUPS", "n_gpus": args.gpus,
            "steps": steps, "warmup": warm, "ms_per_step": dt / steps * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "synthetic random-block porous medium %dx%d (W x H), fingering_periodic step variant"
                                   % (W_PER_GPU * max(1, args.gpus), H_DEFAULT), "grid_W": W_PER_GPU * max(1, args.gpus),
                       "grid_H": H_DEFAULT, "obstacles": "circles r 8-12 on a jittered 40-pitch lattice, seed 1234",
                       "cpu_sample": "each step runs on a %dx%d crop of the workload (same generator, same H)" % (Ws, Hs)},
            "cpu_baseline": {"value": v, "unit": "MLUPS", "cores": cores, "kind": "port",
                             "sample": "%d steps of a %dx%d crop of the workload, oracle/fd_oracle.c with OpenMP" % (steps, Ws, Hs)},
            "e2e": {"value": v, "unit": "MLUPS", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    ref = numpy_reference_arm()
    line["numpy_reference_mlups"] = ref if ref is not None else None
    if ref is None:
        line["numpy_reference_note"] = ("the reference is pure Python and is not on this machine; measured in the build "
                                        "container: profiles/numpy_reference_r2.json (about 1 MLUPS on 1 core)")
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------------------
# product arm
# ------------------------------------------------------------------------------------------------------------
class Job:
    """one engine of a slab-decomposed FP run on the synthetic medium: geometry window, engine, runner"""

    def __init__(self, H, W, world, rank, local, dtype, kernel, halo, solid=None, refl=None):
        from fingering_dynamics_b200 import Engine, synthetic as syn, pinned_empty
        from fingering_dynamics_b200.slab import SlabRunner, slab_bounds
        self.H, self.W, self.world, self.rank, self.dtype = H, W, world, rank, dtype
        self.x0, self.x1 = slab_bounds(W, world, rank)
        self.lo, self.hi = max(0, self.x0 - 2), min(W, self.x1 + 2)
        self.c = c = syn.fp_constants(H)
        if solid is None:
            solid, refl = syn.porous_geometry(H, W, col0=self.lo, ncols=self.hi - self.lo)
        # the job's only host inputs, in page-locked memory: 2 bytes per cell
        self.solid, self.refl = pinned_empty(solid.shape, np.uint8), pinned_empty(refl.shape, np.uint8)
        self.solid[...] = solid
        self.refl[...] = refl
        self.kw = dict(tau=c["tau"], gamma=c["gamma"], a=c["a"], kappa=c["kappa"], Eta_n=c["Eta_n"], M=c["M"],
                       psi_wall=c["psi_wall"], zou_he="fp", inlet_ux=c["inlet_ux"], outlet_ux=c["outlet_ux"], dtype=dtype,
                       kernel=kernel, device=local, slab=(self.x0, self.x1), external_halo=world > 1)
        self.eng = Engine(H, W, **self.kw)
        self.eng.set_geometry(self.solid, self.refl, col0=self.lo)
        self.runner = SlabRunner(self.eng, rank, world, halo=halo)
        self.cells_local = H * (self.x1 - self.x0)

    def own_solid(self):
        return self.solid[:, self.x0 - self.lo:self.x0 - self.lo + (self.x1 - self.x0)]

    def init(self):
        self.runner.init_state(variant="fp", rho0=self.c["rho0"])   # Compute.__init__ on the device

    def close(self):
        from fingering_dynamics_b200 import pinned_free
        self.eng.close()
        pinned_free(self.solid)
        pinned_free(self.refl)


def timed_windows(job, K, warm, reps, align, barrier, allreduce_max):
    """-> (list of window durations in ms, max over ranks each; kernel launches per window; host times of the region)"""
    import torch
    eng, runner = job.eng, job.runner
    stream = torch.cuda.ExternalStream(eng.stream, device=torch.device("cuda", torch.cuda.current_device()))
    runner.step(warm)
    barrier()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(reps)]
    t_host0 = time.time()
    l_tot0 = eng.launch_count
    for a, b in ev:           # enqueued back to back: no host synchronisation until the last window is in
        runner.step(align)
        l0 = eng.launch_count
        a.record(stream)
        runner.step(K)
        b.record(stream)
        per_window = eng.launch_count - l0
    barrier()
    t_host1 = time.time()
    ms = allreduce_max([a.elapsed_time(b) for a, b in ev])
    assert eng.launch_count - l_tot0 == reps * (per_window + align)
    return ms, per_window, (t_host0, t_host1)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--reps", type=int, default=5, help="timed windows of --steps steps each; the median is reported")
    ap.add_argument("--align", type=int, default=3, help="untimed steps enqueued directly before every timed window")
    ap.add_argument("--dtype", default="f64", choices=["f64", "f32"])
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--kernel", default="auto", choices=["auto", "fused", "twopass"])
    ap.add_argument("--workload", default="c4", choices=["c4", "c5"],
                    help="c4: BASELINE configs[3], 8192x2048 per GPU (weak scaling); c5: configs[4], 32768x8192 in total, "
                         "slab-decomposed over the GPUs (strong scaling)")
    ap.add_argument("--H", type=int, default=0)
    ap.add_argument("--W", type=int, default=0, help="global columns (default 8192 per GPU)")
    ap.add_argument("--halo", default="peer", choices=["peer", "nccl"], help="slab halo transport for N>1")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the fp32 / sustained / multi_gpu_check / strong_c5 legs")
    ap.add_argument("--timeline", default="", help="write per-rank per-step event times of a K-step run to this JSON file")
    args = ap.parse_args()
    if args.impl == "reference":
        return reference_arm(args)

    import torch
    from fingering_dynamics_b200 import pinned_empty

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit("--gpus %d but WORLD_SIZE=%d" % (args.gpus, world))
    torch.cuda.set_device(local)
    clocks = ClockSampler(local)   # nvidia-smi takes a moment to start: launch it before the host-side set-up
    clocks.__enter__()
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    if args.workload == "c5":
        H, W, scaling = args.H or C5_H, args.W or C5_W, "strong"
    else:
        H, W, scaling = args.H or H_DEFAULT, args.W or W_PER_GPU * world, "weak"
    K, Wm, reps, align = args.steps, max(args.warmup, 3), max(1, args.reps), max(0, args.align)

    def barrier():
        torch.cuda.synchronize()      # my own engine stream first, then everybody, so that "past the barrier" means
        if world > 1:                 # no rank still pushes halo columns into a neighbour's lattice
            dist.barrier()
            torch.cuda.synchronize()

    def allreduce_max(vals):
        if world == 1:
            return [float(v) for v in vals]
        t = torch.tensor(list(vals), device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return [float(v) for v in t.tolist()]

    halo = args.halo
    if world > 1 and halo == "peer":
        # peer-mapped halos need CUDA IPC between the ranks; if any rank cannot attach, all use NCCL send/recv
        try:
            job = Job(H, W, world, rank, local, args.dtype, args.kernel, "peer")
            ok = 1
        except Exception as ex:  # noqa: BLE001
            print("rank %d: peer halo unavailable (%s), using NCCL" % (rank, ex), file=sys.stderr)
            ok, job = 0, None
        if int(allreduce_max([-ok])[0]) != -1:
            halo = "nccl"
            if job is not None:
                job.close()
            job = Job(H, W, world, rank, local, args.dtype, args.kernel, "nccl")
    else:
        job = Job(H, W, world, rank, local, args.dtype, args.kernel, halo)
    extras = not args.no_extras

    # ---- the N>1 path checked before it is timed -------------------------------------------------------
    check = multi_gpu_check(world, rank, local, halo, dist) if (world > 1 and extras) else None

    # ---- device-resident throughput ("value") -----------------------------------------------------------
    job.init()
    ms_w, launches, (th0, th1) = timed_windows(job, K, Wm, reps, align, barrier, allreduce_max)
    ms = float(np.median(ms_w))
    value = H * W * K / (ms * 1e-3) / 1e6
    sustained = None
    if extras:
        n_sus = int(min(20000, max(K, np.ceil(2000.0 / (ms / K)))))
        ms_s, _, (ts0, ts1) = timed_windows(job, n_sus, 0, 1, align, barrier, allreduce_max)
        sustained = {"value": H * W * n_sus / (ms_s[0] * 1e-3) / 1e6, "unit": "MLUPS", "steps": n_sus,
                     "ms_per_step": ms_s[0] / n_sus, "clocks": clocks.summary(ts0, ts1),
                     "note": "one window of >= 2 s: the board reaches its power cap (sw_power_cap is expected here)"}
        th1 = ts1
    timeline = step_timeline(job, K, align, barrier, dist, world, rank) if args.timeline else None

    # ---- end to end through the public API with host buffers ("e2e") ---------------------------------
    e2e = None
    if not args.no_e2e:
        out = {k: pinned_empty((H, job.x1 - job.x0)) for k in ("psi", "rho", "ux", "uy")}
        h2d = job.solid.nbytes + job.refl.nbytes
        d2h = sum(v.nbytes for v in out.values())
        jobs = []
        for _ in range(3):
            barrier()
            t0 = time.perf_counter()
            job.eng.set_geometry(job.solid, job.refl, col0=job.lo)   # H2D: the job's inputs (solid mask + reflect bits)
            job.init()                                               # Compute.__init__ on the device
            job.runner.step(K)
            job.runner.get_state(("psi", "rho", "ux", "uy"), out=out)   # D2H of the result fields
            barrier()
            jobs.append(time.perf_counter() - t0)
        dt = float(np.median(allreduce_max(jobs)))
        e2e = {"value": H * W * K / dt / 1e6, "unit": "MLUPS", "h2d_bytes_per_step": h2d * world / K,
               "d2h_bytes_per_step": d2h * world / K, "job_ms": dt * 1e3, "jobs_ms": [j * 1e3 for j in jobs],
               "note": "one job = set_geometry (H2D of the solid mask and reflect bits, 2 B/cell, page-locked) + the "
                       "device-side Compute.__init__ (fdlbm_init_state) + K steps + get_state(psi,rho,ux,uy) (D2H, "
                       "float64, page-locked); wall clock, max over ranks, median of 3 jobs; bytes are the job's transfers "
                       "divided by K (the state stays resident between steps, as in the reference's loop)"}
        # what the job costs when only psi is read back -- the one field the reference's main() keeps (its frames,
        # fingering_periodic.py:453,462 / fingering.py:557,565-566); reported next to the headline, not instead of it
        jobs_psi = []
        for _ in range(3):
            barrier()
            t0 = time.perf_counter()
            job.eng.set_geometry(job.solid, job.refl, col0=job.lo)
            job.init()
            job.runner.step(K)
            job.runner.get_state(("psi",), out={"psi": out["psi"]})
            barrier()
            jobs_psi.append(time.perf_counter() - t0)
        dtp = float(np.median(allreduce_max(jobs_psi)))
        e2e["psi_only"] = {"value": H * W * K / dtp / 1e6, "unit": "MLUPS", "job_ms": dtp * 1e3,
                           "d2h_bytes_per_step": out["psi"].nbytes * world / K,
                           "note": "the same job reading back psi alone (what the reference's main() keeps per frame); "
                                   "the read-back of the 4-field job above is 96 % PCIe time (4 x 2.37 ms of 9.9 ms, gpurun_in/e2e_breakdown.py)"}

    # ---- further legs --------------------------------------------------------------------------------------
    fp32 = None
    fluid_fraction = float((job.own_solid() == 0).mean())
    cells_local = job.cells_local
    if extras and world == 1 and args.dtype == "f64" and args.workload == "c4":
        j32 = Job(H, W, 1, 0, local, "f32", args.kernel, "none", solid=job.solid, refl=job.refl)
        j32.init()
        ms32, _, _ = timed_windows(j32, K, Wm, reps, align, barrier, allreduce_max)
        m32 = float(np.median(ms32))
        ach32 = B_ALG["f32"] * cells_local / (m32 / K * 1e-3) / 1e9
        fp32 = {"value": H * W * K / (m32 * 1e-3) / 1e6, "unit": "MLUPS", "ms_per_step": m32 / K, "windows_ms": ms32,
                "roofline": {"bound": "hbm", "achieved": ach32, "peak": peak_hbm()[0], "unit": "GB/s",
                             "frac": ach32 / peak_hbm()[0], "traffic": lookup_traffic("f32", job.x1 - job.x0, H)[0]},
                "kernel": "f32p::k_fused_f32p (packed two-row kernel, lbm_fused_f32.cuh)", "tolerance": FP32_TOLERANCE}
        j32.close()
    strong = None
    if extras and world > 1 and args.workload == "c4":
        job.close()
        job = None
        strong = strong_c5(world, rank, local, args, halo, dist, barrier, allreduce_max)

    if timeline is not None and rank == 0:
        os.makedirs(os.path.dirname(os.path.abspath(args.timeline)), exist_ok=True)
        json.dump(timeline, open(args.timeline, "w"))
    clocks.__exit__()
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peak, how = peak_hbm()
    here = copy_bandwidth_here() if extras else None
    step_ms = ms / K
    achieved = B_ALG[args.dtype] * cells_local / (step_ms * 1e-3) / 1e9
    traffic, traffic_note = lookup_traffic(args.dtype, cells_local // H, H) if world == 1 else \
        (None, "not captured for N > 1 (ncu is a one-GPU tool here; the peer halo adds 2 x 2 columns of stores per step)")
    line = {"metric": "MLUPS (two-phase D2Q9)", "value": value, "unit": "MLUPS", "n_gpus": world, "steps": K,
            "warmup": Wm, "ms_per_step": step_ms, "higher_is_better": True, "scaling": scaling, "vs_baseline": None,
            "dtype": args.dtype, "data": "synthetic",
            "timing": {"reps": reps, "align_steps": align, "windows_ms": ms_w, "statistic": "median window, max over ranks",
                       "note": "windows enqueued back to back without host synchronisation; each = align untimed steps + "
                               "exactly K timed steps between two CUDA events on the engine stream",
                       "value_burst": value, "value_sustained": None if sustained is None else sustained["value"]},
            "config": {"workload": "synthetic random-block porous medium %dx%d (W x H), fingering_periodic step variant"
                                   % (W, H), "grid_W": W, "grid_H": H, "slab_columns_per_gpu": cells_local // H,
                       "obstacles": "circles r 8-12 on a jittered 40-pitch lattice, seed 1234",
                       "initial_state": "Compute.__init__ on the device (fdlbm_init_state)",
                       "kernel": args.kernel, "l2": "state (%.1f GB per lattice copy) far exceeds the 126 MB L2; no flush needed"
                                                    % (cells_local * 18 * (8 if args.dtype == "f64" else 4) / 1e9),
                       "parallelism": "slab%d" % world,
                       "halo": ("in-kernel peer stores over NVLink + stream flags" if halo == "peer" else "NCCL send/recv")
                       if world > 1 else "none"},
            "clocks": clocks.summary(th0, th1), "gpu_launches": launches, "e2e": e2e,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "traffic_note": traffic_note, "peak_source": how,
                         "copy_gbs_this_box": here, "frac_of_this_box": None if not here else achieved / here,
                         "frac_sustained": None if sustained is None else
                         B_ALG[args.dtype] * cells_local / (sustained["ms_per_step"] * 1e-3) / 1e9 / peak,
                         "note": "achieved = %.0f B/LU x %d LU per launch / mean launch duration (CUDA events over the median "
                                 "timed window); burst figure (windows of K steps), frac_sustained from the >= 2 s window; "
                                 "peak is the pod's recorded copy bandwidth, copy_gbs_this_box the same copy measured on this "
                                 "box after the timed region (boxes differ by a few per cent, so frac can touch 1)"
                                 % (B_ALG[args.dtype], cells_local)},
            "fluid_fraction": fluid_fraction}
    if sustained is not None:
        line["sustained"] = sustained
    if fp32 is not None:
        line["fp32"] = fp32
    if check is not None:
        line["multi_gpu_check"] = check
    if strong is not None:
        line["strong_c5"] = strong
    if not args.no_cpu and world == 1:
        cores = os.cpu_count() or 1
        Hs, Ws = H, 1024
        v, dt = cpu_arm(Hs, Ws, 30, 2, cores)
        line["cpu_baseline"] = {"value": v, "unit": "MLUPS", "cores": cores, "kind": "port",
                                "sample": "30 steps of a %dx%d crop of the workload, oracle/fd_oracle.c (OpenMP); the NumPy "
                                          "reference itself: numpy_reference_mlups" % (Ws, Hs)}
        ref = numpy_reference_arm()
        line["numpy_reference_mlups"] = ref
        if ref is None:
            line["numpy_reference_note"] = ("the reference is pure Python and is not on this machine; measured in the build "
                                            "container: profiles/numpy_reference_r2.json (about 1 MLUPS on 1 core)")
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def lookup_traffic(dtype, wl, H):
    """dram bytes per launch of the step kernel from the committed ncu capture of THIS dtype and slab, else None"""
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    key = "k_fused_%s_bytes_per_launch_%dx%d" % (dtype, wl, H)
    try:
        t = json.load(open(tp))
        if key in t:
            return t[key], "ncu --set full capture of this kernel on this grid (%s); not re-measured by this run" % t.get("source", tp)
    except Exception:
        pass
    return None, "no ncu capture for %s on a %dx%d slab" % (dtype, wl, H)


def multi_gpu_check(world, rank, local, halo, dist):
    """256 rows x 96 columns per rank over the REAL ranks (same transport as the timed run): step(1); step(4); step(7),
    then every rank's slab is gathered and compared on rank 0 with ONE engine on the whole grid, bit for bit."""
    from fingering_dynamics_b200 import Engine, synthetic as syn
    H, W = 256, 96 * world
    names = ("f", "g", "psi", "rho", "ux", "uy")
    res = {"grid": "%dx%d" % (W, H), "steps": 12, "halo": halo, "ranks": world, "dtypes": {}}
    for dtype in ("f64", "f32"):
        j = Job(H, W, world, rank, local, dtype, "auto", halo)
        j.init()
        for n in (1, 4, 7):
            j.runner.step(n)
        got = j.runner.get_state(names)
        allg = [None] * world
        dist.all_gather_object(allg, {k: got[k] for k in names})
        bounds = (j.x0, j.x1)
        allb = [None] * world
        dist.all_gather_object(allb, bounds)
        j.close()
        if rank == 0:
            c = syn.fp_constants(H)
            solid, refl = syn.porous_geometry(H, W)
            e = Engine(H, W, tau=c["tau"], gamma=c["gamma"], a=c["a"], kappa=c["kappa"], Eta_n=c["Eta_n"], M=c["M"],
                       psi_wall=c["psi_wall"], zou_he="fp", inlet_ux=c["inlet_ux"], outlet_ux=c["outlet_ux"], dtype=dtype,
                       device=local)
            e.set_geometry(solid, refl)
            e.init_state("fp", rho0=c["rho0"])
            e.step(12)
            want = e.get_state(names)
            e.close()
            same = all(np.array_equal(allg[r][k], want[k][..., allb[r][0]:allb[r][1]]) for r in range(world) for k in names)
            res["dtypes"][dtype] = bool(same)
    res["bit_identical"] = bool(res["dtypes"]) and all(res["dtypes"].values())
    ok = [res["bit_identical"]] if rank == 0 else [None]
    dist.broadcast_object_list(ok, src=0)
    if not ok[0]:
        raise SystemExit("multi_gpu_check FAILED: the %d-rank run differs from one engine (%s)" % (world, res["dtypes"]))
    return res


def strong_c5(world, rank, local, args, halo, dist, barrier, allreduce_max):
    """BASELINE configs[4]: 32768x8192 split over the N GPUs, then the SAME grid on one GPU (rank 0) for the
    efficiency denominator -- both measured in this run."""
    import torch
    H, W = C5_H, C5_W
    K = max(5, min(args.steps, 40))
    reps, align = max(1, min(args.reps, 3)), args.align
    job = Job(H, W, world, rank, local, args.dtype, args.kernel, halo)
    job.init()
    ms_w, _, _ = timed_windows(job, K, 3, reps, align, barrier, allreduce_max)
    ms = float(np.median(ms_w))
    # gather the geometry for the 1-GPU run (uint8, 2 x 268 MB in total)
    own = slice(job.x0 - job.lo, job.x0 - job.lo + (job.x1 - job.x0))
    parts = []
    for plane in (job.solid, job.refl):
        mine = torch.from_numpy(np.ascontiguousarray(plane[:, own].T)).cuda()     # (columns, H): equal sizes per rank
        buf = [torch.empty_like(mine) for _ in range(world)] if rank == 0 else None
        dist.gather(mine, buf, dst=0)
        parts.append(None if rank != 0 else np.ascontiguousarray(torch.cat(buf, 0).T.cpu().numpy()))
        del mine, buf
    job.close()
    torch.cuda.empty_cache()
    one = None
    if rank == 0:
        j1 = Job(H, W, 1, 0, local, args.dtype, args.kernel, "none", solid=parts[0], refl=parts[1])
        j1.init()
        K1 = max(5, K // 2)
        m1, _, _ = timed_windows(j1, K1, 2, 2, 1, lambda: torch.cuda.synchronize(), lambda v: [float(x) for x in v])
        one = {"ms_per_step": float(np.median(m1)) / K1, "value": H * W * K1 / (float(np.median(m1)) * 1e-3) / 1e6,
               "steps": K1, "windows_ms": m1}
        j1.close()
    barrier()
    if rank != 0:
        return None
    v = H * W * K / (ms * 1e-3) / 1e6
    return {"workload": "synthetic porous fingering %dx%d (W x H), BASELINE configs[4], slab-decomposed" % (W, H),
            "scaling": "strong", "n_gpus": world, "slab_columns_per_gpu": W // world, "steps": K, "reps": reps,
            "windows_ms": ms_w, "ms_per_step": ms / K, "value": v, "unit": "MLUPS", "one_gpu": one,
            "speedup": v / one["value"], "efficiency": v / one["value"] / world,
            "roofline_frac": B_ALG[args.dtype] * H * (W // world) / (ms / K * 1e-3) / 1e9 / peak_hbm()[0]}


def step_timeline(job, K, align, barrier, dist, world, rank):
    """per-rank CUDA-event time of every step of one K-step run (each step enqueued by its own call, so the events sit
    between the steps) plus host clock stamps (CLOCK_REALTIME is shared by the ranks of one box)"""
    import torch
    eng, runner = job.eng, job.runner
    stream = torch.cuda.ExternalStream(eng.stream, device=torch.device("cuda", torch.cuda.current_device()))
    out = {}
    for tag, pre in (("after_barrier", 0), ("after_align", align)):
        barrier()
        t_enq0 = time.time()
        runner.step(pre)
        evs = [torch.cuda.Event(enable_timing=True) for _ in range(K + 1)]
        evs[0].record(stream)
        for k in range(K):
            runner.step(1)
            evs[k + 1].record(stream)
        t_enq1 = time.time()
        torch.cuda.synchronize()
        t_done = time.time()
        rec = {"rank": rank, "t_enqueue_start": t_enq0, "t_enqueue_end": t_enq1, "t_done": t_done,
               "step_ms": [evs[k].elapsed_time(evs[k + 1]) for k in range(K)], "total_ms": evs[0].elapsed_time(evs[K])}
        if world > 1:
            allr = [None] * world
            dist.all_gather_object(allr, rec)
        else:
            allr = [rec]
        out[tag] = allr
    return out


if __name__ == "__main__":
    main()
