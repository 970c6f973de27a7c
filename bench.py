#!/usr/bin/env python
"""bench.py -- MLUPS of the two-phase D2Q9 step on synthetic porous domains (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--dtype f64|f32] [--impl ours|reference]

N=1 workload: BASELINE.json configs[3], "synthetic random-block porous medium 8192x2048" (W=8192 columns
along the flow, H=2048 rows), fingering_periodic.py's step variant, fp64.  N>1 (torchrun): the same slab per
GPU concatenated along the flow axis (weak scaling, W = 8192*N), 2-column halos exchanged by NCCL.
One "step" is one lattice-Boltzmann iteration over the whole grid; value = H*W*K / t / 1e6 (all cells:
solids are streamed too), t from CUDA events on the engine stream, max over ranks.

--impl reference times the CPU arm: the oracle port (oracle/fd_oracle.c, OpenMP, all host threads) on a
bounded crop of the same workload.  The reference itself is NumPy and does not travel to the GPU box; its
own speed measured in the build container is ~1.0 MLUPS on one core.
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
B_ALG = {"f64": 288.0, "f32": 144.0}  # algorithmic bytes per lattice update: 18 populations read + written once


def peak_hbm():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi SM clocks / throttle reasons sampled during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.rows, self.proc, self.index = [], None, index
        self.t_mark = None
        self.t_end = None

    def mark(self):
        """start of the timed region: only samples taken after this call are summarised"""
        self.t_mark = time.time()

    def end(self):
        self.t_end = time.time()

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "50"], stdout=subprocess.PIPE,
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
            time.sleep(0.15)
            self.proc.terminate()
            self.thr.join(timeout=2)

    def summary(self):
        sm, mx, reasons = [], [], set()
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        t1 = (self.t_end + 0.1) if self.t_end else float("inf")
        rows = [r for (ts, r) in self.rows if (self.t_mark is None or ts >= self.t_mark) and ts <= t1]
        if len(rows) < 3:  # short timed region: the samples closest to it (warm-up is load too)
            rows = [r for (ts, r) in self.rows if ts <= t1][-5:]
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
                "samples": len(sm)}


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


def reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    Hs, Ws = H_DEFAULT, 1024  # bounded crop of the 2048x8192 workload (same generator, same H)
    steps = max(1, min(args.steps, 40))
    warm = max(1, min(args.warmup, 3))
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
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=400)
    ap.add_argument("--warmup", type=int, default=5)
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
    ap.add_argument("--pageable", action="store_true", help="host arrays in pageable memory (very large grids)")
    args = ap.parse_args()
    if args.impl == "reference":
        return reference_arm(args)

    import torch
    from fingering_dynamics_b200 import Engine, synthetic as syn, pinned_empty
    from fingering_dynamics_b200.slab import SlabRunner, slab_bounds

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit("--gpus %d but WORLD_SIZE=%d" % (args.gpus, world))
    torch.cuda.set_device(local)
    clocks = ClockSampler(local)   # nvidia-smi takes a moment to start: launch it before the host-side set-up
    clocks.__enter__()
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    if args.workload == "c5":
        H, W, scaling = args.H or 8192, args.W or 32768, "strong"
        args.pageable = True   # 38.7 GB of populations per lattice copy: too much to page-lock
    else:
        H, W, scaling = args.H or H_DEFAULT, args.W or W_PER_GPU * world, "weak"
    x0, x1 = slab_bounds(W, world, rank)
    K, Wm = args.steps, max(args.warmup, 3)

    c = syn.fp_constants(H)
    solid, refl = syn.porous_geometry(H, W, col0=max(0, x0 - 2), ncols=min(W, x1 + 2) - max(0, x0 - 2))
    own = slice(x0 - max(0, x0 - 2), x0 - max(0, x0 - 2) + (x1 - x0))
    alloc = np.zeros if args.pageable else pinned_empty
    st = syn.fp_initial_state(np.ascontiguousarray(solid[:, own]), c, col0=x0, alloc=alloc)
    for k in list(st):  # every array of the e2e job lives in page-locked host memory
        if k not in ("f", "g") and not args.pageable:
            pin = pinned_empty(st[k].shape)
            pin[...] = st[k]
            st[k] = pin

    eng = Engine(H, W, tau=c["tau"], gamma=c["gamma"], a=c["a"], kappa=c["kappa"], Eta_n=c["Eta_n"], M=c["M"],
                 psi_wall=c["psi_wall"], zou_he="fp", inlet_ux=c["inlet_ux"], outlet_ux=c["outlet_ux"], dtype=args.dtype,
                 kernel=args.kernel, device=local, slab=(x0, x1), external_halo=world > 1)
    eng.set_geometry(solid, refl, col0=max(0, x0 - 2))
    halo = args.halo
    if world > 1 and halo == "peer":
        # peer-mapped halos need CUDA IPC between the ranks; if any rank cannot attach, all use NCCL send/recv
        try:
            runner = SlabRunner(eng, rank, world, halo="peer")
            ok = 1
        except Exception as ex:  # noqa: BLE001
            print("rank %d: peer halo unavailable (%s), using NCCL" % (rank, ex), file=sys.stderr)
            ok = 0
        t_ok = torch.tensor([ok], device="cuda")
        dist.all_reduce(t_ok, op=dist.ReduceOp.MIN)
        if int(t_ok.item()) == 0:
            halo = "nccl"
            eng.close()
            eng = Engine(H, W, tau=c["tau"], gamma=c["gamma"], a=c["a"], kappa=c["kappa"], Eta_n=c["Eta_n"], M=c["M"],
                         psi_wall=c["psi_wall"], zou_he="fp", inlet_ux=c["inlet_ux"], outlet_ux=c["outlet_ux"],
                         dtype=args.dtype, kernel=args.kernel, device=local, slab=(x0, x1), external_halo=True)
            eng.set_geometry(solid, refl, col0=max(0, x0 - 2))
            runner = SlabRunner(eng, rank, world, halo="nccl")
    else:
        runner = SlabRunner(eng, rank, world, halo=halo)
    stream = torch.cuda.ExternalStream(eng.stream, device=torch.device("cuda", local))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        eng.sync()

    # ---- device-resident throughput ("value") -------------------------------------------------------
    runner.set_state(col0=x0, **st)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    runner.step(Wm)
    barrier()
    l0 = eng.launch_count
    clocks.mark()
    ev0.record(stream)
    runner.step(K)
    ev1.record(stream)
    barrier()
    clocks.end()
    clocks.__exit__()
    ms = ev0.elapsed_time(ev1)
    launches = eng.launch_count - l0
    if world > 1:
        tt = torch.tensor([ms], device="cuda")
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ms = float(tt.item())
    cells_local = H * (x1 - x0)
    value = H * W * K / (ms * 1e-3) / 1e6

    # ---- end to end through the public API with host buffers ("e2e") ---------------------------------
    e2e = None
    if not args.no_e2e:
        out = {k: pinned_empty((H, x1 - x0)) for k in ("psi", "rho", "ux", "uy")}
        h2d = sum(st[k].nbytes for k in st)
        d2h = sum(v.nbytes for v in out.values())
        barrier()
        t0 = time.perf_counter()
        runner.set_state(col0=x0, **st)     # H2D of f, g and the macroscopic arrays (pinned host memory)
        runner.step(K)
        runner.get_state(("psi", "rho", "ux", "uy"), out=out)   # D2H of the result fields
        barrier()
        dt = time.perf_counter() - t0
        if world > 1:
            tt = torch.tensor([dt], device="cuda")
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            dt = float(tt.item())
        e2e = {"value": H * W * K / dt / 1e6, "unit": "MLUPS", "h2d_bytes_per_step": h2d * world / K,
               "d2h_bytes_per_step": d2h * world / K,
               "note": "one job = set_state (H2D) + K steps + get_state(psi,rho,ux,uy) (D2H), wall clock; bytes are the "
                       "job's transfers divided by K (the state stays resident between steps, as in the reference's loop)"}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peak, how = peak_hbm()
    step_ms = ms / K
    achieved = B_ALG[args.dtype] * cells_local / (step_ms * 1e-3) / 1e9
    traffic = None
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tp):
        try:
            traffic = json.load(open(tp)).get("k_fused_%s_bytes_per_launch_%dx%d" % (args.dtype, x1 - x0, H))
        except Exception:
            traffic = None
    line = {"metric": "MLUPS (two-phase D2Q9)", "value": value, "unit": "MLUPS", "n_gpus": world, "steps": K,
            "warmup": Wm, "ms_per_step": step_ms, "higher_is_better": True, "scaling": scaling, "vs_baseline": None,
            "dtype": args.dtype, "data": "synthetic",
            "config": {"workload": "synthetic random-block porous medium %dx%d (W x H), fingering_periodic step variant"
                                   % (W, H), "grid_W": W, "grid_H": H, "slab_columns_per_gpu": x1 - x0,
                       "obstacles": "circles r 8-12 on a jittered 40-pitch lattice, seed 1234",
                       "kernel": args.kernel, "l2": "state (%.1f GB per lattice copy) far exceeds the 126 MB L2; no flush needed"
                                                    % (cells_local * 18 * (8 if args.dtype == "f64" else 4) / 1e9),
                       "parallelism": "slab%d" % world,
                       "halo": ("in-kernel peer stores over NVLink + stream flags" if halo == "peer" else "NCCL send/recv")
                       if world > 1 else "none"},
            "clocks": clocks.summary(), "gpu_launches": launches, "e2e": e2e,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "peak_source": how,
                         "note": "achieved = %.0f B/LU x %d LU per launch / mean launch duration (CUDA events over the timed region)"
                                 % (B_ALG[args.dtype], cells_local)},
            "fluid_fraction": float((solid[:, own] == 0).mean())}
    if not args.no_cpu and world == 1:
        cores = os.cpu_count() or 1
        Hs, Ws = H, 1024
        v, dt = cpu_arm(Hs, Ws, 30, 2, cores)
        line["cpu_baseline"] = {"value": v, "unit": "MLUPS", "cores": cores, "kind": "port",
                                "sample": "30 steps of a %dx%d crop of the workload, oracle/fd_oracle.c (OpenMP); the NumPy "
                                          "reference itself runs ~1.0 MLUPS on 1 core (BASELINE.md)" % (Ws, Hs)}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
