"""CTA durations and column boundaries of the fused fp64 step's chunk balancer on the benchmark grid (8192x2048):
after 1, 2, 4, 8, 12, 40, 200 launches -> gpurun_out/balance_dump.json; plus the step time of windows of 20 steps."""
import json, os, sys, time
import numpy as np
sys.path.insert(0, os.getcwd())
import torch
from fingering_dynamics_b200 import Engine, synthetic as syn

H, W = int(os.environ.get("BD_H", 2048)), int(os.environ.get("BD_W", 8192))
c = syn.fp_constants(H)
solid, refl = syn.porous_geometry(H, W)
e = Engine(H, W, tau=c["tau"], gamma=c["gamma"], a=c["a"], kappa=c["kappa"], Eta_n=c["Eta_n"], M=c["M"], psi_wall=c["psi_wall"],
           zou_he="fp", inlet_ux=c["inlet_ux"], outlet_ux=c["outlet_ux"], dtype="f64")
e.set_geometry(solid, refl)
e.init_state("fp", rho0=c["rho0"])
stream = torch.cuda.ExternalStream(e.stream)
out = {"grid": [W, H], "balance_env": os.environ.get("FDLBM_BALANCE", "1"), "snap": []}
done = 0
for upto in (1, 2, 3, 4, 6, 8, 12, 13, 40, 200):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(stream); e.step(upto - done); b.record(stream); torch.cuda.synchronize()
    ms = a.elapsed_time(b) / (upto - done)
    done = upto
    bi = e.balance_info()
    s = {"launches": done, "ms_per_step_since_last": ms, "nyt": bi["nyt"], "nchunks": bi["nchunks"]}
    if bi["bounds"] is not None:
        t = bi["ticks_ns"].astype(float)
        s["bounds_strip0"] = bi["bounds"][0].tolist()
        s["chunk_len_min_max"] = [int(np.diff(bi["bounds"], axis=1).min()), int(np.diff(bi["bounds"], axis=1).max())]
        s["ticks_us"] = {"min": t.min() / 1e3, "mean": t.mean() / 1e3, "max": t.max() / 1e3, "p5": float(np.percentile(t, 5)) / 1e3,
                         "p95": float(np.percentile(t, 95)) / 1e3}
        s["ticks_by_chunk_us_mean"] = (t.mean(axis=1) / 1e3).round(1).tolist()
        sm = bi["sm_ids"]
        lens = np.diff(bi["bounds"], axis=1).T          # (nchunks, nyt) like ticks / sm_ids
        per = {}
        for k in np.unique(sm):
            sel = sm == k
            per[int(k)] = [int(sel.sum()), int(lens[sel].sum()), float(t[sel].max() / 1e3)]
        full = [v for v in per.values() if v[0] == 3]
        s["sms"] = len(per)
        s["sm_ctas_hist"] = {str(n): sum(1 for v in per.values() if v[0] == n) for n in (1, 2, 3, 4)}
        s["sm_slowest_cta_us"] = {"min": min(v[2] for v in per.values()), "median": float(np.median([v[2] for v in per.values()])),
                                  "max": max(v[2] for v in per.values())}
        if done in (2, 13, 200):
            s["per_sm"] = per
    out["snap"].append(s)
win = []
for _ in range(8):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(stream); e.step(20); b.record(stream); torch.cuda.synchronize()
    win.append(a.elapsed_time(b) / 20)
out["ms_per_step_windows_of_20"] = win
e.close()
os.makedirs("gpurun_out", exist_ok=True)
tag = os.environ.get("BD_TAG", "bal" + out["balance_env"])
json.dump(out, open("gpurun_out/balance_dump_%s.json" % tag, "w"), indent=1)
print(tag, "windows:", [round(x, 4) for x in win])
for s in out["snap"]:
    print(s["launches"], round(s["ms_per_step_since_last"], 4), s.get("chunk_len_min_max"), s.get("ticks_us"))
