"""CTA placement of the fused fp64 step on the benchmark grid (8192x2048): step time before / after the slow SMs are
marked, which SMs they are, per-SM times of the last measuring launch -> gpurun_out/placement_dump_<tag>.json"""
import json, os, sys
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
out = {"grid": [W, H], "placement_env": os.environ.get("FDLBM_PLACEMENT", "1"), "per_launch_ms": []}
for k in range(12):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(stream); e.step(1); b.record(stream); torch.cuda.synchronize()
    out["per_launch_ms"].append(a.elapsed_time(b))
pi = e.placement_info()
out["items"], out["marked_sms"] = pi["items"], pi["marked_sms"]
if pi["sm_ids"] is not None and pi["items"] > 0:
    t, sm = pi["ticks_ns"].astype(float) / 1e3, pi["sm_ids"]
    per = {int(k): [int((sm == k).sum()), float(t[sm == k].max())] for k in np.unique(sm)}
    out["per_sm_ctas_and_slowest_cta_us"] = per
    full = [v[1] for v in per.values() if v[0] == 3]
    out["sm_time_us"] = {"min": min(full), "median": float(np.median(full)), "max": max(full)}
win = []
for _ in range(8):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(stream); e.step(20); b.record(stream); torch.cuda.synchronize()
    win.append(a.elapsed_time(b) / 20)
out["ms_per_step_windows_of_20"] = win
e.close()
os.makedirs("gpurun_out", exist_ok=True)
tag = os.environ.get("BD_TAG", "place" + out["placement_env"])
json.dump(out, open("gpurun_out/placement_dump_%s.json" % tag, "w"), indent=1)
print(tag, "launches:", [round(x, 4) for x in out["per_launch_ms"]])
print(tag, "windows:", [round(x, 4) for x in win], "marked", out["marked_sms"], out.get("sm_time_us"))
