"""Per-CTA timing of the fused step kernels (profiles/README.md).  Build the instrumented library first (here, no GPU needed):
    nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 --shared -Xcompiler -fPIC -DFDLBM_CTA_TIMES \
         -o gpurun_in/variants/lib_dbg.so fingering_dynamics_b200/csrc/fdlbm.cu
then on the GPU box:  python gpurun_in/cta_times.py f64|f32"""
import ctypes, os, sys, json, numpy as np
sys.path.insert(0, os.getcwd())
os.environ["FDLBM_LIB"] = os.path.join(os.getcwd(), "gpurun_in/variants/lib_dbg.so")
from fingering_dynamics_b200 import Engine, synthetic as syn, _native as nat
dtype = sys.argv[1]
H, W = 2048, 8192
c = syn.fp_constants(H)
solid, refl = syn.porous_geometry(H, W)
st = syn.fp_initial_state(solid, c)
e = Engine(H, W, tau=c["tau"], gamma=c["gamma"], a=c["a"], kappa=c["kappa"], Eta_n=c["Eta_n"], M=c["M"], psi_wall=c["psi_wall"],
           zou_he="fp", inlet_ux=c["inlet_ux"], outlet_ux=c["outlet_ux"], dtype=dtype)
e.set_geometry(solid, refl)
e.set_state(**st)
e.step(20)
e.sync()
buf = np.zeros(4096 * 4, np.uint64)
lib = nat.lib()
lib.fdlbm_debug_cta_times.argtypes = [ctypes.c_void_p, ctypes.c_size_t]
rc = lib.fdlbm_debug_cta_times(buf.ctypes.data, buf.nbytes)
t = buf.reshape(-1, 4)
idx = np.where(t[:, 1] > 0)[0]
t = t[idx]
t0 = t[:, 0].min()
start = (t[:, 0] - t0) / 1e3; end = (t[:, 1] - t0) / 1e3; dur = end - start
print(dtype, "rc", rc, "CTAs", len(t), "kernel span us", end.max())
strip = t[:, 3].astype(int)
for s_ in sorted(set(strip)):
    m = strip == s_
    print(" strip %4d n=%3d start %.1f..%.1f  dur mean %.1f min %.1f max %.1f  end max %.1f" % (s_, m.sum(), start[m].min(), start[m].max(), dur[m].mean(), dur[m].min(), dur[m].max(), end[m].max()))
sm = t[:, 2].astype(int)
cnt = np.bincount(sm)
for k in sorted(set(cnt[cnt > 0])):
    ids = np.where(cnt == k)[0]
    m = np.isin(sm, ids) & (strip < 1000)
    print(" SMs with %d CTAs: %d SMs, marching CTA dur mean %.1f max %.1f" % (k, len(ids), dur[m].mean() if m.any() else 0, dur[m].max() if m.any() else 0))
nyt = 8 if dtype == "f32" else 16
chunkid = idx // nyt
march = strip < 1000
print(" by chunk (x position): dur mean / max of chunk k")
for k in sorted(set(chunkid[march])):
    m = march & (chunkid == k)
    print("  %2d:%.0f/%.0f" % (k, dur[m].mean(), dur[m].max()), end="")
print()
per_sm = sorted((dur[march & (sm == s_)].mean(), s_) for s_ in set(sm[march]))
print(" fastest SMs:", [(round(a), b) for a, b in per_sm[:10]])
print(" slowest SMs:", [(round(a), b) for a, b in per_sm[-10:]])
sm_end = np.array([end[march & (sm == s_)].max() for s_ in range(148) if (march & (sm == s_)).any()])
print(" per-SM finishing time: min %.0f  p25 %.0f  median %.0f  p75 %.0f  max %.0f  mean %.0f" % (sm_end.min(), np.percentile(sm_end, 25), np.median(sm_end), np.percentile(sm_end, 75), sm_end.max(), sm_end.mean()))
print(" finishing time by SM id (groups of 8):", [int(sm_end[i:i + 8].mean()) for i in range(0, len(sm_end), 8)])
print(" sum of CTA durations / (n_sm * occ): %.0f" % (dur[march].sum() / 444))
e.close()
