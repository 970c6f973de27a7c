#!/usr/bin/env python
"""Key metrics of one kernel launch from an ncu report:  python profiles/tools/ncu_summary.py <report.ncu-rep> [--top N]
(reads `ncu -i <report> --page raw --csv` and, with --top, the source page: the N instructions with most stall samples)"""
import csv, io, subprocess, sys

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "sm__cycles_elapsed.max"]
STALLS = ["long_scoreboard", "wait", "barrier", "short_scoreboard", "math_pipe_throttle", "mio_throttle", "lg_throttle",
          "not_selected", "dispatch_stall", "branch_resolving", "no_instruction", "membar", "sleeping", "imc_miss"]


def page(rep, name):
    out = subprocess.run(["ncu", "-i", rep, "--page", name, "--csv"], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def main():
    rep = sys.argv[1]
    top = int(sys.argv[sys.argv.index("--top") + 1]) if "--top" in sys.argv else 0
    rows = page(rep, "raw")
    hdr, units, vals = rows[0], rows[1], rows[2]
    ix = {h: i for i, h in enumerate(hdr)}
    print("kernel:", vals[ix["Kernel Name"]])
    for w in WANT:
        if w in ix:
            print("  %-70s %s %s" % (w, vals[ix[w]], units[ix[w]]))
    print("  warps stalled per issued instruction (smsp__average_warps_issue_stalled_*_per_issue_active):")
    for s in STALLS:
        k = "smsp__average_warps_issue_stalled_%s_per_issue_active.ratio" % s
        if k in ix and float(vals[ix[k]] or 0) >= 0.005:
            print("    %-24s %.3f" % (s, float(vals[ix[k]])))
    rd, wr = float(vals[ix["dram__bytes_read.sum"]]), float(vals[ix["dram__bytes_write.sum"]])
    scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}
    tot = rd * scale[units[ix["dram__bytes_read.sum"]]] + wr * scale[units[ix["dram__bytes_write.sum"]]]
    print("  dram bytes per launch: %.0f" % tot)
    if top:
        rows = page(rep, "source")
        hdr = rows[1]
        ix = {h: i for i, h in enumerate(hdr)}
        data = [r for r in rows[2:] if len(r) == len(hdr)]
        tot_s = sum(float(r[ix["# Samples"]] or 0) for r in data) or 1.0
        print("  top %d instructions by stall samples (of %d):" % (top, tot_s))
        reasons = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
        for r in sorted(data, key=lambda r: -float(r[ix["# Samples"]] or 0))[:top]:
            n = float(r[ix["# Samples"]] or 0)
            why = sorted(((float(r[ix[h]] or 0), h[6:]) for h in reasons), reverse=True)[:2]
            print("    %5.2f %%  %-58s %s" % (100 * n / tot_s, r[ix["Source"]][:58], ", ".join("%s %d" % (h, v) for v, h in why if v)))


if __name__ == "__main__":
    main()
