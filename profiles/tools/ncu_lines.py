#!/usr/bin/env python
"""Per source line: warp instructions executed (per warp and column when --per N is given) and stall samples, from
`ncu -i <report> --page source --csv --print-source cuda,sass`:  python profiles/tools/ncu_lines.py <report> [--per N] [--top K]"""
import csv, io, subprocess, sys


def main():
    rep = sys.argv[1]
    per = float(sys.argv[sys.argv.index("--per") + 1]) if "--per" in sys.argv else 1.0
    top = int(sys.argv[sys.argv.index("--top") + 1]) if "--top" in sys.argv else 40
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    fname, lines, ix = "", [], {}
    for r in rows:
        if not r:
            continue
        if r[0] == "File Path":
            fname = r[1].split("/")[-1]
        elif r[0] == "Line No":
            ix = {h: i for i, h in enumerate(r)}
        elif r[0] not in ("", "Function Name") and ix and len(r) > ix["Instructions Executed"]:
            try:
                lines.append((fname, int(r[0]), r[1].strip(), int(r[ix["# Samples"]] or 0), int(r[ix["Instructions Executed"]] or 0)))
            except ValueError:
                pass
    tot_s = sum(l[3] for l in lines) or 1
    tot_i = sum(l[4] for l in lines) or 1
    print("total: %d samples, %.1f instructions per unit" % (tot_s, tot_i / per))
    print("%-22s %6s %9s %7s  %s" % ("file:line", "inst", "inst %", "stall %", "source"))
    for f, n, src, smp, ins in sorted(lines, key=lambda l: -l[4])[:top]:
        print("%-22s %6.1f %8.1f%% %6.1f%%  %s" % ("%s:%d" % (f, n), ins / per, 100.0 * ins / tot_i, 100.0 * smp / tot_s, src[:100]))
    print("-- by stall samples")
    for f, n, src, smp, ins in sorted(lines, key=lambda l: -l[3])[:top // 2]:
        print("%-22s %6.1f %8.1f%% %6.1f%%  %s" % ("%s:%d" % (f, n), ins / per, 100.0 * ins / tot_i, 100.0 * smp / tot_s, src[:100]))


if __name__ == "__main__":
    main()
