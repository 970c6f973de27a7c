#!/usr/bin/env python
"""Scoreboard view of a kernel's SASS: python profiles/tools/sass_scoreboards.py <lib.so> <kernel name substring> [opcode]
Decodes the control field of every instruction (stall count, write / read barrier, wait mask: bits 105..125 of the
128-bit encoding) and prints each instruction matching `opcode` (default SHFL) that waits on a scoreboard some LDG of the
same kernel writes -- the pattern that cost the fp64 step a memory latency per column (profiles/README.md, round 2)."""
import re, subprocess, sys


def decode(lib, kern):
    out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout.split("\n")
    ins, cur, i = [], False, 0
    while i < len(out):
        l = out[i]
        if "Function :" in l:
            cur = kern in l
        m = re.match(r"\s+/\*([0-9a-f]{4,6})\*/\s+(.*?);\s+/\* 0x([0-9a-f]{16}) \*/", l) if cur else None
        if m and i + 1 < len(out):
            m2 = re.match(r"\s+/\* 0x([0-9a-f]{16}) \*/", out[i + 1])
            if m2:
                c = (((int(m2.group(1), 16) << 64) | int(m.group(3), 16)) >> 105) & ((1 << 21) - 1)
                ins.append(dict(addr=int(m.group(1), 16), text=m.group(2), stall=c & 15, wr=(c >> 5) & 7, rd=(c >> 8) & 7,
                                wait=(c >> 11) & 63))
                i += 2
                continue
        i += 1
    return ins


def main():
    lib, kern = sys.argv[1], sys.argv[2]
    op = sys.argv[3] if len(sys.argv) > 3 else "SHFL"
    ins = decode(lib, kern)
    ldg_sb = sorted({x["wr"] for x in ins if "LDG" in x["text"] and "LDGSTS" not in x["text"] and x["wr"] != 7})
    print("%d instructions; scoreboards written by LDG: %s" % (len(ins), ldg_sb))
    # loops = backward branches; an instruction is reported with its innermost enclosing loop
    loops = []
    for x in ins:
        m = re.search(r"\bBRA\S*\s+(?:\S+,\s*)?0x([0-9a-f]+)", x["text"])
        if m and int(m.group(1), 16) < x["addr"]:
            loops.append((int(m.group(1), 16), x["addr"]))
    for a, b in sorted(loops, key=lambda l: l[1] - l[0]):
        if b - a > 0x800:
            print("  loop %#x..%#x: %d instructions" % (a, b, (b - a) // 16 + 1))
    n = 0
    for x in ins:
        if op in x["text"] and any((x["wait"] >> b) & 1 for b in ldg_sb):
            n += 1
            inner = min((l for l in loops if l[0] <= x["addr"] <= l[1]), key=lambda l: l[1] - l[0], default=None)
            print("  %#07x %-56s wait %s  in loop %s" % (x["addr"], x["text"][:56], format(x["wait"], "06b"),
                                                         "%#x..%#x" % inner if inner else "-"))
    print("%d %s instructions wait on an LDG scoreboard" % (n, op))


if __name__ == "__main__":
    main()
