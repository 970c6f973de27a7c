#!/usr/bin/env python
"""SASS excerpt of the shipped library: python profiles/tools/sass_excerpt.py > profiles/r2/sass_excerpt_<tag>.txt
Mnemonic counts of both fused step kernels and, for the plain column loop of each, every bulk copy, mbarrier op,
barrier, shuffle and global access with its scoreboard / wait mask (decoded by sass_scoreboards.decode)."""
import collections, os, re, subprocess, sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from sass_scoreboards import decode  # noqa: E402

lib = os.path.join(HERE, "..", "..", "fingering_dynamics_b200", "csrc", "libfdlbm.so")
print("SASS of the shipped library (cuobjdump -sass fingering_dynamics_b200/csrc/libfdlbm.so, sm_100a)")
print("arch lines:", subprocess.run("cuobjdump -lelf %s | head -3" % lib, shell=True, capture_output=True, text=True).stdout.strip().replace("\n", " | "))
keys = ["UBLKCP", "SYNCS", "LDGSTS", "LDGDEPBAR", "DEPBAR", "LDG", "STG", "LDS", "SHFL", "BAR", "DFMA", "DADD", "DMUL", "FFMA2", "FADD2",
        "FMUL2", "FFMA", "MUFU", "IMAD", "BRA"]


def ops(seq):
    c = collections.Counter()
    for x in seq:
        t = x["text"].split()
        c[(t[1] if t[0].startswith("@") else t[0]).split(".")[0]] += 1
    return c


for kern in ("k_fusedIdLi128ELi2048", "k_fused_f32pILi2048"):
    ins = decode(lib, kern)
    c = ops(ins)
    print("\n== %s: %d instructions" % (kern, len(ins)))
    print("  mnemonic counts (static): " + ", ".join("%s %d" % (k, c[k]) for k in keys if c[k]))
    loops = []
    for x in ins:
        m = re.search(r"\bBRA\S*\s+(?:\S+,\s*)?0x([0-9a-f]+)", x["text"])
        if m and int(m.group(1), 16) < x["addr"] and 0x2000 < x["addr"] - int(m.group(1), 16) < 0x8000:
            loops.append((int(m.group(1), 16), x["addr"]))
    a, b = min(loops, key=lambda l: l[1] - l[0])
    body = [x for x in ins if a <= x["addr"] <= b]
    cb = ops(body)
    print("  plain column loop %#x..%#x: %d instructions static (%.1f KB); %s"
          % (a, b, len(body), len(body) * 16 / 1024.0, ", ".join("%s %d" % (k, cb[k]) for k in keys if cb[k])))
    print("  every bulk copy / mbarrier / barrier / shuffle / global access of that loop (addr, instruction, write-sb, wait mask):")
    for x in body:
        if re.search(r"UBLKCP|SYNCS|BAR\.SYNC|SHFL|LDG\b|LDG\.|LDGSTS|STG|DEPBAR", x["text"]):
            print("    %#07x  %-74s wr %d wait %s" % (x["addr"], x["text"][:74], x["wr"], format(x["wait"], "06b")))
