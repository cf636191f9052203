#!/usr/bin/env python
"""Instruction-mix summary of every kernel in liblmono_b200.so (cuobjdump -sass): instruction count and the mnemonics that
say how a kernel talks to memory and to its neighbours (global / shared / generic accesses -- the distributed-shared-memory stores of a cluster kernel are generic ST --, shuffles, barriers,
cluster barriers, programmatic-dependent-launch waits, fp64 pipe, atomics).  Usage: python profiles/sass_summary.py > profiles/sass_summary_r02.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "lmono_b200", "csrc", "liblmono_b200.so")
txt = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
groups = [("LDG", r"^LDG"), ("LDG.nc", r"^LDG.*CONSTANT"), ("STG", r"^STG"), ("LDS", r"^LDS"), ("STS", r"^STS"), ("LDL/STL", r"^(LDL|STL)"),
          ("ATOM/RED", r"^(ATOM|RED|ATOMG|ATOMS)"), ("SHFL", r"^SHFL"), ("VOTE", r"^VOTE"), ("BAR", r"^BAR"),
          ("cluster bar", r"^UCGABAR|^CGABAR"), ("generic LD/ST", r"^(ST|LD)(\.|$)"), ("PDL wait", r"^ACQBULK"),
          ("fp64", r"^(DFMA|DMUL|DADD|DSETP|MUFU.*64)"), ("fp32", r"^(FFMA|FMUL|FADD|FSETP|FMNMX)"), ("tcgen05/TMA", r"^(UTC|UTMA|UBLKCP|UTMALDG)")]
kern = None
stats = {}
for line in txt.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        kern = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip().split("(")[0]
        stats[kern] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", line)
    if m and kern:
        op = m.group(1)
        stats[kern]["total"] += 1
        for name, pat in groups:
            if re.search(pat, op):
                stats[kern][name] += 1
names = [g[0] for g in groups]
print("SASS instruction mix of liblmono_b200.so (sm_100a), static counts per kernel")
print(f"{'kernel':44s} {'total':>7s} " + " ".join(f"{n:>11s}" for n in names))
for k in sorted(stats, key=lambda k: -stats[k]["total"]):
    if k.startswith("k_tl_stamp") or k == "k_nop":
        continue
    print(f"{k[:44]:44s} {stats[k]['total']:7d} " + " ".join(f"{stats[k][n]:11d}" for n in names))
