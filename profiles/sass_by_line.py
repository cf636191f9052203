#!/usr/bin/env python
"""Attribute the per-instruction counters of an `ncu --page source --csv` SASS listing to CUDA source lines.

ncu's CSV source page is per SASS instruction; `nvdisasm -g` of the same cubin lists the same instructions in the
same order with `//## File "...", line N` markers (inlined frames included).  Usage:
  ncu -i rep.ncu-rep --page source --csv --kernel-name regex:<k> --launch-count 1 > k.csv
  cuobjdump -xelf all liblmono_b200.so ; python profiles/sass_by_line.py k.csv <file>.cubin <mangled-substr> [top]
"""
import csv, re, subprocess, sys, collections

def main():
    ncu_csv, cubin, fn = sys.argv[1:4]
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
    rows = list(csv.reader(open(ncu_csv)))
    hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    hdr = rows[hi]
    body = rows[hi + 1:]
    ci = {n: hdr.index(n) for n in ("Source", "Instructions Executed", "Thread Instructions Executed", "Warp Stall Sampling (All Samples)")}
    dis = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout.splitlines()
    # find the function
    lines = []          # (srcfile, line, inline_chain, opcode)
    infn = False; cur = ("?", 0)
    for l in dis:
        if l.startswith(".text."):
            infn = fn in l
            continue
        if re.match(r"\s*\.section", l):
            infn = False
        if not infn: continue
        m = re.search(r'//## File "([^"]+)", line (\d+)(.*)', l)
        if m:
            # outermost non-inlined frame is what we attribute to when "inlined at" present: keep both
            inl = re.findall(r'inlined at "([^"]+)", line (\d+)', m.group(3))
            cur = (m.group(1).split("/")[-1], int(m.group(2)), tuple((a.split("/")[-1], int(b)) for a, b in inl))
            continue
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(.*?);", l)
        if m:
            lines.append((cur, m.group(1)))
    print(f"{len(body)} ncu instructions, {len(lines)} nvdisasm instructions")
    n = min(len(body), len(lines))
    agg = collections.defaultdict(lambda: [0, 0, 0, 0])
    tot = [0, 0, 0]
    for k in range(n):
        r = body[k]
        ie, te, ss = int(r[ci["Instructions Executed"]] or 0), int(r[ci["Thread Instructions Executed"]] or 0), int(r[ci["Warp Stall Sampling (All Samples)"]] or 0)
        cur = lines[k][0]
        key = (cur[0], cur[1])
        a = agg[key]; a[0] += ie; a[1] += te; a[2] += ss; a[3] += 1
        tot[0] += ie; tot[1] += te; tot[2] += ss
    print(f"total inst {tot[0]}, samples {tot[2]}")
    srcs = {}
    for (f, ln), a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
        if f not in srcs:
            try: srcs[f] = open(f"/root/repo/lmono_b200/csrc/{f}").read().splitlines()
            except Exception: srcs[f] = []
        text = srcs[f][ln - 1].strip()[:100] if 0 < ln <= len(srcs[f]) else ""
        print(f"{f}:{ln:4d} inst {a[0]:8d} ({100*a[0]/tot[0]:4.1f}%) samples {a[2]:6d} ({100*a[2]/max(tot[2],1):4.1f}%) sass {a[3]:3d} | {text}")

main()
