#!/usr/bin/env python
"""Aggregate an ncu report's per-SASS-instruction warp-stall samples to CUDA source lines of xtb_scf.cu.
usage: tools/ncu_lines.py gpurun_out/scf_rN.ncu-rep [top]"""
import collections, csv, io, re, subprocess, sys, tempfile, os
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tmp = tempfile.mkdtemp()
subprocess.run(f"cd {tmp} && cuobjdump -xelf all {root}/dxtb_b200/_C.so >/dev/null 2>&1 && nvdisasm -g xtb_scf.sm_100a.cubin > scf.sass", shell=True, check=True)
# per function: offset -> (file,line)
maps = {}; cur = None; line = None
for l in open(f"{tmp}/scf.sass"):
    m = re.match(r"\s*\.text\.(\S+):", l)
    if m: cur = m.group(1); maps[cur] = {}; line = None; continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m: line = (m.group(1).split("/")[-1], int(m.group(2))); continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+\S+", l)
    if m and cur and line: maps[cur][int(m.group(1), 16)] = line
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
kname = rows[0][1]
hi = [i for i, r in enumerate(rows) if len(r) > 3 and r[0] == "Address"][0]
h = rows[hi]; si = h.index("# Samples"); ii = h.index("Instructions Executed")
data = [(int(r[0], 16), int(r[si] or 0), int(r[ii] or 0)) for r in rows[hi + 1:] if len(r) > si and r[0]]
base = min(d[0] for d in data)
# choose the map whose size matches best
n = len(data); amap = min(maps.values(), key=lambda m: abs(len(m) - n)) if maps else {}
agg = collections.defaultdict(lambda: [0, 0])
for a, s_, i_ in data:
    ln = amap.get(a - base, ("?", 0)); agg[ln][0] += s_; agg[ln][1] += i_
tot = sum(v[0] for v in agg.values()) or 1; toti = sum(v[1] for v in agg.values()) or 1
src = {f: open(f"{root}/dxtb_b200/csrc/{f}").read().splitlines() for f in ("xtb_scf.cu", "xtb_common.cuh")}
print(kname[:80], "instructions", n)
for ln, v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    text = src[ln[0]][ln[1] - 1].strip()[:105] if ln[0] in src and ln[1] > 0 else ""
    print(f"{100*v[0]/tot:5.1f}% smp {100*v[1]/toti:5.1f}% ins {ln[0]}:{ln[1]:4d} {text}")
