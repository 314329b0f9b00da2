#!/usr/bin/env python
"""Top source lines of the first kernel in an .ncu-rep by warp instructions executed (SASS rows
aggregated per CUDA source line), with the average active lanes and the stall-sample share.

    python tools/ncu_source_top.py gpurun_out/prof.ncu-rep [top_n] > profiles/..._source_top.txt
Needs -lineinfo at compile time and --import-source on at capture time."""
import collections
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
hdr = None
fname = ""
cur = None
nk = 0
inst, thr, samp, srcs = collections.Counter(), collections.Counter(), collections.Counter(), {}
kernel = ""
for row in csv.reader(io.StringIO(raw)):
    if not row:
        continue
    if row[0] == "Kernel Name":
        nk += 1
        if nk > 1:
            break
        kernel = row[1]
        continue
    if row[0] == "File Path":
        fname = row[1].split("/")[-1]
        continue
    if row[0] == "Line No":
        hdr = row
        continue
    if hdr is None:
        continue
    if row[0].isdigit():
        cur = (fname, int(row[0]))
        srcs[cur] = row[1].strip()[:100]
        continue
    if row[0] == "" and cur is not None:
        d = dict(zip(hdr[4:], row[4:]))
        try:
            inst[cur] += int(d["Instructions Executed"])
            thr[cur] += int(d["Thread Instructions Executed"])
            samp[cur] += int(d["# Samples"])
        except (KeyError, ValueError):
            pass
tot, ts = sum(inst.values()) or 1, sum(samp.values()) or 1
print(f"# {rep}: {kernel}")
print(f"# warp instructions attributed to source lines: {tot}; stall samples: {ts}")
byfile = collections.Counter()
for k, v in inst.items():
    byfile[k[0]] += v
print("# by file:", ", ".join(f"{k} {100 * v / tot:.1f}%" for k, v in byfile.most_common()))
print("# inst%  stall-sample%  avg active lanes  file:line  source")
for k, v in inst.most_common(top):
    print(f"{100 * v / tot:5.1f}  {100 * samp[k] / ts:5.1f}  {thr[k] / max(v, 1):5.1f}  {k[0]}:{k[1]}  {srcs[k]}")

bys = collections.Counter()
for k, v in samp.items():
    bys[k[0]] += v
print("# stall samples by file:", ", ".join(f"{k} {100 * v / ts:.1f}%" for k, v in bys.most_common()))
print("# top lines by stall samples")
for k, v in samp.most_common(top // 2):
    print(f"{100 * inst[k] / tot:5.1f}  {100 * v / ts:5.1f}  {thr[k] / max(inst[k], 1):5.1f}  {k[0]}:{k[1]}  {srcs[k]}")
