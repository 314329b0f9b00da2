#!/usr/bin/env python
"""Warp-instruction share of each phase of env_kernel from an .ncu-rep source page.
    python tools/ncu_phase_breakdown.py gpurun_out/prof_env4.ncu-rep [n_ctas]"""
import collections
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
nctas = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
hdr = None
fname = ""
agg = collections.Counter()
cur = None
for row in csv.reader(io.StringIO(raw)):
    if not row:
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
        continue
    if row[0] == "" and cur is not None:  # a SASS row under the current source line
        d = dict(zip(hdr[4:], row[4:]))
        try:
            agg[cur] += int(d["Instructions Executed"])
        except (KeyError, ValueError):
            pass
tot = sum(agg.values())
print("total warp instructions", tot)
src = open("routing-board-generation_b200/csrc/connector_kernel.cu").read().splitlines()


def find(s):
    for i, l in enumerate(src):
        if s in l:
            return i + 1
    raise KeyError(s)


marks = [("helpers (carve, path count, random_action)", 1), ("kernel head + tables", find("env_kernel(const EnvParams p)")), ("phase1 load", find("---- phase 1")),
         ("phase2 agents", find("---- phase 2")), ("collisions+move", find("collisions: same destination")), ("phase3 mask/reward", find("---- phase 3")),
         ("per-env", find("for (int m = tid; m < Ec; m += nt)")), ("phase3b swap", find("if (*s.anyhit)")), ("per-agent out", find("const int term = s.term[m];") - 3),
         ("phase4 bulk", find("---- phase 4")), ("end", len(src) + 1)]
by = collections.Counter()
for (f, l), v in agg.items():
    if f == "connector_kernel.cu":
        for (n, a), (_, b) in zip(marks, marks[1:]):
            if a <= l < b:
                by[n] += v
    else:
        by[f] += v
for k, v in by.most_common():
    print(f"{k:45s} {100 * v / tot:5.1f}%  {v / nctas:9.0f} per CTA")
