#!/usr/bin/env python
"""Aggregate an .ncu-rep source page by CUDA source line: warp instructions executed and stall samples.

    python tools/ncu_source_lines.py gpurun_out/prof_env.ncu-rep [top_n]
Needs -lineinfo at compile time and --import-source on at capture time.  Only the first
kernel launch in the report is read.
"""
import csv
import io
import subprocess
import sys


def main():
    rep = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
    out = []
    fname = ""
    hdr = None
    seen_kernels = 0
    for row in csv.reader(io.StringIO(raw)):
        if not row:
            continue
        if row[0] == "File Path":
            fname = row[1].split("/")[-1]
            continue
        if row[0] == "Function Name":
            continue
        if row[0] == "Kernel Name":
            seen_kernels += 1
            if seen_kernels > 1:
                break
            continue
        if row[0] == "Line No":
            hdr = row
            continue
        if hdr is None or row[0] == "":
            continue
        d = dict(zip(hdr[4:], row[4:]))
        try:
            inst = int(d["Instructions Executed"])
            samples = int(d["# Samples"])
        except (KeyError, ValueError):
            continue
        out.append((inst, samples, fname, row[0], row[1].strip()[:110]))
    tot_i = sum(o[0] for o in out) or 1
    tot_s = sum(o[1] for o in out) or 1
    print(f"total warp instructions {tot_i}, stall samples {tot_s}")
    print("inst%  samp%  file:line  source")
    for inst, samples, fn, ln, src in sorted(out, reverse=True)[:top]:
        print(f"{100 * inst / tot_i:5.1f}  {100 * samples / tot_s:5.1f}  {fn}:{ln}  {src}")


if __name__ == "__main__":
    main()
