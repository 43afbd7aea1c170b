#!/usr/bin/env python
"""Per-CUDA-source-line executed warp instructions and stall-sample share of the (single) kernel in an .ncu-rep captured
with --import-source on.  usage: ncu_src_lines.py <report.ncu-rep> [min inst per warp-step to list] [file filter]
The unit "warp-step" is taken from the line that counts loop iterations (the step loop's header)."""
import csv
import subprocess
import sys

rep = sys.argv[1]
thr = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
filt = sys.argv[3] if len(sys.argv) > 3 else "bgr_rollout_f32.cuh"
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True,
                     text=True).stdout
cur, hdr, rows, kern = None, None, [], None
for r in csv.reader(out.splitlines()):
    if len(r) == 2 and r[0] == "File Path":
        cur = r[1].rsplit("/", 1)[-1]
        continue
    if len(r) == 2 and r[0] == "Function Name":
        kern = r[1]
        continue
    if r and r[0] == "Line No":
        hdr = r
        iI, iS, iT = hdr.index("Instructions Executed"), hdr.index("# Samples"), hdr.index("Thread Instructions Executed")
        continue
    if hdr and len(r) > iI and r[2] == "-":
        try:
            rows.append((cur, int(r[0]), int(r[iI]), int(r[iS]), int(r[iT]), r[1]))
        except ValueError:
            pass
steps = None
for f, l, i, s, t, src in rows:
    if f == "bgr_rollout_f32.cuh" and "steps = ks + 1" in src:
        steps = float(i)
tot, ts, tt = sum(r[2] for r in rows), sum(r[3] for r in rows), sum(r[4] for r in rows)
print(kern)
print(f"warp-steps {steps:.4g}   inst/warp-step {tot / steps:.1f}   active lanes/inst {tt / tot:.1f}")
byfile = {}
for r in rows:
    byfile[r[0]] = byfile.get(r[0], 0) + r[2]
print("  ".join(f"{k}: {v / steps:.1f}" for k, v in sorted(byfile.items(), key=lambda kv: -kv[1])))
for f, l, i, s, t, src in sorted(rows, key=lambda r: (r[0], r[1])):
    if filt in f and i / steps >= thr:
        print(f"{l:5d} {i / steps:7.1f} {100 * s / ts:5.1f}%  {src.strip()[:120]}")
