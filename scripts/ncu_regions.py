#!/usr/bin/env python
"""Executed warp instructions, lane utilisation and stall-sample share per REGION of the fp32 step kernel (regions are
found by marker comments in the source embedded in the report).  usage: ncu_regions.py <report.ncu-rep>"""
import collections
import csv
import subprocess
import sys

out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "source", "--csv", "--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout
cur, hdr, rows = None, None, []
for r in csv.reader(out.splitlines()):
    if len(r) == 2 and r[0] == "File Path":
        cur = r[1].rsplit("/", 1)[-1]
        continue
    if len(r) == 2:
        continue
    if r and r[0] == "Line No":
        hdr = r
        iI, iT, iS = hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed"), hdr.index("# Samples")
        continue
    if hdr and len(r) > iI and r[2] == "-":
        def num(x):
            try:
                return int(x)
            except ValueError:
                return 0
        try:
            rows.append((cur, int(r[0]), num(r[iI]), num(r[iT]), num(r[iS]), r[1]))
        except ValueError:
            pass
MARKS = [("auto nearest_general = ", "general: descent"), ("const int2 tw = __ldg(A.trk.twin", "general: twin"),
         ("} else if (rr.y >= 0 && dw < rival_acc2)", "general: rival walk"), ("// ---- warp-cooperative exact scan", "general: warp scan"),
         ("auto nearest = [&]", "nearest (window / wrapper)"), ("// ---- 2. steering law", "steering law"),
         ("// ---- 3. curvature lookup", "accel law"), ("// ---- 5. error logging", "errors / metrics"),
         ("// ---- 6. cone hits", "cones"), ("// ---- 7. lap counter", "lap / audit / dumps"),
         ("// ---- 8.", "vehicle model + end tests"), ("for (int k = 0; DYN ||", "loop head + noise"),
         ("auto begin_episode = ", "begin / finish episode"), ("template <int STEER, int ACCEL, int MODEL, bool EXTRAS", "prologue")]
f32 = sorted((r for r in rows if r[0] == "bgr_rollout_f32.cuh"), key=lambda r: r[1])
# (the report lists only the lines that own instructions: the marker lines are looked up in the source file itself,
#  which must be the one the profiled library was built from)
import os
src_path = sys.argv[2] if len(sys.argv) > 2 else os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                                                              "bgr_pathplanning_control_b200", "csrc", "bgr_rollout_f32.cuh")
marks = []
for l, src in enumerate(open(src_path), 1):
    for key, name in MARKS:
        if key in src:
            marks.append((l, name))
marks.sort()
def region(f, l):
    if f == "bgr_device.cuh":
        return "philox / gaussians (bgr_device.cuh)"
    if f != "bgr_rollout_f32.cuh":
        return "intrinsics headers"
    name = "helpers (norm2, packed ops, sincos)"
    for ml, mn in marks:
        if l >= ml:
            name = mn
    return name
steps = next((float(r[2]) for r in f32 if "steps = ks + 1" in r[5]), 1.0)
a, b, c = collections.Counter(), collections.Counter(), collections.Counter()
for f, l, i, t, s, _ in rows:
    k = region(f, l)
    a[k] += i
    b[k] += t
    c[k] += s
tot, ts = sum(a.values()), sum(c.values())
print(f"warp-steps {steps:.4g}; {tot / steps:.1f} warp-instructions per warp-step; {sum(b.values()) / tot:.1f} active lanes per instruction")
for k, v in a.most_common():
    print(f"{k:40s} inst/step {v / steps:7.1f} ({100 * v / tot:4.1f} %)  lanes {b[k] / max(v, 1):5.1f}  stall samples {100 * c[k] / ts:4.1f} %")
