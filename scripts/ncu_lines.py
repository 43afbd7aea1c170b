#!/usr/bin/env python
"""Per-SOURCE-LINE executed warp instructions per warp-step for one captured kernel.

ncu's `--page source --csv` lists executed-instruction counts per SASS instruction only; `nvdisasm -g` of the very
cubin that ran (the in-tree .o travels to the GPU box unchanged) carries the line table.  Both list the function's
instructions in address order, so they are joined by position (opcodes are cross-checked).

usage: ncu_lines.py <ncu source csv> <nvdisasm -g listing> <mangled kernel name> <sim steps per launch> [top] [opcodes.json]

With the sixth argument the per-opcode totals are also written as JSON, together with the EXECUTED fp32 arithmetic of
the kernel: warp-instructions per warp-step on the FMA pipe's arithmetic opcodes and the flops per lane-step they
stand for (FADD / FMUL = 1, FFMA = 2; packed FADD2 / FMUL2 = 2, FFMA2 = 4).  ncu's own
smsp__sass_thread_inst_executed_op_f{add,mul,fma} counters do not see the packed opcodes, so bench.py takes its
``executed_fp32_frac`` from this table.
"""
import json
import collections
import csv
import re
import sys

src_csv, listing, kernel, steps = sys.argv[1], sys.argv[2], sys.argv[3], float(sys.argv[4])
top = int(sys.argv[5]) if len(sys.argv) > 5 else 60
ops_json = sys.argv[6] if len(sys.argv) > 6 else None
if steps <= 0:
    steps = 32.0          # no unit count known: the table is then in warp-instructions per launch

rows = list(csv.reader(open(src_csv)))
hdr = rows[1]
ia, isrc = hdr.index("Instructions Executed"), hdr.index("Source")
ncu = []
for r in rows[2:]:
    try:
        ncu.append((int(r[ia]), r[isrc].strip()))
    except (ValueError, IndexError):
        continue

ins, cur, on = [], ("?", 0), False
for line in open(listing):
    if line.startswith(".text."):
        on = line.strip() == f".text.{kernel}:"
        continue
    if not on:
        continue
    m = re.match(r'\s*//## File "(.*)", line (\d+)', line)
    if m:
        cur = (m.group(1).rsplit("/", 1)[-1], int(m.group(2)))
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
    if m:
        ins.append((cur, m.group(2).strip()))

def opcode(s):
    m = re.match(r"(@!?U?P\d+\s+)?([A-Z0-9_]+)", s)
    return m.group(2) if m else s

if len(ins) != len(ncu):
    print(f"warning: {len(ins)} instructions in the listing vs {len(ncu)} in the capture", file=sys.stderr)
bad = sum(opcode(a[1]) != opcode(b[1]) for a, b in zip(ins, ncu))
if bad:
    print(f"warning: {bad} opcode mismatches -- listing and capture are not the same build", file=sys.stderr)

W = steps / 32
per_line, per_line_ops = collections.Counter(), collections.defaultdict(collections.Counter)
for (loc, text), (n, _) in zip(ins, ncu):
    per_line[loc] += n / W
    per_line_ops[loc][opcode(text)] += n / W
total = sum(per_line.values())
print(f"TOTAL warp-instructions per warp-step: {total:.2f}")
if ops_json:
    by_op = collections.Counter()
    for (loc, text), (n, _) in zip(ins, ncu):
        by_op[opcode(text)] += n / W
    flops = {"FADD": 1, "FMUL": 1, "FFMA": 2, "FADD2": 2, "FMUL2": 2, "FFMA2": 4}
    json.dump({"kernel": kernel, "units_per_launch": steps, "warp_inst_per_warp_step": total,
               "by_opcode": {k: round(v, 4) for k, v in by_op.most_common()},
               "fp32_warp_inst_per_warp_step": sum(by_op[k] for k in flops),
               "fp32_flops_per_lane_step": sum(by_op[k] * f for k, f in flops.items()),
               "counting": "FADD/FMUL 1, FFMA 2, FADD2/FMUL2 2, FFMA2 4 flops per lane; compares, min/max, conversions, MUFU not counted"},
              open(ops_json, "w"), indent=1)
for loc, v in per_line.most_common(top):
    ops = " ".join(f"{k}:{x:.1f}" for k, x in per_line_ops[loc].most_common(6))
    print(f"{v:7.2f}  {loc[0]}:{loc[1]:<5d} {ops}")
