#!/usr/bin/env python
"""Static proxy for the step kernel's instructions per step: SASS instructions of one kernel whose source line lies in
the time loop but outside the rare paths (general nearest search, general look-ahead loop, heading wrap).  Lets a kernel
edit be judged without a GPU; the dynamic number comes from ncu (scripts/ncu_lines.py).
usage: hot_static.py <object.o> <mangled-kernel-substring>"""
import os
import re
import subprocess
import sys
import tempfile

root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
src = open(os.path.join(root, "bgr_pathplanning_control_b200", "csrc", "bgr_rollout_f32.cuh")).read().splitlines()


def line_of(pat, start=0):
    for i in range(start, len(src)):
        if pat in src[i]:
            return i + 1
    raise SystemExit(f"marker not found: {pat}")


kernel_a = line_of("rollout_kernel_f32(const __grid_constant__ RolloutArgs A)")     # helpers above it inherit
loop_a = line_of("for (int k = 0; DYN || k < A.max_steps; ++k)")
loop_b = line_of("#undef BGR_D2", loop_a)
rare = [(line_of("auto nearest_general = [&]"), line_of("auto nearest = [&]") - 1),
        (line_of("int rem = n_total - 1 - ti;"), line_of("if (!EXTRAS || pp_on) {") - 1),
        (line_of("const float kw_ = rintf"), line_of("wraps += static_cast<int>(kw_);")),
        (line_of("if constexpr (DYN) {", loop_a), line_of("} else {", line_of("if constexpr (DYN) {", loop_a)))]
out = tempfile.mktemp()
subprocess.check_call([os.path.join(root, "scripts", "sass_kernel.sh"), sys.argv[1], sys.argv[2], out], stdout=subprocess.DEVNULL)
hot, by_op, main, cur_hot = 0, {}, False, False
for ln in open(out):
    m = re.search(r'File "([^"]*)", line (\d+)', ln)
    if m:
        if m.group(1).endswith("bgr_rollout_f32.cuh"):
            L = int(m.group(2))
            if L < kernel_a:
                continue                      # an inlined helper (sub2, norm2, lds2p, ...): belongs to its caller
            cur_hot = loop_a <= L < loop_b and not any(a <= L <= b for a, b in rare)
        continue
    m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(@!?U?P\d+\s+)?([A-Z0-9_]+)", ln)
    if m and cur_hot:
        hot += 1
        by_op[m.group(2)] = by_op.get(m.group(2), 0) + 1
os.remove(out)
print(hot, "static hot-path instructions;", " ".join(f"{k}:{v}" for k, v in sorted(by_op.items(), key=lambda kv: -kv[1])[:18]))
