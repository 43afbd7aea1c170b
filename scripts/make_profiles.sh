#!/bin/bash
# Turn what `gpurun -- bash scripts/gpu_round.sh` (and optionally `gpurun --gpus N -- bash scripts/scaling_round.sh`)
# brought back in gpurun_out/ into the tracked summaries under profiles/<name>/.  Runs in the build container (needs
# ncu, cuobjdump, nvdisasm; no GPU).  The objects under csrc/build must be the build that ran on the box (checked:
# gpurun_out/build_id.txt against bgr_build_id() of the in-tree library) -- the per-source-line listings join ncu's
# per-SASS counts with ITS line table, and bench.py quotes the evidence only for that build id.
#   usage: scripts/make_profiles.sh profiles/r2_final
set -e
out=${1:?usage: scripts/make_profiles.sh profiles/<name>}
root=$(cd "$(dirname "$0")/.." && pwd)
cd "$root"
mkdir -p "$out"
tmp=$(mktemp -d)
here=$(python -c "import bgr_pathplanning_control_b200 as b; print(b.build_id())")
ran=$(cat gpurun_out/build_id.txt 2>/dev/null | head -1)
if [ "$here" != "$ran" ]; then echo "build id mismatch: in-tree library $here, GPU run $ran -- rebuild the run's sources first"; exit 1; fi
echo "$here" > "$out/build_id.txt"
steps=2097152000   # sim-steps of one launch of the default bench (1 Mi episodes x 2000 steps)

listing() {  # <object stem> -> $tmp/<stem>.txt
  mkdir -p "$tmp/$1" && (cd "$tmp/$1" && cuobjdump -xelf all "$root/bgr_pathplanning_control_b200/csrc/build/$1.o" > /dev/null \
    && nvdisasm -g "$tmp/$1"/*.cubin > "$tmp/$1.txt")
}
listing bgr_rollout_inst_f32_lean
listing bgr_rollout_inst_f32_full
listing bgr_mpc

summarise() {   # <report stem> <summary name> <listing stem> <mangled kernel> <units per launch>
  [ -f "gpurun_out/$1.ncu-rep" ] || { echo "skip $1 (no report)"; return; }
  ncu -i "gpurun_out/$1.ncu-rep" --page raw --csv > "$tmp/$1_raw.csv" 2> /dev/null
  ncu -i "gpurun_out/$1.ncu-rep" --page source --csv > "$tmp/$1_src.csv" 2> /dev/null
  python scripts/ncu_summary.py "$tmp/$1_raw.csv" $5 > "$out/ncu_full_$2.json"
  python scripts/ncu_lines.py "$tmp/$1_src.csv" "$tmp/$3.txt" "$4" $5 60 "$out/opcodes_$2.json" > "$out/hot_lines_$2.txt"
  head -1 "$out/hot_lines_$2.txt"
}
summarise prof_pp pp_pid_kinematic bgr_rollout_inst_f32_lean _ZN3bgr18rollout_kernel_f32ILi0ELi0ELi0ELb0ELb1ELb0EEEvNS_11RolloutArgsE $steps
summarise prof_stanley stanley_lqg_dynamic bgr_rollout_inst_f32_lean _ZN3bgr18rollout_kernel_f32ILi1ELi1ELi1ELb0ELb1ELb0EEEvNS_11RolloutArgsE $steps
# full-build kernels: units per launch from the bench line of the same workload (sim-steps actually executed)
units() { python - "$1" "$2" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(int(d["roofline"].get("sim_steps_per_launch") or d["value"] * d["ms_per_step"] * 1e-3 / float(sys.argv[2])))
except Exception:
    print(0)
PY
}
summarise prof_replan pp_replan bgr_rollout_inst_f32_full _ZN3bgr18rollout_kernel_f32ILi0ELi0ELi0ELb1ELb1ELb0EEEvNS_11RolloutArgsE "$(units gpurun_out/plain_replan.log 1)"
summarise prof_mc config5_pp bgr_rollout_inst_f32_full _ZN3bgr18rollout_kernel_f32ILi0ELi0ELi0ELb1ELb1ELb0EEEvNS_11RolloutArgsE 0
summarise prof_mc_stlqg config5_stanley_lqg_queue bgr_rollout_inst_f32_full _ZN3bgr18rollout_kernel_f32ILi1ELi1ELi1ELb1ELb1ELb1EEEvNS_11RolloutArgsE 0
for r in mc:config5_pp mc_stlqg:config5_stanley_lqg_queue; do
  [ -f "gpurun_out/prof_${r%%:*}.ncu-rep" ] && python scripts/ncu_regions.py "gpurun_out/prof_${r%%:*}.ncu-rep" > "$out/regions_${r##*:}.txt"
done
if [ -f gpurun_out/prof_mpc.ncu-rep ]; then
  ncu -i gpurun_out/prof_mpc.ncu-rep --page raw --csv > "$tmp/mpc_raw.csv" 2> /dev/null
  python scripts/ncu_summary.py "$tmp/mpc_raw.csv" > "$out/ncu_full_mpc_factored.json"
fi

for w in default pp_replan ref; do
  [ -f gpurun_out/bench_$w.log ] && tail -1 gpurun_out/bench_$w.log > "$out/bench_$w.json"
done
[ -f gpurun_out/launches.csv ] && cp gpurun_out/launches.csv "$out/ncu_launches_bench_pp.csv"
[ -f gpurun_out/parity_fractions.json ] && cp gpurun_out/parity_fractions.json "$out/parity_fractions.json"
[ -f gpurun_out/mpc_suboptimality.json ] && cp gpurun_out/mpc_suboptimality.json "$out/mpc_suboptimality.json"

if [ -f gpurun_out/clocks.csv ]; then
python - "$out" <<'PY'
import csv, statistics, sys
rows = list(csv.reader(open("gpurun_out/clocks.csv")))
body = [r for r in rows[1:] if len(r) > 8]
sm = [float(r[1].split()[0]) for r in body]
pw = [float(r[3].split()[0]) for r in body]
bad = sum(1 for r in body if any(c.strip() == "Active" for c in r[5:9]))
open(sys.argv[1] + "/clocks_summary.txt", "w").write(
    f"nvidia-smi -lms 200 during the bench runs of scripts/gpu_round.sh: {len(sm)} samples, clocks.sm median "
    f"{statistics.median(sm)} MHz (max {max(sm)}, clocks.max.sm {body[0][2].strip()}), power max {max(pw)} W, samples with "
    f"hw_slowdown/hw_thermal/sw_thermal/sw_power_cap active: {bad}\n")
PY
fi

# SASS evidence: mnemonic counts of the whole library AND the actual listing around the sites that prove the design
# (TMA bulk copies + mbarrier, warp argmin, packed fp32 arithmetic) in the two headline kernels
(echo "# SASS mnemonics of libbgr_rollout.so (cuobjdump -sass | counts): TMA bulk copies (UBLKCP), mbarrier (SYNCS), warp argmin (CREDUX), packed fp32 (FADD2 / FMUL2 / FFMA2), SFU sin / cos (MUFU.SIN / MUFU.COS); no HMMA / UTC*MMA by design"
 cuobjdump -sass bgr_pathplanning_control_b200/libbgr_rollout.so \
   | grep -oE "UBLKCP\.S\.G|SYNCS\.[A-Z0-9.]+|CREDUX\.[A-Z]+|FADD2|FMUL2|FFMA2|MUFU\.SIN|MUFU\.COS|F2F\.[A-Z0-9.]+|HMMA[A-Z0-9.]*|UTC[A-Z0-9.]*MMA[A-Z0-9.]*" \
   | sort | uniq -c) > "$out/sass_evidence.txt"
for k in ILi0ELi0ELi0ELb0ELb1ELb0:pp_pid_kinematic ILi1ELi1ELi1ELb0ELb1ELb0:stanley_lqg_dynamic; do
  scripts/sass_kernel.sh bgr_pathplanning_control_b200/csrc/build/bgr_rollout_inst_f32_lean.o "rollout_kernel_f32${k%%:*}" "$out/sass_${k##*:}.txt" > /dev/null
  python - "$out/sass_${k##*:}.txt" <<'PY'
import re, sys
p = sys.argv[1]
lines = open(p).read().splitlines()
keep = set()
for i, ln in enumerate(lines):
    if re.search(r"UBLKCP|SYNCS\.|CREDUX|FADD2|FMUL2|FFMA2|MUFU\.(SIN|COS|RSQ|EX2)|VOTE\.ANY", ln):
        keep.update(range(max(0, i - 3), min(len(lines), i + 4)))
out, last = [lines[0]], 0
for i in sorted(keep):
    if i > last + 1:
        out.append("        ...")
    out.append(lines[i])
    last = i
open(p, "w").write("# nvdisasm -g excerpt (3 lines of context around every TMA / mbarrier / warp-redux / packed-fp32 / SFU site)\n"
                   + "\n".join(out) + "\n")
PY
done

if ls gpurun_out/scale_*.json > /dev/null 2>&1; then
python - "$out" <<'PY'
import glob, json, re, shutil, sys
out = sys.argv[1]
txt = ("bench.py under torchrun (one process per GPU, NCCL), default command = config 2 headline + secondary configs;\n"
       "one multi-GPU box, runs back to back (scripts/scaling_round.sh).  value = all ranks' units / max-over-ranks device time.\n\n")
for f in sorted(glob.glob("gpurun_out/scale_*.json"), key=lambda p: int(re.findall(r"_(\d+)\.json", p)[0])):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
    except Exception:
        continue
    shutil.copy(f, out)
    n = d["n_gpus"]
    txt += f"N={n}  config2 value {d['value']:.4e}  e2e {d['e2e']['value']:.4e} ({d['e2e']['frac_of_value']:.3f})  gather_ms {d.get('gather_ms')}\n"
    for k, r in (d.get("secondary") or {}).items():
        if "error" in r:
            txt += f"      {k}: ERROR {r['error']}\n"
        else:
            txt += (f"      {k}: value {r['value']:.4e} {r['unit']}  ms/step {r['ms_per_step']:.3f}  e2e {r['e2e']['value']:.4e}"
                    f"  gather_ms {r.get('gather_ms')}  scaling {r['scaling']}\n")
open(out + "/scaling.txt", "w").write(txt)
print(txt)
PY
fi
rm -rf "$tmp"
ls "$out"
