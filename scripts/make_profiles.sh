#!/bin/bash
# Turn what `gpurun -- bash scripts/gpu_round.sh` (and optionally `gpurun --gpus 8 -- bash scripts/scaling_round.sh`)
# brought back in gpurun_out/ into the tracked summaries under profiles/<name>/.  Runs in the build container (needs
# ncu, cuobjdump, nvdisasm; no GPU).  The lean .o under csrc/build must be the build that ran on the box: the
# per-source-line listings join ncu's per-SASS counts with ITS line table.
#   usage: scripts/make_profiles.sh profiles/r2_final
set -e
out=${1:?usage: scripts/make_profiles.sh profiles/<name>}
root=$(cd "$(dirname "$0")/.." && pwd)
cd "$root"
mkdir -p "$out"
tmp=$(mktemp -d)
steps=2097152000   # sim-steps of one launch of the default bench (1 Mi episodes x 2000 steps)

(cd "$tmp" && cuobjdump -xelf all "$root/bgr_pathplanning_control_b200/csrc/build/bgr_rollout_inst_f32_lean.o" > /dev/null \
  && nvdisasm -g "$tmp"/*.cubin > "$tmp/lean.txt")

summarise() {   # <report stem> <summary name> <mangled kernel>
  [ -f "gpurun_out/$1.ncu-rep" ] || { echo "skip $1 (no report)"; return; }
  ncu -i "gpurun_out/$1.ncu-rep" --page raw --csv > "$tmp/$1_raw.csv" 2> /dev/null
  ncu -i "gpurun_out/$1.ncu-rep" --page source --csv > "$tmp/$1_src.csv" 2> /dev/null
  python scripts/ncu_summary.py "$tmp/$1_raw.csv" $steps > "$out/ncu_full_$2.json"
  python scripts/ncu_lines.py "$tmp/$1_src.csv" "$tmp/lean.txt" "$3" $steps 60 > "$out/hot_lines_$2.txt"
  head -1 "$out/hot_lines_$2.txt"
}
summarise prof_pp pp_pid_kinematic _ZN3bgr18rollout_kernel_f32ILi0ELi0ELi0ELb0ELb1EEEvNS_11RolloutArgsE
summarise prof_stanley stanley_lqg_dynamic _ZN3bgr18rollout_kernel_f32ILi1ELi1ELi1ELb0ELb1EEEvNS_11RolloutArgsE
if [ -f gpurun_out/prof_mpc.ncu-rep ]; then
  ncu -i gpurun_out/prof_mpc.ncu-rep --page raw --csv > "$tmp/mpc_raw.csv" 2> /dev/null
  python scripts/ncu_summary.py "$tmp/mpc_raw.csv" > "$out/ncu_mpc_factored.json"
fi

for w in pp stanley mpc ref skidpad_mc pp_replan; do
  [ -f gpurun_out/bench_$w.log ] && tail -1 gpurun_out/bench_$w.log > "$out/bench_$w.json"
done
[ -f gpurun_out/launches.csv ] && cp gpurun_out/launches.csv "$out/ncu_launches_bench_pp.csv"

if [ -f gpurun_out/clocks.csv ]; then
python - "$out" <<'EOF'
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
EOF
fi

(echo "# SASS mnemonics of libbgr_rollout.so (cuobjdump -sass | counts): TMA bulk copies (UBLKCP), mbarrier (SYNCS), warp argmin (CREDUX), packed fp32 (FADD2 / FMUL2), SFU sin / cos (MUFU.SIN / MUFU.COS); no HMMA / UTC*MMA by design"
 cuobjdump -sass bgr_pathplanning_control_b200/libbgr_rollout.so \
   | grep -oE "UBLKCP\.S\.G|SYNCS\.[A-Z0-9.]+|CREDUX\.[A-Z]+|FADD2|FMUL2|FFMA2|MUFU\.SIN|MUFU\.COS|HMMA[A-Z0-9.]*|UTC[A-Z0-9.]*MMA[A-Z0-9.]*" \
   | sort | uniq -c) > "$out/sass_evidence.txt"

if ls gpurun_out/scale_pp_*.json > /dev/null 2>&1; then
python - "$out" <<'EOF'
import glob, json, re, shutil, sys
out = sys.argv[1]
rows, base = [], None
for f in sorted(glob.glob("gpurun_out/scale_*_*.json"), key=lambda p: (("stanley" in p), int(re.findall(r"_(\d+)\.json", p)[0]))):
    d = json.load(open(f))
    wl = "stanley_lqg" if "stanley" in f else "pp_sweep"
    if wl == "pp_sweep" and d["n_gpus"] == 1:
        base = d["value"]
    rows.append((wl, d["n_gpus"], d["value"], d["ms_per_step"], d["e2e"]["value"], d.get("gather_ms")))
    shutil.copy(f, out)
txt = ("bench.py under torchrun (one process per GPU, NCCL, weak scaling: 1 Mi episodes x 2000 steps per GPU), 20 timed "
       "steps, 3 warm-up;\none 8-GPU box, runs back to back (scripts/scaling_round.sh).  value = all ranks' sim-steps / "
       "max-over-ranks device time;\ne2e = through bgr_rollout_host with pinned host buffers; the final NCCL gather is "
       "outside the timed region.\n\nworkload      n_gpus  value [sim-steps/s]  ms/step   e2e [sim-steps/s]   "
       "final gather [ms]   vs pp N=1\n")
for wl, n, v, ms, e, g in rows:
    txt += f"{wl:<13} {n:^6}  {v:.4e}           {ms:6.2f}    {e:.3e}           {('%.2f' % g) if g else '-':<18} {v / base:.2f}\n"
open(out + "/scaling.txt", "w").write(txt)
print(txt)
EOF
fi
rm -rf "$tmp"
ls "$out"
