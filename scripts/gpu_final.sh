#!/bin/bash
# Closing GPU session of a round, AFTER scripts/gpu_round.sh + scripts/make_profiles.sh have put the ncu summaries of
# the loaded build under profiles/<dir>: re-capture the MPC evaluator, run the parity tests, then print the default
# bench line with the evidence matched to the build (that line is the tracked profiles/<dir>/bench_default.json).
#   usage: gpurun --timeout 600 -- bash scripts/gpu_final.sh profiles/r2_final
dir=${1:-profiles/r2_final}
mkdir -p gpurun_out
python -c "import bgr_pathplanning_control_b200 as b; print(b.build_id())" > gpurun_out/build_id.txt; cat gpurun_out/build_id.txt "$dir/build_id.txt"
python bench.py --workload mpc --steps 2 --warmup 1 --no-cpu-baseline --no-secondary > gpurun_out/plain_mpc.log 2>&1 && \
timeout 300 ncu --set full --clock-control none --import-source on -k regex:mpc_factored_kernel -s 1 -c 1 -o gpurun_out/prof_mpc -f \
  python bench.py --workload mpc --steps 2 --warmup 1 --no-cpu-baseline --no-secondary > gpurun_out/ncu_mpc.log 2>&1
echo "ncu mpc rc=$?"
if [ -f gpurun_out/prof_mpc.ncu-rep ]; then
  ncu -i gpurun_out/prof_mpc.ncu-rep --page raw --csv > gpurun_out/mpc_raw.csv 2> /dev/null
  python scripts/ncu_summary.py gpurun_out/mpc_raw.csv > "$dir/ncu_full_mpc_factored.json" && cp "$dir/ncu_full_mpc_factored.json" gpurun_out/
fi
timeout 900 python -m pytest tests -m gpu -q --timeout 600 --tb=short > gpurun_out/pytest_full.log 2>&1; echo "pytest rc=$?"
grep -E "passed|failed|error" gpurun_out/pytest_full.log | tail -3
timeout 600 python bench.py > gpurun_out/bench_default.log 2>&1; echo "bench default rc=$?"; tail -1 gpurun_out/bench_default.log | cut -c1-300
