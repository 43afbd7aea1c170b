#!/bin/bash
# One GPU session (1 x B200): parity tests, the default bench line (config 2 + every other BASELINE config as secondary),
# the reference arm, then ncu: a launch list of the default command and one `--set full` capture (with source counters)
# of each kernel DESIGN.md quotes.  Everything lands in gpurun_out/; scripts/make_profiles.sh turns it into profiles/<dir>.
#   usage: gpurun --timeout 2400 -- bash scripts/gpu_round.sh [notests]
mkdir -p gpurun_out
python -c "import bgr_pathplanning_control_b200 as b; print(b.build_id())" > gpurun_out/build_id.txt; cat gpurun_out/build_id.txt
if [ "$1" != notests ]; then
  timeout 1500 python -m pytest tests -m gpu -q --timeout 900 --tb=short -rP > gpurun_out/pytest_full.log 2>&1; echo "pytest rc=$?"
  grep -E "passed|failed|error" gpurun_out/pytest_full.log | tail -3; grep -E "^\[(parity fractions|mpc sub-optimality)" gpurun_out/pytest_full.log
fi
# (no second nvidia-smi poller next to bench.py's own clock sampler: with one running at -lms 200 beside it the headline
#  loop of one round averaged 16.5 ms per step against 14.5 ms without -- bench.py's own samples are the record)
timeout 900 python bench.py > gpurun_out/bench_default.log 2>&1; echo "bench default rc=$?"; tail -1 gpurun_out/bench_default.log | cut -c1-300
for w in pp_replan; do timeout 600 python bench.py --workload $w --no-cpu-baseline > gpurun_out/bench_$w.log 2>&1; echo "bench $w rc=$?"; tail -1 gpurun_out/bench_$w.log | cut -c1-200; done
# config 5 on the rows other ranks of an 8-GPU run would get (the step time of a multi-GPU run is the slowest rank's:
# which rows hold the long / lost episodes decides it)
for r in ${MC_ROW_RANKS:-3 6}; do
  BGR_MC_ROW_RANK=$r timeout 300 python bench.py --workload skidpad_mc --steps 5 --warmup 2 --no-cpu-baseline --no-secondary \
    > gpurun_out/bench_mc_rows$r.log 2>&1; echo "bench skidpad_mc rows of rank $r rc=$?"; tail -1 gpurun_out/bench_mc_rows$r.log | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('   value %.4e  ms/step %.3f  e2e %.4e (sync %.4e)' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['one_call_at_a_time']['value']))"
done
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.log 2>&1; echo "bench ref rc=$?"; tail -1 gpurun_out/bench_ref.log | cut -c1-200
CMD="python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-secondary"
$CMD > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
echo "ncu launches rc=$?"
cap() {  # <report name> <kernel regex> <skip> <bench args...>
  name=$1; regex=$2; skip=$3; shift 3
  python bench.py "$@" --steps 2 --warmup 1 --no-cpu-baseline --no-secondary > gpurun_out/plain_$name.log 2>&1 && \
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$regex -s $skip -c 1 -o gpurun_out/prof_$name -f \
    python bench.py "$@" --steps 2 --warmup 1 --no-cpu-baseline --no-secondary > gpurun_out/ncu_$name.log 2>&1
  echo "ncu $name rc=$?"
}
cap pp rollout_kernel 1 --workload pp_sweep
cap stanley rollout_kernel 1 --workload stanley_lqg
cap replan rollout_kernel 1 --workload pp_replan
# config 5: the pure-pursuit kind is the launch that fills the GPU.  Kinds launch in the order stanley+lqg, stanley+pid,
# pp, all instantiations of rollout_kernel_f32 (ncu matches base names): launch #6 is the pp kind of the first timed step
cap mc rollout_kernel_f32 5 --workload skidpad_mc
cap mc_stlqg rollout_kernel_f32 3 --workload skidpad_mc     # launch #4: stanley+lqg+dynamic on the dynamic queue
cap mpc mpc_factored_kernel 1 --workload mpc
ls -la gpurun_out | head -60
