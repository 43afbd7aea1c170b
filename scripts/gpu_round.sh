#!/bin/bash
# One GPU session: parity tests, bench lines, ncu launch list + one full capture of the step kernel.
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/pytest.log
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap --format=csv -lms 200 > gpurun_out/clocks.csv &
SMI=$!
timeout 600 python bench.py > gpurun_out/bench_pp.log 2>&1; echo "bench rc=$?"; tail -1 gpurun_out/bench_pp.log
timeout 600 python bench.py --workload stanley_lqg --no-cpu-baseline > gpurun_out/bench_stanley.log 2>&1; echo "bench stanley rc=$?"; tail -1 gpurun_out/bench_stanley.log
timeout 600 python bench.py --workload mpc > gpurun_out/bench_mpc.log 2>&1; echo "bench mpc rc=$?"; tail -1 gpurun_out/bench_mpc.log
for w in skidpad_mc pp_replan; do timeout 600 python bench.py --workload $w --no-cpu-baseline > gpurun_out/bench_$w.log 2>&1; echo "bench $w rc=$?"; tail -1 gpurun_out/bench_$w.log; done
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.log 2>&1; echo "bench ref rc=$?"; tail -1 gpurun_out/bench_ref.log
kill $SMI
CMD="python bench.py --steps 2 --warmup 1 --no-cpu-baseline"
$CMD > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
echo "ncu launches rc=$?"
$CMD > gpurun_out/plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:rollout_kernel -s 1 -c 1 -o gpurun_out/prof_pp -f $CMD > gpurun_out/ncu_full.log 2>&1
echo "ncu full rc=$?"; tail -3 gpurun_out/ncu_full.log
CMD2="python bench.py --workload stanley_lqg --steps 2 --warmup 1 --no-cpu-baseline"
$CMD2 > gpurun_out/plain3.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:rollout_kernel -s 1 -c 1 -o gpurun_out/prof_stanley -f $CMD2 > gpurun_out/ncu_full2.log 2>&1
echo "ncu full2 rc=$?"
ls -la gpurun_out
