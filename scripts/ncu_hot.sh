#!/bin/bash
# usage: scripts/ncu_hot.sh <report.ncu-rep> <sim_steps_per_launch> -> /tmp/<name>_hot.txt + summary on stdout
rep=$1; steps=${2:-2097152000}; name=$(basename $rep .ncu-rep)
ncu -i $rep --page raw --csv > /tmp/${name}_raw.csv 2>/dev/null
ncu -i $rep --page source --csv > /tmp/${name}_src.csv 2>/dev/null
python scripts/ncu_summary.py /tmp/${name}_raw.csv $steps | python -c "
import json,sys
for d in json.load(sys.stdin):
    for k,v in d.items():
        if k in ('kernel','gpu__time_duration.sum','sm__throughput.avg.pct_of_peak_sustained_elapsed','smsp__issue_active.avg.pct_of_peak_sustained_active','smsp__thread_inst_executed_per_inst_executed.ratio','sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active','launch__registers_per_thread','smsp__warps_eligible.avg.per_cycle_active','derived','l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum'):
            print(k, v if not isinstance(v,dict) or 'value' not in v else (v['value'], v['unit']))"
python - <<EOF > /tmp/${name}_hot.txt
import csv
rows=list(csv.reader(open('/tmp/${name}_src.csv')))
hdr=rows[1]; ia=hdr.index("Instructions Executed"); isrc=hdr.index("Source"); isamp=hdr.index("# Samples"); it=hdr.index("Avg. Threads Executed")
W=$steps/32
tot=0
for i,r in enumerate(rows[2:]):
    try: n=int(r[ia])
    except: continue
    tot+=n
    print(f"{i:5d} {n/W:7.3f} {float(r[it]):5.1f} {int(r[isamp]):6d}  {r[isrc].strip()}")
print("TOTAL per step", tot/W)
EOF
tail -1 /tmp/${name}_hot.txt
