#!/bin/bash
# One GPU-box visit for the committed profiles: per config the launch list (gpu__time_duration of every kernel of two steps)
# and one `ncu --set full` capture of one step, condensed by tools/ncu_summary.py.
# usage: gpurun --timeout 2400 -- 'bash tools/gpu_profile.sh r2 "c2 c3 c4"'
TAG=${1:-rX}
O=gpurun_out
mkdir -p $O
for C in ${2:-c2}; do
  CONFIG=$C STEPS=2 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/${TAG}_${C}_launches.csv python tools/ncu_step.py > $O/${TAG}_${C}_ncu1.log 2>&1
  CONFIG=$C STEPS=1 timeout 900 ncu --set full --clock-control none -c 40 -f -o $O/${TAG}_${C}_full python tools/ncu_step.py > $O/${TAG}_${C}_ncu2.log 2>&1
  ncu -i $O/${TAG}_${C}_full.ncu-rep --page raw --csv > $O/${TAG}_${C}_ncu_full_raw.csv 2>/dev/null
  python tools/ncu_summary.py < $O/${TAG}_${C}_ncu_full_raw.csv > $O/${TAG}_${C}_ncu_summary.json 2>/dev/null
  rm -f $O/${TAG}_${C}_ncu_full_raw.csv
  if [ -z "$KEEP_REP" ]; then rm -f $O/${TAG}_${C}_full.ncu-rep; fi     # gpurun brings back at most 64 MiB
  tail -2 $O/${TAG}_${C}_ncu2.log
done
echo done
