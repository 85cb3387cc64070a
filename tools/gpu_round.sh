#!/bin/bash
# One GPU-box visit: parity tests, bench (both arms), ncu launch list and a full capture of one step.
# usage: gpurun --timeout 1500 -- 'bash tools/gpu_round.sh r1d'
TAG=${1:-rX}
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/${TAG}_smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > $O/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" >> $O/${TAG}_pytest.log
tail -3 $O/${TAG}_pytest.log
timeout 600 python bench.py > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err; echo "bench rc=$?"
tail -c 600 $O/${TAG}_bench.json
if [ -z "$SKIP_REF" ]; then timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $O/${TAG}_bench_ref.json 2>> $O/${TAG}_bench.err; fi
timeout 300 python tools/layer_timing.py > $O/${TAG}_layer_timing.txt 2>&1
if [ -z "$SKIP_NCU" ]; then
STEPS=2 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/${TAG}_launches.csv python tools/ncu_step.py > $O/${TAG}_ncu1.log 2>&1
STEPS=1 timeout 900 ncu --set full --clock-control none --import-source on -c 40 -f -o $O/${TAG}_full python tools/ncu_step.py > $O/${TAG}_ncu2.log 2>&1
ncu -i $O/${TAG}_full.ncu-rep --page raw --csv > $O/${TAG}_ncu_full_raw.csv 2>/dev/null
python tools/ncu_summary.py < $O/${TAG}_ncu_full_raw.csv > $O/${TAG}_ncu_summary.json 2>/dev/null
fi
echo done
