#!/bin/bash
# One GPU-box visit that regenerates everything under profiles/ for a round: parity tests, the bench lines of the three
# configs and the reference arm, launch lists + ncu summaries (tools/gpu_profile.sh), sanitizer logs (tools/gpu_sanitize.sh),
# CCL timings.  usage: gpurun --timeout 2400 -- 'bash tools/gpu_round.sh r3'      (then copy gpurun_out/<tag>_* to profiles/)
TAG=${1:-rX}
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/${TAG}_smi.txt 2>&1
timeout 1200 python -m pytest tests -m gpu -x -q > $O/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" >> $O/${TAG}_pytest.log
tail -3 $O/${TAG}_pytest.log
for C in c2 c3 c4; do
  timeout 600 python bench.py --config $C > $O/${TAG}_bench_$C.json 2> $O/${TAG}_bench_$C.err; echo "bench $C rc=$?"
done
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > $O/${TAG}_bench_ref.json 2>> $O/${TAG}_bench_c2.err
timeout 300 python tools/layer_timing.py > $O/${TAG}_layer_timing.txt 2>&1
timeout 300 python tools/ccl_timing.py > $O/${TAG}_ccl_timing.txt 2>&1
if [ -z "$SKIP_NCU" ]; then bash tools/gpu_profile.sh $TAG "c2 c3 c4"; fi
if [ -z "$SKIP_SANITIZER" ]; then bash tools/gpu_sanitize.sh $TAG; fi
echo done
