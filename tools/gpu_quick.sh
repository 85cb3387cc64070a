#!/bin/bash
# Short GPU-box visit during development: parity tests, then the per-layer timing table with its variants.
# usage: gpurun --timeout 900 -- 'bash tools/gpu_quick.sh tag'
TAG=${1:-q}
O=gpurun_out
mkdir -p $O
timeout 600 python -m pytest tests -m gpu -x -q > $O/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" >> $O/${TAG}_pytest.log
tail -5 $O/${TAG}_pytest.log
timeout 300 python tools/layer_timing.py > $O/${TAG}_layer_timing.txt 2>&1
cat $O/${TAG}_layer_timing.txt
