#!/bin/bash
# compute-sanitizer evidence for profiles/: memcheck over the GPU parity suite (the big-batch tests are left out: under the
# sanitizer every kernel runs 10-100x slower) and racecheck over the shared-memory CCL kernel and one whole-path run.
# usage: gpurun --timeout 2400 -- 'bash tools/gpu_sanitize.sh r2'
TAG=${1:-rX}
O=gpurun_out
mkdir -p $O
SMALL='not stress and not c5 and not full_size and not c3 and not c4 and not two_threads and not back_to_back and not 135'
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 --log-file $O/${TAG}_memcheck.log \
    python -m pytest tests -m gpu -x -q -k "$SMALL" > $O/${TAG}_memcheck_pytest.log 2>&1
echo "memcheck rc=$?" >> $O/${TAG}_memcheck_pytest.log
tail -3 $O/${TAG}_memcheck_pytest.log; tail -5 $O/${TAG}_memcheck.log
timeout 900 compute-sanitizer --tool racecheck --racecheck-report all --error-exitcode 9 --log-file $O/${TAG}_racecheck.log \
    python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "bboxcc_matches_opencv_golden or pipeline_ccl_batch or deterministic" > $O/${TAG}_racecheck_pytest.log 2>&1
echo "racecheck rc=$?" >> $O/${TAG}_racecheck_pytest.log
tail -3 $O/${TAG}_racecheck_pytest.log; tail -5 $O/${TAG}_racecheck.log
