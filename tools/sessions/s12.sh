#!/bin/bash
# GPU session 12: compute-sanitizer over the round-2 kernels (memcheck, racecheck, synccheck), both decode rings
set -x
O=gpurun_out
mkdir -p $O
export PYTORCH_NO_CUDA_MEMORY_CACHING=1
for tool in memcheck racecheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool --error-exitcode 7 python tools/sanitize_small.py > $O/r02_sanitizer_$tool.log 2>&1; echo "rc=$?" >> $O/r02_sanitizer_$tool.log; tail -4 $O/r02_sanitizer_$tool.log
done
BT709_RING=3l timeout 900 compute-sanitizer --tool racecheck --error-exitcode 7 python tools/sanitize_small.py > $O/r02_sanitizer_racecheck_ring3l.log 2>&1; echo "rc=$?" >> $O/r02_sanitizer_racecheck_ring3l.log; tail -3 $O/r02_sanitizer_racecheck_ring3l.log
BT709_RING=2e timeout 900 compute-sanitizer --tool racecheck --error-exitcode 7 python tools/sanitize_small.py > $O/r02_sanitizer_racecheck_ring2e.log 2>&1; echo "rc=$?" >> $O/r02_sanitizer_racecheck_ring2e.log; tail -3 $O/r02_sanitizer_racecheck_ring2e.log
timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_gpu_round2.py -x -q -k "fused_decode or paired or linear_tile" > $O/r02_sanitizer_memcheck_pytest.log 2>&1; echo "rc=$?" >> $O/r02_sanitizer_memcheck_pytest.log; tail -4 $O/r02_sanitizer_memcheck_pytest.log
echo done
