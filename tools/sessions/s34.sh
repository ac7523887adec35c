#!/bin/bash
# GPU session 34: ragged widths on k_decode_lin (partial last tile column)
set -x
O=gpurun_out
mkdir -p $O
timeout 1200 python -m pytest tests/test_gpu_decode.py tests/test_gpu_round2.py tests/test_gpu_reference_tests.py tests/test_gpu_full_configs.py -x -q > $O/s34_pytest.log 2>&1; tail -3 $O/s34_pytest.log
export PYTORCH_NO_CUDA_MEMORY_CACHING=1
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_gpu_round2.py -x -q -k "ragged or linear_tile" > $O/s34_memcheck.log 2>&1; echo "rc=$?" >> $O/s34_memcheck.log; tail -4 $O/s34_memcheck.log
unset PYTORCH_NO_CUDA_MEMORY_CACHING
for mode in new old; do
  if [ $mode = old ]; then export BT709_NO_RAGGED_LIN=1; else unset BT709_NO_RAGGED_LIN; fi
  echo "== $mode"
  for s in "720 480 2000" "960 540 1500" "1440 1080 500" "352 288 6000" "1920 1080 400" "2000 1000 400"; do python tools/time_decode.py $s; done
  python tools/time_small.py 1440 1080 1 200 3 decode 2>/dev/null || true
done > $O/s34_time_ragged.txt 2>&1
grep -v "^+" $O/s34_time_ragged.txt
echo done
