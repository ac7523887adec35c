#!/bin/bash
# GPU session 36: strided batches on k_decode_lin (frame wrap)
set -x
O=gpurun_out
mkdir -p $O
timeout 1500 python -m pytest tests/test_gpu_decode.py tests/test_gpu_round2.py tests/test_gpu_reference_tests.py tests/test_gpu_full_configs.py -x -q > $O/s36_pytest.log 2>&1; tail -3 $O/s36_pytest.log
export PYTORCH_NO_CUDA_MEMORY_CACHING=1
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_gpu_round2.py -x -q -k "strided_batches" > $O/s36_memcheck.log 2>&1; echo "rc=$?" >> $O/s36_memcheck.log; tail -4 $O/s36_memcheck.log
unset PYTORCH_NO_CUDA_MEMORY_CACHING
for mode in new old; do
  if [ $mode = old ]; then export BT709_STRIDED_LIN=0; else unset BT709_STRIDED_LIN; fi
  echo "== $mode"
  for s in "1920 1080 400" "3840 2160 100" "1280 720 800" "720 480 2000" "1920 1080 400 2" "2048 1536 200 1"; do python tools/time_decode_nv12buf.py $s; done
done > $O/s36_time_strided.txt 2>&1
unset BT709_STRIDED_LIN
grep -v "^+" $O/s36_time_strided.txt
for rep in 1 2; do timeout 300 python bench.py --steps 20 --warmup 5 --no-configs --no-e2e --no-live-traffic | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('c2', d['value'], d['roofline']['frac'])"; done
echo done
