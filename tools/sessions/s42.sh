#!/bin/bash
# GPU session 42: encode with the linear tile order for rows of fewer than 32 tiles that do not divide 32
set -x
O=gpurun_out
mkdir -p $O
timeout 1200 python -m pytest tests -m gpu -x -q -k "encode or cli or example or reference_tests or full_configs" > $O/s42_pytest.log 2>&1; tail -2 $O/s42_pytest.log
for mode in new old; do
  if [ $mode = old ]; then export BT709_NO_ENC_LINEAR=1; else unset BT709_NO_ENC_LINEAR; fi
  echo "== $mode"
  for w in 1280 1920 3840 720 1440 2048; do f=$((800000000 / (w * 1080))); python tools/time_decode.py $w 1080 $f 0 encode 1 0; done
  python tools/time_decode.py 1920 1080 385 0 encode 1 1
  python tools/time_decode.py 1920 1080 385 0 encode 2 2
  python tools/time_small.py 1920 1080 1 200 3 encode
done > $O/s42_time_encode.txt 2>&1
grep -v "^+" $O/s42_time_encode.txt
echo done
