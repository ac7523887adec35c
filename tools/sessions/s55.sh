#!/bin/bash
# GPU session 55: encode with a conflict-free chroma LUT (biased log cells / uniform cells, one replica per lane)
set -x
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q -k "encode or cli or example or reference_tests or full_configs or average or histogram or back_to_back or graph" > $O/s55_pytest.log 2>&1; tail -2 $O/s55_pytest.log
for mode in 2 0; do
  echo "== BT709_ENC_OMODE=$mode (2 = one replica per lane where available, 0 = 8 replicas)"
  for s in "1920 1080 385" "2048 1536 200" "3840 2160 96" "720 480 2000"; do BT709_ENC_OMODE=$mode python tools/time_decode.py $s 0 encode 1 0; done
  BT709_ENC_OMODE=$mode python tools/time_decode.py 1920 1080 385 0 encode 1 1
  BT709_ENC_OMODE=$mode python tools/time_decode.py 1920 1080 385 0 encode 2 2
  BT709_ENC_OMODE=$mode python tools/time_small.py 1920 1080 1 200 3 encode
done > $O/s55_time_encode.txt 2>&1
BT709_ENC_OMODE=1 python tools/time_decode.py 1920 1080 385 0 encode 2 2 >> $O/s55_time_encode.txt 2>&1
grep -v "^+" $O/s55_time_encode.txt
timeout 300 python bench.py --steps 20 --warmup 5 --no-configs --no-e2e --workload c5_encode > $O/s55_c5.json 2>> $O/s55.err
python -c "
import json; d=json.load(open('gpurun_out/s55_c5.json')); print('c5', round(d['value'],1), round(d['roofline']['frac'],4), d['roofline']['traffic'], d['clocks'])"
echo done
