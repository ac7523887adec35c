#!/bin/bash
# GPU session 19: pass 2 with 9 tap conversions per pair and row (float selects) against 12
set -x
O=gpurun_out
mkdir -p $O
V=$PWD/build/variants
for v in taps9; do
  BT709_B200_LIB=$V/$v.so timeout 900 python -m pytest tests/test_gpu_rescale_encode.py tests/test_gpu_round2.py tests/test_gpu_full_configs.py -x -q -k "rescale or pair or fused or two_pass or twopass" > $O/s19_pytest_$v.log 2>&1; tail -1 $O/s19_pytest_$v.log
done
for rep in 1 2; do
  for v in base order64 taps9 taps9_96; do
    for wl in c2_rescale c2_twopass c2_fused; do
      BT709_B200_LIB=$V/$v.so timeout 300 python bench.py --steps 20 --warmup 5 --no-configs --no-e2e --workload $wl > $O/s19_${wl}_${v}_$rep.json 2>> $O/s19.err
    done
  done
done
for v in base taps9; do
  BT709_B200_LIB=$V/$v.so python tools/time_rescale.py 1920 1080 2560 1440 100 smooth opaque
  BT709_B200_LIB=$V/$v.so python tools/time_rescale.py 1920 1080 2560 1440 100 noise
  BT709_B200_LIB=$V/$v.so python tools/time_rescale.py 1280 720 1920 1080 100 noise opaque
  BT709_B200_LIB=$V/$v.so python tools/time_rescale.py 1920 1080 3840 2160 50 noise opaque
done > $O/s19_time_rescale.txt 2>&1
python - <<'PY'
import json, glob
for f in sorted(glob.glob('gpurun_out/s19_c*.json')):
    try:
        d = json.load(open(f))
        print(f.split('/')[-1], round(d['value'], 1), d['unit'], 'frac', round(d['roofline']['frac'], 4), d['clocks']['sm_mhz'])
    except Exception as e:
        print(f, 'ERR', e)
PY
grep -v "^+" $O/s19_time_rescale.txt
echo done
