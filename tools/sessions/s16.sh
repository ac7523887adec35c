#!/bin/bash
# GPU session 16: decode small calls: LUT fill after the first loads + programmatic dependent launch
set -x
O=gpurun_out
mkdir -p $O
V=$PWD/build/variants
BT709_B200_LIB=$V/dec2.so timeout 900 python -m pytest tests/test_gpu_decode.py tests/test_gpu_round2.py tests/test_gpu_reference_tests.py -x -q > $O/s16_pytest_dec2.log 2>&1; tail -1 $O/s16_pytest_dec2.log
BT709_PDL=all BT709_B200_LIB=$V/dec2.so timeout 900 python -m pytest tests/test_gpu_decode.py tests/test_gpu_round2.py -x -q > $O/s16_pytest_dec2_pdlall.log 2>&1; tail -1 $O/s16_pytest_dec2_pdlall.log
(for rep in 1 2; do
for f in 1 2 4 8; do
  echo "base:";   BT709_B200_LIB=$V/base.so python tools/time_small.py 1920 1080 $f
  echo "dec2 pdl:"; BT709_B200_LIB=$V/dec2.so python tools/time_small.py 1920 1080 $f
  echo "dec2 nopdl:"; BT709_PDL=0 BT709_B200_LIB=$V/dec2.so python tools/time_small.py 1920 1080 $f
done
echo "uhd base:"; BT709_B200_LIB=$V/base.so python tools/time_small.py 3840 2160 1
echo "uhd dec2:"; BT709_B200_LIB=$V/dec2.so python tools/time_small.py 3840 2160 1
echo "720p base:"; BT709_B200_LIB=$V/base.so python tools/time_small.py 1280 720 1
echo "720p dec2:"; BT709_B200_LIB=$V/dec2.so python tools/time_small.py 1280 720 1
done) > $O/s16_time_small.txt 2>&1
for rep in 1 2; do
  for v in base dec2; do
    BT709_B200_LIB=$V/$v.so timeout 300 python bench.py --steps 20 --warmup 5 --no-configs --no-e2e > $O/s16_c2_${v}_$rep.json 2>> $O/s16.err
    BT709_B200_LIB=$V/$v.so timeout 300 python bench.py --steps 20 --warmup 5 --no-configs --no-e2e --workload c4_alpha > $O/s16_c4_${v}_$rep.json 2>> $O/s16.err
  done
  BT709_PDL=all BT709_B200_LIB=$V/dec2.so timeout 300 python bench.py --steps 20 --warmup 5 --no-configs --no-e2e > $O/s16_c2_dec2pdlall_$rep.json 2>> $O/s16.err
done
python - <<'PY'
import json, glob
for f in sorted(glob.glob('gpurun_out/s16_c*.json')):
    try:
        d = json.load(open(f))
        print(f.split('/')[-1], round(d['value'], 1), d['unit'], 'frac', round(d['roofline']['frac'], 4), 'sustained', round(d.get('sustained', {}).get('value', 0), 1), d['clocks']['sm_mhz'])
    except Exception as e:
        print(f, 'ERR', e)
PY
cat $O/s16_time_small.txt | grep -v "^+"
echo done
