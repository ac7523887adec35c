#!/bin/bash
# GPU session 15: pass-2 walker with the inner-tile path (immediate tap offsets, no live predicate), tile heights 64/96/128
set -x
O=gpurun_out
mkdir -p $O
for v in inner64 inner96 inner128; do
  BT709_B200_LIB=$PWD/build/variants/$v.so timeout 900 python -m pytest tests/test_gpu_rescale_encode.py tests/test_gpu_round2.py tests/test_gpu_full_configs.py -x -q -k "rescale or pair or fused or two_pass or twopass" > $O/s15_pytest_$v.log 2>&1; tail -1 $O/s15_pytest_$v.log
done
for rep in 1 2; do
  for v in base inner64 inner96 inner128; do
    for wl in c2_rescale c2_twopass c2_fused; do
      BT709_B200_LIB=$PWD/build/variants/$v.so timeout 300 python bench.py --steps 20 --warmup 5 --no-configs --no-e2e --workload $wl > $O/s15_${wl}_${v}_$rep.json 2>> $O/s15.err
    done
  done
done
python - <<'PY'
import json, glob
for f in sorted(glob.glob('gpurun_out/s15_c2_*.json')):
    try:
        d = json.load(open(f))
        print(f.split('/')[-1], round(d['value'], 1), d['unit'], 'frac', round(d['roofline']['frac'], 4), 'sustained', round(d.get('sustained', {}).get('value', 0), 1), d['clocks']['sm_mhz'])
    except Exception as e:
        print(f, 'ERR', e)
PY
echo done
