#!/bin/bash
# GPU session 17: why is the inner-tile walker slower?  timing of diagnostic builds + ncu captures of base / inner64
set -x
O=gpurun_out
mkdir -p $O
V=$PWD/build/variants
for rep in 1 2; do
  for v in base inner64 noinner forceinner; do
    BT709_B200_LIB=$V/$v.so timeout 300 python bench.py --steps 20 --warmup 5 --no-configs --no-e2e --workload c2_rescale > $O/s17_c2_rescale_${v}_$rep.json 2>> $O/s17.err
  done
done
for v in base inner64; do
  BT709_B200_LIB=$V/$v.so timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_rescale_pair -c 1 -o $O/s17_ncu_$v -f python tools/time_rescale.py 1920 1080 2560 1440 100 > $O/s17_ncu_$v.log 2>&1
done
python - <<'PY'
import json, glob
for f in sorted(glob.glob('gpurun_out/s17_c*.json')):
    try:
        d = json.load(open(f))
        print(f.split('/')[-1], round(d['value'], 1), d['unit'], 'frac', round(d['roofline']['frac'], 4), d['clocks']['sm_mhz'])
    except Exception as e:
        print(f, 'ERR', e)
PY
echo done
