#!/bin/bash
# GPU session 22: pass 2 occupancy / prefetch depth sweep on the new walker
set -x
O=gpurun_out
mkdir -p $O
V=$PWD/build/variants
for rep in 1 2; do
  for v in cur w16pd4 w16pd2 w20pd2 w20pd3 w24pd2 w12pd4; do
    for wl in c2_rescale c2_twopass; do
      BT709_B200_LIB=$V/$v.so timeout 300 python bench.py --steps 20 --warmup 5 --no-configs --no-e2e --workload $wl > $O/s22_${wl}_${v}_$rep.json 2>> $O/s22.err
    done
  done
done
python - <<'PY'
import json, glob
for f in sorted(glob.glob('gpurun_out/s22_c*.json')):
    try:
        d = json.load(open(f))
        print(f.split('/')[-1], round(d['value'], 1), d['unit'], 'frac', round(d['roofline']['frac'], 4), d['clocks']['sm_mhz'])
    except Exception as e:
        print(f, 'ERR', e)
PY
echo done
