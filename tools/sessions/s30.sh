#!/bin/bash
# GPU session 30: ring discipline re-check for the alpha variant and the Linear mode after the prologue change; profiler-env check
set -x
O=gpurun_out
mkdir -p $O
ncu --metrics gpu__time_duration.sum -c 1 python -c "import os; print('ENV', sorted(k for k in os.environ if any(t in k for t in ('INJECT','NSIGHT','NCU','PROFIL'))))" 2>&1 | grep ENV
for rep in 1 2; do
  for ring in 2e 3l; do
    BT709_RING=$ring timeout 300 python bench.py --steps 20 --warmup 5 --no-configs --no-e2e --no-live-traffic --workload c4_alpha > $O/s30_c4_${ring}_$rep.json 2>> $O/s30.err
    BT709_RING=$ring timeout 300 python bench.py --steps 20 --warmup 5 --no-configs --no-e2e --no-live-traffic --gamma linear > $O/s30_lin_${ring}_$rep.json 2>> $O/s30.err
    BT709_RING=$ring timeout 300 python bench.py --steps 20 --warmup 5 --no-configs --no-e2e --no-live-traffic --workload c3_uhd > $O/s30_uhd_${ring}_$rep.json 2>> $O/s30.err
    BT709_RING=$ring timeout 300 python bench.py --steps 300 --warmup 5 --no-configs --no-e2e --no-live-traffic --workload c4_alpha > $O/s30_c4s_${ring}_$rep.json 2>> $O/s30.err
  done
done
python - <<'PY'
import json, glob
for f in sorted(glob.glob('gpurun_out/s30_*.json')):
    try:
        d = json.load(open(f))
        print(f.split('/')[-1], round(d['value'], 1), 'frac', round(d['roofline']['frac'], 4), d['clocks']['sm_mhz'])
    except Exception as e:
        print(f, 'ERR', e)
PY
echo done
