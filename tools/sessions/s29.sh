#!/bin/bash
# GPU session 29: bench.py with the live DRAM-traffic probe
set -x
O=gpurun_out
mkdir -p $O
( time timeout 900 python bench.py --steps 20 --warmup 5 > $O/s29_bench_default.json 2>> $O/s29.err ) 2>&1 | grep real
python - <<'PY'
import json
d=json.load(open('gpurun_out/s29_bench_default.json'))
print(d['value'], d['roofline'])
PY
for wl in c2_rescale c2_twopass c2_fused c5_encode c4_alpha c3_uhd; do
  ( time timeout 300 python bench.py --steps 20 --warmup 5 --no-configs --no-e2e --workload $wl > $O/s29_$wl.json 2>> $O/s29.err ) 2>&1 | grep real
  python -c "
import json
d=json.load(open('gpurun_out/s29_$wl.json'))
print('$wl', round(d['value'],1), d['roofline']['traffic'], d['roofline']['algorithmic_bytes_per_launch'], d['roofline']['traffic_source'][:90])"
done
tail -5 $O/s29.err
echo done
