#!/bin/bash
# GPU session 20: encode with the first loads ahead of the table fill; full suite on the combined build
set -x
O=gpurun_out
mkdir -p $O
V=$PWD/build/variants
BT709_B200_LIB=$V/all2.so timeout 1500 python -m pytest tests -m gpu -x -q > $O/s20_pytest_all2.log 2>&1; tail -1 $O/s20_pytest_all2.log
for rep in 1 2 3; do
  for v in base all2; do
    BT709_B200_LIB=$V/$v.so timeout 300 python bench.py --steps 20 --warmup 5 --no-configs --no-e2e --workload c5_encode > $O/s20_c5_${v}_$rep.json 2>> $O/s20.err
  done
done
for v in base all2; do
  (BT709_B200_LIB=$V/$v.so python tools/time_decode.py 2048 1536 200 0 encode 1 0; BT709_B200_LIB=$V/$v.so python tools/time_decode.py 2048 1536 200 0 encode 1 1; BT709_B200_LIB=$V/$v.so python tools/time_decode.py 1920 1080 300 0 encode 1 3;  BT709_B200_LIB=$V/$v.so python tools/time_decode.py 3840 2160 100 0 encode 1 0) > $O/s20_encode_pairs_$v.txt 2>&1
done
python - <<'PY'
import json, glob
for f in sorted(glob.glob('gpurun_out/s20_c*.json')):
    try:
        d = json.load(open(f))
        print(f.split('/')[-1], round(d['value'], 1), d['unit'], 'frac', round(d['roofline']['frac'], 4), 'sustained', round(d.get('sustained', {}).get('value', 0), 1), d['clocks']['sm_mhz'])
    except Exception as e:
        print(f, 'ERR', e)
PY
tail -n 5 $O/s20_encode_pairs_base.txt $O/s20_encode_pairs_all2.txt
echo done
