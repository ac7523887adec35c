#!/bin/bash
# GPU session 26: new small-call tests; pass-2 prefetch depth / warps on the round-robin tile order
set -x
O=gpurun_out
mkdir -p $O
V=$PWD/build/variants
timeout 900 python -m pytest tests/test_gpu_round2.py -x -q -k "tile_height or back_to_back" > $O/s26_pytest_new.log 2>&1; tail -2 $O/s26_pytest_new.log
BT709_PDL=0 timeout 900 python -m pytest tests/test_gpu_round2.py -x -q -k "back_to_back" > $O/s26_pytest_new_nopdl.log 2>&1; tail -1 $O/s26_pytest_new_nopdl.log
for rep in 1 2; do
  for lib in tree w16pd2 w20pd2 w16pd4; do
    if [ $lib = tree ]; then unset BT709_B200_LIB; else export BT709_B200_LIB=$V/$lib.so; fi
    for wl in c2_rescale c2_twopass; do
      timeout 300 python bench.py --steps 20 --warmup 5 --no-configs --no-e2e --workload $wl > $O/s26_${wl}_${lib}_$rep.json 2>> $O/s26.err
    done
  done
done
for lib in tree w16pd2 w20pd2; do
  if [ $lib = tree ]; then unset BT709_B200_LIB; else export BT709_B200_LIB=$V/$lib.so; fi
  echo "== $lib"
  python tools/time_rescale.py 1920 1080 2560 1440 100 smooth opaque
  python tools/time_rescale.py 1920 1080 2560 1440 100 noise
  python tools/time_rescale.py 1280 720 1920 1080 100 noise opaque
  python tools/time_rescale.py 1920 1080 3840 2160 50 noise opaque
  python tools/time_rescale.py 2048 1536 2732 2048 60 noise opaque
  python tools/time_small.py 1920 1080 1 200 3 rescale
done > $O/s26_time_rescale.txt 2>&1
unset BT709_B200_LIB
python - <<'PY'
import json, glob
for f in sorted(glob.glob('gpurun_out/s26_c*.json')):
    try:
        d = json.load(open(f))
        print(f.split('/')[-1], round(d['value'], 1), d['unit'], 'frac', round(d['roofline']['frac'], 4), d['clocks']['sm_mhz'])
    except Exception as e:
        print(f, 'ERR', e)
PY
grep -v "^+" $O/s26_time_rescale.txt
echo done
