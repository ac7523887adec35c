#!/bin/bash
# GPU session 47: k_rescale_pair with the row info inside the alignment gap (191 KB of shared memory: 32 KB of L1 left)
set -x
O=gpurun_out
mkdir -p $O
V=$PWD/build/variants
timeout 900 python -m pytest tests/test_gpu_rescale_encode.py tests/test_gpu_round2.py tests/test_gpu_full_configs.py -x -q -k "rescale or pair or fused or two_pass or twopass or tile_height or back_to_back or graph" > $O/s47_pytest.log 2>&1; tail -1 $O/s47_pytest.log
for rep in 1 2; do
  for lib in prev tree; do
    if [ $lib = tree ]; then unset BT709_B200_LIB; else export BT709_B200_LIB=$V/$lib.so; fi
    for wl in c2_rescale c2_twopass c2_fused; do
      timeout 300 python bench.py --steps 20 --warmup 5 --no-configs --no-e2e --no-live-traffic --workload $wl > $O/s47_${wl}_${lib}_$rep.json 2>> $O/s47.err
    done
  done
done
unset BT709_B200_LIB
python - <<'PY'
import json, glob
for f in sorted(glob.glob('gpurun_out/s47_c*.json')):
    try:
        d = json.load(open(f))
        print(f.split('/')[-1], round(d['value'], 1), 'frac', round(d['roofline']['frac'], 4), d['clocks']['sm_mhz'])
    except Exception as e:
        print(f, 'ERR', e)
PY
python tools/time_rescale.py 1920 1080 2560 1440 100 smooth opaque; python tools/time_rescale.py 1920 1080 2560 1440 100 noise; python tools/time_small.py 1920 1080 1 200 3 rescale
echo done
