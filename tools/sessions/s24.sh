#!/bin/bash
# GPU session 24: fast replicated table fill everywhere, pass 2 with run-time tile height / every SM used by small calls,
# programmatic dependent launch for the pass-2 and encode kernels: full suite + one-frame-per-call times + batch benches
set -x
O=gpurun_out
mkdir -p $O
V=$PWD/build/variants
timeout 1500 python -m pytest tests -m gpu -x -q > $O/s24_pytest.log 2>&1; tail -2 $O/s24_pytest.log
BT709_PDL=all timeout 1500 python -m pytest tests -m gpu -x -q -k "not cli and not example" > $O/s24_pytest_pdlall.log 2>&1; tail -2 $O/s24_pytest_pdlall.log
for lib in base2 tree; do
  if [ $lib = tree ]; then unset BT709_B200_LIB; else export BT709_B200_LIB=$V/$lib.so; fi
  echo "== $lib"
  for m in decode rescale twopass encode; do python tools/time_small.py 1920 1080 1 200 3 $m; done
  for m in decode rescale twopass encode; do python tools/time_small.py 1920 1080 2 100 3 $m; done
  for m in rescale twopass encode; do python tools/time_small.py 1920 1080 4 100 3 $m; done
  python tools/time_small.py 1280 720 1 200 3 rescale
  python tools/time_small.py 3840 2160 1 100 3 rescale
  python tools/time_small.py 3840 2160 1 100 3 encode
done > $O/s24_time_small_modes.txt 2>&1
unset BT709_B200_LIB
echo "== tree, BT709_PDL=0" >> $O/s24_time_small_modes.txt
for m in decode rescale twopass encode; do BT709_PDL=0 python tools/time_small.py 1920 1080 1 200 3 $m; done >> $O/s24_time_small_modes.txt 2>&1
for rep in 1 2; do
  for lib in base2 tree; do
    if [ $lib = tree ]; then unset BT709_B200_LIB; else export BT709_B200_LIB=$V/$lib.so; fi
    for wl in c2_1080p c2_rescale c2_twopass c2_fused c5_encode c4_alpha; do
      timeout 300 python bench.py --steps 20 --warmup 5 --no-configs --no-e2e --workload $wl > $O/s24_${wl}_${lib}_$rep.json 2>> $O/s24.err
    done
  done
done
unset BT709_B200_LIB
python - <<'PY'
import json, glob
for f in sorted(glob.glob('gpurun_out/s24_c*.json')):
    try:
        d = json.load(open(f))
        print(f.split('/')[-1], round(d['value'], 1), d['unit'], 'frac', round(d['roofline']['frac'], 4), d['clocks']['sm_mhz'])
    except Exception as e:
        print(f, 'ERR', e)
PY
grep -v "^+" $O/s24_time_small_modes.txt
echo done
