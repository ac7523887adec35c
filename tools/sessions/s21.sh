#!/bin/bash
# GPU session 21: in-tree build after the decode / pass-2 changes: full suite, smoke, default bench, sustained runs, small calls
set -x
O=gpurun_out
mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -x -q > $O/s21_pytest.log 2>&1; tail -2 $O/s21_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > $O/s21_smoke.log 2>&1; tail -1 $O/s21_smoke.log
timeout 900 python bench.py --steps 20 --warmup 5 > $O/s21_bench_default.json 2>> $O/s21.err
for n in 100 1000; do
  timeout 600 python bench.py --steps $n --warmup 5 --no-configs --no-e2e > $O/s21_bench_c2_${n}.json 2>> $O/s21.err
  BT709_RING=3l timeout 600 python bench.py --steps $n --warmup 5 --no-configs --no-e2e > $O/s21_bench_c2_${n}_3l.json 2>> $O/s21.err
done
timeout 600 python bench.py --steps 20 --warmup 5 --no-configs --no-e2e --gamma linear > $O/s21_bench_c2_linear.json 2>> $O/s21.err
(python tools/time_small.py 1920 1080 1; python tools/time_small.py 1920 1080 1 200 2; python tools/time_small.py 1920 1080 1 200 1; python tools/time_small.py 1920 1080 4; python tools/time_small.py 1920 1080 8) > $O/s21_time_small.txt 2>&1
python - <<'PY'
import json, glob
for f in sorted(glob.glob('gpurun_out/s21_bench_c2*.json')):
    try:
        d = json.load(open(f))
        print(f.split('/')[-1], round(d['value'], 1), d['unit'], 'frac', round(d['roofline']['frac'], 4), 'sustained', round((d.get('sustained') or {}).get('value', 0), 1), d['clocks'])
    except Exception as e:
        print(f, 'ERR', e)
d = json.load(open('gpurun_out/s21_bench_default.json'))
print('default', round(d['value'], 1), d['roofline']['frac'], 'sustained', d['sustained']['value'], d['sustained']['roofline']['frac'], 'e2e', d['e2e']['value'], d['e2e'].get('copy_only_value'))
for c in d['configs']:
    print(c['workload'][:40], round(c['value'], 1), round(c['roofline']['frac'], 4), 'sustained', round(c['sustained']['value'], 1), round(c['sustained']['roofline']['frac'], 4))
PY
grep -v "^+" $O/s21_time_small.txt
echo done
