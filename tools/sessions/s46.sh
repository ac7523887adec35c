#!/bin/bash
# GPU session 46: final tree of round 2: full suite, smoke, default bench line, reference arm, sustained runs, sanitizer
set -x
O=gpurun_out
mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -x -q > $O/s46_pytest.log 2>&1; tail -2 $O/s46_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > $O/s46_smoke.log 2>&1; tail -1 $O/s46_smoke.log
( time timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > $O/r02_bench_default.json 2>> $O/s46.err ) 2>&1 | grep real
timeout 300 python bench.py --impl reference --gpus 1 --steps 3 --warmup 1 > $O/r02_bench_reference.json 2>> $O/s46.err
timeout 600 python bench.py --steps 1000 --warmup 5 --no-configs --no-e2e > $O/r02_bench_c2_1080p_1000steps.json 2>> $O/s46.err
BT709_RING=3l timeout 600 python bench.py --steps 1000 --warmup 5 --no-configs --no-e2e > $O/r02_bench_c2_1080p_1000steps_ring3l.json 2>> $O/s46.err
timeout 300 python bench.py --steps 20 --warmup 5 --no-configs --no-e2e --workload c5_encode > $O/s46_c5.json 2>> $O/s46.err
bash tools/sessions/s12.sh > $O/s46_sanitizer.log 2>&1; grep -E "SUMMARY|rc=|passed|failed|sanitize_small" $O/r02_sanitizer_*.log | head -30
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02_bench_default.json'))
print('default', round(d['value'],1), round(d['roofline']['frac'],4), 'traffic/alg', round(d['roofline']['traffic']/d['roofline']['algorithmic_bytes_per_launch'],4), 'sustained', round(d['sustained']['value'],1), round(d['sustained']['roofline']['frac'],4), 'e2e', d['e2e']['value'], d['e2e'].get('copy_only_value'), 'cpu', d['cpu_baseline']['value'])
for c in d['configs']:
    print(c['workload'][:36], round(c['value'], 1), round(c['roofline']['frac'], 4), 'sust', round(c['sustained']['value'], 1), round(c['sustained']['roofline']['frac'], 4))
for f in ['r02_bench_c2_1080p_1000steps','r02_bench_c2_1080p_1000steps_ring3l','r02_bench_reference','s46_c5']:
    d=json.load(open(f'gpurun_out/{f}.json')); print(f, round(d['value'],3), (d.get('roofline') or {}).get('frac'), (d.get('clocks') or {}).get('sm_mhz'))
PY
echo done
