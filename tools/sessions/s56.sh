#!/bin/bash
# GPU session 56: last run of round 2 on the final tree: full suite, smoke, default bench line
set -x
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/s56_pytest.log 2>&1; tail -2 $O/s56_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > $O/s56_smoke.log 2>&1; tail -1 $O/s56_smoke.log
timeout 300 python bench.py --gpus 1 --steps 20 --warmup 5 > $O/r02_bench_default.json 2>> $O/s56.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02_bench_default.json'))
print('default', round(d['value'],1), round(d['roofline']['frac'],4), 'sustained', round(d['sustained']['value'],1), 'e2e', d['e2e']['value'], d['cpu_baseline'].get('parity'))
for c in d['configs']:
    print(c['workload'][:36], round(c['value'], 1), round(c['roofline']['frac'], 4), 'sust', round(c['sustained']['value'], 1), round(c['sustained']['roofline']['frac'], 4))
PY
echo done
