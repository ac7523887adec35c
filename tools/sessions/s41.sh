#!/bin/bash
# GPU session 41: final tree: full suite, smoke, default bench line (live traffic), reference arm, launch list, small calls
set -x
O=gpurun_out
mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -x -q > $O/s41_pytest.log 2>&1; tail -2 $O/s41_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > $O/s41_smoke.log 2>&1; tail -1 $O/s41_smoke.log
( time timeout 900 python bench.py --steps 20 --warmup 5 > $O/r02_bench_default.json 2>> $O/s41.err ) 2>&1 | grep real
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > $O/r02_bench_reference.json 2>> $O/s41.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file $O/r02_launches.csv python bench.py --steps 2 --warmup 1 --no-e2e > $O/s41_launches_bench.log 2>&1
for m in decode rescale twopass fused encode; do python tools/time_small.py 1920 1080 1 200 3 $m; done > $O/r02_time_small.txt 2>&1
for m in decode rescale twopass fused encode; do python tools/time_small.py 1920 1080 4 100 3 $m; done >> $O/r02_time_small.txt 2>&1
grep -v "^+" $O/r02_time_small.txt | head -5
echo done
