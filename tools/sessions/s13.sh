#!/bin/bash
# GPU session 13: final tree: full suite, encode gamma pairs, refreshed default bench line + launch list
set -x
O=gpurun_out
mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -x -q > $O/s13_pytest.log 2>&1; tail -3 $O/s13_pytest.log
(python tools/time_decode.py 2048 1536 200 0 encode 1 0; python tools/time_decode.py 2048 1536 200 0 encode 1 1; python tools/time_decode.py 2048 1536 200 0 encode 2 2; python tools/time_decode.py 2048 1536 200 0 encode 0 0) > $O/s13_encode_pairs.txt 2>&1
timeout 900 python bench.py --steps 20 --warmup 5 > $O/r02_bench_default.json 2>> $O/s13.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r02_launches.csv python bench.py --steps 2 --warmup 1 --no-e2e > $O/s13_launches_bench.log 2>&1
python -c "import __graft_entry__ as g; g.smoke()" > $O/s13_smoke.log 2>&1; tail -1 $O/s13_smoke.log
echo done
