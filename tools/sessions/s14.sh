#!/bin/bash
# GPU session 14: restored tree after the container was re-created: full suite, smoke, default bench line
set -x
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv > $O/s14_gpu.txt
timeout 1500 python -m pytest tests -m gpu -x -q > $O/s14_pytest.log 2>&1; tail -3 $O/s14_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > $O/s14_smoke.log 2>&1; tail -1 $O/s14_smoke.log
timeout 900 python bench.py --steps 20 --warmup 5 > $O/s14_bench_default.json 2>> $O/s14.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > $O/s14_bench_reference.json 2>> $O/s14.err
(python tools/time_small.py 1920 1080 1; python tools/time_small.py 1920 1080 4; python tools/time_small.py 1920 1080 8) > $O/s14_time_small.txt 2>&1
echo done
