#!/bin/bash
# GPU session 11: final validation of the round-2 tree: full suite, smoke, default bench, launch list, decode capture
set -x
O=gpurun_out
mkdir -p $O
python -c "import __graft_entry__ as g; g.smoke()" > $O/s11_smoke.log 2>&1; tail -2 $O/s11_smoke.log
timeout 1500 python -m pytest tests -m gpu -x -q > $O/s11_pytest.log 2>&1; tail -3 $O/s11_pytest.log
timeout 900 python bench.py --steps 20 --warmup 5 > $O/r02_bench_default.json 2>> $O/s11.err
timeout 300 python bench.py --impl reference --steps 20 --warmup 5 > $O/r02_bench_reference.json 2>> $O/s11.err
timeout 300 python bench.py --steps 1000 --warmup 5 --no-configs --no-e2e > $O/r02_bench_c2_1080p_1000steps.json 2>> $O/s11.err
BT709_RING=3l timeout 300 python bench.py --steps 1000 --warmup 5 --no-configs --no-e2e > $O/r02_bench_c2_1080p_1000steps_ring3l.json 2>> $O/s11.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r02_launches.csv python bench.py --steps 2 --warmup 1 --no-e2e > $O/s11_launches_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_decode_lin -c 1 -s 3 -o $O/r02_decode_lin_2e -f python bench.py --steps 3 --warmup 3 --no-configs --no-e2e --frames 150 > $O/s11_ncu_lin.log 2>&1
BT709_FUSED=1 timeout 300 python bench.py --steps 50 --warmup 5 --no-configs --no-e2e --workload c2_fused > $O/s11_bench_fused.json 2>> $O/s11.err
echo done
