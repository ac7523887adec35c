#!/bin/bash
# GPU session 6: full suite on the final kernels, default bench (burst + sustained per config), launch list, captures for profiles/
set -x
O=gpurun_out
mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -x -q > $O/s6_pytest.log 2>&1; tail -3 $O/s6_pytest.log
timeout 900 python bench.py --steps 20 --warmup 5 > $O/s6_bench_default.json 2>> $O/s6.err
BT709_NO_LIN2=1 timeout 300 python bench.py --steps 20 --warmup 5 --no-configs --no-e2e > $O/s6_bench_20_nolin2.json 2>> $O/s6.err
timeout 300 python bench.py --steps 20 --warmup 5 --no-configs --no-e2e > $O/s6_bench_20.json 2>> $O/s6.err
timeout 300 python bench.py --steps 1000 --warmup 5 --no-configs --no-e2e > $O/s6_bench_1000.json 2>> $O/s6.err
timeout 300 python bench.py --steps 1000 --warmup 5 --no-configs --no-e2e --gamma linear > $O/s6_bench_1000_linear.json 2>> $O/s6.err
timeout 300 python bench.py --steps 100 --warmup 5 --no-configs --no-e2e --gamma linear > $O/s6_bench_100_linear.json 2>> $O/s6.err
timeout 300 python bench.py --steps 100 --warmup 5 --no-configs --no-e2e --workload c3_uhd > $O/s6_bench_c3.json 2>> $O/s6.err
timeout 300 python bench.py --steps 50 --warmup 5 --no-configs --no-e2e --workload c4_alpha > $O/s6_bench_c4.json 2>> $O/s6.err
timeout 300 python bench.py --steps 50 --warmup 5 --no-configs --no-e2e --workload c5_encode > $O/s6_bench_c5.json 2>> $O/s6.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/r02_launches.csv python bench.py --steps 2 --warmup 1 > $O/s6_launches_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_decode_lin -c 1 -s 3 -o $O/r02_decode_lin_alpha -f python bench.py --steps 3 --warmup 3 --no-configs --no-e2e --workload c4_alpha --frames 100 > $O/s6_ncu_alpha.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_rescale_pair -c 1 -s 3 -o $O/r02_rescale_pair3 -f python bench.py --steps 3 --warmup 3 --no-configs --no-e2e --workload c2_rescale --frames 100 > $O/s6_ncu_pair.log 2>&1
for f in 1 2 4 8; do python tools/time_small.py 1920 1080 $f; done > $O/s6_small.txt 2>&1
python -c "import __graft_entry__ as g; g.smoke()" > $O/s6_smoke.log 2>&1; tail -2 $O/s6_smoke.log
echo done
