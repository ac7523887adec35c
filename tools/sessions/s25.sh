#!/bin/bash
# GPU session 25: in-tree build (fused band height for small calls on top of session 24): full suite, smoke, default bench
# line, reference arm, 1000-step runs, launch list, ncu captures of the changed kernels, one-frame-per-call times
set -x
O=gpurun_out
mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -x -q > $O/s25_pytest.log 2>&1; tail -2 $O/s25_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > $O/s25_smoke.log 2>&1; tail -1 $O/s25_smoke.log
timeout 900 python bench.py --steps 20 --warmup 5 > $O/r02_bench_default.json 2>> $O/s25.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > $O/r02_bench_reference.json 2>> $O/s25.err
timeout 600 python bench.py --steps 1000 --warmup 5 --no-configs --no-e2e > $O/r02_bench_c2_1080p_1000steps.json 2>> $O/s25.err
BT709_RING=3l timeout 600 python bench.py --steps 1000 --warmup 5 --no-configs --no-e2e > $O/r02_bench_c2_1080p_1000steps_ring3l.json 2>> $O/s25.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/r02_launches.csv python bench.py --steps 2 --warmup 1 > $O/s25_launches_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_decode_lin -c 1 -s 3 -o $O/r02_decode_lin_2e -f python bench.py --steps 3 --warmup 3 --no-configs --no-e2e --frames 150 > $O/s25_ncu_dec.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_decode_lin -c 1 -s 3 -o $O/r02_decode_lin_alpha -f python bench.py --steps 3 --warmup 3 --no-configs --no-e2e --workload c4_alpha --frames 100 > $O/s25_ncu_alpha.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_rescale_pair -c 1 -s 3 -o $O/r02_rescale_pair -f python bench.py --steps 3 --warmup 3 --no-configs --no-e2e --workload c2_rescale --frames 100 > $O/s25_ncu_pair.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_fused -c 1 -s 3 -o $O/r02_fused -f python bench.py --steps 3 --warmup 3 --no-configs --no-e2e --workload c2_fused --frames 100 > $O/s25_ncu_fused.log 2>&1
for m in decode rescale twopass fused encode; do python tools/time_small.py 1920 1080 1 200 3 $m; done > $O/r02_time_small.txt 2>&1
for m in decode rescale twopass fused encode; do python tools/time_small.py 1920 1080 4 100 3 $m; done >> $O/r02_time_small.txt 2>&1
grep -v "^+" $O/r02_time_small.txt
echo done
