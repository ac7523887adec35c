#!/bin/bash
# GPU session 27: final tree of the session (k_rescale_pair 20 warps / 2 rows ahead): full suite, smoke, default bench line,
# reference arm, 1000-step runs, launch list with the configs, ncu capture of pass 2, one-frame-per-call times, sanitizer
set -x
O=gpurun_out
mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -x -q > $O/s27_pytest.log 2>&1; tail -2 $O/s27_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > $O/s27_smoke.log 2>&1; tail -1 $O/s27_smoke.log
timeout 900 python bench.py --steps 20 --warmup 5 > $O/r02_bench_default.json 2>> $O/s27.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > $O/r02_bench_reference.json 2>> $O/s27.err
timeout 600 python bench.py --steps 1000 --warmup 5 --no-configs --no-e2e > $O/r02_bench_c2_1080p_1000steps.json 2>> $O/s27.err
BT709_RING=3l timeout 600 python bench.py --steps 1000 --warmup 5 --no-configs --no-e2e > $O/r02_bench_c2_1080p_1000steps_ring3l.json 2>> $O/s27.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file $O/r02_launches.csv python bench.py --steps 2 --warmup 1 --no-e2e > $O/s27_launches_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_rescale_pair -c 1 -s 3 -o $O/r02_rescale_pair -f python bench.py --steps 3 --warmup 3 --no-configs --no-e2e --workload c2_rescale --frames 100 > $O/s27_ncu_pair.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_encode_fast -c 1 -s 3 -o $O/r02_encode_fast -f python bench.py --steps 3 --warmup 3 --no-configs --no-e2e --workload c5_encode --frames 100 > $O/s27_ncu_enc.log 2>&1
for m in decode rescale twopass fused encode; do python tools/time_small.py 1920 1080 1 200 3 $m; done > $O/r02_time_small.txt 2>&1
for m in decode rescale twopass fused encode; do python tools/time_small.py 1920 1080 4 100 3 $m; done >> $O/r02_time_small.txt 2>&1
for m in decode rescale twopass encode; do BT709_PDL=0 python tools/time_small.py 1920 1080 1 200 3 $m; done > $O/r02_time_small_nopdl.txt 2>&1
grep -v "^+" $O/r02_time_small.txt
bash tools/sessions/s12.sh > $O/s27_sanitizer.log 2>&1; grep -E "SUMMARY|rc=|passed|failed|sanitize_small" $O/r02_sanitizer_*.log | head -30
echo done
