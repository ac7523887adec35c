#!/bin/bash
# GPU session 1 (round 2): parity, bench, launch list, ncu of k_decode_lin, small-call latency, ring depth 3
set -x
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv > $O/s1_gpu.txt 2>&1
(ls /usr/lib/x86_64-linux-gnu | grep -i -E "nvcuvid|nvidia-encode|nvjpeg"; ldconfig -p | grep -i -E "cuvid|nvenc") > $O/s1_nvdec_probe.txt 2>&1
nproc >> $O/s1_gpu.txt; numactl -H >> $O/s1_gpu.txt 2>&1; nvidia-smi topo -m >> $O/s1_gpu.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q > $O/s1_pytest.log 2>&1; echo "pytest rc=$?" >> $O/s1_pytest.log
tail -5 $O/s1_pytest.log
timeout 600 python bench.py --steps 20 --warmup 5 > $O/s1_bench_default.json 2> $O/s1_bench_default.err
timeout 300 python bench.py --steps 1000 --warmup 5 --no-configs --no-e2e > $O/s1_bench_1000.json 2>> $O/s1_bench_default.err
timeout 300 python bench.py --steps 1000 --warmup 5 --no-configs --no-e2e --gamma linear > $O/s1_bench_1000_linear.json 2>> $O/s1_bench_default.err
timeout 300 python bench.py --steps 100 --warmup 5 --no-configs --no-e2e --gamma apple > $O/s1_bench_100_apple.json 2>> $O/s1_bench_default.err
BT709_NO_LIN2=1 timeout 300 python bench.py --steps 100 --warmup 5 --no-configs --no-e2e > $O/s1_bench_100_nolin2.json 2>> $O/s1_bench_default.err
timeout 300 python bench.py --steps 100 --warmup 5 --no-configs --no-e2e > $O/s1_bench_100.json 2>> $O/s1_bench_default.err
BT709_B200_LIB=$PWD/metalbt709decoder_b200/libbt709_d3.so timeout 300 python bench.py --steps 100 --warmup 5 --no-configs --no-e2e > $O/s1_bench_100_depth3.json 2>> $O/s1_bench_default.err
for f in 1 2 4 8 16 64; do python tools/time_decode.py 1920 1080 $f; done > $O/s1_small.txt 2>&1
for f in 1 2 4 8; do BT709_SMALL_STEPS=0 python tools/time_decode.py 1920 1080 $f; done >> $O/s1_small.txt 2>&1
for f in 1 4; do BT709_NO_LIN2=1 python tools/time_decode.py 1920 1080 $f; done >> $O/s1_small.txt 2>&1
python tools/time_decode.py 3840 2160 64 >> $O/s1_small.txt 2>&1
python tools/time_decode.py 1280 720 256 >> $O/s1_small.txt 2>&1
python tools/time_decode.py 4096 2160 64 >> $O/s1_small.txt 2>&1
python tools/time_decode.py 7680 4320 16 >> $O/s1_small.txt 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r02_launches.csv python bench.py --steps 2 --warmup 1 --no-e2e > $O/s1_launches_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_decode_lin -c 1 -s 3 -o $O/r02_decode_lin -f python bench.py --steps 3 --warmup 3 --no-configs --no-e2e --frames 150 > $O/s1_ncu_lin.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_encode_fast -c 1 -s 3 -o $O/r02_encode_fast -f python bench.py --steps 3 --warmup 3 --no-configs --no-e2e --workload c5_encode --frames 100 > $O/s1_ncu_enc.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_rescale_cols2 -c 1 -s 3 -o $O/r02_rescale_cols2 -f python bench.py --steps 3 --warmup 3 --no-configs --no-e2e --workload c2_rescale --frames 100 > $O/s1_ncu_resc.log 2>&1
echo done
