#!/bin/bash
# GPU session 2: ring depth of k_decode_lin (2/3/4), burst and sustained, BT.709 and Linear curves; small-call kernel time
set -x
O=gpurun_out
mkdir -p $O
L=$PWD/metalbt709decoder_b200
for d in d3 d4 b200; do
  BT709_B200_LIB=$L/libbt709_$d.so timeout 300 python -m pytest tests/test_gpu_round2.py -q -k "linear_tile_stream" > $O/s2_pytest_$d.log 2>&1; tail -1 $O/s2_pytest_$d.log
  for g in bt709 linear; do
    BT709_B200_LIB=$L/libbt709_$d.so timeout 300 python bench.py --steps 100 --warmup 5 --no-configs --no-e2e --gamma $g > $O/s2_bench_100_${g}_$d.json 2>> $O/s2.err
    BT709_B200_LIB=$L/libbt709_$d.so timeout 300 python bench.py --steps 1000 --warmup 5 --no-configs --no-e2e --gamma $g > $O/s2_bench_1000_${g}_$d.json 2>> $O/s2.err
  done
done
BT709_NO_BLOG=1 BT709_B200_LIB=$L/libbt709_b200.so timeout 300 python bench.py --steps 1000 --warmup 5 --no-configs --no-e2e --gamma linear > $O/s2_bench_1000_linear_noblog.json 2>> $O/s2.err
BT709_NO_LIN2=1 timeout 300 python bench.py --steps 1000 --warmup 5 --no-configs --no-e2e > $O/s2_bench_1000_bt709_nolin2.json 2>> $O/s2.err
for f in 1 2 4 8; do BT709_B200_LIB=$L/libbt709_d3.so python tools/time_small.py 1920 1080 $f; done > $O/s2_small.txt 2>&1
for f in 1 2 4 8; do BT709_SMALL_STEPS=0 BT709_B200_LIB=$L/libbt709_d3.so python tools/time_small.py 1920 1080 $f; done >> $O/s2_small.txt 2>&1
for f in 1 4; do BT709_NO_LIN2=1 python tools/time_small.py 1920 1080 $f; done >> $O/s2_small.txt 2>&1
BT709_B200_LIB=$L/libbt709_d3.so ncu --set full --clock-control none --import-source on -k regex:k_decode_lin -c 1 -s 3 -o $O/r02_decode_lin_d3 -f python bench.py --steps 3 --warmup 3 --no-configs --no-e2e --frames 150 > $O/s2_ncu_lin.log 2>&1
echo done
