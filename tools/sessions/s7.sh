#!/bin/bash
# GPU session 7: ring refill order of k_decode_lin (early / late, depth 2 / 3) against the round-1 ring kernel, same box
set -x
O=gpurun_out
mkdir -p $O
L=$PWD/metalbt709decoder_b200
timeout 600 python -m pytest tests/test_gpu_round2.py tests/test_gpu_decode.py -x -q -k "linear_tile or lattice or lut_replication or random_frames or alpha" > $O/s7_pytest.log 2>&1; tail -2 $O/s7_pytest.log
for rep in 1 2; do
for v in b200 d3late d3early; do
  for n in 20 100 1000; do
    BT709_B200_LIB=$L/libbt709_$v.so timeout 300 python bench.py --steps $n --warmup 5 --no-configs --no-e2e > $O/s7_bench_${n}_${v}_$rep.json 2>> $O/s7.err
  done
done
for n in 20 100 1000; do
  BT709_NO_LIN2=1 timeout 300 python bench.py --steps $n --warmup 5 --no-configs --no-e2e > $O/s7_bench_${n}_nolin2_$rep.json 2>> $O/s7.err
done
done
for v in b200 d3late; do
  for n in 20 1000; do
    BT709_B200_LIB=$L/libbt709_$v.so timeout 300 python bench.py --steps $n --warmup 5 --no-configs --no-e2e --gamma linear > $O/s7_bench_${n}_linear_${v}.json 2>> $O/s7.err
    BT709_B200_LIB=$L/libbt709_$v.so timeout 300 python bench.py --steps $n --warmup 5 --no-configs --no-e2e --workload c4_alpha > $O/s7_bench_${n}_alpha_${v}.json 2>> $O/s7.err
    BT709_B200_LIB=$L/libbt709_$v.so timeout 300 python bench.py --steps $n --warmup 5 --no-configs --no-e2e --workload c3_uhd > $O/s7_bench_${n}_uhd_${v}.json 2>> $O/s7.err
  done
done
echo done
