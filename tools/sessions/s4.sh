#!/bin/bash
# GPU session 4: two-pair walker variants, fused kernel parity + speed
set -x
O=gpurun_out
mkdir -p $O
L=$PWD/metalbt709decoder_b200
timeout 1200 python -m pytest tests/test_gpu_round2.py tests/test_gpu_rescale_encode.py tests/test_gpu_reference_tests.py tests/test_gpu_full_configs.py -x -q > $O/s4_pytest.log 2>&1; tail -5 $O/s4_pytest.log
for v in b200 p16_3 p16_2 p24_2; do
  BT709_B200_LIB=$L/libbt709_$v.so timeout 300 python -m pytest tests/test_gpu_round2.py -q -k "paired or fused" > $O/s4_pytest_$v.log 2>&1; tail -1 $O/s4_pytest_$v.log
  for wl in c2_rescale c2_fused c2_twopass; do
    BT709_B200_LIB=$L/libbt709_$v.so timeout 300 python bench.py --steps 50 --warmup 5 --no-configs --no-e2e --workload $wl > $O/s4_bench_${wl}_$v.json 2>> $O/s4.err
  done
done
timeout 300 python bench.py --steps 300 --warmup 5 --no-configs --no-e2e --workload c2_fused > $O/s4_bench_c2_fused_300.json 2>> $O/s4.err
timeout 300 python bench.py --steps 300 --warmup 5 --no-configs --no-e2e --workload c2_rescale > $O/s4_bench_c2_rescale_300.json 2>> $O/s4.err
ncu --set full --clock-control none --import-source on -k regex:k_fused -c 1 -s 3 -o $O/r02_fused -f python bench.py --steps 3 --warmup 3 --no-configs --no-e2e --workload c2_fused --frames 100 > $O/s4_ncu_fused.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_rescale_pair -c 1 -s 3 -o $O/r02_rescale_pair2 -f python bench.py --steps 3 --warmup 3 --no-configs --no-e2e --workload c2_rescale --frames 100 > $O/s4_ncu_pair.log 2>&1
echo done
