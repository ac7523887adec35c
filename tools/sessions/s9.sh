#!/bin/bash
# GPU session 9: register-load decode variant (k_decode_reg) against the ring variants, same box
set -x
O=gpurun_out
mkdir -p $O
for pd in 1 2 3; do
  BT709_REG=$pd timeout 600 python -m pytest tests/test_gpu_round2.py tests/test_gpu_decode.py -x -q -k "linear_tile or lattice_decode or random_frames" > $O/s9_pytest_reg$pd.log 2>&1; tail -1 $O/s9_pytest_reg$pd.log
done
for rep in 1 2; do
  for n in 20 100 1000; do
    for pd in 1 2 3; do
      BT709_REG=$pd timeout 300 python bench.py --steps $n --warmup 5 --no-configs --no-e2e > $O/s9_bench_${n}_reg${pd}_$rep.json 2>> $O/s9.err
    done
    timeout 300 python bench.py --steps $n --warmup 5 --no-configs --no-e2e > $O/s9_bench_${n}_ring2e_$rep.json 2>> $O/s9.err
    BT709_RING=3l timeout 300 python bench.py --steps $n --warmup 5 --no-configs --no-e2e > $O/s9_bench_${n}_ring3l_$rep.json 2>> $O/s9.err
  done
done
for n in 20 1000; do
  BT709_REG=2 timeout 300 python bench.py --steps $n --warmup 5 --no-configs --no-e2e --gamma linear > $O/s9_bench_${n}_linear_reg2.json 2>> $O/s9.err
  timeout 300 python bench.py --steps $n --warmup 5 --no-configs --no-e2e --gamma linear > $O/s9_bench_${n}_linear_ring.json 2>> $O/s9.err
  BT709_REG=2 timeout 300 python bench.py --steps $n --warmup 5 --no-configs --no-e2e --workload c3_uhd > $O/s9_bench_${n}_uhd_reg2.json 2>> $O/s9.err
  timeout 300 python bench.py --steps $n --warmup 5 --no-configs --no-e2e --workload c3_uhd > $O/s9_bench_${n}_uhd_ring.json 2>> $O/s9.err
done
echo done
