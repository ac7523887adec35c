#!/bin/bash
# GPU session 5: fused kernel with L2 policies, full suite, default bench with configs, texture probe
set -x
O=gpurun_out
mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -x -q > $O/s5_pytest.log 2>&1; tail -5 $O/s5_pytest.log
for wl in c2_rescale c2_fused c2_twopass; do
  timeout 300 python bench.py --steps 50 --warmup 5 --no-configs --no-e2e --workload $wl > $O/s5_bench_$wl.json 2>> $O/s5.err
done
timeout 900 python bench.py --steps 20 --warmup 5 > $O/s5_bench_default.json 2>> $O/s5.err
tools/probes/tex_rescale 20 > $O/r02_tex_rescale_probe.json 2>> $O/s5.err
python tools/pcie_probe.py > $O/r02_pcie_probe_1gpu.json 2>> $O/s5.err
ncu --set full --clock-control none --import-source on -k regex:k_fused -c 1 -s 3 -o $O/r02_fused2 -f python bench.py --steps 3 --warmup 3 --no-configs --no-e2e --workload c2_fused --frames 100 > $O/s5_ncu_fused.log 2>&1
echo done
