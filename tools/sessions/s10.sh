#!/bin/bash
# GPU session 10 (8 GPUs): PCIe ceiling at 1/2/4/8 GPUs at once, bench at 8 GPUs (kernel metric + e2e + copy-only ceiling)
set -x
O=gpurun_out
mkdir -p $O
nvidia-smi topo -m > $O/r02_topo_8gpu.txt 2>&1
lscpu | head -25 >> $O/r02_topo_8gpu.txt 2>&1
python tools/pcie_probe.py > $O/r02_pcie_probe_n1.json 2>> $O/s10.err
for n in 2 4 8; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 tools/pcie_probe.py > $O/r02_pcie_probe_n$n.json 2>> $O/s10.err
done
for n in 8 4 2; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $n --steps 20 --warmup 5 > $O/r02_bench_${n}gpu.json 2>> $O/s10.err
done
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 8 --steps 20 --warmup 5 --workload c3_uhd --no-e2e > $O/r02_bench_8gpu_uhd.json 2>> $O/s10.err
echo done
