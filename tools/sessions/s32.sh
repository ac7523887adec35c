#!/bin/bash
# GPU session 32 (8 GPUs): final tree, C2 and C3 frame-sharded, PCIe probe
set -x
O=gpurun_out
mkdir -p $O
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 8 --steps 20 --warmup 5 2> $O/s32_c2.err | grep -v "^NCCL version" > $O/r02_bench_c2_1080p_8gpu.json
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29532 bench.py --gpus 8 --steps 20 --warmup 5 --workload c3_uhd --no-e2e 2> $O/s32_c3.err | grep -v "^NCCL version" > $O/r02_bench_c3_uhd_8gpu.json
python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 4 --steps 20 --warmup 5 2> $O/s32_c2_4.err | grep -v "^NCCL version" > $O/r02_bench_c2_1080p_4gpu.json
python - <<'PY'
import json
for f in ['r02_bench_c2_1080p_8gpu','r02_bench_c3_uhd_8gpu','r02_bench_c2_1080p_4gpu']:
    d=json.load(open(f'gpurun_out/{f}.json'))
    print(f, round(d['value'],1), d['n_gpus'], round(d['roofline']['frac'],4), 'e2e', d['e2e'].get('value'), d['e2e'].get('copy_only_value'), d['clocks'])
PY
echo done
