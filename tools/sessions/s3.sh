#!/bin/bash
# GPU session 3: k_rescale_pair parity + speed, ncu
set -x
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_round2.py tests/test_gpu_rescale_encode.py tests/test_gpu_full_configs.py -x -q -k "rescale or pair or policy or c2_600" > $O/s3_pytest.log 2>&1; tail -3 $O/s3_pytest.log
timeout 600 python -m pytest tests/test_gpu_decode.py -x -q -k "superseded" >> $O/s3_pytest.log 2>&1; tail -3 $O/s3_pytest.log
for wl in c2_rescale c2_twopass; do
  timeout 300 python bench.py --steps 50 --warmup 5 --no-configs --no-e2e --workload $wl > $O/s3_bench_$wl.json 2>> $O/s3.err
  BT709_NO_PAIR=1 timeout 300 python bench.py --steps 50 --warmup 5 --no-configs --no-e2e --workload $wl > $O/s3_bench_${wl}_nopair.json 2>> $O/s3.err
done
(python tools/time_rescale.py 1920 1080 2560 1440 64; python tools/time_rescale.py 1920 1080 2560 1440 64 smooth; python tools/time_rescale.py 1280 720 1920 1080 64; python tools/time_rescale.py 1920 1080 3840 2160 32; python tools/time_rescale.py 1920 1080 1920 1080 64) > $O/s3_time_rescale.txt 2>&1
(BT709_NO_PAIR=1 python tools/time_rescale.py 1920 1080 2560 1440 64; BT709_NO_PAIR=1 python tools/time_rescale.py 1920 1080 2560 1440 64 smooth) >> $O/s3_time_rescale.txt 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_rescale_pair -c 1 -s 3 -o $O/r02_rescale_pair -f python bench.py --steps 3 --warmup 3 --no-configs --no-e2e --workload c2_rescale --frames 100 > $O/s3_ncu.log 2>&1
echo done
