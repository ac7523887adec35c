#!/bin/bash
# GPU session 23: one-frame-per-call kernel times of every entry; pass-2 occupancy variants on other geometries
set -x
O=gpurun_out
mkdir -p $O
V=$PWD/build/variants
for m in decode rescale twopass fused encode; do python tools/time_small.py 1920 1080 1 200 3 $m; done > $O/s23_time_small_modes.txt 2>&1
for m in rescale twopass encode; do python tools/time_small.py 1920 1080 4 100 3 $m; done >> $O/s23_time_small_modes.txt 2>&1
for v in cur w16pd2 w20pd2; do
  echo "== $v"
  BT709_B200_LIB=$V/$v.so python tools/time_rescale.py 1920 1080 2560 1440 100 smooth opaque
  BT709_B200_LIB=$V/$v.so python tools/time_rescale.py 1920 1080 2560 1440 100 noise
  BT709_B200_LIB=$V/$v.so python tools/time_rescale.py 1280 720 1920 1080 100 noise opaque
  BT709_B200_LIB=$V/$v.so python tools/time_rescale.py 1920 1080 3840 2160 50 noise opaque
  BT709_B200_LIB=$V/$v.so python tools/time_rescale.py 1920 1080 1920 1080 100 noise opaque
  BT709_B200_LIB=$V/$v.so python tools/time_rescale.py 2048 1536 2732 2048 60 noise opaque
  BT709_B200_LIB=$V/$v.so python tools/time_rescale.py 3840 2160 1920 2880 30 noise opaque
done > $O/s23_time_rescale.txt 2>&1
grep -v "^+" $O/s23_time_small_modes.txt; grep -v "^+" $O/s23_time_rescale.txt
echo done
