#!/bin/bash
# Build a kernel-experiment variant of the library: tools/build_variant.sh NAME [extra nvcc flags, e.g. -DBT709_PAIR_TH=128]
# -> build/variants/NAME.so (git-ignored, travels to the GPU box; load it with BT709_B200_LIB=build/variants/NAME.so)
set -e
cd "$(dirname "$0")/../metalbt709decoder_b200/csrc"
name=$1; shift
mkdir -p ../../build/variants
nvcc -O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC,-fvisibility=hidden,-ffp-contract=off \
  --fmad=false "$@" -shared -o ../../build/variants/$name.so \
  bt709_decode.cu bt709_fused.cu bt709_rescale.cu bt709_encode.cu bt709_misc.cu bt709_api.cu bt709_tables.cpp bt709_host.cpp y4m.cpp \
  -lcudart_static -lpthread -ldl -lrt
echo built build/variants/$name.so
