#!/bin/bash
# tools/build_variant.sh NAME [nvcc -D flags...] : builds nohuman_b200/variants/libnh_NAME.so for A/B runs
# (python tools/kernel_ab.py picks them up through NH_LIB_PATH).  Built files are git-ignored.
set -e
cd "$(dirname "$0")/.."
name=$1; shift
mkdir -p nohuman_b200/variants
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 \
  -Xcompiler -fPIC,-O3,-Wall,-pthread -Xptxas -v "$@" -ccbin /usr/bin/g++ -shared \
  -o nohuman_b200/variants/libnh_$name.so nohuman_b200/csrc/nh_kernels.cu nohuman_b200/csrc/nh_capi.cu \
  nohuman_b200/csrc/nh_synth.cu nohuman_b200/csrc/nh_pipeline.cc nohuman_b200/csrc/nh_pack.cc -lz -lpthread -ldl \
  > nohuman_b200/variants/build_$name.log 2>&1
grep -A3 "k_stream_classifyILi5ELb0ELb0" nohuman_b200/variants/build_$name.log | grep -E "Used|spill" | tr '\n' ' '
echo " -> $name"
