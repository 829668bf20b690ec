#!/bin/bash
# usage: scripts/build_variant.sh NAME "-DQON_...=.. ..."  -> quanonet_b200/variants/libqon_NAME.so (experiment builds)
set -e
cd "$(dirname "$0")/.."
NAME=$1; FLAGS=$2
B=quanonet_b200/_build; mkdir -p quanonet_b200/variants
nvcc -std=c++17 -O3 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC $FLAGS -I quanonet_b200/csrc \
     -c quanonet_b200/csrc/hea_reg_f32.cu -o $B/hea_reg_f32_$NAME.o
nvcc -shared -o quanonet_b200/variants/libqon_$NAME.so $B/qon_capi.o $B/hea_reg_f32_$NAME.o $B/hea_reg_f32_lanes.o $B/hea_reg_f64.o $B/hea_smem.o $B/hea_hbm.o $B/hea_generic.o \
     -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC -cudart static
echo built $NAME
