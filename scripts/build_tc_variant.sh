#!/bin/bash
# usage: scripts/build_tc_variant.sh NAME "-DQON_TC_...=.. ..."  -> quanonet_b200/variants/libqon_NAME.so
# (experiment builds of the tensor-core tier only; run with QON_LIB_PATH=quanonet_b200/variants/libqon_NAME.so)
set -e
cd "$(dirname "$0")/.."
NAME=$1; FLAGS=$2
B=quanonet_b200/_build; mkdir -p quanonet_b200/variants
nvcc -std=c++17 -O3 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC $FLAGS -I quanonet_b200/csrc \
     -c quanonet_b200/csrc/hea_tc.cu -o $B/hea_tc_$NAME.o
OBJS="$B/hea_generic.o $B/hea_hbm.o $B/hea_reg_f32.o $B/hea_reg_f32_lanes.o $B/hea_reg_f64.o $B/hea_smem.o $B/hea_warp.o $B/qon_capi.o"
nvcc -shared -o quanonet_b200/variants/libqon_$NAME.so $OBJS $B/hea_tc_$NAME.o \
     -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC -cudart static
echo built $NAME
