#!/bin/bash
# build an experimental variant of the library next to the product one:  tools/build_variant.sh wide64 -DSMB_PH_WIDE64=1
# -> stylemesh_b200/lib/libstylemesh_b200_<name>.so (select with SMB_LIB=<path>); only tc_igemm_v5.cu is recompiled
set -e
name=$1; shift
cd "$(dirname "$0")/.."
L=stylemesh_b200/lib
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC --expt-relaxed-constexpr "$@" \
  -c stylemesh_b200/csrc/tc_igemm_v5.cu -o $L/obj/tc_igemm_v5_$name.o
objs=$(ls $L/obj/*.o | grep -v "tc_igemm_v5" | tr '\n' ' ')
nvcc -shared -gencode arch=compute_100a,code=sm_100a -o $L/libstylemesh_b200_$name.so $objs $L/obj/tc_igemm_v5_$name.o
echo built $L/libstylemesh_b200_$name.so
