#!/bin/bash
# build a measurement variant of libmpxb.so: scripts/build_variant.sh NAME unit.cu [nvcc flags...] -> build/variants/libmpxb_NAME.so
# (the other translation units are taken from build/obj; run the variant with MPXB_LIB_PATH=build/variants/libmpxb_NAME.so)
set -e
name=$1; unit=$2; shift 2
mkdir -p build/variants
obj=build/variants/${unit%.cu}_$name.o
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 --extended-lambda -Xcompiler -fPIC "$@" -c -o $obj mpx_b200/csrc/$unit
others=$(ls build/obj/*.o | grep -v "/${unit%.cu}.o")
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o build/variants/libmpxb_$name.so $obj $others
echo built build/variants/libmpxb_$name.so
