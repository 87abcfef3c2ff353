#!/bin/bash
# build an A/B variant of the library: tools/build_variant.sh NAME [extra nvcc flags...]
# -> marllb_b200/_variants/NAME.so ; select at run time with MARLLB_B200_LIB=marllb_b200/_variants/NAME.so
name=$1; shift
mkdir -p marllb_b200/_variants
cd marllb_b200/csrc && nvcc -ccbin /usr/bin/g++ -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC -shared "$@" \
    -o ../_variants/$name.so mlb_api.cu mlb_ops.cu mlb_policy.cu mlb_linear_tc.cu mlb_gemm_tc.cu
