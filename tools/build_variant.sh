#!/bin/bash
# build an A/B variant of the library: tools/build_variant.sh NAME [extra nvcc flags for the env kernels...]
# -> marllb_b200/_variants/NAME.so ; select at run time with MARLLB_B200_LIB=marllb_b200/_variants/NAME.so
# Only mlb_api.cu (which holds the env-step kernels) is recompiled with the extra flags; the other objects are cached.
name=$1; shift
mkdir -p marllb_b200/_variants /tmp/vobj
F="-ccbin /usr/bin/g++ -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC"
cd marllb_b200/csrc
for f in mlb_ops mlb_policy mlb_linear_tc mlb_gemm_tc; do
  [ /tmp/vobj/$f.o -nt $f.cu ] || nvcc $F -c -o /tmp/vobj/$f.o $f.cu &
done
nvcc $F "$@" -c -o /tmp/vobj/api_$name.o mlb_api.cu &
wait
nvcc $F -shared -o ../_variants/$name.so /tmp/vobj/api_$name.o /tmp/vobj/mlb_ops.o /tmp/vobj/mlb_policy.o /tmp/vobj/mlb_linear_tc.o /tmp/vobj/mlb_gemm_tc.o
