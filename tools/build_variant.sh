#!/bin/bash
# Development build of the nt = 10 packed column kernel with extra defines: tools/build_variant.sh TAG -DFOO ...
# -> metada_b200/_obj/libmetada_cuda_TAG.so (use with MDC_LIB=...)
set -e
tag=$1; shift
cd "$(dirname "$0")/../metada_b200"
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -cudart shared \
  "$@" -DNSP_LO=10 -DNSP_HI=10 -c -o _obj/nsp_10_10_var_${tag}.o csrc/nsp_tu.cu 2>/dev/null
objs=$(ls _obj/*.o | grep -v "nsp_10_10.o" | grep -v "_prof" | grep -v "_var_"; echo _obj/nsp_10_10_var_${tag}.o)
nvcc -shared -cudart shared -Xlinker -rpath=/usr/local/cuda/lib64 -o _obj/libmetada_cuda_${tag}.so $objs
ls -la _obj/libmetada_cuda_${tag}.so
