#!/bin/bash
# Development build of the API unit (mdc_api.cu: Jacobi / observation-space / global kernels) with extra defines:
# tools/build_api_variant.sh TAG -DFOO ...  ->  metada_b200/_obj/libmetada_cuda_TAG.so (use with MDC_LIB=...)
set -e
tag=$1; shift
cd "$(dirname "$0")/../metada_b200"
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -cudart shared \
  "$@" -c -o _obj/mdc_api_var_${tag}.o csrc/mdc_api.cu 2>/dev/null
objs=$(ls _obj/*.o | grep -v "mdc_api.o" | grep -v "_var_"; echo _obj/mdc_api_var_${tag}.o)
nvcc -shared -cudart shared -Xlinker -rpath=/usr/local/cuda/lib64 -o _obj/libmetada_cuda_${tag}.so $objs
ls -la _obj/libmetada_cuda_${tag}.so
