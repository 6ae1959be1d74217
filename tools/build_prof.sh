#!/bin/bash
# Instrumented build of the nt = 10 (k = 73..80) packed column kernel: per-phase clock64() ticks (-DNSP_PROFILE).
# Links metada_b200/_obj/libmetada_cuda_prof${NSP_PROF_TAG}.so from the regular objects + the instrumented unit; use it with
#   MDC_LIB=metada_b200/_obj/libmetada_cuda_prof${NSP_PROF_TAG}.so python tools/nsp_phase_profile.py
set -e
cd "$(dirname "$0")/../metada_b200"
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -cudart shared \
  -DNSP_PROFILE $NSP_PROF_DEFS -DNSP_LO=10 -DNSP_HI=10 -c -o _obj/nsp_10_10_prof${NSP_PROF_TAG}.o csrc/nsp_tu.cu 2>/dev/null
objs=$(ls _obj/*.o | grep -v "nsp_10_10.o" | grep -v "_prof" ; echo _obj/nsp_10_10_prof${NSP_PROF_TAG}.o )
nvcc -shared -cudart shared -Xlinker -rpath=/usr/local/cuda/lib64 -o _obj/libmetada_cuda_prof${NSP_PROF_TAG}.so $objs
ls -la _obj/libmetada_cuda_prof${NSP_PROF_TAG}.so
