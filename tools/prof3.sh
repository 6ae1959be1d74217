set -x
for t in prof profu profe; do
  if [ $t = profu ]; then export NSP_PROFILE_UPDATE=1; else unset NSP_PROFILE_UPDATE; fi
  MDC_LIB=metada_b200/_obj/libmetada_cuda_$t.so python tools/nsp_phase_profile.py > gpurun_out/nsp_phase_r02c_$t.json 2>gpurun_out/nsp_phase_r02c_$t.err
done
python tools/nsp_quick.py > gpurun_out/nsp_quick_r02c.txt 2>&1
