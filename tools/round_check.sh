#!/bin/bash
# One GPU-box call: new geographic / multi-variable parity tests, the whole GPU suite, the WRF-shaped probe, the bench line.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/smi.log 2>&1
timeout 300 python -m pytest tests/test_gpu_geo.py -q > gpurun_out/geo_tests.log 2>&1; echo "geo tests rc=$?" | tee -a gpurun_out/summary.log
tail -15 gpurun_out/geo_tests.log
timeout 420 python -m pytest tests -m gpu -q --ignore=tests/test_gpu_geo.py --durations=8 > gpurun_out/gpu_tests.log 2>&1; echo "gpu suite rc=$?" | tee -a gpurun_out/summary.log
tail -25 gpurun_out/gpu_tests.log
timeout 150 python tools/geo_probe.py > gpurun_out/geo_probe.log 2>&1; echo "geo probe rc=$?" | tee -a gpurun_out/summary.log
tail -4 gpurun_out/geo_probe.log
timeout 330 python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench.err; echo "bench rc=$?" | tee -a gpurun_out/summary.log
cut -c1-600 gpurun_out/bench_final.json
