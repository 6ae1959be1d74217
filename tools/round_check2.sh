#!/bin/bash
# GPU-box call 2: ring locate + double-buffered row-slab transfers
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_gpu_geo.py -q > gpurun_out/geo_tests.log 2>&1; echo "geo tests rc=$?" | tee gpurun_out/summary.log
tail -12 gpurun_out/geo_tests.log
timeout 300 python -m pytest tests -m gpu -q --ignore=tests/test_gpu_geo.py > gpurun_out/gpu_tests.log 2>&1; echo "gpu suite rc=$?" | tee -a gpurun_out/summary.log
tail -12 gpurun_out/gpu_tests.log
timeout 120 python tools/geo_probe.py > gpurun_out/geo_probe.log 2>&1; echo "geo probe rc=$?" | tee -a gpurun_out/summary.log
tail -3 gpurun_out/geo_probe.log | cut -c1-700
timeout 420 python tools/e2e_probe.py > gpurun_out/e2e_probe.log 2>&1; echo "e2e probe rc=$?" | tee -a gpurun_out/summary.log
tail -8 gpurun_out/e2e_probe.log | cut -c1-600
