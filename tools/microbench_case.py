"""The FP64 / HBM microbenchmarks behind the roofline denominators (mdc_bench_*), for ncu captures."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import metada_b200 as mb
ctx = mb.Context(0)
print({"fp64_fma_tflops": ctx.bench_fp64_fma(), "fp64_dmma_tflops": ctx.bench_fp64_dmma(), "hbm_copy_gbs": ctx.bench_hbm_copy()})
