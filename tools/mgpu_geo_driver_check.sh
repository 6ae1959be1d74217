#!/bin/bash
# Two processes of the geographic host application (metada_b200/host/_build/geo_letkf_cuda) on two GPUs: the C++
# runtime's sharded geographic analysis (mdc_geo_sharded_analyse: NCCL halo rows selected by box) against the
# one-process, one-store run of the same input.  Usage (2-GPU box): bash tools/mgpu_geo_driver_check.sh
set -e
cd "$(dirname "$0")/.."
T=$(mktemp -d)
python - "$T" <<'PY'
import sys, os, struct
sys.path.insert(0, os.getcwd())
import numpy as np
from metada_b200 import synthetic as syn
nx, ny, nz, k, P, radius = 96, 80, 6, 40, 6000, 60.0
lat, lon = syn.geography(nx, ny, lon0=177.0)
vc = np.array([1000.0, 925.0, 850.0, 700.0, 500.0, 300.0])
o = syn.geo_observations(P, lat, lon, vc, seed=31)
X = syn.ensemble(k, nx, ny, nz, seed=88)
with open(os.path.join(sys.argv[1], "in.bin"), "wb") as f:
    f.write(struct.pack("<6qd", nx, ny, nz, k, P, len(vc), radius))
    for a in (lat, lon, vc, X, o["lat"], o["lon"], o["level"], o["value"], o["err"], o["valid"].astype(np.float64)):
        f.write(np.ascontiguousarray(a, dtype="<f8").tobytes())
PY
EXE=metada_b200/host/_build/geo_letkf_cuda
$EXE $T/in.bin $T/one.bin > /dev/null
MDC_WORLD_SIZE=2 MDC_RANK=0 MDC_COMM_ID_FILE=$T/id $EXE $T/in.bin $T/r0.bin > $T/r0.log 2>&1 &
MDC_WORLD_SIZE=2 MDC_RANK=1 MDC_COMM_ID_FILE=$T/id $EXE $T/in.bin $T/r1.bin > $T/r1.log 2>&1
wait
cat $T/r0.log $T/r1.log
cmp $T/one.bin $T/r0.bin && cmp $T/one.bin $T/r1.bin && echo "2-process geographic driver: both ranks hold the one-store analysis, bit for bit"
