#!/bin/bash
# Two processes of the drop-in driver (metada_b200/host/_build/letkf_cuda) on two GPUs: column sharding with the NCCL
# observation halo behind LETKF<CudaBackendTag>::Analyse (MDC_RANK / MDC_WORLD_SIZE / MDC_COMM_ID_FILE), against the
# one-process run of the same configuration.  Usage (2-GPU box): bash tools/mgpu_driver_check.sh
set -e
cd "$(dirname "$0")/.."
T=$(mktemp -d)
python - "$T" <<'PY'
import sys, os
sys.path.insert(0, os.getcwd())
from tests.test_host_drivers import write_case
from tests.test_oracle_vs_reference import CASES, load
g = load(CASES[1])
for name, extra in (("one", {"streaming": "off"}), ("r0", {"streaming": "on", "slab_rows": 4}), ("r1", {"streaming": "on", "slab_rows": 4})):
    d = os.path.join(sys.argv[1], name); os.makedirs(d)
    write_case(g, d, "canonical", extra)
PY
EXE=metada_b200/host/_build/letkf_cuda
$EXE $T/one/cfg.json --dump $T/one/xa.bin > /dev/null
MDC_WORLD_SIZE=2 MDC_RANK=0 MDC_COMM_ID_FILE=$T/id $EXE $T/r0/cfg.json --dump $T/r0/xa.bin > $T/r0/log 2>&1 &
MDC_WORLD_SIZE=2 MDC_RANK=1 MDC_COMM_ID_FILE=$T/id $EXE $T/r1/cfg.json --dump $T/r1/xa.bin > $T/r1/log 2>&1
wait
cmp $T/one/xa.bin $T/r0/xa.bin && cmp $T/one/xa.bin $T/r1/xa.bin && echo "2-process driver: both ranks hold the one-process analysis, bit for bit"
