"""metada_b200 -- B200 (sm_100a) backend for METADA's ensemble Kalman analysis step.

The product is the C-ABI shared library ``libmetada_cuda.so`` (hand-written CUDA kernels behind
``include/metada_cuda_c_api.h``) plus the C++ host mirror of the reference's trait interfaces in
``metada_b200/host``.  This package is the thin ctypes binding used by tests and bench.py.
There is no CPU fallback: loading fails loudly if the library is missing, and creating a context
fails without a CUDA device.
"""
from .capi import (  # noqa: F401
    Context, Ensemble, Observations, Stream, LetkfParams, LetkfStats, EnkfDiag, Metrics, MdcError,
    MODE_REF_COMPAT, MODE_REF_ETKF, MODE_CANONICAL, LOC_CUTOFF, LOC_GASPARI_COHN, LOC_GAUSSIAN, LOC_EXPONENTIAL, LOC_REF_GASPARI_COHN,
    SOLVER_AUTO, SOLVER_JACOBI, SOLVER_NEWTON_SCHULZ, SOLVER_NEWTON_SCHULZ_FULL,
    lib_path, load_library, build_library, exported_symbols,
)
from . import synthetic  # noqa: F401
from .pipeline import StreamedLetkf  # noqa: F401
