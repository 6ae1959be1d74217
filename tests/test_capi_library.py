"""CPU: the C-ABI shared library builds (nvcc cross-compile), loads, and exports every symbol the
public header declares.  No compute calls (no GPU here)."""
import ctypes

import metada_b200 as mb


def test_library_builds_and_exports_all_header_symbols():
    path = mb.build_library()
    lib = ctypes.CDLL(path)
    names = mb.exported_symbols()
    assert len(names) >= 35
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing


def test_struct_layouts_match_header_sizes():
    # mdc_letkf_params: 3 doubles + 4 ints + double + 4 ints; mdc_letkf_stats: 4 floats + ...
    assert ctypes.sizeof(mb.LetkfParams) == 3 * 8 + 4 * 4 + 8 + 4 * 4
    assert ctypes.sizeof(mb.LetkfStats) == 4 * 4 + 2 * 8 + 2 * 4 + 8 + 2 * 4
    assert ctypes.sizeof(mb.EnkfDiag) == 6 * 8
