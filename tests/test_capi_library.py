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


def test_struct_layouts_match_header_sizes(tmp_path):
    """ctypes mirrors against the real header: a C program prints sizeof / offsetof of every struct."""
    import os
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    src = tmp_path / "sz.c"
    src.write_text('''#include <stdio.h>
#include <stddef.h>
#include "metada_cuda_c_api.h"
int main(void) {
  printf("%zu %zu %zu %zu %zu %zu %zu %zu\\n", sizeof(mdc_letkf_params), sizeof(mdc_letkf_stats), sizeof(mdc_enkf_diag),
         sizeof(mdc_metrics), offsetof(mdc_letkf_params, loc_scale), offsetof(mdc_letkf_params, solver),
         offsetof(mdc_letkf_stats, small_transforms), offsetof(mdc_letkf_stats, sum_sweeps));
  return 0;
}
''')
    exe = tmp_path / "sz"
    subprocess.check_call(["gcc", "-I", os.path.join(root, "include"), str(src), "-o", str(exe)])
    got = [int(v) for v in subprocess.check_output([str(exe)], text=True).split()]
    want = [ctypes.sizeof(mb.LetkfParams), ctypes.sizeof(mb.LetkfStats), ctypes.sizeof(mb.EnkfDiag), ctypes.sizeof(mb.Metrics),
            mb.LetkfParams.loc_scale.offset, mb.LetkfParams.solver.offset, mb.LetkfStats.small_transforms.offset,
            mb.LetkfStats.sum_sweeps.offset]
    assert got == want, (got, want)
