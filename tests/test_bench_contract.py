"""bench.py: the algorithmic work per column it credits and the reference arm's JSON line."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def test_algorithmic_work_per_column_matches_the_survey():
    nx, ny, nz, k, P, radius = bench.WORKLOADS["C5"]
    assert (nx, ny, nz, k, P) == (1500, 1500, 60, 80, 1000000)
    # C5 canonical ~7.0 MFLOP per column at p_loc ~ 90 (9 k^3 credited for the eigensolve), 77.1 KB
    assert abs(bench.flops_per_column(k, 90.0, nz) - 7.0e6) < 0.1e6
    assert abs(bench.bytes_per_column(k, nz, P, nx * ny) - 77.1e3) < 0.1e3
    nx, ny, nz, k, P, radius = bench.WORKLOADS["C3"]
    assert abs(bench.bytes_per_column(k, nz, P, nx * ny) - 32.2e3) < 0.1e3
    assert abs(bench.flops_per_column(k, 93.0, nz) - 1.0e6) < 0.1e6


def test_reference_arm_prints_the_contract_line():
    """--impl reference times the oracle port on the host cores (the reference's LETKF.hpp needs Eigen); here on the
    smallest workload so the test stays short."""
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "C1",
                        "--steps", "1", "--warmup", "0"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["metric"] == "LETKF analysed grid-columns/sec"
    assert line["unit"] == "columns/s" and line["higher_is_better"] is True and line["value"] > 0
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["d2h_bytes_per_step"] == 0
    assert line["e2e"]["value"] == line["value"]
