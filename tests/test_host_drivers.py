"""The C++ host layer: the reference's own adapters/concepts instantiated with CudaBackendTag
(metada_b200/host), driven like the reference's `letkf/etkf/enkf <config>` applications with the
reference's config schema and text file formats.  Binaries are prebuilt by __graft_entry__.build()
in the build container (they need the reference's headers to compile) and travel to the GPU box."""
import json
import os
import struct
import subprocess

import numpy as np
import pytest

from tests.common import analysis_errors
from tests.test_oracle_vs_reference import CASES, load

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BUILD = os.path.join(ROOT, "metada_b200", "host", "_build")


def _need(binary):
    path = os.path.join(BUILD, binary)
    if not os.path.exists(path):
        pytest.fail(f"{path} missing: run __graft_entry__.build() where /root/reference is available")
    return path


def write_case(g, tmp, mode, extra=None):
    """Simple-backend text formats (SimpleState.hpp:248-263, GridObservation.hpp:401-476) + JSON config."""
    k, ny, nx = int(g["k"]), int(g["ny"]), int(g["nx"])
    members = []
    for m in range(k):
        p = os.path.join(tmp, f"ens_{m}.txt")
        with open(p, "w") as f:
            for y in range(ny):
                f.write(" ".join(repr(float(v)) for v in g["X"][m, 0, y]) + "\n")
        members.append({"state": {"variables": "simple", "file": p}})
    op = os.path.join(tmp, "obs.txt")
    with open(op, "w") as f:
        f.write("-" * 80 + "\nTime Z   Y   X   ZNU      XLONG_U    XLAT_U     Simple\n" + "-" * 80 + "\n")
        for i in range(int(g["P"])):
            f.write(f"0 {int(g['oz'][i])} {int(g['oy'][i])} {int(g['ox'][i])} 0.5 0.0 0.0 {float(g['yo'][i])!r}\n")
    analysis = {"algorithm": "letkf", "inflation": float(np.float32(g["inflation"])), "localization_radius": float(g["radius"]),
                "inflation_method": "multiplicative", "format": "txt", "mode": mode,
                "output_base_file": os.path.join(tmp, "analysis")}
    analysis.update(extra or {})
    cfg = {"logger": {"app_name": "test", "level": "error", "color": False, "console": False},
           "geometry": {"x_dim": nx, "y_dim": ny},
           "ensemble": {"members": members},
           "observations": {"types": [{"obs_A": {"if_use": True, "file": op, "coordinate": "grid",
                                                 "variables": [{"simple": {"if_use": True, "error": float(g["err_cfg"]), "missing_value": -999.0}}]}}]},
           "obs_operator": {}, "analysis": analysis}
    cp = os.path.join(tmp, "cfg.json")
    json.dump(cfg, open(cp, "w"))
    return cp


def read_dump(path, ny, nx):
    b = open(path, "rb").read()
    k, n = struct.unpack_from("<qq", b, 0)
    return np.frombuffer(b, dtype="<f8", offset=16).reshape(k, 1, ny, nx)


def test_driver_without_gpu_fails_loudly_no_cpu_fallback(tmp_path):
    import shutil
    if shutil.which("nvidia-smi") and subprocess.run(["nvidia-smi", "-L"], capture_output=True).returncode == 0:
        pytest.skip("a GPU is present")
    exe = _need("letkf_cuda")
    g = load(CASES[1])
    cfg = write_case(g, str(tmp_path), "ref_compat")
    r = subprocess.run([exe, cfg], capture_output=True, text=True)
    assert r.returncode == 1
    assert "no usable CUDA device" in r.stderr and "no CPU fallback" in r.stderr


@pytest.mark.gpu
@pytest.mark.parametrize("name", CASES)
def test_letkf_driver_matches_reference_snapshot(tmp_path, name):
    exe = _need("letkf_cuda")
    g = load(name)
    cfg = write_case(g, str(tmp_path), "ref_compat")
    dump = str(tmp_path / "xa.bin")
    r = subprocess.run([exe, cfg, "--dump", dump], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr + r.stdout
    Xa = read_dump(dump, int(g["ny"]), int(g["nx"]))
    em, ep = analysis_errors(Xa, g["Xa_letkf_snapshot"])
    assert em < 1e-10 and ep < 1e-10, (em, ep)
    # saveEnsemble(): <base>_mean.txt and <base>_member_<i>.txt in the Simple text format (6 decimals)
    mean_txt = np.loadtxt(str(tmp_path / "analysis_mean.txt"))
    assert mean_txt.shape == (int(g["ny"]), int(g["nx"]))
    assert np.abs(mean_txt - g["mean_letkf_snapshot"]).max() < 1e-6
    m3 = np.loadtxt(str(tmp_path / "analysis_member_3.txt"))
    assert np.abs(m3 - g["Xa_letkf_snapshot"][3, 0]).max() < 1e-6


@pytest.mark.gpu
def test_etkf_and_enkf_drivers_match_reference(tmp_path):
    g = load(CASES[1])
    d1 = tmp_path / "etkf"; d1.mkdir()
    cfg = write_case(g, str(d1), "ref_compat")
    dump = str(d1 / "xa.bin")
    r = subprocess.run([_need("etkf_cuda"), cfg, "--dump", dump], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr + r.stdout
    em, ep = analysis_errors(read_dump(dump, int(g["ny"]), int(g["nx"])), g["Xa_etkf"])
    assert em < 1e-10 and ep < 1e-10, (em, ep)
    d2 = tmp_path / "enkf"; d2.mkdir()
    zf = str(d2 / "Z.bin")
    np.ascontiguousarray(g["enkf_Z"], dtype="<f8").tofile(zf)
    cfg = write_case(g, str(d2), "ref_compat", {"perturbation_file": zf})
    dump = str(d2 / "xa.bin")
    r = subprocess.run([_need("enkf_cuda"), cfg, "--dump", dump], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr + r.stdout
    em, ep = analysis_errors(read_dump(dump, int(g["ny"]), int(g["nx"])), g["Xa_enkf"])
    assert em < 1e-10 and ep < 1e-10, (em, ep)
    assert "EnKF diagnostics" in r.stdout


@pytest.mark.gpu
def test_letkf_driver_canonical_3d(tmp_path):
    """z_dim > 1 through the drivers: canonical mode against the oracle."""
    from metada_b200 import synthetic as syn
    from oracle import orc
    nx, ny, nz, k, P = 14, 11, 3, 8, 60
    X = syn.ensemble(k, nx, ny, nz, seed=77)
    o = syn.observations(P, nx, ny, nz, seed=78, sigma=0.25)
    tmp = str(tmp_path)
    members = []
    for m in range(k):
        p = os.path.join(tmp, f"ens_{m}.txt")
        np.savetxt(p, X[m].reshape(nz * ny, nx), fmt="%.17g")
        members.append({"state": {"variables": "simple", "file": p}})
    op = os.path.join(tmp, "obs.txt")
    with open(op, "w") as f:
        f.write("-" * 80 + "\nTime Z   Y   X   ZNU      XLONG_U    XLAT_U     Simple\n" + "-" * 80 + "\n")
        for i in range(P):
            f.write(f"0 {int(o['z'][i])} {int(o['y'][i])} {int(o['x'][i])} 0.5 0.0 0.0 {float(o['value'][i])!r}\n")
    cfg = {"logger": {"app_name": "t", "level": "error", "color": False, "console": False},
           "geometry": {"x_dim": nx, "y_dim": ny, "z_dim": nz}, "ensemble": {"members": members},
           "observations": {"types": [{"obs_A": {"if_use": True, "file": op, "coordinate": "grid",
                                                 "variables": [{"simple": {"if_use": True, "error": 0.25, "missing_value": -999.0}}]}}]},
           "obs_operator": {},
           "analysis": {"inflation": 1.0, "localization_radius": 4.0, "format": "txt", "mode": "canonical",
                        "output_base_file": os.path.join(tmp, "analysis")}}
    cp = os.path.join(tmp, "cfg.json")
    json.dump(cfg, open(cp, "w"))
    dump = os.path.join(tmp, "xa.bin")
    r = subprocess.run([_need("letkf_cuda"), cp, "--dump", dump], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr + r.stdout
    b = open(dump, "rb").read()
    Xa = np.frombuffer(b, dtype="<f8", offset=16).reshape(k, nz, ny, nx)
    ref = orc.letkf(X, o["x"], o["y"], o["z"], o["value"], np.full(P, 0.25), radius=4.0, inflation=1.0)
    em, ep = analysis_errors(Xa, ref["Xa"])
    assert em < 1e-10 and ep < 1e-10, (em, ep)


@pytest.mark.gpu
def test_host_layer_geographic_observations(tmp_path):
    """C++ host layer with the reference's Location(lat, lon, level, GEOGRAPHIC): DeviceObservations +
    DeviceEnsemble::setGeography + mdc_letkf_analyse (metada_b200/host/apps/geo_letkf_cuda.cpp) against the oracle."""
    import struct
    from metada_b200 import synthetic as syn
    from oracle import orc
    nx, ny, nz, k, P, radius = 18, 13, 3, 24, 300, 60.0
    lat, lon = syn.geography(nx, ny)
    vc = np.array([1000.0, 850.0, 500.0])
    o = syn.geo_observations(P, lat, lon, vc, seed=31)
    o["valid"][::11] = 0
    X = syn.ensemble(k, nx, ny, nz, seed=88)
    inp, out = str(tmp_path / "in.bin"), str(tmp_path / "out.bin")
    with open(inp, "wb") as f:
        f.write(struct.pack("<6qd", nx, ny, nz, k, P, len(vc), radius))
        for a in (lat, lon, vc, X, o["lat"], o["lon"], o["level"], o["value"], o["err"], o["valid"].astype(np.float64)):
            f.write(np.ascontiguousarray(a, dtype="<f8").tobytes())
    r = subprocess.run([_need("geo_letkf_cuda"), inp, out], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr + r.stdout
    Xa = np.frombuffer(open(out, "rb").read(), dtype="<f8").reshape(k, nz, ny, nx)
    ex, ey, ez = orc.geo_locate(o["lat"], o["lon"], o["level"], lat, lon, vc)
    ref = orc.letkf_ext(X, ex, ey, ez, o["value"], o["err"], o["valid"], radius=radius, glat=lat, glon=lon,
                        olat=o["lat"], olon=o["lon"])
    em, ep = analysis_errors(Xa, ref["Xa"])
    assert em < 1e-10 and ep < 1e-10, (em, ep)
    assert "columns %d" % (nx * ny) in r.stdout
    # the C++ runtime's sharded geographic path (mdc_geo_sharded_analyse) with one rank: same bits as the one-store
    # analysis (two ranks over NCCL: tools/mgpu_geo_driver_check.sh)
    out2 = str(tmp_path / "out2.bin")
    r = subprocess.run([_need("geo_letkf_cuda"), inp, out2], capture_output=True, text=True, env={**os.environ, "MDC_GEO_SHARDED": "1"})
    assert r.returncode == 0, r.stderr + r.stdout
    assert open(out2, "rb").read() == open(out, "rb").read()


@pytest.mark.gpu
@pytest.mark.parametrize("mode", ["ref_compat", "canonical"])
def test_letkf_driver_streamed_equals_one_shot(tmp_path, mode):
    """LETKF<CudaBackendTag>::Analyse through the C++ streaming runtime (analysis.streaming = "on": the members'
    host arrays go through mdc_stream_analyse in row slabs) writes the same bytes as the one-shot path."""
    exe = _need("letkf_cuda")
    g = load(CASES[1])
    dumps = []
    for streaming, extra in (("off", {}), ("on", {"slab_rows": 5})):
        d = tmp_path / streaming; d.mkdir()
        cfg = write_case(g, str(d), mode, {"streaming": streaming, **extra})
        dump = str(d / "xa.bin")
        r = subprocess.run([exe, cfg, "--dump", dump], capture_output=True, text=True)
        assert r.returncode == 0, r.stderr + r.stdout
        dumps.append(open(dump, "rb").read())
    assert dumps[0] == dumps[1]


@pytest.mark.gpu
def test_lwenkf_driver_matches_reference(tmp_path):
    """lwenkf_cuda <config> (LWEnKF<CudaBackendTag>, the reference's config keys) against the reference's own
    LWEnKF.hpp output for the adaptive-weight / polynomial-localisation case."""
    g = load(CASES[1])
    cases = json.loads(str(g["lwenkf_cases"]))
    i = 1
    weighting, fn, radius = cases[i]
    zf = str(tmp_path / "Z.bin")
    np.ascontiguousarray(g[f"lwenkf{i}_Z"], dtype="<f8").tofile(zf)
    cfg = write_case(g, str(tmp_path), "ref_compat", {"perturbation_file": zf, "weighting_scheme": weighting,
                                                       "localization_function": fn, "localization_radius": radius})
    dump = str(tmp_path / "xa.bin")
    r = subprocess.run([_need("lwenkf_cuda"), cfg, "--dump", dump], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr + r.stdout
    em, ep = analysis_errors(read_dump(dump, int(g["ny"]), int(g["nx"])), g[f"lwenkf{i}_Xa"])
    tol = max(1e-10, 1e-14 * float(g[f"lwenkf{i}_diag"][5]))
    assert em < tol and ep < tol, (em, ep)
    assert "LWEnKF diagnostics" in r.stdout


@pytest.mark.gpu
def test_ensemble_stays_device_resident_across_analyses(tmp_path):
    """Two analyses with a host-side 'forecast' of one member in between (VERDICT r1 missing 6): with
    `resident: true` (default) the first analysis uploads k members, the forecast moves one member down and
    up again (a peek at another member: one more download), the second analysis uploads just that member, and
    saveEnsemble() brings every member home once; `resident: false` re-uploads / re-downloads everything.  Both
    runs must leave byte-identical members, means and text files."""
    exe = _need("resident_cycle_cuda")
    g = load(CASES[1])
    k = int(g["k"])
    out = {}
    for resident in (True, False):
        d = tmp_path / ("res" if resident else "plain"); d.mkdir()
        cfg = write_case(g, str(d), "ref_compat", {"resident": resident, "streaming": "off"})
        dump = str(d / "xa.bin")
        r = subprocess.run([exe, cfg, "--dump", dump], capture_output=True, text=True)
        assert r.returncode == 0, r.stderr + r.stdout
        lines = {ln.split()[1]: ln.split()[2:] for ln in r.stdout.splitlines() if ln.startswith("RESIDENT ")}
        out[resident] = (open(dump, "rb").read(), lines, open(d / "analysis_mean.txt").read(), open(d / "analysis_member_1.txt").read())
    assert out[True][0] == out[False][0]                      # members, bit for bit
    assert out[True][2] == out[False][2] and out[True][3] == out[False][3]
    assert out[True][1]["peek"] == out[False][1]["peek"] and out[True][1]["mean0"] == out[False][1]["mean0"]
    ln = out[True][1]
    up = lambda key: int(ln[key][1])
    down = lambda key: int(ln[key][3])
    assert (up("analysis1"), down("analysis1")) == (k, 0)
    assert (up("forecast"), down("forecast")) == (k, 2)       # member 1 (read-modify-write) and member 0 (peek)
    assert (up("analysis2"), down("analysis2")) == (k + 1, 2) # only the member the host wrote goes up again
    assert (up("saved"), down("saved")) == (k + 1, 2 + k)     # every member comes home once, when it is read
    assert out[False][1]["analysis1"] == ["up", "-1", "down", "-1"]
