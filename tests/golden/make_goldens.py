"""Generates tests/golden/*.npz by running the reference's own headers (oracle/_ref/ref_driver,
built by `make -C oracle ref` from /root/reference + the Eigen-API shim) on
  (1) the reference's tutorial data set (src/backends/simple/tutorial: 36x18 grid, 9 members, 28 obs),
  (2) a seeded synthetic 2-D case written in the Simple backend's text formats.
Run in the build container only (needs /root/reference); the .npz files are committed.
"""
import json
import os
import struct
import subprocess
import sys
import tempfile

import numpy as np
import yaml

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from metada_b200 import synthetic as syn  # noqa: E402

REF = "/root/reference/src/backends/simple/tutorial"
DRIVER = os.path.join(ROOT, "oracle", "_ref", "ref_driver")
OUT = os.path.dirname(os.path.abspath(__file__))


# (weighting_scheme, localization_function, localization_radius): the radius normalises LWEnKF's INDEX distance
# |i - j| / dim (LWEnKF.hpp:569, 593), so values below 1 are what makes the localisation bite
LWENKF_CASES = [("uniform", "gaussian", 0.3), ("adaptive", "gaspari_cohn", 0.4), ("inverse_var", "exponential", 0.25),
                ("likelihood", "cutoff", 0.5), ("uniform", "gaussian", 10.0)]


def read_dump(path):
    b = open(path, "rb").read()
    k, n, P = struct.unpack_from("<qqq", b, 0)
    off, vecs = 24, []
    while off < len(b):
        (m,) = struct.unpack_from("<q", b, off)
        off += 8
        vecs.append(np.frombuffer(b, dtype="<f8", count=m, offset=off).copy())
        off += 8 * m
    return k, n, P, vecs


def run(mode, cfg_path, tmp):
    out = os.path.join(tmp, f"{mode}.bin")
    subprocess.check_call([DRIVER, mode, cfg_path, out], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    return read_dump(out)


def collect(cfg, tmp, nx, ny):
    cfg_path = os.path.join(tmp, "cfg.json")
    json.dump(cfg, open(cfg_path, "w"))
    g = {}
    k, n, P, v = run("hx", cfg_path, tmp)
    g["k"], g["nx"], g["ny"], g["P"] = k, nx, ny, P
    g["HX"] = np.stack(v[:k])                       # [k][P]
    g["yo"], g["var"] = v[k], v[k + 1]
    g["ox"], g["oy"], g["oz"] = (v[k + 2 + i].astype(np.int32) for i in range(3))
    g["counts"] = v[k + 5].astype(np.int32).reshape(ny, nx)
    g["radius"], g["inflation"] = v[k + 6]
    g["mean_b"] = v[k + 7].reshape(ny, nx)
    for mode in ("letkf", "letkf_snapshot", "etkf", "enkf"):
        _, _, _, v = run(mode, cfg_path, tmp)
        g[f"Xa_{mode}"] = v[0].reshape(k, 1, ny, nx)
        if mode.startswith("letkf"):
            g[f"mean_{mode}"] = v[1].reshape(ny, nx)
        if mode == "enkf":
            g["enkf_Z"] = v[1].reshape(P, k)
            g["enkf_diag"] = v[2]
    # LWEnKF<SimpleBackendTag>::Analyse (LWEnKF.hpp:207-334) for every weighting scheme / localisation function
    for i, (weighting, locfn, radius) in enumerate(LWENKF_CASES):
        c = json.loads(json.dumps(cfg))
        c["analysis"].update(weighting_scheme=weighting, localization_function=locfn, localization_radius=radius)
        cp = os.path.join(tmp, f"cfg_lw{i}.json")
        json.dump(c, open(cp, "w"))
        _, _, _, v = run("lwenkf", cp, tmp)
        g[f"lwenkf{i}_Xa"] = v[0].reshape(k, 1, ny, nx)
        g[f"lwenkf{i}_Z"] = v[1].reshape(P, k)
        g[f"lwenkf{i}_diag"] = v[2]        # innovation_norm, background_spread, analysis_spread, max/min K, cond, max/min w, var w
    g["lwenkf_cases"] = np.array(json.dumps(LWENKF_CASES))
    return g


def tutorial(tmp):
    cfg = yaml.safe_load(open(os.path.join(REF, "letkf.yaml")))
    for m in cfg["ensemble"]["members"]:
        m["state"]["file"] = os.path.join(REF, os.path.basename(m["state"]["file"]))
    for t in cfg["observations"]["types"]:
        for name, tc in t.items():
            tc["file"] = os.path.join(REF, os.path.basename(tc["file"]))
    cfg["logger"].update(level="error", console=False)
    cfg["analysis"].update(output_base_file=os.path.join(tmp, "analysis"), format="txt",
                           inflation_method="multiplicative")
    nx, ny = cfg["geometry"]["x_dim"], cfg["geometry"]["y_dim"]
    g = collect(cfg, tmp, nx, ny)
    X = np.stack([np.loadtxt(m["state"]["file"]).reshape(1, ny, nx) for m in cfg["ensemble"]["members"]])
    g["X"] = X
    g["err_cfg"] = 0.1     # config value; the reference narrows it to float (ConfigValue.hpp:108)
    return g


def synthetic_case(tmp, nx=23, ny=17, k=8, P=70, radius=4.5, inflation=1.1, sigma=0.25, seed=5):
    X = syn.ensemble(k, nx, ny, 1, seed=1000 + seed)
    o = syn.observations(P, nx, ny, 1, seed=42 + seed, sigma=sigma)
    members = []
    for m in range(k):
        p = os.path.join(tmp, f"syn_ens_{m}.txt")
        with open(p, "w") as f:
            for y in range(ny):
                f.write(" ".join(repr(float(v)) for v in X[m, 0, y]) + "\n")
        members.append({"state": {"variables": "simple", "file": p}})
    op = os.path.join(tmp, "syn_obs.txt")
    with open(op, "w") as f:
        f.write("-" * 80 + "\nTime Z   Y   X   ZNU      XLONG_U    XLAT_U     Simple\n" + "-" * 80 + "\n")
        for i in range(P):
            f.write(f"0 {int(o['z'][i])} {int(o['y'][i])} {int(o['x'][i])} 0.5 0.0 0.0 {float(o['value'][i])!r}\n")
    cfg = {"logger": {"app_name": "golden", "level": "error", "color": False, "console": False},
           "geometry": {"x_dim": nx, "y_dim": ny},
           "ensemble": {"members": members},
           "observations": {"types": [{"obs_A": {"if_use": True, "file": op, "coordinate": "grid",
                                                 "variables": [{"simple": {"if_use": True, "error": sigma, "missing_value": -999.0}}]}}]},
           "obs_operator": {},
           "analysis": {"algorithm": "letkf", "inflation": inflation, "localization_radius": radius,
                        "inflation_method": "multiplicative", "format": "txt",
                        "output_base_file": os.path.join(tmp, "analysis")}}
    g = collect(cfg, tmp, nx, ny)
    g["X"] = X
    g["err_cfg"] = sigma
    return g


def metrics_case(tmp):
    """The reference's Metrics.hpp (oracle/_ref/ref_metrics) on a seeded ensemble + truth."""
    k, nx, ny, nz = 7, 11, 6, 3
    X = syn.ensemble(k, nx, ny, nz, seed=4242)                       # [k][nz][ny][nx]
    truth = syn.ensemble(3, nx, ny, nz, seed=99).mean(0)             # an independent smooth field
    n = nx * ny * nz
    inp, out = os.path.join(tmp, "m_in.bin"), os.path.join(tmp, "m_out.bin")
    with open(inp, "wb") as f:
        f.write(struct.pack("<qq", k, n))
        f.write(np.ascontiguousarray(X.reshape(k, n)).tobytes())
        f.write(np.ascontiguousarray(truth.reshape(n)).tobytes())
    subprocess.check_call([os.path.join(ROOT, "oracle", "_ref", "ref_metrics"), inp, out])
    b = np.frombuffer(open(out, "rb").read(), dtype="<f8")
    return {"X": X, "truth": truth.reshape(nz, ny, nx), "scalars": b[:5].copy(),
            "mean": b[5:5 + n].reshape(nz, ny, nx).copy(), "spread": b[5 + n:5 + 2 * n].reshape(nz, ny, nx).copy()}


def location_case(tmp):
    """The reference's Location::distance_to (oracle/_ref/ref_location) on seeded GEOGRAPHIC pairs: regional pairs
    (the LETKF selection regime), global pairs, dateline crossings, identical and antipodal points."""
    rng = np.random.default_rng(2024)
    n = 4000
    a = np.empty((n, 4))
    a[:, 0] = rng.uniform(-89, 89, n); a[:, 1] = rng.uniform(-180, 180, n)
    a[:1500, 2] = a[:1500, 0] + rng.uniform(-1.5, 1.5, 1500)            # regional: within ~200 km
    a[:1500, 3] = a[:1500, 1] + rng.uniform(-2.0, 2.0, 1500)
    a[1500:, 2] = rng.uniform(-90, 90, n - 1500); a[1500:, 3] = rng.uniform(-180, 180, n - 1500)
    a[:, 2] = np.clip(a[:, 2], -90, 90)
    a[3000:3200, 1] = rng.uniform(178, 180, 200); a[3000:3200, 3] = rng.uniform(-180, -178, 200)   # dateline
    a[3200:3300, 2:] = a[3200:3300, :2]                                   # identical points
    a[3300:3400, 2] = -a[3300:3400, 0]; a[3300:3400, 3] = a[3300:3400, 1] + 180.0   # antipodes
    inp, out = os.path.join(tmp, "l_in.bin"), os.path.join(tmp, "l_out.bin")
    with open(inp, "wb") as f:
        f.write(struct.pack("<q", n))
        f.write(np.ascontiguousarray(a).tobytes())
    subprocess.check_call([os.path.join(ROOT, "oracle", "_ref", "ref_location"), inp, out])
    # CARTESIAN pairs (Location.hpp:217-225): mixed magnitudes, identical points, one coordinate differing
    m = 1000
    c = rng.uniform(-1.0, 1.0, (m, 6)) * 10.0 ** rng.integers(-3, 7, (m, 1))
    c[800:850, 3:] = c[800:850, :3]
    c[850:900, 3:5] = c[850:900, 0:2]
    cin, cout = os.path.join(tmp, "c_in.bin"), os.path.join(tmp, "c_out.bin")
    with open(cin, "wb") as f:
        f.write(struct.pack("<q", m))
        f.write(np.ascontiguousarray(c).tobytes())
    subprocess.check_call([os.path.join(ROOT, "oracle", "_ref", "ref_location"), cin, cout, "cartesian"])
    return {"pairs": a, "km": np.frombuffer(open(out, "rb").read(), dtype="<f8").copy(),
            "cart_pairs": c, "cart_dist": np.frombuffer(open(cout, "rb").read(), dtype="<f8").copy()}


def obsop_geo_case(tmp, var_nlev=(5, 5, 1)):
    """The reference's IdentityObsOperator (oracle/_ref/ref_obsop_geo, mock WRF-type backends) on GEOGRAPHIC
    observations of a three-variable state ([5, 5, 1] levels) on a curvilinear grid: nearest grid point / level
    (convertGeographicToGrid) + 4-of-8 IDW in the observation's own variable."""
    nx, ny, var_nlev, P = 23, 17, list(var_nlev), 600
    nz = min(v for v in var_nlev if v > 1)        # the geometry's (mass) levels; a W-staggered variable has nz + 1
    lat, lon = syn.geography(nx, ny)
    vc = np.array([1000.0, 925.0, 850.0, 700.0, 500.0])
    o = syn.geo_observations(P, lat, lon, vc, seed=77, margin=0.3)
    rng = np.random.default_rng(5)
    ovar = rng.integers(0, 3, P)
    valid = (rng.random(P) > 0.05)
    state = syn.ensemble(1, nx, ny, sum(var_nlev), seed=31)[0]              # [11][ny][nx]
    # a few observations exactly on grid points and exactly half way between two (ties)
    o["lat"][:5], o["lon"][:5] = lat[3, 4:9], lon[3, 4:9]
    o["lat"][5:9], o["lon"][5:9] = 0.5 * (lat[6, 2:6] + lat[6, 3:7]), 0.5 * (lon[6, 2:6] + lon[6, 3:7])
    inp, out = os.path.join(tmp, "h_in.bin"), os.path.join(tmp, "h_out.bin")
    with open(inp, "wb") as f:
        f.write(struct.pack("<5q", nx, ny, nz, len(var_nlev), P))
        f.write(struct.pack("<%dq" % len(var_nlev), *var_nlev))
        for a in (lat, lon, vc, state, o["lat"], o["lon"], o["level"], ovar.astype(np.float64), valid.astype(np.float64)):
            f.write(np.ascontiguousarray(a, dtype="<f8").tobytes())
    subprocess.check_call([os.path.join(ROOT, "oracle", "_ref", "ref_obsop_geo"), inp, out])
    return {"lat": lat, "lon": lon, "vc": vc, "state": state, "var_nlev": np.array(var_nlev, np.int32),
            "olat": o["lat"], "olon": o["lon"], "olev": o["level"], "ovar": ovar.astype(np.int32),
            "valid": valid.astype(np.uint8), "HX": np.frombuffer(open(out, "rb").read(), dtype="<f8").copy()}


def main():
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "ref"], stdout=subprocess.DEVNULL)
    with tempfile.TemporaryDirectory() as tmp:
        np.savez_compressed(os.path.join(OUT, "metrics_7x11x6x3.npz"), **metrics_case(tmp))
    with tempfile.TemporaryDirectory() as tmp:
        np.savez_compressed(os.path.join(OUT, "location_geographic.npz"), **location_case(tmp))
    with tempfile.TemporaryDirectory() as tmp:
        np.savez_compressed(os.path.join(OUT, "obsop_geographic.npz"), **obsop_geo_case(tmp))
    with tempfile.TemporaryDirectory() as tmp:    # second variable staggered in the vertical (WRF's W: nz + 1 levels)
        np.savez_compressed(os.path.join(OUT, "obsop_geographic_wstag.npz"), **obsop_geo_case(tmp, (5, 6, 1)))
    with tempfile.TemporaryDirectory() as tmp:
        np.savez_compressed(os.path.join(OUT, "tutorial_36x18.npz"), **tutorial(tmp))
    with tempfile.TemporaryDirectory() as tmp:
        np.savez_compressed(os.path.join(OUT, "synthetic_23x17.npz"), **synthetic_case(tmp))
    for f in ("tutorial_36x18.npz", "synthetic_23x17.npz", "metrics_7x11x6x3.npz", "location_geographic.npz", "obsop_geographic.npz"):
        g = np.load(os.path.join(OUT, f))
        print(f, {k: (g[k].shape if g[k].ndim else g[k].item()) for k in g.files})


if __name__ == "__main__":
    main()
