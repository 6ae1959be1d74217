"""Parquet observation ingest: the reference's ParquetObservation file schema
(framework/adapters/common/io/ParquetObservation.cpp:211-330) read with pyarrow into the arrays
mdc_obs_create_geographic takes."""
import numpy as np
import pytest

from metada_b200 import ingest, synthetic as syn


def _write(tmp_path, n=500, seed=3):
    lat, lon = syn.geography(20, 15)
    o = syn.geo_observations(n, lat, lon, np.array([1000.0, 850.0, 500.0]), seed=seed)
    rng = np.random.default_rng(seed)
    value = o["value"].copy()
    value[::17] = np.nan                               # missing values -> nulls in the file
    qc = rng.integers(0, 4, n).astype(np.int32)
    typ = rng.choice([120, 181, 220], n).astype(np.int32)
    path = str(tmp_path / "obs.parquet")
    ingest.write_parquet_observations(path, o["lat"], o["lon"], o["level"], "temperature", value, o["err"],
                                      qc_flag=qc, obs_type=typ, time=np.full(n, 1.7e9))
    return path, o, value, qc, typ


def test_parquet_schema_and_roundtrip(tmp_path):
    import pyarrow.parquet as pq
    path, o, value, qc, typ = _write(tmp_path)
    names = pq.read_schema(path).names
    assert names[:12] == list(ingest.META_COLUMNS) and names[12] == "temperature"
    r = ingest.read_parquet_observations(path, "temperature")
    assert np.array_equal(r["lat"], o["lat"]) and np.array_equal(r["lon"], o["lon"]) and np.array_equal(r["level"], o["level"])
    ok = np.isfinite(value)
    assert np.array_equal(r["valid"], ok.astype(np.uint8))
    assert np.array_equal(r["value"][ok], value[ok]) and (r["value"][~ok] == 0).all()
    assert np.array_equal(r["err"][ok], o["err"][ok])
    assert r["lat"].flags.c_contiguous and r["value"].dtype == np.float64 and r["valid"].dtype == np.uint8


def test_parquet_filters_and_quality_control(tmp_path):
    path, o, value, qc, typ = _write(tmp_path)
    r = ingest.read_parquet_observations(path, "temperature", qc_max=1)
    assert np.array_equal(r["valid"], (np.isfinite(value) & (qc <= 1)).astype(np.uint8))      # QC marks, does not drop
    r = ingest.read_parquet_observations(path, "temperature", obs_type=181, pressure_range=(400.0, 900.0))
    keep = (typ == 181) & (o["level"] >= 400.0) & (o["level"] <= 900.0)
    assert len(r["lat"]) == int(keep.sum()) and np.array_equal(r["lat"], o["lat"][keep])
    with pytest.raises(KeyError):
        ingest.read_parquet_observations(path, "humidity")


@pytest.mark.gpu
def test_parquet_to_device_store_and_analysis(ctx, tmp_path):
    import metada_b200 as mb
    from metada_b200 import capi
    from oracle import orc
    from tests.common import analysis_errors
    path, o, value, qc, typ = _write(tmp_path, n=400, seed=8)
    lat, lon = syn.geography(20, 15)
    vc = np.array([1000.0, 850.0, 500.0])
    X = syn.ensemble(24, 20, 15, 3, seed=55)
    ens = mb.Ensemble(ctx, 20, 15, 3, 24)
    ens.upload(X)
    ens.set_geography(lat, lon, vc)
    obs, r = ingest.observations_from_parquet(ctx, path, "temperature", qc_max=2)
    assert r["valid"].sum() < len(r["valid"])
    capi.letkf_analyse(ens, obs, capi.make_params(60.0, 1.0, mb.MODE_CANONICAL, mb.LOC_GASPARI_COHN))
    ex, ey, ez = orc.geo_locate(r["lat"], r["lon"], r["level"], lat, lon, vc)
    ref = orc.letkf_ext(X, ex, ey, ez, r["value"], r["err"], r["valid"], radius=60.0, glat=lat, glon=lon,
                        olat=r["lat"], olon=r["lon"])
    em, ep = analysis_errors(ens.download(), ref["Xa"])
    assert em < 1e-10 and ep < 1e-10, (em, ep)
    ens.close(); obs.close()
