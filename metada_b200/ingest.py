"""Observation ingest from the reference's column-oriented Parquet files straight into the device SoA.

The reference's ``ParquetObservation`` (framework/adapters/common/io/ParquetObservation.hpp:29-147,
ParquetObservation.cpp:211-330 ``saveToParquet``) writes one row per observation with the metadata columns

    latitude, longitude, pressure, time (float64, unix seconds), obs_type, channel, qc_flag (int32),
    station_id, report_type, instrument_type (utf8), elevation, obs_error (float64)

followed by one column per observed field (float64 / int32 / utf8, e.g. ``temperature``).  A GEOGRAPHIC
observation of the analysis is (latitude, longitude, level = pressure, value = the field, err = obs_error); rows
whose value is missing (null / NaN / ``missing_value``) or whose qc_flag exceeds ``qc_max`` are kept but marked
invalid, as the reference's backends do for missing values (GridObservation.hpp:239-252: infinite variance).
Arrow hands the columns over as contiguous buffers, which go to the device with one copy each
(``mdc_obs_create_geographic``) -- no per-observation objects on the way.

Host-side plumbing only (pyarrow); the filters of ParquetObservation (filterByType / Category / Channel / QC /
PressureRange, :94-98) are the keyword arguments.
"""
from __future__ import annotations

import numpy as np

META_COLUMNS = ("latitude", "longitude", "pressure", "time", "obs_type", "channel", "qc_flag", "station_id",
                "report_type", "instrument_type", "elevation", "obs_error")


def read_parquet_observations(path, field, *, qc_max=None, obs_type=None, channel=None, pressure_range=None,
                              missing_value=None, default_error=None):
    """Returns dict(lat, lon, level, value, err, valid, time, obs_type) of float64 / uint8 / int32 arrays."""
    import pyarrow.parquet as pq
    cols = [c for c in ("latitude", "longitude", "pressure", "time", "obs_type", "channel", "qc_flag", "obs_error", field)]
    schema = pq.read_schema(path)
    missing = [c for c in ("latitude", "longitude", field) if c not in schema.names]
    if missing:
        raise KeyError(f"{path}: columns {missing} not found (has {schema.names})")
    tab = pq.read_table(path, columns=[c for c in cols if c in schema.names])
    n = tab.num_rows

    def col(name, dtype, fill):
        if name not in tab.column_names:
            return np.full(n, fill, dtype=dtype)
        a = tab.column(name).combine_chunks()
        if a.null_count:
            a = a.fill_null(fill)
        return np.ascontiguousarray(a.to_numpy(zero_copy_only=False), dtype=dtype)

    raw = tab.column(field).combine_chunks()
    valid = np.ones(n, dtype=bool) if raw.null_count == 0 else ~np.asarray(raw.is_null())
    value = np.ascontiguousarray(raw.fill_null(0).to_numpy(zero_copy_only=False), dtype=np.float64)
    valid &= np.isfinite(value)
    if missing_value is not None:
        valid &= value != missing_value
    qc = col("qc_flag", np.int32, 0)
    if qc_max is not None:
        valid &= qc <= qc_max
    keep = np.ones(n, dtype=bool)                      # filters DROP rows (ParquetObservation::filterBy*)
    typ, chan, pres = col("obs_type", np.int32, 0), col("channel", np.int32, 0), col("pressure", np.float64, 0.0)
    if obs_type is not None:
        keep &= typ == obs_type
    if channel is not None:
        keep &= chan == channel
    if pressure_range is not None:
        keep &= (pres >= pressure_range[0]) & (pres <= pressure_range[1])
    err = col("obs_error", np.float64, np.nan)
    if default_error is not None:
        err = np.where(np.isfinite(err) & (err > 0), err, default_error)
    valid &= np.isfinite(err) & (err > 0)
    err = np.where(valid, err, 1.0)                    # a placeholder the kernels never use (invalid => weight 0)
    value = np.where(valid, value, 0.0)
    out = {"lat": col("latitude", np.float64, np.nan), "lon": col("longitude", np.float64, np.nan), "level": pres,
           "value": value, "err": err, "valid": valid.astype(np.uint8), "time": col("time", np.float64, 0.0),
           "obs_type": typ}
    bad = ~(np.isfinite(out["lat"]) & np.isfinite(out["lon"]) & (np.abs(out["lat"]) <= 90.0))
    keep &= ~bad                                       # a row without a position cannot be placed at all
    return {k: np.ascontiguousarray(v[keep]) for k, v in out.items()}


def observations_from_parquet(ctx, path, field, **kw):
    """Device observation store (GEOGRAPHIC locations) from a ParquetObservation file."""
    from .capi import Observations
    o = read_parquet_observations(path, field, **kw)
    return Observations.geographic(ctx, o["lat"], o["lon"], o["level"], o["value"], o["err"], o["valid"]), o


def write_parquet_observations(path, lat, lon, pressure, field, value, obs_error, *, time=None, obs_type=None,
                               channel=None, qc_flag=None, station_id=None, report_type="SYNTH",
                               instrument_type="SYNTH", elevation=None):
    """Writes the reference's schema (ParquetObservation.cpp:211-330) -- used by the tests and the synthetic tools."""
    import pyarrow as pa
    import pyarrow.parquet as pq
    n = len(lat)
    z64, z32 = np.zeros(n), np.zeros(n, np.int32)
    tab = pa.table({
        "latitude": np.asarray(lat, np.float64), "longitude": np.asarray(lon, np.float64),
        "pressure": np.asarray(pressure, np.float64), "time": np.asarray(time if time is not None else z64, np.float64),
        "obs_type": np.asarray(obs_type if obs_type is not None else z32, np.int32),
        "channel": np.asarray(channel if channel is not None else z32, np.int32),
        "qc_flag": np.asarray(qc_flag if qc_flag is not None else z32, np.int32),
        "station_id": pa.array(station_id if station_id is not None else [f"S{i:06d}" for i in range(n)]),
        "report_type": pa.array([report_type] * n), "instrument_type": pa.array([instrument_type] * n),
        "elevation": np.asarray(elevation if elevation is not None else z64, np.float64),
        "obs_error": np.asarray(obs_error, np.float64),
        field: pa.array(np.asarray(value, np.float64), mask=~np.isfinite(np.asarray(value, np.float64))),
    })
    pq.write_table(tab, path)
