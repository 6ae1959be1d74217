"""Sharded GEOGRAPHIC analysis on real ranks (torchrun, one process per GPU, NCCL): every rank analyses its slab with
the halo rows received over torch.distributed, rank 0 also runs the one-store analysis and compares bit for bit.
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 tools/mgpu_geo_check.py"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist

import metada_b200 as mb
from metada_b200 import capi, synthetic as syn
from metada_b200.parallel import GeoSlabLetkf

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", rank)))
dist.init_process_group("nccl")
nx, ny, nz, k, P, radius = 120, 96, 12, 40, 9000, 60.0
var_nlev = [5, 6, 1]
vc = np.array([1000.0, 925.0, 850.0, 700.0, 500.0])
lat, lon = syn.geography(nx, ny, lon0=176.0)
o = dict(syn.geo_observations(P, lat, lon, vc, seed=21))
o["var"] = np.random.default_rng(4).integers(0, 3, P).astype(np.int32)
X = syn.ensemble(k, nx, ny, nz, seed=22)
ctx = mb.Context(torch.cuda.current_device())
params = capi.make_params(radius, 1.03, mb.MODE_CANONICAL, mb.LOC_GASPARI_COHN)
job = GeoSlabLetkf(ctx, lat, lon, vc, nz, k, rank, world, radius, var_nlev)
job.ens.upload(np.ascontiguousarray(X[:, :, job.y0:job.y0 + job.ny_loc, :]))
job.set_observations(o)
st = job.analyse(params)
mine = job.ens.download()[:, :, :job.y1 - job.y0, :]
halo = job.halo_rows_last
job.close()
# gather the slabs on rank 0 (through the host: this is a check, not a benchmark)
parts = [None] * world
dist.all_gather_object(parts, (job.y0, job.y1, mine))
if rank == 0:
    ens = mb.Ensemble(ctx, nx, ny, nz, k)
    ens.upload(X)
    ens.set_geography(lat, lon, vc)
    ens.set_variables(var_nlev)
    obs = mb.Observations.geographic(ctx, o["lat"], o["lon"], o["level"], o["value"], o["err"], o["valid"])
    obs.set_variables(o["var"])
    st1 = capi.letkf_analyse(ens, obs, params)
    one = ens.download()
    out = np.empty_like(one)
    for y0, y1, a in parts:
        out[:, :, y0:y1, :] = a
    print(json.dumps({"world": world, "bit_identical": bool(np.array_equal(out, one)), "max_abs_diff": float(np.abs(out - one).max()),
                      "columns": nx * ny, "halo_rows_rank0": halo, "mean_local_obs": st1["sum_local_obs"] / st1["columns"],
                      "changed": float(np.abs(one - X).max())}))
    ens.close(); obs.close()
ctx.close()
dist.barrier()
dist.destroy_process_group()
