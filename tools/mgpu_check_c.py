"""Multi-GPU check of the C++ runtime (csrc/mdc_runtime.cpp): every rank streams its rows through mdc_stream_analyse with
the NCCL observation halo (mdc_comm_init), and the assembled analysis is compared bit for bit with the one-shot analysis
of the whole grid on rank 0.  Launch: torchrun --nproc-per-node N tools/mgpu_check_c.py  (torch.distributed only carries
the NCCL unique id and gathers the result)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist

import metada_b200 as mb
from metada_b200 import capi
from metada_b200.parallel import slab_bounds
from tests.common import make_case


def main():
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dist.init_process_group("gloo")
    nx, ny, nz, k, P, radius = 37, 61, 3, 40, 1300, 5.0
    X, o = make_case(nx, ny, nz, k, P, seed=11, out_of_grid=8, invalid_frac=0.03)
    params = capi.make_params(radius, 1.03, mb.MODE_CANONICAL, mb.LOC_GASPARI_COHN)
    y0, y1 = slab_bounds(ny, rank, world)
    halo = 1 if y1 < ny else 0
    host = np.ascontiguousarray(X[:, :, y0:y1 + halo, :])
    sl = mb.Stream(local, nx, ny, nz, k, radius, row_range=(y0, y1), slab_rows=5, slots=3)
    uid = [mb.Stream.unique_id() if rank == 0 else None]
    dist.broadcast_object_list(uid, 0)
    if world > 1:
        sl.comm_init(uid[0], rank, world)
    st = sl.analyse([host[m].ctypes.data for m in range(k)], o, params, host_row0=y0, host_ny=host.shape[2])
    tm = sl.timings()
    mx = sl.comm_max(float(rank))
    parts = [None] * world
    dist.all_gather_object(parts, (y0, y1, host[:, :, :y1 - y0, :]))
    if rank == 0:
        out = np.empty_like(X)
        for a, b, h in parts:
            out[:, :, a:b, :] = h
        ctx = mb.Context(local)
        ens = mb.Ensemble(ctx, nx, ny, nz, k); ens.upload(X)
        obs = mb.Observations(ctx, o["x"], o["y"], o["z"], o["value"], o["err"], o["valid"])
        capi.letkf_analyse(ens, obs, params)
        ref = ens.download()
        print({"world": world, "bit_identical": bool(np.array_equal(out, ref)), "max_abs_diff": float(np.abs(out - ref).max()),
               "rank0": st["columns"], "timings_rank0": tm, "comm_max": mx})
        assert np.array_equal(out, ref) and mx == world - 1
    sl.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
