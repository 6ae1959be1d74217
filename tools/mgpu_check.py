"""torchrun --nproc-per-node N tools/mgpu_check.py : N-GPU column-sharded LETKF vs the CPU oracle."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
import metada_b200 as mb
from metada_b200 import capi, synthetic as syn
from metada_b200.parallel import SlabLetkf, slab_bounds
rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
nx, ny, nz, k, P, radius = 40, 37, 3, 20, 500, 5.0
X = syn.ensemble(k, nx, ny, nz, seed=1000)
o = syn.observations(P, nx, ny, nz, seed=42)
o["y"][:4] = [-1, ny, ny - 1, 0]
ctx = mb.Context(lr)
job = SlabLetkf(ctx, nx, ny, nz, k, rank, world, radius)
y0, y1 = slab_bounds(ny, rank, world)
job.ens.upload(np.ascontiguousarray(X[:, :, y0:y0 + job.ny_loc, :]))
job.set_observations(o)
st = job.analyse(capi.make_params(radius, 1.0, mb.MODE_CANONICAL, mb.LOC_GASPARI_COHN))
Xa_loc = job.ens.download()[:, :, : y1 - y0, :]
from oracle import orc
ref = orc.letkf(X, o["x"], o["y"], o["z"], o["value"], o["err"], radius=radius)["Xa"][:, :, y0:y1, :]
err = np.abs(Xa_loc - ref).max() / np.abs(ref).max()
t = torch.tensor([err], device="cuda", dtype=torch.float64)
dist.all_reduce(t, op=dist.ReduceOp.MAX)
cols = torch.tensor([st["columns"], job.halo_rows_last], device="cuda", dtype=torch.int64)
dist.all_reduce(cols)
if rank == 0:
    print(f"mgpu_check world={world}: max rel err vs oracle {float(t[0]):.3e}, columns {int(cols[0])} (expect {nx*ny}), halo rows exchanged {int(cols[1])}")
    assert float(t[0]) < 1e-10 and int(cols[0]) == nx * ny
job.close(); ctx.close(); dist.destroy_process_group()
