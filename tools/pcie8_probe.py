"""Aggregate host <-> device ceiling of the box with N ranks copying at once (torchrun --nproc-per-node N): every rank
moves 4 GiB pinned -> device and device -> pinned simultaneously on two streams, all ranks start together.
Rank 0 prints one JSON line: per-rank and aggregate GB/s for H2D alone, D2H alone and both directions at once."""
import json
import os
import time

import torch
import torch.distributed as dist

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n = 1 << 28                       # 2 GiB of float64 per buffer
h_up = torch.empty(n, dtype=torch.float64, pin_memory=True)
h_dn = torch.empty(n, dtype=torch.float64, pin_memory=True)
d_up = torch.empty(n, dtype=torch.float64, device="cuda")
d_dn = torch.empty(n, dtype=torch.float64, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def run(mode, reps=3):
    def once():
        if mode in ("h2d", "both"):
            with torch.cuda.stream(s1):
                d_up.copy_(h_up, non_blocking=True)
        if mode in ("d2h", "both"):
            with torch.cuda.stream(s2):
                h_dn.copy_(d_dn, non_blocking=True)
    once(); torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        once()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    t = torch.tensor([dt], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    nbytes = n * 8 * reps * (2 if mode == "both" else 1)
    return nbytes / float(t[0]) / 1e9


out = {"ranks": world, "per_rank_GBps": {}, "aggregate_GBps": {}}
for mode in ("h2d", "d2h", "both"):
    v = run(mode)
    out["per_rank_GBps"][mode] = v
    out["aggregate_GBps"][mode] = v * world
if rank == 0:
    try:
        out["numa_nodes"] = len([d for d in os.listdir("/sys/devices/system/node") if d.startswith("node")])
    except Exception:  # noqa: BLE001
        pass
    out["host_cores"] = os.cpu_count()
    print(json.dumps(out), flush=True)
if world > 1:
    dist.destroy_process_group()
