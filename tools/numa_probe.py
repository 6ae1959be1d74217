"""Host-memory placement vs PCIe throughput on the GPU box: pinned buffers allocated under different NUMA memory
policies (set_mempolicy via ctypes), H2D / D2H / both directions at once.  Writes gpurun_out/numa_probe.json."""
import ctypes
import glob
import json
import os
import subprocess
import sys
import time

import torch

libc = ctypes.CDLL(None, use_errno=True)
SYS_set_mempolicy = 238            # x86_64
MPOL_DEFAULT, MPOL_PREFERRED, MPOL_BIND, MPOL_INTERLEAVE = 0, 1, 2, 3


def set_mempolicy(mode, nodes):
    mask = ctypes.c_ulong(sum(1 << n for n in nodes))
    rc = libc.syscall(SYS_set_mempolicy, mode, ctypes.byref(mask) if nodes else None, 65 if nodes else 0)
    return rc, ctypes.get_errno()


def sh(cmd):
    try:
        return subprocess.check_output(cmd, shell=True, text=True, stderr=subprocess.STDOUT).strip()
    except Exception as e:  # noqa: BLE001
        return f"ERR {e}"


def measure(nbytes):
    n = nbytes // 8
    h1 = torch.empty(n, dtype=torch.float64, pin_memory=True)
    h2 = torch.empty(n, dtype=torch.float64, pin_memory=True)
    h1.zero_(); h2.zero_()
    d1 = torch.empty(n, dtype=torch.float64, device="cuda")
    d2 = torch.empty(n, dtype=torch.float64, device="cuda")
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    res = {}

    def timed(fn, reps=3):
        fn(); torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(reps):
            fn()
        torch.cuda.synchronize()
        return (time.perf_counter() - t0) / reps

    res["h2d_GBs"] = nbytes / timed(lambda: d1.copy_(h1, non_blocking=True)) / 1e9
    res["d2h_GBs"] = nbytes / timed(lambda: h2.copy_(d2, non_blocking=True)) / 1e9

    def both():
        with torch.cuda.stream(s1):
            d1.copy_(h1, non_blocking=True)
        with torch.cuda.stream(s2):
            h2.copy_(d2, non_blocking=True)
    res["bidir_total_GBs"] = 2 * nbytes / timed(both) / 1e9
    del h1, h2, d1, d2
    torch.cuda.empty_cache()
    return res


def main():
    out = {"nproc": os.cpu_count(), "lscpu_numa": sh("lscpu | grep -i numa"), "topo": sh("nvidia-smi topo -m"),
           "gpu_bus": sh("nvidia-smi --query-gpu=pci.bus_id --format=csv,noheader")}
    bus = out["gpu_bus"].splitlines()[0].strip().lower()
    bus = bus[4:] if len(bus.split(":")[0]) == 8 else bus       # 00000000:1B:00.0 -> 0000:1b:00.0
    out["gpu_numa_node"] = sh(f"cat /sys/bus/pci/devices/{bus}/numa_node")
    nodes = sorted(int(p.rsplit("node", 1)[1]) for p in glob.glob("/sys/devices/system/node/node[0-9]*"))
    out["nodes"] = nodes
    out["node_mem"] = {n: sh(f"grep -E 'MemTotal|MemFree' /sys/devices/system/node/node{n}/meminfo") for n in nodes}
    out["affinity"] = sorted(os.sched_getaffinity(0))[:4] + ["...", len(os.sched_getaffinity(0))]
    torch.cuda.init()
    nbytes = 4 << 30
    runs = [("default", MPOL_DEFAULT, [])] + [(f"bind{n}", MPOL_BIND, [n]) for n in nodes[:4]]
    if len(nodes) > 1:
        runs.append(("interleave", MPOL_INTERLEAVE, nodes))
    out["runs"] = {}
    for name, mode, ns in runs:
        rc, err = set_mempolicy(mode, ns)
        if rc != 0:
            out["runs"][name] = {"error": f"set_mempolicy rc={rc} errno={err}"}
            continue
        try:
            out["runs"][name] = measure(nbytes)
        except Exception as e:  # noqa: BLE001
            out["runs"][name] = {"error": str(e)}
        print(name, out["runs"][name], flush=True)
    set_mempolicy(MPOL_DEFAULT, [])
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(out, open("gpurun_out/numa_probe.json", "w"), indent=1)
    print(json.dumps({k: out[k] for k in ("lscpu_numa", "gpu_numa_node", "nodes", "affinity")}))


if __name__ == "__main__":
    sys.exit(main())
