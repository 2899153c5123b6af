#!/usr/bin/env python
"""pcie_probe.py -- host<->device copy bandwidth per rank, all ranks at once (what bounds bench.py's `e2e`).

  python tools/pcie_probe.py                          # 1 GPU
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/pcie_probe.py

Every rank copies the byte volume a 20k x 20k solve moves for its band (or --mb) H2D and D2H, alone and
with every other rank copying at the same time, from two kinds of pinned memory:
  default : cudaHostAlloc / torch pin_memory -- pages land wherever the allocating thread runs
  near    : b200_host_alloc_near -- pages first-touched on the NUMA node the GPU hangs off
Rank 0 prints one JSON object: per rank the GPU's NUMA node, the CPU set and memory nodes the process may
use, and GB/s for {default, near} x {h2d, d2h} x {alone, concurrent}."""
import argparse
import ctypes as C
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "cuda-spmv-benchmark_b200", "python"))


def allowed(field):
    try:
        for ln in open("/proc/self/status"):
            if ln.startswith(field):
                return ln.split(":", 1)[1].strip()
    except Exception:
        pass
    return None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mb", type=int, default=0, help="MB per copy (default: this rank's share of a 3.2 GB vector)")
    ap.add_argument("--reps", type=int, default=4)
    a = ap.parse_args()
    import torch
    import spmv_b200 as B
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        saved = os.dup(1)
        os.dup2(2, 1)
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    L = B.load()
    nbytes = (a.mb << 20) if a.mb else (3_200_000_000 // world)
    nbytes -= nbytes % 8
    dev = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
    bufs = {"default": torch.empty(nbytes, dtype=torch.uint8).pin_memory()}
    p, node = C.c_void_p(), C.c_int(-2)
    if L.b200_host_alloc_near(local, nbytes, C.byref(p), C.byref(node)) == 0:
        arr = (C.c_ubyte * nbytes).from_address(p.value)
        bufs["near"] = torch.frombuffer(arr, dtype=torch.uint8)

    def sync_all():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
            torch.cuda.synchronize()

    def timed(fn):
        best = None
        for _ in range(a.reps):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1)
            best = ms if best is None else min(best, ms)
        return nbytes / (best * 1e-3) / 1e9

    out = {"rank": rank, "gpu": local, "gpu_numa_node": L.b200_host_node_of_device(local), "near_node": node.value,
           "cpus_allowed": allowed("Cpus_allowed_list"), "mems_allowed": allowed("Mems_allowed_list"),
           "bytes": nbytes, "gb_s": {}}
    for kind, h in bufs.items():
        for name, fn in (("h2d", lambda: dev.copy_(h, non_blocking=True)), ("d2h", lambda: h.copy_(dev, non_blocking=True))):
            # alone: ranks take turns
            for r in range(world):
                sync_all()
                if r == rank:
                    out["gb_s"]["%s_%s_alone" % (kind, name)] = round(timed(fn), 2)
            sync_all()
            out["gb_s"]["%s_%s_concurrent" % (kind, name)] = round(timed(fn), 2)
    sync_all()
    rows = [out]
    if dist is not None:
        rows = [None] * world
        dist.all_gather_object(rows, out)
        sys.stdout.flush()
        os.dup2(saved, 1)
    if rank == 0:
        print(json.dumps({"world": world, "ranks": rows}, indent=1))
    if "near" in bufs:
        del bufs["near"]
        L.b200_host_free(p)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
