"""Pinned host -> device copy ceiling of one box at N ranks (one process per GPU, all copying at the same time).

    python tools/h2d_ceiling.py                                                            # 1 GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/h2d_ceiling.py

Each rank copies a pinned 640 MB buffer (the bf16 volume of 1 M A-scans) to its GPU `--reps` times, ONE
cudaMemcpyAsync per repetition on a side stream; rank 0 prints the per-rank and the aggregate GB/s, with and without
the ranks pinned to disjoint host-core sets.  This is the ceiling bench.py's end-to-end figure is held against
(e2e.h2d_ceiling_gbs): the MSC path moves 640 B per A-scan over PCIe, so e2e A-scans/s <= ceiling / 640 B."""
import argparse
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def measure(dev, nbytes, reps, barrier):
    host = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
    host.zero_()
    dst = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    s = torch.cuda.Stream(device=dev)
    with torch.cuda.stream(s):
        dst.copy_(host, non_blocking=True)
    s.synchronize()
    barrier()
    t0 = time.perf_counter()
    with torch.cuda.stream(s):
        for _ in range(reps):
            dst.copy_(host, non_blocking=True)
    s.synchronize()
    return nbytes * reps / (time.perf_counter() - t0) / 1e9


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mb", type=int, default=640)
    ap.add_argument("--reps", type=int, default=10)
    a = ap.parse_args()
    rank, local_rank = int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("NCCL_DEBUG", "WARN")
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    out = {"n_gpus": world, "mb": a.mb, "reps": a.reps, "cpus": len(os.sched_getaffinity(0))}
    for mode in ("unpinned", "pinned_cores"):
        if mode == "pinned_cores":
            cores = sorted(os.sched_getaffinity(0))
            per = len(cores) // world
            if world > 1 and per >= 2:
                os.sched_setaffinity(0, cores[local_rank * per:(local_rank + 1) * per])
        gbs = measure(dev, a.mb << 20, a.reps, barrier)
        t = torch.tensor([gbs], device=dev, dtype=torch.float64)
        if world > 1:
            allv = [torch.zeros_like(t) for _ in range(world)]
            dist.all_gather(allv, t)
            vals = [float(v[0]) for v in allv]
        else:
            vals = [gbs]
        out[mode] = {"per_rank_gbs": [round(v, 2) for v in vals], "aggregate_gbs": round(sum(vals), 2)}
    if rank == 0:
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
