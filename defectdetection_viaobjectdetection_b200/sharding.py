"""Sharding of a scan volume across the GPUs of one box.

Sets (windows) are independent: attention, the set-axis convolution, the recurrences and the softmax over N
never cross a set, and BatchNorm uses running statistics, so a volume shards by contiguous blocks of sets
(= contiguous scan positions) with NO collective on the data path.  One process per GPU, one library
context per process.  The only exchange is optional: gathering the per-shard detection records, a few
bytes per kept A-scan, with ``torch.distributed`` (NCCL over NVLink on GPUs, gloo on CPU for tests).
"""
from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist

from ._lib import DETECTION


def shard_range(n_sets, world_size, rank, align=1, weights=None):
    """Contiguous, near-equal blocks in scan order: the first ranks get one more unit.

    ``weights`` (one positive number per rank): shard sizes proportional to them instead of equal -- for boxes whose
    GPUs do not share the host-to-device bandwidth equally (tools/h2d_ceiling.py: four of eight GPUs at 23 GB/s, four at
    35 GB/s), where equal shards leave the faster root complex idle in an end-to-end scan.

    ``align`` (sets): shard boundaries fall on multiples of it (the last shard takes the remainder).  The bf16
    convolution models pool per 128-row tile of their flat activation layout, whose period is 64 A-scans:
    ``align = bf16_shard_alignment(n_per_set)`` makes every shard reproduce the single-GPU result bit for bit."""
    if not (0 <= rank < world_size):
        raise ValueError("rank out of range")
    align = max(1, int(align))
    units = -(-int(n_sets) // align)                     # blocks of `align` sets (the last one may be short)
    if weights is not None:
        w = [float(x) for x in weights]
        if len(w) != int(world_size) or min(w) <= 0:
            raise ValueError("weights: one positive number per rank")
        cum = np.concatenate([[0.0], np.cumsum(w)]) / sum(w)
        edges = np.rint(cum * units).astype(np.int64)    # monotone, edges[0] = 0, edges[-1] = units
        return min(int(edges[rank]) * align, int(n_sets)), min(int(edges[rank + 1]) * align, int(n_sets))
    base, extra = divmod(units, int(world_size))
    start = rank * base + min(rank, extra)
    stop = start + base + (1 if rank < extra else 0)
    return min(start * align, int(n_sets)), min(stop * align, int(n_sets))


def bf16_shard_alignment(n_per_set):
    """Sets per period of the flat-row layout (64 A-scans): shard / chunk boundaries on its multiples keep the
    pooled sums of the bf16 convolution models in the same order as an unsharded run."""
    import math
    return 64 // math.gcd(64, int(n_per_set))


def msc_shard_range(n_scans, seq_length, world_size, rank):
    """Shard a run of ``n_scans`` A-scans that will be cut with the signals/ windowing rule
    (json_dataset.py:84-103): whole windows per rank, the end-anchored (overlapping) last window stays
    with the rank that owns the last scans.  Returns (first_window, last_window_exclusive, n_windows)."""
    n_windows = 0 if n_scans < seq_length else -(-n_scans // seq_length)
    lo, hi = shard_range(n_windows, world_size, rank)
    return lo, hi, n_windows


def all_gather_records(records, first_set=0, group=None, device=None):
    """Every rank contributes its structured detection records (dtype DETECTION, ``set_index`` local to the
    shard); returns the records of the whole volume in scan order with global ``set_index``.
    Variable-length: counts are exchanged first, payloads are padded to the longest shard."""
    world = dist.get_world_size(group)
    backend = dist.get_backend(group)
    dev = torch.device(device) if device is not None else (
        torch.device("cuda", torch.cuda.current_device()) if backend == "nccl" else torch.device("cpu"))
    rec = np.ascontiguousarray(records).copy()
    rec["set_index"] += int(first_set)
    n = torch.tensor([len(rec)], dtype=torch.int64, device=dev)
    counts = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(counts, n, group=group)
    counts = [int(c.item()) for c in counts]
    cap = max(max(counts), 1) * DETECTION.itemsize
    payload = torch.zeros(cap, dtype=torch.uint8, device=dev)
    if len(rec):
        raw = torch.from_numpy(rec.view(np.uint8).reshape(-1))
        payload[: raw.numel()] = raw.to(dev)
    gathered = [torch.empty_like(payload) for _ in range(world)]
    dist.all_gather(gathered, payload, group=group)
    parts = [g[: c * DETECTION.itemsize].cpu().numpy().view(DETECTION) for g, c in zip(gathered, counts)]
    return np.concatenate(parts) if parts else np.zeros(0, dtype=DETECTION)


def all_reduce_metrics(m, group=None, device=None):
    """Sum the per-shard detection metrics (the dict of ``metrics_match`` / ``metrics_confusion``) over the ranks:
    sets are independent, so the counts and the fp64 sums of a sharded evaluation add up to those of the whole
    volume (integer counts exactly; the fp64 sums in rank order).  One small all-reduce, off the data path."""
    backend = dist.get_backend(group)
    dev = torch.device(device) if device is not None else (
        torch.device("cuda", torch.cuda.current_device()) if backend == "nccl" else torch.device("cpu"))
    counts = torch.tensor([m["tp"], m["fp"], m["fn"], m["tn"]], dtype=torch.int64, device=dev)
    dist.all_reduce(counts, op=dist.ReduceOp.SUM, group=group)
    # gather (not reduce) the fp64 sums so that they are added in rank order on every rank: reproducible
    sums = torch.tensor([m["sum_iou"], m["sum_position_error"]], dtype=torch.float64, device=dev)
    parts = [torch.zeros_like(sums) for _ in range(dist.get_world_size(group))]
    dist.all_gather(parts, sums, group=group)
    total = torch.zeros(2, dtype=torch.float64)
    for p in parts:
        total += p.cpu()
    tp, fp, fn, tn = (int(v) for v in counts.cpu())
    return dict(tp=tp, fp=fp, fn=fn, tn=tn, sum_iou=float(total[0]), sum_position_error=float(total[1]))
