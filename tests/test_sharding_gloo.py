"""N > 1 path on CPU: world_size-2 gloo run of the shard arithmetic and the detection all-gather."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from defectdetection_viaobjectdetection_b200 import sharding
from defectdetection_viaobjectdetection_b200._lib import DETECTION


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _fake_records(first, last, seed):
    """Deterministic records for sets [first, last): what a shard's predict_records would return."""
    rng = np.random.default_rng(seed)
    rows = []
    for b in range(first, last):
        for i in sorted(rng.choice(50, size=rng.integers(0, 6), replace=False)):
            rows.append((b - first, i, 1, int(i * 3), int(i * 5), 0.1 * i, 0.2 * i, 0.5, 0.0, 0.0, 0.5 + b))
    return np.array(rows, dtype=DETECTION) if rows else np.zeros(0, dtype=DETECTION)


def _worker(rank, world, port, n_sets, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = sharding.shard_range(n_sets, world, rank)
    rec = _fake_records(lo, hi, seed=lo)
    full = sharding.all_gather_records(rec, first_set=lo)
    # sharded evaluation: per-shard metrics add up (counts exactly, fp64 sums in rank order)
    mine = dict(tp=10 + rank, fp=3 * rank, fn=7, tn=100 * (rank + 1), sum_iou=0.1 * (rank + 1), sum_position_error=1e-3 + rank)
    tot = sharding.all_reduce_metrics(mine)
    assert (tot["tp"], tot["fp"], tot["fn"], tot["tn"]) == (sum(10 + r for r in range(world)), sum(3 * r for r in range(world)),
                                                            7 * world, sum(100 * (r + 1) for r in range(world)))
    want_iou = 0.0
    for r in range(world):
        want_iou += 0.1 * (r + 1)
    assert tot["sum_iou"] == want_iou and abs(tot["sum_position_error"] - sum(1e-3 + r for r in range(world))) < 1e-12
    if rank == 0:
        np.save(out, full)
    dist.barrier()
    dist.destroy_process_group()


def test_shard_range_alignment():
    """Aligned shards: boundaries on multiples of the layout period, still an exact partition."""
    assert sharding.bf16_shard_alignment(50) == 32 and sharding.bf16_shard_alignment(300) == 16
    for n in (0, 1, 31, 32, 33, 1000, 20000):
        for w in (1, 2, 4, 8):
            spans = [sharding.shard_range(n, w, r, align=32) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            for (a0, a1), (b0, b1) in zip(spans, spans[1:]):
                assert a1 == b0 and a0 <= a1
            assert all(lo % 32 == 0 for lo, _ in spans if lo < n)


def test_shard_range_weights_partition_in_proportion():
    w = [23.3] * 4 + [35.5] * 4
    for n in (0, 1, 7, 3334, 26667):
        spans = [sharding.shard_range(n, 8, r, weights=w) for r in range(8)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
    spans = [sharding.shard_range(26667, 8, r, align=32, weights=w) for r in range(8)]
    sizes = [hi - lo for lo, hi in spans]
    assert all(lo % 32 == 0 for lo, _ in spans)
    assert abs(sizes[4] / sizes[0] - 35.5 / 23.3) < 0.02
    with pytest.raises(ValueError):
        sharding.shard_range(10, 2, 0, weights=[1.0])


def test_shard_range_partitions_exactly():
    for n in (0, 1, 7, 8, 9, 3334, 160000):
        for w in (1, 2, 4, 8):
            spans = [sharding.shard_range(n, w, r) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1
    lo, hi, nw = sharding.msc_shard_range(1000, 300, 2, 1)
    assert (lo, hi, nw) == (2, 4, 4)


def test_all_gather_records_world2_gloo(tmp_path):
    n_sets, world = 37, 2
    out = str(tmp_path / "gathered.npy")
    mp.spawn(_worker, args=(world, _free_port(), n_sets, out), nprocs=world, join=True)
    got = np.load(out)
    parts = []
    for r in range(world):
        lo, hi = sharding.shard_range(n_sets, world, r)
        p = _fake_records(lo, hi, seed=lo)
        p["set_index"] += lo
        parts.append(p)
    ref = np.concatenate(parts)
    assert len(got) == len(ref) > 0
    for f in DETECTION.names:
        np.testing.assert_array_equal(got[f], ref[f], err_msg=f)
    # scan order: (set, position) non-decreasing
    key = got["set_index"].astype(np.int64) * 1000 + got["position"]
    assert np.all(np.diff(key) > 0)
