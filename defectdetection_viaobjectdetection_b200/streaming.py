"""Whole-volume inference from HOST memory: the end-to-end call of this package.

The reference feeds the model one window at a time (batch 1, one H2D per window, one D2H sync per scalar:
model_tester.py:37-117, predict.py:169-197).  ``VolumeScanner`` streams a host-resident stack of sets through
the drop-in module in resident chunks on a few CUDA streams, so the host->device copy of chunk i+1 and the
device->host copy of the kept detections of chunk i-1 overlap the kernels of chunk i.  Each stream has its
own library context (one paut_ctx <-> one stream); results come back as one structured array in the
reference's (set, position) order, or as the reference's ``list[list[dict]]`` via ``to_predictions``.
"""
from __future__ import annotations

import time

import numpy as np
import torch

from .runtime import DETECTION


class _Lane:
    def __init__(self, device, shape, dtype, capacity):
        self.stream = torch.cuda.Stream(device=device)
        self.x = torch.empty(shape, dtype=dtype, device=device)
        self.host = torch.empty(capacity * DETECTION.itemsize, dtype=torch.uint8).pin_memory()
        self.count_host = torch.zeros(1, dtype=torch.int32).pin_memory()
        self.done = torch.cuda.Event()
        self.pending = None          # (first_set, n_sets, det, count) of the chunk in flight


class VolumeScanner:
    """scanner = VolumeScanner(model, chunk_sets=256); records = scanner.scan(x_host, threshold=0.5)

    ``x_host``: CPU tensor [sets, N, S] (``DefectDetectionModel``: [sets, S, N]), fp32 or bf16; pinned memory
    makes the copies asynchronous.  Returns a numpy structured array (dtype ``DETECTION``) whose
    ``set_index`` is the index into ``x_host``."""

    def __init__(self, model, chunk_sets=256, device=None, lanes=4, reuse_output=False):
        """reuse_output: keep the host record arena between scans (two arenas, used alternately) instead of
        allocating -- and page-faulting -- a fresh one per scan; the array returned by scan() is then only valid
        until the second next scan() of this scanner (copy it to keep it)."""
        self.model = model
        self.reuse_output = bool(reuse_output)
        self._arenas, self._arena_turn = [None, None], 0
        self.chunk_sets = int(chunk_sets)
        self.num_lanes = max(2, int(lanes))
        self.device = torch.device(device) if device is not None else next(iter(model.state_dict().values())).device
        if self.device.type != "cuda":
            raise RuntimeError("VolumeScanner needs the model on a CUDA device (no CPU fallback)")
        self._lanes = None
        self._key = None
        self.h2d_bytes = 0
        self.d2h_bytes = 0

    def close(self):
        """Release what the scanner's lanes own outside torch's allocator: the module's packed weights for the lane
        streams and the lanes' library contexts with their workspaces (they are cached by raw stream handle, which
        CUDA may hand out again once the lane's Stream object is gone).  The scanner can be used again afterwards."""
        from .runtime import release_context
        if self._lanes:
            index = self.device.index if self.device.index is not None else torch.cuda.current_device()
            for lane in self._lanes:
                lane.stream.synchronize()
                self.model.release(index, lane.stream.cuda_stream)
                release_context(index, lane.stream.cuda_stream)
        self._lanes, self._key = None, None

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()
        return False

    def __del__(self):
        import sys
        if sys.is_finalizing():                # interpreter shutdown: the driver reclaims everything, handles may be gone
            return
        try:
            self.close()
        except Exception:                      # noqa: BLE001 -- interpreter shutdown: the driver reclaims everything
            pass

    def _ensure_lanes(self, x_host):
        shape = (self.chunk_sets,) + tuple(x_host.shape[1:])
        key = (shape, x_host.dtype)
        if self._key != key:
            self.close()                       # lanes of the previous shape: their contexts go with them
            n_per = x_host.shape[2] if self.model._kind == "conv1d_msc" else x_host.shape[1]
            self._lanes = [_Lane(self.device, shape, x_host.dtype, self.chunk_sets * n_per)
                           for _ in range(self.num_lanes)]
            self._key = key

    def _harvest(self, lane, out):
        if lane.pending is None:
            return
        first, n_sets, det, count, copied = lane.pending
        t0 = time.perf_counter()
        lane.done.synchronize()
        t1 = time.perf_counter()
        self.stats["wait_s"] += t1 - t0
        n = int(lane.count_host[0])
        if n:
            nbytes = n * DETECTION.itemsize
            if nbytes > copied:                      # more records than the speculative copy covered
                with torch.cuda.stream(lane.stream):
                    lane.host[copied:nbytes].copy_(det[copied:nbytes], non_blocking=True)
                lane.stream.synchronize()
                self.d2h_bytes += nbytes - copied
            # raw byte memcpy out of the pinned buffer (a structured-dtype copy would go field by field), then the
            # chunk-local set indices -> volume indices while the later chunks are still in flight
            self._raw[self._off:self._off + nbytes] = lane.host[:nbytes].numpy()
            if first:
                self._raw[self._off:self._off + nbytes].view(DETECTION)["set_index"] += first
            self._off += nbytes
            self._spec_bytes = max(self._spec_bytes, int(nbytes * 1.25) // 48 * 48 + 48)
        self.d2h_bytes += 4 + copied
        lane.pending = None
        self.stats["unpack_s"] += time.perf_counter() - t1

    @torch.no_grad()
    def scan(self, x_host, threshold=0.5):
        if x_host.is_cuda:
            raise ValueError("scan() takes a host tensor; call the module directly for device-resident data")
        if not x_host.is_contiguous():
            raise RuntimeError("input must be contiguous")
        self._ensure_lanes(x_host)
        self.h2d_bytes = self.d2h_bytes = 0
        self.stats = {"wait_s": 0.0, "unpack_s": 0.0, "launch_s": 0.0}
        self._spec_bytes = getattr(self, "_spec_bytes", 1 << 20)
        out = None
        n_total = x_host.shape[0]
        n_per = x_host.shape[2] if self.model._kind == "conv1d_msc" else x_host.shape[1]
        need = n_total * n_per * DETECTION.itemsize                                  # worst case
        if self.reuse_output:
            self._arena_turn ^= 1
            if self._arenas[self._arena_turn] is None or self._arenas[self._arena_turn].size < need:
                self._arenas[self._arena_turn] = np.empty(need, dtype=np.uint8)
            self._raw = self._arenas[self._arena_turn]
        else:
            self._raw = np.empty(need, dtype=np.uint8)                               # touched lazily
        self._off = 0
        main = torch.cuda.current_stream(self.device)
        for lane in self._lanes:
            lane.stream.wait_stream(main)
        for i, first in enumerate(range(0, n_total, self.chunk_sets)):
            lane = self._lanes[i % self.num_lanes]
            self._harvest(lane, out)
            n_sets = min(self.chunk_sets, n_total - first)
            t_launch = time.perf_counter()
            with torch.cuda.stream(lane.stream):
                xd = lane.x[:n_sets]
                xd.copy_(x_host[first:first + n_sets], non_blocking=True)
                native, (outs, struct, (B, N, S)) = self.model._run(xd)
                det, count = native.postprocess(struct, B, N, S, threshold, self.device)
                lane.count_host.copy_(count, non_blocking=True)
                # speculative copy of the record buffer in the same stream (no second round trip): as many
                # bytes as the previous chunks needed, plus head-room; the rare overflow is fetched at harvest
                copied = min(det.numel(), self._spec_bytes)
                lane.host[:copied].copy_(det[:copied], non_blocking=True)
                lane.done.record(lane.stream)
            lane.pending = (first, n_sets, det, count, copied)
            self.stats["launch_s"] += time.perf_counter() - t_launch
            self.h2d_bytes += xd.numel() * xd.element_size()
        n_chunks = (n_total + self.chunk_sets - 1) // self.chunk_sets
        for j in range(self.num_lanes):                        # flush the chunks in flight, oldest first
            self._harvest(self._lanes[(n_chunks + j) % self.num_lanes], out)
        for lane in self._lanes:
            main.wait_stream(lane.stream)
        rec = self._raw[:self._off].view(DETECTION)
        self._raw = None
        return rec


def to_predictions(records, n_sets):
    """Structured records -> the reference's ``list[list[dict]]`` (two_stage_model.py:490-497 field names
    where they exist; every record also carries the integer sample indices)."""
    out = [[] for _ in range(n_sets)]
    for r in records:
        out[int(r["set_index"])].append({
            "position": int(r["position"]), "class": int(r["cls"]), "score": float(r["score"]),
            "defect_position": np.array([r["start"], r["end"]], dtype=np.float32),
            "start_index": int(r["start_index"]), "end_index": int(r["end_index"]),
            "adjusted_confidence": float(r["confidence"]),
        })
    return out
