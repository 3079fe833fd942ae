"""CPU restatement of the reference's windowing rules (test infrastructure).

Two rules turn a run of ``n`` consecutive A-scans (one beam along the scan
axis, or one scan position across beams) into fixed-length sets:

* signals/ family -- signals/improved_multisignal/json_dataset.py:84-103:
  ceil(n/L) windows, window i < last = [i*L, i*L+L), the last one end-anchored
  [n-L, n) (overlapping its predecessor); runs shorter than L are skipped (:51-52).
* SignalSequenceDetection -- dataset_preparation.py:205,222-225,245-250,280-282:
  all-zero runs dropped; n < L zero-padded to L; n == L as is; n > L overlapping
  windows with total = ceil(n/(L/2)), step = max(1, floor((n-L)/(total-1))) from 0
  while start <= n-L, plus a tail window [n-L, n) when (n-L) % step != 0.
  (The reference additionally filters windows by training annotations; that is
  label handling, not part of the inference path.)

Each rule returns a list of (start, valid_len): the window covers rows
[start, start+valid_len) followed by L-valid_len zero rows.
"""
from __future__ import annotations

import math

import numpy as np


def msc_windows(n, seq_length=50):
    if n < seq_length:
        return []
    num = math.ceil(n / seq_length)
    out = []
    for i in range(num):
        start = i * seq_length if i < num - 1 else n - seq_length
        out.append((start, seq_length))
    return out


def ssd_windows(n, seq_length=50):
    if n < seq_length:
        return [(0, n)]
    if n == seq_length:
        return [(0, n)]
    total = max(1, int(np.ceil(n / (seq_length / 2))))
    step = max(1, int(np.floor((n - seq_length) / (total - 1)))) if total > 1 else seq_length
    out = [(s, seq_length) for s in range(0, n - seq_length + 1, step)]
    if (n - seq_length) % step != 0:
        out.append((n - seq_length, seq_length))
    return out


def gather_windows(volume, rule, seq_length=50, drop_all_zero=None):
    """volume [G, n, S] -> (sets [W, L, S] float32, table int32 [W, 3] = (group, start, valid_len)).
    drop_all_zero defaults to the SSD behaviour (dataset_preparation.py:205) for rule 'ssd'."""
    volume = np.asarray(volume)
    G, n, S = volume.shape
    windows = msc_windows(n, seq_length) if rule == "msc" else ssd_windows(n, seq_length)
    if drop_all_zero is None:
        drop_all_zero = rule == "ssd"
    sets, table = [], []
    for g in range(G):
        if drop_all_zero and np.all(volume[g] == 0):
            continue
        for start, valid in windows:
            w = np.zeros((seq_length, S), dtype=np.float32)
            w[:valid] = volume[g, start:start + valid].astype(np.float32)   # json_dataset.py:112-116 cast
            sets.append(w)
            table.append((g, start, valid))
    if not sets:
        return np.zeros((0, seq_length, S), np.float32), np.zeros((0, 3), np.int32)
    return np.stack(sets), np.asarray(table, dtype=np.int32)
