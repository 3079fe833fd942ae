"""CPU restatement of the parsing half of JsonSignalDataset._load_all_json_files (test infrastructure).

Follows signals/improved_multisignal/json_dataset.py:36-160 line by line for ONE file, with Python's own json
module as the tokenizer; pinned on the fixtures under tests/golden/json_volume (made by the reference class)."""
from __future__ import annotations

import json
import math

import numpy as np


def load_beams(path):
    """[(beam_key, signals list of float32 arrays, labels, defects)] with scans sorted as json_dataset.py:48."""
    with open(path, "r") as f:
        data = json.load(f)
    out = []
    for beam_key in data.keys():
        beam = data[beam_key]
        keys = sorted(beam.keys(), key=lambda x: int(x.split('_')[0]))                 # :48
        signals, labels, defects = [], [], []
        for k in keys:
            scan = beam[k]
            if isinstance(scan, dict) and 'signal' in scan:                            # :113-114
                scan = scan['signal']
            signals.append(np.array(scan, dtype=np.float32))                           # :111-116
            if k.split('_')[1] == "Health":                                            # :69-71
                labels.append(0)
                defects.append([0.0, 0.0])
            else:
                labels.append(1)
                try:                                                                   # :74-79
                    r = k.split('_')[2].split('-')
                    defects.append([float(r[0]), float(r[1])])
                except Exception:
                    defects.append([0.0, 0.0])
        out.append((beam_key, signals, np.array(labels, np.int32), np.array(defects, np.float32).reshape(-1, 2)))
    return out


def signal_sets(path, seq_length=50):
    """(signal_sets, labels, defect_positions) of one file, json_dataset.py:51-52,84-160."""
    sets, labels, defects = [], [], []
    for _, sig, lab, dfx in load_beams(path):
        n = len(sig)
        if n < seq_length:
            continue
        num = math.ceil(n / seq_length)
        for i in range(num):
            start = i * seq_length if i < num - 1 else n - seq_length
            win = sig[start:start + seq_length]
            if any(len(s) != len(win[0]) for s in win):
                continue
            sets.append(np.array(win, dtype=np.float32))
            labels.append(lab[start:start + seq_length].astype(np.float32))
            defects.append(dfx[start:start + seq_length])
    return sets, labels, defects
