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
    """(signal_sets, labels, defect_positions) of one file, json_dataset.py:38-158 statement by statement, including where
    its exceptions are caught: per scan (:108-126, the scan is skipped and its window dropped) and per file (:158, the
    sequences appended before the exception stay)."""
    sets, labels, defects = [], [], []
    try:
        with open(path, "r") as f:
            data = json.load(f)
        for beam_key in data.keys():
            beam = data[beam_key]
            keys = sorted(beam.keys(), key=lambda x: int(x.split('_')[0]))             # :48
            if len(keys) < seq_length:                                                  # :51-52
                continue
            scans, lab, dfx = [], [], []
            for k in keys:
                scans.append(beam[k])
                if k.split('_')[1] == "Health":                                         # :69-71
                    lab.append(0)
                    dfx.append([0.0, 0.0])
                else:
                    lab.append(1)
                    try:                                                                # :74-79
                        r = k.split('_')[2].split('-')
                        dfx.append([float(r[0]), float(r[1])])
                    except Exception:
                        dfx.append([0.0, 0.0])
            n = len(keys)
            num = math.ceil(n / seq_length)
            for i in range(num):
                start = i * seq_length if i < num - 1 else n - seq_length
                seq, l, d = [], [], []
                for j in range(start, start + seq_length):
                    try:                                                                # :108-126
                        scan = scans[j]
                        if isinstance(scan, dict) and 'signal' in scan:
                            scan = scan['signal']
                        seq.append(np.array(scan, dtype=np.float32))
                        l.append(lab[j])
                        d.append(dfx[j])
                    except Exception:
                        continue
                if len(seq) != seq_length or any(len(x) != len(seq[0]) for x in seq):   # :129-146
                    continue
                sets.append(np.array(seq, dtype=np.float32))
                labels.append(np.array(l, np.float32))
                defects.append(np.array(d, np.float32).reshape(-1, 2))
    except Exception:                                                                    # :158
        pass
    return sets, labels, defects
