"""On-disk volume format -> resident sets.

``load_json_volume`` parses one JSON file of the reference's format with the native host loader
(csrc/host_json.cu); ``json_signal_sets`` restates ``JsonSignalDataset`` (signals/improved_multisignal/
json_dataset.py:9-160): the same sequences, labels and defect positions in the same order, with the windows
gathered on the device by ``paut_window_gather`` when a CUDA device is given.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import _lib


BEAM_BAD_ORDER_KEY, BEAM_NO_LABEL, SCAN_SKIPPED = 1, 2, -100          # include/paut.h: PAUT_JSON_BEAM_*, PAUT_JSON_SCAN_SKIPPED


def load_json_volume(path, with_keys=False, strict=True):
    """One file -> list of beams in file order: dict(key, signals float32 [n,S] (None if the scans differ in length or
    one of them is skipped: then ``ragged`` is the list of per-scan arrays, None for a skipped scan), labels int32 [n],
    defects float32 [n,2], scan_order int64 [n], scan_keys (the full key strings, only with ``with_keys=True``),
    status / status_msg) with the scans in the reference's sorted order.  ``status`` holds the BEAM_* bits of a beam on
    which JsonSignalDataset would raise out of its per-file try block (a key without an integer prefix: the sort of
    every beam; a key without a label field: beams with at least seq_length scans); ``strict=True`` raises ValueError
    for such a file like round 1 did, ``strict=False`` leaves the decision to the caller (json_signal_sets keeps the
    earlier beams, as the reference does).  Repeated scan keys keep the last value, like Python's json.
    Beams are parsed on several threads (PAUT_JSON_THREADS, default: all cores up to 16) from a memory-mapped file."""
    lib = _lib.load()
    h = C.c_void_p()
    if lib.paut_json_load_host(os.fsencode(path), C.byref(h)) != 0:
        raise ValueError(lib.paut_json_last_error().decode())
    try:
        beams = []
        key, n, S, msg = C.c_char_p(), C.c_int64(), C.c_int64(), C.c_char_p()
        for b in range(lib.paut_json_num_beams(h)):
            lib.paut_json_beam_info(h, b, C.byref(key), C.byref(n), C.byref(S))
            status = lib.paut_json_beam_status(h, b, C.byref(msg))
            status_msg = msg.value.decode() if status else ""
            if strict and status:
                raise ValueError(f"{path}: {status_msg}")
            labels = np.empty(n.value, np.int32)
            defects = np.empty((n.value, 2), np.float32)
            order = np.empty(n.value, np.int64)
            signals = np.empty((n.value, S.value), np.float32) if S.value >= 0 else None
            rc = lib.paut_json_beam_copy_host(h, b, signals.ctypes.data if signals is not None else None,
                                              labels.ctypes.data, defects.ctypes.data, order.ctypes.data)
            if rc != 0:
                raise ValueError(lib.paut_json_last_error().decode())
            ragged = None
            if signals is None:                                  # scans of different lengths: keep them one by one
                ragged = []
                for i in range(n.value):
                    ln = lib.paut_json_scan_copy_host(h, b, i, None, 0)
                    if ln == SCAN_SKIPPED:                       # an object without "signal": the reference skips the scan
                        ragged.append(None)
                        continue
                    row = np.empty(ln, np.float32)
                    lib.paut_json_scan_copy_host(h, b, i, row.ctypes.data, ln)
                    ragged.append(row)
            keys = [lib.paut_json_scan_key(h, b, i).decode() for i in range(n.value)] if with_keys else None
            beams.append(dict(key=key.value.decode(), signals=signals, labels=labels, defects=defects, scan_order=order,
                              scan_keys=keys, ragged=ragged, status=status, status_msg=status_msg))
        return beams
    finally:
        lib.paut_json_free(h)


def json_signal_sets(json_dir_or_files, seq_length=50, device=None, dtype=None):
    """JsonSignalDataset(json_dir, seq_length) without the Python loops: returns (signal_sets [W, L, S],
    labels float32 [W, L], defect_positions float32 [W, L, 2]) in the dataset's order.  ``signal_sets`` is a CUDA
    tensor gathered on the device when ``device`` is a CUDA device (dtype fp32, or bf16 on request), else numpy.
    Error granularity as in the reference (json_dataset.py:38-158): a file that is not valid JSON contributes nothing; a
    key that breaks the sort of a beam (:48) or -- in a beam with at least seq_length scans -- the label lookup (:69) ends
    the file there, and the sequences of the earlier beams stay; a scan that cannot be converted (:108-126) only drops the
    windows that contain it."""
    import torch
    from .runtime import gather_windows, window_table
    if isinstance(json_dir_or_files, (str, os.PathLike)) and os.path.isdir(json_dir_or_files):
        files = [os.path.join(json_dir_or_files, f) for f in os.listdir(json_dir_or_files) if f.endswith(".json")]
    else:
        files = list(json_dir_or_files)
    on_gpu = device is not None and torch.device(device).type == "cuda"
    sets, labels, defects = [], [], []
    for path in files:
        try:
            beams = load_json_volume(path, strict=False)
        except ValueError as e:
            print(f"Error loading {os.path.basename(path)}: {e}")
            continue
        for beam in beams:
            n = len(beam["labels"])
            if beam["status"] & BEAM_BAD_ORDER_KEY or (beam["status"] & BEAM_NO_LABEL and n >= seq_length):
                print(f"Error loading {os.path.basename(path)}: {beam['status_msg']}")
                break                                               # the reference leaves its per-file try block here
            wins = window_table("msc", n, seq_length)               # json_dataset.py:51-52, 84-103
            if not wins:
                continue                                            # fewer scans than seq_length
            if beam["signals"] is None:
                # scans of different lengths: a window is kept iff all its scans are as long as its first one
                # (json_dataset.py:136-146); such beams are rare, so they take the host path
                for s0, _ in wins:
                    rows = beam["ragged"][s0:s0 + seq_length]
                    if any(r is None for r in rows) or any(len(r) != len(rows[0]) for r in rows):
                        continue
                    w = np.stack(rows)
                    sets.append(torch.from_numpy(w).to(device=device, dtype=dtype or torch.float32)[None] if on_gpu else w[None])
                    labels.append(beam["labels"][s0:s0 + seq_length].astype(np.float32)[None])
                    defects.append(beam["defects"][s0:s0 + seq_length][None])
                continue
            idx = np.array([np.arange(s, s + seq_length) for s, _ in wins])
            labels.append(beam["labels"][idx].astype(np.float32))
            defects.append(beam["defects"][idx])
            if on_gpu:
                vol = torch.from_numpy(beam["signals"]).to(device)[None]
                out, _ = gather_windows(vol, "msc", seq_length, out_dtype=dtype or torch.float32)
                sets.append(out)
            else:
                sets.append(beam["signals"][idx])
    if not sets:
        return (torch.empty(0) if on_gpu else np.zeros((0, seq_length, 0), np.float32),
                np.zeros((0, seq_length), np.float32), np.zeros((0, seq_length, 2), np.float32))
    lengths = {s.shape[-1] for s in sets}
    if len(lengths) != 1:
        raise ValueError(f"beams with different signal lengths {sorted(lengths)} cannot form one batch "
                         "(the reference's DataLoader fails at collation)")
    cat = torch.cat(sets, 0) if on_gpu else np.concatenate(sets, 0)
    return cat, np.concatenate(labels, 0), np.concatenate(defects, 0)


def json_scan_sequences(path):
    """The signal half of SignalSequencePreparation.get_datafile_sequences (SignalSequenceDetection/
    dataset_preparation.py:36-116) -- the SSD-family layout, where one sequence runs ACROSS the beams at a fixed
    scan position: beams sorted by float(beam_key.split('_')[1]), signals grouped by the scan key's first field (the
    string: '07' and '7' are different keys), groups sorted by int(key).  Returns an ordered dict
    {scan_key: float32 [n_beams_with_that_scan, S]} (the reference holds float64 there and casts to float32 in its
    dataset class).  Windows: ``gather_windows(torch.from_numpy(seq)[None].cuda(), "ssd", 50, drop_all_zero=True)``."""
    from collections import OrderedDict
    beams = load_json_volume(path, with_keys=True)
    order = sorted(range(len(beams)), key=lambda i: float(beams[i]["key"].split("_")[1]))      # stable, like sorted()
    groups = OrderedDict()
    for b in order:
        beam = beams[b]
        if beam["signals"] is None:
            raise ValueError(f"beam {beam['key']!r}: scans of different lengths")
        for i, full_key in enumerate(beam["scan_keys"]):
            groups.setdefault(full_key.split("_")[0], []).append(beam["signals"][i])
    out = OrderedDict()
    for k in sorted(groups, key=lambda k: int(k)):                                               # stable
        out[k] = np.stack(groups[k]).astype(np.float32)
    return out


def load_sequences_pickle(path):
    """The prepared-sequence file of the SSD family (SignalSequenceDataset.__init__, dataset_preparation.py:417-427):
    a pickled list of dicts {signals [L,S], annotations [{bbox, label}], file_name, scan_key[, start_idx, end_idx,
    original_length]}.  Plain pickle.load -- load only files you trust, as with the reference."""
    import pickle
    with open(path, "rb") as f:
        return pickle.load(f)


def sequence_targets(sequences):
    """SignalSequenceDataset.__getitem__ for every sequence at once (dataset_preparation.py:429-476): returns
    (signals float32 [W,L,S], label int32 [W,L], defect_position float32 [W,L,2], label_map, file_names, scan_keys).
    label / defect_position are the dense targets paut_metrics_match takes.  Kept quirks of the reference: 'Health'
    is the LAST class id (so a defect class can have id 0), a signal takes the first annotation whose beam range
    contains (start_idx + i) / L, and sequences without 'start_idx' (zero-padded ones) get no defect targets."""
    labels = sorted({a["label"] for s in sequences for a in s["annotations"]})
    label_map = {lab: i for i, lab in enumerate(labels)}
    label_map["Health"] = len(label_map)
    W = len(sequences)
    if W == 0:
        return (np.zeros((0, 0, 0), np.float32), np.zeros((0, 0), np.int32), np.zeros((0, 0, 2), np.float32), label_map, [], [])
    L = len(sequences[0]["signals"])
    signals = np.stack([np.asarray(s["signals"], dtype=np.float32) for s in sequences])
    label = np.full((W, L), label_map["Health"], np.int32)
    pos = np.zeros((W, L, 2), np.float32)
    for w, s in enumerate(sequences):
        if "start_idx" not in s:
            continue
        n = len(s["signals"])
        for i in range(n):
            beam_position = (s["start_idx"] + i) / n
            for a in s["annotations"]:
                if a["bbox"][0] <= beam_position <= a["bbox"][1]:
                    label[w, i] = label_map[a["label"]]
                    pos[w, i] = (a["bbox"][2], a["bbox"][3])
                    break
    return signals, label, pos, label_map, [s["file_name"] for s in sequences], [s["scan_key"] for s in sequences]


# ------------------------------------------------------------------------------------------------ f1: training-side preparation
def json_scan_annotations(path):
    """The annotation half of SignalSequencePreparation.get_datafile_sequences (SignalSequenceDetection/
    dataset_preparation.py:53-116): per scan key the list of defects {"bbox": [beam_first, beam_last, start, end],
    "label"}, where a defect seen in consecutive beams with the same (start, end) extends the last entry instead of
    opening a new one (:85-91), 'Health' scans only create an empty list (:71-73) and a key whose range does not
    parse is reported and skipped (:99-101).  Returns (OrderedDict sorted by int(scan_key), (beam_start, beam_end))."""
    from collections import OrderedDict
    beams = load_json_volume(path, with_keys=True)
    order = sorted(range(len(beams)), key=lambda i: float(beams[i]["key"].split("_")[1]))
    beam_start = float(beams[order[0]]["key"].split("_")[1])
    beam_end = float(beams[order[-1]]["key"].split("_")[1])
    ann = {}
    for b in order:
        beam_idx = float(beams[b]["key"].split("_")[1])
        for scan_file in beams[b]["scan_keys"]:
            parts = scan_file.split("_")
            scan_key = parts[0]
            if parts[1] == "Health":
                ann.setdefault(scan_key, [])
                continue
            try:
                rng = parts[-1].split("-")
                d0, d1 = float(rng[0]), float(rng[1])
            except Exception as ex:                                  # noqa: BLE001 -- the reference catches everything here
                print(f"Error: {ex} in {os.path.splitext(os.path.basename(path))[0]}, {beams[b]['key']}, {scan_file}")
                continue
            cur = ann.get(scan_key)
            if not cur:
                ann[scan_key] = [{"bbox": [beam_idx, beam_idx, d0, d1], "label": parts[1]}]
            else:
                last = cur[-1]["bbox"]
                if last[2] == d0 and last[3] == d1 and last[1] == beam_idx - 1:
                    last[1] += 1
                else:
                    cur.append({"bbox": [beam_idx, beam_idx, d0, d1], "label": parts[1]})
    return OrderedDict(sorted(ann.items(), key=lambda kv: int(kv[0]))), (beam_start, beam_end)


def normalize_annotations(annotations, beam_lims):
    """dataset_preparation.py:118-152: beam positions -> [0, 1] over the file's beam range, defect range unchanged."""
    b0, b1 = beam_lims
    span = b1 - b0
    return {k: [{"bbox": [(d["bbox"][0] - b0) / span, (d["bbox"][1] - b0) / span, d["bbox"][2], d["bbox"][3]],
                 "label": d["label"]} for d in v] for k, v in annotations.items()}


def window_has_objects(annotations, start_idx, end_idx, num_signals):
    """The defect-only window filter of create_beam_sequences (dataset_preparation.py:255-268, :283-296): a window is
    kept iff for some defect some signal position i / num_signals of the window lies in its beam range.  Closed
    form of the reference's double loop (same float arithmetic: i / num_signals in fp64, inclusive bounds)."""
    for d in annotations:
        lo, hi = d["bbox"][0], d["bbox"][1]
        for i in range(start_idx, end_idx):
            if lo <= i / num_signals <= hi:
                return True
    return False


def prepare_beam_sequences(all_sequences, all_annotations, seq_length=50):
    """SignalSequencePreparation.create_beam_sequences (dataset_preparation.py:188-313) on host index arithmetic:
    all_sequences {file: {scan_key: float [n, S]}}, all_annotations {file: {scan_key: [defect]}} (normalised) ->
    the reference's list of sequence dicts (same keys, same order).  All-zero runs and runs without defects are
    skipped, short runs are zero-padded (``original_length``), long runs are cut with the overlapping-window rule
    (``runtime.window_table('ssd', n, seq_length)``) and only windows that contain a defect are kept."""
    from .runtime import window_table
    out = []
    for file_name, sequences in all_sequences.items():
        for scan_key, signals in sequences.items():
            signals = np.asarray(signals)
            if np.all(signals == 0):
                continue
            annotations = all_annotations[file_name].get(scan_key, [])
            if len(annotations) == 0:
                continue
            n = len(signals)
            if n < seq_length:
                padded = np.zeros((seq_length, signals.shape[1]), dtype=signals.dtype)
                padded[:n] = signals
                out.append({"file_name": file_name, "scan_key": scan_key, "signals": padded, "annotations": annotations,
                            "original_length": n, "has_objects": True})
            elif n == seq_length:
                out.append({"file_name": file_name, "scan_key": scan_key, "signals": signals, "annotations": annotations,
                            "has_objects": True})
            else:
                for start, _ in window_table("ssd", n, seq_length):
                    if window_has_objects(annotations, start, start + seq_length, n):
                        out.append({"file_name": file_name, "scan_key": scan_key, "signals": signals[start:start + seq_length],
                                    "annotations": annotations, "start_idx": start, "end_idx": start + seq_length,
                                    "has_objects": True})
    return out
