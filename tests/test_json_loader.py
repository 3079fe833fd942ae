"""the native host loader of the reference's JSON volume format against the reference's
own JsonSignalDataset (fixtures and expected arrays in tests/golden/json_volume, made by running
signals/improved_multisignal/json_dataset.py in the build container)."""
import json
import os

import numpy as np
import pytest
import torch

from defectdetection_viaobjectdetection_b200 import dataio
from tests._golden import GOLDEN_DIR

JDIR = os.path.join(GOLDEN_DIR, "json_volume")


@pytest.fixture(scope="module")
def expected():
    return np.load(os.path.join(JDIR, "expected.npz"))


@pytest.mark.parametrize("name", ["a", "b"])
def test_host_loader_matches_reference_dataset(name, expected):
    sets, labels, defects = dataio.json_signal_sets([os.path.join(JDIR, name + ".json")], seq_length=5)
    np.testing.assert_array_equal(sets, expected[name + "_sets"])            # float32(float64(text)): bit-exact
    np.testing.assert_array_equal(labels, expected[name + "_labels"])
    np.testing.assert_array_equal(defects, expected[name + "_defects"])      # nan == nan positionally


def test_parser_details():
    beams = dataio.load_json_volume(os.path.join(JDIR, "a.json"))
    assert [b["key"] for b in beams] == ["beam_0", "beam_1", "beam_2"]
    assert beams[0]["scan_order"].tolist() == list(range(13))
    # stable sort: "3_Defect_0.125-0.875" was written before "3_Defect_0.5-0.75"
    order = beams[2]["scan_order"].tolist()
    assert order == sorted(order) and order.count(3) == 2
    i = order.index(3)
    assert beams[2]["defects"][i].tolist() == [0.125, 0.875] and beams[2]["defects"][i + 1].tolist() == [0.5, 0.75]
    with open(os.path.join(JDIR, "a.json")) as f:
        raw = json.load(f)
    first_key = sorted(raw["beam_0"], key=lambda k: int(k.split("_")[0]))[0]
    np.testing.assert_array_equal(beams[0]["signals"][0], np.array(raw["beam_0"][first_key], dtype=np.float32))


def test_parallel_parse_equals_sequential(tmp_path, monkeypatch):
    """Files above 1 MB are parsed beam-parallel from a mapped file; the result (and the error a sequential parse
    would report first) does not depend on the thread count or on the I/O mode."""
    rng = np.random.default_rng(4)
    data = {f"beam_{b}": {f"{i}_{'Health' if i % 3 else 'Defect_0.1-0.9'}": [float(x) for x in rng.random(64)]
                          for i in rng.permutation(120)} for b in range(12)}
    p = tmp_path / "big.json"
    p.write_text(json.dumps(data))
    assert p.stat().st_size > (1 << 20)
    results = []
    for threads, io in (("1", "read"), ("3", "mmap"), ("8", "populate")):
        monkeypatch.setenv("PAUT_JSON_THREADS", threads)
        monkeypatch.setenv("PAUT_JSON_IO", io)
        results.append(dataio.load_json_volume(str(p), with_keys=True))
    for other in results[1:]:
        for a, b in zip(results[0], other):
            assert a["key"] == b["key"] and a["scan_keys"] == b["scan_keys"]
            for f in ("signals", "labels", "defects", "scan_order"):
                np.testing.assert_array_equal(a[f], b[f])
    want = np.array([data["beam_5"][k] for k in sorted(data["beam_5"], key=lambda k: int(k.split("_")[0]))], np.float32)
    np.testing.assert_array_equal(results[1][5]["signals"], want)
    bad = dict(data)
    bad["beam_4"] = {"x_Health": [1.0]}
    bad["beam_9"] = {"y_Health": [1.0]}
    p.write_text(json.dumps(bad))
    with pytest.raises(ValueError, match="x_Health"):
        dataio.load_json_volume(str(p))


def test_parser_errors(tmp_path):
    for text in ('{"b": {"x_Health": [1, 2]}}',          # int('x') fails -> the reference aborts the file
                 '{"b": {"3": [1, 2]}}',                 # no label field
                 '{"b": {"3_Health": [1, 2,]}}',         # invalid JSON
                 '{"b": {"3_Health": [0x10]}}'):
        p = tmp_path / "bad.json"
        p.write_text(text)
        with pytest.raises(ValueError):
            dataio.load_json_volume(str(p))
    with pytest.raises(ValueError):
        dataio.load_json_volume(str(tmp_path / "missing.json"))


@pytest.mark.gpu
def test_device_windowing_matches_reference_dataset(expected):
    for name in ("a", "b"):
        sets, labels, defects = dataio.json_signal_sets([os.path.join(JDIR, name + ".json")], seq_length=5, device="cuda")
        assert sets.is_cuda
        np.testing.assert_array_equal(sets.cpu().numpy(), expected[name + "_sets"])
        np.testing.assert_array_equal(labels, expected[name + "_labels"])
    sets16, _, _ = dataio.json_signal_sets([os.path.join(JDIR, "b.json")], seq_length=5, device="cuda", dtype=torch.bfloat16)
    np.testing.assert_array_equal(sets16.float().cpu().numpy(),
                                  torch.from_numpy(expected["b_sets"]).to(torch.bfloat16).float().numpy())


def test_scan_sequences_match_reference_grouping():
    """SSD-family layout: SignalSequencePreparation.get_datafile_sequences (dataset_preparation.py:36-116) run in the
    build container on c_ssd.json -> expected_ssd.npz; beams out of order (one float index), one beam lacking scans."""
    z = np.load(os.path.join(JDIR, "expected_ssd.npz"))
    got = dataio.json_scan_sequences(os.path.join(JDIR, "c_ssd.json"))
    assert list(got.keys()) == [str(k) for k in z["keys"]]
    for k, v in got.items():
        np.testing.assert_array_equal(v, z["seq_" + k])
    assert [v.shape[0] for v in got.values()] == [3, 4, 3, 4, 3, 4]


def test_ragged_beam_is_filtered_window_by_window(tmp_path):
    """A beam whose scans differ in length: the reference drops only the windows that contain the odd scan
    (json_dataset.py:136-146); expected arrays from JsonSignalDataset on each beam of d_ragged.json."""
    z = np.load(os.path.join(JDIR, "expected_ragged.npz"))
    with open(os.path.join(JDIR, "d_ragged.json")) as f:
        data = json.load(f)
    for beam in ("r0", "r1"):
        p = tmp_path / f"{beam}.json"
        p.write_text(json.dumps({beam: data[beam]}))
        sets, labels, defects = dataio.json_signal_sets([str(p)], seq_length=5)
        np.testing.assert_array_equal(sets, z[beam + "_sets"])
        np.testing.assert_array_equal(labels, z[beam + "_labels"])
        np.testing.assert_array_equal(defects, z[beam + "_defects"])
    assert z["r0_sets"].shape[0] == 2                       # windows [0,5) and [8,13); [5,10) holds the short scan
    with pytest.raises(ValueError):                          # both beams at once: S = 8 and S = 6 cannot be one batch
        dataio.json_signal_sets([os.path.join(JDIR, "d_ragged.json")], seq_length=5)


def test_sequence_targets_match_reference_dataset():
    """Prepared-sequence pickle -> dense targets: SignalSequenceDataset.__getitem__ (dataset_preparation.py:429-476)
    over sequences.pkl, run in the build container -> expected_sequences.npz (incl. a padded sequence without
    start_idx and the reference's 'Health is the last class id' mapping)."""
    z = np.load(os.path.join(JDIR, "expected_sequences.npz"))
    signals, label, pos, label_map, files, keys = dataio.sequence_targets(
        dataio.load_sequences_pickle(os.path.join(JDIR, "sequences.pkl")))
    np.testing.assert_array_equal(signals, z["signals"])
    np.testing.assert_array_equal(label, z["label"])
    np.testing.assert_array_equal(pos, z["pos"])
    assert [f"{k}={v}" for k, v in label_map.items()] == [str(s) for s in z["label_map"]]
    assert label_map["Health"] == len(label_map) - 1 and (label[3] == label_map["Health"]).all()
    assert keys == [str(i) for i in range(7)] and files[:2] == ["f0", "f1"]


def test_annotation_merging_and_defect_only_windows_match_reference():
    """f1 remainder: json_scan_annotations / normalize_annotations / prepare_beam_sequences against the reference's
    SignalSequencePreparation run on the same file (tests/golden/make_golden.py --prep; dataset_preparation.py:70-100,
    118-152, 188-313)."""
    import hashlib
    pdir = os.path.join(GOLDEN_DIR, "prep_volume")
    with open(os.path.join(pdir, "expected.json")) as f:
        exp = json.load(f)
    path = os.path.join(pdir, "weld.json")
    ann, lims = dataio.json_scan_annotations(path)
    assert list(ann.keys()) == sorted(exp["annotations"], key=int)
    assert {k: v for k, v in ann.items()} == exp["annotations"]              # merged ranges, labels, order
    assert list(lims) == exp["beam_lims"]
    norm = dataio.normalize_annotations(ann, lims)
    assert norm == exp["normalized"]                                          # same fp64 arithmetic: exact
    seqs = dataio.json_scan_sequences(path)
    got = dataio.prepare_beam_sequences({"weld": seqs}, {"weld": norm}, seq_length=50)
    assert len(got) == len(exp["sequences"])
    for g, e in zip(got, exp["sequences"]):
        assert g["scan_key"] == e["scan_key"]
        assert g.get("start_idx") == e["start_idx"] and g.get("end_idx") == e["end_idx"]
        assert g.get("original_length") == e["original_length"]
        # the reference holds float64 rows (json floats); ours are float32(float64(text)) -- the fixture's samples are
        # float32-representable, so the float64 views are identical
        assert hashlib.sha256(np.ascontiguousarray(g["signals"], dtype=np.float64).tobytes()).hexdigest() == e["sha256"]
    # the filter really filters: the long runs have more windows than were kept
    from defectdetection_viaobjectdetection_b200.runtime import window_table
    assert sum(1 for s in got if s["scan_key"] == "1") < len(window_table("ssd", 140, 50))
