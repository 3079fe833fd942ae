"""Property test of the native JSON volume loader against the oracle restatement (which is itself pinned on the
reference class through tests/golden/json_volume): random beams, keys, labels, ranges and number spellings."""
import json
import os

import numpy as np
import pytest
from hypothesis import HealthCheck, given, settings
from hypothesis import strategies as st

from defectdetection_viaobjectdetection_b200 import dataio
from oracle import jsonload
from tests._golden import GOLDEN_DIR

JDIR = os.path.join(GOLDEN_DIR, "json_volume")


@pytest.mark.parametrize("name,L", [("a", 5), ("b", 5)])
def test_oracle_loader_pinned_on_reference_fixtures(name, L):
    z = np.load(os.path.join(JDIR, "expected.npz"))
    sets, labels, defects = jsonload.signal_sets(os.path.join(JDIR, name + ".json"), L)
    np.testing.assert_array_equal(np.array(sets, np.float32), z[name + "_sets"])
    np.testing.assert_array_equal(np.array(labels, np.float32), z[name + "_labels"])
    np.testing.assert_array_equal(np.array(defects, np.float32), z[name + "_defects"])


@pytest.mark.parametrize("name", ["e_partial1", "e_partial2", "e_partial3"])
def test_partly_usable_files_match_reference_dataset(name):
    """Error granularity of JsonSignalDataset (fixtures made by the reference class, make_golden.py --json-partial): a scan
    that cannot be converted drops its window; a key that breaks the sort or the label lookup ends the file and keeps the
    earlier beams; a short beam is skipped before the label lookup; a repeated key keeps its last value.  Oracle and
    native loader against the reference's arrays."""
    z = np.load(os.path.join(JDIR, "expected_partial.npz"))
    path = os.path.join(JDIR, name + ".json")
    for sets, labels, defects in (jsonload.signal_sets(path, 5), dataio.json_signal_sets([path], seq_length=5)):
        assert len(sets) == len(z[name + "_sets"]) > 0
        np.testing.assert_array_equal(np.array(sets, np.float32), z[name + "_sets"])
        np.testing.assert_array_equal(np.array(labels, np.float32), z[name + "_labels"])
        np.testing.assert_array_equal(np.array(defects, np.float32), z[name + "_defects"])


number = st.one_of(
    st.floats(allow_nan=False, allow_infinity=False, width=64),
    st.floats(min_value=0.0, max_value=1.0),
    st.integers(min_value=-10**6, max_value=10**22),
    st.sampled_from([0.0, -0.0, 1e-320, 5e-324, 1e308, 1.7976931348623157e308, 3.4028235e38, 3.4028236e38, 1e-46,
                     0.1, 0.30000000000000004, 16777217.0]))
label = st.sampled_from(["Health", "Defect", "Defect_0.25-0.5", "Crack_0.1-0.30000001", "Pore_-0.1-0.5", "Defect_abc",
                         "Defect_1e-1-5E-1", "Defect_.5-1.", "Defect_inf-nan", "Defect_ 0.2 -0.7", "Health_0.3-0.4",
                         "Defect_0.1-0.2-0.3", "D_1e400-2", "D_-", "D_+1-+2", "D_0x1-2", "D_1_000-2"])


@st.composite
def volumes(draw):
    S = draw(st.integers(min_value=1, max_value=6))
    beams = {}
    for b in range(draw(st.integers(min_value=1, max_value=4))):
        n = draw(st.integers(min_value=0, max_value=9))
        idx = draw(st.lists(st.integers(min_value=-3, max_value=40), min_size=n, max_size=n))
        scans = {}
        for j, i in enumerate(idx):
            key = f"{i}_{draw(label)}" + ("" if draw(st.booleans()) else f"_x{j}")      # extra fields, unique keys
            vals = [draw(number) for _ in range(S if draw(st.integers(0, 9)) else max(1, S - 1))]
            scans[key] = vals if draw(st.booleans()) else {"gain": 3, "signal": vals, "tags": ["a", {"b": None}]}
        beams[f"beam_{b}"] = scans
    return beams, draw(st.sampled_from([None, 1, 2])), draw(st.integers(min_value=1, max_value=4))


@st.composite
def broken_volumes(draw):
    """volumes() with, somewhere, a scan that cannot be converted, a key without a label field, a key without an integer
    prefix, or a repeated key."""
    beams, indent, L = draw(volumes())
    names = list(beams)
    for _ in range(draw(st.integers(min_value=1, max_value=3))):
        b = beams[draw(st.sampled_from(names))]
        kind = draw(st.sampled_from(["skipped", "nolabel", "badint", "nothing"]))
        pos = draw(st.integers(min_value=0, max_value=30))
        if kind == "skipped":
            b[f"{pos}_Health_s"] = {"amplitude": 1.5, "tags": [1, 2]}
        elif kind == "nolabel":
            b[f"{pos}"] = [0.5]
        elif kind == "badint":
            b[f"q{pos}_Health"] = [0.25]
    return beams, indent, L


@pytest.mark.filterwarnings("ignore:overflow encountered in cast")
@settings(max_examples=150, deadline=None, suppress_health_check=[HealthCheck.too_slow, HealthCheck.function_scoped_fixture])
@given(broken_volumes(), st.booleans())
def test_partly_usable_files_equal_oracle(tmp_path, case, repeat_key):
    beams, indent, L = case
    text = json.dumps(beams, indent=indent)
    first = next(iter(beams.values()))
    if repeat_key and first:                                         # the first beam's first key once more, with another value
        k = next(iter(first))
        again = json.dumps({k: [7.0]})[1:-1]
        at = text.index("{", text.index("{") + 1) + 1                # just inside the first beam's object
        text = text[:at] + again + ", " + text[at:]
    p = tmp_path / "v.json"
    p.write_text(text)
    ws, wl, wd = jsonload.signal_sets(str(p), L)
    lengths = {s.shape[-1] for s in ws}
    if len(lengths) <= 1:
        sets, labels, defects = dataio.json_signal_sets([str(p)], seq_length=L)
        assert len(sets) == len(ws)
        for i in range(len(ws)):
            np.testing.assert_array_equal(sets[i], ws[i])
            np.testing.assert_array_equal(labels[i], wl[i])
            np.testing.assert_array_equal(defects[i], wd[i])
    else:
        with pytest.raises(ValueError):
            dataio.json_signal_sets([str(p)], seq_length=L)


@pytest.mark.filterwarnings("ignore:overflow encountered in cast")
@settings(max_examples=150, deadline=None, suppress_health_check=[HealthCheck.too_slow, HealthCheck.function_scoped_fixture])
@given(volumes())
def test_native_loader_equals_oracle(tmp_path, case):
    beams, indent, L = case
    p = tmp_path / "v.json"
    p.write_text(json.dumps(beams, indent=indent))
    want = jsonload.load_beams(str(p))
    got = dataio.load_json_volume(str(p))
    assert [g["key"] for g in got] == [w[0] for w in want]
    for g, (_, sig, lab, dfx) in zip(got, want):
        rows = g["ragged"] if g["signals"] is None else list(g["signals"])
        assert len(rows) == len(sig)
        for a, b in zip(rows, sig):
            np.testing.assert_array_equal(a, b)                       # float32(float64(text)), inf / denormals included
        np.testing.assert_array_equal(g["labels"], lab)
        np.testing.assert_array_equal(g["defects"], dfx)
    ws, wl, wd = jsonload.signal_sets(str(p), L)
    lengths = {s.shape[-1] for s in ws}
    if len(lengths) <= 1:
        sets, labels, defects = dataio.json_signal_sets([str(p)], seq_length=L)
        assert len(sets) == len(ws)
        for i in range(len(ws)):
            np.testing.assert_array_equal(sets[i], ws[i])
            np.testing.assert_array_equal(labels[i], wl[i])
            np.testing.assert_array_equal(defects[i], wd[i])
    else:
        with pytest.raises(ValueError):
            dataio.json_signal_sets([str(p)], seq_length=L)
