"""Deterministic synthetic weights and inputs used by the tests and golden fixtures.

The generators live in the package (``defectdetection_viaobjectdetection_b200/synthetic.py``) so that ``bench.py``
can build its synthetic volume and random-init weights without importing anything from ``oracle/``; this module
re-exports them for the test infrastructure (``tests/``, ``tests/golden/make_golden.py``)."""
from defectdetection_viaobjectdetection_b200.synthetic import *  # noqa: F401,F403
from defectdetection_viaobjectdetection_b200.synthetic import (KINDS, sinusoidal_pe, state_spec, synth_paut_sets,  # noqa: F401
                                                               synth_state_dict, synth_tensor, synth_uniform_sets)
