"""CPU oracle for the PAUT A-scan inference hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product
path: only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs may import it, and only as the
checker or the CPU baseline, never as the thing shipped.

The oracle restates, with plain fp32 torch/numpy ops on the CPU, the forward
pass and post-processing of the reference models (citations in each function
are relative to the reference repository root).  It is pinned against the
reference itself: ``tests/golden/make_golden.py`` imports the reference
classes from ``/root/reference`` (possible only in the build container), runs
them on the synthetic weights/inputs of ``oracle/synth.py`` (re-exported from the
package's ``synthetic.py``: data generators, no reference arithmetic) and commits
the outputs under ``tests/golden/``; ``tests/test_oracle_golden.py`` checks the
restatement against those files everywhere.

Modules: ``models`` (forwards of the ten model kinds), ``postprocess`` (predict()
loops, keep rules, integer sample indices), ``windowing`` (the two window rules),
``metrics`` (detection-level metrics and the difference matrix, pinned on known
answers from the reference's own functions: ``tests/test_oracle_next.py``).
"""
