"""Build product oracle/_ref/: the reference's own model files, copied VERBATIM from /root/reference so that the CPU
arm of bench.py (`--impl reference`, `cpu_baseline`) can time the real reference classes on the GPU box, where
/root/reference does not exist.  oracle/_ref/ is git-ignored (never part of the history) and travels to the GPU box
with the snapshot like the built .so files.  Run by __graft_entry__.build() when /root/reference is present.

    python oracle/build_ref.py

Test infrastructure / baseline only: nothing in the product path imports oracle/.
"""
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("PAUT_REFERENCE", "/root/reference")
OUT = os.path.join(HERE, "_ref")

FILES = [
    ("signals/multisignalNN/NN_models.py", "NN_models.py"),
    ("SignalSequenceDetection/model.py", "model.py"),
    ("SignalSequenceDetection/enhanced_model.py", "enhanced_model.py"),
    ("SignalSequenceDetection/two_stage_model.py", "two_stage_model.py"),
]


def main():
    if not os.path.isdir(REF):
        print(f"build_ref: {REF} is not available here; oracle/_ref is left as it is")
        return 0
    os.makedirs(OUT, exist_ok=True)
    for src, dst in FILES:
        shutil.copyfile(os.path.join(REF, src), os.path.join(OUT, dst))
    # DefectDetectionModel lives in a script that trains at import: only its class source range travels
    # (signals/MSC_Conv1D_training.py:50-89)
    lines = open(os.path.join(REF, "signals", "MSC_Conv1D_training.py")).read().split("\n")
    with open(os.path.join(OUT, "msc_conv1d_model.py"), "w") as f:
        f.write("import torch\nimport torch.nn as nn\n\n" + "\n".join(lines[49:89]) + "\n")
    print(f"build_ref: {len(FILES) + 1} reference model files -> {OUT}")
    return 0


if __name__ == "__main__":
    sys.exit(main())
