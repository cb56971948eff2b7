"""Stage the UNMODIFIED reference model package for the benchmark's reference arm.

TEST / MEASUREMENT INFRASTRUCTURE ONLY (same import rule as oracle/restatement.py).

    python oracle/stage_reference.py            # run in the build container (needs /root/reference)

Copies ``/root/reference/model`` and ``/root/reference/dataset/constant`` (pure Python, ~930 lines)
byte for byte into ``baseline/_ref/``, which is git-ignored (reference sources never enter the
history) but NOT gpurun-ignored, so it travels to the GPU box with the snapshot.  ``bench.py --impl
reference`` then times the reference's OWN modules on the box's host cores (``cpu_baseline.kind =
"reference"``); if the directory is absent it falls back to the torch.nn port (``kind = "port"``).
``load()`` imports the staged package the way SURVEY.md section 8c describes: an empty ``dataset``
package shell so that ``dataset.constant`` loads without skorch / torchtext.
"""
import os
import shutil
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.environ.get("SLNLP_REFERENCE", "/root/reference")
DST = os.path.join(ROOT, "baseline", "_ref")


def stage():
    """Returns True if baseline/_ref holds the reference model package afterwards."""
    if os.path.isdir(os.path.join(SRC, "model")):
        for rel in ("model", os.path.join("dataset", "constant")):
            dst = os.path.join(DST, rel)
            if os.path.isdir(dst):
                shutil.rmtree(dst)
            shutil.copytree(os.path.join(SRC, rel), dst, ignore=shutil.ignore_patterns("__pycache__"))
    return os.path.isfile(os.path.join(DST, "model", "__init__.py"))


def available():
    return os.path.isfile(os.path.join(DST, "model", "__init__.py"))


def load():
    """Import the staged reference ``model`` package (it must win over the drop-in package of the
    same name: call this only in a process that has not imported the drop-in)."""
    if "model" in sys.modules and not getattr(sys.modules["model"], "__file__", "").startswith(DST):
        raise RuntimeError("the drop-in `model` package is already imported in this process")
    sys.path.insert(0, DST)
    pkg = types.ModuleType("dataset")
    pkg.__path__ = [os.path.join(DST, "dataset")]
    sys.modules["dataset"] = pkg
    import model
    return model


if __name__ == "__main__":
    print("staged" if stage() else "reference not available", DST)
