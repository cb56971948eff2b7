"""Golden vectors for the ASL-Phono composition strategies, from the REFERENCE's own code.

    python tests/golden/make_phono_golden.py          # needs /root/reference (read-only)

``dataset/builder/dataset_builder.py`` imports commons-python and torchtext at module level (both
absent here); they are replaced by empty module shells so that ``DatasetBuilder`` imports - its
``compose_*`` methods are plain Python.  Seeded synthetic frames (field values drawn from the
ASL-Phono value vocabulary shape: '_'-joined direction words, handshape names, nulls) go through all
four strategies; inputs and outputs are stored in ``phono_compose.json``.
"""
import json
import os
import random
import sys
import types

REF = os.environ.get("SLNLP_REFERENCE", "/root/reference")
OUT = os.path.dirname(os.path.abspath(__file__))
FIELDS = ["orientation_dh", "orientation_ndh", "movement_dh", "movement_ndh", "handshape_dh", "handshape_ndh"]


def _import_builder():
    sys.path.insert(0, REF)
    for name, attrs in (("commons", ()), ("commons.log", ("auto_log_progress", "log")),
                        ("commons.util", ("exists", "filename", "filter_files", "read_json", "save_items", "get_hash")),
                        ("torchtext", ()), ("torchtext.data", ("Field", "TabularDataset", "interleave_keys"))):
        m = types.ModuleType(name)
        for a in attrs:
            setattr(m, a, None)
        sys.modules[name] = m
    pkg = types.ModuleType("dataset")
    pkg.__path__ = [os.path.join(REF, "dataset")]
    sys.modules["dataset"] = pkg
    from dataset.builder.dataset_builder import DatasetBuilder
    return DatasetBuilder


def make_rows(n, seed):
    rng = random.Random(seed)
    dirs = [["left", "right", None], ["up", "down", None], ["front", "back", None]]
    shapes = ["A", "B", "C", "5", "L", "open_b", "bent_v", "flat_o", "1", "claw_5"]
    rows = []
    for _ in range(n):
        row = {}
        for f in FIELDS:
            if f.endswith("ndh") and rng.random() < 0.5:
                row[f] = rng.choice([None, ""])     # null in the JSON, or "" after the reference's rewrite
                continue
            if f.startswith("handshape"):
                row[f] = {"value": rng.choice(shapes)}
            else:
                parts = [p for p in (rng.choice(d) for d in dirs) if p]
                rng.shuffle(parts)
                row[f] = {"value": "_".join(parts)} if parts else None
        rows.append(row)
    return rows


def main():
    B = _import_builder()()
    rows = make_rows(40, seed=7)
    out = {"fields": FIELDS, "rows": rows}
    for strategy in ("all_values", "as_words", "as_words_norm", "as_sep_feat"):
        out[strategy] = B.preprocess_src(rows, FIELDS, strategy)
    # a field subset in another order, as the YAML's `fields:` list allows
    sub = ["handshape_dh", "movement_dh", "orientation_dh"]
    out["subset_fields"] = sub
    out["subset_as_words"] = B.preprocess_src(rows, sub, "as_words")
    with open(os.path.join(OUT, "phono_compose.json"), "w") as f:
        json.dump(out, f, indent=0)
    print("wrote phono_compose.json:", out["as_words"][:3], out["as_words_norm"][:2])


if __name__ == "__main__":
    main()
