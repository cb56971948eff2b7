"""ASL-Phono corpus front-end: a directory of per-sample JSON files -> composite-token ids.

Host-side restatement of the reference's torchtext pipeline
(dataset/builder/dataset_builder.py:66-223) without torchtext / commons, producing the
tensors the device path consumes (``SeqDataset``: X [N,T] int64 padded with <pad>=1,
lengths [N], y [N]):

  * one JSON file per sample: ``{"label": str, "frames": [{"phonology": {field: {"value": str}
    | null, ...}}, ...]}``; the file stem up to the first '-' is the sign's prefix, and only
    prefixes with at least ``samples_min_freq`` files are kept (dataset_builder.py:66-94);
  * every frame becomes ONE composite token built from the configured fields by one of the four
    ``composition_strategy`` variants (dataset_builder.py:137-223);
  * vocabularies are torchtext-0.6 ``Field.build_vocab`` order: <unk>=0, <pad>=1, then tokens by
    descending frequency, ties in ascending string order (torchtext/vocab.py sorts by token and
    then stable-sorts by count); the label is whitespace-tokenised like a sequential Field and
    its first token is the class.

torchtext is not installed in this image, so the vocabulary ORDER is "parity unpinned" against a
live torchtext; the composition strategies are pinned to vectors generated from the reference's
own ``DatasetBuilder.compose_*`` methods (tests/golden/phono_compose.json).
"""
from __future__ import annotations

import collections
import json
import os
from typing import Dict, Iterable, List, Sequence

import torch

from .data import SeqDataset
from .vocab import Vocab

STRATEGIES = ("all_values", "as_words", "as_words_norm", "as_sep_feat")


def _value(cell):
    """A field of a frame is {"value": ...} or null / "" (the reference rewrites null to "",
    dataset_builder.py:74); both count as absent."""
    return cell["value"] if cell else None


def _initials(cell):
    v = _value(cell)
    return "".join(k[0] for k in str(v).split("_")) if cell else ""


def _norm_field(field, cell):
    values = str(_value(cell)) if cell else ""
    if field.startswith("orientation") or field.startswith("movement"):
        parts = values.split("_")
        return "".join((("l" if "left" in parts else "r" if "right" in parts else "_"),
                        ("u" if "up" in parts else "d" if "down" in parts else "_"),
                        ("f" if "front" in parts else "b" if "back" in parts else "_")))
    return values


def compose(rows: Iterable[Dict], fields: Sequence[str], strategy: str = "as_words") -> List[str]:
    """Frames (phonology dicts) -> one token per frame (dataset_builder.py:137-223)."""
    assert strategy in STRATEGIES, f"Unknown composition strategy: '{strategy}'"
    if strategy == "all_values":      # fixed-width values joined by '-'
        return ["-".join(f"{(_value(row[f]) if row[f] else ''):<20}" for f in fields) for row in rows]
    if strategy == "as_words":        # initials of the '_'-separated parts: 'lb--ldf--L-'
        return ["-".join(_initials(row[f]) for f in fields) for row in rows]
    if strategy == "as_words_norm":   # orientation / movement as 3 fixed l/r, u/d, f/b slots
        return ["-".join(_norm_field(f, row[f]) for f in fields) for row in rows]
    return [str([_initials(row[f]) for f in fields]) for row in rows]   # as_sep_feat


def build_vocab(token_lists: Iterable[Iterable[str]]) -> Vocab:
    """torchtext-0.6 order: specials, then by descending count, ties by ascending token."""
    counter = collections.Counter()
    for toks in token_lists:
        counter.update(toks)
    ordered = sorted(sorted(counter.items(), key=lambda kv: kv[0]), key=lambda kv: kv[1], reverse=True)
    v = Vocab([w for w, _ in ordered if w not in ("<unk>", "<pad>")])
    v.freqs = counter
    return v


def read_corpus(dataset_dir: str, samples_min_freq: int = 1) -> List[Dict]:
    """Samples of the directory whose sign prefix has >= samples_min_freq files, in file-name order."""
    assert os.path.isdir(dataset_dir), "Invalid dataset directory"
    files = sorted(f for f in os.listdir(dataset_dir) if f.endswith(".json"))
    prefix = lambda f: os.path.splitext(f)[0].split("-")[0]
    counts = collections.Counter(prefix(f) for f in files)
    out = []
    for f in files:
        if counts[prefix(f)] < samples_min_freq:
            continue
        with open(os.path.join(dataset_dir, f)) as fh:
            d = json.load(fh)
        d["file"] = f
        out.append(d)
    return out


def build_dataset(dataset_dir: str, fields: Sequence[str], samples_min_freq: int = 1,
                  composition_strategy: str = "as_words", batch_first: bool = True, max_len: int = None,
                  **_ignored) -> SeqDataset:
    """DatasetBuilder.build + AslDataset(...).stoi() in one step: numericalised, padded tensors."""
    samples = read_corpus(dataset_dir, samples_min_freq)
    assert samples, f"no samples with >= {samples_min_freq} files per sign under {dataset_dir!r}"
    src = [compose([fr["phonology"] for fr in s["frames"]], fields, composition_strategy) for s in samples]
    tgt = [str(s["label"]).split() for s in samples]
    if max_len:
        src = [t[:max_len] for t in src]
    src_vocab, tgt_vocab = build_vocab(src), build_vocab(tgt)
    T = max(len(t) for t in src)
    pad = src_vocab.stoi["<pad>"]
    X = torch.full((len(src), T), pad, dtype=torch.int64)
    for i, toks in enumerate(src):
        X[i, :len(toks)] = torch.tensor([src_vocab.stoi[w] for w in toks], dtype=torch.int64)
    lengths = torch.tensor([len(t) for t in src], dtype=torch.int64)
    y = torch.tensor([tgt_vocab.stoi[t[0]] if t else tgt_vocab.stoi["<unk>"] for t in tgt], dtype=torch.int64)
    ds = SeqDataset(X, lengths, y, src_vocab, tgt_vocab, batch_first=batch_first)
    ds.files = [s["file"] for s in samples]
    return ds
