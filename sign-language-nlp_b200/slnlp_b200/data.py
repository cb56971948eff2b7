"""Synthetic ASL-Phono-shaped data.

The reference composes the six phonology fields of a frame (orientation / movement /
handshape for the dominant and non-dominant hand) into ONE string token
(``composition_strategy: as_words``, dataset/builder/dataset_builder.py:169-182) and
numericalises it with a torchtext vocabulary: specials first (<unk>=0, <pad>=1), then
tokens by descending frequency, ties in lexicographic order.  The corpus itself is not
available offline, so this module reproduces the OUTPUT CONTRACT of that front-end:
``X [N,T] int64`` composite ids padded with 1, ``lengths [N]``, ``y [N]`` labels >= 2.
The per-field Zipf exponent (2.4) is chosen so that ~5% of frames fall outside the 4096
most frequent composites (-> <unk>), a long tail like a real sign corpus.

The generator itself lives in the torch-only ``phono_synth`` module (no CUDA library needed).
"""
from __future__ import annotations

import torch

from phono_synth import (FIELD_CARD, FIELDS, Vocab, compose_as_words, make_fields,  # noqa: F401
                         synthetic_dataset)


# --------------------------------------------------------------------------------------
# Output contract of the reference's AslDataset / AslSliceDataset (dataset/asl_dataset.py:
# 9-303) over already-numericalised data: what main.py, helper.py and skorch touch.
# --------------------------------------------------------------------------------------
import numpy as np


class SeqSlice:
    """One column of a SeqDataset restricted to ``indices`` (skorch SliceDataset contract:
    ``len``, ``shape``, integer / array indexing, so sklearn's CV splitters can index it).
    idx 0 = X: items are (tokens, length, label) triples as the reference's collate expects
    (helper.py:293-304); idx 1 = y."""

    def __init__(self, dataset, idx=0, indices=None):
        self.dataset, self.idx = dataset, idx
        self.indices = None if indices is None else np.asarray(indices, dtype=np.int64)

    @property
    def indices_(self):
        return np.arange(len(self.dataset)) if self.indices is None else self.indices

    def __len__(self):
        return len(self.indices_)

    @property
    def shape(self):
        return (len(self),)

    @property
    def dtype(self):
        return np.dtype("O") if self.idx == 0 else np.dtype("int64")

    def __getitem__(self, i):
        if isinstance(i, tuple):        # sklearn's _array_indexing asks for array[key, ...]
            i = i[0]
        if isinstance(i, (int, np.integer)):
            j = int(self.indices_[i])
            d = self.dataset
            return (d.tokens[j].tolist(), int(d.lengths[j]), int(d.labels_[j])) if self.idx == 0 else int(d.labels_[j])
        if isinstance(i, slice):
            return SeqSlice(self.dataset, self.idx, self.indices_[i])
        i = np.asarray(i)
        if i.dtype == bool:
            i = np.flatnonzero(i)
        return SeqSlice(self.dataset, self.idx, self.indices_[i])

    def to_array(self):                                        # asl_dataset.py:288-303
        d, ind = self.dataset, self.indices_
        if self.idx == 1:
            return d.labels_[ind].numpy().copy()
        return d.tokens[ind].numpy().copy()

    # what the B200 estimator stages on the device (no per-item python lists)
    def tensors(self):
        d, ind = self.dataset, torch.from_numpy(self.indices_)
        return d.tokens[ind], d.lengths[ind], d.labels_[ind]

    def cpu(self):
        return self


class SeqDataset:
    """AslDataset stand-in over numericalised tensors: X [N,T] int64 padded with <pad>=1,
    lengths [N], y [N]; vocabularies torchtext-shaped (vocab.py)."""

    def __init__(self, X, lengths, y, src_vocab, tgt_vocab, batch_first=True):
        self.tokens = torch.as_tensor(X, dtype=torch.int64).contiguous()
        self.lengths = torch.as_tensor(lengths, dtype=torch.int64).contiguous()
        self.labels_ = torch.as_tensor(y, dtype=torch.int64).contiguous()
        assert self.tokens.dim() == 2 and len(self.tokens) == len(self.lengths) == len(self.labels_)
        self.vocab_X, self.vocab_y, self.batch_first = src_vocab, tgt_vocab, batch_first

    @classmethod
    def synthetic(cls, **kw):
        d = synthetic_dataset(**kw)
        return cls(d["X"], d["lengths"], d["y"], d["src_vocab"], d["tgt_vocab"])

    def __len__(self):
        return len(self.tokens)

    def __getitem__(self, i):
        if isinstance(i, (list, tuple, np.ndarray)):
            return [self[int(j)] for j in i]
        return ((self.tokens[i].tolist(), int(self.lengths[i])), int(self.labels_[i]))

    def X(self):
        return SeqSlice(self, 0)

    def y(self):
        return SeqSlice(self, 1)

    def stoi(self):
        return self

    def labels(self, fmt="i"):                                 # asl_dataset.py:210-213
        assert fmt in ("i", "s"), "Unknown format"
        n = len(self.vocab_y)
        return list(range(n)) if fmt == "i" else list(self.vocab_y.itos[:n])

    def _subset(self, ind):
        ind = torch.as_tensor(ind, dtype=torch.int64)
        return SeqDataset(self.tokens[ind], self.lengths[ind], self.labels_[ind], self.vocab_X, self.vocab_y,
                          self.batch_first)

    def truncated(self, length):                               # asl_dataset.py:215-218
        return self._subset(torch.arange(min(length, len(self))))

    def split(self, lengths, indices_only=False, seed=None):   # asl_dataset.py:220-253 (torch random_split)
        from torch.utils.data import random_split
        total = len(self)
        if not isinstance(lengths, list):
            lengths = [lengths]
        lengths = [round(l * total) if isinstance(l, float) else l for l in lengths]
        assert sum(lengths) <= total
        if total - sum(lengths) > 0:
            lengths.append(total - sum(lengths))
        gen = torch.Generator().manual_seed(seed) if seed else None
        parts = random_split(range(total), lengths, generator=gen)
        return [list(p.indices) if indices_only else self._subset(list(p.indices)) for p in parts]
