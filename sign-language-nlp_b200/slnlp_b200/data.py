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
"""
from __future__ import annotations


import torch

from .vocab import Vocab

FIELD_CARD = (27, 27, 27, 27, 88, 88)   # orientation dh/ndh, movement dh/ndh, handshape dh/ndh
FIELDS = ("orientation_dh", "orientation_ndh", "movement_dh", "movement_ndh", "handshape_dh", "handshape_ndh")


def _zipf(card, n, gen, a=2.4):
    w = 1.0 / torch.arange(1, card + 1, dtype=torch.float64) ** a
    return torch.multinomial(w / w.sum(), n, replacement=True, generator=gen)


def make_fields(n_seq, T, seed=1):
    """[n_seq, T, 6] raw field ids; non-dominant-hand fields are empty (id 0) w.p. 0.5."""
    g = torch.Generator().manual_seed(seed)
    cols = [_zipf(c, n_seq * T, g) for c in FIELD_CARD]
    f = torch.stack(cols, dim=1).view(n_seq, T, 6)
    ndh_empty = torch.rand(n_seq, T, generator=g) < 0.5
    for j in (1, 3, 5):
        f[:, :, j] = torch.where(ndh_empty, torch.zeros_like(f[:, :, j]), f[:, :, j])
    return f, g


def compose_as_words(fields, v_src_cap=4098):
    """6-tuple -> composite id by descending frequency after the two specials
    (torchtext order), capped at ``v_src_cap`` (the rest -> <unk> = 0)."""
    n, T, F = fields.shape
    mult = torch.tensor([1, 100, 100 ** 2, 100 ** 3, 100 ** 4, 100 ** 5], dtype=torch.int64)
    key = (fields.view(-1, F) * mult).sum(1)
    uniq, inv, cnt = torch.unique(key, return_inverse=True, return_counts=True)
    # descending count, ties by key ("lexicographic")
    order = sorted(range(len(uniq)), key=lambda i: (-int(cnt[i]), int(uniq[i])))
    rank = torch.empty(len(uniq), dtype=torch.int64)
    rank[torch.tensor(order)] = torch.arange(len(uniq))
    ids = rank[inv] + 2
    ids = torch.where(ids < v_src_cap, ids, torch.zeros_like(ids))
    tokens = [f"w{int(uniq[i])}" for i in order[:v_src_cap - 2]]
    return ids.view(n, T), tokens


def synthetic_dataset(n_seq=5000, T=64, v_src=4098, v_tgt=1026, ragged=False, seed=1):
    """Returns dict(X, lengths, y, src_vocab, tgt_vocab, fields)."""
    fields, g = make_fields(n_seq, T, seed)
    X, tokens = compose_as_words(fields, v_src)
    lengths = torch.randint(5, T + 1, (n_seq,), generator=g) if ragged else torch.full((n_seq,), T, dtype=torch.int64)
    pad = torch.arange(T).unsqueeze(0) >= lengths.unsqueeze(1)
    X = torch.where(pad, torch.ones_like(X), X)
    y = torch.randint(2, v_tgt, (n_seq,), generator=g)
    return dict(X=X.contiguous(), lengths=lengths, y=y, fields=fields,
                src_vocab=Vocab(tokens, size=v_src), tgt_vocab=Vocab(size=v_tgt))


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
