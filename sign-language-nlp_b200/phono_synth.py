"""Torch-only synthetic ASL-Phono-shaped corpus + the torchtext-shaped vocabulary.

This module imports nothing but torch, so the benchmark's reference (CPU) arm and the oracle
tests build the SAME synthetic inputs as the B200 arm without loading libslnlp_b200.so (importing
the ``slnlp_b200`` package loads the CUDA library).  ``slnlp_b200.data`` / ``slnlp_b200.vocab``
re-export these names.

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
import collections

import torch

# ---- vocabulary (dataset/builder/dataset_builder.py:100-135: torchtext-0.6 order, specials first,
#      stoi a defaultdict that maps unknown strings - '<bos>' included - to <unk> = 0)
UNK_WORD, PAD_WORD, BOS_WORD, EOS_WORD = "<unk>", "<pad>", "<bos>", "<eos>"


class Vocab:
    def __init__(self, tokens=(), size=None):
        self.itos = [UNK_WORD, PAD_WORD] + list(tokens)
        if size is not None:
            self.itos += [f"tok{i}" for i in range(len(self.itos), size)]
        self.stoi = collections.defaultdict(int, {w: i for i, w in enumerate(self.itos)})

    def __len__(self):
        return len(self.itos)


# ---- synthetic corpus
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


