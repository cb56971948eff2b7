"""Minimal torchtext-0.6-shaped vocabulary (itos list, defaultdict stoi -> unk).

The reference builds its vocabularies with torchtext Fields
(dataset/builder/dataset_builder.py:100-135): specials first (<unk>=0, <pad>=1),
then tokens by descending frequency, ``stoi`` a defaultdict that maps unknown
strings (including '<bos>') to 0.
"""
import collections

UNK_WORD, PAD_WORD, BOS_WORD, EOS_WORD = "<unk>", "<pad>", "<bos>", "<eos>"


class Vocab:
    def __init__(self, tokens=(), size=None):
        self.itos = [UNK_WORD, PAD_WORD] + list(tokens)
        if size is not None:
            self.itos += [f"tok{i}" for i in range(len(self.itos), size)]
        self.stoi = collections.defaultdict(int, {w: i for i, w in enumerate(self.itos)})

    def __len__(self):
        return len(self.itos)
