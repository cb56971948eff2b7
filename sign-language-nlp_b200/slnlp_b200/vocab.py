"""Minimal torchtext-0.6-shaped vocabulary (itos list, defaultdict stoi -> unk).

The reference builds its vocabularies with torchtext Fields
(dataset/builder/dataset_builder.py:100-135): specials first (<unk>=0, <pad>=1),
then tokens by descending frequency, ``stoi`` a defaultdict that maps unknown
strings (including '<bos>') to 0.
"""
from phono_synth import BOS_WORD, EOS_WORD, PAD_WORD, UNK_WORD, Vocab  # noqa: F401  (torch-only home)
