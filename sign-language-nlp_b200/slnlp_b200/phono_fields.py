"""The six phonological fields of an ASL-Phono frame (dataset/builder/dataset_builder.py:169-182 joins them into
one token; the factored embedding variant of the RNN models keeps them apart: rnn.py ``src_field_vocab_sizes``)."""
from phono_synth import FIELD_CARD, FIELDS  # noqa: F401  (orientation / movement / handshape, dominant and non-dominant hand)
