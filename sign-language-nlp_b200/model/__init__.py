"""Drop-in ``model`` package: the names the reference's YAML resolves with
``pydoc.locate`` (helper.py:93; config/*.yaml:28), backed by the B200 kernels.

    model.EncoderDecoderLSTMAttn   (model/encoder_decoder_lstm_attn.py:4-6)
    model.EncoderDecoderGRUAttn    (model/encoder_decoder_gru_attn.py:4-6)
    model.Transformer              (model/transformer.py:9-109)
    model.EncoderDecoderTransformerAttn  - north_star's name for the Transformer
      (the reference file of that name is dead code)
"""
from slnlp_b200.rnn import RnnEncDecB200


class EncoderDecoderLSTMAttn(RnnEncDecB200):
    def __init__(self, **kwargs):
        super().__init__(rnn_type="lstm", **kwargs)


class EncoderDecoderGRUAttn(RnnEncDecB200):
    def __init__(self, **kwargs):
        super().__init__(rnn_type="gru", **kwargs)


try:  # the Transformer kernels land after the RNN path
    from slnlp_b200.transformer import TransformerB200 as Transformer
    EncoderDecoderTransformerAttn = Transformer
except ImportError:  # pragma: no cover
    pass
