"""CPU: host-side logic of the drop-in modules (construction, state_dict surface,
initialisation stream, flat-parameter views, refusal to compute on CPU)."""
import pytest
import torch

from helpers import GOLDEN_CASES, RNN_CASES, TRANSFORMER_CASES, load_golden
from slnlp_b200.vocab import Vocab
import model as dropin


def build(name, **extra):
    kind, kw = GOLDEN_CASES[name]
    g = load_golden(name)
    key = "model.src_embed.weight" if kind != "transformer" else "src_embedding.weight"
    tkey = "model.trg_embed.weight" if kind != "transformer" else "tgt_embedding.weight"
    cls = {"lstm": "EncoderDecoderLSTMAttn", "gru": "EncoderDecoderGRUAttn", "transformer": "Transformer"}[kind]
    torch.manual_seed(1)
    m = getattr(dropin, cls)(src_vocab=Vocab(size=g["w0"][key].shape[0]), tgt_vocab=Vocab(size=g["w0"][tkey].shape[0]),
                             batch_first=True, dropout=0.0, device=torch.device("cpu"), **kw, **extra)
    return m, g


@pytest.mark.parametrize("name", RNN_CASES)
def test_state_dict_surface_and_init_stream_match_reference(name):
    m, g = build(name)
    sd = m.state_dict()
    assert list(sd.keys()) == list(g["w0"].keys())            # names AND order (Checkpoint files)
    for k, ref in g["w0"].items():
        assert sd[k].shape == ref.shape, k
        # same torch.manual_seed(1) -> the same initial weights as the reference module
        assert torch.equal(sd[k], ref), k


@pytest.mark.parametrize("name", RNN_CASES)
def test_parameters_are_views_of_one_flat_buffer(name):
    m, g = build(name)
    flat = m._flat
    for n, p in m.named_parameters():
        assert p.data_ptr() == flat.data_ptr() + 4 * m._off[n]
    # load_state_dict writes through the views
    m.load_state_dict({k: v + 1 for k, v in g["w0"].items()})
    for n, p in m.named_parameters():
        assert torch.equal(p.data, g["w0"][n] + 1)
        assert p.data_ptr() == flat.data_ptr() + 4 * m._off[n]
    # the two directions of a layer are adjacent (one GEMM covers both)
    G, H = m.G, m.H
    assert m._off["model.encoder.rnn.weight_ih_l0_reverse"] == m._off["model.encoder.rnn.weight_ih_l0"] + G * H * m.E
    assert m._off["model.encoder.rnn.weight_hh_l0_reverse"] == m._off["model.encoder.rnn.weight_hh_l0"] + G * H * H


def test_no_cpu_fallback():
    m, g = build("lstm_small")
    with pytest.raises(RuntimeError, match="CUDA"):
        m(X=g["X"], y=g["y"], lengths=g["lengths"])


def test_constructor_surface():
    with pytest.raises(AssertionError):
        from slnlp_b200.rnn import RnnEncDecB200
        RnnEncDecB200(src_vocab=Vocab(size=5), tgt_vocab=Vocab(size=5), batch_first=True, rnn_type="rnn")
    m, _ = build("gru_small")
    assert m.bos_idx == 0 and m.src_pad == 1 and m.tgt_pad == 1   # <bos> -> <unk>
    assert m.to(torch.device("cpu")).device == torch.device("cpu")


@pytest.mark.parametrize("name", TRANSFORMER_CASES)
def test_transformer_state_dict_surface_and_init_stream(name):
    m, g = build(name)
    sd = m.state_dict()
    # parameters: the reference's names, order and (same seed) initial values
    assert [k for k in sd if not k.endswith(".pe")] == list(g["w0"].keys())
    for k, ref in g["w0"].items():
        assert torch.equal(sd[k], ref), k
    # buffers sit where the reference registers them (transformer.py:32-39; positional_encoding.py:31)
    keys = list(sd.keys())
    assert keys[:4] == ["src_embedding.weight", "src_pos_encoding.pe", "tgt_embedding.weight", "tgt_pos_encoding.pe"]
    assert sd["src_pos_encoding.pe"].shape == (5000, 1, m.E)
    E = m.E
    assert m._off["transformer.encoder.norm.bias"] == m._off["transformer.encoder.norm.weight"] + E
    with pytest.raises(RuntimeError, match="CUDA"):
        m(X=g["X"], y=g["y"], lengths=g["lengths"])
    assert dropin.EncoderDecoderTransformerAttn is dropin.Transformer
