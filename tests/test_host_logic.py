"""CPU: host-side logic of the drop-in modules (construction, state_dict surface,
initialisation stream, flat-parameter views, refusal to compute on CPU)."""
import numpy as np
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


# ---------------------------------------------------------------- grid farm: several fits per GPU (threads)
from sklearn.base import BaseEstimator, ClassifierMixin  # noqa: E402


class _SlowClassifier(ClassifierMixin, BaseEstimator):
    """sklearn estimator whose fit takes a fixed wall time (stands in for a GPU-bound fit)."""

    def __init__(self, C=1.0, delay=0.15):
        self.C, self.delay = C, delay

    def fit(self, X, y):
        import time
        time.sleep(self.delay)
        self.classes_ = np.unique(y)
        self.majority_ = np.bincount(y).argmax()
        return self

    def predict(self, X):
        return np.full(len(X), self.majority_)

    def score(self, X, y):
        return float((self.predict(X) == y).mean()) + 1e-3 * self.C


def test_grid_farm_packs_fits_per_gpu_with_threads(tmp_path):
    import time
    from sklearn.linear_model import LogisticRegression
    from sklearn.model_selection import GridSearchCV
    from slnlp_b200.grid import GridSearchFarm
    rng = np.random.RandomState(0)
    X = rng.randn(120, 5)
    y = (X[:, 0] + 0.5 * X[:, 1] > 0).astype(int)
    grid = {"C": [0.01, 0.1, 1.0, 10.0]}
    # same table as sklearn's own search, packed 3 at a time, with the resume journal written by the parent
    journal = str(tmp_path / "fits.jsonl")
    gs = GridSearchFarm(LogisticRegression(max_iter=200), grid, cv=3, scoring="accuracy", backend="inline", n_gpus=1,
                        fits_per_gpu=3, resume_file=journal).fit(X, y)
    ref = GridSearchCV(LogisticRegression(max_iter=200), grid, cv=3, scoring="accuracy").fit(X, y)
    assert np.allclose(ref.cv_results_["mean_test_score"], gs.cv_results_["mean_test_score"])
    assert gs.best_params_ == ref.best_params_ and len(open(journal).read().splitlines()) == 12
    # the fits really overlap: 8 fits of 0.15 s, 4 at a time
    t0 = time.perf_counter()
    slow = GridSearchFarm(_SlowClassifier(), grid, cv=2, scoring=lambda e, X, y: e.score(X, y), backend="inline", n_gpus=1,
                          fits_per_gpu=4, refit=False).fit(X, y)
    packed = time.perf_counter() - t0
    assert slow.n_fits_ == 8 and packed < 0.9 * 8 * 0.15
    assert slow.best_params_ == {"C": 10.0}
    # a failing fit surfaces in the parent (error_score="raise")
    class Boom(_SlowClassifier):
        def fit(self, X, y):
            raise ValueError("boom")
    with pytest.raises(RuntimeError, match="boom"):
        GridSearchFarm(Boom(), grid, cv=2, scoring=lambda e, X, y: 0.0, backend="inline", n_gpus=1, fits_per_gpu=2,
                       refit=False).fit(X, y)


def test_factored_embedding_variant_surface():
    """F = 6 field tables: parameter names, shapes, adjacency in the flat buffer (one gather
    kernel walks them), and the usual refusal to compute on the CPU."""
    rows, widths = [29, 29, 29, 29, 90, 90], [8, 4, 8, 4, 16, 8]
    m = dropin.EncoderDecoderGRUAttn(src_vocab=Vocab(size=90), tgt_vocab=Vocab(size=7), batch_first=True, embedding_size=48,
                                     hidden_size=16, num_layers=1, dropout=0.0, device=torch.device("cpu"),
                                     src_field_vocab_sizes=rows, src_field_widths=widths)
    names = [n for n, _ in m.named_parameters() if "src_embed" in n]
    assert names == [f"model.src_embed.fields.{i}.weight" for i in range(6)]
    assert "model.src_embed.weight" not in dict(m.named_parameters())
    for i in range(5):
        a, b = names[i], names[i + 1]
        assert m._off[b] >= m._off[a] + rows[i] * widths[i] and m._off[b] - (m._off[a] + rows[i] * widths[i]) < 4
    assert float(dict(m.named_parameters())[names[4]][1].abs().max()) == 0.0      # padding_idx row of every table
    with pytest.raises(AssertionError):
        dropin.EncoderDecoderGRUAttn(src_vocab=Vocab(size=90), tgt_vocab=Vocab(size=7), batch_first=True, embedding_size=50,
                                     hidden_size=16, num_layers=1, dropout=0.0, src_field_vocab_sizes=rows, src_field_widths=widths)
    with pytest.raises(RuntimeError, match="CUDA"):
        m(X=torch.ones(2, 3, 6, dtype=torch.long), y=torch.ones(2, dtype=torch.long), lengths=torch.tensor([3, 3]))


def test_decoder_side_kernel_selection(monkeypatch):
    """Which decoder-side kernels a workspace takes (host logic only, no launch): the one-launch cell backward while every
    layer's contraction G*H x (D + H) stays within 2^18 weights and H, D are multiples of 16; the head kernel at batches
    <= 256; the query / bridge products inside it only on request and only up to H = 256 / 128."""
    import model as dropin
    from slnlp_b200.rnn import _Workspace
    from slnlp_b200.vocab import Vocab
    for k in ("SLNLP_DEC_HEAD", "SLNLP_DEC_HEAD_BWD", "SLNLP_DEC_HEAD_FUSE", "SLNLP_DEC_CELL_BWD", "SLNLP_DEC_TC", "SLNLP_DEC_FUSED"):
        monkeypatch.delenv(k, raising=False)

    def flags(kind, E, H, L, B=50, T=64):
        cls = dropin.EncoderDecoderLSTMAttn if kind == "lstm" else dropin.EncoderDecoderGRUAttn
        m = cls(src_vocab=Vocab(size=100), tgt_vocab=Vocab(size=20), batch_first=True, embedding_size=E, hidden_size=H,
                num_layers=L, dropout=0.1, device=torch.device("cpu"))
        ws = _Workspace(m, B, T, True)
        return dict(head=ws.dec_head, head_bwd=ws.dec_head_bwd, query=ws.fuse_query, bridge=ws.fuse_bridge, cell_bwd=ws.dec_cell_bwd)

    assert flags("lstm", 128, 128, 2) == dict(head=True, head_bwd=False, query=False, bridge=False, cell_bwd=True)     # cfg1
    assert flags("gru", 512, 256, 4)["cell_bwd"] is False          # cfg2: 768 x 1280 weights per contraction
    assert flags("lstm", 24, 20, 3)["cell_bwd"] is False           # H % 16 != 0
    assert flags("lstm", 128, 128, 2, B=4096) == dict(head=False, head_bwd=False, query=False, bridge=False, cell_bwd=False)
    monkeypatch.setenv("SLNLP_DEC_CELL_BWD", "2")
    monkeypatch.setenv("SLNLP_DEC_HEAD_FUSE", "1")
    assert flags("gru", 512, 256, 4) == dict(head=True, head_bwd=True, query=True, bridge=False, cell_bwd=True)
    assert flags("lstm", 128, 128, 2)["bridge"] is True
    monkeypatch.setenv("SLNLP_DEC_HEAD", "0")
    monkeypatch.setenv("SLNLP_DEC_CELL_BWD", "0")
    assert flags("lstm", 128, 128, 2) == dict(head=False, head_bwd=False, query=False, bridge=False, cell_bwd=False)
