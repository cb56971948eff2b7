"""GPU: the drop-in RNN modules (C-ABI kernels) against the reference's golden vectors
and, at BASELINE sizes, against the oracle port on the same seeded inputs.

Tolerances (north_star): logits / loss within 1e-5 relative on the fp32 path, identical
argmax; gradients and post-step weights 2e-5 of the tensor scale (they accumulate one
more contraction)."""
import pytest
import torch

pytestmark = pytest.mark.gpu

from helpers import FP32_RTOL, GOLDEN_CASES, RNN_CASES, grad_rel_err, load_golden, rel_err  # noqa: E402


def build(name, dropout=0.0, **extra):
    import model as dropin
    from slnlp_b200.vocab import Vocab
    kind, kw = GOLDEN_CASES[name]
    g = load_golden(name)
    cls = {"lstm": dropin.EncoderDecoderLSTMAttn, "gru": dropin.EncoderDecoderGRUAttn}[kind]
    m = cls(src_vocab=Vocab(size=g["w0"]["model.src_embed.weight"].shape[0]),
            tgt_vocab=Vocab(size=g["w0"]["model.trg_embed.weight"].shape[0]),
            batch_first=True, dropout=dropout, device=torch.device("cuda"), **kw, **extra)
    m.load_state_dict(g["w0"])
    return m.to(torch.device("cuda")), g


@pytest.mark.parametrize("name", RNN_CASES)
def test_forward_matches_reference_golden(name):
    m, g = build(name)
    X, y, lengths = g["X"].cuda(), g["y"].cuda(), g["lengths"].cuda()
    m.eval()
    with torch.no_grad():
        logp = m(X=X, y=y, lengths=lengths)
    assert rel_err(logp, g["logp_eval"]) < FP32_RTOL
    assert torch.equal(logp.argmax(1).cpu(), g["logp_eval"].argmax(1))        # greedy decode identical
    m.train()
    logp_t = m(X=X, y=y, lengths=lengths)
    assert rel_err(logp_t, g["logp_train"]) < FP32_RTOL
    loss = torch.nn.functional.cross_entropy(logp_t, y, ignore_index=1)
    assert abs(float(loss) - g["loss"][0]) < FP32_RTOL * abs(g["loss"][0])
    # y's values never reach the output
    assert torch.equal(m(X=X, y=(y + 1) % 5 + 2, lengths=lengths), logp_t)


@pytest.mark.parametrize("name", RNN_CASES)
def test_autograd_path_with_stock_clip_and_sgd(name):
    """Drop-in route: loss.backward() through the autograd.Function, then the STOCK
    torch clip_grad_norm_ and SGD exactly as skorch would call them."""
    m, g = build(name)
    X, y, lengths = g["X"].cuda(), g["y"].cuda(), g["lengths"].cuda()
    m.train()
    opt = torch.optim.SGD(m.parameters(), lr=g["lr"], momentum=0.9, nesterov=False)
    scale = max(float(v.abs().max()) for v in g["g0"].values())
    for step in range(3):
        opt.zero_grad()
        loss = torch.nn.functional.cross_entropy(m(X=X, y=y, lengths=lengths), y, ignore_index=1)
        loss.backward()
        if step == 0:
            grads = dict(m.named_parameters())
            assert grads["model.decoder.pre_output_layer.weight"].grad is None    # dead branch
            for k, ref in g["g0"].items():
                assert grad_rel_err(grads[k].grad, ref, scale) < 2e-5, k
            tr = grads["model.trg_embed.weight"].grad
            assert float(tr[1:].abs().max()) == 0.0                              # only <bos>=<unk> row
        gn = torch.nn.utils.clip_grad_norm_(m.parameters(), max_norm=0.5, norm_type=2)
        assert abs(float(loss) - g["loss"][step]) < 2e-5 * abs(g["loss"][step])
        assert abs(float(gn) - g["gnorm"][step]) < 1e-4 * g["gnorm"][step]
        opt.step()
    sd = m.state_dict()
    for k, ref in g["w3"].items():
        assert rel_err(sd[k], ref) < 2e-5, k


@pytest.mark.parametrize("name", RNN_CASES)
@pytest.mark.parametrize("use_graph,host_batch", [(False, False), (True, False), (True, True)])
def test_fused_train_step_matches_reference_golden(name, use_graph, host_batch):
    """host_batch: the batch arrives as host tensors (packed into one pinned H2D copy per step)."""
    from slnlp_b200.rnn import FusedTrainStep
    m, g = build(name)
    m.train()
    B, T = g["X"].shape
    ts = FusedTrainStep(m, B, T, lr=g["lr"], momentum=0.9, max_norm=0.5, use_graph=use_graph)
    X, y, lengths = g["X"], g["y"], g["lengths"]
    if not host_batch:
        X, y, lengths = X.cuda(), y.cuda(), lengths.cuda()
    for step in range(3):
        loss = ts.step(X, y, lengths)
        assert abs(float(loss[0]) - g["loss"][step]) < 2e-5 * abs(g["loss"][step])
        assert abs(float(ts.grad_norm) - g["gnorm"][step]) < 1e-4 * g["gnorm"][step]
    sd = m.state_dict()
    for k, ref in g["w3"].items():
        assert rel_err(sd[k], ref) < 2e-5, k


def _synthetic(B, T, v_src, v_tgt, ragged, seed=3):
    g = torch.Generator().manual_seed(seed)
    X = torch.randint(2, v_src, (B, T), generator=g)
    lengths = torch.randint(5, T + 1, (B,), generator=g) if ragged else torch.full((B,), T)
    for b in range(B):
        X[b, lengths[b]:] = 1
    return X, lengths, torch.randint(2, v_tgt, (B,), generator=g)


@pytest.mark.parametrize("kind,E,H,L,ragged", [("lstm", 128, 128, 2, False), ("lstm", 128, 128, 2, True),
                                              ("gru", 64, 96, 3, True), ("gru", 512, 256, 4, False)])
def test_baseline_size_against_oracle_port(kind, E, H, L, ragged):
    """cfg1 (LSTM 128/128/2, B50, T64, V 4098/1026) and GRU shapes against the torch.nn
    port of the reference run on CPU with identical weights: logits, loss, argmax and
    two full training steps."""
    import model as dropin
    from oracle import port
    from slnlp_b200.rnn import FusedTrainStep
    from slnlp_b200.vocab import Vocab
    B, T, Vs, Vt = 50, 64, 4098, 1026
    torch.manual_seed(1)
    ref = port.build_port(kind, Vs, Vt, E, H, L, dropout=0.0)
    cls = dropin.EncoderDecoderLSTMAttn if kind == "lstm" else dropin.EncoderDecoderGRUAttn
    m = cls(src_vocab=Vocab(size=Vs), tgt_vocab=Vocab(size=Vt), batch_first=True, embedding_size=E,
            hidden_size=H, num_layers=L, dropout=0.0, device=torch.device("cuda"))
    m.load_state_dict(ref.state_dict())
    m = m.to(torch.device("cuda"))
    X, lengths, y = _synthetic(B, T, Vs, Vt, ragged)
    ref.eval()
    with torch.no_grad():
        want = ref(X=X, y=y, lengths=lengths)
    m.eval()
    with torch.no_grad():
        got = m(X=X.cuda(), y=y.cuda(), lengths=lengths.cuda())
    assert rel_err(got, want) < FP32_RTOL
    assert torch.equal(got.argmax(1).cpu(), want.argmax(1))
    m.train()
    ts = FusedTrainStep(m, B, T, lr=0.01)
    opt = torch.optim.SGD(ref.parameters(), lr=0.01, momentum=0.9)
    for step in range(2):
        want_loss = port.reference_train_step(ref, opt, X, y, lengths)
        got_loss = ts.step(X.cuda(), y.cuda(), lengths.cuda())
        assert abs(float(got_loss[0]) - float(want_loss)) < FP32_RTOL * abs(float(want_loss))
    rsd, sd = ref.state_dict(), m.state_dict()
    for k in rsd:
        assert rel_err(sd[k], rsd[k]) < 2e-5, k


def test_dropout_training_is_statistically_sane_and_eval_is_deterministic():
    m, g = build("lstm_small", dropout=0.3)
    X, y, lengths = g["X"].cuda(), g["y"].cuda(), g["lengths"].cuda()
    m.train()
    a = m(X=X, y=y, lengths=lengths).detach()
    b = m(X=X, y=y, lengths=lengths).detach()
    assert not torch.equal(a, b)                      # a fresh mask every call
    m.eval()
    with torch.no_grad():
        c, d = m(X=X, y=y, lengths=lengths), m(X=X, y=y, lengths=lengths)
    assert torch.equal(c, d)
    assert rel_err(c, g["logp_eval"]) < FP32_RTOL     # eval ignores dropout
    # backward replays the forward's masks: finite-difference check on one weight
    m.train()
    p = dict(m.named_parameters())["model.generator.proj.weight"]
    loss = torch.nn.functional.cross_entropy(m(X=X, y=y, lengths=lengths), y, ignore_index=1)
    loss.backward()
    assert torch.isfinite(p.grad).all() and float(p.grad.abs().max()) > 0


def test_input_validation_and_edge_cases():
    m, g = build("gru_small")
    X, y, lengths = g["X"].cuda(), g["y"].cuda(), g["lengths"].cuda()
    bad = lengths.clone()
    bad[0] = 0
    with pytest.raises(ValueError):
        m(X=X, y=y, lengths=bad)                      # pack_padded_sequence rejects length 0
    bad[0] = X.shape[1] + 1
    with pytest.raises(ValueError):
        m(X=X, y=y, lengths=bad)
    with pytest.raises(RuntimeError, match="CUDA"):
        m(X=X.cpu(), y=y, lengths=lengths.cpu()) if False else m.cpu()(X=g["X"], y=g["y"], lengths=g["lengths"])
    m = m.to(torch.device("cuda"))
    # batch of one, length one
    m.eval()
    with torch.no_grad():
        out = m(X=X[:1, :1], y=y[:1], lengths=torch.ones(1, dtype=torch.long, device="cuda"))
    assert out.shape == (1, 12) and abs(float(out.exp().sum()) - 1) < 1e-5
    # lengths omitted -> resolved from the pad count (util.resolve_lengths)
    with torch.no_grad():
        a = m(X=X, y=y)
        b = m(X=X, y=y, lengths=lengths)
    assert torch.equal(a, b)


@pytest.mark.parametrize("kind,E,H,L,ragged", [("lstm", 128, 128, 2, True),      # cfg1: persistent kernel, W_hh in TMEM
                                              ("gru", 512, 256, 4, True),       # cfg2: 8-CTA cluster kernel
                                              ("lstm", 128, 256, 2, False),     # LSTM on the cluster kernel
                                              ("lstm", 64, 512, 1, True),       # 16-CTA clusters
                                              ("gru", 64, 96, 2, True)])        # per-step TMA kernels (H % 32 == 0)
def test_tensor_core_path_within_bf16_tolerance_of_the_reference(kind, E, H, L, ragged):
    """precision='bf16' (tcgen05 recurrence with bf16 operands, TF32 TMA GEMMs, MUFU activations)
    against the torch.nn port of the reference on identical weights and inputs: logits and loss
    within north_star's 2e-2, through two full training steps."""
    import model as dropin
    from helpers import BF16_RTOL
    from oracle import port
    from slnlp_b200.rnn import FusedTrainStep
    from slnlp_b200.vocab import Vocab
    B, T, Vs, Vt = 50, 64, 4098, 1026
    torch.manual_seed(1)
    ref = port.build_port(kind, Vs, Vt, E, H, L, dropout=0.0)
    cls = dropin.EncoderDecoderLSTMAttn if kind == "lstm" else dropin.EncoderDecoderGRUAttn
    m = cls(src_vocab=Vocab(size=Vs), tgt_vocab=Vocab(size=Vt), batch_first=True, embedding_size=E,
            hidden_size=H, num_layers=L, dropout=0.0, device=torch.device("cuda"), precision="bf16")
    m.load_state_dict(ref.state_dict())
    m = m.to(torch.device("cuda"))
    X, lengths, y = _synthetic(B, T, Vs, Vt, ragged)
    ref.eval()
    with torch.no_grad():
        want = ref(X=X, y=y, lengths=lengths)
    m.eval()
    with torch.no_grad():
        got = m(X=X.cuda(), y=y.cuda(), lengths=lengths.cuda())
    assert rel_err(got, want) < BF16_RTOL
    m.train()
    ts = FusedTrainStep(m, B, T, lr=0.01)
    opt = torch.optim.SGD(ref.parameters(), lr=0.01, momentum=0.9)
    for step in range(2):
        want_loss = port.reference_train_step(ref, opt, X, y, lengths)
        got_loss = ts.step(X.cuda(), y.cuda(), lengths.cuda())
        assert abs(float(got_loss[0]) - float(want_loss)) < BF16_RTOL * abs(float(want_loss))
    # the update itself: every tensor within 2e-2 of its own scale (floor 1e-3) after two steps
    rsd, sd = ref.state_dict(), m.state_dict()
    for k in rsd:
        err = float((sd[k].cpu() - rsd[k]).abs().max()) / max(float(rsd[k].abs().max()), 1e-3)
        assert err < BF16_RTOL, k


@pytest.mark.parametrize("kind", ["lstm", "gru"])
def test_factored_six_field_embedding_variant(kind):
    """the true F = 6 factored phonological embedding as a model variant - one table per
    field (orientation / movement / handshape, dominant and non-dominant hand), gathered and concatenated by ONE
    kernel.  No reference counterpart: the oracle is the restatement with torch.cat of per-field embeddings.
    Forward 1e-5 + identical argmax, then two fused training steps (loss, gradient norm, every weight)."""
    import model as dropin
    from oracle import restatement as R
    from slnlp_b200.phono_fields import FIELD_CARD          # (27, 27, 27, 27, 88, 88)
    from slnlp_b200.rnn import FusedTrainStep
    from slnlp_b200.vocab import Vocab
    B, T, Vt, E, H, L = 9, 11, 13, 48, 32, 2
    rows = [c + 2 for c in FIELD_CARD]                      # + <unk>, <pad>
    widths = [8, 4, 8, 4, 16, 8]
    cls = dropin.EncoderDecoderLSTMAttn if kind == "lstm" else dropin.EncoderDecoderGRUAttn
    torch.manual_seed(3)
    m = cls(src_vocab=Vocab(size=max(rows)), tgt_vocab=Vocab(size=Vt), batch_first=True, embedding_size=E, hidden_size=H,
            num_layers=L, dropout=0.0, device=torch.device("cuda"), src_field_vocab_sizes=rows, src_field_widths=widths)
    sd0 = {k: v.clone() for k, v in m.state_dict().items()}
    assert [k for k in sd0 if "src_embed" in k] == [f"model.src_embed.fields.{i}.weight" for i in range(6)]
    assert [tuple(sd0[f"model.src_embed.fields.{i}.weight"].shape) for i in range(6)] == list(zip(rows, widths))
    m = m.to(torch.device("cuda"))
    g = torch.Generator().manual_seed(4)
    X = torch.stack([torch.randint(2, r, (B, T), generator=g) for r in rows], dim=-1)
    lengths = torch.randint(1, T + 1, (B,), generator=g)
    lengths[0] = T
    for b in range(B):
        X[b, lengths[b]:] = 1
    y = torch.randint(2, Vt, (B,), generator=g)
    want = R.rnn_encdec_forward(sd0, X, lengths, kind, L)
    m.eval()
    with torch.no_grad():
        got = m(X=X.cuda(), y=y.cuda(), lengths=lengths.cuda())
        got2 = m(X=X.cuda(), y=y.cuda())                    # lengths resolved from the first field's padding
    assert rel_err(got, want) < FP32_RTOL and torch.equal(got.argmax(1).cpu(), want.argmax(1)) and torch.equal(got, got2)
    m.train()
    ts = FusedTrainStep(m, B, T, lr=0.1)
    sd, bufs = dict(sd0), {}
    for step in range(2):
        sd, want_loss, want_norm = R.train_step(sd, bufs, lambda p: R.rnn_encdec_forward(p, X, lengths, kind, L), y, 0.1)
        loss = ts.step(X.cuda(), y.cuda(), lengths.cuda())
        assert abs(float(loss[0]) - want_loss) < 2e-5 * abs(want_loss)
        assert abs(float(ts.grad_norm) - want_norm) < 1e-4 * want_norm
    got_sd = m.state_dict()
    for k, v in sd.items():
        assert rel_err(got_sd[k], v) < 2e-5, k
    # the padding row of every field table received no gradient
    for i in range(6):
        assert torch.equal(got_sd[f"model.src_embed.fields.{i}.weight"][1].cpu(), sd0[f"model.src_embed.fields.{i}.weight"][1])
    with pytest.raises(ValueError):
        m(X=X[..., :5].cuda(), y=y.cuda(), lengths=lengths.cuda())


@pytest.mark.parametrize("kind,B,T,E,H,L", [("lstm", 512, 16, 256, 256, 3), ("lstm", 300, 40, 256, 384, 2), ("gru", 512, 16, 256, 256, 2)])
def test_large_batch_bf16_operand_path_against_the_reference(kind, B, T, E, H, L):
    """The family BASELINE.json configs[3] (data-parallel LSTM, batch 4096) runs on: CTA-pair bf16 GEMMs for the
    hoisted projections / dX / dW, bf16-operand per-step recurrent kernels writing bf16 copies of h_t and of
    d(pre-activations), bias sums over the bf16 copy, tensor-core decoder cell - against the torch.nn port of the
    reference on identical weights and ragged inputs (2e-2: logits, two training losses, every updated tensor)."""
    import model as dropin
    from helpers import BF16_RTOL
    from oracle import port
    from slnlp_b200.rnn import FusedTrainStep
    from slnlp_b200.vocab import Vocab
    Vs, Vt = 4098, 1026
    torch.manual_seed(7)
    ref = port.build_port(kind, Vs, Vt, E, H, L, dropout=0.0)
    cls = dropin.EncoderDecoderLSTMAttn if kind == "lstm" else dropin.EncoderDecoderGRUAttn
    m = cls(src_vocab=Vocab(size=Vs), tgt_vocab=Vocab(size=Vt), batch_first=True, embedding_size=E,
            hidden_size=H, num_layers=L, dropout=0.0, device=torch.device("cuda"), precision="bf16")
    m.load_state_dict(ref.state_dict())
    m = m.to(torch.device("cuda"))
    X, lengths, y = _synthetic(B, T, Vs, Vt, True)
    ref.eval()
    with torch.no_grad():
        want = ref(X=X, y=y, lengths=lengths)
    m.eval()
    with torch.no_grad():
        got = m(X=X.cuda(), y=y.cuda(), lengths=lengths.cuda())
    assert rel_err(got, want) < BF16_RTOL
    m.train()
    ts = FusedTrainStep(m, B, T, lr=0.01)
    # GRU: CTA-pair GEMMs around the tf32 per-step recurrent kernels (the bf16 step kernels' backward is LSTM-only)
    assert all(ts.ws.pair) and ts.ws.bf_step == (kind == "lstm"), "the shape must take the large-batch kernels this test is about"
    opt = torch.optim.SGD(ref.parameters(), lr=0.01, momentum=0.9)
    for step in range(2):
        want_loss = port.reference_train_step(ref, opt, X, y, lengths)
        got_loss = ts.step(X.cuda(), y.cuda(), lengths.cuda())
        assert abs(float(got_loss[0]) - float(want_loss)) < BF16_RTOL * abs(float(want_loss))
    rsd, sd = ref.state_dict(), m.state_dict()
    for k in rsd:
        err = float((sd[k].cpu() - rsd[k]).abs().max()) / max(float(rsd[k].abs().max()), 1e-3)
        assert err < BF16_RTOL, k


def test_large_batch_path_with_dropout_matches_the_fp32_path_on_the_same_masks():
    """Inter-layer dropout on the large-batch family: bf16 dropout from the bf16 layer output, keep mask as bits applied
    by the BPTT kernels while they read the gradient.  The Philox stream is a function of (seed, step, site, element),
    so the fp32 path of the same module draws the SAME masks: two training steps of both must agree within 2e-2."""
    import model as dropin
    from helpers import BF16_RTOL
    from slnlp_b200.rnn import FusedTrainStep
    from slnlp_b200.vocab import Vocab
    B, T, E, H, L, Vs, Vt = 512, 33, 1024, 512, 2, 4098, 1026      # every GEMM of the layer large enough for the pair kernel
    mods = []
    for prec in ("fp32", "bf16"):
        torch.manual_seed(11)
        m = dropin.EncoderDecoderLSTMAttn(src_vocab=Vocab(size=Vs), tgt_vocab=Vocab(size=Vt), batch_first=True, embedding_size=E,
                                          hidden_size=H, num_layers=L, dropout=0.3, device=torch.device("cuda"), precision=prec, seed=5)
        mods.append(m.to(torch.device("cuda")).train())
    X, lengths, y = _synthetic(B, T, Vs, Vt, True)
    steps = [FusedTrainStep(m, B, T, lr=0.05) for m in mods]
    assert any(steps[1].ws.masked) and any(steps[1].ws.skip_out32), "the bf16 module must take the masked-dropout kernels"
    for it in range(2):
        l32, l16 = (float(ts.step(X.cuda(), y.cuda(), lengths.cuda())[0]) for ts in steps)
        assert abs(l16 - l32) < BF16_RTOL * abs(l32), (it, l16, l32)
    sd32, sd16 = mods[0].state_dict(), mods[1].state_dict()
    for k in sd32:
        err = float((sd16[k] - sd32[k]).abs().max()) / max(float(sd32[k].abs().max()), 1e-3)
        assert err < BF16_RTOL, k


@pytest.mark.parametrize("kind,E,H,L,B,T,dropout,precision", [
    ("lstm", 128, 128, 2, 50, 64, 0.1, "fp32"),     # cfg1: bridge + query fused, both decoder cells
    ("lstm", 128, 128, 2, 50, 64, 0.1, "bf16"),
    ("gru", 64, 96, 3, 37, 20, 0.2, "fp32"),        # GRU gate gradients (x / h halves of the candidate gate), ragged rows
    ("lstm", 256, 256, 2, 50, 33, 0.1, "bf16"),     # query fused, bridge on the GEMM; K = 1024: two chunks
    ("gru", 48, 512, 1, 9, 12, 0.0, "fp32"),        # neither product fused; GH = 1536: a ragged last chunk
    ("lstm", 24, 20, 3, 6, 7, 0.3, "fp32"),         # H % 16 != 0: cells stay on the step kernels, scalar weight rows
])
def test_decoder_fused_kernels_match_the_unfused_chain(monkeypatch, kind, E, H, L, B, T, dropout, precision):
    """dec_head.cu (bridge -> tanh -> query -> attention -> decoder input, and its backward twin) and the one-launch
    decoder-cell backward of dec_cell.cu against the chain of GEMM / element-wise / step kernels they replace: same
    weights, same batch, same dropout masks (same seed and sites) - log-probs and every parameter gradient."""
    import model as dropin
    from slnlp_b200.vocab import Vocab
    Vs, Vt = 300, 40
    cls = dropin.EncoderDecoderLSTMAttn if kind == "lstm" else dropin.EncoderDecoderGRUAttn
    X, lengths, y = _synthetic(B, T, Vs, Vt, ragged=True)
    X, lengths, y = X.cuda(), lengths.cuda(), y.cuda()
    out = {}
    for fused in ("1", "2", "0"):       # 1: the default fused kernels; 2: + query / bridge products inside the head kernels
        monkeypatch.setenv("SLNLP_DEC_HEAD", "0" if fused == "0" else "1")
        monkeypatch.setenv("SLNLP_DEC_CELL_BWD", "0" if fused == "0" else "2")      # 2: at every supported shape
        monkeypatch.setenv("SLNLP_DEC_HEAD_FUSE", "1" if fused == "2" else "0")
        torch.manual_seed(5)
        m = cls(src_vocab=Vocab(size=Vs), tgt_vocab=Vocab(size=Vt), batch_first=True, embedding_size=E, hidden_size=H,
                num_layers=L, dropout=dropout, device=torch.device("cuda"), seed=11, precision=precision)
        m = m.to(torch.device("cuda"))
        m.train()
        logp = m(X=X, y=y, lengths=lengths)
        torch.nn.functional.cross_entropy(logp, y, ignore_index=1).backward()
        out[fused] = (logp.detach().clone(), {k: p.grad.clone() for k, p in m.named_parameters() if p.grad is not None})
    # the fused kernels are fp32 FMA; the chain runs tf32 MMAs on the tensor-core path
    tol = 2e-5 if precision == "fp32" else 5e-3
    scale = max(float(v.abs().max()) for v in out["0"][1].values())
    for fused in ("1", "2"):
        assert rel_err(out[fused][0], out["0"][0]) < tol
        assert set(out[fused][1]) == set(out["0"][1])
        for k, ref in out["0"][1].items():
            assert grad_rel_err(out[fused][1][k], ref, scale) < (5e-5 if precision == "fp32" else 2e-2), (fused, k)
