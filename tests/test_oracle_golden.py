"""CPU: the oracle (restatement + nn port) against the reference's golden vectors."""
import pytest
import torch

from oracle import port, restatement as R
from helpers import GOLDEN_CASES, grad_rel_err, load_golden, rel_err


def _fwd(kind, kw):
    if kind == "transformer":
        return lambda sd, g: R.transformer_forward(sd, g["X"], g["y"], kw["num_heads"], kw["num_layers"])
    return lambda sd, g: R.rnn_encdec_forward(sd, g["X"], g["lengths"], kind, kw["num_layers"])


@pytest.mark.parametrize("name", list(GOLDEN_CASES))
def test_restatement_forward_matches_reference(name):
    kind, kw = GOLDEN_CASES[name]
    g = load_golden(name)
    logp = _fwd(kind, kw)(g["w0"], g)
    assert rel_err(logp, g["logp_train"]) < 1e-5
    assert torch.equal(logp.argmax(1), g["logp_train"].argmax(1))
    # the reference's eval/no_grad path gives the same numbers as train at dropout 0
    assert rel_err(g["logp_eval"], g["logp_train"]) < 1e-5
    loss = R.criterion(logp, g["y"])
    assert abs(float(loss) - g["loss"][0]) < 1e-5 * abs(g["loss"][0])


@pytest.mark.parametrize("name", list(GOLDEN_CASES))
def test_restatement_gradients_and_three_steps(name):
    kind, kw = GOLDEN_CASES[name]
    g = load_golden(name)
    fwd = _fwd(kind, kw)
    sd, bufs = dict(g["w0"]), {}
    for step in range(3):
        if step == 0:
            params = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
            loss = R.criterion(fwd(params, g), g["y"])
            loss.backward()
            scale = max(float(v.abs().max()) for v in g["g0"].values())
            for k, ref in g["g0"].items():
                assert grad_rel_err(params[k].grad, ref, scale) < 2e-5, k
            dead = [k for k, p in params.items() if p.grad is None]
            if kind != "transformer":
                assert dead == ["model.decoder.pre_output_layer.weight"]
        sd, loss, gnorm = R.train_step(sd, bufs, lambda p: fwd(p, g), g["y"], g["lr"])
        assert abs(loss - g["loss"][step]) < 2e-5 * abs(g["loss"][step])
        assert abs(gnorm - g["gnorm"][step]) < 1e-4 * abs(g["gnorm"][step])
    for k, ref in g["w3"].items():
        assert rel_err(sd[k], ref) < 2e-5, k


@pytest.mark.parametrize("name", list(GOLDEN_CASES))
def test_port_matches_reference(name):
    kind, kw = GOLDEN_CASES[name]
    g = load_golden(name)
    v_src = g["w0"]["model.src_embed.weight" if kind != "transformer" else "src_embedding.weight"].shape[0]
    v_tgt = g["w0"]["model.trg_embed.weight" if kind != "transformer" else "tgt_embedding.weight"].shape[0]
    m = port.build_port(kind, v_src, v_tgt, dropout=0.0, **kw)
    missing, unexpected = m.load_state_dict(g["w0"], strict=False)
    assert not unexpected and all(k.endswith(".pe") for k in missing)
    opt = torch.optim.SGD(m.parameters(), lr=g["lr"], momentum=0.9)
    for step in range(3):
        loss = port.reference_train_step(m, opt, g["X"], g["y"], g["lengths"])
        assert abs(float(loss) - g["loss"][step]) < 1e-5 * abs(g["loss"][step])
    for k, ref in g["w3"].items():
        assert rel_err(m.state_dict()[k], ref) < 1e-5, k


def test_quirks_pinned():
    """SURVEY.md section 0 quirks, checked on the restatement (itself pinned above)."""
    g = load_golden("lstm_small")
    kind, kw = GOLDEN_CASES["lstm_small"]
    base = R.rnn_encdec_forward(g["w0"], g["X"], g["lengths"], kind, kw["num_layers"])
    # pad fill value is invisible in the output when mask and lengths agree (quirk 3)
    emb = g["w0"]["model.src_embed.weight"][g["X"]]
    a, _ = R.encoder_forward(g["w0"], emb, g["lengths"], kind, 2, pad_fill=1.0)
    b, _ = R.encoder_forward(g["w0"], emb, g["lengths"], kind, 2, pad_fill=0.0)
    assert not torch.equal(a, b)
    # only trg_embed row 0 (<bos> -> <unk>) is touched (quirk 2)
    w = {k: v.clone() for k, v in g["w0"].items()}
    w["model.trg_embed.weight"][2:] += 1.0
    assert torch.equal(R.rnn_encdec_forward(w, g["X"], g["lengths"], kind, 2), base)
    # transformer: label leakage (quirk 7) - output depends on y
    t = load_golden("transformer_small")
    kw = GOLDEN_CASES["transformer_small"][1]
    o1 = R.transformer_forward(t["w0"], t["X"], t["y"], kw["num_heads"], kw["num_layers"])
    o2 = R.transformer_forward(t["w0"], t["X"], (t["y"] + 1) % 10 + 2, kw["num_heads"], kw["num_layers"])
    assert not torch.allclose(o1, o2)
