"""GPU: the drop-in modules at the BASELINE.json sizes against fixtures generated from the REAL
reference (tests/golden/make_golden_baseline.py; cfg1, cfg2, cfg3 and the cfg4 model E1024/H512/L6 at
batch 50, ragged lengths): eval log-probs and greedy decode, then two fused training steps (loss,
gradient norm, sampled weights + checksums).  fp32 path: 1e-5 relative, identical argmax; bf16
tensor-core path: 2e-2 (north_star)."""
import pytest
import torch

pytestmark = pytest.mark.gpu

from helpers import (BASELINE_CASES, BF16_RTOL, FP32_RTOL, build_baseline_dropin, check_baseline_state,  # noqa: E402
                     load_baseline_golden, rel_err)


def _run(name, precision):
    from slnlp_b200.rnn import FusedTrainStep
    g = load_baseline_golden(name)
    dev = torch.device("cuda")
    m = build_baseline_dropin(name, dev, precision).to(dev)
    X, y, lengths = g["X"].cuda(), g["y"].cuda(), g["lengths"].cuda()
    m.eval()
    with torch.no_grad():
        logp = m(X=X, y=y, lengths=lengths)
    m.train()
    ts = FusedTrainStep(m, X.shape[0], X.shape[1], lr=g["lr"], momentum=0.9, max_norm=0.5)
    losses, norms = [], []
    for _ in range(2):
        losses.append(float(ts.step(X, y, lengths)[0]))
        norms.append(float(ts.grad_norm))
    return g, m, logp, losses, norms


@pytest.mark.parametrize("name", list(BASELINE_CASES))
def test_fp32_path_matches_the_reference_at_baseline_size(name):
    g, m, logp, losses, norms = _run(name, "fp32")
    assert rel_err(logp, g["logp_eval"]) < FP32_RTOL
    assert torch.equal(logp.argmax(1).cpu(), g["logp_eval"].argmax(1))      # greedy decode identical
    for step in range(2):
        assert abs(losses[step] - g["loss"][step]) < FP32_RTOL * abs(g["loss"][step])
        assert abs(norms[step] - g["gnorm"][step]) < 1e-4 * g["gnorm"][step]
    check_baseline_state(dict(m.named_parameters()), g, "w2", 2e-5)


@pytest.mark.parametrize("name", list(BASELINE_CASES))
def test_tensor_core_path_matches_the_reference_at_baseline_size(name):
    g, m, logp, losses, norms = _run(name, "bf16")
    assert rel_err(logp, g["logp_eval"]) < BF16_RTOL
    for step in range(2):
        assert abs(losses[step] - g["loss"][step]) < BF16_RTOL * abs(g["loss"][step])
        assert abs(norms[step] - g["gnorm"][step]) < 5e-2 * g["gnorm"][step]
    check_baseline_state(dict(m.named_parameters()), g, "w2", BF16_RTOL)


def test_first_step_gradients_match_the_reference_at_cfg4_model_size():
    """E1024 / H512 / L6 through the autograd route: sampled elements and checksums of every
    parameter gradient against the reference's (scale = the tensor's max |g|, floored at 1e-3 of
    the largest gradient)."""
    g = load_baseline_golden("cfg4s")
    dev = torch.device("cuda")
    m = build_baseline_dropin("cfg4s", dev).to(dev).train()
    X, y, lengths = g["X"].cuda(), g["y"].cuda(), g["lengths"].cuda()
    loss = torch.nn.functional.cross_entropy(m(X=X, y=y, lengths=lengths), y, ignore_index=1)
    loss.backward()
    assert abs(float(loss) - g["loss"][0]) < FP32_RTOL * abs(g["loss"][0])
    grads = {k: p.grad for k, p in m.named_parameters() if p.grad is not None}
    assert set(grads) == set(g["g0smp"])
    floor = 1e-3 * max(float(v[2]) for v in g["g0sum"].values())
    check_baseline_state(grads, g, "g0", 2e-5, floor=floor)
