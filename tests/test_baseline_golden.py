"""CPU: BASELINE-size fixtures generated from the real reference (tests/golden/make_golden_baseline.py):
the drop-in modules reproduce the reference's seed-1 initial weights at cfg1..cfg4 sizes, and the
oracle port, started from those weights, reproduces the reference's log-probs, losses, gradient norms
and two-step weights - so the GPU tests that compare against the port at full size are one hop from
the reference, not two."""
import numpy as np
import pytest
import torch

from helpers import (BASELINE_CASES, BASELINE_VS, BASELINE_VT, build_baseline_dropin, check_baseline_state, checksum,
                     load_baseline_golden, rel_err)
from oracle import port


@pytest.mark.parametrize("name", list(BASELINE_CASES))
def test_dropin_initial_weights_equal_the_reference_at_baseline_size(name):
    g = load_baseline_golden(name)
    m = build_baseline_dropin(name, torch.device("cpu"))
    assert [n for n, _ in m.named_parameters()] == g["names"]
    for k, p in m.named_parameters():
        assert np.array_equal(checksum(p), g["w0sum"][k]), k


@pytest.mark.parametrize("name", list(BASELINE_CASES))
def test_port_matches_reference_at_baseline_size(name):
    kind, kw = BASELINE_CASES[name]
    g = load_baseline_golden(name)
    m = build_baseline_dropin(name, torch.device("cpu"))
    ref = port.build_port(kind, BASELINE_VS, BASELINE_VT, dropout=0.0, **kw)
    missing, unexpected = ref.load_state_dict(m.state_dict(), strict=False)
    assert not unexpected and all(k.endswith(".pe") for k in missing)
    ref.eval()
    with torch.no_grad():
        logp = ref(X=g["X"], y=g["y"], lengths=g["lengths"])
    assert rel_err(logp, g["logp_eval"]) < 1e-5
    assert torch.equal(logp.argmax(1), g["logp_eval"].argmax(1))
    opt = torch.optim.SGD(ref.parameters(), lr=g["lr"], momentum=0.9)
    for step in range(2):
        loss = port.reference_train_step(ref, opt, g["X"], g["y"], g["lengths"])
        assert abs(float(loss) - g["loss"][step]) < 1e-5 * abs(g["loss"][step])
    check_baseline_state(dict(ref.named_parameters()), g, "w2", 1e-5)
