"""BASELINE-size golden vectors from the REAL reference modules (run in the build container).

    python tests/golden/make_golden_baseline.py          # needs /root/reference (read-only)

``make_golden.py`` pins the oracle at toy sizes with full tensors.  This script pins the
BASELINE.json configurations themselves (batch 50, len 64, V 4098 / 1026, ragged lengths):

    cfg1   EncoderDecoderLSTMAttn  E128  H128 L2        (configs[0])
    cfg2   EncoderDecoderGRUAttn   E512  H256 L4        (configs[1])
    cfg3   Transformer             E512  F256 L4 h8     (configs[2])
    cfg4s  EncoderDecoderLSTMAttn  E1024 H512 L6        (configs[3]'s model at batch 50; also the largest
                                                          member of the configs[4] grid)

The initial weights are NOT stored: the drop-in modules draw the reference's default-initialiser
stream, so ``torch.manual_seed(1)`` + construction reproduces them (tests/test_host_logic.py); the
fixture holds their per-tensor checksums so that a test can prove it.  Stored per case (~0.25 MB):
inputs, eval log-probs [50, 1026], two train losses / gradient norms, and per parameter tensor a
checksum triple (sum, sum |.|, max |.|) plus 64 sampled elements of the first-step gradient and of
the weights after two steps.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_golden import Vocab, _import_reference  # noqa: E402

B, T, VS, VT, LR = 50, 64, 4098, 1026, 0.01
CASES = {
    "cfg1": ("EncoderDecoderLSTMAttn", dict(embedding_size=128, hidden_size=128, num_layers=2)),
    "cfg2": ("EncoderDecoderGRUAttn", dict(embedding_size=512, hidden_size=256, num_layers=4)),
    "cfg3": ("Transformer", dict(embedding_size=512, hidden_size=256, num_layers=4, num_heads=8)),
    "cfg4s": ("EncoderDecoderLSTMAttn", dict(embedding_size=1024, hidden_size=512, num_layers=6)),
}
N_SAMPLE = 64


def make_inputs(seed=3):
    g = torch.Generator().manual_seed(seed)
    X = torch.randint(2, VS, (B, T), generator=g)
    lengths = torch.randint(5, T + 1, (B,), generator=g)
    lengths[0] = T
    for b in range(B):
        X[b, lengths[b]:] = 1
    return X, lengths, torch.randint(2, VT, (B,), generator=g)


def sample_index(name, numel):
    """Fixed pseudo-random element positions of a tensor (seeded by its name and size)."""
    seed = (sum(ord(c) * (i + 1) for i, c in enumerate(name)) * 2654435761 + numel) % (2 ** 31)
    g = torch.Generator().manual_seed(seed)
    return torch.randint(0, numel, (min(N_SAMPLE, numel),), generator=g)


def checksum(t):
    t = t.detach().double().reshape(-1)
    return np.array([float(t.sum()), float(t.abs().sum()), float(t.abs().max())])


def run_case(ref_model, name):
    cls, kw = CASES[name]
    torch.manual_seed(1)
    dev = torch.device("cpu")
    m = getattr(ref_model, cls)(src_vocab=Vocab(VS), tgt_vocab=Vocab(VT), batch_first=True, dropout=0.0,
                                device=dev, **kw).to(dev)
    X, lengths, y = make_inputs()
    out = {"X": X.numpy(), "lengths": lengths.numpy(), "y": y.numpy(), "lr": np.float32(LR)}
    names = [k for k, _ in m.named_parameters()]
    out["names"] = np.array(names)
    for k, p in m.named_parameters():
        out["w0sum/" + k] = checksum(p)
    m.eval()
    with torch.no_grad():
        out["logp_eval"] = m(X=X, y=y, lengths=lengths).numpy().copy()
    m.train()
    opt = torch.optim.SGD(m.parameters(), lr=LR, momentum=0.9, nesterov=False)
    crit = torch.nn.CrossEntropyLoss(ignore_index=1)
    for step in range(2):
        opt.zero_grad()
        loss = crit(m(X=X, y=y, lengths=lengths), y)
        loss.backward()
        if step == 0:
            for k, p in m.named_parameters():
                if p.grad is not None:
                    out["g0sum/" + k] = checksum(p.grad)
                    out["g0smp/" + k] = p.grad.detach().reshape(-1)[sample_index(k, p.numel())].numpy().copy()
        out[f"loss{step}"] = np.float32(loss.item())
        out[f"gnorm{step}"] = np.float32(torch.nn.utils.clip_grad_norm_(m.parameters(), max_norm=0.5, norm_type=2).item())
        opt.step()
    for k, p in m.named_parameters():
        out["w2sum/" + k] = checksum(p)
        out["w2smp/" + k] = p.detach().reshape(-1)[sample_index(k, p.numel())].numpy().copy()
    np.savez_compressed(os.path.join(HERE, f"baseline_{name}.npz"), **out)
    print(name, "loss", float(out["loss0"]), float(out["loss1"]), "gnorm", float(out["gnorm0"]), float(out["gnorm1"]),
          flush=True)


if __name__ == "__main__":
    ref = _import_reference()
    for case in (sys.argv[1:] or list(CASES)):
        run_case(ref, case)
