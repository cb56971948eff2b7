"""Generate golden vectors from the REAL reference modules (run in the build container).

    python tests/golden/make_golden.py          # needs /root/reference (read-only)

The reference ships no tests or known-answer vectors, so the
pins are outputs of the reference itself: for a few small, seeded cases this script
imports ``/root/reference/model`` (with the empty ``dataset`` package shell of
SURVEY.md section 8c so ``dataset.constant`` loads without skorch/torchtext), runs
forward + the skorch-equivalent training step, and stores inputs, initial weights,
log-probs, loss, gradients and the weights after three steps as ``.npz`` fixtures.
The reference cannot travel to the GPU box; these files do.
"""
import collections
import os
import sys
import types

import numpy as np
import torch

REF = os.environ.get("SLNLP_REFERENCE", "/root/reference")
OUT = os.path.dirname(os.path.abspath(__file__))


def _import_reference():
    sys.path.insert(0, REF)
    pkg = types.ModuleType("dataset")
    pkg.__path__ = [os.path.join(REF, "dataset")]
    sys.modules["dataset"] = pkg
    import model  # noqa: the reference package
    return model


class Vocab:
    """torchtext-0.6 duck type: specials first, stoi is a defaultdict -> unk."""

    def __init__(self, n):
        self.itos = ["<unk>", "<pad>"] + [f"tok{i}" for i in range(n - 2)]
        self.stoi = collections.defaultdict(lambda: 0, {w: i for i, w in enumerate(self.itos)})

    def __len__(self):
        return len(self.itos)


def make_inputs(seed, B, T, v_src, v_tgt, ragged=True):
    g = torch.Generator().manual_seed(seed)
    X = torch.randint(2, v_src, (B, T), generator=g)
    if ragged:
        lengths = torch.randint(1, T + 1, (B,), generator=g)
        lengths[0] = T
        lengths[-1] = 1
    else:
        lengths = torch.full((B,), T, dtype=torch.long)
    for b in range(B):
        X[b, lengths[b]:] = 1
    y = torch.randint(2, v_tgt, (B,), generator=g)
    return X, lengths, y


CASES = [
    # name, class, ctor kwargs, B, T, v_src, v_tgt, ragged, lr
    ("lstm_small", "EncoderDecoderLSTMAttn", dict(embedding_size=16, hidden_size=16, num_layers=2), 6, 9, 40, 12, True, 0.1),
    ("lstm_l3_odd", "EncoderDecoderLSTMAttn", dict(embedding_size=24, hidden_size=20, num_layers=3), 5, 7, 33, 9, True, 0.01),
    ("gru_small", "EncoderDecoderGRUAttn", dict(embedding_size=16, hidden_size=16, num_layers=2), 6, 9, 40, 12, True, 0.1),
    ("gru_l1_full", "EncoderDecoderGRUAttn", dict(embedding_size=12, hidden_size=24, num_layers=1), 4, 6, 21, 7, False, 0.1),
    ("transformer_small", "Transformer", dict(embedding_size=16, hidden_size=32, num_layers=2, num_heads=4), 6, 9, 40, 12, True, 0.1),
    ("transformer_h2", "Transformer", dict(embedding_size=24, hidden_size=20, num_layers=1, num_heads=2), 5, 7, 33, 9, False, 0.01),
]


def run_case(model, name, cls, kw, B, T, v_src, v_tgt, ragged, lr):
    torch.manual_seed(1)
    dev = torch.device("cpu")
    m = getattr(model, cls)(src_vocab=Vocab(v_src), tgt_vocab=Vocab(v_tgt), batch_first=True,
                            dropout=0.0, device=dev, **kw).to(dev)
    X, lengths, y = make_inputs(7, B, T, v_src, v_tgt, ragged)
    out = {"X": X.numpy(), "lengths": lengths.numpy(), "y": y.numpy(), "lr": np.float32(lr)}
    for k, v in m.state_dict().items():
        if not k.endswith(".pe"):
            out["w0/" + k] = v.detach().numpy().copy()
    # eval-mode forward (inference path)
    m.eval()
    with torch.no_grad():
        out["logp_eval"] = m(X=X, y=y, lengths=lengths).numpy().copy()
    # three skorch-equivalent train steps
    m.train()
    opt = torch.optim.SGD(m.parameters(), lr=lr, momentum=0.9, nesterov=False)
    crit = torch.nn.CrossEntropyLoss(ignore_index=1)
    for step in range(3):
        opt.zero_grad()
        logp = m(X=X, y=y, lengths=lengths)
        loss = crit(logp, y)
        loss.backward()
        if step == 0:
            out["logp_train"] = logp.detach().numpy().copy()
            for k, p in m.named_parameters():
                if p.grad is not None:
                    out["g0/" + k] = p.grad.detach().numpy().copy()
        out[f"loss{step}"] = np.float32(loss.item())
        out[f"gnorm{step}"] = np.float32(
            torch.nn.utils.clip_grad_norm_(m.parameters(), max_norm=0.5, norm_type=2).item())
        opt.step()
    for k, v in m.state_dict().items():
        if not k.endswith(".pe"):
            out["w3/" + k] = v.detach().numpy().copy()
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)
    print(name, "loss", [float(out[f"loss{i}"]) for i in range(3)], "gnorm0", float(out["gnorm0"]))


if __name__ == "__main__":
    ref_model = _import_reference()
    for case in CASES:
        run_case(ref_model, *case)
