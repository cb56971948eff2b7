"""Shared helpers for the parity tests (golden loading, tolerances)."""
import os

import numpy as np
import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

# name -> (kind, ctor kwargs)   (mirrors tests/golden/make_golden.py:CASES)
GOLDEN_CASES = {
    "lstm_small": ("lstm", dict(embedding_size=16, hidden_size=16, num_layers=2)),
    "lstm_l3_odd": ("lstm", dict(embedding_size=24, hidden_size=20, num_layers=3)),
    "gru_small": ("gru", dict(embedding_size=16, hidden_size=16, num_layers=2)),
    "gru_l1_full": ("gru", dict(embedding_size=12, hidden_size=24, num_layers=1)),
    "transformer_small": ("transformer", dict(embedding_size=16, hidden_size=32, num_layers=2, num_heads=4)),
    "transformer_h2": ("transformer", dict(embedding_size=24, hidden_size=20, num_layers=1, num_heads=2)),
}
RNN_CASES = [k for k, v in GOLDEN_CASES.items() if v[0] != "transformer"]
TRANSFORMER_CASES = [k for k, v in GOLDEN_CASES.items() if v[0] == "transformer"]

# north_star tolerances
FP32_RTOL = 1e-5     # logits / loss, fp32 path
BF16_RTOL = 2e-2     # bf16 tensor-core path


def load_golden(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    g = {"X": torch.from_numpy(z["X"]), "lengths": torch.from_numpy(z["lengths"]),
         "y": torch.from_numpy(z["y"]), "lr": float(z["lr"]),
         "logp_eval": torch.from_numpy(z["logp_eval"]), "logp_train": torch.from_numpy(z["logp_train"]),
         "loss": [float(z[f"loss{i}"]) for i in range(3)],
         "gnorm": [float(z[f"gnorm{i}"]) for i in range(3)],
         "w0": {}, "g0": {}, "w3": {}}
    for k in z.files:
        for pre in ("w0/", "g0/", "w3/"):
            if k.startswith(pre):
                g[pre[:-1]][k[len(pre):]] = torch.from_numpy(z[k])
    return g


def rel_err(a, b):
    """max |a-b| / max |b|  (the 'relative' of north_star: scale of the tensor)."""
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    denom = b.abs().max().clamp_min(1e-30)
    return float((a - b).abs().max() / denom)


def grad_rel_err(a, b, global_scale):
    """Gradient comparison: per-tensor max-abs error over max(|ref|max, 1e-3*global
    gradient scale).  Some gradients (e.g. the attention query layer, whose input
    shifts every score alike) are pure cancellation noise ~1e-8; they are judged
    against the scale of the whole gradient, as the clipped update sees them."""
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    denom = max(float(b.abs().max()), 1e-3 * float(global_scale), 1e-30)
    return float((a - b).abs().max()) / denom
