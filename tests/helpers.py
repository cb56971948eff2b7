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
import os as _os
# $SLNLP_TEST_TOL_SCALE < 1 tightens the fp32 tolerances (margin checks when a kernel changes; the default run uses 1)
FP32_RTOL = 1e-5 * float(_os.environ.get("SLNLP_TEST_TOL_SCALE", "1"))     # logits / loss, fp32 path
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


# ---------------------------------------------------------------- BASELINE-size fixtures (make_golden_baseline.py)
BASELINE_CASES = {
    "cfg1": ("lstm", dict(embedding_size=128, hidden_size=128, num_layers=2)),
    "cfg2": ("gru", dict(embedding_size=512, hidden_size=256, num_layers=4)),
    "cfg3": ("transformer", dict(embedding_size=512, hidden_size=256, num_layers=4, num_heads=8)),
    "cfg4s": ("lstm", dict(embedding_size=1024, hidden_size=512, num_layers=6)),
}
BASELINE_VS, BASELINE_VT = 4098, 1026


def sample_index(name, numel, n_sample=64):
    """The element positions make_golden_baseline.py sampled for tensor ``name``."""
    seed = (sum(ord(c) * (i + 1) for i, c in enumerate(name)) * 2654435761 + numel) % (2 ** 31)
    g = torch.Generator().manual_seed(seed)
    return torch.randint(0, numel, (min(n_sample, numel),), generator=g)


def checksum(t):
    t = t.detach().double().reshape(-1).cpu()
    return np.array([float(t.sum()), float(t.abs().sum()), float(t.abs().max())])


def load_baseline_golden(name):
    z = np.load(os.path.join(GOLDEN, f"baseline_{name}.npz"))
    g = {"X": torch.from_numpy(z["X"]), "lengths": torch.from_numpy(z["lengths"]), "y": torch.from_numpy(z["y"]),
         "lr": float(z["lr"]), "logp_eval": torch.from_numpy(z["logp_eval"]), "names": [str(n) for n in z["names"]],
         "loss": [float(z["loss0"]), float(z["loss1"])], "gnorm": [float(z["gnorm0"]), float(z["gnorm1"])],
         "w0sum": {}, "g0sum": {}, "g0smp": {}, "w2sum": {}, "w2smp": {}}
    for k in z.files:
        pre, _, rest = k.partition("/")
        if rest and pre in g:
            g[pre][rest] = z[k]
    return g


def build_baseline_dropin(name, device, precision="fp32"):
    """The drop-in module of a BASELINE case with the reference's seed-1 initial weights."""
    import model as dropin
    from slnlp_b200.vocab import Vocab
    kind, kw = BASELINE_CASES[name]
    cls = {"lstm": dropin.EncoderDecoderLSTMAttn, "gru": dropin.EncoderDecoderGRUAttn, "transformer": dropin.Transformer}[kind]
    torch.manual_seed(1)
    return cls(src_vocab=Vocab(size=BASELINE_VS), tgt_vocab=Vocab(size=BASELINE_VT), batch_first=True, dropout=0.0,
               device=device, precision=precision, **kw)


def check_baseline_state(sd, g, which, tol, floor=1e-3):
    """Sampled elements and checksums of ``sd`` (weights or gradients by name) against the fixture:
    every sampled element within ``tol`` of the tensor's scale (max |.|, floored)."""
    worst = 0.0
    for k, smp in g[which + "smp"].items():
        t = sd[k].detach().reshape(-1).cpu()
        scale = max(float(g[which + "sum"][k][2]), floor)
        err = float((t[sample_index(k, t.numel())].double() - torch.from_numpy(smp).double()).abs().max()) / scale
        assert err < tol, (k, err)
        # abs-sum checksum: catches an error anywhere in the tensor, not only at the sampled positions
        want = float(g[which + "sum"][k][1])
        got = float(t.double().abs().sum())
        assert abs(got - want) <= max(50 * tol, 1e-3) * max(want, floor), (k, got, want)
        worst = max(worst, err)
    return worst
