"""Error margins of tests/test_gpu_rnn_parity.py::test_decoder_fused_kernels_match_the_unfused_chain: the measured
log-prob and gradient differences between the fused decoder kernels and the chain they replace, next to the test's bars."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "sign-language-nlp_b200")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import model as dropin
from helpers import grad_rel_err, rel_err
from slnlp_b200.vocab import Vocab
from test_gpu_rnn_parity import _synthetic

CASES = [("lstm", 128, 128, 2, 50, 64, 0.1, "fp32"), ("lstm", 128, 128, 2, 50, 64, 0.1, "bf16"), ("gru", 64, 96, 3, 37, 20, 0.2, "fp32"),
         ("lstm", 256, 256, 2, 50, 33, 0.1, "bf16"), ("gru", 48, 512, 1, 9, 12, 0.0, "fp32"), ("lstm", 24, 20, 3, 6, 7, 0.3, "fp32")]
for kind, E, H, L, B, T, dropout, precision in CASES:
    cls = dropin.EncoderDecoderLSTMAttn if kind == "lstm" else dropin.EncoderDecoderGRUAttn
    X, lengths, y = [t.cuda() for t in _synthetic(B, T, 300, 40, ragged=True)]
    out = {}
    for fused in ("1", "2", "0"):
        os.environ["SLNLP_DEC_HEAD"] = "0" if fused == "0" else "1"
        os.environ["SLNLP_DEC_CELL_BWD"] = "0" if fused == "0" else "2"
        os.environ["SLNLP_DEC_HEAD_FUSE"] = "1" if fused == "2" else "0"
        torch.manual_seed(5)
        m = cls(src_vocab=Vocab(size=300), tgt_vocab=Vocab(size=40), batch_first=True, embedding_size=E, hidden_size=H, num_layers=L,
                dropout=dropout, device=torch.device("cuda"), seed=11, precision=precision).to(torch.device("cuda"))
        m.train()
        logp = m(X=X, y=y, lengths=lengths)
        torch.nn.functional.cross_entropy(logp, y, ignore_index=1).backward()
        out[fused] = (logp.detach().clone(), {k: p.grad.clone() for k, p in m.named_parameters() if p.grad is not None})
    scale = max(float(v.abs().max()) for v in out["0"][1].values())
    bars = (2e-5, 5e-5) if precision == "fp32" else (5e-3, 2e-2)
    for fused in ("1", "2"):
        e_logp = rel_err(out[fused][0], out["0"][0])
        worst = max((grad_rel_err(out[fused][1][k], ref, scale), k) for k, ref in out["0"][1].items())
        print(f"{kind} E{E} H{H} L{L} B{B} T{T} p{dropout} {precision} mode {fused}: logp {e_logp:.2e} (bar {bars[0]:.0e}), "
              f"worst gradient {worst[0]:.2e} (bar {bars[1]:.0e}) {worst[1]}")
