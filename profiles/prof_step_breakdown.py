"""Warm per-kernel device-time breakdown of one training step (CUPTI via torch.profiler).

    python profiles/prof_step_breakdown.py [fp32|bf16] [cfg1|cfg2]

ncu launch lists are cold-cache and serialised; this gives kernel durations as they are in
the running step (no CUDA graph here, so gaps between kernels are Python, not the product)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "sign-language-nlp_b200"))
import torch
from torch.profiler import ProfilerActivity, profile

import bench
import model as dropin
from slnlp_b200.data import synthetic_dataset
from slnlp_b200.rnn import FusedTrainStep

prec = sys.argv[1] if len(sys.argv) > 1 else "fp32"
w = bench.WORKLOADS[sys.argv[2] if len(sys.argv) > 2 else "cfg1"]
dev = torch.device("cuda")
data = synthetic_dataset(n_seq=500, T=w["T"], v_src=w["Vs"], v_tgt=w["Vt"])
cls = dropin.EncoderDecoderLSTMAttn if w["kind"] == "lstm" else dropin.EncoderDecoderGRUAttn
torch.manual_seed(1)
m = cls(src_vocab=data["src_vocab"], tgt_vocab=data["tgt_vocab"], batch_first=True, embedding_size=w["E"],
        hidden_size=w["H"], num_layers=w["L"], dropout=w["p"], device=dev, precision=prec).to(dev).train()
B = w["B"]
ts = FusedTrainStep(m, B, w["T"], lr=0.01, use_graph=False)
X, y, l = data["X"][:B].to(dev), data["y"][:B].to(dev), data["lengths"][:B].to(dev)
for _ in range(5):
    ts.step(X, y, l)
torch.cuda.synchronize()
N = 5
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(N):
        ts.step(X, y, l)
    torch.cuda.synchronize()
rows = [(e.key, e.count, e.device_time_total) for e in prof.key_averages() if e.device_time_total > 0]
tot = sum(r[2] for r in rows)
print(f"{prec} {w['name']}: kernel time per step {tot / N:.1f} us over {sum(r[1] for r in rows) // N} launches")
print(f"{'us/step':>9} {'n/step':>6} {'avg us':>8} {'share':>6}  kernel")
for k, n, t in sorted(rows, key=lambda r: -r[2]):
    print(f"{t / N:9.1f} {n / N:6.1f} {t / n:8.2f} {100 * t / tot:5.1f}%  {k[:90]}")
