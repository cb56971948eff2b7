"""Warm per-kernel device times of one fused train step (CUDA-graph replays under torch.profiler /
CUPTI - unlike an ncu launch list these are L2-warm and overlapped as in production):
    python profiles/torch_prof_step.py [cfg1|cfg2|cfg3] [bf16|fp32]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "sign-language-nlp_b200"))
import torch
import bench
import model as dropin
from slnlp_b200.data import synthetic_dataset
from slnlp_b200.rnn import FusedTrainStep
wl = sys.argv[1] if len(sys.argv) > 1 else "cfg1"
prec = sys.argv[2] if len(sys.argv) > 2 else "bf16"
w = bench.WORKLOADS[wl]
dev = torch.device("cuda")
data = synthetic_dataset(n_seq=500, T=w["T"], v_src=w["Vs"], v_tgt=w["Vt"], seed=1)
torch.manual_seed(1)
cls = {"lstm": dropin.EncoderDecoderLSTMAttn, "gru": dropin.EncoderDecoderGRUAttn, "transformer": dropin.Transformer}[w["kind"]]
extra = {"num_heads": w["heads"]} if w["kind"] == "transformer" else {}
m = cls(src_vocab=data["src_vocab"], tgt_vocab=data["tgt_vocab"], batch_first=True, embedding_size=w["E"], hidden_size=w["H"],
        num_layers=w["L"], dropout=w["p"], device=dev, precision=prec, **extra).to(dev).train()
ts = FusedTrainStep(m, w["B"], w["T"], lr=0.01)
X, y, l = data["X"][:50].to(dev), data["y"][:50].to(dev), data["lengths"][:50].to(dev)
for _ in range(5):
    ts.step(X, y, l)
torch.cuda.synchronize()
N = 10
with torch.profiler.profile(activities=[torch.profiler.ProfilerActivity.CUDA]) as prof:
    for _ in range(N):
        ts.step(X, y, l)
    torch.cuda.synchronize()
rows = [(e.key, e.count, e.device_time_total) for e in prof.key_averages() if e.device_time_total > 0]
tot = sum(r[2] for r in rows)
print(f"{wl} {prec}: sum of kernel time per step {tot / N:.1f} us")
print(f"{'us/step':>9} {'n/step':>7} {'avg us':>8} {'share':>6}  kernel")
for k, n, t in sorted(rows, key=lambda r: -r[2])[:28]:
    print(f"{t / N:9.1f} {n / N:7.1f} {t / n:8.2f} {100 * t / tot:5.1f}%  {k[:90]}")
if os.environ.get("SLNLP_PROF_EVENTS"):
    # individual launches of one kernel family, in launch order, for the last step
    pat = os.environ["SLNLP_PROF_EVENTS"]
    evs = [e for e in prof.events() if pat in e.name and e.device_time_total > 0]
    per_step = len(evs) // N
    print(f"last step, {per_step} launches matching {pat!r}:")
    for e in evs[-per_step:]:
        print(f"  {e.device_time_total:8.2f} us  {e.name[:100]}")
