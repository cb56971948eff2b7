"""Per-kernel totals of ONE captured train step at the workload's own batch (CUDA-graph replay under
torch.profiler / CUPTI), sorted by time:   python profiles/kernel_table_step.py [cfg4|cfg1|...] [bf16|fp32] [batch]"""
import os, sys, collections
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "sign-language-nlp_b200"))
import torch
import bench
from phono_synth import synthetic_dataset
from slnlp_b200.rnn import FusedTrainStep
wl = sys.argv[1] if len(sys.argv) > 1 else "cfg4"
prec = sys.argv[2] if len(sys.argv) > 2 else "bf16"
w = bench.WORKLOADS[wl]
B = int(sys.argv[3]) if len(sys.argv) > 3 else w["B"]
dev = torch.device("cuda")
data = synthetic_dataset(n_seq=B, T=w["T"], v_src=w["Vs"], v_tgt=w["Vt"], seed=1)
m = bench.make_module(w, data, dev, prec)
ts = FusedTrainStep(m, B, w["T"], lr=0.01)
X, y, l = data["X"][:B].to(dev), data["y"][:B].to(dev), data["lengths"][:B].to(dev)
for _ in range(3):
    ts.step(X, y, l)
torch.cuda.synchronize()
with torch.profiler.profile(activities=[torch.profiler.ProfilerActivity.CUDA]) as prof:
    ts.run()
    torch.cuda.synchronize()
evs = [e for e in prof.events() if e.device_time_total > 0 and e.device_type == torch.autograd.DeviceType.CUDA]
evs.sort(key=lambda e: e.time_range.start)
span = evs[-1].time_range.end - evs[0].time_range.start
tot, cnt = collections.Counter(), collections.Counter()
for e in evs:
    name = e.name.split("(")[0][-70:]
    tot[name] += e.time_range.end - e.time_range.start
    cnt[name] += 1
print(f"{wl} {prec} batch {B}: one replay, {len(evs)} kernels, span {span / 1e3:.2f} ms, summed kernel time {sum(tot.values()) / 1e3:.2f} ms")
print(f"{'total ms':>9} {'count':>6} {'avg us':>9}  kernel")
for name, t in tot.most_common():
    print(f"{t / 1e3:9.3f} {cnt[name]:6d} {t / cnt[name]:9.2f}  {name}")
