"""Timeline of ONE captured train step (CUDA-graph replay under torch.profiler / CUPTI): every kernel with its
stream, start offset and duration, in start order, plus the idle gaps of the critical (main) stream.
    python profiles/timeline_step.py [cfg1|cfg2|cfg3] [bf16|fp32]"""
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
m = bench.make_module(w, data, dev, prec)
ts = FusedTrainStep(m, w["B"], w["T"], lr=0.01)
X, y, l = data["X"][:50].to(dev), data["y"][:50].to(dev), data["lengths"][:50].to(dev)
for _ in range(5):
    ts.step(X, y, l)
torch.cuda.synchronize()
with torch.profiler.profile(activities=[torch.profiler.ProfilerActivity.CUDA]) as prof:
    for _ in range(3):
        ts.run()
    torch.cuda.synchronize()
evs = [e for e in prof.events() if e.device_time_total > 0 and e.device_type == torch.autograd.DeviceType.CUDA]
evs.sort(key=lambda e: e.time_range.start)
# split into the three replays by the largest gaps
starts = [e.time_range.start for e in evs]
n = len(evs) // 3
step = evs[2 * n:]
t0 = step[0].time_range.start
streams = {}
print(f"{wl} {prec}: last replay, {len(step)} kernels, span {(step[-1].time_range.end - t0):.1f} us")
print(f"{'start':>8} {'dur':>7} {'stream':>6}  kernel")
main_end, gaps = None, []
for e in step:
    sid = getattr(e, "device_resource_id", None)
    if sid is None:
        sid = getattr(e, "stream", -1)
    streams.setdefault(sid, 0.0)
    streams[sid] += e.time_range.end - e.time_range.start
    print(f"{e.time_range.start - t0:8.1f} {e.time_range.end - e.time_range.start:7.2f} {sid!s:>6}  {e.name[:80]}")
print("busy time per stream:", {k: round(v, 1) for k, v in streams.items()})
