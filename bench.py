#!/usr/bin/env python
"""Headline benchmark: training sequences/s of one fit of the reference's
EncoderDecoderLSTMAttn on one B200, next to the reference's CPU path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload cfg1|cfg2|cfg3|cfg4|cfg5] [--precision bf16|fp32]
    python bench.py --impl reference ...        # the reference's own modules on the host cores
    torchrun ... bench.py --gpus N ...          # N > 1: one rank per GPU

cfg1 (default) = BASELINE.json configs[0], the configuration the metric is quoted on; cfg2 = GRU
512/256/4; cfg3 = Transformer 512/256/4/h8; cfg4 = LSTM 1024/512/6 at batch 4096; cfg5 = the LSTM
hyper-parameter grid farmed over the GPUs (fits/hour).

A "step" is one skorch-equivalent training step (forward, CrossEntropyLoss on the log-probs,
backward, global-norm clip 0.5, SGD momentum 0.9) on one batch of synthetic 6-field phonology
sequences.  At N > 1 the headline `value` is N independent fits (the
reference's only parallelism is farming grid-search fits, one per GPU, no collective: weak scaling).

The default (cfg1) line also carries the other BASELINE.json configurations as sub-records, so that
the driver's `bench.py --gpus N` runs put them on record:
    N = 1:  fp32_path (the 1e-5 / identical-argmax path), infer (predict seq/s), dp (cfg4: batch 4096
            on one GPU, tensor roofline), grid (cfg5 slice, fits/hour through GridSearchFarm)
    N > 1:  dp (cfg4, global batch 4096 split over the ranks, overlapped NCCL all-reduce, strong
            scaling, parity against the single-rank step), grid (cfg5 slice over N GPUs, no collective)
`--legs a,b,...` selects sub-records (default all; `--legs none` = headline only).

Prints ONE JSON line (rank 0).  See DESIGN.md "Measurement" for every field.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, "sign-language-nlp_b200")
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)

WORKLOADS = {
    # BASELINE.json configs[0]: the configuration the metric is quoted on ("LSTM-attn, 1 B200")
    "cfg1": dict(kind="lstm", E=128, H=128, L=2, p=0.1, B=50, T=64, Vs=4098, Vt=1026,
                 name="EncoderDecoderLSTMAttn emb128 hidden128 layers2 dropout0.1 batch50 len64"),
    "cfg2": dict(kind="gru", E=512, H=256, L=4, p=0.1, B=50, T=64, Vs=4098, Vt=1026,
                 name="EncoderDecoderGRUAttn emb512 hidden256 layers4 batch50 len64"),
    # config-transformer.yaml grid point: model.Transformer E512, dim_feedforward 256, 4+4 layers, 8 heads
    "cfg3": dict(kind="transformer", E=512, H=256, L=4, p=0.1, heads=8, B=50, T=64, Vs=4098, Vt=1026,
                 name="Transformer emb512 ffn256 layers4 heads8 dropout0.1 batch50 len64"),
    "cfg4": dict(kind="lstm", E=1024, H=512, L=6, p=0.5, B=4096, T=64, Vs=4098, Vt=1026,
                 name="EncoderDecoderLSTMAttn emb1024 hidden512 layers6 dropout0.5 batch4096 len64"),
    # BASELINE.json configs[4]: the full config-enc-dec-lstm-attn grid, farmed over the GPUs (no collective)
    "cfg5": dict(kind="grid", B=50, T=64, Vs=4098, Vt=52, E=0, H=0, L=0, p=0.0,
                 name="config-enc-dec-lstm-attn grid: 3 lr x 3 emb x 3 hidden x 3 layers x 2 dropout x 5-fold CV = 810 fits"),
}
METRIC, UNIT = "train_seq_per_s", "sequences/s"
L2_NOTE = "B200 arm: 256 MiB zero-fill between timed steps (outside the timed events)"
ALL_LEGS = ("fp32_path", "infer", "dp", "grid")


def common_config(w, args, world):
    """The `config` object - a function of the workload and the command line only, so the B200 arm and
    the reference arm print the SAME object."""
    return {"workload": w["name"], "batch": w["B"], "seq_len": w["T"], "v_src": w["Vs"], "v_tgt": w["Vt"],
            "optimizer": "SGD momentum 0.9, global-norm clip 0.5, lr 0.01", "criterion": "CrossEntropyLoss(ignore_index=1) on log-probs",
            "parallelism": ("dp%d: one global batch split over the ranks, NCCL all-reduce" % world) if args.dp else
                           ("%d independent fit(s), one per GPU (grid-search farm, no collective)" % world),
            "l2": "none (--no-flush)" if args.no_flush else L2_NOTE}


def train_flops_per_seq(w):
    """3 x forward GEMM FLOPs of the live graph."""
    E, H, L, T, V = w["E"], w["H"], w["L"], w["T"], w["Vt"]
    if w["kind"] == "transformer":
        F = H
        enc = L * (2 * T * (3 * E * E + E * E + 2 * E * F) + 4 * T * T * E)          # qkv, out, ffn, QK^T + PV
        dec = L * (2 * (3 * E * E + E * E) + 2 * (E * E + E * E) + 2 * T * 2 * E * E + 4 * T * E + 2 * 2 * E * F)
        return 3 * (enc + dec + 2 * E * V)
    G = 4 if w["kind"] == "lstm" else 3
    enc = sum(2 * T * (2 * G * H * (E if l == 0 else 2 * H) + 2 * G * H * H) for l in range(L))
    key, bridge = T * 2 * 2 * H * H, L * 2 * 2 * H * H
    att = 2 * H * H + 2 * T * H + 4 * T * H
    dec = sum(2 * G * H * ((E + 2 * H) if l == 0 else H) + 2 * G * H * H for l in range(L))
    return 3 * (enc + key + bridge + att + dec + 2 * H * V)


def load_peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except (OSError, ValueError):
        return {}


class ClockSampler:
    """nvidia-smi clocks + throttle reasons sampled DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def __exit__(self, *a):
        if self.proc:
            time.sleep(0.15)
            self.proc.terminate()
            self.th.join(timeout=2)

    def summary(self):
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
                for n, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except (ValueError, IndexError):
                pass
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm)}


# ======================================================================================
# reference arm: the reference's own CPU implementation on the box's host cores
# ======================================================================================
def cpu_model_name():
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("model name"):
                return line.split(":", 1)[1].strip()
    except OSError:
        pass
    return "unknown"


def build_reference_module(w, data):
    """The UNMODIFIED reference modules staged under baseline/_ref (oracle/stage_reference.py) when
    present - kind "reference" - else the torch.nn port of them (oracle/port.py) - kind "port"."""
    import torch
    from oracle import stage_reference
    torch.manual_seed(1)
    if stage_reference.available():
        ref = stage_reference.load()
        cls = {"lstm": "EncoderDecoderLSTMAttn", "gru": "EncoderDecoderGRUAttn", "transformer": "Transformer"}[w["kind"]]
        extra = {"num_heads": w["heads"]} if w["kind"] == "transformer" else {}
        dev = torch.device("cpu")
        m = getattr(ref, cls)(src_vocab=data["src_vocab"], tgt_vocab=data["tgt_vocab"], batch_first=True,
                              embedding_size=w["E"], hidden_size=w["H"], num_layers=w["L"], dropout=w["p"], device=dev,
                              **extra).to(dev)
        return m, "reference", "the reference's own model package (baseline/_ref, unmodified) driven by the skorch train step"
    from oracle import port
    m = port.build_port(w["kind"], w["Vs"], w["Vt"], w["E"], w["H"], w["L"], dropout=w["p"], num_heads=w.get("heads"))
    return m, "port", "oracle/port.py = the reference modules on stock torch.nn (baseline/_ref not staged)"


def time_cpu_reference(w, data, steps, warmup, batch=None, infer_batches=0):
    """Median seconds per skorch-equivalent training step of the reference's CPU path."""
    import torch
    from oracle import port
    cores = os.cpu_count() or 1
    B = batch or min(w["B"], 50)
    ref, kind, what = build_reference_module(w, data)
    opt = torch.optim.SGD(ref.parameters(), lr=0.01, momentum=0.9, nesterov=False)
    X, y, lengths = data["X"], data["y"], data["lengths"]
    nb = X.shape[0] // B
    # "all the host threads it can use": at batch 50 the reference's thousands of small ATen ops do not
    # always scale with threads (16 threads measured SLOWER than 1 on the GPU box), so the arm is given its
    # best intra-op thread count - a short calibration over {all, half, 4, 1} - and reports which it took
    cand = sorted({cores, max(1, cores // 2), min(4, cores), 1}, reverse=True)
    if len(cand) > 1 and w["H"] <= 256:
        best = None
        for n in cand:
            torch.set_num_threads(n)
            port.reference_train_step(ref, opt, X[:B], y[:B], lengths[:B])
            t0 = time.perf_counter()
            for i in range(2):
                port.reference_train_step(ref, opt, X[:B], y[:B], lengths[:B])
            dt = time.perf_counter() - t0
            if best is None or dt < best[0]:
                best = (dt, n)
        cores = best[1]
    torch.set_num_threads(cores)
    times = []
    for i in range(warmup + steps):
        j = (i % nb) * B
        t0 = time.perf_counter()
        port.reference_train_step(ref, opt, X[j:j + B], y[j:j + B], lengths[j:j + B])
        if i >= warmup:
            times.append(time.perf_counter() - t0)
    sec = statistics.median(times)
    infer = None
    if infer_batches:
        ref.eval()
        tt = []
        with torch.no_grad():
            for i in range(2 + infer_batches):
                j = (i % nb) * B
                t0 = time.perf_counter()
                ref(X=X[j:j + B], y=y[j:j + B], lengths=lengths[j:j + B]).argmax(1)
                if i >= 2:
                    tt.append(time.perf_counter() - t0)
        infer = B / statistics.median(tt)
    return dict(value=B / sec, sec=sec, cores=cores, batch=B, kind=kind, what=what, infer=infer)


def run_reference(args, w, world):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch
    from phono_synth import synthetic_dataset      # torch-only: the reference arm maps no product .so
    data = synthetic_dataset(n_seq=max(500, 50 * 10), T=w["T"], v_src=w["Vs"], v_tgt=w["Vt"], seed=1)
    steps, warmup = max(1, args.steps), max(1, args.warmup)
    if w["B"] > 50 or w["H"] > 256:  # cfg4 on CPU: batch 50 sample (7.5 s/step), reported per sequence
        steps, warmup = min(steps, 3), 1
    r = time_cpu_reference(w, data, steps, warmup, infer_batches=(8 if w["H"] <= 256 else 0))
    val = r["value"]
    line = {"metric": METRIC, "value": val, "unit": UNIT, "impl": "reference", "n_gpus": args.gpus,
            "steps": steps, "warmup": warmup, "ms_per_step": r["sec"] * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": common_config(w, args, world),
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": r["cores"], "kind": r["kind"], "cpu": cpu_model_name(),
                             "sample": f"{steps} training steps of batch {r['batch']} (median), torch {torch.__version__} CPU, "
                                       f"{r['cores']} intra-op threads (the fastest of all / half / 4 / 1 on this box, {os.cpu_count()} host cores); {r['what']}"},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    if r["infer"] is not None:
        line["infer"] = {"value": r["infer"], "unit": UNIT, "sample": f"median of 8 eval forwards + argmax of batch {r['batch']} (no_grad)"}
    print(json.dumps(line), flush=True)


def cpu_baseline_subprocess(args, workload):
    """The reference arm, run as its own process (its `model` package and the drop-in's share a name), on
    the same box in the same run.  Returns its parsed JSON line."""
    small = workload in ("cfg1",)
    cmd = [sys.executable, os.path.abspath(__file__), "--impl", "reference", "--workload", workload,
           "--steps", "12" if small else "3", "--warmup", "2"]
    env = {k: v for k, v in os.environ.items() if k not in ("RANK", "WORLD_SIZE", "LOCAL_RANK")}
    env["CUDA_VISIBLE_DEVICES"] = ""
    out = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=900)
    for ln in reversed(out.stdout.strip().splitlines()):
        if ln.startswith("{"):
            return json.loads(ln)
    raise RuntimeError("reference arm printed no line: " + out.stderr[-400:])


# ======================================================================================
# B200 arm
# ======================================================================================
class Env:
    def __init__(self):
        import torch
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        torch.cuda.set_device(self.local)
        self.dev = torch.device("cuda", self.local)

    def barrier(self):
        import torch
        import torch.distributed as dist
        if self.world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(self, x):
        import torch
        import torch.distributed as dist
        t = torch.tensor([x], device=self.dev, dtype=torch.float64)
        if self.world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t)


def make_module(w, data, dev, precision, dropout=None):
    import torch
    import model as dropin
    torch.manual_seed(1)
    cls = {"lstm": dropin.EncoderDecoderLSTMAttn, "gru": dropin.EncoderDecoderGRUAttn, "transformer": dropin.Transformer}[w["kind"]]
    extra = {"num_heads": w["heads"]} if w["kind"] == "transformer" else {}
    return cls(src_vocab=data["src_vocab"], tgt_vocab=data["tgt_vocab"], batch_first=True, embedding_size=w["E"],
               hidden_size=w["H"], num_layers=w["L"], dropout=w["p"] if dropout is None else dropout, device=dev,
               precision=precision, **extra).to(dev).train()


def measure_train(env, w, data, m, B, W, K, flush, grad_sync=None, split=False, e2e=True, clocks=False):
    """W warm-up + K timed training steps of module `m` at per-rank batch B.
    Device-resident leg: dataset in HBM, CUDA events per step (L2 flushed between steps, outside the
    events), max over ranks.  e2e leg: pinned host batches through FusedTrainStep.step + loss read-back."""
    import torch
    from slnlp_b200 import _lib
    from slnlp_b200.rnn import FusedTrainStep
    dev, world, rank = env.dev, env.world, env.rank
    ts = FusedTrainStep(m, B, w["T"], lr=0.01, momentum=0.9, max_norm=0.5, grad_sync=grad_sync)
    Xd, yd, ld = data["X"].to(dev), data["y"].to(dev), data["lengths"].to(dev)
    if split and world > 1:  # each rank takes its slice of every global batch
        Xd, yd, ld = Xd[rank::world].contiguous(), yd[rank::world].contiguous(), ld[rank::world].contiguous()
    nb = Xd.shape[0] // B

    def batch(i):
        j = (i % nb) * B
        return Xd[j:j + B], yd[j:j + B], ld[j:j + B]

    for i in range(W):
        ts.step(*batch(i))
    env.barrier()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
    env.barrier()
    sampler = ClockSampler(env.local) if clocks else None
    if sampler:
        sampler.__enter__()
    t_wall0 = time.perf_counter()
    for i in range(K):
        if flush is not None:
            flush.zero_()                       # evict L2 between timed steps (not timed)
        ts.load_batch(*batch(W + i))
        ev[i][0].record()
        ts.run()
        ev[i][1].record()
    env.barrier()
    t_wall = time.perf_counter() - t_wall0
    if sampler:
        sampler.__exit__()
    dev_s = env.max_over_ranks(sum(a.elapsed_time(b) for a, b in ev) / 1e3)
    out = {"value": K * B * world / dev_s, "ms_per_step": dev_s / K * 1e3, "final_loss": float(ts.ws.loss[0]),
           "wall_ms_per_step_incl_flush": t_wall / K * 1e3, "cuda_graph": bool(ts.use_graph)}
    if sampler:
        out["clocks"] = sampler.summary()
    # kernels per step: count one un-captured step through the C ABI
    env.barrier()
    c0 = _lib.lib.slnlp_launch_count()
    ts._step()
    torch.cuda.synchronize()
    out["launches_per_step"] = int(_lib.lib.slnlp_launch_count() - c0)
    if e2e:
        src = (data["X"], data["y"], data["lengths"])
        if split and world > 1:
            src = tuple(t[rank::world].contiguous() for t in src)
        Xh, yh, lh = (t.pin_memory() for t in src)
        for i in range(3):
            j = (i % nb) * B
            float(ts.step(Xh[j:j + B], yh[j:j + B], lh[j:j + B])[0])
        env.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        host_loss = torch.empty(2, pin_memory=True)
        e0.record()
        for i in range(K):
            j = ((W + i) % nb) * B
            loss = ts.step(Xh[j:j + B], yh[j:j + B], lh[j:j + B])   # H2D of the batch inside the timed region
            host_loss.copy_(loss, non_blocking=False)                # D2H of the step's loss, every step
        e1.record()
        env.barrier()
        te = env.max_over_ranks(e0.elapsed_time(e1) / 1e3)
        out["e2e"] = {"value": K * B * world / te, "unit": UNIT, "h2d_bytes_per_step": B * w["T"] * 8 + B * 8 + B * 8,
                      "d2h_bytes_per_step": 8,
                      "api": "FusedTrainStep.step(pinned host X,y,lengths) + loss read-back every step"}
    return out, ts


def _time_graph(call, reps=20):
    """Device time of one `call()` (CUDA-graph replay, CUDA events on the launching stream)."""
    import torch
    for _ in range(3):
        call()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        call()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    a.record()
    for _ in range(reps):
        g.replay()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / 1e3 / reps


def _ncu_traffic(kernel):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed `ncu --set full`
    capture of this kernel (profiles/traffic.json), or None."""
    try:
        return json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get(kernel)
    except (OSError, ValueError):
        return None


def gemm_roofline(m, w, B, peaks):
    """The hoisted input projection of encoder layer 0 (x W_ih^T for all timesteps and both
    directions; Transformer: the fused QKV projection), timed alone.  HBM-bound at these K."""
    import torch
    from slnlp_b200 import _lib
    from slnlp_b200.flat import _TC_GEMM
    lib = _lib.lib
    T, E, H = w["T"], w["E"], w["H"]
    if w["kind"] == "transformer":
        M, N, K, name = B * T, 3 * E, E, "transformer.encoder.layers.0.self_attn.in_proj_weight"
    else:
        G = 4 if w["kind"] == "lstm" else 3
        M, N, K, name = B * T, 2 * G * H, E, "model.encoder.rnn.weight_ih_l0"
    x = torch.randn(M, K, device=m._flat.device)
    out = torch.empty(M, N, device=m._flat.device)
    gws = m._gemm_ws()
    fn, kname = (_TC_GEMM, "gemm_tma_kernel") if m.precision == "bf16" else (lib.slnlp_gemm_f32, "gemm_f32_vec_kernel")
    sec = _time_graph(lambda: _lib.check(fn(0, 1, M, N, K, x.data_ptr(), K, m._ptr(name), K, out.data_ptr(), N, None, 0.0,
                                            gws.data_ptr(), gws.numel(), torch.cuda.current_stream().cuda_stream)))
    by = 4.0 * (M * K + K * N + M * N)
    peak = peaks.get("hbm_gbs", 6650.0)
    return {"kernel": f"{kname} [{M}x{K}]x[{K}x{N}]", "bound": "hbm", "achieved": by / sec / 1e9, "unit": "GB/s",
            "peak": peak, "peak_source": "MEASURED_PEAKS.json hbm_gbs" if peaks else "fallback 6.65 TB/s",
            "frac": by / sec / 1e9 / peak, "us_per_launch": sec * 1e6, "bytes_per_launch": by,
            "tflops": 2.0 * M * N * K / sec / 1e12, "traffic": _ncu_traffic(kname),
            "note": "algorithmic bytes 4(MK+KN+MN); L2-warm graph replay, so a fraction above the HBM-only ceiling is possible"}


def dominant_kernel_roofline(m, ts, w, B, peaks, sustained=False):
    """Time the dominant kernel - the recurrence of encoder layer 0, both directions - alone with
    CUDA events on its launching stream.  bf16 path: ONE persistent launch covers all T steps when the
    batch fits the persistent / cluster kernels, else one rnn_step_fwd_tc_kernel launch per step;
    fp32 path: one rnn_step_fwd_kernel launch per step."""
    import torch
    from slnlp_b200 import _lib
    lib = _lib.lib
    T, H, G = w["T"], w["H"], (4 if w["kind"] == "lstm" else 3)
    ws = ts.ws
    mode = 0 if w["kind"] == "lstm" else 1
    prec = 1 if m.precision == "bf16" else 0

    def call():
        _lib.check(lib.slnlp_rnn_layer_fwd(mode, prec, T, B, H, 2, ws.enc_gates[0].data_ptr(),
                                           m._ptr("model.encoder.rnn.weight_hh_l0"), m._ptr("model.encoder.rnn.bias_hh_l0"),
                                           ts.lengths.data_ptr(), None, None, ws.enc_out[0].data_ptr(),
                                           ws.enc_stash[0].data_ptr(), ws.enc_hfin[0].data_ptr(),
                                           torch.cuda.current_stream().cuda_stream))
    if getattr(ws, "bf_step", False):
        return _large_batch_rooflines(m, ts, w, B, peaks)
    ws.enc_gates[0].normal_(0, 0.5)
    c0 = lib.slnlp_launch_count()
    call()
    launches = max(1, int(lib.slnlp_launch_count() - c0))
    layer_s = _time_graph(call, reps=20 if B <= 256 else 3)
    per_launch_s = layer_s / launches
    flops = 2.0 * B * (G * H) * H * 2 * (T / launches)   # h_{t-1} W_hh^T, both directions, per launch
    if prec == 0:
        kernel = "rnn_step_fwd_kernel"
    else:
        kernel = ("rnn_persistent_fwd_kernel" if H == 128 else "rnn_cluster_fwd_kernel") if launches == 1 else "rnn_step_fwd_tc_kernel"
    key = "bf16_tflops_sustained" if sustained else "bf16_tflops"
    peak = peaks.get(key, 1370.0 if sustained else 1590.0)
    return {"kernel": kernel, "bound": "tensor", "achieved": flops / per_launch_s / 1e12, "unit": "TFLOP/s",
            "peak": peak, "peak_source": f"MEASURED_PEAKS.json {key}" if peaks else "fallback",
            "frac": flops / per_launch_s / 1e12 / peak,
            "us_per_launch": per_launch_s * 1e6, "launches_per_layer": launches, "us_per_timestep": layer_s / T * 1e6,
            "flops_per_launch": flops, "traffic": _ncu_traffic(kernel),
            "note": ("one launch per timestep: [B,H]x[H,4H] per direction, tensor-bound at this batch" if launches > 1 else
                     "at batch 50 the 2*L*T strictly dependent recurrence steps, not FLOPs or bytes, bound this kernel "
                     "(SURVEY.md 8d): read us_per_timestep; the HBM-bound hoisted GEMM is under roofline_gemm")}


def _large_batch_rooflines(m, ts, w, B, peaks):
    """Data-parallel batch sizes (cfg4): the recurrence runs as one persistent CTA-pair kernel launch per timestep
    (rnn_step_pair.cu) whose epilogue moves 46 B per (sequence, unit, direction) - hoisted projection and c_{t-1} in,
    activated gates, c_t, h_t (fp32 + bf16) out: HBM-bound.  The hoisted projection runs on the CTA-pair bf16 GEMM
    (gemm_pair.cu): tensor-bound, reported under `gemm`."""
    import torch
    from slnlp_b200 import _lib
    lib = _lib.lib
    T, H, E = w["T"], w["H"], w["E"]
    ws = ts.ws
    S = lambda: torch.cuda.current_stream().cuda_stream
    ws.enc_gates[0].normal_(0, 0.5)

    def call():
        _lib.check(lib.slnlp_rnn_layer_fwd_bf16(0, T, B, H, 2, ws.enc_gates[0].data_ptr(), ws.w_hh_bf[0].data_ptr(),
                                                m._ptr("model.encoder.rnn.bias_hh_l0"), ts.lengths.data_ptr(),
                                                ws.enc_out[0].data_ptr(), ws.out_bf[0].data_ptr(), ws.enc_stash[0].data_ptr(),
                                                ws.enc_hfin[0].data_ptr(), S()))
    layer_s = _time_graph(call, reps=3)
    per_step = layer_s / T
    by = 46.0 * B * H * 2
    flops = 2.0 * B * (4 * H) * H * 2
    hbm = peaks.get("hbm_gbs", 6650.0)
    rec = {"kernel": "lstm_step_fwd_pair_kernel", "bound": "hbm", "achieved": by / per_step / 1e9, "unit": "GB/s", "peak": hbm,
           "peak_source": "MEASURED_PEAKS.json hbm_gbs" if peaks else "fallback 6.65 TB/s", "frac": by / per_step / 1e9 / hbm,
           "us_per_launch": per_step * 1e6, "launches_per_layer": T, "us_per_timestep": per_step * 1e6,
           "bytes_per_launch": by, "tflops": flops / per_step / 1e12, "traffic": _ncu_traffic("lstm_step_fwd_pair_kernel"),
           "note": "one persistent CTA-pair launch per timestep; algorithmic bytes 46 B per (sequence, unit, direction): "
                   "16 hoisted projection + 4 c_{t-1} in, 16 activated gates + 4 c_t + 4 h_t + 2 h_t(bf16) out"}
    M, N, K = T * B, 8 * H, E
    xb, wb, out = ws.xin_bf[0], ws.w_ih_bf[0], ws.enc_gates[0]
    sec = _time_graph(lambda: _lib.check(lib.slnlp_gemm_bf16(0, 1, M, N, K, xb.data_ptr(), K, wb.data_ptr(), K, out.data_ptr(), N,
                                                              None, 0.0, S())), reps=3)
    tf = peaks.get("bf16_tflops_sustained", 1370.0)
    rec["gemm"] = {"kernel": f"gemm_pair_kernel [{M}x{K}]x[{K}x{N}]", "bound": "tensor", "achieved": 2.0 * M * N * K / sec / 1e12,
                   "unit": "TFLOP/s", "peak": tf, "peak_source": "MEASURED_PEAKS.json bf16_tflops_sustained" if peaks else "fallback",
                   "frac": 2.0 * M * N * K / sec / 1e12 / tf, "us_per_launch": sec * 1e6,
                   "note": "hoisted x W_ih^T of encoder layer 0, bf16 operands, CTA pairs (tcgen05 cta_group::2); the fp32 C "
                           "write (M*N*4 bytes) shares L2 bandwidth with the operand tiles"}
    return rec


# ---------------------------------------------------------------- sub-records
def leg_fp32(env, args, w, data, flush):
    """cfg1 on the fp32 path - the one that carries north_star's 1e-5 / identical-argmax targets."""
    m = make_module(w, data, env.dev, "fp32")
    r, ts = measure_train(env, w, data, m, w["B"], max(3, args.warmup), args.steps, flush)
    return {"workload": w["name"], "dtype": "f32", "value": r["value"], "unit": UNIT, "ms_per_step": r["ms_per_step"],
            "e2e": r["e2e"], "launches_per_step": r["launches_per_step"], "cuda_graph": r["cuda_graph"],
            "parity": "logits/loss 1e-5 relative, greedy decode identical (tests/test_gpu_baseline_golden.py)"}


def leg_infer(env, args, w, data, m):
    """Scoring throughput (main.py:116-117 -> estimator.predict): eval forward + argmax.
    value: batches of 50 resident in HBM, one captured forward per batch, CUDA events.
    e2e: NeuralNetClassifier.predict(host tensors) - H2D of the tokens, forwards, argmax, D2H."""
    import torch
    import model as dropin
    from slnlp_b200.net import NeuralNetClassifier
    from slnlp_b200.rnn import InferStep
    B, T, dev = w["B"], w["T"], env.dev
    Xd, yd, ld = data["X"].to(dev), data["y"].to(dev), data["lengths"].to(dev)
    nb = Xd.shape[0] // B
    step = InferStep(m, B, T)
    pred = torch.empty(nb * B, dtype=torch.int64, device=dev)
    K = max(args.steps, 50)

    def one(i):
        j = (i % nb) * B
        logp = step.step(Xd[j:j + B], yd[j:j + B], ld[j:j + B])
        torch.argmax(logp, dim=1, out=pred[j:j + B])
    for i in range(5):
        one(i)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for i in range(K):
        one(i)
    b.record()
    torch.cuda.synchronize()
    sec = a.elapsed_time(b) / 1e3
    cls = {"lstm": dropin.EncoderDecoderLSTMAttn, "gru": dropin.EncoderDecoderGRUAttn, "transformer": dropin.Transformer}[w["kind"]]
    extra = {"module__num_heads": w["heads"]} if w["kind"] == "transformer" else {}
    net = NeuralNetClassifier(module=cls, batch_size=B, device=str(dev), verbose=0, precision=m.precision,
                              module__src_vocab=data["src_vocab"], module__tgt_vocab=data["tgt_vocab"], module__batch_first=True,
                              module__embedding_size=w["E"], module__hidden_size=w["H"], module__num_layers=w["L"],
                              module__dropout=w["p"], **extra).initialize()
    host = {"X": data["X"], "lengths": data["lengths"], "y": data["y"]}
    net.predict(host)
    torch.cuda.synchronize()
    n = data["X"].shape[0]
    t0 = time.perf_counter()
    reps = 3
    for _ in range(reps):
        out = net.predict(host)
    te = (time.perf_counter() - t0) / reps
    return {"metric": "predict_seq_per_s", "value": K * B / sec, "unit": UNIT, "ms_per_batch": sec / K * 1e3, "batch": B,
            "dtype": "bf16" if m.precision == "bf16" else "f32",
            "e2e": {"value": n / te, "unit": UNIT, "h2d_bytes_per_call": n * T * 8 + 2 * n * 8, "d2h_bytes_per_call": int(out.nbytes),
                    "api": f"NeuralNetClassifier.predict(host X [{n},{T}]): H2D, {n // B} captured forwards, argmax, D2H (wall clock)"}}


def leg_dp(env, args, flush):
    """BASELINE.json configs[3]: LSTM 1024/512/6 dropout 0.5, global batch 4096 split over the ranks
    (strong scaling), gradients exchanged per layer over NCCL while backward runs (slnlp_b200/dp.py)."""
    import torch
    import torch.distributed as dist
    from phono_synth import synthetic_dataset
    from slnlp_b200.dp import BucketedGradSync
    from slnlp_b200.rnn import FusedTrainStep
    w = dict(WORKLOADS["cfg4"])
    Bg = args.dp_batch
    world, rank, dev = env.world, env.rank, env.dev
    assert Bg % world == 0
    B = Bg // world
    peaks = load_peaks()
    data = synthetic_dataset(n_seq=4 * Bg, T=w["T"], v_src=w["Vs"], v_tgt=w["Vt"], seed=1)
    W, K = 3, max(3, min(args.steps, 6))
    m = make_module(w, data, dev, args.precision)
    sync = BucketedGradSync() if world > 1 else None
    r, ts = measure_train(env, w, data, m, B, W, K, flush, grad_sync=sync, split=True, e2e=False)
    flops_seq = train_flops_per_seq(w)
    tf_peak = peaks.get("bf16_tflops_sustained", 1370.0)
    rec = {"workload": w["name"], "global_batch": Bg, "batch_per_gpu": B, "n_gpus": world, "scaling": "strong",
           "value": r["value"], "unit": UNIT, "ms_per_step": r["ms_per_step"], "steps": K, "warmup": W,
           "dtype": "bf16" if args.precision == "bf16" else "f32", "cuda_graph": r["cuda_graph"],
           "launches_per_step": r["launches_per_step"], "params": m._numel,
           "model_tflops_per_gpu": r["value"] * flops_seq / 1e12 / world,
           "model_frac_of_bf16_sustained": r["value"] * flops_seq / 1e12 / world / tf_peak}
    if world == 1:
        rec["roofline"] = dominant_kernel_roofline(m, ts, w, B, peaks, sustained=True)
    else:
        # the exchange alone, and the same step without it: names the limiter
        g = ts.gflat
        for _ in range(2):
            dist.all_reduce(g)
        env.barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(5):
            dist.all_reduce(g)
        b.record()
        env.barrier()
        ar_ms = env.max_over_ranks(a.elapsed_time(b) / 5)
        ts.graph = None
        del ts
        torch.cuda.empty_cache()
        r0, ts0 = measure_train(env, w, data, m, B, 2, K, flush, grad_sync=None, split=True, e2e=False)
        by = 4 * m._numel
        rec["allreduce"] = {"bytes_per_step": by, "collectives_per_step": sync.n_collectives + 1, "ms_alone_one_buffer": ar_ms,
                            "busbw_gbs": 2.0 * (world - 1) / world * by / (ar_ms / 1e3) / 1e9,
                            "overlap": "per encoder layer, issued as soon as that layer's BPTT + dW are enqueued (NCCL stream)"}
        rec["ms_per_step_without_exchange"] = r0["ms_per_step"]
        rec["exposed_exchange_ms"] = r["ms_per_step"] - r0["ms_per_step"]
        rec["limiter"] = ("BPTT / recurrent step kernels (exchange hidden behind backward)"
                          if rec["exposed_exchange_ms"] < 0.1 * r["ms_per_step"] else "gradient all-reduce")
        ts0.graph = None
        del ts0
    del m
    torch.cuda.empty_cache()
    if world > 1 and not args.no_dp_parity:
        rec["dp_parity"] = dp_parity(env, args, w, data, Bg)
    return rec


def dp_parity(env, args, w, data, Bg, steps=2):
    """N ranks of FusedTrainStep + the bucketed exchange against ONE rank running the same global batch:
    loss after `steps` steps and the weights, relative (dropout 0 - masks are drawn per local batch)."""
    import torch
    import torch.distributed as dist
    from slnlp_b200.dp import BucketedGradSync
    from slnlp_b200.rnn import FusedTrainStep
    world, rank, dev = env.world, env.rank, env.dev
    B = Bg // world
    X, y, ln = data["X"][:Bg].clone(), data["y"][:Bg].clone(), data["lengths"][:Bg].clone()
    y[:Bg // 16] = 1                             # ignored labels, all in rank 0's slice: unequal valid counts
    m = make_module(w, data, dev, args.precision, dropout=0.0)
    w0 = m._flat.clone()
    ts = FusedTrainStep(m, B, w["T"], lr=0.01, grad_sync=BucketedGradSync())
    sl = slice(rank * B, (rank + 1) * B)
    Xl, yl, ll = X[sl].to(dev), y[sl].to(dev), ln[sl].to(dev)
    for _ in range(steps):
        loss = ts.step(Xl, yl, ll)
    torch.cuda.synchronize()
    dp_loss, dp_w = float(loss[0]), m._flat.clone()
    ts.graph = None
    del ts
    torch.cuda.empty_cache()
    out = None
    if rank == 0:
        m._flat.copy_(w0)
        ts1 = FusedTrainStep(m, Bg, w["T"], lr=0.01)
        for _ in range(steps):
            loss1 = ts1.step(X.to(dev), y.to(dev), ln.to(dev))
        torch.cuda.synchronize()
        ref_w = m._flat
        upd = (ref_w - w0).abs().max()
        out = {"steps": steps, "global_batch": Bg, "ranks": world, "dropout": 0.0,
               "loss_rel": abs(dp_loss - float(loss1[0])) / abs(float(loss1[0])),
               "w_rel": float((dp_w - ref_w).abs().max() / ref_w.abs().max()),
               "update_rel": float((dp_w - ref_w).abs().max() / upd),
               "note": "rank-0 loss / flat weights after the data-parallel steps vs ONE rank stepping the same global batch; "
                       "update_rel = max |w_dp - w_single| / max |w_single - w_0|; 1/16 of the labels ignored, all on rank 0"}
        ts1.graph = None
        del ts1
    # every rank's weights are identical after the exchange
    chk = torch.stack([dp_w.double().sum(), dp_w.double().abs().sum()])
    lo, hi = chk.clone(), chk.clone()
    dist.all_reduce(lo, op=dist.ReduceOp.MIN)
    dist.all_reduce(hi, op=dist.ReduceOp.MAX)
    if out is not None:
        out["ranks_identical"] = bool(torch.equal(lo, hi))
    del m
    torch.cuda.empty_cache()
    env.barrier()
    return out


def grid_search(env, args, w=None):
    """cfg5: the reference's LSTM grid (config-enc-dec-lstm-attn.yaml:45-51; 162 candidates x 5 folds
    = 810 independent fits) through the estimator + GridSearchFarm, one worker per GPU, fits claimed
    from a store counter - no NCCL.  Bounded so that it finishes in minutes: every fit trains
    --grid-epochs epochs (early stopping off) on a --grid-seqs-sequence synthetic corpus; state it
    when quoting fits/hour."""
    import numpy as np
    import torch
    import helper as h
    from slnlp_b200.data import SeqDataset
    from slnlp_b200.grid import GridSearchFarm
    from slnlp_b200.net import NeuralNetClassifier
    from slnlp_b200 import _lib
    from slnlp_b200 import callbacks as cbs
    import model as dropin
    w = w or WORKLOADS["cfg5"]
    world, local = env.world, env.local
    ds = SeqDataset.synthetic(n_seq=args.grid_seqs, T=w["T"], v_src=w["Vs"], v_tgt=w["Vt"], ragged=True, seed=1)
    grid = {"lr": [0.1, 0.01, 0.001], "module__embedding_size": [1024, 512, 128], "module__hidden_size": [512, 256, 128],
            "module__num_layers": [6, 4, 2], "module__dropout": [0.5, 0.1]}
    if args.grid_fraction < 1.0:      # a deterministic slice of the candidate list, same folds
        grid = {"lr": [0.01], "module__embedding_size": [1024, 512, 128], "module__hidden_size": [512, 256, 128],
                "module__num_layers": [6, 4, 2], "module__dropout": [0.1]}
    net = NeuralNetClassifier(module=dropin.EncoderDecoderLSTMAttn, lr=0.01, max_epochs=args.grid_epochs, batch_size=w["B"],
                              device=f"cuda:{local}", verbose=0, precision=args.precision,
                              module__src_vocab=ds.vocab_X, module__tgt_vocab=ds.vocab_y, module__batch_first=True,
                              optimizer__momentum=0.9, optimizer__nesterov=False, criterion__ignore_index=1,
                              callbacks=[("gradient_clipping", cbs.GradientNormClipping(gradient_clip_value=0.5))])
    y = ds.y().to_array()
    gs = GridSearchFarm(net, grid, cv=5, scoring=h.build_scoring("neg_log_loss", ds.labels(), allow_multiple=False),
                        refit=False, backend="torchrun" if world > 1 else "inline", per_fit_checkpoint_dirs=False,
                        fits_per_gpu=args.fits_per_gpu, procs_per_gpu=args.procs_per_gpu if world == 1 else 1)
    l0 = _lib.lib.slnlp_launch_count()
    env.barrier()
    t0 = time.perf_counter()
    gs.fit(ds.X(), y)
    torch.cuda.synchronize()
    sec = env.max_over_ranks(time.perf_counter() - t0)
    n_fits = gs.n_fits_
    steps_per_fit = args.grid_epochs * int(np.ceil(0.8 * 0.8 * args.grid_seqs / w["B"]))
    busy = {}
    for r in gs.fit_results_.values():
        busy[r.get("gpu", 0)] = busy.get(r.get("gpu", 0), 0.0) + r["fit_time"] + r["score_time"]
    return {"metric": "grid_fits_per_hour", "value": 3600.0 * n_fits / sec, "unit": "fits/hour", "n_gpus": world,
            "fits": n_fits, "search_seconds": sec, "ms_per_fit": sec / n_fits * 1e3, "scaling": "strong",
            "dtype": "f32" if args.precision == "fp32" else "bf16",
            "workload": w["name"] if args.grid_fraction >= 1.0 else w["name"] + " (lr 0.01, dropout 0.1 slice: 27 candidates x 5 folds = 135 fits)",
            "epochs_per_fit": args.grid_epochs, "sequences": args.grid_seqs, "batch": w["B"],
            "train_steps_per_fit": steps_per_fit, "seq_len": w["T"], "v_src": w["Vs"], "v_tgt": w["Vt"],
            "early_stopping": "off (fixed epochs)", "fits_per_gpu": args.fits_per_gpu,
            "procs_per_gpu": args.procs_per_gpu if world == 1 else 1,
            "parallelism": (f"{world} rank(s), one per GPU, x {args.fits_per_gpu} fit(s) at a time per GPU (threads, private streams, "
                            "one captured step graph each), longest-first, fits claimed from a store counter, no collective"
                            if (world > 1 or args.procs_per_gpu <= 1) else
                            f"1 GPU shared by {args.procs_per_gpu} worker processes x {args.fits_per_gpu} fit(s) at a time each (threads, "
                            "private streams, one captured step graph each), longest-first from one task queue, no collective"),
            "gpu_launches_rank0": int(_lib.lib.slnlp_launch_count() - l0) + int(getattr(gs, "worker_launches_", 0)),
            "worker_busy_seconds": {str(k): round(v, 2) for k, v in sorted(busy.items())},
            "best_params": {k: v for k, v in gs.best_params_.items()}, "best_score": gs.best_score_,
            "mean_test_score": [float(v) for v in gs.cv_results_["mean_test_score"]]}


def run_grid(env, args, w):
    """--workload cfg5: the grid search as the headline line."""
    g = grid_search(env, args, w)
    if env.rank != 0:
        return
    line = {"metric": g["metric"], "value": g["value"], "unit": g["unit"], "n_gpus": env.world, "steps": g["fits"], "warmup": 0,
            "ms_per_step": g["ms_per_fit"], "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": g["dtype"], "data": "synthetic", "gpu_launches": g["gpu_launches_rank0"],
            "config": {k: g[k] for k in ("workload", "fits", "epochs_per_fit", "sequences", "batch", "train_steps_per_fit",
                                         "seq_len", "v_src", "v_tgt", "early_stopping", "parallelism", "fits_per_gpu")},
            "search_seconds": g["search_seconds"], "worker_busy_seconds": g["worker_busy_seconds"],
            "best_params": g["best_params"], "best_score": g["best_score"], "mean_test_score": g["mean_test_score"]}
    print(json.dumps(line), flush=True)


def guarded(name, fn, env):
    """A sub-record must never take the headline down: failures are reported in place.  Under
    torchrun a failure on one rank would leave the others in a collective, so there it re-raises."""
    try:
        return fn()
    except Exception as e:  # noqa: BLE001
        if env.world > 1:
            raise
        import traceback
        return {"error": f"{type(e).__name__}: {e}", "trace": traceback.format_exc()[-600:]}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="cfg1", choices=list(WORKLOADS))
    ap.add_argument("--precision", default=os.environ.get("SLNLP_PRECISION", "bf16"), choices=["fp32", "bf16"])
    ap.add_argument("--batch", type=int, default=None)
    ap.add_argument("--dp", action="store_true", help="headline = data-parallel one global batch (NCCL all-reduce)")
    ap.add_argument("--legs", default=None, help="comma list of sub-records (fp32_path,infer,dp,grid), 'all' or 'none'; "
                                                 "default: all for the default cfg1 run, none otherwise")
    ap.add_argument("--dp-batch", type=int, default=4096, help="dp leg: global batch")
    ap.add_argument("--no-dp-parity", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-flush", action="store_true")
    ap.add_argument("--grid-epochs", type=int, default=2, help="cfg5: epochs per fit")
    ap.add_argument("--grid-seqs", type=int, default=500, help="cfg5: sequences in the synthetic corpus")
    ap.add_argument("--grid-fraction", type=float, default=None, help="cfg5: < 1 runs the 27-candidate lr=0.01/dropout=0.1 slice "
                                                                      "(default: the slice as a sub-record, the full grid for --workload cfg5)")
    ap.add_argument("--fits-per-gpu", type=int, default=int(os.environ.get("SLNLP_FITS_PER_GPU", "4")),
                    help="cfg5: fits packed on each GPU (worker threads, private streams)")
    ap.add_argument("--procs-per-gpu", type=int, default=int(os.environ.get("SLNLP_PROCS_PER_GPU", "1")),
                    help="cfg5, one GPU: worker processes sharing the GPU, each with --fits-per-gpu threads")
    args = ap.parse_args()
    w = dict(WORKLOADS[args.workload])
    if args.batch:
        w["B"] = args.batch
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        if w["kind"] == "grid":
            w = dict(WORKLOADS["cfg1"])
        return run_reference(args, w, world)

    import torch
    import torch.distributed as dist
    env = Env()
    if env.world > 1:
        import datetime
        dist.init_process_group("nccl", device_id=env.dev, timeout=datetime.timedelta(minutes=10))
    if w["kind"] == "grid":
        if args.grid_fraction is None:
            args.grid_fraction = 1.0
        run_grid(env, args, w)
        return finish(env)
    if args.grid_fraction is None:
        args.grid_fraction = 1.0        # the sub-record runs the FULL 162-candidate x 5-fold grid (810 fits)
    default_run = args.workload == "cfg1" and not args.dp and not args.batch and args.precision == "bf16"
    legs = args.legs if args.legs is not None else ("all" if default_run else "none")
    legs = set(ALL_LEGS) if legs == "all" else set() if legs == "none" else set(legs.split(","))

    from phono_synth import synthetic_dataset
    from slnlp_b200 import _lib   # noqa: F401
    rank, dev = env.rank, env.dev
    W, K = max(3, args.warmup), args.steps
    B = w["B"] // env.world if args.dp else w["B"]
    n_seq = max(5000 if w["B"] <= 50 else 4 * w["B"], B * 4)
    data = synthetic_dataset(n_seq=n_seq, T=w["T"], v_src=w["Vs"], v_tgt=w["Vt"], seed=1 + (0 if args.dp else rank))
    flush = None if args.no_flush else torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    m = make_module(w, data, dev, args.precision)
    sync = None
    if args.dp and env.world > 1:
        from slnlp_b200.dp import BucketedGradSync
        sync = BucketedGradSync()
    r, ts = measure_train(env, w, data, m, B, W, K, flush, grad_sync=sync, split=args.dp, e2e=True, clocks=True)
    peaks = load_peaks()
    roof_gemm = gemm_roofline(m, w, B, peaks)
    roof = dominant_kernel_roofline(m, ts, w, B, peaks) if w["kind"] != "transformer" else dict(roof_gemm)
    flops_seq = train_flops_per_seq(w)
    tf_peak = peaks.get("bf16_tflops_sustained", 1400.0)
    line = {
        "metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": env.world, "steps": K, "warmup": W,
        "ms_per_step": r["ms_per_step"], "higher_is_better": True,
        "scaling": "strong" if args.dp else "weak", "vs_baseline": None,
        "dtype": "f32" if args.precision == "fp32" else "bf16", "data": "synthetic",
        "config": common_config(w, args, env.world),
        "run": {"batch_per_gpu": B, "global_batch": B * env.world, "params": m._numel, "cuda_graph": r["cuda_graph"]},
        "e2e": r["e2e"],
        "gpu_launches": int(r["launches_per_step"] * K),
        "launches_per_step": r["launches_per_step"],
        "clocks": r["clocks"],
        "wall_ms_per_step_incl_flush": r["wall_ms_per_step_incl_flush"],
        "final_loss": r["final_loss"],
        "model_tflops": r["value"] * flops_seq / 1e12,
        "model_frac_of_tensor_peak": r["value"] * flops_seq / 1e12 / tf_peak / env.world,
        "train_mflop_per_seq": flops_seq / 1e6,
        "roofline": roof,
        "roofline_gemm": roof_gemm,
    }
    if "infer" in legs and env.world == 1:
        line["infer"] = guarded("infer", lambda: leg_infer(env, args, w, data, m), env)
    ts.graph = None
    del ts, m
    torch.cuda.empty_cache()
    if "fp32_path" in legs and env.world == 1 and args.precision == "bf16":
        line["fp32_path"] = guarded("fp32_path", lambda: leg_fp32(env, args, w, data, flush), env)
        torch.cuda.empty_cache()
    if "dp" in legs:
        line["dp"] = guarded("dp", lambda: leg_dp(env, args, flush), env)
        torch.cuda.empty_cache()
    if "grid" in legs:
        line["grid"] = guarded("grid", lambda: grid_search(env, args), env)
    if rank == 0 and not args.no_cpu_baseline and env.world == 1:   # the CPU baseline is an N = 1 figure (rank 0 only)
        try:
            ref = cpu_baseline_subprocess(args, args.workload)
            line["cpu_baseline"] = ref["cpu_baseline"]
            line["speedup_e2e_vs_cpu"] = r["e2e"]["value"] / ref["value"]
            if "infer" in ref and isinstance(line.get("infer"), dict) and "value" in line["infer"]:
                line["infer"]["cpu_baseline"] = dict(ref["infer"], cores=ref["cpu_baseline"]["cores"], kind=ref["cpu_baseline"]["kind"])
                line["infer"]["speedup_e2e_vs_cpu"] = line["infer"]["e2e"]["value"] / ref["infer"]["value"]
        except Exception as e:  # noqa: BLE001
            line["cpu_baseline"] = {"error": f"{type(e).__name__}: {e}"}
    if rank == 0:
        print(json.dumps(line), flush=True)
    finish(env)


def finish(env):
    """Leave without tearing CUDA / NCCL down under live graphs: everything measured is printed; a rank
    that is done synchronises, meets the others and exits hard (no communicator or graph destructors
    racing interpreter shutdown, no hang at exit)."""
    import torch
    sys.stdout.flush()
    sys.stderr.flush()
    try:
        torch.cuda.synchronize()
        if env.world > 1:
            import torch.distributed as dist
            dist.barrier()
            torch.cuda.synchronize()
    finally:
        os._exit(0)


if __name__ == "__main__":
    main()
