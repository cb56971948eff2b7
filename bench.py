#!/usr/bin/env python
"""Headline benchmark: training sequences/s of one fit of the reference's
EncoderDecoderLSTMAttn on one B200, next to the reference's CPU path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload cfg1|cfg2|cfg3|cfg4|cfg5] [--precision bf16|fp32]
    python bench.py --impl reference ...        # the reference's CPU path (oracle port)

cfg1 (default) = BASELINE.json configs[0], the configuration the metric is quoted on; cfg2 = GRU
512/256/4; cfg3 = Transformer 512/256/4/h8; cfg4 = LSTM 1024/512/6 at batch 4096 (``--dp`` splits it
over the ranks with an NCCL gradient all-reduce); cfg5 = the LSTM hyper-parameter grid farmed over
the GPUs (fits/hour).

A "step" is one skorch-equivalent training step (forward, CrossEntropyLoss on the
log-probs, backward, global-norm clip 0.5, SGD momentum 0.9) on one batch of synthetic
6-field phonology sequences.  At N > 1 every rank runs its own
independent fit (the reference's only parallelism is farming grid-search fits, one per
GPU, with no collective: SURVEY.md section 8e) - weak scaling; ``--dp`` instead splits
one global batch across ranks with an NCCL gradient all-reduce.

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
}
    # BASELINE.json configs[4]: the full config-enc-dec-lstm-attn grid, farmed over the GPUs (no collective)
WORKLOADS["cfg5"] = dict(kind="grid", B=50, T=64, Vs=4098, Vt=52, E=0, H=0, L=0, p=0.0,
                         name="config-enc-dec-lstm-attn grid: 3 lr x 3 emb x 3 hidden x 3 layers x 2 dropout x 5-fold CV = 810 fits")
METRIC, UNIT = "train_seq_per_s", "sequences/s"


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


def build_reference_port(w):
    import torch
    from oracle import port
    torch.manual_seed(1)
    return port.build_port(w["kind"], w["Vs"], w["Vt"], w["E"], w["H"], w["L"], dropout=w["p"], num_heads=w.get("heads"))


def time_cpu_port(w, data, steps, warmup, batch=None):
    """The reference's CPU path (torch.nn port of its modules, oracle/port.py) on the host cores."""
    import torch
    from oracle import port
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    B = batch or min(w["B"], 50)
    ref = build_reference_port(w)
    opt = torch.optim.SGD(ref.parameters(), lr=0.01, momentum=0.9, nesterov=False)
    X, y, lengths = data["X"], data["y"], data["lengths"]
    nb = X.shape[0] // B
    times = []
    for i in range(warmup + steps):
        j = (i % nb) * B
        t0 = time.perf_counter()
        port.reference_train_step(ref, opt, X[j:j + B], y[j:j + B], lengths[j:j + B])
        if i >= warmup:
            times.append(time.perf_counter() - t0)
    sec = statistics.median(times)
    return B / sec, sec, cores, B


def cpu_model_name():
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("model name"):
                return line.split(":", 1)[1].strip()
    except OSError:
        pass
    return "unknown"


def run_reference(args, w):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from slnlp_b200.data import synthetic_dataset
    data = synthetic_dataset(n_seq=max(500, 50 * 10), T=w["T"], v_src=w["Vs"], v_tgt=w["Vt"], seed=1)
    steps, warmup = max(1, args.steps), max(1, args.warmup)
    if w["B"] > 50:  # cfg4 on CPU: batch 50 sample, reported per sequence
        steps, warmup = min(steps, 3), 1
    val, sec, cores, B = time_cpu_port(w, data, steps, warmup)
    line = {"metric": METRIC, "value": val, "unit": UNIT, "impl": "reference", "n_gpus": args.gpus,
            "steps": steps, "warmup": warmup, "ms_per_step": sec * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": w["name"], "batch": B, "seq_len": w["T"], "v_src": w["Vs"], "v_tgt": w["Vt"]},
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port", "cpu": cpu_model_name(),
                             "sample": f"{steps} training steps of batch {B} (median), torch {__import__('torch').__version__} CPU, "
                                       "oracle/port.py = the reference modules on stock torch.nn"},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def run_grid(args, w):
    """cfg5: the reference's LSTM grid (config-enc-dec-lstm-attn.yaml:45-51; 162 candidates x 5 folds
    = 810 independent fits) through the estimator + GridSearchFarm, one worker per GPU, fits claimed
    from a store counter - no NCCL.  Bounded so that it finishes in minutes: every fit trains
    --grid-epochs epochs (early stopping off) on a --grid-seqs-sequence synthetic corpus; state it
    when quoting fits/hour.  --impl reference runs the same fits on the CPU port for a sample of the grid."""
    import numpy as np
    import torch
    import helper as h
    from slnlp_b200.data import SeqDataset
    from slnlp_b200.grid import GridSearchFarm
    from slnlp_b200.net import NeuralNetClassifier
    from slnlp_b200 import _lib
    import model as dropin
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    ds = SeqDataset.synthetic(n_seq=args.grid_seqs, T=w["T"], v_src=w["Vs"], v_tgt=w["Vt"], ragged=True, seed=1)
    grid = {"lr": [0.1, 0.01, 0.001], "module__embedding_size": [1024, 512, 128], "module__hidden_size": [512, 256, 128],
            "module__num_layers": [6, 4, 2], "module__dropout": [0.5, 0.1]}
    if args.grid_fraction < 1.0:      # a deterministic slice of the candidate list (every k-th), same folds
        grid = {"lr": [0.01], "module__embedding_size": [1024, 512, 128], "module__hidden_size": [512, 256, 128],
                "module__num_layers": [6, 4, 2], "module__dropout": [0.1]}
    from slnlp_b200 import callbacks as cbs
    net = NeuralNetClassifier(module=dropin.EncoderDecoderLSTMAttn, lr=0.01, max_epochs=args.grid_epochs, batch_size=w["B"],
                              device=f"cuda:{local}", verbose=0, precision=args.precision,
                              module__src_vocab=ds.vocab_X, module__tgt_vocab=ds.vocab_y, module__batch_first=True,
                              optimizer__momentum=0.9, optimizer__nesterov=False, criterion__ignore_index=1,
                              callbacks=[("gradient_clipping", cbs.GradientNormClipping(gradient_clip_value=0.5))])
    y = ds.y().to_array()
    gs = GridSearchFarm(net, grid, cv=5, scoring=h.build_scoring("neg_log_loss", ds.labels(), allow_multiple=False),
                        refit=False, backend="torchrun" if world > 1 else "inline", per_fit_checkpoint_dirs=False,
                        fits_per_gpu=args.fits_per_gpu)
    l0 = _lib.lib.slnlp_launch_count()
    t0 = time.perf_counter()
    gs.fit(ds.X(), y)
    torch.cuda.synchronize()
    sec = time.perf_counter() - t0
    if rank != 0:
        return
    n_fits = gs.n_fits_
    steps_per_fit = args.grid_epochs * int(np.ceil(0.8 * 0.8 * args.grid_seqs / w["B"]))
    busy = {}
    for r in gs.fit_results_.values():
        busy[r.get("gpu", 0)] = busy.get(r.get("gpu", 0), 0.0) + r["fit_time"] + r["score_time"]
    line = {"metric": "grid_fits_per_hour", "value": 3600.0 * n_fits / sec, "unit": "fits/hour", "n_gpus": world,
            "steps": n_fits, "warmup": 0, "ms_per_step": sec / n_fits * 1e3, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f32" if args.precision == "fp32" else "bf16", "data": "synthetic",
            "config": {"workload": w["name"] if args.grid_fraction >= 1.0 else w["name"] + " (lr 0.01, dropout 0.1 slice: 27 candidates x 5 folds)",
                       "fits": n_fits, "epochs_per_fit": args.grid_epochs, "sequences": args.grid_seqs, "batch": w["B"],
                       "train_steps_per_fit": steps_per_fit, "seq_len": w["T"], "v_src": w["Vs"], "v_tgt": w["Vt"], "early_stopping": "off (fixed epochs)",
                       "parallelism": f"{world} worker(s), one per GPU, longest-first, fits claimed from a TCPStore counter, no collective",
                       "fits_per_gpu": args.fits_per_gpu if world == 1 else 1},
            "gpu_launches": int(_lib.lib.slnlp_launch_count() - l0), "search_seconds": sec,
            "worker_busy_seconds": {str(k): round(v, 2) for k, v in sorted(busy.items())},
            "best_params": {k: v for k, v in gs.best_params_.items()}, "best_score": gs.best_score_}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="cfg1", choices=list(WORKLOADS))
    ap.add_argument("--precision", default=os.environ.get("SLNLP_PRECISION", "bf16"), choices=["fp32", "bf16"])
    ap.add_argument("--batch", type=int, default=None)
    ap.add_argument("--dp", action="store_true", help="data-parallel one global batch (NCCL all-reduce)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-flush", action="store_true")
    ap.add_argument("--grid-epochs", type=int, default=2, help="cfg5: epochs per fit")
    ap.add_argument("--grid-seqs", type=int, default=500, help="cfg5: sequences in the synthetic corpus")
    ap.add_argument("--grid-fraction", type=float, default=1.0, help="cfg5: < 1 runs the 27-candidate lr=0.01/dropout=0.1 slice")
    ap.add_argument("--fits-per-gpu", type=int, default=1, help="cfg5, one GPU: fits packed on the GPU (worker threads, private streams)")
    args = ap.parse_args()
    w = dict(WORKLOADS[args.workload])
    if args.batch:
        w["B"] = args.batch
    if w["kind"] == "grid":
        return run_grid(args, w)
    if args.impl == "reference":
        return run_reference(args, w)

    import torch
    import torch.distributed as dist
    import model as dropin
    from slnlp_b200 import _lib
    from slnlp_b200.data import synthetic_dataset
    from slnlp_b200.rnn import FusedTrainStep

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    W = max(3, args.warmup)
    K = args.steps
    B = w["B"] // world if args.dp else w["B"]

    n_seq = max(5000 if w["B"] <= 50 else 8 * w["B"], B * 4)
    data = synthetic_dataset(n_seq=n_seq, T=w["T"], v_src=w["Vs"], v_tgt=w["Vt"], seed=1 + (0 if args.dp else rank))
    torch.manual_seed(1)
    cls = {"lstm": dropin.EncoderDecoderLSTMAttn, "gru": dropin.EncoderDecoderGRUAttn, "transformer": dropin.Transformer}[w["kind"]]
    extra = {"num_heads": w["heads"]} if w["kind"] == "transformer" else {}
    m = cls(src_vocab=data["src_vocab"], tgt_vocab=data["tgt_vocab"], batch_first=True, embedding_size=w["E"],
            hidden_size=w["H"], num_layers=w["L"], dropout=w["p"], device=dev, precision=args.precision, **extra).to(dev).train()

    grad_sync = None
    if args.dp and world > 1:
        from slnlp_b200.dp import sync_gradients     # count-weighted all-reduce: exact global-batch gradient
        grad_sync = sync_gradients
    ts = FusedTrainStep(m, B, w["T"], lr=0.01, momentum=0.9, max_norm=0.5, grad_sync=grad_sync)

    # ---------------- device-resident leg: whole dataset in HBM, batches sliced on device
    Xd, yd, ld = data["X"].to(dev), data["y"].to(dev), data["lengths"].to(dev)
    if args.dp and world > 1:  # each rank takes its slice of every global batch
        Xd, yd, ld = Xd[rank::world].contiguous(), yd[rank::world].contiguous(), ld[rank::world].contiguous()
    nb = Xd.shape[0] // B
    flush = None if args.no_flush else torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def batch(i):
        j = (i % nb) * B
        return Xd[j:j + B], yd[j:j + B], ld[j:j + B]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for i in range(W):
        ts.step(*batch(i))
    barrier()
    l0 = _lib.lib.slnlp_launch_count()
    if not ts.use_graph:
        ts.step(*batch(0))
    launches_per_step = None
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
    losses = []
    barrier()
    with ClockSampler(local) as clocks:
        t_wall0 = time.perf_counter()
        for i in range(K):
            if flush is not None:
                flush.zero_()                       # evict L2 between timed steps (not timed)
            ts.load_batch(*batch(W + i))
            ev[i][0].record()
            ts.run()
            ev[i][1].record()
        barrier()
        t_wall = time.perf_counter() - t_wall0
    step_ms = [a.elapsed_time(b) for a, b in ev]
    dev_s = sum(step_ms) / 1e3
    tt = torch.tensor([dev_s], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    dev_s = float(tt)
    total_seqs = K * B * world
    value = total_seqs / dev_s
    final_loss = float(ts.ws.loss[0])

    # kernels per step: count one un-captured step through the C ABI
    c0 = _lib.lib.slnlp_launch_count()
    ts._step()
    torch.cuda.synchronize()
    launches_per_step = _lib.lib.slnlp_launch_count() - c0

    # ---------------- e2e leg: host (pinned) batches through the public step API, loss read back
    Xh, yh, lh = data["X"].pin_memory(), data["y"].pin_memory(), data["lengths"].pin_memory()
    if args.dp and world > 1:
        Xh, yh, lh = Xh[rank::world].contiguous().pin_memory(), yh[rank::world].contiguous().pin_memory(), lh[rank::world].contiguous().pin_memory()
    Ke = K
    for i in range(3):
        j = (i % nb) * B
        float(ts.step(Xh[j:j + B], yh[j:j + B], lh[j:j + B])[0])
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    host_loss = torch.empty(2, pin_memory=True)
    e0.record()
    for i in range(Ke):
        j = ((W + i) % nb) * B
        loss = ts.step(Xh[j:j + B], yh[j:j + B], lh[j:j + B])   # H2D of the batch inside the timed region
        host_loss.copy_(loss, non_blocking=False)                # D2H of the step's loss, every step
    e1.record()
    barrier()
    te = torch.tensor([e0.elapsed_time(e1) / 1e3], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e = Ke * B * world / float(te)
    h2d = B * w["T"] * 8 + B * 8 + B * 8
    d2h = 8

    # ---------------- dominant kernel, timed alone: the encoder layer-0 recurrence (T launches)
    peaks0 = {}
    try:
        peaks0 = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except (OSError, ValueError):
        pass
    roof_gemm = gemm_roofline(m, w, B, peaks0)
    roof = dominant_kernel_roofline(m, ts, w, B, dev) if w["kind"] != "transformer" else dict(roof_gemm)

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except OSError:
            pass
        flops_seq = train_flops_per_seq(w)
        tf_peak = peaks.get("bf16_tflops_sustained", 1400.0)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": dev_s / K * 1e3, "higher_is_better": True,
            "scaling": "strong" if args.dp else "weak", "vs_baseline": None,
            "dtype": "f32" if args.precision == "fp32" else "bf16", "data": "synthetic",
            "config": {"workload": w["name"], "batch_per_gpu": B, "global_batch": B * world, "seq_len": w["T"],
                       "v_src": w["Vs"], "v_tgt": w["Vt"], "params": m._numel,
                       "parallelism": ("dp%d (NCCL all-reduce)" % world) if args.dp else
                                      ("%d independent fits (grid-search farm, no collective)" % world),
                       "l2": "none (--no-flush)" if flush is None else "256 MiB zero-fill between timed steps (outside the timed events)",
                       "cuda_graph": bool(ts.use_graph), "optimizer": "SGD momentum 0.9, global-norm clip 0.5, lr 0.01"},
            "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "api": "FusedTrainStep.step(pinned host X,y,lengths) + loss read-back every step"},
            "gpu_launches": int(launches_per_step * K),
            "launches_per_step": int(launches_per_step),
            "clocks": clocks.summary(),
            "wall_ms_per_step_incl_flush": t_wall / K * 1e3,
            "final_loss": final_loss,
            "model_tflops": value * flops_seq / 1e12,
            "model_frac_of_tensor_peak": value * flops_seq / 1e12 / tf_peak,
            "train_mflop_per_seq": flops_seq / 1e6,
            "roofline": roof,
            "roofline_gemm": roof_gemm,
        }
        if roof is not None and roof.get("bound") == "tensor":
            roof["peak"] = peaks.get("bf16_tflops", 1590.0)
            roof["peak_source"] = "MEASURED_PEAKS.json bf16_tflops (burst)" if peaks else "fallback 1.59 PFLOP/s"
            roof["frac"] = roof["achieved"] / roof["peak"]
        if not args.no_cpu_baseline and world == 1:   # the CPU baseline is an N = 1 figure (rank 0 only)
            cdata = {k: (v[:1000] if hasattr(v, "shape") else v) for k, v in data.items()}
            cval, csec, cores, cB = time_cpu_port(w, cdata, steps=12 if w["B"] <= 50 and w["H"] <= 128 else 3, warmup=2)
            line["cpu_baseline"] = {"value": cval, "unit": UNIT, "cores": cores, "kind": "port", "cpu": cpu_model_name(),
                                    "sample": f"median of {12 if w['B'] <= 50 and w['H'] <= 128 else 3} training steps of batch {cB} "
                                              f"({csec * 1e3:.0f} ms/step) on the box's host cores; oracle/port.py = the reference "
                                              "modules on stock torch.nn (the Python reference cannot travel to the box)"}
            line["speedup_e2e_vs_cpu"] = e2e / cval
        print(json.dumps(line), flush=True)
    if world > 1:
        # a captured data-parallel step holds NCCL kernels: release the graph before the communicator
        ts.graph = None
        torch.cuda.synchronize()
        dist.destroy_process_group()


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
    fn, kname = (lib.slnlp_gemm_tf32, "gemm_tma_kernel") if m.precision == "bf16" else (lib.slnlp_gemm_f32, "gemm_f32_vec_kernel")
    sec = _time_graph(lambda: _lib.check(fn(0, 1, M, N, K, x.data_ptr(), K, m._ptr(name), K, out.data_ptr(), N, None, 0.0,
                                            gws.data_ptr(), gws.numel(), torch.cuda.current_stream().cuda_stream)))
    by = 4.0 * (M * K + K * N + M * N)
    peak = peaks.get("hbm_gbs", 6650.0)
    return {"kernel": f"{kname} [{M}x{K}]x[{K}x{N}]", "bound": "hbm", "achieved": by / sec / 1e9, "unit": "GB/s",
            "peak": peak, "peak_source": "MEASURED_PEAKS.json hbm_gbs" if peaks else "fallback 6.65 TB/s",
            "frac": by / sec / 1e9 / peak, "us_per_launch": sec * 1e6, "bytes_per_launch": by,
            "tflops": 2.0 * M * N * K / sec / 1e12, "traffic": _ncu_traffic(kname),
            "note": "algorithmic bytes 4(MK+KN+MN); L2-warm graph replay, so a fraction above the HBM-only ceiling is possible"}


def dominant_kernel_roofline(m, ts, w, B, dev):
    """Time the dominant kernel - the recurrence of encoder layer 0, both directions - alone with
    CUDA events on its launching stream.  bf16 path: ONE persistent launch covers all T steps when
    H = 128 (rnn_persistent_fwd_kernel), else one rnn_step_fwd_tc_kernel launch per step; fp32 path:
    one rnn_step_fwd_kernel launch per step."""
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
    ws.enc_gates[0].normal_(0, 0.5)
    c0 = lib.slnlp_launch_count()
    call()
    launches = max(1, int(lib.slnlp_launch_count() - c0))
    layer_s = _time_graph(call)
    per_launch_s = layer_s / launches
    flops = 2.0 * B * (G * H) * H * 2 * (T / launches)   # h_{t-1} W_hh^T, both directions, per launch
    if prec == 0:
        kernel = "rnn_step_fwd_kernel"
    else:
        kernel = ("rnn_persistent_fwd_kernel" if H == 128 else "rnn_cluster_fwd_kernel") if launches == 1 else "rnn_step_fwd_tc_kernel"
    return {"kernel": kernel, "bound": "tensor", "achieved": flops / per_launch_s / 1e12, "unit": "TFLOP/s",
            "us_per_launch": per_launch_s * 1e6, "launches_per_layer": launches, "us_per_timestep": layer_s / T * 1e6,
            "flops_per_launch": flops, "traffic": _ncu_traffic(kernel),
            "note": "at batch 50 the 2*L*T strictly dependent recurrence steps, not FLOPs or bytes, bound this kernel "
                    "(SURVEY.md 8d): read us_per_timestep; the HBM-bound hoisted GEMM is under roofline_gemm"}


if __name__ == "__main__":
    main()
