"""Achieved HBM GB/s of the bandwidth-bound kernels (embedding gather / scatter, Bahdanau attention
step, fused log-softmax + CE, gradient norm, clip + SGD-momentum) against the measured HBM peak,
at the headline config (cfg1: B 50, T 64, E=H=128, 1.99 M params) and at the large config
(cfg4: B 4096, E 1024, H 512, 61.9 M params).  Each kernel is timed as a CUDA-graph replay with a
256 MiB L2 flush between replays (outside the events):   python profiles/bench_hbm_kernels.py"""
import ctypes, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "sign-language-nlp_b200"))
import torch
from slnlp_b200 import _lib as L
lib = L.lib
peak = 6650.0
try:
    peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    pass
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
S = lambda: torch.cuda.current_stream().cuda_stream


def timed(call, reps=10):
    for _ in range(3):
        call()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        call()
    tot = 0.0
    for _ in range(reps):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); g.replay(); b.record(); torch.cuda.synchronize()
        tot += a.elapsed_time(b)
    return tot / reps * 1e-3


def row(name, by, sec):
    print(f"{name:44s} {by / 1e6:10.2f} MB {sec * 1e6:9.1f} us {by / sec / 1e9:8.0f} GB/s {100 * by / sec / 1e9 / peak:6.1f} % of measured HBM peak")


for tag, B, T, E, H, V, Vt, nparam in (("cfg1", 50, 64, 128, 128, 4098, 1026, 1_989_632), ("cfg4", 4096, 64, 1024, 512, 4098, 1026, 61_900_000)):
    print(f"--- {tag}: B={B} T={T} E={E} H={H}  (HBM peak {peak:.0f} GB/s)")
    table = torch.randn(V, E, device="cuda")
    idx = torch.randint(2, V, (B, T), device="cuda")
    out = torch.empty(T, B, E, device="cuda")
    f_off, f_w, f_rows = (ctypes.c_int64 * 1)(0), (ctypes.c_int * 1)(E), (ctypes.c_int64 * 1)(V)
    sec = timed(lambda: L.check(lib.slnlp_embed_gather_fwd(table.data_ptr(), idx.data_ptr(), out.data_ptr(), B, T, 1, f_off, f_w, f_rows, 1, 1.0, None, S())))
    row("embed_gather_fwd (idx + row read + row write)", B * T * (8 + 2 * 4 * E), sec)
    dtab = torch.zeros_like(table)
    sec = timed(lambda: L.check(lib.slnlp_embed_gather_bwd(dtab.data_ptr(), idx.data_ptr(), out.data_ptr(), B, T, 1, f_off, f_w, f_rows, 1, 1.0, 1, S())))
    row("embed_gather_bwd (atomic scatter-add)", B * T * (8 + 3 * 4 * E), sec)
    q, pk, v = torch.randn(B, H, device="cuda"), torch.randn(T, B, H, device="cuda"), torch.randn(H, device="cuda")
    val = torch.randn(T, B, 2 * H, device="cuda")
    alpha, ctx = torch.empty(B, T, device="cuda"), torch.empty(B, 2 * H, device="cuda")
    sec = timed(lambda: L.check(lib.slnlp_attn_step_fwd(q.data_ptr(), pk.data_ptr(), v.data_ptr(), val.data_ptr(), idx.data_ptr(), 1, T, B, H, 2 * H,
                                                        alpha.data_ptr(), ctx.data_ptr(), S())))
    row("attn_step_fwd (K + enc_out + mask read once)", B * T * (12 * H + 8), sec)
    dctx = torch.randn(B, 2 * H, device="cuda")
    dval, dpk, dq, dvp = torch.empty_like(val), torch.empty_like(pk), torch.empty_like(q), torch.empty_like(q)
    sec = timed(lambda: L.check(lib.slnlp_attn_step_bwd(dctx.data_ptr(), q.data_ptr(), pk.data_ptr(), v.data_ptr(), val.data_ptr(), alpha.data_ptr(), T, B, H,
                                                        2 * H, dval.data_ptr(), dpk.data_ptr(), dq.data_ptr(), dvp.data_ptr(), S())))
    row("attn_step_bwd (read K, enc_out; write dK, dV)", B * T * (2 * 12 * H), sec)
    logits, y = torch.randn(B, Vt, device="cuda"), torch.randint(2, Vt, (B,), device="cuda")
    logp, loss, dl, ws = torch.empty_like(logits), torch.zeros(2, device="cuda"), torch.empty(B, Vt + 2, device="cuda"), torch.empty(3 * B, device="cuda")
    sec = timed(lambda: L.check(lib.slnlp_logsoftmax_ce_fused(logits.data_ptr(), y.data_ptr(), 1, B, Vt, logp.data_ptr(), loss.data_ptr(), dl.data_ptr(), Vt + 2,
                                                              ws.data_ptr(), S())))
    row("logsoftmax_ce_fused (logits -> logp, dlogits)", B * Vt * 12, sec)
    p, g, buf = torch.randn(nparam, device="cuda"), torch.randn(nparam, device="cuda"), torch.zeros(nparam, device="cuda")
    part, norm = torch.zeros(lib.slnlp_sumsq_partials(), device="cuda"), torch.zeros(1, device="cuda")
    hyper = torch.tensor([0.01, 0.9, 0.5, 0.0], device="cuda")
    sec = timed(lambda: L.check(lib.slnlp_gradnorm(g.data_ptr(), nparam, part.data_ptr(), norm.data_ptr(), S())))
    row("gradnorm (4 B/param)", 4 * nparam, sec)
    sec = timed(lambda: L.check(lib.slnlp_sgd_momentum_clip(p.data_ptr(), g.data_ptr(), buf.data_ptr(), nparam, hyper.data_ptr(), norm.data_ptr(), 1.0, S())))
    row("sgd_momentum_clip (20 B/param)", 20 * nparam, sec)
    del table, idx, out, dtab, pk, val, dval, dpk, p, g, buf
    torch.cuda.empty_cache()
