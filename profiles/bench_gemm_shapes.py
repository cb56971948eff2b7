"""Time every distinct GEMM shape of one cfg1 / cfg2 / cfg3 train step in isolation (CUDA-graph
replay of 20 calls, L2-warm), for the TMA-fed tf32 and the fp32-FMA kernels:
    python profiles/bench_gemm_shapes.py [cfg1|cfg2|cfg3]
Prints us per call, achieved TFLOP/s and GB/s of algorithmic bytes 4(MK+KN+MN)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "sign-language-nlp_b200"))
import torch
from slnlp_b200 import _lib as L

cfg = sys.argv[1] if len(sys.argv) > 1 else "cfg1"
B, T = 50, 64
R = B * T
if cfg == "cfg1":
    E, H, G = 128, 128, 4
elif cfg == "cfg2":
    E, H, G = 512, 256, 3
if cfg in ("cfg1", "cfg2"):
    GH = G * H
    shapes = [("inproj l0", 0, 1, R, 2 * GH, E), ("inproj l1", 0, 1, R, 2 * GH, 2 * H), ("key", 0, 1, R, H, 2 * H),
              ("dec ih", 0, 1, B, GH, E + 2 * H), ("gen", 0, 1, B, 1026, H), ("gen dW", 1, 0, 1026, H, B),
              ("gen dx", 0, 0, B, H, 1026), ("key dW", 1, 0, H, 2 * H, R), ("key dx", 0, 0, R, 2 * H, H),
              ("dW_ih l1", 1, 0, 2 * GH, 2 * H, R), ("dW_ih l0", 1, 0, 2 * GH, E, R), ("dW_hh", 1, 0, GH, H, R - B),
              ("dx l1", 0, 0, R, 2 * H, 2 * GH), ("d_emb", 0, 0, R, E, 2 * GH)]
else:
    E, F = 512, 256
    shapes = [("qkv", 0, 1, R, 3 * E, E), ("out", 0, 1, R, E, E), ("ffn1", 0, 1, R, F, E), ("ffn2", 0, 1, R, E, F),
              ("kv cross", 0, 1, R, 2 * E, E), ("qkv dW", 1, 0, 3 * E, E, R), ("qkv dx", 0, 0, R, E, 3 * E),
              ("out dW", 1, 0, E, E, R), ("ffn1 dW", 1, 0, F, E, R), ("ffn2 dW", 1, 0, E, F, R), ("ffn2 dx", 0, 0, R, F, E)]
ws = torch.empty(L.lib.slnlp_gemm_workspace_floats(), device="cuda")
S = torch.cuda.current_stream().cuda_stream
print(f"{'gemm':12s} tA tB {'M':>6} {'N':>6} {'K':>6} | {'tf32/TMA us':>11} {'TF/s':>7} {'GB/s':>7} | {'f32 us':>8}")
for name, tA, tB, M, N, K in shapes:
    A = torch.randn((K, M) if tA else (M, K), device="cuda")
    Bm = torch.randn((N, K) if tB else (K, N), device="cuda")
    C = torch.zeros(M, N, device="cuda")
    bias = torch.randn(N, device="cuda")
    res = []
    for fn in (L.lib.slnlp_gemm_tf32, L.lib.slnlp_gemm_f32):
        call = lambda: L.check(fn(tA, tB, M, N, K, A.data_ptr(), A.shape[1], Bm.data_ptr(), Bm.shape[1], C.data_ptr(), N,
                                  bias.data_ptr(), 0.0, ws.data_ptr(), ws.numel(), torch.cuda.current_stream().cuda_stream))
        for _ in range(3):
            call()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            for _ in range(20):
                call()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        a.record()
        for _ in range(5):
            g.replay()
        b.record()
        torch.cuda.synchronize()
        res.append(a.elapsed_time(b) * 1e3 / 100)
    fl, by = 2.0 * M * N * K, 4.0 * (M * K + K * N + M * N)
    print(f"{name:12s} {tA:2d} {tB:2d} {M:6d} {N:6d} {K:6d} | {res[0]:11.2f} {fl / res[0] / 1e6:7.1f} {by / res[0] / 1e3:7.0f} | {res[1]:8.2f}")
