"""Weight-gradient GEMMs of one cfg1 train step (C += A^T B over K = T*B rows, split-K with in-place red.global.add)
in isolation, as a function of the cap on the number of K-slices ($SLNLP_SPLITK_MAX):
    python profiles/bench_gemm_dw.py
CUDA-graph replay of 20 calls, L2-warm; us per call."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "sign-language-nlp_b200"))
import torch
from slnlp_b200 import _lib as L

B, T, E, H, G = 50, 64, 128, 128, 4
R, GH = B * T, G * H
shapes = [("dW_ih l0", 2 * GH, E, R), ("dW_ih l1", 2 * GH, 2 * H, R), ("dW_hh", GH, H, R - B), ("key dW", H, 2 * H, R)]
ws = torch.empty(L.lib.slnlp_gemm_workspace_floats(), device="cuda")
caps = [0, 1, 2, 3, 4, 6, 9, 12]
print(f"{'gemm':10s} {'M':>5} {'N':>5} {'K':>5} | " + " ".join(f"{('cap ' + str(c)) if c else 'default':>8}" for c in caps))
for name, M, N, K in shapes:
    A = torch.randn(K, M, device="cuda") * 0.01
    Bm = torch.randn(K, N, device="cuda") * 0.01
    C = torch.zeros(M, N, device="cuda")
    row = []
    for cap in caps:
        if cap:
            os.environ["SLNLP_SPLITK_MAX"] = str(cap)
        else:
            os.environ.pop("SLNLP_SPLITK_MAX", None)
        call = lambda: L.check(L.lib.slnlp_gemm_tf32(1, 0, M, N, K, A.data_ptr(), M, Bm.data_ptr(), N, C.data_ptr(), N, None, 1.0,
                                                     ws.data_ptr(), ws.numel(), torch.cuda.current_stream().cuda_stream))
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
        row.append(a.elapsed_time(b) * 1e3 / 100)
    print(f"{name:10s} {M:5d} {N:5d} {K:5d} | " + " ".join(f"{t:8.2f}" for t in row))

# (An 8-stage operand ring for the unsplit K = 1024 d(input) products - 192 KB, one CTA per SM - was measured here too:
# 10.2 -> 10.4 us for [3200 x 1024] x [1024 x 256], no gain.  100 CTAs x 32 k-steps x 24 KB = 79 MB in ~8 us is ~10 TB/s of
# L2 reads: these products are bound by re-reading the operands from L2 (A 4x, B 25x), not by load latency; halving the bytes
# (bf16 operands) or sharing a tile across a cluster (TMA multicast) is what would move them.)
