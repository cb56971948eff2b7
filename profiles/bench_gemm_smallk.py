"""The K = batch (50) weight-gradient GEMMs of the decoder / generator on the TMA kernel:
python profiles/bench_gemm_smallk.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "sign-language-nlp_b200"))
import torch
from slnlp_b200 import _lib as L
ws = torch.empty(L.lib.slnlp_gemm_workspace_floats(), device="cuda")
for name, M, N, K, lda, beta in (("gen dW", 1026, 128, 50, 1028, 1.0), ("gen dW beta0", 1026, 128, 50, 1028, 0.0), ("dec dW_ih", 512, 384, 50, 512, 1.0),
                                 ("dec dW_hh", 512, 128, 50, 512, 1.0), ("K=64", 512, 128, 64, 512, 1.0), ("K=96", 512, 128, 96, 512, 1.0), ("K=128", 512, 128, 128, 512, 1.0)):
    A = torch.randn(K, lda, device="cuda")
    B = torch.randn(K, N, device="cuda")
    C = torch.zeros(M, N, device="cuda")
    for fn, tag in ((L.lib.slnlp_gemm_tf32, "tma"), (L.lib.slnlp_gemm_f32, "f32")):
        call = lambda: L.check(fn(1, 0, M, N, K, A.data_ptr(), lda, B.data_ptr(), N, C.data_ptr(), N, None, beta, ws.data_ptr(), ws.numel(),
                                  torch.cuda.current_stream().cuda_stream))
        for _ in range(3):
            call()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            for _ in range(10):
                call()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); a.record()
        for _ in range(5):
            g.replay()
        b.record(); torch.cuda.synchronize()
        print(f"{name:14s} M{M} N{N} K{K} {tag}: {a.elapsed_time(b) * 1e3 / 50:7.2f} us")
