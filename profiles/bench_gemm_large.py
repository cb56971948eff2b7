"""TMA tf32 GEMM on the large-batch (cfg4) shapes: python profiles/bench_gemm_large.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "sign-language-nlp_b200"))
import torch
from slnlp_b200 import _lib as L
ws = torch.empty(L.lib.slnlp_gemm_workspace_floats(), device="cuda")
shapes = [("inproj l1 (B512)", 0, 1, 32768, 4096, 1024), ("inproj l0 (B512)", 0, 1, 32768, 4096, 1024),
          ("dx (B512)", 0, 0, 32768, 1024, 4096), ("dW_ih (B512)", 1, 0, 4096, 1024, 32768),
          ("inproj (B4096)", 0, 1, 262144, 4096, 1024), ("square 8192", 0, 1, 8192, 8192, 8192)]
for name, tA, tB, M, N, K in shapes:
    A = torch.randn((K, M) if tA else (M, K), device="cuda")
    B = torch.randn((N, K) if tB else (K, N), device="cuda")
    C = torch.empty(M, N, device="cuda")
    call = lambda: L.check(L.lib.slnlp_gemm_tf32(tA, tB, M, N, K, A.data_ptr(), A.shape[1], B.data_ptr(), B.shape[1], C.data_ptr(), N,
                                                 None, 0.0, ws.data_ptr(), ws.numel(), torch.cuda.current_stream().cuda_stream))
    for _ in range(2):
        call()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); a.record()
    for _ in range(5):
        call()
    b.record(); torch.cuda.synchronize()
    t = a.elapsed_time(b) / 5 * 1e-3
    print(f"{name:20s} tA{tA} tB{tB} {M:7d} {N:5d} {K:6d}  {t * 1e3:8.3f} ms  {2.0 * M * N * K / t / 1e12:7.1f} TFLOP/s")
    del A, B, C
