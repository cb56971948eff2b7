"""CTA-pair bf16 GEMM (gemm_pair.cu) on the large-batch (cfg4) shapes, next to the TMA tf32 kernel:
python profiles/bench_gemm_pair.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "sign-language-nlp_b200"))
import torch
from slnlp_b200 import _lib as L
ws = torch.empty(L.lib.slnlp_gemm_workspace_floats(), device="cuda")
S = lambda: torch.cuda.current_stream().cuda_stream
shapes = [("inproj (B4096)", 0, 1, 262144, 4096, 1024), ("dx (B4096)", 0, 0, 262144, 1024, 4096),
          ("dW_ih (B4096)", 1, 0, 4096, 1024, 262144), ("dW_hh (B4096)", 1, 0, 2048, 512, 258048),
          ("inproj (B512)", 0, 1, 32768, 4096, 1024), ("square 8192", 0, 1, 8192, 8192, 8192),
          ("cfg3 qkv", 0, 1, 3200, 1536, 512)]
only = sys.argv[1:] 
def timeit(call, n=5):
    for _ in range(2):
        call()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); a.record()
    for _ in range(n):
        call()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n * 1e-3
for name, tA, tB, M, N, K in shapes:
    if only and not any(o in name for o in only):
        continue
    A = torch.randn((K, M) if tA else (M, K), device="cuda")
    B = torch.randn((N, K) if tB else (K, N), device="cuda")
    Ab, Bb = A.to(torch.bfloat16), B.to(torch.bfloat16)
    C = torch.empty(M, N, device="cuda")
    t32 = timeit(lambda: L.check(L.lib.slnlp_gemm_tf32(tA, tB, M, N, K, A.data_ptr(), A.shape[1], B.data_ptr(), B.shape[1], C.data_ptr(), N,
                                                        None, 0.0, ws.data_ptr(), ws.numel(), S())))
    C32 = C.clone()
    tb = timeit(lambda: L.check(L.lib.slnlp_gemm_bf16(tA, tB, M, N, K, Ab.data_ptr(), Ab.shape[1], Bb.data_ptr(), Bb.shape[1], C.data_ptr(), N,
                                                       None, 0.0, S())))
    err = float((C - C32).abs().max() / C32.abs().max())
    tc = timeit(lambda: L.check(L.lib.slnlp_cast_bf16(A.data_ptr(), A.shape[1], Ab.data_ptr(), A.shape[1], A.shape[0], A.shape[1], 0, S())))
    fl = 2.0 * M * N * K
    print(f"{name:16s} tA{tA} tB{tB} {M:7d} {N:5d} {K:6d}  tf32 {t32 * 1e3:8.3f} ms {fl / t32 / 1e12:7.1f} TF/s | pair bf16 {tb * 1e3:8.3f} ms "
          f"{fl / tb / 1e12:7.1f} TF/s | cast A {tc * 1e3:6.3f} ms {A.numel() * 6 / tc / 1e9:6.0f} GB/s | max diff vs tf32 {err:.2e}", flush=True)
    del A, B, Ab, Bb, C, C32
