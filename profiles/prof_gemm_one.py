"""One tensor-core GEMM shape for ncu: python profiles/prof_gemm_one.py tA tB M N K [tf32|bf16]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "sign-language-nlp_b200"))
import torch
from slnlp_b200 import _lib as L
tA, tB, M, N, K = (int(x) for x in sys.argv[1:6])
A = torch.randn((K, M) if tA else (M, K), device="cuda")
Bm = torch.randn((N, K) if tB else (K, N), device="cuda")
C = torch.zeros(M, N, device="cuda")
bias = torch.randn(N, device="cuda")
ws = torch.empty(L.lib.slnlp_gemm_workspace_floats(), device="cuda")
fn = L.lib.slnlp_gemm_f32 if (len(sys.argv) > 6 and sys.argv[6] == "f32") else L.lib.slnlp_gemm_tf32
for _ in range(5):
    L.check(fn(tA, tB, M, N, K, A.data_ptr(), A.shape[1], Bm.data_ptr(), Bm.shape[1], C.data_ptr(), N,
                                  bias.data_ptr(), 0.0, ws.data_ptr(), ws.numel(), torch.cuda.current_stream().cuda_stream))
torch.cuda.synchronize()
print("ok")
