"""One flash-style MHA forward + backward at the cfg3 shape for ncu: python profiles/prof_mha.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "sign-language-nlp_b200"))
import torch
from slnlp_b200 import _lib as L
B, S, nh, dh = 50, 64, 8, 64
E = nh * dh
x = torch.randn(B, S, 3 * E, device="cuda")
tok = torch.full((B, S), 5, dtype=torch.int64, device="cuda")
o, lse = torch.empty(B, S, E, device="cuda"), torch.empty(B, nh, S, device="cuda")
do, dx, dvec = torch.randn(B, S, E, device="cuda"), torch.empty_like(x), torch.empty(B, nh, S, device="cuda")
q, d = x.data_ptr(), dx.data_ptr()
st = torch.cuda.current_stream().cuda_stream
for _ in range(3):
    a, b, c = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    a.record()
    L.check(L.lib.slnlp_mha_fwd(q, 3 * E, q + 4 * E, 3 * E, q + 8 * E, 3 * E, o.data_ptr(), E, lse.data_ptr(), B, S, S, nh, dh, 1,
                                tok.data_ptr(), 1, 0.0, None, 0, st))
    b.record()
    L.check(L.lib.slnlp_mha_bwd(q, 3 * E, q + 4 * E, 3 * E, q + 8 * E, 3 * E, o.data_ptr(), do.data_ptr(), E, lse.data_ptr(),
                                dvec.data_ptr(), d, d + 4 * E, d + 8 * E, B, S, S, nh, dh, 1, tok.data_ptr(), 1, 0.0, None, 0, st))
    c.record()
    torch.cuda.synchronize()
    print(f"mha fwd {a.elapsed_time(b) * 1e3:.1f} us, bwd {b.elapsed_time(c) * 1e3:.1f} us")
