"""Run one encoder-layer forward + BPTT call of the recurrent kernels (cfg1 shape) so that
ncu can capture them in isolation:  python profiles/prof_rnn_layer.py [fp32|bf16] [lstm|gru] [H]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "sign-language-nlp_b200"))
import torch
from slnlp_b200 import _lib

L = _lib
prec = 1 if (len(sys.argv) > 1 and sys.argv[1] == "bf16") else 0
mode = 1 if (len(sys.argv) > 2 and sys.argv[2] == "gru") else 0
H = int(sys.argv[3]) if len(sys.argv) > 3 else 128
T, B, G = 64, 50, (3 if mode else 4)
S = torch.cuda.current_stream().cuda_stream
torch.manual_seed(0)
w_hh = (torch.rand(2, G * H, H, device="cuda") * 2 - 1) / H ** 0.5
b_hh = (torch.rand(2 * G * H, device="cuda") * 2 - 1) / H ** 0.5
lengths = torch.full((B,), T, dtype=torch.int64, device="cuda")
dout, dfin = torch.randn(T, B, 2 * H, device="cuda"), torch.randn(2, B, H, device="cuda")
carry = torch.zeros(4, B, H, device="cuda")
for it in range(3):
    gates = torch.randn(T, B, 2, G, H, device="cuda")
    out, stash, hfin = torch.empty(T, B, 2 * H, device="cuda"), torch.empty(T, B, 2, H, device="cuda"), torch.empty(2, B, H, device="cuda")
    a, b, c = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    a.record()
    L.check(L.lib.slnlp_rnn_layer_fwd(mode, prec, T, B, H, 2, gates.data_ptr(), w_hh.data_ptr(), b_hh.data_ptr(),
                                      lengths.data_ptr(), None, None, out.data_ptr(), stash.data_ptr(), hfin.data_ptr(), S))
    b.record()
    L.check(L.lib.slnlp_rnn_layer_bwd(mode, prec, T, B, H, 2, gates.data_ptr(), stash.data_ptr(), out.data_ptr(),
                                      w_hh.data_ptr(), lengths.data_ptr(), None, None, dout.data_ptr(), dfin.data_ptr(),
                                      None, None, None, carry.data_ptr(), S))
    c.record()
    torch.cuda.synchronize()
    print(f"iter {it}: layer fwd {a.elapsed_time(b) * 1e3:.1f} us ({a.elapsed_time(b) * 1e3 / T:.2f} us/step), "
          f"bwd {b.elapsed_time(c) * 1e3:.1f} us ({b.elapsed_time(c) * 1e3 / T:.2f} us/step)")
