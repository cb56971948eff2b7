"""One bidirectional LSTM layer at the data-parallel shape (T steps, B 4096, H 512) through slnlp_rnn_layer_fwd/bwd_bf16:
us per timestep of the forward and backward step kernels (CUDA events).   python profiles/prof_step_pair.py [T] [B] [H] [reps]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "sign-language-nlp_b200"))
import torch
from slnlp_b200 import _lib as L
T = int(sys.argv[1]) if len(sys.argv) > 1 else 16
B = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
H = int(sys.argv[3]) if len(sys.argv) > 3 else 512
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 5
G = 4
S = lambda: torch.cuda.current_stream().cuda_stream
g = torch.Generator().manual_seed(1)
f = lambda *s: (torch.randn(*s, generator=g) * 0.5).cuda()
bf = lambda *s: torch.empty(*s, device="cuda", dtype=torch.bfloat16)
gates0 = f(T, B, 2, G, H)
w_hh, b_hh = f(2, G * H, H) * 0.1, f(2 * G * H) * 0.1
lengths = torch.randint(T // 2, T + 1, (B,), generator=g).cuda()
out, stash, hfin = torch.empty(T, B, 2 * H, device="cuda"), torch.empty(T, B, 2, H, device="cuda"), torch.empty(2, B, H, device="cuda")
w_bf, wT_bf, out_bf, dg_bf = bf(2, G * H, H), bf(2, H, G * H), bf(T, B, 2 * H), bf(T, B, 2, G, H)
L.check(L.lib.slnlp_cast_bf16(w_hh.data_ptr(), H, w_bf.data_ptr(), H, 2 * G * H, H, 0, S()))
for d in range(2):
    L.check(L.lib.slnlp_cast_bf16(w_hh[d].data_ptr(), H, wT_bf[d].data_ptr(), G * H, G * H, H, 1, S()))
dout, dfin, carry = f(T, B, 2 * H), f(2, B, H), torch.zeros(4, B, H, device="cuda")
gates = gates0.clone()
GB = L.lib.slnlp_rnn_bf16_pair_supported(0, T, B, H, 2) and os.environ.get("GATES_BF16", "0") == "1"   # bf16 gate stash in dg_bf
def fwd():
    L.check(L.lib.slnlp_rnn_layer_fwd_bf16_ex(0, T, B, H, 2, gates.data_ptr(), w_bf.data_ptr(), b_hh.data_ptr(), lengths.data_ptr(),
                                              out.data_ptr(), out_bf.data_ptr(), stash.data_ptr(), hfin.data_ptr(),
                                              dg_bf.data_ptr() if GB else None, S()))
def bwd(wf):
    L.check(L.lib.slnlp_rnn_layer_bwd_bf16(0, T, B, H, 2, gates.data_ptr(), dg_bf.data_ptr(), stash.data_ptr(), out.data_ptr(),
                                           wT_bf.data_ptr(), lengths.data_ptr(), dout.data_ptr(), dfin.data_ptr(), None,
                                           carry.data_ptr(), wf, None, 1.0, 1 if GB else 0, S()))
def timeit(fn):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); a.record(); fn(); b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) * 1e3 / T
tf, tb, tb1 = [], [], []
for i in range(reps):
    gates.copy_(gates0)
    tf.append(timeit(fwd))
    keep, keepb = gates.clone(), dg_bf.clone()
    tb.append(timeit(lambda: bwd(0)))
    gates.copy_(keep); dg_bf.copy_(keepb)
    tb1.append(timeit(lambda: bwd(1)))
el = B * H * 2
print(("bf16 gate stash: " if GB else "") + f"T {T} B {B} H {H}: fwd {min(tf):.1f} us/step ({el * 46 / min(tf) / 1e3:.0f} GB/s of the 46 B/element), "
      f"bwd {min(tb):.1f} us/step bf16-only ({el * 44 / min(tb) / 1e3:.0f} GB/s of 44 B/element), bwd + fp32 dG {min(tb1):.1f} us/step")
