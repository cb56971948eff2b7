"""Per-phase cycle table of the persistent recurrent kernel's step loop (cfg1 layer: LSTM / GRU, T 64,
B 50, H 128, both directions) and the layer time of each loop variant:

    python profiles/prof_persist_phases.py [lstm|gru]

For every forward variant (slnlp_debug_persist_config): (1) the un-instrumented layer timed as a CUDA-graph
replay with CUDA events (us per dependent timestep), (2) the instrumented instantiation's %clock table,
cycles per step averaged over the T steps, for three threads of CTA (0,0): the MMA issuer (thread 0),
thread 160 (warp 5, column group 1) and thread 511.  Variant bits: 1 = one accumulator per gate tile in the
forward kernel (instead of two partial ones),
$SLNLP_PERSIST_PSEQ = 1 | 2 | 4 | 16 forces the sequences per CTA (default: the smallest that fits one wave)."""
import ctypes
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "sign-language-nlp_b200"))
import torch
from slnlp_b200 import _lib

L = _lib
mode = 1 if (len(sys.argv) > 1 and sys.argv[1] == "gru") else 0
H, T, B, G = 128, 64, 50, (3 if mode else 4)
S = lambda: torch.cuda.current_stream().cuda_stream
torch.manual_seed(0)
w_hh = (torch.rand(2, G * H, H, device="cuda") * 2 - 1) / H ** 0.5
b_hh = (torch.rand(2 * G * H, device="cuda") * 2 - 1) / H ** 0.5
lengths = torch.full((B,), T, dtype=torch.int64, device="cuda")
dout, dfin = torch.randn(T, B, 2 * H, device="cuda"), torch.randn(2, B, H, device="cuda")
carry = torch.zeros(4, B, H, device="cuda")
gates0 = torch.randn(T, B, 2, G, H, device="cuda")
gates = gates0.clone()
out, stash, hfin = torch.empty(T, B, 2 * H, device="cuda"), torch.empty(T, B, 2, H, device="cuda"), torch.empty(2, B, H, device="cuda")
PHASES = ["mma issue", "stores+prefetch issue", "mma wait", "tmem ld", "gate math+h store", "proxy fence", "cta barrier", "-"]


def fwd():
    L.check(L.lib.slnlp_rnn_layer_fwd(mode, 1, T, B, H, 2, gates.data_ptr(), w_hh.data_ptr(), b_hh.data_ptr(),
                                      lengths.data_ptr(), None, None, out.data_ptr(), stash.data_ptr(), hfin.data_ptr(), S()))


def bwd():
    L.check(L.lib.slnlp_rnn_layer_bwd(mode, 1, T, B, H, 2, gates.data_ptr(), stash.data_ptr(), out.data_ptr(),
                                      w_hh.data_ptr(), lengths.data_ptr(), None, None, dout.data_ptr(), dfin.data_ptr(),
                                      None, None, None, carry.data_ptr(), S()))


def timed(fn, reps=20):
    for _ in range(3):
        fn()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        fn()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    a.record()
    for _ in range(reps):
        g.replay()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) * 1e3 / reps


def table(which, buf):
    print(f"  {'phase':24s} {'mma warp':>10s} {'epi A':>8s} {'epi B':>8s}   (cycles per step)")
    tot = [0.0, 0.0, 0.0]
    for ph in range(7):
        v = [buf[(which * 3 + th) * 8 + ph] / T for th in range(3)]
        tot = [a + b for a, b in zip(tot, v)]
        print(f"  {PHASES[ph]:24s} {v[0]:10.0f} {v[1]:8.0f} {v[2]:8.0f}")
    print(f"  {'sum':24s} {tot[0]:10.0f} {tot[1]:8.0f} {tot[2]:8.0f}")


buf = (ctypes.c_uint32 * 48)()
print(f"{'gru' if mode else 'lstm'} layer T {T} B {B} H {H}, both directions; sequences per CTA: {os.environ.get('SLNLP_PERSIST_PSEQ', 'auto')}")
for var in (0, 1):
    L.check(L.lib.slnlp_debug_persist_config(var, 0, None))
    gates.copy_(gates0)
    us_f = timed(fwd)
    gates.copy_(gates0); fwd()
    us_b = timed(bwd)
    print(f"variant {var}: fwd {us_f:.1f} us = {us_f / T:.3f} us/step, bwd {us_b:.1f} us = {us_b / T:.3f} us/step")
    L.check(L.lib.slnlp_debug_persist_config(var, 1, None))
    gates.copy_(gates0); fwd()
    L.check(L.lib.slnlp_debug_persist_config(-1, -1, buf))
    print(" forward, instrumented:")
    table(0, buf)
    bwd()
    L.check(L.lib.slnlp_debug_persist_config(-1, 0, buf))
    if var == 0:
        print(" backward, instrumented:")
        table(1, buf)
