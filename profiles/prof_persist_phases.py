"""Debug build only (SLNLP_NVCC_FLAGS=-DSLNLP_PERSIST_TIMING): per-phase clock64 cycles of the
persistent forward kernel, CTA (0,0), averaged over steps."""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "sign-language-nlp_b200"))
import torch
from slnlp_b200 import _lib as L
mode = 1 if (len(sys.argv) > 1 and sys.argv[1] == "gru") else 0
T, B, H, G = 64, 50, 128, (3 if mode else 4)
S = torch.cuda.current_stream().cuda_stream
w_hh = (torch.rand(2, G * H, H, device="cuda") * 2 - 1) / H ** 0.5
b_hh = torch.zeros(2 * G * H, device="cuda")
lengths = torch.full((B,), T, dtype=torch.int64, device="cuda")
out8 = (ctypes.c_ulonglong * 8)()
for it in range(3):
    gates = torch.randn(T, B, 2, G, H, device="cuda")
    out, stash, hfin = torch.empty(T, B, 2 * H, device="cuda"), torch.empty(T, B, 2, H, device="cuda"), torch.empty(2, B, H, device="cuda")
    torch.cuda.synchronize()
    L.lib.slnlp_debug_persist_clocks(out8, 1)
    L.check(L.lib.slnlp_rnn_layer_fwd(mode, 1, T, B, H, 2, gates.data_ptr(), w_hh.data_ptr(), b_hh.data_ptr(),
                                      lengths.data_ptr(), None, None, out.data_ptr(), stash.data_ptr(), hfin.data_ptr(), S))
    torch.cuda.synchronize()
    L.lib.slnlp_debug_persist_clocks(out8, 0)
    n = max(1, out8[6])
    names = ["mma_issue(t0)", "wait_mma", "tmem_ld", "compute+ld+st", "fence+sync", "step_total"]
    print("  ".join(f"{nm}={out8[i] / n:.0f}" for i, nm in enumerate(names)), f"steps={n}")
