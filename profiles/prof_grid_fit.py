"""cProfile of single grid-search fits (host side): where the per-fit wall time goes.
    python profiles/prof_grid_fit.py [E] [H] [L]"""
import cProfile, io, os, pstats, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "sign-language-nlp_b200"))
import numpy as np
import torch
import helper as h
import model as dropin
from slnlp_b200.data import SeqDataset
from slnlp_b200.grid import _fit_and_score
from slnlp_b200.net import NeuralNetClassifier
from slnlp_b200 import callbacks as cbs
E, H, L = (int(x) for x in (sys.argv[1:4] if len(sys.argv) > 3 else (512, 256, 4)))
ds = SeqDataset.synthetic(n_seq=500, T=64, v_src=4098, v_tgt=52, ragged=True, seed=1)
net = NeuralNetClassifier(module=dropin.EncoderDecoderLSTMAttn, lr=0.01, max_epochs=2, batch_size=50, device="cuda:0", verbose=0,
                          precision="bf16", module__src_vocab=ds.vocab_X, module__tgt_vocab=ds.vocab_y, module__batch_first=True,
                          optimizer__momentum=0.9, optimizer__nesterov=False, criterion__ignore_index=1,
                          callbacks=[("gradient_clipping", cbs.GradientNormClipping(gradient_clip_value=0.5))])
y = ds.y().to_array(); Xall = ds.X()
scorer = h.build_scoring("neg_log_loss", ds.labels(), allow_multiple=False)
idx = np.arange(len(y)); train, test = idx[:400], idx[400:]
params = {"lr": 0.01, "module__embedding_size": E, "module__hidden_size": H, "module__num_layers": L, "module__dropout": 0.1}
_fit_and_score(net, params, Xall, y, train, test, scorer, seed=1)      # warm: library load, first-use attributes
torch.cuda.synchronize()
t0 = time.perf_counter()
pr = cProfile.Profile(); pr.enable()
for i in range(3):
    r = _fit_and_score(net, params, Xall, y, train, test, scorer, seed=2 + i)
pr.disable()
torch.cuda.synchronize()
print(f"E{E} H{H} L{L}: {(time.perf_counter() - t0) / 3 * 1e3:.1f} ms per fit (host wall), fit_time {r['fit_time'] * 1e3:.1f} ms, score_time {r['score_time'] * 1e3:.1f} ms")
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(45); print(s.getvalue()[:9000])
