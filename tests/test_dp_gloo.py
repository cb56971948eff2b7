"""CPU, world_size 2, gloo: the data-parallel gradient exchange (slnlp_b200/dp.py).  Two ranks take
unequal numbers of valid labels; after sync_gradients both hold exactly the single-process gradient
and loss of the global batch (computed with the CPU oracle)."""
import os
import pickle
import socket
import subprocess
import sys
import textwrap

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = textwrap.dedent("""
    import os, pickle, sys
    sys.path.insert(0, {root!r}); sys.path.insert(0, os.path.join({root!r}, "sign-language-nlp_b200"))
    import importlib.util, torch, torch.distributed as dist
    # dp.py only needs torch: load it without importing the CUDA-only package __init__
    spec = importlib.util.spec_from_file_location("dp", os.path.join({root!r}, "sign-language-nlp_b200", "slnlp_b200", "dp.py"))
    dp = importlib.util.module_from_spec(spec); spec.loader.exec_module(dp)
    from oracle import port
    rank = int(os.environ["RANK"])
    dist.init_process_group("gloo", rank=rank, world_size=2)
    torch.manual_seed(0)
    ref = port.build_port("gru", 30, 9, 8, 8, 1, dropout=0.0)
    g = torch.Generator().manual_seed(1)
    X = torch.randint(2, 30, (12, 6), generator=g); L = torch.full((12,), 6); y = torch.randint(2, 9, (12,), generator=g)
    y[0] = 1; y[1] = 1; y[2] = 1           # ignored labels, all on rank 0's slice: counts 3 vs 6
    sl = slice(0, 6) if rank == 0 else slice(6, 12)
    loss = torch.nn.functional.cross_entropy(ref(X=X[sl], y=y[sl], lengths=L[sl]), y[sl], ignore_index=1)
    loss.backward()
    params = [p for p in ref.parameters() if p.grad is not None]
    gflat = torch.cat([p.grad.reshape(-1) for p in params])
    lc = torch.tensor([float(loss), float((y[sl] != 1).sum())])
    dp.sync_gradients(gflat, lc)
    pickle.dump((gflat, lc), open({out!r} + str(rank), "wb"))
    dist.destroy_process_group()
""")


def test_two_rank_gradient_sync_equals_single_process_global_batch(tmp_path):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port_no = s.getsockname()[1]
    out = str(tmp_path / "dp")
    script = tmp_path / "w.py"
    script.write_text(WORKER.format(root=ROOT, out=out))
    procs = [subprocess.Popen([sys.executable, str(script)],
                              env=dict(os.environ, RANK=str(r), WORLD_SIZE="2", MASTER_ADDR="127.0.0.1",
                                       MASTER_PORT=str(port_no), CUDA_VISIBLE_DEVICES=""),
                              stdout=subprocess.PIPE, stderr=subprocess.STDOUT) for r in range(2)]
    for p in procs:
        o, _ = p.communicate(timeout=240)
        assert p.returncode == 0, o.decode()
    (g0, lc0), (g1, lc1) = (pickle.load(open(out + str(r), "rb")) for r in range(2))
    assert torch.equal(g0, g1) and torch.equal(lc0, lc1)
    sys.path.insert(0, ROOT)
    from oracle import port
    torch.manual_seed(0)
    ref = port.build_port("gru", 30, 9, 8, 8, 1, dropout=0.0)
    g = torch.Generator().manual_seed(1)
    X = torch.randint(2, 30, (12, 6), generator=g)
    L = torch.full((12,), 6)
    y = torch.randint(2, 9, (12,), generator=g)
    y[0] = 1; y[1] = 1; y[2] = 1
    loss = torch.nn.functional.cross_entropy(ref(X=X, y=y, lengths=L), y, ignore_index=1)
    loss.backward()
    want = torch.cat([p.grad.reshape(-1) for p in ref.parameters() if p.grad is not None])
    assert float(lc0[1]) == 9.0
    assert abs(float(lc0[0]) - float(loss)) < 1e-6 * abs(float(loss))
    assert float((g0 - want).abs().max()) < 1e-6 * float(want.abs().max())


BUCKET_WORKER = textwrap.dedent("""
    import os, pickle, sys
    import importlib.util, torch, torch.distributed as dist
    spec = importlib.util.spec_from_file_location("dp", os.path.join({root!r}, "sign-language-nlp_b200", "slnlp_b200", "dp.py"))
    dp = importlib.util.module_from_spec(spec); spec.loader.exec_module(dp)
    rank = int(os.environ["RANK"])
    dist.init_process_group("gloo", rank=rank, world_size=2)
    g = torch.Generator().manual_seed(10 + rank)
    grad = torch.randn(1000, generator=g)             # this rank's MEAN gradient over its n_r valid labels
    lc = torch.tensor([1.5 + rank, 3.0 if rank == 0 else 6.0])
    sync = dp.BucketedGradSync()
    out = []
    for step in range(2):                               # the object is reused step after step
        gflat, l = grad.clone(), lc.clone()
        sync.begin(l)
        sync.ready(gflat, 700, 1000)                    # announced out of order, as backward finishes them
        sync.ready(gflat, 100, 400)
        sync.ready(gflat, 0, 100)                       # [400, 700) is never announced: finish() sends it
        sync.finish(gflat, l)
        out.append((gflat, l, sync.n_collectives))
    pickle.dump(out, open({out!r} + str(rank), "wb"))
    dist.destroy_process_group()
""")


def test_bucketed_overlapped_exchange_equals_the_count_weighted_global_gradient(tmp_path):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port_no = s.getsockname()[1]
    out = str(tmp_path / "bk")
    script = tmp_path / "w.py"
    script.write_text(BUCKET_WORKER.format(root=ROOT, out=out))
    procs = [subprocess.Popen([sys.executable, str(script)],
                              env=dict(os.environ, RANK=str(r), WORLD_SIZE="2", MASTER_ADDR="127.0.0.1",
                                       MASTER_PORT=str(port_no), CUDA_VISIBLE_DEVICES=""),
                              stdout=subprocess.PIPE, stderr=subprocess.STDOUT) for r in range(2)]
    for p in procs:
        o, _ = p.communicate(timeout=240)
        assert p.returncode == 0, o.decode()
    res = [pickle.load(open(out + str(r), "rb")) for r in range(2)]
    g0 = torch.randn(1000, generator=torch.Generator().manual_seed(10))
    g1 = torch.randn(1000, generator=torch.Generator().manual_seed(11))
    want = (3.0 * g0 + 6.0 * g1) / 9.0
    want_loss = (3.0 * 1.5 + 6.0 * 2.5) / 9.0
    for step in range(2):
        (a, la, na), (b, lb, nb) = res[0][step], res[1][step]
        assert torch.equal(a, b) and torch.equal(la, lb) and na == nb == 4
        assert float((a - want).abs().max()) < 1e-6
        assert abs(float(la[0]) - want_loss) < 1e-6 and float(la[1]) == 9.0
