"""CPU, world_size 2: the torchrun backend of the grid-search farm.  Two processes claim
(candidate, fold) fits from the store counter with no collective; both end with the full,
identical result table, equal to sklearn's own GridSearchCV on the same estimator."""
import os
import pickle
import socket
import subprocess
import sys
import textwrap

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = textwrap.dedent("""
    import os, pickle, sys
    sys.path.insert(0, os.path.join({root!r}, "sign-language-nlp_b200"))
    import numpy as np
    from sklearn.linear_model import LogisticRegression
    from slnlp_b200.grid import GridSearchFarm
    rng = np.random.RandomState(0)
    X = rng.randn(120, 5); y = (X[:, 0] + 0.5 * X[:, 1] > 0).astype(int)
    gs = GridSearchFarm(LogisticRegression(max_iter=200), {{"C": [0.01, 0.1, 1.0, 10.0]}}, cv=3, scoring="accuracy",
                        refit=True, backend="torchrun")
    gs.fit(X, y)
    owners = sorted(r["gpu"] for r in gs.fit_results_.values())
    pickle.dump({{"mean": gs.cv_results_["mean_test_score"], "best": gs.best_params_, "owners": owners,
                  "refit": hasattr(gs, "best_estimator_")}}, open({out!r} + os.environ["RANK"], "wb"))
""")


def test_two_ranks_share_the_grid_without_collectives(tmp_path):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    out = str(tmp_path / "res")
    script = tmp_path / "w.py"
    script.write_text(WORKER.format(root=ROOT, out=out))
    procs = []
    for rank in range(2):
        env = dict(os.environ, RANK=str(rank), WORLD_SIZE="2", MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port),
                   CUDA_VISIBLE_DEVICES="")
        procs.append(subprocess.Popen([sys.executable, str(script)], env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT))
    for p in procs:
        o, _ = p.communicate(timeout=240)
        assert p.returncode == 0, o.decode()
    r0, r1 = (pickle.load(open(out + str(r), "rb")) for r in range(2))
    assert np.array_equal(r0["mean"], r1["mean"]) and r0["best"] == r1["best"]
    assert len(r0["owners"]) == 12 and set(r0["owners"]) <= {0, 1}          # every fit ran exactly once
    assert r0["refit"] and not r1["refit"]                                   # refit on rank 0 only
    from sklearn.linear_model import LogisticRegression
    from sklearn.model_selection import GridSearchCV
    rng = np.random.RandomState(0)
    X = rng.randn(120, 5)
    y = (X[:, 0] + 0.5 * X[:, 1] > 0).astype(int)
    ref = GridSearchCV(LogisticRegression(max_iter=200), {"C": [0.01, 0.1, 1.0, 10.0]}, cv=3, scoring="accuracy").fit(X, y)
    assert np.allclose(ref.cv_results_["mean_test_score"], r0["mean"]) and ref.best_params_ == r0["best"]
