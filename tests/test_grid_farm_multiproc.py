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
                  "refit": hasattr(gs.best_estimator_, "coef_"), "skipped": gs.refit_skipped_}},
                open({out!r} + os.environ["RANK"], "wb"))
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
    assert r0["refit"] and not r1["refit"]                                   # refit on rank 0 only ...
    assert not r0["skipped"] and r1["skipped"]        # ... rank 1 holds the configured, unfitted estimator (no AttributeError)
    from sklearn.linear_model import LogisticRegression
    from sklearn.model_selection import GridSearchCV
    rng = np.random.RandomState(0)
    X = rng.randn(120, 5)
    y = (X[:, 0] + 0.5 * X[:, 1] > 0).astype(int)
    ref = GridSearchCV(LogisticRegression(max_iter=200), {"C": [0.01, 0.1, 1.0, 10.0]}, cv=3, scoring="accuracy").fit(X, y)
    assert np.allclose(ref.cv_results_["mean_test_score"], r0["mean"]) and ref.best_params_ == r0["best"]


FAIL_WORKER = textwrap.dedent("""
    import os, sys
    sys.path.insert(0, os.path.join({root!r}, "sign-language-nlp_b200"))
    import numpy as np
    from sklearn.base import BaseEstimator, ClassifierMixin
    from slnlp_b200.grid import GridSearchFarm

    class Flaky(ClassifierMixin, BaseEstimator):
        def __init__(self, C=1.0):
            self.C = C
        def fit(self, X, y):
            import time
            time.sleep(0.05)
            if self.C == 10.0:
                raise ValueError("boom at C=10")
            self.classes_ = np.unique(y)
            return self
        def predict(self, X):
            return np.zeros(len(X), dtype=int)

    rng = np.random.RandomState(0)
    X = rng.randn(60, 3); y = (X[:, 0] > 0).astype(int)
    try:
        GridSearchFarm(Flaky(), {{"C": [0.1, 1.0, 10.0, 100.0]}}, cv=3, scoring="accuracy", refit=False,
                       backend="torchrun", fits_per_gpu=int(os.environ["K"])).fit(X, y)
    except RuntimeError as e:
        assert "boom at C=10" in str(e), str(e)
        sys.exit(7)
    sys.exit(0)
""")


def test_a_failing_fit_reaches_every_rank_instead_of_a_store_timeout(tmp_path):
    """error_score='raise' under torchrun: the rank whose fit throws publishes an error record; the
    other rank re-raises the same error within seconds (it used to block on a key that never came).
    Also with two worker threads per rank (fits_per_gpu=2)."""
    import time
    for k in ("1", "2"):
        with socket.socket() as s:
            s.bind(("127.0.0.1", 0))
            port = s.getsockname()[1]
        script = tmp_path / f"f{k}.py"
        script.write_text(FAIL_WORKER.format(root=ROOT))
        t0 = time.perf_counter()
        procs = [subprocess.Popen([sys.executable, str(script)],
                                  env=dict(os.environ, RANK=str(r), WORLD_SIZE="2", MASTER_ADDR="127.0.0.1",
                                           MASTER_PORT=str(port), CUDA_VISIBLE_DEVICES="", K=k),
                                  stdout=subprocess.PIPE, stderr=subprocess.STDOUT) for r in range(2)]
        for p in procs:
            o, _ = p.communicate(timeout=120)
            assert p.returncode == 7, o.decode()
        assert time.perf_counter() - t0 < 90
