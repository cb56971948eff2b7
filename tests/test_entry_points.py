"""CPU: the entry-point surface (args / YAML, helper.build_*, estimator parameter protocol,
grid-search scheduling and result table, callbacks) - everything around the kernels that can be
checked without a GPU."""
import os
import sys

import numpy as np
import pytest
import torch

import args as A
import helper as h
from slnlp_b200 import callbacks as cbs
from slnlp_b200.data import SeqDataset
from slnlp_b200.grid import GridSearchFarm, estimate_cost
from slnlp_b200.net import CVSplit, History, NeuralNetClassifier

REF_CFG = "/root/reference/config"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OWN_CFG = os.path.join(ROOT, "sign-language-nlp_b200", "config")


def _args(path, extra=()):
    return vars(A.load_args("t", A.ARGUMENTS, ["--config", path, *extra]))


@pytest.mark.skipif(not os.path.isdir(REF_CFG), reason="reference checkout not present")
@pytest.mark.parametrize("name,model,ncand", [("config-enc-dec-lstm-attn.yaml", "model.EncoderDecoderLSTMAttn", 162),
                                             ("config-enc-dec-gru-attn.yaml", "model.EncoderDecoderGRUAttn", 162),
                                             ("config-transformer.yaml", "model.Transformer", 324)])
def test_reference_yaml_loads_unchanged_and_builds_the_reference_grid(name, model, ncand):
    a = _args(os.path.join(REF_CFG, name))
    assert a["model"] == model and a["batch_size"] == 50 and a["max_epochs"] == 200 and a["cv"] == 5
    assert a["gradient_clipping"] == {"gradient_clip_value": 0.5}
    assert a["early_stopping"] == {"patience": 30, "threshold": 1e-4, "threshold_mode": "rel"}
    assert a["criterion_args"] == {} and a["mode"] == "grid"             # defaults commons would supply
    ds = SeqDataset.synthetic(n_seq=60, T=8, v_src=30, v_tgt=8)
    a["workdir"] = "/tmp/slnlp_test_workdir"
    callbacks, names = h.build_callbacks(dataset=ds, **a)
    assert names[:5] == ["checkpoint", "early_stopping", "gradient_clipping", "lr_scoring", "lr_scheduler"]
    assert len(names) == 5 + 2 * 5                                        # 5 metrics x (valid, train)
    gp = h.build_grid_params(callbacks_names=names, data=ds, **a)
    from sklearn.model_selection import ParameterGrid
    assert len(ParameterGrid(gp["param_grid"])) == ncand
    assert set(gp["param_grid"]) >= {"lr", "module__embedding_size", "module__hidden_size", "module__num_layers", "module__dropout"}
    assert gp["cv"] == 5 and gp["error_score"] == "raise" and gp["refit"] is True
    assert repr(gp["scoring"]) == "ScoringWrapper('neg_log_loss')"
    net_params = h.build_net_params(callbacks=callbacks, callbacks_names=names, device=torch.device("cuda"), dataset=ds, **a)
    assert net_params["criterion__ignore_index"] == 1 and net_params["optimizer__momentum"] == 0.9
    assert net_params["optimizer__nesterov"] is False and net_params["module__batch_first"] is True
    import model as dropin
    assert net_params["module"] is getattr(dropin, model.split(".")[1])
    net = NeuralNetClassifier(**net_params)
    assert net.get_params()["module__src_vocab"] is ds.vocab_X


def test_own_configs_load():
    for f in sorted(os.listdir(OWN_CFG)):
        a = _args(os.path.join(OWN_CFG, f), ["--precision", "bf16", "--gpus", "2"])
        assert a["precision"] == "bf16" and a["gpus"] == 2 and a["model"].startswith("model.")


def test_cli_overrides_yaml_and_required_keys():
    f = os.path.join(OWN_CFG, sorted(os.listdir(OWN_CFG))[0])
    a = _args(f, ["--batch_size", "7", "--grid_args", "{lr: [0.5], model_args: {hidden_size: [32]}}"])
    assert a["batch_size"] == 7 and a["grid_args"]["lr"] == [0.5]
    with pytest.raises(SystemExit):
        A.load_args("t", A.ARGUMENTS, ["--lr", "0.1"])                      # seed / max_epochs / ... missing


def test_prefix_args_and_collate():
    out = h.prefix_args("module", ensure_list=True, a=1, b={"c": [2, 3], "d": {"e": 4}})
    assert out == {"module__a": [1], "module__b__c": [2, 3], "module__b__d__e": [4]}
    assert h.prefix_args(None, x=1) == {"x": 1}
    batch, y = h.collate_data([(([5, 6, 1], 2, 3), 3), (([7, 1, 1], 1, 4), 4)])
    assert batch["X"].tolist() == [[5, 6, 1], [7, 1, 1]] and batch["lengths"].tolist() == [2, 1] and y.tolist() == [3, 4]
    assert batch["X"].dtype == torch.long


def test_estimator_parameter_protocol_and_clone():
    from sklearn.base import clone, is_classifier
    import model as dropin
    ds = SeqDataset.synthetic(n_seq=40, T=8, v_src=30, v_tgt=8)
    net = NeuralNetClassifier(module=dropin.EncoderDecoderLSTMAttn, lr=0.1, max_epochs=3, batch_size=10,
                              module__src_vocab=ds.vocab_X, module__tgt_vocab=ds.vocab_y, module__batch_first=True,
                              module__embedding_size=8, module__hidden_size=8, module__num_layers=1, module__dropout=0.0,
                              optimizer__momentum=0.9, criterion__ignore_index=1,
                              callbacks=[("gradient_clipping", cbs.GradientNormClipping(gradient_clip_value=0.5))])
    assert is_classifier(net)
    p = net.get_params()
    assert p["module__hidden_size"] == 8 and p["lr"] == 0.1 and p["optimizer__momentum"] == 0.9
    c = clone(net).set_params(lr=0.01, module__hidden_size=16)
    assert c.get_params()["module__hidden_size"] == 16 and net.get_params()["module__hidden_size"] == 8
    assert c.callbacks[0][1] is not net.callbacks[0][1]                     # deep-copied per fit
    with pytest.raises(ValueError):
        net.set_params(nonsense=1)
    with pytest.raises(TypeError):
        NeuralNetClassifier(module=dropin.EncoderDecoderLSTMAttn, bogus__x=1)
    with pytest.raises(RuntimeError, match="CUDA only"):
        NeuralNetClassifier(module=dropin.EncoderDecoderLSTMAttn, device="cpu").initialize()


def test_dataset_contract_and_sklearn_indexing():
    from sklearn.utils import _safe_indexing
    ds = SeqDataset.synthetic(n_seq=50, T=6, v_src=20, v_tgt=6, ragged=True)
    X, y = ds.X(), ds.y()
    assert len(X) == 50 and X.shape == (50,) and y.to_array().shape == (50,)
    sub = _safe_indexing(X, np.array([3, 7, 9]))
    assert len(sub) == 3 and sub[1][0] == ds.tokens[7].tolist() and sub[1][1] == int(ds.lengths[7]) and sub[1][2] == int(ds.labels_[7])
    (tok, ln), lab = ds[7]
    assert tok == ds.tokens[7].tolist() and lab == int(ds.labels_[7])
    test, train = ds.split(lengths=0.2, seed=1)
    assert len(test) == 10 and len(train) == 40
    assert ds.labels() == list(range(6)) and ds.truncated(5).tokens.shape[0] == 5
    bal = h.balance_dataset(ds, seed=1)
    cnt = np.bincount(bal.y().to_array(), minlength=6)
    orig = np.bincount(y.to_array(), minlength=6)
    assert cnt[2:].max() - cnt[2:].min() <= orig[2:].max() - orig[2:].min()  # flatter than before


def test_cvsplit_is_first_stratified_fold():
    from sklearn.model_selection import StratifiedKFold
    y = np.array([2, 3] * 25)
    tr, va = CVSplit(5)(50, y)
    tr2, va2 = next(iter(StratifiedKFold(5).split(np.arange(50), y)))
    assert np.array_equal(tr, tr2) and np.array_equal(va, va2)
    tr, va = CVSplit(5)(7, np.arange(7))                                    # singletons: plain KFold fallback
    assert len(va) == 2 and len(tr) == 5


def test_grid_schedule_is_longest_first_and_result_table_has_gridsearchcv_columns():
    grid = {"lr": [0.1, 0.01], "module__hidden_size": [128, 512], "module__num_layers": [2, 6], "module__embedding_size": [128]}
    import model as dropin
    gs = GridSearchFarm(NeuralNetClassifier(module=dropin.EncoderDecoderLSTMAttn), grid, cv=3, scoring="accuracy")
    y = np.array([2, 3, 4] * 10)
    cands, folds, tasks, order = gs._tasks(np.arange(30), y)
    assert len(cands) == 8 and len(folds) == 3 and len(tasks) == 24
    costs = [estimate_cost(cands[tasks[t][0]]) for t in order]
    assert costs == sorted(costs, reverse=True) and costs[0] > 20 * costs[-1]
    rng = np.random.RandomState(0)
    results = {t: {"score": float(rng.rand()), "fit_time": 1.0 + t, "score_time": 0.1} for t in range(len(tasks))}
    gs._collect(cands, folds, tasks, results)
    r = gs.cv_results_
    for col in ("mean_fit_time", "std_fit_time", "mean_score_time", "std_score_time", "params", "split0_test_score",
                "split2_test_score", "mean_test_score", "std_test_score", "rank_test_score", "param_lr",
                "param_module__hidden_size"):
        assert col in r, col
    assert r["rank_test_score"][gs.best_index_] == 1 and gs.best_params_ == cands[gs.best_index_]
    assert abs(gs.best_score_ - r["mean_test_score"].max()) < 1e-12
    import pandas as pd
    assert pd.DataFrame(r).shape[0] == 8


class _FakeNet:
    verbose = 0

    def __init__(self):
        self.history = History()
        self._stop_training = False


def test_early_stopping_and_history_semantics():
    net = _FakeNet()
    es = cbs.EarlyStopping(monitor="valid_loss", patience=3, threshold=1e-4, threshold_mode="rel", lower_is_better=True)
    es.on_train_begin(net)
    losses = [1.0, 0.9, 0.89999, 0.9, 0.91, 0.5]
    stopped_at = None
    for i, l in enumerate(losses):
        net.history.append({"epoch": i + 1, "valid_loss": l})
        es.on_epoch_end(net)
        if net._stop_training:
            stopped_at = i + 1
            break
    assert stopped_at == 5                      # 0.89999 is within the 1e-4 relative threshold of 0.9: a miss
    assert net.history[-1, "valid_loss"] == 0.91 and net.history[:, "valid_loss"][:2] == [1.0, 0.9]


def test_lr_scheduler_is_torch_reduce_on_plateau(tmp_path):
    net = _FakeNet()
    holder = torch.nn.Parameter(torch.zeros(1))
    net.optimizer_ = torch.optim.SGD([holder], lr=0.1, momentum=0.9)
    sch = cbs.LRScheduler(policy="ReduceLROnPlateau", monitor="valid_loss", factor=0.2, patience=2)
    sch.on_train_begin(net)
    for i in range(6):
        net.history.append({"epoch": i + 1, "valid_loss": 1.0})
        sch.on_epoch_end(net)
    assert abs(h.lr_score(net) - 0.1 * 0.2) < 1e-12
    cp = cbs.Checkpoint(monitor="valid_loss_best", dirname=str(tmp_path / "ck"), f_params=None, f_optimizer=None)
    net.history[-1]["valid_loss_best"] = True
    cp.on_epoch_end(net)
    assert (tmp_path / "ck" / "history.json").exists()


def test_grid_search_resumes_from_its_journal(tmp_path):
    """A search restarted with the same grid skips the (candidate, fold) fits its journal holds."""
    from sklearn.linear_model import LogisticRegression
    rng = np.random.RandomState(0)
    X = rng.randn(90, 4)
    y = (X[:, 0] > 0).astype(int)
    journal = str(tmp_path / "grid.jsonl")
    grid = {"C": [0.1, 1.0, 10.0]}
    a = GridSearchFarm(LogisticRegression(max_iter=200), grid, cv=3, scoring="accuracy", refit=False, backend="inline",
                       resume_file=journal).fit(X, y)
    assert a.n_resumed_ == 0 and sum(1 for _ in open(journal)) == 9
    lines = open(journal).read().splitlines()
    open(journal, "w").write("\n".join(lines[:5]) + "\n{\"params\": \"torn")          # interrupted run: 5 fits + a torn line
    b = GridSearchFarm(LogisticRegression(max_iter=200), grid, cv=3, scoring="accuracy", refit=False, backend="inline",
                       resume_file=journal).fit(X, y)
    assert b.n_resumed_ == 5 and np.allclose(a.cv_results_["mean_test_score"], b.cv_results_["mean_test_score"])
    c = GridSearchFarm(LogisticRegression(max_iter=200), grid, cv=3, scoring="accuracy", refit=False, backend="inline",
                       resume_file=journal).fit(X, y)
    assert c.n_resumed_ == 9 and c.best_params_ == a.best_params_


def test_grid_search_over_several_worker_processes_matches_the_inline_search():
    """procs_per_gpu = 2 (worker processes fed by one queue, here on the CPU with a scikit-learn estimator) returns the
    same table as the search run one fit at a time in this process."""
    from sklearn.linear_model import LogisticRegression
    rng = np.random.RandomState(1)
    X = rng.randn(120, 5)
    y = (X[:, 0] + 0.3 * X[:, 1] > 0).astype(int)
    grid = {"C": [0.01, 0.1, 1.0, 10.0]}
    a = GridSearchFarm(LogisticRegression(max_iter=200), grid, cv=3, scoring="accuracy", refit=False, backend="inline").fit(X, y)
    b = GridSearchFarm(LogisticRegression(max_iter=200), grid, cv=3, scoring="accuracy", refit=False, backend="inline",
                       procs_per_gpu=2, fits_per_gpu=2).fit(X, y)
    assert b.n_fits_ == 12 and np.allclose(a.cv_results_["mean_test_score"], b.cv_results_["mean_test_score"])
    assert b.best_params_ == a.best_params_ and b.worker_launches_ == 0        # no kernels on a CPU-only host


def test_product_package_never_imports_the_oracle():
    """The oracle is test infrastructure: nothing under the product package may import it."""
    import re
    pkg = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "sign-language-nlp_b200")
    for root, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(root, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, re.M), os.path.join(root, f)


def test_bench_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference` (the CPU arm the driver runs next to ours) needs no GPU: one JSON line
    with the base contract's keys, `impl: reference`, its own cpu_baseline and a zero-copy e2e object."""
    import json
    import subprocess
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1"],
                         capture_output=True, text=True, timeout=600, env=dict(os.environ, CUDA_VISIBLE_DEVICES=""))
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "e2e", "cpu_baseline", "impl"):
        assert k in line, k
    assert line["impl"] == "reference" and line["metric"] == "train_seq_per_s" and line["unit"] == "sequences/s"
    assert line["value"] > 0 and line["higher_is_better"] is True and line["vs_baseline"] is None
    assert "workload" in line["config"] and "EncoderDecoderLSTMAttn" in line["config"]["workload"]
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["d2h_bytes_per_step"] == 0
    assert line["e2e"]["value"] == line["value"] and line["cpu_baseline"]["value"] == line["value"]
    assert line["cpu_baseline"]["kind"] in ("port", "reference") and line["cpu_baseline"]["cores"] >= 1
    assert line["infer"]["value"] > 0
    # both arms print the SAME config object (a function of workload + command line only) ...
    import argparse
    import importlib.util
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(ROOT, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    assert line["config"] == bench.common_config(bench.WORKLOADS["cfg1"], argparse.Namespace(dp=False, no_flush=False), 1)
    # ... and the reference arm maps no product library (it times CPU code only)
    code = ("import sys, runpy\n"
            f"sys.argv = [{os.path.join(ROOT, 'bench.py')!r}, '--impl', 'reference', '--steps', '1', '--warmup', '1']\n"
            f"runpy.run_path({os.path.join(ROOT, 'bench.py')!r}, run_name='__main__')\n"
            "print('MAPPED', 'libslnlp' in open('/proc/self/maps').read(), any(m.startswith('slnlp_b200') for m in sys.modules))\n")
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=600,
                         env=dict(os.environ, CUDA_VISIBLE_DEVICES=""))
    assert out.returncode == 0, out.stderr[-2000:]
    assert out.stdout.strip().splitlines()[-1] == "MAPPED False False"


# ---------------------------------------------------------------- advisor findings, round 1
class _FakeModule:
    """Stands in for a CUDA module: initialize() needs only V_tgt / tgt_pad on a CPU-only host."""
    V_tgt, tgt_pad = 5, 1

    def __init__(self, **kw):
        self.kw = kw

    def to(self, dev):
        return self

    def parameters(self):
        return [torch.nn.Parameter(torch.zeros(1))]


def _cpu_initialize(net):
    """Run NeuralNetClassifier.initialize() up to the optimizer on a host without CUDA."""
    import unittest.mock as mock
    from slnlp_b200 import net as netmod
    with mock.patch.object(netmod, "OptimState", lambda *a, **k: None):
        return net.initialize()


def test_callbacks_params_are_routed_to_the_named_callback():
    es = cbs.EarlyStopping(patience=30)
    clip = cbs.GradientNormClipping(gradient_clip_value=0.5)
    net = NeuralNetClassifier(module=_FakeModule, device="cuda:0", verbose=0,
                              callbacks=[("early_stopping", es), ("gradient_clipping", clip)],
                              callbacks__early_stopping__patience=3, callbacks__gradient_clipping__gradient_clip_value=0.25)
    _cpu_initialize(net)
    routed = dict(net.callbacks_)
    assert routed["early_stopping"].patience == 3 and routed["gradient_clipping"].gradient_clip_value == 0.25
    assert es.patience == 30 and clip.gradient_clip_value == 0.5      # per-fit copies: the prototypes are untouched
    assert net.max_norm_ == 0.25                                      # the fused clip reads the routed value
    assert net.get_params()["callbacks__early_stopping__patience"] == 3
    with pytest.raises(ValueError, match="unknown callbacks"):
        _cpu_initialize(NeuralNetClassifier(module=_FakeModule, device="cuda:0", verbose=0, callbacks=[("early_stopping", es)],
                                            callbacks__nope__patience=3))
    with pytest.raises(ValueError, match="Invalid parameter"):
        _cpu_initialize(NeuralNetClassifier(module=_FakeModule, device="cuda:0", verbose=0, callbacks=[("early_stopping", es)],
                                            callbacks__early_stopping__patiense=3))


def test_set_params_keeps_the_trained_module_for_soft_parameters():
    net = NeuralNetClassifier(module=_FakeModule, device="cuda:0", verbose=0, lr=0.1)
    _cpu_initialize(net)
    assert net.initialized_
    net.set_params(lr=0.01, max_epochs=3)
    assert net.initialized_ and net.optimizer_.param_groups[0]["lr"] == 0.01     # skorch: lr does not re-initialise
    net.set_params(module__hidden_size=64)
    assert not net.initialized_                                                   # a structural parameter does


def test_load_dataset_refuses_a_missing_corpus_directory(tmp_path):
    with pytest.raises(FileNotFoundError):
        h.load_dataset(dataset_args={"dataset_dir": str(tmp_path / "nope"), "fields": ["handshape_dh"]})
    with pytest.raises(FileNotFoundError):
        h.load_dataset(dataset_args={})
    ds = h.load_dataset(dataset_args={"dataset_dir": str(tmp_path / "nope"), "synthetic": {"n_seq": 20, "T": 6, "v_src": 12, "v_tgt": 5}})
    assert len(ds) == 20
