"""GPU: the skorch-shaped estimator and the entry points.  The fit loop (inner stratified
80/20 split, in-order batches with a ragged tail, CE on the log-probs, norm clip 0.5,
SGD-momentum, epoch losses) is compared with the same loop written around the torch.nn port of
the reference; sklearn's own GridSearchCV drives the estimator and agrees with the farm."""
import json
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from helpers import FP32_RTOL  # noqa: E402


def _net(ds, kind="lstm", epochs=2, **kw):
    import model as dropin
    from slnlp_b200 import callbacks as cbs
    from slnlp_b200.net import NeuralNetClassifier
    cls = {"lstm": dropin.EncoderDecoderLSTMAttn, "gru": dropin.EncoderDecoderGRUAttn, "transformer": dropin.Transformer}[kind]
    extra = {"module__num_heads": 2} if kind == "transformer" else {}
    params = dict(module=cls, lr=0.05, max_epochs=epochs, batch_size=16, device="cuda", verbose=0,
                  module__src_vocab=ds.vocab_X, module__tgt_vocab=ds.vocab_y, module__batch_first=True,
                  module__embedding_size=16, module__hidden_size=16, module__num_layers=2, module__dropout=0.0,
                  optimizer__momentum=0.9, optimizer__nesterov=False, criterion__ignore_index=1,
                  callbacks=[("gradient_clipping", cbs.GradientNormClipping(gradient_clip_value=0.5))], **extra)
    params.update(kw)
    return NeuralNetClassifier(**params)


def _port_loop(kind, ds, sd, epochs, lr, bs, heads=2):
    """The reference's training loop around the torch.nn port, on CPU."""
    from oracle import port
    from slnlp_b200.net import CVSplit
    ref = port.build_port(kind, len(ds.vocab_X), len(ds.vocab_y), 16, 16, 2, dropout=0.0, num_heads=heads)
    ref.load_state_dict(sd)
    opt = torch.optim.SGD(ref.parameters(), lr=lr, momentum=0.9)
    X, L, y = ds.tokens, ds.lengths, ds.labels_
    tr, va = CVSplit(5)(len(X), y.numpy())
    tr, va = torch.from_numpy(tr), torch.from_numpy(va)
    hist = []
    for _ in range(epochs):
        tot = 0.0
        for j in range(0, len(tr), bs):
            idx = tr[j:j + bs]
            loss = port.reference_train_step(ref, opt, X[idx], y[idx], L[idx])
            tot += float(loss) * len(idx)
        ref.eval()
        with torch.no_grad():
            vt = 0.0
            for j in range(0, len(va), bs):
                idx = va[j:j + bs]
                vt += float(torch.nn.functional.cross_entropy(ref(X=X[idx], y=y[idx], lengths=L[idx]), y[idx], ignore_index=1)) * len(idx)
        hist.append((tot / len(tr), vt / len(va)))
    ref.eval()
    with torch.no_grad():
        proba = torch.softmax(ref(X=X, y=y, lengths=L), -1)
    return hist, proba


@pytest.mark.parametrize("kind", ["lstm", "gru", "transformer"])
def test_fit_loop_matches_reference_loop(kind):
    from slnlp_b200.data import SeqDataset
    ds = SeqDataset.synthetic(n_seq=90, T=12, v_src=40, v_tgt=7, ragged=True, seed=5)
    torch.manual_seed(3)
    net = _net(ds, kind).initialize()
    sd = {k: v.detach().cpu().clone() for k, v in net.module_.state_dict().items()}
    net.warm_start = True
    net.fit(ds.X(), ds.y().to_array())
    want, want_proba = _port_loop(kind, ds, sd, 2, 0.05, 16)
    assert len(net.history) == 2
    for row, (tl, vl) in zip(net.history, want):
        assert abs(row["train_loss"] - tl) < 2e-5 * abs(tl)
        assert abs(row["valid_loss"] - vl) < 2e-5 * abs(vl)
    got = net.predict_proba(ds.X())
    assert got.shape == (90, 7) and np.allclose(got.sum(1), 1, atol=1e-5)
    assert np.abs(got - want_proba.numpy()).max() < 5e-5
    assert np.array_equal(net.predict(ds.X()), want_proba.argmax(1).numpy())
    assert 0.0 <= net.score(ds.X(), ds.y().to_array()) <= 1.0


def test_callbacks_history_checkpoint_and_lr_schedule(tmp_path):
    import helper as h
    from slnlp_b200.data import SeqDataset
    ds = SeqDataset.synthetic(n_seq=80, T=10, v_src=30, v_tgt=6, ragged=True, seed=2)
    callbacks, names = h.build_callbacks(mode="grid", workdir=str(tmp_path), scoring=["neg_log_loss", "accuracy", "f1_weighted"],
                                         dataset=ds, early_stopping=dict(patience=2, threshold=0.5, threshold_mode="rel"),
                                         gradient_clipping=dict(gradient_clip_value=0.5),
                                         lr_scheduler=dict(policy="ReduceLROnPlateau", factor=0.2, patience=0, threshold=0.9))
    net = _net(ds, "lstm", epochs=10, callbacks=callbacks)
    net.fit(ds.X(), ds.y().to_array())
    row = net.history[0]
    for k in ("train_loss", "valid_loss", "valid_loss_best", "lr", "valid_neg_log_loss", "train_neg_log_loss",
              "valid_accuracy", "train_accuracy", "valid_f1_weighted", "dur"):
        assert k in row, k
    assert abs(row["valid_neg_log_loss"] + row["valid_loss"]) < 1e-4        # the same quantity, two routes
    assert len(net.history) == 3            # threshold 0.5 rel: epochs 2, 3 cannot "improve" -> stop at patience 2
    assert net.history[-1]["lr"] < net.history[0]["lr"]                     # ReduceLROnPlateau (patience 0) cut it
    assert os.path.exists(tmp_path / "params.pt") and os.path.exists(tmp_path / "history.json")
    sd = torch.load(tmp_path / "params.pt")
    assert "model.encoder.rnn.weight_hh_l0_reverse" in sd                    # reference state_dict names


def test_sklearn_gridsearchcv_drives_the_estimator_and_agrees_with_the_farm():
    import helper as h
    from sklearn.model_selection import GridSearchCV
    from slnlp_b200.data import SeqDataset
    from slnlp_b200.grid import GridSearchFarm
    ds = SeqDataset.synthetic(n_seq=60, T=8, v_src=30, v_tgt=5, ragged=True, seed=4)
    grid = {"lr": [0.1, 0.01], "module__hidden_size": [8, 16]}
    scoring = h.build_scoring("neg_log_loss", ds.labels(), allow_multiple=False)
    y = ds.y().to_array()
    torch.manual_seed(0)
    a = GridSearchCV(_net(ds, "gru", epochs=2), grid, cv=2, scoring=scoring, refit=False, error_score="raise").fit(ds.X(), y)
    torch.manual_seed(0)
    b = GridSearchFarm(_net(ds, "gru", epochs=2), grid, cv=2, scoring=scoring, refit=True, backend="inline").fit(ds.X(), y)
    # same candidates / folds; each fit seeds its own weights from the global stream, so compare structure + ranges
    assert [p for p in a.cv_results_["params"]] == b.cv_results_["params"]
    assert a.cv_results_["mean_test_score"].shape == b.cv_results_["mean_test_score"].shape == (4,)
    assert np.isfinite(b.cv_results_["mean_test_score"]).all() and (b.cv_results_["mean_test_score"] < 0).all()
    assert b.best_estimator_.predict(ds.X()).shape == (60,)


def test_stock_autograd_route_for_other_optimizers():
    from slnlp_b200.data import SeqDataset
    ds = SeqDataset.synthetic(n_seq=40, T=8, v_src=30, v_tgt=5, ragged=True, seed=4)
    net = _net(ds, "lstm", epochs=3, optimizer=torch.optim.Adam, lr=0.01)
    net._kwargs_keys = [k for k in net._kwargs_keys if not k.startswith("optimizer__")]
    net.fit(ds.X(), ds.y().to_array())
    assert not net.fused_ and net.history[-1]["train_loss"] < net.history[0]["train_loss"]


def test_main_run_end_to_end(tmp_path):
    import args as A
    import main
    cfg = os.path.join(os.path.dirname(os.path.abspath(main.__file__)), "config", "b200-lstm-attn.yaml")
    a = vars(A.load_args("t", A.ARGUMENTS, ["--config", cfg, "--workdir", str(tmp_path / "run"), "--max_epochs", "2",
                                            "--cv", "2", "--gpus", "1", "--precision", "fp32", "--verbose", "0",
                                            "--dataset_args", "{synthetic: {n_seq: 120, T: 16, v_src: 60, v_tgt: 8}}",
                                            "--grid_args", "{lr: [0.1, 0.01], model_args: {embedding_size: [16], hidden_size: [16], num_layers: [1], dropout: [0.0]}}"]))
    a["workdir"] = main.h.format_dir(a["workdir"], **a)
    main.h.dump_args(a)
    out = main.run(a)
    assert set(out) >= {"test_accuracy", "test_neg_log_loss", "test_f1_weighted"}
    wd = a["workdir"]
    for f in ("config.yaml", "grid_search_grid_params.csv", "grid_search_results.csv", "grid_search_output.json",
              "test_output.json", "test_profile.json"):
        assert os.path.exists(os.path.join(wd, f)), f
    go = json.load(open(os.path.join(wd, "grid_search_output.json")))
    assert go["n_fits"] == 4 and go["best_params"]["lr"] in (0.1, 0.01) and go["fits_per_hour"] > 0
    prof = json.load(open(os.path.join(wd, "test_profile.json")))
    assert prof["kernel_launches"] > 0


def test_packed_grid_fits_reproduce_the_one_at_a_time_search():
    """fits_per_gpu = 3 (worker threads on private streams, one captured step graph each, fits seeded from
    (candidate, fold)) against the same search one fit at a time, fp32 path with dropout: per-candidate scores agree
    (up to the order of atomic adds in the embedding / bias gradients), twice in a row."""
    import helper as h
    from slnlp_b200.data import SeqDataset
    from slnlp_b200.grid import GridSearchFarm
    ds = SeqDataset.synthetic(n_seq=150, T=12, v_src=40, v_tgt=6, ragged=True, seed=4)
    grid = {"lr": [0.1, 0.01], "module__hidden_size": [16, 32], "module__num_layers": [1, 2], "module__dropout": [0.2]}
    scoring = h.build_scoring("neg_log_loss", ds.labels(), allow_multiple=False)
    y = ds.y().to_array()
    runs = []
    for k in (1, 3, 3):
        gs = GridSearchFarm(_net(ds, "lstm", epochs=2), grid, cv=3, scoring=scoring, refit=False, backend="inline",
                            fits_per_gpu=k).fit(ds.X(), y)
        assert gs.n_fits_ == 24
        runs.append(gs.cv_results_["mean_test_score"])
    assert np.abs(runs[1] - runs[0]).max() < 1e-5 and np.abs(runs[2] - runs[1]).max() < 1e-5
    assert len(set(np.round(runs[0], 6))) > 4            # the candidates really differ
