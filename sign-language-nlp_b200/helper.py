"""Train / score glue between the YAML arguments and the estimator: the functions of the
reference's helper.py that define the drop-in boundary, same names
and argument meaning, built on the B200 estimator instead of skorch + Dask.

    build_net_params  helper.py:41-105      build_grid_params  helper.py:108-180
    build_callbacks   helper.py:197-273     build_scoring      helper.py:276-283
    collate_data      helper.py:293-304     prefix_args        helper.py:325-341
    ScoringWrapper    helper.py:529-554     create_worker_farm ~ create_dask_client helper.py:490-526
"""
import json
import os
import random
from pydoc import locate

import numpy as np
import torch

from slnlp_b200.callbacks import Checkpoint, EarlyStopping, EpochScoring, GradientNormClipping, LRScheduler
from slnlp_b200.data import SeqDataset

PAD_WORD = "<pad>"


def log(*a):
    print(*a, flush=True)


def setup_seed(seed, **kwargs):
    torch.manual_seed(seed)
    random.seed(seed)
    np.random.seed(seed)


def prepare_device(cuda):
    if not cuda:
        raise RuntimeError("the B200 build has no CPU path: run with --cuda True")
    if not torch.cuda.is_available():
        raise RuntimeError("--cuda was given but no CUDA device is visible")
    return torch.device("cuda")


def dump_args(args):
    import yaml
    os.makedirs(args["workdir"], exist_ok=True)
    plain = {k: v for k, v in args.items() if isinstance(v, (str, int, float, bool, list, dict, type(None)))}
    with open(os.path.join(args["workdir"], "config.yaml"), "w") as f:
        yaml.safe_dump(plain, f)


def prefix_args(prefix, ensure_list=False, output=None, **kwargs):
    """{"a": {"b": 1}} under prefix "p" -> {"p__a__b": 1} (skorch / sklearn routing names)."""
    output = {} if output is None else output
    for key, value in kwargs.items():
        name = key if prefix is None else f"{prefix}__{key}"
        if isinstance(value, dict):
            prefix_args(prefix=name, ensure_list=ensure_list, output=output, **value)
        else:
            output[name] = [value] if (ensure_list and not isinstance(value, list)) else value
    return output


def filter_by_keys(map, keys_to_filter, not_in=False):
    return {k: v for k, v in map.items() if (k in keys_to_filter) != not_in}


def collate_data(data):
    """list of ((tokens, length[, label]), y) -> ({"X","lengths","y"} LongTensors, y).  The B200
    estimator stages whole splits on the device instead, but the function keeps working for any
    caller that batches on the host."""
    X, y = zip(*data)
    cols = list(zip(*X))
    X, X_lengths = cols[0], cols[1]
    y = torch.tensor(y, dtype=torch.long)
    return {"X": torch.tensor(X, dtype=torch.long), "lengths": torch.tensor(X_lengths, dtype=torch.long), "y": y}, y


class ScoringWrapper:
    """A named sklearn scorer, callable as scorer(estimator, X, y) (helper.py:529-554)."""

    def __init__(self, score_func, labels=None):
        from sklearn.metrics import get_scorer
        self._score_func = score_func
        self.scorer = get_scorer(score_func)
        if score_func == "neg_log_loss":
            self.scorer._kwargs["labels"] = labels
        elif score_func != "accuracy":
            self.scorer._kwargs["zero_division"] = 0

    def __call__(self, estimator, X, y_true, sample_weight=None):
        if sample_weight is None:
            return self.scorer(estimator, X, y_true)
        return self.scorer(estimator, X, y_true, sample_weight=sample_weight)

    def __repr__(self):
        return f"{type(self).__name__}('{self._score_func}')"

    @property
    def greater_is_better(self):
        return self.scorer._sign == 1

    @property
    def score(self):
        return self._score_func


def build_scoring(scoring, labels=None, allow_multiple=True):
    scoring = scoring if isinstance(scoring, list) else [scoring]
    wrappers = [ScoringWrapper(s, labels) for s in scoring]
    return wrappers if allow_multiple else wrappers[0]


def lr_score(net, X=None, y=None):
    return net.optimizer_.param_groups[0]["lr"]


def build_callbacks(mode, workdir, scoring, dataset, early_stopping=None, gradient_clipping=None,
                    lr_scheduler=None, **kwargs):
    monitor = "valid"
    callbacks = [("checkpoint", Checkpoint(monitor=f"{monitor}_loss_best", dirname=workdir))]
    if early_stopping:
        callbacks.append(("early_stopping", EarlyStopping(**early_stopping, monitor=f"{monitor}_loss",
                                                          lower_is_better=True, sink=log)))
    if gradient_clipping:
        callbacks.append(("gradient_clipping", GradientNormClipping(**gradient_clipping)))
    callbacks.append(("lr_scoring", EpochScoring(scoring=lr_score, name="lr", on_train=False)))
    if lr_scheduler:
        callbacks.append(("lr_scheduler", LRScheduler(monitor=f"{monitor}_loss", step_every="epoch", **lr_scheduler)))
    for wrapper in build_scoring(scoring, dataset.labels(), allow_multiple=True):
        for on_train in (False, True):
            split = "train" if on_train else "valid"
            callbacks.append((f"score_{split}_{wrapper.score}",
                              EpochScoring(scoring=wrapper, name=f"{split}_{wrapper.score}", on_train=on_train,
                                           lower_is_better=not wrapper.greater_is_better)))
    return callbacks, [name for name, _ in callbacks]


def build_callbacks_args(callbacks_names, ensure_list=False, **kwargs):
    wanted = filter_by_keys(kwargs, list(callbacks_names) + ["print_log"])
    return prefix_args("callbacks", ensure_list=ensure_list, **wanted)


def build_net_params(model_args, model, optimizer, criterion, callbacks, callbacks_names, device, dataset,
                     optimizer_args, criterion_args=None, **kwargs):
    criterion_args = dict(criterion_args or {})
    criterion_args["ignore_index"] = dataset.vocab_y.stoi[PAD_WORD]      # model/util/util.py:5-6
    iterators = {"collate_fn": collate_data}
    net_args = filter_by_keys(kwargs, ["lr", "max_epochs", "batch_size", "predict_nonlinearity", "warm_start",
                                       "verbose", "precision"])
    net_args = {k: v for k, v in net_args.items() if v is not None or k == "lr"}
    if net_args.get("lr") is None:
        net_args["lr"] = 0.01            # "tuned in grid search": any placeholder, the grid overrides it
    return {
        "device": device,
        "module": locate(model),
        "optimizer": locate(optimizer),
        "criterion": locate(criterion),
        "callbacks": callbacks,
        "dataset": SeqDataset,
        **net_args,
        **prefix_args("module", batch_first=dataset.batch_first, src_vocab=dataset.vocab_X,
                      tgt_vocab=dataset.vocab_y, device=device, **{k: v for k, v in model_args.items()}),
        **prefix_args("optimizer", **optimizer_args),
        **prefix_args("criterion", **criterion_args),
        **build_callbacks_args(callbacks_names=callbacks_names, model=model, **kwargs),
        **prefix_args("iterator_train", **iterators),
        **prefix_args("iterator_valid", **iterators),
    }


def build_grid_params(grid_args, callbacks_names, model, workdir, scoring, verbose, n_jobs, cv, data, **kwargs):
    grid_args = dict(grid_args)
    model_args = grid_args.pop("model_args", {})
    optimizer_args = grid_args.pop("optimizer_args", {})
    criterion_args = grid_args.pop("criterion_args", {})
    grid_args.pop("training_args", None)
    passthrough = filter_by_keys(grid_args, ["n_jobs", "refit", "verbose", "pre_dispatch", "return_train_score"])
    general = filter_by_keys(grid_args, list(passthrough), not_in=True)
    return {
        "refit": True, "cv": cv, "verbose": verbose, "n_jobs": n_jobs, "error_score": "raise",
        "scoring": build_scoring(scoring, data.labels(), allow_multiple=False),
        **passthrough,
        "param_grid": {
            **prefix_args("module", ensure_list=True, **model_args),
            **prefix_args("optimizer", ensure_list=True, **optimizer_args),
            **prefix_args("criterion", ensure_list=True, **criterion_args),
            **build_callbacks_args(callbacks_names=callbacks_names, ensure_list=True, **general),
            **prefix_args(None, ensure_list=True, **filter_by_keys(general, list(callbacks_names) + ["print_log"], not_in=True)),
        },
    }


def format_dir(dir, **kwargs):
    from datetime import datetime
    if dir is None:
        return ""
    return os.path.normpath(dir.format(**{"datetime": datetime.now(), **kwargs}))


def balance_dataset(dataset, seed):
    """Log-smoothed under- then over-sampling towards the mean class count (helper.py:344-388;
    imbalanced-learn's RandomUnderSampler / RandomOverSampler semantics with numpy draws -
    the exact picks are "parity unpinned")."""
    import math
    from collections import Counter
    y = dataset.y().to_array()
    counts = Counter(y.tolist())
    u = sum(counts.values()) / len(counts)

    def smooth(v, sign):
        t = round(u + math.log(v))
        return v if v * sign > t * sign else t
    under = {k: smooth(v, -1) for k, v in counts.items()}
    over = {k: smooth(v, +1) for k, v in under.items()}
    rng = np.random.RandomState(seed)
    picked = []
    for k in sorted(counts):
        idx = np.flatnonzero(y == k)
        if under[k] < len(idx):
            idx = rng.choice(idx, size=under[k], replace=False)
        if over[k] > len(idx):
            idx = np.concatenate([idx, rng.choice(idx, size=over[k] - len(idx), replace=True)])
        picked.append(idx)
    return dataset._subset(np.concatenate(picked))


# ---------------------------------------------------------------------- outputs (helper.py:399-439)
def _jsonable(o):
    if isinstance(o, (np.integer,)):
        return int(o)
    if isinstance(o, (np.floating,)):
        return float(o)
    if isinstance(o, np.ndarray):
        return o.tolist()
    return str(o)


def save_param_grid(grid_params, phase, workdir, **kwargs):
    import itertools
    import pandas as pd
    cols = list(grid_params.keys())
    df = pd.DataFrame(list(itertools.product(*grid_params.values())), columns=cols)
    df.to_csv(f"{workdir}/{phase}_grid_params.csv")


def save_cv_results(cv_results, phase, workdir, **kwargs):
    import pandas as pd
    pd.DataFrame(cv_results).to_csv(f"{workdir}/{phase}_results.csv")


def save_output(output, phase, workdir, **kwargs):
    log(output)
    with open(f"{workdir}/{phase}_output.json", "w") as f:
        json.dump(output, f, indent=1, default=_jsonable)


class KernelProfile:
    """Stand-in for the torch.profiler block around ONE estimator.predict (main.py:116-117;
    helper.py:391-396,442-487): device time of the call from CUDA events + kernels launched
    through the C ABI, written as <phase>_profile.json."""

    def __init__(self, cuda=True):
        self.cuda = cuda

    def __enter__(self):
        from slnlp_b200._lib import lib
        self._l0 = lib.slnlp_launch_count()
        self._e0, self._e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        import time
        self._t0 = time.perf_counter()
        self._e0.record()
        return self

    def __exit__(self, *exc):
        from slnlp_b200._lib import lib
        import time
        self._e1.record()
        torch.cuda.synchronize()
        self.details = {"cuda_time_total_us": self._e0.elapsed_time(self._e1) * 1e3,
                        "wall_time_total_us": (time.perf_counter() - self._t0) * 1e6,
                        "kernel_launches": int(lib.slnlp_launch_count() - self._l0),
                        "device_type": "CUDA"}


def create_profiler(cuda):
    return KernelProfile(cuda)


def save_profile(profiler, phase, workdir, **kwargs):
    with open(f"{workdir}/{phase}_profile.json", "w") as f:
        json.dump(profiler.details, f, indent=1)


def create_worker_farm(gpus=None, dask_args=None, **kwargs):
    """What create_dask_client (helper.py:490-526) becomes on one B200 box: the number of
    per-GPU worker processes the grid search farms its fits over."""
    n = torch.cuda.device_count() if torch.cuda.is_available() else 0
    n = min(n, gpus) if gpus else n
    log(f" > Worker farm: {n} GPU worker(s), one process per GPU, no collectives")
    return n


def load_dataset(dataset_args, debug=False, seed=1, **kwargs):
    """``AslDataset(device, batch_first=True, **args).stoi()`` (main.py:25).  Three sources produce
    its output contract (X [N,T] int64 padded with <pad>, lengths [N], y [N], torchtext-shaped vocabs):
      dataset_args.tensor_file: a torch.save'd dict {X, lengths, y, src_itos, tgt_itos};
      dataset_args.dataset_dir (an existing directory of ASL-Phono JSON samples): read and composed
        like dataset/builder/dataset_builder.py does, without torchtext (slnlp_b200/phono.py);
      dataset_args.synthetic ({n_seq, T, v_src, v_tgt, ragged}; bench and tests): a synthetic corpus shaped
        like it.  A dataset_dir that does not exist raises FileNotFoundError."""
    from slnlp_b200.vocab import Vocab
    da = dataset_args or {}
    tf = da.get("tensor_file")
    if tf:
        d = torch.load(tf)
        return SeqDataset(d["X"], d["lengths"], d["y"], Vocab(d["src_itos"][2:]), Vocab(d["tgt_itos"][2:]))
    if da.get("dataset_dir") and not da.get("synthetic"):
        if os.path.isdir(da["dataset_dir"]):
            from slnlp_b200.phono import build_dataset
            return build_dataset(da["dataset_dir"], da["fields"], da.get("samples_min_freq", 1),
                                 da.get("composition_strategy", "as_words"))
        # a mis-set path must not silently train, grid-search and "test" on random labels
        raise FileNotFoundError(f"corpus directory {da['dataset_dir']!r} does not exist "
                                "(give dataset_args.synthetic explicitly to use the synthetic corpus)")
    if "synthetic" not in da:
        raise FileNotFoundError("dataset_args names no corpus: set dataset_dir, tensor_file, or synthetic: {...}")
    syn = dict(n_seq=2000, T=64, v_src=4098, v_tgt=1026, ragged=True, seed=seed)
    syn.update(da.get("synthetic") or {})
    return SeqDataset.synthetic(**syn)
