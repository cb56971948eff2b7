"""The skorch callbacks the reference configures (helper.py:197-273), restated for the B200
estimator (skorch is not a dependency here: SURVEY.md section 8c "parity unpinned" - the
semantics below follow skorch 0.10's public source).

    Checkpoint(monitor="valid_loss_best", dirname=workdir)              helper.py:211
    EarlyStopping(patience, threshold, threshold_mode, monitor="valid_loss",
                  lower_is_better=True, sink=log)                       helper.py:219-224
    GradientNormClipping(gradient_clip_value=0.5)                       helper.py:227-229
    EpochScoring(scoring, name, on_train, lower_is_better)              helper.py:235-268
    LRScheduler(policy="ReduceLROnPlateau", monitor="valid_loss", ...)  helper.py:241-245
"""
from __future__ import annotations

import json
import os

import numpy as np
import torch


class Callback:
    def initialize(self):
        return self

    def on_train_begin(self, net, **kw):
        pass

    def on_train_end(self, net, **kw):
        pass

    def on_epoch_begin(self, net, **kw):
        pass

    def on_epoch_end(self, net, **kw):
        pass

    def on_grad_computed(self, net, named_parameters, **kw):
        pass

    def get_params(self, deep=True):
        import inspect
        return {k: getattr(self, k) for k in inspect.signature(type(self).__init__).parameters
                if k not in ("self", "kwargs") and hasattr(self, k)}

    def set_params(self, **params):
        for k, v in params.items():
            setattr(self, k, v)
        return self


class GradientNormClipping(Callback):
    """clip_grad_norm_(parameters, max_norm=gradient_clip_value, norm_type) after backward.
    On the fused train step the estimator reads ``gradient_clip_value`` and the clip runs inside
    the gradnorm + SGD kernels (K11/K12); this hook serves the stock-autograd route."""

    def __init__(self, gradient_clip_value=None, gradient_clip_norm_type=2):
        self.gradient_clip_value = gradient_clip_value
        self.gradient_clip_norm_type = gradient_clip_norm_type

    def on_grad_computed(self, net, named_parameters, **kw):
        if self.gradient_clip_value is None:
            return
        torch.nn.utils.clip_grad_norm_((p for _, p in named_parameters), max_norm=self.gradient_clip_value,
                                       norm_type=self.gradient_clip_norm_type)


class _CachedNet:
    """What skorch's EpochScoring(use_caching=True) hands to the scorer: the net with its
    predictions for this epoch already computed (no second forward pass)."""

    def __init__(self, net, proba):
        self._net, self._proba = net, proba
        self.classes_ = net.classes_

    def predict_proba(self, X):
        return self._proba

    def predict(self, X):
        return self._proba.argmax(axis=1)

    def __getattr__(self, name):
        return getattr(self._net, name)

    # sklearn's scorers check the estimator type through tags
    def __sklearn_tags__(self):
        return self._net.__sklearn_tags__()


class EpochScoring(Callback):
    def __init__(self, scoring, lower_is_better=True, on_train=False, name=None, use_caching=True):
        self.scoring, self.lower_is_better, self.on_train = scoring, lower_is_better, on_train
        self.name, self.use_caching = name, use_caching

    def initialize(self):
        self.best_score_ = np.inf if self.lower_is_better else -np.inf
        self.name_ = self.name or getattr(self.scoring, "__name__", str(self.scoring))
        return self

    def on_train_begin(self, net, **kw):
        self.initialize()

    def on_epoch_end(self, net, dataset_train=None, dataset_valid=None, **kw):
        which = "train" if self.on_train else "valid"
        cache = net._epoch_cache.get(which)
        if cache is None:
            return
        proba, y_true, X = cache["proba"](), cache["y"], cache["X"]
        if getattr(self.scoring, "scorer", None) is not None or isinstance(self.scoring, str):
            est = _CachedNet(net, proba) if self.use_caching else net
            if isinstance(self.scoring, str):
                from sklearn.metrics import get_scorer
                score = get_scorer(self.scoring)(est, X, y_true)
            else:
                score = self.scoring(est, X, y_true)
        else:
            score = self.scoring(net, X, y_true)          # plain callable, e.g. the reference's lr_score
        score = float(score)
        is_best = score < self.best_score_ if self.lower_is_better else score > self.best_score_
        if is_best:
            self.best_score_ = score
        net.history[-1][self.name_] = score
        net.history[-1][self.name_ + "_best"] = bool(is_best)


class EarlyStopping(Callback):
    def __init__(self, monitor="valid_loss", patience=5, threshold=1e-4, threshold_mode="rel",
                 lower_is_better=True, sink=print):
        self.monitor, self.patience, self.threshold = monitor, patience, threshold
        self.threshold_mode, self.lower_is_better, self.sink = threshold_mode, lower_is_better, sink

    def on_train_begin(self, net, **kw):
        if self.threshold_mode not in ("rel", "abs"):
            raise ValueError("Invalid threshold mode: '{}'".format(self.threshold_mode))
        self.misses_ = 0
        self.dynamic_threshold_ = np.inf if self.lower_is_better else -np.inf

    def _is_improved(self, score):
        return score < self.dynamic_threshold_ if self.lower_is_better else score > self.dynamic_threshold_

    def _new_threshold(self, score):
        change = float(self.threshold) * score if self.threshold_mode == "rel" else float(self.threshold)
        return score - change if self.lower_is_better else score + change

    def on_epoch_end(self, net, **kw):
        current = net.history[-1][self.monitor]
        if not self._is_improved(current):
            self.misses_ += 1
        else:
            self.misses_ = 0
            self.dynamic_threshold_ = self._new_threshold(current)
        if self.misses_ == self.patience:
            if net.verbose and self.sink is not None:
                self.sink("Stopping since {} has not improved in the last {} epochs.".format(self.monitor, self.patience))
            net._stop_training = True


class LRScheduler(Callback):
    """policy: a torch.optim.lr_scheduler class or its name.  The scheduler is the STOCK torch
    class stepping ``net.optimizer_`` (a torch.optim.SGD whose lr the fused step reads)."""

    def __init__(self, policy="ReduceLROnPlateau", monitor="valid_loss", step_every="epoch", **kwargs):
        self.policy, self.monitor, self.step_every = policy, monitor, step_every
        self.kwargs = kwargs
        for k, v in kwargs.items():
            setattr(self, k, v)

    def get_params(self, deep=True):
        return dict(policy=self.policy, monitor=self.monitor, step_every=self.step_every, **self.kwargs)

    def on_train_begin(self, net, **kw):
        policy = getattr(torch.optim.lr_scheduler, self.policy) if isinstance(self.policy, str) else self.policy
        self.lr_scheduler_ = policy(net.optimizer_, **self.kwargs)

    def on_epoch_end(self, net, **kw):
        if self.step_every != "epoch":
            return
        if isinstance(self.lr_scheduler_, torch.optim.lr_scheduler.ReduceLROnPlateau):
            self.lr_scheduler_.step(net.history[-1][self.monitor])
        else:
            self.lr_scheduler_.step()


class Checkpoint(Callback):
    def __init__(self, monitor="valid_loss_best", dirname="", f_params="params.pt", f_optimizer="optimizer.pt",
                 f_history="history.json", sink=None):
        self.monitor, self.dirname = monitor, dirname
        self.f_params, self.f_optimizer, self.f_history, self.sink = f_params, f_optimizer, f_history, sink

    def on_epoch_end(self, net, **kw):
        if self.monitor is not None and not net.history[-1].get(self.monitor, False):
            return
        net.history[-1]["event_cp"] = True
        if not self.dirname and self.dirname != "":
            return
        d = self.dirname or "."
        os.makedirs(d, exist_ok=True)

        def atomic(path, writer):      # concurrent fits share `dirname` in the reference
            tmp = f"{path}.{os.getpid()}.tmp"
            writer(tmp)
            os.replace(tmp, path)
        if self.f_params:
            atomic(os.path.join(d, self.f_params), lambda p: torch.save(net.module_.state_dict(), p))
        if self.f_optimizer:
            atomic(os.path.join(d, self.f_optimizer), lambda p: torch.save(net.optimizer_state_dict(), p))
        if self.f_history:
            atomic(os.path.join(d, self.f_history), lambda p: json.dump(list(net.history), open(p, "w"), indent=1))


class PrintLog(Callback):
    def __init__(self, keys=("train_loss", "valid_loss", "dur"), sink=print):
        self.keys, self.sink = keys, sink

    def on_epoch_end(self, net, **kw):
        if not net.verbose:
            return
        h = net.history[-1]
        cols = ["epoch"] + [k for k in h if k in self.keys or (k.startswith(("valid_", "train_")) and not k.endswith("_best"))]
        cols = list(dict.fromkeys(c for c in cols if c in h))
        if h["epoch"] == 1:
            self.sink("  ".join(f"{c:>12}" for c in cols))
        self.sink("  ".join(f"{h[c]:12.4f}" if isinstance(h[c], float) else f"{h[c]:>12}" for c in cols))
