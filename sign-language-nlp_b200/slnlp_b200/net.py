"""skorch-shaped estimator for the B200 modules: ``NeuralNetClassifier`` with the surface the
reference's main.py / helper.py use (main.py:44,70-78,98-117; helper.py:41-105):

    NeuralNetClassifier(module=<class>, criterion=..., optimizer=..., lr, max_epochs, batch_size,
                        device, callbacks=[(name, cb), ...], module__*, optimizer__*, criterion__*,
                        iterator_train__*, iterator_valid__*, dataset=...)
    .fit(X, y) / .partial_fit / .predict / .predict_proba / .score / .get_params / .set_params

so sklearn's ``clone`` / ``GridSearchCV`` and the reference's ``ScoringWrapper`` drive it
unchanged.  skorch itself is not a dependency (absent offline); its train-step order is the
one restated in SURVEY.md section 3.2 ("parity unpinned", section 8c).

The training loop is device-resident: the whole (tiny) split is staged in HBM once, every
batch is one CUDA-graph replay of the fused train step (forward, CE on the log-probs,
backward, global-norm clip, SGD-momentum), losses and predictions stay on the device and come
back once per epoch.  Configurations the fused step does not cover (another criterion or
optimizer) run the same kernels through the autograd route with the stock torch classes.
"""
from __future__ import annotations

import time
from pydoc import locate

import numpy as np
import torch
from sklearn.base import BaseEstimator, ClassifierMixin

from . import callbacks as cbs
from ._lib import check, lib
from .data import SeqDataset, SeqSlice
from .flat import _stream
from .rnn import FusedTrainStep, InferStep, OptimState


import threading as _threading
_INIT_LOCK = _threading.Lock()


class History(list):
    """skorch History subset: ``h[-1]['valid_loss']``, ``h[-1, 'valid_loss']``, ``h[:, 'train_loss']``."""

    def __getitem__(self, i):
        if isinstance(i, tuple):
            idx, key = i
            rows = list.__getitem__(self, idx)
            return [r[key] for r in rows] if isinstance(idx, slice) else rows[key]
        return list.__getitem__(self, i)


class CVSplit:
    """skorch.dataset.CVSplit: the inner train/valid split of every fit.  cv = int k ->
    first fold of (Stratified)KFold(k) is the validation part; float -> that fraction."""

    def __init__(self, cv=5, stratified=True, random_state=None):
        self.cv, self.stratified, self.random_state = cv, stratified, random_state

    def __call__(self, n, y):
        from sklearn.model_selection import KFold, ShuffleSplit, StratifiedKFold, StratifiedShuffleSplit
        idx = np.arange(n)
        if isinstance(self.cv, float):
            cls = StratifiedShuffleSplit if self.stratified else ShuffleSplit
            cv = cls(n_splits=1, test_size=self.cv, random_state=self.random_state)
        else:
            cv = (StratifiedKFold if self.stratified else KFold)(n_splits=self.cv)
        try:
            return next(iter(cv.split(idx, y)))
        except ValueError:      # a class with fewer members than folds: skorch falls back the same way
            return next(iter(KFold(n_splits=self.cv if isinstance(self.cv, int) else 5).split(idx)))


def _as_tensors(X, y=None):
    """Accepts what the reference passes (AslSliceDataset -> SeqSlice), a SeqDataset, a dict
    {"X","lengths"[,"y"]} or a (tokens, lengths) tuple.  Returns CPU int64 (tokens, lengths, labels|None)."""
    if isinstance(X, SeqSlice):
        t, l, lab = X.tensors()
    elif isinstance(X, SeqDataset):
        t, l, lab = X.tokens, X.lengths, X.labels_
    elif isinstance(X, dict):
        t, l, lab = X["X"], X.get("lengths"), X.get("y")
    elif isinstance(X, (tuple, list)) and len(X) in (2, 3) and hasattr(X[0], "shape"):
        t, l, lab = X[0], X[1], (X[2] if len(X) == 3 else None)
    else:
        t, l, lab = X, None, None
    t = torch.as_tensor(np.asarray(t) if not torch.is_tensor(t) else t, dtype=torch.int64)
    if l is None:
        l = (t != 1).sum(1)                                    # util.resolve_lengths with <pad> = 1
    l = torch.as_tensor(np.asarray(l) if not torch.is_tensor(l) else l, dtype=torch.int64)
    if y is not None:
        lab = y.to_array() if hasattr(y, "to_array") else y
    if lab is not None:
        lab = torch.as_tensor(np.asarray(lab) if not torch.is_tensor(lab) else lab, dtype=torch.int64)
    return t.contiguous(), l.contiguous(), (None if lab is None else lab.contiguous())


class NeuralNetClassifier(ClassifierMixin, BaseEstimator):
    prefixes_ = ("module", "optimizer", "criterion", "callbacks", "iterator_train", "iterator_valid", "dataset")
    _soft_params = ("lr", "max_epochs", "batch_size", "verbose", "warm_start", "train_split", "predict_nonlinearity")

    def __init__(self, module, criterion=torch.nn.CrossEntropyLoss, optimizer=torch.optim.SGD, lr=0.01,
                 max_epochs=10, batch_size=128, device="cuda", callbacks=None, train_split="default",
                 predict_nonlinearity="auto", warm_start=False, verbose=1, dataset=None, classes=None,
                 precision="fp32", use_graph=True, **kwargs):
        self.module, self.criterion, self.optimizer = module, criterion, optimizer
        self.lr, self.max_epochs, self.batch_size, self.device = lr, max_epochs, batch_size, device
        self.callbacks, self.train_split, self.predict_nonlinearity = callbacks, train_split, predict_nonlinearity
        self.warm_start, self.verbose, self.dataset, self.classes = warm_start, verbose, dataset, classes
        self.precision, self.use_graph = precision, use_graph
        self._kwargs_keys = []
        for k, v in kwargs.items():
            if "__" not in k or k.split("__", 1)[0] not in self.prefixes_:
                raise TypeError(f"__init__() got an unexpected keyword argument {k!r}")
            setattr(self, k, v)
            self._kwargs_keys.append(k)
        self.initialized_ = False
        self.history = History()

    # ------------------------------------------------------------------ sklearn parameter protocol
    def get_params(self, deep=True, **kw):
        params = BaseEstimator.get_params(self, deep=False)
        for k in self._kwargs_keys:
            params[k] = getattr(self, k)
        return params

    def set_params(self, **params):
        for k, v in params.items():
            if "__" in k:
                if k.split("__", 1)[0] not in self.prefixes_:
                    raise ValueError(f"Invalid parameter {k!r} for estimator {type(self).__name__}")
                if k not in self._kwargs_keys:
                    self._kwargs_keys.append(k)
            elif k not in BaseEstimator.get_params(self, deep=False):
                raise ValueError(f"Invalid parameter {k!r} for estimator {type(self).__name__}")
            setattr(self, k, v)
            # skorch semantics: only parameters that change what initialize() builds force a fresh
            # initialisation; lr / max_epochs / batch_size / verbose / warm_start / iterator_* keep
            # the trained module (partial_fit / warm_start continue from it)
            if k not in self._soft_params and not k.startswith(("iterator_train__", "iterator_valid__")):
                self.initialized_ = False
            elif k == "lr" and getattr(self, "initialized_", False):
                for grp in self.optimizer_.param_groups:
                    grp["lr"] = v
        return self

    def _prefixed(self, prefix):
        n = len(prefix) + 2
        return {k[n:]: getattr(self, k) for k in self._kwargs_keys if k.startswith(prefix + "__")}

    # ------------------------------------------------------------------ initialisation
    @staticmethod
    def _resolve(obj):
        return locate(obj) if isinstance(obj, str) else obj

    def initialize(self):
        dev = torch.device(self.device) if not isinstance(self.device, torch.device) else self.device
        if dev.type != "cuda":
            raise RuntimeError("slnlp_b200: the estimator trains on CUDA only (no CPU fallback); got device=%r" % (self.device,))
        if dev.index is None:
            dev = torch.device("cuda", torch.cuda.current_device())
        self.device_ = dev
        mod_cls = self._resolve(self.module)
        mkw = dict(self._prefixed("module"))
        mkw.setdefault("precision", self.precision)
        mkw["device"] = dev
        fit_seed = getattr(self, "_fit_seed", None)
        if fit_seed is None:
            self.module_ = mod_cls(**mkw).to(dev)
        elif getattr(mod_cls, "accepts_init_generator", False):
            # a seeded fit (grid farm) of a module that draws its initial weights from a generator it is GIVEN: a private
            # one per fit - the same stream torch.manual_seed(fit_seed) would give, without touching (or queueing on) the
            # process-global generator the other worker threads share.  (Under the lock below the seeded construction of
            # every fit was one critical section: ~100 ms per fit on average over the LSTM grid, which capped the packed
            # farm at the same fits/hour for 2, 4 or 6 fits per GPU.)
            mkw.setdefault("seed", fit_seed)
            mkw["init_generator"] = torch.Generator().manual_seed(fit_seed)
            self.module_ = mod_cls(**mkw).to(dev)
        else:
            # a seeded fit (grid farm): the process-global torch generator is shared by every worker thread,
            # so the seeded construction is one critical section and leaves the generator as it found it
            with _INIT_LOCK:
                gen = torch.default_generator        # the CPU generator only: CUDA generators stay untouched
                state = gen.get_state()
                gen.manual_seed(fit_seed)
                mkw.setdefault("seed", fit_seed)
                try:
                    module = mod_cls(**mkw)
                finally:
                    gen.set_state(state)
            self.module_ = module.to(dev)
        self.V_ = self.module_.V_tgt
        self.classes_ = np.arange(self.V_) if self.classes is None else np.asarray(self.classes)
        # criterion / optimizer: the reference's configuration runs fused; anything else through autograd
        crit_cls, opt_cls = self._resolve(self.criterion), self._resolve(self.optimizer)
        ckw, okw = self._prefixed("criterion"), self._prefixed("optimizer")
        self.criterion_ = crit_cls(**ckw)
        clip = None
        self.callbacks_ = []
        # callbacks__<name>__<param> (helper.build_callbacks_args; a grid over e.g.
        # callbacks__early_stopping__patience) is routed to the callback of that name, on a per-fit
        # copy so that candidates of one search never share callback state - skorch's routing
        routed = {}
        for key, val in self._prefixed("callbacks").items():
            cb_name, sep, param = key.partition("__")
            if not sep:
                raise ValueError(f"callbacks__{key}: expected callbacks__<name>__<param>")
            routed.setdefault(cb_name, {})[param] = val
        for item in (self.callbacks or []):
            name, cb = item if isinstance(item, tuple) else (type(item).__name__, item)
            if name in routed:
                import copy
                cb = copy.copy(cb)
                known = cb.get_params() if hasattr(cb, "get_params") else {}
                for param in routed[name]:
                    if param not in known and not hasattr(cb, param):
                        raise ValueError(f"Invalid parameter {param!r} for callback {name!r} ({type(cb).__name__})")
                cb.set_params(**routed.pop(name))
            self.callbacks_.append((name, cb))
            if isinstance(cb, cbs.GradientNormClipping):
                clip = cb
        if routed:
            raise ValueError(f"callbacks__ parameters name unknown callbacks: {sorted(routed)} "
                             f"(have {[n for n, _ in self.callbacks_]})")
        if self.verbose and not any(isinstance(cb, cbs.PrintLog) for _, cb in self.callbacks_):
            self.callbacks_.append(("print_log", cbs.PrintLog()))
        for _, cb in self.callbacks_:
            cb.initialize()
        sgd_ok = (opt_cls is torch.optim.SGD and not okw.get("nesterov", False) and not okw.get("weight_decay", 0)
                  and not okw.get("dampening", 0))
        ce_ok = (crit_cls is torch.nn.CrossEntropyLoss and set(ckw) <= {"ignore_index"}
                 and ckw.get("ignore_index", self.module_.tgt_pad) == self.module_.tgt_pad)
        clip_ok = clip is None or clip.gradient_clip_norm_type == 2
        self.fused_ = bool(sgd_ok and ce_ok and clip_ok)
        self.max_norm_ = float(clip.gradient_clip_value) if (clip is not None and clip.gradient_clip_value) else 0.0
        if self.fused_:
            # a stock torch SGD over a dummy parameter carries lr for LR schedulers / lr_score;
            # the update itself runs in the fused kernel from OptimState
            self._lr_holder = torch.nn.Parameter(torch.zeros(1))
            self.optimizer_ = torch.optim.SGD([self._lr_holder], lr=self.lr, **okw)
            self.opt_state_ = OptimState(self.module_, self.lr, float(okw.get("momentum", 0.0)), self.max_norm_)
            self.steps_ = {}
        else:
            self.optimizer_ = opt_cls(self.module_.parameters(), lr=self.lr, **okw)
        self.history = History()
        self._epoch_cache = {}
        self.infer_steps_ = {}
        self._stop_training = False
        self.initialized_ = True
        return self

    def release_graphs(self):
        """Drop every captured step / scoring graph of this estimator (they are re-captured on demand).
        The grid farm calls it when a fit is done: graph destruction must not race another worker
        thread's capture (flat.CAPTURE_LOCK)."""
        for step in list(getattr(self, "steps_", {}).values()) + list(getattr(self, "infer_steps_", {}).values()):
            step.release()
        if hasattr(self, "steps_"):
            self.steps_ = {}
        self.infer_steps_ = {}

    def optimizer_state_dict(self):
        """torch.optim.SGD.state_dict() layout on both routes (skorch Checkpoint's optimizer.pt)."""
        return self.opt_state_.state_dict(self.module_) if self.fused_ else self.optimizer_.state_dict()

    def load_optimizer_state_dict(self, sd):
        if self.fused_:
            self.opt_state_.load_state_dict(sd, self.module_)
        else:
            self.optimizer_.load_state_dict(sd)

    # ------------------------------------------------------------------ training
    def _fused_step(self, B, T):
        key = (B, T)
        if key not in self.steps_:
            self.steps_[key] = FusedTrainStep(self.module_, B, T, use_graph=self.use_graph, state=self.opt_state_)
        return self.steps_[key]

    def _infer_step(self, B, T):
        """Full batches of the scoring / validation forward: one captured graph per (batch, len) shape."""
        if not self.use_graph:
            return None
        cache = self.__dict__.setdefault("infer_steps_", {})
        if (B, T) not in cache:
            cache[(B, T)] = InferStep(self.module_, B, T)
        return cache[(B, T)]

    def _eval_loss_and_logp(self, Xd, ld, yd, out_logp, train_mode=False):
        """Forward-only pass over a device-resident split; returns the device scalar
        sum_b loss_b * n_b and fills out_logp [N, V]."""
        m, bs, N = self.module_, self.batch_size, Xd.shape[0]
        tot = torch.zeros((), device=Xd.device)
        loss = torch.zeros(2, device=Xd.device)
        row_ws = torch.empty(3 * bs, device=Xd.device)
        m.eval()
        infer = self._infer_step(bs, Xd.shape[1]) if N >= bs else None
        for j in range(0, N, bs):
            k = min(N, j + bs)
            if infer is not None and k - j == bs:
                logp = infer.step(Xd[j:k], yd[j:k], ld[j:k])
            else:
                logp = m.predict_logp(Xd[j:k], ld[j:k], yd[j:k])
            out_logp[j:k].copy_(logp)
            if self.fused_:
                check(lib.slnlp_ce_on_logp(logp.data_ptr(), yd[j:k].data_ptr(), m.tgt_pad, k - j, m.V_tgt, loss.data_ptr(),
                                           None, 0, row_ws.data_ptr(), _stream()), "ce")
                tot += loss[0] * (k - j)
            else:
                tot += self.criterion_(logp, yd[j:k]) * (k - j)
        return tot

    def fit(self, X, y=None, **fit_params):
        if not self.warm_start or not self.initialized_:
            self.initialize()
        return self.partial_fit(X, y, **fit_params)

    def partial_fit(self, X, y=None, **fit_params):
        if not self.initialized_:
            self.initialize()
        tok, lens, lab = _as_tensors(X, y)
        if lab is None:
            raise ValueError("fit needs labels")
        n = tok.shape[0]
        split = CVSplit(5) if self.train_split == "default" else self.train_split
        if split is None:
            itr, iva = np.arange(n), np.arange(0)
        else:
            itr, iva = split(n, lab.numpy())
        dev = self.device_
        itr_t, iva_t = torch.from_numpy(itr), torch.from_numpy(iva)
        # the whole split lives in HBM for the fit (helper.py:293-304 collates per batch on the host)
        Xtr, ltr, ytr = tok[itr_t].to(dev), lens[itr_t].to(dev), lab[itr_t].to(dev)
        Xva, lva, yva = tok[iva_t].to(dev), lens[iva_t].to(dev), lab[iva_t].to(dev)
        ntr, nva, T, bs, V = len(itr), len(iva), tok.shape[1], self.batch_size, self.V_
        m = self.module_
        logp_tr = torch.empty(ntr, V, device=dev)
        logp_va = torch.empty(nva, V, device=dev)
        nb = (ntr + bs - 1) // bs
        batch_loss = torch.zeros(nb, device=dev)
        sizes = torch.tensor([min(bs, ntr - j * bs) for j in range(nb)], device=dev, dtype=torch.float32)
        train_slice = SeqSlice(SeqDataset(tok[itr_t], lens[itr_t], lab[itr_t], m.src_vocab, m.tgt_vocab), 0)
        valid_slice = SeqSlice(SeqDataset(tok[iva_t], lens[iva_t], lab[iva_t], m.src_vocab, m.tgt_vocab), 0)
        ytr_np, yva_np = lab[itr_t].numpy(), lab[iva_t].numpy()
        best = {"train_loss": np.inf, "valid_loss": np.inf}
        self._stop_training = False
        for _, cb in self.callbacks_:
            cb.on_train_begin(self, X=X, y=y)
        for _ in range(self.max_epochs):
            t0 = time.perf_counter()
            for _, cb in self.callbacks_:
                cb.on_epoch_begin(self)
            m.train()
            if self.fused_:
                self.opt_state_.set_lr(float(self.optimizer_.param_groups[0]["lr"]))
                for j in range(nb):
                    a, b = j * bs, min(ntr, (j + 1) * bs)
                    ts = self._fused_step(b - a, T)
                    loss = ts.step(Xtr[a:b], ytr[a:b], ltr[a:b])
                    batch_loss[j].copy_(loss[0])
                    logp_tr[a:b].copy_(ts.ws.logp)
            else:
                for j in range(nb):
                    a, b = j * bs, min(ntr, (j + 1) * bs)
                    self.optimizer_.zero_grad()
                    logp = m(X=Xtr[a:b], y=ytr[a:b], lengths=ltr[a:b])
                    loss = self.criterion_(logp, ytr[a:b])
                    loss.backward()
                    for _, cb in self.callbacks_:
                        cb.on_grad_computed(self, named_parameters=list(m.named_parameters()))
                    self.optimizer_.step()
                    batch_loss[j].copy_(loss.detach())
                    logp_tr[a:b].copy_(logp.detach())
            row = {"epoch": len(self.history) + 1}
            train_loss = (batch_loss * sizes).sum() / ntr
            if nva:
                with torch.no_grad():
                    valid_loss = self._eval_loss_and_logp(Xva, lva, yva, logp_va) / nva
                tl, vl = torch.stack([train_loss, valid_loss]).tolist()       # ONE read-back per epoch
                row["train_loss"], row["valid_loss"] = tl, vl
            else:
                row["train_loss"] = float(train_loss)
            for k in ("train_loss", "valid_loss"):
                if k in row:
                    row[k + "_best"] = bool(row[k] < best[k])
                    best[k] = min(best[k], row[k])
            self.history.append(row)
            # predictions of this epoch for the scoring callbacks (skorch caching semantics):
            # fetched lazily, at most once per split
            self._epoch_cache = {
                "train": {"proba": _Lazy(lambda: logp_tr.exp().cpu().numpy()), "y": ytr_np, "X": train_slice},
                "valid": ({"proba": _Lazy(lambda: logp_va.exp().cpu().numpy()), "y": yva_np, "X": valid_slice} if nva else None),
            }
            row["dur"] = time.perf_counter() - t0
            for _, cb in self.callbacks_:
                cb.on_epoch_end(self, dataset_train=train_slice, dataset_valid=valid_slice)
            if self._stop_training:
                break
        for _, cb in self.callbacks_:
            cb.on_train_end(self, X=X, y=y)
        self._epoch_cache = {}
        return self

    # ------------------------------------------------------------------ inference
    @torch.no_grad()
    def forward_logp(self, X):
        if not self.initialized_:
            raise RuntimeError("This NeuralNetClassifier instance is not initialized yet. Call 'initialize' or 'fit'.")
        tok, lens, lab = _as_tensors(X)
        m, dev, bs = self.module_, self.device_, self.batch_size
        if lab is None:
            if type(m).__name__.startswith("Transformer"):
                raise ValueError("model.Transformer feeds the label to its decoder (transformer.py:65): pass X with labels")
            lab = torch.zeros(tok.shape[0], dtype=torch.int64)
        Xd, ld, yd = tok.to(dev), lens.to(dev), lab.to(dev)
        out = torch.empty(tok.shape[0], self.V_, device=dev)
        was = m.training
        m.eval()
        n, T = tok.shape
        infer = self._infer_step(bs, T) if n >= bs else None
        for j in range(0, n, bs):
            k = min(n, j + bs)
            if infer is not None and k - j == bs:
                out[j:k].copy_(infer.step(Xd[j:k], yd[j:k], ld[j:k]))
            else:
                out[j:k].copy_(m.predict_logp(Xd[j:k], ld[j:k], yd[j:k]))
        m.train(was)
        return out

    def predict_proba(self, X):
        """predict_nonlinearity='auto' with CrossEntropyLoss = softmax over the module output; the
        output is already log-probabilities, so softmax(logp) == exp(logp)."""
        logp = self.forward_logp(X)
        nl = self.predict_nonlinearity
        if nl is None:
            return logp.cpu().numpy()
        if callable(nl):
            return nl(logp).cpu().numpy()
        return torch.softmax(logp, dim=-1).cpu().numpy()

    def predict(self, X):
        return self.forward_logp(X).argmax(dim=1).cpu().numpy()


class _Lazy:
    def __init__(self, fn):
        self.fn, self.v = fn, None

    def __call__(self):
        if self.v is None:
            self.v = self.fn()
        return self.v
