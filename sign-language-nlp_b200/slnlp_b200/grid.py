"""Grid search over independent fits, farmed one worker process per GPU with NO collective
(SURVEY.md section 8e): the B200-box replacement of the reference's
``GridSearchCV(n_jobs=-1)`` under ``joblib.parallel_backend('dask')`` with one Dask worker per
GPU (main.py:70-78; helper.py:490-526; cluster/az-start-workers.sh:11-15).

``GridSearchFarm`` takes GridSearchCV's keyword arguments (helper.py:154-172) and exposes its
result surface: ``cv_results_`` (same column names), ``best_index_``, ``best_score_``,
``best_params_``, ``best_estimator_`` (refit on all the data).

``fits_per_gpu`` > 1 (opt-in; default 1 or $SLNLP_FITS_PER_GPU) packs several fits on one GPU: at the
reference's batch of 50 a fit's recurrent kernels occupy 26-50 of 148 SMs and its step is a chain
of dependent launches, so k worker THREADS per process, each with its own CUDA stream, estimator
and captured step graph, run side by side (one process per GPU: no context switches, no MPS).

Units of work = (candidate, fold) fits.  They are ordered longest-first by a FLOP estimate and
claimed dynamically:
  * ``backend="spawn"``: this process spawns one worker per GPU and feeds a queue;
  * ``backend="torchrun"``: every rank of a torchrun launch calls ``fit`` with the same
    arguments; a fetch-and-add counter in the rendezvous store hands out work and carries the
    results back - host TCP only, no NCCL, no GPU-GPU traffic.
"""
from __future__ import annotations

import contextlib
import json
import os
import pickle
import queue
import threading
import time
import traceback
from typing import Dict

import numpy as np
from sklearn.base import clone, is_classifier
from sklearn.model_selection import ParameterGrid, check_cv


def estimate_cost(params: Dict) -> float:
    """Relative training cost of one fit (SURVEY.md section 8d FLOP formulas, G folded in)."""
    E = params.get("module__embedding_size", 128) or 128
    H = params.get("module__hidden_size", 128) or 128
    L = params.get("module__num_layers", 1) or 1
    return float(sum(H * ((E if l == 0 else 2 * H) + H) for l in range(L)))


def _fit_and_score(estimator, params, X, y, train, test, scorer, fit_subdir=None, seed=None):
    """One (candidate, fold): clone -> set_params -> fit(train) -> scorer(test).  ``seed`` (derived from the
    candidate and fold, not from scheduling) seeds the fit's weight initialisation and dropout stream, so a
    fit's history does not depend on which worker ran it or on what ran beside it."""
    from sklearn.utils import _safe_indexing
    est = clone(estimator).set_params(**params)
    if seed is not None:
        est._fit_seed = int(seed)
    if fit_subdir is not None:      # per-fit checkpoint directory (the reference's fits clobber one dir)
        for item in (getattr(est, "callbacks", None) or []):
            cb = item[1] if isinstance(item, tuple) else item
            if hasattr(cb, "dirname") and cb.dirname:
                cb.dirname = os.path.join(cb.dirname, fit_subdir)
    Xtr, ytr = _safe_indexing(X, train), _safe_indexing(y, train)
    Xte, yte = _safe_indexing(X, test), _safe_indexing(y, test)
    t0 = time.perf_counter()
    try:
        est.fit(Xtr, ytr)
        t1 = time.perf_counter()
        score = float(scorer(est, Xte, yte))
        t2 = time.perf_counter()
    finally:
        if hasattr(est, "release_graphs"):     # captured CUDA graphs go now, under the capture lock
            est.release_graphs()
    return {"score": score, "fit_time": t1 - t0, "score_time": t2 - t1,
            "epochs": len(getattr(est, "history", [])), "n_train": len(train)}


def _stream_scope():
    """A private CUDA stream for a packed worker thread (nothing on a CPU-only host)."""
    import torch
    if not torch.cuda.is_available():
        return contextlib.nullcontext()
    from .flat import thread_stream     # a stream of this thread's own, not one from torch's shared pool
    return torch.cuda.stream(thread_stream("main"))


def _pack_env(fits_per_gpu):
    if fits_per_gpu > 1:
        # several threads capture step graphs in one process: captures must not see each other's CUDA calls
        os.environ.setdefault("SLNLP_CAPTURE_MODE", "thread_local")


def _worker_loop(gpu, task_q, result_q, payload_bytes, own_stream):
    import torch
    try:
        if torch.cuda.is_available():
            torch.cuda.set_device(gpu)        # the current device is per thread
        # a spawned process receives pickled bytes; threads of this process share the objects themselves
        # (the estimator prototype is only ever cloned, never fitted)
        estimator, X, y, scorer = pickle.loads(payload_bytes) if isinstance(payload_bytes, bytes) else payload_bytes
        if gpu is not None and torch.cuda.is_available() and "device" in estimator.get_params():
            estimator.set_params(device=f"cuda:{gpu}")
        with (_stream_scope() if own_stream else contextlib.nullcontext()):
            while True:
                task = task_q.get()
                if task is None:
                    break
                tid, params, train, test, sub, seed = task
                try:
                    out = _fit_and_score(estimator, params, X, y, train, test, scorer, sub, seed)
                    out["gpu"] = gpu
                    result_q.put((tid, out, None))
                except Exception:            # error_score="raise": reported to the parent, which raises
                    result_q.put((tid, None, traceback.format_exc()))
    except Exception:
        result_q.put((-1, None, traceback.format_exc()))


def _worker_main(gpu, task_q, result_q, payload_bytes, fits_per_gpu=1):
    """One process per GPU; fits_per_gpu threads inside it (each ends on its own None sentinel)."""
    _pack_env(fits_per_gpu)
    if fits_per_gpu <= 1:
        _worker_loop(gpu, task_q, result_q, payload_bytes, False)
    else:
        threads = [threading.Thread(target=_worker_loop, args=(gpu, task_q, result_q, payload_bytes, True), daemon=True)
                   for _ in range(fits_per_gpu)]
        for t in threads:
            t.start()
        for t in threads:
            t.join()
    try:        # this process's kernel launches, for the parent's account (bench.py's gpu_launches claim)
        from ._lib import lib
        result_q.put((-2, {"launches": int(lib.slnlp_launch_count())}, None))
    except Exception:
        pass


class GridSearchFarm:
    def __init__(self, estimator, param_grid, *, scoring=None, n_jobs=None, refit=True, cv=None, verbose=0,
                 pre_dispatch=None, error_score="raise", return_train_score=False, n_gpus=None, backend="auto",
                 per_fit_checkpoint_dirs=True, resume_file=None, fits_per_gpu=None, random_state=1, procs_per_gpu=None):
        self.estimator, self.param_grid, self.scoring, self.n_jobs = estimator, param_grid, scoring, n_jobs
        self.refit, self.cv, self.verbose, self.pre_dispatch = refit, cv, verbose, pre_dispatch
        self.error_score, self.return_train_score = error_score, return_train_score
        self.n_gpus, self.backend, self.per_fit_checkpoint_dirs = n_gpus, backend, per_fit_checkpoint_dirs
        # resume_file: JSON-lines journal of finished (candidate, fold) fits.  A search that is started
        # again with the same grid and folds skips what the journal already holds (the reference aborts
        # the whole grid on any failure, helper.py:162, and has no resume - SURVEY.md section 5)
        self.resume_file = resume_file
        self.fits_per_gpu = fits_per_gpu
        # procs_per_gpu worker PROCESSES share each GPU, each with fits_per_gpu threads (spawn / inline backends; the torchrun
        # backend keeps one process per rank).  The idea: a small fit is a few ms of GPU work inside ~50 ms of Python, and
        # threads of one process queue for its interpreter lock.  MEASURED on one B200 without MPS, full 810-fit cfg5 grid
        # (profiles/r02_grid_procs_*.json): 1 process x 4 threads 48.3 s; 3 x 4: 57.3 s; 4 x 2: 58.8 s; 6 x 1: 202.9 s - contexts
        # of different processes time-slice the GPU instead of sharing it, which costs more than the lock.  Default 1.
        self.procs_per_gpu = procs_per_gpu
        # every fit is seeded from (random_state, candidate, fold): the reference seeds the process once
        # (main.py:21, seed 1) and its fits then draw from whatever the worker's RNG holds; None restores that
        self.random_state = random_state

    # ------------------------------------------------------------------ scheduling
    def _tasks(self, X, y):
        cands = list(ParameterGrid(self.param_grid))
        cv = check_cv(self.cv, y, classifier=is_classifier(self.estimator))
        folds = [(tr, te) for tr, te in cv.split(np.arange(len(y)), y)]
        tasks = [(ci, fi) for ci in range(len(cands)) for fi in range(len(folds))]
        # longest first, so no GPU is left holding a 6-layer fit at the end
        order = sorted(range(len(tasks)), key=lambda t: -estimate_cost(cands[tasks[t][0]]) * len(folds[tasks[t][1]][0]))
        return cands, folds, tasks, order

    def _scorer(self):
        if callable(self.scoring):
            return self.scoring
        from sklearn.metrics import get_scorer
        return get_scorer(self.scoring or "accuracy")

    def _resolve_backend(self):
        if self.backend != "auto":
            return self.backend
        if int(os.environ.get("WORLD_SIZE", "1")) > 1:
            return "torchrun"
        return "spawn"

    def fit(self, X, y):
        import torch
        y = np.asarray(y.to_array() if hasattr(y, "to_array") else y)
        cands, folds, tasks, order = self._tasks(X, y)
        scorer = self._scorer()
        backend = self._resolve_backend()
        done = self._load_journal(cands, len(folds))
        if done:
            order = [t for t in order if t not in done]
        self._journal_keys = {t: (json.dumps(cands[tasks[t][0]], sort_keys=True, default=str), tasks[t][1]) for t in range(len(tasks))}
        t0 = time.perf_counter()
        if backend == "torchrun":
            results = self._run_torchrun(cands, folds, tasks, order, X, y, scorer)
        else:
            n = self.n_gpus or (torch.cuda.device_count() if torch.cuda.is_available() else 0)
            ppg = self._procs_per_gpu()
            if n <= 1 or backend == "inline":
                if ppg > 1:      # one GPU (the current device) or none, several worker processes on it
                    gpus = [torch.cuda.current_device()] if torch.cuda.is_available() else [None]
                    results = self._run_spawn(gpus, cands, folds, tasks, order, X, y, scorer, ppg)
                else:
                    results = self._run_inline(cands, folds, tasks, order, X, y, scorer)
            else:
                results = self._run_spawn(list(range(n)), cands, folds, tasks, order, X, y, scorer, ppg)
        results.update(done)
        self.search_time_ = time.perf_counter() - t0
        self.n_fits_ = len(tasks)
        self.n_resumed_ = len(done)
        self._collect(cands, folds, tasks, results)
        if self.refit:
            # every rank holds the full result table (the store carries all fits), so best_params_ is
            # the same everywhere; the refit itself runs once, on the main rank.  The other ranks of a
            # torchrun launch get the configured-but-unfitted estimator and ``refit_skipped_ = True``
            # (main.py skips test_model there) instead of an AttributeError on best_estimator_.
            self.best_estimator_ = clone(self.estimator).set_params(**self.best_params_)
            self.refit_skipped_ = not self._is_main()
            if not self.refit_skipped_:
                t1 = time.perf_counter()
                self.best_estimator_.fit(X, y)
                self.refit_time_ = time.perf_counter() - t1
        return self

    def _load_journal(self, cands, n_folds):
        """{task id: result} of fits a previous, interrupted search already finished."""
        done = {}
        if not self.resume_file or not os.path.exists(self.resume_file):
            return done
        index = {json.dumps(c, sort_keys=True, default=str): ci for ci, c in enumerate(cands)}
        with open(self.resume_file) as f:
            for line in f:
                try:
                    rec = json.loads(line)
                except ValueError:
                    continue            # a torn last line of an interrupted run
                ci = index.get(rec.get("params"))
                if ci is not None and 0 <= rec.get("fold", -1) < n_folds:
                    done[ci * n_folds + rec["fold"]] = rec["result"]
        with open(self.resume_file, "rb+") as f:      # terminate a torn last line so that appends start clean
            f.seek(0, os.SEEK_END)
            if f.tell() > 0:
                f.seek(-1, os.SEEK_END)
                if f.read(1) != b"\n":
                    f.write(b"\n")
        return done

    def _journal(self, t, res):
        if not self.resume_file or not self._is_main_process_for_journal():
            return
        params, fold = self._journal_keys[t]
        with open(self.resume_file, "a") as f:
            f.write(json.dumps({"params": params, "fold": fold, "result": res}) + "\n")
            f.flush()

    def _is_main_process_for_journal(self):
        return True

    def _is_main(self):
        return int(os.environ.get("RANK", "0")) == 0 or self._resolve_backend() != "torchrun"

    def _seed(self, ci, fi):
        """Per-fit seed: a function of (random_state, candidate, fold) only."""
        if self.random_state is None:
            return None
        return (int(self.random_state) * 1000003 + ci * 1009 + fi * 7919 + 12345) & 0x7FFFFFFF

    def _sub(self, ci, fi):
        return f"cand{ci:04d}_fold{fi}" if self.per_fit_checkpoint_dirs else None

    def _fits_per_gpu(self):
        k = self.fits_per_gpu if self.fits_per_gpu is not None else int(os.environ.get("SLNLP_FITS_PER_GPU", "1"))
        return max(1, int(k))

    def _procs_per_gpu(self):
        p = self.procs_per_gpu if self.procs_per_gpu is not None else int(os.environ.get("SLNLP_PROCS_PER_GPU", "1"))
        return max(1, int(p))

    def _run_packed(self, k, cands, folds, tasks, order, X, y, scorer):
        """One GPU (or none), k fits at a time: worker threads of THIS process on private streams."""
        import torch
        _pack_env(k)
        task_q, result_q = queue.Queue(), queue.Queue()
        payload = (self.estimator, X, y, scorer)
        gpu = torch.cuda.current_device() if torch.cuda.is_available() else None
        threads = [threading.Thread(target=_worker_loop, args=(gpu, task_q, result_q, payload, True), daemon=True)
                   for _ in range(k)]
        for t in order:
            ci, fi = tasks[t]
            task_q.put((t, cands[ci], folds[fi][0], folds[fi][1], self._sub(ci, fi), self._seed(ci, fi)))
        for _ in threads:
            task_q.put(None)
        for th in threads:
            th.start()
        out = {}
        while len(out) < len(order):
            tid, res, err = self._next_result(result_q, threads)
            if err is not None:
                raise RuntimeError(f"grid-search fit failed (error_score='raise'):\n{err}")
            out[tid] = res
            self._journal(tid, res)
            if self.verbose:
                ci, fi = tasks[tid]
                print(f"[grid] {len(out)}/{len(tasks)} cand {ci} fold {fi}: score {res['score']:.4f} "
                      f"fit {res['fit_time']:.2f}s", flush=True)
        for th in threads:
            th.join(timeout=30)
        return out

    @staticmethod
    def _next_result(result_q, workers, poll=2.0):
        """Next (task id, result, error) from the workers.  A worker that died without reporting
        (OOM kill, CUDA fault, a crash inside the CUDA library) would leave a bare ``get()`` waiting
        forever: poll, and raise once nobody alive is left to produce the missing result or a
        process ended abnormally (the journal allows the search to resume)."""
        while True:
            try:
                return result_q.get(timeout=poll)
            except queue.Empty:
                pass
            codes = [getattr(w, "exitcode", None) for w in workers]
            bad = [c for c in codes if c not in (None, 0)]
            if bad or not any(w.is_alive() for w in workers):
                try:                        # a result may have landed between the timeout and the check
                    return result_q.get(timeout=0.2)
                except queue.Empty:
                    raise RuntimeError(f"grid-search worker(s) ended without reporting (exit codes {codes}); "
                                       "finished fits are in the resume journal") from None

    def _run_inline(self, cands, folds, tasks, order, X, y, scorer):
        k = self._fits_per_gpu()
        if k > 1:
            return self._run_packed(k, cands, folds, tasks, order, X, y, scorer)
        out = {}
        for t in order:
            ci, fi = tasks[t]
            out[t] = _fit_and_score(self.estimator, cands[ci], X, y, folds[fi][0], folds[fi][1], scorer, self._sub(ci, fi),
                                    self._seed(ci, fi))
            self._journal(t, out[t])
            if self.verbose:
                print(f"[grid] {len(out)}/{len(tasks)} cand {ci} fold {fi}: score {out[t]['score']:.4f} "
                      f"fit {out[t]['fit_time']:.2f}s", flush=True)
        return out

    def _run_spawn(self, gpus, cands, folds, tasks, order, X, y, scorer, procs_per_gpu=1):
        """Worker processes fed by one queue: ``procs_per_gpu`` per entry of ``gpus`` (device indices; None = CPU),
        fits_per_gpu threads in each."""
        import torch.multiprocessing as mp
        ctx = mp.get_context("spawn")
        task_q, result_q = ctx.Queue(), ctx.Queue()
        payload = pickle.dumps((self.estimator, X, y, scorer))
        k = self._fits_per_gpu()
        procs = [ctx.Process(target=_worker_main, args=(g, task_q, result_q, payload, k), daemon=True)
                 for g in gpus for _ in range(max(1, procs_per_gpu))]
        for p in procs:
            p.start()
        for t in order:
            ci, fi = tasks[t]
            task_q.put((t, cands[ci], folds[fi][0], folds[fi][1], self._sub(ci, fi), self._seed(ci, fi)))
        for _ in range(len(procs) * k):      # one sentinel per worker thread
            task_q.put(None)
        out = {}
        self.worker_launches_, reported = 0, 0
        try:
            while len(out) < len(order):
                tid, res, err = self._next_result(result_q, procs)
                if err is not None:
                    raise RuntimeError(f"grid-search fit failed (error_score='raise'):\n{err}")
                if tid == -2:          # a worker process that ran out of tasks reports its launch count and leaves
                    self.worker_launches_ += res["launches"]
                    reported += 1
                    continue
                out[tid] = res
                self._journal(tid, res)
                if self.verbose:
                    ci, fi = tasks[tid]
                    print(f"[grid] {len(out)}/{len(tasks)} cand {ci} fold {fi} on gpu {res['gpu']}: "
                          f"score {res['score']:.4f} fit {res['fit_time']:.2f}s", flush=True)
        except BaseException:
            for p in procs:               # a failed search must not leave workers training
                if p.is_alive():
                    p.terminate()
            raise
        finally:
            deadline = time.time() + 30
            while reported < len(procs) and time.time() < deadline and len(out) >= len(order):
                try:                    # the workers' last words (drained before the joins: a child flushes its queue at exit)
                    tid, res, _ = result_q.get(timeout=0.5)
                    if tid == -2:
                        self.worker_launches_ += res["launches"]
                        reported += 1
                except queue.Empty:
                    if not any(p.is_alive() for p in procs):
                        break
            for p in procs:
                p.join(timeout=30)
                if p.is_alive():
                    p.terminate()
        return out

    def _run_torchrun(self, cands, folds, tasks, order, X, y, scorer):
        """SPMD under torchrun: work is claimed with a store counter, results travel through the
        store.  No process-group collective is issued (no NCCL on the grid path).  ``fits_per_gpu``
        worker threads per rank claim from the same counter.  A fit that raises publishes an error
        record under its own key, so every rank re-raises the SAME error instead of timing out on a
        key that never arrives; waiting polls ``store.check`` (no store-timeout dependence: an
        L6/H512 fit at 200 epochs outlasts the 300 s default)."""
        import datetime
        import torch
        import torch.distributed as dist
        rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
        store = getattr(self, "store", None)
        if store is None:
            store = dist.TCPStore(os.environ.get("MASTER_ADDR", "127.0.0.1"), int(os.environ["MASTER_PORT"]) + 17,
                                  world, is_master=(rank == 0), wait_for_workers=True,
                                  timeout=datetime.timedelta(days=7))
            self.store = store
        gen = getattr(self, "_generation", 0)
        self._generation = gen + 1
        key = f"grid{gen}"
        k_threads = self._fits_per_gpu()
        _pack_env(k_threads)
        gpu = torch.cuda.current_device() if torch.cuda.is_available() else None
        failed = threading.Event()

        def claim_loop(own_stream):
            if gpu is not None:
                torch.cuda.set_device(gpu)
            est = clone(self.estimator)
            if gpu is not None and "device" in est.get_params(deep=False):
                est.set_params(device=f"cuda:{gpu}")
            with (_stream_scope() if own_stream else contextlib.nullcontext()):
                while not failed.is_set():
                    k = store.add(f"{key}/next", 1) - 1
                    if k >= len(order):
                        break
                    t = order[k]
                    ci, fi = tasks[t]
                    try:
                        res = _fit_and_score(est, cands[ci], X, y, folds[fi][0], folds[fi][1], scorer, self._sub(ci, fi),
                                             self._seed(ci, fi))
                    except Exception:
                        failed.set()
                        store.set(f"{key}/res/{t}", pickle.dumps({"error": traceback.format_exc(), "rank": rank}))
                        store.set(f"{key}/failed", str(t))
                        return
                    res["gpu"] = rank
                    store.set(f"{key}/res/{t}", pickle.dumps(res))
                    self._journal(t, res)   # every rank appends its own fits (O_APPEND lines, one writer per line)

        if k_threads <= 1:
            claim_loop(False)
        else:
            threads = [threading.Thread(target=claim_loop, args=(True,), daemon=True) for _ in range(k_threads)]
            for th in threads:
                th.start()
            for th in threads:
                th.join()
        def leave(tag):
            # the store lives in rank 0's process: rank 0 stays until every rank has read what it needs
            # (a rank that returns or raises early would otherwise take the store down under the others)
            store.add(f"{key}/{tag}", 1)
            if rank == 0:
                deadline = time.monotonic() + 120.0
                while int(store.add(f"{key}/{tag}", 0)) < world and time.monotonic() < deadline:
                    time.sleep(0.02)

        out = {}
        for t in order:                 # wait until the owning rank has published it (or a failure)
            while not store.check([f"{key}/res/{t}"]):
                if store.check([f"{key}/failed"]):
                    t = int(store.get(f"{key}/failed"))
                    break
                time.sleep(0.05)
            res = pickle.loads(store.get(f"{key}/res/{t}"))
            if "error" in res:
                leave("saw_failure")
                raise RuntimeError(f"grid-search fit failed on rank {res['rank']} (error_score='raise'):\n{res['error']}")
            out[t] = res
        leave("done")
        return out

    # ------------------------------------------------------------------ GridSearchCV result surface
    def _collect(self, cands, folds, tasks, results):
        nc, nf = len(cands), len(folds)
        score = np.full((nc, nf), np.nan)
        fit_t, score_t = np.zeros((nc, nf)), np.zeros((nc, nf))
        for t, (ci, fi) in enumerate(tasks):
            r = results[t]
            score[ci, fi], fit_t[ci, fi], score_t[ci, fi] = r["score"], r["fit_time"], r["score_time"]
        res: Dict[str, object] = {
            "mean_fit_time": fit_t.mean(1), "std_fit_time": fit_t.std(1),
            "mean_score_time": score_t.mean(1), "std_score_time": score_t.std(1),
        }
        keys = sorted({k for c in cands for k in c})
        for k in keys:
            res[f"param_{k}"] = np.ma.MaskedArray([c.get(k) for c in cands], mask=[k not in c for c in cands], dtype=object)
        res["params"] = cands
        for fi in range(nf):
            res[f"split{fi}_test_score"] = score[:, fi]
        res["mean_test_score"] = score.mean(1)
        res["std_test_score"] = score.std(1)
        from scipy.stats import rankdata
        res["rank_test_score"] = rankdata(-res["mean_test_score"], method="min").astype(np.int32)
        self.cv_results_ = res
        self.best_index_ = int(np.argmin(res["rank_test_score"]))
        self.best_score_ = float(res["mean_test_score"][self.best_index_])
        self.best_params_ = cands[self.best_index_]
        self.fit_results_ = results
        self.n_splits_ = nf
