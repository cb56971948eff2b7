"""CLI / YAML schema of the run (the reference's args.py:3-53 consumed by
``commons.util.load_args``, main.py:131-132; commons is not vendored, so the loader lives here).

Every key can come from the YAML file given with ``--config`` (config/*.yaml of the reference
load unchanged) and be overridden on the command line; dict-typed options take YAML/JSON text.
Keys the reference's YAMLs leave out get the defaults commons would supply: ``mode`` "grid",
``criterion_args`` {}, ``dask_args`` {} (there is no Dask here; ``gpus`` sizes the worker farm).
"""
import argparse
from collections import namedtuple

import yaml

Argument = namedtuple("Argument", "flag name type default help required options")


def _arg(flag, name, type=str, default=None, help="", required=False, options=None):
    return Argument(flag, name, type, default, help, required, options)


def _bool(s):
    return s if isinstance(s, bool) else str(s).lower() in ("1", "true", "yes", "y")


ARGUMENTS = [
    _arg("-m", "--model", help="Model class"),
    _arg("-o", "--optimizer", help="Optimizer class"),
    _arg("-f", "--criterion", help="Criterion class"),
    _arg("-cv", "--cv", type=int, help="Cross-validation folds"),
    _arg("-sc", "--scoring", type=yaml.safe_load, help="Scoring metric(s) to use"),
    _arg("-vb", "--verbose", type=int, help="Verbosity level"),
    _arg("-j", "--n_jobs", type=int, default=1, help="Number of jobs"),
    _arg("-n", "--mode", options=["grid", "train"], default="grid", help="Mode"),
    _arg("-w", "--workdir", help="Working directory"),
    _arg("-d", "--debug", type=_bool, default=False, help="Debug flag"),
    _arg("-nv", "--cuda", type=_bool, default=False, help="Enable cuda"),
    _arg("-k", "--seed", type=int, required=True, help="Seed"),
    _arg("-lr", "--lr", type=float, required=True, help="Learning rate"),
    _arg("-ep", "--max_epochs", type=int, required=True, help="Max epochs"),
    _arg("-bs", "--batch_size", type=int, required=True, help="Batch size"),
    _arg("-ts", "--test_size", type=float, required=True, help="Test size"),
    _arg("-es", "--early_stopping", type=dict, help="Options for early stopping"),
    _arg("-gcl", "--gradient_clipping", type=dict, help="Options for gradient clipping"),
    _arg("-lrs", "--lr_scheduler", type=dict, help="Options for learning rate scheduler"),
    _arg("-ds", "--dataset_args", type=dict, default={}, help="Options for the dataset"),
    _arg("-ma", "--model_args", type=dict, default={}, help="Options for the model"),
    _arg("-oa", "--optimizer_args", type=dict, default={}, help="Options for the optimizer"),
    _arg("-ca", "--criterion_args", type=dict, default={}, help="Options for the criterion"),
    _arg("-gr", "--grid_args", type=dict, default={}, help="Options for the grid search"),
    _arg("-dask", "--dask_args", type=dict, default={}, help="Accepted for CLI compatibility; unused (no Dask)"),
    # B200 additions
    _arg("-g", "--gpus", type=int, default=None, help="GPUs (worker processes) for the grid search; default all"),
    _arg("-p", "--precision", options=["fp32", "bf16"], default="fp32", help="fp32 FMA path or bf16 tcgen05 path"),
]


def load_args(description, arguments=ARGUMENTS, argv=None):
    ap = argparse.ArgumentParser(description=description)
    ap.add_argument("-c", "--config", help="YAML configuration file")
    for a in arguments:
        kw = dict(dest=a.name.lstrip("-"), help=a.help, default=None)
        kw["type"] = yaml.safe_load if a.type is dict else a.type
        if a.options:
            kw["choices"] = a.options
        ap.add_argument(a.flag, a.name, **kw)
    ns = vars(ap.parse_args(argv))
    cfg = {}
    if ns.get("config"):
        with open(ns["config"]) as f:
            cfg = yaml.safe_load(f) or {}
    out = {}
    for a in arguments:
        key = a.name.lstrip("-")
        if ns.get(key) is not None:
            out[key] = ns[key]
        elif key in cfg and (cfg[key] is not None or not a.required):
            out[key] = cfg[key]
        else:
            out[key] = a.default
        # YAML writes 1e-4 style floats as strings inside nested dicts; normalise numerics
        if isinstance(out[key], dict):
            out[key] = _numeric(out[key])
    for k, v in cfg.items():                       # unknown YAML keys pass through, as commons does
        out.setdefault(k, v)
    # `lr` is "tuned in grid search" (empty) in the shipped YAMLs; required only when nothing tunes it
    missing = [a.name for a in arguments if a.required and out.get(a.name.lstrip("-")) is None
               and not (a.name == "--lr" and "lr" in (out.get("grid_args") or {}))]
    if missing:
        ap.error("missing required option(s): " + ", ".join(missing))
    out["config"] = ns.get("config")
    return argparse.Namespace(**out)


def _numeric(d):
    def conv(v):
        if isinstance(v, dict):
            return _numeric(v)
        if isinstance(v, list):
            return [conv(x) for x in v]
        if isinstance(v, str):
            try:
                return float(v) if any(c in v for c in ".eE") else int(v)
            except ValueError:
                return v
        return v
    return {k: conv(v) for k, v in d.items()}
