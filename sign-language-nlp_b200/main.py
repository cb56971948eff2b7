"""Entry point with the reference's surface (main.py:12-143): ``run`` -> ``tune_hyperparams``
(grid search over independent fits, farmed over the GPUs) -> ``test_model`` (scorers + one
profiled predict).  Usage:

    python main.py --config config-enc-dec-lstm-attn.yaml --cuda True [--gpus 8] [--precision bf16]
"""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
if HERE not in sys.path:
    sys.path.insert(0, HERE)

import helper as h  # noqa: E402
from args import ARGUMENTS, load_args  # noqa: E402
from helper import log  # noqa: E402
from slnlp_b200.grid import GridSearchFarm  # noqa: E402
from slnlp_b200.net import NeuralNetClassifier  # noqa: E402


def run(args):
    seed = args["seed"]
    h.setup_seed(seed)
    device = h.prepare_device(args["cuda"])
    dataset = h.load_dataset(**args).stoi()
    if args["debug"]:
        dataset = dataset.truncated(args["cv"] * 10)
    if should_balance_dataset(args):
        dataset = h.balance_dataset(dataset=dataset, seed=seed)
    log(f"{len(dataset)} entries of data")
    callbacks, callbacks_names = h.build_callbacks(dataset=dataset, **args)
    net_params = h.build_net_params(callbacks=callbacks, callbacks_names=callbacks_names, device=device,
                                    dataset=dataset, **args)
    net = NeuralNetClassifier(**net_params)
    test_data, train_data = dataset.split(lengths=args["test_size"], indices_only=False, seed=seed)
    log(f"> Train data: {len(train_data)} entries")
    log(f"> Test data: {len(test_data)} entries")
    best_estimator = tune_hyperparams(estimator=net, callbacks_names=callbacks_names, train_data=train_data, **args)
    if best_estimator is None:      # a non-main rank of a torchrun launch: the main rank refits, tests and writes
        return None
    return test_model(estimator=best_estimator, test_data=test_data, **args)


def tune_hyperparams(estimator, callbacks_names, train_data, cuda, gpus=None, **kwargs):
    log("\n==================== TUNING HYPERPARAMETERS ====================\n")
    phase = "grid_search"
    gs_params = h.build_grid_params(callbacks_names=callbacks_names, data=train_data, **kwargs)
    gs = GridSearchFarm(estimator=estimator, n_gpus=gpus, **gs_params)
    log(gs_params)
    h.save_param_grid(gs.param_grid, phase=phase, **kwargs)
    gs.fit(X=train_data.X(), y=train_data.y().to_array())
    if getattr(gs, "refit_skipped_", False):
        return None
    gs_output = {"best_score": float(gs.best_score_), "best_params": gs.best_params_, "best_index": int(gs.best_index_),
                 "scoring": str(gs.scoring), "n_fits": gs.n_fits_, "search_seconds": gs.search_time_,
                 "fits_per_hour": 3600.0 * gs.n_fits_ / gs.search_time_}
    h.save_output(gs_output, phase=phase, **kwargs)
    h.save_cv_results(gs.cv_results_, phase=phase, **kwargs)
    return gs.best_estimator_


def test_model(estimator, test_data, scoring, cuda, **kwargs):
    log("\n==================== TESTING MODEL ====================\n")
    phase = "test"
    if "accuracy" not in scoring:
        scoring = ["accuracy", *scoring]
    scorers = h.build_scoring(scoring=scoring, labels=test_data.labels())
    test_output = {f"test_{s.score}": s(estimator, test_data.X(), test_data.y().to_array()) for s in scorers}
    with h.create_profiler(cuda) as prof:
        estimator.predict(test_data.X())
    h.save_output(test_output, phase=phase, **kwargs)
    h.save_profile(prof, phase=phase, **kwargs)
    return test_output


def should_balance_dataset(args):
    return bool(args["dataset_args"].get("balance_dataset", False))


if __name__ == "__main__":
    args = vars(load_args("SL Transformer", ARGUMENTS))
    args["workdir"] = h.format_dir(args["workdir"], **args)
    h.dump_args(args)
    args["gpus"] = h.create_worker_farm(**args)
    run(args)
