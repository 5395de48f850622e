"""Experiment CLI with the argument surface of the reference's psvi/experiments/flow_psvi.py (flags :52-284, `inf_dict`
keys :306-354, driver :357-456, results files :552-564), running the B200-native PSVI path.

Differences from the reference script, all on the host side: arguments are parsed in main() instead of at import time
(import the module freely); `torch.autograd.set_detect_anomaly(True)` (:50) is not set -- there is no autograd graph;
the json dump is best-effort because PSVI results hold numpy arrays (the reference raises TypeError after writing the
pickle).  Methods outside the hot path resolve to callables that raise NotImplementedError."""
from __future__ import annotations

import argparse
import json
import os
import pickle
from collections import defaultdict
from typing import Any, Dict, List

from psvi.experiments.experiments_utils import read_dataset
from psvi.inference.baselines import (run_giga, run_mfvi, run_mfvi_regressor, run_mfvi_subset,
                                      run_mfvi_subset_regressor, run_opsvi, run_random, run_sparsevi)
from psvi.inference.psvi_classes import (PSVI, PSVI_Ablated, PSVI_No_IW, PSVI_No_Rescaling, PSVI_regressor, PSVIAFixedU,
                                         PSVIAV, PSVIAV_regressor, PSVIFixedU, PSVIFreeV, PSVILearnV,
                                         PSVILearnV_regressor)


def build_parser() -> argparse.ArgumentParser:
    parser = argparse.ArgumentParser()
    B = argparse.BooleanOptionalAction
    parser.add_argument("--fnm", default="results", type=str, help="Filename where results are stored")
    parser.add_argument("--datasets", default=["phishing"], nargs="+", type=str, help="List of dataset names",
                        choices=["webspam", "phishing", "adult", "MNIST", "halfmoon", "four_blobs", "sinus", "concrete",
                                 "energy", "power", "kin8nm", "protein", "naval", "yacht", "boston", "wine", "year",
                                 "synth_lr_10", "synth_lr_50", "synth_lr_200"])
    parser.add_argument("--methods", default=["psvi_learn_v", "mfvi", "mfvi_subset"], nargs="+", type=str,
                        help="List of inference method names")
    parser.add_argument("--mc_samples", default=10, type=int, help="Monte Carlo samples")
    parser.add_argument("--num_epochs", default=301, type=int, help="Training epochs")
    parser.add_argument("--num_trials", default=3, type=int, help="Trials executed for each inference method")
    parser.add_argument("--data_minibatch", default=128, type=int, help="Data minibatch size")
    parser.add_argument("--inner_it", default=100, type=int,
                        help="Gradient steps in the inner problem of nested optimization")
    parser.add_argument("--outer_it", default=100, type=int,
                        help="Gradient steps in the outer problem of nested optimization")
    parser.add_argument("--trainer", default="nested", choices=["nested", "hyper", "joint"], type=str,
                        help="Method for computation of hypergradient")
    parser.add_argument("--diagonal", action=B, help="Diagonal approximation of Gaussian covariance matrices used")
    parser.add_argument("--architecture", default="logistic_regression", type=str, help="Model architecture",
                        choices=["logistic_regression", "logistic_regression_fullcov", "fn", "fn2", "lenet",
                                 "regressor_net"])
    parser.add_argument("--n_hidden", default=40, type=int,
                        help="Number of hidden units in feedforward neural architectures")
    parser.add_argument("--n_layers", default=1, type=int, help="Number of layers in feedforward neural architectures")
    parser.add_argument("--log_every", default=150, type=int,
                        help="Frequency of logging evaluation results throughout training (in number of outer gradient iterations)")
    parser.add_argument("--register_elbos", action=B,
                        help="Saving variational objectives values throughout inference for plotting")
    parser.add_argument("--init_sd", default=1e-6, type=float,
                        help="Initialization of standard deviation for variational parameters")
    parser.add_argument("--lr0net", default=1e-3, type=float, help="Initial learning rate for model parameters optimizer")
    parser.add_argument("--lr0u", default=1e-4, type=float,
                        help="Initial learning rate for optimizer of pseudocoreset point input coordinates u")
    parser.add_argument("--lr0v", default=1e-3, type=float,
                        help="Initial learning rate for optimizer of coreset support coefficients")
    parser.add_argument("--lr0z", default=1e-3, type=float,
                        help="Initial learning rate for optimizer of coreset points labels")
    parser.add_argument("--lr0alpha", default=1e-3, type=float,
                        help="Initial learning rate for coreset likelihood rescaling coefficient")
    parser.add_argument("--init_at", default="subsample", choices=["subsample", "random"], type=str,
                        help="Method for coreset points initialization")
    parser.add_argument("--compute_weights_entropy", action=B, help="Comput entropy of weights for plotting")
    parser.add_argument("--coreset_sizes", default=[100], nargs="+", type=int,
                        help="List of sizes for coresets computed throughout the experiment, or subsamples used for baselines mfvi_subset and random")
    parser.add_argument("--reset", action=B, help="Reset model parameters over intervals during training")
    parser.add_argument("--prune", action=B, help="Prune to coreset of smaller size")
    parser.add_argument("--prune_interval", default=400, type=int,
                        help="Gradient steps in the outer problem of nested optimization between prunning steps")
    parser.add_argument("--prune_sizes", default=[20], nargs="+", type=int,
                        help="List of sizes for coresets in a pruning experiment (decreasing)")
    parser.add_argument("--increment", action=B, help="Learn tasks incrementally")
    parser.add_argument("--increment_interval", default=1000, type=int,
                        help="Gradient steps in the outer problem of nested optimization between incremental learning stages")
    parser.add_argument("--increment_sizes", default=[20], nargs="+", type=int,
                        help="List of sizes for coresets in the incremental learning setting (non-decreasing)")
    parser.add_argument("--retrain_on_coreset", action=B,
                        help="Retrain the variational model restricted only on the extracted coreset datapoints for the same number of epochs")
    parser.add_argument("--save_input_data", action=B, help="Save input dataset")
    parser.add_argument("--test_ratio", default=0.2, type=float, help="Ratio of test dataset size")
    parser.add_argument("--log_pseudodata", action=B, help="Store pseudodata for visualisation")
    parser.add_argument("--data_folder", default="../data", type=str, help="Folder where dataset gets stored")
    parser.add_argument("--results_folder", default="../results", type=str,
                        help="Folder where evaluation files get stored")
    parser.add_argument("--learn_z", action=B, help="Learn soft labels for distilled data")
    parser.add_argument("--gamma", default=1.0, type=float, help="Decay factor of learning rate")
    parser.set_defaults(diagonal=True, reset=False, compute_weights_entropy=False, register_elbos=False,
                        save_input_data=False, prune=False, increment=False, log_pseudodata=False,
                        retrain_on_coreset=False, learn_z=False)
    return parser


parser = build_parser()


def rec_dd():
    return defaultdict(rec_dd)


def _method(cls):
    return lambda *args, **kwargs: cls(*args, **kwargs).run_psvi(*args, **kwargs)


from psvi.inference.sparsebbvi import run_sparsevi_with_bb_elbo as _sparsebbvi  # noqa: E402


# Inference methods (reference :306-354): every key resolves
inf_dict = {
    "psvi": _method(PSVI),
    "psvi_ablated": _method(PSVI_Ablated),
    "psvi_learn_v": _method(PSVILearnV),
    "psvi_alpha_v": _method(PSVIAV),
    "psvi_no_iw": _method(PSVI_No_IW),
    "psvi_free_v": _method(PSVIFreeV),
    "psvi_no_rescaling": _method(PSVI_No_Rescaling),
    "psvi_fixed_u": _method(PSVIFixedU),
    "psvi_alpha_fixed_u": _method(PSVIAFixedU),
    "psvi_regressor": _method(PSVI_regressor),
    "psvi_alpha_v_regressor": _method(PSVIAV_regressor),
    "psvi_learn_v_regressor": _method(PSVILearnV_regressor),
    "sparsebbvi": _sparsebbvi,
    "opsvi": run_opsvi,
    "random": run_random,
    "sparsevi": run_sparsevi,
    "giga": run_giga,
    "mfvi": run_mfvi,
    "mfvi_subset": run_mfvi_subset,
    "mfvi_regressor": run_mfvi_regressor,
    "mfvi_subset_regressor": run_mfvi_subset_regressor,
}


def experiment_driver(datasets: List[str], methods: List[str], method_args: Dict[str, Any], results=None):
    """Run the experiment grid datasets x methods x trials x coreset sizes (reference :357-456)."""
    results = rec_dd() if results is None else results
    for dnm in datasets:
        print(f"\nReading/Generating the dataset {dnm.upper()}")
        x, y, xt, yt, N, D, train_dataset, test_dataset, num_classes = read_dataset(dnm, method_args)
        print(f"Details about dataset: Total Datapoints: {N}, Dimensions: {D}, num classes: {num_classes}")
        for nm_alg in methods:
            print(f"\n\nRunning {nm_alg}\n")
            logistic_regression = method_args.get("logistic_regression", method_args.get("architecture") == "logreg")
            inf_alg = inf_dict[nm_alg]
            compute_weights_entropy = (not nm_alg.startswith(("opsvi", "mfvi_subset"))) and method_args[
                "compute_weights_entropy"]
            tps = method_args["coreset_sizes"] if nm_alg.startswith(("psvi", "opsvi", "mfvi_subset")) else [-1]
            for t in range(method_args["num_trials"]):
                print(f"Trial #{t}")
                for ps in tps:
                    results[dnm][nm_alg][ps][t] = inf_alg(
                        mc_samples=method_args["mc_samples"], num_epochs=method_args["num_epochs"],
                        data_minibatch=method_args["data_minibatch"], D=D, N=N, tr=t, diagonal=method_args["diagonal"],
                        x=x, y=y, xt=xt, yt=yt, inner_it=method_args["inner_it"], outer_it=method_args["outer_it"],
                        scatterplot_coreset=method_args.get("scatterplot_coreset"),
                        logistic_regression=logistic_regression, trainer=method_args["trainer"],
                        log_every=method_args["log_every"], register_elbos=method_args["register_elbos"],
                        lr0u=method_args["lr0u"], lr0net=method_args["lr0net"], lr0v=method_args["lr0v"],
                        lr0z=method_args["lr0z"], lr0alpha=method_args["lr0alpha"], init_args=method_args["init_at"],
                        init_sd=method_args["init_sd"], num_pseudo=ps, seed=t,
                        compute_weights_entropy=compute_weights_entropy, reset=method_args.get("reset"),
                        reset_interval=method_args.get("reset_interval"), architecture=method_args.get("architecture"),
                        log_pseudodata=method_args.get("log_pseudodata"), n_hidden=method_args.get("n_hidden", 40),
                        n_layers=method_args.get("n_layers", 1), train_dataset=train_dataset,
                        test_dataset=test_dataset, dnm=dnm, nc=num_classes, prune=method_args.get("prune"),
                        prune_interval=method_args.get("prune_interval"), prune_sizes=method_args.get("prune_sizes"),
                        increment=method_args.get("increment"),
                        increment_interval=method_args.get("increment_interval"),
                        increment_sizes=method_args.get("increment_sizes"),
                        retrain_on_coreset=method_args.get("retrain_on_coreset"), learn_z=method_args["learn_z"],
                        gamma=method_args.get("gamma", 1.0))
                    print("Trial completed!\n")
    return write_to_files(results, method_args["fnm"], method_args["results_folder"])


def _plain(o, json_safe=False):
    """defaultdict -> dict (pickle keeps the reference's keys); json_safe also stringifies keys and unpacks arrays."""
    if isinstance(o, dict):
        return {(str(k) if json_safe else k): _plain(v, json_safe) for k, v in o.items()}
    if json_safe and isinstance(o, (list, tuple)):
        return [_plain(v, json_safe) for v in o]
    if json_safe and hasattr(o, "tolist"):
        return o.tolist()
    return o


def write_to_files(results: Dict[str, Any], fnm: str, results_folder: str = "../results"):
    """results[dnm][method][size][trial] -> {results_folder}/{fnm}.pk (+ best-effort .json); reference :552-564."""
    res_fnm = f"{results_folder}/{fnm}.pk"
    print(f"Storing results in {res_fnm}")
    with open(res_fnm, "wb") as outfile:
        pickle.dump(_plain(results), outfile)
    try:
        with open(f"{results_folder}/{fnm}.json", "w") as fp:
            json.dump(_plain(results, json_safe=True), fp)
    except TypeError:
        pass
    return results


def main(argv=None):
    method_args = vars(parser.parse_args(argv))
    method_args["logistic_regression"] = method_args["architecture"] == "logistic_regression"
    for fold in (method_args["data_folder"], method_args["results_folder"]):
        os.makedirs(fold, exist_ok=True)
    if method_args.get("architecture") == "regressor_net":
        raise NotImplementedError("the regression flow reads UCI benchmark files that need a download (no network here); the "
                                  "regressor classes are built: psvi.inference.psvi_classes.PSVI_regressor / "
                                  "PSVILearnV_regressor / PSVIAV_regressor with (train, val, test) BaseDatasets, y_mean, y_std, tau")
    return experiment_driver(method_args["datasets"], method_args["methods"], method_args)


if __name__ == "__main__":
    main()
