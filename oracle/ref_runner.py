"""Drive the UNMODIFIED reference's PSVI hot path for timing (TEST / BENCH INFRASTRUCTURE ONLY -- never imported by the
product; bench.py uses it for the `--impl reference` arm and the `cpu_baseline` leg only).

The reference tree is found by oracle.ref_import (baseline/_ref, the `pip install --target` copy that travels to the GPU
box, or /root/reference in the build container).  What is timed is the reference's own public call sequence for one
outer step (psvi/inference/psvi_classes.py:895-927):  xbatch, ybatch = next(iter(train_loader)) ; nested_step(xbatch,
ybatch) ; loss.item().  Nothing of this repo's kernels, engine or oracle port is on that path.
"""
from __future__ import annotations

import contextlib
import io
import os
import time


def _quiet():
    return contextlib.redirect_stdout(io.StringIO())


def build_reference_chain(cfg, device="cpu", seed=0):
    """PSVILearnV of the reference, set up by its own run_psvi(num_epochs=0) on `device` ("cpu" | "cuda")."""
    from oracle.ref_import import import_reference
    import_reference()
    import numpy as np
    import torch
    from psvi.experiments.experiments_utils import SynthDataset
    from psvi.inference.psvi_classes import PSVILearnV
    from sklearn.datasets import make_moons

    # the reference's halfmoon generator (experiments_utils.py:759-767,796-804): make_moons(1000, noise=.1, seed 42), 80/20
    X, Y = make_moons(n_samples=1000, noise=0.1, random_state=42)
    X, Y = torch.from_numpy(X.astype(np.float32)), torch.from_numpy(Y.astype(np.float32))
    N = cfg["N"]
    tr, te = SynthDataset(X[:N], Y[:N]), SynthDataset(X[N:], Y[N:])
    kw = dict(mc_samples=cfg["S"], num_epochs=0, data_minibatch=cfg["B"], D=cfg["D"], N=N, inner_it=cfg["T"],
              trainer="nested", log_every=150, lr0u=1e-4, lr0net=1e-3, lr0v=1e-3, init_args="subsample",
              init_sd=cfg["init_sd"], num_pseudo=cfg["M"], seed=seed, architecture="fn", n_hidden=cfg["H"], n_layers=1,
              logistic_regression=False, train_dataset=tr, test_dataset=te, dnm="halfmoon", nc=cfg["C"],
              data_folder="/tmp/psvi_data", compute_weights_entropy=False, register_elbos=False)
    real = torch.cuda.is_available
    if device == "cpu":
        torch.cuda.is_available = lambda: False     # psvi_classes.py:141 picks "cuda" whenever it is available
    try:
        with _quiet(), contextlib.redirect_stderr(io.StringIO()):
            obj = PSVILearnV(**kw)
            obj.run_psvi(**kw)
    finally:
        torch.cuda.is_available = real
    assert obj.device.type == device, (obj.device, device)
    return obj


def time_reference(cfg, steps, warmup, device="cpu", anomaly=False, threads=None, budget_s=None):
    """steps/s of the reference's outer step.  `anomaly`: torch.autograd.set_detect_anomaly(True), which is how
    flow_psvi.py ships (flow_psvi.py:50).  `budget_s`: stop early once that much wall time was spent (bounded sample)."""
    import torch
    if threads:
        torch.set_num_threads(int(threads))
    obj = build_reference_chain(cfg, device=device)
    torch.autograd.set_detect_anomaly(bool(anomaly))
    sync = torch.cuda.synchronize if device == "cuda" else (lambda: None)

    def one():
        with _quiet():
            xb, yb = next(iter(obj.train_loader))
            loss = obj.nested_step(xb.to(obj.device), yb.to(obj.device))
            return float(loss.item())
    try:
        for _ in range(warmup):
            one()
        sync()
        t0 = time.perf_counter()
        done = 0
        for _ in range(steps):
            one()
            done += 1
            if budget_s is not None and time.perf_counter() - t0 > budget_s:
                break
        sync()
        dt = time.perf_counter() - t0
    finally:
        torch.autograd.set_detect_anomaly(False)
    from oracle.ref_import import reference_root
    return {"steps_per_s": done / dt, "s_per_step": dt / done, "steps": done, "threads": torch.get_num_threads(),
            "root": reference_root(),
            "cores": os.cpu_count(), "device": device, "anomaly": bool(anomaly)}
