"""Mean-field VI baselines of the BNN flows on the B200 path: run_mfvi_subset (reference psvi/inference/baselines.py
:923-1062) and run_mfvi (:824-920).  Same keyword surface and results dict; training runs through psvi_mf_unroll
(fused sampled forward + backward + torch.optim.Adam arithmetic, Adam moments resident on the device) and the test
loop through psvi_mf_evaluate in its mean-of-logits mode.  The remaining functions of the reference file (Laplace /
GIGA / SparseVI / OPSVI logistic-regression coresets, k-means / EL2N selection, regression MFVI) are different
algorithms on CPU tensors and out of scope."""
from __future__ import annotations

import random
import time
from typing import Any, Dict

import numpy as np
import torch

from psvi import _native
from psvi.experiments.experiments_utils import set_up_model
from psvi.inference.utils import pseudo_rand_init, pseudo_subsample_init
from psvi.models.neural_net import FullCovMLP, MeanFieldLeNet, MeanFieldMLP, categorical_fn


def _not_built(name, where):
    def f(*a, **k):
        raise NotImplementedError(f"{name} ({where}) is outside the PSVI hot-path scope (SURVEY.md section 2 row 5b)")
    f.__name__ = name
    return f


run_random = _not_built("run_random", "baselines.py:35")
run_giga = _not_built("run_giga", "baselines.py:178")
run_sparsevi = _not_built("run_sparsevi", "baselines.py:330")
run_opsvi = _not_built("run_opsvi", "baselines.py:560")


class _Trainer:
    """Device-resident state of one mean-field VI fit."""

    def __init__(self, net, seed, noise_source=None):
        if not isinstance(net, MeanFieldMLP):
            raise NotImplementedError("the CUDA path covers mean-field MLPs (logistic_regression, fn)")
        _native.require_cuda()
        net.check_supported()
        self.net, self.seed, self.noise_source, self.domain = net, seed, noise_source, 0
        self.device = next(net.parameters()).device
        self.desc = _native.make_model(net.dims, net.n_samples())
        self.mu, self.rho = net.flat()
        P = self.mu.numel()
        self.am, self.av = torch.zeros(2 * P, device=self.device), torch.zeros(2 * P, device=self.device)
        self.steps = 0
        self._scratch = None

    def noise(self, n):
        if self.noise_source is not None:
            return _native.make_noise(self.noise_source.take(n, self.device))
        self.domain += 1
        return _native.make_noise(None, seed=self.seed, domain=self.domain)

    def train(self, x, y32, scale, T, lr):
        """T Adam steps on  -scale * sum_{s,m} log p + sum KL  (baselines.py:1023-1030); returns the T losses."""
        losses = torch.zeros(T, device=self.device)
        roww = torch.full((x.shape[0],), float(scale), device=self.device)
        _native.unroll(self.desc, self.noise(T), self.mu, self.rho, self.am, self.av, self.steps, x, y32, roww, None,
                       1.0, 0, 0.0, T, lr, _native.ADAM_TORCH, losses)
        self.steps += T
        return losses

    def test(self, xt, yt32, batch):
        """(acc, nll) with test_logits = net(xt).mean(0) per batch (baselines.py:1035-1043)."""
        n = xt.shape[0]
        need = _native.eval_scratch_floats(self.desc, n, batch)
        if self._scratch is None or self._scratch.numel() < need:
            self._scratch = torch.zeros(need, device=self.device)
        out = torch.zeros(8, device=self.device)
        _native.evaluate(self.desc, self.noise(-(-n // batch)), self.mu, self.rho, None, None, None, xt, yt32, batch, 0,
                         1.0, 0, 0.0, 2, out, self._scratch)
        o = out.cpu()
        return (o[1] / o[2]).item(), (o[0] / o[2]).item()


class _GaussTrainer:
    """Mean-field VI of a one-output regressor net under the Gaussian likelihood (reference baselines.py:1283-1346, `fit`): the
    per-sample passes are psvi_net_pass_gaussian under the streaming engine, torch.optim.Adam arithmetic on the flat (mu, rho)."""

    def __init__(self, net, tau, seed, noise_source=None):
        from psvi.inference.stream import GaussMlpNet, MeanFieldFamily, StreamEngine
        if not isinstance(net, MeanFieldMLP) or net.dims[-1] != 1:
            raise NotImplementedError("the regression baselines run a mean-field regressor_net with one output")
        _native.require_cuda()
        net.check_supported()
        self.S = net.n_samples()
        self.eng = StreamEngine(MeanFieldFamily(net), net.dims, self.S, net=GaussMlpNet(net.dims, self.S, tau))
        self.net, self.seed, self.noise_source, self.domain = net, seed, noise_source, 0
        self.device = next(net.parameters()).device
        self.phi = self.eng.fam.get_phi().clone().contiguous()
        self.m, self.v, self.steps = torch.zeros_like(self.phi), torch.zeros_like(self.phi), 0

    _eps = None      # (bound below: same noise plumbing as _StreamTrainer)

    def step(self, x, y, scale, lr):
        """One Adam step on  -scale * sum_{s,r} log N(y_r | o_s(x_r), 1/tau) + sum KL; returns the loss (a 0-dim tensor)."""
        eng, fam, R = self.eng, self.eng.fam, x.shape[0]
        eps = fam.fix_eps(self._eps(1))[0]
        theta = fam.sample(self.phi, eps)
        nll, tbar = torch.empty(self.S, R, device=self.device), torch.empty(self.S, eng.Pt, device=self.device)
        eng.net.pass_(theta, None, x, y, torch.full((self.S, R), float(scale), device=self.device), nll=nll, tbar=tbar)
        loss = (float(scale) * nll.double().sum() + fam.kl(self.phi).double()).float()
        g = fam.grad(self.phi, eps, tbar, 1.0, 0.0)
        self.steps += 1
        self.m = 0.9 * self.m + (1.0 - 0.9) * g
        self.v = 0.999 * self.v + (1.0 - 0.999) * g * g
        den = self.v.sqrt() / (1.0 - 0.999 ** self.steps) ** 0.5 + 1e-8
        self.phi = self.phi - (lr / (1.0 - 0.9 ** self.steps)) * self.m / den
        fam.set_phi(self.phi)
        return loss

    def outputs(self, x):
        """net(x) with a fresh noise draw: [S, R]."""
        theta = self.eng.fam.sample(self.phi, self.eng.fam.fix_eps(self._eps(1))[0])
        return self.eng.net.logits(theta, x)[..., 0]


class _StreamTrainer:
    """The same fit for the model families without a flat (mu, rho) cluster-engine form -- fn2 (full covariance) and lenet --
    through the streaming engine's per-sample network kernels (psvi/inference/stream.py), torch.optim.Adam arithmetic on the
    flat variational vector.  Quirk Q5 (reference baselines.py:1027): the KL sum filters on VILinear only, so fn2 trains
    WITHOUT a KL term and lenet's conv layers contribute none."""

    def __init__(self, net, seed, noise_source=None):
        from psvi.inference.stream import FullCovFamily, LenetFamily, LenetNet, StreamEngine
        _native.require_cuda()
        net.check_supported()
        S = net.n_samples()
        if isinstance(net, MeanFieldLeNet):
            self.eng, self.kl_coef = StreamEngine(LenetFamily(net), net.dims, S, net=LenetNet(S)), 1.0
        else:
            self.eng, self.kl_coef = StreamEngine(FullCovFamily(net), net.dims, S), 0.0
        self.net, self.seed, self.noise_source, self.domain, self.S = net, seed, noise_source, 0, S
        self.device = next(net.parameters()).device
        self.phi = self.eng.fam.get_phi().clone().contiguous()
        self.m, self.v, self.steps = torch.zeros_like(self.phi), torch.zeros_like(self.phi), 0

    def _eps(self, n):
        if self.noise_source is not None:
            return self.noise_source.take(n, self.device)
        self.domain += 1
        eps = torch.empty(n, self.S, self.eng.Pt, device=self.device)
        _native.philox_normal(self.seed, self.domain, 0, n, self.S, self.eng.Pt, eps)
        return eps

    def train(self, x, y32, scale, T, lr):
        eng, fam, M = self.eng, self.eng.fam, x.shape[0]
        cw = torch.full((self.S, M), float(scale), device=self.device)
        eps_all = fam.fix_eps(self._eps(T))
        losses = torch.zeros(T, device=self.device)
        for t in range(T):
            theta = fam.sample(self.phi, eps_all[t])
            nll, tbar = torch.empty(self.S, M, device=self.device), torch.empty(self.S, eng.Pt, device=self.device)
            eng.net.pass_(theta, None, x, y32, cw, nll=nll, tbar=tbar)
            losses[t] = (float(scale) * nll.double().sum() + self.kl_coef * fam.kl(self.phi).double()).float()
            g = fam.grad(self.phi, eps_all[t], tbar, self.kl_coef, 0.0)
            self.steps += 1
            self.m = 0.9 * self.m + (1.0 - 0.9) * g
            self.v = 0.999 * self.v + (1.0 - 0.999) * g * g
            den = self.v.sqrt() / (1.0 - 0.999 ** self.steps) ** 0.5 + 1e-8
            self.phi = self.phi - (lr / (1.0 - 0.9 ** self.steps)) * self.m / den
        fam.set_phi(self.phi)
        return losses

    def test(self, xt, yt32, batch):
        n = xt.shape[0]
        out = self.eng.evaluate(self.phi, self._eps(-(-n // batch)), None, None, None, xt, yt32, batch, mode=2)
        o = out.cpu()
        return (o[1] / o[2]).item(), (o[0] / o[2]).item()


_GaussTrainer._eps = _StreamTrainer._eps


def _make_trainer(net, seed, noise_source):
    if isinstance(net, MeanFieldMLP) and not isinstance(net, MeanFieldLeNet):
        return _Trainer(net, seed, noise_source)
    if isinstance(net, (FullCovMLP, MeanFieldLeNet)):
        return _StreamTrainer(net, seed, noise_source)
    raise NotImplementedError(f"no CUDA path for {type(net).__name__}")


def run_mfvi_subset(x=None, y=None, xt=None, yt=None, mc_samples=4, data_minibatch=128, num_epochs=100, log_every=10,
                    D=None, lr0net=1e-3, mul_fact=2, seed=0, distr_fn=categorical_fn, log_pseudodata=False,
                    train_dataset=None, test_dataset=None, num_pseudo=100, init_args="subsample", architecture=None,
                    n_hidden=None, nc=2, dnm=None, init_sd=None, noise_source=None, **kwargs) -> Dict[str, Any]:
    """Mean-field VI on a random class-balanced subset (reference baselines.py:923-1062)."""
    device = torch.device("cuda" if torch.cuda.is_available() else "cpu")
    random.seed(seed), np.random.seed(seed), torch.manual_seed(seed)
    nlls, accs, times, elbos = [], [], [0], []
    t_start = time.time()
    net = set_up_model(architecture=architecture, D=D, n_hidden=n_hidden, nc=nc, mc_samples=mc_samples,
                       init_sd=init_sd).to(device)
    if dnm == "MNIST" and not torch.is_tensor(x):
        raise NotImplementedError("vision datasets need a download (no network; SURVEY.md section 2 row 9): pass the images "
                                  "as tensors x [N, 784] / y [N]")
    xbatch, ybatch = (pseudo_rand_init(x, y, num_pseudo=num_pseudo, seed=seed, nc=nc) if init_args == "random"
                      else pseudo_subsample_init(x, y, num_pseudo=num_pseudo, seed=seed, nc=nc))
    n_train = len(train_dataset)
    tr = _make_trainer(net, seed, noise_source)
    xs = xbatch.detach().to(device, torch.float32).reshape(xbatch.shape[0], -1).contiguous()
    ys = ybatch.detach().to(device).to(torch.int32).contiguous()
    xtd = torch.as_tensor(test_dataset.data).to(device, torch.float32).reshape(len(test_dataset), -1).contiguous()
    ytd = torch.as_tensor(test_dataset.targets).to(device).to(torch.int32).contiguous()
    sum_scaling = n_train / num_pseudo
    total = mul_fact * num_epochs
    i = 0
    while i < total:
        # iterations up to and including the next evaluation point (evaluations happen after steps 0, k, 2k, ...)
        nxt = i if i % log_every == 0 else min(((i // log_every) + 1) * log_every, total - 1)
        T = nxt - i + 1
        losses = tr.train(xs, ys, sum_scaling, T, lr0net)
        elbos += [-v for v in losses.cpu().tolist()]
        i += T
        if (i - 1) % log_every == 0:
            acc, nll = tr.test(xtd, ytd, int(data_minibatch))
            times.append(times[-1] + time.time() - t_start)
            nlls.append(nll)
            accs.append(acc)
            if not kwargs.get("quiet", False):
                print(f"predictive accuracy: {(100*accs[-1]):.2f}%")
    results = {"accs": accs, "nlls": nlls, "times": times[1:], "elbos": elbos, "csizes": [num_pseudo] * total}
    if log_pseudodata:
        results["us"], results["zs"], results["vs"] = xbatch.detach(), ybatch.detach(), [sum_scaling] * num_pseudo
        results["grid_preds"] = []
    return results


def run_mfvi(xt=None, yt=None, mc_samples=4, data_minibatch=128, num_epochs=100, log_every=10, N=None, D=None,
             lr0net=1e-3, mul_fact=2, seed=0, distr_fn=categorical_fn, architecture=None, n_hidden=None, nc=2,
             log_pseudodata=False, train_dataset=None, test_dataset=None, init_sd=None, noise_source=None,
             **kwargs) -> Dict[str, Any]:
    """Mean-field VI on the full training set with random minibatches (reference baselines.py:824-920)."""
    device = torch.device("cuda" if torch.cuda.is_available() else "cpu")
    random.seed(seed), np.random.seed(seed), torch.manual_seed(seed)
    nlls, accs, times, elbos = [], [], [0], []
    t_start = time.time()
    net = set_up_model(architecture=architecture, D=D, n_hidden=n_hidden, nc=nc, mc_samples=mc_samples,
                       init_sd=init_sd).to(device)
    tr = _make_trainer(net, seed, noise_source)
    xd = torch.as_tensor(train_dataset.data).to(device, torch.float32).reshape(len(train_dataset), -1).contiguous()
    yd = torch.as_tensor(train_dataset.targets).to(device).to(torch.int32).contiguous()
    xtd = torch.as_tensor(test_dataset.data).to(device, torch.float32).reshape(len(test_dataset), -1).contiguous()
    ytd = torch.as_tensor(test_dataset.targets).to(device).to(torch.int32).contiguous()
    n_train = xd.shape[0]
    total = mul_fact * num_epochs
    for i in range(total):
        idx = torch.randperm(n_train)[: min(int(data_minibatch), n_train)].to(device)
        xb, yb = xd[idx].contiguous(), yd[idx].contiguous()
        losses = tr.train(xb, yb, n_train / xb.shape[0], 1, lr0net)
        elbos.append(-losses.item())
        if i % log_every == 0 or i == total - 1:
            acc, nll = tr.test(xtd, ytd, int(data_minibatch))
            times.append(times[-1] + time.time() - t_start)
            nlls.append(nll)
            accs.append(acc)
            if not kwargs.get("quiet", False):
                print(f"predictive accuracy: {(100*accs[-1]):.2f}%")
    results = {"accs": accs, "nlls": nlls, "times": times[1:], "elbos": elbos, "csizes": None}
    if log_pseudodata:
        results["grid_preds"] = []
    return results


def fit(net=None, optim_vi=None, train_loader=None, pred_loader=None, revert_norm=None, log_every=-1, tau=1e-2, epochs=40,
        device=None, seed=0, noise_source=None, quiet=True):
    """Fit a mean-field regressor BNN with the standard ELBO and log its predictive performance (reference baselines.py:1283-
    1346).  As in the reference every step takes `next(iter(train_loader))` -- with an unshuffled loader that is the FIRST
    minibatch each time -- and scales its log-likelihood by len(train_loader.dataset) / batch; predictions are the plain mean
    over samples of the de-normalised outputs, scored with N(., 1 / tau)."""
    dev = torch.device("cuda") if device is None else device
    tr = _GaussTrainer(net, tau, seed, noise_source)
    lr = float(optim_vi.param_groups[0]["lr"]) if optim_vi is not None else 1e-3
    n_train = len(train_loader.dataset)
    pred = [(xt.to(dev, torch.float32).reshape(xt.shape[0], -1).contiguous(), yt.to(dev, torch.float32).reshape(-1))
            for xt, yt in pred_loader]
    checkpoint = (lambda it: (it % log_every) == 0) if log_every > 0 else (lambda it: it == epochs - 1)
    lls, rmses, times, elbos = [], [], [0], []
    t_start = time.time()
    for e in range(epochs):
        xb, yb = next(iter(train_loader))
        xb = xb.to(dev, torch.float32).reshape(xb.shape[0], -1).contiguous()
        yb = yb.to(dev, torch.float32).reshape(-1).contiguous()
        elbos.append(-tr.step(xb, yb, n_train / xb.shape[0], lr).item())
        if checkpoint(e):
            se = torch.zeros((), device=dev, dtype=torch.float64)
            ll = torch.zeros((), device=dev, dtype=torch.float64)
            total = 0
            for xt, yt in pred:
                d = (revert_norm(tr.outputs(xt)).mean(0) - yt).double()
                se += (d * d).sum()
                ll += (-0.5 * tau * d * d - 0.5 * np.log(2.0 * np.pi / tau)).sum()
                total += yt.shape[0]
            times.append(times[-1] + time.time() - t_start)
            lls.append((ll / total).item())
            rmses.append((se / total).sqrt().item())
            if not quiet:
                print(f"Predictive rmse {rmses[-1]:.2f} | pred ll {lls[-1]:.2f}")
    return {"rmses": rmses, "lls": lls, "times": times[1:], "elbos": elbos, "scale": 1.0 / np.sqrt(tau)}


def _regressor_run(train_loader, n_train, mc_samples, data_minibatch, num_epochs, log_every, D, lr0net, seed, architecture, n_hidden,
                   val_dataset, test_dataset, nc, y_mean, y_std, taus, init_sd, model_selection, noise_source, quiet):
    from torch.utils.data import DataLoader
    device = torch.device("cuda")
    test_loader = DataLoader(test_dataset, batch_size=data_minibatch, shuffle=False)
    val_loader = DataLoader(val_dataset, batch_size=data_minibatch, shuffle=False)
    bpe = max(1, int(n_train / data_minibatch))
    revert_norm = lambda y_pred: y_pred * float(y_std) + float(y_mean)

    def new_net():
        net = set_up_model(architecture=architecture, D=D, n_hidden=n_hidden, nc=nc, mc_samples=mc_samples, init_sd=init_sd).to(device)
        return net, torch.optim.Adam(net.parameters(), lr0net)
    best_tau, best_ll = taus[0], -float("inf")
    if model_selection:          # grid search of the precision on the validation set (reference :1116-1146)
        for tau in taus:
            net, opt = new_net()
            r = fit(net=net, optim_vi=opt, train_loader=train_loader, pred_loader=val_loader, revert_norm=revert_norm, log_every=-1,
                    tau=tau, epochs=num_epochs * bpe, device=device, seed=seed, noise_source=noise_source, quiet=quiet)
            if r["lls"][-1] > best_ll:
                best_tau, best_ll = tau, r["lls"][-1]
    net, opt = new_net()
    res = fit(net=net, optim_vi=opt, train_loader=train_loader, pred_loader=test_loader, revert_norm=revert_norm, log_every=log_every,
              tau=best_tau, epochs=num_epochs * bpe, device=device, seed=seed, noise_source=noise_source, quiet=quiet)
    res["selected_tau"] = best_tau
    return res


def run_mfvi_regressor(mc_samples=4, data_minibatch=128, num_epochs=100, log_every=10, D=None, lr0net=1e-3, seed=0,
                       architecture=None, n_hidden=None, train_dataset=None, val_dataset=None, test_dataset=None, nc=1, y_mean=None,
                       y_std=None, taus=None, init_sd=1e-6, model_selection=True, dnm=None, noise_source=None,
                       **kwargs) -> Dict[str, Any]:
    """Mean-field VI for BNN regression on the full training set (reference baselines.py:1066-1170)."""
    from torch.utils.data import DataLoader
    random.seed(seed), np.random.seed(seed), torch.manual_seed(seed)
    train_loader = DataLoader(train_dataset, batch_size=data_minibatch, shuffle=False)
    return _regressor_run(train_loader, len(train_dataset), mc_samples, data_minibatch, num_epochs, log_every, D, lr0net, seed,
                          architecture, n_hidden, val_dataset, test_dataset, nc, y_mean, y_std, taus, init_sd, model_selection,
                          noise_source, kwargs.get("quiet", True))


def run_mfvi_subset_regressor(mc_samples=4, data_minibatch=128, num_epochs=100, log_every=10, D=None, lr0net=1e-3, seed=0,
                              architecture=None, n_hidden=None, train_dataset=None, val_dataset=None, test_dataset=None, nc=1,
                              y_mean=None, y_std=None, init_sd=1e-6, num_pseudo=100, taus=None, model_selection=False,
                              noise_source=None, **kwargs) -> Dict[str, Any]:
    """The same on a random subset of `num_pseudo` training rows (reference baselines.py:1173-1278; the subset's log-likelihood
    is NOT rescaled to the full data size: `fit` scales by len(loader.dataset) / batch = 1)."""
    from torch.utils.data import DataLoader, Subset
    random.seed(seed), np.random.seed(seed), torch.manual_seed(seed)
    idx = random.sample(range(len(train_dataset)), num_pseudo)
    loader = DataLoader(Subset(train_dataset, idx), batch_size=num_pseudo, shuffle=False)
    res = _regressor_run(loader, len(train_dataset), mc_samples, data_minibatch, num_epochs, log_every, D, lr0net, seed,
                         architecture, n_hidden, val_dataset, test_dataset, nc, y_mean, y_std, taus, init_sd, model_selection,
                         noise_source, kwargs.get("quiet", True))
    res["csizes"] = [num_pseudo]
    return res
