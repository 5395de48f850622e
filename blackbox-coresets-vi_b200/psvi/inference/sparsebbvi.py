"""Incremental variational coreset with the black-box PSVI objective ("sparse-BBVI"): reference psvi/inference/sparsebbvi.py:28-198
with the helpers of psvi/inference/utils.py:85-141 (elbo, sparsevi_psvi_elbo, forward_through_coreset, predict_through_coreset).

B200 path: a mean-field net with ONE logit under a Bernoulli likelihood; every network pass is psvi_net_pass_bernoulli (csrc/
psvi_mf_stream.cu) on sampled weights, the reparameterisation / sampled-nkl maps are the fused family kernels of the streaming
engine.  The coreset weights w only need the per-sample NLLs of a FORWARD pass (their gradient is closed form: the importance
weights depend on w through the pseudo term alone), so the outer loop costs one forward launch per step.

Upstream behaviour that changes the numbers and is kept (see oracle/sparsebbvi_oracle.py, pinned by tests/golden/
sparsebbvi_hm.npz): the inner gradients ACCUMULATE (zero_grad once per outer iteration, :133-140); `elbo` counts its data term S
times (a scalar minus an [S] vector, utils.py:91); the point added to the coreset is the FIRST index of the minibatch
(`argmax(max(corrs))`, :170)."""
from __future__ import annotations

import time

import numpy as np
import torch
import torch.nn as nn

from psvi import _native
from psvi.models.neural_net import MeanFieldMLP, VILinear, make_fcnet


class _BernoulliNet:
    """Per-sample pass with the Bernoulli likelihood (labels 0. / 1.)."""

    def __init__(self, dims, S):
        self.desc = _native.make_model(dims, S)

    def pass_(self, theta, thetad, x, y, cw, logits=None, **out):
        _native.net_pass_bernoulli(self.desc, theta, thetad, x, y, cw, outputs=logits, **out)


class _Engine:
    def __init__(self, net, S, seed, noise_source):
        from psvi.inference.stream import MeanFieldFamily
        self.fam, self.net, self.S = MeanFieldFamily(net), _BernoulliNet(net.dims, S), S
        self.Pt, self.dev = self.fam.Pt, next(net.parameters()).device
        self.seed, self.noise_source, self.domain = seed, noise_source, 0

    def eps(self):
        if self.noise_source is not None:
            return self.noise_source.take(1, self.dev)[0]
        self.domain += 1
        e = torch.empty(1, self.S, self.Pt, device=self.dev)
        _native.philox_normal(self.seed, self.domain, 0, 1, self.S, self.Pt, e)
        return e[0]

    def forward(self, phi, x, y, want_logits=False):
        """One sampled forward over rows x: (nll [S, R], nkl [S], logits [S, R] or None)."""
        eps = self.eps()
        theta = self.fam.sample(phi, eps)
        nkl = self.fam.nkl(phi, eps, theta)
        R = x.shape[0]
        nll = torch.empty(self.S, R, device=self.dev)
        lg = torch.empty(self.S, R, 1, device=self.dev) if want_logits else None
        if R:
            self.net.pass_(theta, None, x, y, None, nll=nll, logits=lg)
        return nll, nkl, (lg[..., 0] if want_logits else None)

    def elbo_grad(self, phi, u, z, w):
        """value and d/dphi of utils.elbo = S * sum_s sum_m w_m nll[s, m] - sum_s sampled_nkl_s."""
        eps = self.eps()
        theta = self.fam.sample(phi, eps)
        S, M = self.S, u.shape[0]
        tbar = torch.zeros(S, self.Pt, device=self.dev)
        data = torch.zeros((), device=self.dev, dtype=torch.float64)
        if M:
            nll = torch.empty(S, M, device=self.dev)
            cw = (float(S) * w)[None, :].expand(S, M).contiguous()
            self.net.pass_(theta, None, u, z, cw, nll=nll, tbar=tbar)
            data = float(S) * (nll.double() @ w.double()).sum()
        beta = torch.full((S,), -1.0, device=self.dev, dtype=torch.float64)
        g = self.fam.grad_with_nkl(phi, eps, tbar, beta, theta, -float(S))
        return (data - self.fam.nkl(phi, eps, theta).sum()).float(), g


def run_sparsevi_with_bb_elbo(n_layers=1, logistic_regression=True, n_hidden=40, log_every=10, lr0=1e-3, register_elbos=False, seed=0,
                              noise_source=None, **kwargs):
    """Same keyword surface and results dict as the reference (:28-198)."""
    _native.require_cuda()
    np.random.seed(seed), torch.manual_seed(seed)
    dev = torch.device("cuda")
    elbos, results = [], {}
    num_epochs, inner_it, outer_it = kwargs["num_epochs"], kwargs["inner_it"], kwargs["outer_it"]
    mc_samples, data_minibatch = kwargs["mc_samples"], kwargs["data_minibatch"]
    f32 = lambda t: torch.as_tensor(t).to(dev, torch.float32).contiguous()
    x, y, xt, yt = f32(kwargs["x"]), f32(kwargs["y"]).reshape(-1), f32(kwargs["xt"]), f32(kwargs["yt"]).reshape(-1)
    N, D = x.shape
    net = (MeanFieldMLP(VILinear(D, 1, mc_samples=mc_samples)) if logistic_regression
           else make_fcnet(D, n_hidden, 1, n_layers=n_layers, linear_class=VILinear, nonl_class=nn.ReLU, mc_samples=mc_samples)).to(dev)
    net.flat()
    eng = _Engine(net, mc_samples, seed, noise_source)
    phi = eng.fam.get_phi().clone().contiguous()
    if kwargs.get("_init") is not None:            # (parity tests inject the reference's initial weights)
        phi = f32(kwargs["_init"])
    mN, vN, tN = torch.zeros_like(phi), torch.zeros_like(phi), 0
    w = torch.zeros(N, device=dev)
    mW, vW, tW = torch.zeros_like(w), torch.zeros_like(w), 0

    def adam(p, g, m, v, t):
        m = 0.9 * m + (1.0 - 0.9) * g
        v = 0.999 * v + (1.0 - 0.999) * g * g
        return p - (lr0 / (1.0 - 0.9 ** t)) * m / (v.sqrt() / (1.0 - 0.999 ** t) ** 0.5 + 1e-8), m, v
    nlls_s, accs_s, csizes, core, times = [], [], [], [], [0]
    t_start = time.time()
    fe = float(np.finfo(np.float32).eps)
    for it in range(num_epochs):
        if it % log_every == 0:                    # predict_through_coreset (utils.py:125-141): ALL data rows weighted by w
            nll_a, nkl, lg = eng.forward(phi, torch.cat([xt, x]), torch.cat([yt, y]), want_logits=True)
            nt = xt.shape[0]
            wt = torch.softmax(-(nll_a[:, nt:].double() @ w.double()) + nkl, 0)
            probs = (wt @ torch.sigmoid(lg[:, :nt].double())).clamp(max=1.0).float()
            accs_s.append(probs.gt(0.5).float().eq(yt).float().mean().item())
            pc = probs.clamp(fe, 1.0 - fe)
            nlls_s.append((-(yt * pc.log() + (1.0 - yt) * torch.log1p(-pc))).mean().item())
            csizes.append(len(core))
            times.append(times[-1] + time.time() - t_start)
        ci = torch.as_tensor(core, device=dev, dtype=torch.long)
        sub = torch.as_tensor(np.random.randint(N, size=data_minibatch), device=dev, dtype=torch.long)
        scale = N / data_minibatch
        # 1. coreset posterior: inner_it Adam steps on the coreset ELBO, gradients accumulating across the steps
        g_acc = torch.zeros_like(phi)
        for in_it in range(inner_it):
            val, g = eng.elbo_grad(phi, x[ci].contiguous(), y[ci].contiguous(), w[ci].contiguous())
            if register_elbos and in_it % log_every == 0:
                elbos.append((1, -val.item()))
            g_acc = g_acc + g
            tN += 1
            phi, mN, vN = adam(phi, g_acc, mN, vN, tN)
        # 2. centred log-likelihoods of the coreset and of a minibatch under the coreset posterior
        nll_a, nkl, _ = eng.forward(phi, torch.cat([x[ci], x[sub]]), torch.cat([y[ci], y[sub]]))
        M = len(core)
        ll = -nll_a.double()
        lw = (ll[:, :M] @ w[ci].double() if M else 0.0) + nkl
        wt = torch.softmax(lw, 0)
        ll_core, ll_data = ll[:, :M].T, ll[:, M:].T
        cd, cc = ll_data - wt[None, :] * ll_data, ll_core - wt[None, :] * ll_core
        resid = scale * cd.sum(0) - (w[ci].double() @ cc if M else 0.0)
        corrs = cd @ resid / (cd ** 2).sum(1).sqrt() / cd.shape[1]
        cmax = ((cc @ resid).abs() / (cc ** 2).sum(1).sqrt() / cc.shape[1]).max() if M else None
        # 3. selection: argmax of a scalar is 0 -- the first index of the minibatch (reference :170)
        if cmax is None or bool(corrs.max() > cmax):
            pt = int(sub[0])
            if pt not in core:
                core.append(pt)
        ci = torch.as_tensor(core, device=dev, dtype=torch.long)
        sub = torch.as_tensor(np.random.randint(N, size=data_minibatch), device=dev, dtype=torch.long)
        # 4. coreset weights: outer_it projected Adam steps on the PSVI objective (utils.py:94-105); closed-form d/dw
        for out_it in range(outer_it):
            nll_a, nkl, _ = eng.forward(phi, torch.cat([x[ci], x[sub]]), torch.cat([y[ci], y[sub]]))
            M, B, S = len(core), data_minibatch, mc_samples
            nd = nll_a.double()
            ps, ds = (N / M) * (nd[:, :M] @ w[ci].double()), nd[:, M:].sum(1)
            lw = -ps + nkl
            wt = torch.softmax(lw, 0)
            e = (N / B) * ds - ps
            ebar = (wt * e).sum()
            if register_elbos and out_it % log_every == 0:
                elbos.append((0, -(ebar - lw.mean()).item()))
            gp = -wt - (wt * (e - ebar) - 1.0 / S)
            g = torch.zeros_like(w)
            g[ci] = ((N / M) * (gp @ nd[:, :M])).float()
            tW += 1
            w, mW, vW = adam(w, g, mW, vW, tW)
            w = w.clamp(min=0.0)
    eng.fam.set_phi(phi)
    results["accs"], results["nlls"], results["csizes"] = accs_s, nlls_s, csizes
    results["times"], results["elbos"] = times[1:], elbos
    results["core_idcs"], results["w"] = list(core), w.detach().cpu()
    return results
