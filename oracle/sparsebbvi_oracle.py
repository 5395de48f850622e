"""CPU restatement of the reference's sparse-BBVI coreset construction (psvi/inference/sparsebbvi.py:28-198 with the helpers
psvi/inference/utils.py:85-141) -- TEST INFRASTRUCTURE ONLY (never imported by the product).  numpy, any float dtype.

What the reference does per outer iteration (quirks kept, they change the numbers):
  1. `inner_it` Adam steps of the net on  elbo = S * sum_s sum_m w_m nll[s, m] - sum_s sampled_nkl_s  over the current coreset
     (utils.py:85-91: `(pseudo_nll.sum() - sampled_nkl).sum()` subtracts an [S] vector from a scalar, so the data term counts S
     times); `optim_net0.zero_grad()` is called ONCE before the loop, so the gradients of the inner iterations
     ACCUMULATE (sparsebbvi.py:133-140);
  2. one forward over cat(coreset, minibatch): centred log-likelihoods, residual, correlations (:144-163);
  3. selection: `pt_idx = sub_idcs[torch.argmax(torch.max(corrs))]` (:170) -- the argmax of a SCALAR is 0, so the point added is
     the FIRST index of the minibatch (if the data correlations beat the coreset ones), not the best-correlated one;
  4. `outer_it` Adam steps of the coreset weights w on the PSVI objective with pseudo term N / |core| * nll_core @ w
     (utils.py:94-105), clamped at 0 after every step; the net's gradients of that loss are never used.
Likelihood: Bernoulli on ONE logit, nll = softplus(o) - y o."""
import numpy as np

from oracle import psvi_oracle as po


def bern_nll(out, y):
    o = out[..., 0]
    return np.logaddexp(0.0, o) - y[None, :] * o, 1.0 / (1.0 + np.exp(-o))


def elbo_grad(mu, rho, eps, u, z, w, dims):
    """value and d/dmu, d/drho of utils.elbo."""
    theta = po.mf_sample(mu, rho, eps)
    S = eps.shape[0]
    if u.shape[0]:
        # `(pseudo_nll.sum() - sampled_nkl).sum()`: a scalar minus an [S] vector, summed -- the data term counts S times
        out, cache = po.mlp_forward(theta, u, dims)
        nll, sg = bern_nll(out, z)
        tb, _ = po.mlp_backward(theta, cache, dims, (S * w[None, :] * (sg - z[None, :]))[..., None])
        data = S * np.sum(nll @ w)
    else:
        tb, data = np.zeros_like(theta), 0.0
    beta = -np.ones(S, dtype=mu.dtype)
    tb = tb - beta[:, None] * theta                       # d(-nkl_s)/dtheta_s = +theta_s
    mu_bar, rho_bar = po.reparam_grad(mu, rho, eps, tb, kl_coef=0.0, rho_extra=beta.sum() / po.softplus(rho))
    return data - po.mf_sampled_nkl(mu, rho, eps, theta).sum(), mu_bar, rho_bar


def forward_through_coreset(mu, rho, eps, u, x, z, y, w, dims):
    theta = po.mf_sample(mu, rho, eps)
    out, _ = po.mlp_forward(theta, np.concatenate([u, x], 0), dims)
    ll = -bern_nll(out, np.concatenate([z, y]))[0]
    M = u.shape[0]
    lw = (ll[:, :M] @ w if M else 0.0) + po.mf_sampled_nkl(mu, rho, eps, theta)
    return ll[:, :M].T, ll[:, M:].T, po.softmax(lw, 0)


def psvi_w_grad(mu, rho, eps, x, u, y, z, w, N, dims):
    """value and d/dw of utils.sparsevi_psvi_elbo."""
    theta = po.mf_sample(mu, rho, eps)
    M, B, S = u.shape[0], x.shape[0], eps.shape[0]
    out, _ = po.mlp_forward(theta, np.concatenate([u, x], 0), dims)
    nll = bern_nll(out, np.concatenate([z, y]))[0]
    ps, ds = (N / M) * (nll[:, :M] @ w), nll[:, M:].sum(-1)
    lw = -ps + po.mf_sampled_nkl(mu, rho, eps, theta)
    wt = po.softmax(lw, 0)
    e = (N / B) * ds - ps
    loss = np.sum(wt * e) - lw.mean()
    beta = wt * (e - np.sum(wt * e)) - 1.0 / S
    gp = -wt - beta
    return loss, (N / M) * (gp @ nll[:, :M])


def predict_through_coreset(mu, rho, eps, xt, x, y, w, dims):
    theta = po.mf_sample(mu, rho, eps)
    out, _ = po.mlp_forward(theta, np.concatenate([xt, x], 0), dims)
    nt = xt.shape[0]
    pn = bern_nll(out[:, nt:], y)[0]
    wt = po.softmax(-(pn @ w) + po.mf_sampled_nkl(mu, rho, eps, theta), 0)
    return out[:, :nt, 0], wt


def run(mu, rho, eps_iter, x, y, xt, yt, dims, num_epochs, inner_it, outer_it, data_minibatch, log_every, lr0, seed, dtype=np.float32):
    """Returns dict(accs, nlls, csizes, core_idcs, w, mu, rho); eps_iter yields one [S, P] noise slab per forward."""
    dt = dtype
    rng = np.random.RandomState(seed)              # the reference seeds numpy's GLOBAL generator: np.random.seed(seed)
    N = x.shape[0]
    mu, rho = mu.astype(dt), rho.astype(dt)
    P = len(mu)
    mN, vN, tN = np.zeros(2 * P, dt), np.zeros(2 * P, dt), 0
    w = np.zeros(N, dt); mW, vW, tW = np.zeros(N, dt), np.zeros(N, dt), 0
    core, accs, nlls, csizes = [], [], [], []
    for it in range(num_epochs):
        if it % log_every == 0:
            lg, wt = predict_through_coreset(mu, rho, next(eps_iter), xt, x, y, w, dims)
            probs = np.minimum(wt @ (1.0 / (1.0 + np.exp(-lg))), 1.0)
            accs.append(np.mean((probs > 0.5) == (yt > 0.5)))
            pc = np.clip(probs, np.finfo(dt).eps, 1 - np.finfo(dt).eps)
            nlls.append(-np.mean(yt * np.log(pc) + (1 - yt) * np.log1p(-pc)))
            csizes.append(len(core))
        sub = rng.randint(N, size=data_minibatch)
        scale = N / data_minibatch
        g_acc = np.zeros(2 * P, dt)
        for _ in range(inner_it):
            _, gmu, grho = elbo_grad(mu, rho, next(eps_iter), x[core], y[core], w[core], dims)
            g_acc = g_acc + np.concatenate([gmu, grho]).astype(dt)          # no zero_grad inside the loop
            tN += 1
            phi, mN, vN = po.torch_adam_step(np.concatenate([mu, rho]), g_acc, mN, vN, tN, dt(lr0))
            mu, rho = phi[:P].astype(dt), phi[P:].astype(dt)
        ll_core, ll_data, wt = forward_through_coreset(mu, rho, next(eps_iter), x[core], x[sub], y[core], y[sub], w[core], dims)
        cd, cc = ll_data - wt[None, :] * ll_data, ll_core - wt[None, :] * ll_core
        resid = scale * cd.sum(0) - (w[core] @ cc if len(core) else 0.0)
        corrs = cd @ resid / np.sqrt((cd ** 2).sum(1)) / cd.shape[1]
        cmax = (np.abs(cc @ resid) / np.sqrt((cc ** 2).sum(1)) / cc.shape[1]).max() if len(core) else None
        if cmax is None or corrs.max() > cmax:
            if sub[0] not in core:
                core.append(sub[0])
        sub = rng.randint(N, size=data_minibatch)
        for _ in range(outer_it):
            _, gw = psvi_w_grad(mu, rho, next(eps_iter), x[sub], x[core], y[sub], y[core], w[core], N, dims)
            g = np.zeros(N, dt); g[core] = gw       # (w[core_idcs] with repeated indices cannot occur: core has no duplicates)
            tW += 1
            w, mW, vW = po.torch_adam_step(w, g, mW, vW, tW, dt(lr0))
            w = np.maximum(w, 0).astype(dt)
    return dict(accs=accs, nlls=nlls, csizes=csizes, core_idcs=list(core), w=w, mu=mu, rho=rho)
