"""CPU oracle for the PSVI hot path -- TEST INFRASTRUCTURE ONLY.

This file is a numpy restatement of the reference algorithm (souravc83/Blackbox-Coresets-VI,
mounted at /root/reference while developing).  It is the *checker* for the CUDA path: only
`tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference` legs may
import it.  The product (`blackbox-coresets-vi_b200/`) never imports or calls anything in here.

Parity status: the reference ships NO tests / golden vectors for this path, so the
oracle is pinned against outputs of the reference itself, run in the build container with a recorded
noise stream (`oracle/make_goldens.py` -> `tests/golden/*.npz`; `tests/test_oracle_vs_golden.py`).

Everything is written for an arbitrary float dtype (fp64 to decide "who is wrong", fp32 to mimic the
reference's arithmetic).  Reference citations use paths relative to the reference root.

Layouts (shared with include/psvi_b200.h):
  * dims = [d0, d1, ..., dL]  -- an MLP with L VILinear layers (L=1: logistic_regression, L=2: fn with one
    hidden layer, ...), ReLU between layers (psvi/models/neural_net.py:267-297, psvi_classes.py:694-717).
  * theta layout ("TL"): per layer l: W_l [d_l, d_{l-1}] row-major, then b_l [d_l].  P_theta = sum_l d_l*(d_{l-1}+1).
  * mu[P_theta], rho[P_theta] in TL.  sigma = softplus(rho)  (neural_net.py:129-131).
  * eps[S, P_theta] in TL: the standard-normal draws of one forward (weight draw first, then bias draw, per
    layer in module order: neural_net.py:155-162).
  * phi (torch `parameters_to_vector` order, neural_net.py:63-69 / Q10): per layer weight, bias, _weight_sd, _bias_sd.
"""
from __future__ import annotations

import numpy as np

LOG2PI_HALF = 0.5 * np.log(2.0 * np.pi)


# --------------------------------------------------------------------------------------------- helpers
def softplus(x):
    # F.softplus(beta=1, threshold=20)  (neural_net.py:131)
    return np.where(x > 20.0, x, np.log1p(np.exp(np.minimum(x, 20.0))))


def sigmoid(x):
    return 1.0 / (1.0 + np.exp(-x))


def inverse_softplus(x):
    # neural_net.py:32-35
    return np.log(np.expm1(x))


def p_theta(dims):
    return int(sum(dims[l] * (dims[l - 1] + 1) for l in range(1, len(dims))))


def layer_slices(dims):
    """[(w_slice, b_slice, out, in)] into a TL vector."""
    out, off = [], 0
    for l in range(1, len(dims)):
        o, i = dims[l], dims[l - 1]
        out.append((slice(off, off + o * i), slice(off + o * i, off + o * i + o), o, i))
        off += o * i + o
    return out


def phi_to_mu_rho(phi, dims):
    """torch parameters_to_vector order -> (mu, rho) in TL."""
    mu, rho, off = [], [], 0
    for l in range(1, len(dims)):
        o, i = dims[l], dims[l - 1]
        w = phi[off:off + o * i]; off += o * i
        b = phi[off:off + o]; off += o
        ws = phi[off:off + o * i]; off += o * i
        bs = phi[off:off + o]; off += o
        mu += [w, b]; rho += [ws, bs]
    return np.concatenate(mu), np.concatenate(rho)


def mu_rho_to_phi(mu, rho, dims):
    out = []
    for (ws, bs, o, i) in layer_slices(dims):
        out += [mu[ws], mu[bs], rho[ws], rho[bs]]
    return np.concatenate(out)


def log_softmax(o):
    m = o.max(-1, keepdims=True)
    return o - m - np.log(np.exp(o - m).sum(-1, keepdims=True))


def softmax(o, axis=-1):
    m = o.max(axis, keepdims=True)
    e = np.exp(o - m)
    return e / e.sum(axis, keepdims=True)


# --------------------------------------------------------------------------- mean-field layer (a1, A.1)
def mf_sample(mu, rho, eps):
    """theta[S,P] = mu + softplus(rho) * eps   (neural_net.py:155-162 via Normal.rsample)."""
    return mu[None, :] + softplus(rho)[None, :] * eps


def mf_kl(mu, rho):
    """sum_layers VIMixin.kl() with prior N(0,1)  (neural_net.py:101-108)."""
    sg = softplus(rho)
    return np.sum(0.5 * (sg * sg + mu * mu - 1.0) - np.log(sg))


def mf_sampled_nkl(mu, rho, eps, theta=None):
    """sum_layers VIMixin.sampled_nkl() -> [S]  (neural_net.py:110-115).
    log p(theta) - log q(theta) = -theta^2/2 + eps^2/2 + log sigma   (the 0.5*log(2pi) cancels)."""
    sg = softplus(rho)
    if theta is None:
        theta = mf_sample(mu, rho, eps)
    return np.sum(-0.5 * theta * theta + 0.5 * eps * eps + np.log(sg)[None, :], axis=1)


# ------------------------------------------------------------------------------------ MLP forward (a2)
def mlp_forward(theta, X, dims):
    """Per-sample MLP: VILinear.forward `x @ W_s^T + b_s` (neural_net.py:176-179) + nn.ReLU between layers.
    theta [S,P], X [R,d0] -> logits [S,R,C], cache of layer inputs / masks."""
    S = theta.shape[0]
    x = np.broadcast_to(X[None], (S,) + X.shape)
    xs, ks = [x], []
    sl = layer_slices(dims)
    for li, (ws, bs, o, i) in enumerate(sl):
        W = theta[:, ws].reshape(S, o, i)
        b = theta[:, bs]
        a = np.einsum("sri,soi->sro", x, W) + b[:, None, :]
        if li < len(sl) - 1:
            k = (a > 0).astype(theta.dtype)
            x = a * k
            ks.append(k); xs.append(x)
        else:
            logits = a
    return logits, (xs, ks)


def nll_rows(logits, labels):
    """-Categorical(logits).log_prob(labels)  (neural_net.py:22-23; psvi_classes.py:466,497) -> [S,R], plus softmax p."""
    ls = log_softmax(logits)
    lab = labels.astype(np.int64)
    nll = -np.take_along_axis(ls, np.broadcast_to(lab[None, :, None], ls.shape[:2] + (1,)), axis=2)[..., 0]
    return nll, np.exp(ls)


def mlp_backward(theta, cache, dims, obar):
    """Reverse pass of mlp_forward for output adjoint obar [S,R,C] -> (theta_bar [S,P], X_bar [S,R,d0])."""
    xs, ks = cache
    S = theta.shape[0]
    tb = np.zeros_like(theta)
    sl = layer_slices(dims)
    abar = obar
    for li in range(len(sl) - 1, -1, -1):
        ws, bs, o, i = sl[li]
        W = theta[:, ws].reshape(S, o, i)
        tb[:, ws] = np.einsum("sro,sri->soi", abar, xs[li]).reshape(S, -1)
        tb[:, bs] = abar.sum(1)
        xbar = np.einsum("sro,soi->sri", abar, W)
        if li > 0:
            abar = xbar * ks[li - 1]
    return tb, xbar


def mlp_dual_forward(theta, theta_dot, X, dims):
    """Primal + tangent forward (tangent of X is zero).  SURVEY Appendix A.6 generalised to L layers."""
    S = theta.shape[0]
    x = np.broadcast_to(X[None], (S,) + X.shape)
    xd = np.zeros_like(x)
    xs, xds, ks = [x], [xd], []
    sl = layer_slices(dims)
    for li, (ws, bs, o, i) in enumerate(sl):
        W = theta[:, ws].reshape(S, o, i); b = theta[:, bs]
        Wd = theta_dot[:, ws].reshape(S, o, i); bd = theta_dot[:, bs]
        a = np.einsum("sri,soi->sro", x, W) + b[:, None, :]
        ad = np.einsum("sri,soi->sro", xd, W) + np.einsum("sri,soi->sro", x, Wd) + bd[:, None, :]
        if li < len(sl) - 1:
            k = (a > 0).astype(theta.dtype)
            x, xd = a * k, ad * k
            ks.append(k); xs.append(x); xds.append(xd)
        else:
            o_, od_ = a, ad
    return o_, od_, (xs, xds, ks)


def mlp_dual_backward(theta, theta_dot, cache, dims, A_o, A_od):
    """Adjoint of the scalar Ldot wrt (theta, theta_dot, X) given adjoints of (o, odot)."""
    xs, xds, ks = cache
    S = theta.shape[0]
    A_t = np.zeros_like(theta); A_td = np.zeros_like(theta)
    sl = layer_slices(dims)
    A_a, A_ad = A_o, A_od
    for li in range(len(sl) - 1, -1, -1):
        ws, bs, o, i = sl[li]
        W = theta[:, ws].reshape(S, o, i); Wd = theta_dot[:, ws].reshape(S, o, i)
        A_t[:, ws] = (np.einsum("sro,sri->soi", A_a, xs[li]) + np.einsum("sro,sri->soi", A_ad, xds[li])).reshape(S, -1)
        A_t[:, bs] = A_a.sum(1)
        A_td[:, ws] = np.einsum("sro,sri->soi", A_ad, xs[li]).reshape(S, -1)
        A_td[:, bs] = A_ad.sum(1)
        A_x = np.einsum("sro,soi->sri", A_a, W) + np.einsum("sro,soi->sri", A_ad, Wd)
        A_xd = np.einsum("sro,soi->sri", A_ad, W)
        if li > 0:
            A_a, A_ad = A_x * ks[li - 1], A_xd * ks[li - 1]
    return A_t, A_td, A_x


# ------------------------------------------------------------------------------ coreset weights a = N f(v)
def coreset_weights(v, N, vmode, alpha=0.0):
    """a = N * f(v).  vmode 0: f = identity (PSVI, psvi_classes.py:111); 1: f = softmax (PSVILearnV :1358-1360);
    2: f = exp(alpha) * softmax (PSVIAV :1486-1488)."""
    if vmode == 0:
        return N * v
    f = softmax(v, 0)
    return N * (np.exp(alpha) if vmode == 2 else 1.0) * f


def coreset_weights_vjp(v, N, vmode, abar, alpha=0.0):
    """(vbar, alphabar) given abar = dLoss/da."""
    if vmode == 0:
        return N * abar, 0.0
    f = softmax(v, 0)
    sc = N * (np.exp(alpha) if vmode == 2 else 1.0)
    vbar = sc * f * (abar - np.dot(f, abar))
    return vbar, (sc * np.dot(f, abar) if vmode == 2 else 0.0)


# -------------------------------------------------------------------------------- inner ELBO (a6, A.2, Q1)
def inner_elbo(mu, rho, eps, u, z, a, dims):
    """PSVI.inner_elbo (psvi_classes.py:488-511): sum_s sum_m a_m nll[s,m] + sum_layers kl()."""
    theta = mf_sample(mu, rho, eps)
    logits, _ = mlp_forward(theta, u, dims)
    nll, _ = nll_rows(logits, z)
    return np.sum(nll @ a) + mf_kl(mu, rho)


def reparam_grad(mu, rho, eps, theta_bar, kl_coef=1.0, rho_extra=None):
    """theta_bar[S,P] -> (mu_bar, rho_bar) incl. analytic KL gradient (SURVEY A.6 'reparam')."""
    sg, sig = softplus(rho), sigmoid(rho)
    mu_bar = theta_bar.sum(0) + kl_coef * mu
    r = (theta_bar * eps).sum(0) + kl_coef * (sg - 1.0 / sg)
    if rho_extra is not None:
        r = r + rho_extra
    return mu_bar, sig * r


def inner_grad(mu, rho, eps, u, z, a, dims):
    """value, d/dmu, d/drho, d/du, d/da of inner_elbo."""
    theta = mf_sample(mu, rho, eps)
    logits, cache = mlp_forward(theta, u, dims)
    nll, p = nll_rows(logits, z)
    q = p.copy()
    np.add.at(q, (slice(None), np.arange(len(z)), z.astype(np.int64)), -1.0)
    obar = a[None, :, None] * q
    tb, xbar = mlp_backward(theta, cache, dims, obar)
    mu_bar, rho_bar = reparam_grad(mu, rho, eps, tb)
    val = np.sum(nll @ a) + mf_kl(mu, rho)
    return val, mu_bar, rho_bar, xbar.sum(0), nll.sum(0)


def inner_hvp(mu, rho, eps, u, z, a, dims, mu_dot, rho_dot):
    """(H_phiphi g, H_uphi g, H_aphi g) for g = (mu_dot, rho_dot)  -- SURVEY A.6 (reverse over forward)."""
    sg, sig = softplus(rho), sigmoid(rho)
    theta = mu[None] + sg[None] * eps
    theta_dot = mu_dot[None] + (sig * rho_dot)[None] * eps
    o, od, cache = mlp_dual_forward(theta, theta_dot, u, dims)
    _, p = nll_rows(o, z)
    q = p.copy()
    np.add.at(q, (slice(None), np.arange(len(z)), z.astype(np.int64)), -1.0)
    c = a[None, :, None]
    A_od = c * q
    A_o = c * p * (od - (p * od).sum(-1, keepdims=True))
    A_c = (q * od).sum(-1).sum(0)  # [R]
    A_t, A_td, A_x = mlp_dual_backward(theta, theta_dot, cache, dims, A_o, A_od)
    hmu = A_t.sum(0) + mu_dot
    hrho = (sig * (A_t * eps).sum(0)
            + sig * (1 - sig) * rho_dot * (A_td * eps).sum(0)
            + ((1 + 1 / (sg * sg)) * sig * sig + (sg - 1 / sg) * sig * (1 - sig)) * rho_dot)
    return hmu, hrho, A_x.sum(0), A_c


# ------------------------------------------------------------------------------ outer PSVI ELBO (a7, A.2)
def psvi_elbo_parts(mu, rho, eps, u, z, a, xb, yb, N, dims):
    theta = mf_sample(mu, rho, eps)
    X = np.concatenate([u, xb], 0)
    lab = np.concatenate([z, yb], 0)
    logits, cache = mlp_forward(theta, X, dims)
    nll, p = nll_rows(logits, lab)
    M, B = u.shape[0], xb.shape[0]
    ps = nll[:, :M] @ a
    ds = (N / B) * nll[:, M:].sum(-1)
    nkl = mf_sampled_nkl(mu, rho, eps, theta)
    lw = -ps + nkl
    w = softmax(lw, 0)
    loss = np.sum(w * (ds - ps)) - lw.mean()
    return loss, dict(theta=theta, cache=cache, nll=nll, p=p, ps=ps, ds=ds, nkl=nkl, lw=lw, w=w, lab=lab, X=X)


def psvi_elbo(mu, rho, eps, u, z, a, xb, yb, N, dims):
    """PSVI.psvi_elbo (psvi_classes.py:445-486), non-learn_z branch."""
    return psvi_elbo_parts(mu, rho, eps, u, z, a, xb, yb, N, dims)[0]


def psvi_elbo_grad(mu, rho, eps, u, z, a, xb, yb, N, dims, kappa=1.0, n_total_rows=None):
    """value, d/dmu, d/drho, d/du, d/da of psvi_elbo (SURVEY A.2 closed forms).
    kappa / n_total_rows describe one rank's share when the minibatch rows are sharded over R ranks:
    L_r = sum_s w_s (d_s^r - kappa p_s) - kappa mean(lw) with kappa = 1/R and d_s^r = (N / n_total_rows) sum over the
    rank's rows; the shares (values and gradients) sum to the unsharded objective because w is rank-independent."""
    loss, t = psvi_elbo_parts(mu, rho, eps, u, z, a, xb, yb, N, dims)
    S = eps.shape[0]
    M, B = u.shape[0], xb.shape[0]
    if n_total_rows is not None:
        t["ds"] = t["ds"] * (B / n_total_rows)
    Btot = B if n_total_rows is None else n_total_rows
    w, e = t["w"], t["ds"] - kappa * t["ps"]
    loss = np.sum(w * e) - kappa * t["lw"].mean()
    beta = w * (e - np.sum(w * e)) - kappa / S        # dLoss/dlw_s
    gp = -kappa * w - beta                              # dLoss/dp_s
    q = t["p"].copy()
    np.add.at(q, (slice(None), np.arange(M + B), t["lab"].astype(np.int64)), -1.0)
    rw = np.concatenate([gp[:, None] * a[None, :], np.broadcast_to((w * N / Btot)[:, None], (S, B))], 1)  # [S,R]
    obar = rw[:, :, None] * q
    tb, xbar = mlp_backward(t["theta"], t["cache"], dims, obar)
    tb = tb - beta[:, None] * t["theta"]               # d nkl_s / d theta = -theta
    sg = softplus(rho)
    mu_bar, rho_bar = reparam_grad(mu, rho, eps, tb, kl_coef=0.0, rho_extra=beta.sum() / sg)
    a_bar = gp @ t["nll"][:, :M]
    return loss, mu_bar, rho_bar, xbar[:, :M].sum(0), a_bar, t


# ----------------------------------------------------------- differentiable Adam + reverse sweep (a8,a9, A.4)
def robust_adam_step(p, g, m, v, t, lr, b1=0.9, b2=0.999, eps=1e-8):
    """DifferentiableAdam._update (psvi/robust_higher/optim.py:303-367); t is 1-based."""
    m = b1 * m + (1 - b1) * g
    v = b2 * v + (1 - b2) * g * g
    q = np.sqrt(v + 1e-8)
    den = q / np.sqrt(1 - b2 ** t) + eps
    p_new = p - (lr / (1 - b1 ** t)) * m / den
    return p_new, m, v


def robust_adam_step_vjp(pbar_next, mbar_next, vbar_next, g, m, v, t, lr, b1=0.9, b2=0.999, eps=1e-8):
    """VJP of one step.  m, v are the post-update moments of step t.
    Returns (gbar, mbar_carry, vbar_carry) where the carries are adjoints of m_{t-1}, v_{t-1}."""
    k = lr / (1 - b1 ** t)
    q = np.sqrt(v + 1e-8)
    sq2 = np.sqrt(1 - b2 ** t)
    den = q / sq2 + eps
    mbar = mbar_next - k * pbar_next / den
    denbar = k * pbar_next * m / (den * den)
    vbar = vbar_next + denbar / (2.0 * q * sq2)
    vbar = np.where(v == 0.0, 0.0, vbar)            # _maybe_mask hook (optim.py:40-52,346-347)
    gbar = (1 - b1) * mbar + 2 * (1 - b2) * g * vbar
    return gbar, b1 * mbar, b2 * vbar


def nested_step(mu, rho, eps_inner, eps_outer, u, z, v, xb, yb, N, dims, lr, vmode=1, alpha=0.0):
    """PSVI.nested_step (psvi_classes.py:541-600) minus the optimiser steps on u, v:
    T = len(eps_inner) unrolled robust-Adam steps on inner_elbo, then psvi_elbo and its hypergradient.
    Returns dict(loss, u_grad, v_grad, alpha_grad, mu_T, rho_T, inner_losses, traj)."""
    T = eps_inner.shape[0]
    a = coreset_weights(v, N, vmode, alpha)
    P = mu.shape[0]
    phi = np.concatenate([mu, rho])
    m = np.zeros_like(phi); vv = np.zeros_like(phi)
    traj, inner_losses = [], []
    for t in range(T):
        val, gmu, grho, _, _ = inner_grad(phi[:P], phi[P:], eps_inner[t], u, z, a, dims)
        g = np.concatenate([gmu, grho])
        phi_new, m, vv = robust_adam_step(phi, g, m, vv, t + 1, lr)
        traj.append((phi, g, m, vv))
        inner_losses.append(val)
        phi = phi_new
    loss, mu_bar, rho_bar, u_bar, a_bar, parts = psvi_elbo_grad(phi[:P], phi[P:], eps_outer, u, z, a, xb, yb, N, dims)
    pbar = np.concatenate([mu_bar, rho_bar])
    mbar = np.zeros_like(pbar); vbar = np.zeros_like(pbar)
    for t in range(T - 1, -1, -1):
        phi_t, g, m_t, v_t = traj[t]
        gbar, mbar, vbar = robust_adam_step_vjp(pbar, mbar, vbar, g, m_t, v_t, t + 1, lr)
        hmu, hrho, hu, ha = inner_hvp(phi_t[:P], phi_t[P:], eps_inner[t], u, z, a, dims, gbar[:P], gbar[P:])
        pbar = pbar + np.concatenate([hmu, hrho])
        u_bar = u_bar + hu
        a_bar = a_bar + ha
    v_bar, alpha_bar = coreset_weights_vjp(v, N, vmode, a_bar, alpha)
    return dict(loss=loss, u_grad=u_bar, v_grad=v_bar, alpha_grad=alpha_bar, mu_T=phi[:P], rho_T=phi[P:],
                inner_losses=np.array(inner_losses), phi0_grad=pbar, parts=parts)


# ---------------------------------------------------------------------------------------- torch.optim.Adam
def torch_adam_step(p, g, m, v, t, lr, b1=0.9, b2=0.999, eps=1e-8):
    """torch.optim.Adam single-tensor update (used for u, v: psvi_classes.py:860-868; mfvi: baselines.py:1013)."""
    m = b1 * m + (1 - b1) * g
    v = b2 * v + (1 - b2) * g * g
    den = np.sqrt(v) / np.sqrt(1 - b2 ** t) + eps
    return p - (lr / (1 - b1 ** t)) * m / den, m, v


# ------------------------------------------------------------------- --trainer joint / alternating
def _outer_grads(mu, rho, eps, u, z, v, xb, yb, N, dims, vmode, alpha):
    a = coreset_weights(v, N, vmode, alpha)
    loss, mb, rb, ub, ab, _ = psvi_elbo_grad(mu, rho, eps, u, z, a, xb, yb, N, dims)
    return loss, mb, rb, ub, coreset_weights_vjp(v, N, vmode, ab, alpha)[0]


def joint_steps(mu, rho, eps_steps, u, z, v, xb, yb, N, dims, lr, vmode=1, alpha=0.0, learn_v=True):
    """PSVI.joint_step (psvi_classes.py:517-526) called len(eps_steps) times: ONE torch Adam (lr0joint) over the model
    parameters, u and (if learnt) v on psvi_elbo; one noise draw per step.  Returns the losses and the final state."""
    st = {k: [x.copy(), 0 * x, 0 * x] for k, x in (("mu", mu), ("rho", rho), ("u", u), ("v", v))}
    losses = []
    for t, eps in enumerate(eps_steps, 1):
        loss, mb, rb, ub, vb = _outer_grads(st["mu"][0], st["rho"][0], eps, st["u"][0], z, st["v"][0], xb, yb, N, dims, vmode, alpha)
        losses.append(loss)
        for k, g in (("mu", mb), ("rho", rb), ("u", ub)) + ((("v", vb),) if learn_v else ()):
            st[k] = list(torch_adam_step(st[k][0], g, st[k][1], st[k][2], t, lr))
    return np.array(losses), st["mu"][0], st["rho"][0], st["u"][0], st["v"][0]


def alternating_steps(mu, rho, eps_steps, u, z, v, xb, yb, N, dims, lr_net, lr_u, vmode=1, alpha=0.0):
    """PSVI.alternating_step (psvi_classes.py:528-539): per call, an Adam step of the model (optim_net) on a fresh psvi_elbo
    draw, then an Adam step of u (optim_u) on another draw at the updated model; v is never stepped.  eps_steps holds two
    draws per call; the returned loss of a call is the second one."""
    net = {k: [x.copy(), 0 * x, 0 * x] for k, x in (("mu", mu), ("rho", rho))}
    us = [u.copy(), 0 * u, 0 * u]
    losses = []
    for t in range(1, len(eps_steps) // 2 + 1):
        _, mb, rb, _, _ = _outer_grads(net["mu"][0], net["rho"][0], eps_steps[2 * t - 2], us[0], z, v, xb, yb, N, dims, vmode, alpha)
        for k, g in (("mu", mb), ("rho", rb)):
            net[k] = list(torch_adam_step(net[k][0], g, net[k][1], net[k][2], t, lr_net))
        loss, _, _, ub, _ = _outer_grads(net["mu"][0], net["rho"][0], eps_steps[2 * t - 1], us[0], z, v, xb, yb, N, dims, vmode, alpha)
        us = list(torch_adam_step(us[0], ub, us[1], us[2], t, lr_u))
        losses.append(loss)
    return np.array(losses), net["mu"][0], net["rho"][0], us[0], v.copy()


# ------------------------------------------------------------------- learn_z: soft pseudo-labels (KLDiv branch)
def soft_targets(L):
    """`labels.softmax(0)` of the reference (psvi_classes.py:469-470,501-502): the label matrix [R, C] is normalised over the
    ROWS (dim 0), per class column -- the rows of a column sum to 1, the classes of a row do not."""
    return softmax(L, 0)


def soft_targets_vjp(L, G):
    """dLoss/dL given G = dLoss/dt for t = softmax(L, 0)."""
    t = softmax(L, 0)
    return t * (G - (t * G).sum(0, keepdims=True))


def _xlogx(t):
    return np.where(t > 0, t * np.log(np.where(t > 0, t, 1.0)), 0.0)


def soft_nll_rows(logits, t):
    """KLDivLoss(reduction="none")(log_softmax(logits), t).sum(classes) (psvi_classes.py:467-474): [S, R], plus softmax p and
    log-softmax."""
    ls = log_softmax(logits)
    return _xlogx(t).sum(-1)[None, :] - (t[None] * ls).sum(-1), np.exp(ls), ls


def inner_grad_soft(mu, rho, eps, u, t, a, dims):
    """value, d/dmu, d/drho of inner_elbo with soft targets t [M, C] (psvi_classes.py:488-511, learn_z branch)."""
    theta = mf_sample(mu, rho, eps)
    logits, cache = mlp_forward(theta, u, dims)
    nll, p, _ = soft_nll_rows(logits, t)
    obar = a[None, :, None] * (t.sum(-1)[None, :, None] * p - t[None])
    tb, _ = mlp_backward(theta, cache, dims, obar)
    mu_bar, rho_bar = reparam_grad(mu, rho, eps, tb)
    return np.sum(nll @ a) + mf_kl(mu, rho), mu_bar, rho_bar


def inner_hvp_soft(mu, rho, eps, u, t, a, dims, mu_dot, rho_dot):
    """(H_phiphi g, H_uphi g, H_aphi g, H_tphi g) of the soft-target inner objective for g = (mu_dot, rho_dot)."""
    sg, sig = softplus(rho), sigmoid(rho)
    theta = mu[None] + sg[None] * eps
    theta_dot = mu_dot[None] + (sig * rho_dot)[None] * eps
    o, od, cache = mlp_dual_forward(theta, theta_dot, u, dims)
    _, p, _ = soft_nll_rows(o, t)
    tau = t.sum(-1)[None, :, None]
    q = tau * p - t[None]
    c = a[None, :, None]
    lsd = od - (p * od).sum(-1, keepdims=True)          # d/d eps of log_softmax
    A_od = c * q
    A_o = c * tau * p * lsd
    A_c = (q * od).sum(-1).sum(0)                        # [M]  (= sum_s d/d eps of nll[s, m])
    A_tt = -(c * lsd).sum(0)                             # [M, C]  d/d eps of d inner / d t
    A_t, A_td, A_x = mlp_dual_backward(theta, theta_dot, cache, dims, A_o, A_od)
    hmu = A_t.sum(0) + mu_dot
    hrho = (sig * (A_t * eps).sum(0)
            + sig * (1 - sig) * rho_dot * (A_td * eps).sum(0)
            + ((1 + 1 / (sg * sg)) * sig * sig + (sg - 1 / sg) * sig * (1 - sig)) * rho_dot)
    return hmu, hrho, A_x.sum(0), A_c, A_tt


def psvi_elbo_grad_soft(mu, rho, eps, u, t_all, a, xb, N, dims):
    """value and d/dmu, d/drho, d/du, d/da, d/dt_all of psvi_elbo with soft targets t_all [M + B, C] (psvi_classes.py:445-486,
    learn_z branch: the data rows carry nc * one_hot(y) pushed through the same softmax over rows)."""
    theta = mf_sample(mu, rho, eps)
    M, B = u.shape[0], xb.shape[0]
    X = np.concatenate([u, xb], 0)
    logits, cache = mlp_forward(theta, X, dims)
    nll, p, ls = soft_nll_rows(logits, t_all)
    S = eps.shape[0]
    ps, ds = nll[:, :M] @ a, (N / B) * nll[:, M:].sum(-1)
    lw = -ps + mf_sampled_nkl(mu, rho, eps, theta)
    w = softmax(lw, 0)
    e = ds - ps
    loss = np.sum(w * e) - lw.mean()
    beta = w * (e - np.sum(w * e)) - 1.0 / S
    gp = -w - beta
    rw = np.concatenate([gp[:, None] * a[None, :], np.broadcast_to((w * N / B)[:, None], (S, B))], 1)
    obar = rw[:, :, None] * (t_all.sum(-1)[None, :, None] * p - t_all[None])
    tb, xbar = mlp_backward(theta, cache, dims, obar)
    tb = tb - beta[:, None] * theta
    sg = softplus(rho)
    mu_bar, rho_bar = reparam_grad(mu, rho, eps, tb, kl_coef=0.0, rho_extra=beta.sum() / sg)
    a_bar = gp @ nll[:, :M]
    dnll_dt = np.where(t_all > 0, np.log(np.where(t_all > 0, t_all, 1.0)) + 1.0, 0.0)[None] - ls      # [S, R, C]
    t_bar = (rw[:, :, None] * dnll_dt).sum(0)
    return loss, mu_bar, rho_bar, xbar[:, :M].sum(0), a_bar, t_bar


def nested_step_learn_z(mu, rho, eps_inner, eps_outer, u, z, v, xb, yb, N, dims, lr, nc, vmode=1, alpha=0.0):
    """PSVI.nested_step with learn_z=True (psvi_classes.py:541-600 with the KLDiv branches :455-474,499-504): the inner
    objective sees t_in = softmax(z, 0), the outer one t_all = softmax(cat(z, nc * one_hot(y)), 0); hypergradients on u, v and
    the soft labels z."""
    T, P, M = eps_inner.shape[0], mu.shape[0], u.shape[0]
    a = coreset_weights(v, N, vmode, alpha)
    t_in = soft_targets(z)
    L_all = np.concatenate([z, nc * np.eye(nc)[yb.astype(np.int64)]], 0)
    t_all = soft_targets(L_all)
    phi = np.concatenate([mu, rho])
    m = np.zeros_like(phi); vv = np.zeros_like(phi)
    traj = []
    for k in range(T):
        _, gmu, grho = inner_grad_soft(phi[:P], phi[P:], eps_inner[k], u, t_in, a, dims)
        g = np.concatenate([gmu, grho])
        phi_new, m, vv = robust_adam_step(phi, g, m, vv, k + 1, lr)
        traj.append((phi, g, m, vv))
        phi = phi_new
    loss, mu_bar, rho_bar, u_bar, a_bar, tall_bar = psvi_elbo_grad_soft(phi[:P], phi[P:], eps_outer, u, t_all, a, xb, N, dims)
    pbar = np.concatenate([mu_bar, rho_bar])
    mbar = np.zeros_like(pbar); vbar = np.zeros_like(pbar)
    tin_bar = np.zeros_like(t_in)
    for k in range(T - 1, -1, -1):
        phi_t, g, m_t, v_t = traj[k]
        gbar, mbar, vbar = robust_adam_step_vjp(pbar, mbar, vbar, g, m_t, v_t, k + 1, lr)
        hmu, hrho, hu, ha, ht = inner_hvp_soft(phi_t[:P], phi_t[P:], eps_inner[k], u, t_in, a, dims, gbar[:P], gbar[P:])
        pbar = pbar + np.concatenate([hmu, hrho])
        u_bar, a_bar, tin_bar = u_bar + hu, a_bar + ha, tin_bar + ht
    v_bar, _ = coreset_weights_vjp(v, N, vmode, a_bar, alpha)
    z_bar = soft_targets_vjp(z, tin_bar) + soft_targets_vjp(L_all, tall_bar)[:M]
    return dict(loss=loss, u_grad=u_bar, v_grad=v_bar, z_grad=z_bar, mu_T=phi[:P], rho_T=phi[P:])


def evaluate_learn_z(mu, rho, eps_batches, xt, yt, dims, batch):
    """PSVI.evaluate with learn_z=True (psvi_classes.py:1049-1056): the pseudo term is summed over classes AND samples before
    it meets the weights, so it shifts every log-weight equally and the importance weights reduce to softmax(sampled_nkl)."""
    nll_sum, correct, k, w = 0.0, 0, 0, None
    for r0 in range(0, xt.shape[0], batch):
        eps = eps_batches[k]; k += 1
        theta = mf_sample(mu, rho, eps)
        logits, _ = mlp_forward(theta, xt[r0:r0 + batch], dims)
        w = softmax(mf_sampled_nkl(mu, rho, eps, theta), 0)
        probs = (softmax(logits, -1) * w[:, None, None]).sum(0)
        lab = yt[r0:r0 + batch].astype(np.int64)
        pn = probs / probs.sum(-1, keepdims=True)
        pn = np.clip(pn, np.finfo(np.float32).eps, 1 - np.finfo(np.float32).eps)
        nll_sum += -np.log(pn[np.arange(len(lab)), lab]).sum()
        correct += (probs.argmax(-1) == lab).sum()
    n = xt.shape[0]
    S = w.shape[0]
    return correct / n, nll_sum / n, -(w * np.log(w)).sum(), (w.sum() ** 2 / (w * w).sum()) / S


# ------------------------------------------------------------------- regressors: Gaussian likelihood, learnable targets z
def gauss_nll_rows(out, y, tau):
    """-Normal(out, 1 / sqrt(tau)).log_prob(y)  (psvi_classes.py:1986, neural_net.py:18-19): [S, R] and the residuals."""
    r = out[..., 0] - y[None, :]
    return 0.5 * tau * r * r + 0.5 * np.log(2.0 * np.pi / tau), r


def inner_grad_gauss(mu, rho, eps, u, z, a, dims, tau):
    """value, d/dmu, d/drho of PSVI_regressor.inner_elbo (psvi_classes.py:2051-2057)."""
    theta = mf_sample(mu, rho, eps)
    out, cache = mlp_forward(theta, u, dims)
    nll, r = gauss_nll_rows(out, z, tau)
    tb, _ = mlp_backward(theta, cache, dims, (a[None, :] * tau * r)[..., None])
    mu_bar, rho_bar = reparam_grad(mu, rho, eps, tb)
    return np.sum(nll @ a) + mf_kl(mu, rho), mu_bar, rho_bar


def inner_hvp_gauss(mu, rho, eps, u, z, a, dims, tau, mu_dot, rho_dot):
    """(H_phiphi g, H_uphi g, H_aphi g, H_zphi g) of the Gaussian inner objective for g = (mu_dot, rho_dot)."""
    sg, sig = softplus(rho), sigmoid(rho)
    theta = mu[None] + sg[None] * eps
    theta_dot = mu_dot[None] + (sig * rho_dot)[None] * eps
    o, od, cache = mlp_dual_forward(theta, theta_dot, u, dims)
    _, r = gauss_nll_rows(o, z, tau)
    d = od[..., 0]
    A_od = (a[None, :] * tau * r)[..., None]
    A_o = (a[None, :] * tau * d)[..., None]
    A_c = (tau * r * d).sum(0)
    A_z = -(a[None, :] * tau * d).sum(0)
    A_t, A_td, A_x = mlp_dual_backward(theta, theta_dot, cache, dims, A_o, A_od)
    hmu = A_t.sum(0) + mu_dot
    hrho = (sig * (A_t * eps).sum(0)
            + sig * (1 - sig) * rho_dot * (A_td * eps).sum(0)
            + ((1 + 1 / (sg * sg)) * sig * sig + (sg - 1 / sg) * sig * (1 - sig)) * rho_dot)
    return hmu, hrho, A_x.sum(0), A_c, A_z


def psvi_elbo_grad_gauss(mu, rho, eps, u, z, a, xb, yb, N, dims, tau):
    """value and d/dmu, d/drho, d/du, d/da, d/dz of PSVI_regressor.psvi_elbo (psvi_classes.py:2034-2048)."""
    theta = mf_sample(mu, rho, eps)
    M, B, S = u.shape[0], xb.shape[0], eps.shape[0]
    out, cache = mlp_forward(theta, np.concatenate([u, xb], 0), dims)
    nll, r = gauss_nll_rows(out, np.concatenate([z, yb]), tau)
    ps, ds = nll[:, :M] @ a, (N / B) * nll[:, M:].sum(-1)
    lw = -ps + mf_sampled_nkl(mu, rho, eps, theta)
    w = softmax(lw, 0)
    e = ds - ps
    loss = np.sum(w * e) - lw.mean()
    beta = w * (e - np.sum(w * e)) - 1.0 / S
    gp = -w - beta
    rw = np.concatenate([gp[:, None] * a[None, :], np.broadcast_to((w * N / B)[:, None], (S, B))], 1)
    tb, xbar = mlp_backward(theta, cache, dims, (rw * tau * r)[..., None])
    tb = tb - beta[:, None] * theta
    mu_bar, rho_bar = reparam_grad(mu, rho, eps, tb, kl_coef=0.0, rho_extra=beta.sum() / softplus(rho))
    return loss, mu_bar, rho_bar, xbar[:, :M].sum(0), gp @ nll[:, :M], -(rw * tau * r)[:, :M].sum(0)


def nested_step_regressor(mu, rho, eps_inner, eps_outer, u, z, v, xb, yb, N, dims, lr, tau, vmode=1, alpha=0.0):
    """PSVI_regressor.nested_step (psvi_classes.py:2059-2093; PSVIAV_regressor :2303-2335): hypergradients on u, v[, alpha] and
    the learnable targets z through the unrolled inner Adam loop."""
    T, P = eps_inner.shape[0], mu.shape[0]
    a = coreset_weights(v, N, vmode, alpha)
    phi = np.concatenate([mu, rho])
    m = np.zeros_like(phi); vv = np.zeros_like(phi)
    traj = []
    for k in range(T):
        _, gmu, grho = inner_grad_gauss(phi[:P], phi[P:], eps_inner[k], u, z, a, dims, tau)
        g = np.concatenate([gmu, grho])
        phi_new, m, vv = robust_adam_step(phi, g, m, vv, k + 1, lr)
        traj.append((phi, g, m, vv))
        phi = phi_new
    loss, mu_bar, rho_bar, u_bar, a_bar, z_bar = psvi_elbo_grad_gauss(phi[:P], phi[P:], eps_outer, u, z, a, xb, yb, N, dims, tau)
    pbar = np.concatenate([mu_bar, rho_bar])
    mbar = np.zeros_like(pbar); vbar = np.zeros_like(pbar)
    for k in range(T - 1, -1, -1):
        phi_t, g, m_t, v_t = traj[k]
        gbar, mbar, vbar = robust_adam_step_vjp(pbar, mbar, vbar, g, m_t, v_t, k + 1, lr)
        hmu, hrho, hu, ha, hz = inner_hvp_gauss(phi_t[:P], phi_t[P:], eps_inner[k], u, z, a, dims, tau, gbar[:P], gbar[P:])
        pbar = pbar + np.concatenate([hmu, hrho])
        u_bar, a_bar, z_bar = u_bar + hu, a_bar + ha, z_bar + hz
    v_bar, alpha_bar = coreset_weights_vjp(v, N, vmode, a_bar, alpha)
    return dict(loss=loss, u_grad=u_bar, v_grad=v_bar, z_grad=z_bar, alpha_grad=alpha_bar, mu_T=phi[:P], rho_T=phi[P:])


def evaluate_regressor(mu, rho, eps_batches, xt, yt, dims, batch, tau, y_mean, y_std):
    """PSVI_regressor.evaluate (psvi_classes.py:2221-2264): the pseudo term is `.sum()`-med over samples before it meets the
    log-weights, so the weights are softmax(sampled_nkl); predictions are de-normalised, the log-likelihood uses them as they
    are.  Returns (rmse, mean log-likelihood)."""
    se, ll, k = 0.0, 0.0, 0
    for r0 in range(0, xt.shape[0], batch):
        eps = eps_batches[k]; k += 1
        theta = mf_sample(mu, rho, eps)
        out, _ = mlp_forward(theta, xt[r0:r0 + batch], dims)
        w = softmax(mf_sampled_nkl(mu, rho, eps, theta), 0)
        yp = ((out[..., 0] * y_std + y_mean) * w[:, None]).sum(0)
        y = yt[r0:r0 + batch]
        se += ((yp - y) ** 2).sum()
        ll += (-0.5 * tau * (yp - y) ** 2 - 0.5 * np.log(2.0 * np.pi / tau)).sum()
    return np.sqrt(se / xt.shape[0]), ll / xt.shape[0]


# ------------------------------------------------------------------------------------ evaluate (a11, Q3, Q12)
def evaluate(mu, rho, eps_batches, u, z, a, xt, yt, dims, batch, correction=True):
    """PSVI.evaluate (psvi_classes.py:1031-1108).  eps_batches[k] is the draw of test batch k.
    Returns (acc, nll, iw_entropy, ness) -- weights-based diagnostics from the LAST batch (Q12)."""
    M = u.shape[0]
    tot, nll_sum, corr = 0, 0.0, 0.0
    fe = np.finfo(mu.dtype).eps
    for k, s0 in enumerate(range(0, xt.shape[0], batch)):
        xb, yb = xt[s0:s0 + batch], yt[s0:s0 + batch]
        theta = mf_sample(mu, rho, eps_batches[k])
        logits, _ = mlp_forward(theta, np.concatenate([u, xb], 0), dims)
        ll_pseudo = -nll_rows(logits[:, :M], z)[0]             # +log p  (Q3: the sign quirk)
        pseudo = ll_pseudo @ a if M > 0 else 0.0
        lw = -pseudo + mf_sampled_nkl(mu, rho, eps_batches[k], theta)
        w = softmax(lw, 0)
        pr = softmax(logits[:, M:], -1)
        probs = (pr * w[:, None, None]).sum(0) if correction else pr.mean(0)
        corr += np.sum(probs.argmax(-1) == yb.astype(np.int64))
        tot += yb.shape[0]
        pn = probs / probs.sum(-1, keepdims=True)               # Categorical(probs=...) normalises + clamps
        pn = np.clip(pn, fe, 1 - fe)
        nll_sum += -np.sum(np.log(pn[np.arange(len(yb)), yb.astype(np.int64)]))
    wp = w[w > 0]
    iw_ent = -np.sum(np.log(wp) * wp)
    ness = w.sum() ** 2 / np.sum(w * w) / w.shape[0]
    return corr / tot, nll_sum / tot, iw_ent, ness


# -------------------------------------------------------------------------------- mfvi_subset step (a13)
def mfvi_grad(mu, rho, eps, x, y, scale, dims, kl_on=True):
    """loss = -scale * sum_{s,m} log p + sum kl   (baselines.py:1023-1028)."""
    a = np.full((x.shape[0],), scale, dtype=mu.dtype)
    val, gmu, grho, _, _ = inner_grad(mu, rho, eps, x, y, a, dims)
    if not kl_on:
        raise NotImplementedError
    return val, gmu, grho


def mfvi_predict(mu, rho, eps, xt, yt, dims):
    """test_logits = net(xt).mean(0); acc / nll (baselines.py:1039-1043). Returns (#correct, nll_sum)."""
    theta = mf_sample(mu, rho, eps)
    logits, _ = mlp_forward(theta, xt, dims)
    ml = logits.mean(0)
    ls = log_softmax(ml)
    yi = yt.astype(np.int64)
    return np.sum(ml.argmax(-1) == yi), -np.sum(ls[np.arange(len(yi)), yi])
