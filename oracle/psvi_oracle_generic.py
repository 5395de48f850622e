"""CPU oracle, family-generic part -- TEST INFRASTRUCTURE ONLY (see oracle/psvi_oracle.py for the rules).

The PSVI objectives only see the variational family through four maps (phi = flat variational parameters in torch's
`parameters_to_vector` order, eps = one [S, P_theta] noise slab in theta layout):
    sample(phi, eps) -> theta [S, P_theta]            kl(phi)            nkl(phi, eps, theta) -> [S]
and their first / second order adjoints.  Two families:
  * MeanField  -- psvi/models/neural_net.py:60-179   (phi per layer: weight, bias, _weight_sd, _bias_sd)
  * FullCov    -- psvi/models/neural_net.py:408-491  (phi per layer: mean[n], _sd[n], _corr[(n-1)(n-2)/2]; theta_s =
                  mean + L eps_s with L = scale_tril :452-461: diag softplus(_sd), strictly-lower entries of the top-left
                  (n-1)x(n-1) block in tril_indices(n-1, n-1, -1) order, last row diagonal only (Q6);
                  kl :435-436 = 0.5(|L|_F^2 + |m|^2 - n) - sum log L_ii; sampled_nkl :438-442 =
                  -|theta|^2/2 + |eps|^2/2 + sum log L_ii, which the reference obtains through a triangular solve).
The generic objective / gradient / HVP / nested-step routines below reduce to oracle/psvi_oracle.py for MeanField
(asserted in tests) and are pinned against the reference's fn2 outputs in tests/golden/fn2_*.npz.
"""
from __future__ import annotations

import numpy as np

from oracle import psvi_oracle as po


class MeanField:
    def __init__(self, dims):
        self.dims = list(dims)
        self.Pt = po.p_theta(dims)
        self.Pphi = 2 * self.Pt

    def split(self, phi):
        return po.phi_to_mu_rho(phi, self.dims)

    def join(self, gmu, grho):
        return po.mu_rho_to_phi(gmu, grho, self.dims)

    def sample(self, phi, eps):
        mu, rho = self.split(phi)
        return po.mf_sample(mu, rho, eps)

    def tangent(self, phi, phidot, eps):
        mu, rho = self.split(phi)
        md, rd = self.split(phidot)
        return md[None] + (po.sigmoid(rho) * rd)[None] * eps

    def kl(self, phi):
        return po.mf_kl(*self.split(phi))

    def nkl(self, phi, eps, theta):
        mu, rho = self.split(phi)
        return po.mf_sampled_nkl(mu, rho, eps, theta)

    def grad(self, phi, eps, tbar, kl_coef, nkl_coef):
        """phi_bar from theta_bar [S,P]; + kl_coef * dKL/dphi + nkl_coef * d(sum_i log sigma_i)/dphi."""
        mu, rho = self.split(phi)
        sg, sig = po.softplus(rho), po.sigmoid(rho)
        gmu = tbar.sum(0) + kl_coef * mu
        grho = sig * ((tbar * eps).sum(0) + kl_coef * (sg - 1 / sg) + nkl_coef / sg)
        return self.join(gmu, grho)

    def hvp(self, phi, phidot, eps, A_t, A_td):
        mu, rho = self.split(phi)
        md, rd = self.split(phidot)
        sg, sig = po.softplus(rho), po.sigmoid(rho)
        hmu = A_t.sum(0) + md
        hrho = (sig * (A_t * eps).sum(0) + sig * (1 - sig) * rd * (A_td * eps).sum(0)
                + ((1 + 1 / (sg * sg)) * sig * sig + (sg - 1 / sg) * sig * (1 - sig)) * rd)
        return self.join(hmu, hrho)


class FullCov:
    def __init__(self, dims):
        self.dims = list(dims)
        self.ns = [dims[l] * (dims[l - 1] + 1) for l in range(1, len(dims))]
        self.ncs = [max((n - 1) * (n - 2) // 2, 0) for n in self.ns]
        self.Pt = sum(self.ns)
        self.Pphi = sum(2 * n + c for n, c in zip(self.ns, self.ncs))

    def layers(self, phi):
        out, off = [], 0
        for n, c in zip(self.ns, self.ncs):
            out.append((phi[off:off + n], phi[off + n:off + 2 * n], phi[off + 2 * n:off + 2 * n + c], n))
            off += 2 * n + c
        return out

    @staticmethod
    def tril_idx(n):
        r, c = np.tril_indices(n - 1, -1)     # same order as torch.tril_indices(n-1, n-1, -1)
        return r, c

    def dense_L(self, sd, corr, n, diag_map=po.softplus):
        L = np.zeros((n, n), dtype=sd.dtype)
        L[np.arange(n), np.arange(n)] = diag_map(sd)
        r, c = self.tril_idx(n)
        L[r, c] = corr
        return L

    def sample(self, phi, eps):
        out, off = [], 0
        for (m, sd, corr, n) in self.layers(phi):
            L = self.dense_L(sd, corr, n)
            out.append(m[None] + eps[:, off:off + n] @ L.T)
            off += n
        return np.concatenate(out, 1)

    def tangent(self, phi, phidot, eps):
        out, off = [], 0
        for (m, sd, corr, n), (md, sdd, corrd, _) in zip(self.layers(phi), self.layers(phidot)):
            Ld = self.dense_L(po.sigmoid(sd) * sdd, corrd, n, diag_map=lambda x: x)
            out.append(md[None] + eps[:, off:off + n] @ Ld.T)
            off += n
        return np.concatenate(out, 1)

    def kl(self, phi):
        t = 0.0
        for (m, sd, corr, n) in self.layers(phi):
            d = po.softplus(sd)
            t += 0.5 * (np.sum(d * d) + np.sum(corr * corr) + np.sum(m * m) - n) - np.sum(np.log(d))
        return t

    def nkl(self, phi, eps, theta):
        logdet = sum(np.sum(np.log(po.softplus(sd))) for (_, sd, _, _) in self.layers(phi))
        return -0.5 * np.sum(theta * theta, 1) + 0.5 * np.sum(eps * eps, 1) + logdet

    def grad(self, phi, eps, tbar, kl_coef, nkl_coef):
        out, off = [], 0
        for (m, sd, corr, n) in self.layers(phi):
            A, e = tbar[:, off:off + n], eps[:, off:off + n]
            d, sig = po.softplus(sd), po.sigmoid(sd)
            G = A.T @ e                                   # dLoss/dL (dense); only the structural entries are parameters
            r, c = self.tril_idx(n)
            gm = A.sum(0) + kl_coef * m
            gsd = sig * (np.diag(G) + kl_coef * (d - 1 / d) + nkl_coef / d)
            gcorr = G[r, c] + kl_coef * corr
            out += [gm, gsd, gcorr]
            off += n
        return np.concatenate(out)

    def hvp(self, phi, phidot, eps, A_t, A_td):
        out, off = [], 0
        for (m, sd, corr, n), (md, sdd, corrd, _) in zip(self.layers(phi), self.layers(phidot)):
            A, Ad, e = A_t[:, off:off + n], A_td[:, off:off + n], eps[:, off:off + n]
            d, sig = po.softplus(sd), po.sigmoid(sd)
            G, Gd = A.T @ e, Ad.T @ e
            r, c = self.tril_idx(n)
            hm = A.sum(0) + md
            hsd = (sig * np.diag(G) + sig * (1 - sig) * sdd * np.diag(Gd)
                   + ((1 + 1 / (d * d)) * sig * sig + (d - 1 / d) * sig * (1 - sig)) * sdd)
            hcorr = G[r, c] + corrd
            out += [hm, hsd, hcorr]
            off += n
        return np.concatenate(out)


# ---------------------------------------------------------------------------------------------- generic objectives
class _MLP:
    """The network seen through four maps; families may carry their own (`fam.net`, e.g. oracle.lenet_oracle.LeNet)."""

    def __init__(self, dims):
        self.dims = dims

    def forward(self, theta, X):
        return po.mlp_forward(theta, X, self.dims)

    def backward(self, theta, cache, obar):
        return po.mlp_backward(theta, cache, self.dims, obar)

    def dual_forward(self, theta, thetad, X):
        return po.mlp_dual_forward(theta, thetad, X, self.dims)

    def dual_backward(self, theta, thetad, cache, A_o, A_od):
        return po.mlp_dual_backward(theta, thetad, cache, self.dims, A_o, A_od)


def _net(fam):
    return getattr(fam, "net", None) or _MLP(fam.dims)


def _nkl_theta_grad(fam, theta):
    f = getattr(fam, "nkl_theta_grad", None)
    return f(theta) if f is not None else -theta


def _q(p, lab):
    q = p.copy()
    np.add.at(q, (slice(None), np.arange(len(lab)), lab.astype(np.int64)), -1.0)
    return q


def inner_grad(fam, phi, eps, u, z, a):
    theta = fam.sample(phi, eps)
    logits, cache = _net(fam).forward(theta, u)
    nll, p = po.nll_rows(logits, z)
    tb, xbar = _net(fam).backward(theta, cache, a[None, :, None] * _q(p, z))
    return np.sum(nll @ a) + fam.kl(phi), fam.grad(phi, eps, tb, 1.0, 0.0), xbar.sum(0), nll.sum(0)


def outer_grad(fam, phi, eps, u, z, a, xb, yb, N):
    theta = fam.sample(phi, eps)
    X, lab = np.concatenate([u, xb], 0), np.concatenate([z, yb], 0)
    logits, cache = _net(fam).forward(theta, X)
    nll, p = po.nll_rows(logits, lab)
    S, M, B = eps.shape[0], u.shape[0], xb.shape[0]
    ps, ds = nll[:, :M] @ a, (N / B) * nll[:, M:].sum(-1)
    lw = -ps + fam.nkl(phi, eps, theta)
    w = po.softmax(lw, 0)
    e = ds - ps
    loss = np.sum(w * e) - lw.mean()
    beta = w * (e - np.sum(w * e)) - 1.0 / S
    gp = -w - beta
    rw = np.concatenate([gp[:, None] * a[None, :], np.broadcast_to((w * N / B)[:, None], (S, B))], 1)
    tb, xbar = _net(fam).backward(theta, cache, rw[:, :, None] * _q(p, lab))
    tb = tb + beta[:, None] * _nkl_theta_grad(fam, theta)
    return loss, fam.grad(phi, eps, tb, 0.0, beta.sum()), xbar[:, :M].sum(0), gp @ nll[:, :M]


def outer_grad_ablated(fam, phi, eps, xb, yb, N):
    """PSVI_Ablated.psvi_elbo (psvi_classes.py:1397-1408): mean_s (N/B) sum_b nll[s, b] - mean_s sampled_nkl_s -- no importance
    weights, no pseudo-data term; value and d/dphi (the direct partials wrt u and v are zero)."""
    theta = fam.sample(phi, eps)
    logits, cache = _net(fam).forward(theta, xb)
    nll, p = po.nll_rows(logits, yb)
    S, B = eps.shape[0], xb.shape[0]
    loss = np.mean((N / B) * nll.sum(-1)) - np.mean(fam.nkl(phi, eps, theta))
    tb, _ = _net(fam).backward(theta, cache, (N / B / S) * _q(p, yb))
    tb = tb - (1.0 / S) * _nkl_theta_grad(fam, theta)
    return loss, fam.grad(phi, eps, tb, 0.0, -1.0)


def inner_hvp(fam, phi, eps, u, z, a, phidot):
    theta, thetad = fam.sample(phi, eps), fam.tangent(phi, phidot, eps)
    o, od, cache = _net(fam).dual_forward(theta, thetad, u)
    _, p = po.nll_rows(o, z)
    q = _q(p, z)
    c = a[None, :, None]
    A_t, A_td, A_x = _net(fam).dual_backward(theta, thetad, cache, c * p * (od - (p * od).sum(-1, keepdims=True)), c * q)
    return fam.hvp(phi, phidot, eps, A_t, A_td), A_x.sum(0), (q * od).sum(-1).sum(0)


def no_iw_expand(u, z, a, C):
    """Quirk of inner_elbo with mc_samples == 1 (psvi_classes.py:493-494,497,505): logits [M, C] are unsqueezed at dim 1, so
    Categorical(logits [M, 1, C]).log_prob(z [M]) broadcasts to [M, M] -- pseudo-point i is scored against EVERY label z_j --
    and the matmul with N f(v) weights column j by a_j:  inner = sum_i sum_j a_j nll(u_i, z_j) + kl
    = sum_i sum_c A_c nll(u_i, c) with A_c = sum_{j: z_j = c} a_j.  Restated as an ordinary weighted objective over M C rows."""
    M = len(z)
    A = np.bincount(z.astype(np.int64), weights=a, minlength=C)
    z2 = np.tile(np.arange(C), M)
    return np.repeat(u, C, 0), z2.astype(z.dtype), A[z2]


def no_iw_collapse(u_bar2, a_bar2, z, C):
    M = len(z)
    return u_bar2.reshape(M, C, -1).sum(1), a_bar2.reshape(M, C).sum(0)[z.astype(np.int64)]


def nested_step(fam, phi0, eps_inner, eps_outer, u, z, v, xb, yb, N, lr, vmode=1, outer="psvi", no_iw_classes=None):
    """outer = "psvi" (PSVI.psvi_elbo) or "ablated" (PSVI_Ablated.psvi_elbo, psvi_classes.py:1388-1408);
    no_iw_classes = C reproduces the mc_samples == 1 quirk of inner_elbo (see no_iw_expand)."""
    T = len(eps_inner)
    a = po.coreset_weights(v, N, vmode)
    u_in, z_in, a_in = (u, z, a) if no_iw_classes is None else no_iw_expand(u, z, a, no_iw_classes)
    phi = phi0.copy()
    m, vv = np.zeros_like(phi), np.zeros_like(phi)
    traj, inner_losses = [], []
    for t in range(T):
        val, g, _, _ = inner_grad(fam, phi, eps_inner[t], u_in, z_in, a_in)
        phi_new, m, vv = po.robust_adam_step(phi, g, m, vv, t + 1, lr)
        traj.append((phi, g, m, vv))
        inner_losses.append(val)
        phi = phi_new
    if outer == "ablated":
        loss, pbar = outer_grad_ablated(fam, phi, eps_outer, xb, yb, N)
        u_bar, a_bar = np.zeros_like(u), np.zeros_like(a)
    else:
        loss, pbar, u_bar, a_bar = outer_grad(fam, phi, eps_outer, u, z, a, xb, yb, N)
    mbar, vbar = np.zeros_like(pbar), np.zeros_like(pbar)
    for t in range(T - 1, -1, -1):
        phi_t, g, m_t, v_t = traj[t]
        gbar, mbar, vbar = po.robust_adam_step_vjp(pbar, mbar, vbar, g, m_t, v_t, t + 1, lr)
        h, hu, ha = inner_hvp(fam, phi_t, eps_inner[t], u_in, z_in, a_in, gbar)
        if no_iw_classes is not None:
            hu, ha = no_iw_collapse(hu, ha, z, no_iw_classes)
        pbar, u_bar, a_bar = pbar + h, u_bar + hu, a_bar + ha
    v_bar, _ = po.coreset_weights_vjp(v, N, vmode, a_bar)
    return dict(loss=loss, u_grad=u_bar, v_grad=v_bar, phi_T=phi, inner_losses=np.array(inner_losses))


def evaluate(fam, phi, eps_batches, u, z, a, xt, yt, batch):
    """PSVI.evaluate (psvi_classes.py:1031-1108) for any family; returns (acc, nll, iw_entropy, ness)."""
    M, tot, nll_sum, corr = u.shape[0], 0, 0.0, 0.0
    fe = np.finfo(phi.dtype).eps
    for k, s0 in enumerate(range(0, xt.shape[0], batch)):
        xb, yb = xt[s0:s0 + batch], yt[s0:s0 + batch]
        theta = fam.sample(phi, eps_batches[k])
        logits, _ = _net(fam).forward(theta, np.concatenate([u, xb], 0))
        lw = (po.nll_rows(logits[:, :M], z)[0] @ a) + fam.nkl(phi, eps_batches[k], theta)   # sign quirk Q3
        w = po.softmax(lw, 0)
        probs = (po.softmax(logits[:, M:], -1) * w[:, None, None]).sum(0)
        corr += np.sum(probs.argmax(-1) == yb.astype(np.int64))
        tot += len(yb)
        pn = np.clip(probs / probs.sum(-1, keepdims=True), fe, 1 - fe)
        nll_sum += -np.sum(np.log(pn[np.arange(len(yb)), yb.astype(np.int64)]))
    wp = w[w > 0]
    return corr / tot, nll_sum / tot, -np.sum(np.log(wp) * wp), w.sum() ** 2 / np.sum(w * w) / len(w)
