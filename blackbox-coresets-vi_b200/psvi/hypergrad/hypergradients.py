"""Approximate implicit differentiation on the B200 path.

Reference: psvi/hypergrad/hypergradients.py -- CG_normaleq :199-244, fixed_point :83-140, jvp :308-311,
CG_torch.cg (psvi/hypergrad/CG_torch.py:9-45).  With the fixed-point map Phi(w, lam) = w - eta * grad_w inner(w, lam; eps)
(GradientDescent, eta = linsys_lr) every Jacobian product the reference obtains from autograd is a Hessian-vector
product of the inner objective:
    J^T x = x - eta * H_eps x          (VJP through w_mapped, fixed noise eps_A)
    J   x = x - eta * H_eps' x         (jvp() = double VJP; it re-evaluates Phi twice, i.e. draws noise twice and uses the
                                        second draw)
    (dPhi/dlam)^T x = -eta * H_{lam,w} x
and each H x is one fused CUDA pass (psvi_mf_inner_hvp).  Vector algebra on the P-length vectors is plain torch.
"""
from __future__ import annotations

import torch

from psvi import _native


class _Hvp:
    def __init__(self, psvi, desc, mu, rho, u, z32, v):
        self.p, self.desc, self.mu, self.rho, self.u, self.z32, self.v = psvi, desc, mu, rho, u, z32, v
        self.P = mu.numel()
        self.M, self.D = u.shape

    def __call__(self, noise, vec):
        dev = self.mu.device
        hphi, hu, hv, ha = (torch.zeros(2 * self.P, device=dev), torch.zeros(self.M, self.D, device=dev),
                            torch.zeros(self.M, device=dev), torch.zeros(1, device=dev))
        p = self.p
        _native.inner_hvp(self.desc, noise, self.mu, self.rho, self.u, self.z32, self.v, float(p.N), p._vmode,
                          p._alpha_value(), vec.contiguous(), hphi, hu, hv, ha)
        return hphi, hu, hv, ha


def _outer(psvi, desc, mu, rho, u, z32, v, xbatch, ybatch):
    dev = mu.device
    M, D = u.shape
    P = mu.numel()
    gout = torch.zeros(_native.gout_floats(desc, M), device=dev)
    ug, vg, ag, loss = (torch.zeros(M, D, device=dev), torch.zeros(M, device=dev), torch.zeros(1, device=dev),
                        torch.zeros(1, device=dev))
    xb = xbatch.detach().to(dev, torch.float32).contiguous()
    yb = ybatch.detach().to(dev).to(torch.int32).contiguous()
    _native.outer_grad(desc, psvi._noise(1), mu, rho, u, z32, v, xb, yb, xb.shape[0], float(psvi.N), psvi._vmode,
                       psvi._alpha_value(), 1.0, gout, ug, vg, ag, loss)
    return gout[:2 * P].clone(), ug, vg, ag


def cg_normaleq_native(psvi, desc, mu, rho, u, z32, v, xbatch, ybatch, K, eta, tol=1e-10):
    """CG on the normal equations (I - J)(I - J^T) x = (I - J) g  (reference :199-244), K iterations."""
    hvp = _Hvp(psvi, desc, mu, rho, u, z32, v)
    g, ug0, vg0, ag0 = _outer(psvi, desc, mu, rho, u, z32, v, xbatch, ybatch)   # o_loss and its gradients
    noise_a = psvi._noise(1)                                                    # w_mapped = fp_map(params, hparams)

    def fresh():
        psvi._noise(1)           # first Phi evaluation inside jvp(): its draw is discarded (hypergradients.py:308-311)
        return psvi._noise(1)

    def A(x):
        t = eta * hvp(noise_a, x)[0]              # x - J^T x
        return eta * hvp(fresh(), t)[0]           # t - J t
    b = eta * hvp(fresh(), g)[0]                  # g - J g
    x_last, r_last, p_last = torch.zeros_like(b), b.clone(), b.clone()
    for _ in range(K):                            # CG_torch.cg, incl. its "break before x_last is updated" behaviour
        Ap = A(p_last)
        rTr = torch.sum(r_last * r_last)
        alpha = rTr / torch.sum(p_last * Ap)
        x = x_last + alpha * p_last
        r = r_last - alpha * Ap
        if float(torch.norm(r)) < tol:
            break
        beta = torch.sum(r * r) / rTr
        p_last = r + beta * p_last
        x_last, r_last = x, r
    _, hu, hv, ha = hvp(noise_a, x_last)          # (dPhi/dlam)^T x = -eta * H_{lam w} x
    return ug0 - eta * hu, vg0 - eta * hv, ag0 - eta * ha


def fixed_point_native(psvi, desc, mu, rho, u, z32, v, xbatch, ybatch, K, eta, tol=1e-10):
    """Stochastic fixed-point iteration v <- J^T v + g (reference :83-140 with stochastic=True)."""
    hvp = _Hvp(psvi, desc, mu, rho, u, z32, v)
    g, ug0, vg0, ag0 = _outer(psvi, desc, mu, rho, u, z32, v, xbatch, ybatch)
    vs = torch.zeros_like(g)
    for _ in range(K):
        prev = vs
        vs = vs - eta * hvp(psvi._noise(1), vs)[0] + g
        if float(torch.norm(vs - prev)) < tol:
            break
    _, hu, hv, ha = hvp(psvi._noise(1), vs)
    return ug0 - eta * hu, vg0 - eta * hv, ag0 - eta * ha


def CG_normaleq(*args, **kwargs):
    raise NotImplementedError("the autograd-callable form of CG_normaleq (reference hypergradients.py:199-244) is "
                              "replaced by cg_normaleq_native, which PSVI.hyper_step drives; generic fp_map/outer_loss "
                              "callables have no fused-kernel equivalent")


def fixed_point(*args, **kwargs):
    raise NotImplementedError("replaced by fixed_point_native (see CG_normaleq)")
