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


def _cg_normaleq(hvp, outer, new_noise, K, eta, tol=1e-10):
    """CG on the normal equations (I - J)(I - J^T) x = (I - J) g  (reference :199-244), K iterations.
    hvp(noise, vec) -> (H_phiphi vec, H_uphi vec, H_vphi vec, H_alphaphi vec); outer() -> (g, u_grad0, v_grad0, alpha_grad0);
    new_noise() draws one noise slab (token) in the reference's consumption order."""
    g, ug0, vg0, ag0 = outer()                    # o_loss and its gradients
    noise_a = new_noise()                         # w_mapped = fp_map(params, hparams)

    def fresh():
        new_noise()              # first Phi evaluation inside jvp(): its draw is discarded (hypergradients.py:308-311)
        return new_noise()

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


def _fixed_point(hvp, outer, new_noise, K, eta, tol=1e-10):
    """Stochastic fixed-point iteration v <- J^T v + g (reference :83-140 with stochastic=True)."""
    g, ug0, vg0, ag0 = outer()
    vs = torch.zeros_like(g)
    for _ in range(K):
        prev = vs
        vs = vs - eta * hvp(new_noise(), vs)[0] + g
        if float(torch.norm(vs - prev)) < tol:
            break
    _, hu, hv, ha = hvp(new_noise(), vs)
    return ug0 - eta * hu, vg0 - eta * hv, ag0 - eta * ha


def cg_normaleq_native(psvi, desc, mu, rho, u, z32, v, xbatch, ybatch, K, eta, tol=1e-10):
    hvp = _Hvp(psvi, desc, mu, rho, u, z32, v)
    return _cg_normaleq(hvp, lambda: _outer(psvi, desc, mu, rho, u, z32, v, xbatch, ybatch), lambda: psvi._noise(1), K, eta, tol)


def fixed_point_native(psvi, desc, mu, rho, u, z32, v, xbatch, ybatch, K, eta, tol=1e-10):
    hvp = _Hvp(psvi, desc, mu, rho, u, z32, v)
    return _fixed_point(hvp, lambda: _outer(psvi, desc, mu, rho, u, z32, v, xbatch, ybatch), lambda: psvi._noise(1), K, eta, tol)


def hyper_stream(psvi, eng, S, phi, u, z32, a, xb, yb, K, eta, approx="CG_normaleq", tol=1e-10):
    """The same solvers over the streaming engine (fn2, lenet, medium / large mean-field models): Hessian-vector products by
    StreamEngine.hvp, the outer gradient by StreamEngine.outer_grad; dL/da is mapped to (dL/dv, dL/dalpha) by the caller's f."""
    dev = phi.device

    def to_v(abar):
        vg, ag = psvi._v_grad_from_abar(abar)
        return vg, (ag if ag is not None else torch.zeros(1, device=dev))

    def hvp(eps, vec):
        h, hu, ha = eng.hvp(phi, eps, u, z32, a, vec.contiguous())
        return (h, hu) + to_v(ha)

    def outer():
        _, pbar, ubar, abar, _ = eng.outer_grad(phi, new_noise(), u, z32, a, xb, yb, float(psvi.N))
        return (pbar, ubar) + to_v(abar)

    def new_noise():
        return psvi._noise_tensor(1, eng.Pt, S)[0]
    solver = _cg_normaleq if approx == "CG_normaleq" else _fixed_point
    return solver(hvp, outer, new_noise, K, eta, tol)


def CG_normaleq(*args, **kwargs):
    raise NotImplementedError("the autograd-callable form of CG_normaleq (reference hypergradients.py:199-244) is "
                              "replaced by cg_normaleq_native, which PSVI.hyper_step drives; generic fp_map/outer_loss "
                              "callables have no fused-kernel equivalent")


def fixed_point(*args, **kwargs):
    raise NotImplementedError("replaced by fixed_point_native (see CG_normaleq)")
