"""Streaming PSVI path: the bilevel step, objectives and predictive pass for models that do not fit the fused
cluster engine -- medium-size mean-field MLPs and the full-covariance family (fn2).

Structure: the objectives only see the variational family through sample / tangent / kl / nkl
and their adjoints, so the heavy per-sample work -- forward, softmax-NLL, backward and the Hessian-vector (dual) pass of
the network on sampled weights -- is ONE native kernel (`psvi_net_pass`, one CTA per MC sample, weights and activations in
shared memory), the dense full-covariance products are native kernels on the packed triangle (`psvi_fc_matvec`,
`psvi_fc_outer`), and this file holds the host-side sequencing: the unrolled robust-Adam loop of
psvi/robust_higher/optim.py:303-367 and its reverse sweep, with the P-length elementwise updates
expressed as torch tensor ops on the device.  Nothing here runs on the CPU.

Reference map: PSVI.inner_elbo / psvi_elbo / nested_step / evaluate (psvi/inference/psvi_classes.py:488-511, 445-486,
541-600, 1031-1108); VIMixin (neural_net.py:60-173); MultivariateNormalVIMixin (neural_net.py:408-491).
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F

from psvi import _native

B1, B2 = 0.9, 0.999
OMB1, OMB2, AEPS = float(1.0 - B1), float(1.0 - B2), 1e-8


class MeanFieldFamily:
    """phi = [mu | rho] in theta layout (the flat buffers that back the module parameters).  `mask` (None = all ones) marks
    the parameters that enter the KL / sampled-nkl sums (the reference filters those sums on VILinear)."""

    mask = None

    def __init__(self, model):
        self.model = model
        self.mu, self.rho = model.flat()
        self.Pt = self.mu.numel()

    def _m(self, t):
        return t if self.mask is None else t * self.mask

    def fix_eps(self, eps):
        return eps

    def get_phi(self):
        return torch.cat([self.mu, self.rho])

    def set_phi(self, phi):
        self.mu.copy_(phi[:self.Pt])
        self.rho.copy_(phi[self.Pt:])

    # the [S][P] maps are fused native kernels (csrc/psvi_family.cu): each reads every slab once
    def sample(self, phi, eps):
        return _native.mf_sample(phi[:self.Pt], phi[self.Pt:], eps.contiguous())

    def tangent(self, phi, phidot, eps):
        return _native.mf_tangent(phi[self.Pt:], phidot[:self.Pt].contiguous(), phidot[self.Pt:].contiguous(), eps.contiguous())

    def kl(self, phi):
        mu, sg = phi[:self.Pt], F.softplus(phi[self.Pt:])
        return self._m(0.5 * (sg * sg + mu * mu - 1.0) - torch.log(sg)).sum()

    def nkl(self, phi, eps, theta):
        return _native.mf_nkl_kl(phi[:self.Pt], phi[self.Pt:], eps.contiguous(), theta, self.mask)[:eps.shape[0]]

    def nkl_theta_grad(self, theta):
        """d nkl_s / d theta_s through the sample."""
        return -self._m(theta)

    def grad(self, phi, eps, tbar, kl_coef, nkl_coef):
        return _native.mf_reparam_grad(phi[:self.Pt], phi[self.Pt:], eps.contiguous(), tbar, kl_coef, nkl_coef, mask=self.mask)

    def grad_with_nkl(self, phi, eps, tbar, beta, theta, beta_sum):
        """phi_bar of the outer objective: tbar plus the d nkl_s / d theta_s path (weights beta_s) and the log-sigma term.
        `beta_sum` = sum_s beta_s as a HOST number: it is -kappa identically (sum_s w_s = 1), so no device read-back is
        needed -- the step stays free of host synchronisation (CUDA-graph capturable)."""
        return _native.mf_reparam_grad(phi[:self.Pt], phi[self.Pt:], eps.contiguous(), tbar, 0.0, float(beta_sum), mask=self.mask,
                                       beta=beta.float().contiguous(), theta=theta)

    def hvp(self, phi, phidot, eps, A_t, A_td):
        return _native.mf_reparam_hvp(phi[self.Pt:], phidot[:self.Pt].contiguous(), phidot[self.Pt:].contiguous(), eps.contiguous(),
                                      A_t, A_td, mask=self.mask)


class LenetFamily(MeanFieldFamily):
    """Mean-field family of make_lenet (reference neural_net.py:334-359): conv layers carry no KL / nkl (Q5) and the last
    VILinear has mc_samples = 1, i.e. ONE noise draw shared by all samples (Q4): its block of every [S, P] noise slab is
    sample 0's row."""

    def __init__(self, model):
        super().__init__(model)
        self.mask = model.kl_mask()
        self.tail = model.shared_tail()

    def fix_eps(self, eps):
        if self.tail:
            eps = eps.clone()
            eps[..., :, -self.tail:] = eps[..., :1, -self.tail:]
        return eps


class FullCovFamily:
    """phi = per layer [mean | _sd | _corr] (torch parameters_to_vector order of the fn2 model)."""

    def __init__(self, model):
        self.model = model
        self.layers = model.vi_layers()
        self.ns = [m.num_params for m in self.layers]
        self.ncs = [m._corr.numel() for m in self.layers]
        self.Pt = sum(self.ns)
        self.offs, self.toffs, o, t = [], [], 0, 0
        for n, c in zip(self.ns, self.ncs):
            self.offs.append(o)
            self.toffs.append(t)
            o += 2 * n + c
            t += n
        # the per-layer launches of a family map are independent: while the step is being captured into a CUDA graph the small
        # layers go to side streams (parallel branches next to the one big layer); eagerly they stay in line (stream juggling
        # from Python would cost more than the microsecond kernels it overlaps)
        self._sides = [torch.cuda.Stream() for _ in self.ns] if torch.cuda.is_available() else []

    def _per_layer(self, fn):
        items = list(zip(self.offs, self.ns, self.toffs))
        if len(items) < 2 or not self._sides or not torch.cuda.is_current_stream_capturing():
            for it in items:
                fn(*it)
            return
        cur = torch.cuda.current_stream()
        big = max(range(len(items)), key=lambda i: self.ns[i])
        fork = torch.cuda.Event()
        fork.record(cur)
        joins = []
        for i, it in enumerate(items):
            if i == big:
                continue
            side = self._sides[i]
            side.wait_event(fork)
            with torch.cuda.stream(side):
                fn(*it)
                ev = torch.cuda.Event()
                ev.record(side)
            joins.append(ev)
        fn(*items[big])
        for ev in joins:
            cur.wait_event(ev)

    def fix_eps(self, eps):
        return eps

    def nkl_theta_grad(self, theta):
        return -theta

    def get_phi(self):
        return torch.cat([p.detach().reshape(-1).float() for m in self.layers for p in (m.mean, m._sd, m._corr)])

    def set_phi(self, phi):
        with torch.no_grad():
            for (m, sd, corr, n), layer in zip(self._split(phi), self.layers):
                layer.mean.copy_(m)
                layer._sd.copy_(sd)
                layer._corr.copy_(corr)

    def _split(self, phi):
        out = []
        for o, n, c in zip(self.offs, self.ns, self.ncs):
            out.append((phi[o:o + n], phi[o + n:o + 2 * n], phi[o + 2 * n:o + 2 * n + c], n))
        return out

    @staticmethod
    def _f32c(t):
        if not (t.is_cuda and t.dtype == torch.float32 and t.is_contiguous()):
            raise _native.NativeError("expected a contiguous CUDA float32 tensor")
        return t

    def _sample(self, phi, phidot, eps):
        S = eps.shape[0]
        phi, eps = self._f32c(phi), self._f32c(eps)
        pd = None if phidot is None else self._f32c(phidot).data_ptr()
        out = torch.empty(S, self.Pt, device=eps.device)
        self._per_layer(lambda o, n, t: _native.fc_sample(n, S, phi.data_ptr() + 4 * o, None if pd is None else pd + 4 * o,
                                                          eps.data_ptr() + 4 * t, self.Pt, out.data_ptr() + 4 * t, self.Pt))
        return out

    def sample(self, phi, eps):
        return self._sample(phi, None, eps)

    def tangent(self, phi, phidot, eps):
        return self._sample(phi, phidot, eps)

    def kl(self, phi):
        t = 0.0
        for (m, sd, corr, n) in self._split(phi):
            d = F.softplus(sd)
            t = t + 0.5 * ((d * d).sum() + (corr * corr).sum() + (m * m).sum() - n) - torch.log(d).sum()
        return t

    def nkl(self, phi, eps, theta):
        logdet = sum(torch.log(F.softplus(sd)).double().sum() for (_, sd, _, _) in self._split(phi))
        return -0.5 * (theta.double() ** 2).sum(1) + 0.5 * (eps.double() ** 2).sum(1) + logdet

    def grad(self, phi, eps, tbar, kl_coef, nkl_coef):
        S = eps.shape[0]
        phi, eps, tbar = self._f32c(phi), self._f32c(eps), self._f32c(tbar)
        g = torch.empty_like(phi)
        self._per_layer(lambda o, n, t: _native.fc_reparam_grad(n, S, phi.data_ptr() + 4 * o, tbar.data_ptr() + 4 * t, self.Pt,
                                                                eps.data_ptr() + 4 * t, self.Pt, kl_coef, nkl_coef,
                                                                g.data_ptr() + 4 * o))
        return g

    def hvp(self, phi, phidot, eps, A_t, A_td):
        S = eps.shape[0]
        phi, phidot, eps, A_t, A_td = (self._f32c(x) for x in (phi, phidot, eps, A_t, A_td))
        h = torch.empty_like(phi)
        self._per_layer(lambda o, n, t: _native.fc_reparam_hvp(n, S, phi.data_ptr() + 4 * o, phidot.data_ptr() + 4 * o,
                                                               A_t.data_ptr() + 4 * t, A_td.data_ptr() + 4 * t, self.Pt,
                                                               eps.data_ptr() + 4 * t, self.Pt, h.data_ptr() + 4 * o))
        return h


class MlpNet:
    """Per-sample MLP pass on sampled weights: psvi_net_pass / psvi_net_predict (one CTA per MC sample)."""

    def __init__(self, dims, S):
        self.desc = _native.make_model(dims, S)

    def pass_(self, theta, thetad, x, y, cw, **out):
        _native.net_pass(self.desc, theta, thetad, x, y, cw, **out)

    def predict(self, theta, lw, mode, xt, yt, out):
        _native.net_predict(self.desc, theta, lw, mode, xt, yt, out)

    def logits(self, theta, x):
        S, C = theta.shape[0], self.desc.dims[self.desc.n_layers]
        lg = torch.empty(S, x.shape[0], C, device=x.device)
        _native.net_pass(self.desc, theta, None, x, torch.zeros(x.shape[0], device=x.device, dtype=torch.int32), None, logits=lg)
        return lg


class GaussMlpNet(MlpNet):
    """The same per-sample MLP pass with the Gaussian likelihood of the regressors (psvi_net_pass_gaussian: one output, precision
    tau, float targets).  `wants_ybar`: the engine also collects d/dy -- the targets z of the pseudo-points are learnable
    (reference psvi_classes.py:2064-2087)."""

    wants_ybar = True

    def __init__(self, dims, S, tau):
        super().__init__(dims, S)
        self.tau = float(tau)

    def pass_(self, theta, thetad, x, y, cw, logits=None, **out):
        _native.net_pass_gaussian(self.desc, theta, thetad, x, y, cw, self.tau, outputs=logits, **out)

    def logits(self, theta, x):
        out = torch.empty(theta.shape[0], x.shape[0], 1, device=x.device)
        _native.net_pass_gaussian(self.desc, theta, None, x, torch.zeros(x.shape[0], device=x.device), None, self.tau, outputs=out)
        return out

    def predict(self, theta, lw, mode, xt, yt, out):
        raise NotImplementedError("the regressors evaluate through PSVI_regressor.evaluate (RMSE / log-likelihood)")


class FnLargeNet:
    """Per-sample pass of fn with one hidden layer in the large regime: batched TMA + tcgen05 GEMMs (csrc/psvi_fn_large.cu).
    `precision`: _native.PREC_TF32X3 (default; fp32-class accuracy, needed by the unrolled hypergradient), PREC_BF16X3
    (split-bf16 operand pairs: twice the MMA rate and half the operand bytes of tf32x3, ~2^-17 per operand; hypergradients
    within 1e-2..1e-1, cosine >= 0.999, of the fp64 oracle -- opt-in) or PREC_BF16 (6x the tensor rate, ~1e-2 relative error
    per pass: values, first-order training, prediction)."""

    def __init__(self, dims, S, precision=_native.PREC_TF32X3, precision_dual=None):
        self.desc, self.S, self.C, self.precision = _native.make_model(dims, S), S, dims[-1], precision
        # arithmetic of the Hessian-vector ("dual") passes of the reverse sweep.  Measured (profiles/r2_mixed_precision_study.md):
        # the hypergradient's sensitivity to operand rounding sits in the GRADIENT passes (their g_i become Adam's denominators
        # and fix the trajectory); split-bf16 dual passes leave it at the tf32x3 level.
        self.precision_dual = precision if precision_dual is None else precision_dual
        # ... and so is the outer objective's value / gradient pass (one per step): same study, row "outer pass = bf16x3"
        self.precision_outer = self.precision_dual
        self.phase = None          # StreamEngine.outer_grad sets "outer" around its passes

    @staticmethod
    def fits(dims, S):
        return len(dims) == 3 and dims[0] % 64 == 0 and dims[1] % 128 == 0 and dims[2] <= 16 and S <= 64

    def pass_(self, theta, thetad, x, y, cw, **out):
        prec = self.precision_dual if thetad is not None else (self.precision_outer if self.phase == "outer" else self.precision)
        _native.fnl_pass(self.desc, prec, theta, thetad, x, y, cw, **out)

    def logits(self, theta, x):
        lg = torch.empty(self.S, x.shape[0], self.C, device=x.device)
        _native.fnl_pass(self.desc, self.precision, theta, None, x, torch.zeros(x.shape[0], device=x.device, dtype=torch.int32),
                         None, logits=lg)
        return lg

    def predict(self, theta, lw, mode, xt, yt, out):
        _native.logits_predict(self.logits(theta, xt), lw, mode, yt, out)

    def data_grad(self, theta, x_bf16, y32, coef):
        """Data term of the outer objective over many rows on the fused tensor path (csrc/psvi_fn_grad_tc.cuh): returns
        (sum_r nll[s, r] [S], coef[s] * sum_r d nll[s, r] / d theta_s [S][P]).  bf16 operands, fp32 accumulation."""
        S, dev = self.S, theta.device
        dsum, tbar = torch.empty(S, device=dev), torch.empty(S, theta.shape[1], device=dev)
        n = _native.fn_data_grad_scratch_floats(self.desc, x_bf16.shape[0])
        ws = getattr(self, "_dg_ws", None)
        if ws is None or ws.numel() < n or ws.device != dev:
            ws = self._dg_ws = torch.empty(n, device=dev)
        _native.fn_data_grad_tc(self.desc, theta.contiguous(), x_bf16, y32, coef.contiguous(), dsum, tbar, ws)
        return dsum, tbar


class LenetNet:
    """Per-sample lenet pass: fused conv + ReLU + pool kernels and the fc kernels of csrc/psvi_lenet.cu."""

    def __init__(self, S):
        self.S = S

    def pass_(self, theta, thetad, x, y, cw, **out):
        _native.lenet_pass(self.S, theta, thetad, x, y, cw, **out)

    def logits(self, theta, x):
        lg = torch.empty(self.S, x.shape[0], 10, device=x.device)
        _native.lenet_pass(self.S, theta, None, x, torch.zeros(x.shape[0], device=x.device, dtype=torch.int32), None, logits=lg)
        return lg

    def predict(self, theta, lw, mode, xt, yt, out):
        _native.logits_predict(self.logits(theta, xt), lw, mode, yt, out)


class StreamEngine:
    def __init__(self, fam, dims, S, net=None):
        self.fam, self.S, self.dims = fam, S, list(dims)
        self.net = net if net is not None else MlpNet(dims, S)
        self.Pt = fam.Pt

    # ---- objectives --------------------------------------------------------------------------------------------------
    def inner_grad(self, phi, eps, u, z32, a, want_val=True, a_exp=None, fixed=False):
        """`a_exp` (a broadcast to [S, M], contiguous) and `fixed` (eps already went through fix_eps) let the unrolled loop
        hoist the per-call preparation; `want_val=False` skips the objective value."""
        S, M = self.S, u.shape[0]
        if not fixed:
            eps = self.fam.fix_eps(eps)
        theta = self.fam.sample(phi, eps)
        nll, tbar = torch.empty(S, M, device=u.device), torch.empty(S, self.Pt, device=u.device)
        self.net.pass_(theta, None, u, z32, a.expand(S, M).contiguous() if a_exp is None else a_exp, nll=nll, tbar=tbar)
        val = (nll.double() @ a.double()).sum() + self.fam.kl(phi).double() if want_val else None
        return val, self.fam.grad(phi, eps, tbar, 1.0, 0.0)

    ROW_CHUNK = 8192   # data rows per network pass when the minibatch is large (full-data term, sharded over ranks)

    def outer_grad(self, phi, eps, u, z32, a, xb, yb32, N, kappa=1.0, n_total=None, xb_bf16=None, data_w=None, extras=None):
        """psvi_elbo value and gradients.  `kappa` / `n_total` describe one rank's share when the data rows are sharded over
        R ranks: L_r = sum_s w_s (d_s^r - kappa p_s) - kappa mean(lw), kappa = 1/R, d_s^r = (N / n_total) * sum over
        this rank's rows; the shares (value and every gradient) add up to the unsharded objective because the importance
        weights depend on the pseudo-data only.  Large minibatches are processed in chunks of ROW_CHUNK rows.
        `data_w` [B] (optional): per-row weights of the data term, d_s = (N / n_total) sum_b data_w_b nll[s, b] (the soft-label
        rows of learn_z); `extras` then receives "dwbar" = dLoss / d data_w."""
        if getattr(self.net, "phase", 0) is None:      # nets with per-phase arithmetic (FnLargeNet): mark the outer passes
            self.net.phase = "outer"
            try:
                return self.outer_grad(phi, eps, u, z32, a, xb, yb32, N, kappa=kappa, n_total=n_total, xb_bf16=xb_bf16,
                                       data_w=data_w, extras=extras)
            finally:
                self.net.phase = None
        if xb_bf16 is not None:
            return self._outer_grad_fulldata(phi, eps, u, z32, a, xb_bf16, yb32, N, kappa, n_total)
        S, M, B, dev = self.S, u.shape[0], xb.shape[0], u.device
        n_total = B if n_total is None else n_total
        eps = self.fam.fix_eps(eps)
        theta = self.fam.sample(phi, eps)
        if M + B <= self.ROW_CHUNK:
            X, lab = torch.cat([u, xb]).contiguous(), torch.cat([z32, yb32]).contiguous()
            nll = torch.empty(S, M + B, device=dev)
            self.net.pass_(theta, None, X, lab, None, nll=nll)
            nd = nll.double()
            nll_u, ds_sum = nll[:, :M], (nd[:, M:].sum(1) if data_w is None else nd[:, M:] @ data_w.double())
        else:
            if data_w is not None:
                raise NotImplementedError("per-row data weights are built for minibatches of at most ROW_CHUNK rows")
            nll_u = torch.empty(S, M, device=dev)
            self.net.pass_(theta, None, u, z32, None, nll=nll_u)
            ds_sum = torch.zeros(S, device=dev, dtype=torch.float64)
            for r0 in range(0, B, self.ROW_CHUNK):
                xc, yc = xb[r0:r0 + self.ROW_CHUNK].contiguous(), yb32[r0:r0 + self.ROW_CHUNK].contiguous()
                nc = torch.empty(S, xc.shape[0], device=dev)
                self.net.pass_(theta, None, xc, yc, None, nll=nc)
                ds_sum += nc.double().sum(1)
        ps, ds = nll_u.double() @ a.double(), (N / n_total) * ds_sum
        lw = -ps + self.fam.nkl(phi, eps, theta)
        w = torch.softmax(lw, 0)
        e = ds - kappa * ps
        ebar = (w * e).sum()
        loss = ebar - kappa * lw.mean()
        beta = w * (e - ebar) - kappa / S
        gp = -kappa * w - beta
        cw_u = (gp[:, None] * a.double()[None, :]).float()
        wd = (w * N / n_total).float()
        tbar = torch.empty(S, self.Pt, device=dev)
        if M + B <= self.ROW_CHUNK:
            cw = torch.cat([cw_u, wd[:, None].expand(S, B) if data_w is None else wd[:, None] * data_w[None, :]], 1).contiguous()
            if extras is not None and data_w is not None:
                extras["dwbar"] = ((w * N / n_total) @ nd[:, M:]).float()
            xbar = torch.empty(S, M + B, X.shape[1], device=dev)
            yk = {}
            if extras is not None and getattr(self.net, "wants_ybar", False):
                yk["ybar"] = torch.empty(S, M + B, device=dev)
            self.net.pass_(theta, None, X, lab, cw, nll=nll, tbar=tbar, xbar=xbar, **yk)
            if yk:
                extras["zbar_outer"] = yk["ybar"][:, :M].sum(0)
            xbar_u, nll_u = xbar[:, :M], nll[:, :M]
        else:
            xbar_u = torch.empty(S, M, u.shape[1], device=dev)
            self.net.pass_(theta, None, u, z32, cw_u.contiguous(), nll=nll_u, tbar=tbar, xbar=xbar_u)
            tb = torch.empty_like(tbar)
            for r0 in range(0, B, self.ROW_CHUNK):
                xc, yc = xb[r0:r0 + self.ROW_CHUNK].contiguous(), yb32[r0:r0 + self.ROW_CHUNK].contiguous()
                nc = torch.empty(S, xc.shape[0], device=dev)
                self.net.pass_(theta, None, xc, yc, wd[:, None].expand(S, xc.shape[0]).contiguous(), nll=nc, tbar=tb)
                tbar += tb
        if hasattr(self.fam, "grad_with_nkl"):
            pbar = self.fam.grad_with_nkl(phi, eps, tbar, beta, theta, -kappa)
        else:
            tbar = tbar + beta.float()[:, None] * self.fam.nkl_theta_grad(theta)
            pbar = self.fam.grad(phi, eps, tbar, 0.0, -kappa)     # sum_s beta_s = -kappa identically
        return loss.float(), pbar, xbar_u.sum(0), (gp.float() @ nll_u), ds.float()

    def _outer_grad_fulldata(self, phi, eps, u, z32, a, xb_bf16, yb32, N, kappa, n_total):
        """outer_grad with the data rows (bf16, many) on the fused tensor path: the importance weights need the pseudo-data
        forward only, so the order is  pseudo values -> w -> ONE data pass (nll sums + w_s N / n_total weighted adjoints) ->
        beta, dL/dp -> pseudo gradient pass.  Same return values and sharding contract (kappa, n_total) as outer_grad."""
        S, M, B, dev = self.S, u.shape[0], xb_bf16.shape[0], u.device
        n_total = B if n_total is None else n_total
        eps = self.fam.fix_eps(eps)
        theta = self.fam.sample(phi, eps)
        nll_u = torch.empty(S, M, device=dev)
        self.net.pass_(theta, None, u, z32, None, nll=nll_u)
        ps = nll_u.double() @ a.double()
        lw = -ps + self.fam.nkl(phi, eps, theta)
        w = torch.softmax(lw, 0)
        wd = (w * N / n_total).float()
        if B > 0:
            dsum, tbar_d = self.net.data_grad(theta, xb_bf16, yb32, wd)
            ds = (N / n_total) * dsum.double()
        else:   # a rank without rows still takes part in the exchange
            tbar_d, ds = None, torch.zeros(S, device=dev, dtype=torch.float64)
        e = ds - kappa * ps
        ebar = (w * e).sum()
        loss = ebar - kappa * lw.mean()
        beta = w * (e - ebar) - kappa / S
        gp = -kappa * w - beta
        cw_u = (gp[:, None] * a.double()[None, :]).float()
        tbar = torch.empty(S, self.Pt, device=dev)
        xbar_u = torch.empty(S, M, u.shape[1], device=dev)
        self.net.pass_(theta, None, u, z32, cw_u.contiguous(), nll=nll_u, tbar=tbar, xbar=xbar_u)
        if tbar_d is not None:
            tbar += tbar_d
        if hasattr(self.fam, "grad_with_nkl"):
            pbar = self.fam.grad_with_nkl(phi, eps, tbar, beta, theta, -kappa)
        else:
            tbar = tbar + beta.float()[:, None] * self.fam.nkl_theta_grad(theta)
            pbar = self.fam.grad(phi, eps, tbar, 0.0, -kappa)     # sum_s beta_s = -kappa identically
        return loss.float(), pbar, xbar_u.sum(0), (gp.float() @ nll_u), ds.float()

    def outer_grad_ablated(self, phi, eps, xb, yb32, N, kappa=1.0, n_total=None):
        """PSVI_Ablated.psvi_elbo (reference psvi_classes.py:1397-1408): mean_s (N/B) sum_b nll[s, b] - mean_s sampled_nkl_s --
        no importance weights and no pseudo-data term (its direct partials wrt u, v are zero).  kappa / n_total: one rank's
        share when the data rows are sharded (the nkl term is counted kappa times)."""
        S, B, dev = self.S, xb.shape[0], xb.device
        n_total = B if n_total is None else n_total
        eps = self.fam.fix_eps(eps)
        theta = self.fam.sample(phi, eps)
        tbar, tb = torch.zeros(S, self.Pt, device=dev), torch.empty(S, self.Pt, device=dev)
        ds = torch.zeros(S, device=dev, dtype=torch.float64)
        for r0 in range(0, B, self.ROW_CHUNK):
            xc, yc = xb[r0:r0 + self.ROW_CHUNK].contiguous(), yb32[r0:r0 + self.ROW_CHUNK].contiguous()
            nc = torch.empty(S, xc.shape[0], device=dev)
            cw = torch.full((S, xc.shape[0]), N / n_total / S, device=dev)
            self.net.pass_(theta, None, xc, yc, cw, nll=nc, tbar=tb)
            tbar += tb
            ds += nc.double().sum(1)
        nkl = self.fam.nkl(phi, eps, theta)
        loss = (N / n_total) * ds.mean() - kappa * nkl.mean()
        beta = torch.full((S,), -kappa / S, device=dev)
        if hasattr(self.fam, "grad_with_nkl"):
            pbar = self.fam.grad_with_nkl(phi, eps, tbar, beta, theta, -kappa)
        else:
            pbar = self.fam.grad(phi, eps, tbar + beta[:, None] * self.fam.nkl_theta_grad(theta), 0.0, -kappa)
        return loss.float(), pbar

    def hvp(self, phi, eps, u, z32, a, phidot, a_exp=None, fixed=False):
        S, M, dev = self.S, u.shape[0], u.device
        if not fixed:
            eps = self.fam.fix_eps(eps)
        theta, thetad = self.fam.sample(phi, eps), self.fam.tangent(phi, phidot, eps)
        tbar, tdbar = torch.empty(S, self.Pt, device=dev), torch.empty(S, self.Pt, device=dev)
        xbar, ac = torch.empty(S, M, u.shape[1], device=dev), torch.empty(S, M, device=dev)
        yk = {"ybar": torch.empty(S, M, device=dev)} if getattr(self.net, "wants_ybar", False) else {}
        self.net.pass_(theta, thetad, u, z32, a.expand(S, M).contiguous() if a_exp is None else a_exp, tbar=tbar, tdbar=tdbar,
                       xbar=xbar, acbar=ac, **yk)
        self._last_hz = yk["ybar"].sum(0) if yk else None       # mixed derivative wrt the (learnable) targets
        return self.fam.hvp(phi, phidot, eps, tbar, tdbar), xbar.sum(0), ac.sum(0)

    # ---- unrolled robust Adam + reverse sweep (optim.py:303-367) -----------------------------------------
    def nested(self, phi, eps_all, u, z32, a, xb, yb32, N, T, lr, want_losses=False, kappa=1.0, n_total=None, reduce_fn=None,
               outer="psvi", xb_bf16=None, a_outer=None, data_w=None, extras=None):
        """eps_all [T+1, S, P].  Returns loss, ubar [M,D], abar [M], phi_T, inner losses (list or None).
        Sharded data term: pass this rank's rows with kappa = 1/world, n_total = rows over all ranks and a `reduce_fn` that
        all-reduces (loss, pbar, ubar, abar) -- the ONE exchange step of the bilevel step; the inner loop and
        the reverse sweep are replicated (identical seeds => identical trajectories on every rank).
        `a_outer` (pseudo-row weights of the OUTER objective when they differ from the inner ones), `data_w` and `extras`
        serve the soft-label rows of learn_z: extras gets "abar_outer" (the outer objective's direct dLoss/da) and "dwbar"."""
        eps_all = self.fam.fix_eps(eps_all)
        a_exp = a.expand(self.S, u.shape[0]).contiguous()
        phi = phi.contiguous()
        m, v = torch.zeros_like(phi), torch.zeros_like(phi)
        traj, losses = [], []
        for t in range(T):
            val, g = self.inner_grad(phi, eps_all[t], u, z32, a, want_val=want_losses, a_exp=a_exp, fixed=True)
            if want_losses:
                losses.append(val)
            phi_new, m, v = _native.adam_unroll_step(phi, g, m, v, lr / (1.0 - B1 ** (t + 1)), math.sqrt(1.0 - B2 ** (t + 1)))
            traj.append((phi, g, m, v))
            phi = phi_new
        if outer == "ablated":
            loss, pbar = self.outer_grad_ablated(phi, eps_all[T], xb, yb32, N, kappa=kappa, n_total=n_total)
            ubar, abar = torch.zeros_like(u), torch.zeros_like(a)
        else:
            loss, pbar, ubar, abar, _ = self.outer_grad(phi, eps_all[T], u, z32, a if a_outer is None else a_outer, xb, yb32, N,
                                                        kappa=kappa, n_total=n_total, xb_bf16=xb_bf16, data_w=data_w,
                                                        extras=extras)
            if extras is not None:
                extras["abar_outer"] = abar.clone()
        if reduce_fn is not None:
            loss, pbar, ubar, abar = reduce_fn(loss, pbar, ubar, abar)
        phi_T = phi
        pbar = pbar.contiguous()
        mbar, vbar = torch.zeros_like(pbar), torch.zeros_like(pbar)
        for t in range(T - 1, -1, -1):
            phi_t, g, m_t, v_t = traj[t]
            gbar = _native.adam_unroll_reverse(pbar, g, m_t, v_t, mbar, vbar, lr / (1.0 - B1 ** (t + 1)),
                                               math.sqrt(1.0 - B2 ** (t + 1)))
            h, hu, ha = self.hvp(phi_t, eps_all[t], u, z32, a, gbar, a_exp=a_exp, fixed=True)
            pbar, ubar, abar = pbar + h, ubar + hu, abar + ha
            if extras is not None and self._last_hz is not None:
                extras["zbar"] = extras.get("zbar", extras.get("zbar_outer", 0.0)) + self._last_hz
        return loss, ubar, abar, phi_T, (torch.stack(losses).float() if want_losses else None)

    # ---- the same step as ONE CUDA graph ---------------------------------------------------------------------------------
    # A bilevel step on this path is hundreds of kernel launches sequenced by Python (676 at BASELINE cfg5, 1365 at cfg4, 596
    # at cfg3): ~45 us of host work per launch, i.e. a host-bound step as soon as the kernels get faster.  The step has no
    # data-dependent control flow and no host synchronisation, so it is captured once per (shapes, T, lr, ...) into a CUDA
    # graph -- every launch of libpsvi_b200 goes to torch's current stream, which is the capturing stream -- and replayed:
    # inputs are copied into the graph's static buffers, the outputs are cloned out.  Falls back to the eager sequence when
    # the ranks exchange data inside the step (reduce_fn), when capture fails, or with PSVI_NO_GRAPH=1.
    use_graphs = True
    MAX_GRAPHS = 4

    def nested_cached(self, phi, eps_all, u, z32, a, xb, yb32, N, T, lr, want_losses=False, kappa=1.0, n_total=None,
                      reduce_fn=None, outer="psvi", xb_bf16=None, a_outer=None, data_w=None, extras=None):
        import os
        scal = dict(N=N, T=T, lr=lr, want_losses=want_losses, kappa=kappa, n_total=n_total, outer=outer)
        tens = dict(phi=phi, eps_all=eps_all, u=u, z32=z32, a=a, xb=xb, yb32=yb32, xb_bf16=xb_bf16, a_outer=a_outer, data_w=data_w)
        # (a full-data term -- xb_bf16, millions of rows -- is device-bound and would only be copied around: eager)
        if reduce_fn is not None or xb_bf16 is not None or not self.use_graphs or os.environ.get("PSVI_NO_GRAPH"):
            return self.nested(**tens, **scal, reduce_fn=reduce_fn, extras=extras)
        graphs = self.__dict__.setdefault("_graphs", {})
        key = (T, float(lr), float(N), bool(want_losses), float(kappa), n_total, outer, extras is not None,
               tuple((k, None if t is None else (tuple(t.shape), t.dtype)) for k, t in tens.items()))
        ent = graphs.get(key)
        if ent is None:
            if len(graphs) >= self.MAX_GRAPHS:
                graphs.pop(next(iter(graphs)))
            try:
                ent = self._capture_nested(tens, scal, extras is not None)
            except Exception as e:      # e.g. a host synchronisation inside a family map: keep the eager sequence for this key
                import warnings
                warnings.warn(f"CUDA-graph capture of the bilevel step failed ({e!r:.200}); running it eagerly")
                ent = False
            graphs[key] = ent
        if ent is False:
            return self.nested(**tens, **scal, extras=extras)
        static_in, graph, outs, ex = ent
        for k, t in tens.items():
            if t is not None:
                static_in[k].copy_(t)
        graph.replay()
        if extras is not None:
            extras.update({k: v.clone() for k, v in ex.items()})
        return tuple(None if o is None else o.clone() for o in outs)

    def _capture_nested(self, tens, scal, want_extras):
        static_in = {k: (None if t is None else t.detach().clone().contiguous()) for k, t in tens.items()}
        cur = torch.cuda.current_stream()
        side = torch.cuda.Stream()
        side.wait_stream(cur)
        with torch.cuda.stream(side):       # warm-up outside capture: workspaces, function attributes, lazy module loads
            self.nested(**static_in, **scal, extras={} if want_extras else None)
        cur.wait_stream(side)
        torch.cuda.synchronize()
        graph = torch.cuda.CUDAGraph()
        ex = {} if want_extras else None
        with torch.cuda.graph(graph):
            outs = self.nested(**static_in, **scal, extras=ex)
        return static_in, graph, outs, ex

    def predict_probs(self, phi, eps, u, z32, a, x, correction=True, chunk=8192):
        """Predictive class probabilities of rows x under ONE noise slab (PSVI.pred_on_grid, psvi_classes.py:1130-1175):
        importance-weighted mixture sum_s w_s softmax(logits_s) (w from the pseudo-data, sign quirk Q3) or the plain mean."""
        eps = self.fam.fix_eps(eps)
        theta = self.fam.sample(phi, eps)
        lw = None
        if correction:
            nll = torch.empty(self.S, u.shape[0], device=x.device)
            self.net.pass_(theta, None, u, z32, None, nll=nll)
            lw = ((nll.double() @ a.double()) + self.fam.nkl(phi, eps, theta)).float().contiguous()
        probs, out = [], torch.zeros(8, device=x.device)
        for r0 in range(0, x.shape[0], chunk):
            xc = x[r0:r0 + chunk].contiguous()
            lg = self.net.logits(theta, xc)
            pc = torch.empty(xc.shape[0], lg.shape[2], device=x.device)
            _native.logits_predict(lg, lw, 0 if correction else 1, None, out, probs_out=pc)
            probs.append(pc)
        return torch.cat(probs)

    # ---- predictive pass (psvi_classes.py:1031-1108) -----------------------------------------------------------------
    def evaluate(self, phi, eps_slabs, u, z32, a, xt, yt32, batch, mode=0):
        """eps_slabs [n_slabs, S, P]; returns out[8] accumulated over slabs (diagnostics of the last slab, Q12)."""
        dev, S = xt.device, self.S
        eps_slabs = self.fam.fix_eps(eps_slabs)
        tot = torch.zeros(8, device=dev)
        out = torch.zeros(8, device=dev)
        n = xt.shape[0]
        for k, r0 in enumerate(range(0, n, batch)):
            theta = self.fam.sample(phi, eps_slabs[k])
            lw = None
            if mode == 0 and (u is None or u.shape[0] == 0):
                lw = self.fam.nkl(phi, eps_slabs[k], theta).float().contiguous()     # no pseudo term: weights from nkl alone
            elif mode == 0:
                M = u.shape[0]
                nll = torch.empty(S, M, device=dev)
                self.net.pass_(theta, None, u, z32, None, nll=nll)
                lw = ((nll.double() @ a.double()) + self.fam.nkl(phi, eps_slabs[k], theta)).float().contiguous()  # Q3
            self.net.predict(theta, lw, mode, xt[r0:r0 + batch].contiguous(), yt32[r0:r0 + batch].contiguous(), out)
            tot[:3] += out[:3]
            tot[3:5] = out[3:5]
            if mode != 0 and r0 + batch >= n and u is not None and u.shape[0] > 0:
                # correction=False still reports the importance-weight diagnostics of the LAST batch (reference
                # psvi_classes.py:1047-1057,1085-1092: the weights are computed regardless of `correction`)
                M = u.shape[0]
                nll = torch.empty(S, M, device=dev)
                self.net.pass_(theta, None, u, z32, None, nll=nll)
                lwl = (nll.double() @ a.double()) + self.fam.nkl(phi, eps_slabs[k], theta)
                w = torch.softmax(lwl, 0)
                tot[3] = -(w * torch.log(w.clamp_min(1e-300))).sum().float()
                tot[4] = (w.sum() ** 2 / (w * w).sum() / S).float()
        return tot
