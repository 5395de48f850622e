"""Bayesian layers and model factories of the PSVI hot path -- same names, constructor arguments, parameter names and
RNG-consumption order as the reference's psvi/models/neural_net.py, with the math executed by libpsvi_b200 (CUDA).

Reference map (psvi/models/neural_net.py):  VIMixin :60-173, VILinear :176-179, make_fcnet :267-297,
categorical_fn :22-23, set_mc_samples :26-29, inverse_softplus :32-35.  `state_dict`s interchange with the reference
(parameter names weight, bias, _weight_sd, _bias_sd; module names lin{i}, nonl{i}, classifier).

What is native here: a whole mean-field MLP (nn.Sequential of VILinear / ReLU, as built by make_fcnet or by
PSVI.set_up_model for `logistic_regression`) is evaluated by ONE fused kernel (sampling -> per-sample GEMMs -> ReLU ->
logits), see MeanFieldMLP.forward.  Gradients are not obtained with autograd: the PSVI objectives, their gradients and
the hypergradient are separate fused kernels driven by psvi.inference.psvi_classes.
The full-covariance family (fn2) evaluates through the streaming path (packed-triangle products + the per-sample
network kernel).  The convolutional family (lenet: VIConv2d / BatchMaxPool2d / make_lenet, reference :194-255,334-359)
evaluates through the same streaming path with the fused conv + ReLU + pool kernels of csrc/psvi_lenet.cu.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from psvi import _native


def categorical_fn(logits=None, probs=None):
    # reference neural_net.py:22-23 (kept for callers that pass `distr_fn`; the fused kernels implement it directly)
    return torch.distributions.Categorical(logits=logits, probs=probs)


def gaussian_fn(loc=None, scale=None):
    return torch.distributions.normal.Normal(loc, scale)


def set_mc_samples(net, mc_samples):
    for module in net.modules():
        if isinstance(module, VIMixin):
            module.mc_samples = mc_samples


def inverse_softplus(x):
    if torch.is_tensor(x):
        return x.expm1().log()
    return np.log(np.expm1(x))


class _NoiseCounter:
    """Hands out Philox stream ids ("domains") so that no two native calls of a process reuse noise."""
    value = 0

    @classmethod
    def next(cls):
        cls.value += 1
        return cls.value


class VIMixin(nn.Module):
    """Mean-field Gaussian over weight and bias; sigma = softplus(rho); prior N(0, prior_sd) -- reference :60-173."""

    def __init__(self, *args, init_sd=0.01, prior_sd=1.0, mc_samples=1, **kwargs):
        super().__init__(*args, **kwargs)
        self._weight_sd = nn.Parameter(inverse_softplus(torch.full_like(self.weight, init_sd)))
        if self.bias is not None:
            self._bias_sd = nn.Parameter(inverse_softplus(torch.full_like(self.bias, init_sd)))
        else:
            self.register_parameter("_bias_sd", None)
        self.prior_sd, self.mc_samples, self._cached_weight, self._cached_bias, self._init_sd = (
            prior_sd, mc_samples, None, None, init_sd)
        self.reset_parameters_variational()

    def reset_parameters_variational(self) -> None:
        super().reset_parameters()  # second nn.Linear init draw, as in the reference (:87-88)
        self._weight_sd.data.copy_(inverse_softplus(torch.full_like(self.weight, self._init_sd)))
        if self.bias is not None:
            self._bias_sd.data.copy_(inverse_softplus(torch.full_like(self.bias, self._init_sd)))
        self._cached_weight, self._cached_bias = None, None

    @property
    def weight_sd(self):
        return F.softplus(self._weight_sd)

    @property
    def bias_sd(self):
        return F.softplus(self._bias_sd) if self.bias is not None else None

    def kl(self):
        """KL(q || N(0, prior_sd)) in closed form (reference :101-108 via torch.distributions)."""
        def one(mu, sd):
            ps = self.prior_sd
            return (0.5 * ((sd * sd + mu * mu) / (ps * ps) - 1.0) - torch.log(sd / ps)).sum()
        out = one(self.weight, self.weight_sd)
        if self.bias is not None:
            out = out + one(self.bias, self.bias_sd)
        return out

    def sampled_nkl(self):
        """log p(theta_s) - log q(theta_s) of the cached sample (reference :110-115) -> [S] (scalar if mc_samples==1)."""
        if self._cached_weight is None:
            raise RuntimeError("sampled_nkl() needs a forward pass first")
        c = 0.5 * float(np.log(2 * np.pi))

        def one(th, mu, sd, nd):
            ps = self.prior_sd
            lp = -0.5 * (th / ps) ** 2 - float(np.log(ps)) - c
            lq = -0.5 * ((th - mu) / sd) ** 2 - torch.log(sd) - c
            return (lp - lq).flatten(-nd).sum(-1)
        out = one(self._cached_weight, self.weight, self.weight_sd, self.weight.ndim)
        if self.bias is not None:
            b = self._cached_bias.squeeze(1) if self.mc_samples > 1 else self._cached_bias
            out = out + one(b, self.bias, self.bias_sd, self.bias.ndim)
        return out

    @property
    def weight_batch_shape(self):
        return torch.Size((self.mc_samples,) if self.mc_samples > 1 else ())

    @property
    def bias_batch_shape(self):
        return torch.Size((self.mc_samples, 1) if self.mc_samples > 1 else ())

    def extra_repr(self):
        return f"{super().extra_repr()}, mc_samples={self.mc_samples}"


class VILinear(VIMixin, nn.Linear):
    """reference :176-179.  A lone VILinear evaluates through the same fused kernel as a 1-layer MeanFieldMLP."""

    def rsample(self):
        _forward_stack([self], torch.zeros(1, self.in_features, device=self.weight.device))
        return self._cached_weight, self._cached_bias

    def forward(self, x):
        if x.dim() != 2:
            raise NotImplementedError(
                "a stand-alone VILinear takes [rows, in_features] inputs; per-sample [S, rows, in] activations only "
                "occur inside a stack, which MeanFieldMLP (make_fcnet) evaluates as one fused kernel")
        return _forward_stack([self], x)


def _layer_tensors(layers):
    mu = torch.cat([t.detach().reshape(-1) for m in layers for t in (m.weight, m.bias)]).float().contiguous()
    rho = torch.cat([t.detach().reshape(-1) for m in layers for t in (m._weight_sd, m._bias_sd)]).float().contiguous()
    return mu, rho


def _forward_stack(layers, x, noise=None, flat=None):
    """Sampled forward of a stack of VILinear layers (ReLU between) on the GPU: returns logits [S, R, C] ([R, C] if
    mc_samples == 1) and fills every layer's _cached_weight / _cached_bias with views of the sampled weights."""
    _native.require_cuda()
    S = int(layers[0].mc_samples)
    for m in layers:
        if m.bias is None:
            raise NotImplementedError("bias=False VILinear layers are not supported by the fused kernels")
        if float(m.prior_sd) != 1.0:
            raise NotImplementedError("prior_sd != 1 is not supported by the fused kernels")
        if int(m.mc_samples) != S:
            raise ValueError("all VI layers of a stack must share mc_samples")
    dims = [layers[0].in_features] + [m.out_features for m in layers]
    model = _native.make_model(dims, max(S, 1))
    mu, rho = flat if flat is not None else _layer_tensors(layers)
    x = x.detach().to(device=mu.device, dtype=torch.float32).contiguous()
    R, P = x.shape[0], mu.numel()
    logits = torch.empty(max(S, 1), R, dims[-1], device=mu.device)
    theta = torch.empty(max(S, 1), P, device=mu.device)
    if noise is None:
        noise = _native.make_noise(None, seed=torch.initial_seed(), domain=_NoiseCounter.next())
    _native.forward(model, noise, mu, rho, x, logits, theta_out=theta)
    off = 0
    for m in layers:
        nw, nb = m.weight.numel(), m.bias.numel()
        w = theta[:, off:off + nw].view(max(S, 1), *m.weight.shape)
        b = theta[:, off + nw:off + nw + nb].view(max(S, 1), 1, nb)
        m._cached_weight, m._cached_bias = (w, b) if S > 1 else (w[0], b[0, 0])
        off += nw + nb
    return logits if S > 1 else logits[0]


class MeanFieldMLP(nn.Sequential):
    """nn.Sequential of VILinear / ReLU evaluated by one fused kernel.  Also owns the flat (mu, rho) device buffers the
    PSVI kernels update in place; the modules' nn.Parameters are views into them, so `state_dict`, `parameters()` and
    torch optimisers keep working unchanged."""

    def vi_layers(self):
        return [m for m in self if isinstance(m, VILinear)]

    def check_supported(self):
        mods = list(self)
        ok = len(mods) >= 1 and isinstance(mods[-1], VILinear)
        for i, m in enumerate(mods[:-1]):
            ok = ok and (isinstance(m, VILinear) if i % 2 == 0 else isinstance(m, nn.ReLU))
        if not ok or len(self.vi_layers()) > _native.MAX_LAYERS:
            raise NotImplementedError("fused kernels cover VILinear (ReLU VILinear)* stacks with at most "
                                      f"{_native.MAX_LAYERS} VI layers; got {self}")

    @property
    def dims(self):
        ls = self.vi_layers()
        return [ls[0].in_features] + [m.out_features for m in ls]

    def n_samples(self):
        return int(self.vi_layers()[0].mc_samples)

    def flat(self):
        """(mu, rho): contiguous fp32 device tensors in theta layout whose storage backs the layer parameters."""
        ls = self.vi_layers()
        f = getattr(self, "_flat", None)
        if f is not None:
            ok, off = f[0].device == ls[0].weight.device, 0
            for m in ls:
                for t, base in ((m.weight, f[0]), (m.bias, f[0])):
                    ok = ok and t.data_ptr() == base.data_ptr() + 4 * off
                    off += t.numel()
            off = 0
            for m in ls:
                for t in (m._weight_sd, m._bias_sd):
                    ok = ok and t.data_ptr() == f[1].data_ptr() + 4 * off
                    off += t.numel()
            if ok:
                return f
        mu, rho = _layer_tensors(ls)
        off = 0
        for m in ls:
            for t_mu, t_rho in ((m.weight, m._weight_sd), (m.bias, m._bias_sd)):
                n = t_mu.numel()
                t_mu.data = mu[off:off + n].view(t_mu.shape)
                t_rho.data = rho[off:off + n].view(t_rho.shape)
                off += n
        object.__setattr__(self, "_flat", (mu, rho))
        return self._flat

    def forward(self, x, noise=None):
        self.check_supported()
        return _forward_stack(self.vi_layers(), x, noise=noise, flat=self.flat() if self.vi_layers()[0].weight.is_cuda else None)


def make_fcnet(in_dim, h_dim, out_dim, n_layers=2, linear_class=None, nonl_class=None, mc_samples=4, residual=False,
               **kwargs):
    """reference :267-297 (module names lin{i}, nonl{i}, classifier; default n_layers=2)."""
    if linear_class is None:
        linear_class = VILinear
    if nonl_class is None:
        nonl_class = nn.ReLU
    net = MeanFieldMLP() if (linear_class is VILinear and nonl_class is nn.ReLU) else nn.Sequential()
    for i in range(n_layers):
        net.add_module(f"lin{i}", linear_class(in_dim if i == 0 else h_dim, h_dim, **kwargs))
        net.add_module(f"nonl{i}", nonl_class())
    net.add_module("classifier", linear_class(h_dim, out_dim, **kwargs))
    for module in net.modules():
        module.mc_samples = mc_samples
    return net


def make_logistic_regression(in_dim, out_dim, **kwargs):
    """nn.Sequential(VILinear(D, nc, ...)) of PSVI.set_up_model (psvi_classes.py:694-699)."""
    return MeanFieldMLP(VILinear(in_dim, out_dim, **kwargs))


# ---- full-covariance family (fn2): reference neural_net.py:408-524 ------------------------------------------------------
class MultivariateNormalVIMixin(nn.Module):
    """Full-covariance Gaussian over ALL parameters of a layer jointly (reference :408-482): theta_s = mean + L eps_s with
    L = scale_tril: diag = softplus(_sd); the strictly-lower entries of the top-left (n-1)x(n-1) block are `_corr`
    (torch.tril_indices(n-1, n-1, -1) order), the last row has no off-diagonals (Q6).  Parameter names mean / _sd / _corr
    and their sizes match the reference, so state_dicts interchange."""

    def __init__(self, *args, init_sd=0.01, prior_sd=1.0, mc_samples=1, **kwargs):
        super().__init__(*args, **kwargs)   # nn.Linear init: consumes the RNG exactly as the reference does
        self.mc_samples, self.prior_sd = mc_samples, prior_sd
        self.param_names, self.param_shapes = [], []
        for n, p in list(self.named_parameters()):
            self.param_names.append(n)
            self.param_shapes.append(p.shape)
            delattr(self, n)
        self.param_numels = [int(np.prod(s)) for s in self.param_shapes]
        n = sum(self.param_numels)
        self.mean = nn.Parameter(p.new_zeros(n))
        self._sd = nn.Parameter(inverse_softplus(p.new_full((n,), init_sd)))
        self._corr = nn.Parameter(p.new_zeros(torch.tril_indices(n - 1, n - 1, offset=-1)[0].numel()))
        self.num_params = n
        self._cached = None

    def reset_parameters_variational(self) -> None:
        raise NotImplementedError

    @property
    def scale_tril(self):
        """Dense L, built the way the reference does (:452-461) -- for inspection only; the kernels keep L packed."""
        k = self.mean.new_zeros(self.num_params, self.num_params)
        k[torch.arange(self.num_params), torch.arange(self.num_params)] = F.softplus(self._sd)
        d = self.mean.size(-1) - 1
        i = torch.tril_indices(d, d, offset=-1)
        k[i[0], i[1]] = self._corr
        return k

    def kl(self):
        d = F.softplus(self._sd)
        return 0.5 * ((d * d).sum() + (self._corr ** 2).sum() + (self.mean ** 2).sum() - self.num_params) - torch.log(d).sum()

    def sampled_nkl(self):
        if self._cached is None:
            raise RuntimeError("sampled_nkl() needs a forward pass first")
        theta, eps = self._cached
        return -0.5 * (theta ** 2).sum(-1) + 0.5 * (eps ** 2).sum(-1) + torch.log(F.softplus(self._sd)).sum()


class VILinearMultivariateNormal(MultivariateNormalVIMixin, nn.Linear):
    """reference :485-491.  Evaluated inside a FullCovMLP stack (one fused per-sample network kernel)."""

    def extra_repr(self):
        return f"num_params={self.num_params}, mc_samples={self.mc_samples}"

    def forward(self, x, **kwargs):
        if x.dim() != 2:
            raise NotImplementedError("a stand-alone VILinearMultivariateNormal takes [rows, in_features] inputs")
        return FullCovMLP(self).forward(x)


class FullCovMLP(nn.Sequential):
    """nn.Sequential of VILinearMultivariateNormal / ReLU (fn2).  Forward: packed L eps products (psvi_fc_matvec) to
    sample every layer's weights, then one per-sample network kernel (psvi_net_pass)."""

    def vi_layers(self):
        return [m for m in self if isinstance(m, VILinearMultivariateNormal)]

    def check_supported(self):
        mods = list(self)
        ok = len(mods) >= 1 and isinstance(mods[-1], VILinearMultivariateNormal)
        for i, m in enumerate(mods[:-1]):
            ok = ok and (isinstance(m, VILinearMultivariateNormal) if i % 2 == 0 else isinstance(m, nn.ReLU))
        if not ok or len(self.vi_layers()) > _native.MAX_LAYERS:
            raise NotImplementedError(f"fused kernels cover VILinearMultivariateNormal (ReLU ...)* stacks; got {self}")
        for m in self.vi_layers():
            if float(m.prior_sd) != 1.0:
                raise NotImplementedError("prior_sd != 1 is not supported by the fused kernels")

    @property
    def dims(self):
        shapes = [m.param_shapes[0] for m in self.vi_layers()]      # weight shapes [out, in]
        return [int(shapes[0][1])] + [int(s[0]) for s in shapes]

    def n_samples(self):
        return int(self.vi_layers()[0].mc_samples)

    def forward(self, x, eps=None):
        from psvi.inference.stream import FullCovFamily
        _native.require_cuda()
        self.check_supported()
        fam = FullCovFamily(self)
        S = max(self.n_samples(), 1)
        dev = self.vi_layers()[0].mean.device
        if eps is None:
            eps = torch.empty(1, S, fam.Pt, device=dev)
            _native.philox_normal(torch.initial_seed(), _NoiseCounter.next(), 0, 1, S, fam.Pt, eps)
            eps = eps[0]
        theta = fam.sample(fam.get_phi(), eps)
        x = x.detach().to(dev, torch.float32).contiguous()
        logits = torch.empty(S, x.shape[0], self.dims[-1], device=dev)
        y = torch.zeros(x.shape[0], device=dev, dtype=torch.int32)
        _native.net_pass(_native.make_model(self.dims, S), theta, None, x, y, None, logits=logits)
        for m, t in zip(fam.layers, fam.toffs):
            m._cached = (theta[:, t:t + m.num_params], eps[:, t:t + m.num_params])
        return logits if self.n_samples() > 1 else logits[0]


def make_fc2net(in_dim, h_dim, out_dim, n_layers=2, linear_class=None, nonl_class=None, mc_samples=4, residual=False,
                **kwargs):
    """reference :494-524 (default n_layers=2 hidden layers, Q7)."""
    if linear_class is None:
        linear_class = VILinearMultivariateNormal
    if nonl_class is None:
        nonl_class = nn.ReLU
    net = FullCovMLP() if (linear_class is VILinearMultivariateNormal and nonl_class is nn.ReLU) else nn.Sequential()
    for i in range(n_layers):
        net.add_module(f"lin{i}", linear_class(in_dim if i == 0 else h_dim, h_dim, mc_samples=mc_samples, **kwargs))
        net.add_module(f"nonl{i}", nonl_class())
    net.add_module("classifier", linear_class(h_dim, out_dim, mc_samples=mc_samples, **kwargs))
    for module in net.modules():
        module.mc_samples = mc_samples
    return net


# ---- convolutional family (lenet): reference neural_net.py:194-255,334-359 ---------------------------------------------
class VIConv2d(VIMixin, nn.Conv2d):
    """reference :194-246 (per-sample convolution; the reference stacks the S samples on the channel axis and uses
    groups = S).  Evaluated inside a MeanFieldLeNet stack by the fused conv + ReLU + pool kernels of libpsvi_b200."""

    def __init__(self, *args, **kwargs):
        if "groups" in kwargs:
            raise ValueError("Cannot use groups argument for variational conv layer as this is used for parallelizing "
                             "across samples.")
        super().__init__(*args, **kwargs)

    def forward(self, x):
        raise NotImplementedError("a stand-alone VIConv2d is not evaluated on its own: the CUDA path covers the lenet stack "
                                  "(make_lenet), whose conv + ReLU + pool stages are fused kernels")


class BatchMaxPool2d(nn.MaxPool2d):
    """reference :249-255 (2x2 max-pool over the flattened (S, N) batch); fused into the conv kernels of the lenet stack."""


class MeanFieldLeNet(MeanFieldMLP):
    """nn.Sequential built by make_lenet, evaluated by libpsvi_b200's lenet pass (psvi_lenet_pass).  The flat (mu, rho)
    buffers of MeanFieldMLP back the parameters of all five VI layers (module order = theta layout)."""

    def vi_layers(self):
        return [m for m in self if isinstance(m, VIMixin)]

    def check_supported(self):
        mods = list(self)
        kinds = [VIConv2d, nn.ReLU, BatchMaxPool2d, VIConv2d, nn.ReLU, BatchMaxPool2d, nn.Flatten, VILinear, nn.ReLU, VILinear,
                 nn.ReLU, VILinear]
        ok = len(mods) == len(kinds) and all(isinstance(m, k) for m, k in zip(mods, kinds))
        if ok:
            c1, c2, f1, f2, f3 = self.vi_layers()
            ok = (tuple(c1.weight.shape) == (6, 1, 5, 5) and tuple(c1.padding) == (2, 2) and tuple(c2.weight.shape) == (16, 6, 5, 5)
                  and tuple(c2.padding) == (0, 0) and tuple(f1.weight.shape) == (120, 400) and tuple(f2.weight.shape) == (84, 120)
                  and tuple(f3.weight.shape) == (10, 84) and all(m.bias is not None and float(m.prior_sd) == 1.0
                                                                 for m in self.vi_layers()))
        if not ok:
            raise NotImplementedError(f"the convolutional CUDA path covers the lenet stack of make_lenet; got {self}")

    @property
    def dims(self):
        return [784, 10]

    def shared_tail(self):
        """Number of trailing theta entries whose noise is ONE draw shared by all samples: make_lenet builds the last
        VILinear without kwargs, so it keeps mc_samples = 1 (reference :358)."""
        last = self.vi_layers()[-1]
        return last.weight.numel() + last.bias.numel() if int(last.mc_samples) == 1 and self.n_samples() > 1 else 0

    def kl_mask(self):
        """1 for parameters of VILinear layers, 0 for VIConv2d ones: the reference's KL / sampled-nkl sums filter on
        isinstance(VILinear) (psvi_classes.py:479-483,506-510; quirk Q5)."""
        dev = self.vi_layers()[0].weight.device
        return torch.cat([torch.full((m.weight.numel() + m.bias.numel(),), 1.0 if isinstance(m, VILinear) else 0.0, device=dev)
                          for m in self.vi_layers()])

    def forward(self, x, eps=None):
        from psvi.inference.stream import LenetFamily
        _native.require_cuda()
        self.check_supported()
        fam = LenetFamily(self)
        S = max(self.n_samples(), 1)
        dev = fam.mu.device
        if eps is None:
            eps = torch.empty(1, S, fam.Pt, device=dev)
            _native.philox_normal(torch.initial_seed(), _NoiseCounter.next(), 0, 1, S, fam.Pt, eps)
            eps = eps[0]
        eps = fam.fix_eps(eps)
        theta = fam.sample(fam.get_phi(), eps)
        x = x.detach().to(dev, torch.float32).reshape(x.shape[0], -1).contiguous()
        logits = torch.empty(S, x.shape[0], 10, device=dev)
        y = torch.zeros(x.shape[0], device=dev, dtype=torch.int32)
        _native.lenet_pass(S, theta, None, x, y, None, logits=logits)
        off = 0
        for m in self.vi_layers():
            nw, nb = m.weight.numel(), m.bias.numel()
            shared = int(m.mc_samples) == 1 and S > 1
            w = theta[:, off:off + nw].view(S, *m.weight.shape)
            b = theta[:, off + nw:off + nw + nb].view(S, 1, nb)
            m._cached_weight, m._cached_bias = (w[0], b[0, 0]) if (shared or S == 1) else (w, b)
            off += nw + nb
        return logits if S > 1 else logits[0]


def make_lenet(conv_class=None, linear_class=None, pool_class=None, nonl_class=None, **kwargs):
    """reference :334-359 (NB the last linear layer is built WITHOUT kwargs: mc_samples = 1, init_sd = 0.01)."""
    conv_class = VIConv2d if conv_class is None else conv_class
    linear_class = VILinear if linear_class is None else linear_class
    pool_class = BatchMaxPool2d if pool_class is None else pool_class
    nonl_class = nn.ReLU if nonl_class is None else nonl_class
    native = conv_class is VIConv2d and linear_class is VILinear and pool_class is BatchMaxPool2d and nonl_class is nn.ReLU
    return (MeanFieldLeNet if native else nn.Sequential)(
        conv_class(1, 6, 5, padding=2, **kwargs),
        nonl_class(),
        pool_class(2, 2),
        conv_class(6, 16, 5, padding=0, **kwargs),
        nonl_class(),
        pool_class(2, 2),
        nn.Flatten(-3, -1),
        linear_class(400, 120, **kwargs),
        nonl_class(),
        linear_class(120, 84, **kwargs),
        nonl_class(),
        linear_class(84, 10),
    )


def make_alexnet(*args, **kwargs):
    raise NotImplementedError("alexnet is outside the PSVI hot-path scope (SURVEY.md section 2, row 1)")


def gaussian_fn(loc=None, scale=None):
    """reference :18-19"""
    return torch.distributions.normal.Normal(loc, scale)


def make_regressor_net(in_dim, h_dim, out_dim=1, n_layers=2, linear_class=None, nonl_class=None, mc_samples=4, residual=False,
                       **kwargs):
    """reference :300-331 (module names lin{i}, nonl{i}, regressor; default TWO hidden layers): a mean-field stack, evaluated by
    the per-sample network kernels with the Gaussian likelihood (psvi_net_pass_gaussian)."""
    if linear_class is None:
        linear_class = VILinear
    if nonl_class is None:
        nonl_class = nn.ReLU
    net = MeanFieldMLP() if (linear_class is VILinear and nonl_class is nn.ReLU) else nn.Sequential()
    for i in range(n_layers):
        net.add_module(f"lin{i}", linear_class(in_dim if i == 0 else h_dim, h_dim, **kwargs))
        net.add_module(f"nonl{i}", nonl_class())
    net.add_module("regressor", linear_class(h_dim, out_dim, **kwargs))
    for module in net.modules():
        module.mc_samples = mc_samples
    return net


def make_resnet(*args, **kwargs):
    raise NotImplementedError("resnet is outside the PSVI hot-path scope (SURVEY.md section 2, row 1)")
