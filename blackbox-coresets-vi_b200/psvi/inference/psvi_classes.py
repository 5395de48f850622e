"""Black-box PSVI on the B200: same class / method names and argument surface as the reference's
psvi/inference/psvi_classes.py, with the objectives, their gradients, the unrolled inner loop and its reverse-mode
hypergradient executed by hand-written CUDA (libpsvi_b200) instead of autograd + `higher`.

Reference map (psvi/inference/psvi_classes.py):
  PSVI.__init__ :89-227        pseudo_subsample_init :229-285   pseudo_rand_init :287-308
  psvi_elbo :445-486           inner_elbo :488-511              nested_step :541-600      hyper_step :602-687
  set_up_model :689-758        run_psvi :761-1028               evaluate :1031-1108       weight_reset :1110-1128
  PSVILearnV :1344-1360        PSVI_No_Rescaling :1363          PSVIFreeV :1376           PSVIAV :1475-1619
  PSVIFixedU :1622             PSVIAFixedU :1743
Deviations, all deliberate and listed in DESIGN.md: model noise comes from an in-kernel Philox stream (or an injected
exact-noise source for parity tests) instead of the CUDA generator; minibatches are gathered on the device from a
device-resident copy of the dataset; when torch.distributed is initialised the data term and the test set are sharded
over the ranks with a single all-reduce per outer step / per evaluate.
"""
from __future__ import annotations

import time

import math

import numpy as np
import torch
import torch.nn as nn
from torch.utils.data import DataLoader, Dataset
from tqdm import tqdm

from psvi import _native
from psvi.inference.utils import LogResource, compute_empirical_mean
from psvi.models.neural_net import (FullCovMLP, MeanFieldLeNet, MeanFieldMLP, VILinear, VILinearMultivariateNormal, categorical_fn, make_fc2net,
                                    make_fcnet, make_lenet, make_logistic_regression, set_mc_samples)


class SubsetPreservingTransforms(Dataset):
    """Subset of a dataset at given indices (reference :51-80): rows of `.data` for tabular datasets; for MNIST-like image
    datasets (uint8 `.data` [N, 28, 28] + `.transform`) the transformed image, as the reference's image branch does."""

    def __init__(self, dataset, indices=None, dim=2, dnm="Cifar10"):
        self.dataset, self.indices, self.dnm, self.dim = dataset, indices, dnm, dim

    def __getitem__(self, idx):
        if self.dnm in {"MNIST", "FashionMNIST"}:
            from PIL import Image
            import numpy as np
            im = Image.fromarray(np.reshape(self.dataset.data[self.indices[idx]].numpy(), (28, 28)), mode="L")
            return self.dataset.transform(im)
        if self.dnm == "Cifar10":
            raise NotImplementedError("Cifar10 needs the architectures outside the PSVI hot-path scope (SURVEY.md section 2)")
        return self.dataset.data[self.indices[idx]].reshape((self.dim,))

    def __len__(self):
        return len(self.indices)


class ExternalNoise:
    """Exact-noise source for parity tests: hands out pre-drawn standard-normal slabs [n, S, P] in consumption order."""

    def __init__(self, slabs):
        self.slabs, self.pos = slabs, 0

    def take(self, n, device):
        out = self.slabs[self.pos:self.pos + n]
        assert len(out) == n, "external noise exhausted"
        self.pos += n
        return torch.as_tensor(np.stack(out)).to(device=device, dtype=torch.float32).contiguous()


def _adam(params, lr):
    """torch.optim.Adam on the outer variables (reference :860-870) as ONE fused kernel per step() (`fused=True`: same update
    rule; the default foreach form launches seven kernels per optimiser, ~5 % of a cfg2 outer step)."""
    return torch.optim.Adam(params, lr, fused=all(p.is_cuda for p in params))


def _dist_info():
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        return dist, dist.get_rank(), dist.get_world_size()
    return None, 0, 1


def shard_bounds(n, rank, world):
    """Contiguous split of n items over `world` ranks (first ranks get the remainder)."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


class PSVI(object):
    r"""PSVI with fixed rescaled coefficients on pseudodata (reference class PSVI)."""

    _vmode = _native.VMODE_IDENTITY

    def __init__(
        self, u=None, z=None, train_dataset=None, test_dataset=None, N=None, D=None, model=None, optim=None,
        optim_u=None, optim_net=None, optim_v=None, optim_z=None, register_elbos=True, num_pseudo=None, seed=0,
        compute_weights_entropy=True, mc_samples=None, reset=False, reset_interval=10, learn_v=False,
        f=lambda *x: x[0], distr_fn=categorical_fn, dnm="MNIST", nc=10, init_dataset=None, parameterised=False,
        learn_z=False, prune=False, prune_interval=None, prune_sizes=None, increment=False, increment_interval=None,
        increment_sizes=None, lr0alpha=1e-3, retrain_on_coreset=False, device_id=None, data_folder=None,
        results_folder=None, mfvi_selection_method="random", load_from_saved=False, pretrain_epochs=5, lr0net=1e-3,
        multiple_pts_per_cluster=True, loaded_from_psvi=True, alpha_dirichlet=0, choose_difficult=True,
        scoring_run=False, noise_source=None, **kwargs,
    ):
        np.random.seed(seed), torch.manual_seed(seed)
        if not torch.cuda.is_available():
            raise _native.NativeError("PSVI on this package runs on CUDA only (libpsvi_b200, sm_100a): no CUDA device is "
                                      "visible; there is no CPU fallback")
        self.device = torch.device(f"cuda:{device_id}" if device_id else "cuda")
        self.u, self.z = u, z
        self.train_dataset, self.test_dataset = train_dataset, test_dataset
        self.N, self.D, self.dnm, self.nc = N, D, dnm, nc
        self.distr_fn = distr_fn
        self.model, self.optim, self.optim_u, self.optim_net, self.optim_v, self.optim_z = (
            model, optim, optim_u, optim_net, optim_v, optim_z)
        self.register_elbos, self.compute_weights_entropy = register_elbos, compute_weights_entropy
        self.elbos = []
        self.num_pseudo, self.mc_samples = (num_pseudo if not increment else increment_sizes[0]), mc_samples
        self.reset, self.reset_interval, self.learn_v, self.learn_z = reset, reset_interval, learn_v, learn_z
        self.increment_interval, self.increment_sizes = increment_interval, increment_sizes
        for flag, name in ((scoring_run, "scoring_run"),):
            if flag:
                raise NotImplementedError(f"{name}=True is outside the hot path built so far (SURVEY.md section 8f)")
        with torch.no_grad():
            self.v = 1.0 / self.num_pseudo * torch.ones(self.num_pseudo, device=self.device)
        self.v.requires_grad_(self.learn_v)
        self.f, self.parameterised = f, parameterised
        self.init_dataset = init_dataset
        self.results = {}
        self.prune, self.increment, self.retrain_on_coreset = prune, increment, retrain_on_coreset
        self.prune_interval, self.prune_sizes = prune_interval, prune_sizes
        self.lr0alpha, self.lr0net = lr0alpha, lr0net
        self.data_folder, self.results_folder = data_folder, results_folder
        self.chosen_indices = []
        self.seed = seed
        self.alpha = None
        self.noise_source = noise_source
        self._noise_domain = 0
        self._ws = {}
        self._dev_data = {}

    # ------------------------------------------------------------------------------------------------ noise
    def _noise(self, n_slabs):
        if self.noise_source is not None:
            return _native.make_noise(self.noise_source.take(n_slabs, self.device))
        self._noise_domain += 1
        return _native.make_noise(None, seed=self.seed, domain=self._noise_domain)

    # ------------------------------------------------------------------------------------------------ pseudo-data init
    def pseudo_subsample_init(self):
        """Class-balanced random subset of the training data (reference :229-285, same loader-driven RNG use)."""
        chosen = self.train_dataset if self.init_dataset is None else self.init_dataset
        ppc = [self.num_pseudo // self.nc] * self.nc
        ppc[-1] = self.num_pseudo - sum(ppc[:-1])
        with torch.no_grad():
            self.z = torch.tensor([c for c, k in enumerate(ppc) for _ in range(k)]).float().to(self.device)
        lst = []
        for c in range(self.nc):
            idx = (torch.as_tensor(chosen.targets).clone().detach() == c).nonzero()
            loader = DataLoader(SubsetPreservingTransforms(chosen, indices=idx, dnm=self.dnm, dim=self.D),
                                batch_size=ppc[c], shuffle=True)
            lst.append(next(iter(loader)).to(device=self.device))
        self.u = torch.cat(lst).float().requires_grad_(True)
        if self.learn_z:
            # target logits initialised at the one-hot encoding of the class labels (reference :259-264)
            self.z = torch.nn.functional.one_hot(self.z.to(torch.int64), num_classes=self.nc).float().requires_grad_(True)

    def pseudo_rand_init(self, variance=1.0):
        """Noisy empirical mean + labels split equally among classes (reference :287-308)."""
        self.u = ((compute_empirical_mean(self.train_loader) + variance * torch.randn(self.num_pseudo, self.D))
                  .clone()).to(self.device).float().requires_grad_(True)
        z = [c * torch.ones(self.num_pseudo // self.nc if c < self.nc - 1
                            else self.num_pseudo - (self.nc - 1) * (self.num_pseudo // self.nc)) for c in range(self.nc)]
        self.z = torch.cat(z).to(self.device)

    # ------------------------------------------------------------------------------------------------ native plumbing
    def _model_desc(self, model=None):
        model = self.model if model is None else model
        if not isinstance(model, (MeanFieldMLP, FullCovMLP)):
            raise NotImplementedError("the CUDA path covers mean-field MLPs (logistic_regression, fn) and fn2; got "
                                      f"{type(model).__name__}")
        model.check_supported()
        S = model.n_samples()
        if isinstance(model, MeanFieldLeNet):
            return model, None, S          # evaluated by the streaming path only (no psvi_mf_model descriptor)
        return model, _native.make_model(model.dims, S), S

    # ---- engine choice: fused cluster kernel when the model fits its shared-memory budget, streaming path otherwise ----
    # arithmetic of the bilevel step in the large regime, set before the first step:
    #   "mixed"  (default) gradient / outer passes in tf32x3, Hessian-vector passes of the reverse sweep in split-bf16 pairs:
    #            hypergradients at the tf32x3 level (the rounding sensitivity sits in the gradient passes), ~20 % faster step
    #   "tf32x3" every pass in tf32x3 (fp32-class)
    #   "bf16x3" every pass in split-bf16 pairs: ~1.5x faster than tf32x3, hypergradient cosine >= 0.999 instead of >= 0.9999
    large_precision = "mixed"

    def _stream(self, model):
        """StreamEngine for `model` (fn2, or a mean-field MLP the fused engine reported as PSVI_ERR_UNSUPPORTED)."""
        from psvi.inference.stream import FullCovFamily, LenetFamily, LenetNet, MeanFieldFamily, StreamEngine
        key = (id(model), model.n_samples())    # PSVI_No_IW switches mc_samples between training and evaluation
        eng = self._ws.get(("stream", key))
        if eng is None:
            if isinstance(model, MeanFieldLeNet):
                eng = StreamEngine(LenetFamily(model), model.dims, model.n_samples(), net=LenetNet(model.n_samples()))
            elif isinstance(model, MeanFieldMLP) and self._is_large_fn(model):
                # large regime (BASELINE config 5): batched TMA + tcgen05 GEMMs, bf16 operands
                from psvi.inference.stream import FnLargeNet
                prec, prec_dual = {"tf32x3": (_native.PREC_TF32X3, _native.PREC_TF32X3),
                                   "bf16x3": (_native.PREC_BF16X3, _native.PREC_BF16X3),
                                   "mixed": (_native.PREC_TF32X3, _native.PREC_BF16X3)}[self.large_precision]
                eng = StreamEngine(MeanFieldFamily(model), model.dims, model.n_samples(),
                                   net=FnLargeNet(model.dims, model.n_samples(), precision=prec, precision_dual=prec_dual))
            else:
                fam = FullCovFamily(model) if isinstance(model, FullCovMLP) else MeanFieldFamily(model)
                eng = StreamEngine(fam, model.dims, model.n_samples())
            self._ws[("stream", key)] = eng
        if isinstance(model, MeanFieldMLP):
            eng.fam.mu, eng.fam.rho = model.flat()
        return eng

    @staticmethod
    def _is_large_fn(model):
        """One-hidden-layer fn whose per-sample weights (> 40 k floats) fit no CTA and whose shape suits the tensor path."""
        from psvi.inference.stream import FnLargeNet
        dims = model.dims
        return (FnLargeNet.fits(dims, model.n_samples()) and dims[1] * (dims[0] + 1) + dims[2] * (dims[1] + 1) > 40000)

    _outer_kind = "psvi"      # which outer objective nested_step differentiates ("ablated": PSVI_Ablated / PSVI_No_IW)
    fulldata_min_rows = 8192  # large fn: data terms with at least this many rows (over all ranks) take the bf16 tensor path

    def _inner_pseudo(self, u, z32, a):
        """(u, z, a) as the inner objective sees them (hook for the mc_samples == 1 quirk of PSVI_No_IW)."""
        return u, z32, a

    def _collapse_pseudo(self, ubar, abar):
        return ubar, abar

    def _use_stream(self, model):
        return (isinstance(model, (FullCovMLP, MeanFieldLeNet)) or self.learn_z
                or self._ws.get(("force_stream", id(model)), False))

    # ---- learn_z: soft pseudo-labels (reference :455-474,499-504, the KLDivLoss branch) ---------------------------------
    # A soft-label row r with targets t[r, :] contributes  sum_c t[r,c] (log t[r,c] - log p_s[r,c])  =  kappa_r + sum_c t[r,c]
    # * nll_s(x_r, label c): the kernels see it as C hard-label rows (x_r, c) with row weights t[r,c]; kappa_r = sum_c t log t
    # does not depend on the network.  t = labels.softmax(0): normalised over the ROWS of the label matrix (per class column),
    # exactly as the reference does -- the inner objective uses softmax(z, 0), the outer one the joint softmax over
    # cat(z, nc * one_hot(y)).  Gradients reach z through torch autograd on those two (tiny) softmaxes.
    def _soft_rows(self, x, t):
        """[R, D] rows and [R, C] targets -> (R C rows, labels 0..C-1 repeated, flattened targets)."""
        R, C = t.shape
        return (x.repeat_interleave(C, 0).contiguous(), torch.arange(C, device=x.device, dtype=torch.int32).repeat(R),
                t.reshape(-1).float().contiguous())

    def _soft_targets(self, yb=None):
        z = self.z
        t_in = torch.softmax(z, 0)
        if yb is None:
            return t_in, None, None
        L = torch.cat([z, self.nc * torch.nn.functional.one_hot(yb.to(torch.int64), num_classes=self.nc).to(z.dtype)])
        t_all = torch.softmax(L, 0)
        return t_in, t_all[:z.shape[0]], t_all[z.shape[0]:]

    def _require_single_rank_learn_z(self):
        if _dist_info()[2] > 1:
            raise NotImplementedError("learn_z couples all rows of a minibatch through softmax(0): not sharded over ranks")

    def _inner_elbo_learn_z(self, model):
        eng, S = self._stream(model), model.n_samples()
        u, _ = self._uv()
        a = self._a()
        t_in = self._soft_targets()[0].detach().float()
        ue, le, te = self._soft_rows(u, t_in)
        val, g = eng.inner_grad(eng.fam.get_phi(), self._noise_tensor(1, eng.Pt, S)[0], ue, le, (a[:, None] * t_in).reshape(-1))
        self._last_inner = g
        return (val + S * (a.double() * torch.xlogy(t_in, t_in).sum(1).double()).sum()).float()

    def _psvi_elbo_learn_z(self, model, xb, yb):
        eng, S = self._stream(model), model.n_samples()
        u, _ = self._uv()
        a, N, B = self._a(), float(self.N), xb.shape[0]
        _, t_p, t_d = (t.detach().float() for t in self._soft_targets(yb))
        ue, le, _ = self._soft_rows(u, t_p)
        xe, ye, dw = self._soft_rows(xb, t_d)
        ex = {}
        loss, pbar, ubar, abar, _ = eng.outer_grad(eng.fam.get_phi(), self._noise_tensor(1, eng.Pt, S)[0], ue, le,
                                                   (a[:, None] * t_p).reshape(-1), xe, ye, N, n_total=B, data_w=dw, extras=ex)
        M, C = t_p.shape
        vg, ag = self._v_grad_from_abar((t_p * abar.reshape(M, C)).sum(1))
        self._last_outer = dict(phi_grad=pbar, u_grad=ubar.reshape(M, C, -1).sum(1), v_grad=vg, alpha_grad=ag)
        return loss + (N / B) * torch.xlogy(t_d, t_d).sum()

    def _nested_step_learn_z(self, model, S, xbatch, ybatch):
        """nested_step with soft pseudo-labels (reference :541-600 with :455-474,499-504): hypergradients on u, v AND z."""
        self._require_single_rank_learn_z()
        eng = self._stream(model)
        T, lr, N = int(self.inner_it), float(self.optim_net.param_groups[0]["lr"]), float(self.N)
        u, _ = self._uv()
        xb = xbatch.detach().to(self.device, torch.float32).reshape(xbatch.shape[0], -1).contiguous()
        yb = ybatch.detach().to(self.device)
        a, B = self._a(), xb.shape[0]
        t_in, t_p, t_d = self._soft_targets(yb)                      # autograd graph z -> targets
        tin, tp, td = t_in.detach().float(), t_p.detach().float(), t_d.detach().float()
        M, C = tin.shape
        ue, le, _ = self._soft_rows(u, tin)
        xe, ye, dw = self._soft_rows(xb, td)
        ex = {}
        loss, ubar, abar, phi_T, il = eng.nested_cached(eng.fam.get_phi(), self._noise_tensor(T + 1, eng.Pt, S), ue, le,
                                                 (a[:, None] * tin).reshape(-1), xe, ye, N, T, lr,
                                                 want_losses=self.register_elbos, n_total=B,
                                                 a_outer=(a[:, None] * tp).reshape(-1), data_w=dw, extras=ex)
        eng.fam.set_phi(phi_T)
        ab_out = ex["abar_outer"].reshape(M, C)
        ab_in = abar.reshape(M, C) - ab_out
        kd = (N / B) * torch.xlogy(t_d, t_d).sum()                   # the data rows' sum t log t (times sum_s w_s = 1)
        loss = loss + kd.detach().float()
        if self.register_elbos:
            kin = S * (a.double() * torch.xlogy(tin, tin).sum(1).double()).sum()
            ilc = (il.double() + kin).cpu()
            for in_it in range(0, T, max(int(self.log_every), 1)):
                self.elbos.append((1, -ilc[in_it].item()))
            self.elbos.append((0, -loss.item()))
        # dLoss/dz: chain dLoss/dt (pseudo rows: a_m * d/d(a_m t_mc); data rows: dwbar) through the two softmaxes over rows
        surrogate = ((t_in * (a[:, None] * ab_in).to(t_in.dtype)).sum() + (t_p * (a[:, None] * ab_out).to(t_p.dtype)).sum()
                     + (t_d * ex["dwbar"].reshape(B, C).to(t_d.dtype)).sum() + kd)
        self.z.grad = torch.autograd.grad(surrogate, self.z)[0]
        vg, ag = self._v_grad_from_abar((tin * ab_in).sum(1) + (tp * ab_out).sum(1))
        self.u.grad = ubar.reshape(M, C, -1).sum(1).to(self.u.dtype).reshape(self.u.shape)
        if self.learn_v:
            self.v.grad = vg.to(self.v.dtype)
        if self.alpha is not None and self.alpha.requires_grad and ag is not None:
            self.alpha.grad = ag.to(self.alpha.dtype)
        self._step_outer_optimisers()
        if self.scheduler_optim_net:
            self.scheduler_optim_net.step()
        if self.optim_z is not None:
            self.optim_z.step()
        return loss

    def _evaluate_learn_z(self, model, S, correction):
        """evaluate() with soft labels (reference :1049-1056): the pseudo term is summed over classes AND samples before it
        meets N f(v), so it shifts every log-weight equally -- the importance weights are softmax(sampled_nkl)."""
        eng = self._stream(model)
        xt, yt = self._device_dataset(self._test_ds(), "test")
        batch = int(self.data_minibatch)
        n_slabs = -(-xt.shape[0] // batch)
        out = eng.evaluate(eng.fam.get_phi(), self._noise_tensor(n_slabs, eng.Pt, S), None, None, None, xt, yt, batch,
                           mode=0 if correction else 1)
        vs = self.f(self.v.detach(), 0)
        v_entropy = vs.sum().square() / vs.square().sum() / self.num_pseudo if self.compute_weights_entropy else None
        return (out[1] / out[2], out[0] / out[2], out[3] if self.compute_weights_entropy else None, out[4], v_entropy)

    def _fused(self, model, fn):
        """Run fn() on the fused engine; if the model does not fit it, remember that and return None."""
        pos, dom = getattr(self.noise_source, "pos", None), self._noise_domain
        try:
            return fn()
        except _native.NativeError as e:
            if e.code != _native.ERR_UNSUPPORTED:
                raise
            self._ws[("force_stream", id(model))] = True
            if pos is not None:
                self.noise_source.pos = pos      # the refused call consumed no noise
            self._noise_domain = dom
            return None

    def _fits_fused(self, model, desc):
        """Does the model fit the shared-memory-resident cluster engines?  Probed once per model with a throw-away noise
        stream (nothing of this object's state or noise order is touched), so that callers can pick the streaming path
        BEFORE they mutate anything."""
        key = ("fits_fused", id(model), model.n_samples())
        fit = self._ws.get(key)
        if fit is None:
            mu, rho = model.flat()
            u, v = self._uv()
            g, val = torch.empty(2 * mu.numel(), device=self.device), torch.empty(1, device=self.device)
            try:
                _native.inner_grad(desc, _native.make_noise(None, seed=0, domain=0), mu, rho, u, self._z32(), v, float(self.N),
                                   self._vmode, self._alpha_value(), g, val)
                fit = True
            except _native.NativeError as e:
                if e.code != _native.ERR_UNSUPPORTED:
                    raise
                fit = False
                self._ws[("force_stream", id(model))] = True
            self._ws[key] = fit
        return fit

    def _noise_tensor(self, n_slabs, Pt, S):
        if self.noise_source is not None:
            return self.noise_source.take(n_slabs, self.device)
        self._noise_domain += 1
        eps = torch.empty(n_slabs, S, Pt, device=self.device)
        _native.philox_normal(self.seed, self._noise_domain, 0, n_slabs, S, Pt, eps)
        return eps

    def _a(self):
        return _native.coreset_weights(self.v.detach().float(), float(self.N), self._vmode, self._alpha_value())

    def _v_grad_from_abar(self, abar):
        """dLoss/dv (and dalpha) from dLoss/da through a = N f(v)."""
        N = float(self.N)
        if self._vmode == _native.VMODE_IDENTITY:
            return N * abar, None
        f = torch.softmax(self.v.detach().float(), 0)
        sc = N * (float(torch.exp(self.alpha.detach())) if self._vmode == _native.VMODE_EXPALPHA_SOFTMAX else 1.0)
        dot = (f * abar).sum()
        return sc * f * (abar - dot), (sc * dot).reshape(1)

    def _buf(self, name, n):
        t = self._ws.get(name)
        if t is None or t.numel() < n or t.device != self.device:
            t = torch.zeros(max(int(n), 1), device=self.device, dtype=torch.float32)
            self._ws[name] = t
        return t

    def _z32(self):
        """Labels of the pseudo-data as int32 (what the kernels read); converted once per label tensor, not per step."""
        c = self._ws.get("z32")
        if c is None or c[0] is not self.z or c[1] != self.z._version:
            c = (self.z, self.z._version, self.z.detach().to(torch.int32).contiguous())
            self._ws["z32"] = c
        return c[2]

    def _alpha_value(self):
        return float(self.alpha.item()) if self.alpha is not None else 0.0

    def _uv(self):
        # image pseudo-data [M, 1, 28, 28] (lenet) is handed to the kernels as rows [M, 784]
        return (self.u.detach().float().reshape(self.u.shape[0], -1).contiguous(), self.v.detach().float().contiguous())

    # ------------------------------------------------------------------------------------------------ objectives
    def psvi_elbo(self, xbatch, ybatch, model=None, params=None, hyperopt=False):
        """Negative PSVI-ELBO (reference :445-486).  Returns a 0-dim tensor; its gradients wrt the variational
        parameters, u and v are left in `self._last_outer` (the fused kernel produces them in the same pass)."""
        assert self.mc_samples > 1
        model, desc, S = self._model_desc(model)
        xb = xbatch.detach().to(self.device, torch.float32).reshape(xbatch.shape[0], -1).contiguous()
        yb = ybatch.detach().to(self.device).to(torch.int32).contiguous()
        if self.learn_z:
            return self._psvi_elbo_learn_z(model, xb, yb)
        if not self._use_stream(model):
            out = self._fused(model, lambda: self._psvi_elbo_fused(model, desc, xb, yb))
            if out is not None:
                return out
        eng = self._stream(model)
        u, _ = self._uv()
        phi = eng.fam.get_phi()
        loss, pbar, ubar, abar, _ = eng.outer_grad(phi, self._noise_tensor(1, eng.Pt, S)[0], u, self._z32(), self._a(), xb,
                                                   yb, float(self.N))
        vg, ag = self._v_grad_from_abar(abar)
        self._last_outer = dict(phi_grad=pbar, u_grad=ubar, v_grad=vg, alpha_grad=ag)
        return loss

    def _psvi_elbo_fused(self, model, desc, xb, yb):
        mu, rho = model.flat()
        u, v = self._uv()
        M, D = u.shape
        P = mu.numel()
        gout = self._buf("gout", _native.gout_floats(desc, M))
        ug, vg, ag, loss = torch.zeros(M, D, device=self.device), torch.zeros(M, device=self.device), \
            torch.zeros(1, device=self.device), torch.zeros(1, device=self.device)
        _native.outer_grad(desc, self._noise(1), mu, rho, u, self._z32(), v, xb, yb, xb.shape[0], float(self.N),
                           self._vmode, self._alpha_value(), 1.0, gout, ug, vg, ag, loss)
        self._last_outer = dict(phi_grad=gout[:2 * P].clone(), u_grad=ug, v_grad=vg, alpha_grad=ag)
        return loss[0]

    def inner_elbo(self, model=None, params=None, hyperopt=False):
        """Negative ELBO on the pseudo-data (reference :488-511).  Gradient wrt (mu, rho) in `self._last_inner`."""
        model, desc, S = self._model_desc(model)
        if self.learn_z:
            return self._inner_elbo_learn_z(model)
        if not self._use_stream(model):
            out = self._fused(model, lambda: self._inner_elbo_fused(model, desc))
            if out is not None:
                return out
        eng = self._stream(model)
        u, _ = self._uv()
        val, g = eng.inner_grad(eng.fam.get_phi(), self._noise_tensor(1, eng.Pt, S)[0], u, self._z32(), self._a())
        self._last_inner = g
        return val.float()

    def _inner_elbo_fused(self, model, desc):
        mu, rho = model.flat()
        u, v = self._uv()
        grad, val = torch.zeros(2 * mu.numel(), device=self.device), torch.zeros(1, device=self.device)
        _native.inner_grad(desc, self._noise(1), mu, rho, u, self._z32(), v, float(self.N), self._vmode,
                           self._alpha_value(), grad, val)
        self._last_inner = grad
        return val[0]

    # ------------------------------------------------------------------------------------------------ optimisation
    def _zero_grads(self):
        for o in (self.optim_u, self.optim_net, self.optim_v if self.learn_v else None, getattr(self, "optim_alpha", None),
                  self.optim_z if self.learn_z else None):
            if o is not None:
                o.zero_grad()

    def nested_step(self, xbatch, ybatch, truncated=False, K=5):
        """One bilevel step (reference :541-600): T differentiable Adam steps on inner_elbo, psvi_elbo, hypergradient on
        (u, v[, alpha]) -- one fused CUDA launch (two around an all-reduce when sharded) -- then the Adam steps on u, v."""
        if truncated:
            raise NotImplementedError("truncated=True is never taken by run_psvi (SURVEY.md section 8a, a8)")
        self._zero_grads()
        model, desc, S = self._model_desc()
        if self.learn_z:
            return self._nested_step_learn_z(model, S, xbatch, ybatch)
        if not self._use_stream(model):
            out = self._fused(model, lambda: self._nested_step_fused(model, desc, S, xbatch, ybatch))
            if out is not None:
                return out
        return self._nested_step_stream(model, S, xbatch, ybatch)

    def _nested_step_stream(self, model, S, xbatch, ybatch):
        """Same step through the streaming path (fn2 / lenet / medium and large mean-field models): per-sample network kernels
        sequenced on the host (psvi/inference/stream.py).  Under torch.distributed the data term is sharded over the ranks."""
        dist, rank, world = _dist_info()
        eng = self._stream(model)
        T, lr = int(self.inner_it), float(self.optim_net.param_groups[0]["lr"])
        u, _ = self._uv()
        xb = xbatch.detach().to(self.device, torch.float32).reshape(xbatch.shape[0], -1).contiguous()
        yb = ybatch.detach().to(self.device).to(torch.int32).contiguous()
        kappa, n_total, reduce_fn = 1.0, None, None
        if world > 1:
            # the data term shards over ranks (contiguous row ranges of the SAME global minibatch); one all-reduce of
            # [loss | dL/dphi_T | direct dL/du | direct dL/da]; inner loop and reverse sweep replicated (same seeds)
            n_total = xb.shape[0]
            lo, hi = shard_bounds(n_total, rank, world)
            xb, yb, kappa = xb[lo:hi].contiguous(), yb[lo:hi].contiguous(), 1.0 / world

            def reduce_fn(loss, pbar, ubar, abar):
                flat = torch.cat([loss.reshape(1), pbar.reshape(-1), ubar.reshape(-1), abar.reshape(-1)]).float()
                dist.all_reduce(flat)
                n1, n2 = pbar.numel(), ubar.numel()
                return flat[0], flat[1:1 + n1], flat[1 + n1:1 + n1 + n2].reshape(ubar.shape), flat[1 + n1 + n2:]
        u_in, z_in, a_in = self._inner_pseudo(u, self._z32(), self._a())
        # many data rows on a large fn (the full-data / sharded data term of BASELINE configs[4]): fused tensor-core pass over
        # bf16 rows (psvi_fn_data_grad_tc); below the threshold the exact tf32x3 pass of the minibatch stays
        xb16 = None
        if (self._outer_kind == "psvi" and isinstance(model, MeanFieldMLP) and not isinstance(model, MeanFieldLeNet)
                and self._is_large_fn(model) and self._fits_fn_tc(model)
                and (n_total if n_total is not None else xb.shape[0]) >= self.fulldata_min_rows):
            xb16 = torch.empty(xb.shape, device=xb.device, dtype=torch.bfloat16)
            if xb.numel():
                _native.f32_to_bf16(xb, xb16)
        loss, ubar, abar, phi_T, il = eng.nested_cached(eng.fam.get_phi(), self._noise_tensor(T + 1, eng.Pt, S), u_in, z_in, a_in,
                                                 xb, yb, float(self.N), T, lr, want_losses=self.register_elbos,
                                                 kappa=kappa, n_total=n_total, reduce_fn=reduce_fn, outer=self._outer_kind,
                                                 xb_bf16=xb16)
        ubar, abar = self._collapse_pseudo(ubar, abar)
        eng.fam.set_phi(phi_T)                 # copy-back of the fast weights (reference :596-599)
        if self.register_elbos:
            ilc = il.cpu()
            for in_it in range(0, T, max(int(self.log_every), 1)):
                self.elbos.append((1, -ilc[in_it].item()))
            self.elbos.append((0, -loss.item()))
        vg, ag = self._v_grad_from_abar(abar)
        self.u.grad = ubar.to(self.u.dtype).reshape(self.u.shape)
        if self.learn_v:
            self.v.grad = vg.to(self.v.dtype)
        if self.alpha is not None and self.alpha.requires_grad and ag is not None:
            self.alpha.grad = ag.to(self.alpha.dtype)
        self._step_outer_optimisers()
        if self.scheduler_optim_net:
            self.scheduler_optim_net.step()
        return loss

    def _nested_step_fused(self, model, desc, S, xbatch, ybatch):
        mu, rho = model.flat()
        u, v = self._uv()
        M, D = u.shape
        P, T = mu.numel(), int(self.inner_it)
        lr = float(self.optim_net.param_groups[0]["lr"])
        traj = self._buf("traj", _native.traj_floats(desc, T))
        gout = self._buf("gout", _native.gout_floats(desc, M))
        # outputs the kernel overwrites completely: fresh (uninitialised) storage per step, because u.grad / v.grad keep
        # referring to it; one allocation instead of four zero-fills
        obuf = torch.empty(M * D + M + 2, device=self.device)
        ug, vg = obuf[:M * D].view(M, D), obuf[M * D:M * D + M]
        ag, loss = obuf[M * D + M:M * D + M + 1], obuf[M * D + M + 1:]
        if self._vmode != _native.VMODE_EXPALPHA_SOFTMAX:
            ag = None
        il = self._buf("il", T) if self.register_elbos else None
        xb = xbatch.detach().to(self.device, torch.float32)
        yb = ybatch.detach().to(self.device)
        if yb.dtype != torch.int32:
            yb = yb.to(torch.int32)
        dist, rank, world = _dist_info()
        n_total = xb.shape[0]
        noise = self._noise(T + 1)
        z32 = self._z32()
        args = lambda xs, ys, kappa: (desc, noise, mu, rho, u, z32, v, xs, ys, n_total, float(self.N),  # noqa: E731
                                      self._vmode, self._alpha_value(), T, lr, kappa)
        if world == 1:
            _native.nested_step(*args(xb.contiguous(), yb.contiguous(), 1.0),
                                _native.PHASE_UNROLL | _native.PHASE_REVERSE, traj, None, ug, vg, ag, loss, il)
        else:
            lo, hi = shard_bounds(n_total, rank, world)
            xs, ys = xb[lo:hi].contiguous(), yb[lo:hi].contiguous()
            a = args(xs if hi > lo else None, ys if hi > lo else None, 1.0 / world)
            _native.nested_step(*a, _native.PHASE_UNROLL, traj, gout, ug, vg, ag, loss, il)
            n_red = 2 * P + M * D + M + S + 4      # [dL/dphi_T | direct du | direct da | d_s | loss terms]
            dist.all_reduce(gout[:n_red])           # the ONE collective of the step (NCCL over NVLink)
            loss = gout[2 * P + M * D + M + S:2 * P + M * D + M + S + 1].clone()
            _native.nested_step(*a, _native.PHASE_REVERSE, traj, gout, ug, vg, ag, None, None)
        if self.register_elbos:
            ilc = il[:T].cpu()
            for in_it in range(0, T, max(int(self.log_every), 1)):
                self.elbos.append((1, -ilc[in_it].item()))
            self.elbos.append((0, -loss.item()))
        self.u.grad = ug.to(self.u.dtype)
        if self.learn_v:
            self.v.grad = vg.to(self.v.dtype)
        if self.alpha is not None and self.alpha.requires_grad and ag is not None:
            self.alpha.grad = ag.to(self.alpha.dtype)
        self._step_outer_optimisers()
        if self.scheduler_optim_net:
            self.scheduler_optim_net.step()
        # the fast weights phi_T were written back into the model's parameters by the kernel (reference :596-599)
        return loss[0] if loss.dim() else loss

    def _step_outer_optimisers(self):
        self.optim_u.step()
        if self.learn_v:
            self.optim_v.step()
            if not self.parameterised:
                with torch.no_grad():
                    torch.clamp_(self.v, min=0.0)
        if getattr(self, "optim_alpha", None) is not None and self.learn_v:
            self.optim_alpha.step()

    def hyper_step(self, xbatch, ybatch, T=50, inner_opt_class=None, K=30, linsys_lr=1e-4,
                   hypergrad_approx="CG_normaleq", **kwargs):
        """Implicit-differentiation step (reference :602-687 with hypergrad.CG_normaleq :199-244 / fixed_point :83-140):
        T plain Adam steps (hypergrad.DifferentiableAdam arithmetic), then K iterations of CG on the normal equations of
        the fixed-point map  Phi(w) = w - linsys_lr * grad inner(w)  whose Jacobian products are Hessian-vector products
        of the inner objective -- each one fused CUDA pass.  Noise is consumed in the reference's order (the JVP's
        double-VJP evaluates Phi twice, the first draw is discarded)."""
        from psvi.hypergrad.hypergradients import cg_normaleq_native, fixed_point_native
        if self.learn_z:
            raise NotImplementedError("--trainer hyper with learn_z (soft pseudo-labels) is not built; use --trainer nested")
        if self._outer_kind != "psvi":
            # PSVI_Ablated / PSVI_No_IW: the reference's outer_loss_function would call the ablated psvi_elbo (and, for
            # No_IW, the mc_samples == 1 label-broadcast quirk of inner_elbo); the implicit solvers here differentiate the
            # importance-weighted objective only -- refuse instead of returning a different hypergradient
            raise NotImplementedError(f"--trainer hyper is not built for {type(self).__name__} (ablated outer objective); "
                                      "use --trainer nested")
        T = int(self.inner_it)
        self._zero_grads()
        model, desc, S = self._model_desc()
        if (self._use_stream(model) or (isinstance(model, MeanFieldMLP) and self._is_large_fn(model))
                or not self._fits_fused(model, desc)):
            return self._hyper_step_stream(model, S, xbatch, ybatch, K, linsys_lr, hypergrad_approx)
        mu, rho = model.flat()
        P = mu.numel()
        u, v = self._uv()
        z32 = self._z32()
        lr = float(self.optim_net.param_groups[0]["lr"])
        am, av = torch.zeros(2 * P, device=self.device), torch.zeros(2 * P, device=self.device)
        _native.unroll(desc, self._noise(T), mu, rho, am, av, 0, u, z32, None, v, float(self.N), self._vmode,
                       self._alpha_value(), T, lr, _native.ADAM_HYPERGRAD, None)
        solver = cg_normaleq_native if hypergrad_approx == "CG_normaleq" else fixed_point_native
        ug, vg, ag = solver(self, desc, mu, rho, u, z32, v, xbatch, ybatch, K, linsys_lr)
        self.u.grad = ug if self.u.grad is None else self.u.grad + ug
        if self.learn_v:
            self.v.grad = vg if self.v.grad is None else self.v.grad + vg
        if self.alpha is not None and self.alpha.requires_grad:
            self.alpha.grad = ag
        self._step_outer_optimisers()
        ll = self.psvi_elbo(xbatch, ybatch, model=self.model)
        return ll.item()

    def _hyper_step_stream(self, model, S, xbatch, ybatch, K, linsys_lr, hypergrad_approx):
        """hyper_step through the streaming engine: T plain steps with the hypergrad.DifferentiableAdam arithmetic
        (psvi/hypergrad/diff_optimizers.py:184-213: u += 1e-12, sqrt(u / (1 - beta2^t)) + eps, no mask), then the implicit
        hypergradient solver over StreamEngine.hvp / outer_grad."""
        from psvi.hypergrad.hypergradients import hyper_stream
        eng = self._stream(model)
        T, lr = int(self.inner_it), float(self.optim_net.param_groups[0]["lr"])
        u, _ = self._uv()
        z32, a = self._z32(), self._a()
        xb = xbatch.detach().to(self.device, torch.float32).reshape(xbatch.shape[0], -1).contiguous()
        yb = ybatch.detach().to(self.device).to(torch.int32).contiguous()
        phi = eng.fam.get_phi()
        m, uu = torch.zeros_like(phi), torch.zeros_like(phi)
        eps_all = self._noise_tensor(T, eng.Pt, S)
        for t in range(T):
            _, g = eng.inner_grad(phi, eps_all[t], u, z32, a)
            m = 0.9 * m + (1.0 - 0.9) * g
            uu = 0.999 * uu + (1.0 - 0.999) * g * g + 1e-12
            phi = phi - lr * (m / (1.0 - 0.9 ** (t + 1)) / (torch.sqrt(uu / (1.0 - 0.999 ** (t + 1))) + 1e-8))
        ug, vg, ag = hyper_stream(self, eng, S, phi, u, z32, a, xb, yb, K, linsys_lr, hypergrad_approx)
        eng.fam.set_phi(phi)
        ug = ug.to(self.u.dtype).reshape(self.u.shape)
        self.u.grad = ug if self.u.grad is None else self.u.grad + ug
        if self.learn_v:
            self.v.grad = vg if self.v.grad is None else self.v.grad + vg
        if self.alpha is not None and self.alpha.requires_grad:
            self.alpha.grad = ag
        self._step_outer_optimisers()
        ll = self.psvi_elbo(xbatch, ybatch, model=self.model)
        return ll.item()

    # --trainer joint / alternating (reference :517-539): first-order steps on psvi_elbo.  The fused outer pass returns the
    # value AND the gradients wrt (mu, rho), u and v in one launch (self._last_outer), which is what loss.backward() gave
    # the reference; the Adam updates are torch.optim.Adam on a flat leaf that mirrors the model's variational parameters
    # (Adam is elementwise, so one flat tensor == the reference's per-layer parameter list).
    def _phi_get(self):
        model = self.model
        if isinstance(model, MeanFieldMLP):
            return torch.cat(model.flat())
        return self._stream(model).fam.get_phi()

    def _phi_set(self, phi):
        model = self.model
        if isinstance(model, MeanFieldMLP):
            mu, rho = model.flat()
            mu.copy_(phi[:mu.numel()])
            rho.copy_(phi[mu.numel():])
        else:
            self._stream(model).fam.set_phi(phi)

    def _phi_leaf(self):
        phi = self._phi_get()
        leaf = getattr(self, "_phi_param", None)
        if leaf is None or leaf.shape != phi.shape or leaf.device != phi.device:
            leaf = self._phi_param = torch.nn.Parameter(phi.clone())
        else:
            leaf.data.copy_(phi)        # (weight_reset / increment_coreset may have touched the model in between)
        return leaf

    def _first_order_step(self, xbatch, ybatch, optim, tag, net, pseudo):
        """optim.zero_grad(); loss = psvi_elbo(...); loss.backward(); optim.step() of the reference for the parameter groups
        `net` (model) / `pseudo` (u and, if learnt, v) that `optim` holds."""
        leaf = self._phi_leaf()
        optim.zero_grad()
        loss = self.psvi_elbo(xbatch, ybatch, model=self.model)
        if self.register_elbos:
            self.elbos.append((tag, -loss.item()))
        g = self._last_outer
        if net:
            leaf.grad = g["phi_grad"].to(leaf.dtype).reshape(leaf.shape)
        if pseudo:
            self.u.grad = g["u_grad"].to(self.u.dtype).reshape(self.u.shape)
            if self.learn_v and pseudo == "uv":
                self.v.grad = g["v_grad"].to(self.v.dtype)
        optim.step()
        if net:
            with torch.no_grad():
                self._phi_set(leaf.data)
        return loss

    def joint_step(self, xbatch, ybatch):
        """reference :517-526: one Adam(lr0joint) over model parameters, u and (if learnt) v."""
        if getattr(self, "_optim_joint_for", None) is not self.model:
            params = [self._phi_leaf(), self.u] + ([self.v] if self.learn_v else [])
            self.optim = torch.optim.Adam(params, self._lr0joint)
            self._optim_joint_for = self.model
        return self._first_order_step(xbatch, ybatch, self.optim, 2, net=True, pseudo="uv")

    def alternating_step(self, xbatch, ybatch):
        """reference :528-539: an Adam step of the model (lr of optim_net) on one draw, then an Adam step of u (optim_u) on
        a fresh draw at the updated model; v is not stepped by this trainer."""
        if getattr(self, "_optim_alt_for", None) is not self.model:
            self._optim_alt_net = torch.optim.Adam([self._phi_leaf()], float(self.optim_net.param_groups[0]["lr"]))
            self._optim_alt_for = self.model
        self._first_order_step(xbatch, ybatch, self._optim_alt_net, 0, net=True, pseudo=None)
        return self._first_order_step(xbatch, ybatch, self.optim_u, 1, net=False, pseudo="u")

    # ------------------------------------------------------------------------------------------------ model
    def set_up_model(self):
        """reference :689-758 (same constructor order, hence the same CPU RNG stream and initial weights)."""
        if self.logistic_regression:
            self.model = make_logistic_regression(self.D, self.nc, init_sd=self.init_sd,
                                                  mc_samples=self.mc_samples).to(self.device)
        elif self.architecture in {"fn", "residual_fn"}:
            self.model = make_fcnet(self.D, self.n_hidden, self.nc, n_layers=self.n_layers, linear_class=VILinear,
                                    nonl_class=nn.ReLU, mc_samples=self.mc_samples,
                                    residual=(self.architecture == "residual_fn"), init_sd=self.init_sd).to(self.device)
        elif self.architecture == "fn2":
            self.model = make_fc2net(self.D, self.n_hidden, self.nc, mc_samples=self.mc_samples,
                                     init_sd=self.init_sd).to(self.device)
        elif self.architecture == "lenet":
            self.model = make_lenet(linear_class=VILinear, nonl_class=nn.ReLU, mc_samples=self.mc_samples,
                                    init_sd=self.init_sd).to(self.device)
        else:
            raise NotImplementedError(f"architecture {self.architecture!r} is outside the PSVI hot-path scope")
        if isinstance(self.model, MeanFieldMLP) and self.device.type == "cuda":
            self.model.flat()    # (MeanFieldLeNet is a MeanFieldMLP: same flat parameter buffers)

    # ------------------------------------------------------------------------------------------------ data
    def _device_dataset(self, ds, key):
        c = self._dev_data.get(key)
        if c is None or c[0] is not ds:
            raw = torch.as_tensor(ds.data)
            if raw.dtype == torch.uint8 and getattr(ds, "transform", None) is not None:
                # image datasets (MNIST-like): the reference's loaders apply ds.transform (ToTensor + Normalize,
                # experiments_utils.py:42-46) per item; do it once and keep the result on the device
                x = torch.stack([ds[i][0] for i in range(len(ds))]).to(self.device, torch.float32).reshape(len(ds), -1).contiguous()
            else:
                x = raw.to(self.device, torch.float32).reshape(len(ds), -1).contiguous()
            y = torch.as_tensor(ds.targets).to(self.device).to(torch.int32).contiguous()
            c = (ds, x, y)
            self._dev_data[key] = c
        return c[1], c[2]

    def _train_ds(self):
        return getattr(self, "_active_train", None) or self.train_dataset

    def _test_ds(self):
        return getattr(self, "_active_test", None) or self.test_dataset

    def _next_minibatch(self):
        """`next(iter(train_loader))` of the reference (:895): a fresh uniformly random batch of min(B, N) rows every
        outer step -- gathered on the device from the resident copy of the (current) training set."""
        x, y = self._device_dataset(self._train_ds(), "train")
        n = x.shape[0]
        if int(self.data_minibatch) >= n:      # full batch: the data term is a sum over all rows, their order is immaterial
            return x, y
        idx = torch.randperm(n, device=self.device)[: int(self.data_minibatch)]   # drawn on the device (no H2D per step)
        return x[idx], y[idx]

    # ------------------------------------------------------------------------------------------------ main loop
    def run_psvi(self, init_args="subsample", trainer="nested", n_layers=1, logistic_regression=True, n_hidden=None,
                 architecture=None, log_every=10, inner_it=10, data_minibatch=None, lr0net=1e-3, lr0u=1e-3,
                 lr0joint=1e-3, lr0v=1e-2, lr0z=1e-2, init_sd=1e-3, num_epochs=1000, log_pseudodata=False,
                 prune_idx=0, increment_idx=0, gamma=1.0, **kwargs):
        """Run inference (reference :761-1028); returns the same results dict."""
        self.init_args, self.trainer, self.logistic_regression = init_args, trainer, logistic_regression
        self.architecture, self.n_hidden, self.n_layers, self.init_sd = architecture, n_hidden, n_layers, init_sd
        self.log_every, self.log_pseudodata = log_every, log_pseudodata
        self.data_minibatch = data_minibatch
        self.inner_it, self.num_epochs = inner_it, num_epochs
        self.scheduler_optim_net = None
        self.gamma = gamma
        epoch_quarter = (self.N // self.data_minibatch) // 4
        scheduler_kwargs = {"step_size": epoch_quarter if epoch_quarter > 0 else 10000, "gamma": self.gamma}
        self.train_loader = DataLoader(self.train_dataset, batch_size=self.data_minibatch, shuffle=True)
        self.test_loader = DataLoader(self.test_dataset, batch_size=self.data_minibatch, shuffle=False)
        if self.increment:
            # incremental learning (reference :823-832): start with classes {0, 1}; each later task adds one class
            self.incremental_train_datasets = [self.train_dataset.subset_where(cs=list(range(c + 1)) if c == 1 else [c])
                                               for c in range(1, self.nc)]
            self.incremental_test_datasets = [self.test_dataset.subset_where(cs=list(range(c + 1))) for c in range(1, self.nc)]
            self._active_train, self._active_test = self.incremental_train_datasets[0], self.incremental_test_datasets[0]
            self.train_loader = DataLoader(self._active_train, batch_size=self.data_minibatch, shuffle=True)
            self.test_loader = DataLoader(self._active_test, batch_size=self.data_minibatch, shuffle=False)
            self.train_data_so_far = len(self._active_train)
            self.nc = 2
        self.set_up_model()
        nlls_psvi, accs_psvi, core_idcs_psvi, iws_entropy, nesses, vs_entropy, us, zs, vs, grid_preds, times = (
            [], [], [], [], [], [], [], [], [], [], [0])
        {"random": self.pseudo_rand_init, "subsample": self.pseudo_subsample_init}[self.init_args]()
        self.optim_net = torch.optim.Adam(list(self.model.parameters()), lr0net)
        self.optim_u = _adam([self.u], lr0u)
        self.scheduler_optim_net = torch.optim.lr_scheduler.StepLR(self.optim_net, **scheduler_kwargs)
        if self.learn_v:
            self.optim_v = _adam([self.v], lr0v)
        if self.learn_z:
            self.optim_z = _adam([self.z], lr0z)          # reference :869-870
        self._lr0joint = lr0joint
        optimizers = {"alternating": self.alternating_step, "nested": self.nested_step, "hyper": self.hyper_step,
                      "joint": self.joint_step}        # reference :871-886
        psvi_step = optimizers[self.trainer]
        total_checkpts = list(range(self.num_epochs))[::max(int(log_every), 1)]
        downsample = 1
        lpit = total_checkpts[::downsample]     # iterations at which the predictive grid is logged (reference :888-890)
        t_start = time.time()
        log_resource = LogResource()
        for it in tqdm(range(self.num_epochs), disable=kwargs.get("quiet", False)):
            xbatch, ybatch = self._next_minibatch()
            if it % self.log_every == 0:
                test_acc, test_nll, iw_ent, ness, v_ent = self.evaluate()
                if (self.log_pseudodata and it in lpit
                        and self.dnm not in {"MNIST", "FashionMNIST", "Cifar10", "adult", "phishing", "webspam"} and self.D == 2):
                    grid_preds.append(self.pred_on_grid().detach().cpu().numpy().T)      # reference :903-909
                with torch.no_grad():
                    nlls_psvi.append(test_nll.item())
                    accs_psvi.append(test_acc.item())
                    if not kwargs.get("quiet", False):
                        print(f"\npredictive accuracy: {(100*test_acc.item()):.2f}%")
                    core_idcs_psvi.append(self.num_pseudo)
                    times.append(times[-1] + time.time() - t_start)
                    vs.append(self.v.clone().cpu().detach().numpy())
                    if iw_ent is not None:
                        iws_entropy.append(iw_ent.item())
                    if ness is not None:
                        nesses.append(ness.item())
                    if v_ent is not None:
                        vs_entropy.append(v_ent.item())
                    if self.log_pseudodata:
                        us.append(self.u.clone().cpu().detach().numpy())
                        zs.append(self.z.clone().cpu().detach().numpy())
            if self.reset and it % self.reset_interval == 0:
                self.weight_reset()
            psvi_step(xbatch, ybatch)
            log_resource.update()
            # prune the coreset to smaller sizes (reference :935-944)
            if self.prune and it > 0 and it % self.prune_interval == 0 and prune_idx < len(self.prune_sizes):
                self.prune_coreset(to_size=self.prune_sizes[prune_idx], lr0v=lr0v, lr0net=lr0net)
                prune_idx += 1
                self.weight_reset()
            # add a new learning task and grow the coreset to fit it (reference :945-966)
            if (self.increment and it > 0 and it % self.increment_interval == 0
                    and increment_idx < len(self.increment_sizes) - 1):
                increment_idx += 1
                samples = torch.multinomial(self.f(self.v.detach(), 0), self.train_data_so_far, replacement=True)
                self.nc += 1
                self.set_up_model()          # the model is re-initialised with one more class
                self.increment_coreset(to_size=self.increment_sizes[increment_idx], lr0v=lr0v, lr0u=lr0u, lr0net=lr0net,
                                       new_class=increment_idx + 1, increment_idx=increment_idx)
                self._active_train = self.incremental_train_datasets[increment_idx].concatenate(
                    self.u.detach()[samples].cpu().clone(), self.z.detach()[samples].cpu().clone().to(
                        self.incremental_train_datasets[increment_idx].targets.dtype))
                self._active_test = self.incremental_test_datasets[increment_idx]
                self.train_loader = DataLoader(self._active_train, batch_size=self.data_minibatch, shuffle=True)
                self.test_loader = DataLoader(self._active_test, batch_size=self.data_minibatch, shuffle=False)
                self.train_data_so_far = len(self._active_train)
        # retrain the model on the extracted coreset only, for the same number of epochs (reference :969-997)
        if self.retrain_on_coreset:
            if not isinstance(self.model, MeanFieldMLP):
                raise NotImplementedError("retrain_on_coreset covers the mean-field models (flat (mu, rho) parameter buffers)")
            self.weight_reset()
            flat = [torch.nn.Parameter(t) for t in self.model.flat()]      # share storage with the flat (mu, rho) buffers
            opt_retrain = torch.optim.Adam(flat, lr0joint)
            for it in tqdm(range(self.num_epochs), disable=kwargs.get("quiet", False)):
                if it % self.log_every == 0:
                    test_acc, test_nll, iw_ent, ness, v_ent = self.evaluate(correction=False)
                    nlls_psvi.append(test_nll.item())
                    accs_psvi.append(test_acc.item())
                    core_idcs_psvi.append(self.num_pseudo)
                    times.append(times[-1] + time.time() - t_start)
                    vs.append(self.f(self.v.detach(), 0).clone().cpu().numpy())
                    if iw_ent is not None:
                        iws_entropy.append(iw_ent.item())
                    if ness is not None:
                        nesses.append(ness.item())
                    if v_ent is not None:
                        vs_entropy.append(v_ent.item())
                opt_retrain.zero_grad()
                self._retrain_step(opt_retrain, flat)
        resource_data = log_resource.get_resources()
        self.results["accs"] = accs_psvi
        self.results["nlls"] = nlls_psvi
        self.results["csizes"] = core_idcs_psvi
        self.results["times"] = times[1:]
        self.results["elbos"] = self.elbos
        self.results["went"] = iws_entropy
        self.results["ness"] = nesses
        self.results["vent"] = vs_entropy
        self.results["vs"] = vs
        self.results["avg_epoch_time"] = resource_data["time"]
        self.results["gpu_memory"] = resource_data["memory"]
        self.results["chosen_indices"] = self.chosen_indices
        if self.log_pseudodata:
            self.results["us"], self.results["zs"], self.results["grid_preds"] = us, zs, grid_preds
        return self.results

    # ------------------------------------------------------------------------------------------------ evaluation
    def evaluate(self, correction=True, **kwargs):
        """Importance-weighted predictive metrics over the test set (reference :1031-1108): one streaming CUDA pass,
        a fresh noise slab per test batch as in the reference; sharded by test batch over the ranks when distributed.
        Returns (acc, nll, iw_entropy, ness, v_entropy) as 0-dim tensors."""
        assert self.mc_samples > 1
        model, desc, S = self._model_desc()
        if self.learn_z:
            return self._evaluate_learn_z(model, S, correction)
        xt, yt = self._device_dataset(self._test_ds(), "test")
        large_fn = isinstance(model, MeanFieldMLP) and not isinstance(model, MeanFieldLeNet) and self._is_large_fn(model)
        if large_fn:
            self._ws[("force_stream", id(model))] = True
        use_fn_tc = large_fn and self._fits_fn_tc(model) and self.noise_source is None
        if self._use_stream(model) and not use_fn_tc:
            eng = self._stream(model)
            u, _ = self._uv()
            batch = int(self.data_minibatch)
            n_slabs = -(-xt.shape[0] // batch)
            out = eng.evaluate(eng.fam.get_phi(), self._noise_tensor(n_slabs, eng.Pt, S), u, self._z32(), self._a(), xt, yt,
                               batch, mode=0 if correction else 1)
            vs = self.f(self.v.detach(), 0)
            v_entropy = vs.sum().square() / vs.square().sum() / self.num_pseudo if self.compute_weights_entropy else None
            return (out[1] / out[2], out[0] / out[2], out[3] if self.compute_weights_entropy else None, out[4], v_entropy)
        mu, rho = model.flat()
        u, v = self._uv()
        batch = int(self.data_minibatch)
        n = xt.shape[0]
        n_slabs = -(-n // batch)
        dist, rank, world = _dist_info()
        out = torch.zeros(8, device=self.device)
        noise = self._noise(n_slabs)
        lo, hi = shard_bounds(n_slabs, rank, world)
        if hi > lo:
            r0, r1 = lo * batch, min(hi * batch, n)
            if noise.mode == _native.NOISE_EXTERNAL and lo > 0:   # external slabs are indexed from this rank's first
                noise = _native.make_noise(noise._keepalive[lo:hi].contiguous())
            first = 0 if noise.mode == _native.NOISE_EXTERNAL else lo
            if use_fn_tc or self._use_tensor_core_eval(model, r1 - r0, batch):
                # tensor path (large fn: psvi_fn_predictive_tc; large single-layer model: psvi_lr_predictive_tc, DESIGN.md 4.5 /
                # 4.7), bf16 operands: ONE native call walks this rank's test batches (one noise slab each)
                xb16 = self._device_bf16(xt, "test")
                rows_slab = min(batch, r1 - r0)
                need = (_native.fn_tc_scratch_floats(desc, rows_slab, u.shape[0]) if use_fn_tc
                        else _native.lr_predictive_tc_scratch_floats(desc)) + 256
                scratch = self._buf("eval_tc", need)
                _native.predictive_tc_slabs(desc, noise, mu, rho, u, self._z32(), v, xb16[r0:r1], yt[r0:r1], batch, first,
                                            float(self.N), self._vmode, self._alpha_value(), 0 if correction else 1, out, scratch)
            else:
                scratch = self._buf("eval", _native.eval_scratch_floats(desc, r1 - r0, batch))
                _native.evaluate(desc, noise, mu, rho, u, self._z32(), v, xt[r0:r1], yt[r0:r1], batch, first,
                                 float(self.N), self._vmode, self._alpha_value(), 0 if correction else 1, out, scratch)
            if hi != n_slabs:
                out[3:5] = 0.0   # Q12: the weight diagnostics are those of the globally last batch
        if dist is not None:
            dist.all_reduce(out)
        vs = self.f(self.v.detach(), 0)
        v_entropy = vs.sum().square() / vs.square().sum() / self.num_pseudo if self.compute_weights_entropy else None
        return (out[1] / out[2], out[0] / out[2], out[3] if self.compute_weights_entropy else None, out[4], v_entropy)

    tensor_core_eval = "auto"   # "auto": use the tcgen05 predictive kernel for >= 64k-row slabs; True / False force it

    def _use_tensor_core_eval(self, model, rows, batch):
        dims = model.dims
        fits = (len(dims) == 2 and dims[0] % 64 == 0 and 64 <= dims[0] <= 256 and dims[1] <= 16
                and model.n_samples() <= 16)
        if self.tensor_core_eval is True:
            if not fits:
                raise NotImplementedError("tensor-core predictive kernel: needs a single-layer model, D % 64 == 0, "
                                          "D <= 256, C <= 16, S <= 16")
            return True
        return bool(fits and self.tensor_core_eval == "auto" and min(rows, batch) >= 65536)

    @staticmethod
    def _fits_fn_tc(model):
        d = model.dims
        return len(d) == 3 and d[0] % 64 == 0 and 64 <= d[0] <= 256 and d[1] % 128 == 0 and d[2] <= 16 and model.n_samples() <= 64

    def _device_bf16(self, x, key):
        c = self._dev_data.get(key + "_bf16")
        if c is None or c[0] is not x:
            xb = torch.empty(x.shape, device=x.device, dtype=torch.bfloat16)
            _native.f32_to_bf16(x.contiguous(), xb)
            c = (x, xb)
            self._dev_data[key + "_bf16"] = c
        return c[1]

    def weight_reset(self):
        """Reset variational parameters to initialisation (reference :1110-1128)."""
        for layer in self.model.modules():
            if isinstance(layer, (VILinear, VILinearMultivariateNormal)) and hasattr(layer, "reset_parameters_variational"):
                layer.reset_parameters_variational()   # in-place inits: the flat-buffer views stay valid
            elif isinstance(layer, nn.Conv2d) and hasattr(layer, "reset_parameters"):
                layer.reset_parameters()               # reference :1121-1126 (VIConv2d falls in the nn.Conv2d branch)

    def pred_on_grid(self, n_test_per_dim=250, correction=True, **kwargs):
        """Predictive probabilities over the 2-d grid [-3, 4] x [-2, 3] (reference :1130-1175): importance-weighted mixture
        under one noise draw.  Returns [n_test_per_dim**2, nc] (row-major over (x0, x1), as the reference's view(-1, 2))."""
        model, desc, S = self._model_desc()
        x0 = torch.linspace(-3, 4, n_test_per_dim)
        x1 = torch.linspace(-2, 3, n_test_per_dim)
        grid = torch.stack(torch.meshgrid(x0, x1, indexing="ij"), dim=-1).reshape(-1, 2).to(self.device).contiguous()
        eng = self._stream(model)
        u, _ = self._uv()
        return eng.predict_probs(eng.fam.get_phi(), self._noise_tensor(1, eng.Pt, S)[0], u, self._z32(), self._a(), grid,
                                 correction=correction)

    def prune_coreset(self, to_size, lr0v=1e-3, lr0net=1e-4):
        """Prune the coreset to a smaller size by sampling points without replacement from f(v) (reference :1177-1193;
        designed for the fixed-u methods)."""
        self.num_pseudo = to_size
        keep_v = torch.multinomial(self.f(self.v.detach(), 0), to_size, replacement=False)
        self.v = torch.zeros_like(self.v[keep_v]).clone().detach().requires_grad_(True)
        self.optim_v = _adam([self.v], lr0v)
        self.u = torch.index_select(self.u.detach(), 0, keep_v).requires_grad_(True)
        self.z = torch.index_select(self.z, 0, keep_v)
        self.optim_u = _adam([self.u], self.optim_u.param_groups[0]["lr"])
        self.optim_net = torch.optim.Adam(list(self.model.parameters()), lr0net)

    def increment_coreset(self, to_size, lr0v=1e-3, lr0u=1e-3, lr0net=1e-4, variance=1.0, new_class=2, increment_idx=1):
        """Grow the coreset to `to_size` points for a new learning task (reference :1194-1217): the new points get the mean
        weight of the old ones and come from the new task's data (or from the noisy empirical mean with init_args="random")."""
        n_old = len(self.v)
        self.num_pseudo, extra = to_size, to_size - n_old
        with torch.no_grad():
            vv = self.v.detach()
            self.v = torch.cat((vv, 1.0 / (n_old + extra) * vv.sum() * torch.ones(extra, device=self.device)))
        self.v = self.v.detach().requires_grad_(True)
        self.optim_v = _adam([self.v], lr0v)
        if self.init_args == "random":
            new_us = (compute_empirical_mean(self.train_loader) + variance * torch.randn(extra, self.D)).clone()
            new_zs = new_class * torch.ones(extra)
        else:
            ds = self.incremental_train_datasets[increment_idx]
            new_us, new_zs = ds[torch.randperm(len(ds))[:extra]]
        self.u = torch.cat((self.u.detach(), new_us.to(self.device).float())).detach().requires_grad_(True)
        self.z = torch.cat((self.z, new_zs.to(self.device).to(self.z.dtype)))
        self.optim_u = _adam([self.u], lr0u)
        self.optim_net = torch.optim.Adam(list(self.model.parameters()), lr0net)

    def _retrain_step(self, opt, params):
        """One Adam step of the model on inner_elbo over the (fixed) coreset (reference :994-997)."""
        self.inner_elbo(model=self.model)
        g, P = self._last_inner, params[0].numel()
        params[0].grad, params[1].grad = g[:P].clone(), g[P:].clone()
        opt.step()


class PSVILearnV(PSVI):
    r"""PSVI with learnable v on a simplex: f = softmax (reference :1344-1360)."""

    _vmode = _native.VMODE_SOFTMAX

    def __init__(self, learn_v=True, parameterised=True, **kwargs):
        super().__init__(**kwargs)
        self.learn_v, self.parameterised = learn_v, parameterised
        with torch.no_grad():
            self.v = torch.zeros(self.num_pseudo, device=self.device)
        self.v.requires_grad_(True)
        self.f = torch.softmax


class PSVI_No_Rescaling(PSVI):
    r"""PSVI without any rescaling of the coreset likelihood (reference :1363-1373)."""

    def __init__(self, **kwargs):
        super().__init__(**kwargs)
        self.v *= 1.0 / self.N


class PSVIFreeV(PSVI):
    r"""PSVI with learnable non-negative v (reference :1376-1385)."""

    def __init__(self, learn_v=True, **kwargs):
        super().__init__(**kwargs)
        self.learn_v = True
        self.v.requires_grad_(True)


class PSVIAV(PSVILearnV):
    r"""Learnable simplex weights and learnable total evidence: f = exp(alpha) softmax(v) (reference :1475-1619)."""

    _vmode = _native.VMODE_EXPALPHA_SOFTMAX

    def __init__(self, learn_v=True, **kwargs):
        super().__init__(**kwargs)
        self.alpha = torch.tensor([0.0], device=self.device)
        self.alpha.requires_grad_(True)
        self.f = lambda *x: torch.exp(self.alpha.detach()) * torch.softmax(x[0], x[1])
        self.optim_alpha = _adam([self.alpha], self.lr0alpha)
        self.results["alpha"] = []

    def evaluate(self, **kwargs):
        self.results["alpha"].append(self.alpha.clone().cpu().detach().numpy())
        return super().evaluate(**kwargs)

    def increment_coreset(self, lr0alpha=1e-3, **kwargs):
        super().increment_coreset(**kwargs)          # reference :1501-1503
        self.optim_alpha = _adam([self.alpha], lr0alpha)

    def hyper_step(self, xbatch, ybatch, T=10, inner_opt_class=None, K=10, linsys_lr=1e-1, hypergrad_approx="CG_normaleq",
                   **kwargs):
        """Reference :1505-1585: the same implicit step as PSVI.hyper_step with alpha as a third hyper-parameter, but with
        its own defaults (K = 10 solver iterations, linear-system step 1e-1) -- run_psvi calls it without arguments."""
        return super().hyper_step(xbatch, ybatch, T=T, inner_opt_class=inner_opt_class, K=K, linsys_lr=linsys_lr,
                                  hypergrad_approx=hypergrad_approx, **kwargs)


def _no_hyper_for_fixed_u(self, *args, **kwargs):
    # reference :1660-1740 / :1790-1883: the fixed-u variants solve the linear system with an ADAM fixed-point map
    # (hypergrad.DifferentiableAdam(step_size=linsys_lr), K = 20) whose Jacobian products are not Hessian-vector products of
    # the inner objective.  Upstream that call cannot run: the map is handed the weights WITHOUT the Adam moment tensors it
    # unpacks as params[n:2n], params[2n:] (diff_optimizers.py:144-146), so `hyper_step` dies with an IndexError in
    # robust_higher/patch.py:117 (checked against the unmodified reference).  There is no behaviour to mirror; substituting the
    # gradient-descent map would invent one.
    raise NotImplementedError(f"--trainer hyper is not available for {type(self).__name__}: the reference's own hyper_step for "
                              "the fixed-u variants (psvi_classes.py:1660-1740) raises an IndexError; use --trainer nested")


class PSVIFixedU(PSVILearnV):
    r"""Fixed coreset locations, learnable weights (reference :1622-1740): the u update is skipped."""

    hyper_step = _no_hyper_for_fixed_u

    def _step_outer_optimisers(self):
        self.u.grad = None
        if self.learn_v:
            self.optim_v.step()


class PSVIAFixedU(PSVIAV):
    r"""Fixed locations, learnable weights and evidence scale (reference :1743-1883)."""

    hyper_step = _no_hyper_for_fixed_u

    def _step_outer_optimisers(self):
        self.u.grad = None
        if self.learn_v:
            self.optim_v.step()
            self.optim_alpha.step()


def _out_of_scope(name, where):
    class _Stub(PSVI):
        def __init__(self, *a, **k):
            raise NotImplementedError(f"{name} ({where}) is outside the PSVI hot path built so far "
                                      "(SURVEY.md section 8f item 1)")
    _Stub.__name__ = name
    return _Stub


class PSVI_Ablated(PSVILearnV):
    r"""PSVI with ablated importance sampling (reference :1388-1408): the outer objective is
    mean_s (N/B) sum_b nll[s, b] - mean_s sampled_nkl_s.  Runs on the streaming engine (per-sample network kernels)."""

    _outer_kind = "ablated"

    def _use_stream(self, model):
        return True

    def psvi_elbo(self, xbatch, ybatch, model=None, params=None, hyperopt=False):
        model, desc, S = self._model_desc(model)
        xb = xbatch.detach().to(self.device, torch.float32).reshape(xbatch.shape[0], -1).contiguous()
        yb = ybatch.detach().to(self.device).to(torch.int32).contiguous()
        eng = self._stream(model)
        loss, pbar = eng.outer_grad_ablated(eng.fam.get_phi(), self._noise_tensor(1, eng.Pt, S)[0], xb, yb, float(self.N))
        u, v = self._uv()
        self._last_outer = dict(phi_grad=pbar, u_grad=torch.zeros_like(u), v_grad=torch.zeros_like(v), alpha_grad=None)
        return loss


class PSVI_No_IW(PSVI_Ablated):
    r"""Single-sample training / multi-sample testing (reference :1411-1472).  With mc_samples == 1 the reference's
    inner_elbo scores pseudo-point i against EVERY label z_j (logits.unsqueeze_(1) at :493-494 makes
    Categorical.log_prob broadcast to [M, M]; the matmul with N f(v) weights column j by a_j): inner = sum_i sum_c A_c
    nll(u_i, c) + kl with A_c = sum_{j: z_j = c} a_j.  Reproduced by expanding the pseudo-data to M C weighted rows."""

    def __init__(self, **kwargs):
        super().__init__(**kwargs)
        self.mc_samples = 1

    def _inner_pseudo(self, u, z32, a):
        C, M = int(self.nc), u.shape[0]
        A = torch.zeros(C, device=a.device, dtype=a.dtype).index_add_(0, z32.long(), a)
        z2 = torch.arange(C, device=u.device, dtype=torch.int32).repeat(M)
        return u.repeat_interleave(C, 0).contiguous(), z2.contiguous(), A[z2.long()].contiguous()

    def _collapse_pseudo(self, ubar, abar):
        C, M = int(self.nc), self.u.shape[0]
        return ubar.reshape(M, C, -1).sum(1), abar.reshape(M, C).sum(0)[self._z32().long()]

    def inner_elbo(self, model=None, params=None, hyperopt=False):
        model, desc, S = self._model_desc(model)
        eng = self._stream(model)
        u, _ = self._uv()
        u_in, z_in, a_in = self._inner_pseudo(u, self._z32(), self._a())
        val, g = eng.inner_grad(eng.fam.get_phi(), self._noise_tensor(1, eng.Pt, S)[0], u_in, z_in, a_in)
        self._last_inner = g
        return val.float()

    def evaluate(self, correction=True, mc_samples_eval=5, mc_samples_train=1, **kwargs):
        self.mc_samples = mc_samples_eval
        set_mc_samples(self.model, self.mc_samples)       # multi-sample for testing
        try:
            return super().evaluate(correction=True, **kwargs)
        finally:
            self.mc_samples = 1
            set_mc_samples(self.model, mc_samples_train)  # single-sample for training


class PSVIEvaluate(PSVI):
    r"""Evaluation of a FIXED coreset (reference :1885-1938): only the network is (re)trained -- nested_step runs the T inner
    Adam steps on the soft-label inner objective, reports psvi_elbo at the result and copies the fast weights back; u, z, v and
    alpha stay as they are (no hypergradient is taken: the reference calls backward() on frozen leaves).  The coreset comes from
    the standard initialisers here; the reference's loaders of saved distilled sets (custom_init*, :310-443) are out of scope."""

    _vmode = _native.VMODE_EXPALPHA_SOFTMAX

    def __init__(self, learn_v=False, parameterised=False, **kwargs):
        kwargs.pop("learn_z", None)
        super().__init__(learn_z=True, **kwargs)
        self.learn_v = False
        self.alpha = torch.tensor([0.0], device=self.device)
        self.f = lambda *x: torch.exp(self.alpha.detach()) * torch.softmax(x[0], x[1])
        self.ablated_weights = self.ablated_alpha = self.ablated_labels = True

    def nested_step(self, xbatch, ybatch):
        from psvi.inference.stream import B1, B2
        self._require_single_rank_learn_z()
        model, _, S = self._model_desc()
        eng = self._stream(model)
        T, lr = int(self.inner_it), float(self.optim_net.param_groups[0]["lr"])
        u, _ = self._uv()
        a = self._a()
        tin = self._soft_targets()[0].detach().float()
        ue, le, _ = self._soft_rows(u, tin)
        ae = (a[:, None] * tin).reshape(-1)
        a_exp = ae.expand(S, ae.numel()).contiguous()
        kin = S * (a.double() * torch.xlogy(tin, tin).sum(1).double()).sum()
        phi = eng.fam.get_phi().contiguous()
        m, v = torch.zeros_like(phi), torch.zeros_like(phi)
        eps_all = eng.fam.fix_eps(self._noise_tensor(T, eng.Pt, S))
        for t in range(T):
            log = self.register_elbos and t % max(int(self.log_every), 1) == 0
            val, g = eng.inner_grad(phi, eps_all[t], ue, le, ae, want_val=log, a_exp=a_exp, fixed=True)
            if log:
                self.elbos.append((1, -(val + kin).item()))
            phi, m, v = _native.adam_unroll_step(phi, g, m, v, lr / (1.0 - B1 ** (t + 1)), math.sqrt(1.0 - B2 ** (t + 1)))
        eng.fam.set_phi(phi)
        loss = self.psvi_elbo(xbatch, ybatch, model=self.model)
        if self.register_elbos:
            self.elbos.append((0, -loss.item()))
        if self.scheduler_optim_net:
            self.scheduler_optim_net.step()
        return loss


class PSVI_regressor(PSVI):
    r"""PSVI for BNN regression (reference :1940-2264): Gaussian likelihood of precision `tau` on a one-output `regressor_net`,
    learnable pseudo-inputs u AND pseudo-targets z.  The per-sample network passes are psvi_net_pass_gaussian (csrc/
    psvi_mf_stream.cu) under the streaming engine, so nested_step is the same unrolled-Adam / reverse-sweep step with one more
    hypergradient (on z).  NB upstream these classes cannot be constructed (`device_id` NameError at :1975; PSVIAV_regressor reads an
    unset `scheduler_optim_net` at :2329); the goldens were taken with both supplied from outside (oracle/make_goldens_r2.py)."""

    def __init__(self, u=None, z=None, train_dataset=None, val_dataset=None, test_dataset=None, y_mean=None, y_std=None, N=None,
                 D=None, optim=None, optim_u=None, optim_net=None, optim_v=None, optim_z=None, register_elbos=False,
                 num_pseudo=None, seed=0, compute_weights_entropy=True, mc_samples=None, learn_v=False, f=lambda *x: x[0],
                 dnm=None, nc=1, init_dataset=None, parameterised=False, learn_z=True, lr0alpha=1e-3, tau=0.1,
                 logistic_regression=False, **kwargs):
        super().__init__(u=u, z=z, train_dataset=train_dataset, test_dataset=test_dataset, N=N, D=D, optim=optim, optim_u=optim_u,
                         optim_net=optim_net, optim_v=optim_v, optim_z=optim_z, register_elbos=register_elbos,
                         num_pseudo=num_pseudo, seed=seed, compute_weights_entropy=compute_weights_entropy, mc_samples=mc_samples,
                         learn_v=learn_v, f=f, dnm=dnm, nc=nc, init_dataset=init_dataset, parameterised=parameterised,
                         learn_z=False, lr0alpha=lr0alpha, device_id=kwargs.get("device_id"),
                         noise_source=kwargs.get("noise_source"))
        from functools import partial
        from psvi.models.neural_net import gaussian_fn
        self.val_dataset, self.y_mean, self.y_std, self.tau = val_dataset, y_mean, y_std, float(tau)
        self.logistic_regression = logistic_regression
        self.distr_fn = partial(gaussian_fn, scale=1.0 / np.sqrt(tau))
        self.learn_targets = bool(learn_z)      # (`learn_z` of the base class means SOFT LABELS; here z are real-valued targets)
        self.scheduler_optim_net = None

    # ---- plumbing
    def _use_stream(self, model):
        return True

    def _stream(self, model):
        from psvi.inference.stream import GaussMlpNet, MeanFieldFamily, StreamEngine
        key = ("stream_gauss", id(model), model.n_samples())
        eng = self._ws.get(key)
        if eng is None:
            if model.dims[-1] != 1:
                raise NotImplementedError("the Gaussian likelihood is built for one network output (nc = 1)")
            eng = StreamEngine(MeanFieldFamily(model), model.dims, model.n_samples(),
                               net=GaussMlpNet(model.dims, model.n_samples(), self.tau))
            self._ws[key] = eng
        eng.fam.mu, eng.fam.rho = model.flat()
        return eng

    def _zf(self):
        return self.z.detach().to(self.device, torch.float32).reshape(-1).contiguous()

    def _xy(self, xbatch, ybatch):
        return (xbatch.detach().to(self.device, torch.float32).reshape(xbatch.shape[0], -1).contiguous(),
                ybatch.detach().to(self.device, torch.float32).reshape(-1).contiguous())

    def set_up_model(self):
        from psvi.models.neural_net import make_regressor_net
        if self.architecture != "regressor_net":
            raise NotImplementedError(f"the regressors run architecture 'regressor_net' (got {self.architecture!r})")
        self.model = make_regressor_net(self.D, self.n_hidden, self.nc, linear_class=VILinear, nonl_class=nn.ReLU,
                                        mc_samples=self.mc_samples, init_sd=self.init_sd).to(self.device)   # reference :743-753
        self.model.flat()

    def pseudo_subsample_init(self):
        """reference :2019-2031: `random.sample` of the training rows; u and z both learnable."""
        import random
        idx = random.sample(range(len(self.train_dataset)), self.num_pseudo)
        x, y = torch.as_tensor(self.train_dataset.data), torch.as_tensor(self.train_dataset.targets)
        self.u = x[idx].clone().to(self.device).float().requires_grad_(True)
        self.z = y[idx].clone().to(self.device).float().requires_grad_(True)

    def _next_minibatch(self):
        x, y = self._device_regression(self.train_dataset, "train")
        n, B = x.shape[0], int(self.data_minibatch)
        if B >= n:
            return x, y
        idx = torch.randperm(n, device=self.device)[:B]
        return x[idx], y[idx]

    def _device_regression(self, ds, key):
        c = self._dev_data.get(key)
        if c is None or c[0] is not ds:
            c = (ds, torch.as_tensor(ds.data).to(self.device, torch.float32).reshape(len(ds), -1).contiguous(),
                 torch.as_tensor(ds.targets).to(self.device, torch.float32).reshape(-1).contiguous())
            self._dev_data[key] = c
        return c[1], c[2]

    # ---- objectives (reference :2034-2057)
    def psvi_elbo(self, xbatch, ybatch, model=None, params=None, hyperopt=False):
        assert self.mc_samples > 1
        model = self.model if model is None else model
        eng, S = self._stream(model), model.n_samples()
        u, _ = self._uv()
        xb, yb = self._xy(xbatch, ybatch)
        ex = {}
        loss, pbar, ubar, abar, _ = eng.outer_grad(eng.fam.get_phi(), self._noise_tensor(1, eng.Pt, S)[0], u, self._zf(), self._a(),
                                                   xb, yb, float(self.N), extras=ex)
        vg, ag = self._v_grad_from_abar(abar)
        self._last_outer = dict(phi_grad=pbar, u_grad=ubar, v_grad=vg, alpha_grad=ag, z_grad=ex.get("zbar_outer"))
        return loss

    def inner_elbo(self, model=None, params=None, hyperopt=False):
        model = self.model if model is None else model
        eng, S = self._stream(model), model.n_samples()
        u, _ = self._uv()
        val, g = eng.inner_grad(eng.fam.get_phi(), self._noise_tensor(1, eng.Pt, S)[0], u, self._zf(), self._a())
        self._last_inner = g
        return val.float()

    # ---- one bilevel step (reference :2059-2093; PSVIAV_regressor :2303-2335)
    def nested_step(self, xbatch, ybatch):
        if _dist_info()[2] > 1:
            raise NotImplementedError("the regressors run on one rank")
        self._zero_grads()
        if self.learn_targets and self.optim_z is not None:
            self.optim_z.zero_grad()
        model = self.model
        eng, S = self._stream(model), model.n_samples()
        T, lr = int(self.inner_it), float(self.optim_net.param_groups[0]["lr"])
        u, _ = self._uv()
        xb, yb = self._xy(xbatch, ybatch)
        ex = {}
        loss, ubar, abar, phi_T, il = eng.nested_cached(eng.fam.get_phi(), self._noise_tensor(T + 1, eng.Pt, S), u, self._zf(),
                                                        self._a(), xb, yb, float(self.N), T, lr,
                                                        want_losses=self.register_elbos, extras=ex)
        eng.fam.set_phi(phi_T)
        if self.register_elbos:
            ilc = il.cpu()
            for in_it in range(0, T, max(int(self.log_every), 1)):
                self.elbos.append((1, -ilc[in_it].item()))
            self.elbos.append((0, -loss.item()))
        vg, ag = self._v_grad_from_abar(abar)
        self.u.grad = ubar.to(self.u.dtype).reshape(self.u.shape)
        if self.learn_v:
            self.v.grad = vg.to(self.v.dtype)
        if self.alpha is not None and self.alpha.requires_grad and ag is not None:
            self.alpha.grad = ag.to(self.alpha.dtype)
        zbar = ex.get("zbar", ex.get("zbar_outer"))
        self.z.grad = zbar.to(self.z.dtype).reshape(self.z.shape)
        self._step_outer_optimisers()           # u, v (+ clamp), alpha
        if self.learn_targets and self.optim_z is not None:
            self.optim_z.step()
        if self.scheduler_optim_net:
            self.scheduler_optim_net.step()
        return loss

    # ---- predictive metrics (reference :2221-2264)
    def evaluate(self, correction=True, **kwargs):
        """(rmse, mean log-likelihood) over the test set.  As in the reference the pseudo term is summed over samples before
        it meets the log-weights, so the weights are softmax(sampled_nkl); predictions are de-normalised with (y_mean, y_std)
        and scored against the raw test targets."""
        assert self.mc_samples > 1
        model = self.model
        eng, S = self._stream(model), model.n_samples()
        xt, yt = self._device_regression(self.test_dataset, "test")
        batch = int(self.data_minibatch)
        n_slabs = -(-xt.shape[0] // batch)
        eps = eng.fam.fix_eps(self._noise_tensor(n_slabs, eng.Pt, S))
        phi = eng.fam.get_phi()
        ym, ys = float(self.y_mean), float(self.y_std)
        se = torch.zeros((), device=self.device, dtype=torch.float64)
        ll = torch.zeros((), device=self.device, dtype=torch.float64)
        for k, r0 in enumerate(range(0, xt.shape[0], batch)):
            theta = eng.fam.sample(phi, eps[k])
            w = torch.softmax(eng.fam.nkl(phi, eps[k], theta).double(), 0)
            out = eng.net.logits(theta, xt[r0:r0 + batch].contiguous())[..., 0].double()
            yp = ((out * ys + ym) * w[:, None]).sum(0)
            d = yp - yt[r0:r0 + batch].double()
            se += (d * d).sum()
            ll += (-0.5 * self.tau * d * d - 0.5 * math.log(2.0 * math.pi / self.tau)).sum()
        n = float(xt.shape[0])
        return (se / n).sqrt().float(), (ll / n).float()

    # ---- main loop (reference :2095-2218)
    def run_psvi(self, init_args="subsample", trainer="nested", n_layers=1, n_hidden=None, architecture=None, log_every=10,
                 inner_it=10, data_minibatch=None, lr0net=1e-3, lr0u=1e-3, lr0v=1e-2, lr0z=1e-2, init_sd=1e-3, num_epochs=1000,
                 log_pseudodata=False, **kwargs):
        self.init_args, self.trainer = init_args, trainer
        self.architecture, self.n_hidden, self.n_layers, self.init_sd = architecture, n_hidden, n_layers, init_sd
        self.log_every, self.log_pseudodata, self.data_minibatch = log_every, log_pseudodata, data_minibatch
        self.inner_it, self.num_epochs = inner_it, num_epochs
        self.set_up_model()
        lls, rmses, csizes, us, zs, vs, times = [], [], [], [], [], [], [0]
        self.train_loader = DataLoader(self.train_dataset, batch_size=self.data_minibatch, shuffle=True)
        self.test_loader = DataLoader(self.test_dataset, batch_size=self.data_minibatch, shuffle=False)
        {"subsample": self.pseudo_subsample_init}[self.init_args]()
        self.optim_net = torch.optim.Adam(list(self.model.parameters()), lr0net)
        self.optim_u = _adam([self.u], lr0u)
        if self.learn_v:
            self.optim_v = _adam([self.v], lr0v)
        if self.learn_targets:
            self.optim_z = _adam([self.z], lr0z)
        psvi_step = {"nested": self.nested_step}[self.trainer]
        t_start = time.time()
        for it in tqdm(range(self.num_epochs), disable=kwargs.get("quiet", False)):
            xbatch, ybatch = self._next_minibatch()
            if it % self.log_every == 0:
                test_rmse, test_ll = self.evaluate(**kwargs)
                lls.append(test_ll.item())
                rmses.append(test_rmse.item())
                csizes.append(self.num_pseudo)
                times.append(times[-1] + time.time() - t_start)
                vs.append(self.f(self.v.detach(), 0).clone().cpu().numpy())
                if self.log_pseudodata:
                    us.append(self.u.clone().cpu().detach().numpy())
                    zs.append(self.z.clone().cpu().detach().numpy())
            psvi_step(xbatch, ybatch)
        self.results.update(rmses=rmses, lls=lls, csizes=csizes, times=times[1:], went=[], ness=[], vent=[], vs=vs)
        return self.results


class PSVILearnV_regressor(PSVI_regressor):
    r"""Learnable simplex weights (reference :2268-2280)."""

    _vmode = _native.VMODE_SOFTMAX

    def __init__(self, learn_v=True, parameterised=True, **kwargs):
        super().__init__(**kwargs)
        self.learn_v, self.parameterised = learn_v, parameterised
        with torch.no_grad():
            self.v = torch.zeros(self.num_pseudo, device=self.device)
        self.v.requires_grad_(True)
        self.f = torch.softmax


class PSVIAV_regressor(PSVILearnV_regressor):
    r"""... and a learnable total evidence exp(alpha) (reference :2283-2335)."""

    _vmode = _native.VMODE_EXPALPHA_SOFTMAX

    def __init__(self, learn_v=True, **kwargs):
        super().__init__(**kwargs)
        self.alpha = torch.tensor([0.0], device=self.device)
        self.alpha.requires_grad_(True)
        self.f = lambda *x: torch.exp(self.alpha.detach()) * torch.softmax(x[0], x[1])
        self.optim_alpha = _adam([self.alpha], self.lr0alpha)
        self.results["alpha"] = []

    def evaluate(self, **kwargs):
        self.results["alpha"].append(self.alpha.clone().cpu().detach().numpy())
        return super().evaluate(**kwargs)
