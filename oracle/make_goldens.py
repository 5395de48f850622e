"""Generate tests/golden/*.npz by running the UNMODIFIED reference (TEST INFRASTRUCTURE ONLY).

Run in the build container (where /root/reference exists):   python oracle/make_goldens.py
Every golden holds the exact inputs (initial variational parameters, pseudo-data, minibatch, noise seed) and
the reference's outputs for: inner_elbo (+autograd grads), psvi_elbo (+grads), one full nested_step
(hypergradients on u and v, final fast weights, u/v after their Adam step) and evaluate -- in fp32 (the
reference's arithmetic) and in fp64 (same fp32 noise, used to decide "who is wrong": SURVEY.md section 4).
Noise is injected through oracle.ref_import.NoiseFeeder, so the fixture only stores the seed.
"""
from __future__ import annotations

import contextlib
import io
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from oracle.ref_import import import_reference, NoiseFeeder  # noqa: E402

import_reference()
import torch  # noqa: E402
from psvi.inference.psvi_classes import PSVILearnV, PSVI  # noqa: E402
from psvi.inference.baselines import run_mfvi_subset  # noqa: E402
from psvi.inference.utils import pseudo_subsample_init  # noqa: E402
from psvi.experiments.experiments_utils import read_dataset  # noqa: E402
from psvi.models.neural_net import VILinear  # noqa: E402

CASES = [
    # name, dataset, architecture, H, n_layers, M, S, T, B, init_sd, lr0net, cls
    dict(name="logreg_hm_m10", dnm="halfmoon", arch="logistic_regression", H=0, n_layers=0, M=10, S=10, T=8, B=128,
         init_sd=1e-3, lr0net=1e-3, cls="learn_v"),
    dict(name="fn_hm_m50_t10", dnm="halfmoon", arch="fn", H=100, n_layers=1, M=50, S=10, T=10, B=128,
         init_sd=1e-3, lr0net=1e-3, cls="learn_v"),
    dict(name="fn_hm_m50_t100", dnm="halfmoon", arch="fn", H=100, n_layers=1, M=50, S=10, T=100, B=128,
         init_sd=1e-3, lr0net=1e-3, cls="learn_v"),
    dict(name="fn_hm_m10_sd1e-6", dnm="halfmoon", arch="fn", H=100, n_layers=1, M=10, S=10, T=20, B=128,
         init_sd=1e-6, lr0net=1e-3, cls="learn_v"),
    dict(name="fn_fb_m10", dnm="four_blobs", arch="fn", H=100, n_layers=1, M=10, S=10, T=10, B=128,
         init_sd=1e-3, lr0net=1e-3, cls="learn_v"),
    dict(name="fn_fb_l2_m13", dnm="four_blobs", arch="fn", H=24, n_layers=2, M=13, S=6, T=6, B=37,
         init_sd=1e-2, lr0net=3e-3, cls="learn_v"),
    dict(name="fn_hm_psvi_fixedv", dnm="halfmoon", arch="fn", H=40, n_layers=1, M=12, S=4, T=5, B=64,
         init_sd=1e-3, lr0net=1e-3, cls="psvi"),
]


def get_data(dnm):
    torch.manual_seed(0)  # four_blobs draws from the global RNG before any seeding; pin it
    with contextlib.redirect_stdout(io.StringIO()):
        return read_dataset(dnm, {"data_folder": "/tmp/psvi_data", "test_ratio": 0.2})


def model_dims(model):
    lins = [m for m in model.modules() if isinstance(m, VILinear)]
    return [lins[0].in_features] + [m.out_features for m in lins]


def get_mu_rho(model):
    mu, rho = [], []
    for m in model.modules():
        if isinstance(m, VILinear):
            mu += [m.weight.detach().reshape(-1), m.bias.detach().reshape(-1)]
            rho += [m._weight_sd.detach().reshape(-1), m._bias_sd.detach().reshape(-1)]
    return torch.cat(mu).double().numpy().copy(), torch.cat(rho).double().numpy().copy()


def run_case(c, dt):
    x, y, xt, yt, N, D, tr, te, nc = get_data(c["dnm"])
    kw = dict(mc_samples=c["S"], num_epochs=0, data_minibatch=c["B"], D=D, N=N, inner_it=c["T"], trainer="nested",
              log_every=10, lr0u=1e-4, lr0net=c["lr0net"], lr0v=1e-3, init_args="subsample", init_sd=c["init_sd"],
              num_pseudo=c["M"], seed=0, architecture=c["arch"], n_hidden=c["H"], n_layers=c["n_layers"],
              logistic_regression=(c["arch"] == "logistic_regression"), train_dataset=tr, test_dataset=te,
              dnm=c["dnm"], nc=nc, data_folder="/tmp/psvi_data", compute_weights_entropy=True, register_elbos=True)
    Cls = PSVILearnV if c["cls"] == "learn_v" else PSVI
    with contextlib.redirect_stdout(io.StringIO()), contextlib.redirect_stderr(io.StringIO()), np.errstate(all="ignore"):
        import warnings
        warnings.simplefilter("ignore")
        obj = Cls(**kw)
        obj.run_psvi(**kw)
    tdt = torch.float32 if dt == "32" else torch.float64
    obj.model.to(tdt)
    rng = np.random.default_rng(1234)
    v0 = (0.3 * rng.standard_normal(c["M"])).astype(np.float32) if c["cls"] == "learn_v" else obj.v.detach().numpy()
    obj.v = torch.tensor(v0, dtype=tdt).requires_grad_(c["cls"] == "learn_v")
    obj.u = obj.u.detach().to(tdt).requires_grad_(True)
    obj.z = obj.z.to(tdt)
    # optimisers must point at the re-typed leaves
    obj.optim_net = torch.optim.Adam(list(obj.model.parameters()), c["lr0net"])
    obj.optim_u = torch.optim.Adam([obj.u], 1e-4)
    if obj.learn_v:
        obj.optim_v = torch.optim.Adam([obj.v], 1e-3)
    obj.scheduler_optim_net = None
    xb, yb = x[: c["B"]].to(tdt), y[: c["B"]].to(tdt)
    dims = model_dims(obj.model)
    mu0, rho0 = get_mu_rho(obj.model)
    out = dict(dims=np.array(dims), N=N, S=c["S"], T=c["T"], M=c["M"], B=c["B"], lr0net=c["lr0net"], noise_seed=777,
               vmode=1 if c["cls"] == "learn_v" else 0,
               mu0=mu0, rho0=rho0, u0=obj.u.detach().double().numpy().copy(), z=obj.z.double().numpy().copy(), v0=v0.astype(np.float64),
               xb=xb.double().numpy().copy(), yb=yb.double().numpy().copy(),
               xt=xt.double().numpy().copy(), yt=yt.double().numpy().copy())
    res = {}
    params = list(obj.model.parameters())
    with NoiseFeeder(dims, c["S"], 777) as nf:
        # (i) inner_elbo + grads
        L = obj.inner_elbo(model=obj.model)
        gs = torch.autograd.grad(L, params + [obj.u] + ([obj.v] if obj.learn_v else []))
        res["inner_val"] = L.item()
        res["inner_gparams"] = torch.cat([g.reshape(-1) for g in gs[:len(params)]]).double().numpy()
        res["inner_gu"] = gs[len(params)].double().numpy()
        if obj.learn_v:
            res["inner_gv"] = gs[len(params) + 1].double().numpy()
        # (ii) psvi_elbo + grads
        L = obj.psvi_elbo(xb, yb, model=obj.model)
        gs = torch.autograd.grad(L, params + [obj.u] + ([obj.v] if obj.learn_v else []))
        res["outer_val"] = L.item()
        res["outer_gparams"] = torch.cat([g.reshape(-1) for g in gs[:len(params)]]).double().numpy()
        res["outer_gu"] = gs[len(params)].double().numpy()
        if obj.learn_v:
            res["outer_gv"] = gs[len(params) + 1].double().numpy()
        # (iii) nested_step
        obj.elbos = []
        loss = obj.nested_step(xb, yb)
        res["nested_loss"] = loss.item()
        res["nested_gu"] = obj.u.grad.double().numpy().copy()
        if obj.learn_v:
            res["nested_gv"] = obj.v.grad.double().numpy().copy()
        res["nested_params"] = torch.nn.utils.parameters_to_vector(obj.model.parameters()).detach().double().numpy()
        res["nested_u_after"] = obj.u.detach().double().numpy().copy()
        res["nested_v_after"] = obj.v.detach().double().numpy().copy()
        res["nested_elbos"] = np.array([e[1] for e in obj.elbos], dtype=np.float64)
        # (iv) evaluate with the post-step state
        acc, nll, went, ness, vent = obj.evaluate()
        res["eval"] = np.array([acc.item(), nll.item(), went.item(), ness.item(), vent.item()])
        out["n_forwards"] = len(nf.history)
    return out, res


def run_mfvi_case():
    """run_mfvi_subset (baselines.py:923-1062) on halfmoon, fn H=40, M=20, S=8, 6 iterations."""
    x, y, xt, yt, N, D, tr, te, nc = get_data("halfmoon")
    M, S, H = 20, 8, 40
    with contextlib.redirect_stdout(io.StringIO()), contextlib.redirect_stderr(io.StringIO()):
        xs, ys = pseudo_subsample_init(x, y, num_pseudo=M, seed=3, nc=nc)
        # initial parameters: same seeding as run_mfvi_subset does before set_up_model
        import random
        random.seed(3), np.random.seed(3), torch.manual_seed(3)
        from psvi.experiments.experiments_utils import set_up_model
        net0 = set_up_model(architecture="fn", D=D, n_hidden=H, nc=nc, mc_samples=S, init_sd=1e-3)
        mu0, rho0 = get_mu_rho(net0)
        # NB experiments_utils.set_up_model does not forward n_layers, so "fn" here has make_fcnet's default of
        # TWO hidden layers (neural_net.py:271) -- unlike PSVI.set_up_model (psvi_classes.py:707-717).
        dims = model_dims(net0)
        with NoiseFeeder(dims, S, 4242) as nf:
            res = run_mfvi_subset(x=x, y=y, xt=xt, yt=yt, mc_samples=S, data_minibatch=256, num_epochs=3, log_every=2,
                                  D=D, lr0net=1e-3, seed=3, train_dataset=tr, test_dataset=te, num_pseudo=M,
                                  init_args="subsample", architecture="fn", n_hidden=H, nc=nc, dnm="halfmoon",
                                  init_sd=1e-3)
            nfw = len(nf.history)
    return dict(dims=np.array(dims), N=N, S=S, M=M, noise_seed=4242, n_forwards=nfw, lr0net=1e-3,
                mu0=mu0, rho0=rho0, xs=xs.detach().double().numpy(), ys=ys.double().numpy(),
                xt=xt.double().numpy(), yt=yt.double().numpy(),
                ref_elbos=np.array(res["elbos"]), ref_accs=np.array(res["accs"]), ref_nlls=np.array(res["nlls"]))


def run_fn2_case(name="fn2_hm_h6", H=6, M=8, S=5, T=3, B=32, init_sd=0.05, lr0net=1e-3):
    """fn2 (full covariance, reference neural_net.py:408-524) in fp64 -- the reference is numerically unstable in fp32 and
    at small init_sd, so the golden is taken in the well-conditioned regime."""
    x, y, xt, yt, N, D, tr, te, nc = get_data("halfmoon")
    kw = dict(mc_samples=S, num_epochs=0, data_minibatch=B, D=D, N=N, inner_it=T, trainer="nested", log_every=10,
              lr0u=1e-4, lr0net=lr0net, lr0v=1e-3, init_args="subsample", init_sd=init_sd, num_pseudo=M, seed=0,
              architecture="fn2", n_hidden=H, n_layers=1, logistic_regression=False, train_dataset=tr, test_dataset=te,
              dnm="halfmoon", nc=nc, data_folder="/tmp/psvi_data", compute_weights_entropy=True, register_elbos=True)
    with contextlib.redirect_stdout(io.StringIO()), contextlib.redirect_stderr(io.StringIO()), np.errstate(all="ignore"):
        import warnings
        warnings.simplefilter("ignore")
        obj = PSVILearnV(**kw)
        obj.run_psvi(**kw)
    tdt = torch.float64
    obj.model.to(tdt)
    rng = np.random.default_rng(99)
    # non-trivial state: random means and correlations (the reference initialises both to zero)
    with torch.no_grad():
        for n_, p_ in obj.model.named_parameters():
            if n_.endswith("mean"):
                p_.copy_(torch.tensor(0.3 * rng.standard_normal(p_.shape)))
            if n_.endswith("_corr"):
                p_.copy_(torch.tensor(0.01 * rng.standard_normal(p_.shape)))
    v0 = (0.3 * rng.standard_normal(M)).astype(np.float32)
    obj.v = torch.tensor(v0, dtype=tdt).requires_grad_(True)
    obj.u = obj.u.detach().to(tdt).requires_grad_(True)
    obj.z = obj.z.to(tdt)
    obj.optim_net = torch.optim.Adam(list(obj.model.parameters()), lr0net)
    obj.optim_u = torch.optim.Adam([obj.u], 1e-4)
    obj.optim_v = torch.optim.Adam([obj.v], 1e-3)
    obj.scheduler_optim_net = None
    xb, yb = x[:B].to(tdt), y[:B].to(tdt)
    dims = [D, H, H, nc]     # make_fc2net default n_layers=2 (Q7)
    params = list(obj.model.parameters())
    phi0 = torch.nn.utils.parameters_to_vector(params).detach().numpy().copy()
    out = dict(dims=np.array(dims), N=N, S=S, T=T, M=M, B=B, lr0net=lr0net, noise_seed=555, vmode=1, phi0=phi0,
               u0=obj.u.detach().numpy().copy(), z=obj.z.numpy().copy(), v0=v0.astype(np.float64),
               xb=xb.numpy().copy(), yb=yb.numpy().copy(), xt=xt.double().numpy().copy(), yt=yt.double().numpy().copy())
    te.data = te.data.double()
    with NoiseFeeder(dims, S, 555, fullcov=True) as nf:
        L = obj.inner_elbo(model=obj.model)
        gs = torch.autograd.grad(L, params + [obj.u, obj.v])
        out["ref64_inner_val"] = L.item()
        out["ref64_inner_gparams"] = torch.cat([g.reshape(-1) for g in gs[:len(params)]]).numpy()
        out["ref64_inner_gu"], out["ref64_inner_gv"] = gs[-2].numpy(), gs[-1].numpy()
        L = obj.psvi_elbo(xb, yb, model=obj.model)
        gs = torch.autograd.grad(L, params + [obj.u, obj.v])
        out["ref64_outer_val"] = L.item()
        out["ref64_outer_gparams"] = torch.cat([g.reshape(-1) for g in gs[:len(params)]]).numpy()
        out["ref64_outer_gu"], out["ref64_outer_gv"] = gs[-2].numpy(), gs[-1].numpy()
        loss = obj.nested_step(xb, yb)
        out["ref64_nested_loss"] = loss.item()
        out["ref64_nested_gu"], out["ref64_nested_gv"] = obj.u.grad.numpy().copy(), obj.v.grad.numpy().copy()
        out["ref64_nested_params"] = torch.nn.utils.parameters_to_vector(obj.model.parameters()).detach().numpy().copy()
        out["ref64_nested_u_after"], out["ref64_nested_v_after"] = obj.u.detach().numpy().copy(), obj.v.detach().numpy().copy()
        acc, nll, went, ness, vent = obj.evaluate()
        out["ref64_eval"] = np.array([acc.item(), nll.item(), went.item(), ness.item(), vent.item()])
        out["n_forwards"] = len(nf.history)
    pth = os.path.join(ROOT, "tests", "golden", name + ".npz")
    np.savez_compressed(pth, **out)
    print(name, "forwards", out["n_forwards"], "nested_loss", out["ref64_nested_loss"], "size", os.path.getsize(pth))


def run_hyper_case(name="hyper_fn_hm", H=20, M=10, S=6, T=6, B=64, K=4, init_sd=1e-2, lr0net=1e-3):
    """PSVI.hyper_step (psvi_classes.py:602-687) with hypergrad.CG_normaleq (hypergradients.py:199-244), fp64."""
    x, y, xt, yt, N, D, tr, te, nc = get_data("halfmoon")
    kw = dict(mc_samples=S, num_epochs=0, data_minibatch=B, D=D, N=N, inner_it=T, trainer="hyper", log_every=10,
              lr0u=1e-4, lr0net=lr0net, lr0v=1e-3, init_args="subsample", init_sd=init_sd, num_pseudo=M, seed=0,
              architecture="fn", n_hidden=H, n_layers=1, logistic_regression=False, train_dataset=tr, test_dataset=te,
              dnm="halfmoon", nc=nc, data_folder="/tmp/psvi_data", compute_weights_entropy=True, register_elbos=False)
    with contextlib.redirect_stdout(io.StringIO()), contextlib.redirect_stderr(io.StringIO()), np.errstate(all="ignore"):
        import warnings
        warnings.simplefilter("ignore")
        obj = PSVILearnV(**kw)
        obj.run_psvi(**kw)
    tdt = torch.float64
    obj.model.to(tdt)
    rng = np.random.default_rng(5)
    v0 = (0.3 * rng.standard_normal(M)).astype(np.float32)
    obj.v = torch.tensor(v0, dtype=tdt).requires_grad_(True)
    obj.u = obj.u.detach().to(tdt).requires_grad_(True)
    obj.z = obj.z.to(tdt)
    obj.optim_net = torch.optim.Adam(list(obj.model.parameters()), lr0net)
    obj.optim_u = torch.optim.Adam([obj.u], 1e-4)
    obj.optim_v = torch.optim.Adam([obj.v], 1e-3)
    obj.scheduler_optim_net = None
    xb, yb = x[:B].to(tdt), y[:B].to(tdt)
    dims = model_dims(obj.model)
    mu0, rho0 = get_mu_rho(obj.model)
    out = dict(dims=np.array(dims), N=N, S=S, T=T, M=M, B=B, K=K, lr0net=lr0net, linsys_lr=1e-2, noise_seed=4321, vmode=1,
               mu0=mu0, rho0=rho0, u0=obj.u.detach().numpy().copy(), z=obj.z.numpy().copy(), v0=v0.astype(np.float64),
               xb=xb.numpy().copy(), yb=yb.numpy().copy())
    with NoiseFeeder(dims, S, 4321) as nf:
        ll = obj.hyper_step(xb, yb, K=K, linsys_lr=1e-2)
        out["ref64_ll"] = ll
        out["ref64_gu"], out["ref64_gv"] = obj.u.grad.numpy().copy(), obj.v.grad.numpy().copy()
        out["ref64_params"] = torch.nn.utils.parameters_to_vector(obj.model.parameters()).detach().numpy().copy()
        out["ref64_u_after"], out["ref64_v_after"] = obj.u.detach().numpy().copy(), obj.v.detach().numpy().copy()
        out["n_forwards"] = len(nf.history)
    pth = os.path.join(ROOT, "tests", "golden", name + ".npz")
    np.savez_compressed(pth, **out)
    print(name, "forwards", out["n_forwards"], "ll", ll, "|gu|", np.abs(out["ref64_gu"]).max(), "size", os.path.getsize(pth))


from tests.fake_mnist import FakeMNIST  # noqa: E402


def run_lenet_case(name="lenet_m10", M=10, S=3, T=2, B=6, init_sd=1e-2, lr0net=1e-3, n_test=24):
    """lenet (VIConv2d / BatchMaxPool2d / make_lenet, neural_net.py:194-255,334-359) in fp64 through PSVILearnV."""
    from oracle.ref_import import LeNetNoiseFeeder
    tr, te = FakeMNIST(64, 0), FakeMNIST(n_test, 1)
    N, D, nc = len(tr), 784, 10
    kw = dict(mc_samples=S, num_epochs=0, data_minibatch=B, D=D, N=N, inner_it=T, trainer="nested", log_every=10,
              lr0u=1e-4, lr0net=lr0net, lr0v=1e-3, init_args="subsample", init_sd=init_sd, num_pseudo=M, seed=0,
              architecture="lenet", n_hidden=0, n_layers=1, logistic_regression=False, train_dataset=tr, test_dataset=te,
              dnm="MNIST", nc=nc, data_folder="/tmp/psvi_data", compute_weights_entropy=True, register_elbos=True)
    with contextlib.redirect_stdout(io.StringIO()), contextlib.redirect_stderr(io.StringIO()), np.errstate(all="ignore"):
        import warnings
        warnings.simplefilter("ignore")
        obj = PSVILearnV(**kw)
        obj.run_psvi(**kw)
    tdt = torch.float64
    obj.model.to(tdt)
    rng = np.random.default_rng(11)
    v0 = (0.3 * rng.standard_normal(M)).astype(np.float32)
    obj.v = torch.tensor(v0, dtype=tdt).requires_grad_(True)
    obj.u = obj.u.detach().to(tdt).requires_grad_(True)
    obj.z = obj.z.to(tdt)
    obj.optim_net = torch.optim.Adam(list(obj.model.parameters()), lr0net)
    obj.optim_u = torch.optim.Adam([obj.u], 1e-4)
    obj.optim_v = torch.optim.Adam([obj.v], 1e-3)
    obj.scheduler_optim_net = None
    xb = torch.stack([tr[i][0] for i in range(B)]).to(tdt)
    yb = torch.tensor([tr[i][1] for i in range(B)]).to(tdt)
    xt = torch.stack([te[i][0] for i in range(n_test)]).to(tdt)
    yt = torch.tensor([te[i][1] for i in range(n_test)])
    params = list(obj.model.parameters())
    phi0 = torch.nn.utils.parameters_to_vector(params).detach().numpy().copy()
    out = dict(N=N, S=S, T=T, M=M, B=B, lr0net=lr0net, noise_seed=909, vmode=1, phi0=phi0,
               u0=obj.u.detach().numpy().reshape(M, 784).copy(), z=obj.z.numpy().copy(), v0=v0.astype(np.float64),
               xb=xb.numpy().reshape(B, 784).copy(), yb=yb.numpy().copy(), xt=xt.numpy().reshape(n_test, 784).copy(),
               yt=yt.double().numpy().copy())
    # evaluate() iterates the reference's own test loader: batches of B rows in dataset order (shuffle=False)
    with LeNetNoiseFeeder(S, 909) as nf:
        L = obj.inner_elbo(model=obj.model)
        gs = torch.autograd.grad(L, params + [obj.u, obj.v])
        out["ref64_inner_val"] = L.item()
        out["ref64_inner_gparams"] = torch.cat([g.reshape(-1) for g in gs[:len(params)]]).numpy()
        out["ref64_inner_gu"], out["ref64_inner_gv"] = gs[-2].numpy().reshape(M, 784), gs[-1].numpy()
        L = obj.psvi_elbo(xb, yb, model=obj.model)
        gs = torch.autograd.grad(L, params + [obj.u, obj.v])
        out["ref64_outer_val"] = L.item()
        out["ref64_outer_gparams"] = torch.cat([g.reshape(-1) for g in gs[:len(params)]]).numpy()
        out["ref64_outer_gu"], out["ref64_outer_gv"] = gs[-2].numpy().reshape(M, 784), gs[-1].numpy()
        loss = obj.nested_step(xb, yb)
        out["ref64_nested_loss"] = loss.item()
        out["ref64_nested_gu"], out["ref64_nested_gv"] = obj.u.grad.numpy().reshape(M, 784).copy(), obj.v.grad.numpy().copy()
        out["ref64_nested_params"] = torch.nn.utils.parameters_to_vector(obj.model.parameters()).detach().numpy().copy()
        out["ref64_nested_u_after"] = obj.u.detach().numpy().reshape(M, 784).copy()
        out["ref64_nested_v_after"] = obj.v.detach().numpy().copy()
        obj.test_loader = [(xt[i:i + B], yt[i:i + B]) for i in range(0, n_test, B)]
        acc, nll, went, ness, vent = obj.evaluate()
        out["ref64_eval"] = np.array([acc.item(), nll.item(), went.item(), ness.item(), vent.item()])
        out["n_forwards"] = len(nf.history)
    pth = os.path.join(ROOT, "tests", "golden", name + ".npz")
    np.savez_compressed(pth, **out)
    print(name, "forwards", out["n_forwards"], "inner", out["ref64_inner_val"], "nested_loss", out["ref64_nested_loss"], "eval",
          out["ref64_eval"], "size", os.path.getsize(pth))


def run_ablated_case(name, cls_name, S, H=30, M=10, T=4, B=32, init_sd=1e-2, lr0net=1e-3):
    """PSVI_Ablated / PSVI_No_IW (psvi_classes.py:1388-1472): outer objective without importance weights; fp64."""
    import psvi.inference.psvi_classes as rc
    x, y, xt, yt, N, D, tr, te, nc = get_data("halfmoon")
    kw = dict(mc_samples=S, num_epochs=0, data_minibatch=B, D=D, N=N, inner_it=T, trainer="nested", log_every=10,
              lr0u=1e-4, lr0net=lr0net, lr0v=1e-3, init_args="subsample", init_sd=init_sd, num_pseudo=M, seed=0,
              architecture="fn", n_hidden=H, n_layers=1, logistic_regression=False, train_dataset=tr, test_dataset=te,
              dnm="halfmoon", nc=nc, data_folder="/tmp/psvi_data", compute_weights_entropy=True, register_elbos=False)
    with contextlib.redirect_stdout(io.StringIO()), contextlib.redirect_stderr(io.StringIO()), np.errstate(all="ignore"):
        import warnings
        warnings.simplefilter("ignore")
        obj = getattr(rc, cls_name)(**kw)
        obj.run_psvi(**kw)
    S = int(obj.mc_samples)
    tdt = torch.float64
    obj.model.to(tdt)
    rng = np.random.default_rng(21)
    v0 = (0.3 * rng.standard_normal(M)).astype(np.float32)
    obj.v = torch.tensor(v0, dtype=tdt).requires_grad_(True)
    obj.u = obj.u.detach().to(tdt).requires_grad_(True)
    obj.z = obj.z.to(tdt)
    obj.optim_net = torch.optim.Adam(list(obj.model.parameters()), lr0net)
    obj.optim_u = torch.optim.Adam([obj.u], 1e-4)
    obj.optim_v = torch.optim.Adam([obj.v], 1e-3)
    obj.scheduler_optim_net = None
    xb, yb = x[:B].to(tdt), y[:B].to(tdt)
    dims = model_dims(obj.model)
    mu0, rho0 = get_mu_rho(obj.model)
    params = list(obj.model.parameters())
    out = dict(dims=np.array(dims), N=N, S=S, T=T, M=M, B=B, lr0net=lr0net, noise_seed=8080, vmode=1, mu0=mu0, rho0=rho0,
               u0=obj.u.detach().numpy().copy(), z=obj.z.numpy().copy(), v0=v0.astype(np.float64), xb=xb.numpy().copy(),
               yb=yb.numpy().copy())
    with NoiseFeeder(dims, S, 8080) as nf:
        L = obj.psvi_elbo(xb, yb, model=obj.model)
        gs = torch.autograd.grad(L, params)
        out["ref64_outer_val"] = L.item()
        out["ref64_outer_gparams"] = torch.cat([g.reshape(-1) for g in gs]).numpy()
        loss = obj.nested_step(xb, yb)
        out["ref64_nested_loss"] = loss.item()
        out["ref64_nested_gu"], out["ref64_nested_gv"] = obj.u.grad.numpy().copy(), obj.v.grad.numpy().copy()
        out["ref64_nested_params"] = torch.nn.utils.parameters_to_vector(obj.model.parameters()).detach().numpy().copy()
        out["n_forwards"] = len(nf.history)
    pth = os.path.join(ROOT, "tests", "golden", name + ".npz")
    np.savez_compressed(pth, **out)
    print(name, "S", S, "forwards", out["n_forwards"], "outer", out["ref64_outer_val"], "nested_loss", out["ref64_nested_loss"],
          "|gu|", np.abs(out["ref64_nested_gu"]).max(), "size", os.path.getsize(pth))


def run_grid_case(name="grid_fn_hm", H=20, M=10, S=6, n=12, init_sd=5e-2):
    """PSVI.pred_on_grid (psvi_classes.py:1130-1175): importance-weighted predictive probabilities over a 2-d grid, fp64."""
    x, y, xt, yt, N, D, tr, te, nc = get_data("halfmoon")
    kw = dict(mc_samples=S, num_epochs=0, data_minibatch=32, D=D, N=N, inner_it=2, trainer="nested", log_every=10, lr0u=1e-4,
              lr0net=1e-3, lr0v=1e-3, init_args="subsample", init_sd=init_sd, num_pseudo=M, seed=0, architecture="fn", n_hidden=H,
              n_layers=1, logistic_regression=False, train_dataset=tr, test_dataset=te, dnm="halfmoon", nc=nc,
              data_folder="/tmp/psvi_data", compute_weights_entropy=True, register_elbos=False)
    with contextlib.redirect_stdout(io.StringIO()), contextlib.redirect_stderr(io.StringIO()), np.errstate(all="ignore"):
        import warnings
        warnings.simplefilter("ignore")
        obj = PSVILearnV(**kw)
        obj.run_psvi(**kw)
    obj.model.to(torch.float64)
    rng = np.random.default_rng(31)
    v0 = (0.3 * rng.standard_normal(M)).astype(np.float32)
    obj.v = torch.tensor(v0, dtype=torch.float64)
    obj.u = obj.u.detach().to(torch.float64)
    obj.z = obj.z.to(torch.float64)
    obj.N = 30.0     # keep the importance weights away from a one-hot vector so that the mixture is actually exercised
    dims = model_dims(obj.model)
    mu0, rho0 = get_mu_rho(obj.model)
    out = dict(dims=np.array(dims), N=obj.N, S=S, M=M, n=n, noise_seed=6060, vmode=1, mu0=mu0, rho0=rho0,
               u0=obj.u.numpy().copy(), z=obj.z.numpy().copy(), v0=v0.astype(np.float64))
    _orig_linspace = torch.linspace
    torch.linspace = lambda *a, **k: _orig_linspace(*a, **k).double()
    try:
        with NoiseFeeder(dims, S, 6060) as nf:
            out["ref64_grid_iw"] = obj.pred_on_grid(n_test_per_dim=n, correction=True).numpy().copy()
            out["ref64_grid_mean"] = obj.pred_on_grid(n_test_per_dim=n, correction=False).numpy().copy()
            out["n_forwards"] = len(nf.history)
    finally:
        torch.linspace = _orig_linspace
    pth = os.path.join(ROOT, "tests", "golden", name + ".npz")
    np.savez_compressed(pth, **out)
    print(name, out["ref64_grid_iw"].shape, out["ref64_grid_iw"][:2], "size", os.path.getsize(pth))


def main():
    os.makedirs(os.path.join(ROOT, "tests", "golden"), exist_ok=True)
    for c in CASES:
        out32, r32 = run_case(c, "32")
        out64, r64 = run_case(c, "64")
        for k in ("mu0", "rho0", "u0", "xb"):
            assert np.array_equal(out32[k], out64[k]), k
        blob = dict(out32)
        blob.update({"ref32_" + k: v for k, v in r32.items()})
        blob.update({"ref64_" + k: v for k, v in r64.items()})
        p = os.path.join(ROOT, "tests", "golden", c["name"] + ".npz")
        np.savez_compressed(p, **blob)
        print(c["name"], "forwards", out32["n_forwards"], "nested_loss32/64", r32["nested_loss"], r64["nested_loss"],
              "size", os.path.getsize(p))
    run_fn2_case()
    run_ablated_case("ablated_fn_hm", "PSVI_Ablated", S=5)
    run_ablated_case("noiw_fn_hm", "PSVI_No_IW", S=5)      # PSVI_No_IW forces mc_samples = 1 for training
    run_lenet_case()
    run_grid_case()
    run_hyper_case()
    blob = run_mfvi_case()
    p = os.path.join(ROOT, "tests", "golden", "mfvi_subset_hm.npz")
    np.savez_compressed(p, **blob)
    print("mfvi_subset_hm", blob["ref_elbos"], blob["ref_accs"], blob["ref_nlls"])


if __name__ == "__main__":
    main()
