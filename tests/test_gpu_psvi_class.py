"""GPU parity tests at the drop-in Python boundary (psvi.inference.psvi_classes / baselines / flow_psvi): the same
calls a user of the reference makes, with the golden noise stream injected, against the reference's own outputs."""
import os
import pickle

import numpy as np
import pytest
import torch

from oracle import psvi_oracle as po
from tests.gpu_util import CASES, GOLDEN, load, rel_l2

pytestmark = pytest.mark.gpu

DNM = {"hm": "halfmoon", "fb": "four_blobs"}


def make_obj(name, g, dims, S, T, eps, cls=None, init_sd=1e-3):
    from psvi.experiments.experiments_utils import read_dataset
    from psvi.inference.psvi_classes import PSVI, ExternalNoise, PSVILearnV
    dnm = DNM[name.split("_")[1]]
    torch.manual_seed(0)
    x, y, xt, yt, N, D, tr, te, nc = read_dataset(dnm, {"test_ratio": 0.2})
    L = len(dims) - 1
    arch = "logistic_regression" if L == 1 else "fn"
    kw = dict(mc_samples=S, num_epochs=0, data_minibatch=int(g["B"]), D=D, N=N, inner_it=T, trainer="nested",
              log_every=10, lr0u=1e-4, lr0net=float(g["lr0net"]), lr0v=1e-3, init_args="subsample", init_sd=init_sd,
              num_pseudo=int(g["M"]), seed=0, architecture=arch, n_hidden=dims[1] if L > 1 else 0, n_layers=L - 1,
              logistic_regression=(arch == "logistic_regression"), train_dataset=tr, test_dataset=te, dnm=dnm, nc=nc,
              compute_weights_entropy=True, register_elbos=True, quiet=True)
    Cls = cls if cls is not None else (PSVILearnV if int(g["vmode"]) == 1 else PSVI)
    obj = Cls(**kw)
    obj.run_psvi(**kw)
    if "xt" in g:
        np.testing.assert_allclose(te.data.numpy(), g["xt"], atol=1e-6)   # same generated dataset as the golden run
    # inject the golden state
    mu, rho = obj.model.flat()
    mu.copy_(torch.as_tensor(g["mu0"])), rho.copy_(torch.as_tensor(g["rho0"]))
    with torch.no_grad():
        obj.u.copy_(torch.as_tensor(g["u0"]))
        obj.v.copy_(torch.as_tensor(g["v0"]))
    obj.z = torch.as_tensor(g["z"]).float().cuda()
    obj.scheduler_optim_net = None
    obj.noise_source = ExternalNoise(eps)
    return obj


@pytest.mark.parametrize("name", CASES)
def test_psvi_methods_match_reference(name):
    g, dims, S, T, eps = load(name)
    obj = make_obj(name, g, dims, S, T, eps)
    sd_small = "sd1e-6" in name
    tol = 5e-3 if sd_small else 1e-3
    xb, yb = torch.as_tensor(g["xb"]).float().cuda(), torch.as_tensor(g["yb"]).cuda()
    # state_dict layout == reference parameters_to_vector order
    vec = torch.nn.utils.parameters_to_vector(obj.model.parameters()).detach().cpu().numpy()
    np.testing.assert_allclose(vec, po.mu_rho_to_phi(g["mu0"], g["rho0"], dims), rtol=1e-6)
    assert abs(obj.inner_elbo(model=obj.model).item() - g["ref32_inner_val"]) <= 1e-4 * abs(g["ref32_inner_val"])
    assert abs(obj.psvi_elbo(xb, yb, model=obj.model).item() - g["ref32_outer_val"]) <= 1e-4 * abs(g["ref32_outer_val"])
    obj.elbos = []
    loss = obj.nested_step(xb, yb)
    assert abs(loss.item() - g["ref32_nested_loss"]) <= 2e-4 * abs(g["ref32_nested_loss"])
    assert rel_l2(obj.u.grad.cpu().numpy(), g["ref64_nested_gu"]) < tol
    if obj.learn_v:
        assert rel_l2(obj.v.grad.cpu().numpy(), g["ref64_nested_gv"]) < tol
        np.testing.assert_allclose(obj.v.detach().cpu().numpy(), g["ref32_nested_v_after"], atol=2e-6)
    np.testing.assert_allclose(obj.u.detach().cpu().numpy(), g["ref32_nested_u_after"], atol=2e-6)
    vec = torch.nn.utils.parameters_to_vector(obj.model.parameters()).detach().cpu().numpy()
    assert rel_l2(vec, g["ref32_nested_params"]) < 1e-5
    np.testing.assert_allclose([e[1] for e in obj.elbos], g["ref32_nested_elbos"], rtol=2e-4)
    assert [e[0] for e in obj.elbos] == [1] * len(range(0, T, 10)) + [0]
    acc, nll, went, ness, vent = obj.evaluate()
    # diagnostics are compared with the reference run in fp64 (same noise): at init_sd=1e-6 the fp32 reference itself
    # is 9% off on the importance-weight entropy (ref32 0.003774 vs ref64 0.003442), the CUDA path is not
    ref = g["ref64_eval"]
    assert abs(acc.item() - ref[0]) <= 1.0 / len(g["yt"]) + 1e-6
    np.testing.assert_allclose([nll.item(), ness.item(), vent.item()], [ref[1], ref[3], ref[4]], rtol=3e-3)
    np.testing.assert_allclose(went.item(), ref[2], rtol=5e-3, atol=1e-4)


def test_run_mfvi_subset_matches_reference_trace():
    from oracle.ref_import import NoiseFeeder
    from psvi.experiments.experiments_utils import read_dataset
    from psvi.inference.baselines import run_mfvi_subset
    from psvi.inference.psvi_classes import ExternalNoise
    g = dict(np.load(os.path.join(GOLDEN, "mfvi_subset_hm.npz")))
    dims = [int(d) for d in g["dims"]]
    S, M = int(g["S"]), int(g["M"])
    eps = NoiseFeeder.stream(dims, S, int(g["noise_seed"]), int(g["n_forwards"]))
    x, y, xt, yt, N, D, tr, te, nc = read_dataset("halfmoon", {"test_ratio": 0.2})
    res = run_mfvi_subset(x=x, y=y, xt=xt, yt=yt, mc_samples=S, data_minibatch=256, num_epochs=3, log_every=2, D=D,
                          lr0net=1e-3, seed=3, train_dataset=tr, test_dataset=te, num_pseudo=M, init_args="subsample",
                          architecture="fn", n_hidden=dims[1], nc=nc, dnm="halfmoon", init_sd=1e-3,
                          noise_source=ExternalNoise(eps), quiet=True)
    np.testing.assert_allclose(res["elbos"], g["ref_elbos"], rtol=5e-5)
    np.testing.assert_allclose(res["accs"], g["ref_accs"], atol=1e-6)
    np.testing.assert_allclose(res["nlls"], g["ref_nlls"], rtol=5e-5)
    assert res["csizes"] == [M] * 6


def test_run_psvi_end_to_end_learns_halfmoon():
    """cfg-2-like run with the in-kernel Philox noise: results dict has the reference's keys and accuracy improves."""
    from psvi.experiments.experiments_utils import read_dataset
    from psvi.inference.psvi_classes import PSVILearnV
    x, y, xt, yt, N, D, tr, te, nc = read_dataset("halfmoon", {"test_ratio": 0.2})
    kw = dict(mc_samples=10, num_epochs=61, data_minibatch=128, D=D, N=N, inner_it=20, trainer="nested", log_every=30,
              lr0u=1e-3, lr0net=1e-2, lr0v=1e-2, init_args="subsample", init_sd=1e-3, num_pseudo=20, seed=0,
              architecture="fn", n_hidden=100, n_layers=1, logistic_regression=False, train_dataset=tr,
              test_dataset=te, dnm="halfmoon", nc=nc, compute_weights_entropy=True, register_elbos=False, quiet=True)
    res = PSVILearnV(**kw).run_psvi(**kw)
    assert set(res) >= {"accs", "nlls", "csizes", "times", "elbos", "went", "ness", "vent", "vs", "avg_epoch_time",
                        "gpu_memory", "chosen_indices"}
    assert len(res["accs"]) == 3 and all(np.isfinite(res["nlls"]))
    assert res["accs"][-1] >= 0.8 and res["accs"][-1] > res["accs"][0]


def test_hyper_trainer_runs_and_reduces_loss():
    from psvi.experiments.experiments_utils import read_dataset
    from psvi.inference.psvi_classes import PSVILearnV
    x, y, xt, yt, N, D, tr, te, nc = read_dataset("halfmoon", {"test_ratio": 0.2})
    kw = dict(mc_samples=8, num_epochs=4, data_minibatch=128, D=D, N=N, inner_it=10, trainer="hyper", log_every=2,
              lr0u=1e-4, lr0net=1e-3, lr0v=1e-3, init_args="subsample", init_sd=1e-3, num_pseudo=10, seed=1,
              architecture="fn", n_hidden=20, n_layers=1, logistic_regression=False, train_dataset=tr, test_dataset=te,
              dnm="halfmoon", nc=nc, compute_weights_entropy=False, register_elbos=False, quiet=True)
    res = PSVILearnV(**kw).run_psvi(**kw)
    assert len(res["accs"]) == 2 and all(np.isfinite(res["nlls"]))


def test_flow_psvi_cli_writes_results(tmp_path):
    from psvi.experiments import flow_psvi
    out = flow_psvi.main(["--datasets", "halfmoon", "--architecture", "logistic_regression", "--methods",
                          "psvi_learn_v", "mfvi_subset", "--coreset_sizes", "10", "--num_epochs", "5", "--inner_it",
                          "5", "--num_trials", "1", "--log_every", "2", "--results_folder", str(tmp_path),
                          "--data_folder", str(tmp_path), "--fnm", "r"])
    with open(tmp_path / "r.pk", "rb") as f:
        res = pickle.load(f)
    assert set(res["halfmoon"]) == {"psvi_learn_v", "mfvi_subset"}
    r = res["halfmoon"]["psvi_learn_v"][10][0]
    assert len(r["accs"]) == 3 and np.isfinite(r["nlls"]).all()
    assert os.path.exists(tmp_path / "r.json")


@pytest.mark.parametrize("stream", [False, True])
def test_hyper_step_matches_reference(stream):
    """--trainer hyper: PSVI.hyper_step + CG_normaleq (reference psvi_classes.py:602-687, hypergradients.py:199-244) with
    the reference's noise-consumption order (T inner, outer, w_mapped, 2 per JVP, final outer); through the fused engine
    and through the streaming engine (the path fn2 / lenet / large models take)."""
    from oracle.ref_import import NoiseFeeder
    from psvi.experiments.experiments_utils import read_dataset
    from psvi.inference.psvi_classes import ExternalNoise, PSVILearnV
    g = dict(np.load(os.path.join(GOLDEN, "hyper_fn_hm.npz")))
    dims = [int(d) for d in g["dims"]]
    S, T, M, B, K = int(g["S"]), int(g["T"]), int(g["M"]), int(g["B"]), int(g["K"])
    eps = NoiseFeeder.stream(dims, S, int(g["noise_seed"]), int(g["n_forwards"]))
    x, y, xt, yt, N, D, tr, te, nc = read_dataset("halfmoon", {"test_ratio": 0.2})
    kw = dict(mc_samples=S, num_epochs=0, data_minibatch=B, D=D, N=N, inner_it=T, trainer="hyper", log_every=10, lr0u=1e-4,
              lr0net=float(g["lr0net"]), lr0v=1e-3, init_args="subsample", init_sd=1e-2, num_pseudo=M, seed=0,
              architecture="fn", n_hidden=dims[1], n_layers=1, logistic_regression=False, train_dataset=tr, test_dataset=te,
              dnm="halfmoon", nc=nc, compute_weights_entropy=True, register_elbos=False, quiet=True)
    obj = PSVILearnV(**kw)
    obj.run_psvi(**kw)
    mu, rho = obj.model.flat()
    mu.copy_(torch.as_tensor(g["mu0"])), rho.copy_(torch.as_tensor(g["rho0"]))
    with torch.no_grad():
        obj.u.copy_(torch.as_tensor(g["u0"]))
        obj.v.copy_(torch.as_tensor(g["v0"]))
    obj.z = torch.as_tensor(g["z"]).float().cuda()
    obj.scheduler_optim_net = None
    obj.noise_source = ExternalNoise(eps)
    if stream:
        obj._ws[("force_stream", id(obj.model))] = True
    xb, yb = torch.as_tensor(g["xb"]).float().cuda(), torch.as_tensor(g["yb"]).cuda()
    ll = obj.hyper_step(xb, yb, K=K, linsys_lr=float(g["linsys_lr"]))
    assert obj.noise_source.pos == int(g["n_forwards"])            # same number of forwards as the reference
    assert abs(ll - g["ref64_ll"]) <= 2e-4 * abs(g["ref64_ll"])
    assert rel_l2(obj.u.grad.cpu().numpy(), g["ref64_gu"]) < 2e-3
    assert rel_l2(obj.v.grad.cpu().numpy(), g["ref64_gv"]) < 2e-3
    vec = torch.nn.utils.parameters_to_vector(obj.model.parameters()).detach().cpu().numpy()
    assert rel_l2(vec, g["ref64_params"]) < 1e-5
    np.testing.assert_allclose(obj.u.detach().cpu().numpy(), g["ref64_u_after"], atol=2e-6)


@pytest.mark.parametrize("name", ["ablated_fn_hm", "noiw_fn_hm"])
def test_ablated_and_no_iw_match_reference(name):
    """PSVI_Ablated / PSVI_No_IW (reference psvi_classes.py:1388-1472) through the streaming engine against the reference's
    fp64 run: outer objective without importance weights; PSVI_No_IW trains with ONE sample, which makes the reference's
    inner_elbo score every pseudo-point against every label (reproduced).  Hypergradients rel-L2 2e-3."""
    import os
    from oracle.ref_import import NoiseFeeder
    from psvi.inference.psvi_classes import PSVI_Ablated, PSVI_No_IW
    g = dict(np.load(os.path.join(GOLDEN, name + ".npz")))
    dims = [int(d) for d in g["dims"]]
    S, T = int(g["S"]), int(g["T"])
    eps = NoiseFeeder.stream(dims, S, int(g["noise_seed"]), int(g["n_forwards"]))
    cls = PSVI_No_IW if name.startswith("noiw") else PSVI_Ablated
    obj = make_obj("x_hm", g, dims, 5, T, eps, cls=cls, init_sd=1e-2)
    assert obj.model.n_samples() == S
    xb, yb = torch.as_tensor(g["xb"]).float().cuda(), torch.as_tensor(g["yb"]).cuda()
    assert abs(obj.psvi_elbo(xb, yb, model=obj.model).item() - g["ref64_outer_val"]) <= 1e-4 * abs(g["ref64_outer_val"])
    got = obj._last_outer["phi_grad"].cpu().numpy()
    ref = g["ref64_outer_gparams"]
    from oracle import psvi_oracle_generic as pg
    assert rel_l2(pg.MeanField(dims).join(got[:len(got) // 2], got[len(got) // 2:]), ref) < 2e-4
    loss = obj.nested_step(xb, yb)
    assert abs(loss.item() - g["ref64_nested_loss"]) <= 2e-4 * abs(g["ref64_nested_loss"])
    assert rel_l2(obj.u.grad.cpu().numpy(), g["ref64_nested_gu"]) < 2e-3
    assert rel_l2(obj.v.grad.cpu().numpy(), g["ref64_nested_gv"]) < 2e-3
    vec = torch.nn.utils.parameters_to_vector(obj.model.parameters()).detach().cpu().numpy()
    assert rel_l2(vec, g["ref64_nested_params"]) < 1e-5
    obj.noise_source = None
    acc, nll, went, ness, vent = obj.evaluate()
    assert 0.0 <= acc.item() <= 1.0 and np.isfinite(nll.item())
    assert obj.model.n_samples() == S      # PSVI_No_IW evaluates with 5 samples and switches back to 1


def test_pred_on_grid_matches_reference():
    """PSVI.pred_on_grid (reference :1130-1175) against the reference's fp64 output with the same injected noise: importance-
    weighted mixture and plain mean over a 12 x 12 grid."""
    import os
    from oracle.ref_import import NoiseFeeder
    g = dict(np.load(os.path.join(GOLDEN, "grid_fn_hm.npz")))
    dims = [int(d) for d in g["dims"]]
    S, n = int(g["S"]), int(g["n"])
    eps = NoiseFeeder.stream(dims, S, int(g["noise_seed"]), int(g["n_forwards"]))
    g2 = dict(g, B=32, lr0net=1e-3)
    obj = make_obj("x_hm", g2, dims, S, 2, eps, init_sd=5e-2)
    obj.N = float(g["N"])
    p_iw = obj.pred_on_grid(n_test_per_dim=n, correction=True).cpu().numpy()
    p_mean = obj.pred_on_grid(n_test_per_dim=n, correction=False).cpu().numpy()
    assert p_iw.shape == (n * n, 2)
    np.testing.assert_allclose(p_iw, g["ref64_grid_iw"], rtol=2e-4, atol=2e-5)
    np.testing.assert_allclose(p_mean, g["ref64_grid_mean"], rtol=2e-4, atol=2e-5)


def test_prune_and_retrain_on_coreset_run():
    """run_psvi with prune=True (reference :935-944,1177-1193) and retrain_on_coreset=True (:969-997): the coreset shrinks to
    the requested sizes, the retraining phase appends its own evaluations, and the model still classifies halfmoon."""
    from psvi.experiments.experiments_utils import read_dataset
    from psvi.inference.psvi_classes import PSVIFixedU
    torch.manual_seed(0)
    x, y, xt, yt, N, D, tr, te, nc = read_dataset("halfmoon", {"test_ratio": 0.2})
    kw = dict(mc_samples=8, num_epochs=41, data_minibatch=128, D=D, N=N, inner_it=20, trainer="nested", log_every=10, lr0u=1e-3,
              lr0net=1e-2, lr0v=1e-2, lr0joint=1e-2, init_args="subsample", init_sd=1e-3, num_pseudo=40, seed=0, architecture="fn",
              n_hidden=30, n_layers=1, logistic_regression=False, train_dataset=tr, test_dataset=te, dnm="halfmoon", nc=nc,
              compute_weights_entropy=True, register_elbos=False, quiet=True, prune=True, prune_interval=15, prune_sizes=[30, 20],
              retrain_on_coreset=True)
    obj = PSVIFixedU(**kw)
    res = obj.run_psvi(**kw)
    assert obj.num_pseudo == 20 and obj.u.shape[0] == 20 and obj.v.shape[0] == 20 and obj.z.shape[0] == 20
    assert res["csizes"][:5] == [40, 40, 30, 30, 20] and len(res["accs"]) == 10      # 5 PSVI + 5 retraining evaluations
    assert res["accs"][-1] > 0.8 and all(np.isfinite(res["nlls"]))


@pytest.mark.parametrize("arch", ["fn2", "lenet"])
def test_hyper_trainer_runs_on_the_streaming_families(arch):
    """--trainer hyper (implicit hypergradient, CG on the normal equations) through the streaming engine for the
    full-covariance and the convolutional family: runs, consumes noise, returns finite metrics and moves u and v."""
    from psvi.inference.psvi_classes import PSVILearnV
    if arch == "lenet":
        from tests.fake_mnist import FakeMNIST
        tr, te = FakeMNIST(300, 0), FakeMNIST(64, 1)
        extra = dict(D=784, N=len(tr), dnm="MNIST", nc=10, n_hidden=0, num_pseudo=20, mc_samples=4)
    else:
        from psvi.experiments.experiments_utils import read_dataset
        x, y, xt, yt, N, D, tr, te, nc = read_dataset("halfmoon", {"test_ratio": 0.2})
        extra = dict(D=D, N=N, dnm="halfmoon", nc=nc, n_hidden=8, num_pseudo=10, mc_samples=6)
    kw = dict(num_epochs=3, data_minibatch=64, inner_it=4, trainer="hyper", log_every=2, lr0u=1e-3, lr0net=1e-3, lr0v=1e-2,
              init_args="subsample", init_sd=1e-2, seed=2, architecture=arch, n_layers=1, logistic_regression=False,
              train_dataset=tr, test_dataset=te, compute_weights_entropy=False, register_elbos=False, quiet=True, **extra)
    obj = PSVILearnV(**kw)
    res = obj.run_psvi(**kw)
    assert len(res["accs"]) == 2 and np.isfinite(res["nlls"]).all()
    assert torch.isfinite(obj.u).all() and torch.isfinite(obj.v).all()
    assert float(obj.v.detach().abs().max()) > 0.0          # v started at zero: the outer optimiser moved it
