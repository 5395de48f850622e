"""GPU parity tests that close the round-1 holes: the remaining PSVI variants, the fixed-point implicit solver, run_mfvi,
BASELINE configs[0] at its own T = 100 (through the generic CASES of the other test files), the tensor-core full-data
passes at the exact cfg5 shape, and "same seed => same initial weights and pseudo-data as the reference".
Goldens: tests/golden/variant_*.npz, fixedpoint_fn_hm.npz, mfvi_hm.npz (oracle/make_goldens_r2.py, unmodified reference)."""
import os

import numpy as np
import pytest
import torch

from oracle import psvi_oracle as po
from tests.gpu_util import GOLDEN, dev, rel_l2

pytestmark = pytest.mark.gpu


def _halfmoon():
    from psvi.experiments.experiments_utils import read_dataset
    torch.manual_seed(0)
    return read_dataset("halfmoon", {"test_ratio": 0.2})


@pytest.mark.parametrize("cls_name", ["PSVIAV", "PSVIAFixedU", "PSVIFixedU", "PSVIFreeV", "PSVI_No_Rescaling"])
def test_variant_nested_step_matches_reference(cls_name):
    """One nested_step of every remaining PSVI variant (reference psvi_classes.py:1363-1385,1475-1883) against the reference's
    fp64 run: hypergradients on u, v and alpha (vmode 2: f = exp(alpha) softmax), which optimisers step, the clamp of the
    free weights.  Tolerances as for PSVILearnV (rel-L2 1e-3 on hypergradients, 2e-4 on the loss)."""
    from oracle.ref_import import NoiseFeeder
    import psvi.inference.psvi_classes as pc
    g = dict(np.load(os.path.join(GOLDEN, f"variant_{cls_name}.npz")))
    dims = [int(d) for d in g["dims"]]
    S, T, M, B = int(g["S"]), int(g["T"]), int(g["M"]), int(g["B"])
    eps = NoiseFeeder.stream(dims, S, int(g["noise_seed"]), int(g["n_forwards"]))
    x, y, xt, yt, N, D, tr, te, nc = _halfmoon()
    kw = dict(mc_samples=S, num_epochs=0, data_minibatch=B, D=D, N=N, inner_it=T, trainer="nested", log_every=10, lr0u=1e-4,
              lr0net=float(g["lr0net"]), lr0v=1e-3, lr0alpha=1e-3, init_args="subsample", init_sd=1e-2, num_pseudo=M, seed=0,
              architecture="fn", n_hidden=dims[1], n_layers=1, logistic_regression=False, train_dataset=tr, test_dataset=te,
              dnm="halfmoon", nc=nc, compute_weights_entropy=True, register_elbos=False, quiet=True)
    obj = getattr(pc, cls_name)(**kw)
    obj.run_psvi(**kw)
    assert bool(obj.learn_v) == bool(int(g["learn_v"]))
    mu, rho = obj.model.flat()
    mu.copy_(torch.as_tensor(g["mu0"])), rho.copy_(torch.as_tensor(g["rho0"]))
    with torch.no_grad():
        obj.u.copy_(torch.as_tensor(g["u0"]))
        obj.v.copy_(torch.as_tensor(g["v0"]))
        if obj.alpha is not None:
            obj.alpha.fill_(float(g["alpha0"]))
    obj.z = torch.as_tensor(g["z"]).float().cuda()
    obj.scheduler_optim_net = None
    obj.noise_source = pc.ExternalNoise(eps)
    xb, yb = torch.as_tensor(g["xb"]).float().cuda(), torch.as_tensor(g["yb"]).cuda()
    loss = obj.nested_step(xb, yb)
    assert obj.noise_source.pos == int(g["n_forwards"])
    assert abs(loss.item() - g["ref64_nested_loss"]) <= 2e-4 * abs(g["ref64_nested_loss"])
    vec = torch.nn.utils.parameters_to_vector(obj.model.parameters()).detach().cpu().numpy()
    assert rel_l2(vec, g["ref64_nested_params"]) < 1e-5
    if int(g["ref64_has_gu"]):
        assert rel_l2(obj.u.grad.cpu().numpy(), g["ref64_nested_gu"]) < 1e-3
    np.testing.assert_allclose(obj.u.detach().cpu().numpy(), g["ref64_nested_u_after"], atol=2e-6)   # fixed-u: unchanged
    if int(g["learn_v"]):
        assert rel_l2(obj.v.grad.cpu().numpy(), g["ref64_nested_gv"]) < 1e-3
    np.testing.assert_allclose(obj.v.detach().cpu().numpy(), g["ref64_nested_v_after"], atol=2e-6)
    if obj.alpha is not None:
        np.testing.assert_allclose(obj.alpha.grad.cpu().numpy(), g["ref64_nested_galpha"], rtol=1e-3)
        np.testing.assert_allclose(obj.alpha.detach().cpu().numpy(), g["ref64_alpha_after"], atol=2e-6)


def test_fixed_point_hyper_step_matches_reference():
    """--trainer hyper with hypergrad_approx="fixed_point" (reference hypergradients.py:83-140, stochastic=True: a fresh noise
    draw per iteration); fp64 reference, same number of forwards."""
    from oracle.ref_import import NoiseFeeder
    from psvi.inference.psvi_classes import ExternalNoise, PSVILearnV
    g = dict(np.load(os.path.join(GOLDEN, "fixedpoint_fn_hm.npz")))
    dims = [int(d) for d in g["dims"]]
    S, T, M, B, K = int(g["S"]), int(g["T"]), int(g["M"]), int(g["B"]), int(g["K"])
    eps = NoiseFeeder.stream(dims, S, int(g["noise_seed"]), int(g["n_forwards"]))
    x, y, xt, yt, N, D, tr, te, nc = _halfmoon()
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
    xb, yb = torch.as_tensor(g["xb"]).float().cuda(), torch.as_tensor(g["yb"]).cuda()
    ll = obj.hyper_step(xb, yb, K=K, linsys_lr=float(g["linsys_lr"]), hypergrad_approx="fixed_point")
    assert obj.noise_source.pos == int(g["n_forwards"])
    assert abs(ll - g["ref64_ll"]) <= 2e-4 * abs(g["ref64_ll"])
    assert rel_l2(obj.u.grad.cpu().numpy(), g["ref64_gu"]) < 2e-3
    assert rel_l2(obj.v.grad.cpu().numpy(), g["ref64_gv"]) < 2e-3
    vec = torch.nn.utils.parameters_to_vector(obj.model.parameters()).detach().cpu().numpy()
    assert rel_l2(vec, g["ref64_params"]) < 1e-5


@pytest.mark.parametrize("trainer", ["joint", "alternating"])
def test_joint_and_alternating_trainers_match_reference(trainer):
    """--trainer joint / alternating (reference psvi_classes.py:517-539,871-880): two steps against the fp64 reference with the
    same injected noise; then the trainer runs end to end through run_psvi."""
    from oracle.ref_import import NoiseFeeder
    from psvi.inference.psvi_classes import ExternalNoise, PSVILearnV
    g = dict(np.load(os.path.join(GOLDEN, f"{trainer}_fn_hm.npz")))
    dims = [int(d) for d in g["dims"]]
    S, M, B = int(g["S"]), int(g["M"]), int(g["B"])
    eps = NoiseFeeder.stream(dims, S, int(g["noise_seed"]), int(g["n_forwards"]))
    x, y, xt, yt, N, D, tr, te, nc = _halfmoon()
    kw = dict(mc_samples=S, num_epochs=0, data_minibatch=B, D=D, N=N, inner_it=1, trainer=trainer, log_every=10, lr0u=1e-4,
              lr0net=float(g["lr0net"]), lr0v=1e-3, lr0joint=float(g["lr0joint"]), init_args="subsample", init_sd=1e-2,
              num_pseudo=M, seed=0, architecture="fn", n_hidden=dims[1], n_layers=1, logistic_regression=False, train_dataset=tr,
              test_dataset=te, dnm="halfmoon", nc=nc, compute_weights_entropy=True, register_elbos=True, quiet=True)
    obj = PSVILearnV(**kw)
    obj.run_psvi(**kw)
    mu, rho = obj.model.flat()
    mu.copy_(torch.as_tensor(g["mu0"])), rho.copy_(torch.as_tensor(g["rho0"]))
    with torch.no_grad():
        obj.u.copy_(torch.as_tensor(g["u0"]))
        obj.v.copy_(torch.as_tensor(g["v0"]))
    obj.z = torch.as_tensor(g["z"]).float().cuda()
    obj.noise_source = ExternalNoise(eps)
    xb, yb = torch.as_tensor(g["xb"]).float().cuda(), torch.as_tensor(g["yb"]).cuda()
    step = obj.joint_step if trainer == "joint" else obj.alternating_step
    losses = [float(step(xb, yb)) for _ in range(int(g["steps"]))]
    assert obj.noise_source.pos == int(g["n_forwards"])
    np.testing.assert_allclose(losses, g["ref64_losses"], rtol=2e-5)
    vec = torch.nn.utils.parameters_to_vector(obj.model.parameters()).detach().cpu().numpy()
    # Adam's first steps move every coordinate by ~lr: compare the MOVES (1e-3 / 1e-4 per step), not just the values
    ref0 = po.mu_rho_to_phi(g["mu0"], g["rho0"], dims)
    assert rel_l2(vec - ref0, g["ref64_params"] - ref0) < 2e-3
    assert rel_l2(obj.u.detach().cpu().numpy() - g["u0"], g["ref64_u_after"] - g["u0"]) < 2e-3
    if trainer == "joint":
        assert rel_l2(obj.v.detach().cpu().numpy() - g["v0"], g["ref64_v_after"] - g["v0"]) < 2e-3
    else:
        np.testing.assert_array_equal(obj.v.detach().cpu().numpy(), g["v0"].astype(np.float32))
    assert [t for t, _ in obj.elbos] == ([2, 2] if trainer == "joint" else [0, 1, 0, 1])
    # end to end: the flag value drives run_psvi (reference :871-886)
    kw.update(num_epochs=12, log_every=6, register_elbos=False)
    obj2 = PSVILearnV(**kw)
    res = obj2.run_psvi(**kw)
    assert len(res["accs"]) == 2 and np.isfinite(res["nlls"]).all()


def test_run_mfvi_matches_reference_trace():
    """run_mfvi (reference baselines.py:824-920) with full-batch steps: ELBO trace, accuracies and NLLs."""
    from oracle.ref_import import NoiseFeeder
    from psvi.inference.baselines import run_mfvi
    from psvi.inference.psvi_classes import ExternalNoise
    g = dict(np.load(os.path.join(GOLDEN, "mfvi_hm.npz")))
    dims = [int(d) for d in g["dims"]]
    S = int(g["S"])
    eps = NoiseFeeder.stream(dims, S, int(g["noise_seed"]), int(g["n_forwards"]))
    x, y, xt, yt, N, D, tr, te, nc = _halfmoon()
    res = run_mfvi(xt=xt, yt=yt, mc_samples=S, data_minibatch=N, num_epochs=3, log_every=2, N=N, D=D, lr0net=1e-3, seed=5,
                   architecture="fn", n_hidden=dims[1], nc=nc, train_dataset=tr, test_dataset=te, init_sd=1e-3,
                   noise_source=ExternalNoise(eps), quiet=True)
    np.testing.assert_allclose(res["elbos"], g["ref_elbos"], rtol=5e-5)
    np.testing.assert_allclose(res["accs"], g["ref_accs"], atol=1e-6)
    np.testing.assert_allclose(res["nlls"], g["ref_nlls"], rtol=5e-5)
    assert res["csizes"] is None


def test_same_seed_gives_the_reference_initial_state():
    """the boundary classes consume the CPU generator in the reference's order, so a fresh
    PSVILearnV(seed=0) holds the reference's initial weights and pseudo-data -- nothing injected here."""
    from psvi.inference.psvi_classes import PSVILearnV
    g = dict(np.load(os.path.join(GOLDEN, "fn_hm_m50_t10.npz")))
    dims = [int(d) for d in g["dims"]]
    x, y, xt, yt, N, D, tr, te, nc = _halfmoon()
    kw = dict(mc_samples=int(g["S"]), num_epochs=0, data_minibatch=int(g["B"]), D=D, N=N, inner_it=int(g["T"]), trainer="nested",
              log_every=10, lr0u=1e-4, lr0net=float(g["lr0net"]), lr0v=1e-3, init_args="subsample", init_sd=1e-3,
              num_pseudo=int(g["M"]), seed=0, architecture="fn", n_hidden=dims[1], n_layers=1, logistic_regression=False,
              train_dataset=tr, test_dataset=te, dnm="halfmoon", nc=nc, compute_weights_entropy=True, register_elbos=True,
              quiet=True)
    obj = PSVILearnV(**kw)
    obj.run_psvi(**kw)
    mu, rho = obj.model.flat()
    np.testing.assert_allclose(mu.cpu().numpy(), g["mu0"], rtol=0, atol=1e-7)
    np.testing.assert_allclose(rho.cpu().numpy(), g["rho0"], rtol=1e-6)
    np.testing.assert_allclose(obj.u.detach().cpu().numpy(), g["u0"], rtol=0, atol=1e-7)
    np.testing.assert_array_equal(obj.z.cpu().numpy(), g["z"])


def test_fn_tc_forward_at_the_exact_cfg5_shape():
    """psvi_fn_nll_tc at BASELINE configs[4]'s exact shape (D=256, H=1024, C=10, S=64) over 300 rows, against the oracle on
    the same bf16-rounded operands (the first set stops at S=16 for H=1024)."""
    from psvi import _native as nat
    D, H, C, S, R = 256, 1024, 10, 64, 300
    rng = np.random.default_rng(64)
    dims = [D, H, C]
    P = po.p_theta(dims)
    mu = np.concatenate([rng.standard_normal(H * D) / np.sqrt(D), np.zeros(H), rng.standard_normal(C * H) / np.sqrt(H),
                         np.zeros(C)]).astype(np.float32)
    rho = np.full(P, po.inverse_softplus(0.02), np.float32)
    eps = rng.standard_normal((1, S, P)).astype(np.float32)
    X, y = rng.standard_normal((R, D)).astype(np.float32), rng.integers(0, C, R)
    bf = lambda a: torch.as_tensor(np.asarray(a, np.float32)).bfloat16().double().numpy()  # noqa: E731
    m = nat.make_model(dims, S)
    wsum, nkl, nll = torch.zeros(S, device="cuda"), torch.zeros(S, device="cuda"), torch.zeros(S, R, device="cuda")
    scr = torch.zeros(nat.fn_tc_scratch_floats(m, R, 0), device="cuda")
    nat.fn_nll_tc(m, nat.make_noise(dev(eps)), dev(mu), dev(rho), dev(X).bfloat16().contiguous(), dev(y, torch.int32), None, 0,
                  wsum, nkl, nll, scr)
    torch.cuda.synchronize()
    th = po.mf_sample(mu.astype(np.float64), rho.astype(np.float64), eps[0].astype(np.float64))
    W1 = bf(th[:, :H * D]).reshape(S, H, D)
    W2 = bf(th[:, H * D + H:H * D + H + C * H]).reshape(S, C, H)
    hid = bf(np.maximum(np.einsum("rd,shd->srh", bf(X), W1).astype(np.float32) + th[:, None, H * D:H * D + H], 0))
    ref = po.nll_rows(np.einsum("srh,sch->src", hid, W2) + th[:, None, H * D + H + C * H:], y)[0]
    got = nll.cpu().numpy()
    assert np.abs(got - ref).max() < 1e-2 and rel_l2(got, ref) < 2e-3
    np.testing.assert_allclose(wsum.cpu().numpy(), ref.sum(1), rtol=2e-3)


def test_incremental_learning_grows_coreset_and_classes():
    """increment=True (reference psvi_classes.py:823-832,945-966,1194-1217): start with two classes, add one class and a block
    of coreset points every `increment_interval` outer steps; the model is rebuilt with one more output each time."""
    from psvi.experiments.experiments_utils import read_dataset
    from psvi.inference.psvi_classes import PSVILearnV
    torch.manual_seed(0)
    x, y, xt, yt, N, D, tr, te, nc = read_dataset("four_blobs", {"test_ratio": 0.2})
    assert nc == 4
    kw = dict(mc_samples=6, num_epochs=13, data_minibatch=64, D=D, N=N, inner_it=5, trainer="nested", log_every=2, lr0u=1e-3,
              lr0net=1e-2, lr0v=1e-2, init_args="subsample", init_sd=1e-3, num_pseudo=8, seed=0, architecture="fn", n_hidden=30,
              n_layers=1, logistic_regression=False, train_dataset=tr, test_dataset=te, dnm="four_blobs", nc=nc,
              increment=True, increment_interval=4, increment_sizes=[8, 12, 16], compute_weights_entropy=True,
              register_elbos=False, quiet=True)
    obj = PSVILearnV(**kw)
    res = obj.run_psvi(**kw)
    assert res["csizes"][0] == 8 and res["csizes"][-1] == 16 and sorted(set(res["csizes"])) == [8, 12, 16]
    assert obj.nc == 4 and obj.model.dims[-1] == 4
    assert obj.u.shape == (16, D) and obj.v.shape == (16,) and obj.z.shape == (16,)
    assert set(obj.z.cpu().numpy().astype(int).tolist()) <= {0, 1, 2, 3} and 3 in set(obj.z.cpu().numpy().astype(int).tolist())
    assert all(np.isfinite(res["nlls"])) and all(0.0 <= a <= 1.0 for a in res["accs"])


@pytest.mark.parametrize("arch", ["fn2", "lenet"])
def test_run_mfvi_subset_for_fn2_and_lenet(arch):
    """mfvi_subset beyond mean-field MLPs (reference baselines.py:923-1062 runs them with the Q5 KL filter: no KL term for fn2,
    none for lenet's conv layers): the ELBO improves over the run and the metrics are finite."""
    from psvi.inference.baselines import run_mfvi_subset
    if arch == "fn2":
        x, y, xt, yt, N, D, tr, te, nc = _halfmoon()
        kw = dict(D=D, n_hidden=8, nc=nc, dnm="halfmoon", init_sd=5e-2, lr0net=1e-2, num_pseudo=40)
    else:
        from psvi.experiments.experiments_utils import SynthDataset
        from tests.fake_mnist import FakeMNIST
        ftr, fte = FakeMNIST(256, 0), FakeMNIST(64, 1)
        x = torch.stack([ftr[i][0] for i in range(len(ftr))]).reshape(len(ftr), -1)
        y = torch.tensor([ftr[i][1] for i in range(len(ftr))]).float()
        xt = torch.stack([fte[i][0] for i in range(len(fte))]).reshape(len(fte), -1)
        yt = torch.tensor([fte[i][1] for i in range(len(fte))]).float()
        tr, te, nc, D = SynthDataset(x, y), SynthDataset(xt, yt), 10, 784
        kw = dict(D=D, n_hidden=0, nc=nc, dnm="MNIST", init_sd=None, lr0net=1e-3, num_pseudo=50)
    res = run_mfvi_subset(x=x, y=y, xt=xt, yt=yt, mc_samples=4, data_minibatch=64, num_epochs=15, log_every=10, seed=1,
                          train_dataset=tr, test_dataset=te, init_args="subsample", architecture=arch, quiet=True, **kw)
    assert len(res["elbos"]) == 30 and all(np.isfinite(res["elbos"])) and all(np.isfinite(res["nlls"]))
    if arch == "lenet":
        assert np.mean(res["elbos"][-5:]) > np.mean(res["elbos"][:5])
    else:   # fn2 starts from zero means (make_fc2net): a symmetric point the first steps leave only slowly; stay finite / sane
        assert abs(np.mean(res["elbos"][-5:]) - np.mean(res["elbos"][:5])) < 0.05 * abs(np.mean(res["elbos"][:5]))
    assert all(0.0 <= a <= 1.0 for a in res["accs"])


def test_learn_z_soft_labels_match_reference():
    """learn_z=True (reference psvi_classes.py:455-474,499-504,546-547,594-595,1049-1056): inner_elbo, psvi_elbo, one
    nested_step with hypergradients on u, v AND the soft labels z, the optim_z update, evaluate(); then PSVIEvaluate and
    run_psvi end to end.  The kernels see every soft-label row as C weighted hard-label rows."""
    from oracle.ref_import import NoiseFeeder
    from psvi.experiments.experiments_utils import SynthDataset, read_dataset
    from psvi.inference.psvi_classes import ExternalNoise, PSVIEvaluate, PSVILearnV
    g = dict(np.load(os.path.join(GOLDEN, "learnz_fn_fb.npz")))
    dims = [int(d) for d in g["dims"]]
    S, T, M, B = int(g["S"]), int(g["T"]), int(g["M"]), int(g["B"])
    eps = NoiseFeeder.stream(dims, S, int(g["noise_seed"]), int(g["n_forwards"]))
    torch.manual_seed(0)
    x, y, xt, yt, N, D, tr, te, nc = read_dataset("four_blobs", {"test_ratio": 0.2})
    te = SynthDataset(torch.as_tensor(g["xt"]).float(), torch.as_tensor(g["yt"]).float())     # the golden's own test rows
    kw = dict(mc_samples=S, num_epochs=0, data_minibatch=B, D=D, N=N, inner_it=T, trainer="nested", log_every=10, lr0u=1e-4,
              lr0net=float(g["lr0net"]), lr0v=1e-3, lr0z=float(g["lr0z"]), init_args="subsample", init_sd=1e-2, num_pseudo=M,
              seed=0, architecture="fn", n_hidden=dims[1], n_layers=1, logistic_regression=False, train_dataset=tr,
              test_dataset=te, dnm="four_blobs", nc=nc, compute_weights_entropy=True, register_elbos=False, quiet=True,
              learn_z=True)
    obj = PSVILearnV(**kw)
    obj.run_psvi(**kw)
    assert obj.z.shape == (M, nc) and obj.z.requires_grad and float(obj.z.sum()) == M      # one-hot initial soft labels
    mu, rho = obj.model.flat()
    mu.copy_(torch.as_tensor(g["mu0"])), rho.copy_(torch.as_tensor(g["rho0"]))
    with torch.no_grad():
        obj.u.copy_(torch.as_tensor(g["u0"]))
        obj.v.copy_(torch.as_tensor(g["v0"]))
        obj.z.copy_(torch.as_tensor(g["z0"]))
    obj.scheduler_optim_net = None
    obj.noise_source = ExternalNoise(eps)
    xb, yb = torch.as_tensor(g["xb"]).float().cuda(), torch.as_tensor(g["yb"]).cuda()
    ie = float(obj.inner_elbo(model=obj.model))
    assert abs(ie - g["ref64_inner_elbo"]) <= 2e-5 * abs(g["ref64_inner_elbo"])
    oe = float(obj.psvi_elbo(xb, yb, model=obj.model))
    assert abs(oe - g["ref64_psvi_elbo"]) <= 2e-5 * abs(g["ref64_psvi_elbo"])
    loss = float(obj.nested_step(xb, yb))
    assert obj.noise_source.pos == int(g["n_forwards_step"])
    assert abs(loss - g["ref64_nested_loss"]) <= 2e-5 * abs(g["ref64_nested_loss"])
    assert rel_l2(obj.u.grad.cpu().numpy(), g["ref64_gu"]) < 2e-3
    assert rel_l2(obj.v.grad.cpu().numpy(), g["ref64_gv"]) < 2e-3
    assert rel_l2(obj.z.grad.cpu().numpy(), g["ref64_gz"]) < 2e-3
    vec = torch.nn.utils.parameters_to_vector(obj.model.parameters()).detach().cpu().numpy()
    assert rel_l2(vec, g["ref64_params"]) < 1e-5
    np.testing.assert_allclose(obj.z.detach().cpu().numpy(), g["ref64_z_after"], atol=2e-5)    # Adam(lr0z) step on z
    acc, nll, went, ness, vent = obj.evaluate()
    assert obj.noise_source.pos == int(g["n_forwards"])
    np.testing.assert_allclose([float(acc), float(nll), float(went), float(ness), float(vent)], g["ref32_eval"], rtol=2e-3)
    # PSVIEvaluate: only the network trains; u, z, v stay put
    kw.update(register_elbos=True)
    ev = PSVIEvaluate(**kw)
    ev.run_psvi(**kw)
    u0, z0, v0 = ev.u.detach().clone(), ev.z.detach().clone(), ev.v.detach().clone()
    p0 = torch.nn.utils.parameters_to_vector(ev.model.parameters()).detach().clone()
    l1 = float(ev.nested_step(xb, yb))
    assert np.isfinite(l1) and torch.equal(ev.u.detach(), u0) and torch.equal(ev.z.detach(), z0) and torch.equal(ev.v.detach(), v0)
    assert not torch.equal(torch.nn.utils.parameters_to_vector(ev.model.parameters()).detach(), p0)
    assert [t for t, _ in ev.elbos] == [1, 0]
    # end to end
    kw.update(num_epochs=6, log_every=3, register_elbos=False)
    res = PSVILearnV(**kw).run_psvi(**kw)
    assert len(res["accs"]) == 2 and np.isfinite(res["nlls"]).all()


def test_meanfieldvi_class_matches_reference(tmp_path):
    """MeanFieldVI (reference psvi/inference/utils.py:221-450): the reference's own run() loop -- sequential minibatch Adam
    steps, forgetting-score bookkeeping after every epoch, mean-of-logits tests -- with the injected noise stream; then save /
    load of the fitted net."""
    from oracle.ref_import import NoiseFeeder
    from psvi.experiments.experiments_utils import SynthDataset
    from psvi.inference.psvi_classes import ExternalNoise
    from psvi.inference.utils import MeanFieldVI
    g = dict(np.load(os.path.join(GOLDEN, "meanfieldvi_hm.npz")))
    dims = [int(d) for d in g["dims"]]
    S, B = int(g["S"]), int(g["B"])
    eps = NoiseFeeder.stream(dims, S, int(g["noise_seed"]), int(g["n_forwards"]))
    tr = SynthDataset(torch.as_tensor(g["x"]).float(), torch.as_tensor(g["y"]).float())
    te = SynthDataset(torch.as_tensor(g["xt"]).float(), torch.as_tensor(g["yt"]).float())
    kw = dict(mc_samples=S, data_minibatch=B, num_epochs=2, log_every=2, N=int(g["N"]), D=dims[0], lr0net=float(g["lr0net"]),
              mul_fact=2, seed=3, architecture="fn", n_hidden=dims[1], nc=dims[-1], train_dataset=tr, test_dataset=te,
              init_sd=1e-2, forgetting_score_flag=True, data_path=str(tmp_path), dnm="halfmoon", quiet=True)
    m = MeanFieldVI(noise_source=ExternalNoise(eps), **kw)
    m.before_train()
    mu, rho = m.net.flat()
    mu.copy_(torch.as_tensor(g["mu0"])), rho.copy_(torch.as_tensor(g["rho0"]))
    for i in range(m.total_iterations):                 # reference run(), :396-401
        m.train_an_epoch()
        m.after_epoch()
        if i % m.log_every == 0 or i == m.total_iterations - 1:
            m.test()
    assert m.noise_source.pos == int(g["n_forwards"])
    forgetting = torch.max(m.total_iterations * m.never_learnt_events, m.forgetting_events).cpu().numpy()
    np.testing.assert_allclose(m.elbos_mfvi, g["ref_elbos"], rtol=1e-4)
    np.testing.assert_allclose(m.accs_mfvi, g["ref_accs"], atol=5.1e-3)          # (one test row of 200 may flip)
    np.testing.assert_allclose(m.nlls_mfvi, g["ref_nlls"], rtol=1e-3)
    assert np.mean(forgetting != g["ref_forgetting"]) <= 0.01
    assert np.mean(m.last_acc.cpu().numpy() != g["ref_last_acc"]) <= 0.01
    # run() end to end with Philox noise, then load_from_saved picks the stored net up
    m2 = MeanFieldVI(**kw)
    m2.run()
    assert len(m2.accs_mfvi) == 3 and np.isfinite(m2.nlls_mfvi).all() and m2.elbos_mfvi[-1] > m2.elbos_mfvi[0]
    m3 = MeanFieldVI(load_from_saved=True, **kw)
    m3.run()
    assert m3.accs_mfvi == []                                                       # nothing was retrained
    p2 = torch.nn.utils.parameters_to_vector(m2.net.parameters())
    p3 = torch.nn.utils.parameters_to_vector(m3.net.parameters())
    assert torch.equal(p2, p3) and torch.equal(m2.forgetting_events, m3.forgetting_events)


@pytest.mark.parametrize("rows,H,S,L", [(8, 16, 5, 1), (200, 16, 5, 2), (400, 100, 10, 1)])
def test_module_forward_matches_oracle(rows, H, S, L):
    """`model(x)` of a mean-field stack (reference neural_net.py:176-179 through nn.Sequential): psvi_mf_forward against the
    oracle's forward on the sampled weights the kernel reports (tolerance: fp32 accumulation order, 1e-5 relative)."""
    import torch.nn as nn
    from psvi.models.neural_net import VILinear, make_fcnet
    torch.manual_seed(1)
    net = make_fcnet(2, H, 3, n_layers=L, linear_class=VILinear, nonl_class=nn.ReLU, mc_samples=S, init_sd=0.1).cuda()
    x = torch.randn(rows, 2, device="cuda")
    lg = net(x)
    assert lg.shape == (S, rows, 3)
    dims = [2] + [H] * L + [3]
    theta = torch.cat([torch.cat([m._cached_weight.reshape(S, -1), m._cached_bias.reshape(S, -1)], 1) for m in net.vi_layers()], 1)
    ref, _ = po.mlp_forward(theta.double().cpu().numpy(), x.double().cpu().numpy(), dims)
    assert rel_l2(lg.cpu().numpy(), ref) < 1e-5


@pytest.mark.parametrize("cls_name", ["PSVILearnV_regressor", "PSVIAV_regressor"])
def test_gaussian_regressors_match_reference(cls_name):
    """The Gaussian-likelihood regressors (reference psvi_classes.py:1940-2335) on psvi_net_pass_gaussian: inner_elbo, psvi_elbo,
    one nested_step with hypergradients on u, v, the learnable targets z (and alpha), the Adam steps, evaluate() (RMSE / mean
    log-likelihood), then run_psvi end to end.  Goldens: `python oracle/make_goldens_r2.py regressor`."""
    from oracle.ref_import import NoiseFeeder
    from psvi.experiments.experiments_utils import BaseDataset
    from psvi.inference import psvi_classes as pc
    g = dict(np.load(os.path.join(GOLDEN, f"regressor_{cls_name}.npz")))
    dims = [int(d) for d in g["dims"]]
    S, T, M, B, tau = int(g["S"]), int(g["T"]), int(g["M"]), int(g["B"]), float(g["tau"])
    eps = NoiseFeeder.stream(dims, S, int(g["noise_seed"]), int(g["n_forwards"]))
    T_ = lambda a: torch.as_tensor(np.asarray(a, dtype=np.float32))
    tr, va, te = BaseDataset(T_(g["x"]), T_(g["y"])), BaseDataset(T_(g["xv"]), T_(g["yv"])), BaseDataset(T_(g["xt"]), T_(g["yt"]))
    kw = dict(mc_samples=S, num_epochs=0, data_minibatch=B, D=dims[0], N=int(g["N"]), inner_it=T, trainer="nested",
              architecture="regressor_net", n_hidden=dims[1], nc=1, train_dataset=tr, val_dataset=va, test_dataset=te,
              y_mean=torch.tensor(float(g["y_mean"])), y_std=torch.tensor(float(g["y_std"])), tau=tau, num_pseudo=M,
              init_args="subsample", lr0net=float(g["lr0net"]), lr0u=1e-3, lr0v=1e-2, lr0z=float(g["lr0z"]), init_sd=1e-2,
              log_every=10, seed=0, quiet=True)
    obj = getattr(pc, cls_name)(**kw)
    obj.run_psvi(**kw)
    assert obj.model.dims == dims and obj.u.shape == (M, dims[0]) and obj.z.shape == (M,) and obj.z.requires_grad
    mu, rho = obj.model.flat()
    mu.copy_(torch.as_tensor(g["mu0"])), rho.copy_(torch.as_tensor(g["rho0"]))
    with torch.no_grad():
        obj.u.copy_(torch.as_tensor(g["u0"])), obj.v.copy_(torch.as_tensor(g["v0"])), obj.z.copy_(torch.as_tensor(g["z0"]))
        if obj.alpha is not None:
            obj.alpha.fill_(float(g["alpha0"]))
    obj.noise_source = pc.ExternalNoise(eps)
    xb, yb = T_(g["xb"]).cuda(), T_(g["yb"]).cuda()
    ie = float(obj.inner_elbo(model=obj.model))
    assert abs(ie - g["ref64_inner_elbo"]) <= 2e-5 * abs(g["ref64_inner_elbo"])
    oe = float(obj.psvi_elbo(xb, yb, model=obj.model))
    assert abs(oe - g["ref64_psvi_elbo"]) <= 2e-5 * abs(g["ref64_psvi_elbo"])
    loss = float(obj.nested_step(xb, yb))
    assert obj.noise_source.pos == int(g["n_forwards_step"])
    assert abs(loss - g["ref64_nested_loss"]) <= 2e-5 * abs(g["ref64_nested_loss"])
    assert rel_l2(obj.u.grad.cpu().numpy(), g["ref64_gu"]) < 2e-3
    assert rel_l2(obj.v.grad.cpu().numpy(), g["ref64_gv"]) < 2e-3
    assert rel_l2(obj.z.grad.cpu().numpy(), g["ref64_gz"]) < 2e-3
    if obj.alpha is not None:
        np.testing.assert_allclose(obj.alpha.grad.cpu().numpy(), g["ref64_galpha"], rtol=2e-3)
    vec = torch.nn.utils.parameters_to_vector(obj.model.parameters()).detach().cpu().numpy()
    assert rel_l2(vec, g["ref64_params"]) < 1e-5
    np.testing.assert_allclose(obj.z.detach().cpu().numpy(), g["ref64_z_after"], atol=2e-5)
    np.testing.assert_allclose(obj.u.detach().cpu().numpy(), g["ref64_u_after"], atol=2e-5)
    rmse, ll = obj.evaluate()
    assert obj.noise_source.pos == int(g["n_forwards"])
    np.testing.assert_allclose([float(rmse), float(ll)], g["ref32_eval"], rtol=1e-3)
    kw.update(num_epochs=30, log_every=10, inner_it=5, lr0u=1e-2)
    res = getattr(pc, cls_name)(**kw).run_psvi(**kw)
    assert len(res["rmses"]) == 3 and np.isfinite(res["rmses"]).all() and np.isfinite(res["lls"]).all()


def test_regression_baselines_match_reference():
    """run_mfvi_regressor (precision selected on the validation set) and run_mfvi_subset_regressor (reference baselines.py:1066-
    1346) with the injected noise stream: same seeds => the same initial nets as the reference (asserted), same ELBO / RMSE /
    log-likelihood traces (fp32 reference; 1e-3)."""
    from oracle.ref_import import NoiseFeeder
    from psvi.experiments.experiments_utils import BaseDataset
    from psvi.inference import baselines as bl
    from psvi.inference.psvi_classes import ExternalNoise
    g = dict(np.load(os.path.join(GOLDEN, "regbase_mfvi.npz")))
    dims = [int(d) for d in g["dims"]]
    S, B = int(g["S"]), int(g["B"])
    eps = NoiseFeeder.stream(dims, S, int(g["noise_seed"]), int(g["n_forwards"]))
    T_ = lambda a: torch.as_tensor(np.asarray(a, dtype=np.float32))
    tr, va, te = BaseDataset(T_(g["x"]), T_(g["y"])), BaseDataset(T_(g["xv"]), T_(g["yv"])), BaseDataset(T_(g["xt"]), T_(g["yt"]))
    common = dict(mc_samples=S, data_minibatch=B, num_epochs=3, log_every=2, D=dims[0], lr0net=1e-2, seed=0,
                  architecture="regressor_net", n_hidden=dims[1], train_dataset=tr, val_dataset=va, test_dataset=te, nc=1,
                  y_mean=torch.tensor(float(g["y_mean"])), y_std=torch.tensor(float(g["y_std"])), init_sd=1e-2)
    inits, real_setup = [], bl.set_up_model

    def recording_setup(**kw):
        net = real_setup(**kw)
        inits.append(np.concatenate([np.concatenate([m.weight.detach().reshape(-1).numpy(), m.bias.detach().reshape(-1).numpy()])
                                     for m in net.vi_layers()]))
        return net
    bl.set_up_model = recording_setup
    try:
        src = ExternalNoise(eps)
        r1 = bl.run_mfvi_regressor(taus=[0.3, 0.9], model_selection=True, dnm="synthetic", noise_source=src, **common)
        assert src.pos == int(g["n_forwards_full"])
        r2 = bl.run_mfvi_subset_regressor(taus=[0.5], model_selection=False, num_pseudo=40, noise_source=src, **common)
        assert src.pos == int(g["n_forwards"])
    finally:
        bl.set_up_model = real_setup
    np.testing.assert_allclose(np.stack(inits), g["mu0s"], atol=1e-7)          # same seeds, same initial means as the reference
    assert abs(r1["scale"] - float(g["ref_full_scale"])) < 1e-6 and r1["selected_tau"] == 0.3
    for tag, r in (("full", r1), ("subset", r2)):
        np.testing.assert_allclose(r["elbos"], g[f"ref_{tag}_elbos"], rtol=1e-4)
        np.testing.assert_allclose(r["rmses"], g[f"ref_{tag}_rmses"], rtol=1e-3)
        np.testing.assert_allclose(r["lls"], g[f"ref_{tag}_lls"], rtol=1e-3)
    assert r2["csizes"] == [40]


def test_sparsebbvi_matches_reference():
    """run_sparsevi_with_bb_elbo (reference psvi/inference/sparsebbvi.py:28-198) on psvi_net_pass_bernoulli with the injected noise
    stream: coreset sizes, accuracy and NLL trace of the fp32 reference; and through the flow's method table."""
    from oracle.ref_import import NoiseFeeder
    from psvi.experiments.flow_psvi import inf_dict
    from psvi.inference.psvi_classes import ExternalNoise
    from psvi.inference.sparsebbvi import run_sparsevi_with_bb_elbo
    g = dict(np.load(os.path.join(GOLDEN, "sparsebbvi_hm.npz")))
    dims = [int(d) for d in g["dims"]]
    S = int(g["S"])
    eps = NoiseFeeder.stream(dims, S, int(g["noise_seed"]), int(g["n_forwards"]))
    src = ExternalNoise(eps)
    kw = dict(n_layers=1, logistic_regression=False, n_hidden=dims[1], log_every=int(g["log_every"]), lr0=float(g["lr0"]),
              register_elbos=True, seed=int(g["seed"]), num_epochs=int(g["num_epochs"]), inner_it=int(g["inner_it"]),
              outer_it=int(g["outer_it"]), x=torch.as_tensor(g["x"]), y=torch.as_tensor(g["y"]), xt=torch.as_tensor(g["xt"]),
              yt=torch.as_tensor(g["yt"]), mc_samples=S, data_minibatch=int(g["data_minibatch"]), scatterplot_coreset=False)
    res = run_sparsevi_with_bb_elbo(noise_source=src, _init=np.concatenate([g["mu0"], g["rho0"]]), **kw)
    assert src.pos == int(g["n_forwards"])
    assert res["csizes"] == [int(c) for c in g["ref_csizes"]]
    np.testing.assert_allclose(res["accs"], g["ref_accs"], atol=5.1e-3)
    np.testing.assert_allclose(res["nlls"], g["ref_nlls"], rtol=5e-4)
    assert len(res["core_idcs"]) == len(set(res["core_idcs"])) and float(res["w"].min()) >= 0.0
    assert [t for t, _ in res["elbos"]][:2] == [1, 0]
    assert inf_dict["sparsebbvi"] is run_sparsevi_with_bb_elbo
    r2 = run_sparsevi_with_bb_elbo(**{**kw, "logistic_regression": True, "register_elbos": False})      # Philox noise, logistic model
    assert np.isfinite(r2["nlls"]).all() and r2["csizes"][0] == 0
