"""The convolutional family (lenet) on the GPU: the per-sample network pass of csrc/psvi_lenet.cu (forward / gradient /
Hessian-vector flavours) against the fp64 oracle, and the drop-in PSVILearnV methods (inner_elbo, psvi_elbo, nested_step,
evaluate) against the UNMODIFIED reference's fp64 run stored in tests/golden/lenet_m10.npz.

Tolerances (fp32 kernels, different summation order than the fp64 reference): single passes rel-L2 2e-4, nested
hypergradients rel-L2 2e-3, values rtol 1e-4."""
import os

import numpy as np
import pytest
import torch

from oracle import lenet_oracle as lo
from oracle import psvi_oracle as po
from oracle.ref_import import LeNetNoiseFeeder
from tests.gpu_util import GOLDEN, dev, rel_l2, zeros

pytestmark = pytest.mark.gpu


def _case(S, R, seed):
    rng = np.random.default_rng(seed)
    th = []
    for (_, wshape, nb, _) in lo.LAYERS:
        fan_in = int(np.prod(wshape[1:]))
        th += [rng.standard_normal((S, int(np.prod(wshape)))) / np.sqrt(fan_in), 0.1 * rng.standard_normal((S, nb))]
    theta = np.concatenate(th, 1)
    thetad = 0.3 * rng.standard_normal(theta.shape) * np.abs(theta).mean()
    X = rng.standard_normal((R, 784))
    y = rng.integers(0, 10, R)
    cw = rng.uniform(0.5, 1.5, (S, R))
    return theta, thetad, X, y, cw


# (10, 200) is the pass of BASELINE configs[3] at full size (S = 10 MC samples, M = 200 pseudo-images): ~20 s of fp64 oracle
@pytest.mark.parametrize("S,R", [(1, 3), (3, 7), (4, 19), (10, 200)])
def test_lenet_pass_matches_oracle(S, R):
    from psvi import _native as nat
    nat.require_cuda()
    theta, thetad, X, y, cw = _case(S, R, 10 * S + R)
    P = lo.P
    assert nat.lenet_num_theta() == P
    th, thd, x_, y_, cw_ = dev(theta), dev(thetad), dev(X), dev(y, torch.int32), dev(cw)
    # forward
    nll, logits = zeros(S, R), zeros(S, R, 10)
    nat.lenet_pass(S, th, None, x_, y_, None, nll=nll, logits=logits)
    o, cache = lo.LeNet.forward(theta, X)
    ref_nll, p = po.nll_rows(o, y)
    assert rel_l2(logits.cpu().numpy(), o) < 2e-5
    np.testing.assert_allclose(nll.cpu().numpy(), ref_nll, rtol=1e-4, atol=1e-5)
    # gradient pass
    tbar, xbar = zeros(S, P), zeros(S, R, 784)
    nat.lenet_pass(S, th, None, x_, y_, cw_, nll=nll, tbar=tbar, xbar=xbar)
    q = p.copy()
    q[:, np.arange(R), y] -= 1.0
    At, Ax = lo.LeNet.backward(theta, cache, cw[:, :, None] * q)
    assert rel_l2(tbar.cpu().numpy(), At) < 2e-4
    assert rel_l2(xbar.cpu().numpy(), Ax) < 2e-4
    # dual (Hessian-vector) pass
    tbar, tdbar, xbar, ac = zeros(S, P), zeros(S, P), zeros(S, R, 784), zeros(S, R)
    nat.lenet_pass(S, th, thd, x_, y_, cw_, tbar=tbar, tdbar=tdbar, xbar=xbar, acbar=ac)
    torch.cuda.synchronize()
    o, od, c2 = lo.LeNet.dual_forward(theta, thetad, X)
    c = cw[:, :, None]
    At, Atd, Ax = lo.LeNet.dual_backward(theta, thetad, c2, c * p * (od - (p * od).sum(-1, keepdims=True)), c * q)
    assert rel_l2(tbar.cpu().numpy(), At) < 3e-4
    assert rel_l2(tdbar.cpu().numpy(), Atd) < 2e-4
    assert rel_l2(xbar.cpu().numpy(), Ax) < 3e-4
    assert rel_l2(ac.cpu().numpy(), (q * od).sum(-1)) < 2e-4


def make_lenet_obj(g, S, T, eps):
    from psvi.inference.psvi_classes import ExternalNoise, PSVILearnV
    from psvi.models.neural_net import MeanFieldLeNet
    from tests.fake_mnist import FakeMNIST
    tr, te = FakeMNIST(64, 0), FakeMNIST(len(g["yt"]), 1)
    kw = dict(mc_samples=S, num_epochs=0, data_minibatch=int(g["B"]), D=784, N=len(tr), inner_it=T, trainer="nested",
              log_every=10, lr0u=1e-4, lr0net=float(g["lr0net"]), lr0v=1e-3, init_args="subsample", init_sd=1e-2,
              num_pseudo=int(g["M"]), seed=0, architecture="lenet", n_hidden=0, n_layers=1, logistic_regression=False,
              train_dataset=tr, test_dataset=te, dnm="MNIST", nc=10, compute_weights_entropy=True, register_elbos=False,
              quiet=True)
    obj = PSVILearnV(**kw)
    obj.run_psvi(**kw)
    assert isinstance(obj.model, MeanFieldLeNet)
    names = [n for n, _ in obj.model.named_parameters()]
    assert names[:4] == ["0.weight", "0.bias", "0._weight_sd", "0._bias_sd"] and len(names) == 20
    assert tuple(obj.u.shape) == (int(g["M"]), 1, 28, 28)            # image pseudo-data, as in the reference
    torch.nn.utils.vector_to_parameters(torch.as_tensor(g["phi0"]).float().cuda(), obj.model.parameters())
    with torch.no_grad():
        obj.u.copy_(torch.as_tensor(g["u0"]).reshape(obj.u.shape))
        obj.v.copy_(torch.as_tensor(g["v0"]))
    obj.z = torch.as_tensor(g["z"]).float().cuda()
    obj.scheduler_optim_net = None
    obj.noise_source = ExternalNoise(eps)
    return obj


def test_lenet_psvi_methods_match_reference():
    g = dict(np.load(os.path.join(GOLDEN, "lenet_m10.npz")))
    S, T, M = int(g["S"]), int(g["T"]), int(g["M"])
    eps = LeNetNoiseFeeder.stream(S, int(g["noise_seed"]), int(g["n_forwards"]))
    obj = make_lenet_obj(g, S, T, eps)
    # the synthetic test set reaches the device through the dataset's own transform, like the reference's loaders
    xt_dev, _ = obj._device_dataset(obj.test_dataset, "test")
    np.testing.assert_allclose(xt_dev.cpu().numpy(), g["xt"], atol=1e-6)
    B = int(g["B"])
    xb = torch.as_tensor(g["xb"]).float().cuda().reshape(B, 1, 28, 28)
    yb = torch.as_tensor(g["yb"]).cuda()
    assert abs(obj.inner_elbo(model=obj.model).item() - g["ref64_inner_val"]) <= 1e-4 * abs(g["ref64_inner_val"])
    assert rel_l2(_to_phi(obj._last_inner), g["ref64_inner_gparams"]) < 2e-4
    assert abs(obj.psvi_elbo(xb, yb, model=obj.model).item() - g["ref64_outer_val"]) <= 1e-4 * abs(g["ref64_outer_val"])
    assert rel_l2(_to_phi(obj._last_outer["phi_grad"]), g["ref64_outer_gparams"]) < 5e-4
    assert rel_l2(obj._last_outer["u_grad"].cpu().numpy(), g["ref64_outer_gu"]) < 5e-4
    loss = obj.nested_step(xb, yb)
    assert abs(loss.item() - g["ref64_nested_loss"]) <= 2e-4 * abs(g["ref64_nested_loss"])
    assert tuple(obj.u.grad.shape) == (M, 1, 28, 28)
    assert rel_l2(obj.u.grad.cpu().numpy().reshape(M, 784), g["ref64_nested_gu"]) < 2e-3
    assert rel_l2(obj.v.grad.cpu().numpy(), g["ref64_nested_gv"]) < 2e-3
    vec = torch.nn.utils.parameters_to_vector(obj.model.parameters()).detach().cpu().numpy()
    assert rel_l2(vec, g["ref64_nested_params"]) < 1e-5
    np.testing.assert_allclose(obj.u.detach().cpu().numpy().reshape(M, 784), g["ref64_nested_u_after"], atol=2e-6)
    acc, nll, went, ness, vent = obj.evaluate()
    ref = g["ref64_eval"]
    assert abs(acc.item() - ref[0]) <= 1.0 / len(g["yt"]) + 1e-6
    np.testing.assert_allclose([nll.item(), ness.item()], ref[[1, 3]], rtol=3e-3)
    assert abs(went.item() - ref[2]) < 1e-5
    # module-level forward: [S, R, 10] logits, last layer's cached sample shared by all samples (Q4)
    lg = obj.model(xb)
    assert lg.shape == (S, B, 10) and torch.isfinite(lg).all()
    assert obj.model[-1]._cached_weight.shape == (10, 84) and obj.model[0]._cached_weight.shape == (S, 6, 1, 5, 5)


def _to_phi(g_mu_rho):
    """[dmu | drho] in theta layout -> torch parameters_to_vector order of the lenet modules."""
    v = g_mu_rho.detach().cpu().numpy()
    return lo.LeNetMeanField().join(v[:lo.P], v[lo.P:])


def test_lenet_psvi_run_learns_synthetic_digits():
    """End to end through run_psvi (reference :761-1028): 31 outer steps of psvi_learn_v with lenet on the MNIST-shaped
    synthetic set take the test accuracy from chance to > 0.9 (in-kernel Philox noise, no injected quantities)."""
    from psvi.inference.psvi_classes import PSVILearnV
    from tests.fake_mnist import FakeMNIST
    tr, te = FakeMNIST(1000, 0), FakeMNIST(300, 1)
    kw = dict(mc_samples=8, num_epochs=31, data_minibatch=128, D=784, N=len(tr), inner_it=10, trainer="nested", log_every=10,
              lr0u=1e-3, lr0net=1e-3, lr0v=1e-2, init_args="subsample", init_sd=1e-3, num_pseudo=50, seed=0,
              architecture="lenet", n_hidden=0, n_layers=1, logistic_regression=False, train_dataset=tr, test_dataset=te,
              dnm="MNIST", nc=10, compute_weights_entropy=True, register_elbos=False, quiet=True)
    res = PSVILearnV(**kw).run_psvi(**kw)
    accs = [float(a) for a in res["accs"]]
    assert accs[0] < 0.3 and accs[-1] > 0.9, accs
    assert all(np.isfinite(float(x)) for x in res["nlls"])


@pytest.mark.parametrize("S,R", [(1, 1), (2, 11), (3, 6)])
def test_lenet_pass_writes_inside_its_outputs_only(S, R):
    """Guard regions around every output of the gradient and dual passes stay untouched for row counts that leave partial
    image groups / row chunks / 32-wide tiles in every kernel (R = 1, 11, 6); NaN-filled outputs are fully overwritten."""
    from psvi import _native as nat
    nat.require_cuda()
    theta, thetad, X, y, cw = _case(S, R, 5)
    P, G = theta.shape[1], 1024
    th, thd, x_, y_, cw_ = dev(theta), dev(thetad), dev(X), dev(y, torch.int32), dev(cw)

    def guarded(*shape):
        n = int(np.prod(shape))
        buf = torch.full((n + 2 * G,), 12345.0, device="cuda")
        view = buf[G:G + n].view(*shape)
        view.fill_(float("nan"))
        return buf, view
    outs = {k: guarded(*shp) for k, shp in dict(nll=(S, R), tbar=(S, P), xbar=(S, R, 784)).items()}
    nat.lenet_pass(S, th, None, x_, y_, cw_, **{k: v[1] for k, v in outs.items()})
    outs2 = {k: guarded(*shp) for k, shp in dict(tbar=(S, P), tdbar=(S, P), xbar=(S, R, 784), acbar=(S, R)).items()}
    nat.lenet_pass(S, th, thd, x_, y_, cw_, **{k: v[1] for k, v in outs2.items()})
    torch.cuda.synchronize()
    for group in (outs, outs2):
        for k, (buf, view) in group.items():
            assert torch.all(buf[:G] == 12345.0) and torch.all(buf[-G:] == 12345.0), k
            assert torch.isfinite(view).all(), k
