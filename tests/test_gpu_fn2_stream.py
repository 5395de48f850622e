"""Streaming path on the GPU: full-covariance fn2 (reference neural_net.py:408-524) against the reference's fp64 outputs
(tests/golden/fn2_hm_h6.npz) and the generic oracle; and the streaming mean-field path against the fused engine."""
import os

import numpy as np
import pytest
import torch

from oracle import psvi_oracle as po
from oracle import psvi_oracle_generic as pg
from oracle.ref_import import NoiseFeeder
from tests.gpu_util import GOLDEN, dev, load, rel_l2, zeros

pytestmark = pytest.mark.gpu


def test_fc_matvec_and_outer_match_numpy():
    from psvi import _native as nat
    nat.require_cuda()
    rng = np.random.default_rng(0)
    for n, S in ((1, 3), (2, 4), (3, 5), (42, 7), (300, 33)):
        nc = max((n - 1) * (n - 2) // 2, 0)
        base, dg = rng.standard_normal(n), rng.random(n) + 0.1
        off = rng.standard_normal(nc)
        eps = rng.standard_normal((S, n + 5))       # leading dimension larger than n on purpose
        fam = pg.FullCov([1, 1]); L = np.zeros((n, n)); L[np.arange(n), np.arange(n)] = dg
        if nc:
            r, c = fam.tril_idx(n); L[r, c] = off
        ref = base[None] + eps[:, 2:2 + n] @ L.T
        e_t = dev(eps); out = zeros(S, n + 3)
        nat.fc_matvec(n, S, dev(base), dev(dg), dev(off) if nc else dev(np.zeros(1)), e_t.data_ptr() + 8, n + 5,
                      out.data_ptr() + 4, n + 3)
        torch.cuda.synchronize()
        np.testing.assert_allclose(out.cpu().numpy()[:, 1:1 + n], ref, rtol=2e-5, atol=2e-5)
        A = rng.standard_normal((S, n + 2))
        a_t = dev(A); gb, gd, go = zeros(n), zeros(n), zeros(max(nc, 1))
        nat.fc_outer(n, S, a_t.data_ptr(), n + 2, e_t.data_ptr() + 8, n + 5, gb, gd, go)
        torch.cuda.synchronize()
        G = A[:, :n].T @ eps[:, 2:2 + n]
        np.testing.assert_allclose(gb.cpu().numpy(), A[:, :n].sum(0), rtol=2e-5, atol=2e-5)
        np.testing.assert_allclose(gd.cpu().numpy(), np.diag(G), rtol=2e-5, atol=2e-5)
        if nc:
            np.testing.assert_allclose(go.cpu().numpy()[:nc], G[r, c], rtol=2e-5, atol=2e-5)


def make_fn2_obj(g, dims, S, T, eps):
    from psvi.experiments.experiments_utils import read_dataset
    from psvi.inference.psvi_classes import ExternalNoise, PSVILearnV
    x, y, xt, yt, N, D, tr, te, nc = read_dataset("halfmoon", {"test_ratio": 0.2})
    kw = dict(mc_samples=S, num_epochs=0, data_minibatch=int(g["B"]), D=D, N=N, inner_it=T, trainer="nested", log_every=10,
              lr0u=1e-4, lr0net=float(g["lr0net"]), lr0v=1e-3, init_args="subsample", init_sd=0.05, num_pseudo=int(g["M"]),
              seed=0, architecture="fn2", n_hidden=dims[1], n_layers=1, logistic_regression=False, train_dataset=tr,
              test_dataset=te, dnm="halfmoon", nc=nc, compute_weights_entropy=True, register_elbos=False, quiet=True)
    obj = PSVILearnV(**kw)
    obj.run_psvi(**kw)
    assert obj.model.dims == dims
    # the reference's parameter names / sizes
    names = [n for n, _ in obj.model.named_parameters()]
    assert names[:3] == ["lin0.mean", "lin0._sd", "lin0._corr"]
    torch.nn.utils.vector_to_parameters(torch.as_tensor(g["phi0"]).float().cuda(), obj.model.parameters())
    with torch.no_grad():
        obj.u.copy_(torch.as_tensor(g["u0"]))
        obj.v.copy_(torch.as_tensor(g["v0"]))
    obj.z = torch.as_tensor(g["z"]).float().cuda()
    obj.scheduler_optim_net = None
    obj.noise_source = ExternalNoise(eps)
    return obj


def test_fn2_psvi_methods_match_reference():
    g = dict(np.load(os.path.join(GOLDEN, "fn2_hm_h6.npz")))
    dims = [int(d) for d in g["dims"]]
    S, T = int(g["S"]), int(g["T"])
    eps = NoiseFeeder.stream(dims, S, int(g["noise_seed"]), int(g["n_forwards"]))
    obj = make_fn2_obj(g, dims, S, T, eps)
    xb, yb = torch.as_tensor(g["xb"]).float().cuda(), torch.as_tensor(g["yb"]).cuda()
    assert abs(obj.inner_elbo(model=obj.model).item() - g["ref64_inner_val"]) <= 1e-4 * abs(g["ref64_inner_val"])
    assert rel_l2(obj._last_inner.cpu().numpy(), g["ref64_inner_gparams"]) < 2e-4
    assert abs(obj.psvi_elbo(xb, yb, model=obj.model).item() - g["ref64_outer_val"]) <= 1e-4 * abs(g["ref64_outer_val"])
    assert rel_l2(obj._last_outer["phi_grad"].cpu().numpy(), g["ref64_outer_gparams"]) < 5e-4
    assert rel_l2(obj._last_outer["u_grad"].cpu().numpy(), g["ref64_outer_gu"]) < 5e-4
    loss = obj.nested_step(xb, yb)
    assert abs(loss.item() - g["ref64_nested_loss"]) <= 2e-4 * abs(g["ref64_nested_loss"])
    assert rel_l2(obj.u.grad.cpu().numpy(), g["ref64_nested_gu"]) < 2e-3
    assert rel_l2(obj.v.grad.cpu().numpy(), g["ref64_nested_gv"]) < 2e-3
    vec = torch.nn.utils.parameters_to_vector(obj.model.parameters()).detach().cpu().numpy()
    assert rel_l2(vec, g["ref64_nested_params"]) < 1e-5
    np.testing.assert_allclose(obj.u.detach().cpu().numpy(), g["ref64_nested_u_after"], atol=2e-6)
    acc, nll, went, ness, vent = obj.evaluate()
    ref = g["ref64_eval"]
    assert abs(acc.item() - ref[0]) <= 1.0 / len(g["yt"]) + 1e-6
    np.testing.assert_allclose([nll.item(), went.item(), ness.item()], ref[1:4], rtol=3e-3, atol=1e-5)
    # module-level forward
    lg = obj.model(torch.as_tensor(g["xt"][:7]).float().cuda())
    assert lg.shape == (S, 7, dims[-1]) and torch.isfinite(lg).all()


@pytest.mark.parametrize("name", ["fn_hm_m50_t10", "fn_fb_l2_m13", "logreg_hm_m10"])
def test_stream_meanfield_nested_equals_fused(name):
    """The streaming mean-field path (medium-size models) on a model small enough for the fused engine too."""
    from tests.test_gpu_psvi_class import make_obj
    g, dims, S, T, eps = load(name)
    xb, yb = torch.as_tensor(g["xb"]).float().cuda(), torch.as_tensor(g["yb"]).cuda()
    res = []
    for force in (False, True):
        obj = make_obj(name, g, dims, S, T, eps)
        obj.noise_source.pos = 2
        if force:
            obj._ws[("force_stream", id(obj.model))] = True
        loss = obj.nested_step(xb, yb)
        vec = torch.nn.utils.parameters_to_vector(obj.model.parameters()).detach().cpu().numpy()
        res.append((loss.item(), obj.u.grad.cpu().numpy(), obj.v.grad.cpu().numpy() if obj.learn_v else np.zeros(1), vec))
    assert abs(res[0][0] - res[1][0]) <= 1e-4 * abs(res[0][0])
    assert rel_l2(res[1][1], res[0][1]) < 1e-3
    if int(g["vmode"]):
        assert rel_l2(res[1][2], res[0][2]) < 1e-3
    assert rel_l2(res[1][3], res[0][3]) < 1e-5
    assert rel_l2(res[1][1], g["ref64_nested_gu"]) < 1e-3


def test_medium_meanfield_nested_step_two_hidden_layers_of_100():
    """PSVI nested step on a model the fused engine refuses (P = 10 602): automatic fall-through to the streaming path,
    checked against the fp64 oracle."""
    from psvi.experiments.experiments_utils import read_dataset
    from psvi.inference.psvi_classes import ExternalNoise, PSVILearnV
    x, y, xt, yt, N, D, tr, te, nc = read_dataset("halfmoon", {"test_ratio": 0.2})
    S, M, T, B = 4, 12, 3, 32
    kw = dict(mc_samples=S, num_epochs=0, data_minibatch=B, D=D, N=N, inner_it=T, trainer="nested", log_every=10, lr0u=1e-4,
              lr0net=1e-3, lr0v=1e-3, init_args="subsample", init_sd=1e-2, num_pseudo=M, seed=0, architecture="fn",
              n_hidden=100, n_layers=2, logistic_regression=False, train_dataset=tr, test_dataset=te, dnm="halfmoon", nc=nc,
              compute_weights_entropy=True, register_elbos=False, quiet=True)
    obj = PSVILearnV(**kw)
    obj.run_psvi(**kw)
    dims = obj.model.dims
    assert dims == [2, 100, 100, 2]
    eps = NoiseFeeder.stream(dims, S, 31, T + 1)
    obj.noise_source = ExternalNoise(eps)
    obj.scheduler_optim_net = None
    mu, rho = [t.cpu().numpy().astype(np.float64) for t in obj.model.flat()]
    u0, z, v0 = obj.u.detach().cpu().numpy().astype(np.float64), obj.z.cpu().numpy(), obj.v.detach().cpu().numpy().astype(np.float64)
    xb, yb = x[:B].cuda(), y[:B].cuda()
    loss = obj.nested_step(xb, yb)
    assert obj._ws.get(("force_stream", id(obj.model))) is True
    e64 = [e.astype(np.float64) for e in eps]
    r = po.nested_step(mu, rho, np.stack(e64[:T]), e64[T], u0, z, v0, x[:B].numpy().astype(np.float64), y[:B].numpy(),
                       float(N), dims, 1e-3, vmode=1)
    assert abs(loss.item() - r["loss"]) <= 1e-4 * abs(r["loss"])
    assert rel_l2(obj.u.grad.cpu().numpy(), r["u_grad"]) < 2e-3
    assert rel_l2(obj.v.grad.cpu().numpy(), r["v_grad"]) < 2e-3


@pytest.mark.parametrize("n,S", [(5, 3), (70, 5), (150, 40), (257, 32)])
def test_fc_sample_and_reparam_match_numpy_and_stay_inside_their_outputs(n, S):
    """psvi_fc_sample / psvi_fc_reparam_grad / psvi_fc_reparam_hvp (the packed-triangle family maps of fn2, reference
    neural_net.py:408-491) against a dense fp64 restatement, for sizes that span one to five 64 x 64 tiles of the tiled rank-S
    update and sample counts on both sides of its 32-sample stage; guard bands around every output stay untouched."""
    from psvi import _native as nat
    nat.require_cuda()
    rng = np.random.default_rng(n + S)
    ntri = (n - 1) * (n - 2) // 2
    P = 2 * n + ntri
    phi = np.concatenate([rng.standard_normal(n), rng.standard_normal(n) - 1.0, 0.1 * rng.standard_normal(ntri)]).astype(np.float32)
    phid = (0.3 * rng.standard_normal(P)).astype(np.float32)
    eps = rng.standard_normal((S, n)).astype(np.float32)
    A = rng.standard_normal((S, n)).astype(np.float32)
    Ad = rng.standard_normal((S, n)).astype(np.float32)
    f64 = lambda a: a.astype(np.float64)
    mean, sd, corr = f64(phi[:n]), f64(phi[n:2 * n]), f64(phi[2 * n:])
    d, sig = np.log1p(np.exp(sd)), 1.0 / (1.0 + np.exp(-sd))

    def tril(v):   # strictly-lower entries of the top-left (n-1) x (n-1) block, row-major; the last row has none (quirk Q6)
        L = np.zeros((n, n))
        if n >= 3:
            r, c = np.tril_indices(n - 1, -1)
            L[r, c] = v
        return L

    G = 64
    guard = lambda size: torch.full((size + 2 * G,), 7.25, device="cuda")
    ok = lambda t, size: bool((t[:G] == 7.25).all() and (t[G + size:] == 7.25).all())
    phi_d, phid_d, eps_d, A_d, Ad_d = dev(phi), dev(phid), dev(eps), dev(A), dev(Ad)
    # sample and tangent
    out = guard(S * n)
    nat.fc_sample(n, S, phi_d.data_ptr(), None, eps_d.data_ptr(), n, out.data_ptr() + 4 * G, n)
    ref = mean + f64(eps) * d + f64(eps) @ tril(corr).T
    assert ok(out, S * n) and rel_l2(out[G:G + S * n].cpu().numpy().reshape(S, n), ref) < 1e-5
    out = guard(S * n)
    nat.fc_sample(n, S, phi_d.data_ptr(), phid_d.data_ptr(), eps_d.data_ptr(), n, out.data_ptr() + 4 * G, n)
    md, sdd, cd = f64(phid[:n]), f64(phid[n:2 * n]), f64(phid[2 * n:])
    ref = md + f64(eps) * (sig * sdd) + f64(eps) @ tril(cd).T
    assert ok(out, S * n) and rel_l2(out[G:G + S * n].cpu().numpy().reshape(S, n), ref) < 1e-5
    # gradient
    kl, nkl = 0.7, -0.3
    g = guard(P)
    nat.fc_reparam_grad(n, S, phi_d.data_ptr(), A_d.data_ptr(), n, eps_d.data_ptr(), n, kl, nkl, g.data_ptr() + 4 * G)
    AE = f64(A).T @ f64(eps)                                    # [r][c] = sum_s A[s][r] eps[s][c]
    rr, cc = np.tril_indices(n - 1, -1) if n >= 3 else (np.zeros(0, int), np.zeros(0, int))
    ref = np.concatenate([f64(A).sum(0) + kl * mean, sig * ((f64(A) * f64(eps)).sum(0) + kl * (d - 1 / d) + nkl / d),
                          AE[rr, cc] + kl * corr])
    assert ok(g, P) and rel_l2(g[G:G + P].cpu().numpy(), ref) < 1e-5
    # Hessian-vector product
    h = guard(P)
    nat.fc_reparam_hvp(n, S, phi_d.data_ptr(), phid_d.data_ptr(), A_d.data_ptr(), Ad_d.data_ptr(), n, eps_d.data_ptr(), n,
                       h.data_ptr() + 4 * G)
    ref = np.concatenate([f64(A).sum(0) + md,
                          sig * (f64(A) * f64(eps)).sum(0) + sig * (1 - sig) * sdd * (f64(Ad) * f64(eps)).sum(0)
                          + ((1 + 1 / d ** 2) * sig ** 2 + (d - 1 / d) * sig * (1 - sig)) * sdd,
                          AE[rr, cc] + cd])
    assert ok(h, P) and rel_l2(h[G:G + P].cpu().numpy(), ref) < 1e-5
