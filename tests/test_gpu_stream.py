"""Streaming ("medium regime") path: models whose parameters do not fit the shared-memory-resident engine -- here the
baselines' fn with two 100-unit hidden layers (P = 10 602, experiments_utils.py:346-371) -- against the oracle."""
import numpy as np
import pytest
import torch

from oracle import psvi_oracle as po
from tests.gpu_util import dev, rel_l2, zeros

pytestmark = pytest.mark.gpu


def setup(dims, S, M, seed=0):
    rng = np.random.default_rng(seed)
    P = po.p_theta(dims)
    mu = (0.2 * rng.standard_normal(P)).astype(np.float32)
    rho = np.full(P, po.inverse_softplus(1e-2), np.float32)
    x = rng.standard_normal((M, dims[0])).astype(np.float32)
    y = rng.integers(0, dims[-1], M)
    return rng, P, mu, rho, x, y


@pytest.mark.parametrize("dims,S,M", [([2, 100, 100, 2], 10, 50), ([5, 64, 64, 64, 3], 4, 37)])
def test_unroll_stream_matches_oracle(dims, S, M):
    from psvi import _native as nat
    nat.require_cuda()
    rng, P, mu, rho, x, y = setup(dims, S, M)
    T, lr, scale = 4, 1e-3, 16.0
    eps = rng.standard_normal((T, S, P)).astype(np.float32)
    model = nat.make_model(dims, S)
    dmu, drho, am, av = dev(mu), dev(rho), zeros(2 * P), zeros(2 * P)
    losses = zeros(T)
    roww = torch.full((M,), scale, device="cuda")
    # the engine must refuse this size, the binding then streams
    rc = nat.lib().psvi_mf_unroll(__import__("ctypes").byref(model), __import__("ctypes").byref(nat.make_noise(dev(eps))),
                                  dmu.data_ptr(), drho.data_ptr(), am.data_ptr(), av.data_ptr(), 0, dev(x).data_ptr(),
                                  dev(y, torch.int32).data_ptr(), roww.data_ptr(), None, M, 1.0, 0, 0.0, 0, lr, 1, None, None)
    assert rc == nat.ERR_UNSUPPORTED
    nat.unroll(model, nat.make_noise(dev(eps)), dmu, drho, am, av, 0, dev(x), dev(y, torch.int32), roww, None, 1.0, 0, 0.0,
               T, lr, nat.ADAM_TORCH, losses)
    torch.cuda.synchronize()
    m64, r64 = mu.astype(np.float64), rho.astype(np.float64)
    m = np.zeros(2 * P); v = np.zeros(2 * P)
    ref_losses = []
    for t in range(T):
        val, gmu, grho = po.mfvi_grad(m64, r64, eps[t].astype(np.float64), x.astype(np.float64), y, scale, dims)
        phi, m, v = po.torch_adam_step(np.concatenate([m64, r64]), np.concatenate([gmu, grho]), m, v, t + 1, lr)
        m64, r64 = phi[:P], phi[P:]
        ref_losses.append(val)
    np.testing.assert_allclose(losses.cpu().numpy(), ref_losses, rtol=5e-5)
    assert rel_l2(dmu.cpu().numpy(), m64) < 1e-5
    assert rel_l2(drho.cpu().numpy(), r64) < 1e-5
    assert rel_l2(am.cpu().numpy(), m) < 2e-4


@pytest.mark.parametrize("mode", [0, 2])
def test_evaluate_stream_matches_oracle(mode):
    from psvi import _native as nat
    nat.require_cuda()
    dims, S, M = [2, 100, 100, 2], 8, 20
    rng, P, mu, rho, u, z = setup(dims, S, M, seed=3)
    n_rows, batch, N = 300, 128, 800.0
    nb = -(-n_rows // batch)
    eps = rng.standard_normal((nb, S, P)).astype(np.float32)
    xt = rng.standard_normal((n_rows, 2)).astype(np.float32)
    yt = rng.integers(0, 2, n_rows)
    v = (0.3 * rng.standard_normal(M)).astype(np.float32)
    model = nat.make_model(dims, S)
    out = zeros(8)
    scratch = zeros(nat.eval_scratch_floats(model, n_rows, batch))
    nat.evaluate(model, nat.make_noise(dev(eps)), dev(mu), dev(rho), dev(u), dev(z, torch.int32), dev(v), dev(xt),
                 dev(yt, torch.int32), batch, 0, N, 1, 0.0, mode, out, scratch)
    torch.cuda.synchronize()
    o = out.cpu().numpy()
    m64, r64 = mu.astype(np.float64), rho.astype(np.float64)
    if mode == 0:
        a = po.coreset_weights(v.astype(np.float64), N, 1)
        acc, nll, went, ness = po.evaluate(m64, r64, eps.astype(np.float64), u.astype(np.float64), z, a,
                                           xt.astype(np.float64), yt, dims, batch)
        np.testing.assert_allclose([o[3], o[4]], [went, ness], rtol=2e-3, atol=1e-5)
    else:
        c = nl = 0.0
        for k in range(nb):
            ck, nk = po.mfvi_predict(m64, r64, eps[k].astype(np.float64), xt[k * batch:(k + 1) * batch].astype(np.float64),
                                     yt[k * batch:(k + 1) * batch], dims)
            c += ck; nl += nk
        acc, nll = c / n_rows, nl / n_rows
    assert o[2] == n_rows
    assert abs(o[1] / o[2] - acc) <= 1.0 / n_rows + 1e-6
    np.testing.assert_allclose(o[0] / o[2], nll, rtol=2e-4)


def test_run_mfvi_subset_n_hidden_100():
    """BASELINE config 2: mfvi_subset with --n_hidden 100 (two hidden layers in the baselines' set_up_model)."""
    from psvi.experiments.experiments_utils import read_dataset
    from psvi.inference.baselines import run_mfvi_subset
    x, y, xt, yt, N, D, tr, te, nc = read_dataset("halfmoon", {"test_ratio": 0.2})
    res = run_mfvi_subset(x=x, y=y, xt=xt, yt=yt, mc_samples=10, data_minibatch=128, num_epochs=150, log_every=100, D=D,
                          lr0net=1e-2, seed=0, train_dataset=tr, test_dataset=te, num_pseudo=50, init_args="subsample",
                          architecture="fn", n_hidden=100, nc=nc, dnm="halfmoon", init_sd=1e-3, quiet=True)
    assert len(res["elbos"]) == 300 and len(res["accs"]) == 3
    assert np.isfinite(res["nlls"]).all() and res["accs"][-1] > 0.8 and res["elbos"][-1] > res["elbos"][0]


def test_stream_outer_grad_row_chunks_and_rank_shares_add_up():
    """StreamEngine.outer_grad: (i) processing the data rows in chunks equals one concatenated pass; (ii) the shares of R
    emulated ranks (kappa = 1/R, contiguous row shards of the same minibatch) add up to the unsharded value and gradients --
    the identity behind the ONE all-reduce of the sharded bilevel step."""
    from psvi.inference.stream import MeanFieldFamily, StreamEngine
    from psvi.models.neural_net import make_fcnet
    torch.manual_seed(0)
    D, H, C, S, M, B = 6, 40, 3, 5, 12, 50
    net = make_fcnet(D, H, C, n_layers=1, mc_samples=S, init_sd=1e-2).cuda()
    net.flat()
    eng = StreamEngine(MeanFieldFamily(net), net.dims, S)
    g = torch.Generator(device="cuda").manual_seed(1)
    u, xb = torch.randn(M, D, device="cuda", generator=g), torch.randn(B, D, device="cuda", generator=g)
    z = torch.randint(0, C, (M,), device="cuda", generator=g, dtype=torch.int32)
    yb = torch.randint(0, C, (B,), device="cuda", generator=g, dtype=torch.int32)
    a = torch.rand(M, device="cuda", generator=g) * 40
    eps = torch.randn(S, eng.Pt, device="cuda", generator=g)
    phi = eng.fam.get_phi()
    full = eng.outer_grad(phi, eps, u, z, a, xb, yb, 500.0)
    eng.ROW_CHUNK = 16
    chunked = eng.outer_grad(phi, eps, u, z, a, xb, yb, 500.0)
    for x1, x2 in zip(full[:4], chunked[:4]):
        assert rel_l2(x2.cpu().numpy(), x1.cpu().numpy()) < 2e-5
    for R in (2, 3):
        for chunk in (8192, 16):
            eng.ROW_CHUNK = chunk
            parts = []
            for r in range(R):
                lo, hi = (B * r) // R, (B * (r + 1)) // R
                parts.append(eng.outer_grad(phi, eps, u, z, a, xb[lo:hi], yb[lo:hi], 500.0, kappa=1.0 / R, n_total=B))
            for i in range(4):
                tot = sum(p[i].double() for p in parts)
                assert rel_l2(tot.cpu().numpy(), full[i].double().cpu().numpy()) < 5e-5


@pytest.mark.parametrize("masked", [False, True])
def test_fused_meanfield_family_kernels_match_formulas(masked):
    """csrc/psvi_family.cu (sample / tangent / nkl / reparameterisation gradient and HVP tails over [S][P] slabs) against the
    closed forms of SURVEY Appendix A.1 / A.6 written with torch ops in fp64."""
    from psvi import _native as nat
    import torch.nn.functional as F
    nat.require_cuda()
    g = torch.Generator(device="cuda").manual_seed(3)
    S, P = 5, 1003
    mu, rho = torch.randn(P, device="cuda", generator=g) * 0.3, torch.randn(P, device="cuda", generator=g) - 3.0
    mud, rhod = torch.randn(P, device="cuda", generator=g), torch.randn(P, device="cuda", generator=g)
    eps, tbar = torch.randn(S, P, device="cuda", generator=g), torch.randn(S, P, device="cuda", generator=g)
    Atd, beta = torch.randn(S, P, device="cuda", generator=g), torch.randn(S, device="cuda", generator=g)
    mask = (torch.rand(P, device="cuda", generator=g) > 0.4).float() if masked else None
    mk = (mask if masked else torch.ones(P, device="cuda")).double()
    m64, r64, e64, t64 = mu.double(), rho.double(), eps.double(), tbar.double()
    sg, sig = F.softplus(r64), torch.sigmoid(r64)
    theta = nat.mf_sample(mu, rho, eps)
    assert rel_l2(theta.cpu().numpy(), (m64 + sg * e64).cpu().numpy()) < 1e-6
    thd = nat.mf_tangent(rho, mud, rhod, eps)
    assert rel_l2(thd.cpu().numpy(), (mud.double() + sig * rhod.double() * e64).cpu().numpy()) < 1e-6
    th64 = theta.double()
    out = nat.mf_nkl_kl(mu, rho, eps, theta, mask)
    nkl = (mk * (-0.5 * th64 ** 2 + 0.5 * e64 ** 2 + torch.log(sg))).sum(1)
    kl = (mk * (0.5 * (sg ** 2 + m64 ** 2 - 1) - torch.log(sg))).sum()
    np.testing.assert_allclose(out[:S].cpu().numpy(), nkl.cpu().numpy(), rtol=1e-5)
    np.testing.assert_allclose(out[S].item(), kl.item(), rtol=1e-5)
    for kc, nc_, use_beta in ((1.0, 0.0, False), (0.0, 0.7, True)):
        gg = nat.mf_reparam_grad(mu, rho, eps, tbar, kc, nc_, mask=mask, beta=beta if use_beta else None,
                                 theta=theta if use_beta else None)
        tb = t64 - (beta.double()[:, None] * mk * th64 if use_beta else 0.0)
        ref = torch.cat([tb.sum(0) + mk * kc * m64, sig * ((tb * e64).sum(0) + mk * (kc * (sg - 1 / sg) + nc_ / sg))])
        assert rel_l2(gg.cpu().numpy(), ref.cpu().numpy()) < 2e-6
    h = nat.mf_reparam_hvp(rho, mud, rhod, eps, tbar, Atd, mask=mask)
    rd = rhod.double()
    ref = torch.cat([t64.sum(0) + mk * mud.double(),
                     sig * (t64 * e64).sum(0) + sig * (1 - sig) * rd * (Atd.double() * e64).sum(0)
                     + mk * ((1 + 1 / sg ** 2) * sig ** 2 + (sg - 1 / sg) * sig * (1 - sig)) * rd])
    assert rel_l2(h.cpu().numpy(), ref.cpu().numpy()) < 2e-6


@pytest.mark.gpu
@pytest.mark.parametrize("t", [1, 7, 100])
def test_fused_unrolled_adam_kernels_match_the_tensor_expressions(t):
    """psvi_adam_unroll_step / _reverse against the reference's tensor expressions (robust_higher/optim.py:303-367 and what
    autograd replays through them), evaluated in fp32 by torch op by op and in float64 by autograd."""
    import math
    from psvi import _native
    torch.manual_seed(t)
    n, lr, B1, B2 = 5003, 1e-3, 0.9, 0.999
    phi, g = torch.randn(n, device="cuda"), torch.randn(n, device="cuda") * 10 ** torch.randint(-6, 2, (n,), device="cuda").float()
    m, v = torch.randn(n, device="cuda") * 0.1, torch.rand(n, device="cuda") * 0.01
    g[:17], v[:17] = 0.0, 0.0          # exercises the v' == 0 mask of the reverse sweep
    k, sq2 = lr / (1.0 - B1 ** t), math.sqrt(1.0 - B2 ** t)
    # forward, fp32 op by op (the order stream.py used before the fused kernel)
    m_ref = m * B1 + float(1.0 - B1) * g
    v_ref = v * B2 + float(1.0 - B2) * g * g
    phi_ref = phi - k * (m_ref / (torch.sqrt(v_ref + 1e-8) / sq2 + 1e-8))
    phi_n, m_n, v_n = _native.adam_unroll_step(phi, g, m, v, k, sq2)
    assert torch.equal(m_n, m_ref) and torch.equal(v_n, v_ref)
    # (ATen divides by a host scalar as a multiplication by its reciprocal: last-ulp differences in the denominator)
    torch.testing.assert_close(phi_n, phi_ref, rtol=5e-6, atol=1e-7)
    # reverse against float64 autograd of the same step
    pbar, mbar, vbar = torch.randn(n, device="cuda"), torch.randn(n, device="cuda"), torch.randn(n, device="cuda") * 100
    gd = g.double().requires_grad_(True)
    md = m.double() * B1 + (1.0 - B1) * gd
    vd = v.double() * B2 + (1.0 - B2) * gd * gd
    vd.register_hook(lambda gr: torch.where(vd.detach() == 0, torch.zeros_like(gr), gr))
    phid = phi.double() - k * (md / (torch.sqrt(vd + 1e-8) / sq2 + 1e-8))
    mdn, vdn = md * 1.0, vd * 1.0
    (gbar_ref,) = torch.autograd.grad([phid, mdn, vdn], [gd], [pbar.double(), mbar.double(), vbar.double()])
    mb_in, vb_in = mbar.clone(), vbar.clone()
    gbar = _native.adam_unroll_reverse(pbar, g, m_n, v_n, mb_in, vb_in, k, sq2)
    assert rel_l2(gbar.cpu().numpy(), gbar_ref.cpu().numpy()) < 2e-6
    torch.testing.assert_close(gbar[:17], float(1.0 - B1) * (mbar[:17] - k * pbar[:17] / (math.sqrt(1e-8) / sq2 + 1e-8)), rtol=1e-5,
                               atol=0)
    torch.testing.assert_close(vb_in[:17], torch.zeros(17, device="cuda"), rtol=0, atol=0)


@pytest.mark.parametrize("dims,S,R", [([6, 8, 8, 1], 4, 37), ([13, 50, 50, 1], 5, 700)])
def test_net_pass_gaussian_matches_oracle(dims, S, R):
    """psvi_net_pass_gaussian (the regressors' likelihood, reference psvi_classes.py:1986,2034-2057): value, gradient and
    forward-over-reverse flavours on sampled weights theta [S][P] against the fp64 oracle -- nll, weight adjoints, input adjoints,
    the mixed row-weight term, and d/dy (gradient pass) / its directional derivative (dual pass).  R = 700 spans several row
    chunks of the kernel.  fp32 kernel: 2e-5 relative."""
    from psvi import _native as nat
    nat.require_cuda()
    rng = np.random.default_rng(S + R)
    P, tau = po.p_theta(dims), 0.7
    theta = (0.3 * rng.standard_normal((S, P))).astype(np.float32)
    thetad = (0.3 * rng.standard_normal((S, P))).astype(np.float32)
    X = rng.standard_normal((R, dims[0])).astype(np.float32)
    y = rng.standard_normal(R).astype(np.float32)
    cw = rng.uniform(0.5, 1.5, (S, R)).astype(np.float32)
    model = nat.make_model(dims, S)
    th, thd, x_, y_, cw_ = dev(theta), dev(thetad), dev(X), dev(y), dev(cw)
    t64, td64, X64, y64, cw64 = (a.astype(np.float64) for a in (theta, thetad, X, y, cw))
    # forward: outputs + nll
    out, nll = zeros(S, R, 1), zeros(S, R)
    nat.net_pass_gaussian(model, th, None, x_, y_, None, tau, nll=nll, outputs=out)
    o, cache = po.mlp_forward(t64, X64, dims)
    ref_nll, r = po.gauss_nll_rows(o, y64, tau)
    assert rel_l2(out.cpu().numpy(), o) < 2e-5 and rel_l2(nll.cpu().numpy(), ref_nll) < 2e-5
    # gradient pass
    tbar, xbar, ybar = zeros(S, P), zeros(S, R, dims[0]), zeros(S, R)
    nat.net_pass_gaussian(model, th, None, x_, y_, cw_, tau, nll=nll, tbar=tbar, xbar=xbar, ybar=ybar)
    tb, xb = po.mlp_backward(t64, cache, dims, (cw64 * tau * r)[..., None])
    assert rel_l2(tbar.cpu().numpy(), tb) < 2e-5 and rel_l2(xbar.cpu().numpy(), xb) < 2e-5
    assert rel_l2(ybar.cpu().numpy(), -cw64 * tau * r) < 2e-5
    # dual pass
    tbar, tdbar, xbar, ac, ybar = zeros(S, P), zeros(S, P), zeros(S, R, dims[0]), zeros(S, R), zeros(S, R)
    nat.net_pass_gaussian(model, th, thd, x_, y_, cw_, tau, tbar=tbar, tdbar=tdbar, xbar=xbar, acbar=ac, ybar=ybar)
    o, od, cache2 = po.mlp_dual_forward(t64, td64, X64, dims)
    d = od[..., 0]
    A_t, A_td, A_x = po.mlp_dual_backward(t64, td64, cache2, dims, (cw64 * tau * d)[..., None], (cw64 * tau * r)[..., None])
    assert rel_l2(tbar.cpu().numpy(), A_t) < 5e-5 and rel_l2(tdbar.cpu().numpy(), A_td) < 5e-5
    assert rel_l2(xbar.cpu().numpy(), A_x) < 5e-5
    assert rel_l2(ac.cpu().numpy(), tau * r * d) < 5e-5 and rel_l2(ybar.cpu().numpy(), -cw64 * tau * d) < 5e-5


@pytest.mark.parametrize("dims,S,R", [([2, 40, 40, 2], 32, 100), ([5, 24, 16, 3], 6, 229), ([7, 12, 3], 3, 65)])
def test_net_pass_row_split_over_a_cluster(dims, S, R, monkeypatch):
    """psvi_net_pass splits the rows of a sample over a thread-block cluster (Z = R / 16 <= 8 CTAs, partial weight adjoints
    summed through distributed shared memory in fixed rank order): gradient and dual passes equal the one-CTA-per-sample form
    (PSVI_NET_PASS_Z=1) to fp32 summation-order accuracy, match the fp64 oracle, and are bit-reproducible run to run."""
    from psvi import _native as nat
    nat.require_cuda()
    rng = np.random.default_rng(S + R)
    P = po.p_theta(dims)
    theta = (0.3 * rng.standard_normal((S, P))).astype(np.float32)
    thetad = (0.3 * rng.standard_normal((S, P))).astype(np.float32)
    X = rng.standard_normal((R, dims[0])).astype(np.float32)
    y = rng.integers(0, dims[-1], R)
    cw = rng.uniform(0.5, 1.5, (S, R)).astype(np.float32)
    model = nat.make_model(dims, S)
    th, thd, x_, y_, cw_ = dev(theta), dev(thetad), dev(X), dev(y, torch.int32), dev(cw)

    def run():
        nll, tb1, xb1 = zeros(S, R), zeros(S, P), zeros(S, R, dims[0])
        nat.net_pass(model, th, None, x_, y_, cw_, nll=nll, tbar=tb1, xbar=xb1)
        tb2, tdb2, xb2, ac = zeros(S, P), zeros(S, P), zeros(S, R, dims[0]), zeros(S, R)
        nat.net_pass(model, th, thd, x_, y_, cw_, tbar=tb2, tdbar=tdb2, xbar=xb2, acbar=ac)
        torch.cuda.synchronize()
        return [t.cpu().numpy() for t in (nll, tb1, xb1, tb2, tdb2, xb2, ac)]

    split = run()
    again = run()
    for a, b in zip(split, again):
        np.testing.assert_array_equal(a, b)
    monkeypatch.setenv("PSVI_NET_PASS_Z", "1")
    single = run()
    for a, b in zip(split, single):
        assert rel_l2(a, b) < 2e-6
    t64, td64, X64, cw64 = (a.astype(np.float64) for a in (theta, thetad, X, cw))
    o, cache = po.mlp_forward(t64, X64, dims)
    ref_nll, p = po.nll_rows(o, y)
    q = p.copy()
    q[:, np.arange(R), y] -= 1.0
    At, Ax = po.mlp_backward(t64, cache, dims, cw64[:, :, None] * q)
    assert rel_l2(split[0], ref_nll) < 2e-5 and rel_l2(split[1], At) < 2e-5 and rel_l2(split[2], Ax) < 2e-5
