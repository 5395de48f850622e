"""Size-independent properties of the CUDA path at BASELINE.json's FULL sizes, where the fp64 oracle is too slow to be the
checker (cfg2: fn H=100 M=50 S=10 T=100; cfg3: fn2 [2,40,40,2] S=32; cfg4: lenet M=200 S=10; cfg5: fn D=256 H=1024 S=64 M=1000;
full-data passes over 0.5 M rows).  Each per-sample network pass (the three `net` plug-ins of the streaming engine) must satisfy,
for the loss L_s(theta) = sum_r cw[s, r] nll[s, r]:

  P1  the dual pass returns A_thetadot = grad L_s          (A_Wdot = Wbar), i.e. tdbar(dual) == tbar(grad)
  P2  forward-mode == reverse-mode directional derivative: sum_r cw[s, r] acbar[s, r] == <tbar[s], thetad[s]>
  P3  rows add up: a pass over all rows == the sum of passes over two row blocks (the identity the sharded step rests on)
  P4  the Hessian-vector product is linear in the direction
  P5  the Hessian is symmetric: <w, H v> == <v, H w>

and the full-data tensor-core passes must be invariant under splitting / permuting rows.  Tolerances are fp32 accumulation
order (1e-4 class) and, for the large regime, the tf32x3 arithmetic (1e-3 class; x10 more for the opt-in split-bf16
arithmetic); they are written next to each assert."""
import numpy as np
import pytest
import torch

from oracle import psvi_oracle as po
from tests.gpu_util import rel_l2

pytestmark = pytest.mark.gpu


def _rand_theta_mlp(dims, S, g):
    parts = []
    for din, dout in zip(dims[:-1], dims[1:]):
        parts += [torch.randn(S, dout * din, device="cuda", generator=g) / din ** 0.5,
                  0.1 * torch.randn(S, dout, device="cuda", generator=g)]
    return torch.cat(parts, 1).contiguous()


def _rand_theta_lenet(S, g):
    from oracle import lenet_oracle as lo
    parts = []
    for (_, wshape, nb, _) in lo.LAYERS:
        fan_in = int(np.prod(wshape[1:]))
        parts += [torch.randn(S, int(np.prod(wshape)), device="cuda", generator=g) / fan_in ** 0.5,
                  0.1 * torch.randn(S, nb, device="cuda", generator=g)]
    return torch.cat(parts, 1).contiguous()


def _nets():
    from psvi import _native
    from psvi.inference.stream import FnLargeNet, LenetNet, MlpNet
    return {
        # name: (net factory, theta factory, D_in, classes, S, rows, tolerance scale)
        "cfg2_fn": (lambda: MlpNet([2, 100, 2], 10), lambda g: _rand_theta_mlp([2, 100, 2], 10, g), 2, 2, 10, 178, 1.0),
        "cfg3_fn2_net": (lambda: MlpNet([2, 40, 40, 2], 32), lambda g: _rand_theta_mlp([2, 40, 40, 2], 32, g), 2, 2, 32, 228, 1.0),
        "cfg4_lenet": (lambda: LenetNet(10), lambda g: _rand_theta_lenet(10, g), 784, 10, 10, 328, 1.0),
        "cfg5_fn_large": (lambda: FnLargeNet([256, 1024, 10], 64), lambda g: _rand_theta_mlp([256, 1024, 10], 64, g), 256, 10, 64,
                          1000, 10.0),
        "cfg5_fn_large_bf16x3": (lambda: FnLargeNet([256, 1024, 10], 64, precision=_native.PREC_BF16X3),
                                 lambda g: _rand_theta_mlp([256, 1024, 10], 64, g), 256, 10, 64, 1000, 100.0),
    }


def _passes(net, theta, thetad, x, y, cw, want_x=False):
    S, P, R = theta.shape[0], theta.shape[1], x.shape[0]
    z = lambda *s: torch.zeros(*s, device="cuda")
    out = dict(nll=z(S, R), tbar=z(S, P))
    if want_x:
        out["xbar"] = z(S, R, x.shape[1])
    if thetad is not None:
        out.update(tdbar=z(S, P), acbar=z(S, R))
        out.setdefault("xbar", z(S, R, x.shape[1]))
        out.pop("nll")
    net.pass_(theta, thetad, x, y, cw, **out)
    torch.cuda.synchronize()
    return out


@pytest.mark.parametrize("name", ["cfg2_fn", "cfg3_fn2_net", "cfg4_lenet", "cfg5_fn_large", "cfg5_fn_large_bf16x3"])
def test_network_pass_properties_at_full_size(name):
    from psvi import _native
    _native.require_cuda()
    mk_net, mk_theta, D, C, S, R, tol = _nets()[name]
    g = torch.Generator(device="cuda").manual_seed(len(name))
    net, theta = mk_net(), mk_theta(g)
    scale = theta.abs().mean()
    v = (torch.randn(theta.shape, device="cuda", generator=g) * scale).contiguous()
    w = (torch.randn(theta.shape, device="cuda", generator=g) * scale).contiguous()
    x = torch.randn(R, D, device="cuda", generator=g).contiguous()
    y = torch.randint(0, C, (R,), device="cuda", generator=g, dtype=torch.int32)
    cw = (0.5 + torch.rand(S, R, device="cuda", generator=g)).contiguous()

    grad = _passes(net, theta, None, x, y, cw)
    dv = _passes(net, theta, v, x, y, cw)
    dw = _passes(net, theta, w, x, y, cw)
    assert torch.isfinite(grad["tbar"]).all() and torch.isfinite(dv["tbar"]).all()

    # P1: A_thetadot of the dual pass is the gradient                                  (fp32: 2e-5; tf32x3: 2e-4)
    assert rel_l2(dv["tdbar"].cpu().numpy(), grad["tbar"].cpu().numpy()) < 2e-5 * tol
    # P2: sum_r cw acbar == <grad, v> per sample                                       (relative to |grad| |v|: 1e-5 / 1e-4)
    fwd = (cw.double() * dv["acbar"].double()).sum(1)
    rev = (grad["tbar"].double() * v.double()).sum(1)
    bound = grad["tbar"].double().norm(dim=1) * v.double().norm(dim=1)
    assert float(((fwd - rev).abs() / bound).max()) < 1e-5 * tol
    # P3: row blocks add up (nll concatenates, weight adjoints add)                    (2e-5 / 2e-4)
    h = R // 2 + 1
    g1 = _passes(net, theta, None, x[:h].contiguous(), y[:h].contiguous(), cw[:, :h].contiguous())
    g2 = _passes(net, theta, None, x[h:].contiguous(), y[h:].contiguous(), cw[:, h:].contiguous())
    assert rel_l2((g1["tbar"].double() + g2["tbar"].double()).cpu().numpy(), grad["tbar"].double().cpu().numpy()) < 2e-5 * tol
    assert rel_l2(torch.cat([g1["nll"], g2["nll"]], 1).cpu().numpy(), grad["nll"].cpu().numpy()) < 1e-5 * tol
    # P4: H (2 v - 3 w) == 2 H v - 3 H w, also for the input adjoint                   (1e-4 / 1e-3)
    dc = _passes(net, theta, (2.0 * v - 3.0 * w).contiguous(), x, y, cw)
    for key in ("tbar", "xbar", "acbar"):
        comb = 2.0 * dv[key].double() - 3.0 * dw[key].double()
        assert rel_l2(dc[key].double().cpu().numpy(), comb.cpu().numpy()) < 1e-4 * tol, key
    # P5: <w, H v> == <v, H w> per sample                                              (relative to |w| |H v|: 1e-4 / 1e-3)
    a = (w.double() * dv["tbar"].double()).sum(1)
    b = (v.double() * dw["tbar"].double()).sum(1)
    bound = w.double().norm(dim=1) * dv["tbar"].double().norm(dim=1)
    assert float(((a - b).abs() / bound).max()) < 1e-4 * tol


def test_cfg2_bilevel_step_is_deterministic_and_engines_agree_at_full_size():
    """cfg2 at full size (fn H=100, M=50, S=10, B=128, T=100): (i) the cluster engine is bit-reproducible run to run
    (fixed-order reductions, no atomics); (ii) the streaming engine -- an independent implementation of the same step
    (different kernels, host-sequenced fused Adam) -- agrees on the loss and the final parameters, and on the hypergradients
    up to what 100 unrolled Adam steps make of fp32 summation-order differences (measured tolerance written below; at
    T = 10 both engines meet the fp64 oracle at 1e-3, tests/test_gpu_engine.py and test_gpu_fn2_stream.py)."""
    from psvi import _native as nat
    from psvi.inference.stream import MeanFieldFamily, StreamEngine
    from psvi.models.neural_net import make_fcnet
    from tests.gpu_util import zeros
    nat.require_cuda()
    torch.manual_seed(0)
    D, H, C, S, M, B, T, N, lr = 2, 100, 2, 10, 50, 128, 100, 800.0, 1e-3
    net = make_fcnet(D, H, C, n_layers=1, mc_samples=S, init_sd=1e-3).cuda()
    net.flat()
    eng = StreamEngine(MeanFieldFamily(net), net.dims, S)
    g = torch.Generator(device="cuda").manual_seed(3)
    u, xb = torch.randn(M, D, device="cuda", generator=g).contiguous(), torch.randn(B, D, device="cuda", generator=g).contiguous()
    z = (u[:, 0] * u[:, 1] > 0).to(torch.int32).contiguous()
    yb = (xb[:, 0] * xb[:, 1] > 0).to(torch.int32).contiguous()
    v = torch.zeros(M, device="cuda")
    eps = torch.randn(T + 1, S, eng.Pt, device="cuda", generator=g).contiguous()
    phi = eng.fam.get_phi().contiguous()
    P = eng.Pt
    model = nat.make_model(net.dims, S)

    def fused():
        mu, rho = phi[:P].clone(), phi[P:].clone()
        traj, gout = zeros(max(nat.traj_floats(model, T), 1)), zeros(nat.gout_floats(model, M))
        ug, vg, loss, il = zeros(M, D), zeros(M), zeros(1), zeros(T)
        nat.nested_step(model, nat.make_noise(eps), mu, rho, u, z, v, xb, yb, B, N, nat.VMODE_SOFTMAX, 0.0, T, lr, 1.0,
                        nat.PHASE_UNROLL | nat.PHASE_REVERSE, traj, gout, ug, vg, None, loss, il)
        torch.cuda.synchronize()
        return mu, rho, ug, vg, loss, il

    r1, r2 = fused(), fused()
    for t1, t2 in zip(r1, r2):
        assert torch.equal(t1, t2)
    mu_f, rho_f, ug_f, vg_f, loss_f, _ = r1
    a = nat.coreset_weights(v, N, nat.VMODE_SOFTMAX, 0.0)
    loss_s, ubar_s, abar_s, phiT_s, _ = eng.nested(phi, eps, u, z, a, xb, yb, N, T, lr)
    torch.cuda.synchronize()
    assert abs(loss_s.item() - loss_f.item()) <= 1e-4 * abs(loss_f.item())
    assert rel_l2(phiT_s.cpu().numpy(), torch.cat([mu_f, rho_f]).cpu().numpy()) < 1e-5
    ug_s, ug = ubar_s.cpu().numpy(), ug_f.cpu().numpy()
    cosine = float((ug_s * ug).sum() / np.linalg.norm(ug_s) / np.linalg.norm(ug))
    assert rel_l2(ug_s, ug) < 2e-2 and cosine > 0.999, (rel_l2(ug_s, ug), cosine)
    # abar is the adjoint of a = N softmax(v); the fused kernel returns the adjoint of v itself
    sm = torch.softmax(v, 0).double()
    vg_s = N * sm * (abar_s.double() - (sm * abar_s.double()).sum())
    assert rel_l2(vg_s.cpu().numpy(), vg_f.double().cpu().numpy()) < 2e-2


def _fulldata_case(D, H, C, S, n_rows, g):
    dims = [D, H, C] if H else [D, C]
    P = po.p_theta(dims)
    mu = (torch.randn(P, device="cuda", generator=g) * 0.05).contiguous()
    rho = torch.full((P,), float(po.inverse_softplus(0.02)), device="cuda")
    x = torch.randn(n_rows, D, device="cuda", generator=g).bfloat16().contiguous()
    y = torch.randint(0, C, (n_rows,), device="cuda", generator=g, dtype=torch.int32)
    return dims, P, mu, rho, x, y


def test_fn_tc_full_data_pass_rows_split_and_permute_at_full_size():
    """psvi_fn_nll_tc at cfg5 shapes (D=256, H=1024, C=10, S=64) over 262 144 rows (Philox noise: the same weights in every
    call): per-sample NLL sums over all rows == the sums over two row blocks == the sums over a row permutation (rtol 2e-5:
    fp32 partial sums in a different order); the per-row values are bit-identical run to run and for rows that stay in the same
    work item, and agree to the last ulp or two otherwise (the order in which a work item walks the hidden-unit chunks of
    GEMM2 rotates with the item index, so the fp32 accumulation order of the logits differs: rtol 1e-6, atol 4e-6)."""
    from psvi import _native as nat
    from tests.gpu_util import zeros
    nat.require_cuda()
    g = torch.Generator(device="cuda").manual_seed(11)
    D, H, C, S, n = 256, 1024, 10, 64, 262144
    dims, P, mu, rho, x, y = _fulldata_case(D, H, C, S, n, g)
    model = nat.make_model(dims, S)
    noise = nat.make_noise(None, seed=1234, domain=7)   # Philox: the same sampled weights in every call
    scratch = zeros(nat.fn_tc_scratch_floats(model, n, 0))
    ones = torch.ones(n, device="cuda")

    def run(xx, yy):
        ws, nk, nl = zeros(S), zeros(S), zeros(S, xx.shape[0])
        nat.fn_nll_tc(model, noise, mu, rho, xx, yy, ones[:xx.shape[0]].contiguous(), 0, ws, nk, nl, scratch)
        torch.cuda.synchronize()
        return ws, nk, nl
    ws, nk, nl = run(x, y)
    assert torch.isfinite(ws).all() and float(nl.min()) >= 0.0
    h = 100003   # not a multiple of the 128-row tile
    ws1, nk1, nl1 = run(x[:h].contiguous(), y[:h].contiguous())
    ws2, _, nl2 = run(x[h:].contiguous(), y[h:].contiguous())
    torch.testing.assert_close(ws1.double() + ws2.double(), ws.double(), rtol=2e-5, atol=0)
    assert torch.equal(nk1, nk)
    assert torch.equal(nl1, nl[:, :h])
    torch.testing.assert_close(nl2, nl[:, h:], rtol=1e-6, atol=4e-6)
    assert torch.equal(run(x, y)[2], nl)
    perm = torch.randperm(n, device="cuda", generator=g)
    wsp, _, nlp = run(x[perm].contiguous(), y[perm].contiguous())
    torch.testing.assert_close(wsp.double(), ws.double(), rtol=2e-5, atol=0)
    torch.testing.assert_close(nlp, nl[:, perm], rtol=1e-6, atol=4e-6)


def _families():
    from psvi.inference.stream import FullCovFamily, LenetFamily, MeanFieldFamily
    from psvi.models.neural_net import make_fc2net, make_fcnet, make_lenet

    def mf():
        net = make_fcnet(256, 1024, 10, n_layers=1, mc_samples=64, init_sd=1e-2).cuda()
        net.flat()
        return MeanFieldFamily(net), 64

    def ln():
        net = make_lenet(mc_samples=10, init_sd=1e-2).cuda()
        net.flat()
        return LenetFamily(net), 10

    def fc():
        return FullCovFamily(make_fc2net(2, 40, 2, n_layers=2, mc_samples=32, init_sd=1e-2).cuda()), 32
    return {"cfg5_meanfield": mf, "cfg4_lenet_family": ln, "cfg3_fullcov": fc}


@pytest.mark.parametrize("name", ["cfg5_meanfield", "cfg4_lenet_family", "cfg3_fullcov"])
def test_family_maps_are_adjoint_and_consistent_at_full_size(name):
    """The variational family seen through its fused maps, at full size (P = 273 k x 64 samples; lenet 61.7 k x 10; the packed
    1 640 x 1 640 triangle x 32):
      * `grad` is the adjoint of `tangent`:  sum_s <A_s, tangent_s(phidot)> == <grad(A; kl = nkl = 0), phidot>   (1e-5 of the norms)
      * `tangent` is the derivative of `sample`:  (sample(phi + d phidot) - sample(phi - d phidot)) / 2d == tangent(phidot) (2e-3)
      * `hvp(A_theta, A_thetadot)` is the derivative of `grad` (with the KL term): with thetabar(d) = A_thetadot + d A_theta (A_thetadot
        IS the gradient wrt theta, A_theta its derivative along the direction -- SURVEY A.6), the central difference of
        grad(phi + d phidot; thetabar(d); kl = 1)                                                                          (2e-3)"""
    from psvi import _native
    _native.require_cuda()
    fam, S = _families()[name]()
    g = torch.Generator(device="cuda").manual_seed(17)
    phi = fam.get_phi().contiguous()
    phidot = (torch.randn(phi.shape, device="cuda", generator=g) * 0.1).contiguous()
    eps = fam.fix_eps(torch.randn(S, fam.Pt, device="cuda", generator=g)).contiguous()
    A = torch.randn(S, fam.Pt, device="cuda", generator=g).contiguous()
    Ad = torch.randn(S, fam.Pt, device="cuda", generator=g).contiguous()
    tang = fam.tangent(phi, phidot, eps)
    gr = fam.grad(phi, eps, A, 0.0, 0.0)
    lhs, rhs = (A.double() * tang.double()).sum(), (gr.double() * phidot.double()).sum()
    assert abs(float(lhs - rhs)) < 1e-5 * float(A.double().norm() * tang.double().norm())
    d = 1e-2
    fd = (fam.sample((phi + d * phidot).contiguous(), eps).double() - fam.sample((phi - d * phidot).contiguous(), eps).double()) / (2 * d)
    assert rel_l2(tang.double().cpu().numpy(), fd.cpu().numpy()) < 2e-3
    h = fam.hvp(phi, phidot, eps, A, Ad)
    gp = fam.grad((phi + d * phidot).contiguous(), eps, (Ad + d * A).contiguous(), 1.0, 0.0).double()
    gm = fam.grad((phi - d * phidot).contiguous(), eps, (Ad - d * A).contiguous(), 1.0, 0.0).double()
    assert rel_l2(h.double().cpu().numpy(), ((gp - gm) / (2 * d)).cpu().numpy()) < 2e-3


def test_lr_tc_full_data_pass_rows_split_and_permute_at_full_size():
    """psvi_lr_predictive_tc (logistic regression D=256, C=10, S=10, M=50) over 2 M bf16 rows, Philox noise: correct-count and
    NLL sums over all rows == the sums over two row blocks == the sums over a row permutation (counts exactly, NLL rtol 2e-5)."""
    from psvi import _native as nat
    from tests.gpu_util import zeros
    nat.require_cuda()
    g = torch.Generator(device="cuda").manual_seed(5)
    D, C, S, M, n = 256, 10, 10, 50, 2_000_000
    dims, P, mu, rho, x, y = _fulldata_case(D, 0, C, S, n, g)
    model = nat.make_model(dims, S)
    noise = nat.make_noise(None, seed=99, domain=3)
    u = torch.randn(M, D, device="cuda", generator=g).contiguous()
    z = torch.randint(0, C, (M,), device="cuda", generator=g, dtype=torch.int32)
    v = torch.zeros(M, device="cuda")
    scratch = zeros(nat.lr_predictive_tc_scratch_floats(model))

    def run(xx, yy):
        out = zeros(8)
        nat.lr_predictive_tc(model, noise, mu, rho, u, z, v, xx, yy, 0, 1000.0, nat.VMODE_SOFTMAX, 0.0, 0, out, scratch)
        torch.cuda.synchronize()
        return out.double()
    full = run(x, y)
    h = 777_777
    a, b = run(x[:h].contiguous(), y[:h].contiguous()), run(x[h:].contiguous(), y[h:].contiguous())
    perm = torch.randperm(n, device="cuda", generator=g)
    p = run(x[perm].contiguous(), y[perm].contiguous())
    # out[0] = sum of the row NLLs, out[1] = number of correct rows, out[2] = number of rows (the psvi_mf_evaluate contract)
    assert float(full[2]) == n and float(a[2] + b[2]) == n
    assert float(a[1] + b[1]) == float(full[1]) == float(p[1]) and 0 < float(full[1]) < n
    torch.testing.assert_close(a[0] + b[0], full[0], rtol=2e-5, atol=0)
    torch.testing.assert_close(p[0], full[0], rtol=2e-5, atol=0)
    assert torch.equal(run(x, y), full)
