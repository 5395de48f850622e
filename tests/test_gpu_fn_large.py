"""Large-regime per-sample pass of fn (batched TMA + tcgen05 GEMMs, csrc/psvi_fn_large.cu) against the fp64 oracle.

Operands are bf16 (inputs, sampled weights, hidden activations, output adjoints), accumulation fp32.  Stated tolerance: the
forward (logits, NLL) is compared with the oracle evaluated on the same bf16-rounded operands (rel-L2 2e-3); gradients,
Hessian-vector products and input adjoints carry the bf16 rounding of the intermediate adjoints and are compared with the
un-rounded fp64 oracle at rel-L2 1.5e-2 and cosine >= 0.9999."""
import numpy as np
import pytest
import torch

from oracle import psvi_oracle as po
from tests.gpu_util import dev, rel_l2, zeros

pytestmark = pytest.mark.gpu


def bf16(x):
    return torch.as_tensor(np.asarray(x, dtype=np.float32)).bfloat16().double().numpy()


def cos(a, b):
    a, b = np.asarray(a, np.float64).ravel(), np.asarray(b, np.float64).ravel()
    return float(a @ b / (np.linalg.norm(a) * np.linalg.norm(b)))


def make_case(D, H, C, S, R, seed):
    rng = np.random.default_rng(seed)
    dims = [D, H, C]
    theta = np.concatenate([rng.standard_normal((S, H * D)) / np.sqrt(D), 0.1 * rng.standard_normal((S, H)),
                            rng.standard_normal((S, C * H)) / np.sqrt(H), 0.1 * rng.standard_normal((S, C))], 1)
    thetad = rng.standard_normal(theta.shape) * np.abs(theta).mean() * 0.5
    # operands that are exactly representable in bf16, so that the oracle sees the same inputs
    theta, thetad = bf16(theta), bf16(thetad)
    X = bf16(rng.standard_normal((R, D)))
    y = rng.integers(0, C, R)
    cw = rng.uniform(0.5, 1.5, (S, R))
    return dims, theta, thetad, X, y, cw


def unambiguous_relu(theta, X, dims, margin=1e-3):
    """Nudges first-layer biases so that no pre-activation lies within `margin` of zero: a unit whose pre-activation is ~1e-6
    from zero takes either side of the ReLU depending on the arithmetic (fp64 oracle vs split-precision GEMM), and ONE flipped
    (row, unit) pair moves the rel-L2 of the first-layer gradient by ~sqrt(1 / (H R)) ~ 3e-3 -- a property of ReLU, not of the
    kernel under test."""
    D, H, _ = dims
    for _ in range(8):
        a1 = np.einsum("rd,shd->srh", X.astype(np.float64), theta[:, :H * D].astype(np.float64).reshape(-1, H, D)) \
            + theta[:, None, H * D:H * D + H].astype(np.float64)
        bad = (np.abs(a1) < margin).any(axis=1)
        if not bad.any():
            break
        theta[:, H * D:H * D + H][bad] += np.float32(3.1 * margin)
    return theta


# the last three shapes have even 128 x 128 tile grids and K >= 256 in (some of) their products: those run on CTA pairs
# (cta_group::2, 256 x 256 tiles); (256, 512, 4, 2, 300) mixes pair and single-CTA launches inside one pass
@pytest.mark.parametrize("prec", ["tf32x3", "bf16x3"])
@pytest.mark.parametrize("D,H,C,S,R", [(64, 128, 3, 2, 100), (128, 256, 10, 3, 300), (256, 384, 4, 2, 129), (192, 128, 16, 2, 260),
                                       (256, 256, 10, 3, 200), (256, 512, 4, 2, 300), (512, 256, 16, 2, 500),
                                       (256, 1024, 10, 4, 1000)])   # the last: BASELINE configs[4]'s exact D / H / C / M (S = 4 of 64)
def test_fnl_pass_tf32x3_matches_oracle(D, H, C, S, R, prec):
    """precision 1 (three kind::tf32 MMAs on (hi, lo) operand pairs): fp32-class accuracy -- rel-L2 2e-5 on every output
    against the fp64 oracle with un-rounded operands.  precision 2 (the same split with bf16 pairs, three kind::f16 MMAs):
    16 mantissa bits per operand -- the same checks at 10x the tolerance (2e-4 / 5e-4)."""
    from psvi import _native as nat
    nat.require_cuda()
    PREC, k = (nat.PREC_TF32X3, 1.0) if prec == "tf32x3" else (nat.PREC_BF16X3, 10.0)
    rng = np.random.default_rng(D + H + C + S + R)
    dims = [D, H, C]
    theta = np.concatenate([rng.standard_normal((S, H * D)) / np.sqrt(D), 0.1 * rng.standard_normal((S, H)),
                            rng.standard_normal((S, C * H)) / np.sqrt(H), 0.1 * rng.standard_normal((S, C))], 1).astype(np.float32)
    thetad = (rng.standard_normal(theta.shape) * np.abs(theta).mean() * 0.5).astype(np.float32)
    X = rng.standard_normal((R, D)).astype(np.float32)
    theta = unambiguous_relu(theta, X, dims)
    y = rng.integers(0, C, R)
    cw = rng.uniform(0.5, 1.5, (S, R)).astype(np.float32)
    model = nat.make_model(dims, S)
    P = theta.shape[1]
    th, thd, x_, y_, cw_ = dev(theta), dev(thetad), dev(X), dev(y, torch.int32), dev(cw)
    t64, td64, X64, cw64 = theta.astype(np.float64), thetad.astype(np.float64), X.astype(np.float64), cw.astype(np.float64)
    nll, logits = zeros(S, R), zeros(S, R, C)
    nat.fnl_pass(model, PREC, th, None, x_, y_, None, nll=nll, logits=logits)
    o, cache = po.mlp_forward(t64, X64, dims)
    ref_nll, p = po.nll_rows(o, y)
    assert rel_l2(logits.cpu().numpy(), o) < 2e-5 * k
    np.testing.assert_allclose(nll.cpu().numpy(), ref_nll, rtol=1e-4 * k, atol=2e-5 * k)
    q = p.copy()
    q[:, np.arange(R), y] -= 1.0
    At, Ax = po.mlp_backward(t64, cache, dims, cw64[:, :, None] * q)
    tbar, xbar = zeros(S, P), zeros(S, R, D)
    nat.fnl_pass(model, PREC, th, None, x_, y_, cw_, nll=nll, tbar=tbar, xbar=xbar)
    blocks = ((0, H * D), (H * D, H * D + H), (H * D + H, H * D + H + C * H), (H * D + H + C * H, P))
    for lo, hi in blocks:
        assert rel_l2(tbar.cpu().numpy()[:, lo:hi], At[:, lo:hi]) < 2e-5 * k
    assert rel_l2(xbar.cpu().numpy(), Ax) < 2e-5 * k
    o, od, c2 = po.mlp_dual_forward(t64, td64, X64, dims)
    c = cw64[:, :, None]
    At, Atd, Ax = po.mlp_dual_backward(t64, td64, c2, dims, c * p * (od - (p * od).sum(-1, keepdims=True)), c * q)
    tbar, tdbar, xbar, ac = zeros(S, P), zeros(S, P), zeros(S, R, D), zeros(S, R)
    nat.fnl_pass(model, PREC, th, thd, x_, y_, cw_, tbar=tbar, tdbar=tdbar, xbar=xbar, acbar=ac)
    torch.cuda.synchronize()
    for lo, hi in blocks:
        assert rel_l2(tbar.cpu().numpy()[:, lo:hi], At[:, lo:hi]) < 5e-5 * k
        assert rel_l2(tdbar.cpu().numpy()[:, lo:hi], Atd[:, lo:hi]) < 2e-5 * k
    assert rel_l2(xbar.cpu().numpy(), Ax) < 5e-5 * k
    assert rel_l2(ac.cpu().numpy(), (q * od).sum(-1)) < 2e-5 * k


@pytest.mark.parametrize("D,H,C,S,R", [(64, 128, 3, 2, 100), (128, 256, 10, 3, 300), (256, 384, 4, 2, 129), (192, 128, 16, 2, 260)])
def test_fnl_pass_bf16_matches_oracle(D, H, C, S, R):
    from psvi import _native as nat
    nat.require_cuda()
    dims, theta, thetad, X, y, cw = make_case(D, H, C, S, R, D + H + C + S + R)
    model = nat.make_model(dims, S)
    P = theta.shape[1]
    th, thd, x_, y_, cw_ = dev(theta), dev(thetad), dev(X), dev(y, torch.int32), dev(cw)
    # forward
    nll, logits = zeros(S, R), zeros(S, R, C)
    nat.fnl_pass(model, nat.PREC_BF16, th, None, x_, y_, None, nll=nll, logits=logits)
    # oracle with the hidden layer rounded to bf16 like the kernel does
    W1 = theta[:, :H * D].reshape(S, H, D)
    b1 = theta[:, H * D:H * D + H]
    W2 = theta[:, H * D + H:H * D + H + C * H].reshape(S, C, H)
    b2 = theta[:, H * D + H + C * H:]
    hr = bf16(np.maximum(np.einsum("rd,shd->srh", X, W1) + b1[:, None, :], 0))
    o_r = np.einsum("srh,sch->src", hr, W2) + b2[:, None, :]
    assert rel_l2(logits.cpu().numpy(), o_r) < 2e-3
    np.testing.assert_allclose(nll.cpu().numpy(), po.nll_rows(o_r, y)[0], rtol=5e-3, atol=5e-3)
    # gradient pass vs the un-rounded oracle
    o, cache = po.mlp_forward(theta, X, dims)
    _, p = po.nll_rows(o, y)
    q = p.copy()
    q[:, np.arange(R), y] -= 1.0
    At, Ax = po.mlp_backward(theta, cache, dims, cw[:, :, None] * q)
    tbar, xbar = zeros(S, P), zeros(S, R, D)
    nat.fnl_pass(model, nat.PREC_BF16, th, None, x_, y_, cw_, nll=nll, tbar=tbar, xbar=xbar)
    assert rel_l2(tbar.cpu().numpy(), At) < 1.5e-2 and cos(tbar.cpu().numpy(), At) > 0.9999
    assert rel_l2(xbar.cpu().numpy(), Ax) < 1.5e-2
    # per-block check (the small blocks must not hide behind the large first-layer block)
    for lo, hi in ((H * D, H * D + H), (H * D + H, H * D + H + C * H), (H * D + H + C * H, P)):
        assert rel_l2(tbar.cpu().numpy()[:, lo:hi], At[:, lo:hi]) < 1.5e-2
    # dual pass
    o, od, c2 = po.mlp_dual_forward(theta, thetad, X, dims)
    c = cw[:, :, None]
    At, Atd, Ax = po.mlp_dual_backward(theta, thetad, c2, dims, c * p * (od - (p * od).sum(-1, keepdims=True)), c * q)
    tbar, tdbar, xbar, ac = zeros(S, P), zeros(S, P), zeros(S, R, D), zeros(S, R)
    nat.fnl_pass(model, nat.PREC_BF16, th, thd, x_, y_, cw_, tbar=tbar, tdbar=tdbar, xbar=xbar, acbar=ac)
    torch.cuda.synchronize()
    assert rel_l2(tbar.cpu().numpy(), At) < 1.5e-2 and cos(tbar.cpu().numpy(), At) > 0.9999
    assert rel_l2(tdbar.cpu().numpy(), Atd) < 1.5e-2 and cos(tdbar.cpu().numpy(), Atd) > 0.9999
    assert rel_l2(xbar.cpu().numpy(), Ax) < 1.5e-2
    assert rel_l2(ac.cpu().numpy(), (q * od).sum(-1)) < 1.5e-2
    for lo, hi in ((H * D, H * D + H), (H * D + H, H * D + H + C * H), (H * D + H + C * H, P)):
        assert rel_l2(tbar.cpu().numpy()[:, lo:hi], At[:, lo:hi]) < 2e-2
        assert rel_l2(tdbar.cpu().numpy()[:, lo:hi], Atd[:, lo:hi]) < 2e-2


@pytest.mark.parametrize("prec,tol,min_cos", [("mixed", 5e-3, 0.9999), ("tf32x3", 5e-3, 0.9999), ("bf16x3", 5e-2, 0.999)])
def test_large_fn_nested_step_and_evaluate_through_psvi_class(prec, tol, min_cos):
    """PSVILearnV on a model in the large regime (P = 50 691 per sample): the class picks the batched-GEMM tensor path
    (FnLargeNet) for inner_elbo / psvi_elbo / nested_step and the fused tcgen05 forward for evaluate; checked against the
    fp64 oracle.  Default arithmetic "mixed" (tf32x3 gradient / outer passes, split-bf16 Hessian-vector passes) and all-tf32x3:
    hypergradients rel-L2 5e-3, cosine > 0.9999 -- the SAME tolerance for both; all-split-bf16 (`large_precision = "bf16x3"`):
    rel-L2 5e-2, cosine > 0.999."""
    from oracle.ref_import import NoiseFeeder
    from psvi.experiments.experiments_utils import SynthDataset, make_synthetic_rows
    from psvi.inference.psvi_classes import ExternalNoise, PSVILearnV
    from psvi.inference.stream import FnLargeNet
    D, H, C, S, M, T, B = 128, 384, 3, 3, 24, 3, 32
    X, Y = make_synthetic_rows(600, D, C, seed=0)
    tr, te = SynthDataset(X[:500], Y[:500].float()), SynthDataset(X[500:], Y[500:].float())
    kw = dict(mc_samples=S, num_epochs=0, data_minibatch=B, D=D, N=500, inner_it=T, trainer="nested", log_every=10, lr0u=1e-4,
              lr0net=1e-3, lr0v=1e-3, init_args="subsample", init_sd=1e-2, num_pseudo=M, seed=0, architecture="fn",
              n_hidden=H, n_layers=1, logistic_regression=False, train_dataset=tr, test_dataset=te, dnm="synthetic", nc=C,
              compute_weights_entropy=True, register_elbos=False, quiet=True)
    obj = PSVILearnV(**kw)
    obj.large_precision = prec
    obj.run_psvi(**kw)
    dims = obj.model.dims
    assert dims == [D, H, C] and obj._is_large_fn(obj.model)
    eps = NoiseFeeder.stream(dims, S, 77, T + 1)
    obj.noise_source = ExternalNoise(eps)
    obj.scheduler_optim_net = None
    mu, rho = [t.cpu().numpy().astype(np.float64) for t in obj.model.flat()]
    u0, z = obj.u.detach().cpu().numpy().astype(np.float64), obj.z.cpu().numpy()
    v0 = obj.v.detach().cpu().numpy().astype(np.float64)
    xb, yb = X[:B].cuda(), Y[:B].cuda()
    loss = obj.nested_step(xb, yb)
    assert isinstance(obj._stream(obj.model).net, FnLargeNet)
    e64 = [e.astype(np.float64) for e in eps]
    r = po.nested_step(mu, rho, np.stack(e64[:T]), e64[T], u0, z, v0, X[:B].numpy().astype(np.float64), Y[:B].numpy(), 500.0,
                       dims, 1e-3, vmode=1)
    assert abs(loss.item() - r["loss"]) <= 1e-4 * abs(r["loss"])
    gu, gv = obj.u.grad.cpu().numpy(), obj.v.grad.cpu().numpy()
    assert rel_l2(gu, r["u_grad"]) < tol and cos(gu, r["u_grad"]) > min_cos
    assert rel_l2(gv, r["v_grad"]) < tol and cos(gv, r["v_grad"]) > min_cos
    muT, rhoT = [t.cpu().numpy() for t in obj.model.flat()]
    assert rel_l2(muT, r["mu_T"]) < 1e-5 and rel_l2(rhoT, r["rho_T"]) < 1e-5
    # evaluate: fused tcgen05 forward (in-kernel Philox noise) -- sanity of the metrics on a learnable synthetic problem
    obj.noise_source = None
    acc, nll, went, ness, vent = obj.evaluate()
    assert 0.0 <= acc.item() <= 1.0 and np.isfinite(nll.item()) and nll.item() > 0


@pytest.mark.parametrize("prec", ["bf16", "tf32x3", "bf16x3"])
def test_fnl_pass_writes_inside_its_outputs_only(prec):
    """Guard regions around every output of the gradient and dual passes stay untouched when the row count is not a multiple
    of the 128-row tile (R = 37) and C = 3 pads to 16; NaN-filled outputs are fully overwritten."""
    from psvi import _native as nat
    nat.require_cuda()
    PREC = {"bf16": nat.PREC_BF16, "tf32x3": nat.PREC_TF32X3, "bf16x3": nat.PREC_BF16X3}[prec]
    D, H, C, S, R, G = 64, 128, 3, 2, 37, 1024
    dims, theta, thetad, X, y, cw = make_case(D, H, C, S, R, 9)
    P = theta.shape[1]
    model = nat.make_model(dims, S)
    th, thd, x_, y_, cw_ = dev(theta), dev(thetad), dev(X), dev(y, torch.int32), dev(cw)

    def guarded(*shape):
        n = int(np.prod(shape))
        buf = torch.full((n + 2 * G,), 12345.0, device="cuda")
        view = buf[G:G + n].view(*shape)
        view.fill_(float("nan"))
        return buf, view
    outs = {k: guarded(*shp) for k, shp in dict(nll=(S, R), tbar=(S, P), xbar=(S, R, D), logits=(S, R, C)).items()}
    nat.fnl_pass(model, PREC, th, None, x_, y_, cw_, **{k: v[1] for k, v in outs.items()})
    outs2 = {k: guarded(*shp) for k, shp in dict(tbar=(S, P), tdbar=(S, P), xbar=(S, R, D), acbar=(S, R)).items()}
    nat.fnl_pass(model, PREC, th, thd, x_, y_, cw_, **{k: v[1] for k, v in outs2.items()})
    torch.cuda.synchronize()
    for group in (outs, outs2):
        for k, (buf, view) in group.items():
            assert torch.all(buf[:G] == 12345.0) and torch.all(buf[-G:] == 12345.0), k
            assert torch.isfinite(view).all(), k


@pytest.mark.parametrize("prec", ["tf32x3", "bf16x3"])
def test_fnl_pass_kernel_variants_agree(prec, monkeypatch):
    """The alternative forms of the batched GEMM kernel compute the same thing: CTA pairs (cta_group::2, 256 x 256 tiles) against
    single CTAs (PSVI_FNL_CG2=0) bit for bit -- the K order of the accumulation is the same --, the ReLU mask as bits against the
    mask read from the stored activations (PSVI_FNL_NO_MBITS=1) and the bias adjoint from the epilogue's column sums against the
    row-sum sweep (PSVI_FNL_NO_COLPART=1) to fp32 summation-order accuracy.  Shape with even tile grids and K >= 256, so that the
    default run does use the pair kernel."""
    from psvi import _native as nat
    nat.require_cuda()
    D, H, C, S, R = 256, 256, 10, 3, 250
    PREC = nat.PREC_TF32X3 if prec == "tf32x3" else nat.PREC_BF16X3
    rng = np.random.default_rng(11)
    dims = [D, H, C]
    theta = np.concatenate([rng.standard_normal((S, H * D)) / np.sqrt(D), 0.1 * rng.standard_normal((S, H)),
                            rng.standard_normal((S, C * H)) / np.sqrt(H), 0.1 * rng.standard_normal((S, C))], 1).astype(np.float32)
    thetad = (rng.standard_normal(theta.shape) * np.abs(theta).mean() * 0.5).astype(np.float32)
    X = rng.standard_normal((R, D)).astype(np.float32)
    theta = unambiguous_relu(theta, X, dims)
    y = rng.integers(0, C, R)
    cw = rng.uniform(0.5, 1.5, (S, R)).astype(np.float32)
    model = nat.make_model(dims, S)
    P = theta.shape[1]
    th, thd, x_, y_, cw_ = dev(theta), dev(thetad), dev(X), dev(y, torch.int32), dev(cw)

    def run():
        nll, tb1, xb1 = zeros(S, R), zeros(S, P), zeros(S, R, D)
        nat.fnl_pass(model, PREC, th, None, x_, y_, cw_, nll=nll, tbar=tb1, xbar=xb1)
        tb2, tdb2, xb2, ac = zeros(S, P), zeros(S, P), zeros(S, R, D), zeros(S, R)
        nat.fnl_pass(model, PREC, th, thd, x_, y_, cw_, tbar=tb2, tdbar=tdb2, xbar=xb2, acbar=ac)
        torch.cuda.synchronize()
        return [t.cpu().numpy() for t in (nll, tb1, xb1, tb2, tdb2, xb2, ac)]

    base = run()
    monkeypatch.setenv("PSVI_FNL_CG2", "0")
    single = run()
    for a, b in zip(base, single):
        np.testing.assert_array_equal(a, b)
    monkeypatch.delenv("PSVI_FNL_CG2")
    monkeypatch.setenv("PSVI_FNL_NO_MBITS", "1")
    for a, b in zip(base, run()):
        assert rel_l2(a, b) < 1e-6
    monkeypatch.delenv("PSVI_FNL_NO_MBITS")
    monkeypatch.setenv("PSVI_FNL_NO_COLPART", "1")
    for a, b in zip(base, run()):
        assert rel_l2(a, b) < 1e-6
