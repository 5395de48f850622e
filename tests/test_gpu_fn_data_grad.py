"""Full-data data-term gradient on the tensor path (psvi_fn_data_grad_tc) against an fp64
restatement of what autograd differentiates in the reference: data_nll = sum_r nll[s, r] inside psvi_elbo
(psvi/inference/psvi_classes.py:477,484-486) for externally sampled weights.  The restatement applies the kernel's bf16
operand roundings (weights, rows, hidden activations, output seeds, hidden adjoints), so the comparison is tight; a second
check against the un-rounded fp64 gradient bounds the total bf16 error."""
import numpy as np
import pytest
import torch

from psvi import _native as nat
from tests.gpu_util import dev

pytestmark = pytest.mark.gpu


def bf(a):
    return torch.as_tensor(np.asarray(a, np.float32)).bfloat16().double().numpy()


def reference(theta, X, y, coef, D, H, C, rounded=True):
    r = bf if rounded else (lambda a: np.asarray(a, np.float64))
    S = theta.shape[0]
    HD = H * D
    dsum, tbar = np.zeros(S), np.zeros_like(theta, dtype=np.float64)
    Xb = r(X)
    for s in range(S):
        W1, b1 = r(theta[s, :HD]).reshape(H, D), theta[s, HD:HD + H].astype(np.float64)
        W2, b2 = r(theta[s, HD + H:HD + H + C * H]).reshape(C, H), theta[s, HD + H + C * H:].astype(np.float64)
        a = Xb @ W1.T + b1
        mask = a > 0
        h = r(np.maximum(a, 0))
        o = h @ W2.T + b2
        o = o - o.max(1, keepdims=True)
        p = np.exp(o) / np.exp(o).sum(1, keepdims=True)
        nll = -np.log(p[np.arange(len(y)), y])
        dsum[s] = nll.sum()
        ob = p.copy()
        ob[np.arange(len(y)), y] -= 1.0
        ob = r(coef[s] * ob)
        apre = (ob @ W2) * mask
        ab = r(apre)
        tbar[s, :HD] = (ab.T @ Xb).reshape(-1)
        tbar[s, HD:HD + H] = apre.sum(0)
        tbar[s, HD + H:HD + H + C * H] = (ob.T @ h).reshape(-1)
        tbar[s, HD + H + C * H:] = ob.sum(0)
    return dsum, tbar


@pytest.mark.parametrize("D,H,C,S,R", [(64, 128, 3, 2, 100), (128, 256, 10, 3, 300), (256, 1024, 10, 4, 777),
                                       (256, 1024, 2, 64, 200)])
def test_data_grad_matches_fp64_restatement(D, H, C, S, R):
    rng = np.random.default_rng(D + H + C + S + R)
    P = H * D + H + C * H + C
    theta = np.concatenate([rng.standard_normal((S, H * D)) / np.sqrt(D), 0.1 * rng.standard_normal((S, H)),
                            rng.standard_normal((S, C * H)) / np.sqrt(H), 0.1 * rng.standard_normal((S, C))], 1).astype(np.float32)
    X = rng.standard_normal((R, D)).astype(np.float32)
    y = rng.integers(0, C, R)
    coef = rng.uniform(0.5, 1.5, S).astype(np.float32)
    model = nat.make_model([D, H, C], S)
    th, xb = dev(theta), dev(X).bfloat16().contiguous()
    yy, cf = dev(y, torch.int32), dev(coef)
    dsum, tbar = torch.zeros(S, device="cuda"), torch.full((S, P), 7.0, device="cuda")   # (tbar is overwritten)
    scr = torch.zeros(nat.fn_data_grad_scratch_floats(model, R), device="cuda")
    nat.fn_data_grad_tc(model, th, xb, yy, cf, dsum, tbar, scr)
    torch.cuda.synchronize()
    rd, rt = reference(theta, X, y, coef, D, H, C, rounded=True)
    ud, ut = reference(theta, X, y, coef, D, H, C, rounded=False)
    got_d, got_t = dsum.cpu().double().numpy(), tbar.cpu().double().numpy()
    assert np.allclose(got_d, rd, rtol=2e-3), (got_d, rd)
    HD = H * D
    blocks = {"W1bar": slice(0, HD), "b1bar": slice(HD, HD + H), "W2bar": slice(HD + H, HD + H + C * H),
              "b2bar": slice(HD + H + C * H, P)}
    for name, sl in blocks.items():
        for s in range(S):
            g, r_, u_ = got_t[s, sl], rt[s, sl], ut[s, sl]
            rel = np.linalg.norm(g - r_) / np.linalg.norm(r_)
            # tolerance: fp32 accumulation order + ReLU units within rounding of zero (bf16-rounded restatement)
            assert rel < 6e-3, (name, s, rel)
    # total bf16 error against the un-rounded fp64 gradient, on the whole per-sample gradient (single blocks such as b2bar
    # are sums with heavy cancellation: their relative error is not meaningful on its own)
    for s in range(S):
        relu_ = np.linalg.norm(got_t[s] - ut[s]) / np.linalg.norm(ut[s])
        assert relu_ < 8e-2, (s, relu_)


def test_data_grad_is_deterministic_and_additive_over_row_shards():
    """Bit-reproducible run to run; and shard gradients add up to the full-data gradient (the identity behind the one
    all-reduce of SURVEY.md section 8e)."""
    D, H, C, S, R = 128, 256, 4, 5, 1000
    rng = np.random.default_rng(5)
    P = H * D + H + C * H + C
    theta = (rng.standard_normal((S, P)) * 0.08).astype(np.float32)
    X = rng.standard_normal((R, D)).astype(np.float32)
    y = rng.integers(0, C, R)
    coef = rng.uniform(0.5, 1.5, S).astype(np.float32)
    model = nat.make_model([D, H, C], S)
    th, xb, yy, cf = dev(theta), dev(X).bfloat16().contiguous(), dev(y, torch.int32), dev(coef)

    def run(lo, hi):
        d, t = torch.zeros(S, device="cuda"), torch.zeros(S, P, device="cuda")
        scr = torch.zeros(nat.fn_data_grad_scratch_floats(model, hi - lo), device="cuda")
        nat.fn_data_grad_tc(model, th, xb[lo:hi].contiguous(), yy[lo:hi].contiguous(), cf, d, t, scr)
        torch.cuda.synchronize()
        return d, t
    d0, t0 = run(0, R)
    d1, t1 = run(0, R)
    assert torch.equal(d0, d1) and torch.equal(t0, t1)
    da, ta = run(0, 384)
    db, tb = run(384, R)
    assert torch.allclose(da + db, d0, rtol=1e-5)
    assert float((ta + tb - t0).norm() / t0.norm()) < 1e-5


def _large_obj(B, min_rows, T=2, seed=0):
    from psvi.experiments.experiments_utils import SynthDataset, make_synthetic_rows
    from psvi.inference.psvi_classes import PSVILearnV
    D, H, C, S, M = 128, 384, 3, 3, 24
    X, Y = make_synthetic_rows(B + 200, D, C, seed=seed)
    tr, te = SynthDataset(X[:B], Y[:B].float()), SynthDataset(X[B:], Y[B:].float())
    kw = dict(mc_samples=S, num_epochs=0, data_minibatch=B, D=D, N=B, inner_it=T, trainer="nested", log_every=10, lr0u=1e-4,
              lr0net=1e-3, lr0v=1e-3, init_args="subsample", init_sd=1e-2, num_pseudo=M, seed=0, architecture="fn",
              n_hidden=H, n_layers=1, logistic_regression=False, train_dataset=tr, test_dataset=te, dnm="synthetic", nc=C,
              compute_weights_entropy=True, register_elbos=False, quiet=True)
    obj = PSVILearnV(**kw)
    obj.fulldata_min_rows = min_rows
    obj.run_psvi(**kw)
    return obj, X, Y, (D, H, C, S, M, T)


def test_nested_step_with_fulldata_tensor_path_matches_oracle():
    """PSVILearnV.nested_step whose data term covers the whole (2 048-row) training set: the class routes the data rows
    through psvi_fn_data_grad_tc (bf16 operands); hypergradients against the fp64 oracle of the reference's nested_step
    (psvi_classes.py:541-600).  Stated tolerance: loss 2e-3 relative, hypergradients rel-L2 5e-2 and cosine >= 0.998 (the
    data-term gradient carries bf16 operand rounding; the unrolled inner loop stays in tf32x3)."""
    from oracle import psvi_oracle as po
    from oracle.ref_import import NoiseFeeder
    from psvi.inference.psvi_classes import ExternalNoise
    B = 2048
    obj, X, Y, (D, H, C, S, M, T) = _large_obj(B, min_rows=1024)
    dims = obj.model.dims
    eps = NoiseFeeder.stream(dims, S, 77, T + 1)
    obj.noise_source = ExternalNoise(eps)
    obj.scheduler_optim_net = None
    mu, rho = [t.cpu().numpy().astype(np.float64) for t in obj.model.flat()]
    u0, z = obj.u.detach().cpu().numpy().astype(np.float64), obj.z.cpu().numpy()
    v0 = obj.v.detach().cpu().numpy().astype(np.float64)
    calls = nat.launch_count()
    xb, yb = obj._next_minibatch()          # full batch: the resident training set itself
    assert xb.shape[0] == B
    loss = obj.nested_step(xb, yb)
    e64 = [e.astype(np.float64) for e in eps]
    r = po.nested_step(mu, rho, np.stack(e64[:T]), e64[T], u0, z, v0, X[:B].numpy().astype(np.float64), Y[:B].numpy(), float(B),
                       dims, 1e-3, vmode=1)
    assert nat.launch_count() > calls
    assert abs(loss.item() - r["loss"]) <= 2e-3 * abs(r["loss"]), (loss.item(), r["loss"])
    gu, gv = obj.u.grad.cpu().double().numpy(), obj.v.grad.cpu().double().numpy()
    cos = lambda a, b: float(a.ravel() @ b.ravel() / (np.linalg.norm(a) * np.linalg.norm(b)))  # noqa: E731
    rel = lambda a, b: float(np.linalg.norm(a - b) / np.linalg.norm(b))  # noqa: E731
    assert rel(gu, r["u_grad"]) < 5e-2 and cos(gu, r["u_grad"]) > 0.998, (rel(gu, r["u_grad"]), cos(gu, r["u_grad"]))
    assert rel(gv, r["v_grad"]) < 5e-2 and cos(gv, r["v_grad"]) > 0.998, (rel(gv, r["v_grad"]), cos(gv, r["v_grad"]))


def test_outer_grad_fulldata_shares_add_up():
    """Two ranks' shares of the outer objective (kappa = 1/2, contiguous row shards) add up to the
    unsharded value and gradients -- what the one all-reduce of the sharded step relies on -- on the tensor data path."""
    B = 3000
    obj, X, Y, (D, H, C, S, M, T) = _large_obj(B, min_rows=1024)
    eng = obj._stream(obj.model)
    phi = eng.fam.get_phi()
    eps = obj._noise_tensor(1, eng.Pt, S)[0]
    u, _ = obj._uv()
    z32, a = obj._z32(), obj._a()
    xb, yb = obj._next_minibatch()
    x16 = xb.bfloat16().contiguous()
    full = eng.outer_grad(phi, eps, u, z32, a, None, yb, float(B), xb_bf16=x16)
    cut = 1700
    s0 = eng.outer_grad(phi, eps, u, z32, a, None, yb[:cut].contiguous(), float(B), kappa=0.5, n_total=B, xb_bf16=x16[:cut].contiguous())
    s1 = eng.outer_grad(phi, eps, u, z32, a, None, yb[cut:].contiguous(), float(B), kappa=0.5, n_total=B, xb_bf16=x16[cut:].contiguous())
    for k in range(4):
        tot = s0[k] + s1[k]
        assert float((tot - full[k]).norm() / full[k].norm()) < 2e-4, k
