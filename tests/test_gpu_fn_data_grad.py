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
