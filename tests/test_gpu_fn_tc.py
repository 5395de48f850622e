"""Tensor-core sampled-GEMM forward for `fn` with one hidden layer (TMA + tcgen05, hidden activations kept in TMEM)
vs the fp64 oracle.

Operands are bf16 (X, the sampled W1 / W2 and the ReLU activations feeding the second GEMM), accumulation fp32: the
oracle is evaluated on the SAME bf16-rounded operands (incl. the rounded hidden layer), so what remains is accumulation
order and ex2.approx: per-row NLL atol 2e-3 + rtol 2e-3 (a hidden unit whose pre-activation sits within one ulp of a bf16
rounding boundary may round the other way), sums rtol 5e-4.  Against the un-rounded fp64 oracle the bar is the bf16 input
rounding: sums rtol 3e-2."""
import numpy as np
import pytest
import torch

from oracle import psvi_oracle as po
from tests.gpu_util import dev, zeros

pytestmark = pytest.mark.gpu


def bf16(x):
    return torch.as_tensor(np.asarray(x, dtype=np.float32)).bfloat16().double().numpy()


def rounded_logits(theta, X, dims):
    D, H, C = dims
    S = theta.shape[0]
    W1 = bf16(theta[:, :H * D]).reshape(S, H, D)
    b1 = theta[:, H * D:H * D + H]
    W2 = bf16(theta[:, H * D + H:H * D + H + C * H]).reshape(S, C, H)
    b2 = theta[:, H * D + H + C * H:]
    a1 = np.einsum("rd,shd->srh", bf16(X), W1).astype(np.float32).astype(np.float64) + b1[:, None, :]
    h = bf16(np.maximum(a1, 0.0))
    return np.einsum("srh,sch->src", h, W2) + b2[:, None, :]


def make_case(D, H, C, S, n_rows, M, seed):
    rng = np.random.default_rng(seed)
    dims = [D, H, C]
    P = po.p_theta(dims)
    mu = np.concatenate([rng.standard_normal(H * D) / np.sqrt(D), 0.1 * rng.standard_normal(H),
                         rng.standard_normal(C * H) / np.sqrt(H), 0.1 * rng.standard_normal(C)]).astype(np.float32)
    rho = np.full(P, po.inverse_softplus(0.02), np.float32)
    eps = rng.standard_normal((1, S, P)).astype(np.float32)
    X = rng.standard_normal((n_rows, D)).astype(np.float32)
    y = rng.integers(0, C, n_rows)
    u = rng.standard_normal((M, D)).astype(np.float32)
    z = rng.integers(0, C, M)
    v = (0.3 * rng.standard_normal(M)).astype(np.float32)
    return dims, P, mu, rho, eps, X, y, u, z, v


@pytest.mark.parametrize("D,H,C,S,n_rows", [(64, 128, 3, 2, 100), (128, 256, 10, 5, 300), (256, 512, 10, 8, 1000),
                                             (256, 1024, 2, 3, 129), (192, 384, 16, 7, 260)])
def test_fn_nll_tc_matches_oracle(D, H, C, S, n_rows):
    from psvi import _native as nat
    nat.require_cuda()
    dims, P, mu, rho, eps, X, y, *_ = make_case(D, H, C, S, n_rows, 4, D + H + C + S)
    model = nat.make_model(dims, S)
    xb = dev(X).bfloat16().contiguous()
    cw = np.random.default_rng(1).uniform(0.5, 1.5, n_rows).astype(np.float32)
    wsum, nkl, nll = zeros(S), zeros(S), zeros(S, n_rows)
    scratch = zeros(nat.fn_tc_scratch_floats(model, n_rows, 0))
    nat.fn_nll_tc(model, nat.make_noise(dev(eps)), dev(mu), dev(rho), xb, dev(y, torch.int32), dev(cw), 0, wsum, nkl, nll,
                  scratch)
    torch.cuda.synchronize()
    mu64, rho64, e64 = mu.astype(np.float64), rho.astype(np.float64), eps[0].astype(np.float64)
    theta = po.mf_sample(mu64, rho64, e64)
    ref = po.nll_rows(rounded_logits(theta, X, dims), y)[0]
    np.testing.assert_allclose(nll.cpu().numpy(), ref, rtol=2e-3, atol=2e-3)
    np.testing.assert_allclose(wsum.cpu().numpy(), ref @ cw.astype(np.float64), rtol=5e-4)
    np.testing.assert_allclose(nkl.cpu().numpy(), po.mf_sampled_nkl(mu64, rho64, e64, theta), rtol=1e-5)
    # un-rounded oracle: only the bf16 input rounding separates the two
    full = po.nll_rows(po.mlp_forward(theta, X.astype(np.float64), dims)[0], y)[0]
    np.testing.assert_allclose(wsum.cpu().numpy(), full @ cw.astype(np.float64), rtol=3e-2)


@pytest.mark.parametrize("D,H,C,S,n_rows,M,mode", [(64, 128, 3, 4, 300, 24, 0), (256, 256, 10, 8, 2000, 40, 0),
                                                    (128, 512, 4, 64, 700, 16, 1), (256, 1024, 10, 16, 2000, 200, 0),
                                                    (64, 128, 3, 4, 19300, 24, 0), (64, 128, 5, 3, 40000, 8, 1)])
def test_fn_predictive_tc_matches_oracle(D, H, C, S, n_rows, M, mode):
    from psvi import _native as nat
    nat.require_cuda()
    dims, P, mu, rho, eps, X, y, u, z, v = make_case(D, H, C, S, n_rows, M, 7 * D + H + C + S)
    N = 5000.0
    model = nat.make_model(dims, S)
    xb = dev(X).bfloat16().contiguous()
    out = zeros(8)
    scratch = zeros(nat.fn_tc_scratch_floats(model, n_rows, M))
    nat.fn_predictive_tc(model, nat.make_noise(dev(eps)), dev(mu), dev(rho), dev(u), dev(z, torch.int32), dev(v), xb,
                         dev(y, torch.int32), 0, N, 1, 0.0, mode, out, scratch)
    torch.cuda.synchronize()
    o = out.cpu().numpy()
    mu64, rho64, e64 = mu.astype(np.float64), rho.astype(np.float64), eps[0].astype(np.float64)
    a = po.coreset_weights(v.astype(np.float64), N, 1)
    theta = po.mf_sample(mu64, rho64, e64)
    lw = po.nll_rows(rounded_logits(theta, u, dims), z)[0] @ a + po.mf_sampled_nkl(mu64, rho64, e64, theta)
    w = po.softmax(lw, 0) if mode == 0 else np.full(S, 1.0 / S)
    probs = (po.softmax(rounded_logits(theta, X, dims), -1) * w[:, None, None]).sum(0)
    fe = np.finfo(np.float32).eps
    pn = np.clip(probs / probs.sum(-1, keepdims=True), fe, 1 - fe)
    nll = -np.log(pn[np.arange(n_rows), y]).sum()
    corr = np.sum(probs.argmax(-1) == y)
    assert o[2] == n_rows
    np.testing.assert_allclose(o[0], nll, rtol=5e-4)
    assert abs(o[1] - corr) <= 1 + n_rows // 1000
    if mode == 0:
        wp = w[w > 0]
        np.testing.assert_allclose(o[3], -np.sum(np.log(wp) * wp), rtol=2e-3, atol=1e-4)
        np.testing.assert_allclose(o[4], w.sum() ** 2 / np.sum(w * w) / S, rtol=2e-3)


def test_fn_predictive_tc_philox_matches_external():
    """PHILOX mode consumes the same normals psvi_philox_normal materialises (noise fixed from the same seeds)."""
    from psvi import _native as nat
    nat.require_cuda()
    D, H, C, S, n_rows, M = 128, 256, 5, 6, 500, 12
    dims, P, mu, rho, _, X, y, u, z, v = make_case(D, H, C, S, n_rows, M, 3)
    model = nat.make_model(dims, S)
    eps = zeros(3, S, P)
    nat.philox_normal(1234, 9, 0, 3, S, P, eps)
    xb = dev(X).bfloat16().contiguous()
    args = (dev(mu), dev(rho), dev(u), dev(z, torch.int32), dev(v), xb, dev(y, torch.int32))
    o1, o2 = zeros(8), zeros(8)
    scratch = zeros(nat.fn_tc_scratch_floats(model, n_rows, M))
    nat.fn_predictive_tc(model, nat.make_noise(None, seed=1234, domain=9), *args, 2, 800.0, 1, 0.0, 0, o1, scratch)
    nat.fn_predictive_tc(model, nat.make_noise(eps), *args, 2, 800.0, 1, 0.0, 0, o2, scratch)
    torch.cuda.synchronize()
    np.testing.assert_allclose(o1.cpu().numpy(), o2.cpu().numpy(), rtol=1e-6)
