"""Tensor-core (TMA + tcgen05 + TMEM) full-data predictive kernel for the single-layer model vs the fp64 oracle.

Operands are bf16 (X and the sampled weights), accumulation fp32: the oracle is evaluated on the SAME bf16-rounded
operands, so the remaining difference is accumulation order and __expf: NLL rtol 2e-4, at most one flipped argmax per
2000 rows.  Against the un-rounded fp32 kernel (psvi_mf_evaluate) the bar is the bf16 input rounding of SURVEY.md
section 4: NLL rtol 2e-2."""
import numpy as np
import pytest
import torch

from oracle import psvi_oracle as po
from tests.gpu_util import dev, zeros

pytestmark = pytest.mark.gpu


def bf16(x):
    return torch.as_tensor(np.asarray(x, dtype=np.float32)).bfloat16().double().numpy()


@pytest.mark.parametrize("D,C,S,n_rows,mode", [(64, 3, 4, 300, 0), (128, 10, 8, 1000, 0), (256, 10, 10, 5000, 0),
                                                (256, 2, 16, 777, 1), (192, 16, 5, 129, 0)])
def test_lr_predictive_tc_matches_oracle(D, C, S, n_rows, mode):
    from psvi import _native as nat
    nat.require_cuda()
    rng = np.random.default_rng(D + C + S)
    dims = [D, C]
    P = po.p_theta(dims)
    M, N = 24, 5000.0
    mu = (0.15 * rng.standard_normal(P)).astype(np.float32)
    rho = np.full(P, po.inverse_softplus(0.05), np.float32)
    eps = rng.standard_normal((1, S, P)).astype(np.float32)
    u = rng.standard_normal((M, D)).astype(np.float32)
    z = rng.integers(0, C, M)
    v = (0.3 * rng.standard_normal(M)).astype(np.float32)
    X = rng.standard_normal((n_rows, D)).astype(np.float32)
    y = rng.integers(0, C, n_rows)
    model = nat.make_model(dims, S)
    xb = dev(X).bfloat16().contiguous()
    out = zeros(8)
    scratch = zeros(nat.lr_predictive_tc_scratch_floats(model))
    nat.lr_predictive_tc(model, nat.make_noise(dev(eps)), dev(mu), dev(rho), dev(u), dev(z, torch.int32), dev(v), xb,
                         dev(y, torch.int32), 0, N, 1, 0.0, mode, out, scratch)
    torch.cuda.synchronize()
    o = out.cpu().numpy()
    # oracle on the same rounded operands
    a = po.coreset_weights(v.astype(np.float64), N, 1)
    theta = po.mf_sample(mu.astype(np.float64), rho.astype(np.float64), eps[0].astype(np.float64))
    lg_u, _ = po.mlp_forward(theta, u.astype(np.float64), dims)
    lw = (po.nll_rows(lg_u, z)[0] @ a) + po.mf_sampled_nkl(mu.astype(np.float64), rho.astype(np.float64),
                                                             eps[0].astype(np.float64), theta)
    w = po.softmax(lw, 0) if mode == 0 else np.full(S, 1.0 / S)
    W = bf16(theta[:, :C * D]).reshape(S, C, D)
    b = theta[:, C * D:]
    logits = np.einsum("rd,scd->src", bf16(X), W) + b[:, None, :]
    probs = (po.softmax(logits, -1) * w[:, None, None]).sum(0)
    pn = np.clip(probs / probs.sum(-1, keepdims=True), np.finfo(np.float32).eps, 1 - np.finfo(np.float32).eps)
    nll = -np.log(pn[np.arange(n_rows), y]).sum()
    corr = np.sum(probs.argmax(-1) == y)
    assert o[2] == n_rows
    np.testing.assert_allclose(o[0], nll, rtol=2e-4)
    assert abs(o[1] - corr) <= 1 + n_rows // 2000
    # un-rounded fp32 kernel on the same inputs
    out2 = zeros(8)
    sc2 = zeros(nat.eval_scratch_floats(model, n_rows, n_rows))
    nat.evaluate(model, nat.make_noise(dev(eps)), dev(mu), dev(rho), dev(u), dev(z, torch.int32), dev(v), dev(X),
                 dev(y, torch.int32), n_rows, 0, N, 1, 0.0, mode, out2, sc2)
    torch.cuda.synchronize()
    o2 = out2.cpu().numpy()
    np.testing.assert_allclose(o[0], o2[0], rtol=2e-2)
    if mode == 0:
        np.testing.assert_allclose(o[3:5], o2[3:5], rtol=1e-5)


@pytest.mark.parametrize("kind", ["lr", "fn"])
def test_predictive_tc_slabs_equals_the_per_slab_calls(kind):
    """psvi_predictive_tc_slabs (one native call over all test batches of a rank: PSVI.evaluate, reference psvi_classes.py:1038-1092,
    a fresh noise slab per batch) == the per-slab entry points called in a loop: sums bit-identical, the importance-weight
    diagnostics those of the LAST slab; ragged last batch (n_rows % batch != 0) and a non-zero first slab index."""
    from psvi import _native as nat
    nat.require_cuda()
    rng = np.random.default_rng(3)
    if kind == "lr":
        dims, S, n_rows, batch = [128, 10], 8, 1000, 384
    else:
        dims, S, n_rows, batch = [64, 128, 3], 4, 700, 256
    D, C = dims[0], dims[-1]
    P = po.p_theta(dims)
    M, N, first = 16, 3000.0, 2
    mu = dev((0.15 * rng.standard_normal(P)).astype(np.float32))
    rho = dev(np.full(P, po.inverse_softplus(0.05), np.float32))
    u, z = dev(rng.standard_normal((M, D)).astype(np.float32)), dev(rng.integers(0, C, M), torch.int32)
    v = dev((0.3 * rng.standard_normal(M)).astype(np.float32))
    xb = dev(rng.standard_normal((n_rows, D)).astype(np.float32)).bfloat16().contiguous()
    y = dev(rng.integers(0, C, n_rows), torch.int32)
    model = nat.make_model(dims, S)
    noise = nat.make_noise(None, seed=5, domain=9)
    n_scr = (nat.lr_predictive_tc_scratch_floats(model) if kind == "lr" else nat.fn_tc_scratch_floats(model, batch, M)) + 256
    one = nat.lr_predictive_tc if kind == "lr" else nat.fn_predictive_tc
    for mode in (0, 1):
        acc, last = torch.zeros(3, device="cuda", dtype=torch.float64), None
        for k, r0 in enumerate(range(0, n_rows, batch)):
            out = zeros(8)
            one(model, noise, mu, rho, u, z, v, xb[r0:r0 + batch], y[r0:r0 + batch], first + k, N, 1, 0.0, mode, out, zeros(n_scr))
            acc += out[:3].double()
            last = out[3:5].clone()
        out = zeros(8)
        nat.predictive_tc_slabs(model, noise, mu, rho, u, z, v, xb, y, batch, first, N, 1, 0.0, mode, out, zeros(n_scr))
        torch.cuda.synchronize()
        assert out[2].item() == n_rows
        np.testing.assert_allclose(out[:3].cpu().numpy(), acc.cpu().numpy(), rtol=1e-6)
        assert torch.equal(out[3:5], last)
