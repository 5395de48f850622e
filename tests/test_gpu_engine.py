"""GPU parity tests: the CUDA engine (through the C ABI) against the fp64 oracle on the golden inputs, and against the
reference's own fp32/fp64 outputs stored in tests/golden.

Tolerances (fp32 arithmetic in a different summation order than ATen):
  values (ELBOs)                 rtol 1e-4   (asserted at 5e-5)
  gradients / HVPs (one pass)    rel-L2 2e-4
  nested hypergradients          rel-L2 5e-3 at init_sd 1e-6, 1e-3 otherwise; cosine >= 0.9999
"""
import numpy as np
import pytest
import torch

from oracle import psvi_oracle as po
from tests.gpu_util import CASES, dev, load, rel_l2, zeros

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def nat():
    from psvi import _native
    _native.require_cuda()
    return _native


@pytest.mark.parametrize("name", CASES)
def test_inner_grad(nat, name):
    g, dims, S, T, eps = load(name)
    N, vmode = float(g["N"]), int(g["vmode"])
    model = nat.make_model(dims, S)
    P = nat.num_theta(model)
    assert P == po.p_theta(dims)
    mu, rho, u, z, v = dev(g["mu0"]), dev(g["rho0"]), dev(g["u0"]), dev(g["z"], torch.int32), dev(g["v0"])
    grad, val = zeros(2 * P), zeros(1)
    nat.inner_grad(model, nat.make_noise(dev(eps[0][None])), mu, rho, u, z, v, N, vmode, 0.0, grad, val)
    torch.cuda.synchronize()
    a = po.coreset_weights(g["v0"], N, vmode)
    oval, gmu, grho, _, _ = po.inner_grad(g["mu0"], g["rho0"], eps[0].astype(np.float64), g["u0"], g["z"], a, dims)
    assert abs(val.item() - oval) <= 5e-5 * abs(oval)
    assert rel_l2(grad.cpu().numpy(), np.concatenate([gmu, grho])) < 2e-4
    # and the reference's own autograd result
    ref = po.phi_to_mu_rho(g["ref64_inner_gparams"], dims)
    assert rel_l2(grad.cpu().numpy(), np.concatenate(ref)) < 2e-4


@pytest.mark.parametrize("name", CASES)
def test_outer_grad(nat, name):
    g, dims, S, T, eps = load(name)
    N, vmode = float(g["N"]), int(g["vmode"])
    model = nat.make_model(dims, S)
    P = nat.num_theta(model)
    M, D = g["u0"].shape
    mu, rho, u, z, v = dev(g["mu0"]), dev(g["rho0"]), dev(g["u0"]), dev(g["z"], torch.int32), dev(g["v0"])
    xb, yb = dev(g["xb"]), dev(g["yb"], torch.int32)
    gout, ug, vg, loss = zeros(nat.gout_floats(model, M)), zeros(M, D), zeros(M), zeros(1)
    nat.outer_grad(model, nat.make_noise(dev(eps[1][None])), mu, rho, u, z, v, xb, yb, xb.shape[0], N, vmode, 0.0, 1.0,
                   gout, ug, vg, None, loss)
    torch.cuda.synchronize()
    a = po.coreset_weights(g["v0"], N, vmode)
    oval, gmu, grho, gu, ga, parts = po.psvi_elbo_grad(g["mu0"], g["rho0"], eps[1].astype(np.float64), g["u0"], g["z"],
                                                        a, g["xb"], g["yb"], N, dims)
    assert abs(loss.item() - oval) <= 5e-5 * abs(oval)
    go = gout.cpu().numpy()
    assert rel_l2(go[:2 * P], np.concatenate([gmu, grho])) < 2e-4
    # the direct u/a partials cancel heavily when sigma -> 0 (all samples alike): the bar is the reference's own
    # fp32-vs-fp64 noise floor on the same quantity
    floor = rel_l2(g["ref32_outer_gu"], g["ref64_outer_gu"])
    tol_u = max(2e-4, 2 * floor)
    assert rel_l2(ug.cpu().numpy(), gu) < tol_u
    # direct dLoss/da_m = sum_s gp_s nll[s,m]: terms of size |gp_s| nll cancel, so the floor is absolute
    gp = -parts["w"] - (parts["w"] * ((parts["ds"] - parts["ps"]) - np.sum(parts["w"] * (parts["ds"] - parts["ps"]))) - 1.0 / S)
    atol = 1e-6 * (np.abs(gp) @ parts["nll"][:, :M])
    assert np.all(np.abs(go[2 * P + M * D:2 * P + M * D + M] - ga) <= tol_u * np.abs(ga) + atol)
    np.testing.assert_allclose(go[2 * P + M * D + M:2 * P + M * D + M + S], parts["ds"], rtol=5e-5)
    gv = po.coreset_weights_vjp(g["v0"], N, vmode, ga)[0]
    fmax = po.softmax(g["v0"], 0).max() if vmode else 1.0
    assert np.all(np.abs(vg.cpu().numpy() - gv) <= max(5e-4, tol_u) * np.abs(gv) + 2 * N * fmax * atol.max())
    assert rel_l2(ug.cpu().numpy(), g["ref64_outer_gu"]) < tol_u


@pytest.mark.parametrize("name", CASES)
def test_inner_hvp(nat, name):
    g, dims, S, T, eps = load(name)
    N, vmode = float(g["N"]), int(g["vmode"])
    model = nat.make_model(dims, S)
    P = nat.num_theta(model)
    M, D = g["u0"].shape
    rng = np.random.default_rng(5)
    gd = rng.standard_normal(2 * P)
    mu, rho, u, z, v = dev(g["mu0"]), dev(g["rho0"]), dev(g["u0"]), dev(g["z"], torch.int32), dev(g["v0"])
    hphi, hu, hv = zeros(2 * P), zeros(M, D), zeros(M)
    nat.inner_hvp(model, nat.make_noise(dev(eps[0][None])), mu, rho, u, z, v, N, vmode, 0.0, dev(gd), hphi, hu, hv)
    torch.cuda.synchronize()
    a = po.coreset_weights(g["v0"], N, vmode)
    gd32 = gd.astype(np.float32).astype(np.float64)
    hmu, hrho, ohu, oha = po.inner_hvp(g["mu0"], g["rho0"], eps[0].astype(np.float64), g["u0"], g["z"], a, dims,
                                       gd32[:P], gd32[P:])
    assert rel_l2(hphi.cpu().numpy(), np.concatenate([hmu, hrho])) < 2e-4
    assert rel_l2(hu.cpu().numpy(), ohu) < 2e-4
    assert rel_l2(hv.cpu().numpy(), po.coreset_weights_vjp(g["v0"], N, vmode, oha)[0]) < 5e-4


def run_nested(nat, g, dims, S, T, eps, phase_split=False):
    N, vmode = float(g["N"]), int(g["vmode"])
    model = nat.make_model(dims, S)
    M, D = g["u0"].shape
    mu, rho, u, z, v = dev(g["mu0"]), dev(g["rho0"]), dev(g["u0"]), dev(g["z"], torch.int32), dev(g["v0"])
    xb, yb = dev(g["xb"]), dev(g["yb"], torch.int32)
    noise = nat.make_noise(dev(np.stack(eps[2:3 + T])))
    traj, gout = zeros(max(nat.traj_floats(model, T), 1)), zeros(nat.gout_floats(model, M))
    ug, vg, loss, il = zeros(M, D), zeros(M), zeros(1), zeros(T)
    args = (model, noise, mu, rho, u, z, v, xb, yb, xb.shape[0], N, vmode, 0.0, T, float(g["lr0net"]), 1.0)
    if phase_split:
        nat.nested_step(*args, nat.PHASE_UNROLL, traj, gout, ug, vg, None, loss, il)
        nat.nested_step(*args, nat.PHASE_REVERSE, traj, gout, ug, vg, None, None, None)
    else:
        nat.nested_step(*args, nat.PHASE_UNROLL | nat.PHASE_REVERSE, traj, gout, ug, vg, None, loss, il)
    torch.cuda.synchronize()
    return dict(mu=mu.cpu().numpy(), rho=rho.cpu().numpy(), ug=ug.cpu().numpy(), vg=vg.cpu().numpy(),
                loss=loss.item(), il=il.cpu().numpy(), traj=traj.cpu().numpy())


@pytest.mark.parametrize("split", [False, True])
@pytest.mark.parametrize("name", CASES)
def test_nested_step(nat, name, split):
    g, dims, S, T, eps = load(name)
    N, vmode = float(g["N"]), int(g["vmode"])
    out = run_nested(nat, g, dims, S, T, eps, phase_split=split)
    e64 = [e.astype(np.float64) for e in eps]
    r = po.nested_step(g["mu0"], g["rho0"], np.stack(e64[2:2 + T]), e64[2 + T], g["u0"], g["z"], g["v0"], g["xb"],
                       g["yb"], N, dims, float(g["lr0net"]), vmode=vmode)
    sd_small = "sd1e-6" in name
    tol = 5e-3 if sd_small else 1e-3
    np.testing.assert_allclose(out["il"], r["inner_losses"], rtol=5e-5)
    assert rel_l2(out["mu"], r["mu_T"]) < 1e-5
    assert rel_l2(out["rho"], r["rho_T"]) < 1e-5
    assert abs(out["loss"] - r["loss"]) <= 1e-4 * abs(r["loss"])
    assert rel_l2(out["ug"], r["u_grad"]) < tol, (rel_l2(out["ug"], r["u_grad"]), rel_l2(g["ref32_nested_gu"], r["u_grad"]))
    cos = np.sum(out["ug"] * r["u_grad"]) / np.linalg.norm(out["ug"]) / np.linalg.norm(r["u_grad"])
    assert cos >= 0.9999
    if vmode:
        assert rel_l2(out["vg"], r["v_grad"]) < tol
    # against the reference's own runs (fp64 and fp32) as stored in the golden
    assert rel_l2(out["ug"], g["ref64_nested_gu"]) < tol
    assert abs(out["loss"] - g["ref32_nested_loss"]) <= 2e-4 * abs(out["loss"])
    assert rel_l2(po.mu_rho_to_phi(out["mu"], out["rho"], dims), g["ref32_nested_params"]) < 1e-5


def test_philox_generator_matches_numpy_restatement(nat):
    from oracle.philox import philox_normal_np
    S, P = 5, 37
    out = zeros(3, S, P)
    nat.philox_normal(0x1234567887654321, 7, 2, 3, S, P, out)
    torch.cuda.synchronize()
    ref = philox_normal_np(0x1234567887654321, 7, 2, 3, S, P)
    np.testing.assert_allclose(out.cpu().numpy(), ref, rtol=0, atol=2e-5)
    big = zeros(64, 16, 1000)
    nat.philox_normal(99, 0, 0, 64, 16, 1000, big)
    x = big.cpu().numpy().ravel().astype(np.float64)
    assert abs(x.mean()) < 5e-3 and abs(x.std() - 1) < 5e-3
    assert abs(((x - x.mean()) ** 4).mean() / x.var() ** 2 - 3) < 0.05


@pytest.mark.parametrize("name", ["fn_hm_m50_t10", "logreg_hm_m10"])
def test_nested_step_philox_mode_equals_external_mode_on_dumped_noise(nat, name):
    """PHILOX mode consumes exactly the slabs psvi_philox_normal materialises."""
    g, dims, S, T, _ = load(name)
    N, vmode = float(g["N"]), int(g["vmode"])
    model = nat.make_model(dims, S)
    P = nat.num_theta(model)
    M, D = g["u0"].shape
    slabs = zeros(T + 1, S, P)
    nat.philox_normal(4321, 3, 0, T + 1, S, P, slabs)
    outs = []
    for noise in (nat.make_noise(None, seed=4321, domain=3), nat.make_noise(slabs)):
        mu, rho, u, z, v = dev(g["mu0"]), dev(g["rho0"]), dev(g["u0"]), dev(g["z"], torch.int32), dev(g["v0"])
        xb, yb = dev(g["xb"]), dev(g["yb"], torch.int32)
        traj = zeros(nat.traj_floats(model, T))
        ug, vg, loss = zeros(M, D), zeros(M), zeros(1)
        nat.nested_step(model, noise, mu, rho, u, z, v, xb, yb, xb.shape[0], N, vmode, 0.0, T, 1e-3, 1.0, 3, traj, None,
                        ug, vg, None, loss, None)
        torch.cuda.synchronize()
        outs.append((ug.cpu().numpy(), vg.cpu().numpy(), loss.item(), mu.cpu().numpy()))
    for a, b in zip(outs[0], outs[1]):
        np.testing.assert_allclose(a, b, rtol=1e-5, atol=1e-7)


def test_unroll_torch_adam_matches_mfvi_golden(nat):
    import os
    from oracle.ref_import import NoiseFeeder
    from tests.gpu_util import GOLDEN
    g = dict(np.load(os.path.join(GOLDEN, "mfvi_subset_hm.npz")))
    dims = [int(d) for d in g["dims"]]
    S, N, M = int(g["S"]), float(g["N"]), int(g["M"])
    eps = NoiseFeeder.stream(dims, S, int(g["noise_seed"]), int(g["n_forwards"]))
    model = nat.make_model(dims, S)
    P = nat.num_theta(model)
    mu, rho = dev(g["mu0"]), dev(g["rho0"])
    am, av = zeros(2 * P), zeros(2 * P)
    xs, ys, xt, yt = dev(g["xs"]), dev(g["ys"], torch.int32), dev(g["xt"]), dev(g["yt"], torch.int32)
    roww = torch.full((M,), N / M, device="cuda")
    k, elbos, accs, nlls = 0, [], [], []
    scratch = zeros(nat.eval_scratch_floats(model, xt.shape[0], 256))
    for i in range(6):
        loss = zeros(1)
        nat.unroll(model, nat.make_noise(dev(eps[k][None])), mu, rho, am, av, i, xs, ys, roww, None, N, 0, 0.0, 1,
                   float(g["lr0net"]), nat.ADAM_TORCH, loss)
        k += 1
        elbos.append(-loss.item())
        if i % 2 == 0:
            out = zeros(8)
            nat.evaluate(model, nat.make_noise(dev(eps[k][None])), mu, rho, None, None, None, xt, yt, 256, 0, N, 0, 0.0,
                         2, out, scratch)
            k += 1
            o = out.cpu().numpy()
            accs.append(o[1] / o[2]); nlls.append(o[0] / o[2])
    np.testing.assert_allclose(elbos, g["ref_elbos"], rtol=5e-5)
    np.testing.assert_allclose(accs, g["ref_accs"], atol=1e-6)
    np.testing.assert_allclose(nlls, g["ref_nlls"], rtol=5e-5)


@pytest.mark.parametrize("name", CASES)
def test_evaluate(nat, name):
    g, dims, S, T, eps = load(name)
    N, vmode = float(g["N"]), int(g["vmode"])
    model = nat.make_model(dims, S)
    B = int(g["B"])
    nb = -(-g["xt"].shape[0] // B)
    P = po.p_theta(dims)
    mu_rho = po.phi_to_mu_rho(g["ref64_nested_params"], dims)
    u1, v1 = g["ref64_nested_u_after"], g["ref64_nested_v_after"]
    e = eps[3 + T:3 + T + nb]
    out = zeros(8)
    scratch = zeros(nat.eval_scratch_floats(model, g["xt"].shape[0], B))
    nat.evaluate(model, nat.make_noise(dev(np.stack(e))), dev(mu_rho[0]), dev(mu_rho[1]), dev(u1), dev(g["z"], torch.int32),
                 dev(v1), dev(g["xt"]), dev(g["yt"], torch.int32), B, 0, N, vmode, 0.0, 0, out, scratch)
    torch.cuda.synchronize()
    o = out.cpu().numpy()
    ref = g["ref64_eval"]
    assert o[2] == g["xt"].shape[0]
    assert abs(o[1] / o[2] - ref[0]) <= 1.0 / o[2] + 1e-6       # at most one borderline row flips
    np.testing.assert_allclose(o[0] / o[2], ref[1], rtol=2e-4)
    np.testing.assert_allclose(o[3], ref[2], rtol=2e-3, atol=1e-5)
    np.testing.assert_allclose(o[4], ref[3], rtol=2e-3)


@pytest.mark.parametrize("D,C,H,S,n_rows,batch", [(2, 2, 100, 10, 5000, 1536), (2, 4, 37, 6, 3001, 1024), (4, 2, 128, 3, 777, 8192)])
@pytest.mark.parametrize("mode", [0, 1, 2])
def test_evaluate_register_rows_kernel_equals_generic(nat, D, C, H, S, n_rows, batch, mode, monkeypatch):
    """psvi_mf_evaluate for one hidden layer and D, C in {2, 4} runs its rows pass in the register-form kernel
    (psvi_mf_eval_rows_fn1_kernel: several 1024-row chunks per noise slab, ragged last chunk and slab); with the same Philox
    noise it must give the generic shared-memory kernel's sums (PSVI_EVAL_GENERIC=1): row count exact, correct count within
    one borderline row, NLL sum to fp32 summation accuracy -- in all three modes (IW-corrected, mean of softmax, mean logits)."""
    rng = np.random.default_rng(D + C + H + S + n_rows)
    dims, M = [D, H, C], 13
    P = po.p_theta(dims)
    mu = (0.5 * rng.standard_normal(P)).astype(np.float32)
    rho = (rng.standard_normal(P) - 2.0).astype(np.float32)
    u = rng.standard_normal((M, D)).astype(np.float32)
    z = rng.integers(0, C, M)
    v = rng.uniform(0.5, 1.5, M).astype(np.float32)
    xt = rng.standard_normal((n_rows, D)).astype(np.float32)
    yt = rng.integers(0, C, n_rows)
    model = nat.make_model(dims, S)
    args = (dev(mu), dev(rho), dev(u), dev(z, torch.int32), dev(v), dev(xt), dev(yt, torch.int32), batch, 0, 800.0, 0, 0.0, mode)

    def run():
        out = zeros(8)
        scratch = zeros(nat.eval_scratch_floats(model, n_rows, batch))
        nat.evaluate(model, nat.make_noise(None, seed=5, domain=3), *args, out, scratch)
        torch.cuda.synchronize()
        return out.cpu().numpy()

    fast = run()
    monkeypatch.setenv("PSVI_EVAL_GENERIC", "1")
    ref = run()
    assert fast[2] == ref[2] == n_rows
    assert abs(fast[1] - ref[1]) <= 1.0
    np.testing.assert_allclose(fast[0], ref[0], rtol=2e-5)
    np.testing.assert_allclose(fast[3:5], ref[3:5], rtol=1e-5, atol=1e-7)
