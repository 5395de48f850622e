"""Pins the CPU oracle (oracle/psvi_oracle.py) against outputs of the UNMODIFIED reference
(tests/golden/*.npz, produced by oracle/make_goldens.py with an injected noise stream).

Tolerances: the fp64 oracle is compared with the fp64 reference run (same fp32 noise) -> tight (1e-9 rel);
the fp32 reference is then compared at the float noise floor SURVEY.md section 4 calibrated
(ELBO rtol 1e-4, hypergradients rel-L2 5e-3 at init_sd 1e-6 / 2e-4 at init_sd >= 1e-3)."""
import glob
import os

import numpy as np
import pytest

from oracle import psvi_oracle as po
from oracle.ref_import import NoiseFeeder

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CASES = sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN, "*.npz"))
               if not os.path.basename(p).startswith(("mfvi", "meanfieldvi", "regressor", "regbase", "sparsebbvi", "fn2", "hyper", "lenet", "ablated", "noiw", "grid", "variant", "fixedpoint", "joint", "alternating", "learnz")))


def rel_l2(a, b):
    return np.linalg.norm(np.asarray(a) - np.asarray(b)) / max(np.linalg.norm(b), 1e-300)


def load(name):
    g = dict(np.load(os.path.join(GOLDEN, name + ".npz")))
    dims = [int(d) for d in g["dims"]]
    S, T = int(g["S"]), int(g["T"])
    eps = NoiseFeeder.stream(dims, S, int(g["noise_seed"]), int(g["n_forwards"]))
    return g, dims, S, T, [e.astype(np.float64) for e in eps]


@pytest.mark.parametrize("name", CASES)
def test_inner_and_outer_elbo_and_grads(name):
    g, dims, S, T, eps = load(name)
    N, vmode = float(g["N"]), int(g["vmode"])
    a = po.coreset_weights(g["v0"], N, vmode)
    val, gmu, grho, gu, ga = po.inner_grad(g["mu0"], g["rho0"], eps[0], g["u0"], g["z"], a, dims)
    assert abs(val - g["ref64_inner_val"]) <= 1e-9 * abs(val)
    assert rel_l2(po.mu_rho_to_phi(gmu, grho, dims), g["ref64_inner_gparams"]) < 1e-9
    assert rel_l2(gu, g["ref64_inner_gu"]) < 1e-9
    if vmode:
        assert rel_l2(po.coreset_weights_vjp(g["v0"], N, vmode, ga)[0], g["ref64_inner_gv"]) < 1e-9
    val, gmu, grho, gu, ga, _ = po.psvi_elbo_grad(g["mu0"], g["rho0"], eps[1], g["u0"], g["z"], a, g["xb"], g["yb"], N, dims)
    assert abs(val - g["ref64_outer_val"]) <= 1e-9 * abs(val)
    assert rel_l2(po.mu_rho_to_phi(gmu, grho, dims), g["ref64_outer_gparams"]) < 1e-8
    assert rel_l2(gu, g["ref64_outer_gu"]) < 1e-8
    if vmode:
        assert rel_l2(po.coreset_weights_vjp(g["v0"], N, vmode, ga)[0], g["ref64_outer_gv"]) < 1e-8
    # fp32 reference sits at its own float noise floor around the fp64 answer
    assert abs(g["ref32_outer_val"] - val) <= 1e-4 * abs(val)


@pytest.mark.parametrize("name", CASES)
def test_nested_step_hypergradient_and_evaluate(name):
    g, dims, S, T, eps = load(name)
    N, vmode = float(g["N"]), int(g["vmode"])
    r = po.nested_step(g["mu0"], g["rho0"], np.stack(eps[2:2 + T]), eps[2 + T], g["u0"], g["z"], g["v0"],
                       g["xb"], g["yb"], N, dims, float(g["lr0net"]), vmode=vmode)
    assert abs(r["loss"] - g["ref64_nested_loss"]) <= 1e-9 * abs(r["loss"])
    assert rel_l2(po.mu_rho_to_phi(r["mu_T"], r["rho_T"], dims), g["ref64_nested_params"]) < 1e-10
    assert rel_l2(r["u_grad"], g["ref64_nested_gu"]) < 1e-7
    if vmode:
        assert rel_l2(r["v_grad"], g["ref64_nested_gv"]) < 1e-7
    # logged inner ELBOs (psvi_classes.py:553-559): every log_every(=10)-th inner loss then the outer one
    ref_elbos = g["ref64_nested_elbos"]
    mine = [-r["inner_losses"][t] for t in range(0, T, 10)] + [-r["loss"]]
    np.testing.assert_allclose(mine, ref_elbos, rtol=1e-9)
    # u, v after their first torch.optim.Adam step (psvi_classes.py:584-586)
    u1, _, _ = po.torch_adam_step(g["u0"], r["u_grad"], 0 * g["u0"], 0 * g["u0"], 1, 1e-4)
    np.testing.assert_allclose(u1, g["ref64_nested_u_after"], rtol=0, atol=1e-9)
    # fp32 reference vs fp64 oracle: SURVEY section 4 tolerances
    sd_small = "sd1e-6" in name
    tol = 5e-3 if sd_small else 2e-4
    assert abs(g["ref32_nested_loss"] - r["loss"]) <= (2e-5 if sd_small else 1e-5) * abs(r["loss"])
    assert rel_l2(g["ref32_nested_gu"], r["u_grad"]) < tol
    if vmode:
        assert rel_l2(g["ref32_nested_gv"], r["v_grad"]) < tol
    # evaluate() on the post-step state
    v1 = g["ref64_nested_v_after"]
    a1 = po.coreset_weights(v1, N, vmode)
    nb = -(-g["xt"].shape[0] // int(g["B"]))
    acc, nll, went, ness = po.evaluate(r["mu_T"], r["rho_T"], eps[3 + T:3 + T + nb], g["ref64_nested_u_after"], g["z"], a1,
                                       g["xt"], g["yt"], dims, int(g["B"]))
    ref = g["ref64_eval"]
    assert abs(acc - ref[0]) < 1e-7  # the reference accumulates `corrects` in a float32 tensor
    np.testing.assert_allclose([nll, went, ness], ref[1:4], rtol=1e-7)
    f = po.softmax(v1, 0) if vmode else v1
    np.testing.assert_allclose(f.sum() ** 2 / (f * f).sum() / len(f), ref[4], rtol=1e-9)


def test_mfvi_subset_trace():
    g = dict(np.load(os.path.join(GOLDEN, "mfvi_subset_hm.npz")))
    dims = [int(d) for d in g["dims"]]
    S, N, M = int(g["S"]), float(g["N"]), int(g["M"])
    eps = [e.astype(np.float32) for e in NoiseFeeder.stream(dims, S, int(g["noise_seed"]), int(g["n_forwards"]))]
    dt = np.float32
    mu, rho = g["mu0"].astype(dt), g["rho0"].astype(dt)
    m = np.zeros(2 * len(mu), dt); v = np.zeros(2 * len(mu), dt)
    xs, ys, xt, yt = g["xs"].astype(dt), g["ys"], g["xt"].astype(dt), g["yt"]
    k, elbos, accs, nlls = 0, [], [], []
    for i in range(6):
        val, gmu, grho = po.mfvi_grad(mu, rho, eps[k], xs, ys, dt(N / M), dims); k += 1
        phi, m, v = po.torch_adam_step(np.concatenate([mu, rho]), np.concatenate([gmu, grho]).astype(dt), m, v, i + 1,
                                       dt(g["lr0net"]))
        mu, rho = phi[:len(mu)].astype(dt), phi[len(mu):].astype(dt)
        elbos.append(-val)
        if i % 2 == 0:
            c, nl = po.mfvi_predict(mu, rho, eps[k], xt, yt, dims); k += 1
            accs.append(c / len(yt)); nlls.append(nl / len(yt))
    np.testing.assert_allclose(elbos, g["ref_elbos"], rtol=2e-5)
    np.testing.assert_allclose(accs, g["ref_accs"], atol=1e-6)
    np.testing.assert_allclose(nlls, g["ref_nlls"], rtol=2e-5)


def test_generic_oracle_reduces_to_meanfield_oracle():
    from oracle import psvi_oracle_generic as pg
    g, dims, S, T, eps = load("fn_fb_l2_m13")
    N, vmode = float(g["N"]), int(g["vmode"])
    fam = pg.MeanField(dims)
    phi0 = po.mu_rho_to_phi(g["mu0"], g["rho0"], dims)
    r = pg.nested_step(fam, phi0, eps[2:2 + T], eps[2 + T], g["u0"], g["z"], g["v0"], g["xb"], g["yb"], N,
                       float(g["lr0net"]), vmode=vmode)
    assert abs(r["loss"] - g["ref64_nested_loss"]) <= 1e-9 * abs(r["loss"])
    assert rel_l2(r["u_grad"], g["ref64_nested_gu"]) < 1e-7
    assert rel_l2(r["v_grad"], g["ref64_nested_gv"]) < 1e-7
    assert rel_l2(r["phi_T"], g["ref64_nested_params"]) < 1e-10


def test_fullcov_oracle_matches_reference_fn2():
    """fn2 (MultivariateNormalVIMixin, neural_net.py:408-491): the closed forms (no triangular solve) against the
    reference's autograd / higher run in fp64."""
    from oracle import psvi_oracle_generic as pg
    g = dict(np.load(os.path.join(GOLDEN, "fn2_hm_h6.npz")))
    dims = [int(d) for d in g["dims"]]
    S, T, N = int(g["S"]), int(g["T"]), float(g["N"])
    eps = [e.astype(np.float64) for e in NoiseFeeder.stream(dims, S, int(g["noise_seed"]), int(g["n_forwards"]))]
    fam = pg.FullCov(dims)
    assert fam.Pphi == len(g["phi0"])
    a = po.coreset_weights(g["v0"], N, 1)
    val, gphi, gu, ga = pg.inner_grad(fam, g["phi0"], eps[0], g["u0"], g["z"], a)
    assert abs(val - g["ref64_inner_val"]) <= 1e-9 * abs(val)
    assert rel_l2(gphi, g["ref64_inner_gparams"]) < 1e-8
    assert rel_l2(gu, g["ref64_inner_gu"]) < 1e-8
    val, gphi, gu, ga = pg.outer_grad(fam, g["phi0"], eps[1], g["u0"], g["z"], a, g["xb"], g["yb"], N)
    assert abs(val - g["ref64_outer_val"]) <= 1e-9 * abs(val)
    assert rel_l2(gphi, g["ref64_outer_gparams"]) < 1e-7
    assert rel_l2(gu, g["ref64_outer_gu"]) < 1e-7
    assert rel_l2(po.coreset_weights_vjp(g["v0"], N, 1, ga)[0], g["ref64_outer_gv"]) < 1e-7
    r = pg.nested_step(fam, g["phi0"], eps[2:2 + T], eps[2 + T], g["u0"], g["z"], g["v0"], g["xb"], g["yb"], N,
                       float(g["lr0net"]), vmode=1)
    assert abs(r["loss"] - g["ref64_nested_loss"]) <= 1e-8 * abs(r["loss"])
    assert rel_l2(r["phi_T"], g["ref64_nested_params"]) < 1e-9
    assert rel_l2(r["u_grad"], g["ref64_nested_gu"]) < 1e-6
    assert rel_l2(r["v_grad"], g["ref64_nested_gv"]) < 1e-6
    a1 = po.coreset_weights(g["ref64_nested_v_after"], N, 1)
    B = int(g["B"])
    nb = -(-g["xt"].shape[0] // B)
    acc, nll, went, ness = pg.evaluate(fam, r["phi_T"], eps[3 + T:3 + T + nb], g["ref64_nested_u_after"], g["z"], a1,
                                       g["xt"], g["yt"], B)
    assert abs(acc - g["ref64_eval"][0]) < 1e-7
    np.testing.assert_allclose([nll, went, ness], g["ref64_eval"][1:4], rtol=1e-6)


@pytest.mark.parametrize("name", ["ablated_fn_hm", "noiw_fn_hm"])
def test_ablated_outer_objective_matches_reference(name):
    """PSVI_Ablated / PSVI_No_IW (psvi_classes.py:1388-1472): outer objective mean_s data_nll - mean_s sampled_nkl (no importance
    weights); PSVI_No_IW trains with mc_samples = 1.  fp64 reference run vs the generic oracle with outer="ablated"."""
    from oracle import psvi_oracle_generic as pg
    g = dict(np.load(os.path.join(GOLDEN, name + ".npz")))
    dims = [int(d) for d in g["dims"]]
    S, T, N = int(g["S"]), int(g["T"]), float(g["N"])
    eps = [e.astype(np.float64) for e in NoiseFeeder.stream(dims, S, int(g["noise_seed"]), int(g["n_forwards"]))]
    fam = pg.MeanField(dims)
    phi0 = po.mu_rho_to_phi(g["mu0"], g["rho0"], dims)
    val, gphi = pg.outer_grad_ablated(fam, phi0, eps[0], g["xb"], g["yb"], N)
    assert abs(val - g["ref64_outer_val"]) <= 1e-9 * abs(val)
    assert rel_l2(gphi, g["ref64_outer_gparams"]) < 1e-8
    r = pg.nested_step(fam, phi0, eps[1:1 + T], eps[1 + T], g["u0"], g["z"], g["v0"], g["xb"], g["yb"], N, float(g["lr0net"]),
                       vmode=1, outer="ablated", no_iw_classes=dims[-1] if S == 1 else None)
    assert abs(r["loss"] - g["ref64_nested_loss"]) <= 1e-8 * abs(r["loss"])
    assert rel_l2(r["phi_T"], g["ref64_nested_params"]) < 1e-9
    assert rel_l2(r["u_grad"], g["ref64_nested_gu"]) < 1e-6
    assert rel_l2(r["v_grad"], g["ref64_nested_gv"]) < 1e-6


VARIANTS = {"PSVIAV": 2, "PSVIAFixedU": 2, "PSVIFixedU": 1, "PSVIFreeV": 0, "PSVI_No_Rescaling": 0}


@pytest.mark.parametrize("cls", sorted(VARIANTS))
def test_variant_nested_step_matches_reference(cls):
    """The remaining PSVI variants (reference psvi_classes.py:1363-1385,1475-1883) through the oracle's nested_step: f = identity
    / softmax / exp(alpha) softmax, with and without updates of u.  Goldens: oracle/make_goldens_r2.py."""
    from oracle.ref_import import NoiseFeeder
    g = dict(np.load(os.path.join(GOLDEN, f"variant_{cls}.npz")))
    dims = [int(d) for d in g["dims"]]
    S, T, N, vmode = int(g["S"]), int(g["T"]), float(g["N"]), VARIANTS[cls]
    eps = [e.astype(np.float64) for e in NoiseFeeder.stream(dims, S, int(g["noise_seed"]), int(g["n_forwards"]))]
    r = po.nested_step(g["mu0"], g["rho0"], np.stack(eps[:T]), eps[T], g["u0"], g["z"], g["v0"], g["xb"], g["yb"], N, dims,
                       float(g["lr0net"]), vmode=vmode, alpha=float(g["alpha0"]))
    assert abs(r["loss"] - g["ref64_nested_loss"]) <= 1e-9 * abs(r["loss"])
    assert rel_l2(po.mu_rho_to_phi(r["mu_T"], r["rho_T"], dims), g["ref64_nested_params"]) < 1e-10
    if int(g["ref64_has_gu"]):
        assert rel_l2(r["u_grad"], g["ref64_nested_gu"]) < 1e-7
        u1, _, _ = po.torch_adam_step(g["u0"], r["u_grad"], 0 * g["u0"], 0 * g["u0"], 1, 1e-4)
        np.testing.assert_allclose(u1, g["ref64_nested_u_after"], rtol=0, atol=1e-9)
    else:   # fixed-u variants never touch u
        np.testing.assert_array_equal(g["ref64_nested_u_after"], g["u0"])
    if int(g["learn_v"]):
        assert rel_l2(r["v_grad"], g["ref64_nested_gv"]) < 1e-7
        v1, _, _ = po.torch_adam_step(g["v0"], r["v_grad"], 0 * g["v0"], 0 * g["v0"], 1, 1e-3)
        if cls == "PSVIFreeV":
            v1 = np.maximum(v1, 0.0)     # clamp of the non-parameterised weights (psvi_classes.py:587-591)
        np.testing.assert_allclose(v1, g["ref64_nested_v_after"], rtol=0, atol=1e-9)
    if vmode == 2:
        np.testing.assert_allclose(r["alpha_grad"], g["ref64_nested_galpha"][0], rtol=1e-7)


@pytest.mark.parametrize("trainer", ["joint", "alternating"])
def test_joint_and_alternating_trainers_match_reference(trainer):
    """`--trainer joint` / `alternating` (reference psvi_classes.py:517-539,871-880): two steps through the oracle's closed-form
    psvi_elbo gradients + torch Adam against the fp64 reference.  Goldens: `python oracle/make_goldens_r2.py trainers`."""
    g = dict(np.load(os.path.join(GOLDEN, f"{trainer}_fn_hm.npz")))
    dims = [int(d) for d in g["dims"]]
    S, N = int(g["S"]), float(g["N"])
    eps = [e.astype(np.float64) for e in NoiseFeeder.stream(dims, S, int(g["noise_seed"]), int(g["n_forwards"]))]
    if trainer == "joint":
        r = po.joint_steps(g["mu0"], g["rho0"], eps, g["u0"], g["z"], g["v0"], g["xb"], g["yb"], N, dims, float(g["lr0joint"]))
    else:
        r = po.alternating_steps(g["mu0"], g["rho0"], eps, g["u0"], g["z"], g["v0"], g["xb"], g["yb"], N, dims,
                                 float(g["lr0net"]), float(g["lr0u"]))
    np.testing.assert_allclose(r[0], g["ref64_losses"], rtol=1e-9)
    assert rel_l2(po.mu_rho_to_phi(r[1], r[2], dims), g["ref64_params"]) < 1e-9
    np.testing.assert_allclose(r[3], g["ref64_u_after"], rtol=0, atol=1e-9)
    np.testing.assert_allclose(r[4], g["ref64_v_after"], rtol=0, atol=1e-9)


def test_learn_z_soft_labels_match_reference():
    """learn_z=True (reference psvi_classes.py:455-474,499-504,546-547,594-595,1049-1056): soft pseudo-labels through the KLDiv
    branch -- inner_elbo, psvi_elbo, one nested_step with the hypergradient on z, and evaluate().  Golden: `python
    oracle/make_goldens_r2.py learnz`."""
    g = dict(np.load(os.path.join(GOLDEN, "learnz_fn_fb.npz")))
    dims = [int(d) for d in g["dims"]]
    S, T, N, nc = int(g["S"]), int(g["T"]), float(g["N"]), int(dims[-1])
    eps = [e.astype(np.float64) for e in NoiseFeeder.stream(dims, S, int(g["noise_seed"]), int(g["n_forwards"]))]
    a = po.coreset_weights(g["v0"], N, 1)
    t_in = po.soft_targets(g["z0"])
    t_all = po.soft_targets(np.concatenate([g["z0"], nc * np.eye(nc)[g["yb"].astype(np.int64)]], 0))
    val = po.inner_grad_soft(g["mu0"], g["rho0"], eps[0], g["u0"], t_in, a, dims)[0]
    assert abs(val - g["ref64_inner_elbo"]) <= 1e-9 * abs(val)
    out = po.psvi_elbo_grad_soft(g["mu0"], g["rho0"], eps[1], g["u0"], t_all, a, g["xb"], N, dims)[0]
    assert abs(out - g["ref64_psvi_elbo"]) <= 1e-9 * abs(out)
    r = po.nested_step_learn_z(g["mu0"], g["rho0"], np.stack(eps[2:2 + T]), eps[2 + T], g["u0"], g["z0"], g["v0"], g["xb"],
                               g["yb"], N, dims, float(g["lr0net"]), nc)
    assert abs(r["loss"] - g["ref64_nested_loss"]) <= 1e-9 * abs(r["loss"])
    assert rel_l2(po.mu_rho_to_phi(r["mu_T"], r["rho_T"], dims), g["ref64_params"]) < 1e-10
    assert rel_l2(r["u_grad"], g["ref64_gu"]) < 1e-7
    assert rel_l2(r["v_grad"], g["ref64_gv"]) < 1e-7
    assert rel_l2(r["z_grad"], g["ref64_gz"]) < 1e-7
    z1, _, _ = po.torch_adam_step(g["z0"], r["z_grad"], 0 * g["z0"], 0 * g["z0"], 1, float(g["lr0z"]))
    np.testing.assert_allclose(z1, g["ref64_z_after"], rtol=0, atol=1e-9)
    # evaluate(): fp32 reference run, importance weights = softmax(sampled_nkl)
    n0 = int(g["n_forwards_step"])
    e32 = [e.astype(np.float32) for e in eps[n0:]]
    acc, nll, went, ness = po.evaluate_learn_z(r["mu_T"].astype(np.float32), r["rho_T"].astype(np.float32), e32,
                                               g["xt"].astype(np.float32), g["yt"], dims, int(g["B"]))
    np.testing.assert_allclose([acc, nll, went, ness], g["ref32_eval"][:4], rtol=2e-4)


def test_meanfieldvi_class_trace_and_forgetting_scores():
    """MeanFieldVI (reference psvi/inference/utils.py:221-450): sequential minibatches, torch Adam, mean-of-logits test, and the
    forgetting-event / never-learnt bookkeeping after every epoch.  Golden: `python oracle/make_goldens_r2.py meanfieldvi`."""
    g = dict(np.load(os.path.join(GOLDEN, "meanfieldvi_hm.npz")))
    dims = [int(d) for d in g["dims"]]
    S, B, dt = int(g["S"]), int(g["B"]), np.float32
    eps = [e.astype(dt) for e in NoiseFeeder.stream(dims, S, int(g["noise_seed"]), int(g["n_forwards"]))]
    mu, rho = g["mu0"].astype(dt), g["rho0"].astype(dt)
    m = np.zeros(2 * len(mu), dt); v = np.zeros(2 * len(mu), dt)
    x, y, xt, yt = g["x"], g["y"].astype(np.int64), g["xt"], g["yt"].astype(np.int64)
    n, k, step, total = x.shape[0], 0, 0, 4
    forgetting, last_acc, never = np.zeros(n), np.zeros(n), np.ones(n)
    elbos, accs, nlls = [], [], []
    for i in range(total):
        for r0 in range(0, n, B):
            val, gmu, grho = po.mfvi_grad(mu, rho, eps[k], x[r0:r0 + B], y[r0:r0 + B], dt(n / len(x[r0:r0 + B])), dims); k += 1
            step += 1
            phi, m, v = po.torch_adam_step(np.concatenate([mu, rho]), np.concatenate([gmu, grho]).astype(dt), m, v, step,
                                           dt(g["lr0net"]))
            mu, rho = phi[:len(mu)].astype(dt), phi[len(mu):].astype(dt)
            elbos.append(-val)
        for r0 in range(0, n, B):
            logits, _ = po.mlp_forward(po.mf_sample(mu, rho, eps[k]), x[r0:r0 + B], dims); k += 1
            cur = (logits.mean(0).argmax(-1) == y[r0:r0 + B]).astype(np.float64)
            forgetting[r0:r0 + B] += last_acc[r0:r0 + B] > cur
            last_acc[r0:r0 + B] = cur
            never[r0:r0 + B] = np.minimum(never[r0:r0 + B], 1.0 - cur)
        if i % 2 == 0 or i == total - 1:
            c, nl = po.mfvi_predict(mu, rho, eps[k], xt, yt, dims); k += 1
            accs.append(c / len(yt)); nlls.append(nl / len(yt))
    assert k == int(g["n_forwards"])
    forgetting = np.maximum(total * never, forgetting)
    np.testing.assert_allclose(elbos, g["ref_elbos"], rtol=5e-5)
    np.testing.assert_allclose(accs, g["ref_accs"], atol=1e-6)
    np.testing.assert_allclose(nlls, g["ref_nlls"], rtol=5e-5)
    assert np.mean(forgetting != g["ref_forgetting"]) <= 0.01      # (a row whose two top logits tie within fp32 noise may flip)
    assert np.mean(last_acc != g["ref_last_acc"]) <= 0.01


@pytest.mark.parametrize("cls", ["PSVILearnV_regressor", "PSVIAV_regressor"])
def test_gaussian_regressors_match_reference(cls):
    """The Gaussian-likelihood regressors (reference psvi_classes.py:1940-2335; constructible only with `device_id` /
    `scheduler_optim_net` supplied from outside, see oracle/make_goldens_r2.py: run_regressor): inner_elbo, psvi_elbo, one
    nested_step with hypergradients on u, v, the learnable targets z (and alpha), the optimiser steps, and evaluate()."""
    g = dict(np.load(os.path.join(GOLDEN, f"regressor_{cls}.npz")))
    dims = [int(d) for d in g["dims"]]
    S, T, N, tau, vmode, alpha = int(g["S"]), int(g["T"]), float(g["N"]), float(g["tau"]), int(g["vmode"]), float(g["alpha0"])
    eps = [e.astype(np.float64) for e in NoiseFeeder.stream(dims, S, int(g["noise_seed"]), int(g["n_forwards"]))]
    a = po.coreset_weights(g["v0"], N, vmode, alpha)
    val = po.inner_grad_gauss(g["mu0"], g["rho0"], eps[0], g["u0"], g["z0"], a, dims, tau)[0]
    assert abs(val - g["ref64_inner_elbo"]) <= 1e-9 * abs(val)
    out = po.psvi_elbo_grad_gauss(g["mu0"], g["rho0"], eps[1], g["u0"], g["z0"], a, g["xb"], g["yb"], N, dims, tau)[0]
    assert abs(out - g["ref64_psvi_elbo"]) <= 1e-9 * abs(out)
    r = po.nested_step_regressor(g["mu0"], g["rho0"], np.stack(eps[2:2 + T]), eps[2 + T], g["u0"], g["z0"], g["v0"], g["xb"],
                                 g["yb"], N, dims, float(g["lr0net"]), tau, vmode=vmode, alpha=alpha)
    assert abs(r["loss"] - g["ref64_nested_loss"]) <= 1e-9 * abs(r["loss"])
    assert rel_l2(po.mu_rho_to_phi(r["mu_T"], r["rho_T"], dims), g["ref64_params"]) < 1e-10
    for k, ref in (("u_grad", "ref64_gu"), ("v_grad", "ref64_gv"), ("z_grad", "ref64_gz")):
        assert rel_l2(r[k], g[ref]) < 1e-7, k
    if vmode == 2:
        np.testing.assert_allclose(r["alpha_grad"], g["ref64_galpha"][0], rtol=1e-7)
    z1, _, _ = po.torch_adam_step(g["z0"], r["z_grad"], 0 * g["z0"], 0 * g["z0"], 1, float(g["lr0z"]))
    np.testing.assert_allclose(z1, g["ref64_z_after"], rtol=0, atol=1e-9)
    e32 = [e.astype(np.float32) for e in eps[int(g["n_forwards_step"]):]]
    rmse, ll = po.evaluate_regressor(r["mu_T"].astype(np.float32), r["rho_T"].astype(np.float32), e32, g["xt"], g["yt"], dims,
                                     int(g["B"]), np.float32(tau), np.float32(g["y_mean"]), np.float32(g["y_std"]))
    np.testing.assert_allclose([rmse, ll], g["ref32_eval"], rtol=2e-5)


def test_regression_baselines_match_reference():
    """run_mfvi_regressor (precision picked on the validation set) and run_mfvi_subset_regressor (reference baselines.py:1066-1346,
    `fit` :1283-1346): every step trains on the FIRST minibatch of the unshuffled loader, scaled by len(dataset) / batch (so the
    subset run is not rescaled at all); predictions = mean over samples of the de-normalised outputs.  fp32 reference."""
    import random
    g = dict(np.load(os.path.join(GOLDEN, "regbase_mfvi.npz")))
    dims = [int(d) for d in g["dims"]]
    S, B, dt = int(g["S"]), int(g["B"]), np.float32
    eps = [e.astype(dt) for e in NoiseFeeder.stream(dims, S, int(g["noise_seed"]), int(g["n_forwards"]))]
    ym, ys = dt(g["y_mean"]), dt(g["y_std"])
    k = [0]

    def fit(init, xtr, ytr, n_train, xp, yp, tau, epochs, log_every):
        mu, rho = g["mu0s"][init].astype(dt), g["rho0s"][init].astype(dt)
        m = np.zeros(2 * len(mu), dt); v = np.zeros(2 * len(mu), dt)
        a = np.full(len(xtr), n_train / len(xtr), dt)
        rm, ll, el = [], [], []
        for e in range(epochs):
            val, gmu, grho = po.inner_grad_gauss(mu, rho, eps[k[0]], xtr, ytr, a, dims, dt(tau)); k[0] += 1
            phi, m, v = po.torch_adam_step(np.concatenate([mu, rho]), np.concatenate([gmu, grho]).astype(dt), m, v, e + 1, dt(1e-2))
            mu, rho = phi[:len(mu)].astype(dt), phi[len(mu):].astype(dt)
            el.append(-val)
            if (e % log_every == 0) if log_every > 0 else (e == epochs - 1):
                se = l = 0.0
                for r0 in range(0, len(xp), B):
                    out, _ = po.mlp_forward(po.mf_sample(mu, rho, eps[k[0]]), xp[r0:r0 + B], dims); k[0] += 1
                    d = (out[..., 0] * ys + ym).mean(0) - yp[r0:r0 + B]
                    se += (d * d).sum(); l += (-0.5 * tau * d * d - 0.5 * np.log(2 * np.pi / tau)).sum()
                rm.append(np.sqrt(se / len(xp))); ll.append(l / len(xp))
        return rm, ll, el
    x, y, xv, yv, xt, yt = (g[n] for n in ("x", "y", "xv", "yv", "xt", "yt"))
    epochs = 3 * max(1, int(len(x) / B))
    best, best_ll = None, -np.inf
    for i, tau in enumerate((0.3, 0.9)):
        _, ll, _ = fit(i, x[:B], y[:B], len(x), xv, yv, tau, epochs, -1)
        if ll[-1] > best_ll:
            best, best_ll = tau, ll[-1]
    assert abs(1.0 / np.sqrt(best) - float(g["ref_full_scale"])) < 1e-9
    rm, ll, el = fit(2, x[:B], y[:B], len(x), xt, yt, best, epochs, 2)
    assert k[0] == int(g["n_forwards_full"])
    np.testing.assert_allclose(rm, g["ref_full_rmses"], rtol=2e-5)
    np.testing.assert_allclose(ll, g["ref_full_lls"], rtol=2e-5)
    np.testing.assert_allclose(el, g["ref_full_elbos"], rtol=5e-5)
    random.seed(0)
    idx = random.sample(range(len(x)), 40)
    rm, ll, el = fit(3, x[idx], y[idx], 40, xt, yt, 0.5, epochs, 2)
    assert k[0] == int(g["n_forwards"])
    np.testing.assert_allclose(rm, g["ref_subset_rmses"], rtol=2e-5)
    np.testing.assert_allclose(ll, g["ref_subset_lls"], rtol=2e-5)
    np.testing.assert_allclose(el, g["ref_subset_elbos"], rtol=5e-5)


def test_sparsebbvi_matches_reference():
    """Sparse-BBVI coreset construction (reference psvi/inference/sparsebbvi.py:28-198 + utils.py:85-141) through the numpy
    restatement oracle/sparsebbvi_oracle.py (Bernoulli likelihood; accumulating inner gradients, the S-fold data term of `elbo`,
    the first-index selection): accuracy / NLL trace and coreset sizes of the fp32 reference."""
    from oracle import sparsebbvi_oracle as so
    g = dict(np.load(os.path.join(GOLDEN, "sparsebbvi_hm.npz")))
    dims = [int(d) for d in g["dims"]]
    eps = iter([e.astype(np.float32) for e in NoiseFeeder.stream(dims, int(g["S"]), int(g["noise_seed"]), int(g["n_forwards"]))])
    r = so.run(g["mu0"], g["rho0"], eps, g["x"], g["y"], g["xt"], g["yt"], dims, int(g["num_epochs"]), int(g["inner_it"]),
               int(g["outer_it"]), int(g["data_minibatch"]), int(g["log_every"]), float(g["lr0"]), int(g["seed"]))
    assert next(eps, None) is None                       # every forward of the reference run was consumed
    assert r["csizes"] == [int(c) for c in g["ref_csizes"]]
    np.testing.assert_allclose(r["accs"], g["ref_accs"], atol=1e-6)
    np.testing.assert_allclose(r["nlls"], g["ref_nlls"], rtol=2e-5)
