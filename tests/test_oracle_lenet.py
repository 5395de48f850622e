"""The lenet oracle (oracle/lenet_oracle.py: per-sample conv / ReLU+pool / fc layers as linear maps + fixed selections, the
mean-field family with the conv-KL and shared-last-layer quirks Q4/Q5) against the UNMODIFIED reference's fp64 run of
VIConv2d / BatchMaxPool2d / make_lenet through PSVILearnV (tests/golden/lenet_m10.npz, made by oracle/make_goldens.py)."""
import os

import numpy as np

from oracle import lenet_oracle as lo
from oracle import psvi_oracle as po
from oracle import psvi_oracle_generic as pg
from oracle.ref_import import LeNetNoiseFeeder

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def rel_l2(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300)


def load():
    g = dict(np.load(os.path.join(GOLDEN, "lenet_m10.npz")))
    S, T = int(g["S"]), int(g["T"])
    eps = [e.astype(np.float64) for e in LeNetNoiseFeeder.stream(S, int(g["noise_seed"]), int(g["n_forwards"]))]
    return g, S, T, eps


def test_lenet_sizes():
    assert lo.P == 61706 and lo.kl_mask().sum() == 48120 + 10164 + 850


def test_lenet_oracle_matches_reference():
    g, S, T, eps = load()
    N = float(g["N"])
    fam = lo.LeNetMeanField()
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
    np.testing.assert_allclose([nll, ness], g["ref64_eval"][[1, 3]], rtol=1e-6)
    assert abs(went - g["ref64_eval"][2]) < 1e-9
