"""N>1 path, host side (runs on CPU with the gloo backend, world_size 2): the data term of psvi_elbo is sharded over
ranks, every rank computes its share L_r = sum_s w_s (d_s^r - p_s / R) - mean(lw) / R, and ONE all-reduce(sum) of
[dL/dphi_T | direct du | direct da | d_s | loss] reproduces the unsharded quantities.  The
per-rank compute here is the CPU oracle (test infrastructure); the CUDA path's own sharded-vs-unsharded check is the GPU
test tests/test_gpu_sharded.py."""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import psvi_oracle as po
from oracle.ref_import import NoiseFeeder

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _worker(rank, world, port, name, out):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, os.path.join(root, "blackbox-coresets-vi_b200"))
    from psvi.inference.psvi_classes import _dist_info, shard_bounds
    d, r, w = _dist_info()
    assert (r, w) == (rank, world) and d is not None
    g = dict(np.load(os.path.join(GOLDEN, name + ".npz")))
    dims = [int(x) for x in g["dims"]]
    eps = NoiseFeeder.stream(dims, int(g["S"]), int(g["noise_seed"]), 2)[1].astype(np.float64)
    N, vmode = float(g["N"]), int(g["vmode"])
    a = po.coreset_weights(g["v0"], N, vmode)
    B = g["xb"].shape[0]
    lo, hi = shard_bounds(B, rank, world)
    loss, gmu, grho, gu, ga, t = po.psvi_elbo_grad(g["mu0"], g["rho0"], eps, g["u0"], g["z"], a, g["xb"][lo:hi],
                                                   g["yb"][lo:hi], N, dims, kappa=1.0 / world, n_total_rows=B)
    buf = torch.from_numpy(np.concatenate([gmu, grho, gu.ravel(), ga, t["ds"], [loss]]))
    dist.all_reduce(buf)                       # the one collective of the outer step
    if rank == 0:
        np.save(out, buf.numpy())
    dist.destroy_process_group()


@pytest.mark.parametrize("name", ["fn_hm_m50_t10", "fn_fb_l2_m13"])
def test_sharded_outer_gradient_allreduce_equals_unsharded(name, tmp_path):
    world, port = 2, 29500 + (os.getpid() % 500)
    out = str(tmp_path / "red.npy")
    mp.spawn(_worker, args=(world, port, name, out), nprocs=world, join=True)
    red = np.load(out)
    g = dict(np.load(os.path.join(GOLDEN, name + ".npz")))
    dims = [int(x) for x in g["dims"]]
    eps = NoiseFeeder.stream(dims, int(g["S"]), int(g["noise_seed"]), 2)[1].astype(np.float64)
    N, vmode = float(g["N"]), int(g["vmode"])
    a = po.coreset_weights(g["v0"], N, vmode)
    loss, gmu, grho, gu, ga, t = po.psvi_elbo_grad(g["mu0"], g["rho0"], eps, g["u0"], g["z"], a, g["xb"], g["yb"], N, dims)
    full = np.concatenate([gmu, grho, gu.ravel(), ga, t["ds"], [loss]])
    np.testing.assert_allclose(red, full, rtol=1e-9, atol=1e-9)
    # and the unsharded oracle equals the reference's autograd (pins the kappa=1 path to the golden as well)
    assert abs(loss - g["ref64_outer_val"]) <= 1e-9 * abs(loss)


def test_shard_bounds_cover_everything_once():
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, os.path.join(root, "blackbox-coresets-vi_b200"))
    from psvi.inference.psvi_classes import shard_bounds
    for n in (0, 1, 7, 128, 1000):
        for world in (1, 2, 3, 8):
            spans = [shard_bounds(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            assert max(h - l for l, h in spans) - min(h - l for l, h in spans) <= 1
