"""The sharded CUDA path on ONE GPU: the two ranks of a world_size-2 run are emulated back to back (the driver's GPU box
has one device for tests): each "rank" runs PHASE_UNROLL on its shard of the minibatch with pseudo_scale 1/2, the gout
buffers are summed (what the NCCL all-reduce does), and PHASE_REVERSE must reproduce the unsharded hypergradients."""
import numpy as np
import pytest
import torch

from tests.gpu_util import dev, load, rel_l2, zeros

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", ["fn_hm_m50_t10", "logreg_hm_m10", "fn_fb_l2_m13"])
def test_two_emulated_ranks_match_single_rank(name):
    from psvi import _native as nat
    from psvi.inference.psvi_classes import shard_bounds
    nat.require_cuda()
    g, dims, S, T, eps = load(name)
    N, vmode = float(g["N"]), int(g["vmode"])
    model = nat.make_model(dims, S)
    P = nat.num_theta(model)
    M, D = g["u0"].shape
    noise_t = dev(np.stack(eps[2:3 + T]))
    xb, yb = dev(g["xb"]), dev(g["yb"], torch.int32)
    B = xb.shape[0]
    u, z, v = dev(g["u0"]), dev(g["z"], torch.int32), dev(g["v0"])

    def fresh():
        return dev(g["mu0"]), dev(g["rho0"]), zeros(max(nat.traj_floats(model, T), 1)), zeros(nat.gout_floats(model, M))

    # single rank, fused
    mu, rho, traj, gout = fresh()
    ug1, vg1, loss1 = zeros(M, D), zeros(M), zeros(1)
    nat.nested_step(model, nat.make_noise(noise_t), mu, rho, u, z, v, xb, yb, B, N, vmode, 0.0, T, float(g["lr0net"]), 1.0,
                    3, traj, None, ug1, vg1, None, loss1, None)
    # two emulated ranks
    world, gouts, trajs, losses = 2, [], [], []
    for r in range(world):
        lo, hi = shard_bounds(B, r, world)
        mu_r, rho_r, traj_r, gout_r = fresh()
        l_r = zeros(1)
        nat.nested_step(model, nat.make_noise(noise_t), mu_r, rho_r, u, z, v, xb[lo:hi].contiguous(),
                        yb[lo:hi].contiguous(), B, N, vmode, 0.0, T, float(g["lr0net"]), 1.0 / world, nat.PHASE_UNROLL,
                        traj_r, gout_r, None, None, None, l_r, None)
        gouts.append(gout_r); trajs.append(traj_r); losses.append(l_r)
        assert torch.equal(mu_r, mu)          # the replicated inner loop is bit-identical on every rank
    n_red = 2 * P + M * D + M + S + 4
    red = gouts[0].clone()
    red[:n_red] = gouts[0][:n_red] + gouts[1][:n_red]          # == dist.all_reduce(sum)
    assert torch.equal(trajs[0], trajs[1])
    ug2, vg2 = zeros(M, D), zeros(M)
    nat.nested_step(model, nat.make_noise(noise_t), mu_r, rho_r, u, z, v, None, None, B, N, vmode, 0.0, T,
                    float(g["lr0net"]), 1.0 / world, nat.PHASE_REVERSE, trajs[0], red, ug2, vg2, None, None, None)
    torch.cuda.synchronize()
    assert abs((losses[0] + losses[1]).item() - loss1.item()) <= 2e-5 * abs(loss1.item())
    assert rel_l2(ug2.cpu().numpy(), ug1.cpu().numpy()) < 2e-4
    assert rel_l2(vg2.cpu().numpy(), vg1.cpu().numpy()) < 2e-4
