"""Real-NCCL check of the sharded path (not collected by pytest: needs >= 2 GPUs and torchrun):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tests/multi_gpu_check.py

Every rank runs PSVILearnV.nested_step / evaluate with torch.distributed initialised (minibatch rows and test batches
sharded, ONE all-reduce each) and rank 0 compares with the same step computed unsharded on its own GPU."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "blackbox-coresets-vi_b200"))


def build(seed):
    from psvi.experiments.experiments_utils import read_dataset
    from psvi.inference.psvi_classes import PSVILearnV
    x, y, xt, yt, N, D, tr, te, nc = read_dataset("halfmoon", {"test_ratio": 0.2})
    kw = dict(mc_samples=10, num_epochs=0, data_minibatch=64, D=D, N=N, inner_it=20, trainer="nested", log_every=10,
              lr0u=1e-4, lr0net=1e-3, lr0v=1e-3, init_args="subsample", init_sd=1e-3, num_pseudo=20, seed=seed,
              architecture="fn", n_hidden=50, n_layers=1, logistic_regression=False, train_dataset=tr, test_dataset=te,
              dnm="halfmoon", nc=nc, compute_weights_entropy=True, register_elbos=False, quiet=True)
    obj = PSVILearnV(**kw)
    obj.run_psvi(**kw)
    obj.scheduler_optim_net = None
    return obj, x, y


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
    dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ["LOCAL_RANK"])))
    import psvi.inference.psvi_classes as pc
    obj, x, y = build(seed=0)                       # identical on every rank (same seed)
    xb, yb = x[:64].cuda(), y[:64].cuda()
    loss = obj.nested_step(xb, yb)                  # sharded: rows split over ranks + one all-reduce
    ev = [t.item() for t in obj.evaluate()[:4]]
    ug, vg = obj.u.grad.clone(), obj.v.grad.clone()
    # every rank must hold identical results (replicated reverse sweep)
    g = [torch.zeros_like(ug) for _ in range(world)]
    dist.all_gather(g, ug)
    same = all(torch.equal(g[0], t) for t in g)
    if rank == 0:
        real = pc._dist_info
        pc._dist_info = lambda: (None, 0, 1)       # unsharded reference on this GPU
        ref, _, _ = build(seed=0)
        loss_ref = ref.nested_step(xb, yb)
        ev_ref = [t.item() for t in ref.evaluate()[:4]]
        pc._dist_info = real
        rel = lambda a, b: float((a - b).norm() / b.norm())  # noqa: E731
        print(f"world={world} identical_across_ranks={same} loss {loss.item():.5f} vs {loss_ref.item():.5f} "
              f"u_grad rel {rel(ug, ref.u.grad):.2e} v_grad rel {rel(vg, ref.v.grad):.2e} eval {ev} vs {ev_ref}")
        ok = (same and abs(loss.item() - loss_ref.item()) <= 2e-5 * abs(loss_ref.item()) and rel(ug, ref.u.grad) < 2e-4
              and rel(vg, ref.v.grad) < 2e-4 and np.allclose(ev, ev_ref, rtol=2e-4, atol=1e-5))
        print("MULTI_GPU_CHECK", "PASS" if ok else "FAIL")
    # ---- the streaming path (fn2 / lenet / medium and large models): data term sharded, one all-reduce
    obj2, x, y = build(seed=0)
    obj2._ws[("force_stream", id(obj2.model))] = True
    loss2 = obj2.nested_step(xb, yb)
    ug2, vg2 = obj2.u.grad.clone(), obj2.v.grad.clone()
    g = [torch.zeros_like(ug2) for _ in range(world)]
    dist.all_gather(g, ug2)
    same2 = all(torch.equal(g[0], t) for t in g)
    if rank == 0:
        real = pc._dist_info
        pc._dist_info = lambda: (None, 0, 1)
        ref2, _, _ = build(seed=0)
        ref2._ws[("force_stream", id(ref2.model))] = True
        loss2_ref = ref2.nested_step(xb, yb)
        pc._dist_info = real
        rel = lambda a, b: float((a - b).norm() / b.norm())  # noqa: E731
        print(f"stream: world={world} identical_across_ranks={same2} loss {loss2.item():.5f} vs {loss2_ref.item():.5f} "
              f"u_grad rel {rel(ug2, ref2.u.grad):.2e} v_grad rel {rel(vg2, ref2.v.grad):.2e}")
        ok2 = (same2 and abs(loss2.item() - loss2_ref.item()) <= 2e-5 * abs(loss2_ref.item())
               and rel(ug2, ref2.u.grad) < 5e-4 and rel(vg2, ref2.v.grad) < 5e-4)
        print("MULTI_GPU_CHECK_STREAM", "PASS" if ok2 else "FAIL")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
