"""cfg4 timing: lenet, M=200, S=10, B=128, T in argv (default 20 and 100)."""
import sys, os, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "blackbox-coresets-vi_b200"))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
from psvi.inference.psvi_classes import PSVILearnV
from tests.fake_mnist import FakeMNIST
Ts = [int(a) for a in sys.argv[1:]] or [20, 100]
tr, te = FakeMNIST(2000, 0), FakeMNIST(512, 1)
for T in Ts:
    kw = dict(mc_samples=10, num_epochs=0, data_minibatch=128, D=784, N=len(tr), inner_it=T, trainer="nested", log_every=10,
              lr0u=1e-4, lr0net=1e-3, lr0v=1e-3, init_args="subsample", init_sd=1e-3, num_pseudo=200, seed=0,
              architecture="lenet", n_hidden=0, n_layers=1, logistic_regression=False, train_dataset=tr, test_dataset=te,
              dnm="MNIST", nc=10, compute_weights_entropy=True, register_elbos=False, quiet=True)
    obj = PSVILearnV(**kw)
    obj.run_psvi(**kw)
    xb, yb = obj._next_minibatch()
    for _ in range(2):
        obj.nested_step(xb, yb)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 3
    t0 = time.time(); e0.record()
    for _ in range(n):
        loss = obj.nested_step(xb, yb)
    e1.record(); torch.cuda.synchronize()
    print(f"lenet cfg4 T={T}: {e0.elapsed_time(e1)/n:.1f} ms/outer step (wall {1e3*(time.time()-t0)/n:.1f} ms), loss {loss.item():.1f}", flush=True)
    t0 = time.time()
    acc, nll, *_ = obj.evaluate()
    torch.cuda.synchronize()
    print(f"  evaluate(512 rows): {1e3*(time.time()-t0):.1f} ms acc {acc.item():.3f} nll {nll.item():.3f}", flush=True)
