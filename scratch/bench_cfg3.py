"""cfg3 timing: fn2 (full covariance, two hidden layers of 40 units), synthetic 2-d data N=100k, M=100, S=32, T=20, B=128."""
import sys, os, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "blackbox-coresets-vi_b200"))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
from psvi.experiments.experiments_utils import SynthDataset, make_synthetic_rows
from psvi.inference.psvi_classes import PSVILearnV
H = int(sys.argv[1]) if len(sys.argv) > 1 else 40
X, Y = make_synthetic_rows(100000, 2, 2, seed=0)
tr, te = SynthDataset(X[:90000], Y[:90000].float()), SynthDataset(X[90000:], Y[90000:].float())
kw = dict(mc_samples=32, num_epochs=0, data_minibatch=128, D=2, N=90000, inner_it=20, trainer="nested", log_every=1000, lr0u=1e-4,
          lr0net=1e-3, lr0v=1e-3, init_args="subsample", init_sd=1e-2, num_pseudo=100, seed=0, architecture="fn2", n_hidden=H,
          n_layers=1, logistic_regression=False, train_dataset=tr, test_dataset=te, dnm="synthetic", nc=2,
          compute_weights_entropy=False, register_elbos=False, quiet=True)
obj = PSVILearnV(**kw)
obj.run_psvi(**kw)
print("dims", obj.model.dims, "params", sum(p.numel() for p in obj.model.parameters()))
xb, yb = obj._next_minibatch()
obj.nested_step(xb, yb)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
n = 3
e0.record()
for _ in range(n):
    loss = obj.nested_step(xb, yb)
e1.record(); torch.cuda.synchronize()
print(f"cfg3 fn2 H={H} nested_step: {e0.elapsed_time(e1)/n:.1f} ms/outer step, loss {loss.item():.1f}", flush=True)
obj.data_minibatch = 10000
t0 = time.time(); acc, nll, *_ = obj.evaluate(); torch.cuda.synchronize()
print(f"evaluate 10000 rows: {1e3*(time.time()-t0):.1f} ms acc {acc.item():.3f} nll {nll.item():.3f}")
