"""cfg1 timing: halfmoon, logistic_regression (one layer, P = 6), psvi_learn_v, M = 10, S = 10, T = 100, B = 128."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "blackbox-coresets-vi_b200"))
import torch
from psvi.experiments.experiments_utils import read_dataset
from psvi.inference.psvi_classes import PSVILearnV
x, y, xt, yt, N, D, tr, te, nc = read_dataset("halfmoon", {"test_ratio": 0.2})
kw = dict(mc_samples=10, num_epochs=0, data_minibatch=128, D=D, N=N, inner_it=100, trainer="nested", log_every=1000, lr0u=1e-4,
          lr0net=1e-3, lr0v=1e-3, init_args="subsample", init_sd=1e-3, num_pseudo=10, seed=0, architecture="logistic_regression",
          n_hidden=0, n_layers=1, logistic_regression=True, train_dataset=tr, test_dataset=te, dnm="halfmoon", nc=nc,
          compute_weights_entropy=True, register_elbos=False, quiet=True)
obj = PSVILearnV(**kw)
obj.run_psvi(**kw)
xb, yb = obj._next_minibatch()
for _ in range(5):
    obj.nested_step(xb, yb)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
n = 100
a.record()
for _ in range(n):
    loss = obj.nested_step(xb, yb)
b.record(); torch.cuda.synchronize()
print(f"cfg1 logistic_regression nested_step (M=10, S=10, T=100): {a.elapsed_time(b) / n:.3f} ms per outer step = "
      f"{n / a.elapsed_time(b) * 1e3:.0f} steps/s, loss {loss.item():.3f}")
