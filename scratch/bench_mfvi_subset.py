"""cfg2 baseline arm: run_mfvi_subset, fn with --n_hidden 100 (two hidden layers in the baselines' set_up_model), M=50, S=10."""
import sys, os, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "blackbox-coresets-vi_b200"))
import torch
from psvi.experiments.experiments_utils import read_dataset
from psvi.inference.baselines import run_mfvi_subset
x, y, xt, yt, N, D, tr, te, nc = read_dataset("halfmoon", {"test_ratio": 0.2})
for epochs in (50, 301):
    torch.cuda.synchronize(); t0 = time.time()
    res = run_mfvi_subset(x=x, y=y, xt=xt, yt=yt, mc_samples=10, data_minibatch=128, num_epochs=epochs, log_every=100, D=D,
                          lr0net=1e-3, seed=0, train_dataset=tr, test_dataset=te, num_pseudo=50, init_args="subsample",
                          architecture="fn", n_hidden=100, nc=nc, dnm="halfmoon", init_sd=1e-3, quiet=True)
    torch.cuda.synchronize(); dt = time.time() - t0
    print(f"run_mfvi_subset num_epochs={epochs}: {2*epochs} Adam steps in {dt*1e3:.1f} ms = {2*epochs/dt:.0f} it/s; acc {res['accs'][-1]:.3f}")
