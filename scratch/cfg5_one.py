import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "blackbox-coresets-vi_b200"))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
from psvi.experiments.experiments_utils import SynthDataset, make_synthetic_rows
from psvi.inference.psvi_classes import PSVILearnV
T, S = 2, 64
D, H, C, M, B = 256, 1024, 10, 1000, 128
X, Y = make_synthetic_rows(3000, D, C, seed=0)
tr, te = SynthDataset(X[:2500], Y[:2500].float()), SynthDataset(X[2500:], Y[2500:].float())
kw = dict(mc_samples=S, num_epochs=0, data_minibatch=B, D=D, N=2500, inner_it=T, trainer="nested", log_every=1000, lr0u=1e-4,
          lr0net=1e-3, lr0v=1e-3, init_args="subsample", init_sd=1e-3, num_pseudo=M, seed=0, architecture="fn", n_hidden=H,
          n_layers=1, logistic_regression=False, train_dataset=tr, test_dataset=te, dnm="synthetic", nc=C,
          compute_weights_entropy=False, register_elbos=False, quiet=True)
obj = PSVILearnV(**kw)
obj.large_precision = sys.argv[1] if len(sys.argv) > 1 else 'tf32x3'
obj.run_psvi(**kw)
xb, yb = obj._next_minibatch()
obj.nested_step(xb, yb)
torch.cuda.synchronize()
