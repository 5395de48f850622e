import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "blackbox-coresets-vi_b200"))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
from psvi.inference.psvi_classes import PSVILearnV
from tests.fake_mnist import FakeMNIST
T = int(sys.argv[1]) if len(sys.argv) > 1 else 2
tr, te = FakeMNIST(600, 0), FakeMNIST(64, 1)
kw = dict(mc_samples=10, num_epochs=0, data_minibatch=128, D=784, N=len(tr), inner_it=T, trainer="nested", log_every=10,
          lr0u=1e-4, lr0net=1e-3, lr0v=1e-3, init_args="subsample", init_sd=1e-3, num_pseudo=200, seed=0,
          architecture="lenet", n_hidden=0, n_layers=1, logistic_regression=False, train_dataset=tr, test_dataset=te,
          dnm="MNIST", nc=10, compute_weights_entropy=True, register_elbos=False, quiet=True)
obj = PSVILearnV(**kw)
obj.run_psvi(**kw)
xb, yb = obj._next_minibatch()
for _ in range(int(sys.argv[2]) if len(sys.argv) > 2 else 2):
    obj.nested_step(xb, yb)
torch.cuda.synchronize()
