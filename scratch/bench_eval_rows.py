"""PSVI.evaluate over 2 M synthetic rows at cfg2 shapes (D=2, fn H=100, S=10, M=50, batch 8192): register-form rows kernel
against the generic one (PSVI_EVAL_GENERIC=1); prints both results (acc, nll) and the median time."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "blackbox-coresets-vi_b200"))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
from psvi.experiments.experiments_utils import SynthDataset, make_synthetic_rows, read_dataset
from psvi.inference.psvi_classes import PSVILearnV
x, y, xt, yt, N, D, tr, te, nc = read_dataset("halfmoon", {"test_ratio": 0.2})
kw = dict(mc_samples=10, num_epochs=0, data_minibatch=128, D=D, N=N, inner_it=100, trainer="nested", log_every=1000, lr0u=1e-4,
          lr0net=1e-3, lr0v=1e-3, init_args="subsample", init_sd=1e-3, num_pseudo=50, seed=0, architecture="fn", n_hidden=100,
          n_layers=1, logistic_regression=False, train_dataset=tr, test_dataset=te, dnm="halfmoon", nc=nc,
          compute_weights_entropy=True, register_elbos=False, quiet=True)
obj = PSVILearnV(**kw)
obj.run_psvi(**kw)
n_big = 2_000_000
xb, yb = make_synthetic_rows(n_big, D, nc, seed=0)
obj.test_dataset = SynthDataset(xb, yb.float())
obj.data_minibatch = 8192
obj._dev_data.pop("test", None)
for mode in ("fn1", "generic"):
    if mode == "generic":
        os.environ["PSVI_EVAL_GENERIC"] = "1"
    res = [float(t) for t in obj.evaluate()]
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    per = []
    for _ in range(10):
        a.record(); obj.evaluate(); b.record(); torch.cuda.synchronize()
        per.append(a.elapsed_time(b))
    ms = sorted(per)[len(per) // 2]
    print(f"{mode}: {ms:.3f} ms per pass, {n_big * 10 / ms / 1e6:.2f} G row-samples/s; acc/nll/went/ness/vent = {res}", flush=True)
