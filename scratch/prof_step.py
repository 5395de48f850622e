"""Host vs device time of one nested_step for cfg3 (fn2) / cfg5 (large fn): python scratch/prof_step.py cfg3|cfg5"""
import sys, os, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "blackbox-coresets-vi_b200"))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
from collections import defaultdict
from torch.profiler import profile, ProfilerActivity
from psvi.experiments.experiments_utils import SynthDataset, make_synthetic_rows
from psvi.inference.psvi_classes import PSVILearnV
which = sys.argv[1] if len(sys.argv) > 1 else "cfg3"
if which == "cfg3":
    X, Y = make_synthetic_rows(100000, 2, 2, seed=0)
    tr, te = SynthDataset(X[:90000], Y[:90000].float()), SynthDataset(X[90000:], Y[90000:].float())
    kw = dict(mc_samples=32, num_epochs=0, data_minibatch=128, D=2, N=90000, inner_it=20, trainer="nested", log_every=1000,
              lr0u=1e-4, lr0net=1e-3, lr0v=1e-3, init_args="subsample", init_sd=1e-2, num_pseudo=100, seed=0, architecture="fn2",
              n_hidden=40, n_layers=1, logistic_regression=False, train_dataset=tr, test_dataset=te, dnm="synthetic", nc=2,
              compute_weights_entropy=False, register_elbos=False, quiet=True)
else:
    X, Y = make_synthetic_rows(20000, 256, 10, seed=0)
    tr, te = SynthDataset(X[:16000], Y[:16000].float()), SynthDataset(X[16000:], Y[16000:].float())
    kw = dict(mc_samples=64, num_epochs=0, data_minibatch=128, D=256, N=16000, inner_it=10, trainer="nested", log_every=1000,
              lr0u=1e-4, lr0net=1e-3, lr0v=1e-3, init_args="subsample", init_sd=1e-2, num_pseudo=1000, seed=0, architecture="fn",
              n_hidden=1024, n_layers=1, logistic_regression=False, train_dataset=tr, test_dataset=te, dnm="synthetic", nc=10,
              compute_weights_entropy=False, register_elbos=False, quiet=True)
obj = PSVILearnV(**kw)
if len(sys.argv) > 2:
    obj.large_precision = sys.argv[2]
obj.run_psvi(**kw)
xb, yb = obj._next_minibatch()
for _ in range(2):
    obj.nested_step(xb, yb)
torch.cuda.synchronize()
t0 = time.time()
for _ in range(3):
    obj.nested_step(xb, yb)
t_enq = (time.time() - t0) / 3
torch.cuda.synchronize()
print(f"{which}: host enqueue {1e3*t_enq:.1f} ms/step, wall {1e3*(time.time()-t0)/3:.1f} ms/step")
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    obj.nested_step(xb, yb)
    torch.cuda.synchronize()
d, tot = defaultdict(lambda: [0, 0.0]), 0.0
for e in prof.events():
    if e.device_type == torch.autograd.DeviceType.CUDA:
        d[e.name[:70]][0] += 1; d[e.name[:70]][1] += e.device_time; tot += e.device_time
print(f"sum of device activity {tot/1e3:.1f} ms over {sum(v[0] for v in d.values())} launches")
for k, v in sorted(d.items(), key=lambda kv: -kv[1][1])[:18]:
    print(f"  {v[1]/tot*100:5.1f}% n={v[0]:5d} avg {v[1]/v[0]:7.1f} us  {k}")
