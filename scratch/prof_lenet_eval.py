import sys, os, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "blackbox-coresets-vi_b200"))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
from collections import defaultdict
from torch.profiler import profile, ProfilerActivity
from psvi.inference.psvi_classes import PSVILearnV
from tests.fake_mnist import FakeMNIST
tr, te = FakeMNIST(2000, 0), FakeMNIST(2048, 1)
kw = dict(mc_samples=10, num_epochs=0, data_minibatch=128, D=784, N=len(tr), inner_it=20, trainer="nested", log_every=10,
          lr0u=1e-4, lr0net=1e-3, lr0v=1e-3, init_args="subsample", init_sd=1e-3, num_pseudo=200, seed=0,
          architecture="lenet", n_hidden=0, n_layers=1, logistic_regression=False, train_dataset=tr, test_dataset=te,
          dnm="MNIST", nc=10, compute_weights_entropy=True, register_elbos=False, quiet=True)
obj = PSVILearnV(**kw)
obj.run_psvi(**kw)
for _ in range(2):
    obj.evaluate()
torch.cuda.synchronize()
t0 = time.time()
for _ in range(3):
    obj.evaluate()
t_enq = (time.time() - t0) / 3
torch.cuda.synchronize()
print(f"evaluate(2048 rows, batch 128): host {1e3*t_enq:.1f} ms, wall {1e3*(time.time()-t0)/3:.1f} ms")
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    obj.evaluate(); torch.cuda.synchronize()
d, tot = defaultdict(lambda: [0, 0.0]), 0.0
for e in prof.events():
    if e.device_type == torch.autograd.DeviceType.CUDA:
        d[e.name[:70]][0] += 1; d[e.name[:70]][1] += e.device_time; tot += e.device_time
print(f"device {tot/1e3:.1f} ms over {sum(v[0] for v in d.values())} launches")
for k, v in sorted(d.items(), key=lambda kv: -kv[1][1])[:10]:
    print(f"  {v[1]/tot*100:5.1f}% n={v[0]:5d} avg {v[1]/v[0]:7.1f} us  {k}")
print(prof.key_averages().table(sort_by="self_cpu_time_total", row_limit=12, max_name_column_width=50))
