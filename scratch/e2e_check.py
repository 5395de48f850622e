"""End-to-end sanity on the GPU: (a) the flow_psvi CLI on cfg1, (b) a short lenet PSVI run learns the synthetic digits."""
import sys, os, subprocess, time, glob, pickle
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, os.path.join(ROOT, "blackbox-coresets-vi_b200")); sys.path.insert(0, ROOT)
env = dict(os.environ, PYTHONPATH=os.path.join(ROOT, "blackbox-coresets-vi_b200"))
t0 = time.time()
r = subprocess.run([sys.executable, "-m", "psvi.experiments.flow_psvi", "--datasets", "halfmoon", "--architecture", "logistic_regression",
                    "--methods", "psvi_learn_v", "mfvi_subset", "--coreset_sizes", "10", "--num_trials", "1", "--num_epochs", "61",
                    "--results_folder", "/tmp/psvi_res", "--data_folder", "/tmp/psvi_data", "--log_every", "20"], env=env,
                   capture_output=True, text=True)
print("flow_psvi rc", r.returncode, f"{time.time()-t0:.1f}s")
print(r.stdout[-600:]); print(r.stderr[-400:] if r.returncode else "")
for f in glob.glob("/tmp/psvi_res/*.pk"):
    res = pickle.load(open(f, "rb"))
    for d, m in res.items():
        for meth, sz in m.items():
            for s_, tr in sz.items():
                for t_, rr in tr.items():
                    print(d, meth, s_, t_, {k: (v[-1] if hasattr(v, "__len__") and len(v) else v) for k, v in rr.items() if k in ("accs", "nlls")})
import torch
from psvi.inference.psvi_classes import PSVILearnV
from tests.fake_mnist import FakeMNIST
tr, te = FakeMNIST(2000, 0), FakeMNIST(500, 1)
kw = dict(mc_samples=8, num_epochs=31, data_minibatch=128, D=784, N=len(tr), inner_it=10, trainer="nested", log_every=10,
          lr0u=1e-3, lr0net=1e-3, lr0v=1e-2, init_args="subsample", init_sd=1e-3, num_pseudo=50, seed=0, architecture="lenet",
          n_hidden=0, n_layers=1, logistic_regression=False, train_dataset=tr, test_dataset=te, dnm="MNIST", nc=10,
          compute_weights_entropy=True, register_elbos=False, quiet=True)
t0 = time.time()
res = PSVILearnV(**kw).run_psvi(**kw)
print("lenet psvi_learn_v: accs", [round(float(a), 3) for a in res["accs"]], "nlls", [round(float(a), 3) for a in res["nlls"]], f"{time.time()-t0:.1f}s")
