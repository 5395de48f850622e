"""cfg5 nested step timing: fn D=256 H=1024 C=10, M=1000, S=64, B=128 (per-rank shard of the minibatch), T in argv."""
import sys, os, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "blackbox-coresets-vi_b200"))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
from psvi.experiments.experiments_utils import SynthDataset, make_synthetic_rows
from psvi.inference.psvi_classes import PSVILearnV
T = int(sys.argv[1]) if len(sys.argv) > 1 else 10
S = int(sys.argv[2]) if len(sys.argv) > 2 else 64
D, H, C, M, B = 256, 1024, 10, 1000, 128
X, Y = make_synthetic_rows(20000, D, C, seed=0)
tr, te = SynthDataset(X[:16000], Y[:16000].float()), SynthDataset(X[16000:], Y[16000:].float())
kw = dict(mc_samples=S, num_epochs=0, data_minibatch=B, D=D, N=16000, inner_it=T, trainer="nested", log_every=1000, lr0u=1e-4,
          lr0net=1e-3, lr0v=1e-3, init_args="subsample", init_sd=1e-3, num_pseudo=M, seed=0, architecture="fn", n_hidden=H,
          n_layers=1, logistic_regression=False, train_dataset=tr, test_dataset=te, dnm="synthetic", nc=C,
          compute_weights_entropy=False, register_elbos=False, quiet=True)
obj = PSVILearnV(**kw)
obj.large_precision = sys.argv[3] if len(sys.argv) > 3 else 'tf32x3'
print('precision', obj.large_precision)
obj.run_psvi(**kw)
xb, yb = obj._next_minibatch()
obj.nested_step(xb, yb)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
n = int(os.environ.get("REPS", "2"))
e0.record()
for _ in range(n):
    loss = obj.nested_step(xb, yb)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / n
f = lambda R: 2.0 * S * R * (D * H + H * C)
fl = T * 9 * f(M) + 3 * f(M + B)
print(f"cfg5 nested_step T={T} S={S}: {ms:.1f} ms/outer step, {fl/ms/1e9:.1f} TFLOP/s algorithmic, loss {loss.item():.1f}, "
      f"mem {torch.cuda.max_memory_allocated()/2**30:.1f} GiB", flush=True)
t0 = time.time(); acc, nll, *_ = obj.evaluate(); torch.cuda.synchronize()
print(f"evaluate 4000 rows: {1e3*(time.time()-t0):.1f} ms acc {acc.item():.3f} nll {nll.item():.3f}")
# ---- phase timing
import collections
eng = obj._stream(obj.model)
acc_t = collections.defaultdict(float)
def wrap(name):
    f = getattr(eng, name)
    def g(*a, **k):
        torch.cuda.synchronize(); t0 = time.time()
        r = f(*a, **k)
        torch.cuda.synchronize(); acc_t[name] += time.time() - t0
        return r
    setattr(eng, name, g)
for nme in ("inner_grad", "outer_grad", "hvp"):
    wrap(nme)
orig_noise = obj._noise_tensor
def nt(*a, **k):
    torch.cuda.synchronize(); t0 = time.time(); r = orig_noise(*a, **k); torch.cuda.synchronize(); acc_t["noise"] += time.time() - t0; return r
obj._noise_tensor = nt
torch.cuda.synchronize(); t0 = time.time()
obj.nested_step(xb, yb)
torch.cuda.synchronize(); tot = time.time() - t0
print("phases (ms):", {k: round(1e3 * v, 1) for k, v in acc_t.items()}, "total", round(1e3 * tot, 1))
