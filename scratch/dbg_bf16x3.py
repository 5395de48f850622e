"""Hypergradient error of a split-bf16 ("bf16x3") large-regime path vs the fp64 oracle as the problem grows (S, M), next to the
tf32x3 path: python scratch/dbg_bf16x3.py [tf32x3|bf16x3]  (PSVI.large_precision)."""
import sys, os, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "blackbox-coresets-vi_b200"))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np, torch
from oracle import psvi_oracle as po
from oracle.ref_import import NoiseFeeder
from psvi.experiments.experiments_utils import SynthDataset, make_synthetic_rows
from psvi import _native
mode = sys.argv[1] if len(sys.argv) > 1 else 'tf32x3'
print('arithmetic:', mode, flush=True)
from psvi.inference.psvi_classes import ExternalNoise, PSVILearnV
def rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64); return np.linalg.norm(a-b)/max(np.linalg.norm(b),1e-300)
def cos(a, b):
    a, b = np.asarray(a, np.float64).ravel(), np.asarray(b, np.float64).ravel(); return a@b/np.linalg.norm(a)/np.linalg.norm(b)
D, H, C, T, B = 128, 384, 3, 3, 64
for S, M, init_sd in [(3, 24, 1e-2), (8, 96, 1e-2), (16, 384, 1e-2), (16, 384, 1e-3)]:
    X, Y = make_synthetic_rows(4000, D, C, seed=0)
    tr, te = SynthDataset(X[:3000], Y[:3000].float()), SynthDataset(X[3000:], Y[3000:].float())
    kw = dict(mc_samples=S, num_epochs=0, data_minibatch=B, D=D, N=3000, inner_it=T, trainer="nested", log_every=10, lr0u=1e-4,
              lr0net=1e-3, lr0v=1e-3, init_args="subsample", init_sd=init_sd, num_pseudo=M, seed=0, architecture="fn",
              n_hidden=H, n_layers=1, logistic_regression=False, train_dataset=tr, test_dataset=te, dnm="synthetic", nc=C,
              compute_weights_entropy=True, register_elbos=False, quiet=True)
    obj = PSVILearnV(**kw); obj.large_precision = mode; obj.run_psvi(**kw)
    dims = obj.model.dims
    eps = NoiseFeeder.stream(dims, S, 77, T + 1)
    obj.noise_source = ExternalNoise(eps); obj.scheduler_optim_net = None
    mu, rho = [t.cpu().numpy().astype(np.float64) for t in obj.model.flat()]
    u0, z = obj.u.detach().cpu().numpy().astype(np.float64), obj.z.cpu().numpy()
    v0 = obj.v.detach().cpu().numpy().astype(np.float64)
    loss = obj.nested_step(X[:B].cuda(), Y[:B].cuda())
    e64 = [e.astype(np.float64) for e in eps]
    t0 = time.time()
    r = po.nested_step(mu, rho, np.stack(e64[:T]), e64[T], u0, z, v0, X[:B].numpy().astype(np.float64), Y[:B].numpy(), 3000.0, dims, 1e-3, vmode=1)
    gu, gv = obj.u.grad.cpu().numpy(), obj.v.grad.cpu().numpy()
    print(f"S={S} M={M} sd={init_sd}: loss rel {abs(loss.item()-r['loss'])/abs(r['loss']):.2e}  gu rel {rel(gu, r['u_grad']):.3f} cos {cos(gu, r['u_grad']):.4f}"
          f"  gv rel {rel(gv, r['v_grad']):.3f} cos {cos(gv, r['v_grad']):.4f}  |gu| {np.abs(r['u_grad']).mean():.3g} (oracle {time.time()-t0:.1f}s)", flush=True)
