import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "blackbox-coresets-vi_b200"))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np, torch
from oracle import psvi_oracle as po, psvi_oracle_generic as pg
from oracle.ref_import import NoiseFeeder
from psvi.experiments.experiments_utils import SynthDataset, make_synthetic_rows
from psvi.inference.psvi_classes import ExternalNoise, PSVILearnV
def rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64); return np.linalg.norm(a-b)/max(np.linalg.norm(b),1e-300)
D, H, C, S, M, T, B = 128, 384, 3, 3, 24, 3, 32
X, Y = make_synthetic_rows(600, D, C, seed=0)
tr, te = SynthDataset(X[:500], Y[:500].float()), SynthDataset(X[500:], Y[500:].float())
kw = dict(mc_samples=S, num_epochs=0, data_minibatch=B, D=D, N=500, inner_it=T, trainer="nested", log_every=10, lr0u=1e-4,
          lr0net=1e-3, lr0v=1e-3, init_args="subsample", init_sd=1e-2, num_pseudo=M, seed=0, architecture="fn",
          n_hidden=H, n_layers=1, logistic_regression=False, train_dataset=tr, test_dataset=te, dnm="synthetic", nc=C,
          compute_weights_entropy=True, register_elbos=False, quiet=True)
obj = PSVILearnV(**kw); obj.run_psvi(**kw)
dims = obj.model.dims
eps = NoiseFeeder.stream(dims, S, 77, T + 1)
eng = obj._stream(obj.model)
print(type(eng.net).__name__)
phi = eng.fam.get_phi()
mu, rho = [t.cpu().numpy().astype(np.float64) for t in obj.model.flat()]
fam = pg.MeanField(dims)
phi64 = np.concatenate([mu, rho])   # TL order [mu|rho]
# oracle works in torch param order; use po-level functions with mu, rho directly
u0, z = obj.u.detach().cpu().numpy().astype(np.float64), obj.z.cpu().numpy()
a = po.coreset_weights(obj.v.detach().cpu().numpy().astype(np.float64), 500.0, 1)
e0 = eps[0].astype(np.float64)
xb, yb = X[:B], Y[:B]
u_t, z32, a_t = obj._uv()[0], obj._z32(), obj._a()
# inner grad
val, g = eng.inner_grad(phi, torch.as_tensor(eps[0]).cuda(), u_t, z32, a_t)
v64, gmu, grho, gu, ga = po.inner_grad(mu, rho, e0, u0, z, a, dims)
print("inner val", val.item(), v64, "g", rel(g.cpu().numpy(), np.concatenate([gmu, grho])))
# outer grad
loss, pbar, ubar, abar, _ = eng.outer_grad(phi, torch.as_tensor(eps[0]).cuda(), u_t, z32, a_t, xb.cuda(), yb.cuda().int(), 500.0)
l64, mb, rb, ub, ab, _ = po.psvi_elbo_grad(mu, rho, e0, u0, z, a, xb.numpy().astype(np.float64), yb.numpy(), 500.0, dims)
print("outer loss", loss.item(), l64, "pbar", rel(pbar.cpu().numpy(), np.concatenate([mb, rb])), "ubar", rel(ubar.cpu().numpy(), ub),
      "abar", rel(abar.cpu().numpy(), ab))
print(" ubar col0 gpu", ubar.cpu().numpy()[:4, 0], "ref", ub[:4, 0])
# hvp along a direction like the reverse sweep's (scaled pbar)
gdir = (pbar / pbar.abs().max()).contiguous()
h, hu, ha = eng.hvp(phi, torch.as_tensor(eps[0]).cuda(), u_t, z32, a_t, gdir)
gd = gdir.cpu().numpy().astype(np.float64); P = len(mu)
hmu, hrho, hx, hc = po.inner_hvp(mu, rho, e0, u0, z, a, dims, gd[:P], gd[P:])
print("hvp h", rel(h.cpu().numpy(), np.concatenate([hmu, hrho])), "hu", rel(hu.cpu().numpy(), hx), "ha", rel(ha.cpu().numpy(), hc))
print(" hu col0 gpu", hu.cpu().numpy()[:4, 0], "ref", hx[:4, 0])
# random direction
rng = np.random.default_rng(0)
gd = rng.standard_normal(2 * P) * 1e-3
h, hu, ha = eng.hvp(phi, torch.as_tensor(eps[0]).cuda(), u_t, z32, a_t, torch.as_tensor(gd).float().cuda())
hmu, hrho, hx, hc = po.inner_hvp(mu, rho, e0, u0, z, a, dims, gd[:P], gd[P:])
print("hvp(rand) h", rel(h.cpu().numpy(), np.concatenate([hmu, hrho])), "hu", rel(hu.cpu().numpy(), hx), "ha", rel(ha.cpu().numpy(), hc))
