"""Per-phase timeline of the fn1 engine (rank 0 / thread 0 clock64 stamps)."""
import os, sys, ctypes
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "blackbox-coresets-vi_b200"))
import numpy as np, torch
from psvi import _native as nat
nat.require_cuda()
C, M, S, T, B, H = 2, int(os.environ.get("M", 50)), 10, 100, 128, 100
torch.manual_seed(0)
model = nat.make_model([2, H, C], S)
P = int(nat.lib().psvi_mf_num_theta(model))
mu0 = (torch.randn(P, device="cuda") * 0.3).contiguous()
rho0 = torch.full((P,), float(np.log(np.expm1(1e-3))), device="cuda")
u = torch.randn(M, 2, device="cuda"); z = torch.randint(0, C, (M,), device="cuda", dtype=torch.int32)
v = torch.zeros(M, device="cuda")
xb = torch.randn(B, 2, device="cuda"); yb = torch.randint(0, C, (B,), device="cuda", dtype=torch.int32)
traj = torch.zeros(nat.traj_floats(model, T), device="cuda")
ug, vg, loss = torch.zeros(M, 2, device="cuda"), torch.zeros(M, device="cuda"), torch.zeros(1, device="cuda")
tl = torch.zeros(2 * 4096, device="cuda", dtype=torch.int64)
L = ctypes.CDLL(nat.LIB_PATH)
for i in range(3):
    mu, rho = mu0.clone(), rho0.clone()
    L.psvi_internal_set_timeline(ctypes.c_void_p(tl.data_ptr() if i == 2 else 0))
    nat.nested_step(model, nat.make_noise(None, seed=1, domain=7), mu, rho, u, z, v, xb, yb, B, 800.0, 1, 0.0, T, 1e-3,
                    1.0, 3, traj, None, ug, vg, None, loss, None)
torch.cuda.synchronize()
t = tl.cpu().numpy().reshape(-1, 2)
t = t[t[:, 1] > 0]
codes, clk = t[:, 0], t[:, 1]
# average delta from previous stamp, per code, over the steady state
import collections
d = collections.defaultdict(list)
for i in range(1, len(codes)):
    d[(int(codes[i - 1]), int(codes[i]))].append(clk[i] - clk[i - 1])
print(f"M={M}: total stamps {len(codes)}, total cycles {clk[-1]-clk[0]}")
for k in sorted(d):
    a = np.array(d[k][2:]) if len(d[k]) > 4 else np.array(d[k])
    print(f"  {k[0]:3d} -> {k[1]:3d}: n={len(d[k]):4d} median {np.median(a):8.0f} cyc  mean {a.mean():8.0f}")
