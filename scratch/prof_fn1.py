"""Three cfg2 nested steps through the C ABI (for ncu: -k regex:fn1 -s 2 -c 1)."""
import os, sys
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "blackbox-coresets-vi_b200"))
import numpy as np, torch
from psvi import _native as nat
nat.require_cuda()
C, M, S, T, B, H = 2, 50, 10, 100, 128, 100
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
for i in range(3):
    mu, rho = mu0.clone(), rho0.clone()
    nat.nested_step(model, nat.make_noise(None, seed=1, domain=7), mu, rho, u, z, v, xb, yb, B, 800.0, 1, 0.0, T, 1e-3,
                    1.0, 3, traj, None, ug, vg, None, loss, None)
torch.cuda.synchronize()
print("ok", loss.item())
