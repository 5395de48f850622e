"""cfg2 nested step at the C ABI: specialised fn1 engine vs the generic cluster engine (PSVI_DISABLE_FN1=1)."""
import os, sys
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "blackbox-coresets-vi_b200"))
import numpy as np, torch
from psvi import _native as nat

def run(C=2, M=50, S=10, T=100, B=128, H=100, reps=20):
    torch.manual_seed(0)
    dims = [2, H, C]
    model = nat.make_model(dims, S)
    P = int(nat.lib().psvi_mf_num_theta(model))
    mu0 = (torch.randn(P, device="cuda") * 0.3).contiguous()
    rho0 = torch.full((P,), float(np.log(np.expm1(1e-3))), device="cuda")
    u = torch.randn(M, 2, device="cuda"); z = torch.randint(0, C, (M,), device="cuda", dtype=torch.int32)
    v = torch.zeros(M, device="cuda")
    xb = torch.randn(B, 2, device="cuda"); yb = torch.randint(0, C, (B,), device="cuda", dtype=torch.int32)
    traj = torch.zeros(nat.traj_floats(model, T), device="cuda")
    out = {}
    for name, flag in (("fn1", "0"), ("generic", "1")):
        os.environ["PSVI_DISABLE_FN1"] = flag
        ug, vg, loss = torch.zeros(M, 2, device="cuda"), torch.zeros(M, device="cuda"), torch.zeros(1, device="cuda")
        ts = []
        for i in range(reps + 3):
            mu, rho = mu0.clone(), rho0.clone()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            nat.nested_step(model, nat.make_noise(None, seed=1, domain=7), mu, rho, u, z, v, xb, yb, B, 800.0, 1, 0.0, T, 1e-3,
                            1.0, 3, traj, None, ug, vg, None, loss, None)
            e1.record(); torch.cuda.synchronize()
            if i >= 3:
                ts.append(e0.elapsed_time(e1))
        out[name] = (np.median(ts), ug.clone(), vg.clone(), loss.item(), mu.clone())
        print(f"C={C} M={M} S={S} T={T} H={H} {name}: median {np.median(ts):.3f} ms  min {np.min(ts):.3f} ms  -> {1e3/np.median(ts):.0f} steps/s")
    a, b = out["fn1"], out["generic"]
    rel = lambda x, y: float((x - y).norm() / y.norm())
    print(f"   fn1 vs generic: u_grad rel {rel(a[1], b[1]):.2e}  v_grad rel {rel(a[2], b[2]):.2e}  loss {a[3]:.5f} / {b[3]:.5f}  phi_T rel {rel(a[4], b[4]):.2e}")

if __name__ == "__main__":
    nat.require_cuda()
    run()
    run(C=4)
    run(M=10)
    run(T=10)
    run(H=40, S=4)
