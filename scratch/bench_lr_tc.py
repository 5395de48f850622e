import sys, os, json
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "blackbox-coresets-vi_b200"))
import numpy as np, torch
from psvi import _native as nat
nat.require_cuda()
D, C, S, M = 256, 10, int(os.environ.get("S", 10)), 50
n_rows = int(os.environ.get("ROWS", 8_000_000))
dims = [D, C]
model = nat.make_model(dims, S)
P = nat.num_theta(model)
g = torch.Generator(device="cuda").manual_seed(0)
mu = 0.1 * torch.randn(P, device="cuda", generator=g)
rho = torch.full((P,), float(np.log(np.expm1(0.05))), device="cuda")
u = torch.randn(M, D, device="cuda", generator=g); z = torch.randint(0, C, (M,), device="cuda", dtype=torch.int32, generator=g)
v = torch.zeros(M, device="cuda")
xb = torch.randn(n_rows, D, device="cuda", generator=g, dtype=torch.bfloat16)
y = torch.randint(0, C, (n_rows,), device="cuda", dtype=torch.int32, generator=g)
out = torch.zeros(8, device="cuda"); scratch = torch.zeros(nat.lr_predictive_tc_scratch_floats(model), device="cuda")
noise = nat.make_noise(None, seed=1, domain=1)
def run():
    nat.lr_predictive_tc(model, noise, mu, rho, u, z, v, xb, y, 0, 10000.0, 1, 0.0, 0, out, scratch)
for _ in range(3): run()
torch.cuda.synchronize()
ts = []
for _ in range(10):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); run(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
ts.sort()
byt = n_rows * (D * 2 + 4)
print(json.dumps({"rows": n_rows, "S": S, "ms_min": ts[0], "ms_med": ts[len(ts)//2], "GBps_med": byt / ts[len(ts)//2] / 1e6,
                  "GBps_best": byt / ts[0] / 1e6, "frac_of_6556": byt / ts[len(ts)//2] / 1e6 / 6556.2,
                  "row_samples_per_s": n_rows * S / (ts[len(ts)//2] * 1e-3), "out": out.cpu().tolist()[:5]}))
