import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "blackbox-coresets-vi_b200"))
import numpy as np, torch
from psvi import _native as nat
nat.LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'libpsvi_b200_prof.so')  # -DPSVI_FN_PROF build
D, H, C, S, M = 256, 1024, 10, 64, 1000
n = int(sys.argv[1]) if len(sys.argv) > 1 else 131072
model = nat.make_model([D, H, C], S)
P = nat.num_theta(model)
g = torch.Generator(device="cuda").manual_seed(0)
mu = (torch.randn(P, device="cuda", generator=g) * 0.05).contiguous()
rho = torch.full((P,), float(np.log(np.expm1(1e-3))), device="cuda")
X = torch.randn(n, D, device="cuda", generator=g).bfloat16().contiguous()
y = torch.randint(0, C, (n,), device="cuda", generator=g, dtype=torch.int32)
out = torch.zeros(8, device="cuda")
scratch = torch.zeros(nat.fn_tc_scratch_floats(model, n, M), device="cuda")
noise = nat.make_noise(None, seed=1, domain=0)
os.environ.pop("PSVI_FN_PROF", None)
for _ in range(2):
    nat.fn_predictive_tc(model, noise, mu, rho, None, None, None, X, y, 0, 1e7, 1, 0.0, 1, out, scratch)
torch.cuda.synchronize()
os.environ["PSVI_FN_PROF"] = "1"
nat.fn_predictive_tc(model, noise, mu, rho, None, None, None, X, y, 0, 1e7, 1, 0.0, 1, out, scratch)
torch.cuda.synchronize()
