"""Timing of the tensor-core fn forward (cfg5 shapes): python scratch/bench_fn_tc.py [n_rows ...]"""
import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "blackbox-coresets-vi_b200"))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np, torch
from psvi import _native as nat

D, H, C, S, M = 256, 1024, 10, 64, 1000
rows = [int(a) for a in sys.argv[1:]] or [1000, 18944, 1 << 20]
dims = [D, H, C]
model = nat.make_model(dims, S)
P = nat.num_theta(model)
g = torch.Generator(device="cuda").manual_seed(0)
mu = torch.cat([torch.randn(H * D, device="cuda", generator=g) / D ** 0.5, torch.zeros(H, device="cuda"),
                torch.randn(C * H, device="cuda", generator=g) / H ** 0.5, torch.zeros(C, device="cuda")]).contiguous()
rho = torch.full((P,), float(np.log(np.expm1(1e-3))), device="cuda")
u = torch.randn(M, D, device="cuda", generator=g)
z = torch.randint(0, C, (M,), device="cuda", generator=g, dtype=torch.int32)
v = torch.zeros(M, device="cuda")
noise = nat.make_noise(None, seed=1, domain=0)
for n in rows:
    X = torch.randn(n, D, device="cuda", generator=g).bfloat16().contiguous()
    y = torch.randint(0, C, (n,), device="cuda", generator=g, dtype=torch.int32)
    out = torch.zeros(8, device="cuda")
    scratch = torch.zeros(nat.fn_tc_scratch_floats(model, n, M), device="cuda")
    for mode in (1, 0):
        for _ in range(2):
            nat.fn_predictive_tc(model, noise, mu, rho, u, z, v, X, y, 0, 1e7, 1, 0.0, mode, out, scratch)
        torch.cuda.synchronize()
        reps = 5 if n < 200000 else 2
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            nat.fn_predictive_tc(model, noise, mu, rho, u, z, v, X, y, 0, 1e7, 1, 0.0, mode, out, scratch)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        fl = 2.0 * S * (n + (M if mode == 0 else 0)) * (D * H + H * 16)
        print(f"rows={n} mode={mode} {ms:.3f} ms/call  {fl / ms / 1e9:.1f} TFLOP/s (incl. prep)  out={out[:5].tolist()}", flush=True)
