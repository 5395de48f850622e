"""psvi_fn_data_grad_tc at cfg5 shapes: time and algorithmic TFLOP/s (3 F_fwd per row)."""
import os, sys
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "blackbox-coresets-vi_b200"))
import torch
from psvi import _native as nat
nat.require_cuda()
D, H, C, S = 256, 1024, 10, 64
P = H * D + H + C * H + C
model = nat.make_model([D, H, C], S)
g = torch.Generator(device="cuda").manual_seed(0)
theta = torch.cat([torch.randn(S, H * D, device="cuda", generator=g) / D ** 0.5, torch.zeros(S, H, device="cuda"),
                   torch.randn(S, C * H, device="cuda", generator=g) / H ** 0.5, torch.zeros(S, C, device="cuda")], 1).contiguous()
coef = torch.ones(S, device="cuda")
for R in [int(a) for a in (sys.argv[1:] or ["131072", "524288"])]:
    x = torch.randn(R, D, device="cuda", generator=g, dtype=torch.bfloat16)
    y = torch.randint(0, C, (R,), device="cuda", dtype=torch.int32, generator=g)
    dsum, tbar = torch.zeros(S, device="cuda"), torch.zeros(S, P, device="cuda")
    scr = torch.zeros(nat.fn_data_grad_scratch_floats(model, R), device="cuda")
    for _ in range(2):
        nat.fn_data_grad_tc(model, theta, x, y, coef, dsum, tbar, scr)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 3
    a.record()
    for _ in range(reps):
        nat.fn_data_grad_tc(model, theta, x, y, coef, dsum, tbar, scr)
    b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b) / reps
    fl = 3 * 2.0 * S * R * (D * H + H * C)
    print(f"R={R}: {ms:.2f} ms per call, {fl / ms / 1e9:.1f} algorithmic TFLOP/s, scratch {scr.numel() * 4 / 2**30:.2f} GiB")
    del x, y, scr
