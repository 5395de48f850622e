"""Small invocations of the kernels written late in the round, for compute-sanitizer (memcheck):
compute-sanitizer --tool memcheck python scratch/sanitize_small.py"""
import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "blackbox-coresets-vi_b200"))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import math
import torch
from psvi import _native as nat
from psvi.inference.stream import FnLargeNet, FullCovFamily, LenetNet, MlpNet
from psvi.models.neural_net import make_fc2net
g = torch.Generator(device="cuda").manual_seed(0)
z = lambda *s: torch.zeros(*s, device="cuda")
rn = lambda *s: torch.randn(*s, device="cuda", generator=g)
# lenet: odd row counts (partial image groups, partial row chunks)
for S, R in ((1, 1), (3, 7), (2, 11)):
    net = LenetNet(S)
    P = nat.lenet_num_theta()
    th, thd = (rn(S, P) * 0.05).contiguous(), (rn(S, P) * 0.01).contiguous()
    x, y = rn(R, 784).contiguous(), torch.randint(0, 10, (R,), device="cuda", generator=g, dtype=torch.int32)
    cw = (0.5 + torch.rand(S, R, device="cuda", generator=g)).contiguous()
    net.pass_(th, None, x, y, cw, nll=z(S, R), tbar=z(S, P), xbar=z(S, R, 784))
    net.pass_(th, thd, x, y, cw, tbar=z(S, P), tdbar=z(S, P), xbar=z(S, R, 784), acbar=z(S, R))
    net.logits(th, x)
torch.cuda.synchronize(); print("lenet ok", flush=True)
# unrolled-Adam kernels, odd length
n = 1003
phi, gg, m, v = rn(n), rn(n), rn(n) * 0.1, torch.rand(n, device="cuda", generator=g)
p2, m2, v2 = nat.adam_unroll_step(phi, gg, m, v, 1e-3 / (1 - 0.9), math.sqrt(1 - 0.999))
nat.adam_unroll_reverse(rn(n), gg, m2, v2, rn(n), rn(n), 1e-3 / (1 - 0.9), math.sqrt(1 - 0.999))
torch.cuda.synchronize(); print("adam ok", flush=True)
# full-covariance family maps
fam = FullCovFamily(make_fc2net(2, 7, 2, n_layers=2, mc_samples=5, init_sd=1e-2).cuda())
phi = fam.get_phi().contiguous()
eps, A, Ad = rn(5, fam.Pt).contiguous(), rn(5, fam.Pt).contiguous(), rn(5, fam.Pt).contiguous()
pd = (rn(phi.numel()) * 0.1).contiguous()
fam.sample(phi, eps); fam.tangent(phi, pd, eps); fam.grad(phi, eps, A, 1.0, 0.5); fam.hvp(phi, pd, eps, A, Ad)
torch.cuda.synchronize(); print("fullcov ok", flush=True)
# large-regime pass, all three precisions, rows not a multiple of the tile
D, H, C, S, R = 64, 128, 3, 2, 37
for prec in (nat.PREC_BF16, nat.PREC_TF32X3, nat.PREC_BF16X3):
    net = FnLargeNet([D, H, C], S, precision=prec)
    P = H * D + H + C * H + C
    th, thd = (rn(S, P) * 0.1).contiguous(), (rn(S, P) * 0.01).contiguous()
    x, y = rn(R, D).contiguous(), torch.randint(0, C, (R,), device="cuda", generator=g, dtype=torch.int32)
    cw = (0.5 + torch.rand(S, R, device="cuda", generator=g)).contiguous()
    net.pass_(th, None, x, y, cw, nll=z(S, R), tbar=z(S, P), xbar=z(S, R, D))
    net.pass_(th, thd, x, y, cw, tbar=z(S, P), tdbar=z(S, P), xbar=z(S, R, D), acbar=z(S, R))
torch.cuda.synchronize(); print("large ok", flush=True)
