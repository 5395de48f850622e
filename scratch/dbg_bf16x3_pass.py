import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "blackbox-coresets-vi_b200"))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np, torch
from oracle import psvi_oracle as po
from tests.gpu_util import dev, rel_l2, zeros
from psvi import _native as nat
for (D, H, C, S, R) in [(128, 256, 3, 3, 300), (128, 256, 10, 2, 300), (128, 256, 10, 3, 100), (64, 128, 10, 2, 100), (64, 128, 8, 2, 100), (64, 128, 9, 2, 100), (64, 128, 16, 2, 100)]:
    rng = np.random.default_rng(D + H + C + S + R)
    dims = [D, H, C]
    theta = np.concatenate([rng.standard_normal((S, H * D)) / np.sqrt(D), 0.1 * rng.standard_normal((S, H)),
                            rng.standard_normal((S, C * H)) / np.sqrt(H), 0.1 * rng.standard_normal((S, C))], 1).astype(np.float32)
    X = rng.standard_normal((R, D)).astype(np.float32)
    y = rng.integers(0, C, R)
    cw = rng.uniform(0.5, 1.5, (S, R)).astype(np.float32)
    model = nat.make_model(dims, S)
    P = theta.shape[1]
    th, x_, y_, cw_ = dev(theta), dev(X), dev(y, torch.int32), dev(cw)
    t64, X64, cw64 = theta.astype(np.float64), X.astype(np.float64), cw.astype(np.float64)
    o, cache = po.mlp_forward(t64, X64, dims)
    ref_nll, p = po.nll_rows(o, y)
    q = p.copy(); q[:, np.arange(R), y] -= 1.0
    At, Ax = po.mlp_backward(t64, cache, dims, cw64[:, :, None] * q)
    blocks = ((0, H * D), (H * D, H * D + H), (H * D + H, H * D + H + C * H), (H * D + H + C * H, P))
    for prec in (nat.PREC_TF32X3, nat.PREC_BF16X3):
        nll, logits, tbar, xbar = zeros(S, R), zeros(S, R, C), zeros(S, P), zeros(S, R, D)
        nat.fnl_pass(model, prec, th, None, x_, y_, None, nll=nll, logits=logits)
        nat.fnl_pass(model, prec, th, None, x_, y_, cw_, nll=nll, tbar=tbar, xbar=xbar)
        torch.cuda.synchronize()
        print((D, H, C, S, R), "prec", prec, "logits %.1e" % rel_l2(logits.cpu().numpy(), o),
              "tbar W1 %.1e b1 %.1e W2 %.1e b2 %.1e" % tuple(rel_l2(tbar.cpu().numpy()[:, lo:hi], At[:, lo:hi]) for lo, hi in blocks),
              "xbar %.1e" % rel_l2(xbar.cpu().numpy(), Ax), flush=True)
