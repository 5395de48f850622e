import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "blackbox-coresets-vi_b200"))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
from psvi import _native as nat
from tests.test_gpu_fullsize_properties import _fulldata_case
from tests.gpu_util import zeros
g = torch.Generator(device="cuda").manual_seed(11)
D, H, C, S, n = 256, 1024, 10, 64, 262144
dims, P, mu, rho, x, y = _fulldata_case(D, H, C, S, n, g)
model = nat.make_model(dims, S)
noise = nat.make_noise(None, seed=1234, domain=7)
scratch = zeros(nat.fn_tc_scratch_floats(model, n, 0))
ones = torch.ones(n, device="cuda")
def run(xx, yy):
    ws, nk, nl = zeros(S), zeros(S), zeros(S, xx.shape[0])
    nat.fn_nll_tc(model, noise, mu, rho, xx, yy, ones[:xx.shape[0]].contiguous(), 0, ws, nk, nl, scratch)
    torch.cuda.synchronize()
    return ws, nk, nl
ws, nk, nl = run(x, y)
wsb, nkb, nlb = run(x, y)
print("repeat identical:", torch.equal(nl, nlb), torch.equal(ws, wsb))
for h in (100003, 100096, 128 * 700):
    _, _, nl1 = run(x[:h].contiguous(), y[:h].contiguous())
    _, _, nl2 = run(x[h:].contiguous(), y[h:].contiguous())
    d1 = (nl1 != nl[:, :h]); d2 = (nl2 != nl[:, h:])
    print(h, "first block differing:", int(d1.sum()), "second:", int(d2.sum()), "max abs", float((nl2 - nl[:, h:]).abs().max()),
          "rows with diffs in block 2:", int(d2.any(0).sum()), "samples:", int(d2.any(1).sum()))
    if d1.any():
        idx = d1.nonzero()[:5]; print(" block1 idx", idx.tolist(), float((nl1 - nl[:, :h]).abs().max()))
