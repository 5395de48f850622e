import os, sys
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "blackbox-coresets-vi_b200"))
import numpy as np, torch
from psvi import _native as nat
from psvi.experiments.experiments_utils import set_up_model
from psvi.inference.baselines import _make_trainer
M = int(sys.argv[1]) if len(sys.argv) > 1 else 400
H = int(sys.argv[2]) if len(sys.argv) > 2 else 16
S = int(sys.argv[3]) if len(sys.argv) > 3 else 5
torch.manual_seed(0)
net = set_up_model(architecture="fn", D=2, n_hidden=H, nc=2, mc_samples=S, init_sd=1e-2).cuda()
tr = _make_trainer(net, 0, None)
x = torch.randn(800, 2, device="cuda"); y = (x[:, 0] > 0).to(torch.int32)
for r0 in (0, M):
    l = tr.train(x[r0:r0 + M], y[r0:r0 + M], 800 / M, 1, 1e-2)
    torch.cuda.synchronize()
    print("train ok", r0, l.item())
lg = net(x[:M]); torch.cuda.synchronize(); print("fwd ok", lg.shape)
print(tr.test(x, y, M))
