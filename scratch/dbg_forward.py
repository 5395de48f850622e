import os, sys
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "blackbox-coresets-vi_b200"))
import torch
from psvi.models.neural_net import make_fcnet, VILinear
import torch.nn as nn
R, H, S, L = (int(a) for a in sys.argv[1:5])
torch.manual_seed(0)
net = make_fcnet(2, H, 2, n_layers=L, linear_class=VILinear, nonl_class=nn.ReLU, mc_samples=S, init_sd=1e-2).cuda()
x = torch.randn(R, 2, device="cuda")
lg = net(x); torch.cuda.synchronize()
print("fwd ok", R, H, S, L, tuple(lg.shape), float(lg.abs().mean()))
