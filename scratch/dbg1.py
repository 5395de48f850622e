import sys, os
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "blackbox-coresets-vi_b200"))
import numpy as np, torch
from oracle import psvi_oracle as po
from tests.gpu_util import load, dev, zeros, rel_l2
from psvi import _native as nat
np.set_printoptions(precision=5, linewidth=200)
for name in ["logreg_hm_m10", "fn_fb_l2_m13"]:
    g, dims, S, T, eps = load(name)
    N, vmode = float(g["N"]), int(g["vmode"])
    model = nat.make_model(dims, S); P = nat.num_theta(model); M, D = g["u0"].shape
    mu, rho, u, z, v = dev(g["mu0"]), dev(g["rho0"]), dev(g["u0"]), dev(g["z"], torch.int32), dev(g["v0"])
    xb, yb = dev(g["xb"]), dev(g["yb"], torch.int32)
    gout, ug, vg, loss = zeros(nat.gout_floats(model, M)), zeros(M, D), zeros(M), zeros(1)
    noise = nat.make_noise(dev(eps[1][None]))
    nat.outer_grad(model, noise, mu, rho, u, z, v, xb, yb, xb.shape[0], N, vmode, 0.0, 1.0, gout, ug, vg, None, loss)
    torch.cuda.synchronize()
    a = po.coreset_weights(g["v0"], N, vmode)
    oval, gmu, grho, gu, ga, t = po.psvi_elbo_grad(g["mu0"], g["rho0"], eps[1].astype(np.float64), g["u0"], g["z"], a, g["xb"], g["yb"], N, dims)
    go = gout.cpu().numpy(); o = 2*P + M*D + M
    print(name, "loss", loss.item(), oval)
    print(" ds  ", go[o:o+S]); print(" ds* ", t["ds"])
    print(" w   ", go[o+S+4:o+2*S+4]); print(" w*  ", t["w"])
    w, e = t["w"], t["ds"] - t["ps"]; beta = w*(e-np.sum(w*e)) - 1.0/S
    print(" beta", go[o+2*S+4:o+3*S+4]); print(" b*  ", beta)
    print(" gp  ", go[o+3*S+4:o+4*S+4]); print(" gp* ", -w-beta)
    print(" terms", go[o+S:o+S+4])
    print(" gmu ", go[:8]); print(" gmu*", gmu[:8])
    print(" grho", go[P:P+8]); print(" gr* ", grho[:8])
    print(" ug rel", rel_l2(ug.cpu().numpy(), gu), "abar rel", rel_l2(go[2*P+M*D:2*P+M*D+M], ga))
