"""CPU oracle for the convolutional family (lenet) -- TEST INFRASTRUCTURE ONLY (rules: oracle/psvi_oracle.py header).

numpy restatement of the reference's lenet path:
  * make_lenet                psvi/models/neural_net.py:334-359  conv(1->6,5,pad 2) ReLU pool2 conv(6->16,5) ReLU pool2
                              Flatten fc(400->120) ReLU fc(120->84) ReLU fc(84->10)
  * VIConv2d.forward          :194-246  per-sample convolution (the reference stacks samples on channels, groups = S)
  * BatchMaxPool2d            :249-255  2x2 max-pool over the flattened (S, N) batch
  * quirks Q4 / Q5: the last VILinear is built without kwargs -> mc_samples = 1 (ONE weight draw shared
    by all S samples, scalar sampled_nkl broadcast over s) and init_sd = 0.01; the KL / sampled-nkl sums run over VILinear
    layers only, so the two conv layers contribute neither.
The network is exposed through the four maps the family-generic objectives need (forward / backward / dual_forward /
dual_backward, same contracts as psvi_oracle.mlp_*), the family through oracle.psvi_oracle_generic's interface.
theta layout (TL): per layer weight then bias, module order; phi = torch parameters_to_vector order
(weight, bias, _weight_sd, _bias_sd per layer).  Pinned against the reference in tests/golden/lenet_*.npz.
"""
from __future__ import annotations

import numpy as np

from oracle import psvi_oracle as po

# (kind, weight shape, bias size)
LAYERS = [("conv", (6, 1, 5, 5), 6, 2), ("conv", (16, 6, 5, 5), 16, 0), ("fc", (120, 400), 120, None), ("fc", (84, 120), 84, None),
          ("fc", (10, 84), 10, None)]
SIZES = [(int(np.prod(w)), b) for _, w, b, _ in LAYERS]
P = sum(nw + nb for nw, nb in SIZES)          # 61 706
KL_LAYERS = (2, 3, 4)                         # VILinear layers only (Q5)
SHARED_LAYER = 4                              # mc_samples = 1 (Q4)


def layer_slices():
    out, off = [], 0
    for nw, nb in SIZES:
        out.append((slice(off, off + nw), slice(off + nw, off + nw + nb)))
        off += nw + nb
    return out


def kl_mask(dtype=np.float64):
    m = np.zeros(P, dtype)
    for l in KL_LAYERS:
        ws, bs = layer_slices()[l]
        m[ws] = 1
        m[bs] = 1
    return m


def share_last_layer(eps):
    """eps [..., S, P] with the last layer's block replaced by sample 0's draw (the single draw of an mc_samples=1 layer)."""
    eps = eps.copy()
    ws, bs = layer_slices()[SHARED_LAYER]
    eps[..., :, ws] = eps[..., :1, ws]
    eps[..., :, bs] = eps[..., :1, bs]
    return eps


# ------------------------------------------------------------------------------------------------ conv as a linear map
def _cols(x, pad):
    """x [..., C, H, W] -> [..., Ho, Wo, C, 5, 5] windows."""
    if pad:
        x = np.pad(x, [(0, 0)] * (x.ndim - 2) + [(pad, pad), (pad, pad)])
    w = np.lib.stride_tricks.sliding_window_view(x, (5, 5), axis=(-2, -1))      # [..., C, Ho, Wo, 5, 5]
    return np.moveaxis(w, -5, -3)


def conv_apply(W, x, pad):
    """W [S, Co, Ci, 5, 5], x [S or 1, R, Ci, H, W] -> [S, R, Co, Ho, Wo]  (F.conv2d with groups = S, neural_net.py:233-241)."""
    return np.einsum("srhwcij,socij->srohw", _cols(x, pad), W, optimize=True)


def conv_applyT(W, g, pad, hin):
    """adjoint of conv_apply wrt x: g [S, R, Co, Ho, Wo] -> [S, R, Ci, hin, hin]."""
    S, R, Co, Ho, Wo = g.shape
    Ci = W.shape[2]
    out = np.zeros((S, R, Ci, hin + 2 * pad, hin + 2 * pad), g.dtype)
    for i in range(5):
        for j in range(5):
            out[:, :, :, i:i + Ho, j:j + Wo] += np.einsum("srohw,soc->srchw", g, W[:, :, :, i, j], optimize=True)
    return out[:, :, :, pad:pad + hin, pad:pad + hin] if pad else out


def conv_wgrad(g, x, pad):
    """adjoint of conv_apply wrt W: -> [S, Co, Ci, 5, 5]."""
    return np.einsum("srohw,srhwcij->socij", g, np.broadcast_to(_cols(x, pad), (g.shape[0],) + _cols(x, pad).shape[1:]), optimize=True)


# ------------------------------------------------------------------------------------------------ ReLU + 2x2 max-pool
def _windows(a):
    S, R, C, H, W = a.shape
    return a.reshape(S, R, C, H // 2, 2, W // 2, 2).transpose(0, 1, 2, 3, 5, 4, 6).reshape(S, R, C, H // 2, W // 2, 4)


def relu_pool(a):
    """-> pooled, (k, mask): k = first argmax of relu(a) in the window (torch max_pool2d), mask = max > 0 (ReLU backward)."""
    w = np.maximum(_windows(a), 0.0)
    k = w.argmax(-1)
    p = np.take_along_axis(w, k[..., None], -1)[..., 0]
    return p, (k, p > 0)


def pool_select(a, sel):
    k, m = sel
    return np.take_along_axis(_windows(a), k[..., None], -1)[..., 0] * m


def pool_scatter(g, sel):
    k, m = sel
    S, R, C, Hq, Wq = g.shape
    w = np.zeros((S, R, C, Hq, Wq, 4), g.dtype)
    np.put_along_axis(w, k[..., None], (g * m)[..., None], -1)
    return w.reshape(S, R, C, Hq, Wq, 2, 2).transpose(0, 1, 2, 3, 5, 4, 6).reshape(S, R, C, 2 * Hq, 2 * Wq)


# ------------------------------------------------------------------------------------------------ the network
def _unpack(theta):
    S = theta.shape[0]
    out = []
    for (kind, wshape, nb, _), (ws, bs) in zip(LAYERS, layer_slices()):
        out.append((theta[:, ws].reshape((S,) + wshape), theta[:, bs]))
    return out


class LeNet:
    """forward / backward / dual_forward / dual_backward on sampled weights theta [S, P]; X [R, 784]."""
    P = P

    @staticmethod
    def forward(theta, X):
        S = theta.shape[0]
        (W1, b1), (W2, b2), (W3, b3), (W4, b4), (W5, b5) = _unpack(theta)
        x0 = X.reshape(1, X.shape[0], 1, 28, 28)
        a1 = conv_apply(W1, x0, 2) + b1[:, None, :, None, None]
        p1, s1 = relu_pool(a1)
        a2 = conv_apply(W2, p1, 0) + b2[:, None, :, None, None]
        p2, s2 = relu_pool(a2)
        f = p2.reshape(S, X.shape[0], 400)
        a3 = np.einsum("sri,soi->sro", f, W3) + b3[:, None, :]
        h3 = np.maximum(a3, 0)
        a4 = np.einsum("sri,soi->sro", h3, W4) + b4[:, None, :]
        h4 = np.maximum(a4, 0)
        o = np.einsum("sri,soi->sro", h4, W5) + b5[:, None, :]
        return o, dict(x0=x0, p1=p1, s1=s1, p2=p2, s2=s2, f=f, h3=h3, h4=h4)

    @staticmethod
    def _back(theta, thetad, c, G5, G5d):
        """Shared backward: adjoints G (wrt primal pre-activations) and Gd (wrt tangent pre-activations); thetad / Gd may
        be None (plain gradient).  Returns A_theta, A_thetadot (or None), A_x."""
        S, R = G5.shape[:2]
        Ws = _unpack(theta)
        Wd = _unpack(thetad) if thetad is not None else [(None, None)] * 5
        At = np.zeros_like(theta)
        Atd = np.zeros_like(theta) if thetad is not None else None
        sl = layer_slices()
        dual = thetad is not None
        # fully connected layers 5, 4, 3
        ins = [None, None, (c["f"], c.get("fd")), (c["h3"], c.get("h3d")), (c["h4"], c.get("h4d"))]
        masks = [None, None, None, c["h3"] > 0, c["h4"] > 0]
        G, Gd = G5, G5d
        for l in (4, 3, 2):
            x, xd = ins[l]
            ws, bs = sl[l]
            gw = np.einsum("sro,sri->soi", G, x)
            if dual:
                gw = gw + np.einsum("sro,sri->soi", Gd, xd)
                Atd[:, ws] = np.einsum("sro,sri->soi", Gd, x).reshape(S, -1)
                Atd[:, bs] = Gd.sum(1)
            At[:, ws] = gw.reshape(S, -1)
            At[:, bs] = G.sum(1)
            Ax = np.einsum("sro,soi->sri", G, Ws[l][0])
            if dual:
                Ax = Ax + np.einsum("sro,soi->sri", Gd, Wd[l][0])
                Axd = np.einsum("sro,soi->sri", Gd, Ws[l][0])
            if masks[l] is not None:
                Ax = Ax * masks[l]
                if dual:
                    Axd = Axd * masks[l]
            G, Gd = Ax, (Axd if dual else None)
        # conv 2
        G = pool_scatter(G.reshape(S, R, 16, 5, 5), c["s2"])
        Gd = pool_scatter(Gd.reshape(S, R, 16, 5, 5), c["s2"]) if dual else None
        ws, bs = sl[1]
        gw = conv_wgrad(G, c["p1"], 0)
        if dual:
            gw = gw + conv_wgrad(Gd, c["p1d"], 0)
            Atd[:, ws] = conv_wgrad(Gd, c["p1"], 0).reshape(S, -1)
            Atd[:, bs] = Gd.sum((1, 3, 4))
        At[:, ws] = gw.reshape(S, -1)
        At[:, bs] = G.sum((1, 3, 4))
        Ap = conv_applyT(Ws[1][0], G, 0, 14)
        if dual:
            Ap = Ap + conv_applyT(Wd[1][0], Gd, 0, 14)
            Apd = conv_applyT(Ws[1][0], Gd, 0, 14)
        # conv 1
        G = pool_scatter(Ap, c["s1"])
        Gd = pool_scatter(Apd, c["s1"]) if dual else None
        ws, bs = sl[0]
        At[:, ws] = conv_wgrad(G, c["x0"], 2).reshape(S, -1)
        At[:, bs] = G.sum((1, 3, 4))
        Ax = conv_applyT(Ws[0][0], G, 2, 28)
        if dual:
            Atd[:, ws] = conv_wgrad(Gd, c["x0"], 2).reshape(S, -1)
            Atd[:, bs] = Gd.sum((1, 3, 4))
            Ax = Ax + conv_applyT(Wd[0][0], Gd, 2, 28)
        return At, Atd, Ax.reshape(S, R, 784)

    @staticmethod
    def backward(theta, cache, obar):
        At, _, Ax = LeNet._back(theta, None, cache, obar, None)
        return At, Ax

    @staticmethod
    def dual_forward(theta, thetad, X):
        S, R = theta.shape[0], X.shape[0]
        o, c = LeNet.forward(theta, X)
        Ws, Wd = _unpack(theta), _unpack(thetad)
        a1d = conv_apply(Wd[0][0], c["x0"], 2) + Wd[0][1][:, None, :, None, None]
        p1d = pool_select(a1d, c["s1"])
        a2d = conv_apply(Ws[1][0], p1d, 0) + conv_apply(Wd[1][0], c["p1"], 0) + Wd[1][1][:, None, :, None, None]
        p2d = pool_select(a2d, c["s2"])
        fd = p2d.reshape(S, R, 400)
        h3d = (np.einsum("sri,soi->sro", fd, Ws[2][0]) + np.einsum("sri,soi->sro", c["f"], Wd[2][0]) + Wd[2][1][:, None, :]) * (c["h3"] > 0)
        h4d = (np.einsum("sri,soi->sro", h3d, Ws[3][0]) + np.einsum("sri,soi->sro", c["h3"], Wd[3][0]) + Wd[3][1][:, None, :]) * (c["h4"] > 0)
        od = np.einsum("sri,soi->sro", h4d, Ws[4][0]) + np.einsum("sri,soi->sro", c["h4"], Wd[4][0]) + Wd[4][1][:, None, :]
        c.update(p1d=p1d, fd=fd, h3d=h3d, h4d=h4d)
        return o, od, c

    @staticmethod
    def dual_backward(theta, thetad, cache, A_o, A_od):
        return LeNet._back(theta, thetad, cache, A_o, A_od)


# ------------------------------------------------------------------------------------------------ the family
class LeNetMeanField:
    """Mean-field family of make_lenet: phi in torch parameters_to_vector order; conv layers carry no KL / nkl (Q5); the last
    layer's noise is one draw shared by all samples (Q4; pass eps through share_last_layer)."""
    net = LeNet

    def __init__(self):
        self.Pt, self.Pphi = P, 2 * P
        self.mask = kl_mask()

    def split(self, phi):
        mu, rho, off = [], [], 0
        for nw, nb in SIZES:
            mu += [phi[off:off + nw], phi[off + nw:off + nw + nb]]
            rho += [phi[off + nw + nb:off + 2 * nw + nb], phi[off + 2 * nw + nb:off + 2 * (nw + nb)]]
            off += 2 * (nw + nb)
        return np.concatenate(mu), np.concatenate(rho)

    def join(self, gmu, grho):
        out, off = [], 0
        for nw, nb in SIZES:
            out += [gmu[off:off + nw + nb], grho[off:off + nw + nb]]
            off += nw + nb
        return np.concatenate(out)

    def sample(self, phi, eps):
        mu, rho = self.split(phi)
        return po.mf_sample(mu, rho, eps)

    def tangent(self, phi, phidot, eps):
        mu, rho = self.split(phi)
        md, rd = self.split(phidot)
        return md[None] + (po.sigmoid(rho) * rd)[None] * eps

    def kl(self, phi):
        mu, rho = self.split(phi)
        sg = po.softplus(rho)
        return np.sum(self.mask * (0.5 * (sg * sg + mu * mu - 1.0) - np.log(sg)))

    def nkl(self, phi, eps, theta):
        mu, rho = self.split(phi)
        sg = po.softplus(rho)
        return np.sum(self.mask[None] * (-0.5 * theta * theta + 0.5 * eps * eps + np.log(sg)[None, :]), axis=1)

    def nkl_theta_grad(self, theta):
        """d nkl_s / d theta_s (through the sample)."""
        return -theta * self.mask[None]

    def grad(self, phi, eps, tbar, kl_coef, nkl_coef):
        mu, rho = self.split(phi)
        sg, sig = po.softplus(rho), po.sigmoid(rho)
        gmu = tbar.sum(0) + kl_coef * self.mask * mu
        grho = sig * ((tbar * eps).sum(0) + self.mask * (kl_coef * (sg - 1 / sg) + nkl_coef / sg))
        return self.join(gmu, grho)

    def hvp(self, phi, phidot, eps, A_t, A_td):
        mu, rho = self.split(phi)
        md, rd = self.split(phidot)
        sg, sig = po.softplus(rho), po.sigmoid(rho)
        hmu = A_t.sum(0) + self.mask * md
        hrho = (sig * (A_t * eps).sum(0) + sig * (1 - sig) * rd * (A_td * eps).sum(0)
                + self.mask * ((1 + 1 / (sg * sg)) * sig * sig + (sg - 1 / sg) * sig * (1 - sig)) * rd)
        return self.join(hmu, hrho)
