"""cfg5-style precision study on the pinned fn oracle (oracle/psvi_oracle.py, fp64): operand formats per PASS TYPE.
The bilevel step has two kinds of network passes: the inner-gradient pass (trajectory; its g_i end up in Adam's denominators)
and the Hessian-vector ("dual") pass of the reverse sweep.  Question: can the dual pass run in the cheaper split-bf16 format
while the gradient pass keeps tf32x3?  Operands of every layer product are rounded (hi, lo pairs, three products); accumulation
exact.  usage: python scratch/fn_mixed_precision_study.py [D H C S M T B init_sd]"""
import os, sys, time
sys.path.insert(0, os.getcwd())
import numpy as np
from oracle import psvi_oracle as po

D, H, C, S, M, T, B = (int(a) for a in (sys.argv[1:8] if len(sys.argv) > 7 else "32 128 10 8 96 10 32".split()))
init_sd = float(sys.argv[8]) if len(sys.argv) > 8 else 1e-2
dims = [D, H, C]
rng = np.random.default_rng(0)
P = po.p_theta(dims)
mu = np.concatenate([rng.uniform(-1, 1, H * D) / np.sqrt(D), rng.uniform(-1, 1, H) / np.sqrt(D),
                     rng.uniform(-1, 1, C * H) / np.sqrt(H), rng.uniform(-1, 1, C) / np.sqrt(H)])
rho = np.full(P, np.log(np.expm1(init_sd)))
Wt = rng.standard_normal((D, C))
u = rng.standard_normal((M, D)); z = (u @ Wt).argmax(1).astype(np.float64)
xb = rng.standard_normal((B, D)); yb = (xb @ Wt).argmax(1).astype(np.float64)
v = 0.3 * rng.standard_normal(M)
eps = rng.standard_normal((T + 1, S, P))
N = 10000.0

def rnd_bits(x, keep):
    x32 = np.ascontiguousarray(x, dtype=np.float32)
    uu = x32.view(np.uint32).astype(np.uint64)
    drop = 23 - keep
    uu = (uu + ((1 << (drop - 1)) - 1) + ((uu >> drop) & 1)) >> drop << drop
    return uu.astype(np.uint32).view(np.float32).astype(np.float64)

def split(x, keep):
    hi = rnd_bits(x, keep)
    return hi, rnd_bits(x - hi, keep)

real_einsum = np.einsum
MODE = ["exact"]
def ein(expr, *ops, **kw):
    m = MODE[0]
    if isinstance(m, dict):
        m = m.get(expr, m["default"])
    if len(ops) != 2 or m == "exact":
        return real_einsum(expr, *ops, **kw)
    A, Bm = ops
    if m == "fp32":
        return real_einsum(expr, A.astype(np.float32).astype(np.float64), Bm.astype(np.float32).astype(np.float64), **kw)
    keep = {"bf16": 7, "bf16x3": 7, "tf32x3": 10}[m]
    if m == "bf16":
        return real_einsum(expr, rnd_bits(A, keep), rnd_bits(Bm, keep), **kw)
    Ah, Al = split(A, keep); Bh, Bl = split(Bm, keep)
    return real_einsum(expr, Ah, Bh, **kw) + real_einsum(expr, Ah, Bl, **kw) + real_einsum(expr, Al, Bh, **kw)

real_hvp, real_grad, real_outer = po.inner_hvp, po.inner_grad, po.psvi_elbo_grad
def run(grad_mode, hvp_mode, outer_mode=None):
    outer_mode = grad_mode if outer_mode is None else outer_mode
    def w(fn, mode):
        def f(*a, **k):
            old = MODE[0]; MODE[0] = mode
            try: return fn(*a, **k)
            finally: MODE[0] = old
        return f
    po.inner_hvp, po.inner_grad, po.psvi_elbo_grad = w(real_hvp, hvp_mode), w(real_grad, grad_mode), w(real_outer, outer_mode)
    np.einsum = ein
    try:
        return po.nested_step(mu, rho, eps[:T], eps[T], u, z, v, xb, yb, N, dims, 1e-3, vmode=1)
    finally:
        np.einsum = real_einsum
        po.inner_hvp, po.inner_grad, po.psvi_elbo_grad = real_hvp, real_grad, real_outer

def rel(a, b): return np.linalg.norm(a - b) / np.linalg.norm(b)
def cos(a, b): return float(np.dot(a.ravel(), b.ravel()) / np.linalg.norm(a) / np.linalg.norm(b))
t0 = time.time(); ex = run("exact", "exact")
print(f"case D={D} H={H} C={C} S={S} M={M} T={T} B={B} init_sd={init_sd}: loss {ex['loss']:.4f} ({time.time()-t0:.1f}s per run)")
print("| gradient pass | dual pass | outer pass | u_grad rel-L2 | u_grad cos | v_grad rel-L2 | v_grad cos |")
print("|---|---|---|---|---|---|---|")
FWD, WG, XG = "sri,soi->sro", "sro,sri->soi", "sro,soi->sri"
extra = [({"default": "tf32x3", WG: "bf16x3"}, "bf16x3", None), ({"default": "tf32x3", XG: "bf16x3"}, "bf16x3", None),
         ({"default": "tf32x3", FWD: "bf16x3"}, "bf16x3", None), ({"default": "bf16x3", FWD: "tf32x3"}, "bf16x3", None),
         ("tf32x3", {"default": "bf16x3", WG: "bf16"}, None), ("tf32x3", {"default": "bf16x3", XG: "bf16"}, None),
         ("tf32x3", {"default": "bf16x3", WG: "bf16", XG: "bf16"}, None), ("tf32x3", {"default": "bf16x3", FWD: "bf16"}, None)]
for gm, hm, om in [("fp32", "fp32", None), ("tf32x3", "tf32x3", None), ("bf16x3", "bf16x3", None), ("tf32x3", "bf16x3", None),
                   ("bf16x3", "tf32x3", None), ("tf32x3", "bf16x3", "bf16x3"), ("tf32x3", "bf16", None)] + extra:
    r = run(gm, hm, om)
    om = om or (gm if isinstance(gm, str) else gm["default"])
    if not isinstance(gm, str):
        gm = gm["default"] + ", " + ", ".join(f"{v} for " + {FWD: "fwd", WG: "wgrad", XG: "xgrad"}[k] for k, v in gm.items() if k != "default")
    hm = hm if isinstance(hm, str) else "bf16x3, bf16 for " + "+".join({FWD: "fwd", WG: "wgrad", XG: "xgrad"}[k] for k in hm if k != "default")
    print(f"| {gm} | {hm} | {om} | {rel(r['u_grad'], ex['u_grad']):.2e} | {cos(r['u_grad'], ex['u_grad']):.6f} | "
          f"{rel(r['v_grad'], ex['v_grad']):.2e} | {cos(r['v_grad'], ex['v_grad']):.6f} |", flush=True)
