"""Verdict item 8 ("lenet on tensor cores, or measure why not"): what would tensor-core operand formats do to the lenet
hypergradient?  CPU emulation on the pinned lenet oracle (oracle/lenet_oracle.py + psvi_oracle_generic.nested_step, fp64):
every layer product (np.einsum with two operands) of the selected layers has its OPERANDS rounded to the tensor format
(accumulation stays exact, as on the tensor pipe with fp32 accumulators it is ~1e-7):
  fp32      operands rounded to fp32                     (what the CUDA-core kernels of psvi_lenet.cu do)
  tf32x3    (hi, lo) TF32 pairs, hi.hi + hi.lo + lo.hi   (3 kind::tf32 MMAs)
  bf16x3    (hi, lo) BF16 pairs, three products          (3 kind::f16 MMAs)
  bf16      one BF16 product
for the fc layers only, and for fc + conv layers.  Output: rel-L2 and cosine of the nested-step hypergradients on u and v
against the exact fp64 run (tests/golden/lenet_m10.npz case: S, M, T of the golden)."""
import os, sys, time
sys.path.insert(0, os.getcwd())
import numpy as np
from oracle import lenet_oracle as lo, psvi_oracle as po, psvi_oracle_generic as pg
from oracle.ref_import import LeNetNoiseFeeder

g = dict(np.load("tests/golden/lenet_m10.npz"))
S, T = int(g["S"]), int(g["T"])
eps = [e.astype(np.float64) for e in LeNetNoiseFeeder.stream(S, int(g["noise_seed"]), int(g["n_forwards"]))]
N = float(g["N"])
fam = lo.LeNetMeanField()
if len(sys.argv) > 1:            # synthetic larger case: S T M B  (weights of the golden, images resampled from its rows + noise)
    S, T, M, B = (int(a) for a in sys.argv[1:5])
    rng = np.random.default_rng(0)
    pool = np.concatenate([g["u0"].reshape(len(g["u0"]), -1), g["xb"].reshape(len(g["xb"]), -1)])
    g["u0"] = pool[rng.integers(0, len(pool), M)] + 0.3 * rng.standard_normal((M, pool.shape[1]))
    g["z"] = rng.integers(0, 10, M).astype(np.float64)
    g["v0"] = 0.3 * rng.standard_normal(M)
    g["xb"] = pool[rng.integers(0, len(pool), B)] + 0.3 * rng.standard_normal((B, pool.shape[1]))
    g["yb"] = rng.integers(0, 10, B).astype(np.float64)
    eps = [lo.share_last_layer(rng.standard_normal((S, lo.P))) for _ in range(T + 3)]

def rnd_bits(x, keep):           # round fp32 mantissa to `keep` explicit bits (nearest even)
    x32 = np.ascontiguousarray(x, dtype=np.float32)
    u = x32.view(np.uint32).astype(np.uint64)
    drop = 23 - keep
    u = (u + ((1 << (drop - 1)) - 1) + ((u >> drop) & 1)) >> drop << drop
    return u.astype(np.uint32).view(np.float32).astype(np.float64)

def split(x, keep):
    hi = rnd_bits(x, keep)
    lo_ = rnd_bits(x - hi, keep)
    return hi, lo_

FC = {"sri,soi->sro", "sro,sri->soi", "sro,soi->sri"}
real_einsum = np.einsum
def make(mode, layers):
    def ein(expr, *ops, **kw):
        if len(ops) != 2 or mode == "exact":
            return real_einsum(expr, *ops, **kw)
        is_fc = expr in FC
        if (layers == "fc" and not is_fc):
            return real_einsum(expr, *ops, **kw)
        A, B = ops
        if mode == "fp32":
            return real_einsum(expr, A.astype(np.float32).astype(np.float64), B.astype(np.float32).astype(np.float64), **kw)
        keep = {"bf16": 7, "bf16x3": 7, "tf32x3": 10}[mode]
        if mode == "bf16":
            return real_einsum(expr, rnd_bits(A, keep), rnd_bits(B, keep), **kw)
        Ah, Al = split(A, keep); Bh, Bl = split(B, keep)
        return real_einsum(expr, Ah, Bh, **kw) + real_einsum(expr, Ah, Bl, **kw) + real_einsum(expr, Al, Bh, **kw)
    return ein

def run(mode, layers):
    np.einsum = make(mode, layers)
    try:
        r = pg.nested_step(fam, g["phi0"], eps[2:2 + T], eps[2 + T], g["u0"], g["z"], g["v0"], g["xb"], g["yb"], N,
                           float(g["lr0net"]), vmode=1)
    finally:
        np.einsum = real_einsum
    return r

def rel(a, b): return np.linalg.norm(a - b) / np.linalg.norm(b)
def cos(a, b): return float(np.dot(a.ravel(), b.ravel()) / np.linalg.norm(a) / np.linalg.norm(b))
t0 = time.time()
ex = run("exact", "all")
print(f"case: S={S} T={T} M={g['u0'].shape[0]} B={g['xb'].shape[0]}  exact loss {ex['loss']:.6f}  ({time.time()-t0:.1f}s per run)")
print("| operands | layers | loss rel | u_grad rel-L2 | u_grad cos | v_grad rel-L2 | v_grad cos |")
print("|---|---|---|---|---|---|---|")
for layers in ("fc", "all"):
    for mode in ("fp32", "tf32x3", "bf16x3", "bf16"):
        r = run(mode, layers)
        print(f"| {mode} | {layers} | {abs(r['loss']-ex['loss'])/abs(ex['loss']):.2e} | {rel(r['u_grad'], ex['u_grad']):.2e} | "
              f"{cos(r['u_grad'], ex['u_grad']):.6f} | {rel(r['v_grad'], ex['v_grad']):.2e} | {cos(r['v_grad'], ex['v_grad']):.6f} |",
              flush=True)
