"""Helpers shared by the GPU parity tests (call the CUDA path through the C ABI binding)."""
import glob
import os

import numpy as np
import torch

from oracle import psvi_oracle as po
from oracle.ref_import import NoiseFeeder

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CASES = sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN, "*.npz"))
               if not os.path.basename(p).startswith(("mfvi", "meanfieldvi", "regressor", "regbase", "sparsebbvi", "fn2", "hyper", "lenet", "ablated", "noiw", "grid", "variant", "fixedpoint", "joint", "alternating", "learnz")))


def rel_l2(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300)


def load(name):
    g = dict(np.load(os.path.join(GOLDEN, name + ".npz")))
    dims = [int(d) for d in g["dims"]]
    S, T = int(g["S"]), int(g["T"])
    eps = NoiseFeeder.stream(dims, S, int(g["noise_seed"]), int(g["n_forwards"]))
    return g, dims, S, T, eps


def dev(x, dtype=torch.float32):
    return torch.as_tensor(np.ascontiguousarray(x)).to(device="cuda", dtype=dtype).contiguous()


def zeros(*shape):
    return torch.zeros(*shape, device="cuda", dtype=torch.float32)
