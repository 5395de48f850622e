"""Import shim for the UNMODIFIED reference (TEST INFRASTRUCTURE ONLY; never imported by the product).

The reference is pure Python.  Two third-party modules it imports at module scope are absent in
this image and unused by the hot path: `faiss` (psvi/inference/utils.py:28) and `arff`
(psvi/experiments/experiments_utils.py:14); they are stubbed.  The reference tree is looked up at
/root/reference (build container) or baseline/_ref (a pip --target install that travels to the GPU box).
"""
from __future__ import annotations

import os
import sys
import types

_HERE = os.path.dirname(os.path.abspath(__file__))
# the pip --target install (python -m pip install --no-index --no-build-isolation --no-deps --target baseline/_ref
# /root/reference; __graft_entry__.build() does it) comes first: it is what travels to the GPU box
_CANDIDATES = [os.path.join(os.path.dirname(_HERE), "baseline", "_ref"), "/root/reference"]


def reference_root():
    for c in _CANDIDATES:
        if os.path.isdir(os.path.join(c, "psvi", "inference")):
            return c
    return None


def import_reference():
    """Returns the reference `psvi` package (raises ImportError when it is not reachable)."""
    root = reference_root()
    if root is None:
        raise ImportError("reference tree not found (looked in %s)" % _CANDIDATES)
    for m in ("faiss", "arff"):
        if m not in sys.modules:
            try:
                __import__(m)
            except Exception:
                sys.modules[m] = types.ModuleType(m)
    # make sure OUR drop-in `psvi` package is not shadowing the reference one
    for k in [k for k in sys.modules if k == "psvi" or k.startswith("psvi.")]:
        mod = sys.modules[k]
        f = getattr(mod, "__file__", "") or ""
        if not f.startswith(root):
            del sys.modules[k]
    if root not in sys.path:
        sys.path.insert(0, root)
    else:
        sys.path.remove(root)
        sys.path.insert(0, root)
    import psvi  # noqa: F401
    import psvi.inference.psvi_classes  # noqa: F401
    assert sys.modules["psvi"].__file__.startswith(root), sys.modules["psvi"].__file__
    return sys.modules["psvi"]


class NoiseFeeder:
    """Replaces torch.distributions.normal._standard_normal so that the reference consumes a numpy-seeded
    noise stream in its own draw order (per VI layer: weight draw [S,out,in], then bias draw [S,1,out];
    neural_net.py:155-170).  One [S, P_theta] array ("TL" layout, see psvi_oracle.py) is generated per forward."""

    def __init__(self, dims, S, seed, fullcov=False):
        import numpy as np
        self.dims, self.S = list(dims), S
        self.rng = np.random.default_rng(seed)
        self.history = []
        self._pos = 0
        self._shapes = []
        for l in range(1, len(dims)):
            if fullcov:   # MultivariateNormal.rsample: ONE [S, n] draw per layer (neural_net.py:467-472)
                self._shapes += [(S, 1, dims[l] * (dims[l - 1] + 1))]
            else:
                self._shapes += [(S, dims[l], dims[l - 1]), (S, 1, dims[l])]
        self._fullcov = fullcov
        self._call = 0

    @staticmethod
    def stream(dims, S, seed, n):
        """The same arrays a feeder with this seed hands out for its first n forwards."""
        import numpy as np
        rng = np.random.default_rng(seed)
        P = sum(dims[l] * (dims[l - 1] + 1) for l in range(1, len(dims)))
        return [rng.standard_normal((S, P)).astype(np.float32) for _ in range(n)]

    def __call__(self, shape, dtype, device):
        import numpy as np
        import torch
        if self._call == 0:
            P = sum(s[1] * s[2] for s in self._shapes)
            self.history.append(self.rng.standard_normal((self.S, P)).astype(np.float32))
            self._pos = 0
        exp = self._shapes[self._call]
        if self.S == 1 and not self._fullcov and tuple(shape) != exp:
            # mc_samples == 1: the reference draws without the sample dimension (neural_net.py:164-170)
            assert tuple(shape) == (exp[1], exp[2]) or (exp[1] == 1 and tuple(shape) == (exp[2],)), (tuple(shape), exp)
            n = exp[1] * exp[2]
            out = self.history[-1][0, self._pos:self._pos + n].reshape(tuple(shape))
            self._pos += n
            self._call = (self._call + 1) % len(self._shapes)
            return torch.from_numpy(np.ascontiguousarray(out)).to(dtype=dtype, device=device)
        if self._fullcov:
            assert tuple(shape) == (exp[0], exp[2]), (tuple(shape), exp)
            shape_out = (exp[0], exp[2])
        else:
            assert tuple(shape) == exp, (tuple(shape), exp)
            shape_out = exp
        n = exp[1] * exp[2]
        out = self.history[-1][:, self._pos:self._pos + n].reshape(shape_out)
        self._pos += n
        self._call = (self._call + 1) % len(self._shapes)
        return torch.from_numpy(np.ascontiguousarray(out)).to(dtype=dtype, device=device)

    def __enter__(self):
        import torch.distributions.multivariate_normal as tdm
        import torch.distributions.normal as tdn
        self._orig = (tdn._standard_normal, tdm._standard_normal)
        tdn._standard_normal = self
        tdm._standard_normal = self
        return self

    def __exit__(self, *a):
        import torch.distributions.multivariate_normal as tdm
        import torch.distributions.normal as tdn
        tdn._standard_normal, tdm._standard_normal = self._orig


class LeNetNoiseFeeder(NoiseFeeder):
    """NoiseFeeder for make_lenet (neural_net.py:334-359): per layer weight draw [S, *weight.shape] then bias draw
    [S, 1, n_bias]; the last VILinear has mc_samples = 1 (Q4) and draws [10, 84], [10] ONCE -- its block of the [S, P] slab is
    sample 0's row replicated, which is how the CUDA path and the oracle consume a shared draw."""

    SHAPES = [((6, 1, 5, 5), False), ((6,), False), ((16, 6, 5, 5), False), ((16,), False), ((120, 400), False), ((120,), False),
              ((84, 120), False), ((84,), False), ((10, 84), True), ((10,), True)]

    def __init__(self, S, seed):
        import numpy as np
        self.S = S
        self.rng = np.random.default_rng(seed)
        self.history = []
        self._pos = 0
        self._call = 0
        self.P = sum(int(np.prod(s)) for s, _ in self.SHAPES)

    @staticmethod
    def _share(e):
        e[:, -850:] = e[:1, -850:]
        return e

    @staticmethod
    def stream(S, seed, n):
        import numpy as np
        rng = np.random.default_rng(seed)
        P = sum(int(np.prod(s)) for s, _ in LeNetNoiseFeeder.SHAPES)
        return [LeNetNoiseFeeder._share(rng.standard_normal((S, P)).astype(np.float32)) for _ in range(n)]

    def __call__(self, shape, dtype, device):
        import numpy as np
        import torch
        if self._call == 0:
            self.history.append(self._share(self.rng.standard_normal((self.S, self.P)).astype(np.float32)))
            self._pos = 0
        base, shared = self.SHAPES[self._call]
        n = int(np.prod(base))
        blk = self.history[-1][:, self._pos:self._pos + n]
        if shared:
            exp = tuple(base)
            out = blk[0].reshape(exp)
        else:
            exp = (self.S,) + tuple(base) if len(base) > 1 else (self.S, 1) + tuple(base)
            out = blk.reshape(exp)
        assert tuple(shape) == exp, (tuple(shape), exp)
        self._pos += n
        self._call = (self._call + 1) % len(self.SHAPES)
        return torch.from_numpy(np.ascontiguousarray(out)).to(dtype=dtype, device=device)
