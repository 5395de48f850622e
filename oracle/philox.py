"""numpy restatement of the Philox4x32-10 + Box-Muller generator of csrc/psvi_common.cuh (TEST INFRASTRUCTURE ONLY).

Counter (idx//4, sample, slab, domain), key = 64-bit seed; the 4 outputs give normals for TL indices 4*(idx//4)..+3.
Philox4x32-10 itself is the published algorithm of Salmon et al., "Parallel random numbers: as easy as 1, 2, 3" (SC'11)."""
import numpy as np

M0, M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
W0, W1 = np.uint32(0x9E3779B9), np.uint32(0xBB67AE85)


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    c0, c1, c2, c3 = (np.asarray(c, dtype=np.uint32) for c in (c0, c1, c2, c3))
    k0, k1 = np.uint32(k0), np.uint32(k1)
    with np.errstate(over="ignore"):
        for _ in range(10):
            p0 = M0 * c0.astype(np.uint64)
            p1 = M1 * c2.astype(np.uint64)
            hi0, lo0 = (p0 >> np.uint64(32)).astype(np.uint32), p0.astype(np.uint32)
            hi1, lo1 = (p1 >> np.uint64(32)).astype(np.uint32), p1.astype(np.uint32)
            c0, c1, c2, c3 = hi1 ^ c1 ^ k0, lo1, hi0 ^ c3 ^ k1, lo0
            k0, k1 = np.uint32(k0 + W0), np.uint32(k1 + W1)
    return c0, c1, c2, c3


def _bm(a, b):
    f = np.float32
    u1 = a.astype(f) * f(2.3283064365386963e-10) + f(1.1641532182693481e-10)
    u2 = b.astype(f) * f(2.3283064365386963e-10) + f(1.1641532182693481e-10)
    r = np.sqrt(f(-2.0) * np.log(u1)).astype(f)
    ang = (f(6.283185307179586) * u2).astype(f)
    return (r * np.cos(ang)).astype(f), (r * np.sin(ang)).astype(f)


def philox_normal_np(seed, domain, first_slab, n_slabs, S, P):
    n4 = (P + 3) // 4
    sl, s, q4 = np.meshgrid(np.arange(n_slabs) + first_slab, np.arange(S), np.arange(n4), indexing="ij")
    shp = sl.shape
    x, y, z, w = philox4x32_10(q4.ravel(), s.ravel(), sl.ravel(), np.full(q4.size, domain, dtype=np.uint32),
                               seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
    n0, n1 = _bm(x, y)
    n2, n3 = _bm(z, w)
    out = np.stack([n0, n1, n2, n3], -1).reshape(shp + (4,)).reshape(n_slabs, S, n4 * 4)
    return out[:, :, :P]
