"""NumPy restatement of the library's Philox4x32-10 noise streams (csrc/common.cuh) -- TEST
INFRASTRUCTURE (see oracle/__init__.py).  Philox4x32-10 is Salmon et al., SC'11; the known-answer
vectors of Random123 are checked in tests/test_oracle.py."""
import numpy as np

M0, M1, W0, W1 = 0xD2511F53, 0xCD9E8D57, 0x9E3779B9, 0xBB67AE85


def philox4x32_10(c, k):
    """c: (...,4) uint32 counters, k: (2,) key -> (...,4) uint32."""
    c = [c[..., i].astype(np.uint64) for i in range(4)]
    k0, k1 = np.uint64(k[0]), np.uint64(k[1])
    mask = np.uint64(0xFFFFFFFF)
    for _ in range(10):
        p0 = np.uint64(M0) * c[0]
        p1 = np.uint64(M1) * c[2]
        hi0, lo0 = p0 >> np.uint64(32), p0 & mask
        hi1, lo1 = p1 >> np.uint64(32), p1 & mask
        c = [hi1 ^ c[1] ^ k0, lo1, hi0 ^ c[3] ^ k1, lo0]
        k0 = (k0 + np.uint64(W0)) & mask
        k1 = (k1 + np.uint64(W1)) & mask
    return np.stack(c, axis=-1).astype(np.uint32)


def philox_normal4(seed, rows, t, kind, j):
    """NumPy restatement of csrc/common.cuh `normal4`: 4 unit normals per (row, t, kind, j)
    from one Philox block via Box-Muller.  rows: int64 array -> (len(rows), 4) float32."""
    rows = np.asarray(rows, np.int64)
    c = np.stack([np.full(rows.shape, t & 0xFFFFFFFF, np.uint64), (rows & 0xFFFFFFFF).astype(np.uint64),
                  ((rows >> 32) & 0xFFFFFFFF).astype(np.uint64),
                  np.full(rows.shape, ((kind << 24) | j) & 0xFFFFFFFF, np.uint64)], axis=-1).astype(np.uint32)
    r = philox4x32_10(c, np.array([seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF], np.uint64))

    def bm(a, b):
        u0 = (a.astype(np.float64) * 2.3283064365386963e-10 + 1.1641532182693481e-10).astype(np.float32)
        u1 = ((b >> np.uint32(8)).astype(np.float32) * np.float32(5.9604644775390625e-08))
        rad = np.sqrt(np.float32(-2.0) * np.log(u0)).astype(np.float32)
        ang = (2.0 * np.pi * u1.astype(np.float64))
        return (rad * np.cos(ang)).astype(np.float32), (rad * np.sin(ang)).astype(np.float32)
    n0, n1 = bm(r[..., 0], r[..., 1])
    n2, n3 = bm(r[..., 2], r[..., 3])
    return np.stack([n0, n1, n2, n3], axis=-1)


