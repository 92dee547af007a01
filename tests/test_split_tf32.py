"""CPU check of the fp32 -> (hi, lo) splits the tensor-core kernels are built on (csrc/umma.cuh):
`hi` must be exactly representable in tf32 (11 significant bits), `hi + lo` must reproduce the value
exactly, and a product evaluated as hi*hi + lo*hi + hi*lo must be fp32-accurate.  NumPy float32
restatement of the device arithmetic (round-to-nearest, one rounding per operation)."""
import numpy as np

f32 = np.float32


def split_bits(v):
    """umma::split_tf32: integer round-half-up at bit 13, lo = v - hi."""
    b = v.view(np.uint32)
    hi = ((b + np.uint32(0x1000)) & np.uint32(0xFFFFE000)).view(np.float32)
    return hi, (v - hi).astype(f32)


def split_veltkamp(v):
    """umma::split_tf32_x2: c = v * (2^13 + 1); hi = c - (c - v); lo = v - hi (each op rounded to fp32)."""
    c = (v * f32(8193.0)).astype(f32)
    t = (c - v).astype(f32)
    hi = (c - t).astype(f32)
    return hi, (v - hi).astype(f32)


def tf32_trunc(a):
    return (a.view(np.uint32) & np.uint32(0xFFFFE000)).view(np.float32)


def values():
    rs = np.random.RandomState(0)
    v = np.concatenate([rs.standard_normal(20000), rs.standard_normal(2000) * 1e-6, rs.standard_normal(2000) * 1e6,
                        [0.0, 1.0, -1.0, 0.1, 3.0, 1e-30, 65504.0, 1.0 + 2.0 ** -11, 1.0 + 2.0 ** -12]]).astype(f32)
    return v


def test_splits_are_exact_and_tf32_representable():
    v = values()
    for split in (split_bits, split_veltkamp):
        hi, lo = split(v)
        assert np.array_equal((hi.astype(np.float64) + lo.astype(np.float64)).astype(f32), v)   # hi + lo == v
        assert np.array_equal(tf32_trunc(hi), hi)                         # 13 low mantissa bits of hi are zero
        nz = v != 0
        assert np.all(np.abs(lo[nz]) <= np.abs(v[nz]) * 2.0 ** -11 * 1.0001)   # lo is the rounding remainder


def test_three_pass_product_is_fp32_accurate():
    rs = np.random.RandomState(1)
    a = rs.standard_normal((256, 64)).astype(f32)
    w = (rs.standard_normal((64, 64)) * 0.3).astype(f32)
    want = a.astype(np.float64) @ w.astype(np.float64)
    for split in (split_bits, split_veltkamp):
        ah, al = split(a)
        wh, wl = split(w)
        al, wl = tf32_trunc(al), tf32_trunc(wl)        # the tensor core reads tf32: low bits of lo are dropped
        got = (ah.astype(np.float64) @ wh.astype(np.float64) + al.astype(np.float64) @ wh.astype(np.float64) +
               ah.astype(np.float64) @ wl.astype(np.float64))
        one = tf32_trunc(a).astype(np.float64) @ tf32_trunc(w).astype(np.float64)
        scale = np.abs(want).max()
        assert np.abs(got - want).max() / scale < 2e-6     # three passes: fp32 level
        assert np.abs(one - want).max() / scale > 1e-4     # a single TF32 pass is not
