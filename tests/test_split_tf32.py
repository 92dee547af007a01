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
    """Veltkamp's split (umma::split_tf32_x2 until round 2): c = v * (2^13 + 1); hi = c - (c - v); lo = v - hi
    (each op rounded to fp32)."""
    c = (v * f32(8193.0)).astype(f32)
    t = (c - v).astype(f32)
    hi = (c - t).astype(f32)
    return hi, (v - hi).astype(f32)


def fma32(a, b, c):
    """a * b + c with ONE rounding to fp32 (the float64 product of two fp32 values is exact; the float64 sum rounds
    far below the fp32 ulp for the magnitudes used here)."""
    return (a.astype(np.float64) * b.astype(np.float64) + c.astype(np.float64)).astype(f32)


def split_fma3(v):
    """umma::split_tf32_x2 (three FFMA2): s = rn(v + 8192 v); hi = s - 8192 v (exact); lo = v - hi (exact)."""
    k = np.full_like(v, 8192.0)
    s = fma32(v, k, v)
    hi = fma32(v, -k, s)
    return hi, fma32(hi, np.full_like(v, -1.0), v)


def tf32_trunc(a):
    return (a.view(np.uint32) & np.uint32(0xFFFFE000)).view(np.float32)


def values():
    rs = np.random.RandomState(0)
    v = np.concatenate([rs.standard_normal(20000), rs.standard_normal(2000) * 1e-6, rs.standard_normal(2000) * 1e6,
                        [0.0, 1.0, -1.0, 0.1, 3.0, 1e-30, 65504.0, 1.0 + 2.0 ** -11, 1.0 + 2.0 ** -12]]).astype(f32)
    return v


def test_splits_are_exact_and_tf32_representable():
    v = values()
    for split in (split_bits, split_veltkamp, split_fma3):
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
    for split in (split_bits, split_veltkamp, split_fma3):
        ah, al = split(a)
        wh, wl = split(w)
        al, wl = tf32_trunc(al), tf32_trunc(wl)        # the tensor core reads tf32: low bits of lo are dropped
        got = (ah.astype(np.float64) @ wh.astype(np.float64) + al.astype(np.float64) @ wh.astype(np.float64) +
               ah.astype(np.float64) @ wl.astype(np.float64))
        one = tf32_trunc(a).astype(np.float64) @ tf32_trunc(w).astype(np.float64)
        scale = np.abs(want).max()
        assert np.abs(got - want).max() / scale < 2e-6     # three passes: fp32 level
        assert np.abs(one - want).max() / scale > 1e-4     # a single TF32 pass is not


def test_three_instruction_split_rounds_to_nearest_like_the_integer_form():
    """Over 40 decades: hi of the three-FFMA2 split is tf32-exact, hi + lo == v exactly, |lo| <= 2^-11 |v|, and hi equals
    the integer round-half-up form except at exact ties (round-to-even there) and for mantissas within 2.4e-4 of 2,
    where 8193 v crosses a binade and the rounding step doubles (hi is then one of the two neighbours, lo still exact)."""
    rs = np.random.RandomState(2)
    v = (rs.standard_normal(400000) * 10.0 ** rs.randint(-20, 20, 400000)).astype(f32)
    hi3, lo3 = split_fma3(v)
    hib, _ = split_bits(v)
    assert np.array_equal(tf32_trunc(hi3), hi3)
    assert np.array_equal(hi3.astype(np.float64) + lo3.astype(np.float64), v.astype(np.float64))
    assert np.all(np.abs(lo3) <= np.abs(v) * 2.0 ** -11 * 1.001)
    diff = hi3 != hib
    assert diff.mean() < 1e-3
    ties = (v.view(np.uint32) & np.uint32(0x1FFF)) == np.uint32(0x1000)
    top = (v.view(np.uint32) & np.uint32(0x7FFFFF)) >= np.uint32(int((16384.0 / 8193.0 - 1.0) * 2 ** 23) - 1)
    assert np.all(ties[diff] | top[diff])
