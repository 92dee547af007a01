"""csrc/host_rng.cu continues NumPy's legacy global generator bit-exactly (CPU-only test)."""
import numpy as np

from bayesgm_b200._hostrng import LegacyStream, EgmProducer


def test_choice_normal_rand_match_numpy_bit_for_bit():
    for seed in (1024, 7, 123456):
        np.random.seed(seed)
        np.random.normal(size=3)                      # leaves a cached gaussian in the state
        s = LegacyStream.from_numpy()
        for n, bs in ((1000, 32), (20000, 32), (33, 33), (5, 1), (100000, 64)):
            want = np.random.choice(n, bs, replace=False)
            np.testing.assert_array_equal(s.choice(n, bs), want)
            wz = np.random.normal(np.zeros(7), 1.0, (bs, 7)).astype('float32')
            np.testing.assert_array_equal(s.normal(0.0, 1.0, (bs, 7)), wz)
            np.testing.assert_array_equal(s.rand(5), np.random.rand(5))
        # handing the state back: both generators continue identically
        a = np.random.choice(50, 50, replace=False)
        s.to_numpy()
        np.testing.assert_array_equal(np.random.choice(50, 50, replace=False), a)


def test_egm_stream_matches_reference_call_order():
    n, bs, zd, freq, iters = 5000, 32, 10, 5, 7
    np.random.seed(1024)
    want_idx = np.empty((iters, freq + 1, bs), np.int32)
    want_z = np.empty((iters, freq + 1, bs, zd), np.float32)
    for c in range(iters):                            # causalbgm/base.py:405-413
        for k in range(freq):
            want_idx[c, k] = np.random.choice(n, bs, replace=False)
            want_z[c, k] = np.random.normal(np.zeros(zd), 1.0, (bs, zd)).astype('float32')
        want_z[c, freq] = np.random.normal(np.zeros(zd), 1.0, (bs, zd)).astype('float32')
        want_idx[c, freq] = np.random.choice(n, bs, replace=False)
    after = np.random.choice(n, n, replace=False)    # what fit() draws next (:489)
    np.random.seed(1024)
    prod = EgmProducer(n, bs, zd, freq, iters, chunk=3, depth=2)
    got_idx, got_z = [], []
    done = 0
    while done < iters:
        i, z = prod.get()
        got_idx.append(i)
        got_z.append(z)
        done += len(i)
    prod.close()
    np.testing.assert_array_equal(np.concatenate(got_idx), want_idx)
    np.testing.assert_array_equal(np.concatenate(got_z), want_z)
    np.testing.assert_array_equal(np.random.choice(n, n, replace=False), after)


def test_free_memory_estimate_asks_the_driver_once():
    """_lib.free_memory_estimate (the memory-budget checks of predict()): cudaMemGetInfo once per device, afterwards
    capacity - live tensor bytes from the allocator's counters."""
    from bayesgm_b200 import _lib

    class FakeCuda(object):
        def __init__(self):
            self.calls, self.allocated, self.reserved, self.device = 0, 0, 0, 0

        def current_device(self):
            return self.device

        def mem_get_info(self):
            self.calls += 1
            return (100 - self.reserved, 128)

        def memory_reserved(self):
            return self.reserved

        def memory_allocated(self):
            return self.allocated

    class FakeTorch(object):
        cuda = FakeCuda()

    t = FakeTorch()
    saved = dict(_lib._MEM_CAP)
    _lib._MEM_CAP.clear()
    try:
        t.cuda.reserved, t.cuda.allocated = 30, 20
        assert _lib.free_memory_estimate(t) == 100 - 20          # driver-free 70 + cached 30 = capacity 100
        t.cuda.allocated = 90
        assert _lib.free_memory_estimate(t) == 10
        t.cuda.allocated = 500
        assert _lib.free_memory_estimate(t) == 0
        assert t.cuda.calls == 1
        t.cuda.device = 1                                          # another device: asked once more
        t.cuda.allocated = 0
        assert _lib.free_memory_estimate(t) == 100 and t.cuda.calls == 2
    finally:
        _lib._MEM_CAP.clear()
        _lib._MEM_CAP.update(saved)
