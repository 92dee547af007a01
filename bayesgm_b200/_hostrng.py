"""NumPy's legacy global generator, continued natively (csrc/host_rng.cu, include/bgm_b200.h).

`LegacyStream.from_numpy()` takes over the state of `np.random` (the generator the reference's
training loops draw from), produces `np.random.choice(n, bs, replace=False)` / `np.random.normal` /
the whole per-iteration draw pattern of `egm_init` bit-exactly in C -- callable from a background
thread, ctypes releases the GIL -- and `to_numpy()` hands the advanced state back so that whatever
draws from `np.random` next continues the stream exactly as in the reference.
"""
import ctypes as C
import queue
import threading

import numpy as np

from . import _lib


class LegacyStream(object):
    def __init__(self, state):
        name, key, pos, has_gauss, gauss = state
        assert name == 'MT19937'
        self.st = _lib.MtState()
        C.memmove(self.st.key, np.ascontiguousarray(key, np.uint32).ctypes.data, 624 * 4)
        self.st.pos, self.st.has_gauss, self.st.gauss = int(pos), int(has_gauss), float(gauss)
        self._work = None

    @classmethod
    def from_numpy(cls):
        return cls(np.random.get_state())

    def state(self):
        key = np.frombuffer(bytes(self.st.key), dtype=np.uint32).copy()
        return ('MT19937', key, int(self.st.pos), int(self.st.has_gauss), float(self.st.gauss))

    def to_numpy(self):
        np.random.set_state(self.state())

    def _scratch(self, n):
        if self._work is None or self._work.size < n:
            self._work = np.empty(n, np.int32)
        return self._work

    def choice(self, n, size):
        out = np.empty(size, np.int32)
        _lib.call("bgm_host_choice", C.byref(self.st), int(n), int(size), out.ctypes.data_as(C.c_void_p),
                  self._scratch(n).ctypes.data_as(C.c_void_p))
        return out

    def normal(self, loc, scale, shape):
        out = np.empty(shape, np.float32)
        _lib.call("bgm_host_normal", C.byref(self.st), float(loc), float(scale), int(out.size),
                  out.ctypes.data_as(C.c_void_p))
        return out

    def rand(self, count):
        out = np.empty(count, np.float64)
        _lib.call("bgm_host_rand", C.byref(self.st), int(count), out.ctypes.data_as(C.c_void_p))
        return out

    def egm_chunk(self, n, bs, zd, freq, iters):
        idx = np.empty((iters, freq + 1, bs), np.int32)
        z = np.empty((iters, freq + 1, bs, zd), np.float32)
        _lib.call("bgm_host_egm_stream", C.byref(self.st), int(n), int(bs), int(zd), int(freq), int(iters),
                  idx.ctypes.data_as(C.c_void_p), z.ctypes.data_as(C.c_void_p),
                  self._scratch(n).ctypes.data_as(C.c_void_p))
        return idx, z


class EgmProducer(object):
    """Background thread that runs `total` iterations of egm_init's draw pattern `chunk` iterations at a
    time, `depth` chunks ahead of the consumer.  `close()` joins the thread and writes the final state
    back into `np.random` (the stream position after exactly `total` iterations, as in the reference)."""

    def __init__(self, n, bs, zd, freq, total, chunk=64, depth=4):
        self.stream = LegacyStream.from_numpy()
        self.q = queue.Queue(maxsize=depth)
        self._err = None

        def work():
            try:
                done = 0
                while done < total:
                    cnt = min(chunk, total - done)
                    self.q.put(self.stream.egm_chunk(n, bs, zd, freq, cnt))
                    done += cnt
            except BaseException as e:      # surfaced in get()
                self._err = e
                self.q.put(None)
        self.thread = threading.Thread(target=work, daemon=True)
        self.thread.start()

    def get(self):
        item = self.q.get()
        if item is None:
            raise self._err
        return item

    def close(self, drain=False):
        if drain:
            while self.thread.is_alive() or not self.q.empty():
                try:
                    self.q.get(timeout=0.05)
                except queue.Empty:
                    pass
        self.thread.join()
        self.stream.to_numpy()


class FloydProducer(object):
    """Same hand-over interface as EgmProducer for `egm_init(index_stream='floyd')`: mini-batch indices
    drawn WITHOUT permuting all n rows (NumPy's Generator.choice: Floyd's subset sampling + a shuffle of
    the subset) and the prior draws from the same PCG64 generator, seeded by one draw from `np.random`
    (so `np.random.seed` still fixes the run).  The same distribution as the reference's
    `np.random.choice(n, bs, replace=False)` (bgm/base.py:406) at O(bs) instead of O(n) per draw -- not
    the same stream: the bit-exact continuation of NumPy's legacy generator is EgmProducer."""

    def __init__(self, n, bs, zd, freq, total, chunk=64, depth=4):
        self.rng = np.random.Generator(np.random.PCG64(int(np.random.randint(0, 2 ** 31 - 1))))
        self.q = queue.Queue(maxsize=depth)
        self._err = None
        self._stop = False

        def work():
            try:
                done = 0
                while done < total and not self._stop:
                    cnt = min(chunk, total - done)
                    idx = np.empty((cnt, freq + 1, bs), np.int32)
                    for c in range(cnt):
                        for k in range(freq + 1):
                            idx[c, k] = self.rng.choice(n, bs, replace=False)
                    z = self.rng.standard_normal((cnt, freq + 1, bs, zd), dtype=np.float32)
                    self.q.put((idx, z))
                    done += cnt
            except BaseException as e:
                self._err = e
                self.q.put(None)
        self.thread = threading.Thread(target=work, daemon=True)
        self.thread.start()

    def get(self):
        item = self.q.get()
        if item is None:
            raise self._err
        return item

    def close(self, drain=False):
        self._stop = True
        while self.thread.is_alive() or not self.q.empty():
            try:
                self.q.get(timeout=0.05)
            except queue.Empty:
                pass
        self.thread.join()
