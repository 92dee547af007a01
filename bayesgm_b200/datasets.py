"""Synthetic-input generators with the reference's exact NumPy legacy-RNG call order.

Only what the benchmark / tests need to build inputs on a box where /root/reference
is absent.  Written from the formulas in `src/bayesgm/datasets/causal_samplers.py:40-67`,
`base_sampler.py:29-46` and `prior_samplers.py:20-59`; tests/test_golden.py checks bit-equality
against outputs of the reference modules themselves (tests/golden/reference_datasets.npz, made by
tests/golden/make_golden.py where /root/reference is importable).
"""
import math

import numpy as np


def _standardize(v):
    """sklearn StandardScaler().fit_transform: (v - mean) / population std."""
    v = np.asarray(v, dtype=np.float32)
    mean = v.mean(axis=0, dtype=np.float64)
    var = v.astype(np.float64).var(axis=0)
    scale = np.sqrt(var)
    scale[scale == 0.0] = 1.0
    return ((v - mean) / scale).astype(np.float32)


class Base_sampler(object):
    """base_sampler.py:6-84 -- stores (x, y, v) as float32 and cycles shuffled
    mini-batches; `normalize` standardises v."""

    def __init__(self, x, y, v, batch_size=32, normalize=False, random_seed=123):
        assert len(x) == len(y) == len(v)
        np.random.seed(random_seed)
        self.data_x = np.array(x, dtype='float32')
        self.data_y = np.array(y, dtype='float32')
        self.data_v = np.array(v, dtype='float32')
        if self.data_x.ndim == 1:
            self.data_x = self.data_x.reshape(-1, 1)
        if self.data_y.ndim == 1:
            self.data_y = self.data_y.reshape(-1, 1)
        self.batch_size = batch_size
        if normalize:
            from sklearn.preprocessing import StandardScaler
            self.data_v = StandardScaler().fit_transform(self.data_v)
        self.sample_size = len(x)
        self.full_index = np.arange(self.sample_size)
        np.random.shuffle(self.full_index)
        self.idx_gen = self._idx_generator(self.sample_size)

    def _idx_generator(self, sample_size):
        bs = self.batch_size
        while True:
            for step in range(math.ceil(sample_size / bs)):
                if (step + 1) * bs <= sample_size:
                    yield self.full_index[step * bs:(step + 1) * bs]
                else:
                    yield np.hstack([self.full_index[step * bs:],
                                     self.full_index[:((step + 1) * bs - sample_size)]])
                    np.random.shuffle(self.full_index)

    def next_batch(self):
        idx = next(self.idx_gen)
        return self.data_x[idx, :], self.data_y[idx, :], self.data_v[idx, :]

    def load_all(self):
        return self.data_x, self.data_y, self.data_v


class Sim_Hirano_Imbens_sampler(Base_sampler):
    """causal_samplers.py:40-67: v ~ Exp(1); x ~ Exp(scale 1/(v0+v1));
    y ~ N(x + (v0+v2) exp(-x (v0+v2)), 1); v standardised."""

    def __init__(self, batch_size=32, N=20000, v_dim=200, seed=0):
        np.random.seed(seed)
        v = np.random.exponential(scale=1.0, size=(N, v_dim))
        rate = v[:, 0] + v[:, 1]
        x = np.random.exponential(scale=1 / rate)
        y = np.random.normal(x + (v[:, 0] + v[:, 2]) * np.exp(-x * (v[:, 0] + v[:, 2])), 1)
        super().__init__(x.reshape(-1, 1), y.reshape(-1, 1), v, batch_size=batch_size, normalize=True)


class Gaussian_sampler(object):
    """prior_samplers.py:4-69.  NOTE the reference reseeds NumPy's global generator
    to 1024 in the constructor (:24); kept, because it fixes the mini-batch index
    stream of every model constructed afterwards."""

    def __init__(self, mean, sd=1, N=20000):
        self.total_size = N
        self.mean = mean
        self.sd = sd
        np.random.seed(1024)
        self.X = np.random.normal(self.mean, self.sd, (self.total_size, len(self.mean))).astype('float32')

    def train(self, batch_size, label=False):
        indx = np.random.randint(low=0, high=self.total_size, size=batch_size)
        return self.X[indx, :]

    def get_batch(self, batch_size):
        return np.random.normal(self.mean, self.sd, (batch_size, len(self.mean))).astype('float32')

    def load_all(self):
        return self.X


def acic_shaped_binary(n=20000, p=100, seed=0):
    """ACIC-*shaped* synthetic binary-treatment data (the ACIC csv
    files are not in the reference repo): v ~ N(0,1) standardised,
    x ~ Bernoulli(sigmoid(v0+v1)), y = x + v0 + 0.5 v2 + N(0,1)."""
    rs = np.random.RandomState(seed)
    v = rs.standard_normal((n, p))
    x = (rs.uniform(size=n) < 1.0 / (1.0 + np.exp(-(v[:, 0] + v[:, 1])))).astype(np.float32)
    y = (x + v[:, 0] + 0.5 * v[:, 2] + rs.standard_normal(n)).astype(np.float32)
    return x.reshape(-1, 1), y.reshape(-1, 1), _standardize(v)


def simulate_z_hetero(n=20000, k=3, d=20 - 1, seed=42):
    """simulators.py:163-204 (same NumPy global-generator call order): X (n, d) = noisy low-rank
    projection of a k-dim latent Z, Y (n,) = sin(Z w) + heteroskedastic noise."""
    np.random.seed(seed)
    Z = np.random.randn(n, k)
    A = np.random.randn(d, k)
    X = 0.2 * Z @ A.T + 0.1 * np.random.randn(n, d)
    w = np.random.randn(k)
    u = np.random.randn(k)
    mean_Y = np.sin(Z @ w)
    std_Y = 0.1 + 0.5 * 1 / (1 + np.exp(-(Z @ u)))
    Y = mean_Y + std_Y * np.random.randn(n)
    return X, Y
