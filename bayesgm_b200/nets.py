"""Host-side weight containers for the reference's deterministic networks.

The reference builds Keras models (`networks/base.py`); here a net is just its
float32 arrays in Keras layout (kernel[in,out], bias[out]) -- the device kernels
consume a packed image built from them by the C library.  Initialisation follows
Keras defaults: glorot-uniform kernels, zero biases.
"""
import ctypes as C

import numpy as np

from . import _lib


class DenseNet(object):
    """`BaseFullyConnectedNet` (networks/base.py:4-51): Dense stack, LeakyReLU(0.2)
    between layers, linear last layer, batchnorm=False."""

    def __init__(self, input_dim, output_dim, model_name, nb_units, rng=None):
        self.input_dim = int(input_dim)
        self.output_dim = int(output_dim)
        self.model_name = model_name
        self.nb_units = [int(u) for u in nb_units]
        self.dims = [self.input_dim] + self.nb_units + [self.output_dim]
        rng = rng if rng is not None else np.random
        self.layers = []
        for i in range(len(self.dims) - 1):
            fan_in, fan_out = self.dims[i], self.dims[i + 1]
            lim = np.sqrt(6.0 / (fan_in + fan_out))
            W = rng.uniform(-lim, lim, size=(fan_in, fan_out)).astype(np.float32)
            self.layers.append([W, np.zeros(fan_out, np.float32)])

    # Keras-style accessors: [kernel0, bias0, kernel1, bias1, ...]
    def get_weights(self):
        return [a.copy() for layer in self.layers for a in layer]

    def set_weights(self, weights):
        assert len(weights) == 2 * len(self.layers), "expected kernel/bias per layer"
        for i, layer in enumerate(self.layers):
            W = np.asarray(weights[2 * i], np.float32)
            b = np.asarray(weights[2 * i + 1], np.float32)
            assert W.shape == layer[0].shape and b.shape == layer[1].shape, \
                "%s layer %d: shape mismatch" % (self.model_name, i)
            layer[0], layer[1] = W.copy(), b.copy()

    @property
    def trainable_variables(self):
        return [a for layer in self.layers for a in layer]

    def flat_params(self):
        return np.ascontiguousarray(
            np.concatenate([np.concatenate([W.ravel(), b.ravel()]) for W, b in self.layers]).astype(np.float32))

    def desc(self):
        """(bgm_net_desc, keep-alive tuple) for the C ABI."""
        dims = (C.c_int * len(self.dims))(*self.dims)
        flat = self.flat_params()
        d = _lib.NetDesc(len(self.layers), C.cast(dims, C.POINTER(C.c_int)),
                         flat.ctypes.data_as(C.POINTER(C.c_float)))
        return d, (dims, flat)

    def as_oracle_layers(self):
        return [(W, b) for W, b in self.layers]

    def load_flat(self, flat):
        """Inverse of flat_params()."""
        o = 0
        for layer in self.layers:
            for i in (0, 1):
                a = layer[i]
                layer[i] = np.array(flat[o:o + a.size], np.float32).reshape(a.shape)
                o += a.size


class BayesDenseNet(object):
    """`BayesianFullyConnectedNet` (networks/bnn.py:4-38): BatchNormalization on the input, then
    `tfp.layers.DenseFlipout` layers (LeakyReLU(0.2) between them).  Per layer: kernel posterior loc
    and untransformed scale rho (sigma = finfo(float32).eps + softplus(rho)), deterministic bias loc.
    TFP 0.18 default initialisers: loc, bias ~ N(0, 0.1^2), rho ~ N(-3, 0.1^2); Keras BN defaults."""

    def __init__(self, input_dim, output_dim, model_name, nb_units, rng=None):
        self.input_dim, self.output_dim = int(input_dim), int(output_dim)
        self.model_name = model_name
        self.nb_units = [int(u) for u in nb_units]
        self.dims = [self.input_dim] + self.nb_units + [self.output_dim]
        rng = rng if rng is not None else np.random
        k = self.input_dim
        self.bn = dict(gamma=np.ones(k, np.float32), beta=np.zeros(k, np.float32),
                       mean=np.zeros(k, np.float32), var=np.ones(k, np.float32))
        self.layers = []
        for i in range(len(self.dims) - 1):
            fi, fo = self.dims[i], self.dims[i + 1]
            self.layers.append([(0.1 * rng.standard_normal((fi, fo))).astype(np.float32),
                                (-3.0 + 0.1 * rng.standard_normal((fi, fo))).astype(np.float32),
                                (0.1 * rng.standard_normal(fo)).astype(np.float32)])

    # Keras order: BN [gamma, beta, moving_mean, moving_variance], then per DenseFlipout layer
    # [kernel_posterior_loc, kernel_posterior_untransformed_scale, bias_posterior_loc]
    def get_weights(self):
        out = [self.bn[k].copy() for k in ('gamma', 'beta', 'mean', 'var')]
        for layer in self.layers:
            out += [a.copy() for a in layer]
        return out

    def set_weights(self, weights):
        assert len(weights) == 4 + 3 * len(self.layers), "expected BN(4) + loc/rho/bias per DenseFlipout layer"
        for k, w in zip(('gamma', 'beta', 'mean', 'var'), weights[:4]):
            w = np.asarray(w, np.float32)
            assert w.shape == (self.input_dim,)
            self.bn[k] = w.copy()
        for i, layer in enumerate(self.layers):
            for j in range(3):
                w = np.asarray(weights[4 + 3 * i + j], np.float32)
                assert w.shape == layer[j].shape, "%s layer %d: shape mismatch" % (self.model_name, i)
                layer[j] = w.copy()

    def flat_params(self):
        """Trainable parameters, Keras trainable_variables order: gamma | beta | (loc, rho, bias) per layer."""
        parts = [self.bn['gamma'], self.bn['beta']] + [a.ravel() for layer in self.layers for a in layer]
        return np.ascontiguousarray(np.concatenate(parts).astype(np.float32))

    def load_flat(self, flat):
        k = self.input_dim
        self.bn['gamma'], self.bn['beta'] = np.array(flat[:k], np.float32), np.array(flat[k:2 * k], np.float32)
        o = 2 * k
        for layer in self.layers:
            for j in range(3):
                a = layer[j]
                layer[j] = np.array(flat[o:o + a.size], np.float32).reshape(a.shape)
                o += a.size

    def desc(self):
        """(bgm_bnn_net_desc, keep-alive tuple) for the C ABI."""
        dims = (C.c_int * len(self.dims))(*self.dims)
        bn = np.ascontiguousarray(np.concatenate([self.bn['gamma'], self.bn['beta']]), np.float32)
        flat = np.ascontiguousarray(np.concatenate([a.ravel() for layer in self.layers for a in layer]), np.float32)
        d = _lib.BnnNetDesc(len(self.layers), C.cast(dims, C.POINTER(C.c_int)), bn.ctypes.data_as(C.POINTER(C.c_float)),
                            flat.ctypes.data_as(C.POINTER(C.c_float)))
        return d, (dims, bn, flat)

    def as_oracle_params(self):
        return dict(bn=dict(self.bn), layers=[tuple(layer) for layer in self.layers])


class VariationalNet(object):
    """`BaseVariationalNet` (networks/base.py:53-117): BatchNormalization on the input,
    Dense+LeakyReLU(0.2) hidden layers, a mean head and a softplus(+eps) variance head.
    Keras defaults: BN gamma 1, beta 0, moving mean 0, moving variance 1, eps 1e-3."""

    def __init__(self, input_dim, output_dim, model_name, nb_units, rng=None):
        self.input_dim, self.output_dim = int(input_dim), int(output_dim)
        self.model_name = model_name
        self.nb_units = [int(u) for u in nb_units]
        rng = rng if rng is not None else np.random
        self.bn = dict(gamma=np.ones(self.input_dim, np.float32), beta=np.zeros(self.input_dim, np.float32),
                       mean=np.zeros(self.input_dim, np.float32), var=np.ones(self.input_dim, np.float32))
        dims = [self.input_dim] + self.nb_units

        def dense(fi, fo):
            lim = np.sqrt(6.0 / (fi + fo))
            return [rng.uniform(-lim, lim, size=(fi, fo)).astype(np.float32), np.zeros(fo, np.float32)]
        self.hidden = [dense(dims[i], dims[i + 1]) for i in range(len(self.nb_units))]
        self.mean = dense(dims[-1], self.output_dim)
        self.var = dense(dims[-1], self.output_dim)

    # Keras order: BN [gamma, beta, moving_mean, moving_variance], hidden kernel/bias..., mean, var
    def get_weights(self):
        out = [self.bn[k].copy() for k in ('gamma', 'beta', 'mean', 'var')]
        for W, b in self.hidden + [self.mean, self.var]:
            out += [W.copy(), b.copy()]
        return out

    def set_weights(self, weights):
        assert len(weights) == 4 + 2 * (len(self.hidden) + 2), "expected BN(4) + kernel/bias per Dense layer"
        for k, w in zip(('gamma', 'beta', 'mean', 'var'), weights[:4]):
            w = np.asarray(w, np.float32)
            assert w.shape == (self.input_dim,)
            self.bn[k] = w.copy()
        layers = self.hidden + [self.mean, self.var]
        for i, layer in enumerate(layers):
            W = np.asarray(weights[4 + 2 * i], np.float32)
            b = np.asarray(weights[5 + 2 * i], np.float32)
            assert W.shape == layer[0].shape and b.shape == layer[1].shape, \
                "%s layer %d: shape mismatch" % (self.model_name, i)
            layer[0], layer[1] = W.copy(), b.copy()

    def desc(self):
        """(bgm_varnet_desc, keep-alive tuple) for the C ABI."""
        units = (C.c_int * len(self.nb_units))(*self.nb_units)
        bn = np.ascontiguousarray(np.concatenate([self.bn[k] for k in ('gamma', 'beta', 'mean', 'var')]), np.float32)
        flat = lambda layers: np.ascontiguousarray(
            np.concatenate([np.concatenate([W.ravel(), b.ravel()]) for W, b in layers]).astype(np.float32))
        hp, mp, vp = flat(self.hidden), flat([self.mean]), flat([self.var])
        fp = lambda a: a.ctypes.data_as(C.POINTER(C.c_float))
        d = _lib.VarNetDesc(self.input_dim, self.output_dim, len(self.nb_units),
                            C.cast(units, C.POINTER(C.c_int)), fp(bn), fp(hp), fp(mp), fp(vp))
        return d, (units, bn, hp, mp, vp)

    def with_extra_columns(self, cols):
        """A copy whose mean / variance heads carry, after the x_dim real outputs, one more output per entry of
        `cols` (a copy of that output's head column): a feature listed m times in `ind_x1` is observed through
        m identical outputs, which adds its likelihood term m times like the gather of bgm/base.py:689-700."""
        cols = np.asarray(cols, dtype=np.int64)
        out = VariationalNet.__new__(VariationalNet)
        out.input_dim, out.output_dim = self.input_dim, self.output_dim + len(cols)
        out.model_name, out.nb_units = self.model_name, list(self.nb_units)
        out.bn = {k: v.copy() for k, v in self.bn.items()}
        out.hidden = [[W.copy(), b.copy()] for W, b in self.hidden]
        ext = lambda layer: [np.concatenate([layer[0], layer[0][:, cols]], axis=1).astype(np.float32),
                             np.concatenate([layer[1], layer[1][cols]]).astype(np.float32)]
        out.mean, out.var = ext(self.mean), ext(self.var)
        return out

    def as_oracle_params(self):
        return dict(bn=dict(self.bn), hidden=[(W, b) for W, b in self.hidden],
                    mean=(self.mean[0], self.mean[1]), var=(self.var[0], self.var[1]))


class DiscNet(object):
    """`Discriminator` (networks/base.py:338-385): Dense -> BatchNormalization (batch
    statistics) -> tanh blocks, then Dense(1).  Keras defaults: glorot-uniform kernels, zero
    biases, gamma 1, beta 0."""

    def __init__(self, input_dim, model_name, nb_units, rng=None):
        self.input_dim = int(input_dim)
        self.model_name = model_name
        self.nb_units = [int(u) for u in nb_units]
        self.dims = [self.input_dim] + self.nb_units + [1]
        rng = rng if rng is not None else np.random
        self.layers, self.bns = [], []
        for i in range(len(self.dims) - 1):
            fi, fo = self.dims[i], self.dims[i + 1]
            lim = np.sqrt(6.0 / (fi + fo))
            self.layers.append([rng.uniform(-lim, lim, size=(fi, fo)).astype(np.float32), np.zeros(fo, np.float32)])
            if i < len(self.nb_units):
                self.bns.append([np.ones(fo, np.float32), np.zeros(fo, np.float32)])

    def trainable_list(self):
        """Keras trainable_variables order: per block kernel, bias, gamma, beta; then output kernel, bias."""
        out = []
        for (W, b), (g, be) in zip(self.layers[:-1], self.bns):
            out += [W, b, g, be]
        return out + list(self.layers[-1])

    def flat_params(self):
        return np.ascontiguousarray(np.concatenate([a.ravel() for a in self.trainable_list()]).astype(np.float32))

    def load_flat(self, flat):
        o = 0
        for a in self.trainable_list():
            a[...] = flat[o:o + a.size].reshape(a.shape)
            o += a.size

    def set_trainable(self, arrays):
        for a, src in zip(self.trainable_list(), arrays):
            a[...] = np.asarray(src, np.float32).reshape(a.shape)

    def desc(self):
        dims = (C.c_int * len(self.dims))(*self.dims)
        flat = self.flat_params()
        d = _lib.DiscDesc(len(self.nb_units), C.cast(dims, C.POINTER(C.c_int)), flat.ctypes.data_as(C.POINTER(C.c_float)))
        return d, (dims, flat)

    def as_oracle_params(self):
        return dict(layers=[(W, b) for W, b in self.layers],
                    bns=[dict(gamma=g, beta=be) for g, be in self.bns])
