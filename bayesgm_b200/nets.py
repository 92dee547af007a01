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
