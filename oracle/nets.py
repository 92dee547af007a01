"""Oracle restatement of the reference's deterministic networks (test infrastructure).

Follows `src/bayesgm/models/networks/base.py`:
  * BaseFullyConnectedNet  :4-51   -> `mlp_forward`
  * BaseVariationalNet     :53-117 -> `variational_forward`, `reparameterize`
  * Discriminator          :338-385 -> `discriminator_forward`
Weights are plain float32 arrays in Keras layout: kernel[in, out], bias[out].
Keras defaults restated (TF 2.10, not vendored): Dense glorot-uniform kernel / zero
bias; LeakyReLU(alpha=0.2); BatchNormalization momentum .99, eps 1e-3, gamma 1,
beta 0, moving mean 0, moving var 1, biased batch variance in training mode.
"""
import numpy as np

LEAKY_SLOPE = np.float32(0.2)
BN_EPS = np.float32(1e-3)


def glorot_uniform(rs, fan_in, fan_out):
    """Keras `glorot_uniform`: U(-l, l), l = sqrt(6 / (fan_in + fan_out))."""
    lim = np.sqrt(6.0 / (fan_in + fan_out))
    return rs.uniform(-lim, lim, size=(fan_in, fan_out)).astype(np.float32)


def init_mlp(rs, dims, bias_scale=0.0):
    """[(W[in,out], b[out])] for a Dense stack; `dims` = [in, h1, ..., out].

    networks/base.py:17-26 builds `len(nb_units)+1` Dense layers.  Keras biases
    start at zero; `bias_scale` > 0 draws small random biases so that parity tests
    exercise the bias path too.
    """
    layers = []
    for i in range(len(dims) - 1):
        W = glorot_uniform(rs, dims[i], dims[i + 1])
        if bias_scale > 0:
            b = (bias_scale * rs.standard_normal(dims[i + 1])).astype(np.float32)
        else:
            b = np.zeros(dims[i + 1], np.float32)
        layers.append((W, b))
    return layers


def leaky_relu(x):
    return np.where(x > 0, x, LEAKY_SLOPE * x).astype(np.float32)


def softplus(t):
    """tf.nn.softplus = log(1 + exp(t)), evaluated stably in float32."""
    t = np.asarray(t, np.float32)
    return (np.maximum(t, 0) + np.log1p(np.exp(-np.abs(t)))).astype(np.float32)


def sigmoid(t):
    t = np.asarray(t, np.float32)
    return (1.0 / (1.0 + np.exp(-t))).astype(np.float32)


def mlp_forward(layers, x):
    """BaseFullyConnectedNet.call, networks/base.py:30-51 (batchnorm=False, so the
    BatchNormalization layers created at :25 are never applied, :43-44)."""
    h = np.asarray(x, np.float32)
    for W, b in layers[:-1]:
        h = leaky_relu(h @ W + b)
    W, b = layers[-1]
    return (h @ W + b).astype(np.float32)


def mlp_forward_cache(layers, x):
    """Forward pass keeping pre-activations (for the hand-written backward)."""
    h = np.asarray(x, np.float32)
    pre = []
    for W, b in layers[:-1]:
        a = (h @ W + b).astype(np.float32)
        pre.append(a)
        h = leaky_relu(a)
    return h, pre


def init_variational(rs, z_dim, x_dim, units, bias_scale=0.0, bn_random=False):
    """Parameters of BaseVariationalNet (networks/base.py:58-96)."""
    hidden = init_mlp(rs, [z_dim] + list(units), bias_scale)
    last = units[-1]
    mean = (glorot_uniform(rs, last, x_dim),
            (bias_scale * rs.standard_normal(x_dim)).astype(np.float32))
    var = (glorot_uniform(rs, last, x_dim),
           (bias_scale * rs.standard_normal(x_dim)).astype(np.float32))
    if bn_random:  # a "trained" BN: non-trivial moving statistics
        bn = dict(gamma=(1 + 0.1 * rs.standard_normal(z_dim)).astype(np.float32),
                  beta=(0.1 * rs.standard_normal(z_dim)).astype(np.float32),
                  mean=(0.1 * rs.standard_normal(z_dim)).astype(np.float32),
                  var=(1 + 0.2 * rs.uniform(size=z_dim)).astype(np.float32))
    else:
        bn = dict(gamma=np.ones(z_dim, np.float32), beta=np.zeros(z_dim, np.float32),
                  mean=np.zeros(z_dim, np.float32), var=np.ones(z_dim, np.float32))
    return dict(bn=bn, hidden=hidden, mean=mean, var=var)


def bn_inference(bn, x):
    """Keras BatchNormalization, training=False: gamma*(x-mu)/sqrt(var+eps)+beta."""
    inv = (bn['gamma'] / np.sqrt(bn['var'] + BN_EPS)).astype(np.float32)
    return ((x - bn['mean']) * inv + bn['beta']).astype(np.float32)


def variational_forward(p, z, eps=1e-6):
    """BaseVariationalNet.call(training=False), networks/base.py:98-111."""
    h = bn_inference(p['bn'], np.asarray(z, np.float32))
    for W, b in p['hidden']:
        h = leaky_relu(h @ W + b)
    mean = (h @ p['mean'][0] + p['mean'][1]).astype(np.float32)
    var = softplus(h @ p['var'][0] + p['var'][1]) + np.float32(eps)
    return mean, var.astype(np.float32)


def reparameterize(mean, var, noise):
    """networks/base.py:113-117 with the N(0,1) draw passed in."""
    return (noise * np.sqrt(var) + mean).astype(np.float32)


def bn_training(x, gamma, beta):
    """Keras BN, training=True: biased batch statistics, eps=1e-3."""
    mu = x.mean(axis=0)
    var = ((x - mu) ** 2).mean(axis=0)
    return ((x - mu) / np.sqrt(var + BN_EPS) * gamma + beta).astype(np.float32)


def init_discriminator(rs, in_dim, units):
    layers = init_mlp(rs, [in_dim] + list(units) + [1])
    bns = [dict(gamma=np.ones(u, np.float32), beta=np.zeros(u, np.float32)) for u in units]
    return dict(layers=layers, bns=bns)


def discriminator_forward(p, x):
    """Discriminator.call, networks/base.py:364-385: Dense -> BN(training) -> tanh,
    last Dense linear.  BN is in training mode in every call."""
    h = np.asarray(x, np.float32)
    for (W, b), bn in zip(p['layers'][:-1], p['bns']):
        h = np.tanh(bn_training(h @ W + b, bn['gamma'], bn['beta'])).astype(np.float32)
    W, b = p['layers'][-1]
    return (h @ W + b).astype(np.float32)
