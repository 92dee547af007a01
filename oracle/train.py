"""Oracle restatement of CausalBGM's EGM training steps (test infrastructure).

Follows `src/bayesgm/models/causalbgm/base.py`:
  * train_disc_step   :305-330 -> `disc_step`
  * train_gen_step    :332-377 -> `gen_step`
  * egm_init          :380-431 -> `egm_init`
  * update_g/h/f_net  :156-243 -> `update_g`, `update_h`, `update_f`
and Keras `Adam` (TF 2.10 optimizer_v2) -> `Adam`.  Gradients come from torch
autograd on the CPU in float32 (an independent differentiator: the CUDA kernels use
hand-derived backward and double-backward formulas).  Networks are the Keras-layout arrays
of oracle.nets; the discriminator is Dense -> BatchNormalization(training) -> tanh
(networks/base.py:364-385) with batch statistics in every call.
"""
import numpy as np
import torch

BN_EPS = 1e-3


def to_t(layers, requires_grad=False):
    return [(torch.tensor(W, dtype=torch.float32, requires_grad=requires_grad),
             torch.tensor(b, dtype=torch.float32, requires_grad=requires_grad)) for W, b in layers]


def mlp(layers, x):
    """BaseFullyConnectedNet.call, networks/base.py:30-51."""
    h = x
    for W, b in layers[:-1]:
        h = torch.nn.functional.leaky_relu(h @ W + b, 0.2)
    W, b = layers[-1]
    return h @ W + b


def disc(p, x):
    """Discriminator.call, networks/base.py:364-385 (BN in training mode, biased variance)."""
    h = x
    for (W, b), (gamma, beta) in zip(p['layers'][:-1], p['bn']):
        a = h @ W + b
        mu = a.mean(dim=0)
        var = ((a - mu) ** 2).mean(dim=0)
        h = torch.tanh((a - mu) / torch.sqrt(var + BN_EPS) * gamma + beta)
    W, b = p['layers'][-1]
    return h @ W + b


def disc_to_t(p, requires_grad=True):
    return dict(layers=to_t(p['layers'], requires_grad),
                bn=[(torch.tensor(bn['gamma'], dtype=torch.float32, requires_grad=requires_grad),
                     torch.tensor(bn['beta'], dtype=torch.float32, requires_grad=requires_grad)) for bn in p['bns']])


def disc_param_list(pt):
    """Keras trainable_variables order of the Discriminator: per hidden block Dense kernel,
    bias, BN gamma, beta; then the output Dense kernel, bias."""
    out = []
    for (W, b), (g, be) in zip(pt['layers'][:-1], pt['bn']):
        out += [W, b, g, be]
    out += list(pt['layers'][-1])
    return out


def mlp_param_list(layers_t):
    return [a for W, b in layers_t for a in (W, b)]


class Adam(object):
    """Keras Adam, dense update: m,v moments, lr_t = lr*sqrt(1-b2^t)/(1-b1^t),
    theta -= lr_t * m / (sqrt(v) + 1e-7)."""

    def __init__(self, lr, beta_1=0.9, beta_2=0.99, epsilon=1e-7):
        self.lr, self.b1, self.b2, self.eps, self.t = lr, beta_1, beta_2, epsilon, 0
        self.m, self.v = None, None

    def apply(self, params, grads):
        """params, grads: lists of float32 NumPy arrays; params updated in place."""
        if self.m is None:
            self.m = [np.zeros_like(p) for p in params]
            self.v = [np.zeros_like(p) for p in params]
        self.t += 1
        f32 = np.float32
        lr_t = f32(self.lr * np.sqrt(1.0 - self.b2 ** self.t) / (1.0 - self.b1 ** self.t))
        for p, g, m, v in zip(params, grads, self.m, self.v):
            g = g.astype(f32)
            m += (g - m) * f32(1 - self.b1)
            v += (g * g - v) * f32(1 - self.b2)
            p -= lr_t * m / (np.sqrt(v) + f32(self.eps))


def disc_step(params, nets, dz, batch_z, batch_v, epsilon):
    """causalbgm/base.py:305-330.  Returns (dz_loss, d_loss, grads of dz_net's trainable
    variables in Keras order).  `epsilon` is the tf.random.uniform([]) draw (:307)."""
    e = to_t(nets['e'])
    pt = disc_to_t(dz)
    z = torch.tensor(batch_z, dtype=torch.float32)
    v = torch.tensor(batch_v, dtype=torch.float32)
    z_ = mlp(e, v).detach()                                               # :310 (only dz_net is trained)
    z_hat = (z * epsilon + z_ * (1 - epsilon)).requires_grad_(True)       # :311
    d_hat = disc(pt, z_hat)                                               # :312
    d_ = disc(pt, z_)                                                     # :314
    d = disc(pt, z)                                                       # :315
    dz_loss = -d.mean() + d_.mean()                                       # :316
    grad_z = torch.autograd.grad(d_hat.sum(), z_hat, create_graph=True)[0]    # :319
    grad_norm = torch.sqrt((grad_z ** 2).sum(dim=1))                      # :320
    gp = ((grad_norm - 1.0) ** 2).mean()                                  # :321
    d_loss = dz_loss + 10 * gp                                            # :323
    plist = disc_param_list(pt)
    grads = torch.autograd.grad(d_loss, plist, allow_unused=True)
    grads = [np.zeros(tuple(p.shape), np.float32) if g is None else g.numpy() for g, p in zip(grads, plist)]
    return float(dz_loss.detach()), float(d_loss.detach()), grads


def _split_z(params, z):
    d0, d1, d2, _ = params['z_dims']
    return z[:, :d0], z[:, d0:d0 + d1], z[:, d0 + d1:d0 + d1 + d2]


def gen_step(params, nets, dz, batch_z, batch_v, batch_x, batch_y):
    """causalbgm/base.py:332-377.  Returns (losses tuple as :377, grads dict per net in
    Keras order: g, e, f, h)."""
    p = params['v_dim']
    g, e, f, h = [to_t(nets[k], True) for k in ('g', 'e', 'f', 'h')]
    pt = disc_to_t(dz, requires_grad=False)
    z = torch.tensor(batch_z, dtype=torch.float32)
    v = torch.tensor(batch_v, dtype=torch.float32)
    x = torch.tensor(batch_x, dtype=torch.float32)
    y = torch.tensor(batch_y, dtype=torch.float32)
    g_out = mlp(g, z)
    v_ = g_out[:, :p]                                                     # :336
    sig_loss = (g_out[:, -1] ** 2).mean()                                 # :337
    z_ = mlp(e, v)                                                        # :338
    z0, z1, z2 = _split_z(params, z_)
    z__ = mlp(e, v_)                                                      # :344
    v__ = mlp(g, z_)[:, :p]                                               # :345
    d_ = disc(pt, z_)                                                     # :347
    l2_v = ((v - v__) ** 2).mean()                                        # :349
    l2_z = ((z - z__) ** 2).mean()                                        # :350
    e_adv = -d_.mean()                                                    # :352
    f_out = mlp(f, torch.cat([z0, z1, x], dim=-1))
    y_ = f_out[:, :1]                                                     # :354
    sig_loss = sig_loss + (f_out[:, -1] ** 2).mean()                      # :355
    h_out = mlp(h, torch.cat([z0, z2], dim=-1))
    x_ = h_out[:, :1]                                                     # :357
    sig_loss = sig_loss + (h_out[:, -1] ** 2).mean()                      # :358
    if params['binary_treatment']:                                        # :361
        l2_x = torch.nn.functional.binary_cross_entropy_with_logits(x_, x, reduction='mean')
    else:
        l2_x = ((x_ - x) ** 2).mean()                                     # :365
    l2_y = ((y_ - y) ** 2).mean()                                         # :366
    use_z_rec = float(params.get('use_z_rec', True))
    loss = e_adv + (l2_v + use_z_rec * l2_z) + (l2_x + l2_y) + 0.001 * sig_loss   # :367
    plist = mlp_param_list(g) + mlp_param_list(e) + mlp_param_list(f) + mlp_param_list(h)
    grads = torch.autograd.grad(loss, plist)
    grads = [a.numpy() for a in grads]
    out, i = {}, 0
    for name, net in (('g', g), ('e', e), ('f', f), ('h', h)):
        k = 2 * len(net)
        out[name] = grads[i:i + k]
        i += k
    losses = tuple(float(a.detach()) for a in (e_adv, l2_v, l2_z, l2_x, l2_y, loss))
    return losses, out


def flat_params(layers):
    return [a for W, b in layers for a in (W, b)]


def disc_flat_params(dz):
    out = []
    for (W, b), bn in zip(dz['layers'][:-1], dz['bns']):
        out += [W, b, bn['gamma'], bn['beta']]
    out += list(dz['layers'][-1])
    return out


class EgmTrainer(object):
    """Holds nets + the two Keras-Adam optimizers of the EGM phase (:86-87) and applies
    the steps in place.  `nets` / `dz` arrays are modified."""

    def __init__(self, params, nets, dz):
        self.params, self.nets, self.dz = params, nets, dz
        for k in ('g', 'e', 'f', 'h'):
            nets[k] = [(np.array(W, np.float32), np.array(b, np.float32)) for W, b in nets[k]]
        self.g_opt = Adam(params['lr'], 0.9, 0.99)
        self.d_opt = Adam(params['lr'], 0.9, 0.99)

    def train_disc_step(self, batch_z, batch_v, epsilon):
        dz_loss, d_loss, grads = disc_step(self.params, self.nets, self.dz, batch_z, batch_v, epsilon)
        self.d_opt.apply(disc_flat_params(self.dz), grads)
        return dz_loss, d_loss

    def train_gen_step(self, batch_z, batch_v, batch_x, batch_y):
        losses, grads = gen_step(self.params, self.nets, self.dz, batch_z, batch_v, batch_x, batch_y)
        plist = sum([flat_params(self.nets[k]) for k in ('g', 'e', 'f', 'h')], [])
        glist = grads['g'] + grads['e'] + grads['f'] + grads['h']
        self.g_opt.apply(plist, glist)
        return losses

    def egm_init(self, data, egm_n_iter, batch_size, z_sampler, eps_fn):
        """causalbgm/base.py:403-417, with the reference's NumPy call order:
        g_d_freq x (choice, get_batch) then (get_batch, choice)."""
        data_x, data_y, data_v = data
        out = None
        for _ in range(egm_n_iter + 1):
            for _ in range(self.params['g_d_freq']):
                idx = np.random.choice(len(data_x), batch_size, replace=False)     # :406
                bz = z_sampler.get_batch(batch_size)                               # :407
                dl = self.train_disc_step(bz, data_v[idx, :], eps_fn())
            bz = z_sampler.get_batch(batch_size)                                   # :412
            idx = np.random.choice(len(data_x), batch_size, replace=False)         # :413
            gl = self.train_gen_step(bz, data_v[idx, :], data_x[idx, :], data_y[idx, :])
            out = (dl, gl)
        return out


def update_net_grads(params, nets, which, batch_z, batch_x, batch_y, batch_v, eps=1e-6):
    """update_g_net / update_h_net / update_f_net (:156-243): (loss, mse, grads)."""
    p = params['v_dim']
    z = torch.tensor(batch_z, dtype=torch.float32)
    net = to_t(nets[which], True)
    sp = torch.nn.functional.softplus
    if which == 'g':
        out = mlp(net, z)
        mu = out[:, :p]
        s2 = params['sigma_v'] ** 2 if 'sigma_v' in params else sp(out[:, -1]) + eps
        v = torch.tensor(batch_v, dtype=torch.float32)
        mse = ((v - mu) ** 2).mean()
        loss = (((v - mu) ** 2).sum(dim=1) / (2 * s2) + p * torch.log(torch.as_tensor(s2)) / 2).mean()
    else:
        z0, z1, z2 = _split_z(params, z)
        x = torch.tensor(batch_x, dtype=torch.float32)
        if which == 'h':
            out = mlp(net, torch.cat([z0, z2], dim=-1))
            mu, tgt = out[:, :1], x
            key = 'sigma_x'
        else:
            out = mlp(net, torch.cat([z0, z1, x], dim=-1))
            mu, tgt = out[:, :1], torch.tensor(batch_y, dtype=torch.float32)
            key = 'sigma_y'
        if which == 'h' and params['binary_treatment']:
            mse = torch.nn.functional.binary_cross_entropy_with_logits(mu, tgt, reduction='mean')
            loss = mse
        else:
            s2 = params[key] ** 2 if key in params else sp(out[:, -1]) + eps
            mse = ((tgt - mu) ** 2).mean()
            loss = (((tgt - mu) ** 2).sum(dim=1) / (2 * s2) + torch.log(torch.as_tensor(s2)) / 2).mean()
    grads = torch.autograd.grad(loss, mlp_param_list(net))
    return float(loss.detach()), float(mse.detach()), [a.numpy() for a in grads]


def latent_grad(params, nets, batch_z, batch_x, batch_y, batch_v, eps=1e-6):
    """update_latent_variable_sgd (:246-295): (loss_postrior_z, gradient w.r.t. the batch rows)."""
    p = params['v_dim']
    sp = torch.nn.functional.softplus
    g, f, h = [to_t(nets[k]) for k in ('g', 'f', 'h')]
    z = torch.tensor(batch_z, dtype=torch.float32, requires_grad=True)
    x = torch.tensor(batch_x, dtype=torch.float32)
    y = torch.tensor(batch_y, dtype=torch.float32)
    v = torch.tensor(batch_v, dtype=torch.float32)
    z0, z1, z2 = _split_z(params, z)
    go = mlp(g, z)
    s2v = params['sigma_v'] ** 2 if 'sigma_v' in params else sp(go[:, -1]) + eps
    loss_v = (((v - go[:, :p]) ** 2).sum(dim=1) / (2 * s2v) + p * torch.log(torch.as_tensor(s2v)) / 2).mean()
    ho = mlp(h, torch.cat([z0, z2], dim=-1))
    if params['binary_treatment']:
        loss_x = torch.nn.functional.binary_cross_entropy_with_logits(ho[:, :1], x, reduction='mean')
    else:
        s2x = params['sigma_x'] ** 2 if 'sigma_x' in params else sp(ho[:, -1]) + eps
        loss_x = (((x - ho[:, :1]) ** 2).sum(dim=1) / (2 * s2x) + torch.log(torch.as_tensor(s2x)) / 2).mean()
    fo = mlp(f, torch.cat([z0, z1, x], dim=-1))
    s2y = params['sigma_y'] ** 2 if 'sigma_y' in params else sp(fo[:, -1]) + eps
    loss_y = (((y - fo[:, :1]) ** 2).sum(dim=1) / (2 * s2y) + torch.log(torch.as_tensor(s2y)) / 2).mean()
    prior = ((z ** 2).sum(dim=1) / 2).mean()
    loss = loss_v + loss_x + loss_y + prior
    gz = torch.autograd.grad(loss, z)[0]
    return float(loss.detach()), gz.numpy()


class SparseAdam(object):
    """Keras Adam applied to tf.gather'ed rows of a variable (TF 2.10 optimizer_v2
    `_resource_apply_sparse`): m, v of the WHOLE variable decay, the gathered
    rows receive the scaled gradient, and the WHOLE variable moves."""

    def __init__(self, lr, shape, beta_1=0.9, beta_2=0.99, epsilon=1e-7):
        self.lr, self.b1, self.b2, self.eps, self.t = lr, beta_1, beta_2, epsilon, 0
        self.m = np.zeros(shape, np.float32)
        self.v = np.zeros(shape, np.float32)

    def apply(self, table, idx, grad_rows):
        f32 = np.float32
        self.t += 1
        lr_t = f32(self.lr * np.sqrt(1.0 - self.b2 ** self.t) / (1.0 - self.b1 ** self.t))
        self.m *= f32(self.b1)
        self.v *= f32(self.b2)
        self.m[idx] += f32(1 - self.b1) * grad_rows
        self.v[idx] += f32(1 - self.b2) * grad_rows * grad_rows
        table -= lr_t * self.m / (np.sqrt(self.v) + f32(self.eps))


class IterTrainer(object):
    """The iterative phase of CausalBGM.fit (:488-514) on host arrays."""

    def __init__(self, params, nets, data_z):
        self.params, self.nets = params, nets
        for k in ('g', 'f', 'h'):
            nets[k] = [(np.array(W, np.float32), np.array(b, np.float32)) for W, b in nets[k]]
        self.data_z = np.array(data_z, np.float32)
        self.opt = {k: Adam(params['lr_theta'], 0.9, 0.99) for k in ('g', 'h', 'f')}
        self.z_opt = SparseAdam(params['lr_z'], self.data_z.shape)

    def step(self, data, batch_idx):
        data_x, data_y, data_v = data
        bz = self.data_z[batch_idx]
        bx, by, bv = data_x[batch_idx], data_y[batch_idx], data_v[batch_idx]
        out = []
        for k in ('g', 'h', 'f'):                                          # :500-502
            loss, mse, grads = update_net_grads(self.params, self.nets, k, bz, bx, by, bv)
            self.opt[k].apply(flat_params(self.nets[k]), grads)
            out += [loss, mse]
        lz, gz = latent_grad(self.params, self.nets, bz, bx, by, bv)       # :505 (updated nets)
        self.z_opt.apply(self.data_z, batch_idx, gz)
        return out, lz

    def epoch(self, data, batch_size):
        n = len(data[0])
        sample_idx = np.random.choice(n, n, replace=False)                 # :489
        last = None
        for i in range(0, n, batch_size):
            last = self.step(data, sample_idx[i:i + batch_size])
        return last
