"""Oracle restatement of BGM's EGM training steps (test infrastructure).

Follows `src/bayesgm/models/bgm/base.py`:
  * train_disc_step :190-245 -> `disc_step`   (dz_net + dx_net, LSGAN targets .9/.1, GP weight `gamma`)
  * train_gen_step  :247-291 -> `gen_step`    (g_net + e_net)
with the generator `BaseVariationalNet` (networks/base.py:53-117) in TRAINING mode as the
reference calls it there (`self.g_net(data_z)`, default training=True): its input
BatchNormalization uses batch statistics and updates the moving statistics
(momentum .99, biased variance) on every call.  Gradients: torch autograd, CPU, float32.
`g` = dict(bn=dict(gamma,beta,mean,var), hidden=[(W,b)], mean=(W,b), var=(W,b)) as in
oracle.nets.init_variational; the N(0,1) draws of `reparameterize` (:113-117) and the
U(0,1) draws of :199-200 are passed in.
"""
import numpy as np
import torch

from .train import to_t, mlp, disc, disc_to_t, disc_param_list, mlp_param_list

BN_EPS = 1e-3
MOMENTUM = 0.99


def g_to_t(g, requires_grad):
    t = lambda a: torch.tensor(a, dtype=torch.float32, requires_grad=requires_grad)
    return dict(gamma=t(g['bn']['gamma']), beta=t(g['bn']['beta']), hidden=to_t(g['hidden'], requires_grad),
                mean=(t(g['mean'][0]), t(g['mean'][1])), var=(t(g['var'][0]), t(g['var'][1])))


def g_param_list(gt):
    """Keras trainable_variables order: BN gamma, beta; hidden kernel/bias...; mean; var."""
    return [gt['gamma'], gt['beta']] + mlp_param_list(gt['hidden']) + list(gt['mean']) + list(gt['var'])


def g_flat_params(g):
    out = [g['bn']['gamma'], g['bn']['beta']]
    for W, b in g['hidden']:
        out += [W, b]
    return out + list(g['mean']) + list(g['var'])


def var_net_train(gt, z, stats):
    """BaseVariationalNet.call(training=True): returns (mean, var); appends the batch
    (mean, biased variance) of the input BN to `stats`."""
    mu = z.mean(dim=0)
    var = ((z - mu) ** 2).mean(dim=0)
    stats.append((mu.detach().numpy().copy(), var.detach().numpy().copy()))
    h = (z - mu) / torch.sqrt(var + BN_EPS) * gt['gamma'] + gt['beta']
    for W, b in gt['hidden']:
        h = torch.nn.functional.leaky_relu(h @ W + b, 0.2)
    mean = h @ gt['mean'][0] + gt['mean'][1]
    v = torch.nn.functional.softplus(h @ gt['var'][0] + gt['var'][1]) + 1e-6
    return mean, v


def update_moving(g, stats):
    """Keras BN moving-statistics update, once per training-mode call, in call order."""
    for mu, var in stats:
        g['bn']['mean'] = (g['bn']['mean'] * np.float32(MOMENTUM) + mu * np.float32(1 - MOMENTUM)).astype(np.float32)
        g['bn']['var'] = (g['bn']['var'] * np.float32(MOMENTUM) + var * np.float32(1 - MOMENTUM)).astype(np.float32)


def disc_step(params, g, e, dz, dx, batch_z, batch_x, eps_z, eps_x, noise):
    """bgm/base.py:190-245 -> ((dz_loss, dx_loss, d_loss), grads [dz..., dx...], bn stats)."""
    gt = g_to_t(g, False)
    et = to_t(e)
    dzt, dxt = disc_to_t(dz), disc_to_t(dx)
    z = torch.tensor(batch_z, dtype=torch.float32)
    x = torch.tensor(batch_x, dtype=torch.float32)
    stats = []
    z_ = mlp(et, x)                                                        # :203
    z_hat = (z * eps_z + z_ * (1 - eps_z)).requires_grad_(True)            # :204
    dz_hat = disc(dzt, z_hat)
    mu_x, s2_x = var_net_train(gt, z, stats)                               # :207
    x_ = torch.tensor(noise, dtype=torch.float32) * torch.sqrt(s2_x) + mu_x    # :208
    x_hat = (x * eps_x + x_ * (1 - eps_x)).requires_grad_(True)            # :209
    dx_hat = disc(dxt, x_hat)
    d_x_, d_z_, d_x, d_z = disc(dxt, x_), disc(dzt, z_), disc(dxt, x), disc(dzt, z)
    dz_loss = (((0.9 - d_z) ** 2).mean() + ((0.1 - d_z_) ** 2).mean()) / 2.0     # :221
    dx_loss = (((0.9 - d_x) ** 2).mean() + ((0.1 - d_x_) ** 2).mean()) / 2.0     # :223
    gz = torch.autograd.grad(dz_hat.sum(), z_hat, create_graph=True)[0]
    gpz = ((torch.sqrt((gz ** 2).sum(dim=1)) - 1.0) ** 2).mean()           # :227-229
    gx = torch.autograd.grad(dx_hat.sum(), x_hat, create_graph=True)[0]
    gpx = ((torch.sqrt((gx ** 2).sum(dim=1)) - 1.0) ** 2).mean()           # :232-234
    d_loss = dx_loss + dz_loss + params['gamma'] * (gpz + gpx)             # :236
    plist = disc_param_list(dzt) + disc_param_list(dxt)
    grads = torch.autograd.grad(d_loss, plist, allow_unused=True)
    grads = [np.zeros(tuple(p.shape), np.float32) if a is None else a.numpy() for a, p in zip(grads, plist)]
    return (float(dz_loss.detach()), float(dx_loss.detach()), float(d_loss.detach())), grads, stats


def gen_step(params, g, e, dz, dx, batch_z, batch_x, noise1, noise2):
    """bgm/base.py:247-291 -> ((g_loss_adv, e_loss_adv, l2_loss_z, l2_loss_x, reg_loss, g_e_loss),
    grads [g (Keras order)..., e...], bn stats of the two g_net calls)."""
    gt = g_to_t(g, True)
    et = to_t(e, True)
    dzt, dxt = disc_to_t(dz, False), disc_to_t(dx, False)
    z = torch.tensor(batch_z, dtype=torch.float32)
    x = torch.tensor(batch_x, dtype=torch.float32)
    stats = []
    mu1, s1 = var_net_train(gt, z, stats)                                  # :258
    x_ = torch.tensor(noise1, dtype=torch.float32) * torch.sqrt(s1) + mu1  # :259
    reg = (s1 ** 2).mean()                                                 # :260
    z_ = mlp(et, x)                                                        # :262
    z__ = mlp(et, x_)                                                      # :264
    mu2, s2 = var_net_train(gt, z_, stats)                                 # :266
    x__ = torch.tensor(noise2, dtype=torch.float32) * torch.sqrt(s2) + mu2     # :267
    d_x_, d_z_ = disc(dxt, x_), disc(dzt, z_)                              # :269-270
    l2_x = ((x - x__) ** 2).mean()                                         # :272
    l2_z = ((z - z__) ** 2).mean()                                         # :273
    g_adv = ((0.9 - d_x_) ** 2).mean()                                     # :277
    e_adv = ((0.9 - d_z_) ** 2).mean()                                     # :278
    loss = g_adv + e_adv + 10 * (l2_x + l2_z) + params['alpha'] * reg      # :279
    plist = g_param_list(gt) + mlp_param_list(et)
    grads = [a.numpy() for a in torch.autograd.grad(loss, plist)]
    losses = tuple(float(a.detach()) for a in (g_adv, e_adv, l2_z, l2_x, reg, loss))
    return losses, grads, stats


# ----------------------------------------------------------------------------------------
# Iterative phase of BGM.fit (bgm/base.py:145-187 and the loop :397-415)
def iter_g_grads(g, batch_z, batch_x):
    """update_g_net (:145-164): returns (loss_x, loss_mse, grads in g_param_list order, BN batch stats).
    The generator runs in training mode (`self.g_net(data_z)`, default training=True)."""
    gt = g_to_t(g, True)
    z = torch.tensor(batch_z, dtype=torch.float32)
    x = torch.tensor(batch_x, dtype=torch.float32)
    stats = []
    mu, s2 = var_net_train(gt, z, stats)
    loss_mse = ((x - mu) ** 2).mean()
    loss_x = (((x - mu) ** 2) / (2 * s2) + 0.5 * torch.log(s2)).sum(dim=1).mean()
    grads = torch.autograd.grad(loss_x, g_param_list(gt))
    return float(loss_x.detach()), float(loss_mse.detach()), [a.numpy() for a in grads], stats


def iter_latent_grad(g, batch_z, batch_x):
    """update_latent_variable_sgd (:167-187): returns (loss_postrior_z, d loss / d batch_z, BN stats)."""
    gt = g_to_t(g, False)
    z = torch.tensor(batch_z, dtype=torch.float32, requires_grad=True)
    x = torch.tensor(batch_x, dtype=torch.float32)
    stats = []
    mu, s2 = var_net_train(gt, z, stats)
    loss = (((x - mu) ** 2) / (2 * s2) + 0.5 * torch.log(s2)).sum(dim=1).mean() + ((z ** 2).sum(dim=1) / 2).mean()
    (gz,) = torch.autograd.grad(loss, [z])
    return float(loss.detach()), gz.numpy(), stats


def var_net_infer(g, z):
    """BaseVariationalNet.call(training=False): BN with the moving statistics; NumPy float32."""
    f32 = np.float32
    h = (z - g['bn']['mean']) / np.sqrt(g['bn']['var'] + f32(BN_EPS)) * g['bn']['gamma'] + g['bn']['beta']
    for W, b in g['hidden']:
        h = h @ W + b
        h = np.where(h > 0, h, f32(0.2) * h).astype(f32)
    mean = h @ g['mean'][0] + g['mean'][1]
    raw = h @ g['var'][0] + g['var'][1]
    return mean.astype(f32), (np.logaddexp(0, raw) + f32(1e-6)).astype(f32)


class BgmIterTrainer(object):
    """The iterative phase of BGM.fit on host arrays: Keras Adam(lr_theta, .9, .99) on the generator,
    and for the latent rows Adam(lr_z, .9, .99) applied to a FRESH variable per batch (zero slots,
    shared step count) followed by scatter_nd_update (:410-413)."""

    def __init__(self, params, g, data_z):
        from .train import Adam
        self.params, self.g = params, g
        self.data_z = np.array(data_z, np.float32)
        self.g_opt = Adam(params['lr_theta'], 0.9, 0.99)
        self.lr_z, self.t_z = params['lr_z'], 0

    def step(self, data, batch_idx):
        bz, bx = self.data_z[batch_idx].copy(), data[batch_idx]
        loss_x, mse, grads, stats = iter_g_grads(self.g, bz, bx)
        flat = g_flat_params(self.g)
        self.g_opt.apply(flat, grads)
        # write the updated arrays back into the dict (Adam updates in place, keep references in sync)
        self.g['bn']['gamma'], self.g['bn']['beta'] = flat[0], flat[1]
        update_moving(self.g, stats)
        lz, gz, stats = iter_latent_grad(self.g, bz, bx)
        update_moving(self.g, stats)
        f32 = np.float32
        self.t_z += 1
        lr_t = f32(self.lr_z * np.sqrt(1.0 - 0.99 ** self.t_z) / (1.0 - 0.9 ** self.t_z))
        m, v = f32(0.1) * gz, f32(1 - 0.99) * gz * gz
        self.data_z[batch_idx] = bz - lr_t * m / (np.sqrt(v) + f32(1e-7))
        return (loss_x, mse), lz, gz

    def epoch(self, data, batch_size):
        n = len(data)
        sample_idx = np.random.choice(n, n, replace=False)                 # :397
        last = None
        for i in range(0, n - batch_size + 1, batch_size):                 # incomplete last batch skipped (:401)
            last = self.step(data, sample_idx[i:i + batch_size])
        return last

    def mse(self, data):
        mu, _ = var_net_infer(self.g, self.data_z)
        return float(np.mean((data - mu) ** 2))
