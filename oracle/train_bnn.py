"""Oracle restatement of CausalBGM's training steps on BAYESIAN networks (test infrastructure, see
oracle/__init__.py): torch autograd on the CPU over the DenseFlipout / batch-statistics BatchNorm
forward of oracle/bnn.py, with the network noise regenerated from the Philox streams of the CUDA
kernels (`oracle.bnn.PhiloxFlipout`), so that losses and gradients can be compared value for value
with the hand-derived backward of csrc/layered.cuh.

Follows `src/bayesgm/models/causalbgm/base.py`: train_disc_step :305-330, train_gen_step :332-377,
update_g/h/f_net :156-243 (with the `kl_weight * sum(net.losses)` term of :171-173, :205-207,
:234-236), update_latent_variable_sgd :246-302, evaluate :534-556; networks/bnn.py:4-38.

Noise ids: a training step with counter c draws call ids 16*c + k per net; signs are keyed by the row's
position in the batch (evaluate: by the data row).  Which call is which:
  gen step     g: 0 = g(z) for v_ (:336), 1 = g(z) again for its sigma column (:337), 2 = g(z_) (:345)
               e: 0 = e(v) (:338), 1 = e(v_) (:344);  f, h: 0 = mean column, 1 = sigma column (:354-359)
  disc step    e: 0
  update_*     g, h, f: 0
  latent step  g, h, f: 0 = mean, 1 = sigma (:259-286)
  evaluate     e, g, h, f: 0
"""
import numpy as np
import torch

from .bnn import PhiloxFlipout, SCALE_EPS
from .train import disc, disc_to_t, disc_param_list, BN_EPS


def net_to_t(net, requires_grad=True):
    t = lambda a: torch.tensor(np.asarray(a), dtype=torch.float32, requires_grad=requires_grad)
    return dict(gamma=t(net['bn']['gamma']), beta=t(net['bn']['beta']),
                layers=[(t(loc), t(rho), t(b)) for loc, rho, b in net['layers']])


def param_list(nt):
    """Keras trainable_variables order: BN gamma, beta; then loc, rho, bias per DenseFlipout layer."""
    out = [nt['gamma'], nt['beta']]
    for loc, rho, b in nt['layers']:
        out += [loc, rho, b]
    return out


def forward(nt, x, prov, name, call, rows_offset=0, stats=None):
    """BayesianFullyConnectedNet.call (bnn.py:25-38), differentiable."""
    if stats is None:
        mu = x.mean(dim=0)
        var = ((x - mu) ** 2).mean(dim=0)
    else:
        mu, var = stats
    h = (x - mu) / torch.sqrt(var + BN_EPS) * nt['gamma'] + nt['beta']
    L = len(nt['layers'])
    rows = x.shape[0]
    for l, (loc, rho, b) in enumerate(nt['layers']):
        K, N = loc.shape
        eps, s_in, s_out = prov.flipout(name, l, call, rows, K, N)
        sigma = float(SCALE_EPS) + torch.nn.functional.softplus(rho)
        dW = sigma * torch.from_numpy(eps)
        out = h @ loc + ((h * torch.from_numpy(s_in)) @ dW) * torch.from_numpy(s_out) + b
        h = torch.nn.functional.leaky_relu(out, 0.2) if l < L - 1 else out
    return h


def kl(nt):
    tot = 0.0
    for loc, rho, _ in nt['layers']:
        s = float(SCALE_EPS) + torch.nn.functional.softplus(rho)
        tot = tot + (-torch.log(s) + 0.5 * (s * s + loc * loc) - 0.5).sum()
    return tot


def _grads(loss, plist):
    gs = torch.autograd.grad(loss, plist, allow_unused=True, retain_graph=True)
    return [np.zeros(tuple(p.shape), np.float32) if g is None else g.detach().numpy() for g, p in zip(gs, plist)]


def _split_z(params, z):
    d0, d1, d2, _ = params['z_dims']
    return z[:, :d0], z[:, d0:d0 + d1], z[:, d0 + d1:d0 + d1 + d2]


def encode(params, nets, batch_v, seed, ctr):
    """e_net(v) as train_disc_step draws it (:310): call id 16*ctr."""
    prov = PhiloxFlipout(seed)
    with torch.no_grad():
        return forward(net_to_t(nets['e'], False), torch.tensor(batch_v, dtype=torch.float32), prov, 'e', 16 * ctr).numpy()


def disc_step(params, nets, dz, batch_z, batch_v, epsilon, seed, ctr):
    """causalbgm/base.py:305-330 with a Bayesian e_net: (dz_loss, d_loss, dz_net gradients)."""
    pt = disc_to_t(dz)
    z = torch.tensor(batch_z, dtype=torch.float32)
    z_ = torch.tensor(encode(params, nets, batch_v, seed, ctr))
    z_hat = (z * epsilon + z_ * (1 - epsilon)).requires_grad_(True)
    d_hat, d_, d = disc(pt, z_hat), disc(pt, z_), disc(pt, z)
    dz_loss = -d.mean() + d_.mean()
    grad_z = torch.autograd.grad(d_hat.sum(), z_hat, create_graph=True)[0]
    gp = ((torch.sqrt((grad_z ** 2).sum(dim=1)) - 1.0) ** 2).mean()
    d_loss = dz_loss + 10 * gp
    return float(dz_loss.detach()), float(d_loss.detach()), _grads(d_loss, disc_param_list(pt))


def gen_step(params, nets, dz, batch_z, batch_v, batch_x, batch_y, seed, ctr):
    """causalbgm/base.py:332-377 on Bayesian nets -> (losses as :377, gradients per net in Keras order)."""
    p = params['v_dim']
    prov = PhiloxFlipout(seed)
    c = 16 * ctr
    g, e, f, h = [net_to_t(nets[k]) for k in ('g', 'e', 'f', 'h')]
    pt = disc_to_t(dz, requires_grad=False)
    z = torch.tensor(batch_z, dtype=torch.float32)
    v = torch.tensor(batch_v, dtype=torch.float32)
    x = torch.tensor(batch_x, dtype=torch.float32).reshape(-1, 1)
    y = torch.tensor(batch_y, dtype=torch.float32).reshape(-1, 1)
    v_ = forward(g, z, prov, 'g', c)[:, :p]                                  # :336
    sig = (forward(g, z, prov, 'g', c + 1)[:, -1] ** 2).mean()               # :337
    z_ = forward(e, v, prov, 'e', c)                                         # :338
    z0, z1, z2 = _split_z(params, z_)
    z__ = forward(e, v_, prov, 'e', c + 1)                                   # :344
    v__ = forward(g, z_, prov, 'g', c + 2)[:, :p]                            # :345
    d_ = disc(pt, z_)                                                        # :347
    l2_v = ((v - v__) ** 2).mean()
    l2_z = ((z - z__) ** 2).mean()
    e_adv = -d_.mean()
    fin = torch.cat([z0, z1, x], dim=-1)
    hin = torch.cat([z0, z2], dim=-1)
    y_ = forward(f, fin, prov, 'f', c)[:, :1]                                # :354
    sig = sig + (forward(f, fin, prov, 'f', c + 1)[:, -1] ** 2).mean()       # :355-356
    x_ = forward(h, hin, prov, 'h', c)[:, :1]                                # :357
    sig = sig + (forward(h, hin, prov, 'h', c + 1)[:, -1] ** 2).mean()       # :358-359
    if params['binary_treatment']:
        l2_x = torch.nn.functional.binary_cross_entropy_with_logits(x_, x)
    else:
        l2_x = ((x_ - x) ** 2).mean()
    l2_y = ((y_ - y) ** 2).mean()
    total = e_adv + (l2_v + float(params['use_z_rec']) * l2_z) + (l2_x + l2_y) + 0.001 * sig   # :367-368
    grads = {k: _grads(total, param_list(n)) for k, n in (('g', g), ('e', e), ('f', f), ('h', h))}
    losses = tuple(float(a.detach()) for a in (e_adv, l2_v, l2_z, l2_x, l2_y, total))
    return losses, grads


def _nll(target, mu, raw, fixed_sigma, eps=1e-6):
    s2 = torch.tensor(fixed_sigma ** 2) if fixed_sigma is not None else torch.nn.functional.softplus(raw) + eps
    D = target.shape[1]
    return (((target - mu) ** 2).sum(dim=1) / (2 * s2) + D * torch.log(s2) / 2).mean()


def iter_nets_step(params, nets, batch_z, batch_x, batch_y, batch_v, seed, ctr):
    """update_g_net, update_h_net, update_f_net (:156-243) -> (losses[6], gradients of g, h, f)."""
    p = params['v_dim']
    prov = PhiloxFlipout(seed)
    c = 16 * ctr
    klw = float(params['kl_weight'])
    z = torch.tensor(batch_z, dtype=torch.float32)
    v = torch.tensor(batch_v, dtype=torch.float32)
    x = torch.tensor(batch_x, dtype=torch.float32).reshape(-1, 1)
    y = torch.tensor(batch_y, dtype=torch.float32).reshape(-1, 1)
    z0, z1, z2 = _split_z(params, z)
    g, h, f = [net_to_t(nets[k]) for k in ('g', 'h', 'f')]
    go = forward(g, z, prov, 'g', c)
    loss_v = _nll(v, go[:, :p], go[:, -1], params.get('sigma_v')) + klw * kl(g)
    mse_v = ((v - go[:, :p]) ** 2).mean()
    ho = forward(h, torch.cat([z0, z2], dim=-1), prov, 'h', c)
    if params['binary_treatment']:
        ce = torch.nn.functional.binary_cross_entropy_with_logits(ho[:, :1], x)
        loss_x, mse_x = ce + klw * kl(h), ce
    else:
        loss_x = _nll(x, ho[:, :1], ho[:, -1], params.get('sigma_x')) + klw * kl(h)
        mse_x = ((x - ho[:, :1]) ** 2).mean()
    fo = forward(f, torch.cat([z0, z1, x], dim=-1), prov, 'f', c)
    loss_y = _nll(y, fo[:, :1], fo[:, -1], params.get('sigma_y')) + klw * kl(f)
    mse_y = ((y - fo[:, :1]) ** 2).mean()
    grads = dict(g=_grads(loss_v, param_list(g)), h=_grads(loss_x, param_list(h)), f=_grads(loss_y, param_list(f)))
    losses = tuple(float(a.detach()) for a in (loss_v, mse_v, loss_x, mse_x, loss_y, mse_y))
    return losses, grads


def latent_step(params, nets, batch_z, batch_x, batch_y, batch_v, seed, ctr):
    """update_latent_variable_sgd (:246-302) -> (loss_postrior_z, d loss / d batch rows)."""
    p = params['v_dim']
    prov = PhiloxFlipout(seed)
    c = 16 * ctr
    z = torch.tensor(batch_z, dtype=torch.float32, requires_grad=True)
    v = torch.tensor(batch_v, dtype=torch.float32)
    x = torch.tensor(batch_x, dtype=torch.float32).reshape(-1, 1)
    y = torch.tensor(batch_y, dtype=torch.float32).reshape(-1, 1)
    z0, z1, z2 = _split_z(params, z)
    g, h, f = [net_to_t(nets[k], False) for k in ('g', 'h', 'f')]
    mu_v = forward(g, z, prov, 'g', c)[:, :p]                                # :259
    raw_v = forward(g, z, prov, 'g', c + 1)[:, -1]                           # :263
    loss = _nll(v, mu_v, raw_v, params.get('sigma_v'))
    hin = torch.cat([z0, z2], dim=-1)
    mu_x = forward(h, hin, prov, 'h', c)[:, :1]
    raw_x = forward(h, hin, prov, 'h', c + 1)[:, -1]
    if params['binary_treatment']:
        loss = loss + torch.nn.functional.binary_cross_entropy_with_logits(mu_x, x)
    else:
        loss = loss + _nll(x, mu_x, raw_x, params.get('sigma_x'))
    fin = torch.cat([z0, z1, x], dim=-1)
    mu_y = forward(f, fin, prov, 'f', c)[:, :1]
    raw_y = forward(f, fin, prov, 'f', c + 1)[:, -1]
    loss = loss + _nll(y, mu_y, raw_y, params.get('sigma_y'))
    loss = loss + ((z ** 2).sum(dim=1) / 2).mean()                           # :291-292
    gz = torch.autograd.grad(loss, z)[0]
    return float(loss.detach()), gz.numpy()


def evaluate_mse(params, nets, data, seed, ctr, data_z=None):
    """The per-row part of evaluate (:534-556): (mse_x, mse_y, mse_v, z used)."""
    x, y, v = [torch.tensor(np.asarray(a), dtype=torch.float32) for a in data]
    x, y = x.reshape(-1, 1), y.reshape(-1, 1)
    p = params['v_dim']
    prov = PhiloxFlipout(seed)
    c = 16 * ctr
    with torch.no_grad():
        e, g, h, f = [net_to_t(nets[k], False) for k in ('e', 'g', 'h', 'f')]
        z = forward(e, v, prov, 'e', c) if data_z is None else torch.tensor(data_z, dtype=torch.float32)
        z0, z1, z2 = _split_z(params, z)
        v_pred = forward(g, z, prov, 'g', c)[:, :p]
        x_pred = forward(h, torch.cat([z0, z2], dim=-1), prov, 'h', c)[:, :1]
        y_pred = forward(f, torch.cat([z0, z1, x], dim=-1), prov, 'f', c)[:, :1]
        if params['binary_treatment']:
            x_pred = torch.sigmoid(x_pred)
        return (float(((x - x_pred) ** 2).mean()), float(((y - y_pred) ** 2).mean()), float(((v - v_pred) ** 2).mean()),
                z.numpy())
