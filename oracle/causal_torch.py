"""Multi-threaded torch-CPU restatement of the reference's RW-MH loop (TEST / BASELINE
INFRASTRUCTURE, see oracle/__init__.py): the CPU timing arm BASELINE.md section 3 describes.

Same structure as `causalbgm/base.py:860-898`: NumPy legacy-RNG proposal
`normal(0, q_sd, (n, zd)).astype(float32)` (:862), TWO fp32 log-posterior forward passes per
iteration (:865-866; `get_log_posterior` :765-817 with the MLPs of networks/base.py:30-51) as
torch-CPU GEMMs on all host threads, `exp(min(d, 0))`, `rand(n) <` accept (:868-870), masked copy
(:871), `current.copy()` per kept sample (:896).  `oracle/causal.py` (NumPy) stays the CHECKER;
tests/test_oracle.py pins this restatement to it.
"""
import numpy as np
import torch
import torch.nn.functional as F


def to_torch_nets(nets):
    return {k: [(torch.from_numpy(np.ascontiguousarray(W)), torch.from_numpy(np.ascontiguousarray(b))) for W, b in v]
            for k, v in nets.items()}


def mlp(layers, h):
    for W, b in layers[:-1]:
        h = F.leaky_relu(torch.addmm(b, h, W), 0.2, inplace=True)
    W, b = layers[-1]
    return torch.addmm(b, h, W)


def log_posterior(params, tnets, x, y, v, z, eps=1e-6):
    """causalbgm/base.py:765-817 (float32, torch CPU)."""
    d0, d1, d2, _ = params['z_dims']
    p = params['v_dim']
    z0, z1, z2 = z[:, :d0], z[:, d0:d0 + d1], z[:, d0 + d1:d0 + d1 + d2]
    g = mlp(tnets['g'], z)
    s2v = torch.tensor(params['sigma_v'] ** 2) if 'sigma_v' in params else F.softplus(g[:, -1]) + eps
    h = mlp(tnets['h'], torch.cat([z0, z2], dim=-1))
    s2x = torch.tensor(params['sigma_x'] ** 2) if 'sigma_x' in params else F.softplus(h[:, -1]) + eps
    f = mlp(tnets['f'], torch.cat([z0, z1, x], dim=-1))
    s2y = torch.tensor(params['sigma_y'] ** 2) if 'sigma_y' in params else F.softplus(f[:, -1]) + eps
    loss_pv = ((v - g[:, :p]) ** 2).sum(dim=1) / (2 * s2v) + p * torch.log(s2v) / 2
    if params['binary_treatment']:
        loss_px = F.binary_cross_entropy_with_logits(h[:, 0], x[:, 0], reduction='none')
    else:
        loss_px = ((x - h[:, :1]) ** 2).sum(dim=1) / (2 * s2x) + torch.log(s2x) / 2
    loss_py = ((y - f[:, :1]) ** 2).sum(dim=1) / (2 * s2y) + torch.log(s2y) / 2
    loss_prior = (z ** 2).sum(dim=1) / 2
    return -(loss_pv + loss_px + loss_py + loss_prior)


def mh_sampler(params, nets, data, q_sd=1.0, burn_in=0, n_keep=5, rs=None):
    """causalbgm/base.py:820-904 with a fixed proposal scale; returns (n_keep, n, zd) float32."""
    rs = rs if rs is not None else np.random
    tn = to_torch_nets(nets)
    x, y, v = [torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)) for a in data]
    n = x.shape[0]
    zd = sum(params['z_dims'])
    cur = torch.from_numpy(rs.normal(0, 1, size=(n, zd)).astype('float32'))                       # :842
    samples = []
    t = 0
    with torch.no_grad():
        while len(samples) < n_keep:
            prop = cur + torch.from_numpy(rs.normal(0, q_sd, size=(n, zd)).astype('float32'))       # :862
            lp_p = log_posterior(params, tn, x, y, v, prop)                                         # :865
            lp_c = log_posterior(params, tn, x, y, v, cur)                                          # :866
            ratio = torch.exp(torch.clamp(lp_p - lp_c, max=0.0))                                    # :868
            idx = torch.from_numpy(rs.rand(n)) < ratio.double()                                     # :870
            cur[idx] = prop[idx]                                                                    # :871
            if t >= burn_in:
                samples.append(cur.clone())                                                         # :896
            t += 1
    return torch.stack(samples).numpy()
