"""BGM with the HMC posterior-sampling path on B200 (sm_100a) kernels.

Drop-in for the method surface of `bayesgm.models.bgm.BGM`
(`src/bayesgm/models/bgm/base.py`): same constructor, method names, kwargs, return
shapes and error behaviour for the path this package accelerates --
`get_log_posterior` (:665), `tfp_mcmc_sampler` (:709), `predict_on_posteriors` (:511),
`predict` (:527) and `generate` (:483).  The HMC integrator / accept step / shared
step-size adaptation the reference delegates to TFP 0.18 runs in `hmc_kernel`
(csrc/hmc.cuh); inputs and outputs are host NumPy arrays like the reference's.
There is no CPU fallback.  Deterministic generator only (`use_bnn=False`).
"""
import ctypes as C
import datetime
import os

import numpy as np

from . import _lib
from .datasets import Gaussian_sampler
from .nets import DenseNet, VariationalNet, DiscNet
from .datasets import Base_sampler

_DEFAULTS = dict(use_bnn=False, g_units=[64] * 5, e_units=[64] * 5, dz_units=[64, 32, 8],
                 dx_units=[64, 32, 8], lr=0.001, lr_theta=0.005, lr_z=0.005, gamma=0.0, alpha=0.0,
                 g_d_freq=1, save_model=True, save_res=True, kl_weight=0.00005)


def quantile_dim0(torch, a, q):
    """np.quantile(a, q, axis=0) (linear interpolation) on the device."""
    srt = torch.sort(a, dim=0).values
    pos = q * (a.shape[0] - 1)
    lo = int(np.floor(pos))
    hi = min(lo + 1, a.shape[0] - 1)
    return srt[lo] + (srt[hi] - srt[lo]) * float(pos - lo)


def nan_coded(data, ind_x1, n, x_dim, with_extra=False):
    """The dense equivalent of tfp_mcmc_sampler's `ind_x1` forms (bgm/base.py:741-775):
    a copy of `data` with NaN at every entry that is NOT listed as observed.

    A feature may be listed several times for a row (legal in the reference: the gathered terms of
    :689-700 simply add).  With `with_extra=True` the function returns (data', extra_cols): data' has one
    more column per (feature, repeat level) that occurs, holding the value where the row lists the feature
    at least that often and NaN elsewhere; the caller evaluates it with a generator whose heads repeat those
    output columns (VariationalNet.with_extra_columns).  Without it duplicates raise."""
    data = np.array(data, dtype=np.float32, copy=True)
    if ind_x1 is None:
        return (data, []) if with_extra else data
    counts = np.zeros((n, x_dim), dtype=np.int32)
    if isinstance(ind_x1, (list, tuple)) and len(ind_x1) > 0 and isinstance(ind_x1[0], (list, tuple)):
        assert len(ind_x1) == n, "len(ind_x1)=%d != n_samples=%d" % (len(ind_x1), n)
        assert max(len(r) for r in ind_x1) > 0, "No observed features"
        for i, row in enumerate(ind_x1):
            np.add.at(counts[i], np.asarray(row, dtype=np.int64), 1)
    else:
        ind = np.asarray(ind_x1, dtype=np.int64)
        if ind.ndim == 1:
            ind = np.broadcast_to(ind[None, :], (n, ind.shape[0]))
        elif ind.ndim != 2:
            raise ValueError("ind_x1 must be rank 1 or 2 if tensor-like.")
        np.add.at(counts, (np.repeat(np.arange(n), ind.shape[1]), ind.reshape(-1)), 1)
    extra_cols, extra_data = [], []
    if counts.max() > 1:
        if not with_extra:
            raise NotImplementedError("bayesgm_b200: duplicate feature indices in ind_x1 need the caller to pass "
                                      "with_extra=True and an extended generator")
        for level in range(2, int(counts.max()) + 1):
            for j in np.nonzero((counts >= level).any(axis=0))[0]:
                extra_cols.append(int(j))
                extra_data.append(np.where(counts[:, j] >= level, data[:, j], np.float32(np.nan)))
    data[counts == 0] = np.nan
    if with_extra:
        if extra_cols:
            data = np.concatenate([data, np.stack(extra_data, axis=1).astype(np.float32)], axis=1)
        return data, extra_cols
    return data


class BGM(object):
    """See the reference docstring, bgm/base.py:19-57, for `params`."""

    def __init__(self, params, timestamp=None, random_seed=None):
        self.params = params
        self.timestamp = timestamp
        p = dict(_DEFAULTS)
        p.update(params)
        self._p = p
        if p['use_bnn']:
            raise NotImplementedError(
                "bayesgm_b200: use_bnn=True (BayesianVariationalNet, networks/bnn.py) is not built; "
                "pass use_bnn=False (BaseVariationalNet, networks/base.py).")
        rng = np.random.RandomState(random_seed) if random_seed is not None else np.random
        self.g_net = VariationalNet(p['z_dim'], p['x_dim'], 'g_net', p['g_units'], rng)      # :70
        self.e_net = DenseNet(p['x_dim'], p['z_dim'], 'e_net', p['e_units'], rng)             # :73
        self.dz_net = DiscNet(p['z_dim'], 'dz_net', p['dz_units'], rng)                       # :76
        self.dx_net = DiscNet(p['x_dim'], 'dx_net', p['dx_units'], rng)                       # :78
        self.z_sampler = Gaussian_sampler(mean=np.zeros(p['z_dim']), sd=1.0)                 # :85
        self._trainer = None
        self._trainer_dirty = False
        self._layered = False
        self._noise_rng = np.random.RandomState(0 if random_seed is None else random_seed)
        if self.timestamp is None:
            self.timestamp = datetime.datetime.now().strftime('%Y%m%d_%H%M%S')
        self.checkpoint_path = "{}/checkpoints/{}/{}".format(p['output_dir'], p['dataset'], self.timestamp)
        if p['save_model'] and not os.path.exists(self.checkpoint_path):
            os.makedirs(self.checkpoint_path)
        self.save_dir = "{}/results/{}/{}".format(p['output_dir'], p['dataset'], self.timestamp)
        if p['save_res'] and not os.path.exists(self.save_dir):
            os.makedirs(self.save_dir)
        self._handle = None
        self.last_acceptance_rate = None
        self.last_step_size = None

    # ------------------------------------------------------------------ plumbing
    def get_config(self):
        return {"params": self.params}

    def initialize_nets(self, print_summary=False):
        if print_summary:
            print(self.g_net.model_name, [self.g_net.input_dim] + self.g_net.nb_units + [self.g_net.output_dim])

    def set_weights(self, g=None, e=None, dz=None, dx=None):
        """Keras-layout weights (`net.get_weights()` of a trained reference model; dz / dx: the
        Discriminators' trainable_variables)."""
        self._sync_from_trainer()
        if g is not None:
            self.g_net.set_weights(g)
        if e is not None:
            self.e_net.set_weights(e)
        if dz is not None:
            self.dz_net.set_trainable(dz)
        if dx is not None:
            self.dx_net.set_trainable(dx)
        self._drop_handle()
        self._drop_trainer()

    def get_weights(self):
        self._sync_from_trainer()
        return dict(g=self.g_net.get_weights(), e=self.e_net.get_weights(),
                    dz=[a.copy() for a in self.dz_net.trainable_list()],
                    dx=[a.copy() for a in self.dx_net.trainable_list()])

    def save_weights(self, path):
        """All network weights (Keras-layout arrays) into one .npz -- stands in for the reference's
        `*.weights.h5` files (:335-338, :433-436); `load_weights` restores them."""
        w = self.get_weights()
        np.savez(path, **{"%s_%d" % (k, i): a for k, arrs in w.items() for i, a in enumerate(arrs)})

    def load_weights(self, path):
        z = np.load(path)
        got = {}
        for k in ('g', 'e', 'dz', 'dx'):
            got[k] = [z["%s_%d" % (k, i)] for i in range(sum(1 for name in z.files if name.startswith(k + "_")))]
        self.set_weights(**got)

    # Two training engines behind the same calls (same device parameter layout): the fused single-CTA kernels
    # (csrc/train_bgm.cuh: batch <= 32, x_dim up to ~110) and the layered engine (csrc/layered.cuh: any x_dim,
    # any batch size, gamma == 0).
    _LT_NAMES = dict(bgm_bgmtrainer_create="bgm_ltb_create", bgm_trainer_destroy="bgm_ltb_destroy",
                     bgm_trainer_buffers="bgm_ltb_buffers", bgm_trainer_get_params="bgm_ltb_get_params",
                     bgm_trainer_bn_moving="bgm_ltb_bn_moving", bgm_bgm_train_disc_grad="bgm_ltb_disc_grad",
                     bgm_bgm_train_gen_grad="bgm_ltb_gen_grad", bgm_train_adam="bgm_ltb_adam",
                     bgm_bgmtrainer_set_iter="bgm_ltb_set_iter", bgm_bgm_iter_g="bgm_ltb_iter_g",
                     bgm_bgm_iter_latent="bgm_ltb_iter_latent", bgm_bgm_evaluate="bgm_ltb_evaluate")

    def _tfn(self, name):
        """Symbol of the active training engine.  The engine is only known once the trainer exists (a model too
        wide for the fused kernels falls back to the layered engine at creation)."""
        if self._trainer is None and name not in ("bgm_trainer_destroy", "bgm_bgmtrainer_create"):
            self._device_trainer()
        return self._LT_NAMES[name] if self._layered else name

    def _set_layered(self, on):
        if bool(on) != self._layered:
            self._sync_from_trainer()
            self._drop_trainer()
            self._layered = bool(on)

    def _drop_trainer(self):
        if self._trainer is not None:
            getattr(_lib.load(), self._tfn("bgm_trainer_destroy"))(self._trainer)
            self._trainer = None
            self._trainer_dirty = False

    def _create_trainer(self, layered):
        p = self._p
        gd, gk = self.g_net.desc()
        ed, ek = self.e_net.desc()
        zd_, zk = self.dz_net.desc()
        xd_, xk = self.dx_net.desc()
        h = C.c_void_p()
        _lib.call("bgm_ltb_create" if layered else "bgm_bgmtrainer_create", C.byref(h), C.byref(gd), C.byref(ed),
                  C.byref(zd_), C.byref(xd_), float(p['lr']), 0.5, 0.9, float(p['alpha']), float(p['gamma']))   # Adam betas :83-85
        return h

    def _device_trainer(self):
        if self._trainer is None:
            _lib.require_cuda()
            if not self._layered:
                try:
                    self._trainer = self._create_trainer(False)
                except _lib.BgmError as e:
                    if e.code != -4:          # BGM_ERR_NOMEM: x_dim too wide for one SM's shared memory
                        raise
                    self._layered = True
            if self._trainer is None:
                self._trainer = self._create_trainer(True)
        return self._trainer

    def _sync_from_trainer(self):
        """Pull trained parameters (device layout: gamma | beta | hidden | [W_mean|W_var] | [b_mean|b_var] | e)
        and the BN moving statistics back into the Keras-layout host arrays."""
        if self._trainer is None or not self._trainer_dirty:
            return
        n = C.c_int()
        zd, xd = self._p['z_dim'], self._p['x_dim']
        _lib.call(self._tfn("bgm_trainer_buffers"), self._trainer, 0, C.byref(n), None, None)
        flat = np.empty(n.value, np.float32)
        _lib.call(self._tfn("bgm_trainer_get_params"), self._trainer, 0, flat.ctypes.data_as(C.c_void_p))
        g = self.g_net
        g.bn['gamma'], g.bn['beta'] = flat[:zd].copy(), flat[zd:2 * zd].copy()
        o = 2 * zd
        for layer in g.hidden:
            for i in (0, 1):
                a = layer[i]
                layer[i] = flat[o:o + a.size].reshape(a.shape).copy()
                o += a.size
        last = g.nb_units[-1]
        wcat = flat[o:o + last * 2 * xd].reshape(last, 2 * xd)
        o += last * 2 * xd
        bcat = flat[o:o + 2 * xd]
        o += 2 * xd
        g.mean = [wcat[:, :xd].copy(), bcat[:xd].copy()]
        g.var = [wcat[:, xd:].copy(), bcat[xd:].copy()]
        self.e_net.load_flat(flat[o:])
        mv = np.empty(2 * zd, np.float32)
        _lib.call(self._tfn("bgm_trainer_bn_moving"), self._trainer, mv.ctypes.data_as(C.c_void_p), 0)
        g.bn['mean'], g.bn['var'] = mv[:zd].copy(), mv[zd:].copy()
        _lib.call(self._tfn("bgm_trainer_buffers"), self._trainer, 1, C.byref(n), None, None)
        flat = np.empty(n.value, np.float32)
        _lib.call(self._tfn("bgm_trainer_get_params"), self._trainer, 1, flat.ctypes.data_as(C.c_void_p))
        k = self.dz_net.flat_params().size
        self.dz_net.load_flat(flat[:k])
        self.dx_net.load_flat(flat[k:])
        self._trainer_dirty = False
        self._drop_handle()

    def _drop_handle(self):
        if self._handle is not None:
            _lib.load().bgm_hmc_destroy(self._handle)
            self._handle = None

    def __del__(self):
        try:
            self._drop_handle()
            self._drop_trainer()
        except Exception:
            pass

    def _extended_model(self, extra_cols):
        """A throw-away device model whose heads repeat `extra_cols` (duplicate indices in ind_x1); the caller
        destroys it with bgm_hmc_destroy."""
        self._sync_from_trainer()
        _lib.require_cuda()
        d, keep = self.g_net.with_extra_columns(extra_cols).desc()
        h = C.c_void_p()
        _lib.call("bgm_hmc_create", C.byref(h), C.byref(d))
        del keep
        if getattr(self, 'hmc_engine', 'auto') == 'simt':
            _lib.call("bgm_hmc_set_engine", h, 1)
        return h

    def _device_model(self):
        if getattr(self, '_handle_override', None) is not None:
            return self._handle_override
        self._sync_from_trainer()
        if self._handle is None:
            _lib.require_cuda()
            d, keep = self.g_net.desc()
            h = C.c_void_p()
            _lib.call("bgm_hmc_create", C.byref(h), C.byref(d))
            self._handle = h
            kind = {'auto': 0, 'simt': 1, 'tensor': 2}[getattr(self, 'hmc_engine', 'auto')]
            if kind:
                _lib.call("bgm_hmc_set_engine", h, kind)
        return self._handle

    def set_hmc_engine(self, engine='auto'):
        """'auto' (tensor-core engine when every hidden layer of g_net is 64 wide), 'simt' or 'tensor' (raises if
        unavailable).  Both run the same algorithm on the same Philox streams; DESIGN.md 4.3 / 4.3b."""
        if engine not in ('auto', 'simt', 'tensor'):
            raise ValueError("engine must be 'auto', 'simt' or 'tensor'")
        self.hmc_engine = engine
        if self._handle is not None:
            _lib.call("bgm_hmc_set_engine", self._handle, {'auto': 0, 'simt': 1, 'tensor': 2}[engine])

    def hmc_engine_info(self):
        kind, avail, smem = C.c_int(), C.c_int(), C.c_int()
        issued = C.c_longlong()
        _lib.call("bgm_hmc_engine_info", self._device_model(), C.byref(kind), C.byref(avail), C.byref(smem), C.byref(issued))
        return dict(engine={1: 'simt', 2: 'tensor'}[kind.value], tensor_available=bool(avail.value),
                    tensor_smem_bytes=smem.value, tensor_issued_macs_per_grad=issued.value)

    def kernel_info(self):
        smem, nops = C.c_int(), C.c_int()
        macs, issued = C.c_longlong(), C.c_longlong()
        _lib.call("bgm_hmc_info", self._device_model(), C.byref(smem), C.byref(nops), C.byref(macs),
                  C.byref(issued))
        return dict(smem_bytes=smem.value, n_ops=nops.value, macs_per_grad=macs.value,
                    issued_macs_per_grad=issued.value)

    def _stage_x(self, data, torch, n_extra=0):
        """Host (n,x_dim) array with NaN = missing -> device (n,ldx), ldx % 4 == 0,
        pad columns NaN (= not observed).  n_extra: virtual columns of nan_coded(with_extra=True)."""
        xd = self._p['x_dim'] + int(n_extra)
        if isinstance(data, torch.Tensor):
            t = data.float()
        else:
            t = torch.from_numpy(np.ascontiguousarray(data, dtype=np.float32))
        if t.ndim != 2 or t.shape[1] != xd:
            raise ValueError("data must have shape (n, %d)" % xd)
        ldx = (xd + 3) // 4 * 4
        d = t.contiguous() if t.is_cuda else t.contiguous().to('cuda', non_blocking=True)
        if ldx != xd:
            pad = torch.full((d.shape[0], ldx), float('nan'), dtype=torch.float32, device='cuda')
            pad[:, :xd] = d
            d = pad
        return d, ldx, d.shape[0]

    @staticmethod
    def _dev(a, torch, dtype=None):
        if isinstance(a, torch.Tensor):
            t = a
        else:
            t = torch.from_numpy(np.ascontiguousarray(a))
        if dtype is not None and t.dtype != dtype:
            t = t.to(dtype)
        return t.contiguous().to('cuda', non_blocking=True) if not t.is_cuda else t.contiguous()

    # --------------------------------------------------------------- hot path
    def get_log_posterior(self, data_z, data_x, ind_x1=None, obs_mask=None, *, return_grad=False):
        """bgm/base.py:665-705 -> (n,) float32.  Missing observations: either NaN entries in
        `data_x` or the reference's (`ind_x1`, `obs_mask`) padded-gather form (:689-700).
        `return_grad=True` also returns d log p / d z, the quantity TFP's HMC differentiates for."""
        torch = _lib.require_cuda()
        data_x = np.asarray(data_x, dtype=np.float32)
        n, xd = data_x.shape
        if ind_x1 is not None:
            ind = np.asarray(ind_x1, dtype=np.int64)
            msk = np.ones(ind.shape, bool) if obs_mask is None else np.asarray(obs_mask) > 0
            lists = [ind[i][msk[i]].tolist() for i in range(n)]
            data_x, extra = nan_coded(data_x, lists, n, xd, with_extra=True)
        else:
            extra = []
        x, ldx, n = self._stage_x(data_x, torch, len(extra))
        z = self._dev(data_z, torch, torch.float32)
        if z.shape != (n, self._p['z_dim']):
            raise ValueError("data_z must have shape (%d, %d)" % (n, self._p['z_dim']))
        lp = torch.empty(n, dtype=torch.float32, device='cuda')
        g = torch.empty((n, self._p['z_dim']), dtype=torch.float32, device='cuda') if return_grad else None
        model = self._extended_model(extra) if extra else self._device_model()
        try:
            _lib.call("bgm_hmc_logpost_grad", model, _lib.ptr(x), ldx, _lib.ptr(z), n, _lib.ptr(lp),
                      _lib.ptr(g), _lib.stream_ptr())
            res = (lp.cpu().numpy(), g.cpu().numpy()) if return_grad else lp.cpu().numpy()
        finally:
            if extra:
                _lib.load().bgm_hmc_destroy(model)
        return res

    def _hmc_device(self, x, ldx, n, n_mcmc, burn_in, step_size, num_leapfrog_steps, seed, row_offset=0,
                    noise=None, trace=False, group=None, n_total=None, keep_samples=True,
                    adaptation_rate=0.01, target_accept=0.75):
        """Runs the sampler on a staged device buffer; returns a dict of device tensors."""
        torch = _lib.require_cuda()
        zd = self._p['z_dim']
        T = int(burn_in) + int(n_mcmc)
        n_adapt = int(burn_in * 0.8)                                              # :807
        m = self._device_model()
        dev = 'cuda'
        z_state = torch.empty((n, zd), dtype=torch.float32, device=dev)
        g_state = torch.empty((n, zd), dtype=torch.float32, device=dev)
        lp_state = torch.empty(n, dtype=torch.float32, device=dev)
        samples = torch.empty((n_mcmc, n, zd), dtype=torch.float32, device=dev) if keep_samples else None
        stat = torch.zeros(max(T, 1), dtype=torch.float64, device=dev)
        count = torch.zeros(max(T, 1), dtype=torch.int32, device=dev)
        step = torch.tensor([step_size], dtype=torch.float32, device=dev)
        a = _lib.HmcArgs()
        a.x_dev, a.ldx, a.n = x.data_ptr(), ldx, n
        a.z_state_dev, a.g_state_dev, a.lp_state_dev = z_state.data_ptr(), g_state.data_ptr(), lp_state.data_ptr()
        a.burn_in, a.num_leapfrog = int(burn_in), int(num_leapfrog_steps)
        a.step_dev = step.data_ptr()
        a.seed, a.row_offset = int(seed) & (2 ** 64 - 1), int(row_offset)
        a.out_samples_dev = samples.data_ptr() if keep_samples else None
        a.accept_stat_dev, a.accept_count_dev = stat.data_ptr(), count.data_ptr()
        keep = []
        if noise is not None:
            z0 = self._dev(noise['z0'], torch, torch.float32)
            mom = self._dev(noise['momentum'], torch, torch.float32)
            logu = self._dev(noise['log_u'], torch, torch.float32)
            assert z0.shape == (n, zd) and mom.shape == (T, n, zd) and logu.shape == (T, n)
            z_state.copy_(z0)
            a.mom_dev, a.logu_dev = mom.data_ptr(), logu.data_ptr()
            a.init_mode = 1
            keep += [mom, logu]
        else:
            a.init_mode = 2
        out = dict(samples=samples, z_state=z_state, g_state=g_state, lp_state=lp_state, accept_stat=stat,
                   accept_count=count, step=step, _keep=keep)
        if trace:
            out['accept_mask'] = torch.zeros((T, n), dtype=torch.uint8, device=dev)
            out['log_accept'] = torch.zeros((T, n), dtype=torch.float32, device=dev)
            out['step_trace'] = torch.zeros(T, dtype=torch.float32, device=dev)
            a.accept_mask_dev, a.log_accept_dev = out['accept_mask'].data_ptr(), out['log_accept'].data_ptr()
        st = _lib.stream_ptr()
        n_total = int(n_total) if n_total is not None else n
        # one launch per step while the shared step size adapts (:805-809), then one launch
        t = 0
        while t < min(n_adapt, T):
            a.t_begin, a.t_end = t, t + 1
            if trace:
                out['step_trace'][t] = step[0]
            _lib.call("bgm_hmc_run", m, C.byref(a), st)
            a.init_mode = 0
            if group is not None:
                import torch.distributed as dist
                dist.all_reduce(stat[t:t + 1], group=group)
            _lib.call("bgm_hmc_adapt", C.c_void_p(stat.data_ptr()), t, n_total, float(target_accept),
                      float(adaptation_rate), C.c_void_p(step.data_ptr()), st)
            t += 1
        if t < T or a.init_mode != 0:
            a.t_begin, a.t_end = t, T
            if trace and t < T:
                out['step_trace'][t:] = step[0]
            _lib.call("bgm_hmc_run", m, C.byref(a), st)
        return out

    def tfp_mcmc_sampler(self, data, ind_x1=None, n_mcmc=3000, burn_in=5000, step_size=0.01,
                         num_leapfrog_steps=10, seed=42, *, noise=None, return_trace=False, verbose=1):
        """bgm/base.py:709-830 -> np.ndarray (n_mcmc, n, z_dim).

        `data` may carry NaN for missing entries (as BGM.predict receives it) and/or
        `ind_x1` may list the observed features per row (list of lists, (n,K) or (K,)
        indices).  Noise: in-kernel Philox4x32-10 keyed by `seed`, or
        `noise=dict(z0, momentum, log_u)` of shapes (n,zd), (T,n,zd), (T,n) for
        state-for-state parity tests.
        """
        torch = _lib.require_cuda()
        data = np.asarray(data, dtype=np.float32)
        n, xd = data.shape
        coded, extra = nan_coded(data, ind_x1, n, xd, with_extra=True)
        x, ldx, n = self._stage_x(coded, torch, len(extra))
        # duplicate indices: sample with a generator whose heads repeat the duplicated outputs
        self._handle_override = self._extended_model(extra) if extra else None
        try:
            r = self._hmc_device(x, ldx, n, int(n_mcmc), int(burn_in), float(step_size), int(num_leapfrog_steps),
                                 seed, noise=noise, trace=return_trace)
            torch.cuda.synchronize()
        finally:
            if self._handle_override is not None:
                _lib.load().bgm_hmc_destroy(self._handle_override)
            self._handle_override = None
        T = int(burn_in) + int(n_mcmc)
        counts = r['accept_count'].cpu().numpy()[:T]
        self.last_acceptance_rate = float(counts[int(burn_in):].sum()) / max(1, int(n_mcmc) * n)   # :825
        self.last_step_size = float(r['step'].cpu()[0])
        if verbose:
            print(f"TFP MCMC Acceptance Rate: {self.last_acceptance_rate:.4f}")
        samples = r['samples'].cpu().numpy()
        if return_trace:
            tr = dict(accept=r['accept_mask'].cpu().numpy().astype(bool), log_accept=r['log_accept'].cpu().numpy(),
                      step=r['step_trace'].cpu().numpy(), step_final=self.last_step_size, accept_count=counts,
                      z_final=r['z_state'].cpu().numpy(), lp_final=r['lp_state'].cpu().numpy(),
                      g_final=r['g_state'].cpu().numpy())
            return samples, tr
        return samples

    def philox_noise(self, seed, n, T, row_offset=0):
        """The exact noise `tfp_mcmc_sampler(seed=...)` draws in-kernel."""
        torch = _lib.require_cuda()
        zd = self._p['z_dim']
        z0 = torch.empty((n, zd), dtype=torch.float32, device='cuda')
        mom = torch.empty((T, n, zd), dtype=torch.float32, device='cuda')
        logu = torch.empty((T, n), dtype=torch.float32, device='cuda')
        _lib.call("bgm_hmc_noise", int(seed) & (2 ** 64 - 1), int(row_offset), n, zd, 0, T, _lib.ptr(z0),
                  _lib.ptr(mom), _lib.ptr(logu), _lib.stream_ptr())
        return dict(z0=z0.cpu().numpy(), momentum=mom.cpu().numpy(), log_u=logu.cpu().numpy())

    def _predict_device(self, zs, n_keep, n, seed, row_offset=0, noise=None, sample0=0):
        torch = _lib.require_cuda()
        out = torch.empty((n_keep, n, self._p['x_dim']), dtype=torch.float32, device='cuda')
        nz = self._dev(noise, torch, torch.float32) if noise is not None else None
        _lib.call("bgm_hmc_predict", self._device_model(), _lib.ptr(zs), n_keep, n, int(sample0),
                  int(seed) & (2 ** 64 - 1), int(row_offset), _lib.ptr(nz), _lib.ptr(out), _lib.stream_ptr())
        return out

    def predict_on_posteriors(self, data_posterior_z, *, seed=0, noise=None):
        """bgm/base.py:511-525 -> (n_mcmc, n, x_dim) posterior-predictive draws.
        `noise` (n_mcmc, n, x_dim) injects the N(0,1) draws of `reparameterize` (:113-117)."""
        torch = _lib.require_cuda()
        zs = self._dev(data_posterior_z, torch, torch.float32)
        n_keep, n, _ = zs.shape
        return self._predict_device(zs, n_keep, n, seed, noise=noise).cpu().numpy()

    def generate(self, nb_samples=1000, use_x_sd=True, *, seed=None):
        """bgm/base.py:483-509: z ~ N(0,I) through the generator -> (x, sigma^2).  Like the reference
        every call draws fresh z (and fresh N(0,1) for use_x_sd) -- from the model's private generator,
        or reproducibly from `seed`."""
        torch = _lib.require_cuda()
        if seed is None:
            seed = int(self._noise_rng.randint(0, 2 ** 31 - 1))
        z = torch.from_numpy(np.random.RandomState(seed).standard_normal((nb_samples, self._p['z_dim']))
                             .astype(np.float32)).cuda()
        mu = torch.empty((nb_samples, self._p['x_dim']), dtype=torch.float32, device='cuda')
        var = torch.empty_like(mu)
        _lib.call("bgm_hmc_heads", self._device_model(), _lib.ptr(z), nb_samples, _lib.ptr(mu), _lib.ptr(var),
                  _lib.stream_ptr())
        if use_x_sd:
            draw = self._predict_device(z[None], 1, nb_samples, seed)[0]
            return draw.cpu().numpy(), var.cpu().numpy()
        return mu.cpu().numpy(), var.cpu().numpy()

    def predict(self, data, alpha=0.05, return_samples=False, bs=100, n_mcmc=5000, burn_in=5000, step_size=0.01,
                num_leapfrog_steps=10, seed=42, *, group=None, row_offset=0, n_total=None, verbose=1):
        """bgm/base.py:527-663 -> (imputed data | posterior-predictive samples, intervals).

        The chain states and the predictive draws stay on the device; each `bs` slice of
        rows is reduced there (mean over samples, quantiles on the missing dims) and only
        the results come back.  Under torch.distributed pass `group`, this rank's global
        `row_offset` and the global row count `n_total`: rows are sharded, the only
        collective is the scalar all-reduce of the shared step-size statistic while TFP's
        SimpleStepSizeAdaptation is active.
        """
        assert 0 < alpha < 1, "The significance level 'alpha' must be greater than 0 and less than 1."
        torch = _lib.require_cuda()
        data_np = np.asarray(data, dtype=np.float32)
        n, xd = data_np.shape
        miss = np.isnan(data_np)
        x, ldx, n = self._stage_x(data_np, torch)
        r = self._hmc_device(x, ldx, n, int(n_mcmc), int(burn_in), float(step_size), int(num_leapfrog_steps),
                             seed, row_offset=row_offset, group=group, n_total=n_total)
        T = int(burn_in) + int(n_mcmc)
        counts = r['accept_count'][int(burn_in):T].sum()
        self.last_step_size = float(r['step'].cpu()[0])
        self.last_acceptance_rate = float(counts.item()) / max(1, int(n_mcmc) * n)
        if verbose:
            print(f"TFP MCMC Acceptance Rate: {self.last_acceptance_rate:.4f}")
        zs = r['samples']
        return self._predictive_reduce(zs, data_np, miss, n, xd, int(n_mcmc), alpha, return_samples, bs, seed,
                                       row_offset, x_dev=x[:, :xd])

    def _predictive_reduce(self, zs, data_np, miss, n, xd, n_mcmc, alpha, return_samples, bs, seed, row_offset,
                           x_dev=None):
        """bgm/base.py:603-663 on the device: posterior-predictive draws for every kept state, their mean
        (imputation) and the alpha/2, 1-alpha/2 quantiles of the missing entries.  The reference loops over
        `bs`-row slices on the host and concatenates (n_mcmc, n, x_dim) draws; here the rows are processed in
        chunks sized to a memory budget (the Philox noise is keyed by the global row, so the chunking does not
        change any draw), each chunk is reduced on the device with ONE sort, and the results come back in one
        device-to-host copy at the end.  `bs` only sets the chunk granularity."""
        torch = _lib.require_cuda()
        same_pattern = bool(np.all(miss == miss[0]))                               # :622-623
        miss_idx = np.where(miss[0])[0]
        bs = max(1, int(bs))
        free = _lib.free_memory_estimate(torch)
        per_row = 4.0 * n_mcmc * xd * 3.2            # draws + sorted copy + temporaries
        rows = max(bs, int(min(12e9, 0.3 * free) // per_row) // bs * bs)
        q_lo, q_hi = alpha / 2.0, 1.0 - alpha / 2.0

        def q_pair(srt):
            out = []
            for q in (q_lo, q_hi):
                pos = q * (srt.shape[0] - 1)
                lo = int(np.floor(pos))
                hi = min(lo + 1, srt.shape[0] - 1)
                out.append(srt[lo] + (srt[hi] - srt[lo]) * float(pos - lo))
            return out
        imputed_d = torch.empty((n, xd), dtype=torch.float32, device='cuda')
        k_cols = miss_idx.size if same_pattern else xd
        lo_d = torch.empty((n, k_cols), dtype=torch.float32, device='cuda')
        up_d = torch.empty((n, k_cols), dtype=torch.float32, device='cuda')
        midx_d = torch.from_numpy(miss_idx).cuda() if same_pattern and miss_idx.size else None
        all_draws = [] if return_samples else None
        for i in range(0, n, rows):                                                # :607-612
            j = min(i + rows, n)
            zb = zs[:, i:j, :].contiguous()
            draws = self._predict_device(zb, n_mcmc, j - i, seed, row_offset=row_offset + i)
            if return_samples:
                all_draws.append(draws.cpu().numpy())
            # one pass over the draws: thread = (row, feature) column keeps the few smallest / largest values
            # (bgm_column_quantiles); a quantile deep inside the sample (> 16 order statistics from an end, e.g.
            # n_mcmc = 3000 at alpha = 0.05) falls back to a device sort
            M = (j - i) * xd
            mean_c = imputed_d[i:j]
            lo_c = torch.empty((j - i, xd), dtype=torch.float32, device='cuda') if k_cols else None
            up_c = torch.empty((j - i, xd), dtype=torch.float32, device='cuda') if k_cols else None
            deep = k_cols and max(int(np.floor(q_lo * (n_mcmc - 1))) + 2, n_mcmc - int(np.floor(q_hi * (n_mcmc - 1)))) > 16
            if deep:
                imputed_d[i:j] = draws.mean(dim=0)                                 # :660
                part = draws[:, :, midx_d] if same_pattern else draws
                lo_d[i:j], up_d[i:j] = q_pair(torch.sort(part, dim=0).values)
            else:
                _lib.call("bgm_column_quantiles", _lib.ptr(draws), int(n_mcmc), int(M), float(q_lo), float(q_hi),
                          _lib.ptr(mean_c), _lib.ptr(lo_c), _lib.ptr(up_c), _lib.stream_ptr())
                if k_cols:
                    lo_d[i:j] = lo_c[:, midx_d] if same_pattern else lo_c
                    up_d[i:j] = up_c[:, midx_d] if same_pattern else up_c
            del draws
        if return_samples:
            draws_np = np.concatenate(all_draws, axis=1)
        # the rest stays on the device too: observed entries are put back (:662) and, for a ragged missing
        # pattern, the interval bounds of the missing entries are compacted before the one copy to the host
        if x_dev is None:
            x_dev = torch.from_numpy(np.ascontiguousarray(data_np)).cuda()
        miss_d = torch.isnan(x_dev)
        if same_pattern:
            lo, up = lo_d.cpu().numpy(), up_d.cpu().numpy()
            pred_interval = np.stack([lo, up], axis=-1) if k_cols else np.zeros((n, 0, 2), dtype=np.float32)
        else:
            # :637-649, one (k_i, 2) array per row -- cut from the row-major list of all missing entries
            pairs = torch.stack([lo_d[miss_d], up_d[miss_d]], dim=-1).cpu().numpy()
            ends = np.cumsum(miss.sum(axis=1)).tolist()          # plain slices: np.split spends 2.4 us per piece on axis juggling
            pred_interval = [pairs[a:b] for a, b in zip([0] + ends[:-1], ends)]
        if return_samples:
            return draws_np, pred_interval
        data_imputed = torch.where(miss_d, imputed_d, x_dev).cpu().numpy()                        # :662
        return data_imputed, pred_interval

    # ------------------------------------------------------------ training: fit
    def _encode_device(self, data):
        """e_net(data) with the current weights -> device tensor (n, z_dim) (`data_z_init`, :388; the
        `data_z=None` branch of evaluate, :452-453).  Runs on the layered engine's Dense kernels."""
        torch = _lib.require_cuda()
        x = self._dev(data, torch, torch.float32).contiguous()
        n = x.shape[0]
        z = torch.empty((n, self._p['z_dim']), dtype=torch.float32, device='cuda')
        if self._layered and self._trainer is not None:
            _lib.call("bgm_ltb_encode", self._trainer, _lib.ptr(x), n, _lib.ptr(z), _lib.stream_ptr())
            return z
        self._sync_from_trainer()
        p = dict(self._p)
        saved_gamma = self._p['gamma']
        self._p['gamma'] = 0.0                     # the encoder pass does not depend on it
        try:
            h = self._create_trainer(True)
        finally:
            self._p['gamma'] = saved_gamma
        try:
            _lib.call("bgm_ltb_encode", h, _lib.ptr(x), n, _lib.ptr(z), _lib.stream_ptr())
            torch.cuda.synchronize()
        finally:
            _lib.load().bgm_ltb_destroy(h)
        return z

    def _encode_host(self, data):
        return self._encode_device(data).cpu().numpy()

    def evaluate(self, data, data_z=None, use_x_sd=True, *, seed=0):
        """bgm/base.py:446-471: MSE between the data and its reconstruction from `data_z` (or from
        e_net(data)), generator in inference mode.  With use_x_sd=True the reference adds
        sigma * N(0,1) (`reparameterize`, TF RNG); here the draw comes from `seed`, so the value
        agrees with the reference in distribution, exactly for use_x_sd=False."""
        torch = _lib.require_cuda()
        x = self._dev(data, torch, torch.float32).contiguous()
        n = x.shape[0]
        z = self._encode_device(data) if data_z is None else self._dev(data_z, torch, torch.float32).contiguous()
        if not use_x_sd:
            out = torch.zeros(1, dtype=torch.float64, device='cuda')
            _lib.call(self._tfn("bgm_bgm_evaluate"), self._device_trainer(), _lib.ptr(z), _lib.ptr(x), n, _lib.ptr(out),
                      _lib.stream_ptr())
            return float(out.cpu()[0]) / (n * self._p['x_dim'])
        self._sync_from_trainer()
        pred = self._predict_device(z[None], 1, n, seed)[0]
        return float(((x - pred) ** 2).mean().cpu())

    def fit(self, data, batch_size=32, epochs=100, epochs_per_eval=5, use_egm_init=True, egm_n_iter=20000,
            egm_batches_per_eval=500, verbose=1):
        """bgm/base.py:343-442: optional EGM warm-start, latent table from e(X) (or N(0,1)), then
        `epochs+1` epochs of mini-batch `update_g_net` + `update_latent_variable_sgd` (the incomplete
        last batch is skipped, :401), evaluating every `epochs_per_eval` epochs.  Data, latent table,
        parameters and optimizer state stay on the device; the epoch permutation comes from NumPy's
        global generator (`np.random.choice(n, n, replace=False)`, :397) -- bit-exact index stream."""
        torch = _lib.require_cuda()
        n = len(data)
        bs = int(batch_size)
        if bs > 32:
            self._set_layered(True)      # the fused single-CTA kernels take at most 32 rows
        if self._p['save_res']:
            with open('{}/params.txt'.format(self.save_dir), 'w') as f_params:
                f_params.write(str(self.params))
        if use_egm_init:
            self.egm_init(data, egm_n_iter=egm_n_iter, egm_batches_per_eval=egm_batches_per_eval,
                          batch_size=batch_size, verbose=verbose)
            if verbose:
                print('Initialize latent variables Z with e(V)...')
            data_z_init = self._encode_device(data)                                            # :388
        else:
            if verbose:
                print('Random initialization of latent variables Z...')
            data_z_init = np.random.normal(0, 1, size=(n, self._p['z_dim'])).astype('float32')   # :391
        xd = self._dev(data, torch, torch.float32).contiguous()
        z = data_z_init.clone() if isinstance(data_z_init, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(data_z_init)).cuda()
        self.data_z = z
        tr = self._device_trainer()
        st = _lib.stream_ptr()
        _lib.call(self._tfn("bgm_bgmtrainer_set_iter"), tr, float(self._p['lr_theta']), float(self._p['lr_z']))
        gl = torch.zeros(2, dtype=torch.float32, device='cuda')
        zl = torch.zeros(1, dtype=torch.float32, device='cuda')
        self.history_loss = []
        if verbose:
            print('Iterative Updating Starts ...')
        for epoch in range(int(epochs) + 1):
            sample_idx = np.random.choice(n, n, replace=False)                                # :397
            idx_d = torch.from_numpy(sample_idx.astype(np.int32)).cuda()
            base = idx_d.data_ptr()
            for i in range(0, n - bs + 1, bs):                                                # :401
                ip = C.c_void_p(base + 4 * i)
                _lib.call(self._tfn("bgm_bgm_iter_g"), tr, _lib.ptr(z), _lib.ptr(xd), ip, bs, 1, 1.0, _lib.ptr(gl), st)   # :406
                _lib.call(self._tfn("bgm_bgm_iter_latent"), tr, _lib.ptr(z), _lib.ptr(xd), ip, bs, _lib.ptr(zl), None, st)  # :409-413
            self._trainer_dirty = True
            if epoch % epochs_per_eval == 0:                                                  # :425-442
                mse_x = self.evaluate(data=xd, data_z=z)
                self.history_loss.append(mse_x)
                if verbose:
                    print('Epoch [%d/%d]: MSE_x: %.4f\n' % (epoch, epochs, mse_x))
                if self._p['save_model']:                                                     # :433-436
                    self.save_weights(self.checkpoint_path + "/weights_at_%d.npz" % epoch)
                if self._p['save_res']:
                    gen1, var1 = self.generate(nb_samples=5000)
                    gen12, var12 = self.generate(nb_samples=5000, use_x_sd=False)
                    np.savez('%s/data_gen_at_%d.npz' % (self.save_dir, epoch), gen1=gen1, gen12=gen12,
                             z=z.cpu().numpy(), var1=var1, var12=var12)
        self.last_iter_losses = tuple(float(a) for a in gl.cpu().numpy()) + (float(zl.cpu()[0]),)

    def iter_step(self, data_z_table, data, batch_idx, *, apply=True):
        """One mini-batch of the iterative phase on device copies of the given host arrays (parity
        tests): returns ((loss_x, loss_mse_x), loss_postrior_z, grad rows, updated latent table)."""
        torch = _lib.require_cuda()
        tr = self._device_trainer()
        st = _lib.stream_ptr()
        if not getattr(self, '_iter_ready', False) or self._trainer_epoch != id(tr):
            _lib.call(self._tfn("bgm_bgmtrainer_set_iter"), tr, float(self._p['lr_theta']), float(self._p['lr_z']))
            self._iter_ready, self._trainer_epoch = True, id(tr)
        z = self._dev(data_z_table, torch, torch.float32).contiguous().clone()
        xd = self._dev(data, torch, torch.float32).contiguous()
        idx = torch.from_numpy(np.asarray(batch_idx, np.int32)).cuda()
        bs = len(batch_idx)
        gl = torch.zeros(2, dtype=torch.float32, device='cuda')
        zl = torch.zeros(1, dtype=torch.float32, device='cuda')
        gz = torch.zeros((bs, self._p['z_dim']), dtype=torch.float32, device='cuda')
        _lib.call(self._tfn("bgm_bgm_iter_g"), tr, _lib.ptr(z), _lib.ptr(xd), _lib.ptr(idx), bs, 1 if apply else 0, 1.0,
                  _lib.ptr(gl), st)
        _lib.call(self._tfn("bgm_bgm_iter_latent"), tr, _lib.ptr(z), _lib.ptr(xd), _lib.ptr(idx), bs, _lib.ptr(zl), _lib.ptr(gz), st)
        self._trainer_dirty = True
        g = gl.cpu().numpy()
        return (float(g[0]), float(g[1])), float(zl.cpu()[0]), gz.cpu().numpy(), z.cpu().numpy()

    # ------------------------------------------------------------ EGM training
    def _offset_streams(self, group):
        """Data-parallel contract (`group=`): see CausalBGM._offset_streams -- rank 0 keeps the
        single-GPU host streams, rank r > 0 reseeds NumPy's global generator (mini-batch order, prior
        z) and its private noise generator once per model."""
        import torch.distributed as dist
        r = dist.get_rank(group)
        if r > 0 and not getattr(self, '_streams_offset', False):
            np.random.seed((1024 + 7919 * r) % (2 ** 32))
            self._noise_rng = np.random.RandomState((self._noise_rng.randint(0, 2 ** 31 - 1) + 104729 * r) % (2 ** 32))
        self._streams_offset = True

    def _grad_tensor(self, group):
        torch = _lib.require_cuda()
        n, ptr = C.c_int(), C.c_void_p()
        _lib.call(self._tfn("bgm_trainer_buffers"), self._device_trainer(), group, C.byref(n), None, C.byref(ptr))

        class _View(object):
            __cuda_array_interface__ = dict(shape=(n.value,), typestr='<f4', data=(ptr.value, False), version=2)
        return torch.as_tensor(_View(), device='cuda')

    def _apply(self, group_id, dist_group):
        scale = 1.0
        if dist_group is not None:
            import torch.distributed as dist
            dist.all_reduce(self._grad_tensor(group_id), group=dist_group)
            scale = 1.0 / dist.get_world_size(dist_group)
        _lib.call(self._tfn("bgm_train_adam"), self._device_trainer(), group_id, float(scale), _lib.stream_ptr())
        self._trainer_dirty = True

    def _disc_call(self, z, x, eps_z, eps_x, noise, losses):
        _lib.call(self._tfn("bgm_bgm_train_disc_grad"), self._device_trainer(), _lib.ptr(z), _lib.ptr(x), z.shape[0],
                  float(eps_z), float(eps_x), _lib.ptr(noise), _lib.ptr(losses), _lib.stream_ptr())

    def _gen_call(self, z, x, n1, n2, losses):
        _lib.call(self._tfn("bgm_bgm_train_gen_grad"), self._device_trainer(), _lib.ptr(z), _lib.ptr(x), z.shape[0],
                  _lib.ptr(n1), _lib.ptr(n2), _lib.ptr(losses), _lib.stream_ptr())

    def gradients(self, which, data_z, data_x, *, eps_z=0.5, eps_x=0.5, noise=None, noise2=None):
        """(losses, flat gradient in DEVICE layout) of one step without the optimizer update
        (test hook).  'disc': group 1 = [dz | dx]; 'gen': group 0 (see bgm_bgmtrainer_create)."""
        torch = _lib.require_cuda()
        z = self._dev(data_z, torch, torch.float32)
        x = self._dev(data_x, torch, torch.float32)
        n1 = self._dev(noise, torch, torch.float32)
        if which == 'disc':
            losses = torch.empty(3, dtype=torch.float32, device='cuda')
            self._disc_call(z, x, eps_z, eps_x, n1, losses)
            return losses.cpu().numpy(), self._grad_tensor(1).cpu().numpy()
        n2 = self._dev(noise2, torch, torch.float32)
        losses = torch.empty(6, dtype=torch.float32, device='cuda')
        self._gen_call(z, x, n1, n2, losses)
        return losses.cpu().numpy(), self._grad_tensor(0).cpu().numpy()

    def train_disc_step(self, data_z, data_x, *, eps_z=None, eps_x=None, noise=None, group=None):
        """bgm/base.py:190-245 -> (dz_loss, dx_loss, d_loss).  The U(0,1) draws of :199-200 and the
        N(0,1) draws of reparameterize (:208) come from a private RandomState unless given."""
        torch = _lib.require_cuda()
        z = self._dev(data_z, torch, torch.float32)
        x = self._dev(data_x, torch, torch.float32)
        rs = self._noise_rng
        eps_z = float(rs.uniform()) if eps_z is None else eps_z
        eps_x = float(rs.uniform()) if eps_x is None else eps_x
        noise = rs.standard_normal(tuple(x.shape)).astype(np.float32) if noise is None else noise
        losses = torch.empty(3, dtype=torch.float32, device='cuda')
        self._disc_call(z, x, eps_z, eps_x, self._dev(noise, torch, torch.float32), losses)
        self._apply(1, group)
        return tuple(float(a) for a in losses.cpu().numpy())

    def train_gen_step(self, data_z, data_x, *, noise1=None, noise2=None, group=None):
        """bgm/base.py:247-291 -> (g_loss_adv, e_loss_adv, l2_loss_z, l2_loss_x, reg_loss, g_e_loss)."""
        torch = _lib.require_cuda()
        z = self._dev(data_z, torch, torch.float32)
        x = self._dev(data_x, torch, torch.float32)
        rs = self._noise_rng
        noise1 = rs.standard_normal(tuple(x.shape)).astype(np.float32) if noise1 is None else noise1
        noise2 = rs.standard_normal(tuple(x.shape)).astype(np.float32) if noise2 is None else noise2
        losses = torch.empty(6, dtype=torch.float32, device='cuda')
        self._gen_call(z, x, self._dev(noise1, torch, torch.float32), self._dev(noise2, torch, torch.float32), losses)
        self._apply(0, group)
        return tuple(float(a) for a in losses.cpu().numpy())

    def egm_init(self, data, egm_n_iter=10000, batch_size=32, egm_batches_per_eval=500, verbose=1, *, group=None,
                 eval_during=True):
        """bgm/base.py:294-340: mini-batches from `Base_sampler` (its shuffled index stream is
        NumPy's legacy generator, bit-exact), prior draws from `z_sampler.get_batch`.  Every
        `egm_batches_per_eval` iterations evaluate() (and, with save_res, generate()/np.savez) runs
        like :317-339 (`eval_during=False` skips it); history in `self.egm_history`."""
        torch = _lib.require_cuda()
        if group is not None:
            self._offset_streams(group)
        if int(batch_size) > 32:
            self._set_layered(True)      # the fused single-CTA kernels take at most 32 rows
        data = np.asarray(data, dtype=np.float32)
        ds_seed = 123
        if group is not None:
            import torch.distributed as dist
            ds_seed = 123 + 7919 * dist.get_rank(group)
        self.data_sampler = Base_sampler(x=data, y=data, v=data, batch_size=batch_size, normalize=False,
                                         random_seed=ds_seed)                                        # :295
        dloss = torch.zeros(3, dtype=torch.float32, device='cuda')
        gloss = torch.zeros(6, dtype=torch.float32, device='cuda')
        rs = self._noise_rng
        self.egm_history = []
        if verbose:
            print('EGM Initialization Starts ...')
        for it in range(int(egm_n_iter) + 1):
            for _ in range(int(self._p['g_d_freq'])):
                bx, _, _ = self.data_sampler.next_batch()                                    # :300
                bz = self.z_sampler.get_batch(batch_size)                                   # :301
                x, z = self._dev(bx, torch, torch.float32), self._dev(bz, torch, torch.float32)
                nz = self._dev(rs.standard_normal(bx.shape).astype(np.float32), torch, torch.float32)
                self._disc_call(z, x, rs.uniform(), rs.uniform(), nz, dloss)
                self._apply(1, group)
            bx, _, _ = self.data_sampler.next_batch()                                        # :304
            bz = self.z_sampler.get_batch(batch_size)                                       # :305
            x, z = self._dev(bx, torch, torch.float32), self._dev(bz, torch, torch.float32)
            n1 = self._dev(rs.standard_normal(bx.shape).astype(np.float32), torch, torch.float32)
            n2 = self._dev(rs.standard_normal(bx.shape).astype(np.float32), torch, torch.float32)
            self._gen_call(z, x, n1, n2, gloss)
            self._apply(0, group)
            if it % egm_batches_per_eval == 0:                                             # :310-337
                if verbose:
                    d, g = dloss.cpu().numpy(), gloss.cpu().numpy()
                    print('EGM Initialization Iter [%d] : g_loss_adv[%.4f], e_loss_adv [%.4f], l2_loss_z [%.4f], '
                          'l2_loss_x [%.4f], sd^2_loss[%.4f], g_e_loss [%.4f], dz_loss [%.4f], dx_loss[%.4f], d_loss [%.4f]'
                          % (it, g[0], g[1], g[2], g[3], g[4], g[5], d[0], d[1], d[2]))
                if eval_during:
                    self._trainer_dirty = True
                    mse_sd = self.evaluate(data=data, use_x_sd=True)
                    mse = self.evaluate(data=data, use_x_sd=False)
                    self.egm_history.append((it, mse_sd, mse))
                    if verbose:
                        print('iter [%d/%d]: MSE_x: %.4f\n' % (it, egm_n_iter, mse_sd))
                        print('iter [%d/%d]: MSE_x no x_sd: %.4f\n' % (it, egm_n_iter, mse))
                    if self._p['save_res']:
                        gen1, var1 = self.generate(nb_samples=5000)
                        gen12, var12 = self.generate(nb_samples=5000, use_x_sd=False)
                        np.savez('%s/init_data_gen_at_%d.npz' % (self.save_dir, it), gen1=gen1, gen12=gen12,
                                 z=self._encode_host(data), var1=var1, var12=var12)
                    if self._p['save_model']:                                                 # :334-338
                        self.save_weights(self.checkpoint_path + "/weights_at_egm_init_%d.npz" % it)
        if verbose:
            print('EGM Initialization Ends.')
        return tuple(float(a) for a in dloss.cpu().numpy()), tuple(float(a) for a in gloss.cpu().numpy())
