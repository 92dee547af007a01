"""CausalBGM with the posterior-sampling path on B200 (sm_100a) kernels.

Drop-in for the method surface of `bayesgm.models.causalbgm.CausalBGM`
(`src/bayesgm/models/causalbgm/base.py`): same constructor, same method names,
kwargs, return types and error behaviour for the path this package accelerates --
`get_log_posterior` (:765), `metropolis_hastings_sampler` (:820),
`infer_from_latent_posterior` (:671) and `predict` (:573).  Inputs and outputs are
host NumPy arrays like the reference's; underneath every method calls the C ABI of
libbgm_b200.so (include/bgm_b200.h) on device buffers.  There is no CPU fallback.

`use_bnn=False`: deterministic networks (networks/base.py), persistent tcgen05 / SIMT sampler.
`use_bnn=True` (the shipped default): Bayesian networks (networks/bnn.py: input BatchNormalization on
batch statistics + DenseFlipout) -- every evaluation draws fresh network noise and couples the rows
of a `bs` slice through the batch statistics, see csrc/bnn.cuh.
"""
import ctypes as C
import datetime
import os

import numpy as np

from . import _lib
from .datasets import Gaussian_sampler
from .nets import DenseNet, DiscNet, BayesDenseNet
from .shard import merge_adrf, finish_adrf

_DEFAULTS = dict(use_bnn=True, g_units=[64] * 5, e_units=[64] * 5, f_units=[64, 32, 8],
                 h_units=[64, 32, 8], dz_units=[64, 32, 8], lr=0.0002, lr_theta=0.0001,
                 lr_z=0.0001, g_d_freq=5, save_model=False, save_res=True, kl_weight=0.0001,
                 use_z_rec=True)


def _quantile_dim0(torch, a, q):
    """np.quantile(a, q, axis=0) (linear interpolation) on the device, without
    torch.quantile's input-size limit."""
    srt = torch.sort(a, dim=0).values
    pos = q * (a.shape[0] - 1)
    lo = int(np.floor(pos))
    hi = min(lo + 1, a.shape[0] - 1)
    frac = float(pos - lo)
    return srt[lo] + (srt[hi] - srt[lo]) * frac


def _ite_summary(torch, eff, alpha):
    """causalbgm/base.py:640-642 on the device: mean and the alpha/2, 1 - alpha/2 quantiles over the kept states of
    the per-subject draws eff (n_keep, n) -> three (n,) NumPy arrays.  One pass (bgm_column_quantiles: thread =
    subject, the few smallest / largest draws in registers); a device sort only when a quantile sits more than 16
    order statistics from an end of the sample."""
    S, n = eff.shape
    q_lo, q_hi = alpha / 2.0, 1.0 - alpha / 2.0
    need = max(int(np.floor(q_lo * (S - 1))) + 2, S - int(np.floor(q_hi * (S - 1))))
    if need > 16:
        return (eff.mean(dim=0).cpu().numpy(), _quantile_dim0(torch, eff, q_lo).cpu().numpy(),
                _quantile_dim0(torch, eff, q_hi).cpu().numpy())
    eff = eff.contiguous()
    out = torch.empty((3, n), dtype=torch.float32, device='cuda')
    _lib.call("bgm_column_quantiles", _lib.ptr(eff), int(S), int(n), float(q_lo), float(q_hi), _lib.ptr(out[0]), _lib.ptr(out[1]),
              _lib.ptr(out[2]), _lib.stream_ptr())
    h = out.cpu().numpy()
    return h[0], h[1], h[2]


class CausalBGM(object):
    """See the reference docstring, causalbgm/base.py:12-54, for `params`."""

    def __init__(self, params, timestamp=None, random_seed=None):
        self.params = params
        self.timestamp = timestamp
        p = dict(_DEFAULTS)
        p.update(params)
        self._p = p
        self._bnn = bool(p['use_bnn'])
        if random_seed is not None:
            np.random.seed(random_seed)
        zd = sum(p['z_dims'])
        z0, z1, z2, _ = p['z_dims']
        rng = np.random.RandomState(random_seed) if random_seed is not None else np.random
        Net = BayesDenseNet if self._bnn else DenseNet                                 # :64-81
        self.g_net = Net(zd, p['v_dim'] + 1, 'g_net', p['g_units'], rng)               # :65 / :74
        self.e_net = Net(p['v_dim'], zd, 'e_net', p['e_units'], rng)                   # :67 / :76
        self.f_net = Net(z0 + z1 + 1, 2, 'f_net', p['f_units'], rng)                   # :69 / :78
        self.h_net = Net(z0 + z2, 2, 'h_net', p['h_units'], rng)                       # :71 / :80
        self._bnn_calls = 0              # get_log_posterior call counter (keys the network-noise stream)
        self.dz_net = DiscNet(zd, 'dz_net', p['dz_units'], rng)                        # :83
        self.z_sampler = Gaussian_sampler(mean=np.zeros(zd), sd=1.0)                  # :88 (reseeds to 1024)
        self._trainer = None
        self._trainer_dirty = False      # device parameters newer than the host arrays
        self._layered = self._bnn        # which training engine _device_trainer() builds
        self._lt_seed = int(np.random.RandomState(random_seed).randint(1, 2 ** 31 - 1)) if random_seed is not None \
            else 20240229
        self._eps_rng = np.random.RandomState(0 if random_seed is None else random_seed)
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
        self.last_q_sd = None

    # ------------------------------------------------------------------ plumbing
    def get_config(self):
        return {"params": self.params}

    def initialize_nets(self, print_summary=False):
        if print_summary:
            for net in (self.g_net, self.f_net, self.h_net):
                print(net.model_name, net.dims)

    def set_weights(self, g=None, e=None, f=None, h=None, dz=None):
        """Load Keras-layout weights ([kernel, bias, ...] per net; dz: the Discriminator's
        trainable_variables), e.g. exported from a trained reference model."""
        self._sync_from_trainer()
        for net, w in ((self.g_net, g), (self.e_net, e), (self.f_net, f), (self.h_net, h)):
            if w is not None:
                net.set_weights(w)
        if dz is not None:
            self.dz_net.set_trainable(dz)
        self._drop_handle()
        self._drop_trainer()

    # Two training engines behind the same calls: the fused single-CTA kernels (csrc/train.cuh: deterministic
    # nets, batch <= 32) and the layered engine (csrc/layered.cuh: Bayesian nets, any batch size / width).
    _LT_NAMES = dict(bgm_trainer_destroy="bgm_lt_destroy", bgm_trainer_buffers="bgm_lt_buffers",
                     bgm_trainer_get_params="bgm_lt_get_params", bgm_trainer_set_params="bgm_lt_set_params",
                     bgm_train_disc_grad="bgm_lt_disc_grad", bgm_train_gen_grad="bgm_lt_gen_grad",
                     bgm_train_adam="bgm_lt_adam", bgm_trainer_set_iter="bgm_lt_set_iter",
                     bgm_train_iter_nets="bgm_lt_iter_nets", bgm_train_iter_latent="bgm_lt_iter_latent",
                     bgm_causal_evaluate="bgm_lt_evaluate")

    def _tfn(self, name):
        """Symbol of the active training engine for the generic entry-point name."""
        return self._LT_NAMES[name] if self._layered else name

    def _set_layered(self, on):
        """Chooses the training engine (Bayesian nets always train on the layered one)."""
        on = bool(on) or self._bnn
        if on != self._layered:
            self._sync_from_trainer()
            self._drop_trainer()
            self._layered = on

    def _drop_trainer(self):
        if self._trainer is not None:
            getattr(_lib.load(), self._tfn("bgm_trainer_destroy"))(self._trainer)
            self._trainer = None
            self._trainer_dirty = False

    def _sync_from_trainer(self):
        """Pull the trained parameters back into the host arrays (and invalidate the packed
        sampler model) -- done lazily, when something reads the weights."""
        if self._trainer is None or not self._trainer_dirty:
            return
        n = C.c_int()
        for group in (0, 1):
            _lib.call(self._tfn("bgm_trainer_buffers"), self._trainer, group, C.byref(n), None, None)
            flat = np.empty(n.value, np.float32)
            _lib.call(self._tfn("bgm_trainer_get_params"), self._trainer, group, flat.ctypes.data_as(C.c_void_p))
            if group == 0:
                o = 0
                for net in (self.g_net, self.e_net, self.f_net, self.h_net):
                    k = net.flat_params().size
                    net.load_flat(flat[o:o + k])
                    o += k
            else:
                self.dz_net.load_flat(flat)
        self._trainer_dirty = False
        self._drop_handle()

    def _lt_desc(self, net):
        """bgm_bnn_net_desc of a net for the layered engine (deterministic nets: bn = NULL)."""
        if self._bnn:
            return net.desc()
        dims = (C.c_int * len(net.dims))(*net.dims)
        flat = net.flat_params()
        d = _lib.BnnNetDesc(len(net.layers), C.cast(dims, C.POINTER(C.c_int)), None, flat.ctypes.data_as(C.POINTER(C.c_float)))
        return d, (dims, flat)

    def _device_trainer(self):
        if self._trainer is None and self._layered:
            _lib.require_cuda()
            p = self._p
            zd4 = (C.c_int * 4)(*[int(d) for d in p['z_dims']])
            descs = [self._lt_desc(net) for net in (self.g_net, self.e_net, self.f_net, self.h_net)]
            dd, dk = self.dz_net.desc()
            h = C.c_void_p()
            _lib.call("bgm_lt_create", C.byref(h), zd4, int(p['v_dim']), int(bool(p['binary_treatment'])),
                      int(bool(p['use_z_rec'])), int(self._bnn), C.byref(descs[0][0]), C.byref(descs[1][0]),
                      C.byref(descs[2][0]), C.byref(descs[3][0]), C.byref(dd), float(p['lr']), 0.9, 0.99,
                      float(p['kl_weight']) if self._bnn else 0.0, int(self._lt_seed))
            self._trainer = h
        if self._trainer is None:
            _lib.require_cuda()
            p = self._p
            zd4 = (C.c_int * 4)(*[int(d) for d in p['z_dims']])
            descs = [net.desc() for net in (self.g_net, self.e_net, self.f_net, self.h_net)]
            dd, dk = self.dz_net.desc()
            h = C.c_void_p()
            _lib.call("bgm_trainer_create", C.byref(h), zd4, int(p['v_dim']), int(bool(p['binary_treatment'])),
                      int(bool(p['use_z_rec'])), C.byref(descs[0][0]), C.byref(descs[1][0]),
                      C.byref(descs[2][0]), C.byref(descs[3][0]), C.byref(dd), float(p['lr']), 0.9, 0.99)
            self._trainer = h
        return self._trainer

    def _drop_handle(self):
        if self._handle is not None:
            if self._bnn:
                _lib.load().bgm_bnn_destroy(self._handle)
            else:
                _lib.load().bgm_causal_destroy(self._handle)
            self._handle = None

    def __del__(self):
        try:
            self._drop_handle()
            self._drop_trainer()
        except Exception:
            pass

    def _device_model(self):
        """Packs g/f/h for the kernels (once per weight change)."""
        self._sync_from_trainer()
        if self._handle is None and self._bnn:
            _lib.require_cuda()
            p = self._p
            zd4 = (C.c_int * 4)(*[int(d) for d in p['z_dims']])
            gd, gk = self.g_net.desc()
            fd, fk = self.f_net.desc()
            hd, hk = self.h_net.desc()
            h = C.c_void_p()
            sig = [float(p[k]) if k in p else -1.0 for k in ('sigma_v', 'sigma_x', 'sigma_y')]
            _lib.call("bgm_bnn_create", C.byref(h), zd4, int(p['v_dim']), int(bool(p['binary_treatment'])),
                      sig[0], sig[1], sig[2], C.byref(gd), C.byref(fd), C.byref(hd))
            self._handle = h
            if getattr(self, 'bnn_plan', None):          # 1: thread = row, 2 (default): two threads per row
                _lib.call("bgm_bnn_set_plan", h, int(self.bnn_plan))
        if self._handle is None:
            _lib.require_cuda()
            p = self._p
            zd4 = (C.c_int * 4)(*[int(d) for d in p['z_dims']])
            gd, gk = self.g_net.desc()
            fd, fk = self.f_net.desc()
            hd, hk = self.h_net.desc()
            h = C.c_void_p()
            sig = [float(p[k]) if k in p else -1.0 for k in ('sigma_v', 'sigma_x', 'sigma_y')]
            _lib.call("bgm_causal_create", C.byref(h), zd4, int(p['v_dim']), int(bool(p['binary_treatment'])),
                      sig[0], sig[1], sig[2], C.byref(gd), C.byref(fd), C.byref(hd))
            self._handle = h
            kind = {'auto': 0, 'simt': 1, 'tensor': 2}[getattr(self, 'sampler_engine', 'auto')]
            if kind:
                _lib.call("bgm_causal_set_sampler", h, kind)
        return self._handle

    def set_sampler_engine(self, engine='auto'):
        """'auto' (tensor-core engine when the net shape allows it), 'simt' (fp32 FMA-pipe
        engine) or 'tensor' (raises if unavailable).  Both run the same algorithm on the
        same Philox streams; see DESIGN.md 4.1 / 4.1b."""
        if engine not in ('auto', 'simt', 'tensor'):
            raise ValueError("engine must be 'auto', 'simt' or 'tensor'")
        self.sampler_engine = engine
        if self._handle is not None:
            _lib.call("bgm_causal_set_sampler", self._handle, {'auto': 0, 'simt': 1, 'tensor': 2}[engine])

    def sampler_info(self):
        if self._bnn:
            smem, rows = C.c_int(), C.c_int()
            macs = C.c_longlong()
            _lib.call("bgm_bnn_info", self._device_model(), C.byref(smem), C.byref(rows), C.byref(macs))
            return dict(engine='bnn', tensor_available=False, tensor_smem_bytes=0, smem_bytes=smem.value,
                        rows_per_cta=rows.value, macs_per_eval=macs.value,
                        kernel=('bnn_mh_kernel<%d>' if getattr(self, 'bnn_plan', 2) == 1 else 'bnn_mh2_kernel<%d>') % (
                            (8 if sum(self._p['z_dims']) <= 8 else 16 if sum(self._p['z_dims']) <= 16 else 32) //
                            (1 if getattr(self, 'bnn_plan', 2) == 1 else 2)))
        kind, avail, smem = C.c_int(), C.c_int(), C.c_int()
        issued = C.c_longlong()
        _lib.call("bgm_causal_sampler_info", self._device_model(), C.byref(kind), C.byref(avail), C.byref(smem),
                  C.byref(issued))
        buf = C.create_string_buffer(64)
        _lib.call("bgm_causal_kernel_name", self._device_model(), buf, 64)
        return dict(engine={1: 'simt', 2: 'tensor'}[kind.value], tensor_available=bool(avail.value),
                    tensor_smem_bytes=smem.value, tensor_issued_macs_per_row=issued.value,
                    kernel=buf.value.decode())

    def launch_smem_bytes(self):
        """Dynamic shared memory of the sampler launch (what ncu reports as dynamic smem per block)."""
        si = self.sampler_info()
        if si['engine'] == 'tensor':
            return si['tensor_smem_bytes']
        return si['smem_bytes'] if si['engine'] == 'bnn' else self.kernel_info()['smem_bytes']

    def kernel_info(self):
        smem, warps, nops, proj = C.c_int(), C.c_int(), C.c_int(), C.c_int()
        macs, issued = C.c_longlong(), C.c_longlong()
        _lib.call("bgm_causal_info", self._device_model(), C.byref(smem), C.byref(warps), C.byref(nops),
                  C.byref(macs), C.byref(issued), C.byref(proj))
        return dict(smem_bytes=smem.value, warps_per_cta=warps.value, n_ops=nops.value,
                    macs_per_row=macs.value, issued_macs_per_row=issued.value, proj_dim=proj.value)

    def _aux(self, v, ldv, n):
        """Per-data-set device buffers: the projected covariates (models with proj_dim > 0,
        see bgm_causal_project) and the scratch of the in-kernel work scheduler."""
        torch = _lib.require_cuda()
        if self._bnn:
            nd = _lib.load().bgm_bnn_scratch_doubles(self._device_model(), n)
            return dict(vproj=None, r0=None, ldvproj=0, sched=None,
                        scratch=torch.empty(int(nd), dtype=torch.float64, device='cuda'))
        sched = torch.empty((n + 31) // 32 + 1, dtype=torch.int32, device='cuda')
        if not hasattr(self, '_proj_dim') or self._handle is None:
            self._proj_dim = self.kernel_info()['proj_dim']
        if not self._proj_dim:
            return dict(vproj=None, r0=None, ldvproj=0, sched=sched)
        ldp = (self._proj_dim + 3) // 4 * 4
        vproj = torch.empty((n, ldp), dtype=torch.float32, device='cuda')
        r0 = torch.empty(n, dtype=torch.float32, device='cuda')
        _lib.call("bgm_causal_project", self._device_model(), _lib.ptr(v), ldv, n, _lib.ptr(vproj), ldp,
                  _lib.ptr(r0), _lib.stream_ptr())
        return dict(vproj=vproj, r0=r0, ldvproj=ldp, sched=sched)

    @staticmethod
    def _to_device(a, torch, cols=None):
        """Host array (NumPy or torch CPU, possibly pinned) -> contiguous float32 device
        tensor; `cols` pads the row length with zeros (the kernels need ldv % 4 == 0)."""
        if isinstance(a, torch.Tensor):
            t = a
        else:
            t = torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32))
        if t.dtype != torch.float32:
            t = t.float()
        if t.is_cuda:
            d = t.contiguous()
        else:
            d = t.contiguous().to('cuda', non_blocking=True)
        if cols is not None and d.shape[1] != cols:
            pad = torch.zeros((d.shape[0], cols), dtype=torch.float32, device='cuda')
            pad[:, :d.shape[1]] = d
            d = pad
        return d

    def _stage(self, data):
        torch = _lib.require_cuda()
        data_x, data_y, data_v = data
        n = len(data_x)
        p = self._p['v_dim']
        if data_v.shape[1] != p:
            raise ValueError("data_v has %d columns, params['v_dim'] is %d" % (data_v.shape[1], p))
        ldv = (p + 3) // 4 * 4
        x = self._to_device(data_x, torch).reshape(-1)
        y = self._to_device(data_y, torch).reshape(-1)
        v = self._to_device(data_v, torch, cols=ldv)
        assert x.numel() == n and y.numel() == n and v.shape[0] == n
        return torch, x, y, v, ldv, n

    # --------------------------------------------------------------- hot path
    def get_log_posterior(self, data_x, data_y, data_v, data_z, eps=1e-6, *, seed=None, call=None):
        """causalbgm/base.py:765-817 -> (n,) float32 NumPy array.  With Bayesian nets every call
        draws fresh network noise (Philox stream `seed`, call id `call`; by default a per-model seed
        and a running call counter) and normalises the inputs with the statistics of this batch."""
        torch, x, y, v, ldv, n = self._stage((data_x, data_y, data_v))
        z = self._to_device(data_z, torch)
        zd = sum(self._p['z_dims'])
        if z.shape != (n, zd):
            raise ValueError("data_z must have shape (%d, %d)" % (n, zd))
        out = torch.empty(n, dtype=torch.float32, device='cuda')
        aux = self._aux(v, ldv, n)
        if self._bnn:
            if seed is None:
                if not hasattr(self, '_bnn_seed'):
                    self._bnn_seed = int(np.random.randint(0, 2 ** 31 - 1))
                seed = self._bnn_seed
            if call is None:
                call = self._bnn_calls
                self._bnn_calls += 1
            _lib.call("bgm_bnn_logpost", self._device_model(), _lib.ptr(x), _lib.ptr(y), _lib.ptr(v), ldv, _lib.ptr(z),
                      n, int(seed) & (2 ** 64 - 1), 0, 0, int(call) & 0xFFFFFFFF, _lib.ptr(aux['scratch']),
                      _lib.ptr(out), _lib.stream_ptr())
            return out.cpu().numpy()
        _lib.call("bgm_causal_logpost", self._device_model(), _lib.ptr(x), _lib.ptr(y), _lib.ptr(v), ldv,
                  _lib.ptr(aux['vproj']), aux['ldvproj'], _lib.ptr(aux['r0']), _lib.ptr(z), n, _lib.ptr(out),
                  _lib.ptr(aux['sched']), _lib.stream_ptr())
        return out.cpu().numpy()

    def _mh_device(self, x, y, v, ldv, n, burn_in, n_keep, q_sd, adaptive_sd, initial_q_sd,
                   target_acceptance_rate, tolerance, adjustment_interval, window_size,
                   seed, row_offset, noise=None, trace=False, keep_samples=True, aux=None, slice_id=0, prior=None):
        """Runs the sampler on staged device buffers; returns a dict of device tensors.  `prior`: device
        rows (n, zd+1) of a conditional prior (IdentifiableCausalBGM), see bgm_mh_args.prior_dev."""
        torch = _lib.require_cuda()
        aux = aux if aux is not None else self._aux(v, ldv, n)
        zd = sum(self._p['z_dims'])
        T = burn_in + n_keep
        m = self._device_model()
        dev = 'cuda'
        z_state = torch.empty((n, zd), dtype=torch.float32, device=dev)
        lp_state = torch.empty(n, dtype=torch.float32, device=dev)
        samples = torch.empty((n_keep, n, zd), dtype=torch.float32, device=dev) if keep_samples else None
        acc_count = torch.zeros(T, dtype=torch.int32, device=dev)
        q = torch.tensor([initial_q_sd if adaptive_sd else q_sd], dtype=torch.float64, device=dev)
        a = _lib.MhArgs()
        a.x_dev, a.y_dev, a.v_dev = x.data_ptr(), y.data_ptr(), v.data_ptr()
        a.ldv, a.n = ldv, n
        if aux['vproj'] is not None:
            a.vproj_dev, a.r0_dev, a.ldvproj = aux['vproj'].data_ptr(), aux['r0'].data_ptr(), aux['ldvproj']
        if aux['sched'] is not None:
            a.sched_dev = aux['sched'].data_ptr()
        a.z_state_dev, a.lp_state_dev = z_state.data_ptr(), lp_state.data_ptr()
        a.burn_in = burn_in
        a.q_sd_dev = q.data_ptr()
        a.seed, a.row_offset = int(seed) & (2 ** 64 - 1), int(row_offset)
        a.out_samples_dev = samples.data_ptr() if keep_samples else None
        a.accept_count_dev = acc_count.data_ptr()
        keep = [aux]
        if prior is not None:
            if self._bnn:
                raise NotImplementedError("bayesgm_b200: conditional prior with Bayesian nets")
            assert prior.shape == (n, zd + 1) and prior.is_contiguous()
            a.prior_dev, a.ldprior = prior.data_ptr(), zd + 1
            keep.append(prior)
        if noise is not None:
            z0 = self._to_device(noise['z0'], torch)
            eps = self._to_device(noise['eps'], torch)
            u = torch.from_numpy(np.ascontiguousarray(noise['u'], dtype=np.float64)).to(dev)
            assert z0.shape == (n, zd) and eps.shape == (T, n, zd) and u.shape == (T, n)
            z_state.copy_(z0)
            a.eps_dev, a.u_dev = eps.data_ptr(), u.data_ptr()
            a.init_mode = 1
            keep += [eps, u]
        else:
            a.init_mode = 2
        out = dict(samples=samples, z_state=z_state, lp_state=lp_state, accept_count=acc_count, q_sd=q)
        if trace:
            out['accept_mask'] = torch.zeros((T, n), dtype=torch.uint8, device=dev)
            out['lp_trace'] = torch.zeros((T, n), dtype=torch.float32, device=dev)
            a.accept_mask_dev = out['accept_mask'].data_ptr()
            a.lp_trace_dev = out['lp_trace'].data_ptr()
        st = _lib.stream_ptr()
        if self._bnn:
            # Bayesian nets: one launch per iteration inside bgm_bnn_mh (batch statistics couple the rows)
            lpc = None
            if trace:
                out['lp_cur_trace'] = torch.zeros((T, n), dtype=torch.float32, device=dev)
                lpc = out['lp_cur_trace']

            def run(args):
                _lib.call("bgm_bnn_mh", m, C.byref(args), int(slice_id), _lib.ptr(aux['scratch']), _lib.ptr(lpc), st)
        else:
            def run(args):
                _lib.call("bgm_causal_mh", m, C.byref(args), st)
        if not adaptive_sd:
            a.t_begin, a.t_end = 0, T
            run(a)
        else:
            # q_sd changes after iterations 50, 100, ... < burn_in (:880): one launch per
            # constant-q_sd stretch, the rule itself runs on the device in between.
            cuts = [c for c in range(adjustment_interval, burn_in, adjustment_interval)]
            begin = 0
            for c in cuts + [None]:
                end = T if c is None else c + 1
                a.t_begin, a.t_end = begin, end
                run(a)
                a.init_mode = 0
                if c is not None:
                    _lib.call("bgm_mh_adapt_qsd", C.c_void_p(acc_count.data_ptr()), c, window_size, n,
                              float(target_acceptance_rate), float(tolerance), C.c_void_p(q.data_ptr()), st)
                begin = end
        out['_keep'] = keep
        return out

    def metropolis_hastings_sampler(self, data, initial_q_sd=1.0, q_sd=None, burn_in=5000, n_keep=3000,
                                    target_acceptance_rate=0.25, tolerance=0.05, adjustment_interval=50,
                                    adaptive_sd=None, window_size=100, *, seed=None, noise=None,
                                    return_trace=False, verbose=1):
        """causalbgm/base.py:820-904 -> np.ndarray (n_keep, n, zd).

        Noise: by default an in-kernel Philox4x32-10 stream keyed by `seed` (drawn from
        NumPy's global generator when None, so `np.random.seed` still makes runs
        repeatable).  `noise=dict(z0, eps, u)` injects pre-drawn N(0,1) / U(0,1) values
        (shapes (n,zd), (T,n,zd), (T,n)) for state-for-state parity tests.
        """
        torch, x, y, v, ldv, n = self._stage(data)
        if adaptive_sd is None:                                                   # :852-853
            adaptive_sd = (q_sd is None or q_sd <= 0)
        if seed is None:
            seed = int(np.random.randint(0, 2 ** 31 - 1)) * (2 ** 31) + int(np.random.randint(0, 2 ** 31 - 1))
        r = self._mh_device(x, y, v, ldv, n, int(burn_in), int(n_keep), q_sd, adaptive_sd, initial_q_sd,
                            target_acceptance_rate, tolerance, adjustment_interval, window_size,
                            seed, 0, noise=noise, trace=return_trace)
        samples = r['samples'].cpu().numpy()
        counts = r['accept_count'].cpu().numpy()
        T = burn_in + n_keep
        w = min(window_size, T)
        self.last_acceptance_rate = float(counts[T - w:].sum()) / (w * n)         # :901
        self.last_q_sd = float(r['q_sd'].cpu()[0])
        if verbose:
            print(f"Final MCMC Acceptance Rate: {self.last_acceptance_rate:.4f}")
        if return_trace:
            tr = dict(accept=r['accept_mask'].cpu().numpy().astype(bool), lp_prop=r['lp_trace'].cpu().numpy(),
                      accept_count=counts, q_sd_final=self.last_q_sd,
                      z_final=r['z_state'].cpu().numpy(), lp_final=r['lp_state'].cpu().numpy())
            if 'lp_cur_trace' in r:
                tr['lp_cur'] = r['lp_cur_trace'].cpu().numpy()
            return samples, tr
        return samples

    def _effect_device(self, z_samples, n_keep, n, x_values, sample_y, seed, row_offset, noise=None, memoise=True):
        """ITE draws (n_keep, n) or ADRF partial sums (n_x, n_keep) on the device.  memoise: evaluate
        f_net once per DISTINCT kept state of a row (a rejected proposal repeats the state) and
        combine -- identical results, ~acceptance-rate of the f_net work (bgm_causal_effect_index /
        _compact / _heads / _combine); falls back to the direct kernel for a single kept state."""
        torch = _lib.require_cuda()
        m = self._device_model()
        st = _lib.stream_ptr()
        binary = bool(self._p['binary_treatment'])
        nz = self._to_device(noise, torch) if noise is not None else None
        seed = int(seed) & (2 ** 64 - 1)
        if binary:
            out = torch.empty((n_keep, n), dtype=torch.float32, device='cuda')
            xv, n_x = None, 2
        else:
            xv = self._dose_grid(x_values, torch)
            n_x = len(x_values)
            out = torch.zeros((n_x, n_keep), dtype=torch.float64, device='cuda')
        total = n_keep * n
        if self._bnn:
            # one f_net call per (kept state, dose) with its own network noise: nothing to memoise
            stats = torch.empty((n_keep, 2 * (self._p['z_dims'][0] + self._p['z_dims'][1])), dtype=torch.float32,
                                device='cuda')
            for s0 in range(0, n_keep, 65535):
                s1 = min(s0 + 65535, n_keep)
                assert s0 == 0, "more than 65535 kept states per call are not supported with Bayesian nets"
                _lib.call("bgm_bnn_effect", m, _lib.ptr(z_samples), s1 - s0, n, _lib.ptr(xv), n_x, int(bool(sample_y)),
                          seed, int(row_offset), _lib.ptr(nz), _lib.ptr(stats), None if binary else _lib.ptr(out),
                          _lib.ptr(out) if binary else None, st)
            return out
        if not memoise or n_keep < 2 or total >= 2 ** 31:
            _lib.call("bgm_causal_effect", m, _lib.ptr(z_samples), n_keep, n, _lib.ptr(xv), n_x, int(bool(sample_y)),
                      seed, int(row_offset), _lib.ptr(nz), None if binary else _lib.ptr(out),
                      _lib.ptr(out) if binary else None, st)
            return out
        zd = sum(self._p['z_dims'])
        local = torch.empty(total, dtype=torch.int32, device='cuda')
        rowtot = torch.empty(n, dtype=torch.int32, device='cuda')
        rowend = torch.empty(n, dtype=torch.int32, device='cuda')
        scratch = torch.empty((n + 2047) // 2048, dtype=torch.int32, device='cuda')
        _lib.call("bgm_causal_effect_index", _lib.ptr(z_samples), n_keep, n, zd, _lib.ptr(local), _lib.ptr(rowtot),
                  _lib.ptr(rowend), _lib.ptr(scratch), st)
        n_distinct = int(rowend[-1].item())                   # the one host sync of the memoised path
        if 8.0 * n_distinct * n_x > min(32e9, 0.4 * _lib.free_memory_estimate(torch)):
            # the (mu, sigma) table of the distinct states would not fit comfortably: evaluate directly
            del local, rowtot, rowend, scratch
            return self._effect_device(z_samples, n_keep, n, x_values, sample_y, seed, row_offset, noise=noise,
                                       memoise=False)
        # the number of distinct states changes from call to call: carve these two out of workspaces that only
        # grow (geometrically), so that repeated predict() calls do not cudaMalloc / cudaFree gigabytes each time
        zlist = self._workspace('zlist', n_distinct * zd, torch).view(n_distinct, zd)
        _lib.call("bgm_causal_effect_compact", _lib.ptr(z_samples), n_keep, n, zd, _lib.ptr(local), _lib.ptr(rowend),
                  _lib.ptr(zlist), st)
        heads = self._workspace('heads', n_distinct * n_x * 2, torch).view(n_distinct, n_x, 2)
        _lib.call("bgm_causal_effect_heads", m, _lib.ptr(zlist), n_distinct, _lib.ptr(xv), n_x, _lib.ptr(heads), st)
        _lib.call("bgm_causal_effect_combine", m, _lib.ptr(heads), _lib.ptr(local), _lib.ptr(rowend), n_keep, n, n_x,
                  int(bool(sample_y)), seed, int(row_offset), _lib.ptr(nz), None if binary else _lib.ptr(out),
                  _lib.ptr(out) if binary else None, st)
        self.last_distinct_fraction = n_distinct / float(total)
        return out

    def _dose_grid(self, x_values, torch):
        """x_values on the device; the last grid is kept (a copy from pageable memory is a synchronous driver call)."""
        xa = np.ascontiguousarray(np.asarray(x_values, dtype=np.float32))
        key = xa.tobytes()
        cached = getattr(self, '_xv_cache', None)
        if cached is None or cached[0] != key:
            cached = self._xv_cache = (key, torch.tensor(xa, device='cuda'))
        return cached[1]

    def _workspace(self, name, numel, torch):
        """float32 device workspace of at least `numel` elements, kept on the model and grown by >= 1.3x."""
        ws = getattr(self, '_ws', None)
        if ws is None:
            ws = self._ws = {}
        t = ws.get(name)
        if t is None or t.numel() < numel:
            ws[name] = None
            t = ws[name] = torch.empty(int(max(numel, 1) * 1.3) + 1024, dtype=torch.float32, device='cuda')
        return t[:numel]

    def infer_from_latent_posterior(self, data_posterior_z, x_values=None, sample_y=True, eps=1e-6, *,
                                    seed=None, noise=None):
        """causalbgm/base.py:671-763.  Binary: ITE (n_keep, n); continuous: ADRF draws
        (len(x_values), n_keep).  `noise` injects the N(0,1) draws of :704/:725/:753
        (binary (2,n_keep,n): x=1 then x=0; continuous (len(x_values), n_keep, n))."""
        torch = _lib.require_cuda()
        zs = self._to_device(data_posterior_z, torch)
        n_keep, n, zd = zs.shape
        if seed is None:
            seed = int(np.random.randint(0, 2 ** 31 - 1))
        if self._p['binary_treatment']:
            return self._effect_device(zs, n_keep, n, None, sample_y, seed, 0, noise).cpu().numpy()
        x_values = np.atleast_1d(np.asarray(x_values, dtype=float))
        sums = self._effect_device(zs, n_keep, n, x_values, sample_y, seed, 0, noise)
        return (sums / float(n)).float().cpu().numpy()

    def predict(self, data, alpha=0.01, n_mcmc=3000, burn_in=5000, x_values=None, q_sd=1.0, sample_y=True,
                bs=10000, *, seed=None, group=None, row_offset=0, verbose=1):
        """causalbgm/base.py:573-668 -> (effect, posterior interval).

        The kept states never leave the device: the `bs` slices are sampled (several per
        launch when q_sd is fixed) and reduced to ITE draws / ADRF partial sums on the GPU.  Under torch.distributed pass
        `group` (and this rank's global `row_offset`): every rank processes its own
        shard of rows and the ADRF sums are all-reduced once at the end (no collective
        inside the sampling loop).
        """
        assert 0 < alpha < 1, "The significance level 'alpha' must be greater than 0 and less than 1."
        if not self._p['binary_treatment']:
            if x_values is None:
                raise ValueError("For continuous treatment, 'x_values' must not be None. "
                                 "Provide a list or a single treatment value.")
        if x_values is not None:
            x_values = np.array([x_values], dtype=float) if np.isscalar(x_values) else np.array(x_values, dtype=float)
        torch = _lib.require_cuda()
        data_x, data_y, data_v = data
        n_test = len(data_x)
        bs = max(1, int(bs))
        if seed is None:
            seed = int(np.random.randint(0, 2 ** 31 - 1)) * (2 ** 31) + int(np.random.randint(0, 2 ** 31 - 1))
        if verbose:
            print('MCMC Latent Variable Sampling ...')
        adaptive = (q_sd is None or q_sd <= 0)
        binary = self._p['binary_treatment']
        if binary:
            ite_mean = np.zeros(n_test, dtype=np.float32)
            upper = np.zeros(n_test, dtype=np.float32)
            lower = np.zeros(n_test, dtype=np.float32)
        else:
            sums = torch.zeros((len(x_values), n_mcmc), dtype=torch.float64, device='cuda')
            n_seen = 0
        # The reference runs one independent MH per `bs` slice (:630, :650).  With a fixed q_sd the
        # chains of different slices do not interact and the Philox noise is keyed by the global
        # row, so several slices are sampled in ONE launch (identical results, a full grid instead
        # of ceil(bs/128) row tiles); with the adaptive rule the acceptance window is per slice.
        step_rows = bs
        if not adaptive and not self._bnn:     # Bayesian nets: the rows of a slice share batch statistics
            zd_ = sum(self._p['z_dims'])
            n_x_ = 2 if binary else len(x_values)
            # kept states + effect draws + the memoised path's index arrays and (worst case) heads
            per_row = float(n_mcmc) * (4.0 * (zd_ + 2) + 8.0 + 8.0 * n_x_)
            budget = min(24e9, 0.3 * _lib.free_memory_estimate(torch))
            step_rows = max(bs, int(budget // per_row) // bs * bs)
        acc_tail = []
        for start in range(0, n_test, step_rows):
            end = min(start + step_rows, n_test)
            _, x, y, v, ldv, n = self._stage((data_x[start:end], data_y[start:end], data_v[start:end]))
            r = self._mh_device(x, y, v, ldv, n, int(burn_in), int(n_mcmc), q_sd, adaptive, 1.0, 0.25, 0.05,
                                50, 100, seed, row_offset + start, slice_id=(row_offset + start) // bs)
            eff = self._effect_device(r['samples'], int(n_mcmc), n, x_values, sample_y, seed,
                                      row_offset + start)
            T_ = int(burn_in) + int(n_mcmc)
            w_ = min(100, T_)
            acc_tail.append((r['accept_count'][T_ - w_:].sum(), w_ * n))                # :901, read back once below
            if binary:                                                            # :640-642
                ite_mean[start:end], lower[start:end], upper[start:end] = _ite_summary(torch, eff, alpha)
            else:                                                                 # :660-661
                sums += eff
                n_seen += n
        if acc_tail:
            tot = torch.stack([a for a, _ in acc_tail]).sum()
            self.last_acceptance_rate = float(tot.item()) / float(sum(c for _, c in acc_tail))
            if verbose:
                print(f"Final MCMC Acceptance Rate: {self.last_acceptance_rate:.4f}")
        if binary:
            return ite_mean, np.stack([lower, upper], axis=1)
        ce = merge_adrf(sums, n_seen, group) if group is not None else \
            (sums / float(n_seen)).float().cpu().numpy()                          # :663
        return finish_adrf(ce, alpha)                                             # :665-667

    def philox_noise(self, seed, n, T, row_offset=0):
        """The exact noise `metropolis_hastings_sampler(seed=...)` draws in-kernel, as
        host arrays dict(z0, eps, u) -- lets a Philox run be replayed through any other
        implementation of the algorithm (used by the parity tests)."""
        torch = _lib.require_cuda()
        zd = sum(self._p['z_dims'])
        z0 = torch.empty((n, zd), dtype=torch.float32, device='cuda')
        eps = torch.empty((T, n, zd), dtype=torch.float32, device='cuda')
        u = torch.empty((T, n), dtype=torch.float64, device='cuda')
        _lib.call("bgm_mh_noise", int(seed) & (2 ** 64 - 1), int(row_offset), n, zd, 0, T, _lib.ptr(z0),
                  _lib.ptr(eps), _lib.ptr(u), _lib.stream_ptr())
        return dict(z0=z0.cpu().numpy(), eps=eps.cpu().numpy(), u=u.cpu().numpy())

    # ------------------------------------------------- fit: EGM + iterative phase
    @staticmethod
    def _save_data(fname, data, delimiter='\t'):
        """utils/data_io.py:8-31."""
        if fname.endswith('.npy'):
            np.save(fname, data)
        elif fname.endswith('.txt') or fname.endswith('.csv'):
            np.savetxt(fname, data, fmt='%.6f', delimiter=delimiter)
        else:
            raise ValueError("Wrong saving format, please specify either .npy, .txt, or .csv")

    @property
    def data_z(self):
        """The latent table of `fit` (:484) as a host array."""
        z = getattr(self, '_data_z', None)
        return None if z is None else z.cpu().numpy()

    def _sigmas(self):
        return [float(self._p[k]) if k in self._p else -1.0 for k in ('sigma_v', 'sigma_x', 'sigma_y')]

    def evaluate(self, data, data_z=None, nb_intervals=200):
        """causalbgm/base.py:534-570 -> (causal_pre, mse_x, mse_y, mse_v): ITE (n,1) for a binary
        treatment, else the ADRF on `nb_intervals` doses between the 5th and 95th percentile of x."""
        torch, x, y, v, ldv, n = self._stage(data)
        p, zd = self._p['v_dim'], sum(self._p['z_dims'])
        vc = v[:, :p].contiguous() if ldv != p else v
        tr = self._device_trainer()
        sums = torch.zeros(3, dtype=torch.float64, device='cuda')
        if data_z is None:
            z = torch.empty((n, zd), dtype=torch.float32, device='cuda')
            _lib.call(self._tfn("bgm_causal_evaluate"), tr, None, _lib.ptr(x), _lib.ptr(y), _lib.ptr(vc), n, _lib.ptr(sums),
                      _lib.ptr(z), _lib.stream_ptr())
        else:
            z = self._to_device(data_z, torch)
            _lib.call(self._tfn("bgm_causal_evaluate"), tr, _lib.ptr(z), _lib.ptr(x), _lib.ptr(y), _lib.ptr(vc), n,
                      _lib.ptr(sums), None, _lib.stream_ptr())
        s = sums.cpu().numpy()
        mse_v, mse_x, mse_y = np.float32(s[0] / (n * p)), np.float32(s[1] / n), np.float32(s[2] / n)
        zs = z.reshape(1, n, zd)
        if self._p['binary_treatment']:
            ite = self._effect_device(zs, 1, n, None, False, 0, 0)
            return ite.reshape(n, 1).cpu().numpy(), mse_x, mse_y, mse_v
        xs = np.sort(np.asarray(x.cpu().numpy()).ravel())                       # tfp.stats.percentile, 'nearest'
        x_min = xs[int(np.round((n - 1) * 0.05))]
        x_max = xs[int(np.round((n - 1) * 0.95))]
        x_values = np.linspace(x_min, x_max, nb_intervals).astype(np.float32)
        adrf = self._effect_device(zs, 1, n, x_values, False, 0, 0)
        return (adrf[:, 0] / float(n)).float().cpu().numpy(), mse_x, mse_y, mse_v

    def fit(self, data, epochs=100, epochs_per_eval=5, batch_size=32, startoff=0, use_egm_init=True,
            egm_n_iter=30000, egm_batches_per_eval=500, save_format='txt', verbose=1):
        """causalbgm/base.py:434-532: optional EGM warm-start, latent table from e(V) (or N(0,1)),
        then `epochs+1` epochs of mini-batch updates of g, h, f and of the latent rows (dense
        Keras-Adam sweep), evaluating every `epochs_per_eval` epochs.  Data, latent table and all
        optimizer state stay on the device; the epoch permutation comes from NumPy's global
        generator (`np.random.choice(n, n, replace=False)`, :489) -- bit-exact index stream."""
        torch = _lib.require_cuda()
        data_x, data_y, data_v = data
        n = len(data_x)
        p, zd = self._p['v_dim'], sum(self._p['z_dims'])
        bs = int(batch_size)
        if bs > 32:
            # the fused single-CTA kernels take at most 32 rows: larger mini-batches train on the layered engine
            # (the WGAN-GP discriminator step through csrc/disc_big.cuh)
            self._set_layered(True)
        if self._p['save_res']:
            with open('{}/params.txt'.format(self.save_dir), 'w') as f_params:
                f_params.write(str(self.params))
        xd = self._to_device(data_x, torch).reshape(-1).contiguous()
        yd = self._to_device(data_y, torch).reshape(-1).contiguous()
        vd = self._to_device(data_v, torch).contiguous()
        tr = self._device_trainer()
        st = _lib.stream_ptr()
        if use_egm_init:
            self.egm_init(data, egm_n_iter=egm_n_iter, egm_batches_per_eval=egm_batches_per_eval,
                          batch_size=batch_size, verbose=verbose)
            if verbose:
                print('Initialize latent variables Z with e(V)...')
            z = torch.empty((n, zd), dtype=torch.float32, device='cuda')                    # :479
            sums = torch.zeros(3, dtype=torch.float64, device='cuda')
            _lib.call(self._tfn("bgm_causal_evaluate"), tr, None, _lib.ptr(xd), _lib.ptr(yd), _lib.ptr(vd), n, _lib.ptr(sums),
                      _lib.ptr(z), st)
        else:
            if verbose:
                print('Random initialization of latent variables Z...')
            z = torch.from_numpy(np.random.normal(0, 1, size=(n, zd)).astype('float32')).cuda()   # :482
        self._data_z = z
        m_z, v_z = torch.zeros_like(z), torch.zeros_like(z)
        slot = torch.full((n,), -1, dtype=torch.int32, device='cuda')
        sig = self._sigmas()
        _lib.call(self._tfn("bgm_trainer_set_iter"), tr, float(self._p['lr_theta']), float(self._p['lr_z']), sig[0], sig[1], sig[2])
        nl = torch.zeros(6, dtype=torch.float32, device='cuda')
        zl = torch.zeros(1, dtype=torch.float32, device='cuda')
        best_loss = np.inf
        if verbose:
            print('Iterative Updating Starts ...')
        for epoch in range(int(epochs) + 1):
            sample_idx = np.random.choice(n, n, replace=False)                               # :489
            idx_d = torch.from_numpy(sample_idx.astype(np.int32)).cuda()
            base = idx_d.data_ptr()
            for i in range(0, n, bs):
                b = min(bs, n - i)
                ip = C.c_void_p(base + 4 * i)
                _lib.call(self._tfn("bgm_train_iter_nets"), tr, _lib.ptr(z), _lib.ptr(xd), _lib.ptr(yd), _lib.ptr(vd), ip, b, 1,
                          1.0, _lib.ptr(nl), st)                                             # :500-502
                if self._layered:
                    _lib.call("bgm_lt_iter_latent", tr, _lib.ptr(z), _lib.ptr(m_z), _lib.ptr(v_z), _lib.ptr(slot), n,
                              _lib.ptr(xd), _lib.ptr(yd), _lib.ptr(vd), ip, b, _lib.ptr(zl), None, st)
                else:
                    _lib.call("bgm_train_iter_latent", tr, _lib.ptr(z), _lib.ptr(m_z), _lib.ptr(v_z), _lib.ptr(slot), n,
                              _lib.ptr(xd), _lib.ptr(yd), _lib.ptr(vd), ip, b, _lib.ptr(zl), st)     # :505
            self._trainer_dirty = True
            if epoch % epochs_per_eval == 0:                                                 # :517-532
                causal_pre, mse_x, mse_y, mse_v = self.evaluate(data=(xd, yd, vd), data_z=z)
                if verbose:
                    print('Epoch [%d/%d]: MSE_x: %.4f, MSE_y: %.4f, MSE_v: %.4f\n' % (epoch, epochs, mse_x, mse_y, mse_v))
                if epoch >= startoff and mse_y < best_loss:
                    best_loss = mse_y
                    self.best_causal_pre = causal_pre
                    self.best_epoch = epoch
                    if self._p['save_model']:                                                # :527-530
                        self.save_weights('{}/weights_at_{}.npz'.format(self.checkpoint_path, epoch))
                if self._p['save_res']:
                    self._save_data('{}/causal_pre_at_{}.{}'.format(self.save_dir, epoch, save_format), causal_pre)
        self.last_iter_losses = tuple(float(a) for a in nl.cpu().numpy()) + (float(zl.cpu()[0]),)

    # ------------------------------------------------------------ EGM training
    def _offset_streams(self, group):
        """Data-parallel contract (`group=`): every rank trains on its own mini-batches, so the host
        streams must differ per rank -- `Gaussian_sampler.__init__` reseeds NumPy's global generator to
        1024 on EVERY rank (prior_samplers.py:24), which would make all ranks draw the same indices,
        prior z and epsilon and turn the gradient all-reduce into a no-op.  Rank 0 keeps the single-GPU
        streams; rank r > 0 reseeds the global generator and its private epsilon generator once per
        model with a rank-dependent seed."""
        import torch.distributed as dist
        r = dist.get_rank(group)
        if r > 0 and not getattr(self, '_streams_offset', False):
            np.random.seed((1024 + 7919 * r) % (2 ** 32))
            self._eps_rng = np.random.RandomState((self._eps_rng.randint(0, 2 ** 31 - 1) + 104729 * r) % (2 ** 32))
        self._streams_offset = True

    def _grad_tensor(self, group):
        """torch view (no copy) of the trainer's flat gradient buffer, for all-reduce."""
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

    def gradients(self, which, data_z, data_v, data_x=None, data_y=None, epsilon=0.5):
        """(losses, flat gradient) of one step WITHOUT the optimizer update; `which` is
        'disc' (dz_net, Keras trainable_variables order) or 'gen' (g|e|f|h).  Test hook."""
        torch = _lib.require_cuda()
        z = self._to_device(data_z, torch)
        v = self._to_device(data_v, torch)
        if which == 'disc':
            losses = torch.empty(2, dtype=torch.float32, device='cuda')
            _lib.call(self._tfn("bgm_train_disc_grad"), self._device_trainer(), _lib.ptr(z), _lib.ptr(v), z.shape[0],
                      float(epsilon), 10.0, _lib.ptr(losses), _lib.stream_ptr())
            return losses.cpu().numpy(), self._grad_tensor(1).cpu().numpy()
        x = self._to_device(data_x, torch).reshape(-1)
        y = self._to_device(data_y, torch).reshape(-1)
        losses = torch.empty(6, dtype=torch.float32, device='cuda')
        _lib.call(self._tfn("bgm_train_gen_grad"), self._device_trainer(), _lib.ptr(z), _lib.ptr(v), _lib.ptr(x), _lib.ptr(y),
                  z.shape[0], _lib.ptr(losses), _lib.stream_ptr())
        return losses.cpu().numpy(), self._grad_tensor(0).cpu().numpy()

    def set_noise_counter(self, counter):
        """Layered engine: the step counter that keys the network noise of the next training step
        (call ids 16*counter + k, csrc/layered_api.cuh) -- lets tests replay a step through the oracle."""
        self._set_layered(True)
        _lib.call("bgm_lt_set_call", self._device_trainer(), int(counter) & 0xFFFFFFFF)

    def iter_gradients(self, data_z_table, data, batch_idx):
        """(losses[6], flat group-0 gradient) of update_g/h/f_net (:156-243) on the rows `batch_idx`,
        WITHOUT the optimizer updates.  Test hook."""
        torch = _lib.require_cuda()
        data_x, data_y, data_v = data
        z = self._to_device(data_z_table, torch).contiguous()
        xd = self._to_device(data_x, torch).reshape(-1).contiguous()
        yd = self._to_device(data_y, torch).reshape(-1).contiguous()
        vd = self._to_device(data_v, torch).contiguous()
        idx = torch.from_numpy(np.asarray(batch_idx, np.int32)).cuda()
        tr = self._device_trainer()
        sig = self._sigmas()
        if not getattr(self, '_iter_ready', None) is tr:
            _lib.call(self._tfn("bgm_trainer_set_iter"), tr, float(self._p['lr_theta']), float(self._p['lr_z']), sig[0],
                      sig[1], sig[2])
            self._iter_ready = tr
        nl = torch.zeros(6, dtype=torch.float32, device='cuda')
        _lib.call(self._tfn("bgm_train_iter_nets"), tr, _lib.ptr(z), _lib.ptr(xd), _lib.ptr(yd), _lib.ptr(vd), _lib.ptr(idx),
                  len(batch_idx), 0, 1.0, _lib.ptr(nl), _lib.stream_ptr())
        return nl.cpu().numpy(), self._grad_tensor(0).cpu().numpy()

    def latent_step(self, data_z_table, data, batch_idx):
        """One update_latent_variable_sgd (:246-302) on copies of the given arrays (layered engine):
        returns (loss_postrior_z, gradient rows (bs, zd), updated latent table).  Test hook."""
        torch = _lib.require_cuda()
        self._set_layered(True)
        data_x, data_y, data_v = data
        z = self._to_device(data_z_table, torch).contiguous().clone()
        n, zd = z.shape
        xd = self._to_device(data_x, torch).reshape(-1).contiguous()
        yd = self._to_device(data_y, torch).reshape(-1).contiguous()
        vd = self._to_device(data_v, torch).contiguous()
        idx = torch.from_numpy(np.asarray(batch_idx, np.int32)).cuda()
        tr = self._device_trainer()
        sig = self._sigmas()
        if not getattr(self, '_iter_ready', None) is tr:
            _lib.call("bgm_lt_set_iter", tr, float(self._p['lr_theta']), float(self._p['lr_z']), sig[0], sig[1], sig[2])
            self._iter_ready = tr
        m_z, v_z = torch.zeros_like(z), torch.zeros_like(z)
        slot = torch.full((n,), -1, dtype=torch.int32, device='cuda')
        zl = torch.zeros(1, dtype=torch.float32, device='cuda')
        gz = torch.zeros((len(batch_idx), zd), dtype=torch.float32, device='cuda')
        _lib.call("bgm_lt_iter_latent", tr, _lib.ptr(z), _lib.ptr(m_z), _lib.ptr(v_z), _lib.ptr(slot), n, _lib.ptr(xd),
                  _lib.ptr(yd), _lib.ptr(vd), _lib.ptr(idx), len(batch_idx), _lib.ptr(zl), _lib.ptr(gz), _lib.stream_ptr())
        assert int((slot != -1).sum().item()) == 0
        return float(zl.cpu()[0]), gz.cpu().numpy(), z.cpu().numpy()

    def get_weights(self):
        """dict of Keras-layout weight lists of g, e, f, h and dz (trainable_variables order)."""
        self._sync_from_trainer()
        return dict(g=self.g_net.get_weights(), e=self.e_net.get_weights(), f=self.f_net.get_weights(),
                    h=self.h_net.get_weights(), dz=[a.copy() for a in self.dz_net.trainable_list()])

    def save_weights(self, path):
        """All network weights (Keras-layout arrays) into one .npz -- the counterpart of the
        reference's tf.train.Checkpoint (:100-113, :527-530); `load_weights` restores them."""
        w = self.get_weights()
        np.savez(path, **{"%s_%d" % (k, i): a for k, arrs in w.items() for i, a in enumerate(arrs)})

    def load_weights(self, path):
        z = np.load(path)
        got = {}
        for k in ('g', 'e', 'f', 'h', 'dz'):
            got[k] = [z["%s_%d" % (k, i)] for i in range(sum(1 for name in z.files if name.startswith(k + "_")))]
        self.set_weights(**got)

    def load_tf_checkpoint(self, path):
        """Restores the nets from a checkpoint written by the reference (`tf.train.Checkpoint`, :112-127): `path` is a
        checkpoint prefix (`.../ckpt-5`) or the directory a CheckpointManager wrote.  Deterministic nets only; see
        bayesgm_b200/tf_checkpoint.py for the format reader and its validation status."""
        from .tf_checkpoint import load_tf_checkpoint
        return load_tf_checkpoint(self, path)

    def train_disc_step(self, data_z, data_v, *, epsilon=None, group=None):
        """causalbgm/base.py:305-330 -> (dz_loss, d_loss).  `epsilon` is the U(0,1) draw of
        :307 (TensorFlow's stream in the reference; here a private RandomState unless given).
        Under torch.distributed pass `group`: gradients are averaged over the ranks."""
        torch = _lib.require_cuda()
        z = self._to_device(data_z, torch)
        v = self._to_device(data_v, torch)
        bs = z.shape[0]
        if epsilon is None:
            epsilon = float(self._eps_rng.uniform())
        losses = torch.empty(2, dtype=torch.float32, device='cuda')
        _lib.call(self._tfn("bgm_train_disc_grad"), self._device_trainer(), _lib.ptr(z), _lib.ptr(v), bs, float(epsilon), 10.0,
                  _lib.ptr(losses), _lib.stream_ptr())
        self._apply(1, group)
        l = losses.cpu().numpy()
        return float(l[0]), float(l[1])

    def train_gen_step(self, data_z, data_v, data_x, data_y, *, group=None):
        """causalbgm/base.py:332-377 -> (e_loss_adv, l2_loss_v, l2_loss_z, l2_loss_x, l2_loss_y, g_e_loss)."""
        torch = _lib.require_cuda()
        z = self._to_device(data_z, torch)
        v = self._to_device(data_v, torch)
        x = self._to_device(data_x, torch).reshape(-1)
        y = self._to_device(data_y, torch).reshape(-1)
        losses = torch.empty(6, dtype=torch.float32, device='cuda')
        _lib.call(self._tfn("bgm_train_gen_grad"), self._device_trainer(), _lib.ptr(z), _lib.ptr(v), _lib.ptr(x), _lib.ptr(y),
                  z.shape[0], _lib.ptr(losses), _lib.stream_ptr())
        self._apply(0, group)
        return tuple(float(a) for a in losses.cpu().numpy())

    def egm_init(self, data, egm_n_iter=30000, batch_size=32, egm_batches_per_eval=500, verbose=1, *,
                 group=None, chunk=64, eval_during=True, index_stream='numpy'):
        """causalbgm/base.py:380-431.  The data set stays on the device; mini-batch indices
        and prior draws come from NumPy's global generator in the reference's exact call
        order (g_d_freq x [choice, get_batch], then [get_batch, choice]) -- bit-exact index
        streams -- generated `chunk` iterations ahead and uploaded in one copy; the batches
        are gathered on the device.  Returns the last (dz_loss, d_loss) and generator losses.
        Every `egm_batches_per_eval` iterations the model is evaluated like :425-430 (`eval_during=False`
        skips it); the (iteration, mse_x, mse_y, mse_v) history is kept in `self.egm_history`.
        `index_stream='floyd'` draws the mini-batches by subset sampling instead (same distribution,
        O(batch) instead of O(n) host work per draw, NOT NumPy's legacy stream; _hostrng.FloydProducer)."""
        torch = _lib.require_cuda()
        if index_stream not in ('numpy', 'floyd'):
            raise ValueError("index_stream must be 'numpy' or 'floyd'")
        if group is not None:
            self._offset_streams(group)
        data_x, data_y, data_v = data
        n = len(data_x)
        p, zd = self._p['v_dim'], sum(self._p['z_dims'])
        freq = int(self._p['g_d_freq'])
        bs = int(batch_size)
        if bs > 32:          # more than the fused kernels' 32 rows: layered engine, WGAN-GP step through csrc/disc_big.cuh
            self._set_layered(True)
        xd = self._to_device(data_x, torch).reshape(-1).contiguous()
        yd = self._to_device(data_y, torch).reshape(-1).contiguous()
        vd = self._to_device(data_v, torch).contiguous()
        tr = self._device_trainer()
        st = _lib.stream_ptr()
        dloss = torch.zeros(2, dtype=torch.float32, device='cuda')
        gloss = torch.zeros(6, dtype=torch.float32, device='cuda')
        bz = torch.empty((bs, zd), dtype=torch.float32, device='cuda')
        bv = torch.empty((bs, p), dtype=torch.float32, device='cuda')
        bx = torch.empty(bs, dtype=torch.float32, device='cuda')
        by = torch.empty(bs, dtype=torch.float32, device='cuda')
        if verbose:
            print('EGM Initialization Starts ...')
        self.egm_history = []
        total = int(egm_n_iter) + 1
        it = 0
        # NumPy's global generator is continued natively on a background thread (csrc/host_rng.cu): the
        # reference's draw order -- g_d_freq x [choice(n, bs) :406, get_batch :407], then [get_batch :412,
        # choice :413] -- bit-exact, `chunk` iterations per hand-over; the state goes back into np.random
        # when the loop ends (`choice` without replacement permutes all n indices per call: at n >= 1e5 that
        # is more host time than the training step it feeds)
        from ._hostrng import EgmProducer, FloydProducer
        prod = (EgmProducer if index_stream == 'numpy' else FloydProducer)(n, bs, zd, freq, total, chunk=int(chunk))
        try:
            self._egm_loop(prod, total, freq, bs, p, zd, xd, yd, vd, tr, st, dloss, gloss, bz, bv, bx, by, group,
                           egm_batches_per_eval, verbose, eval_during, torch)
        finally:
            prod.close(drain=True)
        if verbose:
            print('EGM Initialization Ends.')
        d, g = dloss.cpu().numpy(), gloss.cpu().numpy()
        return (float(d[0]), float(d[1])), tuple(float(a) for a in g)

    def _egm_loop(self, prod, total, freq, bs, p, zd, xd, yd, vd, tr, st, dloss, gloss, bz, bv, bx, by, group,
                  egm_batches_per_eval, verbose, eval_during, torch):
        it = 0
        while it < total:
            idx, zz = prod.get()
            cnt = idx.shape[0]
            eps = np.array([[self._eps_rng.uniform() for _ in range(freq)] for _ in range(cnt)], np.float32).reshape(cnt, freq)
            idx_d = torch.from_numpy(idx).cuda()
            zz_d = torch.from_numpy(zz).cuda()
            for c in range(cnt):
                for k in range(freq):
                    ip = C.c_void_p(idx_d[c, k].data_ptr())
                    _lib.call("bgm_gather_rows", _lib.ptr(vd), p, ip, bs, p, _lib.ptr(bv), st)
                    _lib.call(self._tfn("bgm_train_disc_grad"), tr, C.c_void_p(zz_d[c, k].data_ptr()), _lib.ptr(bv), bs,
                              float(eps[c, k]), 10.0, _lib.ptr(dloss), st)
                    self._apply(1, group)
                ip = C.c_void_p(idx_d[c, freq].data_ptr())
                _lib.call("bgm_gather_rows", _lib.ptr(vd), p, ip, bs, p, _lib.ptr(bv), st)
                _lib.call("bgm_gather_rows", _lib.ptr(xd), 1, ip, bs, 1, _lib.ptr(bx), st)
                _lib.call("bgm_gather_rows", _lib.ptr(yd), 1, ip, bs, 1, _lib.ptr(by), st)
                _lib.call(self._tfn("bgm_train_gen_grad"), tr, C.c_void_p(zz_d[c, freq].data_ptr()), _lib.ptr(bv), _lib.ptr(bx),
                          _lib.ptr(by), bs, _lib.ptr(gloss), st)
                self._apply(0, group)
                if (it + c) % egm_batches_per_eval == 0:                                     # :418-430
                    if verbose:
                        d, g = dloss.cpu().numpy(), gloss.cpu().numpy()
                        print('EGM Initialization Iter [%d] : e_loss_adv [%.4f], l2_loss_v [%.4f], l2_loss_z [%.4f], '
                              'l2_loss_x [%.4f], l2_loss_y [%.4f], g_e_loss [%.4f], dz_loss [%.4f], d_loss [%.4f]'
                              % (it + c, g[0], g[1], g[2], g[3], g[4], g[5], d[0], d[1]))
                    if eval_during:
                        self._trainer_dirty = True
                        causal_pre, mse_x, mse_y, mse_v = self.evaluate(data=(xd, yd, vd))
                        self.egm_history.append((it + c, float(mse_x), float(mse_y), float(mse_v)))
                        if self._p['save_res']:
                            self._save_data('{}/causal_pre_egm_init_iter-{}.txt'.format(self.save_dir, it + c), causal_pre)
            it += cnt
