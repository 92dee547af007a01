"""The one HBM-bound kernel of the path: Keras Adam on a gathered variable = a dense sweep over the whole latent
table per mini-batch (latent_adam_sweep_kernel).  Times bgm_train_iter_latent at n rows and reports the
sweep's share and bandwidth: algorithmic traffic = 6 floats per table element (read and write z, m, v) + 4 bytes of
slot per row."""
import ctypes as C, json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from bayesgm_b200 import CausalBGM, _lib

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4000000
zd_dims = [1, 1, 1, 7]
zd = sum(zd_dims)
p = 200
m = CausalBGM(params=bench.params(zd_dims), random_seed=1)
rs = np.random.RandomState(0)
nd = 4096                      # the data rows the batch indices point at; the TABLE has n rows
xd = torch.from_numpy(rs.standard_normal(n).astype(np.float32)).cuda()
yd = torch.from_numpy(rs.standard_normal(n).astype(np.float32)).cuda()
vd = torch.zeros((nd, p), dtype=torch.float32, device='cuda')
z = torch.randn((n, zd), device='cuda')
m_z, v_z = torch.zeros_like(z), torch.zeros_like(z)
slot = torch.full((n,), -1, dtype=torch.int32, device='cuda')
tr = m._device_trainer()
_lib.call("bgm_trainer_set_iter", tr, 1e-4, 1e-4, -1.0, -1.0, -1.0)
zl = torch.zeros(1, dtype=torch.float32, device='cuda')
st = _lib.stream_ptr()
idx = torch.from_numpy(rs.randint(0, nd, size=(64, 32)).astype(np.int32)).cuda()

def step(i):
    _lib.call("bgm_train_iter_latent", tr, _lib.ptr(z), _lib.ptr(m_z), _lib.ptr(v_z), _lib.ptr(slot), n, _lib.ptr(xd), _lib.ptr(yd),
              _lib.ptr(vd), C.c_void_p(idx[i % 64].data_ptr()), 32, _lib.ptr(zl), st)
for i in range(5):
    step(i)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
reps = 50
e0.record()
for i in range(reps):
    step(i)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
# the same step on a tiny table = everything but the sweep
n_small = 4096
z2 = torch.randn((n_small, zd), device='cuda'); m2, v2 = torch.zeros_like(z2), torch.zeros_like(z2)
s2 = torch.full((n_small,), -1, dtype=torch.int32, device='cuda')
def step_small(i):
    _lib.call("bgm_train_iter_latent", tr, _lib.ptr(z2), _lib.ptr(m2), _lib.ptr(v2), _lib.ptr(s2), n_small, _lib.ptr(xd), _lib.ptr(yd),
              _lib.ptr(vd), C.c_void_p(idx[i % 64].data_ptr()), 32, _lib.ptr(zl), st)
for i in range(5):
    step_small(i)
torch.cuda.synchronize()
e0.record()
for i in range(reps):
    step_small(i)
e1.record()
torch.cuda.synchronize()
ms_small = e0.elapsed_time(e1) / reps
bytes_alg = n * zd * 24 + n * 4
peaks = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json"))) \
    if os.path.exists(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")) else {}
# (the difference of the two step times also carries launch gaps of the step's other kernels: the kernel's own
# duration comes from `ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum -k regex:latent_adam_sweep`)
print(json.dumps({"kernel": "latent_adam_sweep_kernel", "table_rows": n, "zd": zd, "ms_per_latent_step": ms,
                  "ms_per_latent_step_small_table": ms_small, "sweep_ms": ms - ms_small,
                  "algorithmic_bytes": bytes_alg, "achieved_GBps_lower_bound": bytes_alg / ((ms - ms_small) * 1e-3) / 1e9,
                  "measured_peaks": peaks}))
