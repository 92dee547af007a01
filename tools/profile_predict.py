"""Host-side profile of CausalBGM.predict at the bench workload (cProfile) + a synchronised stage breakdown."""
import cProfile, pstats, io, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from bayesgm_b200 import CausalBGM

x, y, v = bench.make_data(0)
m = CausalBGM(params=bench.params(), random_seed=123)
xh, yh, vh = [torch.from_numpy(np.ascontiguousarray(a)).pin_memory() for a in (x, y, v)]
kw = dict(alpha=0.01, n_mcmc=500, burn_in=500, x_values=bench.X_VALUES, q_sd=1.0, sample_y=True, bs=100000, verbose=0)
for i in range(3):
    m.predict((xh, yh, vh), seed=i, **kw)
torch.cuda.synchronize()
t0 = time.perf_counter()
for i in range(5):
    m.predict((xh, yh, vh), seed=10 + i, **kw)
print("predict ms/step", (time.perf_counter() - t0) / 5 * 1e3, "OMP", os.environ.get("OMP_NUM_THREADS"), "torch threads", torch.get_num_threads())
pr = cProfile.Profile()
pr.enable()
for i in range(5):
    m.predict((xh, yh, vh), seed=20 + i, **kw)
pr.disable()
s = io.StringIO()
pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(14)
print(s.getvalue()[:3500])
# stage breakdown with synchronisation
def stage(name, fn):
    torch.cuda.synchronize(); t = time.perf_counter(); r = fn(); torch.cuda.synchronize()
    print("%-28s %.2f ms" % (name, (time.perf_counter() - t) * 1e3)); return r
_, xd, yd, vd, ldv, n = stage("H2D (_stage)", lambda: m._stage((xh, yh, vh)))
aux = stage("project (_aux)", lambda: m._aux(vd, ldv, n))
r = stage("sampler", lambda: m._mh_device(xd, yd, vd, ldv, n, 500, 500, 1.0, False, 1.0, 0.25, 0.05, 50, 100, 5, 0, aux=aux))
eff = stage("effect (memoised)", lambda: m._effect_device(r['samples'], 500, n, bench.X_VALUES, True, 5, 0))
from bayesgm_b200.shard import finish_adrf
stage("D2H + quantiles", lambda: finish_adrf((eff / float(n)).float().cpu().numpy(), 0.01))
