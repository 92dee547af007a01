"""Where do the occasional +10..80 ms predict() calls on some boxes come from?  Per call: host time, CPU time of the
process, device time between two events, cgroup throttling counters.  usage: python tools/e2e_jitter.py [calls]"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from bayesgm_b200 import CausalBGM


def cg(name):
    for p in ("/sys/fs/cgroup/" + name, "/sys/fs/cgroup/cpu/" + name):
        try:
            return open(p).read().strip().replace("\n", " ")
        except OSError:
            pass
    return None


calls = int(sys.argv[1]) if len(sys.argv) > 1 else 30
print("cpus", os.cpu_count(), "affinity", len(os.sched_getaffinity(0)), "cpu.max", cg("cpu.max"), "OMP", os.environ.get("OMP_NUM_THREADS"),
      "torch threads", torch.get_num_threads())
print("cpu.stat before", cg("cpu.stat"))
x, y, v = bench.make_data(0)
m = CausalBGM(params=bench.params(), random_seed=123)
xh, yh, vh = [torch.from_numpy(np.ascontiguousarray(a)).pin_memory() for a in (x, y, v)]
kw = dict(alpha=0.01, n_mcmc=500, burn_in=500, x_values=bench.X_VALUES, q_sd=1.0, sample_y=True, bs=100000, verbose=0)
rows = []
for i in range(calls):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a0 = torch.cuda.memory_stats()["num_alloc_retries"], torch.cuda.memory_reserved()
    t0, c0 = time.perf_counter(), time.process_time()
    e0.record()
    m.predict((xh, yh, vh), seed=i, **kw)
    e1.record()
    t1, c1 = time.perf_counter(), time.process_time()
    torch.cuda.synchronize()
    rows.append((1e3 * (t1 - t0), 1e3 * (c1 - c0), e0.elapsed_time(e1), torch.cuda.memory_reserved() - a0[1]))
for i, r in enumerate(rows):
    print("call %2d host %.1f ms  cpu %.1f ms  device %.1f ms  reserved %+d" % ((i,) + r))
print("cpu.stat after", cg("cpu.stat"))
# the same with the sampler alone (device-resident inputs)
_, xd, yd, vd, ldv, n = m._stage((xh, yh, vh))
aux = m._aux(vd, ldv, n)
ts = []
for i in range(10):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    m._mh_device(xd, yd, vd, ldv, n, 500, 500, 1.0, False, 1.0, 0.25, 0.05, 50, 100, seed=i, row_offset=0, aux=aux)
    torch.cuda.synchronize(); ts.append(1e3 * (time.perf_counter() - t0))
print("sampler alone ms", [round(t, 1) for t in ts])
# stage by stage, synchronised, to see which stage the slow calls stretch
from bayesgm_b200.shard import finish_adrf
def stage(fn):
    torch.cuda.synchronize(); t = time.perf_counter(); r = fn(); torch.cuda.synchronize()
    return r, 1e3 * (time.perf_counter() - t)
print("stages: H2D, project, sampler, effect, D2H+quantiles (ms)")
for i in range(calls):
    (_, xd, yd, vd, ldv, n), t_h2d = stage(lambda: m._stage((xh, yh, vh)))
    aux, t_aux = stage(lambda: m._aux(vd, ldv, n))
    r, t_mh = stage(lambda: m._mh_device(xd, yd, vd, ldv, n, 500, 500, 1.0, False, 1.0, 0.25, 0.05, 50, 100, 5, 0, aux=aux))
    eff, t_eff = stage(lambda: m._effect_device(r['samples'], 500, n, bench.X_VALUES, True, 5, 0))
    _, t_fin = stage(lambda: finish_adrf((eff / float(n)).float().cpu().numpy(), 0.01))
    print("call %2d  %.1f  %.1f  %.1f  %.1f  %.1f" % (i, t_h2d, t_aux, t_mh, t_eff, t_fin))
# the suspect itself: cudaMemGetInfo latency on this box
lat = []
for i in range(300):
    t = time.perf_counter(); torch.cuda.mem_get_info(); lat.append(1e3 * (time.perf_counter() - t))
    time.sleep(0.005)
lat = np.array(lat)
print("cudaMemGetInfo x300: median %.3f ms  p90 %.3f  max %.3f  calls > 5 ms: %d" % (np.median(lat), np.percentile(lat, 90), lat.max(), int((lat > 5).sum())))
