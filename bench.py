#!/usr/bin/env python
"""Benchmark of the posterior-sampling hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config cfg3|cfg2|cfg4|cfg5|cfg3bnn]

Default workload = BASELINE.json configs[2] (the one the metric is quoted on): CausalBGM,
continuous treatment, Sim_Hirano_Imbens-shaped synthetic data, n=100000 rows PER GPU, p=200
covariates, z_dims=[1,1,1,2] (z_dim=5), deterministic nets (use_bnn=False), glorot-uniform weights
from RandomState(123); one STEP = one random-walk MH run of T = 1000 iterations (burn_in 500 + 500
kept) over all rows = n*T posterior samples.

  value : n*T*gpus / step time, inputs resident in HBM, kept states written to HBM (one launch of
          the persistent sampler kernel per step per GPU).
  e2e   : the same through `CausalBGM.predict(data, x_values=linspace(0,3,20))` with HOST (pinned)
          x, y, v: H2D copies, sampler, effect kernels, the ADRF all-reduce (N > 1) and the D2H of
          the result all inside the timed region.
Multi-GPU: one process per GPU (torchrun).  ONE data set of n*N rows is generated (same seed on
every rank) and sharded by rows (`shard_rows`); the sampler needs no collective, `predict` all-reduces
the (20, 500) ADRF partial sums once (NCCL) -- weak scaling; timing = max over ranks.
Other configs (parity-size / secondary workloads, lines committed under profiles/):
  cfg2    binary treatment, ACIC-shaped n=20000 p=100 z_dims=[3,6,3,6]            (samples/s)
  cfg3bnn cfg3 with the shipped Bayesian nets (use_bnn=True), n=20000 per GPU      (samples/s)
  cfg4    EGM training, n=1e6 rows sharded over the GPUs, batch 32 per GPU, NCCL gradient
          all-reduce in every discriminator / generator step                       (EGM iterations/s)
  cfg5    BGM imputation, n=500000/8 rows per GPU, x_dim=500, 30% MCAR, HMC burn_in 100 + 100 kept,
          L=10, shared step size all-reduced per adaptation step, streaming predictive (samples/s)
`--impl reference`: the reference algorithm's CPU restatement on the host cores (rank 0 only) --
a multi-threaded torch-CPU port (oracle/causal_torch.py) of causalbgm/base.py:860-898 with the
reference's two forward passes per iteration, on a bounded sample of the same workload.  TensorFlow
2.10 / TFP 0.18 are not installable here.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

if "--impl" in sys.argv and "reference" in sys.argv:
    # the CPU arm uses every host core; torchrun exports OMP_NUM_THREADS=1, which would silently make it a
    # single-threaded run
    for k in ("OMP_NUM_THREADS", "MKL_NUM_THREADS", "OPENBLAS_NUM_THREADS"):
        os.environ[k] = str(os.cpu_count() or 1)

import numpy as np  # noqa: E402

N_ROWS, V_DIM, Z_DIMS = 100000, 200, [1, 1, 1, 2]
BURN_IN, N_MCMC = 500, 500
X_VALUES = np.linspace(0, 3, 20)
METRIC = "posterior samples/sec (n=100k, z_dim=5) at 1/2/4/8 B200 vs CPU ref"
UNIT = "posterior samples/s"
WORKLOAD = ("CausalBGM continuous-treatment Sim_Hirano_Imbens n=100000 p=200 z_dims=[1,1,1,2] "
            "(z_dim=5), 1000 posterior iters (burn_in 500 + 500 kept), use_bnn=False")


def params(z_dims=None, v_dim=V_DIM, binary=False, use_bnn=False):
    return dict(dataset='Sim_Hirano_Imbens', output_dir='/tmp/bgm_b200_bench', save_res=False,
                save_model=False, binary_treatment=binary, use_bnn=use_bnn, z_dims=list(z_dims or Z_DIMS), v_dim=v_dim,
                lr_theta=1e-4, lr_z=1e-4, g_units=[64] * 5, f_units=[64, 32, 8], h_units=[64, 32, 8],
                kl_weight=1e-4, lr=2e-4, g_d_freq=5, use_z_rec=True, e_units=[64] * 5,
                dz_units=[64, 32, 8])


def make_data(seed, n=N_ROWS, v_dim=V_DIM):
    from bayesgm_b200.datasets import Sim_Hirano_Imbens_sampler
    return Sim_Hirano_Imbens_sampler(N=n, v_dim=v_dim, seed=seed).load_all()


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        return json.load(open(path)), "measured (MEASURED_PEAKS.json)"
    return dict(hbm_gbs=6650.0, bf16_tflops=1590.0), "fallback (B200_PROFILING.md)"


class ClockSampler(object):
    """nvidia-smi clocks / throttle reasons of one GPU during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                 "--format=csv,noheader,nounits", "-lms", "50"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), [c.strip() for c in line.split(",")]))

    def stop(self, t_begin=None, t_end=None):
        if self.proc is None:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["nvidia-smi unavailable"])
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        # an exiting nvidia-smi holds driver locks for a while: launches of this process were seen to stall by
        # 50-90 ms for up to half a second afterwards (per_step_ms of the first end-to-end pass on a fresh box:
        # [52.8, 139.0, 55.2, 105.7, 102.3] against a steady 52.7 two seconds later) -- let it settle before the
        # next timed region
        time.sleep(2.0)
        sm, smax, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ts, r in self.rows:
            if t_begin is not None and not (t_begin <= ts <= t_end + 0.05):
                continue
            try:
                sm.append(float(r[0]))
                smax = float(r[1])
                for nm, val in zip(names, r[3:7]):
                    if val.lower().startswith("active"):
                        reasons.add(nm)
            except (ValueError, IndexError):
                pass
        return dict(sm_mhz=float(np.median(sm)) if sm else None, sm_max_mhz=smax,
                    reasons=sorted(reasons), samples=len(sm))


def host_nets(P, seed=123):
    """The glorot draws of CausalBGM(params, random_seed=123), in its order (g, e, f, h)."""
    from bayesgm_b200.nets import DenseNet
    rs = np.random.RandomState(seed)
    zdims = P['z_dims']
    zd = sum(zdims)
    g = DenseNet(zd, P['v_dim'] + 1, 'g', P['g_units'], rs)
    e = DenseNet(P['v_dim'], zd, 'e', P['e_units'], rs)
    f = DenseNet(zdims[0] + zdims[1] + 1, 2, 'f', P['f_units'], rs)
    h = DenseNet(zdims[0] + zdims[2], 2, 'h', P['h_units'], rs)
    return dict(g=g.as_oracle_layers(), e=e.as_oracle_layers(), f=f.as_oracle_layers(), h=h.as_oracle_layers())


def cpu_reference_rate(iters, data, P=None, warm=1):
    """Reference-faithful RW-MH loop on ALL host cores (multi-threaded torch-CPU port): samples/s."""
    import torch
    from oracle import causal_torch
    torch.set_num_threads(os.cpu_count() or 1)
    P = P or params()
    nets = host_nets(P)
    n = len(data[0])
    if warm:
        causal_torch.mh_sampler(P, nets, data, q_sd=1.0, burn_in=0, n_keep=warm)
    t0 = time.perf_counter()
    causal_torch.mh_sampler(P, nets, data, q_sd=1.0, burn_in=0, n_keep=iters)
    dt = time.perf_counter() - t0
    return n * iters / dt, dt, torch.get_num_threads()


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    iters = 30
    data = make_data(0)
    for _ in range(args.warmup):
        cpu_reference_rate(1, data, warm=0)
    t0 = time.perf_counter()
    threads = 1
    for _ in range(args.steps):
        _, _, threads = cpu_reference_rate(iters, data, warm=0)
    dt = time.perf_counter() - t0
    value = N_ROWS * iters * args.steps / dt
    sample = "%d MH iterations per step over all n=%d rows (of T=%d)" % (iters, N_ROWS, BURN_IN + N_MCMC)
    emit({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "config": {"workload": WORKLOAD, "sample": sample},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample,
                         "host_cpus": os.cpu_count(), "omp_num_threads": os.environ.get("OMP_NUM_THREADS")},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "multi-threaded torch-CPU restatement (oracle/causal_torch.py) of causalbgm/base.py:820-904, two fp32 "
                "forward passes per iteration like the reference; TF 2.10/TFP 0.18 not installable here",
    })


class Dist(object):
    def __init__(self):
        import torch
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        torch.cuda.set_device(self.local)
        self.group = None
        if self.world > 1:
            import torch.distributed as dist
            dist.init_process_group("nccl", device_id=torch.device("cuda", self.local))
            self.group = dist.group.WORLD

    def barrier(self):
        import torch
        if self.world > 1:
            import torch.distributed as dist
            dist.barrier()
        torch.cuda.synchronize()

    def max(self, *vals):
        import torch
        t = torch.tensor(list(vals), dtype=torch.float64, device='cuda')
        if self.world > 1:
            import torch.distributed as dist
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return [float(a) for a in t]

    def close(self):
        if self.world > 1:
            import torch.distributed as dist
            dist.destroy_process_group()


def timed_passes(D, steps, step_fn, ratio_limit=1.3, per_step_ms=None):
    """K steps between barriers; a pass in which one step takes > ratio_limit x the fastest one was disturbed
    from outside (observed right after another CUDA process exits on the box) and is re-measured ONCE.
    per_step_ms(): device-side per-step times (CUDA events) for steps that return before the GPU is done;
    default: host time per step (steps that end with their device-to-host copy)."""
    def one():
        import gc
        gc.collect()
        gc.disable()            # a generation-2 collection inside a 50 ms step shows up as a +10-40 ms outlier
        try:
            D.barrier()
            t0 = time.perf_counter()
            per = []
            for i in range(steps):
                ts = time.perf_counter()
                step_fn(i)
                per.append(time.perf_counter() - ts)
            D.barrier()
            wall = time.perf_counter() - t0
        finally:
            gc.enable()
        if per_step_ms is not None:
            per = per_step_ms()
        return wall, per, t0
    wall, per, t0 = one()
    first = None
    scale = 1.0 if per_step_ms is not None else 1e3
    if D.max(max(per) / max(min(per), 1e-9))[0] > ratio_limit:
        first = {"ms_per_step": 1e3 * wall / steps, "per_step_ms": [round(scale * p, 2) for p in per]}
        time.sleep(2.0)
        wall, per, t0 = one()
    timed_passes.last_per_step_ms = [round(scale * p, 2) for p in per]
    return D.max(wall)[0], first, t0, t0 + wall


def run_sampler(args, cfg):
    """cfg3 (default), cfg2 and cfg3bnn: the MH sampler + predict()."""
    import torch
    from bayesgm_b200 import CausalBGM, _lib
    from bayesgm_b200.shard import shard_rows
    import ctypes as C

    D = Dist()
    rank, world, local = D.rank, D.world, D.local
    n_gpus = world
    binary = cfg == "cfg2"
    bnn = cfg == "cfg3bnn"
    if binary:
        from bayesgm_b200.datasets import acic_shaped_binary
        n_per, v_dim, z_dims = 20000, 100, [3, 6, 3, 6]
        workload = ("CausalBGM binary-treatment ACIC-shaped synthetic n=20000 p=100 z_dims=[3,6,3,6], 1000 posterior "
                    "iters (burn_in 500 + 500 kept), use_bnn=False")
        x, y, v = acic_shaped_binary(n=n_per * world, p=v_dim, seed=0)
        x, y = x.reshape(-1, 1), y.reshape(-1, 1)
    else:
        n_per, v_dim, z_dims = (20000 if bnn else N_ROWS), V_DIM, Z_DIMS
        workload = WORKLOAD if not bnn else WORKLOAD.replace("n=100000", "n=20000").replace("use_bnn=False", "use_bnn=True "
                                                                                          "(DenseFlipout + batch-stat BN)")
        x, y, v = make_data(0, n=n_per * world, v_dim=v_dim)          # ONE data set, sharded by rows
    lo, hi = shard_rows(n_per * world, rank, world)
    x, y, v = x[lo:hi], y[lo:hi], v[lo:hi]
    P = params(z_dims, v_dim, binary, use_bnn=bnn)
    model = CausalBGM(params=P, random_seed=123)
    if not bnn:
        model.set_sampler_engine(args.engine)
    sinfo = model.sampler_info()
    info = model.kernel_info() if not bnn else dict(macs_per_row=sinfo['macs_per_eval'] // 2, issued_macs_per_row=2 * sinfo['macs_per_eval'],
                                                     proj_dim=0, smem_bytes=sinfo['smem_bytes'])
    tensor = sinfo['engine'] == 'tensor'
    kname = sinfo['kernel']
    T = BURN_IN + N_MCMC
    zd = sum(z_dims)
    # ---- device-resident arm ----
    _, xd, yd, vd, ldv, n = model._stage((x, y, v))
    flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')   # > 126 MB L2
    ev0 = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    ev1 = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    last = {}

    def step(i, timed=True):
        flush.zero_()                              # L2 flush between steps
        aux = model._aux(vd, ldv, n)               # causal_project_kernel: part of the step (wall clock) ...
        if timed:
            ev0[i].record()                        # ... the events bracket the sampler launch(es) alone (roofline)
        last['r'] = model._mh_device(xd, yd, vd, ldv, n, BURN_IN, N_MCMC, 1.0, False, 1.0, 0.25, 0.05, 50, 100,
                                     seed=1000 + i, row_offset=lo, aux=aux)
        if timed:
            ev1[i].record()

    # W warm-up steps, then keep warming until ~2.5 s of GPU work has run: a step is ~35 ms, and an idle B200
    # needs far longer than 3 such steps to reach its boost clock.  nvidia-smi is started BEFORE the warm-up:
    # its start-up stalled CUDA calls of the first timed step when it was started at the timed region's edge
    clocks = ClockSampler(local)
    clocks.start()
    t_w = time.perf_counter()
    warm_run = 0
    while warm_run < args.warmup or (time.perf_counter() - t_w < 2.5 and warm_run < 200):
        step(warm_run % max(args.steps, 1), False)
        warm_run += 1
        torch.cuda.synchronize()
    wall, first_attempt, t0, t1 = timed_passes(D, args.steps, step,
                                               per_step_ms=lambda: [a.elapsed_time(b) for a, b in zip(ev0, ev1)])
    kern_ms = [a.elapsed_time(b) for a, b in zip(ev0, ev1)]
    clk = clocks.stop(t0, t1)
    accept = float(last['r']['accept_count'].sum().item()) / (T * n)
    kern_ms_mean = D.max(float(np.mean(kern_ms)))[0]
    value = n * T * n_gpus * args.steps / wall

    # ---- end-to-end arm: predict() from pinned host buffers, ADRF partial sums all-reduced over the ranks ----
    xh, yh, vh = [torch.from_numpy(np.ascontiguousarray(a)).pin_memory() for a in (x, y, v)]

    def e2e_step(i):
        return model.predict((xh, yh, vh), alpha=0.01, n_mcmc=N_MCMC, burn_in=BURN_IN,
                             x_values=None if binary else X_VALUES, q_sd=1.0, sample_y=True, bs=n, seed=2000 + i,
                             row_offset=lo, group=D.group, verbose=0)
    # warm-up: at least W calls, then until the call time has settled (the first CUDA process on a fresh box needs
    # about a second of predict() calls before it does: per-step 110, 111, 125, 118, 103, 58, 53 ms ... were observed)
    best, n_warm = float("inf"), 0
    while n_warm < 20:
        torch.cuda.synchronize()
        ts = time.perf_counter()
        e2e_step(n_warm)
        dt = time.perf_counter() - ts
        n_warm += 1
        settled = dt <= 1.1 * best
        best = min(best, dt)
        if n_warm >= max(args.warmup, 3) and D.max(0.0 if settled else 1.0)[0] == 0.0:
            break
    e2e_wall, e2e_first, _, _ = timed_passes(D, args.steps, e2e_step)
    e2e_per_step = list(timed_passes.last_per_step_ms)
    e2e_value = n * T * n_gpus * args.steps / e2e_wall

    if rank == 0:
        peaks, peak_src = measured_peaks()
        tf = C.c_double()
        _lib.call("bgm_fp32_peak_tflops", C.byref(tf), _lib.stream_ptr())
        fp32_peak = tf.value
        evals = (T + 1) if not bnn else 2 * T
        flop_per_launch = 2.0 * info['macs_per_row'] * n * evals
        achieved_tflops = flop_per_launch / (kern_ms_mean * 1e-3) / 1e12
        issued_macs = sinfo['tensor_issued_macs_per_row'] if tensor else info['issued_macs_per_row']
        issued_tflops = 2.0 * issued_macs * n * (evals if not bnn else T) / (kern_ms_mean * 1e-3) / 1e12
        bytes_per_launch = 4.0 * n * (v_dim + 2) + 4.0 * N_MCMC * n * zd + 8.0 * n * (zd + 1)
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tpath) and cfg == "cfg3":
            traffic = json.load(open(tpath)).get("causal_mh_tc_kernel_dram_bytes_per_launch" if tensor
                                                 else "causal_mh_kernel_dram_bytes_per_launch")
        if world == 1:
            cpu_iters = 100
            if bnn:
                cpu_baseline = {"value": None, "unit": UNIT, "cores": os.cpu_count(), "kind": "port",
                                "sample": "see the cfg3 line (deterministic nets); the Bayesian-net CPU oracle is NumPy, single-threaded"}
            else:
                cpu_rate, cpu_dt, threads = cpu_reference_rate(cpu_iters, (x, y, v), P)
                cpu_baseline = {"value": cpu_rate, "unit": UNIT, "cores": threads, "kind": "port",
                                "sample": "%d MH iterations over all n=%d rows, multi-threaded torch-CPU port, two forward "
                                          "passes per iteration like the reference (%.1f s)" % (cpu_iters, n, cpu_dt)}
        else:   # the CPU baseline is timed at N=1 only
            cpu_baseline = {"value": None, "unit": UNIT, "cores": os.cpu_count(), "kind": "port",
                            "sample": "not timed at N>1 (see the N=1 line)"}
        if tensor:
            # wide layers on the tensor pipe as 3xTF32: three TF32 MMAs per fp32-accurate product.
            # MEASURED_PEAKS.json holds the dense bf16 peak only; kind::tf32 runs at half that rate.
            bf16_peak = peaks["bf16_tflops"]
            roofline = {"bound": "tensor", "achieved": achieved_tflops, "peak": bf16_peak, "unit": "TFLOP/s",
                        "frac": achieved_tflops / bf16_peak, "traffic": traffic,
                        "peak_source": peak_src + " bf16_tflops (burst; the kernel is timed alone)",
                        "issued": issued_tflops, "issued_frac_of_tf32_peak": issued_tflops / (bf16_peak / 2.0),
                        "fp32_pipe_peak": fp32_peak, "achieved_over_fp32_pipe_peak": achieved_tflops / fp32_peak,
                        "note": "achieved = ALGORITHMIC 2*%d FLOP per row-iteration (the reference's log-posterior, "
                                "evaluated once per iteration). The 64-wide layers run on tcgen05 as error-compensated "
                                "3xTF32 (fp32-level error): the kernel ISSUES 2*%d FLOP-equivalents per row-iteration, "
                                "nearly all on the tensor pipe at the TF32 rate (= bf16 peak / 2), so an fp32-accurate "
                                "run of this net cannot exceed ~%.0f algorithmic TFLOP/s; the fp32 FMA pipe peak "
                                "(%.1f TFLOP/s, measured live) is what the SIMT engine is bound by"
                                % (info['macs_per_row'], issued_macs,
                                   bf16_peak / 2.0 * info['macs_per_row'] / max(issued_macs, 1), fp32_peak)}
        else:
            roofline = {"bound": "fp32", "achieved": achieved_tflops, "peak": fp32_peak, "unit": "TFLOP/s",
                        "frac": achieved_tflops / fp32_peak, "traffic": traffic,
                        "peak_source": "bgm_fp32_peak_tflops (dependent-FFMA micro-benchmark, measured live)",
                        "issued": issued_tflops, "issued_frac": issued_tflops / fp32_peak,
                        "note": ("compute-bound SIMT kernel. achieved = ALGORITHMIC 2*%d FLOP per row per network evaluation"
                                 % info['macs_per_row']) +
                                ("; Bayesian nets: 2 evaluations per iteration (proposal and current state, fresh noise each, "
                                 "like causalbgm/base.py:865-866), each issuing 2x the multiply-adds (loc and perturbation "
                                 "products of DenseFlipout)" if bnn else
                                 "; the kernel ISSUES 2*%d: the v_dim-wide last layer of g_net is evaluated in its %d-dim "
                                 "row space (exact QR identity, DESIGN.md 4.1)" % (issued_macs, info['proj_dim']))}
        launches = ["causal_project_kernel", kname] if not bnn else [kname + " x%d (one per iteration) + 1 statistics launch" % T]
        out = {
            "metric": METRIC if cfg == "cfg3" else "posterior samples/sec (%s)" % cfg, "value": value, "unit": UNIT,
            "n_gpus": n_gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * wall / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload, "rows_per_gpu": n, "rows_total": n * n_gpus, "iterations": T, "q_sd": 1.0,
                       "noise": "in-kernel Philox4x32-10", "engine": sinfo['engine'], "l2": "flushed between steps (256 MB memset)",
                       "parallelism": "one data set of %d rows sharded by rows x%d; no collective in the sampler; e2e: one NCCL "
                                      "all-reduce of the ADRF partial sums per predict()" % (n * n_gpus, n_gpus),
                       "acceptance_rate": accept},
            "e2e": {"value": e2e_value, "unit": UNIT,
                    "h2d_bytes_per_step": int(4 * n * (v_dim + 2)),
                    "d2h_bytes_per_step": int(4 * len(X_VALUES) * N_MCMC) if not binary else int(4 * 3 * n),
                    "api": "CausalBGM.predict(%s, sample_y=True, bs=n%s)" % ("x_values=linspace(0,3,20)" if not binary else "binary",
                                                                            ", group=WORLD" if world > 1 else ""),
                    "collectives_per_step": 0 if world == 1 else (2 if not binary else 0),
                    "ms_per_step": 1e3 * e2e_wall / args.steps, "per_step_ms": e2e_per_step, "warmup_calls": n_warm,
                    "remeasured_after_disturbed_pass": e2e_first},
            "gpu_launches": (2 if not bnn else T + 1) * args.steps,
            "launches_per_step": launches,
            "kernel": {"name": kname, "engine": sinfo['engine'], "ms_per_launch": kern_ms_mean,
                       "ms_per_launch_min_max": [float(min(kern_ms)), float(max(kern_ms))], "warmup_steps_run": warm_run,
                       "remeasured_after_disturbed_pass": first_attempt,
                       "warps_per_cta": 16 if 'tc16' in kname else 8,
                       "smem_bytes": model.launch_smem_bytes() if not bnn else sinfo['smem_bytes']},
            "roofline": roofline,
            "roofline_hbm": {"bound": "hbm", "achieved": bytes_per_launch / (kern_ms_mean * 1e-3) / 1e9,
                             "peak": peaks["hbm_gbs"], "unit": "GB/s",
                             "frac": bytes_per_launch / (kern_ms_mean * 1e-3) / 1e9 / peaks["hbm_gbs"],
                             "peak_source": peak_src,
                             "note": "north_star asks for the HBM fraction; the kernel is not HBM-bound (SURVEY F7)"},
            "cpu_baseline": cpu_baseline,
            "clocks": clk,
        }
        emit(out)
    D.close()


def run_cfg4(args):
    """EGM training steps (causalbgm/base.py:380-431) on n = 1e6 rows sharded over the GPUs, batch 32 per GPU,
    gradients all-reduced (NCCL) in every discriminator / generator step."""
    import torch
    from bayesgm_b200 import CausalBGM
    from bayesgm_b200.shard import shard_rows
    D = Dist()
    n_total = args.rows or 1000000
    iters = args.iters or 200
    x, y, v = make_data(0, n=n_total)
    lo, hi = shard_rows(n_total, D.rank, D.world)
    x, y, v = x[lo:hi], y[lo:hi], v[lo:hi]
    P = params([1, 1, 1, 7])
    model = CausalBGM(params=P, random_seed=123)
    xd, yd, vd = [torch.from_numpy(np.ascontiguousarray(a)).cuda() for a in (x, y, v)]
    clocks = ClockSampler(D.local)
    clocks.start()

    def step(i, group=D.group, index_stream='numpy'):
        model.egm_init((xd, yd, vd), egm_n_iter=iters - 1, batch_size=32, egm_batches_per_eval=10 ** 9, verbose=0,
                       group=group, eval_during=False, index_stream=index_stream)
        torch.cuda.synchronize()
    for i in range(max(1, min(args.warmup, 2))):
        step(i)
    wall, first, t0, t1 = timed_passes(D, args.steps, step, ratio_limit=1.5)
    clk = clocks.stop(t0, t1)
    wall_nc = None
    if D.world > 1:        # the same without the collective: its share of the step
        wall_nc, _, _, _ = timed_passes(D, args.steps, lambda i: step(i, None), ratio_limit=1.5)
    # the same steps fed by subset-sampled mini-batches (O(batch) host work per draw instead of the O(n)
    # permutation behind np.random.choice(n, 32, replace=False)): what the device path does when the host stream is not the limit
    step(0, D.group, 'floyd')
    wall_fl, _, _, _ = timed_passes(D, args.steps, lambda i: step(i, D.group, 'floyd'), ratio_limit=1.5)
    if D.rank == 0:
        value = iters * args.steps / wall
        emit({"metric": "EGM training iterations/sec (5 discriminator + 1 generator steps, batch 32 per GPU)", "value": value,
              "unit": "EGM iterations/s", "n_gpus": D.world, "steps": args.steps, "warmup": args.warmup,
              "ms_per_step": 1e3 * wall / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
              "dtype": "f32", "data": "synthetic",
              "config": {"workload": "cfg4: CausalBGM EGM training, Sim_Hirano_Imbens n=%d p=200 z_dims=[1,1,1,7] sharded by rows over "
                                     "%d GPU(s), %d EGM iterations per step, batch 32 per GPU (global batch %d), use_bnn=False"
                                     % (n_total, D.world, iters, 32 * D.world),
                         "rows_per_gpu": hi - lo,
                         "collectives_per_iteration": 0 if D.world == 1 else 6,
                         "collective": "NCCL all-reduce of the flat gradient buffer (dz: ~3k floats x5, g|e|f|h: ~66k floats x1)",
                         "index_stream": "NumPy legacy generator continued natively on a background thread (bit-exact)"},
              "mini_batches_per_s_per_gpu": 6 * value,
              "index_stream_floyd": {"value": iters * args.steps / wall_fl, "unit": "EGM iterations/s",
                                     "ms_per_step": 1e3 * wall_fl / args.steps,
                                     "note": "egm_init(index_stream='floyd'): same distribution of mini-batches, not NumPy's stream"},
              "iterations_per_s_without_collective": None if wall_nc is None else iters * args.steps / wall_nc,
              "allreduce_share_of_step": None if wall_nc is None else max(0.0, 1.0 - wall_nc / wall),
              "reference_level": "tutorial tqdm: ~55 mini-batches/s in the iterative phase (docs/source/causalbgm/tutorial_py.ipynb:372)",
              "e2e": {"value": value, "unit": "EGM iterations/s", "h2d_bytes_per_step": int(iters * 6 * 32 * (4 + 4 * 10)),
                      "d2h_bytes_per_step": 32, "api": "CausalBGM.egm_init(data, group=WORLD)"},
              "gpu_launches": int(iters * args.steps * (5 * 3 + 5)), "clocks": clk})
    D.close()


def run_cfg5(args):
    """BGM missing-data imputation (bgm/base.py:527-663): HMC over the latent of every row with the shared step size
    all-reduced per adaptation step, then streaming posterior-predictive mean / intervals."""
    import torch
    from bayesgm_b200 import BGM
    from bayesgm_b200.datasets import simulate_z_hetero
    from bayesgm_b200.shard import shard_rows
    D = Dist()
    n_total = args.rows or (500000 // 8) * D.world         # weak scaling: 62500 rows per GPU
    burn_in, n_mcmc, L = 100, 100, 10
    X, Y = simulate_z_hetero(n=n_total, k=10, d=499, seed=42)
    data = np.c_[X, Y].astype(np.float32)
    data[np.random.RandomState(1).rand(*data.shape) < 0.3] = np.nan
    lo, hi = shard_rows(n_total, D.rank, D.world)
    data = data[lo:hi]
    n = hi - lo
    P = dict(dataset='cfg5', output_dir='/tmp/bgm_b200_bench', save_res=False, save_model=False, use_bnn=False, x_dim=500,
             z_dim=10, g_units=[64] * 5, e_units=[64] * 5, dz_units=[64, 32, 8], dx_units=[64, 32, 8], lr=1e-3, lr_theta=5e-3,
             lr_z=5e-3, gamma=0.0, alpha=0.0, g_d_freq=1, kl_weight=5e-5)
    model = BGM(params=P, random_seed=123)
    if args.engine != "auto":
        model.set_hmc_engine(args.engine)
    info = model.kernel_info()
    einfo = model.hmc_engine_info()
    xdev, ldx, _ = model._stage_x(data, torch)
    ev0 = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    ev1 = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    clocks = ClockSampler(D.local)
    clocks.start()

    def step(i, timed=True):
        if timed:
            ev0[i].record()
        r = model._hmc_device(xdev, ldx, n, n_mcmc, burn_in, 0.01, L, seed=42 + i, row_offset=lo, group=D.group, n_total=n_total)
        if timed:
            ev1[i].record()
        return r
    for i in range(max(1, min(args.warmup, 2))):
        step(0, False)
        torch.cuda.synchronize()
    wall, first, t0, t1 = timed_passes(D, args.steps, step,
                                       per_step_ms=lambda: [a.elapsed_time(b) for a, b in zip(ev0, ev1)])
    kern_ms = D.max(float(np.mean([a.elapsed_time(b) for a, b in zip(ev0, ev1)])))[0]
    clk = clocks.stop(t0, t1)
    T = burn_in + n_mcmc
    value = n * T * D.world * args.steps / wall

    def e2e_step(i):
        return model.predict(data, alpha=0.05, bs=1000, n_mcmc=n_mcmc, burn_in=burn_in, step_size=0.01, num_leapfrog_steps=L,
                             seed=42 + i, group=D.group, row_offset=lo, n_total=n_total, verbose=0)
    e2e_step(0)
    e2e_wall, e2e_first, _, _ = timed_passes(D, args.steps, e2e_step, ratio_limit=1.5)
    if D.rank == 0:
        import ctypes as C
        from bayesgm_b200 import _lib
        tf = C.c_double()
        _lib.call("bgm_fp32_peak_tflops", C.byref(tf), _lib.stream_ptr())
        grads = T * L + 1
        achieved = 2.0 * info['macs_per_grad'] * n * grads / (kern_ms * 1e-3) / 1e12
        note = "algorithmic 2*%d FLOP per gradient evaluation (forward + d/dz), %d evaluations per row" % (info['macs_per_grad'], grads)
        if einfo['engine'] == 'tensor':
            peaks, peak_src = measured_peaks()
            peak = peaks.get("bf16_tflops_sustained") or peaks.get("bf16_tflops") or 1648.1
            issued = 2.0 * einfo['tensor_issued_macs_per_grad'] * n * grads / (kern_ms * 1e-3) / 1e12
            roof = {"bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak, "traffic": None,
                    "peak_source": peak_src + " bf16 TFLOP/s, the sustained figure where the file has one (the kernel runs for about half a second)",
                    "issued": issued, "issued_frac_of_tf32_peak": issued / (peak / 2.0),
                    "fp32_pipe_peak": tf.value, "achieved_over_fp32_pipe_peak": achieved / tf.value,
                    "note": note + "; every 64-wide product runs on tcgen05 as error-compensated 3xTF32 (3 passes at the TF32 rate = "
                                   "bf16 peak / 2), the likelihood terms (exp, log1p, two reciprocals per feature) on the FMA / SFU pipes"}
        else:
            roof = {"bound": "fp32", "achieved": achieved, "peak": tf.value, "unit": "TFLOP/s", "frac": achieved / tf.value,
                    "traffic": None, "peak_source": "bgm_fp32_peak_tflops (measured live)", "note": note}
        emit({"metric": "posterior samples/sec (cfg5: BGM HMC, x_dim=500, z_dim=10)", "value": value, "unit": UNIT, "n_gpus": D.world,
              "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * wall / args.steps, "higher_is_better": True,
              "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
              "config": {"workload": "cfg5: BGM missing-data imputation, simulate_z_hetero n=%d (x_dim=500, z_dim=10), 30%% MCAR, HMC "
                                     "burn_in 100 + 100 kept, L=10, step 0.01, shared step size adapted for 80 steps" % n_total,
                         "rows_per_gpu": n, "collectives_per_step": 0 if D.world == 1 else int(0.8 * burn_in),
                         "collective": "NCCL all-reduce of ONE float64 (sum of accept probabilities) per adaptation step"},
              "e2e": {"value": n * T * D.world * args.steps / e2e_wall, "unit": UNIT, "h2d_bytes_per_step": int(4 * n * 500),
                      "d2h_bytes_per_step": int(4 * n * 500 + 8 * int(np.isnan(data).sum())),     # imputed matrix + (lower, upper) of every missing entry
                      "api": "BGM.predict(data_with_NaN, bs=1000, group=WORLD)",
                      "ms_per_step": 1e3 * e2e_wall / args.steps},
              "gpu_launches": int((0.8 * burn_in * 2 + 1) * args.steps),
              "kernel": {"name": "hmc_tc_kernel<%d>" % (4 * ((10 + 3) // 4)) if einfo['engine'] == 'tensor' else "hmc_kernel", "engine": einfo['engine'], "ms_per_launch_set": kern_ms,
                         "smem_bytes": einfo['tensor_smem_bytes'] if einfo['engine'] == 'tensor' else info['smem_bytes']},
              "roofline": roof,
              "chain_steps_per_s": n * T * D.world * args.steps / wall, "acceptance_rate": model.last_acceptance_rate,
              "step_size_final": model.last_step_size, "clocks": clk})
    D.close()


def emit(obj):
    """The ONE JSON line on the real stdout (see main(): fd 1 is pointed at stderr while the
    benchmark runs so that library chatter -- e.g. NCCL's version banner -- cannot precede it)."""
    os.write(_REAL_STDOUT, (json.dumps(obj) + "\n").encode())


_REAL_STDOUT = 1


def main():
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="cfg3", choices=["cfg3", "cfg2", "cfg3bnn", "cfg4", "cfg5"])
    ap.add_argument("--rows", type=int, default=0, help="cfg4 / cfg5: total rows (default: the BASELINE size)")
    ap.add_argument("--iters", type=int, default=0, help="cfg4: EGM iterations per step")
    ap.add_argument("--engine", default="auto", choices=["auto", "simt", "tensor"],
                    help="sampler engine (bgm_causal_set_sampler); auto = tensor cores when the net shape allows")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    elif args.config == "cfg4":
        run_cfg4(args)
    elif args.config == "cfg5":
        run_cfg5(args)
    else:
        run_sampler(args, args.config)


if __name__ == "__main__":
    main()
