#!/usr/bin/env python
"""Benchmark of the posterior-sampling hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Workload (BASELINE.json configs[2], the one the metric is quoted on): CausalBGM,
continuous treatment, Sim_Hirano_Imbens-shaped synthetic data n=100000 rows per GPU,
p=200 covariates, z_dims=[1,1,1,2] (z_dim=5), deterministic nets (use_bnn=False),
glorot-uniform weights from RandomState(123); one STEP = one random-walk MH run of
T = 1000 iterations (burn_in 500 + 500 kept) over all rows = n*T posterior samples.

  value : n*T*gpus / step time, inputs resident in HBM, kept states written to HBM
          (one launch of the persistent sampler kernel per step per GPU).
  e2e   : the same through `CausalBGM.predict(data, x_values=linspace(0,3,20))` with
          HOST (pinned) x, y, v: H2D copies, sampler, effect kernel, D2H of the ADRF
          draws all inside the timed region.
Multi-GPU: one process per GPU (torchrun), rows sharded, no collective in the sampler
(weak scaling); timing = max over ranks of CUDA-event time.
`--impl reference`: the reference algorithm's CPU restatement (oracle/, NumPy BLAS on
all host cores) on a bounded sample of the same workload -- TensorFlow 2.10 / TFP 0.18
are not installable here, see DESIGN.md.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_ROWS, V_DIM, Z_DIMS = 100000, 200, [1, 1, 1, 2]
BURN_IN, N_MCMC = 500, 500
X_VALUES = np.linspace(0, 3, 20)
METRIC = "posterior samples/sec (n=100k, z_dim=5) at 1/2/4/8 B200 vs CPU ref"
UNIT = "posterior samples/s"
WORKLOAD = ("CausalBGM continuous-treatment Sim_Hirano_Imbens n=100000 p=200 z_dims=[1,1,1,2] "
            "(z_dim=5), 1000 posterior iters (burn_in 500 + 500 kept), use_bnn=False")


def params():
    return dict(dataset='Sim_Hirano_Imbens', output_dir='/tmp/bgm_b200_bench', save_res=False,
                save_model=False, binary_treatment=False, use_bnn=False, z_dims=Z_DIMS, v_dim=V_DIM,
                lr_theta=1e-4, lr_z=1e-4, g_units=[64] * 5, f_units=[64, 32, 8], h_units=[64, 32, 8],
                kl_weight=1e-4, lr=2e-4, g_d_freq=5, use_z_rec=True, e_units=[64] * 5,
                dz_units=[64, 32, 8])


def make_data(seed):
    from bayesgm_b200.datasets import Sim_Hirano_Imbens_sampler
    return Sim_Hirano_Imbens_sampler(N=N_ROWS, v_dim=V_DIM, seed=seed).load_all()


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


def cpu_reference_rate(iters, n_rows=N_ROWS, data=None, warm=1):
    """Reference-faithful RW-MH loop on the host cores (oracle port): samples/s."""
    from oracle import causal
    from bayesgm_b200.nets import DenseNet
    P = params()
    # same glorot draws, same order (g, e, f, h) as CausalBGM(params, random_seed=123)
    rs = np.random.RandomState(123)
    zd = sum(Z_DIMS)
    g = DenseNet(zd, V_DIM + 1, 'g', P['g_units'], rs)
    e = DenseNet(V_DIM, zd, 'e', P['e_units'], rs)
    f = DenseNet(Z_DIMS[0] + Z_DIMS[1] + 1, 2, 'f', P['f_units'], rs)
    h = DenseNet(Z_DIMS[0] + Z_DIMS[2], 2, 'h', P['h_units'], rs)
    nets = dict(g=g.as_oracle_layers(), e=e.as_oracle_layers(), f=f.as_oracle_layers(), h=h.as_oracle_layers())
    if data is None:
        data = make_data(0)
    data = tuple(a[:n_rows] for a in data)
    if warm:
        causal.mh_sampler(P, nets, data, q_sd=1.0, burn_in=0, n_keep=warm)
    t0 = time.perf_counter()
    causal.mh_sampler(P, nets, data, q_sd=1.0, burn_in=0, n_keep=iters)
    dt = time.perf_counter() - t0
    return n_rows * iters / dt, dt


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count()
    iters = 5
    data = make_data(0)
    for _ in range(args.warmup):
        cpu_reference_rate(1, data=data, warm=0)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        cpu_reference_rate(iters, data=data, warm=0)
    dt = time.perf_counter() - t0
    value = N_ROWS * iters * args.steps / dt
    sample = "%d MH iterations per step over all n=%d rows (of T=%d)" % (iters, N_ROWS, BURN_IN + N_MCMC)
    emit({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "config": {"workload": WORKLOAD, "sample": sample},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "CPU restatement (oracle/) of causalbgm/base.py:820-904; TF 2.10/TFP 0.18 not installable here",
    })


def run_ours(args):
    import torch
    import torch.distributed as dist
    from bayesgm_b200 import CausalBGM, _lib
    import ctypes as C

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    n_gpus = world

    P = params()
    model = CausalBGM(params=P, random_seed=123)
    model.set_sampler_engine(args.engine)
    info = model.kernel_info()
    sinfo = model.sampler_info()
    tensor = sinfo['engine'] == 'tensor'
    kname = sinfo['kernel']
    x, y, v = make_data(rank)                      # weak scaling: every rank its own n rows
    T = BURN_IN + N_MCMC
    # ---- device-resident arm ----
    _, xd, yd, vd, ldv, n = model._stage((x, y, v))
    flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')   # > 126 MB L2
    ev0 = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    ev1 = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]

    def step(i, timed):
        flush.zero_()                              # L2 flush between steps
        aux = model._aux(vd, ldv, n)               # causal_project_kernel: part of the step (wall clock) ...
        if timed:
            ev0[i].record()                        # ... the events bracket the sampler launch alone (roofline)
        r = model._mh_device(xd, yd, vd, ldv, n, BURN_IN, N_MCMC, 1.0, False, 1.0, 0.25, 0.05, 50, 100,
                             seed=1000 + i, row_offset=rank * N_ROWS, aux=aux)
        if timed:
            ev1[i].record()
        return r

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # W warm-up steps, then keep warming until ~1.5 s of GPU work has run: a step is ~35 ms, and an
    # idle B200 (120 MHz) needs far longer than 3 such steps to reach its boost clock
    # nvidia-smi is started BEFORE the warm-up: its start-up (NVML initialisation, driver locks) stalled
    # CUDA calls of the first timed step for 100+ ms when it was started at the timed region's edge
    clocks = ClockSampler(local)
    clocks.start()
    t_w = time.perf_counter()
    warm_run = 0
    while warm_run < args.warmup or (time.perf_counter() - t_w < 2.5 and warm_run < 200):
        step(warm_run % max(args.steps, 1), False)
        warm_run += 1
        torch.cuda.synchronize()
    def timed_pass():
        barrier()
        t0 = time.perf_counter()
        for i in range(args.steps):
            rr = step(i, True)
        barrier()
        t1 = time.perf_counter()
        return rr, t0, t1, [a.elapsed_time(b) for a, b in zip(ev0, ev1)]

    r, t0, t1, kern_ms = timed_pass()
    first_attempt = None
    # A pass in which one step takes >1.3x the fastest one was disturbed from outside (observed: single
    # steps of 70-180 ms next to 34 ms ones right after another CUDA process exits on the box, clocks at
    # max, no throttle reason).  Like a throttled run it is re-measured ONCE; both are reported.
    med_all = torch.tensor([max(kern_ms) / max(min(kern_ms), 1e-9)], dtype=torch.float64, device='cuda')
    if world > 1:
        dist.all_reduce(med_all, op=dist.ReduceOp.MAX)
    if float(med_all[0]) > 1.3:
        first_attempt = {"ms_per_step": 1e3 * (t1 - t0) / args.steps, "ms_per_launch_min_max": [float(min(kern_ms)), float(max(kern_ms))]}
        time.sleep(2.0)
        r, t0, t1, kern_ms = timed_pass()
    wall = t1 - t0
    clk = clocks.stop(t0, t1)
    accept = float(r['accept_count'].sum().item()) / (T * n)
    tm = torch.tensor([wall, float(np.mean(kern_ms))], dtype=torch.float64, device='cuda')
    if world > 1:
        dist.all_reduce(tm, op=dist.ReduceOp.MAX)
    wall, kern_ms_mean = float(tm[0]), float(tm[1])
    value = n * T * n_gpus * args.steps / wall

    # ---- end-to-end arm: predict() from pinned host buffers ----
    xh, yh, vh = [torch.from_numpy(a).pin_memory() for a in (x, y, v)]
    def e2e_step(i):
        return model.predict((xh, yh, vh), alpha=0.01, n_mcmc=N_MCMC, burn_in=BURN_IN, x_values=X_VALUES,
                             q_sd=1.0, sample_y=True, bs=N_ROWS, seed=2000 + i, row_offset=rank * N_ROWS,
                             verbose=0)
    for i in range(min(args.warmup, 2)):
        e2e_step(i)

    def e2e_pass():
        barrier()
        t0 = time.perf_counter()
        per = []
        for i in range(args.steps):
            ts = time.perf_counter()
            e2e_step(i)                 # returns host arrays: every step ends with its D2H
            per.append(time.perf_counter() - ts)
        barrier()
        return time.perf_counter() - t0, per

    e2e_wall, per = e2e_pass()
    e2e_first = None
    ratio = torch.tensor([max(per) / max(min(per), 1e-9)], dtype=torch.float64, device='cuda')
    if world > 1:
        dist.all_reduce(ratio, op=dist.ReduceOp.MAX)
    if float(ratio[0]) > 1.3:           # same rule as the device-resident arm
        e2e_first = {"ms_per_step": 1e3 * e2e_wall / args.steps}
        time.sleep(2.0)
        e2e_wall, per = e2e_pass()
    te = torch.tensor([e2e_wall], dtype=torch.float64, device='cuda')
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = n * T * n_gpus * args.steps / float(te[0])

    if rank == 0:
        peaks, peak_src = measured_peaks()
        tf = C.c_double()
        _lib.call("bgm_fp32_peak_tflops", C.byref(tf), _lib.stream_ptr())
        fp32_peak = tf.value
        flop_per_launch = 2.0 * info['macs_per_row'] * n * (T + 1)
        achieved_tflops = flop_per_launch / (kern_ms_mean * 1e-3) / 1e12
        issued_macs = sinfo['tensor_issued_macs_per_row'] if tensor else info['issued_macs_per_row']
        issued_tflops = 2.0 * issued_macs * n * (T + 1) / (kern_ms_mean * 1e-3) / 1e12
        bytes_per_launch = 4.0 * n * (V_DIM + 2) + 4.0 * N_MCMC * n * sum(Z_DIMS) + 8.0 * n * (sum(Z_DIMS) + 1)
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tpath):
            traffic = json.load(open(tpath)).get("causal_mh_tc_kernel_dram_bytes_per_launch" if tensor
                                                 else "causal_mh_kernel_dram_bytes_per_launch")
        cpu_iters = 10
        if world == 1:
            cpu_rate, cpu_dt = cpu_reference_rate(cpu_iters, data=(x, y, v))
            cpu_baseline = {"value": cpu_rate, "unit": UNIT, "cores": os.cpu_count(), "kind": "port",
                            "sample": "%d MH iterations over all n=%d rows, NumPy BLAS (%.1f s)" % (cpu_iters, n, cpu_dt)}
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
                        "note": "compute-bound SIMT kernel. achieved = ALGORITHMIC 2*%d FLOP per row-iteration "
                                "(the reference's log-posterior, evaluated once per iteration); the kernel ISSUES "
                                "2*%d: the v_dim-wide last layer of g_net is evaluated in its %d-dim row space "
                                "(exact QR identity, DESIGN.md 4.1)"
                                % (info['macs_per_row'], issued_macs, info['proj_dim'])}
        out = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": n_gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * wall / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "rows_per_gpu": n, "iterations": T, "q_sd": 1.0,
                       "noise": "in-kernel Philox4x32-10", "engine": sinfo['engine'], "l2": "flushed between steps (256 MB memset)",
                       "parallelism": "rows sharded x%d, no collective in the sampler" % n_gpus,
                       "acceptance_rate": accept},
            "e2e": {"value": e2e_value, "unit": UNIT,
                    "h2d_bytes_per_step": int(4 * n * (V_DIM + 2)),
                    "d2h_bytes_per_step": int(4 * len(X_VALUES) * N_MCMC),
                    "api": "CausalBGM.predict(x_values=linspace(0,3,20), sample_y=True, bs=n)",
                    "ms_per_step": 1e3 * e2e_wall / args.steps, "remeasured_after_disturbed_pass": e2e_first},
            "gpu_launches": 2 * args.steps,
            "launches_per_step": ["causal_project_kernel", kname],
            "kernel": {"name": kname, "engine": sinfo['engine'], "ms_per_launch": kern_ms_mean,
                       "ms_per_launch_min_max": [float(min(kern_ms)), float(max(kern_ms))], "warmup_steps_run": warm_run,
                       "remeasured_after_disturbed_pass": first_attempt,
                       "warps_per_cta": 16 if 'tc16' in kname else 8, "smem_bytes": sinfo['tensor_smem_bytes'] if tensor else info['smem_bytes']},
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
    if world > 1:
        dist.destroy_process_group()


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
    ap.add_argument("--engine", default="auto", choices=["auto", "simt", "tensor"],
                    help="sampler engine (bgm_causal_set_sampler); auto = tensor cores when the net shape allows")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
