#!/usr/bin/env python
"""bench.py -- weighted phase-space samples/s of the fused hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--samples S] [--jobs J] [--precision f32|f64]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...
    python bench.py --impl reference ...      # the reference's CPU algorithm on the host cores

Workload (BASELINE.json configs[2], the one the metric is quoted on; it fits one
GPU because samples are never stored): Soft Drop (z_cut=0.1, beta_SD=0 i.e. the
mMDT form of C_groomed) critical x subsequent emission MLL ECF e_2^(2)
distribution of quark jets, log sampling with epsilon=1e-10, the reference's
verbatim running-coupling twoEmissionWeight with np.nan_to_num, 100 fixed
logspace edges (sudakov_utils.py:533), boundary condition [1,'minus'].
One JOB = S samples per GPU -> (sum w, sum w^2, count) -> [all-reduce] ->
density +- error -> cdf +- error.  One STEP = J such jobs back to back on fresh
sample indices (J = 50 by default, so that the timed region of a 20-step run
lasts seconds and the GPU reaches its steady clocks and power).  Weak scaling:
S per GPU is fixed.

Besides the headline, rank 0 of a single-GPU run times the other BASELINE
configurations and the SURVEY 8(f) rows (per-event Sudakov sampling, RSS shower
and record replay, fused EWOC) briefly and reports them under "sub_results"
(each with its kernel time and, where one applies, roofline fractions).
Prints ONE JSON line (rank 0).
"""
import argparse
import ctypes as C
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "weighted phase-space samples/sec"
UNIT = "samples/s"
EPS, Z_CUT, BETA = 1e-10, .1, 2.
N_EDGES = 100
SEED = 20260101
# algorithmic work per weighted sample, flop-equivalents / SFU operations (the contract figures)
CONTRACT = {"c1": (80., 4.), "c2i": (110., 7.), "c3": (390., 28.), "c4": (460., 32.)}
# Work of a C3 sample whose reference weight is NaN -> 0 (99.7 % of the running-coupling samples), same counting rule
# (DESIGN.md section 4.1): Philox 100; counted mode: 4 convert+FMA 12, ex2+lg2 10, observable 9, Landau test 13, log-bin
# 12, count 1 = 157; counts on request: 3 conversions 3, 2 FMA 4, 1 add, 4 compares, 3 predicate ops = 115.
ZERO_WEIGHT = {"all": (157., 2.), "weighted": (115., 0.)}
ACTIVE_SHARE_C3_RC = 0.003


def edges_c3():
    return np.logspace(np.log10(EPS) - 1, np.log10(.5), N_EDGES)


def workload_config(args, world):
    """identical in both arms (the reference arm times a bounded sample of this workload, see its cpu_baseline)"""
    return {
        "workload": "C3: Soft Drop (z_cut=0.1, beta_sd=0, mMDT C_groomed) critical x subsequent MLL ECF e2^(2), "
                    "quark, running coupling verbatim twoEmissionWeight + nan_to_num, log sampling eps=1e-10",
        "samples_per_job_per_gpu": int(args.samples), "jobs_per_step": int(args.jobs),
        "samples_per_gpu_per_step": int(args.samples) * int(args.jobs),
        "global_samples_per_step": int(args.samples) * int(args.jobs) * world,
        "bins": N_EDGES - 1, "edges": "np.logspace(-11, log10(.5), 100)", "boundary_condition": "[1,'minus']",
        "rng": "philox4x32-10, counter = global sample index",
        "sharding": "sample index ranges per rank, one all-reduce of [sum w, sum w^2, count] per job" if world > 1
                    else "single GPU",
        "l2": "256 MiB memset between timed steps; the path reads no per-sample HBM data "
              "(inputs: 800 B of bin edges per job)",
    }


# ---------------------------------------------------------------------------
# clocks
# ---------------------------------------------------------------------------
class ClockSampler:
    """SM clock and throttle reasons during the timed region: NVML polled every few ms from a thread; nvidia-smi
    as the fallback."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}

    def __init__(self, gpu_index):
        self.rows, self.proc, self.nvml, self.stop_flag = [], None, None, False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nvml = pynvml
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(gpu_index)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM))
            self.thread = threading.Thread(target=self._poll, daemon=True)
            self.thread.start()
            return
        except Exception:
            self.nvml = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                          "-lms", "20", "-i", str(gpu_index)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _poll(self):
        n = self.nvml
        while not self.stop_flag:
            try:
                mhz = n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM)
                try:
                    mask = n.nvmlDeviceGetCurrentClocksEventReasons(self.handle)
                except Exception:
                    mask = n.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle)
                watts = n.nvmlDeviceGetPowerUsage(self.handle) / 1000.
                self.rows.append((time.perf_counter(), float(mhz), int(mask), watts))
            except Exception:
                pass
            time.sleep(0.005)

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), [x.strip() for x in line.split(",")]))

    def stop(self, t0, t1):
        if self.nvml is not None:
            self.stop_flag = True
            self.thread.join(timeout=1.)
            rows = [r for r in self.rows if t0 <= r[0] <= t1] or self.rows
            if not rows:
                return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["no samples"]}
            reasons = sorted({name for r in rows for bit, name in self.REASONS.items() if r[2] & bit})
            return {"sm_mhz": statistics.median(r[1] for r in rows), "sm_max_mhz": self.max_mhz, "reasons": reasons,
                    "samples": len(rows), "power_w_max": max(r[3] for r in rows),
                    "power_w_median": statistics.median(r[3] for r in rows), "source": "nvml"}
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        rows = [r for t, r in self.rows if t0 <= t <= t1 and len(r) >= 9] or [r for _, r in self.rows if len(r) >= 9]
        if not rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        sm = [float(r[1]) for r in rows]
        reasons = set()
        for r in rows:
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": float(rows[0][2]), "reasons": sorted(reasons),
                "samples": len(rows), "power_w_max": max(float(r[3]) for r in rows), "source": "nvidia-smi"}


# ---------------------------------------------------------------------------
# CPU baseline: the reference's algorithm restated by the oracle
# ---------------------------------------------------------------------------
def _cpu_chunk(args):
    """One process: the reference's path on `n` samples -- Python-loop sampling
    (sampler.py:30-34), NumPy weights and np.histogram (oracle restatement)."""
    n, seed, loop = args
    from oracle import jmc_oracle as O
    np.seterr(all='ignore')
    t0 = time.perf_counter()
    if loop:
        crit, jc = O.sample_loop_critical('log', n, Z_CUT, EPS, seed)
        sub, js = O.sample_loop_ungroomed('log', n, EPS, seed + 1)
    else:
        u = O.mt_uniforms(seed, 4 * n)
        crit, jc, _ = O.sample_critical('log', u[:2 * n].reshape(n, 2), Z_CUT, EPS)
        sub, js, _ = O.sample_ungroomed('log', u[2 * n:].reshape(n, 2), EPS)
    csub = sub[:, 0] * sub[:, 1] ** BETA * sub[:, 1] ** BETA
    obs = np.maximum(O.ecf_groomed(crit[:, 0], crit[:, 1], Z_CUT, BETA, z_pre=Z_CUT, f=0., acc='MLL'), csub)
    w = np.nan_to_num(O.weight_two_emission(crit, sub, Z_CUT, BETA, 'quark', False)) * jc * js
    h, h2, c = O.histograms(obs, w, edges_c3())
    return h, h2, c, time.perf_counter() - t0


def cpu_reference_step(n_per_proc, procs, seed, loop, pool):
    t0 = time.perf_counter()
    parts = pool.map(_cpu_chunk, [(n_per_proc, seed + 7919 * i, loop) for i in range(procs)])
    from oracle import jmc_oracle as O
    h = sum(p[0] for p in parts)
    h2 = sum(p[1] for p in parts)
    c = sum(p[2] for p in parts)
    area = np.log(1. / EPS) ** 4.
    dens, err = O.density_from_hist(h, h2, c, edges_c3(), area, n_per_proc * procs)
    O.cumulative(dens, err, edges_c3(), last_bc=[1., 'minus'])
    return time.perf_counter() - t0


def run_reference_arm(args):
    """bench.py --impl reference: the reference's own CPU algorithm (oracle port;
    the Python reference itself cannot travel to the GPU box) on all host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import multiprocessing as mp
    cores = os.cpu_count() or 1
    n_per_proc = int(args.cpu_samples) if args.cpu_samples else 60000
    ctx = mp.get_context("fork")
    with ctx.Pool(cores) as pool:
        for w in range(args.warmup):
            cpu_reference_step(n_per_proc, cores, 100 + w, True, pool)
        times = [cpu_reference_step(n_per_proc, cores, 1000 + k, True, pool) for k in range(args.steps)]
        vec_n = n_per_proc * 20
        vec_t = cpu_reference_step(vec_n, cores, 5000, False, pool)
    total = n_per_proc * cores
    value = total * args.steps / sum(times)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * sum(times) / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args, max(1, args.gpus)),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": "each step is a bounded sample of the workload: %d samples (%d per process x %d "
                                   "processes); Python-loop sampling as sampler.py:30-34 + NumPy weights + "
                                   "np.histogram, fanned over all cores (the reference itself is single-threaded)"
                                   % (total, n_per_proc, cores),
                         "vectorised_port_value": vec_n * cores / vec_t,
                         "vectorised_port_note": "same path with NumPy-vectorised sampling (not what the reference "
                                                 "does; an upper bound for a NumPy implementation), all cores"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def cpu_baseline_quick():
    """single-thread port of the reference path on a bounded sample (~10-20 s)"""
    n = 1000000          # the full reference path at N = 1e6 (about 8 s) ...
    _, _, _, t_loop = _cpu_chunk((n, 42, True))
    nv = 10000000        # ... and its vectorised stage alone at N = 1e7 (about 6 s)
    _, _, _, t_vec = _cpu_chunk((nv, 43, False))
    return {"value": n / t_loop, "unit": UNIT, "cores": 1, "kind": "port",
            "sample": "%d C3 samples, single thread: Python-loop sampling (sampler.py:30-34 restated) + NumPy "
                      "weights + np.histogram; host has %d cores" % (n, os.cpu_count() or 1),
            "vectorised_port_value": nv / t_vec,
            "vectorised_port_note": "%d samples with NumPy-vectorised sampling instead of the reference's loop, "
                                    "1 thread" % nv}


# ---------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------
class Job:
    """device buffers of one fused job and its three stages on torch's current stream"""

    def __init__(self, ig, edges, bc, precision):
        import torch
        from jetmontecarlo_b200 import _device as D
        from jetmontecarlo_b200 import _lib, fused
        from jetmontecarlo_b200.montecarlo.integrator import _bc_args
        self.D, self.L, self.torch = D, _lib, torch
        self.ig, self.s = ig, ig.c_struct()
        self.h_edges = np.ascontiguousarray(edges, dtype=np.float64)
        self.edges = D.to_device(self.h_edges)
        self.ne, self.nb = len(edges), len(edges) - 1
        self.prec = fused.PRECISIONS[precision]
        self.sum_w, self.sum_w2 = D.zeros((self.nb,)), D.zeros((self.nb,))
        self.count = D.zeros((self.nb,), dtype=torch.int64)
        self.out = [D.empty((self.nb,)), D.empty((self.nb,)), D.empty((self.nb + 1,)), D.empty((self.nb,))]
        self.bc_kind, self.bc_value = _bc_args(None, list(bc)) if not isinstance(bc, float) else _bc_args(bc, None)
        self.handle = _lib.default_handle(torch.cuda.current_device())
        self.area = float(ig.area)

    def run(self, n, first):
        D, L = self.D, self.L
        D.zero_(self.sum_w); D.zero_(self.sum_w2); D.zero_(self.count)     # cudaMemsetAsync (jmc_memset)
        L.check(L.lib.jmc_handle_run(self.handle.ptr, C.byref(self.s), n, SEED, first, self.prec,
                                     self.h_edges.ctypes.data_as(C.c_void_p), self.ne, D.ptr(self.sum_w),
                                     D.ptr(self.sum_w2), D.ptr(self.count), None, D.stream()))

    def finalize(self, n_total):
        D, L = self.D, self.L
        L.check(L.lib.jmc_finalize(D.ptr(self.sum_w), D.ptr(self.sum_w2), D.ptr(self.count), D.ptr(self.edges), self.ne,
                                   self.area, n_total, self.bc_kind, self.bc_value, *[D.ptr(o) for o in self.out],
                                   D.stream()))


def fp32_peaks(sm_count, sm_mhz):
    return 2 * 128 * sm_count * sm_mhz * 1e6 / 1e12, 16 * sm_count * sm_mhz * 1e6       # TFLOP/s, SFU op/s


def time_kind(name, ig, edges, bc, precision, n, sm_count, sm_mhz, contract=None, reps=5):
    """one fused integrand alone: CUDA events around `reps` whole jobs (zero, run, finalize) after 3 warm jobs"""
    import torch
    job = Job(ig, edges, bc, precision)
    for k in range(3):
        job.run(min(n, 10 ** 7), k * n)
        job.finalize(n)
    torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
          for _ in range(reps)]
    for k in range(reps):
        ev[k][0].record()
        job.run(n, (10 + k) * n)
        ev[k][1].record()
        job.finalize(n)
        ev[k][2].record()
    torch.cuda.synchronize()
    kern = statistics.median(a.elapsed_time(b) for a, b, _ in ev)
    whole = statistics.median(a.elapsed_time(c) for a, _, c in ev)
    rec = {"name": name, "samples_per_job": n, "precision": precision, "value": n / (whole * 1e-3), "unit": UNIT,
           "job_ms": whole, "kernel_ms": kern, "integral_first": float(job.out[2][0].item())}
    if contract:
        peak, sfu_peak = fp32_peaks(sm_count, sm_mhz)
        rate = n / (kern * 1e-3)
        rec["roofline"] = {"flop_eq_per_sample": contract[0], "sfu_per_sample": contract[1],
                           "frac_fp32": rate * contract[0] / 1e12 / peak, "frac_sfu": rate * contract[1] / sfu_peak,
                           "figure": "SURVEY 8(d) contract figure" if precision == "f32" else
                                     "SURVEY 8(d) figure against the FP32 peak (the FP64 mode runs on the FP64 pipe)"}
    return rec


def next_row_results(hbm_gbs):    # noqa: C901
    """SURVEY 8 "next" rows on the device, one short measurement each (CUDA events around the library call):
    f1 per-event Sudakov sampling, f3 RSS shower + replay of a stored event record, f4 fused EWOC."""
    import torch
    from jetmontecarlo_b200 import _device as D
    from jetmontecarlo_b200.fused import Integrand
    from jetmontecarlo_b200 import fused
    from jetmontecarlo_b200.numerics import sudakov_sampling as S
    from jetmontecarlo_b200.analytics.radiators.fixedcoupling import preRadAnalytic_fc_LL
    from jetmontecarlo_b200.utils import partonshower_utils as PS
    from jetmontecarlo_b200.montecarlo.ewoc import fused_ewoc
    from jetmontecarlo_b200.processes import ee2hadrons_LO as EE
    out = []

    def timed(fn, reps=3):
        best = None
        for _ in range(reps):
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
            ev[0].record(); r = fn(); ev[1].record(); torch.cuda.synchronize()
            ms = ev[0].elapsed_time(ev[1])
            best = ms if best is None else min(best, ms)
        return best, r

    # f1: one z_pre per event from exp(-R_pre(z, theta_crit)) on [0, z_cut], R_pre a 100 x 100 table
    gz, gt = np.logspace(-10, np.log10(Z_CUT), 100), np.logspace(-8, 0, 100)
    table = np.asarray(preRadAnalytic_fc_LL(gz[:, None] * np.ones((1, 100)), np.ones((100, 1)) * gt[None, :], Z_CUT, 'quark'))
    rad = S.TableRadiator(gz, gt, np.nan_to_num(table))
    n1 = 1_000_000
    theta = torch.exp(torch.rand(n1, dtype=torch.float64, device="cuda") * np.log(1e-6))
    ms, _ = timed(lambda: S.generate_precritical_samples(rad, theta, Z_CUT, seed=3))
    out.append({"name": "f1: per-event inverse-transform Sudakov samples (pre-critical, 100 x 100 radiator table, one event "
                        "per thread block, 1e6 events)", "kernel_ms": ms, "value": n1 / (ms * 1e-3), "unit": "events/s",
                "reference_note": "the reference loops over events in Python, one samples_from_cdf call each "
                                  "(~1e3 probe evaluations of a scipy interpolant): O(1e2) events/s"})
    # f3: RSS-groomed shower (grooming in the kernel's accumulator) and the replay of a stored event record
    edges = np.insert(np.logspace(-10, np.log10(.5), 100), 0, 1e-100)
    ig = Integrand.shower(beta=2., jet_type='quark', acc='MLL', groomer='rss', z_cut=Z_CUT, f=1., obs_acc='MLL', n_emissions=2)
    rec = time_kind("f3: veto shower with RSS grooming (f = 1, two-emission ECF), quark MLL, jets", ig, edges, 0., 'f32',
                    50_000_000, 0, 0.)
    rec["unit"] = "jets/s"
    out.append(rec)
    jets = PS.gen_jets(10_000_000, beta=2., jet_type='quark', acc='MLL', verbose=0, seed=5)
    lens, chains = jets.chains_device(max_emissions=32, precision='f32')
    ms, _ = timed(lambda: PS.ecfs_from_chains(lens, chains, beta=2., obs_acc='MLL', n_emissions=2, groomer='rss',
                                              z_cut=Z_CUT, f=1.))
    nbytes = chains.numel() * 8 + lens.numel() * 4 + lens.numel() * 8
    out.append({"name": "f3: getECFs_rss on a stored event record (1e7 jets x 32 emission slots), replay kernel", "kernel_ms": ms,
                "value": lens.numel() / (ms * 1e-3), "unit": "jets/s",
                "roofline": {"bound": "hbm", "achieved": nbytes / ms / 1e6, "peak": hbm_gbs, "unit": "GB/s",
                             "frac": nbytes / ms / 1e6 / hbm_gbs, "bytes_per_jet": nbytes / lens.numel()}})
    del chains, lens
    # f4: energy-weighted correlator of e+e- -> hadrons at LO, fused (nothing stored per event)
    proc = EE.ee2hadrons_LO(energy=1000., num_samples=8, sampleMethod='lin', seed=1)
    n4 = 100_000_000
    ms, _ = timed(lambda: fused_ewoc(proc, n4, EE.costheta_ij, energy_weights=(1, 1),
                                     edges=np.linspace(-1., 1., 101), seed=9))
    out.append({"name": "f4: fused EWOC of ee -> hadrons (hit-or-miss events -> 3 pairs -> energy-weighted histogram), "
                        "1e8 accepted events", "kernel_ms": ms, "value": n4 / (ms * 1e-3), "unit": "events/s"})
    return out


def sub_results(args, sm_count, sm_mhz):
    """the other BASELINE configurations, briefly, one GPU (driver-measured numbers for C1, C2, C4, C5,
    FP64 and the fixed-coupling control)"""
    import torch
    from jetmontecarlo_b200 import _device as D
    from jetmontecarlo_b200 import _lib, fused
    from jetmontecarlo_b200.fused import Integrand
    from jetmontecarlo_b200.utils.interpolation import lin_log_mixed_list
    e3, out = edges_c3(), []
    c3 = lambda **kw: Integrand.crit_sub('log', Z_CUT, EPS, beta=BETA, jet_type='quark', obs_acc='MLL', **kw)
    rad = lambda **kw: Integrand.radiator('log', EPS, beta=BETA, **kw)
    e1 = np.logspace(-11, np.log10(.5), 100)
    plan = [
        ("c3_rc_counted (np.histogram counts of every sample: the parity mode)", c3(fixedcoupling=False), e3,
         [1., 'minus'], 'f32', 10 ** 9, None),
        ("c3_fc (fixed-coupling control: every sample carries a weight)", c3(fixedcoupling=True), e3, [1., 'minus'],
         'f32', 10 ** 9, CONTRACT["c3"]),
        ("c3_rc_f64 (NumPy-order FP64 twin)", c3(fixedcoupling=False), e3, [1., 'minus'], 'f64', 10 ** 8, None),
        ("c3_fc_f64", c3(fixedcoupling=True), e3, [1., 'minus'], 'f64', 10 ** 8, CONTRACT["c3"]),
        ("c1 (ungroomed quark LL fixed coupling radiator), 1e9 samples", rad(jet_type='quark', fixedcoupling=True), e1,
         [0., 'plus'], 'f32', 10 ** 9, CONTRACT["c1"]),
        ("c1, 1e6 samples (BASELINE size)", rad(jet_type='quark', fixedcoupling=True), e1, [0., 'plus'], 'f32', 10 ** 6,
         CONTRACT["c1"]),
        ("c2 integration form, quark rc MLL, 1e8 samples", rad(jet_type='quark', fixedcoupling=False, obs_acc='MLL',
                                                              split_acc='MLL'), e1, [0., 'plus'], 'f32', 10 ** 8,
         CONTRACT["c2i"]),
        ("c2 integration form, gluon rc MLL, 1e8 samples", rad(jet_type='gluon', fixedcoupling=False, obs_acc='MLL',
                                                              split_acc='MLL'), e1, [0., 'plus'], 'f32', 10 ** 8,
         CONTRACT["c2i"]),
        ("c4 (pre x crit x sub, zero bin), 1.25e9 samples", Integrand.pre_crit_sub('log', Z_CUT, EPS), np.logspace(-12, np.log10(.5), 100),
         [1., 'minus'], 'f32', 1_250_000_000, CONTRACT["c4"]),
    ]
    for name, ig, edges, bc, prec, n, contract in plan:
        rec = time_kind(name, ig, edges, bc, prec, n, sm_count, sm_mhz, contract)
        if name.startswith("c3_rc_counted"):
            a = ACTIVE_SHARE_C3_RC
            fe = a * CONTRACT["c3"][0] + (1 - a) * ZERO_WEIGHT["all"][0]
            peak, _ = fp32_peaks(sm_count, sm_mhz)
            rec["roofline"] = {"flop_eq_per_sample": fe, "frac_fp32": n / (rec["kernel_ms"] * 1e-3) * fe / 1e12 / peak,
                               "figure": "blend: 0.3 % active samples at the 8(d) figure, the rest at the "
                                         "zero-weight figure of DESIGN.md 4.1"}
        out.append(rec)
    # C2 veto form: jets/s
    eshow = np.insert(np.logspace(-10, np.log10(.5), 100), 0, 1e-100)
    for jet in ('quark', 'gluon'):
        ig = Integrand.shower(beta=2., jet_type=jet, acc='MLL', groomer='ungroomed', obs_acc='MLL', n_emissions=1)
        rec = time_kind("c2 veto-algorithm shower, %s MLL, jets" % jet, ig, eshow, 0., 'f32', 5 * 10 ** 7, sm_count,
                        sm_mhz, None, reps=3)
        trials = 30. if jet == 'quark' else 57.
        peak, sfu_peak = fp32_peaks(sm_count, sm_mhz)
        rate = rec["samples_per_job"] / (rec["kernel_ms"] * 1e-3)
        rec["unit"] = "jets/s"
        rec["roofline"] = {"flop_eq_per_jet": 130. * trials + 40., "sfu_per_jet": 8. * trials,
                           "frac_fp32": rate * (130. * trials + 40.) / 1e12 / peak,
                           "frac_sfu": rate * 8. * trials / sfu_peak, "figure": "SURVEY 8(d): 130 flop-eq + 8 SFU per trial"}
        out.append(rec)
    ig = Integrand.shower(beta=2., jet_type='quark', acc='MLL', groomer='softdrop', z_cut=.1, obs_acc='MLL', n_emissions=2)
    rec = time_kind("c3 (iii) veto shower, Soft Drop two-emission ECF, quark MLL, jets", ig, eshow, 0., 'f32',
                    5 * 10 ** 7, sm_count, sm_mhz, None, reps=3)
    rec["unit"] = "jets/s"
    out.append(rec)
    # C1 at the BASELINE size through the host-buffer entry point: wall time per job
    ig = rad(jet_type='quark', fixedcoupling=True)
    s, nb = ig.c_struct(), len(e1) - 1
    bufs = [np.empty(nb), np.empty(nb), np.empty(nb + 1), np.empty(nb), np.empty(nb), np.empty(nb), np.empty(nb, dtype=np.uint64)]
    P = lambda a: a.ctypes.data_as(C.c_void_p)
    # (argument objects built once: a caller that issues many jobs does not re-wrap its buffers per call)
    fn, s_ref, e_ptr, b_ptrs = _lib.lib.jmc_integrate_host, C.byref(s), P(e1), [P(b) for b in bufs]
    call = lambda k: _lib.check(fn(s_ref, 10 ** 6, SEED, k * 10 ** 6, _lib.F32, e_ptr, nb + 1, _lib.BC_LAST_PLUS, 0., *b_ptrs))
    for k in range(20):
        call(k)
    t0 = time.perf_counter()
    for k in range(200):
        call(20 + k)
    dt = (time.perf_counter() - t0) / 200
    out.append({"name": "c1, 1e6 samples per job through jmc_integrate_host (host edges in, host results out)",
                "wall_us_per_job": dt * 1e6, "value": 1e6 / dt, "unit": UNIT, "integral_first": float(bufs[2][0])})
    # C5: theta_crit grid of gen_crit_sub_num_rad, the reference's defaults (examples/params.py:67,86)
    npts, nbins, n5 = 5000, 5000, 5_000_000
    grid = lin_log_mixed_list(1e-10, 1., npts)
    igs = [rad(jet_type='quark', fixedcoupling=True, theta_scale=float(t)) for t in grid]
    kw = dict(numbins=nbins, binspacing='log', min_log_bins=grid ** 2. * 1e-15, seed=1, precision='f32', last_bc=[0., 'plus'])
    fused.integrate_batch(igs[:4], 100000, **dict(kw, min_log_bins=kw['min_log_bins'][:4]))
    torch.cuda.synchronize(); t0 = time.perf_counter()
    fused.integrate_batch(igs, n5, **kw)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    out.append({"name": "c5: %d theta_crit points x %.0e samples x %d bins (min/max pass + histogram pass + results to "
                        "the host)" % (len(grid), n5, nbins), "wall_s": dt,
                "value": 2 * len(grid) * n5 / dt, "unit": "sample-evaluations/s"})
    # replay path: jmc_hist on device arrays larger than L2 (HBM roofline)
    nh = 250_000_000
    g = torch.Generator(device="cuda").manual_seed(1)
    obs = torch.exp(torch.rand(nh, dtype=torch.float64, device="cuda", generator=g) * np.log(1e-11))
    w = torch.rand(nh, dtype=torch.float64, device="cuda", generator=g)
    e = D.to_device(e1)
    sw, sw2, cnt = D.zeros((nb,)), D.zeros((nb,)), D.zeros((nb,), dtype=torch.int64)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    ms = []
    for rep in range(4):
        ev[0].record()
        _lib.check(_lib.lib.jmc_hist(D.ptr(obs), D.ptr(w), nh, D.ptr(e), nb + 1, D.ptr(sw), D.ptr(sw2), D.ptr(cnt), D.stream()))
        ev[1].record(); torch.cuda.synchronize()
        ms.append(ev[0].elapsed_time(ev[1]))
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except (OSError, ValueError):
        pass
    hbm = peaks.get("hbm_gbs", 6550.)
    gbs = 16 * nh / min(ms[1:]) / 1e6
    out.append({"name": "replay: jmc_hist (setDensity on device arrays), %.1e samples, every weight non-zero, 99 log bins" % nh,
                "kernel_ms": min(ms[1:]), "value": nh / (min(ms[1:]) * 1e-3), "unit": UNIT,
                "roofline": {"bound": "hbm", "achieved": gbs, "peak": hbm, "unit": "GB/s", "frac": gbs / hbm,
                             "bytes_per_sample": 16,
                             "peak_source": "MEASURED_PEAKS.json hbm_gbs" if "hbm_gbs" in peaks else "recipe fallback 6550"}})
    try:
        out.extend(next_row_results(hbm))
    except Exception as exc:          # the SURVEY 8(f) rows must never take the C1-C5 records down
        out.append({"name": "next rows (SURVEY 8 f1-f4)", "error": repr(exc)})
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--samples", type=float, default=1e9, help="samples per GPU per job")
    ap.add_argument("--jobs", type=int, default=50, help="jobs per step")
    ap.add_argument("--precision", default="f32", choices=["f32", "f64"])
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--cpu-samples", type=float, default=0, help="reference arm: samples per process per step")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-sub-results", action="store_true")
    ap.add_argument("--variant", default="c3_rc", choices=["c3_rc", "c3_fc"])
    ap.add_argument("--counts", default="weighted", choices=["weighted", "all"],
                    help="per-bin counts over the weighted samples only (default) or np.histogram's count of every sample")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup

    if args.impl == "reference":
        return run_reference_arm(args)

    import torch
    import torch.distributed as dist
    from jetmontecarlo_b200 import _lib, fused
    from jetmontecarlo_b200.fused import Integrand

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product has no CPU path")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    assert world == args.gpus or world == 1, "launch with torchrun --nproc-per-node = --gpus"

    n, J = int(args.samples), int(args.jobs)
    ig = Integrand.crit_sub('log', Z_CUT, EPS, beta=BETA, jet_type='quark', fixedcoupling=(args.variant == "c3_fc"),
                            obs_acc='MLL', counts=args.counts)
    h_edges = edges_c3()
    nb = N_EDGES - 1
    job = Job(ig, h_edges, [1., 'minus'], args.precision)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    packed = torch.empty((3, nb), dtype=torch.float64, device="cuda")

    def one_job(index):
        """zero -> jmc_handle_run on this rank's shard -> [all-reduce] -> jmc_finalize; all on torch's current stream"""
        job.run(n, (index * world + rank) * n)       # every job and rank draws fresh sample indices
        if world > 1:
            torch.stack([job.sum_w, job.sum_w2, job.count.to(torch.float64)], out=packed)
            dist.all_reduce(packed)
            job.sum_w.copy_(packed[0]); job.sum_w2.copy_(packed[1]); job.count.copy_(packed[2].to(torch.int64))
        job.finalize(n * world)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for k in range(args.warmup):
        for j in range(J):
            one_job(k * J + j)
    barrier()

    # ---- device-timed region: K steps of J jobs, CUDA events per step, L2 flushed between steps
    clocks = ClockSampler(local_rank) if rank == 0 else None
    launches0 = _lib.launch_count()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
          for _ in range(args.steps)]
    barrier()
    t_wall0 = time.perf_counter()
    for k in range(args.steps):
        job.D.zero_(flush)
        ev[k][0].record()
        base = (args.warmup + k) * J
        # the first job of the step carries the kernel-only events
        job.D.zero_(job.sum_w); job.D.zero_(job.sum_w2); job.D.zero_(job.count)
        kev0, kev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        kev0.record()
        _lib.check(_lib.lib.jmc_handle_run(job.handle.ptr, C.byref(job.s), n, SEED, (base * world + rank) * n, job.prec,
                                           job.h_edges.ctypes.data_as(C.c_void_p), job.ne, job.D.ptr(job.sum_w),
                                           job.D.ptr(job.sum_w2), job.D.ptr(job.count), None, job.D.stream()))
        kev1.record()
        ev[k] = (ev[k][0], (kev0, kev1), ev[k][2])
        if world > 1:
            torch.stack([job.sum_w, job.sum_w2, job.count.to(torch.float64)], out=packed)
            dist.all_reduce(packed)
            job.sum_w.copy_(packed[0]); job.sum_w2.copy_(packed[1]); job.count.copy_(packed[2].to(torch.int64))
        job.finalize(n * world)
        for j in range(1, J):
            one_job(base + j)
        ev[k][2].record()
    barrier()
    t_wall1 = time.perf_counter()
    launches = _lib.launch_count() - launches0
    step_ms = [a.elapsed_time(c) for a, _, c in ev]
    kern_ms = [k0.elapsed_time(k1) for _, (k0, k1), _ in ev]
    t_rank = torch.tensor([sum(step_ms), statistics.median(kern_ms)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t_rank, op=dist.ReduceOp.MAX)
    total_ms, kern_job_ms = t_rank.tolist()
    clock_info = clocks.stop(t_wall0, t_wall1) if clocks else None
    value = n * J * world * args.steps / (total_ms * 1e-3)
    cdf0 = float(job.out[2][0].item())

    # ---- end to end through the public API with host buffers: J jobs per step, host edges in, host results out
    barrier()
    e2e_steps = args.steps
    if world == 1:
        hd, hde, hi, hie = (np.empty(nb), np.empty(nb), np.empty(nb + 1), np.empty(nb))
        hsw, hsw2, hc = np.empty(nb), np.empty(nb), np.empty(nb, dtype=np.uint64)
        P = lambda a: a.ctypes.data_as(C.c_void_p)
        nbytes_in = h_edges.nbytes * J
        nbytes_out = (8 * (4 * nb + 1) + 8 * 3 * nb) * J
        e2e_path = "jmc_integrate_host (C ABI, host buffers), %d jobs per step" % J

        def e2e_job(index):
            _lib.check(_lib.lib.jmc_integrate_host(C.byref(job.s), n, SEED, (10 ** 6 + index) * n, job.prec, P(h_edges),
                                                   N_EDGES, _lib.BC_LAST_MINUS, 1., P(hd), P(hde), P(hi), P(hie),
                                                   P(hsw), P(hsw2), P(hc)))
            return hi[0]
    else:
        pinned = torch.from_numpy(h_edges).pin_memory()
        nbytes_in = (h_edges.nbytes * 2 + 8 * 2 * nb) * J
        nbytes_out = (8 * (2 * nb) + 8 * (2 * nb + 1)) * J
        e2e_path = ("jetmontecarlo_b200.fused.integrate (host edges -> device, jmc_handle_run, NCCL all-reduce, "
                    "jmc_finalize/jmc_cumulative, results -> host), %d jobs per step" % J)

        def e2e_job(index):
            it = fused.integrate(ig, n * world, edges=pinned.numpy(), seed=SEED + 1000 + index,
                                 precision=args.precision, last_bc=[1., 'minus'])
            return it.integral[0]
    e2e_job(-1)
    barrier()
    t0 = time.perf_counter()
    for k in range(e2e_steps):
        for j in range(J):
            e2e_job(k * J + j)
    torch.cuda.synchronize()
    t_e2e = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t_e2e, op=dist.ReduceOp.MAX)
    e2e_value = n * J * world * e2e_steps / t_e2e.item()

    if rank == 0:
        kern_s = kern_job_ms * 1e-3
        per_gpu_rate = n / kern_s
        sm_mhz = (clock_info or {}).get("sm_mhz") or 1965.
        sm_count = _lib.lib.jmc_sm_count()
        fp32_peak, sfu_peak = fp32_peaks(sm_count, sm_mhz)
        if args.variant == "c3_rc":
            a = ACTIVE_SHARE_C3_RC       # measured share of non-NaN weights (SURVEY 8d caveat; tests check it)
            flop_eq = a * CONTRACT["c3"][0] + (1 - a) * ZERO_WEIGHT[args.counts][0]
            sfu = a * CONTRACT["c3"][1] + (1 - a) * ZERO_WEIGHT[args.counts][1]
        else:
            flop_eq, sfu = CONTRACT["c3"]
        achieved = per_gpu_rate * flop_eq / 1e12
        roofline = {
            "bound": "fp32", "achieved": achieved, "peak": fp32_peak, "unit": "TFLOP/s", "frac": achieved / fp32_peak,
            "frac_is": "samples/s x algorithmic flop-equivalents per sample (SURVEY 8(d) counting rule; blend of 0.3 % "
                       "active samples at the contract figure 390 and 99.7 % zero-weight samples at what the "
                       "reference's NaN -> 0 semantics make necessary, DESIGN.md 4.1) / FP32 peak",
            "traffic": None,
            "kernel": "run_fast_kernel" if args.precision == "f32" else "run_exact_kernel",
            "kernel_ms_per_launch": kern_s * 1e3, "flop_eq_per_sample": flop_eq, "sfu_per_sample": sfu,
            "sfu_frac": per_gpu_rate * sfu / sfu_peak,
            "frac_contract_390": per_gpu_rate * CONTRACT["c3"][0] / 1e12 / fp32_peak,
            "frac_contract_390_is": "the same with SURVEY 8(d)'s 390 flop-eq charged to EVERY sample: above 1 by "
                                    "construction, because the kernel does not evaluate radiators for samples whose "
                                    "reference weight is NaN -> 0",
            "peak_source": "2 x 128 FP32 lanes x %d SMs x %.0f MHz (clock sampled during the timed region); "
                           "MEASURED_PEAKS.json has no FP32/SFU figure, its HBM peak does not bound a kernel "
                           "with ~0 B/sample of DRAM traffic" % (sm_count, sm_mhz),
        }
        try:      # ncu-derived companions of the figure above (DRAM traffic, issue-slot utilisation, pipes)
            with open(os.path.join(ROOT, "profiles", "roofline_inputs.json")) as f:
                prof = json.load(f)["run_fast_kernel" if args.counts == "weighted" else "run_fast_kernel_counted"]
            if args.precision == "f32" and args.variant == "c3_rc":
                roofline["traffic"] = prof.get("dram_bytes_per_launch")
                roofline["ncu"] = prof
                ipw = prof.get("warp_instructions_per_sample")
                if ipw:
                    # executed warp-instructions per warp of 32 samples (ncu) x warps/s / issue slots/s
                    roofline["issue_slot_utilisation"] = (per_gpu_rate / 32.) * ipw / (4 * sm_count * sm_mhz * 1e6)
                    roofline["issue_slot_utilisation_is"] = ("executed warp-instructions per sample (ncu source page of "
                                                             "the committed profile) x measured samples/s / (4 "
                                                             "schedulers x SMs x clock)")
                # the pipe that actually binds: the generator's 32x32->64 multiplies on the FMA-heavy pipe.  A
                # kernel that only generates the Philox blocks takes generator_only_ms on this GPU; this kernel
                # (generator + sampling + Landau test [+ observable + histogram]) is measured against it.
                floor_ms = prof.get("generator_only_ms_per_1e9_samples")
                if floor_ms:
                    roofline["generator_floor"] = {
                        "ms_per_1e9_samples": floor_ms, "frac": floor_ms * (n / 1e9) / (kern_s * 1e3),
                        "source": prof.get("generator_only_source")}
        except (OSError, KeyError, ValueError):
            pass
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": total_ms / args.steps, "ms_per_job": total_ms / args.steps / J,
            "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": args.precision, "data": "synthetic",
            "config": workload_config(args, world),
            "mode": {"variant": args.variant, "counts": args.counts,
                     "counts_note": "weighted: the per-bin count is taken over the samples with a non-zero weight "
                                    "(density, error and integral are unchanged: the reference uses the count only to "
                                    "zero empty bins, integrator.py:101,123-124); all: np.histogram's count of every "
                                    "sample (sub_results: c3_rc_counted)",
                     "timed_region_s": total_ms * 1e-3, "e2e_path": e2e_path},
            "result": {"cdf_first_bin": cdf0},
            "clocks": clock_info,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(nbytes_in),
                    "d2h_bytes_per_step": int(nbytes_out)},
            "gpu_launches": int(launches),
            "roofline": roofline,
        }
        if world == 1 and not args.no_sub_results:
            try:
                line["sub_results"] = sub_results(args, sm_count, sm_mhz)
            except Exception as exc:      # a sub-measurement must never take the headline down
                line["sub_results_error"] = repr(exc)
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline_quick()
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
