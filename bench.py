#!/usr/bin/env python
"""Benchmark of the Krotov sweep hot path.

Metric (BASELINE.json): Krotov iterations/sec on the 128-objective, nt=1000
two-level ensemble (configs[3], "C4").  One *step* = one Krotov iteration =
chi boundary -> backward sweep -> fused update/forward sweep -> tau
(/root/reference/src/krotov/optimize.py:393-508).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

Our arm prints ONE JSON line with
  value      iterations/s of the ONE 128-objective problem with all inputs
             resident in HBM, timed with CUDA events on the launching stream,
             max over ranks.  With --gpus N > 1 the objectives of that one
             problem are sharded over the GPUs (GPUShards mode 'sharded': one
             in-kernel NVLink exchange of nt doubles per fixed-point round):
             strong scaling, reported as measured;
  e2e        the same metric through the public krotov_b200.optimize_pulses
             call with HOST (numpy) objectives and a per-iteration host hook;
  roofline   algorithmic HBM bytes of the dominant kernel / its measured
             duration against MEASURED_PEAKS.json, plus what actually limits
             the kernel and an FP64 fraction against a measured DFMA peak;
  cpu_baseline  the reference's loop timed on this box's host cores: numpy
             port (serial sample and multi-process on all cores), the
             UNMODIFIED reference from baseline/_ref (bounded sample) and a
             C/OpenMP port;
  parity     max relative deviation of the updated pulses (iterations 1..4,
             full workload) between the CUDA path and the CPU port of the
             same run (tolerance 1e-10);
  configs    (N = 1) the other BASELINE configs C1, C2, C3, C5: it/s, e2e and
             their CPU baselines from the same run;
  weak_scaling_ensemble, replicas, sharded_parity  (N > 1) extra keys: one
             ensemble of 512*N objectives sharded over the GPUs against 512 on
             one GPU; N independent 128-objective optimisations; the sharded
             pulses against the single-GPU pulses.
The reference arm (--impl reference) times the reference's CPU path: the
faster of the numpy port (serial / multi-process on all host cores) and the
unmodified reference package, see DESIGN.md section 6.
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

METRIC = "krotov_iterations_per_sec"
UNIT = "it/s"
WORKLOAD = dict(workload="C4_tls_ensemble", K=128, N=2, nt=1000, L=1,
                chi="chis_re", propagator="expm")
# factory in krotov_b200.workloads, objectives sampled for the CPU legs
CONFIGS = {
    'C1': ('tls_state_to_state', {}),
    'C2': ('transmon_xgate', {}),
    'C3': ('two_qubit_gate', {}),
    'C4': ('tls_ensemble', {'K': 128, 'nt': 1000}),
    'C5': ('dissipative_qubit_reset', {}),
}
WEAK_K_PER_GPU = 512     # 4 objectives per CTA on 128 SMs


def load_peaks():
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(path):
        with open(path) as fh:
            return json.load(fh), "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0}, "fallback"


class ClockSampler:
    """SM clock / throttle-reason sampling during the timed region.  NVML is
    polled from a thread every ~2 ms (the timed region of the fused kernel is
    only a few milliseconds long, too short for `nvidia-smi -lms`); falls back
    to one nvidia-smi query if NVML is unavailable."""

    REASONS = ((0x8, 'hw_slowdown'), (0x40, 'hw_thermal_slowdown'),
               (0x20, 'sw_thermal_slowdown'), (0x4, 'sw_power_cap'))

    def __init__(self, index=0):
        self.index, self.rows, self.thread = index, [], None
        self.stop_flag = threading.Event()
        self.nvml = self.handle = None
        self.sm_max = None

    def _sample(self):
        n = self.nvml
        sm = n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM)
        try:
            mask = n.nvmlDeviceGetCurrentClocksEventReasons(self.handle)
        except Exception:
            mask = n.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle)
        self.rows.append((float(sm), int(mask)))

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nvml = pynvml
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(self.index)
            self.sm_max = float(pynvml.nvmlDeviceGetMaxClockInfo(
                self.handle, pynvml.NVML_CLOCK_SM))
            self._sample()

            def loop():
                while not self.stop_flag.is_set():
                    try:
                        self._sample()
                    except Exception:
                        break
                    time.sleep(0.002)
            self.thread = threading.Thread(target=loop, daemon=True)
            self.thread.start()
        except Exception:
            self.nvml = None

    def stop(self):
        if self.nvml is not None:
            self.stop_flag.set()
            if self.thread is not None:
                self.thread.join(timeout=2)
            try:
                self._sample()
            except Exception:
                pass
            sm = [r[0] for r in self.rows]
            reasons = sorted({name for _, mask in self.rows
                              for bit, name in self.REASONS if mask & bit})
            return {"sm_mhz": float(np.median(sm)) if sm else None,
                    "sm_max_mhz": self.sm_max, "reasons": reasons,
                    "samples": len(sm), "source": "nvml"}
        try:
            out = subprocess.run(
                ['nvidia-smi', '-i', str(self.index),
                 '--query-gpu=clocks.sm,clocks.max.sm',
                 '--format=csv,noheader,nounits'],
                capture_output=True, text=True, timeout=10).stdout
            sm, smax = [float(x) for x in out.strip().split(',')]
            return {"sm_mhz": sm, "sm_max_mhz": smax, "reasons": [],
                    "samples": 1, "source": "nvidia-smi (after the run)"}
        except Exception:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [],
                    "samples": 0}


def make_workload(label, replica=0, **override):
    """Workload of a BASELINE config label ('C1'..'C5'); `replica` > 0: the
    same C4 ensemble with a different guess amplitude (an independent
    optimisation per GPU); `override`: factory arguments (e.g. K)."""
    import krotov_b200 as krotov
    factory, kwargs = CONFIGS[label]
    kwargs = dict(kwargs, **override)
    if label == 'C4' and replica:
        kwargs['ampl0'] = 0.2 * (1 + 0.05 * replica)
    return getattr(krotov.workloads, factory)(**kwargs)


def build_workload(replica=0):
    """The contract workload described by WORKLOAD (C4 unless --workload)."""
    label = WORKLOAD['workload'][:2]
    if label == 'C4':
        return make_workload('C4', replica, K=WORKLOAD['K'], nt=WORKLOAD['nt'])
    wl = make_workload(label)
    low = wl.lowered()
    WORKLOAD.update(K=wl.K, N=len(low['psi0'][0]), nt=wl.nt,
                    chi='chis_' + wl.chi)
    return wl


def chi_of(krotov, wl):
    if wl.chi == 'qubit_reset':
        fixed = wl.meta['chi_fixed']
        return lambda fw_states_T, objectives, tau_vals: [
            fixed.copy() for _ in fw_states_T]
    return getattr(krotov.functionals, 'chis_' + wl.chi)


# --------------------------------------------------------------------------
# CPU side: ports of the reference loop and the unmodified reference

def _oracle_chi(wl):
    from oracle import krotov_oracle as orc
    if wl.chi == 'qubit_reset':
        fixed = wl.meta['chi_fixed'].reshape(-1, order='F')
        return lambda fw, tg, tau, w: [fixed.copy() for _ in fw]
    return getattr(orc, 'chis_' + wl.chi)


def time_oracle(wl, iters, k_sample=None):
    """Seconds per Krotov iteration of the numpy oracle (serial, 1 core) on
    the first `k_sample` objectives of `wl`, scaled to the full K (the
    reference loop is linear in the number of objectives)."""
    from oracle import krotov_oracle as orc
    try:
        import threadpoolctl
        limiter = threadpoolctl.threadpool_limits(limits=1)
    except Exception:
        limiter = None
    low = wl.lowered()
    K = len(low['terms'])
    ks = K if k_sample is None else max(1, min(K, k_sample))
    terms, psi0, targets = low['terms'][:ks], low['psi0'][:ks], \
        low['targets'][:ks]
    pulses = [p.copy() for p in low['pulses']]
    sup = bool(low['is_super'])
    chi = _oracle_chi(wl)
    fw_T = [orc.forward_propagation(terms[k], pulses, low['tlist'], psi0[k],
                                    sup, store_all=False)
            for k in range(ks)]
    tau = np.array([np.vdot(targets[k], fw_T[k]) for k in range(ks)])
    times = []
    for _ in range(iters):
        t0 = time.perf_counter()
        rec = orc.krotov_iteration(
            terms, psi0, targets, pulses, low['shapes'], low['lambdas'],
            low['tlist'], fw_T, tau, chi, sup)
        times.append(time.perf_counter() - t0)
        pulses = rec['optimized_pulses']
        fw_T, tau = rec['fw_states_T'], rec['tau_vals']
    if limiter is not None:
        limiter.restore_original_limits()
    per_iter = float(np.mean(times)) * (K / ks)
    return per_iter, ks


def time_oracle_parallel(wl, iters, warmup=1, dump=None):
    """Seconds per Krotov iteration of the multi-process numpy port (the
    reference's parallel mode, parallelization.py:51-57: backward sweep
    parallel over the objectives, update/forward sweep synchronised per time
    step) on ALL host cores, full workload."""
    from oracle import krotov_oracle as orc
    from oracle.krotov_oracle_mp import ParallelOracle
    low = wl.lowered()
    K = len(low['terms'])
    pulses = [p.copy() for p in low['pulses']]
    with ParallelOracle(low['terms'], low['psi0'], low['targets'],
                        low['shapes'], low['lambdas'], low['tlist']) as po:
        fw_T = po.forward(pulses)
        tau = np.array([np.vdot(low['targets'][k], fw_T[k])
                        for k in range(K)])
        times, history = [], []
        for it in range(warmup + iters):
            t0 = time.perf_counter()
            rec = po.iteration(pulses, fw_T, tau, orc.chis_re)
            if it >= warmup:
                times.append(time.perf_counter() - t0)
            pulses, fw_T, tau = rec['optimized_pulses'], \
                rec['fw_states_T'], rec['tau_vals']
            history.append(np.array(pulses))
        nproc = po.nproc
    if dump:   # pulses after Krotov iterations 1, 2, ... (parity check)
        np.save(dump, np.array(history))
    return float(np.mean(times)), nproc


def parallel_leg(wl, iters, warmup=1, in_subprocess=False, dump=None):
    """cpu_baseline-style dict for the multi-process port.  With
    `in_subprocess` the workers are forked from a fresh interpreter (never
    from a process that holds a CUDA context)."""
    if in_subprocess:
        out = subprocess.run(
            [sys.executable, os.path.abspath(__file__), '--cpu-parallel-leg',
             '--steps', str(iters), '--warmup', str(warmup)] +
            (['--dump-pulses', dump] if dump else []),
            capture_output=True, text=True, timeout=600)
        for ln in reversed(out.stdout.strip().splitlines()):
            if ln.startswith('{'):
                return json.loads(ln)
        raise RuntimeError('parallel leg failed: ' + out.stderr[-300:])
    per_iter, nproc = time_oracle_parallel(wl, iters, warmup, dump)
    return {"value": 1.0 / per_iter, "unit": UNIT, "cores": nproc,
            "kind": "port", "host_cpus": os.cpu_count(),
            "sample": "%d Krotov iterations of the full workload (K=%d, "
                      "nt=%d), numpy port in the reference's multi-process "
                      "mode (parallelization.py:51-57): %d worker processes, "
                      "one pipe round trip per time step"
                      % (iters, WORKLOAD['K'], WORKLOAD['nt'], nproc)}


def time_c_oracle(wl, iters=3):
    """Iterations/s of the multi-threaded C restatement of the path
    (oracle/krotov_oracle_c.c, OpenMP over objectives, own Pade expm) on the
    full workload: best of 1 thread and all host threads."""
    os.environ.setdefault('OMP_WAIT_POLICY', 'ACTIVE')
    os.environ.setdefault('GOMP_SPINCOUNT', '100000000')
    os.environ.setdefault('OMP_PROC_BIND', 'true')
    from oracle import krotov_oracle as orc
    from oracle import krotov_oracle_c as coc
    low = wl.lowered()
    ncpu = min(coc.max_threads(), os.cpu_count() or 1)
    best = None
    for threads in sorted({1, ncpu}):
        co = coc.COracle(low, nthreads=threads)
        pulses = np.array(low['pulses'])
        phiT = co.forward(pulses)
        tau = np.einsum('kn,kn->k', co.targets.conj(), phiT)
        times = []
        for _ in range(iters):
            chis = orc.chis_re(list(phiT), list(co.targets), list(tau), None)
            t0 = time.perf_counter()
            rec = co.iteration(pulses, chis)
            times.append(time.perf_counter() - t0)
            pulses, phiT, tau = rec['optimized_pulses'], \
                rec['fw_states_T'], rec['tau_vals']
        val = 1.0 / float(np.min(times))
        if best is None or val > best[0]:
            best = (val, threads)
    return {"value": best[0], "unit": UNIT, "cores": best[1], "kind": "port",
            "host_cpus": os.cpu_count(),
            "sample": "%d Krotov iterations of the full workload, C/OpenMP "
                      "restatement (oracle/krotov_oracle_c.c), best of 1 and "
                      "%d threads" % (iters, ncpu)}


def time_unmodified_reference(label, k_sample, iters, **override):
    """cpu_baseline-style dict for the UNMODIFIED reference package
    (baseline/_ref, or /root/reference/src in the build container), run in a
    process of its own through oracle/run_reference.py: its own
    optimize_pulses loop with the numpy plugins of its notebook 09, serial
    (optimize.py:233-238 pins BLAS to one thread).  Bounded sample: the first
    `k_sample` objectives, time scaled linearly in K."""
    factory, kwargs = CONFIGS[label]
    kwargs = dict(kwargs, **override)
    try:
        out = subprocess.run(
            [sys.executable, '-m', 'oracle.run_reference', factory,
             str(k_sample), str(iters), json.dumps(kwargs)],
            capture_output=True, text=True, timeout=600, cwd=ROOT)
        d = None
        for ln in reversed(out.stdout.strip().splitlines()):
            if ln.startswith('{'):
                d = json.loads(ln)
                break
        if d is None:
            return {"unavailable": out.stderr.strip()[-300:] or "no output"}
    except Exception as exc:  # pragma: no cover
        return {"unavailable": repr(exc)}
    return {"value": 1.0 / d['seconds_per_iteration'], "unit": UNIT,
            "cores": 1, "kind": "reference", "host_cpus": os.cpu_count(),
            "sample": "%d Krotov iteration(s) of the first %d of %d "
                      "objectives at nt=%d through the unmodified "
                      "krotov.optimize_pulses (%s, numpy plugins of notebook "
                      "09), time scaled by %d/%d"
                      % (d['iterations'], d['k_sample'], d['K'], d['nt'],
                         d['reference_path'], d['K'], d['k_sample'])}


def cpu_legs_for(label, wl, budget_s=6.0):
    """Bounded CPU baselines of one config: numpy port (serial) and the
    unmodified reference."""
    K = wl.K
    # ~seconds per (objective, iteration) of the port: scale the sample
    per_obj = {'C1': 0.02, 'C2': 0.04, 'C3': 0.08, 'C4': 0.04, 'C5': 0.5}
    ks = int(max(1, min(K, budget_s / per_obj.get(label, 0.1))))
    per_iter, ks = time_oracle(wl, 1, k_sample=ks)
    port = {"value": 1.0 / per_iter, "unit": UNIT, "cores": 1,
            "kind": "port", "host_cpus": os.cpu_count(),
            "sample": "1 Krotov iteration of the first %d of %d objectives "
                      "at nt=%d (numpy port, serial), time scaled by %d/%d"
                      % (ks, K, wl.nt, K, ks)}
    ref = time_unmodified_reference(label, ks, 1)
    return port, ref


def run_reference_arm(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    wl = build_workload()
    label = WORKLOAD['workload'][:2]
    total = args.steps + args.warmup
    # bounded sample: aim at <= ~150 s for the whole run
    per_obj_iter = 0.05   # ~s per (objective, iteration), refined below
    budget = 150.0 / max(total, 1)
    ks = int(max(1, min(WORKLOAD['K'], budget / per_obj_iter)))
    per_iter_w, ks = time_oracle(wl, max(args.warmup, 1), ks)
    per_iter, ks = time_oracle(wl, args.steps, ks)
    value = 1.0 / per_iter
    sample = ("%d Krotov iterations of the first %d of %d objectives, "
              "nt=%d, time scaled by %d/%d (loop is linear in K)"
              % (args.steps, ks, WORKLOAD['K'], WORKLOAD['nt'],
                 WORKLOAD['K'], ks))
    serial = {"value": value, "unit": UNIT, "cores": 1, "kind": "port",
              "sample": sample, "host_cpus": os.cpu_count()}
    best = serial
    parallel = None
    if label == 'C4':
        try:   # the reference's multi-process mode on all host cores
            parallel = parallel_leg(wl, args.steps, max(args.warmup, 1))
            if parallel["value"] > value:
                best = parallel
                value, per_iter = parallel["value"], 1.0 / parallel["value"]
        except Exception as exc:  # pragma: no cover
            parallel = {"unavailable": repr(exc)}
    # the unmodified reference package (bounded sample)
    override = dict(K=WORKLOAD['K'], nt=WORKLOAD['nt']) if label == 'C4' else {}
    unmodified = time_unmodified_reference(
        label, min(WORKLOAD['K'], 16), min(max(args.steps, 1), 2), **override)
    if unmodified.get("value", 0.0) > value:
        best = unmodified
        value, per_iter = unmodified["value"], 1.0 / unmodified["value"]
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT,
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": per_iter * 1e3, "higher_is_better": True,
        "scaling": "strong" if args.gpus > 1 else "weak",
        "vs_baseline": None, "dtype": "c128",
        "data": "synthetic", "config": dict(WORKLOAD),
        "cpu_baseline": best,
        "cpu_baseline_serial": serial,
        "cpu_baseline_parallel": parallel,
        "cpu_baseline_reference": unmodified,
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0,
                "d2h_bytes_per_step": 0},
    }
    if label == 'C4':
        try:   # extra information: an optimised multi-threaded C port
            line["cpu_baseline_c"] = time_c_oracle(wl)
        except Exception as exc:  # pragma: no cover
            line["cpu_baseline_c"] = {"unavailable": repr(exc)}
    print(json.dumps(line))


# --------------------------------------------------------------------------
# GPU side

def measure_fp64_peak():
    """Measured FP64 FMA throughput of the device (TFLOP/s) from the
    micro-benchmark profiles/micro/fp64_peak (built by
    __graft_entry__.build()); None if the binary is missing."""
    exe = os.path.join(ROOT, 'profiles', 'micro', 'fp64_peak')
    if not os.path.exists(exe):
        return None
    try:
        out = subprocess.run([exe], capture_output=True, text=True,
                             timeout=60).stdout
        for ln in reversed(out.strip().splitlines()):
            if ln.startswith('{'):
                return json.loads(ln)
    except Exception:
        pass
    return None


def algorithmic_flops(K, N, NT, M, L):
    """F_alg of SURVEY.md section 8(d) (dense-expm formulation, real flops,
    complex MAC = 8, Pade-13 without squaring as the upper bound)."""
    f_step = 8.0 * M * N * N + (6 + 4.0 / 3) * 8 * N ** 3 + 8 * N * N
    return 2.0 * K * NT * f_step + K * NT * 8.0 * L * (N * N + N)


class DeviceRun:
    """Device-resident Krotov iterations of one workload on this rank's GPU:
    the whole problem (dist None), or this rank's block of objectives of a
    problem sharded over the ranks (`sharded`)."""

    def __init__(self, krotov, wl, dist=None, sharded=False, engine='auto',
                 flush_t=None):
        import torch
        from krotov_b200.compiler import compile_problem, initialize_controls
        from krotov_b200.engine import SweepEngine
        from krotov_b200.parallelization import ShardComm, shard_bounds
        self.torch, self.krotov, self.wl, self.dist = torch, krotov, wl, dist
        objectives = wl.objectives(krotov.Objective)
        (controls, _, guess_pulses, mapping, lam, shp) = initialize_controls(
            objectives, wl.pulse_options, wl.tlist)
        K = len(objectives)
        self.K_total = K
        lo, hi = 0, K
        world = dist.get_world_size() if (dist is not None and sharded) else 1
        if world > 1:
            lo, hi = shard_bounds(K, world, dist.get_rank())
        self.cp = cp = compile_problem(objectives[lo:hi], controls,
                                       mapping[lo:hi], wl.tlist)
        self.eng = eng = SweepEngine(cp, shp, lam)
        self.shard = None
        if world > 1:
            self.shard = ShardComm(dist, None, eng.device).attach(eng)
            eng.K_total = K
        self.stream = torch.cuda.current_stream()
        self.guess_t = eng.pulses_to_device(guess_pulses)
        self.opt_t = self.guess_t.clone()
        self.phiT = eng.propagate_forward(self.guess_t)
        self.tau_t = eng.overlaps(eng.t_targets, self.phiT)
        self.flush = flush_t
        self.fixed_chi = None
        if wl.chi == 'qubit_reset':
            self.fixed_chi = [cp.vec(wl.meta['chi_fixed'])
                              for _ in range(cp.K)]
            eng.chi_from_host(self.fixed_chi)
        # one launch per Krotov iteration (csrc/kq_picard.cuh) where the
        # problem allows, else chi boundary / backward sweep / fused sweep /
        # tau launches
        self.fused = engine != 'sweeps' and eng.fused_supported()
        self.spare_phiT = eng.new_states()
        self.spare_tau = torch.empty_like(self.tau_t)
        self.hint = False

    def one_iteration(self, ev=None):
        eng, wl, stream = self.eng, self.wl, self.stream
        if self.fused:
            if ev:
                ev[0].record(stream)
                ev[1].record(stream)
            try:
                eng.krotov_iteration(
                    None if self.fixed_chi is not None else wl.chi,
                    self.guess_t, self.opt_t, self.phiT, self.tau_t,
                    self.spare_phiT, self.spare_tau,
                    prev_guess_t=self.opt_t if self.hint else None)
                self.hint = True
            except self.krotov._lib.KqError as exc:
                if exc.status != -3:
                    raise
                self.fused = False
                return self.one_iteration(ev)
            if ev:
                ev[2].record(stream)
            self.phiT, self.spare_phiT = self.spare_phiT, self.phiT
            self.tau_t, self.spare_tau = self.spare_tau, self.tau_t
            self.guess_t, self.opt_t = self.opt_t, self.guess_t
            return
        if self.fixed_chi is None:
            eng.chi_builtin(wl.chi, self.phiT, self.tau_t,
                            K_total=self.K_total, shard=self.shard)
        if ev:
            ev[0].record(stream)
        eng.sweep_backward(self.guess_t)
        if ev:
            ev[1].record(stream)
        self.phiT = eng.sweep_forward_update(self.guess_t, self.opt_t,
                                             phiT=self.phiT)
        if ev:
            ev[2].record(stream)
        self.tau_t = eng.overlaps(eng.t_targets, self.phiT)
        self.guess_t, self.opt_t = self.opt_t, self.guess_t

    def barrier(self):
        if self.dist is not None:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def run(self, steps, warmup, sampler=None):
        """`warmup` untimed iterations, then `steps` iterations each timed
        with CUDA events on the launching stream, L2 flushed (untimed) before
        every one.  Returns a dict; ms_per_step is the max over ranks."""
        torch, eng = self.torch, self.eng
        for _ in range(warmup):
            self.one_iteration()
        self.barrier()
        if sampler is not None:
            sampler.start()
        launches0 = eng.launches
        t_iter, t_bw, t_fw = [], [], []
        self.barrier()
        wall0 = time.perf_counter()
        for _ in range(steps):
            if self.flush is not None:
                self.flush.zero_()   # evict L2 between timed iterations
            e = [torch.cuda.Event(enable_timing=True) for _ in range(5)]
            e[3].record(self.stream)
            self.one_iteration(e)
            e[4].record(self.stream)
            e[4].synchronize()
            t_iter.append(e[3].elapsed_time(e[4]))
            t_bw.append(e[0].elapsed_time(e[1]))
            t_fw.append(e[1].elapsed_time(e[2]))
        self.barrier()
        wall = time.perf_counter() - wall0
        launches = eng.launches - launches0
        clocks = sampler.stop() if sampler is not None else None
        if eng.status() != 0:
            raise RuntimeError("exchange failure in sweep kernel")
        fb_epoch, pic_iters = eng.sweep_diagnostics()
        if self.fused and eng.first_failed_epoch() != 0:
            raise RuntimeError("time-parallel iteration did not converge")
        total_ms = float(np.sum(t_iter))
        if self.dist is not None:
            t = torch.tensor([total_ms, float(launches)], dtype=torch.float64,
                             device=eng.device)
            self.dist.all_reduce(t[:1], op=self.dist.ReduceOp.MAX)
            self.dist.all_reduce(t[1:], op=self.dist.ReduceOp.SUM)
            total_ms, launches = float(t[0].item()), int(t[1].item())
        return dict(ms_per_step=total_ms / steps, launches=launches,
                    fw_ms=float(np.mean(t_fw)), bw_ms=float(np.mean(t_bw)),
                    rounds=pic_iters, fused=self.fused, wall=wall,
                    clocks=clocks,
                    sequential_fallback=bool(fb_epoch == eng.epoch))

    def pulses(self):
        """The current (optimised) pulses [L][NT] on the host."""
        return self.guess_t.cpu().numpy().copy()

    def close(self):
        if self.shard is not None:
            self.shard.close()
            self.shard = None


def e2e_run(krotov, wl, steps, warmup, parallel_map=None, dist=None,
            chi_constructor=None):
    """The metric through the public API: krotov_b200.optimize_pulses with
    host (numpy) objectives and a per-iteration host hook."""
    stamps = []

    def hook(**kw):
        stamps.append(time.perf_counter())
        return 1 - np.mean(kw['tau_vals']).real

    n_e2e = warmup + steps
    # one untimed call first: the first optimize_pulses of a process also pays
    # for CUDA's lazy loading of every kernel it launches (once per process,
    # not per call); reported separately as whole_call_first_seconds
    t0 = time.perf_counter()
    krotov.optimize_pulses(
        wl.objectives(krotov.Objective), wl.pulse_options, wl.tlist,
        propagator=krotov.propagators.expm,
        chi_constructor=chi_constructor or chi_of(krotov, wl),
        info_hook=lambda **kw: None, iter_stop=1, parallel_map=parallel_map)
    t_first = time.perf_counter() - t0
    t0 = time.perf_counter()
    res = krotov.optimize_pulses(
        wl.objectives(krotov.Objective), wl.pulse_options, wl.tlist,
        propagator=krotov.propagators.expm,
        chi_constructor=chi_constructor or chi_of(krotov, wl),
        info_hook=hook, iter_stop=n_e2e, parallel_map=parallel_map)
    t_call = time.perf_counter() - t0
    steady = (stamps[-1] - stamps[warmup]) / steps
    if dist is not None:
        import torch
        t = torch.tensor([steady], dtype=torch.float64, device='cuda')
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        steady = float(t.item())
    return {"value": 1.0 / steady, "unit": UNIT,
            "h2d_bytes_per_step": int(round(res.h2d_bytes_loop / n_e2e)),
            "d2h_bytes_per_step": int(round(res.d2h_bytes_loop / n_e2e)),
            "h2d_bytes_setup": int(res.h2d_bytes - res.h2d_bytes_loop),
            "whole_call_value": n_e2e / t_call,
            "whole_call_seconds": t_call,
            "whole_call_first_seconds": t_first,
            "fused_iterations": int(getattr(res, 'fused_iterations', 0)),
            "measured_total_h2d_bytes": res.h2d_bytes,
            "measured_total_d2h_bytes": res.d2h_bytes}, res


def notebook_entry(krotov, args, factory, kwargs, flush):
    """One of the reference's notebook-shaped problems on one B200: device
    timed iterations (one kq_krotov_iteration call each), the same through
    optimize_pulses with a host hook, and the numpy port on one core."""
    w = getattr(krotov.workloads, factory)(**kwargs)
    d = DeviceRun(krotov, w, engine=args.engine, flush_t=flush)
    steps = min(args.steps, 10)
    r = d.run(steps, 3)
    ee, _ = e2e_run(krotov, w, steps, 3)
    entry = {
        "workload": w.name, "K": w.K, "N": d.cp.N, "L": d.cp.L, "nt": w.nt,
        "value": 1e3 / r['ms_per_step'], "unit": UNIT,
        "ms_per_step": r['ms_per_step'],
        "engine": "fused" if r['fused'] else "sweeps",
        "gpu_launches": r['launches'],
        "e2e": {k: ee[k] for k in ("value", "unit", "h2d_bytes_per_step",
                                   "d2h_bytes_per_step", "whole_call_value")},
    }
    d.close()
    if not args.no_cpu:
        per_iter, ks = time_oracle(w, 1)
        entry["cpu_baseline"] = {
            "value": 1.0 / per_iter, "unit": UNIT, "cores": 1, "kind": "port",
            "host_cpus": os.cpu_count(),
            "sample": "1 Krotov iteration of the full workload (numpy port, "
                      "serial, one BLAS thread)"}
    return entry


def saturating_entry(krotov, args, peaks, which, K=131072):
    """One B200, C4's physics with K = 131 072 objectives (backward-state
    store 4.2 GB): whole Krotov iterations through the sweep calls, device
    timed; achieved = algorithmic bytes of both sweeps / iteration time."""
    import torch
    w = make_workload('C4', K=K, nt=WORKLOAD['nt'])
    d = DeviceRun(krotov, w, engine='sweeps')
    r = d.run(3, 3)
    cp = d.cp
    alg = 32.0 * cp.K * (cp.NT + 1) * cp.N
    it_ms = r['ms_per_step']
    entry = {
        "workload": "C4_tls_ensemble", "K": cp.K, "N": cp.N, "nt": w.nt,
        "value": 1e3 / it_ms, "unit": UNIT, "ms_per_step": it_ms,
        "bw_sweep_ms": r['bw_ms'], "fw_sweep_ms": r['fw_ms'],
        "gpu_launches": r['launches'],
        "roofline": {
            "bound": "hbm", "unit": "GB/s",
            "algorithmic_bytes_per_iteration": alg,
            "achieved": alg / (it_ms * 1e-3) / 1e9,
            "peak": peaks["hbm_gbs"], "peak_source": which + " (burst copy)",
            "frac": alg / (it_ms * 1e-3) / 1e9 / peaks["hbm_gbs"],
            "bw_sweep_frac": 0.5 * alg / (r['bw_ms'] * 1e-3) / 1e9
            / peaks["hbm_gbs"],
            "fw_sweep_frac": 0.5 * alg / (r['fw_ms'] * 1e-3) / 1e9
            / peaks["hbm_gbs"],
            "traffic": None,
            "limited_by": "update sweep: one grid-wide all-reduce per time "
                          "step (148 CTAs through L2) -- csrc/kq_sat.cuh; "
                          "backward sweep: HBM writes",
        },
    }
    traffic_file = os.path.join(ROOT, 'profiles', 'traffic.json')
    if os.path.exists(traffic_file):
        try:
            with open(traffic_file) as fh:
                t = json.load(fh)
            entry["roofline"]["traffic"] = (t.get("fw_sat_K131072", 0.0)
                                            + t.get("bw_sat_K131072", 0.0))
        except Exception:
            pass
    d.close()
    del d
    torch.cuda.empty_cache()
    return entry


def roofline_of(run, cp, peaks, which, fp64_peak):
    N, NT, K = cp.N, cp.NT, cp.K
    fused = run['fused']
    fw_ms, bw_ms = run['fw_ms'], run['bw_ms']
    if fused:
        dominant = ("time-parallel Krotov iteration kernel (chi boundary + "
                    "backward sweep + update/forward sweep + tau)")
        # the reference algorithm's state traffic of BOTH sweeps (backward
        # states written, then read): the kernel keeps them in shared memory
        alg_bytes = 32.0 * K * (NT + 1) * N
    else:
        dominant = "fused update+forward sweep kernel" \
            if fw_ms >= bw_ms else "backward sweep kernel"
        # X rows read (fw) or written (bw) by this rank's kernel
        alg_bytes = 16.0 * K * (NT + 1) * N
    dom_ms = max(fw_ms, bw_ms)
    achieved = alg_bytes / (dom_ms * 1e-3) / 1e9
    flops = algorithmic_flops(K, N, NT, cp.M, cp.L)
    if not fused:
        flops *= 0.5
    fp64 = {"algorithmic_gflop_per_launch": flops / 1e9,
            "achieved_tflops": flops / (dom_ms * 1e-3) / 1e12,
            "peak_tflops": None, "frac": None,
            "peak_source": "profiles/micro/fp64_peak.cu (measured DFMA "
                           "throughput of this device, same run)"}
    if fp64_peak and fp64_peak.get('tflops'):
        fp64["peak_tflops"] = fp64_peak['tflops']
        fp64["frac"] = fp64["achieved_tflops"] / fp64_peak['tflops']
    return {
        "bound": "hbm", "achieved": achieved, "peak": peaks["hbm_gbs"],
        "unit": "GB/s", "frac": achieved / peaks["hbm_gbs"],
        "traffic": None, "peak_source": which + " (burst copy)",
        "limited_by": ("latency: fixed-point rounds x cross-CTA exchange "
                       "through L2 (DRAM traffic ~0, FP64 pipe ~15 % busy); "
                       "the HBM figure is the formal bound of SURVEY 8(d), "
                       "not what limits this kernel") if fused else
                      ("dependency chain of nt-1 sequential steps"),
        "fp64": fp64,
        "kernel": dominant, "kernel_ms": dom_ms,
        "algorithmic_bytes_per_launch": alg_bytes,
        "fw_sweep_ms": fw_ms, "bw_sweep_ms": bw_ms,
        "picard_iterations_last_sweep": run['rounds'],
        "sequential_fallback_used": run['sequential_fallback'],
        "ns_per_time_step_fw": fw_ms * 1e6 / NT,
        "ns_per_time_step_bw": bw_ms * 1e6 / NT,
        "engine": "fused" if fused else "sweeps",
    }


def run_ours(args):
    import torch
    import krotov_b200 as krotov
    from krotov_b200.parallelization import GPUShards

    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        # NCCL announces its version on stdout when the first communicator is
        # created: send that to stderr so that stdout carries the JSON line only
        sys.stdout.flush()
        saved_fd = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group('nccl', device_id=torch.device(
                'cuda', local_rank))
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved_fd, 1)
            os.close(saved_fd)

    lib = krotov._lib.load()
    if args.picard is not None:
        krotov._lib.check(lib.kq_set_option(b"picard", args.picard))
    if args.dpoly is not None:
        krotov._lib.check(lib.kq_set_option(b"dpoly", args.dpoly))
    if args.picard_rtol_e15 is not None:
        krotov._lib.check(lib.kq_set_option(b"picard_rtol_e15",
                                            args.picard_rtol_e15))
    if args.picard_history is not None:
        krotov._lib.check(lib.kq_set_option(b"picard_history",
                                            args.picard_history))
    if args.pdl:
        for opt, val in ((b"cooperative_launch", 0),
                         (b"programmatic_launch", 1)):
            krotov._lib.check(lib.kq_set_option(opt, val))
    device = torch.device('cuda', local_rank)
    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32,
                        device=device)   # 256 MB > 126 MB L2
    label = WORKLOAD['workload'][:2]
    wl = build_workload()
    K = wl.K
    peaks, which = load_peaks()
    fp64_peak = measure_fp64_peak() if rank == 0 else None

    # ---- the contract measurement: ONE problem, sharded when world > 1 -----
    sharded = world > 1 and args.shard_mode in ('auto', 'sharded')
    main = DeviceRun(krotov, wl, dist=dist, sharded=sharded,
                     engine=args.engine, flush_t=flush)
    sampler = ClockSampler(local_rank) if rank == 0 else None
    run = main.run(args.steps, args.warmup, sampler)
    value = 1e3 / run['ms_per_step']
    roofline = roofline_of(run, main.cp, peaks, which, fp64_peak)
    traffic_file = os.path.join(ROOT, 'profiles', 'traffic.json')
    if os.path.exists(traffic_file):
        try:
            with open(traffic_file) as fh:
                roofline["traffic"] = json.load(fh).get(
                    "fused" if run['fused'] else (
                        "fw" if run['fw_ms'] >= run['bw_ms'] else "bw"))
        except Exception:
            pass
    sharded_pulses = main.pulses()
    main.close()

    extras = {}
    if world > 1:
        # -- parity of the sharded pulses against the same problem on one GPU
        single = DeviceRun(krotov, wl, engine=args.engine)
        for _ in range(args.warmup + args.steps):
            single.one_iteration()
        torch.cuda.synchronize()
        ref_p = single.pulses()
        dev = float(np.max(np.abs(sharded_pulses - ref_p)) /
                    np.max(np.abs(ref_p)))
        t = torch.tensor(sharded_pulses, device=device)
        t0 = t.clone()
        dist.broadcast(t0, 0)
        same = torch.tensor([1.0 if torch.equal(t, t0) else 0.0],
                            device=device)
        dist.all_reduce(same, op=dist.ReduceOp.MIN)
        extras["sharded_parity"] = {
            "max_rel_pulse_deviation_vs_one_gpu": dev, "tolerance": 1e-10,
            "ok": bool(dev <= 1e-10),
            "identical_on_all_ranks": bool(same.item() == 1.0),
            "after_iterations": args.warmup + args.steps}
        # one GPU, same problem, same run (every rank measures it; max)
        single1 = DeviceRun(krotov, wl, engine=args.engine, flush_t=flush)
        r1 = single1.run(args.steps, args.warmup)
        t = torch.tensor([r1['ms_per_step']], dtype=torch.float64,
                         device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        extras["one_gpu_same_run"] = {
            "value": 1e3 / float(t.item()), "unit": UNIT,
            "ms_per_step": float(t.item()),
            "note": "the complete 128-objective problem on each GPU alone "
                    "(no communication), slowest rank"}
        # -- N independent optimisations (the old 'replicas' number)
        rep = DeviceRun(krotov, make_workload('C4', replica=rank,
                                              K=WORKLOAD['K'],
                                              nt=WORKLOAD['nt']),
                        dist=dist, sharded=False, engine=args.engine,
                        flush_t=flush) if label == 'C4' else None
        if rep is not None:
            rr = rep.run(args.steps, args.warmup)
            extras["replicas"] = {
                "value": world * 1e3 / rr['ms_per_step'], "unit": UNIT,
                "note": "%d independent %d-objective optimisations, one per "
                        "GPU, no communication (not the contract metric)"
                        % (world, K)}
        # -- weak scaling of ONE ensemble: 512 objectives per GPU
        if label == 'C4':
            try:
                big = make_workload('C4', K=WEAK_K_PER_GPU * world,
                                    nt=WORKLOAD['nt'])
                w = DeviceRun(krotov, big, dist=dist, sharded=True,
                              engine=args.engine, flush_t=flush)
                wr = w.run(args.steps, args.warmup)
                w.close()
                one = DeviceRun(krotov, make_workload(
                    'C4', K=WEAK_K_PER_GPU, nt=WORKLOAD['nt']),
                    engine=args.engine, flush_t=flush)
                o1 = one.run(args.steps, args.warmup)
                t = torch.tensor([o1['ms_per_step']], dtype=torch.float64,
                                 device=device)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                ms1 = float(t.item())
                extras["weak_scaling_ensemble"] = {
                    "K": WEAK_K_PER_GPU * world, "nt": WORKLOAD['nt'],
                    "objectives_per_gpu": WEAK_K_PER_GPU,
                    "value": 1e3 / wr['ms_per_step'], "unit": UNIT,
                    "ms_per_step": wr['ms_per_step'],
                    "objective_iterations_per_sec":
                        WEAK_K_PER_GPU * world * 1e3 / wr['ms_per_step'],
                    "one_gpu_K%d_ms_per_step" % WEAK_K_PER_GPU: ms1,
                    "efficiency": ms1 / wr['ms_per_step'],
                    "rounds": wr['rounds'], "engine":
                        "fused" if wr['fused'] else "sweeps",
                    "note": "ONE ensemble of %d objectives sharded over %d "
                            "GPUs (one in-kernel NVLink exchange per "
                            "fixed-point round) against %d objectives on one "
                            "GPU; efficiency = t(1 GPU, K=%d) / t(N GPUs, "
                            "K=%d*N)" % (WEAK_K_PER_GPU * world, world,
                                         WEAK_K_PER_GPU, WEAK_K_PER_GPU,
                                         WEAK_K_PER_GPU)}
            except Exception as exc:  # pragma: no cover
                extras["weak_scaling_ensemble"] = {"unavailable": repr(exc)}

    # ---- end to end through the public API, host buffers -------------------
    selector = GPUShards(mode='sharded') if sharded else None
    e2e, res = e2e_run(krotov, wl, args.steps, args.warmup,
                       parallel_map=selector, dist=dist)
    e2e["api"] = "krotov_b200.optimize_pulses(numpy objectives, info_hook)"
    e2e["note"] = (
        "hooked iterations are pipelined: two iterations are launched ahead "
        "of the one whose hook runs and results are fetched on a copy "
        "stream, so iterations run back to back with a warm L2 -- e2e can "
        "exceed `value`, which flushes L2 before every (isolated, "
        "event-timed) iteration")

    cpu = cpu_serial = cpu_c = cpu_ref = parity = None
    configs = None
    e2e_host_chi = None
    if rank == 0 and world == 1:
        # ---- end to end with a HOST chi_constructor (the reference's plugin
        # boundary, optimize.py:404-406): phi(T) down, chi(T) up every
        # iteration, no launch-ahead
        try:
            builtin = chi_of(krotov, wl)

            def host_chi(fw_states_T, objectives, tau_vals):
                return builtin(fw_states_T=fw_states_T,
                               objectives=objectives, tau_vals=tau_vals)
            e2e_host_chi, _ = e2e_run(krotov, wl, args.steps, args.warmup,
                                      chi_constructor=host_chi)
            e2e_host_chi["api"] = ("optimize_pulses with a Python "
                                   "chi_constructor callback + info_hook")
        except Exception as exc:  # pragma: no cover
            e2e_host_chi = {"unavailable": repr(exc)}

    # ---- CPU baselines (rank 0, N=1 only) ----------------------------------
    if rank == 0 and world == 1 and not args.no_cpu:
        per_iter, ks = time_oracle(wl, 2, k_sample=48 if label == 'C4'
                                   else None)
        cpu = {"value": 1.0 / per_iter, "unit": UNIT, "cores": 1,
               "kind": "port", "host_cpus": os.cpu_count(),
               "sample": "2 Krotov iterations of the first %d of %d "
                         "objectives at nt=%d, time scaled by %d/%d"
                         % (ks, K, wl.nt, K, ks)}
        cpu_serial = cpu
        override = dict(K=WORKLOAD['K'], nt=WORKLOAD['nt']) \
            if label == 'C4' else {}
        cpu_ref = time_unmodified_reference(label, min(K, 16), 2, **override)
        if label == 'C4':
            try:   # reference's multi-process mode, all host cores
                import tempfile
                dump = os.path.join(tempfile.mkdtemp(), 'cpu_pulses.npy')
                par = parallel_leg(wl, 3, 1, in_subprocess=True, dump=dump)
                if par["value"] > cpu["value"]:
                    cpu = par
                # parity in the same run: the pulses after Krotov iterations
                # 1..4 of the full workload, CUDA path vs the CPU port above
                want = np.load(dump)[:, 0, :]
                got = krotov.optimize_pulses(
                    wl.objectives(krotov.Objective), wl.pulse_options,
                    wl.tlist, propagator=krotov.propagators.expm,
                    chi_constructor=chi_of(krotov, wl), iter_stop=len(want),
                    store_all_pulses=True).all_pulses[-len(want):]
                devs = [float(np.max(np.abs(np.array(g)[0] - w))
                              / np.max(np.abs(w))) for g, w in zip(got, want)]
                parity = {"max_rel_pulse_deviation": max(devs),
                          "per_iteration": devs, "tolerance": 1e-10,
                          "ok": bool(max(devs) <= 1e-10),
                          "against": "numpy port (cpu_baseline leg of this "
                                     "run), full workload, Krotov iterations "
                                     "1..%d from the same guess" % len(want)}
            except Exception as exc:  # pragma: no cover
                cpu_serial = dict(cpu_serial, parallel_leg_error=repr(exc))
            try:
                cpu_c = time_c_oracle(wl)
            except Exception as exc:  # pragma: no cover
                cpu_c = {"unavailable": repr(exc)}
        if cpu_ref and cpu_ref.get("value", 0.0) > cpu["value"]:
            cpu = cpu_ref

    # ---- the other BASELINE configs, same run (N = 1) ----------------------
    if rank == 0 and world == 1 and label == 'C4' and not args.no_configs:
        configs = {}
        for lab in ('C1', 'C2', 'C3', 'C5'):
            try:
                w2 = make_workload(lab)
                d = DeviceRun(krotov, w2, engine=args.engine, flush_t=flush)
                steps2 = min(args.steps, 10)
                r2 = d.run(steps2, 3)
                ee, _ = e2e_run(krotov, w2, steps2, 3)
                entry = {
                    "workload": w2.name, "K": w2.K, "N": d.cp.N, "nt": w2.nt,
                    "value": 1e3 / r2['ms_per_step'], "unit": UNIT,
                    "ms_per_step": r2['ms_per_step'],
                    "engine": "fused" if r2['fused'] else "sweeps",
                    "rounds": r2['rounds'], "gpu_launches": r2['launches'],
                    "e2e": {k: ee[k] for k in (
                        "value", "unit", "h2d_bytes_per_step",
                        "d2h_bytes_per_step", "whole_call_value")},
                }
                if not args.no_cpu:
                    port, ref = cpu_legs_for(lab, w2)
                    entry["cpu_baseline"] = port
                    entry["cpu_baseline_reference"] = ref
                configs[lab] = entry
            except Exception as exc:  # pragma: no cover
                configs[lab] = {"unavailable": repr(exc)}
        # -- the reference's notebook-shaped problems: several controls per
        # objective (notebooks 03 / 08) and the 17-level transmon (notebook 05)
        for lab, factory, kw in (
                ('nb03_lambda_4controls', 'lambda_system',
                 dict(nt=500, gamma=0.5)),
                ('nb08_lambda_ensemble', 'lambda_system',
                 dict(nt=500, gamma=0.0, lambda_a=0.5,
                      ensemble_mu=[0.9, 0.95, 1.0, 1.05, 1.1])),
                ('nb05_transmon_N17', 'transmon_xgate',
                 dict(nstates=8, nt=1000))):
            try:
                configs[lab] = notebook_entry(krotov, args, factory, kw, flush)
            except Exception as exc:  # pragma: no cover
                configs[lab] = {"unavailable": repr(exc)}
        # -- the same physics with K = 131 072 objectives: the
        # regime where the sweeps stream the 4.2 GB backward-state store from
        # HBM -- the number that can be read against the HBM roof
        try:
            configs['C4sat'] = saturating_entry(krotov, args, peaks, which)
        except Exception as exc:  # pragma: no cover
            configs['C4sat'] = {"unavailable": repr(exc)}

    if rank == 0:
        if world == 1:
            par = "1 GPU"
        elif sharded:
            par = ("%d GPUs, the objectives of ONE problem sharded over the "
                   "ranks (GPUShards mode 'sharded': %d per GPU, one "
                   "in-kernel NVLink exchange of nt doubles per fixed-point "
                   "round, no NCCL call on the data path)"
                   % (world, -(-K // world)))
        else:
            par = "%d GPUs, GPUShards mode '%s'" % (world, args.shard_mode)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT,
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": run['ms_per_step'], "higher_is_better": True,
            "scaling": "weak" if world == 1 else "strong",
            "vs_baseline": None, "dtype": "c128",
            "data": "synthetic",
            "config": dict(WORKLOAD, l2="flushed between timed iterations "
                           "(256 MB write)", parallelism=par),
            "e2e": e2e, "gpu_launches": run['launches'],
            "clocks": run['clocks'],
            "roofline": roofline, "cpu_baseline": cpu,
            "cpu_baseline_serial": cpu_serial, "cpu_baseline_c": cpu_c,
            "cpu_baseline_reference": cpu_ref,
            "parity": parity, "e2e_host_chi": e2e_host_chi,
            "fp64_peak": fp64_peak,
            "wall_seconds_timed_region": run['wall'],
        }
        if cpu_c and cpu_c.get("value"):
            line["ratio_vs_cpu_baseline_c"] = {
                "device": value / cpu_c["value"],
                "e2e": e2e["value"] / cpu_c["value"]}
        if configs is not None:
            line["configs"] = configs
        line.update(extras)
        print(json.dumps(line))
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--no-cpu', action='store_true',
                    help='skip the CPU baseline legs')
    ap.add_argument('--no-configs', action='store_true',
                    help='skip the C1/C2/C3/C5 measurements of the N=1 line')
    ap.add_argument('--shard-mode', default='auto',
                    choices=['auto', 'sharded'],
                    help='multi-GPU distribution of the one problem')
    ap.add_argument('--engine', default='auto', choices=['auto', 'sweeps'],
                    help="'sweeps' forces the four-launch sweep sequence")
    ap.add_argument('--picard', type=int, default=None, choices=[0, 1, 2],
                    help='time-parallel fused sweep: 0 off (sequential '
                         'kernel), 1 on (library default)')
    ap.add_argument('--dpoly', type=int, default=None, choices=[0, 1, 2],
                    help="delta-polynomial iteration: 0 never, 1 where the "
                    "engine asks for it (library default), 2 wherever it fits")
    ap.add_argument('--picard-rtol-e15', type=int, default=None,
                    help="fixed-point tolerance of the time-parallel kernel in "
                    "units of 1e-15 (library default 20 = 2e-14)")
    ap.add_argument('--picard-history', type=int, default=None,
                    choices=[0, 1], help='update-history first iterate of '
                    'the fixed-point kernel (library default: on)')
    ap.add_argument('--pdl', action='store_true',
                    help='experimental: regular launches with programmatic '
                         'stream serialization instead of cooperative ones')
    ap.add_argument('--workload', default='C4',
                    help='C4 (contract workload) or C1/C2/C3/C5/C4sat for '
                         'additional measurements')
    ap.add_argument('--cpu-parallel-leg', action='store_true',
                    help='internal: print the multi-process CPU leg and exit')
    ap.add_argument('--dump-pulses', default=None, help='internal')
    args = ap.parse_args()
    if args.cpu_parallel_leg:
        print(json.dumps(parallel_leg(build_workload(), args.steps,
                                      max(args.warmup, 1),
                                      dump=args.dump_pulses)))
        return
    if args.workload == 'C4sat':
        WORKLOAD.update(workload='C4_tls_ensemble', K=131072)
        args.no_cpu = True
        args.no_configs = True
    elif args.workload != 'C4':
        names = {'C1': 'C1_tls_state_to_state', 'C2': 'C2_transmon_xgate',
                 'C3': 'C3_two_qubit_gate', 'C5': 'C5_dissipative_qubit_reset'}
        WORKLOAD.update(workload=names[args.workload])
        args.no_configs = True
    if args.warmup < 3 and args.impl == 'ours':
        args.warmup = 3
    if args.impl == 'reference':
        run_reference_arm(args)
    else:
        run_ours(args)


if __name__ == '__main__':
    main()
