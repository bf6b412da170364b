#!/usr/bin/env python
"""Benchmark of the Krotov sweep hot path.

Metric (BASELINE.json): Krotov iterations/sec on the 128-objective, nt=1000
two-level ensemble (configs[3], "C4").  One *step* = one Krotov iteration =
chi boundary -> backward sweep -> fused update/forward sweep -> tau
(/root/reference/src/krotov/optimize.py:393-508).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

Our arm prints ONE JSON line with
  value      iterations/s with all inputs resident in HBM, timed with CUDA
             events on the launching stream (max over ranks),
  e2e        the same metric through the public krotov_b200.optimize_pulses
             call with HOST (numpy) objectives and a per-iteration host hook:
             every iteration copies its results device->host and the pulses
             host->device inside the timed region,
  roofline   algorithmic HBM bytes of the dominant kernel / its measured
             duration against MEASURED_PEAKS.json,
  cpu_baseline  the numpy oracle port of the reference loop timed on this
             box's host cores: the faster of the serial loop (bounded sample)
             and the reference's multi-process mode on all host cores.
  parity     max relative deviation of the updated pulses (Krotov iterations
             1..4, full workload) between the CUDA path and the CPU port
             timed in the same run (tolerance 1e-10).
The reference arm (--impl reference) times the oracle port (the reference is
pure Python + QuTiP and cannot travel to the GPU box): the
faster of its serial loop and its multi-process mode on all host cores.
With --gpus N > 1 the default is N independent replicas (one ensemble
optimisation per GPU, "replicas only", scaling "weak"); --shard-mode
exchange|gather|replicate select the distributions of one problem.
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


def build_workload(replica=0):
    """`replica` > 0 (multi-GPU 'independent' mode): the same ensemble with a
    different guess amplitude -- an independent optimisation per GPU."""
    import krotov_b200 as krotov
    name = WORKLOAD['workload']
    if name == 'C4_tls_ensemble':
        return krotov.workloads.tls_ensemble(
            K=WORKLOAD['K'], nt=WORKLOAD['nt'], ampl0=0.2 * (1 + 0.05 * replica))
    wl = krotov.workloads.by_name(name[:2])
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
# CPU side: the oracle port of the reference loop

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
    fw_T = [orc.forward_propagation(terms[k], pulses, low['tlist'], psi0[k],
                                    False, store_all=False)
            for k in range(ks)]
    tau = np.array([np.vdot(targets[k], fw_T[k]) for k in range(ks)])
    times = []
    for _ in range(iters):
        t0 = time.perf_counter()
        rec = orc.krotov_iteration(
            terms, psi0, targets, pulses, low['shapes'], low['lambdas'],
            low['tlist'], fw_T, tau, orc.chis_re, False)
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


def run_reference_arm(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    wl = build_workload()
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
    if WORKLOAD['workload'] == 'C4_tls_ensemble':
        try:   # the reference's multi-process mode on all host cores
            parallel = parallel_leg(wl, args.steps, max(args.warmup, 1))
            if parallel["value"] > value:
                best = parallel
                value, per_iter = parallel["value"], 1.0 / parallel["value"]
        except Exception as exc:  # pragma: no cover
            parallel = {"unavailable": repr(exc)}
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT,
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": per_iter * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "c128",
        "data": "synthetic", "config": dict(WORKLOAD),
        "cpu_baseline": best,
        "cpu_baseline_serial": serial,
        "cpu_baseline_parallel": parallel,
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0,
                "d2h_bytes_per_step": 0},
    }
    try:   # extra information: an optimised multi-threaded C port
        line["cpu_baseline_c"] = time_c_oracle(wl)
    except Exception as exc:  # pragma: no cover
        line["cpu_baseline_c"] = {"unavailable": repr(exc)}
    print(json.dumps(line))


# --------------------------------------------------------------------------
# GPU side

def run_ours(args):
    import torch
    import krotov_b200 as krotov
    from krotov_b200.compiler import compile_problem, initialize_controls
    from krotov_b200.engine import SweepEngine

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

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    from krotov_b200.parallelization import (GPUShards, ShardComm,
                                             shard_bounds)
    if args.picard is not None:
        krotov._lib.check(krotov._lib.load().kq_set_option(
            b"picard", args.picard))
    # Multi-GPU: where the library would only replicate the problem (the path
    # does not shard at this size) bench.py runs N
    # INDEPENDENT replicas, one ensemble optimisation per GPU, no collective
    # on the data path ("replicas only"); --shard-mode selects the sharded
    # (exchange / gather) or the redundant (replicate) distributions instead.
    if args.picard_history is not None:
        krotov._lib.check(krotov._lib.load().kq_set_option(
            b"picard_history", args.picard_history))
    if args.pdl:
        for opt, val in ((b"cooperative_launch", 0),
                         (b"programmatic_launch", 1)):
            krotov._lib.check(krotov._lib.load().kq_set_option(opt, val))
    wl = build_workload()
    K = len(wl.Hs)
    n_state = len(wl.lowered()['psi0'][0])
    independent = False
    if world > 1:
        if args.shard_mode in ('auto', 'independent'):
            independent = (args.shard_mode == 'independent' or (
                args.engine != 'sweeps' and
                GPUShards(mode='auto').choose(K, n_state) == 'replicate'))
        if independent:
            wl = build_workload(replica=rank)
    objectives = wl.objectives(krotov.Objective)
    (controls, _, guess_pulses, mapping, lam, shp) = initialize_controls(
        objectives, wl.pulse_options, wl.tlist)
    selector = GPUShards(mode='auto' if args.shard_mode == 'independent'
                         else args.shard_mode)
    mode = selector.choose(K, n_state) if world > 1 else None
    if independent:
        mode = 'replicate'   # every rank runs its own complete problem
    if args.engine == 'sweeps' and mode == 'replicate':
        mode = 'gather'
    lo, hi = shard_bounds(K, world, rank) if mode == 'exchange' else (0, K)
    cp = compile_problem(objectives[lo:hi], controls, mapping[lo:hi],
                         wl.tlist)
    eng = SweepEngine(cp, shp, lam)
    shard = gather_comm = None
    if mode == 'replicate' and not independent and \
            not eng.fused_supported():
        mode = 'gather'
    if mode == 'exchange':
        shard = ShardComm(dist, None, eng.device).attach(eng)
    elif mode == 'gather':
        gather_comm = ShardComm(dist, None, eng.device).attach_gather(eng)
    N, NT, L = cp.N, cp.NT, cp.L
    stream = torch.cuda.current_stream()

    # ---- device-resident steps ---------------------------------------------
    guess_t = eng.pulses_to_device(guess_pulses)
    opt_t = guess_t.clone()
    phiT = eng.propagate_forward(guess_t)
    tau_t = eng.overlaps(eng.t_targets, phiT)
    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32,
                        device=eng.device)   # 256 MB > 126 MB L2

    fixed_chi = None
    if wl.chi == 'qubit_reset':
        fixed_chi = [cp.vec(wl.meta['chi_fixed']) for _ in range(cp.K)]
        eng.chi_from_host(fixed_chi)

    # one launch per Krotov iteration (csrc/kq_picard.cuh) where the problem
    # allows, else chi boundary / backward sweep / fused sweep / tau launches
    fused = (args.engine != 'sweeps' and (world == 1 or mode == 'replicate')
             and eng.fused_supported())
    spare_phiT, spare_tau = eng.new_states(), torch.empty_like(tau_t)
    hint = [False]   # after the first iteration opt_t holds the previous guess

    def one_iteration(ev=None):
        nonlocal guess_t, opt_t, phiT, tau_t, spare_phiT, spare_tau, fused
        if fused:
            if ev:
                ev[0].record(stream)
                ev[1].record(stream)
            try:
                eng.krotov_iteration(
                    None if fixed_chi is not None else wl.chi, guess_t, opt_t,
                    phiT, tau_t, spare_phiT, spare_tau,
                    prev_guess_t=opt_t if hint[0] else None)
                hint[0] = True
            except krotov._lib.KqError:
                fused = False
                return one_iteration(ev)
            if ev:
                ev[2].record(stream)
            phiT, spare_phiT = spare_phiT, phiT
            tau_t, spare_tau = spare_tau, tau_t
            guess_t, opt_t = opt_t, guess_t
            return
        if fixed_chi is None:
            eng.chi_builtin(wl.chi, phiT, tau_t, K_total=K, shard=shard)
        if ev:
            ev[0].record(stream)
        eng.sweep_backward(guess_t)
        if ev:
            ev[1].record(stream)
        phiT = eng.sweep_forward_update(guess_t, opt_t, phiT=phiT)
        if ev:
            ev[2].record(stream)
        tau_t = eng.overlaps(eng.t_targets, phiT)
        guess_t, opt_t = opt_t, guess_t

    for _ in range(args.warmup):
        one_iteration()
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    launches0 = eng.launches
    t_iter, t_bw, t_fw = [], [], []
    barrier()
    wall0 = time.perf_counter()
    for _ in range(args.steps):
        flush.zero_()          # evict L2 between timed iterations (untimed)
        e = [torch.cuda.Event(enable_timing=True) for _ in range(5)]
        e[3].record(stream)
        one_iteration(e)
        e[4].record(stream)
        e[4].synchronize()
        t_iter.append(e[3].elapsed_time(e[4]))
        t_bw.append(e[0].elapsed_time(e[1]))
        t_fw.append(e[1].elapsed_time(e[2]))
    barrier()
    wall = time.perf_counter() - wall0
    launches = eng.launches - launches0
    clocks = sampler.stop() if rank == 0 else None
    if eng.status() != 0:
        raise RuntimeError("exchange failure in sweep kernel")
    fb_epoch, pic_iters = eng.sweep_diagnostics()
    if fused and eng.first_failed_epoch() != 0:
        raise RuntimeError("time-parallel iteration did not converge")
    total_ms = float(np.sum(t_iter))
    if dist is not None:
        t = torch.tensor([total_ms], dtype=torch.float64, device=eng.device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms = float(t.item())
    ms_per_step = total_ms / args.steps
    # independent replicas: every rank completed its own iteration per step
    units = world if independent else 1
    value = units * 1e3 / ms_per_step
    if dist is not None:
        t = torch.tensor([launches], dtype=torch.float64, device=eng.device)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        launches = int(t.item())

    # ---- end to end through the public API, host buffers ------------------
    stamps = []

    def hook(**kw):
        stamps.append(time.perf_counter())
        return 1 - np.mean(kw['tau_vals']).real

    n_e2e = args.warmup + args.steps
    t0 = time.perf_counter()
    res = krotov.optimize_pulses(
        wl.objectives(krotov.Objective), wl.pulse_options, wl.tlist,
        propagator=krotov.propagators.expm,
        chi_constructor=chi_of(krotov, wl), info_hook=hook,
        iter_stop=n_e2e,
        parallel_map=selector if (world > 1 and not independent) else None)
    t_call = time.perf_counter() - t0
    steady = (stamps[-1] - stamps[args.warmup]) / args.steps
    e2e_value = 1.0 / steady
    # measured inside the iteration loop of optimize_pulses: per iteration one
    # pinned device->host copy of pulses | g_a | tau | status words; the
    # pulses stay on the device between iterations (the next iteration's input
    # is this iteration's output) and are uploaded again only if a hook
    # modifies them, so the steady-state host->device traffic is zero -- the
    # upload of the host objectives/pulses is in whole_call_seconds
    d2h_step = int(round(res.d2h_bytes_loop / max(n_e2e, 1)))
    h2d_step = int(round(res.h2d_bytes_loop / max(n_e2e, 1)))
    e2e = {"value": e2e_value, "unit": UNIT,
           "h2d_bytes_per_step": h2d_step, "d2h_bytes_per_step": d2h_step,
           "h2d_bytes_setup": int(res.h2d_bytes - res.h2d_bytes_loop),
           "whole_call_value": n_e2e / t_call,
           "whole_call_seconds": t_call,
           "api": "krotov_b200.optimize_pulses(numpy objectives, info_hook)",
           "note": "hooked iterations are pipelined: two iterations are "
                   "launched ahead of the one whose hook runs and results "
                   "are fetched on a copy stream, so iterations run back to "
                   "back with a warm L2 -- e2e can exceed `value`, which "
                   "flushes L2 before every (isolated, event-timed) iteration",
           "measured_total_h2d_bytes": res.h2d_bytes,
           "measured_total_d2h_bytes": res.d2h_bytes}
    if dist is not None:
        t = torch.tensor([steady], dtype=torch.float64, device=eng.device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e["value"] = units / float(t.item())

    # ---- roofline of the dominant kernel -----------------------------------
    peaks, which = load_peaks()
    fw_ms, bw_ms = float(np.mean(t_fw)), float(np.mean(t_bw))
    if fused:
        dominant = ("time-parallel Krotov iteration kernel (chi boundary + "
                    "backward sweep + update/forward sweep + tau)")
        # the reference algorithm's state traffic of BOTH sweeps (backward
        # states written, then read): the kernel keeps them in shared memory
        alg_bytes = 32.0 * cp.K * (NT + 1) * N
    else:
        dominant = "fused update+forward sweep kernel" \
            if fw_ms >= bw_ms else "backward sweep kernel"
        # X rows read (fw) or written (bw) by this rank's kernel
        alg_bytes = 16.0 * cp.K * (NT + 1) * N
    dom_ms = max(fw_ms, bw_ms)
    achieved = alg_bytes / (dom_ms * 1e-3) / 1e9
    roofline = {
        "bound": "hbm", "achieved": achieved, "peak": peaks["hbm_gbs"],
        "unit": "GB/s", "frac": achieved / peaks["hbm_gbs"],
        "traffic": None, "peak_source": which + " (burst copy)",
        "kernel": dominant, "kernel_ms": dom_ms,
        "algorithmic_bytes_per_launch": alg_bytes,
        "fw_sweep_ms": fw_ms, "bw_sweep_ms": bw_ms,
        "picard_iterations_last_sweep": pic_iters,
        "sequential_fallback_used": bool(fb_epoch == eng.epoch),
        "ns_per_time_step_fw": fw_ms * 1e6 / NT,
        "ns_per_time_step_bw": bw_ms * 1e6 / NT,
        "engine": "fused" if fused else "sweeps",
        "note": ("one launch per iteration; ~%d fixed-point rounds, each "
                 "parallel in time; backward states stay in shared memory "
                 "(DRAM traffic ~0), bound by the cross-CTA exchange "
                 "latency, see DESIGN.md" % pic_iters) if fused else
                ("sequential chain of nt-1 dependent steps; working set "
                 "(%.1f MB) is L2-resident, see DESIGN.md" % (
                     2 * alg_bytes / 1e6)),
    }
    traffic_file = os.path.join(ROOT, 'profiles', 'traffic.json')
    if os.path.exists(traffic_file):
        try:
            with open(traffic_file) as fh:
                roofline["traffic"] = json.load(fh).get(
                    "fused" if fused else (
                        "fw" if fw_ms >= bw_ms else "bw"))
        except Exception:
            pass

    # ---- CPU baseline (rank 0, N=1 only) -----------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        per_iter, ks = time_oracle(wl, 2, k_sample=48)
        cpu = {"value": 1.0 / per_iter, "unit": UNIT, "cores": 1,
               "kind": "port", "host_cpus": os.cpu_count(),
               "sample": "2 Krotov iterations of the first %d of %d "
                         "objectives at nt=%d, time scaled by %d/%d"
                         % (ks, K, NT + 1, K, ks)}

    cpu_serial = cpu
    parity = None
    if cpu is not None and WORKLOAD['workload'] == 'C4_tls_ensemble':
        try:   # reference's multi-process mode, all host cores
            import tempfile
            dump = os.path.join(tempfile.mkdtemp(), 'cpu_pulses.npy')
            par = parallel_leg(wl, 3, 1, in_subprocess=True, dump=dump)
            if par["value"] > cpu["value"]:
                cpu = par
            # parity in the same run: the pulses after Krotov iterations
            # 1..4 of the full workload, CUDA path vs the CPU port timed above
            want = np.load(dump)[:, 0, :]
            got = krotov.optimize_pulses(
                wl.objectives(krotov.Objective), wl.pulse_options, wl.tlist,
                propagator=krotov.propagators.expm,
                chi_constructor=chi_of(krotov, wl), iter_stop=len(want),
                store_all_pulses=True).all_pulses[-len(want):]
            devs = [float(np.max(np.abs(np.array(g)[0] - w))
                          / np.max(np.abs(w))) for g, w in zip(got, want)]
            parity = {"max_rel_pulse_deviation": max(devs),
                      "per_iteration": devs, "tolerance": 1e-10,
                      "ok": bool(max(devs) <= 1e-10),
                      "against": "numpy port (cpu_baseline leg of this run), "
                                 "full workload, Krotov iterations 1..%d "
                                 "from the same guess" % len(want)}
        except Exception as exc:  # pragma: no cover
            cpu_serial = dict(cpu_serial, parallel_leg_error=repr(exc))

    cpu_c = None
    if cpu is not None and WORKLOAD['workload'] == 'C4_tls_ensemble':
        try:
            cpu_c = time_c_oracle(wl)
        except Exception as exc:  # pragma: no cover
            cpu_c = {"unavailable": repr(exc)}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT,
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "weak" if (independent or world == 1) else "strong",
            "vs_baseline": None, "dtype": "c128",
            "data": "synthetic",
            "config": dict(WORKLOAD, l2="flushed between timed iterations "
                           "(256 MB write)", parallelism=(
                               "1 GPU" if world == 1 else
                               "%d independent replicas, one %d-objective "
                               "ensemble optimisation per GPU (replicas only: "
                               "the path does not shard at this size), value "
                               "= iterations of all replicas / max time over "
                               "ranks" % (world, K) if independent else
                               "%d GPUs, GPUShards mode '%s'" % (world, mode))),
            "e2e": e2e, "gpu_launches": launches, "clocks": clocks,
            "roofline": roofline, "cpu_baseline": cpu,
            "cpu_baseline_serial": cpu_serial, "cpu_baseline_c": cpu_c,
            "parity": parity,
            "wall_seconds_timed_region": wall,
        }
        print(json.dumps(line))
    if shard is not None:
        shard.close()
    if gather_comm is not None:
        gather_comm.close()
    if dist is not None:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--no-cpu', action='store_true',
                    help='skip the CPU baseline leg')
    ap.add_argument('--shard-mode', default='auto',
                    choices=['auto', 'independent', 'exchange', 'gather',
                             'replicate'],
                    help='multi-GPU distribution of the fused sweep')
    ap.add_argument('--engine', default='auto', choices=['auto', 'sweeps'],
                    help="'sweeps' forces the four-launch sweep sequence")
    ap.add_argument('--picard', type=int, default=None, choices=[0, 1, 2],
                    help='time-parallel fused sweep: 0 off (sequential '
                         'kernel), 1 on (library default)')
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
    elif args.workload != 'C4':
        names = {'C1': 'C1_tls_state_to_state', 'C2': 'C2_transmon_xgate',
                 'C3': 'C3_two_qubit_gate', 'C5': 'C5_dissipative_qubit_reset'}
        WORKLOAD.update(workload=names[args.workload])
        args.no_cpu = True
    if args.warmup < 3 and args.impl == 'ours':
        args.warmup = 3
    if args.impl == 'reference':
        run_reference_arm(args)
    else:
        run_ours(args)


if __name__ == '__main__':
    main()
