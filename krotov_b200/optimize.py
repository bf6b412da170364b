"""``optimize_pulses``: Krotov's method with the time loops lowered to B200.

Same keyword plugin surface, bookkeeping and :class:`Result` semantics as
``krotov.optimize_pulses`` (/root/reference/src/krotov/optimize.py:33-590).
What differs is *where* the work happens: the three per-time-step Python
loops of the reference (initial forward propagation :806-846, backward
propagation :849-886, sequential update + forward step :449-500) are one
CUDA kernel launch each (``libkrotov_b200.so``), and the built-in
chi-constructors and tau overlaps run on the device too, so that an iteration
without hooks needs no host round trip at all.  Host callbacks
(`chi_constructor`, `info_hook`, `modify_params_after_iter`,
`check_convergence`, ``sigma.refresh``) keep their reference signatures and
run once per iteration.
"""
import collections
import copy
import logging
import time

import numpy as np

from . import functionals as _functionals
from .compiler import compile_problem, initialize_controls
from .conversions import control_onto_interval, pulse_onto_tlist
from ._lib import KqError
from ._lib import check as _check
from .engine import SweepEngine
from .info_hooks import chain
from .mu import derivative_wrt_pulse
from .parallelization import GPUShards, ShardComm, shard_bounds
from .propagators import DensityMatrixODEPropagator
from .propagators import expm as _expm_marker
from .result import Result
from .second_order import _overlap

__all__ = ['optimize_pulses']

_BUILTIN_CHI = {
    _functionals.chis_re: 're',
    _functionals.chis_ss: 'ss',
    _functionals.chis_sm: 'sm',
    _functionals.chis_hs: 'hs',
}


class _PackedResults:
    """Pulses | g_a | tau | status words of one iteration in ONE device buffer
    (three of them, iteration j uses buffer j % 3: up to two iterations are
    launched ahead of the one whose hooks run), so that an iteration with host
    hooks needs a single device->host copy into pinned memory and one
    synchronisation."""

    def __init__(self, torch, device, L, NT, K):
        # K: complex numbers in the tau region (sharded objectives: the
        # gathered phi(T) | tau blocks of all ranks)
        def up16(x):
            return (x + 15) // 16 * 16
        self.torch, self.L, self.NT, self.K = torch, L, NT, K
        self.o_ga = up16(L * NT * 8)
        self.o_tau = self.o_ga + up16(max(L, 1) * 8)
        self.o_diag = self.o_tau + K * 16
        self.nbytes = self.o_diag + 16
        self.dev = [torch.zeros(self.nbytes, dtype=torch.uint8, device=device)
                    for _ in (0, 1, 2)]
        self.host = torch.zeros(self.nbytes, dtype=torch.uint8).pin_memory()
        self.hnp = self.host.numpy()
        self.views = [self._views(r) for r in self.dev]
        # results are copied on a stream of their own, behind an event that
        # follows the iteration's launch: iterations already queued behind it
        # on the compute stream do not delay the copy
        self.copy_stream = torch.cuda.Stream(device=device)
        self.events = [torch.cuda.Event() for _ in self.dev]
        self.marked = [False] * len(self.dev)
        self._host_ptr = self.host.data_ptr()
        self._dev_ptr = [r.data_ptr() for r in self.dev]

    def mark(self, i, eng):
        """Record 'buffer i is complete' behind the launch just made."""
        self.events[i].record(self.torch.cuda.current_stream(eng.device))
        self.marked[i] = True

    def _views(self, r):
        t, L, NT, K = self.torch, self.L, self.NT, self.K
        return dict(
            pulses=r[:L * NT * 8].view(t.float64).view(L, NT),
            g_a=r[self.o_ga:self.o_ga + max(L, 1) * 8].view(t.float64),
            tau=r[self.o_tau:self.o_tau + K * 16].view(t.complex128),
            diag=r[self.o_diag:self.o_diag + 16])

    def fetch(self, i, eng, diag_written=False):
        """Copy buffer `i` to the host (after appending the engine's status
        words unless the kernel wrote them itself); returns (pulses [L][NT],
        g_a [L], tau [K], status words)."""
        if not diag_written:
            self.views[i]['diag'].copy_(eng.workspace[:16])
            self.marked[i] = False
        if not self.marked[i]:
            self.events[i].record(self.torch.cuda.current_stream(eng.device))
        self.marked[i] = False
        self.copy_stream.wait_event(self.events[i])
        _check(eng.lib.kq_fetch_results(
            self._host_ptr, self._dev_ptr[i], self.nbytes,
            self.copy_stream.cuda_stream))
        self.copy_stream.synchronize()
        eng.d2h_bytes += self.nbytes
        h, L, NT, K = self.hnp, self.L, self.NT, self.K
        pulses = h[:L * NT * 8].view(np.float64).reshape(L, NT)
        g_a = h[self.o_ga:self.o_ga + max(L, 1) * 8].view(np.float64)
        tau = h[self.o_tau:self.o_tau + K * 16].view(np.complex128)
        diag = h[self.o_diag:self.o_diag + 16].view(np.int32)
        return pulses, g_a, tau, diag


class _LazyStates:
    """Per-objective view ``store[k][n]`` of a device state store
    ``[nt, K, N]``; the tensor is downloaded once, on first access.  Valid
    during the iteration in which it is handed to a hook (the device buffers
    are re-used by the next iteration)."""

    def __init__(self, tensor, cp, eng=None):
        self._tensor, self._cp, self._host = tensor, cp, None
        self._eng = eng
        self._expired = False

    def _data(self):
        if self._host is None:
            if self._expired:
                raise RuntimeError(
                    "state storage is only valid during the iteration it "
                    "was passed to a hook; copy what you need inside the hook")
            self._host = (self._eng.download(self._tensor) if self._eng
                          else self._tensor.cpu().numpy())
        return self._host

    def _expire(self):
        self._expired = True
        self._tensor = None

    def __len__(self):
        return self._cp.K

    def __getitem__(self, k):
        return _LazyObjectiveStates(self, k)

    def __iter__(self):
        return (self[k] for k in range(len(self)))


class _LazyObjectiveStates:
    def __init__(self, parent, k):
        self._p, self._k = parent, k

    def __len__(self):
        return self._p._cp.NT + 1

    def __getitem__(self, n):
        cp = self._p._cp
        data = self._p._data()
        if isinstance(n, slice):
            return [self[i] for i in range(*n.indices(len(self)))]
        return cp.unvec(data[n, self._k].copy(), cp.state_templates[self._k])

    def __iter__(self):
        return (self[n] for n in range(len(self)))


class _LazyFinalStates:
    """List-like ``fw_states_T``: the K final states phi_k(T), downloaded from
    the device (and, when sharded, gathered over the ranks) only when a hook,
    a custom chi_constructor or the caller touches them.  The snapshot is
    taken on first access or, at the latest, by :meth:`freeze` right before
    the device buffer is overwritten by the next iteration."""

    def __init__(self, fetch, K):
        self._fetch, self._K, self._items = fetch, K, None

    def freeze(self, needed):
        """Drop the device reference; materialise first if `needed`."""
        if self._items is None and self._fetch is not None and needed:
            self._items = self._fetch()
        self._fetch = None

    def _list(self):
        if self._items is None:
            if self._fetch is None:
                raise RuntimeError(
                    "fw_states_T of a past iteration were not kept; copy "
                    "them inside the hook if they are needed later")
            self._items = self._fetch()
        return self._items

    def __len__(self):
        return self._K

    def __getitem__(self, i):
        return self._list()[i]

    def __iter__(self):
        return iter(self._list())

    def __eq__(self, other):
        return list(self) == list(other)

    def __reduce__(self):
        return (list, (list(self),))


def _check_lowerable(propagator, objectives):
    props = propagator if isinstance(propagator, list) else [propagator]
    if isinstance(propagator, list):
        assert len(props) == len(objectives)
    for p in props:
        if p is _expm_marker or isinstance(p, DensityMatrixODEPropagator):
            continue
        raise NotImplementedError(
            "krotov_b200 lowers the time loop to CUDA kernels and therefore "
            "only accepts krotov_b200.propagators.expm or a "
            "DensityMatrixODEPropagator instance as `propagator`; a custom "
            "per-step Python propagator (%r) cannot be lowered and there is "
            "no CPU fallback" % (p,))


def _check_overlap(overlap, cp):
    """A custom `overlap` is accepted only if it is the standard one
    (verified on random states); it is then lowered like the default."""
    if overlap is None or overlap is _overlap:
        return
    rng = np.random.default_rng(0)
    for _ in range(2):
        a = rng.normal(size=cp.N) + 1j * rng.normal(size=cp.N)
        b = rng.normal(size=cp.N) + 1j * rng.normal(size=cp.N)
        tmpl = cp.state_templates[0]
        got = overlap(cp.unvec(a, tmpl), cp.unvec(b, tmpl))
        if got is None or abs(complex(got) - np.vdot(a, b)) > 1e-10 * abs(
                np.vdot(a, b)):
            raise NotImplementedError(
                "custom `overlap` differs from <a|b> / tr(a^dag b) and "
                "cannot be lowered to the B200 sweep kernels")


def _restore_from_previous_result(result, objectives, tlist, store_all_pulses):
    """Guess controls/pulses from a previous Result, with the reference's
    validation and messages (optimize.py:707-774)."""
    if not isinstance(result, Result):
        raise ValueError("Continuation is only possible from a Result object")
    if len(objectives) != len(result.objectives):
        raise ValueError(
            "When continuing from a previous Result, the number of "
            "objectives must be the same")
    for a, b in zip(objectives, result.objectives):
        if a != b:
            raise ValueError(
                "When continuing from a previous Result, the objectives must "
                "remain unchanged")
    if store_all_pulses and len(result.all_pulses) == 0:
        raise ValueError(
            "The store_all_pulses parameter cannot be changed when "
            "continuing from a previous Result. Pass it as False.")
    if not store_all_pulses and len(result.all_pulses) > 0:
        raise ValueError(
            "The store_all_pulses parameter cannot be changed when "
            "continuing from a previous Result. Pass it as True.")
    same_grid = len(tlist) == len(result.tlist) and np.max(
        np.abs(np.array(tlist) - np.array(result.tlist))) <= 1e-5
    if not same_grid:
        raise ValueError(
            "When continuing from a previous Result, the controls must be "
            "defined on the same time grid")
    nt = len(tlist)
    guess_controls = []
    for control in result.optimized_controls:
        if len(control) == nt - 1:   # dumped before finalisation: pulses
            guess_controls.append(pulse_onto_tlist(control))
        elif len(control) == nt:
            guess_controls.append(control)
        else:
            raise ValueError(
                "Invalid Result: optimized_controls and tlist are incongruent")
    return guess_controls, [control_onto_interval(c) for c in guess_controls]


def _check_finite(pulses, iteration):
    """The kernels never hang on non-finite input, but their output is then
    meaningless: say so instead of returning NaN pulses (e.g. a
    chi_constructor that returned NaN, or a lambda_a of zero)."""
    if not all(np.all(np.isfinite(p)) for p in pulses):
        raise FloatingPointError(
            "non-finite pulse values after Krotov iteration %d (check "
            "lambda_a, the update shapes and the chi_constructor)" % iteration)


def optimize_pulses(objectives, pulse_options, tlist, *, propagator,
                    chi_constructor, mu=None, sigma=None, iter_start=0,
                    iter_stop=5000, check_convergence=None, info_hook=None,
                    modify_params_after_iter=None, storage='array',
                    parallel_map=None, store_all_pulses=False,
                    continue_from=None,
                    skip_initial_forward_propagation=False, norm=None,
                    overlap=None, limit_thread_pool=None, device=None,
                    engine_mode=None):
    """Use Krotov's method to optimize towards the given `objectives`.

    Arguments have the meaning documented for the reference
    (optimize.py:56-228), with these engine-specific notes:

    * `propagator`: :func:`krotov_b200.propagators.expm` or a
      :class:`~krotov_b200.propagators.DensityMatrixODEPropagator` (or a list
      of those).  Both select exact piecewise-constant propagation on the
      device.  Any other callable raises NotImplementedError.
    * `chi_constructor`: ``krotov_b200.functionals.chis_re/ss/sm/hs`` run on
      the device; any other callable is a host callback per iteration.
    * `mu`: None (default derivative) or a custom callable that is linear in
      the state and independent of time/pulse values (evaluated once).
    * `norm`: ignored -- the engine normalises chi with the L2/Frobenius
      norm; the update is invariant under this choice (optimize.py:410,467).
    * `overlap`: None or a callable equal to the default overlap.
    * `parallel_map`: a :class:`krotov_b200.parallelization.GPUShards`
      instance shards the objectives over the GPUs of a ``torchrun`` job
      (every rank calls ``optimize_pulses`` with the full `objectives` list
      and obtains identical pulses; hooks see gathered ``tau_vals`` and
      ``fw_states_T`` but only the local part of the state stores).  The
      reference's process-pool maps are accepted and ignored.
    * `storage`, `limit_thread_pool`: accepted and ignored (state stores live
      in HBM; objectives are batched on the GPU).
    * `device`: CUDA device (default: current torch device).
    * `engine_mode`: None (default) picks the fastest kernels: the fused
      time-parallel iteration kernel (``kq_krotov_iteration``, one launch per
      Krotov iteration) where the problem allows, else the sweep kernels;
      ``'sweeps'`` forces the four-launch sweep sequence.

    Returns:
        Result

    Raises:
        ValueError: as the reference, for invalid controls, shapes,
            `pulse_options` or `continue_from`.
        NotImplementedError: for plugins that cannot be lowered.
        krotov_b200.EngineUnavailable: no CUDA device / library not built.
    """
    logger = logging.getLogger('krotov')
    logger.info("Initializing optimization with Krotov's method")
    second_order = sigma is not None
    if modify_params_after_iter is not None:
        info_hook = (modify_params_after_iter if info_hook is None
                     else chain(modify_params_after_iter, info_hook))
    _check_lowerable(propagator, objectives)

    (controls, guess_controls, guess_pulses, pulses_mapping, lambda_vals,
     shape_arrays) = initialize_controls(objectives, pulse_options, tlist)
    if continue_from is not None:
        guess_controls, guess_pulses = _restore_from_previous_result(
            continue_from, objectives, tlist, store_all_pulses)
    if skip_initial_forward_propagation and second_order:
        raise ValueError(
            "skip_initial_forward_propagation is incompatible with "
            "second order Krotov (sigma is not None)")

    # ---- sharding over GPUs (parallel_map=GPUShards()) ---------------------
    K_total = len(objectives)
    shard = None
    shard_mode = None
    lo, hi = 0, K_total
    if isinstance(parallel_map, GPUShards):
        if second_order:
            raise NotImplementedError(
                "second-order Krotov is not available with GPUShards yet")
        dist, group, rank, world = parallel_map.resolve()
        if world > 1:
            if K_total < world:
                raise ValueError("fewer objectives than GPUs")
            from ._dense import dense as _dense
            n_state = _dense(objectives[0].initial_state).size
            shard_mode = parallel_map.choose(K_total, n_state, world)
            if shard_mode in ('exchange', 'sharded'):
                lo, hi = shard_bounds(K_total, world, rank)
    local_objectives = objectives[lo:hi]
    cp = compile_problem(local_objectives, controls, pulses_mapping[lo:hi],
                         tlist, mu=None if mu is derivative_wrt_pulse else mu,
                         pulses_for_mu=guess_pulses)
    _check_overlap(overlap, cp)
    eng = SweepEngine(cp, shape_arrays, lambda_vals, device=device)
    torch = eng.torch
    gather_comm = None
    if (hi - lo) != K_total:
        shard = ShardComm(dist, group, eng.device).attach(eng)
        eng.K_total = K_total
    if shard_mode == 'replicate' and not (
            engine_mode is None and eng.fused_supported()):
        shard_mode = 'gather'     # outside the one-launch kernel family
    if shard_mode == 'gather':
        # every rank holds the complete problem; only the backward sweep is
        # sharded (engine.sweep_backward)
        gather_comm = ShardComm(dist, group, eng.device).attach_gather(eng)
    chi_kind = _BUILTIN_CHI.get(chi_constructor)
    if chi_kind is not None and cp.targets is None:
        chi_kind = None  # built-ins need state targets; let the host raise
    if engine_mode not in (None, 'sweeps'):
        raise ValueError("engine_mode must be None or 'sweeps'")
    # one-launch-per-iteration kernel family (csrc/kq_picard.cuh); falls back
    # to the sweep kernels when the library declines or does not converge
    # ('sharded': the kernel itself sums over the GPUs in every round)
    use_fused = (engine_mode is None and gather_comm is None
                 and (shard is None or shard_mode == 'sharded')
                 and eng.fused_supported())
    L, NT, K = cp.L, cp.NT, K_total
    has_targets = cp.targets is not None
    templates = [obj.initial_state for obj in objectives]

    def states_to_host(t):
        if shard is not None:
            t = shard.all_gather_rows(t, K_total)
        arr = eng.download(t)
        return [cp.unvec(arr[k].copy(), templates[k]) for k in range(K)]

    def tau_to_host(tau_t):
        if tau_t is None:
            return np.array([None] * K)
        if shard is not None:
            tau_t = shard.all_gather_rows(tau_t, K_total)
        return eng.download(tau_t).copy()

    def pulses_to_host(p_t):
        arr = eng.download(p_t)
        return [arr[l].copy() for l in range(L)]

    def lazy_states(t):
        if shard is not None:
            # the gather is a collective: do it now, download lazily
            t = shard.all_gather_rows(t, K_total)

            def fetch(tt=t):
                arr = eng.download(tt)
                return [cp.unvec(arr[k].copy(), templates[k])
                        for k in range(K)]
            return _LazyFinalStates(fetch, K)
        return _LazyFinalStates(lambda: states_to_host(t), K)

    g_a_integrals = np.zeros(L)
    if continue_from is None:
        result = Result()
        result.start_local_time = time.localtime()
    else:
        result = copy.deepcopy(continue_from)

    # ---- initial forward propagation (optimize.py:295-322) ----------------
    guess_t = eng.pulses_to_device(guess_pulses)
    opt_t = guess_t.clone()
    Phi0 = Phi1 = None
    tic = time.time()
    if skip_initial_forward_propagation:
        if continue_from is not None:
            fw_states_T = list(continue_from.states)
            phiT = eng.upload(
                np.array([cp.vec(s) for s in fw_states_T[lo:hi]]),
                torch.complex128)
        else:
            logger.warning(
                "You should not use `skip_initial_forward_propagation` "
                "unless you are also passing `continue_from`")
            fw_states_T = [None] * K
            phiT = None
    else:
        if second_order:
            Phi0 = eng.new_state_store()
            Phi1 = eng.new_state_store()
        phiT = eng.propagate_forward(guess_t, store=Phi0)
        fw_states_T = None  # materialised on demand
    tau_t = None
    if has_targets and phiT is not None:
        tau_t = eng.overlaps(eng.t_targets, phiT)
    torch.cuda.synchronize(eng.device)
    toc = time.time()

    host_loop = (info_hook is not None or check_convergence is not None
                 or chi_kind is None or second_order)
    tau_vals = tau_to_host(tau_t)
    if fw_states_T is None:
        fw_states_T = lazy_states(phiT)

    forward_states = forward_states0 = None
    if second_order:
        forward_states0 = forward_states = _LazyStates(Phi0, cp, eng)

    info = None
    optimized_pulses = copy.deepcopy(guess_pulses)
    adjoint_objectives = ([obj.adjoint() for obj in objectives]
                          if info_hook is not None else None)
    static = dict(
        objectives=objectives, adjoint_objectives=adjoint_objectives,
        lambda_vals=lambda_vals, shape_arrays=shape_arrays, tlist=tlist,
        propagator=propagator, chi_constructor=chi_constructor,
        mu=derivative_wrt_pulse if mu is None else mu, sigma=sigma,
        iter_start=iter_start, iter_stop=iter_stop,
    )
    if info_hook is not None:
        info = info_hook(
            backward_states=None, forward_states=forward_states,
            forward_states0=forward_states0, guess_pulses=guess_pulses,
            optimized_pulses=optimized_pulses, g_a_integrals=g_a_integrals,
            fw_states_T=fw_states_T, tau_vals=tau_vals, start_time=tic,
            stop_time=toc, iteration=0, info_vals=[], shared_data={},
            **static)

    result.tlist = tlist
    result.objectives = objectives
    result.guess_controls = guess_controls
    result.optimized_controls = optimized_pulses
    result.controls_mapping = pulses_mapping
    if continue_from is None:
        if info is not None:
            result.info_vals.append(info)
        result.iters.append(0)
        result.iter_seconds.append(int(toc - tic))
        result.iter_seconds_device.append(toc - tic)
        if not np.all(tau_vals == None):  # noqa: E711
            result.tau_vals.append(tau_vals)
        if store_all_pulses:
            result.all_pulses.append(guess_pulses)
    else:
        iter_start = continue_from.iters[-1]
        logger.info("Continuing from previous result, with iteration %d",
                    iter_start + 1)
    result.states = fw_states_T

    # hooks may already have modified lambda_vals / guess_pulses at iteration 0
    # (tests/test_infohooks.py:30-37 halves lambda_a in every call)
    lam_snapshot = np.array(lambda_vals, dtype=np.float64)
    first_iteration = iter_start + 1
    if info_hook is not None:
        eng.set_lambda(lambda_vals)
        new_guess = eng.pulses_to_device(guess_pulses)
        if not torch.equal(new_guess, guess_t):
            guess_t.copy_(new_guess)
    deferred = []   # (iteration, tau_t, pulses_t or None, ev0, ev1) fast path
    finished_by_break = False
    # phi(T) and tau are double-buffered: the iteration reads the previous
    # ones (chis_ss/sm/hs) while it writes the new ones
    spare = {'phiT': eng.new_states(),
             'tau': torch.empty(cp.K, dtype=torch.complex128,
                                device=eng.device)}
    any_fused = False
    n_fused = 0             # iterations done by the one-launch kernel
    prev_guess_ref = None   # device pulses the previous iteration started from
    # hooked iterations: everything the host needs in one pinned copy
    packed = None
    # objectives sharded over GPUs, one-launch kernel: every rank's
    # phi(T) | tau block is all-gathered (ONE NCCL call per iteration, on the
    # compute stream) straight into the packed buffer, so hooks on every rank
    # see all objectives after the single device->host copy
    sharded_packed = host_loop and shard is not None and use_fused
    Kmax = -(-K_total // shard.world) if shard is not None else cp.K
    if host_loop and shard is None:
        packed = _PackedResults(torch, eng.device, L, NT, cp.K)
    elif sharded_packed:
        packed = _PackedResults(torch, eng.device, L, NT,
                                shard.world * Kmax * (cp.N + 1))
    loc_ring = [None, None, None]

    def ring_loc(m):
        """Rank-local phi(T) [Kmax, N] | tau [Kmax] of iteration m in one
        buffer (the unit of the all-gather)."""
        if loc_ring[m % 3] is None:
            flat = torch.zeros(Kmax * (cp.N + 1), dtype=torch.complex128,
                               device=eng.device)
            loc_ring[m % 3] = dict(
                flat=flat,
                phiT=flat[:Kmax * cp.N].view(Kmax, cp.N)[:cp.K],
                tau=flat[Kmax * cp.N:Kmax * cp.N + cp.K])
        return loc_ring[m % 3]

    def gather_into_packed(m):
        shard.all_gather_flat(packed.views[m % 3]['tau'],
                              ring_loc(m)['flat'])

    def split_gathered(block):
        """Host copy of the gathered region -> (tau [K], phi(T) [K, N])."""
        G = np.array(block).reshape(shard.world, Kmax * (cp.N + 1))
        taus, phis = [], []
        for r in range(shard.world):
            a, b = shard_bounds(K_total, shard.world, r)
            phis.append(G[r, :Kmax * cp.N].reshape(Kmax, cp.N)[:b - a])
            taus.append(G[r, Kmax * cp.N:Kmax * cp.N + (b - a)])
        return np.concatenate(taus), np.concatenate(phis, axis=0)
    ri = 0
    # Hooked iterations rotate through three sets of output buffers (iteration
    # j: packed buffer, phi(T) and backward-state store number j % 3), so that
    # iterations j+1 and j+2 can be in flight while the hooks of j read j's.
    ahead = collections.deque()   # iterations launched ahead, oldest first
    phiT_ring = [None, None, None]
    X_ring = [None, None, None]

    def ring_phiT(m):
        if phiT_ring[m % 3] is None:
            phiT_ring[m % 3] = eng.new_states()
        return phiT_ring[m % 3]

    def ring_X(m):
        if X_ring[m % 3] is None:
            X_ring[m % 3] = eng.X if all(x is None for x in X_ring) \
                else eng.new_state_store()
        return X_ring[m % 3]

    h2d_setup, d2h_setup = eng.h2d_bytes, eng.d2h_bytes   # traffic before the loop

    # ---- main loop (optimize.py:393-577) ----------------------------------
    for krotov_iteration in range(iter_start + 1, iter_stop + 1):
        logger.info("Started Krotov iteration %d", krotov_iteration)
        tic = time.time()
        ev0 = ev1 = None
        if not host_loop:
            ev0 = torch.cuda.Event(enable_timing=True)
            ev1 = torch.cuda.Event(enable_timing=True)
            ev0.record(torch.cuda.current_stream(eng.device))

        # boundary condition chi(T), normalised (optimize.py:404-410)
        chi_states = chi_norms = None
        if chi_kind is None:
            chis = chi_constructor(fw_states_T=fw_states_T,
                                   objectives=objectives, tau_vals=tau_vals)
            chi_norms = list(eng.chi_from_host(
                [cp.vec(c) for c in chis[lo:hi]]))

        # phi(T) of the previous iteration is overwritten by the sweep below
        if isinstance(fw_states_T, _LazyFinalStates):
            fw_states_T.freeze(needed=False)

        # forward propagation and pulse update need sigma at the midpoints
        sigma_t = None
        if second_order:
            sig = np.array([
                float(sigma(tlist[n] + 0.5 * (tlist[n + 1] - tlist[n])))
                for n in range(NT)])
            sigma_t = eng.upload(sig, torch.float64)

        def sweep_iteration():
            """backward propagation under the guess pulses
            (optimize.py:413-425), forward propagation and pulse update
            (:427-500), tau (:503-508): one kernel launch each"""
            if chi_kind is not None:
                eng.chi_builtin(chi_kind, phiT, tau_t, K_total=K_total,
                                shard=shard)
            eng.sweep_backward(guess_t)
            new_phiT = eng.sweep_forward_update(
                guess_t, opt_t, phiT=spare_phiT, sigma_t=sigma_t, Phi0=Phi0,
                Phi1=Phi1)
            new_tau = eng.overlaps(eng.t_targets, new_phiT, out=spare_tau) \
                if has_targets else None
            return new_phiT, new_tau

        if packed is not None:
            # this iteration's outputs live in one buffer
            ri = krotov_iteration % 3
            pv = packed.views[ri]
            opt_t = pv['pulses']
            eng.g_a = pv['g_a']
            if sharded_packed:
                spare['tau'] = ring_loc(krotov_iteration)['tau']
                spare['phiT'] = ring_loc(krotov_iteration)['phiT']
            else:
                spare['tau'] = pv['tau']
                spare['phiT'] = ring_phiT(krotov_iteration)
            if info_hook is not None:
                eng.X = ring_X(krotov_iteration)
        spare_phiT = spare['phiT']
        spare_tau = spare['tau'] if has_targets else None
        # the buffer about to receive the optimized pulses still holds the
        # guess of the iteration before: a starting hint for the fused kernel
        prev_guess_t = prev_guess_ref
        ran_fused = False
        launch_epoch = None
        if ahead and ahead[0]['iteration'] != krotov_iteration:
            ahead.clear()   # cannot happen; never consume a stale launch
        if ahead:
            # this iteration was launched ahead, while the hooks of an
            # earlier one were running; its buffers are exactly the ones
            # selected above
            ran_fused = True
            launch_epoch = ahead.popleft()['epoch']
        elif use_fused:
            try:
                # chi boundary, both sweeps and tau in ONE launch
                eng.krotov_iteration(
                    chi_kind, guess_t, opt_t, phiT, tau_t, spare_phiT,
                    spare_tau, store_X=info_hook is not None,
                    sigma_t=sigma_t, Phi0=Phi0, Phi1=Phi1,
                    prev_guess_t=prev_guess_t,
                    diag_t=pv['diag'] if packed is not None else None)
                ran_fused = True
                launch_epoch = eng.epoch
                if packed is not None:
                    if sharded_packed:
                        gather_into_packed(krotov_iteration)
                    packed.mark(ri, eng)
            except KqError as exc:
                if exc.status != -3:     # KQ_ERR_UNSUPPORTED
                    raise
                use_fused = False      # outside the fused kernel family
        fetched = None
        if ran_fused and host_loop:
            if packed is not None:
                fetched = packed.fetch(ri, eng, diag_written=True)
                fb_epoch = int(fetched[3][1])
            else:
                fb_epoch, _ = eng.sweep_diagnostics()  # synchronises
            if fb_epoch == (launch_epoch & 0xFFFFFFFF):
                # the fixed-point iteration did not converge: outputs are
                # untouched; repeat with the sweep kernels and stay there
                ran_fused = use_fused = False
                fetched = None
                ahead.clear()   # launched with the outputs this one never wrote
                eng.clear_fused_failure()
        if ran_fused:
            new_phiT, new_tau = spare_phiT, spare_tau
        else:
            if chi_kind is None:
                pass    # chi was uploaded by chi_from_host above
            new_phiT, new_tau = sweep_iteration()
            if sharded_packed:
                if new_phiT is not spare_phiT:
                    spare_phiT.copy_(new_phiT)
                gather_into_packed(krotov_iteration)
        spare['phiT'] = phiT if phiT is not None else eng.new_states()
        if has_targets:
            spare['tau'] = tau_t if tau_t is not None else torch.empty(
                cp.K, dtype=torch.complex128, device=eng.device)
        phiT, tau_t = new_phiT, new_tau
        any_fused = any_fused or ran_fused
        n_fused += int(ran_fused)
        prev_guess_ref = guess_t

        if not host_loop:
            ev1.record(torch.cuda.current_stream(eng.device))
            deferred.append((krotov_iteration,
                             None if tau_t is None else tau_t.clone(),
                             opt_t.clone() if store_all_pulses else None,
                             ev0, ev1))
            # prepare next iteration (optimize.py:564): guess <- optimized
            guess_t, opt_t = opt_t, guess_t
            continue

        # ---- host bookkeeping (hooks present) -----------------------------
        # the guess pulses of this iteration are the optimized pulses of the
        # previous one and are already on the host
        guess_pulses_host = optimized_pulses if krotov_iteration > \
            first_iteration else guess_pulses
        X_this = eng.X
        if packed is not None:
            if fetched is None:
                fetched = packed.fetch(ri, eng)    # synchronises the stream
            if (ran_fused and use_fused and chi_kind is not None
                    and not second_order):
                # Launch the NEXT TWO iterations now, so that the device never
                # waits for the host: while the hooks of iteration j run, j+1
                # executes and j+2 is queued behind it.  Each reads the
                # outputs of the one before and writes only buffers nothing
                # else refers to (sets (j+1) % 3, (j+2) % 3); if a hook then
                # modifies the pulses or lambda_a, or the loop ends, the
                # launches are simply discarded (and repeated with the
                # modified inputs).
                last = ahead[-1] if ahead else dict(
                    iteration=krotov_iteration, guess=guess_t, opt=opt_t,
                    phiT=phiT, tau=tau_t)
                while len(ahead) < 2 and \
                        last['iteration'] < static['iter_stop']:
                    m = last['iteration'] + 1
                    pvm = packed.views[m % 3]
                    eng.g_a = pvm['g_a']
                    if info_hook is not None:
                        eng.X = ring_X(m)
                    if sharded_packed:
                        phiT_m = ring_loc(m)['phiT']
                        tau_m = ring_loc(m)['tau'] if has_targets else None
                    else:
                        phiT_m = ring_phiT(m)
                        tau_m = pvm['tau'] if has_targets else None
                    eng.krotov_iteration(
                        chi_kind, last['opt'], pvm['pulses'], last['phiT'],
                        last['tau'], phiT_m, tau_m,
                        store_X=info_hook is not None,
                        prev_guess_t=last['guess'], diag_t=pvm['diag'])
                    if sharded_packed:
                        gather_into_packed(m)
                    packed.mark(m % 3, eng)
                    last = dict(iteration=m, guess=last['opt'],
                                opt=pvm['pulses'], phiT=phiT_m, tau=tau_m,
                                epoch=eng.epoch)
                    ahead.append(last)
            optimized_pulses = [fetched[0][l].copy() for l in range(L)]
            g_a_integrals[:] = fetched[1][:L]
            phi_host = None
            if sharded_packed:
                tau_all, phi_host = split_gathered(fetched[2])
                tau_vals = tau_all if tau_t is not None \
                    else np.array([None] * K)
            else:
                tau_vals = fetched[2].copy() if tau_t is not None \
                    else np.array([None] * K)
            st = int(fetched[3][0])
        else:
            optimized_pulses = pulses_to_host(opt_t)   # synchronises
            g_a_integrals[:] = eng.download(eng.g_a)[:L]
            tau_vals = tau_to_host(tau_t)
            st = eng.status()
        if st != 0:
            raise RuntimeError("sweep kernel reported exchange failure %d"
                               % st)
        _check_finite(optimized_pulses, krotov_iteration)
        if packed is not None and sharded_packed:
            # the final states of all ranks came with the packed copy
            fw_states_T = _LazyFinalStates(
                lambda ph=phi_host: [cp.unvec(ph[k].copy(), templates[k])
                                     for k in range(K)], K)
        else:
            fw_states_T = lazy_states(phiT)
        backward_states = _LazyStates(X_this, cp, eng)
        if second_order:
            forward_states = _LazyStates(Phi1, cp, eng)
            forward_states0 = _LazyStates(Phi0, cp, eng)
        toc = time.time()

        if info_hook is not None:
            opt_snapshot = [p.copy() for p in optimized_pulses]
            info = info_hook(
                backward_states=backward_states,
                forward_states=forward_states,
                forward_states0=forward_states0, fw_states_T=fw_states_T,
                guess_pulses=guess_pulses_host,
                optimized_pulses=optimized_pulses,
                g_a_integrals=g_a_integrals, tau_vals=tau_vals,
                start_time=tic, stop_time=toc, info_vals=result.info_vals,
                shared_data={}, iteration=krotov_iteration, **static)
            # hooks may have modified lambda_vals / optimized_pulses
            if not np.array_equal(lam_snapshot, np.asarray(lambda_vals)):
                eng.set_lambda(lambda_vals)
                lam_snapshot = np.array(lambda_vals, dtype=np.float64)
                ahead.clear()    # launched ahead with the old lambda_a
            if not all(np.array_equal(a, b) for a, b
                       in zip(optimized_pulses, opt_snapshot)):
                opt_t.copy_(eng.pulses_to_device(optimized_pulses))
                ahead.clear()    # launched ahead with the unmodified pulses
        result.iters.append(krotov_iteration)
        result.iter_seconds.append(int(toc - tic))
        result.iter_seconds_device.append(toc - tic)
        if info is not None:
            result.info_vals.append(info)
        if not np.all(tau_vals == None):  # noqa: E711
            result.tau_vals.append(tau_vals)
        result.optimized_controls = optimized_pulses
        if store_all_pulses:
            result.all_pulses.append(copy.deepcopy(optimized_pulses))
        result.states = fw_states_T
        logger.info("Finished Krotov iteration %d", krotov_iteration)

        msg = None
        if check_convergence is not None:
            msg = check_convergence(result)
        if krotov_iteration >= static['iter_stop']:
            iter_stop = static['iter_stop']
            result.message = "Reached %d iterations" % iter_stop
            finished_by_break = True
        elif bool(msg) is True:
            result.message = "Reached convergence"
            if isinstance(msg, str):
                result.message += ": " + msg
            finished_by_break = True
        if finished_by_break:
            backward_states._expire()
            break
        # prepare next iteration
        guess_t, opt_t = opt_t, guess_t
        if second_order:
            if chi_states is None:
                chi_arr = eng.download(eng.chi)
                chi_states = [cp.unvec(chi_arr[k].copy(), templates[k])
                              for k in range(K)]
                if chi_norms is None:
                    chi_norms = list(eng.download(eng.chi_norms))
            sigma.refresh(
                forward_states=forward_states,
                forward_states0=forward_states0, chi_states=chi_states,
                chi_norms=chi_norms, optimized_pulses=optimized_pulses,
                guess_pulses=optimized_pulses, objectives=objectives,
                result=result)
            forward_states._expire()
            forward_states0._expire()
            Phi0, Phi1 = Phi1, Phi0
        backward_states._expire()

    if not finished_by_break:
        result.message = "Reached %d iterations" % max(iter_start, iter_stop)

    # ---- fast path epilogue: one synchronisation for all iterations --------
    if deferred:
        torch.cuda.synchronize(eng.device)
        st = eng.status()
        if st != 0:
            raise RuntimeError("sweep kernel reported exchange failure %d"
                               % st)
        if any_fused and eng.first_failed_epoch() != 0:
            # a fused iteration did not converge somewhere along the way; no
            # hook has seen anything yet, so simply redo the run with the
            # sweep kernels
            logger.info("time-parallel iteration did not converge; "
                        "repeating with the sweep kernels")
            if shard is not None:
                shard.close()
            if gather_comm is not None:
                gather_comm.close()
            return optimize_pulses(
                objectives, pulse_options, tlist, propagator=propagator,
                chi_constructor=chi_constructor, mu=mu, sigma=sigma,
                iter_start=iter_start, iter_stop=iter_stop,
                check_convergence=check_convergence, info_hook=info_hook,
                storage=storage, parallel_map=parallel_map,
                store_all_pulses=store_all_pulses,
                continue_from=continue_from,
                skip_initial_forward_propagation=(
                    skip_initial_forward_propagation), norm=norm,
                overlap=overlap, limit_thread_pool=limit_thread_pool,
                device=device, engine_mode='sweeps')
        for (it, tau_i, pulses_i, e0, e1) in deferred:
            secs = e0.elapsed_time(e1) * 1e-3
            result.iters.append(it)
            result.iter_seconds.append(int(secs))
            result.iter_seconds_device.append(secs)
            if tau_i is not None:
                result.tau_vals.append(tau_to_host(tau_i))
            if pulses_i is not None:
                result.all_pulses.append(pulses_to_host(pulses_i))
        # after the swap at the end of the loop the optimized pulses of the
        # last iteration are in guess_t
        optimized_pulses = pulses_to_host(guess_t)
        _check_finite(optimized_pulses, deferred[-1][0])
        result.optimized_controls = optimized_pulses
        result.states = states_to_host(phiT)

    # ---- finalize (optimize.py:583-590) -----------------------------------
    if isinstance(result.states, _LazyFinalStates):
        result.states = list(result.states)
    result.end_local_time = time.localtime()
    result.optimized_controls = [
        pulse_onto_tlist(np.asarray(p)) for p in result.optimized_controls]
    result.gpu_launches = eng.launches
    # diagnostics of the last update sweep: fixed-point rounds (0: sequential
    # kernel) and whether the sequential kernel had to take over
    _fb, _rounds = eng.sweep_diagnostics()
    result.update_sweep_rounds = _rounds
    result.sequential_fallback = bool(_fb != 0 and _fb == (eng.epoch & 0xFFFFFFFF))
    result.fused_iterations = n_fused
    result.h2d_bytes, result.d2h_bytes = eng.h2d_bytes, eng.d2h_bytes
    result.h2d_bytes_loop = eng.h2d_bytes - h2d_setup
    result.d2h_bytes_loop = eng.d2h_bytes - d2h_setup
    if shard is not None:
        shard.close()
    if gather_comm is not None:
        gather_comm.close()
    return result
