"""Device side of the sweep engine: owns the torch CUDA tensors of one
compiled problem and issues the C-ABI calls (include/krotov_b200.h).

PyTorch is used for device memory and streams only; every sweep is one launch
of a hand-written sm_100a kernel from ``libkrotov_b200.so``.  All calls are
asynchronous on the current torch stream.  There is no CPU fallback: without
a CUDA device or without the built library construction raises
:class:`krotov_b200._lib.EngineUnavailable`.
"""
import ctypes

import numpy as np

from . import _lib
from ._lib import EngineUnavailable, KqComm, KqProblem, KqSparse, check

__all__ = ['SweepEngine', 'CHI_KINDS']

CHI_KINDS = {'re': 0, 'ss': 1, 'sm': 2, 'hs': 3}


def _ptr(t):
    return ctypes.c_void_p(0 if t is None else t.data_ptr())


class SweepEngine:
    """Device tensors + launches for one :class:`CompiledProblem`.

    Attributes (torch tensors on the device):
        X: backward states ``[NT+1, K, N]`` (time-major), filled by
            :meth:`sweep_backward`.
        chi, chi_norms: normalised boundary states ``[K, N]`` and their norms.
    """

    def __init__(self, cp, shape_arrays, lambda_vals, device=None):
        import torch
        self.torch = torch
        self.lib = _lib.load()
        if not torch.cuda.is_available():
            raise EngineUnavailable(
                "krotov_b200 needs a CUDA device (sm_100a); there is no CPU "
                "fallback for the sweep kernels")
        self.device = torch.device(
            'cuda', torch.cuda.current_device()) if device is None \
            else torch.device(device)
        self.cp = cp
        self.launches = 0       # kernels launched through the C ABI
        self.h2d_bytes = 0      # host->device traffic issued by the engine
        self.d2h_bytes = 0
        dev = self.device
        c128, f64 = torch.complex128, torch.float64

        def up(a, dtype):
            t = torch.as_tensor(np.ascontiguousarray(a), dtype=dtype,
                                device=dev)
            self.h2d_bytes += t.numel() * t.element_size()
            return t

        self._sparse = None
        if cp.sparse is not None:
            # N > 64: CSR matrices (csrc/kq_csr.cuh); no dense copies
            self.t_ops = self.t_ops_adj = self.t_mu = None
            sp = cp.sparse
            self.t_csr = [up(sp['row_ptr'], torch.int32),
                          up(sp['mat_off'], torch.int64),
                          up(sp['col'], torch.int32), up(sp['val'], c128)]
            coded = sp['dict'] is not None
            if coded:
                # uint16 arrays travel as int16 bit patterns
                self.t_csr += [
                    up(sp['col16'].view(np.int16), torch.int16),
                    up(sp['code16'].view(np.int16), torch.int16),
                    up(sp['dict'], c128)]
            ptrs = [t.data_ptr() for t in self.t_csr] + [0] * (
                0 if coded else 3)
            self._sparse = KqSparse(
                *ptrs, len(sp['dict']) if coded else 0,
                sp['stage_nnz_update'], sp['stage_nnz_prop'])
        else:
            self.t_ops = up(cp.ops, c128)
            self.t_ops_adj = up(cp.ops_adj, c128)
            self.t_mu = up(cp.mu, c128)
        self.t_t2p = up(cp.term2pulse, torch.int32)
        self.t_opn = up(cp.op_norm, f64)
        self.t_dt = up(cp.dt, f64)
        self.t_shape = up(np.array(shape_arrays, dtype=np.float64).reshape(
            cp.L, cp.NT), f64)
        self.t_lambda = up(np.asarray(lambda_vals, dtype=np.float64), f64)
        self.t_psi0 = up(cp.psi0, c128)
        self.t_targets = None if cp.targets is None else up(cp.targets, c128)
        self.t_weights = None if cp.weights is None else up(cp.weights, f64)
        # few objectives with N = 3, 4: the delta-polynomial iteration
        # (csrc/kq_dpoly.cuh) runs its sequential chain in ONE warp when
        # K (N + 1) <= 32 lanes and is then at least as fast as the
        # time-parallel fixed point, which would leave most of the GPU idle;
        # strongly coupled problems (the number of fixed-point rounds grows
        # with T * (S/lambda) * sum_k ||chi_k|| ||mu_k||^2) prefer it up to
        # K = 16 anyway
        update_sweep = 0
        if cp.M == 2 and cp.L == 1 and cp.N in (3, 4) and cp.K <= 16:
            lam0 = float(np.asarray(lambda_vals, dtype=np.float64)[0])
            smax = float(np.max(np.abs(np.asarray(shape_arrays[0]))))
            mu2 = max(np.linalg.norm(np.asarray(cp.mu[k, 0]).reshape(
                cp.N, cp.N), 2) for k in range(cp.K)) ** 2
            coupling = float(np.sum(cp.dt)) * smax / lam0 * 0.5 * mu2
            if coupling > 4.0 or cp.K * (cp.N + 1) <= 32:
                update_sweep = 1
        self.update_sweep = update_sweep
        self.problem = KqProblem(
            K=cp.K, N=cp.N, NT=cp.NT, L=cp.L, M=cp.M,
            is_super=1 if cp.is_super else 0,
            ops=0 if self.t_ops is None else self.t_ops.data_ptr(),
            ops_adj=0 if self.t_ops_adj is None else self.t_ops_adj.data_ptr(),
            mu=0 if self.t_mu is None else self.t_mu.data_ptr(),
            term2pulse=self.t_t2p.data_ptr(),
            op_norm=self.t_opn.data_ptr(), dt=self.t_dt.data_ptr(),
            shape=self.t_shape.data_ptr(), lambda_a=self.t_lambda.data_ptr(),
            real_ops=1 if cp.real_ops else 0, reserved=0,
            update_sweep=update_sweep,
            row_nnz=int(getattr(cp, 'row_nnz', 0) or 0),
            sparse=0 if self._sparse is None else ctypes.addressof(
                self._sparse))
        self._p = ctypes.byref(self.problem)
        nbytes = self.lib.kq_workspace_bytes(self._p)
        self.workspace = torch.zeros(nbytes, dtype=torch.uint8, device=dev)
        self._iter_args = {}    # marshalled kq_krotov_iteration arguments
        self.epoch = 0
        self.comm = None      # KqComm: cross-GPU exchange (objectives sharded)
        self.shard = None     # ShardComm owning `comm` (NCCL collectives)
        self.K_total = cp.K   # objectives over all ranks
        self.gather = None    # dict set by ShardComm.attach_gather
        K, N, NT = cp.K, cp.N, cp.NT
        self.X = torch.empty((NT + 1, K, N), dtype=c128, device=dev)
        self.chi = torch.empty((K, N), dtype=c128, device=dev)
        self.chi_norms = torch.empty(K, dtype=f64, device=dev)
        self.g_a = torch.zeros(max(cp.L, 1), dtype=f64, device=dev)

    # -- helpers -----------------------------------------------------------
    def _stream(self):
        return ctypes.c_void_p(self.raw_stream())

    def raw_stream(self):
        """cudaStream_t of torch's current stream on this device."""
        try:
            return self.torch._C._cuda_getCurrentRawStream(self.device.index)
        except Exception:   # pragma: no cover
            return self.torch.cuda.current_stream(self.device).cuda_stream

    def new_state_store(self):
        return self.torch.empty((self.cp.NT + 1, self.cp.K, self.cp.N),
                                dtype=self.torch.complex128,
                                device=self.device)

    def new_states(self):
        return self.torch.empty((self.cp.K, self.cp.N),
                                dtype=self.torch.complex128,
                                device=self.device)

    def upload(self, array, dtype):
        t = self.torch.as_tensor(np.ascontiguousarray(array), dtype=dtype,
                                 device=self.device)
        self.h2d_bytes += t.numel() * t.element_size()
        return t

    def download(self, tensor):
        self.d2h_bytes += tensor.numel() * tensor.element_size()
        return tensor.cpu().numpy()

    def pulses_to_device(self, pulses):
        arr = np.array(pulses, dtype=np.float64).reshape(self.cp.L,
                                                         self.cp.NT)
        return self.upload(arr, self.torch.float64)

    def set_lambda(self, lambda_vals):
        """Refresh the device copy of lambda_a (hooks may change it between
        iterations, optimize.py:552-555 / tests/test_infohooks.py:30-37)."""
        self.t_lambda.copy_(self.torch.as_tensor(
            np.asarray(lambda_vals, dtype=np.float64)))
        self.h2d_bytes += 8 * self.cp.L

    def plan(self):
        """(family, grid, block, smem) the fused sweep uses."""
        vals = [ctypes.c_int32() for _ in range(4)]
        check(self.lib.kq_plan(self._p, *[ctypes.byref(v) for v in vals]))
        return tuple(v.value for v in vals)

    # -- sweeps ------------------------------------------------------------
    def propagate_forward(self, pulses_t, phiT=None, store=None,
                          state0=None):
        """Initial forward propagation (optimize.py:806-846)."""
        if phiT is None:
            phiT = self.new_states()
        s0 = self.t_psi0 if state0 is None else state0
        check(self.lib.kq_propagate_forward(
            self._p, _ptr(pulses_t), _ptr(s0), _ptr(phiT), _ptr(store),
            self._stream()))
        self.launches += 1
        return phiT

    def sweep_backward(self, pulses_t):
        """Backward propagation of :attr:`chi` into :attr:`X`
        (optimize.py:849-886)."""
        g = self.gather
        if g is None:
            check(self.lib.kq_sweep_backward(
                self._p, _ptr(pulses_t), _ptr(self.chi), _ptr(self.X),
                self._stream()))
        else:
            # 'gather' mode: this rank's block of objectives, stored on every
            # GPU; the stores alternate so that a fast rank cannot overwrite
            # rows a slower rank is still reading in its fused sweep
            g['flip'] ^= 1
            self.X = g['stores'][g['flip']]
            check(self.lib.kq_sweep_backward_range(
                self._p, _ptr(pulses_t), _ptr(self.chi), g['tables'][g['flip']],
                g['comm'].world, g['comm'].rank, g['lo'], g['hi'] - g['lo'],
                self._stream()))
            self.launches += 1          # the column-block scatter kernel
            g['comm'].stream_barrier()
        self.launches += 1
        return self.X

    def sweep_forward_update(self, guess_t, opt_t, phiT=None, sigma_t=None,
                             Phi0=None, Phi1=None):
        """Fused pulse update + forward propagation (optimize.py:449-500)."""
        if phiT is None:
            phiT = self.new_states()
        self.epoch += 1
        comm = ctypes.byref(self.comm) if self.comm is not None else None
        check(self.lib.kq_sweep_forward_update(
            self._p, _ptr(guess_t), _ptr(opt_t), _ptr(self.X),
            _ptr(self.chi_norms), _ptr(self.t_psi0), _ptr(phiT),
            _ptr(sigma_t), _ptr(Phi0), _ptr(Phi1), _ptr(self.g_a), comm,
            _ptr(self.workspace), ctypes.c_uint32(self.epoch & 0xFFFFFFFF),
            self._stream()))
        self.launches += 1
        return phiT

    def krotov_iteration(self, chi_kind, guess_t, opt_t, phiT_in, tau_in,
                         phiT_out, tau_out, store_X=False, sigma_t=None,
                         Phi0=None, Phi1=None, prev_guess_t=None,
                         diag_t=None):
        """One whole Krotov iteration (optimize.py:393-508) in one launch of
        the time-parallel kernel family: chi boundary (`chi_kind` 're', 'ss',
        'sm', 'hs', or None for the states already in :attr:`chi` /
        :attr:`chi_norms`), backward sweep, update + forward sweep, tau.
        Raises :class:`KqError` (KQ_ERR_UNSUPPORTED) if the problem is outside
        that family.  `prev_guess_t` (may be `opt_t`): guess pulses of the
        iteration before, a starting hint for the fixed-point iteration."""
        self.epoch += 1
        kind = -1 if chi_kind is None else CHI_KINDS[chi_kind]
        # the pointer arguments repeat with the rotating buffer sets of the
        # caller: marshal them once per combination (the cache keeps the
        # tensors alive, so an id cannot be re-used by another tensor)
        tau_sum = None
        if kind == CHI_KINDS['sm'] and self.comm is not None:
            # sum_j w_j tau_j over the objectives of ALL ranks (NCCL, on the
            # compute stream; no host synchronisation)
            w = tau_in if self.t_weights is None else tau_in * self.t_weights
            tau_sum = self._tau_sum = w.sum().reshape(1)
            self.shard.all_reduce_sum(tau_sum)
        tensors = (tau_in, phiT_in, guess_t, prev_guess_t, opt_t, phiT_out,
                   tau_out, self.X if store_X else None, sigma_t, Phi0, Phi1,
                   self.g_a, diag_t, tau_sum)
        key = (kind,) + tuple(map(id, tensors))
        hit = self._iter_args.get(key)
        if hit is None:
            if len(self._iter_args) >= 32:
                self._iter_args.clear()

            def ptr(t):
                return 0 if t is None else t.data_ptr()
            hit = (tensors, (
                self._p, kind, self.K_total, ptr(self.t_targets),
                ptr(self.t_weights),
                ptr(self.chi if kind < 0 else None),
                ptr(self.chi_norms if kind < 0 else None),
                ptr(tau_in), ptr(phiT_in), ptr(guess_t), ptr(prev_guess_t),
                ptr(opt_t), ptr(self.t_psi0), ptr(phiT_out), ptr(tau_out),
                ptr(self.X if store_X else None),
                ptr(None if kind < 0 else self.chi),
                ptr(None if kind < 0 else self.chi_norms),
                ptr(sigma_t), ptr(Phi0), ptr(Phi1), ptr(self.g_a),
                ptr(diag_t),
                None if self.comm is None else ctypes.byref(self.comm),
                ptr(tau_sum), ptr(self.workspace)))
            self._iter_args[key] = hit
        check(self.lib.kq_krotov_iteration(
            *hit[1], self.epoch & 0xFFFFFFFF, self.raw_stream()))
        self.launches += 1

    def fused_supported(self):
        """True if :meth:`krotov_iteration` handles this problem."""
        cp = self.cp
        # the library declines (KQ_ERR_UNSUPPORTED) what its kernel families
        # do not cover: N <= 4 runs the time-parallel fixed-point kernel, few
        # objectives with N >= 3 the delta-polynomial sweep (csrc/kq_dpoly.cuh)
        if self.gather is None and cp.M == 2 and cp.L == 1 and 2 <= cp.N <= 16:
            return True
        # few objectives with several controls or sparse rows: the
        # entries-in-registers kernels (csrc/kq_lanes.cuh), one call per
        # iteration as well
        nnz = getattr(cp, 'row_nnz', 0) or cp.N
        return (self.gather is None and self.comm is None
                and cp.ops is not None and 2 <= cp.N <= 32 and 1 <= cp.L <= 4
                and cp.M <= 5 and nnz <= 4)

    def clear_fused_failure(self):
        """Reset the 'first failed epoch' status word."""
        self.workspace[12:16].zero_()

    def overlaps(self, a, b, out=None):
        """tau_k = <a_k|b_k> (optimize.py:316-322, 503-508)."""
        if out is None:
            out = self.torch.empty(self.cp.K, dtype=self.torch.complex128,
                                   device=self.device)
        check(self.lib.kq_overlaps(self.cp.K, self.cp.N, _ptr(a), _ptr(b),
                                   _ptr(out), self._stream()))
        self.launches += 1
        return out

    def chi_builtin(self, kind, phiT, tau_t, K_total=None, shard=None):
        """Device chi-constructor + normalisation (optimize.py:404-410).
        With `shard` (a ShardComm) the 1/N prefactors use the global number
        of objectives and the chis_sm sum runs over all ranks."""
        tau_sum = None
        if kind == 'sm':
            w = tau_t if self.t_weights is None else tau_t * self.t_weights
            tau_sum = w.sum().reshape(1)
            if shard is not None:
                shard.all_reduce_sum(tau_sum)
        check(self.lib.kq_chi_boundary(
            self._p, CHI_KINDS[kind],
            self.cp.K if K_total is None else K_total, _ptr(phiT),
            _ptr(self.t_targets), _ptr(tau_t), _ptr(self.t_weights),
            _ptr(tau_sum), _ptr(self.chi), _ptr(self.chi_norms),
            self._stream()))
        self.launches += 1

    def chi_from_host(self, chi_vectors):
        """Upload host chi states [K,N] (custom chi_constructor), normalise
        with the L2/Frobenius norm."""
        arr = np.asarray(chi_vectors, dtype=np.complex128).reshape(
            self.cp.K, self.cp.N)
        norms = np.sqrt((arr.real ** 2 + arr.imag ** 2).sum(axis=1))
        # a zero chi (target reached exactly) stays zero instead of 0/0
        arr = arr / np.where(norms > 0, norms, 1.0)[:, None]
        self.chi.copy_(self.torch.as_tensor(arr))
        self.chi_norms.copy_(self.torch.as_tensor(norms))
        self.h2d_bytes += arr.nbytes + norms.nbytes
        return norms

    def status(self):
        """Exchange status word (0 = ok); synchronises."""
        self.d2h_bytes += 4
        return int(self.workspace[:4].view(self.torch.int32).item())

    def sweep_diagnostics(self):
        """(fallback_epoch, picard_iterations) of the last fused sweep:
        status words 1 and 2 of the workspace (kq_picard.cuh); synchronises.
        ``fallback_epoch == self.epoch`` means the time-parallel sweep did not
        converge and the sequential kernel produced the result."""
        w = self.workspace[:16].view(self.torch.int32).cpu()
        self.d2h_bytes += 16
        return int(w[1]), int(w[2])

    def first_failed_epoch(self):
        """Epoch of the first fused iteration that did not converge (0 =
        none); synchronises."""
        self.d2h_bytes += 4
        return int(self.workspace[12:16].view(self.torch.int32).item())
