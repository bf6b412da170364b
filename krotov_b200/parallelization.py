"""Multi-GPU execution: objectives sharded over the ranks of a
``torch.distributed`` process group (one process per GPU).

The reference parallelises over objectives with process pools
(/root/reference/src/krotov/parallelization.py:233-604): whole state histories
are pickled back after the backward sweep and, in the sequential update sweep,
one queue round trip per time step sends the new pulse values to K worker
processes.  Here the same data-parallel split runs on the GPUs of one
NVLink/NVSwitch box:

* each rank owns a contiguous block of objectives (operators, chi store and
  forward states never leave its HBM);
* backward sweep and initial forward propagation need no communication;
* the fused update/forward sweep needs the sum over ALL objectives at every
  time step (optimize.py:454-470).  Each rank's kernel writes its partial sum
  straight into every peer's exchange buffer (flag-tagged 16-byte stores over
  NVLink, CUDA-IPC mapped memory) and polls its own buffer; ranks add the
  partials in rank order, so every GPU computes bit-identical pulses.  No
  host or NCCL call sits inside the time loop;
* once per iteration NCCL all-gathers tau / final states for the host
  callbacks and all-reduces the scalar needed by ``chis_sm``.

Pass ``parallel_map=GPUShards()`` to :func:`krotov_b200.optimize_pulses` from
every rank of a ``torchrun`` job.  The process-pool maps of the reference
(``parallel_map``, ``parallel_map_fw_prop_step``) are accepted by
``optimize_pulses`` for signature compatibility and ignored.
"""
import ctypes

import numpy as np

from . import _lib
from ._lib import KqComm, check

__all__ = ['GPUShards', 'shard_bounds', 'ShardComm', 'USE_THREADPOOL_LIMITS',
           'serial_map', 'parallel_map', 'parallel_map_fw_prop_step',
           'set_parallelization']

USE_THREADPOOL_LIMITS = True


def shard_bounds(K, world, rank):
    """Half-open range of the objectives owned by `rank`: contiguous blocks
    whose sizes differ by at most one (the first ``K % world`` ranks hold one
    more)."""
    base, extra = divmod(K, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


class GPUShards:
    """Selector for multi-GPU execution (value of `parallel_map`).

    Args:
        group: ``torch.distributed`` process group (default: WORLD).  The
            group must already be initialised (backend ``nccl`` on GPUs).
        mode: how the sequential update sweep is distributed.

            ``'exchange'``: every rank owns a block of objectives for all
            sweeps; the fused sweep kernels exchange their partial sums over
            NVLink at every time step (about 2 microseconds per step).  Pays
            off when the per-step work of a block is much larger than that.

            ``'gather'``: the backward sweep is sharded and each kernel writes
            the backward states it produces into the stores of ALL GPUs (P2P
            stores); after one NCCL barrier every GPU runs the complete fused
            sweep redundantly, with no per-step exchange.  Best for small
            state vectors, where the sweep is bound by the chain of nt-1
            dependent steps and not by throughput.

            ``'sharded'``: every rank owns a block of objectives for all
            sweeps and runs the one-launch time-parallel iteration kernel
            (``kq_krotov_iteration``) on it; in every fixed-point round of
            that kernel the per-time-step sums over the objectives cross the
            GPUs once (nt doubles per GPU written straight into the peers'
            exchange buffers over NVLink, added in rank order), so all ranks
            obtain bit-identical pulses.  Iterations the kernel family does
            not cover fall back to ``'exchange'``.

            ``'replicate'``: every rank runs the complete problem with the
            one-launch time-parallel iteration kernel (``kq_krotov_iteration``)
            and no communication at all; all ranks obtain identical pulses
            (the kernel is deterministic).  For problems that fit that kernel
            family a single GPU is already latency-bound, so sharding them
            only adds exchange latency: this is the fastest correct choice.

            ``'auto'`` (default): ``'sharded'`` for N <= 4 and at most
            1184 objectives per GPU, else ``'gather'`` if
            ``K * N * N <= 65536``, else ``'exchange'``.
    """

    def __init__(self, group=None, mode='auto'):
        if mode not in ('auto', 'exchange', 'gather', 'replicate', 'sharded'):
            raise ValueError("mode must be 'auto', 'sharded', 'exchange', "
                             "'gather' or 'replicate'")
        self.group = group
        self.mode = mode

    def choose(self, K, N, world=1):
        if self.mode != 'auto':
            return self.mode
        if N <= 4 and -(-K // max(world, 1)) <= 1184:
            return 'sharded'
        return 'gather' if K * N * N <= 65536 else 'exchange'

    def resolve(self):
        import torch.distributed as dist
        if not dist.is_initialized():
            raise RuntimeError(
                "GPUShards needs an initialised torch.distributed process "
                "group (launch with torchrun)")
        return dist, self.group, dist.get_rank(self.group), \
            dist.get_world_size(self.group)


class ShardComm:
    """Exchange buffers and collectives of one rank."""

    def __init__(self, dist, group, device):
        import torch
        self.torch, self.dist, self.group = torch, dist, group
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)
        self.device = device
        self._own = None
        self._owned = []     # further IPC allocations owned by this rank
        self._peers = []
        self.slots_t = None

    def attach(self, eng, for_exchange=True):
        """Allocate this rank's IPC exchange buffer, map the peers' buffers
        and (for 'exchange' mode) hand the pointer table to the engine."""
        lib = _lib.load()
        nbytes = lib.kq_comm_slot_bytes(eng._p)
        ptr = ctypes.c_void_p()
        handle = ctypes.create_string_buffer(64)
        check(lib.kq_comm_alloc(nbytes, ctypes.byref(ptr), handle))
        self._own = ptr
        handles = [None] * self.world
        self.dist.all_gather_object(handles, handle.raw, group=self.group)
        ptrs = []
        for r in range(self.world):
            if r == self.rank:
                ptrs.append(ptr.value)
            else:
                peer = ctypes.c_void_p()
                check(lib.kq_comm_open(handles[r], ctypes.byref(peer)))
                self._peers.append(peer)
                ptrs.append(peer.value)
        self.slots_t = self.torch.tensor(ptrs, dtype=self.torch.int64,
                                         device=self.device)
        self.kqcomm = KqComm(rank=self.rank, world=self.world,
                             slots=self.slots_t.data_ptr())
        if for_exchange:
            eng.comm = self.kqcomm
            eng.shard = self
        self._eng = eng
        self._barrier_tag = 0
        self.dist.barrier(group=self.group)
        return self

    def attach_gather(self, eng):
        """'gather' mode: make both backward-state stores of `eng` writable by
        every peer (CUDA IPC handles of the torch allocations) and tell the
        engine which block of objectives this rank propagates backward."""
        torch = self.torch
        lib = _lib.load()
        self.attach(eng, for_exchange=False)   # flag slots for the barrier
        shape = tuple(eng.X.shape)
        nbytes = eng.X.numel() * eng.X.element_size()

        class _Raw:   # __cuda_array_interface__ view of an IPC allocation
            def __init__(self, ptr):
                self.__cuda_array_interface__ = dict(
                    data=(ptr, False), shape=shape + (2,), typestr='<f8',
                    version=2, strides=None)

        own, handles = [], []
        for _ in range(2):
            ptr = ctypes.c_void_p()
            handle = ctypes.create_string_buffer(64)
            check(lib.kq_comm_alloc(nbytes, ctypes.byref(ptr), handle))
            self._owned.append(ptr)
            own.append(ptr.value)
            handles.append(handle.raw)
        stores = tuple(
            torch.view_as_complex(torch.as_tensor(_Raw(p), device=self.device))
            for p in own)
        eng.X, eng.X2 = stores
        gathered = [None] * self.world
        self.dist.all_gather_object(gathered, handles, group=self.group)
        tables = []
        for b in range(2):
            ptrs = []
            for r in range(self.world):
                if r == self.rank:
                    ptrs.append(own[b])
                    continue
                peer = ctypes.c_void_p()
                check(lib.kq_comm_open(gathered[r][b], ctypes.byref(peer)))
                self._peers.append(peer)
                ptrs.append(peer.value)
            tables.append((ctypes.c_void_p * self.world)(*ptrs))
        lo, hi = shard_bounds(eng.cp.K, self.world, self.rank)
        eng.gather = dict(tables=tables, lo=lo, hi=hi, flip=0,
                          stores=(eng.X, eng.X2), comm=self)
        self.dist.barrier(group=self.group)
        return self

    def stream_barrier(self):
        """Order the kernels of all ranks on their streams: every rank's P2P
        stores of the backward sweep are complete before any rank starts the
        fused sweep.  One tiny kernel per rank (flags over NVLink), no NCCL
        call and no host synchronisation."""
        self._barrier_tag += 1
        eng = self._eng
        check(_lib.load().kq_comm_barrier(
            ctypes.byref(self.kqcomm),
            ctypes.c_uint32(self._barrier_tag & 0xFFFFFFFF or 1),
            ctypes.c_void_p(eng.workspace.data_ptr()), eng._stream()))
        eng.launches += 1

    def close(self):
        lib = _lib.load()
        self.torch.cuda.synchronize(self.device)
        self.dist.barrier(group=self.group)
        for p in self._peers:
            lib.kq_comm_close(p)
        self._peers = []
        if self._own is not None:
            lib.kq_comm_free(self._own)
            self._own = None
        for p in self._owned:
            lib.kq_comm_free(p)
        self._owned = []

    # -- once-per-iteration collectives (NCCL) ------------------------------
    def all_gather_rows(self, local, K_total):
        """Concatenate per-rank row blocks ``[K_r, ...]`` into ``[K, ...]``
        (ranks own :func:`shard_bounds` blocks)."""
        torch = self.torch
        kmax = -(-K_total // self.world)
        pad = torch.zeros((kmax,) + tuple(local.shape[1:]), dtype=local.dtype,
                          device=local.device)
        pad[:local.shape[0]] = local
        if local.dtype.is_complex:
            send = torch.view_as_real(pad).contiguous()
        else:
            send = pad.contiguous()
        out = [torch.empty_like(send) for _ in range(self.world)]
        self.dist.all_gather(out, send, group=self.group)
        parts = []
        for r in range(self.world):
            lo, hi = shard_bounds(K_total, self.world, r)
            blk = out[r][:hi - lo]
            parts.append(torch.view_as_complex(blk)
                         if local.dtype.is_complex else blk)
        return torch.cat(parts, dim=0)

    def all_gather_flat(self, out, local):
        """``out[r * n : (r + 1) * n] = local of rank r`` for equally sized
        complex vectors: one NCCL all-gather on the current stream, no
        temporaries."""
        v = self.torch.view_as_real
        self.dist.all_gather_into_tensor(v(out).view(-1), v(local).view(-1),
                                         group=self.group)
        return out

    def all_reduce_sum(self, t):
        """In-place sum over ranks of a real or complex tensor."""
        v = self.torch.view_as_real(t) if t.dtype.is_complex else t
        self.dist.all_reduce(v, group=self.group)
        return t


# --- names of the reference's process-level maps ------------------------------
# The reference parallelises over objectives with Python process pools
# (parallelization.py:233-311, 433-495).  Here all objectives of a sweep run in
# one kernel launch, so `optimize_pulses` accepts these maps in `parallel_map`
# for source compatibility and ignores them; called directly they are plain
# serial maps with the interface of qutip.parallel.serial_map.

def serial_map(task, values, task_args=None, task_kwargs=None, **kwargs):
    """``[task(value, *task_args, **task_kwargs) for value in values]``
    (interface of :func:`qutip.parallel.serial_map`, the reference's default
    map, optimize.py:266-269)."""
    task_args = () if task_args is None else task_args
    task_kwargs = {} if task_kwargs is None else task_kwargs
    return [task(value, *task_args, **task_kwargs) for value in values]


def parallel_map(task, values, task_args=None, task_kwargs=None,
                 num_cpus=None, progress_bar=None):
    """Signature of the reference's process-pool map
    (parallelization.py:233-240); evaluated serially -- the GPU engine does
    not use process-level parallelism."""
    return serial_map(task, values, task_args, task_kwargs)


def parallel_map_fw_prop_step(shared, values, task_args):
    """Marker for the third entry of the reference's ``parallel_map`` tuple
    (parallelization.py:433-495).  The per-time-step synchronisation it
    implements with consumer processes happens inside the update kernel
    here; `optimize_pulses` never calls it."""
    raise NotImplementedError(
        "parallel_map_fw_prop_step is only accepted as a marker in "
        "optimize_pulses(parallel_map=...); the time-step exchange runs "
        "inside the CUDA kernels")


def set_parallelization(use_loky=False, start_method=None, loky_pickler=None,
                        use_threadpool_limits=True):
    """Accepted for compatibility (parallelization.py:176-230); only
    `use_threadpool_limits` is recorded, nothing else applies to the GPU
    engine."""
    global USE_THREADPOOL_LIMITS
    USE_THREADPOOL_LIMITS = bool(use_threadpool_limits)
