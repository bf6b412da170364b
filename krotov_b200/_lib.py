"""ctypes binding of ``libkrotov_b200.so`` (C ABI in ``include/krotov_b200.h``).

The shared library is built in-tree by ``__graft_entry__.build()`` (plain
``nvcc -shared`` for sm_100a, no torch headers).  There is NO fallback: if the
library is missing or a CUDA device is unavailable the engine raises
:class:`EngineUnavailable` -- the sweeps never run on the CPU.
"""
import ctypes
import os
import subprocess

__all__ = ['load', 'EngineUnavailable', 'KqError', 'KqProblem', 'KqComm',
           'LIB_PATH', 'build_library', 'check']

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
LIB_PATH = os.path.join(CSRC, 'libkrotov_b200.so')
INCLUDE = os.path.join(os.path.dirname(HERE), 'include')

COMPILE_FLAGS = [
    '-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo', '-O3',
    '-std=c++17', '-Xcompiler', '-fPIC',
]
LINK_FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-shared',
              '-Xcompiler', '-fPIC']


class EngineUnavailable(RuntimeError):
    """The CUDA library or device required by the sweep engine is missing."""


class KqError(RuntimeError):
    """An ABI call returned a negative status (``.status``: the kq_status
    value, e.g. -3 = KQ_ERR_UNSUPPORTED)."""

    def __init__(self, message, status=None):
        super().__init__(message)
        self.status = status


class KqProblem(ctypes.Structure):
    """Mirror of ``kq_problem`` (include/krotov_b200.h)."""
    _fields_ = [
        ('K', ctypes.c_int32), ('N', ctypes.c_int32), ('NT', ctypes.c_int32),
        ('L', ctypes.c_int32), ('M', ctypes.c_int32),
        ('is_super', ctypes.c_int32),
        ('ops', ctypes.c_void_p), ('ops_adj', ctypes.c_void_p),
        ('mu', ctypes.c_void_p), ('term2pulse', ctypes.c_void_p),
        ('op_norm', ctypes.c_void_p), ('dt', ctypes.c_void_p),
        ('shape', ctypes.c_void_p), ('lambda_a', ctypes.c_void_p),
        ('real_ops', ctypes.c_int32), ('reserved', ctypes.c_int32),
        ('update_sweep', ctypes.c_int32), ('row_nnz', ctypes.c_int32),
        ('sparse', ctypes.c_void_p),
    ]


class KqSparse(ctypes.Structure):
    """Mirror of ``kq_sparse``."""
    _fields_ = [('row_ptr', ctypes.c_void_p), ('mat_off', ctypes.c_void_p),
                ('col', ctypes.c_void_p), ('val', ctypes.c_void_p),
                ('col16', ctypes.c_void_p), ('code16', ctypes.c_void_p),
                ('dict', ctypes.c_void_p), ('n_dict', ctypes.c_int32),
                ('stage_nnz_update', ctypes.c_int32),
                ('stage_nnz_prop', ctypes.c_int32)]


class KqComm(ctypes.Structure):
    """Mirror of ``kq_comm``."""
    _fields_ = [('rank', ctypes.c_int32), ('world', ctypes.c_int32),
                ('slots', ctypes.c_void_p)]


_SIGNATURES = {
    'kq_version': (ctypes.c_int, []),
    'kq_last_error': (ctypes.c_char_p, []),
    'kq_set_option': (ctypes.c_int, [ctypes.c_char_p, ctypes.c_int]),
    'kq_comm_alloc': (ctypes.c_int, [ctypes.c_size_t,
                                     ctypes.POINTER(ctypes.c_void_p),
                                     ctypes.c_char_p]),
    'kq_comm_open': (ctypes.c_int, [ctypes.c_char_p,
                                    ctypes.POINTER(ctypes.c_void_p)]),
    'kq_comm_close': (ctypes.c_int, [ctypes.c_void_p]),
    'kq_comm_free': (ctypes.c_int, [ctypes.c_void_p]),
    'kq_comm_barrier': (ctypes.c_int, [ctypes.POINTER(KqComm),
                                       ctypes.c_uint32, ctypes.c_void_p,
                                       ctypes.c_void_p]),
    'kq_workspace_bytes': (ctypes.c_size_t, [ctypes.POINTER(KqProblem)]),
    'kq_dpoly_header_offset': (ctypes.c_size_t, [ctypes.POINTER(KqProblem)]),
    'kq_comm_slot_bytes': (ctypes.c_size_t, [ctypes.POINTER(KqProblem)]),
    'kq_propagate_forward': (ctypes.c_int, [
        ctypes.POINTER(KqProblem), ctypes.c_void_p, ctypes.c_void_p,
        ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]),
    'kq_sweep_backward': (ctypes.c_int, [
        ctypes.POINTER(KqProblem), ctypes.c_void_p, ctypes.c_void_p,
        ctypes.c_void_p, ctypes.c_void_p]),
    'kq_sweep_backward_range': (ctypes.c_int, [
        ctypes.POINTER(KqProblem), ctypes.c_void_p, ctypes.c_void_p,
        ctypes.POINTER(ctypes.c_void_p), ctypes.c_int32, ctypes.c_int32,
        ctypes.c_int32, ctypes.c_int32, ctypes.c_void_p]),
    'kq_sweep_forward_update': (ctypes.c_int, [
        ctypes.POINTER(KqProblem), ctypes.c_void_p, ctypes.c_void_p,
        ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
        ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
        ctypes.POINTER(KqComm), ctypes.c_void_p, ctypes.c_uint32,
        ctypes.c_void_p]),
    'kq_krotov_iteration': (ctypes.c_int, [
        ctypes.POINTER(KqProblem), ctypes.c_int, ctypes.c_int32] +
        [ctypes.c_void_p] * 20 + [ctypes.POINTER(KqComm), ctypes.c_void_p,
                                ctypes.c_void_p, ctypes.c_uint32,
                                ctypes.c_void_p]),
    'kq_chi_boundary': (ctypes.c_int, [
        ctypes.POINTER(KqProblem), ctypes.c_int, ctypes.c_int32,
        ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
        ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]),
    'kq_fetch_results': (ctypes.c_int, [
        ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p]),
    'kq_overlaps': (ctypes.c_int, [
        ctypes.c_int32, ctypes.c_int32, ctypes.c_void_p, ctypes.c_void_p,
        ctypes.c_void_p, ctypes.c_void_p]),
    'kq_plan_fused': (ctypes.c_int, [
        ctypes.POINTER(KqProblem), ctypes.POINTER(ctypes.c_int32),
        ctypes.POINTER(ctypes.c_int32), ctypes.POINTER(ctypes.c_int32),
        ctypes.POINTER(ctypes.c_int32)]),
    'kq_plan': (ctypes.c_int, [
        ctypes.POINTER(KqProblem), ctypes.POINTER(ctypes.c_int32),
        ctypes.POINTER(ctypes.c_int32), ctypes.POINTER(ctypes.c_int32),
        ctypes.POINTER(ctypes.c_int32)]),
}

EXPORTS = tuple(_SIGNATURES)

_lib = None


def build_library(verbose=False, jobs=None):
    """Compile every ``csrc/*.cu`` translation unit for sm_100a (in parallel)
    and link them into ``csrc/libkrotov_b200.so``.  nvcc cross-compiles
    without a GPU."""
    from concurrent.futures import ThreadPoolExecutor
    sources = sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC)
                     if f.endswith('.cu'))
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC)
               if f.endswith(('.cuh', '.inc'))]
    headers.append(os.path.join(INCLUDE, 'krotov_b200.h'))
    newest_hdr = max(os.path.getmtime(h) for h in headers)
    objdir = os.path.join(CSRC, 'build')
    os.makedirs(objdir, exist_ok=True)
    nvcc = os.environ.get('NVCC', '/usr/local/cuda/bin/nvcc')
    todo, objects = [], []
    for src in sources:
        obj = os.path.join(objdir, os.path.basename(src)[:-3] + '.o')
        objects.append(obj)
        if not os.path.exists(obj) or os.path.getmtime(obj) < max(
                os.path.getmtime(src), newest_hdr):
            todo.append((src, obj))

    def compile_one(job):
        src, obj = job
        cmd = [nvcc] + COMPILE_FLAGS + ['-c', '-o', obj, src]
        if verbose:
            print(' '.join(cmd), flush=True)
        subprocess.run(cmd, check=True)

    if todo:
        with ThreadPoolExecutor(jobs or min(len(todo), os.cpu_count() or 4)) \
                as pool:
            list(pool.map(compile_one, todo))
    if todo or not os.path.exists(LIB_PATH) or any(
            os.path.getmtime(LIB_PATH) < os.path.getmtime(o) for o in objects):
        cmd = [nvcc] + LINK_FLAGS + ['-o', LIB_PATH] + objects
        if verbose:
            print(' '.join(cmd), flush=True)
        subprocess.run(cmd, check=True)
    return LIB_PATH


def load():
    """Load the library (once) and declare every entry point's signature.

    Raises:
        EngineUnavailable: the shared library has not been built.
    """
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise EngineUnavailable(
            "%s not found: build it with `python -c 'import __graft_entry__ "
            "as g; g.build()'` (needs nvcc). krotov_b200 has no CPU fallback."
            % LIB_PATH)
    try:
        lib = ctypes.CDLL(LIB_PATH)
    except OSError as exc:
        raise EngineUnavailable("cannot load %s: %s" % (LIB_PATH, exc))
    for name, (restype, argtypes) in _SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = restype
        fn.argtypes = argtypes
    _lib = lib
    return lib


def check(status):
    """Raise :class:`KqError` with the library's message if `status` < 0."""
    if status != 0:
        msg = load().kq_last_error()
        raise KqError("libkrotov_b200 error %d: %s"
                      % (status, msg.decode() if msg else '?'), status)
