"""CPU tests: host logic of the facade, the problem compiler, and the C-ABI
library's exported symbols (no compute calls; no GPU needed)."""
import ctypes
import os
import re

import numpy as np
import pytest

import krotov_b200 as krotov
from krotov_b200 import _lib
from krotov_b200.compiler import compile_problem, initialize_controls

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    """include/krotov_b200.h and the built .so agree."""
    _lib.build_library()
    header = open(os.path.join(ROOT, 'include', 'krotov_b200.h')).read()
    declared = set(re.findall(r'\b(kq_[a-z_]+)\s*\(', header))
    assert declared == set(_lib.EXPORTS)
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), name
    assert _lib.load().kq_version() >= 100


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    wl = krotov.workloads.tls_ensemble(K=2, nt=10)
    with pytest.raises(krotov.EngineUnavailable):
        krotov.optimize_pulses(
            wl.objectives(krotov.Objective), wl.pulse_options, wl.tlist,
            propagator=krotov.propagators.expm,
            chi_constructor=krotov.functionals.chis_re, iter_stop=1)


def test_custom_propagator_rejected():
    wl = krotov.workloads.tls_ensemble(K=2, nt=10)

    def my_prop(H, state, dt, c_ops=None, backwards=False, initialize=False):
        return state

    with pytest.raises(NotImplementedError):
        krotov.optimize_pulses(
            wl.objectives(krotov.Objective), wl.pulse_options, wl.tlist,
            propagator=my_prop, chi_constructor=krotov.functionals.chis_re,
            iter_stop=1)


def test_discretisation_inverse_property():
    """tests/test_structural_conversions.py:18-60 of the reference: pulse ->
    control -> pulse is the identity; discretize() argument checking."""
    from functools import partial
    from krotov_b200.conversions import discretize
    from krotov_b200.shapes import blackman, qutip_callback
    tlist = np.linspace(0, 10, 20)
    ctrl = qutip_callback(blackman, t_start=0, t_stop=10)
    pulse_orig = krotov.conversions.control_onto_interval(
        discretize(ctrl, tlist))
    control = krotov.conversions.pulse_onto_tlist(pulse_orig)
    pulse = krotov.conversions.control_onto_interval(control)
    assert np.max(np.abs(pulse - pulse_orig)) < 1e-14
    with pytest.raises(TypeError):
        discretize(partial(blackman, t_start=0, t_stop=10), tlist)
    with pytest.raises(TypeError):
        discretize('sin(t)', tlist)
    with pytest.raises(ValueError):
        discretize(np.array([ctrl(t, None) for t in tlist[:-1]]), tlist)
    arr = discretize(ctrl, tlist)
    assert len(arr) == len(tlist)
    assert abs(arr[0]) < 1e-15 and abs(arr[-1]) < 1e-15
    assert np.max(np.abs(np.array([ctrl(t, None) for t in tlist]) - arr)) \
        < 1e-15


def test_controls_mapping_doc_example():
    """conversions.py:186-236 of the reference."""
    X, Y, Z = np.eye(2), np.eye(2) * 2, np.eye(2) * 3
    u1, u2 = np.array([]), np.array([])
    psi = np.zeros((2, 1))
    krotov.Objective.type_checking = False
    try:
        c_ops = [[[X, u1]], [[Y, u2]]]
        objs = [
            krotov.Objective(initial_state=psi, target=psi,
                             H=[X, [Y, u1], [Z, u1]], c_ops=c_ops),
            krotov.Objective(initial_state=psi, target=psi,
                             H=[X, [Y, u2]], c_ops=c_ops),
        ]
        controls = krotov.conversions.extract_controls(objs)
        assert len(controls) == 2 and controls[0] is u1 and controls[1] is u2
        m = krotov.conversions.extract_controls_mapping(objs, controls)
        assert m == [[[[1, 2], []], [[0], []], [[], [0]]],
                     [[[], [1]], [[0], []], [[], [0]]]]
    finally:
        krotov.Objective.type_checking = True
    H = ['X', ['X', None], ['Y', None], ['Z', None]]
    pulses = [np.array([0, 10, 0]), np.array([0, 20, 0])]
    out = krotov.conversions.plug_in_pulse_values(H, pulses, [[1, 2], [3]], 1)
    assert out == ['X', ['X', 10], ['Y', 10], ['Z', 20]]


def test_pulse_options_validation():
    """tests/test_pulse_options.py:9-70 of the reference."""
    wl = krotov.workloads.tls_ensemble(K=1, nt=10)
    objs = wl.objectives(krotov.Objective)
    ctrl = list(wl.pulse_options)[0]
    with pytest.raises(ValueError, match="lambda_a"):
        initialize_controls(objs, {ctrl: dict(update_shape=1)}, wl.tlist)
    with pytest.raises(ValueError, match="update_shape"):
        initialize_controls(objs, {ctrl: dict(lambda_a=1)}, wl.tlist)
    with pytest.raises(ValueError, match=r"range \[0, 1\]"):
        initialize_controls(
            objs, {ctrl: dict(lambda_a=1, update_shape=lambda t: 2.0)},
            wl.tlist)
    with pytest.raises(ValueError, match="pulse options"):
        initialize_controls(objs, {}, wl.tlist)
    with pytest.raises(ValueError, match="real-valued"):
        initialize_controls(
            objs, {ctrl: dict(lambda_a=1, update_shape=lambda t: 1j)},
            wl.tlist)


def test_compile_problem_tables():
    """Term merging, mu table, adjoints, column-major layout."""
    rng = np.random.default_rng(3)
    N = 3
    mats = [rng.normal(size=(N, N)) + 1j * rng.normal(size=(N, N))
            for _ in range(4)]
    u1 = lambda t, a: 1.0  # noqa: E731
    u2 = lambda t, a: 2.0  # noqa: E731
    psi = np.ones((N, 1), dtype=complex)
    objs = [
        krotov.Objective(initial_state=psi, target=psi,
                         H=[mats[0], [mats[1], u1], [mats[2], u1],
                            [mats[3], u2]]),
        krotov.Objective(initial_state=psi, target=psi,
                         H=[mats[0], [mats[3], u2]]),
    ]
    tlist = np.linspace(0, 1, 5)
    opts = {u1: dict(lambda_a=1, update_shape=1),
            u2: dict(lambda_a=2, update_shape=1)}
    controls, _, pulses, mapping, lam, shp = initialize_controls(
        objs, opts, tlist)
    cp = compile_problem(objs, controls, mapping, tlist)
    assert (cp.K, cp.N, cp.L, cp.M, cp.NT) == (2, 3, 2, 3, 4)
    assert cp.term2pulse.tolist() == [[-1, 0, 1], [-1, 1, -2]]
    # column-major: stored[c, r] == op[r, c]
    assert np.allclose(cp.ops[0, 1].T, mats[1] + mats[2])
    assert np.allclose(cp.ops_adj[0, 1].T, (mats[1] + mats[2]).conj().T)
    assert np.allclose(cp.mu[0, 0].T, mats[1] + mats[2])
    assert np.allclose(cp.mu[1, 0], 0)
    assert np.allclose(cp.mu[1, 1].T, mats[3])
    assert np.allclose(cp.op_norm[0, 0], np.linalg.norm(mats[0], 1))
    assert np.allclose(lam, [1, 2])
    # custom mu given as callable is probed into the same table
    def my_mu(objectives, i_objective, pulses, pulses_mapping, i_pulse,
              time_index):
        op = krotov.mu.derivative_wrt_pulse(
            objectives, i_objective, pulses, pulses_mapping, i_pulse,
            time_index)
        if callable(op) and not isinstance(op, np.ndarray):
            return op
        return lambda state: op @ state
    cp2 = compile_problem(objs, controls, mapping, tlist, mu=my_mu,
                          pulses_for_mu=pulses)
    assert np.allclose(cp2.mu, cp.mu)


def test_liouvillian_matches_lindblad_rhs():
    rng = np.random.default_rng(5)
    d = 3
    H = rng.normal(size=(d, d)) + 1j * rng.normal(size=(d, d))
    H = H + H.conj().T
    C = rng.normal(size=(d, d)) + 1j * rng.normal(size=(d, d))
    rho = rng.normal(size=(d, d)) + 1j * rng.normal(size=(d, d))
    L = krotov.liouvillian(H, [C])
    rhs = -1j * (H @ rho - rho @ H) + C @ rho @ C.conj().T - 0.5 * (
        C.conj().T @ C @ rho + rho @ C.conj().T @ C)
    got = (L @ rho.reshape(-1, order='F')).reshape(d, d, order='F')
    assert np.allclose(got, rhs)
    nested = krotov.liouvillian([H, [H, None]], [C])
    assert np.allclose(nested[0], L) and nested[1][1] is None


def test_gate_and_ensemble_objectives():
    b = [np.eye(2)[:, [i]].astype(complex) for i in range(2)]
    X = np.array([[0, 1], [1, 0]], dtype=complex)
    H = [np.eye(2), [X, lambda t, a: 0.0]]
    objs = krotov.gate_objectives(b, X, H)
    assert len(objs) == 2
    assert objs[0].target is b[1] and objs[1].target is b[0]
    objs_w = krotov.gate_objectives(b, X, H, weights=[1, 3])
    assert [o.weight for o in objs_w] == [0.5, 1.5]
    with pytest.raises(ValueError):
        krotov.gate_objectives(b, np.eye(3), H)
    ens = krotov.ensemble_objectives(objs, [H, H])
    assert len(ens) == 6 and ens[2].H is H
    b4 = [np.eye(4)[:, [i]].astype(complex) for i in range(4)]
    pe = krotov.gate_objectives(b4, 'PE', [np.eye(4)])
    assert len(pe) == 4 and pe[0].target == 'PE'
    assert np.allclose(pe[1].initial_state,
                       1j * (b4[1] + b4[2]) / np.sqrt(2))
    rho = krotov.gate_objectives(b, X, H, liouville_states_set='3states')
    assert len(rho) == 3 and rho[0].initial_state.shape == (2, 2)
    adj = objs[0].adjoint()
    assert np.allclose(adj.H[1][0], X.conj().T) and adj.H[1][1] is H[1][1]


def test_qobj_duck_typing_through_compiler():
    """Qobj-shaped inputs (the shim's dense Qobj) compile to the same tables
    as numpy inputs."""
    import sys
    sys.path.insert(0, os.path.join(ROOT, 'oracle', 'ref_shims'))
    try:
        import qutip
    finally:
        sys.path.pop(0)
    wl = krotov.workloads.transmon_xgate(nstates=2, nt=20)
    o_np = wl.objectives(krotov.Objective)
    o_q = wl.objectives(krotov.Objective, wrap=lambda a: qutip.Qobj(a))
    outs = []
    for objs in (o_np, o_q):
        controls, _, _, mapping, _, _ = initialize_controls(
            objs, wl.pulse_options, wl.tlist)
        outs.append(compile_problem(objs, controls, mapping, wl.tlist))
    assert np.array_equal(outs[0].ops, outs[1].ops)
    assert np.array_equal(outs[0].psi0, outs[1].psi0)
    back = outs[1].unvec(outs[1].psi0[0], o_q[0].initial_state)
    assert isinstance(back, qutip.Qobj) and back.type == 'ket'


def test_result_dump_load_roundtrip(tmp_path):
    r = krotov.Result()
    wl = krotov.workloads.tls_ensemble(K=1, nt=10)
    r.objectives = wl.objectives(krotov.Objective)
    r.tlist = wl.tlist
    r.optimized_controls = [np.arange(9.0)]
    r.guess_controls = [np.arange(10.0)]
    r.iters = [0, 1]
    r.message = 'x'
    fn = str(tmp_path / 'r.dump')
    r.dump(fn)
    r2 = krotov.Result.load(fn, finalize=True)
    assert len(r2.optimized_controls[0]) == 10
    assert r2.objectives[0].H[1][1] is None   # lambda control dropped
    r3 = krotov.Result.load(fn, objectives=r.objectives)
    assert r3.objectives is r.objectives


def test_fused_iteration_plan_geometry():
    """Launch geometry of the one-launch iteration kernel (host-only query):
    one objective per CTA up to 148, power-of-two chunk lengths, and a clean
    refusal outside the kernel family."""
    import ctypes
    from krotov_b200 import _lib
    lib = _lib.load()

    def plan(K, N, nt, L=1, M=2):
        p = _lib.KqProblem(K=K, N=N, NT=nt - 1, L=L, M=M, is_super=0)
        vals = [ctypes.c_int32() for _ in range(4)]
        rc = lib.kq_plan_fused(ctypes.byref(p), *[ctypes.byref(v) for v in vals])
        return rc, tuple(v.value for v in vals)

    rc, (grid, block, chunk, smem) = plan(128, 2, 1000)      # C4
    assert rc == 0 and (grid, block, chunk) == (128, 256, 4)
    assert smem < 100 * 1024
    rc, (grid, block, chunk, _) = plan(1, 2, 500)            # C1
    assert rc == 0 and (grid, block, chunk) == (1, 256, 2)
    rc, (grid, block, chunk, _) = plan(4, 4, 2000)           # C3
    assert rc == 0 and (grid, block, chunk) == (4, 256, 8)
    rc, (grid, block, chunk, _) = plan(300, 2, 1000)         # 3 objectives per CTA
    assert rc == 0 and grid == 100 and block == 192 and chunk == 16
    assert plan(128, 2, 5000)[0] == -3     # state stores exceed shared memory
    assert plan(128, 5, 1000)[0] == -3     # N > 4: lane-per-row kernels
    assert plan(8, 2, 100, L=2, M=3)[0] == -3
    assert plan(2000, 2, 1000)[0] == -3    # more than 8 objectives per CTA


def test_reference_process_maps_are_accepted_names():
    """parallelization.py:233-311 of the reference: the process maps exist
    under the same names with the serial_map interface."""
    from krotov_b200 import parallelization as par
    assert par.serial_map(lambda v, a, b=0: v + a + b, [1, 2], (10,),
                          {'b': 5}) == [16, 17]
    assert par.parallel_map(lambda v: 2 * v, [1, 2], num_cpus=4) == [2, 4]
    par.set_parallelization(use_threadpool_limits=False)
    assert par.USE_THREADPOOL_LIMITS is False
    par.set_parallelization()
    with pytest.raises(NotImplementedError):
        par.parallel_map_fw_prop_step(None, [0], ())


def test_bench_cpu_legs_on_a_small_ensemble(monkeypatch, tmp_path):
    """bench.py's CPU legs (serial numpy port, multi-process port with the
    pulse dump used for the same-run parity key) on a 4-objective ensemble."""
    import numpy as np
    import bench
    monkeypatch.setitem(bench.WORKLOAD, 'K', 4)
    monkeypatch.setitem(bench.WORKLOAD, 'nt', 60)
    wl = bench.build_workload()
    per_iter, ks = bench.time_oracle(wl, 1)
    assert per_iter > 0 and ks == 4
    dump = str(tmp_path / 'pulses.npy')
    leg = bench.parallel_leg(wl, 1, 1, dump=dump)
    assert leg['kind'] == 'port' and 1 <= leg['cores'] <= 4
    assert leg['value'] > 0 and leg['unit'] == bench.UNIT
    assert np.load(dump).shape == (2, 1, 59)
    # replica r > 0 is a different optimisation problem (guess amplitude)
    p0 = np.array(bench.build_workload().lowered()['pulses'])
    p1 = np.array(bench.build_workload(replica=1).lowered()['pulses'])
    assert np.allclose(p1, 1.05 * p0)


def test_convergence_specs_follow_reference():
    """convergence.py:109-297 of the reference: glom-style specs (default
    ('info_vals', T[-1])), error propagation, Or()."""
    from krotov_b200 import convergence as cv
    from krotov_b200.result import Result
    r = Result()
    chk = cv.value_below('1e-4', spec=lambda res: res.info_vals[-1],
                         name='J_T')
    r.info_vals.append(1e-4)
    assert chk(r) is None
    r.info_vals.append(9e-5)
    assert chk(r) == 'J_T < 1e-4'
    # default spec, tuple entries are NOT unwrapped
    assert cv.value_below(1e-3)(r) == "%s < 0.001" % (
        str(('info_vals', cv.T[-1])),)
    # a spec that selects another attribute is honoured, not ignored
    r.tau_vals.append(np.array([0.5 + 0j]))
    chk = cv.value_above('0.4', spec=('tau_vals', cv.T[-1], cv.T[0],
                                      lambda z: z.real), name='F')
    assert chk(r) == 'F > 0.4'
    assert cv.value_below('0.4', spec=('tau_vals', cv.T[-1], cv.T[0],
                                       abs))(r) is None
    assert cv.extract(r, 'info_vals')[-1] == 9e-5
    # no info_hook -> empty info_vals -> the reference raises IndexError
    with pytest.raises(IndexError):
        cv.value_below(1e-3)(Result())
    with pytest.raises(TypeError):
        cv.value_below(1e-3, spec={'a': 'info_vals'})(r)
    # delta_below: one missing value passes, two missing values re-raise
    d = cv.delta_below('1e-4', name='ΔJ_T')
    r2 = Result()
    with pytest.raises(IndexError):
        d(r2)
    for v, want in [(9e-1, None), (1e-1, None), (4e-4, None), (2e-4, None),
                    (1e-6, None), (1e-7, 'ΔJ_T < 1e-4')]:
        r2.info_vals.append(v)
        assert d(r2) == want
    r3 = Result()
    for v, want in [(9e-1, None), (1e-1, None), (2e-1, (
            'Loss of monotonic convergence; error decrease < 0'))]:
        r3.info_vals.append(v)
        assert cv.check_monotonic_error(r3) == want
    assert cv.check_monotonic_fidelity(r3) is None
    assert cv.Or(cv.check_monotonic_fidelity, cv.check_monotonic_error)(
        r3).startswith('Loss of monotonic')
    with pytest.raises(ValueError):
        cv.dump_result('x', every=0)


def test_import_krotov_alias_package():
    """`import krotov` resolves to the engine when compat/ is on the path: the
    reference's package, submodule and top-level names
    (/root/reference/src/krotov/__init__.py:40-65)."""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = (
        "import krotov, krotov_b200\n"
        "import krotov.propagators as P\n"
        "from krotov.functionals import chis_re, J_T_re\n"
        "from krotov.objectives import Objective, gate_objectives\n"
        "from krotov.convergence import check_monotonic_error, value_below\n"
        "from krotov.info_hooks import print_table, chain\n"
        "from krotov.shapes import flattop, blackman\n"
        "from krotov.second_order import numerical_estimate_A\n"
        "from krotov.conversions import pulse_onto_tlist\n"
        "from krotov.parallelization import parallel_map\n"
        "assert krotov.optimize_pulses is krotov_b200.optimize_pulses\n"
        "assert krotov.Objective is krotov_b200.Objective\n"
        "assert krotov.result.Result is krotov.Result\n"
        "assert P.expm is krotov_b200.propagators.expm\n"
        "assert callable(krotov.mu.derivative_wrt_pulse)\n"
        "print('ok')\n")
    env = dict(os.environ)
    env['PYTHONPATH'] = os.pathsep.join(
        [os.path.join(root, 'compat'), root, env.get('PYTHONPATH', '')])
    out = subprocess.run([sys.executable, '-c', code], env=env,
                         capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr
    assert out.stdout.strip() == 'ok'


def test_print_table_api_matches_reference():
    """print_table: keyword-only signature with J_T_prev, validation of
    col_formats / col_headers (/root/reference/tests/test_infohooks.py:78-127),
    default and custom layouts (rows of tests/test_infohooks/
    custom_format_out.txt and of the SciPost example in the docstring,
    info_hooks.py:366-372), '*' markers, J_T_prev from info_vals."""
    import io
    from krotov_b200.info_hooks import print_table
    J_T = lambda **kwargs: 1.0  # noqa: E731
    for bad, msg in (((".2e"), '8 elements'), ("eeeeeeee", 'percent format string'),
                     ((".2e", ".2e"), '8 elements'),
                     (('%d', '%.2', '%.2e', '%.2e', '%.2e', '%.2e', '%.2e', '%d'),
                      'Invalid col_formats')):
        with pytest.raises(ValueError) as exc:
            print_table(J_T=J_T, col_formats=bad)
        assert msg in str(exc.value)
    for bad in ("header", ("header", "header")):
        with pytest.raises(ValueError) as exc:
            print_table(J_T=J_T, col_headers=bad)
        assert '8 elements' in str(exc.value)
    with pytest.raises(ValueError) as exc:
        print_table(J_T=J_T, show_g_a_int_per_pulse=True, col_headers=(
            "it", "J_T", "∫gₐ(ϵ{i})dt", "∑∫gₐ(t)dt", "J", "ΔJ_T", "ΔJ", "secs"))
    assert "must support '.format(l=l)'" in str(exc.value)
    with pytest.raises(TypeError):
        print_table(J_T)          # keyword-only, like the reference

    def rows(hook, vals, g_as, n_pulses=1, info_vals=None):
        info_vals = [] if info_vals is None else info_vals
        for it, (v, g) in enumerate(zip(vals, g_as)):
            info_vals.append(hook(
                iteration=it, g_a_integrals=np.atleast_1d(g),
                guess_pulses=[None] * n_pulses, iter_stop=10, start_time=0.0,
                stop_time=float(it), info_vals=info_vals, value=v))
    out = io.StringIO()
    hook = print_table(J_T=lambda **kw: kw['value'], out=out)
    rows(hook, [1.0, 0.765, 0.556], [0.0, 2.33e-2, 2.07e-2])
    assert out.getvalue().splitlines() == [
        "iter.      J_T    ∫gₐ(t)dt          J       ΔJ_T         ΔJ  secs",
        "0     1.00e+00    0.00e+00   1.00e+00        n/a        n/a     0",
        "1     7.65e-01    2.33e-02   7.88e-01  -2.35e-01  -2.12e-01     1",
        "2     5.56e-01    2.07e-02   5.77e-01  -2.09e-01  -1.88e-01     2"]
    out = io.StringIO()
    hook = print_table(
        J_T=lambda **kw: kw['value'], out=out,
        col_formats=('%8d', '%12.4e', '%12.4e', '%12.4e', '%12.4e', '%12.4e', '%12.4e',
                     '%05d'))
    rows(hook, [1.0, 0.76485], [0.0, 0.11758])
    assert out.getvalue().splitlines() == [
        "iter.              J_T     ∫gₐ(t)dt            J         ΔJ_T"
        "           ΔJ  secs",
        "       0    1.0000e+00   0.0000e+00   1.0000e+00          n/a"
        "          n/a 00000",
        "       1    7.6485e-01   1.1758e-01   8.8243e-01  -2.3515e-01"
        "  -1.1757e-01 00001"]
    # two pulses, per-pulse columns, loss of monotonic convergence marked
    out = io.StringIO()
    hook = print_table(J_T=lambda **kw: kw['value'], out=out,
                       show_g_a_int_per_pulse=True, unicode=False)
    rows(hook, [1.0, 1.1], [[0.0, 0.0], [1e-2, 2e-2]], n_pulses=2)
    lines = out.getvalue().splitlines()
    assert lines[0].split() == ["iter.", "J_T", "g_a_int_1", "g_a_int_2",
                                "g_a_int", "J", "Delta", "J_T", "Delta", "J",
                                "secs"]
    assert lines[2].endswith(" **")
    assert lines[2].split()[:5] == ['1', '1.10e+00', '1.00e-02', '2.00e-02',
                                    '3.00e-02']
    # custom headers (tests/test_infohooks/custom_header_out.txt of the reference)
    out = io.StringIO()
    hook = print_table(
        J_T=lambda **kw: kw['value'], out=out, show_g_a_int_per_pulse=True,
        col_headers=('iteration', ' final time functional',
                     ' running cost (pulse {l})', ' total running cost',
                     ' total functional', ' change in final time functional',
                     ' change in total functional', ' seconds for iteration'))
    rows(hook, [1.0, 0.76485], [0.0, 0.11758])
    assert out.getvalue().splitlines() == [
        "iteration   final time functional  total running cost  total functional"
        "  change in final time functional  change in total functional"
        "  seconds for iteration",
        "0                        1.00e+00            0.00e+00          1.00e+00"
        "                              n/a                         n/a"
        "                      0",
        "1                        7.65e-01            1.18e-01          8.82e-01"
        "                        -2.35e-01                   -1.18e-01"
        "                      1"]
    # continuation: the first printed row uses info_vals[-1] (or J_T_prev)
    out = io.StringIO()
    hook = print_table(J_T=lambda **kw: 0.5, out=out)
    hook(iteration=4, g_a_integrals=np.array([0.1]), guess_pulses=[None],
         iter_stop=10, start_time=0.0, stop_time=0.0, info_vals=[0.75])
    assert "-2.50e-01  -1.50e-01" in out.getvalue()


def test_perfect_entangler_functional():
    """krotov_b200.perfect_entanglers (restatement of what notebook 07 takes
    from the absent `weylchamber`): local invariants of known gates, F_PE of
    notebook 07's guess (cell 39: 1.447335), Wirtinger gradient against
    central differences, chi constructor = -dF/d<phi| in the Bell basis of
    gate_objectives(..., 'PE')."""
    from krotov_b200 import perfect_entanglers as pe
    import krotov_b200 as krotov
    cnot = np.array([[1, 0, 0, 0], [0, 1, 0, 0], [0, 0, 0, 1], [0, 0, 1, 0]])
    swap = np.array([[1, 0, 0, 0], [0, 0, 1, 0], [0, 1, 0, 0], [0, 0, 0, 1]])
    for U, want in ((np.eye(4), (1, 0, 3)), (cnot, (0, 0, 1)),
                    (swap, (-1, 0, -3)),
                    (np.exp(0.7j) * cnot, (0, 0, 1))):    # any global phase
        assert np.allclose(pe.g1g2g3(U), want, atol=1e-14)
    assert abs(pe.F_PE(*pe.g1g2g3(cnot))) < 1e-14
    rng = np.random.default_rng(3)
    q, _ = np.linalg.qr(rng.normal(size=(4, 4)) + 1j * rng.normal(size=(4, 4)))
    a = q + 0.05 * (rng.normal(size=(4, 4)) + 1j * rng.normal(size=(4, 4)))
    F, dFc = pe.F_PE_gradient(a)

    def F_of(b):
        return pe.F_PE_gradient(b)[0]
    h = 1e-6
    for k in range(4):
        for l in range(4):
            e = np.zeros((4, 4), dtype=complex)
            e[k, l] = 1
            num = 0.5 * ((F_of(a + h * e) - F_of(a - h * e)) / (2 * h)
                         + 1j * (F_of(a + 1j * h * e) - F_of(a - 1j * h * e))
                         / (2 * h))
            assert abs(num - dFc[k, l]) < 1e-9
    # Bell basis of the objectives = columns of MAGIC; chi = -dF/d<phi|
    basis = [np.eye(4, dtype=complex)[:, [i]] for i in range(4)]
    objs = krotov.gate_objectives(basis, 'PE', [np.eye(4)])
    bell = np.array([o.initial_state.ravel() for o in objs])
    assert np.allclose(bell, pe.MAGIC.T)
    states = [(q @ b.reshape(4, 1)) for b in bell]
    chis = pe.make_PE_krotov_chi_constructor(basis)(states, objs, None)
    a_q = bell.conj() @ np.array([s.ravel() for s in states]).T
    _, d = pe.F_PE_gradient(a_q)
    for l in range(4):
        assert chis[l].shape == (4, 1)
        assert np.allclose(chis[l].ravel(), -(d[:, l] @ bell))


def test_row_nnz_of_compiled_problems():
    """kq_problem.row_nnz (largest number of non-zero columns in a row of an
    objective's terms taken together, diagonal included) selects the
    entries-in-registers kernels (csrc/kq_lanes.cuh) for sparse rows."""
    import krotov_b200 as krotov
    from krotov_b200.compiler import compile_problem, initialize_controls

    def row_nnz(wl):
        objs = wl.objectives(krotov.Objective)
        controls, _, _, mapping, _, _ = initialize_controls(
            objs, wl.pulse_options, wl.tlist)
        return compile_problem(objs, controls, mapping, wl.tlist).row_nnz

    W = krotov.workloads
    # ladder: diagonal drift + tridiagonal drive
    assert row_nnz(W.transmon_xgate(nstates=8, nt=20)) == 3
    # Lambda system: 1-2 and 2-3 couplings, diagonal detunings
    assert row_nnz(W.lambda_system(nt=20, gamma=0.5)) == 3
    # two-level systems are dense
    assert row_nnz(W.tls_ensemble(K=3, nt=20)) == 2
    # the two-qubit gate Hamiltonian of C3 couples every level to three others
    assert 2 <= row_nnz(W.two_qubit_gate(nt=20)) <= 4


def test_kq_plan_reports_the_kernel_families():
    """kq_plan is host-only introspection: which family the update sweep of a
    problem runs on (include/krotov_b200.h)."""
    import ctypes
    from krotov_b200 import _lib
    lib = _lib.load()

    def family(K, N, NT, L, M, real=0, nnz=0):
        p = _lib.KqProblem(K=K, N=N, NT=NT, L=L, M=M, is_super=0, ops=1,
                           ops_adj=1, mu=1, term2pulse=1, op_norm=1, dt=1,
                           shape=1, lambda_a=1, real_ops=real, reserved=0,
                           update_sweep=0, row_nnz=nnz, sparse=None)
        v = [ctypes.c_int32() for _ in range(4)]
        assert lib.kq_plan(ctypes.byref(p),
                           *[ctypes.byref(x) for x in v]) == 0
        return v[0].value, v[1].value, v[2].value

    assert family(128, 2, 999, 1, 2, real=1)[0] == 0          # C4
    assert family(131072, 2, 999, 1, 2, real=1) == (4, 148, 512)   # kq_sat
    assert family(5, 3, 499, 4, 5, nnz=3) == (3, 1, 32)       # Lambda ensemble
    assert family(2, 17, 999, 1, 2, real=1, nnz=3) == (3, 1, 64)   # transmon
    assert family(2, 17, 999, 1, 2, real=1, nnz=0)[0] == 1    # dense rows
    assert family(300, 3, 499, 4, 5, nnz=3)[0] == 0           # too many lanes
