"""Debug aid: the delta-polynomial update sweep alone (no sequential kernel
behind it) against the sequential kernels, through the 4-call C ABI."""
import sys, os
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import krotov_b200 as krotov
from krotov_b200.compiler import compile_problem, initialize_controls
from krotov_b200.engine import SweepEngine

lib = krotov._lib.load()
name = sys.argv[1] if len(sys.argv) > 1 else 'C2'
wl = {'C2': lambda: krotov.workloads.transmon_xgate(),
      'C3': lambda: krotov.workloads.two_qubit_gate(nt=500),
      'C5': lambda: krotov.workloads.dissipative_qubit_reset(nt=500),
      'T5': lambda: krotov.workloads.transmon_xgate(nstates=2, nt=100)}[name]()
objectives = wl.objectives(krotov.Objective)
controls, _, guess, mapping, lam, shp = initialize_controls(
    objectives, wl.pulse_options, wl.tlist)
cp = compile_problem(objectives, controls, mapping, wl.tlist)


def run(dpoly, debug, iters=3):
    lib.kq_set_option(b"dpoly", dpoly)
    lib.kq_set_option(b"dpoly_debug", debug)
    lib.kq_set_option(b"picard", 0)
    eng = SweepEngine(cp, shp, lam)
    g = eng.pulses_to_device(guess)
    o = g.clone()
    phiT = eng.propagate_forward(g)
    tau = eng.overlaps(eng.t_targets, phiT)
    out = []
    for it in range(iters):
        if wl.chi == 'qubit_reset':
            eng.chi_from_host([cp.vec(wl.meta['chi_fixed'])] * cp.K)
        else:
            eng.chi_builtin(wl.chi, phiT, tau)
        eng.sweep_backward(g)
        phiT = eng.sweep_forward_update(g, o, phiT=eng.new_states())
        tau = eng.overlaps(eng.t_targets, phiT)
        torch.cuda.synchronize()
        nb = lib.kq_workspace_bytes(eng._p)
        hdr = eng.workspace[nb - 64:nb].cpu().numpy()
        st = eng.workspace[:16].view(torch.int32).cpu().numpy()
        out.append((o.cpu().numpy().copy(), phiT.cpu().numpy().copy(),
                    hdr[:16].view(np.int32).copy(), hdr[16:32].view(np.float64).copy(),
                    st.copy(), eng.epoch, eng.g_a.cpu().numpy().copy()))
        g, o = o, g
    return out


ref = run(0, 0)
got = run(2, 1)
for it, (a, b) in enumerate(zip(ref, got)):
    err = np.max(np.abs(a[0] - b[0])) / np.max(np.abs(a[0]))
    upd = np.max(np.abs(a[0] - (ref[it - 1][0] if it else np.array(guess))))
    first_bad = np.argmax(np.abs(a[0] - b[0]) > 1e-9 * np.max(np.abs(a[0])))
    print("it %d: rel pulse err %.3e  (true max update %.3e)  phiT err %.3e  "
          "hdr J=%d valid_epoch=%d bound=%.3e last_max=%.3e  status=%s epoch=%d "
          "first bad n=%d  g_a %s vs %s"
          % (it + 1, err, upd, np.max(np.abs(a[1] - b[1])), b[2][0], b[2][1],
             b[3][0], b[3][1], b[4], b[5], first_bad, a[6], b[6]))
    if it == 0:
        print("  ref[:6]", a[0][0, :6])
        print("  got[:6]", b[0][0, :6])


def run_composite(iters=6):
    """kq_krotov_iteration (composite path), sequential kernels queued."""
    lib.kq_set_option(b"dpoly", 1)
    lib.kq_set_option(b"dpoly_debug", 0)
    lib.kq_set_option(b"picard", 1)
    eng = SweepEngine(cp, shp, lam)
    g = eng.pulses_to_device(guess)
    o = g.clone()
    phiT = eng.propagate_forward(g)
    tau = eng.overlaps(eng.t_targets, phiT)
    sp, st_ = eng.new_states(), torch.empty_like(tau)
    fixed = wl.chi == 'qubit_reset'
    if fixed:
        eng.chi_from_host([cp.vec(wl.meta['chi_fixed'])] * cp.K)
    for it in range(iters):
        e0 = torch.cuda.Event(enable_timing=True)
        e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        eng.krotov_iteration(None if fixed else wl.chi, g, o, phiT, tau, sp, st_)
        e1.record()
        torch.cuda.synchronize()
        nb = lib.kq_workspace_bytes(eng._p)
        hdr = eng.workspace[nb - 64:nb].cpu().numpy()
        st = eng.workspace[:16].view(torch.int32).cpu().numpy()
        err = float('nan')
        if it < len(ref):
            err = np.max(np.abs(ref[it][0] - o.cpu().numpy())) / np.max(np.abs(ref[it][0]))
        print("composite it %d: %.3f ms  rel err vs sequential %.2e  J=%d valid_epoch=%d "
              "bound=%.3e last_max=%.3e status=%s epoch=%d"
              % (it + 1, e0.elapsed_time(e1), err, hdr[:16].view(np.int32)[0],
                 hdr[:16].view(np.int32)[1], hdr[16:32].view(np.float64)[0],
                 hdr[16:32].view(np.float64)[1], st, eng.epoch))
        phiT, sp = sp, phiT
        tau, st_ = st_, tau
        g, o = o, g


run_composite()
