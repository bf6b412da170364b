"""Per-iteration hooks (host callbacks, once per Krotov iteration).

``chain`` and ``print_table`` follow the calling convention of
``krotov.info_hooks`` (/root/reference/src/krotov/info_hooks.py:24-56,
352-641): every hook receives the keyword arguments listed at
info_hooks.py:59-86; the state stores are lazy sequences backed by device
tensors (downloaded only if a hook touches them).
"""
import sys
import time

import numpy as np

__all__ = ['chain', 'print_debug_information', 'print_table']


def chain(*hooks):
    """Call `hooks` in order with shared kwargs; return the tuple of their
    non-None results (a single result unwrapped, None if all are None)."""

    def info_hook(**kwargs):
        results = []
        for hook in hooks:
            res = hook(**kwargs)
            if res is not None:
                results.append(res)
        if not results:
            return None
        return results[0] if len(results) == 1 else tuple(results)

    return info_hook


def print_debug_information(*, objectives, adjoint_objectives,
                            backward_states, forward_states, forward_states0,
                            guess_pulses, optimized_pulses, g_a_integrals,
                            lambda_vals, shape_arrays, fw_states_T, tlist,
                            tau_vals, start_time, stop_time, iteration,
                            info_vals, shared_data, propagator,
                            chi_constructor, mu, sigma, iter_start, iter_stop,
                            out=sys.stdout):
    """Print a block of diagnostics for one iteration (abridged version of
    info_hooks.py:59-293)."""
    pr = lambda *a: print(*a, file=out)  # noqa: E731
    pr("Iteration %d" % iteration)
    pr("    duration: %.1f secs (started at %s)" % (
        stop_time - start_time,
        time.strftime('%Y-%m-%d %H:%M:%S', time.localtime(start_time))))
    pr("    number of objectives: %d" % len(objectives))
    for l, (g, o) in enumerate(zip(guess_pulses, optimized_pulses)):
        pr("    pulse %d: guess amplitude [%.2e, %.2e], optimized [%.2e, %.2e]"
           % (l, np.min(g), np.max(g), np.min(o), np.max(o)))
        pr("    lambda_a[%d] = %.2e, ∫gₐ(t)dt = %.2e"
           % (l, lambda_vals[l], g_a_integrals[l]))
    if tau_vals is not None and not np.all(np.asarray(tau_vals) == None):  # noqa: E711
        pr("    τ: " + ", ".join("(%.2e:%.2fπ)" % (abs(t), np.angle(t) / np.pi)
                                 for t in tau_vals))
    out.flush()


def print_table(J_T, show_g_a_int_per_pulse=False, unicode=True,
                col_formats=None, col_headers=None, out=None):
    """Return an info_hook that prints one table row per iteration
    (``iter. J_T ∫gₐ(t)dt J ΔJ_T ΔJ secs``) and returns the value of
    ``J_T(**kwargs)`` so that it ends up in ``Result.info_vals``
    (info_hooks.py:352-641)."""
    if out is None:
        out = sys.stdout
    state = {'J_T_prev': None, 'first': True}
    g_lbl = "∫gₐ(t)dt" if unicode else "g_a_int"
    d1, d2 = ("ΔJ_T", "ΔJ") if unicode else ("Delta J_T", "Delta J")

    def info_hook(**kwargs):
        J_T_val = J_T(**kwargs)
        iteration = kwargs['iteration']
        g_a = np.asarray(kwargs['g_a_integrals'])
        g_sum = float(np.sum(g_a))
        secs = int(kwargs['stop_time'] - kwargs['start_time'])
        if state['first']:
            hdr = "%-5s %9s" % ("iter.", "J_T")
            if show_g_a_int_per_pulse:
                for l in range(len(g_a)):
                    hdr += " %11s" % ("%s_%d" % (g_lbl, l + 1))
            hdr += " %11s %10s %10s %10s %5s" % (g_lbl, "J", d1, d2, "secs")
            print(hdr, file=out)
            state['first'] = False
        row = "%-5d %9.2e" % (iteration, J_T_val)
        if show_g_a_int_per_pulse:
            for v in g_a:
                row += " %11.2e" % v
        row += " %11.2e %10.2e" % (g_sum, J_T_val + g_sum)
        if iteration == 0 or state['J_T_prev'] is None:
            row += " %10s %10s" % ("n/a", "n/a")
        else:
            dJ_T = J_T_val - state['J_T_prev']
            row += " %10.2e %10.2e" % (dJ_T, dJ_T + g_sum)
        row += " %5d" % secs
        print(row, file=out)
        out.flush()
        state['J_T_prev'] = J_T_val
        return J_T_val

    return info_hook
