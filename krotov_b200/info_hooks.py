"""Per-iteration hooks (host callbacks, once per Krotov iteration).

``chain`` and ``print_table`` follow the calling convention of
``krotov.info_hooks`` (/root/reference/src/krotov/info_hooks.py:24-56,
352-641): every hook receives the keyword arguments listed at
info_hooks.py:59-86; the state stores are lazy sequences backed by device
tensors (downloaded only if a hook touches them).
"""
import sys
import unicodedata
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


def _width(text):
    """Printed length of `text`: code points that are not combining marks
    (stands in for the reference's `grapheme.length`, info_hooks.py:296-314)."""
    return sum(1 for ch in text if not unicodedata.combining(ch))


def _right(text, width):
    """Right-justify to `width` printed characters (info_hooks.py:317-333)."""
    return ' ' * max(0, width - _width(text)) + text


class _PerPulseLabel:
    """Header of the per-pulse g_a columns with a subscript index
    (info_hooks.py:336-349): ``.format(l=12)`` -> ``∫gₐ(ϵ₁₂)dt``."""

    def format(self, l):
        sub = ''.join(chr(ord('₀') + int(d)) for d in str(int(l)))
        return "∫gₐ(ϵ" + sub + ")dt"


_DEFAULT_FORMATS = ('%d', '%.2e', '%.2e', '%.2e', '%.2e', '%.2e', '%.2e', '%d')


def print_table(*, J_T, show_g_a_int_per_pulse=False, J_T_prev=None,
                unicode=True, col_formats=_DEFAULT_FORMATS, col_headers=None,
                out=sys.stdout):
    """Return an info_hook that prints one table row per iteration --
    iteration, ``J_T``, [one ``∫gₐ(t)dt`` column per pulse,] their sum,
    ``J = J_T + ∑∫gₐ(t)dt``, ``ΔJ_T``, ``ΔJ``, seconds -- and returns
    ``J_T(**kwargs)`` so that it ends up in ``Result.info_vals``.  Same
    keyword-only signature, validation, column widths and ``*`` markers for a
    loss of monotonic convergence as the reference (info_hooks.py:352-641):
    `J_T_prev` (default: ``info_vals[-1]``) supplies the previous value, also
    for the first row of a continued optimisation; `col_formats` /
    `col_headers` are 8-tuples (the third header is ``.format(l=...)``-ed with
    the pulse number)."""
    if J_T_prev is None:
        def J_T_prev(**kwargs):
            vals = kwargs['info_vals']
            return vals[-1] if len(vals) > 0 else 0

    default_labels = col_headers is None
    single_label = None
    if default_labels:
        if unicode:
            widths = [5, 9, 12, 12, 11, 11, 11, 6]
            col_headers = ["iter.", "J_T", _PerPulseLabel(), "∑∫gₐ(t)dt", "J",
                           "ΔJ_T", "ΔJ", "secs"]
            single_label = "∫gₐ(t)dt"
        else:
            widths = [5, 9, 11, 11, 11, 11, 11, 6]
            col_headers = ["iter.", "J_T", "g_a_int_{l}", "g_a_int", "J",
                           "Delta J_T", "Delta J", "secs"]
            single_label = "g_a_int"
    else:
        widths = [2, 4, 4, 4, 4, 4, 4, 3]
    if len(col_formats) != 8 or len(col_headers) != 8:
        raise ValueError(
            "col_formats, and col_headers must each have exactly 8 elements")
    samples = [10, 1e-15, 1e-15, 1e-15, 1e-15, -1e-15, -1e-15, 30]
    per_pulse = col_headers[2]
    try:
        if show_g_a_int_per_pulse:
            widths[2] = max(_width(col_formats[2] % samples[2]) + 1,
                            _width(per_pulse.format(l=10)) + 1)
    except (AttributeError, NameError, TypeError, KeyError) as exc:
        raise ValueError(
            "The third label %r in col_headers must support '.format(l=l)' "
            "where l is an integer: %r" % (per_pulse, exc))
    try:
        widths = [
            max(w, _width(fmt % v) + 1,
                (_width(lbl) + 1) if isinstance(lbl, str) else 0)
            for (w, fmt, lbl, v) in zip(widths, col_formats, col_headers,
                                        samples)]
    except TypeError:
        raise ValueError(
            "Invalid col_formats %r: Each element must specify a percent "
            "format string for a single value" % (col_formats, ))
    except ValueError as exc:
        raise ValueError("Invalid col_formats %r: %s" % (col_formats, exc))
    if default_labels and col_formats[0] == '%d':
        widths[0] = 5
    fmt_it, fmt_JT, _, fmt_ga, fmt_J, fmt_dJT, fmt_dJ, fmt_sec = col_formats
    (h_it, h_JT, _, h_ga, h_J, h_dJT, h_dJ, h_sec) = col_headers
    w_it, w_JT, w_gal, w_ga, w_J, w_dJT, w_dJ, w_sec = widths

    def info_hook(**kwargs):
        iteration = kwargs['iteration']
        g_a = kwargs['g_a_integrals']
        n_pulses = len(kwargs['guess_pulses'])
        wi = max(w_it, len(str(kwargs['iter_stop'])) + 1)
        wl = max(w_gal, _width(per_pulse.format(l=n_pulses)) + 1)
        split = n_pulses > 1 and show_g_a_int_per_pulse
        if iteration == 0:
            line = h_it.ljust(wi) + _right(h_JT, w_JT)
            if split:
                for l in range(n_pulses):
                    line += _right(per_pulse.format(l=l + 1), wl)
            if n_pulses > 1 or not default_labels:
                line += _right(h_ga, w_ga)
            else:
                line += _right(single_label, w_ga)
            line += _right(h_J, w_J) + _right(h_dJT, w_dJT)
            line += _right(h_dJ, w_dJ) + _right(h_sec, w_sec)
            out.write(line + "\n")
        J_T_val = J_T(**kwargs)
        g_sum = np.sum(g_a)
        secs = int(kwargs['stop_time'] - kwargs['start_time'])
        line = str(fmt_it % iteration).ljust(wi) + _right(fmt_JT % J_T_val, w_JT)
        if split:
            for l in range(n_pulses):
                line += _right(fmt_ga % g_a[l], wl)
        line += _right(fmt_ga % g_sum, w_ga)
        line += _right(fmt_J % (J_T_val + g_sum), w_J)
        marks = ""
        if iteration == 0:
            line += _right("n/a", w_dJT) + _right("n/a", w_dJ)
        else:
            dJ_T = J_T_val - J_T_prev(**kwargs)
            dJ = dJ_T + g_sum
            line += _right(fmt_dJT % dJ_T, w_dJT) + _right(fmt_dJ % dJ, w_dJ)
            if dJ_T > 0 or dJ > 0:   # loss of monotonic convergence
                marks = " " + ("*" if dJ_T > 0 else "") + ("*" if dJ > 0 else "")
        line += " " + _right(fmt_sec % secs, w_sec - 1) + marks
        out.write(line + "\n")
        out.flush()
        return J_T_val

    return info_hook


