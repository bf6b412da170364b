"""CPU probe: causal fixed-point (Picard) iteration for the fused update/forward
sweep.  Counts iterations until the pulse matches the sequential sweep."""
import sys, numpy as np, scipy.linalg
sys.path.insert(0, '/root/repo')
from krotov_b200 import workloads
from oracle import krotov_oracle as orc

def probe(name, iters=2, **kw):
    wl = workloads.by_name(name, **kw)
    low = wl.lowered()
    if wl.chi == 'qubit_reset':
        fixed = wl.meta['chi_fixed'].reshape(-1, order='F')
        chi = lambda fw, targets, tau, weights=None: [fixed.copy() for _ in fw]
    else:
        chi = {'re': orc.chis_re, 'ss': orc.chis_ss, 'sm': orc.chis_sm, 'hs': orc.chis_hs}[wl.chi]
    rec = orc.optimize(low['terms'], low['psi0'], low['targets'], low['pulses'], low['shapes'],
                       low['lambdas'], low['tlist'], chi, iter_stop=iters, is_super=low['is_super'],
                       operator_norm='fro')
    terms = low['terms']; K = len(terms); tl = low['tlist']; NT = len(tl) - 1
    dt = np.diff(tl)
    for it in range(1, iters + 1):
        guess = rec[it - 1]['optimized_pulses'][0]
        want = rec[it]['optimized_pulses'][0]
        X = np.array(rec[it]['backward_states'])       # K, nt, N
        cn = np.array(rec[it]['chi_norms'])
        H0 = np.array([sum(op for op, l in t if l < 0) for t in terms])
        H1 = np.array([sum(op for op, l in t if l == 0) for t in terms])
        mu = H1
        f = 1 if low['is_super'] else -1j
        if low['is_super']:
            mu = 1j * H1
        psi0 = np.array(low['psi0'])
        sl = low['shapes'][0] / low['lambdas'][0]
        eps = guess.copy()
        # eta[k,n] = mu^dag chi
        for j in range(1, 60):
            A = f * (H0[:, None] + eps[None, :, None, None] * H1[:, None]) * dt[None, :, None, None]
            U = scipy.linalg.expm(A)                   # K, NT, N, N
            phi = np.empty((K, NT, psi0.shape[1]), complex)
            cur = psi0.copy()
            for n in range(NT):
                phi[:, n] = cur
                cur = np.einsum('kab,kb->ka', U[:, n], cur)
            d = np.einsum('k,kna,kab,knb->n', cn, X[:, :NT].conj(), mu, phi).imag
            new = guess + sl * d
            delta = np.max(np.abs(new - eps)) / np.max(np.abs(want))
            err = np.max(np.abs(new - want)) / np.max(np.abs(want))
            eps = new
            print(f"{name} it{it} picard {j}: change {delta:.2e}  err {err:.2e}")
            if err < 1e-14 and delta < 1e-15:
                break

if __name__ == '__main__':
    probe('C4', K=16)
    probe('C1')
    probe('C2')


def probe_windows(name, nwins=(1, 2, 4, 8, 16, 32), it=2, **kw):
    """Windowed fixed point: the time grid is cut into windows that are solved
    one after the other (each a Picard problem of its own, started from the
    final states of the window before).  Prints rounds per window and the
    work in units of one full-grid round (sum rounds_w * len_w / NT)."""
    wl = workloads.by_name(name, **kw)
    low = wl.lowered()
    if wl.chi == 'qubit_reset':
        fixed = wl.meta['chi_fixed'].reshape(-1, order='F')
        chi = lambda fw, targets, tau, weights=None: [fixed.copy() for _ in fw]
    else:
        chi = {'re': orc.chis_re, 'ss': orc.chis_ss, 'sm': orc.chis_sm, 'hs': orc.chis_hs}[wl.chi]
    rec = orc.optimize(low['terms'], low['psi0'], low['targets'], low['pulses'], low['shapes'],
                       low['lambdas'], low['tlist'], chi, iter_stop=it, is_super=low['is_super'],
                       operator_norm='fro')
    terms = low['terms']; K = len(terms); tl = low['tlist']; NT = len(tl) - 1
    dt = np.diff(tl)
    guess = rec[it - 1]['optimized_pulses'][0]
    want = rec[it]['optimized_pulses'][0]
    X = np.array(rec[it]['backward_states']); cn = np.array(rec[it]['chi_norms'])
    H0 = np.array([sum(op for op, l in t if l < 0) for t in terms])
    H1 = np.array([sum(op for op, l in t if l == 0) for t in terms])
    f = 1 if low['is_super'] else -1j
    mu = 1j * H1 if low['is_super'] else H1
    sl = low['shapes'][0] / low['lambdas'][0]
    for nwin in nwins:
        bounds = np.linspace(0, NT, nwin + 1).astype(int)
        cur0 = np.array(low['psi0']).copy()
        eps = guess.copy()
        rounds = []
        for a, b in zip(bounds, bounds[1:]):
            for j in range(1, 200):
                A = f * (H0[:, None] + eps[None, a:b, None, None] * H1[:, None]) * dt[None, a:b, None, None]
                U = scipy.linalg.expm(A)
                cur = cur0.copy()
                phi = np.empty((K, b - a, cur.shape[1]), complex)
                for n in range(b - a):
                    phi[:, n] = cur
                    cur = np.einsum('kab,kb->ka', U[:, n], cur)
                d = np.einsum('k,kna,kab,knb->n', cn, X[:, a:b].conj(), mu, phi).imag
                new = guess[a:b] + sl[a:b] * d
                delta = np.max(np.abs(new - eps[a:b])) / np.max(np.abs(want))
                eps[a:b] = new
                if delta < 2e-14:
                    break
            rounds.append(j)
            # final states of the window under the converged pulse
            A = f * (H0[:, None] + eps[None, a:b, None, None] * H1[:, None]) * dt[None, a:b, None, None]
            U = scipy.linalg.expm(A)
            for n in range(b - a):
                cur0 = np.einsum('kab,kb->ka', U[:, n], cur0)
        work = sum(r * (b - a) for r, a, b in zip(rounds, bounds, bounds[1:])) / NT
        err = np.max(np.abs(eps - want)) / np.max(np.abs(want))
        print(f"{name} it{it} windows {nwin:3d}: rounds/window max {max(rounds)} mean {np.mean(rounds):.1f} "
              f"sequential depth {sum(rounds)}  work {work:.1f} full rounds  err {err:.1e}", flush=True)
