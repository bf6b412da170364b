"""CPU probe: Anderson acceleration of the causal fixed-point iteration."""
import sys, numpy as np, scipy.linalg
sys.path.insert(0, '/root/repo')
from krotov_b200 import workloads
from oracle import krotov_oracle as orc

def setup(name, it=1, **kw):
    wl = workloads.by_name(name, **kw)
    low = wl.lowered()
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
    psi0 = np.array(low['psi0'])
    sl = low['shapes'][0] / low['lambdas'][0]
    def F(eps):
        A = -1j * (H0[:, None] + eps[None, :, None, None] * H1[:, None]) * dt[None, :, None, None]
        U = scipy.linalg.expm(A)
        phi = np.empty((K, NT, psi0.shape[1]), complex)
        cur = psi0.copy()
        for n in range(NT):
            phi[:, n] = cur
            cur = np.einsum('kab,kb->ka', U[:, n], cur)
        d = np.einsum('k,kna,kab,knb->n', cn, X[:, :NT].conj(), H1, phi).imag
        return guess + sl * d
    return F, guess, want

def run(name, m, it=1, verbose=False, **kw):
    F, guess, want = setup(name, it, **kw)
    scale = np.max(np.abs(want))
    xs, fs = [], []
    x = guess.copy()
    for j in range(1, 80):
        fx = F(x)
        r = fx - x
        res = np.max(np.abs(r)) / scale
        err = np.max(np.abs(fx - want)) / scale
        if verbose:
            print(f"{name} AA({m}) eval {j}: residual {res:.2e} err(F(x)) {err:.2e}")
        if res < 2e-14:
            break
        xs.append(x.copy()); fs.append(fx.copy())
        xs, fs = xs[-(m + 1):], fs[-(m + 1):]
        if m == 0 or len(xs) < 2:
            x = fx
            continue
        R = np.array([f - xx for f, xx in zip(fs, xs)])    # residuals
        dR = (R[1:] - R[:-1]).T                            # NT x mk
        dF = (np.array(fs)[1:] - np.array(fs)[:-1]).T
        gamma, *_ = np.linalg.lstsq(dR, R[-1], rcond=None)
        x = fs[-1] - dF @ gamma
    return j, err

if __name__ == '__main__':
    for name, kw in (('C4', dict(K=16)), ('C2', {}), ('C1', {})):
        for m in (0, 1, 2, 3, 5):
            n, err = run(name, m, **kw)
            print('==>', name, 'AA', m, 'evaluations', n, 'final err %.1e' % err, flush=True)
