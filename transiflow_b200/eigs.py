'''Generalized eigenvalues of  J v = lambda M v  near a target, for Interface.eigs.

The reference delegates this to jadapy's JDQZ (BaseInterface.py:294-361), a dependency that is
neither vendored nor installed.  This module implements the same contract -- the
``'Number of Eigenvalues'`` eigenvalues closest to ``'Target'``, sorted by descending real part
(BaseInterface.py:349-361) -- with a shift-and-invert Arnoldi process whose operator
``v -> (J - sigma M)^-1 M v`` is one preconditioned Krylov solve on the device per step.  The
small Hessenberg eigenproblem is solved on the host with numpy.

Host-side driver only: the vectors are numpy arrays, every product with J / M and every solve
goes through the callables the Interface passes in.
'''

import numpy


def shift_invert_arnoldi(apply_op, n, num=5, target=0.0, tol=1e-7, max_dim=60, v0=None, real_cap=100.0):
    '''Eigenvalues (and Ritz vectors) of the pencil closest to ``target``.

    apply_op(v) must return ``(J - target M)^-1 M v``; the Arnoldi basis is complex when the target is.
    Returns ``(eigenvalues[num], vectors[n, num], converged)``; eigenvalues are sorted like the
    reference (descending real part, values with real part >= real_cap last,
    BaseInterface.py:349-361).
    '''
    rng = numpy.random.default_rng(1234)
    dtype = numpy.complex128 if isinstance(target, complex) else float
    v = numpy.array(v0, dtype=dtype) if v0 is not None else rng.standard_normal(n).astype(dtype)
    # start in the range of the operator: removes the components the mass matrix annihilates
    v = apply_op(v)
    nrm = numpy.linalg.norm(v)
    if not numpy.isfinite(nrm) or nrm == 0.0:
        raise RuntimeError('shift-invert operator returned a zero / non-finite vector')
    m = max(int(max_dim), num + 2)
    V = numpy.zeros((n, m + 1), dtype=dtype)
    H = numpy.zeros((m + 1, m), dtype=dtype)
    V[:, 0] = v / nrm
    theta = y = None
    sel = None
    converged = False
    k = 0
    for j in range(m):
        w = apply_op(V[:, j])
        # classical Gram-Schmidt, two sweeps
        h = V[:, :j + 1].conj().T @ w
        w = w - V[:, :j + 1] @ h
        h2 = V[:, :j + 1].conj().T @ w
        w = w - V[:, :j + 1] @ h2
        H[:j + 1, j] = h + h2
        hn = numpy.linalg.norm(w)
        H[j + 1, j] = hn
        k = j + 1
        breakdown = hn <= 1e-14 * max(1.0, numpy.abs(H[:j + 1, j]).max())
        if not breakdown:
            V[:, j + 1] = w / hn
        if breakdown or (k >= num + 2 and (k % 2 == 0 or k == m)):
            theta, y = numpy.linalg.eig(H[:k, :k])
            sel = numpy.argsort(-numpy.abs(theta))[:min(num, k)]      # largest |theta| = closest to the target
            # Ritz residual |h_{k+1,k}| |e_k^T y| relative to |theta|
            res = numpy.abs(H[k, k - 1]) * numpy.abs(y[k - 1, sel]) / numpy.maximum(numpy.abs(theta[sel]), 1e-300)
            if breakdown or (len(sel) >= num and numpy.all(res <= tol)):
                converged = True
                break
    if theta is None:
        theta, y = numpy.linalg.eig(H[:k, :k])
        sel = numpy.argsort(-numpy.abs(theta))[:min(num, k)]
    lam = target + 1.0 / theta[sel]
    vec = V[:, :k] @ y[:, sel]
    key = [(-l.real if l.real < real_cap else real_cap) for l in lam]
    idx = numpy.argsort(key, kind='stable')
    return lam[idx], vec[:, idx], converged
