'''Numpy model of the blocked in-place Gauss-Jordan inversion of csrc/tfb_direct.cu (k_gj_blocked): panels of NB columns are
eliminated on their own (pivot search among the rows not yet used), the rest of the matrix receives one rank-NB update per
panel.  Checks the algebra (incl. the row swaps and the final column gather) against numpy.linalg.inv.'''
import numpy


def gj_panel(P, k0):
    '''In-place Gauss-Jordan steps on the m x nb panel P (columns k0..k0+nb of the matrix); returns the pivot rows.'''
    m, nb = P.shape
    piv = []
    for s in range(nb):
        k = k0 + s
        p = k + int(numpy.argmax(numpy.abs(P[k:, s])))
        piv.append(p)
        if p != k:
            P[[k, p], :] = P[[p, k], :]
        rowk = P[k, :].copy()
        colk = P[:, s].copy()
        pivinv = 1.0 / colk[k]
        f = colk * pivinv
        P -= numpy.outer(f, rowk)
        P[:, s] = -f
        P[k, :] = rowk * pivinv
        P[k, s] = pivinv
    return piv


def blocked_inverse(A, NB):
    A = A.copy()
    m = A.shape[0]
    pivots = []
    for k0 in range(0, m, NB):
        nb = min(NB, m - k0)
        K = slice(k0, k0 + nb)
        P = A[:, K].copy()
        piv = gj_panel(P, k0)
        pivots += piv
        # the other columns: the same row swaps, then  A_J <- [rows outside K] A_J + P_new * A_K,J(after the swaps)
        J = numpy.r_[0:k0, k0 + nb:m]
        AJ = A[:, J]
        for s, p in enumerate(piv):
            if p != k0 + s:
                AJ[[k0 + s, p], :] = AJ[[p, k0 + s], :]
        AK = AJ[K, :].copy()
        AJ[K, :] = 0.0
        AJ += P @ AK
        A[:, J] = AJ
        A[:, K] = P
    colsrc = list(range(m))
    for k in range(m - 1, -1, -1):
        p = pivots[k]
        if p != k:
            colsrc[k], colsrc[p] = colsrc[p], colsrc[k]
    return A[:, colsrc]


if __name__ == '__main__':
    rng = numpy.random.default_rng(0)
    for m, NB in ((7, 3), (40, 8), (96, 16), (130, 24)):
        A = rng.standard_normal((m, m))
        A[rng.integers(0, m, m // 3), rng.integers(0, m, m // 3)] = 0.0
        for d in range(0, m, 3):
            A[d, d] = 0.0           # zero diagonal entries, as the pressure rows have
        err = numpy.abs(blocked_inverse(A, NB) @ A - numpy.eye(m)).max()
        print(m, NB, 'max |inv A - I| = %.2e' % err)
        assert err < 1e-9
