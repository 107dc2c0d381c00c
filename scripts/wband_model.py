"""numpy model of K3e (kernels_wband.cu): the exact solve of a WIDE block-banded reduced system, cut into C chunks
with separators of w poses, every chunk factored as a bordered band (border rows = [left separator | rhs |
right separator]) in LAPACK-style band storage, panels of 48 columns.  Same storage formulas and the same loop
structure as the CUDA kernels (fill -> per panel: factor + trailing update -> separator assembly -> dense
separator solve -> per chunk back-substitution), so an indexing mistake shows up here, on the CPU.

Usage: python scripts/wband_model.py   (random SPD block-banded systems, checks against numpy.linalg.solve)"""
import numpy as np

NB = 48


def plan(n_free, w, C):
    """Chunk layout: interior lengths (poses), first pose of every chunk / separator, padded interior size."""
    inter = n_free - (C - 1) * w
    base, extra = divmod(inter, C)
    lens = [base + (1 if c < extra else 0) for c in range(C)]
    p0, pos = [], 0
    for c in range(C):
        p0.append(pos)
        pos += lens[c] + (w if c < C - 1 else 0)
    assert pos == n_free
    m_pad = (6 * max(lens) + NB - 1) // NB * NB
    return lens, p0, m_pad


def solve(S_blocks, rhs, n_free, w, C):
    """S_blocks: dict (a, b) -> 6x6 block, a <= b <= a + w (upper block band).  Returns y."""
    lens, p0, m_pad = plan(n_free, w, C)
    assert min(lens) >= w + 1
    sepw = 6 * w if C > 1 else 0
    nbr = 2 * sepw + 1
    NBRP = nbr + (nbr & 1)
    bw = 6 * w + 5
    BWR = (bw + 7) // 8 * 8
    ld = BWR + NB                      # band storage: element (i, j), i >= j, at A[j * ld + i]
    ldB = NBRP
    # border rows: [left separator | rhs | right separator]; the right separator's rows stay zero until the chunk's
    # last w poses, so panels before r_start carry only the first sepw + 1 border rows
    r_start = 6 * (min(lens) - w) // NB * NB if C > 1 else 0
    A = np.zeros((C, m_pad * (ld + 1)))
    Bd = np.zeros((C, (m_pad + NBRP) * ldB))
    ns = (C - 1) * sepw
    ns_pad = max(NB, (ns + NB - 1) // NB * NB)
    T = np.zeros((ns_pad + 1, ns_pad))   # dense lower triangle [row][col], last row = rhs

    # which chunk / separator a pose belongs to
    kind = np.zeros(n_free, dtype=int)   # 0 interior, 1 separator
    owner = np.zeros(n_free, dtype=int)
    local = np.zeros(n_free, dtype=int)
    for c in range(C):
        for a in range(lens[c]):
            kind[p0[c] + a], owner[p0[c] + a], local[p0[c] + a] = 0, c, a
        if c < C - 1:
            for a in range(w):
                q = p0[c] + lens[c] + a
                kind[q], owner[q], local[q] = 1, c, a

    # ---- fill (wband_fill_kernel) ----
    for (a, b), blk in S_blocks.items():
        assert a <= b <= a + w
        for r in range(6):
            for cc in range(6):
                v = blk[r, cc]
                if kind[a] == 0 and kind[b] == 0:
                    assert owner[a] == owner[b]
                    i, j = 6 * local[b] + cc, 6 * local[a] + r
                    if i >= j:
                        A[owner[a], j * ld + i] = v
                elif kind[a] == 0 and kind[b] == 1:       # right separator of a's chunk
                    assert owner[b] == owner[a]
                    assert 6 * local[a] + r >= r_start
                    Bd[owner[a], (6 * local[a] + r) * ldB + sepw + 1 + 6 * local[b] + cc] = v
                elif kind[a] == 1 and kind[b] == 0:       # left separator of b's chunk
                    assert owner[b] == owner[a] + 1
                    Bd[owner[b], (6 * local[b] + cc) * ldB + 6 * local[a] + r] = v
                else:
                    assert owner[a] == owner[b]
                    i, j = owner[b] * sepw + 6 * local[b] + cc, owner[a] * sepw + 6 * local[a] + r
                    if i >= j:
                        T[i, j] = v
    for c in range(C):
        for j in range(m_pad):
            if j < 6 * lens[c]:
                Bd[c, j * ldB + sepw] = rhs[6 * p0[c] + j]
            else:
                A[c, j * ld + j] = 1.0
    for j in range(ns_pad):
        if j < ns:
            s, o = divmod(j, sepw)
            T[ns_pad, j] = rhs[6 * (p0[s] + lens[s]) + o]
        else:
            T[j, j] = 1.0

    def row_addr(c, j0, mb, rl, col):
        """address of (local row rl of the panel at j0, band column `col`) -> (array, index)"""
        t0 = j0 + NB
        if rl < mb:
            return A[c], col * ld + t0 + rl
        return Bd[c], col * ldB + (rl - mb)

    # ---- factorisation: per panel, all chunks ----
    Ldiag = np.zeros((C, m_pad // NB, NB, NB))
    for j0 in range(0, m_pad, NB):
        t0 = j0 + NB
        mb = min(BWR, m_pad - t0)
        m = mb + (nbr if j0 >= r_start else sepw + 1)
        for c in range(C):
            # panel kernel: diagonal block + rows below
            D = np.zeros((NB, NB))
            for cc in range(NB):
                for r in range(cc, NB):
                    D[r, cc] = A[c, (j0 + cc) * ld + j0 + r]
            L11 = np.linalg.cholesky(D + np.tril(D, -1).T)
            Ldiag[c, j0 // NB] = L11
            X = np.zeros((m, NB))
            for rl in range(m):
                for cc in range(NB):
                    arr, ix = row_addr(c, j0, mb, rl, j0 + cc)
                    X[rl, cc] = arr[ix]
            X = np.linalg.solve(L11, X.T).T          # l L11^T = a
            for rl in range(m):
                for cc in range(NB):
                    arr, ix = row_addr(c, j0, mb, rl, j0 + cc)
                    arr[ix] = X[rl, cc]
            # syrk kernel: trailing update over the local index space [0, m), lower triangle only
            U = X @ X.T
            for jl in range(m):
                for il in range(jl, m):
                    if jl < mb:
                        arr, ix = row_addr(c, j0, mb, il, t0 + jl)
                    else:
                        arr, ix = Bd[c], (m_pad + (jl - mb)) * ldB + (il - mb)
                    arr[ix] -= U[il, jl]

    # ---- separator assembly (gather: deterministic) ----
    def dp(c, bi, bj):
        return Bd[c, (m_pad + bj) * ldB + bi]
    for gj in range(ns):
        sj, oj = divmod(gj, sepw)
        for gi in range(gj, ns):
            si, oi = divmod(gi, sepw)
            if si == sj:
                T[gi, gj] += dp(si, sepw + 1 + oi, sepw + 1 + oj) + dp(si + 1, oi, oj)
            elif si == sj + 1:
                T[gi, gj] += dp(si, sepw + 1 + oi, oj)
        # (rhs, R_j) lies ABOVE the diagonal of D' in this order: D' is symmetric, the kernels keep (R_j, rhs)
        T[ns_pad, gj] += dp(sj, sepw + 1 + oj, sepw) + dp(sj + 1, sepw, oj)

    # ---- dense separator solve (K3d) ----
    xsep = np.zeros(ns)
    if ns:
        Tm = np.tril(T[:ns_pad]) + np.tril(T[:ns_pad], -1).T
        xsep = np.linalg.solve(Tm, T[ns_pad])[:ns]

    # ---- back-substitution per chunk ----
    y = np.zeros(6 * n_free)
    for c in range(C):
        xs = np.zeros(nbr)       # (the rhs slot stays zero)
        if C > 1:
            if c > 0:
                xs[:sepw] = xsep[(c - 1) * sepw:c * sepw]
            if c < C - 1:
                xs[sepw + 1:] = xsep[c * sepw:(c + 1) * sepw]
        xw = np.zeros(m_pad)
        for j in range(m_pad):
            xw[j] = Bd[c, j * ldB + sepw] - Bd[c, j * ldB:j * ldB + nbr] @ xs
        for j0 in range(m_pad - NB, -1, -NB):
            L11 = Ldiag[c, j0 // NB]
            xp = np.linalg.solve(L11.T, xw[j0:j0 + NB])
            xw[j0:j0 + NB] = xp
            for j in range(max(0, j0 - BWR), j0):
                xw[j] -= A[c, j * ld + j0:j * ld + j0 + NB] @ xp
        y[6 * p0[c]:6 * (p0[c] + lens[c])] = xw[:6 * lens[c]]
        if c < C - 1:
            q = 6 * (p0[c] + lens[c])
            y[q:q + sepw] = xsep[c * sepw:(c + 1) * sepw]
    return y


def random_system(n_free, w, seed):
    rng = np.random.default_rng(seed)
    n = 6 * n_free
    J = np.zeros((n + 40, n))
    # SPD with exactly block half-bandwidth w: sum of "landmark" terms over windows of w + 1 poses
    M = np.zeros((n, n))
    for a in range(n_free):
        hi = min(n_free, a + w + 1)
        k = 6 * (hi - a)
        G = rng.standard_normal((k, 4))
        M[6 * a:6 * hi, 6 * a:6 * hi] += G @ G.T
    M += np.eye(n) * 0.5
    blocks = {}
    for a in range(n_free):
        for b in range(a, min(n_free, a + w + 1)):
            blocks[(a, b)] = M[6 * a:6 * a + 6, 6 * b:6 * b + 6].copy()
    return M, blocks, rng.standard_normal(n)


if __name__ == "__main__":
    for (n_free, w, C) in [(40, 3, 1), (60, 3, 3), (75, 5, 4), (90, 13, 2)]:
        M, blocks, rhs = random_system(n_free, w, 7 + n_free)
        y = solve(blocks, rhs, n_free, w, C)
        ref = np.linalg.solve(M, rhs)
        err = np.abs(y - ref).max() / np.abs(ref).max()
        print("n_free %d w %d C %d  rel err %.2e" % (n_free, w, C, err))
        assert err < 1e-9
