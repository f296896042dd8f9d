"""TEST INFRASTRUCTURE ONLY (never imported by the product path).

CPU restatement of the hierarchical-clustering step clusttraj runs on the condensed RMSD vector:
`scipy.cluster.hierarchy.linkage(distmat, method)` as called at clusttraj/classify.py:30,99,127 (SciPy is a
third-party dependency of the reference, `scipy` in requirements.txt, not vendored under /root/reference; the
version importable in this image is the pin).  What is restated is SciPy's published algorithm for the
methods the GPU path covers:

  * complete / average / weighted / ward : the nearest-neighbour-chain algorithm (Murtagh; Müllner 2011, as in
    fastcluster) — grow a chain of nearest neighbours from the lowest-index live cluster, preferring the previous
    chain element on ties, merge the first reciprocal pair, keep the merged cluster in the HIGHER of the two slots,
    update its distances with the Lance-Williams formula of the method;
  * single : Prim's minimum spanning tree from point 0;
  * both followed by a STABLE sort of the merges by distance and a union-find relabelling (new cluster ids n, n+1, ...
    in sorted order, the two ids of a merge in ascending order).

Pinned against the live module: tests/test_oracle_golden.py compares every method bit for bit with
`scipy.cluster.hierarchy.linkage` on random, tied and duplicate-laden inputs, so the step order, the tie rules and the
floating-point operation order below are the ones SciPy executes — which is what the GPU kernel has to reproduce."""
import numpy as np

METHODS = {"single": 0, "complete": 1, "average": 2, "weighted": 6, "ward": 5}   # SciPy's method codes


def _cidx(n, i, j):
    if i < j:
        return n * i - (i * (i + 1)) // 2 + (j - i - 1)
    return n * j - (j * (j + 1)) // 2 + (i - j - 1)


def _update(method, d_xi, d_yi, d_xy, nx, ny, ni):
    if method == "complete":
        return max(d_xi, d_yi)
    if method == "average":
        return (nx * d_xi + ny * d_yi) / (nx + ny)
    if method == "weighted":
        return 0.5 * (d_xi + d_yi)
    if method == "ward":
        t = 1.0 / (nx + ny + ni)
        return np.sqrt((ni + nx) * t * d_xi * d_xi + (ni + ny) * t * d_yi * d_yi - ni * t * d_xy * d_xy)
    raise ValueError(method)


def _label(Z, n):
    parent = list(range(2 * n - 1))
    size = [1] * n + [0] * (n - 1)
    nxt = n

    def find(x):
        r = x
        while parent[r] != r:
            r = parent[r]
        while parent[x] != r:
            parent[x], x = r, parent[x]
        return r

    for k in range(n - 1):
        xr, yr = find(int(Z[k, 0])), find(int(Z[k, 1]))
        Z[k, 0], Z[k, 1] = (xr, yr) if xr < yr else (yr, xr)
        parent[xr] = nxt
        parent[yr] = nxt
        size[nxt] = size[xr] + size[yr]
        Z[k, 3] = size[nxt]
        nxt += 1


def raw_merges(dists, method):
    """The merges in the order the algorithm finds them (before the sort / relabelling): float64[n-1, 4]."""
    dists = np.asarray(dists, dtype=np.float64)
    n = int(round((1 + np.sqrt(1 + 8 * dists.size)) / 2))
    Z = np.empty((n - 1, 4))
    if method == "single":
        merged = np.zeros(n, dtype=bool)
        D = np.full(n, np.inf)
        x = 0
        y = 0
        for k in range(n - 1):
            cur = np.inf
            merged[x] = True
            for i in range(n):
                if merged[i]:
                    continue
                d = dists[_cidx(n, x, i)]
                if D[i] > d:
                    D[i] = d
                if D[i] < cur:
                    y = i
                    cur = D[i]
            Z[k] = (x, y, cur, 0)
            x = y
        return Z
    D = dists.copy()
    size = np.ones(n, dtype=np.int64)
    chain = np.empty(n, dtype=np.int64)
    clen = 0
    y = 0
    for k in range(n - 1):
        if clen == 0:
            clen = 1
            chain[0] = int(np.nonzero(size > 0)[0][0])
        while True:
            x = int(chain[clen - 1])
            if clen > 1:
                y = int(chain[clen - 2])
                cur = D[_cidx(n, x, y)]
            else:
                cur = np.inf
            for i in range(n):
                if size[i] == 0 or i == x:
                    continue
                d = D[_cidx(n, x, i)]
                if d < cur:
                    cur = d
                    y = i
            if clen > 1 and y == chain[clen - 2]:
                break
            chain[clen] = y
            clen += 1
        clen -= 2
        if x > y:
            x, y = y, x
        nx, ny = int(size[x]), int(size[y])
        Z[k] = (x, y, cur, nx + ny)
        size[x] = 0
        size[y] = nx + ny
        for i in range(n):
            ni = int(size[i])
            if ni == 0 or i == y:
                continue
            D[_cidx(n, i, y)] = _update(method, D[_cidx(n, i, x)], D[_cidx(n, i, y)], cur, nx, ny, ni)
    return Z


def finish(Zraw):
    """Stable sort by distance + union-find relabelling (what SciPy does after either algorithm)."""
    n = Zraw.shape[0] + 1
    Z = Zraw[np.argsort(Zraw[:, 2], kind="mergesort")].copy()
    _label(Z, n)
    return Z


def linkage(dists, method="average"):
    return finish(raw_merges(dists, method))
