"""TEST INFRASTRUCTURE: exact site-pattern probabilities by Felsenstein pruning (new ground truth; the
reference has no likelihood code).  P(pattern) = sum_c w_c/sum(w) sum_x pi_c[x] * L_root(x) where the root's own
stem matrix is applied first, exactly as the simulator does (MarkovProcessAlongTree.hs:94-98)."""
from __future__ import annotations

import itertools

import numpy as np


def pattern_probabilities(is_mixture, ws, ds, p, parent, leaf_row):
    """p: [N][C][K][K]; returns probs [K^T] indexed by sum_t state_t * K^(T-1-t) (row 0 most significant)."""
    p = np.asarray(p, dtype=np.float64)
    n, c, k, _ = p.shape
    t = int((np.asarray(leaf_row) >= 0).sum())
    pats = np.array(list(itertools.product(range(k), repeat=t)), dtype=np.int64)  # [K^T][T]
    npat = len(pats)
    ws = np.asarray(ws, dtype=np.float64) if is_mixture else np.ones(1)
    ds = np.asarray(ds, dtype=np.float64).reshape(c, k)
    ds = ds / ds.sum(axis=1, keepdims=True)
    children = [[] for _ in range(n)]
    for i in range(1, n):
        children[parent[i]].append(i)
    total = np.zeros(npat)
    for ci in range(c):
        # partial[node] = [npat][K]: prob of the observed leaves below `node` given the state AT node
        partial = [None] * n
        for node in range(n - 1, -1, -1):
            if leaf_row[node] >= 0:
                l = np.zeros((npat, k))
                l[np.arange(npat), pats[:, leaf_row[node]]] = 1.0
            else:
                l = np.ones((npat, k))
                for ch in children[node]:
                    l = l * (partial[ch] @ p[ch, ci].T)  # sum_y P_ch[x][y] partial_ch[y]
            partial[node] = l
        root_like = partial[0] @ p[0, ci].T  # the root's stem jump from the drawn root state
        total += ws[ci] / ws.sum() * (root_like @ ds[ci])
    return total
