"""Slow, obviously-correct numpy / pure-Python cross-checks of the C oracle.  TEST
INFRASTRUCTURE ONLY (small inputs).

* ``dtw_rows``       -- row-by-row numpy subsequence DTW + Python back-trace following the
                        published mlpy 3.5.0 rules (the algorithm behind MotifSeq.py:437).
* ``brute_min_cost`` -- true minimum warping cost by exhaustive path enumeration (tiny inputs).
* ``zscale_np`` / ``medmad_np`` -- the literal numpy/sklearn expressions of MotifSeq.py:186-200.
"""
from __future__ import annotations

import functools

import numpy as np


def dtw_rows(x, y):
    x = np.asarray(x, dtype=np.float64)
    y = np.asarray(y, dtype=np.float64)
    n, m = x.size, y.size
    cost = np.empty((n, m))
    cost[0, :] = np.abs(x[0] - y)
    for i in range(1, n):
        local = np.abs(x[i] - y)
        cost[i, 0] = local[0] + cost[i - 1, 0]
        prev = cost[i - 1]
        row = cost[i]
        for j in range(1, m):
            best = prev[j]
            if prev[j - 1] < best:
                best = prev[j - 1]
            if row[j - 1] < best:
                best = row[j - 1]
            row[j] = local[j] + best
    end = int(np.argmin(cost[-1]))
    i, j = n - 1, end
    path = [(i, j)]
    while i > 0 or j > 0:
        if i == 0:
            j -= 1
        elif j == 0:
            i -= 1
        else:
            up, dg, lf = cost[i - 1, j], cost[i - 1, j - 1], cost[i, j - 1]
            mc = min(up, dg, lf)
            if dg == mc:
                i -= 1
                j -= 1
            elif lf == mc:
                j -= 1
            else:
                i -= 1
        path.append((i, j))
    path.reverse()
    lead = 0
    for k in range(1, len(path)):
        if path[k][0] == 0:
            lead += 1
        else:
            break
    path = path[lead:]
    return float(cost[-1, end]), cost, path


def brute_min_cost(x, y):
    """min over all warping paths (steps (1,0),(0,1),(1,1)) from any (0,s) to any (n-1,e)."""
    x = [float(v) for v in x]
    y = [float(v) for v in y]
    n, m = len(x), len(y)
    best = [float("inf")]

    def walk(i, j, acc):
        # explicit enumeration, no DP: every monotone path is summed left to right
        acc = acc + abs(x[i] - y[j])
        if i == n - 1:
            best[0] = min(best[0], acc)      # free end: may stop at any column of the last row
        if i + 1 < n:
            walk(i + 1, j, acc)
        if j + 1 < m:
            walk(i, j + 1, acc)
        if i + 1 < n and j + 1 < m:
            walk(i + 1, j + 1, acc)

    for s in range(m):                        # free start: any column of row 0
        walk(0, s, 0.0)
    return best[0]


def zscale_np(sig):
    import sklearn.preprocessing
    return sklearn.preprocessing.scale(np.array(sig), axis=0, with_mean=True, with_std=True, copy=True)


def medmad_np(sig):
    sig = np.array(sig)
    arr = np.ma.array(sig).compressed()
    med = np.median(arr)
    mad = np.median(np.abs(arr - med))
    with np.errstate(divide="ignore", invalid="ignore"):
        return (sig - med) / (mad * 1.4826)
