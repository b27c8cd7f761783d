"""Oracle: min-cost assignment (TEST INFRASTRUCTURE ONLY, see oracle/__init__).

The reference calls ``scipy.optimize.linear_sum_assignment(C)`` at
lib/core/tracking_engine.py:237 and accepts EVERY returned pair (no gating,
:244-246).  scipy is an un-vendored third-party dependency: the reference pins
scipy 1.0.1 (all_pkg_versions.txt:234, a pure-python Munkres); this image
ships scipy 1.18.1, whose solver is the rectangular shortest-augmenting-path
algorithm of D. F. Crouse, "On implementing 2D rectangular assignment
algorithms", IEEE T-AES 52(4), 2016 (scipy/optimize/rectangular_lsap).

Tracking cost matrices are dominated by exact ties (every non-overlapping pair
costs exactly 1.0f), so *which* optimum is returned depends on the solver's
visiting order.  ``lsap_crouse`` restates the published algorithm with the
visiting order of the scipy >= 1.6 implementation (remaining-column list filled
in reverse, swap-with-last removal, "prefer an unassigned column on ties"),
in float64 like scipy.  It is pinned live against the installed scipy in
tests/test_oracle_lsa.py (thousands of random, tie-heavy, rectangular cases),
and the device solver mirrors the same order, so indices match scipy>=1.6
bit-for-bit.  Against the reference's pinned scipy 1.0.1 the guarantee is
equal total cost, and identical indices whenever the optimum is unique.
"""
import numpy as np


def lsap_crouse(cost):
    """Returns (row_ind, col_ind) exactly like scipy.optimize.linear_sum_assignment."""
    cost = np.asarray(cost, dtype=np.float64)
    nr, nc = cost.shape
    if nr == 0 or nc == 0:
        return np.zeros(0, np.int64), np.zeros(0, np.int64)
    transpose = nc < nr
    if transpose:
        cost = cost.T.copy()
        nr, nc = nc, nr
    INF = np.inf
    u = [0.0] * nr
    v = [0.0] * nc
    spc = [INF] * nc
    path = [-1] * nc
    col4row = [-1] * nr
    row4col = [-1] * nc
    c = cost.tolist()
    for cur_row in range(nr):
        # ---- augmenting_path ----
        min_val = 0.0
        remaining = [nc - it - 1 for it in range(nc)]
        num_remaining = nc
        SR = [False] * nr
        SC = [False] * nc
        for j in range(nc):
            spc[j] = INF
        sink = -1
        i = cur_row
        while sink == -1:
            index = -1
            lowest = INF
            SR[i] = True
            ci = c[i]
            ui = u[i]
            for it in range(num_remaining):
                j = remaining[it]
                r = min_val + ci[j] - ui - v[j]
                if r < spc[j]:
                    path[j] = i
                    spc[j] = r
                if spc[j] < lowest or (spc[j] == lowest and row4col[j] == -1):
                    lowest = spc[j]
                    index = it
            min_val = lowest
            if min_val == INF:
                raise ValueError('cost matrix is infeasible')
            j = remaining[index]
            if row4col[j] == -1:
                sink = j
            else:
                i = row4col[j]
            SC[j] = True
            num_remaining -= 1
            remaining[index] = remaining[num_remaining]
        # ---- dual update ----
        u[cur_row] += min_val
        for i2 in range(nr):
            if SR[i2] and i2 != cur_row:
                u[i2] += min_val - spc[col4row[i2]]
        for j2 in range(nc):
            if SC[j2]:
                v[j2] -= min_val - spc[j2]
        # ---- augment ----
        j = sink
        while True:
            i2 = path[j]
            row4col[j] = i2
            col4row[i2], j = j, col4row[i2]
            if i2 == cur_row:
                break
    col4row = np.asarray(col4row, dtype=np.int64)
    if transpose:
        order = np.argsort(col4row, kind='stable')
        return col4row[order], order.astype(np.int64)
    return np.arange(nr, dtype=np.int64), col4row


def bipartite_matching_greedy(C):
    """lib/core/tracking_engine.py:184-206 (argmin of the shrinking matrix)."""
    C = np.array(C, copy=True)
    prev_ids, cur_ids = [], []
    row_ids = np.arange(C.shape[0])
    col_ids = np.arange(C.shape[1])
    while C.size > 0:
        i, j = np.unravel_index(C.argmin(), C.shape)
        prev_ids.append(int(row_ids[i]))
        cur_ids.append(int(col_ids[j]))
        C = np.delete(np.delete(C, i, 0), j, 1)
        row_ids = np.delete(row_ids, i, 0)
        col_ids = np.delete(col_ids, j, 0)
    return prev_ids, cur_ids
