"""Pure-Python model of the long-pair fill kernel's schedule and arithmetic (ga_long.cuh,
long_fill_kernel): chains of rows skewed inside a lane, values in the gap-shifted domain,
band-boundary rows and column checkpoints as the traceback kernels read them.  It mirrors the
CUDA code statement by statement (same state variables, same update order) so that the
algorithm -- who reads which value from which step -- is checked against the oracle's full
table on the CPU, where there is no GPU.  Test infrastructure only.

Layout: a band = LANES lanes x LR rows, LR = NC chains x ELL rows.  Chain q = g*NC + k (lane g,
chain k) works at step t on column j = t - q + 1; its ELL rows are updated top to bottom within
the step (dependent), different chains of a lane are independent within a step.

Shifted domain: X^(i,j) = 4 * (X(i,j) - GH(j) - GV(i)), GH(j) = sum_{t<=j} gap_h[s2[t-1]],
GV(i) = sum_{t<=i} gap_v[s1[t-1]].  Then
    d0 = h^(i-1,j-1) + 4*(sub(a,b) - gap_h[b] - gap_v[a])
    h  = min(d0, L^(i,j-1), U^(i-1,j))
    L^(i,j) = min(h + 4o, L^(i,j-1)),   U^(i,j) = min(h + 4o, U^(i-1,j))
(the kernel evaluates the U recurrence of a chain of rows in an equivalent form that keeps U off the
row-to-row dependency, see chain() below)
where L(i,j) = min(D0+o, D1, D2+o)(i,j) is the level-1 candidate handed to column j+1 and
U(i,j) = min(D0+o, D1+o, D2)(i,j) the level-2 candidate handed to row i+1: the gap-extension
additions disappear from the cell update.
"""
from __future__ import annotations

import numpy as np

INVALID = -(2 ** 31) + 0x00808080      # 0x80808080 as int32: "boundary entry not yet written"
L_INF, L_BIGADD = 0x3F000000, 0x4F000000


class Scheme:
    def __init__(self, A, sub, gap_h, gap_v, o, max_cost):
        self.A, self.sub, self.gap_h, self.gap_v, self.o, self.max_cost = A, sub, gap_h, gap_v, o, max_cost


def fill_long(s1, s2, sc: Scheme, ELL=2, NC=2, LANES=4, LSYNC=4, LCC=8):
    """-> (cost, brow, vchk) with brow[band][j] = (h^, U^) of the band's last row and
    vchk[band][c][g][r] = (h^, L^) of row g*LR+r of the band at column c*LCC (shifted, x4)."""
    m, n = len(s1), len(s2)
    LR = ELL * NC
    LBAND = LANES * LR
    nbands = (m + LBAND - 1) // LBAND
    nck = n // LCC + 2
    o = sc.o
    o4 = 4 * o
    big = (sc.max_cost + 1) * max(m, n)
    colsum = [o]
    for i in range(1, m + 1):
        colsum.append(colsum[-1] + sc.gap_v[s1[i - 1]])
    rowsum = [o]
    for j in range(1, n + 1):
        rowsum.append(rowsum[-1] + sc.gap_h[s2[j - 1]])
    gv_pad = list(sc.gap_v) + [0]
    # table of shifted substitution costs, row A = padding letter (rows past the end of seq_1)
    tab = [[4 * ((sc.sub[a * sc.A + b] if a < sc.A else 0) - sc.gap_h[b] - gv_pad[a]) for b in range(sc.A)]
           for a in range(sc.A + 1)]
    brow = [[None] * (n + 1) for _ in range(nbands)]
    vchk = [[[[None] * LR for _ in range(LANES)] for _ in range(nck)] for _ in range(nbands)]
    cost = None

    def col0(i):                       # shifted column-0 values of DP row i: (h^, L^)
        c = colsum[min(i, m)]
        gvs = c - o
        return 4 * (min(big, c) - gvs), 4 * (min(big, c + o) - gvs)

    assert LANES * NC <= LCC, "at most one chain of a band is on a checkpoint column per step"
    for band in range(nbands):
        # ---- per-lane state
        aw = [[0] * LR for _ in range(LANES)]
        hl = [[0] * LR for _ in range(LANES)]
        Lp = [[0] * LR for _ in range(LANES)]
        ub = [[0] * NC for _ in range(LANES)]
        hd = [[0] * NC for _ in range(LANES)]
        bv = [[0] * NC for _ in range(LANES)]      # letter of chain k at the CURRENT step (runs one step ahead of the cells)
        sv = [[0] * LR for _ in range(LANES)]      # shifted substitution cost of row r at the current step
        rh_prev = [0] * LANES
        for g in range(LANES):
            ib = band * LBAND + g * LR
            for r in range(LR):
                row = ib + r
                aw[g][r] = s1[row] if row < m else sc.A
                hl[g][r], Lp[g][r] = col0(ib + r + 1)
            rh_prev[g] = 0 if ib == 0 else col0(ib)[0]
        bv[0][0] = s2[0]                           # lane 0, chain 0 starts on column 1
        for g in range(LANES):
            for r in range(LR):
                sv[g][r] = tab[aw[g][r]][bv[g][r // ELL]]
        nsteps = n + LANES * NC - 1
        rh_in = rh_prev[:]                         # what a shuffle at the end of a virtual step -1 would deliver
        ru_in = [None] * LANES
        for t0 in range(0, nsteps, LSYNC):
            # staged inputs of lane 0, chain 0: entry u = (h^, U^ of the row above at column c, letter of
            # column c + 1) for c = t0 + 1 + u
            stage = []
            for u in range(LSYNC):
                cj = t0 + 1 + u
                nxt = s2[cj] if cj + 1 <= n else 0
                if cj <= n:
                    if band > 0:
                        v = brow[band - 1][cj]
                        assert v is not None and v[0] != INVALID
                    else:
                        rj = rowsum[cj]
                        ghs = rj - o
                        v = (4 * (min(big, rj) - ghs), 4 * (min(big, rj + o) - ghs))
                    stage.append((v[0], v[1], nxt))
                else:
                    stage.append((0, 0, 0))
            for u in range(LSYNC):
                t = t0 + u
                # The kernel shuffles the bottom row of a lane's LAST chain right after that chain's
                # update and the lane below consumes it one step later (rh_in / ru_in); the letter shuffle
                # reads the letters of the current step.  All lanes move in lock-step: phases.
                nb = [bv[g - 1][NC - 1] if g > 0 else stage[u][2] for g in range(LANES)]

                def chain(g, k, diag, up):
                    """cells of chain k of lane g at this step.  The kernel does not skip an inactive
                    chain: it runs the update with inputs that make it the identity (d0 = own h, U from
                    above = +inf, infinite gap open for the U handed down); hd / ub become garbage."""
                    j = t - (g * NC + k) + 1
                    # The kernel keeps U off the row-to-row path: e_r = min(d0_r, L_r),
                    # P_r = min(e_r + 4o, P_{r-1}) (never sees U_in), h_r = min3(e_r, P_{r-1}, U_in),
                    # U handed down = min(P_last, U_in).  `ref_*` below is the textbook order it must equal.
                    if not (1 <= j <= n):
                        up = L_INF
                        P = L_INF
                        for r in range(k * ELL, (k + 1) * ELL):
                            e = min(hl[g][r], Lp[g][r])
                            h = min(e, P, up)
                            P = min(e + L_BIGADD, P)
                            assert h == hl[g][r] and min(h + o4, Lp[g][r]) == Lp[g][r], "identity"
                            assert P == L_INF
                        assert min(P, up) == L_INF
                        hd[g][k] = ub[g][k] = None          # garbage, must never be read
                        return
                    assert bv[g][k] == s2[j - 1]
                    P = L_INF
                    ref_up = up
                    for r in range(k * ELL, (k + 1) * ELL):
                        d0 = diag + sv[g][r]
                        e = min(d0, Lp[g][r])
                        h = min(e, P, up)
                        P = min(e + o4, P)
                        assert h == min(d0, Lp[g][r], ref_up), "U-free form of h"
                        ref_up = min(h + o4, ref_up)
                        assert min(P, up) == ref_up, "U-free form of U"
                        diag = hl[g][r]
                        hl[g][r] = h
                        Lp[g][r] = min(h + o4, Lp[g][r])
                    hd[g][k] = diag
                    ub[g][k] = min(P, up)
                    if k == NC - 1 and g == LANES - 1:
                        brow[band][j] = (hl[g][LR - 1], ub[g][k])

                def shuffle():
                    return ([hl[g - 1][LR - 1] if g > 0 else None for g in range(LANES)],
                            [ub[g - 1][NC - 1] if g > 0 else None for g in range(LANES)])

                if NC > 1:
                    old_hd = [row[:] for row in hd]
                    old_ub = [row[:] for row in ub]
                    for g in range(LANES):
                        chain(g, NC - 1, old_hd[g][NC - 2], old_ub[g][NC - 2])
                    rh_next, ru_next = shuffle()
                    for g in range(LANES):
                        for k in range(NC - 2, 0, -1):
                            chain(g, k, old_hd[g][k - 1], old_ub[g][k - 1])
                for g in range(LANES):
                    rh = stage[u][0] if g == 0 else rh_in[g]
                    ru = stage[u][1] if g == 0 else ru_in[g]
                    diag = rh_prev[g]
                    rh_prev[g] = rh                      # unconditional (see long_fill_step)
                    chain(g, 0, diag, ru)
                if NC == 1:
                    rh_next, ru_next = shuffle()
                rh_in, ru_in = rh_next, ru_next
                for g in range(LANES):
                    nbv = [nb[g]] + [bv[g][k - 1] for k in range(1, NC)]
                    bv[g] = nbv
                    sv[g] = [tab[aw[g][r]][nbv[r // ELL]] for r in range(LR)]
                # at most one chain of the band is on a checkpoint column, and which one follows from t alone
                qs = (t + 1) % LCC
                if qs < LANES * NC:
                    gs, ks = qs // NC, qs % NC
                    j = t - qs + 1
                    if 1 <= j <= n:
                        for r in range(ks * ELL, (ks + 1) * ELL):
                            vchk[band][j // LCC][gs][r] = (hl[gs][r], Lp[gs][r])
        if band == (m - 1) // LBAND:
            g, r = ((m - 1) % LBAND) // LR, (m - 1) % LR
            cost = hl[g][r] // 4 + (rowsum[n] - o) + (colsum[m] - o)
    return cost, brow, vchk, (rowsum, colsum, LBAND, LR)


def unshift(v, i, j, rowsum, colsum, o, m):
    """Stored (shifted, x4) value of cell (i, j) -> plain cost x4."""
    return v + 4 * ((rowsum[j] - o) + (colsum[min(i, m)] - o))
