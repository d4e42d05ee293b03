"""Lane-level Python model of the CUDA fill + walk kernels (test aid, CPU only).

Mirrors globalign_b200/csrc/ga_fill.cuh and ga_walk.cuh step by step -- lanes
that own R consecutive rows, the anti-diagonal step loop, values scaled by 4
with the level index in the two low bits ("tagged min3"), the skewed
direction-code layout, multi-pass bands with a boundary row handed from pass to
pass -- so that the *algorithm* (boundary initialisation, tie order, layout
arithmetic, walker decoding) is checked against the oracle without a GPU.
tests/test_kernel_model.py runs it; the CUDA code is a transliteration.
"""

from __future__ import annotations

TAG = 3


def layout(R, G, rows_per_word):
    wpp = -(-R // rows_per_word)            # words per plane per step
    nw = 2 * wpp                            # words per step: plane 0 = c0, plane 1 = c1 ^ c2
    # The fill stages 2 steps and stores, per row-word w, one uint4 =
    # (plane 0, step 0), (plane 0, step 1), (plane 1, step 0), (plane 1, step 1) of that word;
    # the CUDA code lays the uint4s out [pass][pair of blocks][row-word][lane][block of the pair]
    # (four consecutive steps of a lane's row-word per 32-byte sector); the model keeps them in a
    # dict keyed (pass, block, row-word, lane), the linear address map is checked on the GPU.
    return dict(R=R, G=G, RPW=rows_per_word, WPP=wpp, NW=nw, S=2, NQ=wpp)


def fill_model(s1, s2, sub, gap_h, gap_v, o, max_cost, R=4, G=8, rows_per_word=16):
    """One pair, one 'warp' of G lanes.  Returns (cost, codes dict, layout)."""
    m, n = len(s1), len(s2)
    A = len(gap_h)
    L = layout(R, G, rows_per_word)
    B = G * R
    passes = -(-m // B)
    nsteps = n + G - 1
    nTB = -(-nsteps // L["S"])
    big4 = 4 * (max_cost + 1) * max(m, n)
    o4 = 4 * o
    codes = {}                               # (pass, 2-step block, row-word, lane) -> [p0u0, p0u1, p1u0, p1u1]
    bnd = [None] * (n + 2)                   # boundary row between passes: (h, U) per column j
    # prefix sums of vertical gap costs: col_i = o + sum_{t<=i} gap_v[s1[t-1]]
    col = [0] * (m + B + 2)
    run = o
    for i in range(1, m + B + 2):
        a = s1[i - 1] if i <= m else None
        run += gap_v[a] if a is not None else 0
        col[i] = run
    cost = None
    for p in range(passes):
        lanes = []
        for g in range(G):
            ib = p * B + g * R               # row above this lane's first row
            st = dict(ib=ib)
            st["a"] = [s1[ib + r] if ib + r < m else None for r in range(R)]
            st["hl"] = [min(big4, 4 * col[ib + r + 1]) for r in range(R)]
            st["Lp"] = [min(big4, 4 * (col[ib + r + 1] + o)) for r in range(R)]
            st["rh_prev"] = 0 if ib == 0 else min(big4, 4 * col[ib])
            st["hb"] = 0
            st["ub"] = 0
            st["bv"] = None
            lanes.append(st)
        rowj = o                              # running top-boundary value (pass 0, lane 0)
        new_bnd = [None] * (n + 2)
        for tb in range(nTB):
            stage = [[0] * (L["S"] * L["NW"]) for _ in range(G)]
            for u in range(L["S"]):
                t = tb * L["S"] + u
                # "shuffles": read the previous step's values of the lane above
                recv = [(lanes[g - 1]["hb"], lanes[g - 1]["ub"], lanes[g - 1]["bv"]) if g > 0
                        else None for g in range(G)]
                bnew = s2[t] if t < n else None
                jt = t + 1
                if bnew is not None:
                    rowj_next = rowj + gap_h[bnew]
                else:
                    rowj_next = rowj
                for g in range(G):
                    st = lanes[g]
                    j = t - g + 1
                    if g == 0:
                        if p == 0:
                            rh = min(big4, 4 * rowj_next)
                            ru = min(4 * (rowj_next + o), big4)
                        else:
                            rh, ru = bnd[jt] if 1 <= jt <= n else (0, 0)
                        bv = bnew
                    else:
                        rh, ru, bv = recv[g]
                    st["bv"] = bv
                    if not (1 <= j <= n):
                        continue
                    b = bv
                    assert b == s2[j - 1]
                    diag = st["rh_prev"]
                    st["rh_prev"] = rh
                    up = ru
                    ghT = 4 * gap_h[b] + 1
                    words = [0] * L["NW"]
                    for r in range(R):
                        a = st["a"][r]
                        s4 = 4 * sub[a * A + b] if a is not None else 0
                        gvT = (4 * gap_v[a] if a is not None else 0) + 2
                        d0 = diag + s4
                        d1 = st["Lp"][r] + ghT
                        d2 = up + gvT
                        hq = min(d0, d1, d2)          # tagged: low 2 bits = c0
                        Lq = min(hq + o4, d1)         # VIADDMNMX; tag is c0 or 1
                        Uq = min(hq + o4, d2)         # tag is c0 or 2
                        w = r // L["RPW"]
                        lu = (Lq ^ Uq) & TAG          # c1 ^ c2; with c0 the walker recovers c1 and c2
                        words[w] = words[w] * 4 + (hq & TAG)
                        words[L["WPP"] + w] = words[L["WPP"] + w] * 4 + lu
                        diag = st["hl"][r]
                        st["hl"][r] = hq & ~TAG
                        st["Lp"][r] = Lq & ~TAG
                        up = Uq & ~TAG
                    st["hb"] = st["hl"][R - 1]
                    st["ub"] = up
                    if g == G - 1:
                        new_bnd[j] = (st["hb"], st["ub"])
                    for x in range(L["NW"]):
                        stage[g][u * L["NW"] + x] = words[x]
                    # final cost: row m lives in lane (m-1-pB)//R at r=(m-1)%R
                    if j == n and p == passes - 1 and g == (m - 1 - p * B) // R:
                        cost = st["hl"][(m - 1) % R] >> 2
                rowj = rowj_next
            for g in range(G):
                for w in range(L["WPP"]):
                    codes[(p, tb, w, g)] = [stage[g][w], stage[g][L["NW"] + w],
                                            stage[g][L["WPP"] + w], stage[g][L["NW"] + L["WPP"] + w]]
        bnd = new_bnd
    return cost, codes, L


def code_at(codes, L, m, i, j, level):
    """Direction at DP cell (i, j) (1-based) for a successor on `level`, as the walker
    decodes it: c0 from plane 0; plane 1 holds c1 ^ c2 (c1 in {c0, 1}, c2 in {c0, 2})."""
    c0 = _plane_at(codes, L, m, i, j, 0)
    if level == 0 or c0 == level:
        return c0
    x = _plane_at(codes, L, m, i, j, 1)
    take = (x & level) != 0 if c0 == 0 else x != 0
    return level if take else c0


def _plane_at(codes, L, m, i, j, plane):
    R, G = L["R"], L["G"]
    B = G * R
    p = (i - 1) // B
    g = ((i - 1) % B) // R
    r = (i - 1) % R
    t = (j - 1) + g
    w = r // L["RPW"]
    word = codes[(p, t >> 1, w, g)][plane * 2 + (t & 1)]
    rows_in_word = min(L["RPW"], R - w * L["RPW"])
    shift = 2 * (rows_in_word - 1 - (r % L["RPW"]))
    return (word >> shift) & 3


def walk_model(codes, L, s1, s2, gap_h, gap_v, o, max_cost):
    """Walker: (ops in walk order, status)."""
    m, n = len(s1), len(s2)
    big = (max_cost + 1) * max(m, n)
    i, j, level = m, n, 0
    ops = []
    k = code_at(codes, L, m, i, j, level)
    ops.append(k)
    i, j = (i - 1, j - 1) if k == 0 else (i, j - 1) if k == 1 else (i - 1, j)
    level = k
    if i == 0 and j == 0:
        return ops, 0
    if i == 0 or j == 0:
        # boundary cell evaluated once by the reference (only when m == 1 or n == 1)
        if i == 0:
            row = o + sum(gap_h[b] for b in s2[:j])
            v = [big, row, big]
            want = 1
        else:
            colv = o + sum(gap_v[a] for a in s1[:i])
            v = [big, big, colv]
            want = 2
        if level == 1:
            v[0] += o
            v[2] += o
        elif level == 2:
            v[0] += o
            v[1] += o
        if v.index(min(v)) != want:
            return ops, 1
        ops.extend([want] * (j if i == 0 else i))
        return ops, 0
    while True:
        k = code_at(codes, L, m, i, j, level)
        ops.append(k)
        i, j = (i - 1, j - 1) if k == 0 else (i, j - 1) if k == 1 else (i - 1, j)
        level = k
        if i == 0:
            ops.extend([1] * j)
            return ops, 0
        if j == 0:
            ops.extend([2] * i)
            return ops, 0
