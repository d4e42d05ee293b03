"""TEST INFRASTRUCTURE ONLY -- CPU restatement (pure Python) of globalign's DP hot path.

This module is the *oracle*: an independent restatement of the reference's
three-level min-cost dynamic programme and of its value-comparing traceback.
Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import it.  The product package
(``globalign_b200``) never imports anything under ``oracle/``.

Pinning status: PINNED.  ``tests/golden/make_golden.py`` (run in the build
container, where the real reference is importable from /root/reference/src)
checks this file against the unmodified reference on the reference's own
known-answer tests, the tutorial vectors and several thousand random cases, and
freezes the reference's outputs as ``tests/golden/*.json``; the CPU test-suite
re-checks this oracle against those frozen vectors.

Reference lines restated here (paths relative to /root/reference/src/globalign):
  * boundary initialisation          globaligner.py:756-821  (make_dp_array)
  * cell recurrence                  globaligner.py:317-363  (get_next_best_costs)
  * fill order                       globaligner.py:366-392  (dp_array_forward)
  * traceback walk                   globaligner.py:395-593  (dp_array_backward)
  * tie table                        globaligner.py:595-685  (cost_ranks_dispatcher)
  * column emitters                  globaligner.py:688-753
  * cost -> score                    conclude.py:154-177     (final_cost_to_score)
  * score matrix -> cost matrix      start.py:500-557        (scoring_mat_to_costing_mat)

Levels: 0 = match/mismatch column, 1 = gap in seq_1 (consumes a seq_2 letter,
horizontal move), 2 = gap in seq_2 (consumes a seq_1 letter, vertical move).
Canonical tie order = the first option listed in the reference's table:
level 0 before level 1 before level 2 (globaligner.py:599, 605-607).
"""

from __future__ import annotations

import math
import random as _random

GAP = "-"

OP_DIAG = 0  # level 0
OP_GAP1 = 1  # level 1: '-' in seq_1 row
OP_GAP2 = 2  # level 2: '-' in seq_2 row


# --------------------------------------------------------------------------
# scheme helpers (start.py:488-497, 500-557; conclude.py:154-177)
# --------------------------------------------------------------------------
def max_entry(mat):
    """Largest value of a dict-of-dict matrix (start.py:488-497)."""
    return max(max(row.values()) for row in mat.values())


def score_to_cost_matrix(scoring_mat, max_score):
    """start.py:500-557 with the default deltas floor(b/2) / ceil(b/2)."""
    dd = math.floor(max_score / 2)
    di = math.ceil(max_score / 2)
    out = {}
    for x, row in scoring_mat.items():
        new_row = {}
        for y, s in row.items():
            if x == GAP and y != GAP:
                new_row[y] = dd - s          # horizontal step
            elif y == GAP and x != GAP:
                new_row[y] = di - s          # vertical step
            else:
                new_row[y] = dd + di - s     # (mis)match, and the '-','-' corner
        out[x] = new_row
    return out


def cost_to_score(cost, m, n, max_score):
    """conclude.py:154-177."""
    return n * math.floor(max_score / 2) + m * math.ceil(max_score / 2) - cost


# --------------------------------------------------------------------------
# fill (globaligner.py:756-821, 317-392)
# --------------------------------------------------------------------------
def boundary_table(seq_1, seq_2, cm, gap_open_cost):
    """(m+1) x (n+1) table with row 0 / column 0 set; interior cells are None."""
    m, n = len(seq_1), len(seq_2)
    sentinel = (max_entry(cm) + 1) * max(m, n)   # finite "infinity", :777
    table = [[None] * (n + 1) for _ in range(m + 1)]
    table[0][0] = (0, 0, 0)
    run = gap_open_cost
    for j in range(1, n + 1):
        run += cm[GAP][seq_2[j - 1]]
        table[0][j] = (sentinel, run, sentinel)
    run = gap_open_cost
    for i in range(1, m + 1):
        run += cm[seq_1[i - 1]][GAP]
        table[i][0] = (sentinel, sentinel, run)
    return table


def fill_in_place(table, seq_1, seq_2, cm, gap_open_cost):
    """Row-major sweep of the recurrence over a table whose boundary is set."""
    o = gap_open_cost
    gap_row = cm[GAP]
    for i in range(1, len(seq_1) + 1):
        a = seq_1[i - 1]
        sub_row = cm[a]
        vert = sub_row[GAP]
        above = table[i - 1]
        here = table[i]
        for j in range(1, len(seq_2) + 1):
            b = seq_2[j - 1]
            nw = above[j - 1]
            w = here[j - 1]
            nn = above[j]
            here[j] = (
                min(nw) + sub_row[b],
                min(w[0] + o, w[1], w[2] + o) + gap_row[b],
                min(nn[0] + o, nn[1] + o, nn[2]) + vert,
            )
    return table


def fill(seq_1, seq_2, cm, gap_open_cost):
    return fill_in_place(boundary_table(seq_1, seq_2, cm, gap_open_cost),
                         seq_1, seq_2, cm, gap_open_cost)


# --------------------------------------------------------------------------
# traceback (globaligner.py:395-685)
# --------------------------------------------------------------------------
def _candidates(cell, level, o):
    """Values ranked at a cell when the successor node sits on `level`.

    The reference also adds a step cost common to all three entries
    (globaligner.py:490-505); a common constant never changes integer ranks,
    so it is omitted.
    """
    if level == 0:
        return (cell[0], cell[1], cell[2])
    if level == 1:
        return (cell[0] + o, cell[1], cell[2] + o)
    return (cell[0] + o, cell[1] + o, cell[2])


def _candidates_exact(cell, level, o, cm, a, b, first):
    """The same for float-valued costs, where the common addend DOES matter: exactly the sums
    the reference forms (globaligner.py:490-510) -- nothing at the first evaluation (:425-435);
    on level 0 cm[a][b] on all three; on levels 1 and 2 cm["-"][b] (the "-" row on both), with
    gap_open_cost added to it first on the entries whose level differs."""
    if first:
        return tuple(cell)
    if level == 0:
        add = (cm[a][b],) * 3
    elif level == 1:
        add = (o + cm[GAP][b], cm[GAP][b], o + cm[GAP][b])
    else:
        add = (o + cm[GAP][b], o + cm[GAP][b], cm[GAP][b])
    return tuple(sum(cs) for cs in zip(cell, add))


def tied_minima(cell, level, o, exact=None):
    v = _candidates(cell, level, o) if exact is None else _candidates_exact(cell, level, o, *exact)
    lo = min(v)
    return tuple(k for k in range(3) if v[k] == lo)


def first_listed(options):
    """Canonical tie-break: the first option of the reference's table."""
    return options[0]


def walk(table, seq_1, seq_2, gap_open_cost, choose=first_listed, cm=None):
    """Return (ops, cost): ops in walk order (end of alignment first).

    `choose` receives the tuple of tied minimal levels (ascending) and picks
    one.  `cm`: pass the costing matrix when costs are floats (see _candidates_exact).
    Raises IndexError-like ValueError where the reference would leave the
    table (only possible outside the domain max(m, n) > gap_open_cost).
    """
    m, n = len(seq_1), len(seq_2)
    o = gap_open_cost
    i, j, level = m, n, 0
    cost = min(table[m][n])
    ops = []
    steps = 0
    while True:
        # One evaluation + move.  The very first one is the pre-loop block
        # (:425-458); later ones are loop iterations (:475-537).
        if (i == 0 and j == 0) or i < 0 or j < 0:
            raise ValueError("walk left the table (reference-undefined input)")
        # cm given: float-valued costs, rank the reference's exact sums (seq[i-1] wraps at i == 0
        # like Python's negative index)
        exact = None if cm is None else (cm, seq_1[i - 1], seq_2[j - 1], steps == 0)
        k = choose(tied_minima(table[i][j], level, o, exact))
        # A boundary cell can only be evaluated right after the first move
        # (when m == 1 or n == 1).  Inside the parity domain it always picks the
        # boundary-consistent level; anything else sends the reference to
        # negative (wrapping) indices.
        if (i == 0 and k != 1) or (j == 0 and k != 2):
            raise ValueError("walk left the table (reference-undefined input)")
        ops.append(k)
        if k == 0:
            i -= 1
            j -= 1
        elif k == 1:
            j -= 1
        else:
            i -= 1
        level = k
        steps += 1
        if steps == 1:
            if i == 0 and j == 0:          # :460
                break
            continue                       # no shortcut test after the first move
        if i == 0:                         # :542-561
            ops.extend([OP_GAP1] * j)
            break
        if j == 0:                         # :562-581
            ops.extend([OP_GAP2] * i)
            break
    return ops, cost


def render(ops, seq_1, seq_2):
    """Walk-order ops -> (seq_1_aligned, middle_part, seq_2_aligned) (:688-753, :584-593)."""
    i, j = len(seq_1), len(seq_2)
    top, mid, bot = [], [], []
    for k in ops:
        if k == OP_DIAG:
            i -= 1
            j -= 1
            top.append(seq_1[i])
            bot.append(seq_2[j])
            mid.append("|" if seq_1[i] == seq_2[j] else "*")
        elif k == OP_GAP1:
            j -= 1
            top.append(GAP)
            bot.append(seq_2[j])
            mid.append(" ")
        else:
            i -= 1
            top.append(seq_1[i])
            bot.append(GAP)
            mid.append(" ")
    assert i == 0 and j == 0, (i, j)
    top.reverse()
    mid.reverse()
    bot.reverse()
    return "".join(top), "".join(mid), "".join(bot)


def align(seq_1, seq_2, cm, gap_open_cost, choose=first_listed):
    """Full hot path: (seq_1_aligned, middle_part, seq_2_aligned, cost, ops)."""
    table = fill(seq_1, seq_2, cm, gap_open_cost)
    ops, cost = walk(table, seq_1, seq_2, gap_open_cost, choose)
    s1a, mid, s2a = render(ops, seq_1, seq_2)
    return s1a, mid, s2a, cost, ops


# --------------------------------------------------------------------------
# admissibility: is an alignment reachable by the *unpatched* reference?
# --------------------------------------------------------------------------
def admissible(ops, seq_1, seq_2, cm, gap_open_cost, table=None):
    """True iff every step of `ops` picks one of the tied minima the reference
    would choose among at that cell (admissibility check)."""
    if table is None:
        table = fill(seq_1, seq_2, cm, gap_open_cost)
    ops = list(ops)
    state = {"pos": 0, "ok": True}

    def replay(options):
        if state["pos"] >= len(ops):
            raise ValueError("ops too short")
        k = ops[state["pos"]]
        state["pos"] += 1
        if k not in options:
            state["ok"] = False
            return options[0]
        return k

    try:
        got, _ = walk(table, seq_1, seq_2, gap_open_cost, replay)
    except ValueError:
        return False
    return state["ok"] and list(got) == ops


# --------------------------------------------------------------------------
# seeded replay of the reference's random tie-breaking
# --------------------------------------------------------------------------
_DRAW_SIZES = (3, 2, 2, 2, 3, 2, 2, 2, 3)          # dict-literal order, one half
_LIVE_SLOT = {(0, 1, 2): 0, (0, 1): 1, (0, 2): 2, (1, 2): 3}


class SeededChooser:
    """Consumes a ``random.Random`` exactly like cost_ranks_dispatcher does:
    18 ``choice`` draws per call (9 for the is_match=True keys, then 9 for the
    False keys; globaligner.py:599-671), of which at most one is live."""

    def __init__(self, rng: _random.Random, seq_1, seq_2):
        self.rng = rng
        self.seq_1, self.seq_2 = seq_1, seq_2
        self.i, self.j = len(seq_1), len(seq_2)

    def __call__(self, options):
        is_match = self.seq_1[self.i - 1] == self.seq_2[self.j - 1]
        draws = []
        for _half in range(2):
            for size in _DRAW_SIZES:
                draws.append(self.rng.choice(range(size)))
        if len(options) == 1:
            k = options[0]
        else:
            slot = _LIVE_SLOT[tuple(options)] + (0 if is_match else 9)
            k = options[draws[slot]]
        if k == 0:
            self.i -= 1
            self.j -= 1
        elif k == 1:
            self.j -= 1
        else:
            self.i -= 1
        return k
