/*
 * TEST INFRASTRUCTURE ONLY -- plain-C restatement of globalign's DP hot path.
 *
 * Oracle for the CUDA path at sizes the pure-Python restatement
 * (oracle/py_oracle.py) cannot reach.  Nothing in the product package links,
 * loads or executes this file; it is built into oracle/_build/libga_oracle.so
 * by oracle/Makefile and used only by tests/, __graft_entry__.smoke() and the
 * CPU-baseline legs of bench.py.
 *
 * Pinning status: PINNED -- tests/test_oracle.py checks it against the golden
 * vectors frozen from the real reference (tests/golden/), and
 * tests/golden/make_golden.py checks it against the live reference in the
 * build container.
 *
 * Reference lines restated (under /root/reference/src/globalign):
 *   boundary   globaligner.py:756-821      cell   globaligner.py:317-363
 *   sweep      globaligner.py:366-392      walk   globaligner.py:395-593
 *   tie table  globaligner.py:595-685 (canonical = first listed option:
 *              level 0, then level 1, then level 2)
 *
 * Inputs are "lowered": letters are small integer codes, the cost matrix is
 * dense.  sub[a*A+b] = cm[a][b], gap_h[b] = cm["-"][b] (horizontal step,
 * level 1), gap_v[a] = cm[a]["-"] (vertical step, level 2), o = gap_open_cost,
 * max_cost = largest entry of cm (for the reference's finite sentinel).
 *
 * ops are 0 (diagonal), 1 (gap in seq_1), 2 (gap in seq_2), in walk order
 * (last alignment column first), one byte each.
 */
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct { int32_t v[3]; } cell_t;

static inline int32_t min2(int32_t a, int32_t b) { return a < b ? a : b; }
static inline int32_t min3(int32_t a, int32_t b, int32_t c) { return min2(min2(a, b), c); }

static inline int32_t sentinel_of(int64_t m, int64_t n, int32_t max_cost) {
    return (int32_t)((int64_t)(max_cost + 1) * (m > n ? m : n));   /* :777 */
}

/* Row 0 and column 0 exactly as make_dp_array writes them (:778-819). */
static void set_boundary(cell_t *t, int64_t m, int64_t n, const uint8_t *s1, const uint8_t *s2,
                         const int32_t *gap_h, const int32_t *gap_v, int32_t o, int32_t max_cost) {
    const int64_t W = n + 1;
    const int32_t big = sentinel_of(m, n, max_cost);
    t[0].v[0] = t[0].v[1] = t[0].v[2] = 0;
    int32_t run = o;
    for (int64_t j = 1; j <= n; ++j) {
        run += gap_h[s2[j - 1]];
        t[j].v[0] = big; t[j].v[1] = run; t[j].v[2] = big;
    }
    run = o;
    for (int64_t i = 1; i <= m; ++i) {
        run += gap_v[s1[i - 1]];
        t[i * W].v[0] = big; t[i * W].v[1] = big; t[i * W].v[2] = run;
    }
}

/* Interior sweep (:380-390 calling :331-363).  Boundary must already be set. */
static void sweep(cell_t *t, int64_t m, int64_t n, const uint8_t *s1, const uint8_t *s2, int A,
                  const int32_t *sub, const int32_t *gap_h, const int32_t *gap_v, int32_t o) {
    const int64_t W = n + 1;
    for (int64_t i = 1; i <= m; ++i) {
        const int a = s1[i - 1];
        const int32_t *srow = sub + (int64_t)a * A;
        const int32_t gv = gap_v[a];
        const cell_t *up = t + (i - 1) * W;
        cell_t *cur = t + i * W;
        for (int64_t j = 1; j <= n; ++j) {
            const int b = s2[j - 1];
            const cell_t nw = up[j - 1], w = cur[j - 1], nn = up[j];
            cur[j].v[0] = min3(nw.v[0], nw.v[1], nw.v[2]) + srow[b];
            cur[j].v[1] = min3(w.v[0] + o, w.v[1], w.v[2] + o) + gap_h[b];
            cur[j].v[2] = min3(nn.v[0] + o, nn.v[1] + o, nn.v[2]) + gv;
        }
    }
}

/* First-listed minimum of the three candidates ranked when the successor node
 * is on `level` (:490-514 with the common step cost dropped; :599-607). */
static inline int pick(const cell_t *c, int level, int32_t o) {
    int32_t v0 = c->v[0], v1 = c->v[1], v2 = c->v[2];
    if (level == 1) { v0 += o; v2 += o; }
    else if (level == 2) { v0 += o; v1 += o; }
    int k = 0; int32_t best = v0;
    if (v1 < best) { best = v1; k = 1; }
    if (v2 < best) { k = 2; }
    return k;
}

/* Value-comparing walk (:425-593).  Returns 0, or 1 if the reference would
 * step off the table (only possible outside max(m,n) > gap_open_cost). */
static int walk(const cell_t *t, int64_t m, int64_t n, int32_t o, uint8_t *ops, int64_t *nops,
                int32_t *cost) {
    const int64_t W = n + 1;
    int64_t i = m, j = n, k_out = 0, steps = 0;
    int level = 0;
    *cost = min3(t[m * W + n].v[0], t[m * W + n].v[1], t[m * W + n].v[2]);
    for (;;) {
        if ((i == 0 && j == 0) || i < 0 || j < 0) return 1;
        const int k = pick(&t[i * W + j], level, o);
        if ((i == 0 && k != 1) || (j == 0 && k != 2)) return 1;
        ops[k_out++] = (uint8_t)k;
        if (k == 0) { --i; --j; } else if (k == 1) { --j; } else { --i; }
        level = k;
        if (++steps == 1) {                 /* :460 -- only (0,0) ends the walk here */
            if (i == 0 && j == 0) break;
            continue;
        }
        if (i == 0) { while (j-- > 0) ops[k_out++] = 1; break; }   /* :542-561 */
        if (j == 0) { while (i-- > 0) ops[k_out++] = 2; break; }   /* :562-581 */
    }
    *nops = k_out;
    return 0;
}

/* ---- exported -------------------------------------------------------- */

/* Fill a caller-provided (m+1)x(n+1)x3 int32 table.  If keep_boundary != 0 the
 * caller's row 0 / column 0 are used as they are (the reference's own
 * dp_array_forward test passes hand-written boundary values). */
int gao_fill_table(int64_t m, int64_t n, const uint8_t *s1, const uint8_t *s2, int A,
                   const int32_t *sub, const int32_t *gap_h, const int32_t *gap_v, int32_t o,
                   int32_t max_cost, int keep_boundary, int32_t *table) {
    cell_t *t = (cell_t *)table;
    if (!keep_boundary) set_boundary(t, m, n, s1, s2, gap_h, gap_v, o, max_cost);
    sweep(t, m, n, s1, s2, A, sub, gap_h, gap_v, o);
    return 0;
}

/* Whole hot path for one pair.  ops must hold m+n bytes. */
int gao_align(int64_t m, int64_t n, const uint8_t *s1, const uint8_t *s2, int A,
              const int32_t *sub, const int32_t *gap_h, const int32_t *gap_v, int32_t o,
              int32_t max_cost, uint8_t *ops, int64_t *nops, int32_t *cost) {
    cell_t *t = (cell_t *)malloc((size_t)(m + 1) * (size_t)(n + 1) * sizeof(cell_t));
    if (!t) return -1;
    set_boundary(t, m, n, s1, s2, gap_h, gap_v, o, max_cost);
    sweep(t, m, n, s1, s2, A, sub, gap_h, gap_v, o);
    const int rc = walk(t, m, n, o, ops, nops, cost);
    free(t);
    return rc;
}

/* Batch driver over pthreads (CPU baseline).  Offsets are in letters;
 * pair p owns ops[ops_off[p] .. ops_off[p] + m_p + n_p). */
typedef struct {
    int64_t lo, hi;
    const uint8_t *s1; const int64_t *off1; const uint8_t *s2; const int64_t *off2;
    int A; const int32_t *sub, *gap_h, *gap_v; int32_t o, max_cost;
    uint8_t *ops; const int64_t *ops_off; int64_t *nops; int32_t *cost; int32_t *status;
} job_t;

static void *job_main(void *arg) {
    job_t *jb = (job_t *)arg;
    for (int64_t p = jb->lo; p < jb->hi; ++p) {
        const int64_t m = jb->off1[p + 1] - jb->off1[p], n = jb->off2[p + 1] - jb->off2[p];
        jb->status[p] = gao_align(m, n, jb->s1 + jb->off1[p], jb->s2 + jb->off2[p], jb->A, jb->sub,
                                  jb->gap_h, jb->gap_v, jb->o, jb->max_cost,
                                  jb->ops + jb->ops_off[p], &jb->nops[p], &jb->cost[p]);
    }
    return NULL;
}

int gao_align_batch(int64_t n_pairs, const uint8_t *s1, const int64_t *off1, const uint8_t *s2,
                    const int64_t *off2, int A, const int32_t *sub, const int32_t *gap_h,
                    const int32_t *gap_v, int32_t o, int32_t max_cost, uint8_t *ops,
                    const int64_t *ops_off, int64_t *nops, int32_t *cost, int32_t *status,
                    int n_threads) {
    if (n_threads < 1) n_threads = 1;
    if (n_threads > 256) n_threads = 256;
    pthread_t th[256];
    job_t jobs[256];
    for (int k = 0; k < n_threads; ++k) {
        job_t jb = { n_pairs * k / n_threads, n_pairs * (k + 1) / n_threads, s1, off1, s2, off2, A,
                     sub, gap_h, gap_v, o, max_cost, ops, ops_off, nops, cost, status };
        jobs[k] = jb;
        if (pthread_create(&th[k], NULL, job_main, &jobs[k]) != 0) return -1;
    }
    for (int k = 0; k < n_threads; ++k) pthread_join(th[k], NULL);
    return 0;
}

/* Linear-memory check that `ops` (walk order) is exactly the canonical walk.
 *
 * The canonical walk only inspects the cells it passes through, so one
 * row-by-row sweep that keeps two rows and records the three values at the
 * cells named by `ops` is enough to re-decide every step.  Returns 0 when ops
 * is the canonical walk and *cost its cost, 2 when some step differs, 1 for
 * reference-undefined input, -1 on allocation failure. */
int gao_verify_path(int64_t m, int64_t n, const uint8_t *s1, const uint8_t *s2, int A,
                    const int32_t *sub, const int32_t *gap_h, const int32_t *gap_v, int32_t o,
                    int32_t max_cost, const uint8_t *ops, int64_t nops, int32_t *cost) {
    /* cells visited by ops, in walk order */
    int64_t *pi = (int64_t *)malloc(sizeof(int64_t) * (size_t)(nops + 1));
    int64_t *pj = (int64_t *)malloc(sizeof(int64_t) * (size_t)(nops + 1));
    cell_t *pv = (cell_t *)malloc(sizeof(cell_t) * (size_t)(nops + 1));
    cell_t *rows = (cell_t *)malloc(sizeof(cell_t) * 2 * (size_t)(n + 1));
    if (!pi || !pj || !pv || !rows) { free(pi); free(pj); free(pv); free(rows); return -1; }
    int64_t i = m, j = n;
    int rc = 0;
    for (int64_t s = 0; s < nops; ++s) {
        pi[s] = i; pj[s] = j;
        if (ops[s] == 0) { --i; --j; } else if (ops[s] == 1) { --j; } else { --i; }
        if (i < 0 || j < 0) { rc = 2; break; }
    }
    if (rc == 0 && (i != 0 || j != 0)) rc = 2;
    if (rc == 0) {
        const int32_t big = sentinel_of(m, n, max_cost);
        cell_t *prev = rows, *cur = rows + (n + 1);
        prev[0].v[0] = prev[0].v[1] = prev[0].v[2] = 0;
        int32_t run = o;
        for (int64_t c = 1; c <= n; ++c) {
            run += gap_h[s2[c - 1]];
            prev[c].v[0] = big; prev[c].v[1] = run; prev[c].v[2] = big;
        }
        /* path cells are visited with non-increasing row index: consume from the tail */
        int64_t s = nops - 1;
        while (s >= 0 && pi[s] == 0) { pv[s] = prev[pj[s]]; --s; }
        int32_t col = o;
        for (int64_t r = 1; r <= m; ++r) {
            const int a = s1[r - 1];
            const int32_t *srow = sub + (int64_t)a * A;
            const int32_t gv = gap_v[a];
            col += gv;
            cur[0].v[0] = big; cur[0].v[1] = big; cur[0].v[2] = col;
            for (int64_t c = 1; c <= n; ++c) {
                const int b = s2[c - 1];
                const cell_t nw = prev[c - 1], w = cur[c - 1], nn = prev[c];
                cur[c].v[0] = min3(nw.v[0], nw.v[1], nw.v[2]) + srow[b];
                cur[c].v[1] = min3(w.v[0] + o, w.v[1], w.v[2] + o) + gap_h[b];
                cur[c].v[2] = min3(nn.v[0] + o, nn.v[1] + o, nn.v[2]) + gv;
            }
            while (s >= 0 && pi[s] == r) { pv[s] = cur[pj[s]]; --s; }
            cell_t *tmp = prev; prev = cur; cur = tmp;
        }
        *cost = min3(prev[n].v[0], prev[n].v[1], prev[n].v[2]);
        /* re-decide every step exactly like walk() */
        int level = 0; int64_t steps = 0, k_at = 0;
        i = m; j = n;
        for (;;) {
            if (i == 0 && j == 0) { rc = 1; break; }
            if (k_at >= nops) { rc = 2; break; }
            const int k = pick(&pv[k_at], level, o);
            if ((i == 0 && k != 1) || (j == 0 && k != 2)) { rc = 1; break; }
            if (k != ops[k_at]) { rc = 2; break; }
            ++k_at;
            if (k == 0) { --i; --j; } else if (k == 1) { --j; } else { --i; }
            level = k;
            if (++steps == 1) { if (i == 0 && j == 0) break; continue; }
            if (i == 0) {
                for (; j > 0; --j) { if (k_at >= nops || ops[k_at] != 1) { rc = 2; break; } ++k_at; }
                break;
            }
            if (j == 0) {
                for (; i > 0; --i) { if (k_at >= nops || ops[k_at] != 2) { rc = 2; break; } ++k_at; }
                break;
            }
        }
        if (rc == 0 && k_at != nops) rc = 2;
    }
    free(pi); free(pj); free(pv); free(rows);
    return rc;
}

/* Score-only, linear memory: the final cost (for very long pairs). */
int gao_cost_only(int64_t m, int64_t n, const uint8_t *s1, const uint8_t *s2, int A,
                  const int32_t *sub, const int32_t *gap_h, const int32_t *gap_v, int32_t o,
                  int32_t max_cost, int32_t *cost) {
    cell_t *rows = (cell_t *)malloc(sizeof(cell_t) * 2 * (size_t)(n + 1));
    if (!rows) return -1;
    const int32_t big = sentinel_of(m, n, max_cost);
    cell_t *prev = rows, *cur = rows + (n + 1);
    prev[0].v[0] = prev[0].v[1] = prev[0].v[2] = 0;
    int32_t run = o;
    for (int64_t c = 1; c <= n; ++c) {
        run += gap_h[s2[c - 1]];
        prev[c].v[0] = big; prev[c].v[1] = run; prev[c].v[2] = big;
    }
    int32_t col = o;
    for (int64_t r = 1; r <= m; ++r) {
        const int a = s1[r - 1];
        const int32_t *srow = sub + (int64_t)a * A;
        const int32_t gv = gap_v[a];
        col += gv;
        cur[0].v[0] = big; cur[0].v[1] = big; cur[0].v[2] = col;
        for (int64_t c = 1; c <= n; ++c) {
            const int b = s2[c - 1];
            const cell_t nw = prev[c - 1], w = cur[c - 1], nn = prev[c];
            cur[c].v[0] = min3(nw.v[0], nw.v[1], nw.v[2]) + srow[b];
            cur[c].v[1] = min3(w.v[0] + o, w.v[1], w.v[2] + o) + gap_h[b];
            cur[c].v[2] = min3(nn.v[0] + o, nn.v[1] + o, nn.v[2]) + gv;
        }
        cell_t *tmp = prev; prev = cur; cur = tmp;
    }
    *cost = min3(prev[n].v[0], prev[n].v[1], prev[n].v[2]);
    free(rows);
    return 0;
}
