// Full value table and value-ranking walk for ONE pair (sm_100a): what dp_array_forward leaves in
// dp_array and what dp_array_backward does with it (reference globaligner.py:366-392, :395-593).
// Used by the seam functions, the seeded replay of the reference's random tie-breaking
//  and float-valued penalties (f2).  Not a throughput path.
#pragma once
#include "ga_common.cuh"

namespace ga {

// What dp_array_forward leaves in dp_array (reference globaligner.py:366-392), for ONE
// small pair; anti-diagonals in parallel, one block.  Not a throughput path.  Templated on the
// value type: int32 for the reference's own (int()-coerced) entry point, f64 for float-valued
// penalties, which the reference accepts when its inner functions are called directly (type
// hints globaligner.py:371,761).  The f64 instance adds and compares in the reference's own
// order -- min of the three predecessors (with gap_open added where levels differ), then the
// step cost (:359-363) -- so its doubles are bit-identical to CPython's.
struct IntCosts {
    SchemeDev sc;
    typedef int value_t;
    typedef long long acc_t;
    static constexpr bool exact_adds = false;
    __device__ int sub(int a, int b) const { return sc.sub4[a * sc.A + b] / 4; }
    __device__ int gh(int b) const { return sc.gh_raw[b]; }
    __device__ int gv(int a) const { return sc.gv_raw[a]; }
    __device__ int open() const { return sc.o; }
    __device__ int big(int m, int n) const { return (sc.max_cost + 1) * max(m, n); }
};
struct F64Costs {
    int A;
    const double *s, *h, *v;
    double o, max_cost;
    typedef double value_t;
    typedef double acc_t;
    static constexpr bool exact_adds = true;   // floating point: a common addend can change ranks
    __device__ double sub(int a, int b) const { return s[a * A + b]; }
    __device__ double gh(int b) const { return h[b]; }
    __device__ double gv(int a) const { return v[a]; }
    __device__ double open() const { return o; }
    __device__ double big(int m, int n) const { return (max_cost + 1.0) * (double)max(m, n); }
};

template <typename C>
__global__ void values_kernel(const C c, const uint8_t *s1, int m, const uint8_t *s2, int n,
                              int keep_boundary, typename C::value_t *tab) {
    typedef typename C::value_t T;
    const int W = n + 1;
    auto at = [&](int i, int j, int k) -> T & { return tab[((size_t)i * W + j) * 3 + k]; };
    if (!keep_boundary) {
        if (threadIdx.x == 0) {
            const T big = c.big(m, n);
            at(0, 0, 0) = at(0, 0, 1) = at(0, 0, 2) = 0;
            T run = c.open();
            for (int j = 1; j <= n; ++j) { run += c.gh(s2[j - 1]); at(0, j, 0) = big; at(0, j, 1) = run; at(0, j, 2) = big; }
            run = c.open();
            for (int i = 1; i <= m; ++i) { run += c.gv(s1[i - 1]); at(i, 0, 0) = big; at(i, 0, 1) = big; at(i, 0, 2) = run; }
        }
        __syncthreads();
    }
    const T o = c.open();
    for (int d = 2; d <= m + n; ++d) {
        const int ilo = max(1, d - n), ihi = min(m, d - 1);
        for (int i = ilo + threadIdx.x; i <= ihi; i += blockDim.x) {
            const int j = d - i;
            const int a = s1[i - 1], b = s2[j - 1];
            at(i, j, 0) = min(min(at(i - 1, j - 1, 0), at(i - 1, j - 1, 1)), at(i - 1, j - 1, 2)) + c.sub(a, b);
            at(i, j, 1) = min(min(at(i, j - 1, 0) + o, at(i, j - 1, 1)), at(i, j - 1, 2) + o) + c.gh(b);
            at(i, j, 2) = min(min(at(i - 1, j, 0) + o, at(i - 1, j, 1) + o), at(i - 1, j, 2)) + c.gv(a);
        }
        __syncthreads();
    }
}

// The reference breaks traceback ties with random.choice (globaligner.py:595-685): every
// dispatcher call draws 18 choices (9 for the is_match=True keys, 9 for the False keys), of
// which at most one is live.  The host emulates CPython's draws and passes, per evaluated
// cell, eight 2-bit option indices: bits 0-7 for is_match=True, bits 8-15 for False; within a
// byte: all three tied (options 0,1,2) | levels 0,1 tied (0,1) | 0,2 tied (0,2) | 1,2 tied (1,2).
// One thread walks the full value table (ga_fill_values' kernel) ranking values as
// dp_array_backward does (:425-581).  Single pairs only; not a throughput path.
// `draws` == nullptr: canonical order (first-listed option of every tie).
//
// What is ranked (:490-514): the cell's three values plus, per level, the cost of the step to the
// successor -- on level 0 cm[a][b] for all three, on levels 1 and 2 cm["-"][b] (the reference
// uses the "-" ROW on both gap levels) with gap_open_cost added FIRST on the levels that differ
// from the successor's -- and nothing at the first evaluation (:425-435).  A common addend
// cannot change integer ranks, so the integer instance adds gap_open only; the f64 instance
// adds exactly what the reference adds, in its order: value + (gap_open + step).
template <typename C>
__global__ void replay_walk_kernel(const C cs, const uint8_t *s1, int m, const uint8_t *s2, int n,
                                   const typename C::value_t *tab, const uint16_t *draws, uint32_t *ops_out,
                                   long long *n_ops_out, int *n_eval_out, int *status_out,
                                   typename C::value_t *cost_out) {
    typedef typename C::value_t T;
    typedef typename C::acc_t ACC;
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    const int W = n + 1;
    const ACC o = cs.open();
    long long nops = 0;
    uint32_t word = 0;
    int n_eval = 0, status = 0;
    auto emit = [&](int op) {
        word |= (uint32_t)op << (2 * (int)(nops & 15));
        if ((++nops & 15) == 0) { ops_out[(nops >> 4) - 1] = word; word = 0; }
    };
    const T *corner = tab + ((size_t)m * W + n) * 3;
    *cost_out = min(min(corner[0], corner[1]), corner[2]);
    int i = m, j = n, level = 0;
    while (status == 0) {
        if (i == 0 && j == 0) break;                       // :460 and the loop's natural end
        // the shortcut tests (:542-581) follow every move except the first one, so a boundary
        // cell reached by the very first move is still ranked once (18 more draws)
        if (n_eval >= 2 && (i == 0 || j == 0)) break;
        const T *c = tab + ((size_t)i * W + j) * 3;
        // Python's negative index wraps: seq[-1] is the last letter (only at a boundary cell)
        const int ai = i >= 1 ? i - 1 : m - 1, bj = j >= 1 ? j - 1 : n - 1;
        ACC v0 = c[0], v1 = c[1], v2 = c[2];
        if (C::exact_adds) {
            if (n_eval > 0) {
                if (level == 0) {
                    const ACC st = cs.sub(s1[ai], s2[bj]);
                    v0 += st; v1 += st; v2 += st;
                } else {
                    const ACC g = cs.gh(s2[bj]);
                    const ACC og = o + g;
                    if (level == 1) { v0 += og; v1 += g; v2 += og; } else { v0 += og; v1 += og; v2 += g; }
                }
            }
        } else {
            if (level == 1) { v0 += o; v2 += o; } else if (level == 2) { v0 += o; v1 += o; }
        }
        const ACC lo = min(min(v0, v1), v2);
        const int t0 = v0 == lo, t1 = v1 == lo, t2 = v2 == lo;
        const int is_match = s1[ai] == s2[bj];
        const unsigned d = draws ? draws[n_eval] >> (is_match ? 0 : 8) : 0u;
        ++n_eval;
        int k;
        if (t0 && t1 && t2) k = (int)(d & 3u);                       // options (0, 1, 2)
        else if (t0 && t1) k = (int)((d >> 2) & 3u);                  // (0, 1)
        else if (t0 && t2) k = ((d >> 4) & 3u) ? 2 : 0;               // (0, 2)
        else if (t1 && t2) k = ((d >> 6) & 3u) ? 2 : 1;               // (1, 2)
        else k = t0 ? 0 : (t1 ? 1 : 2);
        if ((i == 0 && k != 1) || (j == 0 && k != 2)) { status = 1; break; }   // reference leaves the table
        emit(k);
        if (k == 0) { --i; --j; } else if (k == 1) { --j; } else { --i; }
        level = k;
    }
    if (status == 0) {
        for (; j > 0 && i == 0; --j) emit(1);
        for (; i > 0 && j == 0; --i) emit(2);
    }
    if (nops & 15) ops_out[nops >> 4] = word;
    *n_ops_out = nops;
    *n_eval_out = n_eval;
    *status_out = status;
}

}  // namespace ga
