// K2 + K5: one long pair on the whole chip (sm_100a).
//
// K2, long_fill_kernel -- fill without storing direction codes.  The DP matrix is cut
// into bands of LBAND = 32 lanes x LR rows; one warp sweeps one band with the same
// anti-diagonal wavefront as the batched kernel (ga_fill.cuh: lane g owns LR rows and
// works on column t-g at step t; h/U of its bottom row go to lane g+1 by shuffle).
// (LR = 4 rows per lane: 128-row bands.)  Band k reads the bottom row of band k-1 (h, U per column) from global memory, 32
// columns at a time, behind a progress counter that band k-1 publishes with a
// release fence -- a systolic pipeline of bands across all SMs.  Bands are claimed in
// increasing order from an atomic counter, so a band's predecessor is always running
// or finished: no deadlock, no co-residency requirement.  Kept for the traceback:
//   * brow[k][j]      (h, U) of the last row of band k, every column  (8 B / column / band)
//   * vchk[k][c][...] (h, L) of every row of band k at column c*LCC   (8 B / row / checkpoint)
// i.e. 8/LBAND + 8/LCC = 0.09 bytes per cell instead of the 0.5 bytes per cell of stored
// direction codes (0.94 GB instead of 5 GB for 100 kb x 100 kb); taller bands and sparser
// checkpoints shrink it further at the price of more recomputation in K5.
//
// K5, long_trace_kernel -- linear-space traceback by checkpoint + recompute (the
// reference has no such mode; SURVEY.md 7.3 option (a)).  One warp repeats: re-fill,
// WITH direction codes into shared memory, the tile (band of the current row) x (the
// <= LCC columns left of the current column) from its two checkpoints; walk it with the
// same decoding as ga_walk.cuh until the path leaves the tile.  The path crosses about
// m/LBAND + n/LCC tiles, so the recomputation is O((m+n) * LBAND) cells.  Because the
// recomputed values are the forward values, the walk is exactly the canonical walk of
// the full matrix (same ties), which a Hirschberg split would not give for free.
//
// Replaces, for one pair too large for a stored code matrix: make_dp_array
// (globaligner.py:756-821), dp_array_forward (:366-392), dp_array_backward (:395-593).
#pragma once
#include "ga_common.cuh"
#include "ga_fill.cuh"

namespace ga {

constexpr int LR = 4;               // rows per lane (more, shorter bands: the pipeline is latency-bound)
constexpr int LBAND = 32 * LR;      // rows per band
constexpr int LCC = 128;            // columns between vertical checkpoints
constexpr int LSYNC = 16;           // columns per boundary-row fetch (sets the lag between bands)

struct LongArgs {
    const uint8_t *s1, *s2;         // device letters
    int m, n;
    const int *colsum;              // m+1: o + sum_{t<=i} gap_v[s1[t-1]]  (column-0 level-2 values), [0] = o
    const int *rowsum;              // n+1: o + sum_{t<=j} gap_h[s2[t-1]]  (row-0 level-1 values)
    int2 *brow;                     // nbands x (n+1)
    int2 *vchk;                     // nbands x nck x 32 x LR
    int *band_counter;
    int nbands, nck;
    unsigned big;                   // reference sentinel (max_cost+1)*max(m,n)
    int *out_cost;
    int debug_nosync;               // timing experiments only: do not wait for the band above
};

template <bool SIMPLE>
struct LongLane {
    unsigned aw[LR], gvT[LR], hl[LR], Lp[LR];
};

__device__ __forceinline__ int2 ld_volatile_int2(const int2 *p) {
    int2 v;
    asm volatile("ld.volatile.global.v2.s32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(p));
    return v;
}

// letters / per-row constants of the LR rows starting at DP row ib+1
template <bool SIMPLE>
__device__ __forceinline__ void long_load_rows(const SchemeDev &sc, const LongArgs &a, int ib,
                                               LongLane<SIMPLE> &ln) {
#pragma unroll
    for (int r = 0; r < LR; ++r) {
        const int row = ib + r;
        if (SIMPLE) {
            ln.aw[r] = row < a.m ? (unsigned)a.s1[row] : (unsigned)PAD_A;
            ln.gvT[r] = 4u * sc.gv + 2u;
        } else {
            const int x = row < a.m ? (int)a.s1[row] : sc.A;
            ln.aw[r] = (unsigned)(x * sc.A);
            ln.gvT[r] = (unsigned)sc.gv4t[x];
        }
    }
}

// one column of LR cells for this lane; returns (hb, ub) through references
template <bool SIMPLE, bool CODES>
__device__ __forceinline__ void long_cells(const SchemeDev &sc, const int *s_sub, LongLane<SIMPLE> &ln,
                                           unsigned &diag_io, unsigned up, unsigned bv, unsigned ghT,
                                           unsigned o4, unsigned mm4, unsigned &hb, unsigned &ub,
                                           unsigned &w0, unsigned &w1) {
    unsigned diag = diag_io;
    w0 = 0u;
    w1 = 0u;
#pragma unroll
    for (int r = 0; r < LR; ++r) {
        unsigned d0;
        if (SIMPLE) d0 = diag + (ln.aw[r] != bv ? mm4 : 0u);
        else d0 = diag + (unsigned)s_sub[ln.aw[r] + bv];
        const unsigned d1 = ln.Lp[r] + ghT;
        const unsigned d2 = up + ln.gvT[r];
        const unsigned hq = (unsigned)__vimin3_s32((int)d0, (int)d1, (int)d2);
        const unsigned Lq = (unsigned)__viaddmin_s32((int)hq, (int)o4, (int)d1);
        const unsigned Uq = (unsigned)__viaddmin_s32((int)hq, (int)o4, (int)d2);
        if (CODES) {
            w0 = (w0 << 2) | (hq & 3u);
            w1 = (w1 << 2) | ((Lq & 1u) | (Uq & 2u));
        }
        diag = ln.hl[r];
        ln.hl[r] = hq & ~3u;
        ln.Lp[r] = Lq & ~3u;
        up = Uq & ~3u;
    }
    hb = ln.hl[LR - 1];
    ub = up;
}

// ---------------------------------------------------------------------------- K2
constexpr int LWARPS = 4;           // independent band-sweeping warps per block (one per scheduler)

template <bool SIMPLE>
__global__ void __launch_bounds__(32 * LWARPS)
long_fill_kernel(const SchemeDev sc, const LongArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    int *s_sub = reinterpret_cast<int *>(smem_raw);
    int *s_gh = s_sub;
    if (!SIMPLE) {
        const int n_sub = (sc.A + 1) * sc.A;
        s_gh = s_sub + n_sub;
        for (int k = threadIdx.x; k < n_sub; k += blockDim.x) s_sub[k] = sc.sub4[k];
        for (int k = threadIdx.x; k < sc.A; k += blockDim.x) s_gh[k] = sc.gh4t[k];
        __syncthreads();
    }
    const int g = threadIdx.x & 31;     // the warps of a block never synchronise with each other again
    const unsigned o4 = 4u * sc.o, mm4 = 4u * sc.mm;
    const unsigned ghTc = 4u * sc.gh + 1u;
    const unsigned big = a.big;
    // hot fields in registers (kernel parameters live in constant memory; an LDC inside
    // the step loop sits on the critical path of a lone warp)
    const int N = a.n, M = a.m;
    const uint8_t *const S2 = a.s2;
    const int *const ROWSUM = a.rowsum;
    const unsigned O = (unsigned)sc.o;
    const int nosync = a.debug_nosync;
    for (;;) {
        int band = 0;
        if (g == 0) band = atomicAdd(a.band_counter, 1);
        band = __shfl_sync(FULL, band, 0);
        if (band >= a.nbands) break;
        const int ib = band * LBAND + g * LR;
        LongLane<SIMPLE> ln;
        long_load_rows<SIMPLE>(sc, a, ib, ln);
#pragma unroll
        for (int r = 0; r < LR; ++r) {
            const unsigned c = (unsigned)a.colsum[min(ib + r + 1, a.m)];
            ln.hl[r] = min(big, c) << 2;
            ln.Lp[r] = min(big, c + sc.o) << 2;
        }
        unsigned rh_prev = 0;
        if (ib != 0) rh_prev = min(big, (unsigned)a.colsum[min(ib, a.m)]) << 2;
        const int pm = (M - 1) / LBAND, gm = ((M - 1) % LBAND) / LR, rm = (M - 1) % LR;
        int2 *my_brow = a.brow + (size_t)band * (N + 1);
        const int2 *up_brow = a.brow + (size_t)(band > 0 ? band - 1 : 0) * (N + 1);
        int2 *my_vchk = a.vchk + ((size_t)band * a.nck) * (32 * LR) + (size_t)g * LR;

        unsigned hb = 0, ub = 0, bv = 0;
        unsigned bblk = 0;
        int2 tblk = make_int2(0, 0);
        const int nsteps = N + 31;
        for (int t = 0; t < nsteps; ++t) {
            if ((t & (LSYNC - 1)) == 0) {
                // next 32 columns: letters of seq_2 and the boundary row above this band
                const int c0 = t + 1;                       // first column of the block (1-based)
                const int cj = c0 + (g & (LSYNC - 1));      // lanes >= LSYNC duplicate a valid column
                bblk = cj <= N ? (unsigned)S2[cj - 1] : (SIMPLE ? (unsigned)PAD_B : 0u);
                if (band > 0 && c0 <= N) {
                    if (cj <= N) {
                        do { tblk = ld_volatile_int2(up_brow + cj); } while (tblk.x < 0 && !nosync);   // -1 = not yet written
                    }
                    __syncwarp();
                } else if (band == 0) {
                    const unsigned rj = (unsigned)ROWSUM[min(cj, N)];
                    tblk = make_int2((int)(min(big, rj) << 2), (int)(min(big, rj + O) << 2));
                }
            }
            const int j = t - g + 1;
            unsigned rh = __shfl_up_sync(FULL, hb, 1);
            unsigned ru = __shfl_up_sync(FULL, ub, 1);
            unsigned bin = __shfl_up_sync(FULL, bv, 1);
            const unsigned b0 = __shfl_sync(FULL, bblk, t & (LSYNC - 1));
            const unsigned th = (unsigned)__shfl_sync(FULL, tblk.x, t & (LSYNC - 1));
            const unsigned tu = (unsigned)__shfl_sync(FULL, tblk.y, t & (LSYNC - 1));
            if (g == 0) { bin = b0; rh = th; ru = tu; }
            bv = bin;
            if (j >= 1 && j <= N) {
                unsigned diag = rh_prev;
                rh_prev = rh;
                const unsigned ghT = SIMPLE ? ghTc : (unsigned)s_gh[bv];
                unsigned w0, w1;
                long_cells<SIMPLE, false>(sc, s_sub, ln, diag, ru, bv, ghT, o4, mm4, hb, ub, w0, w1);
                if (g == 31) __stcg(my_brow + j, make_int2((int)hb, (int)ub));   // one 64-bit store
                if ((j & (LCC - 1)) == 0) {
                    int2 *dst = my_vchk + (size_t)(j / LCC) * (32 * LR);
#pragma unroll
                    for (int r = 0; r < LR; ++r) __stcg(dst + r, make_int2((int)ln.hl[r], (int)ln.Lp[r]));
                }
                if (j == N && band == pm && g == gm) {
                    unsigned v = 0;
#pragma unroll
                    for (int r = 0; r < LR; ++r)
                        if (r == rm) v = ln.hl[r];
                    *a.out_cost = (int)(v >> 2);
                }
            }
        }
    }
}

// ---------------------------------------------------------------------------- K5
constexpr int LT_NTB = (LCC + 31 + 1) / 2;          // store blocks of 2 steps (2 words per step)

struct TraceOut {
    uint32_t *ops;                  // packed 2-bit ops, walk order
    long long *n_ops;
    int *status;
    int *tiles_refilled;
};

template <bool SIMPLE>
__global__ void __launch_bounds__(32)
long_trace_kernel(const SchemeDev sc, const LongArgs a, const TraceOut out) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    // layout: [tile codes: LT_NTB x 32 lanes x uint4][TABLE: sub | gh]
    uint4 *s_codes = reinterpret_cast<uint4 *>(smem_raw);
    int *s_sub = reinterpret_cast<int *>(smem_raw + (size_t)LT_NTB * 32 * sizeof(uint4));
    int *s_gh = s_sub;
    if (!SIMPLE) {
        const int n_sub = (sc.A + 1) * sc.A;
        s_gh = s_sub + n_sub;
        for (int k = threadIdx.x; k < n_sub; k += 32) s_sub[k] = sc.sub4[k];
        for (int k = threadIdx.x; k < sc.A; k += 32) s_gh[k] = sc.gh4t[k];
        __syncwarp();
    }
    const int g = threadIdx.x;
    const unsigned o4 = 4u * sc.o, mm4 = 4u * sc.mm, ghTc = 4u * sc.gh + 1u;
    const unsigned big = a.big;

    int i = a.m, j = a.n, level = 0, status = 0, tiles = 0;
    long long nops = 0;
    uint32_t word = 0;
    bool first = true;
    auto emit = [&](int op) {          // lane 0 only
        word |= (uint32_t)op << (2 * (int)(nops & 15));
        if ((++nops & 15) == 0) { out.ops[(nops >> 4) - 1] = word; word = 0; }
    };

    while (i > 0 && j > 0 && status == 0) {
        const int band = (i - 1) / LBAND;
        const int cblk = (j - 1) / LCC;
        const int jc = cblk * LCC;                  // columns jc+1 .. j are re-filled
        const int width = j - jc;
        // ---- re-fill the tile with direction codes
        const int ib = band * LBAND + g * LR;
        LongLane<SIMPLE> ln;
        long_load_rows<SIMPLE>(sc, a, ib, ln);
        unsigned rh_prev;
        if (cblk == 0) {
#pragma unroll
            for (int r = 0; r < LR; ++r) {
                const unsigned c = (unsigned)a.colsum[min(ib + r + 1, a.m)];
                ln.hl[r] = min(big, c) << 2;
                ln.Lp[r] = min(big, c + sc.o) << 2;
            }
            rh_prev = ib == 0 ? 0u : (min(big, (unsigned)a.colsum[min(ib, a.m)]) << 2);
        } else {
            const int2 *src = a.vchk + (((size_t)band * a.nck + cblk) * 32 + g) * LR;
#pragma unroll
            for (int r = 0; r < LR; ++r) {
                const int2 v = __ldcg(src + r);
                ln.hl[r] = (unsigned)v.x;
                ln.Lp[r] = (unsigned)v.y;
            }
            // h(ib, jc): bottom-row h of the lane above at the checkpoint column
            unsigned above = __shfl_up_sync(FULL, ln.hl[LR - 1], 1);
            if (g == 0) {
                if (band > 0) above = (unsigned)__ldcg(a.brow + (size_t)(band - 1) * (a.n + 1) + jc).x;
                else above = min(big, (unsigned)a.rowsum[jc]) << 2;
            }
            rh_prev = above;
        }
        const int2 *up_brow = a.brow + (size_t)(band > 0 ? band - 1 : 0) * (a.n + 1);
        unsigned hb = 0, ub = 0, bv = 0;
        const int nsteps = width + 31;
        unsigned st0 = 0, st1 = 0;
        unsigned bblk = 0;
        int2 tblk = make_int2(0, 0);
        for (int t = 0; t < nsteps; ++t) {
            if ((t & 31) == 0) {
                // letters of seq_2 and the boundary row above the band, 32 columns at a time
                const int cj = min(jc + t + 1 + g, a.n);
                bblk = (unsigned)a.s2[cj - 1];
                if (band > 0) {
                    tblk = __ldcg(up_brow + cj);
                } else {
                    const unsigned rj = (unsigned)a.rowsum[cj];
                    tblk = make_int2((int)(min(big, rj) << 2), (int)(min(big, rj + sc.o) << 2));
                }
            }
            const int jl = t - g + 1;                // column within the tile, 1-based
            unsigned rh = __shfl_up_sync(FULL, hb, 1);
            unsigned ru = __shfl_up_sync(FULL, ub, 1);
            unsigned bin = __shfl_up_sync(FULL, bv, 1);
            const unsigned b0 = __shfl_sync(FULL, bblk, t & 31);
            const unsigned th = (unsigned)__shfl_sync(FULL, tblk.x, t & 31);
            const unsigned tu = (unsigned)__shfl_sync(FULL, tblk.y, t & 31);
            if (g == 0) { bin = b0; rh = th; ru = tu; }
            bv = bin;
            unsigned w0 = 0, w1 = 0;
            if (jl >= 1 && jl <= width) {
                unsigned diag = rh_prev;
                rh_prev = rh;
                const unsigned ghT = SIMPLE ? ghTc : (unsigned)s_gh[bv];
                long_cells<SIMPLE, true>(sc, s_sub, ln, diag, ru, bv, ghT, o4, mm4, hb, ub, w0, w1);
            }
            if (t & 1) {
                s_codes[(t >> 1) * 32 + g] = make_uint4(st0, st1, w0, w1);
            } else {
                st0 = w0;
                st1 = w1;
            }
        }
        if (nsteps & 1) s_codes[(nsteps >> 1) * 32 + g] = make_uint4(st0, st1, 0u, 0u);
        __syncwarp();
        ++tiles;
        // ---- walk inside the tile (lane 0)
        if (g == 0) {
            const uint32_t *cw = reinterpret_cast<const uint32_t *>(s_codes);
            auto plane = [&](int ii, int jj, int pl) -> unsigned {
                const int rem = (ii - 1) - band * LBAND;
                const int lg = rem / LR, r = rem - lg * LR;
                const int t = (jj - jc - 1) + lg;
                const int flat = (t & 1) * 2 + pl;
                const uint32_t w = cw[((size_t)(t >> 1) * 32 + lg) * 4 + flat];
                return (w >> (2 * (LR - 1 - r))) & 3u;
            };
            while (i > 0 && j > jc && (i - 1) / LBAND == band) {
                int op;
                if (level != 0 && (plane(i, j, 1) & (unsigned)level)) op = level;
                else op = (int)plane(i, j, 0);
                emit(op);
                if (op == 0) { --i; --j; } else if (op == 1) { --j; } else { --i; }
                level = op;
                if (first) {
                    first = false;
                    if ((i == 0) != (j == 0)) {
                        // boundary cell ranked once by the reference (m == 1 or n == 1)
                        long long v0 = big, v1 = big, v2 = big;
                        int want;
                        if (i == 0) { v1 = a.rowsum[j]; want = 1; } else { v2 = a.colsum[i]; want = 2; }
                        if (level == 1) { v0 += sc.o; v2 += sc.o; } else if (level == 2) { v0 += sc.o; v1 += sc.o; }
                        int k = 0;
                        long long best = v0;
                        if (v1 < best) { best = v1; k = 1; }
                        if (v2 < best) { k = 2; }
                        if (k != want) status = 1;
                    }
                }
            }
        }
        i = __shfl_sync(FULL, i, 0);
        j = __shfl_sync(FULL, j, 0);
        level = __shfl_sync(FULL, level, 0);
        status = __shfl_sync(FULL, status, 0);
        first = __shfl_sync(FULL, (int)first, 0) != 0;
        __syncwarp();
    }
    if (g == 0) {
        if (status == 0) {
            for (; j > 0; --j) emit(1);
            for (; i > 0; --i) emit(2);
        }
        if (nops & 15) out.ops[nops >> 4] = word;
        *out.n_ops = nops;
        *out.status = status;
        *out.tiles_refilled = tiles;
    }
}

}  // namespace ga
