// K2 + K5: one long pair on the whole chip (sm_100a).
//
// K2, long_fill_kernel -- fill without storing direction codes.  The DP matrix is cut into
// bands of LBAND = 32 lanes x LR rows; one warp sweeps one band as an anti-diagonal wavefront,
// bands pipelined across all SMs (band k reads the bottom row of band k-1 from global memory).
//
//   * ILP inside a lane.  The LR rows of a lane form LNC chains of LELL rows.  Chain q = g*LNC + k
//     (lane g, chain k) works at step t on column t - q + 1: the chains of a lane sit on different
//     columns, so their cell updates are independent and interleave in the instruction stream; only
//     the LELL rows of one chain depend on each other within a step.  (Round 1 ran all rows of a
//     lane on one column: one serial chain of 4 cells, a lone warp issuing every ~3.75 cycles.)
//     The chains of a step run bottom-up (chain LNC-1 first), so the shuffled inputs of chain 0
//     are needed last and the shuffle latency hides behind the other chains.
//   * Gap-shifted values.  X^(i,j) = 4 * (X(i,j) - GH(j) - GV(i)) with GH / GV the running sums of
//     the per-letter horizontal / vertical gap-extension costs.  In this domain
//         d0 = h^(i-1,j-1) + 4*(sub(a,b) - gap_h[b] - gap_v[a])          (one table load + add)
//         h  = min3(d0, L^(i,j-1), U^(i-1,j))
//         L^ = min(h + 4o, L^(i,j-1)),  U^ = min(h + 4o, U^(i-1,j))
//     -- the two gap-extension additions per cell are gone, and per-letter gap costs cost nothing.
//     (long_chain evaluates these in an equivalent order that keeps U off the row-to-row
//     dependency: 3 x VIADDMNMX + 1 x VIMNMX3 per cell.)
//   * Sequence tiles by TMA.  seq_2 reaches a band in tiles of LTCOLS letters copied by the TMA
//     engine (cp.async.bulk, mbarrier completion) into two shared-memory slots per warp, two tiles
//     ahead of the wavefront.
//   * Every spin loop is warp-uniform (votes): behind a per-lane spin loop the hardware may leave
//     the warp split for the rest of the band, and a split warp takes the collective slow path of
//     every shuffle.
//     L(i,j) = min(D0+o, D1, D2+o) is the level-1 candidate handed to column j+1, U(i,j) =
//     min(D0+o, D1+o, D2) the level-2 candidate handed to row i+1 (reference cell update,
//     globaligner.py:331-363).  The shift is the same for the three candidates of a cell, so every
//     comparison -- and every tie -- is unchanged.  Exact integers throughout.
//   * Boundary hand-off.  (h^, U^) of a band's last row go to global memory with one 64-bit store
//     per column; the consumer fetches them LSYNC columns at a time, half a block AHEAD of its
//     wavefront (the load latency hides behind half a block of steps), synchronised by the data itself:
//     entries are pre-set to L_INVALID and re-polled only if they have not arrived.  Bands are claimed in
//     increasing order from an atomic counter, so a band's predecessor is always running or
//     finished: no deadlock, no co-residency requirement.
// Kept for the traceback (shifted values; K5 adds the shift back):
//   * brow[k][j]      (h^, U^) of the last row of band k, every column  (8 B / column / band)
//   * vchk[k][c][...] (h^, L^) of every row of band k at column c*LCC   (8 B / row / checkpoint)
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

#ifndef GA_LONG_ELL
#define GA_LONG_ELL 6
#endif
#ifndef GA_LONG_NC
#define GA_LONG_NC 1
#endif
constexpr int LELL = GA_LONG_ELL;   // rows per chain (updated one after the other within a step)
constexpr int LNC = GA_LONG_NC;     // chains per lane (independent within a step)
constexpr int LR = LELL * LNC;      // rows per lane
constexpr int LBAND = 32 * LR;      // rows per band
constexpr int LCC = 128;            // columns between vertical checkpoints
#ifndef GA_LONG_LSYNC
#define GA_LONG_LSYNC 16
#endif
constexpr int LSYNC = GA_LONG_LSYNC; // columns per boundary-row fetch = steps per unrolled block
constexpr int L_INVALID = (int)0x80808080;   // boundary entry not yet written (memset 0x80)
static_assert(LNC <= 16 && LR <= 16 && (LSYNC == 8 || LSYNC == 16 || LSYNC == 32), "chains per lane / rows per lane / block length");

struct LongArgs {
    const uint8_t *s1, *s2;         // device letters
    int m, n;
    const int *colsum;              // m+1: o + sum_{t<=i} gap_v[s1[t-1]]  (column-0 level-2 values), [0] = o
    const int *rowsum;              // n+1: o + sum_{t<=j} gap_h[s2[t-1]]  (row-0 level-1 values)
    int2 *brow;                     // nbands x (n+1), shifted values
    int2 *vchk;                     // nbands x nck x 32 x LR, shifted values
    int *band_counter;
    int nbands, nck;
    unsigned big;                   // reference sentinel (max_cost+1)*max(m,n)
    int *out_cost;
    int o4, o4b;                    // both 4 * gap_open, as two parameters (see FillConsts)
};

// 4 * (GH(j) + GV(i)): what the fill subtracted from the values of cell (i, j)
__device__ __forceinline__ int long_shift(const LongArgs &a, int o, int i, int j) {
    return 4 * ((a.rowsum[j] - o) + (a.colsum[min(i, a.m)] - o));
}

template <bool SIMPLE>
struct LongLane {
    unsigned aw[LR], gvT[LR], hl[LR], Lp[LR];
};

__device__ __forceinline__ int2 ld_volatile_int2(const int2 *p) {
    int2 v;
    asm volatile("ld.relaxed.gpu.global.v2.s32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(p) : "memory");
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

// one column of LR cells for this lane (traceback kernels: plain x4 domain with level tags);
// returns (hb, ub) through references
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

constexpr bool LONG_PAIR_TABLE_OK = LELL % 2 == 0;     // MODE 2 needs the rows of a chain in pairs
// MODE 1: substitution costs (already shifted by the gap costs) from a table in shared memory --
// any scheme with A <= 64 letters.  MODE 0: compare-and-select for simple schemes over larger
// alphabets (the table would not fit).
//
// The column letters run ONE STEP AHEAD of the cells: at step t every chain already holds the
// letter of the column it works on (bv) and the substitution costs of its rows against it (sv),
// fetched during step t-1; during step t it fetches those of step t+1.  Letters simply shift
// along the chains (chain q's letter at step t+1 is chain q-1's at step t), so neither the
// shuffle that moves them between lanes nor the table load sits on the path of the cell update.
template <int MODE>
struct FillLane {
    unsigned aw[LR];        // MODE 1: shared-space byte address of the row letter's table row; MODE 0: the letter
    int hl[LR], Lp[LR];     // h^ and L^ of my rows at the last column their chain computed
    int sv[LR];             // shifted substitution cost of my rows at their chain's current column
    int ub[LNC];            // U^ out of the last row of chain k (input of the row below)
    int hd[LNC];            // h^ of the last row of chain k one column earlier (diagonal input of the row below)
    unsigned bv[LNC];       // letter of chain k's current column (MODE 1: its byte offset in a table row)
    int rh_prev;            // h^ of the row above my first row, one column earlier
    int rh_in, ru_in;       // h^ / U^ of the row above my first row at my chain 0's next column (shuffled a step ahead)
};

// (o4 and o4b hold the same number but arrive as two kernel parameters: computed in the kernel
// as 4 * o the assembler folds the multiply into an IMAD and emits h + 4o once plus two plain
// minima -- three instructions -- instead of two fused VIADDMNMX.)
struct FillConsts {
    int o4, o4b;            // 4 * gap_open
    int cm4, cx4;           // MODE 0: shifted match / mismatch cost
    unsigned k1;            // runtime 1 (IMAD forms, see SchemeDev::k_one)
};

// shifted substitution cost of row r against column letter b
template <int MODE>
__device__ __forceinline__ int long_sub(const FillLane<MODE> &ln, int r, unsigned b, const FillConsts &c) {
    if (MODE != 0) {
        int v;
        asm("ld.shared.s32 %0, [%1];" : "=r"(v) : "r"(ln.aw[r] * c.k1 + b));        // address add as IMAD (FMA pipe)
        return v;
    }
    return ln.aw[r] != b ? c.cx4 : c.cm4;
}

// MODE 2 (alphabets up to 8 letters, an even number of rows per chain): the table is indexed by the
// letters of TWO consecutive rows, an entry holds both costs -- one 8-byte load and one address per
// two rows.  (Shared-memory instructions are what four band-sweeping warps of an SM contend for.)
template <int MODE>
__host__ __device__ constexpr unsigned long_letter_scale() { return MODE == 1 ? 4u : MODE == 2 ? 8u : 1u; }

// substitution costs of rows [r0, r1) against the column letter with (scaled) offset b
template <int MODE>
__device__ __forceinline__ void long_sub_rows(FillLane<MODE> &ln, int r0, int r1, unsigned b, const FillConsts &c) {
    if (MODE == 2) {
#pragma unroll
        for (int r = r0; r < r1; r += 2) {
            int v0, v1;
            asm("ld.shared.v2.s32 {%0, %1}, [%2];" : "=r"(v0), "=r"(v1) : "r"(ln.aw[r] * c.k1 + b));
            ln.sv[r] = v0;
            ln.sv[r + 1] = v1;
        }
    } else {
#pragma unroll
        for (int r = r0; r < r1; ++r) ln.sv[r] = long_sub<MODE>(ln, r, b, c);
    }
}

// The LELL cells of chain K at its current column.  diag / up: h^(i-1, j-1) and U^(i-1, j) of the
// row above the chain's first row.
//
// CHECK (blocks at the edges of the matrix, where a chain's column may not exist yet or any more):
// the update of an inactive chain is made the IDENTITY by its inputs instead of being skipped --
// no branch, so a whole unrolled block stays one basic block, and no select behind the update on
// the dependency chain.  A row's state always satisfies h <= L <= h + 4o; feeding d0 = h (its own
// value), U-from-above = +inf and an infinite gap-open for the U it hands down gives
//     e = min(h, L) = h,  h' = min3(h, inf, inf) = h,  L' = min(h + 4o, L) = L,  U' = min(h + inf, inf) = inf
// so hl / Lp keep their values (the column-0 boundary before the chain starts, column N after it
// ends) and the chain below sees +inf, which it ignores for the same reason.  What an inactive
// chain leaves in hd / ub is never read: the chain below becomes active one step after it.
#ifndef GA_LONG_UFREE
#define GA_LONG_UFREE 1
#endif
constexpr int L_INF = 0x3f000000;       // above every value the range check of ga_align_long admits (|v| < 2^28)
constexpr int L_BIGADD = 0x4f000000;    // v + L_BIGADD >= L_INF for every admitted v, without overflow
template <int MODE, int K, int CHECK>
__device__ __forceinline__ void long_chain(FillLane<MODE> &ln, int diag, int up, bool act, const FillConsts &c) {
    // The textbook order  h_r = min(d0_r, L_r, U_{r-1});  U_r = min(h_r + 4o, U_{r-1})  threads U through
    // all rows of the chain: 8 cycles of dependent latency per row, then a 24-cycle shuffle to the lane
    // below -- which needs it for ITS first row of the next step.  With e_r = min(d0_r, L_r) (no U in
    // it) and gap_open >= 0:
    //     U_r = min(e_r + 4o, U_{r-1} + 4o, U_{r-1}) = min(e_r + 4o, U_{r-1})
    //         = min(P_r, U_in),      P_r = min_{k <= r} e_k + 4o      (a chain that never sees U_in)
    //     h_r = min(e_r, U_{r-1}) = min3(e_r, P_{r-1}, U_in)
    // so every row's h is ONE instruction behind U_in, the U handed down is min(P_last, U_in), and
    // what is left on the path from lane to lane is that minimum plus the shuffle (28 cycles per
    // step instead of 72).  One more ALU instruction per cell (4 instead of 3); exact -- plain
    // integer minima, no tags in this kernel.
    const int ob = (!CHECK || act) ? c.o4b : L_BIGADD;   // inactive: e + ob >= inf, so the U handed down stays +inf
    if (CHECK && !act) up = L_INF;
#if GA_LONG_UFREE
    int P = L_INF;
#pragma unroll
    for (int r = K * LELL; r < (K + 1) * LELL; ++r) {
        // e = min(diag + sub', L): one fused add-min (inactive: min(own h + 0, L) = own h)
        const int dg = (CHECK && !act) ? ln.hl[r] : diag;
        const int sv = (CHECK && !act) ? 0 : ln.sv[r];
        const int e = __viaddmin_s32(dg, sv, ln.Lp[r]);
        const int h = __vimin3_s32(e, P, up);
        P = __viaddmin_s32(e, ob, P);
        diag = ln.hl[r];
        ln.hl[r] = h;
        ln.Lp[r] = __viaddmin_s32(h, c.o4, ln.Lp[r]);
    }
    ln.hd[K] = diag;
    ln.ub[K] = min(P, up);
#else
#pragma unroll
    for (int r = K * LELL; r < (K + 1) * LELL; ++r) {
        int d0 = (int)((unsigned)diag * c.k1 + (unsigned)ln.sv[r]);                 // IMAD
        if (CHECK && !act) d0 = ln.hl[r];
        const int h = __vimin3_s32(d0, ln.Lp[r], up);
        diag = ln.hl[r];
        ln.hl[r] = h;
        ln.Lp[r] = __viaddmin_s32(h, c.o4, ln.Lp[r]);
        up = __viaddmin_s32(h, ob, up);
    }
    ln.hd[K] = diag;
    ln.ub[K] = up;
#endif
}

// a value the compiler must keep in a register (it cannot re-derive it from where it came from)
__device__ __forceinline__ int pin_reg(int v) {
    int r;
    asm volatile("mov.b32 %0, %1;" : "=r"(r) : "r"(v));
    return r;
}
template <typename T>
__device__ __forceinline__ T *pin_reg(T *v) {
    T *r;
    asm volatile("mov.b64 %0, %1;" : "=l"(r) : "l"(v));
    return r;
}

__device__ __forceinline__ void st_cg_pred4(bool p, int2 *dst, int x, int y, int z, int w) {
    asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.s32 q, %0, 0;\n\t@q st.global.cg.v4.s32 [%1], {%2, %3, %4, %5};\n\t}"
                 :: "r"((int)p), "l"(dst), "r"(x), "r"(y), "r"(z), "r"(w) : "memory");
}
__device__ __forceinline__ void st_cg_pred(bool p, int2 *dst, int x, int y) {
    asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.s32 q, %0, 0;\n\t@q st.global.cg.v2.s32 [%1], {%2, %3};\n\t}"
                 :: "r"((int)p), "l"(dst), "r"(x), "r"(y) : "memory");
}

// chain K (>= 1) of one step: my column is j (chain K lags chain 0 by K columns); it reads the state
// its upper neighbour chain had one step ago
template <int MODE, int CHECK, int K>
__device__ __forceinline__ void long_one_chain(FillLane<MODE> &ln, int j, int N, bool last_lane, int2 *brow_j,
                                               const FillConsts &c) {
    const bool act = CHECK == 0 || (j >= 1 && (CHECK == 1 || j <= N));
    long_chain<MODE, K, CHECK>(ln, ln.hd[K - 1], ln.ub[K - 1], act, c);
    if (K == LNC - 1) st_cg_pred(last_lane && act, brow_j, ln.hl[LR - 1], ln.ub[LNC - 1]);
}

// chains K, K-1, ..., 1 of one step
template <int MODE, int CHECK, int K>
__device__ __forceinline__ void long_mid_chains(FillLane<MODE> &ln, int j0, int N, bool last_lane, int2 *brow_j,
                                                const FillConsts &c) {
    if constexpr (K >= 1) {
        long_one_chain<MODE, CHECK, K>(ln, j0 - K, N, last_lane, brow_j, c);
        long_mid_chains<MODE, CHECK, K - 1>(ln, j0, N, last_lane, brow_j, c);
    }
}

// One block of LSYNC steps of a band sweep, fully unrolled and BRANCH-FREE (compile-time store
// offsets, predicated stores, selects).  CHECK = 0 is the steady state (every chain inside the
// matrix for the whole block), 1 the ramp-in of a band (columns >= 1 to be tested only), 2 the general edge block.  `slot`: shared-space address of this block's LSYNC staged inputs of
// lane 0: entry u = (h^, U^ of the row above the band at column t0 + 1 + u, letter of the NEXT
// column), 16 bytes; every lane loads entry u + 1 during step u (a broadcast), lane 0 keeps it.
//
// A lone warp issues in order, so every value that crosses lanes is requested long before it is
// used: the bottom row of my last chain is shuffled to the lane below right after that chain's
// update and consumed by the lane below's chain 0 in its NEXT step (ln.rh_in / ln.ru_in carry it
// across); the letter shuffle is consumed only by the loads of the next step's substitution costs.
// Without a branch in the block no shuffle needs a convergence check, and the assembler overlaps
// the chains of a step and consecutive steps.
//
// Column checkpoints (CK blocks only): at step t the only chain of the band that can sit on a
// multiple of LCC is chain q* = (t + 1) mod LCC (if the band has that many chains).  q* is warp-
// uniform and, with LNC a power of two, its chain index within the lane is known at compile time:
// LELL predicated stores in the owning lane.  Blocks in which no step can hit are instantiated
// without that code.
// ---- TMA staging of the sequence tile (cp.async.bulk + mbarrier; SASS: UBLKCP / SYNCS)
// The letters of seq_2 a band sweeps over arrive in TILES of LTCOLS columns: tile S is the 16-byte
// aligned piece seq_2[S * LTCOLS .. + LTILE) -- the letters of columns S*LTCOLS + 1 .. (S+1)*LTCOLS + 16,
// what the blocks of steps t0 in [S*LTCOLS, (S+1)*LTCOLS) look up.  The TMA engine copies a tile
// into one of the warp's two shared-memory slots and signals the slot's mbarrier; no thread holds
// the data in flight.  One elected lane requests tile S+2 when the warp moves from tile S to S+1
// (a whole tile of steps ahead); the wait at that point finds tile S+1 long complete.
// (Waiting once per BLOCK on a 32-byte tile cost 10 % of the fill: a lone in-order warp pays the
// mbarrier round trip and the dependent shared-memory load every 16 steps.)
constexpr int LTCOLS = 512;             // columns per tile
constexpr int LTILE = LTCOLS + 16;      // bytes per tile: one more 16-byte piece for the look-ahead letters of the last block
static_assert(LSYNC <= 16 && LTCOLS % LSYNC == 0, "the look-ahead of a tile's last block stays inside the extra 16 bytes");
__device__ __forceinline__ void tma_bar_init(unsigned bar) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(bar) : "memory");
}
__device__ __forceinline__ void tma_bar_init_done() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void tma_load_tile(unsigned dst, const void *src, unsigned bar) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"(LTILE) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(dst), "l"(src), "r"(LTILE), "r"(bar) : "memory");
}
__device__ __forceinline__ unsigned ld_shared_u8(unsigned addr) {
    unsigned v;
    asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
    return v;
}
// (warp-uniform spin, like every loop of the fill kernel whose trip count the compiler cannot see:
// behind a per-lane spin loop the hardware may let part of the warp run ahead for good, and a split
// warp takes the slow collective path of every shuffle)
__device__ __forceinline__ void tma_bar_wait(unsigned bar, unsigned parity) {
    unsigned done;
    do {
        asm volatile("{\n\t.reg .pred p;\n\t"
                     "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
                     "selp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    } while (!__all_sync(FULL, done != 0u));
}

// (Measured and rejected, round 2: handing the last row of a lane to the lane below through a
// shared-memory exchange area (one STS.64 + one LDS.64 per step, volatile so that the assembler keeps
// their order) with every lane reading its own column letter from a TMA-fed ring -- no shuffle and no
// lane-0 select left, 42 instead of 52 instructions per step, 25 instead of 28 on the ALU pipe -- ran
// 8.53 ms against 8.16 ms for this form at 100 kb x 100 kb: the four band-sweeping warps of an SM
// queue on the shared-memory pipe (10 of its instructions per step each), which the shuffles do not.)
// Raw loads for the LSYNC columns of the NEXT block (lanes >= LSYNC duplicate a valid column): the
// sequence tile (TMA) and the boundary-row entry above this band (which may still be
// L_INVALID; band 0: the row-0 running sum instead).  Nothing here USES a loaded value -- a lone
// warp would stall on it -- that happens at the start of the next block.  Issued from the MIDDLE of
// a block: early enough to hide the L2 latency behind half a block of steps, late enough that the
// band above, which runs only a few dozen steps ahead, has already written the entries.
struct FetchCtx {
    const uint8_t *s2;
    const int2 *up_brow;
    const int *rowsum;
    int band, cj_next;              // my column of the next block
    int2 traw;
    unsigned tile0, bar0;           // shared-space addresses: the warp's two sequence-tile slots and their mbarriers
    unsigned seq0;                  // number (per warp, over all its bands) of this band's tile 0: tile k uses slot k & 1, phase (k >> 1) & 1
    unsigned laddr;                 // shared-space address of the letter my lane looks up for the next block
    unsigned braw;
    __device__ __forceinline__ unsigned tile(unsigned k) const { return tile0 + (k & 1u) * LTILE; }
    __device__ __forceinline__ unsigned bar(unsigned k) const { return bar0 + (k & 1u) * 8u; }
    __device__ __forceinline__ void request(int S) const {        // one lane only
        tma_load_tile(tile(seq0 + S), s2 + (size_t)S * LTCOLS, bar(seq0 + S));
    }
    __device__ __forceinline__ void wait(int S) const { tma_bar_wait(bar(seq0 + S), ((seq0 + S) >> 1) & 1u); }
};
__device__ __forceinline__ void long_fetch(FetchCtx &f, int N) {
    // the letter of the column after mine (0-based index cj_next of seq_2) out of its tile: the
    // block loop has waited for that tile.  Past the end of the sequence this reads slack bytes,
    // discarded at the use.
    f.braw = ld_shared_u8(f.laddr);
    f.traw = make_int2(0, 0);
    if (f.cj_next <= N) {
        if (f.band > 0) f.traw = ld_volatile_int2(f.up_brow + f.cj_next);
        else f.traw.x = f.rowsum[f.cj_next];
    }
}

template <int MODE, int CHECK, bool CK, int U>
__device__ __forceinline__ void long_fill_step(FillLane<MODE> &ln, int4 &e, unsigned slot, int g, int q0,
                                               bool last_lane, int t0, int N, int2 *brow_j0, int2 *band_vchk,
                                               const FillConsts &c, FetchCtx &f) {
    if constexpr (U == LSYNC / 2) long_fetch(f, N);
    const unsigned nbs = __shfl_up_sync(FULL, ln.bv[LNC - 1], 1);      // my chain 0's letter of the NEXT step
    const int j0 = t0 + U - q0 + 1;                    // column of my chain 0 at this step
    int rh_next = 0, ru_next = 0;
    if constexpr (LNC > 1) {
        long_one_chain<MODE, CHECK, LNC - 1>(ln, j0 - (LNC - 1), N, last_lane, brow_j0 + U, c);
        rh_next = __shfl_up_sync(FULL, ln.hl[LR - 1], 1);
        ru_next = __shfl_up_sync(FULL, ln.ub[LNC - 1], 1);
        long_mid_chains<MODE, CHECK, LNC - 2>(ln, j0, N, last_lane, brow_j0 + U, c);
    }
    {   // chain 0: the row above is the last row of the lane above (lane 0: of the band above)
        const int rh = g == 0 ? e.x : ln.rh_in;
        const int ru = g == 0 ? e.y : ln.ru_in;
        const bool act = CHECK == 0 || (j0 >= 1 && (CHECK == 1 || j0 <= N));
        // (rh_prev follows the lane above unconditionally: while that lane's last chain is inactive
        // its hl[LR-1] is h^(ib, 0), exactly what my chain 0 needs as its first diagonal input)
        const int diag = ln.rh_prev;
        ln.rh_prev = rh;
        long_chain<MODE, 0, CHECK>(ln, diag, ru, act, c);
        if (LNC == 1) st_cg_pred(last_lane && act, brow_j0 + U, ln.hl[LR - 1], ln.ub[0]);
    }
    if constexpr (LNC == 1) {
        rh_next = __shfl_up_sync(FULL, ln.hl[LR - 1], 1);
        ru_next = __shfl_up_sync(FULL, ln.ub[0], 1);
    }
    ln.rh_in = rh_next;
    ln.ru_in = ru_next;
    // letters and substitution costs of the next step (loads issued now, used a step later)
    const unsigned nb0 = g == 0 ? (unsigned)e.z : nbs;
#pragma unroll
    for (int k = LNC - 1; k >= 0; --k) ln.bv[k] = k == 0 ? nb0 : ln.bv[k - 1];
#pragma unroll
    for (int k = 0; k < LNC; ++k) long_sub_rows<MODE>(ln, k * LELL, (k + 1) * LELL, ln.bv[k], c);
    if constexpr (U + 1 < LSYNC)
        asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(e.x), "=r"(e.y), "=r"(e.z), "=r"(e.w)
                     : "r"(slot + 16u * (U + 1)));
    if constexpr (CK || CHECK == 2) {
        constexpr int KS = (U + 1) & (LNC - 1);        // q* mod LNC (t0 and LCC are multiples of LNC)
        const int tt = t0 + U + 1;
        const int qs = tt & (LCC - 1);                 // q*, warp-uniform
        const int j = tt - qs;                         // its column, a multiple of LCC
        bool hit = qs < 32 * LNC && (qs / LNC) == g;
        if (CHECK == 2) hit = hit && j >= 1 && j <= N;
        int2 *dst = band_vchk + ((size_t)(tt / LCC) * 32 + g) * LR;
        if constexpr (LELL % 2 == 0) {      // two rows per 16-byte store (a lane's LR entries are 16-byte aligned)
#pragma unroll
            for (int r = KS * LELL; r < (KS + 1) * LELL; r += 2)
                st_cg_pred4(hit, dst + r, ln.hl[r], ln.Lp[r], ln.hl[r + 1], ln.Lp[r + 1]);
        } else {
#pragma unroll
            for (int r = KS * LELL; r < (KS + 1) * LELL; ++r) st_cg_pred(hit, dst + r, ln.hl[r], ln.Lp[r]);
        }
    }
    if constexpr (U + 1 < LSYNC)
        long_fill_step<MODE, CHECK, CK, U + 1>(ln, e, slot, g, q0, last_lane, t0, N, brow_j0, band_vchk, c, f);
}

template <int MODE, int CHECK, bool CK>
__device__ __forceinline__ void long_fill_block(FillLane<MODE> &ln, unsigned slot, int g, int t0,
                                                int N, int2 *my_brow, int2 *band_vchk, const FillConsts &c, FetchCtx &f) {
    static_assert(32 * LNC <= LCC && (LNC & (LNC - 1)) == 0 && LSYNC % LNC == 0,
                  "at most one chain of a band on a checkpoint column per step; its index in the lane is static");
    const int q0 = g * LNC;
    int2 *brow_j0 = my_brow + (t0 - (q0 + LNC - 1) + 1);   // column of my last chain at the block's first step
    int4 e;
    asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(e.x), "=r"(e.y), "=r"(e.z), "=r"(e.w) : "r"(slot));
    long_fill_step<MODE, CHECK, CK, 0>(ln, e, slot, g, q0, g == 31, t0, N, brow_j0, band_vchk, c, f);
}

template <int MODE>
__global__ void __launch_bounds__(32 * LWARPS)
long_fill_kernel(const SchemeDev sc, const LongArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ int4 s_stage[LWARPS][2][LSYNC];          // lane 0's inputs, two blocks per warp
    __shared__ __align__(16) uint8_t s_tile[LWARPS][2][LTILE];   // TMA-staged tiles of seq_2, two per warp
    __shared__ __align__(8) unsigned long long s_bar[LWARPS][2];  // their mbarriers
    if (threadIdx.x == 0) {
#pragma unroll
        for (int k = 0; k < 2 * LWARPS; ++k) tma_bar_init((unsigned)__cvta_generic_to_shared(&s_bar[0][0] + k));
        tma_bar_init_done();
    }
    int *s_tab = reinterpret_cast<int *>(smem_raw);
    if (MODE == 1) {
        // shifted substitution costs 4*(sub(a,b) - gap_h[b] - gap_v[a]); row A = padding letter
        // (rows past the end of seq_1: sub4 row A and gv_raw[A] are zero)
        const int n_sub = (sc.A + 1) * sc.A;
        for (int k = threadIdx.x; k < n_sub; k += blockDim.x) {
            const int ai = k / sc.A, b = k - ai * sc.A;
            s_tab[k] = sc.sub4[k] - 4 * sc.gh_raw[b] - 4 * sc.gv_raw[ai];
        }
    } else if (MODE == 2) {
        // the same costs for every PAIR of row letters (a0, a1): entry b of pair row a0 * (A+1) + a1
        const int A1 = sc.A + 1, n_ent = A1 * A1 * sc.A;
        for (int k = threadIdx.x; k < n_ent; k += blockDim.x) {
            const int b = k % sc.A, pr = k / sc.A, a0 = pr / A1, a1 = pr - a0 * A1;
            s_tab[2 * k] = sc.sub4[a0 * sc.A + b] - 4 * sc.gh_raw[b] - 4 * sc.gv_raw[a0];
            s_tab[2 * k + 1] = sc.sub4[a1 * sc.A + b] - 4 * sc.gh_raw[b] - 4 * sc.gv_raw[a1];
        }
    }
    __syncthreads();
    const unsigned tab_base = (unsigned)__cvta_generic_to_shared(s_tab);
    const int g = threadIdx.x & 31;     // the warps of a block never synchronise with each other again
    int4 *my_stage = &s_stage[threadIdx.x >> 5][0][0];
    const unsigned stage_base = (unsigned)__cvta_generic_to_shared(my_stage);
    FillConsts c;
    c.o4 = a.o4;
    c.o4b = a.o4b;
    c.cm4 = -4 * (sc.gh + sc.gv);
    c.cx4 = 4 * (sc.mm - sc.gh - sc.gv);
    c.k1 = sc.k_one;
    const unsigned big = a.big;
    // hot fields in registers (kernel parameters live in constant memory; an LDC inside
    // the step loop sits on the critical path of a lone warp)
    // (pinned: a plain copy is rematerialised as a constant-bank load wherever it is used -- 3 % of
    // the kernel's time sat on those loads in the block prologue)
    const int N = pin_reg(a.n), M = a.m;
    const uint8_t *const S2 = a.s2;
    const int *const ROWSUM = a.rowsum;
    const int O = sc.o;
    const int wi = threadIdx.x >> 5;
    unsigned tile_seq = 0;              // sequence tiles this warp has consumed (slot = seq & 1, barrier phase = (seq >> 1) & 1)
    for (;;) {
        int band = 0;
        if (g == 0) band = atomicAdd(a.band_counter, 1);
        band = __shfl_sync(FULL, band, 0);
        if (band >= a.nbands) break;
        const int ib = band * LBAND + g * LR;
        FillLane<MODE> ln;
#pragma unroll
        for (int k = 0; k < LNC; ++k) { ln.ub[k] = 0; ln.hd[k] = 0; ln.bv[k] = 0u; }
        if (g == 0) ln.bv[0] = long_letter_scale<MODE>() * (unsigned)S2[0];     // chain 0 of lane 0 starts on column 1
#pragma unroll
        for (int r = 0; r < LR; ++r) {
            const int row = ib + r;
            const unsigned la = row < M ? (unsigned)a.s1[row] : (unsigned)sc.A;
            if (MODE == 2) {        // even rows: the pair row of (my letter, the next row's letter)
                const unsigned lb = row + 1 < M ? (unsigned)a.s1[min(row + 1, M - 1)] : (unsigned)sc.A;
                ln.aw[r] = tab_base + 8u * (unsigned)sc.A * (la * (unsigned)(sc.A + 1) + lb);
            } else if (MODE == 1) ln.aw[r] = tab_base + 4u * (unsigned)sc.A * la;
            else ln.aw[r] = row < M ? la : (unsigned)PAD_A;
            const int cs = a.colsum[min(ib + r + 1, M)];
            const int gvs = cs - O;
            ln.hl[r] = 4 * ((int)min(big, (unsigned)cs) - gvs);
            ln.Lp[r] = 4 * ((int)min(big, (unsigned)(cs + O)) - gvs);
        }
#pragma unroll
        for (int k = 0; k < LNC; ++k) long_sub_rows<MODE>(ln, k * LELL, (k + 1) * LELL, ln.bv[k], c);
        ln.rh_prev = 0;
        if (ib != 0) {
            const int cs = a.colsum[min(ib, M)];
            ln.rh_prev = 4 * ((int)min(big, (unsigned)cs) - (cs - O));
        }
        int2 *my_brow = pin_reg(a.brow + (size_t)band * (N + 1));
        const int2 *up_brow = pin_reg(a.brow + (size_t)(band > 0 ? band - 1 : 0) * (N + 1));
        int2 *band_vchk = pin_reg(a.vchk + ((size_t)band * a.nck) * (32 * LR));

        FetchCtx f;
        f.s2 = S2;
        f.up_brow = up_brow;
        f.rowsum = ROWSUM;
        f.band = band;
        f.cj_next = 1 + (g & (LSYNC - 1));
        f.tile0 = (unsigned)__cvta_generic_to_shared(&s_tile[wi][0][0]);
        f.bar0 = (unsigned)__cvta_generic_to_shared(&s_bar[wi][0]);
        f.seq0 = tile_seq;
        if (g == 0) { f.request(0); f.request(1); }
        f.wait(0);
        f.laddr = f.tile(f.seq0) + 1u + (unsigned)(g & (LSYNC - 1));
        long_fetch(f, N);                   // block 0
        const int nsteps = N + 32 * LNC - 1;
        // what the shuffle at the end of a (virtual) step -1 would have delivered: the lane above is
        // still on column 0 (rh_prev follows rh_in unconditionally, see long_fill_step)
        ln.rh_in = ln.rh_prev;
        ln.ru_in = 0;
        __syncwarp();                       // the previous band's last reads of the staging area are done
        for (int t0 = 0, blk = 0; t0 < nsteps; t0 += LSYNC, ++blk) {
            int2 tblk = f.traw;
            const int cj = f.cj_next;       // = t0 + 1 + (g & (LSYNC - 1))
            const unsigned bnext = cj + 1 <= N ? f.braw : 0u;      // letter of the column after mine
            if (band > 0) {
                // entries requested half a block ago: poll only what has not arrived.  The loop is
                // warp-uniform (vote), the reload predicated: no lane leaves the others behind, the
                // warp enters the unrolled block converged (a split warp runs every shuffle of the
                // block through the slow collective path -- 7x slower).
                bool miss = cj <= N && tblk.x == L_INVALID;
                while (__any_sync(FULL, miss)) {
                    if (miss) {
                        tblk = ld_volatile_int2(up_brow + cj);
                        miss = tblk.x == L_INVALID;
                    }
                }
            } else if (cj <= N) {
                const int rj = tblk.x;
                const int ghs = rj - O;
                tblk = make_int2(4 * ((int)min(big, (unsigned)rj) - ghs), 4 * ((int)min(big, (unsigned)(rj + O)) - ghs));
            }
            if (g < LSYNC) my_stage[(blk & 1) * LSYNC + g] = make_int4(tblk.x, tblk.y, (int)(long_letter_scale<MODE>() * bnext), 0);
            __syncwarp();
            f.cj_next = cj + LSYNC;         // (past the last block the loads are guarded by cj <= N)
            f.laddr += LSYNC;
            if (((t0 + LSYNC) & (LTCOLS - 1)) == 0) {
                // this block's fetch crosses into the next tile: it must have landed; the tile behind
                // us (whose letters all sit in registers or the staging area by now) gives its slot
                // to the tile after the next
                const int S = t0 / LTCOLS;
                f.wait(S + 1);
                if (g == 0) f.request(S + 2);
                f.laddr = f.tile(f.seq0 + S + 1) + 1u + (unsigned)(g & (LSYNC - 1));
            }
            const unsigned slot = stage_base + (unsigned)((blk & 1) * LSYNC * 16);
            const int tq = t0 & (LCC - 1);
            const bool ck = tq < 32 * LNC || tq + LSYNC >= LCC;        // can a step of this block hit a checkpoint column?
            if (t0 >= 32 * LNC && t0 + LSYNC <= N) {
                if (ck) long_fill_block<MODE, 0, true>(ln, slot, g, t0, N, my_brow, band_vchk, c, f);
                else long_fill_block<MODE, 0, false>(ln, slot, g, t0, N, my_brow, band_vchk, c, f);
            } else if (t0 + LSYNC <= N && t0 + LSYNC < LCC) {
                // ramp-in: chains may not have started yet, none can be past column N or on a
                // checkpoint column.  These blocks set the lag between consecutive bands.
                long_fill_block<MODE, 1, false>(ln, slot, g, t0, N, my_brow, band_vchk, c, f);
            } else {
                long_fill_block<MODE, 2, true>(ln, slot, g, t0, N, my_brow, band_vchk, c, f);
            }
        }
        {   // the last tile requested is never used; retire it so that the barrier phases stay in step
            const int S_last = ((nsteps + LSYNC - 1) / LSYNC * LSYNC) / LTCOLS + 1;    // tiles 0 .. S_last were requested
            f.wait(S_last);
            tile_seq = f.seq0 + (unsigned)S_last + 1u;
        }
        // the last active step of every chain was column N: hl[] now holds h^(row, N)
        if (band == (M - 1) / LBAND && g == ((M - 1) % LBAND) / LR) {
            const int rm = (M - 1) % LR;
            int v = 0;
#pragma unroll
            for (int r = 0; r < LR; ++r)
                if (r == rm) v = ln.hl[r];
            *a.out_cost = (v >> 2) + (ROWSUM[N] - O) + (a.colsum[M] - O);
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

// ---- shared pieces of the traceback kernels -------------------------------------------------

struct TraceConsts {
    unsigned o4, mm4, ghTc, big;
};

// Left-boundary state of the lanes' rows at column jc (a multiple of LCC): column-0 formulas
// for jc == 0, the fill's checkpoint otherwise.  Returns h(ib_lane, jc), the diagonal input of
// the lane's first row at column jc+1.
template <bool SIMPLE>
__device__ __forceinline__ unsigned long_restore(const SchemeDev &sc, const LongArgs &a, int band, int jc, int g,
                                                 LongLane<SIMPLE> &ln, unsigned big) {
    const int ib = band * LBAND + g * LR;
    long_load_rows<SIMPLE>(sc, a, ib, ln);
    if (jc == 0) {
#pragma unroll
        for (int r = 0; r < LR; ++r) {
            const unsigned c = (unsigned)a.colsum[min(ib + r + 1, a.m)];
            ln.hl[r] = min(big, c) << 2;
            ln.Lp[r] = min(big, c + sc.o) << 2;
        }
        return ib == 0 ? 0u : (min(big, (unsigned)a.colsum[min(ib, a.m)]) << 2);
    }
    // (the fill keeps gap-shifted values, see long_shift)
    const int2 *src = a.vchk + (((size_t)band * a.nck + jc / LCC) * 32 + g) * LR;
#pragma unroll
    for (int r = 0; r < LR; ++r) {
        const int2 v = __ldcg(src + r);
        const int sh = long_shift(a, sc.o, ib + r + 1, jc);
        ln.hl[r] = (unsigned)(v.x + sh);
        ln.Lp[r] = (unsigned)(v.y + sh);
    }
    unsigned above = __shfl_up_sync(FULL, ln.hl[LR - 1], 1);   // bottom-row h of the lane above
    if (g == 0) {
        if (band > 0) above = (unsigned)(__ldcg(a.brow + (size_t)(band - 1) * (a.n + 1) + jc).x + long_shift(a, sc.o, band * LBAND, jc));
        else above = min(big, (unsigned)a.rowsum[jc]) << 2;
    }
    return above;
}

// Re-fill columns jc+1 .. jc+width of `band` with direction codes into shared memory
// (two steps per uint4: plane 0 / plane 1 of the even step, then of the odd step).
template <bool SIMPLE>
__device__ __forceinline__ void long_refill_codes(const SchemeDev &sc, const LongArgs &a, const int *s_sub,
                                                  const int *s_gh, uint4 *s_codes, int band, int jc, int width,
                                                  int g, const TraceConsts &k) {
    LongLane<SIMPLE> ln;
    unsigned rh_prev = long_restore<SIMPLE>(sc, a, band, jc, g, ln, k.big);
    const int2 *up_brow = a.brow + (size_t)(band > 0 ? band - 1 : 0) * (a.n + 1);
    unsigned hb = 0, ub = 0, bv = 0, st0 = 0, st1 = 0, bblk = 0;
    int2 tblk = make_int2(0, 0);
    const int nsteps = width + 31;
    for (int t = 0; t < nsteps; ++t) {
        if ((t & 31) == 0) {       // letters of seq_2 and the boundary row above, 32 columns at a time
            const int cj = min(jc + t + 1 + g, a.n);
            bblk = (unsigned)a.s2[cj - 1];
            if (band > 0) {
                tblk = __ldcg(up_brow + cj);
                const int sh = long_shift(a, sc.o, band * LBAND, cj);
                tblk.x += sh;
                tblk.y += sh;
            } else {
                const unsigned rj = (unsigned)a.rowsum[cj];
                tblk = make_int2((int)(min(k.big, rj) << 2), (int)(min(k.big, rj + sc.o) << 2));
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
            const unsigned ghT = SIMPLE ? k.ghTc : (unsigned)s_gh[bv];
            long_cells<SIMPLE, true>(sc, s_sub, ln, diag, ru, bv, ghT, k.o4, k.mm4, hb, ub, w0, w1);
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
}

// direction at cell (ii, jj) of the tile in shared memory, for a successor on `level`
__device__ __forceinline__ int long_tile_dir(const uint4 *s_codes, int band, int jc, int ii, int jj, int level) {
    const uint32_t *cw = reinterpret_cast<const uint32_t *>(s_codes);
    const int rem = (ii - 1) - band * LBAND;
    const int lg = rem / LR, r = rem - lg * LR;
    const int t = (jj - jc - 1) + lg;
    const uint32_t *base = cw + ((size_t)(t >> 1) * 32 + lg) * 4 + (t & 1) * 2;
    const int sh = 2 * (LR - 1 - r);
    if (level != 0 && (((base[1] >> sh) & 3u) & (unsigned)level)) return level;
    return (int)((base[0] >> sh) & 3u);
}

template <bool SIMPLE>
__device__ __forceinline__ void long_load_tables(const SchemeDev &sc, int *s_sub, int *&s_gh) {
    s_gh = s_sub;
    if (!SIMPLE) {
        const int n_sub = (sc.A + 1) * sc.A;
        s_gh = s_sub + n_sub;
        for (int k = threadIdx.x; k < n_sub; k += blockDim.x) s_sub[k] = sc.sub4[k];
        for (int k = threadIdx.x; k < sc.A; k += blockDim.x) s_gh[k] = sc.gh4t[k];
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------- K5, sequential form
// One warp re-fills and walks tile after tile.  Exact for every input; used for tiny pairs,
// with GA_F_SERIAL, and as the fallback of the parallel form below.
template <bool SIMPLE>
__global__ void __launch_bounds__(32)
long_trace_kernel(const SchemeDev sc, const LongArgs a, const TraceOut out, const int *skip_unless_failed) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint4 *s_codes = reinterpret_cast<uint4 *>(smem_raw);
    int *s_sub = reinterpret_cast<int *>(smem_raw + (size_t)LT_NTB * 32 * sizeof(uint4));
    int *s_gh;
    if (skip_unless_failed && *skip_unless_failed == 0) return;      // the parallel form succeeded
    long_load_tables<SIMPLE>(sc, s_sub, s_gh);
    const int g = threadIdx.x;
    const TraceConsts k{4u * sc.o, 4u * sc.mm, 4u * sc.gh + 1u, a.big};

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
        const int jc = ((j - 1) / LCC) * LCC;       // columns jc+1 .. j are re-filled
        long_refill_codes<SIMPLE>(sc, a, s_sub, s_gh, s_codes, band, jc, j - jc, g, k);
        ++tiles;
        if (g == 0) {
            while (i > 0 && j > jc && (i - 1) / LBAND == band) {
                const int op = long_tile_dir(s_codes, band, jc, i, j, level);
                emit(op);
                if (op == 0) { --i; --j; } else if (op == 1) { --j; } else { --i; }
                level = op;
                if (first) {
                    first = false;
                    if ((i == 0) != (j == 0)) {
                        // boundary cell ranked once by the reference (m == 1 or n == 1)
                        long long v0 = k.big, v1 = k.big, v2 = k.big;
                        int want;
                        if (i == 0) { v1 = a.rowsum[j]; want = 1; } else { v2 = a.colsum[i]; want = 2; }
                        if (level == 1) { v0 += sc.o; v2 += sc.o; } else if (level == 2) { v0 += sc.o; v1 += sc.o; }
                        int pick = 0;
                        long long best = v0;
                        if (v1 < best) { best = v1; pick = 1; }
                        if (v2 < best) { pick = 2; }
                        if (pick != want) status = 1;
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

// ---------------------------------------------------------------------------- K5, parallel form
// The sequential form spends its time on a chain of ~m/LBAND + n/LCC dependent tile refills.
// The parallel form removes the dependency between bands:
//   (a) long_xprop_kernel, all bands at once: re-fill each band over a corridor of columns
//       around the diagonal and propagate, with every cell and level L, the record
//       X_L(i,j) = where the canonical walk from state (i, j, L) leaves the band:
//       TOP(j', L') = crossing into the band above at column j' with successor level L',
//       LEFT(i')    = reaching column 0 at row i', OUT = running out of the corridor.
//       X follows the same decisions as the codes: X_0 = X of the c0-predecessor,
//       X_1 = bL ? X_1(left) : X_0, X_2 = bU ? X_2(up) : X_0.  Kept: X of the band's last row.
//   (b) long_chain_kernel, one thread: follow the records from (m, n, 0) band by band -> the
//       entry state of every band (one lookup per band).
//   (c) long_detail_kernel, all bands at once: each band re-fills and walks only its own
//       segment of the path (the sequential form's loop, confined to one band).
//   (d) long_pack_kernel: concatenate the segments into the 2-bit op stream.
// If a crossing falls outside the corridor (very dissimilar sequences) the chain sets `failed`
// and the sequential form runs instead: the result is exact either way.
constexpr int X_OUT = (int)0x80000000;
__host__ __device__ inline int x_top(int j, int L) { return j * 4 + L; }
__host__ __device__ inline int x_left(int i) { return -i - 1; }

struct ParTrace {
    const int *run_if;          // nullptr: always; else the X-propagation and the chain run only if *run_if != 0
                                // (second attempt with a wider corridor, only when the first one failed)
    int W;                      // corridor half-width (columns)
    int XW;                     // row stride of xrow (entries)
    int *xrow;                  // nbands x XW x 3
    int *corr;                  // nbands x 2: the corridor [lo, hi] of every band (written by the X-propagation,
                                // so that the serial chain does no 64-bit divisions)
    int *xlast;                 // X_0(m, n)
    int *entry;                 // nbands x 4: i, j, L, kind (0 walk, 1 vertical run only, 2 not on the path)
    long long *seg_off;         // nbands: start of the band's ops in seg_ops
    int *seg_len;               // nbands: ops the band produced
    long long *seg_pos;         // nbands: position of the band's ops in the final stream
    uint8_t *seg_ops;           // one byte per op
    long long seg_cap;
    int *failed;
    int *tiles;
};

__host__ __device__ inline void long_corridor(int band, int m, int n, int W, int &lo, int &hi) {
    const long long r0 = (long long)band * LBAND, r1 = (long long)min((band + 1) * LBAND, m);
    long long c0 = r0 * n / m - W, c1 = (r1 * n + m - 1) / m + W;
    if (c0 < 0) c0 = 0;
    if (c1 > n) c1 = n;
    lo = (int)(c0 / LCC) * LCC;
    hi = (int)c1;
}

__device__ __forceinline__ int selp_s32(bool p, int a, int b) {
    int r;
    asm("{\n\t.reg .pred q;\n\tsetp.ne.s32 q, %3, 0;\n\tselp.s32 %0, %1, %2, q;\n\t}" : "=r"(r) : "r"(a), "r"(b), "r"((int)p));
    return r;
}

template <bool SIMPLE>
__global__ void __launch_bounds__(32 * LWARPS)
long_xprop_kernel(const SchemeDev sc, const LongArgs a, const ParTrace pt) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    int *s_sub = reinterpret_cast<int *>(smem_raw);
    int *s_gh;
    if (pt.run_if && *pt.run_if == 0) return;
    long_load_tables<SIMPLE>(sc, s_sub, s_gh);
    const int g = threadIdx.x & 31;
    const int band = blockIdx.x * LWARPS + (threadIdx.x >> 5);
    if (band >= a.nbands) return;
    const TraceConsts k{4u * sc.o, 4u * sc.mm, 4u * sc.gh + 1u, a.big};
    int lo, hi;
    long_corridor(band, a.m, a.n, pt.W, lo, hi);
    if (g == 0) { pt.corr[2 * band] = lo; pt.corr[2 * band + 1] = hi; }
    const int width = hi - lo;
    const int ib = band * LBAND + g * LR;
    LongLane<SIMPLE> ln;
    unsigned rh_prev = long_restore<SIMPLE>(sc, a, band, lo, g, ln, k.big);
    // X state: x0[r] = X_0 of my row r at the previous column (diagonal input of row r+1),
    // x1[r] = X_1 of my row r at the previous column (left input).  At the corridor's left edge
    // every walk that continues leftwards is LEFT (lo == 0) or OUT.
    int x0[LR], x1[LR];
#pragma unroll
    for (int r = 0; r < LR; ++r) x0[r] = x1[r] = (lo == 0) ? x_left(ib + r + 1) : X_OUT;
    int xh_prev = (g == 0) ? x_top(lo, 0) : ((lo == 0) ? x_left(ib) : X_OUT);
    const bool last_band = band == a.nbands - 1;
    const int gm = ((a.m - 1) % LBAND) / LR, rm = (a.m - 1) % LR;
    int *my_xrow = pt.xrow + (size_t)band * pt.XW * 3;
    const int2 *up_brow = a.brow + (size_t)(band > 0 ? band - 1 : 0) * (a.n + 1);
    unsigned hb = 0, ub = 0, bv = 0, bblk = 0;
    int xb0 = 0, xb2 = 0;
    int2 tblk = make_int2(0, 0);
    const int nsteps = width + 31;
    for (int t = 0; t < nsteps; ++t) {
        if ((t & 31) == 0) {
            const int cj = min(lo + t + 1 + g, a.n);
            bblk = (unsigned)a.s2[cj - 1];
            if (band > 0) {
                tblk = __ldcg(up_brow + cj);
                const int sh = long_shift(a, sc.o, band * LBAND, cj);
                tblk.x += sh;
                tblk.y += sh;
            } else {
                const unsigned rj = (unsigned)a.rowsum[cj];
                tblk = make_int2((int)(min(k.big, rj) << 2), (int)(min(k.big, rj + sc.o) << 2));
            }
        }
        const int jl = t - g + 1;
        unsigned rh = __shfl_up_sync(FULL, hb, 1);
        unsigned ru = __shfl_up_sync(FULL, ub, 1);
        unsigned bin = __shfl_up_sync(FULL, bv, 1);
        int rx0 = __shfl_up_sync(FULL, xb0, 1);
        int rx2 = __shfl_up_sync(FULL, xb2, 1);
        const unsigned b0 = __shfl_sync(FULL, bblk, t & 31);
        const unsigned th = (unsigned)__shfl_sync(FULL, tblk.x, t & 31);
        const unsigned tu = (unsigned)__shfl_sync(FULL, tblk.y, t & 31);
        if (g == 0) {
            bin = b0; rh = th; ru = tu;
            rx0 = x_top(lo + jl, 0);       // leaving through the top of the band at this column
            rx2 = x_top(lo + jl, 2);
        }
        bv = bin;
        if (jl >= 1 && jl <= width) {
            unsigned diag = rh_prev, up = ru;
            rh_prev = rh;
            int xdiag = xh_prev, xup = rx2;
            xh_prev = rx0;
            const unsigned ghT = SIMPLE ? k.ghTc : (unsigned)s_gh[bv];
#pragma unroll
            for (int r = 0; r < LR; ++r) {
                unsigned d0;
                if (SIMPLE) d0 = diag + (ln.aw[r] != bv ? k.mm4 : 0u);
                else d0 = diag + (unsigned)s_sub[ln.aw[r] + bv];
                const unsigned d1 = ln.Lp[r] + ghT;
                const unsigned d2 = up + ln.gvT[r];
                const unsigned hq = (unsigned)__vimin3_s32((int)d0, (int)d1, (int)d2);
                const unsigned Lq = (unsigned)__viaddmin_s32((int)hq, (int)k.o4, (int)d1);
                const unsigned Uq = (unsigned)__viaddmin_s32((int)hq, (int)k.o4, (int)d2);
                // (selp, not ?: -- the compiler turned the nested conditional into a divergent branch
                // per select, seven reconvergence points per row)
                const unsigned c0 = hq & 3u;
                const int X0 = selp_s32(c0 == 0u, xdiag, selp_s32(c0 == 1u, x1[r], xup));
                const int X1 = selp_s32((Lq & 1u) != 0u, x1[r], X0);
                const int X2 = selp_s32((Uq & 2u) != 0u, xup, X0);
                xdiag = x0[r];
                x0[r] = X0;
                x1[r] = X1;
                xup = X2;
                diag = ln.hl[r];
                ln.hl[r] = hq & ~3u;
                ln.Lp[r] = Lq & ~3u;
                up = Uq & ~3u;
            }
            hb = ln.hl[LR - 1];
            ub = up;
            xb0 = x0[LR - 1];
            xb2 = xup;
            if (g == 31) {                  // the band's last row: keep all three records
                int *dst = my_xrow + (size_t)jl * 3;
                dst[0] = x0[LR - 1];
                dst[1] = x1[LR - 1];
                dst[2] = xup;
            }
            if (last_band && g == gm && lo + jl == a.n) {
#pragma unroll
                for (int r = 0; r < LR; ++r)
                    if (r == rm) *pt.xlast = x0[r];
            }
        }
    }
}

// One warp.  The chain is one dependent lookup per band -- X of band b names the column at which
// the walk enters band b-1 -- and a lone thread pays an L2 round trip for each (0.28 ms for the
// 521 bands of a 100 kb pair).  The path moves on by about LBAND * n / m columns per band, so the
// records it will need are predictable: the warp keeps CHAIN_D windows of 32 columns x 3 levels
// in flight (cp.async into shared memory, issued CHAIN_D bands ahead around the extrapolated
// column) and the lookup of a band is a shared-memory read; a miss (the path drifted out of the
// window) falls back to the global load.  All lanes follow the same control flow on the same
// values; lane 0 writes.
constexpr int CHAIN_D = 8;
__global__ void __launch_bounds__(32)
long_chain_kernel(const LongArgs a, const ParTrace pt) {
    constexpr int CORR_CAP = 4096;                  // bands whose corridor is kept in shared memory
    __shared__ int s_win[CHAIN_D][32 * 3];
    __shared__ int s_base[CHAIN_D], s_band[CHAIN_D];
    __shared__ int s_corr[2 * CORR_CAP];
    if (blockIdx.x != 0) return;
    if (pt.run_if && *pt.run_if == 0) return;
    const int lane = threadIdx.x;
    const int nb = a.nbands;
    for (int b = lane; b < nb; b += 32) pt.entry[b * 4 + 3] = 2;
    for (int k = lane; k < 2 * min(nb, CORR_CAP); k += 32) s_corr[k] = pt.corr[k];
    if (lane < CHAIN_D) s_band[lane] = -1;
    __syncwarp();
    auto corr = [&](int band, int k) { return band < CORR_CAP ? s_corr[2 * band + k] : pt.corr[2 * band + k]; };
    const long long dj = (long long)LBAND * a.n / a.m;          // columns per band along the main diagonal
    // request the records of band `band` around entry column `col` (index space: column - lo(band))
    auto request = [&](int band, long long col) {
        if (band >= 0) {
            const int slot = band % CHAIN_D;
            const int lo = corr(band, 0), hi = corr(band, 1);
            int base = (int)max(0LL, min((long long)(hi - lo) - 31, col - lo - 16));
            if (base < 0) base = 0;
            if (lane == 0) { s_base[slot] = base; s_band[slot] = band; }
            const int idx = base + lane;
            if (idx <= hi - lo && idx < pt.XW) {
                const int *src = pt.xrow + ((size_t)band * pt.XW + idx) * 3;
                const unsigned dst = (unsigned)__cvta_generic_to_shared(&s_win[slot][lane * 3]);
#pragma unroll
                for (int k = 0; k < 3; ++k)
                    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" :: "r"(dst + 4u * k), "l"(src + k) : "memory");
            }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    int b = nb - 1, X = *pt.xlast;
    int ei = a.m, ej = a.n, eL = 0;
    long long off = 0;
    bool ok = true;
    // band nb-1 is looked up through xlast; windows for the CHAIN_D bands below it
    for (int d = 1; d <= CHAIN_D; ++d) request(nb - 1 - d, (long long)(nb - d) * LBAND * a.n / a.m);     // the diagonal at the band's last row
    for (;;) {
        const int ib = b * LBAND;
        if (lane == 0) {
            pt.entry[b * 4 + 0] = ei;
            pt.entry[b * 4 + 1] = ej;
            pt.entry[b * 4 + 2] = eL;
            pt.entry[b * 4 + 3] = 0;
        }
        if (X == X_OUT) { ok = false; break; }
        int jx = X < 0 ? 0 : (X >> 2);                 // column at which the walk leaves the band
        long long cap = (long long)(ei - ib) + (ej - jx) + 2;
        if (b == 0 && X >= 0) cap += jx;               // trailing run along row 0
        if (lane == 0) pt.seg_off[b] = off;
        off += cap;
        if (X < 0 || (b > 0 && jx == 0)) {             // reached column 0 (inside band b, or exactly on its upper edge): the rest is vertical
            for (int c = b - 1 - lane; c >= 0; c -= 32) {
                pt.entry[c * 4 + 0] = (c + 1) * LBAND;
                pt.entry[c * 4 + 1] = 0;
                pt.entry[c * 4 + 2] = 2;
                pt.entry[c * 4 + 3] = 1;
                pt.seg_off[c] = off + (long long)(b - 1 - c) * LBAND;
            }
            off += (long long)b * LBAND;
            break;
        }
        if (b == 0) break;
        --b;
        ei = (b + 1) * LBAND;
        ej = jx;
        eL = X & 3;
        // the window of band b was requested CHAIN_D iterations ago: everything but the newest CHAIN_D - 1 groups has landed
        asm volatile("cp.async.wait_group %0;" :: "n"(CHAIN_D - 1) : "memory");
        __syncwarp();
        const int slot = b % CHAIN_D;
        const int lo = corr(b, 0), hi = corr(b, 1);
        if (!(ej > lo && ej <= hi)) { ok = false; break; }
        const int idx = ej - lo, base = s_base[slot];
        if (s_band[slot] == b && idx >= base && idx < base + 32) X = s_win[slot][(idx - base) * 3 + eL];
        else X = pt.xrow[((size_t)b * pt.XW + idx) * 3 + eL];
        __syncwarp();                                   // the slot is read before it is given to band b - CHAIN_D
        request(b - CHAIN_D, (long long)ej - (long long)CHAIN_D * dj);
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    if (off > pt.seg_cap) ok = false;
    if (lane == 0) *pt.failed = ok ? 0 : 1;
}

template <bool SIMPLE>
__global__ void __launch_bounds__(32)
long_detail_kernel(const SchemeDev sc, const LongArgs a, const ParTrace pt) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint4 *s_codes = reinterpret_cast<uint4 *>(smem_raw);
    int *s_sub = reinterpret_cast<int *>(smem_raw + (size_t)LT_NTB * 32 * sizeof(uint4));
    int *s_gh;
    if (*pt.failed) return;
    const int band = blockIdx.x;
    const int kind = pt.entry[band * 4 + 3];
    const int g = threadIdx.x;
    if (kind == 2) { if (g == 0) pt.seg_len[band] = 0; return; }
    uint8_t *seg = pt.seg_ops + pt.seg_off[band];
    const int ib = band * LBAND;
    int i = pt.entry[band * 4 + 0], j = pt.entry[band * 4 + 1], level = pt.entry[band * 4 + 2];
    if (kind == 1) {                                   // column 0: gaps in seq_2 only
        for (int q = g; q < i - ib; q += 32) seg[q] = 2;
        if (g == 0) pt.seg_len[band] = i - ib;
        return;
    }
    long_load_tables<SIMPLE>(sc, s_sub, s_gh);
    const TraceConsts k{4u * sc.o, 4u * sc.mm, 4u * sc.gh + 1u, a.big};
    int cnt = 0, tiles = 0;
    while (i > ib && j > 0) {
        const int jc = ((j - 1) / LCC) * LCC;
        long_refill_codes<SIMPLE>(sc, a, s_sub, s_gh, s_codes, band, jc, j - jc, g, k);
        ++tiles;
        if (g == 0) {
            while (i > ib && j > jc) {
                const int op = long_tile_dir(s_codes, band, jc, i, j, level);
                seg[cnt++] = (uint8_t)op;
                if (op == 0) { --i; --j; } else if (op == 1) { --j; } else { --i; }
                level = op;
            }
        }
        i = __shfl_sync(FULL, i, 0);
        j = __shfl_sync(FULL, j, 0);
        level = __shfl_sync(FULL, level, 0);
        __syncwarp();
    }
    if (g == 0) {
        for (; j == 0 && i > ib; --i) seg[cnt++] = 2;          // reached column 0 inside the band
        if (band == 0) for (; j > 0; --j) seg[cnt++] = 1;      // reached row 0: gaps in seq_1
        pt.seg_len[band] = cnt;
        atomicAdd(pt.tiles, tiles);
    }
}

// positions of the segments in the final stream (walk order: last band first) and the total
__global__ void long_prefix_kernel(const LongArgs a, const ParTrace pt, long long *n_ops, int *status) {
    if (threadIdx.x != 0 || blockIdx.x != 0 || *pt.failed) return;
    long long pos = 0;
    for (int b = a.nbands - 1; b >= 0; --b) {
        pt.seg_pos[b] = pos;
        pos += pt.seg_len[b];
    }
    *n_ops = pos;
    *status = 0;
}

__global__ void long_pack_kernel(const LongArgs a, const ParTrace pt, const long long *n_ops, uint32_t *ops) {
    if (*pt.failed) return;
    const long long total = *n_ops;
    const long long w = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (w * 16 >= total) return;
    // band holding op w*16: seg_pos decreases with the band index
    int lo = 0, hi = a.nbands - 1;
    while (lo < hi) {                                   // smallest band index b with seg_pos[b] <= w*16
        const int mid = (lo + hi) >> 1;
        if (pt.seg_pos[mid] <= w * 16) hi = mid; else lo = mid + 1;
    }
    int b = lo;
    while (b > 0 && pt.seg_len[b] == 0) --b;            // skip empty segments sharing the position
    long long within = w * 16 - pt.seg_pos[b];
    uint32_t word = 0;
    for (int q = 0; q < 16 && w * 16 + q < total; ++q) {
        while (within >= pt.seg_len[b]) { within -= pt.seg_len[b]; --b; }
        word |= (uint32_t)pt.seg_ops[pt.seg_off[b] + within] << (2 * q);
        ++within;
    }
    ops[w] = word;
}

}  // namespace ga
