// K1: batched three-level DP fill + direction-code store (sm_100a).
//
// Replaces, for every pair of a batch, the reference's make_dp_array
// (globaligner.py:756-821), get_next_best_costs (:317-363) and dp_array_forward
// (:366-392), and produces what dp_array_backward (:395-593) needs instead of
// the value table: three 2-bit direction codes per cell.
//
// One warp fills one tile.  Lanes are split into groups of G; a group owns one
// item (PACKED: two pairs, one per 16-bit half of every register; otherwise one
// pair).  Lane g of a group owns DP rows  pass*G*R + g*R + 1 .. + R  and at step
// t computes column j = t - g + 1 of those rows (anti-diagonal wavefront).  The
// bottom row of a lane (h = min of the three levels, and U = the level-2
// candidate for the row below) travels to lane g+1 by warp shuffle; nothing but
// the letters of seq_2 goes through shared memory (unpacked there from the 8-, 4- or 2-bit
// input stream, see ld_letter).
//
// Arithmetic (all integer, exact):  values are kept multiplied by 4; before a
// three-way minimum the level index (0,1,2) is added in the two low bits, so a
// single VIMNMX3 returns both the minimum and its first-listed argmin -- ties go
// to the lowest level, which is the canonical order of the reference's tie
// table (globaligner.py:599-607).  With (d0,d1,d2) the three level values of a
// cell, o the gap-open cost and hq = min3(d0, d1|1, d2|2):
//     c0 = hq & 3                      (successor on level 0),     h = hq & ~3
//     Lq = min(hq + o, d1|1)           value min(d0+o, d1, d2+o),  L = Lq & ~3
//     Uq = min(hq + o, d2|2)           value min(d0+o, d1+o, d2),  U = Uq & ~3
// (since d1 <= d1+o, min(d0+o,d1,d2+o) = min(h+o, d1); the tag of hq carried
// through the fused add-min VIADDMNMX keeps every tie on the reference's side:
// Lq's tag is c1 in {c0, 1}, Uq's tag is c2 in {c0, 2}).  Stored per cell: c0 and
// x = c1 ^ c2 (2 bits), from which the walker recovers c1 and c2.  The next cells use
//     d0' = h(diag) + sub,  d1' = L(left) + gap_h,  d2' = U(up) + gap_v.
// The boundary row/column use the reference's finite sentinel
// (max_cost+1)*max(m,n) so every value equals the reference's.
#pragma once
#include <type_traits>

#include "ga_common.cuh"

namespace ga {

// VOps<PACKED, UNSIGNED16>: the min / add-min flavour.  Packed SIMPLE kernels compare signed
// halves (values < 2^15); the packed TABLE kernels compare unsigned halves, which doubles the
// range (values < 2^16) -- what 1000-residue proteins under BLOSUM62 need: 4 * 16 * 1000 = 64000.
template <bool PACKED, bool UNS = false>
struct VOps;

template <bool UNS>
struct VOps<false, UNS> {
    static constexpr unsigned TAGM = 3u;
    static __device__ __forceinline__ unsigned min3(unsigned a, unsigned b, unsigned c) {
        return (unsigned)__vimin3_s32((int)a, (int)b, (int)c);
    }
    static __device__ __forceinline__ unsigned addmin(unsigned a, unsigned b, unsigned c) {
        return (unsigned)__viaddmin_s32((int)a, (int)b, (int)c);      // min(a + b, c)
    }
    static __device__ __forceinline__ unsigned pk(unsigned lo, unsigned) { return lo; }
};

template <>
struct VOps<true, true> {
    static constexpr unsigned TAGM = 0x00030003u;
    static __device__ __forceinline__ unsigned min3(unsigned a, unsigned b, unsigned c) {
        return __vimin3_u16x2(a, b, c);
    }
    static __device__ __forceinline__ unsigned addmin(unsigned a, unsigned b, unsigned c) {
        return __viaddmin_u16x2(a, b, c);                             // per half: min(a + b, c)
    }
    static __device__ __forceinline__ unsigned pk(unsigned lo, unsigned hi) { return lo | (hi << 16); }
};

template <>
struct VOps<true, false> {
    static constexpr unsigned TAGM = 0x00030003u;
    static __device__ __forceinline__ unsigned min3(unsigned a, unsigned b, unsigned c) {
        return __vimin3_s16x2(a, b, c);
    }
    static __device__ __forceinline__ unsigned addmin(unsigned a, unsigned b, unsigned c) {
        return __viaddmin_s16x2(a, b, c);                             // per half: min(a + b, c)
    }
    static __device__ __forceinline__ unsigned pk(unsigned lo, unsigned hi) { return lo | (hi << 16); }
};

// BANKED (packed table kernels only): the cost table is replicated per lane -- entry (b, a) of lane L
// lives in word (b * AP + a) * 32 + L, i.e. always in shared-memory bank L -- so the two table
// loads of a cell never conflict (the plain layout averages 3.1 wavefronts per load on random
// protein letters and runs the shared-memory pipe at 78 %).  Costs AP^2 * 128 bytes per block.
// BANKED = 4: 32-bit entries, one wavefront per load.  BANKED = 2: 16-bit entries (half the memory,
// alphabets up to ~28 letters at three blocks per SM); lanes 2k and 2k+1 then share a bank and
// conflict when their entries differ in more than parity: ~1.5 wavefronts per load.
// (Measured and rejected, cfg3 sample, round 2: holding the register allocation to four resident
// blocks per SM with __launch_bounds__(128, 4) -- 128 registers, 16-96 bytes spilled, 16-bit table so
// that four blocks fit -- 1,666 vs 1,947 GCUPS; 12 rows per lane 1,805.  The kernel is issue-bound,
// 3 warps per scheduler already cover most dependency stalls.)
template <bool PACKED, bool SIMPLE, int R, int G, int BANKED = 0>
__global__ void __launch_bounds__(128)
fill_kernel(const SchemeDev sc, const TileDesc *__restrict__ tiles, const UniformTiles ut, int n_tiles,
            int *tile_counter, const int *__restrict__ perm,
            const uint8_t *__restrict__ s1, const long long *__restrict__ off1,
            const uint8_t *__restrict__ s2, const long long *__restrict__ off2,
            uint4 *__restrict__ codes, int2 *bnd, int bnd_stride, int sb_stride,
            int *__restrict__ out_cost, int persistent) {
    using L = Lay<PACKED, R, G>;
    constexpr bool PT = PACKED && !SIMPLE;       // packed table kernel: two pairs per register, costs from tables
    constexpr bool PB = PT && BANKED != 0;       // ... with the lane-replicated (conflict-free) table
    constexpr int EB = BANKED == 2 ? 2 : 4;      // PB: bytes per table entry
    using V = VOps<PACKED, PT>;
    using BT = typename std::conditional<PACKED, uint32_t, uint8_t>::type;
    static_assert(G >= 1 && G <= 32 && R >= 1, "bad configuration");

    extern __shared__ __align__(16) unsigned char smem_raw[];
    int *s_sub = reinterpret_cast<int *>(smem_raw);
    int *s_gh = s_sub;
    size_t table_bytes = 0;
    const int AP = sc.A + 1;                     // PT: letters 0..A-1 plus the zero-cost padding letter A
    unsigned pt_lo = 0, pt_hi = 0;               // PT: shared-space byte addresses of the two tables
    if (PB) {
        unsigned *tab = reinterpret_cast<unsigned *>(smem_raw);          // [b][a][lane], EB bytes per entry
        const int n_ent = AP * AP * 32;
        const int n_words = n_ent * EB / 4;
        s_gh = reinterpret_cast<int *>(tab + n_words);
        for (int k = threadIdx.x; k < n_ent; k += blockDim.x) {
            const int e = k >> 5, b = e / AP, a = e - b * AP;
            const unsigned v = (a < sc.A && b < sc.A) ? (unsigned)sc.sub4[a * sc.A + b] : 0u;
            if (EB == 4) tab[k] = v;
            else reinterpret_cast<unsigned short *>(tab)[k] = (unsigned short)v;
        }
        for (int k = threadIdx.x; k < AP; k += blockDim.x) s_gh[k] = k < sc.A ? sc.gh4t[k] : 1;
        table_bytes = ((size_t)(n_words + AP) * 4 + 15) & ~(size_t)15;
        pt_lo = (unsigned)__cvta_generic_to_shared(tab) + (unsigned)EB * (threadIdx.x & 31);   // my column
        pt_hi = pt_lo;
        __syncthreads();
    } else if (PT) {
        // transposed [b][a], padding letter A in both directions (all zero, so padded cells only
        // copy values and every half stays below 2^16); s_hi is the same table shifted into the
        // upper half so that d0 = diag + s_lo[..] + s_hi[..] is one three-input add
        unsigned *s_lo = reinterpret_cast<unsigned *>(smem_raw);
        unsigned *s_hi = s_lo + AP * AP;
        s_gh = reinterpret_cast<int *>(s_hi + AP * AP);
        for (int k = threadIdx.x; k < AP * AP; k += blockDim.x) {
            const int b = k / AP, a = k - b * AP;
            const unsigned v = (a < sc.A && b < sc.A) ? (unsigned)sc.sub4[a * sc.A + b] : 0u;
            s_lo[k] = v;
            s_hi[k] = v << 16;
        }
        for (int k = threadIdx.x; k < AP; k += blockDim.x) s_gh[k] = k < sc.A ? sc.gh4t[k] : 1;
        table_bytes = ((size_t)(2 * AP * AP + AP) * 4 + 15) & ~(size_t)15;
        pt_lo = (unsigned)__cvta_generic_to_shared(s_lo);
        pt_hi = (unsigned)__cvta_generic_to_shared(s_hi);
        __syncthreads();
    } else if (!SIMPLE) {
        const int n_sub = (sc.A + 1) * sc.A;
        s_gh = s_sub + n_sub;
        for (int k = threadIdx.x; k < n_sub; k += blockDim.x) s_sub[k] = sc.sub4[k];
        for (int k = threadIdx.x; k < sc.A; k += blockDim.x) s_gh[k] = sc.gh4t[k];
        table_bytes = ((size_t)(n_sub + sc.A) * 4 + 15) & ~(size_t)15;
        __syncthreads();
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int grp = lane / G;
    const int g = lane - grp * G;
    const bool lane_ok = grp < L::IPW;
    const int my_item = lane_ok ? grp : 0;
    BT *sb = reinterpret_cast<BT *>(smem_raw + table_bytes) + (size_t)warp * L::IPW * sb_stride;
    const BT *myb = sb + (size_t)my_item * sb_stride;
    const int gwarp = blockIdx.x * (blockDim.x >> 5) + warp;
    int2 *my_bnd = bnd + ((size_t)gwarp * L::IPW + my_item) * bnd_stride;

    const unsigned o4 = V::pk(4u * sc.o, 4u * sc.o);
    const unsigned mm4 = 4u * sc.mm;
    const unsigned ghTc = V::pk(4u * sc.gh + 1u, 4u * sc.gh + 1u);
    const unsigned gvTc = V::pk(4u * sc.gv + 2u, 4u * sc.gv + 2u);
    const unsigned padb = SIMPLE ? (unsigned)PAD_B : (PT ? (unsigned)sc.A : 0u);
    const unsigned blim = SIMPLE ? 0xFFu : (PT ? (unsigned)sc.A : (unsigned)(sc.A - 1));   // largest usable seq_2 letter
    // a*k+b forms that must stay IMADs (FMA pipe); see SchemeDev::k_one
    const unsigned K1 = sc.k_one, K4 = sc.k_four, KM1 = sc.k_minus_one;

    // persistent: a machine-filling grid whose warps claim tiles from an atomic counter (needed
    // for multi-pass tiles: the boundary-row scratch is per resident warp).  Otherwise one tile
    // per warp and the block retires, which lets the block scheduler slip the (high-priority)
    // walk blocks of the previous chunk in between.
    for (int round = 0;; ++round) {
        int tile = gwarp;
        if (persistent) {
            if (lane == 0) tile = atomicAdd(tile_counter, 1);
            tile = __shfl_sync(FULL, tile, 0);
        } else if (round > 0) {
            break;
        }
        if (tile >= n_tiles) break;
        const TileDesc td = tile_desc(tiles, ut, tile, L::PPT);

        // ---- the pair(s) of my item
        const int ia = my_item * L::PPI, ibx = ia + (PACKED ? 1 : 0);
        const bool haveA = lane_ok && ia < td.n_pairs;
        const bool haveB = PACKED && lane_ok && ibx < td.n_pairs;
        const int pA = pair_at(perm, td.first_pair + min(ia, td.n_pairs - 1));
        const int pB = pair_at(perm, td.first_pair + min(ibx, td.n_pairs - 1));
        const long long a1 = pair_off(off1, pA, ut.M), a2 = pair_off(off2, pA, ut.N);
        const int mA = (int)(pair_off(off1, pA + 1, ut.M) - a1), nA = (int)(pair_off(off2, pA + 1, ut.N) - a2);
        long long b1 = a1;
        int mB = mA, nB = nA;
        if (PACKED) {
            b1 = pair_off(off1, pB, ut.M);
            mB = (int)(pair_off(off1, pB + 1, ut.M) - b1);
            nB = (int)(pair_off(off2, pB + 1, ut.N) - pair_off(off2, pB, ut.N));
        }
        const unsigned bigA = (unsigned)(sc.max_cost + 1) * (unsigned)max(mA, nA);
        const unsigned bigB = (unsigned)(sc.max_cost + 1) * (unsigned)max(mB, nB);

        const int nsteps = td.N + G - 1;
        const int nTB = (nsteps + L::S - 1) / L::S;        // store blocks of 2 steps
        const int sb_len = nTB * L::S + 1;

        // ---- stage the seq_2 letters of every item of the tile in shared memory
        __syncwarp();
        for (int q = 0; q < L::IPW; ++q) {
            const int qa = pair_at(perm, td.first_pair + min(q * L::PPI, td.n_pairs - 1));
            const int qb = pair_at(perm, td.first_pair + min(q * L::PPI + (PACKED ? 1 : 0), td.n_pairs - 1));
            const long long oa = pair_off(off2, qa, ut.N), ob = pair_off(off2, qb, ut.N);
            const int na = (int)(pair_off(off2, qa + 1, ut.N) - oa), nb = (int)(pair_off(off2, qb + 1, ut.N) - ob);
            BT *row = sb + (size_t)q * sb_stride;
            for (int c = lane; c < sb_len; c += 32) {
                // table kernels index with the letters: an out-of-range letter (caller error) is
                // clamped so that it cannot address outside the tables
                const unsigned va = c < na ? min(ld_letter(s2, oa + c, sc.seq_bits), blim) : padb;
                if (PACKED) {
                    const unsigned vb = c < nb ? min(ld_letter(s2, ob + c, sc.seq_bits), blim) : padb;
                    row[c] = (BT)(va | (vb << 16));
                } else {
                    row[c] = (BT)va;
                }
            }
        }
        __syncwarp();

        // where the final cell (m, n) of each pair lives
        // (the tile's last pass may run with fewer rows per lane, see tail_rows)
        const int npass = (td.M + L::BAND - 1) / L::BAND;
        const int rt_last = tail_rows<R, G>(td.M);
        const int pmA = (mA - 1) / L::BAND, pmB = (mB - 1) / L::BAND;
        const int rpA = pmA == npass - 1 ? rt_last : R, rpB = pmB == npass - 1 ? rt_last : R;
        const int gmA = ((mA - 1) % L::BAND) / rpA, rmA = ((mA - 1) % L::BAND) % rpA;
        const int gmB = ((mB - 1) % L::BAND) / rpB, rmB = ((mB - 1) % L::BAND) % rpB;
        unsigned costA = 0, costB = 0;

        int colbase = sc.o;    // TABLE: o + sum of gap_v over the rows of earlier passes
        int colbaseB = sc.o;   // PT: the same for pair B
        for (int p = 0; p < npass; ++p) {
            auto run_pass = [&](auto rt_c) {
                constexpr int RT = decltype(rt_c)::value;          // DP rows per lane in this pass
                constexpr int WPPT = (RT + L::RPW - 1) / L::RPW;   // code words per plane per step
                constexpr int NWT = 2 * WPPT;
                const int ib = p * L::BAND + g * RT;    // DP row just above my first row
                unsigned aw[RT], gvT[RT], hl[RT], Lp[RT];
                unsigned awB[PT ? RT : 1];              // PT: byte offset of pair B's row letter in a table row
                int colv[RT + 1];                       // column-0 value of rows ib .. ib+R (level 2)
                int colvB[PT ? RT + 1 : 1];             // PT: the same for pair B (per-letter gap costs)
                if (PT) {
                    int mineA = 0, mineB = 0, gvrA[RT], gvrB[RT];
    #pragma unroll
                    for (int r = 0; r < RT; ++r) {
                        const int row = ib + r;
                        const int la = row < mA ? min((int)ld_letter(s1, a1 + row, sc.seq_bits), sc.A) : sc.A;
                        const int lb = row < mB ? min((int)ld_letter(s1, b1 + row, sc.seq_bits), sc.A) : sc.A;
                        aw[r] = (PB ? 32u * EB : 4u) * (unsigned)la;  // byte offset of the letter in a table row
                        awB[r] = (PB ? 32u * EB : 4u) * (unsigned)lb;
                        gvT[r] = (unsigned)sc.gv4t[la] | ((unsigned)sc.gv4t[lb] << 16);
                        gvrA[r] = sc.gv_raw[la];
                        gvrB[r] = sc.gv_raw[lb];
                        mineA += gvrA[r];
                        mineB += gvrB[r];
                    }
                    int inclA = mineA, inclB = mineB;  // inclusive scans over the lanes of my group
    #pragma unroll
                    for (int d = 1; d < G; d <<= 1) {
                        const int va = __shfl_up_sync(FULL, inclA, d), vb = __shfl_up_sync(FULL, inclB, d);
                        if (g >= d) { inclA += va; inclB += vb; }
                    }
                    const int totA = __shfl_sync(FULL, inclA, grp * G + G - 1);
                    const int totB = __shfl_sync(FULL, inclB, grp * G + G - 1);
                    colv[0] = colbase + inclA - mineA;
                    colvB[0] = colbaseB + inclB - mineB;
    #pragma unroll
                    for (int r = 0; r < RT; ++r) {
                        colv[r + 1] = colv[r] + gvrA[r];
                        colvB[r + 1] = colvB[r] + gvrB[r];
                    }
                    colbase += totA;
                    colbaseB += totB;
                } else if (SIMPLE) {
    #pragma unroll
                    for (int r = 0; r < RT; ++r) {
                        const int row = ib + r;
                        const unsigned la = row < mA ? ld_letter(s1, a1 + row, sc.seq_bits) : (unsigned)PAD_A;
                        if (PACKED) {
                            const unsigned lb = row < mB ? ld_letter(s1, b1 + row, sc.seq_bits) : (unsigned)PAD_A;
                            aw[r] = la | (lb << 16);
                        } else {
                            aw[r] = la;
                        }
                        gvT[r] = gvTc;
                    }
    #pragma unroll
                    for (int r = 0; r <= RT; ++r) colv[r] = sc.o + sc.gv * (ib + r);
                } else {
                    int mine = 0, gvr[RT];
    #pragma unroll
                    for (int r = 0; r < RT; ++r) {
                        const int row = ib + r;
                        const int a = row < mA ? min((int)ld_letter(s1, a1 + row, sc.seq_bits), sc.A) : sc.A;
                        aw[r] = (unsigned)(a * sc.A);
                        gvT[r] = (unsigned)sc.gv4t[a];
                        gvr[r] = sc.gv_raw[a];
                        mine += gvr[r];
                    }
                    int incl = mine;                   // inclusive scan over the lanes of my group
    #pragma unroll
                    for (int d = 1; d < G; d <<= 1) {
                        const int v = __shfl_up_sync(FULL, incl, d);
                        if (g >= d) incl += v;
                    }
                    const int band_total = __shfl_sync(FULL, incl, grp * G + G - 1);
                    colv[0] = colbase + incl - mine;
    #pragma unroll
                    for (int r = 0; r < RT; ++r) colv[r + 1] = colv[r] + gvr[r];
                    colbase += band_total;
                }
    #pragma unroll
                for (int r = 0; r < RT; ++r) {
                    const unsigned c = (unsigned)colv[r + 1];
                    const unsigned cB = PT ? (unsigned)colvB[r + 1] : c;
                    hl[r] = V::pk(min(bigA, c), min(bigB, cB)) << 2;                // min(D(i,0))
                    Lp[r] = V::pk(min(bigA, c + sc.o), min(bigB, cB + sc.o)) << 2;  // level-1 candidate
                }
                unsigned rh_prev = 0;                  // h(ib, 0): diagonal input of my first row at j = 1
                if (ib != 0)
                    rh_prev = V::pk(min(bigA, (unsigned)colv[0]), min(bigB, (unsigned)(PT ? colvB[0] : colv[0]))) << 2;

                unsigned hb = 0, ub = 0, bv = 0;
                int rowj = sc.o;                       // TABLE: running row-0 value (pass 0, lane g == 0)
                int rowjB = sc.o;                      // PT: the same for pair B
                // PT kernels with one group per warp, pass 0: the row-0 values h(0, j), U(0, j) -- running sums
                // of per-letter gap costs, clamped by the sentinel -- used to be computed step by step by
                // lane 0 alone, a divergent stretch of ~50 instructions with two dependent loads in which
                // the other 31 lanes waited (ncu source counters, cfg3: 4 % of the instructions, 10 % of
                // the time).  Now every 32 steps each lane computes them for ONE column (its letters, a
                // warp scan over the gap costs) and lane 0 picks up the column of the step by shuffle.
                constexpr bool ROW0_AHEAD = PT && G == 32;
                unsigned r0h = 0, r0u = 0;             // ROW0_AHEAD: packed h / U of row 0 at column t0 + lane + 1
                int r0baseA = sc.o, r0baseB = sc.o;    // ... running sums up to column t0
                int2 nxt = make_int2(0, 0);
                if (p > 0 && g == 0 && !ROW0_AHEAD) nxt = __ldcg(my_bnd + 1);
                // ROW0_AHEAD, later passes: the boundary row of the pass above, 32 columns per coalesced load, fetched
                // 32 steps before lane 0 needs them (lane l holds column t0 + l + 1)
                int2 bcur = make_int2(0, 0), bnext = make_int2(0, 0);
                if (ROW0_AHEAD && p > 0) bnext = __ldcg(my_bnd + min(lane + 1, td.N));
                uint4 *cbase = codes + td.code_off + (size_t)p * ((nTB + 1) / 2) * (L::WPP * 64);

                for (int tb = 0; tb < nTB; ++tb) {
                    unsigned stg[L::S * NWT];
    #pragma unroll
                    for (int u = 0; u < L::S; ++u) {
                        const int t = tb * L::S + u;
                        const int j = t - g + 1;
                        if (ROW0_AHEAD && p == 0 && (t & 31) == 0) {
                            const int c = t + lane;                        // letter index of column c + 1
                            const unsigned lw = c < sb_len ? (unsigned)myb[c] : (padb | (padb << 16));
                            int ga = (s_gh[lw & 0xFFFFu] - 1) >> 2, gb = (s_gh[lw >> 16] - 1) >> 2;   // (padding letter: 0)
#pragma unroll
                            for (int d = 1; d < 32; d <<= 1) {
                                const int va = __shfl_up_sync(FULL, ga, d), vb = __shfl_up_sync(FULL, gb, d);
                                if (lane >= d) { ga += va; gb += vb; }
                            }
                            const unsigned rjA = (unsigned)(r0baseA + ga), rjB = (unsigned)(r0baseB + gb);
                            r0h = V::pk(min(bigA, rjA), min(bigB, rjB)) << 2;
                            r0u = V::pk(min(bigA, rjA + sc.o), min(bigB, rjB + sc.o)) << 2;
                            r0baseA += __shfl_sync(FULL, ga, 31);
                            r0baseB += __shfl_sync(FULL, gb, 31);
                        }
                        if (ROW0_AHEAD && p > 0 && (t & 31) == 0) {
                            bcur = bnext;
                            bnext = __ldcg(my_bnd + min(t + 32 + lane + 1, td.N));
                        }
                        unsigned row0h = 0, row0u = 0;
                        if (ROW0_AHEAD) {
                            row0h = __shfl_sync(FULL, p == 0 ? r0h : (unsigned)bcur.x, t & 31);
                            row0u = __shfl_sync(FULL, p == 0 ? r0u : (unsigned)bcur.y, t & 31);
                        }
                        unsigned rh = __shfl_up_sync(FULL, hb, 1);
                        unsigned ru = __shfl_up_sync(FULL, ub, 1);
                        unsigned bin = __shfl_up_sync(FULL, bv, 1);
                        const unsigned bnew = (unsigned)myb[t];
                        if (g == 0) {
                            bin = bnew;
                            if (ROW0_AHEAD) {
                                rh = row0h;
                                ru = row0u;
                            } else if (p == 0) {
                                unsigned rj, rjB;      // row-0 value at column t+1 (level 1)
                                if (PT) {
                                    // the padding letter A costs nothing: past the end of a pair the
                                    // running value stays put (those columns are never read back)
                                    const unsigned la = bnew & 0xFFFFu, lb = bnew >> 16;
                                    if (la < (unsigned)sc.A) rowj += sc.gh_raw[la];
                                    if (lb < (unsigned)sc.A) rowjB += sc.gh_raw[lb];
                                    rj = (unsigned)rowj;
                                    rjB = (unsigned)rowjB;
                                } else if (SIMPLE) {
                                    rj = rjB = (unsigned)(sc.o + sc.gh * (t + 1));
                                } else {
                                    rowj += sc.gh_raw[bnew];
                                    rj = rjB = (unsigned)rowj;
                                }
                                rh = V::pk(min(bigA, rj), min(bigB, rjB)) << 2;
                                ru = V::pk(min(bigA, rj + sc.o), min(bigB, rjB + sc.o)) << 2;
                            } else {
                                rh = (unsigned)nxt.x;
                                ru = (unsigned)nxt.y;
                                nxt = __ldcg(my_bnd + min(t + 2, td.N));
                            }
                        }
                        bv = bin;
    #pragma unroll
                        for (int x = 0; x < NWT; ++x) stg[u * NWT + x] = 0u;
                        if (lane_ok && j >= 1 && j <= td.N) {
                            unsigned diag = rh_prev, up = ru;
                            rh_prev = rh;
                            unsigned ghT, rowA = 0, rowB = 0;
                            if (PT) {
                                const unsigned la = bv & 0xFFFFu, lb = bv >> 16;
                                ghT = (unsigned)s_gh[la] | ((unsigned)s_gh[lb] << 16);
                                rowA = pt_lo + la * (unsigned)((PB ? 32 * EB : 4) * AP);
                                rowB = pt_hi + lb * (unsigned)((PB ? 32 * EB : 4) * AP);
                            } else {
                                ghT = SIMPLE ? ghTc : (unsigned)s_gh[bv];
                            }
                            unsigned acc[NWT];
    #pragma unroll
                            for (int x = 0; x < NWT; ++x) acc[x] = 0u;
    #pragma unroll
                            for (int r = 0; r < RT; ++r) {
                                unsigned d0;
                                if (PT) {
                                    // two table loads (address adds as IMADs) and one three-input add
                                    unsigned sa, sb2;
                                    if (PB && EB == 2) {
                                        unsigned short ha, hb2;
                                        asm("ld.shared.u16 %0, [%1];" : "=h"(ha) : "r"(aw[r] * K1 + rowA));
                                        asm("ld.shared.u16 %0, [%1];" : "=h"(hb2) : "r"(awB[r] * K1 + rowB));
                                        sa = ha;
                                        sb2 = hb2;
                                    } else {
                                        asm("ld.shared.u32 %0, [%1];" : "=r"(sa) : "r"(aw[r] * K1 + rowA));
                                        asm("ld.shared.u32 %0, [%1];" : "=r"(sb2) : "r"(awB[r] * K1 + rowB));
                                    }
                                    d0 = PB ? diag + sa + (sb2 << 16) : diag + sa + sb2;
                                } else if (PACKED) {
                                    d0 = __vminu2(aw[r] ^ bv, 0x00010001u) * mm4 + diag;
                                } else if (SIMPLE) {
                                    d0 = diag + (aw[r] != bv ? mm4 : 0u);
                                } else {
                                    d0 = diag + (unsigned)s_sub[aw[r] + bv];
                                }
                                const unsigned d1 = Lp[r] * K1 + ghT;            // IMAD
                                const unsigned d2 = up * K1 + gvT[r];            // IMAD
                                const unsigned hq = V::min3(d0, d1, d2);
                                const unsigned Lq = V::addmin(hq, o4, d1);
                                const unsigned Uq = V::addmin(hq, o4, d2);
                                const int w = r / L::RPW;
                                // plane 0: c0.  A code word takes at most RPW inserts, so a half
                                // never shifts into its neighbour.
                                const unsigned th = hq & V::TAGM;                 // LOP3
                                acc[w] = acc[w] * K4 + th;                       // IMAD
                                // plane 1: x = tag(Lq) ^ tag(Uq), ONE LOP3.  tag(Lq) is c0 or 1 and tag(Uq) is
                                // c0 or 2, so together with c0 the walker recovers both (ga_walk.cuh): c0 = 0:
                                // bit 0 = [c1 == 1], bit 1 = [c2 == 2]; c0 = 1: x != 0 <=> c2 == 2; c0 = 2:
                                // x != 0 <=> c1 == 1.
                                const unsigned lu = (Lq ^ Uq) & V::TAGM;
                                acc[WPPT + w] = acc[WPPT + w] * K4 + lu;     // IMAD
                                diag = hl[r];
                                hl[r] = th * KM1 + hq;                           // hq - tag, IMAD
                                Lp[r] = Lq & ~V::TAGM;
                                up = Uq & ~V::TAGM;
                            }
                            hb = hl[RT - 1];
                            ub = up;
    #pragma unroll
                            for (int x = 0; x < NWT; ++x) stg[u * NWT + x] = acc[x];
                            if (g == G - 1 && p + 1 < npass)
                                __stcg(my_bnd + j, make_int2((int)hb, (int)ub));
                            if (j == nA && p == pmA && g == gmA) {
    #pragma unroll
                                for (int r = 0; r < RT; ++r)
                                    if (r == rmA) costA = hl[r];
                            }
                            if (PACKED && j == nB && p == pmB && g == gmB) {
    #pragma unroll
                                for (int r = 0; r < RT; ++r)
                                    if (r == rmB) costB = hl[r];
                            }
                        }
                    }
                    // one 128-bit store per row-word: both planes of the two staged steps.
                    // sc.code_pairs: blocks 2k and 2k+1 of a lane share a 32-byte sector (4 steps per
                    // sector for the walk); otherwise the words of a block are lane-adjacent.
    #pragma unroll
                    for (int w = 0; w < WPPT; ++w) {
                        const size_t at = sc.code_pairs ? (size_t)((tb >> 1) * WPPT + w) * 64 + 2 * lane + (tb & 1)
                                                        : (size_t)(tb * WPPT + w) * 32 + lane;
                        cbase[at] = make_uint4(stg[w], stg[NWT + w], stg[WPPT + w], stg[NWT + WPPT + w]);
                    }
                }
                __syncwarp();
                    };
            // The last pass of a tile covers only the rows that are left: it runs with the
            // smallest of R/4, R/2, 3R/4, R rows per lane that holds them (tail_rows), so a
            // 600-row pair costs 512 + 96 rows of work instead of 1024.
            const int rt = p == npass - 1 ? rt_last : R;
            if (!TailRows<R, G>::ON || rt == R) run_pass(std::integral_constant<int, R>{});
            else if (rt == TailRows<R, G>::Q3) run_pass(std::integral_constant<int, TailRows<R, G>::Q3>{});
            else if (rt == TailRows<R, G>::Q2) run_pass(std::integral_constant<int, TailRows<R, G>::Q2>{});
            else run_pass(std::integral_constant<int, TailRows<R, G>::Q1>{});
        }
        if (haveA && g == gmA) out_cost[pA] = (int)((PACKED ? (costA & 0xFFFFu) : costA) >> 2);
        if (haveB && g == gmB) out_cost[pB] = (int)((costB >> 16) >> 2);
    }
}

}  // namespace ga
