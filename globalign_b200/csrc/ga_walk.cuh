// K4: traceback walk over the direction codes (sm_100a).
//
// (Measured and rejected: prefetch.global.L2 of the cell 8 diagonal steps ahead made the walk
// 14 % slower, 1.26 -> 1.44 ms per 1M 150 x 150 pairs -- it is bound by DRAM sectors, not latency.)
//
// Replaces the reference's dp_array_backward (globaligner.py:395-593),
// cost_ranks_dispatcher (:595-685, canonical = first listed option) and the four
// take_* emitters (:688-753).  One thread walks one pair: at cell (i, j) with
// the successor node on level L it decodes c_L(i, j) -- plane 0 holds c_0, plane 1
// holds c_1 ^ c_2 (c_1 is c_0 or 1, c_2 is c_0 or 2, so the pair is recoverable) --
// emits it as the operation, moves, and sets L to it.  Reaching row 0 / column 0 ends the
// walk with a run of gap operations (:542-581).  Operations are emitted in walk
// order, 16 per 32-bit word.
//
// The only boundary cell the reference ever *ranks* is the one reached by the
// very first move when m == 1 or n == 1 (:460 tests just (0,0)); that cell is
// re-ranked here from its closed-form values with the reference's finite
// sentinel, and if the reference would not pick the boundary-consistent level
// (it would then index with -1) the pair is flagged status = 1.
#pragma once
#include "ga_common.cuh"

namespace ga {

__device__ __forceinline__ int first_min3(long long v0, long long v1, long long v2) {
    int k = 0;
    long long best = v0;
    if (v1 < best) { best = v1; k = 1; }
    if (v2 < best) { k = 2; }
    return k;
}

template <bool PACKED, int R, int G>
__global__ void __launch_bounds__(128)
walk_kernel(const SchemeDev sc, const TileDesc *__restrict__ tiles, const UniformTiles ut, int pair_lo,
            int n_pairs, const int *__restrict__ perm,
            const uint8_t *__restrict__ s1, const long long *__restrict__ off1,
            const uint8_t *__restrict__ s2, const long long *__restrict__ off2,
            const uint32_t *__restrict__ codes, uint8_t *__restrict__ out_ops,
            const long long *__restrict__ ops_off, int *__restrict__ out_len,
            int *__restrict__ out_status) {
    using L = Lay<PACKED, R, G>;
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n_pairs) return;
    const int p = pair_at(perm, pair_lo + k);  // the pair's index in the caller's arrays
    const int tile = k / L::PPT;               // tiles hold PPT consecutive positions of the chunk's order
    const int within = k - tile * L::PPT;
    const int item = within / L::PPI;
    const int half = within - item * L::PPI;
    const TileDesc td = tile_desc(tiles, ut, tile, L::PPT);
    const long long a1 = pair_off(off1, p, ut.M), a2 = pair_off(off2, p, ut.N);
    const int m = (int)(pair_off(off1, p + 1, ut.M) - a1), n = (int)(pair_off(off2, p + 1, ut.N) - a2);
    const int nsteps = td.N + G - 1;
    const int nB4 = ((nsteps + 1) / 2 + 1) / 2;     // pairs of 2-step blocks
    const uint32_t *cw = codes + 4 * td.code_off;
    const int lane0 = item * G;
    const int hshift = (PACKED && half) ? 16 : 0;

    uint32_t *ow = reinterpret_cast<uint32_t *>(out_ops + pair_off(ops_off, p, ut.ops_stride));
    uint32_t word = 0;
    int nops = 0;
    auto emit = [&](int op) {
        word |= (uint32_t)op << (2 * (nops & 15));
        if ((++nops & 15) == 0) { ow[(nops >> 4) - 1] = word; word = 0; }
    };
    // the tile's last pass may hold fewer rows per lane (tail_rows, shared with the fill kernel)
    const int last_pass = (td.M - 1) / L::BAND;
    const int rt_last = tail_rows<R, G>(td.M);
    const int wpp_last = (rt_last + L::RPW - 1) / L::RPW;
    const float inv_rt_last = 1.0f / (float)rt_last;
    // index (in 32-bit words) of cell (i, j)'s plane-0 entry and the shift of its 2 bits
    auto locate = [&](int i, int j, int &sh) -> size_t {
        const int row = i - 1;
        const int pass = row / L::BAND;
        const int rem = row - pass * L::BAND;
        int g, rp, wpp;
        if (TailRows<R, G>::ON && pass == last_pass) {
            rp = rt_last;
            wpp = wpp_last;
            g = __float2int_rz(((float)rem + 0.5f) * inv_rt_last);    // rem / rp, exact for rem < 2^15
        } else {
            rp = R;
            wpp = L::WPP;
            g = rem / R;
        }
        const int r = rem - g * rp;
        const int t = (j - 1) + g;
        const int w = r / L::RPW;
        const int rows_in_word = min(L::RPW, rp - w * L::RPW);
        sh = 2 * (rows_in_word - 1 - (r - w * L::RPW)) + hshift;
        // 128-bit word (pass, block pair, row-word, lane, block of the pair) = {p0 s0, p0 s1, p1 s0, p1 s1}
        const int tb = t >> 1;
        const size_t pass0 = (size_t)pass * nB4 * L::WPP * 64;            // uint4 words before this pass
        const size_t u4 = sc.code_pairs ? pass0 + ((size_t)(tb >> 1) * wpp + w) * 64 + 2 * (lane0 + g) + (tb & 1)
                                        : pass0 + ((size_t)tb * wpp + w) * 32 + lane0 + g;
        return u4 * 4 + (t & 1);
    };
    auto code_at = [&](int i, int j, int level) -> int {
        int sh;
        const size_t q = locate(i, j, sh);
        auto plane = [&](int pl) -> unsigned { return (cw[q + 2 * pl] >> sh) & 3u; };
        const int c0 = (int)plane(0);
        if (level == 0 || c0 == level) return c0;      // c_L is c0 or L
        // plane 1 holds x = c1 ^ c2 with c1 in {c0, 1}, c2 in {c0, 2}:
        //   c0 = 0: bit 0 of x = [c1 == 1], bit 1 = [c2 == 2];  otherwise (c0 is the other gap
        //   level, whose own code equals c0): x != 0 <=> c_L == L
        const unsigned x = plane(1);
        const bool take = c0 == 0 ? (x & (unsigned)level) != 0 : x != 0;
        return take ? level : c0;
    };

    int i = m, j = n, level = 0, status = 0;
    // first move (globaligner.py:425-458)
    int op = code_at(i, j, level);
    emit(op);
    if (op == 0) { --i; --j; } else if (op == 1) { --j; } else { --i; }
    level = op;
    if (!(i == 0 && j == 0)) {
        if (i == 0 || j == 0) {
            // boundary cell ranked once by the reference (m == 1 or n == 1)
            const long long big = (long long)(sc.max_cost + 1) * max(m, n);
            long long v0 = big, v1 = big, v2 = big;
            int want;
            if (i == 0) {
                long long row = sc.o;
                for (int c = 0; c < j; ++c) row += sc.simple ? sc.gh : sc.gh_raw[min(ld_letter(s2, a2 + c, sc.seq_bits), (unsigned)(sc.A - 1))];
                v1 = row;
                want = 1;
            } else {
                long long col = sc.o;
                for (int c = 0; c < i; ++c) col += sc.simple ? sc.gv : sc.gv_raw[min(ld_letter(s1, a1 + c, sc.seq_bits), (unsigned)sc.A)];
                v2 = col;
                want = 2;
            }
            if (level == 1) { v0 += sc.o; v2 += sc.o; }
            else if (level == 2) { v0 += sc.o; v1 += sc.o; }
            if (first_min3(v0, v1, v2) != want) status = 1;
        }
        if (status == 0) {
            while (i > 0 && j > 0) {
                op = code_at(i, j, level);
                emit(op);
                if (op == 0) { --i; --j; } else if (op == 1) { --j; } else { --i; }
                level = op;
            }
            for (; j > 0; --j) emit(1);       // row 0: gaps in seq_1 (:542-561)
            for (; i > 0; --i) emit(2);       // column 0: gaps in seq_2 (:562-581)
        }
    }
    if (nops & 15) ow[nops >> 4] = word;
    out_len[p] = nops;
    out_status[p] = status;
}

}  // namespace ga
