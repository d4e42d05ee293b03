// Shared device-side definitions of the globalign_b200 kernels (sm_100a).
//
// Data layout in HBM ("Direction-code layout"):
//   * sequences: one uint8 letter index per residue, CSR offsets (int64);
//   * direction codes: the reference's traceback ranks, at every cell, the three
//     level values with an extra gap_open on the levels that differ from the
//     successor's level L, and takes the first-listed minimum c_L
//     (reference globaligner.py:490-514, tie table :595-685).  Because
//     c_1 is either c_0 or 1 and c_2 is either c_0 or 2, a cell needs 4 bits:
//     plane 0 holds c_0 (2 bits), plane 1 holds c_1 ^ c_2 (2 bits; one LOP3 in the
//     fill, and with c_0 the walker recovers c_1 and c_2);
//   * a "tile" is the set of pairs one warp fills together.  The warp's 32
//     lanes are split into groups of G lanes; each lane owns R consecutive DP
//     rows; a group sweeps the columns of its pair(s) as an anti-diagonal
//     wavefront (lane g works on column t-g at step t).  Codes are stored in
//     that skewed (step, lane) order.  The fill stages 2 steps in registers and
//     writes, per row-word w (8 or 16 rows of a lane), one 128-bit word =
//     {plane 0 step 0, plane 0 step 1, plane 1 step 0, plane 1 step 1}; these are
//     laid out [pass][pair of 2-step blocks][row-word][lane 0..31][block of the pair]:
//     the traceback walk, bound by DRAM sectors, finds both planes of FOUR
//     consecutive steps of a cell's row-word in one 32-byte sector, and a store
//     instruction of a warp writes the 16-byte halves of 32 consecutive sectors
//     (the other halves follow one block later and merge in L2).
//     (Measured alternatives: 8 steps of a lane per 32-byte sector made the walk
//     2.4x faster but its strided stores cost the fill 12 %; a 4-step unrolled
//     body cost the fill 6 % -- both rejected.)
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace ga {

struct SchemeDev {
    int A;               // alphabet size (letters 0..A-1)
    int o;               // gap_open_cost
    int max_cost;        // largest costing-matrix entry (reference sentinel, globaligner.py:777)
    int simple;          // 1: sub = (a == b) ? 0 : mm, one gap cost per direction
    int mm;              // SIMPLE: mismatch cost
    int gh, gv;          // SIMPLE: horizontal / vertical gap-extension cost
    const int *sub4;     // TABLE: (A+1)*A ints, 4*cm[a][b]; row A (padding letter) is all zero
    const int *gh4t;     // TABLE: A ints, 4*cm["-"][b] + 1  (level-1 tag folded in)
    const int *gv4t;     // TABLE: A+1 ints, 4*cm[a]["-"] + 2 (level-2 tag), entry A = 2
    const int *gv_raw;   // TABLE: A+1 ints, cm[a]["-"], entry A = 0
    const int *gh_raw;   // TABLE: A ints
    // Runtime multipliers (1, 4, 0xffffffff).  The fill kernel writes its additions,
    // tag clears and code packing as a*k+b with these so that ptxas emits IMAD (FMA
    // pipe) instead of IADD3/LOP3/LEA (ALU pipe): the ALU pipe is the bottleneck of
    // the cell update (profiles/README.md), the FMA pipe would otherwise idle.
    unsigned k_one, k_four, k_minus_one;
    int code_pairs;      // direction-code layout: 1 = two consecutive 2-step blocks of a lane per 32-byte sector
    int seq_bits;        // bits per input letter as the kernels see them: 8 (one letter per byte), 4 or 2 (bit-packed stream)
};

struct TileDesc {
    long long code_off;  // start of this tile's direction codes, in uint4 units (chunk-relative)
    int first_pair;      // position of the tile's first pair in the pair order (see pair_at)
    int n_pairs;         // pairs in the tile (<= items per warp * pairs per item)
    int M, N;            // longest seq_1 / seq_2 in the tile
};

// Letter k of a sequence buffer.  seq_bits = 8: one letter per byte.  4 / 2: the letters form a
// little-endian bit stream (letter k in bits [k*bits, (k+1)*bits) of the byte array), the usual
// packed nucleotide layouts; offsets stay in letters.  (5-bit streams, GA_F_SEQ_5BIT, never reach
// the kernels: the host side unpacks every chunk to bytes on the device first, unpack5_kernel --
// a second load in here for letters that straddle two bytes cost the DNA kernels 2.3 %.)
__device__ __forceinline__ unsigned ld_letter(const uint8_t *__restrict__ s, long long k, int bits) {
    const long long bit = k * bits;
    return ((unsigned)s[bit >> 3] >> (unsigned)(bit & 7)) & ((1u << bits) - 1u);
}

// Letters [off[pair_lo], off[pair_hi]) of a 5-bit letter stream (letter k in bits [5k, 5k + 5)) as
// bytes: out[k] = letter k.  Grid-stride over groups of eight letters (five bytes in, eight out);
// only letters of the range are written (its edges letter by letter), so the chunks of a pipeline
// can unpack concurrently.
__global__ void unpack5_kernel(const uint8_t *__restrict__ packed, const long long *__restrict__ off, long long uni_len,
                               int pair_lo, int pair_hi, uint8_t *__restrict__ out) {
    const long long lo = off ? off[pair_lo] : (long long)pair_lo * uni_len;
    const long long hi = off ? off[pair_hi] : (long long)pair_hi * uni_len;
    const long long q0 = lo >> 3, q1 = (hi + 7) >> 3;
    for (long long q = q0 + blockIdx.x * (long long)blockDim.x + threadIdx.x; q < q1; q += (long long)gridDim.x * blockDim.x) {
        const long long k0 = q << 3;
        if (k0 >= lo && k0 + 8 <= hi) {
            const uint8_t *p = packed + 5 * q;
            const unsigned long long w = (unsigned long long)p[0] | ((unsigned long long)p[1] << 8) | ((unsigned long long)p[2] << 16) |
                                         ((unsigned long long)p[3] << 24) | ((unsigned long long)p[4] << 32);
            uint2 v;
            v.x = (unsigned)(w & 31) | ((unsigned)((w >> 5) & 31) << 8) | ((unsigned)((w >> 10) & 31) << 16) | ((unsigned)((w >> 15) & 31) << 24);
            v.y = (unsigned)((w >> 20) & 31) | ((unsigned)((w >> 25) & 31) << 8) | ((unsigned)((w >> 30) & 31) << 16) | ((unsigned)((w >> 35) & 31) << 24);
            *reinterpret_cast<uint2 *>(out + k0) = v;
        } else {                                   // an edge of the range: only the bytes that hold its letters are read
            for (long long k = max(k0, lo); k < min(k0 + 8, hi); ++k) {
                const long long bit = 5 * k;
                unsigned v = packed[bit >> 3];
                if ((bit & 7) > 3) v |= (unsigned)packed[(bit >> 3) + 1] << 8;
                out[k] = (uint8_t)((v >> (bit & 7)) & 31u);
            }
        }
    }
}

// When every pair of a chunk has the same (m, n) the tile descriptors are arithmetic and are
// not uploaded: tiles == nullptr and this struct (a kernel parameter) describes them all.
struct UniformTiles {
    int M, N;             // the common lengths
    int pair_lo, pair_hi; // pairs of the chunk
    long long u4;         // direction-code words (uint4) per tile
    long long ops_stride; // bytes of a pair's packed-ops slice when the offset arrays are implicit
};

// Offset arrays may be absent (nullptr) for fixed-length batches: pair p then starts at p * len.
__device__ __forceinline__ long long pair_off(const long long *__restrict__ off, long long p, long long len) {
    return off ? off[p] : p * len;
}

// Tiles hold consecutive positions of the planner's pair ORDER; `perm` maps a position to the
// pair's index in the caller's arrays (nullptr: identity).  The planner orders the pairs of a
// chunk by length so that the pairs sharing a tile (and, packed, a register) are alike.
__device__ __forceinline__ int pair_at(const int *__restrict__ perm, int k) { return perm ? perm[k] : k; }

__device__ __forceinline__ TileDesc tile_desc(const TileDesc *tiles, const UniformTiles &ut, int tile, int ppt) {
    if (tiles) return tiles[tile];
    TileDesc td;
    td.code_off = (long long)tile * ut.u4;
    td.first_pair = ut.pair_lo + tile * ppt;
    td.n_pairs = min(ppt, ut.pair_hi - td.first_pair);
    td.M = ut.M;
    td.N = ut.N;
    return td;
}

// Compile-time layout of one kernel configuration.
template <bool PACKED, int R_, int G_>
struct Lay {
    static constexpr int R = R_;                       // DP rows per lane
    static constexpr int G = G_;                       // lanes per group (one item per group)
    static constexpr int RPW = PACKED ? 8 : 16;        // rows per 32-bit code word (per plane)
    static constexpr int WPP = (R + RPW - 1) / RPW;    // words per plane per step
    static constexpr int NW = 2 * WPP;                 // code words per lane per step (2 planes)
    static constexpr int S = 2;                        // steps staged in registers per store block
    static constexpr int NQ = WPP;                     // 128-bit stores per lane per block (one per row-word)
    static constexpr int IPW = 32 / G;                 // items (lane groups) per warp
    static constexpr int PPI = PACKED ? 2 : 1;         // pairs per item
    static constexpr int PPT = IPW * PPI;              // pairs per tile
    static constexpr int BAND = G * R;                 // DP rows per pass
};

// Rows per lane of a tile's LAST pass.  Multi-pass configurations (G = 32, R divisible by 4)
// run it with R/4, R/2, 3R/4 or R rows per lane, whichever is the smallest that holds the
// rows left over after the full passes; the direction codes of that pass use
// ceil(rows / RPW) words per plane per step.  Fill and walk kernels share this function.
template <int R, int G>
struct TailRows {
    static constexpr bool ON = (G == 32) && (R % 4 == 0) && (R >= 8);
    static constexpr int Q1 = ON ? R / 4 : R, Q2 = ON ? R / 2 : R, Q3 = ON ? 3 * R / 4 : R;
};
template <int R, int G>
__host__ __device__ inline int tail_rows(int M) {
    if (!TailRows<R, G>::ON) return R;
    const int band = R * G;
    const int left = M - ((M - 1) / band) * band;        // 1 .. band rows in the last pass
    const int need = (left + G - 1) / G;
    return need <= TailRows<R, G>::Q1 ? TailRows<R, G>::Q1
           : need <= TailRows<R, G>::Q2 ? TailRows<R, G>::Q2
           : need <= TailRows<R, G>::Q3 ? TailRows<R, G>::Q3 : R;
}

// Runtime mirror of Lay<> for host-side planning.
struct LayInfo {
    int packed, R, G, RPW, WPP, NW, S, NQ, IPW, PPI, PPT, BAND;
};
// uint4 words of direction codes of one tile: passes x 2-step blocks x row-words x 32 lanes
__host__ __device__ inline long long tile_code_u4(int M, int N, int G, int BAND, int NW) {
    const long long npass = (M + BAND - 1) / BAND;
    const long long nb2 = (N + G - 1 + 1) / 2;
    const long long nb4 = (nb2 + 1) / 2;                 // 2-step blocks are stored in pairs
    return npass * nb4 * 2 * (NW / 2) * 32;
}
template <bool PACKED, int R, int G>
inline LayInfo lay_info() {
    using L = Lay<PACKED, R, G>;
    return LayInfo{PACKED ? 1 : 0, L::R, L::G, L::RPW, L::WPP, L::NW, L::S, L::NQ, L::IPW,
                   L::PPI, L::PPT, L::BAND};
}

__host__ __device__ inline int ceil_div(int a, int b) { return (a + b - 1) / b; }

// steps of one pass over a tile whose longest seq_2 has N letters
__host__ __device__ inline int tile_steps(int N, int G) { return N + G - 1; }

constexpr unsigned FULL = 0xffffffffu;
constexpr int PAD_A = 0xFE;   // padding letter for rows past the end of seq_1 (SIMPLE kernels)
constexpr int PAD_B = 0xFF;   // padding letter for columns past the end of seq_2

}  // namespace ga
