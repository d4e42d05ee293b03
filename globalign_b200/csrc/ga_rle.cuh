// Compact result stream of a batch (sm_100a): the alignment of every pair as run-length encoded
// operations, packed back to back, plus one 8-byte record per pair.
//
// What the reference returns per pair is three strings and a cost (dp_array_backward,
// globaligner.py:584-593); ga_align_batch hands back the operations those strings are made of
// (2 bits per alignment column in fixed-size slices).  For reads the slices are mostly empty
// -- 150 x 150 pairs reserve 76 bytes for ~39 bytes of operations -- and with eight GPUs behind one
// host the device-to-host copy of those slices is what limits the end-to-end rate.  An alignment of
// similar sequences is a few long runs of "match/mismatch" broken by short gaps, so the same
// information fits in ~10 bytes per pair as runs.
//
//   RLE byte: bits 0-1 = operation (0 match/mismatch column, 1 gap in seq_1, 2 gap in seq_2),
//             bits 2-7 = run length - 1 (runs longer than 64 continue in the next byte),
//   in WALK order (last alignment column first), exactly the order of the 2-bit operations.
//
// Kernels (one thread per pair, after the walk kernel of a chunk):
//   rle_count_kernel   runs of the pair's 2-bit operations -> bytes needed
//   (cub exclusive scan of the byte counts -> offsets relative to the chunk)
//   rle_base_kernel    chunk base = end of the previous chunk (chunks finish in order)
//   rle_write_kernel   the RLE bytes at base + offset; the pair's record and absolute offset
#pragma once
#include "ga_common.cuh"

namespace ga {

struct PairResult {
    int cost;             // alignment cost (min of the corner cell, globaligner.py:425)
    unsigned nops_status; // bits 0-30: alignment columns; bit 31: status (1 = reference-undefined input)
};

// 2-bit internal slices when the caller gives no offsets: bytes of pair p's slice
__device__ __forceinline__ long long ops_slice_bytes(long long m, long long n) {
    return ((((m + n) + 3) / 4) + 3) & ~3LL;
}

__global__ void ops_need_kernel(const long long *__restrict__ off1, const long long *__restrict__ off2, int n_pairs,
                                long long *__restrict__ need) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p < n_pairs) need[p] = ops_slice_bytes(off1[p + 1] - off1[p], off2[p + 1] - off2[p]);
}

// iterate the runs of a pair's 2-bit ops; F(op, run) is called once per RLE byte
template <typename F>
__device__ __forceinline__ int rle_runs(const uint32_t *__restrict__ w, int nops, F emit) {
    int bytes = 0, cur = -1, run = 0;
    uint32_t word = 0;
    for (int k = 0; k < nops; ++k) {
        if ((k & 15) == 0) word = w[k >> 4];
        const int op = (int)(word & 3u);
        word >>= 2;
        if (op == cur && run < 64) {
            ++run;
        } else {
            if (run) { emit(bytes, cur, run); ++bytes; }
            cur = op;
            run = 1;
        }
    }
    if (run) { emit(bytes, cur, run); ++bytes; }
    return bytes;
}

__global__ void rle_count_kernel(int pair_lo, int n_pairs, const uint8_t *__restrict__ ops,
                                 const long long *__restrict__ ops_off, long long ops_stride,
                                 const int *__restrict__ ops_len, const int *__restrict__ status,
                                 int *__restrict__ rle_len) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n_pairs) return;
    const int p = pair_lo + k;
    const uint32_t *w = reinterpret_cast<const uint32_t *>(ops + pair_off(ops_off, p, ops_stride));
    const int nops = status[p] == 0 ? ops_len[p] : 0;      // undefined pairs carry no operations
    rle_len[p] = rle_runs(w, nops, [](int, int, int) {});
}

// base[c + 1] = base[c] + bytes of chunk c (the scan left relative offsets in rel[])
__global__ void rle_base_kernel(long long *base, int c, int pair_lo, int n_pairs, const long long *__restrict__ rel,
                                const int *__restrict__ rle_len) {
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        const int last = pair_lo + n_pairs - 1;
        base[c + 1] = base[c] + rel[last] + rle_len[last];
    }
}

__global__ void rle_write_kernel(int pair_lo, int n_pairs, const uint8_t *__restrict__ ops,
                                 const long long *__restrict__ ops_off, long long ops_stride,
                                 const int *__restrict__ ops_len, const int *__restrict__ status,
                                 const int *__restrict__ cost, const long long *__restrict__ rel,
                                 const long long *__restrict__ base, int c, long long capacity,
                                 uint8_t *__restrict__ rle, long long *__restrict__ out_off,
                                 PairResult *__restrict__ out_res, int *overflow, int last_chunk, long long n_total) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n_pairs) return;
    const int p = pair_lo + k;
    const long long at = base[c] + rel[p];
    out_off[p] = at;
    if (last_chunk && k == n_pairs - 1) out_off[n_total] = base[c + 1];
    PairResult r;
    r.cost = cost[p];
    r.nops_status = (unsigned)ops_len[p] | (status[p] != 0 ? 0x80000000u : 0u);
    out_res[p] = r;
    const uint32_t *w = reinterpret_cast<const uint32_t *>(ops + pair_off(ops_off, p, ops_stride));
    const int nops = status[p] == 0 ? ops_len[p] : 0;
    if (base[c + 1] > capacity) {       // the caller's buffer is too small: write nothing past it
        rle_runs(w, nops, [&](int b, int op, int run) {
            if (at + b < capacity) rle[at + b] = (uint8_t)(op | ((run - 1) << 2));
            else *overflow = 1;
        });
        return;
    }
    rle_runs(w, nops, [&](int b, int op, int run) { rle[at + b] = (uint8_t)(op | ((run - 1) << 2)); });
}

}  // namespace ga
