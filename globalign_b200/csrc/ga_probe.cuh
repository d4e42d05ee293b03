// Integer-ALU roofline microbenchmark kernels (sm_100a); host side: ga_alu_peak in ga_api.cu.
#pragma once
#include <cuda_runtime.h>

namespace ga {

// No memory traffic: how many of the fill kernel's instruction kinds the SMs retire per
// second.  Throughput modes use 8 independent chains per thread:
//   0 IADD   1 VIMNMX3.S16x2   2 VIMNMX3 (s32)   3 LOP3   4 IMAD   5 the packed cell's own mix
//   6 VIADDMNMX.S16x2   7 VIMNMX.U16x2   8 VIADD.16x2 (__vadd2)   9 VIADDMNMX (s32)
// Latency modes use ONE dependent chain per thread (reported as giga-instr/s; with
// 8 warps per scheduler resident the rate is issue-limited unless latency > 8 cycles):
//   10 VIMNMX3.S16x2 chain   11 VIADDMNMX.S16x2 chain   12 IADD chain
template <int MODE>
__global__ void __launch_bounds__(256) alu_kernel(unsigned *out, int iters, unsigned seed) {
    unsigned a[8], b[8], c[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        a[k] = seed + threadIdx.x * 7u + k;
        b[k] = seed * 3u + threadIdx.x + 11u * k;
        c[k] = seed ^ (threadIdx.x * 5u + k);
    }
    if constexpr (MODE >= 10) {
        unsigned x = a[0], y = b[0], z = c[0];
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int k = 0; k < 16; ++k) {
                if (MODE == 10) x = __vimin3_s16x2(x, y, z) + 1u;
                else if (MODE == 11) x = __viaddmin_s16x2(x, y, z);
                else x = x + y;
                if (MODE == 10) { /* the +1 keeps the chain from collapsing; counted below */ }
            }
        }
        out[blockIdx.x * blockDim.x + threadIdx.x] = x ^ y ^ z;
    } else {
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            if (MODE == 0) { a[k] += b[k]; b[k] += a[k]; }
            else if (MODE == 1) { a[k] = __vimin3_s16x2(a[k], b[k], c[k]); b[k] = __vimin3_s16x2(b[k], c[k], a[k]); }
            else if (MODE == 2) { a[k] = (unsigned)__vimin3_s32((int)a[k], (int)b[k], (int)c[k]); b[k] = (unsigned)__vimin3_s32((int)b[k], (int)c[k], (int)a[k]); }
            else if (MODE == 3) { a[k] = (a[k] & b[k]) ^ c[k]; b[k] = (b[k] | c[k]) ^ a[k]; }
            else if (MODE == 4) { a[k] = a[k] * b[k] + c[k]; b[k] = b[k] * c[k] + a[k]; }
            else if (MODE == 6) { a[k] = __viaddmin_s16x2(a[k], b[k], c[k]); b[k] = __viaddmin_s16x2(b[k], c[k], a[k]); }
            else if (MODE == 7) { a[k] = __vminu2(a[k], b[k]); b[k] = __vmaxu2(b[k], c[k] ^ a[k]); }
            else if (MODE == 8) { a[k] = __vadd2(a[k], b[k]); b[k] = __vadd2(b[k], a[k]); }
            else if (MODE == 9) { a[k] = (unsigned)__viaddmin_s32((int)a[k], (int)b[k], (int)c[k]); b[k] = (unsigned)__viaddmin_s32((int)b[k], (int)c[k], (int)a[k]); }
            else {
                // one packed cell: a = diag/h, b = L, c = U  (same instruction mix as fill_kernel)
                const unsigned d0 = __vminu2(a[k] ^ seed, 0x00010001u) * 20u + a[k];
                const unsigned d1 = b[k] + 0x00090009u, d2 = c[k] + 0x000a000au;
                const unsigned hq = __vimin3_s16x2(d0, d1, d2);
                const unsigned Lq = __viaddmin_s16x2(hq, 0x00280028u, d1);
                const unsigned Uq = __viaddmin_s16x2(hq, 0x00280028u, d2);
                const unsigned lu = (Lq & 0x00010001u) | (Uq & ~0x00010001u);
                seed = (seed << 2) | (hq & 0x00030003u);
                seed ^= (seed << 2) | (lu & 0x00030003u);
                a[k] = hq & ~0x00030003u; b[k] = Lq & ~0x00030003u; c[k] = Uq & ~0x00030003u;
            }
        }
    }
    unsigned acc = seed;
#pragma unroll
    for (int k = 0; k < 8; ++k) acc ^= a[k] ^ b[k] ^ c[k];
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    }
}

}  // namespace ga
