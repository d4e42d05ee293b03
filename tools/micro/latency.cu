// Single-warp dependent-chain latencies of the instructions the long-pair fill is made of (sm_100a).
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o latency latency.cu && ./latency
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>
__global__ void chain(long long *out, int *sink, int iters, int a, int b) {
    __shared__ int sm[1024];
    for (int k = threadIdx.x; k < 1024; k += blockDim.x) sm[k] = (k * 4 + 4) & 4095;   // pointer chase: next byte offset
    __syncthreads();
    int x = a + threadIdx.x, y = b, z = a * 3, w = b + 7;
    unsigned addr = (unsigned)__cvta_generic_to_shared(sm);
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int k = 0; k < 32; ++k) {
            if (MODE == 0) x = __vimin3_s32(x, y, z) + 0;                 // VIMNMX3
            else if (MODE == 1) x = __viaddmin_s32(x, a, y);            // VIADDMNMX
            else if (MODE == 2) { x = __vimin3_s32(x, y, z); x = __viaddmin_s32(x, a, w); }   // min3 -> addmin
            else if (MODE == 3) x = __shfl_up_sync(0xffffffffu, x, 1);  // SHFL.UP
            else if (MODE == 4) { int v; asm volatile("ld.shared.s32 %0, [%1];" : "=r"(v) : "r"(addr + (unsigned)x)); x = v; }  // LDS
            else if (MODE == 5) x = x * a + b;                          // IMAD
            else if (MODE == 6) { x = __shfl_up_sync(0xffffffffu, x, 1); x = __vimin3_s32(x, y, z); x = __viaddmin_s32(x, a, w); }  // lane-crossing link
            else if (MODE == 7) x = (x == y) ? z : x + 1;               // ISETP + SEL-ish
        }
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) out[0] = t1 - t0;
    sink[threadIdx.x] = x;
}

int main() {
    long long *d;
    int *s;
    cudaMalloc(&d, 8);
    cudaMalloc(&s, 4 * 32);
    const char *names[] = {"VIMNMX3 s32", "VIADDMNMX s32", "VIMNMX3 -> VIADDMNMX (pair)", "SHFL.UP", "LDS (dependent address)",
                           "IMAD", "SHFL.UP -> VIMNMX3 -> VIADDMNMX (triple)", "compare + select"};
    const int iters = 2000;
    for (int mode = 0; mode < 8; ++mode) {
        for (int rep = 0; rep < 2; ++rep) {
            switch (mode) {
                case 0: chain<0><<<1, 32>>>(d, s, iters, 4, 9); break;
                case 1: chain<1><<<1, 32>>>(d, s, iters, 4, 9); break;
                case 2: chain<2><<<1, 32>>>(d, s, iters, 4, 9); break;
                case 3: chain<3><<<1, 32>>>(d, s, iters, 4, 9); break;
                case 4: chain<4><<<1, 32>>>(d, s, iters, 4, 9); break;
                case 5: chain<5><<<1, 32>>>(d, s, iters, 4, 9); break;
                case 6: chain<6><<<1, 32>>>(d, s, iters, 4, 9); break;
                default: chain<7><<<1, 32>>>(d, s, iters, 4, 9); break;
            }
            cudaDeviceSynchronize();
        }
        long long c = 0;
        cudaMemcpy(&c, d, 8, cudaMemcpyDeviceToHost);
        printf("%-45s %.2f cycles per chain link\n", names[mode], (double)c / (iters * 32.0));
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
