// C-ABI of libglobalign_b200.so (see include/globalign_b200.h).
//
// Host side: context, scheme upload, planner (pairs -> tiles -> chunks that fit
// the direction-code scratch), kernel launches on one CUDA stream, CUDA-event
// timing.  No CPU fallback exists: every entry point needs a CUDA device.
#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <atomic>
#include <functional>
#include <chrono>
#include <string>
#include <thread>
#include <vector>

#include "../../include/globalign_b200.h"
#include "ga_fill.cuh"
#include "ga_walk.cuh"
#include "ga_long.cuh"
#include "ga_values.cuh"
#include "ga_probe.cuh"
#include "ga_rle.cuh"
#include <cub/device/device_scan.cuh>

using namespace ga;

// ------------------------------------------------------------------ kernels registry
struct KernelCfg {
    const char *name;
    LayInfo lay;
    bool simple;
    int banked;      // packed table kernel with the lane-replicated cost table: bytes per entry (0 = plain table)
    const void *fill_fn;
    void (*launch_fill)(int grid, size_t smem, cudaStream_t st, const SchemeDev &sc,
                        const TileDesc *tiles, const UniformTiles &ut, int n_tiles, int *counter, const int *perm,
                        const uint8_t *s1,
                        const long long *off1, const uint8_t *s2, const long long *off2,
                        uint4 *codes, int2 *bnd, int bnd_stride, int sb_stride, int *out_cost, int persistent);
    void (*launch_walk)(cudaStream_t st, const SchemeDev &sc, const TileDesc *tiles, const UniformTiles &ut, int pair_lo,
                        int n_pairs, const int *perm, const uint8_t *s1, const long long *off1, const uint8_t *s2,
                        const long long *off2, const uint32_t *codes, uint8_t *out_ops,
                        const long long *ops_off, int *out_len, int *out_status);
};

template <bool P, bool S, int R, int G, int B>
static void launch_fill_t(int grid, size_t smem, cudaStream_t st, const SchemeDev &sc,
                          const TileDesc *tiles, const UniformTiles &ut, int n_tiles, int *counter, const int *perm,
                          const uint8_t *s1,
                          const long long *off1, const uint8_t *s2, const long long *off2,
                          uint4 *codes, int2 *bnd, int bnd_stride, int sb_stride, int *out_cost,
                          int persistent) {
    fill_kernel<P, S, R, G, B><<<grid, 128, smem, st>>>(sc, tiles, ut, n_tiles, counter, perm, s1, off1, s2, off2,
                                                     codes, bnd, bnd_stride, sb_stride, out_cost, persistent);
}
template <bool P, int R, int G>
static void launch_walk_t(cudaStream_t st, const SchemeDev &sc, const TileDesc *tiles, const UniformTiles &ut, int pair_lo,
                          int n_pairs, const int *perm, const uint8_t *s1, const long long *off1, const uint8_t *s2,
                          const long long *off2, const uint32_t *codes, uint8_t *out_ops,
                          const long long *ops_off, int *out_len, int *out_status) {
    const int grid = (n_pairs + 127) / 128;
    walk_kernel<P, R, G><<<grid, 128, 0, st>>>(sc, tiles, ut, pair_lo, n_pairs, perm, s1, off1, s2, off2, codes,
                                               out_ops, ops_off, out_len, out_status);
}
template <bool P, bool S, int R, int G, int B = 0>
static KernelCfg make_cfg(const char *name) {
    KernelCfg c;
    c.name = name;
    c.lay = lay_info<P, R, G>();
    c.simple = S;
    c.banked = B;
    c.fill_fn = (const void *)fill_kernel<P, S, R, G, B>;
    c.launch_fill = launch_fill_t<P, S, R, G, B>;
    c.launch_walk = launch_walk_t<P, R, G>;
    return c;
}

static const std::vector<KernelCfg> &registry() {
    static const std::vector<KernelCfg> cfgs = {
        // packed 16-bit, two pairs per register (simple schemes, short reads)
        make_cfg<true, true, 7, 8>("fill_p16_r7_g8"),
        make_cfg<true, true, 10, 8>("fill_p16_r10_g8"),
        make_cfg<true, true, 13, 8>("fill_p16_r13_g8"),
        make_cfg<true, true, 16, 8>("fill_p16_r16_g8"),
        make_cfg<true, true, 13, 16>("fill_p16_r13_g16"),
        make_cfg<true, true, 19, 16>("fill_p16_r19_g16"),
        make_cfg<true, true, 5, 32>("fill_p16_r5_g32"),
        make_cfg<true, true, 10, 16>("fill_p16_r10_g16"),
        make_cfg<true, true, 15, 10>("fill_p16_r15_g10"),
        make_cfg<true, true, 19, 8>("fill_p16_r19_g8"),
        make_cfg<true, true, 16, 16>("fill_p16_r16_g16"),
        make_cfg<true, true, 8, 32>("fill_p16_r8_g32"),
        make_cfg<true, true, 16, 32>("fill_p16_r16_g32"),
        // packed unsigned 16-bit, two pairs per register, table-driven schemes (proteins up to
        // ~1000 residues under BLOSUM62)
        make_cfg<true, false, 8, 8>("fill_p16t_r8_g8"),
        make_cfg<true, false, 8, 16>("fill_p16t_r8_g16"),
        make_cfg<true, false, 8, 32>("fill_p16t_r8_g32"),
        make_cfg<true, false, 12, 32>("fill_p16t_r12_g32"),
        make_cfg<true, false, 16, 32>("fill_p16t_r16_g32"),
        // ... with the cost table replicated per lane (no shared-memory bank conflicts)
        make_cfg<true, false, 16, 32, 4>("fill_p16b_r16_g32"),
        make_cfg<true, false, 16, 32, 2>("fill_p16h_r16_g32"),      // 16-bit entries
        // int32, simple schemes (compare instead of table)
        make_cfg<false, true, 8, 8>("fill_i32s_r8_g8"),
        make_cfg<false, true, 8, 16>("fill_i32s_r8_g16"),
        make_cfg<false, true, 8, 32>("fill_i32s_r8_g32"),
        // int32, table-driven schemes (BLOSUM, custom matrices, per-letter gaps)
        make_cfg<false, false, 8, 8>("fill_i32t_r8_g8"),
        make_cfg<false, false, 8, 16>("fill_i32t_r8_g16"),
        make_cfg<false, false, 8, 32>("fill_i32t_r8_g32"),
    };
    return cfgs;
}

static const KernelCfg *find_cfg(const char *name) {
    for (const auto &c : registry())
        if (strcmp(c.name, name) == 0) return &c;
    return nullptr;
}

// ------------------------------------------------------------------ context
struct Chunk {
    int tile_lo, tile_hi;     // tiles [lo, hi)
    int pair_lo, pair_hi;     // pairs [lo, hi)
    long long code_u4;        // uint4 words of direction codes
};

// events of one chunk of the pipeline
struct ChunkEvents {
    cudaEvent_t in0, in1, f0, f1, w0, w1, out0, out1;
    cudaEvent_t base, r0, r1, x1;  // compact (RLE) output: chunk size known on the host / its copy / its kernels done
};

struct DevBuf {
    void *p = nullptr;
    size_t cap = 0;
    void *raw = nullptr;     // what cudaMalloc returned (== p unless guard bands are on)
    size_t req = 0;          // guard mode: bytes of the last request (the rear band starts there)
};

// GA_DEBUG_GUARD=1 (tests): every scratch buffer of a context is allocated at exactly the requested
// size between two bands of GUARD_BYTES filled with a pattern, and the bands are checked after
// every call (guard_check) -- compute-sanitizer is not available on the GPU pool, this catches a
// kernel writing past either end of any internal buffer.
constexpr size_t GUARD_BYTES = 4096;
constexpr int GUARD_PATTERN = 0xC5;
static bool guard_mode() {
    static const bool on = getenv("GA_DEBUG_GUARD") != nullptr && atoi(getenv("GA_DEBUG_GUARD")) != 0;
    return on;
}

struct ga_ctx {
    int device = 0;
    cudaStream_t own_stream = nullptr, stream = nullptr;
    // pipeline streams: two fill streams (alternating chunks, so the tail of one chunk
    // overlaps the start of the next), two walk streams, copy-in and copy-out
    cudaStream_t aux[6] = {};
    cudaEvent_t in_piece[32] = {}, in_begin = nullptr;   // host-buffer entry point: upload pieces
    int in_pieces = 1;                  // pieces of the current call
    std::string err;
    int64_t scratch_limit = (int64_t)64 << 30;
    int sm_count = 0;
    // scheme
    bool have_scheme = false;
    SchemeDev sc{};
    bool uniform_gaps = false;
    DevBuf d_sub4, d_gh4t, d_gv4t, d_gvraw, d_ghraw;
    std::vector<int> h_gh, h_gv;       // host copies (prefix sums of the long-pair path)
    // plan
    bool have_plan = false;
    const KernelCfg *cfg = nullptr;
    int64_t n_pairs = 0;
    std::vector<TileDesc> tiles;
    std::vector<int> perm;             // pair order of a ragged plan (empty: identity)
    DevBuf d_perm;
    std::vector<Chunk> chunks;
    int sb_stride = 0, bnd_stride = 0, grid = 0;
    size_t smem = 0;
    int64_t cells = 0, code_bytes = 0;
    uint32_t plan_flags = 0;
    bool serial = false;
    bool uniform = false;             // every pair has the same (m, n): arithmetic tile descriptors
    bool plan_for_io = false;         // planning on behalf of the host-buffer entry point
    bool plan_rle = false;            // ... with the compact (RLE) output
    bool persistent = true;           // fill grid: machine-filling + tile counter, or one tile per warp
    int uniM = 0, uniN = 0;
    long long uni_u4 = 0;
    long long uni_ops_stride = 0;      // bytes per pair of packed ops when offsets are implicit
    DevBuf d_tiles, d_codes[2], d_bnd[2], d_counter;
    int n_bufs = 1;                    // code buffers of the current plan (1 serial, else 2)
    // staging for the host-buffer entry point
    DevBuf d_s1, d_s2, d_off1, d_off2, d_cost, d_ops, d_opsoff, d_len, d_status;
    DevBuf d_u1, d_u2;                  // GA_F_SEQ_5BIT: the letters of the batch unpacked to bytes (chunk by chunk)
    int in_bits = 8;                    // bits per letter of the caller's sequence buffers (5: unpacked before the kernels)
    // long-pair path
    DevBuf l_colsum, l_rowsum, l_brow, l_vchk, l_prog, l_misc, l_xrow, l_meta, l_seg;
    // single-pair value-table paths (ga_fill_values, ga_align_replay): kept between calls
    DevBuf v_s1, v_s2, v_tab, v_draws, v_ops, v_misc;
    // compact (RLE) output of ga_align_batch_rle
    DevBuf r_rle, r_len, r_rel, r_off, r_res, r_base, r_tmp, r_flag;
    long long *h_base = nullptr;        // pinned: running end of the compact stream after each chunk
    size_t h_base_cap = 0;
    int64_t plan_ops_bytes = 0;         // bytes of the internal fixed-slice 2-bit ops of the current plan
    // timing
    cudaEvent_t ev[8] = {};
    std::vector<ChunkEvents> cev;
    float ms[GA_T_COUNT] = {};
    bool timing_valid = false, timing_io = false, timing_rle = false;
};

// GA_DEBUG_TIMING=1: host-side phase times on stderr (diagnostics)
struct HostTimer {
    bool on;
    std::chrono::steady_clock::time_point t0;
    HostTimer() : on(getenv("GA_DEBUG_TIMING") != nullptr), t0(std::chrono::steady_clock::now()) {}
    void lap(const char *what) {
        if (!on) return;
        auto t1 = std::chrono::steady_clock::now();
        fprintf(stderr, "[ga host] %-28s %8.3f ms\n", what, std::chrono::duration<double, std::milli>(t1 - t0).count());
        t0 = t1;
    }
};

// Split [0, n) over a few host threads (planner loops over 10^6 pairs).
template <typename F>
static void parallel_for(int64_t n, F body) {
    const int64_t min_grain = 1 << 16;
    int T = (int)std::min<int64_t>(std::max(1u, std::min(8u, std::thread::hardware_concurrency())),
                                   (n + min_grain - 1) / min_grain);
    if (T <= 1) { body(0, (int64_t)0, n); return; }
    std::vector<std::thread> th;
    for (int t = 0; t < T; ++t) th.emplace_back([=] { body(t, n * t / T, n * (t + 1) / T); });
    for (auto &x : th) x.join();
}

static int fail(ga_ctx *c, int code, const char *fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    if (c) c->err = buf;
    return code;
}
#define CK(call)                                                                              \
    do {                                                                                      \
        cudaError_t e_ = (call);                                                              \
        if (e_ != cudaSuccess)                                                                \
            return fail(ctx, GA_E_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), \
                        __FILE__, __LINE__);                                                  \
    } while (0)

static int ensure(ga_ctx *ctx, DevBuf &b, size_t bytes) {
    if (guard_mode()) {
        const size_t want = (std::max(bytes, (size_t)16) + 15) & ~(size_t)15;
        if (want == b.req && b.p) return GA_OK;
        if (b.raw) CK(cudaFree(b.raw));
        b = DevBuf{};
        CK(cudaMalloc(&b.raw, want + 2 * GUARD_BYTES));
        CK(cudaMemset(b.raw, GUARD_PATTERN, GUARD_BYTES));
        CK(cudaMemset((uint8_t *)b.raw + GUARD_BYTES + want, GUARD_PATTERN, GUARD_BYTES));
        b.p = (uint8_t *)b.raw + GUARD_BYTES;
        b.cap = b.req = want;
        // GA_DEBUG_GUARD=2: self-test -- plant one stray byte behind the buffer; the next check must report it
        if (atoi(getenv("GA_DEBUG_GUARD")) == 2) CK(cudaMemset((uint8_t *)b.p + want, 0, 1));
        return GA_OK;
    }
    if (bytes <= b.cap) return GA_OK;
    if (b.raw) CK(cudaFree(b.raw));
    b = DevBuf{};
    const size_t want = std::max(bytes, (size_t)256);
    CK(cudaMalloc(&b.raw, want));
    b.p = b.raw;
    b.cap = want;
    return GA_OK;
}

static std::vector<std::pair<const char *, DevBuf *>> all_buffers(ga_ctx *ctx);

__global__ void guard_scan_kernel(const uint8_t *front, const uint8_t *rear, int n, int which, int *fault) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    if (front[k] != (uint8_t)GUARD_PATTERN) atomicCAS(fault, 0, 2 * which + 1);
    if (rear[k] != (uint8_t)GUARD_PATTERN) atomicCAS(fault, 0, 2 * which + 2);
}

static int guard_check(ga_ctx *ctx);
static inline int guarded(ga_ctx *ctx, int rc) { return rc == GA_OK ? guard_check(ctx) : rc; }

// after a call has finished on the device: are all guard bands intact?
static int guard_check(ga_ctx *ctx) {
    if (!guard_mode()) return GA_OK;
    CK(cudaDeviceSynchronize());
    int *d_fault = nullptr, fault = 0;
    CK(cudaMalloc(&d_fault, sizeof(int)));
    CK(cudaMemset(d_fault, 0, sizeof(int)));
    const auto bufs = all_buffers(ctx);
    for (size_t k = 0; k < bufs.size(); ++k) {
        const DevBuf *b = bufs[k].second;
        if (!b->raw) continue;
        guard_scan_kernel<<<(int)(GUARD_BYTES / 256), 256>>>((const uint8_t *)b->raw, (const uint8_t *)b->p + b->req,
                                                            (int)GUARD_BYTES, (int)k, d_fault);
    }
    CK(cudaMemcpy(&fault, d_fault, sizeof(int), cudaMemcpyDeviceToHost));
    cudaFree(d_fault);
    if (fault)
        return fail(ctx, GA_E_CUDA, "GA_DEBUG_GUARD: a kernel wrote %s scratch buffer %s", (fault & 1) ? "in front of" : "behind",
                    bufs[(size_t)((fault - 1) / 2)].first);
    return GA_OK;
}

extern "C" int ga_version(void) { return 100; }

extern "C" int ga_device_count(int *count) {
    if (!count) return GA_E_ARG;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) { *count = 0; return GA_E_CUDA; }
    *count = n;
    return GA_OK;
}

extern "C" int ga_ctx_create(int device, ga_ctx **out) {
    if (!out) return GA_E_ARG;
    *out = nullptr;
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n <= 0 || device < 0 || device >= n) return GA_E_CUDA;
    ga_ctx *ctx = new ga_ctx();
    ctx->device = device;
    if (cudaSetDevice(device) != cudaSuccess ||
        cudaStreamCreateWithFlags(&ctx->own_stream, cudaStreamNonBlocking) != cudaSuccess) {
        delete ctx;
        return GA_E_CUDA;
    }
    ctx->stream = ctx->own_stream;
    {   // walk / copy streams outrank the fill streams: their blocks take freed SM slots first
        int lo_pri = 0, hi_pri = 0;
        cudaDeviceGetStreamPriorityRange(&lo_pri, &hi_pri);
        for (int k = 0; k < 6; ++k)
            cudaStreamCreateWithPriority(&ctx->aux[k], cudaStreamNonBlocking, k < 2 ? lo_pri : hi_pri);
    }
    cudaDeviceGetAttribute(&ctx->sm_count, cudaDevAttrMultiProcessorCount, device);
    for (auto &e : ctx->ev) cudaEventCreate(&e);
    for (auto &e : ctx->in_piece) cudaEventCreate(&e);
    cudaEventCreate(&ctx->in_begin);
    *out = ctx;
    return GA_OK;
}

static std::vector<std::pair<const char *, DevBuf *>> all_buffers(ga_ctx *ctx) {
    return {
        {"d_sub4", &ctx->d_sub4},
        {"d_gh4t", &ctx->d_gh4t},
        {"d_gv4t", &ctx->d_gv4t},
        {"d_gvraw", &ctx->d_gvraw},
        {"d_ghraw", &ctx->d_ghraw},
        {"d_tiles", &ctx->d_tiles},
        {"d_perm", &ctx->d_perm},
        {"d_codes[0]", &ctx->d_codes[0]},
        {"d_codes[1]", &ctx->d_codes[1]},
        {"d_bnd[0]", &ctx->d_bnd[0]},
        {"d_bnd[1]", &ctx->d_bnd[1]},
        {"d_counter", &ctx->d_counter},
        {"d_s1", &ctx->d_s1},
        {"d_s2", &ctx->d_s2},
        {"d_u1", &ctx->d_u1},
        {"d_u2", &ctx->d_u2},
        {"d_off1", &ctx->d_off1},
        {"d_off2", &ctx->d_off2},
        {"d_cost", &ctx->d_cost},
        {"d_ops", &ctx->d_ops},
        {"d_opsoff", &ctx->d_opsoff},
        {"d_len", &ctx->d_len},
        {"d_status", &ctx->d_status},
        {"l_colsum", &ctx->l_colsum},
        {"l_rowsum", &ctx->l_rowsum},
        {"l_brow", &ctx->l_brow},
        {"l_vchk", &ctx->l_vchk},
        {"l_prog", &ctx->l_prog},
        {"l_misc", &ctx->l_misc},
        {"l_xrow", &ctx->l_xrow},
        {"l_meta", &ctx->l_meta},
        {"l_seg", &ctx->l_seg},
        {"v_s1", &ctx->v_s1},
        {"v_s2", &ctx->v_s2},
        {"v_tab", &ctx->v_tab},
        {"v_draws", &ctx->v_draws},
        {"v_ops", &ctx->v_ops},
        {"v_misc", &ctx->v_misc},
        {"r_rle", &ctx->r_rle},
        {"r_len", &ctx->r_len},
        {"r_rel", &ctx->r_rel},
        {"r_off", &ctx->r_off},
        {"r_res", &ctx->r_res},
        {"r_base", &ctx->r_base},
        {"r_tmp", &ctx->r_tmp},
        {"r_flag", &ctx->r_flag},
    };
}

extern "C" int ga_ctx_destroy(ga_ctx *ctx) {
    if (!ctx) return GA_OK;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    for (auto &nb : all_buffers(ctx))
        if (nb.second->raw) cudaFree(nb.second->raw);
    if (ctx->h_base) cudaFreeHost(ctx->h_base);
    for (auto &e : ctx->ev) if (e) cudaEventDestroy(e);
    for (auto &e : ctx->in_piece) if (e) cudaEventDestroy(e);
    if (ctx->in_begin) cudaEventDestroy(ctx->in_begin);
    for (auto &c : ctx->cev)
        for (cudaEvent_t e : {c.in0, c.in1, c.f0, c.f1, c.w0, c.w1, c.out0, c.out1, c.base, c.r0, c.r1, c.x1}) cudaEventDestroy(e);
    for (auto &a : ctx->aux) if (a) cudaStreamDestroy(a);
    if (ctx->own_stream) cudaStreamDestroy(ctx->own_stream);
    delete ctx;
    return GA_OK;
}

extern "C" const char *ga_last_error(const ga_ctx *ctx) { return ctx ? ctx->err.c_str() : "null context"; }

extern "C" int ga_ctx_set_stream(ga_ctx *ctx, void *cuda_stream) {
    if (!ctx) return GA_E_ARG;
    ctx->stream = cuda_stream ? (cudaStream_t)cuda_stream : ctx->own_stream;
    return GA_OK;
}

extern "C" int ga_ctx_set_scratch_limit(ga_ctx *ctx, int64_t bytes) {
    if (!ctx || bytes < (1 << 20)) return fail(ctx, GA_E_ARG, "scratch limit must be >= 1 MiB");
    ctx->scratch_limit = bytes;
    ctx->have_plan = false;
    return GA_OK;
}

extern "C" int ga_sync(ga_ctx *ctx) {
    if (!ctx) return GA_E_ARG;
    CK(cudaSetDevice(ctx->device));
    CK(cudaStreamSynchronize(ctx->stream));
    return guard_check(ctx);
}

// ------------------------------------------------------------------ scheme
extern "C" int ga_set_scheme(ga_ctx *ctx, int A, const int32_t *sub, const int32_t *gap_h,
                             const int32_t *gap_v, int32_t gap_open, int32_t max_cost) {
    if (!ctx) return GA_E_ARG;
    if (A < 1 || A > 250 || !sub || !gap_h || !gap_v) return fail(ctx, GA_E_ARG, "bad scheme (1 <= A <= 250)");
    if (gap_open < 0) return fail(ctx, GA_E_ARG, "gap_open must be >= 0");
    int64_t mx = 0;
    for (int k = 0; k < A * A; ++k) {
        if (sub[k] < 0) return fail(ctx, GA_E_ARG, "costs must be non-negative");
        mx = std::max<int64_t>(mx, sub[k]);
    }
    for (int k = 0; k < A; ++k) {
        if (gap_h[k] < 0 || gap_v[k] < 0) return fail(ctx, GA_E_ARG, "costs must be non-negative");
        mx = std::max<int64_t>(mx, std::max(gap_h[k], gap_v[k]));
    }
    if (max_cost < mx) return fail(ctx, GA_E_ARG, "max_cost is smaller than a matrix entry");
    if ((int64_t)max_cost > (1 << 24) || gap_open > (1 << 24)) return fail(ctx, GA_E_RANGE, "costs too large");
    CK(cudaSetDevice(ctx->device));
    // simple scheme: zero diagonal, one off-diagonal value, one gap cost per direction
    bool simple = true, uniform = true;
    int mm = (A > 1) ? sub[1] : 0;
    for (int a = 0; a < A && simple; ++a)
        for (int b = 0; b < A; ++b)
            if (sub[a * A + b] != (a == b ? 0 : mm)) { simple = false; break; }
    for (int k = 1; k < A; ++k)
        if (gap_h[k] != gap_h[0] || gap_v[k] != gap_v[0]) uniform = false;
    simple = simple && uniform;
    std::vector<int> sub4((size_t)(A + 1) * A, 0), gh4t(A), gv4t(A + 1), gvraw(A + 1), ghraw(A);
    for (int k = 0; k < A * A; ++k) sub4[k] = 4 * sub[k];
    for (int k = 0; k < A; ++k) {
        gh4t[k] = 4 * gap_h[k] + 1;
        gv4t[k] = 4 * gap_v[k] + 2;
        gvraw[k] = gap_v[k];
        ghraw[k] = gap_h[k];
    }
    gv4t[A] = 2;
    gvraw[A] = 0;
    int rc;
    if ((rc = ensure(ctx, ctx->d_sub4, sub4.size() * 4))) return rc;
    if ((rc = ensure(ctx, ctx->d_gh4t, gh4t.size() * 4))) return rc;
    if ((rc = ensure(ctx, ctx->d_gv4t, gv4t.size() * 4))) return rc;
    if ((rc = ensure(ctx, ctx->d_gvraw, gvraw.size() * 4))) return rc;
    if ((rc = ensure(ctx, ctx->d_ghraw, ghraw.size() * 4))) return rc;
    CK(cudaStreamSynchronize(ctx->stream));
    CK(cudaMemcpy(ctx->d_sub4.p, sub4.data(), sub4.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(ctx->d_gh4t.p, gh4t.data(), gh4t.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(ctx->d_gv4t.p, gv4t.data(), gv4t.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(ctx->d_gvraw.p, gvraw.data(), gvraw.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(ctx->d_ghraw.p, ghraw.data(), ghraw.size() * 4, cudaMemcpyHostToDevice));
    SchemeDev &sc = ctx->sc;
    sc.A = A;
    sc.o = gap_open;
    sc.max_cost = max_cost;
    sc.simple = simple ? 1 : 0;
    sc.mm = mm;
    sc.gh = gap_h[0];
    sc.gv = gap_v[0];
    sc.sub4 = (const int *)ctx->d_sub4.p;
    sc.gh4t = (const int *)ctx->d_gh4t.p;
    sc.gv4t = (const int *)ctx->d_gv4t.p;
    sc.gv_raw = (const int *)ctx->d_gvraw.p;
    sc.gh_raw = (const int *)ctx->d_ghraw.p;
    sc.k_one = 1u;
    sc.k_four = 4u;
    sc.k_minus_one = 0xffffffffu;
    sc.seq_bits = 8;
    ctx->uniform_gaps = uniform;
    ctx->h_gh.assign(gap_h, gap_h + A);
    ctx->h_gv.assign(gap_v, gap_v + A);
    ctx->have_scheme = true;
    ctx->have_plan = false;
    ctx->timing_valid = false;
    return GA_OK;
}

// ------------------------------------------------------------------ planner
// Does the value range of a packed kernel hold these lengths?
static bool cfg_fits16(const ga_ctx *ctx, const KernelCfg *c, int maxM, int maxN) {
    const SchemeDev &sc = ctx->sc;
    const long long big = (long long)(sc.max_cost + 1) * std::max(maxM, maxN);
    if (c->simple) {
        // largest value any cell (padding included) can take, times 4 plus tag, must fit int16
        const int band = c->lay.BAND, G = c->lay.G;
        const long long Mp = ((maxM + band - 1) / band) * (long long)band, Np = maxN + G;
        const long long vmax = big + (std::min(Mp, Np) + 2) * sc.max_cost + 3LL * sc.o;
        return 4 * vmax + 3 < 32767 && 4LL * sc.mm < 65536;
    }
    // The packed table kernels compare UNSIGNED halves and give padded rows and columns a zero-cost
    // letter, so padded cells only copy values: h <= big + o (a path of max(i,j) steps plus one
    // gap open), L / U <= h + o + max_cost, their inputs d1 / d2 one more step cost.
    return 4 * (big + 2LL * sc.o + 2LL * sc.max_cost) + 3 <= 65535 && sc.A <= 64;
}

static const KernelCfg *choose_cfg(const ga_ctx *ctx, int maxM, int maxN, uint32_t flags) {
    const SchemeDev &sc = ctx->sc;
    auto fits16 = [&](const KernelCfg *c) { return cfg_fits16(ctx, c, maxM, maxN); };
    auto fits16u = [&]() { return cfg_fits16(ctx, find_cfg("fill_p16t_r8_g8"), maxM, maxN); };
    if (const char *force = getenv("GA_FORCE_CFG")) {      // tests and tuning
        const KernelCfg *c = find_cfg(force);
        if (!c || (c->simple && !sc.simple)) return nullptr;
        if (c->lay.packed && !(c->simple ? fits16(c) : fits16u())) return nullptr;
        return c;
    }
    if (sc.simple) {
        if (!(flags & GA_F_FORCE_GENERIC)) {
            // single-pass configurations by rows covered (R x G); measured on B200
            // (tools/tune_cfgs.py): for 150-row reads r19_g8 > r15_g10 > r10_g16 -- few lanes per
            // group (short ramp), many rows per lane.  The smallest band that holds the reads wins.
            const char *pick = maxM <= 56 ? "fill_p16_r7_g8"
                               : maxM <= 80 ? "fill_p16_r10_g8"
                               : maxM <= 104 ? "fill_p16_r13_g8"
                               : maxM <= 128 ? "fill_p16_r16_g8"
                               : maxM <= 152 ? "fill_p16_r19_g8"
                               : maxM <= 160 ? "fill_p16_r10_g16"
                               : maxM <= 208 ? "fill_p16_r13_g16"
                               : maxM <= 256 ? "fill_p16_r16_g16"
                               : maxM <= 304 ? "fill_p16_r19_g16" : "fill_p16_r16_g32";
            const KernelCfg *c = find_cfg(pick);
            if (fits16(c)) return c;
        }
        return find_cfg(maxM <= 64 ? "fill_i32s_r8_g8" : maxM <= 128 ? "fill_i32s_r8_g16" : "fill_i32s_r8_g32");
    }
    if (!(flags & GA_F_FORCE_GENERIC) && fits16u()) {
        if (maxM > 256) {
            // the lane-replicated table needs (A+1)^2 * 128 bytes per block; it pays (+4.5 % on
            // cfg3) only while three blocks still fit an SM, i.e. for alphabets of ~20 letters
            const size_t ent = (size_t)(sc.A + 1) * (sc.A + 1) * 32;
            const size_t stage = (size_t)4 * (((size_t)maxN + 32 + 2 + 1 + 16 + 15) & ~(size_t)15) * 4 + 1024;
            if (3 * (ent * 4 + stage) <= 224 * 1024) return find_cfg("fill_p16b_r16_g32");
            if (3 * (ent * 2 + stage) <= 224 * 1024) return find_cfg("fill_p16h_r16_g32");   // 16-bit entries
            return find_cfg("fill_p16t_r16_g32");
        }
        return find_cfg(maxM <= 64 ? "fill_p16t_r8_g8" : maxM <= 128 ? "fill_p16t_r8_g16" : "fill_p16t_r8_g32");
    }
    return find_cfg(maxM <= 64 ? "fill_i32t_r8_g8" : maxM <= 128 ? "fill_i32t_r8_g16" : "fill_i32t_r8_g32");
}

// dynamic shared memory of a fill kernel: cost table (table-driven schemes) + the staged seq_2
// letters of every item of the block's four warps
constexpr size_t SMEM_LIMIT = 200 * 1024;
static size_t fill_smem_bytes(const ga_ctx *ctx, const KernelCfg *cfg, int maxN) {
    const LayInfo &L = cfg->lay;
    const size_t sb_stride = (size_t)(((maxN + L.G - 1 + L.S + 1 + 16) + 15) & ~15);
    const size_t elt = L.packed ? 4 : 1;
    size_t table = 0;
    if (!cfg->simple) {
        const size_t A = (size_t)ctx->sc.A;
        table = L.packed ? ((2 * (A + 1) * (A + 1) + (A + 1)) * 4 + 15) & ~(size_t)15
                         : (((A + 1) * A + A) * 4 + 15) & ~(size_t)15;
        if (cfg->banked) table = ((A + 1) * (A + 1) * 32 * (size_t)cfg->banked + (A + 1) * 4 + 15) & ~(size_t)15;
    }
    return table + (size_t)4 * L.IPW * sb_stride * elt;
}

static int plan_common(ga_ctx *ctx, const int64_t *s1_off, const int64_t *s2_off, int64_t n_pairs,
                       uint32_t flags, int maxM, int maxN, int minM, int minN, int64_t cells, HostTimer &ht);

// bits per input letter named by the flags (-1: more than one packing flag)
static int seq_bits_of(uint32_t flags) {
    const int n = ((flags & GA_F_SEQ_2BIT) != 0) + ((flags & GA_F_SEQ_4BIT) != 0) + ((flags & GA_F_SEQ_5BIT) != 0);
    if (n > 1) return -1;
    return (flags & GA_F_SEQ_2BIT) ? 2 : (flags & GA_F_SEQ_4BIT) ? 4 : (flags & GA_F_SEQ_5BIT) ? 5 : 8;
}

extern "C" int ga_plan_batch(ga_ctx *ctx, const int64_t *s1_off, const int64_t *s2_off,
                             int64_t n_pairs, const int64_t *out_ops_off, uint32_t flags) {
    if (!ctx) return GA_E_ARG;
    if (!ctx->have_scheme) return fail(ctx, GA_E_STATE, "ga_set_scheme has not been called");
    if (!s1_off || !s2_off || n_pairs < 1 || n_pairs > 0x7fffff00LL)
        return fail(ctx, GA_E_ARG, "bad batch (1 <= n_pairs < 2^31)");
    CK(cudaSetDevice(ctx->device));
    HostTimer ht;
    ctx->have_plan = false;
    int maxM = 0, maxN = 0, minM = 1 << 30, minN = 1 << 30;
    int64_t cells = 0, ops_bytes = 0;
    {
        struct Part { int maxM = 0, maxN = 0, minM = 1 << 30, minN = 1 << 30; int64_t cells = 0, bad = -1, ops_bytes = 0; int why = 0; };
        Part parts[8];
        const bool check_ops = out_ops_off && !(flags & GA_F_COST_ONLY);
        parallel_for(n_pairs, [&](int t, int64_t lo, int64_t hi) {
            Part a;
            for (int64_t p = lo; p < hi; ++p) {
                const int64_t m = s1_off[p + 1] - s1_off[p], n = s2_off[p + 1] - s2_off[p];
                int why = 0;
                if (m < 1 || n < 1) why = 1;
                else if (m > (1 << 24) || n > (1 << 15)) why = 2;
                else if (check_ops) {
                    const int64_t need = (((m + n + 3) / 4) + 3) & ~(int64_t)3;
                    if ((out_ops_off[p] & 3) || out_ops_off[p + 1] - out_ops_off[p] < need) why = 3;
                }
                if (why) { if (a.bad < 0) { a.bad = p; a.why = why; } continue; }
                a.maxM = std::max<int>(a.maxM, (int)m);
                a.maxN = std::max<int>(a.maxN, (int)n);
                a.minM = std::min<int>(a.minM, (int)m);
                a.minN = std::min<int>(a.minN, (int)n);
                a.cells += m * n;
                a.ops_bytes += (((m + n + 3) / 4) + 3) & ~(int64_t)3;
            }
            parts[t] = a;
        });
        for (const Part &a : parts) {
            if (a.bad >= 0) {
                static const char *msg[] = {"", "has an empty sequence",
                    "is too long for the batched kernels (m <= 2^24, n <= 32768); use ga_align_long",
                    "has an ops slice that is too small or misaligned"};
                return fail(ctx, GA_E_ARG, "pair %lld %s", (long long)a.bad, msg[a.why]);
            }
            maxM = std::max(maxM, a.maxM);
            maxN = std::max(maxN, a.maxN);
            minM = std::min(minM, a.minM);
            minN = std::min(minN, a.minN);
            cells += a.cells;
            ops_bytes += a.ops_bytes;
        }
    }
    int rc = plan_common(ctx, s1_off, s2_off, n_pairs, flags, maxM, maxN, minM, minN, cells, ht);
    ctx->plan_ops_bytes = ops_bytes;
    return rc;
}

// Planning for a fixed-length batch without offset arrays (every pair m x n).
static int plan_uniform(ga_ctx *ctx, int64_t m, int64_t n, int64_t n_pairs, uint32_t flags) {
    if (!ctx->have_scheme) return fail(ctx, GA_E_STATE, "ga_set_scheme has not been called");
    if (n_pairs < 1 || n_pairs > 0x7fffff00LL || m < 1 || n < 1 || m > (1 << 24) || n > (1 << 15))
        return fail(ctx, GA_E_ARG, "bad uniform batch (1 <= n_pairs < 2^31, 1 <= m <= 2^24, 1 <= n <= 32768)");
    CK(cudaSetDevice(ctx->device));
    HostTimer ht;
    ctx->have_plan = false;
    int rc = plan_common(ctx, nullptr, nullptr, n_pairs, flags, (int)m, (int)n, (int)m, (int)n, m * n * n_pairs, ht);
    ctx->plan_ops_bytes = ((((m + n) + 3) / 4 + 3) & ~(int64_t)3) * n_pairs;
    return rc;
}

static int plan_common(ga_ctx *ctx, const int64_t *s1_off, const int64_t *s2_off, int64_t n_pairs,
                       uint32_t flags, int maxM, int maxN, int minM, int minN, int64_t cells, HostTimer &ht) {
    for (float &v : ctx->ms) v = 0.f;
    ctx->timing_valid = false;          // the chunk events now belong to a plan that has not run
    {
        const int bits = seq_bits_of(flags);
        if (bits < 0) return fail(ctx, GA_E_ARG, "choose one sequence packing");
        if (ctx->sc.A > (1 << bits)) return fail(ctx, GA_E_ARG, "alphabet of %d letters does not fit %d-bit packed sequences", ctx->sc.A, bits);
        ctx->in_bits = bits;
        ctx->sc.seq_bits = bits == 5 ? 8 : bits;        // 5-bit streams are unpacked to bytes per chunk (run_pipeline)
        if (bits == 5) {
            const int64_t t1 = s1_off ? s1_off[n_pairs] : (int64_t)maxM * n_pairs, t2 = s2_off ? s2_off[n_pairs] : (int64_t)maxN * n_pairs;
            int rc;
            if ((rc = ensure(ctx, ctx->d_u1, (size_t)t1 + 32)) || (rc = ensure(ctx, ctx->d_u2, (size_t)t2 + 32))) return rc;
        }
        const char *e = getenv("GA_CODE_PAIRS");
        ctx->sc.code_pairs = e ? (atoi(e) != 0) : 1;
    }
    ht.lap("plan: scan pairs");
    const KernelCfg *cfg = choose_cfg(ctx, maxM, maxN, flags);
    if (!cfg) return fail(ctx, GA_E_ARG, "no usable kernel configuration (GA_FORCE_CFG names an unknown kernel or one whose value range / scheme kind does not fit)");
    // choose_cfg looks at seq_1 only (rows); the seq_2 letters of every item of a warp are staged
    // in shared memory, so a short seq_1 against a long seq_2 may not fit the few-lanes-per-group
    // kernels (4 items per warp).  Fall back to wider groups of the same family, then from the
    // packed to the int32 family (one byte per staged letter): n <= 32768 always fits the latter.
    if (fill_smem_bytes(ctx, cfg, maxN) > SMEM_LIMIT && !getenv("GA_FORCE_CFG")) {
        const KernelCfg *alt = nullptr;
        for (int pass = 0; pass < 2 && !alt; ++pass)
            for (const auto &c : registry()) {
                if (c.simple != cfg->simple || c.banked) continue;
                if (pass == 0 ? (c.lay.packed != cfg->lay.packed || c.lay.G <= cfg->lay.G) : (c.lay.packed != 0)) continue;
                if (c.lay.packed && !cfg_fits16(ctx, &c, maxM, maxN)) continue;
                if (fill_smem_bytes(ctx, &c, maxN) > SMEM_LIMIT) continue;
                if (!alt || c.lay.G < alt->lay.G || (c.lay.G == alt->lay.G && c.lay.R > alt->lay.R)) alt = &c;
            }
        if (alt) cfg = alt;
    }
    const LayInfo &L = cfg->lay;
    {   // int32 range
        const long long Mp = ((maxM + L.BAND - 1) / L.BAND) * (long long)L.BAND;
        const long long big = (long long)(ctx->sc.max_cost + 1) * std::max(maxM, maxN);
        const long long vmax = big + (std::min<long long>(Mp, maxN + L.G) + 2) * ctx->sc.max_cost + 3LL * ctx->sc.o;
        if (4 * vmax + 3 >= 0x7fffffffLL) return fail(ctx, GA_E_RANGE, "costs could overflow int32 for these lengths");
    }
    if (!cfg->simple && ctx->sc.A > 64)
        return fail(ctx, GA_E_ARG, "table-driven schemes support at most 64 letters");
    // tiles: consecutive groups of PPT positions of the pair order
    const int64_t n_tiles = (n_pairs + L.PPT - 1) / L.PPT;
    // launch geometry (before the chunks are cut: their number depends on how many warps are resident)
    const int warps = 4;
    ctx->sb_stride = ((maxN + L.G - 1 + L.S + 1 + 16) + 15) & ~15;
    ctx->bnd_stride = maxN + 8;
    ctx->smem = fill_smem_bytes(ctx, cfg, maxN);
    if (ctx->smem > SMEM_LIMIT) return fail(ctx, GA_E_ARG, "sequences too long for the batched kernels' shared-memory staging");
    CK(cudaFuncSetAttribute(cfg->fill_fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ctx->smem));
    int per_sm = 0;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, cfg->fill_fn, 128, ctx->smem));
    if (per_sm < 1) return fail(ctx, GA_E_ARG, "fill kernel does not fit on an SM");
    const long long want_blocks = (n_tiles + warps - 1) / warps;
    if (const char *e = getenv("GA_FILL_BLOCKS_PER_SM")) per_sm = std::max(1, std::min(per_sm, atoi(e)));   // experiments
    ctx->grid = (int)std::min<long long>(want_blocks, (long long)per_sm * ctx->sm_count);
    const bool uniform = (minM == maxM && minN == maxN);
    ctx->chunks.clear();
    ctx->perm.clear();
    // Chunks: the direction codes of a chunk live in one of two scratch buffers, so
    // the walk of chunk c overlaps the fill of chunk c+1.  Large batches are cut into
    // ~4 chunks (more if the scratch limit requires); GA_F_SERIAL keeps one stream
    // and as few chunks as the limit allows.
    const bool serial = (flags & GA_F_SERIAL) != 0;
    const long long hard_u4 = std::max<long long>((serial ? ctx->scratch_limit : ctx->scratch_limit / 2) / 16, 1);
    long long total_u4 = 0, uni_u4 = 0, max_chunk_u4 = 0;
    // Where to cut, as cumulative fractions of the batch.  Every chunk boundary costs a fill tail
    // (warps idle until the slowest tile of the chunk is done), and the walk of a chunk only
    // overlaps the next fill in the gaps the persistent fill blocks leave, so chunks must hold
    // many tiles per resident warp (>= ~12; measured on cfg2 and cfg3, tools/probe_cfg3.py):
    // device-resident runs get 1-4 equal chunks.  The host-buffer call also hides its copies
    // behind the kernels: a small first chunk (its upload is all that precedes the first fill),
    // a small last chunk (its download is all that follows the last walk), the rest as above.
    std::vector<double> cuts;
    {
        const double tiles_per_warp = (double)n_tiles / std::max(1, ctx->grid * warps);
        // a boundary costs more when tiles and walks are long (cfg3: ~0.4 ms, cfg2: ~0.05 ms)
        const bool short_pairs = maxM + maxN <= 512;
        const double depth = short_pairs ? (ctx->plan_for_io ? 8.0 : 12.0) : (ctx->plan_for_io ? 20.0 : 40.0);
        int mid = (int)std::max(1.0, std::min(ctx->plan_for_io ? 6.0 : 4.0, tiles_per_warp / depth + 0.5));
        if (const char *e = getenv("GA_CHUNKS")) mid = std::max(1, atoi(e));
        if (serial) {
            cuts = {1.0};
        } else if (ctx->plan_for_io && tiles_per_warp >= 16.0 && !getenv("GA_CHUNKS")) {
            double edge = 0.125;                       // share of the first and of the last chunk
            if (const char *e = getenv("GA_EDGE_CHUNK")) edge = std::min(0.4, std::max(0.01, atof(e)));
            // (the compact output's download is a few bytes per pair: no small last chunk needed)
            bool last_edge = !ctx->plan_rle;
            if (const char *e = getenv("GA_EDGE_LAST")) last_edge = atoi(e) != 0;
            cuts.push_back(edge);
            const double rest = last_edge ? 1.0 - 2 * edge : 1.0 - edge;
            for (int k = 1; k <= mid; ++k) cuts.push_back(edge + rest * k / mid);
            if (last_edge) cuts.push_back(1.0);
        } else {
            for (int k = 1; k <= mid; ++k) cuts.push_back((double)k / mid);
        }
    }
    if (uniform) {
        // every tile is the same: no descriptor array, the kernels compute them (UniformTiles)
        uni_u4 = tile_code_u4(maxM, maxN, L.G, L.BAND, L.NW);
        total_u4 = uni_u4 * n_tiles;
        ctx->tiles.clear();
        if (uni_u4 > hard_u4)
            return fail(ctx, GA_E_NOMEM, "one tile needs %lld bytes of direction codes, above the scratch limit", uni_u4 * 16);
        ht.lap("plan: tiles");
        const int64_t max_tiles = std::max<int64_t>(1, hard_u4 / uni_u4);        // tiles per chunk the scratch allows
        int64_t t = 0;
        for (size_t k = 0; k < cuts.size() && t < n_tiles; ++k) {
            const int64_t stop = k + 1 == cuts.size() ? n_tiles
                                 : std::min<int64_t>(n_tiles, std::max<int64_t>(t + 1, (int64_t)(cuts[k] * (double)n_tiles + 0.5)));
            const int64_t parts = (stop - t + max_tiles - 1) / max_tiles;        // split further if above the limit
            for (int64_t q = 0; q < parts; ++q) {
                const int64_t thi = q + 1 == parts ? stop : t + (stop - t) / (parts - q);
                Chunk c{(int)t, (int)thi, (int)(t * L.PPT), (int)std::min<int64_t>(n_pairs, thi * L.PPT),
                        (thi - t) * uni_u4};
                ctx->chunks.push_back(c);
                max_chunk_u4 = std::max(max_chunk_u4, c.code_u4);
                t = thi;
            }
        }
    } else {
        // Ragged batch.  (1) Cut the batch, in the caller's order, into chunks of whole tiles by
        // an estimate of each pair's share of direction codes (chunks stay contiguous pair ranges
        // so that their results can be copied out as slices).  (2) Inside each chunk order the
        // pairs by length, longest first: the pairs of a tile -- and the two pairs packed into one
        // register -- then have nearly the same lengths, and the padding to the tile's longest
        // pair mostly disappears.  (3) Form tiles over that order.  If a chunk ends up above the
        // scratch limit (the estimate ignores padding) the cut is redone with a lower target.
        const bool sort_pairs = !(flags & GA_F_KEEP_ORDER) && (L.PPT > 1 || maxM > L.BAND);
        std::vector<double> est((size_t)n_pairs);
        double est_total = 0;
        {
            double part[8] = {};
            parallel_for(n_pairs, [&](int tid, int64_t lo, int64_t hi) {
                double tot = 0;
                for (int64_t p = lo; p < hi; ++p) {
                    est[(size_t)p] = (double)tile_code_u4((int)(s1_off[p + 1] - s1_off[p]), (int)(s2_off[p + 1] - s2_off[p]),
                                                         L.G, L.BAND, L.NW) / L.PPT;
                    tot += est[(size_t)p];
                }
                part[tid] = tot;
            });
            for (double v : part) est_total += v;
        }
        ctx->tiles.resize((size_t)n_tiles);
        if (sort_pairs) ctx->perm.resize((size_t)n_pairs);
        TileDesc *tiles = ctx->tiles.data();
        double target = (double)hard_u4 * 0.98;          // estimated codes a chunk may hold
        for (int attempt = 0;; ++attempt) {
            ctx->chunks.clear();
            {   // (1)
                Chunk cur{0, 0, 0, 0, 0};
                double acc = 0, done = 0;
                size_t k = 0;
                for (int64_t t = 0; t < n_tiles; ++t) {
                    double e = 0;
                    const int64_t lo = t * L.PPT, hi = std::min<int64_t>(n_pairs, lo + L.PPT);
                    for (int64_t p = lo; p < hi; ++p) e += est[(size_t)p];
                    const bool at_cut = k + 1 < cuts.size() && done + acc >= cuts[k] * est_total;
                    if ((at_cut || acc + e > target) && acc > 0) {
                        cur.tile_hi = (int)t;
                        cur.pair_hi = (int)lo;
                        ctx->chunks.push_back(cur);
                        cur = Chunk{(int)t, 0, (int)lo, 0, 0};
                        done += acc;
                        acc = 0;
                        while (k + 1 < cuts.size() && done >= cuts[k] * est_total) ++k;
                    }
                    acc += e;
                }
                cur.tile_hi = (int)n_tiles;
                cur.pair_hi = (int)n_pairs;
                ctx->chunks.push_back(cur);
            }
            // (2) + (3), chunks in parallel
            const int64_t n_chunks = (int64_t)ctx->chunks.size();
            std::atomic<int64_t> next{0};
            auto work = [&]() {
                std::vector<int> tmp;
                std::vector<size_t> hist;
                for (;;) {
                    const int64_t c = next.fetch_add(1);
                    if (c >= n_chunks) break;
                    Chunk &ch = ctx->chunks[(size_t)c];
                    if (sort_pairs) {
                        // order by (m, n) descending: two stable counting passes (n, then m), O(pairs)
                        const size_t cnt = (size_t)(ch.pair_hi - ch.pair_lo);
                        int *out = ctx->perm.data() + ch.pair_lo;
                        if (maxM <= (1 << 20)) {
                            tmp.resize(cnt);
                            hist.assign((size_t)std::max(maxM, maxN) + 2, 0);
                            for (size_t k = 0; k < cnt; ++k) {
                                const int64_t p = ch.pair_lo + (int64_t)k;
                                ++hist[(size_t)(maxN - (s2_off[p + 1] - s2_off[p])) + 1];
                            }
                            for (size_t v = 1; v < hist.size(); ++v) hist[v] += hist[v - 1];
                            for (size_t k = 0; k < cnt; ++k) {
                                const int64_t p = ch.pair_lo + (int64_t)k;
                                tmp[hist[(size_t)(maxN - (s2_off[p + 1] - s2_off[p]))]++] = (int)p;
                            }
                            hist.assign(hist.size(), 0);
                            for (size_t k = 0; k < cnt; ++k) {
                                const int64_t p = tmp[k];
                                ++hist[(size_t)(maxM - (s1_off[p + 1] - s1_off[p])) + 1];
                            }
                            for (size_t v = 1; v < hist.size(); ++v) hist[v] += hist[v - 1];
                            for (size_t k = 0; k < cnt; ++k) {
                                const int64_t p = tmp[k];
                                out[hist[(size_t)(maxM - (s1_off[p + 1] - s1_off[p]))]++] = (int)p;
                            }
                        } else {
                            for (size_t k = 0; k < cnt; ++k) out[k] = ch.pair_lo + (int)k;
                            std::stable_sort(out, out + cnt, [&](int x, int y) {
                                const int64_t mx = s1_off[x + 1] - s1_off[x], my = s1_off[y + 1] - s1_off[y];
                                if (mx != my) return mx > my;
                                return s2_off[x + 1] - s2_off[x] > s2_off[y + 1] - s2_off[y];
                            });
                        }
                    }
                    long long u4sum = 0;
                    for (int64_t t = ch.tile_lo; t < ch.tile_hi; ++t) {
                        TileDesc &td = tiles[t];
                        const int64_t lo = t * L.PPT, hi = std::min<int64_t>(n_pairs, lo + L.PPT);
                        int M = 0, N = 0;
                        for (int64_t k = lo; k < hi; ++k) {
                            const int64_t p = sort_pairs ? ctx->perm[(size_t)k] : k;
                            M = std::max<int>(M, (int)(s1_off[p + 1] - s1_off[p]));
                            N = std::max<int>(N, (int)(s2_off[p + 1] - s2_off[p]));
                        }
                        td.first_pair = (int)lo;
                        td.n_pairs = (int)(hi - lo);
                        td.M = M;
                        td.N = N;
                        td.code_off = u4sum;
                        u4sum += tile_code_u4(M, N, L.G, L.BAND, L.NW);
                    }
                    ch.code_u4 = u4sum;
                }
            };
            {
                const int nt = (int)std::min<int64_t>(8, n_chunks);
                std::vector<std::thread> th;
                for (int k = 1; k < nt; ++k) th.emplace_back(work);
                work();
                for (auto &t : th) t.join();
            }
            total_u4 = 0;
            max_chunk_u4 = 0;
            for (const Chunk &ch : ctx->chunks) {
                total_u4 += ch.code_u4;
                max_chunk_u4 = std::max(max_chunk_u4, ch.code_u4);
            }
            if (max_chunk_u4 <= hard_u4) break;
            bool single = false;
            for (const Chunk &ch : ctx->chunks) single = single || (ch.code_u4 > hard_u4 && ch.tile_hi - ch.tile_lo == 1);
            if (single || attempt >= 12)
                return fail(ctx, GA_E_NOMEM, "one tile needs more direction-code scratch than the scratch limit allows");
            target *= 0.7;
        }
        ht.lap("plan: tiles");
    }

    ht.lap("plan: chunks");
    int rc;
    if (!uniform && (rc = ensure(ctx, ctx->d_tiles, (size_t)n_tiles * sizeof(TileDesc)))) return rc;
    if (!ctx->perm.empty() && (rc = ensure(ctx, ctx->d_perm, (size_t)n_pairs * sizeof(int)))) return rc;
    ht.lap("plan: occupancy query");
    const int n_bufs = (serial || ctx->chunks.size() == 1) ? 1 : (int)std::min<size_t>(2, ctx->chunks.size());
    ctx->n_bufs = n_bufs;
    for (int b = 0; b < n_bufs; ++b) {
        if ((rc = ensure(ctx, ctx->d_codes[b], (size_t)max_chunk_u4 * 16))) return rc;
        if ((rc = ensure(ctx, ctx->d_bnd[b], (size_t)ctx->grid * warps * L.IPW * ctx->bnd_stride * sizeof(int2)))) return rc;
    }
    if ((rc = ensure(ctx, ctx->d_counter, ctx->chunks.size() * sizeof(int) + 256))) return rc;
    if (!uniform) {
        CK(cudaMemcpyAsync(ctx->d_tiles.p, ctx->tiles.data(), (size_t)n_tiles * sizeof(TileDesc),
                           cudaMemcpyHostToDevice, ctx->stream));
        if (!ctx->perm.empty())
            CK(cudaMemcpyAsync(ctx->d_perm.p, ctx->perm.data(), (size_t)n_pairs * sizeof(int),
                               cudaMemcpyHostToDevice, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));  // ctx->tiles may be re-planned by the caller
    }
    ht.lap("plan: scratch + tile upload");
    while (ctx->cev.size() < ctx->chunks.size()) {
        ChunkEvents c;
        for (cudaEvent_t *e : {&c.in0, &c.in1, &c.f0, &c.f1, &c.w0, &c.w1, &c.out0, &c.out1, &c.base, &c.r0, &c.r1, &c.x1}) CK(cudaEventCreate(e));
        ctx->cev.push_back(c);
    }
    ctx->cfg = cfg;
    ctx->n_pairs = n_pairs;
    ctx->cells = cells;
    ctx->code_bytes = total_u4 * 16;
    ctx->plan_flags = flags;
    ctx->serial = serial;
    ctx->uniform = uniform;
    // one-tile-per-warp grids measured slower than the persistent grid (10.7 vs 10.1 ms per 1M
    // pairs); kept behind an environment switch for experiments.  Multi-pass tiles always need
    // the persistent form (per-warp boundary scratch).
    ctx->persistent = maxM > L.BAND || getenv("GA_NONPERSISTENT_FILL") == nullptr;
    ctx->uniM = maxM;
    ctx->uniN = maxN;
    ctx->uni_u4 = uni_u4;
    ctx->uni_ops_stride = ((((long long)maxM + maxN + 3) / 4) + 3) & ~3LL;
    ctx->have_plan = true;
    return GA_OK;
}

extern "C" int ga_get_plan_info(const ga_ctx *ctx, ga_plan_info *info) {
    if (!ctx || !info || !ctx->have_plan) return GA_E_STATE;
    memset(info, 0, sizeof *info);
    snprintf(info->fill_kernel, sizeof info->fill_kernel, "%s", ctx->cfg->name);
    info->rows_per_lane = ctx->cfg->lay.R;
    info->lanes_per_group = ctx->cfg->lay.G;
    info->packed16 = ctx->cfg->lay.packed;
    info->launches_per_run = (int)ctx->chunks.size() * ((ctx->plan_flags & GA_F_COST_ONLY) ? 1 : 2);
    info->code_bytes = ctx->code_bytes;
    info->cells = ctx->cells;
    info->n_chunks = (int64_t)ctx->chunks.size();
    return GA_OK;
}

// Host buffers of the host-buffer entry point (null for ga_run_batch).
constexpr int MAX_IN_PIECES = 32;   // the letters are uploaded in up to this many pieces (about 1 MB each), started before planning

struct HostIO {
    int64_t total1 = 0, total2 = 0;     // letters in s1 / s2 (piece k ends at total * (k+1) / pieces)
    int pieces = 1;
    const uint8_t *s1 = nullptr, *s2 = nullptr;
    const int64_t *off1 = nullptr, *off2 = nullptr, *ops_off = nullptr;
    int32_t *cost = nullptr, *len = nullptr, *status = nullptr;
    uint8_t *ops = nullptr;
    // compact (RLE) output instead of cost / ops / len / status
    PairResult *res = nullptr;
    uint8_t *rle = nullptr;
    int64_t *rle_off = nullptr;
    int64_t rle_cap = 0;
};

// The chunk pipeline.  For chunk c (buffers / streams chosen by c & 1):
//   copy-in (host entry only)  ->  fill  ->  walk  ->  copy-out (host entry only)
// fill(c) waits for walk(c-2) (it reuses that chunk's code buffer).  (Three rotating buffers were
// measured: the 0.12 ms bubble per pair of chunks disappears from the timeline, GA_DEBUG_TIMELINE=1,
// but the step time does not change -- fill and walk share the same SM slots either way.)  Everything is
// bracketed by ev[0] / ev[1] on the caller-visible stream, which finally waits for
// all auxiliary streams.
static int run_pipeline(ga_ctx *ctx, const uint8_t *d_s1_in, const int64_t *d_s1_off, const uint8_t *d_s2_in,
                        const int64_t *d_s2_off, int32_t *d_cost, uint8_t *d_ops, const int64_t *d_ops_off,
                        int32_t *d_len, int32_t *d_status, const HostIO *io) {
    const KernelCfg *cfg = ctx->cfg;
    // (5-bit letter streams are unpacked chunk by chunk into d_u1 / d_u2, see below)
    const uint8_t *d_s1 = ctx->in_bits == 5 ? (const uint8_t *)ctx->d_u1.p : d_s1_in;
    const uint8_t *d_s2 = ctx->in_bits == 5 ? (const uint8_t *)ctx->d_u2.p : d_s2_in;
    const bool cost_only = (ctx->plan_flags & GA_F_COST_ONLY) != 0;
    const bool serial = ctx->serial || ctx->chunks.size() == 1;
    cudaStream_t main = ctx->stream;
    const size_t nc = ctx->chunks.size();
    CK(cudaMemsetAsync(ctx->d_counter.p, 0, nc * sizeof(int), main));
    CK(cudaEventRecord(ctx->ev[0], main));
    if (!serial || io)
        for (auto &a : ctx->aux) CK(cudaStreamWaitEvent(a, ctx->ev[0], 0));
    for (size_t c = 0; c < nc; ++c) {
        const Chunk &ch = ctx->chunks[c];
        ChunkEvents &E = ctx->cev[c];
        const int b = serial ? 0 : (int)(c % (size_t)ctx->n_bufs);      // code / boundary buffer
        const int s2i = (int)(c & 1);                                   // stream pair
        cudaStream_t sF = serial ? main : ctx->aux[s2i], sW = serial ? main : ctx->aux[2 + s2i];
        cudaStream_t sOut = ctx->aux[5];
        const int n_tiles = ch.tile_hi - ch.tile_lo, np = ch.pair_hi - ch.pair_lo;
        if (io) {   // the letters are already in flight (ga_align_batch): wait for the piece that
                    // holds this chunk's last letter of either sequence set
            const int64_t a1 = io->off1 ? io->off1[ch.pair_hi] : (int64_t)ch.pair_hi * ctx->uniM;
            const int64_t b1 = io->off2 ? io->off2[ch.pair_hi] : (int64_t)ch.pair_hi * ctx->uniN;
            const int P = io->pieces;
            auto piece_of = [P](int64_t end, int64_t total) {
                int k = 0;
                while (k < P - 1 && total * (k + 1) / P < end) ++k;
                return k;
            };
            const int k = std::max(piece_of(a1, io->total1), piece_of(b1, io->total2));
            CK(cudaStreamWaitEvent(sF, ctx->in_piece[k], 0));
        }
        if (!serial && c >= (size_t)ctx->n_bufs) CK(cudaStreamWaitEvent(sF, ctx->cev[c - ctx->n_bufs].w1, 0));   // code buffer free
        CK(cudaEventRecord(E.f0, sF));
        if (ctx->in_bits == 5) {    // this chunk's letters as bytes; the kernels below read those
            const int ug = ctx->sm_count * 8;
            unpack5_kernel<<<ug, 256, 0, sF>>>(d_s1_in, (const long long *)d_s1_off, ctx->uniM, ch.pair_lo, ch.pair_hi, (uint8_t *)ctx->d_u1.p);
            unpack5_kernel<<<ug, 256, 0, sF>>>(d_s2_in, (const long long *)d_s2_off, ctx->uniN, ch.pair_lo, ch.pair_hi, (uint8_t *)ctx->d_u2.p);
            CK(cudaGetLastError());
        }
        const bool persistent = ctx->persistent || serial;
        const int grid = persistent ? std::min(ctx->grid, (n_tiles + 3) / 4) : (n_tiles + 3) / 4;
        const TileDesc *d_tiles = ctx->uniform ? nullptr : (const TileDesc *)ctx->d_tiles.p + ch.tile_lo;
        const UniformTiles ut{ctx->uniM, ctx->uniN, ch.pair_lo, ch.pair_hi, ctx->uni_u4, ctx->uni_ops_stride};
        const int *d_perm = ctx->perm.empty() ? nullptr : (const int *)ctx->d_perm.p;
        cfg->launch_fill(grid, ctx->smem, sF, ctx->sc, d_tiles, ut,
                         n_tiles, (int *)ctx->d_counter.p + c, d_perm, d_s1, (const long long *)d_s1_off, d_s2,
                         (const long long *)d_s2_off, (uint4 *)ctx->d_codes[b].p, (int2 *)ctx->d_bnd[b].p,
                         ctx->bnd_stride, ctx->sb_stride, d_cost, persistent ? 1 : 0);
        CK(cudaGetLastError());
        CK(cudaEventRecord(E.f1, sF));
        if (sW != sF) CK(cudaStreamWaitEvent(sW, E.f1, 0));
        CK(cudaEventRecord(E.w0, sW));
        if (!cost_only) {
            cfg->launch_walk(sW, ctx->sc, d_tiles, ut, ch.pair_lo, np, d_perm,
                             d_s1, (const long long *)d_s1_off, d_s2, (const long long *)d_s2_off,
                             (const uint32_t *)ctx->d_codes[b].p, d_ops, (const long long *)d_ops_off,
                             d_len, d_status);
            CK(cudaGetLastError());
        }
        CK(cudaEventRecord(E.w1, sW));          // the chunk's direction codes are free again from here
        if (io && io->rle) {
            // compact result stream (ga_rle.cuh): runs per pair, scan, chunk base, write.  The chunk
            // base chains through the chunks, which finish in order but on alternating streams.
            if (c > 0 && !serial) CK(cudaStreamWaitEvent(sW, ctx->cev[c - 1].x1, 0));
            const int g1 = (np + 255) / 256;
            rle_count_kernel<<<g1, 256, 0, sW>>>(ch.pair_lo, np, d_ops, (const long long *)d_ops_off, ctx->uni_ops_stride,
                                                 d_len, d_status, (int *)ctx->r_len.p);
            size_t tmp_bytes = ctx->r_tmp.cap;
            CK(cub::DeviceScan::ExclusiveScan(ctx->r_tmp.p, tmp_bytes, (const int *)ctx->r_len.p + ch.pair_lo,
                                              (long long *)ctx->r_rel.p + ch.pair_lo, cuda::std::plus<>{}, 0LL, np, sW));
            rle_base_kernel<<<1, 32, 0, sW>>>((long long *)ctx->r_base.p, (int)c, ch.pair_lo, np,
                                              (const long long *)ctx->r_rel.p, (const int *)ctx->r_len.p);
            rle_write_kernel<<<g1, 256, 0, sW>>>(ch.pair_lo, np, d_ops, (const long long *)d_ops_off, ctx->uni_ops_stride,
                                                 d_len, d_status, d_cost, (const long long *)ctx->r_rel.p,
                                                 (const long long *)ctx->r_base.p, (int)c, (long long)io->rle_cap,
                                                 (uint8_t *)ctx->r_rle.p, (long long *)ctx->r_off.p,
                                                 (PairResult *)ctx->r_res.p, (int *)ctx->r_flag.p,
                                                 c + 1 == nc ? 1 : 0, (long long)ctx->n_pairs);
            CK(cudaGetLastError());
            CK(cudaEventRecord(E.x1, sW));
        }
        if (io && io->rle) {
            CK(cudaStreamWaitEvent(sOut, E.x1, 0));
            CK(cudaEventRecord(E.out0, sOut));
            CK(cudaMemcpyAsync(ctx->h_base + c + 1, (long long *)ctx->r_base.p + c + 1, 8, cudaMemcpyDeviceToHost, sOut));
            CK(cudaEventRecord(E.base, sOut));
            CK(cudaMemcpyAsync(io->res + ch.pair_lo, (PairResult *)ctx->r_res.p + ch.pair_lo, (size_t)np * sizeof(PairResult),
                               cudaMemcpyDeviceToHost, sOut));
            const size_t extra = c + 1 == nc ? 1 : 0;          // the closing offset travels with the last chunk
            CK(cudaMemcpyAsync(io->rle_off + ch.pair_lo, (long long *)ctx->r_off.p + ch.pair_lo, ((size_t)np + extra) * 8,
                               cudaMemcpyDeviceToHost, sOut));
            CK(cudaEventRecord(E.out1, sOut));
        } else if (io) {
            CK(cudaStreamWaitEvent(sOut, E.w1, 0));
            CK(cudaEventRecord(E.out0, sOut));
            CK(cudaMemcpyAsync(io->cost + ch.pair_lo, d_cost + ch.pair_lo, (size_t)np * 4, cudaMemcpyDeviceToHost, sOut));
            if (!cost_only) {
                const int64_t o0 = io->ops_off ? io->ops_off[ch.pair_lo] : (int64_t)ch.pair_lo * ctx->uni_ops_stride;
                const int64_t o1 = io->ops_off ? io->ops_off[ch.pair_hi] : (int64_t)ch.pair_hi * ctx->uni_ops_stride;
                CK(cudaMemcpyAsync(io->ops + o0, d_ops + o0, (size_t)(o1 - o0), cudaMemcpyDeviceToHost, sOut));
                CK(cudaMemcpyAsync(io->len + ch.pair_lo, d_len + ch.pair_lo, (size_t)np * 4, cudaMemcpyDeviceToHost, sOut));
                CK(cudaMemcpyAsync(io->status + ch.pair_lo, d_status + ch.pair_lo, (size_t)np * 4, cudaMemcpyDeviceToHost, sOut));
            }
            CK(cudaEventRecord(E.out1, sOut));
        }
    }
    if (io && io->rle) {
        // The compact stream's size is known only on the device: as each chunk's running end
        // arrives (8 bytes), its part of the stream is copied out.  The host waits per chunk
        // while later chunks are still being filled and walked.
        cudaStream_t sOut = ctx->aux[5];
        ctx->h_base[0] = 0;
        for (size_t c = 0; c < nc; ++c) {
            ChunkEvents &E = ctx->cev[c];
            CK(cudaEventSynchronize(E.base));
            const long long lo = ctx->h_base[c], hi = std::min<long long>(ctx->h_base[c + 1], io->rle_cap);
            CK(cudaEventRecord(E.r0, sOut));
            if (hi > lo)
                CK(cudaMemcpyAsync(io->rle + lo, (const uint8_t *)ctx->r_rle.p + lo, (size_t)(hi - lo), cudaMemcpyDeviceToHost, sOut));
            CK(cudaEventRecord(E.r1, sOut));
        }
        CK(cudaStreamWaitEvent(main, ctx->cev[nc - 1].r1, 0));
    }
    // join: the caller-visible stream waits for the tails of every auxiliary stream
    if (!serial || io) {
        for (size_t c = (nc >= 2 ? nc - 2 : 0); c < nc; ++c) {
            CK(cudaStreamWaitEvent(main, ctx->cev[c].w1, 0));
            if (io) CK(cudaStreamWaitEvent(main, ctx->cev[c].out1, 0));
        }
    }
    CK(cudaEventRecord(ctx->ev[1], main));
    ctx->timing_valid = true;
    ctx->timing_io = io != nullptr;
    ctx->timing_rle = io && io->rle;
    return GA_OK;
}

extern "C" int ga_run_batch(ga_ctx *ctx, const uint8_t *d_s1, const int64_t *d_s1_off,
                            const uint8_t *d_s2, const int64_t *d_s2_off, int32_t *d_out_cost,
                            uint8_t *d_out_ops, const int64_t *d_out_ops_off, int32_t *d_out_ops_len,
                            int32_t *d_out_status) {
    if (!ctx) return GA_E_ARG;
    if (!ctx->have_plan) return fail(ctx, GA_E_STATE, "ga_plan_batch has not been called");
    const bool cost_only = (ctx->plan_flags & GA_F_COST_ONLY) != 0;
    if (!d_s1 || !d_s1_off || !d_s2 || !d_s2_off || !d_out_cost ||
        (!cost_only && (!d_out_ops || !d_out_ops_off || !d_out_ops_len || !d_out_status)))
        return fail(ctx, GA_E_ARG, "null device buffer");
    CK(cudaSetDevice(ctx->device));
    return run_pipeline(ctx, d_s1, d_s1_off, d_s2, d_s2_off, d_out_cost, d_out_ops, d_out_ops_off,
                        d_out_ops_len, d_out_status, nullptr);
}

extern "C" int ga_get_timing(ga_ctx *ctx, float *ms) {
    if (!ctx || !ms) return GA_E_ARG;
    if (!ctx->timing_valid || !ctx->have_plan) return fail(ctx, GA_E_STATE, "nothing has run yet");
    CK(cudaSetDevice(ctx->device));
    CK(cudaStreamSynchronize(ctx->stream));
    float fill = 0, walk = 0, h2d = 0, d2h = 0, total = 0;
    for (size_t c = 0; c < ctx->chunks.size(); ++c) {
        float t = 0;
        CK(cudaEventElapsedTime(&t, ctx->cev[c].f0, ctx->cev[c].f1));
        fill += t;
        CK(cudaEventElapsedTime(&t, ctx->cev[c].w0, ctx->cev[c].w1));
        walk += t;
        if (ctx->timing_io) {
            CK(cudaEventElapsedTime(&t, ctx->cev[c].out0, ctx->cev[c].out1));
            d2h += t;
            if (ctx->timing_rle) {
                CK(cudaEventElapsedTime(&t, ctx->cev[c].r0, ctx->cev[c].r1));
                d2h += t;
            }
        }
    }
    if (ctx->timing_io) CK(cudaEventElapsedTime(&h2d, ctx->in_begin, ctx->in_piece[ctx->in_pieces - 1]));
    if (getenv("GA_DEBUG_TIMELINE")) {      // per-chunk event times relative to the start of the call (ms)
        auto at = [&](cudaEvent_t e) { float t = 0; cudaEventElapsedTime(&t, ctx->ev[0], e); return t; };
        if (ctx->timing_io)
            fprintf(stderr, "[ga timeline] upload: first piece %.3f, last piece %.3f (%d pieces)\n",
                    at(ctx->in_piece[0]), at(ctx->in_piece[ctx->in_pieces - 1]), ctx->in_pieces);
        for (size_t c = 0; c < ctx->chunks.size(); ++c) {
            const ChunkEvents &E = ctx->cev[c];
            fprintf(stderr, "[ga timeline] chunk %zu pairs %d: fill %.3f-%.3f walk %.3f-%.3f", c,
                    ctx->chunks[c].pair_hi - ctx->chunks[c].pair_lo, at(E.f0), at(E.f1), at(E.w0), at(E.w1));
            if (ctx->timing_io) fprintf(stderr, " copy-out %.3f-%.3f", at(E.out0), at(E.out1));
            fprintf(stderr, "\n");
        }
        fprintf(stderr, "[ga timeline] end %.3f\n", at(ctx->ev[1]));
    }
    CK(cudaEventElapsedTime(&total, ctx->ev[0], ctx->ev[1]));
    // per-phase sums; with the pipeline on, phases overlap each other, so they add up
    // to more than `total` (the bracketing events on the caller-visible stream)
    ms[GA_T_H2D] = h2d;
    ms[GA_T_FILL] = fill;
    ms[GA_T_WALK] = walk;
    ms[GA_T_D2H] = d2h;
    ms[GA_T_TOTAL] = total;
    return GA_OK;
}

// ------------------------------------------------------------------ host-buffer entry points
// Shared by ga_align_batch (offset arrays) and ga_align_batch_uniform (every pair um x un, offset
// arrays implicit: s1_off == s2_off == out_ops_off == nullptr).
// compact output of ga_align_batch_rle (instead of cost / ops / ops_len / status)
struct RleOut {
    PairResult *res;
    uint8_t *rle;
    int64_t cap;
    int64_t *off;
};

static int align_batch_impl(ga_ctx *ctx, const uint8_t *s1, const int64_t *s1_off, const uint8_t *s2,
                            const int64_t *s2_off, int64_t um, int64_t un, int64_t n_pairs, int32_t *out_cost,
                            uint8_t *out_ops, const int64_t *out_ops_off, int32_t *out_ops_len,
                            int32_t *out_status, uint32_t flags, const RleOut *ro = nullptr) {
    const bool cost_only = (flags & GA_F_COST_ONLY) != 0;
    const bool implicit = s1_off == nullptr;
    if (ro && cost_only) return fail(ctx, GA_E_ARG, "the compact output has no cost-only form");
    if (n_pairs < 1 || n_pairs > 0x7fffff00LL) return fail(ctx, GA_E_ARG, "bad batch (1 <= n_pairs < 2^31)");
    CK(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream, sIn = ctx->aux[4];
    const int64_t t1 = implicit ? um * n_pairs : s1_off[n_pairs], t2 = implicit ? un * n_pairs : s2_off[n_pairs];
    if (t1 < n_pairs || t2 < n_pairs || (!implicit && (s1_off[0] != 0 || s2_off[0] != 0)))
        return fail(ctx, GA_E_ARG, "offset arrays must start at 0 and be increasing");
    const int bits = seq_bits_of(flags);
    if (bits < 0) return fail(ctx, GA_E_ARG, "choose one sequence packing");
    const size_t b1 = (size_t)((t1 * bits + 7) / 8), b2 = (size_t)((t2 * bits + 7) / 8);
    const size_t offb = (size_t)(n_pairs + 1) * 8;
    const int64_t ops_stride = ((((um + un) + 3) / 4) + 3) & ~(int64_t)3;
    // (compact output of a ragged batch: the internal slices are sized after planning)
    const size_t opsb = cost_only ? 0 : (size_t)(implicit ? ops_stride * n_pairs : (ro ? 0 : out_ops_off[n_pairs]));
    int rc;
    if ((rc = ensure(ctx, ctx->d_s1, b1 + 16)) || (rc = ensure(ctx, ctx->d_s2, b2 + 16)) ||
        (rc = ensure(ctx, ctx->d_cost, (size_t)n_pairs * 4)))
        return rc;
    if (!implicit && ((rc = ensure(ctx, ctx->d_off1, offb)) || (rc = ensure(ctx, ctx->d_off2, offb)))) return rc;
    if (!cost_only &&
        ((rc = ensure(ctx, ctx->d_ops, opsb + 16)) || (!implicit && (rc = ensure(ctx, ctx->d_opsoff, offb))) ||
         (rc = ensure(ctx, ctx->d_len, (size_t)n_pairs * 4)) ||
         (rc = ensure(ctx, ctx->d_status, (size_t)n_pairs * 4))))
        return rc;
    HostTimer ht0;
    // Start the uploads right away on the copy-in stream -- offset arrays, then the letters in
    // IN_PIECES pieces with an event each -- so that host-side planning overlaps with them and
    // the first fill can start as soon as its own piece has landed.
    CK(cudaEventRecord(ctx->ev[2], st));
    CK(cudaStreamWaitEvent(sIn, ctx->ev[2], 0));
    CK(cudaEventRecord(ctx->in_begin, sIn));
    if (!implicit) {
        CK(cudaMemcpyAsync(ctx->d_off1.p, s1_off, offb, cudaMemcpyHostToDevice, sIn));
        CK(cudaMemcpyAsync(ctx->d_off2.p, s2_off, offb, cudaMemcpyHostToDevice, sIn));
        if (ro) {
            // the internal 2-bit slices of a compact-output call: offsets made on the device
            size_t t_a = 0;
            CK(cub::DeviceScan::ExclusiveScan(nullptr, t_a, (long long *)nullptr, (long long *)nullptr, cuda::std::plus<>{}, 0LL, (int)n_pairs, sIn));
            if ((rc = ensure(ctx, ctx->r_tmp, t_a + 256))) return rc;
            ops_need_kernel<<<(int)((n_pairs + 255) / 256), 256, 0, sIn>>>((const long long *)ctx->d_off1.p, (const long long *)ctx->d_off2.p,
                                                                          (int)n_pairs, (long long *)ctx->d_opsoff.p);
            size_t tb = ctx->r_tmp.cap;
            CK(cub::DeviceScan::ExclusiveScan(ctx->r_tmp.p, tb, (long long *)ctx->d_opsoff.p, (long long *)ctx->d_opsoff.p,
                                              cuda::std::plus<>{}, 0LL, (int)n_pairs, sIn));
        } else if (!cost_only) {
            CK(cudaMemcpyAsync(ctx->d_opsoff.p, out_ops_off, offb, cudaMemcpyHostToDevice, sIn));
        }
    }
    // about 1 MB per piece: fine enough that the first chunk waits for little more than its own
    // letters, coarse enough that small batches are one copy per sequence set
    const int IN_PIECES = (int)std::max<size_t>(1, std::min<size_t>(MAX_IN_PIECES, (b1 + b2) >> 20));
    ctx->in_pieces = IN_PIECES;
    auto upload_piece = [&](int k) -> int {
        // piece k = letters [t*k/P, t*(k+1)/P) -> the bytes that hold them (edges may be sent twice)
        const int64_t a0 = (t1 * k / IN_PIECES) * bits / 8, a1 = ((t1 * (k + 1) / IN_PIECES) * bits + 7) / 8;
        const int64_t c0 = (t2 * k / IN_PIECES) * bits / 8, c1 = ((t2 * (k + 1) / IN_PIECES) * bits + 7) / 8;
        if (a1 > a0) CK(cudaMemcpyAsync((uint8_t *)ctx->d_s1.p + a0, s1 + a0, (size_t)(a1 - a0), cudaMemcpyHostToDevice, sIn));
        if (c1 > c0) CK(cudaMemcpyAsync((uint8_t *)ctx->d_s2.p + c0, s2 + c0, (size_t)(c1 - c0), cudaMemcpyHostToDevice, sIn));
        CK(cudaEventRecord(ctx->in_piece[k], sIn));
        return GA_OK;
    };
    // only the first pieces go out before planning: the copy engine works in issue order and the
    // (small) tile-descriptor upload of a non-uniform plan must not queue behind the whole batch
    const int EARLY = std::max(1, IN_PIECES / 4);
    for (int k = 0; k < EARLY; ++k)
        if ((rc = upload_piece(k))) return rc;
    ht0.lap("align: enqueue uploads");
    HostTimer ht;
    ctx->plan_for_io = true;
    ctx->plan_rle = ro != nullptr;
    rc = implicit ? plan_uniform(ctx, um, un, n_pairs, flags)
                  : ga_plan_batch(ctx, s1_off, s2_off, n_pairs, (cost_only || ro) ? nullptr : out_ops_off, flags);
    ctx->plan_for_io = false;
    ctx->plan_rle = false;
    if (rc) { cudaStreamSynchronize(sIn); return rc; }
    for (int k = EARLY; k < IN_PIECES; ++k)
        if ((rc = upload_piece(k))) return rc;
    ht.lap("align: plan total");
    HostIO io;
    io.total1 = t1; io.total2 = t2; io.pieces = IN_PIECES;
    io.s1 = s1; io.s2 = s2; io.off1 = s1_off; io.off2 = s2_off; io.ops_off = out_ops_off;
    io.cost = out_cost; io.len = out_ops_len; io.status = out_status; io.ops = out_ops;
    if (ro) {
        const size_t nc = ctx->chunks.size();
        size_t t_b = 0;
        int max_np = 0;
        for (const Chunk &ch : ctx->chunks) max_np = std::max(max_np, ch.pair_hi - ch.pair_lo);
        CK(cub::DeviceScan::ExclusiveScan(nullptr, t_b, (const int *)nullptr, (long long *)nullptr, cuda::std::plus<>{}, 0LL, max_np, st));
        if ((rc = ensure(ctx, ctx->d_ops, (size_t)ctx->plan_ops_bytes + 16)) || (rc = ensure(ctx, ctx->r_tmp, t_b + 256)) ||
            (rc = ensure(ctx, ctx->r_len, (size_t)n_pairs * 4)) || (rc = ensure(ctx, ctx->r_rel, (size_t)n_pairs * 8)) ||
            (rc = ensure(ctx, ctx->r_off, (size_t)(n_pairs + 1) * 8)) || (rc = ensure(ctx, ctx->r_res, (size_t)n_pairs * sizeof(PairResult))) ||
            (rc = ensure(ctx, ctx->r_base, (nc + 1) * 8)) || (rc = ensure(ctx, ctx->r_flag, 16)) ||
            (rc = ensure(ctx, ctx->r_rle, (size_t)ro->cap + 16)))
            return rc;
        if (ctx->h_base_cap < nc + 1) {
            if (ctx->h_base) CK(cudaFreeHost(ctx->h_base));
            ctx->h_base = nullptr;
            ctx->h_base_cap = 0;
            CK(cudaMallocHost(&ctx->h_base, (nc + 1 + 64) * 8));
            ctx->h_base_cap = nc + 1 + 64;
        }
        CK(cudaMemsetAsync(ctx->r_base.p, 0, 8, st));
        CK(cudaMemsetAsync(ctx->r_flag.p, 0, 16, st));
        io.res = ro->res; io.rle = ro->rle; io.rle_cap = ro->cap; io.rle_off = ro->off;
    }
    rc = run_pipeline(ctx, (const uint8_t *)ctx->d_s1.p, implicit ? nullptr : (const int64_t *)ctx->d_off1.p,
                      (const uint8_t *)ctx->d_s2.p, implicit ? nullptr : (const int64_t *)ctx->d_off2.p,
                      (int32_t *)ctx->d_cost.p, (uint8_t *)ctx->d_ops.p,
                      implicit ? nullptr : (const int64_t *)ctx->d_opsoff.p,
                      (int32_t *)ctx->d_len.p, (int32_t *)ctx->d_status.p, &io);
    if (rc) return rc;
    ht.lap("align: enqueue pipeline");
    CK(cudaStreamSynchronize(st));
    ht.lap("align: wait for the device");
    if (ro && ctx->h_base[ctx->chunks.size()] > ro->cap)
        return fail(ctx, GA_E_NOMEM, "the compact operation stream needs %lld bytes, the buffer holds %lld",
                    ctx->h_base[ctx->chunks.size()], (long long)ro->cap);
    return GA_OK;
}

extern "C" int ga_align_batch_rle(ga_ctx *ctx, const uint8_t *s1, const int64_t *s1_off, int64_t m, const uint8_t *s2,
                                  const int64_t *s2_off, int64_t n, int64_t n_pairs, ga_pair_result *out_res,
                                  uint8_t *out_rle, int64_t rle_capacity, int64_t *out_rle_off, uint32_t flags) {
    if (!ctx) return GA_E_ARG;
    static_assert(sizeof(ga_pair_result) == sizeof(PairResult), "record layout");
    if (!s1 || !s2 || !out_res || !out_rle || !out_rle_off || rle_capacity < 1 || (s1_off == nullptr) != (s2_off == nullptr))
        return fail(ctx, GA_E_ARG, "null host buffer");
    if (!s1_off && (m < 1 || n < 1)) return fail(ctx, GA_E_ARG, "fixed-length form needs m >= 1 and n >= 1");
    const RleOut ro{reinterpret_cast<PairResult *>(out_res), out_rle, rle_capacity, out_rle_off};
    return guarded(ctx, align_batch_impl(ctx, s1, s1_off, s2, s2_off, s1_off ? 0 : m, s1_off ? 0 : n, n_pairs, nullptr, nullptr,
                            nullptr, nullptr, nullptr, flags, &ro));
}

extern "C" int ga_align_batch(ga_ctx *ctx, const uint8_t *s1, const int64_t *s1_off,
                              const uint8_t *s2, const int64_t *s2_off, int64_t n_pairs,
                              int32_t *out_cost, uint8_t *out_ops, const int64_t *out_ops_off,
                              int32_t *out_ops_len, int32_t *out_status, uint32_t flags) {
    if (!ctx) return GA_E_ARG;
    const bool cost_only = (flags & GA_F_COST_ONLY) != 0;
    if (!s1 || !s2 || !s1_off || !s2_off || !out_cost ||
        (!cost_only && (!out_ops || !out_ops_off || !out_ops_len || !out_status)))
        return fail(ctx, GA_E_ARG, "null host buffer");
    return guarded(ctx, align_batch_impl(ctx, s1, s1_off, s2, s2_off, 0, 0, n_pairs, out_cost, out_ops, out_ops_off,
                            out_ops_len, out_status, flags));
}

extern "C" int ga_align_batch_uniform(ga_ctx *ctx, const uint8_t *s1, int64_t m, const uint8_t *s2, int64_t n,
                                      int64_t n_pairs, int32_t *out_cost, uint8_t *out_ops,
                                      int32_t *out_ops_len, int32_t *out_status, uint32_t flags) {
    if (!ctx) return GA_E_ARG;
    const bool cost_only = (flags & GA_F_COST_ONLY) != 0;
    if (!s1 || !s2 || !out_cost || m < 1 || n < 1 || (!cost_only && (!out_ops || !out_ops_len || !out_status)))
        return fail(ctx, GA_E_ARG, "null host buffer or empty sequences");
    return guarded(ctx, align_batch_impl(ctx, s1, nullptr, s2, nullptr, m, n, n_pairs, out_cost, out_ops, nullptr,
                            out_ops_len, out_status, flags));
}

// ------------------------------------------------------------------ one long pair
extern "C" int ga_align_long(ga_ctx *ctx, const uint8_t *s1, int64_t m, const uint8_t *s2, int64_t n,
                             int32_t *out_cost, uint8_t *out_ops, int64_t *out_n_ops,
                             int32_t *out_status, uint32_t flags, ga_long_info *info) {
    if (!ctx) return GA_E_ARG;
    if (!ctx->have_scheme) return fail(ctx, GA_E_STATE, "ga_set_scheme has not been called");
    const bool cost_only = (flags & GA_F_COST_ONLY) != 0;
    if (!s1 || !s2 || !out_cost || m < 1 || n < 1 || m > (1 << 28) || n > (1 << 28) ||
        (!cost_only && (!out_ops || !out_n_ops || !out_status)))
        return fail(ctx, GA_E_ARG, "bad arguments to ga_align_long");
    SchemeDev sc = ctx->sc;             // a copy: this entry point takes one letter per byte, whatever
    sc.seq_bits = 8;                    // packing the current batch plan (kept valid) was made for
    if (!sc.simple && sc.A > 64) return fail(ctx, GA_E_ARG, "table-driven schemes support at most 64 letters");
    {
        const long long big = (long long)(sc.max_cost + 1) * std::max(m, n);
        // plain values (traceback kernels) and the fill's gap-shifted values, which go down to
        // -4 * (sum of all gap-extension costs); both must stay clear of L_INVALID and below L_INF
        const long long vmax = big + (m + n + LBAND + 2) * (long long)sc.max_cost + 3LL * sc.o;
        if (4 * vmax + 3 >= 0x10000000LL) return fail(ctx, GA_E_RANGE, "costs too large for these lengths (4 * value must stay below 2^28)");
    }
    CK(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    // column-0 / row-0 running sums (reference make_dp_array, globaligner.py:802-819)
    std::vector<int> colsum((size_t)m + 1), rowsum((size_t)n + 1);
    colsum[0] = sc.o;
    for (int64_t i = 1; i <= m; ++i) {
        if (s1[i - 1] >= sc.A) return fail(ctx, GA_E_ARG, "letter index out of range in seq_1");
        colsum[(size_t)i] = colsum[(size_t)i - 1] + ctx->h_gv[s1[i - 1]];
    }
    rowsum[0] = sc.o;
    for (int64_t j = 1; j <= n; ++j) {
        if (s2[j - 1] >= sc.A) return fail(ctx, GA_E_ARG, "letter index out of range in seq_2");
        rowsum[(size_t)j] = rowsum[(size_t)j - 1] + ctx->h_gh[s2[j - 1]];
    }
    const int nbands = (int)((m + LBAND - 1) / LBAND);
    const int nck = (int)(n / LCC + 2);
    const size_t brow_b = (size_t)nbands * (size_t)(n + 1) * sizeof(int2);
    const size_t vchk_b = (size_t)nbands * nck * 32 * LR * sizeof(int2);
    if ((int64_t)(brow_b + vchk_b) > ctx->scratch_limit)
        return fail(ctx, GA_E_NOMEM, "checkpoints need %zu bytes, above the scratch limit", brow_b + vchk_b);
    const size_t ops_words = (size_t)((m + n + 15) / 16 + 1);
    int rc;
    // (seq_2 carries zeroed slack: the fill kernel's TMA tiles of the run-out steps reach up to
    // 2 * LTCOLS + 96 bytes behind the last letter; lanes whose column does not exist look letters
    // up there -- their results are discarded, but a table address is formed from them)
    if ((rc = ensure(ctx, ctx->d_s1, (size_t)m + 16)) || (rc = ensure(ctx, ctx->d_s2, (size_t)n + 2560)) ||
        (rc = ensure(ctx, ctx->l_colsum, colsum.size() * 4)) || (rc = ensure(ctx, ctx->l_rowsum, rowsum.size() * 4)) ||
        (rc = ensure(ctx, ctx->l_brow, brow_b)) || (rc = ensure(ctx, ctx->l_vchk, vchk_b)) ||
        (rc = ensure(ctx, ctx->l_misc, 256)) ||
        (rc = ensure(ctx, ctx->d_ops, ops_words * 4 + 16)))
        return rc;
    CK(cudaMemcpyAsync(ctx->d_s1.p, s1, (size_t)m, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(ctx->d_s2.p, s2, (size_t)n, cudaMemcpyHostToDevice, st));
    CK(cudaMemsetAsync((uint8_t *)ctx->d_s2.p + n, 0, 2560, st));
    CK(cudaMemcpyAsync(ctx->l_colsum.p, colsum.data(), colsum.size() * 4, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(ctx->l_rowsum.p, rowsum.data(), rowsum.size() * 4, cudaMemcpyHostToDevice, st));
    CK(cudaMemsetAsync(ctx->l_brow.p, 0x80, brow_b, st));      // L_INVALID = boundary entry not yet written
    CK(cudaMemsetAsync(ctx->l_misc.p, 0, 256, st));
    // l_misc: [0] band counter, [1] cost, [2] status, [3] tiles, [4..5] n_ops
    int *misc = (int *)ctx->l_misc.p;
    LongArgs a;
    a.s1 = (const uint8_t *)ctx->d_s1.p;
    a.s2 = (const uint8_t *)ctx->d_s2.p;
    a.m = (int)m;
    a.n = (int)n;
    a.colsum = (const int *)ctx->l_colsum.p;
    a.rowsum = (const int *)ctx->l_rowsum.p;
    a.brow = (int2 *)ctx->l_brow.p;
    a.vchk = (int2 *)ctx->l_vchk.p;
    a.band_counter = misc;
    a.nbands = nbands;
    a.nck = nck;
    a.big = (unsigned)((long long)(sc.max_cost + 1) * std::max(m, n));
    a.out_cost = misc + 1;
    a.o4 = a.o4b = 4 * sc.o;
    const size_t table = sc.simple ? 0 : ((((size_t)(sc.A + 1) * sc.A + sc.A) * 4 + 15) & ~(size_t)15);
    // one band-sweeping warp per scheduler: the chains of a lane give a lone warp enough
    // independent work to keep issuing; more bands than warps queue on the atomic band counter
    int blocks_per_sm = 1;
    if (const char *e = getenv("GA_LONG_BLOCKS_PER_SM")) blocks_per_sm = std::max(1, std::min(8, atoi(e)));
    const int grid = std::min((nbands + LWARPS - 1) / LWARPS, ctx->sm_count * blocks_per_sm);
    const bool fill_table = sc.A <= 64;          // shifted substitution table in shared memory
    CK(cudaEventRecord(ctx->ev[2], st));
    // small alphabets: one table entry per pair of rows (GA_LONG_PAIR_TABLE=0: the single-row table, for A/B runs)
    bool pair_table = LONG_PAIR_TABLE_OK && sc.A <= 8;
    if (const char *e = getenv("GA_LONG_PAIR_TABLE")) pair_table = pair_table && atoi(e) != 0;
    if (pair_table) long_fill_kernel<2><<<grid, 32 * LWARPS, (size_t)(sc.A + 1) * (sc.A + 1) * sc.A * 8, st>>>(sc, a);
    else if (fill_table) long_fill_kernel<1><<<grid, 32 * LWARPS, (size_t)(sc.A + 1) * sc.A * 4, st>>>(sc, a);
    else long_fill_kernel<0><<<grid, 32 * LWARPS, 0, st>>>(sc, a);
    CK(cudaGetLastError());
    CK(cudaEventRecord(ctx->ev[3], st));
    int launches = 1;
    // The band-parallel traceback is five dependent launches; for a pair of a few bands the single
    // warp that re-fills and walks tile after tile is quicker (GA_LONG_SEQ_BANDS: tuning override).
    // Measured (tools/prof_long.py with GA_LONG_SEQ_BANDS = 6 / 0), sequential vs band-parallel traceback:
    // 300 x 300 0.16 vs 0.20 ms, 500 x 500 0.24 vs 0.27 ms, 700 x 700 0.35 vs 0.30 ms, 1000 x 1000 0.50 vs 0.33 ms.
    int seq_bands = 3;
    if (const char *e = getenv("GA_LONG_SEQ_BANDS")) seq_bands = atoi(e);
    const bool parallel = !cost_only && m >= 2 && n >= 2 && !(flags & GA_F_SERIAL) && nbands > seq_bands;
    if (!cost_only) {
        TraceOut to;
        to.ops = (uint32_t *)ctx->d_ops.p;
        to.n_ops = (long long *)(misc + 4);
        to.status = misc + 2;
        to.tiles_refilled = misc + 3;
        const size_t tsmem = (size_t)LT_NTB * 32 * sizeof(uint4) + table;
        const int *skip = nullptr;
        if (parallel) {
            // parallel traceback: X-propagation over a diagonal corridor, chain, per-band detail, pack
            // The corridor is tried narrow first (similar sequences stay within a few hundred
            // columns of the straight diagonal; its width is the cost of the X-propagation), then
            // wide, on the device without a host round trip: the second pair of launches returns
            // at once when the first chain succeeded.
            ParTrace pt;
            const int W_try[2] = {192 + (int)((m + n) / 1024), 512 + (int)((m + n) / 128)};
            pt.run_if = nullptr;
            pt.W = W_try[1];
            pt.XW = 2 * pt.W + (int)(((int64_t)LBAND * n + m - 1) / m) + LCC + 8;
            pt.seg_cap = 2 * (m + n) + 4096 + (int64_t)nbands * 4;
            const size_t xrow_b = (size_t)nbands * pt.XW * 3 * 4;
            const size_t meta_b = (size_t)nbands * (16 + 8 + 4 + 8 + 8) + 64;
            if ((rc = ensure(ctx, ctx->l_xrow, xrow_b)) || (rc = ensure(ctx, ctx->l_meta, meta_b)) ||
                (rc = ensure(ctx, ctx->l_seg, (size_t)pt.seg_cap + 64)))
                return rc;
            uint8_t *meta = (uint8_t *)ctx->l_meta.p;
            pt.xrow = (int *)ctx->l_xrow.p;
            pt.seg_off = (long long *)meta;
            pt.seg_pos = (long long *)(meta + (size_t)nbands * 8);
            pt.entry = (int *)(meta + (size_t)nbands * 16);
            pt.seg_len = (int *)(meta + (size_t)nbands * 32);
            pt.corr = (int *)(meta + (size_t)nbands * 36);
            pt.seg_ops = (uint8_t *)ctx->l_seg.p;
            pt.xlast = misc + 7;
            pt.failed = misc + 6;
            pt.tiles = misc + 3;
            int n_try = 2;
            if (const char *e = getenv("GA_LONG_CORRIDORS")) n_try = atoi(e) == 1 ? 1 : 2;      // 1: the wide corridor only
            for (int k = 2 - n_try; k < 2; ++k) {
                ParTrace q = pt;
                q.W = W_try[k];
                q.XW = 2 * q.W + (int)(((int64_t)LBAND * n + m - 1) / m) + LCC + 8;
                q.run_if = (k == 1 && n_try == 2) ? pt.failed : nullptr;
                if (sc.simple) long_xprop_kernel<true><<<(nbands + LWARPS - 1) / LWARPS, 32 * LWARPS, 0, st>>>(sc, a, q);
                else long_xprop_kernel<false><<<(nbands + LWARPS - 1) / LWARPS, 32 * LWARPS, table, st>>>(sc, a, q);
                long_chain_kernel<<<1, 32, 0, st>>>(a, q);
                launches += 2;
            }
            launches -= 2;              // (the five counted below include one X-propagation + chain)
            if (sc.simple) {
                CK(cudaFuncSetAttribute(long_detail_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tsmem));
                long_detail_kernel<true><<<nbands, 32, tsmem, st>>>(sc, a, pt);
            } else {
                CK(cudaFuncSetAttribute(long_detail_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tsmem));
                long_detail_kernel<false><<<nbands, 32, tsmem, st>>>(sc, a, pt);
            }
            long_prefix_kernel<<<1, 32, 0, st>>>(a, pt, to.n_ops, to.status);
            long_pack_kernel<<<(int)((ops_words + 255) / 256), 256, 0, st>>>(a, pt, to.n_ops, to.ops);
            CK(cudaGetLastError());
            launches += 5;
            skip = pt.failed;           // the sequential form below runs only if the chain failed
        }
        if (sc.simple) {
            CK(cudaFuncSetAttribute(long_trace_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tsmem));
            long_trace_kernel<true><<<1, 32, tsmem, st>>>(sc, a, to, skip);
        } else {
            CK(cudaFuncSetAttribute(long_trace_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tsmem));
            long_trace_kernel<false><<<1, 32, tsmem, st>>>(sc, a, to, skip);
        }
        CK(cudaGetLastError());
        ++launches;
    }
    CK(cudaEventRecord(ctx->ev[4], st));
    int host_misc[16] = {};
    CK(cudaMemcpyAsync(host_misc, misc, sizeof host_misc, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    *out_cost = host_misc[1];
    if (!cost_only) {
        long long nops = 0;
        memcpy(&nops, host_misc + 4, 8);
        *out_n_ops = nops;
        *out_status = host_misc[2];
        CK(cudaMemcpy(out_ops, ctx->d_ops.p, (size_t)((nops + 15) / 16) * 4, cudaMemcpyDeviceToHost));
    }
    if (info) {
        memset(info, 0, sizeof *info);
        CK(cudaEventElapsedTime(&info->fill_ms, ctx->ev[2], ctx->ev[3]));
        CK(cudaEventElapsedTime(&info->trace_ms, ctx->ev[3], ctx->ev[4]));
        info->launches = launches;
        info->tile_rows = LBAND;
        info->tile_cols = LCC;
        info->retraced_tiles = host_misc[3];
        info->trace_mode = cost_only ? -1 : (!parallel ? 0 : (host_misc[6] ? 2 : 1));
        info->checkpoint_bytes = (int64_t)(brow_b + vchk_b);
        info->cells = m * n;
    }
    return guard_check(ctx);
}

// ------------------------------------------------------------------ full value table (kernels: ga_values.cuh)
// letters of a single pair must index the scheme's tables
static int check_letters(ga_ctx *ctx, const uint8_t *s, int64_t len, const char *which) {
    const int A = ctx->sc.A;
    for (int64_t k = 0; k < len; ++k)
        if (s[k] >= A) return fail(ctx, GA_E_ARG, "letter index out of range in %s", which);
    return GA_OK;
}

extern "C" int ga_fill_values(ga_ctx *ctx, const uint8_t *s1, int64_t m, const uint8_t *s2, int64_t n,
                              int keep_boundary, int32_t *table) {
    if (!ctx) return GA_E_ARG;
    if (!ctx->have_scheme) return fail(ctx, GA_E_STATE, "ga_set_scheme has not been called");
    if (!s1 || !s2 || !table || m < 1 || n < 1 || (m + 1) * (n + 1) > (int64_t)50'000'000)
        return fail(ctx, GA_E_ARG, "ga_fill_values is for small pairs");
    int rc;
    if ((rc = check_letters(ctx, s1, m, "seq_1")) || (rc = check_letters(ctx, s2, n, "seq_2"))) return rc;
    CK(cudaSetDevice(ctx->device));
    const size_t tb = (size_t)(m + 1) * (n + 1) * 3 * 4;
    if ((rc = ensure(ctx, ctx->v_s1, (size_t)m)) || (rc = ensure(ctx, ctx->v_s2, (size_t)n)) ||
        (rc = ensure(ctx, ctx->v_tab, tb)))
        return rc;
    uint8_t *d1 = (uint8_t *)ctx->v_s1.p, *d2 = (uint8_t *)ctx->v_s2.p;
    int *dt = (int *)ctx->v_tab.p;
    cudaStream_t st = ctx->stream;
    SchemeDev sc = ctx->sc;
    sc.seq_bits = 8;
    CK(cudaMemcpyAsync(d1, s1, m, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(d2, s2, n, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(dt, table, tb, cudaMemcpyHostToDevice, st));
    values_kernel<IntCosts><<<1, 256, 0, st>>>(IntCosts{sc}, d1, (int)m, d2, (int)n, keep_boundary, dt);
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(table, dt, tb, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    return GA_OK;
}

// ------------------------------------------------------------------ seeded replay of random ties (kernel: ga_values.cuh)
// Shared by ga_align_replay (table filled here from the sequences) and ga_walk_values (table given
// by the caller: the reference's dp_array_backward walks whatever dp_array it is handed).
static int replay_impl(ga_ctx *ctx, const uint8_t *s1, int64_t m, const uint8_t *s2, int64_t n,
                       const int32_t *given_table, const uint16_t *draws, int64_t n_draws, int32_t *out_cost,
                       uint8_t *out_ops, int64_t *out_n_ops, int32_t *out_n_eval, int32_t *out_status) {
    int rc;
    if ((rc = check_letters(ctx, s1, m, "seq_1")) || (rc = check_letters(ctx, s2, n, "seq_2"))) return rc;
    CK(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    const size_t tb = (size_t)(m + 1) * (n + 1) * 3 * 4;
    const size_t ops_words = (size_t)((m + n + 15) / 16 + 1);
    if ((rc = ensure(ctx, ctx->v_s1, (size_t)m)) || (rc = ensure(ctx, ctx->v_s2, (size_t)n)) ||
        (rc = ensure(ctx, ctx->v_tab, tb)) || (rc = ensure(ctx, ctx->v_draws, (size_t)std::max<int64_t>(n_draws, 1) * 2)) ||
        (rc = ensure(ctx, ctx->v_ops, ops_words * 4)) || (rc = ensure(ctx, ctx->v_misc, 64)))
        return rc;
    uint8_t *d1 = (uint8_t *)ctx->v_s1.p, *d2 = (uint8_t *)ctx->v_s2.p;
    int *dt = (int *)ctx->v_tab.p, *dmisc = (int *)ctx->v_misc.p;
    uint16_t *dd = draws ? (uint16_t *)ctx->v_draws.p : nullptr;
    uint32_t *dops = (uint32_t *)ctx->v_ops.p;
    SchemeDev sc = ctx->sc;
    sc.seq_bits = 8;
    CK(cudaMemcpyAsync(d1, s1, m, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(d2, s2, n, cudaMemcpyHostToDevice, st));
    if (draws) CK(cudaMemcpyAsync(dd, draws, (size_t)n_draws * 2, cudaMemcpyHostToDevice, st));
    if (given_table) {
        CK(cudaMemcpyAsync(dt, given_table, tb, cudaMemcpyHostToDevice, st));
    } else {
        values_kernel<IntCosts><<<1, 1024, 0, st>>>(IntCosts{sc}, d1, (int)m, d2, (int)n, 0, dt);
        CK(cudaGetLastError());
    }
    replay_walk_kernel<IntCosts><<<1, 32, 0, st>>>(IntCosts{sc}, d1, (int)m, d2, (int)n, dt, dd, dops,
                                         (long long *)(dmisc + 2), dmisc, dmisc + 1, dmisc + 4);
    CK(cudaGetLastError());
    int hm[8] = {};
    CK(cudaMemcpyAsync(hm, dmisc, sizeof hm, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    long long nops = 0;
    memcpy(&nops, hm + 2, 8);
    *out_n_eval = hm[0];
    *out_status = hm[1];
    *out_cost = hm[4];
    *out_n_ops = nops;
    CK(cudaMemcpy(out_ops, dops, (size_t)((nops + 15) / 16) * 4, cudaMemcpyDeviceToHost));
    return GA_OK;
}

extern "C" int ga_align_replay(ga_ctx *ctx, const uint8_t *s1, int64_t m, const uint8_t *s2, int64_t n,
                               const uint16_t *draws, int64_t n_draws, int32_t *out_cost, uint8_t *out_ops,
                               int64_t *out_n_ops, int32_t *out_n_eval, int32_t *out_status) {
    if (!ctx) return GA_E_ARG;
    if (!ctx->have_scheme) return fail(ctx, GA_E_STATE, "ga_set_scheme has not been called");
    if (!s1 || !s2 || !draws || !out_cost || !out_ops || !out_n_ops || !out_n_eval || !out_status || m < 1 ||
        n < 1 || n_draws < m + n || (m + 1) * (n + 1) > (int64_t)60'000'000)
        return fail(ctx, GA_E_ARG, "ga_align_replay: single pairs up to 6e7 table cells, m+n draws");
    {
        const long long big = (long long)(ctx->sc.max_cost + 1) * std::max(m, n);
        if (big + (m + n + 2) * (long long)ctx->sc.max_cost + 3LL * ctx->sc.o >= 0x7fffffffLL)
            return fail(ctx, GA_E_RANGE, "costs could overflow int32 for these lengths");
    }
    return guarded(ctx, replay_impl(ctx, s1, m, s2, n, nullptr, draws, n_draws, out_cost, out_ops, out_n_ops, out_n_eval,
                       out_status));
}

extern "C" int ga_walk_values(ga_ctx *ctx, const uint8_t *s1, int64_t m, const uint8_t *s2, int64_t n,
                              const int32_t *table, const uint16_t *draws, int64_t n_draws, int32_t *out_cost,
                              uint8_t *out_ops, int64_t *out_n_ops, int32_t *out_n_eval, int32_t *out_status) {
    if (!ctx) return GA_E_ARG;
    if (!ctx->have_scheme) return fail(ctx, GA_E_STATE, "ga_set_scheme has not been called");
    if (!s1 || !s2 || !table || !out_cost || !out_ops || !out_n_ops || !out_n_eval || !out_status || m < 1 ||
        n < 1 || (draws && n_draws < m + n) || (m + 1) * (n + 1) > (int64_t)60'000'000)
        return fail(ctx, GA_E_ARG, "ga_walk_values: single pairs up to 6e7 table cells, m+n draws or none");
    return guarded(ctx, replay_impl(ctx, s1, m, s2, n, table, draws, n_draws, out_cost, out_ops, out_n_ops, out_n_eval,
                       out_status));
}

// ------------------------------------------------------------------ float-valued penalties (f64)
// SURVEY.md section 8 row f2.  The scheme travels with the call (it is not the ctx's integer
// scheme).  One pair; value table in HBM (24 B per cell), value-ranking walk as above.
struct F64Pair {
    uint8_t *d1 = nullptr, *d2 = nullptr;
    double *dsch = nullptr, *dt = nullptr;
    uint16_t *dd = nullptr;
    uint32_t *dops = nullptr;
    int *dmisc = nullptr;
    ~F64Pair() { cudaFree(d1); cudaFree(d2); cudaFree(dsch); cudaFree(dt); cudaFree(dd); cudaFree(dops); cudaFree(dmisc); }
};

static int f64_setup(ga_ctx *ctx, F64Pair &w, F64Costs &c, int32_t A, const double *sub, const double *gap_h,
                     const double *gap_v, double gap_open, double max_cost, const uint8_t *s1, int64_t m,
                     const uint8_t *s2, int64_t n, cudaStream_t st) {
    for (int64_t k = 0; k < m; ++k) if (s1[k] >= A) return fail(ctx, GA_E_ARG, "seq_1 letter index outside the alphabet");
    for (int64_t k = 0; k < n; ++k) if (s2[k] >= A) return fail(ctx, GA_E_ARG, "seq_2 letter index outside the alphabet");
    const size_t nsch = (size_t)A * A + 2 * (size_t)A;
    std::vector<double> h(nsch);
    memcpy(h.data(), sub, sizeof(double) * A * A);
    memcpy(h.data() + (size_t)A * A, gap_h, sizeof(double) * A);
    memcpy(h.data() + (size_t)A * A + A, gap_v, sizeof(double) * A);
    for (double v : h) if (!(v == v) || v - v != 0.0) return fail(ctx, GA_E_ARG, "costs must be finite");
    if (!(gap_open == gap_open) || gap_open - gap_open != 0.0 || !(max_cost == max_cost) || max_cost - max_cost != 0.0)
        return fail(ctx, GA_E_ARG, "costs must be finite");
    CK(cudaMalloc(&w.d1, m));
    CK(cudaMalloc(&w.d2, n));
    CK(cudaMalloc(&w.dsch, nsch * sizeof(double)));
    CK(cudaMalloc(&w.dt, (size_t)(m + 1) * (n + 1) * 3 * sizeof(double)));
    CK(cudaMemcpyAsync(w.d1, s1, m, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(w.d2, s2, n, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(w.dsch, h.data(), nsch * sizeof(double), cudaMemcpyHostToDevice, st));
    CK(cudaStreamSynchronize(st));   // h goes out of scope with the caller's frame only, but keep it simple
    c = F64Costs{A, w.dsch, w.dsch + (size_t)A * A, w.dsch + (size_t)A * A + A, gap_open, max_cost};
    return GA_OK;
}

extern "C" int ga_fill_values_f64(ga_ctx *ctx, int32_t A, const double *sub, const double *gap_h,
                                  const double *gap_v, double gap_open, double max_cost, const uint8_t *s1,
                                  int64_t m, const uint8_t *s2, int64_t n, int keep_boundary, double *table) {
    if (!ctx) return GA_E_ARG;
    if (!sub || !gap_h || !gap_v || !s1 || !s2 || !table || A < 1 || A > 250 || m < 1 || n < 1 ||
        (m + 1) * (n + 1) > (int64_t)50'000'000)
        return fail(ctx, GA_E_ARG, "ga_fill_values_f64: single pairs up to 5e7 table cells");
    CK(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    F64Pair w;
    F64Costs c;
    int rc = f64_setup(ctx, w, c, A, sub, gap_h, gap_v, gap_open, max_cost, s1, m, s2, n, st);
    if (rc) return rc;
    const size_t tb = (size_t)(m + 1) * (n + 1) * 3 * sizeof(double);
    if (keep_boundary) CK(cudaMemcpyAsync(w.dt, table, tb, cudaMemcpyHostToDevice, st));
    values_kernel<F64Costs><<<1, 1024, 0, st>>>(c, w.d1, (int)m, w.d2, (int)n, keep_boundary, w.dt);
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(table, w.dt, tb, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    return GA_OK;
}

static int align_f64_impl(ga_ctx *ctx, int32_t A, const double *sub, const double *gap_h, const double *gap_v,
                          double gap_open, double max_cost, const uint8_t *s1, int64_t m, const uint8_t *s2,
                          int64_t n, const double *given_table, const uint16_t *draws, int64_t n_draws,
                          double *out_cost, uint8_t *out_ops, int64_t *out_n_ops, int32_t *out_n_eval,
                          int32_t *out_status) {
    CK(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    F64Pair w;
    F64Costs c;
    int rc = f64_setup(ctx, w, c, A, sub, gap_h, gap_v, gap_open, max_cost, s1, m, s2, n, st);
    if (rc) return rc;
    const size_t ops_words = (size_t)((m + n + 15) / 16 + 1);
    CK(cudaMalloc(&w.dops, ops_words * 4));
    CK(cudaMalloc(&w.dmisc, 64));
    if (draws) {
        CK(cudaMalloc(&w.dd, (size_t)n_draws * 2));
        CK(cudaMemcpyAsync(w.dd, draws, (size_t)n_draws * 2, cudaMemcpyHostToDevice, st));
    }
    if (given_table) {
        CK(cudaMemcpyAsync(w.dt, given_table, (size_t)(m + 1) * (n + 1) * 3 * sizeof(double), cudaMemcpyHostToDevice, st));
    } else {
        values_kernel<F64Costs><<<1, 1024, 0, st>>>(c, w.d1, (int)m, w.d2, (int)n, 0, w.dt);
        CK(cudaGetLastError());
    }
    // dmisc: [0] n_eval, [1] status, [2..3] n_ops (int64), [4..5] cost (f64)
    replay_walk_kernel<F64Costs><<<1, 32, 0, st>>>(c, w.d1, (int)m, w.d2, (int)n, w.dt, w.dd, w.dops,
                                                         (long long *)(w.dmisc + 2), w.dmisc, w.dmisc + 1,
                                                         (double *)(w.dmisc + 4));
    CK(cudaGetLastError());
    int hm[8] = {};
    CK(cudaMemcpyAsync(hm, w.dmisc, sizeof hm, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    long long nops = 0;
    memcpy(&nops, hm + 2, 8);
    memcpy(out_cost, hm + 4, 8);
    *out_n_eval = hm[0];
    *out_status = hm[1];
    *out_n_ops = nops;
    CK(cudaMemcpy(out_ops, w.dops, (size_t)((nops + 15) / 16) * 4, cudaMemcpyDeviceToHost));
    return GA_OK;
}

extern "C" int ga_align_f64(ga_ctx *ctx, int32_t A, const double *sub, const double *gap_h, const double *gap_v,
                            double gap_open, double max_cost, const uint8_t *s1, int64_t m, const uint8_t *s2,
                            int64_t n, const uint16_t *draws, int64_t n_draws, double *out_cost,
                            uint8_t *out_ops, int64_t *out_n_ops, int32_t *out_n_eval, int32_t *out_status) {
    if (!ctx) return GA_E_ARG;
    if (!sub || !gap_h || !gap_v || !s1 || !s2 || !out_cost || !out_ops || !out_n_ops || !out_n_eval ||
        !out_status || A < 1 || A > 250 || m < 1 || n < 1 || (draws && n_draws < m + n) ||
        (m + 1) * (n + 1) > (int64_t)50'000'000)
        return fail(ctx, GA_E_ARG, "ga_align_f64: single pairs up to 5e7 table cells, m+n draws or none");
    return align_f64_impl(ctx, A, sub, gap_h, gap_v, gap_open, max_cost, s1, m, s2, n, nullptr, draws, n_draws,
                          out_cost, out_ops, out_n_ops, out_n_eval, out_status);
}

extern "C" int ga_walk_values_f64(ga_ctx *ctx, int32_t A, const double *sub, const double *gap_h,
                                  const double *gap_v, double gap_open, double max_cost, const uint8_t *s1,
                                  int64_t m, const uint8_t *s2, int64_t n, const double *table,
                                  const uint16_t *draws, int64_t n_draws, double *out_cost, uint8_t *out_ops,
                                  int64_t *out_n_ops, int32_t *out_n_eval, int32_t *out_status) {
    if (!ctx) return GA_E_ARG;
    if (!sub || !gap_h || !gap_v || !s1 || !s2 || !table || !out_cost || !out_ops || !out_n_ops || !out_n_eval ||
        !out_status || A < 1 || A > 250 || m < 1 || n < 1 || (draws && n_draws < m + n) ||
        (m + 1) * (n + 1) > (int64_t)50'000'000)
        return fail(ctx, GA_E_ARG, "ga_walk_values_f64: single pairs up to 5e7 table cells, m+n draws or none");
    return align_f64_impl(ctx, A, sub, gap_h, gap_v, gap_open, max_cost, s1, m, s2, n, table, draws, n_draws,
                          out_cost, out_ops, out_n_ops, out_n_eval, out_status);
}

// ------------------------------------------------------------------ integer-ALU roofline probe (kernels: ga_probe.cuh)
extern "C" int ga_alu_peak(ga_ctx *ctx, int mode, double *gops) {
    if (!ctx || !gops || mode < 0 || mode > 12) return GA_E_ARG;
    CK(cudaSetDevice(ctx->device));
    const int blocks = ctx->sm_count * 8, threads = 256, iters = 4096;
    unsigned *d = nullptr;
    CK(cudaMalloc(&d, (size_t)blocks * threads * 4));
    cudaStream_t st = ctx->stream;
    auto launch = [&](int it) {
        switch (mode) {
#define GA_ALU_CASE(M) case M: alu_kernel<M><<<blocks, threads, 0, st>>>(d, it, 12345u); break;
            GA_ALU_CASE(0) GA_ALU_CASE(1) GA_ALU_CASE(2) GA_ALU_CASE(3) GA_ALU_CASE(4) GA_ALU_CASE(5)
            GA_ALU_CASE(6) GA_ALU_CASE(7) GA_ALU_CASE(8) GA_ALU_CASE(9) GA_ALU_CASE(10) GA_ALU_CASE(11)
            default: alu_kernel<12><<<blocks, threads, 0, st>>>(d, it, 12345u); break;
#undef GA_ALU_CASE
        }
    };
    launch(64);
    float best = 1e30f;
    for (int rep = 0; rep < 5; ++rep) {
        CK(cudaEventRecord(ctx->ev[6], st));
        launch(iters);
        CK(cudaEventRecord(ctx->ev[7], st));
        CK(cudaStreamSynchronize(st));
        float ms = 0;
        CK(cudaEventElapsedTime(&ms, ctx->ev[6], ctx->ev[7]));
        best = std::min(best, ms);
    }
    CK(cudaGetLastError());
    cudaFree(d);
    // instructions per thread per iteration
    double per_iter = 8.0 * 2.0;
    if (mode == 5) per_iter = 8.0 * 16.0;
    if (mode == 10) per_iter = 32.0;      // min3 + add per link, 16 links
    if (mode == 11 || mode == 12) per_iter = 16.0;
    *gops = per_iter * iters * (double)blocks * threads / (best * 1e-3) / 1e9;
    return GA_OK;
}
