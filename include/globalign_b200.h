/*
 * globalign_b200 -- C-ABI of the B200 (sm_100a) implementation of globalign's
 * Gotoh affine-gap Needleman-Wunsch hot path (three-level DP fill + traceback).
 *
 * The reference (iamgiddyaboutgit/globalign) is pure Python and has no FFI
 * layer; the seam this library replaces is the three plain calls inside
 * find_global_alignment (reference src/globalign/globaligner.py:256-293):
 *
 *     make_dp_array      globaligner.py:756-821   boundary row/column
 *     dp_array_forward   globaligner.py:366-392   fill (cell: :317-363)
 *     dp_array_backward  globaligner.py:395-593   traceback (tie table :595-685)
 *
 * A maintainer of the reference binds these entry points with ctypes (see
 * INTEGRATION.md for the stub).  All functions return 0 on success or a
 * negative GA_E_* code; ga_last_error() gives the message.  No function
 * throws, allocates memory that the caller must free, or takes torch types.
 *
 * Conventions
 *   - Sequences are "lowered": one uint8 letter index (0 .. A-1) per residue,
 *     pairs concatenated, CSR offsets (int64, n_pairs+1 entries, in letters).
 *     With GA_F_SEQ_5BIT / GA_F_SEQ_4BIT / GA_F_SEQ_2BIT (batch entry points) the letter buffers
 *     are little-endian bit streams instead: letter k of the concatenation occupies bits
 *     [k*b, (k+1)*b) of the byte array (b = 5, 4 or 2; a 5-bit letter may straddle two bytes);
 *     offsets stay in letters.
 *   - The scheme is a dense costing matrix (ga_set_scheme), exactly the
 *     reference's costing_mat with its "-" row/column split out:
 *         sub[a*A+b] = costing_mat[a][b]
 *         gap_h[b]   = costing_mat["-"][b]   horizontal step = gap in seq_1 (level 1)
 *         gap_v[a]   = costing_mat[a]["-"]   vertical step   = gap in seq_2 (level 2)
 *         gap_open   = gap_open_cost, max_cost = get_max_val(costing_mat)
 *   - ops: the alignment as 2-bit operations in WALK order (last alignment
 *     column first, as dp_array_backward produces them before its final
 *     reverse): 0 = match/mismatch column, 1 = gap in seq_1, 2 = gap in seq_2.
 *     Four ops per byte, op k of a pair in bits 2*(k%4) of byte k/4 of that
 *     pair's slice.  ops_off has n_pairs+1 entries (multiples of 4); pair p's slice is
 *     [ops_off[p], ops_off[p+1]) and must hold ceil((m_p+n_p)/4) bytes rounded up to a
 *     multiple of 4.
 *   - Ties are broken in the reference's canonical order (first option listed
 *     in cost_ranks_dispatcher: level 0, then 1, then 2).
 *   - status[p]: 0 ok; 1 = the reference's walk would leave the table for this
 *     pair (only possible when max(m,n) <= gap_open_cost; the reference raises
 *     IndexError or wraps around there) -- cost is still exact, ops undefined.
 */
#ifndef GLOBALIGN_B200_H
#define GLOBALIGN_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct ga_ctx ga_ctx;

enum {
    GA_OK = 0,
    GA_E_CUDA = -1,        /* a CUDA runtime call failed */
    GA_E_ARG = -2,         /* invalid argument */
    GA_E_RANGE = -3,       /* costs could overflow int32 for these lengths */
    GA_E_NOMEM = -4,       /* device scratch budget exceeded */
    GA_E_STATE = -5        /* call order (no scheme / no plan) */
};

/* flags for ga_align_batch / ga_plan_batch */
enum {
    GA_F_DEFAULT = 0,
    GA_F_FORCE_GENERIC = 1,   /* never use the packed 16-bit kernel */
    GA_F_COST_ONLY = 2,       /* skip the traceback walk (ops / ops_len untouched) */
    GA_F_SERIAL = 4,          /* one stream, no fill/walk/copy overlap (isolated kernel timing) */
    GA_F_SEQ_4BIT = 8,        /* s1/s2 are 4-bit packed letter streams (alphabet <= 16) */
    GA_F_SEQ_2BIT = 16,       /* s1/s2 are 2-bit packed letter streams (alphabet <= 4) */
    GA_F_SEQ_5BIT = 64,       /* s1/s2 are 5-bit packed letter streams (alphabet <= 32: amino acids) */
    GA_F_KEEP_ORDER = 32      /* ragged batches: form tiles in the caller's pair order instead of ordering
                                 each chunk's pairs by length first (results are the same either way) */
};

/* phases reported by ga_get_timing (milliseconds, CUDA events on the ctx stream) */
enum { GA_T_H2D = 0, GA_T_FILL = 1, GA_T_WALK = 2, GA_T_D2H = 3, GA_T_TOTAL = 4, GA_T_COUNT = 5 };

int ga_version(void);
int ga_device_count(int *count);

/* One context per (host thread, device).  Calls on a context are serialised by
 * the caller.  The context owns only scratch (direction codes, plan, staging). */
int ga_ctx_create(int device, ga_ctx **ctx);
int ga_ctx_destroy(ga_ctx *ctx);
const char *ga_last_error(const ga_ctx *ctx);

/* Launch on this cudaStream_t (e.g. torch.cuda.current_stream().cuda_stream);
 * NULL/0 = the context's own stream. */
int ga_ctx_set_stream(ga_ctx *ctx, void *cuda_stream);
/* Cap on library-owned direction-code scratch in bytes (default 64 GiB).  Batches are
 * processed in chunks whose codes ping-pong between two scratch buffers of at most
 * half this size each (one buffer of at most this size with GA_F_SERIAL). */
int ga_ctx_set_scratch_limit(ga_ctx *ctx, int64_t bytes);

/* Replaces: the costing_mat / gap_open_cost arguments of make_dp_array,
 * dp_array_forward, dp_array_backward (globaligner.py:260-287). */
int ga_set_scheme(ga_ctx *ctx, int A, const int32_t *sub, const int32_t *gap_h,
                  const int32_t *gap_v, int32_t gap_open, int32_t max_cost);

/* Whole hot path for a batch of pairs, host buffers in, host buffers out
 * (H2D and D2H inside).  Replaces, per pair, make_dp_array + dp_array_forward +
 * dp_array_backward.  out_ops may be NULL with GA_F_COST_ONLY. */
int ga_align_batch(ga_ctx *ctx, const uint8_t *s1, const int64_t *s1_off, const uint8_t *s2,
                   const int64_t *s2_off, int64_t n_pairs, int32_t *out_cost, uint8_t *out_ops,
                   const int64_t *out_ops_off, int32_t *out_ops_len, int32_t *out_status,
                   uint32_t flags);

/* The same for a fixed-length batch (every seq_1 has m letters, every seq_2 n letters -- the
 * usual shape of sequencing reads): no offset arrays; pair p's letters start at p*m / p*n and
 * its packed-ops slice at byte p * (ceil((m+n)/4) rounded up to a multiple of 4).  Saves the
 * host-side scan of the offset arrays and their upload. */
int ga_align_batch_uniform(ga_ctx *ctx, const uint8_t *s1, int64_t m, const uint8_t *s2, int64_t n,
                           int64_t n_pairs, int32_t *out_cost, uint8_t *out_ops, int32_t *out_ops_len,
                           int32_t *out_status, uint32_t flags);

/* The same hot path with a COMPACT result stream: one 8-byte record per pair and the alignments as
 * run-length encoded operations packed back to back -- what dp_array_backward's three strings
 * (globaligner.py:584-593) are made of, in ~10 bytes per read pair instead of a fixed slice of
 * ceil((m+n)/4) bytes.  With several GPUs behind one host the device-to-host copy of the result is
 * what limits the end-to-end rate; this form cuts it ~4x.
 *   out_res[p]      cost and number of alignment columns (bit 31 of nops_status = status[p])
 *   out_rle         RLE bytes, pair p at [out_rle_off[p], out_rle_off[p+1]): bits 0-1 operation
 *                   (0 / 1 / 2 as above), bits 2-7 run length - 1 (longer runs continue in the next
 *                   byte), walk order.  Capacity rle_capacity bytes; m+n bytes per pair always
 *                   suffice; GA_E_NOMEM (nothing written past the end) if it is too small.
 *   out_rle_off     n_pairs + 1 entries, written by the call.
 * s1_off == s2_off == NULL: fixed-length batch (every pair m x n, as ga_align_batch_uniform);
 * otherwise m and n are ignored.  flags: the GA_F_SEQ_* packings, GA_F_KEEP_ORDER, GA_F_FORCE_GENERIC. */
typedef struct {
    int32_t cost;
    uint32_t nops_status;
} ga_pair_result;
int ga_align_batch_rle(ga_ctx *ctx, const uint8_t *s1, const int64_t *s1_off, int64_t m, const uint8_t *s2,
                       const int64_t *s2_off, int64_t n, int64_t n_pairs, ga_pair_result *out_res,
                       uint8_t *out_rle, int64_t rle_capacity, int64_t *out_rle_off, uint32_t flags);

/* Same work on DEVICE-resident buffers, split into planning (host side: reads
 * the two offset arrays, sizes scratch, uploads tile metadata) and running
 * (only kernels; asynchronous on the ctx stream).  d_* are device pointers the
 * caller owns; offsets are passed both as host (for planning) and device
 * (for the kernels) copies. */
int ga_plan_batch(ga_ctx *ctx, const int64_t *s1_off, const int64_t *s2_off, int64_t n_pairs,
                  const int64_t *out_ops_off, uint32_t flags);
int ga_run_batch(ga_ctx *ctx, const uint8_t *d_s1, const int64_t *d_s1_off, const uint8_t *d_s2,
                 const int64_t *d_s2_off, int32_t *d_out_cost, uint8_t *d_out_ops,
                 const int64_t *d_out_ops_off, int32_t *d_out_ops_len, int32_t *d_out_status);
int ga_sync(ga_ctx *ctx);

/* Facts about the current plan: kernel name, number of kernel launches one
 * ga_run_batch makes, direction-code bytes written per run, cells per run. */
typedef struct {
    char fill_kernel[64];
    int32_t rows_per_lane, lanes_per_group, packed16;
    int32_t launches_per_run;
    int64_t code_bytes;
    int64_t cells;
    int64_t n_chunks;
} ga_plan_info;
int ga_get_plan_info(const ga_ctx *ctx, ga_plan_info *info);

/* CUDA-event milliseconds of the last ga_align_batch / ga_run_batch: per-phase sums
 * over the chunks (phases overlap each other unless GA_F_SERIAL was planned, so they
 * can add up to more than GA_T_TOTAL) and the bracketed total on the context stream. */
int ga_get_timing(ga_ctx *ctx, float *ms /* GA_T_COUNT entries */);

/* One long pair on the whole GPU, CHECKPOINTED traceback.  Replaces make_dp_array +
 * dp_array_forward + dp_array_backward (globaligner.py:756-821, :366-392, :395-593) for a
 * pair whose direction-code matrix should not be stored: the fill (bands of 192 rows
 * pipelined across all SMs) keeps only band-boundary rows and column checkpoints; the
 * traceback re-fills 192 x <=128 tiles under the path from those checkpoints and walks them
 * -- all bands in parallel once a cheap pass has located the path's crossing of every band
 * boundary (GA_F_SERIAL: tile after tile).  The result is the same canonical alignment as
 * ga_align_batch gives.  Host buffers; out_ops must hold 4 * ceil((m + n) / 16) + 4 bytes
 * (2-bit ops in walk order, as for ga_align_batch).
 *
 * Memory: NOT O(m + n).  The checkpoints take
 *     8 * (n + 1) * ceil(m / 192)  +  8 * 192 * ceil(m / 192) * (n / 128 + 2)   bytes
 * = about m * n / 9.6 bytes (0.104 bytes per cell: 1.04 GB for 100 kb x 100 kb, where stored
 * direction codes would take 5 GB; 104 GB for 1 Mb x 1 Mb).  They count against the scratch
 * limit (ga_ctx_set_scratch_limit, default 64 GiB: pairs up to about 780 kb x 780 kb; a B200
 * holds about 1.2 Mb x 1.2 Mb); beyond it GA_E_NOMEM is returned.  With GA_F_COST_ONLY the
 * same checkpoints are written (the fill kernel is the same).  A divide-and-conquer
 * (Hirschberg / Myers-Miller) form with O(m + n) memory is not implemented: reproducing the
 * reference's tie order exactly needs the forward values on the path, which is what the
 * checkpoints provide.
 * Values: 4 * (every cost on a path + the sentinel) must stay below 2^28 (GA_E_RANGE otherwise),
 * e.g. DNA with costs <= 5 up to about 4 M letters per sequence.
 * The reference refuses len(seq_1) * len(seq_2) >= 2e7 (start.py:213); this entry point
 * has no such guard. */
typedef struct {
    float fill_ms, trace_ms;          /* CUDA events on the context stream */
    int32_t launches;                 /* kernels launched */
    int32_t tile_rows, tile_cols;     /* checkpoint grid */
    int32_t retraced_tiles;           /* tiles the traceback re-filled */
    int64_t checkpoint_bytes;         /* device memory held by the checkpoints */
    int64_t cells;
    int32_t trace_mode;               /* 1 parallel across bands, 0 sequential, 2 parallel fell back, -1 none */
    int32_t reserved;
} ga_long_info;
int ga_align_long(ga_ctx *ctx, const uint8_t *s1, int64_t m, const uint8_t *s2, int64_t n,
                  int32_t *out_cost, uint8_t *out_ops, int64_t *out_n_ops, int32_t *out_status,
                  uint32_t flags, ga_long_info *info /* may be NULL */);

/* Debug / parity aid: the full (m+1) x (n+1) x 3 int32 value table of ONE pair,
 * i.e. what dp_array_forward leaves in dp_array (globaligner.py:366-392).  If
 * keep_boundary != 0, row 0 and column 0 of `table` (host, in/out) are used as
 * given instead of make_dp_array's values -- the reference's own
 * test_dp_array_forward passes hand-written boundary values. */
int ga_fill_values(ga_ctx *ctx, const uint8_t *s1, int64_t m, const uint8_t *s2, int64_t n,
                   int keep_boundary, int32_t *table);

/* Seeded replay of the reference's RANDOM tie-breaking for one pair.
 * cost_ranks_dispatcher (globaligner.py:595-685) draws 18 random.choice values per traceback
 * step; the caller emulates those draws (CPython: choice = seq[_randbelow(len)], rejection
 * sampling on getrandbits(2)) and passes per evaluated cell a uint16: bits 0-7 for
 * is_match = True, bits 8-15 for False; within the byte four 2-bit option indices for the tie
 * patterns {all three tied -> options (0,1,2); levels 0,1 tied -> (0,1); 0,2 tied -> (0,2);
 * 1,2 tied -> (1,2)}.  The device fills the full value table and one thread walks it ranking
 * values exactly as dp_array_backward does (:425-581).  out_n_eval = dispatcher calls made
 * (= draws consumed).  Not a throughput path: (m+1)*(n+1) <= 6e7. */
int ga_align_replay(ga_ctx *ctx, const uint8_t *s1, int64_t m, const uint8_t *s2, int64_t n,
                    const uint16_t *draws, int64_t n_draws, int32_t *out_cost, uint8_t *out_ops,
                    int64_t *out_n_ops, int32_t *out_n_eval, int32_t *out_status);

/* The traceback alone, on a value table the CALLER provides: the reference's dp_array_backward
 * (globaligner.py:395-593) walks whatever dp_array it is handed -- its own test builds one with
 * hand-written boundary values (tests/globaligner_test.py:6-37) -- ranking the three values of
 * each cell (:425-581).  `table` is (m+1) x (n+1) x 3 int32, row-major, as ga_fill_values leaves
 * it.  draws == NULL: canonical tie order; otherwise as ga_align_replay.  out_cost = min of the
 * corner cell (:425).  status 1: the walk leaves the table (the reference raises IndexError or
 * wraps to negative indices).  Not a throughput path: (m+1)*(n+1) <= 6e7. */
int ga_walk_values(ga_ctx *ctx, const uint8_t *s1, int64_t m, const uint8_t *s2, int64_t n,
                   const int32_t *table, const uint16_t *draws, int64_t n_draws, int32_t *out_cost,
                   uint8_t *out_ops, int64_t *out_n_ops, int32_t *out_n_eval, int32_t *out_status);

/* Float-valued penalties.  The reference's entry point coerces every score
 * and cost with int() (start.py:51-69,121-133,423), but its inner functions take floats when
 * called directly (type hints globaligner.py:371,761).  These two calls serve that case for
 * ONE pair in f64, adding and comparing in the reference's own order (min over the three
 * predecessors, then the step cost; :359-363), so the doubles are bit-identical to CPython's.
 * The scheme travels with the call: sub[a*A+b] = costing_mat[a][b], gap_h[b] = costing_mat["-"][b],
 * gap_v[a] = costing_mat[a]["-"], max_cost = get_max_val(costing_mat).
 *   ga_fill_values_f64: the table dp_array_forward leaves ((m+1)*(n+1)*3 doubles; keep_boundary as
 *     in ga_fill_values).
 *   ga_align_f64: cost + traceback ranking VALUES as dp_array_backward does (:425-581);
 *     draws == NULL -> canonical order (first-listed option of every tie), else as ga_align_replay.
 * Not a throughput path: (m+1)*(n+1) <= 5e7. */
int ga_fill_values_f64(ga_ctx *ctx, int32_t A, const double *sub, const double *gap_h, const double *gap_v,
                       double gap_open, double max_cost, const uint8_t *s1, int64_t m, const uint8_t *s2,
                       int64_t n, int keep_boundary, double *table);
int ga_align_f64(ga_ctx *ctx, int32_t A, const double *sub, const double *gap_h, const double *gap_v,
                 double gap_open, double max_cost, const uint8_t *s1, int64_t m, const uint8_t *s2, int64_t n,
                 const uint16_t *draws, int64_t n_draws, double *out_cost, uint8_t *out_ops,
                 int64_t *out_n_ops, int32_t *out_n_eval, int32_t *out_status);
/* ga_walk_values for an f64 table (the walk of ga_align_f64 on the caller's values). */
int ga_walk_values_f64(ga_ctx *ctx, int32_t A, const double *sub, const double *gap_h, const double *gap_v,
                       double gap_open, double max_cost, const uint8_t *s1, int64_t m, const uint8_t *s2,
                       int64_t n, const double *table, const uint16_t *draws, int64_t n_draws,
                       double *out_cost, uint8_t *out_ops, int64_t *out_n_ops, int32_t *out_n_eval,
                       int32_t *out_status);

/* Integer-ALU roofline microbenchmark (no memory traffic): giga-instructions/s the
 * device retires of one instruction kind.  Throughput modes (8 independent chains per
 * thread): 0 IADD, 1 VIMNMX3.S16x2, 2 VIMNMX3 (s32), 3 LOP3, 4 IMAD, 5 the packed fill
 * cell's own instruction mix, 6 VIADDMNMX.S16x2, 7 VIMNMX.U16x2, 8 VIADD.16x2,
 * 9 VIADDMNMX (s32).  Dependent-chain modes: 10 VIMNMX3.S16x2, 11 VIADDMNMX.S16x2, 12 IADD. */
int ga_alu_peak(ga_ctx *ctx, int mode, double *gops);

#ifdef __cplusplus
}
#endif
#endif /* GLOBALIGN_B200_H */
