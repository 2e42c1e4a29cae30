/*
 * seekr_b200.h -- C ABI of libseekr_b200.so, the B200 (sm_100a) implementation of SEEKR's hot path.
 *
 * The reference (CalabreseLab/seekr 2.0.2) is pure Python and has no FFI of its own; its
 * boundary for this path is the Python API.  Each entry point below names the reference
 * lines it replaces; the modules under seekr_b200/ bind them with ctypes and re-create that Python API
 * on top (INTEGRATION.md shows the stub a reference maintainer would add).
 *
 * Conventions
 *   - every function returns an int status (SKR_OK == 0); skr_last_error() gives the text of
 *     the last failure on the calling thread;
 *   - plain pointers and sizes only; "d_" pointers are device memory owned by the caller
 *     (the Python layer allocates them with torch), host pointers are plain host memory;
 *   - `stream` is a cudaStream_t passed as void* (NULL = default stream); all device entry
 *     points are asynchronous on it and never synchronise;
 *   - the current CUDA device must be the one that owns the buffers;
 *   - matrices are row-major; `ld` is the row pitch in elements.
 */
#ifndef SEEKR_B200_H
#define SEEKR_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SKR_ABI_VERSION 1

enum {
    SKR_OK = 0,
    SKR_ERR_ARG = 1,           /* bad argument (unsupported k, misaligned pointer, ...) */
    SKR_ERR_CUDA = 2,          /* a CUDA call failed; skr_last_error() has the CUDA message */
    SKR_ERR_IO = 3,            /* cannot open / read the FASTA file */
    SKR_ERR_NOMEM = 4,
    SKR_ERR_FASTA_BLANK = 5,   /* blank line: the reference raises IndexError (fasta_reader.py:53) */
    SKR_ERR_FASTA_HEADER = 6,  /* header without a sequence: AssertionError (fasta_reader.py:58) */
    SKR_ERR_CAPACITY = 7,      /* streamed packer: the record estimate was too small; restart on the finished handle */
};

const char* skr_last_error(void);
int skr_abi_version(void);

/* ------------------------------------------------------------------------------------------
 * Host ingest: FASTA text -> 2-bit codes + invalid mask   (replaces seekr/fasta_reader.py:41-78)
 *
 * Packed layout ("blocks" of 64 bases; every record starts on a block boundary):
 *   codes : 4 x uint32 per block, base p of a record in word p/16 at bits [30-2*(p%16), +2),
 *           i.e. earlier bases in higher bits, so a k-mer read as an integer has its first
 *           letter most significant like the reference's column order (kmer_counts.py:121-122)
 *   mask  : 2 x uint32 per block, base p in word p/32 at bit 31-(p%32); 1 = letter outside the
 *           alphabet, or padding past the end of the record (its code is 0)
 *   block_offsets[m+1] : first block of each record; lengths[m] : bases per record
 * One extra all-invalid block follows the last record so kernels may read one word ahead.
 * `lut[256]` maps a byte of the *upper-cased* sequence to its digit 0..3 or 255 (not in the
 * alphabet); the packer upper-cases ASCII letters itself, as Reader does (fasta_reader.py:55,62).
 * ------------------------------------------------------------------------------------------ */
typedef struct SkrPacked SkrPacked;

/* Parse FASTA text held in memory.  Line breaks are \n, \r\n or \r (text-mode open); lines are
 * stripped of leading/trailing ASCII whitespace; a line whose first character is '>' starts a
 * record.  nthreads <= 0 picks hardware_concurrency.  pinned != 0 allocates the packed arrays
 * with cudaHostAlloc (from a reusable pool) so they can be streamed to the GPU asynchronously. */
int skr_pack_fasta_buffer(const void* text, size_t nbytes, const uint8_t* lut, int nthreads, int pinned,
                          SkrPacked** out);
int skr_pack_fasta_file(const char* path, const uint8_t* lut, int nthreads, int pinned, SkrPacked** out);
/* Already-joined sequences (BasicCounter.seqs assigned by hand, occurrences(row, seq)):
 * letters = concatenated bytes, offs[m+1]. */
int skr_pack_sequences(const void* letters, const int64_t* offs, int64_t m, const uint8_t* lut, int nthreads,
                       int pinned, SkrPacked** out);
void skr_packed_free(SkrPacked* p);
/* skr_pack_fasta_buffer in two halves for the streamed path: the call returns once the text has been scanned (the
 * record table -- lengths, block offsets, header / body spans' offsets -- is final and every error of the
 * synchronous call has been reported), while `nthreads` background threads fill codes and mask in record order.
 * `text` must stay valid until skr_packed_wait returns.  skr_packed_wait_records blocks until records [0, upto)
 * are packed (upto < 0: all); skr_packed_wait also ends the background threads; skr_packed_free waits first. */
int skr_pack_fasta_buffer_async(const void* text, size_t nbytes, const uint8_t* lut, int nthreads, int pinned,
                                SkrPacked** out);
/* Texts of 32 MB and more (SEEKR_B200_WAVE_MIN_BYTES) are scanned in the background too, wave by wave: the call
 * returns after the first wave with a slab sized for an estimated record count (skr_packed_capacity_records); only
 * the errors of that first part are reported by the call itself, later ones by skr_packed_wait / _wait_records /
 * _wait_scanned (num_records and friends return -1 then).  skr_packed_wait_scanned blocks until records [0, want)
 * are in the record table (want < 0: until the table is complete), *avail = records in the table so far, *finished
 * = 1 once the count is final; SKR_ERR_CAPACITY when the estimate proved too small (the job then rebuilds an exact
 * slab -- codes / mask / offsets pointers CHANGE -- and the handle is complete after skr_packed_wait).
 * skr_packed_num_records / _num_blocks / _total_bases / _header_spans / _body_spans wait for the table. */
int skr_packed_wait_records(SkrPacked* p, int64_t upto);
int skr_packed_wait(SkrPacked* p);
int skr_packed_wait_scanned(SkrPacked* p, int64_t want, int64_t* avail, int* finished);
int64_t skr_packed_capacity_records(const SkrPacked* p);

int64_t skr_packed_num_records(const SkrPacked* p);
int64_t skr_packed_num_blocks(const SkrPacked* p);  /* including the trailing pad block */
int64_t skr_packed_total_bases(const SkrPacked* p);
const uint32_t* skr_packed_codes(const SkrPacked* p);          /* 4 * num_blocks words */
const uint32_t* skr_packed_mask(const SkrPacked* p);           /* 2 * num_blocks words */
const uint64_t* skr_packed_block_offsets(const SkrPacked* p);  /* m + 1 */
const uint32_t* skr_packed_lengths(const SkrPacked* p);        /* m */
/* byte spans into the source text, 2 per record: (offset, length) of the stripped header line
 * (Reader.get_headers, fasta_reader.py:75-78) and of the record body (first to last sequence line) */
const uint64_t* skr_packed_header_spans(const SkrPacked* p);
const uint64_t* skr_packed_body_spans(const SkrPacked* p);
/* the four arrays live in one contiguous host slab (codes first), so one copy moves them all */
const void* skr_packed_slab(const SkrPacked* p);
size_t skr_packed_slab_bytes(const SkrPacked* p);
/* 1-based line number of the offending line after SKR_ERR_FASTA_BLANK / _HEADER (0 otherwise) */
int64_t skr_pack_error_line(void);

/* ------------------------------------------------------------------------------------------
 * k-mer counting with fused normalisation   (replaces seekr/kmer_counts.py:140-151, 189-209)
 * ------------------------------------------------------------------------------------------ */

/* Device cell for the matrix-wide minimum Log2.post needs (kmer_counts.py:208). */
typedef struct {
    uint32_t min_ordered; /* order-preserving encoding of the smallest non-NaN value seen */
    uint32_t nan_seen;    /* non-zero once any NaN was seen: np.min propagates NaN */
} SkrMinCell;

int skr_min_reset(SkrMinCell* d_cell, void* stream);

/* One row per record: overlapping k-mer histogram, scaled to counts per kb exactly as the
 * reference does (c-fold binary64 sum of 1000/(L-k+1), rounded once), then optionally
 * log2(x+1) [log2_pre], -mean, /std.  d_mean/d_std may be NULL (step skipped); vec_is_f64
 * says whether they hold doubles (numpy then computes in binary64 and rounds) or floats.
 * out_is_f64 != 0 writes doubles (raw counts only: occurrences() on a float64 row).
 * d_min, when given, receives the running minimum / NaN flag of everything written.
 * d_post, when given, holds the matrix-wide minimum already (see skr_count_colmin) and the Log2.post
 * tail (+ |min|, + 1, log2; kmer_counts.py:207-209) is applied in the same epilogue.
 * d_rstd, when given (fp32 vectors only), holds RN(1/std) from skr_reciprocal and switches the division to
 * a 5-instruction correctly rounded form; the caller guarantees 2^-40 <= std <= 2^40 and |mean| <= 2^40.
 * Records with L < k give a zero-count row; the caller must reject L == k-1 beforehand
 * (ZeroDivisionError in the reference).  1 <= k <= 8. */
int skr_count(const uint32_t* d_codes, const uint32_t* d_mask, const uint64_t* d_block_offsets,
              const uint32_t* d_lengths, int64_t m, int k, int log2_pre, const void* d_mean, const void* d_std,
              int vec_is_f64, void* d_out, int out_is_f64, int64_t ld_out, SkrMinCell* d_min, const SkrMinCell* d_post,
              const float* d_rstd, void* stream);

/* Speculative Log2.post (seekr_kmer_counts -mv -sv, kmer_counts.py:207-209 in ONE pass over the matrix).
 * With supplied vectors (finite mean, finite std > 0) the z-score fl(fl(x - mean_j) / std_j) is monotone in the
 * count value x >= 0, so the matrix-wide minimum is min_j fl(fl(0 - mean_j) / std_j) whenever the arg-min column j*
 * holds a zero count in some record -- which any realistic input does.  skr_post_spec derives that shift and j*
 * from the vectors alone (identical on every rank of a sharded run: no collective before counting);
 * skr_count_ex applies the Log2.post tail with it in its epilogue and writes the run's epoch (spec_epoch, a
 * non-zero number the caller changes from run to run, so the cell -- shift and j* depend on the vectors only --
 * is set up once and never reset) into zero_seen when it meets a zero count in column j*.  The fallback launches
 * (count with a running minimum, skr_post_log2_skip) are enqueued behind it with d_skip = &zero_seen and
 * skip_value = that epoch: they return at once when the speculation held, and redo the matrix the two-pass way
 * when it did not -- the choice is made on the device, the host never waits. */
typedef struct {
    SkrMinCell shift;   /* the speculated matrix minimum, in the form the Log2.post tail reads (d_post) */
    int32_t zero_col;   /* j*; -1 when there is nothing to speculate on */
    uint32_t zero_seen; /* epoch of the last run in which a record with a zero count in column j* was seen */
} SkrPostSpec;
int skr_post_spec(const void* d_mean, const void* d_std, int vec_is_f64, int64_t cols, SkrPostSpec* d_spec, void* stream);
/* ... and, when both vectors are given, the tail ((x - mean_j)/std_j + shift) + 1 of kmer_counts.py:169,175,208 as one
 * multiply-add x * a_j + b_j: d_post_a[j] = RN(1/std_j); d_post_b[j] = the tail of a ZERO count in column j, evaluated
 * with the reference's own fp32 operations, so empty bins keep the reference's bits and the matrix minimum is exactly
 * log2(1) (16-byte aligned arrays of `cols` floats).  Every term is >= 0, nothing cancels: counted bins differ from the
 * step-by-step tail by a few ulp before the log2.  skr_count_ex takes the arrays as d_post_a / d_post_b next to d_spec. */
int skr_post_spec_affine(const void* d_mean, const void* d_std, int vec_is_f64, int64_t cols, SkrPostSpec* d_spec,
                         float* d_post_a, float* d_post_b, void* stream);

/* skr_count with every optional piece in one argument block (zero-initialise, then fill what is needed).
 * Beyond skr_count's arguments:
 *   d_colmin   per-column minima of the un-normalised values (what skr_count_colmin returns; no vectors then);
 *              d_out may be NULL for a minima-only pass
 *   d_colsum / d_colsq   accurate column statistics in the same pass (k = 4, 5, 6, plain counts): every thread sums
 *              the values and squares of its own columns over its records in fp32 and adds them here (binary64
 *              atomics, arrays zeroed by the caller) when it is done; skr_colstat_finish turns them into the
 *              fp32 mean / std vectors (binary64: mean = S1/rows, var = S2/rows - mean^2).  Closer to the exact
 *              value than numpy's sequential fp32 sums, hence not bit-identical to the reference
 *              (skr_col_pass is the order-exact route); one exchange of 2 * 4^k doubles when sharded
 *   d_spec     speculative Log2.post (above): d_post must point at d_spec->shift; spec_epoch (0 = 1) is what the
 *              kernel writes into d_spec->zero_seen
 *   d_skip     the launch returns at once when *d_skip == skip_value (0 = 1)
 *   d_min_reset  a minimum cell reset by the launch's first thread (the cell a launch BEHIND this one tracks:
 *              saves the separate skr_min_reset launch of the fallback route) */
typedef struct {
    const uint32_t* d_codes;
    const uint32_t* d_mask;
    const uint64_t* d_block_offsets;
    const uint32_t* d_lengths;
    int64_t m;
    int32_t k;
    int32_t log2_pre;
    const void* d_mean;
    const void* d_std;
    const float* d_rstd;
    int32_t vec_is_f64;
    int32_t out_is_f64;
    void* d_out;
    int64_t ld_out;
    SkrMinCell* d_min;
    const SkrMinCell* d_post;
    uint32_t* d_colmin;
    double* d_colsum;
    double* d_colsq;
    SkrPostSpec* d_spec;
    const uint32_t* d_skip;
    uint32_t max_length; /* longest record of this launch if the caller knows it (0 = unknown): lets k <= 6 skip the
                            launch that drains the list of records too long for the batch / team kernels */
    uint32_t skip_value;
    uint32_t spec_epoch;
    uint32_t reserved;
    SkrMinCell* d_min_reset;
    const float* d_post_a; /* with d_spec: the Log2.post tail folded into log2(x * a_j + b_j) (skr_post_spec_affine) */
    const float* d_post_b;
} SkrCountArgs;
int skr_count_ex(const SkrCountArgs* args, void* stream);
int skr_colstat_finish(const double* d_colsum, const double* d_colsq, int64_t cols, int64_t total_rows, float* d_mean,
                       float* d_std, int* d_flags /* [2]: mean, std; bit 0 not finite, bit 1 not positive */, void* stream);
/* skr_post_log2 that returns at once when *d_skip == skip_value (0 = 1) */
int skr_post_log2_skip(float* d_a, int64_t m, int64_t cols, int64_t ld, const SkrMinCell* d_min, const uint32_t* d_skip,
                       uint32_t skip_value, void* stream);

/* Streamed get_counts() (replaces the read-everything / count-everything / copy-everything sequence of
 * fasta_reader.py:41-63 + kmer_counts.py:196-200 when rows are independent once the vectors are known):
 * follows a packer started with skr_pack_fasta_buffer_async chunk by chunk -- pack chunk i+2 || H2D chunk i+1 ||
 * count chunk i || D2H chunk i-1 -- on two internal copy streams and the caller's `stream`.
 *   count        template for the per-chunk skr_count_ex calls: k, log2_pre, vectors, d_out (the WHOLE m x ld_out
 *                device matrix, rows are filled chunk by chunk), d_post / d_spec (speculative Log2.post), ...;
 *                the four packed-input pointers and m are filled in per chunk
 *   d_slab       device buffer of skr_packed_slab_bytes bytes: the packed arrays arrive here, same layout as the
 *                host slab (afterwards it is a complete device copy, usable with skr_count)
 *   h_out/h_ld   host destination (pitch in elements), or NULL to leave the result on the device;
 *                h_out_pinned != 0: page-locked memory, written directly by the copy engine; 0: pageable memory,
 *                reached through a ring of four small pinned slots drained by copy_threads host threads
 *   chunk_records  0 = about 32 MB of output rows per chunk
 * Returns when every row has reached the host (the copy streams are synchronised; `stream` is not). */
typedef struct {
    SkrCountArgs count;
    void* d_slab;
    float* h_out;
    int64_t h_ld;
    int32_t h_out_pinned;
    int32_t copy_threads;
    int64_t chunk_records;
    int64_t capacity_records; /* rows count.d_out / h_out can take; 0 = the handle's record count (which waits for a
                               * background scan).  A handle that turns out larger ends the call with SKR_ERR_CAPACITY */
    int64_t records_done;     /* out: records counted and copied out when the call returned */
} SkrStreamArgs;
int skr_stream_counts(SkrPacked* packed, SkrStreamArgs* args, void* stream);

/* Deferred normalisation for Log2.post with known, finite mean / positive std vectors (the
 * seekr_kmer_counts -mv -sv path): the count kernel writes the un-normalised values and keeps the
 * per-column minimum (float bits, +inf initialised by skr_colmin_reset; completed by skr_colmin_scan); because rounded subtraction and
 * division by a positive number are monotone, min_ij fl(fl(x_ij - mean_j)/std_j) = min_j fl(fl(xmin_j -
 * mean_j)/std_j), which skr_colmin_finish puts into the min cell; skr_normalize_post_log2 then applies
 * -mean, /std, +|min|, +1, log2 in one element-wise pass (same roundings as kmer_counts.py:169,175,207-209).
 * Sharded runs all-reduce(min) the column array before finishing.  skr_vec_check sets bit 0 of *d_flag
 * if a vector element is not finite and bit 1 if one is <= 0 (callers fall back to the fused path then). */
int skr_colmin_reset(uint32_t* d_colmin, int64_t cols, void* stream);
int skr_count_colmin(const uint32_t* d_codes, const uint32_t* d_mask, const uint64_t* d_block_offsets,
                     const uint32_t* d_lengths, int64_t m, int k, int log2_pre, float* d_out, int64_t ld_out,
                     uint32_t* d_colmin, void* stream);
/* columns in which no record had a zero count still hold +inf after skr_count_colmin: reduce them over the rows */
int skr_colmin_scan(const float* d_a, int64_t m, int64_t cols, int64_t ld, uint32_t* d_colmin, void* stream);
int skr_colmin_finish(const uint32_t* d_colmin, int64_t cols, const void* d_mean, const void* d_std, int vec_is_f64,
                      SkrMinCell* d_min, void* stream);
int skr_normalize_post_log2(float* d_a, int64_t m, int64_t cols, int64_t ld, const void* d_mean, const void* d_std,
                            int vec_is_f64, const SkrMinCell* d_min, void* stream);
int skr_vec_check(const void* d_vec, int vec_is_f64, int64_t n, int* d_flag, void* stream);
/* out[j] = RN(1 / v[j]) */
int skr_reciprocal(const float* d_vec, int64_t n, float* d_out, void* stream);
/* counts operand pairs (of n generated on the device) for which the reciprocal-based division differs from
 * IEEE division; *d_mismatches must be zeroed by the caller */
int skr_selftest_division(uint64_t n, uint64_t seed, uint64_t* d_mismatches, void* stream);

/* a = log2(a + 1)                                   (BasicCounter.log2_norm, kmer_counts.py:189-192) */
int skr_log2_norm(float* d_a, int64_t m, int64_t cols, int64_t ld, void* stream);
/* a = log2((a + |min|) + 1), min read from the device cell; NaN min -> all NaN   (kmer_counts.py:207-209) */
int skr_post_log2(float* d_a, int64_t m, int64_t cols, int64_t ld, const SkrMinCell* d_min, void* stream);
/* a -= vec (center with a supplied vector, kmer_counts.py:169);  a /= vec (standardize, :175).
 * Both update d_min (optional) with what they write. */
int skr_sub_vec(float* d_a, int64_t m, int64_t cols, int64_t ld, const void* d_vec, int vec_is_f64,
                SkrMinCell* d_min, void* stream);
int skr_div_vec(float* d_a, int64_t m, int64_t cols, int64_t ld, const void* d_vec, int vec_is_f64,
                SkrMinCell* d_min, void* stream);
/* a = fl(fl(a - mean) / std) in one pass, either vector optional (both share vec_is_f64) */
int skr_normalize(float* d_a, int64_t m, int64_t cols, int64_t ld, const void* d_mean, const void* d_std,
                  int vec_is_f64, SkrMinCell* d_min, void* stream);
/* running minimum / NaN flag of an existing matrix */
int skr_min_scan(const float* d_a, int64_t m, int64_t cols, int64_t ld, SkrMinCell* d_min, void* stream);

/* ------------------------------------------------------------------------------------------
 * Column statistics for mean=True / std=True and seekr_norm_vectors
 * (replaces np.mean / np.std(axis=0) at seekr/kmer_counts.py:168,174)
 *
 * Order-exact passes: numpy reduces axis 0 row after row in fp32, so each column is summed
 * sequentially in row order with plain fp32 adds; d_acc[cols] carries the running sums in and
 * out, which lets row shards on several GPUs be chained (rank r continues from rank r-1).
 * ------------------------------------------------------------------------------------------ */
enum {
    SKR_COLPASS_SUM = 0,      /* acc += a[i][j] */
    SKR_COLPASS_CENTERED = 1, /* acc += fl(a[i][j] - vec[j])                     (sum of the centred matrix) */
    SKR_COLPASS_SQDEV = 2,    /* acc += fl(y - vec2[j])^2, y = vec ? fl(a[i][j] - vec[j]) : a[i][j] */
};
/* None of the passes writes the matrix: the centred values are recomputed where they are needed
 * (each with the same single fp32 rounding the reference's in-place `counts -= mean` applies), and
 * skr_normalize writes the final matrix once.  d_vec: mean vector (fp32 or fp64), d_vec2: fp32. */
int skr_col_pass(int kind, const float* d_a, int64_t m, int64_t cols, int64_t ld, const void* d_vec, int vec_is_f64,
                 const float* d_vec2, float* d_acc, void* stream);
/* mean[j] = (float)((double)acc[j] / rows);  std[j] = sqrtf((float)((double)acc[j] / rows)).
 * d_flag (optional, one int): bit 0 set if a result is not finite, bit 1 if a result is <= 0. */
int skr_col_finish(const float* d_acc, int64_t cols, int64_t total_rows, int take_sqrt, float* d_out, int* d_flag,
                   void* stream);

/* Scalable passes for sharded runs: per-column binary64 partial sums computed row-parallel
 * (to be summed across ranks with one all-reduce), kinds as above.  More accurate than the
 * reference, not bit-identical to it. */
int skr_col_partial_f64(int kind, const float* d_a, int64_t m, int64_t cols, int64_t ld, const void* d_vec,
                        int vec_is_f64, const float* d_vec2, double* d_acc, void* stream);
int skr_col_finish_f64(const double* d_acc, int64_t cols, int64_t total_rows, int take_sqrt, float* d_out,
                       int* d_flag, void* stream);

/* ------------------------------------------------------------------------------------------
 * Pearson   (replaces seekr/pearson.py:32-44)
 *
 * skr_pearson_prepare row-standardises (mean, std ddof=0; pearson.py:35-38) and splits every
 * value into two fp16 planes hi + lo (22 significant bits) after scaling the row by a power of
 * two; skr_pearson_gemm forms hi*hi' + hi*lo' + lo*hi' with tcgen05 MMAs (fp32 accumulation in
 * TMEM), un-scales and multiplies by alpha (= 1/K, pearson.py:41).  symmetric != 0 (same operands,
 * m == n, whole matrix in one call) computes only the tiles on and above the diagonal and mirrors them.
 * Planes are [rows_padded][k_padded] fp16, rows_padded % 128 == 0, k_padded % 64 == 0, zero filled.
 * ------------------------------------------------------------------------------------------ */
int64_t skr_pearson_rows_padded(int64_t rows);
int64_t skr_pearson_k_padded(int64_t K);
int skr_pearson_prepare(const void* d_a, int a_is_f64, int64_t rows, int64_t K, int64_t ld, int row_standardize,
                        uint16_t* d_hi, uint16_t* d_lo, float* d_row_scale, void* stream);
int skr_pearson_gemm(const uint16_t* d_a_hi, const uint16_t* d_a_lo, const float* d_a_scale, int64_t m,
                     const uint16_t* d_b_hi, const uint16_t* d_b_lo, const float* d_b_scale, int64_t n, int64_t K,
                     double alpha, void* d_c, int c_is_f64, int64_t ldc, int symmetric, void* stream);

/* ------------------------------------------------------------------------------------------
 * Consumers of the r matrix (the reference runs them as Python loops)
 *
 * skr_pval_empirical   replaces find_pval.py:157-159: p[i][j] = count(background > r[i][j]) / N, the
 *                      division in binary64, stored in r's type.  d_sorted_bg is the background in
 *                      ascending order (float32 or float64, no NaN).
 * skr_pval_dist        replaces find_pval.py:126-128 for the closed-form scipy.stats families below:
 *                      p = 1 - dist(shape, loc, scale).cdf(r), evaluated in binary64 (support handling and
 *                      invalid parameters as rv_continuous.cdf).  shape is ignored by families without one.
 * skr_triu_extract     replaces find_dist.py:163: c[np.triu_indices(n, k=1)] (row-major) into d_out, which
 *                      holds skr_triu_count(n) = n(n-1)/2 values.
 * skr_pearson_pairs    r of npairs (i, j) pairs from prepared planes (find_dist.py:160-169 needs only a random
 *                      subset of the triangle): out[q] = alpha * <a_i, b_j>, binary64 accumulation.
 * ------------------------------------------------------------------------------------------ */
enum {
    SKR_DIST_NORM = 0,     /* scipy.stats.norm                */
    SKR_DIST_LOGNORM = 1,  /* lognorm(s)      shape = s       */
    SKR_DIST_CAUCHY = 2,
    SKR_DIST_EXPON = 3,
    SKR_DIST_RAYLEIGH = 4,
    SKR_DIST_UNIFORM = 5,
    SKR_DIST_PARETO = 6,   /* pareto(b)       shape = b       */
    SKR_DIST_EXPONPOW = 7, /* exponpow(b)     shape = b       */
    SKR_DIST_GAMMA = 8,    /* gamma(a)        shape = a   cdf = gammainc(a, x)        (regularised incomplete gamma) */
    SKR_DIST_CHI2 = 9      /* chi2(df)        shape = df  cdf = gammainc(df/2, x/2)                                  */
};
int skr_pval_empirical(const void* d_r, int r_is_f64, int64_t m, int64_t n, int64_t ld, const void* d_sorted_bg,
                       int bg_is_f64, int64_t N, void* d_p, int64_t ldp, void* stream);
int skr_pval_dist(const void* d_r, int r_is_f64, int64_t m, int64_t n, int64_t ld, int kind, double shape, double loc,
                  double scale, void* d_p, int64_t ldp, void* stream);
int64_t skr_triu_count(int64_t n);
int skr_triu_extract(const void* d_c, int c_is_f64, int64_t n, int64_t ld, void* d_out, void* stream);
int skr_pearson_pairs(const uint16_t* d_a_hi, const uint16_t* d_a_lo, const float* d_a_scale, const uint16_t* d_b_hi,
                      const uint16_t* d_b_lo, const float* d_b_scale, int64_t K, const int64_t* d_i, const int64_t* d_j,
                      int64_t npairs, double alpha, float* d_out, void* stream);

/* ------------------------------------------------------------------------------------------
 * Similarity graph of an r matrix (the numeric front half of seekr/kmer_leiden.py)
 *
 * skr_sim_threshold    replaces kmer_leiden.py:91-94 in place on the device:
 *                      c[c < cutoff] = 0, then (zero_diagonal) np.fill_diagonal(c, 0).  The comparison is made
 *                      in the matrix's type with the cutoff converted to it, as numpy does for a Python scalar;
 *                      NaN entries stay NaN.
 * skr_sim_edge_offsets replaces the counting half of kmer_leiden.py:103-104 ((df.values > 0)): entry (i, j) of
 *                      the thresholded, zero-diagonal matrix is positive iff the ORIGINAL value x has
 *                      !(x < cutoff) && x > 0 && i != j, so d_c is the untouched r matrix.  Writes offsets per
 *                      (row, column slice): d_offsets holds m * SKR_SIM_SLICES + 1 values, d_offsets[0] = 0, row i's
 *                      edges start at d_offsets[i * SKR_SIM_SLICES] (every SKR_SIM_SLICES-th value is the CSR row
 *                      offset) and the last value is the edge count (upper_only != 0: only j > i, one entry per
 *                      undirected edge).
 *                      d_c may be a block of rows of a larger matrix (a row shard of a rank, or one of the row
 *                      blocks a 250 000 x 250 000 result is produced in): row0 is the index of its first row in the
 *                      whole matrix, which places the diagonal (column row0 + i) and the upper half; d_src holds
 *                      whole-matrix row indices.  The same row0 applies to all three entry points.
 * skr_sim_edge_fill    replaces kmer_leiden.py:104 (df.values[df.values > 0].flatten()) and np.nonzero of the
 *                      adjacency: edges in row-major order, d_src (may be NULL for CSR form) / d_dst int32,
 *                      d_weight in the matrix's type; each array holds d_offsets[m * SKR_SIM_SLICES] entries.
 * ------------------------------------------------------------------------------------------ */
#define SKR_SIM_SLICES 8 /* column slices per row in d_offsets */
int skr_sim_threshold(void* d_c, int c_is_f64, int64_t m, int64_t n, int64_t ld, int64_t row0, double cutoff,
                      int zero_diagonal, void* stream);
int skr_sim_edge_offsets(const void* d_c, int c_is_f64, int64_t m, int64_t n, int64_t ld, int64_t row0, double cutoff,
                         int upper_only, int64_t* d_offsets, void* stream);
int skr_sim_edge_fill(const void* d_c, int c_is_f64, int64_t m, int64_t n, int64_t ld, int64_t row0, double cutoff,
                      int upper_only, const int64_t* d_offsets, int32_t* d_src, int32_t* d_dst, void* d_weight,
                      void* stream);
/* The counting half fused into the GEMM ("emit the edge list from the GEMM epilogue"):
 * skr_pearson_gemm_edges is skr_pearson_gemm for a float32 result whose epilogue also counts, per (row, column
 * slice), the finished r values that are edges (same predicate as skr_sim_edge_offsets) and then scans the counts
 * into d_offsets -- the offsets pass never reads the matrix; skr_sim_edge_fill follows as usual.  symmetric != 0
 * needs upper_only != 0 (only the tiles on and above the diagonal are computed there).
 * skr_sim_slice_width: columns per slice for an n-column matrix; skr_sim_offsets_scan: the scan alone, for counts
 * already sitting in d_offsets[1 ..]. */
int skr_pearson_gemm_edges(const uint16_t* d_a_hi, const uint16_t* d_a_lo, const float* d_a_scale, int64_t m,
                           const uint16_t* d_b_hi, const uint16_t* d_b_lo, const float* d_b_scale, int64_t n, int64_t K,
                           double alpha, float* d_c, int64_t ldc, int symmetric, int64_t row0, double cutoff,
                           int upper_only, int64_t* d_offsets, void* stream);
int64_t skr_sim_slice_width(int64_t n);
int skr_sim_offsets_scan(int64_t* d_offsets, int64_t m, void* stream);

/* ------------------------------------------------------------------------------------------
 * Collectives over NVLink peer memory (one process per GPU; SURVEY section 8e: the Log2.post minimum
 * of kmer_counts.py:207-208 spans all row shards)
 *
 * skr_peer_alloc / open / close / free   an exchange buffer (cudaMalloc, zeroed) and its 64-byte CUDA IPC
 *                      handle; a peer process maps it with skr_peer_open.
 * skr_min_exchange     all-reduce of the minimum cell in ONE single-warp kernel: every rank stores its cell,
 *                      tagged with `epoch`, into every peer's buffer (P2P stores) and reduces the `world` cells
 *                      that arrive in its own.  d_peers[t] = rank t's buffer as mapped here (d_peers[rank] = own
 *                      buffer), each 2 * world 64-bit words.  epoch starts at 1 and grows by 1 per call on every
 *                      rank.  *d_err becomes 1 if a peer did not show up within the spin limit.
 * ------------------------------------------------------------------------------------------ */
int skr_peer_alloc(size_t bytes, void** d_out, unsigned char* handle64);
int skr_peer_open(const unsigned char* handle64, void** d_out);
int skr_peer_close(void* d_ptr);
int skr_peer_free(void* d_ptr);
int skr_min_exchange(SkrMinCell* d_cell, void* const* d_peers, int world, int rank, uint64_t epoch, int* d_err,
                     void* stream);
/* the same, but the launch returns at once when *d_skip == skip_value (d_skip may be NULL; 0 = 1); every rank must
 * hold the same value there -- the speculative Log2.post route exchanges its flag first.  flag_value != 0: the
 * cell's second word is an epoch flag, set when it equals flag_value, and the OR over the ranks comes back as
 * flag_value / 0 (SkrPostSpec.zero_seen seen as a cell together with zero_col).  The spin limit of both exchanges
 * is 60 s, or SEEKR_B200_PEER_TIMEOUT_S. */
int skr_min_exchange_skip(SkrMinCell* d_cell, void* const* d_peers, int world, int rank, uint64_t epoch,
                          const uint32_t* d_skip, uint32_t skip_value, uint32_t flag_value, int* d_err, void* stream);
/* skr_flag_or_exchange   OR of one flag per rank (set = *d_flag == flag_value; the OR comes back as flag_value / 0) that
 *                      does not wait in the common case: a rank whose own flag is set stores its word into every peer and
 *                      returns -- the OR is already known to it; only a rank whose flag is clear waits for the others.
 *                      Every word also carries the outcomes of the writer's previous 31 epochs, so that a straggler that
 *                      finds a peer's slot overwritten by a later epoch still learns what its own epoch turned out to
 *                      be; d_state (one zero-initialised word per rank, kept by the caller) holds that history.  The
 *                      peer buffers are those of the minimum exchange, skr_min_exchange_bytes(world) bytes each; the
 *                      epochs of this exchange count on their own, from 1. */
int64_t skr_min_exchange_bytes(int world);
int skr_flag_or_exchange(uint32_t* d_flag, uint32_t flag_value, void* const* d_peers, int world, int rank, uint64_t epoch,
                         uint32_t* d_state, int* d_err, void* stream);
/* skr_colstat_exchange   all-reduce(sum) of the n binary64 column partials of skr_col_partial_f64 over the ranks,
 *                      fused with skr_col_finish_f64 (divide by the total row count, optional sqrt, fp32, quality
 *                      flag) in ONE kernel: P2P stores of the partials into every peer, an epoch flag per rank, a
 *                      rank-ordered sum (all ranks get the same bits).  Peer buffers hold
 *                      skr_colstat_exchange_bytes(world, n_cap) bytes, zero-initialised (skr_peer_alloc). */
/* skr_colsum_exchange    the ONE exchange of the accurate column statistics (d_colsum / d_colsq of skr_count_ex, laid out as
 *                      [2][cols] doubles in d_acc) fused with skr_colstat_finish: same protocol and buffers as
 *                      skr_colstat_exchange with n_cap >= 2 * cols; d_flags[2] is NOT cleared here. */
int skr_colsum_exchange(const double* d_acc, void* const* d_peers, int world, int rank, uint64_t epoch, int64_t cols,
                        int64_t n_cap, int64_t total_rows, float* d_mean, float* d_std, int* d_flags, int* d_err,
                        void* stream);
int64_t skr_colstat_exchange_bytes(int world, int64_t n_cap);
int skr_colstat_exchange(const double* d_acc, void* const* d_peers, int world, int rank, uint64_t epoch, int64_t n,
                         int64_t n_cap, int64_t total_rows, int take_sqrt, float* d_out, int* d_flag, int* d_err,
                         void* stream);

/* ------------------------------------------------------------------------------------------
 * Text output of a host float32 matrix (replaces the DataFrame.to_csv / np.savetxt
 * calls of kmer_counts.py:235-241), formatted on `threads` host threads (0 = all), byte-identical to them:
 *   style 0  pandas' cell text for float32 (numpy's shortest round-trip repr; NaN -> empty cell)
 *   style 1  "%1.6f" (np.savetxt)
 *   style 2  data is a float64 matrix (ld in elements): pandas' cell text for float64 (shortest round-trip repr,
 *            positional for 1e-4 <= |x| < 1e16; NaN -> empty cell) -- the seekr_pearson CSV output of
 *            console_scripts.py:636-638
 * header (header_len bytes, may be NULL) is written first; labels + label_offs[m+1] (may be NULL) give the
 * already CSV-quoted first field of every row.  skr_format_f32 formats n values, one per line (for tests).
 * ------------------------------------------------------------------------------------------ */
int skr_csv_write(const char* path, const void* data, int64_t m, int64_t cols, int64_t ld, const char* header,
                  int64_t header_len, const char* labels, const int64_t* label_offs, int style, int threads);
int skr_format_f32(const float* values, int64_t n, int style, char* out, int64_t capacity, int64_t* written);
int skr_format_f64(const double* values, int64_t n, char* out, int64_t capacity, int64_t* written); /* style 2 */

/* ------------------------------------------------------------------------------------------
 * Text input of a labelled count matrix (replaces pd.read_csv(path, index_col=0) of
 * console_scripts.py:628-629 for the files kmer_counts.py:235-238 writes): parsed on `threads` host threads
 * (0 = all) into binary64 with the same bits as pandas' default C parser (its 17-digit accumulate-and-scale
 * procedure is restated, including its rounding errors).  Files outside the plain form (quotes, labels pandas
 * would convert, cells that are not plain numbers, ragged rows) return SKR_CSV_UNSUPPORTED: use pandas.
 * The table owns its memory until skr_csv_free; labels / columns are concatenated bytes with rows+1 / cols+1 offsets.
 * ------------------------------------------------------------------------------------------ */
#define SKR_CSV_UNSUPPORTED 100
typedef struct SkrCsvTable SkrCsvTable;
int skr_csv_read(const char* path, int threads, SkrCsvTable** out);
void skr_csv_free(SkrCsvTable* t);
int64_t skr_csv_rows(const SkrCsvTable* t);
int64_t skr_csv_cols(const SkrCsvTable* t);
const double* skr_csv_values(const SkrCsvTable* t);
const char* skr_csv_labels(const SkrCsvTable* t);
const int64_t* skr_csv_label_offsets(const SkrCsvTable* t);
const char* skr_csv_columns(const SkrCsvTable* t);
const int64_t* skr_csv_column_offsets(const SkrCsvTable* t);
int skr_csv_all_integer(const SkrCsvTable* t); /* 1: every cell is an integer literal (pandas: int64 columns) */

/* ------------------------------------------------------------------------------------------
 * Host <-> device plumbing used by the Python layer (thin wrappers; no reference counterpart)
 * ------------------------------------------------------------------------------------------ */
int skr_host_alloc(size_t bytes, void** out); /* pinned host memory from the library's pool */
int skr_host_alloc_pooled(size_t bytes, void** out); /* only if the pool holds a slab that large: *out = NULL otherwise */
void skr_host_free(void* p);                  /* returns it to the pool */
void skr_host_pool_trim(void);                /* releases every pooled slab */
int skr_copy_h2d(void* d_dst, const void* h_src, size_t bytes, void* stream);
int skr_copy_d2h(void* h_dst, const void* d_src, size_t bytes, void* stream);
/* rows x row_bytes with different pitches on either side */
int skr_copy_d2h_2d(void* h_dst, size_t h_pitch, const void* d_src, size_t d_pitch, size_t row_bytes, size_t rows,
                    void* stream);
int skr_copy_h2d_2d(void* d_dst, size_t d_pitch, const void* h_src, size_t h_pitch, size_t row_bytes, size_t rows,
                    void* stream);
int skr_stream_sync(void* stream);
int skr_device_count(int* out);

/* number of kernels launched by this library on the calling thread since the last reset
 * (bench.py reports it as gpu_launches) */
int64_t skr_launch_count(int reset);

/* Host-side evaluation of the same binade-jumping routine the count kernel uses for the
 * c-fold binary64 sum (unit-tested against the literal loop on the CPU). */
double skr_chain_sum_host(double increment, uint32_t count);

#ifdef __cplusplus
}
#endif
#endif /* SEEKR_B200_H */
