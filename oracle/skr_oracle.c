/*
 * CPU oracle for the SEEKR hot path, C restatement -- TEST INFRASTRUCTURE ONLY.
 *
 * Same algorithm as oracle/seekr_oracle.py (which follows the reference file by
 * file), written in C so the parity tests and bench.py's cpu_baseline leg can
 * run it at sizes where the reference's pure-Python loop takes minutes.
 * Nothing under seekr_b200/ links or loads this file.
 *
 * Parity status: PINNED -- tests/test_oracle.py checks every entry point
 * against the Python oracle, the reference's golden files and the fixtures
 * generated from the unmodified reference (tests/golden/make_golden.py).
 *
 * Reference lines restated here:
 *   seekr/kmer_counts.py:140-151  occurrences   -> orc_count_row / orc_count_matrix
 *   seekr/kmer_counts.py:189-192  log2_norm     -> orc_log2_norm
 *   seekr/kmer_counts.py:165-169  center        -> orc_col_mean + orc_sub_vec
 *   seekr/kmer_counts.py:171-174  standardize   -> orc_col_std  + orc_div_vec
 *   seekr/kmer_counts.py:207-209  Log2.post     -> orc_post_log2
 *   seekr/pearson.py:35-38        row standardise -> orc_row_standardize
 *
 * Build: see oracle/Makefile (gcc -O2 -fopenmp -ffp-contract=off; no -ffast-math:
 * every float operation below must stay one IEEE operation).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

int orc_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* kmer_counts.py:143-150: `counts[kmer] += increment` once per window, in binary64,
 * then one rounding to the row's dtype on assignment. */
static double chain_sum(double inc, int64_t c) {
    double acc = 0.0;
    for (int64_t i = 0; i < c; ++i) acc += inc;
    return acc;
}

/* One sequence.  lut[ch] = column digit of letter ch in the alphabet, 255 = not
 * in the alphabet.  The sequence is upper-cased by the Reader before counting
 * (fasta_reader.py:55,62), so `lut` is built from upper-case letters and `seq`
 * must already be upper case.  hist is scratch of 4^k int64.  Returns -1 for
 * L == k-1 (ZeroDivisionError in the reference), 0 otherwise. */
static int count_row(const uint8_t* seq, int64_t L, int k, const uint8_t* lut, int64_t* hist,
                     int64_t nbins, float* row_f32, double* row_f64) {
    int64_t n = L - k + 1;
    if (n == 0) return -1;
    if (n < 0) return 0; /* range(negative) is empty: row stays zero */
    double inc = 1000.0 / (double)n;
    memset(hist, 0, sizeof(int64_t) * (size_t)nbins);
    int64_t mask = nbins - 1;
    int64_t idx = 0;
    int valid = 0; /* consecutive valid letters ending at position p */
    for (int64_t p = 0; p < L; ++p) {
        uint8_t d = lut[seq[p]];
        if (d == 255) {
            valid = 0;
            idx = 0;
        } else {
            idx = ((idx << 2) | d) & mask;
            if (valid < k) ++valid;
            if (valid == k) hist[idx] += 1;
        }
    }
    for (int64_t b = 0; b < nbins; ++b) {
        if (hist[b]) {
            double v = chain_sum(inc, hist[b]);
            if (row_f32) row_f32[b] = (float)v;
            if (row_f64) row_f64[b] = v;
        }
    }
    return 0;
}

int orc_count_row(const uint8_t* seq, int64_t L, int k, const uint8_t* lut, float* row_f32, double* row_f64) {
    int64_t nbins = (int64_t)1 << (2 * k);
    int64_t* hist = (int64_t*)malloc(sizeof(int64_t) * (size_t)nbins);
    if (!hist) return -2;
    int rc = count_row(seq, L, k, lut, hist, nbins, row_f32, row_f64);
    free(hist);
    return rc;
}

/* Integer histogram only (what the CUDA kernel accumulates before scaling). */
int orc_int_counts(const uint8_t* seq, int64_t L, int k, const uint8_t* lut, int64_t* hist) {
    int64_t nbins = (int64_t)1 << (2 * k);
    memset(hist, 0, sizeof(int64_t) * (size_t)nbins);
    int64_t mask = nbins - 1, idx = 0;
    int valid = 0;
    for (int64_t p = 0; p < L; ++p) {
        uint8_t d = lut[seq[p]];
        if (d == 255) { valid = 0; idx = 0; }
        else {
            idx = ((idx << 2) | d) & mask;
            if (valid < k) ++valid;
            if (valid == k) hist[idx] += 1;
        }
    }
    return 0;
}

/* kmer_counts.py:196-200.  seqs = concatenated upper-case letters, offs[m+1].
 * out is m x 4^k float32, zero-initialised by the caller (np.zeros).
 * Sequences are independent, so rows are spread over OpenMP threads; the
 * reference itself is single-threaded.  Returns the index+1 of the first
 * sequence with L == k-1 (negated), or 0. */
int64_t orc_count_matrix(const uint8_t* seqs, const int64_t* offs, int64_t m, int k, const uint8_t* lut,
                         float* out, int nthreads) {
    int64_t nbins = (int64_t)1 << (2 * k);
    int64_t bad = 0;
#ifdef _OPENMP
    if (nthreads <= 0) nthreads = omp_get_max_threads();
#pragma omp parallel num_threads(nthreads)
#endif
    {
        int64_t* hist = (int64_t*)malloc(sizeof(int64_t) * (size_t)nbins);
#ifdef _OPENMP
#pragma omp for schedule(dynamic, 16)
#endif
        for (int64_t i = 0; i < m; ++i) {
            int rc = count_row(seqs + offs[i], offs[i + 1] - offs[i], k, lut, hist, nbins,
                               out + (size_t)i * (size_t)nbins, NULL);
            if (rc == -1) {
#ifdef _OPENMP
#pragma omp critical
#endif
                { if (bad == 0 || i + 1 < bad) bad = i + 1; }
            }
        }
        free(hist);
    }
    return -bad;
}

/* kmer_counts.py:189-192: counts += 1 ; log2, element by element in fp32. */
void orc_log2_norm(float* a, int64_t n) {
#ifdef _OPENMP
#pragma omp parallel for
#endif
    for (int64_t i = 0; i < n; ++i) {
        float v = a[i] + 1.0f;
        a[i] = log2f(v);
    }
}

/* np.mean(float32, axis=0): numpy reduces axis 0 of a C-ordered matrix row after row
 * into an fp32 accumulator (add.reduce), then divides in binary64 and rounds to fp32
 * (numpy/_core/_methods.py:_mean).  Columns are independent -> threads over columns. */
void orc_col_mean(const float* a, int64_t m, int64_t cols, float* mean) {
#ifdef _OPENMP
#pragma omp parallel for schedule(static)
#endif
    for (int64_t jb = 0; jb < cols; jb += 64) {
        int64_t je = jb + 64 < cols ? jb + 64 : cols;
        float acc[64];
        for (int64_t j = jb; j < je; ++j) acc[j - jb] = 0.0f;
        for (int64_t i = 0; i < m; ++i) {
            const float* r = a + (size_t)i * (size_t)cols;
            for (int64_t j = jb; j < je; ++j) acc[j - jb] = acc[j - jb] + r[j];
        }
        for (int64_t j = jb; j < je; ++j) mean[j] = (float)((double)acc[j - jb] / (double)m);
    }
}

/* np.std(float32, axis=0, ddof=0) (numpy/_core/_methods.py:_var,_std):
 * arrmean = sum/m (as above); x = a - arrmean; x = x*x; sum again; /m in binary64 -> fp32; sqrt in fp32. */
void orc_col_std(const float* a, int64_t m, int64_t cols, float* std) {
    float* arrmean = (float*)malloc(sizeof(float) * (size_t)cols);
    orc_col_mean(a, m, cols, arrmean);
#ifdef _OPENMP
#pragma omp parallel for schedule(static)
#endif
    for (int64_t jb = 0; jb < cols; jb += 64) {
        int64_t je = jb + 64 < cols ? jb + 64 : cols;
        float acc[64];
        for (int64_t j = jb; j < je; ++j) acc[j - jb] = 0.0f;
        for (int64_t i = 0; i < m; ++i) {
            const float* r = a + (size_t)i * (size_t)cols;
            for (int64_t j = jb; j < je; ++j) {
                float d = r[j] - arrmean[j];
                float q = d * d;
                acc[j - jb] = acc[j - jb] + q;
            }
        }
        for (int64_t j = jb; j < je; ++j) {
            float var = (float)((double)acc[j - jb] / (double)m);
            std[j] = sqrtf(var);
        }
    }
    free(arrmean);
}

/* counts -= mean (kmer_counts.py:169) / counts /= std (:175).  numpy computes an fp32
 * matrix against a binary64 vector in binary64 and rounds once to fp32; against an fp32
 * vector it is one fp32 operation.  vec_is_f64 selects which. */
void orc_sub_vec(float* a, int64_t m, int64_t cols, const void* vec, int vec_is_f64) {
#ifdef _OPENMP
#pragma omp parallel for
#endif
    for (int64_t i = 0; i < m; ++i) {
        float* r = a + (size_t)i * (size_t)cols;
        if (vec_is_f64) {
            const double* v = (const double*)vec;
            for (int64_t j = 0; j < cols; ++j) r[j] = (float)((double)r[j] - v[j]);
        } else {
            const float* v = (const float*)vec;
            for (int64_t j = 0; j < cols; ++j) r[j] = r[j] - v[j];
        }
    }
}

void orc_div_vec(float* a, int64_t m, int64_t cols, const void* vec, int vec_is_f64) {
#ifdef _OPENMP
#pragma omp parallel for
#endif
    for (int64_t i = 0; i < m; ++i) {
        float* r = a + (size_t)i * (size_t)cols;
        if (vec_is_f64) {
            const double* v = (const double*)vec;
            for (int64_t j = 0; j < cols; ++j) r[j] = (float)((double)r[j] / v[j]);
        } else {
            const float* v = (const float*)vec;
            for (int64_t j = 0; j < cols; ++j) r[j] = r[j] / v[j];
        }
    }
}

/* np.min over the whole matrix; NaN propagates (kmer_counts.py:208). */
float orc_min(const float* a, int64_t n) {
    float mn = INFINITY;
    int has_nan = 0;
#ifdef _OPENMP
#pragma omp parallel for reduction(min : mn) reduction(| : has_nan)
#endif
    for (int64_t i = 0; i < n; ++i) {
        float v = a[i];
        if (v != v) has_nan = 1;
        else if (v < mn) mn = v;
    }
    return has_nan ? NAN : mn;
}

/* kmer_counts.py:207-209: counts += |min| ; counts += 1 ; log2 -- three separate fp32 steps. */
void orc_post_log2(float* a, int64_t n) {
    float shift = fabsf(orc_min(a, n));
#ifdef _OPENMP
#pragma omp parallel for
#endif
    for (int64_t i = 0; i < n; ++i) {
        float v = a[i] + shift;
        v = v + 1.0f;
        a[i] = log2f(v);
    }
}

/* pearson.py:35-38 for fp32 input: row mean and std (ddof=0) by numpy's pairwise fp32
 * summation are within a few ulp of the binary64 value; the oracle computes them in
 * binary64 and rounds, which is inside the 1e-5 band the tests allow for r. */
void orc_row_standardize(const float* a, int64_t m, int64_t cols, float* out) {
#ifdef _OPENMP
#pragma omp parallel for
#endif
    for (int64_t i = 0; i < m; ++i) {
        const float* r = a + (size_t)i * (size_t)cols;
        float* o = out + (size_t)i * (size_t)cols;
        double s = 0.0;
        for (int64_t j = 0; j < cols; ++j) s += r[j];
        float mean = (float)(s / (double)cols);
        double q = 0.0;
        for (int64_t j = 0; j < cols; ++j) {
            float d = r[j] - mean;
            o[j] = d;
            q += (double)d * (double)d;
        }
        /* np.std of the centred row: its own mean is ~0; numpy subtracts it again */
        double s2 = 0.0;
        for (int64_t j = 0; j < cols; ++j) s2 += o[j];
        float m2 = (float)(s2 / (double)cols);
        q = 0.0;
        for (int64_t j = 0; j < cols; ++j) {
            float d = o[j] - m2;
            q += (double)(d * d);
        }
        float sd = sqrtf((float)(q / (double)cols));
        for (int64_t j = 0; j < cols; ++j) o[j] = o[j] / sd;
    }
}
