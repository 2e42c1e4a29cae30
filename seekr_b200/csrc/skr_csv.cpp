// Text output of the count matrix: once counting runs on the GPU, writing the CSV is
// what the default `seekr_kmer_counts` spends its time on (pandas DataFrame.to_csv / np.savetxt format every
// value through Python objects; kmer_counts.py:235-241).  This writer formats rows on all host threads and
// produces the same bytes:
//   style 0  what pandas writes for a float32 frame: numpy's shortest round-trip text of the float32 value
//            (positional for 1e-4 <= |x| < 1e6, otherwise d.ddde+XX), '' for NaN, 'inf' / '-inf';
//   style 1  np.savetxt(fmt="%1.6f"): C "%.6f" of the value widened to double, 'nan' / 'inf' / '-inf'.
// Row labels (already CSV-quoted by the caller) and the header line are optional.
#include <charconv>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "skr_common.h"

namespace {

// numpy float32 -> str (np.float32.__str__ / what pandas puts into a CSV cell); returns the new end pointer.
// Needs up to 24 characters.
inline char* format_f32_repr(char* out, float v) {
    if (v != v) return out;  // pandas na_rep = ''
    if (std::isinf(v)) {
        if (v < 0) *out++ = '-';
        memcpy(out, "inf", 3);
        return out + 3;
    }
    if (v == 0.0f) {
        if (std::signbit(v)) *out++ = '-';
        memcpy(out, "0.0", 3);
        return out + 3;
    }
    char sci[32];
    // shortest digits that round-trip as float32, in the form d[.ddd]e[+-]XX
    auto res = std::to_chars(sci, sci + sizeof(sci), v, std::chars_format::scientific);
    char* p = sci;
    if (*p == '-') { *out++ = '-'; ++p; }
    char digits[16];
    int nd = 0;
    for (; p < res.ptr && *p != 'e'; ++p)
        if (*p != '.') digits[nd++] = *p;
    ++p;  // 'e'
    int esign = 1;
    if (*p == '-') { esign = -1; ++p; } else if (*p == '+') { ++p; }
    int e10 = 0;
    for (; p < res.ptr; ++p) e10 = e10 * 10 + (*p - '0');
    e10 *= esign;  // value = d.ddd x 10^e10
    // numpy (2.x) prints a float32 positionally for 1e-4 <= |x| < 1e6, tested on the value itself
    // (float32(1e-4) = 9.9999997e-05 is below the limit and prints as 1e-04)
    const double av = std::fabs((double)v);
    if (av >= 1e-4 && av < 1e6) {
        if (e10 >= 0) {
            for (int i = 0; i <= e10; ++i) *out++ = i < nd ? digits[i] : '0';
            *out++ = '.';
            if (nd > e10 + 1) {
                for (int i = e10 + 1; i < nd; ++i) *out++ = digits[i];
            } else {
                *out++ = '0';
            }
        } else {
            *out++ = '0';
            *out++ = '.';
            for (int i = 0; i < -e10 - 1; ++i) *out++ = '0';
            for (int i = 0; i < nd; ++i) *out++ = digits[i];
        }
    } else {
        *out++ = digits[0];
        if (nd > 1) {
            *out++ = '.';
            for (int i = 1; i < nd; ++i) *out++ = digits[i];
        }
        *out++ = 'e';
        *out++ = e10 < 0 ? '-' : '+';
        const int a = e10 < 0 ? -e10 : e10;
        if (a < 10) *out++ = '0';
        auto r2 = std::to_chars(out, out + 4, a);
        out = r2.ptr;
    }
    return out;
}

// float64 cell text of DataFrame.to_csv: the shortest digits that round-trip as binary64 (Python / numpy repr),
// positional for 1e-4 <= |x| < 1e16, otherwise d.ddde+XX; NaN -> empty cell.  Needs up to 26 characters.
inline char* format_f64_repr(char* out, double v) {
    if (v != v) return out;
    if (std::isinf(v)) {
        if (v < 0) *out++ = '-';
        memcpy(out, "inf", 3);
        return out + 3;
    }
    if (v == 0.0) {
        if (std::signbit(v)) *out++ = '-';
        memcpy(out, "0.0", 3);
        return out + 3;
    }
    char sci[40];
    auto res = std::to_chars(sci, sci + sizeof(sci), v, std::chars_format::scientific);
    char* p = sci;
    if (*p == '-') { *out++ = '-'; ++p; }
    char digits[24];
    int nd = 0;
    for (; p < res.ptr && *p != 'e'; ++p)
        if (*p != '.') digits[nd++] = *p;
    ++p;
    int esign = 1;
    if (*p == '-') { esign = -1; ++p; } else if (*p == '+') { ++p; }
    int e10 = 0;
    for (; p < res.ptr; ++p) e10 = e10 * 10 + (*p - '0');
    e10 *= esign;
    if (e10 >= -4 && e10 < 16) {
        if (e10 >= 0) {
            for (int i = 0; i <= e10; ++i) *out++ = i < nd ? digits[i] : '0';
            *out++ = '.';
            if (nd > e10 + 1) {
                for (int i = e10 + 1; i < nd; ++i) *out++ = digits[i];
            } else {
                *out++ = '0';
            }
        } else {
            *out++ = '0';
            *out++ = '.';
            for (int i = 0; i < -e10 - 1; ++i) *out++ = '0';
            for (int i = 0; i < nd; ++i) *out++ = digits[i];
        }
    } else {
        *out++ = digits[0];
        if (nd > 1) {
            *out++ = '.';
            for (int i = 1; i < nd; ++i) *out++ = digits[i];
        }
        *out++ = 'e';
        *out++ = e10 < 0 ? '-' : '+';
        const int a = e10 < 0 ? -e10 : e10;
        if (a < 10) *out++ = '0';
        auto r2 = std::to_chars(out, out + 4, a);
        out = r2.ptr;
    }
    return out;
}

inline char* format_f32_fixed6(char* out, float v) {
    if (v != v) { memcpy(out, "nan", 3); return out + 3; }
    if (std::isinf(v)) {
        if (v < 0) *out++ = '-';
        memcpy(out, "inf", 3);
        return out + 3;
    }
    // "%.6f" of the exact binary value; float32 magnitudes need at most 39 + 1 + 6 digits
    const int n = snprintf(out, 56, "%.6f", (double)v);
    return out + n;
}

constexpr size_t kMaxCell = 56;

void format_rows(const void* data_any, int64_t r0, int64_t r1, int64_t cols, int64_t ld, const char* labels,
                 const int64_t* label_offs, int style, std::string& buf) {
    std::vector<char> line;
    for (int64_t r = r0; r < r1; ++r) {
        const size_t label_len = labels ? (size_t)(label_offs[r + 1] - label_offs[r]) : 0;
        line.resize(label_len + 1 + (size_t)cols * (kMaxCell + 1) + 2);
        char* p = line.data();
        if (labels) {
            memcpy(p, labels + label_offs[r], label_len);
            p += label_len;
            *p++ = ',';
        }
        if (style == 2) {
            const double* row = (const double*)data_any + r * ld;
            for (int64_t c = 0; c < cols; ++c) {
                p = format_f64_repr(p, row[c]);
                *p++ = c + 1 < cols ? ',' : '\n';
            }
        } else {
            const float* row = (const float*)data_any + r * ld;
            for (int64_t c = 0; c < cols; ++c) {
                p = style == 0 ? format_f32_repr(p, row[c]) : format_f32_fixed6(p, row[c]);
                *p++ = c + 1 < cols ? ',' : '\n';
            }
        }
        if (cols == 0) *p++ = '\n';
        buf.append(line.data(), (size_t)(p - line.data()));
    }
}

}  // namespace

extern "C" int skr_format_f32(const float* values, int64_t n, int style, char* out, int64_t capacity, int64_t* written) {
    if (!values || !out || !written || n < 0) return skr::fail(SKR_ERR_ARG, "skr_format_f32: bad argument");
    if (capacity < n * (int64_t)(kMaxCell + 1)) return skr::fail(SKR_ERR_ARG, "skr_format_f32: buffer too small");
    char* p = out;
    for (int64_t i = 0; i < n; ++i) {
        p = style == 0 ? format_f32_repr(p, values[i]) : format_f32_fixed6(p, values[i]);
        *p++ = '\n';
    }
    *written = p - out;
    return SKR_OK;
}

extern "C" int skr_format_f64(const double* values, int64_t n, char* out, int64_t capacity, int64_t* written) {
    if (!values || !out || !written || n < 0) return skr::fail(SKR_ERR_ARG, "skr_format_f64: bad argument");
    if (capacity < n * (int64_t)(kMaxCell + 1)) return skr::fail(SKR_ERR_ARG, "skr_format_f64: buffer too small");
    char* p = out;
    for (int64_t i = 0; i < n; ++i) {
        p = format_f64_repr(p, values[i]);
        *p++ = '\n';
    }
    *written = p - out;
    return SKR_OK;
}

extern "C" int skr_csv_write(const char* path, const void* data, int64_t m, int64_t cols, int64_t ld, const char* header,
                             int64_t header_len, const char* labels, const int64_t* label_offs, int style, int threads) {
    if (!path || (!data && m * cols > 0) || m < 0 || cols < 0 || ld < cols || (labels && !label_offs) || style < 0 || style > 2)
        return skr::fail(SKR_ERR_ARG, "skr_csv_write: bad argument");
    FILE* f = fopen(path, "wb");
    if (!f) return skr::fail(SKR_ERR_IO, "skr_csv_write: cannot open %s", path);
    bool ok = true;
    if (header && header_len > 0) ok = fwrite(header, 1, (size_t)header_len, f) == (size_t)header_len;
    if (threads < 1) threads = (int)std::thread::hardware_concurrency();
    if (threads < 1) threads = 1;
    // slabs of rows: formatted in parallel, written in order; the slab size bounds the memory held (~64 MB of text)
    const int64_t row_text = cols * 12 + 64;
    int64_t rows_per_task = (4 << 20) / (row_text > 0 ? row_text : 1);
    if (rows_per_task < 1) rows_per_task = 1;
    const int64_t slab = rows_per_task * threads;
    for (int64_t s0 = 0; s0 < m && ok; s0 += slab) {
        const int64_t s1 = s0 + slab < m ? s0 + slab : m;
        const int ntask = (int)((s1 - s0 + rows_per_task - 1) / rows_per_task);
        std::vector<std::string> bufs((size_t)ntask);
        std::vector<std::thread> pool;
        for (int t = 0; t < ntask; ++t) {
            const int64_t r0 = s0 + t * rows_per_task, r1 = r0 + rows_per_task < s1 ? r0 + rows_per_task : s1;
            pool.emplace_back([&, t, r0, r1]() { format_rows(data, r0, r1, cols, ld, labels, label_offs, style, bufs[(size_t)t]); });
        }
        for (auto& th : pool) th.join();
        for (int t = 0; t < ntask && ok; ++t) ok = fwrite(bufs[(size_t)t].data(), 1, bufs[(size_t)t].size(), f) == bufs[(size_t)t].size();
    }
    if (fclose(f) != 0) ok = false;
    if (!ok) return skr::fail(SKR_ERR_IO, "skr_csv_write: short write to %s", path);
    return SKR_OK;
}
