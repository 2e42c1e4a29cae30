// Text input of a labelled count matrix: `seekr_pearson a.csv b.csv` reads both files
// with pd.read_csv(path, index_col=0) (console_scripts.py:628-629), minutes at 50 000 x 4 096 cells.  This reader
// splits the file into lines, parses the lines on all host threads and returns the cells as binary64 -- the type
// pandas gives them -- with the SAME bits: pandas' default C parser does not round correctly, it accumulates at
// most 17 digits into a double and scales by one power of ten (its `precise_xstrtod`); `parse_cell` below restates that
// procedure, so a value that pandas gets one ulp off comes out one ulp off here too.
//
// Only the plain shape seekr itself writes is handled (kmer_counts.py:235-238): one header line, an index label
// first on every row, unquoted fields, decimal / scientific numbers, empty cells (NaN), inf / -inf.  Anything
// else -- a quote character anywhere, a label pandas would convert (numbers, booleans, NA spellings), a cell
// that is not a plain number, ragged rows -- reports SKR_CSV_UNSUPPORTED and the caller uses pandas instead.
#include <cfloat>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <string>
#include <thread>
#include <vector>

#include "skr_common.h"

struct SkrCsvTable {
    int64_t rows = 0, cols = 0;
    double* values = nullptr;
    std::string labels;                 // index labels, concatenated
    std::vector<int64_t> label_offs;    // rows + 1
    std::string columns;                // column names (without the index name), concatenated
    std::vector<int64_t> column_offs;   // cols + 1
    int all_integer = 0;                // every cell is a plain integer literal (pandas would give int64 columns)
};

namespace {

inline bool is_digit(char c) { return c >= '0' && c <= '9'; }

// 1e0 .. 1e308, each the correctly rounded binary64 value of the literal (strtod at start-up)
struct Pow10Table {
    double v[309];
    Pow10Table() {
        char lit[16];
        for (int i = 0; i <= 308; ++i) {
            snprintf(lit, sizeof(lit), "1e%d", i);
            v[i] = strtod(lit, nullptr);
        }
    }
    double operator[](int i) const { return v[i]; }
};
const Pow10Table kPow10;

// pandas/_libs/src/parser/tokenizer.c precise_xstrtod(decimal='.', sci='E', tsep='\0') -- the converter behind
// float_precision=None / 'high' -- on the field [p, end).
// Returns false when the field is not consumed completely.  *integer: no '.', no exponent.
inline bool parse_cell(const char* p, const char* end, double* out, bool* integer) {
    const int max_digits = 17;
    bool negative = false;
    if (p < end && (*p == '-' || *p == '+')) negative = *p++ == '-';
    int exponent = 0, num_digits = 0, num_decimals = 0;
    double number = 0.;
    *integer = true;
    while (p < end && is_digit(*p)) {
        if (num_digits < max_digits) {
            number = number * 10. + (*p - '0');
            num_digits++;
        } else {
            ++exponent;
        }
        p++;
    }
    if (p < end && *p == '.') {
        *integer = false;
        p++;
        while (num_digits < max_digits && p < end && is_digit(*p)) {
            number = number * 10. + (*p - '0');
            p++;
            num_digits++;
            num_decimals++;
        }
        if (num_digits >= max_digits)
            while (p < end && is_digit(*p)) ++p;
        exponent -= num_decimals;
    }
    if (num_digits == 0) return false;
    if (negative) number = -number;
    if (p < end && (*p == 'e' || *p == 'E')) {
        *integer = false;
        ++p;
        bool eneg = false;
        if (p < end && (*p == '-' || *p == '+')) eneg = *p++ == '-';
        int n = 0, nd = 0;
        while (p < end && is_digit(*p)) {
            if (n < 100000) n = n * 10 + (*p - '0');
            nd++;
            p++;
        }
        if (nd == 0) return false;
        exponent += eneg ? -n : n;
    }
    if (p != end) return false;
    // one multiplication or division by a correctly rounded power of ten
    if (exponent > 308) return false;  // pandas reports a range error and re-parses: left to pandas
    if (exponent > 0) {
        number *= kPow10[exponent];
    } else if (exponent < -308) {
        if (exponent < -616) {
            number = 0.;
        } else {
            number /= kPow10[-308 - exponent];
            number /= kPow10[308];
        }
    } else {
        number /= kPow10[-exponent];
    }
    if (std::isinf(number)) return false;
    *out = number;
    return true;
}

inline bool ieq(const char* p, size_t n, const char* lit) {
    if (strlen(lit) != n) return false;
    for (size_t i = 0; i < n; ++i) {
        char c = p[i];
        if (c >= 'A' && c <= 'Z') c = (char)(c - 'A' + 'a');
        if (c != lit[i]) return false;
    }
    return true;
}

// A label pandas keeps as the string it is: not empty, not a number, boolean or NA spelling, no blanks at the ends.
bool plain_label(const char* p, size_t n) {
    if (n == 0) return false;
    const char c = p[0];
    if (is_digit(c) || c == '+' || c == '-' || c == '.' || c == '#' || c == '<' || c == ' ' || c == '\t') return false;
    if (p[n - 1] == ' ' || p[n - 1] == '\t') return false;
    static const char* const reserved[] = {"n/a", "na", "null", "nan", "none", "true", "false", "inf", "infinity"};
    for (const char* r : reserved)
        if (ieq(p, n, r)) return false;
    return true;
}

struct Line {
    const char* begin;
    const char* end;  // without the line terminator
};

}  // namespace

extern "C" int skr_csv_read(const char* path, int threads, SkrCsvTable** out) {
    if (!path || !out) return skr::fail(SKR_ERR_ARG, "skr_csv_read: null argument");
    *out = nullptr;
    FILE* f = fopen(path, "rb");
    if (!f) return skr::fail(SKR_ERR_IO, "skr_csv_read: cannot open %s", path);
    std::string text;
    {
        if (fseek(f, 0, SEEK_END) != 0) { fclose(f); return skr::fail(SKR_ERR_IO, "skr_csv_read: cannot seek in %s", path); }
        const long long size = ftell(f);
        rewind(f);
        text.resize(size > 0 ? (size_t)size : 0);
        const size_t got = text.empty() ? 0 : fread(&text[0], 1, text.size(), f);
        fclose(f);
        if (got != text.size()) return skr::fail(SKR_ERR_IO, "skr_csv_read: short read from %s", path);
    }
    if (threads < 1) threads = (int)std::thread::hardware_concurrency();
    if (threads < 1) threads = 1;
    if (text.empty() || memchr(text.data(), '"', text.size()) || memchr(text.data(), '\0', text.size()))
        return skr::fail(SKR_CSV_UNSUPPORTED, "skr_csv_read: empty file or quoted fields");

    // lines (blank lines are skipped as pandas does); a lone '\r' inside a line is left to pandas
    std::vector<Line> lines;
    {
        const char* p = text.data();
        const char* const end = p + text.size();
        while (p < end) {
            const char* nl = (const char*)memchr(p, '\n', (size_t)(end - p));
            const char* stop = nl ? nl : end;
            const char* e = stop;
            if (e > p && e[-1] == '\r') --e;
            if (memchr(p, '\r', (size_t)(e - p))) return skr::fail(SKR_CSV_UNSUPPORTED, "skr_csv_read: bare carriage return");
            if (e > p) lines.push_back({p, e});
            p = nl ? nl + 1 : end;
        }
    }
    if (lines.empty()) return skr::fail(SKR_CSV_UNSUPPORTED, "skr_csv_read: no header line");

    auto table = new SkrCsvTable();
    // header: index name, then the column names
    {
        const char* p = lines[0].begin;
        const char* const e = lines[0].end;
        const char* comma = (const char*)memchr(p, ',', (size_t)(e - p));
        table->column_offs.push_back(0);
        int64_t cols = 0;
        while (comma) {
            p = comma + 1;
            comma = (const char*)memchr(p, ',', (size_t)(e - p));
            const char* fe = comma ? comma : e;
            table->columns.append(p, (size_t)(fe - p));
            table->column_offs.push_back((int64_t)table->columns.size());
            ++cols;
        }
        table->cols = cols;
    }
    const int64_t rows = (int64_t)lines.size() - 1, cols = table->cols;
    table->rows = rows;
    if (cols == 0) { delete table; return skr::fail(SKR_CSV_UNSUPPORTED, "skr_csv_read: no data columns"); }
    table->values = (double*)malloc(sizeof(double) * (size_t)(rows > 0 ? rows : 1) * (size_t)cols);
    if (!table->values) { delete table; return skr::fail(SKR_ERR_NOMEM, "skr_csv_read: out of memory"); }

    std::vector<int> bad((size_t)threads, 0), integer((size_t)threads, 1);
    std::vector<Line> label_spans((size_t)rows);
    auto work = [&](int t) {
        const int64_t r0 = rows * t / threads, r1 = rows * (t + 1) / threads;
        const double nan = std::numeric_limits<double>::quiet_NaN(), inf = std::numeric_limits<double>::infinity();
        for (int64_t r = r0; r < r1; ++r) {
            const char* p = lines[(size_t)r + 1].begin;
            const char* const e = lines[(size_t)r + 1].end;
            const char* comma = (const char*)memchr(p, ',', (size_t)(e - p));
            if (!comma || !plain_label(p, (size_t)(comma - p))) { bad[(size_t)t] = 1; return; }
            label_spans[(size_t)r] = {p, comma};
            double* row = table->values + r * cols;
            p = comma + 1;
            for (int64_t c = 0; c < cols; ++c) {
                const char* fe = (c + 1 < cols) ? (const char*)memchr(p, ',', (size_t)(e - p)) : e;
                if (!fe || (c + 1 == cols && memchr(p, ',', (size_t)(e - p)))) { bad[(size_t)t] = 1; return; }
                bool is_int = false;
                if (fe == p) {
                    row[c] = nan;  // empty cell: NaN (what to_csv writes for NaN)
                    is_int = false;
                } else if (!parse_cell(p, fe, &row[c], &is_int)) {
                    const size_t n = (size_t)(fe - p);
                    if (n == 3 && memcmp(p, "inf", 3) == 0) row[c] = inf;
                    else if (n == 4 && memcmp(p, "-inf", 4) == 0) row[c] = -inf;
                    else { bad[(size_t)t] = 1; return; }
                } else if (is_int && (fe - p) > 15) {
                    bad[(size_t)t] = 1;  // pandas parses integer columns exactly as int64: keep to the range where both agree
                    return;
                }
                if (!is_int) integer[(size_t)t] = 0;
                p = fe + 1;
            }
        }
    };
    {
        std::vector<std::thread> pool;
        for (int t = 1; t < threads; ++t) pool.emplace_back(work, t);
        work(0);
        for (auto& th : pool) th.join();
    }
    for (int t = 0; t < threads; ++t)
        if (bad[(size_t)t]) {
            free(table->values);
            delete table;
            return skr::fail(SKR_CSV_UNSUPPORTED, "skr_csv_read: a row of %s is not in the plain labelled form", path);
        }
    table->all_integer = 1;
    for (int t = 0; t < threads; ++t) table->all_integer &= integer[(size_t)t];
    table->label_offs.reserve((size_t)rows + 1);
    table->label_offs.push_back(0);
    for (int64_t r = 0; r < rows; ++r) {
        table->labels.append(label_spans[(size_t)r].begin, (size_t)(label_spans[(size_t)r].end - label_spans[(size_t)r].begin));
        table->label_offs.push_back((int64_t)table->labels.size());
    }
    *out = table;
    return SKR_OK;
}

extern "C" void skr_csv_free(SkrCsvTable* t) {
    if (!t) return;
    free(t->values);
    delete t;
}
extern "C" int64_t skr_csv_rows(const SkrCsvTable* t) { return t ? t->rows : 0; }
extern "C" int64_t skr_csv_cols(const SkrCsvTable* t) { return t ? t->cols : 0; }
extern "C" const double* skr_csv_values(const SkrCsvTable* t) { return t ? t->values : nullptr; }
extern "C" const char* skr_csv_labels(const SkrCsvTable* t) { return t ? t->labels.data() : nullptr; }
extern "C" const int64_t* skr_csv_label_offsets(const SkrCsvTable* t) { return t ? t->label_offs.data() : nullptr; }
extern "C" const char* skr_csv_columns(const SkrCsvTable* t) { return t ? t->columns.data() : nullptr; }
extern "C" const int64_t* skr_csv_column_offsets(const SkrCsvTable* t) { return t ? t->column_offs.data() : nullptr; }
extern "C" int skr_csv_all_integer(const SkrCsvTable* t) { return t ? t->all_integer : 0; }
