// Host FASTA ingest: text -> 2-bit codes + invalid mask in (pinned) host memory.
//
// Replaces seekr/fasta_reader.py:41-78 (Reader._read_data, _upper_seq_per_line, get_seqs,
// get_headers) for the counting path.  Semantics kept from the reference:
//   * text-mode line splitting: "\n", "\r\n" and a lone "\r" all end a line;
//   * every line is stripped of leading/trailing whitespace (str.strip: 0x09-0x0D, 0x1C-0x20);
//   * a stripped line starting with '>' is a header, everything else is sequence, joined per record
//     and upper-cased; every remaining byte counts as one base (inner blanks included);
//   * a blank line raises IndexError there (line[0] on an empty string)   -> SKR_ERR_FASTA_BLANK;
//   * a header directly after a header (except at line 0) trips the assert -> SKR_ERR_FASTA_HEADER;
//   * the last record may be empty.
// The packed layout is described in include/seekr_b200.h.
#include <cuda_runtime.h>
#include <immintrin.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <memory>
#include <mutex>
#include <thread>
#include <unordered_map>
#include <vector>

#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include "skr_common.h"

namespace skr {

std::string& last_error() {
    thread_local std::string s;
    return s;
}
int64_t& launch_counter() {
    thread_local int64_t n = 0;
    return n;
}
int fail(int code, const char* fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    last_error() = buf;
    return code;
}

}  // namespace skr

extern "C" const char* skr_last_error(void) { return skr::last_error().c_str(); }
extern "C" int skr_abi_version(void) { return SKR_ABI_VERSION; }
extern "C" int64_t skr_launch_count(int reset) {
    int64_t v = skr::launch_counter();
    if (reset) skr::launch_counter() = 0;
    return v;
}

namespace {

thread_local int64_t g_error_line = 0;

// ---------------------------------------------------------------------------------------------
// slab allocation (pinned slabs are pooled: cudaHostAlloc costs milliseconds per call)
// ---------------------------------------------------------------------------------------------
struct Slab {
    void* ptr = nullptr;
    size_t cap = 0;
    bool pinned = false;
};

std::mutex g_pool_mu;
std::vector<Slab> g_pool;
constexpr size_t kPoolMaxSlabs = 8;

int slab_alloc(size_t bytes, bool pinned, Slab* out) {
    bytes = std::max<size_t>(bytes, 4096);
    if (pinned) {
        std::lock_guard<std::mutex> lock(g_pool_mu);
        int best = -1;
        for (size_t i = 0; i < g_pool.size(); ++i)
            if (g_pool[i].cap >= bytes && (best < 0 || g_pool[i].cap < g_pool[best].cap)) best = (int)i;
        if (best >= 0) {
            *out = g_pool[best];
            g_pool.erase(g_pool.begin() + best);
            return SKR_OK;
        }
    }
    out->pinned = pinned;
    if (pinned) {
        size_t cap = bytes + bytes / 8;  // head-room so a slightly larger next input reuses the slab
        cudaError_t e = cudaHostAlloc(&out->ptr, cap, cudaHostAllocDefault);
        if (e != cudaSuccess)
            return skr::fail(SKR_ERR_CUDA, "cudaHostAlloc(%zu) failed: %s", cap, cudaGetErrorString(e));
        out->cap = cap;
    } else {
        if (posix_memalign(&out->ptr, 4096, bytes) != 0) return skr::fail(SKR_ERR_NOMEM, "out of host memory");
        out->cap = bytes;
    }
    return SKR_OK;
}

void slab_free(Slab& s) {
    if (!s.ptr) return;
    if (s.pinned) {
        std::lock_guard<std::mutex> lock(g_pool_mu);
        if (g_pool.size() < kPoolMaxSlabs) {
            g_pool.push_back(s);
        } else {
            // drop the smallest pooled slab in favour of this one if it is larger
            size_t small = 0;
            for (size_t i = 1; i < g_pool.size(); ++i)
                if (g_pool[i].cap < g_pool[small].cap) small = i;
            if (g_pool[small].cap < s.cap) {
                cudaFreeHost(g_pool[small].ptr);
                g_pool[small] = s;
            } else {
                cudaFreeHost(s.ptr);
            }
        }
    } else {
        free(s.ptr);
    }
    s.ptr = nullptr;
}

inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

std::mutex g_live_mu;
std::unordered_map<void*, Slab> g_live;  // slabs handed out through skr_host_alloc

}  // namespace

extern "C" int skr_host_alloc(size_t bytes, void** out) {
    if (!out) return skr::fail(SKR_ERR_ARG, "skr_host_alloc: null out");
    Slab s;
    int rc = slab_alloc(bytes, true, &s);
    if (rc != SKR_OK) return rc;
    {
        std::lock_guard<std::mutex> lock(g_live_mu);
        g_live[s.ptr] = s;
    }
    *out = s.ptr;
    return SKR_OK;
}

extern "C" int skr_host_alloc_pooled(size_t bytes, void** out) {
    if (!out) return skr::fail(SKR_ERR_ARG, "skr_host_alloc_pooled: null out");
    *out = nullptr;
    Slab s;
    {
        std::lock_guard<std::mutex> lock(g_pool_mu);
        int best = -1;
        for (size_t i = 0; i < g_pool.size(); ++i)
            if (g_pool[i].cap >= bytes && (best < 0 || g_pool[i].cap < g_pool[best].cap)) best = (int)i;
        if (best < 0) return SKR_OK;
        s = g_pool[best];
        g_pool.erase(g_pool.begin() + best);
    }
    {
        std::lock_guard<std::mutex> lock(g_live_mu);
        g_live[s.ptr] = s;
    }
    *out = s.ptr;
    return SKR_OK;
}

extern "C" void skr_host_free(void* p) {
    if (!p) return;
    Slab s;
    {
        std::lock_guard<std::mutex> lock(g_live_mu);
        auto it = g_live.find(p);
        if (it == g_live.end()) return;
        s = it->second;
        g_live.erase(it);
    }
    slab_free(s);
}

extern "C" void skr_host_pool_trim(void) {
    std::lock_guard<std::mutex> lock(g_pool_mu);
    for (auto& s : g_pool) cudaFreeHost(s.ptr);
    g_pool.clear();
}

extern "C" int skr_copy_h2d(void* d_dst, const void* h_src, size_t bytes, void* stream) {
    if (!bytes) return SKR_OK;
    SKR_CUDA_CHECK(cudaMemcpyAsync(d_dst, h_src, bytes, cudaMemcpyHostToDevice, (cudaStream_t)stream));
    return SKR_OK;
}
extern "C" int skr_copy_d2h(void* h_dst, const void* d_src, size_t bytes, void* stream) {
    if (!bytes) return SKR_OK;
    SKR_CUDA_CHECK(cudaMemcpyAsync(h_dst, d_src, bytes, cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    return SKR_OK;
}
extern "C" int skr_copy_d2h_2d(void* h_dst, size_t h_pitch, const void* d_src, size_t d_pitch, size_t row_bytes,
                               size_t rows, void* stream) {
    if (!rows || !row_bytes) return SKR_OK;
    SKR_CUDA_CHECK(cudaMemcpy2DAsync(h_dst, h_pitch, d_src, d_pitch, row_bytes, rows, cudaMemcpyDeviceToHost,
                                     (cudaStream_t)stream));
    return SKR_OK;
}
extern "C" int skr_copy_h2d_2d(void* d_dst, size_t d_pitch, const void* h_src, size_t h_pitch, size_t row_bytes,
                               size_t rows, void* stream) {
    if (!rows || !row_bytes) return SKR_OK;
    SKR_CUDA_CHECK(cudaMemcpy2DAsync(d_dst, d_pitch, h_src, h_pitch, row_bytes, rows, cudaMemcpyHostToDevice,
                                     (cudaStream_t)stream));
    return SKR_OK;
}
extern "C" int skr_stream_sync(void* stream) {
    SKR_CUDA_CHECK(cudaStreamSynchronize((cudaStream_t)stream));
    return SKR_OK;
}
extern "C" int skr_device_count(int* out) {
    if (!out) return skr::fail(SKR_ERR_ARG, "null out");
    SKR_CUDA_CHECK(cudaGetDeviceCount(out));
    return SKR_OK;
}

struct PackJob;
struct WaveJob;

struct SkrPacked {
    PackJob* job = nullptr;  // packing still running in the background (skr_pack_fasta_buffer_async)
    WaveJob* wave = nullptr; // scan AND pack still running in the background, wave by wave (large texts)
    Slab retired;            // wave mode, capacity estimate too small: the first slab, kept until the handle goes
    int64_t m = 0;
    int64_t nblocks = 0;  // including the trailing pad block
    int64_t total_bases = 0;
    Slab slab;
    size_t slab_bytes = 0;
    uint32_t* codes = nullptr;
    uint32_t* mask = nullptr;
    uint64_t* blk_off = nullptr;
    uint32_t* len = nullptr;
    std::vector<uint64_t> header_spans;
    std::vector<uint64_t> body_spans;
};

namespace {

inline bool is_space(unsigned char c) { return (c >= 0x09 && c <= 0x0D) || (c >= 0x1C && c <= 0x20); }

// Calls fn(line_begin, line_end) for every line that STARTS in [from, to); a line may extend past
// `to` (never past `end`).  Line breaks: \n, \r\n, \r (text-mode "universal newlines").  fn returns
// false to stop.  Returns the start of the first line not visited, or nullptr when stopped.
template <class Fn>
const char* for_each_line(const char* base, const char* from, const char* to, const char* end, Fn&& fn) {
    const char* s = from;
    if (s > base) {  // move to the first line start at or after `from`
        while (s < end) {
            char prev = s[-1];
            if (prev == '\n' || (prev == '\r' && *s != '\n')) break;
            ++s;
        }
    }
    const char* nl = nullptr;  // cached position of the next '\n' at or after s (end if none)
    while (s < to && s < end) {
        if (!nl || nl < s) {
            nl = (const char*)memchr(s, '\n', (size_t)(end - s));
            if (!nl) nl = end;
        }
        const char* e = nl;
        const char* nxt = nl < end ? nl + 1 : end;
        const char* cr = (const char*)memchr(s, '\r', (size_t)(nl - s));
        if (cr && !(cr + 1 == nl && nl < end)) {  // a lone \r ends the line (a \r\n pair is left to strip())
            e = cr;
            nxt = cr + 1;
        }
        if (!fn(s, e)) return nullptr;
        s = nxt;
    }
    return s;
}

// AVX2 variant of for_each_line: 32 bytes per step, every '\n' / '\r' found through one movemask.
#define SKR_AVX2 __attribute__((target("avx2,bmi,bmi2,lzcnt,popcnt")))

template <class Fn>
SKR_AVX2 const char* for_each_line_avx2(const char* base, const char* from, const char* to, const char* end, Fn&& fn) {
    const char* s = from;
    if (s > base) {
        while (s < end) {
            char prev = s[-1];
            if (prev == '\n' || (prev == '\r' && *s != '\n')) break;
            ++s;
        }
    }
    if (s >= to || s >= end) return s;
    const __m256i v_nl = _mm256_set1_epi8('\n'), v_cr = _mm256_set1_epi8('\r');
    const char* p = s;          // scan position
    const char* skip = nullptr; // the '\n' of a "\r\n" pair already consumed
    while (p < end) {
        uint32_t mask;
        size_t n = (size_t)(end - p);
        if (n >= 32) {
            const __m256i v = _mm256_loadu_si256((const __m256i*)p);
            mask = (uint32_t)_mm256_movemask_epi8(_mm256_or_si256(_mm256_cmpeq_epi8(v, v_nl), _mm256_cmpeq_epi8(v, v_cr)));
            n = 32;
        } else {
            mask = 0;
            for (size_t i = 0; i < n; ++i)
                if (p[i] == '\n' || p[i] == '\r') mask |= 1u << i;
        }
        while (mask) {
            const char* t = p + __builtin_ctz(mask);
            mask &= mask - 1;
            if (t == skip) continue;
            const char* nxt = t + 1;
            if (*t == '\r' && nxt < end && *nxt == '\n') { skip = nxt; nxt = t + 2; }
            if (!fn(s, t)) return nullptr;
            s = nxt;
            if (s >= to) return s;
        }
        p += n;
    }
    if (s < end && s < to) {  // last line without a terminator
        if (!fn(s, end)) return nullptr;
        s = end;
    }
    return s;
}

// Line iteration with a fast lane for the bulk of a FASTA file: runs of "clean" sequence lines -- no '>' , no
// whitespace other than the terminating '\n' (so nothing to strip, no '\r'), no blank line -- are consumed 32 bytes
// at a time with three compares and a popcount, and reported once per run through clean(first_line_start,
// bases, end_of_last_line); every other line goes through fn exactly as in for_each_line_avx2.  The effect on the
// caller's state is the same as visiting the lines one by one (the per-line callback was the whole cost of the
// scan: ~25 ns per 61-byte line).
template <class Fn, class Clean>
SKR_AVX2 const char* for_each_line_fast(const char* base, const char* from, const char* to, const char* end, Fn&& fn,
                                        Clean&& clean) {
    const char* s = from;
    if (s > base) {
        while (s < end) {
            char prev = s[-1];
            if (prev == '\n' || (prev == '\r' && *s != '\n')) break;
            ++s;
        }
    }
    const __m256i v_nl = _mm256_set1_epi8('\n'), v_gt = _mm256_set1_epi8('>');
    const __m256i v_sp = _mm256_set1_epi8(0x20);
    const char* lim = to < end ? to : end;
    while (s < to && s < end) {
        const char* p = s;
        const char* last_nl = nullptr;
        uint64_t nls = 0;
        uint32_t carry = 1;  // s is a line start: a '\n' right here is a blank line
        while (p + 64 <= lim) {  // two vectors per step, no data-dependent branch but the exit
            const __m256i v0 = _mm256_loadu_si256((const __m256i*)p);
            const __m256i v1 = _mm256_loadu_si256((const __m256i*)(p + 32));
            const uint64_t nlm = (uint64_t)(uint32_t)_mm256_movemask_epi8(_mm256_cmpeq_epi8(v0, v_nl)) |
                                 ((uint64_t)(uint32_t)_mm256_movemask_epi8(_mm256_cmpeq_epi8(v1, v_nl)) << 32);
            const __m256i s0 = _mm256_or_si256(_mm256_cmpeq_epi8(_mm256_min_epu8(v0, v_sp), v0), _mm256_cmpeq_epi8(v0, v_gt));
            const __m256i s1 = _mm256_or_si256(_mm256_cmpeq_epi8(_mm256_min_epu8(v1, v_sp), v1), _mm256_cmpeq_epi8(v1, v_gt));
            const uint64_t spm = ((uint64_t)(uint32_t)_mm256_movemask_epi8(s0) | ((uint64_t)(uint32_t)_mm256_movemask_epi8(s1) << 32)) & ~nlm;
            const uint64_t blank = nlm & ((nlm << 1) | carry);
            if (spm | blank) break;
            nls += (uint64_t)__builtin_popcountll(nlm);
            const char* cand = p + (63 - (int)_lzcnt_u64(nlm));  // p - 1 when the step holds no '\n': never selected
            last_nl = nlm ? cand : last_nl;
            carry = (uint32_t)(nlm >> 63);
            p += 64;
        }
        while (p + 32 <= lim) {
            const __m256i v = _mm256_loadu_si256((const __m256i*)p);
            const uint32_t nlm = (uint32_t)_mm256_movemask_epi8(_mm256_cmpeq_epi8(v, v_nl));
            // every byte <= 0x20 (all of str.strip's whitespace, and the other control characters for good measure)
            const __m256i ws = _mm256_cmpeq_epi8(_mm256_min_epu8(v, v_sp), v);
            const uint32_t spm = (uint32_t)_mm256_movemask_epi8(_mm256_or_si256(ws, _mm256_cmpeq_epi8(v, v_gt))) & ~nlm;
            const uint32_t blank = nlm & ((nlm << 1) | carry);
            if (spm | blank) break;
            if (nlm) {
                nls += (uint64_t)__builtin_popcount(nlm);
                last_nl = p + (31 - __builtin_clz(nlm));
            }
            carry = nlm >> 31;
            p += 32;
        }
        if (last_nl) {  // the complete clean lines s .. last_nl
            if (!clean(s, (uint64_t)(last_nl - s) - (nls - 1), last_nl)) return nullptr;
            s = last_nl + 1;
            continue;
        }
        const char* nxt = for_each_line_avx2(base, s, s + 1, end, fn);  // exactly one line, the careful way
        if (!nxt) return nullptr;
        s = nxt;
    }
    return s;
}

inline void strip(const char*& a, const char*& b) {
    while (a < b && is_space((unsigned char)*a)) ++a;
    while (b > a && is_space((unsigned char)b[-1])) --b;
}

struct Rec {
    uint64_t hdr_off, hdr_len;
    uint64_t body_off, body_len;  // first sequence line start .. last sequence line end (unstripped span)
    uint64_t bases;
};

struct ChunkResult {
    std::vector<Rec> recs;
    uint64_t err_off = UINT64_MAX;
    int err_code = 0;
    bool first_line_not_header = false;
};

void build_lut2(const uint8_t* lut, uint8_t* lut2) {
    for (int c = 0; c < 256; ++c) {
        int u = (c >= 'a' && c <= 'z') ? c - 32 : c;
        lut2[c] = lut[u];
    }
}

// Packs stripped line segments of one record.
struct BitWriter {
    uint32_t* cw;
    uint32_t* mw;
    // pending bases, top aligned: cacc holds cn < 16 bases (2 bits each), macc holds mn < 32 mask bits
    uint64_t cacc = 0, macc = 0;
    int cn = 0, mn = 0;
    inline void put(uint8_t d) {
        const uint64_t inv = d > 3;
        cacc |= (uint64_t)(inv ? 0u : d) << (62 - 2 * cn);
        macc |= inv << (63 - mn);
        if (++cn == 16) { *cw++ = (uint32_t)(cacc >> 32); cacc = 0; cn = 0; }
        if (++mn == 32) { *mw++ = (uint32_t)(macc >> 32); macc = 0; mn = 0; }
    }
    // n <= 32 bases at once: codes top aligned in c64 (2n bits), invalid flags top aligned in i32 (n bits)
    inline void append(uint64_t c64, uint32_t i32, int n) {
        // pending cn < 16 bases (< 32 bits) + up to 64 new bits: hi takes what fits, lo the overflow
        const int sh = 2 * cn;
        uint64_t hi = cacc | (c64 >> sh);
        uint64_t lo = sh ? (c64 << (64 - sh)) : 0;
        int total = cn + n;
        if (total >= 16) {
            *cw++ = (uint32_t)(hi >> 32);
            if (total >= 32) {
                *cw++ = (uint32_t)hi;
                hi = lo;
                total -= 32;
                if (total >= 16) { *cw++ = (uint32_t)(hi >> 32); hi <<= 32; total -= 16; }
            } else {
                hi = (hi << 32) | (lo >> 32);
                total -= 16;
            }
        }
        cacc = hi;
        cn = total;
        macc |= ((uint64_t)i32 << 32) >> mn;
        mn += n;
        if (mn >= 32) {
            *mw++ = (uint32_t)(macc >> 32);
            macc <<= 32;
            mn -= 32;
        }
    }
    // pad to the end of the record's last block: codes 0, mask 1
    void finish(uint32_t* cend, uint32_t* mend) {
        if (cn) { *cw++ = (uint32_t)(cacc >> 32); cn = 0; cacc = 0; }
        if (mn) { *mw++ = (uint32_t)(macc >> 32) | (0xFFFFFFFFu >> mn); mn = 0; macc = 0; }
        while (cw < cend) *cw++ = 0;
        while (mw < mend) *mw++ = 0xFFFFFFFFu;
    }
};

// The four alphabet letters when they are plain upper-case ASCII letters (then a byte matches letter i
// iff (byte | 0x20) == (letter | 0x20), which also upper-cases the input); otherwise the LUT path is used.
struct SimdAlphabet {
    bool ok = false;
    uint8_t low[4] = {0, 0, 0, 0};  // letter | 0x20 for digit 0..3 (0 = digit unused)
};

SimdAlphabet simd_alphabet(const uint8_t* lut) {
    SimdAlphabet a;
    int found = 0;
    for (int c = 0; c < 256; ++c) {
        if (lut[c] <= 3) {
            if (c < 'A' || c > 'Z' || a.low[lut[c]] != 0) return a;  // not a letter / digit used twice
            a.low[lut[c]] = (uint8_t)(c | 0x20);
            ++found;
        }
    }
    a.ok = found > 0 && __builtin_cpu_supports("avx2") && __builtin_cpu_supports("bmi2") && !getenv("SKR_PACK_NO_AVX2");
    return a;
}

SKR_AVX2 void pack_segment_avx2(BitWriter& w, const char* a, const char* b, const SimdAlphabet& al, const char* buf_end) {
    const __m256i rev = _mm256_setr_epi8(15, 14, 13, 12, 11, 10, 9, 8, 7, 6, 5, 4, 3, 2, 1, 0,
                                         15, 14, 13, 12, 11, 10, 9, 8, 7, 6, 5, 4, 3, 2, 1, 0);
    const __m256i lower = _mm256_set1_epi8(0x20);
    const __m256i l0 = _mm256_set1_epi8((char)al.low[0]), l1 = _mm256_set1_epi8((char)al.low[1]);
    const __m256i l2 = _mm256_set1_epi8((char)al.low[2]), l3 = _mm256_set1_epi8((char)al.low[3]);
    alignas(32) char tail[32];
    while (a < b) {
        int n = (int)std::min<ptrdiff_t>(32, b - a);
        __m256i v;
        if (a + 32 <= buf_end) {  // bytes past b are masked off below
            v = _mm256_loadu_si256((const __m256i*)a);
        } else {
            memset(tail, 0, 32);
            memcpy(tail, a, (size_t)n);
            v = _mm256_load_si256((const __m256i*)tail);
        }
        // byte i -> byte 31-i
        v = _mm256_shuffle_epi8(v, rev);
        v = _mm256_permute2x128_si256(v, v, 0x01);
        const __m256i vl = _mm256_or_si256(v, lower);
        const __m256i m0 = _mm256_cmpeq_epi8(vl, l0), m1 = _mm256_cmpeq_epi8(vl, l1);
        const __m256i m2 = _mm256_cmpeq_epi8(vl, l2), m3 = _mm256_cmpeq_epi8(vl, l3);
        const uint32_t keep = n == 32 ? 0xFFFFFFFFu : ~(0xFFFFFFFFu >> n);  // top n bits
        const uint32_t b0 = (uint32_t)_mm256_movemask_epi8(_mm256_or_si256(m1, m3)) & keep;
        const uint32_t b1 = (uint32_t)_mm256_movemask_epi8(_mm256_or_si256(m2, m3)) & keep;
        const uint32_t valid = (uint32_t)_mm256_movemask_epi8(_mm256_or_si256(_mm256_or_si256(m0, m1), _mm256_or_si256(m2, m3)));
        const uint32_t inv = ~valid & keep;
        const uint64_t c64 = _pdep_u64(b1, 0xAAAAAAAAAAAAAAAAull) | _pdep_u64(b0, 0x5555555555555555ull);
        w.append(c64, inv, n);
        a += n;
    }
}


// Pass 1 over one byte range of the text: the records whose header line STARTS in [from, to), each followed to its
// end (past `to` if need be).  Sequence lines at the start of the range belong to a record of an earlier range.
struct ScanCtx {
    const char* text;
    const char* end;
    bool use_avx2;
    bool fast;
};

void scan_slice(const ScanCtx& cx, const char* from, const char* to, bool first_slice, ChunkResult& R) {
    const char* text = cx.text;
    const char* end = cx.end;
    bool in_record = false;  // a header that started in this range is open
    bool beyond = false;     // phase 2: lines that start past `to` (they finish our last record)
    auto on_line = [&](const char* a, const char* b) -> bool {
        const char* la = a;
        const char* lb = b;
        strip(la, lb);
        if (la == lb) {
            if (beyond) return false;  // the range that owns this line reports it
            uint64_t off = (uint64_t)(a - text);
            if (off < R.err_off) { R.err_off = off; R.err_code = SKR_ERR_FASTA_BLANK; }
            return false;
        }
        if (*la == '>') {
            if (beyond) return false;  // next range's record: our last record is complete
            Rec r;
            r.hdr_off = (uint64_t)(la - text);
            r.hdr_len = (uint64_t)(lb - la);
            r.body_off = (uint64_t)(b - text);
            r.body_len = 0;
            r.bases = 0;
            R.recs.push_back(r);
            in_record = true;
            return true;
        }
        if (!in_record) {  // sequence line of a record opened in an earlier range
            if (first_slice) { R.first_line_not_header = true; return false; }
            return true;
        }
        Rec& r = R.recs.back();
        if (r.body_len == 0) r.body_off = (uint64_t)(a - text);
        r.body_len = (uint64_t)(b - text) - r.body_off;
        r.bases += (uint64_t)(lb - la);
        return true;
    };
    auto on_clean = [&](const char* a, uint64_t nbases, const char* last_end) -> bool {
        if (!in_record) {  // sequence lines of a record opened in an earlier range
            if (first_slice) { R.first_line_not_header = true; return false; }
            return true;
        }
        Rec& r = R.recs.back();
        if (r.body_len == 0) r.body_off = (uint64_t)(a - text);
        r.body_len = (uint64_t)(last_end - text) - r.body_off;
        r.bases += nbases;
        return true;
    };
    const char* next = cx.fast ? for_each_line_fast(text, from, to, end, on_line, on_clean)
                               : (cx.use_avx2 ? for_each_line_avx2(text, from, to, end, on_line)
                                              : for_each_line(text, from, to, end, on_line));
    if (next && in_record && next < end) {
        beyond = true;
        if (cx.fast) for_each_line_fast(text, next, end, end, on_line, on_clean);
        else if (cx.use_avx2) for_each_line_avx2(text, next, end, end, on_line);
        else for_each_line(text, next, end, end, on_line);
    }
}

// 1-based number of the line that holds byte `off`
int64_t line_of_offset(const char* text, uint64_t off) {
    int64_t line = 1;
    for (const char* p = text; p < text + off; ++p)
        if (*p == '\n' || (*p == '\r' && p[1] != '\n')) ++line;
    return line;
}

int run_threads(int nthreads, const std::function<void(int)>& fn) {
    if (nthreads <= 1) {
        fn(0);
        return 0;
    }
    std::vector<std::thread> th;
    th.reserve(nthreads - 1);
    for (int t = 1; t < nthreads; ++t) th.emplace_back(fn, t);
    fn(0);
    for (auto& t : th) t.join();
    return 0;
}

int pick_threads(int nthreads, size_t work_bytes) {
    if (nthreads > 0) return std::min(nthreads, 256);  // explicit request: honoured as is
    nthreads = (int)std::thread::hardware_concurrency();
    if (nthreads <= 0) nthreads = 1;
    nthreads = std::min(nthreads, 64);
    size_t by_work = std::max<size_t>(1, work_bytes / (256 * 1024));
    return (int)std::min<size_t>((size_t)nthreads, by_work);
}

int alloc_packed(SkrPacked* P, int64_t m, const std::vector<uint64_t>& bases, bool pinned) {
    P->m = m;
    uint64_t blocks = 0;
    for (int64_t i = 0; i < m; ++i) blocks += (bases[i] + 63) / 64;
    P->nblocks = (int64_t)blocks + 1;
    size_t off_codes = 0;
    size_t off_mask = align_up(off_codes + (size_t)P->nblocks * 16, 256);
    size_t off_blk = align_up(off_mask + (size_t)P->nblocks * 8, 256);
    size_t off_len = align_up(off_blk + (size_t)(m + 1) * 8, 256);
    size_t total = align_up(off_len + (size_t)std::max<int64_t>(m, 1) * 4, 256);
    int rc = slab_alloc(total, pinned, &P->slab);
    if (rc != SKR_OK) return rc;
    P->slab_bytes = total;
    char* base = (char*)P->slab.ptr;
    P->codes = (uint32_t*)(base + off_codes);
    P->mask = (uint32_t*)(base + off_mask);
    P->blk_off = (uint64_t*)(base + off_blk);
    P->len = (uint32_t*)(base + off_len);
    uint64_t b = 0;
    uint64_t tot = 0;
    for (int64_t i = 0; i < m; ++i) {
        P->blk_off[i] = b;
        P->len[i] = (uint32_t)bases[i];
        b += (bases[i] + 63) / 64;
        tot += bases[i];
    }
    P->blk_off[m] = b;
    P->total_bases = (int64_t)tot;
    // trailing pad block: codes 0, mask all ones
    memset(P->codes + b * 4, 0, 16);
    memset(P->mask + b * 2, 0xFF, 8);
    return SKR_OK;
}


}  // namespace

namespace {

// Pass 2 for one record: its sequence lines, stripped, into the 2-bit codes and the validity mask of its blocks.
void pack_record(SkrPacked* P, const Rec& r, int64_t i, const char* text, const char* end, const SimdAlphabet& al,
                 const uint8_t* lut2, bool use_avx2) {
    uint64_t b0 = P->blk_off[i], b1 = P->blk_off[i + 1];
    BitWriter w{P->codes + b0 * 4, P->mask + b0 * 2};
    const char* bs = text + r.body_off;
    const char* be = bs + r.body_len;
    if (r.body_len) {
        auto pack_line = [&](const char* a, const char* b) -> bool {
            strip(a, b);
            if (al.ok) pack_segment_avx2(w, a, b, al, end);
            else for (const char* p = a; p < b; ++p) w.put(lut2[(unsigned char)*p]);
            return true;
        };
        if (use_avx2) for_each_line_avx2(bs, bs, be, be, pack_line);
        else for_each_line(bs, bs, be, be, pack_line);
    }
    w.finish(P->codes + b1 * 4, P->mask + b1 * 2);
}

}  // namespace

// The pack pass over the records found by the scan: groups of 64 records are drawn in record order by however many
// threads run work(); done[g] is raised when group g is complete, so a consumer can follow the packed prefix.
struct PackJob {
    SkrPacked* P;
    std::vector<Rec> recs;
    const char* text;
    const char* end;
    SimdAlphabet al;
    uint8_t lut2[256];
    bool use_avx2;
    int64_t m;
    int64_t ngroups;
    std::atomic<int64_t> next_rec{0};
    std::unique_ptr<std::atomic<uint8_t>[]> done;
    std::atomic<int64_t> watermark{0};  // groups [0, watermark) are known to be complete
    std::vector<std::thread> threads;

    static constexpr int64_t kGroup = 64;

    PackJob(SkrPacked* P_, std::vector<Rec>&& recs_, const char* text_, const char* end_, const SimdAlphabet& al_,
            const uint8_t* lut2_, bool use_avx2_)
        : P(P_), recs(std::move(recs_)), text(text_), end(end_), al(al_), use_avx2(use_avx2_) {
        memcpy(lut2, lut2_, 256);
        m = (int64_t)recs.size();
        ngroups = (m + kGroup - 1) / kGroup;
        done.reset(new std::atomic<uint8_t>[(size_t)std::max<int64_t>(ngroups, 1)]);
        for (int64_t g = 0; g < ngroups; ++g) done[g].store(0, std::memory_order_relaxed);
    }

    void work() {
        for (;;) {
            const int64_t i0 = next_rec.fetch_add(kGroup);
            if (i0 >= m) break;
            const int64_t i1 = std::min(m, i0 + kGroup);
            for (int64_t i = i0; i < i1; ++i) pack_record(P, recs[i], i, text, end, al, lut2, use_avx2);
            done[i0 / kGroup].store(1, std::memory_order_release);
        }
    }

    void start(int nthreads) {
        threads.reserve((size_t)nthreads);
        for (int t = 0; t < nthreads; ++t) threads.emplace_back([this] { work(); });
    }

    // blocks until records [0, upto) are packed
    void wait_records(int64_t upto) {
        const int64_t need = std::min(ngroups, (upto + kGroup - 1) / kGroup);
        int64_t w = watermark.load(std::memory_order_relaxed);
        int spins = 0;
        while (w < need) {
            if (done[w].load(std::memory_order_acquire)) { ++w; continue; }
            if (++spins < 2000) _mm_pause(); else std::this_thread::yield();
        }
        watermark.store(w, std::memory_order_relaxed);
    }

    void join() {
        for (auto& t : threads) t.join();
        threads.clear();
    }
};


// ---- scan and pack in waves (large texts, background mode) -----------------------------------------------------
// The serial part of the streamed get_counts() used to be the scan: the record count sizes every allocation, so
// nothing could start before the whole text had been looked at (2.5-3.3 ms for 165 MB on 16 threads).  Here the
// text is cut into slices of ~nbytes / (16 T); the T workers go through them T at a time: scan a slice each,
// barrier, worker 0 appends the wave's records to the table (prefix of block offsets, the checks of the reference's
// parser), barrier, everybody packs the records that are final, next wave.  The slab is allocated after the first
// wave for an ESTIMATED record count (records per byte of wave 0, + 25 % + 1024) and a block count that is a bound
// given that record count (bases <= bytes); the call returns at that point.  Consumers follow `scanned` (table
// entries final) and the per-group done flags (codes and mask written).  A text whose later part is denser in
// records than the estimate allows falls back: the waves finish the scan only, then an exact slab is packed as in
// the one-shot mode, and streaming consumers are told SKR_ERR_CAPACITY (they restart on the finished handle).
struct SpinBarrier {
    std::atomic<int> count{0};
    std::atomic<int> sense{0};
    int n = 1;
    void wait(int& local) {
        local ^= 1;
        if (count.fetch_add(1, std::memory_order_acq_rel) == n - 1) {
            count.store(0, std::memory_order_relaxed);
            sense.store(local, std::memory_order_release);
        } else {
            int spins = 0;
            while (sense.load(std::memory_order_acquire) != local) {
                if (++spins < 4000) _mm_pause(); else std::this_thread::yield();
            }
        }
    }
};

namespace {
int layout_packed(SkrPacked* P, int64_t cap_records, int64_t cap_blocks, bool pinned);
}

struct WaveJob {
    SkrPacked* P;
    ScanCtx cx;
    size_t nbytes;
    SimdAlphabet al;
    uint8_t lut2[256];
    bool pinned;
    int device = -1;  // the caller's CUDA device: worker 0 allocates the pinned slab, which must not wake device 0
    int T;
    size_t slice_bytes;
    int64_t nslices, nwaves;
    std::vector<ChunkResult> res;  // one per worker, reused every wave
    std::vector<Rec> recs;         // reserved to the capacity: addresses stay put while packers read them
    int64_t cap_records = 0;
    uint64_t blocks = 0, total_bases = 0;
    std::atomic<int64_t> scanned{0};      // records whose table entries are final
    std::atomic<int64_t> pack_groups{0};  // groups of 64 records that may be packed
    std::atomic<int64_t> next_group{0};
    std::atomic<int64_t> watermark{0};
    std::unique_ptr<std::atomic<uint8_t>[]> done;
    std::atomic<int> wave0{0};     // the handle is usable (slab allocated, first records in the table) or the job failed
    std::atomic<int> finished{0};  // the record table is complete (or the job failed)
    std::atomic<int> overflow{0};
    std::atomic<int> failed{0};
    int err_code = 0;
    int64_t err_line = 0;
    std::string err_msg;
    SpinBarrier bar;
    std::vector<std::thread> threads;
    static constexpr int64_t kGroup = 64;

    void fail(int code, int64_t line, const char* fmt, ...) {
        char buf[512];
        va_list ap;
        va_start(ap, fmt);
        vsnprintf(buf, sizeof buf, fmt, ap);
        va_end(ap);
        err_code = code;
        err_line = line;
        err_msg = buf;
        failed.store(1, std::memory_order_release);
        finished.store(1, std::memory_order_release);
        wave0.store(1, std::memory_order_release);
    }

    // worker 0, between the barriers of wave w: the wave's records join the table
    void merge(int64_t w) {
        const char* text = cx.text;
        if (w == 0 && res[0].first_line_not_header) {
            fail(SKR_ERR_ARG, 0, "FASTA text does not start with a '>' header line");
            return;
        }
        uint64_t err_off = UINT64_MAX;
        int code = 0;
        for (int t = 0; t < T; ++t)
            if (res[t].err_off < err_off) { err_off = res[t].err_off; code = res[t].err_code; }
        const bool last = w == nwaves - 1;
        if (w == 0) {
            int64_t m0 = 0;
            for (int t = 0; t < T; ++t) m0 += (int64_t)res[t].recs.size();
            if (last) {
                cap_records = m0;
            } else {
                const double seen = (double)std::min(nbytes, (size_t)T * slice_bytes);
                cap_records = (int64_t)((double)m0 * ((double)nbytes / seen) * 1.25) + 1024;
            }
            const int64_t cap_blocks = (int64_t)(nbytes / 64) + cap_records + 2;
            int rc = layout_packed(P, cap_records, cap_blocks, pinned);
            if (rc != SKR_OK) { fail(rc, 0, "%s", skr::last_error().c_str()); return; }
            recs.reserve((size_t)std::max<int64_t>(cap_records, 1));
            P->header_spans.resize((size_t)cap_records * 2);
            P->body_spans.resize((size_t)cap_records * 2);
            const int64_t groups = cap_records / kGroup + 1;
            done.reset(new std::atomic<uint8_t>[(size_t)groups]);
            for (int64_t g = 0; g < groups; ++g) done[g].store(0, std::memory_order_relaxed);
        }
        for (int t = 0; t < T && !code; ++t) {
            for (const Rec& r : res[t].recs) {
                if (r.hdr_off >= err_off) break;  // the reference stops at the first offending line
                // empty record anywhere but last -> the reference's assert, at the header that follows it
                if (!recs.empty() && recs.back().bases == 0) {
                    err_off = r.hdr_off;
                    code = SKR_ERR_FASTA_HEADER;
                    break;
                }
                if (r.bases > 0xFFFFFFFFull) {
                    fail(SKR_ERR_ARG, 0, "record %lld longer than 2^32-1 bases", (long long)recs.size());
                    return;
                }
                const int64_t i = (int64_t)recs.size();
                if (i >= cap_records && !overflow.load(std::memory_order_relaxed)) overflow.store(1, std::memory_order_release);
                recs.push_back(r);
                if (!overflow.load(std::memory_order_relaxed)) {
                    P->blk_off[i] = blocks;
                    P->len[i] = (uint32_t)r.bases;
                    P->header_spans[2 * i] = r.hdr_off;
                    P->header_spans[2 * i + 1] = r.hdr_len;
                    P->body_spans[2 * i] = r.body_off;
                    P->body_spans[2 * i + 1] = r.body_len;
                }
                blocks += (r.bases + 63) / 64;
                total_bases += r.bases;
            }
        }
        if (code || err_off != UINT64_MAX) {
            if (!code) code = SKR_ERR_FASTA_BLANK;
            const int64_t line = line_of_offset(text, err_off);
            if (code == SKR_ERR_FASTA_BLANK) fail(code, line, "string index out of range (blank line %lld in FASTA)", (long long)line);
            else fail(code, line, "There may be a header without a sequence at line %lld.", (long long)(line - 1));
            return;
        }
        const int64_t n = (int64_t)recs.size();
        if (!overflow.load(std::memory_order_relaxed)) {
            P->blk_off[n] = blocks;
            if (last) finalize(n);
            pack_groups.store(last ? (n + kGroup - 1) / kGroup : n / kGroup, std::memory_order_release);
            scanned.store(n, std::memory_order_release);
            if (last) finished.store(1, std::memory_order_release);
        }
        wave0.store(1, std::memory_order_release);
    }

    void finalize(int64_t n) {
        P->m = n;
        P->nblocks = (int64_t)blocks + 1;
        P->total_bases = (int64_t)total_bases;
        P->header_spans.resize((size_t)n * 2);
        P->body_spans.resize((size_t)n * 2);
        memset(P->codes + blocks * 4, 0, 16);    // trailing pad block: codes 0, mask all ones
        memset(P->mask + blocks * 2, 0xFF, 8);
    }

    // the estimate was too small: an exact slab, filled by everybody after the last wave
    void rebuild_exact() {
        const int64_t n = (int64_t)recs.size();
        P->retired = P->slab;  // a consumer may still be copying from it
        P->slab = Slab();
        int rc = layout_packed(P, n, (int64_t)blocks + 1, pinned);
        if (rc != SKR_OK) { fail(rc, 0, "%s", skr::last_error().c_str()); return; }
        P->header_spans.resize((size_t)n * 2);
        P->body_spans.resize((size_t)n * 2);
        uint64_t b = 0;
        for (int64_t i = 0; i < n; ++i) {
            const Rec& r = recs[(size_t)i];
            P->blk_off[i] = b;
            P->len[i] = (uint32_t)r.bases;
            P->header_spans[2 * i] = r.hdr_off;
            P->header_spans[2 * i + 1] = r.hdr_len;
            P->body_spans[2 * i] = r.body_off;
            P->body_spans[2 * i + 1] = r.body_len;
            b += (r.bases + 63) / 64;
        }
        P->blk_off[n] = b;
        finalize(n);
        const int64_t groups = n / kGroup + 1;
        done.reset(new std::atomic<uint8_t>[(size_t)groups]);
        for (int64_t g = 0; g < groups; ++g) done[g].store(0, std::memory_order_relaxed);
        next_group.store(0, std::memory_order_relaxed);
        watermark.store(0, std::memory_order_relaxed);
        pack_groups.store((n + kGroup - 1) / kGroup, std::memory_order_release);
    }

    void pack_available() {
        for (;;) {
            int64_t g = next_group.load(std::memory_order_relaxed);
            if (g >= pack_groups.load(std::memory_order_acquire)) break;
            if (!next_group.compare_exchange_weak(g, g + 1, std::memory_order_relaxed)) continue;
            const int64_t i1 = std::min<int64_t>((g + 1) * kGroup, (int64_t)recs.size());
            for (int64_t i = g * kGroup; i < i1; ++i) pack_record(P, recs[(size_t)i], i, cx.text, cx.end, al, lut2, cx.use_avx2);
            done[g].store(1, std::memory_order_release);
        }
    }

    void worker(int t) {
        int sense = 0;
        const char* text = cx.text;
        if (t == 0 && pinned && device >= 0) cudaSetDevice(device);
        for (int64_t w = 0; w < nwaves; ++w) {
            const int64_t s = w * T + t;
            res[t] = ChunkResult();
            if (s < nslices && !failed.load(std::memory_order_relaxed)) {
                const size_t a = (size_t)s * slice_bytes, b = std::min(nbytes, a + slice_bytes);
                scan_slice(cx, text + a, text + b, s == 0, res[t]);
            }
            bar.wait(sense);
            if (t == 0 && !failed.load(std::memory_order_relaxed)) merge(w);
            bar.wait(sense);
            if (failed.load(std::memory_order_acquire)) return;
            if (!overflow.load(std::memory_order_acquire)) pack_available();
        }
        if (overflow.load(std::memory_order_acquire)) {
            if (t == 0) rebuild_exact();
            bar.wait(sense);
            if (failed.load(std::memory_order_acquire)) return;
            pack_available();
            bar.wait(sense);
            if (t == 0) {  // only now: the table AND the words of the exact slab are there
                scanned.store((int64_t)recs.size(), std::memory_order_release);
                finished.store(1, std::memory_order_release);
            }
        }
    }

    void start() {
        bar.n = T;
        res.resize((size_t)T);
        threads.reserve((size_t)T);
        for (int t = 0; t < T; ++t) threads.emplace_back([this, t] { worker(t); });
    }

    void join() {
        for (auto& t : threads) t.join();
        threads.clear();
    }

    static void pause(int& spins) {
        if (++spins < 2000) _mm_pause(); else std::this_thread::yield();
    }

    // blocks until records [0, want) are in the table, or the table is complete; SKR_ERR_CAPACITY after a fallback
    int wait_scanned(int64_t want, int64_t* avail, int* fin) {
        int spins = 0;
        for (;;) {
            if (overflow.load(std::memory_order_acquire)) return SKR_ERR_CAPACITY;
            const int f = finished.load(std::memory_order_acquire);
            const int64_t n = scanned.load(std::memory_order_acquire);
            if (f || (want >= 0 && n >= want)) {
                if (failed.load(std::memory_order_acquire)) return err_code;
                *avail = n;
                *fin = f;
                return SKR_OK;
            }
            pause(spins);
        }
    }

    void wait_finished() {
        int spins = 0;
        while (!finished.load(std::memory_order_acquire)) pause(spins);
    }

    // blocks until records [0, upto) are packed (upto <= scanned)
    void wait_records(int64_t upto) {
        const int64_t need = (upto + kGroup - 1) / kGroup;
        int64_t w = watermark.load(std::memory_order_relaxed);
        int spins = 0;
        while (w < need && !failed.load(std::memory_order_acquire)) {
            if (overflow.load(std::memory_order_acquire)) { wait_finished(); return; }  // rebuilt: everything is packed by then
            if (w < pack_groups.load(std::memory_order_acquire) && done[w].load(std::memory_order_acquire)) { ++w; spins = 0; continue; }
            pause(spins);
        }
        if (!overflow.load(std::memory_order_acquire)) watermark.store(w, std::memory_order_relaxed);
    }
};

namespace {

// one slab: [codes | mask | block offsets (cap + 1) | lengths (cap)], the same order alloc_packed uses
int layout_packed(SkrPacked* P, int64_t cap_records, int64_t cap_blocks, bool pinned) {
    size_t off_codes = 0;
    size_t off_mask = align_up(off_codes + (size_t)cap_blocks * 16, 256);
    size_t off_blk = align_up(off_mask + (size_t)cap_blocks * 8, 256);
    size_t off_len = align_up(off_blk + (size_t)(cap_records + 1) * 8, 256);
    size_t total = align_up(off_len + (size_t)std::max<int64_t>(cap_records, 1) * 4, 256);
    int rc = slab_alloc(total, pinned, &P->slab);
    if (rc != SKR_OK) return rc;
    P->slab_bytes = total;
    char* base = (char*)P->slab.ptr;
    P->codes = (uint32_t*)(base + off_codes);
    P->mask = (uint32_t*)(base + off_mask);
    P->blk_off = (uint64_t*)(base + off_blk);
    P->len = (uint32_t*)(base + off_len);
    return SKR_OK;
}

// the error a background job ended with, raised on the thread that asks
int wave_error(WaveJob* w) {
    g_error_line = w->err_line;
    return skr::fail(w->err_code, "%s", w->err_msg.c_str());
}

}  // namespace

extern "C" int64_t skr_pack_error_line(void) { return g_error_line; }

static int pack_fasta_buffer(const void* text_v, size_t nbytes, const uint8_t* lut, int nthreads, int pinned, bool async,
                             SkrPacked** out);

extern "C" int skr_pack_fasta_buffer(const void* text_v, size_t nbytes, const uint8_t* lut, int nthreads, int pinned,
                                     SkrPacked** out) {
    return pack_fasta_buffer(text_v, nbytes, lut, nthreads, pinned, false, out);
}

extern "C" int skr_pack_fasta_buffer_async(const void* text_v, size_t nbytes, const uint8_t* lut, int nthreads, int pinned,
                                           SkrPacked** out) {
    return pack_fasta_buffer(text_v, nbytes, lut, nthreads, pinned, true, out);
}

extern "C" int skr_packed_wait_records(SkrPacked* p, int64_t upto) {
    if (!p) return skr::fail(SKR_ERR_ARG, "skr_packed_wait_records: null handle");
    if (p->wave) {
        WaveJob* w = p->wave;
        int64_t avail = 0;
        int fin = 0;
        int rc = w->wait_scanned(upto, &avail, &fin);  // upto < 0: the whole table
        if (rc == SKR_ERR_CAPACITY) { w->wait_finished(); rc = w->failed.load() ? w->err_code : SKR_OK; avail = w->scanned.load(); }
        if (rc != SKR_OK) return wave_error(w);
        w->wait_records(upto < 0 ? avail : std::min(upto, avail));
        return w->failed.load() ? wave_error(w) : SKR_OK;
    }
    if (p->job) p->job->wait_records(upto < 0 ? p->m : std::min(upto, p->m));
    return SKR_OK;
}

extern "C" int skr_packed_wait_scanned(SkrPacked* p, int64_t want, int64_t* avail, int* finished) {
    if (!p || !avail || !finished) return skr::fail(SKR_ERR_ARG, "skr_packed_wait_scanned: null argument");
    if (!p->wave) {
        *avail = p->m;
        *finished = 1;
        return SKR_OK;
    }
    int rc = p->wave->wait_scanned(want, avail, finished);
    if (rc == SKR_ERR_CAPACITY)
        return skr::fail(rc, "the record estimate of the streamed packer was too small; the handle is complete after skr_packed_wait");
    return rc == SKR_OK ? SKR_OK : wave_error(p->wave);
}

extern "C" int64_t skr_packed_capacity_records(const SkrPacked* p) {
    if (!p) return 0;
    return p->wave ? p->wave->cap_records : p->m;
}

extern "C" int skr_packed_wait(SkrPacked* p) {
    if (!p) return skr::fail(SKR_ERR_ARG, "skr_packed_wait: null handle");
    if (p->job) {
        p->job->join();
        delete p->job;
        p->job = nullptr;
    }
    if (p->wave) {
        p->wave->join();
        if (p->wave->failed.load()) return wave_error(p->wave);  // the job stays: later calls report the same error
        delete p->wave;
        p->wave = nullptr;
    }
    return SKR_OK;
}

// the record table of a handle whose scan still runs is not final: everything that depends on it waits here
static bool table_ready(const SkrPacked* p) {
    if (!p->wave) return true;
    p->wave->wait_finished();
    return !p->wave->failed.load();
}

static int pack_fasta_buffer(const void* text_v, size_t nbytes, const uint8_t* lut, int nthreads, int pinned, bool async,
                             SkrPacked** out) {
    g_error_line = 0;
    if (!out || !lut || (!text_v && nbytes)) return skr::fail(SKR_ERR_ARG, "skr_pack_fasta_buffer: null argument");
    *out = nullptr;
    const char* text = (const char*)text_v;
    const char* end = text + nbytes;
    uint8_t lut2[256];
    build_lut2(lut, lut2);
    const SimdAlphabet al = simd_alphabet(lut);
    const bool use_avx2 = __builtin_cpu_supports("avx2") && __builtin_cpu_supports("bmi2") && !getenv("SKR_PACK_NO_AVX2");
    int T = pick_threads(nthreads, nbytes);

    const bool profile = getenv("SKR_PACK_PROFILE") != nullptr;
    auto now = [] { return std::chrono::steady_clock::now(); };
    auto ms = [](std::chrono::steady_clock::time_point a, std::chrono::steady_clock::time_point b) {
        return std::chrono::duration<double, std::milli>(b - a).count();
    };
    // large texts in background mode: scan and pack in waves, the call returns after the first one
    size_t wave_min = (size_t)32 << 20;
    if (const char* env = getenv("SEEKR_B200_WAVE_MIN_BYTES")) wave_min = (size_t)strtoull(env, nullptr, 10);
    if (async && T >= 2 && nbytes >= wave_min && !getenv("SEEKR_B200_NO_WAVES")) {
        const auto t0 = now();
        SkrPacked* P = new SkrPacked();
        WaveJob* job = new WaveJob();
        job->P = P;
        job->cx = ScanCtx{text, end, use_avx2, use_avx2 && !getenv("SKR_PACK_NO_FAST_SCAN")};
        job->nbytes = nbytes;
        job->al = al;
        memcpy(job->lut2, lut2, 256);
        job->pinned = pinned != 0;
        if (job->pinned && cudaGetDevice(&job->device) != cudaSuccess) job->device = -1;
        job->T = T;
        int waves = 16;
        if (const char* env = getenv("SEEKR_B200_WAVES")) waves = std::max(1, atoi(env));
        size_t slice_min = (size_t)64 << 10;
        if (const char* env = getenv("SEEKR_B200_WAVE_SLICE_MIN")) slice_min = std::max<size_t>(64, (size_t)strtoull(env, nullptr, 10));
        job->slice_bytes = std::max<size_t>(slice_min, (nbytes + (size_t)T * waves - 1) / ((size_t)T * waves));
        job->nslices = (int64_t)((nbytes + job->slice_bytes - 1) / job->slice_bytes);
        job->nwaves = (job->nslices + T - 1) / T;
        P->wave = job;
        job->start();
        int spins = 0;
        while (!job->wave0.load(std::memory_order_acquire)) WaveJob::pause(spins);
        if (job->failed.load(std::memory_order_acquire)) {  // wave 0 holds what a one-shot scan reports first
            job->join();
            const int rc = wave_error(job);
            skr_packed_free(P);
            return rc;
        }
        if (profile)
            fprintf(stderr, "skr_pack (waves): %d threads, %lld slices of %zu bytes in %lld waves, first wave + slab %.2f ms, "
                    "capacity %lld records (%zu bytes)\n", T, (long long)job->nslices, job->slice_bytes, (long long)job->nwaves,
                    ms(t0, now()), (long long)job->cap_records, nbytes);
        *out = P;
        return SKR_OK;
    }
    const auto t_start = now();
    // ---- pass 1: records and their lengths, chunked by byte range ----------------------------
    std::vector<ChunkResult> res(T);
    const ScanCtx cx{text, end, use_avx2, use_avx2 && !getenv("SKR_PACK_NO_FAST_SCAN")};
    run_threads(T, [&](int t) {
        scan_slice(cx, text + nbytes * (size_t)t / (size_t)T, text + nbytes * (size_t)(t + 1) / (size_t)T, t == 0, res[t]);
    });

    // the earliest error wins: the reference stops at the first offending line
    uint64_t err_off = UINT64_MAX;
    int err_code = 0;
    for (int t = 0; t < T; ++t)
        if (res[t].err_off < err_off) { err_off = res[t].err_off; err_code = res[t].err_code; }
    if (res[0].first_line_not_header)
        return skr::fail(SKR_ERR_ARG, "FASTA text does not start with a '>' header line");

    std::vector<Rec> recs;
    size_t nrec = 0;
    for (auto& r : res) nrec += r.recs.size();
    recs.reserve(nrec);
    for (auto& r : res) recs.insert(recs.end(), r.recs.begin(), r.recs.end());
    // empty record anywhere but last -> the reference's assert (only i == 0 is exempt, which is the
    // first header itself, never an empty record followed by a header)
    for (size_t i = 0; i + 1 < recs.size(); ++i) {
        if (recs[i].bases == 0) {
            uint64_t off = recs[i + 1].hdr_off;
            if (off < err_off) { err_off = off; err_code = SKR_ERR_FASTA_HEADER; }
            break;
        }
    }
    if (err_code) {
        const int64_t line = line_of_offset(text, err_off);
        g_error_line = line;
        if (err_code == SKR_ERR_FASTA_BLANK)
            return skr::fail(err_code, "string index out of range (blank line %lld in FASTA)", (long long)line);
        return skr::fail(err_code, "There may be a header without a sequence at line %lld.", (long long)(line - 1));
    }

    const auto t_pass1 = now();
    // ---- pass 2: allocate + pack ---------------------------------------------------------------
    SkrPacked* P = new SkrPacked();
    int64_t m = (int64_t)recs.size();
    std::vector<uint64_t> bases((size_t)m);
    for (int64_t i = 0; i < m; ++i) {
        bases[i] = recs[i].bases;
        if (bases[i] > 0xFFFFFFFFull) {
            delete P;
            return skr::fail(SKR_ERR_ARG, "record %lld longer than 2^32-1 bases", (long long)i);
        }
    }
    int rc = alloc_packed(P, m, bases, pinned != 0);
    if (rc != SKR_OK) { delete P; return rc; }
    P->header_spans.resize((size_t)m * 2);
    P->body_spans.resize((size_t)m * 2);
    for (int64_t i = 0; i < m; ++i) {  // the spans belong to the record table: final before packing starts
        P->header_spans[2 * i] = recs[i].hdr_off;
        P->header_spans[2 * i + 1] = recs[i].hdr_len;
        P->body_spans[2 * i] = recs[i].body_off;
        P->body_spans[2 * i + 1] = recs[i].body_len;
    }
    const auto t_alloc = now();
    PackJob* job = new PackJob(P, std::move(recs), text, end, al, lut2, use_avx2);
    if (async) {
        // the record table (lengths, block offsets) is final; codes and mask are filled by T background threads in
        // record order, consumers follow the progress with skr_packed_wait_records
        P->job = job;
        job->start(T);
        *out = P;
        if (profile)
            fprintf(stderr, "skr_pack (async): %d threads, scan %.2f ms, merge+alloc %.2f ms (%zu bytes, %lld records)\n", T,
                    ms(t_start, t_pass1), ms(t_pass1, t_alloc), nbytes, (long long)m);
        return SKR_OK;
    }
    run_threads(T, [&](int) { job->work(); });
    delete job;
    if (profile)
        fprintf(stderr, "skr_pack: %d threads, scan %.2f ms, merge+alloc %.2f ms, pack %.2f ms (%zu bytes, %lld records)\n", T,
                ms(t_start, t_pass1), ms(t_pass1, t_alloc), ms(t_alloc, now()), nbytes, (long long)m);
    *out = P;
    return SKR_OK;
}

extern "C" int skr_pack_fasta_file(const char* path, const uint8_t* lut, int nthreads, int pinned, SkrPacked** out) {
    if (!path) return skr::fail(SKR_ERR_ARG, "skr_pack_fasta_file: null path");
    int fd = open(path, O_RDONLY);
    if (fd < 0) return skr::fail(SKR_ERR_IO, "cannot open %s", path);
    struct stat st;
    if (fstat(fd, &st) != 0) { close(fd); return skr::fail(SKR_ERR_IO, "cannot stat %s", path); }
    size_t n = (size_t)st.st_size;
    if (n == 0) { close(fd); return skr_pack_fasta_buffer("", 0, lut, nthreads, pinned, out); }
    void* map = mmap(nullptr, n, PROT_READ, MAP_PRIVATE | MAP_POPULATE, fd, 0);
    close(fd);
    if (map == MAP_FAILED) return skr::fail(SKR_ERR_IO, "cannot mmap %s", path);
    madvise(map, n, MADV_SEQUENTIAL);
    int rc = skr_pack_fasta_buffer(map, n, lut, nthreads, pinned, out);
    munmap(map, n);
    return rc;
}

extern "C" int skr_pack_sequences(const void* letters_v, const int64_t* offs, int64_t m, const uint8_t* lut,
                                  int nthreads, int pinned, SkrPacked** out) {
    g_error_line = 0;
    if (!out || !lut || !offs || m < 0) return skr::fail(SKR_ERR_ARG, "skr_pack_sequences: bad argument");
    *out = nullptr;
    const unsigned char* letters = (const unsigned char*)letters_v;
    uint8_t lut2[256];
    build_lut2(lut, lut2);
    const SimdAlphabet al = simd_alphabet(lut);
    std::vector<uint64_t> bases((size_t)m);
    for (int64_t i = 0; i < m; ++i) {
        if (offs[i + 1] < offs[i]) return skr::fail(SKR_ERR_ARG, "offsets must be non-decreasing");
        bases[i] = (uint64_t)(offs[i + 1] - offs[i]);
        if (bases[i] > 0xFFFFFFFFull) return skr::fail(SKR_ERR_ARG, "record %lld too long", (long long)i);
    }
    SkrPacked* P = new SkrPacked();
    int rc = alloc_packed(P, m, bases, pinned != 0);
    if (rc != SKR_OK) { delete P; return rc; }
    int T = pick_threads(nthreads, m ? (size_t)(offs[m] - offs[0]) : 0);
    std::atomic<int64_t> next_rec{0};
    run_threads(T, [&](int) {
        for (;;) {
            int64_t i0 = next_rec.fetch_add(64);
            if (i0 >= m) break;
            int64_t i1 = std::min(m, i0 + 64);
            for (int64_t i = i0; i < i1; ++i) {
                uint64_t b0 = P->blk_off[i], b1 = P->blk_off[i + 1];
                BitWriter w{P->codes + b0 * 4, P->mask + b0 * 2};
                if (al.ok) pack_segment_avx2(w, (const char*)letters + offs[i], (const char*)letters + offs[i + 1], al,
                                             (const char*)letters + offs[m]);
                else for (int64_t p = offs[i]; p < offs[i + 1]; ++p) w.put(lut2[letters[p]]);
                w.finish(P->codes + b1 * 4, P->mask + b1 * 2);
            }
        }
    });
    *out = P;
    return SKR_OK;
}

extern "C" void skr_packed_free(SkrPacked* p) {
    if (!p) return;
    skr_packed_wait(p);  // background packing writes into the slab
    if (p->wave) {       // a failed job is kept for its error message
        delete p->wave;
        p->wave = nullptr;
    }
    slab_free(p->slab);
    slab_free(p->retired);
    delete p;
}

// -1 when a background scan ended with an error (skr_packed_wait reports it)
extern "C" int64_t skr_packed_num_records(const SkrPacked* p) { return table_ready(p) ? p->m : -1; }
extern "C" int64_t skr_packed_num_blocks(const SkrPacked* p) { return table_ready(p) ? p->nblocks : -1; }
extern "C" int64_t skr_packed_total_bases(const SkrPacked* p) { return table_ready(p) ? p->total_bases : -1; }
extern "C" const uint32_t* skr_packed_codes(const SkrPacked* p) { return p->codes; }
extern "C" const uint32_t* skr_packed_mask(const SkrPacked* p) { return p->mask; }
extern "C" const uint64_t* skr_packed_block_offsets(const SkrPacked* p) { return p->blk_off; }
extern "C" const uint32_t* skr_packed_lengths(const SkrPacked* p) { return p->len; }
extern "C" const uint64_t* skr_packed_header_spans(const SkrPacked* p) { return table_ready(p) ? p->header_spans.data() : nullptr; }
extern "C" const uint64_t* skr_packed_body_spans(const SkrPacked* p) { return table_ready(p) ? p->body_spans.data() : nullptr; }
extern "C" const void* skr_packed_slab(const SkrPacked* p) { return p->slab.ptr; }
extern "C" size_t skr_packed_slab_bytes(const SkrPacked* p) { return p->slab_bytes; }
