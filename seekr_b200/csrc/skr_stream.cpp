// Streamed get_counts(): FASTA text -> host count matrix with packing, H2D, counting and D2H overlapped.
//
// Replaces the serial sequence of seekr/fasta_reader.py:41-63 (read everything) + kmer_counts.py:196-200 (count
// everything) + the copy of the finished matrix for the cases where rows are independent once the vectors are
// known: Log2.none / Log2.pre with or without supplied vectors, and Log2.post with supplied vectors through the
// speculated shift (skr_post_spec).  The packer fills the pinned slab in record order on background threads
// (skr_pack_fasta_buffer_async); this driver follows its progress chunk by chunk:
//
//      pack chunk i+2   ||   H2D chunk i+1 (copy-in stream)   ||   count chunk i   ||   D2H chunk i-1 (copy-out stream)
//
// so the end-to-end time is the longest of the four, which is the D2H of the 4 * 4^k bytes per record.  The host
// destination is written directly when it is pinned; a pageable destination is reached through a small ring of
// pinned slots that host threads drain with memcpy (no allocation of result-sized pinned memory on a cold start).
#include <cuda_runtime.h>
#include <sys/mman.h>
#include <unistd.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "skr_common.h"

namespace {

struct StreamRes {  // per-device streams, created once
    cudaStream_t in = nullptr, out = nullptr;
    int dev = -1;
};
thread_local StreamRes g_res;

int get_streams(StreamRes** out) {
    int dev = 0;
    SKR_CUDA_CHECK(cudaGetDevice(&dev));
    if (g_res.dev != dev) {
        if (g_res.in) { cudaStreamDestroy(g_res.in); cudaStreamDestroy(g_res.out); }
        SKR_CUDA_CHECK(cudaStreamCreateWithFlags(&g_res.in, cudaStreamNonBlocking));
        SKR_CUDA_CHECK(cudaStreamCreateWithFlags(&g_res.out, cudaStreamNonBlocking));
        g_res.dev = dev;
    }
    *out = &g_res;
    return SKR_OK;
}

#ifndef MADV_POPULATE_WRITE
#define MADV_POPULATE_WRITE 23
#endif

// A pageable destination that has never been touched costs one page fault per 4 KB inside the copier's memcpy
// (1.6 s for the 4.1 GB result of 250 000 transcripts); asking the kernel for the whole range of a chunk in one call
// maps the pages without the per-page trap.  Best effort: older kernels return EINVAL and the memcpy faults as before.
void prefault_for_write(char* begin, char* end, char* lo, char* hi) {
    static const uintptr_t page = (uintptr_t)sysconf(_SC_PAGESIZE);
    uintptr_t a = (uintptr_t)begin & ~(page - 1), b = ((uintptr_t)end + page - 1) & ~(page - 1);
    if (a < (uintptr_t)lo) a = ((uintptr_t)lo + page - 1) & ~(page - 1);
    if (b > (uintptr_t)hi) b = (uintptr_t)hi & ~(page - 1);
    if (b > a) (void)madvise((void*)a, (size_t)(b - a), MADV_POPULATE_WRITE);
}

struct Chunk {
    int64_t r0 = 0, r1 = 0;
    cudaEvent_t in_done = nullptr, k_done = nullptr, out_done = nullptr;
    int slot = -1;
    double t_issue = 0;  // host time the chunk's work was enqueued (profile)
};

double now_ms() {
    return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

}  // namespace

extern "C" int skr_stream_counts(SkrPacked* packed, SkrStreamArgs* sa, void* stream_v) {
    if (!packed || !sa) return skr::fail(SKR_ERR_ARG, "skr_stream_counts: null argument");
    sa->records_done = 0;
    // a handle whose scan still runs (large texts) has no record count yet: the chunks follow the scan
    int64_t avail = 0;
    int fin = 0;
    int rc = skr_packed_wait_scanned(packed, 0, &avail, &fin);
    if (rc != SKR_OK) return rc;
    const int64_t capacity = sa->capacity_records > 0 ? sa->capacity_records : skr_packed_num_records(packed);
    if (capacity < 0) return skr_packed_wait(packed);  // the scan failed: its error
    if (fin && avail == 0) return SKR_OK;
    if (!sa->d_slab || !sa->count.d_out) return skr::fail(SKR_ERR_ARG, "skr_stream_counts: device slab and device matrix are required");
    if (sa->count.out_is_f64) return skr::fail(SKR_ERR_ARG, "skr_stream_counts: float32 matrices only");
    const int k = sa->count.k;
    if (k < 1 || k > 8) return skr::fail(SKR_ERR_ARG, "skr_stream_counts: k=%d not supported", k);
    const int64_t cols = (int64_t)1 << (2 * k);
    if (sa->h_out && sa->h_ld < cols) return skr::fail(SKR_ERR_ARG, "skr_stream_counts: host pitch < 4^k");
    cudaStream_t s_k = (cudaStream_t)stream_v;
    StreamRes* res = nullptr;
    rc = get_streams(&res);
    if (rc != SKR_OK) return rc;
    const bool profile = getenv("SKR_STREAM_PROFILE") != nullptr;
    const double t_begin = now_ms();

    // host views of the slab and the matching device addresses (same layout on both sides)
    const char* h_slab = (const char*)skr_packed_slab(packed);
    char* d_slab = (char*)sa->d_slab;
    const uint32_t* h_codes = skr_packed_codes(packed);
    const uint32_t* h_mask = skr_packed_mask(packed);
    const uint64_t* h_blk = skr_packed_block_offsets(packed);
    const uint32_t* h_len = skr_packed_lengths(packed);
    const size_t off_codes = (size_t)((const char*)h_codes - h_slab), off_mask = (size_t)((const char*)h_mask - h_slab);
    const size_t off_blk = (size_t)((const char*)h_blk - h_slab), off_len = (size_t)((const char*)h_len - h_slab);

    // chunks: a small first one so the pipeline fills quickly, then ~chunk_bytes of output each
    int64_t chunk_mb = 32;
    // experiment knob; 16 ... 96 MB all end within the run-to-run noise of the 32 MB default (17.5 - 19 ms at S50k)
    if (const char* env = getenv("SEEKR_B200_STREAM_CHUNK_MB")) chunk_mb = std::max(1, atoi(env));
    int64_t per = sa->chunk_records > 0 ? sa->chunk_records : std::max<int64_t>(256, (chunk_mb << 20) / (cols * 4));
    const bool ring = sa->h_out && !sa->h_out_pinned;
    if (ring) per = std::min<int64_t>(per, std::max<int64_t>(64, ((int64_t)8 << 20) / (cols * 4)));
    const int64_t first = std::max<int64_t>(64, per / 4);
    const int max_chunks = (int)(capacity / per + 3);
    std::vector<Chunk> chunks((size_t)max_chunks);
    std::atomic<int> nchunks{-1};  // known when the last chunk has been issued

    // whatever the caller enqueued on its stream (vectors, the speculation cell) comes before the first kernel, and
    // the copy-in stream must not overwrite a slab an earlier launch on that stream still reads
    cudaEvent_t start_ev;
    const unsigned ev_flags = profile ? cudaEventDefault : cudaEventDisableTiming;  // the profile prints a device timeline
    SKR_CUDA_CHECK(cudaEventCreateWithFlags(&start_ev, ev_flags));
    SKR_CUDA_CHECK(cudaEventRecord(start_ev, s_k));
    SKR_CUDA_CHECK(cudaStreamWaitEvent(res->in, start_ev, 0));
    SKR_CUDA_CHECK(cudaStreamWaitEvent(res->out, start_ev, 0));

    // pinned ring for a pageable destination
    constexpr int kSlots = 4;
    size_t slot_bytes = 0;
    char* slots[kSlots] = {nullptr, nullptr, nullptr, nullptr};
    if (ring) {
        slot_bytes = (size_t)std::max(per, first) * cols * 4;
        for (int i = 0; i < kSlots; ++i) {
            rc = skr_host_alloc(slot_bytes, (void**)&slots[i]);
            if (rc != SKR_OK) {
                for (int j = 0; j < i; ++j) skr_host_free(slots[j]);
                cudaEventDestroy(start_ev);
                return rc;
            }
        }
    }
    // copier threads for the ring: each drains whole chunks (event wait, then memcpy of its rows)
    std::atomic<int> copy_next{0}, copy_fail{0};
    std::vector<std::atomic<int>> issued((size_t)max_chunks), drained((size_t)max_chunks);
    for (int i = 0; i < max_chunks; ++i) { issued[i].store(0); drained[i].store(0); }
    std::vector<std::thread> copiers;
    const int dev = res->dev;
    if (ring) {
        int nthreads = sa->copy_threads > 0 ? sa->copy_threads : 4;
        for (int t = 0; t < nthreads; ++t) {
            copiers.emplace_back([&, dev] {
                cudaSetDevice(dev);
                for (;;) {
                    const int i = copy_next.fetch_add(1);
                    for (;;) {  // chunk i issued, or no chunk i
                        const int n = nchunks.load(std::memory_order_acquire);
                        if ((n >= 0 && i >= n) || copy_fail.load()) return;
                        if (i < max_chunks && issued[i].load(std::memory_order_acquire)) break;
                        std::this_thread::yield();
                    }
                    const Chunk& c = chunks[i];
                    const size_t row = (size_t)cols * 4;
                    char* dst = (char*)sa->h_out + (size_t)c.r0 * (size_t)sa->h_ld * 4;
                    // map the chunk's pages while its rows are still on their way
                    prefault_for_write(dst, dst + (size_t)(c.r1 - c.r0) * (size_t)sa->h_ld * 4, (char*)sa->h_out,
                                       (char*)sa->h_out + (size_t)capacity * (size_t)sa->h_ld * 4);
                    if (cudaEventSynchronize(c.out_done) != cudaSuccess) { copy_fail.store(1); drained[i].store(1); return; }
                    const char* src = slots[c.slot];
                    if ((size_t)sa->h_ld == (size_t)cols) {
                        memcpy(dst, src, (size_t)(c.r1 - c.r0) * row);
                    } else {
                        for (int64_t r = 0; r < c.r1 - c.r0; ++r) memcpy(dst + (size_t)r * sa->h_ld * 4, src + (size_t)r * row, row);
                    }
                    drained[i].store(1, std::memory_order_release);
                }
            });
        }
    }

    int status = SKR_OK;
    double t_first_kernel = 0, t_last_issue = 0;
    int64_t r = 0;
    int i = 0;
    for (; status == SKR_OK; ++i) {
        // ---- the next chunk's records: in the table (scan), then packed
        // (chunks of 4 x `per` were tried for the pinned destination: fewer hand-overs, but 134 MB copies ran at
        // 49.5 GB/s while the packer was still at work against 52.5 GB/s for 33 MB ones, and the last chunk's copy
        // starts later: 15.8 against 15.2 ms, profiles/r02_e2e_timeline.txt)
        const int64_t want = r + (i == 0 ? first : per);
        status = skr_packed_wait_scanned(packed, want, &avail, &fin);
        if (status != SKR_OK) break;
        const int64_t r1 = std::min(want, avail);
        if (r1 == r) break;  // fin: nothing left
        if (r1 > capacity || i >= max_chunks) {
            status = skr::fail(SKR_ERR_CAPACITY, "skr_stream_counts: more records than the %lld the buffers were sized for", (long long)capacity);
            break;
        }
        Chunk& c = chunks[i];
        c.r0 = r;
        c.r1 = r1;
        cudaEventCreateWithFlags(&c.in_done, ev_flags);
        cudaEventCreateWithFlags(&c.k_done, ev_flags);
        cudaEventCreateWithFlags(&c.out_done, ev_flags);
        c.t_issue = now_ms() - t_begin;
        status = skr_packed_wait_records(packed, c.r1);
        if (status != SKR_OK) break;
        // ---- its table entries and its code and mask words go in
        const uint64_t b0 = h_blk[c.r0], b1 = h_blk[c.r1];
        uint32_t longest = 0;
        for (int64_t q = c.r0; q < c.r1; ++q) longest = std::max(longest, h_len[q]);
        cudaError_t e = cudaMemcpyAsync(d_slab + off_blk + (size_t)c.r0 * 8, h_blk + c.r0, (size_t)(c.r1 - c.r0 + 1) * 8,
                                        cudaMemcpyHostToDevice, res->in);
        if (e == cudaSuccess)
            e = cudaMemcpyAsync(d_slab + off_len + (size_t)c.r0 * 4, h_len + c.r0, (size_t)(c.r1 - c.r0) * 4, cudaMemcpyHostToDevice, res->in);
        if (b1 > b0 && e == cudaSuccess) {
            e = cudaMemcpyAsync(d_slab + off_codes + b0 * 16, (const char*)h_codes + b0 * 16, (size_t)(b1 - b0) * 16,
                                cudaMemcpyHostToDevice, res->in);
            if (e == cudaSuccess)
                e = cudaMemcpyAsync(d_slab + off_mask + b0 * 8, (const char*)h_mask + b0 * 8, (size_t)(b1 - b0) * 8,
                                    cudaMemcpyHostToDevice, res->in);
        }
        if (e == cudaSuccess) e = cudaEventRecord(c.in_done, res->in);
        if (e == cudaSuccess) e = cudaStreamWaitEvent(s_k, c.in_done, 0);
        if (e != cudaSuccess) { status = skr::fail(SKR_ERR_CUDA, "skr_stream_counts: H2D failed: %s", cudaGetErrorString(e)); break; }
        // ---- count the chunk's records into their rows of the device matrix
        SkrCountArgs a = sa->count;
        a.d_codes = (const uint32_t*)(d_slab + off_codes);
        a.d_mask = (const uint32_t*)(d_slab + off_mask);
        a.d_block_offsets = (const uint64_t*)(d_slab + off_blk) + c.r0;
        a.d_lengths = (const uint32_t*)(d_slab + off_len) + c.r0;
        a.m = c.r1 - c.r0;
        a.max_length = (int64_t)longest;
        a.d_out = (char*)sa->count.d_out + (size_t)c.r0 * (size_t)sa->count.ld_out * 4;
        status = skr_count_ex(&a, s_k);
        if (status != SKR_OK) break;
        if (i == 0) t_first_kernel = now_ms();
        e = cudaEventRecord(c.k_done, s_k);
        // ---- and out again
        if (sa->h_out && e == cudaSuccess) {
            e = cudaStreamWaitEvent(res->out, c.k_done, 0);
            const char* src = (const char*)a.d_out;
            const size_t rows = (size_t)(c.r1 - c.r0), row = (size_t)cols * 4;
            if (ring) {
                c.slot = i % kSlots;
                if (i >= kSlots) {  // the slot's previous tenant must have been drained
                    while (!drained[i - kSlots].load(std::memory_order_acquire)) {
                        if (copy_fail.load()) break;
                        std::this_thread::yield();
                    }
                }
                if (e == cudaSuccess)
                    e = cudaMemcpy2DAsync(slots[c.slot], row, src, (size_t)sa->count.ld_out * 4, row, rows, cudaMemcpyDeviceToHost, res->out);
            } else if (e == cudaSuccess) {
                char* dst = (char*)sa->h_out + (size_t)c.r0 * (size_t)sa->h_ld * 4;
                if (sa->h_ld == cols && sa->count.ld_out == cols)  // both sides dense: one linear copy
                    e = cudaMemcpyAsync(dst, src, rows * row, cudaMemcpyDeviceToHost, res->out);
                else
                    e = cudaMemcpy2DAsync(dst, (size_t)sa->h_ld * 4, src, (size_t)sa->count.ld_out * 4, row, rows,
                                          cudaMemcpyDeviceToHost, res->out);
            }
            if (e == cudaSuccess) e = cudaEventRecord(c.out_done, res->out);
        }
        if (e != cudaSuccess) { status = skr::fail(SKR_ERR_CUDA, "skr_stream_counts: %s", cudaGetErrorString(e)); break; }
        issued[i].store(1, std::memory_order_release);
        r = c.r1;
    }
    const int issued_chunks = i;  // chunk i (if the loop stopped inside it) was not issued
    nchunks.store(issued_chunks, std::memory_order_release);
    t_last_issue = now_ms();
    const std::string first_error = status != SKR_OK ? skr::last_error() : std::string();
    if (status != SKR_OK) copy_fail.store(1);
    for (auto& t : copiers) t.join();
    cudaError_t e1 = cudaStreamSynchronize(res->out);
    cudaError_t e2 = cudaStreamSynchronize(res->in);
    if (status == SKR_OK) {
        // the trailing pad block (written when the table was completed, not by the pack threads)
        const int64_t nblocks = skr_packed_num_blocks(packed);
        cudaMemcpyAsync(d_slab + off_codes + (size_t)(nblocks - 1) * 16, (const char*)h_codes + (size_t)(nblocks - 1) * 16, 16,
                        cudaMemcpyHostToDevice, s_k);
        cudaMemcpyAsync(d_slab + off_mask + (size_t)(nblocks - 1) * 8, (const char*)h_mask + (size_t)(nblocks - 1) * 8, 8,
                        cudaMemcpyHostToDevice, s_k);
    }
    // the caller's stream continues after the last chunk; nothing later on it may race with our copy-out reads
    if (status == SKR_OK && issued_chunks > 0 && chunks[issued_chunks - 1].out_done && sa->h_out)
        cudaStreamWaitEvent(s_k, chunks[issued_chunks - 1].out_done, 0);
    if (profile && status == SKR_OK && sa->h_out) {
        fprintf(stderr, "skr_stream timeline (ms after the start event; host issue time in brackets):\n");
        for (int q = 0; q < issued_chunks; ++q) {
            float a = 0, b = 0, d = 0;
            cudaEventElapsedTime(&a, start_ev, chunks[q].in_done);
            cudaEventElapsedTime(&b, start_ev, chunks[q].k_done);
            cudaEventElapsedTime(&d, start_ev, chunks[q].out_done);
            fprintf(stderr, "  chunk %2d records %6lld-%6lld [%6.2f]  in %6.2f  counted %6.2f  out %6.2f\n", q, (long long)chunks[q].r0,
                    (long long)chunks[q].r1, chunks[q].t_issue, a, b, d);
        }
    }
    for (auto& c : chunks) {
        if (c.in_done) cudaEventDestroy(c.in_done);
        if (c.k_done) cudaEventDestroy(c.k_done);
        if (c.out_done) cudaEventDestroy(c.out_done);
    }
    cudaEventDestroy(start_ev);
    for (int q = 0; q < kSlots; ++q) if (slots[q]) skr_host_free(slots[q]);
    if (status != SKR_OK) {
        skr::last_error() = first_error;
        return status;
    }
    if (e1 != cudaSuccess || e2 != cudaSuccess)
        return skr::fail(SKR_ERR_CUDA, "skr_stream_counts: %s", cudaGetErrorString(e1 != cudaSuccess ? e1 : e2));
    if (copy_fail.load()) return skr::fail(SKR_ERR_CUDA, "skr_stream_counts: copy-out failed");
    sa->records_done = r;
    if (profile)
        fprintf(stderr, "skr_stream: %d chunks (%s destination), first kernel issued at %.2f ms, last chunk issued at %.2f ms, done at %.2f ms\n",
                issued_chunks, ring ? "pageable, pinned ring" : (sa->h_out ? "pinned" : "no host"), t_first_kernel - t_begin,
                t_last_issue - t_begin, now_ms() - t_begin);
    return SKR_OK;
}
