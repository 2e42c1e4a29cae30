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

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstdlib>
#include <cstring>
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

struct Chunk {
    int64_t r0 = 0, r1 = 0;
    cudaEvent_t in_done = nullptr, k_done = nullptr, out_done = nullptr;
    int slot = -1;
};

double now_ms() {
    return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

}  // namespace

extern "C" int skr_stream_counts(SkrPacked* packed, const SkrStreamArgs* sa, void* stream_v) {
    if (!packed || !sa) return skr::fail(SKR_ERR_ARG, "skr_stream_counts: null argument");
    const int64_t m = skr_packed_num_records(packed);
    if (m == 0) return SKR_OK;
    if (!sa->d_slab || !sa->count.d_out) return skr::fail(SKR_ERR_ARG, "skr_stream_counts: device slab and device matrix are required");
    if (sa->count.out_is_f64) return skr::fail(SKR_ERR_ARG, "skr_stream_counts: float32 matrices only");
    const int k = sa->count.k;
    if (k < 1 || k > 8) return skr::fail(SKR_ERR_ARG, "skr_stream_counts: k=%d not supported", k);
    const int64_t cols = (int64_t)1 << (2 * k);
    if (sa->h_out && sa->h_ld < cols) return skr::fail(SKR_ERR_ARG, "skr_stream_counts: host pitch < 4^k");
    cudaStream_t s_k = (cudaStream_t)stream_v;
    StreamRes* res = nullptr;
    int rc = get_streams(&res);
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
    const int64_t nblocks = skr_packed_num_blocks(packed);

    // chunks: a small first one so the pipeline fills quickly, then ~chunk_bytes of output each
    int64_t per = sa->chunk_records > 0 ? sa->chunk_records : std::max<int64_t>(256, ((int64_t)32 << 20) / (cols * 4));
    const bool ring = sa->h_out && !sa->h_out_pinned;
    if (ring) per = std::min<int64_t>(per, std::max<int64_t>(64, ((int64_t)8 << 20) / (cols * 4)));
    std::vector<Chunk> chunks;
    for (int64_t r = 0; r < m;) {
        int64_t n = chunks.empty() ? std::max<int64_t>(64, per / 4) : per;
        Chunk c;
        c.r0 = r;
        c.r1 = std::min(m, r + n);
        chunks.push_back(c);
        r = c.r1;
    }
    const int nchunks = (int)chunks.size();

    // the record table first: it is complete before any code word is packed
    SKR_CUDA_CHECK(cudaMemcpyAsync(d_slab + off_blk, h_blk, (size_t)(m + 1) * 8, cudaMemcpyHostToDevice, res->in));
    SKR_CUDA_CHECK(cudaMemcpyAsync(d_slab + off_len, h_len, (size_t)m * 4, cudaMemcpyHostToDevice, res->in));
    // trailing pad block (written by the allocator, not by the pack threads)
    SKR_CUDA_CHECK(cudaMemcpyAsync(d_slab + off_codes + (size_t)(nblocks - 1) * 16, (const char*)h_codes + (size_t)(nblocks - 1) * 16, 16,
                                   cudaMemcpyHostToDevice, res->in));
    SKR_CUDA_CHECK(cudaMemcpyAsync(d_slab + off_mask + (size_t)(nblocks - 1) * 8, (const char*)h_mask + (size_t)(nblocks - 1) * 8, 8,
                                   cudaMemcpyHostToDevice, res->in));
    // whatever the caller enqueued on its stream (vectors, the speculation cell) comes before the first kernel, and
    // the copy-in stream must not overwrite a slab an earlier launch on that stream still reads
    cudaEvent_t start_ev;
    SKR_CUDA_CHECK(cudaEventCreateWithFlags(&start_ev, cudaEventDisableTiming));
    SKR_CUDA_CHECK(cudaEventRecord(start_ev, s_k));
    SKR_CUDA_CHECK(cudaStreamWaitEvent(res->in, start_ev, 0));
    SKR_CUDA_CHECK(cudaStreamWaitEvent(res->out, start_ev, 0));

    // pinned ring for a pageable destination
    constexpr int kSlots = 4;
    size_t slot_bytes = 0;
    char* slots[kSlots] = {nullptr, nullptr, nullptr, nullptr};
    if (ring) {
        int64_t max_rows = 0;
        for (auto& c : chunks) max_rows = std::max(max_rows, c.r1 - c.r0);
        slot_bytes = (size_t)max_rows * cols * 4;
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
    std::vector<std::atomic<int>> issued(nchunks), drained(nchunks);
    for (int i = 0; i < nchunks; ++i) { issued[i].store(0); drained[i].store(0); }
    std::vector<std::thread> copiers;
    const int dev = res->dev;
    if (ring) {
        int nthreads = sa->copy_threads > 0 ? sa->copy_threads : 4;
        for (int t = 0; t < nthreads; ++t) {
            copiers.emplace_back([&, dev] {
                cudaSetDevice(dev);
                for (;;) {
                    const int i = copy_next.fetch_add(1);
                    if (i >= nchunks) break;
                    while (!issued[i].load(std::memory_order_acquire)) {
                        if (copy_fail.load()) return;
                        std::this_thread::yield();
                    }
                    const Chunk& c = chunks[i];
                    if (cudaEventSynchronize(c.out_done) != cudaSuccess) { copy_fail.store(1); drained[i].store(1); return; }
                    const size_t row = (size_t)cols * 4;
                    const char* src = slots[c.slot];
                    char* dst = (char*)sa->h_out + (size_t)c.r0 * (size_t)sa->h_ld * 4;
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
    for (int i = 0; i < nchunks && status == SKR_OK; ++i) {
        Chunk& c = chunks[i];
        cudaEventCreateWithFlags(&c.in_done, cudaEventDisableTiming);
        cudaEventCreateWithFlags(&c.k_done, cudaEventDisableTiming);
        cudaEventCreateWithFlags(&c.out_done, cudaEventDisableTiming);
        // ---- wait for the packer, then copy the chunk's code and mask words in
        skr_packed_wait_records(packed, c.r1);
        const uint64_t b0 = h_blk[c.r0], b1 = h_blk[c.r1];
        cudaError_t e = cudaSuccess;
        if (b1 > b0) {
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
                e = cudaMemcpy2DAsync((char*)sa->h_out + (size_t)c.r0 * (size_t)sa->h_ld * 4, (size_t)sa->h_ld * 4, src,
                                      (size_t)sa->count.ld_out * 4, row, rows, cudaMemcpyDeviceToHost, res->out);
            }
            if (e == cudaSuccess) e = cudaEventRecord(c.out_done, res->out);
        }
        if (e != cudaSuccess) { status = skr::fail(SKR_ERR_CUDA, "skr_stream_counts: %s", cudaGetErrorString(e)); break; }
        issued[i].store(1, std::memory_order_release);
    }
    t_last_issue = now_ms();
    if (status != SKR_OK) copy_fail.store(1);
    for (auto& t : copiers) t.join();
    cudaError_t e1 = cudaStreamSynchronize(res->out);
    cudaError_t e2 = cudaStreamSynchronize(res->in);
    // the caller's stream continues after the last chunk; nothing later on it may race with our copy-out reads
    if (status == SKR_OK && !chunks.empty() && chunks.back().out_done && sa->h_out) cudaStreamWaitEvent(s_k, chunks.back().out_done, 0);
    for (auto& c : chunks) {
        if (c.in_done) cudaEventDestroy(c.in_done);
        if (c.k_done) cudaEventDestroy(c.k_done);
        if (c.out_done) cudaEventDestroy(c.out_done);
    }
    cudaEventDestroy(start_ev);
    for (int i = 0; i < kSlots; ++i) if (slots[i]) skr_host_free(slots[i]);
    if (status == SKR_OK && (e1 != cudaSuccess || e2 != cudaSuccess))
        status = skr::fail(SKR_ERR_CUDA, "skr_stream_counts: %s", cudaGetErrorString(e1 != cudaSuccess ? e1 : e2));
    if (status == SKR_OK && copy_fail.load()) status = skr::fail(SKR_ERR_CUDA, "skr_stream_counts: copy-out failed");
    if (profile)
        fprintf(stderr, "skr_stream: %d chunks (%s destination), first kernel issued at %.2f ms, last chunk issued at %.2f ms, done at %.2f ms\n",
                nchunks, ring ? "pageable, pinned ring" : (sa->h_out ? "pinned" : "no host"), t_first_kernel - t_begin, t_last_issue - t_begin,
                now_ms() - t_begin);
    return status;
}
