// Host-side half of the input path: a persistent worker pool that assembles the rows of the next
// minibatches into pinned staging buffers, so the copy engine can move each batch to HBM as ONE
// contiguous DMA while the SMs run the step.
//
// Replaces the reference's per-sample Python __getitem__ + transform + default_collate
// (reference solver_worker.py:805-832, transform.py:25-38): here the only per-sample host work is a
// memcpy of the raw row, done by native threads outside the GIL; the per-sample arithmetic runs on
// the device afterwards (K5).  Why not let the GPU gather over PCIe (K8)?  It can, and that path
// stays: but any CTA that sits on an SM for the ~1.3 ms a 67 MB batch needs on PCIe costs the
// step's cluster-scheduled GEMMs far more than its share of SMs (measured: 4 CTAs -> GEMMs +35 %).
// The DMA engines cost the SMs nothing.
//
// Jobs are FIFO; a job is split into chunks of rows that workers claim with an atomic counter.
// Stores to the staging buffer are non-temporal (no read-for-ownership traffic, the CPU never
// reads the staging buffer back) and fenced before the job is reported complete.
#include <emmintrin.h>

#include <atomic>
#include <condition_variable>
#include <cstring>
#include <deque>
#include <memory>
#include <mutex>
#include <thread>
#include <vector>

#include "frl_common.cuh"

namespace frl {

static inline void copy_row_nt(uint8_t* __restrict__ d, const uint8_t* __restrict__ s, int64_t bytes) {
    if (((reinterpret_cast<uintptr_t>(d) | static_cast<uintptr_t>(bytes)) & 15u) != 0) {
        memcpy(d, s, static_cast<size_t>(bytes));
        return;
    }
    int64_t i = 0;
    for (; i + 64 <= bytes; i += 64) {
        const __m128i a = _mm_loadu_si128(reinterpret_cast<const __m128i*>(s + i));
        const __m128i b = _mm_loadu_si128(reinterpret_cast<const __m128i*>(s + i + 16));
        const __m128i c = _mm_loadu_si128(reinterpret_cast<const __m128i*>(s + i + 32));
        const __m128i e = _mm_loadu_si128(reinterpret_cast<const __m128i*>(s + i + 48));
        _mm_stream_si128(reinterpret_cast<__m128i*>(d + i), a);
        _mm_stream_si128(reinterpret_cast<__m128i*>(d + i + 16), b);
        _mm_stream_si128(reinterpret_cast<__m128i*>(d + i + 32), c);
        _mm_stream_si128(reinterpret_cast<__m128i*>(d + i + 48), e);
    }
    for (; i < bytes; i += 16)
        _mm_stream_si128(reinterpret_cast<__m128i*>(d + i),
                         _mm_loadu_si128(reinterpret_cast<const __m128i*>(s + i)));
}

struct GatherJob {
    const uint8_t* src;
    uint8_t* dst;
    std::vector<int64_t> idx;          // private copy: the caller's index buffer may be reused
    int64_t row_bytes;
    int64_t rows_per_chunk;
    int64_t n_chunks;
    std::atomic<int64_t> next{0};
    std::atomic<int64_t> done{0};
    int64_t ticket;
};

}  // namespace frl

struct frl_gather_pool {
    std::mutex mu;
    std::condition_variable work_cv, done_cv;
    std::deque<std::shared_ptr<frl::GatherJob>> queue;     // jobs with unclaimed chunks
    std::vector<std::thread> workers;
    int64_t last_ticket = 0;
    int64_t completed_upto = 0;                            // every ticket <= this is complete
    std::vector<int64_t> completed_out_of_order;
    bool stop = false;

    void run() {
        for (;;) {
            std::shared_ptr<frl::GatherJob> job;
            int64_t chunk = -1;
            {
                std::unique_lock<std::mutex> lk(mu);
                for (;;) {
                    while (!queue.empty()) {
                        auto& front = queue.front();
                        const int64_t c = front->next.fetch_add(1, std::memory_order_relaxed);
                        if (c < front->n_chunks) { job = front; chunk = c; break; }
                        queue.pop_front();                 // fully claimed
                    }
                    if (job || stop) break;
                    work_cv.wait(lk);
                }
                if (!job) return;
            }
            // keep claiming chunks of this job without the lock
            for (;;) {
                const int64_t n = static_cast<int64_t>(job->idx.size());
                const int64_t lo = chunk * job->rows_per_chunk;
                int64_t hi = lo + job->rows_per_chunk;
                if (hi > n) hi = n;
                for (int64_t i = lo; i < hi; ++i)
                    frl::copy_row_nt(job->dst + i * job->row_bytes, job->src + job->idx[i] * job->row_bytes,
                                     job->row_bytes);
                _mm_sfence();
                const int64_t finished = job->done.fetch_add(1, std::memory_order_acq_rel) + 1;
                if (finished == job->n_chunks) {
                    std::lock_guard<std::mutex> lk(mu);
                    completed_out_of_order.push_back(job->ticket);
                    bool advanced = true;
                    while (advanced) {
                        advanced = false;
                        for (size_t k = 0; k < completed_out_of_order.size(); ++k)
                            if (completed_out_of_order[k] == completed_upto + 1) {
                                ++completed_upto;
                                completed_out_of_order.erase(completed_out_of_order.begin() + k);
                                advanced = true;
                                break;
                            }
                    }
                    done_cv.notify_all();
                }
                chunk = job->next.fetch_add(1, std::memory_order_relaxed);
                if (chunk >= job->n_chunks) break;
            }
        }
    }
};

extern "C" frl_gather_pool* frl_gather_pool_create(int n_threads) {
    if (n_threads < 1) n_threads = 1;
    if (n_threads > 256) n_threads = 256;
    frl_gather_pool* p = new (std::nothrow) frl_gather_pool();
    if (!p) { frl::set_error("frl_gather_pool_create: out of memory"); return nullptr; }
    try {
        for (int t = 0; t < n_threads; ++t) p->workers.emplace_back([p] { p->run(); });
    } catch (...) {
        frl::set_error("frl_gather_pool_create: cannot start %d threads", n_threads);
        {
            std::lock_guard<std::mutex> lk(p->mu);
            p->stop = true;
        }
        p->work_cv.notify_all();
        for (auto& th : p->workers) th.join();
        delete p;
        return nullptr;
    }
    return p;
}

extern "C" void frl_gather_pool_destroy(frl_gather_pool* p) {
    if (!p) return;
    {
        std::lock_guard<std::mutex> lk(p->mu);
        p->stop = true;
    }
    p->work_cv.notify_all();
    for (auto& th : p->workers) th.join();
    delete p;
}

extern "C" int frl_gather_pool_threads(const frl_gather_pool* p) {
    return p ? static_cast<int>(p->workers.size()) : 0;
}

extern "C" int64_t frl_gather_pool_submit(frl_gather_pool* p, const void* src_host, int64_t src_rows,
                                          const int64_t* idx_host, void* dst_host, int64_t n_rows,
                                          int64_t row_bytes) {
    if (!p || n_rows < 0 || row_bytes < 0 || src_rows < 1 || (n_rows > 0 && (!src_host || !idx_host || !dst_host))) {
        frl::set_error("frl_gather_pool_submit: bad arguments");
        return FRL_E_ARG;
    }
    for (int64_t i = 0; i < n_rows; ++i)
        if (idx_host[i] < 0 || idx_host[i] >= src_rows) {
            frl::set_error("frl_gather_pool_submit: index %lld at position %lld outside [0, %lld)",
                           static_cast<long long>(idx_host[i]), static_cast<long long>(i),
                           static_cast<long long>(src_rows));
            return FRL_E_ARG;
        }
    auto job = std::make_shared<frl::GatherJob>();
    job->src = static_cast<const uint8_t*>(src_host);
    job->dst = static_cast<uint8_t*>(dst_host);
    job->idx.assign(idx_host, idx_host + n_rows);
    job->row_bytes = row_bytes;
    // ~256 KB of rows per chunk: fine-grained enough to balance, coarse enough to amortise the atomics
    int64_t rpc = row_bytes > 0 ? (256 * 1024) / row_bytes : n_rows;
    if (rpc < 1) rpc = 1;
    job->rows_per_chunk = rpc;
    job->n_chunks = (n_rows + rpc - 1) / rpc;
    std::lock_guard<std::mutex> lk(p->mu);
    job->ticket = ++p->last_ticket;
    if (job->n_chunks == 0 || row_bytes == 0) {
        job->n_chunks = 0;
        p->completed_out_of_order.push_back(job->ticket);
        bool advanced = true;
        while (advanced) {
            advanced = false;
            for (size_t k = 0; k < p->completed_out_of_order.size(); ++k)
                if (p->completed_out_of_order[k] == p->completed_upto + 1) {
                    ++p->completed_upto;
                    p->completed_out_of_order.erase(p->completed_out_of_order.begin() + k);
                    advanced = true;
                    break;
                }
        }
        p->done_cv.notify_all();
        return job->ticket;
    }
    p->queue.push_back(job);
    p->work_cv.notify_all();
    return job->ticket;
}

extern "C" int frl_gather_pool_wait(frl_gather_pool* p, int64_t ticket) {
    if (!p || ticket < 1) {
        frl::set_error("frl_gather_pool_wait: bad arguments");
        return FRL_E_ARG;
    }
    std::unique_lock<std::mutex> lk(p->mu);
    if (ticket > p->last_ticket) {
        frl::set_error("frl_gather_pool_wait: ticket %lld was never issued", static_cast<long long>(ticket));
        return FRL_E_ARG;
    }
    p->done_cv.wait(lk, [&] { return p->completed_upto >= ticket; });
    return 0;
}
