// stream_b200.hpp -- file <-> pinned host memory on helper threads, for the servers ("streaming
// chunk scheduler overlapping file/PCIe I/O with compute").
//
// The reference's servers read a ciphertext, compute, write a ciphertext (server_jpeg.cpp:113-153,
// fhe_resize.h:333-391).  On a GPU the three activities have different owners -- the file system and the record
// parser (CPU), the copy engines, the SMs -- so the servers run them as a pipeline over UNITS (a chunk of tiles,
// a row of pixels):
//   Reader thread:  parses unit i + 1, i + 2, ... of the input stream into a ring of pinned buffers
//   main thread:    fhe_upload_async(unit i + 1) | kernels(unit i) | fhe_download_async(unit i - 1)
//   Writer thread:  serialises unit i - 2, ... from a second ring of pinned buffers into the output stream
// A pinned buffer returns to its ring when the transfer that used it has completed (fhe_upload_fence /
// fhe_transfer_fence on the main thread).  Errors on a helper thread are rethrown on the main thread.
#pragma once
#include <atomic>
#include <condition_variable>
#include <cstdint>
#include <deque>
#include <exception>
#include <functional>
#include <mutex>
#include <stdexcept>
#include <string>
#include <thread>
#include <vector>

#include <fcntl.h>
#include <unistd.h>

#include "../../include/fhe_b200.h"

namespace stream_b200 {

struct Slot {
    uint64_t *host = nullptr;  // pinned, `words` u64
    size_t words = 0;
    long unit = -1;            // index of the unit it holds
    size_t count = 0;          // ciphertexts (or whatever the filler counts) in it
    uint32_t size = 0;         // free for the user (ciphertext size of an output unit)
};

class SlotQueue {
public:
    void push(Slot *s)
    {
        {
            std::lock_guard<std::mutex> g(m_);
            q_.push_back(s);
        }
        cv_.notify_all();
    }
    // null when closed and empty
    Slot *pop()
    {
        std::unique_lock<std::mutex> g(m_);
        cv_.wait(g, [&] { return !q_.empty() || closed_; });
        if (q_.empty()) return nullptr;
        Slot *s = q_.front();
        q_.pop_front();
        return s;
    }
    void close()
    {
        {
            std::lock_guard<std::mutex> g(m_);
            closed_ = true;
        }
        cv_.notify_all();
    }

private:
    std::mutex m_;
    std::condition_variable cv_;
    std::deque<Slot *> q_;
    bool closed_ = false;
};

class Ring {
protected:
    Ring(size_t nslots, size_t words) : slots_(nslots)
    {
        for (Slot &s : slots_) {
            void *p = nullptr;
            if (fhe_host_alloc(words * sizeof(uint64_t), &p) != FHE_OK) {
                const std::string why = fhe_last_error();
                for (Slot &t : slots_)  // the destructor does not run for a half-built ring
                    if (t.host) fhe_host_free(t.host);
                throw std::runtime_error(why);
            }
            s.host = (uint64_t *)p;
            s.words = words;
            free_.push(&s);
        }
    }
    ~Ring()
    {
        for (Slot &s : slots_)
            if (s.host) fhe_host_free(s.host);
    }
    void rethrow()
    {
        std::lock_guard<std::mutex> g(em_);
        if (error_) std::rethrow_exception(error_);
    }
    void fail(std::exception_ptr e)
    {
        std::lock_guard<std::mutex> g(em_);
        if (!error_) error_ = e;
    }
    std::vector<Slot> slots_;
    SlotQueue free_, ready_;
    std::mutex em_;
    std::exception_ptr error_;
    std::thread th_;
};

// fill(unit, slot) parses the next unit into slot.host and returns false at the end of the stream
class Reader : public Ring {
public:
    Reader(size_t nslots, size_t words, std::function<bool(long, Slot &)> fill) : Ring(nslots, words)
    {
        th_ = std::thread([this, fill] {
            try {
                for (long u = 0;; ++u) {
                    Slot *s = free_.pop();
                    if (!s || stop_) break;
                    s->unit = u;
                    if (!fill(u, *s)) break;
                    ready_.push(s);
                }
            } catch (...) {
                fail(std::current_exception());
            }
            ready_.close();
        });
    }
    ~Reader()
    {
        stop_ = true;  // do not parse further units into the buffers that are still free
        free_.close();
        if (th_.joinable()) th_.join();
    }
    // next filled unit, in order; null at the end of the stream
    Slot *next()
    {
        Slot *s = ready_.pop();
        rethrow();
        return s;
    }
    // the upload that read `s` has completed
    void recycle(Slot *s) { free_.push(s); }

private:
    std::atomic<bool> stop_{false};
};

// drain(slot) serialises one unit; units are drained in the order they are submitted
class Writer : public Ring {
public:
    Writer(size_t nslots, size_t words, std::function<void(Slot &)> drain) : Ring(nslots, words)
    {
        th_ = std::thread([this, drain] {
            try {
                while (Slot *s = ready_.pop()) {
                    drain(*s);
                    free_.push(s);
                }
            } catch (...) {
                fail(std::current_exception());
                while (Slot *s = ready_.pop()) free_.push(s);  // keep the producer from blocking
            }
        });
    }
    ~Writer()
    {
        ready_.close();
        if (th_.joinable()) th_.join();
    }
    // a buffer to download the next unit into (blocks while every buffer is waiting to be written)
    Slot *acquire()
    {
        Slot *s = free_.pop();
        rethrow();
        return s;
    }
    // the download into `s` has landed
    void submit(Slot *s) { ready_.push(s); }
    // everything submitted has been written
    void finish()
    {
        ready_.close();
        if (th_.joinable()) th_.join();
        rethrow();
    }
};

// ---------------------------------------------------------------------------------------------------------------
// Fixed-size records of a file, several threads at once.  A ciphertext stream is a concatenation of records of one
// size (seal_wire_b200.hpp record_bytes), so record i of a unit sits at a computable offset: the helper threads
// pread / pwrite disjoint ranges and convert them in parallel -- the single-threaded parser managed 0.35 GB/s,
// two orders of magnitude below what the device consumes.
// ---------------------------------------------------------------------------------------------------------------
inline unsigned io_threads()
{
    const unsigned hc = std::thread::hardware_concurrency();
    return hc >= 16 ? 8 : hc >= 4 ? hc / 2 : 1;
}
// body(begin, end) on `threads` threads over [0, count); the first exception is rethrown
inline void parallel_ranges(size_t count, unsigned threads, const std::function<void(size_t, size_t)> &body)
{
    if (threads > count) threads = (unsigned)(count ? count : 1);
    if (threads <= 1) {
        body(0, count);
        return;
    }
    std::vector<std::thread> pool;
    std::mutex m;
    std::exception_ptr err;
    for (unsigned t = 0; t < threads; ++t)
        pool.emplace_back([&, t] {
            try {
                body(count * t / threads, count * (t + 1) / threads);
            } catch (...) {
                std::lock_guard<std::mutex> g(m);
                if (!err) err = std::current_exception();
            }
        });
    for (auto &th : pool) th.join();
    if (err) std::rethrow_exception(err);
}
// pread of exactly `bytes` at `offset`; false at a clean end of file (nothing read), throws on a short record
inline bool pread_exact(int fd, void *buf, size_t bytes, uint64_t offset)
{
    size_t got = 0;
    while (got < bytes) {
        const ssize_t r = ::pread(fd, (char *)buf + got, bytes - got, (off_t)(offset + got));
        if (r < 0) throw std::runtime_error("ciphertext stream: read failed");
        if (r == 0) {
            if (got == 0) return false;
            throw std::invalid_argument("ciphertext stream: truncated record");
        }
        got += (size_t)r;
    }
    return true;
}
inline void pwrite_exact(int fd, const void *buf, size_t bytes, uint64_t offset)
{
    size_t put = 0;
    while (put < bytes) {
        const ssize_t r = ::pwrite(fd, (const char *)buf + put, bytes - put, (off_t)(offset + put));
        if (r <= 0) throw std::runtime_error("ciphertext stream: write failed");
        put += (size_t)r;
    }
}

}  // namespace stream_b200
