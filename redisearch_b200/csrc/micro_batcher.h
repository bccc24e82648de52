// Combine concurrent single-query callers into one batched call (leader / follower).
//
// RediSearch's worker threads call VecSimIndex_TopKQuery one query at a time (src/util/workers.c,
// src/iterators/hybrid_reader.c:374); on the device one corpus pass serves hundreds of queries for the price of one
// (DESIGN.md §4).  The first caller to arrive becomes the leader of a batch: it waits until `max_batch` requests have
// gathered or `window` has passed since it arrived, takes the requests, lets the next arrival lead the next batch, runs
// the batch function outside the lock and wakes its followers.  Plain C++: no CUDA in here, unit-tested on the host
// (tests/test_micro_batcher.py).
#pragma once
#include <chrono>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <vector>

namespace rsb200 {

template <class Req>
class MicroBatcher {
  public:
    using BatchFn = std::function<void(std::vector<Req *> &)>;

    MicroBatcher(size_t max_batch, std::chrono::microseconds window, BatchFn fn, std::function<void(Req &)> on_error = nullptr)
        : max_batch_(max_batch ? max_batch : 1), window_(window), fn_(std::move(fn)), on_error_(std::move(on_error)) {}

    // Blocks until `r` has been processed by some batch (possibly one this thread led).
    void submit(Req &r) {
        std::unique_lock<std::mutex> lk(mu_);
        Ticket t{&r, false, false};
        pending_.push_back(&t);
        if (pending_.size() >= max_batch_) arrivals_.notify_all(); // a waiting leader may go now
        while (!t.done) {
            if (t.taken) { // in a batch some leader is executing right now
                done_.wait(lk, [&] { return t.done; });
                break;
            }
            if (leader_active_) { // follower: wait to be taken, or for the leadership to become vacant
                done_.wait(lk, [&] { return t.taken || t.done || !leader_active_; });
                continue;
            }
            // leader of the next batch
            leader_active_ = true;
            const auto deadline = std::chrono::steady_clock::now() + window_;
            arrivals_.wait_until(lk, deadline, [&] { return pending_.size() >= max_batch_; });
            const size_t take = std::min(pending_.size(), max_batch_);
            std::vector<Ticket *> batch(pending_.begin(), pending_.begin() + take);
            pending_.erase(pending_.begin(), pending_.begin() + take);
            for (Ticket *b : batch) b->taken = true;
            leader_active_ = false; // whoever is still pending (or arrives next) elects the next leader
            done_.notify_all();
            if (batch.empty()) continue;
            lk.unlock();
            std::vector<Req *> reqs;
            reqs.reserve(batch.size());
            for (Ticket *b : batch) reqs.push_back(b->req);
            bool threw = false;
            try {
                fn_(reqs);
            } catch (...) { // followers must never be left waiting: mark the batch done, report through on_error
                threw = true;
            }
            lk.lock();
            if (threw && on_error_)
                for (Ticket *b : batch) on_error_(*b->req);
            for (Ticket *b : batch) b->done = true;
            batches_++;
            requests_ += batch.size();
            done_.notify_all();
        }
    }

    // statistics (under the lock)
    void stats(size_t *batches, size_t *requests) {
        std::lock_guard<std::mutex> g(mu_);
        *batches = batches_;
        *requests = requests_;
    }

  private:
    struct Ticket {
        Req *req;
        bool taken, done;
    };
    const size_t max_batch_;
    const std::chrono::microseconds window_;
    BatchFn fn_;
    std::function<void(Req &)> on_error_; // called (under the lock) for every request of a batch whose fn_ threw
    std::mutex mu_;
    std::condition_variable arrivals_, done_;
    std::vector<Ticket *> pending_;
    bool leader_active_ = false;
    size_t batches_ = 0, requests_ = 0;
};

} // namespace rsb200
