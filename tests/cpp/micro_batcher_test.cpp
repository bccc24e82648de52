// Host-only unit test of redisearch_b200/csrc/micro_batcher.h: many threads submit, every request is processed exactly
// once with the right answer, batches never exceed max_batch, and concurrent arrivals really are combined.
#include "../../redisearch_b200/csrc/micro_batcher.h"

#include <atomic>
#include <cstdio>
#include <thread>

struct Req {
    int in;
    int out;
    int processed;
};

int main() {
    using namespace rsb200;
    int failures = 0;
    for (int round = 0; round < 3; round++) {
        const size_t max_batch = round == 0 ? 16 : (round == 1 ? 1 : 256);
        std::atomic<size_t> largest{0};
        MicroBatcher<Req> mb(max_batch, std::chrono::microseconds(round == 1 ? 0 : 3000), [&](std::vector<Req *> &b) {
            size_t cur = largest.load();
            while (b.size() > cur && !largest.compare_exchange_weak(cur, b.size())) {
            }
            std::this_thread::sleep_for(std::chrono::microseconds(200)); // the "device" is busy for a while
            for (Req *r : b) {
                r->out = r->in * 3 + 1;
                r->processed++;
            }
        });
        const int T = 64, PER = 20;
        std::vector<std::thread> th;
        std::atomic<int> bad{0};
        std::atomic<bool> go{false};
        for (int t = 0; t < T; t++)
            th.emplace_back([&, t] {
                while (!go.load()) std::this_thread::yield();
                for (int i = 0; i < PER; i++) {
                    Req r{t * 1000 + i, 0, 0};
                    mb.submit(r);
                    if (r.out != r.in * 3 + 1 || r.processed != 1) bad++;
                }
            });
        go = true;
        for (auto &x : th) x.join();
        size_t batches = 0, requests = 0;
        mb.stats(&batches, &requests);
        const bool ok = bad == 0 && requests == (size_t)T * PER && largest <= max_batch && batches <= requests &&
                        (max_batch == 1 ? batches == requests : batches < requests);
        printf("round %d: max_batch %zu -> %zu requests in %zu batches, largest %zu, bad %d: %s\n", round, max_batch, requests, batches,
               largest.load(), bad.load(), ok ? "ok" : "FAIL");
        failures += !ok;
    }
    return failures;
}
