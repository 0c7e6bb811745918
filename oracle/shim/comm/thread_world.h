// oracle/shim/comm/thread_world.h -- TEST INFRASTRUCTURE. The "MPI world" of the compiled-in-place reference:
// every rank is a thread of this process; messages are pointers posted to a mailbox between two barriers.
#ifndef ORACLE_SHIM_THREAD_WORLD_H
#define ORACLE_SHIM_THREAD_WORLD_H
#include <condition_variable>
#include <mutex>
#include <vector>

namespace shim {
    struct Barrier {
        std::mutex m;
        std::condition_variable cv;
        int n = 1, waiting = 0;
        unsigned long gen = 0;
        void wait() {
            if (n <= 1) return;
            std::unique_lock<std::mutex> lk(m);
            const unsigned long g = gen;
            if (++waiting == n) { waiting = 0; gen++; cv.notify_all(); }
            else cv.wait(lk, [&] { return gen != g; });
        }
    };
    struct Slot { void *ptr; unsigned long count; };
    struct ThreadWorld {
        int n = 1;
        Barrier barrier;
        std::vector<Slot> box;        // [rank][dir]
        std::vector<double> reduce;   // [rank][16]
        explicit ThreadWorld(int ranks) : n(ranks), box(2 * ranks), reduce(16 * ranks) { barrier.n = ranks; }
    };
    extern thread_local int tl_rank;
    extern thread_local ThreadWorld *tl_world;
}
#endif
