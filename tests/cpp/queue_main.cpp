// The single-producer / single-consumer queue behind ModalSolver's messages, exercised from one source compiled twice by
// tests/test_host_logic.py: against the mirror's external/readerwriterqueue.h and against the reference's own (moodycamel),
// printing the same trace: capacity reached by try_enqueue for the sizes modal_solver.h uses (and a few more), FIFO order
// across wrap-around, size_approx, try_dequeue on empty.
#include <cstdio>
#include "readerwriterqueue.h"

int main() {
    const int sizes[] = {1, 2, 3, 4, 7, 8, 15, 16, 100, 512, 1023};
    for (int maxSize : sizes) {
        moodycamel::ReaderWriterQueue<int> q(maxSize);
        int pushed = 0;
        while (q.try_enqueue(pushed)) ++pushed;                       // never allocates: fails at the block's capacity
        printf("maxSize %d capacity %d size_approx %d\n", maxSize, pushed, (int)q.size_approx());
        long long sum = 0; int v = -1, popped = 0, next = pushed;
        for (int round = 0; round < 3 * pushed + 5; ++round) {        // interleave to wrap around several times
            if (q.try_dequeue(v)) { sum = sum * 31 + v; ++popped; }
            if (round % 3 != 2 && q.try_enqueue(next)) ++next;
        }
        while (q.try_dequeue(v)) { sum = (sum * 31 + v) % 1000000007LL; ++popped; }
        printf("  popped %d next %d checksum %lld empty_dequeue %d size_approx %d\n", popped, next, sum % 1000000007LL, (int)q.try_dequeue(v), (int)q.size_approx());
    }
    return 0;
}
