// Host micro-benchmark behind the design of the Euler walk's records: a dependent chain over random 64-byte lines (one
// line per step, like the walk), where every line also names the line `depth` steps ahead so that it can be prefetched.
// Prints ns per step for depth 0 (pure dependent misses) .. 8: the curve the walk's lookahead depth is chosen from.
#include <sys/mman.h>
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <numeric>
#include <random>
#include <vector>
static double now() { return std::chrono::duration<double, std::nano>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
struct alignas(64) Line {
    uint32_t next, ahead[8], pad[7];
};
int main(int argc, char** argv) {
    const size_t mib = argc > 1 ? atoi(argv[1]) : 512;
    const size_t n = (mib << 20) / sizeof(Line);
    Line* p = (Line*)mmap(nullptr, n * sizeof(Line), PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0);
    madvise(p, n * sizeof(Line), MADV_HUGEPAGE);
    memset(p, 0, n * sizeof(Line));
    std::vector<uint32_t> perm(n);
    std::iota(perm.begin(), perm.end(), 0u);
    std::mt19937_64 rng(1);
    for (size_t i = n - 1; i > 0; i--) std::swap(perm[i], perm[rng() % (i + 1)]);
    for (size_t i = 0; i < n; i++) {
        p[perm[i]].next = perm[(i + 1) % n];
        for (int d = 0; d < 8; d++) p[perm[i]].ahead[d] = perm[(i + d + 2) % n];  // ahead[d]: the line d + 2 steps from here
    }
    const size_t hops = 4000000;
    for (int depth = 0; depth <= 8; depth++) {
        uint32_t cur = perm[0];
        uint64_t sink = 0;
        double t0 = now();
        for (size_t i = 0; i < hops; i++) {
            const Line& l = p[cur];
            if (depth >= 2) __builtin_prefetch(&p[l.ahead[depth - 2]]);
            else if (depth == 1) __builtin_prefetch(&p[l.next]);
            sink += l.pad[0];
            cur = l.next;
        }
        double t1 = now();
        printf("{\"footprint_MiB\": %zu, \"lookahead_steps\": %d, \"ns_per_step\": %.1f, \"sink\": %llu}\n", mib, depth, (t1 - t0) / hops,
               (unsigned long long)(sink + cur) & 1);
    }
    return 0;
}
