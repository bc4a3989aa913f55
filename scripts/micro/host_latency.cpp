// Host micro-benchmark: dependent-load latency over a footprint like the Euler walk's row array (context for the
// per-step cost of the host walk: one dependent miss per step, prefetched up to two steps ahead).
#include <sys/mman.h>
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <numeric>
#include <random>
#include <vector>
static double now() { return std::chrono::duration<double, std::nano>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
int main() {
    for (size_t mib : {16, 64, 256, 1024}) {
        for (int huge = 1; huge >= 0; huge--) {
            const size_t bytes = mib << 20, n = bytes / 64;
            char* p = (char*)mmap(nullptr, bytes, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0);
            madvise(p, bytes, huge ? MADV_HUGEPAGE : MADV_NOHUGEPAGE);
            memset(p, 0, bytes);
            std::vector<uint32_t> perm(n);
            std::iota(perm.begin(), perm.end(), 0u);
            std::mt19937_64 rng(1);
            for (size_t i = n - 1; i > 0; i--) std::swap(perm[i], perm[rng() % (i + 1)]);
            for (size_t i = 0; i < n; i++) *(uint32_t*)(p + (size_t)perm[i] * 64) = perm[(i + 1) % n];  // one cycle over all lines
            const size_t hops = 4000000;
            uint32_t cur = perm[0];
            double t0 = now();
            for (size_t i = 0; i < hops; i++) cur = *(volatile uint32_t*)(p + (size_t)cur * 64);
            double t1 = now();
            printf("{\"footprint_MiB\": %zu, \"huge_pages\": %d, \"ns_per_dependent_load\": %.1f, \"sink\": %u}\n", mib, huge, (t1 - t0) / hops, cur & 1);
            munmap(p, bytes);
        }
    }
    return 0;
}
