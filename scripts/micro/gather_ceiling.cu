// Micro-benchmark: ceiling for random 32-byte-sector gathers, the access pattern of the bounded
// Dijkstra (row_ptr pair, target bit, col/weight row of a settled node).  Two footprints: far larger than L2
// (HBM random-access ceiling) and L2-resident (the regime of the E. coli-size configs); two access shapes:
// independent gathers (all memory-level parallelism the SMs can hold) and dependent chains of length 8
// (each search is a chain: the next row is known only after the previous one arrived).
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <vector>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)
__device__ __forceinline__ uint64_t mix(uint64_t z) {
    z += 0x9E3779B97F4A7C15ull; z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull; z = (z ^ (z >> 27)) * 0x94D049BB133111EBull; return z ^ (z >> 31);
}
__global__ void fill(uint32_t* a, uint64_t n_words, uint64_t n_sectors) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_words; i += (uint64_t)gridDim.x * blockDim.x)
        a[i] = (uint32_t)(mix(i) % n_sectors);  // every word names a random sector
}
template <int CHAIN>
__global__ void gather(const uint32_t* __restrict__ a, uint64_t n_sectors, uint64_t per_thread, uint32_t* out) {
    const uint64_t tid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t acc = 0;
    for (uint64_t r = 0; r < per_thread; r++) {
        uint32_t s = (uint32_t)(mix(tid * per_thread + r) % n_sectors);
#pragma unroll
        for (int c = 0; c < CHAIN; c++) s = a[(uint64_t)s * 8 + (c & 7)];  // next sector comes from the loaded word
        acc ^= s;
    }
    if (acc == 0xFFFFFFFFu) out[0] = acc;
}
template <int CHAIN>
static int run(const char* name, const uint32_t* a, uint64_t n_sectors, uint32_t* out) {
    const int threads = 256, blocks = 148 * 8;
    const uint64_t per_thread = 64;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int w = 0; w < 2; w++) gather<CHAIN><<<blocks, threads>>>(a, n_sectors, per_thread, out);
    CK(cudaEventRecord(e0));
    const int reps = 5;
    for (int w = 0; w < reps; w++) gather<CHAIN><<<blocks, threads>>>(a, n_sectors, per_thread, out);
    CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double sectors = (double)reps * blocks * threads * per_thread * CHAIN;
    printf("{\"case\": \"%s\", \"chain\": %d, \"footprint_MiB\": %.0f, \"Gsectors_per_s\": %.2f, \"GBps_32B_sectors\": %.1f}\n", name, CHAIN,
           n_sectors * 32.0 / (1 << 20), sectors / (ms * 1e-3) / 1e9, sectors * 32 / (ms * 1e-3) / 1e9);
    return 0;
}
int main() {
    uint32_t* out; CK(cudaMalloc(&out, 4));
    for (uint64_t mib : {8192ull, 32ull}) {
        const uint64_t n_sectors = mib * (1 << 20) / 32, n_words = n_sectors * 8;
        uint32_t* a; CK(cudaMalloc(&a, n_words * 4));
        fill<<<148 * 16, 256>>>(a, n_words, n_sectors); CK(cudaDeviceSynchronize());
        const char* name = mib > 1000 ? "hbm_random" : "l2_resident";
        if (run<1>(name, a, n_sectors, out)) return 1;
        if (run<8>(name, a, n_sectors, out)) return 1;
        CK(cudaFree(a));
    }
    return 0;
}
