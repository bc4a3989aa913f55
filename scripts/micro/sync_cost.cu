// Micro-benchmark: cost of reading one device scalar on the host (diagnostic for the small-input step).
#include <cuda_runtime.h>
#include <chrono>
#include <cstdio>
__global__ void bump(unsigned* p) { *p += 1; }
__global__ void bump_mapped(unsigned* p, volatile unsigned* h) { *p += 1; *h = *p; }
static double now() { return std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
int main() {
    cudaStream_t s; cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking);
    unsigned* d; cudaMalloc(&d, 4); cudaMemset(d, 0, 4);
    unsigned pageable = 0, *pinned, *mapped, *mapped_d;
    cudaHostAlloc(&pinned, 4, cudaHostAllocDefault);
    cudaHostAlloc(&mapped, 4, cudaHostAllocMapped); cudaHostGetDevicePointer(&mapped_d, mapped, 0);
    const int R = 2000;
    for (int variant = 0; variant < 5; variant++) {
        for (int w = 0; w < 2; w++) {
            double t0 = now();
            for (int i = 0; i < R; i++) {
                switch (variant) {
                    case 0: bump<<<1, 1, 0, s>>>(d); cudaStreamSynchronize(s); break;
                    case 1: bump<<<1, 1, 0, s>>>(d); cudaMemcpyAsync(&pageable, d, 4, cudaMemcpyDeviceToHost, s); cudaStreamSynchronize(s); break;
                    case 2: bump<<<1, 1, 0, s>>>(d); cudaMemcpyAsync(pinned, d, 4, cudaMemcpyDeviceToHost, s); cudaStreamSynchronize(s); break;
                    case 3: bump_mapped<<<1, 1, 0, s>>>(d, mapped_d); cudaStreamSynchronize(s); break;
                    case 4: for (int j = 0; j < 4; j++) bump<<<1, 1, 0, s>>>(d); cudaStreamSynchronize(s); break;
                }
            }
            double t1 = now();
            const char* names[] = {"kernel+sync", "kernel+D2H pageable+sync", "kernel+D2H pinned+sync", "kernel writes mapped+sync", "4 kernels+sync"};
            if (w) printf("%-28s %.2f us/iter\n", names[variant], (t1 - t0) / R);
        }
    }
    return 0;
}
