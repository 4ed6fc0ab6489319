// experiment: cost of pinning host memory -- cudaHostAlloc vs transparent-huge-page backed mmap + cudaHostRegister
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <sys/mman.h>
#include <cuda_runtime.h>
static double ms(std::chrono::steady_clock::time_point a) { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - a).count(); }
int main()
{
    cudaFree(nullptr);
    const size_t n = 128u << 20;
    for (int rep = 0; rep < 3; rep++) {
        void* p = nullptr;
        auto t0 = std::chrono::steady_clock::now();
        cudaHostAlloc(&p, n, cudaHostAllocDefault);
        printf("cudaHostAlloc 128 MB: %.1f ms\n", ms(t0));
        t0 = std::chrono::steady_clock::now();
        cudaFreeHost(p);
        printf("  cudaFreeHost: %.1f ms\n", ms(t0));
        for (int huge = 0; huge < 2; huge++) {
            t0 = std::chrono::steady_clock::now();
            void* q = mmap(nullptr, n + (2u << 20), PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0);
            char* a = (char*)(((uintptr_t)q + (2u << 20) - 1) & ~(uintptr_t)((2u << 20) - 1));
            if (huge) madvise(a, n, MADV_HUGEPAGE);
            memset(a, 0, n);
            const double t_touch = ms(t0);
            t0 = std::chrono::steady_clock::now();
            cudaError_t e = cudaHostRegister(a, n, cudaHostRegisterDefault);
            printf("mmap%s + touch %.1f ms, cudaHostRegister %.1f ms (%s)\n", huge ? " + MADV_HUGEPAGE" : "", t_touch, ms(t0), cudaGetErrorString(e));
            void* d = nullptr; cudaMalloc(&d, n);
            cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
            cudaEventRecord(e0); cudaMemcpyAsync(d, a, n, cudaMemcpyHostToDevice); cudaEventRecord(e1); cudaEventSynchronize(e1);
            float t; cudaEventElapsedTime(&t, e0, e1);
            printf("  H2D from it: %.1f GB/s\n", n / 1e6 / t);
            t0 = std::chrono::steady_clock::now();
            cudaHostUnregister(a); munmap(q, n + (2u << 20));
            printf("  unregister + munmap: %.1f ms\n", ms(t0));
            cudaFree(d);
        }
    }
    system("cat /sys/kernel/mm/transparent_hugepage/enabled");
    return 0;
}
