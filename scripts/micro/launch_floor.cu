// Microbenchmark: floor of a dependent-launch chain (with/without programmatic dependent launch) versus a
// persistent cooperative kernel with a global-memory grid barrier.  Build: nvcc -O3 -arch=sm_100a -o launch_floor launch_floor.cu
#include <cstdio>
#include <cuda_runtime.h>
#include <cooperative_groups.h>

__global__ void k_empty(double* buf, int pdl)
{
    if (pdl) {
        asm volatile("griddepcontrol.launch_dependents;");
        asm volatile("griddepcontrol.wait;" ::: "memory");
    }
    if (buf == nullptr) return;
}

// one global round trip + one atomic per thread-group, like the small-t symv
__global__ void k_work(double* buf, double* y, int pdl)
{
    if (pdl) {
        asm volatile("griddepcontrol.launch_dependents;");
        asm volatile("griddepcontrol.wait;" ::: "memory");
    }
    const size_t base = ((size_t)blockIdx.x * blockDim.x + threadIdx.x);
    double s = 0.0;
#pragma unroll
    for (int q = 0; q < 16; ++q) s += buf[base + (size_t)q * gridDim.x * blockDim.x];
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) atomicAdd(y + (blockIdx.x & 63) * 8 + (threadIdx.x >> 5), s);
}

__device__ __forceinline__ unsigned ld_acq(const unsigned* p)
{
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

__global__ void k_persist(double* buf, double* y, unsigned* ctr, int iters, int work)
{
    unsigned target = 0;
    for (int it = 0; it < iters; ++it) {
        if (work) {
            const size_t base = ((size_t)blockIdx.x * blockDim.x + threadIdx.x);
            double s = 0.0;
#pragma unroll
            for (int q = 0; q < 16; ++q) s += buf[base + (size_t)q * gridDim.x * blockDim.x];
            for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
            if ((threadIdx.x & 31) == 0) atomicAdd(y + (blockIdx.x & 63) * 8 + (threadIdx.x >> 5), s);
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            target += gridDim.x;
            __threadfence();
            atomicAdd(ctr, 1u);
            while (ld_acq(ctr) < target) { }
        }
        __syncthreads();
    }
}

int main()
{
    const int N = 2000;
    double *buf, *y;
    unsigned* ctr;
    cudaMalloc(&buf, sizeof(double) * 16 * 1024 * 256 * 4);
    cudaMemset(buf, 0, sizeof(double) * 16 * 1024 * 256 * 4);
    cudaMalloc(&y, sizeof(double) * 1024);
    cudaMemset(y, 0, sizeof(double) * 1024);
    cudaMalloc(&ctr, 4);
    cudaStream_t s;
    cudaStreamCreate(&s);
    cudaEvent_t a, b;
    cudaEventCreate(&a);
    cudaEventCreate(&b);
    float ms;
    for (int grid : {20, 120, 720}) {
        for (int pdl = 0; pdl < 2; ++pdl) {
            for (int work = 0; work < 2; ++work) {
                cudaLaunchAttribute at[1];
                at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
                at[0].val.programmaticStreamSerializationAllowed = 1;
                cudaLaunchConfig_t cfg = {};
                cfg.stream = s; cfg.attrs = at; cfg.numAttrs = pdl ? 1 : 0;
                cfg.gridDim = dim3(grid); cfg.blockDim = dim3(256);
                for (int rep = 0; rep < 2; ++rep) {
                    cudaEventRecord(a, s);
                    for (int i = 0; i < N; ++i) {
                        if (work) cudaLaunchKernelEx(&cfg, k_work, buf, y, pdl);
                        else cudaLaunchKernelEx(&cfg, k_empty, buf, pdl);
                    }
                    cudaEventRecord(b, s);
                    cudaEventSynchronize(b);
                    cudaEventElapsedTime(&ms, a, b);
                }
                printf("chain grid=%d pdl=%d work=%d : %.3f us/launch  (%s)\n", grid, pdl, work, ms * 1e3 / N,
                       cudaGetErrorString(cudaGetLastError()));
            }
        }
    }
    for (int grid : {20, 148, 444}) {
        for (int work = 0; work < 2; ++work) {
            int iters = N;
            void* args[] = {&buf, &y, &ctr, &iters, &work};
            for (int rep = 0; rep < 2; ++rep) {
                cudaMemsetAsync(ctr, 0, 4, s);
                cudaEventRecord(a, s);
                cudaLaunchCooperativeKernel((void*)k_persist, dim3(grid), dim3(256), args, 0, s);
                cudaEventRecord(b, s);
                cudaEventSynchronize(b);
                cudaEventElapsedTime(&ms, a, b);
            }
            printf("persist grid=%d work=%d : %.3f us/iter  (%s)\n", grid, work, ms * 1e3 / N,
                   cudaGetErrorString(cudaGetLastError()));
        }
    }
    return 0;
}
