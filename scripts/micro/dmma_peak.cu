// DMMA (mma.sync.m8n8k4.f64) issue-rate microbenchmark: register operands only.
// build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o dmma_peak dmma_peak.cu ; run: ./dmma_peak
#include <cstdio>
#include <cuda_runtime.h>

template <int NACC>
__global__ void dmma_loop(double* out, int iters, double a0, double b0)
{
    double acc[NACC][2];
#pragma unroll
    for (int i = 0; i < NACC; ++i) { acc[i][0] = threadIdx.x; acc[i][1] = i; }
    double a = a0 + threadIdx.x * 1e-9, b = b0 - threadIdx.x * 1e-9;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < NACC; ++i)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                         : "+d"(acc[i][0]), "+d"(acc[i][1]) : "d"(a), "d"(b));
    }
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < NACC; ++i) s += acc[i][0] + acc[i][1];
    if (s == 1.2345e300) out[0] = s;
}

template <int NACC>
void run(int warps, int ctas_per_sm, int sms, double* out)
{
    const int iters = 4096;
    dim3 grid(sms * ctas_per_sm), block(32 * warps);
    dmma_loop<NACC><<<grid, block>>>(out, 64, 1.0, 1.0);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    dmma_loop<NACC><<<grid, block>>>(out, iters, 1.0, 1.0);
    cudaEventRecord(e1);
    cudaDeviceSynchronize();
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double flops = 512.0 * NACC * iters * warps * (double)grid.x;
    printf("{\"nacc\": %d, \"warps_per_cta\": %d, \"ctas_per_sm\": %d, \"ms\": %.4f, \"tflops\": %.2f}\n", NACC, warps,
           ctas_per_sm, ms, flops / ms / 1e9);
}

int main()
{
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    double* out; cudaMalloc(&out, 8);
    const int ws[] = {4, 8, 16, 32};
    for (int w : ws) { run<1>(w, 1, sms, out); run<2>(w, 1, sms, out); run<4>(w, 1, sms, out); run<8>(w, 1, sms, out); run<16>(w, 1, sms, out); }
    run<8>(8, 2, sms, out); run<8>(8, 3, sms, out); run<8>(8, 4, sms, out);
    return 0;
}
