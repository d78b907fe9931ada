// gg_common.cuh -- shared device helpers for the gglasso_b200 kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define GG_WARP 32

// ---- control block (device resident, one per ADMM problem) -----------------------------
// All per-iteration scalars live on the device so the ADMM loop needs no host round trip:
// kernels read rho / the pending dual rescale / the done flag from here.
#define GG_CTRL_STRIDE 16
#define GG_C_RHO      0   // current penalty parameter rho
#define GG_C_XSCALE   1   // pending factor for the scaled dual X (rho_old/rho_new), applied lazily
#define GG_C_DONE     2   // 1.0 once the stopping test has passed (kernels become no-ops)
#define GG_C_ITER     3   // number of completed iterations
#define GG_C_R        4
#define GG_C_S        5
#define GG_C_EPRI     6
#define GG_C_EDUAL    7
#define GG_C_STATUS   8   // 1 = optimal (set together with DONE)
#define GG_C_LAM1     9   // > 0: lambda1 of the MGL prox is taken from here instead of the kernel argument, so that a
#define GG_C_LAM2     10  //      captured CUDA graph of the iteration can be replayed for another grid point

#define GG_HIST_STRIDE 5  // r, s, e_pri, e_dual, rho (value used during the iteration)
#define GG_NPART 5        // |Omega|^2, |Theta-L|^2, |X|^2, |Omega-Theta+L|^2, |Omega-Omega_prev|^2

// kernel-launch counter of the library (gg_launch_count() in the C ABI): every launch site calls it once per launch
void gg_count_launch(int n);

#define GG_CHECK_LAUNCH()                                   \
    do {                                                    \
        cudaError_t e__ = cudaGetLastError();               \
        if (e__ != cudaSuccess) return (int)e__;            \
    } while (0)

__device__ __forceinline__ double gg_warp_sum(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

__device__ __forceinline__ double gg_warp_max(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// Block-wide sum of N values per thread; result valid in thread 0. `scratch` >= N*32 doubles.
template <int N>
__device__ __forceinline__ void gg_block_sum(double (&v)[N], double* scratch)
{
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int nw = (blockDim.x + 31) >> 5;
#pragma unroll
    for (int i = 0; i < N; ++i) v[i] = gg_warp_sum(v[i]);
    __syncthreads();
    if (lane == 0) {
#pragma unroll
        for (int i = 0; i < N; ++i) scratch[i * 32 + wid] = v[i];
    }
    __syncthreads();
    if (wid == 0) {
#pragma unroll
        for (int i = 0; i < N; ++i) {
            double t = (lane < nw) ? scratch[i * 32 + lane] : 0.0;
            v[i] = gg_warp_sum(t);
        }
    }
}

__device__ __forceinline__ double gg_block_max(double v, double* scratch)
{
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int nw = (blockDim.x + 31) >> 5;
    v = gg_warp_max(v);
    __syncthreads();
    if (lane == 0) scratch[wid] = v;
    __syncthreads();
    double t = (lane < nw) ? scratch[lane] : -1.0e300;
    t = gg_warp_max(t);
    return t;  // valid in warp 0 (all lanes)
}

// ---- cp.async (LDGSTS) -------------------------------------------------------------------
__device__ __forceinline__ void gg_cp_async8(void* smem, const void* gmem)
{
    uint32_t s = (uint32_t)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void gg_cp_async16(void* smem, const void* gmem)
{
    uint32_t s = (uint32_t)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void gg_cp_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void gg_cp_wait()
{
    asm volatile("cp.async.wait_group %0;\n" ::"n"(N));
}

// ---- FP64 tensor core MMA (DMMA m8n8k4) ---------------------------------------------------
// A fragment: a = A[lane/4][lane%4]; B fragment: b = B[lane%4][lane/4];
// C fragment: c0,c1 = C[lane/4][2*(lane%4) + {0,1}].
__device__ __forceinline__ void gg_dmma(double& c0, double& c1, double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}
