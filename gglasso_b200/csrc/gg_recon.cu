// gg_recon.cu -- spectral reconstruction  Out = V diag(f(D)) V^T  on FP64 tensor cores (sm_100a).
//
// Replaces phiplus / prox_rank_norm of the reference:
//   mode 0  f(d) = (sqrt(d^2 + 4 beta) + d)/2, beta = bnum[m]/rho   ggl_helper.py:272-303, admm_solver.py:183-187
//   mode 1  f(d) = max(d - beta, 0),          beta = bnum[m]/rho   ggl_helper.py:29-36,   admm_solver.py:197-205
//   mode 2  f(d) = d                                               (plain reassembly, used by tests)
// Input is Vt (rows = eigenvectors) as produced by gg_eigh, so Out[i][j] = sum_c f_c Vt[c][i] Vt[c][j]:
// a "TN" rank-p update.  Only tile pairs I <= J are computed (SYRK: p^3 flops instead of 2p^3) and the
// mirror tile is written through shared memory, so the result is exactly symmetric.
#include "gg_common.cuh"
#include <stdlib.h>

#define RT 64          // output tile
#define RKC 16         // eigenvector chunk
#define RLD (RT + 4)   // (4*c + i) mod 16 distinct -> conflict-free DMMA fragment loads

template <bool VEC>
__global__ void __launch_bounds__(256)
recon_kernel(const double* __restrict__ Vt, const double* __restrict__ D, const double* __restrict__ bnum,
             const double* __restrict__ ctrl, int mpp, int mode, int p, double* __restrict__ Out)
{
    __shared__ __align__(16) double sm[2 * 2 * RKC * RLD];   // As[2][RKC][RLD], Bs[2][RKC][RLD]; reused as Cs[64][65]
    __shared__ double fs[2][RKC];
    const int I = blockIdx.y, J = blockIdx.x, m = blockIdx.z;
    if (I > J) return;
    double rho = 1.0;
    if (ctrl) {
        const double* c = ctrl + (size_t)(m / mpp) * GG_CTRL_STRIDE;
        if (c[GG_C_DONE] != 0.0) return;
        rho = c[GG_C_RHO];
    }
    const double beta = (bnum ? bnum[m] : 1.0) / rho;
    const double* V = Vt + (size_t)m * p * p;
    const double* Dm = D + (size_t)m * p;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int fr = lane >> 2, fc = lane & 3;
    const int wr = wid >> 1, wc = wid & 1;          // 4 x 2 warps -> 16 x 32 outputs per warp
    const int i0 = I * RT, j0 = J * RT;
    const int nchunks = (p + RKC - 1) / RKC;

    auto load_chunk = [&](int c, int buf) {
        double* As = sm + (size_t)buf * RKC * RLD;
        double* Bs = sm + (size_t)(2 + buf) * RKC * RLD;
        const int c0 = c * RKC;
        if (VEC) {
            for (int idx = tid; idx < 2 * RKC * (RT / 2); idx += 256) {
                const int which = idx / (RKC * (RT / 2));
                const int rem = idx % (RKC * (RT / 2));
                const int kk = rem / (RT / 2), e = (rem % (RT / 2)) * 2;
                const int gc = (which ? j0 : i0) + e, gr = c0 + kk;
                double* dst = (which ? Bs : As) + kk * RLD + e;
                if (gr < p && gc < p) gg_cp_async16(dst, V + (size_t)gr * p + gc);
                else { dst[0] = 0.0; dst[1] = 0.0; }
            }
        } else {
            for (int idx = tid; idx < 2 * RKC * RT; idx += 256) {
                const int which = idx / (RKC * RT);
                const int rem = idx % (RKC * RT);
                const int kk = rem / RT, e = rem % RT;
                const int gc = (which ? j0 : i0) + e, gr = c0 + kk;
                double* dst = (which ? Bs : As) + kk * RLD + e;
                if (gr < p && gc < p) gg_cp_async8(dst, V + (size_t)gr * p + gc);
                else dst[0] = 0.0;
            }
        }
        if (tid < RKC) {
            const int cc = c0 + tid;
            double f = 0.0;
            if (cc < p) {
                const double d = Dm[cc];
                if (mode == 0) f = 0.5 * (sqrt(d * d + 4.0 * beta) + d);
                else if (mode == 1) f = fmax(d - beta, 0.0);
                else f = d;
            }
            fs[buf][tid] = f;
        }
        gg_cp_commit();
    };

    double acc[2][4][2];
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) { acc[a][b][0] = 0.0; acc[a][b][1] = 0.0; }

    load_chunk(0, 0);
    for (int c = 0; c < nchunks; ++c) {
        if (c + 1 < nchunks) { load_chunk(c + 1, (c + 1) & 1); gg_cp_wait<1>(); }
        else gg_cp_wait<0>();
        __syncthreads();
        const int buf = c & 1;
        const double* As = sm + (size_t)buf * RKC * RLD;
        const double* Bs = sm + (size_t)(2 + buf) * RKC * RLD;
#pragma unroll
        for (int k0 = 0; k0 < RKC; k0 += 4) {
            const double f = fs[buf][k0 + fc];
            double fa[2], fb[4];
#pragma unroll
            for (int a = 0; a < 2; ++a) fa[a] = As[(k0 + fc) * RLD + (wr * 2 + a) * 8 + fr] * f;
#pragma unroll
            for (int b = 0; b < 4; ++b) fb[b] = Bs[(k0 + fc) * RLD + (wc * 4 + b) * 8 + fr];
#pragma unroll
            for (int a = 0; a < 2; ++a)
#pragma unroll
                for (int b = 0; b < 4; ++b) gg_dmma(acc[a][b][0], acc[a][b][1], fa[a], fb[b]);
        }
        __syncthreads();
    }

    // ---- epilogue through shared memory: tile (I,J) and its mirror (J,I) --------------------
    double* Cs = sm;                         // 64 x 65
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            const int r = (wr * 2 + a) * 8 + fr, cc = (wc * 4 + b) * 8 + 2 * fc;
            Cs[r * 65 + cc] = acc[a][b][0];
            Cs[r * 65 + cc + 1] = acc[a][b][1];
        }
    __syncthreads();
    double* O = Out + (size_t)m * p * p;
    for (int idx = tid; idx < RT * RT; idx += 256) {
        const int r = idx / RT, cc = idx % RT;
        const int gi = i0 + r, gj = j0 + cc;
        if (gi < p && gj < p) {
            const double v = (I == J && cc < r) ? Cs[cc * 65 + r] : Cs[r * 65 + cc];
            O[(size_t)gi * p + gj] = v;
        }
    }
    if (I < J) {
        for (int idx = tid; idx < RT * RT; idx += 256) {
            const int r = idx / RT, cc = idx % RT;        // element (j0 + r, i0 + cc) of the mirror tile
            const int gi = j0 + r, gj = i0 + cc;
            if (gi < p && gj < p) O[(size_t)gi * p + gj] = Cs[cc * 65 + r];
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// Bulk-copy (TMA engine) variant for even p: the operand rows -- 64 consecutive doubles of an eigenvector row, 512
// bytes, 16-byte aligned when p is even -- are fetched with cp.async.bulk (SASS UBLKCP) by ONE warp and land in a
// three-stage ring guarded by mbarriers (expect_tx / complete_tx); the other warps never touch the load path, they
// wait on the stage's barrier and issue DMMAs.  2-D tensor maps are not used: they write dense boxes, and the m8n8k4
// fragment loads need the 68-double row pitch to stay bank-conflict free (a 128-byte swizzle leaves them 2-way
// conflicted), so each row is its own 1-D bulk copy into the padded layout.
#define RS 3           // ring stages

__device__ __forceinline__ uint32_t rc_smem(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void rc_mbar_init(uint64_t* bar, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(rc_smem(bar)), "r"(count));
}
__device__ __forceinline__ void rc_mbar_expect(uint64_t* bar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(rc_smem(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void rc_mbar_wait(uint64_t* bar, unsigned parity)
{
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "RC_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra RC_DONE;\n"
        "bra RC_WAIT;\n"
        "RC_DONE:\n"
        "}\n" ::"r"(rc_smem(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void rc_bulk_g2s(void* dst, const void* src, unsigned bytes, uint64_t* bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(rc_smem(dst)), "l"(src), "r"(bytes), "r"(rc_smem(bar)) : "memory");
}

__global__ void __launch_bounds__(256)
recon_bulk_kernel(const double* __restrict__ Vt, const double* __restrict__ D, const double* __restrict__ bnum,
                  const double* __restrict__ ctrl, int mpp, int mode, int p, double* __restrict__ Out)
{
    extern __shared__ __align__(16) double rsm[];     // As[RS][RKC][RLD], Bs[RS][RKC][RLD], fs[RS][RKC], full[RS]
    double* As = rsm;
    double* Bs = rsm + RS * RKC * RLD;
    double* fs = rsm + 2 * RS * RKC * RLD;
    uint64_t* full = reinterpret_cast<uint64_t*>(fs + RS * RKC);
    const int I = blockIdx.y, J = blockIdx.x, m = blockIdx.z;
    if (I > J) return;
    double rho = 1.0;
    if (ctrl) {
        const double* c = ctrl + (size_t)(m / mpp) * GG_CTRL_STRIDE;
        if (c[GG_C_DONE] != 0.0) return;
        rho = c[GG_C_RHO];
    }
    const double beta = (bnum ? bnum[m] : 1.0) / rho;
    const double* V = Vt + (size_t)m * p * p;
    const double* Dm = D + (size_t)m * p;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int fr = lane >> 2, fc = lane & 3;
    const int wr = wid >> 1, wc = wid & 1;
    const int i0 = I * RT, j0 = J * RT;
    const int nchunks = (p + RKC - 1) / RKC;
    const int wA = min(RT, p - i0), wB = min(RT, p - j0);          // even: p, i0, j0 are even
    // columns beyond p (edge tiles) and rows beyond p (last chunk) must read as zero: clear the ring once; every chunk
    // rewrites the same column range of a stage, the producer clears rows beyond p explicitly
    for (int e = tid; e < 2 * RS * RKC * RLD; e += 256) rsm[e] = 0.0;
    if (tid == 0) {
        for (int s = 0; s < RS; ++s) rc_mbar_init(full + s, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");    // the zero fill precedes the async-proxy writes

    auto issue = [&](int c) {                     // warp 0: rows of chunk c into stage c % RS
        const int stg = c % RS, c0 = c * RKC;
        const int which = lane >> 4, kk = lane & 15, gr = c0 + kk;
        const int rows = min(RKC, p - c0);
        if (lane == 0) rc_mbar_expect(full + stg, (unsigned)(rows * (wA + wB) * sizeof(double)));
        __syncwarp();
        double* dst = (which ? Bs : As) + ((size_t)stg * RKC + kk) * RLD;
        if (gr < p) {
            rc_bulk_g2s(dst, V + (size_t)gr * p + (which ? j0 : i0), (unsigned)((which ? wB : wA) * sizeof(double)), full + stg);
        } else {
            for (int e = 0; e < RT; ++e) dst[e] = 0.0;
        }
        if (lane < RKC) {
            const int cc = c0 + lane;
            double f = 0.0;
            if (cc < p) {
                const double d = Dm[cc];
                if (mode == 0) f = 0.5 * (sqrt(d * d + 4.0 * beta) + d);
                else if (mode == 1) f = fmax(d - beta, 0.0);
                else f = d;
            }
            fs[stg * RKC + lane] = f;
        }
    };

    if (wid == 0)
        for (int c = 0; c < RS && c < nchunks; ++c) issue(c);
    __syncthreads();

    double acc[2][4][2];
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) { acc[a][b][0] = 0.0; acc[a][b][1] = 0.0; }

    for (int c = 0; c < nchunks; ++c) {
        const int stg = c % RS;
        rc_mbar_wait(full + stg, (unsigned)((c / RS) & 1));
        const double* Ac = As + (size_t)stg * RKC * RLD;
        const double* Bc = Bs + (size_t)stg * RKC * RLD;
#pragma unroll
        for (int k0 = 0; k0 < RKC; k0 += 4) {
            const double f = fs[stg * RKC + k0 + fc];
            double fa[2], fb[4];
#pragma unroll
            for (int a = 0; a < 2; ++a) fa[a] = Ac[(k0 + fc) * RLD + (wr * 2 + a) * 8 + fr] * f;
#pragma unroll
            for (int b = 0; b < 4; ++b) fb[b] = Bc[(k0 + fc) * RLD + (wc * 4 + b) * 8 + fr];
#pragma unroll
            for (int a = 0; a < 2; ++a)
#pragma unroll
                for (int b = 0; b < 4; ++b) gg_dmma(acc[a][b][0], acc[a][b][1], fa[a], fb[b]);
        }
        __syncthreads();                          // every warp is done with this stage
        if (wid == 0 && c + RS < nchunks) {
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            issue(c + RS);
        }
    }

    // ---- epilogue through shared memory: tile (I,J) and its mirror (J,I) --------------------
    double* Cs = rsm;                        // 64 x 65
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            const int r = (wr * 2 + a) * 8 + fr, cc = (wc * 4 + b) * 8 + 2 * fc;
            Cs[r * 65 + cc] = acc[a][b][0];
            Cs[r * 65 + cc + 1] = acc[a][b][1];
        }
    __syncthreads();
    double* O = Out + (size_t)m * p * p;
    for (int idx = tid; idx < RT * RT; idx += 256) {
        const int r = idx / RT, cc = idx % RT;
        const int gi = i0 + r, gj = j0 + cc;
        if (gi < p && gj < p) {
            const double v = (I == J && cc < r) ? Cs[cc * 65 + r] : Cs[r * 65 + cc];
            O[(size_t)gi * p + gj] = v;
        }
    }
    if (I < J) {
        for (int idx = tid; idx < RT * RT; idx += 256) {
            const int r = idx / RT, cc = idx % RT;        // element (j0 + r, i0 + cc) of the mirror tile
            const int gi = j0 + r, gj = i0 + cc;
            if (gi < p && gj < p) O[(size_t)gi * p + gj] = Cs[cc * 65 + r];
        }
    }
}

int gg_launch_recon(const double* Vt, const double* D, const double* bnum, const double* ctrl, int mpp, int mode,
                    int M, int p, double* Out, cudaStream_t st)
{
    const int nt = (p + RT - 1) / RT;
    dim3 grid(nt, nt, M);
    gg_count_launch(1);
    static int recon_old = -1;
    if (recon_old < 0) { const char* ev = getenv("GG_RECON_OLD"); recon_old = ev ? atoi(ev) : 0; }
    if ((p & 1) == 0 && !recon_old) {
        const size_t smem = sizeof(double) * (2 * RS * RKC * RLD + RS * RKC) + sizeof(uint64_t) * RS;
        cudaError_t e = cudaFuncSetAttribute(recon_bulk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
        recon_bulk_kernel<<<grid, 256, smem, st>>>(Vt, D, bnum, ctrl, mpp, mode, p, Out);
    } else if ((p & 1) == 0) recon_kernel<true><<<grid, 256, 0, st>>>(Vt, D, bnum, ctrl, mpp, mode, p, Out);
    else recon_kernel<false><<<grid, 256, 0, st>>>(Vt, D, bnum, ctrl, mpp, mode, p, Out);
    GG_CHECK_LAUNCH();
    return 0;
}
