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

int gg_launch_recon(const double* Vt, const double* D, const double* bnum, const double* ctrl, int mpp, int mode,
                    int M, int p, double* Out, cudaStream_t st)
{
    const int nt = (p + RT - 1) / RT;
    dim3 grid(nt, nt, M);
    gg_count_launch(1);
    if ((p & 1) == 0) recon_kernel<true><<<grid, 256, 0, st>>>(Vt, D, bnum, ctrl, mpp, mode, p, Out);
    else recon_kernel<false><<<grid, 256, 0, st>>>(Vt, D, bnum, ctrl, mpp, mode, p, Out);
    GG_CHECK_LAUNCH();
    return 0;
}
