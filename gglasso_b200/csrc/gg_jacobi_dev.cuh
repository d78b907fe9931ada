// gg_jacobi_dev.cuh -- shared-memory one-sided Jacobi building block (used by the small-matrix path,
// the block-Jacobi pair kernel and the divide & conquer leaves).
#pragma once
#include "gg_common.cuh"

// ---- round-robin tournament: n even, round r in [0,n-1), slot s in [0,n/2) -------------------
__device__ __forceinline__ void rr_pair(int n, int r, int s, int& a, int& b)
{
    const int m = n - 1;
    if (s == 0) { a = m; b = r; }
    else { a = (r + s) % m; b = (r - s + m) % m; }
}

// ---- shared-memory one-sided Jacobi on the rows of G (n x n, row stride ld) --------------------
// LP lanes cooperate on one row pair; every lane keeps its NE = ceil(n/LP) elements of both rows in registers
// between the dot-product and the rotation phase (one shared-memory read and one write per element and
// rotation).  LP = 16 makes every half-warp access one contiguous 128-byte row segment (conflict free).
// The three dot products run on two accumulators each to halve the dependent FP64 chains, which -- not
// throughput -- bound this latency-limited kernel (ncu: 41 % short-scoreboard, 23 % barrier stalls).
// Returns the number of sweeps executed.
template <int LP, int NE>
__device__ int jacobi_rows_smem(double* G, int n, int ld, double tol, int max_sweeps)
{
    const int nn = n + (n & 1);
    const int half = nn >> 1;
    const int ngroups = blockDim.x / LP;
    const int gid = threadIdx.x / LP, gl = threadIdx.x % LP;
    const double tol2 = tol * tol;
    int sweep = 0;
    for (; sweep < max_sweeps; ++sweep) {
        int rot = 0;
        for (int r = 0; r < nn - 1; ++r) {
            for (int s0 = 0; s0 < half; s0 += ngroups) {
                const int s = s0 + gid;
                bool act = s < half;
                int i = 0, j = 0;
                if (act) {
                    rr_pair(nn, r, s, i, j);
                    if (i > j) { const int t = i; i = j; j = t; }
                    act = j < n;
                }
                double x[NE], y[NE];
                double a0 = 0.0, a1 = 0.0, b0 = 0.0, b1 = 0.0, g0 = 0.0, g1 = 0.0;
                double* gi = G + (size_t)i * ld;
                double* gj = G + (size_t)j * ld;
#pragma unroll
                for (int q = 0; q < NE; ++q) {
                    const int e = gl + q * LP;
                    const bool ok = act && e < n;
                    x[q] = ok ? gi[e] : 0.0;
                    y[q] = ok ? gj[e] : 0.0;
                }
#pragma unroll
                for (int q = 0; q < NE; q += 2) {
                    a0 = fma(x[q], x[q], a0); b0 = fma(y[q], y[q], b0); g0 = fma(x[q], y[q], g0);
                    if (q + 1 < NE) {
                        a1 = fma(x[q + 1], x[q + 1], a1); b1 = fma(y[q + 1], y[q + 1], b1);
                        g1 = fma(x[q + 1], y[q + 1], g1);
                    }
                }
                double a = a0 + a1, b = b0 + b1, g = g0 + g1;
#pragma unroll
                for (int o = LP >> 1; o > 0; o >>= 1) {
                    a += __shfl_xor_sync(0xffffffffu, a, o);
                    b += __shfl_xor_sync(0xffffffffu, b, o);
                    g += __shfl_xor_sync(0xffffffffu, g, o);
                }
                if (act && g * g > tol2 * (a * b)) {
                    // tan(theta) = 2g sign(d) / (|d| + sqrt(d^2 + 4g^2)), d = b - a   (smaller root)
                    const double dd = b - a;
                    const double hh = sqrt(fma(dd, dd, 4.0 * g * g));
                    const double t = copysign(2.0 * g, dd * g) / (fabs(dd) + hh);
                    const double c = rsqrt(fma(t, t, 1.0));
                    const double sn = c * t;
#pragma unroll
                    for (int q = 0; q < NE; ++q) {
                        const int e = gl + q * LP;
                        if (e < n) {
                            gi[e] = c * x[q] - sn * y[q];
                            gj[e] = sn * x[q] + c * y[q];
                        }
                    }
                    rot = 1;
                }
            }
            __syncthreads();
        }
        if (__syncthreads_count(rot) == 0) { ++sweep; break; }
    }
    return sweep;
}
