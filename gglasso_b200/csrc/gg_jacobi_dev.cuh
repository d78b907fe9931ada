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
// LP lanes cooperate on one row pair.  Returns the number of sweeps executed.
template <int LP>
__device__ int jacobi_rows_smem(double* G, int n, int ld, double tol, int max_sweeps)
{
    const int nn = n + (n & 1);
    const int half = nn >> 1;
    const int ngroups = blockDim.x / LP;
    const int gid = threadIdx.x / LP, gl = threadIdx.x % LP;
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
                double a = 0.0, b = 0.0, g = 0.0;
                if (act) {
                    const double* gi = G + (size_t)i * ld;
                    const double* gj = G + (size_t)j * ld;
                    for (int e = gl; e < n; e += LP) {
                        const double x = gi[e], y = gj[e];
                        a = fma(x, x, a); b = fma(y, y, b); g = fma(x, y, g);
                    }
                }
#pragma unroll
                for (int o = LP >> 1; o > 0; o >>= 1) {
                    a += __shfl_xor_sync(0xffffffffu, a, o);
                    b += __shfl_xor_sync(0xffffffffu, b, o);
                    g += __shfl_xor_sync(0xffffffffu, g, o);
                }
                if (act && g * g > (tol * tol) * (a * b)) {
                    // tan(theta) = 2g sign(d) / (|d| + sqrt(d^2 + 4g^2)), d = b - a   (smaller root; one sqrt,
                    // one division, one rsqrt -- these FP64 chains dominate the latency of a round)
                    const double dd = b - a;
                    const double hh = sqrt(fma(dd, dd, 4.0 * g * g));
                    const double t = copysign(2.0 * g, dd * g) / (fabs(dd) + hh);
                    const double c = rsqrt(fma(t, t, 1.0));
                    const double sn = c * t;
                    double* gi = G + (size_t)i * ld;
                    double* gj = G + (size_t)j * ld;
                    for (int e = gl; e < n; e += LP) {
                        const double x = gi[e], y = gj[e];
                        gi[e] = c * x - sn * y;
                        gj[e] = sn * x + c * y;
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

