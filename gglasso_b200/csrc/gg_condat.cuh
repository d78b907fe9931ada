// gg_condat.cuh -- in-place direct 1-D total-variation prox on a strided K-vector.
//
// Replaces the per-entry call chain prox_phi_fgl -> prox_tv -> condat_method of the reference
// (src/gglasso/solver/ggl_helper.py:126-134, src/gglasso/solver/fgl_helper.py:11-68).
// Same taut-string state machine (k, k0, k-, k+, vmin, vmax, umin, umax) and the same
// floating point expressions, so the result is bit-identical to the reference for the same
// input vector.  In place: a segment [k0, k+-] is only written after the scan has moved past
// it, and a restart resumes at k+- + 1, so no value is read after it was overwritten.
//
// __host__ __device__ so the exact device logic is unit-tested on the CPU (tests/test_host_logic.py
// via gg_host_tv1d in gg_capi.cu).
#pragma once

#ifdef __CUDACC__
#define GG_HD __host__ __device__ __forceinline__
#else
#define GG_HD inline
#endif

// v[i*stride], i = 0..n-1
GG_HD void gg_tv1d_inplace(double* v, int n, int stride, double lam)
{
#define Y(i) v[(i) * stride]
    int k = 0, k0 = 0, kp = 0, km = 0;
    double vmin = Y(0) - lam, vmax = Y(0) + lam;
    double umin = lam, umax = -lam;
    for (;;) {
        if (k == n - 1) {
            if (umin < 0.0) {
                for (int i = k0; i <= km; ++i) Y(i) = vmin;
                km += 1; k = k0 = km;
                const double yk = Y(k);
                umin = lam; vmin = yk; umax = yk + lam - vmax;
            } else if (umax > 0.0) {
                for (int i = k0; i <= kp; ++i) Y(i) = vmax;
                kp += 1; k = k0 = kp;
                const double yk = Y(k);
                umax = -lam; vmax = yk; umin = yk - lam - vmin;
            } else {
                const double val = vmin + umin / (double)(k - k0 + 1);
                for (int i = k0; i < n; ++i) Y(i) = val;
                return;
            }
            if (k == n - 1) { Y(k) = vmin + umin; return; }
            continue;
        }
        const double yn = Y(k + 1);
        if (yn + umin - vmin < -lam) {
            for (int i = k0; i <= km; ++i) Y(i) = vmin;
            km += 1; k = kp = k0 = km;
            const double yk = Y(k);
            vmin = yk; vmax = yk + 2.0 * lam;
            umin = lam; umax = -lam;
        } else if (yn + umax - vmax > lam) {
            for (int i = k0; i <= kp; ++i) Y(i) = vmax;
            kp += 1; k = km = k0 = kp;
            const double yk = Y(k);
            vmin = yk - 2.0 * lam; vmax = yk;
            umin = lam; umax = -lam;
        } else {
            k += 1;
            umin = umin + yn - vmin;
            umax = umax + yn - vmax;
            if (umin >= lam)  { vmin += (umin - lam) / (double)(k - k0 + 1); umin = lam;  km = k; }
            if (umax <= -lam) { vmax += (umax + lam) / (double)(k - k0 + 1); umax = -lam; kp = k; }
        }
    }
#undef Y
}

GG_HD double gg_soft(double v, double l)
{
    // sign(v) * max(|v| - l, 0)   (reference: ggl_helper.py:12-14)
    double a = fabs(v) - l;
    a = a > 0.0 ? a : 0.0;
    return v > 0.0 ? a : (v < 0.0 ? -a : 0.0);
}
