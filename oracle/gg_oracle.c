/*
 * oracle/gg_oracle.c -- TEST INFRASTRUCTURE ONLY (parity checker + CPU baseline leg).
 *
 * Plain-C restatement of the per-entry proximal operators on the GGLasso ADMM hot path.
 * Nothing in the product package (gglasso_b200/) links or loads this file; only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may use it.
 *
 * Reference behaviour restated (file:line relative to /root/reference):
 *   gg_tv1d        : src/gglasso/solver/fgl_helper.py:11-68   (condat_method, Condat 2013 taut string)
 *   gg_prox_p_fgl  : src/gglasso/solver/ggl_helper.py:131-134 (prox_phi_fgl = soft(TV(v,l2), l1))
 *                    src/gglasso/solver/ggl_helper.py:190-207 (prox_p: upper triangle + mirror, diag kept)
 *   gg_prox_p_ggl  : src/gglasso/solver/ggl_helper.py:68-71   (prox_phi_ggl = group-shrink(soft(v,l1), l2))
 *                    src/gglasso/solver/ggl_helper.py:38-43   (prox_2norm)
 *   gg_pval        : src/gglasso/solver/ggl_helper.py:162-176 (P_val)
 *
 * Pinned against the real reference through tests/golden/*.npz (tests/test_oracle_golden.py).
 */
#include <math.h>
#include <stddef.h>
#include <stdlib.h>

/* 1-D total-variation prox  argmin_x 0.5||x-y||^2 + lam*sum|x[i+1]-x[i]|  (direct algorithm).
 * x may alias y?  No: kept separate here for clarity. */
void gg_tv1d(const double *y, int n, double lam, double *x)
{
    if (n <= 0) return;
    int k = 0, k0 = 0, kp = 0, km = 0;
    double vmin = y[0] - lam, vmax = y[0] + lam;
    double umin = lam, umax = -lam;
    for (;;) {
        if (k == n - 1) {
            /* end of signal reached with the current segment still open */
            if (umin < 0.0) {
                for (int i = k0; i <= km; ++i) x[i] = vmin;
                km += 1; k = k0 = km;
                umin = lam; vmin = y[k]; umax = y[k] + lam - vmax;
            } else if (umax > 0.0) {
                for (int i = k0; i <= kp; ++i) x[i] = vmax;
                kp += 1; k = k0 = kp;
                umax = -lam; vmax = y[k]; umin = y[k] - lam - vmin;
            } else {
                double v = vmin + umin / (double)(k - k0 + 1);
                for (int i = k0; i < n; ++i) x[i] = v;
                return;
            }
            if (k == n - 1) { x[k] = vmin + umin; return; }
            continue;
        }
        if (y[k + 1] + umin - vmin < -lam) {            /* negative jump */
            for (int i = k0; i <= km; ++i) x[i] = vmin;
            km += 1; k = kp = k0 = km;
            vmin = y[k]; vmax = y[k] + 2.0 * lam;
            umin = lam; umax = -lam;
        } else if (y[k + 1] + umax - vmax > lam) {      /* positive jump */
            for (int i = k0; i <= kp; ++i) x[i] = vmax;
            kp += 1; k = km = k0 = kp;
            vmin = y[k] - 2.0 * lam; vmax = y[k];
            umin = lam; umax = -lam;
        } else {                                        /* no jump: extend segment */
            k += 1;
            umin = umin + y[k] - vmin;
            umax = umax + y[k] - vmax;
            if (umin >= lam)  { vmin += (umin - lam) / (double)(k - k0 + 1); umin = lam;  km = k; }
            if (umax <= -lam) { vmax += (umax + lam) / (double)(k - k0 + 1); umax = -lam; kp = k; }
        }
    }
}

static inline double soft(double v, double l)
{
    double a = fabs(v) - l;
    if (a < 0.0) a = 0.0;
    return (v > 0.0) ? a : ((v < 0.0) ? -a : 0.0 * a);
}

/* X, M: (K,p,p) C-contiguous. Only the upper triangle of X is read (as the reference does). */
void gg_prox_p_fgl(const double *X, int K, int p, double l1, double l2, double *M)
{
    double *v = (double *)malloc(sizeof(double) * (size_t)K * 2);
    double *t = v + K;
    size_t pp = (size_t)p * p;
    for (int i = 0; i < p; ++i) {
        for (int j = i; j < p; ++j) {
            if (i == j) {
                for (int k = 0; k < K; ++k) M[k * pp + (size_t)i * p + i] = X[k * pp + (size_t)i * p + i];
                continue;
            }
            for (int k = 0; k < K; ++k) v[k] = X[k * pp + (size_t)i * p + j];
            gg_tv1d(v, K, l2, t);
            for (int k = 0; k < K; ++k) {
                double r = soft(t[k], l1);
                M[k * pp + (size_t)i * p + j] = r;
                M[k * pp + (size_t)j * p + i] = r;
            }
        }
    }
    free(v);
}

void gg_prox_p_ggl(const double *X, int K, int p, double l1, double l2, double *M)
{
    double *u = (double *)malloc(sizeof(double) * (size_t)K);
    size_t pp = (size_t)p * p;
    for (int i = 0; i < p; ++i) {
        for (int j = i; j < p; ++j) {
            if (i == j) {
                for (int k = 0; k < K; ++k) M[k * pp + (size_t)i * p + i] = X[k * pp + (size_t)i * p + i];
                continue;
            }
            double ss = 0.0;
            for (int k = 0; k < K; ++k) { u[k] = soft(X[k * pp + (size_t)i * p + j], l1); ss += u[k] * u[k]; }
            double nrm = sqrt(ss);
            double a = (nrm > l2) ? nrm : l2;
            for (int k = 0; k < K; ++k) {
                double r = (u[k] * (a - l2)) / a;
                M[k * pp + (size_t)i * p + j] = r;
                M[k * pp + (size_t)j * p + i] = r;
            }
        }
    }
    free(u);
}

/* regulariser value; reg: 0 = GGL, 1 = FGL */
double gg_pval(const double *X, int K, int p, double l1, double l2, int reg)
{
    size_t pp = (size_t)p * p;
    double res = 0.0;
    for (int i = 0; i < p; ++i) {
        for (int j = i + 1; j < p; ++j) {
            double n1 = 0.0, n2 = 0.0;
            for (int k = 0; k < K; ++k) {
                double a = X[k * pp + (size_t)i * p + j];
                n1 += fabs(a);
                if (reg == 0) n2 += a * a;
                else if (k > 0) n2 += fabs(a - X[(k - 1) * pp + (size_t)i * p + j]);
            }
            if (reg == 0) n2 = sqrt(n2);
            res += l1 * n1 + l2 * n2;
        }
    }
    return 2.0 * res;
}
