// gg_sytrd_blocked.cuh -- blocked Householder tridiagonalisation for the batched large-p eigensolver (sm_100a).
// Included by gg_tridiag.cu (needs TrWs, tr_block_allsum, gg_dmma, gg_cp_async8).
//
// Replaces the first stage of np.linalg.eigh (LAPACK dsytrd) at admm_solver.py:181,199 / single_admm_solver.py:164,174.
// LAPACK's blocked scheme (dlatrd + dsyr2k), laid out for a batch of M matrices on 148 SMs:
//
//   panel   sytrd_panel_kernel   ONE launch reduces PB columns of every matrix.  Each matrix belongs to a thread-block
//                                CLUSTER of CS CTAs (CS in {4,6,8,16}); the upper triangle of the trailing matrix A0
//                                (as of the panel start) is dealt to the CTAs in blocks of 8 rows, block-cyclically.
//                                As much of a CTA's strip as fits is kept RESIDENT in shared memory for the whole
//                                panel, the rest is streamed once per column -- with an L2 evict_last hint for as
//                                much of it as the L2 can hold and evict_first for the remainder, so that the
//                                cyclic sweep does not thrash the cache.  Per column j (v_j known to every CTA):
//                                  y   = A0 v_j          every CTA multiplies its strip, both ways (row sums and
//                                                        mirrored column sums: one pass over the half matrix)
//                                  X1  reduce-scatter of the partial y over distributed shared memory (+ partial
//                                      W^T v, V^T v, and the pivot row of V, W from its owner)
//                                  y  -= V (W^T v) + W (V^T v)   (pending panel updates, dlatrd), p = tau y,
//                                  z   = (A0[j+1,:] - V W[j+1,:]^T - W V[j+1,:]^T) - p        on the own entries
//                                  X2  all-gather of z, of the partial p.v and of p at the pivot row
//                                  every CTA: alpha = (tau/2) p.v, next row a' = z + (2 alpha - p_piv) v_j, its norm,
//                                  v_{j+1}; own rows: w_j = p - alpha v_j
//                                two cluster barriers per column instead of two kernel launches; no global memory
//                                traffic inside the column loop except the streamed part of the strip.
//   update  sytrd_syr2k_kernel   A[T,T] -= V W^T + W V^T on the upper triangle of the remaining block, FP64 tensor
//                                cores (DMMA m8n8k4), one CTA per 64x64 tile, K = 2 PB.
//
// The last TR_TAIL columns are finished by tr_tail_kernel as before.
#pragma once
#include <cooperative_groups.h>
namespace cg = cooperative_groups;

#define PB_THREADS 256
#define PB_CG 4                        // column groups of the strip product (a warp takes every PB_CG-th 32-column chunk)
#define PB_WR (PB_THREADS / 32 / PB_CG)  // row groups: warps = PB_WR x PB_CG
#define PB_H 8                 // rows per ownership block
#define PB_MAXCS 16
#define PB_NMAX 2048           // largest matrix size on this path (column accumulators live in registers)

struct PanelGeom {
    int n, p0, nbk;            // matrix size, first column of the panel, columns in this panel
    int t, TP;                 // trailing size n - p0 - 1, padded to a multiple of 32
    int nblk;                  // number of 8-row blocks of the trailing matrix
    int NO;                    // own-slot capacity per CTA: 8 * ceil(nblk / CS)
    int res_doubles;           // shared-memory doubles available for the resident part of the strip
};

// shared-memory carve-up (in doubles) -- must match sytrd_panel_fixed_doubles() on the host
template <int NB>
struct PanelSmem {
    double *vfull, *vnext, *ycol, *rowp, *yin, *gin, *gbuf, *gtot, *bc2, *piv, *Vs, *Ws, *pown, *vown, *red, *arows, *sc, *strip;
    int *boff, *otab, *itab;
    __device__ __forceinline__ void carve(double* base, int TP, int NO, int CS, int nqmax)
    {
        double* q = base;
        vfull = q; q += TP;           // v_j, every entry
        vnext = q; q += TP;           // z of the next pivot row, every entry (written by all CTAs of the cluster)
        ycol = q; q += PB_WR * TP;    // column sums of the strip product, one slice per warp row group
        rowp = q; q += PB_CG * NO;    // row sums, one slice per warp column group
        yin = q; q += CS * NO;        // partial y of the own rows from every CTA
        gin = q; q += CS * 2 * NB;    // partial [W^T v ; V^T v] from every CTA
        gbuf = q; q += 2 * NB;
        gtot = q; q += 2 * NB;
        bc2 = q; q += PB_MAXCS;       // partial p.v from every CTA
        piv = q; q += 2 * (2 * NB + 2);   // pivot row, two buffers (column parity): [0] p, [1..NB] V, [NB+1..2NB] W
        Vs = q; q += NO * (NB + 1);   // own rows of the panel's V and W
        Ws = q; q += NO * (NB + 1);
        pown = q; q += NO;
        vown = q; q += NO;            // v_j on the own rows
        red = q; q += 64;
        arows = q; q += NB * NO;      // own entries of the panel's pivot rows of A0
        sc = q; q += 4 * NB;          // d, e, tau of the panel and the diagonal entries of the pivot rows
        boff = (int*)q; q += (nqmax + 1) / 2 + 1;
        otab = (int*)q; q += TP / 2;  // entry i -> (owner << 16) | slot
        itab = (int*)q; q += (NO + 1) / 2 + 1;   // own slot -> entry
        strip = q;
    }
};

static inline size_t sytrd_panel_fixed_doubles(int NB, int TP, int NO, int CS, int nqmax)
{
    return (size_t)(2 + PB_WR) * TP + PB_CG * NO + (size_t)CS * NO + (size_t)CS * 2 * NB + 4 * NB + PB_MAXCS + 2 * (2 * NB + 2) +
           (size_t)2 * NO * (NB + 1) + 2 * NO + 64 + (size_t)NB * NO + 4 * NB + (nqmax + 1) / 2 + 1 + TP / 2 +
           (NO + 1) / 2 + 1;
}

// 8 per-lane partial sums -> lane l holds the warp total of element (l >> 2)
__device__ __forceinline__ double pb_reduce8(double (&r)[8], int lane)
{
    double s4[4], s2[2];
    const bool b4 = (lane & 16) != 0, b3 = (lane & 8) != 0, b2 = (lane & 4) != 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const double send = b4 ? r[k] : r[k + 4];
        const double keep = b4 ? r[k + 4] : r[k];
        s4[k] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
    }
#pragma unroll
    for (int k = 0; k < 2; ++k) {
        const double send = b3 ? s4[k] : s4[k + 2];
        const double keep = b3 ? s4[k + 2] : s4[k];
        s2[k] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
    }
    const double send = b2 ? s2[0] : s2[1];
    const double keep = b2 ? s2[1] : s2[0];
    double s = keep + __shfl_xor_sync(0xffffffffu, send, 4);
    s += __shfl_xor_sync(0xffffffffu, s, 2);
    s += __shfl_xor_sync(0xffffffffu, s, 1);
    return s;
}

// One 8-row block of the strip times v, both ways, for the column chunks ch = g, g + 4, ... of this warp (32 columns
// per chunk, one per lane).  src: element (i, cc) of the block at src[i*pitch + (cc - r0)].
//   racc[i]  += sum_cc a(i,cc) v[cc]            (cc >= r0+i, diagonal included)
//   ycw[cc]  += sum_i  a(i,cc) v[r0+i]          (cc >  r0+i)      ycw: this warp row-group's column sums (shared)
// The kernel is bound by instruction issue, so the chunks strictly between the diagonal chunk and the ragged last
// chunk of a full block run without any predicate (two chunks per trip: 16 independent loads in flight per lane).
template <bool RES>
__device__ __forceinline__ double pb_ld(const double* p, unsigned long long pol)
{
    if (RES) return *p;
    double v;
    asm volatile("ld.global.L1::no_allocate.L2::cache_hint.f64 %0, [%1], %2;" : "=d"(v) : "l"(p), "l"(pol));
    return v;
}

template <bool RES>
__device__ __forceinline__ void pb_block_matvec(const double* __restrict__ src, int pitch, int r0, int t, int g, int lane,
                                                const double* __restrict__ vfull, double* __restrict__ ycw,
                                                double (&racc)[8], unsigned long long pol)
{
    const int ch0 = r0 >> 5;                         // chunk holding the diagonal of this block
    const int chl = (t - 1) >> 5;                    // last chunk with valid columns
    const int rows = min(PB_H, t - r0);
    const bool ragged = (t & 31) != 0;
    double vr[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) vr[i] = vfull[r0 + i];
    int ch = g + PB_CG * ((ch0 - g + PB_CG - 1) / PB_CG);   // first chunk of this warp at or right of the diagonal chunk
    // masked trip: diagonal chunk, ragged last chunk, or a block with fewer than 8 rows
    auto masked = [&](int chm) {
        const int cc = 32 * chm + lane;
        const bool on = cc < t;
        double a[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const bool ok = on && (i < rows) && (cc >= r0 + i);
            a[i] = ok ? pb_ld<RES>(src + (size_t)i * pitch + (cc - r0), pol) : 0.0;
        }
        const double vcc = on ? vfull[cc] : 0.0;
        double cs = 0.0;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            racc[i] = fma(a[i], vcc, racc[i]);
            cs = fma((cc == r0 + i) ? 0.0 : a[i], vr[i], cs);      // the diagonal is counted once (row sum)
        }
        if (on) ycw[cc] += cs;
    };
    if (rows < PB_H) {
        for (; ch <= chl; ch += PB_CG) masked(ch);
        return;
    }
    if (ch == ch0) { masked(ch); ch += PB_CG; }
    const int chi = (ragged ? chl - 1 : chl);        // last chunk that may take the unmasked path
    const double* p = src + (32 * ch + lane - r0);
    for (; ch + PB_CG <= chi; ch += 2 * PB_CG, p += 64 * PB_CG) {
        double a[8], b[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) a[i] = pb_ld<RES>(p + (size_t)i * pitch, pol);
#pragma unroll
        for (int i = 0; i < 8; ++i) b[i] = pb_ld<RES>(p + (size_t)i * pitch + 32 * PB_CG, pol);
        const int cc = 32 * ch + lane;
        const double va = vfull[cc], vb = vfull[cc + 32 * PB_CG];
        double ca = 0.0, cb = 0.0;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            racc[i] = fma(a[i], va, racc[i]);
            ca = fma(a[i], vr[i], ca);
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            racc[i] = fma(b[i], vb, racc[i]);
            cb = fma(b[i], vr[i], cb);
        }
        ycw[cc] += ca;
        ycw[cc + 32 * PB_CG] += cb;
    }
    if (ch <= chi) {
        double a[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) a[i] = pb_ld<RES>(p + (size_t)i * pitch, pol);
        const int cc = 32 * ch + lane;
        const double va = vfull[cc];
        double ca = 0.0;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            racc[i] = fma(a[i], va, racc[i]);
            ca = fma(a[i], vr[i], ca);
        }
        ycw[cc] += ca;
        ch += PB_CG;
    }
    if (ch <= chl) masked(ch);                       // ragged last chunk
}

// phase clock of the panel kernel (diagnostics, GG_TR_TIMING=1): nanoseconds summed over all columns by thread 0 of
// the first CTA; [0] X3 exchange  [1] Householder + g partials  [2] strip product  [3] X1 exchange  [4] correction
// [5] X2 exchange  [6] w and next row  [7] panel prologue (strip load)  [8] columns
__device__ unsigned long long pb_phase_ns[16];
__device__ __forceinline__ unsigned long long pb_now()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
#define PB_STAMP(slot)                                                   \
    do {                                                                 \
        if (clock_on) {                                                  \
            const unsigned long long now_ = pb_now();                    \
            acc_ns[slot] += now_ - last_ns;                              \
            last_ns = now_;                                              \
        }                                                                \
    } while (0)

template <int NB>
__global__ void __launch_bounds__(PB_THREADS, 1)
sytrd_panel_kernel(const double* __restrict__ A, int n, int p0, int nbk, TrWs ws, double* __restrict__ Wp,
                   const int* __restrict__ skip, int res_doubles, int l2_doubles, int timing)
{
    extern __shared__ __align__(16) double pbsm[];
    cg::cluster_group cluster = cg::this_cluster();
    const int CS = (int)cluster.num_blocks();
    const int c = (int)cluster.block_rank();
    const int m = blockIdx.x / CS;
    if (skip && skip[m]) return;                    // uniform over the cluster: no barrier is left waiting
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const bool clock_on = timing && blockIdx.x == 0 && tid == 0;
    unsigned long long acc_ns[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0}, last_ns = clock_on ? pb_now() : 0;
    const int base = p0 + 1, t = n - base;
    const int TP = (t + 31) & ~31;
    const int nblk = (t + PB_H - 1) / PB_H;
    const int nqmax = (nblk + CS - 1) / CS;
    const int NO = PB_H * nqmax;
    const int nq = (c < nblk) ? (nblk - c + CS - 1) / CS : 0;      // my blocks: B = c + CS*q
    const int nown = PB_H * nq;
    PanelSmem<NB> S;
    S.carve(pbsm, TP, NO, CS, nqmax);
    const double* Am = A + (size_t)m * n * n;
    const double* A0 = Am + (size_t)base * n + base;              // trailing block as of the panel start
    double* Vh = ws.Vh + (size_t)m * n * n;
    double* Wm = Wp + (size_t)m * NB * n;
    const int LDV = NB + 1;
    unsigned long long pol_last, pol_first;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol_last));
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol_first));

    // ---- tables; resident / L2-pinned split of the strip: blocks in order until each budget is used up ------------
    for (int i = tid; i < TP; i += PB_THREADS) {
        const int B = i >> 3;
        S.otab[i] = ((B % CS) << 16) | (((B / CS) << 3) | (i & 7));
    }
    for (int sl = tid; sl < nown; sl += PB_THREADS) S.itab[sl] = PB_H * (c + CS * (sl >> 3)) + (sl & 7);
    if (tid == 0) {
        int used = 0, l2used = 0;
        for (int q = 0; q < nq; ++q) {
            const int r0 = PB_H * (c + CS * q);
            const int pitch = (t - r0 + 1) & ~1;
            const int need = PB_H * pitch;
            if (used + need <= res_doubles) { S.boff[q] = used; used += need; }
            else if (l2used + need <= l2_doubles) { S.boff[q] = -1; l2used += need; }      // streamed, kept in L2
            else S.boff[q] = -2;                                                           // streamed, evict first
        }
    }
    __syncthreads();
    for (int q = 0; q < nq; ++q) {
        const int off = S.boff[q];
        if (off < 0) continue;
        const int r0 = PB_H * (c + CS * q);
        const int len = t - r0, pitch = (len + 1) & ~1;
        for (int cc = tid; cc < pitch; cc += 2 * PB_THREADS) {
            double v[2][PB_H];
#pragma unroll
            for (int u = 0; u < 2; ++u)
#pragma unroll
                for (int i = 0; i < PB_H; ++i) {
                    const int cu = cc + u * PB_THREADS;
                    v[u][i] = (r0 + i < t && cu < len && cu >= i) ? __ldcg(A0 + (size_t)(r0 + i) * n + r0 + cu) : 0.0;
                }
#pragma unroll
            for (int u = 0; u < 2; ++u)
#pragma unroll
                for (int i = 0; i < PB_H; ++i)
                    if (cc + u * PB_THREADS < pitch) S.strip[off + i * pitch + cc + u * PB_THREADS] = v[u][i];
        }
    }
    // own entries of the pivot rows of A0 used inside the panel (row local k is eliminated at column k + 1) and their
    // diagonal entries: staged now, so that no global load is outstanding at any cluster barrier of the column loop
    for (int sl = tid; sl < nown; sl += PB_THREADS) {
        const int i = S.itab[sl];
        for (int k0 = 0; k0 < nbk - 1; k0 += 8) {
            double v[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int k = k0 + u;
                v[u] = (k < nbk - 1 && i > k && i < t) ? __ldcg(A0 + (size_t)k * n + i) : 0.0;
            }
#pragma unroll
            for (int u = 0; u < 8; ++u)
                if (k0 + u < nbk - 1) S.arows[(k0 + u) * NO + sl] = v[u];
        }
    }
    if (tid < nbk - 1) S.sc[3 * NB + tid] = __ldcg(A0 + (size_t)tid * n + tid);
    if (tid == 0) S.sc[0] = Am[(size_t)p0 * n + p0];               // d[p0]
    // the first row to eliminate, row p0 right of the diagonal: every CTA reads all of it (z of "column -1")
    for (int i = tid; i < TP; i += PB_THREADS) {
        S.vnext[i] = (i < t) ? Am[(size_t)p0 * n + base + i] : 0.0;
        S.vfull[i] = 0.0;
    }
    cluster.sync();                                 // every CTA of the cluster is running: remote shared memory is valid
    PB_STAMP(7);

    double tau_prev = 0.0;                          // tau of column jj - 1
    for (int jj = 0; jj <= nbk; ++jj) {
        // ================= G: finish column jj-1 (alpha, w), form the next row a' and v_jj ==========================
        // the pivot-row buffer alternates with the column parity: the owner of the next pivot row may already be
        // writing this column's values (X1) while a slower CTA still reads the previous ones here
        const double* pivp = S.piv + ((jj + 1) & 1) * (2 * NB + 2);     // written during column jj-1
        double* pivc = S.piv + (jj & 1) * (2 * NB + 2);                 // written during this column
        double al = 0.0, beta2 = 0.0;
        if (jj > 0) {
            double dot = 0.0;
            for (int r = 0; r < CS; ++r) dot += S.bc2[r];
            al = 0.5 * tau_prev * dot;
            beta2 = 2.0 * al - pivp[0];
            for (int sl = tid; sl < nown; sl += PB_THREADS) {
                const int i = S.itab[sl];
                S.Ws[sl * LDV + jj - 1] = (i >= jj - 1 && i < t) ? S.pown[sl] - al * S.vfull[i] : 0.0;      // w_{jj-1}
            }
        }
        if (jj == nbk) break;
        double areg[PB_NMAX / PB_THREADS];
        double part = 0.0;
#pragma unroll
        for (int u = 0; u < PB_NMAX / PB_THREADS; ++u) {
            const int i = tid + PB_THREADS * u;
            areg[u] = 0.0;
            if (i < t && i >= jj) {
                areg[u] = fma(beta2, S.vfull[i], S.vnext[i]);                      // a' = z + (2 alpha - p_piv) v
                if (i > jj) part = fma(areg[u], areg[u], part);
            }
        }
        const double alpha = fma(beta2, S.vfull[jj], S.vnext[jj]);
        const double xn2 = tr_block_allsum(part, S.red);                           // (barrier: old v no longer needed)
        double beta = alpha, scale = 0.0, tau_j = 0.0;
        if (xn2 > 0.0) {
            beta = -copysign(sqrt(alpha * alpha + xn2), alpha);
            tau_j = (beta - alpha) / beta;
            scale = 1.0 / (alpha - beta);
        }
#pragma unroll
        for (int u = 0; u < PB_NMAX / PB_THREADS; ++u) {
            const int i = tid + PB_THREADS * u;
            if (i < TP) S.vfull[i] = (i == jj) ? 1.0 : ((i > jj && i < t) ? areg[u] * scale : 0.0);
        }
        if (tid == 0) {
            S.sc[2 * NB + jj] = tau_j;
            S.sc[NB + jj] = beta;
        }
        __syncthreads();
        for (int sl = tid; sl < nown; sl += PB_THREADS) {
            const int i = S.itab[sl];
            const double v = (i < t) ? S.vfull[i] : 0.0;
            S.Vs[sl * LDV + jj] = v;
            S.vown[sl] = v;
        }
        __syncthreads();
        // partial g = [W^T v ; V^T v] over the own rows (columns k < jj): eight lanes per column, each over every
        // eighth own row, combined with three shuffles
        for (int k = tid >> 3; k < 2 * jj; k += PB_THREADS / 8) {
            const double* P = (k < jj) ? (S.Ws + k) : (S.Vs + (k - jj));
            double a0 = 0.0, a1 = 0.0;
            int sl = tid & 7;
            for (; sl + 8 < nown; sl += 16) {
                a0 = fma(P[sl * LDV], S.vown[sl], a0);
                a1 = fma(P[(sl + 8) * LDV], S.vown[sl + 8], a1);
            }
            if (sl < nown) a0 = fma(P[sl * LDV], S.vown[sl], a0);
            double acc = a0 + a1;
            const unsigned gm = 0xffu << (lane & 24);          // the eight lanes that share this column (same trip count)
            acc += __shfl_xor_sync(gm, acc, 1);
            acc += __shfl_xor_sync(gm, acc, 2);
            acc += __shfl_xor_sync(gm, acc, 4);
            if ((tid & 7) == 0) S.gbuf[k] = acc;      // gbuf[0..jj) = W^T v, gbuf[jj..2jj) = V^T v
        }
        PB_STAMP(1);
        // ================= A: y = A0 v over the strip ==============================================================
        {
            const int wr = warp / PB_CG, g = warp % PB_CG;
            double* ycw = S.ycol + wr * TP;
            for (int cc = 32 * g + lane; cc < TP; cc += 32 * PB_CG) ycw[cc] = 0.0;      // this warp's columns of its row group
            __syncwarp();
            for (int q = wr; q < nq; q += PB_WR) {
                const int r0 = PB_H * (c + CS * q);
                double racc[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) racc[i] = 0.0;
                if (r0 + PB_H > jj) {                 // rows above the current column carry v = 0
                    const int off = S.boff[q];
                    if (off >= 0)
                        pb_block_matvec<true>(S.strip + off, (t - r0 + 1) & ~1, r0, t, g, lane, S.vfull, ycw, racc, 0ull);
                    else
                        pb_block_matvec<false>(A0 + (size_t)r0 * n + r0, n, r0, t, g, lane, S.vfull, ycw, racc,
                                               off == -1 ? pol_last : pol_first);
                }
                const double tot = pb_reduce8(racc, lane);
                if ((lane & 3) == 0) S.rowp[g * NO + q * PB_H + (lane >> 2)] = tot;
            }
        }
        __syncthreads();
        PB_STAMP(2);
        // ================= X1: reduce-scatter of the partial y, all-gather of the partial g, pivot row of V, W =====
        for (int i = 2 * tid; i < t; i += 2 * PB_THREADS) {        // pairs (i, i+1): same 8-row block, same owner
            const int ot = S.otab[i], owner = ot >> 16, sl = ot & 0xffff;
            double2 v = make_double2(0.0, 0.0);
#pragma unroll
            for (int w = 0; w < PB_WR; ++w) {
                const double2 yc = *reinterpret_cast<const double2*>(S.ycol + w * TP + i);
                v.x += yc.x; v.y += yc.y;
            }
            if (owner == c) {
#pragma unroll
                for (int w = 0; w < PB_CG; ++w) {
                    const double2 rp = *reinterpret_cast<const double2*>(S.rowp + w * NO + sl);
                    v.x += rp.x; v.y += rp.y;
                }
            }
            *reinterpret_cast<double2*>(cluster.map_shared_rank(S.yin, owner) + c * NO + sl) = v;
        }
        if (tid < 2 * jj)
            for (int r = 0; r < CS; ++r) cluster.map_shared_rank(S.gin, r)[c * 2 * NB + tid] = S.gbuf[tid];
        const int otp = S.otab[jj], powner = otp >> 16, psl = otp & 0xffff;       // next pivot row: local row jj
        if (powner == c && tid >= 1 && tid <= 2 * NB) {
            double v;
            if (tid <= NB) v = (tid - 1 <= jj) ? S.Vs[psl * LDV + tid - 1] : 0.0;
            else v = (tid - 1 - NB < jj) ? S.Ws[psl * LDV + tid - 1 - NB] : 0.0;
            for (int r = 0; r < CS; ++r) cluster.map_shared_rank(pivc, r)[tid] = v;
        }
        cluster.sync();
        PB_STAMP(3);
        // ================= C: y on the own rows, panel correction, p = tau y, z of the next pivot row ===============
        if (tid < 2 * jj) {
            double gsum = 0.0;
            for (int r = 0; r < CS; ++r) gsum += S.gin[r * 2 * NB + tid];
            S.gtot[tid] = gsum;
        }
        __syncthreads();
        double dp = 0.0;
        for (int sl = tid; sl < nown; sl += PB_THREADS) {
            const int i = S.itab[sl];
            double y = 0.0;
            for (int r = 0; r < CS; ++r) y += S.yin[r * NO + sl];
            double c0 = 0.0, c1 = 0.0, b0 = 0.0, b1 = 0.0;
            for (int k = 0; k < jj; ++k) {
                const double vk = S.Vs[sl * LDV + k], wk = S.Ws[sl * LDV + k];
                c0 = fma(vk, S.gtot[k], c0);                      // V (W^T v)
                c1 = fma(wk, S.gtot[jj + k], c1);                 // W (V^T v)
                b0 = fma(pivc[1 + k], wk, b0);                    // V[piv][k] W[i][k]
                b1 = fma(pivc[1 + NB + k], vk, b1);               // W[piv][k] V[i][k]
            }
            y -= c0 + c1;
            double pv = 0.0, z = 0.0;
            if (i >= jj && i < t) {
                pv = tau_j * y;
                dp = fma(pv, S.vfull[i], dp);
            }
            S.pown[sl] = pv;
            if (jj + 1 < nbk && i > jj && i < t) z = (S.arows[jj * NO + sl] - (b0 + b1)) - pv;
            if (jj + 1 < nbk)
                for (int r = 0; r < CS; ++r) cluster.map_shared_rank(S.vnext, r)[i] = z;
        }
        const double dp_c = tr_block_allsum(dp, S.red);               // (barrier: pown complete)
        PB_STAMP(4);
        // ================= X2: partial dots, p at the pivot row ======================================================
        if (tid < CS) cluster.map_shared_rank(S.bc2, tid)[c] = dp_c;
        if (powner == c && tid < CS) cluster.map_shared_rank(pivc, tid)[0] = S.pown[psl];
        cluster.sync();
        PB_STAMP(5);
        tau_prev = tau_j;
    }
    PB_STAMP(6);
    // ---- results of the panel: Householder vectors (rows of Vh), w vectors (for the trailing update), d, e, tau ----
    __syncthreads();
    for (int e = tid; e < nbk * nown; e += PB_THREADS) {
        const int k = e / nown, sl = e - k * nown;
        const int i = S.itab[sl];
        if (i >= k && i < t) {
            Vh[(size_t)(p0 + k) * n + base + i] = S.Vs[sl * LDV + k];
            Wm[(size_t)k * n + base + i] = S.Ws[sl * LDV + k];
        }
    }
    if (c == 0 && tid < nbk) {
        if (tid == 0) ws.d[(size_t)m * n + p0] = S.sc[0];
        ws.e[(size_t)m * n + p0 + tid] = S.sc[NB + tid];
        ws.tau[(size_t)m * n + p0 + tid] = S.sc[2 * NB + tid];
    }
    // diagonal entry of local row r (pivot of column r+1 of this panel): d = A0[r][r] - 2 sum_{k<=r} V[r][k] W[r][k],
    // from the finished panel rows of the CTA that owns row r
    if (tid < nbk - 1) {
        const int r = tid, ot = S.otab[r];
        if ((ot >> 16) == c) {
            const int sl = ot & 0xffff;
            double s2 = 0.0;
            for (int k = 0; k <= r; ++k) s2 = fma(S.Vs[sl * LDV + k], S.Ws[sl * LDV + k], s2);
            ws.d[(size_t)m * n + base + r] = S.sc[3 * NB + r] - 2.0 * s2;
        }
    }
    cluster.sync();                                   // no CTA exits while a peer may still write into its shared memory
    if (clock_on) {
        PB_STAMP(8);
        for (int q = 0; q < 9; ++q) atomicAdd(&pb_phase_ns[q], acc_ns[q]);
    }
}

// ---- trailing update on the FP64 tensor cores --------------------------------------------------------------------
// C[r][cc] -= sum_k V_k[r] W_k[cc] + W_k[r] V_k[cc]   for the upper-triangle tiles of C = A[g0:, g0:], g0 = p0 + nbk.
// TN form: As[k][r] = Acat[k][g0 + r], Bs[k][cc] = Bcat[k][g0 + cc] with Acat = [V ; W], Bcat = [W ; V] (rows of Vh
// and of the panel's W buffer, global column indexing).
#define SY_T 64
#define SY_KC 16
#define SY_LD (SY_T + 4)
template <int NB>
__global__ void __launch_bounds__(256)
sytrd_syr2k_kernel(double* __restrict__ A, int n, int p0, int nbk, const double* __restrict__ Vh,
                   const double* __restrict__ Wp, const int* __restrict__ skip, int nt)
{
    __shared__ __align__(16) double sm[2 * 2 * SY_KC * SY_LD];
    const int m = blockIdx.y;
    if (skip && skip[m]) return;
    int idx = blockIdx.x, I = 0;
    while (idx >= nt - I) { idx -= nt - I; ++I; }
    const int J = I + idx;
    const int g0 = p0 + nbk, tt = n - g0;
    const int i0 = I * SY_T, j0 = J * SY_T;
    const double* Vm = Vh + (size_t)m * n * n + (size_t)p0 * n + g0;       // row k: v_{p0+k}, columns from g0
    const double* Wm = Wp + (size_t)m * NB * n + g0;                      // row k: w_{p0+k}
    const int K = 2 * nbk;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int fr = lane >> 2, fc = lane & 3;
    const int wr = wid >> 1, wc = wid & 1;
    const int nchunks = (K + SY_KC - 1) / SY_KC;
    auto load_chunk = [&](int ch, int buf) {
        double* As = sm + (size_t)buf * SY_KC * SY_LD;
        double* Bs = sm + (size_t)(2 + buf) * SY_KC * SY_LD;
        const int k0 = ch * SY_KC;
        for (int e = tid; e < 2 * SY_KC * SY_T; e += 256) {
            const int which = e / (SY_KC * SY_T);
            const int rem = e - which * (SY_KC * SY_T);
            const int kk = rem / SY_T, x = rem - kk * SY_T;
            const int k = k0 + kk;
            const int col = (which ? j0 : i0) + x;
            double* dst = (which ? Bs : As) + kk * SY_LD + x;
            if (k < K && col < tt) {
                // A operand: [V ; W]; B operand: [W ; V]
                const bool useV = which ? (k >= nbk) : (k < nbk);
                const int kr = (k < nbk) ? k : k - nbk;
                gg_cp_async8(dst, (useV ? Vm : Wm) + (size_t)kr * n + col);
            } else *dst = 0.0;
        }
        gg_cp_commit();
    };
    double acc[2][4][2];
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) { acc[a][b][0] = 0.0; acc[a][b][1] = 0.0; }
    load_chunk(0, 0);
    for (int ch = 0; ch < nchunks; ++ch) {
        if (ch + 1 < nchunks) { load_chunk(ch + 1, (ch + 1) & 1); gg_cp_wait<1>(); }
        else gg_cp_wait<0>();
        __syncthreads();
        const double* As = sm + (size_t)(ch & 1) * SY_KC * SY_LD;
        const double* Bs = sm + (size_t)(2 + (ch & 1)) * SY_KC * SY_LD;
#pragma unroll
        for (int k0 = 0; k0 < SY_KC; k0 += 4) {
            double fa[2], fb[4];
#pragma unroll
            for (int a = 0; a < 2; ++a) fa[a] = As[(k0 + fc) * SY_LD + (wr * 2 + a) * 8 + fr];
#pragma unroll
            for (int b = 0; b < 4; ++b) fb[b] = Bs[(k0 + fc) * SY_LD + (wc * 4 + b) * 8 + fr];
#pragma unroll
            for (int a = 0; a < 2; ++a)
#pragma unroll
                for (int b = 0; b < 4; ++b) gg_dmma(acc[a][b][0], acc[a][b][1], fa[a], fb[b]);
        }
        __syncthreads();
    }
    // epilogue through shared memory: the read-modify-write of C runs over whole 512-byte row segments
    double* Cs = sm;                                  // 64 x 65 (the operand buffers are free now)
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            const int r = (wr * 2 + a) * 8 + fr, cc = (wc * 4 + b) * 8 + 2 * fc;
            Cs[r * 65 + cc] = acc[a][b][0];
            Cs[r * 65 + cc + 1] = acc[a][b][1];
        }
    __syncthreads();
    double* C = A + (size_t)m * n * n + (size_t)g0 * n + g0;
    for (int e0 = tid; e0 < SY_T * SY_T; e0 += 4 * 256) {
        double cur[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int e = e0 + 256 * u, r = i0 + (e >> 6), cc = j0 + (e & 63);
            cur[u] = (r < tt && cc < tt && cc >= r) ? __ldcg(C + (size_t)r * n + cc) : 0.0;
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int e = e0 + 256 * u, r = i0 + (e >> 6), cc = j0 + (e & 63);
            if (r < tt && cc < tt && cc >= r) C[(size_t)r * n + cc] = cur[u] - Cs[(e >> 6) * 65 + (e & 63)];
        }
    }
}

// ---- host side -----------------------------------------------------------------------------------------------------
struct SytrdBlockedPlan {
    int CS, NB;
    int max_smem;          // opt-in dynamic shared memory limit of the device, bytes
};

static int sytrd_blocked_env(const char* name, int dflt)
{
    const char* ev = getenv(name);
    return ev ? atoi(ev) : dflt;
}

static int sytrd_pick_cs(int M)
{
    int cs = sytrd_blocked_env("GG_TR_CS", 0);
    if (cs == 4 || cs == 6 || cs == 8 || cs == 16) return cs;
    const int cand[4] = {16, 8, 6, 4};
    for (int i = 0; i < 4; ++i)
        if (M * cand[i] <= 148) return cand[i];
    return 4;
}

template <int NB>
static int sytrd_panel_launch(const double* A, int n, int p0, int nbk, TrWs tw, double* Wp, const int* skip, int M, int CS,
                              int max_smem, cudaStream_t s)
{
    const int t = n - p0 - 1, TP = (t + 31) & ~31, nblk = (t + PB_H - 1) / PB_H, nqmax = (nblk + CS - 1) / CS;
    const int NO = PB_H * nqmax;
    const size_t fixed = sytrd_panel_fixed_doubles(NB, TP, NO, CS, nqmax);
    const size_t cap = (size_t)max_smem / sizeof(double);
    if (fixed + 64 > cap) return -6;
    // whole strip of the CTA with the most data (rank 0): sum over its blocks of 8 * even(t - r0)
    size_t strip = 0;
    for (int q = 0; q < nqmax; ++q) {
        const int r0 = PB_H * (CS * q);
        if (r0 < t) strip += (size_t)PB_H * ((t - r0 + 1) & ~1);
    }
    size_t res = cap - fixed;
    if (res > strip) res = strip;
    const size_t smem = (fixed + res) * sizeof(double);
    auto kern = sytrd_panel_kernel<NB>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem);
    if (e != cudaSuccess) return (int)e;
    if (CS > 8) {
        e = cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
        if (e != cudaSuccess) return (int)e;
    }
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = CS; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(M * CS); cfg.blockDim = dim3(PB_THREADS); cfg.dynamicSmemBytes = smem; cfg.stream = s;
    cfg.attrs = at; cfg.numAttrs = 1;
    gg_count_launch(1);
    // L2 share of one CTA for the streamed part of its strip (evict_last): GG_TR_L2MB megabytes over all CTAs
    const long long l2_total = (long long)sytrd_blocked_env("GG_TR_L2MB", 64) << 20;
    const int l2_doubles = (int)(l2_total / 8 / ((long long)M * CS));
    e = cudaLaunchKernelEx(&cfg, kern, A, n, p0, nbk, tw, Wp, skip, (int)res, l2_doubles,
                           sytrd_blocked_env("GG_TR_TIMING", 0));
    return e == cudaSuccess ? 0 : (int)e;
}

// diagnostics: read and clear the phase clock of the panel kernels
int gg_sytrd_phase_times(unsigned long long* out16)
{
    cudaError_t e = cudaMemcpyFromSymbol(out16, pb_phase_ns, sizeof(unsigned long long) * 16);
    if (e != cudaSuccess) return (int)e;
    unsigned long long z[16] = {0};
    e = cudaMemcpyToSymbol(pb_phase_ns, z, sizeof(z));
    return e == cudaSuccess ? 0 : (int)e;
}

// workspace of the W panel: (M, 32, n) doubles
static inline size_t sytrd_blocked_ws_doubles(int M, int n) { return (size_t)M * 32 * n; }

// Reduces columns [0, js) of every matrix; `which`: 0 = everything, 1 = panel kernels only, 2 = trailing updates only
// (profiling: the values are then meaningless, the memory traffic and launch sequence are the real ones).
static int sytrd_blocked_run(double* A, int n, int M, int js, TrWs tw, double* Wp, const int* skip, cudaStream_t s, int which)
{
    int dev = 0, max_smem = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    const int CS = sytrd_pick_cs(M);
    const int NB = sytrd_blocked_env("GG_TR_NB", 16) == 32 ? 32 : 16;
    for (int p0 = 0; p0 < js; p0 += NB) {
        const int nbk = (js - p0 < NB) ? js - p0 : NB;
        if (which != 2) {
            const int rc = (NB == 32) ? sytrd_panel_launch<32>(A, n, p0, nbk, tw, Wp, skip, M, CS, max_smem, s)
                                      : sytrd_panel_launch<16>(A, n, p0, nbk, tw, Wp, skip, M, CS, max_smem, s);
            if (rc != 0) return rc;
        }
        if (which != 1) {
            const int tt = n - (p0 + nbk);
            const int nt = (tt + SY_T - 1) / SY_T;
            gg_count_launch(1);
            if (NB == 32)
                sytrd_syr2k_kernel<32><<<dim3(nt * (nt + 1) / 2, M), 256, 0, s>>>(A, n, p0, nbk, tw.Vh, Wp, skip, nt);
            else
                sytrd_syr2k_kernel<16><<<dim3(nt * (nt + 1) / 2, M), 256, 0, s>>>(A, n, p0, nbk, tw.Vh, Wp, skip, nt);
        }
    }
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? 0 : (int)e;
}
