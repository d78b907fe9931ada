/*
 * gglasso_b200.h -- C ABI of the B200 (sm_100a) implementation of GGLasso's ADMM hot path.
 *
 * The reference (fabian-sp/GGLasso v0.2.1) is pure Python and has no FFI; its boundary for this
 * path is three Python callables (ADMM_MGL, ADMM_SGL, block_SGL).  The host side of this repo
 * (gglasso_b200/solver/) re-implements those callables and drives the entry points below through
 * ctypes.  Each entry point cites the reference code it replaces (paths relative to the
 * reference repository root).
 *
 * Conventions
 *   - all matrices are FP64, C-contiguous stacks (M, p, p) in DEVICE memory; M = number of
 *     matrices in the launch (K instances of one MGL problem, or M independent SGL problems);
 *   - `mpp` = matrices per ADMM problem (K for MGL, 1 for SGL batches); problem q = m / mpp;
 *   - `ctrl` = device control block, GG_CTRL_STRIDE doubles per problem (layout below); rho, the
 *     pending dual rescale and the done flag live there so the loop never needs the host;
 *   - `stream` is a cudaStream_t passed as void*;
 *   - return value: 0 on success, a positive cudaError_t, or a negative argument error;
 *   - no entry point allocates device memory: scratch comes from the caller
 *     (gg_eigh_workspace_bytes).
 */
#ifndef GGLASSO_B200_H
#define GGLASSO_B200_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GG_CTRL_STRIDE 16 /* doubles per problem */
/* ctrl[0]=rho  ctrl[1]=pending X scale  ctrl[2]=done  ctrl[3]=iterations done
 * ctrl[4..7]=r,s,eps_pri,eps_dual of the last iteration  ctrl[8]=1 if 'optimal'
 * ctrl[9], ctrl[10]: if ctrl[9] > 0, lambda1 / lambda2 of gg_prox_mgl and gg_prox_band are read from here instead of
 * the call arguments (lets a captured CUDA graph of one iteration serve every point of a lambda grid) */
#define GG_HIST_STRIDE 5  /* per iteration: r, s, eps_pri, eps_dual, rho */
#define GG_NPART 5        /* partial sums per CTA feeding gg_stop_update */

int gg_version(void);

/* number of kernels this library has launched in the calling process so far (monotonic; bench.py reports the
 * difference over its timed region as `gpu_launches`).  New; no reference counterpart. */
long long gg_launch_count(void);

/* W = Theta - L - X - (n_k/rho) S ; also folds the pending rescale of X (rho_old/rho_new) into X.
 * reference: src/gglasso/solver/admm_solver.py:180,236 ; single_admm_solver.py:163,205.
 * L may be NULL (non-latent); nk may be NULL (all ones) or a device array (M). */
int gg_build_w(const double* Theta, const double* L, double* X, const double* S, const double* nk,
               const double* ctrl, int M, int p, int mpp, double* W, void* stream);

/* Batched symmetric eigendecomposition (replaces np.linalg.eigh at admm_solver.py:181,199 and
 * single_admm_solver.py:164,174).  A (M,p,p) is overwritten by Vt: row c of Vt[m] is the unit
 * eigenvector belonging to D[m][c] (order and signs arbitrary).  vectors=0 skips normalisation
 * (eigenvalues only).  Path selection: p <= 48 shared-memory Jacobi (one CTA per matrix), larger p Householder
 * tridiagonalisation + divide & conquer + back-transformation.  block_nb2 = 0: default; 1: force the shared-memory
 * Jacobi kernel (p <= 160); 32/64/128: block-Jacobi with that many rows per block pair (p > 160).
 * tol<=0, max_sweeps<=0: defaults.  quad_tol: a sweep whose largest measured off-diagonal cosine is
 * below quad_tol is taken as the last one (quadratic convergence); 0 disables.  info[0] (host) = sweeps.
 * Vt_warm (optional, (M,p,p), rows orthonormal; used for p <= 160): warm start from the eigenvectors of the
 * previous ADMM iteration (identity for the first call); overwritten with the new eigenvectors. */
size_t gg_eigh_workspace_bytes(int M, int p);
int gg_eigh(double* A, double* D, int M, int p, const double* ctrl, int mpp, void* ws, size_t ws_bytes,
            int vectors, int block_nb2, double tol, int max_sweeps, double quad_tol, int* info, double* Vt_warm,
            void* stream);

/* Stage 1 of the large-p eigensolver only (Householder tridiagonalisation of the batch), for profiling:
 * which = 0 full sytrd, 1 only the per-column kernels, 2 only the trailing-matrix (symv + rank-2 update)
 * kernels.  Results of which != 0 are meaningless; A is destroyed.  p must exceed 32. */
int gg_sytrd_profile(double* A, double* D, int M, int p, void* ws, size_t ws_bytes, int which, void* stream);

/* Write-back depth q of the trailing-matrix kernel: pass j stores the updated trailing block only when q rank-2
 * updates are pending (every q-th pass moves 16 B per element, the others 8 B).  Needed to state the algorithmic
 * bytes of a launch; 1 <= q <= 4 (environment GG_TR_LAZY, default 3). */
int gg_sytrd_write_depth(void);
/* diagnostics (GG_TR_TIMING=1): nanoseconds per phase of the blocked tridiagonalisation's panel kernel, summed by one
 * CTA since the last call (16 counters, see gg_sytrd_blocked.cuh); reads and clears. */
int gg_sytrd_phase_clock(unsigned long long* out16);

/* Out = V diag(f(D)) V^T with V^T = Vt from gg_eigh (FP64 tensor cores, exactly symmetric result).
 * mode 0: f = phi+(d, beta) = (sqrt(d^2+4 beta)+d)/2   src/gglasso/solver/ggl_helper.py:272-303
 * mode 1: f = max(d-beta, 0)                            src/gglasso/solver/ggl_helper.py:29-36
 * mode 2: f = d
 * beta = bnum[m]/rho (bnum NULL -> 1; ctrl NULL -> rho = 1). */
int gg_recon(const double* Vt, const double* D, const double* bnum, const double* ctrl, int mpp, int mode,
             int M, int p, double* Out, void* stream);

/* Theta = prox_od_1norm(Omega + L + X, lam/rho)   src/gglasso/solver/ggl_helper.py:16-27,
 * single_admm_solver.py:169.  lam_mat (M,p,p) optional elementwise penalty (lambda1*lambda1_mask).
 * C == NULL (non-latent): fused with X += Omega - Theta (single_admm_solver.py:177) and the partial
 * sums for gg_stop_update (partials: M * gg_sgl_nparts(p,M) * GG_NPART doubles).
 * C != NULL (latent): writes C = Theta - X - Omega (single_admm_solver.py:173), X untouched.
 * pvec (optional, device int[M]): true size of problem m inside a padded (M,p,p) ragged batch (block_SGL's
 * connected components, single_admm_solver.py:441-459); entries beyond it are excluded from the norms. */
int gg_sgl_nparts(int p, int M);
int gg_prox_sgl(const double* Omega, const double* Omega_prev, const double* L, double* X, double* Theta,
                double* C, const double* ctrl, double lam, const double* lam_mat, int M, int p,
                double* partials, const int* pvec, void* stream);

/* Functional SGL: Theta = prox_sum_Frob(Omega + L + X, Mblk, lam/rho)   src/gglasso/solver/ggl_helper.py:45-66,
 * src/gglasso/solver/functional_sgl_admm.py:148.  Off-diagonal Mblk x Mblk blocks are shrunk in Frobenius norm.
 * blk_nrm: scratch, M * (p/Mblk)^2 doubles.  Same fusion / latent convention as gg_prox_sgl. */
int gg_prox_fsgl(const double* Omega, const double* Omega_prev, const double* L, double* X, double* Theta,
                 double* C, const double* ctrl, double lam, int Mblk, int M, int p, double* partials,
                 double* blk_nrm, void* stream);

/* Theta = prox_p(Omega + L + X, lambda1/rho, lambda2/rho, reg)   src/gglasso/solver/ggl_helper.py:190-207
 * reg 0 = GGL (ggl_helper.py:68-71,38-43), 1 = FGL (ggl_helper.py:131-134, fgl_helper.py:11-68).
 * Same fusion/latent convention as gg_prox_sgl; partials: gg_mgl_ntile(p)^2 * GG_NPART doubles,
 * zero-initialised by the caller once. */
int gg_mgl_ntile(int p);
int gg_prox_mgl(const double* Omega, const double* Omega_prev, const double* L, double* X, double* Theta,
                double* C, const double* ctrl, double lambda1, double lambda2, int reg, int K, int p,
                double* partials, void* stream);

/* The same prox + dual update + residual sums on the UPPER triangles only (non-latent loop; 2.5 A bytes instead of
 * 5 A): Theta and X are written for i <= j, the residual sums count off-diagonal entries twice.  The consumers inside
 * the loop (gg_build_w_upper, the tridiagonal path of gg_eigh) read upper triangles only; gg_mirror_upper fills the
 * lower triangles of up to two (M,p,p) stacks afterwards -- the mirroring of prox_p (ggl_helper.py:198-205) done once
 * instead of every iteration.  partials: gg_mgl_upper_nparts(p) * GG_NPART doubles.
 * gg_build_w_upper: W = Theta - X - (n_k/rho) S for i <= j (admm_solver.py:180), one problem of K instances. */
int gg_jacobi_max(void);   /* largest p solved by the shared-memory Jacobi kernel (which reads the full matrix) */
int gg_mgl_upper_nparts(int p);
int gg_prox_mgl_upper(const double* Omega, const double* Omega_prev, double* X, double* Theta, const double* ctrl,
                      double lambda1, double lambda2, int reg, int K, int p, double* partials, void* stream);
int gg_build_w_upper(const double* Theta, double* X, const double* S, const double* nk, const double* ctrl, int K,
                     int p, double* W, void* stream);
int gg_mirror_upper(double* A0, double* A1, int M, int p, void* stream);

/* K-sharded MGL (one process per GPU): V = (Omega + L) + X on the local instances
 * (L may be NULL), and the cross-instance prox on a row band: V, Theta are (K, nb, p) slabs holding global rows
 * row0..row0+nb-1 of all K instances (after the all-to-all re-tile).  Same prox as gg_prox_mgl. */
int gg_add3(const double* Omega, const double* L, const double* X, double* V, size_t total, void* stream);
/* The two re-tile passes of the K-sharded solve, fused with their neighbours (no torch.cat / strided copies on the
 * hot path).  The p rows are cut into `world` contiguous bands (first p % world bands one row longer); band d of the
 * K_loc local instances is the contiguous block [K_loc][rows_d][p] at offset K_loc*p*lo_d of `send` / `recv`, i.e.
 * exactly the split layout of the two all-to-all exchanges.
 *   gg_pack_bands : send = (Omega + L) + X   (admm_solver.py:190 argument of prox_p), L may be NULL
 *   gg_unpack_dual: Theta (instance layout) from `recv`; C == NULL: X += Omega - Theta and the five residual partial
 *                   sums (admm_solver.py:208,316-331; partials (K_loc * gg_sgl_nparts(p,K_loc), 5));
 *                   C != NULL (latent): C = Theta - X - Omega (admm_solver.py:197), dual update follows the L step. */
int gg_pack_bands(const double* Omega, const double* L, const double* X, const double* ctrl, int K_loc, int p, int world,
                  double* send, void* stream);
int gg_unpack_dual(const double* recv, const double* Omega, const double* Omega_prev, double* X, double* Theta,
                   double* C, const double* ctrl, int K_loc, int p, int world, double* partials, void* stream);

/* Peer-memory variants of the two exchange steps (buffers in symmetric memory, peer_*[r] = device pointer of rank
 * r's buffer as mapped into this process, world <= 16): gg_pack_bands_p2p stores V = (Omega+L)+X of this rank's
 * instances (global indices k_lo ..) straight into every rank's band buffer (K_total, rows_d, p); gg_prox_band_p2p
 * runs the band prox and stores Theta of instance k into its owner's receive buffer (layout of gg_unpack_dual's
 * `recv`).  The caller places a cross-rank barrier after each.  Same arithmetic as gg_pack_bands / gg_prox_band. */
int gg_pack_bands_p2p(const double* Omega, const double* L, const double* X, const double* ctrl, int K_loc, int p,
                      int world, int k_lo, double* const* peer_band, void* stream);
int gg_prox_band_p2p(const double* V, double* const* peer_back, const double* ctrl, double lambda1, double lambda2,
                     int reg, int K, int nb, int p, int row0, int world, void* stream);
int gg_prox_band(const double* V, double* Theta, const double* ctrl, double lambda1, double lambda2, int reg,
                 int K, int nb, int p, int row0, void* stream);

/* ext_ADMM_MGL, non-conforming group graphical lasso (src/gglasso/solver/ext_admm_solver.py:191-273, :330-347,
 * :394-453).  The K matrices of size p_k are padded to (K,p,p); pvec[k] = p_k; one problem (ctrl block 0), rho fixed.
 *  gg_ext_theta : Theta = prox_od_1norm((Omega+L+X0+Lambda-X1)/2, lam1[k]/(2 rho)); C = Theta-X0-Omega if C != NULL
 *  gg_ext_lambda: Lambda = prox_2norm_G(Theta + X1, G, lambda2/rho); G is the reference's (2,Lg,K) int array
 *  gg_ext_dual  : X0 += Omega-Theta+L, X1 += Theta-Lambda, partial sums laid out for gg_stop_update (nparts =
 *                 K * gg_sgl_nparts(p,K)) */
int gg_ext_theta(const double* Omega, const double* L, const double* X0, const double* Lam, const double* X1,
                 const double* lam1, const double* ctrl, int K, int p, double* Theta, double* C, void* stream);
int gg_ext_lambda(const double* Theta, const double* X1, const int* G, int Lg, int K, int p, double lambda2,
                  const double* ctrl, double* Lam, void* stream);
int gg_ext_dual(double* X0, double* X1, const double* Omega, const double* Omega_prev, const double* Theta,
                const double* L, const double* Lam, const double* Lam_prev, const double* ctrl, const int* pvec,
                int K, int p, double* partials, void* stream);

/* X += Omega - Theta + L and partial sums (latent variants; L may be NULL for the K-sharded non-latent loop)   admm_solver.py:208, single_admm_solver.py:177.
 * sgl_order selects the association order of the reference's SGL expression. */
int gg_dual_update(double* X, const double* Omega, const double* Omega_prev, const double* Theta,
                   const double* L, const double* ctrl, int M, int p, int mpp, int sgl_order,
                   double* partials, void* stream);

/* Boyd residuals, stopping test and rho update on the device   admm_solver.py:216-246,316-331.
 * hist: nprob * hist_cap * GG_HIST_STRIDE doubles; pdim[q] = K(p^2+p)/2 of problem q. */
int gg_stop_update(const double* partials, int nparts, double* ctrl, double* hist, int hist_cap,
                   const double* pdim, double tol, double rtol, int update_rho, int nprob, void* stream);

/* apply the pending rescale to X after the loop (admm_solver.py:236) and clear it */
int gg_scale_pending(double* X, double* ctrl, int M, int p, int mpp, void* stream);

/* objective pieces for measure=True   ggl_helper.py:162-176,266-270 ; basic_linalg.py:20-33.
 * partials: 2 * gg_objective_nparts(p) doubles: {<Omega,S>, P(Theta)} per CTA.  reg -1 skips P. */
int gg_objective_nparts(int p);
int gg_objective(const double* Omega, const double* S, const double* Theta, double lambda1, double lambda2,
                 int reg, int K, int p, double* partials, void* stream);

/* max |A - A^T| partials (symmetry warnings, admm_solver.py:284-291); out: M * gg_sgl_nparts(p,M) doubles */
int gg_asym_max(const double* A, int M, int p, double* out, void* stream);

/* out[m] = min_i (a_ii - sum_{j!=i} |a_ij|): Gershgorin lower bound of the spectrum.  A positive bound proves
 * positive definiteness, letting the post-loop PD check (admm_solver.py:294-296) skip its eigendecomposition.
 * ws: at least the gg_eigh workspace of the same (M,p). */
int gg_gershgorin_min(const double* A, int M, int p, void* ws, size_t ws_bytes, double* out, void* stream);

/* Host-side execution of the exact device TV-prox routine (unit tests without a GPU). */
void gg_host_tv1d(double* v, int n, int stride, double lam);

#ifdef __cplusplus
}
#endif
#endif
