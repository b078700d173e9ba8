/*
 * gpx.h -- C ABI of libgpx.so, the B200-native (sm_100a) exact-Gaussian-process engine.
 *
 * Drop-in boundary: the reference (happyjin/Gaussian_process) has no FFI; its
 * boundary is the set of module-level NumPy functions its five scripts call.  libgpx replaces the
 * linear algebra *inside* those functions.  Every entry point below names the reference lines whose
 * arithmetic it replaces.  The Python host side (gaussian_process_b200/*.py, same module / function
 * names as the reference) binds these with ctypes; see INTEGRATION.md for the stub a maintainer of
 * the reference would add.
 *
 * Conventions
 *   - extern "C", plain pointers and sizes, no C++/torch types, no exceptions cross the boundary.
 *   - all matrices are IEEE float64, ROW-MAJOR, leading dimension `ld` in elements.
 *   - "device" entry points (gpx_*) take DEVICE pointers (e.g. torch tensor .data_ptr()) and run on
 *     the handle's stream; "host" entry points (gpx_host_*) take HOST pointers, do the H2D/D2H
 *     copies themselves and synchronise before returning.
 *   - return value: 0 = ok; >0 = LAPACK-style index (1-based) of the first non-positive pivot
 *     (the Python layer raises numpy.linalg.LinAlgError, as np.linalg.cholesky does); <0 = error
 *     (-k = bad argument k, or GPX_E_*), text via gpx_last_error().
 *   - dense kernels run on padded tiles: dimensions given to the device-level linear-algebra calls
 *     must be multiples of GPX_TILE (128).  gpx_cov_build pads for you (identity on the padded
 *     diagonal, zeros elsewhere) so a padded factorisation equals the unpadded one.
 */
#ifndef GPX_H
#define GPX_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GPX_VERSION 200
#define GPX_TILE 128

#define GPX_E_CUDA   (-100)   /* a CUDA runtime call failed            */
#define GPX_E_ARG    (-101)   /* generic bad argument                  */
#define GPX_E_NOMEM  (-102)   /* workspace allocation failed           */
#define GPX_E_NCCL   (-103)   /* NCCL not initialised / call failed    */

typedef struct gpx_ctx* gpx_handle;

/* covariance families (theta layout in brackets) */
enum gpx_cov_kind {
    GPX_COV_SE   = 0,  /* [sigma, l]   sigma^2 exp(-.5 d / l^2)            GP_regression.py:8-19      */
    GPX_COV_LIN  = 1,  /* [c]          (a-c).(b-c)                          GP_regression.py:22-33     */
    GPX_COV_PER  = 2,  /* [p, l]       exp(-2 sin^2(pi r/p)/l^2)            GP_regression.py:36-50     */
    GPX_COV_CO2  = 3   /* [theta1..11] SE + SE*periodic + RQ + SE + delta   CO2_example.py:9-94        */
};

/* flags for gpx_cov_build */
#define GPX_COV_SAME_X   1   /* X1 is X2: square block; theta11^2 delta (CO2) and `diag_add` go on the diagonal */
#define GPX_COV_LOWER    2   /* only tiles on/below the diagonal are computed; strictly-upper tiles are zeroed  */
#define GPX_COV_SKIP_UPPER 8 /* with GPX_COV_LOWER: strictly-upper tiles are not written at all (the caller's factorisation
                                never reads them and gpx_potrf zeroes them at the end): saves half of the HBM writes       */
#define GPX_COV_DELTA    4   /* n1 == n2 but X1 is not X2 (a square CROSS block): the CO2 delta term still goes on
                                the diagonal (CO2_example.py:58-66 tests the shape only); padding stays zero        */

/* ---- library / handle ------------------------------------------------------------------- */
int         gpx_version(void);
const char* gpx_last_error(void);
int         gpx_create(int device, gpx_handle* out);
int         gpx_destroy(gpx_handle h);
int         gpx_set_stream(gpx_handle h, void* cuda_stream);      /* cudaStream_t; NULL = default */
int         gpx_synchronize(gpx_handle h);
int64_t     gpx_padded_dim(int64_t n);                             /* round up to GPX_TILE */
/* number of kernels launched by this handle since creation (bench.py's gpu_launches) */
int64_t     gpx_launch_count(gpx_handle h);

/* ---- A1-A3: fused covariance builder ----------------------------------------------------
 * K[i,j] = k(X1[i,:], X2[j,:]; theta) for i<n1, j<n2; rows/cols up to (n1p,n2p) are padding:
 * zero, except K[i,i] = 1 for i >= n1 when GPX_COV_SAME_X (identity padding).  `diag_add` (the
 * reference's `+ s*np.eye(N)`, e.g. GP_regression.py:138) is added on the true diagonal when
 * SAME_X.  If dK != NULL, the ntheta derivative matrices dK/dtheta_j (SURVEY Appendix C;
 * tune_hyperparms_regression.py:48,54) are written in the same pass to dK + j*dk_stride.
 * Replaces GP_regression.py:18-19,32,48-49 and CO2_example.py:79-93. */
int gpx_cov_build(gpx_handle h, int kind, const double* X1, int64_t n1, const double* X2, int64_t n2,
                  int D, const double* theta_host, int ntheta, double diag_add, int flags,
                  double* K, int64_t n1p, int64_t n2p, int64_t ldk, double* dK, int64_t dk_stride);

/* One additive term of the CO2 composite as an element-wise map of a precomputed squared-distance matrix: replaces
 * kernel_1 .. kernel_4 (CO2_example.py:9-66), which take `sqdist` (and kernel_2 also `l2_norm`, NULL = sqrt(sqdist)).
 * term 1: t0^2 exp(-.5 d/t1^2); 2: t0^2 exp(-.5 d/t1^2 - 2 (sin(pi r)/t2)^2); 3: t0^2 (1 + .5 d/(t2 t1^2))^-t2;
 * 4: t0^2 exp(-.5 d/t1^2) + t2^2 [i == j] when rows == cols (the reference tests the shape only, :60).  Device pointers. */
int gpx_co2_term(gpx_handle h, int term, int64_t rows, int64_t cols, const double* sqdist, int64_t ldd,
                 const double* l2_norm, int64_t ldr, double t0, double t1, double t2, double* out, int64_t ldo);

/* ---- A4: Cholesky (np.linalg.cholesky call sites) ------------------------
 * In-place lower Cholesky of the n x n (n % 128 == 0) matrix A; on return the lower triangle holds
 * L and the strict upper triangle is zero (NumPy convention).  Blocked recursive right-looking:
 * 128x128 leaf factor (warp-shuffle/shared-memory kernel) + FP64 DMMA (mma.sync m8n8k4) TRSM/SYRK; for
 * 1024 <= n <= 16384 a three-stream look-ahead variant (substitution panel chain + grouped K = G*128 updates).
 * `dinv` (n/128 tiles of 128x128, device) receives the inverses of the diagonal blocks of L; they
 * are required by the solve / inverse routines below. */
int gpx_potrf(gpx_handle h, double* A, int64_t n, int64_t lda, double* dinv);
/* Same factorisation without the host synchronisation: returns after enqueueing; the first failing pivot of all
 * gpx_potrf_async calls since the last query is fetched and cleared (with a stream synchronisation) by gpx_potrf_info.  Lets several handles
 * (streams) factor independent matrices concurrently, e.g. the per-class B_c of GP_multi_classification.py:93. */
int gpx_potrf_async(gpx_handle h, double* A, int64_t n, int64_t lda, double* dinv);
int gpx_potrf_info(gpx_handle h, int* info_out);
/* Tuning / test hook: panels per group of the look-ahead factorisation used for 1024 <= n <= 16384 (0 = chosen by size, the
 * default; 1 = ungrouped; G > 1 = bulk trailing updates with K = G*128).  Process-wide. */
int gpx_potrf_set_group(int panels_per_group);

/* ---- A5: triangular solves (np.linalg.solve(L,.) / inv(L) call sites, row A5) -------------- */
/* x <- L^-1 x (trans=0) or L^-T x (trans=1), one right-hand side, HBM-bound blocked TRSV. */
int gpx_trsv(gpx_handle h, const double* L, int64_t n, int64_t ldl, const double* dinv, int trans, double* x);
/* B <- L^-1 B (trans=0) or L^-T B (trans=1); B is n x nrhs row-major, nrhs % 128 == 0. */
int gpx_trsm(gpx_handle h, const double* L, int64_t n, int64_t ldl, const double* dinv, int trans,
             double* B, int64_t nrhs, int64_t ldb);
/* Explicit inverses of the bs x bs diagonal blocks of L (bs = gpx_block_size_for(n): 1024, 512, 256 or 128) built from
 * the leaf inverses; D holds n/bs blocks of bs x bs doubles, work n*bs/4 doubles.  gpx_trsv_big / gpx_trsm_big are the
 * same solves as gpx_trsv / gpx_trsm with a bs/128 times shorter serial chain (tmp: bs, resp. bs*nrhs doubles). */
int gpx_block_size_for(int64_t n);
int gpx_block_inverses(gpx_handle h, const double* L, int64_t n, int64_t ldl, const double* dinv, int bs, double* D, double* work);
int gpx_trsv_big(gpx_handle h, const double* L, int64_t n, int64_t ldl, const double* D, int bs, int trans, double* x, double* tmp);
int gpx_trsm_big(gpx_handle h, const double* L, int64_t n, int64_t ldl, const double* D, int bs, int trans, double* B,
                 int64_t nrhs, int64_t ldb, double* tmp);
/* In-place inverse of the lower-triangular factor: L <- L^-1 (np.linalg.inv(L), tune...:144).
 * `work` must hold n*n/4 doubles. */
int gpx_trtri(gpx_handle h, double* L, int64_t n, int64_t ldl, const double* dinv, double* work);
/* out(lower) <- Linv^T Linv = (L L^T)^-1   (np.dot(inv(L.T), inv(L)), tune...:144); out != Linv. */
int gpx_lauum(gpx_handle h, const double* Linv, int64_t n, int64_t ldl, double* out, int64_t ldo);

/* general FP64 DMMA GEMM: C = alpha * op(A) op(B) + beta * C.  a_kmajor: A stored [M][K] (1) or
 * [K][M] (0); b_kmajor: B stored [N][K] (1) or [K][N] (0).  M,N % 128 == 0, K % 16 == 0. */
int gpx_gemm(gpx_handle h, int a_kmajor, int b_kmajor, int64_t M, int64_t N, int64_t K, double alpha,
             const double* A, int64_t lda, const double* B, int64_t ldb, double beta, double* C, int64_t ldc);
/* y = alpha * op(A) x + beta * y for row-major A (m x n); trans=1 uses A^T. */
int gpx_gemv(gpx_handle h, int trans, int64_t m, int64_t n, double alpha, const double* A, int64_t lda,
             const double* x, double beta, double* y);

/* ---- A6/A7: predictive moments and log marginal likelihood --------------------------------
 * out[0] = -.5 y.alpha - sum_i log L_ii - n/2 log(2 pi)   (tune...:141,312; CO2...:148)
 * out[1] = y.alpha, out[2] = sum log L_ii.   n = true (unpadded) size. */
int gpx_lml(gpx_handle h, const double* L, int64_t n, int64_t ldl, const double* y, const double* alpha, double* out3);
/* mu[j] = sum_i Ks[i,j] alpha[i]; var[j] = kss_diag[j] - sum_i V[i,j]^2 (GP_regression.py:143-147). */
int gpx_predict_moments(gpx_handle h, const double* Ks, const double* V, int64_t n, int64_t m, int64_t ld,
                        const double* alpha, const double* kss_diag, double* mu, double* var);

/* ---- A8: fused LML gradient -----------------------------------------------------------------
 * grad[j] = .5 * sum_ik (alpha_i alpha_k - Kinv_ik) dK_ik/dtheta_j with dK recomputed on the fly
 * from X (never materialised); Kinv is read once (lower triangle, symmetric weights).
 * Replaces tune_hyperparms_regression.py:43-57 (N x N x N GEMM + trace).  grad is DEVICE memory. */
int gpx_lml_grad(gpx_handle h, int kind, const double* X, int64_t n, int D, const double* theta_host, int ntheta,
                 const double* Kinv, int64_t ldk, const double* alpha, double* grad);

/* ---- fused drivers used by the Python drop-in modules and bench.py -------------------------
 * gp_fit: K = cov(X,X;theta) + s I  ->  L (in A, padded np x np), dinv, alpha = K^-1 y (np doubles, y is
 * padded with zeros by the callee), out3 = {lml, y.alpha, sum log L_ii}.  Follows
 * tune_hyperparms_regression.py:306-312 / CO2_example.py:142-148. */
int gpx_gp_fit(gpx_handle h, int kind, const double* X, int64_t n, int D, const double* theta_host, int ntheta,
               double s, const double* y, double* A, int64_t np_, int64_t lda, double* dinv,
               double* alpha, double* out3);
/* gp_fit + K^-1 (trtri + lauum into Kinv) + fused gradient; `A` ends up holding L^-1.
 * Follows tune_hyperparms_regression.py:123-145.  Kinv: np x np doubles (ld = lda). */
int gpx_gp_fit_grad(gpx_handle h, int kind, const double* X, int64_t n, int D, const double* theta_host, int ntheta,
                    double s, const double* y, double* A, int64_t np_, int64_t lda, double* dinv,
                    double* Kinv, double* alpha, double* out3, double* grad);

/* ---- N1: device-resident hyper-parameter optimiser -------------------------------------------------------------
 * Gradient ascent on the log marginal likelihood over the hyper-parameters j with mask[j] != 0: per iteration
 * K(theta) + sI -> Cholesky -> alpha -> LML -> K^-1 -> dLML/dtheta -> theta_j += step * grad_j, until
 * |LML - LML_prev| <= tol (LML_prev starts at 0; the step of the converging iteration is still taken) or max_iter.
 * Replaces the loop of tune_hyperparms_regression.py:121-153 (mask = {0, 1}: only l moves, as shipped) and extends it to
 * sigma and to the 11 hyper-parameters of CO2_example.py.  theta, history and convergence state live on the device;
 * with use_graph != 0 one iteration is ONE CUDA-graph launch (captured on the second iteration) + a 40-byte read-back.
 * X, y: device; ws: gpx_gp_ascent_ws_elems(n) doubles (device); theta_io / mask / theta_used_out / out4 / history: HOST.
 * out4 = {iterations, LML of the last iteration, its |LML - LML_prev|, converged}; theta_io returns theta after the last
 * step, theta_used_out the theta the last iteration was evaluated at.  Returns > 0 if a K(theta) is not positive definite. */
int64_t gpx_gp_ascent_ws_elems(int64_t n);
int gpx_gp_ascent(gpx_handle h, int kind, const double* X, int64_t n, int D, double* theta_io, int ntheta, const int* mask,
                  double s, const double* y, double step, double tol, int max_iter, int use_graph, double* ws,
                  double* theta_used_out, double* out4, double* history);

/* ---- A10/A11: Laplace building blocks --------------------------------------------------------*/
/* binary (GP_binary_classification.py:48-83,104-105): mode 0 = reference-faithful gradient
 * t - sigmoid(y f), mode 1 = textbook t - sigmoid(f); w = sigmoid(f)(1-sigmoid(f)); sw = sqrt(w). */
int gpx_logistic_terms(gpx_handle h, int mode, int64_t n, const double* y, const double* f,
                       double* grad, double* w, double* sw);
/* B = I + diag(sw) K diag(sw)  (GP_binary...:107, GP_multi...:93), np x np padded, identity padding. */
int gpx_build_B(gpx_handle h, const double* K, const double* sw, int64_t n, int64_t np_, int64_t ld, double* B);
/* softmax over classes: f, pi are class-major (C, stride) (GP_multi...:26-63). */
int gpx_softmax_classes(gpx_handle h, int C, int64_t n, int64_t stride, const double* f, double* pi);
/* Reference-faithful multiclass pieces (GP_multi...:150-157): the reference's pi_matrix is POINT-major
 * (row a = i*C + c) while D = diag(pi_vector) is CLASS-major (index c*stride + i, stride = literal 60).
 * out = Kinv + c_diag*I + diag(pi_vec) - Pi Pi^T on the padded np x np buffer (identity padding). */
int gpx_multi_ref_hessian(gpx_handle h, int C, int64_t n, int64_t stride, const double* Kinv, int64_t ld,
                          const double* pi_vec, double c_diag, int64_t np_, double* out);
/* out = (D - Pi Pi^T) f with the same conventions (GP_multi...:157). */
int gpx_multi_ref_wf(gpx_handle h, int C, int64_t n, int64_t stride, const double* pi_vec, const double* f, double* out);
/* Textbook Alg. 3.3 (class-major Pi): b = (D - Pi Pi^T) f + y - pi  (GP_multi...:113 with b = W f + y - pi). */
int gpx_multi_b(gpx_handle h, int C, int64_t n, const double* pi, const double* f, const double* y, double* b);
/* E(lower tiles) (+)= diag(sd) X diag(sd)   (E_c of GP_multi...:95 and its sum :101). */
int gpx_scale_sym_acc(gpx_handle h, const double* X, const double* sd, int64_t n, int64_t np_, int64_t ld,
                      int accumulate, double* E);
/* M[i,:] *= s[i] (W^1/2 k_* of GP_binary...:151). */
int gpx_scale_rows(gpx_handle h, int64_t rows, int64_t cols, int64_t ld, const double* s, double* M);
/* copy the strictly-lower triangle of the n x n matrix A into its upper triangle. */
int gpx_symmetrize(gpx_handle h, int64_t n, double* A, int64_t ld);
/* y = S x for a symmetric S stored in its lower triangle. */
int gpx_symv_lower(gpx_handle h, int64_t n, const double* S, int64_t ld, const double* x, double* y);
/* small vector algebra on device (n doubles): see gpx_vec_op codes in vec.cu */
int gpx_vec_op(gpx_handle h, int op, int64_t n, double a, const double* x, const double* y, const double* z, double* out);
/* dst[i*dst_stride] = src[i*src_stride], i < n (e.g. diagonal extraction with src_stride = ld+1) */
int gpx_copy_strided(gpx_handle h, int64_t n, const double* src, int64_t src_stride, double* dst, int64_t dst_stride);
/* out[0] = sum_i x_i*y_i (y may equal x) -- device scalar */
int gpx_dot(gpx_handle h, int64_t n, const double* x, const double* y, double* out);

/* ---- A10/A11: device-resident Laplace iterations (one call = build B -> factor -> solves -> f update -> error; caller
 * workspace, no allocation, no host synchronisation inside a step) -------------------------------------------------------*/
int64_t gpx_laplace_binary_ws_elems(int64_t np_);
/* One textbook Newton iteration of GP_binary_classification.py:104-111 (W and gradient at the current f; R&W Alg. 3.1).
 * K: np x np full symmetric covariance; y, f, f_new: np doubles (zero padded), f_new != f; B (np x np) receives
 * chol(I + W^1/2 K W^1/2), dinv its leaf inverses; ws: gpx_laplace_binary_ws_elems(np) doubles, on return ws[0..np) =
 * gradient, ws[np..2np) = W, ws[2np..3np) = W^1/2 of this iteration; err_dev[0] = |f_new - f|_2 (device).  A failed
 * pivot is reported by gpx_potrf_info. */
int gpx_laplace_binary_step(gpx_handle h, const double* K, int64_t n, int64_t np_, int64_t ld, const double* y,
                            const double* f, double* B, double* dinv, double* ws, double* f_new, double* err_dev);
/* The whole as-shipped loop of GP_binary_classification.py:86-133: W and the gradient are evaluated at f_prior (never at
 * f), B is factored once, inv(L) is formed explicitly (:108) and applied as two mat-vecs, iteration stops at
 * |f_new - f| <= tol.  Linv (np x np) receives inv(L); g_out / w_out / sw_out / f_out: np doubles (device);
 * errors_host[max_iter] and *iters_out are HOST memory.  Returns > 0 when B is not positive definite. */
int gpx_laplace_binary_ref_fit(gpx_handle h, const double* K, int64_t n, int64_t np_, int64_t ld, const double* y,
                               const double* f_prior, double tol, int max_iter, double* B, double* dinv, double* Linv,
                               double* ws, double* f_out, double* g_out, double* w_out, double* sw_out,
                               double* errors_host, int* iters_out);
int64_t gpx_laplace_multi_ws_elems(int64_t np_, int C, int nlanes);
/* One textbook multiclass iteration (GP_multi_classification.py:66-126 / R&W Alg. 3.3) over the nloc classes
 * classes[0..nloc) (HOST ints) owned by this rank.  K: np x np shared covariance block; y, f, f_new, pi: C x n class-major;
 * Linv_store: nloc x np x np (receives L_c^-1); lanes[nlanes]: HOST array of handles, each bound to its own stream, through
 * which the independent per-class factorisations (:88-101) are issued; sums over classes (sum_c E_c, R^T c, f) are
 * all-reduced over h's NCCL communicator when it has one.  err_dev[0] = |f_new - f|_2.  No host synchronisation. */
int gpx_laplace_multi_step(gpx_handle h, const gpx_handle* lanes, int nlanes, const double* K, int64_t n, int64_t np_,
                           int64_t ld, int C, const int* classes, int nloc, const double* y, const double* f, double* ws,
                           double* Linv_store, double* f_new, double* pi, double* err_dev);

/* ---- measurement helpers --------------------------------------------------------------------
 * register-resident DMMA.8x8x4 / DFMA issue-rate microbenchmarks -> measured FP64 peaks (TFLOP/s). */
int gpx_bench_fp64_peak(gpx_handle h, int use_dmma, int iters, double* tflops_out, double* ms_out);

/* CUDA-event instrumentation (off by default).  After gpx_timing_enable(h,1), every DMMA GEMM launch and
 * every phase of the fused drivers is bracketed by events; gpx_timing_collect synchronises and returns
 * out[0] = GEMM kernel ms, out[1] = #GEMM launches, out[2] = flops executed by them, out[3+p] = ms in
 * phase p (0 cov build, 1 potrf, 2 solves+LML, 3 trtri, 4 lauum, 5 gradient).  nout >= 11. */
int gpx_timing_enable(gpx_handle h, int on);
int gpx_timing_collect(gpx_handle h, double* out, int nout);
/* per-launch CSV (phase, M, N, K, flops executed, ms, start offset) of the instrumented region; after gpx_timing_collect */
int gpx_timing_dump(gpx_handle h, const char* path);

/* debug hook: cycles spent by the last potrf leaf kernel in {load, factor, inverse, store} (clock64) */
int gpx_debug_leaf_cycles(long long* out4);

/* ---- multi-GPU (one process per GPU; NCCL communicator owned by the handle) -----------------
 * Layout: 1-D block-cyclic block columns of width nb (a P x 1 grid of the 2-D block-cyclic scheme), panels
 * broadcast with NCCL and kept in a replicated factor, trailing updates grouped to K = 1024; K^-1 by two local
 * prefix-structured TRSMs (no all-gather); see nccl_mg.cu.  NCCL is dlopen'ed at run time. */
int     gpx_nccl_load(const char* libnccl_path);                        /* optional explicit path */
int     gpx_nccl_unique_id(void* id128);                                /* rank 0: fills 128 bytes */
int     gpx_nccl_init(gpx_handle h, const void* id128, int rank, int world);
int     gpx_mg_set_group_k(int k);                                      /* K of the grouped trailing update (default 1024) */
int     gpx_mg_set_layout(int snake);                                   /* 1 (default): boustrophedon block->rank map; 0: plain block-cyclic */
int     gpx_mg_block_owner(int64_t j, int world);                       /* rank that owns global block column j */
int64_t gpx_mg_block_global(int64_t q, int world, int rank);            /* global block of local block q */
int64_t gpx_mg_blocks_below(int64_t j, int world, int rank);            /* # of the rank's blocks with global index < j */
int64_t gpx_mg_padded_dim(int64_t n, int nb, int world);                /* n rounded up to nb*world (2*nb*world with the default map) */
int64_t gpx_mg_workspace_elems(int64_t n, int nb, int world);           /* doubles of device workspace per rank */
/* Element offsets of the pieces inside the workspace: out6 = {Aloc, Kloc (local block columns of K^-1 after a fit with
 * gradient; rows on/below each block's diagonal block are valid), Lfull (replicated factor: npad x npad row-major, clean lower
 * triangle), dinv (its leaf inverses), npad, wloc}. */
int gpx_mg_workspace_layout(int64_t n, int nb, int world, int64_t* out6);
/* This rank's part of the distributed fit + LML (+ gradient when with_grad): tune...:123-145 on P GPUs.
 * X (n x D), y (n), alpha (npad), out3 (3), grad (ntheta) are device pointers; every rank gets the same
 * alpha / out3 / grad.  Returns >0 (first bad pivot) on every rank if the matrix is not positive definite.  After the
 * call the workspace holds the replicated factor (with_grad = 0 or 1) and, with_grad = 1, this rank's block columns of K^-1. */
int gpx_mg_fit_grad(gpx_handle h, int kind, const double* X, int64_t n, int D, const double* theta_host, int ntheta,
                    double s, const double* y, int nb, double* ws, double* alpha, double* out3, double* grad, int with_grad);
/* The distributed (block-cyclic, NCCL panel broadcast) Cholesky on its own: factors
 *   M = diag(scale) k(X, X; theta) diag(scale) + diag_add I      (scale may be NULL)
 * and leaves the replicated factor in the workspace.  scale = W^1/2, diag_add = 1 is the Laplace matrix
 * B = I + W^1/2 K W^1/2 of GP_binary_classification.py:107. */
int gpx_mg_factor(gpx_handle h, int kind, const double* X, int64_t n, int D, const double* theta_host, int ntheta,
                  double diag_add, const double* scale, int nb, double* ws);
/* x <- (L L^T)^-1 x with the replicated factor in the workspace (x: npad doubles, zero padded); no communication. */
int gpx_mg_potrs_vec(gpx_handle h, int64_t n, int nb, double* ws, double* x);
/* One textbook Newton iteration of the binary Laplace approximation (GP_binary_classification.py:104-111) with B factored
 * by the distributed Cholesky: K (np x np, ld) is the replicated covariance used for the two O(N^2) mat-vecs, X / theta its
 * inputs (B is built block-cyclically from them).  y, f, f_new: npad doubles; vws: 8 npad doubles (vws[0..npad) = gradient,
 * [npad..2 npad) = W, [2 npad..3 npad) = W^1/2 on return); err_dev[0] = |f_new - f|_2. */
int gpx_mg_laplace_binary_step(gpx_handle h, int kind, const double* X, int64_t n, int D, const double* theta_host,
                               int ntheta, const double* K, int64_t ld, const double* y, const double* f, int nb,
                               double* ws, double* vws, double* f_new, double* err_dev);
/* Test helper: the same per-rank routines driven for P virtual ranks on ONE GPU (collectives become copies);
 * ws_all = P * gpx_mg_workspace_elems(n, nb, P) doubles. */
int gpx_mg_emulate_fit_grad(gpx_handle h, int P, int kind, const double* X, int64_t n, int D, const double* theta_host,
                            int ntheta, double s, const double* y, int nb, double* ws_all, double* alpha, double* out3,
                            double* grad);

/* ---- host-buffer drop-in calls (HOST pointers; copies inside; synchronous) ------------------*/
/* LML (+ optional gradient wrt all theta when grad_host != NULL) of y ~ GP(0, cov + s I). */
int gpx_host_lml(gpx_handle h, int kind, const double* X, int64_t n, int D, const double* theta, int ntheta,
                 double s, const double* y, double* lml_out, double* grad_host);
/* Whole posterior of GP_regression.py:109-156 (== tune...:67-101, CO2...:182-214) for SMALL problems in ONE kernel
 * launch: requires N <= gpx_small_max() training points, and n <= gpx_small_max() test points when sampling (nf > 0;
 * without sampling any n: one thread block per gpx_small_max() test points).  HOST pointers: X[N,D], y[N],
 * Xs[n,D], Z[n,nf] (the caller's standard normals, drawn on the host so the NumPy RNG stream stays the reference's;
 * nf = 0 skips sampling) -> mu[n], var[n] (NOT square-rooted; may be negative exactly where the reference yields NaN),
 * fpost[n,nf] = mu + chol(K** + jitter I - V^T V) Z, lml.  Returns > 0 when K + s I (GP_regression.py:138) or the
 * posterior covariance (:154) is not positive definite (index of the failing pivot). */
int gpx_gp_small_posterior_host(gpx_handle h, int kind, const double* X, int64_t N, int D, const double* y,
                                const double* Xs, int64_t n, const double* theta, int ntheta, double s, double jitter,
                                const double* Z, int nf, double* mu, double* var, double* fpost, double* lml);
int gpx_small_max(void);
/* The same posterior in two steps, for hosts that must draw their normals AFTER the linear algebra succeeded (the
 * reference raises LinAlgError at GP_regression.py:138/:154 before it consumes the RNG at :155):
 * gpx_gp_small_fit_host -> mu, var, lml and keeps chol(K** + jitter I - V^T V) on the device (n <= gpx_small_max());
 * gpx_gp_small_sample_host -> fpost[n,nf] = mu + L_ Z for the factor of the last successful fit on this handle. */
int gpx_gp_small_fit_host(gpx_handle h, int kind, const double* X, int64_t N, int D, const double* y, const double* Xs,
                          int64_t n, const double* theta, int ntheta, double s, double jitter, double* mu, double* var,
                          double* lml);
int gpx_gp_small_sample_host(gpx_handle h, int64_t n, int nf, const double* Z, double* fpost);
/* Prior draws (GP_regression.py:71-92) for n <= gpx_small_max(): factor k(Xs,Xs) + s I in one launch and keep it on the
 * device; gpx_gp_small_sample_host(h, n, nf, Z, out) then returns L Z (the caller adds its prior mean).  Returns > 0
 * when the matrix is not positive definite (:90). */
int gpx_gp_small_prior_factor_host(gpx_handle h, int kind, const double* Xs, int64_t n, int D, const double* theta,
                                   int ntheta, double s);
/* LML (tune...:292-313, CO2...:131-149) and, when grad != NULL, dLML/dtheta for every hyper-parameter
 * (tune...:31-64,144 generalised per SURVEY Appendix C) for N <= gpx_small_max() in one kernel launch. */
int gpx_gp_small_lml_grad_host(gpx_handle h, int kind, const double* X, int64_t N, int D, const double* y,
                               const double* theta, int ntheta, double s, double* lml, double* grad);
/* The reference's whole gradient-ascent loop on the SE length-scale (tune_hyperparms_regression.py:121-153) inside ONE
 * kernel launch, all state in shared memory: l <- l + step * dLML/dl until |LML - LML_old| <= tol (LML_old starts at 0;
 * the step of the converging iteration is still taken) or max_iter iterations.  out6 = [iterations, l after the last
 * step, l the last iteration was evaluated at, its LML, its |LML - LML_old|, converged flag]. */
int gpx_gp_small_ascent_host(gpx_handle h, const double* X, int64_t N, int D, const double* y, double sigma, double l0,
                             double s, double step, double tol, int max_iter, double* out6);

#ifdef __cplusplus
}
#endif
#endif /* GPX_H */
