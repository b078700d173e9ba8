// Device-resident Laplace iterations (SURVEY.md 8a rows A10/A11, 8b "single-call Laplace steps"): one C-ABI call runs
// build-B -> potrf -> solves -> f update -> error on caller-provided workspace, with no allocation and no host
// synchronisation inside a step.  The O(N^3) work is potrf.cu / gemm*.cu; the element-wise pieces are laplace.cu / vec.cu.
//   gpx_laplace_binary_step     : one textbook Newton iteration (R&W Alg. 3.1; GP_binary_classification.py:104-111 with
//                                 W and the gradient evaluated at the current f)
//   gpx_laplace_binary_ref_fit  : the whole as-shipped loop (GP_binary_classification.py:86-133: W, gradient frozen at
//                                 f_prior, B factored once, inv(L) formed explicitly), one 16-byte read-back per iteration
//   gpx_laplace_multi_step      : one textbook multiclass iteration (R&W Alg. 3.3; GP_multi_classification.py:66-126) over
//                                 the classes this rank owns, per-class factorisations issued through several handles
//                                 (streams), sums over classes completed by NCCL when the handle has a communicator
#include <vector>
#include "common.cuh"

int gpx_nccl_allreduce_sum(gpx_ctx* h, double* buf, size_t count);   // nccl_mg.cu (no-op when world == 1)

namespace {

__global__ void sqrt_scalar_kernel(double* x) { x[0] = sqrt(x[0]); }

// err[0] = |a - b|_2 over n entries (two launches + a scalar sqrt; deterministic)
int diff_norm(gpx_ctx* h, int64_t n, const double* a, const double* b, double* tmp, double* err) {
    GPX_TRY(gpx_vec_op(h, 6, n, 0.0, a, b, nullptr, tmp));
    GPX_TRY(gpx_dot(h, n, tmp, tmp, err));
    sqrt_scalar_kernel<<<1, 1, 0, h->stream>>>(err);
    GPX_CHECK_LAUNCH(h);
    return 0;
}

struct BinWs {
    double *g, *w, *sw, *b, *t, *a, *u, *d;   // np each
    double *Dbig, *work, *tmp;                // block inverses of the current factor
    double* tri;                              // trtri workspace (np^2/4) of the reference-faithful loop
    int bs;
};

BinWs carve_bin(double* ws, int64_t np_) {
    BinWs v;
    double* p = ws;
    v.g = p; p += np_; v.w = p; p += np_; v.sw = p; p += np_; v.b = p; p += np_;
    v.t = p; p += np_; v.a = p; p += np_; v.u = p; p += np_; v.d = p; p += np_;
    v.bs = gpx_block_size_for(np_);
    v.Dbig = p; p += (size_t)np_ * v.bs;
    v.work = p; p += (size_t)np_ * v.bs / 4 + GPX_T;
    v.tmp = p; p += v.bs + 64;
    v.tri = p;
    return v;
}

// t <- (L L^T)^-1 t with the short-chain solves when the matrix is large enough
int potrs_vec(gpx_ctx* h, const double* L, int64_t np_, int64_t ld, const double* dinv, const BinWs& v, double* x) {
    if (v.bs > GPX_T && np_ >= 2 * v.bs) {
        GPX_TRY(gpx_block_inverses(h, L, np_, ld, dinv, v.bs, v.Dbig, v.work));
        GPX_TRY(gpx_trsv_big(h, L, np_, ld, v.Dbig, v.bs, 0, x, v.tmp));
        return gpx_trsv_big(h, L, np_, ld, v.Dbig, v.bs, 1, x, v.tmp);
    }
    GPX_TRY(gpx_trsv(h, L, np_, ld, dinv, 0, x));
    return gpx_trsv(h, L, np_, ld, dinv, 1, x);
}

}  // namespace

extern "C" int64_t gpx_laplace_binary_ws_elems(int64_t np_) {
    const int64_t bs = gpx_block_size_for(np_);
    return 8 * np_ + np_ * bs + np_ * bs / 4 + GPX_T + bs + 64 + np_ * np_ / 4 + GPX_T;
}

// One Newton iteration.  K: np x np (full symmetric covariance in [:n,:n]); y, f: np (zero padded); B receives the factor
// of I + W^1/2 K W^1/2, dinv its leaf inverses; f_new (np, != f); err_dev[0] = |f_new - f|_2.  Pivot failures are sticky
// in the handle (gpx_potrf_info).  Vectors g, w, sw of this iteration stay at ws[0 .. 3 np).
extern "C" int gpx_laplace_binary_step(gpx_handle h, const double* K, int64_t n, int64_t np_, int64_t ld, const double* y,
                                       const double* f, double* B, double* dinv, double* ws, double* f_new, double* err_dev) {
    GPX_ENTER(h);
    GPX_REQUIRE(np_ % GPX_T == 0 && np_ >= n && n > 0, 4);
    GPX_REQUIRE(f_new != f, 11);
    BinWs v = carve_bin(ws, np_);
    GPX_TRY(gpx_logistic_terms(h, 1, n, y, f, v.g, v.w, v.sw));                    // :104-105 at the current f
    GPX_TRY(gpx_build_B(h, K, v.sw, n, np_, ld, B));                               // :107
    GPX_TRY(gpx_potrf_async(h, B, np_, ld, dinv));
    GPX_TRY(gpx_vec_op(h, 5, n, 0.0, v.w, f, v.g, v.b));                           // b = W f + grad          (:109)
    GPX_TRY(gpx_gemv(h, 0, n, n, 1.0, K, ld, v.b, 0.0, v.t));                      // K b
    GPX_CUDA(cudaMemsetAsync(v.t + n, 0, (np_ - n) * sizeof(double), h->stream));
    GPX_TRY(gpx_vec_op(h, 2, n, 0.0, v.sw, v.t, nullptr, v.t));                    // W^1/2 K b
    GPX_TRY(potrs_vec(h, B, np_, ld, dinv, v, v.t));                               // B^-1 .
    GPX_TRY(gpx_vec_op(h, 3, n, 0.0, v.b, v.sw, v.t, v.a));                        // a = b - W^1/2 B^-1 W^1/2 K b (:110)
    GPX_TRY(gpx_gemv(h, 0, n, n, 1.0, K, ld, v.a, 0.0, f_new));                    // f = K a                 (:111)
    return diff_norm(h, n, f_new, f, v.d, err_dev);
}

// The as-shipped loop.  Linv (np x np) receives inv(L) (the reference forms and returns it, :108,:133); g/w/sw out: np
// each; errors_host[max_iter]; *iters_out = iterations run.  f_out (np) = last iterate.
extern "C" int gpx_laplace_binary_ref_fit(gpx_handle h, const double* K, int64_t n, int64_t np_, int64_t ld, const double* y,
                                          const double* f_prior, double tol, int max_iter, double* B, double* dinv,
                                          double* Linv, double* ws, double* f_out, double* g_out, double* w_out, double* sw_out,
                                          double* errors_host, int* iters_out) {
    GPX_ENTER(h);
    GPX_REQUIRE(np_ % GPX_T == 0 && np_ >= n && n > 0, 4);
    GPX_REQUIRE(Linv != nullptr && errors_host != nullptr && iters_out != nullptr, 12);
    BinWs v = carve_bin(ws, np_);
    cudaStream_t S = h->stream;
    GPX_CUDA(cudaMemsetAsync(g_out, 0, np_ * sizeof(double), S));
    GPX_CUDA(cudaMemsetAsync(w_out, 0, np_ * sizeof(double), S));
    GPX_CUDA(cudaMemsetAsync(sw_out, 0, np_ * sizeof(double), S));
    GPX_TRY(gpx_logistic_terms(h, 0, n, y, f_prior, g_out, w_out, sw_out));        // :104-105, evaluated at f_prior
    GPX_TRY(gpx_build_B(h, K, sw_out, n, np_, ld, B));
    int info = gpx_potrf(h, B, np_, ld, dinv);                                      // :107
    if (info != 0) return info;
    GPX_CUDA(cudaMemcpy2DAsync(Linv, ld * sizeof(double), B, ld * sizeof(double), np_ * sizeof(double), np_,
                               cudaMemcpyDeviceToDevice, S));
    GPX_TRY(gpx_trtri(h, Linv, np_, ld, dinv, v.tri));                              // :108
    double* f = f_out;
    double* fn = v.u;                                                               // ping-pong: f_out <-> ws.u
    GPX_CUDA(cudaMemsetAsync(f, 0, np_ * sizeof(double), S));
    GPX_CUDA(cudaMemsetAsync(fn, 0, np_ * sizeof(double), S));
    GPX_CUDA(cudaMemsetAsync(v.t, 0, np_ * sizeof(double), S));
    double* err_dev = v.tmp;
    int it = 0;
    for (; it < max_iter;) {
        GPX_TRY(gpx_vec_op(h, 5, n, 0.0, w_out, f, g_out, v.b));                   // b = W f + grad
        GPX_TRY(gpx_gemv(h, 0, n, n, 1.0, K, ld, v.b, 0.0, v.t));
        GPX_TRY(gpx_vec_op(h, 2, n, 0.0, sw_out, v.t, nullptr, v.t));
        GPX_TRY(gpx_gemv(h, 0, n, n, 1.0, Linv, ld, v.t, 0.0, v.a));               // L_inv .
        GPX_TRY(gpx_gemv(h, 1, n, n, 1.0, Linv, ld, v.a, 0.0, v.t));               // L_inv^T .
        GPX_TRY(gpx_vec_op(h, 3, n, 0.0, v.b, sw_out, v.t, v.a));
        GPX_TRY(gpx_gemv(h, 0, n, n, 1.0, K, ld, v.a, 0.0, fn));
        GPX_TRY(diff_norm(h, n, fn, f, v.d, err_dev));
        double e = 0.0;
        GPX_CUDA(cudaMemcpyAsync(&e, err_dev, sizeof(double), cudaMemcpyDeviceToHost, S));
        GPX_CUDA(cudaStreamSynchronize(S));
        errors_host[it++] = e;
        double* sw_ = f; f = fn; fn = sw_;
        if (e <= tol) break;
    }
    if (f != f_out) GPX_CUDA(cudaMemcpyAsync(f_out, f, np_ * sizeof(double), cudaMemcpyDeviceToDevice, S));
    *iters_out = it;
    return 0;
}

// ------------------------------------------------------------------------------------------------ multiclass (Alg. 3.3)
extern "C" int64_t gpx_laplace_multi_ws_elems(int64_t np_, int C, int nlanes) {
    // per lane: Ec + Esum (np^2 each) + trtri work (np^2 / 4 + slack); shared: M block inverses + vectors
    const int64_t bs = gpx_block_size_for(np_);
    return (int64_t)nlanes * (2 * np_ * np_ + np_ * np_ / 4 + GPX_T) + np_ * bs + np_ * bs / 4 + GPX_T + bs +
           (int64_t)(4 * C + 8) * np_ + np_ * GPX_T + 64;
}

// One iteration over the classes `classes[0..nloc)` this rank owns.  K: np x np shared covariance block (the reference's
// block_diag(K_sub x C), :233-238); y, f, f_new, pi: C x n class-major (contiguous); Linv_store: nloc x np x np (L_c^-1
// kept for the E_c mat-vecs); lanes[nlanes]: handles (each bound to its own stream by the caller) through which the
// per-class factorisations are issued; err_dev[0] = |f_new - f|_2.  Sums over classes are all-reduced over h's NCCL
// communicator when h->world > 1.  No host synchronisation; pivot failures are sticky in the lane handles.
extern "C" int gpx_laplace_multi_step(gpx_handle h, const gpx_handle* lanes, int nlanes, const double* K, int64_t n,
                                      int64_t np_, int64_t ld, int C, const int* classes, int nloc, const double* y,
                                      const double* f, double* ws, double* Linv_store, double* f_new, double* pi,
                                      double* err_dev) {
    GPX_ENTER(h);
    GPX_REQUIRE(nlanes >= 1 && lanes != nullptr, 3);
    GPX_REQUIRE(np_ % GPX_T == 0 && np_ >= n && n > 0, 6);
    GPX_REQUIRE(C >= 1 && nloc >= 0 && nloc <= C, 8);
    cudaStream_t S = h->stream;
    const size_t sq = (size_t)np_ * np_;
    const int bs = gpx_block_size_for(np_);
    // ---- carve the workspace
    double* p = ws;
    std::vector<double*> Ec(nlanes), Es(nlanes), Wk(nlanes);
    for (int k = 0; k < nlanes; ++k) {
        Ec[k] = p; p += sq;
        Es[k] = p; p += sq;
        Wk[k] = p; p += sq / 4 + GPX_T;
    }
    BinWs mv;
    mv.bs = bs;
    mv.Dbig = p; p += (size_t)np_ * bs;
    mv.work = p; p += (size_t)np_ * bs / 4 + GPX_T;
    mv.tmp = p; p += bs + 64;
    double* sd = p; p += (size_t)C * np_;        // D_c^1/2 per class (padded rows)
    double* b = p; p += (size_t)C * np_;         // class c at b + c*n (contiguous C x n inside)
    double* cv = p; p += (size_t)C * np_;        // c_c = E_c K b_c, class c at cv + c*np
    double* a = p; p += (size_t)C * np_;
    double* rsum = p; p += np_;
    double* t1 = p; p += np_;
    double* t2 = p; p += np_;
    double* t3 = p; p += np_;
    double* dinvM = p; p += (size_t)np_ * GPX_T;  // leaf inverses of M
    GPX_TRY(gpx_softmax_classes(h, C, n, n, f, pi));                                          // :51-58
    GPX_CUDA(cudaMemsetAsync(sd, 0, (size_t)C * np_ * sizeof(double), S));
    GPX_CUDA(cudaEventRecord(h->ev_a, S));
    // ---- per-class factorisations on the lanes (:88-101)
    std::vector<int> used(nlanes, 0), first_on_lane(nlanes, 1);
    auto lanes_body = [&]() -> int {
        for (int i = 0; i < nloc; ++i) {
            const int c = classes[i], k = i % nlanes;
            gpx_ctx* e = lanes[k];
            if (!used[k]) GPX_CUDA(cudaStreamWaitEvent(e->stream, h->ev_a, 0));
            used[k] = 1;                                                                           // from here on the lane must be joined
            double* sdc = sd + (size_t)c * np_;
            double* Lc = Linv_store + (size_t)i * sq;
            GPX_TRY(gpx_vec_op(e, 10, n, 0.0, pi + (size_t)c * n, nullptr, nullptr, sdc));          // D_c^1/2
            GPX_TRY(gpx_build_B(e, K, sdc, n, np_, ld, Lc));                                       // :92
            double* dv = Ec[k];                                                                    // leaf inverses: head of Ec (overwritten by lauum later)
            GPX_TRY(gpx_potrf_async(e, Lc, np_, np_, dv));                                         // :93  L_c
            GPX_TRY(gpx_trtri(e, Lc, np_, np_, dv, Wk[k]));                                        // :94  L_c^-1
            GPX_TRY(gpx_lauum(e, Lc, np_, np_, Ec[k], np_));                                       // B_c^-1 (lower tiles)
            GPX_TRY(gpx_scale_sym_acc(e, Ec[k], sdc, n, np_, np_, first_on_lane[k] ? 0 : 1, Es[k])); // :95,:101
            first_on_lane[k] = 0;
        }
        return 0;
    };
    const int lanes_rc = lanes_body();
    // join every lane that was started back into S -- also on error, so that the caller never recycles the workspace while a
    // lane is still writing it
    for (int k = 0; k < nlanes; ++k)
        if (used[k]) {
            gpx_ctx* e = lanes[k];
            if (cudaEventRecord(e->ev_b, e->stream) == cudaSuccess) cudaStreamWaitEvent(S, e->ev_b, 0);
        }
    GPX_TRY(lanes_rc);
    double* Esum = Es[0];
    if (!used[0]) GPX_CUDA(cudaMemsetAsync(Esum, 0, sq * sizeof(double), S));                  // this rank owns no class
    for (int k = 1; k < nlanes; ++k)
        if (used[k]) GPX_TRY(gpx_vec_op(h, 1, (int64_t)sq, 1.0, Esum, Es[k], nullptr, Esum));
    GPX_TRY(gpx_nccl_allreduce_sum(h, Esum, sq));
    if (np_ > n) {  // the padding block must be the identity again (it may hold a sum over lanes / ranks)
        GPX_TRY(gpx_vec_op(h, 4, np_ - n, 1.0, nullptr, nullptr, nullptr, t1));
        GPX_TRY(gpx_copy_strided(h, np_ - n, t1, 1, Esum + (size_t)n * np_ + n, np_ + 1));
    }
    GPX_TRY(gpx_potrf_async(h, Esum, np_, np_, dinvM));                                        // :107  M = chol(sum E_c)
    GPX_TRY(gpx_multi_b(h, C, n, pi, f, y, b));                                                // :113
    auto apply_E = [&](int i, const double* x, double* out) -> int {                           // out = E_c x (np entries)
        const int c = classes[i];
        const double* sdc = sd + (size_t)c * np_;
        const double* Lc = Linv_store + (size_t)i * sq;
        GPX_TRY(gpx_vec_op(h, 2, n, 0.0, sdc, x, nullptr, t1));
        GPX_TRY(gpx_gemv(h, 0, n, n, 1.0, Lc, np_, t1, 0.0, t2));
        GPX_TRY(gpx_gemv(h, 1, n, n, 1.0, Lc, np_, t2, 0.0, t3));
        return gpx_vec_op(h, 2, n, 0.0, sdc, t3, nullptr, out);
    };
    GPX_CUDA(cudaMemsetAsync(rsum, 0, np_ * sizeof(double), S));
    GPX_CUDA(cudaMemsetAsync(cv, 0, (size_t)C * np_ * sizeof(double), S));
    for (int i = 0; i < nloc; ++i) {
        const int c = classes[i];
        double* kb = a + (size_t)c * np_;                                                      // scratch until a_c is formed
        GPX_TRY(gpx_gemv(h, 0, n, n, 1.0, K, ld, b + (size_t)c * n, 0.0, kb));
        GPX_TRY(apply_E(i, kb, cv + (size_t)c * np_));                                         // :114  c = E K b
        GPX_TRY(gpx_vec_op(h, 1, n, 1.0, rsum, cv + (size_t)c * np_, nullptr, rsum));          // R^T c
    }
    GPX_TRY(gpx_nccl_allreduce_sum(h, rsum, np_));
    GPX_TRY(potrs_vec(h, Esum, np_, np_, dinvM, mv, rsum));                                    // M^-T M^-1 R^T c  (:115)
    GPX_CUDA(cudaMemsetAsync(f_new, 0, (size_t)C * n * sizeof(double), S));
    for (int i = 0; i < nloc; ++i) {
        const int c = classes[i];
        double* ac = a + (size_t)c * np_;
        GPX_TRY(apply_E(i, rsum, ac));                                                         // E_c R M^-T M^-1 R^T c
        GPX_TRY(gpx_vec_op(h, 6, n, 0.0, b + (size_t)c * n, cv + (size_t)c * np_, nullptr, t1));
        GPX_TRY(gpx_vec_op(h, 1, n, 1.0, t1, ac, nullptr, ac));                                // :116
        GPX_TRY(gpx_gemv(h, 0, n, n, 1.0, K, ld, ac, 0.0, f_new + (size_t)c * n));             // :117
    }
    GPX_TRY(gpx_nccl_allreduce_sum(h, f_new, (size_t)C * n));
    return diff_norm(h, (int64_t)C * n, f_new, f, b, err_dev);     // b is dead here: C*n <= C*np scratch
}
