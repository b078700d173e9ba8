// Device-resident hyper-parameter optimiser: gradient ascent on the log marginal likelihood over any
// subset of the hyper-parameters of any covariance family -- tune_hyperparms_regression.py:121-153 (which updates only l;
// its sigma update is commented out at :46-62) generalised to sigma, l and the 11 CO2 hyper-parameters
// (CO2_example.py:330-379 optimises those by Bayesian optimisation only).
//
// One iteration = K(theta) -> Cholesky -> alpha -> LML -> K^-1 -> dLML/dtheta -> theta_j += step * grad_j (masked), with
// theta, the LML history and the convergence state in DEVICE memory: the covariance and gradient kernels read theta from
// there, so the whole iteration is captured ONCE into a CUDA graph (second iteration) and replayed; the host only reads
// 40 bytes per iteration for the convergence decision.  All buffers come from the caller's workspace (no allocation
// inside the loop: the first, eager iteration sizes the handle's scratch before the capture).
#include "common.cuh"

namespace {

struct AscentState {        // device
    double theta[16];       // current hyper-parameters (read by the kernels)
    double theta_used[16];  // hyper-parameters the last iteration was evaluated at
    double lml_prev;        // LML of the previous iteration (starts at 0, tune...:116)
    double lml, err;        // LML / |LML - LML_prev| of the last iteration
    double iters;
};

__global__ void ascent_update_kernel(AscentState* st, const double* __restrict__ out3, const double* __restrict__ grad,
                                     const int* __restrict__ mask, int ntheta, double step) {
    if (threadIdx.x != 0) return;
    const double lml = out3[0];
    for (int j = 0; j < ntheta; ++j) {
        st->theta_used[j] = st->theta[j];
        if (mask[j]) st->theta[j] += step * grad[j];             // tune...:35-36,60-61 (always taken, before the test)
    }
    st->err = sqrt((lml - st->lml_prev) * (lml - st->lml_prev));  // tune...:147
    st->lml_prev = lml;                                           // tune...:148
    st->lml = lml;
    st->iters += 1.0;
}

struct AscentWs {
    double *A, *Kinv, *dinv, *alpha, *ypad, *out, *Dbig, *work, *tmp;
    AscentState* st;
    int* mask;
    int bs;
};

AscentWs carve(double* ws, int64_t np_) {
    AscentWs w;
    double* p = ws;
    w.A = p; p += (size_t)np_ * np_;
    w.Kinv = p; p += (size_t)np_ * np_;
    w.dinv = p; p += (size_t)np_ * GPX_T;
    w.alpha = p; p += np_;
    w.ypad = p; p += np_;
    w.out = p; p += 32;
    w.st = (AscentState*)p; p += 64;
    w.mask = (int*)p; p += 16;
    w.bs = gpx_block_size_for(np_);
    w.Dbig = p; p += (size_t)np_ * w.bs;
    w.work = p; p += (size_t)np_ * w.bs / 4 + GPX_T;
    w.tmp = p;
    return w;
}

// enqueue one iteration on h->stream (no host synchronisation, no allocation once the handle's scratch is warm)
int enqueue_iteration(gpx_ctx* h, int kind, const double* X, int64_t n, int D, const double* theta_host, int ntheta, double s,
                      const double* y, int64_t np_, const AscentWs& w, double step) {
    const double* th_dev = w.st->theta;
    GPX_TRY(gpx_cov_build_block(h, kind, X, n, D, theta_host, ntheta, s, GPX_COV_SAME_X | GPX_COV_LOWER | GPX_COV_SKIP_UPPER, w.A,
                                np_, np_, np_, 0, 0, nullptr, th_dev));                                  // tune...:123,127
    GPX_TRY(gpx_potrf_async(h, w.A, np_, np_, w.dinv));
    GPX_CUDA(cudaMemcpyAsync(w.alpha, w.ypad, np_ * sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
    if (w.bs > GPX_T && np_ >= 2 * w.bs) {
        GPX_TRY(gpx_block_inverses(h, w.A, np_, np_, w.dinv, w.bs, w.Dbig, w.work));
        GPX_TRY(gpx_trsv_big(h, w.A, np_, np_, w.Dbig, w.bs, 0, w.alpha, w.tmp));                         // tune...:128
        GPX_TRY(gpx_trsv_big(h, w.A, np_, np_, w.Dbig, w.bs, 1, w.alpha, w.tmp));                         // tune...:129
    } else {
        GPX_TRY(gpx_trsv(h, w.A, np_, np_, w.dinv, 0, w.alpha));
        GPX_TRY(gpx_trsv(h, w.A, np_, np_, w.dinv, 1, w.alpha));
    }
    GPX_TRY(gpx_lml(h, w.A, n, np_, y, w.alpha, w.out));                                                  // tune...:141
    GPX_TRY(gpx_trtri(h, w.A, np_, np_, w.dinv, w.Kinv));                                                 // tune...:144
    GPX_TRY(gpx_lauum(h, w.A, np_, np_, w.Kinv, np_));
    GPX_TRY(gpx_lml_grad_block(h, kind, X, n, D, theta_host, ntheta, w.Kinv, np_, w.alpha, w.out + 3, np_, np_, 0, 0, th_dev));
    ascent_update_kernel<<<1, 32, 0, h->stream>>>(w.st, w.out, w.out + 3, w.mask, ntheta, step);          // tune...:145-148
    GPX_CHECK_LAUNCH(h);
    return 0;
}

}  // namespace

extern "C" int64_t gpx_gp_ascent_ws_elems(int64_t n) {
    const int64_t np_ = gpx_padded_dim(n);
    const int64_t bs = gpx_block_size_for(np_);
    return 2 * np_ * np_ + np_ * GPX_T + 2 * np_ + 32 + 64 + 16 + np_ * bs + np_ * bs / 4 + GPX_T + bs + 64;
}

// X (n x D), y (n): device.  theta_io (ntheta), mask (ntheta, 0 = keep fixed), out4, history (max_iter or NULL): HOST.
// theta_io returns theta AFTER the last step; theta_used_out (ntheta, HOST) the theta the last iteration was evaluated at
// (the one whose LML is reported: tune...:155-157); out4 = {iterations, its LML, its |LML - LML_prev|, converged}.
// use_graph = 0 enqueues every iteration eagerly (reference point for the CUDA-graph replay).
extern "C" int gpx_gp_ascent(gpx_handle h, int kind, const double* X, int64_t n, int D, double* theta_io, int ntheta,
                             const int* mask, double s, const double* y, double step, double tol, int max_iter, int use_graph,
                             double* ws, double* theta_used_out, double* out4, double* history) {
    GPX_ENTER(h);
    GPX_REQUIRE(n > 0 && ntheta >= 1 && ntheta <= 11, 7);
    GPX_REQUIRE(theta_io && mask && out4 && ws, 6);
    GPX_REQUIRE(max_iter >= 1, 13);
    const int64_t np_ = gpx_padded_dim(n);
    AscentWs w = carve(ws, np_);
    // The loop runs on a private stream: the caller's stream may be the legacy default stream, which cannot be captured.
    cudaStream_t caller = h->stream;
    if (!h->graph_stream) GPX_CUDA(cudaStreamCreateWithFlags(&h->graph_stream, cudaStreamNonBlocking));
    cudaStream_t S = h->graph_stream;
    GPX_CUDA(cudaEventRecord(h->ev_a, caller));           // X, y were produced on the caller's stream
    GPX_CUDA(cudaStreamWaitEvent(S, h->ev_a, 0));
    struct Restore { gpx_ctx* h; cudaStream_t s; ~Restore() { h->stream = s; } } restore{h, caller};
    h->stream = S;
    AscentState init;
    memset(&init, 0, sizeof(init));
    for (int j = 0; j < ntheta; ++j) init.theta[j] = init.theta_used[j] = theta_io[j];
    int mask16[16] = {0};
    for (int j = 0; j < ntheta; ++j) mask16[j] = mask[j];
    GPX_CUDA(cudaMemcpyAsync(w.st, &init, sizeof(init), cudaMemcpyHostToDevice, S));
    GPX_CUDA(cudaMemcpyAsync(w.mask, mask16, sizeof(mask16), cudaMemcpyHostToDevice, S));
    GPX_CUDA(cudaMemsetAsync(w.ypad, 0, np_ * sizeof(double), S));
    GPX_CUDA(cudaMemcpyAsync(w.ypad, y, n * sizeof(double), cudaMemcpyDeviceToDevice, S));
    GPX_CUDA(cudaMemsetAsync(h->d_info, 0, sizeof(int), S));
    GPX_CUDA(cudaStreamSynchronize(S));      // `init` / `mask16` are stack buffers

    cudaGraph_t graph = nullptr;
    cudaGraphExec_t exec = nullptr;
    int rc = 0, it = 0, converged = 0;
    AscentState host;
    memset(&host, 0, sizeof(host));
    const int timing_was = h->timing_on;
    for (; it < max_iter && rc == 0;) {
        if (it == 0 || !use_graph) {
            rc = enqueue_iteration(h, kind, X, n, D, theta_io, ntheta, s, y, np_, w, step);          // eager (warms the scratch)
        } else {
            if (!exec) {
                h->timing_on = 0;                                                                     // no event pairs inside a capture
                cudaError_t eb = cudaStreamBeginCapture(S, cudaStreamCaptureModeRelaxed);
                if (eb != cudaSuccess) {
                    gpx_set_error("gpx_gp_ascent: cudaStreamBeginCapture failed: %s", cudaGetErrorString(eb));
                    h->timing_on = timing_was;
                    rc = GPX_E_CUDA;
                    break;
                }
                rc = enqueue_iteration(h, kind, X, n, D, theta_io, ntheta, s, y, np_, w, step);
                cudaError_t e = cudaStreamEndCapture(S, &graph);
                h->timing_on = timing_was;
                if (rc == 0 && (e != cudaSuccess || graph == nullptr)) {
                    gpx_set_error("gpx_gp_ascent: stream capture failed: %s", cudaGetErrorString(e));
                    rc = GPX_E_CUDA;
                }
                if (rc == 0 && cudaGraphInstantiate(&exec, graph, 0) != cudaSuccess) {
                    gpx_set_error("gpx_gp_ascent: cudaGraphInstantiate failed: %s", cudaGetErrorString(cudaGetLastError()));
                    rc = GPX_E_CUDA;
                }
                if (rc != 0) break;
            }
            if (cudaGraphLaunch(exec, S) != cudaSuccess) {
                gpx_set_error("gpx_gp_ascent: cudaGraphLaunch failed: %s", cudaGetErrorString(cudaGetLastError()));
                rc = GPX_E_CUDA;
                break;
            }
            h->launches += 1;
        }
        if (rc != 0) break;
        int info = 0;
        if (cudaMemcpyAsync(&host, w.st, sizeof(host), cudaMemcpyDeviceToHost, S) != cudaSuccess) { rc = GPX_E_CUDA; break; }
        rc = gpx_read_info(h, &info);                                                                 // synchronises S
        if (rc != 0) break;
        if (info > 0) {
            gpx_set_error("gpx_gp_ascent: K + sI is not positive definite at iteration %d (leading minor %d)", it + 1, info);
            cudaMemsetAsync(h->d_info, 0, sizeof(int), S);
            rc = info;
            break;
        }
        if (history) history[it] = host.lml;
        ++it;
        if (host.err <= tol) { converged = 1; break; }                                                // tune...:149
    }
    if (exec) cudaGraphExecDestroy(exec);
    if (graph) cudaGraphDestroy(graph);
    h->timing_on = timing_was;
    if (rc != 0) return rc;
    for (int j = 0; j < ntheta; ++j) {
        theta_io[j] = host.theta[j];
        if (theta_used_out) theta_used_out[j] = host.theta_used[j];
    }
    out4[0] = it; out4[1] = host.lml; out4[2] = host.err; out4[3] = converged;
    return 0;
}
