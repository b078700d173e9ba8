// Handle management, error reporting, fused fit / fit+gradient drivers, host-buffer entry points and
// the FP64 peak micro-benchmarks of libgpx.
#include <stdarg.h>
#include "common.cuh"

static thread_local char g_err[1024] = "";

void gpx_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

extern "C" const char* gpx_last_error(void) { return g_err; }
extern "C" int gpx_version(void) { return GPX_VERSION; }
extern "C" int64_t gpx_padded_dim(int64_t n) { return ((n + GPX_T - 1) / GPX_T) * GPX_T; }

extern "C" int gpx_create(int device, gpx_handle* out) {
    GPX_REQUIRE(out != nullptr, 2);
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) {
        gpx_set_error("gpx_create: no CUDA device available (%s); libgpx has no CPU fallback", cudaGetErrorString(e));
        return GPX_E_CUDA;
    }
    GPX_REQUIRE(device >= 0 && device < ndev, 1);
    GPX_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    GPX_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10) {
        gpx_set_error("gpx_create: device %d is sm_%d%d; libgpx is built for sm_100a (B200) only", device, prop.major, prop.minor);
        return GPX_E_CUDA;
    }
    gpx_ctx* h = new gpx_ctx();
    memset(h, 0, sizeof(*h));
    h->device = device;
    h->stream = 0;
    h->world = 1;
    GPX_CUDA(cudaMalloc(&h->d_info, sizeof(int)));
    GPX_CUDA(cudaMemset(h->d_info, 0, sizeof(int)));
    GPX_CUDA(cudaMalloc(&h->d_theta, 16 * sizeof(double)));
    {   // communication stream: highest priority so NCCL blocks are scheduled as soon as an SM frees up
        int lo = 0, hi = 0;
        GPX_CUDA(cudaDeviceGetStreamPriorityRange(&lo, &hi));
        GPX_CUDA(cudaStreamCreateWithPriority(&h->aux_stream, cudaStreamNonBlocking, hi));
        GPX_CUDA(cudaStreamCreateWithPriority(&h->aux2_stream, cudaStreamNonBlocking, hi));
    }
    GPX_CUDA(cudaEventCreateWithFlags(&h->ev_a, cudaEventDisableTiming));
    GPX_CUDA(cudaEventCreateWithFlags(&h->ev_b, cudaEventDisableTiming));
    *out = h;
    return 0;
}

int gpx_enter(gpx_ctx* h) {
    int cur = -1;
    GPX_CUDA(cudaGetDevice(&cur));
    if (cur != h->device) GPX_CUDA(cudaSetDevice(h->device));
    return 0;
}

extern "C" int gpx_destroy(gpx_handle h) {
    if (!h) return 0;
    cudaSetDevice(h->device);
    cudaDeviceSynchronize();
    gpx_timing_destroy(h);
    if (h->scratch) cudaFree(h->scratch);
    if (h->scratch2) cudaFree(h->scratch2);
    if (h->pinned) cudaFreeHost(h->pinned);
    if (h->d_small) cudaFree(h->d_small);
    if (h->d_info) cudaFree(h->d_info);
    if (h->d_partial) cudaFree(h->d_partial);
    if (h->d_theta) cudaFree(h->d_theta);
    if (h->aux_stream) cudaStreamDestroy(h->aux_stream);
    if (h->aux2_stream) cudaStreamDestroy(h->aux2_stream);
    if (h->graph_stream) cudaStreamDestroy(h->graph_stream);
    if (h->ev_a) cudaEventDestroy(h->ev_a);
    if (h->ev_b) cudaEventDestroy(h->ev_b);
    delete h;
    return 0;
}

extern "C" int gpx_set_stream(gpx_handle h, void* s) {
    GPX_ENTER(h);
    h->stream = (cudaStream_t)s;
    return 0;
}

extern "C" int gpx_synchronize(gpx_handle h) {
    GPX_ENTER(h);
    GPX_CUDA(cudaStreamSynchronize(h->stream));
    return 0;
}

extern "C" int64_t gpx_launch_count(gpx_handle h) { return h ? h->launches : 0; }

int gpx_scratch(gpx_ctx* h, size_t bytes, void** out) {
    if (h->scratch_bytes < bytes) {
        if (h->scratch) {
            cudaStreamSynchronize(h->stream);
            cudaFree(h->scratch);
            h->scratch = nullptr;
            h->scratch_bytes = 0;
        }
        cudaError_t e = cudaMalloc(&h->scratch, bytes);
        if (e != cudaSuccess) {
            gpx_set_error("gpx: cudaMalloc(%zu) for scratch failed: %s", bytes, cudaGetErrorString(e));
            return GPX_E_NOMEM;
        }
        h->scratch_bytes = bytes;
    }
    *out = h->scratch;
    return 0;
}

int gpx_scratch2(gpx_ctx* h, size_t bytes, void** out) {
    if (h->scratch2_bytes < bytes) {
        if (h->scratch2) {
            cudaStreamSynchronize(h->stream);
            cudaFree(h->scratch2);
            h->scratch2 = nullptr;
            h->scratch2_bytes = 0;
        }
        cudaError_t e = cudaMalloc(&h->scratch2, bytes);
        if (e != cudaSuccess) {
            gpx_set_error("gpx: cudaMalloc(%zu) for scratch2 failed: %s", bytes, cudaGetErrorString(e));
            return GPX_E_NOMEM;
        }
        h->scratch2_bytes = bytes;
    }
    *out = h->scratch2;
    return 0;
}

int gpx_read_info(gpx_ctx* h, int* info_host) {
    GPX_CUDA(cudaMemcpyAsync(info_host, h->d_info, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    GPX_CUDA(cudaStreamSynchronize(h->stream));
    return 0;
}

// ---------------------------------------------------------------------------------------------
// fused drivers
// ---------------------------------------------------------------------------------------------
extern "C" int gpx_gp_fit(gpx_handle h, int kind, const double* X, int64_t n, int D, const double* theta_host, int ntheta,
                          double s, const double* y, double* A, int64_t np_, int64_t lda, double* dinv, double* alpha,
                          double* out3) {
    GPX_ENTER(h);
    GPX_REQUIRE(np_ == gpx_padded_dim(n), 11);
    // K + s I, lower tiles only, identity padding            (tune...:306-307, CO2...:142-143)
    gpx_phase_mark(h, GPX_PH_COV);
    GPX_TRY(gpx_cov_build(h, kind, X, n, X, n, D, theta_host, ntheta, s, GPX_COV_SAME_X | GPX_COV_LOWER | GPX_COV_SKIP_UPPER, A, np_, np_, lda,
                          nullptr, 0));
    gpx_phase_mark(h, GPX_PH_POTRF);
    int info = gpx_potrf(h, A, np_, lda, dinv);
    if (info != 0) return info;
    gpx_phase_mark(h, GPX_PH_SOLVE);
    // alpha = L^-T (L^-1 y)                                   (tune...:308-309)
    GPX_CUDA(cudaMemsetAsync(alpha, 0, np_ * sizeof(double), h->stream));
    GPX_CUDA(cudaMemcpyAsync(alpha, y, n * sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
    const int bs = gpx_block_size_for(np_);
    if (bs > GPX_T && np_ >= 2 * bs) {
        // shorten the serial chain of the two solves with explicit inverses of the bs x bs diagonal blocks
        void* sc = nullptr;
        GPX_TRY(gpx_scratch2(h, ((size_t)np_ * bs + (size_t)np_ * bs / 4 + bs) * sizeof(double), &sc));
        double* Dbig = (double*)sc;
        double* work = Dbig + (size_t)np_ * bs;
        double* tmp = work + (size_t)np_ * bs / 4;
        GPX_TRY(gpx_block_inverses(h, A, np_, lda, dinv, bs, Dbig, work));
        GPX_TRY(gpx_trsv_big(h, A, np_, lda, Dbig, bs, 0, alpha, tmp));
        GPX_TRY(gpx_trsv_big(h, A, np_, lda, Dbig, bs, 1, alpha, tmp));
    } else {
        GPX_TRY(gpx_trsv(h, A, np_, lda, dinv, 0, alpha));
        GPX_TRY(gpx_trsv(h, A, np_, lda, dinv, 1, alpha));
    }
    int r = gpx_lml(h, A, n, lda, y, alpha, out3);           // tune...:312
    gpx_phase_mark(h, GPX_PH_END);
    return r;
}

extern "C" int gpx_gp_fit_grad(gpx_handle h, int kind, const double* X, int64_t n, int D, const double* theta_host,
                               int ntheta, double s, const double* y, double* A, int64_t np_, int64_t lda, double* dinv,
                               double* Kinv, double* alpha, double* out3, double* grad) {
    int r = gpx_gp_fit(h, kind, X, n, D, theta_host, ntheta, s, y, A, np_, lda, dinv, alpha, out3);
    if (r != 0) return r;
    // K_y^-1 = L^-T L^-1 (tune...:144): in-place triangular inverse, then one triangular SYRK into Kinv.
    // Kinv doubles as the trtri workspace (it is overwritten by lauum afterwards).
    gpx_phase_mark(h, GPX_PH_TRTRI);
    GPX_TRY(gpx_trtri(h, A, np_, lda, dinv, Kinv));
    gpx_phase_mark(h, GPX_PH_LAUUM);
    GPX_TRY(gpx_lauum(h, A, np_, lda, Kinv, lda));
    gpx_phase_mark(h, GPX_PH_GRAD);
    r = gpx_lml_grad(h, kind, X, n, D, theta_host, ntheta, Kinv, lda, alpha, grad);  // tune...:43-57
    gpx_phase_mark(h, GPX_PH_END);
    return r;
}

extern "C" int gpx_host_lml(gpx_handle h, int kind, const double* X, int64_t n, int D, const double* theta, int ntheta,
                            double s, const double* y, double* lml_out, double* grad_host) {
    GPX_ENTER(h);
    GPX_REQUIRE(n > 0, 4);
    const int64_t np_ = gpx_padded_dim(n);
    const int64_t nt = np_ / GPX_T;
    size_t bytes_A = (size_t)np_ * np_ * sizeof(double);
    size_t bytes = bytes_A * (grad_host ? 2 : 1) + (size_t)nt * GPX_T * GPX_T * sizeof(double) +
                   ((size_t)n * D + 2 * np_ + 32) * sizeof(double);
    void* base = nullptr;
    GPX_TRY(gpx_scratch(h, bytes, &base));
    double* A = (double*)base;
    double* Kinv = grad_host ? A + (size_t)np_ * np_ : nullptr;
    double* dinv = A + (size_t)np_ * np_ * (grad_host ? 2 : 1);
    double* dX = dinv + (size_t)nt * GPX_T * GPX_T;
    double* dy = dX + (size_t)n * D;
    double* dalpha = dy + np_;
    double* dout = dalpha + np_;
    GPX_CUDA(cudaMemcpyAsync(dX, X, (size_t)n * D * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    GPX_CUDA(cudaMemcpyAsync(dy, y, (size_t)n * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    int r;
    if (grad_host)
        r = gpx_gp_fit_grad(h, kind, dX, n, D, theta, ntheta, s, dy, A, np_, np_, dinv, Kinv, dalpha, dout, dout + 3);
    else
        r = gpx_gp_fit(h, kind, dX, n, D, theta, ntheta, s, dy, A, np_, np_, dinv, dalpha, dout);
    if (r != 0) return r;
    double host[3 + 11];
    GPX_CUDA(cudaMemcpyAsync(host, dout, (3 + (grad_host ? ntheta : 0)) * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    GPX_CUDA(cudaStreamSynchronize(h->stream));
    *lml_out = host[0];
    if (grad_host)
        for (int i = 0; i < ntheta; ++i) grad_host[i] = host[3 + i];
    return 0;
}

// ---------------------------------------------------------------------------------------------
// FP64 peak micro-benchmarks (register-resident issue loops)
// ---------------------------------------------------------------------------------------------
namespace {
// One 1024-thread block per SM (32 warps, 8 per scheduler), 8 independent accumulator chains per warp: exactly one
// wave on any SM count, so the result cannot depend on how the block scheduler packs several blocks per SM.
constexpr int PEAK_CHAINS = 8;
__global__ void __launch_bounds__(1024, 1) dmma_peak_kernel(int iters, double* out) {
    double c[PEAK_CHAINS][2];
#pragma unroll
    for (int i = 0; i < PEAK_CHAINS; ++i) c[i][0] = c[i][1] = 0.0;
    double a = 1.0 + threadIdx.x * 1e-9, b = 1.0 - threadIdx.x * 1e-9;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < PEAK_CHAINS; ++i)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                         : "+d"(c[i][0]), "+d"(c[i][1])
                         : "d"(a), "d"(b));
    }
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < PEAK_CHAINS; ++i) s += c[i][0] + c[i][1];
    if (s == 12345.678) out[0] = s;
}
__global__ void __launch_bounds__(1024, 1) dfma_peak_kernel(int iters, double* out) {
    double c[PEAK_CHAINS];
#pragma unroll
    for (int i = 0; i < PEAK_CHAINS; ++i) c[i] = threadIdx.x * 1e-3 + i;
    double a = 1.0 + threadIdx.x * 1e-9, b = 1e-9;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < PEAK_CHAINS; ++i) c[i] = fma(c[i], a, b);
    }
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < PEAK_CHAINS; ++i) s += c[i];
    if (s == 12345.678) out[0] = s;
}
// mixed: even warps issue DMMA, odd warps issue DFMA (8x the iterations: one DFMA warp-instruction is
// 2 issue cycles, one DMMA 16) -- tells whether the two FP64 pipes run concurrently (they do not: 35.5 TF).
__global__ void __launch_bounds__(1024, 1) mixed_peak_kernel(int iters, double* out) {
    const int warp = threadIdx.x >> 5;
    double s = 0.0;
    double a = 1.0 + threadIdx.x * 1e-9, b = 1.0 - threadIdx.x * 1e-9;
    if (warp & 1) {
        double c[PEAK_CHAINS];
#pragma unroll
        for (int i = 0; i < PEAK_CHAINS; ++i) c[i] = threadIdx.x * 1e-3 + i;
        for (int it = 0; it < iters * 8; ++it) {
#pragma unroll
            for (int i = 0; i < PEAK_CHAINS; ++i) c[i] = fma(c[i], a, 1e-9);
        }
#pragma unroll
        for (int i = 0; i < PEAK_CHAINS; ++i) s += c[i];
    } else {
        double c[PEAK_CHAINS][2];
#pragma unroll
        for (int i = 0; i < PEAK_CHAINS; ++i) c[i][0] = c[i][1] = 0.0;
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int i = 0; i < PEAK_CHAINS; ++i)
                asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                             : "+d"(c[i][0]), "+d"(c[i][1])
                             : "d"(a), "d"(b));
        }
#pragma unroll
        for (int i = 0; i < PEAK_CHAINS; ++i) s += c[i][0] + c[i][1];
    }
    if (s == 12345.678) out[0] = s;
}
}  // namespace

extern "C" int gpx_bench_fp64_peak(gpx_handle h, int use_dmma, int iters, double* tflops_out, double* ms_out) {
    GPX_ENTER(h);
    cudaEvent_t e0, e1;
    GPX_CUDA(cudaEventCreate(&e0));
    GPX_CUDA(cudaEventCreate(&e1));
    cudaDeviceProp prop;
    GPX_CUDA(cudaGetDeviceProperties(&prop, h->device));
    const int blocks = prop.multiProcessorCount, threads = 1024;
    double best = 1e30;
    // An idle GPU needs up to ~1 s of load to reach its boost clocks: repeat until the best time stops improving.
    int stale = 0;
    for (int rep = 0; rep < 400 && stale < 12; ++rep) {
        // one untimed launch keeps the queue busy so that host-side launch latency cannot leak between the events
        for (int k = 0; k < 2; ++k) {
            if (k == 1) GPX_CUDA(cudaEventRecord(e0, h->stream));
            if (use_dmma == 2) mixed_peak_kernel<<<blocks, threads, 0, h->stream>>>(iters, h->d_theta);
            else if (use_dmma) dmma_peak_kernel<<<blocks, threads, 0, h->stream>>>(iters, h->d_theta);
            else dfma_peak_kernel<<<blocks, threads, 0, h->stream>>>(iters, h->d_theta);
            GPX_CHECK_LAUNCH(h);
        }
        GPX_CUDA(cudaEventRecord(e1, h->stream));
        GPX_CUDA(cudaEventSynchronize(e1));
        float ms = 0;
        GPX_CUDA(cudaEventElapsedTime(&ms, e0, e1));
        if (ms < best * 0.995) { best = ms; stale = 0; } else { if (ms < best) best = ms; ++stale; }
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    const double warps = (double)blocks * threads / 32.0;
    const double ch = PEAK_CHAINS;
    double flops = use_dmma ? warps * iters * ch * 512.0 : warps * iters * ch * 64.0;
    if (use_dmma == 2) flops = 0.5 * warps * iters * ch * 512.0 + 0.5 * warps * iters * 8.0 * ch * 64.0;
    *ms_out = best;
    *tflops_out = flops / (best * 1e-3) / 1e12;
    return 0;
}
