// Shared declarations for libgpx (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include "../../include/gpx.h"

#define GPX_T 128  // tile edge used by every dense kernel

struct gpx_ctx {
    int device;
    cudaStream_t stream;
    int64_t launches;
    // scratch owned by the handle (grown on demand)
    void* scratch;       size_t scratch_bytes;
    void* scratch2;      size_t scratch2_bytes;   // block inverses of the fused fit drivers
    void* pinned;        size_t pinned_bytes;     // host staging buffer of the host-pointer entry points (small.cu)
    double* d_small;     int small_n;             // posterior factor kept by gpx_gp_small_fit_host (small.cu)
    int* d_info;         // device int: first failing pivot (1-based) or 0
    double* d_partial;   // reduction partials
    size_t partial_elems;
    double* d_theta;     // 16 doubles of hyper-parameters for kernels
    // NCCL (optional, dlopen'ed)
    void* nccl_comm; int rank, world;
    cudaStream_t aux_stream, aux2_stream; cudaEvent_t ev_a, ev_b;
    cudaStream_t graph_stream;          // private capturable stream of the optimiser loop (ascent.cu), created on demand
    // optional instrumentation (timing.cu)
    int timing_on; void* timing;
};

// phases of the fused drivers (gpx_timing_collect out[3 + phase])
enum { GPX_PH_COV = 0, GPX_PH_POTRF = 1, GPX_PH_SOLVE = 2, GPX_PH_TRTRI = 3, GPX_PH_LAUUM = 4, GPX_PH_GRAD = 5,
       GPX_PH_END = 6, GPX_NPHASES = 8 };
void gpx_timing_gemm_begin(gpx_ctx* h, double flops_exec, int M = 0, int N = 0, int K = 0);
void gpx_timing_gemm_end(gpx_ctx* h);
void gpx_timing_leaf_begin(gpx_ctx* h);
void gpx_timing_leaf_end(gpx_ctx* h);
void gpx_phase_mark(gpx_ctx* h, int phase);
void gpx_timing_destroy(gpx_ctx* h);

void gpx_set_error(const char* fmt, ...);

#define GPX_CUDA(call)                                                                       \
    do {                                                                                     \
        cudaError_t e__ = (call);                                                            \
        if (e__ != cudaSuccess) {                                                            \
            gpx_set_error("%s:%d %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e__)); \
            return GPX_E_CUDA;                                                               \
        }                                                                                    \
    } while (0)

#define GPX_CHECK_LAUNCH(h)                                                                  \
    do {                                                                                     \
        (h)->launches++;                                                                     \
        cudaError_t e__ = cudaGetLastError();                                                \
        if (e__ != cudaSuccess) {                                                            \
            gpx_set_error("%s:%d kernel launch -> %s", __FILE__, __LINE__, cudaGetErrorString(e__)); \
            return GPX_E_CUDA;                                                               \
        }                                                                                    \
    } while (0)

#define GPX_REQUIRE(cond, argno)                                                             \
    do {                                                                                     \
        if (!(cond)) {                                                                       \
            gpx_set_error("%s:%d bad argument %d: %s", __FILE__, __LINE__, (argno), #cond);  \
            return -(argno);                                                                 \
        }                                                                                    \
    } while (0)

// every extern "C" entry point: validate the handle and make its device current (a process may hold handles on
// several GPUs; kernel attributes, streams and allocations are per device)
#define GPX_MAX_DEVICES 64
int gpx_enter(gpx_ctx* h);
#define GPX_ENTER(h)                                                                         \
    do {                                                                                     \
        GPX_REQUIRE((h) != nullptr, 1);                                                      \
        int e__ = gpx_enter(h);                                                              \
        if (e__ != 0) return e__;                                                            \
    } while (0)

#define GPX_TRY(call)                                                                        \
    do {                                                                                     \
        int r__ = (call);                                                                    \
        if (r__ != 0) return r__;                                                            \
    } while (0)

// ---- internal (C++) interfaces between translation units -------------------------------------
struct GemmArgs {
    const double* A; const double* B; double* C;
    int M, N, K;                // M,N multiples of 128; K multiple of 16
    int64_t lda, ldb, ldc;
    double alpha, beta;
    int64_t sA, sB, sC; int batch;   // strided batch (elements)
    int64_t sA2, sB2, sC2; int batch2;  // optional outer batch level: blockIdx.z = z2 * batch + z1
    int a_kmajor, b_kmajor;     // 1: operand stored with k contiguous ([M][K] / [N][K]); 0: [K][M] / [K][N]
    int lower_only;             // skip output tiles strictly above the diagonal
    int kb_mode;                // first k: 0 -> 0, 1 -> tile row0, 2 -> tile col0, 3 -> block-cyclic mapped column position
    int ke_mode;                // last  k: 0 -> K, 1 -> tile row0+128, 2 -> tile col0+128
    int rev_rows;               // schedule tile rows in reverse (longest k-loops first)
    int64_t kb_batch;           // added to the first k per batch index (batched triangular products)
    int kb_const;               // constant added to the first k
    // block-cyclic column map (multi-GPU trailing update): local column tile bn of C corresponds to the
    // global tile ((bn / cyc_tpb + cyc_q0) * cyc_P + cyc_p) * cyc_tpb + bn % cyc_tpb; its offset relative to
    // cyc_row_base (global row of C row 0) is used for the lower-only test and as B's row offset.
    int cyc_P, cyc_p, cyc_tpb, cyc_q0, cyc_row_base;
    int cyc_snake;              // block -> rank map: 0 plain cyclic (q*P + p), 1 boustrophedon (see gpx_cyc_global)
    int cyc_b_rows;             // 1: the mapped position is also B's row offset (trailing update); 0: B uses the local column
};
// ---- block -> rank maps of the multi-GPU layout.  Plain block-cyclic gives rank 0 the longest columns of a triangular
// matrix (9 % more work than rank 7 at N = 65536, nb = 256, P = 8); the boustrophedon ("snake") order 0..P-1, P-1..0, ...
// pairs a long column with a short one on every rank and balances the triangular work to second order.
__host__ __device__ inline int64_t gpx_cyc_global(int64_t q, int P, int p, int snake) {   // global block of local block q
    if (!snake) return q * P + p;
    return (q >> 1) * 2 * P + ((q & 1) ? 2 * P - 1 - p : p);
}
__host__ __device__ inline int gpx_cyc_owner(int64_t j, int P, int snake) {
    if (!snake) return (int)(j % P);
    const int pos = (int)(j % (2 * P));
    return pos < P ? pos : 2 * P - 1 - pos;
}
__host__ __device__ inline int64_t gpx_cyc_local(int64_t j, int P, int snake) {            // local index of block j on its owner
    if (!snake) return j / P;
    return 2 * (j / (2 * P)) + ((j % (2 * P)) >= P ? 1 : 0);
}
__host__ __device__ inline int64_t gpx_cyc_count_below(int64_t j, int P, int p, int snake) {   // # local blocks with global index < j
    if (j <= 0) return 0;
    if (!snake) return j > p ? (j - p + P - 1) / P : 0;
    const int64_t cyc = j / (2 * P), rem = j % (2 * P);
    return 2 * cyc + (rem > p ? 1 : 0) + (rem > 2 * P - 1 - p ? 1 : 0);
}

int gpx_gemm_launch(gpx_ctx* h, const GemmArgs& a);
int gpx_gemm_tma_try_launch(gpx_ctx* h, const GemmArgs& a, double flops_exec);

int gpx_lml_grad_block(gpx_ctx* h, int kind, const double* X, int64_t n, int D, const double* theta_host, int ntheta,
                       const double* Kinv, int64_t ldk, const double* alpha, double* grad, int64_t rows, int64_t cols, int rg0,
                       int cg0, const double* theta_dev = nullptr);
int gpx_cov_build_block(gpx_ctx* h, int kind, const double* X, int64_t n, int D, const double* theta_host, int ntheta,
                        double diag_add, int flags, double* K, int64_t rows, int64_t cols, int64_t ldk, int rg0, int cg0,
                        const double* scale = nullptr, const double* theta_dev = nullptr);
int gpx_potrf_block(gpx_ctx* h, double* A, int64_t n, int64_t lda, double* dinv, int goff);
int gpx_trsm_right_lt_block(gpx_ctx* h, double* B, int64_t m, int64_t ldb, const double* L, int64_t n, int64_t ldl,
                            const double* dinv);
int gpx_panel_factor_sub(gpx_ctx* h, double* P, int64_t rows, int64_t ld, int nb, double* dinv, int goff);
int gpx_trsm_left_prefix_block(gpx_ctx* h, const double* L, int64_t n, int64_t ldl, const double* dinv, double* B, int64_t ldb,
                               int P, int p, int nb, int snake, const double* D = nullptr, int bs = 0, double* tmp = nullptr);
int gpx_trsm_left_prefix_trans_block(gpx_ctx* h, const double* L, int64_t n, int64_t ldl, const double* dinv, double* B,
                                     int64_t ldb, int P, int p, int nb, int snake, const double* D = nullptr, int bs = 0,
                                     double* tmp = nullptr);
int gpx_scratch(gpx_ctx* h, size_t bytes, void** out);
int gpx_scratch2(gpx_ctx* h, size_t bytes, void** out);
int gpx_read_info(gpx_ctx* h, int* info_host);

// ---- RAII set of timing-less events (multi-stream drivers): destroyed on every exit path (destroying a pending event only
// defers its release)
#ifdef __cplusplus
#include <vector>
struct GpxEventSet {
    std::vector<cudaEvent_t> ev;
    int create(size_t count) {
        ev.assign(count, nullptr);
        for (auto& e : ev)
            if (cudaEventCreateWithFlags(&e, cudaEventDisableTiming) != cudaSuccess) {
                gpx_set_error("gpx: cudaEventCreate failed");
                return GPX_E_CUDA;
            }
        return 0;
    }
    ~GpxEventSet() {
        for (auto e : ev)
            if (e) cudaEventDestroy(e);
    }
};
#endif
