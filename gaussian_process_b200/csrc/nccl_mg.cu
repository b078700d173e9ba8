// Multi-GPU exact-GP fit + LML + gradient (one process per GPU).
//
// Layout: 1-D block-cyclic over block columns of width nb (a P x 1 process grid of the 2-D block-cyclic
// scheme; on NVSwitch every rank sees every panel at full bandwidth, so the second grid dimension buys
// nothing -- SURVEY 8e "Cholesky").  Rank p owns global block columns j = q*P + p, stored side by side in a
// local row-major matrix Aloc[npad][nloc*nb].
//   potrf : right-looking; the owner factors the diagonal block + TRSMs the panel (recursive DMMA kernels),
//           the panel is broadcast with NCCL on a communication stream and *kept* by every rank in a
//           replicated factor Lfull, the trailing update of the local columns is one block-cyclic-mapped DMMA
//           launch.  Look-ahead: the owner of panel j+1 updates and factors it before the rest of update j, so
//           the broadcast of panel j+1 overlaps update j.
//   alpha : TRSVs on the replicated factor (redundant on every rank, no communication).
//   L^-1  : every rank solves L X = E for its own block columns only (prefix-structured recursive TRSM),
//           then one all-gather replicates X.
//   K^-1  : local block columns of X^T X by batched triangular DMMA products.
//   grad  : fused trace kernel over the local block columns of K^-1, all-reduce of the ntheta partial sums.
// NCCL is dlopen'ed (torch-bundled libnccl.so.2); with world == 1 no NCCL symbol is touched.  For tests the same
// per-rank routines can be driven for P virtual ranks inside one process (gpx_mg_emulate_fit_grad).
#include <dlfcn.h>
#include <vector>
#include "common.cuh"

namespace {

// ------------------------------------------------------------------------------------------------ NCCL (dlopen)
typedef struct { char internal[128]; } ncclUniqueId_t;
typedef void* ncclComm_p;
typedef int (*fn_GetUniqueId)(ncclUniqueId_t*);
typedef int (*fn_CommInitRank)(ncclComm_p*, int, ncclUniqueId_t, int);
typedef int (*fn_CommDestroy)(ncclComm_p);
typedef int (*fn_Broadcast)(const void*, void*, size_t, int, int, ncclComm_p, cudaStream_t);
typedef int (*fn_AllGather)(const void*, void*, size_t, int, ncclComm_p, cudaStream_t);
typedef int (*fn_AllReduce)(const void*, void*, size_t, int, int, ncclComm_p, cudaStream_t);
typedef const char* (*fn_GetErrorString)(int);
struct NcclApi {
    void* lib = nullptr;
    fn_GetUniqueId GetUniqueId = nullptr;
    fn_CommInitRank CommInitRank = nullptr;
    fn_CommDestroy CommDestroy = nullptr;
    fn_Broadcast Broadcast = nullptr;
    fn_AllGather AllGather = nullptr;
    fn_AllReduce AllReduce = nullptr;
    fn_GetErrorString GetErrorString = nullptr;
} g_nccl;
constexpr int NCCL_F64 = 8, NCCL_I32 = 2, NCCL_SUM = 0, NCCL_MAX = 2;

int nccl_load(const char* path) {
    if (g_nccl.lib) return 0;
    void* lib = nullptr;
    if (path && path[0]) lib = dlopen(path, RTLD_NOW | RTLD_GLOBAL);
    if (!lib) lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!lib) {
        gpx_set_error("gpx: cannot dlopen libnccl.so.2 (%s)", dlerror());
        return GPX_E_NCCL;
    }
    g_nccl.lib = lib;
#define LOAD(name) g_nccl.name = (fn_##name)dlsym(lib, "nccl" #name); if (!g_nccl.name) { gpx_set_error("gpx: nccl" #name " missing"); return GPX_E_NCCL; }
    LOAD(GetUniqueId) LOAD(CommInitRank) LOAD(CommDestroy) LOAD(Broadcast) LOAD(AllGather) LOAD(AllReduce) LOAD(GetErrorString)
#undef LOAD
    return 0;
}
#define GPX_NCCL(call)                                                                          \
    do {                                                                                        \
        int r__ = (call);                                                                       \
        if (r__ != 0) {                                                                         \
            gpx_set_error("%s:%d %s -> %s", __FILE__, __LINE__, #call, g_nccl.GetErrorString(r__)); \
            return GPX_E_NCCL;                                                                  \
        }                                                                                       \
    } while (0)

// ------------------------------------------------------------------------------------------------ per-rank state
struct MgRank {
    gpx_ctx* h;         // handle whose stream the rank's kernels run on
    int P, p;           // world, rank
    int64_t n, npad;    // true / padded size (npad % (nb*P) == 0)
    int nb, tpb;        // block width, tiles per block
    int64_t nblk, nloc, wloc;   // global blocks, local blocks, local width (elements)
    double *Aloc, *Lfull, *Xall, *Kloc, *dinv, *stage[2];
};

__global__ void set_identity_blocks_kernel(double* X, int64_t ld, int64_t nloc, int nb, int P, int p) {
    // X[(q*P+p)*nb + i][q*nb + i] = 1 for every local block q
    int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= nloc * nb) return;
    int64_t q = idx / nb, i = idx - q * nb;
    X[((q * P + p) * nb + i) * ld + q * nb + i] = 1.0;
}

size_t stage_elems(const MgRank& r) { return (size_t)r.npad * r.nb + (size_t)r.tpb * GPX_T * GPX_T; }

// owner side: factor diagonal block j and TRSM the panel below it (in Aloc), then pack panel + leaf inverses
int panel_factor_pack(MgRank& r, int64_t j, double* stage) {
    gpx_ctx* h = r.h;
    const int64_t q = j / r.P, r0 = j * r.nb, rows = r.npad - r0;
    double* diag = r.Aloc + r0 * r.wloc + q * r.nb;
    double* dinvj = r.dinv + j * r.tpb * GPX_T * GPX_T;
    GPX_TRY(gpx_potrf_block(h, diag, r.nb, r.wloc, dinvj, (int)r0));
    if (rows > r.nb)
        GPX_TRY(gpx_trsm_right_lt_block(h, diag + (int64_t)r.nb * r.wloc, rows - r.nb, r.wloc, diag, r.nb, r.wloc, dinvj));
    GPX_CUDA(cudaMemcpy2DAsync(stage, r.nb * sizeof(double), diag, r.wloc * sizeof(double), r.nb * sizeof(double), rows,
                               cudaMemcpyDeviceToDevice, h->stream));
    GPX_CUDA(cudaMemcpyAsync(stage + (size_t)rows * r.nb, dinvj, (size_t)r.tpb * GPX_T * GPX_T * sizeof(double),
                             cudaMemcpyDeviceToDevice, h->stream));
    return 0;
}

// every rank: panel j (contiguous rows x nb) -> replicated factor + leaf inverses
int panel_unpack(MgRank& r, int64_t j, const double* stage, cudaStream_t s) {
    const int64_t r0 = j * r.nb, rows = r.npad - r0;
    GPX_CUDA(cudaMemcpy2DAsync(r.Lfull + r0 * r.npad + r0, r.npad * sizeof(double), stage, r.nb * sizeof(double),
                               r.nb * sizeof(double), rows, cudaMemcpyDeviceToDevice, s));
    GPX_CUDA(cudaMemcpyAsync(r.dinv + j * r.tpb * GPX_T * GPX_T, stage + (size_t)rows * r.nb,
                             (size_t)r.tpb * GPX_T * GPX_T * sizeof(double), cudaMemcpyDeviceToDevice, s));
    return 0;
}

// trailing update with panel j of the local blocks q in [q_lo, q_hi): Aloc[k:, q] -= L[k:, j] L[k-block, j]^T
int trailing_update(MgRank& r, int64_t j, int64_t q_lo, int64_t q_hi) {
    if (q_lo >= q_hi) return 0;
    const int64_t gk0 = q_lo * r.P + r.p;            // global block of the first updated local column
    const int64_t r_start = gk0 * r.nb;
    GemmArgs a{};
    a.batch = 1;
    a.alpha = -1.0; a.beta = 1.0;
    a.A = r.Lfull + r_start * r.npad + j * r.nb; a.lda = r.npad; a.a_kmajor = 1;
    a.B = a.A; a.ldb = r.npad; a.b_kmajor = 1;
    a.C = r.Aloc + r_start * r.wloc + q_lo * r.nb; a.ldc = r.wloc;
    a.M = (int)(r.npad - r_start); a.N = (int)((q_hi - q_lo) * r.nb); a.K = r.nb;
    a.lower_only = 1;
    a.cyc_P = r.P; a.cyc_p = r.p; a.cyc_tpb = r.tpb; a.cyc_q0 = (int)q_lo; a.cyc_row_base = (int)r_start; a.cyc_b_rows = 1;
    return gpx_gemm_launch(r.h, a);
}

int64_t first_local_block_after(const MgRank& r, int64_t j) {  // smallest q with q*P + p > j
    if (j < r.p) return 0;
    return (j - r.p) / r.P + 1;
}

// Xall[rank][npad][wloc] (rank-major, as all-gathered) -> Xf[npad][npad] row-major in GLOBAL column order; only rows
// at/below each block's diagonal are copied (the rest of X is zero and never read).  Xf reuses the Lfull buffer.
int reorder_X(MgRank& r) {
    for (int64_t i = 0; i < r.nblk; ++i) {
        const int src = (int)(i % r.P);
        const int64_t q = i / r.P, r0 = i * r.nb;
        const double* from = r.Xall + (size_t)src * r.npad * r.wloc + r0 * r.wloc + q * r.nb;
        double* to = r.Lfull + r0 * r.npad + r0;
        GPX_CUDA(cudaMemcpy2DAsync(to, r.npad * sizeof(double), from, r.wloc * sizeof(double), r.nb * sizeof(double),
                                   r.npad - r0, cudaMemcpyDeviceToDevice, r.h->stream));
    }
    return 0;
}

// local block columns of K^-1 = X^T X (tiles on/below the diagonal): ONE triangular TMA-fed DMMA launch for all owned
// block columns -- C = Kloc[npad][wloc], A(m,k) = X[k][m], B(k,n) = X[k][gcol(n)] with the block-cyclic column map,
// k >= row tile (X is lower triangular), tiles above the diagonal skipped; row tiles are scheduled longest-k first.
int lauum_local(MgRank& r) {
    const double* Xf = r.Lfull;
    GemmArgs a{};
    a.alpha = 1.0; a.beta = 0.0;
    a.batch = 1;
    a.A = Xf; a.lda = r.npad; a.a_kmajor = 0;
    a.B = Xf; a.ldb = r.npad; a.b_kmajor = 0;
    a.C = r.Kloc; a.ldc = r.wloc;
    a.M = (int)r.npad; a.N = (int)r.wloc; a.K = (int)r.npad;
    a.kb_mode = 1;
    a.lower_only = 1;
    a.cyc_P = r.P; a.cyc_p = r.p; a.cyc_tpb = r.tpb; a.cyc_q0 = 0; a.cyc_row_base = 0; a.cyc_b_rows = 1;
    return gpx_gemm_launch(r.h, a);
}

__global__ void accumulate_kernel(int n, const double* __restrict__ x, double* __restrict__ acc) {
    int i = threadIdx.x;
    if (i < n) acc[i] += x[i];
}

int grad_local(MgRank& r, int kind, const double* X, int D, const double* theta, int ntheta, const double* alpha,
               double* grad_acc /* device, ntheta, zeroed by caller */, double* tmp /* device, ntheta */) {
    for (int64_t q = 0; q < r.nloc; ++q) {
        const int64_t j = q * r.P + r.p, r0 = j * r.nb;
        GPX_TRY(gpx_lml_grad_block(r.h, kind, X, r.n, D, theta, ntheta, r.Kloc + r0 * r.wloc + q * r.nb, r.wloc, alpha, tmp,
                                   r.npad - r0, r.nb, (int)r0, (int)r0));
        accumulate_kernel<<<1, 32, 0, r.h->stream>>>(ntheta, tmp, grad_acc);
        GPX_CHECK_LAUNCH(r.h);
    }
    return 0;
}

int cov_local(MgRank& r, int kind, const double* X, int D, const double* theta, int ntheta, double s) {
    for (int64_t q = 0; q < r.nloc; ++q) {
        const int64_t j = q * r.P + r.p;
        GPX_TRY(gpx_cov_build_block(r.h, kind, X, r.n, D, theta, ntheta, s, GPX_COV_SAME_X | GPX_COV_LOWER,
                                    r.Aloc + q * r.nb, r.npad, r.nb, r.wloc, 0, (int)(j * r.nb)));
    }
    return 0;
}

int trtri_local(MgRank& r) {
    double* Xloc = r.Xall + (size_t)r.p * r.npad * r.wloc;
    GPX_CUDA(cudaMemsetAsync(Xloc, 0, (size_t)r.npad * r.wloc * sizeof(double), r.h->stream));
    const int64_t cnt = r.nloc * r.nb;
    set_identity_blocks_kernel<<<(unsigned)((cnt + 255) / 256), 256, 0, r.h->stream>>>(Xloc, r.wloc, r.nloc, r.nb, r.P, r.p);
    GPX_CHECK_LAUNCH(r.h);
    return gpx_trsm_left_prefix_block(r.h, r.Lfull, r.npad, r.npad, r.dinv, Xloc, r.wloc, r.P, r.p, r.nb);
}

int solve_lml(MgRank& r, const double* y, double* alpha, double* out3) {
    gpx_ctx* h = r.h;
    GPX_CUDA(cudaMemsetAsync(alpha, 0, r.npad * sizeof(double), h->stream));
    GPX_CUDA(cudaMemcpyAsync(alpha, y, r.n * sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
    GPX_TRY(gpx_trsv(h, r.Lfull, r.npad, r.npad, r.dinv, 0, alpha));
    GPX_TRY(gpx_trsv(h, r.Lfull, r.npad, r.npad, r.dinv, 1, alpha));
    return gpx_lml(h, r.Lfull, r.n, r.npad, y, alpha, out3);
}

// alpha = X^T (X y) from the replicated inverse factor X = L^-1 (row-major Xf in the Lfull buffer): two sweeps of
// HBM-bound GEMVs with no sequential dependency (the reference's CO2 path forms alpha the same way, CO2...:144-145).
// `diag` holds diag(L) saved before the factor buffer was recycled; out3 as gpx_lml.
int solve_lml_from_inverse(MgRank& r, const double* y, double* alpha, double* tmp, const double* diag, double* out3) {
    gpx_ctx* h = r.h;
    const double* Xf = r.Lfull;
    GPX_CUDA(cudaMemsetAsync(tmp, 0, r.npad * sizeof(double), h->stream));
    GPX_CUDA(cudaMemsetAsync(alpha, 0, r.npad * sizeof(double), h->stream));
    GPX_CUDA(cudaMemcpyAsync(alpha, y, r.n * sizeof(double), cudaMemcpyDeviceToDevice, h->stream));   // padded y
    for (int64_t j = 0; j < r.nblk; ++j) {      // tmp = X y  (panel j contributes to rows >= j*nb)
        const int64_t r0 = j * r.nb;
        GPX_TRY(gpx_gemv(h, 0, r.npad - r0, r.nb, 1.0, Xf + r0 * r.npad + r0, r.npad, alpha + r0, 1.0, tmp + r0));
    }
    for (int64_t j = 0; j < r.nblk; ++j) {      // alpha_j = X_j^T tmp
        const int64_t r0 = j * r.nb;
        GPX_TRY(gpx_gemv(h, 1, r.npad - r0, r.nb, 1.0, Xf + r0 * r.npad + r0, r.npad, tmp + r0, 0.0, alpha + r0));
    }
    return gpx_lml(h, diag, r.n, 0, y, alpha, out3);   // ldl = 0: diag[i*0 + i]
}

int init_rank(MgRank& r, gpx_ctx* h, int P, int p, int64_t n, int nb, double* ws) {
    r.h = h; r.P = P; r.p = p; r.n = n; r.nb = nb; r.tpb = nb / GPX_T;
    const int64_t unit = (int64_t)nb * P;
    r.npad = ((n + unit - 1) / unit) * unit;
    r.nblk = r.npad / nb; r.nloc = r.nblk / P; r.wloc = r.nloc * nb;
    double* w = ws;
    r.Aloc = w; w += (size_t)r.npad * r.wloc;
    r.Kloc = w; w += (size_t)r.npad * r.wloc;
    r.Lfull = w; w += (size_t)r.npad * r.npad;
    r.Xall = w; w += (size_t)r.npad * r.npad;
    r.dinv = w; w += (size_t)(r.npad / GPX_T) * GPX_T * GPX_T;
    r.stage[0] = w; w += stage_elems(r);
    r.stage[1] = w; w += stage_elems(r);
    return 0;
}

}  // namespace

// sum over the ranks of the handle's communicator, in place, on the handle's stream (no-op for a single rank)
int gpx_nccl_allreduce_sum(gpx_ctx* h, double* buf, size_t count) {
    if (h->world <= 1 || h->nccl_comm == nullptr) return 0;
    GPX_NCCL(g_nccl.AllReduce(buf, buf, count, NCCL_F64, NCCL_SUM, (ncclComm_p)h->nccl_comm, h->stream));
    return 0;
}

extern "C" int64_t gpx_mg_padded_dim(int64_t n, int nb, int world) {
    const int64_t unit = (int64_t)nb * world;
    return ((n + unit - 1) / unit) * unit;
}

// doubles of device workspace one rank needs for gpx_mg_fit_grad
extern "C" int64_t gpx_mg_workspace_elems(int64_t n, int nb, int world) {
    const int64_t npad = gpx_mg_padded_dim(n, nb, world);
    const int64_t wloc = npad / world;
    return 2 * npad * wloc + 2 * npad * npad + (npad / GPX_T) * GPX_T * GPX_T + 2 * (npad * nb + (nb / GPX_T) * GPX_T * GPX_T);
}

extern "C" int gpx_nccl_load(const char* path) { return nccl_load(path); }

extern "C" int gpx_nccl_unique_id(void* id128) {
    GPX_TRY(nccl_load(nullptr));
    GPX_NCCL(g_nccl.GetUniqueId((ncclUniqueId_t*)id128));
    return 0;
}

extern "C" int gpx_nccl_init(gpx_handle h, const void* id128, int rank, int world) {
    GPX_ENTER(h);
    GPX_TRY(nccl_load(nullptr));
    ncclUniqueId_t id;
    memcpy(&id, id128, sizeof(id));
    ncclComm_p comm = nullptr;
    GPX_NCCL(g_nccl.CommInitRank(&comm, world, id, rank));
    h->nccl_comm = comm;
    h->rank = rank;
    h->world = world;
    return 0;
}

// One rank's part of the distributed fit + LML + gradient.  `ws` = gpx_mg_workspace_elems doubles of device memory.
// out3 (device) = {lml, y.alpha, sum log diag}; grad (device) = ntheta doubles (already all-reduced).
extern "C" int gpx_mg_fit_grad(gpx_handle h, int kind, const double* X, int64_t n, int D, const double* theta_host, int ntheta,
                               double s, const double* y, int nb, double* ws, double* alpha, double* out3, double* grad,
                               int with_grad) {
    GPX_ENTER(h);
    GPX_REQUIRE(nb >= GPX_T && nb % GPX_T == 0, 10);
    const int P = h->world, p = h->rank;
    GPX_REQUIRE(P == 1 || h->nccl_comm != nullptr, 1);
    MgRank r;
    init_rank(r, h, P, p, n, nb, ws);
    // Three streams: S = bulk trailing updates, H = panel chain (high priority), Cs = NCCL broadcasts + unpack.
    //   after panel j is received:  H: owner(j+1): update column j+1, factor + pack panel j+1   (-> Cs broadcasts it)
    //                                  owner(j+2): [after bulk update j-1] update column j+2
    //                               S: update the local block columns with global index >= j+3
    // so the latency-bound chain runs concurrently with, and up to two panels ahead of, the DMMA-bound bulk updates.
    cudaStream_t S = h->stream, Cs = h->aux_stream, H = h->aux2_stream;
    std::vector<cudaEvent_t> evPanel(r.nblk), evRecv(r.nblk), evS(r.nblk);
    for (int64_t j = 0; j < r.nblk; ++j) {
        GPX_CUDA(cudaEventCreateWithFlags(&evPanel[j], cudaEventDisableTiming));
        GPX_CUDA(cudaEventCreateWithFlags(&evRecv[j], cudaEventDisableTiming));
        GPX_CUDA(cudaEventCreateWithFlags(&evS[j], cudaEventDisableTiming));
    }
    cudaEvent_t ev_start;
    GPX_CUDA(cudaEventCreateWithFlags(&ev_start, cudaEventDisableTiming));
    GPX_CUDA(cudaMemsetAsync(h->d_info, 0, sizeof(int), S));
    GPX_CUDA(cudaMemsetAsync(r.Lfull, 0, (size_t)r.npad * r.npad * sizeof(double), S));
    gpx_phase_mark(h, GPX_PH_COV);
    GPX_TRY(cov_local(r, kind, X, D, theta_host, ntheta, s));
    gpx_phase_mark(h, GPX_PH_POTRF);
    GPX_CUDA(cudaEventRecord(ev_start, S));
    GPX_CUDA(cudaStreamWaitEvent(Cs, ev_start, 0));   // Lfull memset / covariance build precede the chain
    GPX_CUDA(cudaStreamWaitEvent(H, ev_start, 0));
    if (P > 2) {
        // ---- many ranks: the owner of panel j+1 runs its chain (column update, factor, pack) on the bulk stream BEFORE
        // the rest of update j, so the chain never shares SMs with the bulk GEMM and every other rank gets the broadcast
        // as early as possible (measured at 8 GPUs: 464 ms vs 509-516 ms for the concurrent schedule below)
        cudaEvent_t evP2[2], evR2[2];
        for (int i = 0; i < 2; ++i) {
            GPX_CUDA(cudaEventCreateWithFlags(&evP2[i], cudaEventDisableTiming));
            GPX_CUDA(cudaEventCreateWithFlags(&evR2[i], cudaEventDisableTiming));
        }
        // ---- right-looking block-cyclic Cholesky with look-ahead
        if (r.p == 0) GPX_TRY(panel_factor_pack(r, 0, r.stage[0]));
        GPX_CUDA(cudaEventRecord(evP2[0], S));
        for (int64_t j = 0; j < r.nblk; ++j) {
            const int sb = (int)(j & 1);
            const int owner = (int)(j % P);
            const int64_t rows = r.npad - j * r.nb;
            const size_t count = (size_t)rows * r.nb + (size_t)r.tpb * GPX_T * GPX_T;
            GPX_CUDA(cudaStreamWaitEvent(Cs, evP2[sb], 0));
            if (P > 1) GPX_NCCL(g_nccl.Broadcast(r.stage[sb], r.stage[sb], count, NCCL_F64, owner, (ncclComm_p)h->nccl_comm, Cs));
            GPX_TRY(panel_unpack(r, j, r.stage[sb], Cs));
            GPX_CUDA(cudaEventRecord(evR2[sb], Cs));
            GPX_CUDA(cudaStreamWaitEvent(S, evR2[sb], 0));
            const int64_t q_first = first_local_block_after(r, j);
            if (j + 1 < r.nblk && (int)((j + 1) % P) == r.p) {
                const int64_t qn = (j + 1) / P;               // local index of the next panel (== q_first)
                GPX_TRY(trailing_update(r, j, qn, qn + 1));
                GPX_TRY(panel_factor_pack(r, j + 1, r.stage[sb ^ 1]));
                GPX_CUDA(cudaEventRecord(evP2[sb ^ 1], S));
                GPX_TRY(trailing_update(r, j, qn + 1, r.nloc));
            } else {
                if (j + 1 < r.nblk) GPX_CUDA(cudaEventRecord(evP2[sb ^ 1], S));  // keeps the event "fresh" on non-owners
                GPX_TRY(trailing_update(r, j, q_first, r.nloc));
            }
        }

        GPX_CUDA(cudaStreamSynchronize(Cs));
        for (int i = 0; i < 2; ++i) {
            cudaEventDestroy(evP2[i]);
            cudaEventDestroy(evR2[i]);
        }
    } else {
        auto on_H = [&](auto&& fn) -> int { h->stream = H; int rc_ = fn(); h->stream = S; return rc_; };
        if (r.p == 0) {
            GPX_TRY(on_H([&]() { return panel_factor_pack(r, 0, r.stage[0]); }));
            GPX_CUDA(cudaEventRecord(evPanel[0], H));
        }
        for (int64_t j = 0; j < r.nblk; ++j) {
            const int sb = (int)(j & 1);
            const int owner = (int)(j % P);
            const int64_t rows = r.npad - j * r.nb;
            const size_t count = (size_t)rows * r.nb + (size_t)r.tpb * GPX_T * GPX_T;
            // ---- communication stream
            if (owner == r.p) GPX_CUDA(cudaStreamWaitEvent(Cs, evPanel[j], 0));
            if (j >= 2) GPX_CUDA(cudaStreamWaitEvent(Cs, evRecv[j - 2], 0));   // stage[sb] free (trivially true on Cs itself)
            if (P > 1) GPX_NCCL(g_nccl.Broadcast(r.stage[sb], r.stage[sb], count, NCCL_F64, owner, (ncclComm_p)h->nccl_comm, Cs));
            GPX_TRY(panel_unpack(r, j, r.stage[sb], Cs));
            GPX_CUDA(cudaEventRecord(evRecv[j], Cs));
            // ---- chain stream
            GPX_CUDA(cudaStreamWaitEvent(H, evRecv[j], 0));
            if (j + 1 < r.nblk && (int)((j + 1) % P) == r.p) {
                const int64_t q1 = (j + 1) / P;
                GPX_TRY(on_H([&]() {
                    GPX_TRY(trailing_update(r, j, q1, q1 + 1));
                    return panel_factor_pack(r, j + 1, r.stage[sb ^ 1]);
                }));
                GPX_CUDA(cudaEventRecord(evPanel[j + 1], H));
            }
            if (j + 2 < r.nblk && (int)((j + 2) % P) == r.p) {
                const int64_t q2 = (j + 2) / P;
                if (j >= 1) GPX_CUDA(cudaStreamWaitEvent(H, evS[j - 1], 0));
                GPX_TRY(on_H([&]() { return trailing_update(r, j, q2, q2 + 1); }));
            }
            // ---- bulk stream.  With many ranks the owner of panel j+1 gives its chain the whole GPU first (the chain kernels
            // are throughput-bound while the trailing matrix is large, so sharing the SMs with the bulk update only delays the
            // broadcast every other rank waits for); with 1-2 ranks the concurrent schedule wins.
            GPX_CUDA(cudaStreamWaitEvent(S, evRecv[j], 0));
            if (P > 2 && j + 1 < r.nblk && (int)((j + 1) % P) == r.p) GPX_CUDA(cudaStreamWaitEvent(S, evPanel[j + 1], 0));
            GPX_TRY(trailing_update(r, j, first_local_block_after(r, j + 2), r.nloc));
            GPX_CUDA(cudaEventRecord(evS[j], S));
        }
        GPX_CUDA(cudaEventRecord(ev_start, H));
        GPX_CUDA(cudaStreamWaitEvent(S, ev_start, 0));
    }
    GPX_CUDA(cudaStreamSynchronize(Cs));
    GPX_CUDA(cudaStreamSynchronize(H));
    int info = 0;
    GPX_TRY(gpx_read_info(h, &info));
    if (P > 1) {
        // agree on the failure flag: max over ranks of (info > 0 ? info : 0)
        GPX_CUDA(cudaMemcpyAsync(h->d_info, &info, sizeof(int), cudaMemcpyHostToDevice, S));
        GPX_NCCL(g_nccl.AllReduce(h->d_info, h->d_info, 1, NCCL_I32, NCCL_MAX, (ncclComm_p)h->nccl_comm, S));
        GPX_TRY(gpx_read_info(h, &info));
    }
    for (int64_t j = 0; j < r.nblk; ++j) {
        cudaEventDestroy(evPanel[j]);
        cudaEventDestroy(evRecv[j]);
        cudaEventDestroy(evS[j]);
    }
    cudaEventDestroy(ev_start);
    if (info > 0) {
        gpx_set_error("gpx_mg_fit_grad: leading minor of order %d is not positive definite", info);
        return info;
    }
    if (!with_grad) {
        gpx_phase_mark(h, GPX_PH_SOLVE);
        GPX_TRY(solve_lml(r, y, alpha, out3));
        gpx_phase_mark(h, GPX_PH_END);
        return 0;
    }
    // diag(L) is needed for the log-determinant after the factor buffer is recycled for X
    double* diagL = r.stage[0];
    double* tmpv = r.stage[0] + r.npad;
    GPX_TRY(gpx_copy_strided(h, r.npad, r.Lfull, r.npad + 1, diagL, 1));
    gpx_phase_mark(h, GPX_PH_TRTRI);
    GPX_TRY(trtri_local(r));
    if (P > 1) {
        const size_t cnt = (size_t)r.npad * r.wloc;
        GPX_NCCL(g_nccl.AllGather(r.Xall + (size_t)r.p * cnt, r.Xall, cnt, NCCL_F64, (ncclComm_p)h->nccl_comm, S));
    }
    GPX_TRY(reorder_X(r));
    gpx_phase_mark(h, GPX_PH_SOLVE);
    GPX_TRY(solve_lml_from_inverse(r, y, alpha, tmpv, diagL, out3));
    gpx_phase_mark(h, GPX_PH_LAUUM);
    GPX_TRY(lauum_local(r));
    gpx_phase_mark(h, GPX_PH_GRAD);
    GPX_CUDA(cudaMemsetAsync(grad, 0, ntheta * sizeof(double), S));
    GPX_TRY(grad_local(r, kind, X, D, theta_host, ntheta, alpha, grad, h->d_theta));
    if (P > 1) GPX_NCCL(g_nccl.AllReduce(grad, grad, ntheta, NCCL_F64, NCCL_SUM, (ncclComm_p)h->nccl_comm, S));
    gpx_phase_mark(h, GPX_PH_END);
    return 0;
}

// Test helper: run the same per-rank routines for P *virtual* ranks inside one process on one GPU (broadcast /
// all-gather / all-reduce become device copies), phase by phase.  Validates the block-cyclic index maps without
// needing P GPUs.  ws_all = P * gpx_mg_workspace_elems doubles.  Outputs as gpx_mg_fit_grad (rank 0's copy).
extern "C" int gpx_mg_emulate_fit_grad(gpx_handle h, int P, int kind, const double* X, int64_t n, int D,
                                       const double* theta_host, int ntheta, double s, const double* y, int nb, double* ws_all,
                                       double* alpha, double* out3, double* grad) {
    GPX_ENTER(h);
    GPX_REQUIRE(P >= 1 && P <= 16, 2);
    cudaStream_t S = h->stream;
    std::vector<MgRank> R(P);
    const int64_t per = gpx_mg_workspace_elems(n, nb, P);
    for (int p = 0; p < P; ++p) init_rank(R[p], h, P, p, n, nb, ws_all + (size_t)p * per);
    GPX_CUDA(cudaMemsetAsync(h->d_info, 0, sizeof(int), S));
    for (int p = 0; p < P; ++p) {
        GPX_CUDA(cudaMemsetAsync(R[p].Lfull, 0, (size_t)R[p].npad * R[p].npad * sizeof(double), S));
        GPX_TRY(cov_local(R[p], kind, X, D, theta_host, ntheta, s));
    }
    const MgRank& r0 = R[0];
    for (int64_t j = 0; j < r0.nblk; ++j) {
        const int owner = (int)(j % P);
        const int64_t rows = r0.npad - j * r0.nb;
        const size_t count = (size_t)rows * r0.nb + (size_t)r0.tpb * GPX_T * GPX_T;
        GPX_TRY(panel_factor_pack(R[owner], j, R[owner].stage[0]));
        for (int p = 0; p < P; ++p) {
            if (p != owner)
                GPX_CUDA(cudaMemcpyAsync(R[p].stage[0], R[owner].stage[0], count * sizeof(double), cudaMemcpyDeviceToDevice, S));
            GPX_TRY(panel_unpack(R[p], j, R[p].stage[0], S));
            GPX_TRY(trailing_update(R[p], j, first_local_block_after(R[p], j), R[p].nloc));
        }
    }
    int info = 0;
    GPX_TRY(gpx_read_info(h, &info));
    if (info > 0) return info;
    GPX_TRY(solve_lml(R[0], y, alpha, out3));
    for (int p = 0; p < P; ++p) GPX_TRY(trtri_local(R[p]));
    const size_t cnt = (size_t)r0.npad * r0.wloc;
    for (int p = 0; p < P; ++p)
        for (int src = 0; src < P; ++src)
            if (src != p)
                GPX_CUDA(cudaMemcpyAsync(R[p].Xall + src * cnt, R[src].Xall + src * cnt, cnt * sizeof(double),
                                         cudaMemcpyDeviceToDevice, S));
    GPX_CUDA(cudaMemsetAsync(grad, 0, ntheta * sizeof(double), S));
    for (int p = 0; p < P; ++p) {
        GPX_TRY(reorder_X(R[p]));
        GPX_TRY(lauum_local(R[p]));
        GPX_TRY(grad_local(R[p], kind, X, D, theta_host, ntheta, alpha, grad, h->d_theta));  // sums over ranks = all-reduce
    }
    return 0;
}
