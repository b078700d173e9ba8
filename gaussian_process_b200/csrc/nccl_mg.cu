// Multi-GPU exact-GP fit + LML + gradient, distributed Cholesky of any matrix the covariance builder can form, and the
// distributed binary-Laplace Newton step (one process per GPU).
//
// Layout: 1-D block-cyclic over block columns of width nb (a P x 1 process grid of the 2-D block-cyclic scheme; on
// NVSwitch every rank sees every panel at full bandwidth, so the second grid dimension buys nothing -- SURVEY 8e
// "Cholesky").  Rank p owns global block columns j = q*P + p, side by side in a local row-major matrix Aloc[npad][wloc].
//   potrf : right-looking with GROUPED trailing updates.  Panels are nb wide (the broadcast / chain granularity) but the
//           bulk of the trailing matrix is updated once per group of G panels with K = G*nb (default 1024), where the
//           DMMA GEMM runs at its large-K rate (a K = 256 update runs ~20 % below it -- that, not communication, was the
//           scaling loss of the first version).  Inside a group, panel j updates only the local columns of its own group
//           (K = nb, "eager"); when the group is complete its G panels update the columns of the NEXT group first (so
//           the panel chain can go on) and the rest of the trailing matrix in G deferred chunks, one per step of the next
//           group, so every rank's compute stream stays busy while the next panels are factored and broadcast.  The owner
//           factors the diagonal block + TRSMs the panel (recursive DMMA kernels); NCCL broadcast + unpack into the
//           replicated factor Lfull run on a high-priority communication stream.
//   alpha : blocked TRSVs on the replicated factor (redundant on every rank, no communication) on a side stream,
//           concurrently with the inverse.
//   K^-1  : two prefix-structured recursive TRSMs on the rank's own block columns with the replicated factor:
//           L X = E (X = local columns of L^-1), then L^T Z = X restricted to the rows on/below each column block's
//           diagonal block (all a symmetric result needs).  N^3/(3P) flops each, NO communication (the first version
//           all-gathered L^-1: 34 GB at N = 65536, plus a replicated copy of it on every GPU).
//   grad  : fused trace kernel over the local block columns of K^-1, all-reduce of the ntheta partial sums.
// Memory per GPU: N^2 (replicated factor) + 2 N^2 / P doubles.
// NCCL is dlopen'ed (torch-bundled libnccl.so.2); with world == 1 no NCCL symbol is touched.  For tests the same
// per-rank routines can be driven for P virtual ranks inside one process (gpx_mg_emulate_fit_grad).
#include <dlfcn.h>
#include <stdlib.h>
#include <algorithm>
#include <vector>
#include "common.cuh"

namespace {

// ------------------------------------------------------------------------------------------------ NCCL (dlopen)
typedef struct { char internal[128]; } ncclUniqueId_t;
typedef void* ncclComm_p;
typedef int (*fn_GetUniqueId)(ncclUniqueId_t*);
typedef int (*fn_CommInitRank)(ncclComm_p*, int, ncclUniqueId_t, int);
typedef int (*fn_CommDestroy)(ncclComm_p);
typedef int (*fn_CommAbort)(ncclComm_p);
typedef int (*fn_Broadcast)(const void*, void*, size_t, int, int, ncclComm_p, cudaStream_t);
typedef int (*fn_AllGather)(const void*, void*, size_t, int, ncclComm_p, cudaStream_t);
typedef int (*fn_AllReduce)(const void*, void*, size_t, int, int, ncclComm_p, cudaStream_t);
typedef const char* (*fn_GetErrorString)(int);
struct NcclApi {
    void* lib = nullptr;
    fn_GetUniqueId GetUniqueId = nullptr;
    fn_CommInitRank CommInitRank = nullptr;
    fn_CommDestroy CommDestroy = nullptr;
    fn_CommAbort CommAbort = nullptr;
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
    LOAD(GetUniqueId) LOAD(CommInitRank) LOAD(CommDestroy) LOAD(CommAbort) LOAD(Broadcast) LOAD(AllGather) LOAD(AllReduce) LOAD(GetErrorString)
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
struct MgBuild {            // what matrix is factored: scale_i k(x_i, x_j; theta) scale_j + diag_add [i == j]
    int kind; const double* X; int64_t n; int D; const double* theta; int ntheta; double diag_add; const double* scale;
};

struct MgRank {
    gpx_ctx* h;         // handle whose stream the rank's kernels run on
    int P, p, snake;    // world, rank, block -> rank map (gpx_cyc_*)
    int64_t n, npad;    // true / padded size (npad % (nb*P) == 0; nb*2P with the snake map)
    int nb, tpb, G;     // block width, tiles per block, panels per group (bulk K = G*nb)
    int64_t nblk, nloc, wloc;   // global blocks, local blocks, local width (elements)
    double *Aloc, *Xloc, *Lfull, *dinv, *stage[2], *vec, *solve_ws, *trsm_tmp;
};

int g_group_k = -1;
int group_k() {   // K of the grouped bulk update
    if (g_group_k < 0) {
        const char* e = getenv("GPX_MG_GROUP_K");
        g_group_k = e ? atoi(e) : 1024;
        if (g_group_k < GPX_T) g_group_k = GPX_T;
    }
    return g_group_k;
}

// Panels factored by the substitution chain (gpx_panel_factor_sub) or by potrf + GEMM-based TRSM?  Measured at 8 GPUs: the
// substitution chain wins where the factorisation is chain-bound (C3, N = 16384: 41.1 -> 38.5 ms per Newton step) and is
// level / slightly behind where the panels are tall (C5, N = 65536: 1.162 vs 1.166 s) -> default by problem size and, inside a
// large factorisation, by panel height (the short panels of the chain-bound tail); $GPX_MG_PANEL_SUB forces it.
int panel_sub(int64_t npad, int64_t rows) {   // rows = height of this panel
    static int v = -2, tall = -1;
    if (v == -2) {
        const char* e = getenv("GPX_MG_PANEL_SUB");
        v = e ? atoi(e) : -1;
        const char* t = getenv("GPX_MG_PANEL_SUB_ROWS");   // panels at most this tall use the substitution chain
        tall = t ? atoi(t) : 24576;
    }
    return v >= 0 ? v : ((npad <= 32768 || rows <= tall) ? 1 : 0);
}

int g_snake = -1;
int layout_snake() {   // 1: boustrophedon block -> rank map (default), 0: plain block-cyclic
    if (g_snake < 0) {
        const char* e = getenv("GPX_MG_SNAKE");
        g_snake = e ? (atoi(e) != 0) : 1;
    }
    return g_snake;
}

__global__ void set_identity_blocks_kernel(double* X, int64_t ld, int64_t nloc, int nb, int P, int p, int snake) {
    // X[global(q)*nb + i][q*nb + i] = 1 for every local block q
    int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= nloc * nb) return;
    int64_t q = idx / nb, i = idx - q * nb;
    X[(gpx_cyc_global(q, P, p, snake) * nb + i) * ld + q * nb + i] = 1.0;
}

size_t stage_elems(int64_t npad, int nb) { return (size_t)npad * nb + (size_t)(nb / GPX_T) * GPX_T * GPX_T; }
size_t solve_ws_elems(int64_t npad) {   // block inverses (npad*bs) + their build scratch + the TRSV scratch
    const int64_t bs = gpx_block_size_for(npad);
    return (size_t)npad * bs + (size_t)npad * bs / 4 + GPX_T + bs + 64;
}
struct SolveWs { int bs; double *Dbig, *work, *tmp; bool use; };
SolveWs solve_ws_of(const MgRank& r) {
    SolveWs w;
    w.bs = gpx_block_size_for(r.npad);
    w.Dbig = r.solve_ws;
    w.work = w.Dbig + (size_t)r.npad * w.bs;
    w.tmp = w.work + (size_t)r.npad * w.bs / 4 + GPX_T;
    w.use = w.bs > GPX_T && r.npad >= 2 * w.bs;
    return w;
}
// explicit inverses of the bs x bs diagonal blocks of the replicated factor: shared by the TRSVs and the two TRSMs
int build_block_inverses(MgRank& r) {
    SolveWs w = solve_ws_of(r);
    if (!w.use) return 0;
    return gpx_block_inverses(r.h, r.Lfull, r.npad, r.npad, r.dinv, w.bs, w.Dbig, w.work);
}

// owner side: factor diagonal block j and TRSM the panel below it (in Aloc), then pack panel + leaf inverses
int panel_factor_pack(MgRank& r, int64_t j, double* stage) {
    gpx_ctx* h = r.h;
    const int64_t q = gpx_cyc_local(j, r.P, r.snake), r0 = j * r.nb, rows = r.npad - r0;
    double* diag = r.Aloc + r0 * r.wloc + q * r.nb;
    double* dinvj = r.dinv + j * r.tpb * GPX_T * GPX_T;
    if (panel_sub(r.npad, rows)) {
        GPX_TRY(gpx_panel_factor_sub(h, diag, rows, r.wloc, r.nb, dinvj, (int)r0));   // substitution chain (see potrf.cu)
    } else {
        GPX_TRY(gpx_potrf_block(h, diag, r.nb, r.wloc, dinvj, (int)r0));
        if (rows > r.nb)
            GPX_TRY(gpx_trsm_right_lt_block(h, diag + (int64_t)r.nb * r.wloc, rows - r.nb, r.wloc, diag, r.nb, r.wloc, dinvj));
    }
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

// update of the local blocks q in [q_lo, q_hi) with the panels [j_lo, j_lo + np) (K = np*nb) of the replicated factor:
// Aloc[k:, q] -= L[k:, j_lo .. j_lo+np) L[block(q), j_lo .. j_lo+np)^T, lower tiles only
int trailing_update(MgRank& r, int64_t j_lo, int np_panels, int64_t q_lo, int64_t q_hi) {
    if (q_hi > r.nloc) q_hi = r.nloc;
    if (q_lo >= q_hi || np_panels <= 0) return 0;
    const int64_t gk0 = gpx_cyc_global(q_lo, r.P, r.p, r.snake);   // global block of the first updated local column
    const int64_t r_start = gk0 * r.nb;
    GemmArgs a{};
    a.batch = 1;
    a.alpha = -1.0; a.beta = 1.0;
    a.A = r.Lfull + r_start * r.npad + j_lo * r.nb; a.lda = r.npad; a.a_kmajor = 1;
    a.B = a.A; a.ldb = r.npad; a.b_kmajor = 1;
    a.C = r.Aloc + r_start * r.wloc + q_lo * r.nb; a.ldc = r.wloc;
    a.M = (int)(r.npad - r_start); a.N = (int)((q_hi - q_lo) * r.nb); a.K = np_panels * r.nb;
    a.lower_only = 1;
    a.cyc_P = r.P; a.cyc_p = r.p; a.cyc_snake = r.snake; a.cyc_tpb = r.tpb; a.cyc_q0 = (int)q_lo; a.cyc_row_base = (int)r_start; a.cyc_b_rows = 1;
    return gpx_gemm_launch(r.h, a);
}

int64_t first_local_block_after(const MgRank& r, int64_t j) {  // smallest local q whose global block is > j
    return gpx_cyc_count_below(j + 1, r.P, r.p, r.snake);
}
int owner_of(const MgRank& r, int64_t j) { return gpx_cyc_owner(j, r.P, r.snake); }

// Deferred bulk update B(g): the panels of group g applied to the local columns beyond group g+1, cut into `nchunks`
// column ranges of roughly equal area; chunk c is issued at step c of the next group.
struct BulkPlan {
    int64_t j_lo = 0; int np_panels = 0; int nchunks = 0;
    std::vector<int64_t> q_cut;      // nchunks + 1 boundaries
};
BulkPlan plan_bulk(const MgRank& r, int64_t g, int nchunks) {
    BulkPlan bp;
    const int64_t j0 = g * r.G;
    bp.j_lo = j0;
    bp.np_panels = (int)std::min<int64_t>(r.G, r.nblk - j0);
    bp.nchunks = nchunks;
    const int64_t q_lo = first_local_block_after(r, j0 + 2 * (int64_t)r.G - 1);   // beyond group g+1
    bp.q_cut.assign(nchunks + 1, r.nloc);
    bp.q_cut[0] = std::min(q_lo, r.nloc);
    if (nchunks <= 0 || q_lo >= r.nloc) return bp;
    // area of column q ~ rows below its diagonal block
    double total = 0.0;
    for (int64_t q = q_lo; q < r.nloc; ++q) total += (double)(r.nblk - gpx_cyc_global(q, r.P, r.p, r.snake));
    double acc = 0.0;
    int c = 1;
    for (int64_t q = q_lo; q < r.nloc && c < nchunks; ++q) {
        acc += (double)(r.nblk - gpx_cyc_global(q, r.P, r.p, r.snake));
        if (acc >= total * c / nchunks) bp.q_cut[c++] = q + 1;
    }
    for (; c < nchunks; ++c) bp.q_cut[c] = r.nloc;
    return bp;
}
int bulk_chunk(MgRank& r, const BulkPlan& bp, int c) {
    if (c >= bp.nchunks) return 0;
    return trailing_update(r, bp.j_lo, bp.np_panels, bp.q_cut[c], bp.q_cut[c + 1]);
}

// One step of the factorisation on one rank, AFTER panel j (group g = j / G, position w) is in Lfull:
//   eager updates of the local columns of group g beyond j (and, if this rank owns
//   panel j+1 of the same group, its factorisation right after its own update)  ->  at the end of the group: A(g), the
//   update of the NEXT group's local columns with the whole group (K = G*nb), then the first panel of the next group.
// `on_panel(jn, stage)` is called right after panel jn has been factored and packed by this rank.
template <class OnPanel>
int rank_step(MgRank& r, int64_t j, double* next_stage, OnPanel&& on_panel) {
    const int64_t g = j / r.G, j0 = g * r.G;
    const int gw = (int)std::min<int64_t>(r.G, r.nblk - j0);
    const int64_t jl = j0 + gw - 1;                       // last panel of the group
    const int64_t qa = first_local_block_after(r, j), qe = first_local_block_after(r, jl);
    if (j < jl && owner_of(r, j + 1) == r.p) {             // the next panel of this group is mine: it is local block qa
        GPX_TRY(trailing_update(r, j, 1, qa, qa + 1));
        GPX_TRY(panel_factor_pack(r, j + 1, next_stage));
        GPX_TRY(on_panel(j + 1, next_stage));
        GPX_TRY(trailing_update(r, j, 1, qa + 1, qe));
    } else {
        GPX_TRY(trailing_update(r, j, 1, qa, qe));
    }
    if (j == jl && jl + 1 < r.nblk) {                     // group complete: A(g) on the next group's columns
        const int64_t qn = first_local_block_after(r, jl + r.G);
        GPX_TRY(trailing_update(r, j0, gw, qe, qn));
        if (owner_of(r, jl + 1) == r.p) {
            GPX_TRY(panel_factor_pack(r, jl + 1, next_stage));
            GPX_TRY(on_panel(jl + 1, next_stage));
        }
    }
    return 0;
}

int build_local(MgRank& r, const MgBuild& b) {
    for (int64_t q = 0; q < r.nloc; ++q) {
        const int64_t j = gpx_cyc_global(q, r.P, r.p, r.snake);
        GPX_TRY(gpx_cov_build_block(r.h, b.kind, b.X, b.n, b.D, b.theta, b.ntheta, b.diag_add, GPX_COV_SAME_X | GPX_COV_LOWER,
                                    r.Aloc + q * r.nb, r.npad, r.nb, r.wloc, 0, (int)(j * r.nb), b.scale));
    }
    return 0;
}

__global__ void accumulate_kernel(int n, const double* __restrict__ x, double* __restrict__ acc) {
    int i = threadIdx.x;
    if (i < n) acc[i] += x[i];
}

int grad_local(MgRank& r, int kind, const double* X, int D, const double* theta, int ntheta, const double* alpha,
               double* grad_acc /* device, ntheta, zeroed by caller */, double* tmp /* device, ntheta */) {
    for (int64_t q = 0; q < r.nloc; ++q) {
        const int64_t j = gpx_cyc_global(q, r.P, r.p, r.snake), r0 = j * r.nb;
        GPX_TRY(gpx_lml_grad_block(r.h, kind, X, r.n, D, theta, ntheta, r.Xloc + r0 * r.wloc + q * r.nb, r.wloc, alpha, tmp,
                                   r.npad - r0, r.nb, (int)r0, (int)r0));
        accumulate_kernel<<<1, 32, 0, r.h->stream>>>(ntheta, tmp, grad_acc);
        GPX_CHECK_LAUNCH(r.h);
    }
    return 0;
}

// local block columns of K^-1 (rows on/below each block's diagonal block) from the replicated factor, in Xloc
int inverse_local(MgRank& r) {
    GPX_CUDA(cudaMemsetAsync(r.Xloc, 0, (size_t)r.npad * r.wloc * sizeof(double), r.h->stream));
    const int64_t cnt = r.nloc * r.nb;
    set_identity_blocks_kernel<<<(unsigned)((cnt + 255) / 256), 256, 0, r.h->stream>>>(r.Xloc, r.wloc, r.nloc, r.nb, r.P, r.p, r.snake);
    GPX_CHECK_LAUNCH(r.h);
    SolveWs w = solve_ws_of(r);     // block inverses must have been built (build_block_inverses)
    const double* D = w.use ? w.Dbig : nullptr;
    gpx_phase_mark(r.h, GPX_PH_TRTRI);
    GPX_TRY(gpx_trsm_left_prefix_block(r.h, r.Lfull, r.npad, r.npad, r.dinv, r.Xloc, r.wloc, r.P, r.p, r.nb, r.snake, D, w.bs,
                                       r.trsm_tmp));                                                                   // X = L^-1 E
    gpx_phase_mark(r.h, GPX_PH_LAUUM);
    return gpx_trsm_left_prefix_trans_block(r.h, r.Lfull, r.npad, r.npad, r.dinv, r.Xloc, r.wloc, r.P, r.p, r.nb, r.snake, D,
                                            w.bs, r.trsm_tmp);                                                         // Z = L^-T X
}

// alpha = L^-T L^-1 y and {lml, y.alpha, sum log diag} on the replicated factor (short-chain blocked TRSVs)
int solve_lml(MgRank& r, const double* y, double* alpha, double* out3) {
    gpx_ctx* h = r.h;
    GPX_CUDA(cudaMemsetAsync(alpha, 0, r.npad * sizeof(double), h->stream));
    GPX_CUDA(cudaMemcpyAsync(alpha, y, r.n * sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
    SolveWs w = solve_ws_of(r);     // block inverses must have been built (build_block_inverses)
    if (w.use) {
        GPX_TRY(gpx_trsv_big(h, r.Lfull, r.npad, r.npad, w.Dbig, w.bs, 0, alpha, w.tmp));
        GPX_TRY(gpx_trsv_big(h, r.Lfull, r.npad, r.npad, w.Dbig, w.bs, 1, alpha, w.tmp));
    } else {
        GPX_TRY(gpx_trsv(h, r.Lfull, r.npad, r.npad, r.dinv, 0, alpha));
        GPX_TRY(gpx_trsv(h, r.Lfull, r.npad, r.npad, r.dinv, 1, alpha));
    }
    return gpx_lml(h, r.Lfull, r.n, r.npad, y, alpha, out3);
}

int init_rank(MgRank& r, gpx_ctx* h, int P, int p, int64_t n, int nb, double* ws) {
    r.h = h; r.P = P; r.p = p; r.n = n; r.nb = nb; r.tpb = nb / GPX_T;
    r.G = std::max(1, group_k() / nb);
    r.snake = layout_snake();
    r.npad = gpx_mg_padded_dim(n, nb, P);
    r.nblk = r.npad / nb; r.nloc = r.nblk / P; r.wloc = r.nloc * nb;
    double* w = ws;
    r.Aloc = w; w += (size_t)r.npad * r.wloc;
    r.Xloc = w; w += (size_t)r.npad * r.wloc;
    r.Lfull = w; w += (size_t)r.npad * r.npad;
    r.dinv = w; w += (size_t)(r.npad / GPX_T) * GPX_T * GPX_T;
    r.stage[0] = w; w += stage_elems(r.npad, nb);
    r.stage[1] = w; w += stage_elems(r.npad, nb);
    r.vec = w; w += 4 * (size_t)r.npad;
    r.solve_ws = w; w += solve_ws_elems(r.npad);
    r.trsm_tmp = w; w += (size_t)gpx_block_size_for(r.npad) * r.wloc;
    return 0;
}

// ---- the factorisation on a real rank: compute stream S, communication stream Cs (NCCL broadcast + unpack)
int factor_rank(MgRank& r, const MgBuild& b, int* info_out) {
    gpx_ctx* h = r.h;
    const int P = r.P;
    cudaStream_t S = h->stream, Cs = h->aux_stream;
    GpxEventSet es;
    GPX_TRY(es.create(5));
    cudaEvent_t evPanel[2] = {es.ev[0], es.ev[1]}, evRecv[2] = {es.ev[2], es.ev[3]}, ev_start = es.ev[4];
    int rc = 0;
    auto body = [&]() -> int {
        GPX_CUDA(cudaMemsetAsync(h->d_info, 0, sizeof(int), S));
        GPX_CUDA(cudaMemsetAsync(r.Lfull, 0, (size_t)r.npad * r.npad * sizeof(double), S));
        gpx_phase_mark(h, GPX_PH_COV);
        GPX_TRY(build_local(r, b));
        gpx_phase_mark(h, GPX_PH_POTRF);
        GPX_CUDA(cudaEventRecord(ev_start, S));
        GPX_CUDA(cudaStreamWaitEvent(Cs, ev_start, 0));   // Lfull memset / matrix build precede the first unpack
        auto on_panel = [&](int64_t jn, double*) -> int {
            GPX_CUDA(cudaEventRecord(evPanel[jn & 1], S));
            return 0;
        };
        if (r.p == 0) {
            GPX_TRY(panel_factor_pack(r, 0, r.stage[0]));
            GPX_TRY(on_panel(0, r.stage[0]));
        }
        BulkPlan prev, cur;
        bool have_prev = false;
        for (int64_t j = 0; j < r.nblk; ++j) {
            const int sb = (int)(j & 1);
            const int owner = owner_of(r, j);
            const int64_t rows = r.npad - j * r.nb;
            const size_t count = (size_t)rows * r.nb + (size_t)r.tpb * GPX_T * GPX_T;
            const int64_t g = j / r.G, j0 = g * r.G;
            const int gw = (int)std::min<int64_t>(r.G, r.nblk - j0);
            // ---- communication stream: broadcast panel j as soon as its owner has packed it, unpack into the factor.
            // Non-owners gate their side of the broadcast on the point of their OWN compute stream at which the owner
            // (running in lock-step) is about to finish the panel: a receive kernel launched earlier would spin on
            // ~20 SMs' worth of registers for a whole step and slow the bulk GEMM by ~14 % (measured at 2 GPUs).
            GPX_CUDA(cudaStreamWaitEvent(Cs, evPanel[sb], 0));
            if (P > 1) {
                // a one-element broadcast first: its single-CTA kernel is what spins until the owner is ready, so the
                // many-channel kernel of the panel itself starts when the data is there and never idles on the SMs
                GPX_NCCL(g_nccl.Broadcast(r.vec, r.vec, 1, NCCL_F64, owner, (ncclComm_p)h->nccl_comm, Cs));
                GPX_NCCL(g_nccl.Broadcast(r.stage[sb], r.stage[sb], count, NCCL_F64, owner, (ncclComm_p)h->nccl_comm, Cs));
            }
            GPX_TRY(panel_unpack(r, j, r.stage[sb], Cs));
            GPX_CUDA(cudaEventRecord(evRecv[sb], Cs));
            // ---- compute stream: the deferred chunk does not need panel j, so it is issued BEFORE the wait
            if (have_prev) GPX_TRY(bulk_chunk(r, prev, (int)(j - j0)));
            GPX_CUDA(cudaStreamWaitEvent(S, evRecv[sb], 0));
            GPX_TRY(rank_step(r, j, r.stage[sb ^ 1], on_panel));
            if (j + 1 < r.nblk && owner_of(r, j + 1) != r.p) GPX_CUDA(cudaEventRecord(evPanel[sb ^ 1], S));   // the gate above
            if (j == j0 + gw - 1) {   // group complete: its bulk update is deferred into the steps of the next group
                const int64_t jn0 = j0 + gw;
                const int gw_next = jn0 < r.nblk ? (int)std::min<int64_t>(r.G, r.nblk - jn0) : 0;
                prev = plan_bulk(r, g, gw_next);
                have_prev = gw_next > 0;
            }
        }
        return 0;
    };
    rc = body();
    // always join the communication stream back into S before anyone may recycle the buffers (also on error)
    if (cudaEventRecord(ev_start, Cs) == cudaSuccess) cudaStreamWaitEvent(S, ev_start, 0);
    if (rc != 0) {
        // A rank-local failure (launch / allocation / NCCL error) in the middle of the panel loop would leave the other
        // ranks blocked in their next broadcast: abort the communicator so their collectives fail instead of hanging.
        if (P > 1 && h->nccl_comm && g_nccl.CommAbort) {
            g_nccl.CommAbort((ncclComm_p)h->nccl_comm);
            h->nccl_comm = nullptr;
        }
        *info_out = 0;
        return rc;
    }
    int info = 0;
    GPX_TRY(gpx_read_info(h, &info));
    if (P > 1) {
        // agree on the pivot status: max over ranks of (info > 0 ? info : 0)
        GPX_CUDA(cudaMemcpyAsync(h->d_info, &info, sizeof(int), cudaMemcpyHostToDevice, S));
        GPX_NCCL(g_nccl.AllReduce(h->d_info, h->d_info, 1, NCCL_I32, NCCL_MAX, (ncclComm_p)h->nccl_comm, S));
        GPX_TRY(gpx_read_info(h, &info));
    }
    cudaMemsetAsync(h->d_info, 0, sizeof(int), S);
    *info_out = info;
    return rc;
}

}  // namespace

// K of the grouped trailing update (panels per group = K / nb, at least 1); default 1024 or $GPX_MG_GROUP_K
extern "C" int gpx_mg_set_group_k(int k) {
    GPX_REQUIRE(k >= GPX_T && k % GPX_T == 0, 1);
    g_group_k = k;
    return 0;
}

extern "C" int64_t gpx_mg_padded_dim(int64_t n, int nb, int world) {
    const int64_t unit = (int64_t)nb * world * (layout_snake() ? 2 : 1);   // every rank owns the same number of blocks
    return ((n + unit - 1) / unit) * unit;
}

// the map itself (host-side queries; no GPU needed): owner of global block j, global block of local block q on rank p,
// number of rank p's blocks with global index < j
extern "C" int gpx_mg_block_owner(int64_t j, int world) { return gpx_cyc_owner(j, world, layout_snake()); }
extern "C" int64_t gpx_mg_block_global(int64_t q, int world, int rank) { return gpx_cyc_global(q, world, rank, layout_snake()); }
extern "C" int64_t gpx_mg_blocks_below(int64_t j, int world, int rank) { return gpx_cyc_count_below(j, world, rank, layout_snake()); }

// block -> rank map of the multi-GPU layout: 1 = boustrophedon (default, balances the triangular work), 0 = plain cyclic
extern "C" int gpx_mg_set_layout(int snake) {
    g_snake = snake ? 1 : 0;
    return 0;
}

// doubles of device workspace one rank needs for the gpx_mg_* calls
extern "C" int64_t gpx_mg_workspace_elems(int64_t n, int nb, int world) {
    const int64_t npad = gpx_mg_padded_dim(n, nb, world);
    const int64_t wloc = npad / world;
    return 2 * npad * wloc + npad * npad + (npad / GPX_T) * GPX_T * GPX_T + 2 * (int64_t)stage_elems(npad, nb) + 4 * npad +
           (int64_t)solve_ws_elems(npad) + (int64_t)gpx_block_size_for(npad) * wloc;
}

// where the pieces live inside the workspace (element offsets): out6 = {Aloc, Xloc (local block columns of K^-1 after a
// fit with gradient), Lfull (replicated factor, npad x npad row-major, clean lower triangle), dinv (leaf inverses),
// npad, wloc}
extern "C" int gpx_mg_workspace_layout(int64_t n, int nb, int world, int64_t* out6) {
    GPX_REQUIRE(out6 != nullptr, 4);
    const int64_t npad = gpx_mg_padded_dim(n, nb, world);
    const int64_t wloc = npad / world;
    out6[0] = 0;
    out6[1] = npad * wloc;
    out6[2] = 2 * npad * wloc;
    out6[3] = 2 * npad * wloc + npad * npad;
    out6[4] = npad;
    out6[5] = wloc;
    return 0;
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

// sum over the ranks of the handle's communicator, in place, on the handle's stream (no-op for a single rank)
int gpx_nccl_allreduce_sum(gpx_ctx* h, double* buf, size_t count) {
    if (h->world <= 1 || h->nccl_comm == nullptr) return 0;
    GPX_NCCL(g_nccl.AllReduce(buf, buf, count, NCCL_F64, NCCL_SUM, (ncclComm_p)h->nccl_comm, h->stream));
    return 0;
}

// Distributed Cholesky of  M = diag(scale) k(X, X; theta) diag(scale) + diag_add I  (scale may be NULL): on return the
// workspace holds the replicated factor (gpx_mg_workspace_layout) on every rank.  This is the factorisation of
// gpx_mg_fit_grad (scale = NULL, diag_add = s) and of the Laplace matrix B = I + W^1/2 K W^1/2 (scale = W^1/2, diag_add = 1;
// GP_binary_classification.py:107).  Returns > 0 (first bad pivot) on every rank if M is not positive definite.
extern "C" int gpx_mg_factor(gpx_handle h, int kind, const double* X, int64_t n, int D, const double* theta_host, int ntheta,
                             double diag_add, const double* scale, int nb, double* ws) {
    GPX_ENTER(h);
    GPX_REQUIRE(nb >= GPX_T && nb % GPX_T == 0, 10);
    GPX_REQUIRE(h->world == 1 || h->nccl_comm != nullptr, 1);
    MgRank r;
    init_rank(r, h, h->world, h->rank, n, nb, ws);
    MgBuild b{kind, X, n, D, theta_host, ntheta, diag_add, scale};
    int info = 0;
    GPX_TRY(factor_rank(r, b, &info));
    if (info > 0) gpx_set_error("gpx_mg_factor: leading minor of order %d is not positive definite", info);
    return info;
}

// x <- (L L^T)^-1 x on the replicated factor left in `ws` by gpx_mg_factor / gpx_mg_fit_grad(with_grad = 0); x: npad doubles
// (zero padded).  Redundant on every rank, no communication.
extern "C" int gpx_mg_potrs_vec(gpx_handle h, int64_t n, int nb, double* ws, double* x) {
    GPX_ENTER(h);
    MgRank r;
    init_rank(r, h, h->world, h->rank, n, nb, ws);
    SolveWs w = solve_ws_of(r);
    if (w.use) {
        GPX_TRY(build_block_inverses(r));
        GPX_TRY(gpx_trsv_big(h, r.Lfull, r.npad, r.npad, w.Dbig, w.bs, 0, x, w.tmp));
        return gpx_trsv_big(h, r.Lfull, r.npad, r.npad, w.Dbig, w.bs, 1, x, w.tmp);
    }
    GPX_TRY(gpx_trsv(h, r.Lfull, r.npad, r.npad, r.dinv, 0, x));
    return gpx_trsv(h, r.Lfull, r.npad, r.npad, r.dinv, 1, x);
}

// One rank's part of the distributed fit + LML + gradient.  `ws` = gpx_mg_workspace_elems doubles of device memory.
// out3 (device) = {lml, y.alpha, sum log diag}; grad (device) = ntheta doubles (already all-reduced).
extern "C" int gpx_mg_fit_grad(gpx_handle h, int kind, const double* X, int64_t n, int D, const double* theta_host, int ntheta,
                               double s, const double* y, int nb, double* ws, double* alpha, double* out3, double* grad,
                               int with_grad) {
    GPX_ENTER(h);
    GPX_REQUIRE(nb >= GPX_T && nb % GPX_T == 0, 10);
    const int P = h->world;
    GPX_REQUIRE(P == 1 || h->nccl_comm != nullptr, 1);
    MgRank r;
    init_rank(r, h, P, h->rank, n, nb, ws);
    MgBuild b{kind, X, n, D, theta_host, ntheta, s, nullptr};
    int info = 0;
    GPX_TRY(factor_rank(r, b, &info));
    if (info > 0) {
        gpx_set_error("gpx_mg_fit_grad: leading minor of order %d is not positive definite", info);
        return info;
    }
    cudaStream_t S = h->stream, H = h->aux2_stream;
    GPX_TRY(build_block_inverses(r));
    if (!with_grad) {
        gpx_phase_mark(h, GPX_PH_SOLVE);
        GPX_TRY(solve_lml(r, y, alpha, out3));
        gpx_phase_mark(h, GPX_PH_END);
        return 0;
    }
    // alpha / LML on a side stream (HBM-bound TRSVs on the replicated factor) while S computes the inverse (DMMA-bound)
    GpxEventSet es;
    GPX_TRY(es.create(2));
    GPX_CUDA(cudaEventRecord(es.ev[0], S));
    GPX_CUDA(cudaStreamWaitEvent(H, es.ev[0], 0));
    h->stream = H;
    int rc = solve_lml(r, y, alpha, out3);
    h->stream = S;
    if (rc == 0) rc = inverse_local(r);
    // join H before anything else (also on error: alpha / out3 / the factor must not be recycled under the side stream)
    if (cudaEventRecord(es.ev[1], H) == cudaSuccess) cudaStreamWaitEvent(S, es.ev[1], 0);
    GPX_TRY(rc);
    gpx_phase_mark(h, GPX_PH_GRAD);
    GPX_CUDA(cudaMemsetAsync(grad, 0, ntheta * sizeof(double), S));
    GPX_TRY(grad_local(r, kind, X, D, theta_host, ntheta, alpha, grad, h->d_theta));
    GPX_TRY(gpx_nccl_allreduce_sum(h, grad, ntheta));
    gpx_phase_mark(h, GPX_PH_END);
    return 0;
}

// Distributed binary-Laplace Newton iteration (GP_binary_classification.py:104-111 with W and the gradient at the current
// f): B = I + W^1/2 K W^1/2 is built block-cyclically straight from X (never as a whole on one GPU) and factored by the
// distributed Cholesky; the O(N^2) mat-vecs with the replicated K (np x np, leading dimension ld) and the solves on the
// replicated factor are done redundantly on every rank (no communication besides the panel broadcasts).
// vws: 8*npad doubles (ws[0..npad) = gradient, [npad..2npad) = W, [2npad..3npad) = W^1/2 on return); f, f_new, y: npad.
extern "C" int gpx_mg_laplace_binary_step(gpx_handle h, int kind, const double* X, int64_t n, int D, const double* theta_host,
                                          int ntheta, const double* K, int64_t ld, const double* y, const double* f, int nb,
                                          double* ws, double* vws, double* f_new, double* err_dev) {
    GPX_ENTER(h);
    GPX_REQUIRE(nb >= GPX_T && nb % GPX_T == 0, 12);
    GPX_REQUIRE(f_new != f, 15);
    MgRank r;
    init_rank(r, h, h->world, h->rank, n, nb, ws);
    const int64_t np_ = r.npad;
    double *g = vws, *w = vws + np_, *sw = vws + 2 * np_, *b = vws + 3 * np_, *t = vws + 4 * np_, *a = vws + 5 * np_, *d = vws + 6 * np_;
    GPX_CUDA(cudaMemsetAsync(vws, 0, 8 * np_ * sizeof(double), h->stream));
    GPX_TRY(gpx_logistic_terms(h, 1, n, y, f, g, w, sw));
    MgBuild bd{kind, X, n, D, theta_host, ntheta, 1.0, sw};
    int info = 0;
    GPX_TRY(factor_rank(r, bd, &info));
    if (info > 0) {
        gpx_set_error("gpx_mg_laplace_binary_step: B is not positive definite (leading minor %d)", info);
        return info;
    }
    GPX_TRY(gpx_vec_op(h, 5, n, 0.0, w, f, g, b));
    GPX_TRY(gpx_gemv(h, 0, n, n, 1.0, K, ld, b, 0.0, t));
    GPX_TRY(gpx_vec_op(h, 2, n, 0.0, sw, t, nullptr, t));
    GPX_TRY(gpx_mg_potrs_vec(h, n, nb, ws, t));
    GPX_TRY(gpx_vec_op(h, 3, n, 0.0, b, sw, t, a));
    GPX_TRY(gpx_gemv(h, 0, n, n, 1.0, K, ld, a, 0.0, f_new));
    GPX_TRY(gpx_vec_op(h, 6, n, 0.0, f_new, f, nullptr, d));
    GPX_TRY(gpx_dot(h, n, d, d, err_dev));
    return gpx_vec_op(h, 10, 1, 0.0, err_dev, nullptr, nullptr, err_dev);
}

// Test helper: run the same per-rank routines for P *virtual* ranks inside one process on one GPU (broadcasts become
// device copies), step by step in the order of the real driver.  Validates the block-cyclic index maps, the grouped /
// deferred trailing updates and the prefix-structured solves without needing P GPUs.
// ws_all = P * gpx_mg_workspace_elems doubles.  Outputs as gpx_mg_fit_grad (rank 0's copy).
extern "C" int gpx_mg_emulate_fit_grad(gpx_handle h, int P, int kind, const double* X, int64_t n, int D,
                                       const double* theta_host, int ntheta, double s, const double* y, int nb, double* ws_all,
                                       double* alpha, double* out3, double* grad) {
    GPX_ENTER(h);
    GPX_REQUIRE(P >= 1 && P <= 16, 2);
    cudaStream_t S = h->stream;
    std::vector<MgRank> R(P);
    const int64_t per = gpx_mg_workspace_elems(n, nb, P);
    for (int p = 0; p < P; ++p) init_rank(R[p], h, P, p, n, nb, ws_all + (size_t)p * per);
    MgBuild b{kind, X, n, D, theta_host, ntheta, s, nullptr};
    GPX_CUDA(cudaMemsetAsync(h->d_info, 0, sizeof(int), S));
    for (int p = 0; p < P; ++p) {
        GPX_CUDA(cudaMemsetAsync(R[p].Lfull, 0, (size_t)R[p].npad * R[p].npad * sizeof(double), S));
        GPX_TRY(build_local(R[p], b));
    }
    const MgRank& r0 = R[0];
    auto noop = [](int64_t, double*) -> int { return 0; };
    GPX_TRY(panel_factor_pack(R[0], 0, R[0].stage[0]));
    std::vector<BulkPlan> prev(P);
    bool have_prev = false;
    for (int64_t j = 0; j < r0.nblk; ++j) {
        const int sb = (int)(j & 1);
        const int owner = owner_of(r0, j);
        const int64_t rows = r0.npad - j * r0.nb;
        const size_t count = (size_t)rows * r0.nb + (size_t)r0.tpb * GPX_T * GPX_T;
        const int64_t g = j / r0.G, j0 = g * r0.G;
        const int gw = (int)std::min<int64_t>(r0.G, r0.nblk - j0);
        for (int p = 0; p < P; ++p) {
            if (p != owner)
                GPX_CUDA(cudaMemcpyAsync(R[p].stage[sb], R[owner].stage[sb], count * sizeof(double), cudaMemcpyDeviceToDevice, S));
            GPX_TRY(panel_unpack(R[p], j, R[p].stage[sb], S));
        }
        for (int p = 0; p < P; ++p) {
            if (have_prev) GPX_TRY(bulk_chunk(R[p], prev[p], (int)(j - j0)));
            GPX_TRY(rank_step(R[p], j, R[p].stage[sb ^ 1], noop));
        }
        if (j == j0 + gw - 1) {
            const int64_t jn0 = j0 + gw;
            const int gw_next = jn0 < r0.nblk ? (int)std::min<int64_t>(r0.G, r0.nblk - jn0) : 0;
            for (int p = 0; p < P; ++p) prev[p] = plan_bulk(R[p], g, gw_next);
            have_prev = gw_next > 0;
        }
    }
    int info = 0;
    GPX_TRY(gpx_read_info(h, &info));
    if (info > 0) return info;
    GPX_TRY(build_block_inverses(R[0]));
    GPX_TRY(solve_lml(R[0], y, alpha, out3));
    GPX_CUDA(cudaMemsetAsync(grad, 0, ntheta * sizeof(double), S));
    for (int p = 0; p < P; ++p) {
        if (p > 0) GPX_TRY(build_block_inverses(R[p]));
        GPX_TRY(inverse_local(R[p]));
        GPX_TRY(grad_local(R[p], kind, X, D, theta_host, ntheta, alpha, grad, h->d_theta));  // sums over ranks = all-reduce
    }
    return 0;
}
