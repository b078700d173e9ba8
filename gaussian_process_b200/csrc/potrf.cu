// Blocked FP64 Cholesky, triangular solves and inverse for sm_100a.
//
//   potrf  : recursive right-looking lower Cholesky.  Leaves are 128x128 diagonal blocks factored by a
//            single-CTA shared-memory kernel that also emits the inverse of the leaf (used by every
//            triangular solve as a GEMM operand); everything else is the DMMA GEMM of gemm.cu
//            (TRSM against the leaf inverses, SYRK/GEMM trailing updates with k = half the block).
//   trsm   : recursive, left side, lower, N or T.
//   trtri  : recursive in-place inverse of L (two triangular GEMMs per level, batched over the
//            independent sub-problems of that level).
//   lauum  : out = Linv^T Linv (lower) in one triangular-aware launch.
// Replaces np.linalg.cholesky / np.linalg.solve(L, .) / np.linalg.inv(L) / np.dot(inv(L.T), inv(L))
// (SURVEY.md 8a rows A4, A5).
#include "common.cuh"

namespace {

constexpr int LT = 128;          // leaf size
constexpr int LLD = LT + 1;      // padded smem leading dimension
constexpr int LEAF_THREADS = 512;
constexpr int LEAF_SMEM = (LT * LLD + LT) * (int)sizeof(double);

// One CTA factors one 128x128 diagonal block in shared memory.
//  S (lower+diag) <- L ; S (strict upper) <- (L^-1)^T ; dinv_diag <- 1/L_ii
//  A tile <- L (upper zeroed) ; Dinv tile <- L^-1 (upper zeroed)
// info: first failing pivot (global 1-based index) is recorded once.
__global__ void __launch_bounds__(LEAF_THREADS, 1)
potrf_leaf_kernel(double* __restrict__ A, int64_t lda, double* __restrict__ Dinv, int* info, int global_off,
                  int64_t strideA, int64_t strideD) {
    extern __shared__ double sm[];
    double* S = sm;
    double* dd = sm + LT * LLD;
    A += (int64_t)blockIdx.x * strideA;
    Dinv += (int64_t)blockIdx.x * strideD;
    global_off += blockIdx.x * LT;
    const int tid = threadIdx.x;
    // load lower triangle (incl. diagonal)
    for (int idx = tid; idx < LT * LT; idx += LEAF_THREADS) {
        int i = idx >> 7, j = idx & 127;
        S[i * LLD + j] = (j <= i) ? A[(int64_t)i * lda + j] : 0.0;
    }
    __syncthreads();
    // right-looking, un-normalised columns: after step j, S[i][k] -= S[i][j] S[k][j] / S[j][j]
    const int tx = tid & 31, ty = tid >> 5;  // 32 x 16
    bool failed = false;
    for (int j = 0; j < LT; ++j) {
        const double d = S[j * LLD + j];
        if (!(d > 0.0)) {  // also catches NaN
            if (tid == 0) atomicCAS(info, 0, global_off + j + 1);
            failed = true;
            break;
        }
        const double rd = 1.0 / d;
        for (int i = j + 1 + ty; i < LT; i += 16) {
            const double lij = S[i * LLD + j] * rd;
            for (int k = j + 1 + tx; k <= i; k += 32) S[i * LLD + k] -= lij * S[k * LLD + j];
        }
        __syncthreads();
    }
    if (failed) {
        // poison the outputs so downstream results are visibly invalid; host reads `info`
        for (int idx = tid; idx < LT * LT; idx += LEAF_THREADS) {
            int i = idx >> 7, j = idx & 127;
            A[(int64_t)i * lda + j] = (j <= i) ? __longlong_as_double(0x7ff8000000000000LL) : 0.0;
            Dinv[idx] = (j <= i) ? __longlong_as_double(0x7ff8000000000000LL) : 0.0;
        }
        return;
    }
    // normalise: L[j][j] = sqrt(d_j), L[i][j] = S[i][j] / sqrt(d_j)
    if (tid < LT) dd[tid] = sqrt(S[tid * LLD + tid]);
    __syncthreads();
    for (int idx = tid; idx < LT * LT; idx += LEAF_THREADS) {
        int i = idx >> 7, j = idx & 127;
        if (j < i) S[i * LLD + j] /= dd[j];
    }
    __syncthreads();
    if (tid < LT) {
        S[tid * LLD + tid] = dd[tid];
        dd[tid] = 1.0 / dd[tid];
    }
    __syncthreads();
    // inverse: column c of X = L^-1 by forward substitution, 4 threads per column (same warp)
    // X[i][c] is kept at S[c][i] (strict upper), X[c][c] = dd[c].
    {
        const int c = tid >> 2, q = tid & 3;
        for (int i = 1; i < LT; ++i) {  // uniform trip count; columns with c >= i idle
            double part = 0.0;
            if (c < i) {
                for (int k = c + q; k < i; k += 4) {
                    const double xk = (k == c) ? dd[c] : S[c * LLD + k];
                    part += S[i * LLD + k] * xk;
                }
            }
            part += __shfl_xor_sync(0xffffffffu, part, 1);
            part += __shfl_xor_sync(0xffffffffu, part, 2);
            if (c < i && q == 0) S[c * LLD + i] = -part * dd[i];
            __syncwarp();
        }
    }
    __syncthreads();
    for (int idx = tid; idx < LT * LT; idx += LEAF_THREADS) {
        int i = idx >> 7, j = idx & 127;
        A[(int64_t)i * lda + j] = (j <= i) ? S[i * LLD + j] : 0.0;
        Dinv[idx] = (j < i) ? S[j * LLD + i] : (j == i ? dd[i] : 0.0);
    }
}

__global__ void zero_upper_tiles_kernel(double* A, int64_t lda, int nt) {
    // grid (nt, nt): zero tile (bi, bj) for bj > bi
    int bi = blockIdx.y, bj = blockIdx.x;
    if (bj <= bi) return;
    double* T = A + (int64_t)bi * LT * lda + (int64_t)bj * LT;
    for (int idx = threadIdx.x; idx < LT * LT / 2; idx += blockDim.x) {
        int i = idx >> 6, j2 = idx & 63;
        reinterpret_cast<double2*>(T + (int64_t)i * lda)[j2] = make_double2(0.0, 0.0);
    }
}

int leaf(gpx_ctx* h, double* A, int64_t lda, double* dinv_tile, int goff, int batch = 1, int64_t strideA = 0,
         int64_t strideD = 0) {
    static bool configured = false;
    if (!configured) {
        GPX_CUDA(cudaFuncSetAttribute(potrf_leaf_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, LEAF_SMEM));
        configured = true;
    }
    potrf_leaf_kernel<<<batch, LEAF_THREADS, LEAF_SMEM, h->stream>>>(A, lda, dinv_tile, h->d_info, goff, strideA, strideD);
    GPX_CHECK_LAUNCH(h);
    return 0;
}

GemmArgs base_args() {
    GemmArgs a{};
    a.batch = 1;
    a.alpha = 1.0;
    a.beta = 0.0;
    return a;
}

// split point (in elements) for a block of nt tiles
inline int64_t half_tiles(int64_t n) { return ((n / LT) / 2) * LT; }

// ---- right TRSM: X L^T = B in place (B is m x n, L is n x n lower), used by potrf ----------------
int trsm_right_lt(gpx_ctx* h, double* B, int64_t m, int64_t ldb, const double* L, int64_t n, int64_t ldl,
                  const double* dinv) {
    if (m <= 0) return 0;
    if (n == LT) {
        GemmArgs a = base_args();  // B <- B * Dinv^T : C[i][j] = sum_k B[i][k] Dinv[j][k]
        a.A = B; a.lda = ldb; a.a_kmajor = 1;
        a.B = dinv; a.ldb = LT; a.b_kmajor = 1;
        a.C = B; a.ldc = ldb;
        a.M = (int)m; a.N = LT; a.K = LT;
        return gpx_gemm_launch(h, a);
    }
    const int64_t h1 = half_tiles(n), h2 = n - h1;
    GPX_TRY(trsm_right_lt(h, B, m, ldb, L, h1, ldl, dinv));
    {
        GemmArgs a = base_args();  // B2 -= X1 * L21^T
        a.A = B; a.lda = ldb; a.a_kmajor = 1;
        a.B = L + h1 * ldl; a.ldb = ldl; a.b_kmajor = 1;
        a.C = B + h1; a.ldc = ldb;
        a.M = (int)m; a.N = (int)h2; a.K = (int)h1;
        a.alpha = -1.0; a.beta = 1.0;
        GPX_TRY(gpx_gemm_launch(h, a));
    }
    return trsm_right_lt(h, B + h1, m, ldb, L + h1 * ldl + h1, h2, ldl, dinv + (h1 / LT) * LT * LT);
}

int potrf_rec(gpx_ctx* h, double* A, int64_t n, int64_t lda, double* dinv, int goff) {
    if (n == LT) return leaf(h, A, lda, dinv, goff);
    const int64_t h1 = half_tiles(n), h2 = n - h1;
    GPX_TRY(potrf_rec(h, A, h1, lda, dinv, goff));
    double* A21 = A + h1 * lda;
    double* A22 = A21 + h1;
    GPX_TRY(trsm_right_lt(h, A21, h2, lda, A, h1, lda, dinv));
    {
        GemmArgs a = base_args();  // A22 -= A21 A21^T (lower tiles)
        a.A = A21; a.lda = lda; a.a_kmajor = 1;
        a.B = A21; a.ldb = lda; a.b_kmajor = 1;
        a.C = A22; a.ldc = lda;
        a.M = (int)h2; a.N = (int)h2; a.K = (int)h1;
        a.alpha = -1.0; a.beta = 1.0;
        a.lower_only = 1;
        GPX_TRY(gpx_gemm_launch(h, a));
    }
    return potrf_rec(h, A22, h2, lda, dinv + (h1 / LT) * LT * LT, goff + (int)h1);
}

// ---- left TRSM: L X = B (trans=0) or L^T X = B (trans=1), B is n x m in place -------------------
int trsm_left(gpx_ctx* h, const double* L, int64_t n, int64_t ldl, const double* dinv, int trans, double* B,
              int64_t m, int64_t ldb) {
    if (n == LT) {
        GemmArgs a = base_args();  // X = Dinv B or Dinv^T B
        a.A = dinv; a.lda = LT; a.a_kmajor = trans ? 0 : 1;
        a.B = B; a.ldb = ldb; a.b_kmajor = 0;
        a.C = B; a.ldc = ldb;
        a.M = LT; a.N = (int)m; a.K = LT;
        return gpx_gemm_launch(h, a);
    }
    const int64_t h1 = half_tiles(n), h2 = n - h1;
    const double* L21 = L + h1 * ldl;
    const double* L22 = L21 + h1;
    const double* dinv2 = dinv + (h1 / LT) * LT * LT;
    double* B2 = B + h1 * ldb;
    if (!trans) {
        GPX_TRY(trsm_left(h, L, h1, ldl, dinv, 0, B, m, ldb));
        GemmArgs a = base_args();  // B2 -= L21 X1
        a.A = L21; a.lda = ldl; a.a_kmajor = 1;
        a.B = B; a.ldb = ldb; a.b_kmajor = 0;
        a.C = B2; a.ldc = ldb;
        a.M = (int)h2; a.N = (int)m; a.K = (int)h1;
        a.alpha = -1.0; a.beta = 1.0;
        GPX_TRY(gpx_gemm_launch(h, a));
        return trsm_left(h, L22, h2, ldl, dinv2, 0, B2, m, ldb);
    } else {
        GPX_TRY(trsm_left(h, L22, h2, ldl, dinv2, 1, B2, m, ldb));
        GemmArgs a = base_args();  // B1 -= L21^T X2
        a.A = L21; a.lda = ldl; a.a_kmajor = 0;
        a.B = B2; a.ldb = ldb; a.b_kmajor = 0;
        a.C = B; a.ldc = ldb;
        a.M = (int)h1; a.N = (int)m; a.K = (int)h2;
        a.alpha = -1.0; a.beta = 1.0;
        GPX_TRY(gpx_gemm_launch(h, a));
        return trsm_left(h, L, h1, ldl, dinv, 1, B, m, ldb);
    }
}

__global__ void copy_tiles_kernel(double* __restrict__ dst, int64_t ldd, int64_t strideD, const double* __restrict__ src,
                                  int64_t strideS) {
    double* D = dst + (int64_t)blockIdx.x * strideD;
    const double* S = src + (int64_t)blockIdx.x * strideS;
    for (int idx = threadIdx.x; idx < LT * LT; idx += blockDim.x) {
        int i = idx >> 7, j = idx & 127;
        D[(int64_t)i * ldd + j] = S[idx];
    }
}

}  // namespace

extern "C" int gpx_potrf(gpx_handle h, double* A, int64_t n, int64_t lda, double* dinv) {
    GPX_REQUIRE(h != nullptr, 1);
    GPX_REQUIRE(A != nullptr && ((uintptr_t)A % 16) == 0, 2);
    GPX_REQUIRE(n > 0 && n % LT == 0, 3);
    GPX_REQUIRE(lda >= n && lda % 2 == 0, 4);
    GPX_REQUIRE(dinv != nullptr, 5);
    GPX_CUDA(cudaMemsetAsync(h->d_info, 0, sizeof(int), h->stream));
    GPX_TRY(potrf_rec(h, A, n, lda, dinv, 0));
    const int nt = (int)(n / LT);
    if (nt > 1) {
        zero_upper_tiles_kernel<<<dim3(nt, nt), 256, 0, h->stream>>>(A, lda, nt);
        GPX_CHECK_LAUNCH(h);
    }
    int info = 0;
    GPX_TRY(gpx_read_info(h, &info));
    if (info > 0) gpx_set_error("gpx_potrf: leading minor of order %d is not positive definite", info);
    return info;
}

extern "C" int gpx_trsm(gpx_handle h, const double* L, int64_t n, int64_t ldl, const double* dinv, int trans,
                        double* B, int64_t nrhs, int64_t ldb) {
    GPX_REQUIRE(h != nullptr, 1);
    GPX_REQUIRE(n > 0 && n % LT == 0, 3);
    GPX_REQUIRE(nrhs > 0 && nrhs % LT == 0, 8);
    return trsm_left(h, L, n, ldl, dinv, trans, B, nrhs, ldb);
}

// In-place L <- L^-1.  Leaves come from dinv; level by level (bottom-up) the off-diagonal block of every
// sub-problem of size 2*hs is formed as  A21 <- -(A22inv * (A21 * A11inv)) with the intermediate in `work`.
// Requires n/128 to be a power of two times an arbitrary count?  No: handled by explicit recursion below.
namespace {
int trtri_rec(gpx_ctx* h, double* A, int64_t n, int64_t lda, const double* dinv, double* work) {
    if (n == LT) {
        copy_tiles_kernel<<<1, 256, 0, h->stream>>>(A, lda, 0, dinv, 0);
        GPX_CHECK_LAUNCH(h);
        return 0;
    }
    const int64_t h1 = half_tiles(n), h2 = n - h1;
    double* A21 = A + h1 * lda;
    double* A22 = A21 + h1;
    GPX_TRY(trtri_rec(h, A, h1, lda, dinv, work));
    GPX_TRY(trtri_rec(h, A22, h2, lda, dinv + (h1 / LT) * LT * LT, work));
    {
        GemmArgs a = base_args();  // T = A21 * A11inv : T[i][j] = sum_{k>=j} A21[i][k] A11inv[k][j]
        a.A = A21; a.lda = lda; a.a_kmajor = 1;
        a.B = A; a.ldb = lda; a.b_kmajor = 0;
        a.C = work; a.ldc = h1;
        a.M = (int)h2; a.N = (int)h1; a.K = (int)h1;
        a.kb_mode = 2;
        GPX_TRY(gpx_gemm_launch(h, a));
    }
    {
        GemmArgs a = base_args();  // A21 = -A22inv * T : sum_{k<=i}
        a.A = A22; a.lda = lda; a.a_kmajor = 1;
        a.B = work; a.ldb = h1; a.b_kmajor = 0;
        a.C = A21; a.ldc = lda;
        a.M = (int)h2; a.N = (int)h1; a.K = (int)h2;
        a.alpha = -1.0;
        a.ke_mode = 1;
        a.rev_rows = 1;
        GPX_TRY(gpx_gemm_launch(h, a));
    }
    return 0;
}
}  // namespace

extern "C" int gpx_trtri(gpx_handle h, double* L, int64_t n, int64_t ldl, const double* dinv, double* work) {
    GPX_REQUIRE(h != nullptr, 1);
    GPX_REQUIRE(n > 0 && n % LT == 0, 3);
    GPX_REQUIRE(n == LT || work != nullptr, 6);
    return trtri_rec(h, L, n, ldl, dinv, work);
}

extern "C" int gpx_lauum(gpx_handle h, const double* Linv, int64_t n, int64_t ldl, double* out, int64_t ldo) {
    GPX_REQUIRE(h != nullptr, 1);
    GPX_REQUIRE(n > 0 && n % LT == 0, 3);
    GPX_REQUIRE(out != Linv, 5);
    GemmArgs a = base_args();  // out[i][j] = sum_{k >= i} Linv[k][i] Linv[k][j], i >= j
    a.A = Linv; a.lda = ldl; a.a_kmajor = 0;
    a.B = Linv; a.ldb = ldl; a.b_kmajor = 0;
    a.C = out; a.ldc = ldo;
    a.M = (int)n; a.N = (int)n; a.K = (int)n;
    a.lower_only = 1;
    a.kb_mode = 1;
    return gpx_gemm_launch(h, a);
}

// ---- internal entry points used by the multi-GPU driver (nccl_mg.cu) --------------------------------
int gpx_potrf_block(gpx_ctx* h, double* A, int64_t n, int64_t lda, double* dinv, int goff) {
    return potrf_rec(h, A, n, lda, dinv, goff);
}
int gpx_trsm_right_lt_block(gpx_ctx* h, double* B, int64_t m, int64_t ldb, const double* L, int64_t n, int64_t ldl,
                            const double* dinv) {
    return trsm_right_lt(h, B, m, ldb, L, n, ldl, dinv);
}

namespace {
// Left lower TRSM  L X = B  where only a *prefix* of B's columns is non-zero in any given row range: for rows
// below global row r (relative to this sub-problem's row 0 at global row `grow0`), the non-zero columns are the
// local blocks whose global block index is <= block(r): count = prefix(r).  Used for L^-1 on block-cyclic columns.
struct PrefixMap { int P, p, nb; };
inline int64_t prefix_cols(const PrefixMap& pm, int64_t grow_end) {
    // number of local columns (elements) whose global block start is < grow_end
    const int64_t nblk_below = (grow_end + pm.nb - 1) / pm.nb;              // global blocks 0..nblk_below-1 start below grow_end
    const int64_t cnt = nblk_below > pm.p ? (nblk_below - pm.p + pm.P - 1) / pm.P : 0;  // those owned by rank p
    return cnt * pm.nb;
}
int trsm_left_prefix(gpx_ctx* h, const double* L, int64_t n, int64_t ldl, const double* dinv, double* B, int64_t ldb,
                     int64_t grow0, const PrefixMap& pm) {
    const int64_t ncols = prefix_cols(pm, grow0 + n);   // columns that can be non-zero within these rows
    if (ncols <= 0) return 0;
    if (n == LT) {
        GemmArgs a = base_args();
        a.A = dinv; a.lda = LT; a.a_kmajor = 1;
        a.B = B; a.ldb = ldb; a.b_kmajor = 0;
        a.C = B; a.ldc = ldb;
        a.M = LT; a.N = (int)ncols; a.K = LT;
        return gpx_gemm_launch(h, a);
    }
    const int64_t h1 = half_tiles(n), h2 = n - h1;
    GPX_TRY(trsm_left_prefix(h, L, h1, ldl, dinv, B, ldb, grow0, pm));
    const int64_t nc1 = prefix_cols(pm, grow0 + h1);   // X1 is non-zero only in these columns
    if (nc1 > 0) {
        GemmArgs a = base_args();  // B2[:, :nc1] -= L21 X1[:, :nc1]
        a.A = L + h1 * ldl; a.lda = ldl; a.a_kmajor = 1;
        a.B = B; a.ldb = ldb; a.b_kmajor = 0;
        a.C = B + h1 * ldb; a.ldc = ldb;
        a.M = (int)h2; a.N = (int)nc1; a.K = (int)h1;
        a.alpha = -1.0; a.beta = 1.0;
        // X1[k][c] is zero for global rows above the start of c's block: begin the k loop there
        a.kb_mode = 3; a.cyc_P = pm.P; a.cyc_p = pm.p; a.cyc_tpb = pm.nb / LT; a.cyc_q0 = 0; a.cyc_row_base = (int)grow0;
        a.cyc_b_rows = 0;
        GPX_TRY(gpx_gemm_launch(h, a));
    }
    return trsm_left_prefix(h, L + h1 * ldl + h1, h2, ldl, dinv + (h1 / LT) * LT * LT, B + h1 * ldb, ldb, grow0 + h1, pm);
}
}  // namespace

int gpx_trsm_left_prefix_block(gpx_ctx* h, const double* L, int64_t n, int64_t ldl, const double* dinv, double* B, int64_t ldb,
                               int P, int p, int nb) {
    PrefixMap pm{P, p, nb};
    return trsm_left_prefix(h, L, n, ldl, dinv, B, ldb, 0, pm);
}
