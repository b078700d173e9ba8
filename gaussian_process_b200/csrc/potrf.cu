// Blocked FP64 Cholesky, triangular solves and inverse for sm_100a.
//
//   potrf  : recursive right-looking lower Cholesky.  Leaves are 128x128 diagonal blocks factored by a
//            single-CTA shared-memory kernel that also emits the inverse of the leaf (used by every
//            triangular solve as a GEMM operand); everything else is the DMMA GEMM of gemm.cu
//            (TRSM against the leaf inverses, SYRK/GEMM trailing updates with k = half the block).
//            Between 1024 and 16384 a three-stream look-ahead variant takes over (potrf_la_grouped): the panel chain
//            factors leaves WITHOUT their inverses and solves by warp-per-row substitution (trsm_sub_kernel), the
//            bulk of the trailing matrix is updated once per group of panels with K = G*128, and all leaf inverses
//            come from one batched launch at the end.
//   trsm   : recursive, left side, lower, N or T.
//   trtri  : recursive in-place inverse of L (two triangular GEMMs per level, batched over the
//            independent sub-problems of that level).
//   lauum  : out = Linv^T Linv (lower) in one triangular-aware launch.
// Replaces np.linalg.cholesky / np.linalg.solve(L, .) / np.linalg.inv(L) / np.dot(inv(L.T), inv(L))
// (SURVEY.md 8a rows A4, A5).
#include <stdlib.h>
#include <algorithm>
#include <vector>
#include "common.cuh"

namespace {

constexpr int LT = 128;          // leaf size
constexpr int LB = 4;            // register block edge: thread (bi,bj) owns the 4x4 block of the lower triangle
constexpr int NBK = LT / LB;     // 32 block rows -> 528 lower blocks
constexpr int LEAF_THREADS = 544;
constexpr int LEAF_SMEM = (NBK * 578 + LT) * (int)sizeof(double);

// One CTA factors one 128x128 diagonal block and inverts the factor.  The whole lower triangle lives in
// REGISTERS (one 4x4 block per thread); shared memory only carries the finished panel values:
//   step p (32 steps): thread (p,p) factors its 4x4 block and inverts it | panel threads (i,p) multiply by
//   L_pp^-T | every thread (i,j>p) applies the rank-4 update from shared memory.  2 barriers per step.
// Then X = L^-1 right-looking over block rows (1 barrier per step), X stored transposed in the unused upper
// triangle of the same smem tile:  S(lower+diag) = L, S(strict upper) = X^T, dd = diag(X).
//  A tile <- L (upper zeroed) ; Dinv tile <- L^-1 (upper zeroed)
// info: first failing pivot (global 1-based index) is recorded once.
// 4x4 blocks are stored block-major with 16-byte skews (block = 18 doubles, block row = 578 doubles) so that a warp whose
// lanes read the same chunk of 32 different blocks -- along a block row or a block column -- hits distinct bank groups.
__device__ __forceinline__ int soff_full(int a, int b) { return (a >> 2) * 578 + (b >> 2) * 18 + (a & 3) * 4 + (b & 3); }
__device__ long long g_leaf_dbg[8];
// MODE 0: factor + inverse (A <- L, Dinv <- L^-1); MODE 1: factor only (the look-ahead chain: the explicit inverse is taken
// off the critical path and formed later by one batched MODE 2 launch); MODE 2: inverse only (A holds L, Dinv <- L^-1).
template <int MODE>
__device__ __forceinline__ void potrf_leaf_body(double* __restrict__ A, int64_t lda, double* __restrict__ Dinv, int* info,
                                                int global_off, int64_t strideA, int64_t strideD, double* sm, int& fail_col) {
    auto soff = [](int a, int b) -> int { return soff_full(a, b); };
    double* S = sm;
    double* dd = sm + NBK * 578;
    A += (int64_t)blockIdx.x * strideA;
    Dinv += (int64_t)blockIdx.x * strideD;
    global_off += blockIdx.x * LT;
    const int tid = threadIdx.x;
    const long long t_start = clock64();
    // Factorisation phase: blocks are numbered COLUMN-major (all blocks of block column 0, then column 1, ...), so the
    // threads that still have work at step p (bj > p) are a contiguous tail of the CTA -- finished warps skip the step
    // entirely -- and the panel threads (bj == p+1) sit in one or two warps instead of one lane in each of 17 warps.
    // Inverse phase: ROW-major numbering (ri, rj), for the same reason applied to block rows.
    int bi = -1, bj = -1, ri = -1, rj = -1;
    if (tid < NBK * (NBK + 1) / 2) {
        int c = 0, off = 0;
        while (off + (NBK - c) <= tid) { off += NBK - c; ++c; }
        bj = c;
        bi = c + (tid - off);
        ri = (int)((sqrtf(8.f * tid + 1.f) - 1.f) * 0.5f);
        while (ri * (ri + 1) / 2 > tid) --ri;
        while ((ri + 1) * (ri + 2) / 2 <= tid) ++ri;
        rj = tid - ri * (ri + 1) / 2;
    }
    const bool active = bi >= 0;
    double a[LB][LB];
    if (active) {
#pragma unroll
        for (int r = 0; r < LB; ++r) {
            const double2* src = reinterpret_cast<const double2*>(A + (int64_t)(LB * bi + r) * lda + LB * bj);
            double2 v0 = src[0], v1 = src[1];
            a[r][0] = v0.x; a[r][1] = v0.y; a[r][2] = v1.x; a[r][3] = v1.y;
        }
    }
    if (tid == 0) fail_col = -1;
    if (MODE == 2 && active) {
        // inverse only: the block just loaded IS the factor.  S(lower + diag) <- L, and every diagonal thread forms its
        // X_pp = L_pp^-1 (4x4, in registers) -> S(strict upper of the diagonal block) = X_pp^T, dd = diag(X_pp).
        if (bi == bj) {
            double x[LB][LB];
#pragma unroll
            for (int c = 0; c < LB; ++c) x[c][c] = 1.0 / a[c][c];
#pragma unroll
            for (int c = 0; c < LB; ++c)
#pragma unroll
                for (int r = c + 1; r < LB; ++r) {
                    double v = 0.0;
#pragma unroll
                    for (int m = c; m < r; ++m) v += a[r][m] * x[m][c];
                    x[r][c] = -v * x[r][r];
                }
#pragma unroll
            for (int r = 0; r < LB; ++r) {
                dd[LB * bi + r] = x[r][r];
#pragma unroll
                for (int c = 0; c <= r; ++c) S[soff(LB * bi + r, LB * bi + c)] = a[r][c];
#pragma unroll
                for (int c = 0; c < r; ++c) S[soff(LB * bi + c, LB * bi + r)] = x[r][c];
            }
        } else {
#pragma unroll
            for (int r = 0; r < LB; ++r) {
                double2* dst = reinterpret_cast<double2*>(S + soff(LB * bi + r, LB * bj));
                dst[0] = make_double2(a[r][0], a[r][1]);
                dst[1] = make_double2(a[r][2], a[r][3]);
            }
        }
    }
    __syncthreads();
    const long long t_loaded = clock64();

    // ---- factorisation with intra-kernel look-ahead: in every step ALL trailing blocks apply the rank-4 update of
    // panel p, and the owner of the next diagonal block factors it right away (its ~700-cycle serial chain overlaps
    // the bulk update of the other threads); the panel multiply follows after one barrier.
    auto diag_factor = [&](int p) {        // thread (p,p): 4x4 Cholesky + inverse of the diagonal block, in registers
        double l[LB][LB], x[LB][LB];
        int bad = -1;
#pragma unroll
        for (int c = 0; c < LB; ++c) {
            double d = a[c][c];
#pragma unroll
            for (int m = 0; m < c; ++m) d -= l[c][m] * l[c][m];
            if (!(d > 0.0) && bad < 0) bad = c;
            // 1/sqrt(d) by rsqrt + Newton (shorter dependent chain than sqrt followed by a division)
            double rc = rsqrt(d);
            rc = rc * (1.5 - 0.5 * d * rc * rc);
            double lc = d * rc;
            lc = lc + 0.5 * rc * fma(-lc, lc, d);
            rc = rc + rc * fma(-lc, rc, 1.0);
            l[c][c] = lc;
            x[c][c] = rc;
#pragma unroll
            for (int r = c + 1; r < LB; ++r) {
                double v = a[r][c];
#pragma unroll
                for (int m = 0; m < c; ++m) v -= l[r][m] * l[c][m];
                l[r][c] = v * rc;
            }
        }
#pragma unroll
        for (int c = 0; c < LB; ++c)
#pragma unroll
            for (int r = c + 1; r < LB; ++r) {
                double v = 0.0;
#pragma unroll
                for (int m = c; m < r; ++m) v += l[r][m] * x[m][c];
                x[r][c] = -v * x[r][r];
            }
        if (bad >= 0) fail_col = LB * p + bad;
#pragma unroll
        for (int r = 0; r < LB; ++r) {
            dd[LB * p + r] = x[r][r];
#pragma unroll
            for (int c = 0; c <= r; ++c) S[soff(LB * p + r, LB * p + c)] = l[r][c];
#pragma unroll
            for (int c = 0; c < r; ++c) S[soff(LB * p + c, LB * p + r)] = x[r][c];   // X_pp^T above the diagonal
        }
    };
    auto panel_mul = [&](int p) {          // thread (i,p), i>p: A_ip <- A_ip * L_pp^-T = A_ip * X_pp^T
        double x[LB][LB];
#pragma unroll
        for (int r = 0; r < LB; ++r) {
            x[r][r] = dd[LB * p + r];
#pragma unroll
            for (int c = 0; c < r; ++c) x[r][c] = S[soff(LB * p + c, LB * p + r)];
        }
        double o[LB][LB];
#pragma unroll
        for (int r = 0; r < LB; ++r)
#pragma unroll
            for (int c = 0; c < LB; ++c) o[r][c] = a[r][0] * x[c][0];
#pragma unroll
        for (int m = 1; m < LB; ++m)
#pragma unroll
            for (int r = 0; r < LB; ++r)
#pragma unroll
                for (int c = m; c < LB; ++c) o[r][c] = fma(a[r][m], x[c][m], o[r][c]);
#pragma unroll
        for (int r = 0; r < LB; ++r) {
            double2* dst = reinterpret_cast<double2*>(S + soff(LB * bi + r, LB * p));
            dst[0] = make_double2(o[r][0], o[r][1]);
            dst[1] = make_double2(o[r][2], o[r][3]);
        }
    };
    auto rank4_update = [&](int p) {       // thread (i,j), j>p: A_ij -= L_ip L_jp^T
        double li[LB][LB], lj[LB][LB];
#pragma unroll
        for (int r = 0; r < LB; ++r) {
            const double2* pi_ = reinterpret_cast<const double2*>(S + soff(LB * bi + r, LB * p));
            const double2* pj_ = reinterpret_cast<const double2*>(S + soff(LB * bj + r, LB * p));
            double2 u0 = pi_[0], u1 = pi_[1], w0 = pj_[0], w1 = pj_[1];
            li[r][0] = u0.x; li[r][1] = u0.y; li[r][2] = u1.x; li[r][3] = u1.y;
            lj[r][0] = w0.x; lj[r][1] = w0.y; lj[r][2] = w1.x; lj[r][3] = w1.y;
        }
        // m outermost: 16 independent FMAs between dependent ones (DFMA latency ~16 cycles would otherwise serialise)
#pragma unroll
        for (int m = 0; m < LB; ++m)
#pragma unroll
            for (int r = 0; r < LB; ++r)
#pragma unroll
                for (int c = 0; c < LB; ++c) a[r][c] = fma(-li[r][m], lj[c][m], a[r][c]);
    };

    if (MODE != 2) {
    if (active && bi == 0 && bj == 0) diag_factor(0);
    __syncthreads();
    if (fail_col < 0) {
        if (active && bj == 0 && bi > 0) panel_mul(0);
        __syncthreads();
        for (int p = 0; p + 1 < NBK; ++p) {
            if (active && bj > p) {
                rank4_update(p);
                if (bi == p + 1 && bj == p + 1) diag_factor(p + 1);
            }
            __syncthreads();
            if (fail_col >= 0) break;
            if (active && bj == p + 1 && bi > p + 1) panel_mul(p + 1);
            __syncthreads();
        }
    }
    if (fail_col >= 0) {
        if (tid == 0) atomicCAS(info, 0, global_off + fail_col + 1);
        // poison the outputs so downstream results are visibly invalid; the host reads `info`
        const double qnan = __longlong_as_double(0x7ff8000000000000LL);
        for (int idx = tid; idx < LT * LT; idx += LEAF_THREADS) {
            int i = idx >> 7, j = idx & 127;
            A[(int64_t)i * lda + j] = (j <= i) ? qnan : 0.0;
            if (MODE == 0) Dinv[idx] = (j <= i) ? qnan : 0.0;
        }
        return;
    }
    }   // MODE != 2
    if (MODE == 1) {   // factor only: store L (upper triangle zeroed) and leave
        for (int idx = tid; idx < LT * LT; idx += LEAF_THREADS) {
            int i = idx >> 7, j = idx & 127;
            A[(int64_t)i * lda + j] = (j <= i) ? S[soff(i, j)] : 0.0;
        }
        return;
    }

    const long long t_fact = clock64();
    // ---- X = L^-1, right-looking over block rows k with look-ahead: acc_ij += L_ik X_kj for every i > k, and the
    // threads of row k+1 finalise X_{k+1,j} = -X_{k+1,k+1} acc right away (one barrier per step).
    double acc[LB][LB];
#pragma unroll
    for (int r = 0; r < LB; ++r)
#pragma unroll
        for (int c = 0; c < LB; ++c) acc[r][c] = 0.0;
    for (int k = 0; k + 1 < NBK; ++k) {
        if (active && ri > k && rj <= k) {
            double li[LB][LB], xt[LB][LB];     // xt[c][m] = X_kj[m][c]
#pragma unroll
            for (int r = 0; r < LB; ++r) {
                const double2* pi_ = reinterpret_cast<const double2*>(S + soff(LB * ri + r, LB * k));
                double2 u0 = pi_[0], u1 = pi_[1];
                li[r][0] = u0.x; li[r][1] = u0.y; li[r][2] = u1.x; li[r][3] = u1.y;
            }
            if (rj < k) {
#pragma unroll
                for (int c = 0; c < LB; ++c) {
                    const double2* px = reinterpret_cast<const double2*>(S + soff(LB * rj + c, LB * k));
                    double2 u0 = px[0], u1 = px[1];
                    xt[c][0] = u0.x; xt[c][1] = u0.y; xt[c][2] = u1.x; xt[c][3] = u1.y;
                }
            } else {  // rj == k: X_kk (lower triangular)
#pragma unroll
                for (int c = 0; c < LB; ++c)
#pragma unroll
                    for (int m = 0; m < LB; ++m)
                        xt[c][m] = (m == c) ? dd[LB * k + c] : (m > c ? S[soff(LB * k + c, LB * k + m)] : 0.0);
            }
#pragma unroll
            for (int m = 0; m < LB; ++m)
#pragma unroll
                for (int r = 0; r < LB; ++r)
#pragma unroll
                    for (int c = 0; c < LB; ++c) acc[r][c] = fma(li[r][m], xt[c][m], acc[r][c]);
            if (ri == k + 1) {                 // row k+1 is complete for every j <= k: finalise and publish (transposed)
                const int kk = k + 1;
                double x[LB][LB];
#pragma unroll
                for (int r = 0; r < LB; ++r) {
                    x[r][r] = dd[LB * kk + r];
#pragma unroll
                    for (int c = 0; c < r; ++c) x[r][c] = S[soff(LB * kk + c, LB * kk + r)];
                }
#pragma unroll
                for (int c = 0; c < LB; ++c) {
                    double o[LB];
#pragma unroll
                    for (int r = 0; r < LB; ++r) {
                        double v = 0.0;
#pragma unroll
                        for (int m = 0; m <= r; ++m) v += x[r][m] * acc[m][c];
                        o[r] = -v;
                    }
                    double2* dst = reinterpret_cast<double2*>(S + soff(LB * rj + c, LB * kk));
                    dst[0] = make_double2(o[0], o[1]);
                    dst[1] = make_double2(o[2], o[3]);
                }
            }
        }
        __syncthreads();
    }
    __syncthreads();
    const long long t_inv = clock64();
    for (int idx = tid; idx < LT * LT; idx += LEAF_THREADS) {
        int i = idx >> 7, j = idx & 127;
        if (MODE == 0) A[(int64_t)i * lda + j] = (j <= i) ? S[soff(i, j)] : 0.0;
        Dinv[idx] = (j < i) ? S[soff(j, i)] : (j == i ? dd[i] : 0.0);
    }
    if (tid == 0 && blockIdx.x == 0) {
        g_leaf_dbg[0] = t_loaded - t_start;
        g_leaf_dbg[1] = t_fact - t_loaded;
        g_leaf_dbg[2] = t_inv - t_fact;
        g_leaf_dbg[3] = clock64() - t_inv;
    }
}

template <int MODE>
__global__ void __launch_bounds__(LEAF_THREADS, 1)
potrf_leaf_kernel(double* __restrict__ A, int64_t lda, double* __restrict__ Dinv, int* info, int global_off,
                  int64_t strideA, int64_t strideD) {
    extern __shared__ __align__(16) double sm[];
    __shared__ int fail_col;
    potrf_leaf_body<MODE>(A, lda, Dinv, info, global_off, strideA, strideD, sm, fail_col);
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

// mode 0: factor + inverse; 1: factor only; 2: inverse only (see potrf_leaf_kernel)
int leaf(gpx_ctx* h, double* A, int64_t lda, double* dinv_tile, int goff, int batch = 1, int64_t strideA = 0,
         int64_t strideD = 0, int mode = 0) {
    static bool configured_dev[GPX_MAX_DEVICES] = {};
    bool& configured = configured_dev[h->device % GPX_MAX_DEVICES];   // cudaFuncSetAttribute is per device
    if (!configured) {
        GPX_CUDA(cudaFuncSetAttribute(potrf_leaf_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, LEAF_SMEM));
        GPX_CUDA(cudaFuncSetAttribute(potrf_leaf_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, LEAF_SMEM));
        GPX_CUDA(cudaFuncSetAttribute(potrf_leaf_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, LEAF_SMEM));
        configured = true;
    }
    gpx_timing_leaf_begin(h);
    if (mode == 1) potrf_leaf_kernel<1><<<batch, LEAF_THREADS, LEAF_SMEM, h->stream>>>(A, lda, dinv_tile, h->d_info, goff, strideA, strideD);
    else if (mode == 2) potrf_leaf_kernel<2><<<batch, LEAF_THREADS, LEAF_SMEM, h->stream>>>(A, lda, dinv_tile, h->d_info, goff, strideA, strideD);
    else potrf_leaf_kernel<0><<<batch, LEAF_THREADS, LEAF_SMEM, h->stream>>>(A, lda, dinv_tile, h->d_info, goff, strideA, strideD);
    GPX_CHECK_LAUNCH(h);
    gpx_timing_leaf_end(h);
    return 0;
}

// ---- X L^T = B in place for ONE 128 x 128 lower-triangular leaf L (row-major, ldl), any number of rows of B: forward
// substitution, one WARP per row (the row lives in registers, 4 entries per lane), L^T packed in shared memory, no
// barrier inside the 128 steps (one multiply, one shuffle, up to four FMAs each).  ~7 us for up to 4736 rows: this is the
// TRSM of the look-ahead chain, which so far had to wait for the explicit inverse of the leaf (27 us) and then ran as a GEMM.
constexpr int TS_WARPS = 32;
constexpr int TS_SMEM = (LT * (LT + 1) / 2 + LT) * (int)sizeof(double);
__global__ void __launch_bounds__(TS_WARPS * 32) trsm_sub_kernel(double* __restrict__ B, int64_t ldb, int64_t rows,
                                                                  const double* __restrict__ L, int64_t ldl) {
    extern __shared__ __align__(16) double tsm[];
    double* U = tsm;                          // U[off(m) + (k - m)] = L[k][m], k >= m : column m of L from the diagonal down
    double* rd = tsm + LT * (LT + 1) / 2;     // 1 / L[m][m]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int idx = tid; idx < LT * LT; idx += TS_WARPS * 32) {
        const int k = idx >> 7, m = idx & 127;
        if (m <= k) U[m * LT - (m * (m - 1)) / 2 + (k - m)] = L[(int64_t)k * ldl + m];
    }
    if (tid < LT) rd[tid] = 1.0 / L[(int64_t)tid * ldl + tid];
    __syncthreads();
    const int64_t r = (int64_t)blockIdx.x * TS_WARPS + warp;
    if (r >= rows) return;
    double* row = B + r * ldb;
    double a[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) a[i] = row[lane + 32 * i];
#pragma unroll
    for (int slot = 0; slot < 4; ++slot) {
#pragma unroll 8
        for (int mm = 0; mm < 32; ++mm) {
            const int m = slot * 32 + mm;
            const double* u = U + (m * LT - (m * (m - 1)) / 2) - m;   // u[k] = L[k][m]
            const double xm = __shfl_sync(0xffffffffu, a[slot] * rd[m], mm);
            if (lane == mm) a[slot] = xm;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                if (i < slot) continue;
                const int k = lane + 32 * i;
                if (k > m) a[i] = fma(-xm, u[k], a[i]);
            }
        }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) row[lane + 32 * i] = a[i];
}

int trsm_sub(gpx_ctx* h, double* B, int64_t rows, int64_t ldb, const double* L, int64_t ldl) {
    if (rows <= 0) return 0;
    static bool configured_dev[GPX_MAX_DEVICES] = {};
    bool& configured = configured_dev[h->device % GPX_MAX_DEVICES];
    if (!configured) {
        GPX_CUDA(cudaFuncSetAttribute(trsm_sub_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TS_SMEM));
        configured = true;
    }
    trsm_sub_kernel<<<(unsigned)((rows + TS_WARPS - 1) / TS_WARPS), TS_WARPS * 32, TS_SMEM, h->stream>>>(B, ldb, rows, L, ldl);
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

int potrf_la(gpx_ctx* h, double* A, int64_t n, int64_t lda, double* dinv, int goff, int nb);
int la_threshold();
int la_block(int64_t n);

int potrf_rec(gpx_ctx* h, double* A, int64_t n, int64_t lda, double* dinv, int goff) {
    if (n == LT) return leaf(h, A, lda, dinv, goff);
    // mid-size blocks: the serial chain of leaves / small GEMMs dominates -> look-ahead right-looking algorithm
    if (n >= 1024 && n <= la_threshold()) {
        const int nb = la_block(n);
        if (n % nb == 0 && n >= 4 * nb) return potrf_la(h, A, n, lda, dinv, goff, nb);
    }
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


namespace {
// Right-looking blocked Cholesky with LOOK-AHEAD on two streams (used below a size threshold where the serial chain of
// small kernels, not the DMMA pipe, bounds the recursive formulation).  Panel width nb; after panel j is final:
//   chain stream H (high priority): a. update block column j+1 with panel j   b. factor panel j+1 (diag potrf + TRSM)
//                                   c. [after the bulk update j-1 finished] update block column j+2 with panel j
//   bulk stream S                 : update block columns >= j+3 with panel j
// so the latency-bound panel chain runs concurrently with (and up to two panels ahead of) the DMMA-bound bulk updates.
int potrf_la_split(gpx_ctx* h, double* A, int64_t n, int64_t lda, double* dinv, int goff, int nb);
int potrf_la_grouped(gpx_ctx* h, double* A, int64_t n, int64_t lda, double* dinv, int goff, int G);
int la_split();
int la_group(int64_t n);
int potrf_la(gpx_ctx* h, double* A, int64_t n, int64_t lda, double* dinv, int goff, int nb) {
    if (la_split() && h->aux2_stream && nb == LT && la_group(n) > 1) return potrf_la_grouped(h, A, n, lda, dinv, goff, la_group(n));
    if (la_split() && h->aux2_stream) return potrf_la_split(h, A, n, lda, dinv, goff, nb);
    const int64_t nblk = n / nb;
    const int tpb = nb / LT;
    cudaStream_t S = h->stream, H = h->aux_stream;
    GpxEventSet es;
    GPX_TRY(es.create(2 * nblk + 1));
    cudaEvent_t* evP = es.ev.data();
    cudaEvent_t* evS = evP + nblk;
    cudaEvent_t ev0 = es.ev[2 * nblk];
    GPX_CUDA(cudaEventRecord(ev0, S));            // everything queued on S so far (covariance build ...) precedes the chain
    GPX_CUDA(cudaStreamWaitEvent(H, ev0, 0));
    auto on_H = [&](auto&& fn) -> int { h->stream = H; int r = fn(); h->stream = S; return r; };
    auto factor_panel = [&](int64_t j) -> int {
        double* diag = A + j * nb * lda + j * nb;
        double* dj = dinv + j * tpb * LT * LT;
        GPX_TRY(potrf_rec(h, diag, nb, lda, dj, goff + (int)(j * nb)));
        const int64_t below = n - (j + 1) * nb;
        if (below > 0) GPX_TRY(trsm_right_lt(h, diag + (int64_t)nb * lda, below, lda, diag, nb, lda, dj));
        return 0;
    };
    // update block columns [c0, c1) with panel j (rows from c0*nb down, lower tiles only)
    auto update = [&](int64_t j, int64_t c0, int64_t c1) -> int {
        if (c1 > nblk) c1 = nblk;
        if (c0 >= c1) return 0;
        const int64_t r0 = c0 * nb;
        GemmArgs a = base_args();
        a.A = A + r0 * lda + j * nb; a.lda = lda; a.a_kmajor = 1;
        a.B = a.A; a.ldb = lda; a.b_kmajor = 1;
        a.C = A + r0 * lda + r0; a.ldc = lda;
        a.M = (int)(n - r0); a.N = (int)((c1 - c0) * nb); a.K = nb;
        a.alpha = -1.0; a.beta = 1.0;
        a.lower_only = 1;
        return gpx_gemm_launch(h, a);
    };
    int rc = on_H([&]() { return factor_panel(0); });
    if (rc == 0) rc = cudaEventRecord(evP[0], H) == cudaSuccess ? 0 : GPX_E_CUDA;
    for (int64_t j = 0; j < nblk && rc == 0; ++j) {
        // ---- chain (H)
        if (j + 1 < nblk) {
            rc = on_H([&]() {
                GPX_TRY(update(j, j + 1, j + 2));
                return factor_panel(j + 1);
            });
            if (rc == 0 && cudaEventRecord(evP[j + 1], H) != cudaSuccess) rc = GPX_E_CUDA;
            if (rc == 0 && j + 2 < nblk) {
                if (j >= 1) cudaStreamWaitEvent(H, evS[j - 1], 0);
                rc = on_H([&]() { return update(j, j + 2, j + 3); });
            }
        }
        // ---- bulk (S)
        if (rc == 0) {
            cudaStreamWaitEvent(S, evP[j], 0);
            rc = update(j, j + 3, nblk);
            if (rc == 0 && cudaEventRecord(evS[j], S) != cudaSuccess) rc = GPX_E_CUDA;
        }
    }
    // join: S continues only after the chain has finished -- also on error, so that the caller never recycles A / dinv
    // while the auxiliary stream is still writing them
    if (cudaEventRecord(ev0, H) == cudaSuccess) cudaStreamWaitEvent(S, ev0, 0);
    h->stream = S;
    return rc;   // no host synchronisation; EventSet releases the events
}

// The same right-looking factorisation with the panel chain cut down to what the NEXT diagonal block really waits for.
// With panel j final in its diagonal block (Linv_j known), only the first block row below it gates the next leaf:
//   H  (chain, high priority): T_top(j): row block j+1 of column j  <- . Linv_j^T
//                              U_diag(j->j+1): diagonal block j+1    -= L(j+1,j) L(j+1,j)^T
//                              factor diagonal block j+1
//   H2 (column work)         : T_rest(j): row blocks >= j+2 of column j;  U_rest(j->j+1): rows >= j+2 of column j+1;
//                              U2(j->j+2): column j+2 (rows >= j+2)
//   S  (bulk)                : columns >= j+3
// so a chain step is {one-block TRSM, one-block SYRK, diagonal factorisation} instead of three column-high GEMMs plus the
// factorisation.  Every block still receives its updates in panel order (S: panels <= c-3, H2: c-2, H2/H: c-1), hence the
// result is bit-identical to potrf_la's.  Events: Leaf[j] (H), Top[j] (H), P[j] / U1[j] / U2[j] (H2), Sd[j] (S).
int potrf_la_split(gpx_ctx* h, double* A, int64_t n, int64_t lda, double* dinv, int goff, int nb) {
    const int64_t nblk = n / nb;
    const int tpb = nb / LT;
    cudaStream_t S = h->stream, H = h->aux_stream, H2 = h->aux2_stream;
    enum { LEAF = 0, TOP, PAN, U1, U2, SD, NEV };
    GpxEventSet es;
    GPX_TRY(es.create(NEV * nblk + 2));
    auto E = [&](int kind, int64_t j) -> cudaEvent_t { return es.ev[kind * nblk + j]; };
    int rc = 0;
    auto rec = [&](int kind, int64_t j, cudaStream_t st) { if (rc == 0 && cudaEventRecord(E(kind, j), st) != cudaSuccess) rc = GPX_E_CUDA; };
    auto wait = [&](cudaStream_t st, int kind, int64_t j) { if (rc == 0 && j >= 0 && cudaStreamWaitEvent(st, E(kind, j), 0) != cudaSuccess) rc = GPX_E_CUDA; };
    auto on = [&](cudaStream_t st, auto&& fn) { if (rc != 0) return; h->stream = st; rc = fn(); h->stream = S; };
    cudaEvent_t ev0 = es.ev[NEV * nblk], ev1 = es.ev[NEV * nblk + 1];
    GPX_CUDA(cudaEventRecord(ev0, S));            // everything queued on S so far precedes the factorisation
    GPX_CUDA(cudaStreamWaitEvent(H, ev0, 0));
    GPX_CUDA(cudaStreamWaitEvent(H2, ev0, 0));
    auto blk = [&](int64_t i, int64_t j) -> double* { return A + i * nb * lda + j * nb; };
    auto dj = [&](int64_t j) -> double* { return dinv + j * tpb * LT * LT; };
    // C(rows r0.., block column c) -= A(rows r0.., block column j) * A(block row c, block column j)^T
    auto update = [&](int64_t j, int64_t r0, int64_t rows, int64_t c, int64_t ncols, int lower) -> int {
        if (rows <= 0 || ncols <= 0) return 0;
        GemmArgs a = base_args();
        a.A = blk(r0, j); a.lda = lda; a.a_kmajor = 1;
        a.B = blk(c, j); a.ldb = lda; a.b_kmajor = 1;
        a.C = blk(r0, c); a.ldc = lda;
        a.M = (int)(rows * nb); a.N = (int)(ncols * nb); a.K = nb;
        a.alpha = -1.0; a.beta = 1.0;
        a.lower_only = lower;
        return gpx_gemm_launch(h, a);
    };
    // nb == 128: the chain factors leaves WITHOUT their explicit inverses and solves by substitution (trsm_sub); the
    // inverses every later solve needs are formed by one batched launch after the factorisation (off the critical path).
    const bool sub = (nb == LT);
    auto factor_diag = [&](int64_t j) -> int {
        if (sub) return leaf(h, blk(j, j), lda, dj(j), goff + (int)(j * nb), 1, 0, 0, 1);
        return potrf_rec(h, blk(j, j), nb, lda, dj(j), goff + (int)(j * nb));
    };
    auto panel_trsm = [&](int64_t i0, int64_t rows, int64_t j) -> int {
        if (sub) return trsm_sub(h, blk(i0, j), rows, lda, blk(j, j), lda);
        return trsm_right_lt(h, blk(i0, j), rows, lda, blk(j, j), nb, lda, dj(j));
    };
    on(H, [&]() { return factor_diag(0); });
    rec(LEAF, 0, H);
    for (int64_t j = 0; j < nblk && rc == 0; ++j) {
        const int64_t rest = nblk - (j + 2);       // block rows below block row j+1
        // ---- H: the part of panel j the next diagonal block waits for
        if (j + 1 < nblk) {
            wait(H, U1, j - 1); wait(H, U2, j - 2); wait(H, SD, j - 3);          // block (j+1, j) has every earlier update
            on(H, [&]() { return panel_trsm(j + 1, nb, j); });
            rec(TOP, j, H);
            wait(H, U2, j - 1); wait(H, SD, j - 2);                              // block (j+1, j+1) likewise
            on(H, [&]() {
                GPX_TRY(update(j, j + 1, 1, j + 1, 1, 1));
                return factor_diag(j + 1);
            });
            rec(LEAF, j + 1, H);
        }
        // ---- H2: the rest of column j, then its updates of columns j+1 and j+2
        if (rest > 0) {
            wait(H2, LEAF, j); wait(H2, SD, j - 3);
            on(H2, [&]() { return panel_trsm(j + 2, rest * nb, j); });
            rec(PAN, j, H2);
            wait(H2, TOP, j); wait(H2, SD, j - 2);
            on(H2, [&]() { return update(j, j + 2, rest, j + 1, 1, 0); });
            rec(U1, j, H2);
            wait(H2, SD, j - 1);
            on(H2, [&]() { return update(j, j + 2, rest, j + 2, 1, 1); });
            rec(U2, j, H2);
        }
        // ---- S: bulk update of the columns >= j+3
        if (nblk - (j + 3) > 0) {
            wait(S, PAN, j);
            on(S, [&]() { return update(j, j + 3, nblk - (j + 3), j + 3, nblk - (j + 3), 1); });
            rec(SD, j, S);
        }
    }
    // join: S continues only after both chains have finished -- also on error (see potrf_la)
    if (cudaEventRecord(ev0, H) == cudaSuccess) cudaStreamWaitEvent(S, ev0, 0);
    if (cudaEventRecord(ev1, H2) == cudaSuccess) cudaStreamWaitEvent(S, ev1, 0);
    h->stream = S;
    if (rc == 0 && sub)   // all leaf inverses in ONE batched launch (nblk CTAs, ~27 us)
        rc = leaf(h, A, lda, dinv, goff, (int)nblk, (int64_t)nb * (lda + 1), (int64_t)LT * LT, 2);
    return rc;   // no host synchronisation; EventSet releases the events
}


// The look-ahead factorisation with GROUPED bulk updates (the single-GPU twin of the multi-GPU driver's scheme): 128-wide
// panels for the chain, but the bulk of the trailing matrix is updated once per group of G panels with K = G*128, where the
// DMMA GEMM runs near its large-K rate (a K = 128 update reaches ~25 TF).  For panel j of group g = j / G:
//   H  (chain, high priority): TOP(j): block (j+1, j) <- . L_jj^-T by substitution;  [j+1 in the same group:] diagonal
//                              block (j+1, j+1) -= L(j+1,j) L(j+1,j)^T, factor it (no inverse)
//   H2 (column work)         : PAN(j): rows >= j+2 of column j by substitution;  U1: column j+1, U2: column j+2,
//                              E: columns j+3 .. end of the group (all K = 128, inside the group only)
//   S  (bulk)                : when the group is complete  A(g): the NEXT group's columns (K = G*128; the chain of group
//                              g+1 waits for it), then B(g): all columns beyond the next group (K = G*128) -- B(g) overlaps
//                              with the chain of group g+1.
// Every block receives the contributions of all earlier panels exactly once (own-group panels eagerly, earlier groups in
// A / B), in a fixed order, so the result does not depend on timing.
int potrf_la_grouped(gpx_ctx* h, double* A, int64_t n, int64_t lda, double* dinv, int goff, int G) {
    constexpr int nb = LT;
    const int64_t nblk = n / nb;
    const int64_t ngrp = (nblk + G - 1) / G;
    cudaStream_t S = h->stream, H = h->aux_stream, H2 = h->aux2_stream;
    enum { LEAF = 0, TOP, PAN, U1, U2, AG, NEV };
    GpxEventSet es;
    GPX_TRY(es.create(NEV * nblk + 2));
    auto E = [&](int kind, int64_t j) -> cudaEvent_t { return es.ev[kind * nblk + j]; };
    int rc = 0;
    auto rec = [&](int kind, int64_t j, cudaStream_t st) { if (rc == 0 && cudaEventRecord(E(kind, j), st) != cudaSuccess) rc = GPX_E_CUDA; };
    auto wait = [&](cudaStream_t st, int kind, int64_t j) { if (rc == 0 && j >= 0 && cudaStreamWaitEvent(st, E(kind, j), 0) != cudaSuccess) rc = GPX_E_CUDA; };
    auto on = [&](cudaStream_t st, auto&& fn) { if (rc != 0) return; h->stream = st; rc = fn(); h->stream = S; };
    cudaEvent_t ev0 = es.ev[NEV * nblk], ev1 = es.ev[NEV * nblk + 1];
    GPX_CUDA(cudaEventRecord(ev0, S));            // everything queued on S so far precedes the factorisation
    GPX_CUDA(cudaStreamWaitEvent(H, ev0, 0));
    GPX_CUDA(cudaStreamWaitEvent(H2, ev0, 0));
    auto blk = [&](int64_t i, int64_t j) -> double* { return A + i * nb * lda + j * nb; };
    // C(block rows r0.., block columns c..c+ncols) -= A(rows r0.., panels j..j+np) * A(block rows c.., panels j..j+np)^T
    auto update = [&](int64_t j, int64_t np, int64_t r0, int64_t rows, int64_t c, int64_t ncols, int lower) -> int {
        if (rows <= 0 || ncols <= 0 || np <= 0) return 0;
        GemmArgs a = base_args();
        a.A = blk(r0, j); a.lda = lda; a.a_kmajor = 1;
        a.B = blk(c, j); a.ldb = lda; a.b_kmajor = 1;
        a.C = blk(r0, c); a.ldc = lda;
        a.M = (int)(rows * nb); a.N = (int)(ncols * nb); a.K = (int)(np * nb);
        a.alpha = -1.0; a.beta = 1.0;
        a.lower_only = lower;
        return gpx_gemm_launch(h, a);
    };
    auto factor_diag = [&](int64_t j) -> int { return leaf(h, blk(j, j), lda, dinv + j * LT * LT, goff + (int)(j * nb), 1, 0, 0, 1); };
    on(H, [&]() { return factor_diag(0); });
    rec(LEAF, 0, H);
    for (int64_t j = 0; j < nblk && rc == 0; ++j) {
        const int64_t g = j / G, gend = std::min<int64_t>((g + 1) * G, nblk);     // this group's columns are [g*G, gend)
        const bool next_in_group = j + 1 < gend;
        // ---- H: block row j+1 of panel j, then (inside the group) the next diagonal block
        if (j + 1 < nblk) {
            if (j > g * G) wait(H, U1, j - 1);                                   // block (j+1, j) has every in-group update
            on(H, [&]() { return trsm_sub(h, blk(j + 1, j), nb, lda, blk(j, j), lda); });
            rec(TOP, j, H);
            if (next_in_group) {
                if (j > g * G) wait(H, U2, j - 1);                               // block (j+1, j+1) likewise (U2 of panel j-1)
                on(H, [&]() {
                    GPX_TRY(update(j, 1, j + 1, 1, j + 1, 1, 1));
                    return factor_diag(j + 1);
                });
                rec(LEAF, j + 1, H);
            }
        }
        // ---- H2: the rest of column j, then its updates of the remaining columns of the group
        const int64_t rest = nblk - (j + 2);
        if (rest > 0) {
            wait(H2, LEAF, j);
            on(H2, [&]() { return trsm_sub(h, blk(j + 2, j), rest * nb, lda, blk(j, j), lda); });
            rec(PAN, j, H2);
            if (next_in_group) {
                wait(H2, TOP, j);
                on(H2, [&]() { return update(j, 1, j + 2, rest, j + 1, 1, 0); });                       // column j+1, rows >= j+2
                rec(U1, j, H2);
                if (j + 2 < gend) {
                    on(H2, [&]() { return update(j, 1, j + 2, rest, j + 2, 1, 1); });                   // column j+2
                    rec(U2, j, H2);
                    if (j + 3 < gend)
                        on(H2, [&]() { return update(j, 1, j + 3, nblk - (j + 3), j + 3, gend - (j + 3), 1); });   // rest of the group
                } else {
                    rec(U2, j, H2);
                }
            }
        }
        // ---- S: the group is complete -> A(g) on the next group's columns, then the deferred bulk B(g)
        if (j + 1 == gend && gend < nblk) {
            const int64_t g0 = g * G, np = gend - g0;
            const int64_t nend = std::min<int64_t>(gend + G, nblk);              // next group's columns [gend, nend)
            if (rest > 0) wait(S, PAN, j);
            wait(S, TOP, j);
            on(S, [&]() { return update(g0, np, gend, nblk - gend, gend, nend - gend, 1); });
            rec(AG, g, S);
            wait(H, AG, g);
            on(H, [&]() { return factor_diag(gend); });
            rec(LEAF, gend, H);
            if (nend < nblk) on(S, [&]() { return update(g0, np, nend, nblk - nend, nend, nblk - nend, 1); });
        }
    }
    (void)ngrp;
    if (cudaEventRecord(ev0, H) == cudaSuccess) cudaStreamWaitEvent(S, ev0, 0);
    if (cudaEventRecord(ev1, H2) == cudaSuccess) cudaStreamWaitEvent(S, ev1, 0);
    h->stream = S;
    if (rc == 0)   // all leaf inverses in ONE batched launch (nblk CTAs, ~27 us)
        rc = leaf(h, A, lda, dinv, goff, (int)nblk, (int64_t)nb * (lda + 1), (int64_t)LT * LT, 2);
    return rc;
}

int g_la_group_forced = -1;
int la_group(int64_t n) {   // panels per group of the grouped look-ahead factorisation (0 / 1: ungrouped potrf_la_split)
    int& forced = g_la_group_forced;
    if (forced < 0) {
        const char* e = getenv("GPX_POTRF_LA_GROUP");
        forced = e ? atoi(e) : 0;
    }
    if (forced > 0) return forced;
    return n <= 2048 ? 1 : (n <= 4096 ? 2 : (n <= 8192 ? 4 : 8));
}

int la_split() {   // 1: potrf_la_split (default), 0: potrf_la
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("GPX_POTRF_LA_SPLIT");
        v = e ? atoi(e) : 1;
    }
    return v;
}

int la_block(int64_t n) {   // panel width of the look-ahead algorithm
    static int forced = -1;
    if (forced < 0) {
        const char* e = getenv("GPX_POTRF_LA_NB");
        forced = e ? atoi(e) : 0;
    }
    if (forced > 0) return forced;
    return la_split() ? 128 : (n <= 8192 ? 128 : 256);   // the substitution chain (potrf_la_split) works on 128-wide panels
}

int la_threshold() {   // largest n factored by the look-ahead algorithm (0 disables it)
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("GPX_POTRF_LA_MAX");
        v = e ? atoi(e) : 16384;
    }
    return v;
}
}  // namespace

// panels per group of the look-ahead factorisation: 0 = by size (default), 1 = ungrouped, G > 1 forces K = G*128 bulk updates
extern "C" int gpx_potrf_set_group(int panels_per_group) {
    GPX_REQUIRE(panels_per_group >= 0 && panels_per_group <= 64, 1);
    g_la_group_forced = panels_per_group;
    return 0;
}

extern "C" int gpx_potrf_async(gpx_handle h, double* A, int64_t n, int64_t lda, double* dinv) {
    GPX_ENTER(h);
    GPX_REQUIRE(A != nullptr && ((uintptr_t)A % 16) == 0, 2);
    GPX_REQUIRE(n > 0 && n % LT == 0, 3);
    GPX_REQUIRE(lda >= n && lda % 2 == 0, 4);
    GPX_REQUIRE(dinv != nullptr, 5);
    GPX_TRY(potrf_rec(h, A, n, lda, dinv, 0));
    const int nt = (int)(n / LT);
    if (nt > 1) {
        zero_upper_tiles_kernel<<<dim3(nt, nt), 256, 0, h->stream>>>(A, lda, nt);
        GPX_CHECK_LAUNCH(h);
    }
    return 0;
}

extern "C" int gpx_potrf(gpx_handle h, double* A, int64_t n, int64_t lda, double* dinv) {
    GPX_ENTER(h);
    GPX_CUDA(cudaMemsetAsync(h->d_info, 0, sizeof(int), h->stream));
    int r = gpx_potrf_async(h, A, n, lda, dinv);
    if (r != 0) return r;
    int info = 0;
    GPX_TRY(gpx_read_info(h, &info));
    if (info > 0) gpx_set_error("gpx_potrf: leading minor of order %d is not positive definite", info);
    return info;
}

extern "C" int gpx_potrf_info(gpx_handle h, int* info_out) {
    GPX_ENTER(h);
    GPX_REQUIRE(info_out != nullptr, 2);
    int info = 0;
    GPX_TRY(gpx_read_info(h, &info));
    GPX_CUDA(cudaMemsetAsync(h->d_info, 0, sizeof(int), h->stream));   // read-and-clear: the flag is sticky across async calls
    *info_out = info;
    if (info > 0) gpx_set_error("gpx_potrf: leading minor of order %d is not positive definite", info);
    return 0;
}

extern "C" int gpx_trsm(gpx_handle h, const double* L, int64_t n, int64_t ldl, const double* dinv, int trans,
                        double* B, int64_t nrhs, int64_t ldb) {
    GPX_ENTER(h);
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

namespace {
// Level-synchronous in-place inverse for a power-of-two tile count: the independent sub-problems of one recursion level
// are ONE strided-batched GEMM launch (two launches per level, 2 log2(n/128) launches in all) while they are small; from
// BIG_BLOCK on, each sub-problem is large enough for its own (TMA-fed) launch.
int trtri_levels(gpx_ctx* h, double* A, int64_t n, int64_t lda, const double* dinv, double* work) {
    const int nt = (int)(n / LT);
    copy_tiles_kernel<<<nt, 256, 0, h->stream>>>(A, lda, (int64_t)LT * (lda + 1), dinv, (int64_t)LT * LT);
    GPX_CHECK_LAUNCH(h);
    constexpr int64_t BIG_BLOCK = 4096;
    for (int64_t hs = LT; hs < n; hs *= 2) {
        const int nsub = (int)(n / (2 * hs));
        const int64_t sub_stride = 2 * hs * (lda + 1);
        const int groups = hs >= BIG_BLOCK ? nsub : 1;          // separate launches for big sub-problems
        const int per = hs >= BIG_BLOCK ? 1 : nsub;
        for (int gi = 0; gi < groups; ++gi) {
            double* base = A + (int64_t)gi * sub_stride;
            double* wbase = work + (int64_t)gi * hs * hs;
            GemmArgs a = base_args();  // T = A21 * A11inv  (k >= column tile)
            a.A = base + hs * lda; a.lda = lda; a.a_kmajor = 1; a.sA = sub_stride;
            a.B = base; a.ldb = lda; a.b_kmajor = 0; a.sB = sub_stride;
            a.C = wbase; a.ldc = hs; a.sC = hs * hs;
            a.M = (int)hs; a.N = (int)hs; a.K = (int)hs;
            a.batch = per;
            a.kb_mode = 2;
            GPX_TRY(gpx_gemm_launch(h, a));
            GemmArgs b = base_args();  // A21 = -A22inv * T  (k <= row tile)
            b.A = base + hs * lda + hs; b.lda = lda; b.a_kmajor = 1; b.sA = sub_stride;
            b.B = wbase; b.ldb = hs; b.b_kmajor = 0; b.sB = hs * hs;
            b.C = base + hs * lda; b.ldc = lda; b.sC = sub_stride;
            b.M = (int)hs; b.N = (int)hs; b.K = (int)hs;
            b.batch = per;
            b.alpha = -1.0;
            b.ke_mode = 1;
            b.rev_rows = 1;
            GPX_TRY(gpx_gemm_launch(h, b));
        }
    }
    return 0;
}
}  // namespace

extern "C" int gpx_trtri(gpx_handle h, double* L, int64_t n, int64_t ldl, const double* dinv, double* work) {
    GPX_ENTER(h);
    GPX_REQUIRE(n > 0 && n % LT == 0, 3);
    GPX_REQUIRE(n == LT || work != nullptr, 6);
    const int64_t nt = n / LT;
    if (nt > 1 && (nt & (nt - 1)) == 0) return trtri_levels(h, L, n, ldl, dinv, work);
    return trtri_rec(h, L, n, ldl, dinv, work);
}

extern "C" int gpx_lauum(gpx_handle h, const double* Linv, int64_t n, int64_t ldl, double* out, int64_t ldo) {
    GPX_ENTER(h);
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

// One whole panel of the multi-GPU driver (nb columns, `rows` rows from the diagonal block down, leading dimension ld):
// leaf by leaf -- factor the 128 x 128 diagonal leaf WITHOUT its inverse, solve everything below it by substitution
// (warp per row), update the remaining columns of the panel (K = 128) -- and the leaf inverses in one batched launch at the
// end.  Replaces potrf(nb) + a GEMM-based TRSM against explicit inverses (three K = 128 launches over the full height).
int gpx_panel_factor_sub(gpx_ctx* h, double* P, int64_t rows, int64_t ld, int nb, double* dinv, int goff) {
    const int tpb = nb / LT;
    for (int t = 0; t < tpb; ++t) {
        double* D = P + (int64_t)t * LT * ld + t * LT;                       // diagonal leaf t
        GPX_TRY(leaf(h, D, ld, dinv + (int64_t)t * LT * LT, goff + t * LT, 1, 0, 0, 1));
        const int64_t below = rows - (int64_t)(t + 1) * LT;
        if (below <= 0) continue;
        GPX_TRY(trsm_sub(h, D + (int64_t)LT * ld, below, ld, D, ld));
        if (t + 1 < tpb) {
            GemmArgs a = base_args();  // columns (t+1)*128 .. nb of the panel -= P_t P_t^T (rows below leaf t)
            a.A = D + (int64_t)LT * ld; a.lda = ld; a.a_kmajor = 1;
            a.B = a.A; a.ldb = ld; a.b_kmajor = 1;
            a.C = D + (int64_t)LT * ld + LT; a.ldc = ld;
            a.M = (int)below; a.N = (tpb - 1 - t) * LT; a.K = LT;
            a.alpha = -1.0; a.beta = 1.0;
            a.lower_only = 1;
            GPX_TRY(gpx_gemm_launch(h, a));
        }
    }
    return leaf(h, P, ld, dinv, goff, tpb, (int64_t)LT * (ld + 1), (int64_t)LT * LT, 2);
}

namespace {
// Left lower TRSM  L X = B  where only a *prefix* of B's columns is non-zero in any given row range: for rows
// below global row r (relative to this sub-problem's row 0 at global row `grow0`), the non-zero columns are the
// local blocks whose global block index is <= block(r): count = prefix(r).  Used for L^-1 on block-cyclic columns.
struct PrefixMap {
    int P, p, nb, snake;
    // optional explicit inverses of the bs x bs diagonal blocks of L (gpx_block_inverses) + a bs x ncols scratch: the
    // recursion then stops at bs (one triangular GEMM per block instead of ~15 latency-bound launches per 1024 rows)
    const double* D = nullptr; int bs = 0; double* tmp = nullptr;
};
inline int64_t prefix_cols(const PrefixMap& pm, int64_t grow_end) {
    // number of local columns (elements) whose global block start is < grow_end
    const int64_t nblk_below = (grow_end + pm.nb - 1) / pm.nb;              // global blocks 0..nblk_below-1 start below grow_end
    return gpx_cyc_count_below(nblk_below, pm.P, pm.p, pm.snake) * pm.nb;   // those owned by rank p
}
int trsm_left_prefix(gpx_ctx* h, const double* L, int64_t n, int64_t ldl, const double* dinv, double* B, int64_t ldb,
                     int64_t grow0, const PrefixMap& pm) {
    const int64_t ncols = prefix_cols(pm, grow0 + n);   // columns that can be non-zero within these rows
    if (ncols <= 0) return 0;
    if (pm.D && n == pm.bs && grow0 % pm.bs == 0) {
        GemmArgs a = base_args();   // tmp = D_b B (k <= row tile), copied back
        a.A = pm.D + (grow0 / pm.bs) * (int64_t)pm.bs * pm.bs; a.lda = pm.bs; a.a_kmajor = 1;
        a.B = B; a.ldb = ldb; a.b_kmajor = 0;
        a.C = pm.tmp; a.ldc = ncols;
        a.M = pm.bs; a.N = (int)ncols; a.K = pm.bs;
        a.ke_mode = 1; a.rev_rows = 1;
        GPX_TRY(gpx_gemm_launch(h, a));
        GPX_CUDA(cudaMemcpy2DAsync(B, ldb * sizeof(double), pm.tmp, ncols * sizeof(double), ncols * sizeof(double), pm.bs,
                                   cudaMemcpyDeviceToDevice, h->stream));
        return 0;
    }
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
        a.kb_mode = 3; a.cyc_P = pm.P; a.cyc_p = pm.p; a.cyc_snake = pm.snake; a.cyc_tpb = pm.nb / LT; a.cyc_q0 = 0; a.cyc_row_base = (int)grow0;
        a.cyc_b_rows = 0;
        GPX_TRY(gpx_gemm_launch(h, a));
    }
    return trsm_left_prefix(h, L + h1 * ldl + h1, h2, ldl, dinv + (h1 / LT) * LT * LT, B + h1 * ldb, ldb, grow0 + h1, pm);
}
}  // namespace

int gpx_trsm_left_prefix_block(gpx_ctx* h, const double* L, int64_t n, int64_t ldl, const double* dinv, double* B, int64_t ldb,
                               int P, int p, int nb, int snake, const double* D, int bs, double* tmp) {
    PrefixMap pm{P, p, nb, snake};
    if (D && tmp && bs > LT) { pm.D = D; pm.bs = bs; pm.tmp = tmp; }
    return trsm_left_prefix(h, L, n, ldl, dinv, B, ldb, 0, pm);
}

namespace {
// Backward counterpart:  L^T Z = X  restricted to the entries on/below each column block's own diagonal block (row >= start
// of the column's global block) -- all a symmetric result needs.  Back substitution produces the bottom rows first and
// rows >= r of Z depend only on rows >= r of X, so the restriction is exact; rows of this sub-problem use the local
// columns whose global block starts below the end of the row range (the same prefix structure), and output tiles that
// lie entirely above a column's block start are skipped (lower_only with the block-cyclic column map).  N^3/(3P) flops.
int trsm_left_prefix_trans(gpx_ctx* h, const double* L, int64_t n, int64_t ldl, const double* dinv, double* B, int64_t ldb,
                           int64_t grow0, const PrefixMap& pm) {
    const int64_t ncols = prefix_cols(pm, grow0 + n);
    if (ncols <= 0) return 0;
    if (pm.D && n == pm.bs && grow0 % pm.bs == 0) {
        GemmArgs a = base_args();   // tmp = D_b^T X (k >= row tile), copied back
        a.A = pm.D + (grow0 / pm.bs) * (int64_t)pm.bs * pm.bs; a.lda = pm.bs; a.a_kmajor = 0;
        a.B = B; a.ldb = ldb; a.b_kmajor = 0;
        a.C = pm.tmp; a.ldc = ncols;
        a.M = pm.bs; a.N = (int)ncols; a.K = pm.bs;
        a.kb_mode = 1;
        GPX_TRY(gpx_gemm_launch(h, a));
        GPX_CUDA(cudaMemcpy2DAsync(B, ldb * sizeof(double), pm.tmp, ncols * sizeof(double), ncols * sizeof(double), pm.bs,
                                   cudaMemcpyDeviceToDevice, h->stream));
        return 0;
    }
    if (n == LT) {
        GemmArgs a = base_args();  // Z = Dinv^T X (in place; a CTA reads exactly the column range it writes)
        a.A = dinv; a.lda = LT; a.a_kmajor = 0;
        a.B = B; a.ldb = ldb; a.b_kmajor = 0;
        a.C = B; a.ldc = ldb;
        a.M = LT; a.N = (int)ncols; a.K = LT;
        return gpx_gemm_launch(h, a);
    }
    const int64_t h1 = half_tiles(n), h2 = n - h1;
    GPX_TRY(trsm_left_prefix_trans(h, L + h1 * ldl + h1, h2, ldl, dinv + (h1 / LT) * LT * LT, B + h1 * ldb, ldb, grow0 + h1, pm));
    const int64_t nc1 = prefix_cols(pm, grow0 + h1);   // columns that have rows in the top half
    if (nc1 > 0) {
        GemmArgs a = base_args();  // X1[:, :nc1] -= L21^T Z2[:, :nc1]
        a.A = L + h1 * ldl; a.lda = ldl; a.a_kmajor = 0;
        a.B = B + h1 * ldb; a.ldb = ldb; a.b_kmajor = 0;
        a.C = B; a.ldc = ldb;
        a.M = (int)h1; a.N = (int)nc1; a.K = (int)h2;
        a.alpha = -1.0; a.beta = 1.0;
        a.lower_only = 1;   // rows above a column's block start are never needed
        a.cyc_P = pm.P; a.cyc_p = pm.p; a.cyc_snake = pm.snake; a.cyc_tpb = pm.nb / LT; a.cyc_q0 = 0; a.cyc_row_base = (int)grow0; a.cyc_b_rows = 0;
        GPX_TRY(gpx_gemm_launch(h, a));
    }
    return trsm_left_prefix_trans(h, L, h1, ldl, dinv, B, ldb, grow0, pm);
}
}  // namespace

int gpx_trsm_left_prefix_trans_block(gpx_ctx* h, const double* L, int64_t n, int64_t ldl, const double* dinv, double* B,
                                     int64_t ldb, int P, int p, int nb, int snake, const double* D, int bs, double* tmp) {
    PrefixMap pm{P, p, nb, snake};
    if (D && tmp && bs > LT) { pm.D = D; pm.bs = bs; pm.tmp = tmp; }
    return trsm_left_prefix_trans(h, L, n, ldl, dinv, B, ldb, 0, pm);
}

extern "C" int gpx_debug_leaf_cycles(long long* out4) {
    return cudaMemcpyFromSymbol(out4, g_leaf_dbg, 4 * sizeof(long long)) == cudaSuccess ? 0 : GPX_E_CUDA;
}

// ---- block inverses: explicit inverses of the bs x bs diagonal blocks of L (bs = 128 * 2^k <= 1024) ---------------
// They shorten the serial chain of every triangular solve by bs/128: a blocked TRSV / TRSM then has n/bs sequential
// steps, each a full-width GEMV / GEMM with an explicit block inverse.  Built from the 128-leaf inverses with the
// level-synchronous scheme of trtri_levels, batched over (sub-problem, block).
namespace {
__global__ void place_leaf_inverses_kernel(double* __restrict__ D, int bs, const double* __restrict__ dinv) {
    const int t = blockIdx.x, lpb = bs / LT;               // leaf index, leaves per block
    double* dst = D + (int64_t)(t / lpb) * bs * bs + (int64_t)(t % lpb) * LT * (bs + 1);
    const double* src = dinv + (int64_t)t * LT * LT;
    for (int idx = threadIdx.x; idx < LT * LT; idx += blockDim.x) dst[(int64_t)(idx >> 7) * bs + (idx & 127)] = src[idx];
}
}  // namespace

extern "C" int gpx_block_size_for(int64_t n) {
    for (int bs = 1024; bs > LT; bs >>= 1)
        if (n % bs == 0) return bs;
    return LT;
}

// D: (n/bs) blocks of bs x bs doubles; work: n*bs/4 doubles.
extern "C" int gpx_block_inverses(gpx_handle h, const double* L, int64_t n, int64_t ldl, const double* dinv, int bs, double* D,
                                  double* work) {
    GPX_ENTER(h);
    GPX_REQUIRE(n > 0 && n % bs == 0 && bs >= LT && bs % LT == 0 && (bs & (bs - 1)) == 0, 6);
    const int nB = (int)(n / bs);
    GPX_CUDA(cudaMemsetAsync(D, 0, (size_t)n * bs * sizeof(double), h->stream));
    place_leaf_inverses_kernel<<<(unsigned)(n / LT), 256, 0, h->stream>>>(D, bs, dinv);
    GPX_CHECK_LAUNCH(h);
    for (int64_t hs = LT; hs < bs; hs *= 2) {
        const int per = (int)(bs / (2 * hs));
        GemmArgs a = base_args();   // T = L21 * X11   (k >= column tile)
        a.A = L + hs * ldl; a.lda = ldl; a.a_kmajor = 1; a.sA = 2 * hs * (ldl + 1); a.sA2 = (int64_t)bs * (ldl + 1);
        a.B = D; a.ldb = bs; a.b_kmajor = 0; a.sB = 2 * hs * (bs + 1); a.sB2 = (int64_t)bs * bs;
        a.C = work; a.ldc = hs; a.sC = hs * hs; a.sC2 = (int64_t)per * hs * hs;
        a.M = (int)hs; a.N = (int)hs; a.K = (int)hs;
        a.batch = per; a.batch2 = nB;
        a.kb_mode = 2;
        GPX_TRY(gpx_gemm_launch(h, a));
        GemmArgs b = base_args();   // X21 = -X22 * T   (k <= row tile)
        b.A = D + hs * (bs + 1); b.lda = bs; b.a_kmajor = 1; b.sA = 2 * hs * (bs + 1); b.sA2 = (int64_t)bs * bs;
        b.B = work; b.ldb = hs; b.b_kmajor = 0; b.sB = hs * hs; b.sB2 = (int64_t)per * hs * hs;
        b.C = D + hs * bs; b.ldc = bs; b.sC = 2 * hs * (bs + 1); b.sC2 = (int64_t)bs * bs;
        b.M = (int)hs; b.N = (int)hs; b.K = (int)hs;
        b.batch = per; b.batch2 = nB;
        b.alpha = -1.0;
        b.ke_mode = 1;
        b.rev_rows = 1;
        GPX_TRY(gpx_gemm_launch(h, b));
    }
    return 0;
}

namespace {
// x <- L^-1 x / L^-T x with bs-block inverses; tmp: bs doubles
int trsv_big_rec(gpx_ctx* h, const double* L, int64_t n, int64_t ldl, const double* D, int bs, int trans, double* x, double* tmp) {
    if (n == bs) {
        GPX_TRY(gpx_gemv(h, trans, bs, bs, 1.0, D, bs, x, 0.0, tmp));
        GPX_CUDA(cudaMemcpyAsync(x, tmp, (size_t)bs * sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
        return 0;
    }
    const int64_t h1 = ((n / bs) / 2) * bs, h2 = n - h1;
    const double* L21 = L + h1 * ldl;
    const double* L22 = L21 + h1;
    const double* D2 = D + (h1 / bs) * (int64_t)bs * bs;
    if (!trans) {
        GPX_TRY(trsv_big_rec(h, L, h1, ldl, D, bs, 0, x, tmp));
        GPX_TRY(gpx_gemv(h, 0, h2, h1, -1.0, L21, ldl, x, 1.0, x + h1));
        return trsv_big_rec(h, L22, h2, ldl, D2, bs, 0, x + h1, tmp);
    }
    GPX_TRY(trsv_big_rec(h, L22, h2, ldl, D2, bs, 1, x + h1, tmp));
    GPX_TRY(gpx_gemv(h, 1, h2, h1, -1.0, L21, ldl, x + h1, 1.0, x));
    return trsv_big_rec(h, L, h1, ldl, D, bs, 1, x, tmp);
}

// B <- L^-1 B / L^-T B (B is n x m, row-major); tmp: bs * m doubles
int trsm_big_rec(gpx_ctx* h, const double* L, int64_t n, int64_t ldl, const double* D, int bs, int trans, double* B, int64_t m,
                 int64_t ldb, double* tmp) {
    if (n == bs) {
        GemmArgs a = base_args();   // tmp = D B  or  D^T B  (triangular k range), then copied back
        a.A = D; a.lda = bs; a.a_kmajor = trans ? 0 : 1;
        a.B = B; a.ldb = ldb; a.b_kmajor = 0;
        a.C = tmp; a.ldc = m;
        a.M = bs; a.N = (int)m; a.K = bs;
        if (trans) a.kb_mode = 1; else { a.ke_mode = 1; a.rev_rows = 1; }
        GPX_TRY(gpx_gemm_launch(h, a));
        GPX_CUDA(cudaMemcpy2DAsync(B, ldb * sizeof(double), tmp, m * sizeof(double), m * sizeof(double), bs,
                                   cudaMemcpyDeviceToDevice, h->stream));
        return 0;
    }
    const int64_t h1 = ((n / bs) / 2) * bs, h2 = n - h1;
    const double* L21 = L + h1 * ldl;
    const double* L22 = L21 + h1;
    const double* D2 = D + (h1 / bs) * (int64_t)bs * bs;
    double* B2 = B + h1 * ldb;
    if (!trans) {
        GPX_TRY(trsm_big_rec(h, L, h1, ldl, D, bs, 0, B, m, ldb, tmp));
        GemmArgs a = base_args();   // B2 -= L21 X1
        a.A = L21; a.lda = ldl; a.a_kmajor = 1;
        a.B = B; a.ldb = ldb; a.b_kmajor = 0;
        a.C = B2; a.ldc = ldb;
        a.M = (int)h2; a.N = (int)m; a.K = (int)h1;
        a.alpha = -1.0; a.beta = 1.0;
        GPX_TRY(gpx_gemm_launch(h, a));
        return trsm_big_rec(h, L22, h2, ldl, D2, bs, 0, B2, m, ldb, tmp);
    }
    GPX_TRY(trsm_big_rec(h, L22, h2, ldl, D2, bs, 1, B2, m, ldb, tmp));
    GemmArgs a = base_args();   // B1 -= L21^T X2
    a.A = L21; a.lda = ldl; a.a_kmajor = 0;
    a.B = B2; a.ldb = ldb; a.b_kmajor = 0;
    a.C = B; a.ldc = ldb;
    a.M = (int)h1; a.N = (int)m; a.K = (int)h2;
    a.alpha = -1.0; a.beta = 1.0;
    GPX_TRY(gpx_gemm_launch(h, a));
    return trsm_big_rec(h, L, h1, ldl, D, bs, 1, B, m, ldb, tmp);
}
}  // namespace

extern "C" int gpx_trsv_big(gpx_handle h, const double* L, int64_t n, int64_t ldl, const double* D, int bs, int trans, double* x,
                            double* tmp) {
    GPX_ENTER(h);
    GPX_REQUIRE(n > 0 && n % bs == 0, 3);
    return trsv_big_rec(h, L, n, ldl, D, bs, trans, x, tmp);
}

extern "C" int gpx_trsm_big(gpx_handle h, const double* L, int64_t n, int64_t ldl, const double* D, int bs, int trans, double* B,
                            int64_t nrhs, int64_t ldb, double* tmp) {
    GPX_ENTER(h);
    GPX_REQUIRE(n > 0 && n % bs == 0, 3);
    GPX_REQUIRE(nrhs > 0 && nrhs % LT == 0, 9);
    return trsm_big_rec(h, L, n, ldl, D, bs, trans, B, nrhs, ldb, tmp);
}
