// Fused small-problem posterior: the whole of GP_regression.py:109-156 (== tune...:67-101, CO2...:182-214) for
// N <= 128 training and n <= 128 test points in ONE kernel launch of ONE thread block.
//
// At the sizes the reference ships (N = 5, n = 100) the tiled path is nothing but launch latency and host
// round trips (~60 launches, 6 synchronisations, 0.72 ms against 0.16 ms for NumPy).  Here every matrix lives in
// shared memory (packed lower triangles for K/L and for the posterior covariance, the N x (n+1) block [K_s | y] in
// full), the host stages all inputs in one pinned buffer (one H2D copy), and all results come back in one D2H copy:
//
//   K = k(X,X) + s I           -> L (packed, in place)                      GP_regression.py:125-138
//   [V | m] = L^-1 [k(X,X*) | y]                                             :139,144
//   mu = V^T m (== K_s^T alpha, alpha = L^-T m), var = k** - colsum V^2      :140-148
//   LML = -1/2 m^T m - sum log L_kk - N/2 log 2 pi                           tune...:141
//   C = k(X*,X*) + jitter I - V^T V -> L_ (packed, in place)                 :154
//   f_post = mu + L_ Z   (Z: host-drawn normals, global NumPy RNG)           :155
//
// Both factorisations and the forward substitution are written so that a column step needs ONE block barrier:
// the elimination is square-root free (A[i][j] -= A[i][k] A[j][k] / d_k never rewrites column k; the reciprocal
// of the next pivot is produced by the thread that finishes it) and rows / columns are scaled in a final pass.
// The per-entry covariance code is the one the tiled builder uses (cov_eval.cuh), with the same accumulation order
// over the input dimension, so K entries are bit-identical to gpx_cov_build's.
#include "common.cuh"
#include "cov_eval.cuh"

namespace {

using namespace gpx_cov;

constexpr int ST = 1024;     // threads of the single block, viewed as a 32 x 32 grid (tx, ty)
constexpr int SMAX = 128;    // largest N and n served: 4 x 4 sub-tiles of 32 x 32 entries
static_assert(ST == 32 * 32 && SMAX == 4 * 32, "lane tx covers the columns tx + 32b, b < 4; 32 warps cover 128 rows");

__device__ __forceinline__ int tri(int i, int j) { return i * (i + 1) / 2 + j; }  // packed lower triangle, j <= i

// One covariance entry.  Deliberately NOT inlined and called from rolled loops: the kernel runs every code line
// once or a few times, so its cost is instruction fetch -- 16 inlined copies of exp/sin/pow made it 200 KB of SASS.
template <int KIND>
__device__ __noinline__ double pair_value(const CovParams& p, const double* __restrict__ a,
                                          const double* __restrict__ b, bool diag) {
    double acc = 0.0, aux = 0.0;
    for (int d = 0; d < p.D; ++d) {
        if (KIND == GPX_COV_LIN) {
            acc += (a[d] - p.th[0]) * (b[d] - p.th[0]);
            aux += a[d] + b[d];
        } else {
            const double df = a[d] - b[d];
            acc += df * df;
        }
    }
    double dk[1];
    return cov_eval<KIND, false>(p, acc, aux, diag, dk);
}

// The same entry together with d/dtheta_q for every hyper-parameter (dk: NTheta<KIND> doubles).
template <int KIND>
__device__ __noinline__ double pair_value_dk(const CovParams& p, const double* __restrict__ a,
                                             const double* __restrict__ b, bool diag, double* dk) {
    double acc = 0.0, aux = 0.0;
    for (int d = 0; d < p.D; ++d) {
        if (KIND == GPX_COV_LIN) {
            acc += (a[d] - p.th[0]) * (b[d] - p.th[0]);
            aux += a[d] + b[d];
        } else {
            const double df = a[d] - b[d];
            acc += df * df;
        }
    }
    return cov_eval<KIND, true>(p, acc, aux, diag, dk);
}

// In-place Cholesky of a packed lower triangle in shared memory.  rdiag / rinv receive L_kk and 1 / L_kk; pinv is
// scratch for the pivot reciprocals.  Returns 0, or the 1-based index of the first non-positive pivot (uniform
// over the block).  A column step is bound by instruction issue (every warp walks the step's control code), so
// only CW warps take part and the rows below the pivot are dealt to them cyclically: warp w owns rows
// k+1+w, k+1+w+CW, ..., lane tx the columns tx + 32b.
constexpr int CW = 16;
__device__ int chol_packed(double* A, int n, double* rdiag, double* rinv, double* pinv) {
    const int tx = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (threadIdx.x == 0) pinv[0] = 1.0 / A[0];
    for (int k = 0; k < n; ++k) {
        __syncthreads();
        const double d = A[tri(k, k)];
        if (!(d > 0.0)) return k + 1;
        if (w < CW) {
            const double pk = pinv[k];
            double cj[4];
#pragma unroll
            for (int b = 0; b < 4; ++b) {
                const int j = tx + 32 * b;
                cj[b] = (j > k && j < n) ? A[tri(j, k)] * pk : 0.0;
            }
            int i = k + 1 + w;
            if (w == 0 && i < n) {
                // row k+1 first: it finishes the next pivot, whose reciprocal (a long dependent chain) then
                // overlaps the remaining rows of this warp
                const int bi = i * (i + 1) / 2;
                const double aik = A[bi + k];
#pragma unroll
                for (int b = 0; b < 4; ++b) {
                    const int j = tx + 32 * b;
                    if (j > k && j <= i) A[bi + j] -= aik * cj[b];
                }
                if (tx == (i & 31)) pinv[i] = 1.0 / A[bi + i];   // this lane wrote A[i][i] itself
                i += CW;
            }
#pragma unroll 4
            for (; i < n; i += CW) {
                const int bi = i * (i + 1) / 2;
                const double aik = A[bi + k];
#pragma unroll
                for (int b = 0; b < 4; ++b) {
                    const int j = tx + 32 * b;
                    if (j > k && j <= i) A[bi + j] -= aik * cj[b];
                }
            }
        }
    }
    __syncthreads();
    for (int k = threadIdx.x; k < n; k += ST) {
        const double r = sqrt(A[tri(k, k)]);
        rdiag[k] = r;
        rinv[k] = 1.0 / r;
    }
    __syncthreads();
#pragma unroll 1
    for (int i = w; i < n; i += ST / 32) {
        const int bi = i * (i + 1) / 2;
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            const int j = tx + 32 * b;
            if (j < i) A[bi + j] *= rinv[j];
            else if (j == i) A[bi + i] = rdiag[i];
        }
    }
    __syncthreads();
    return 0;
}

// in : [X N*D][y N][Xs n*D][Z n*nf]      out : [mu n][var n][fpost n*nf][lml, info_train, info_post, 0]
// Block b serves the test points [128 b, 128 b + 128) (every block refactors the tiny K itself); the posterior
// factor needs the whole posterior covariance in one block, so it is offered for n_total <= 128 only.
// mode 0: moments only; 1: moments + f_post from Z; 2: moments, and [L_ packed | mu] is left in `keep`.
template <int KIND>
__global__ void __launch_bounds__(ST, 1) gp_small_kernel(const __grid_constant__ CovParams p, int N, int n_total, int nf, int mode, double s,
                                                         double jitter, const double* __restrict__ in,
                                                         double* __restrict__ out, double* __restrict__ keep) {
    extern __shared__ double sm[];
    const int tid = threadIdx.x, tx = tid & 31, ty = tid >> 5;
    const int D = p.D;
    const int c0 = blockIdx.x * SMAX;
    const int n = (n_total - c0) < SMAX ? (n_total - c0) : SMAX;
    const int nv = n + 1;                          // columns of [K_s | y]
    const bool lead = blockIdx.x == 0;
    const double* X = in;
    const double* y = X + (size_t)N * D;
    const double* Xs = y + N + (size_t)c0 * D;
    const double* Z = y + N + (size_t)n_total * D;
    double* o_mu = out + c0;
    double* o_var = out + n_total + c0;
    double* o_f = out + 2 * (size_t)n_total;
    double* o_s = o_f + (size_t)n_total * nf;

    const int big = N > n ? N : n;
    double* Lp = sm;                               // packed L (N), later the packed posterior factor (n)
    double* Vs = Lp + big * (big + 1) / 2;         // N x nv, row-major: [K_s | y] -> [V | m]
    double* mus = Vs + N * nv;                     // mu
    double* rdiag = mus + n;                       // L_kk
    double* rinv = rdiag + big;                    // 1 / L_kk
    double* pinv = rinv + big;                     // pivot reciprocals (chol_packed scratch)

    // ---- K + s I (packed lower) and [K_s | y]
#pragma unroll 1
    for (int ab = 0; ab < 16; ++ab) {
        const int i = ty + 32 * (ab >> 2), j = tx + 32 * (ab & 3);
        if (i < N && j <= i) {
            double v = pair_value<KIND>(p, X + (size_t)i * D, X + (size_t)j * D, i == j);
            if (i == j) v += s;
            Lp[tri(i, j)] = v;
        }
    }
#pragma unroll 1
    for (int i = ty; i < N; i += 32) {
#pragma unroll 1
        for (int c = tx; c < n; c += 32)
            Vs[i * nv + c] = pair_value<KIND>(p, X + (size_t)i * D, Xs + (size_t)c * D,
                                              KIND == GPX_COV_CO2 && N == n_total && i == c);
        if (tx == 0) Vs[i * nv + n] = y[i];
    }
    const int info1 = chol_packed(Lp, N, rdiag, rinv, pinv);
    if (info1) {
        if (tid == 0 && lead) { o_s[0] = 0.0; o_s[1] = info1; o_s[2] = 0.0; o_s[3] = 0.0; }
        return;
    }
    // ---- [V | m] = L^-1 [K_s | y]: row k stays unscaled while rows below consume it (one barrier per step)
    for (int k = 0; k < N; ++k) {
        __syncthreads();
        const double rk = rinv[k];
#pragma unroll 1
        for (int i = k + 1 + ty; i < N; i += ST / 32) {
            const double lik = Lp[tri(i, k)] * rk;
            for (int c = tx; c < nv; c += 32) Vs[i * nv + c] -= lik * Vs[k * nv + c];
        }
    }
    __syncthreads();
#pragma unroll 1
    for (int i = ty; i < N; i += ST / 32)
        for (int c = tx; c < nv; c += 32) Vs[i * nv + c] *= rinv[i];
    __syncthreads();
#pragma unroll 1
    for (int c = tid; c < n; c += ST) {
        double m = 0.0, ss = 0.0;
#pragma unroll 4
        for (int i = 0; i < N; ++i) {
            const double v = Vs[i * nv + c];
            m += v * Vs[i * nv + n];                                           // GP_regression.py:143
            ss += v * v;
        }
        mus[c] = m;
        o_mu[c] = m;
        const double* xc = Xs + (size_t)c * D;
        o_var[c] = pair_value<KIND>(p, xc, xc, true) - ss;                     // :147
    }
    if (lead && tid >= ST - 32) {   // last warp: LML (tune...:141)
        double part = 0.0, lg = 0.0;
        for (int i = tx; i < N; i += 32) {
            const double m = Vs[i * nv + n];
            part += m * m;
            lg += log(rdiag[i]);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            part += __shfl_xor_sync(0xffffffffu, part, o);
            lg += __shfl_xor_sync(0xffffffffu, lg, o);
        }
        if (tx == 0) o_s[0] = -.5 * part - lg - N / 2.0 * log(2.0 * PI_D);
    }
    if (mode == 0) {
        if (tid == 0 && lead) { o_s[1] = 0.0; o_s[2] = 0.0; o_s[3] = 0.0; }
        return;
    }
    __syncthreads();   // every read of L and of its diagonal is done: the storage becomes the posterior covariance
#pragma unroll 1
    for (int ab = 0; ab < 16; ++ab) {
        const int i = ty + 32 * (ab >> 2), j = tx + 32 * (ab & 3);
        if (i < n && j <= i) {
            double dot = 0.0;
#pragma unroll 4
            for (int r = 0; r < N; ++r) dot += Vs[r * nv + i] * Vs[r * nv + j];
            double v = pair_value<KIND>(p, Xs + (size_t)i * D, Xs + (size_t)j * D, i == j);
            if (i == j) v += jitter;
            Lp[tri(i, j)] = v - dot;                                           // :154
        }
    }
    __syncthreads();
    const int info2 = chol_packed(Lp, n, rdiag, rinv, pinv);
    if (info2) {
        if (tid == 0) { o_s[1] = 0.0; o_s[2] = info2; o_s[3] = 0.0; }
        return;
    }
    if (mode == 1) {
        for (int t = tid; t < n * nf; t += ST) {
            const int a = t / nf, f = t - a * nf;
            double acc = 0.0;
            for (int b = 0; b <= a; ++b) acc += Lp[tri(a, b)] * Z[(size_t)b * nf + f];
            o_f[t] = mus[a] + acc;                                             // :155
        }
    } else {
        const int np_ = n * (n + 1) / 2;
        for (int t = tid; t < np_; t += ST) keep[t] = Lp[t];
        for (int t = tid; t < n; t += ST) keep[np_ + t] = mus[t];
    }
    if (tid == 0) { o_s[1] = 0.0; o_s[2] = 0.0; o_s[3] = 0.0; }
}

// Prior factor (GP_regression.py:71-92): chol(k(X*,X*) + s I) for n <= 128, left in `keep` as [L packed | 0] so that
// gp_small_sample_kernel forms L z.  out[0] = 1-based index of a non-positive pivot or 0.
template <int KIND>
__global__ void __launch_bounds__(ST, 1) gp_small_prior_kernel(const __grid_constant__ CovParams p, int n, double s,
                                                               const double* __restrict__ Xs, double* __restrict__ keep,
                                                               double* __restrict__ out) {
    extern __shared__ double sm[];
    const int tid = threadIdx.x, tx = tid & 31, ty = tid >> 5;
    const int D = p.D;
    double* Lp = sm;
    double* rdiag = Lp + n * (n + 1) / 2;
    double* rinv = rdiag + n;
    double* pinv = rinv + n;
#pragma unroll 1
    for (int ab = 0; ab < 16; ++ab) {
        const int i = ty + 32 * (ab >> 2), j = tx + 32 * (ab & 3);
        if (i < n && j <= i) {
            double v = pair_value<KIND>(p, Xs + (size_t)i * D, Xs + (size_t)j * D, i == j);
            if (i == j) v += s;
            Lp[tri(i, j)] = v;
        }
    }
    const int info = chol_packed(Lp, n, rdiag, rinv, pinv);
    if (tid == 0) out[0] = info;
    if (info) return;
    const int np_ = n * (n + 1) / 2;
    for (int t = tid; t < np_; t += ST) keep[t] = Lp[t];
    for (int t = tid; t < n; t += ST) keep[np_ + t] = 0.0;
}

// f_post = mu + L_ Z from the factor kept by a mode-2 launch (GP_regression.py:155).
__global__ void __launch_bounds__(ST) gp_small_sample_kernel(int n, int nf, const double* __restrict__ keep,
                                                             const double* __restrict__ Z, double* __restrict__ fpost) {
    const double* mu = keep + n * (n + 1) / 2;
    for (int t = blockIdx.x * ST + threadIdx.x; t < n * nf; t += gridDim.x * ST) {
        const int a = t / nf, f = t - a * nf;
        double acc = 0.0;
        for (int b = 0; b <= a; ++b) acc += keep[tri(a, b)] * Z[(size_t)b * nf + f];
        fpost[t] = mu[a] + acc;
    }
}

// LML and dLML/dtheta for N <= 128 in one block, optionally the reference's whole gradient-ascent loop on the SE
// length-scale (tune_hyperparms_regression.py:121-153) without leaving the kernel:
//   K + s I -> L;  [L^-1 | m] = L^-1 [I | y];  alpha = L^-T m;  LML (tune...:141);
//   grad_q = 1/2 sum_ij (alpha_i alpha_j - (L^-T L^-1)_ij) dK_ij/dtheta_q            (tune...:55-57,144)
//   ascent: l <- l + step * grad_l, stop once |LML - LML_old| <= tol (the step is still taken), at most max_iter.
// in : [X N*D][y N]     out : [lml, info, iters, l_final, l_used, error, converged, 0, grad[11]]
template <int KIND>
__global__ void __launch_bounds__(ST, 1) gp_small_grad_kernel(const __grid_constant__ CovParams p0, int N, double s,
                                                              int want_grad, int ascent, double step, double tol,
                                                              int max_iter, const double* __restrict__ in,
                                                              double* __restrict__ out) {
    extern __shared__ double sm[];
    constexpr int NT = NTheta<KIND>::value;
    const int tid = threadIdx.x, tx = tid & 31, ty = tid >> 5;
    const int D = p0.D;
    const int nv = N + 1;
    const double* X = in;
    const double* y = X + (size_t)N * D;
    double* Lp = sm;                               // packed K + s I -> L
    double* Vs = Lp + N * (N + 1) / 2;             // N x nv: [I | y] -> [L^-1 | m]
    double* al = Vs + N * nv;                      // alpha
    double* rdiag = al + N;
    double* rinv = rdiag + N;
    double* pinv = rinv + N;
    double* red = pinv + (N < 12 ? 12 : N);        // 32 warps x (NT + 1) partial sums; the totals land in pinv[0..NT]

    CovParams p = p0;
    double lml_old = 0.0, lml = 0.0, err = 0.0, l_used = p.th[1];
    int it = 0, converged = 0;
    for (;;) {
#pragma unroll 1
        for (int ab = 0; ab < 16; ++ab) {
            const int i = ty + 32 * (ab >> 2), j = tx + 32 * (ab & 3);
            if (i < N && j <= i) {
                double v = pair_value<KIND>(p, X + (size_t)i * D, X + (size_t)j * D, i == j);
                if (i == j) v += s;
                Lp[tri(i, j)] = v;
            }
        }
#pragma unroll 1
        for (int i = ty; i < N; i += 32) {
            for (int c = tx; c < N; c += 32) Vs[i * nv + c] = (c == i) ? 1.0 : 0.0;
            if (tx == 0) Vs[i * nv + N] = y[i];
        }
        const int info = chol_packed(Lp, N, rdiag, rinv, pinv);
        if (info) {
            if (tid == 0) { out[0] = 0.0; out[1] = info; out[2] = it; }
            return;
        }
        for (int k = 0; k < N; ++k) {              // [L^-1 | m]: one barrier per step, row k scaled at the end
            __syncthreads();
            const double rk = rinv[k];
#pragma unroll 1
            for (int i = k + 1 + ty; i < N; i += ST / 32) {
                const double lik = Lp[tri(i, k)] * rk;
                for (int c = tx; c < nv; c += 32)
                    if (c <= k || c == N) Vs[i * nv + c] -= lik * Vs[k * nv + c];   // columns > k of row k are zero
            }
        }
        __syncthreads();
#pragma unroll 1
        for (int i = ty; i < N; i += ST / 32)
            for (int c = tx; c < nv; c += 32) Vs[i * nv + c] *= rinv[i];
        __syncthreads();
        for (int j = tid; j < N; j += ST) {        // alpha = L^-T m
            double a = 0.0;
            for (int r = j; r < N; ++r) a += Vs[r * nv + j] * Vs[r * nv + N];
            al[j] = a;
        }
        double part[NT + 1];
#pragma unroll
        for (int q = 0; q <= NT; ++q) part[q] = 0.0;
        if (ty == 31) {                            // last warp: the two LML sums, folded into slot NT
            for (int i = tx; i < N; i += 32) {
                const double m = Vs[i * nv + N];
                part[NT] += -.5 * m * m - log(rdiag[i]);
            }
        }
        __syncthreads();
        if (want_grad) {
#pragma unroll 1
            for (int ab = 0; ab < 16; ++ab) {
                const int i = ty + 32 * (ab >> 2), j = tx + 32 * (ab & 3);
                if (i < N && j <= i) {
                    double kinv = 0.0;
                    for (int r = i; r < N; ++r) kinv += Vs[r * nv + i] * Vs[r * nv + j];
                    double dk[NT];
                    pair_value_dk<KIND>(p, X + (size_t)i * D, X + (size_t)j * D, i == j, dk);
                    const double wgt = (al[i] * al[j] - kinv) * (i == j ? .5 : 1.0);
#pragma unroll
                    for (int q = 0; q < NT; ++q) part[q] += wgt * dk[q];
                }
            }
        }
#pragma unroll
        for (int q = 0; q <= NT; ++q) {
            double v = part[q];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
            if (tx == 0) red[ty * (NT + 1) + q] = v;
        }
        __syncthreads();
        if (tid <= NT) {
            double v = 0.0;
            for (int w = 0; w < ST / 32; ++w) v += red[w * (NT + 1) + tid];
            pinv[tid] = v;                         // pinv is free between factorisations (sized >= 12 by the host)
        }
        __syncthreads();
        lml = pinv[NT] - N / 2.0 * log(2.0 * PI_D);
        ++it;
        if (!ascent) break;
        err = fabs(lml - lml_old);                 // tune...:130
        l_used = p.th[1];
        p.th[1] += step * pinv[1];                 // tune...:63 (sigma's update is commented out in the reference)
        lml_old = lml;
        if (err <= tol) { converged = 1; break; }
        if (it >= max_iter) break;
        __syncthreads();                           // pinv / red are rewritten by the next iteration
    }
    if (tid == 0) {
        out[0] = lml; out[1] = 0.0; out[2] = it; out[3] = p.th[1]; out[4] = l_used; out[5] = err; out[6] = converged;
        out[7] = 0.0;
    }
    if (tid < NT) out[8 + tid] = pinv[tid];
}

int ensure_pinned(gpx_ctx* h, size_t bytes) {
    if (h->pinned_bytes >= bytes) return 0;
    if (h->pinned) {
        cudaStreamSynchronize(h->stream);
        cudaFreeHost(h->pinned);
        h->pinned = nullptr;
        h->pinned_bytes = 0;
    }
    if (bytes < 65536) bytes = 65536;
    cudaError_t e = cudaMallocHost(&h->pinned, bytes);
    if (e != cudaSuccess) {
        gpx_set_error("gpx: cudaMallocHost(%zu) failed: %s", bytes, cudaGetErrorString(e));
        return GPX_E_NOMEM;
    }
    h->pinned_bytes = bytes;
    return 0;
}

template <int KIND>
int launch_small(gpx_ctx* h, const CovParams& p, int N, int n, int nf, int mode, double s, double jitter, const double* in,
                 double* out, double* keep, size_t smem) {
    static bool configured_dev[GPX_MAX_DEVICES] = {};
    bool& configured = configured_dev[h->device % GPX_MAX_DEVICES];   // cudaFuncSetAttribute is per device
    if (!configured) {
        GPX_CUDA(cudaFuncSetAttribute(gp_small_kernel<KIND>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        configured = true;
    }
    gp_small_kernel<KIND><<<(n + SMAX - 1) / SMAX, ST, smem, h->stream>>>(p, N, n, nf, mode, s, jitter, in, out, keep);
    GPX_CHECK_LAUNCH(h);
    return 0;
}

// mode as in gp_small_kernel.  Host pointers in and out; one H2D copy, one launch, one D2H copy, one synchronisation.
int small_run(gpx_ctx* h, int kind, const double* X, int64_t N, int D, const double* y, const double* Xs, int64_t n,
              const double* theta, int ntheta, double s, double jitter, const double* Z, int nf, int mode, double* mu,
              double* var, double* fpost, double* lml) {
    GPX_ENTER(h);
    GPX_REQUIRE(kind >= 0 && kind <= 3, 2);
    GPX_REQUIRE(N >= 1 && N <= SMAX, 4);
    GPX_REQUIRE(D >= 1, 5);
    GPX_REQUIRE(n >= 1 && (n <= SMAX || mode == 0) && n <= (1 << 20), 8);
    GPX_REQUIRE(X && y && Xs && theta && mu && var && lml, 3);
    CovParams p;
    {
        static const int expect[4] = {2, 1, 2, 11};
        GPX_REQUIRE(expect[kind] == ntheta, 10);
        p.kind = kind;
        p.ntheta = ntheta;
        p.D = D;
        for (int i = 0; i < 11; ++i) p.th[i] = i < ntheta ? theta[i] : 0.0;
    }
    if (mode != 1) nf = 0;
    const size_t nX = (size_t)N * D, nXs = (size_t)n * D, nZ = (size_t)n * nf;
    const size_t in_elems = nX + N + nXs + nZ;
    const size_t out_elems = 2 * (size_t)n + nZ + 4;
    const int nl = (int)(n < SMAX ? n : SMAX);        // test points per block
    const int big = (int)(N > nl ? N : nl);
    const size_t smem = ((size_t)big * (big + 1) / 2 + (size_t)N * (nl + 1) + nl + 3 * big) * sizeof(double);
    GPX_REQUIRE(smem <= 227 * 1024, 4);
    GPX_TRY(ensure_pinned(h, (in_elems + out_elems) * sizeof(double)));
    void* dev = nullptr;
    GPX_TRY(gpx_scratch(h, (in_elems + out_elems) * sizeof(double), &dev));
    if (mode == 2 && !h->d_small) GPX_CUDA(cudaMalloc(&h->d_small, ((size_t)SMAX * (SMAX + 1) / 2 + SMAX) * sizeof(double)));
    h->small_n = 0;
    double* hin = (double*)h->pinned;
    double* hout = hin + in_elems;
    double* din = (double*)dev;
    double* dout = din + in_elems;
    memcpy(hin, X, nX * sizeof(double));
    memcpy(hin + nX, y, (size_t)N * sizeof(double));
    memcpy(hin + nX + N, Xs, nXs * sizeof(double));
    if (nZ) memcpy(hin + nX + N + nXs, Z, nZ * sizeof(double));
    GPX_CUDA(cudaMemcpyAsync(din, hin, in_elems * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    int r;
    switch (kind) {
        case GPX_COV_SE: r = launch_small<GPX_COV_SE>(h, p, (int)N, (int)n, nf, mode, s, jitter, din, dout, h->d_small, smem); break;
        case GPX_COV_LIN: r = launch_small<GPX_COV_LIN>(h, p, (int)N, (int)n, nf, mode, s, jitter, din, dout, h->d_small, smem); break;
        case GPX_COV_PER: r = launch_small<GPX_COV_PER>(h, p, (int)N, (int)n, nf, mode, s, jitter, din, dout, h->d_small, smem); break;
        default: r = launch_small<GPX_COV_CO2>(h, p, (int)N, (int)n, nf, mode, s, jitter, din, dout, h->d_small, smem); break;
    }
    if (r != 0) return r;
    GPX_CUDA(cudaMemcpyAsync(hout, dout, out_elems * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    GPX_CUDA(cudaStreamSynchronize(h->stream));
    const double* sc = hout + 2 * n + nZ;
    if (sc[1] != 0.0) return (int)sc[1];            // K + s I not positive definite (GP_regression.py:138)
    memcpy(mu, hout, (size_t)n * sizeof(double));
    memcpy(var, hout + n, (size_t)n * sizeof(double));
    *lml = sc[0];
    if (sc[2] != 0.0) return (int)sc[2];            // posterior covariance not positive definite (:154)
    if (nZ) memcpy(fpost, hout + 2 * n, nZ * sizeof(double));
    if (mode == 2) h->small_n = (int)n;
    return 0;
}

}  // namespace

extern "C" int gpx_small_max(void) { return SMAX; }

extern "C" int gpx_gp_small_posterior_host(gpx_handle h, int kind, const double* X, int64_t N, int D, const double* y,
                                           const double* Xs, int64_t n, const double* theta, int ntheta, double s,
                                           double jitter, const double* Z, int nf, double* mu, double* var, double* fpost,
                                           double* lml) {
    GPX_REQUIRE(nf >= 0 && (nf == 0 || (Z != nullptr && fpost != nullptr)), 14);
    return small_run(h, kind, X, N, D, y, Xs, n, theta, ntheta, s, jitter, Z, nf, nf > 0 ? 1 : 0, mu, var, fpost, lml);
}

extern "C" int gpx_gp_small_fit_host(gpx_handle h, int kind, const double* X, int64_t N, int D, const double* y,
                                     const double* Xs, int64_t n, const double* theta, int ntheta, double s, double jitter,
                                     double* mu, double* var, double* lml) {
    return small_run(h, kind, X, N, D, y, Xs, n, theta, ntheta, s, jitter, nullptr, 0, 2, mu, var, nullptr, lml);
}

extern "C" int gpx_gp_small_sample_host(gpx_handle h, int64_t n, int nf, const double* Z, double* fpost) {
    GPX_ENTER(h);
    GPX_REQUIRE(n >= 1 && n == h->small_n, 2);      // the factor of the last successful gpx_gp_small_fit_host
    GPX_REQUIRE(nf >= 1 && Z && fpost, 3);
    const size_t nZ = (size_t)n * nf;
    GPX_TRY(ensure_pinned(h, 2 * nZ * sizeof(double)));
    void* dev = nullptr;
    GPX_TRY(gpx_scratch(h, 2 * nZ * sizeof(double), &dev));
    double* hz = (double*)h->pinned;
    double* dz = (double*)dev;
    memcpy(hz, Z, nZ * sizeof(double));
    GPX_CUDA(cudaMemcpyAsync(dz, hz, nZ * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    const int blocks = (int)((nZ + ST - 1) / ST);
    gp_small_sample_kernel<<<blocks < 148 ? blocks : 148, ST, 0, h->stream>>>((int)n, nf, h->d_small, dz, dz + nZ);
    GPX_CHECK_LAUNCH(h);
    GPX_CUDA(cudaMemcpyAsync(hz + nZ, dz + nZ, nZ * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    GPX_CUDA(cudaStreamSynchronize(h->stream));
    memcpy(fpost, hz + nZ, nZ * sizeof(double));
    return 0;
}

namespace {

template <int KIND>
int launch_small_grad(gpx_ctx* h, const CovParams& p, int N, double s, int want_grad, int ascent, double step, double tol,
                      int max_iter, const double* in, double* out, size_t smem) {
    static bool configured_dev[GPX_MAX_DEVICES] = {};
    bool& configured = configured_dev[h->device % GPX_MAX_DEVICES];   // cudaFuncSetAttribute is per device
    if (!configured) {
        GPX_CUDA(cudaFuncSetAttribute(gp_small_grad_kernel<KIND>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        configured = true;
    }
    gp_small_grad_kernel<KIND><<<1, ST, smem, h->stream>>>(p, N, s, want_grad, ascent, step, tol, max_iter, in, out);
    GPX_CHECK_LAUNCH(h);
    return 0;
}

int small_grad_run(gpx_ctx* h, int kind, const double* X, int64_t N, int D, const double* y, const double* theta, int ntheta,
                   double s, int want_grad, int ascent, double step, double tol, int max_iter, double* out19) {
    GPX_ENTER(h);
    GPX_REQUIRE(kind >= 0 && kind <= 3, 2);
    GPX_REQUIRE(N >= 1 && N <= SMAX, 4);
    GPX_REQUIRE(D >= 1, 5);
    GPX_REQUIRE(X && y && theta && out19, 3);
    CovParams p;
    {
        static const int expect[4] = {2, 1, 2, 11};
        GPX_REQUIRE(expect[kind] == ntheta, 8);
        p.kind = kind;
        p.ntheta = ntheta;
        p.D = D;
        for (int i = 0; i < 11; ++i) p.th[i] = i < ntheta ? theta[i] : 0.0;
    }
    const size_t nX = (size_t)N * D, in_elems = nX + N, out_elems = 19;
    const size_t pv = N < 12 ? 12 : N;               // pinv doubles as the 12-slot reduction result
    const size_t smem = ((size_t)N * (N + 1) / 2 + (size_t)N * (N + 1) + 3 * N + pv + 32 * 12) * sizeof(double);
    GPX_REQUIRE(smem <= 227 * 1024, 4);
    GPX_TRY(ensure_pinned(h, (in_elems + out_elems) * sizeof(double)));
    void* dev = nullptr;
    GPX_TRY(gpx_scratch(h, (in_elems + out_elems) * sizeof(double), &dev));
    double* hin = (double*)h->pinned;
    double* hout = hin + in_elems;
    double* din = (double*)dev;
    double* dout = din + in_elems;
    memcpy(hin, X, nX * sizeof(double));
    memcpy(hin + nX, y, (size_t)N * sizeof(double));
    GPX_CUDA(cudaMemcpyAsync(din, hin, in_elems * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    int r;
    switch (kind) {
        case GPX_COV_SE: r = launch_small_grad<GPX_COV_SE>(h, p, (int)N, s, want_grad, ascent, step, tol, max_iter, din, dout, smem); break;
        case GPX_COV_LIN: r = launch_small_grad<GPX_COV_LIN>(h, p, (int)N, s, want_grad, ascent, step, tol, max_iter, din, dout, smem); break;
        case GPX_COV_PER: r = launch_small_grad<GPX_COV_PER>(h, p, (int)N, s, want_grad, ascent, step, tol, max_iter, din, dout, smem); break;
        default: r = launch_small_grad<GPX_COV_CO2>(h, p, (int)N, s, want_grad, ascent, step, tol, max_iter, din, dout, smem); break;
    }
    if (r != 0) return r;
    GPX_CUDA(cudaMemcpyAsync(hout, dout, out_elems * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    GPX_CUDA(cudaStreamSynchronize(h->stream));
    memcpy(out19, hout, out_elems * sizeof(double));
    if (hout[1] != 0.0) return (int)hout[1];         // K + s I not positive definite (tune...:127)
    return 0;
}

}  // namespace

extern "C" int gpx_gp_small_lml_grad_host(gpx_handle h, int kind, const double* X, int64_t N, int D, const double* y,
                                          const double* theta, int ntheta, double s, double* lml, double* grad) {
    GPX_REQUIRE(lml != nullptr, 10);
    double out[19];
    GPX_TRY(small_grad_run(h, kind, X, N, D, y, theta, ntheta, s, grad != nullptr, 0, 0.0, 0.0, 1, out));
    *lml = out[0];
    if (grad)
        for (int q = 0; q < ntheta; ++q) grad[q] = out[8 + q];
    return 0;
}

extern "C" int gpx_gp_small_ascent_host(gpx_handle h, const double* X, int64_t N, int D, const double* y, double sigma,
                                        double l0, double s, double step, double tol, int max_iter, double* out6) {
    GPX_REQUIRE(out6 != nullptr && max_iter >= 1, 11);
    const double theta[2] = {sigma, l0};
    double out[19];
    GPX_TRY(small_grad_run(h, GPX_COV_SE, X, N, D, y, theta, 2, s, 1, 1, step, tol, max_iter, out));
    out6[0] = out[2];   // iterations done
    out6[1] = out[3];   // l after the last step
    out6[2] = out[4];   // l the last iteration was evaluated at
    out6[3] = out[0];   // LML of the last iteration
    out6[4] = out[5];   // |LML - LML_old| of the last iteration
    out6[5] = out[6];   // 1 when the tolerance was met
    return 0;
}

namespace {
template <int KIND>
int launch_small_prior(gpx_ctx* h, const CovParams& p, int n, double s, const double* Xs, double* keep, double* out, size_t smem) {
    static bool configured_dev[GPX_MAX_DEVICES] = {};
    bool& configured = configured_dev[h->device % GPX_MAX_DEVICES];   // cudaFuncSetAttribute is per device
    if (!configured) {
        GPX_CUDA(cudaFuncSetAttribute(gp_small_prior_kernel<KIND>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        configured = true;
    }
    gp_small_prior_kernel<KIND><<<1, ST, smem, h->stream>>>(p, n, s, Xs, keep, out);
    GPX_CHECK_LAUNCH(h);
    return 0;
}
}  // namespace

extern "C" int gpx_gp_small_prior_factor_host(gpx_handle h, int kind, const double* Xs, int64_t n, int D, const double* theta,
                                              int ntheta, double s) {
    GPX_ENTER(h);
    GPX_REQUIRE(kind >= 0 && kind <= 3, 2);
    GPX_REQUIRE(Xs != nullptr && theta != nullptr, 3);
    GPX_REQUIRE(n >= 1 && n <= SMAX, 4);
    GPX_REQUIRE(D >= 1, 5);
    CovParams p;
    {
        static const int expect[4] = {2, 1, 2, 11};
        GPX_REQUIRE(expect[kind] == ntheta, 7);
        p.kind = kind;
        p.ntheta = ntheta;
        p.D = D;
        for (int i = 0; i < 11; ++i) p.th[i] = i < ntheta ? theta[i] : 0.0;
    }
    const size_t nXs = (size_t)n * D;
    const size_t smem = ((size_t)n * (n + 1) / 2 + 3 * (size_t)n) * sizeof(double);
    GPX_TRY(ensure_pinned(h, (nXs + 1) * sizeof(double)));
    void* dev = nullptr;
    GPX_TRY(gpx_scratch(h, (nXs + 1) * sizeof(double), &dev));
    if (!h->d_small) GPX_CUDA(cudaMalloc(&h->d_small, ((size_t)SMAX * (SMAX + 1) / 2 + SMAX) * sizeof(double)));
    h->small_n = 0;
    double* hin = (double*)h->pinned;
    double* din = (double*)dev;
    memcpy(hin, Xs, nXs * sizeof(double));
    GPX_CUDA(cudaMemcpyAsync(din, hin, nXs * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    int r;
    switch (kind) {
        case GPX_COV_SE: r = launch_small_prior<GPX_COV_SE>(h, p, (int)n, s, din, h->d_small, din + nXs, smem); break;
        case GPX_COV_LIN: r = launch_small_prior<GPX_COV_LIN>(h, p, (int)n, s, din, h->d_small, din + nXs, smem); break;
        case GPX_COV_PER: r = launch_small_prior<GPX_COV_PER>(h, p, (int)n, s, din, h->d_small, din + nXs, smem); break;
        default: r = launch_small_prior<GPX_COV_CO2>(h, p, (int)n, s, din, h->d_small, din + nXs, smem); break;
    }
    if (r != 0) return r;
    GPX_CUDA(cudaMemcpyAsync(hin + nXs, din + nXs, sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    GPX_CUDA(cudaStreamSynchronize(h->stream));
    if (hin[nXs] != 0.0) return (int)hin[nXs];      // K + s I not positive definite (GP_regression.py:90)
    h->small_n = (int)n;
    return 0;
}
