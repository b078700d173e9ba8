// Element-wise / structured kernels of the Laplace approximations.
// The O(N^3) work of the Newton loops (factor B = I + W^1/2 K W^1/2, inverses, solves) is done by
// potrf.cu / gemm.cu; this file holds the HBM-bound pieces around it.
#include "common.cuh"

namespace {

__device__ __forceinline__ double sigmoid(double x) {
    // numerically stable logistic (scipy.special.expit)
    if (x >= 0.0) {
        return 1.0 / (1.0 + exp(-x));
    }
    const double e = exp(x);
    return e / (1.0 + e);
}

// GP_binary_classification.py:66-83.  mode 0: grad = t - sigmoid(y f) (as shipped), mode 1: t - sigmoid(f).
__global__ void logistic_terms_kernel(int mode, int64_t n, const double* __restrict__ y, const double* __restrict__ f,
                                      double* __restrict__ grad, double* __restrict__ w, double* __restrict__ sw) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double yi = y[i], fi = f[i];
    const double t = (yi + 1.0) / 2.0;
    const double p = sigmoid(fi);
    if (grad) grad[i] = t - (mode == 0 ? sigmoid(yi * fi) : p);
    const double wi = p * (1.0 - p);
    if (w) w[i] = wi;
    if (sw) sw[i] = sqrt(wi);
}

// B = I + diag(sw) K diag(sw) on the true n x n block; identity on the padding.
__global__ void __launch_bounds__(256) build_B_kernel(const double* __restrict__ K, const double* __restrict__ sw, int64_t n,
                                                     int64_t np_, int64_t ld, double* __restrict__ B) {
    const int64_t r = blockIdx.x;   // rows on grid.x (grid.y is limited to 65535, and N = 65536 is a headline size)
    const double sr = r < n ? sw[r] : 0.0;
    for (int64_t c = (int64_t)blockIdx.y * 256 + threadIdx.x; c < np_; c += (int64_t)gridDim.y * 256) {
        double v = (r == c) ? 1.0 : 0.0;
        if (r < n && c < n) v += sr * K[r * ld + c] * sw[c];
        B[r * ld + c] = v;
    }
}

// softmax over classes for every point (GP_multi_classification.py:26-33 applied per point :51-58).  One thread per
// OUTPUT index k = c*stride + i.  With stride < n (the reference's literal 60 when n > 60) several (c, i) pairs map to
// the same k; the reference's sequential loop (i outer, c inner) leaves the value of the pair with the largest i, i.e.
// the smallest c -- reproduced here deterministically instead of racing.
__global__ void softmax_classes_kernel(int C, int64_t n, int64_t stride, const double* __restrict__ f, double* __restrict__ pi) {
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= (int64_t)(C - 1) * stride + n) return;
    int64_t cw = k - (n - 1) > 0 ? (k - (n - 1) + stride - 1) / stride : 0;   // smallest class whose point index is < n
    const int64_t i = k - cw * stride;
    if (i < 0 || cw >= C) return;                                            // gap between class segments (stride > n)
    double mx = f[i];
    for (int c = 1; c < C; ++c) mx = fmax(mx, f[(int64_t)c * stride + i]);
    double s = 0.0;
    for (int c = 0; c < C; ++c) s += exp(f[(int64_t)c * stride + i] - mx);
    pi[k] = exp(f[cw * stride + i] - mx) / s;
}

// Reference-faithful multiclass Hessian pieces (GP_multi_classification.py:150-157).  The reference's
// pi_matrix is filled POINT-major (row a = i*C + c) while D = diag(pi_vector) is CLASS-major (index
// c*stride + i); both conventions are reproduced here.  p_i[c] = pi_vec[c*stride + i].
//   out[a][b] = Kinv[a][b] + (a==b)(c_diag + pi_vec[a]) - [a/C == b/C] p_{a/C}[a%C] p_{a/C}[b%C]
__global__ void __launch_bounds__(256) multi_ref_hessian_kernel(int C, int64_t n, int64_t stride, const double* __restrict__ Kinv,
                                                               int64_t ld, const double* __restrict__ pi_vec, double c_diag,
                                                               int64_t np_, double* __restrict__ out) {
    const int64_t a = blockIdx.x;
    const int64_t N = (int64_t)C * n;
    for (int64_t b = (int64_t)blockIdx.y * 256 + threadIdx.x; b < np_; b += (int64_t)gridDim.y * 256) {
        double v;
        if (a < N && b < N) {
            v = Kinv[a * ld + b];
            if (a == b) v += c_diag + pi_vec[a];
            const int64_t ia = a / C, ib = b / C;
            if (ia == ib) v -= pi_vec[(a % C) * stride + ia] * pi_vec[(b % C) * stride + ia];
        } else {
            v = (a == b) ? 1.0 : 0.0;
        }
        out[a * ld + b] = v;
    }
}

// out[a] = pi_vec[a] f[a] - p_i[a%C] * sum_c p_i[c] f[i*C + c],  i = a / C   (W f with the reference's W)
__global__ void multi_ref_wf_kernel(int C, int64_t n, int64_t stride, const double* __restrict__ pi_vec,
                                    const double* __restrict__ f, double* __restrict__ out) {
    int64_t a = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (a >= (int64_t)C * n) return;
    const int64_t i = a / C;
    double s = 0.0;
    for (int c = 0; c < C; ++c) s += pi_vec[(int64_t)c * stride + i] * f[i * C + c];
    out[a] = pi_vec[a] * f[a] - pi_vec[(a % C) * stride + i] * s;
}

// textbook Alg 3.3 line 9: b = (D - Pi Pi^T) f + y - pi with class-major Pi (stride n)
__global__ void multi_b_kernel(int C, int64_t n, const double* __restrict__ pi, const double* __restrict__ f,
                               const double* __restrict__ y, double* __restrict__ b) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double s = 0.0;
    for (int c = 0; c < C; ++c) s += pi[(int64_t)c * n + i] * f[(int64_t)c * n + i];
    for (int c = 0; c < C; ++c) {
        const int64_t k = (int64_t)c * n + i;
        b[k] = pi[k] * f[k] - pi[k] * s + y[k] - pi[k];
    }
}

// Esum(lower tiles) (+)= diag(sd) X diag(sd) : accumulate=0 overwrites.  Rows/cols >= n: identity/zero padding
// is written when accumulate == 0 so that Esum stays factorable on the padded size.
__global__ void __launch_bounds__(256) scale_sym_acc_kernel(const double* __restrict__ X, const double* __restrict__ sd, int64_t n,
                                                           int64_t np_, int64_t ld, int accumulate, double* __restrict__ E) {
    const int64_t r = blockIdx.x;
    const double sr = r < n ? sd[r] : 0.0;
    const int64_t cend = ((r / GPX_T) + 1) * GPX_T;  // through the end of the diagonal tile
    for (int64_t c = (int64_t)blockIdx.y * 256 + threadIdx.x; c < cend; c += (int64_t)gridDim.y * 256) {
        double v;
        if (r < n && c < n) v = sr * X[r * ld + c] * sd[c];
        else v = (!accumulate && r == c) ? 1.0 : 0.0;
        if (accumulate) E[r * ld + c] += v;
        else E[r * ld + c] = v;
    }
}

// y = S x for symmetric S given by its lower triangle (full diagonal tiles): two-pass, row part + column part
__global__ void __launch_bounds__(256) symv_lower_row_kernel(int64_t n, const double* __restrict__ S, int64_t ld,
                                                            const double* __restrict__ x, double* __restrict__ y) {
    // y[r] = sum_{c<=r} S[r][c] x[c]   (one warp per row)
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int64_t r = (int64_t)blockIdx.x * 8 + warp; r < n; r += (int64_t)gridDim.x * 8) {
        double s = 0.0;
        for (int64_t c = lane; c <= r; c += 32) s += S[r * ld + c] * x[c];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (lane == 0) y[r] = s;
    }
}
__global__ void __launch_bounds__(256) symv_lower_col_kernel(int64_t n, const double* __restrict__ S, int64_t ld,
                                                            const double* __restrict__ x, double* __restrict__ y) {
    // y[c] += sum_{r>c} S[r][c] x[r]  (thread per column; coalesced across c)
    int64_t c = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (c >= n) return;
    double s0 = 0.0, s1 = 0.0;
    int64_t r = c + 1;
    for (; r + 1 < n; r += 2) {
        s0 += S[r * ld + c] * x[r];
        s1 += S[(r + 1) * ld + c] * x[r + 1];
    }
    if (r < n) s0 += S[r * ld + c] * x[r];
    y[c] += s0 + s1;
}

}  // namespace

extern "C" int gpx_logistic_terms(gpx_handle h, int mode, int64_t n, const double* y, const double* f, double* grad,
                                  double* w, double* sw) {
    GPX_ENTER(h);
    GPX_REQUIRE(mode == 0 || mode == 1, 2);
    logistic_terms_kernel<<<(unsigned)((n + 255) / 256), 256, 0, h->stream>>>(mode, n, y, f, grad, w, sw);
    GPX_CHECK_LAUNCH(h);
    return 0;
}

extern "C" int gpx_build_B(gpx_handle h, const double* K, const double* sw, int64_t n, int64_t np_, int64_t ld, double* B) {
    GPX_ENTER(h);
    GPX_REQUIRE(np_ % GPX_T == 0 && np_ >= n, 5);
    dim3 grid((unsigned)np_, (unsigned)((np_ + 255) / 256 > 64 ? 64 : (np_ + 255) / 256));
    build_B_kernel<<<grid, 256, 0, h->stream>>>(K, sw, n, np_, ld, B);
    GPX_CHECK_LAUNCH(h);
    return 0;
}

extern "C" int gpx_softmax_classes(gpx_handle h, int C, int64_t n, int64_t stride, const double* f, double* pi) {
    GPX_ENTER(h);
    GPX_REQUIRE(C >= 1 && n >= 1 && stride >= 1, 2);
    const int64_t outs = (int64_t)(C - 1) * stride + n;
    softmax_classes_kernel<<<(unsigned)((outs + 255) / 256), 256, 0, h->stream>>>(C, n, stride, f, pi);
    GPX_CHECK_LAUNCH(h);
    return 0;
}

extern "C" int gpx_multi_ref_hessian(gpx_handle h, int C, int64_t n, int64_t stride, const double* Kinv, int64_t ld,
                                     const double* pi_vec, double c_diag, int64_t np_, double* out) {
    GPX_ENTER(h);
    GPX_REQUIRE(np_ % GPX_T == 0 && np_ >= (int64_t)C * n, 9);
    dim3 grid((unsigned)np_, (unsigned)((np_ + 255) / 256 > 64 ? 64 : (np_ + 255) / 256));
    multi_ref_hessian_kernel<<<grid, 256, 0, h->stream>>>(C, n, stride, Kinv, ld, pi_vec, c_diag, np_, out);
    GPX_CHECK_LAUNCH(h);
    return 0;
}

extern "C" int gpx_multi_ref_wf(gpx_handle h, int C, int64_t n, int64_t stride, const double* pi_vec, const double* f,
                                double* out) {
    GPX_ENTER(h);
    multi_ref_wf_kernel<<<(unsigned)(((int64_t)C * n + 255) / 256), 256, 0, h->stream>>>(C, n, stride, pi_vec, f, out);
    GPX_CHECK_LAUNCH(h);
    return 0;
}

extern "C" int gpx_multi_b(gpx_handle h, int C, int64_t n, const double* pi, const double* f, const double* y, double* b) {
    GPX_ENTER(h);
    multi_b_kernel<<<(unsigned)((n + 255) / 256), 256, 0, h->stream>>>(C, n, pi, f, y, b);
    GPX_CHECK_LAUNCH(h);
    return 0;
}

extern "C" int gpx_scale_sym_acc(gpx_handle h, const double* X, const double* sd, int64_t n, int64_t np_, int64_t ld,
                                 int accumulate, double* E) {
    GPX_ENTER(h);
    GPX_REQUIRE(np_ % GPX_T == 0 && np_ >= n, 5);
    dim3 grid((unsigned)np_, (unsigned)((np_ + 255) / 256 > 64 ? 64 : (np_ + 255) / 256));
    scale_sym_acc_kernel<<<grid, 256, 0, h->stream>>>(X, sd, n, np_, ld, accumulate, E);
    GPX_CHECK_LAUNCH(h);
    return 0;
}

extern "C" int gpx_symv_lower(gpx_handle h, int64_t n, const double* S, int64_t ld, const double* x, double* y) {
    GPX_ENTER(h);
    int64_t blocks = (n + 7) / 8;
    if (blocks > 148 * 16) blocks = 148 * 16;
    symv_lower_row_kernel<<<(unsigned)blocks, 256, 0, h->stream>>>(n, S, ld, x, y);
    GPX_CHECK_LAUNCH(h);
    symv_lower_col_kernel<<<(unsigned)((n + 255) / 256), 256, 0, h->stream>>>(n, S, ld, x, y);
    GPX_CHECK_LAUNCH(h);
    return 0;
}

namespace {
__global__ void __launch_bounds__(256) scale_rows_kernel(int64_t rows, int64_t cols, int64_t ld, const double* __restrict__ s,
                                                        double* __restrict__ M) {
    const int64_t r = blockIdx.x;
    if (r >= rows) return;
    const double sr = s[r];
    for (int64_t c = (int64_t)blockIdx.y * 256 + threadIdx.x; c < cols; c += (int64_t)gridDim.y * 256) M[r * ld + c] *= sr;
}
// mirror the strictly-lower triangle into the upper one (tile-wise transpose through shared memory)
__global__ void __launch_bounds__(256) symmetrize_kernel(int64_t n, double* __restrict__ A, int64_t ld) {
    __shared__ double t[32][33];
    const int bi = blockIdx.y, bj = blockIdx.x;
    if (bj > bi) return;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
    for (int r = ty; r < 32; r += 8) {
        int64_t i = (int64_t)bi * 32 + r, j = (int64_t)bj * 32 + tx;
        t[r][tx] = (i < n && j < n) ? A[i * ld + j] : 0.0;
    }
    __syncthreads();
    for (int r = ty; r < 32; r += 8) {
        int64_t i = (int64_t)bj * 32 + r, j = (int64_t)bi * 32 + tx;  // destination (upper) element A[i][j] = lower[j][i]
        if (i < n && j < n && j > i) A[i * ld + j] = t[tx][r];
    }
}
}  // namespace

extern "C" int gpx_scale_rows(gpx_handle h, int64_t rows, int64_t cols, int64_t ld, const double* s, double* M) {
    GPX_ENTER(h);
    if (rows <= 0 || cols <= 0) return 0;
    dim3 grid((unsigned)rows, (unsigned)((cols + 255) / 256 > 64 ? 64 : (cols + 255) / 256));
    scale_rows_kernel<<<grid, 256, 0, h->stream>>>(rows, cols, ld, s, M);
    GPX_CHECK_LAUNCH(h);
    return 0;
}

extern "C" int gpx_symmetrize(gpx_handle h, int64_t n, double* A, int64_t ld) {
    GPX_ENTER(h);
    const unsigned nb = (unsigned)((n + 31) / 32);
    symmetrize_kernel<<<dim3(nb, nb), 256, 0, h->stream>>>(n, A, ld);
    GPX_CHECK_LAUNCH(h);
    return 0;
}
