// Fused covariance-matrix builder and fused LML-gradient kernel.
//
// cov_build : one pass computes pairwise squared distances on 128x128 output tiles (X tiles staged in
//             shared memory, X2 transposed so lane-contiguous reads are conflict free), applies the
//             covariance family (SE / linear / periodic / CO2 composite = SE + SE*periodic + RQ + SE +
//             delta), adds the noise diagonal, writes identity padding, and optionally emits every
//             dK/dtheta_j in the same pass.  HBM-write bound: 8*n1p*n2p bytes.
// lml_grad  : reads K^-1 once (lower tiles), recomputes dK/dtheta_j per element from X and accumulates
//             .5 * sum (alpha_i alpha_j - Kinv_ij) dK_ij for all theta at once.  HBM-read bound.
#include "common.cuh"
#include "cov_eval.cuh"

namespace {

using namespace gpx_cov;

constexpr int TM = 128, TN = 128, DCH = 16, CTHREADS = 512, RI = 8;  // thread owns RI x 4 outputs

// One chunk of DC coordinates of a 128x128 tile: stage the X tiles in shared memory and accumulate.  DC > 0 is a compile-time
// chunk width (index arithmetic by shifts, the coordinate loop fully unrolled: the runtime-width loop spent 40 % of its issue
// slots on UMOV / IMAD / ISETP / BRA and an integer division per staged element -- ncu source page, round 2); DC == 0 is the
// generic runtime-width fallback.
template <int KIND, int DC>
__device__ __forceinline__ void tile_chunk(const CovParams& p, const double* __restrict__ X1, int64_t n1,
                                           const double* __restrict__ X2, int64_t n2, int row0, int col0, int d0, int dc_rt,
                                           double* xs1, double* xs2, double (&acc)[RI][4], double (&aux)[RI][4]) {
    const int tid = threadIdx.x, tx = tid & 31, ty = tid >> 5;
    const int D = p.D;
    const int dc = DC > 0 ? DC : dc_rt;
    __syncthreads();
    for (int idx = tid; idx < TM * dc; idx += CTHREADS) {
        const int r = DC > 0 ? idx / DC : idx / dc, dd = idx - r * dc;
        const int64_t gr = row0 + r;
        xs1[r * DCH + dd] = gr < n1 ? X1[gr * D + d0 + dd] : 0.0;
    }
    for (int idx = tid; idx < TN * dc; idx += CTHREADS) {
        const int c = DC > 0 ? idx / DC : idx / dc, dd = idx - c * dc;
        const int64_t gc = col0 + c;
        xs2[dd * TN + c] = gc < n2 ? X2[gc * D + d0 + dd] : 0.0;
    }
    __syncthreads();
    auto body = [&](int dd) {
        double b[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) b[j] = xs2[dd * TN + tx + 32 * j];
#pragma unroll
        for (int i = 0; i < RI; ++i) {
            const double a = xs1[(ty + 16 * i) * DCH + dd];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                if (KIND == GPX_COV_LIN) {
                    acc[i][j] += (a - p.th[0]) * (b[j] - p.th[0]);
                    aux[i][j] += a + b[j];
                } else {
                    const double df = a - b[j];
                    acc[i][j] += df * df;
                }
            }
        }
    };
    if (DC > 0) {
#pragma unroll
        for (int dd = 0; dd < (DC > 0 ? DC : 1); ++dd) body(dd);
    } else {
        for (int dd = 0; dd < dc; ++dd) body(dd);
    }
}

// Accumulate pairwise terms for a 128x128 tile: thread (tx = tid&31, ty = tid>>5) owns rows ty+16i (i<8)
// and columns tx+32j (j<4).  xs1: [128][DCH] row tile, xs2: [DCH][128] transposed column tile.
template <int KIND>
__device__ __forceinline__ void tile_accumulate(const CovParams& p, const double* __restrict__ X1, int64_t n1,
                                                const double* __restrict__ X2, int64_t n2, int row0, int col0,
                                                double* xs1, double* xs2, double (&acc)[RI][4], double (&aux)[RI][4]) {
    const int D = p.D;
#pragma unroll
    for (int i = 0; i < RI; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = aux[i][j] = 0.0;
    for (int d0 = 0; d0 < D; d0 += DCH) {
        const int dc = (D - d0) < DCH ? (D - d0) : DCH;
        if (dc == 16) tile_chunk<KIND, 16>(p, X1, n1, X2, n2, row0, col0, d0, dc, xs1, xs2, acc, aux);
        else if (dc == 8) tile_chunk<KIND, 8>(p, X1, n1, X2, n2, row0, col0, d0, dc, xs1, xs2, acc, aux);
        else if (dc == 1) tile_chunk<KIND, 1>(p, X1, n1, X2, n2, row0, col0, d0, dc, xs1, xs2, acc, aux);
        else tile_chunk<KIND, 0>(p, X1, n1, X2, n2, row0, col0, d0, dc, xs1, xs2, acc, aux);
    }
}

template <int KIND, bool WITH_DK>
__global__ void __launch_bounds__(CTHREADS) cov_build_kernel(CovParams p_in, const double* __restrict__ X1, int64_t n1,
                                                            const double* __restrict__ X2, int64_t n2, double diag_add,
                                                            int flags, double* __restrict__ K, int64_t ldk,
                                                            double* __restrict__ dK, int64_t dk_stride, int rg0, int cg0,
                                                            const double* __restrict__ scale,
                                                            const double* __restrict__ theta_dev) {
    // theta_dev (optional): the hyper-parameters are read from DEVICE memory instead of the launch parameters, so that a
    // captured CUDA graph of an optimiser iteration sees the values the previous iteration wrote (ascent.cu).
    CovParams p = p_in;
    if (theta_dev) {
#pragma unroll
        for (int q = 0; q < NTheta<KIND>::value; ++q) p.th[q] = theta_dev[q];
    }
    // scale (optional, square blocks): K[i,j] <- scale[i] k_ij scale[j] before the diagonal term is added, so that
    // B = I + W^1/2 K W^1/2 (GP_binary_classification.py:107) is built straight from X with diag_add = 1.
    // rg0 / cg0: global row / column index of this launch's element (0,0) -- K points at that element; used by
    // the multi-GPU driver to build only the block columns a rank owns.
    __shared__ double xs1[TM * DCH];
    __shared__ double xs2[DCH * TN];
    const int lrow0 = blockIdx.y * TM, lcol0 = blockIdx.x * TN;
    const int row0 = lrow0 + rg0, col0 = lcol0 + cg0;
    const int tid = threadIdx.x, tx = tid & 31, ty = tid >> 5;
    const bool same = flags & GPX_COV_SAME_X;
    K -= (int64_t)rg0 * ldk + cg0;  // index with global (row, col) below
    if (WITH_DK) dK -= (int64_t)rg0 * ldk + cg0;
    if ((flags & GPX_COV_LOWER) && col0 > row0) {  // strictly-upper tile: zeros (or left untouched)
        if (flags & GPX_COV_SKIP_UPPER) return;
#pragma unroll
        for (int i = 0; i < RI; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                int64_t off = (int64_t)(row0 + ty + 16 * i) * ldk + col0 + tx + 32 * j;
                K[off] = 0.0;
                if (WITH_DK) {
#pragma unroll
                    for (int q = 0; q < NTheta<KIND>::value; ++q) dK[q * dk_stride + off] = 0.0;
                }
            }
        return;
    }
    double acc[RI][4], aux[RI][4];
    tile_accumulate<KIND>(p, X1, n1, X2, n2, row0, col0, xs1, xs2, acc, aux);
#pragma unroll
    for (int i = 0; i < RI; ++i) {
        const int64_t r = row0 + ty + 16 * i;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int64_t c = col0 + tx + 32 * j;
            const int64_t off = r * ldk + c;
            double dk[11];
            double v;
            if (r < n1 && c < n2) {
                const bool diag = (flags & (GPX_COV_SAME_X | GPX_COV_DELTA)) && (r == c);
                double a = acc[i][j];
                v = cov_eval<KIND, WITH_DK>(p, a, aux[i][j], diag, dk);
                if (scale) v = scale[r] * v * scale[c];
                if (diag) v += diag_add;
            } else {
                v = (same && r == c) ? 1.0 : 0.0;
                if (WITH_DK) {
#pragma unroll
                    for (int q = 0; q < NTheta<KIND>::value; ++q) dk[q] = 0.0;
                }
            }
            K[off] = v;
            if (WITH_DK) {
#pragma unroll
                for (int q = 0; q < NTheta<KIND>::value; ++q) dK[q * dk_stride + off] = dk[q];
            }
        }
    }
}

// grad partials: partial[block][q] = sum over this lower tile of w_ij * dK_ij,q.
// The 128 x 128 tile of K^-1 is fetched into shared memory with cp.async at kernel entry and only waited for after the
// distance loop: the first version loaded it element by element inside the epilogue and spent most of its time in
// long-scoreboard stalls (ncu: 5.5 stalled warps per issue, FP64 pipe 29 % active, 8 % of DRAM bandwidth).
constexpr int GRAD_SMEM = TM * TN * (int)sizeof(double);
template <int KIND>
__global__ void __launch_bounds__(CTHREADS) lml_grad_kernel(CovParams p_in, const double* __restrict__ X, int64_t n,
                                                           const double* __restrict__ Kinv, int64_t ldk,
                                                           const double* __restrict__ alpha, double* __restrict__ partial,
                                                           int rg0, int cg0, const double* __restrict__ theta_dev) {
    CovParams p = p_in;
    if (theta_dev) {
#pragma unroll
        for (int q = 0; q < NTheta<KIND>::value; ++q) p.th[q] = theta_dev[q];
    }
    extern __shared__ __align__(16) double ktile[];      // [128][128] tile of K^-1
    __shared__ double xs1[TM * DCH];
    __shared__ double xs2[DCH * TN];
    __shared__ double red[CTHREADS / 32][11];
    const int bm = blockIdx.y + rg0 / TM, bn = blockIdx.x + cg0 / TN;  // global tile indices
    Kinv -= (int64_t)rg0 * ldk + cg0;
    const int tid = threadIdx.x, tx = tid & 31, ty = tid >> 5;
    const int block_id = blockIdx.y * gridDim.x + blockIdx.x;
    constexpr int NTH = NTheta<KIND>::value;
    double sums[NTH];
#pragma unroll
    for (int q = 0; q < NTH; ++q) sums[q] = 0.0;
    if (bn <= bm) {
        const int row0 = bm * TM, col0 = bn * TN;
        {   // 8192 16-byte chunks, 16 per thread, in flight during the whole distance loop
            const double* src = Kinv + (int64_t)row0 * ldk + col0;
#pragma unroll
            for (int c = 0; c < (TM * TN / 2) / CTHREADS; ++c) {
                const int id = tid + c * CTHREADS, r = id >> 6, c2 = id & 63;
                const unsigned dst = (unsigned)__cvta_generic_to_shared(ktile + r * TN + 2 * c2);
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(dst), "l"(src + (int64_t)r * ldk + 2 * c2));
            }
            asm volatile("cp.async.commit_group;\n" ::);
        }
        double acc[RI][4], aux[RI][4];
        tile_accumulate<KIND>(p, X, n, X, n, row0, col0, xs1, xs2, acc, aux);
        asm volatile("cp.async.wait_group 0;\n" ::);
        __syncthreads();
        const double wmul = (bn == bm) ? 0.5 : 1.0;  // .5 * (1 on diagonal tiles | 2 on strictly-lower tiles)
        double aj[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int64_t c = col0 + tx + 32 * j;
            aj[j] = c < n ? alpha[c] : 0.0;
        }
#pragma unroll
        for (int i = 0; i < RI; ++i) {
            const int64_t r = row0 + ty + 16 * i;
            const double ai = r < n ? alpha[r] : 0.0;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int64_t c = col0 + tx + 32 * j;
                if (r < n && c < n) {
                    double dk[11];
                    (void)cov_eval<KIND, true>(p, acc[i][j], aux[i][j], r == c, dk);
                    const double w = wmul * (ai * aj[j] - ktile[(ty + 16 * i) * TN + tx + 32 * j]);
#pragma unroll
                    for (int q = 0; q < NTH; ++q) sums[q] += w * dk[q];
                }
            }
        }
    }
#pragma unroll
    for (int q = 0; q < NTH; ++q) {
        double v = sums[q];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (tx == 0) red[ty][q] = v;
    }
    __syncthreads();
    if (tid < p.ntheta) {
        double v = 0.0;
        for (int w = 0; w < CTHREADS / 32; ++w) v += red[w][tid];
        partial[(int64_t)block_id * p.ntheta + tid] = v;
    }
}

__global__ void grad_finish_kernel(int nblocks, int ntheta, const double* __restrict__ partial, double* __restrict__ grad) {
    __shared__ double red[256];
    const int q = blockIdx.x;
    double s = 0.0;
    for (int b = threadIdx.x; b < nblocks; b += 256) s += partial[(int64_t)b * ntheta + q];
    red[threadIdx.x] = s;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) grad[q] = red[0];
}

int expected_ntheta(int kind) {
    switch (kind) {
        case GPX_COV_SE: return 2;
        case GPX_COV_LIN: return 1;
        case GPX_COV_PER: return 2;
        case GPX_COV_CO2: return 11;
    }
    return -1;
}

int make_params(int kind, int D, const double* theta, int ntheta, CovParams* p) {
    GPX_REQUIRE(expected_ntheta(kind) == ntheta, 7);
    GPX_REQUIRE(D >= 1, 5);
    p->kind = kind;
    p->ntheta = ntheta;
    p->D = D;
    for (int i = 0; i < 11; ++i) p->th[i] = i < ntheta ? theta[i] : 0.0;
    return 0;
}

template <bool WITH_DK>
int launch_build(gpx_ctx* h, const CovParams& p, const double* X1, int64_t n1, const double* X2, int64_t n2, double diag_add,
                 int flags, double* K, int64_t n1p, int64_t n2p, int64_t ldk, double* dK, int64_t dk_stride, int rg0 = 0,
                 int cg0 = 0, const double* scale = nullptr, const double* theta_dev = nullptr) {
    dim3 grid((unsigned)(n2p / TN), (unsigned)(n1p / TM));
    switch (p.kind) {
        case GPX_COV_SE:
            cov_build_kernel<GPX_COV_SE, WITH_DK><<<grid, CTHREADS, 0, h->stream>>>(p, X1, n1, X2, n2, diag_add, flags, K, ldk, dK, dk_stride, rg0, cg0, scale, theta_dev);
            break;
        case GPX_COV_LIN:
            cov_build_kernel<GPX_COV_LIN, WITH_DK><<<grid, CTHREADS, 0, h->stream>>>(p, X1, n1, X2, n2, diag_add, flags, K, ldk, dK, dk_stride, rg0, cg0, scale, theta_dev);
            break;
        case GPX_COV_PER:
            cov_build_kernel<GPX_COV_PER, WITH_DK><<<grid, CTHREADS, 0, h->stream>>>(p, X1, n1, X2, n2, diag_add, flags, K, ldk, dK, dk_stride, rg0, cg0, scale, theta_dev);
            break;
        default:
            cov_build_kernel<GPX_COV_CO2, WITH_DK><<<grid, CTHREADS, 0, h->stream>>>(p, X1, n1, X2, n2, diag_add, flags, K, ldk, dK, dk_stride, rg0, cg0, scale, theta_dev);
            break;
    }
    GPX_CHECK_LAUNCH(h);
    return 0;
}

}  // namespace

namespace {
// The four additive terms of the CO2 composite as element-wise maps of a precomputed squared-distance matrix
// (CO2_example.py:9-66: kernel_1 .. kernel_4 take `sqdist` / `l2_norm`, not the inputs).  Same operation order as cov_eval.
__global__ void __launch_bounds__(256) co2_term_kernel(int term, int64_t rows, int64_t cols, const double* __restrict__ D2,
                                                      int64_t ldd, const double* __restrict__ R, int64_t ldr, double t0, double t1,
                                                      double t2, int delta, double* __restrict__ out, int64_t ldo) {
    const int64_t r = blockIdx.x;
    for (int64_t c = (int64_t)blockIdx.y * 256 + threadIdx.x; c < cols; c += (int64_t)gridDim.y * 256) {
        const double d = D2[r * ldd + c];
        double v;
        if (term == 1) {
            v = (t0 * t0) * exp(-.5 * d / (t1 * t1));                                   // :17
        } else if (term == 2) {
            const double rr = R ? R[r * ldr + c] : sqrt(d);
            const double q = sin(gpx_cov::PI_D * rr) / t2;
            v = (t0 * t0) * exp(-.5 * d / (t1 * t1) + -2.0 * (q * q));                  // :30-32
        } else if (term == 3) {
            const double u = 1.0 + .5 * d / (t2 * (t1 * t1));                           // :44
            v = (t0 * t0) * (1.0 / pow(u, t2));                                         // :45-46
        } else {
            v = (t0 * t0) * exp(-.5 * d / (t1 * t1));                                   // :65
            if (delta && r == c) v += t2 * t2;                                          // :60-66 (delta iff square)
        }
        out[r * ldo + c] = v;
    }
}
}  // namespace

extern "C" int gpx_co2_term(gpx_handle h, int term, int64_t rows, int64_t cols, const double* sqdist, int64_t ldd,
                            const double* l2_norm, int64_t ldr, double t0, double t1, double t2, double* out, int64_t ldo) {
    GPX_ENTER(h);
    GPX_REQUIRE(term >= 1 && term <= 4, 2);
    GPX_REQUIRE(rows > 0 && cols > 0, 3);
    GPX_REQUIRE(sqdist != nullptr && out != nullptr, 5);
    dim3 grid((unsigned)rows, (unsigned)((cols + 255) / 256 > 64 ? 64 : (cols + 255) / 256));
    co2_term_kernel<<<grid, 256, 0, h->stream>>>(term, rows, cols, sqdist, ldd, l2_norm, ldr, t0, t1, t2, rows == cols ? 1 : 0, out, ldo);
    GPX_CHECK_LAUNCH(h);
    return 0;
}

extern "C" int gpx_cov_build(gpx_handle h, int kind, const double* X1, int64_t n1, const double* X2, int64_t n2, int D,
                             const double* theta_host, int ntheta, double diag_add, int flags, double* K, int64_t n1p,
                             int64_t n2p, int64_t ldk, double* dK, int64_t dk_stride) {
    GPX_ENTER(h);
    GPX_REQUIRE(kind >= 0 && kind <= 3, 2);
    GPX_REQUIRE(n1 > 0 && n2 > 0, 4);
    GPX_REQUIRE(n1p >= n1 && n1p % TM == 0 && n2p >= n2 && n2p % TN == 0, 13);
    GPX_REQUIRE(ldk >= n2p, 15);
    GPX_REQUIRE(!(flags & (GPX_COV_SAME_X | GPX_COV_DELTA)) || (n1 == n2), 11);
    CovParams p;
    GPX_TRY(make_params(kind, D, theta_host, ntheta, &p));
    if (dK) return launch_build<true>(h, p, X1, n1, X2, n2, diag_add, flags, K, n1p, n2p, ldk, dK, dk_stride);
    return launch_build<false>(h, p, X1, n1, X2, n2, diag_add, flags, K, n1p, n2p, ldk, nullptr, 0);
}

// Gradient partial sums over the block of K^-1 whose element (0,0) has global index (rg0, cg0) and which spans
// `rows` x `cols` elements (multiples of 128); grad (device, ntheta doubles) receives this block's contribution.
int gpx_lml_grad_block(gpx_ctx* h, int kind, const double* X, int64_t n, int D, const double* theta_host, int ntheta,
                       const double* Kinv, int64_t ldk, const double* alpha, double* grad, int64_t rows, int64_t cols, int rg0,
                       int cg0, const double* theta_dev) {
    GPX_REQUIRE(kind >= 0 && kind <= 3, 2);
    CovParams p;
    GPX_TRY(make_params(kind, D, theta_host, ntheta, &p));
    const int ntr = (int)(rows / TM), ntc = (int)(cols / TN);
    const size_t need = (size_t)ntr * ntc * ntheta;
    if (h->partial_elems < need) {
        if (h->d_partial) cudaFree(h->d_partial);
        h->d_partial = nullptr;
        h->partial_elems = 0;
        if (cudaMalloc(&h->d_partial, need * sizeof(double)) != cudaSuccess) {
            gpx_set_error("cudaMalloc of %zu gradient partials failed", need);
            return GPX_E_NOMEM;
        }
        h->partial_elems = need;
    }
    {
        static bool configured_dev[GPX_MAX_DEVICES] = {};
        bool& configured = configured_dev[h->device % GPX_MAX_DEVICES];   // cudaFuncSetAttribute is per device
        if (!configured) {
            GPX_CUDA(cudaFuncSetAttribute(lml_grad_kernel<GPX_COV_SE>, cudaFuncAttributeMaxDynamicSharedMemorySize, GRAD_SMEM));
            GPX_CUDA(cudaFuncSetAttribute(lml_grad_kernel<GPX_COV_LIN>, cudaFuncAttributeMaxDynamicSharedMemorySize, GRAD_SMEM));
            GPX_CUDA(cudaFuncSetAttribute(lml_grad_kernel<GPX_COV_PER>, cudaFuncAttributeMaxDynamicSharedMemorySize, GRAD_SMEM));
            GPX_CUDA(cudaFuncSetAttribute(lml_grad_kernel<GPX_COV_CO2>, cudaFuncAttributeMaxDynamicSharedMemorySize, GRAD_SMEM));
            configured = true;
        }
    }
    dim3 grid(ntc, ntr);
    switch (kind) {
        case GPX_COV_SE: lml_grad_kernel<GPX_COV_SE><<<grid, CTHREADS, GRAD_SMEM, h->stream>>>(p, X, n, Kinv, ldk, alpha, h->d_partial, rg0, cg0, theta_dev); break;
        case GPX_COV_LIN: lml_grad_kernel<GPX_COV_LIN><<<grid, CTHREADS, GRAD_SMEM, h->stream>>>(p, X, n, Kinv, ldk, alpha, h->d_partial, rg0, cg0, theta_dev); break;
        case GPX_COV_PER: lml_grad_kernel<GPX_COV_PER><<<grid, CTHREADS, GRAD_SMEM, h->stream>>>(p, X, n, Kinv, ldk, alpha, h->d_partial, rg0, cg0, theta_dev); break;
        default: lml_grad_kernel<GPX_COV_CO2><<<grid, CTHREADS, GRAD_SMEM, h->stream>>>(p, X, n, Kinv, ldk, alpha, h->d_partial, rg0, cg0, theta_dev); break;
    }
    GPX_CHECK_LAUNCH(h);
    grad_finish_kernel<<<ntheta, 256, 0, h->stream>>>(ntr * ntc, ntheta, h->d_partial, grad);
    GPX_CHECK_LAUNCH(h);
    return 0;
}

// Covariance block whose element (0,0) has global index (rg0, cg0): rows x cols elements written at K (ld ldk).
int gpx_cov_build_block(gpx_ctx* h, int kind, const double* X, int64_t n, int D, const double* theta_host, int ntheta,
                        double diag_add, int flags, double* K, int64_t rows, int64_t cols, int64_t ldk, int rg0, int cg0,
                        const double* scale, const double* theta_dev) {
    CovParams p;
    GPX_TRY(make_params(kind, D, theta_host, ntheta, &p));
    return launch_build<false>(h, p, X, n, X, n, diag_add, flags, K, rows, cols, ldk, nullptr, 0, rg0, cg0, scale, theta_dev);
}

extern "C" int gpx_lml_grad(gpx_handle h, int kind, const double* X, int64_t n, int D, const double* theta_host, int ntheta,
                            const double* Kinv, int64_t ldk, const double* alpha, double* grad) {
    GPX_ENTER(h);
    const int64_t np_ = ((n + TM - 1) / TM) * TM;
    return gpx_lml_grad_block(h, kind, X, n, D, theta_host, ntheta, Kinv, ldk, alpha, grad, np_, np_, 0, 0, nullptr);
}
