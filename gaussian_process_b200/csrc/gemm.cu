// FP64 tensor-core GEMM for sm_100a: C = alpha * op(A) op(B) + beta * C on 128 x TN output tiles.
//
// Tensor path: mma.sync.aligned.m8n8k4.f64 (SASS DMMA.8x8x4 -- the only FP64 tensor instruction on
// sm_100a; tcgen05.mma has no f64 kind, and DMMA shares the FP64 FMA units so 128 flop/clk/SM is the
// ceiling).  Warp tile 64 x 32 = 8 x 4 DMMA fragments (32 independent accumulators per warp), K slab 16,
// multi-stage cp.async (LDGSTS) pipeline into padded shared memory laid out so every 64-bit fragment
// load is bank-conflict free:
//   k-major operand  : smem[row][16+4]   lane(g,t) reads [r0+g][kk+t]  -> bank8 = 4g+t   (distinct)
//   k-strided operand: smem[k][T+4]      lane(g,t) reads [kk+t][c0+g]  -> bank8 = 4t+g   (distinct)
// Two tile configurations:
//   TN = 128 : one CTA of 8 warps per SM (in-place right-TRSM leaves need the whole 128-wide row block)
//   TN =  64 : CTAs of 4 warps, TWO resident per SM -- while one CTA sits in its prologue / epilogue /
//              __syncthreads the other keeps the DMMA pipe busy (hardware "ping-pong"); default.
// This kernel is the trailing update (SYRK/GEMM), the TRSM/TRTRI/LAUUM work-horse and the prediction
// TRSM; triangular structure is exploited at tile granularity through per-tile k ranges.
#include <stdlib.h>
#include "common.cuh"

namespace {

constexpr int BM = 128, BK = 16;
constexpr int KC_LD = BK + 4;  // 20 doubles per row of a k-major slab

template <int TN>
struct Cfg {
    static constexpr int NT = TN * 2;                     // threads: 8 warps (TN=128) or 4 warps (TN=64)
    static constexpr int WN = TN / 32;                    // warps along N
    static constexpr int STAGES = (TN == 128) ? 4 : 3;
    static constexpr int A_ELEMS = 128 * KC_LD;           // >= 16 * 132
    static constexpr int B_ELEMS = (TN * KC_LD > 16 * (TN + 4)) ? TN * KC_LD : 16 * (TN + 4);
    static constexpr int SMEM = STAGES * (A_ELEMS + B_ELEMS) * (int)sizeof(double);
    static constexpr int MIN_CTAS = (TN == 128) ? 1 : 2;
};

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

// global -> shared for one operand slab of ROWS "rows".  KMAJOR: ROWS x 16 k (row r = P[(r0+r)*ld + k0 ..]);
// else 16 k-rows x ROWS (row kr = P[(k0+kr)*ld + r0 ..]).
template <bool KMAJOR, int ROWS, int NT>
__device__ __forceinline__ void load_slab(double* s, const double* __restrict__ P, int64_t ld, int r0, int k0, int tid) {
    constexpr int CHUNKS = ROWS * 8;  // 16-byte chunks in the slab
#pragma unroll
    for (int i = 0; i < CHUNKS / NT; ++i) {
        int id = tid + i * NT;
        if (KMAJOR) {
            int r = id >> 3, c = id & 7;
            cp_async16(s + r * KC_LD + c * 2, P + (int64_t)(r0 + r) * ld + k0 + c * 2);
        } else {
            constexpr int CPR = ROWS / 2;  // chunks per k-row
            int kr = id / CPR, c = id % CPR;
            cp_async16(s + kr * (ROWS + 4) + c * 2, P + (int64_t)(k0 + kr) * ld + r0 + c * 2);
        }
    }
}

__device__ __forceinline__ int mapped_pos(const GemmArgs& p, int col0) {
    if (p.cyc_P <= 0) return col0;
    const int bw = p.cyc_tpb * 128;  // block width in elements
    const int lb = col0 / bw;
    return (int)gpx_cyc_global(lb + p.cyc_q0, p.cyc_P, p.cyc_p, p.cyc_snake) * bw + col0 % bw - p.cyc_row_base;
}

template <bool A_KM, bool B_KM, int TN>
__global__ void __launch_bounds__(Cfg<TN>::NT, Cfg<TN>::MIN_CTAS) dgemm_dmma_kernel(GemmArgs p) {
    using C_ = Cfg<TN>;
    constexpr int NT = C_::NT, STAGES = C_::STAGES;
    constexpr int KS_LDA = 128 + 4, KS_LDB = TN + 4;
    extern __shared__ __align__(16) double smem[];
    const int tid = threadIdx.x;
    const int bn = blockIdx.x;
    const int bm = p.rev_rows ? (gridDim.y - 1 - blockIdx.y) : blockIdx.y;
    const int row0 = bm * BM, col0 = bn * TN;
    const int gpos = mapped_pos(p, col0);  // column position in global (block-cyclic) numbering
    const int brow0 = (p.cyc_P > 0 && !p.cyc_b_rows) ? col0 : gpos;  // offset of this column tile inside the B operand
    if (p.lower_only && gpos >= row0 + BM) return;                    // tile entirely above the diagonal
    const int z1 = p.batch2 > 1 ? (int)(blockIdx.z % p.batch) : (int)blockIdx.z;
    const int z2 = p.batch2 > 1 ? (int)(blockIdx.z / p.batch) : 0;
    const double* __restrict__ A = p.A + (int64_t)z1 * p.sA + (int64_t)z2 * p.sA2;
    const double* __restrict__ B = p.B + (int64_t)z1 * p.sB + (int64_t)z2 * p.sB2;
    double* C = p.C + (int64_t)z1 * p.sC + (int64_t)z2 * p.sC2;

    int kbeg = (p.kb_mode == 1 ? row0 : (p.kb_mode == 2 ? (col0 / 128) * 128 : (p.kb_mode == 3 ? (gpos / 128) * 128 : 0))) +
               (int)(z1 * p.kb_batch) + p.kb_const;
    if (kbeg < 0) kbeg = 0;
    int kend = p.ke_mode == 1 ? row0 + BM : (p.ke_mode == 2 ? (col0 / 128) * 128 + 128 : p.K);
    if (kend > p.K) kend = p.K;
    const int nk = kend > kbeg ? (kend - kbeg) / BK : 0;

    const int warp = tid >> 5, lane = tid & 31;
    const int g = lane >> 2, t = lane & 3;
    const int wm0 = (warp / C_::WN) * 64, wn0 = (warp % C_::WN) * 32;

    double acc[8][4][2];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

    auto sA = [&](int st) { return smem + st * (C_::A_ELEMS + C_::B_ELEMS); };
    auto sB = [&](int st) { return smem + st * (C_::A_ELEMS + C_::B_ELEMS) + C_::A_ELEMS; };

    // prologue: prefetch STAGES-1 slabs
#pragma unroll
    for (int st = 0; st < STAGES - 1; ++st) {
        if (st < nk) {
            load_slab<A_KM, 128, NT>(sA(st), A, p.lda, row0, kbeg + st * BK, tid);
            load_slab<B_KM, TN, NT>(sB(st), B, p.ldb, brow0, kbeg + st * BK, tid);
        }
        cp_async_commit();
    }

    for (int kt = 0; kt < nk; ++kt) {
        cp_async_wait<STAGES - 2>();
        __syncthreads();
        {
            int nxt = kt + STAGES - 1;
            if (nxt < nk) {
                int st = nxt % STAGES;
                load_slab<A_KM, 128, NT>(sA(st), A, p.lda, row0, kbeg + nxt * BK, tid);
                load_slab<B_KM, TN, NT>(sB(st), B, p.ldb, brow0, kbeg + nxt * BK, tid);
            }
            cp_async_commit();
        }
        const double* a_s = sA(kt % STAGES);
        const double* b_s = sB(kt % STAGES);
#pragma unroll
        for (int kk = 0; kk < BK; kk += 4) {
            double af[8], bf[4];
#pragma unroll
            for (int i = 0; i < 8; ++i)
                af[i] = A_KM ? a_s[(wm0 + i * 8 + g) * KC_LD + kk + t] : a_s[(kk + t) * KS_LDA + wm0 + i * 8 + g];
#pragma unroll
            for (int j = 0; j < 4; ++j)
                bf[j] = B_KM ? b_s[(wn0 + j * 8 + g) * KC_LD + kk + t] : b_s[(kk + t) * KS_LDB + wn0 + j * 8 + g];
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) dmma(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
        }
    }
    cp_async_wait<0>();
    // all of this CTA's operand reads have landed before any of its stores: in-place use is safe when
    // the operand region the CTA reads is exactly its own output tile (TRSM leaves).
    __syncthreads();

    const double alpha = p.alpha, beta = p.beta;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int r = row0 + wm0 + i * 8 + g;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int c = col0 + wn0 + j * 8 + 2 * t;
            double2* dst = reinterpret_cast<double2*>(C + (int64_t)r * p.ldc + c);
            double2 v;
            v.x = alpha * acc[i][j][0];
            v.y = alpha * acc[i][j][1];
            if (beta != 0.0) {
                double2 o = *dst;
                v.x += beta * o.x;
                v.y += beta * o.y;
            }
            *dst = v;
        }
    }
}

int host_mapped_pos(const GemmArgs& a, int col0) {
    if (a.cyc_P <= 0) return col0;
    const int bw = a.cyc_tpb * 128;
    return (int)gpx_cyc_global(col0 / bw + a.cyc_q0, a.cyc_P, a.cyc_p, a.cyc_snake) * bw + col0 % bw - a.cyc_row_base;
}

// flops a launch really executes (tile-granular k ranges), for the instrumentation
double exec_flops(const GemmArgs& a, int TN) {
    const int gx = a.N / TN, gy = a.M / BM, gz = a.batch > 0 ? a.batch : 1;
    double ksum = 0.0;
    for (int bm = 0; bm < gy; ++bm)
        for (int bn = 0; bn < gx; ++bn) {
            const int col0 = bn * TN;
            const int gp = host_mapped_pos(a, col0);
            if (a.lower_only && gp >= bm * BM + BM) continue;
            int kb = a.kb_mode == 1 ? bm * BM : (a.kb_mode == 2 ? (col0 / 128) * 128 : (a.kb_mode == 3 ? (gp / 128) * 128 : 0));
            int ke = a.ke_mode == 1 ? bm * BM + BM : (a.ke_mode == 2 ? (col0 / 128) * 128 + 128 : a.K);
            if (ke > a.K) ke = a.K;
            for (int z = 0; z < gz; ++z) {
                int kbz = kb + (int)(z * a.kb_batch) + a.kb_const;
                if (kbz < 0) kbz = 0;
                if (ke > kbz) ksum += (double)(ke - kbz);
            }
        }
    return 2.0 * BM * TN * ksum * (a.batch2 > 1 ? a.batch2 : 1);
}

template <bool A_KM, bool B_KM, int TN>
int launch_t(gpx_ctx* h, const GemmArgs& a) {
    using C_ = Cfg<TN>;
    static bool configured_dev[GPX_MAX_DEVICES] = {};
    bool& configured = configured_dev[h->device % GPX_MAX_DEVICES];   // cudaFuncSetAttribute is per device
    if (!configured) {
        GPX_CUDA(cudaFuncSetAttribute(dgemm_dmma_kernel<A_KM, B_KM, TN>, cudaFuncAttributeMaxDynamicSharedMemorySize, C_::SMEM));
        GPX_CUDA(cudaFuncSetAttribute(dgemm_dmma_kernel<A_KM, B_KM, TN>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
        configured = true;
    }
    dim3 grid(a.N / TN, a.M / BM, (a.batch > 0 ? a.batch : 1) * (a.batch2 > 1 ? a.batch2 : 1));
    if (h->timing_on) gpx_timing_gemm_begin(h, exec_flops(a, TN), a.M, a.N, a.K);
    dgemm_dmma_kernel<A_KM, B_KM, TN><<<grid, C_::NT, C_::SMEM, h->stream>>>(a);
    GPX_CHECK_LAUNCH(h);
    gpx_timing_gemm_end(h);
    return 0;
}

template <int TN>
int dispatch(gpx_ctx* h, const GemmArgs& a) {
    if (a.a_kmajor && a.b_kmajor) return launch_t<true, true, TN>(h, a);
    if (a.a_kmajor && !a.b_kmajor) return launch_t<true, false, TN>(h, a);
    if (!a.a_kmajor && a.b_kmajor) return launch_t<false, true, TN>(h, a);
    return launch_t<false, false, TN>(h, a);
}

int default_tn() {
    static int tn = 0;
    if (tn == 0) {
        const char* e = getenv("GPX_GEMM_TN");
        tn = (e && atoi(e) == 128) ? 128 : 64;
    }
    return tn;
}

}  // namespace

int gpx_gemm_launch(gpx_ctx* h, const GemmArgs& a) {
    if (a.M <= 0 || a.N <= 0) return 0;
    GPX_REQUIRE(a.M % BM == 0 && a.N % 128 == 0 && a.K % BK == 0, 2);
    GPX_REQUIRE(a.lda % 2 == 0 && a.ldb % 2 == 0 && a.ldc % 2 == 0, 3);
    GPX_REQUIRE(((uintptr_t)a.A % 16) == 0 && ((uintptr_t)a.B % 16) == 0 && ((uintptr_t)a.C % 16) == 0, 4);
    // a right-TRSM leaf updates a 128-wide row block in place: one CTA must own the whole block
    const bool inplace_rows = (a.C == a.A) && a.a_kmajor;
    if (inplace_rows || default_tn() == 128) return dispatch<128>(h, a);
    {   // k-major x k-major (trailing update / TRSM update): TMA-fed kernel when available
        const int r = gpx_gemm_tma_try_launch(h, a, h->timing_on ? exec_flops(a, 64) : 0.0);
        if (r != 0) return r < 0 ? r : 0;
    }
    return dispatch<64>(h, a);
}

extern "C" int gpx_gemm(gpx_handle h, int a_kmajor, int b_kmajor, int64_t M, int64_t N, int64_t K, double alpha,
                        const double* A, int64_t lda, const double* B, int64_t ldb, double beta, double* C, int64_t ldc) {
    GPX_ENTER(h);
    GemmArgs a{};
    a.A = A; a.B = B; a.C = C;
    a.M = (int)M; a.N = (int)N; a.K = (int)K;
    a.lda = lda; a.ldb = ldb; a.ldc = ldc;
    a.alpha = alpha; a.beta = beta;
    a.batch = 1;
    a.a_kmajor = a_kmajor; a.b_kmajor = b_kmajor;
    return gpx_gemm_launch(h, a);
}
