// TMA-fed FP64 DMMA GEMM (all four operand layouts; unbatched, out-of-place launches): operand slabs are fetched by the
// Tensor Memory Accelerator (cp.async.bulk.tensor.2d, SASS UTMALDG) into 128-byte-swizzled shared memory and handed to the
// DMMA warps through an mbarrier full/empty ring -- no per-thread address arithmetic, no __syncthreads in the main loop.
//
//   k-major operand   : one box {16 k, rows} -> smem [rows][16] doubles (row = 128 B = 8 chunks of 16 B, chunk ^= row & 7)
//   k-strided operand : rows/16 boxes {16 m, 16 k} -> smem [box][16 k][16 m]            (chunk ^= k & 7)
//   Both stay bank-conflict free for the 64-bit fragment loads of mma.m8n8k4 with
//     k(t, j) = 8*(j>>1) + ({0,3,4,7}[t] ^ (j&1))        (a permutation of the 16 k of a slab, shared by A and B)
//     k-major fragment rows permuted inside each group of 8:  rho(g) = [0,1,4,5,2,3,6,7][g]
//   (16 lanes of a half warp then hit 16 distinct 8-byte bank pairs in either layout); the permutations only relabel
//   which accumulator a lane owns, the epilogue undoes them (column pairs stay adjacent: rho(2t), rho(2t)+1).
//   Producer = thread 0: refills the stage consumed two slabs ago (waits on its `empty` mbarrier), consumers wait on
//   `full`, arrive on `empty`.  Two 4-warp CTAs per SM as in gemm.cu.
#include <cuda.h>
#include <stdlib.h>
#include "common.cuh"

namespace {

constexpr int BM = 128, BK = 16;
constexpr int STAGES = 4;

template <int TN>
struct TCfg {
    static constexpr int NT = TN * 2;
    static constexpr int WN = TN / 32;
    static constexpr int A_BYTES = BM * BK * 8;       // 16384
    static constexpr int B_BYTES = TN * BK * 8;       // 8192 (TN = 64)
    static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
    static constexpr int SMEM = STAGES * STAGE_BYTES + 1024 /* alignment slack */ + 2 * STAGES * 8;
};

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(unsigned bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(unsigned bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(bar) : "memory");
}
__device__ __forceinline__ unsigned mbar_try_wait(unsigned bar, unsigned parity) {
    unsigned ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    return ok;
}
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity) {
    while (!mbar_try_wait(bar, parity)) {
    }
}
__device__ __forceinline__ void tma_load_2d(unsigned dst, const CUtensorMap* map, int c0, int c1, unsigned bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];\n" ::"r"(dst),
        "l"((unsigned long long)map), "r"(c0), "r"(c1), "r"(bar)
        : "memory");
}
__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

__device__ __forceinline__ int mapped_pos(const GemmArgs& p, int col0) {
    if (p.cyc_P <= 0) return col0;
    const int bw = p.cyc_tpb * 128;
    const int lb = col0 / bw;
    return (int)gpx_cyc_global(lb + p.cyc_q0, p.cyc_P, p.cyc_p, p.cyc_snake) * bw + col0 % bw - p.cyc_row_base;
}

__device__ __forceinline__ int rho8(int g) { return (g & 1) | (((g >> 1) & 1) << 2) | ((g >> 2) << 1); }

template <bool A_KM, bool B_KM, int TN>
__global__ void __launch_bounds__(TCfg<TN>::NT, 2)
dgemm_dmma_tma_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB, GemmArgs p) {
    using C_ = TCfg<TN>;
    extern __shared__ uint8_t smem_raw[];
    const int tid = threadIdx.x;
    const int bn = blockIdx.x;
    const int bm = p.rev_rows ? (gridDim.y - 1 - blockIdx.y) : blockIdx.y;
    const int row0 = bm * BM, col0 = bn * TN;
    const int gpos = mapped_pos(p, col0);
    const int brow0 = (p.cyc_P > 0 && !p.cyc_b_rows) ? col0 : gpos;
    if (p.lower_only && gpos >= row0 + BM) return;
    double* C = p.C;

    int kbeg = (p.kb_mode == 1 ? row0 : (p.kb_mode == 2 ? (col0 / 128) * 128 : (p.kb_mode == 3 ? (gpos / 128) * 128 : 0))) + p.kb_const;
    if (kbeg < 0) kbeg = 0;
    int kend = p.ke_mode == 1 ? row0 + BM : (p.ke_mode == 2 ? (col0 / 128) * 128 + 128 : p.K);
    if (kend > p.K) kend = p.K;
    const int nk = kend > kbeg ? (kend - kbeg) / BK : 0;

    // 1024-byte aligned stage buffers (SWIZZLE_128B atom = 8 rows x 128 B), barriers behind them
    const unsigned base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const unsigned bar_full = base + STAGES * C_::STAGE_BYTES;
    const unsigned bar_empty = bar_full + STAGES * 8;
    if (tid == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(bar_full + s * 8, 1);
            mbar_init(bar_empty + s * 8, C_::NT / 32);
        }
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    __syncthreads();

    const int warp = tid >> 5, lane = tid & 31;
    const int g = lane >> 2, t = lane & 3;
    const int wm0 = (warp / C_::WN) * 64, wn0 = (warp % C_::WN) * 32;

    double acc[8][4][2];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

    auto issue = [&](int slab) {   // thread 0 only
        const int s = slab % STAGES;
        const unsigned fb = bar_full + s * 8;
        mbar_expect_tx(fb, C_::STAGE_BYTES);
        const unsigned dstA = base + s * C_::STAGE_BYTES;
        const int k0 = kbeg + slab * BK;
        if (A_KM) {
            tma_load_2d(dstA, &mapA, k0, row0, fb);
        } else {
#pragma unroll
            for (int b = 0; b < BM / 16; ++b) tma_load_2d(dstA + b * 2048, &mapA, row0 + 16 * b, k0, fb);
        }
        if (B_KM) {
            tma_load_2d(dstA + C_::A_BYTES, &mapB, k0, brow0, fb);
        } else {
#pragma unroll
            for (int b = 0; b < TN / 16; ++b) tma_load_2d(dstA + C_::A_BYTES + b * 2048, &mapB, brow0 + 16 * b, k0, fb);
        }
    };
    if (tid == 0) {
        const int npre = nk < STAGES ? nk : STAGES;
        for (int s = 0; s < npre; ++s) issue(s);
    }

    // per-lane constant parts of the swizzled fragment addresses (generic pointers for plain LDS)
    const uint8_t* gbase = smem_raw + (base - smem_u32(smem_raw));
    const int rg = rho8(g);
    const int s0t = (t == 0) ? 0 : (t == 1 ? 3 : (t == 2 ? 4 : 7));
    // k-major: byte offset = (row)*128 + ((k>>1) ^ (row&7))*16 + (k&1)*8, row = w0 + 8i + rho(g)
    // k-strided: byte offset = (m>>4)*2048 + k*128 + (((m&15)>>1) ^ (k&7))*16 + (m&1)*8, m = w0 + 8i + g
    const int a_km_lane = (wm0 + rg) * 128, b_km_lane = (wn0 + rg) * 128;
    const int a_ks_lane = (wm0 >> 4) * 2048 + (g & 1) * 8, b_ks_lane = (wn0 >> 4) * 2048 + (g & 1) * 8;

    for (int kt = 0; kt < nk; ++kt) {
        if (tid == 0 && kt >= 2) {           // refill the stage consumed two slabs ago
            const int slab = kt - 2 + STAGES;
            if (slab < nk) {
                const int s = slab % STAGES;
                mbar_wait(bar_empty + s * 8, (unsigned)(((slab / STAGES) - 1) & 1));
                issue(slab);
            }
        }
        const int s = kt % STAGES;
        mbar_wait(bar_full + s * 8, (unsigned)((kt / STAGES) & 1));
        const uint8_t* a_s = gbase + s * C_::STAGE_BYTES;
        const uint8_t* b_s = a_s + C_::A_BYTES;
#pragma unroll
        for (int j4 = 0; j4 < 4; ++j4) {
            const int kap = 8 * (j4 >> 1) + (s0t ^ (j4 & 1));                       // this lane's k inside the slab
            const int sw_km = (((kap >> 1) ^ rg) << 4) + (kap & 1) * 8;             // k-major: swizzled chunk + half
            const int sw_ks0 = kap * 128 + ((((g >> 1)) ^ (kap & 7)) << 4);         // k-strided, even fragment (m&15 < 8)
            const int sw_ks1 = kap * 128 + (((4 + (g >> 1)) ^ (kap & 7)) << 4);     // k-strided, odd fragment
            double af[8], bf[4];
#pragma unroll
            for (int i = 0; i < 8; ++i)
                af[i] = A_KM ? *reinterpret_cast<const double*>(a_s + a_km_lane + i * 1024 + sw_km)
                             : *reinterpret_cast<const double*>(a_s + a_ks_lane + (i >> 1) * 2048 + ((i & 1) ? sw_ks1 : sw_ks0));
#pragma unroll
            for (int j = 0; j < 4; ++j)
                bf[j] = B_KM ? *reinterpret_cast<const double*>(b_s + b_km_lane + j * 1024 + sw_km)
                             : *reinterpret_cast<const double*>(b_s + b_ks_lane + (j >> 1) * 2048 + ((j & 1) ? sw_ks1 : sw_ks0));
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) dmma(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_empty + s * 8);
    }

    const double alpha = p.alpha, beta = p.beta;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int r = row0 + wm0 + i * 8 + (A_KM ? rg : g);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int c = col0 + wn0 + j * 8 + (B_KM ? rho8(2 * t) : 2 * t);
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

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn g_encode = nullptr;
int g_tma_state_dev[GPX_MAX_DEVICES] = {};   // per device: 0 unknown, 1 usable, -1 unavailable / disabled

int tma_init(int device) {
    int& g_tma_state = g_tma_state_dev[device % GPX_MAX_DEVICES];
    if (g_tma_state != 0) return g_tma_state;
    const char* e = getenv("GPX_GEMM_TMA");
    if (e && atoi(e) == 0) return g_tma_state = -1;
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) != cudaSuccess || !fn ||
        q != cudaDriverEntryPointSuccess) {
        cudaGetLastError();
        return g_tma_state = -1;
    }
    g_encode = (EncodeTiledFn)fn;
    bool ok = true;
#define CFG(a, b)                                                                                                              \
    ok = ok && cudaFuncSetAttribute(dgemm_dmma_tma_kernel<a, b, 64>, cudaFuncAttributeMaxDynamicSharedMemorySize, TCfg<64>::SMEM) == cudaSuccess && \
         cudaFuncSetAttribute(dgemm_dmma_tma_kernel<a, b, 64>, cudaFuncAttributePreferredSharedMemoryCarveout, 100) == cudaSuccess;
    CFG(true, true) CFG(true, false) CFG(false, true) CFG(false, false)
#undef CFG
    if (!ok) {
        cudaGetLastError();
        return g_tma_state = -1;
    }
    return g_tma_state = 1;
}

// k-strided operand: K x cols doubles with leading dimension ld; box = 16 (cols) x 16 (k)
int make_map_ks(CUtensorMap* m, const double* ptr, int64_t cols, int64_t K, int64_t ld) {
    cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)K};
    cuuint64_t strides[1] = {(cuuint64_t)ld * sizeof(double)};
    cuuint32_t box[2] = {16, (cuuint32_t)BK};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = g_encode(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, (void*)ptr, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                          CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? 0 : -1;
}

// k-major operand: rows x K doubles with leading dimension ld; box = 16 (k) x box_rows
int make_map(CUtensorMap* m, const double* ptr, int64_t rows, int64_t K, int64_t ld, int box_rows) {
    cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)ld * sizeof(double)};
    cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = g_encode(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, (void*)ptr, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                          CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? 0 : -1;
}

}  // namespace

// Returns 1 if the launch was done through the TMA kernel, 0 if the caller must use the cp.async kernel, <0 on error.
int gpx_gemm_tma_try_launch(gpx_ctx* h, const GemmArgs& a, double flops_exec) {
    if (a.batch > 1 || a.batch2 > 1 || a.C == a.A || a.C == a.B) return 0;
    if ((a.lda % 2) || (a.ldb % 2)) return 0;
    if (tma_init(h->device) != 1) return 0;
    constexpr int TN = 64;
    alignas(64) CUtensorMap mapA, mapB;
    const int64_t rowsB = a.cyc_P > 0 && a.cyc_b_rows ? a.M : a.N;
    if ((a.a_kmajor ? make_map(&mapA, a.A, a.M, a.K, a.lda, BM) : make_map_ks(&mapA, a.A, a.M, a.K, a.lda)) != 0) return 0;
    if ((a.b_kmajor ? make_map(&mapB, a.B, rowsB, a.K, a.ldb, TN) : make_map_ks(&mapB, a.B, rowsB, a.K, a.ldb)) != 0) return 0;
    dim3 grid(a.N / TN, a.M / BM, 1);
    if (h->timing_on) gpx_timing_gemm_begin(h, flops_exec, a.M, a.N, a.K);
    if (a.a_kmajor && a.b_kmajor) dgemm_dmma_tma_kernel<true, true, TN><<<grid, TCfg<TN>::NT, TCfg<TN>::SMEM, h->stream>>>(mapA, mapB, a);
    else if (a.a_kmajor) dgemm_dmma_tma_kernel<true, false, TN><<<grid, TCfg<TN>::NT, TCfg<TN>::SMEM, h->stream>>>(mapA, mapB, a);
    else if (a.b_kmajor) dgemm_dmma_tma_kernel<false, true, TN><<<grid, TCfg<TN>::NT, TCfg<TN>::SMEM, h->stream>>>(mapA, mapB, a);
    else dgemm_dmma_tma_kernel<false, false, TN><<<grid, TCfg<TN>::NT, TCfg<TN>::SMEM, h->stream>>>(mapA, mapB, a);
    GPX_CHECK_LAUNCH(h);
    gpx_timing_gemm_end(h);
    return 1;
}
