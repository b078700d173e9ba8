// Optional CUDA-event instrumentation: per-launch timing of the DMMA GEMM kernel and per-phase timing of
// the fused drivers.  Off by default (zero overhead); bench.py switches it on for the roofline numbers.
#include <vector>
#include "common.cuh"

struct gpx_timing {
    std::vector<cudaEvent_t> pool;      // reusable events
    size_t used = 0;
    std::vector<std::pair<size_t, size_t>> gemm_pairs;   // (start,end) event indices
    std::vector<std::pair<size_t, size_t>> leaf_pairs;   // potrf leaf launches
    double gemm_flops_exec = 0.0;       // flops the kernel actually executes (tile-granular k ranges)
    std::vector<std::pair<int, size_t>> marks;            // (phase id, event index)
    struct Shape { int M, N, K, phase; double flops; };
    std::vector<Shape> gemm_shapes;                       // one per gemm_pairs entry
    int cur_phase = -1;
};

static cudaEvent_t next_event(gpx_timing* t, size_t* idx) {
    if (t->used == t->pool.size()) {
        cudaEvent_t e;
        cudaEventCreate(&e);
        t->pool.push_back(e);
    }
    *idx = t->used;
    return t->pool[t->used++];
}

void gpx_timing_gemm_begin(gpx_ctx* h, double flops_exec, int M, int N, int K) {
    if (!h->timing_on) return;
    gpx_timing* t = (gpx_timing*)h->timing;
    size_t i0;
    cudaEvent_t e0 = next_event(t, &i0);
    cudaEventRecord(e0, h->stream);
    t->gemm_pairs.push_back({i0, (size_t)-1});
    t->gemm_shapes.push_back({M, N, K, t->cur_phase, flops_exec});
    t->gemm_flops_exec += flops_exec;
}

void gpx_timing_gemm_end(gpx_ctx* h) {
    if (!h->timing_on) return;
    gpx_timing* t = (gpx_timing*)h->timing;
    size_t i1;
    cudaEvent_t e1 = next_event(t, &i1);
    cudaEventRecord(e1, h->stream);
    t->gemm_pairs.back().second = i1;
}

void gpx_timing_leaf_begin(gpx_ctx* h) {
    if (!h->timing_on) return;
    gpx_timing* t = (gpx_timing*)h->timing;
    size_t i0;
    cudaEventRecord(next_event(t, &i0), h->stream);
    t->leaf_pairs.push_back({i0, (size_t)-1});
}

void gpx_timing_leaf_end(gpx_ctx* h) {
    if (!h->timing_on) return;
    gpx_timing* t = (gpx_timing*)h->timing;
    size_t i1;
    cudaEventRecord(next_event(t, &i1), h->stream);
    t->leaf_pairs.back().second = i1;
}

void gpx_phase_mark(gpx_ctx* h, int phase) {
    if (!h->timing_on) return;
    gpx_timing* t = (gpx_timing*)h->timing;
    size_t i;
    cudaEvent_t e = next_event(t, &i);
    cudaEventRecord(e, h->stream);
    t->marks.push_back({phase, i});
    t->cur_phase = phase;
}

extern "C" int gpx_timing_enable(gpx_handle h, int on) {
    GPX_ENTER(h);
    if (!h->timing) h->timing = new gpx_timing();
    gpx_timing* t = (gpx_timing*)h->timing;
    t->used = 0;
    t->gemm_pairs.clear();
    t->gemm_shapes.clear();
    t->cur_phase = -1;
    t->leaf_pairs.clear();
    t->marks.clear();
    t->gemm_flops_exec = 0.0;
    h->timing_on = on ? 1 : 0;
    return 0;
}

// out[0] = total DMMA-GEMM kernel ms, out[1] = #GEMM launches, out[2] = flops executed by those launches,
// out[3 + p] = ms spent in phase p (time from mark p to the next mark), p < GPX_NPHASES;
// out[11] = potrf-leaf kernel ms, out[12] = #leaf launches (when nout >= 13).
extern "C" int gpx_timing_collect(gpx_handle h, double* out, int nout) {
    GPX_REQUIRE(h != nullptr && h->timing != nullptr, 1);
    GPX_REQUIRE(nout >= 3 + GPX_NPHASES, 3);
    gpx_timing* t = (gpx_timing*)h->timing;
    GPX_CUDA(cudaStreamSynchronize(h->stream));
    for (int i = 0; i < nout; ++i) out[i] = 0.0;
    for (auto& pr : t->gemm_pairs) {
        if (pr.second == (size_t)-1) continue;
        float ms = 0.f;
        GPX_CUDA(cudaEventElapsedTime(&ms, t->pool[pr.first], t->pool[pr.second]));
        out[0] += ms;
    }
    if (nout >= 3 + GPX_NPHASES + 2) {
        for (auto& pr : t->leaf_pairs) {
            if (pr.second == (size_t)-1) continue;
            float ms = 0.f;
            GPX_CUDA(cudaEventElapsedTime(&ms, t->pool[pr.first], t->pool[pr.second]));
            out[3 + GPX_NPHASES] += ms;
        }
        out[3 + GPX_NPHASES + 1] = (double)t->leaf_pairs.size();
    }
    out[1] = (double)t->gemm_pairs.size();
    out[2] = t->gemm_flops_exec;
    for (size_t m = 0; m + 1 < t->marks.size(); ++m) {
        float ms = 0.f;
        GPX_CUDA(cudaEventElapsedTime(&ms, t->pool[t->marks[m].second], t->pool[t->marks[m + 1].second]));
        int p = t->marks[m].first;
        if (p >= 0 && p < GPX_NPHASES) out[3 + p] += ms;
    }
    return 0;
}

// Per-launch CSV of the instrumented region (call after gpx_timing_collect, before the next gpx_timing_enable):
// index, phase, M, N, K, flops executed, ms, start offset (ms after the first recorded event).
extern "C" int gpx_timing_dump(gpx_handle h, const char* path) {
    GPX_REQUIRE(h != nullptr && h->timing != nullptr && path != nullptr, 1);
    gpx_timing* t = (gpx_timing*)h->timing;
    FILE* f = fopen(path, "w");
    GPX_REQUIRE(f != nullptr, 2);
    fprintf(f, "idx,phase,M,N,K,flops,ms,start_ms\n");
    for (size_t i = 0; i < t->gemm_pairs.size(); ++i) {
        auto& pr = t->gemm_pairs[i];
        if (pr.second == (size_t)-1) continue;
        float ms = 0.f, st = 0.f;
        cudaEventElapsedTime(&ms, t->pool[pr.first], t->pool[pr.second]);
        cudaEventElapsedTime(&st, t->pool[0], t->pool[pr.first]);
        const auto& sh = t->gemm_shapes[i];
        fprintf(f, "%zu,%d,%d,%d,%d,%.0f,%.4f,%.4f\n", i, sh.phase, sh.M, sh.N, sh.K, sh.flops, ms, st);
    }
    fclose(f);
    return 0;
}

void gpx_timing_destroy(gpx_ctx* h) {
    if (!h->timing) return;
    gpx_timing* t = (gpx_timing*)h->timing;
    for (auto e : t->pool) cudaEventDestroy(e);
    delete t;
    h->timing = nullptr;
}
