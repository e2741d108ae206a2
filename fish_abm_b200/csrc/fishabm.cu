// libfishabm.so: C ABI (include/fishabm.h) over the sm_100a kernels in kernels.cuh.
// No CPU fallback: every operation on fish state runs on the device; without a device fishabm_create fails.
#include "../../include/fishabm.h"

#include <cuda_runtime.h>
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <string>
#include <vector>

#include "kernels.cuh"
#include "neighbour_v2.cuh"

using namespace fishk;

namespace {

thread_local std::string g_create_error;

struct Buffers {  // one set of cell-sorted arrays
    float4* rec = nullptr;
    float4* cold0 = nullptr;
    int4* cold1 = nullptr;
    int* cell = nullptr;
};

}  // namespace

struct fishabm_ctx {
    fishabm_params p;
    std::string err;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> nb_events;  // around each neighbour kernel of the last call
    int nb_events_used = 0;
    int n = 0, nx = 0, ny = 0, ncell = 0, qrows = 0, qcols = 0;
    double speed_norm = 1.0;
    uint32_t step = 0;
    bool have_state = false;
    bool pending_sample = false;  // sample() of step `step-1` not yet applied (it shares the next move's neighbour pass)
    bool stats_on = false;
    int nb_variant = 2;  // neighbour kernel design: 2 = candidate-parallel (default), 1 = focal-parallel (FISHABM_NEIGHBOUR=v1)
    uint64_t launches = 0;
    float last_ms = 0.f;

    Buffers buf[2];
    int cur = 0;
    int *next_cell = nullptr, *next_rank = nullptr, *cell_count = nullptr, *cell_start = nullptr, *tile_sums = nullptr;
    int4* cnt = nullptr;
    float* turn = nullptr;
    uint32_t* aux = nullptr;
    uint8_t* quality = nullptr;
    int *upd_count = nullptr, *upd_idx = nullptr;
    uint8_t* upd_val = nullptr;
    unsigned long long* stats = nullptr;  // 4 pass counters + 1 scratch for sums
    int* work_counter = nullptr;
    int nb_grid = 0;
    int* bad = nullptr;
    int* pos_of_id = nullptr;
    // host<->device staging in the reference's layout
    double *st_coords = nullptr, *st_dir = nullptr, *st_speed = nullptr, *st_sf = nullptr;
    int *st_lag = nullptr, *st_rate = nullptr, *st_intake = nullptr, *st_i4 = nullptr;
    uint8_t* h_q8 = nullptr;
};

namespace {

int fail(fishabm_ctx* c, int code, const char* fmt, ...) {
    char b[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(b, sizeof b, fmt, ap);
    va_end(ap);
    if (c) c->err = b; else g_create_error = b;
    return code;
}

#define CK(call)                                                                                          \
    do {                                                                                                  \
        cudaError_t e_ = (call);                                                                          \
        if (e_ != cudaSuccess) return fail(ctx, FISHABM_E_CUDA, "%s: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
    } while (0)

template <class T> cudaError_t dalloc(T** p, size_t count) { return cudaMalloc((void**)p, count * sizeof(T) > 0 ? count * sizeof(T) : 1); }

inline int cdiv(int a, int b) { return (a + b - 1) / b; }

// ---- the counting sort: counts (already accumulated in cell_count) -> cell_start; scatter cur -> other
int rebuild(fishabm_ctx* ctx) {
    const int ntiles = cdiv(ctx->ncell, SCAN_TILE);
    scan_tile_sums_kernel<<<ntiles, SCAN_THREADS, 0, ctx->stream>>>(ctx->cell_count, ctx->ncell, ctx->tile_sums);
    scan_tile_offsets_kernel<<<1, SCAN_THREADS, 0, ctx->stream>>>(ctx->tile_sums, ntiles);
    scan_finish_kernel<<<ntiles, SCAN_THREADS, 0, ctx->stream>>>(ctx->cell_count, ctx->ncell, ctx->tile_sums, ctx->cell_start);
    const Buffers& in = ctx->buf[ctx->cur];
    const Buffers& out = ctx->buf[ctx->cur ^ 1];
    ScatterArgs s{in.rec, in.cold0, in.cold1, out.rec, out.cold0, out.cold1, out.cell, ctx->next_cell, ctx->next_rank, ctx->cell_start, ctx->n};
    scatter_kernel<<<cdiv(ctx->n, 256), 256, 0, ctx->stream>>>(s);
    ctx->launches += 4;
    ctx->cur ^= 1;
    CK(cudaGetLastError());
    return 0;
}

enum { PASS_COUNTS = 1, PASS_TURN = 2 };

// K4: one neighbour pass.  active_lag_max < 0 evaluates every fish (frozen-state queries).
int neighbour_pass(fishabm_ctx* ctx, int active_lag_max, bool do_sample, uint32_t step_sample, uint32_t step_move, bool timed) {
    const Buffers& b = ctx->buf[ctx->cur];
    NeighArgs a;
    a.rec = b.rec; a.cold0 = b.cold0; a.cold1 = b.cold1; a.cell_start = ctx->cell_start;
    a.cnt = ctx->cnt; a.turn = ctx->turn; a.aux = ctx->aux; a.stats = ctx->stats_on ? ctx->stats : nullptr;
    a.ncell = ctx->ncell; a.nx = ctx->nx; a.ny = ctx->ny;
    a.active_lag_max = active_lag_max; a.do_sample = do_sample ? 1 : 0;
    a.step_sample = step_sample; a.step_move = step_move; a.seed = ctx->p.seed; a.replicate = 0;
    a.work_counter = ctx->work_counter;
    if (ctx->stats_on) CK(cudaMemsetAsync(ctx->stats, 0, 4 * sizeof(unsigned long long), ctx->stream));
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    if (timed) {
        if (ctx->nb_events_used == (int)ctx->nb_events.size()) {
            cudaEvent_t x, y;
            CK(cudaEventCreate(&x)); CK(cudaEventCreate(&y));
            ctx->nb_events.emplace_back(x, y);
        }
        e0 = ctx->nb_events[ctx->nb_events_used].first; e1 = ctx->nb_events[ctx->nb_events_used].second;
        ctx->nb_events_used++;
        CK(cudaEventRecord(e0, ctx->stream));
    }
    if (ctx->nb_variant == 1) {
        const int grid = cdiv(ctx->ncell, NB_WARPS);
        if (ctx->stats_on) neighbour_kernel<true><<<grid, NB_THREADS, 0, ctx->stream>>>(a);
        else neighbour_kernel<false><<<grid, NB_THREADS, 0, ctx->stream>>>(a);
    } else {
        CK(cudaMemsetAsync(ctx->work_counter, 0, sizeof(int), ctx->stream));
        const int grid = std::min(ctx->nb_grid, cdiv(ctx->ncell, V2_WARPS));
        if (ctx->stats_on) neighbour_kernel_v2<true><<<grid, V2_THREADS, V2_SMEM_BYTES, ctx->stream>>>(a);
        else neighbour_kernel_v2<false><<<grid, V2_THREADS, V2_SMEM_BYTES, ctx->stream>>>(a);
    }
    ctx->launches += 1;
    if (timed) CK(cudaEventRecord(e1, ctx->stream));
    CK(cudaGetLastError());
    return 0;
}

// K5 (+ deferred food-grid writes, + re-sort when fish moved)
int update_pass(fishabm_ctx* ctx, bool do_sample, bool do_move, uint32_t step_move) {
    const Buffers& b = ctx->buf[ctx->cur];
    UpdateArgs u;
    u.rec = b.rec; u.cold0 = b.cold0; u.cold1 = b.cold1; u.cell = b.cell; u.cell_start = ctx->cell_start;
    u.cnt = ctx->cnt; u.turn = ctx->turn; u.aux = ctx->aux;
    u.next_cell = ctx->next_cell; u.next_rank = ctx->next_rank; u.next_count = ctx->cell_count;
    u.quality = ctx->quality; u.upd_count = ctx->upd_count; u.upd_idx = ctx->upd_idx; u.upd_val = ctx->upd_val;
    u.n = ctx->n; u.nx = ctx->nx; u.ny = ctx->ny; u.qcols = ctx->qcols;
    u.do_sample = do_sample; u.do_move = do_move; u.rolling = ctx->p.rolling;
    u.step_move = step_move; u.seed = ctx->p.seed; u.replicate = 0; u.speed_norm = ctx->speed_norm;
    update_kernel<<<cdiv(ctx->n, 256), 256, 0, ctx->stream>>>(u);
    ctx->launches += 1;
    if (do_sample) {
        apply_quality_updates_kernel<<<64, 256, 0, ctx->stream>>>(ctx->quality, ctx->upd_count, ctx->upd_idx, ctx->upd_val);
        reset_counter_kernel<<<1, 1, 0, ctx->stream>>>(ctx->upd_count);
        ctx->launches += 2;
    }
    CK(cudaGetLastError());
    if (do_move) return rebuild(ctx);
    return 0;
}

int begin_timed(fishabm_ctx* ctx) {
    ctx->nb_events_used = 0;
    CK(cudaEventRecord(ctx->ev0, ctx->stream));
    return 0;
}
int end_timed(fishabm_ctx* ctx) {
    CK(cudaEventRecord(ctx->ev1, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    CK(cudaEventElapsedTime(&ctx->last_ms, ctx->ev0, ctx->ev1));
    return 0;
}

int need_state(fishabm_ctx* ctx) {
    if (!ctx) return FISHABM_E_ARG;
    if (!ctx->have_state) return fail(ctx, FISHABM_E_STATE, "no state: call fishabm_set_state or fishabm_initialize first");
    return 0;
}

// fishabm_step leaves the last step's sample() pending so that it can share the next move's neighbour pass;
// anything that observes or replaces state completes it first.
int flush_pending(fishabm_ctx* ctx, bool timed = false) {
    if (!ctx->pending_sample) return 0;
    int rc;
    if ((rc = neighbour_pass(ctx, 0, true, ctx->step - 1, ctx->step - 1, timed))) return rc;
    if ((rc = update_pass(ctx, true, false, ctx->step - 1))) return rc;
    ctx->pending_sample = false;
    return 0;
}

}  // namespace

extern "C" {

int fishabm_params_default(fishabm_params* p) {
    if (!p) return FISHABM_E_ARG;
    memset(p, 0, sizeof *p);
    p->struct_size = (int32_t)sizeof *p;
    p->n = 60;           // #define num 60, fish_mod.cpp:23
    p->w = p->h = 2000;  // Structure, fish_mod.cpp:53-59
    p->mode = FISHABM_MODE_CELL_FP32;
    p->device = 0;
    p->replicates = 1;
    p->rolling = 1;
    p->seed = 0;
    p->speed_norm = 0.0;
    p->rank = 0; p->nranks = 1; p->n_global = 0; p->capacity = 0;
    return 0;
}

const char* fishabm_last_error(const fishabm_ctx* ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }

int fishabm_create(const fishabm_params* p, fishabm_ctx** out) {
    fishabm_ctx* ctx = nullptr;  // for CK/fail before the context exists
    if (!p || !out) return fail(nullptr, FISHABM_E_ARG, "null argument");
    *out = nullptr;
    if (p->struct_size != (int32_t)sizeof(fishabm_params)) return fail(nullptr, FISHABM_E_ARG, "fishabm_params.struct_size mismatch: use fishabm_params_default");
    if (p->n < 1 || p->n >= (1 << 28)) return fail(nullptr, FISHABM_E_ARG, "n must be in [1, 2^28)");
    if (p->w % CELL_I || p->h % CELL_I || p->w < 3 * CELL_I || p->h < 3 * CELL_I)
        return fail(nullptr, FISHABM_E_ARG, "w and h must be multiples of 200 and at least 600 (got %d x %d)", p->w, p->h);
    if (p->mode != FISHABM_MODE_CELL_FP32) return fail(nullptr, FISHABM_E_ARG, "mode %d is not available in this build", p->mode);
    if (p->replicates != 1) return fail(nullptr, FISHABM_E_ARG, "replicates > 1: use the ensemble entry points");
    if (p->nranks != 1) return fail(nullptr, FISHABM_E_ARG, "nranks > 1: use the slab entry points");
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev < 1) return fail(nullptr, FISHABM_E_CUDA, "no CUDA device (%s): libfishabm has no CPU path", cudaGetErrorString(e));
    if (p->device < 0 || p->device >= ndev) return fail(nullptr, FISHABM_E_ARG, "device %d out of range (%d devices)", p->device, ndev);
    e = cudaSetDevice(p->device);
    if (e != cudaSuccess) return fail(nullptr, FISHABM_E_CUDA, "cudaSetDevice: %s", cudaGetErrorString(e));

    fishabm_ctx* c = new (std::nothrow) fishabm_ctx();
    if (!c) return fail(nullptr, FISHABM_E_NOMEM, "out of host memory");
    c->p = *p;
    c->n = p->n;
    c->nx = p->w / CELL_I; c->ny = p->h / CELL_I; c->ncell = c->nx * c->ny;
    c->qrows = p->h / FOOD_I; c->qcols = p->w / FOOD_I;
    c->speed_norm = p->speed_norm > 0 ? p->speed_norm : (double)p->n;
    const size_t n = (size_t)c->n, nq = (size_t)c->qrows * c->qcols;
    bool ok = true;
    auto A = [&](cudaError_t r) { if (r != cudaSuccess) { ok = false; g_create_error = cudaGetErrorString(r); } };
    A(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    A(cudaEventCreate(&c->ev0)); A(cudaEventCreate(&c->ev1));
    for (int k = 0; k < 2 && ok; ++k) {
        A(dalloc(&c->buf[k].rec, n)); A(dalloc(&c->buf[k].cold0, n)); A(dalloc(&c->buf[k].cold1, n)); A(dalloc(&c->buf[k].cell, n));
    }
    A(dalloc(&c->next_cell, n)); A(dalloc(&c->next_rank, n));
    A(dalloc(&c->cell_count, (size_t)c->ncell + 1)); A(dalloc(&c->cell_start, (size_t)c->ncell + 1));
    A(dalloc(&c->tile_sums, (size_t)cdiv(c->ncell, SCAN_TILE) + 1));
    A(dalloc(&c->cnt, n)); A(dalloc(&c->turn, n)); A(dalloc(&c->aux, n));
    A(dalloc(&c->quality, nq));
    A(dalloc(&c->upd_count, 1)); A(dalloc(&c->upd_idx, n)); A(dalloc(&c->upd_val, n));
    A(dalloc(&c->stats, 8)); A(dalloc(&c->work_counter, 1)); A(dalloc(&c->bad, 1)); A(dalloc(&c->pos_of_id, n));
    A(dalloc(&c->st_coords, 2 * n)); A(dalloc(&c->st_dir, n)); A(dalloc(&c->st_speed, n)); A(dalloc(&c->st_sf, n));
    A(dalloc(&c->st_lag, n)); A(dalloc(&c->st_rate, n)); A(dalloc(&c->st_intake, n)); A(dalloc(&c->st_i4, 4 * n));
    if (ok) {
        A(cudaMemset(c->cell_count, 0, ((size_t)c->ncell + 1) * sizeof(int)));
        A(cudaMemset(c->quality, 0, nq));
        A(cudaMemset(c->upd_count, 0, sizeof(int)));
        A(cudaMemset(c->stats, 0, 8 * sizeof(unsigned long long)));
        A(cudaMemset(c->aux, 0, n * sizeof(uint32_t)));
    }
    if (ok) {
        A(cudaFuncSetAttribute(neighbour_kernel_v2<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)V2_SMEM_BYTES));
        A(cudaFuncSetAttribute(neighbour_kernel_v2<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)V2_SMEM_BYTES));
    }
    if (ok) {
        int per_sm = 0, sms = 0;
        A(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, neighbour_kernel_v2<false>, V2_THREADS, V2_SMEM_BYTES));
        A(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, p->device));
        c->nb_grid = std::max(1, per_sm) * std::max(1, sms);  // persistent: one resident wave
    }
    const char* variant = getenv("FISHABM_NEIGHBOUR");
    if (variant && strcmp(variant, "v1") == 0) c->nb_variant = 1;
    c->h_q8 = (uint8_t*)malloc(nq > 0 ? nq : 1);
    if (!ok || !c->h_q8) {
        std::string msg = "device allocation failed: " + g_create_error;
        fishabm_destroy(c);
        return fail(nullptr, FISHABM_E_NOMEM, "%s", msg.c_str());
    }
    (void)ctx;
    *out = c;
    return 0;
}

void fishabm_destroy(fishabm_ctx* c) {
    if (!c) return;
    cudaSetDevice(c->p.device);
    if (c->stream) cudaStreamSynchronize(c->stream);
    for (int k = 0; k < 2; ++k) { cudaFree(c->buf[k].rec); cudaFree(c->buf[k].cold0); cudaFree(c->buf[k].cold1); cudaFree(c->buf[k].cell); }
    cudaFree(c->next_cell); cudaFree(c->next_rank); cudaFree(c->cell_count); cudaFree(c->cell_start); cudaFree(c->tile_sums);
    cudaFree(c->cnt); cudaFree(c->turn); cudaFree(c->aux); cudaFree(c->quality);
    cudaFree(c->upd_count); cudaFree(c->upd_idx); cudaFree(c->upd_val); cudaFree(c->stats); cudaFree(c->work_counter); cudaFree(c->bad); cudaFree(c->pos_of_id);
    cudaFree(c->st_coords); cudaFree(c->st_dir); cudaFree(c->st_speed); cudaFree(c->st_sf);
    cudaFree(c->st_lag); cudaFree(c->st_rate); cudaFree(c->st_intake); cudaFree(c->st_i4);
    for (auto& pr : c->nb_events) { cudaEventDestroy(pr.first); cudaEventDestroy(pr.second); }
    if (c->ev0) cudaEventDestroy(c->ev0);
    if (c->ev1) cudaEventDestroy(c->ev1);
    if (c->stream) cudaStreamDestroy(c->stream);
    free(c->h_q8);
    delete c;
}

int fishabm_set_state(fishabm_ctx* ctx, const double* coords, const double* dir, const double* speed, const double* speed_factor,
                      const int32_t* lag, const int32_t* sample_rate, const int32_t* food_intake) {
    if (!ctx || !coords || !dir || !speed || !speed_factor || !lag || !sample_rate || !food_intake)
        return fail(ctx, FISHABM_E_ARG, "fishabm_set_state: null argument");
    CK(cudaSetDevice(ctx->p.device));
    if (ctx->have_state) { int rcf = flush_pending(ctx); if (rcf) return rcf; }
    ctx->pending_sample = false;
    const size_t n = (size_t)ctx->n;
    cudaStream_t s = ctx->stream;
    CK(cudaMemcpyAsync(ctx->st_coords, coords, 2 * n * sizeof(double), cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(ctx->st_dir, dir, n * sizeof(double), cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(ctx->st_speed, speed, n * sizeof(double), cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(ctx->st_sf, speed_factor, n * sizeof(double), cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(ctx->st_lag, lag, n * sizeof(int), cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(ctx->st_rate, sample_rate, n * sizeof(int), cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(ctx->st_intake, food_intake, n * sizeof(int), cudaMemcpyHostToDevice, s));
    CK(cudaMemsetAsync(ctx->bad, 0, sizeof(int), s));
    CK(cudaMemsetAsync(ctx->cell_count, 0, ((size_t)ctx->ncell + 1) * sizeof(int), s));
    const Buffers& b = ctx->buf[ctx->cur];
    IngestArgs a{ctx->st_coords, ctx->st_dir, ctx->st_speed, ctx->st_sf, ctx->st_lag, ctx->st_rate, ctx->st_intake,
                 b.rec, b.cold0, b.cold1, ctx->next_cell, ctx->next_rank, ctx->cell_count,
                 ctx->n, ctx->nx, ctx->ny, 0, ctx->bad, (double)ctx->p.w, (double)ctx->p.h};
    ingest_kernel<<<cdiv(ctx->n, 256), 256, 0, s>>>(a);
    ctx->launches += 1;
    int rc = rebuild(ctx);
    if (rc) return rc;
    int bad = 0;
    CK(cudaMemcpyAsync(&bad, ctx->bad, sizeof(int), cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    if (bad) { ctx->have_state = false; return fail(ctx, FISHABM_E_ARG, "fishabm_set_state: coordinates must be finite and inside [0,w] x [0,h]"); }
    ctx->have_state = true;
    return 0;
}

int fishabm_get_state(fishabm_ctx* ctx, double* coords, double* dir, double* speed, double* speed_factor, int32_t* lag,
                      int32_t* sample_rate, int32_t* food_intake) {
    int rc = need_state(ctx);
    if (rc) return rc;
    CK(cudaSetDevice(ctx->p.device));
    { int rcf = flush_pending(ctx); if (rcf) return rcf; }
    const size_t n = (size_t)ctx->n;
    cudaStream_t s = ctx->stream;
    const Buffers& b = ctx->buf[ctx->cur];
    ExportArgs a{b.rec, b.cold0, b.cold1, b.cell,
                 coords ? ctx->st_coords : nullptr, dir ? ctx->st_dir : nullptr, speed ? ctx->st_speed : nullptr,
                 speed_factor ? ctx->st_sf : nullptr, lag ? ctx->st_lag : nullptr, sample_rate ? ctx->st_rate : nullptr,
                 food_intake ? ctx->st_intake : nullptr, ctx->n, ctx->nx, 0};
    export_kernel<<<cdiv(ctx->n, 256), 256, 0, s>>>(a);
    ctx->launches += 1;
    CK(cudaGetLastError());
    if (coords) CK(cudaMemcpyAsync(coords, ctx->st_coords, 2 * n * sizeof(double), cudaMemcpyDeviceToHost, s));
    if (dir) CK(cudaMemcpyAsync(dir, ctx->st_dir, n * sizeof(double), cudaMemcpyDeviceToHost, s));
    if (speed) CK(cudaMemcpyAsync(speed, ctx->st_speed, n * sizeof(double), cudaMemcpyDeviceToHost, s));
    if (speed_factor) CK(cudaMemcpyAsync(speed_factor, ctx->st_sf, n * sizeof(double), cudaMemcpyDeviceToHost, s));
    if (lag) CK(cudaMemcpyAsync(lag, ctx->st_lag, n * sizeof(int), cudaMemcpyDeviceToHost, s));
    if (sample_rate) CK(cudaMemcpyAsync(sample_rate, ctx->st_rate, n * sizeof(int), cudaMemcpyDeviceToHost, s));
    if (food_intake) CK(cudaMemcpyAsync(food_intake, ctx->st_intake, n * sizeof(int), cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    return 0;
}

int fishabm_initialize(fishabm_ctx* ctx, double speed) {
    if (!ctx) return FISHABM_E_ARG;
    // initialize(), fish_mod.cpp:183-206.  Tiny and one-off: the state is composed on the host from the same
    // counter-based streams the device uses, then ingested like any other state.
    const int n = ctx->n;
    std::vector<double> coords(2 * (size_t)n), dir(n), sp(n, speed), sf(n, 1.0);
    std::vector<int32_t> z(n, 0);
    for (int i = 0; i < n; ++i) {
        dir[i] = (double)fishrng::draw(ctx->p.seed, 0, 0, (uint32_t)i, STREAM_RANDOMWALK, 0, -45, 45);
        coords[2 * (size_t)i] = ctx->p.w / 2 + fishrng::draw(ctx->p.seed, 0, 0, (uint32_t)i, STREAM_INIT, 0, 0, 200) - 100;
        coords[2 * (size_t)i + 1] = ctx->p.h / 2 + fishrng::draw(ctx->p.seed, 0, 1, (uint32_t)i, STREAM_INIT, 0, 0, 200) - 100;
    }
    return fishabm_set_state(ctx, coords.data(), dir.data(), sp.data(), sf.data(), z.data(), z.data(), z.data());
}

int fishabm_set_quality(fishabm_ctx* ctx, const int32_t* q, int32_t rows, int32_t cols, int32_t row_stride) {
    if (!ctx || !q) return fail(ctx, FISHABM_E_ARG, "fishabm_set_quality: null argument");
    if (rows != ctx->qrows || cols != ctx->qcols || row_stride < cols)
        return fail(ctx, FISHABM_E_ARG, "fishabm_set_quality: grid must be %d x %d (h/20 x w/20), got %d x %d", ctx->qrows, ctx->qcols, rows, cols);
    for (int r = 0; r < rows; ++r)
        for (int c = 0; c < cols; ++c) {
            const int v = q[(size_t)r * row_stride + c];
            if (v < 0 || v > 255) return fail(ctx, FISHABM_E_ARG, "fishabm_set_quality: value %d at (%d,%d) outside 0..255", v, r, c);
            ctx->h_q8[(size_t)r * cols + c] = (uint8_t)v;
        }
    CK(cudaSetDevice(ctx->p.device));
    { int rcf = flush_pending(ctx); if (rcf) return rcf; }
    CK(cudaMemcpyAsync(ctx->quality, ctx->h_q8, (size_t)rows * cols, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return 0;
}

int fishabm_get_quality(fishabm_ctx* ctx, int32_t replicate, int32_t* q, int32_t rows, int32_t cols, int32_t row_stride) {
    if (!ctx || !q) return fail(ctx, FISHABM_E_ARG, "fishabm_get_quality: null argument");
    if (replicate != 0 || rows != ctx->qrows || cols != ctx->qcols || row_stride < cols)
        return fail(ctx, FISHABM_E_ARG, "fishabm_get_quality: grid is %d x %d, replicate 0", ctx->qrows, ctx->qcols);
    CK(cudaSetDevice(ctx->p.device));
    { int rcf = flush_pending(ctx); if (rcf) return rcf; }
    CK(cudaMemcpyAsync(ctx->h_q8, ctx->quality, (size_t)rows * cols, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    for (int r = 0; r < rows; ++r)
        for (int c = 0; c < cols; ++c) q[(size_t)r * row_stride + c] = ctx->h_q8[(size_t)r * cols + c];
    return 0;
}

int fishabm_sum_quality(fishabm_ctx* ctx, int64_t* sums) {
    if (!ctx || !sums) return fail(ctx, FISHABM_E_ARG, "fishabm_sum_quality: null argument");
    CK(cudaSetDevice(ctx->p.device));
    { int rcf = flush_pending(ctx); if (rcf) return rcf; }
    CK(cudaMemsetAsync(ctx->stats + 4, 0, sizeof(unsigned long long), ctx->stream));
    sum_quality_kernel<<<296, 256, 0, ctx->stream>>>(ctx->quality, (size_t)ctx->qrows * ctx->qcols, ctx->stats + 4);
    ctx->launches += 1;
    unsigned long long v = 0;
    CK(cudaMemcpyAsync(&v, ctx->stats + 4, sizeof v, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    sums[0] = (int64_t)v;
    return 0;
}

int fishabm_load_quality_csv(const char* path, int32_t* q, int32_t max_rows, int32_t max_cols, int32_t row_stride,
                             int32_t* rows_out, int32_t* cols_out) {
    // get_environment's parser (fish_mod.cpp:604-629): lines of comma separated ints.  Unlike the reference
    // (which silently yields an undefined grid) a missing file is an error.
    if (!path || !q || !rows_out || !cols_out) return FISHABM_E_ARG;
    FILE* f = fopen(path, "r");
    if (!f) return FISHABM_E_ARG;
    int rows = 0, cols = 0;
    std::string line;
    int ch;
    bool fail_fmt = false;
    auto flush = [&]() {
        if (line.empty()) return;
        if (rows >= max_rows) { fail_fmt = true; return; }
        int c = 0;
        const char* s = line.c_str();
        while (*s) {
            char* endp = nullptr;
            long v = strtol(s, &endp, 10);
            if (endp == s) { fail_fmt = true; return; }
            if (c >= max_cols) { fail_fmt = true; return; }
            q[(size_t)rows * row_stride + c++] = (int32_t)v;
            s = endp;
            while (*s == ',' || *s == ' ' || *s == '\r') ++s;
        }
        if (c > cols) cols = c;
        rows++;
        line.clear();
    };
    while ((ch = fgetc(f)) != EOF) { if (ch == '\n') flush(); else line.push_back((char)ch); }
    flush();
    fclose(f);
    if (fail_fmt) return FISHABM_E_ARG;
    *rows_out = rows; *cols_out = cols;
    return 0;
}

// ---- frozen-state queries ------------------------------------------------------------------------------
static int export_pass(fishabm_ctx* ctx, int32_t* cnt4, double* turn, int32_t* n_avoid, int32_t* n_align, int32_t* n_attract, double* speed) {
    const size_t n = (size_t)ctx->n;
    cudaStream_t s = ctx->stream;
    int* d4 = ctx->st_i4;
    ExportPassArgs a{ctx->buf[ctx->cur].cold1, ctx->cnt, ctx->turn, ctx->aux,
                     cnt4 ? d4 : nullptr, turn ? ctx->st_dir : nullptr, n_avoid ? ctx->st_lag : nullptr,
                     n_align ? ctx->st_rate : nullptr, n_attract ? ctx->st_intake : nullptr, speed ? ctx->st_sf : nullptr,
                     ctx->n, 0, ctx->speed_norm};
    export_pass_kernel<<<cdiv(ctx->n, 256), 256, 0, s>>>(a);
    ctx->launches += 1;
    CK(cudaGetLastError());
    if (cnt4) CK(cudaMemcpyAsync(cnt4, d4, 4 * n * sizeof(int), cudaMemcpyDeviceToHost, s));
    if (turn) CK(cudaMemcpyAsync(turn, ctx->st_dir, n * sizeof(double), cudaMemcpyDeviceToHost, s));
    if (n_avoid) CK(cudaMemcpyAsync(n_avoid, ctx->st_lag, n * sizeof(int), cudaMemcpyDeviceToHost, s));
    if (n_align) CK(cudaMemcpyAsync(n_align, ctx->st_rate, n * sizeof(int), cudaMemcpyDeviceToHost, s));
    if (n_attract) CK(cudaMemcpyAsync(n_attract, ctx->st_intake, n * sizeof(int), cudaMemcpyDeviceToHost, s));
    if (speed) CK(cudaMemcpyAsync(speed, ctx->st_sf, n * sizeof(double), cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    return 0;
}

int fishabm_zone_counts(fishabm_ctx* ctx, int32_t* out4) {
    int rc = need_state(ctx);
    if (rc) return rc;
    if (!out4) return fail(ctx, FISHABM_E_ARG, "null argument");
    CK(cudaSetDevice(ctx->p.device));
    { int rcf = flush_pending(ctx); if (rcf) return rcf; }
    if ((rc = neighbour_pass(ctx, -1, false, ctx->step, ctx->step, false))) return rc;
    return export_pass(ctx, out4, nullptr, nullptr, nullptr, nullptr, nullptr);
}

int fishabm_social(fishabm_ctx* ctx, double* turn, int32_t* n_avoid, int32_t* n_align, int32_t* n_attract) {
    int rc = need_state(ctx);
    if (rc) return rc;
    if (!turn) return fail(ctx, FISHABM_E_ARG, "null argument");
    CK(cudaSetDevice(ctx->p.device));
    { int rcf = flush_pending(ctx); if (rcf) return rcf; }
    if ((rc = neighbour_pass(ctx, -1, false, ctx->step, ctx->step, false))) return rc;
    return export_pass(ctx, nullptr, turn, n_avoid, n_align, n_attract, nullptr);
}

int fishabm_get_speed(fishabm_ctx* ctx, double* out) {
    int rc = need_state(ctx);
    if (rc) return rc;
    if (!out) return fail(ctx, FISHABM_E_ARG, "null argument");
    CK(cudaSetDevice(ctx->p.device));
    { int rcf = flush_pending(ctx); if (rcf) return rcf; }
    if ((rc = neighbour_pass(ctx, -1, false, ctx->step, ctx->step, false))) return rc;
    return export_pass(ctx, nullptr, nullptr, nullptr, nullptr, nullptr, out);
}

int fishabm_in_zone(fishabm_ctx* ctx, int32_t fov_deg, double r, int32_t* out) {
    int rc = need_state(ctx);
    if (rc) return rc;
    if (!out) return fail(ctx, FISHABM_E_ARG, "null argument");
    if (!(r <= 200.0) || fov_deg < 0 || fov_deg > 360) return fail(ctx, FISHABM_E_ARG, "fishabm_in_zone: needs r <= 200 and 0 <= fov <= 360");
    CK(cudaSetDevice(ctx->p.device));
    { int rcf = flush_pending(ctx); if (rcf) return rcf; }
    const Buffers& b = ctx->buf[ctx->cur];
    ZoneArgs a{b.rec, b.cold0, b.cold1, b.cell, ctx->cell_start, ctx->st_lag, ctx->n, ctx->nx, ctx->ny, 0, fov_deg / 2, r};
    in_zone_generic_kernel<<<cdiv(ctx->n, 128), 128, 0, ctx->stream>>>(a);
    ctx->launches += 1;
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(out, ctx->st_lag, (size_t)ctx->n * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return 0;
}

// ---- stepping --------------------------------------------------------------------------------------------
int fishabm_move(fishabm_ctx* ctx) {
    int rc = need_state(ctx);
    if (rc) return rc;
    CK(cudaSetDevice(ctx->p.device));
    if ((rc = begin_timed(ctx))) return rc;
    if ((rc = flush_pending(ctx, true))) return rc;
    if ((rc = neighbour_pass(ctx, 0, false, ctx->step, ctx->step, true))) return rc;
    if ((rc = update_pass(ctx, false, true, ctx->step))) return rc;
    return end_timed(ctx);
}

int fishabm_sample(fishabm_ctx* ctx) {
    int rc = need_state(ctx);
    if (rc) return rc;
    CK(cudaSetDevice(ctx->p.device));
    if ((rc = begin_timed(ctx))) return rc;
    if (ctx->pending_sample) {  // the step's sample() is the pending one
        if ((rc = flush_pending(ctx, true))) return rc;
    } else {
        if ((rc = neighbour_pass(ctx, 0, true, ctx->step, ctx->step, true))) return rc;
        if ((rc = update_pass(ctx, true, false, ctx->step))) return rc;
        ctx->step += 1;
    }
    return end_timed(ctx);
}

static int step_untimed(fishabm_ctx* ctx, int nsteps) {
    // move(s); { sample(s+t-1) + move(s+t) } ...; the trailing sample stays pending (flush_pending)
    int rc;
    for (int t = 0; t < nsteps; ++t) {
        const uint32_t s = ctx->step;
        if (ctx->pending_sample) {
            if ((rc = neighbour_pass(ctx, 1, true, s - 1, s, true))) return rc;
            if ((rc = update_pass(ctx, true, true, s))) return rc;
        } else {
            if ((rc = neighbour_pass(ctx, 0, false, s, s, true))) return rc;
            if ((rc = update_pass(ctx, false, true, s))) return rc;
        }
        ctx->pending_sample = true;
        ctx->step = s + 1;
    }
    return 0;
}

int fishabm_step(fishabm_ctx* ctx, int32_t nsteps) {
    int rc = need_state(ctx);
    if (rc) return rc;
    if (nsteps < 1) return fail(ctx, FISHABM_E_ARG, "nsteps must be >= 1");
    CK(cudaSetDevice(ctx->p.device));
    if ((rc = begin_timed(ctx))) return rc;
    if ((rc = step_untimed(ctx, nsteps))) return rc;
    return end_timed(ctx);
}

int fishabm_flush(fishabm_ctx* ctx) {
    int rc = need_state(ctx);
    if (rc) return rc;
    CK(cudaSetDevice(ctx->p.device));
    if ((rc = begin_timed(ctx))) return rc;
    if ((rc = flush_pending(ctx, true))) return rc;
    return end_timed(ctx);
}

int fishabm_run_logged(fishabm_ctx* ctx, int32_t nsteps, int64_t* sumq, int32_t* intake, int32_t intake_every) {
    int rc = need_state(ctx);
    if (rc) return rc;
    if (nsteps < 1) return fail(ctx, FISHABM_E_ARG, "nsteps must be >= 1");
    CK(cudaSetDevice(ctx->p.device));
    int rows = 0;
    for (int z = 0; z < nsteps; ++z) {
        ctx->nb_events_used = 0;
        if ((rc = step_untimed(ctx, 1))) return rc;
        if (intake && intake_every > 0 && z % intake_every == 0) {
            if ((rc = fishabm_get_state(ctx, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, intake + (size_t)rows * ctx->n))) return rc;
            rows++;
        }
        if (sumq && (rc = fishabm_sum_quality(ctx, sumq + z))) return rc;
    }
    CK(cudaStreamSynchronize(ctx->stream));
    return 0;
}

int fishabm_feed(fishabm_ctx* ctx, const int32_t* ids, int32_t k) {
    int rc = need_state(ctx);
    if (rc) return rc;
    if (!ids || k < 0) return fail(ctx, FISHABM_E_ARG, "null argument");
    if (k > ctx->n) return fail(ctx, FISHABM_E_ARG, "fishabm_feed: at most n ids per call");
    if (k == 0) return 0;
    CK(cudaSetDevice(ctx->p.device));
    { int rcf = flush_pending(ctx); if (rcf) return rcf; }
    const Buffers& b = ctx->buf[ctx->cur];
    CK(cudaMemcpyAsync(ctx->st_lag, ids, (size_t)k * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
    index_by_id_kernel<<<cdiv(ctx->n, 256), 256, 0, ctx->stream>>>(b.cold1, ctx->n, 0, ctx->pos_of_id);
    FeedArgs a{b.cold0, b.cold1, b.rec, b.cell, ctx->st_lag, k, ctx->quality, ctx->n, ctx->nx, ctx->qcols, 0, ctx->pos_of_id};
    feed_list_kernel<<<1, 32, 0, ctx->stream>>>(a);
    ctx->launches += 2;
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(ctx->stream));
    return 0;
}

int fishabm_get_step(const fishabm_ctx* ctx, uint32_t* step) {
    if (!ctx || !step) return FISHABM_E_ARG;
    *step = ctx->step;
    return 0;
}
int fishabm_set_step(fishabm_ctx* ctx, uint32_t step) {
    if (!ctx) return FISHABM_E_ARG;
    ctx->step = step;
    return 0;
}

int fishabm_draw(const fishabm_ctx* ctx, int32_t stream, uint32_t replicate, uint32_t step, uint32_t id, uint32_t k, int32_t a,
                 int32_t b, int32_t* out) {
    if (!ctx || !out || b < a) return FISHABM_E_ARG;
    *out = fishrng::draw(ctx->p.seed, replicate, step, id, (uint32_t)stream, k, a, b);
    return 0;
}

int fishabm_last_step_ms(const fishabm_ctx* ctx, float* ms) {
    if (!ctx || !ms) return FISHABM_E_ARG;
    *ms = ctx->last_ms;
    return 0;
}

int fishabm_last_neighbour_ms(const fishabm_ctx* ctx, float* ms, int32_t* launches) {
    if (!ctx || !ms) return FISHABM_E_ARG;
    float total = 0.f;
    for (int k = 0; k < ctx->nb_events_used; ++k) {
        float t = 0.f;
        if (cudaEventElapsedTime(&t, ctx->nb_events[k].first, ctx->nb_events[k].second) == cudaSuccess) total += t;
    }
    *ms = total;
    if (launches) *launches = ctx->nb_events_used;
    return 0;
}

int fishabm_enable_stats(fishabm_ctx* ctx, int32_t on) {
    if (!ctx) return FISHABM_E_ARG;
    ctx->stats_on = on != 0;
    return 0;
}

int fishabm_last_pass_stats(fishabm_ctx* ctx, fishabm_pass_stats* out) {
    if (!ctx || !out) return FISHABM_E_ARG;
    unsigned long long v[4] = {0, 0, 0, 0};
    CK(cudaSetDevice(ctx->p.device));
    CK(cudaMemcpyAsync(v, ctx->stats, sizeof v, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    out->focal = v[0]; out->candidates = v[1]; out->interacting = v[2]; out->rechecked = v[3];
    return 0;
}

int fishabm_launch_count(const fishabm_ctx* ctx, uint64_t* out) {
    if (!ctx || !out) return FISHABM_E_ARG;
    *out = ctx->launches;
    return 0;
}

// ---- host-side scalar helpers (reference semantics; utilities, not the hot path) ----------------------------
double fishabm_correctangle(double dir) {  // fish_mod.cpp:410-418
    if (dir < 0) return dir + 360;
    if (dir >= 360) return dir - 360;
    return dir;
}

void fishabm_correct_coords(double* xy, int32_t w, int32_t h) {  // fish_mod.cpp:466-480
    if (!xy) return;
    if (xy[0] > w) xy[0] -= w; else if (xy[0] < 0) xy[0] += w;
    if (xy[1] > h) xy[1] -= h; else if (xy[1] < 0) xy[1] += h;
}

double fishabm_averageturn(const double* v, int32_t n, int32_t comb_weight) {  // fish_mod.cpp:379-408
    if (!v || n < 1) return 0.0;
    double xs = 0, ys = 0;
    for (int i = 0; i < n; ++i) {
        const double wgt = comb_weight ? (1.7 - i) : 1.0;
        xs += cos(v[i] * M_PI / 180) * wgt;
        ys += sin(v[i] * M_PI / 180) * wgt;
    }
    const double x = xs / n, y = ys / n;
    if (y == 0) return x >= 0 ? 0.0 : 180.0;
    const double ang = acos(x / sqrt(x * x + y * y)) * 180 / M_PI;
    return y < 0 ? -ang : ang;
}

}  // extern "C"
