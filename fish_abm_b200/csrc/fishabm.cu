// libfishabm.so: C ABI (include/fishabm.h) over the sm_100a kernels in kernels.cuh.
// No CPU fallback: every operation on fish state runs on the device; without a device fishabm_create fails.
#include "../../include/fishabm.h"

#include <cuda_runtime.h>
#include <dlfcn.h>
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>

#include <algorithm>
#include <string>
#include <vector>

#include <nvtx3/nvToolsExt.h>   // header-only NVTX 3: ranges cost nothing unless a profiler injects its library

#include "kernels.cuh"
#include "neighbour_v3.cuh"
#include "neighbour_v4.cuh"
#include "ensemble.cuh"

using namespace fishk;

namespace {

thread_local std::string g_create_error;

// NVTX range around each phase of a step: K4 neighbour pass, K5 update (+ message packing), the slab
// exchange, K1-K3 ingest + counting sort; visible in Nsight Systems / Compute timelines
struct NvtxRange {
    explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
    ~NvtxRange() { nvtxRangePop(); }
};

struct Buffers {  // one set of cell-sorted arrays
    float4* rec = nullptr;
    float4* cold0 = nullptr;
    int4* cold1 = nullptr;
    int* cell = nullptr;
};

}  // namespace

struct fishabm_ctx {
    fishabm_params p;
    std::string err_msg;
    cudaStream_t stream = nullptr;
    cudaStream_t copy_stream = nullptr;   // get_state: downloads that overlap the pending sample() pass
    cudaEvent_t ev_copy = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> nb_events;  // around each neighbour kernel of the last call
    int nb_events_used = 0;
    // optional phase timing of fishabm_step (fishabm_enable_phase_timing): 5 events per step
    bool phases_on = false;
    std::vector<cudaEvent_t> ph_events;
    int ph_steps = 0;
    int n = 0;            // fish ids are 0..n-1 (the whole school, also in slab mode)
    int cap = 0;          // entries the sorted arrays can hold (= n for a whole school)
    Grid g;               // local grid (kernels.cuh)
    int ncell = 0, qrows = 0, qcols = 0;   // local cells; food grid held here: qrows x qcols (owned columns)
    int qcols_g = 0, qcol0 = 0;            // global food columns; first owned food column
    bool ens = false;     // replicate ensemble (ensemble.cuh): R schools of n fish, arrays hold R*n entries
    int R = 1;
    size_t ens_smem = 0;
    long long* d_sums = nullptr;   // [R]
    bool slab = false;    // halo columns + message path
    int own = 0, col0 = 0, left = 0, right = 0;
    int cap_m = 0, cap_g = 0;
    size_t msg_bytes = 0;
    unsigned char *out_left = nullptr, *out_right = nullptr, *in_left = nullptr, *in_right = nullptr;
    void* comm = nullptr;  // ncclComm_t once connected
    // peer-memory transport: my inbox [2 sides][2 slots] + flags; the neighbours' inboxes once attached
    unsigned char* inbox = nullptr;
    size_t msg_stride = 0, inbox_bytes = 0;
    unsigned char *peer_left = nullptr, *peer_right = nullptr;
    void* ipc_opened[2] = {nullptr, nullptr};
    bool p2p = false, published = false;
    int seq = 0;                 // exchanges completed (the same on every rank)
    SlabHeader* ctr = nullptr;   // [2] local slot counters while packing into remote inboxes
    long long timeout_ns = 10000000000LL;
    bool step_open = false;  // between fishabm_slab_begin_step and fishabm_slab_end_step
    int* n_unsorted = nullptr;
    int* err = nullptr;
    int h_n = 0;
    std::vector<int> h_ids;
    std::vector<char> h_tmp;
    double speed_norm = 1.0;
    Model model;                 // the context's model constants in device form
    Model* d_models = nullptr;   // ensembles: one per replicate
    uint32_t step = 0;
    bool have_state = false;
    bool pending_sample = false;  // sample() of step `step-1` not yet applied (it shares the next move's neighbour pass)
    bool stats_on = false;
    int nb_variant = 4;  // neighbour kernel design: 4 = candidate-parallel, one window per vertical cell pair (default);
                         // 3 = candidate-parallel, one window per cell (FISHABM_NEIGHBOUR=v3); 1 = focal-parallel (FISHABM_NEIGHBOUR=v1)
    uint64_t launches = 0;
    float last_ms = 0.f;

    Buffers buf[2];
    int cur = 0;
    float4* rec_tmp = nullptr;   // moved records of a fused neighbour + update pass (kernels read buf[cur].rec meanwhile)
    int apply_up = -1;           // >= 0: the deferred food-grid writes of upd_count[apply_up] ride in the coming scatter
    float4* rec_src = nullptr;   // where the coming sort finds the unsorted records: rec_tmp after a fused pass, else buf[cur].rec
    bool fuse_update = false;    // FISHABM_FUSED_UPDATE=1: K5 inside the neighbour kernel (measured SLOWER: 363 vs 249 + 45 us at C3,
                                 // DESIGN.md section 3; kept as a bit-identical cross-check of the update's data flow)
    int *next_cell = nullptr, *next_rank = nullptr, *cell_count = nullptr, *cell_start = nullptr, *tile_sums = nullptr;
    int4* cnt = nullptr;
    float* turn = nullptr;
    uint32_t* aux = nullptr;
    uint8_t* quality = nullptr;
    int *upd_count = nullptr, *upd_idx = nullptr;  // upd_count[2]: alternate by sample(), each apply clears the other
    int upd_parity = 0, wc_parity = 0;
    int* scan_sync = nullptr;   // [2] ticket + epoch flag of the fused scan, [1] ticket of the chained scan
    unsigned long long* scan_desc = nullptr;   // chained scan: one published word per tile
    bool scan_chained = true;   // FISHABM_SCAN=coop: the cooperative one-launch scan instead (A/B)
    int scan_epoch = 0, scan_resident = 0;
    uint8_t* upd_val = nullptr;
    unsigned long long* stats = nullptr;  // 4 pass counters + 1 scratch for sums
    int* work_counter = nullptr;
    int nb_grid = 0;
    int* bad = nullptr;
    int* pos_of_id = nullptr;
    // host<->device staging in the reference's layout
    double *st_coords = nullptr, *st_dir = nullptr, *st_speed = nullptr, *st_sf = nullptr;
    int *st_lag = nullptr, *st_rate = nullptr, *st_intake = nullptr, *st_i4 = nullptr, *st_ids = nullptr;
    uint8_t* h_q8 = nullptr;
};

namespace {

int fail(fishabm_ctx* c, int code, const char* fmt, ...) {
    char b[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(b, sizeof b, fmt, ap);
    va_end(ap);
    if (c) c->err_msg = b; else g_create_error = b;
    return code;
}

#define CK(call)                                                                                          \
    do {                                                                                                  \
        cudaError_t e_ = (call);                                                                          \
        if (e_ != cudaSuccess) return fail(ctx, FISHABM_E_CUDA, "%s: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
    } while (0)

template <class T> cudaError_t dalloc(T** p, size_t count) { return cudaMalloc((void**)p, count * sizeof(T) > 0 ? count * sizeof(T) : 1); }

inline int cdiv(int a, int b) { return (a + b - 1) / b; }

void model_defaults(fishabm_model_params* m) {
    memset(m, 0, sizeof *m);
    m->lag_after_feed = 10; m->sample_rate_fed = 30; m->sample_rate_unfed = 0; m->sample_draw_max = 35;
    m->error_range = 10; m->randomwalk_range = 45; m->alone_align_min = 5;
    m->fed_speed_factor = 0.5; m->weight_align = 1.7; m->weight_attract = 1.7 - 1;   // (1.7 - i), fish_mod.cpp:385-386
}
const char* model_invalid(const fishabm_model_params& m) {
    if (m.lag_after_feed < 0 || m.lag_after_feed > (1 << 20)) return "lag_after_feed must be in [0, 2^20]";
    if (m.sample_rate_fed < 0 || m.sample_rate_unfed < 0 || m.sample_rate_fed > (1 << 20) || m.sample_rate_unfed > (1 << 20)) return "sample rates must be in [0, 2^20]";
    if (m.sample_draw_max < 0 || m.sample_draw_max > (1 << 20)) return "sample_draw_max must be in [0, 2^20]";
    if (m.error_range < 0 || m.error_range > 180 || m.randomwalk_range < 0 || m.randomwalk_range > 180) return "error_range / randomwalk_range must be in [0, 180] degrees";
    if (m.alone_align_min < 0) return "alone_align_min must be >= 0";
    if (!(m.fed_speed_factor >= 0.0 && m.fed_speed_factor <= 3.0)) return "fed_speed_factor must be in [0, 3]";
    if (!(m.weight_align >= 0.0 && m.weight_attract >= 0.0 && m.weight_align + m.weight_attract > 0.0)) return "combination weights must be >= 0 and not both zero";
    return nullptr;
}
Model to_device_model(const fishabm_model_params& m) {
    Model d;
    d.lag_after_feed = m.lag_after_feed; d.rate_fed = m.sample_rate_fed; d.rate_unfed = m.sample_rate_unfed;
    d.sample_draw_max = m.sample_draw_max; d.error_range = m.error_range; d.randomwalk_range = m.randomwalk_range;
    d.alone_align_min = m.alone_align_min;
    { const uint32_t range = 2u * (uint32_t)m.error_range + 1u; d.error_reject = (0u - range) % range; }
    d.fed_sf = (float)m.fed_speed_factor; d.w_align = (float)m.weight_align; d.w_attract = (float)m.weight_attract; d.pad2 = 0.f;
    return d;
}

// ---- NCCL, loaded on demand (only slab contexts that call fishabm_slab_connect need it).  The handful of
// declarations below restate the public NCCL 2.x C API so that building the library does not need nccl.h.
struct NcclUniqueId { char internal[128]; };
struct NcclApi {
    void* h = nullptr;
    int (*GetUniqueId)(NcclUniqueId*) = nullptr;
    int (*CommInitRank)(void**, int, NcclUniqueId, int) = nullptr;
    int (*CommDestroy)(void*) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    int (*Send)(const void*, size_t, int, int, void*, cudaStream_t) = nullptr;
    int (*Recv)(void*, size_t, int, int, void*, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
    std::string why;
};
constexpr int NCCL_UINT8 = 1;  // ncclUint8

NcclApi& nccl() {
    static NcclApi api;
    if (api.h || !api.why.empty()) return api;
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char* nm : names) { api.h = dlopen(nm, RTLD_NOW | RTLD_GLOBAL); if (api.h) break; }
    if (!api.h) { api.why = std::string("cannot load libnccl.so.2: ") + dlerror(); return api; }
    auto sym = [&](const char* n) { void* f = dlsym(api.h, n); if (!f && api.why.empty()) api.why = std::string("libnccl lacks ") + n; return f; };
    api.GetUniqueId = (int (*)(NcclUniqueId*))sym("ncclGetUniqueId");
    api.CommInitRank = (int (*)(void**, int, NcclUniqueId, int))sym("ncclCommInitRank");
    api.CommDestroy = (int (*)(void*))sym("ncclCommDestroy");
    api.GroupStart = (int (*)())sym("ncclGroupStart");
    api.GroupEnd = (int (*)())sym("ncclGroupEnd");
    api.Send = (int (*)(const void*, size_t, int, int, void*, cudaStream_t))sym("ncclSend");
    api.Recv = (int (*)(void*, size_t, int, int, void*, cudaStream_t))sym("ncclRecv");
    api.GetErrorString = (const char* (*)(int))sym("ncclGetErrorString");
    if (!api.why.empty()) { dlclose(api.h); api.h = nullptr; }
    return api;
}

#define NK(call)                                                                                                   \
    do {                                                                                                           \
        int r_ = (call);                                                                                           \
        if (r_ != 0) return fail(ctx, FISHABM_E_COMM, "%s: %s", #call, nccl().GetErrorString ? nccl().GetErrorString(r_) : "nccl error"); \
    } while (0)

// CUDA loads kernels lazily, and loading one may have to wait for the device to go idle.  The peer-memory transport
// keeps a kernel spinning on the device until a NEIGHBOUR's kernels have run, so a first-use load in between would
// deadlock (until the spin's timeout).  Load every kernel of the library up front, once per device.
cudaError_t preload_kernels() {
    cudaFuncAttributes fa;
    cudaError_t e = cudaSuccess;
    auto L = [&](const void* f) { if (e == cudaSuccess) e = cudaFuncGetAttributes(&fa, f); };
    L((const void*)apply_quality_updates_kernel); L((const void*)copy_int_kernel); L((const void*)ens_broadcast_quality_kernel);
    L((const void*)ens_export_kernel); L((const void*)ens_export_pass_kernel); L((const void*)ens_ingest_kernel);
    L((const void*)ens_sum_quality_kernel); L((const void*)ensemble_kernel); L((const void*)export_kernel);
    L((const void*)export_pass_kernel); L((const void*)feed_list_kernel); L((const void*)halo_refresh_kernel);
    L((const void*)in_zone_generic_kernel); L((const void*)index_by_id_kernel); L((const void*)ingest_incoming_kernel);
    L((const void*)ingest_kernel); L((const void*)neighbour_kernel<false>); L((const void*)neighbour_kernel<true>);
    L((const void*)neighbour_kernel_v4<false>); L((const void*)neighbour_kernel_v4<true>); L((const void*)scan_finish_kernel);
    L((const void*)neighbour_kernel_v3<false, false>); L((const void*)neighbour_kernel_v3<true, false>);
    L((const void*)neighbour_kernel_v3<false, true>); L((const void*)neighbour_kernel_v3<true, true>);
    L((const void*)scan_chained_kernel); L((const void*)scan_fused_kernel); L((const void*)scan_tile_offsets_kernel); L((const void*)scan_tile_sums_kernel);
    L((const void*)scatter_kernel); L((const void*)slab_publish_kernel); L((const void*)slab_wait_kernel);
    L((const void*)sum_quality_kernel); L((const void*)update_kernel); L((const void*)check_fp64_kernel); L((const void*)brute_fp64_pass_kernel);
    return e;
}

// a non-zero identifier of the physical GPU behind a device ordinal (the ordinal itself depends on CUDA_VISIBLE_DEVICES)
int32_t gpu_identity(int device) {
    char bus[64] = {0};
    if (cudaDeviceGetPCIBusId(bus, sizeof bus, device) != cudaSuccess) { cudaGetLastError(); return 1 + device; }
    uint32_t h = 2166136261u;
    for (const char* p = bus; *p; ++p) h = (h ^ (uint32_t)(unsigned char)*p) * 16777619u;
    return (int32_t)(h | 1u);
}

SlabMsg msg(const fishabm_ctx* c, unsigned char* base, SlabHeader* counters = nullptr) { return SlabMsg{base, c->cap_m, c->cap_g, counters}; }
// inbox layout: side 0 = messages from my left neighbour, side 1 = from my right; two slots per side (alternating
// exchanges: a neighbour may already be packing exchange s+1 while exchange s is still being ingested here)
size_t inbox_msg_off(const fishabm_ctx* c, int side, int slot) { return (size_t)(side * 2 + slot) * c->msg_stride; }
size_t inbox_flag_off(const fishabm_ctx* c, int side, int slot) { return 4 * c->msg_stride + (size_t)(side * 2 + slot) * 128; }

// ---- the counting sort: counts (already accumulated in cell_count) -> cell_start; scatter cur -> other
int rebuild(fishabm_ctx* ctx) {
    NvtxRange nv("fishabm: K2 scan + K3 scatter (counting sort)");
    const int ntiles = cdiv(ctx->ncell, SCAN_TILE);
    bool fused_scan = false;
    if (ctx->scan_chained) {
        ctx->scan_epoch = (ctx->scan_epoch + 1) & 0x3FFFFFFF;
        if (ctx->scan_epoch == 0) ctx->scan_epoch = 1;   // (0 is what the words are cleared to)
        ScanChainArgs f{ctx->cell_count, ctx->ncell, ctx->cell_start, ctx->scan_desc, ctx->scan_sync + 2, (unsigned)ctx->scan_epoch, ntiles};
        scan_chained_kernel<<<ntiles, SCAN_THREADS, 0, ctx->stream>>>(f);
        fused_scan = true;
        ctx->launches -= 2;
    } else if (ntiles <= ctx->scan_resident) {
        // one launch: its blocks wait for the block that scans the tile sums, so they must all be resident at the same time.
        // A COOPERATIVE launch makes the driver guarantee that (whatever else shares the GPU) or refuse the launch.
        ScanFusedArgs f{ctx->cell_count, ctx->ncell, ctx->tile_sums, ctx->cell_start, ctx->scan_sync, ctx->scan_sync + 1, ctx->scan_epoch + 1, ntiles};
        void* params[] = {&f};
        if (cudaLaunchCooperativeKernel((const void*)scan_fused_kernel, dim3(ntiles), dim3(SCAN_THREADS), params, 0, ctx->stream) == cudaSuccess) {
            ctx->scan_epoch += 1;
            fused_scan = true;
            ctx->launches -= 2;
        } else {
            cudaGetLastError();          // refused (not enough free SMs for a co-resident grid): the three-launch scan below
            ctx->scan_resident = 0;
        }
    }
    if (!fused_scan) {
        scan_tile_sums_kernel<<<ntiles, SCAN_THREADS, 0, ctx->stream>>>(ctx->cell_count, ctx->ncell, ctx->tile_sums);
        scan_tile_offsets_kernel<<<1, SCAN_THREADS, 0, ctx->stream>>>(ctx->tile_sums, ntiles);
        scan_finish_kernel<<<ntiles, SCAN_THREADS, 0, ctx->stream>>>(ctx->cell_count, ctx->ncell, ctx->tile_sums, ctx->cell_start);
    }
    const Buffers& in = ctx->buf[ctx->cur];
    const Buffers& out = ctx->buf[ctx->cur ^ 1];
    ScatterArgs s{ctx->rec_src ? ctx->rec_src : in.rec, in.cold0, in.cold1, out.rec, out.cold0, out.cold1, out.cell, ctx->next_cell, ctx->next_rank, ctx->cell_start, ctx->n_unsorted, ctx->cap,
                  nullptr, nullptr, nullptr, nullptr, nullptr};
    const int sgrid = cdiv(ctx->cap, 256);
    if (ctx->apply_up >= 0) {
        if (sgrid >= SCATTER_APPLY_BLOCKS) {
            s.quality = ctx->quality; s.upd_count = ctx->upd_count + ctx->apply_up; s.upd_count_next = ctx->upd_count + (ctx->apply_up ^ 1);
            s.upd_idx = ctx->upd_idx; s.upd_val = ctx->upd_val;
        } else {   // tiny school: the scatter has fewer blocks than the apply loop assumes
            apply_quality_updates_kernel<<<64, 256, 0, ctx->stream>>>(ctx->quality, ctx->upd_count + ctx->apply_up, ctx->upd_count + (ctx->apply_up ^ 1), ctx->upd_idx, ctx->upd_val);
            ctx->launches += 1;
        }
        ctx->apply_up = -1;
    }
    scatter_kernel<<<sgrid, 256, 0, ctx->stream>>>(s);
    ctx->launches += 4;
    ctx->cur ^= 1;
    ctx->rec_src = nullptr;
    CK(cudaGetLastError());
    return 0;
}

enum { PASS_COUNTS = 1, PASS_TURN = 2 };

// K4: one neighbour pass.  active_lag_max < 0 evaluates every fish (frozen-state queries).
int split_for(const fishabm_ctx* ctx) {
    // fewer cells than resident warps (small or dense schools): split every cell's fish over several work items
    const int own_cells = ctx->g.own_c1 - ctx->g.own_c0;
    int split = std::max(1, std::min(32, (2 * ctx->nb_grid * V4_WARPS) / std::max(1, own_cells)));
    if (const char* sp = getenv("FISHABM_SPLIT")) split = std::max(1, std::min(32, atoi(sp)));
    return split;
}
// the per-fish update can ride inside the neighbour kernel when a warp owns whole cells
bool can_fuse_update(const fishabm_ctx* ctx) {
    return ctx->fuse_update && ctx->nb_variant == 3 && ctx->p.mode == FISHABM_MODE_CELL_FP32 && !ctx->ens && split_for(ctx) == 1;
}

// fuse: the update (K5) to run inside the pass (see neighbour_kernel_v3<., true>), or null
int neighbour_pass(fishabm_ctx* ctx, int active_lag_max, bool do_sample, uint32_t step_sample, uint32_t step_move, bool timed,
                   const UpdateArgs* fuse = nullptr) {
    NvtxRange nv("fishabm: K4 neighbour pass");
    const Buffers& b = ctx->buf[ctx->cur];
    NeighArgs a;
    a.rec = b.rec; a.cold0 = b.cold0; a.cold1 = b.cold1; a.cell_start = ctx->cell_start;
    a.cnt = ctx->cnt; a.turn = ctx->turn; a.aux = ctx->aux; a.stats = ctx->stats_on ? ctx->stats : nullptr;
    a.g = ctx->g;
    a.active_lag_max = active_lag_max; a.do_sample = do_sample ? 1 : 0;
    a.step_sample = step_sample; a.step_move = step_move; a.seed = ctx->p.seed; a.replicate = ctx->p.replicate0;
    a.split = 1; a.single_cols = 0;
    a.model = ctx->model;
    a.work_counter = ctx->work_counter + ctx->wc_parity;            // the persistent kernel's two cell counters alternate:
    a.work_counter_next = ctx->work_counter + (ctx->wc_parity ^ 1);  // each launch clears the one the next launch uses
    if (ctx->stats_on) CK(cudaMemsetAsync(ctx->stats, 0, 8 * sizeof(unsigned long long), ctx->stream));
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    if (timed && ctx->nb_events_used >= 4096) timed = false;  // per-launch timing covers the first 4096 passes of a call
    if (timed) {
        if (ctx->nb_events_used == (int)ctx->nb_events.size()) {
            cudaEvent_t x, y;
            CK(cudaEventCreate(&x)); CK(cudaEventCreate(&y));
            try { ctx->nb_events.emplace_back(x, y); }
            catch (...) { cudaEventDestroy(x); cudaEventDestroy(y); return fail(ctx, FISHABM_E_NOMEM, "out of host memory"); }
        }
        e0 = ctx->nb_events[ctx->nb_events_used].first; e1 = ctx->nb_events[ctx->nb_events_used].second;
        ctx->nb_events_used++;
        CK(cudaEventRecord(e0, ctx->stream));
    }
    const int own_cells = ctx->g.own_c1 - ctx->g.own_c0;
    if (ctx->p.mode == FISHABM_MODE_BRUTE_FP64) {
        BrutePassArgs ba;
        ba.c = CheckArgs{b.rec, b.cold0, b.cold1, b.cell, nullptr, nullptr, 0, nullptr, nullptr, nullptr,
                         ctx->n, ctx->g, ctx->p.w, ctx->p.h, step_move, ctx->p.seed, ctx->p.replicate0, ctx->model,
                         ctx->p.model.weight_align, ctx->p.model.weight_attract};
        ba.cnt = ctx->cnt; ba.turn = ctx->turn; ba.aux = ctx->aux;
        ba.active_lag_max = active_lag_max; ba.do_sample = do_sample ? 1 : 0; ba.step_sample = step_sample;
        brute_fp64_pass_kernel<<<cdiv(ctx->n * 32, 256), 256, 0, ctx->stream>>>(ba);
    } else if (ctx->nb_variant == 1) {
        const int grid = cdiv(own_cells, NB_WARPS);
        if (ctx->stats_on) neighbour_kernel<true><<<grid, NB_THREADS, 0, ctx->stream>>>(a);
        else neighbour_kernel<false><<<grid, NB_THREADS, 0, ctx->stream>>>(a);
    } else {
        ctx->wc_parity ^= 1;
        a.split = split_for(ctx);
        const int grid = std::min(ctx->nb_grid, cdiv(own_cells * a.split, V4_WARPS));
        if (ctx->nb_variant == 4) {
            // work items are cell PAIRS when split == 1, except in the last columns: when the dispenser runs dry every warp
            // still finishes its item, so the final `tail` items' worth of work per resident warp is handed out cell by cell
            if (a.split == 1) {
                const char* pt = getenv("FISHABM_PAIR_TAIL");   // (tests force 0: every column as cell pairs, however small the school)
                const double tail = pt ? atof(pt) : 2.0;
                const int ncols = own_cells / ctx->g.ny;
                a.single_cols = std::min(ncols, (int)std::ceil(tail * ctx->nb_grid * V4_WARPS / (double)ctx->g.ny));
            }
            if (ctx->stats_on) neighbour_kernel_v4<true><<<grid, V4_THREADS, V4_SMEM_BYTES, ctx->stream>>>(a);
            else neighbour_kernel_v4<false><<<grid, V4_THREADS, V4_SMEM_BYTES, ctx->stream>>>(a);
        } else if (fuse) {
            if (ctx->stats_on) neighbour_kernel_v3<true, true><<<grid, V3_THREADS, V3_SMEM_BYTES, ctx->stream>>>(a, *fuse);
            else neighbour_kernel_v3<false, true><<<grid, V3_THREADS, V3_SMEM_BYTES, ctx->stream>>>(a, *fuse);
        } else {
            UpdateArgs none;
            memset(&none, 0, sizeof none);
            if (ctx->stats_on) neighbour_kernel_v3<true, false><<<grid, V3_THREADS, V3_SMEM_BYTES, ctx->stream>>>(a, none);
            else neighbour_kernel_v3<false, false><<<grid, V3_THREADS, V3_SMEM_BYTES, ctx->stream>>>(a, none);
        }
    }
    ctx->launches += 1;
    if (timed) CK(cudaEventRecord(e1, ctx->stream));
    CK(cudaGetLastError());
    return 0;
}

// the two messages of the coming exchange, emptied: local buffers (NCCL / caller-driven / lone slab) or, with the
// peer-memory transport, the neighbours' inboxes themselves (what leaves through my LEFT edge arrives from the RIGHT there)
int open_out_messages(fishabm_ctx* ctx, SlabMsg* out_left, SlabMsg* out_right) {
    if (ctx->p2p) {
        const int slot = (ctx->seq + 1) & 1;
        *out_left = msg(ctx, ctx->peer_left + inbox_msg_off(ctx, 1, slot), ctx->ctr + 0);
        *out_right = msg(ctx, ctx->peer_right + inbox_msg_off(ctx, 0, slot), ctx->ctr + 1);
        CK(cudaMemsetAsync(ctx->ctr, 0, 2 * sizeof(SlabHeader), ctx->stream));
    } else {
        *out_left = msg(ctx, ctx->out_left); *out_right = msg(ctx, ctx->out_right);
        CK(cudaMemsetAsync(ctx->out_left, 0, sizeof(SlabHeader), ctx->stream));
        CK(cudaMemsetAsync(ctx->out_right, 0, sizeof(SlabHeader), ctx->stream));
    }
    return 0;
}

// peer-memory transport: complete the two remote messages of exchange seq+1 (headers, then arrival flags)
int publish(fishabm_ctx* ctx) {
    const int s = ctx->seq + 1, slot = s & 1;
    SlabPublishArgs pa{ctx->ctr + 0, ctx->ctr + 1,
                       reinterpret_cast<SlabHeader*>(ctx->peer_left + inbox_msg_off(ctx, 1, slot)),
                       reinterpret_cast<SlabHeader*>(ctx->peer_right + inbox_msg_off(ctx, 0, slot)),
                       reinterpret_cast<int*>(ctx->peer_left + inbox_flag_off(ctx, 1, slot)),
                       reinterpret_cast<int*>(ctx->peer_right + inbox_flag_off(ctx, 0, slot)), s};
    slab_publish_kernel<<<1, 32, 0, ctx->stream>>>(pa);
    ctx->launches += 1;
    ctx->published = true;
    CK(cudaGetLastError());
    return 0;
}

// K5 (+ deferred food-grid writes); a move also packs the slab messages
int make_update_args(fishabm_ctx* ctx, bool do_sample, bool do_move, uint32_t step_move, UpdateArgs* out) {
    const Buffers& b = ctx->buf[ctx->cur];
    UpdateArgs& u = *out;
    u.rec = b.rec; u.rec_out = b.rec; u.cold0 = b.cold0; u.cold1 = b.cold1; u.cell = b.cell; u.cell_start = ctx->cell_start;
    u.cnt = ctx->cnt; u.turn = ctx->turn; u.aux = ctx->aux;
    u.next_cell = ctx->next_cell; u.next_rank = ctx->next_rank; u.next_count = ctx->cell_count;
    const int up = ctx->upd_parity;
    u.quality = ctx->quality; u.upd_count = ctx->upd_count + up; u.upd_idx = ctx->upd_idx; u.upd_val = ctx->upd_val;
    u.err = ctx->err;
    u.g = ctx->g; u.qcols = ctx->qcols;
    u.do_sample = do_sample; u.do_move = do_move; u.rolling = ctx->p.rolling;
    u.step_move = step_move; u.seed = ctx->p.seed; u.replicate = ctx->p.replicate0; u.speed_norm_inv2 = (float)(2.0 / ctx->speed_norm);
    u.slab = ctx->slab ? 1 : 0;
    u.model = ctx->model;
    u.out_left = msg(ctx, ctx->out_left); u.out_right = msg(ctx, ctx->out_right);
    if (ctx->slab && do_move) { int rc = open_out_messages(ctx, &u.out_left, &u.out_right); if (rc) return rc; }
    u.part = UPD_ALL;
    return 0;
}

// the deferred food-grid writes of a sample(): inside the coming sort's scatter when one follows (a move), else their own launch
int after_update(fishabm_ctx* ctx, bool do_sample, bool sort_follows) {
    if (do_sample) {
        const int up = ctx->upd_parity;
        if (sort_follows) ctx->apply_up = up;
        else {
            apply_quality_updates_kernel<<<64, 256, 0, ctx->stream>>>(ctx->quality, ctx->upd_count + up, ctx->upd_count + (up ^ 1), ctx->upd_idx, ctx->upd_val);
            ctx->launches += 1;
        }
        ctx->upd_parity ^= 1;
    }
    CK(cudaGetLastError());
    return 0;
}

int update_launch(fishabm_ctx* ctx, bool do_sample, bool do_move, uint32_t step_move) {
    NvtxRange nv("fishabm: K5 update (sample / feed / move, slab message packing)");
    UpdateArgs u;
    { int rc = make_update_args(ctx, do_sample, do_move, step_move, &u); if (rc) return rc; }
    static const bool edges_first = !(getenv("FISHABM_EDGES_FIRST") && atoi(getenv("FISHABM_EDGES_FIRST")) == 0);
    if (ctx->slab && do_move && ctx->p2p && ctx->p.nranks > 1 && edges_first) {
        // edges first: the two outermost owned columns per side hold every migrant and ghost; their messages are
        // published while the interior is still being updated (the neighbours' arrive behind that work)
        u.part = UPD_EDGES;
        update_kernel<<<std::max(1, std::min(cdiv(ctx->cap, 256), cdiv(8 * ctx->cap_g, 256))), 256, 0, ctx->stream>>>(u);
        int rc = publish(ctx);
        if (rc) return rc;
        u.part = UPD_MIDDLE;
        ctx->launches += 1;
    }
    update_kernel<<<cdiv(ctx->cap, 256), 256, 0, ctx->stream>>>(u);
    ctx->launches += 1;
    return after_update(ctx, do_sample, do_move);
}

// after a move: (slab) take in the two received messages, then re-sort
int finish_move(fishabm_ctx* ctx, unsigned char* in_left, unsigned char* in_right) {
    if (ctx->slab) {
        NvtxRange nv("fishabm: K1 ingest of received slab messages");
        const Buffers& b = ctx->buf[ctx->cur];
        IncomingArgs a{msg(ctx, in_left), msg(ctx, in_right), ctx->rec_src ? ctx->rec_src : b.rec, b.cold0, b.cold1, ctx->next_cell, ctx->next_rank, ctx->cell_count,
                       ctx->cell_start, ctx->n_unsorted, ctx->err, ctx->cap, ctx->g};
        ingest_incoming_kernel<<<std::max(1, std::min(1024, cdiv(2 * (ctx->cap_m + ctx->cap_g), 256))), 256, 0, ctx->stream>>>(a);
        ctx->launches += 1;
        CK(cudaGetLastError());
    }
    return rebuild(ctx);
}

// the exchange the library drives itself: a lone slab is its own neighbour on both sides; several slabs use NCCL
int exchange(fishabm_ctx* ctx, unsigned char** in_left, unsigned char** in_right) {
    NvtxRange nv("fishabm: slab exchange (halo columns + migrants)");
    if (ctx->p.nranks == 1) { *in_left = ctx->out_right; *in_right = ctx->out_left; return 0; }
    if (ctx->p2p) {
        if (!ctx->published) { int rc = publish(ctx); if (rc) return rc; }  // (move() published after its edge columns)
        ctx->published = false;
        const int s = ++ctx->seq, slot = s & 1;
        slab_wait_kernel<<<1, 32, 0, ctx->stream>>>(reinterpret_cast<int*>(ctx->inbox + inbox_flag_off(ctx, 0, slot)),
                                                    reinterpret_cast<int*>(ctx->inbox + inbox_flag_off(ctx, 1, slot)), s, ctx->timeout_ns, ctx->err);
        ctx->launches += 1;
        CK(cudaGetLastError());
        *in_left = ctx->inbox + inbox_msg_off(ctx, 0, slot);
        *in_right = ctx->inbox + inbox_msg_off(ctx, 1, slot);
        return 0;
    }
    if (!ctx->comm) return fail(ctx, FISHABM_E_STATE, "slab context is not connected: call fishabm_slab_attach or fishabm_slab_connect, or drive the exchange with fishabm_slab_begin_step / end_step");
    NcclApi& N = nccl();
    NK(N.GroupStart());
    NK(N.Send(ctx->out_left, ctx->msg_bytes, NCCL_UINT8, ctx->left, ctx->comm, ctx->stream));
    NK(N.Send(ctx->out_right, ctx->msg_bytes, NCCL_UINT8, ctx->right, ctx->comm, ctx->stream));
    // with two ranks both neighbours are the same peer: its first message (out_left) is what crosses MY right edge
    NK(N.Recv(ctx->in_right, ctx->msg_bytes, NCCL_UINT8, ctx->right, ctx->comm, ctx->stream));
    NK(N.Recv(ctx->in_left, ctx->msg_bytes, NCCL_UINT8, ctx->left, ctx->comm, ctx->stream));
    NK(N.GroupEnd());
    *in_left = ctx->in_left; *in_right = ctx->in_right;
    return 0;
}

int update_pass(fishabm_ctx* ctx, bool do_sample, bool do_move, uint32_t step_move) {
    int rc = update_launch(ctx, do_sample, do_move, step_move);
    if (rc || !do_move) return rc;
    unsigned char *il = nullptr, *ir = nullptr;
    if (ctx->slab && (rc = exchange(ctx, &il, &ir))) return rc;
    return finish_move(ctx, il, ir);
}

// phase boundaries of one step: 0 start, 1 after the neighbour pass, 2 after update (+publish), 3 after the exchange, 4 after the sort
constexpr int PH_PER_STEP = 5, PH_MAX_STEPS = 512;
int phase_mark(fishabm_ctx* ctx, int k) {
    if (!ctx->phases_on || ctx->ph_steps >= PH_MAX_STEPS) return 0;
    const size_t idx = (size_t)ctx->ph_steps * PH_PER_STEP + k;
    while (ctx->ph_events.size() <= idx) {
        cudaEvent_t e;
        CK(cudaEventCreate(&e));
        try { ctx->ph_events.push_back(e); } catch (...) { cudaEventDestroy(e); return fail(ctx, FISHABM_E_NOMEM, "out of host memory"); }
    }
    CK(cudaEventRecord(ctx->ph_events[idx], ctx->stream));
    if (k == PH_PER_STEP - 1) ctx->ph_steps++;
    return 0;
}

int begin_timed(fishabm_ctx* ctx) {
    ctx->nb_events_used = 0;
    ctx->ph_steps = 0;
    CK(cudaEventRecord(ctx->ev0, ctx->stream));
    return 0;
}
int check_device_errors(fishabm_ctx* ctx) {
    if (!ctx->slab) return 0;
    int e = 0, bad = 0;
    CK(cudaMemcpyAsync(&e, ctx->err, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaMemcpyAsync(&bad, ctx->bad, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    if (bad) {
        ctx->have_state = false;
        CK(cudaMemsetAsync(ctx->bad, 0, sizeof(int), ctx->stream));
        if (bad & 1) return fail(ctx, FISHABM_E_ARG, "set_state: coordinates must be finite and inside [0,w] x [0,h]");
        if (bad & 8) return fail(ctx, FISHABM_E_ARG, "fishabm_set_state_owned: an id is outside [0, %d)", ctx->n);
        if (bad & 16) return fail(ctx, FISHABM_E_ARG, "fishabm_set_state_owned: a speed is too large: slabs need 3*speed < 200");
        if (bad & 4) return fail(ctx, FISHABM_E_ARG, "fishabm_set_state_owned: a fish lies outside this slab's columns [%d, %d)", ctx->col0, ctx->col0 + ctx->own);
        return fail(ctx, FISHABM_E_NOMEM, "set_state: this slab holds more than its capacity of %d entries: raise fishabm_params.capacity", ctx->cap);
    }
    if (e & ERR_SLAB_OVERFLOW) return fail(ctx, FISHABM_E_NOMEM, "slab capacity exceeded (capacity %d, halo_capacity %d, migrant_capacity %d): raise fishabm_params.capacity / halo_capacity / migrant_capacity", ctx->cap, ctx->cap_g, ctx->cap_m);
    if (e & ERR_SLAB_TIMEOUT) return fail(ctx, FISHABM_E_COMM, "a neighbour slab's message did not arrive within %.1f s (peer-memory transport)", ctx->timeout_ns * 1e-9);
    if (e & ERR_STEP_TOO_LONG) return fail(ctx, FISHABM_E_ARG, "a fish moved more than 200 units in one step: not supported with slabs");
    return 0;
}
int end_timed(fishabm_ctx* ctx) {
    CK(cudaEventRecord(ctx->ev1, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    CK(cudaEventElapsedTime(&ctx->last_ms, ctx->ev0, ctx->ev1));
    return check_device_errors(ctx);
}

int need_state(fishabm_ctx* ctx) {
    if (!ctx) return FISHABM_E_ARG;
    if (!ctx->have_state) return fail(ctx, FISHABM_E_STATE, "no state: call fishabm_set_state or fishabm_initialize first");
    if (ctx->step_open) return fail(ctx, FISHABM_E_STATE, "a slab step is open: call fishabm_slab_end_step first");
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

// owned fish now (slab: read back from the device; whole school: n)
int owned_range(fishabm_ctx* ctx, int* count) {
    if (!ctx->slab) { *count = ctx->n; return 0; }
    int be[2] = {0, 0};
    CK(cudaMemcpyAsync(&be[0], ctx->cell_start + ctx->g.own_c0, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaMemcpyAsync(&be[1], ctx->cell_start + ctx->g.own_c1, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    *count = be[1] - be[0];
    return 0;
}

// device staging -> the caller's array.  Whole school: the staging array is already in id order.  Slab: the
// staging array is compact (owned fish in sorted order) and is scattered by id on the host.
template <class T>
int fetch(fishabm_ctx* ctx, T* user, const T* dev, int count, int width) {
    if (!user) return 0;
    if (!ctx->slab) {
        CK(cudaMemcpyAsync(user, dev, (size_t)count * width * sizeof(T), cudaMemcpyDeviceToHost, ctx->stream));
        return 0;
    }
    try { ctx->h_tmp.resize((size_t)count * width * sizeof(T) + 1); }   // nothing may throw across the C boundary
    catch (...) { return fail(ctx, FISHABM_E_NOMEM, "out of host memory"); }
    CK(cudaMemcpyAsync(ctx->h_tmp.data(), dev, (size_t)count * width * sizeof(T), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    const T* src = reinterpret_cast<const T*>(ctx->h_tmp.data());
    for (int k = 0; k < count; ++k)
        for (int c = 0; c < width; ++c) user[(size_t)ctx->h_ids[k] * width + c] = src[(size_t)k * width + c];
    return 0;
}
int fetch_ids(fishabm_ctx* ctx, int count) {
    if (!ctx->slab) return 0;
    try { ctx->h_ids.resize((size_t)count + 1); }
    catch (...) { return fail(ctx, FISHABM_E_NOMEM, "out of host memory"); }
    CK(cudaMemcpyAsync(ctx->h_ids.data(), ctx->st_ids, (size_t)count * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return 0;
}

// ---- replicate ensembles --------------------------------------------------------------------------------------------
EnsembleArgs ens_args(fishabm_ctx* c, int op, int nsteps) {
    EnsembleArgs a;
    memset(&a, 0, sizeof a);
    const Buffers& b = c->buf[0];
    a.rec = b.rec; a.cold0 = b.cold0; a.cold1 = b.cold1; a.quality = c->quality;
    a.n = c->n; a.nx = c->g.nx_g; a.ny = c->g.ny; a.qcols = c->qcols; a.qcells = c->qrows * c->qcols;
    a.replicates = c->R; a.replicate0 = c->p.replicate0;
    a.op = op; a.nsteps = nsteps; a.step0 = c->step;
    a.rolling = c->p.rolling; a.seed = c->p.seed; a.speed_norm = c->speed_norm;
    a.cnt = c->cnt; a.turn = c->turn; a.aux = c->aux;
    a.models = c->d_models;
    return a;
}
int ens_launch(fishabm_ctx* ctx, const EnsembleArgs& a) {
    ensemble_kernel<<<ctx->R, ENS_THREADS, ctx->ens_smem, ctx->stream>>>(a);
    ctx->launches += 1;
    CK(cudaGetLastError());
    return 0;
}
int ens_timed(fishabm_ctx* ctx, int op, int nsteps) {
    int rc;
    if ((rc = begin_timed(ctx))) return rc;
    if ((rc = ens_launch(ctx, ens_args(ctx, op, nsteps)))) return rc;
    return end_timed(ctx);
}
int ens_query(fishabm_ctx* ctx, int32_t* cnt4, double* turn, int32_t* n_avoid, int32_t* n_align, int32_t* n_attract, double* speed) {
    int rc;
    if ((rc = ens_launch(ctx, ens_args(ctx, ENS_OP_QUERY, 0)))) return rc;
    const long long tot = (long long)ctx->R * ctx->n;
    EnsPassOutArgs o{ctx->cnt, ctx->turn, tot, ctx->speed_norm, cnt4 ? ctx->st_i4 : nullptr, turn ? ctx->st_dir : nullptr,
                     n_avoid ? ctx->st_lag : nullptr, n_align ? ctx->st_rate : nullptr, n_attract ? ctx->st_intake : nullptr, speed ? ctx->st_sf : nullptr};
    ens_export_pass_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, ctx->stream>>>(o);
    ctx->launches += 1;
    CK(cudaGetLastError());
    cudaStream_t s = ctx->stream;
    if (cnt4) CK(cudaMemcpyAsync(cnt4, ctx->st_i4, 4 * tot * sizeof(int), cudaMemcpyDeviceToHost, s));
    if (turn) CK(cudaMemcpyAsync(turn, ctx->st_dir, tot * sizeof(double), cudaMemcpyDeviceToHost, s));
    if (n_avoid) CK(cudaMemcpyAsync(n_avoid, ctx->st_lag, tot * sizeof(int), cudaMemcpyDeviceToHost, s));
    if (n_align) CK(cudaMemcpyAsync(n_align, ctx->st_rate, tot * sizeof(int), cudaMemcpyDeviceToHost, s));
    if (n_attract) CK(cudaMemcpyAsync(n_attract, ctx->st_intake, tot * sizeof(int), cudaMemcpyDeviceToHost, s));
    if (speed) CK(cudaMemcpyAsync(speed, ctx->st_sf, tot * sizeof(double), cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
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
    p->halo_capacity = 0; p->migrant_capacity = 0; p->force_slab = 0;
    model_defaults(&p->model);
    return 0;
}

int fishabm_model_params_default(fishabm_model_params* m) {
    if (!m) return FISHABM_E_ARG;
    model_defaults(m);
    return 0;
}

const char* fishabm_last_error(const fishabm_ctx* ctx) { return ctx ? ctx->err_msg.c_str() : g_create_error.c_str(); }

int fishabm_create(const fishabm_params* p, fishabm_ctx** out) {
    fishabm_ctx* ctx = nullptr;  // for CK/fail before the context exists
    if (!p || !out) return fail(nullptr, FISHABM_E_ARG, "null argument");
    *out = nullptr;
    if (p->struct_size != (int32_t)sizeof(fishabm_params)) return fail(nullptr, FISHABM_E_ARG, "fishabm_params.struct_size mismatch: use fishabm_params_default");
    if (p->n < 1 || p->n >= (1 << 28)) return fail(nullptr, FISHABM_E_ARG, "n must be in [1, 2^28)");
    if (p->w % CELL_I || p->h % CELL_I || p->w < 3 * CELL_I || p->h < 3 * CELL_I)
        return fail(nullptr, FISHABM_E_ARG, "w and h must be multiples of 200 and at least 600 (got %d x %d)", p->w, p->h);
    if (p->mode != FISHABM_MODE_CELL_FP32 && p->mode != FISHABM_MODE_ENSEMBLE && p->mode != FISHABM_MODE_BRUTE_FP64)
        return fail(nullptr, FISHABM_E_ARG, "unknown mode %d", p->mode);
    if (p->mode == FISHABM_MODE_BRUTE_FP64 && (p->n > 65536 || p->nranks != 1 || p->force_slab || p->replicates != 1))
        return fail(nullptr, FISHABM_E_ARG, "FISHABM_MODE_BRUTE_FP64 (fp64 all-pairs neighbour pass, O(n^2)) is for one whole school of n <= 65536 fish");
    if (p->replicates < 1) return fail(nullptr, FISHABM_E_ARG, "replicates must be >= 1");
    if (const char* why = model_invalid(p->model)) return fail(nullptr, FISHABM_E_ARG, "fishabm_params.model: %s", why);
    const bool ens = p->replicates > 1 || p->mode == FISHABM_MODE_ENSEMBLE;
    if (ens && (p->n > ENS_MAX_FISH || p->nranks != 1 || p->force_slab))
        return fail(nullptr, FISHABM_E_ARG, "ensembles need n <= %d fish per school and no slabs (replicates shard across GPUs by replicate0)", ENS_MAX_FISH);
    if (ens && (long long)p->replicates * p->n >= (1LL << 31)) return fail(nullptr, FISHABM_E_ARG, "replicates * n must stay below 2^31");
    if (p->nranks < 1 || p->rank < 0 || p->rank >= p->nranks) return fail(nullptr, FISHABM_E_ARG, "rank %d / nranks %d", p->rank, p->nranks);
    // a fish that migrates into a slab's last column must not at the same time be a ghost for the slab beyond it
    // (that slab's halo was packed before the migrant arrived): every slab needs two owned columns
    if (p->nranks > 1 && 2 * p->nranks > p->w / CELL_I)
        return fail(nullptr, FISHABM_E_ARG, "nranks %d: every slab needs at least 2 of the %d cell columns (w/200)", p->nranks, p->w / CELL_I);
    if (p->n_global != 0 && p->n_global != p->n) return fail(nullptr, FISHABM_E_ARG, "n_global must be 0 or n: n is the size of the whole school on every rank");
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
    const int nx_g = p->w / CELL_I, ny = p->h / CELL_I;
    c->ens = ens;
    c->R = p->replicates;
    c->slab = p->nranks > 1 || p->force_slab;
    c->col0 = (int)((long long)p->rank * nx_g / p->nranks);
    const int col1 = (int)((long long)(p->rank + 1) * nx_g / p->nranks);
    c->own = col1 - c->col0;
    c->left = (p->rank + p->nranks - 1) % p->nranks;
    c->right = (p->rank + 1) % p->nranks;
    Grid& g = c->g;
    g.ny = ny; g.nx_g = nx_g;
    if (c->slab) {
        g.lnx = c->own + 2; g.wrap_x = 0; g.gx0 = (c->col0 - 1 + nx_g) % nx_g;
        g.own_c0 = ny; g.own_c1 = (g.lnx - 1) * ny;
    } else {
        g.lnx = nx_g; g.wrap_x = 1; g.gx0 = 0; g.own_c0 = 0; g.own_c1 = nx_g * ny;
    }
    g.ncell = g.lnx * ny;
    c->ncell = g.ncell;
    c->qrows = p->h / FOOD_I; c->qcols = c->own * FOOD_PER_CELL; c->qcols_g = p->w / FOOD_I; c->qcol0 = c->col0 * FOOD_PER_CELL;
    c->speed_norm = p->speed_norm > 0 ? p->speed_norm : (double)p->n;
    c->model = to_device_model(p->model);
    if (c->slab) {
        const double per_col = (double)p->n / nx_g;
        c->cap_g = p->halo_capacity > 0 ? p->halo_capacity : (int)(2.0 * per_col) + 4096;
        c->cap_m = p->migrant_capacity > 0 ? p->migrant_capacity : (int)(0.25 * per_col) + 2048;
        const double want = p->capacity > 0 ? (double)p->capacity : 1.3 * per_col * (c->own + 4) + 16384.0;
        if (want >= (double)(1 << 28)) { delete c; return fail(nullptr, FISHABM_E_ARG, "slab capacity must stay below 2^28 entries"); }
        c->cap = (int)want;
        c->msg_bytes = SlabMsg::bytes(c->cap_m, c->cap_g);
    } else {
        c->cap = p->n;
    }
    if (ens) {
        c->ens_smem = ensemble_smem_bytes(p->n, c->qrows * c->qcols);
        if (c->ens_smem > 227 * 1024) { delete c; return fail(nullptr, FISHABM_E_ARG, "ensemble school does not fit in shared memory (%zu bytes): smaller n or window", c->ens_smem); }
    }
    // an ensemble reuses the single-school fields with R*n entries (one buffer set, no sort structures)
    const size_t n = (size_t)c->n * c->R, cap = ens ? n : (size_t)c->cap, nq = (size_t)c->qrows * c->qcols * c->R;
    bool ok = true;
    auto A = [&](cudaError_t r) { if (r != cudaSuccess && ok) { ok = false; g_create_error = cudaGetErrorString(r); } };
    A(preload_kernels());
    {   // the host libm's unit vectors of integer headings (kernels.cuh: g_unit_deg), written exactly as fish_mod.cpp:232-233
        static double2 tab[UNIT_DEG_MAX - UNIT_DEG_MIN + 1];
        for (int d = UNIT_DEG_MIN; d <= UNIT_DEG_MAX; ++d) {
            volatile double dir = (double)d;  // no constant folding: the run-time libm is the reference's
            tab[d - UNIT_DEG_MIN] = make_double2(cos(dir * M_PI / 180), sin(dir * M_PI / 180));
        }
        A(cudaMemcpyToSymbol(g_unit_deg, tab, sizeof tab));
    }
    A(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    A(cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
    A(cudaEventCreateWithFlags(&c->ev_copy, cudaEventDisableTiming));
    A(cudaEventCreate(&c->ev0)); A(cudaEventCreate(&c->ev1));
    for (int k = 0; k < (ens ? 1 : 2) && ok; ++k) {
        A(dalloc(&c->buf[k].rec, cap)); A(dalloc(&c->buf[k].cold0, cap)); A(dalloc(&c->buf[k].cold1, cap)); A(dalloc(&c->buf[k].cell, cap));
    }
    A(dalloc(&c->next_cell, cap)); A(dalloc(&c->next_rank, cap));
    if (!ens) A(dalloc(&c->rec_tmp, cap));
    if (const char* fu = getenv("FISHABM_FUSED_UPDATE")) c->fuse_update = atoi(fu) != 0;
    A(dalloc(&c->cell_count, (size_t)c->ncell + 1)); A(dalloc(&c->cell_start, (size_t)c->ncell + 1));
    A(dalloc(&c->tile_sums, (size_t)cdiv(c->ncell, SCAN_TILE) + 1));
    A(dalloc(&c->cnt, cap)); A(dalloc(&c->turn, cap)); A(dalloc(&c->aux, cap));
    A(dalloc(&c->quality, nq));
    A(dalloc(&c->upd_count, 2)); A(dalloc(&c->upd_idx, cap)); A(dalloc(&c->upd_val, cap));
    A(dalloc(&c->stats, 8)); A(dalloc(&c->work_counter, 2)); A(dalloc(&c->bad, 1)); A(dalloc(&c->pos_of_id, n));
    A(dalloc(&c->n_unsorted, 1)); A(dalloc(&c->err, 1)); A(dalloc(&c->d_sums, (size_t)c->R)); A(dalloc(&c->scan_sync, 4)); A(dalloc(&c->scan_desc, (size_t)cdiv(c->ncell, SCAN_TILE) + 1));
    A(dalloc(&c->d_models, (size_t)c->R));
    if (ok) {
        std::vector<Model> hm((size_t)c->R, c->model);
        A(cudaMemcpy(c->d_models, hm.data(), hm.size() * sizeof(Model), cudaMemcpyHostToDevice));
    }
    if (ok) { A(cudaMemset(c->scan_sync, 0, 4 * sizeof(int))); A(cudaMemset(c->scan_desc, 0, ((size_t)cdiv(c->ncell, SCAN_TILE) + 1) * sizeof(unsigned long long))); }
    if (const char* sc = getenv("FISHABM_SCAN")) c->scan_chained = strcmp(sc, "coop") != 0;
    A(dalloc(&c->st_coords, 2 * n)); A(dalloc(&c->st_dir, n)); A(dalloc(&c->st_speed, n)); A(dalloc(&c->st_sf, n));
    A(dalloc(&c->st_lag, n)); A(dalloc(&c->st_rate, n)); A(dalloc(&c->st_intake, n)); A(dalloc(&c->st_i4, 4 * n)); A(dalloc(&c->st_ids, n));
    if (c->slab) {
        A(dalloc(&c->out_left, c->msg_bytes)); A(dalloc(&c->out_right, c->msg_bytes));
        A(dalloc(&c->in_left, c->msg_bytes)); A(dalloc(&c->in_right, c->msg_bytes));
        c->msg_stride = (c->msg_bytes + 255) & ~(size_t)255;
        c->inbox_bytes = 4 * c->msg_stride + 8 * 128;
        A(dalloc(&c->inbox, c->inbox_bytes)); A(dalloc(&c->ctr, 2));
        if (ok) A(cudaMemset(c->inbox, 0, c->inbox_bytes));
        if (const char* t = getenv("FISHABM_SLAB_TIMEOUT_MS")) c->timeout_ns = atoll(t) * 1000000LL;
    }
    if (ok) {
        A(cudaMemset(c->cell_count, 0, ((size_t)c->ncell + 1) * sizeof(int)));
        A(cudaMemset(c->cell_start, 0, ((size_t)c->ncell + 1) * sizeof(int)));
        A(cudaMemset(c->quality, 0, nq));
        A(cudaMemset(c->upd_count, 0, 2 * sizeof(int)));
        A(cudaMemset(c->work_counter, 0, 2 * sizeof(int)));
        A(cudaMemset(c->err, 0, sizeof(int)));
        A(cudaMemset(c->stats, 0, 8 * sizeof(unsigned long long)));
        A(cudaMemset(c->aux, 0, cap * sizeof(uint32_t)));
        if (c->slab) {
            A(cudaMemset(c->out_left, 0, sizeof(SlabHeader))); A(cudaMemset(c->out_right, 0, sizeof(SlabHeader)));
            A(cudaMemset(c->in_left, 0, sizeof(SlabHeader))); A(cudaMemset(c->in_right, 0, sizeof(SlabHeader)));
        }
    }
    if (ok && ens) A(cudaFuncSetAttribute(ensemble_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)c->ens_smem));
    if (ok) {
        A(cudaFuncSetAttribute(neighbour_kernel_v4<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)V4_SMEM_BYTES));
        A(cudaFuncSetAttribute(neighbour_kernel_v4<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)V4_SMEM_BYTES));
        A(cudaFuncSetAttribute(neighbour_kernel_v3<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)V3_SMEM_BYTES));
        A(cudaFuncSetAttribute(neighbour_kernel_v3<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)V3_SMEM_BYTES));
        A(cudaFuncSetAttribute(neighbour_kernel_v3<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)V3_SMEM_BYTES));
        A(cudaFuncSetAttribute(neighbour_kernel_v3<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)V3_SMEM_BYTES));
    }
    const char* variant = getenv("FISHABM_NEIGHBOUR");
    if (variant && strcmp(variant, "v1") == 0) c->nb_variant = 1;
    if (variant && strcmp(variant, "v3") == 0) c->nb_variant = 3;
    if (ok) {
        int per_sm = 0, sms = 0;
        A(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, p->device));
        static_assert(V4_WARPS == V3_WARPS && V4_THREADS == V3_THREADS, "the work split assumes the same CTA shape");
        if (c->nb_variant == 4) A(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, neighbour_kernel_v4<false>, V4_THREADS, V4_SMEM_BYTES));
        else A(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, neighbour_kernel_v3<false, true>, V3_THREADS, V3_SMEM_BYTES));
        c->nb_grid = std::max(1, per_sm) * std::max(1, sms);  // persistent: one resident wave
        int scan_per_sm = 0;
        A(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&scan_per_sm, scan_fused_kernel, SCAN_THREADS, 0));
        c->scan_resident = scan_per_sm * sms / 2;  // the fused scan spins on a flag: only with every block resident, with margin
    }
    c->h_q8 = (uint8_t*)malloc((size_t)c->qrows * c->qcols > 0 ? (size_t)c->qrows * c->qcols : 1);
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
    if (c->comm && nccl().CommDestroy) nccl().CommDestroy(c->comm);
    for (void* q : c->ipc_opened) if (q) cudaIpcCloseMemHandle(q);
    cudaFree(c->inbox); cudaFree(c->ctr);
    for (int k = 0; k < 2; ++k) { cudaFree(c->buf[k].rec); cudaFree(c->buf[k].cold0); cudaFree(c->buf[k].cold1); cudaFree(c->buf[k].cell); }
    cudaFree(c->rec_tmp);
    cudaFree(c->next_cell); cudaFree(c->next_rank); cudaFree(c->cell_count); cudaFree(c->cell_start); cudaFree(c->tile_sums);
    cudaFree(c->cnt); cudaFree(c->turn); cudaFree(c->aux); cudaFree(c->quality);
    cudaFree(c->upd_count); cudaFree(c->upd_idx); cudaFree(c->upd_val); cudaFree(c->stats); cudaFree(c->work_counter); cudaFree(c->bad); cudaFree(c->pos_of_id);
    cudaFree(c->n_unsorted); cudaFree(c->err); cudaFree(c->d_sums); cudaFree(c->scan_sync); cudaFree(c->scan_desc); cudaFree(c->d_models);
    cudaFree(c->st_coords); cudaFree(c->st_dir); cudaFree(c->st_speed); cudaFree(c->st_sf);
    cudaFree(c->st_lag); cudaFree(c->st_rate); cudaFree(c->st_intake); cudaFree(c->st_i4); cudaFree(c->st_ids);
    cudaFree(c->out_left); cudaFree(c->out_right); cudaFree(c->in_left); cudaFree(c->in_right);
    for (auto& pr : c->nb_events) { cudaEventDestroy(pr.first); cudaEventDestroy(pr.second); }
    for (cudaEvent_t e : c->ph_events) cudaEventDestroy(e);
    if (c->ev0) cudaEventDestroy(c->ev0);
    if (c->ev1) cudaEventDestroy(c->ev1);
    if (c->stream) cudaStreamDestroy(c->stream);
    if (c->copy_stream) cudaStreamDestroy(c->copy_stream);
    if (c->ev_copy) cudaEventDestroy(c->ev_copy);
    free(c->h_q8);
    delete c;
}

int fishabm_set_state(fishabm_ctx* ctx, const double* coords, const double* dir, const double* speed, const double* speed_factor,
                      const int32_t* lag, const int32_t* sample_rate, const int32_t* food_intake) {
    if (!ctx || !coords || !dir || !speed || !speed_factor || !lag || !sample_rate || !food_intake)
        return fail(ctx, FISHABM_E_ARG, "fishabm_set_state: null argument");
    if (ctx->step_open) return fail(ctx, FISHABM_E_STATE, "a slab step is open: call fishabm_slab_end_step first");
    CK(cudaSetDevice(ctx->p.device));
    if (ctx->have_state) { int rcf = flush_pending(ctx); if (rcf) return rcf; }
    ctx->pending_sample = false;
    const size_t n = (size_t)ctx->n * ctx->R;
    cudaStream_t s = ctx->stream;
    if (ctx->slab) {  // a fish may not outrun its halo: |step| = speed * speed_factor <= 3 * speed must stay below one cell
        int bad_speed = 0;
        for (size_t i = 0; i < n; ++i) bad_speed |= !(speed[i] * 3.0 < (double)CELL_I);
        for (size_t i = 0; bad_speed && i < n; ++i)
            if (!(speed[i] * 3.0 < (double)CELL_I)) return fail(ctx, FISHABM_E_ARG, "fishabm_set_state: speed %g of fish %zu: slabs need 3*speed < 200", speed[i], i);
    }
    CK(cudaMemcpyAsync(ctx->st_coords, coords, 2 * n * sizeof(double), cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(ctx->st_dir, dir, n * sizeof(double), cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(ctx->st_speed, speed, n * sizeof(double), cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(ctx->st_sf, speed_factor, n * sizeof(double), cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(ctx->st_lag, lag, n * sizeof(int), cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(ctx->st_rate, sample_rate, n * sizeof(int), cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(ctx->st_intake, food_intake, n * sizeof(int), cudaMemcpyHostToDevice, s));
    CK(cudaMemsetAsync(ctx->bad, 0, sizeof(int), s));
    CK(cudaMemsetAsync(ctx->err, 0, sizeof(int), s));
    if (ctx->ens) {
        const Buffers& b = ctx->buf[0];
        EnsIoArgs a{ctx->st_coords, ctx->st_dir, ctx->st_speed, ctx->st_sf, ctx->st_lag, ctx->st_rate, ctx->st_intake, b.rec, b.cold0, b.cold1,
                    (long long)n, ctx->g.nx_g, ctx->g.ny, ctx->bad, (double)ctx->p.w, (double)ctx->p.h};
        ens_ingest_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(a);
        ctx->launches += 1;
        int bad = 0;
        CK(cudaMemcpyAsync(&bad, ctx->bad, sizeof(int), cudaMemcpyDeviceToHost, s));
        CK(cudaStreamSynchronize(s));
        if (bad) { ctx->have_state = false; return fail(ctx, FISHABM_E_ARG, "fishabm_set_state: coordinates must be finite and inside [0,w] x [0,h]"); }
        ctx->have_state = true;
        return 0;
    }
    CK(cudaMemsetAsync(ctx->cell_count, 0, ((size_t)ctx->ncell + 1) * sizeof(int), s));
    ctx->h_n = ctx->slab ? 0 : ctx->n;  // whole school: entry i = fish i; slab: entries are claimed
    CK(cudaMemcpyAsync(ctx->n_unsorted, &ctx->h_n, sizeof(int), cudaMemcpyHostToDevice, s));
    const Buffers& b = ctx->buf[ctx->cur];
    IngestArgs a{ctx->st_coords, ctx->st_dir, ctx->st_speed, ctx->st_sf, ctx->st_lag, ctx->st_rate, ctx->st_intake,
                 b.rec, b.cold0, b.cold1, ctx->next_cell, ctx->next_rank, ctx->cell_count, ctx->n_unsorted, nullptr,
                 ctx->n, ctx->cap, ctx->g, ctx->n, ctx->bad, (double)ctx->p.w, (double)ctx->p.h};
    ingest_kernel<<<cdiv(ctx->n, 256), 256, 0, s>>>(a);
    ctx->launches += 1;
    int rc = rebuild(ctx);
    if (rc) return rc;
    int bad = 0;
    CK(cudaMemcpyAsync(&bad, ctx->bad, sizeof(int), cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    if (bad & 1) { ctx->have_state = false; return fail(ctx, FISHABM_E_ARG, "fishabm_set_state: coordinates must be finite and inside [0,w] x [0,h]"); }
    if (bad & 2) { ctx->have_state = false; return fail(ctx, FISHABM_E_NOMEM, "fishabm_set_state: this slab holds more than its capacity of %d entries: raise fishabm_params.capacity", ctx->cap); }
    ctx->have_state = true;
    return 0;
}

int fishabm_set_state_owned(fishabm_ctx* ctx, int32_t count, const int32_t* ids, const double* coords, const double* dir,
                            const double* speed, const double* speed_factor, const int32_t* lag, const int32_t* sample_rate,
                            const int32_t* food_intake) {
    if (!ctx || count < 0 || (count > 0 && (!ids || !coords || !dir || !speed || !speed_factor || !lag || !sample_rate || !food_intake)))
        return fail(ctx, FISHABM_E_ARG, "fishabm_set_state_owned: null argument");
    if (!ctx->slab) return fail(ctx, FISHABM_E_STATE, "fishabm_set_state_owned is for slab contexts (use fishabm_set_state)");
    if (ctx->step_open) return fail(ctx, FISHABM_E_STATE, "a slab step is open: call fishabm_slab_end_step first");
    if (count > ctx->n || count > ctx->cap) return fail(ctx, FISHABM_E_NOMEM, "fishabm_set_state_owned: %d fish exceed the slab's capacity %d", count, ctx->cap);
    CK(cudaSetDevice(ctx->p.device));
    if (ctx->have_state) { int rcf = flush_pending(ctx); if (rcf) return rcf; }
    ctx->pending_sample = false;
    const size_t n = (size_t)count;
    cudaStream_t s = ctx->stream;
    { int rcb = begin_timed(ctx); if (rcb) return rcb; }
    // the uploads go to staging arrays; the state itself is only touched by the ingest kernel below
    CK(cudaMemcpyAsync(ctx->st_coords, coords, 2 * n * sizeof(double), cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(ctx->st_dir, dir, n * sizeof(double), cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(ctx->st_speed, speed, n * sizeof(double), cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(ctx->st_sf, speed_factor, n * sizeof(double), cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(ctx->st_lag, lag, n * sizeof(int), cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(ctx->st_rate, sample_rate, n * sizeof(int), cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(ctx->st_intake, food_intake, n * sizeof(int), cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(ctx->st_ids, ids, n * sizeof(int), cudaMemcpyHostToDevice, s));
    // ids and speeds are validated by the ingest kernel (it reads them anyway; a host-side pass over the 8.4M fish of a 2-GPU
    // rank cost 10 ms in front of a 9 ms upload) and reported by fishabm_wait like every other bad input
    CK(cudaMemsetAsync(ctx->bad, 0, sizeof(int), s));
    CK(cudaMemsetAsync(ctx->err, 0, sizeof(int), s));
    CK(cudaMemsetAsync(ctx->cell_count, 0, ((size_t)ctx->ncell + 1) * sizeof(int), s));
    CK(cudaMemsetAsync(ctx->n_unsorted, 0, sizeof(int), s));
    {   // 1. this slab's own fish -> sorted arrays (no ghosts yet)
        const Buffers& b = ctx->buf[ctx->cur];
        IngestArgs a{ctx->st_coords, ctx->st_dir, ctx->st_speed, ctx->st_sf, ctx->st_lag, ctx->st_rate, ctx->st_intake,
                     b.rec, b.cold0, b.cold1, ctx->next_cell, ctx->next_rank, ctx->cell_count, ctx->n_unsorted, ctx->st_ids,
                     count, ctx->cap, ctx->g, ctx->n, ctx->bad, (double)ctx->p.w, (double)ctx->p.h};
        if (count > 0) ingest_kernel<<<cdiv(count, 256), 256, 0, s>>>(a);
        ctx->launches += 1;
        int rc = rebuild(ctx);
        if (rc) return rc;
    }
    {   // 2. boundary columns -> the neighbours' halos: one exchange, then the ordinary re-sort takes the ghosts in
        const Buffers& b = ctx->buf[ctx->cur];
        HaloRefreshArgs h{b.rec, b.cold1, b.cell, ctx->cell_start, ctx->next_cell, ctx->next_rank, ctx->cell_count, ctx->g, {}, {}};
        int rc = open_out_messages(ctx, &h.out_left, &h.out_right);
        if (rc) return rc;
        halo_refresh_kernel<<<cdiv(ctx->cap, 256), 256, 0, s>>>(h);
        ctx->launches += 1;
        CK(cudaGetLastError());
        unsigned char *il = nullptr, *ir = nullptr;
        if ((rc = exchange(ctx, &il, &ir))) return rc;
        if ((rc = finish_move(ctx, il, ir))) return rc;
    }
    ctx->have_state = true;  // provisional: fishabm_wait reports bad input
    return 0;
}

int fishabm_get_state_owned(fishabm_ctx* ctx, int32_t* count, int32_t* ids, double* coords, double* dir, double* speed,
                            double* speed_factor, int32_t* lag, int32_t* sample_rate, int32_t* food_intake) {
    int rc = need_state(ctx);
    if (rc) return rc;
    if (!count) return fail(ctx, FISHABM_E_ARG, "fishabm_get_state_owned: null count");
    if (ctx->ens) return fail(ctx, FISHABM_E_STATE, "fishabm_get_state_owned is for single schools / slabs");
    CK(cudaSetDevice(ctx->p.device));
    int cnt = 0;   // (a pending sample() moves no fish: the owned range and the sorted order are already final)
    if ((rc = owned_range(ctx, &cnt))) return rc;
    if (cnt > *count) { const int have = *count; *count = cnt; return fail(ctx, FISHABM_E_NOMEM, "fishabm_get_state_owned: %d fish owned, arrays hold %d", cnt, have); }
    *count = cnt;
    cudaStream_t s = ctx->stream;
    const size_t n = (size_t)cnt;
    // as in fishabm_get_state: what a pending sample() cannot change (ids, positions, headings, speeds: 36 of 56 bytes per fish)
    // is exported first and goes to the host on the second stream while the pending pass runs
    const bool early = ctx->pending_sample && cnt > 0 && (ids || coords || dir || speed);
    if (early) {
        const Buffers& b0 = ctx->buf[ctx->cur];
        ExportArgs e{b0.rec, b0.cold0, b0.cold1, b0.cell, ctx->cell_start, coords ? ctx->st_coords : nullptr, dir ? ctx->st_dir : nullptr,
                     speed ? ctx->st_speed : nullptr, nullptr, nullptr, nullptr, nullptr, ids ? ctx->st_ids : nullptr, ctx->g, 1};
        export_kernel<<<cdiv(cnt, 256), 256, 0, s>>>(e);
        ctx->launches += 1;
        CK(cudaGetLastError());
        CK(cudaEventRecord(ctx->ev_copy, s));
        CK(cudaStreamWaitEvent(ctx->copy_stream, ctx->ev_copy, 0));
        if (ids) CK(cudaMemcpyAsync(ids, ctx->st_ids, n * sizeof(int), cudaMemcpyDeviceToHost, ctx->copy_stream));
        if (coords) CK(cudaMemcpyAsync(coords, ctx->st_coords, 2 * n * sizeof(double), cudaMemcpyDeviceToHost, ctx->copy_stream));
        if (dir) CK(cudaMemcpyAsync(dir, ctx->st_dir, n * sizeof(double), cudaMemcpyDeviceToHost, ctx->copy_stream));
        if (speed) CK(cudaMemcpyAsync(speed, ctx->st_speed, n * sizeof(double), cudaMemcpyDeviceToHost, ctx->copy_stream));
        ids = nullptr; coords = nullptr; dir = nullptr; speed = nullptr;   // done: the second export leaves them out
    }
    struct CopyGuard {   // whatever happens below, the early downloads have landed before the call returns
        fishabm_ctx* c; bool on;
        ~CopyGuard() { if (on) cudaStreamSynchronize(c->copy_stream); }
    } guard{ctx, early};
    { int rcf = flush_pending(ctx); if (rcf) return rcf; }
    const Buffers& b = ctx->buf[ctx->cur];
    ExportArgs a{b.rec, b.cold0, b.cold1, b.cell, ctx->cell_start,
                 coords ? ctx->st_coords : nullptr, dir ? ctx->st_dir : nullptr, speed ? ctx->st_speed : nullptr,
                 speed_factor ? ctx->st_sf : nullptr, lag ? ctx->st_lag : nullptr, sample_rate ? ctx->st_rate : nullptr,
                 food_intake ? ctx->st_intake : nullptr, ids ? ctx->st_ids : nullptr, ctx->g, 1};
    if (cnt > 0) export_kernel<<<cdiv(cnt, 256), 256, 0, s>>>(a);
    ctx->launches += 1;
    CK(cudaGetLastError());
    if (ids) CK(cudaMemcpyAsync(ids, ctx->st_ids, n * sizeof(int), cudaMemcpyDeviceToHost, s));
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

int fishabm_get_state(fishabm_ctx* ctx, double* coords, double* dir, double* speed, double* speed_factor, int32_t* lag,
                      int32_t* sample_rate, int32_t* food_intake) {
    int rc = need_state(ctx);
    if (rc) return rc;
    CK(cudaSetDevice(ctx->p.device));
    if (ctx->ens) {
        const size_t n = (size_t)ctx->n * ctx->R;
        cudaStream_t s = ctx->stream;
        const Buffers& b = ctx->buf[0];
        EnsIoArgs a{coords ? ctx->st_coords : nullptr, dir ? ctx->st_dir : nullptr, speed ? ctx->st_speed : nullptr, speed_factor ? ctx->st_sf : nullptr,
                    lag ? ctx->st_lag : nullptr, sample_rate ? ctx->st_rate : nullptr, food_intake ? ctx->st_intake : nullptr, b.rec, b.cold0, b.cold1,
                    (long long)n, ctx->g.nx_g, ctx->g.ny, ctx->bad, (double)ctx->p.w, (double)ctx->p.h};
        ens_export_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(a);
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
    cudaStream_t s = ctx->stream;
    // A pending sample() (fishabm_step leaves the last one to the next neighbour pass) changes neither positions nor headings
    // nor speeds (fish_mod.cpp:540-586 writes lag, sample_rate, food_intake, speed_factor): those three arrays -- 32 of the 52
    // bytes per fish -- are exported first and travel to the host on a second stream WHILE the pending pass runs.
    const bool early = ctx->pending_sample && !ctx->slab && (coords || dir || speed);
    if (early) {
        const Buffers& b0 = ctx->buf[ctx->cur];
        ExportArgs e{b0.rec, b0.cold0, b0.cold1, b0.cell, ctx->cell_start, coords ? ctx->st_coords : nullptr, dir ? ctx->st_dir : nullptr,
                     speed ? ctx->st_speed : nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, ctx->g, 0};
        export_kernel<<<cdiv(ctx->n, 256), 256, 0, s>>>(e);
        ctx->launches += 1;
        CK(cudaGetLastError());
        CK(cudaEventRecord(ctx->ev_copy, s));
        CK(cudaStreamWaitEvent(ctx->copy_stream, ctx->ev_copy, 0));
        const size_t n = (size_t)ctx->n;
        if (coords) CK(cudaMemcpyAsync(coords, ctx->st_coords, 2 * n * sizeof(double), cudaMemcpyDeviceToHost, ctx->copy_stream));
        if (dir) CK(cudaMemcpyAsync(dir, ctx->st_dir, n * sizeof(double), cudaMemcpyDeviceToHost, ctx->copy_stream));
        if (speed) CK(cudaMemcpyAsync(speed, ctx->st_speed, n * sizeof(double), cudaMemcpyDeviceToHost, ctx->copy_stream));
        coords = nullptr; dir = nullptr; speed = nullptr;   // done: the second export below leaves them out
    }
    { int rcf = flush_pending(ctx); if (rcf) { if (early) cudaStreamSynchronize(ctx->copy_stream); return rcf; } }
    int cnt = 0;
    if ((rc = owned_range(ctx, &cnt))) return rc;
    const Buffers& b = ctx->buf[ctx->cur];
    ExportArgs a{b.rec, b.cold0, b.cold1, b.cell, ctx->cell_start,
                 coords ? ctx->st_coords : nullptr, dir ? ctx->st_dir : nullptr, speed ? ctx->st_speed : nullptr,
                 speed_factor ? ctx->st_sf : nullptr, lag ? ctx->st_lag : nullptr, sample_rate ? ctx->st_rate : nullptr,
                 food_intake ? ctx->st_intake : nullptr, ctx->slab ? ctx->st_ids : nullptr, ctx->g, ctx->slab ? 1 : 0};
    if (cnt > 0) export_kernel<<<cdiv(cnt, 256), 256, 0, s>>>(a);
    ctx->launches += 1;
    CK(cudaGetLastError());
    struct CopyGuard {   // whatever happens below, the early downloads have landed before the call returns
        fishabm_ctx* c; bool on;
        ~CopyGuard() { if (on) cudaStreamSynchronize(c->copy_stream); }
    } guard{ctx, early};
    if ((rc = fetch_ids(ctx, cnt))) return rc;
    if ((rc = fetch(ctx, coords, ctx->st_coords, cnt, 2))) return rc;
    if ((rc = fetch(ctx, dir, ctx->st_dir, cnt, 1))) return rc;
    if ((rc = fetch(ctx, speed, ctx->st_speed, cnt, 1))) return rc;
    if ((rc = fetch(ctx, speed_factor, ctx->st_sf, cnt, 1))) return rc;
    if ((rc = fetch(ctx, lag, ctx->st_lag, cnt, 1))) return rc;
    if ((rc = fetch(ctx, sample_rate, ctx->st_rate, cnt, 1))) return rc;
    if ((rc = fetch(ctx, food_intake, ctx->st_intake, cnt, 1))) return rc;
    CK(cudaStreamSynchronize(s));
    return 0;
}

int fishabm_initialize(fishabm_ctx* ctx, double speed) {
    if (!ctx) return FISHABM_E_ARG;
    // initialize(), fish_mod.cpp:183-206.  Tiny and one-off: the state is composed on the host from the same
    // counter-based streams the device uses, then ingested like any other state.
    const size_t n = (size_t)ctx->n * ctx->R;
    std::vector<double> coords, dir, sp, sf;
    std::vector<int32_t> z;
    try { coords.resize(2 * n); dir.resize(n); sp.assign(n, speed); sf.assign(n, 1.0); z.assign(n, 0); }
    catch (...) { return fail(ctx, FISHABM_E_NOMEM, "out of host memory"); }
    for (int r = 0; r < ctx->R; ++r) {
        const uint32_t rep = ctx->p.replicate0 + (uint32_t)r;
        for (int f = 0; f < ctx->n; ++f) {
            const size_t i = (size_t)r * ctx->n + f;
            dir[i] = (double)fishrng::draw(ctx->p.seed, rep, 0, (uint32_t)f, STREAM_RANDOMWALK, 0, -45, 45);   // distrib_randomwalk as the reference declares it (:30,196)
            coords[2 * i] = ctx->p.w / 2 + fishrng::draw(ctx->p.seed, rep, 0, (uint32_t)f, STREAM_INIT, 0, 0, 200) - 100;
            coords[2 * i + 1] = ctx->p.h / 2 + fishrng::draw(ctx->p.seed, rep, 1, (uint32_t)f, STREAM_INIT, 0, 0, 200) - 100;
        }
    }
    return fishabm_set_state(ctx, coords.data(), dir.data(), sp.data(), sf.data(), z.data(), z.data(), z.data());
}

// the food grid arguments are the WHOLE grid (h/20 x w/20); a slab keeps / returns its owned columns
int fishabm_set_quality(fishabm_ctx* ctx, const int32_t* q, int32_t rows, int32_t cols, int32_t row_stride) {
    if (!ctx || !q) return fail(ctx, FISHABM_E_ARG, "fishabm_set_quality: null argument");
    if (rows != ctx->qrows || cols != ctx->qcols_g || row_stride < cols)
        return fail(ctx, FISHABM_E_ARG, "fishabm_set_quality: grid must be %d x %d (h/20 x w/20), got %d x %d", ctx->qrows, ctx->qcols_g, rows, cols);
    for (int r = 0; r < rows; ++r)
        for (int c = 0; c < ctx->qcols; ++c) {
            const int v = q[(size_t)r * row_stride + ctx->qcol0 + c];
            if (v < 0 || v > 255) return fail(ctx, FISHABM_E_ARG, "fishabm_set_quality: value %d at (%d,%d) outside 0..255", v, r, ctx->qcol0 + c);
            ctx->h_q8[(size_t)r * ctx->qcols + c] = (uint8_t)v;
        }
    CK(cudaSetDevice(ctx->p.device));
    { int rcf = flush_pending(ctx); if (rcf) return rcf; }
    CK(cudaMemcpyAsync(ctx->quality, ctx->h_q8, (size_t)rows * ctx->qcols, cudaMemcpyHostToDevice, ctx->stream));
    if (ctx->ens && ctx->R > 1) {  // every replicate starts from the same grid
        const long long tot = (long long)rows * ctx->qcols * ctx->R;
        ens_broadcast_quality_kernel<<<1184, 256, 0, ctx->stream>>>(ctx->quality, ctx->quality, rows * ctx->qcols, tot);
        ctx->launches += 1;
    }
    CK(cudaStreamSynchronize(ctx->stream));
    return 0;
}

int fishabm_get_quality(fishabm_ctx* ctx, int32_t replicate, int32_t* q, int32_t rows, int32_t cols, int32_t row_stride) {
    if (!ctx || !q) return fail(ctx, FISHABM_E_ARG, "fishabm_get_quality: null argument");
    if (replicate < 0 || replicate >= ctx->R || rows != ctx->qrows || cols != ctx->qcols_g || row_stride < cols)
        return fail(ctx, FISHABM_E_ARG, "fishabm_get_quality: grid is %d x %d, replicates 0..%d", ctx->qrows, ctx->qcols_g, ctx->R - 1);
    CK(cudaSetDevice(ctx->p.device));
    { int rcf = flush_pending(ctx); if (rcf) return rcf; }
    CK(cudaMemcpyAsync(ctx->h_q8, ctx->quality + (size_t)replicate * rows * ctx->qcols, (size_t)rows * ctx->qcols, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    for (int r = 0; r < rows; ++r)
        for (int c = 0; c < ctx->qcols; ++c) q[(size_t)r * row_stride + ctx->qcol0 + c] = ctx->h_q8[(size_t)r * ctx->qcols + c];
    return 0;
}

int fishabm_sum_quality(fishabm_ctx* ctx, int64_t* sums) {
    if (!ctx || !sums) return fail(ctx, FISHABM_E_ARG, "fishabm_sum_quality: null argument");
    CK(cudaSetDevice(ctx->p.device));
    { int rcf = flush_pending(ctx); if (rcf) return rcf; }
    if (ctx->ens) {
        ens_sum_quality_kernel<<<ctx->R, 256, 0, ctx->stream>>>(ctx->quality, ctx->qrows * ctx->qcols, ctx->d_sums);
        ctx->launches += 1;
        CK(cudaGetLastError());
        CK(cudaMemcpyAsync(sums, ctx->d_sums, (size_t)ctx->R * sizeof(long long), cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
        return 0;
    }
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

int fishabm_create_environment(uint64_t seed, uint32_t replicate, int32_t n_seeds, int32_t* q, int32_t rows, int32_t cols,
                               int32_t row_stride) {
    // fish_mod.cpp:482-531, the sweeps exactly as written (including `quality[u][i+1]` tested for non-zero rather than
    // `> p`, and the inclusive 0..rows / 0..cols seed ranges whose out-of-grid hits have no effect)
    if (!q || rows < 2 || cols < 2 || row_stride < cols || n_seeds < 0) return FISHABM_E_ARG;
    const uint32_t STREAM_QUALITY = 3, STREAM_ENV = 7;
    uint32_t kq = 0, ke = 0;
    auto Q = [&](int u, int i) -> int32_t& { return q[(size_t)u * row_stride + i]; };
    for (int u = 0; u < rows; ++u)
        for (int i = 0; i < cols; ++i) Q(u, i) = 0;
    for (int s = 0; s < n_seeds; ++s) {
        const int r = fishrng::draw(seed, replicate, 0, 0, STREAM_ENV, ke++, 0, rows);
        const int c = fishrng::draw(seed, replicate, 0, 0, STREAM_ENV, ke++, 0, cols);
        if (r < rows && c < cols) Q(r, c) = 10;
    }
    for (int p = 0; p < 20; ++p) {
        for (int u = 0; u < rows - 1; ++u)
            for (int i = 0; i < cols - 1; ++i)
                if (Q(u + 1, i) > p || Q(u + 1, i + 1) > p || Q(u, i + 1)) Q(u, i) = fishrng::draw(seed, replicate, 0, 0, STREAM_QUALITY, kq++, 0, 10);
        for (int u = rows - 1; u > 0; --u)
            for (int i = cols - 1; i > 0; --i)
                if (Q(u - 1, i) > p || Q(u - 1, i - 1) > p || Q(u, i - 1)) Q(u, i) = fishrng::draw(seed, replicate, 0, 0, STREAM_QUALITY, kq++, 0, 10);
    }
    return 0;
}

// ---- frozen-state queries ------------------------------------------------------------------------------
static int export_pass(fishabm_ctx* ctx, int32_t* cnt4, double* turn, int32_t* n_avoid, int32_t* n_align, int32_t* n_attract, double* speed) {
    int rc, cnt = 0;
    if ((rc = owned_range(ctx, &cnt))) return rc;
    cudaStream_t s = ctx->stream;
    int* d4 = ctx->st_i4;
    ExportPassArgs a{ctx->buf[ctx->cur].cold1, ctx->cnt, ctx->turn, ctx->aux, ctx->cell_start,
                     cnt4 ? d4 : nullptr, turn ? ctx->st_dir : nullptr, n_avoid ? ctx->st_lag : nullptr,
                     n_align ? ctx->st_rate : nullptr, n_attract ? ctx->st_intake : nullptr, speed ? ctx->st_sf : nullptr,
                     ctx->slab ? ctx->st_ids : nullptr, ctx->g, ctx->slab ? 1 : 0, ctx->speed_norm};
    if (cnt > 0) export_pass_kernel<<<cdiv(cnt, 256), 256, 0, s>>>(a);
    ctx->launches += 1;
    CK(cudaGetLastError());
    if ((rc = fetch_ids(ctx, cnt))) return rc;
    if ((rc = fetch(ctx, cnt4, d4, cnt, 4))) return rc;
    if ((rc = fetch(ctx, turn, ctx->st_dir, cnt, 1))) return rc;
    if ((rc = fetch(ctx, n_avoid, ctx->st_lag, cnt, 1))) return rc;
    if ((rc = fetch(ctx, n_align, ctx->st_rate, cnt, 1))) return rc;
    if ((rc = fetch(ctx, n_attract, ctx->st_intake, cnt, 1))) return rc;
    if ((rc = fetch(ctx, speed, ctx->st_sf, cnt, 1))) return rc;
    CK(cudaStreamSynchronize(s));
    return 0;
}

int fishabm_zone_counts(fishabm_ctx* ctx, int32_t* out4) {
    int rc = need_state(ctx);
    if (rc) return rc;
    if (!out4) return fail(ctx, FISHABM_E_ARG, "null argument");
    CK(cudaSetDevice(ctx->p.device));
    if (ctx->ens) return ens_query(ctx, out4, nullptr, nullptr, nullptr, nullptr, nullptr);
    { int rcf = flush_pending(ctx); if (rcf) return rcf; }
    if ((rc = neighbour_pass(ctx, -1, false, ctx->step, ctx->step, false))) return rc;
    return export_pass(ctx, out4, nullptr, nullptr, nullptr, nullptr, nullptr);
}

int fishabm_social(fishabm_ctx* ctx, double* turn, int32_t* n_avoid, int32_t* n_align, int32_t* n_attract) {
    int rc = need_state(ctx);
    if (rc) return rc;
    if (!turn) return fail(ctx, FISHABM_E_ARG, "null argument");
    CK(cudaSetDevice(ctx->p.device));
    if (ctx->ens) return ens_query(ctx, nullptr, turn, n_avoid, n_align, n_attract, nullptr);
    { int rcf = flush_pending(ctx); if (rcf) return rcf; }
    if ((rc = neighbour_pass(ctx, -1, false, ctx->step, ctx->step, false))) return rc;
    return export_pass(ctx, nullptr, turn, n_avoid, n_align, n_attract, nullptr);
}

int fishabm_get_speed(fishabm_ctx* ctx, double* out) {
    int rc = need_state(ctx);
    if (rc) return rc;
    if (!out) return fail(ctx, FISHABM_E_ARG, "null argument");
    CK(cudaSetDevice(ctx->p.device));
    if (ctx->ens) return ens_query(ctx, nullptr, nullptr, nullptr, nullptr, nullptr, out);
    { int rcf = flush_pending(ctx); if (rcf) return rcf; }
    if ((rc = neighbour_pass(ctx, -1, false, ctx->step, ctx->step, false))) return rc;
    return export_pass(ctx, nullptr, nullptr, nullptr, nullptr, nullptr, out);
}

int fishabm_in_zone(fishabm_ctx* ctx, int32_t fov_deg, double r, int32_t* out) {
    int rc = need_state(ctx);
    if (rc) return rc;
    if (!out) return fail(ctx, FISHABM_E_ARG, "null argument");
    if (!(r <= 200.0) || fov_deg < 0 || fov_deg > 360) return fail(ctx, FISHABM_E_ARG, "fishabm_in_zone: needs r <= 200 and 0 <= fov <= 360");
    if (ctx->ens) return fail(ctx, FISHABM_E_STATE, "fishabm_in_zone is not available for ensembles (use fishabm_zone_counts)");
    CK(cudaSetDevice(ctx->p.device));
    { int rcf = flush_pending(ctx); if (rcf) return rcf; }
    int cnt = 0;
    if ((rc = owned_range(ctx, &cnt))) return rc;
    const Buffers& b = ctx->buf[ctx->cur];
    ZoneArgs a{b.rec, b.cold0, b.cold1, b.cell, ctx->cell_start, ctx->st_lag, ctx->slab ? ctx->st_ids : nullptr, ctx->g, ctx->slab ? 1 : 0, fov_deg / 2, r};
    if (cnt > 0) in_zone_generic_kernel<<<cdiv(cnt, 128), 128, 0, ctx->stream>>>(a);
    ctx->launches += 1;
    CK(cudaGetLastError());
    if ((rc = fetch_ids(ctx, cnt))) return rc;
    if ((rc = fetch(ctx, out, ctx->st_lag, cnt, 1))) return rc;
    CK(cudaStreamSynchronize(ctx->stream));
    return 0;
}

int fishabm_check_fp64(fishabm_ctx* ctx, const int32_t* ids, int32_t k, int32_t* counts4, double* turn, int32_t* flags) {
    int rc = need_state(ctx);
    if (rc) return rc;
    if (!ids || k < 0 || !counts4 || !turn || !flags) return fail(ctx, FISHABM_E_ARG, "fishabm_check_fp64: null argument");
    if (ctx->slab || ctx->ens) return fail(ctx, FISHABM_E_STATE, "fishabm_check_fp64 needs a whole-school context (every fish on this device)");
    if (k > ctx->n) return fail(ctx, FISHABM_E_ARG, "fishabm_check_fp64: at most n ids per call");
    if (k == 0) return 0;
    CK(cudaSetDevice(ctx->p.device));
    { int rcf = flush_pending(ctx); if (rcf) return rcf; }
    const Buffers& b = ctx->buf[ctx->cur];
    cudaStream_t s = ctx->stream;
    CK(cudaMemcpyAsync(ctx->st_ids, ids, (size_t)k * sizeof(int), cudaMemcpyHostToDevice, s));
    CK(cudaMemsetAsync(ctx->pos_of_id, 0xFF, (size_t)ctx->n * sizeof(int), s));
    index_by_id_kernel<<<cdiv(ctx->cap, 256), 256, 0, s>>>(b.cold1, ctx->cell_start, ctx->g, ctx->pos_of_id);
    CheckArgs a{b.rec, b.cold0, b.cold1, b.cell, ctx->pos_of_id, ctx->st_ids, k, ctx->st_i4, ctx->st_dir, ctx->st_lag,
                ctx->n, ctx->g, ctx->p.w, ctx->p.h, ctx->step, ctx->p.seed, ctx->p.replicate0, ctx->model,
                ctx->p.model.weight_align, ctx->p.model.weight_attract};
    check_fp64_kernel<<<cdiv(k * 32, 256), 256, 0, s>>>(a);
    ctx->launches += 2;
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(counts4, ctx->st_i4, (size_t)k * 4 * sizeof(int), cudaMemcpyDeviceToHost, s));
    CK(cudaMemcpyAsync(turn, ctx->st_dir, (size_t)k * sizeof(double), cudaMemcpyDeviceToHost, s));
    CK(cudaMemcpyAsync(flags, ctx->st_lag, (size_t)k * sizeof(int), cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    return 0;
}

// ---- stepping --------------------------------------------------------------------------------------------
int fishabm_move(fishabm_ctx* ctx) {
    int rc = need_state(ctx);
    if (rc) return rc;
    CK(cudaSetDevice(ctx->p.device));
    if (ctx->ens) return ens_timed(ctx, ENS_OP_MOVE, 1);
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
    if (ctx->ens) { rc = ens_timed(ctx, ENS_OP_SAMPLE, 1); if (!rc) ctx->step += 1; return rc; }
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

// first half of one step: the neighbour pass and the per-fish update, up to the packed slab messages
static int step_first_half(fishabm_ctx* ctx) {
    int rc;
    const uint32_t s = ctx->step;
    const bool pend = ctx->pending_sample;  // sample(s-1) and move(s) share one neighbour pass
    if (can_fuse_update(ctx)) {              // ... and the per-fish update rides inside it
        UpdateArgs u;
        if ((rc = make_update_args(ctx, pend, true, s, &u))) return rc;
        u.rec_out = ctx->rec_tmp;
        if ((rc = neighbour_pass(ctx, pend ? 1 : 0, pend, pend ? s - 1 : s, s, true, &u))) return rc;
        ctx->rec_src = ctx->rec_tmp;
        if ((rc = phase_mark(ctx, 1))) return rc;
        return after_update(ctx, pend, true);
    }
    if ((rc = neighbour_pass(ctx, pend ? 1 : 0, pend, pend ? s - 1 : s, s, true))) return rc;
    if ((rc = phase_mark(ctx, 1))) return rc;
    return update_launch(ctx, pend, true, s);
}

static int step_untimed(fishabm_ctx* ctx, int nsteps) {
    // move(s); { sample(s+t-1) + move(s+t) } ...; the trailing sample stays pending (flush_pending)
    int rc;
    for (int t = 0; t < nsteps; ++t) {
        if ((rc = phase_mark(ctx, 0))) return rc;
        if ((rc = step_first_half(ctx))) return rc;
        if ((rc = phase_mark(ctx, 2))) return rc;
        unsigned char *il = nullptr, *ir = nullptr;
        if (ctx->slab && (rc = exchange(ctx, &il, &ir))) return rc;
        if ((rc = phase_mark(ctx, 3))) return rc;
        if ((rc = finish_move(ctx, il, ir))) return rc;
        if ((rc = phase_mark(ctx, 4))) return rc;
        ctx->pending_sample = true;
        ctx->step += 1;
    }
    return 0;
}

int fishabm_step(fishabm_ctx* ctx, int32_t nsteps) {
    int rc = need_state(ctx);
    if (rc) return rc;
    if (nsteps < 1) return fail(ctx, FISHABM_E_ARG, "nsteps must be >= 1");
    CK(cudaSetDevice(ctx->p.device));
    if (ctx->ens) { rc = ens_timed(ctx, ENS_OP_STEPS, nsteps); if (!rc) ctx->step += (uint32_t)nsteps; return rc; }
    if ((rc = begin_timed(ctx))) return rc;
    if ((rc = step_untimed(ctx, nsteps))) return rc;
    return end_timed(ctx);
}

int fishabm_flush(fishabm_ctx* ctx) {
    int rc = need_state(ctx);
    if (rc) return rc;
    if (ctx->ens) { ctx->last_ms = 0.f; ctx->nb_events_used = 0; return 0; }  // an ensemble step completes inside the call
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
    if (ctx->ens) {
        // the whole logged run is ONE launch: the loggers' rows are written from shared memory as the steps go by
        const int rows = (intake && intake_every > 0) ? (nsteps + intake_every - 1) / intake_every : 0;
        long long* d_sumq = nullptr;
        int* d_intake = nullptr;
        const size_t sq = (size_t)nsteps * ctx->R, si = (size_t)rows * ctx->R * ctx->n;
        if (sumq) CK(cudaMalloc(&d_sumq, sq * sizeof(long long)));
        if (rows) { cudaError_t e = cudaMalloc(&d_intake, si * sizeof(int)); if (e != cudaSuccess) { cudaFree(d_sumq); return fail(ctx, FISHABM_E_NOMEM, "intake log: %s", cudaGetErrorString(e)); } }
        EnsembleArgs a = ens_args(ctx, ENS_OP_STEPS, nsteps);
        a.sumq = d_sumq; a.intake = d_intake; a.intake_every = rows ? intake_every : 0;
        rc = begin_timed(ctx);
        if (!rc) rc = ens_launch(ctx, a);
        if (!rc) rc = end_timed(ctx);
        cudaError_t e1 = cudaSuccess, e2 = cudaSuccess;
        if (!rc && sumq) e1 = cudaMemcpy(sumq, d_sumq, sq * sizeof(long long), cudaMemcpyDeviceToHost);
        if (!rc && rows) e2 = cudaMemcpy(intake, d_intake, si * sizeof(int), cudaMemcpyDeviceToHost);
        cudaFree(d_sumq); cudaFree(d_intake);
        if (rc) return rc;
        if (e1 != cudaSuccess || e2 != cudaSuccess) return fail(ctx, FISHABM_E_CUDA, "copying the logs back: %s", cudaGetErrorString(e1 != cudaSuccess ? e1 : e2));
        ctx->step += (uint32_t)nsteps;
        return 0;
    }
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
    return check_device_errors(ctx);
}

int fishabm_feed(fishabm_ctx* ctx, const int32_t* ids, int32_t k) {
    int rc = need_state(ctx);
    if (rc) return rc;
    if (!ids || k < 0) return fail(ctx, FISHABM_E_ARG, "null argument");
    if (k > ctx->n) return fail(ctx, FISHABM_E_ARG, "fishabm_feed: at most n ids per call");
    if (ctx->ens) return fail(ctx, FISHABM_E_STATE, "fishabm_feed is not available for ensembles");
    if (k == 0) return 0;
    CK(cudaSetDevice(ctx->p.device));
    { int rcf = flush_pending(ctx); if (rcf) return rcf; }
    const Buffers& b = ctx->buf[ctx->cur];
    CK(cudaMemcpyAsync(ctx->st_lag, ids, (size_t)k * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemsetAsync(ctx->pos_of_id, 0xFF, (size_t)ctx->n * sizeof(int), ctx->stream));
    index_by_id_kernel<<<cdiv(ctx->cap, 256), 256, 0, ctx->stream>>>(b.cold1, ctx->cell_start, ctx->g, ctx->pos_of_id);
    FeedArgs a{b.cold0, b.cold1, b.rec, b.cell, ctx->st_lag, k, ctx->quality, ctx->n, ctx->qcols, ctx->g, ctx->pos_of_id, ctx->model};
    feed_list_kernel<<<1, 32, 0, ctx->stream>>>(a);
    ctx->launches += 2;
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(ctx->stream));
    return 0;
}

// ---- slab decomposition ------------------------------------------------------------------------------------
int fishabm_slab_get_info(fishabm_ctx* ctx, fishabm_slab_info* out) {
    if (!ctx || !out) return FISHABM_E_ARG;
    memset(out, 0, sizeof *out);
    out->rank = ctx->p.rank; out->nranks = ctx->p.nranks; out->left = ctx->left; out->right = ctx->right;
    out->col0 = ctx->col0; out->col1 = ctx->col0 + ctx->own;
    out->capacity = ctx->cap; out->halo_capacity = ctx->cap_g; out->migrant_capacity = ctx->cap_m;
    out->message_bytes = (int64_t)ctx->msg_bytes;
    if (ctx->have_state) {
        CK(cudaSetDevice(ctx->p.device));
        int own = 0, held = 0, rc;
        if ((rc = owned_range(ctx, &own))) return rc;
        CK(cudaMemcpyAsync(&held, ctx->cell_start + ctx->ncell, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
        out->owned = own; out->held = held;
    }
    return 0;
}

int fishabm_slab_unique_id(void* id_out, int32_t bytes) {
    fishabm_ctx* ctx = nullptr;
    if (!id_out || bytes < FISHABM_UNIQUE_ID_BYTES) return fail(nullptr, FISHABM_E_ARG, "fishabm_slab_unique_id: needs a %d-byte buffer", FISHABM_UNIQUE_ID_BYTES);
    NcclApi& N = nccl();
    if (!N.h) return fail(nullptr, FISHABM_E_COMM, "%s", N.why.c_str());
    NcclUniqueId id;
    NK(N.GetUniqueId(&id));
    memcpy(id_out, &id, sizeof id);
    return 0;
}

int fishabm_slab_connect(fishabm_ctx* ctx, const void* id, int32_t bytes) {
    if (!ctx || !id || bytes < FISHABM_UNIQUE_ID_BYTES) return fail(ctx, FISHABM_E_ARG, "fishabm_slab_connect: bad argument");
    if (!ctx->slab || ctx->p.nranks < 2) return fail(ctx, FISHABM_E_STATE, "fishabm_slab_connect: the context is not one of several slabs");
    if (ctx->comm) return fail(ctx, FISHABM_E_STATE, "fishabm_slab_connect: already connected");
    NcclApi& N = nccl();
    if (!N.h) return fail(ctx, FISHABM_E_COMM, "%s", N.why.c_str());
    CK(cudaSetDevice(ctx->p.device));
    NcclUniqueId uid;
    memcpy(&uid, id, sizeof uid);
    NK(N.CommInitRank(&ctx->comm, ctx->p.nranks, uid, ctx->p.rank));
    return 0;
}

int fishabm_slab_begin_step(fishabm_ctx* ctx) {
    int rc = need_state(ctx);
    if (rc) return rc;
    if (!ctx->slab) return fail(ctx, FISHABM_E_STATE, "not a slab context");
    CK(cudaSetDevice(ctx->p.device));
    ctx->nb_events_used = 0;
    if ((rc = step_first_half(ctx))) return rc;
    CK(cudaStreamSynchronize(ctx->stream));
    ctx->step_open = true;
    return 0;
}

int fishabm_slab_buffers(fishabm_ctx* ctx, void** out_left, void** out_right, void** in_left, void** in_right, int64_t* bytes) {
    if (!ctx || !ctx->slab) return fail(ctx, FISHABM_E_STATE, "not a slab context");
    if (out_left) *out_left = ctx->out_left;
    if (out_right) *out_right = ctx->out_right;
    if (in_left) *in_left = ctx->in_left;
    if (in_right) *in_right = ctx->in_right;
    if (bytes) *bytes = (int64_t)ctx->msg_bytes;
    return 0;
}

int fishabm_slab_end_step(fishabm_ctx* ctx) {
    if (!ctx || !ctx->slab || !ctx->step_open) return fail(ctx, FISHABM_E_STATE, "fishabm_slab_end_step without fishabm_slab_begin_step");
    CK(cudaSetDevice(ctx->p.device));
    int rc = finish_move(ctx, ctx->in_left, ctx->in_right);
    if (rc) return rc;
    ctx->step_open = false;
    ctx->pending_sample = true;
    ctx->step += 1;
    CK(cudaStreamSynchronize(ctx->stream));
    return check_device_errors(ctx);
}

int fishabm_slab_get_endpoint(fishabm_ctx* ctx, fishabm_slab_endpoint* out) {
    if (!ctx || !out) return FISHABM_E_ARG;
    if (!ctx->slab || !ctx->inbox) return fail(ctx, FISHABM_E_STATE, "not a slab context");
    CK(cudaSetDevice(ctx->p.device));
    memset(out, 0, sizeof *out);
    cudaIpcMemHandle_t h;
    static_assert(sizeof(h) <= sizeof(out->ipc_handle), "ipc handle size");
    CK(cudaIpcGetMemHandle(&h, ctx->inbox));
    memcpy(out->ipc_handle, &h, sizeof h);
    out->raw_pointer = (uint64_t)(uintptr_t)ctx->inbox;
    out->bytes = (int64_t)ctx->inbox_bytes;
    out->pid = (int32_t)getpid(); out->device = ctx->p.device; out->rank = ctx->p.rank;
    out->pad = gpu_identity(ctx->p.device);   // which physical GPU (hash of its PCI bus id): see fishabm_slab_attach
    return 0;
}

int fishabm_slab_attach(fishabm_ctx* ctx, const fishabm_slab_endpoint* left, const fishabm_slab_endpoint* right) {
    if (!ctx || !left || !right) return fail(ctx, FISHABM_E_ARG, "fishabm_slab_attach: null argument");
    if (!ctx->slab || ctx->p.nranks < 2) return fail(ctx, FISHABM_E_STATE, "fishabm_slab_attach: the context is not one of several slabs");
    if (ctx->p2p) return fail(ctx, FISHABM_E_STATE, "fishabm_slab_attach: already attached");
    if (left->rank != ctx->left || right->rank != ctx->right) return fail(ctx, FISHABM_E_ARG, "fishabm_slab_attach: endpoints of ranks %d / %d given, neighbours are %d / %d", left->rank, right->rank, ctx->left, ctx->right);
    if (left->bytes != (int64_t)ctx->inbox_bytes || right->bytes != (int64_t)ctx->inbox_bytes) return fail(ctx, FISHABM_E_ARG, "fishabm_slab_attach: the neighbours' message capacities differ from mine");
    CK(cudaSetDevice(ctx->p.device));
    // Kernels of ONE GPU cannot be relied on to run at the same time (one slab's wait kernel could keep the neighbour's
    // kernels from ever being scheduled), so the peer-memory transport needs every slab on its own GPU.  Slabs that share a
    // GPU exchange through the caller-driven hooks (fishabm_slab_begin_step / buffers / end_step).
    const int32_t my_gpu = gpu_identity(ctx->p.device);
    if (ctx->p.nranks > 1 && (left->pad == my_gpu || right->pad == my_gpu) && !getenv("FISHABM_ALLOW_SHARED_GPU_P2P"))
        return fail(ctx, FISHABM_E_ARG, "fishabm_slab_attach: a neighbour slab lives on the same GPU; the peer-memory transport needs one GPU per slab "
                                       "(use the caller-driven exchange: fishabm_slab_begin_step / fishabm_slab_buffers / fishabm_slab_end_step)");
    const fishabm_slab_endpoint* eps[2] = {left, right};
    unsigned char* ptr[2] = {nullptr, nullptr};
    for (int k = 0; k < 2; ++k) {
        if (k == 1 && left->pid == right->pid && left->raw_pointer == right->raw_pointer) { ptr[1] = ptr[0]; break; }  // two ranks: one peer
        const fishabm_slab_endpoint* e = eps[k];
        if (e->pid == (int32_t)getpid()) {
            if (e->device != ctx->p.device) {
                cudaError_t r = cudaDeviceEnablePeerAccess(e->device, 0);
                if (r == cudaErrorPeerAccessAlreadyEnabled) cudaGetLastError();
                else if (r != cudaSuccess) return fail(ctx, FISHABM_E_COMM, "no peer access from device %d to %d: %s", ctx->p.device, e->device, cudaGetErrorString(r));
            }
            ptr[k] = reinterpret_cast<unsigned char*>((uintptr_t)e->raw_pointer);
        } else {
            cudaIpcMemHandle_t h;
            memcpy(&h, e->ipc_handle, sizeof h);
            void* q = nullptr;
            cudaError_t r = cudaIpcOpenMemHandle(&q, h, cudaIpcMemLazyEnablePeerAccess);
            if (r != cudaSuccess) return fail(ctx, FISHABM_E_COMM, "cudaIpcOpenMemHandle (rank %d's inbox): %s", e->rank, cudaGetErrorString(r));
            ctx->ipc_opened[k] = q;
            ptr[k] = static_cast<unsigned char*>(q);
        }
    }
    ctx->peer_left = ptr[0]; ctx->peer_right = ptr[1];
    ctx->p2p = true;
    return 0;
}

int fishabm_step_enqueue(fishabm_ctx* ctx, int32_t nsteps) {
    int rc = need_state(ctx);
    if (rc) return rc;
    if (nsteps < 1) return fail(ctx, FISHABM_E_ARG, "nsteps must be >= 1");
    CK(cudaSetDevice(ctx->p.device));
    if (ctx->ens) return fail(ctx, FISHABM_E_STATE, "fishabm_step_enqueue is for single schools / slabs");
    if ((rc = begin_timed(ctx))) return rc;
    return step_untimed(ctx, nsteps);
}

int fishabm_wait(fishabm_ctx* ctx) {
    if (!ctx) return FISHABM_E_ARG;
    CK(cudaSetDevice(ctx->p.device));
    return end_timed(ctx);
}

int fishabm_slab_copy(void* dst, const void* src, int64_t bytes) {
    if (!dst || !src || bytes < 0) return FISHABM_E_ARG;
    if (cudaMemcpy(dst, src, (size_t)bytes, cudaMemcpyDefault) != cudaSuccess) return FISHABM_E_CUDA;
    if (cudaDeviceSynchronize() != cudaSuccess) return FISHABM_E_CUDA;  // device-to-device copies return early
    return 0;
}

int fishabm_set_replicate_models(fishabm_ctx* ctx, const fishabm_model_params* models, int32_t count) {
    if (!ctx || !models) return fail(ctx, FISHABM_E_ARG, "fishabm_set_replicate_models: null argument");
    if (count != ctx->R) return fail(ctx, FISHABM_E_ARG, "fishabm_set_replicate_models: %d models for %d replicates", count, ctx->R);
    if (ctx->step_open) return fail(ctx, FISHABM_E_STATE, "a slab step is open: call fishabm_slab_end_step first");
    std::vector<Model> hm;
    try { hm.resize((size_t)count); } catch (...) { return fail(ctx, FISHABM_E_NOMEM, "out of host memory"); }
    for (int r = 0; r < count; ++r) {
        if (const char* why = model_invalid(models[r])) return fail(ctx, FISHABM_E_ARG, "fishabm_set_replicate_models: model %d: %s", r, why);
        hm[(size_t)r] = to_device_model(models[r]);
    }
    CK(cudaSetDevice(ctx->p.device));
    if (ctx->have_state) { int rc = flush_pending(ctx); if (rc) return rc; }   // the pending sample() belongs to the old constants
    CK(cudaMemcpyAsync(ctx->d_models, hm.data(), hm.size() * sizeof(Model), cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    ctx->model = hm[0];
    ctx->p.model = models[0];
    return 0;
}

int fishabm_get_step(const fishabm_ctx* ctx, uint32_t* step) {
    if (!ctx || !step) return FISHABM_E_ARG;
    *step = ctx->step;
    return 0;
}
int fishabm_set_step(fishabm_ctx* ctx, uint32_t step) {
    if (!ctx) return FISHABM_E_ARG;
    if (ctx->step_open) return fail(ctx, FISHABM_E_STATE, "a slab step is open: call fishabm_slab_end_step first");
    // a pending sample() belongs to the OLD step index (its draws are keyed by it): complete it before renumbering
    if (ctx->have_state && ctx->pending_sample) {
        CK(cudaSetDevice(ctx->p.device));
        int rc = flush_pending(ctx);
        if (rc) return rc;
    }
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

int fishabm_enable_phase_timing(fishabm_ctx* ctx, int32_t on) {
    if (!ctx) return FISHABM_E_ARG;
    ctx->phases_on = on != 0;
    ctx->ph_steps = 0;
    return 0;
}

int fishabm_last_phase_ms(const fishabm_ctx* ctx, float* ms4, int32_t* steps) {
    if (!ctx || !ms4) return FISHABM_E_ARG;
    for (int k = 0; k < 4; ++k) ms4[k] = 0.f;
    for (int t = 0; t < ctx->ph_steps; ++t)
        for (int k = 0; k < 4; ++k) {
            float v = 0.f;
            if (cudaEventElapsedTime(&v, ctx->ph_events[(size_t)t * PH_PER_STEP + k], ctx->ph_events[(size_t)t * PH_PER_STEP + k + 1]) == cudaSuccess) ms4[k] += v;
        }
    if (steps) *steps = ctx->ph_steps;
    return 0;
}

int fishabm_enable_stats(fishabm_ctx* ctx, int32_t on) {
    if (!ctx) return FISHABM_E_ARG;
    ctx->stats_on = on != 0;
    return 0;
}

int fishabm_last_pass_stats(fishabm_ctx* ctx, fishabm_pass_stats* out) {
    if (!ctx || !out) return FISHABM_E_ARG;
    unsigned long long v[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    CK(cudaSetDevice(ctx->p.device));
    CK(cudaMemcpyAsync(v, ctx->stats, sizeof v, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    out->focal = v[0]; out->candidates = v[1]; out->interacting = v[2]; out->rechecked = v[3];
    if (getenv("FISHABM_TAIL_DEBUG") && v[7]) {   // developer aid: how much of the neighbour launch its warps spent waiting for the last one
        const double span = (double)(v[4] - ~v[6]), busy = (double)v[5] / (double)v[7];
        fprintf(stderr, "fishabm: neighbour launch %.1f us from first warp start to last warp exit, mean warp busy %.1f us (%.1f %% idle), %llu warps\n",
                span * 1e-3, busy * 1e-3, 100.0 * (1.0 - busy / span), v[7]);
    }
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
