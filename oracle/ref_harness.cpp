// oracle/ref_harness.cpp -- TEST INFRASTRUCTURE (parity oracle), not product code.
//
// Wraps the UNMODIFIED reference translation unit (fish_mod.cpp) so that its own functions
// can be called from tests on frozen state with reproducible random draws.  The reference is
// not copied into this repository: oracle/build_ref.sh pipes /root/reference/fish_mod.cpp
// through `sed` (only line 23, `#define num 60`, and optionally the grid stride on lines
// 24-25, are rewritten) into a temporary file, names it in REF_TU and deletes it after
// compiling.  Only the resulting libfishref_<num>.so lands in oracle/_ref/.
//
// How the reference is made replayable:
//   * `random_device` and `uniform_int_distribution` are macro-swapped, for the duration of
//     the #include only, with `harness_urbg` / `harness_dist`.  harness_dist knows its (a,b)
//     range, picks the stream from it and serves Philox words addressed by
//     (seed, replicate, step, fish id, stream, k) through libstdc++'s own integer mapping
//     (orc_rng.h), so the reference, the C restatement and the CUDA path can be fed identical
//     draws.
//   * `main` is renamed; state lives in static storage; every call runs on a thread with a
//     stack sized for the reference's by-value `individuals` copies (fish_mod.cpp:64,79).
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <fstream>
#include <iostream>
#include <map>
#include <random>
#include <ratio>
#include <sstream>
#include <string>
#include <thread>
#include <vector>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <pthread.h>
#include <unistd.h>

#include "orc_rng.h"

// ----------------------------------------------------------------------------------------
// replayable RNG
// ----------------------------------------------------------------------------------------
enum HarnessMode {
    HM_FOCAL = 0,  // every draw belongs to fish `focal`
    HM_MOVE = 1,   // inside reference move(): error draws arrive once per fish in id order
    HM_SAMPLE = 2, // inside reference sample(): sample draws arrive for lag==0 fish in id order
    HM_INIT = 3,   // inside reference initialize(): randomwalk draws arrive once per fish
    HM_ZERO = 4,   // every draw returns 0 (or `a` if 0 is outside [a,b])
    HM_ENV = 5     // inside reference create_environment(): distrib_seed returns env_seeds; every other draw is the
                   // k-th of its stream (k = 0,1,2,... in call order), fish id 0
};

struct HarnessCtx {
    uint64_t seed = 0;
    uint32_t replicate = 0;
    uint32_t step = 0;
    int mode = HM_FOCAL;
    int focal = 0;
    int zero_error_except = -1; // HM_MOVE: if >=0, error draws of every other fish return 0
    long n_error = 0, n_sample = 0, n_rw = 0, n_choice = 0;
    long choice_k = 0;               // tie-break draws consumed by the current focal fish
    int social_first = 1;            // evaluation order of `social(...) + distrib_error(rd)`
    const int* lag0_ids = nullptr;   // HM_SAMPLE: ids with lag==0 in id order
    long n_lag0 = 0;
    long census[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    int env_seeds = 0;               // HM_ENV: the value distrib_seed yields (the shipped (0,0) range can only give 0)
    long env_k[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    int probe_seq[4] = {-1, -1, -1, -1}; // first streams drawn since probe_n was reset
    int probe_n = 0;
};
static thread_local HarnessCtx g_ctx;

static int stream_of(int a, int b) {
    if (a == 0 && b == 1) return ORC_STREAM_CHOICE;
    if (a == -10 && b == 10) return ORC_STREAM_ERROR;
    if (a == -45 && b == 45) return ORC_STREAM_RANDOMWALK;
    if (a == 0 && b == 10) return ORC_STREAM_QUALITY;
    if (a == 0 && b == 0) return ORC_STREAM_SEED;
    if (a == 0 && b == 35) return ORC_STREAM_SAMPLE;
    return 7;
}

static int harness_draw(int a, int b) {
    HarnessCtx& c = g_ctx;
    const int stream = stream_of(a, b);
    c.census[stream & 7]++;
    if (c.probe_n < 4) c.probe_seq[c.probe_n++] = stream;
    if (c.mode == HM_ZERO) return (a <= 0 && 0 <= b) ? 0 : a;
    if (c.mode == HM_ENV) {
        if (stream == ORC_STREAM_SEED) return c.env_seeds;
        return orc_draw_from(c.seed, c.replicate, 0, 0, (uint32_t)stream, (uint32_t)c.env_k[stream & 7]++, a, b);
    }
    int id = c.focal;
    uint32_t k0 = 0;
    switch (stream) {
    case ORC_STREAM_ERROR:
        if (c.mode == HM_MOVE) {
            id = (int)c.n_error;
            c.n_error++;
            c.choice_k = 0;
            if (c.zero_error_except >= 0 && id != c.zero_error_except) return 0;
        }
        break;
    case ORC_STREAM_SAMPLE:
        if (c.mode == HM_SAMPLE) {
            id = (c.n_sample < c.n_lag0) ? c.lag0_ids[c.n_sample] : -1;
            c.n_sample++;
        }
        break;
    case ORC_STREAM_RANDOMWALK:
        if (c.mode == HM_INIT) { id = (int)c.n_rw; }
        else if (c.mode == HM_MOVE) { id = (int)c.n_error - (c.social_first ? 0 : 1); }
        c.n_rw++;
        break;
    case ORC_STREAM_CHOICE:
        if (c.mode == HM_MOVE) id = (int)c.n_error - (c.social_first ? 0 : 1);
        k0 = (uint32_t)c.choice_k++; // reference order = candidate order; the restatement keys ties by neighbour id
        c.n_choice++;
        break;
    default: break;
    }
    return orc_draw_from(c.seed, c.replicate, c.step, (uint32_t)id, (uint32_t)stream, k0, a, b);
}

struct harness_urbg {
    typedef uint32_t result_type;
    static constexpr result_type min() { return 0u; }
    static constexpr result_type max() { return 0xFFFFFFFFu; }
    result_type operator()() { return 0u; } // never consulted: harness_dist addresses words itself
};

template <class T> struct harness_dist {
    T lo, hi;
    harness_dist(T a, T b) : lo(a), hi(b) {}
    template <class G> T operator()(G&) { return (T)harness_draw((int)lo, (int)hi); }
};

// ----------------------------------------------------------------------------------------
// the reference, verbatim
// ----------------------------------------------------------------------------------------
#define main ref_main
#define random_device harness_urbg
#define uniform_int_distribution harness_dist
#include REF_TU
#undef main
#undef random_device
#undef uniform_int_distribution

static const int kRefNum = num;
static const int kRefCols = n_cols;
static const int kRefRows = n_rows;
#undef num // keep the macro out of the rest of this file

// ----------------------------------------------------------------------------------------
// static state + big-stack runner
// ----------------------------------------------------------------------------------------
static individuals g_inds;
static Structure g_struc;
static int g_quality[n_rows][n_cols];
static Mat g_env;

namespace {
struct Job { void (*fn)(void*); void* arg; HarnessCtx ctx; };
void* job_trampoline(void* p) {
    Job* j = (Job*)p;
    g_ctx = j->ctx;
    j->fn(j->arg);
    j->ctx = g_ctx;
    return nullptr;
}
size_t big_stack_bytes() { return (size_t)8 * sizeof(individuals) + ((size_t)16 << 20); }

// run fn(arg) on a thread whose stack can hold the reference's by-value copies; the calling
// thread's harness context is carried in and out.
void run_big(void (*fn)(void*), void* arg) {
    Job j{fn, arg, g_ctx};
    pthread_attr_t at;
    pthread_attr_init(&at);
    pthread_attr_setstacksize(&at, big_stack_bytes());
    pthread_t th;
    if (pthread_create(&th, &at, job_trampoline, &j) != 0) {
        fprintf(stderr, "ref_harness: cannot create a %zu-byte stack thread\n", big_stack_bytes());
        abort();
    }
    pthread_join(th, nullptr);
    pthread_attr_destroy(&at);
    g_ctx = j.ctx;
}

// parallel for over [0,n) on `threads` big-stack threads (static interleaved partition)
struct ParJob { void (*body)(int, void*); void* arg; int n, t, nt; HarnessCtx ctx; };
void* par_trampoline(void* p) {
    ParJob* j = (ParJob*)p;
    g_ctx = j->ctx;
    for (int i = j->t; i < j->n; i += j->nt) j->body(i, j->arg);
    return nullptr;
}
void par_for(int n, int threads, void (*body)(int, void*), void* arg) {
    if (threads < 1) threads = 1;
    if (threads > n) threads = n > 0 ? n : 1;
    std::vector<ParJob> jobs(threads);
    std::vector<pthread_t> th(threads);
    pthread_attr_t at;
    pthread_attr_init(&at);
    pthread_attr_setstacksize(&at, big_stack_bytes());
    for (int t = 0; t < threads; ++t) {
        jobs[t] = ParJob{body, arg, n, t, threads, g_ctx};
        if (pthread_create(&th[t], &at, par_trampoline, &jobs[t]) != 0) abort();
    }
    for (int t = 0; t < threads; ++t) pthread_join(th[t], nullptr);
    pthread_attr_destroy(&at);
}
double now_s() {
    return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}
} // namespace

// ----------------------------------------------------------------------------------------
// C entry points (bound with ctypes from tests/ and bench.py's cpu_baseline leg)
// ----------------------------------------------------------------------------------------
extern "C" {

int ref_num(void) { return kRefNum; }
int ref_grid_stride(void) { return kRefCols; }
int ref_grid_rows(void) { return kRefRows; }
long ref_sizeof_individuals(void) { return (long)sizeof(individuals); }

// world geometry: Structure's members are public (fish_mod.cpp:51-62), so other window sizes
// are set directly; w,h,rows,cols keep the reference's int types.
void ref_set_world(int w, int h, int rows, int cols) {
    g_struc.w = w; g_struc.h = h; g_struc.rows = rows; g_struc.cols = cols;
}
void ref_get_world(int* out4) { out4[0] = g_struc.w; out4[1] = g_struc.h; out4[2] = g_struc.rows; out4[3] = g_struc.cols; }

void ref_rng(uint64_t seed, uint32_t replicate, uint32_t step) {
    g_ctx.seed = seed; g_ctx.replicate = replicate; g_ctx.step = step;
}
void ref_census(long* out8) { for (int i = 0; i < 8; ++i) out8[i] = g_ctx.census[i]; }
void ref_census_reset(void) { for (int i = 0; i < 8; ++i) g_ctx.census[i] = 0; }

int ref_set_state(int n, const double* coords, const double* dir, const double* speed,
                  const double* speed_factor, const int* lag, const int* sample_rate, const int* food_intake) {
    if (n > kRefNum) return -1;
    g_inds.n = n; g_inds.dim = 2; g_inds.nfollow = 0;
    for (int i = 0; i < n; ++i) {
        g_inds.coords[i][0] = coords[2 * i]; g_inds.coords[i][1] = coords[2 * i + 1];
        g_inds.dir[i] = dir[i]; g_inds.speed[i] = speed[i]; g_inds.speed_factor[i] = speed_factor[i];
        g_inds.lag[i] = lag[i]; g_inds.sample_rate[i] = sample_rate[i]; g_inds.food_intake[i] = food_intake[i];
        g_inds.sample[i] = false; g_inds.colors[i][0] = g_inds.colors[i][1] = g_inds.colors[i][2] = 0;
    }
    return 0;
}
void ref_get_state(double* coords, double* dir, double* speed, double* speed_factor, int* lag,
                   int* sample_rate, int* food_intake) {
    for (int i = 0; i < g_inds.n; ++i) {
        coords[2 * i] = g_inds.coords[i][0]; coords[2 * i + 1] = g_inds.coords[i][1];
        dir[i] = g_inds.dir[i]; speed[i] = g_inds.speed[i]; speed_factor[i] = g_inds.speed_factor[i];
        lag[i] = g_inds.lag[i]; sample_rate[i] = g_inds.sample_rate[i]; food_intake[i] = g_inds.food_intake[i];
    }
}
int ref_n(void) { return g_inds.n; }

int ref_set_quality(const int* q, int rows, int cols) {
    if (rows > kRefRows || cols > kRefCols) return -1;
    for (int r = 0; r < rows; ++r) for (int c = 0; c < cols; ++c) g_quality[r][c] = q[(size_t)r * cols + c];
    return 0;
}
void ref_get_quality(int* q, int rows, int cols) {
    for (int r = 0; r < rows; ++r) for (int c = 0; c < cols; ++c) q[(size_t)r * cols + c] = g_quality[r][c];
}

// get_environment (fish_mod.cpp:603-637) reads ./quality.csv relative to the CWD
static void job_get_env(void*) { get_environment(g_env, g_quality, g_struc); }
int ref_get_environment(const char* dir_with_quality_csv) {
    char old[4096];
    if (!getcwd(old, sizeof old)) return -1;
    if (chdir(dir_with_quality_csv) != 0) return -2;
    memset(g_quality, 0, sizeof g_quality);
    run_big(job_get_env, nullptr);
    if (chdir(old) != 0) return -3;
    return 0;
}

// create_environment (fish_mod.cpp:482-531): dead code in the reference's main (distrib_seed(0,0) yields 0 seeds);
// replayed here with a live seed count
static void job_create_env(void*) { create_environment(g_env, g_quality, g_struc); }
int ref_create_environment(int n_seeds, uint64_t seed, uint32_t replicate) {
    memset(g_quality, 0, sizeof g_quality);
    g_ctx = HarnessCtx();
    g_ctx.seed = seed; g_ctx.replicate = replicate; g_ctx.mode = HM_ENV; g_ctx.env_seeds = n_seeds;
    std::ostringstream sink;  // the reference prints progress lines
    std::streambuf* old = std::cout.rdbuf(sink.rdbuf());
    run_big(job_create_env, nullptr);
    std::cout.rdbuf(old);
    g_ctx.mode = HM_FOCAL;
    return 0;
}

static void job_sum(void* out) { *(int*)out = sum_array(g_quality, g_struc); }
int ref_sum_array(void) { int s = 0; run_big(job_sum, &s); return s; }

// initialize (fish_mod.cpp:183-206): positions from rand(), headings from distrib_randomwalk
struct InitArg { int n; double speed; unsigned srand_seed; };
static void job_init(void* p) {
    InitArg* a = (InitArg*)p;
    srand(a->srand_seed);
    g_ctx.mode = HM_INIT; g_ctx.n_rw = 0;
    initialize(g_inds, a->n, 2, a->speed, g_struc);
    g_ctx.mode = HM_FOCAL;
}
int ref_initialize(int n, double speed, unsigned srand_seed) {
    if (n > kRefNum) return -1;
    InitArg a{n, speed, srand_seed};
    run_big(job_init, &a);
    return 0;
}

// ---- per-fish known answers on the frozen state -------------------------------------------
struct ZoneArg { int fov; double r; int* out; };
static void body_in_zone(int id, void* p) { ZoneArg* a = (ZoneArg*)p; a->out[id] = in_zone(g_inds, id, a->fov, a->r, g_struc); }
void ref_in_zone_all(int fov, double r, int* out, int threads) {
    ZoneArg a{fov, r, out};
    par_for(g_inds.n, threads, body_in_zone, &a);
}
struct ZoneIdsArg { int fov; double r; const int* ids; int* out; };
static void body_in_zone_ids(int k, void* p) { ZoneIdsArg* a = (ZoneIdsArg*)p; a->out[k] = in_zone(g_inds, a->ids[k], a->fov, a->r, g_struc); }
void ref_in_zone_ids(int fov, double r, const int* ids, int k, int* out, int threads) {
    ZoneIdsArg a{fov, r, ids, out};
    par_for(k, threads, body_in_zone_ids, &a);
}

struct SocialArg { int fov; const int* ids; double* out; };
static void body_social(int k, void* p) {
    SocialArg* a = (SocialArg*)p;
    int id = a->ids ? a->ids[k] : k;
    g_ctx.mode = HM_FOCAL; g_ctx.focal = id; g_ctx.choice_k = 0;
    a->out[k] = social(g_inds, id, a->fov, g_struc);
}
void ref_social_all(int fov, double* out, int threads) {
    SocialArg a{fov, nullptr, out};
    par_for(g_inds.n, threads, body_social, &a);
}
void ref_social_ids(int fov, const int* ids, int k, double* out, int threads) {
    SocialArg a{fov, ids, out};
    par_for(k, threads, body_social, &a);
}

double ref_correctangle(double a) { return correctangle(a); }
void ref_correct_coords(double* xy) { correct_coords(xy, g_struc); }
double ref_getangle(const double* focal_xy, double focal_dir, const double* pos_xy, int focal_id) {
    double f[2] = {focal_xy[0], focal_xy[1]}, p[2] = {pos_xy[0], pos_xy[1]};
    g_ctx.mode = HM_FOCAL; g_ctx.focal = focal_id; g_ctx.choice_k = 0;
    return getangle(f, focal_dir, p, g_struc);
}
double ref_averageturn(const double* v, int n, int comb) {
    std::vector<double> vec(v, v + n);
    return averageturn(vec, comb != 0);
}

// get_speed (fish_mod.cpp:588-591) for every fish on the frozen state: each result is computed
// from the same unmodified snapshot (get_speed only writes speed_factor[id], which in_zone ignores).
static void body_get_speed(int id, void* out) {
    int count = in_zone(g_inds, id, 180, 200, g_struc);
    ((double*)out)[id] = 1 + 2 * (double)count / (double)kRefNum; // fish_mod.cpp:590 (normaliser is the macro)
}
void ref_get_speed_all(double* out, int threads) { par_for(g_inds.n, threads, body_get_speed, out); }
static void job_get_speed(void* p) { get_speed(g_inds, *(int*)p, g_struc); }
void ref_get_speed(int id) { run_big(job_get_speed, &id); } // verbatim, mutates speed_factor[id]

static void job_feed(void* p) { feed(g_inds, *(int*)p, g_quality, g_env, g_struc); }
void ref_feed(int id) { run_big(job_feed, &id); }

// ---- the reference's own sequential step ---------------------------------------------------
static void detect_eval_order(void);
static int g_order_known = 0, g_social_first = 1;

static void job_move(void*) {
    g_ctx.mode = HM_MOVE; g_ctx.n_error = 0; g_ctx.choice_k = 0; g_ctx.social_first = g_social_first;
    move(g_inds, g_struc);
    g_ctx.mode = HM_FOCAL;
}
void ref_move(void) {
    if (!g_order_known) detect_eval_order();
    g_ctx.zero_error_except = -1;
    run_big(job_move, nullptr);
}

static std::vector<int> g_lag0;
static void job_sample(void*) {
    g_lag0.clear();
    for (int i = 0; i < g_inds.n; ++i) if (g_inds.lag[i] == 0) g_lag0.push_back(i);
    g_ctx.mode = HM_SAMPLE; g_ctx.n_sample = 0; g_ctx.lag0_ids = g_lag0.data(); g_ctx.n_lag0 = (long)g_lag0.size();
    sample(g_inds, g_quality, g_env, g_struc);
    g_ctx.mode = HM_FOCAL; g_ctx.lag0_ids = nullptr;
}
void ref_sample(void) { run_big(job_sample, nullptr); }

// `steps` reference steps, move() then sample() (fish_mod.cpp:130-131), draws addressed by
// step = first_step, first_step+1, ...; optionally logs sum_array after each step (the
// commented-out logger, fish_mod.cpp:149-154) and food_intake every `intake_every` steps (:132-141).
struct RunArg { int steps; uint32_t first_step; int* sumq; int* intake; int intake_every; double seconds; };
static void job_run(void* p) {
    RunArg* a = (RunArg*)p;
    double t0 = now_s();
    int rows_logged = 0;
    for (int z = 0; z < a->steps; ++z) {
        g_ctx.step = a->first_step + (uint32_t)z;
        job_move(nullptr);
        job_sample(nullptr);
        if (a->intake && a->intake_every > 0 && z % a->intake_every == 0) {
            for (int i = 0; i < g_inds.n; ++i) a->intake[(size_t)rows_logged * g_inds.n + i] = g_inds.food_intake[i];
            rows_logged++;
        }
        if (a->sumq) a->sumq[z] = sum_array(g_quality, g_struc);
    }
    a->seconds = now_s() - t0;
}
double ref_run(int steps, uint32_t first_step, int* sumq_or_null, int* intake_or_null, int intake_every) {
    if (!g_order_known) detect_eval_order();
    g_ctx.zero_error_except = -1;
    RunArg a{steps, first_step, sumq_or_null, intake_or_null, intake_every, 0.0};
    run_big(job_run, &a);
    return a.seconds;
}

// Synchronous (frozen-state) move of ONE fish through the verbatim move(): every other fish is
// given lag=1 on a scratch copy and a zero error draw, so only fish `id` changes
// (SURVEY.md section 8c).  Outputs the focal fish's new dir, coords and speed_factor; the latter
// follows the reference's "new position, old heading" pose (fish_mod.cpp:171-175).
struct FocalArg { int id; double* out5; };
static void job_move_focal(void* p) {
    FocalArg* a = (FocalArg*)p;
    static individuals saved;
    memcpy(&saved, &g_inds, sizeof saved);
    int keep = g_inds.lag[a->id];
    for (int i = 0; i < g_inds.n; ++i) g_inds.lag[i] = 1;
    g_inds.lag[a->id] = keep;
    g_ctx.mode = HM_MOVE; g_ctx.n_error = 0; g_ctx.choice_k = 0; g_ctx.social_first = g_social_first;
    g_ctx.zero_error_except = a->id;
    move(g_inds, g_struc);
    g_ctx.mode = HM_FOCAL; g_ctx.zero_error_except = -1;
    a->out5[0] = g_inds.dir[a->id]; a->out5[1] = g_inds.coords[a->id][0]; a->out5[2] = g_inds.coords[a->id][1];
    a->out5[3] = g_inds.speed_factor[a->id]; a->out5[4] = (double)g_inds.lag[a->id];
    memcpy(&g_inds, &saved, sizeof saved);
}
void ref_move_focal(int id, double* out5) {
    if (!g_order_known) detect_eval_order();
    FocalArg a{id, out5};
    run_big(job_move_focal, &a);
}

// ---- timing legs for bench.py's cpu_baseline / --impl reference ------------------------------
// (a) whole steps: ref_run above returns seconds.
// (b) N too large for an O(N^2) step: the reference's own per-fish work of one step for `k` focal ids on the
//     full state: if lag==0, social (3 in_zone + pair loop) + get_speed's in_zone; then sample's 3 in_zone:
//     8 all-pairs passes for a moving fish, 3 for a lagged one (fish_mod.cpp:165-179,543-545).
struct FocalTimeArg { const int* ids; volatile double sink; };
static void body_focal_work(int k, void* p) {
    FocalTimeArg* a = (FocalTimeArg*)p;
    int id = a->ids[k];
    g_ctx.mode = HM_FOCAL; g_ctx.focal = id; g_ctx.choice_k = 0;
    double acc = 0;
    if (g_inds.lag[id] == 0) {                       // move(): fish_mod.cpp:165-175
        acc += social(g_inds, id, 330, g_struc);     //   social = 3 in_zone + the pair loop
        acc += in_zone(g_inds, id, 180, 200, g_struc); // get_speed
    }
    acc += in_zone(g_inds, id, 330, 15, g_struc);    // sample(): fish_mod.cpp:543-545, every fish
    acc += in_zone(g_inds, id, 330, 100, g_struc);
    acc += in_zone(g_inds, id, 330, 200, g_struc);
    a->sink = a->sink + acc;
}
double ref_time_focal(const int* ids, int k, int threads) {
    FocalTimeArg a{ids, 0.0};
    double t0 = now_s();
    par_for(k, threads, body_focal_work, &a);
    return now_s() - t0;
}

int ref_social_evaluated_first(void) { if (!g_order_known) detect_eval_order(); return g_social_first; }

} // extern "C"

// Which operand of `social(inds,id,330,struc) + distrib_error(rd)` (fish_mod.cpp:167) did this
// compiler evaluate first?  (Unspecified in C++11.)  Probe with a lone fish: social() falls through
// to distrib_randomwalk (fish_mod.cpp:372-374), so the order of the first two draws tells.
static void job_probe(void*) {
    g_ctx.mode = HM_ZERO;
    g_ctx.probe_n = 0;
    move(g_inds, g_struc);
}
static void detect_eval_order(void) {
    static individuals saved;
    memcpy(&saved, &g_inds, sizeof saved);
    HarnessCtx keep = g_ctx;
    g_inds.n = 1; g_inds.dim = 2; g_inds.lag[0] = 0; g_inds.dir[0] = 0; g_inds.speed[0] = 0;
    g_inds.speed_factor[0] = 1; g_inds.coords[0][0] = 10; g_inds.coords[0][1] = 10;
    run_big(job_probe, nullptr);
    g_social_first = (g_ctx.probe_seq[0] == ORC_STREAM_RANDOMWALK) ? 1 : 0;
    g_order_known = 1;
    memcpy(&g_inds, &saved, sizeof saved);
    g_ctx = keep;
}
