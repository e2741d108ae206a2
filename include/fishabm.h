/* include/fishabm.h -- C ABI of libfishabm.so, the B200 (sm_100a) implementation of the per-step agent
 * update of fritzfrancisco/fish_abm.
 *
 * The reference has no FFI layer: its operator API is the set of free-function prototypes at
 * fish_mod.cpp:64-82 over `individuals&`, `Structure` and `int quality[][n_cols]`.  Every entry point
 * below names the reference function it replaces.  Conventions:
 *   - extern "C", opaque context, plain pointers and sizes, no C++/torch types;
 *   - every call returns 0 on success or a negative FISHABM_E_* code and never throws;
 *     fishabm_last_error() gives the message;
 *   - the library owns all device memory; host buffers are borrowed for the duration of the call and use
 *     the reference's own layout (class individuals, fish_mod.cpp:35-49): double coords[n][2], double dir[n]
 *     (degrees, unwrapped), double speed[n], double speed_factor[n], int lag[n], int sample_rate[n],
 *     int food_intake[n];
 *   - one context = one school (or one ensemble of independent replicate schools, or one slab of a
 *     multi-GPU school).  A context is not thread-safe; distinct contexts are independent;
 *   - calls are synchronous with respect to the host;
 *   - there is no CPU fallback: without a CUDA device fishabm_create fails with FISHABM_E_CUDA.
 *
 * Model constants.  Geometry is fixed: zone radii 15/100/200 (fish_mod.cpp:273-275,543-545; 200 is the neighbour cell),
 * social field of view 330 deg and front cone 180 deg (:167,589), food cell 20 units (:56), tie draw choice(0,1) (:28).
 * Everything else the reference writes as a literal is a run-time parameter with the reference's value as default
 * (fishabm_model_params): lag 10 (:548), sample_rate 30/0 (:570,573), distributions error(-10,10) randomwalk(-45,45)
 * sample(0,35) (:29-33), the "alone" threshold 5 (:546), fed speed_factor 0.5 (:566), combination weights 1.7/0.7 (:385-386).
 */
#ifndef FISHABM_H
#define FISHABM_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FISHABM_VERSION 1

enum {
    FISHABM_OK = 0,
    FISHABM_E_ARG = -1,     /* bad argument / unsupported geometry */
    FISHABM_E_CUDA = -2,    /* CUDA runtime error (message has the detail) */
    FISHABM_E_STATE = -3,   /* call not valid in the current state (e.g. no state uploaded) */
    FISHABM_E_NOMEM = -4,
    FISHABM_E_COMM = -5     /* slab exchange error */
};

enum {
    FISHABM_MODE_CELL_FP32 = 0,  /* product path: cell list + fused fp32 neighbour kernel with fp64 re-check
                                    of borderline predicates */
    FISHABM_MODE_BRUTE_FP64 = 1, /* fp64 check mode as a stepping mode: every neighbour pass is the brute-force, double-precision
                                    evaluation of fishabm_check_fp64 over ALL pairs (no cell list, no fp32 pair arithmetic);
                                    O(n^2), one whole school of n <= 65536 fish.  State and move arithmetic stay fp32. */
    FISHABM_MODE_ENSEMBLE = 2    /* replicate ensemble: one CTA per school, all pairs in shared memory; implied by
                                    replicates > 1, may be forced for a single school (n <= 1024) */
};

/* random streams = the reference's distributions (fish_mod.cpp:28-33) */
enum {
    FISHABM_STREAM_CHOICE = 0, FISHABM_STREAM_ERROR = 1, FISHABM_STREAM_RANDOMWALK = 2,
    FISHABM_STREAM_QUALITY = 3, FISHABM_STREAM_SEED = 4, FISHABM_STREAM_SAMPLE = 5, FISHABM_STREAM_INIT = 6
};

typedef struct fishabm_ctx fishabm_ctx;

/* the behavioural constants of move() / sample() / feed() (literals in the reference; defaults = its values).  One set per
 * context; an ensemble may give every replicate its own (fishabm_set_replicate_models: parameter sweeps). */
typedef struct fishabm_model_params {
    int32_t lag_after_feed;     /* 10   handling time after a feeding attempt, fish_mod.cpp:548 */
    int32_t sample_rate_fed;    /* 30   sample_rate after a successful feed(), :570 */
    int32_t sample_rate_unfed;  /* 0    sample_rate after an unsuccessful feed(), :573 */
    int32_t sample_draw_max;    /* 35   distrib_sample(0, 35), :33: a fish samples when sample_rate > draw */
    int32_t error_range;        /* 10   distrib_error(-10, 10), :29 (degrees added to every turn) */
    int32_t randomwalk_range;   /* 45   distrib_randomwalk(-45, 45), :30 (turn of a fish that sees nobody; initial headings) */
    int32_t alone_align_min;    /* 5    a fish with no avoid-zone neighbour and fewer than 5 align-zone neighbours is "alone"
                                        and does not sample, :546 */
    int32_t reserved;
    double fed_speed_factor;    /* 0.5  speed_factor after a successful feed(), :566 */
    double weight_align;        /* 1.7  averageturn's comb_weight for the alignment turn, :385-386 (1.7 - i, i = 0) */
    double weight_attract;      /* 0.7  ... and for the attraction turn (1.7 - i, i = 1) */
} fishabm_model_params;

typedef struct fishabm_params {
    int32_t struct_size;   /* sizeof(fishabm_params), set by fishabm_params_default */
    int32_t n;             /* fish per school            (#define num / inds.n, fish_mod.cpp:23,37) */
    int32_t w, h;          /* world size                 (Structure.w/.h, fish_mod.cpp:58-59); multiples of 200,
                              at least 600 in cell-list mode */
    int32_t mode;          /* FISHABM_MODE_* */
    int32_t device;        /* CUDA device ordinal */
    int32_t replicates;    /* 1 = one school; >1 = ensemble of independent schools of n <= 1024 fish, each in its own
                              w x h window with its own food grid.  Arrays then hold replicates*n entries,
                              replicate-major.  fishabm_in_zone (arbitrary fov/r) and fishabm_feed are not available
                              for ensembles; fishabm_step completes its last sample() inside the call. */
    int32_t rolling;       /* 1 = get_speed active (the committed reference, fish_mod.cpp:173); 0 = get_speed skipped:
                              speed_factor only changes through feed().  (The code that produced the reference's
                              data/norolling_* files is not in its repository; this switch does not reproduce them.) */
    uint64_t seed;         /* Philox key; draws are addressed by (seed, replicate, step, fish id, stream, k) */
    double speed_norm;     /* the macro `num` in get_speed (fish_mod.cpp:590); 0 = use n */
    /* slab decomposition of one school across GPUs (one context per rank); 1 rank = whole world */
    int32_t rank, nranks;
    int32_t n_global;      /* reserved: 0 or n (n is always the size of the WHOLE school, on every rank) */
    int32_t capacity;      /* per-rank entry capacity (owned fish + ghosts + one step's arrivals) when slabs are used
                              (0 = automatic: 1.3 x the uniform-density expectation) */
    int32_t halo_capacity;    /* ghosts per slab message (0 = automatic: 2 x fish per cell column + 4096) */
    int32_t migrant_capacity; /* migrants per slab message (0 = automatic) */
    uint32_t replicate0;   /* replicate id of the (first) school: part of the Philox key, so that an ensemble sharded
                              across GPUs, or one school of it run alone, draws the same numbers */
    int32_t force_slab;    /* 1 = use halo columns and the message path even with nranks == 1 (the slab is then its own
                              neighbour); for testing the decomposition on one GPU */
    fishabm_model_params model;  /* set to the reference's values by fishabm_params_default */
} fishabm_params;

/* pair counters of the most recent neighbour pass (for the roofline accounting) */
typedef struct fishabm_pass_stats {
    uint64_t focal;        /* focal fish evaluated */
    uint64_t candidates;   /* candidate pairs examined (3x3 stencil, j != i) */
    uint64_t interacting;  /* pairs inside the 330 deg cone and r < 200 that produced a turn */
    uint64_t rechecked;    /* pairs whose predicates were re-evaluated in fp64 */
} fishabm_pass_stats;

int fishabm_params_default(fishabm_params* p);
int fishabm_model_params_default(fishabm_model_params* m);
/* ensembles: one model per replicate (count == replicates), e.g. a sweep of sample_rate_fed x fed_speed_factor x speed;
 * single schools: count == 1 replaces the context's model.  Takes effect from the next pass; completes a pending sample() first. */
int fishabm_set_replicate_models(fishabm_ctx* ctx, const fishabm_model_params* models, int32_t count);
int fishabm_create(const fishabm_params* p, fishabm_ctx** out);
void fishabm_destroy(fishabm_ctx* ctx);
/* message of the last failing call on ctx (ctx == NULL: last fishabm_create failure on this thread) */
const char* fishabm_last_error(const fishabm_ctx* ctx);

/* ---- state (replaces direct member access on `individuals`, fish_mod.cpp:35-49) --------------------
 * Arrays hold replicates*n entries, replicate-major.  Any pointer may be NULL in get_state to skip it. */
int fishabm_set_state(fishabm_ctx* ctx, const double* coords, const double* dir, const double* speed,
                      const double* speed_factor, const int32_t* lag, const int32_t* sample_rate,
                      const int32_t* food_intake);
int fishabm_get_state(fishabm_ctx* ctx, double* coords, double* dir, double* speed, double* speed_factor,
                      int32_t* lag, int32_t* sample_rate, int32_t* food_intake);
/* initialize(), fish_mod.cpp:183-206: lag 0, sample_rate 0, speed_factor 1, intake 0, integer heading from
 * the randomwalk stream, integer-lattice position centre +-100 (from the init stream instead of rand()). */
int fishabm_initialize(fishabm_ctx* ctx, double speed);

/* ---- food grid (int quality[][n_cols], fish_mod.cpp:119; get_environment :603-629; sum_array :593-601) */
int fishabm_set_quality(fishabm_ctx* ctx, const int32_t* q, int32_t rows, int32_t cols, int32_t row_stride);
int fishabm_get_quality(fishabm_ctx* ctx, int32_t replicate, int32_t* q, int32_t rows, int32_t cols, int32_t row_stride);
int fishabm_sum_quality(fishabm_ctx* ctx, int64_t* sums /* [replicates] */);
/* parse a quality.csv the way get_environment does (comma separated ints, one row per line) */
int fishabm_load_quality_csv(const char* path, int32_t* q, int32_t max_rows, int32_t max_cols, int32_t row_stride,
                             int32_t* rows_out, int32_t* cols_out);

/* create_environment(environment, quality, struc), fish_mod.cpp:482-531: a procedural food map -- n_seeds cells of
 * quality 10 spread by 20 forward/backward sweeps that redraw a cell (0..10) whenever a lower/right (resp. upper/left)
 * neighbour is richer than the sweep index.  The reference draws the seed count from distrib_seed(0,0), i.e. always 0
 * (an empty map; its main() loads quality.csv instead); here it is an argument.  Host-side, one-off; draws are the
 * k-th of their stream in call order (fish id 0, step 0).  Values reach 10: they fit the library's 8-bit grid. */
int fishabm_create_environment(uint64_t seed, uint32_t replicate, int32_t n_seeds, int32_t* q, int32_t rows, int32_t cols,
                               int32_t row_stride);

/* ---- frozen-state queries: one result per fish, no mutation ------------------------------------------ */
/* in_zone(inds,id,fov,r,struc) for every id, fish_mod.cpp:226-260.  r <= 200. */
int fishabm_in_zone(fishabm_ctx* ctx, int32_t fov_deg, double r, int32_t* out);
/* the four counts of the step: out[id] = { in_zone(330,15), in_zone(330,100), in_zone(330,200), in_zone(180,200) } */
int fishabm_zone_counts(fishabm_ctx* ctx, int32_t* out4);
/* social(inds,id,330,struc) for every id, fish_mod.cpp:262-377 (random-walk / tie draws use the current step).
 * n_* may be NULL. */
int fishabm_social(fishabm_ctx* ctx, double* turn, int32_t* n_avoid, int32_t* n_align, int32_t* n_attract);
/* get_speed's value 1 + 2*in_zone(180,200)/num for every id, fish_mod.cpp:588-591, without storing it */
int fishabm_get_speed(fishabm_ctx* ctx, double* out);

/* fp64 check mode: the four in_zone counts and the social() turn of the k listed fish, by brute force over ALL fish
 * of the school in double precision, restating fish_mod.cpp:226-260, 262-377, 379-464 literally -- independent of the
 * cell list, of the fp32 pair arithmetic and of the fixed-point sums the product path uses (sums are reduced in a
 * different order than the reference's id order: agreement with a CPU evaluation is ~1e-12 deg, not bitwise).
 * flags[i]: bit 0 random walk, bit 1 tie-break draws consumed, bit 2 a bearing within 1e-4 deg of a discontinuity of
 * the turn rule (the reference's own answer then hinges on rounding noise).  Whole-school contexts only; O(k * n). */
int fishabm_check_fp64(fishabm_ctx* ctx, const int32_t* ids, int32_t k, int32_t* counts4, double* turn, int32_t* flags);

/* ---- stepping ------------------------------------------------------------------------------------------ */
/* move(inds,struc), fish_mod.cpp:163-181, synchronous form (every fish reads the pre-call state) */
int fishabm_move(fishabm_ctx* ctx);
/* sample(inds,quality,env,struc), fish_mod.cpp:540-558 (feeding conflicts resolved lowest id first, as the
 * reference's id-ordered loop does).  Advances the step counter. */
int fishabm_sample(fishabm_ctx* ctx);
/* nsteps x { move; sample } (fish_mod.cpp:126-131) with sample(t) and move(t+1) sharing one neighbour pass;
 * results are bit-identical to calling fishabm_move / fishabm_sample alternately (see fishabm_flush). */
int fishabm_step(fishabm_ctx* ctx, int32_t nsteps);
/* fishabm_step leaves the sample() of its last step pending so that it can share the neighbour pass of the next
 * move(); every call that observes or replaces state completes it implicitly, fishabm_flush does so explicitly. */
int fishabm_flush(fishabm_ctx* ctx);
/* feed(inds,id,...), fish_mod.cpp:560-574, for the listed ids in the given order */
int fishabm_feed(fishabm_ctx* ctx, const int32_t* ids, int32_t k);
/* the reference's commented-out loggers (fish_mod.cpp:132-141,149-154) over nsteps steps:
 * sumq[t] = sum_array after step t ([nsteps][replicates]); intake rows every `intake_every` steps
 * ([rows][replicates*n]); either may be NULL. */
int fishabm_run_logged(fishabm_ctx* ctx, int32_t nsteps, int64_t* sumq, int32_t* intake, int32_t intake_every);

int fishabm_get_step(const fishabm_ctx* ctx, uint32_t* step);
int fishabm_set_step(fishabm_ctx* ctx, uint32_t step);

/* ---- introspection ----------------------------------------------------------------------------------- */
/* the integer the library draws for (stream, step, fish id, k) mapped to [a,b] like
 * std::uniform_int_distribution<int>(a,b) on a 32-bit engine (fish_mod.cpp:27-33) */
int fishabm_draw(const fishabm_ctx* ctx, int32_t stream, uint32_t replicate, uint32_t step, uint32_t id, uint32_t k,
                 int32_t a, int32_t b, int32_t* out);
/* device time of the most recent fishabm_move/sample/step call, CUDA events on the library's stream */
int fishabm_last_step_ms(const fishabm_ctx* ctx, float* ms);
/* device time of the neighbour kernel alone, summed over the (first 4096) launches of the most recent call, and that
 * launch count */
int fishabm_last_neighbour_ms(const fishabm_ctx* ctx, float* ms, int32_t* launches);
int fishabm_last_pass_stats(fishabm_ctx* ctx, fishabm_pass_stats* out);
/* kernels launched by this context since creation (for the benchmark's gpu_launches) */
int fishabm_launch_count(const fishabm_ctx* ctx, uint64_t* out);
/* enable pair counters (small cost) */
int fishabm_enable_stats(fishabm_ctx* ctx, int32_t on);
/* phase timing of fishabm_step / fishabm_step_enqueue (off by default: five event records per step).  ms4 = device time
 * summed over the (first 512) steps of the most recent call, by phase:
 *   [0] neighbour pass   [1] per-fish update (+ packing / publishing the slab messages)
 *   [2] slab exchange (waiting for the neighbours' messages; 0 for a whole school)   [3] ingest + counting sort */
int fishabm_enable_phase_timing(fishabm_ctx* ctx, int32_t on);
int fishabm_last_phase_ms(const fishabm_ctx* ctx, float* ms4, int32_t* steps);

/* ---- slab decomposition: one school split along x over nranks GPUs, one context per rank ---------------------
 * (no reference counterpart: fish_mod.cpp is single-threaded; BASELINE.json north_star asks for it.)
 * Rank r owns the cell columns [r*nx/nranks, (r+1)*nx/nranks) of the nx = w/200 columns (at least 2 each, so
 * nranks <= nx/2), plus one halo column per side.  In slab mode fishabm_set_state / fishabm_set_quality take the GLOBAL arrays (n_global fish, the whole grid)
 * on every rank and keep the rank's part; fishabm_get_state / get_quality / the frozen-state queries write only the
 * entries of the fish (columns) the rank owns and leave the rest of the caller's arrays untouched;
 * fishabm_sum_quality returns the rank's partial sum.  Integer state is bit-identical for any nranks.
 *
 * After every move() each slab sends ONE fixed-size message to each neighbour (migrants + its boundary column as
 * the neighbour's next halo).  With fishabm_slab_connect the library moves them itself with NCCL send/recv on its
 * own stream (no host synchronisation per step).  Without it the caller moves them: begin_step, copy each
 * out_left -> left neighbour's in_right and out_right -> right neighbour's in_left, end_step. */
#define FISHABM_UNIQUE_ID_BYTES 128
typedef struct fishabm_slab_info {
    int32_t rank, nranks, left, right;   /* ring neighbours (periodic world) */
    int32_t col0, col1;                  /* owned global cell columns [col0, col1) */
    int32_t owned, held;                 /* fish owned now; entries held now (owned + ghosts) */
    int32_t capacity, halo_capacity, migrant_capacity;
    int64_t message_bytes;
} fishabm_slab_info;
int fishabm_slab_get_info(fishabm_ctx* ctx, fishabm_slab_info* out);
/* ncclGetUniqueId (rank 0; distribute the bytes to the other ranks by any means) */
int fishabm_slab_unique_id(void* id_out, int32_t bytes);
/* ncclCommInitRank with the context's rank / nranks on its device; collective over all ranks */
int fishabm_slab_connect(fishabm_ctx* ctx, const void* id, int32_t bytes);
/* caller-driven exchange */
int fishabm_slab_begin_step(fishabm_ctx* ctx);
int fishabm_slab_buffers(fishabm_ctx* ctx, void** out_left, void** out_right, void** in_left, void** in_right, int64_t* bytes);
int fishabm_slab_end_step(fishabm_ctx* ctx);
/* device-to-device copy helper for the caller-driven exchange (cudaMemcpy, synchronous) */
int fishabm_slab_copy(void* dst_device, const void* src_device, int64_t bytes);
/* Peer-memory transport (preferred on NVLink / NVSwitch boxes): every slab owns an INBOX (two sides x two slots of one
 * message each, plus arrival flags) in its GPU's memory.  Once a slab is attached to its two neighbours' inboxes, the
 * update kernel stores each migrant / ghost record directly into the neighbour's inbox over NVLink while it packs,
 * a one-warp kernel completes the header and raises the arrival flag, and a one-warp kernel on the receiving side
 * waits for both flags: no copy kernel, no NCCL launch, no host involvement.  fishabm_slab_endpoint describes the
 * inbox (CUDA IPC handle for other processes, raw pointer for contexts of the same process); the caller carries the
 * endpoints between ranks.  Attach replaces the NCCL connection; it is collective in the same sense. */
typedef struct fishabm_slab_endpoint {
    unsigned char ipc_handle[64];  /* cudaIpcMemHandle_t of the inbox allocation */
    uint64_t raw_pointer;          /* the same allocation's device pointer, valid inside the owning process */
    int64_t bytes;                 /* size of the inbox allocation */
    int32_t pid, device, rank;
    int32_t pad;                   /* identifier of the physical GPU (filled by fishabm_slab_get_endpoint) */
} fishabm_slab_endpoint;
int fishabm_slab_get_endpoint(fishabm_ctx* ctx, fishabm_slab_endpoint* out);
/* Every slab must be on its OWN GPU: kernels of one GPU are not guaranteed to run concurrently, so a slab waiting (on the
 * device) for a neighbour on the same GPU could wait forever.  fishabm_slab_attach refuses such an attachment with
 * FISHABM_E_ARG; slabs that share a GPU use the caller-driven exchange above. */
int fishabm_slab_attach(fishabm_ctx* ctx, const fishabm_slab_endpoint* left, const fishabm_slab_endpoint* right);
/* Slab-local state (what a rank of a large run should use: no rank ever holds the whole school).
 * fishabm_set_state_owned: `count` fish with global ids ids[], every one inside this slab's columns (x in
 * [col0*200, col1*200)); afterwards the slabs refresh each other's halo columns with one exchange.  Collective over
 * all ranks; the call only ENQUEUES (so that several slabs of one process can be driven in turn): follow it with
 * fishabm_wait, which also reports bad input (coordinates outside the world or the slab's columns, ids outside [0, n),
 * speeds with 3*speed >= 200: all checked on the device while the arrays are ingested).  fishabm_get_state_owned: the fish this slab owns now, compact, with
 * their ids; *count is the capacity of the arrays on entry and the number of fish on return. */
int fishabm_set_state_owned(fishabm_ctx* ctx, int32_t count, const int32_t* ids, const double* coords, const double* dir,
                            const double* speed, const double* speed_factor, const int32_t* lag, const int32_t* sample_rate,
                            const int32_t* food_intake);
int fishabm_get_state_owned(fishabm_ctx* ctx, int32_t* count, int32_t* ids, double* coords, double* dir, double* speed,
                            double* speed_factor, int32_t* lag, int32_t* sample_rate, int32_t* food_intake);
/* fishabm_step without the final host synchronisation, and the matching wait (several contexts of one process that
 * are each other's neighbours must all be enqueued before any of them can finish) */
int fishabm_step_enqueue(fishabm_ctx* ctx, int32_t nsteps);
int fishabm_wait(fishabm_ctx* ctx);

/* ---- host-side scalar helpers with the reference's semantics (not on the hot path) ---------------------- */
double fishabm_averageturn(const double* v, int32_t n, int32_t comb_weight); /* fish_mod.cpp:379-408 */
double fishabm_correctangle(double dir);                                       /* fish_mod.cpp:410-418 */
void fishabm_correct_coords(double* xy, int32_t w, int32_t h);                 /* fish_mod.cpp:466-480 */

#ifdef __cplusplus
}
#endif
#endif /* FISHABM_H */
