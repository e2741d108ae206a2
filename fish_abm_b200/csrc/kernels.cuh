// Device kernels of the per-step agent update (single school, cell-list mode).
//
// Data layout (all arrays in CELL-SORTED order, rebuilt by a counting sort after every move):
//   rec   float4 {ox, oy, cos(dir), sin(dir)}   position as offset inside the fish's own 200x200 neighbour
//                                                cell (ulp <= 1.5e-5 at any world size) + heading unit vector
//   cold0 float4 {dir_deg (unwrapped), speed, speed_factor, unused}
//   cold1 int4   {lag, sample_rate, food_intake, global id}
//   cell  int    LOCAL cell index of the fish, column-major: c = lcx*ny + cy, so that a cell COLUMN is one
//                contiguous slice of every sorted array (a slab's boundary column needs no gather)
//   cell_start[ncell+1]                          exclusive scan of the per-cell counts; cell_start[ncell] = fish held
// One context holds `lnx` local columns of `ny` cells.  A whole school: lnx = W/200, periodic in x.  One slab of a
// school split across GPUs (nranks > 1): lnx = owned columns + 2, local columns 0 and lnx-1 are HALO columns whose
// fish (ghosts) are candidates only; x is not periodic locally (the halo is), y always is.
// Per-pass outputs (sorted order): cnt int4 {n15, n100, n200, front}, turn float (deg), aux uint32.
//
// Kernels:
//   K1 ingest_kernel / update_kernel  - write the fish's next cell and claim its rank in that cell (atomic)
//   K2 scan_*                         - exclusive scan of the cell counts
//   K3 scatter_kernel                 - counting-sort scatter of the three 16-byte records
//   K4 neighbour_kernel               - fused pair loop: in_zone counts at r=15/100/200 (330 deg), front count
//                                       (180 deg, r=200), avoidance/alignment/attraction turn sums, averageturn,
//                                       random walk, and sample()'s feeding decision
//   K5 update_kernel                  - sample()/feed()/move() state transition per fish
#pragma once
#include <cooperative_groups.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>

#include "philox.cuh"

namespace fishk {

constexpr int CELL_I = 200;          // neighbour cell edge = largest radius (fish_mod.cpp:275,589)
constexpr float CELL_F = 200.f;
constexpr int FOOD_I = 20;           // Structure.res_factor, fish_mod.cpp:56
constexpr int FOOD_PER_CELL = CELL_I / FOOD_I;  // 10 food cells per neighbour-cell edge

constexpr int STREAM_CHOICE = 0, STREAM_ERROR = 1, STREAM_RANDOMWALK = 2, STREAM_SAMPLE = 5, STREAM_INIT = 6;

// ---- neighbour kernel configuration
constexpr int NB_WARPS = 8;              // warps (= focal cells) per CTA
constexpr int NB_THREADS = NB_WARPS * 32;
constexpr int NB_CAP = 256;              // candidate records staged per warp batch

constexpr float PI_F = 3.14159265358979f;
constexpr float COS165_F = -0.96592582628906829f;   // cos(165 deg): the 330 deg field of view (fov/2, fish_mod.cpp:254)
constexpr float BAND_R = 2.0e-4f;        // |d - r| below this -> fp64 re-check
constexpr float COLLINEAR_BAND = 1.0e-6f; // |sin(bearing)| below this, neighbour ahead -> fp64 re-check (acos(q > 1) = NaN drop)
constexpr int FIX_SHIFT = 20;            // fixed-point fraction bits of the cos/sin accumulators
constexpr float FIX_MAGIC = 12.0f;       // x + 12.0f has ulp 2^-20 for x in [-1,1]
constexpr int FIX_MAGIC_BITS = 0x41400000;

// local grid of one context
struct Grid {
    int lnx, ny, ncell;    // local columns (halos included), rows, lnx*ny
    int own_c0, own_c1;    // owned cells [own_c0, own_c1): whole school [0,ncell); slab [ny, (lnx-1)*ny)
    int wrap_x;            // 1 = periodic in x locally (whole school); 0 = slab with halo columns
    int gx0;               // global column of local column 0
    int nx_g;              // global columns (W/200)
};

// the model constants the reference writes as literals, as run-time parameters (include/fishabm.h: fishabm_model_params);
// the defaults are the reference's values
struct Model {
    int lag_after_feed;      // 10,  fish_mod.cpp:548
    int rate_fed;            // 30,  :570
    int rate_unfed;          // 0,   :573
    int sample_draw_max;     // 35,  :33  distrib_sample(0, 35)
    int error_range;         // 10,  :29  distrib_error(-10, 10)
    int randomwalk_range;    // 45,  :30  distrib_randomwalk(-45, 45)
    int alone_align_min;     // 5,   :546 alone = no avoid-zone neighbour and fewer than 5 align-zone neighbours
    uint32_t error_reject;   // 2^32 mod (2 * error_range + 1): the rejection threshold of the error draw (philox.cuh: draw_known)
    float fed_sf;            // 0.5, :566
    float w_align;           // 1.7, :385-386 (1.7 - i, i = 0)
    float w_attract;         // 0.7  (1.7 - i, i = 1)
    float pad2;
};

struct NeighArgs {
    const float4* rec;
    const float4* cold0;
    const int4* cold1;
    const int* cell_start;
    int4* cnt;
    float* turn;
    uint32_t* aux;
    unsigned long long* stats;   // [4] focal, candidates, interacting, rechecked (may be null)
    Grid g;
    int active_lag_max;          // fish with lag > this are skipped (their outputs are not needed); <0: all fish
    int do_sample;               // evaluate sample()'s feeding decision into aux
    uint32_t step_sample, step_move;
    uint64_t seed;
    uint32_t replicate;
    int split;                   // work items per cell (each takes a slice of the cell's fish): 1 when cells outnumber
                                 // the resident warps, up to 32 for small or dense schools that would otherwise leave SMs idle
    int single_cols;             // v4, split == 1: the LAST `single_cols` owned columns are handed out cell by cell instead of as cell
                                 // pairs (same arithmetic, half the item length: a shorter tail when the resident warps run dry)
    int* work_counter;           // work-item dispenser of the persistent kernels, zero at launch
    int* work_counter_next;      // the other dispenser: cleared by this launch for the next one
    Model model;
};

// aux bits
constexpr uint32_t AUX_WANTS = 1u;            // fish decided to feed (fish_mod.cpp:547)
constexpr uint32_t AUX_FOOD_SHIFT = 1;        // bits 1..7: food cell index inside the neighbour cell (0..99)
constexpr uint32_t AUX_FOOD_MASK = 0x7Fu;
constexpr uint32_t AUX_RANDOMWALK = 1u << 8;  // turn came from the random-walk stream (fish_mod.cpp:372-374)
constexpr uint32_t AUX_TIE_SHIFT = 9;         // tie-break draws consumed (saturating 7 bits)
constexpr uint32_t AUX_EVALUATED = 1u << 16;  // the fish was active in this pass

// MUFU.RSQ / MUFU.RCP without the denormal fix-up code nvcc wraps around rsqrtf / __fdividef (three extra
// instructions each); identical results for normal arguments, and every argument here is >= 1e-37.
__device__ __forceinline__ float rsqrt_ftz(float x) { float r; asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ float rcp_ftz(float x) { float r; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }

// atan2 in radians, result in (-pi, pi]; max abs error ~1.5e-7 rad.  (0,0) -> 0.
__device__ __forceinline__ float atan2_fast(float y, float x) {
    const float ax = fabsf(x), ay = fabsf(y);
    const float mx = fmaxf(ax, ay), mn = fminf(ax, ay);
    const float t = mn * rcp_ftz(fmaxf(mx, 1e-37f));
    const float s = t * t;
    float p = -4.054543159e-03f;
    p = fmaf(p, s, 2.186286711e-02f);
    p = fmaf(p, s, -5.591218983e-02f);
    p = fmaf(p, s, 9.642186820e-02f);
    p = fmaf(p, s, -1.390862525e-01f);
    p = fmaf(p, s, 1.994656476e-01f);
    p = fmaf(p, s, -3.332986071e-01f);
    p = fmaf(p, s, 9.999993356e-01f);
    float r = p * t;
    r = (ay > ax) ? (0.5f * PI_F - r) : r;
    r = (x < 0.f) ? (PI_F - r) : r;
    return copysignf(r, y);
}

// cos / sin of every INTEGER heading in [-720, 1080] degrees as the HOST's libm (glibc, the reference's own libm)
// evaluates cos(dir * M_PI / 180) / sin(...) (fish_mod.cpp:232-233), uploaded once per device by fishabm_create.  CUDA's
// fp64 cos/sin may differ from glibc's in the last ulp; on integer-lattice states with integer headings -- the reference's
// own initial state, fish_mod.cpp:196-202 -- exactly collinear neighbours make the reference's `acos(q)` flip between a
// number and NaN on that ulp, so the literal fp64 predicates below take the unit vector from this table
// whenever the heading is an integer; any other heading uses the device's cos/sin (no exact collinearity to decide).
constexpr int UNIT_DEG_MIN = -720, UNIT_DEG_MAX = 1080;
__device__ double2 g_unit_deg[UNIT_DEG_MAX - UNIT_DEG_MIN + 1];
__device__ __forceinline__ void unit_vector_fp64(float dir_deg, double& c, double& s) {
    const float r = rintf(dir_deg);
    if (r == dir_deg && r >= (float)UNIT_DEG_MIN && r <= (float)UNIT_DEG_MAX) {
        const double2 u = g_unit_deg[(int)r - UNIT_DEG_MIN];
        c = u.x; s = u.y;
        return;
    }
    const double ang = __ddiv_rn(__dmul_rn((double)dir_deg, M_PI), 180.0);
    c = cos(ang); s = sin(ang);
}

// (column, row) of a column-major cell index without an integer division (~25 instructions each): the fp32 quotient is
// within one or two of the true one for any grid, and the remainder fixes it up exactly
__device__ __forceinline__ void cell_xy(int c, int ny, int& cx, int& cy) {
    int q = (int)(__int2float_rz(c) * __frcp_rn(__int2float_rn(ny)));
    int r = c - q * ny;
    while (r < 0) { q -= 1; r += ny; }     // (at most one iteration up to ~4M columns; a loop so that it is exact beyond)
    while (r >= ny) { q += 1; r -= ny; }
    cx = q; cy = r;
}

// Literal fp64 restatement of the per-pair predicates of in_zone (fish_mod.cpp:237-254) for a pair whose fp32
// evaluation fell inside a rounding band.  dx,dy are exact (offsets are fp32, cell shifts are multiples of 200).
// Products and sums are kept unfused so the operation order is the reference's.
// bit0 dist<15, bit1 dist<100, bit2 dist<200, bit3 angle<165, bit4 angle<90
// kcode = (kx + 1) + 3 * (ky + 1): the candidate's cell is (kx, ky) cells away from the focal fish's, kx in -1..1, ky >= -10
// (|ky| = 2 occurs when a warp stages a vertical pair of cells: neighbour_v4.cuh)
__device__ __noinline__ unsigned exact_pair_bits(float oxi, float oyi, float dir_deg, float oxj, float oyj, int kcode) {
    const double sx = (double)(((kcode + 30) % 3) - 1) * 200.0, sy = (double)(((kcode + 30) / 3) - 11) * 200.0;
    const double x = __dadd_rn(__dsub_rn((double)oxj, (double)oxi), sx);
    const double y = __dadd_rn(__dsub_rn((double)oyj, (double)oyi), sy);
    double c, s;
    unit_vector_fp64(dir_deg, c, s);
    const double dist = __dsqrt_rn(__dadd_rn(__dmul_rn(x, x), __dmul_rn(y, y)));
    const double q = __ddiv_rn(__dadd_rn(__dmul_rn(c, x), __dmul_rn(s, y)), dist);
    const double angle = __ddiv_rn(__dmul_rn(acos(q), 180.0), M_PI);
    unsigned b = 0;
    if (dist < 15.0) b |= 1u;
    if (dist < 100.0) b |= 2u;
    if (dist < 200.0) b |= 4u;
    if (angle < 165.0) b |= 8u;
    if (angle < 90.0) b |= 16u;
    return b;
}

// getangle's wrap applied to the standard representative beta (rad) for a focal fish whose raw heading is
// outside (-540, 540): world bearing A in [0,360), minus the raw heading, ONE correctangle, fold > 180
// (fish_mod.cpp:440-459).  Everywhere else the result equals beta, which the fast path returns directly.
__device__ __noinline__ float getangle_wrap_exotic(float beta_rad, float dir_raw_deg) {
    const double b = (double)beta_rad * (180.0 / M_PI);
    double dm = fmod((double)dir_raw_deg, 360.0);
    if (dm < 0) dm += 360.0;
    double A = b + dm;
    if (A < 0) A += 360.0; else if (A >= 360.0) A -= 360.0;
    double a = A - (double)dir_raw_deg;
    if (a < 0) a += 360.0; else if (a >= 360.0) a -= 360.0;
    if (a > 180.0) a -= 360.0;
    return (float)(a * (M_PI / 180.0));
}

// The y frame of the pair arithmetic.  A fish of an EVEN cell row computes its pair geometry in the frame of the cell row
// above it: its own offset becomes RN(oy - 200) and a candidate `ky` rows away RN(qy + (ky - 1) * 200); a fish of an odd
// row uses its own cell's frame (offset unchanged, candidates RN(qy + ky * 200)).  Either way |y| < 400 (one rounding of
// <= 1.5e-5 per coordinate, inside the bands' budget), and a warp that owns the vertical cell pair (2p, 2p + 1) stages ONE
// window for the fish of both cells (neighbour_v4.cuh).  Every kernel design and the ensemble kernel use this frame, so
// that they stay bit-identical to each other.
__device__ __forceinline__ float yframe_shift(int cy) { return (cy & 1) ? 0.f : -CELL_F; }

// cell k (0..8) of the 3x3 stencil around cell c: x offset k%3-1, y offset k/3-1.  Owned cells of a slab always
// have both x neighbours inside the local grid (the halo columns).
__device__ __forceinline__ int stencil_cell(const Grid& g, int c, int k) {
    int xx = c / g.ny + (k % 3) - 1, yy = c % g.ny + (k / 3) - 1;
    if (g.wrap_x) xx = xx < 0 ? xx + g.lnx : (xx >= g.lnx ? xx - g.lnx : xx);
    yy = yy < 0 ? yy + g.ny : (yy >= g.ny ? yy - g.ny : yy);
    return xx * g.ny + yy;
}

struct FocalAcc {
    int n15, n100, n200, nf;        // running counts
    int ax0, ay0, ax1, ay1, ax2, ay2;  // fixed-point cos/sin sums of the current batch (avoid, align, attract)
    double sx0, sy0, sx1, sy1, sx2, sy2;  // folded sums (exact integers)
    int ties;
    unsigned cand, inter, rechk;
};

__device__ __forceinline__ void fold_batch(FocalAcc& A) {
    A.sx0 += (double)A.ax0; A.sy0 += (double)A.ay0;
    A.sx1 += (double)A.ax1; A.sy1 += (double)A.ay1;
    A.sx2 += (double)A.ax2; A.sy2 += (double)A.ay2;
    A.ax0 = A.ay0 = A.ax1 = A.ay1 = A.ax2 = A.ay2 = 0;
}

// averageturn (fish_mod.cpp:379-408) from the cos/sin sums; degrees
__device__ __forceinline__ double average_from_sums(double x, double y) {
    if (y == 0.0) return (x >= 0.0) ? 0.0 : 180.0;
    return atan2(y, x) * (180.0 / M_PI);
}

// fp32 form used by the candidate-parallel kernel: integer sums in, degrees out
__device__ __forceinline__ float average_from_sums_f(long long x, long long y) {
    if (y == 0) return (x >= 0) ? 0.f : 180.f;
    return atan2_fast((float)y, (float)x) * (180.f / PI_F);
}

// Per-fish end of the neighbour pass, shared by every kernel design: social()'s selection (fish_mod.cpp:355-376),
// averageturn (:379-408), the random walk (:372-374) and sample()'s feeding decision (:543-547).
// s[6] = fixed-point cos/sin sums (avoid x,y; align x,y; attract x,y), exact integers.
struct FocalResult { int4 cnt; float turn; uint32_t aux; };
// sample()'s decision for one fish (fish_mod.cpp:543-547): AUX_WANTS + the food cell inside its neighbour cell, or 0
__device__ __forceinline__ uint32_t sample_decision(const Model& m, uint64_t seed, uint32_t replicate, uint32_t step_sample, uint32_t id, int lag,
                                                    int sample_rate, float fx, float fy, int n_av, int n_al) {
    if (lag != 0) return 0u;  // the draw is only consumed when lag == 0 (fish_mod.cpp:547)
    const int dr = fishrng::draw(seed, replicate, step_sample, id, STREAM_SAMPLE, 0, 0, m.sample_draw_max);
    const bool alone = (n_av == 0 && n_al < m.alone_align_min);
    if (!(sample_rate > dr && !alone)) return 0u;
    const int fxl = min((int)fx / FOOD_I, FOOD_PER_CELL - 1), fyl = min((int)fy / FOOD_I, FOOD_PER_CELL - 1);
    return AUX_WANTS | ((uint32_t)(fyl * FOOD_PER_CELL + fxl) << AUX_FOOD_SHIFT);
}
__device__ __forceinline__ FocalResult focal_result(const Model& m, uint64_t seed, uint32_t replicate, uint32_t step_sample, uint32_t step_move, int do_sample,
                                                    uint32_t id, int lag, int sample_rate, float fx, float fy, int n15, int n100, int n200,
                                                    int nfr, const long long* s, int ties) {
    uint32_t aux = AUX_EVALUATED;
    const int n_av = n15, n_al = n100 - n15, n_at = n200 - n100;
    float turn;  // degrees
    if (n_av > 0) turn = average_from_sums_f(s[0], s[1]);
    else if (n_al > 0 && n_at > 0) {
        const float aa = average_from_sums_f(s[2], s[3]) * (PI_F / 180.f);
        const float tt = average_from_sums_f(s[4], s[5]) * (PI_F / 180.f);
        float sa, ca, st, ct;
        __sincosf(aa, &sa, &ca);
        __sincosf(tt, &st, &ct);
        const float xs = fmaf(m.w_align, ca, m.w_attract * ct), ys = fmaf(m.w_align, sa, m.w_attract * st);
        turn = (ys == 0.f) ? (xs >= 0.f ? 0.f : 180.f) : atan2_fast(ys, xs) * (180.f / PI_F);
    } else if (n_al > 0) turn = average_from_sums_f(s[2], s[3]);
    else if (n_at > 0) turn = average_from_sums_f(s[4], s[5]);
    else {
        turn = (float)fishrng::draw(seed, replicate, step_move, id, STREAM_RANDOMWALK, 0, -m.randomwalk_range, m.randomwalk_range);
        aux |= AUX_RANDOMWALK;
    }
    aux |= (uint32_t)min(ties, 127) << AUX_TIE_SHIFT;
    if (do_sample) aux |= sample_decision(m, seed, replicate, step_sample, id, lag, sample_rate, fx, fy, n_av, n_al);
    FocalResult r;
    r.cnt = make_int4(n15, n100, n200, nfr); r.turn = turn; r.aux = aux;
    return r;
}

__device__ __forceinline__ void finish_focal(const NeighArgs& a, int idx, int4 k1, float fx, float fy, int n15, int n100, int n200,
                                             int nfr, const long long* s, int ties) {
    const FocalResult r = focal_result(a.model, a.seed, a.replicate, a.step_sample, a.step_move, a.do_sample, (uint32_t)k1.w, k1.x, k1.y, fx, fy,
                                       n15, n100, n200, nfr, s, ties);
    a.cnt[idx] = r.cnt;
    a.turn[idx] = r.turn;
    a.aux[idx] = r.aux;
}

// One candidate pair, one lane: exact classification (with the fp64 re-check of borderline predicates), zone / front
// counts, and the avoidance / alignment / attraction turn of the pair (fish_mod.cpp:316-351) added as fixed-point
// cos/sin to the lane's partial sums.  (qx,qy) is the candidate already shifted into the focal cell's frame, (qc,qs) its
// heading.  `raw(rx, ry, kcode, fxr, fyr)` yields the candidate's unshifted in-cell offset and its stencil code for the
// re-check, and overwrites (fxr, fyr) -- preset to (fx, fy) -- with the focal fish's unshifted offset where that differs;
// `ids(my, nb)` yields the two fish ids for a tie draw.  Returns true when the pair was re-checked in fp64.
// counts are packed two to a word (16 bits each: a lane sees at most 32 candidates per reduction, a warp 1024):
//   c_far = n200 | front << 16 | ties << 27,  c_near = n100 | n15 << 16
// (ties = tie-break draws consumed, a diagnostic kept modulo 32)
struct PairAcc { int c_far, c_near; int ax0, ay0, ax1, ay1, ax2, ay2; };
constexpr int PAIR_TIE_UNIT = 1 << 27;
__device__ __forceinline__ void pair_acc_clear(PairAcc& A) { A.c_far = A.c_near = 0; A.ax0 = A.ay0 = A.ax1 = A.ay1 = A.ax2 = A.ay2 = 0; }
__device__ __forceinline__ int4 unpack_counts(int c_far, int c_near) {  // n15, n100, n200, front
    return make_int4(c_near >> 16, c_near & 0xFFFF, c_far & 0xFFFF, (c_far >> 16) & 0x7FF);
}
__device__ __forceinline__ int unpack_ties(int c_far) { return (int)((unsigned)c_far >> 27); }

template <class Raw, class Ids>
__device__ __forceinline__ bool pair_math(bool live, float fx, float fy, float fc, float fs, float fdir, bool exotic, float qx, float qy,
                                          float qc, float qs, uint64_t seed, uint32_t replicate, uint32_t step_move, const Raw& raw,
                                          const Ids& ids, PairAcc& A) {
    const float dx = qx - fx, dy = qy - fy;
    const float d2 = fmaf(dx, dx, dy * dy);
    // d2 == 0: the fish itself or a coincident fish: acos(0/0) is NaN in the reference, the pair is dropped from every
    // count (fish_mod.cpp:252-254)
    live = live && (d2 > 0.f);
    const float rinv = rsqrt_ftz(fmaxf(d2, 1e-30f));
    const float d = d2 * rinv;
    const float dot = fmaf(fc, dx, fs * dy);
    const float ca = dot * rinv;
    // rounding bands: position quantisation (<= 3e-5 units) dominates at short range; the constant covers the fp32
    // heading unit vector (<= 5e-7), MUFU.RSQ (2.4e-7) and the dot product's rounding with a factor of two to spare
    const float tb = fmaf(rinv, 1.0e-4f, 2.0e-6f);
    const float near_r = fminf(fminf(fabsf(d - 15.f), fabsf(d - 100.f)), fabsf(d - 200.f));
    const float near_c = fminf(fabsf(ca - COS165_F), fabsf(ca));
    // third bearing band: a neighbour EXACTLY dead ahead (integer-lattice states, the reference's own initial state,
    // fish_mod.cpp:196-202).  The reference's fp64 cosine can round to 1 + 1e-16 there, acos() is NaN and the pair drops out
    // of every count (fish_mod.cpp:252-254): only the literal fp64 evaluation can tell.  sin(bearing) = cross / d;
    // an exactly collinear pair evaluates to <= 3e-7 in fp32 (unit-vector rounding), a non-collinear fp32 pair to >= 1e-6
    // only 6e-7 of the time.
    const float cross = fmaf(fc, dy, -fs * dx);
    const bool ahead = fabsf(cross) * rinv < COLLINEAR_BAND && ca > 0.f;
    const bool unc = (near_r < BAND_R || near_c < tb || ahead) && live;
    // classification values: the fp32 distance / bearing cosine, or -- for a borderline pair -- stand-ins that make the
    // comparisons below reproduce the literal fp64 predicates (only two floats cross the rare branch, and the five
    // flags are computed once, as predicates, after it)
    float d_cls = d, ca_cls = ca;
    if (unc) {
        float rx, ry, fxr = fx, fyr = fy; int kcode;
        raw(rx, ry, kcode, fxr, fyr);
        const unsigned bits = exact_pair_bits(fxr, fyr, fdir, rx, ry, kcode);
        d_cls = (bits & 1u) ? 1.f : (bits & 2u) ? 50.f : (bits & 4u) ? 150.f : 300.f;   // dist < 15 / < 100 / < 200 / not
        ca_cls = !(bits & 8u) ? -1.f : (bits & 16u) ? 0.5f : -0.5f;                      // angle >= 165 / < 90 / in between
    }
    const bool in15 = d_cls < 15.f, in100 = d_cls < 100.f, in200 = d_cls < 200.f, vis = ca_cls > COS165_F, front = ca_cls > 0.f;
    const bool inter = live && vis && in200;
    A.c_far += (inter ? 1 : 0) + ((live && front && in200) ? 0x10000 : 0);
    int near_inc = in100 ? 1 : 0;        // in15 implies in100 (also after the fp64 re-check: same distance): one select
    near_inc = in15 ? 0x10001 : near_inc;  // chain instead of two independent conversions and an add
    A.c_near += inter ? near_inc : 0;
    if (inter) {
        const bool al = in100 && !in15;
        const float crossA = fmaf(fc, qs, -fs * qc);
        const float dotA = fmaf(fc, qc, fs * qs);
        float beta = atan2_fast(al ? crossA : cross, al ? dotA : dot);
        bool tie180 = fabsf(beta) >= PI_F;
        if (exotic) { beta = getangle_wrap_exotic(beta, fdir); tie180 = (beta == PI_F); }
        float t = beta;
        if (in15) {
            t = beta - copysignf(PI_F, beta);
            if (beta == 0.f) {  // dead ahead: +-180 at random (fish_mod.cpp:327-329), keyed by the neighbour's id
                uint32_t my, nb;
                ids(my, nb);
                t = (fishrng::draw(seed, replicate, step_move, my, STREAM_CHOICE, nb, 0, 1) ? PI_F : -PI_F);
                A.c_far += PAIR_TIE_UNIT;
            }
        } else if (al && tie180) {  // exactly opposite heading (fish_mod.cpp:460-462)
            uint32_t my, nb;
            ids(my, nb);
            t = (fishrng::draw(seed, replicate, step_move, my, STREAM_CHOICE, nb, 0, 1) ? PI_F : -PI_F);
            A.c_far += PAIR_TIE_UNIT;
        }
        const float ee = d - (in15 ? 0.f : (al ? 15.f : 200.f));
        const float kz = in15 ? 0.1f : 0.004f;
        const float wgt = rcp_ftz(fmaf(kz * ee, ee, 2.f));
        const float th = t * wgt;
        float sn, cs;
        __sincosf(th, &sn, &cs);
        const int vx = __float_as_int(cs + FIX_MAGIC) - FIX_MAGIC_BITS;
        const int vy = __float_as_int(sn + FIX_MAGIC) - FIX_MAGIC_BITS;
        if (in15) { A.ax0 += vx; A.ay0 += vy; }
        else if (al) { A.ax1 += vx; A.ay1 += vy; }
        else { A.ax2 += vx; A.ay2 += vy; }
    }
    return unc;
}

// ------------------------------------------------------------------------------------------------------
// Round form of the pair math (neighbour_kernel_v3; the ensemble kernel and neighbour_kernel_v4 use pair_fast below): 32 pairs of ONE focal fish, one per lane.
// The hot path is straight-line, branch-free code: geometry, rounding bands, classification, bearing, weight, sincos.
// Everything rare -- a borderline predicate (fp64 re-check), a tie draw -- is detected as a flag and, if ANY lane of
// the round raises it (about 1 round in 2500), the whole round is redone by the generic pair_math above before
// anything has been accumulated: one warp-uniform branch per round instead of per-lane divergence.  Focal fish with an
// exotic raw heading (|dir| >= 540) take the generic code for all their rounds (the caller's choice).
// A lane without a pair is given a candidate beyond every radius (no `live` predicate); a coincident fish (d2 == 0)
// turns into NaNs that fail every comparison, exactly the reference's NaN drop (fish_mod.cpp:252-254).
// Sums are kept as RAW float bit patterns (cos + 12.0f as an integer, i.e. value * 2^20 + FIX_MAGIC_BITS) added with
// 32-bit wrap-around; the bias n_zone * FIX_MAGIC_BITS is removed once per focal fish (pair_sums_unbias), saving two
// subtractions per pair.
// Counts.  PACK8 = true : ONE word, n15 | n100 << 8 | n200 << 16 | front << 24 (a reduction covers <= 255 candidates);
//          PACK8 = false: two words as in PairAcc (16 bits each, <= 1024 candidates per reduction), no tie diagnostic.
// ------------------------------------------------------------------------------------------------------
struct PairAccR { unsigned c_far, c_near, ax0, ay0, ax1, ay1, ax2, ay2; };
__device__ __forceinline__ void pair_acc_clear(PairAccR& A) { A.c_far = A.c_near = 0u; A.ax0 = A.ay0 = A.ax1 = A.ay1 = A.ax2 = A.ay2 = 0u; }
template <bool PACK8>
__device__ __forceinline__ int4 unpack_counts_r(unsigned c_far, unsigned c_near) {  // n15, n100, n200, front
    if (PACK8) return make_int4((int)(c_far & 0xFFu), (int)((c_far >> 8) & 0xFFu), (int)((c_far >> 16) & 0xFFu), (int)(c_far >> 24));
    return make_int4((int)(c_near >> 16), (int)(c_near & 0xFFFFu), (int)(c_far & 0xFFFFu), (int)((c_far >> 16) & 0x7FFu));
}

// reduced raw sums r[6] + the counts -> exact integer sums (bias removed); valid while |sum| < 2^31
__device__ __forceinline__ void pair_sums_unbias(const unsigned* r, const int4& cc, long long* s) {
    const unsigned b0 = (unsigned)cc.x * (unsigned)FIX_MAGIC_BITS, b1 = (unsigned)(cc.y - cc.x) * (unsigned)FIX_MAGIC_BITS,
                   b2 = (unsigned)(cc.z - cc.y) * (unsigned)FIX_MAGIC_BITS;
    s[0] = (long long)(int)(r[0] - b0); s[1] = (long long)(int)(r[1] - b0);
    s[2] = (long long)(int)(r[2] - b1); s[3] = (long long)(int)(r[3] - b1);
    s[4] = (long long)(int)(r[4] - b2); s[5] = (long long)(int)(r[5] - b2);
}

// the generic per-lane code for one round, results folded into the raw accumulators
template <bool PACK8, class Raw, class Ids>
__device__ __forceinline__ bool pair_round_generic(bool live, float fx, float fy, float fc, float fs, float fdir, bool exotic, float qx, float qy,
                                                float qc, float qs, uint64_t seed, uint32_t replicate, uint32_t step_move, const Raw& raw,
                                                const Ids& ids, PairAccR& A) {
    PairAcc T;
    pair_acc_clear(T);
    const bool rechecked = pair_math(live, fx, fy, fc, fs, fdir, exotic, qx, qy, qc, qs, seed, replicate, step_move, raw, ids, T);
    const unsigned i15 = (unsigned)T.c_near >> 16, i100 = (unsigned)T.c_near & 0xFFFFu, i200 = (unsigned)T.c_far & 0xFFFFu;
    const unsigned ifr = ((unsigned)T.c_far >> 16) & 0x7FFu;
    if (PACK8) A.c_far += i15 | (i100 << 8) | (i200 << 16) | (ifr << 24);
    else { A.c_far += i200 | (ifr << 16); A.c_near += (unsigned)T.c_near; }
    A.ax0 += (unsigned)T.ax0 + i15 * (unsigned)FIX_MAGIC_BITS; A.ay0 += (unsigned)T.ay0 + i15 * (unsigned)FIX_MAGIC_BITS;
    A.ax1 += (unsigned)T.ax1 + (i100 - i15) * (unsigned)FIX_MAGIC_BITS; A.ay1 += (unsigned)T.ay1 + (i100 - i15) * (unsigned)FIX_MAGIC_BITS;
    A.ax2 += (unsigned)T.ax2 + (i200 - i100) * (unsigned)FIX_MAGIC_BITS; A.ay2 += (unsigned)T.ay2 + (i200 - i100) * (unsigned)FIX_MAGIC_BITS;
    return rechecked;
}

// live: the lane holds a pair (only the generic fallback needs to know; the fast path relies on a far-away candidate)
template <bool PACK8, class Raw, class Ids>
__device__ __forceinline__ bool pair_round(bool live, bool exotic, float fx, float fy, float fc, float fs, float fdir, float qx, float qy,
                                           float qc, float qs, uint64_t seed, uint32_t replicate, uint32_t step_move, const Raw& raw,
                                           const Ids& ids, PairAccR& A) {
    const float dx = qx - fx, dy = qy - fy;
    const float d2 = fmaf(dx, dx, dy * dy);
    const float dot = fmaf(fc, dx, fs * dy);
    const float cross = fmaf(fc, dy, -fs * dx);
    const float rinv = rsqrt_ftz(d2);          // d2 == 0 -> +inf -> d, ca = NaN -> no predicate below holds
    const float d = d2 * rinv, ca = dot * rinv;
    // rounding bands (see pair_math)
    const float tb = fmaf(rinv, 1.0e-4f, 2.0e-6f);
    const float near_r = fminf(fminf(fabsf(d - 15.f), fabsf(d - 100.f)), fabsf(d - 200.f));
    const float near_c = fminf(fabsf(ca - COS165_F), fabsf(ca));
    const bool front = ca > 0.f;
    const bool unc = near_r < BAND_R || near_c < tb || (fabsf(cross) * rinv < COLLINEAR_BAND && front);
    const bool in15 = d < 15.f, in100 = d < 100.f, in200 = d < 200.f, vis = ca > COS165_F;
    const bool inter = vis && in200;
    const bool al = in100 && !in15;
    // bearing of the neighbour (avoid / attract: of its position; align: of its heading), fish_mod.cpp:316-351
    const float crossA = fmaf(fc, qs, -fs * qc), dotA = fmaf(fc, qc, fs * qs);
    const float beta = atan2_fast(al ? crossA : cross, al ? dotA : dot);
    const bool tie = (in15 && beta == 0.f) || (al && fabsf(beta) >= PI_F);
    // warp-uniform: redo the round with the generic per-lane code (`exotic` is the same in every lane: the focal fish's)
    if (__any_sync(0xFFFFFFFFu, unc || (inter && tie)) || exotic)
        return pair_round_generic<PACK8>(live, fx, fy, fc, fs, fdir, exotic, qx, qy, qc, qs, seed, replicate, step_move, raw, ids, A);
    const float t = in15 ? beta - copysignf(PI_F, beta) : beta;
    const float ee = d - (in15 ? 0.f : (al ? 15.f : 200.f));
    const float kz = in15 ? 0.1f : 0.004f;
    const float wgt = rcp_ftz(fmaf(kz * ee, ee, 2.f));
    const float th = t * wgt;
    float sn, cs;
    __sincosf(th, &sn, &cs);
    const unsigned vx = (unsigned)__float_as_int(cs + FIX_MAGIC), vy = (unsigned)__float_as_int(sn + FIX_MAGIC);
    if (PACK8) {
        unsigned inc = in100 ? 0x010100u : 0x010000u;
        inc = in15 ? 0x010101u : inc;
        if (inter) A.c_far += inc;
        if (front && in200) A.c_far += 0x01000000u;
    } else {
        if (inter) A.c_far += 1u;
        if (front && in200) A.c_far += 0x10000u;
        unsigned near_inc = in100 ? 1u : 0u;
        near_inc = in15 ? 0x10001u : near_inc;
        if (inter) A.c_near += near_inc;
    }
    if (inter && in15) { A.ax0 += vx; A.ay0 += vy; }
    if (inter && al) { A.ax1 += vx; A.ay1 += vy; }
    if (inter && !in100) { A.ax2 += vx; A.ay2 += vy; }
    return false;
}

// The round without its fallback branch (neighbour_kernel_v4): the SAME arithmetic as pair_round's fast path, accumulated
// unconditionally; the return value says that this lane's pair needs the generic code (borderline predicate or tie draw).
// The caller ORs it over all rounds of a focal fish and, if any lane raised it, throws the accumulators away and redoes
// the fish's rounds with pair_round_generic -- the rare case costs a few rounds more, the common one loses the vote, the
// branch and its reconvergence bookkeeping from every round.
// if (p) { x += vx; y += vy; } as two predicated adds (left to itself the compiler selects vx / 0 and adds: twice the instructions)
__device__ __forceinline__ void pred_add2(bool p, unsigned& x, unsigned& y, unsigned vx, unsigned vy) {
    asm("{\n\t.reg .pred q;\n\tsetp.ne.u32 q, %4, 0;\n\t@q add.u32 %0, %0, %2;\n\t@q add.u32 %1, %1, %3;\n\t}"
        : "+r"(x), "+r"(y) : "r"(vx), "r"(vy), "r"((unsigned)p));
}
template <bool PACK8>
__device__ __forceinline__ bool pair_fast(float fx, float fy, float fc, float fs, float qx, float qy, float qc, float qs, PairAccR& A) {
    const float dx = qx - fx, dy = qy - fy;
    const float d2 = fmaf(dx, dx, dy * dy);
    const float dot = fmaf(fc, dx, fs * dy);
    const float cross = fmaf(fc, dy, -fs * dx);
    const float rinv = rsqrt_ftz(d2);          // d2 == 0 -> +inf -> d, ca = NaN -> no predicate below holds
    const float d = d2 * rinv, ca = dot * rinv;
    const float tb = fmaf(rinv, 1.0e-4f, 2.0e-6f);
    const float d15 = d - 15.f, d200 = d - 200.f;
    const float near_r = fminf(fminf(fabsf(d15), fabsf(d - 100.f)), fabsf(d200));
    const float near_c = fminf(fabsf(ca - COS165_F), fabsf(ca));
    const bool front = ca > 0.f;
    const bool unc = near_r < BAND_R || near_c < tb || (fabsf(cross) * rinv < COLLINEAR_BAND && front);
    const bool in15 = d < 15.f, in100 = d < 100.f, in200 = d < 200.f, vis = ca > COS165_F;
    const bool inter = vis && in200;
    const bool al = in100 && !in15;
    const float crossA = fmaf(fc, qs, -fs * qc), dotA = fmaf(fc, qc, fs * qs);
    const float beta = atan2_fast(al ? crossA : cross, al ? dotA : dot);
    const bool tie = (in15 && beta == 0.f) || (al && fabsf(beta) >= PI_F);
    const float t = in15 ? beta - copysignf(PI_F, beta) : beta;
    const float ee = in15 ? d : (al ? d15 : d200);   // d - 0 / d - 15 / d - 200: the same values pair_round subtracts
    const float kz = in15 ? 0.1f : 0.004f;
    const float wgt = rcp_ftz(fmaf(kz * ee, ee, 2.f));
    const float th = t * wgt;
    float sn, cs;
    __sincosf(th, &sn, &cs);
    const unsigned vx = (unsigned)__float_as_int(cs + FIX_MAGIC), vy = (unsigned)__float_as_int(sn + FIX_MAGIC);
    if (PACK8) {
        unsigned inc = in100 ? 0x010100u : 0x010000u;
        inc = in15 ? 0x010101u : inc;
        if (inter) A.c_far += inc;
        if (front && in200) A.c_far += 0x01000000u;
    } else {
        if (inter) A.c_far += 1u;
        if (front && in200) A.c_far += 0x10000u;
        unsigned near_inc = in100 ? 1u : 0u;
        near_inc = in15 ? 0x10001u : near_inc;
        if (inter) A.c_near += near_inc;
    }
    pred_add2(inter && in15, A.ax0, A.ay0, vx, vy);
    pred_add2(inter && al, A.ax1, A.ay1, vx, vy);
    pred_add2(inter && !in100, A.ax2, A.ay2, vx, vy);
    return unc || (inter && tie);
}

// heading unit vector of a stored heading (degrees, fp32): THE one place it is computed, so that a state that went
// through get_state / set_state continues bit-identically to one that stayed on the device
// (a true division: multiplying by a rounded 1/180 would triple the error of the unit vector, which the bearing bands of
// pair_math budget for)
__device__ __forceinline__ void heading_vector(float dir_deg, float& cs, float& sn) { sincospif(dir_deg / 180.f, &sn, &cs); }

// move() for one fish (fish_mod.cpp:165-179) in cell/offset form.  In: record, cold state, cell (cx,cy), social turn,
// front count, the error draw.  Out: the new record / heading / speed_factor and the UNWRAPPED new cell (the caller
// applies the world's periodicity).
__device__ __forceinline__ void move_core(float4& r, float4& c0, int lag, int& cx, int& cy, float turn, int front, int err, int rolling,
                                          float speed_norm_inv2) {
    // fp32 throughout: the state is fp32 (heading ulp 3e-5 deg = 5e-7 rad, in-cell offset ulp 1.5e-5), so each result is
    // within one ulp of the fp64 evaluation -- far inside the 1e-5 rad / 1e-4 unit parity bars -- at a fraction of
    // the instructions of the fp64 sincospi / floor.
    // Branch-free: a lagged fish (:177-179) is a moving fish with turn 0 (its heading is not re-wrapped) and step length 0, so
    // that the lanes of a warp -- 40 % moving, 60 % lagged -- run ONE instruction stream (the separate branches left half
    // the lanes idle: 15 threads per instruction in ncu).
    const bool moving = lag == 0;
    float d0 = c0.x;
    if (moving) { if (d0 < 0.f) d0 += 360.f; else if (d0 >= 360.f) d0 -= 360.f; }   // correctangle, :410-418 (moving fish only, :166)
    const float dir = moving ? d0 + (turn + (float)err) : c0.x + (float)err;       // :167-169 / :178
    float sn, cs;
    heading_vector(dir, cs, sn);   // the new heading is also the direction of the step
    const float stepl = moving ? c0.y * c0.z : 0.f;  // speed * OLD speed_factor (:171-172); a lagged fish stays where it is
    const float px = fmaf(stepl, cs, r.x), py = fmaf(stepl, sn, r.y);
    if (moving && rolling) c0.z = fmaf((float)front, speed_norm_inv2, 1.0f);  // get_speed: 1 + 2*count/num, :588-591
    // periodic wrap (correct_coords, :466-480) in cell/offset form
    // (multiplying by 1/200 may misplace a value within an ulp of a cell edge by one cell: the two fix-ups below
    // bring the offset back into [0, 200) either way)
    const float kx = floorf(px * (1.f / CELL_F)), ky = floorf(py * (1.f / CELL_F));
    float ox = fmaf(-kx, CELL_F, px), oy = fmaf(-ky, CELL_F, py);
    cx += (int)kx; cy += (int)ky;
    if (ox >= CELL_F) { ox -= CELL_F; cx += 1; }
    if (oy >= CELL_F) { oy -= CELL_F; cy += 1; }
    if (ox < 0.f) { ox += CELL_F; cx -= 1; }
    if (oy < 0.f) { oy += CELL_F; cy -= 1; }
    r.x = ox; r.y = oy; r.z = cs; r.w = sn;
    c0.x = dir;
}

template <bool STATS>
__device__ __forceinline__ void process_batch(const float4* __restrict__ tile, const uint32_t* __restrict__ tj, int fill,
                                              bool active, float fx, float fy, float fy_raw, float fc, float fs, float dir_raw,
                                              bool exotic, uint32_t my_id, const NeighArgs& a, FocalAcc& A) {
    if (!active) return;
    for (int k = 0; k < fill; ++k) {
        const float4 q = tile[k];  // same address in every lane: shared-memory broadcast
        const float dx = q.x - fx, dy = q.y - fy;
        const float d2 = fmaf(dx, dx, dy * dy);
        if (STATS) A.cand++;
        // d2 == 0: the fish itself or a coincident fish; the reference's acos(0/0) is NaN and the pair is
        // dropped from every count (fish_mod.cpp:252-254)
        if (!(d2 < (CELL_F + BAND_R) * (CELL_F + BAND_R)) || d2 == 0.f) continue;
        const float rinv = rsqrtf(d2);
        const float d = d2 * rinv;
        const float dot = fmaf(fc, dx, fs * dy);
        const float ca = dot * rinv;  // cos of the bearing of j seen from i
        // rounding bands: position quantisation (<= 3e-5 units) dominates at short range
        const float tb = fmaf(rinv, 1.0e-4f, 2.0e-6f);
        bool unc = fabsf(d - 15.f) < BAND_R || fabsf(d - 100.f) < BAND_R || fabsf(d - 200.f) < BAND_R;
        unc = unc || fabsf(ca - COS165_F) < tb || fabsf(ca) < tb;
        const float cross = fmaf(fc, dy, -fs * dx);
        unc = unc || (fabsf(cross) * rinv < COLLINEAR_BAND && ca > 0.f);  // exactly dead ahead: see pair_math
        bool in15 = d < 15.f, in100 = d < 100.f, in200 = d < 200.f, vis = ca > COS165_F, front = ca > 0.f;
        if (unc) {
            const uint32_t code = tj[k];
            const float4 raw = __ldg(&a.rec[code & 0x0FFFFFFFu]);
            const unsigned b = exact_pair_bits(fx, fy_raw, dir_raw, raw.x, raw.y, (int)(code >> 28));
            in15 = b & 1u; in100 = b & 2u; in200 = b & 4u; vis = b & 8u; front = b & 16u;
            if (STATS) A.rechk++;
        }
        A.nf += (front && in200) ? 1 : 0;
        if (!(vis && in200)) continue;
        A.n200++;
        A.n100 += in100 ? 1 : 0;
        A.n15 += in15 ? 1 : 0;
        if (STATS) A.inter++;
        // ---- turn contribution of this neighbour (fish_mod.cpp:316-351), radians
        const bool al = in100 && !in15;
        const float crossA = fmaf(fc, q.w, -fs * q.z);
        const float dotA = fmaf(fc, q.z, fs * q.w);
        float beta = atan2_fast(al ? crossA : cross, al ? dotA : dot);
        bool tie180 = fabsf(beta) >= PI_F;
        if (exotic) { beta = getangle_wrap_exotic(beta, dir_raw); tie180 = (beta == PI_F); }
        float t = beta;
        if (in15) {
            t = beta - copysignf(PI_F, beta);
            if (beta == 0.f) {  // neighbour dead ahead: +-180 at random (fish_mod.cpp:327-329); keyed by the neighbour's id
                const uint32_t nid = (uint32_t)a.cold1[tj[k] & 0x0FFFFFFFu].w;
                t = (fishrng::draw(a.seed, a.replicate, a.step_move, my_id, STREAM_CHOICE, nid, 0, 1) ? PI_F : -PI_F);
                A.ties++;
            }
        } else if (al && tie180) {  // exactly opposite heading: +-180 at random (fish_mod.cpp:460-462)
            const uint32_t nid = (uint32_t)a.cold1[tj[k] & 0x0FFFFFFFu].w;
            t = (fishrng::draw(a.seed, a.replicate, a.step_move, my_id, STREAM_CHOICE, nid, 0, 1) ? PI_F : -PI_F);
            A.ties++;
        }
        const float e = d - (in15 ? 0.f : (al ? 15.f : 200.f));
        const float kz = in15 ? 0.1f : 0.004f;
        const float wgt = __fdividef(1.f, fmaf(kz * e, e, 2.f));
        const float th = t * wgt;
        float sn, cs;
        __sincosf(th, &sn, &cs);
        const int vx = __float_as_int(cs + FIX_MAGIC) - FIX_MAGIC_BITS;
        const int vy = __float_as_int(sn + FIX_MAGIC) - FIX_MAGIC_BITS;
        if (in15) { A.ax0 += vx; A.ay0 += vy; }
        else if (al) { A.ax1 += vx; A.ay1 += vy; }
        else { A.ax2 += vx; A.ay2 += vy; }
    }
}

template <bool STATS>
__global__ void __launch_bounds__(NB_THREADS) neighbour_kernel(const NeighArgs a) {
    __shared__ float4 s_tile[NB_WARPS * NB_CAP];
    __shared__ uint32_t s_tj[NB_WARPS * NB_CAP];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const unsigned FULL = 0xFFFFFFFFu;
    const int c = a.g.own_c0 + blockIdx.x * NB_WARPS + warp;
    if (c >= a.g.own_c1) return;
    const int s0 = a.cell_start[c], e0 = a.cell_start[c + 1];
    if (s0 == e0) return;
    int nb_s = 0, nb_e = 0;
    if (lane < 9) {
        const int cc = stencil_cell(a.g, c, lane);
        nb_s = a.cell_start[cc];
        nb_e = a.cell_start[cc + 1];
    }
    float4* tile = s_tile + warp * NB_CAP;
    uint32_t* tj = s_tj + warp * NB_CAP;
    const float ysh = yframe_shift(c % a.g.ny);

    for (int base = s0; base < e0; base += 32) {
        const int i = base + lane;
        const bool valid = i < e0;
        float4 me = make_float4(0.f, 0.f, 1.f, 0.f);
        int4 c1 = make_int4(0, 0, 0, 0);
        float dir_raw = 0.f;
        if (valid) { me = a.rec[i]; c1 = a.cold1[i]; dir_raw = a.cold0[i].x; }
        const bool active = valid && (a.active_lag_max < 0 || c1.x <= a.active_lag_max);
        if (!__any_sync(FULL, active)) {
            if (valid) a.aux[i] = 0u;
            continue;
        }
        const bool exotic = fabsf(dir_raw) >= 540.f;
        FocalAcc A;
        A.n15 = A.n100 = A.n200 = A.nf = 0;
        A.ax0 = A.ay0 = A.ax1 = A.ay1 = A.ax2 = A.ay2 = 0;
        A.sx0 = A.sy0 = A.sx1 = A.sy1 = A.sx2 = A.sy2 = 0.0;
        A.ties = 0; A.cand = A.inter = A.rechk = 0;

        int fill = 0;
        for (int k = 0; k < 9; ++k) {
            const int ks = __shfl_sync(FULL, nb_s, k), ke = __shfl_sync(FULL, nb_e, k);
            const float sx = (float)((k % 3) - 1) * CELL_F, sy = (float)((k / 3) - 1) * CELL_F + ysh;
            for (int b = ks; b < ke;) {
                const int m = min(ke - b, NB_CAP - fill);
                for (int t = lane; t < m; t += 32) {
                    float4 r = __ldg(&a.rec[b + t]);
                    r.x += sx; r.y += sy;
                    tile[fill + t] = r;
                    tj[fill + t] = (uint32_t)(b + t) | ((uint32_t)k << 28);
                }
                fill += m; b += m;
                if (fill == NB_CAP) {
                    __syncwarp();
                    process_batch<STATS>(tile, tj, fill, active, me.x, me.y + ysh, me.y, me.z, me.w, dir_raw, exotic, (uint32_t)c1.w, a, A);
                    fold_batch(A);
                    __syncwarp();
                    fill = 0;
                }
            }
        }
        __syncwarp();
        process_batch<STATS>(tile, tj, fill, active, me.x, me.y + ysh, me.y, me.z, me.w, dir_raw, exotic, (uint32_t)c1.w, a, A);
        fold_batch(A);
        __syncwarp();

        if (STATS) {
            unsigned f = __reduce_add_sync(FULL, active ? 1u : 0u);
            unsigned cd = __reduce_add_sync(FULL, A.cand), it = __reduce_add_sync(FULL, A.inter), rc = __reduce_add_sync(FULL, A.rechk);
            if (lane == 0 && a.stats) {
                atomicAdd(&a.stats[0], (unsigned long long)f); atomicAdd(&a.stats[1], (unsigned long long)cd);
                atomicAdd(&a.stats[2], (unsigned long long)it); atomicAdd(&a.stats[3], (unsigned long long)rc);
            }
        }
        if (!valid) continue;
        if (!active) { a.aux[i] = 0u; continue; }

        const long long sums[6] = {(long long)A.sx0, (long long)A.sy0, (long long)A.sx1, (long long)A.sy1, (long long)A.sx2, (long long)A.sy2};
        finish_focal(a, i, c1, me.x, me.y, A.n15, A.n100, A.n200, A.nf, sums, A.ties);
    }
}

// ------------------------------------------------------------------------------------------------------
// Slab messages (nranks > 1): what one slab sends to ONE neighbour after its fish have moved.
//   migrants: fish that crossed into the neighbour's first/last owned column (full state)
//   ghosts  : the fish now in this slab's own boundary column (hot record + id), the neighbour's next halo
// Fixed capacity, counts in the header, so that the exchange needs no host round trip.
// ------------------------------------------------------------------------------------------------------
struct SlabHeader { int n_mig, n_ghost, overflow, pad; };
struct MigRec { float4 rec; float4 cold0; int4 cold1; int cy; int pad[3]; };
struct GhostRec { float4 rec; int id; int cy; int pad[2]; };

struct SlabMsg {  // view of one message buffer
    unsigned char* base;   // header + payload (in peer-memory transport: the NEIGHBOUR's inbox, written over NVLink)
    int cap_m, cap_g;
    SlabHeader* counters;  // where slots are claimed while packing: the header itself, or (peer-memory transport) a local
                           // copy that slab_publish_kernel writes into the remote header once the payload is complete
    __host__ __device__ SlabHeader* hdr() const { return counters ? counters : reinterpret_cast<SlabHeader*>(base); }
    __host__ __device__ MigRec* mig() const { return reinterpret_cast<MigRec*>(base + sizeof(SlabHeader)); }
    __host__ __device__ GhostRec* ghost() const { return reinterpret_cast<GhostRec*>(base + sizeof(SlabHeader) + (size_t)cap_m * sizeof(MigRec)); }
    __host__ __device__ static size_t bytes(int cap_m, int cap_g) { return sizeof(SlabHeader) + (size_t)cap_m * sizeof(MigRec) + (size_t)cap_g * sizeof(GhostRec); }
};

// slot claim with one atomic per converged warp (the fish of a boundary column are contiguous in sorted order, so
// whole warps arrive here together)
__device__ __forceinline__ int claim_slot(int* counter) {
    const cooperative_groups::coalesced_group g = cooperative_groups::coalesced_threads();
    int base = 0;
    if (g.thread_rank() == 0) base = atomicAdd(counter, (int)g.size());
    return g.shfl(base, 0) + (int)g.thread_rank();
}

__device__ __forceinline__ void slab_push_migrant(const SlabMsg& m, const float4& r, const float4& c0, const int4& c1, int cy) {
    const int k = claim_slot(&m.hdr()->n_mig);
    if (k < m.cap_m) { MigRec& o = m.mig()[k]; o.rec = r; o.cold0 = c0; o.cold1 = c1; o.cy = cy; }
    else m.hdr()->overflow = 1;
}
__device__ __forceinline__ void slab_push_ghost(const SlabMsg& m, const float4& r, int id, int cy) {
    const int k = claim_slot(&m.hdr()->n_ghost);
    if (k < m.cap_g) { GhostRec& o = m.ghost()[k]; o.rec = r; o.id = id; o.cy = cy; }
    else m.hdr()->overflow = 1;
}

// ------------------------------------------------------------------------------------------------------
// K5: per-fish state transition
// ------------------------------------------------------------------------------------------------------
struct UpdateArgs {
    float4* rec_out;    // where the moved record goes: `rec` itself (update_kernel: in place), or a scratch array when the
                        // update runs inside the neighbour kernel, whose other warps still read `rec` as candidates
    float4* rec;
    float4* cold0;
    int4* cold1;
    int* cell;          // current cell of the fish (sorted order)
    const int* cell_start;
    const int4* cnt;
    const float* turn;
    const uint32_t* aux;
    int* next_cell;     // out: cell after the move (-1: entry dropped by the next sort)
    int* next_rank;     // out: rank claimed in that cell
    int* next_count;    // per-cell counters being built for the next sort (atomic)
    uint8_t* quality;   // food grid of the owned columns, rows x qcols
    int* upd_count;     // deferred food-grid updates (applied after this kernel)
    int* upd_idx;
    uint8_t* upd_val;
    int* err;           // sticky error flags (ERR_*)
    Grid g;
    int qcols;
    int do_sample, do_move, rolling;
    uint32_t step_move;
    uint64_t seed;
    uint32_t replicate;
    float speed_norm_inv2;   // 2 / num (get_speed's normaliser, fish_mod.cpp:590)
    int slab;           // pack migrants / ghosts for the two neighbours
    SlabMsg out_left, out_right;
    int part;           // UPD_ALL, or (slabs, peer-memory transport) UPD_EDGES first and UPD_MIDDLE after the publish
    Model model;
};
// A fish moves less than one cell per step, so only the two outermost owned columns on each side can produce
// migrants or ghosts: updating them FIRST lets the messages leave (slab_publish_kernel) while the bulk of the slab is
// still being updated, and the neighbours' messages arrive behind that work instead of being waited for.
enum { UPD_ALL = 0, UPD_EDGES = 1, UPD_MIDDLE = 2 };

constexpr int ERR_SLAB_OVERFLOW = 1;   // a slab message or the fish arrays overflowed their capacity
constexpr int ERR_STEP_TOO_LONG = 2;   // a fish moved further than one neighbour cell in one step (slab mode)
constexpr int ERR_SLAB_TIMEOUT = 4;    // a neighbour's message did not arrive (peer-memory transport)

// ---- peer-memory transport: the update kernel has stored this step's payload directly into the two neighbours'
// inboxes (NVLink peer stores).  publish: complete each remote message with its header, then raise its flag (release,
// system scope).  wait: spin (bounded) until both of MY inbox slots carry this step's sequence number (acquire).
struct SlabPublishArgs {
    const SlabHeader* counters_left; const SlabHeader* counters_right;
    SlabHeader* remote_left_hdr; SlabHeader* remote_right_hdr;
    int* remote_left_flag; int* remote_right_flag;
    int seq;
};
__global__ void slab_publish_kernel(const SlabPublishArgs a) {
    if (threadIdx.x >= 2) return;
    const SlabHeader h = threadIdx.x == 0 ? *a.counters_left : *a.counters_right;
    SlabHeader* rh = threadIdx.x == 0 ? a.remote_left_hdr : a.remote_right_hdr;
    int* rf = threadIdx.x == 0 ? a.remote_left_flag : a.remote_right_flag;
    *rh = h;
    __threadfence_system();  // payload (previous kernel) and header are ordered before the flag, system-wide
    *reinterpret_cast<volatile int*>(rf) = a.seq;
}
__global__ void slab_wait_kernel(const int* flag_left, const int* flag_right, int seq, long long timeout_ns, int* err) {
    if (threadIdx.x >= 2) return;
    const volatile int* f = threadIdx.x == 0 ? flag_left : flag_right;
    unsigned long long t0, t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t0));
    while (*f - seq < 0) {  // wrap-safe "flag < seq"
        asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
        if ((long long)(t - t0) > timeout_ns) { atomicOr(err, ERR_SLAB_TIMEOUT); break; }
        __nanosleep(200);
    }
    __threadfence_system();
}

// the state transition of the owned fish at sorted index i
__device__ __forceinline__ void update_fish(const UpdateArgs& a, const int i, const int mycell) {
    float4 c0 = a.cold0[i];
    int4 c1 = a.cold1[i];
    const uint32_t aux = a.do_sample ? a.aux[i] : 0u;
    const int ny = a.g.ny;
    int cx0, cy0;
    cell_xy(mycell, ny, cx0, cy0);

    if (a.do_sample) {  // sample(), fish_mod.cpp:546-556
        if (aux & AUX_WANTS) {
            c1.x = a.model.lag_after_feed;
            // feed(), fish_mod.cpp:560-574.  Fish that feed on the same food cell in the same step are served in
            // id order (the reference's loop order, :541): rank = feeders of this food cell with a lower id.
            const uint32_t key = aux & ((AUX_FOOD_MASK << AUX_FOOD_SHIFT) | AUX_WANTS);
            const int s = a.cell_start[mycell], e = a.cell_start[mycell + 1];
            int rank = 0, total = 1;   // the fish itself; a second feeder on the same 20 x 20 food cell is rare (~3 %)
            for (int j = s; j < e; ++j) {
                const uint32_t aj = a.aux[j];
                if (j != i && (aj & ((AUX_FOOD_MASK << AUX_FOOD_SHIFT) | AUX_WANTS)) == key) {
                    total++;
                    rank += (a.cold1[j].w < c1.w) ? 1 : 0;
                }
            }
            const int fl = (int)((aux >> AUX_FOOD_SHIFT) & AUX_FOOD_MASK);
            const int fly = (fl * 205) >> 11, flx = fl - fly * FOOD_PER_CELL;   // fl / 10, fl % 10 for fl < 100
            const int gx = (cx0 - (a.g.wrap_x ? 0 : 1)) * FOOD_PER_CELL + flx;   // owned columns start at local column 0 (whole school) or 1 (slab)
            const int gy = cy0 * FOOD_PER_CELL + fly;
            const int qi = gy * a.qcols + gx;
            const int q = (int)a.quality[qi];
            if (rank < q) { c0.z = a.model.fed_sf; c1.z += 1; c1.y = a.model.rate_fed; }
            else c1.y = a.model.rate_unfed;
            if (rank == 0 && q > 0) {  // the grid itself is only written after every feeder has read it
                const int k = atomicAdd(a.upd_count, 1);
                a.upd_idx[k] = qi;
                a.upd_val[k] = (uint8_t)(q - min(q, total));
            }
        } else if (c1.x == 0) c1.y += 1;
        else c1.x -= 1;
    }

    if (a.do_move) {  // move(), fish_mod.cpp:165-179
        const int err = fishrng::draw_known(a.seed, a.replicate, a.step_move, (uint32_t)c1.w, STREAM_ERROR, 0, -a.model.error_range,
                                            2u * (uint32_t)a.model.error_range + 1u, a.model.error_reject);
        float4 r = a.rec[i];
        int cx = cx0, cy = cy0;
        const bool moving = c1.x == 0;
        move_core(r, c0, c1.x, cx, cy, moving ? a.turn[i] : 0.f, moving ? a.cnt[i].w : 0, err, a.rolling, a.speed_norm_inv2);
        if (a.g.wrap_x) {   // only a fish that crosses the world's edge pays for the modulo
            if (cx < 0 || cx >= a.g.lnx) { cx %= a.g.lnx; if (cx < 0) cx += a.g.lnx; }
        } else {
            // slabs: a fish may advance at most one column per step (the halo is one column wide, and with the edges-first
            // update a fish of an interior column must not reach a boundary or halo column after the messages have left)
            const int cx_old = cx0;
            if (cx < cx_old - 1 || cx > cx_old + 1 || cx < 0 || cx >= a.g.lnx) {
                atomicOr(a.err, ERR_STEP_TOO_LONG);
                cx = max(0, min(cx, a.g.lnx - 1));
            }
        }
        if (cy < 0 || cy >= ny) { cy %= ny; if (cy < 0) cy += ny; }
        const int newcell = cx * ny + cy;
        a.rec_out[i] = r;
        a.next_cell[i] = newcell;
        {   // claim a rank in the new cell: one atomic per group of lanes that go to the SAME cell (neighbouring threads are
            // neighbouring fish, so a warp targets two or three cells) instead of 32 same-address atomics the L2 serialises
            const unsigned act = __activemask();
            const unsigned peers = __match_any_sync(act, newcell);
            const int leader = __ffs(peers) - 1, lane = (int)(threadIdx.x & 31u);
            int base = 0;
            if (lane == leader) base = atomicAdd(&a.next_count[newcell], __popc(peers));
            base = __shfl_sync(peers, base, leader);
            a.next_rank[i] = base + __popc(peers & ((1u << lane) - 1u));
        }
        if (a.slab) {
            // crossing into a halo column = leaving the slab: the full state goes to that neighbour (the local
            // copy stays behind as next step's ghost); standing in a boundary column = being the neighbour's ghost
            if (cx == 0) slab_push_migrant(a.out_left, r, c0, c1, cy);
            if (cx == a.g.lnx - 1) slab_push_migrant(a.out_right, r, c0, c1, cy);
            if (cx == 1) slab_push_ghost(a.out_left, r, c1.w, cy);
            if (cx == a.g.lnx - 2) slab_push_ghost(a.out_right, r, c1.w, cy);
        }
    }
    a.cold0[i] = c0;
    a.cold1[i] = c1;
}

// 8 CTAs per SM (32 registers, a few spills): the kernel is bound by memory latency, and the warps in flight hide more of it
// than the registers would save (measured at C3: 33 us against 40 us with 48 registers / 5 CTAs; loads hoisted above the
// feeding-rank loop did not help)
#ifndef FISHABM_UPDATE_MIN_CTAS
#define FISHABM_UPDATE_MIN_CTAS 8
#endif
__global__ void __launch_bounds__(256, FISHABM_UPDATE_MIN_CTAS) update_kernel(const UpdateArgs a) {
    const int tid = blockIdx.x * blockDim.x + threadIdx.x, nthreads = gridDim.x * blockDim.x;
    const int total = a.cell_start[a.g.ncell];
    const int b1 = a.cell_start[a.g.own_c0], e2 = a.cell_start[a.g.own_c1];
    if (a.part == UPD_ALL) {  // one thread per held entry
        if (tid >= total) return;
        if (tid < b1 || tid >= e2) { if (a.do_move) a.next_cell[tid] = -1; return; }  // ghost: rebuilt by every exchange
        update_fish(a, tid, a.cell[tid]);
        return;
    }
    // owned fish = [b1,e1) (first two columns) + [e1,b2) (middle) + [b2,e2) (last two columns), contiguous in sorted order
    const int edge_c = min(a.g.own_c0 + 2 * a.g.ny, a.g.own_c1);
    const int e1 = a.cell_start[edge_c];
    const int b2 = a.cell_start[max(a.g.own_c1 - 2 * a.g.ny, edge_c)];
    if (a.part == UPD_EDGES) {  // grid-stride over: left ghosts, right ghosts, first edge, last edge
        const int nl = b1, nr = total - e2, n1 = e1 - b1, n2 = e2 - b2;
        for (int t = tid; t < nl + nr + n1 + n2; t += nthreads) {
            if (t < nl) { if (a.do_move) a.next_cell[t] = -1; }
            else if (t < nl + nr) { if (a.do_move) a.next_cell[e2 + (t - nl)] = -1; }
            else if (t < nl + nr + n1) { const int i = b1 + (t - nl - nr); update_fish(a, i, a.cell[i]); }
            else { const int i = b2 + (t - nl - nr - n1); update_fish(a, i, a.cell[i]); }
        }
    } else {                    // UPD_MIDDLE: one thread per fish of the interior columns
        const int i = e1 + tid;
        if (i < b2) update_fish(a, i, a.cell[i]);
    }
}

// the deferred food-grid writes of one sample(); also clears the OTHER update counter (the two alternate by step)
__global__ void apply_quality_updates_kernel(uint8_t* quality, const int* upd_count, int* upd_count_next, const int* upd_idx, const uint8_t* upd_val) {
    const int n = *upd_count;
    if (blockIdx.x == 0 && threadIdx.x == 0) *upd_count_next = 0;
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x) quality[upd_idx[k]] = upd_val[k];
}

// Slab-local set_state: every slab holds only its own fish; this pass keeps them where they are, drops stale ghosts
// and packs the boundary columns as the neighbours' halos (the same message a move() sends, without migrants).
struct HaloRefreshArgs {
    const float4* rec; const int4* cold1; const int* cell; const int* cell_start;
    int* next_cell; int* next_rank; int* next_count;
    Grid g;
    SlabMsg out_left, out_right;
};
__global__ void __launch_bounds__(256) halo_refresh_kernel(const HaloRefreshArgs a) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.cell_start[a.g.ncell]) return;
    if (i < a.cell_start[a.g.own_c0] || i >= a.cell_start[a.g.own_c1]) { a.next_cell[i] = -1; return; }
    const int c = a.cell[i], cx = c / a.g.ny, cy = c % a.g.ny;
    a.next_cell[i] = c;
    a.next_rank[i] = atomicAdd(&a.next_count[c], 1);
    if (cx == 1) slab_push_ghost(a.out_left, a.rec[i], a.cold1[i].w, cy);
    if (cx == a.g.lnx - 2) slab_push_ghost(a.out_right, a.rec[i], a.cold1[i].w, cy);
}

// After the exchange: the two received messages become unsorted entries behind the fish already held
// (left neighbour's migrants -> my first owned column, its ghosts -> my left halo; mirrored for the right).
struct IncomingArgs {
    SlabMsg in_left, in_right;
    float4* rec; float4* cold0; int4* cold1;
    int* next_cell; int* next_rank; int* next_count;
    const int* cell_start;   // cell_start[ncell] = entries held before the exchange
    int* n_unsorted;         // out: entries the scatter has to look at
    int* err;
    int cap_total;
    Grid g;
};
__global__ void __launch_bounds__(256) ingest_incoming_kernel(const IncomingArgs a) {
    const SlabHeader hl = *a.in_left.hdr(), hr = *a.in_right.hdr();
    const int lm = min(hl.n_mig, a.in_left.cap_m), lg = min(hl.n_ghost, a.in_left.cap_g);
    const int rm = min(hr.n_mig, a.in_right.cap_m), rg = min(hr.n_ghost, a.in_right.cap_g);
    const int base = a.cell_start[a.g.ncell];
    const int total = lm + lg + rm + rg;
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        *a.n_unsorted = min(base + total, a.cap_total);
        if (hl.overflow || hr.overflow || base + total > a.cap_total) atomicOr(a.err, ERR_SLAB_OVERFLOW);
    }
    const int ny = a.g.ny;
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < total; k += gridDim.x * blockDim.x) {
        const int i = base + k;
        if (i >= a.cap_total) break;
        float4 r, c0 = make_float4(0.f, 0.f, 1.f, 0.f);
        int4 c1;
        int lcx, cy;
        if (k < lm + lg) {
            if (k < lm) { const MigRec& m = a.in_left.mig()[k]; r = m.rec; c0 = m.cold0; c1 = m.cold1; cy = m.cy; lcx = 1; }
            else { const GhostRec& m = a.in_left.ghost()[k - lm]; r = m.rec; c1 = make_int4(0x7FFFFFFF, 0, 0, m.id); cy = m.cy; lcx = 0; }
        } else {
            const int kk = k - lm - lg;
            if (kk < rm) { const MigRec& m = a.in_right.mig()[kk]; r = m.rec; c0 = m.cold0; c1 = m.cold1; cy = m.cy; lcx = a.g.lnx - 2; }
            else { const GhostRec& m = a.in_right.ghost()[kk - rm]; r = m.rec; c1 = make_int4(0x7FFFFFFF, 0, 0, m.id); cy = m.cy; lcx = a.g.lnx - 1; }
        }
        const int c = lcx * ny + cy;
        a.rec[i] = r; a.cold0[i] = c0; a.cold1[i] = c1;
        a.next_cell[i] = c;
        a.next_rank[i] = atomicAdd(&a.next_count[c], 1);
    }
}

// ------------------------------------------------------------------------------------------------------
// K1 (ingest), K2 (scan), K3 (scatter)
// ------------------------------------------------------------------------------------------------------
struct IngestArgs {
    const double* coords; const double* dir; const double* speed; const double* sf;
    const int* lag; const int* rate; const int* intake;
    float4* rec; float4* cold0; int4* cold1;
    int* next_cell; int* next_rank; int* next_count;
    int* n_unsorted;  // entry counter (atomic); a fish of a slab's edge columns can be held twice (owned + ghost)
    const int* ids;   // null: fish i has id i (whole-school arrays); else the arrays are this slab's own fish only
    int n_global, cap_total;
    Grid g;
    int n_school;      // fish in the whole school (slab-local input: ids are validated against it)
    int* bad;  // bits: 1 a coordinate is outside [0,W] x [0,H] or not finite, 2 the arrays overflow, 4 a fish of a slab-local input
               // lies outside the slab's columns, 8 one of its ids is outside [0, n_school), 16 its speed can outrun the halo
    double w, h;
};

__device__ __forceinline__ void ingest_entry(const IngestArgs& a, int k, int lcx, int cy, const float4& r, const float4& c0, const int4& c1) {
    if (k < 0) k = claim_slot(a.n_unsorted);
    if (k >= a.cap_total) { atomicOr(a.bad, 2); return; }
    const int c = lcx * a.g.ny + cy;
    a.rec[k] = r; a.cold0[k] = c0; a.cold1[k] = c1;
    a.next_cell[k] = c;
    a.next_rank[k] = atomicAdd(&a.next_count[c], 1);
}

__global__ void ingest_kernel(const IngestArgs a) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.n_global) return;
    double x = a.coords[2 * (size_t)i], y = a.coords[2 * (size_t)i + 1];
    if (!(x >= 0.0 && x <= a.w && y >= 0.0 && y <= a.h)) { atomicOr(a.bad, 1); x = 0; y = 0; }
    if (x >= a.w) x -= a.w;  // x == W is kept by the reference's strict compare (fish_mod.cpp:468); here it wraps
    if (y >= a.h) y -= a.h;
    int cx = (int)floor(x / CELL_I), cy = (int)floor(y / CELL_I);
    float ox = (float)(x - (double)cx * CELL_I), oy = (float)(y - (double)cy * CELL_I);
    if (ox >= CELL_F) { ox -= CELL_F; cx += 1; }
    if (oy >= CELL_F) { oy -= CELL_F; cy += 1; }
    if (cx >= a.g.nx_g) cx -= a.g.nx_g;
    if (cy >= a.g.ny) cy -= a.g.ny;
    const float dirf = (float)a.dir[i];
    float sn, cs;
    heading_vector(dirf, cs, sn);
    const int id = a.ids ? a.ids[i] : i;
    if (a.ids) {   // slab-local input (fishabm_set_state_owned): validated here, while it is being read anyway, and reported by fishabm_wait
        // ids index pos_of_id[] on the device and the caller's arrays on the host: an entry with an id outside the school is dropped
        if ((unsigned)id >= (unsigned)a.n_school) { atomicOr(a.bad, 8); return; }
        // a fish may not outrun its halo: |step| = speed * speed_factor <= 3 * speed must stay below one cell
        if (!(a.speed[i] * 3.0 < (double)CELL_I)) atomicOr(a.bad, 16);
    }
    const float4 r = make_float4(ox, oy, cs, sn);
    const float4 c0 = make_float4(dirf, (float)a.speed[i], (float)a.sf[i], 0.f);
    const int4 c1 = make_int4(a.lag[i], a.rate[i], a.intake[i], id);
    if (a.g.wrap_x) { ingest_entry(a, i, cx, cy, r, c0, c1); return; }  // whole school: entry i, n_unsorted preset to n
    // slab: local column of the fish's global column, as an owned entry and/or as a ghost of either halo column
    int lo = cx - a.g.gx0 - 1; if (lo < 0) lo += a.g.nx_g;  // columns past the first owned one
    lo += 1;
    if (lo <= a.g.lnx - 2) ingest_entry(a, -1, lo, cy, r, c0, c1);
    else if (a.ids) atomicOr(a.bad, 4);  // slab-local input: the fish is not in this slab's columns
    if (a.ids) return;                   // halos come from the neighbours (halo_refresh_kernel + exchange)
    const int4 g1 = make_int4(0x7FFFFFFF, 0, 0, i);
    if (cx == a.g.gx0) ingest_entry(a, -1, 0, cy, r, c0, g1);
    if (cx == (a.g.gx0 + a.g.lnx - 1) % a.g.nx_g) ingest_entry(a, -1, a.g.lnx - 1, cy, r, c0, g1);
}

constexpr int SCAN_THREADS = 256;
constexpr int SCAN_ITEMS = 8;
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;

__device__ __forceinline__ int block_exclusive_scan(int v, int* total, int* s_warp) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    int inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { int t = __shfl_up_sync(0xFFFFFFFFu, inc, o); if (lane >= o) inc += t; }
    if (lane == 31) s_warp[w] = inc;
    __syncthreads();
    if (w == 0) {
        int x = (lane < (int)(blockDim.x >> 5)) ? s_warp[lane] : 0;
        int xi = x;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { int t = __shfl_up_sync(0xFFFFFFFFu, xi, o); if (lane >= o) xi += t; }
        s_warp[lane] = xi - x;
        if (lane == 31) s_warp[32] = xi;
    }
    __syncthreads();
    const int res = inc - v + s_warp[w];
    *total = s_warp[32];
    __syncthreads();
    return res;
}

// pass 1: per-tile sums
__global__ void __launch_bounds__(SCAN_THREADS) scan_tile_sums_kernel(const int* __restrict__ count, int n, int* __restrict__ tile_sums) {
    __shared__ int s_warp[33];
    const int base = blockIdx.x * SCAN_TILE + threadIdx.x * SCAN_ITEMS;
    int v = 0;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k) if (base + k < n) v += count[base + k];
    int total;
    block_exclusive_scan(v, &total, s_warp);
    if (threadIdx.x == 0) tile_sums[blockIdx.x] = total;
}
// pass 2: exclusive scan of the tile sums (single block, any length)
__global__ void __launch_bounds__(SCAN_THREADS) scan_tile_offsets_kernel(int* tile_sums, int ntiles) {
    __shared__ int s_warp[33];
    int carry = 0;
    for (int b = 0; b < ntiles; b += SCAN_THREADS) {
        const int i = b + threadIdx.x;
        const int v = i < ntiles ? tile_sums[i] : 0;
        int total;
        const int ex = block_exclusive_scan(v, &total, s_warp);
        if (i < ntiles) tile_sums[i] = carry + ex;
        carry += total;
    }
}
// pass 3: final exclusive scan; clears the counters for the next build; writes start[n] = total
__global__ void __launch_bounds__(SCAN_THREADS) scan_finish_kernel(int* __restrict__ count, int n, const int* __restrict__ tile_sums,
                                                                   int* __restrict__ start) {
    __shared__ int s_warp[33];
    const int base = blockIdx.x * SCAN_TILE + threadIdx.x * SCAN_ITEMS;
    int vals[SCAN_ITEMS];
    int v = 0;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k) { vals[k] = (base + k < n) ? count[base + k] : 0; v += vals[k]; }
    int total;
    int run = block_exclusive_scan(v, &total, s_warp) + tile_sums[blockIdx.x];
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k)
        if (base + k < n) { start[base + k] = run; run += vals[k]; count[base + k] = 0; }
    if (base <= n - 1 && n - 1 < base + SCAN_ITEMS) start[n] = run;
}

// the three passes in ONE launch, for grids whose tiles are all co-resident (the host checks): every block publishes
// its tile sum and takes a ticket; the last block to arrive scans the tile sums and raises the epoch flag; the others
// wait for it and finish their tile.  Saves two launches (~7 us) per step.
struct ScanFusedArgs { int* count; int n; int* tile_sums; int* start; int* ticket; int* ready; int epoch; int ntiles; };
__global__ void __launch_bounds__(SCAN_THREADS) scan_fused_kernel(const ScanFusedArgs a) {
    __shared__ int s_warp[33];
    __shared__ int s_last;
    const int base = blockIdx.x * SCAN_TILE + threadIdx.x * SCAN_ITEMS;
    int vals[SCAN_ITEMS];
    int v = 0;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k) { vals[k] = (base + k < a.n) ? a.count[base + k] : 0; v += vals[k]; }
    int total;
    const int ex = block_exclusive_scan(v, &total, s_warp);
    if (threadIdx.x == 0) {
        a.tile_sums[blockIdx.x] = total;
        __threadfence();
        s_last = (atomicAdd(a.ticket, 1) == a.ntiles - 1) ? 1 : 0;
    }
    __syncthreads();
    if (s_last) {  // exclusive scan of the tile sums (any length), in place
        __threadfence();
        int carry = 0;
        for (int b = 0; b < a.ntiles; b += SCAN_THREADS) {
            const int i = b + threadIdx.x;
            const int t = i < a.ntiles ? *reinterpret_cast<volatile int*>(&a.tile_sums[i]) : 0;
            int tot;
            const int e = block_exclusive_scan(t, &tot, s_warp);
            if (i < a.ntiles) a.tile_sums[i] = carry + e;
            carry += tot;
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            *a.ticket = 0;
            __threadfence();
            *reinterpret_cast<volatile int*>(a.ready) = a.epoch;
        }
    }
    if (threadIdx.x == 0) {
        while (*reinterpret_cast<volatile int*>(a.ready) != a.epoch) __nanosleep(40);
        __threadfence();
    }
    __syncthreads();
    int run = ex + *reinterpret_cast<volatile int*>(&a.tile_sums[blockIdx.x]);
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k)
        if (base + k < a.n) { a.start[base + k] = run; run += vals[k]; a.count[base + k] = 0; }
    if (base <= a.n - 1 && a.n - 1 < base + SCAN_ITEMS) a.start[a.n] = run;
}

// The same scan as ONE ordinary launch that needs no co-residency (default): a chained scan with decoupled look-back.  A block
// takes its tile number from a ticket counter, so a tile only ever waits for tiles whose blocks are already running; it
// publishes its tile sum, then a warp looks back over the published words of the preceding tiles -- 32 at a time, summing
// tile sums until it meets a tile that already knows its inclusive prefix -- and publishes its own prefix.  A word carries
// (epoch << 2 | state) in its high half and the value in its low half, written and read as one 64-bit access; the epoch makes
// last step's words invalid without a reset.
struct ScanChainArgs { int* count; int n; int* start; unsigned long long* desc; int* ticket; unsigned epoch; int ntiles; };
__global__ void __launch_bounds__(SCAN_THREADS) scan_chained_kernel(const ScanChainArgs a) {
    __shared__ int s_warp[33];
    __shared__ int s_tile, s_prefix;
    if (threadIdx.x == 0) s_tile = atomicAdd(a.ticket, 1);
    __syncthreads();
    const int tile = s_tile;
    const int base = tile * SCAN_TILE + threadIdx.x * SCAN_ITEMS;
    int vals[SCAN_ITEMS];
    int v = 0;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k) { vals[k] = (base + k < a.n) ? a.count[base + k] : 0; v += vals[k]; }
    int total;
    const int ex = block_exclusive_scan(v, &total, s_warp);
    const unsigned long long tag_sum = (unsigned long long)((a.epoch << 2) | 1u) << 32, tag_prefix = (unsigned long long)((a.epoch << 2) | 2u) << 32;
    volatile unsigned long long* desc = a.desc;
    if (threadIdx.x == 0) {
        desc[tile] = (tile == 0 ? tag_prefix : tag_sum) | (unsigned)total;
        if (tile == 0) s_prefix = 0;
        if (tile == a.ntiles - 1) *a.ticket = 0;   // every ticket of this launch has been taken: ready for the next one
    }
    if (tile > 0 && threadIdx.x < 32) {
        const int lane = threadIdx.x;
        int prefix = 0;
        for (int look = tile - 1;; look -= 32) {
            const int idx = look - lane;
            unsigned long long d = 0;
            if (idx >= 0) do { d = desc[idx]; } while ((unsigned)(d >> 34) != a.epoch);   // published in THIS scan
            const bool is_prefix = idx >= 0 && ((unsigned)(d >> 32) & 3u) == 2u;
            const unsigned pm = __ballot_sync(0xFFFFFFFFu, is_prefix);
            const int first = pm ? __ffs((int)pm) - 1 : 31;   // tiles up to the nearest one with a known prefix count
            prefix += __reduce_add_sync(0xFFFFFFFFu, (idx >= 0 && lane <= first) ? (int)(unsigned)d : 0);
            if (pm) break;   // (tile 0 always publishes a prefix: the walk ends there at the latest)
        }
        if (lane == 0) { s_prefix = prefix; desc[tile] = tag_prefix | (unsigned)(prefix + total); }
    }
    __syncthreads();
    int run = ex + s_prefix;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k)
        if (base + k < a.n) { a.start[base + k] = run; run += vals[k]; a.count[base + k] = 0; }
    if (base <= a.n - 1 && a.n - 1 < base + SCAN_ITEMS) a.start[a.n] = run;
}

struct ScatterArgs {
    const float4* rec_in; const float4* cold0_in; const int4* cold1_in;
    float4* rec_out; float4* cold0_out; int4* cold1_out; int* cell_out;
    const int* next_cell; const int* next_rank; const int* start;
    const int* n_unsorted; int cap;
    // the deferred food-grid writes of the sample() that preceded this sort ride along (one launch less per step); null: none
    uint8_t* quality; const int* upd_count; int* upd_count_next; const int* upd_idx; const uint8_t* upd_val;
};
constexpr int SCATTER_APPLY_BLOCKS = 64;
__global__ void __launch_bounds__(256) scatter_kernel(const ScatterArgs a) {
    if (a.quality && blockIdx.x < SCATTER_APPLY_BLOCKS) {   // = apply_quality_updates_kernel (nothing in this kernel reads the grid)
        const int n = *a.upd_count;
        if (blockIdx.x == 0 && threadIdx.x == 0) *a.upd_count_next = 0;
        for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += SCATTER_APPLY_BLOCKS * blockDim.x) a.quality[a.upd_idx[k]] = a.upd_val[k];
    }
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= min(*a.n_unsorted, a.cap)) return;  // claims beyond the capacity were refused (and reported)
    // the records are requested together with the keys (not behind the ghost test and the gather of the cell's start): the
    // kernel is bound by memory latency, and this takes one round trip out of its dependency chain
    const int c = a.next_cell[i];
    const int rank = a.next_rank[i];
    const float4 r = a.rec_in[i], c0 = a.cold0_in[i];
    const int4 c1 = a.cold1_in[i];
    if (c < 0) return;  // last step's ghost
    const int dst = a.start[c] + rank;
    a.rec_out[dst] = r;
    a.cold0_out[dst] = c0;
    a.cold1_out[dst] = c1;
    a.cell_out[dst] = c;
}

// ------------------------------------------------------------------------------------------------------
// export / generic queries.  Owned fish only.  pos = id (whole school: the caller's arrays are in id order) or
// the fish's index inside the owned range plus an id list (slab: the host scatters the compact arrays by id).
// ------------------------------------------------------------------------------------------------------
struct ExportArgs {
    const float4* rec; const float4* cold0; const int4* cold1; const int* cell; const int* cell_start;
    double* coords; double* dir; double* speed; double* sf; int* lag; int* rate; int* intake; int* ids;
    Grid g;
    int compact;
};
__global__ void copy_int_kernel(const int* src, int* dst) { *dst = *src; }
__global__ void export_kernel(const ExportArgs a) {
    const int own_b = a.cell_start[a.g.own_c0];
    const int i = own_b + blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.cell_start[a.g.own_c1]) return;
    const float4 r = a.rec[i], c0 = a.cold0[i];
    const int4 c1 = a.cold1[i];
    const int c = a.cell[i];
    const size_t pos = a.compact ? (size_t)(i - own_b) : (size_t)c1.w;
    if (a.ids) a.ids[pos] = c1.w;
    if (a.coords) {
        const int gcx = (a.g.gx0 + c / a.g.ny) % a.g.nx_g;
        a.coords[2 * pos] = (double)gcx * CELL_I + (double)r.x;
        a.coords[2 * pos + 1] = (double)(c % a.g.ny) * CELL_I + (double)r.y;
    }
    if (a.dir) a.dir[pos] = (double)c0.x;
    if (a.speed) a.speed[pos] = (double)c0.y;
    if (a.sf) a.sf[pos] = (double)c0.z;
    if (a.lag) a.lag[pos] = c1.x;
    if (a.rate) a.rate[pos] = c1.y;
    if (a.intake) a.intake[pos] = c1.z;
}

// per-pass outputs from sorted order to id (or compact) order
struct ExportPassArgs {
    const int4* cold1; const int4* cnt; const float* turn; const uint32_t* aux; const int* cell_start;
    int* cnt4; double* turn_out; int* n_avoid; int* n_align; int* n_attract; double* speed_out; int* ids;
    Grid g;
    int compact; double speed_norm;
};
__global__ void export_pass_kernel(const ExportPassArgs a) {
    const int own_b = a.cell_start[a.g.own_c0];
    const int i = own_b + blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.cell_start[a.g.own_c1]) return;
    const int id = a.cold1[i].w;
    const size_t pos = a.compact ? (size_t)(i - own_b) : (size_t)id;
    const int4 c = a.cnt[i];
    if (a.ids) a.ids[pos] = id;
    if (a.cnt4) { a.cnt4[4 * pos] = c.x; a.cnt4[4 * pos + 1] = c.y; a.cnt4[4 * pos + 2] = c.z; a.cnt4[4 * pos + 3] = c.w; }
    if (a.turn_out) a.turn_out[pos] = (double)a.turn[i];
    if (a.n_avoid) a.n_avoid[pos] = c.x;
    if (a.n_align) a.n_align[pos] = c.y - c.x;
    if (a.n_attract) a.n_attract[pos] = c.z - c.y;
    if (a.speed_out) a.speed_out[pos] = 1.0 + 2.0 * (double)c.w / a.speed_norm;
}

// in_zone(inds,id,fov,r) for arbitrary fov / r <= 200: literal fp64 predicate over the 3x3 stencil
struct ZoneArgs {
    const float4* rec; const float4* cold0; const int4* cold1; const int* cell; const int* cell_start;
    int* out; int* ids; Grid g; int compact, fov_half; double r;
};
__global__ void in_zone_generic_kernel(const ZoneArgs a) {
    const int own_b = a.cell_start[a.g.own_c0];
    const int i = own_b + blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.cell_start[a.g.own_c1]) return;
    const float4 me = a.rec[i];
    const int c = a.cell[i];
    double cs, sn;
    unit_vector_fp64(a.cold0[i].x, cs, sn);
    int count = 0;
    for (int k = 0; k < 9; ++k) {
        const int cc = stencil_cell(a.g, c, k);
        const double sx = (double)((k % 3) - 1) * 200.0, sy = (double)((k / 3) - 1) * 200.0;
        for (int j = a.cell_start[cc]; j < a.cell_start[cc + 1]; ++j) {
            if (j == i) continue;
            const float4 o = a.rec[j];
            const double x = __dadd_rn(__dsub_rn((double)o.x, (double)me.x), sx);
            const double y = __dadd_rn(__dsub_rn((double)o.y, (double)me.y), sy);
            const double dist = __dsqrt_rn(__dadd_rn(__dmul_rn(x, x), __dmul_rn(y, y)));
            const double q = __ddiv_rn(__dadd_rn(__dmul_rn(cs, x), __dmul_rn(sn, y)), dist);
            const double angle = __ddiv_rn(__dmul_rn(acos(q), 180.0), M_PI);
            if (angle < (double)a.fov_half && dist < a.r) count++;
        }
    }
    const int id = a.cold1[i].w;
    const size_t pos = a.compact ? (size_t)(i - own_b) : (size_t)id;
    a.out[pos] = count;
    if (a.ids) a.ids[pos] = id;
}

// ------------------------------------------------------------------------------------------------------
// fp64 check mode (FISHABM_MODE_BRUTE_FP64's evaluator, fishabm_check_fp64): in_zone counts and the social() turn of
// listed fish by brute force over ALL fish of the school, in double precision, restating the reference's own formulas
// (in_zone fish_mod.cpp:226-260, social :262-377, averageturn :379-408, getangle :420-464) -- nothing shared with the
// product path: no cell list, no fp32 pair math, no polynomial atan2, no fixed-point sums.  One warp per listed fish,
// candidates across the lanes (so sums are reduced in a different order than the reference's id order: ~1e-13 deg).
// ------------------------------------------------------------------------------------------------------
struct CheckArgs {
    const float4* rec; const float4* cold0; const int4* cold1; const int* cell; const int* pos_of_id;
    const int* ids; int k;
    int* counts4; double* turn; int* flags;
    int n; Grid g; int w, h;
    uint32_t step; uint64_t seed; uint32_t replicate;
    Model model;
    double w_align, w_attract;   // the combination weights in double precision (Model keeps fp32 copies for the fp32 path)
};
enum { CHECK_RANDOMWALK = 1, CHECK_TIES = 2, CHECK_KNIFE = 4 };

__device__ __forceinline__ double check_fold(double d, int extent) {  // :239-250: int halves, strict compares
    if (d > extent / 2) return d - extent;
    if (d < -extent / 2) return d + extent;
    return d;
}
// getangle (:420-464) for an offset (x,y) already folded; *tie: the reference would draw distrib_choice (:460-462)
__device__ __forceinline__ double check_getangle(double x, double y, double fdir, bool* tie) {
    double angle = acos(x / sqrt(x * x + y * y)) * 180.0 / M_PI;
    if (y < 0) angle = 360.0 - angle;
    else if (y == 0) angle = (x >= 0) ? 0.0 : 180.0;
    angle = angle - fdir;
    if (angle < 0) angle += 360.0; else if (angle >= 360.0) angle -= 360.0;  // correctangle, once
    *tie = false;
    if (angle > 180.0) angle -= 360.0;
    else if (angle == 180.0) *tie = true;
    return angle;
}
__device__ __forceinline__ double check_averageturn(double xs, double ys, int count) {  // :379-408 from the cos/sin sums
    const double x = xs / count, y = ys / count;
    double angle = acos(x / sqrt(x * x + y * y)) * 180.0 / M_PI;
    if (y < 0) angle = -angle;
    else if (y == 0) angle = (x >= 0) ? 0.0 : 180.0;
    return angle;
}
__device__ __forceinline__ double warp_sum(double v) {
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, o);
    return v;
}

// the fp64 brute-force evaluation of ONE focal fish (sorted index fi, global id `id`) by one warp: every lane returns the
// reduced counts, the turn and the flags
struct CheckOut { int c15, c100, c200, cf; double turn; int flags; };
__device__ __forceinline__ CheckOut check_eval(const CheckArgs& a, int fi, int id, int lane) {
    const int ny = a.g.ny;
    auto world = [&](int i, double& X, double& Y) {  // exact: cell * 200 + fp32 offset
        const float4 r = a.rec[i];
        const int c = a.cell[i];
        X = (double)((a.g.gx0 + c / ny) % a.g.nx_g) * 200.0 + (double)r.x;
        Y = (double)(c % ny) * 200.0 + (double)r.y;
    };
    double fx, fy;
    world(fi, fx, fy);
    const double fdir = (double)a.cold0[fi].x;
    const double dx0 = cos(fdir * M_PI / 180.0), dy0 = sin(fdir * M_PI / 180.0);
    int c15 = 0, c100 = 0, c200 = 0, cf = 0, ties = 0, knife = 0;
    double avx = 0, avy = 0, alx = 0, aly = 0, atx = 0, aty = 0;
    for (int j = lane; j < a.n; j += 32) {
        if (j == fi) continue;
        double X, Y;
        world(j, X, Y);
        const double x = check_fold(X - fx, a.w), y = check_fold(Y - fy, a.h);
        if (!(x * x + y * y < 40000.5)) continue;  // nothing below can hold beyond r = 200
        const double dist = sqrt(x * x + y * y);
        const double angle = acos((dx0 * x + dy0 * y) / dist) * 180.0 / M_PI;  // NaN (coincident / rounding) fails every test
        if (angle < 90.0 && dist < 200.0) cf++;                                 // get_speed's in_zone(180, 200), :589
        if (!(angle < 165.0 && dist < 200.0)) continue;                         // :254, :315 (fov/2 = 165)
        c200++; if (dist < 100.0) c100++; if (dist < 15.0) c15++;
        const uint32_t nid = (uint32_t)a.cold1[j].w;
        bool tie;
        double t;
        if (dist < 15.0) {            // avoid, :316-332
            double b = check_getangle(x, y, fdir, &tie);
            if (tie) { b = (fishrng::draw(a.seed, a.replicate, a.step, (uint32_t)id, STREAM_CHOICE, nid, 0, 1) - 0.5) * 360.0; ties++; }
            if (fabs(b) < 1e-4 || fabs(fabs(b) - 180.0) < 1e-4) knife++;
            if (b > 0) t = b - 180.0;
            else if (b < 0) t = b + 180.0;
            else { t = b + (fishrng::draw(a.seed, a.replicate, a.step, (uint32_t)id, STREAM_CHOICE, nid, 0, 1) - 0.5) * 360.0; ties++; }
            t = t * (1.0 / (0.1 * (dist * dist) + 2.0));
            avx += cos(t * M_PI / 180.0); avy += sin(t * M_PI / 180.0);
        } else if (dist < 100.0) {    // align, :333-344: the neighbour's heading vector added to the focal position
            const double jd = (double)a.cold0[j].x;
            const double px = cos(jd * M_PI / 180.0) + fx, py = sin(jd * M_PI / 180.0) + fy;
            double b = check_getangle(check_fold(px - fx, a.w), check_fold(py - fy, a.h), fdir, &tie);
            if (tie) { b = (fishrng::draw(a.seed, a.replicate, a.step, (uint32_t)id, STREAM_CHOICE, nid, 0, 1) - 0.5) * 360.0; ties++; }
            if (fabs(fabs(b) - 180.0) < 1e-4) knife++;
            t = b * (1.0 / (0.004 * ((dist - 15.0) * (dist - 15.0)) + 2.0));
            alx += cos(t * M_PI / 180.0); aly += sin(t * M_PI / 180.0);
        } else {                      // attract, :345-351
            double b = check_getangle(x, y, fdir, &tie);
            if (tie) { b = (fishrng::draw(a.seed, a.replicate, a.step, (uint32_t)id, STREAM_CHOICE, nid, 0, 1) - 0.5) * 360.0; ties++; }
            if (fabs(fabs(b) - 180.0) < 1e-4) knife++;
            t = b * (1.0 / (0.004 * ((dist - 200.0) * (dist - 200.0)) + 2.0));
            atx += cos(t * M_PI / 180.0); aty += sin(t * M_PI / 180.0);
        }
    }
    c15 = __reduce_add_sync(0xFFFFFFFFu, c15); c100 = __reduce_add_sync(0xFFFFFFFFu, c100);
    c200 = __reduce_add_sync(0xFFFFFFFFu, c200); cf = __reduce_add_sync(0xFFFFFFFFu, cf);
    ties = __reduce_add_sync(0xFFFFFFFFu, ties); knife = __reduce_add_sync(0xFFFFFFFFu, knife);
    avx = warp_sum(avx); avy = warp_sum(avy); alx = warp_sum(alx); aly = warp_sum(aly); atx = warp_sum(atx); aty = warp_sum(aty);
    const int n_av = c15, n_al = c100 - c15, n_at = c200 - c100;
    double turn;
    int fl = (ties ? CHECK_TIES : 0) | (knife ? CHECK_KNIFE : 0);
    if (n_av > 0) turn = check_averageturn(avx, avy, n_av);             // :356-358
    else if (n_al > 0 && n_at > 0) {                                    // :360-365: weights 1.7 / 0.7
        const double A = check_averageturn(alx, aly, n_al), T = check_averageturn(atx, aty, n_at);
        const double xs = cos(A * M_PI / 180.0) * a.w_align + cos(T * M_PI / 180.0) * a.w_attract;
        const double ys = sin(A * M_PI / 180.0) * a.w_align + sin(T * M_PI / 180.0) * a.w_attract;
        turn = check_averageturn(xs, ys, 2);
    } else if (n_al > 0) turn = check_averageturn(alx, aly, n_al);
    else if (n_at > 0) turn = check_averageturn(atx, aty, n_at);
    else { turn = (double)fishrng::draw(a.seed, a.replicate, a.step, (uint32_t)id, STREAM_RANDOMWALK, 0, -a.model.randomwalk_range, a.model.randomwalk_range); fl |= CHECK_RANDOMWALK; }  // :372-374
    CheckOut o;
    o.c15 = c15; o.c100 = c100; o.c200 = c200; o.cf = cf; o.turn = turn; o.flags = fl;
    return o;
}

__global__ void __launch_bounds__(256) check_fp64_kernel(const CheckArgs a) {
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp >= a.k) return;
    const int id = a.ids[warp];
    const int fi = (id >= 0 && id < a.n) ? a.pos_of_id[id] : -1;
    if (fi < 0) { if (lane == 0) { a.turn[warp] = nan(""); a.flags[warp] = -1; for (int q = 0; q < 4; ++q) a.counts4[4 * warp + q] = -1; } return; }
    const CheckOut o = check_eval(a, fi, id, lane);
    if (lane != 0) return;
    a.counts4[4 * warp] = o.c15; a.counts4[4 * warp + 1] = o.c100; a.counts4[4 * warp + 2] = o.c200; a.counts4[4 * warp + 3] = o.cf;
    a.turn[warp] = o.turn;
    a.flags[warp] = o.flags;
}

// FISHABM_MODE_BRUTE_FP64: the neighbour pass of a whole (small) school as the fp64 brute-force evaluation above -- a warp
// per fish, every other fish a candidate, no cell list, no fp32 pair arithmetic -- writing the same per-pass outputs
// (cnt / turn / aux, sorted order) the update kernel consumes.  O(n^2); the "fp64 check mode" of BASELINE.json north_star
// as a stepping mode.
struct BrutePassArgs {
    CheckArgs c;             // c.step = the move's step (tie / random-walk draws); ids / pos_of_id / counts4 / turn / flags unused
    int4* cnt; float* turn; uint32_t* aux;
    int active_lag_max, do_sample;
    uint32_t step_sample;
};
__global__ void __launch_bounds__(256) brute_fp64_pass_kernel(const BrutePassArgs a) {
    const int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (i >= a.c.n) return;
    const int4 k1 = a.c.cold1[i];
    if (!(a.active_lag_max < 0 || k1.x <= a.active_lag_max)) { if (lane == 0) a.aux[i] = 0u; return; }
    const CheckOut o = check_eval(a.c, i, k1.w, lane);
    if (lane != 0) return;
    uint32_t aux = AUX_EVALUATED | ((o.flags & CHECK_RANDOMWALK) ? AUX_RANDOMWALK : 0u);
    const float4 r = a.c.rec[i];
    if (a.do_sample) aux |= sample_decision(a.c.model, a.c.seed, a.c.replicate, a.step_sample, (uint32_t)k1.w, k1.x, k1.y, r.x, r.y, o.c15, o.c100 - o.c15);
    a.cnt[i] = make_int4(o.c15, o.c100, o.c200, o.cf);
    a.turn[i] = (float)o.turn;
    a.aux[i] = aux;
}

__global__ void sum_quality_kernel(const uint8_t* q, size_t n, unsigned long long* out) {
    unsigned long long s = 0;
    for (size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += (size_t)gridDim.x * blockDim.x) s += q[k];
    for (int o = 16; o > 0; o >>= 1) s += __shfl_down_sync(0xFFFFFFFFu, s, o);
    if ((threadIdx.x & 31) == 0 && s) atomicAdd(out, s);
}

// feed(inds,id,...) for an explicit id list, sequentially (fish_mod.cpp:560-574); API completeness, not hot.
// Ids this context does not own are skipped.
struct FeedArgs {
    float4* cold0; int4* cold1; const float4* rec; const int* cell; const int* ids; int k;
    uint8_t* quality; int n_global, qcols; Grid g; int* pos_of_id;
    Model model;
};
__global__ void index_by_id_kernel(const int4* cold1, const int* cell_start, Grid g, int* pos_of_id) {
    const int i = cell_start[g.own_c0] + blockIdx.x * blockDim.x + threadIdx.x;
    if (i < cell_start[g.own_c1]) pos_of_id[cold1[i].w] = i;
}
__global__ void feed_list_kernel(const FeedArgs a) {
    if (blockIdx.x != 0 || threadIdx.x != 0) return;
    for (int t = 0; t < a.k; ++t) {
        const int id = a.ids[t];
        if (id < 0 || id >= a.n_global) continue;
        const int i = a.pos_of_id[id];
        if (i < 0) continue;
        const float4 r = a.rec[i];
        const int c = a.cell[i];
        const int gx = (c / a.g.ny - a.g.own_c0 / a.g.ny) * FOOD_PER_CELL + min((int)r.x / FOOD_I, FOOD_PER_CELL - 1);
        const int gy = (c % a.g.ny) * FOOD_PER_CELL + min((int)r.y / FOOD_I, FOOD_PER_CELL - 1);
        const int qi = gy * a.qcols + gx;
        const int q = a.quality[qi];
        float4 c0 = a.cold0[i];
        int4 c1 = a.cold1[i];
        if (q > 0) { c0.z = a.model.fed_sf; a.quality[qi] = (uint8_t)(q - 1); c1.z += 1; c1.y = a.model.rate_fed; }
        else c1.y = a.model.rate_unfed;
        a.cold0[i] = c0; a.cold1[i] = c1;
    }
}

}  // namespace fishk
