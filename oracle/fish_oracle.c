/* oracle/fish_oracle.c -- TEST INFRASTRUCTURE (parity oracle), not product code.
 *
 * A plain-C, fp64, runtime-sized restatement of the per-step agent update of
 * fritzfrancisco/fish_abm (move() + sample() and everything they call).  Each function cites the
 * reference lines it follows.  It exists so that the CUDA path can be checked at sizes and world
 * geometries the verbatim reference (oracle/_ref/libfishref_<num>.so, compile-time capacity) does
 * not cover, and on the GPU box where /root/reference is absent.
 *
 * PINNING: tests/test_oracle_vs_ref.py checks every function below against the compiled,
 * unmodified reference on identical frozen states and identical random draws (bit-for-bit: both
 * sides are glibc fp64 with the same operation order), and tests/golden/ holds reference-generated
 * vectors so the same check runs where the reference cannot be built.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this library.  The product (libfishabm.so) never links or calls it.
 *
 * Two update semantics are provided:
 *   *_seq  : the reference's own sequential, in-place, id-ordered loops (fish_mod.cpp:164,541);
 *   *_sync : the synchronous (frozen-state) form the GPU implements: every fish reads the state as
 *            it was before the call.  Per fish the arithmetic is the reference's; the documented
 *            differences are DESIGN.md D1-D3 (neighbours pre-move; get_speed at the pre-move pose).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#ifdef _OPENMP
#include <omp.h>
#endif
#include "orc_rng.h"

/* the reference's literals as parameters (defaults = the reference's values; the product's fishabm_model_params) */
typedef struct {
    int lag_after_feed;   /* 10, fish_mod.cpp:548 */
    int rate_fed;         /* 30, :570 */
    int rate_unfed;       /* 0,  :573 */
    int sample_draw_max;  /* 35, :33 */
    int error_range;      /* 10, :29 */
    int randomwalk_range; /* 45, :30 */
    int alone_align_min;  /* 5,  :546 */
    int pad;
    double fed_sf;        /* 0.5, :566 */
    double w_align;       /* 1.7, :385-386 (1.7 - i, i = 0) */
    double w_attract;     /* 1.7 - 1 */
} orc_model;

typedef struct {
    int n;             /* fish (inds.n) */
    int w, h;          /* Structure.w / .h, ints as in fish_mod.cpp:58-59 */
    int rows, cols;    /* Structure.rows / .cols (food grid) */
    int q_stride;      /* row stride of quality[][] (n_cols, fish_mod.cpp:24) */
    double speed_norm; /* the macro `num` in get_speed (fish_mod.cpp:590) */
    uint64_t seed;
    uint32_t replicate;
    uint32_t pad_;
    double* coords; /* [n][2] */
    double* dir;
    double* speed;
    double* speed_factor;
    int* lag;
    int* sample_rate;
    int* food_intake;
    int* quality; /* [rows][q_stride] */
    orc_model m;
} orc_school;

void orc_model_default(orc_model* m) {
    m->lag_after_feed = 10; m->rate_fed = 30; m->rate_unfed = 0; m->sample_draw_max = 35; m->error_range = 10;
    m->randomwalk_range = 45; m->alone_align_min = 5; m->pad = 0; m->fed_sf = 0.5; m->w_align = 1.7; m->w_attract = 1.7 - 1;
}

/* ---- scalar helpers ----------------------------------------------------------------------- */

/* correctangle, fish_mod.cpp:410-418: one +-360 wrap */
double orc_correctangle(double dir) {
    if (dir < 0) return dir + 360;
    if (dir >= 360) return dir - 360;
    return dir;
}

/* correct_coords, fish_mod.cpp:466-480: one +-w / +-h wrap, `> w` strict */
void orc_correct_coords(double* xy, int w, int h) {
    if (xy[0] > w) xy[0] -= w; else if (xy[0] < 0) xy[0] += w;
    if (xy[1] > h) xy[1] -= h; else if (xy[1] < 0) xy[1] += h;
}

/* minimum-image fold used at fish_mod.cpp:239-250, 298-310, 426-438: int halves, strict compares */
static inline double fold(double d, int extent) {
    if (d > extent / 2) return d - extent;
    if (d < -extent / 2) return d + extent;
    return d;
}

/* averageturn, fish_mod.cpp:379-408, from the running sums of cos/sin (same summation order as the
 * reference's loop over the vector) */
static double averageturn_sums(double x_sum, double y_sum, int count) {
    double x = x_sum / count, y = y_sum / count;
    double angle = acos(x / sqrt(x * x + y * y)) * 180 / M_PI;
    if (y < 0) angle = 0 - angle;
    else if (y == 0) angle = (x >= 0) ? 0 : 180;
    return angle;
}

double orc_averageturn(const double* v, int n, int comb_weight) {
    double xs = 0, ys = 0;
    for (int i = 0; i < n; ++i) {
        double c = cos(v[i] * M_PI / 180), s = sin(v[i] * M_PI / 180);
        if (comb_weight) { xs = xs + c * (1.7 - i); ys = ys + s * (1.7 - i); } /* :385-386 */
        else { xs = xs + c; ys = ys + s; }                                     /* :389-390 */
    }
    return averageturn_sums(xs, ys, n);
}

/* getangle, fish_mod.cpp:420-464.  *tie180 is set when the reference would draw distrib_choice
 * (:460-462); the caller resolves the draw. */
static double getangle_core(double fx, double fy, double fdir, double px, double py, int w, int h, int* tie180) {
    double x = fold(px - fx, w), y = fold(py - fy, h);
    double angle = acos(x / sqrt(x * x + y * y)) * 180 / M_PI;
    if (y < 0) angle = 360 - angle;
    else if (y == 0) angle = (x >= 0) ? 0 : 180;
    angle = angle - fdir;
    double na = orc_correctangle(angle);
    *tie180 = 0;
    if (na > 180) na = na - 360;
    else if (na == 180) *tie180 = 1;
    return na;
}

double orc_getangle(const double* focal_xy, double focal_dir, const double* pos_xy, int w, int h,
                    uint64_t seed, uint32_t replicate, uint32_t step, uint32_t id, uint32_t k) {
    int tie;
    double a = getangle_core(focal_xy[0], focal_xy[1], focal_dir, pos_xy[0], pos_xy[1], w, h, &tie);
    if (tie) a = (orc_draw_from(seed, replicate, step, id, ORC_STREAM_CHOICE, k, 0, 1) - 0.5) * 360;
    return a;
}

/* ---- neighbour queries --------------------------------------------------------------------- */

/* in_zone, fish_mod.cpp:226-260 */
int orc_in_zone(const orc_school* S, int id, int fov, double r) {
    int c = 0;
    double dx0 = cos(S->dir[id] * M_PI / 180), dy0 = sin(S->dir[id] * M_PI / 180);
    for (int i = 0; i < S->n; ++i) {
        if (i == id) continue;
        double x = fold(S->coords[2 * i] - S->coords[2 * id], S->w);
        double y = fold(S->coords[2 * i + 1] - S->coords[2 * id + 1], S->h);
        double dist = sqrt(x * x + y * y);
        double angle = acos((dx0 * x + dy0 * y) / dist) * 180 / M_PI;
        if (angle < (fov / 2) && dist < r) c++;
    }
    return c;
}

/* the four counts the step needs, one pass, same per-pair arithmetic as in_zone:
 * out = { in_zone(330,15), in_zone(330,100), in_zone(330,200), in_zone(180,200) } */
void orc_zone_counts(const orc_school* S, int id, int* out4) {
    int c15 = 0, c100 = 0, c200 = 0, cf = 0;
    double dx0 = cos(S->dir[id] * M_PI / 180), dy0 = sin(S->dir[id] * M_PI / 180);
    for (int i = 0; i < S->n; ++i) {
        if (i == id) continue;
        double x = fold(S->coords[2 * i] - S->coords[2 * id], S->w);
        double y = fold(S->coords[2 * i + 1] - S->coords[2 * id + 1], S->h);
        double dist = sqrt(x * x + y * y);
        if (!(dist < 200)) continue; /* every predicate below includes dist < r <= 200 */
        double angle = acos((dx0 * x + dy0 * y) / dist) * 180 / M_PI;
        if (angle < 165) { c200++; if (dist < 100) c100++; if (dist < 15) c15++; }
        if (angle < 90) cf++;
    }
    out4[0] = c15; out4[1] = c100; out4[2] = c200; out4[3] = cf;
}

typedef struct {
    double turn;
    int n_avoid, n_align, n_attract; /* as fish_mod.cpp:273-275 */
    int n_ties;                      /* distrib_choice draws consumed (:328, :461) */
    int randomwalk;                  /* 1 if the turn came from distrib_randomwalk (:373) */
    int knife;                       /* neighbours whose bearing sits on a discontinuity of the turn rule to within
                                        1e-4 deg (avoid zone: bearing ~ 0; any zone: |bearing| ~ 180): the reference's own
                                        answer then hinges on fp64 rounding noise, so parity tests skip such fish */
} orc_social_out;

/* social, fish_mod.cpp:262-377.  Turns are accumulated as running cos/sin sums in candidate order,
 * which is the order averageturn() later walks the reference's vectors in, so results agree bit for
 * bit.  Tie-break draws are addressed by the neighbour's id (k = i). */
void orc_social(const orc_school* S, int id, int fov, uint32_t step, orc_social_out* out) {
    const double fx = S->coords[2 * id], fy = S->coords[2 * id + 1], fdir = S->dir[id];
    const double dx0 = cos(fdir * M_PI / 180), dy0 = sin(fdir * M_PI / 180);
    int n_av = 0, n_al = 0, n_at = 0, ties = 0, knife = 0;
    double av_x = 0, av_y = 0, al_x = 0, al_y = 0, at_x = 0, at_y = 0;
    for (int i = 0; i < S->n; ++i) {
        if (i == id) continue;
        double y = fold(S->coords[2 * i + 1] - fy, S->h);
        double x = fold(S->coords[2 * i] - fx, S->w);
        double dist = sqrt(x * x + y * y);
        double angle = acos((dx0 * x + dy0 * y) / dist) * 180 / M_PI;
        if (!(angle < (fov / 2) && dist < 200)) continue;
        int tie;
        if (dist < 15) { /* :316-332 */
            double a = getangle_core(fx, fy, fdir, S->coords[2 * i], S->coords[2 * i + 1], S->w, S->h, &tie);
            if (tie) { a = (orc_draw_from(S->seed, S->replicate, step, id, ORC_STREAM_CHOICE, i, 0, 1) - 0.5) * 360; ties++; }
            if (fabs(a) < 1e-4 || fabs(fabs(a) - 180) < 1e-4) knife++;
            double t;
            if (a > 0) t = a - 180;
            else if (a < 0) t = a + 180;
            else { t = a + (orc_draw_from(S->seed, S->replicate, step, id, ORC_STREAM_CHOICE, i, 0, 1) - 0.5) * 360; ties++; }
            t = t * (1 / (0.1 * (dist * dist) + 2));
            av_x = av_x + cos(t * M_PI / 180); av_y = av_y + sin(t * M_PI / 180);
            n_av++;
        } else if (dist < 100) { /* :333-344 */
            double p[2];
            p[0] = cos(S->dir[i] * M_PI / 180) + fx;
            p[1] = sin(S->dir[i] * M_PI / 180) + fy;
            double a = getangle_core(fx, fy, fdir, p[0], p[1], S->w, S->h, &tie);
            if (tie) { a = (orc_draw_from(S->seed, S->replicate, step, id, ORC_STREAM_CHOICE, i, 0, 1) - 0.5) * 360; ties++; }
            if (fabs(fabs(a) - 180) < 1e-4) knife++;
            double t = a * (1 / (0.004 * ((dist - 15) * (dist - 15)) + 2));
            al_x = al_x + cos(t * M_PI / 180); al_y = al_y + sin(t * M_PI / 180);
            n_al++;
        } else { /* :345-351 */
            double a = getangle_core(fx, fy, fdir, S->coords[2 * i], S->coords[2 * i + 1], S->w, S->h, &tie);
            if (tie) { a = (orc_draw_from(S->seed, S->replicate, step, id, ORC_STREAM_CHOICE, i, 0, 1) - 0.5) * 360; ties++; }
            if (fabs(fabs(a) - 180) < 1e-4) knife++;
            double t = a * (1 / (0.004 * ((dist - 200) * (dist - 200)) + 2));
            at_x = at_x + cos(t * M_PI / 180); at_y = at_y + sin(t * M_PI / 180);
            n_at++;
        }
    }
    double turn = 0;
    int rw = 0;
    if (n_av > 0) turn = averageturn_sums(av_x, av_y, n_av); /* :356-358 */
    else if (n_al > 0 && n_at > 0) {                          /* :360-365 */
        /* averageturn(comb_al_at, 1), :362-365: weights (1.7 - i) for i = 0, 1 = w_align, w_attract */
        double a0 = averageturn_sums(al_x, al_y, n_al), a1 = averageturn_sums(at_x, at_y, n_at);
        double xs = 0, ys = 0;
        xs = xs + cos(a0 * M_PI / 180) * S->m.w_align; ys = ys + sin(a0 * M_PI / 180) * S->m.w_align;
        xs = xs + cos(a1 * M_PI / 180) * S->m.w_attract; ys = ys + sin(a1 * M_PI / 180) * S->m.w_attract;
        turn = averageturn_sums(xs, ys, 2);
    } else if (n_al > 0) turn = averageturn_sums(al_x, al_y, n_al);
    else if (n_at > 0) turn = averageturn_sums(at_x, at_y, n_at);
    else { turn = orc_draw(S->seed, S->replicate, step, id, ORC_STREAM_RANDOMWALK, -S->m.randomwalk_range, S->m.randomwalk_range); rw = 1; } /* :372-374 */
    out->turn = turn; out->n_avoid = n_av; out->n_align = n_al; out->n_attract = n_at;
    out->n_ties = ties; out->randomwalk = rw; out->knife = knife;
}

/* ---- feed / sample ------------------------------------------------------------------------- */

/* feed, fish_mod.cpp:560-574 (the repaint at :576-585 is viewer-only) */
void orc_feed(orc_school* S, int id) {
    int x = (int)S->coords[2 * id], y = (int)S->coords[2 * id + 1];
    int* cell = &S->quality[(size_t)(y / (S->h / S->rows)) * S->q_stride + x / (S->w / S->cols)];
    if (*cell > 0) {
        S->speed_factor[id] = S->m.fed_sf;
        *cell = *cell - 1;
        S->food_intake[id] = S->food_intake[id] + 1;
        S->sample_rate[id] = S->m.rate_fed;
    } else {
        S->sample_rate[id] = S->m.rate_unfed;
    }
}

/* sample, fish_mod.cpp:540-558.  Positions and headings do not change inside sample(), so the id-
 * ordered loop IS the frozen-state semantics; only feed() is order dependent (lowest id first).
 * counts_or_null: optional [n][4] precomputed orc_zone_counts (identical values, saves passes). */
void orc_sample(orc_school* S, uint32_t step, const int* counts_or_null) {
    for (int id = 0; id < S->n; ++id) {
        int n_avoid, n_align;
        if (counts_or_null) { n_avoid = counts_or_null[4 * id]; n_align = counts_or_null[4 * id + 1] - n_avoid; }
        else { n_avoid = orc_in_zone(S, id, 330, 15); n_align = orc_in_zone(S, id, 330, 100) - n_avoid; }
        int alone = (n_avoid == 0 && n_align < S->m.alone_align_min);
        if (S->lag[id] == 0 &&
            S->sample_rate[id] > orc_draw(S->seed, S->replicate, step, id, ORC_STREAM_SAMPLE, 0, S->m.sample_draw_max) && alone == 0) {
            S->lag[id] = S->m.lag_after_feed;
            orc_feed(S, id);
        } else if (S->lag[id] == 0) {
            S->sample_rate[id] = S->sample_rate[id] + 1;
        } else {
            S->lag[id] = S->lag[id] - 1;
        }
    }
}

/* sum_array, fish_mod.cpp:593-601 */
/* create_environment, fish_mod.cpp:482-531 (dead code in the reference's main: distrib_seed(0,0) gives 0 seeds; here
 * the seed count is an argument).  Draws: the k-th draw of each stream in call order, fish id 0, step 0
 * (stream 7 = the function's local distrib_rows / distrib_cols, range 0..rows / 0..cols INCLUSIVE as written).
 * Seeds that land on row == rows or col == cols fall outside the grid the sweeps read and are dropped. */
void orc_create_environment(int* q, int rows, int cols, int stride, int n_seeds, uint64_t seed, uint32_t replicate) {
    uint32_t kq = 0, ke = 0;
    for (int u = 0; u < rows; ++u)
        for (int i = 0; i < cols; ++i) q[u * stride + i] = 0;
    for (int s = 0; s < n_seeds; ++s) {
        int r = orc_draw_from(seed, replicate, 0, 0, 7, ke++, 0, rows);
        int c = orc_draw_from(seed, replicate, 0, 0, 7, ke++, 0, cols);
        if (r < rows && c < cols) q[r * stride + c] = 10;
    }
    for (int p = 0; p < 20; ++p) {
        for (int u = 0; u < rows; ++u)
            for (int i = 0; i < cols; ++i)
                if (u < rows - 1 && i < cols - 1)
                    if (q[(u + 1) * stride + i] > p || q[(u + 1) * stride + i + 1] > p || q[u * stride + i + 1])
                        q[u * stride + i] = orc_draw_from(seed, replicate, 0, 0, ORC_STREAM_QUALITY, kq++, 0, 10);
        for (int u = rows - 1; u > 0; --u)
            for (int i = cols - 1; i > 0; --i)
                if (q[(u - 1) * stride + i] > p || q[(u - 1) * stride + i - 1] > p || q[u * stride + i - 1])
                    q[u * stride + i] = orc_draw_from(seed, replicate, 0, 0, ORC_STREAM_QUALITY, kq++, 0, 10);
    }
}

long orc_sum_quality(const orc_school* S) {
    long s = 0;
    for (int r = 0; r < S->rows; ++r) for (int c = 0; c < S->cols; ++c) s += S->quality[(size_t)r * S->q_stride + c];
    return s;
}

/* ---- move ------------------------------------------------------------------------------------ */

/* move, fish_mod.cpp:163-181, the reference's sequential in-place loop */
void orc_move_seq(orc_school* S, uint32_t step) {
    for (int id = 0; id < S->n; ++id) {
        int err = orc_draw(S->seed, S->replicate, step, id, ORC_STREAM_ERROR, -S->m.error_range, S->m.error_range);
        if (S->lag[id] == 0) {
            double dir = orc_correctangle(S->dir[id]);
            orc_social_out so;
            orc_social(S, id, 330, step, &so);
            double turn = so.turn + err;
            dir = dir + turn;
            S->coords[2 * id] = S->coords[2 * id] + S->speed[id] * cos(dir * M_PI / 180) * S->speed_factor[id];
            S->coords[2 * id + 1] = S->coords[2 * id + 1] + S->speed[id] * sin(dir * M_PI / 180) * S->speed_factor[id];
            int count = orc_in_zone(S, id, 180, 200); /* get_speed, :588-591: new position, old heading */
            S->speed_factor[id] = 1 + 2 * (double)count / S->speed_norm;
            orc_correct_coords(&S->coords[2 * id], S->w, S->h);
            S->dir[id] = dir;
        } else {
            S->dir[id] = S->dir[id] + err;
        }
    }
}

/* Synchronous move: per fish the lines of fish_mod.cpp:165-179 on the FROZEN pre-move state.
 * D1: neighbours are all pre-move.  D2: get_speed's count is in_zone(frozen, id, 180, 200), i.e.
 * the pre-move pose.  turn_or_null / counts_or_null: optional precomputed orc_social turns and
 * orc_zone_counts on the same frozen state. */
void orc_move_sync(orc_school* S, uint32_t step, const double* turn_or_null, const int* counts_or_null) {
    const int n = S->n;
    double* nc = (double*)malloc(sizeof(double) * 2 * n);
    double* nd = (double*)malloc(sizeof(double) * n);
    double* nf = (double*)malloc(sizeof(double) * n);
#pragma omp parallel for schedule(dynamic, 16)
    for (int id = 0; id < n; ++id) {
        int err = orc_draw(S->seed, S->replicate, step, id, ORC_STREAM_ERROR, -S->m.error_range, S->m.error_range);
        nc[2 * id] = S->coords[2 * id]; nc[2 * id + 1] = S->coords[2 * id + 1];
        nf[id] = S->speed_factor[id];
        if (S->lag[id] == 0) {
            double dir = orc_correctangle(S->dir[id]);
            double st;
            if (turn_or_null) st = turn_or_null[id];
            else { orc_social_out so; orc_social(S, id, 330, step, &so); st = so.turn; }
            double turn = st + err;
            dir = dir + turn;
            nc[2 * id] = S->coords[2 * id] + S->speed[id] * cos(dir * M_PI / 180) * S->speed_factor[id];
            nc[2 * id + 1] = S->coords[2 * id + 1] + S->speed[id] * sin(dir * M_PI / 180) * S->speed_factor[id];
            int count = counts_or_null ? counts_or_null[4 * id + 3] : orc_in_zone(S, id, 180, 200);
            nf[id] = 1 + 2 * (double)count / S->speed_norm;
            orc_correct_coords(&nc[2 * id], S->w, S->h);
            nd[id] = dir;
        } else {
            nd[id] = S->dir[id] + err;
        }
    }
    memcpy(S->coords, nc, sizeof(double) * 2 * n);
    memcpy(S->dir, nd, sizeof(double) * n);
    memcpy(S->speed_factor, nf, sizeof(double) * n);
    free(nc); free(nd); free(nf);
}

/* ---- whole-school helpers (OpenMP over focal ids; read-only on S) ------------------------------ */

void orc_zone_counts_all(const orc_school* S, int* out /* [n][4] */) {
#pragma omp parallel for schedule(dynamic, 16)
    for (int id = 0; id < S->n; ++id) orc_zone_counts(S, id, out + 4 * id);
}

void orc_zone_counts_ids(const orc_school* S, const int* ids, int k, int* out /* [k][4] */) {
#pragma omp parallel for schedule(dynamic, 4)
    for (int j = 0; j < k; ++j) orc_zone_counts(S, ids[j], out + 4 * j);
}

void orc_in_zone_all(const orc_school* S, int fov, double r, int* out) {
#pragma omp parallel for schedule(dynamic, 16)
    for (int id = 0; id < S->n; ++id) out[id] = orc_in_zone(S, id, fov, r);
}

void orc_social_all(const orc_school* S, uint32_t step, orc_social_out* out) {
#pragma omp parallel for schedule(dynamic, 16)
    for (int id = 0; id < S->n; ++id) orc_social(S, id, 330, step, out + id);
}

void orc_social_ids(const orc_school* S, uint32_t step, const int* ids, int k, orc_social_out* out) {
#pragma omp parallel for schedule(dynamic, 4)
    for (int j = 0; j < k; ++j) orc_social(S, ids[j], 330, step, out + j);
}

/* one synchronous step in the reference's call order: move (frozen) then sample */
void orc_step_sync(orc_school* S, uint32_t step) {
    const int n = S->n;
    int* counts = (int*)malloc(sizeof(int) * 4 * n);
    orc_zone_counts_all(S, counts);
    orc_move_sync(S, step, NULL, counts);
    orc_zone_counts_all(S, counts);
    orc_sample(S, step, counts);
    free(counts);
}

/* one reference-order sequential step */
void orc_step_seq(orc_school* S, uint32_t step) {
    orc_move_seq(S, step);
    orc_sample(S, step, NULL);
}

/* `steps` steps from first_step; logs sum_array per step (fish_mod.cpp:149-154) and food_intake every
 * intake_every steps (:132-141) when the pointers are non-null.  sync: 0 = *_seq, 1 = *_sync. */
void orc_run(orc_school* S, int steps, uint32_t first_step, int sync, long* sumq_or_null,
             int* intake_or_null, int intake_every) {
    int rows_logged = 0;
    for (int z = 0; z < steps; ++z) {
        if (sync) orc_step_sync(S, first_step + z); else orc_step_seq(S, first_step + z);
        if (intake_or_null && intake_every > 0 && z % intake_every == 0) {
            memcpy(intake_or_null + (size_t)rows_logged * S->n, S->food_intake, sizeof(int) * S->n);
            rows_logged++;
        }
        if (sumq_or_null) sumq_or_null[z] = orc_sum_quality(S);
    }
}

/* ---- timing leg for bench.py's cpu_baseline ("port") ---------------------------------------------
 * The reference's per-fish work of one step for k focal ids on the full state: for a moving fish (lag==0)
 * social + get_speed's in_zone + 3 in_zone, then sample's 3 in_zone = 8 all-pairs passes; 3 for a lagged fish
 * (fish_mod.cpp:165-179,543-545), each literally executed.  Returns seconds. */
static double now_s(void) {
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

double orc_time_focal(const orc_school* S, const int* ids, int k, int* threads_used) {
    double sink = 0;
    int nt = 1;
#ifdef _OPENMP
    nt = omp_get_max_threads();
#endif
    if (threads_used) *threads_used = nt;
    double t0 = now_s();
#pragma omp parallel for schedule(dynamic, 1) reduction(+ : sink)
    for (int j = 0; j < k; ++j) {
        int id = ids[j];
        double acc = 0;
        if (S->lag[id] == 0) { /* move(): fish_mod.cpp:165-175 */
            orc_social_out so;
            orc_social(S, id, 330, 0, &so); /* the pair loop of social, :293-354 */
            acc += so.turn;
            acc += orc_in_zone(S, id, 330, 15) + orc_in_zone(S, id, 330, 100) + orc_in_zone(S, id, 330, 200); /* social :273-275 */
            acc += orc_in_zone(S, id, 180, 200);                                                                /* get_speed :589 */
        }
        acc += orc_in_zone(S, id, 330, 15) + orc_in_zone(S, id, 330, 100) + orc_in_zone(S, id, 330, 200); /* sample :543-545 */
        sink += acc;
    }
    double t = now_s() - t0;
    if (sink == 12345.678) t += 1e-30; /* keep the work observable */
    return t;
}

int orc_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* ---- exported views of the inline RNG helpers (for tests) ------------------------------------------ */
void orc_philox_export(const uint32_t* ctr4, const uint32_t* key2, uint32_t* out4) { orc_philox4x32_10(ctr4, key2, out4); }
int orc_draw_export(uint64_t seed, uint32_t replicate, uint32_t step, uint32_t id, uint32_t stream, uint32_t k0, int a, int b) {
    return orc_draw_from(seed, replicate, step, id, stream, k0, a, b);
}
int orc_lemire_accept_export(uint32_t word, int a, int b, int* value) { return orc_lemire_accept(word, a, b, value); }
