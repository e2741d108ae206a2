// K8: replicate ensembles (BASELINE.json configs[4]: 65,536 schools of N=200).  One CTA owns one school for the whole
// launch: the school's state and its food grid live in shared memory, every step is a neighbour pass over all pairs of the
// school (a warp per focal fish: a cheap in-range filter over all fish compacts the few that can interact into a queue,
// the pair math runs on those, 32 at a time) followed by the per-fish update, and only the logged
// observables (sum of the food grid per step, food_intake every k steps: the reference's commented-out loggers,
// fish_mod.cpp:132-141,149-154) touch HBM in between.  Replicates never communicate; the per-pair arithmetic, the
// per-focal epilogue and the move arithmetic are the SAME device functions the cell-list path uses (pair_math,
// focal_result, move_core), and pair sums are integers, so replicate r is bit-identical to a single school created
// with the same seed and replicate id.
#pragma once
#include "kernels.cuh"

namespace fishk {

constexpr int ENS_THREADS = 256;
constexpr int ENS_MAX_FISH = 1024;  // lane partial sums and their warp totals stay inside int32

enum { ENS_OP_STEPS = 0, ENS_OP_MOVE = 1, ENS_OP_SAMPLE = 2, ENS_OP_QUERY = 3 };

struct EnsembleArgs {
    float4* rec;      // [R][n]  in-cell offset + heading unit vector
    float4* cold0;    // [R][n]  dir, speed, speed_factor, -
    int4* cold1;      // [R][n]  lag, sample_rate, food_intake, cell (cx << 16 | cy)
    uint8_t* quality; // [R][qrows*qcols]
    int n, nx, ny, qcols, qcells;
    int replicates;
    uint32_t replicate0;
    int op, nsteps;
    uint32_t step0;
    int rolling;
    uint64_t seed;
    double speed_norm;
    long long* sumq;  // [nsteps][R] or null (ENS_OP_STEPS)
    int* intake;      // [rows][R][n] or null
    int intake_every;
    int4* cnt; float* turn; uint32_t* aux;  // [R][n] outputs of ENS_OP_QUERY
    const Model* models;  // [R]: every replicate has its own model constants (parameter sweeps, BASELINE.json configs[4])
};

__host__ __device__ inline int ens_queue_stride(int n) { return (n + 32 + 7) & ~7; }   // survivor queue of one warp, 16-bit entries
__host__ __device__ inline size_t ensemble_smem_bytes(int n, int qcells) {
    return (size_t)n * (16 + 16 + 16 + 16 + 4 + 4 + 4 + 4) + (size_t)((qcells + 15) / 16) * 16 + 64 +
           (size_t)(ENS_THREADS / 32) * ens_queue_stride(n) * 2;
}
__device__ __forceinline__ int ens_cell_pack(int cx, int cy) { return (cx << 16) | cy; }

__global__ void __launch_bounds__(ENS_THREADS, 4) ensemble_kernel(const EnsembleArgs a) {
    extern __shared__ __align__(16) unsigned char ens_smem[];
    const int n = a.n, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const unsigned FULL = 0xFFFFFFFFu;
    float4* s_rec = reinterpret_cast<float4*>(ens_smem);
    float4* s_c0 = s_rec + n;
    int4* s_c1 = reinterpret_cast<int4*>(s_c0 + n);
    int4* s_cnt = s_c1 + n;
    float* s_turn = reinterpret_cast<float*>(s_cnt + n);
    uint32_t* s_aux = reinterpret_cast<uint32_t*>(s_turn + n);
    int* s_updq = reinterpret_cast<int*>(s_aux + n);
    int* s_updv = s_updq + n;
    uint8_t* s_q = reinterpret_cast<uint8_t*>(s_updv + n);
    int* s_total = reinterpret_cast<int*>(s_q + ((a.qcells + 15) / 16) * 16);
    int* s_dens = s_total + 4;   // [0] interacting pairs and [1] focal fish of the running pass, [2] the next pass goes direct
    unsigned short* s_queue = reinterpret_cast<unsigned short*>(s_total + 16) + (size_t)warp * ens_queue_stride(n);

    const int rep = blockIdx.x;
    const uint32_t replicate = a.replicate0 + (uint32_t)rep;
    __shared__ Model model;   // this replicate's model constants
    if (tid == 0) model = a.models[rep];
    const size_t base = (size_t)rep * n;
    for (int i = tid; i < n; i += ENS_THREADS) { s_rec[i] = a.rec[base + i]; s_c0[i] = a.cold0[base + i]; s_c1[i] = a.cold1[base + i]; }
    const uint8_t* gq = a.quality + (size_t)rep * a.qcells;
    if (tid == 0) { *s_total = 0; s_dens[0] = 0; s_dens[1] = 0; s_dens[2] = 0; }
    __syncthreads();
    {
        int part = 0;
        for (int k = tid; k < a.qcells; k += ENS_THREADS) { const uint8_t v = gq[k]; s_q[k] = v; part += v; }
        part = __reduce_add_sync(FULL, part);
        if (lane == 0 && part) atomicAdd(s_total, part);
    }
    __syncthreads();

    // ---- neighbour pass: a warp per focal fish ------------------------------------------------------------------------
    // candidate j seen from focal fish f: the periodic minimum image in whole cells, then the SAME frame arithmetic as the
    // cell-list kernels (kernels.cuh: yframe_shift)
    auto image = [&](int cj, int cxf, int cyf, int& dcx, int& dcy) {
        dcx = (cj >> 16) - cxf; dcy = (cj & 0xFFFF) - cyf;
        if (dcx > 1) dcx -= a.nx; else if (dcx < -1) dcx += a.nx;
        if (dcy > 1) dcy -= a.ny; else if (dcy < -1) dcy += a.ny;
        return dcx >= -1 && dcx <= 1 && dcy >= -1 && dcy <= 1;
    };
    const unsigned lt_mask = (1u << lane) - 1u;
    // Two ways through the school's fish, bit-identical in their results (integer sums of the same per-pair arithmetic):
    //   filtered  every fish against the interaction range first, in the round's own fp32 arithmetic (the same dx, dy, d2 bits, so
    //             the cut can sit right behind the 200-unit band); the survivors are compacted into the warp's queue and only
    //             they go through the pair rounds -- a dispersed school has a handful per fish;
    //   direct    the rounds take all fish, 32 at a time -- a cohesive school (the reference's lattice start stays one:
    //             BASELINE.json configs[4]) has nearly every pair in range, and the filter would only add to the work.
    // Which one a pass uses follows the interacting fraction the PREVIOUS pass of this school counted (s_dens).
    // The shift that takes candidate j into focal fish f's frame depends only on the two fish's cells.  In a world of at most
    // 32 x 32 cells (6400 units; the default school lives in 10 x 10) every lane holds the shift of ONE cell column and ONE cell
    // row of the focal fish -- (column - cxf) * 200 wrapped to the nearest image, or a far-away value when the image is not
    // adjacent -- and a pair costs two shuffles instead of the wrap arithmetic.  Larger worlds compute it per pair.
    const bool lut_ok = a.nx <= 32 && a.ny <= 32;
    auto pass = [&](int active_lag_max, int do_sample, uint32_t step_sample, uint32_t step_move) {
        const bool direct = s_dens[2] != 0;
        for (int f = warp; f < n; f += ENS_THREADS / 32) {
            const int4 k1 = s_c1[f];
            const bool active = active_lag_max < 0 || k1.x <= active_lag_max;
            if (!active) { if (lane == 0) s_aux[f] = 0u; continue; }
            const float4 F = s_rec[f];
            const float fdir = s_c0[f].x;
            const bool exotic = fabsf(fdir) >= 540.f;
            const int cxf = k1.w >> 16, cyf = k1.w & 0xFFFF;
            const float ysh = yframe_shift(cyf), fyf = F.y + ysh;   // the pair arithmetic's y frame (kernels.cuh)
            float lutx = 1.0e18f, luty = 3.0e17f;                   // this lane's cell column / row seen from the focal fish
            if (lut_ok) {
                int dcx = lane - cxf, dcy = lane - cyf;
                if (dcx > 1) dcx -= a.nx; else if (dcx < -1) dcx += a.nx;
                if (dcy > 1) dcy -= a.ny; else if (dcy < -1) dcy += a.ny;
                if (lane < a.nx && dcx >= -1 && dcx <= 1) lutx = (float)dcx * CELL_F;
                if (lane < a.ny && dcy >= -1 && dcy <= 1) luty = (float)dcy * CELL_F + ysh;
            }
            // the candidate's shift into the focal frame: (sx, sy), far away unless its cell image is adjacent
            auto shift_of = [&](int cj, float& sx, float& sy) {
                if (lut_ok) {
                    sx = __shfl_sync(FULL, lutx, cj >> 16);
                    sy = __shfl_sync(FULL, luty, cj & 0xFFFF);
                } else {
                    int dcx, dcy;
                    const bool inr = image(cj, cxf, cyf, dcx, dcy);
                    sx = inr ? (float)dcx * CELL_F : 1.0e18f;
                    sy = inr ? (float)dcy * CELL_F + ysh : 3.0e17f;
                }
            };
            int count = n;   // entries the rounds go through: all fish (direct) or the queue's survivors
            if (!direct) {
                int qn = 0;
                __syncwarp();
                for (int j0 = 0; j0 < n; j0 += 32) {
                    const int j = min(j0 + lane, n - 1);
                    const float2 q = *reinterpret_cast<const float2*>(&s_rec[j]);
                    float sx, sy;
                    shift_of(s_c1[j].w, sx, sy);
                    const float dx = (q.x + sx) - F.x, dy = (q.y + sy) - fyf;
                    const bool keep = j0 + lane < n && fmaf(dx, dx, dy * dy) < (CELL_F + 2.f * BAND_R) * (CELL_F + 2.f * BAND_R);
                    const unsigned m = __ballot_sync(FULL, keep);
                    if (keep) s_queue[qn + __popc(m & lt_mask)] = (unsigned short)j;
                    qn += __popc(m);
                }
                __syncwarp();
                count = qn;
            }
            PairAccR P;
            pair_acc_clear(P);
            // pass 0: the branch-free round (pair_fast) on 32 entries at a time.  A lane beyond the last entry takes the focal
            // fish itself: a pair with d2 == 0 fails every predicate (the reference's NaN drop), like a candidate whose shift is
            // the far-away value.  A borderline pair, a tie or an exotic heading anywhere (rare) discards the sums and pass 1
            // redoes the fish with the generic per-lane code.
            unsigned dirty = exotic ? 1u : 0u;
            for (int pass_no = exotic ? 1 : 0; pass_no < 2; ++pass_no) {
                for (int q0 = 0; q0 < count; q0 += 32) {
                    const int e = q0 + lane;
                    const int jj = e < count ? (direct ? e : (int)s_queue[e]) : f;
                    const float4 Q = s_rec[jj];
                    const int cj = s_c1[jj].w;
                    if (pass_no == 0) {
                        float sx, sy;
                        shift_of(cj, sx, sy);
                        if (pair_fast<false>(F.x, fyf, F.z, F.w, Q.x + sx, Q.y + sy, Q.z, Q.w, P)) dirty = 1u;
                    } else {
                        int dcx, dcy;
                        const bool live = image(cj, cxf, cyf, dcx, dcy) && e < count;
                        const float qx = live ? Q.x + (float)dcx * CELL_F : 1.0e18f, qy = live ? Q.y + ((float)dcy * CELL_F + ysh) : 3.0e17f;
                        auto raw = [&](float& rx, float& ry, int& kcode, float&, float& fyr) { rx = Q.x; ry = Q.y; kcode = (dcy + 1) * 3 + (dcx + 1); fyr = F.y; };
                        auto ids = [&](uint32_t& my, uint32_t& nb) { my = (uint32_t)f; nb = (uint32_t)jj; };
                        pair_round_generic<false>(live, F.x, fyf, F.z, F.w, fdir, exotic, qx, qy, Q.z, Q.w, a.seed, replicate, step_move, raw, ids, P);
                    }
                }
                if (pass_no == 0) {
                    if (!__any_sync(FULL, dirty != 0u)) break;
                    pair_acc_clear(P);
                }
            }
            const unsigned red[6] = {__reduce_add_sync(FULL, P.ax0), __reduce_add_sync(FULL, P.ay0), __reduce_add_sync(FULL, P.ax1),
                                     __reduce_add_sync(FULL, P.ay1), __reduce_add_sync(FULL, P.ax2), __reduce_add_sync(FULL, P.ay2)};
            const unsigned cfar = __reduce_add_sync(FULL, P.c_far), cnear = __reduce_add_sync(FULL, P.c_near);
            const int4 cc = unpack_counts_r<false>(cfar, cnear);
            long long sums[6];
            pair_sums_unbias(red, cc, sums);
            const int c15 = cc.x, c100 = cc.y, c200 = cc.z, cfr = cc.w, ties = 0;
            if (lane == 0) {
                const FocalResult r = focal_result(model, a.seed, replicate, step_sample, step_move, do_sample, (uint32_t)f, k1.x, k1.y, F.x, F.y,
                                                   c15, c100, c200, cfr, sums, ties);
                s_cnt[f] = r.cnt; s_turn[f] = r.turn; s_aux[f] = r.aux;
                atomicAdd(&s_dens[0], c200); atomicAdd(&s_dens[1], 1);
            }
        }
        __syncthreads();
        if (tid == 0) {   // the next pass goes direct when more than half of all pairs interacted in this one
            s_dens[2] = (2 * s_dens[0] > s_dens[1] * n) ? 1 : 0;
            s_dens[0] = 0; s_dens[1] = 0;
        }
        __syncthreads();
    };

    // ---- per-fish update: sample() / feed() / move(), the same transition as update_kernel ---------------------------
    auto update = [&](int do_sample, int do_move, uint32_t step_move) {
        for (int i = tid; i < n; i += ENS_THREADS) {
            float4 c0 = s_c0[i];
            int4 c1 = s_c1[i];
            const uint32_t aux = do_sample ? s_aux[i] : 0u;
            s_updq[i] = -1;
            if (do_sample) {  // sample(), fish_mod.cpp:546-556
                if (aux & AUX_WANTS) {
                    c1.x = model.lag_after_feed;
                    // feed(), fish_mod.cpp:560-574: feeders of one food cell are served lowest id first
                    const uint32_t key = aux & ((AUX_FOOD_MASK << AUX_FOOD_SHIFT) | AUX_WANTS);
                    int rank = 0, total = 0;
                    for (int j = 0; j < n; ++j) {
                        if ((s_aux[j] & ((AUX_FOOD_MASK << AUX_FOOD_SHIFT) | AUX_WANTS)) == key && s_c1[j].w == c1.w) {
                            total++;
                            rank += (j < i) ? 1 : 0;
                        }
                    }
                    const int fl = (int)((aux >> AUX_FOOD_SHIFT) & AUX_FOOD_MASK);
                    const int gx = (c1.w >> 16) * FOOD_PER_CELL + fl % FOOD_PER_CELL;
                    const int gy = (c1.w & 0xFFFF) * FOOD_PER_CELL + fl / FOOD_PER_CELL;
                    const int qi = gy * a.qcols + gx;
                    const int q = (int)s_q[qi];
                    if (rank < q) { c0.z = model.fed_sf; c1.z += 1; c1.y = model.rate_fed; atomicSub(s_total, 1); }
                    else c1.y = model.rate_unfed;
                    if (rank == 0 && q > 0) { s_updq[i] = qi; s_updv[i] = q - min(q, total); }  // written after every feeder has read
                } else if (c1.x == 0) c1.y += 1;
                else c1.x -= 1;
            }
            if (do_move) {  // move(), fish_mod.cpp:165-179
                const int err = fishrng::draw(a.seed, replicate, step_move, (uint32_t)i, STREAM_ERROR, 0, -model.error_range, model.error_range);
                float4 r = s_rec[i];
                int cx = c1.w >> 16, cy = c1.w & 0xFFFF;
                const bool moving = c1.x == 0;
                move_core(r, c0, c1.x, cx, cy, moving ? s_turn[i] : 0.f, moving ? s_cnt[i].w : 0, err, a.rolling, (float)(2.0 / a.speed_norm));
                cx %= a.nx; if (cx < 0) cx += a.nx;
                cy %= a.ny; if (cy < 0) cy += a.ny;
                c1.w = ens_cell_pack(cx, cy);
                s_rec[i] = r;
            }
            s_c0[i] = c0;
            // the cell (c1.w) of OTHER fish is read by the feeding-rank loop above: publish c1 after the barrier
            s_cnt[i] = c1;
        }
        __syncthreads();
        for (int i = tid; i < n; i += ENS_THREADS) {
            s_c1[i] = s_cnt[i];
            if (s_updq[i] >= 0) s_q[s_updq[i]] = (uint8_t)s_updv[i];
        }
        __syncthreads();
    };

    auto log_after_sample = [&](int z) {  // observables after sample() of launch-local step z
        if (a.sumq && tid == 0) a.sumq[(size_t)z * a.replicates + rep] = (long long)*s_total;
        if (a.intake && a.intake_every > 0 && z % a.intake_every == 0) {
            int* row = a.intake + ((size_t)(z / a.intake_every) * a.replicates + rep) * n;
            for (int i = tid; i < n; i += ENS_THREADS) row[i] = s_c1[i].z;
        }
    };

    if (a.op == ENS_OP_STEPS) {
        // move(s); { sample(s+t-1) + move(s+t) sharing one pass } ...; the last sample() completes inside the launch
        for (int t = 0; t < a.nsteps; ++t) {
            const uint32_t s = a.step0 + (uint32_t)t;
            const int pend = t > 0;
            pass(pend ? 1 : 0, pend, s - 1, s);
            update(pend, 1, s);
            if (pend) log_after_sample(t - 1);
        }
        const uint32_t sl = a.step0 + (uint32_t)a.nsteps - 1;
        pass(0, 1, sl, sl);
        update(1, 0, sl);
        log_after_sample(a.nsteps - 1);
    } else if (a.op == ENS_OP_MOVE) {
        pass(0, 0, a.step0, a.step0);
        update(0, 1, a.step0);
    } else if (a.op == ENS_OP_SAMPLE) {
        pass(0, 1, a.step0, a.step0);
        update(1, 0, a.step0);
    } else {  // ENS_OP_QUERY: frozen-state outputs for every fish
        pass(-1, 0, a.step0, a.step0);
        for (int i = tid; i < n; i += ENS_THREADS) { a.cnt[base + i] = s_cnt[i]; a.turn[base + i] = s_turn[i]; a.aux[base + i] = s_aux[i]; }
        return;
    }
    for (int i = tid; i < n; i += ENS_THREADS) { a.rec[base + i] = s_rec[i]; a.cold0[base + i] = s_c0[i]; a.cold1[base + i] = s_c1[i]; }
    uint8_t* gqo = a.quality + (size_t)rep * a.qcells;
    for (int k = tid; k < a.qcells; k += ENS_THREADS) gqo[k] = s_q[k];
}

// host arrays (reference layout, replicate-major) <-> ensemble arrays
struct EnsIoArgs {
    double* coords; double* dir; double* speed; double* sf; int* lag; int* rate; int* intake;
    float4* rec; float4* cold0; int4* cold1;
    long long total; int nx, ny; int* bad; double w, h;
};
__global__ void ens_ingest_kernel(const EnsIoArgs a) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.total) return;
    double x = a.coords[2 * i], y = a.coords[2 * i + 1];
    if (!(x >= 0.0 && x <= a.w && y >= 0.0 && y <= a.h)) { atomicOr(a.bad, 1); x = 0; y = 0; }
    if (x >= a.w) x -= a.w;
    if (y >= a.h) y -= a.h;
    int cx = (int)floor(x / CELL_I), cy = (int)floor(y / CELL_I);
    float ox = (float)(x - (double)cx * CELL_I), oy = (float)(y - (double)cy * CELL_I);
    if (ox >= CELL_F) { ox -= CELL_F; cx += 1; }
    if (oy >= CELL_F) { oy -= CELL_F; cy += 1; }
    if (cx >= a.nx) cx -= a.nx;
    if (cy >= a.ny) cy -= a.ny;
    const float dirf = (float)a.dir[i];
    float sn, cs;
    heading_vector(dirf, cs, sn);
    a.rec[i] = make_float4(ox, oy, cs, sn);
    a.cold0[i] = make_float4(dirf, (float)a.speed[i], (float)a.sf[i], 0.f);
    a.cold1[i] = make_int4(a.lag[i], a.rate[i], a.intake[i], ens_cell_pack(cx, cy));
}
__global__ void ens_export_kernel(const EnsIoArgs a) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.total) return;
    const float4 r = a.rec[i], c0 = a.cold0[i];
    const int4 c1 = a.cold1[i];
    if (a.coords) {
        a.coords[2 * i] = (double)(c1.w >> 16) * CELL_I + (double)r.x;
        a.coords[2 * i + 1] = (double)(c1.w & 0xFFFF) * CELL_I + (double)r.y;
    }
    if (a.dir) a.dir[i] = (double)c0.x;
    if (a.speed) a.speed[i] = (double)c0.y;
    if (a.sf) a.sf[i] = (double)c0.z;
    if (a.lag) a.lag[i] = c1.x;
    if (a.rate) a.rate[i] = c1.y;
    if (a.intake) a.intake[i] = c1.z;
}

struct EnsPassOutArgs {
    const int4* cnt; const float* turn; long long total; double speed_norm;
    int* cnt4; double* turn_out; int* n_avoid; int* n_align; int* n_attract; double* speed_out;
};
__global__ void ens_export_pass_kernel(const EnsPassOutArgs a) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.total) return;
    const int4 c = a.cnt[i];
    if (a.cnt4) { a.cnt4[4 * i] = c.x; a.cnt4[4 * i + 1] = c.y; a.cnt4[4 * i + 2] = c.z; a.cnt4[4 * i + 3] = c.w; }
    if (a.turn_out) a.turn_out[i] = (double)a.turn[i];
    if (a.n_avoid) a.n_avoid[i] = c.x;
    if (a.n_align) a.n_align[i] = c.y - c.x;
    if (a.n_attract) a.n_attract[i] = c.z - c.y;
    if (a.speed_out) a.speed_out[i] = 1.0 + 2.0 * (double)c.w / a.speed_norm;
}

// per-replicate sums of the food grids
__global__ void ens_sum_quality_kernel(const uint8_t* q, int qcells, long long* out) {
    __shared__ int s_part;
    if (threadIdx.x == 0) s_part = 0;
    __syncthreads();
    const uint8_t* g = q + (size_t)blockIdx.x * qcells;
    int part = 0;
    for (int k = threadIdx.x; k < qcells; k += blockDim.x) part += g[k];
    part = __reduce_add_sync(0xFFFFFFFFu, part);
    if ((threadIdx.x & 31) == 0 && part) atomicAdd(&s_part, part);
    __syncthreads();
    if (threadIdx.x == 0) out[blockIdx.x] = s_part;
}

__global__ void ens_broadcast_quality_kernel(const uint8_t* one, uint8_t* all, int qcells, long long total) {
    for (long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x; k < total; k += (long long)gridDim.x * blockDim.x) all[k] = one[k % qcells];
}

}  // namespace fishk
