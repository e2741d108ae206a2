// K4, second design: candidate-parallel neighbour kernel.
//
// One warp owns one focal cell.  The <= 256 candidates of a window batch (the fish of the 3x3 cell stencil, shifted
// into the focal cell's frame) sit one per lane in registers (8 chunks of 32).  For each ACTIVE focal fish of the
// cell (broadcast from shared memory):
//   filter  - every lane tests its candidates with a cheap, conservative, sqrt-free predicate (inside the r=200
//             disc and not clearly behind the 330 deg cone); survivors are compacted into a per-focal queue in
//             shared memory with ballot/popc;
//   process - the queue is consumed 32 entries at a time, one pair per lane, all lanes working for the SAME focal
//             fish: exact classification (with the fp64 re-check of borderline pairs), zone counts by ballot, and the
//             avoidance / alignment / attraction turn of the pair accumulated as fixed-point cos/sin in registers;
//   reduce  - six warp REDUX adds fold the lanes' partial sums into the focal fish's 64-bit accumulators.
// Lanes therefore stay full regardless of how many fish of the cell are active, and the expensive per-pair math only
// runs on (almost only) interacting pairs.  Sums are integers, hence independent of candidate order: results are
// bit-identical to the first design (neighbour_kernel) and across any re-sorting of the fish.
#pragma once
#include "kernels.cuh"

namespace fishk {

constexpr int V2_WARPS = 8;
constexpr int V2_THREADS = V2_WARPS * 32;
constexpr int V2_WIN = 256;           // candidates per window batch
constexpr int V2_CH = V2_WIN / 32;    // register chunks per lane
constexpr int V2_FMAX = 32;           // focal fish per group
constexpr float V2_FAR = 1.0e18f;     // padding coordinate: fails the disc test

struct V2WarpSmem {
    float wx[V2_WIN], wy[V2_WIN], wc[V2_WIN], ws[V2_WIN];  // window candidates in the focal cell's frame
    uint32_t wcode[V2_WIN];                                // sorted index | stencil code << 28
    long long acc[V2_FMAX][6];                             // fixed-point cos/sin sums: avoid x,y, align x,y, attract x,y
    int cnt[V2_FMAX][4];                                   // n15, n100, n200, front
    float4 frec[V2_FMAX];
    float fdir[V2_FMAX];
    int fidx[V2_FMAX];
    int fties[V2_FMAX];
    unsigned char queue[V2_WIN];
};

constexpr size_t V2_SMEM_BYTES = sizeof(V2WarpSmem) * V2_WARPS;

// conservative cone test constant: cos^2(165 deg) + margin (see the band analysis in DESIGN.md)
constexpr float V2_KCONE = 0.93301270189f + 2.5e-4f;
constexpr float V2_R2HI = (CELL_F + 2.5e-4f) * (CELL_F + 2.5e-4f);

template <bool STATS>
__global__ void __launch_bounds__(V2_THREADS, 3) neighbour_kernel_v2(const NeighArgs a) {
    extern __shared__ __align__(16) unsigned char v2_smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const unsigned FULL = 0xFFFFFFFFu;
    const unsigned lt_mask = (1u << lane) - 1u;
    V2WarpSmem& S = reinterpret_cast<V2WarpSmem*>(v2_smem_raw)[warp];

    unsigned long long st_cand = 0, st_inter = 0, st_rechk = 0, st_focal = 0;
    // persistent warps: cells are handed out in order by an atomic counter (work per cell varies with occupancy)
    while (true) {
    int c = 0;
    if (lane == 0) c = atomicAdd(a.work_counter, 1);
    c = a.g.own_c0 + __shfl_sync(FULL, c, 0);
    if (c >= a.g.own_c1) break;
    const int s0 = a.cell_start[c], e0 = a.cell_start[c + 1];
    if (s0 == e0) continue;
    int nb_s = 0, nb_e = 0;
    if (lane < 9) {
        const int cc = stencil_cell(a.g, c, lane);
        nb_s = a.cell_start[cc];
        nb_e = a.cell_start[cc + 1];
    }

    for (int g0 = s0; g0 < e0; g0 += V2_FMAX) {
        // ---- focal group: the active fish among 32 consecutive fish of the cell, compacted
        const int i = g0 + lane;
        const bool valid = i < e0;
        float4 me = make_float4(0.f, 0.f, 1.f, 0.f);
        int4 c1 = make_int4(0, 0, 0, 0);
        float dir_raw = 0.f;
        if (valid) { me = a.rec[i]; c1 = a.cold1[i]; dir_raw = a.cold0[i].x; }
        const bool active = valid && (a.active_lag_max < 0 || c1.x <= a.active_lag_max);
        const unsigned amask = __ballot_sync(FULL, active);
        const int nf = __popc(amask);
        if (valid && !active) a.aux[i] = 0u;
        if (nf == 0) continue;
        __syncwarp();
        if (active) {
            const int slot = __popc(amask & lt_mask);
            S.frec[slot] = me; S.fdir[slot] = dir_raw; S.fidx[slot] = i;
        }
        __syncwarp();
        if (STATS) st_focal += (unsigned)nf;

        // ---- window batches: the 9 stencil cells' fish are streamed through a WIN-entry window
        int fill = 0, k = 0;
        bool first_batch = true;
        int ks = __shfl_sync(FULL, nb_s, 0), ke = __shfl_sync(FULL, nb_e, 0);
        int b = ks;
        bool more = true;
        while (more) {
            while (fill < V2_WIN && k < 9) {
                if (b >= ke) {
                    ++k;
                    if (k < 9) { ks = __shfl_sync(FULL, nb_s, k); ke = __shfl_sync(FULL, nb_e, k); b = ks; }
                    continue;
                }
                const float sx = (float)((k % 3) - 1) * CELL_F, sy = (float)((k / 3) - 1) * CELL_F;
                const int m = min(ke - b, V2_WIN - fill);
                for (int t = lane; t < m; t += 32) {
                    const float4 r = __ldg(&a.rec[b + t]);
                    S.wx[fill + t] = r.x + sx; S.wy[fill + t] = r.y + sy; S.wc[fill + t] = r.z; S.ws[fill + t] = r.w;
                    S.wcode[fill + t] = (uint32_t)(b + t) | ((uint32_t)k << 28);
                }
                fill += m; b += m;
            }
            more = (k < 9);
            if (fill > 0) {
                // ======== one window batch of `fill` candidates ========
                __syncwarp();
                const int nch = (fill + 31) >> 5;
                float cxr[V2_CH], cyr[V2_CH];
#pragma unroll
                for (int ch = 0; ch < V2_CH; ++ch) {
                    const int j = ch * 32 + lane;
                    const bool in = j < fill;
                    cxr[ch] = in ? S.wx[j] : V2_FAR;
                    cyr[ch] = in ? S.wy[j] : V2_FAR;
                }
                for (int f = 0; f < nf; ++f) {
                    const float4 F = S.frec[f];  // broadcast
                    const float fx = F.x, fy = F.y, fc = F.z, fs = F.w;
                    // ---- filter (all candidates, cheap, conservative)
                    int qn = 0;
#pragma unroll
                    for (int ch = 0; ch < V2_CH; ++ch) {
                        if (ch < nch) {
                            const float dx = cxr[ch] - fx, dy = cyr[ch] - fy;
                            const float d2 = fmaf(dx, dx, dy * dy);
                            const float dot = fmaf(fc, dx, fs * dy);
                            // inside the disc, and not clearly behind the cone: dot > -|cos165| d, written sqrt-free with the
                            // signed square; the +1 keeps every pair closer than 1 unit (its rounding band is wide)
                            const bool p = (d2 < V2_R2HI) && (fmaf(V2_KCONE, d2, dot * fabsf(dot)) > -1.0f);
                            const unsigned bal = __ballot_sync(FULL, p);
                            if (p) S.queue[qn + __popc(bal & lt_mask)] = (unsigned char)(ch * 32 + lane);
                            qn += __popc(bal);
                        }
                    }
                    __syncwarp();
                    if (STATS) st_cand += (unsigned)fill;
                    // ---- process (compacted pairs, one per lane, same focal fish in every lane)
                    const float fdir = S.fdir[f];
                    const bool exotic = fabsf(fdir) >= 540.f;
                    PairAcc P;
                    pair_acc_clear(P);
                    for (int q0 = 0; q0 < qn; q0 += 32) {
                        const int e = q0 + lane;
                        const bool live = e < qn;
                        const int j = live ? (int)S.queue[e] : 0;
                        const bool unc = pair_math(
                            live, fx, fy, fc, fs, fdir, exotic, S.wx[j], S.wy[j], S.wc[j], S.ws[j], a.seed, a.replicate, a.step_move,
                            [&](float& rx, float& ry, int& kcode) {
                                const uint32_t code = S.wcode[j];
                                const float4 raw = __ldg(&a.rec[code & 0x0FFFFFFFu]);
                                rx = raw.x; ry = raw.y; kcode = (int)(code >> 28);
                            },
                            [&](uint32_t& my, uint32_t& nb) {
                                nb = (uint32_t)a.cold1[S.wcode[j] & 0x0FFFFFFFu].w;
                                my = (uint32_t)a.cold1[S.fidx[f]].w;
                            },
                            P);
                        if (STATS) st_rechk += __popc(__ballot_sync(FULL, unc));
                    }
                    // ---- reduce the lanes' partial sums; lane 0 owns the focal fish's accumulators
                    const int r0 = __reduce_add_sync(FULL, P.ax0), r1 = __reduce_add_sync(FULL, P.ay0);
                    const int r2 = __reduce_add_sync(FULL, P.ax1), r3 = __reduce_add_sync(FULL, P.ay1);
                    const int r4 = __reduce_add_sync(FULL, P.ax2), r5 = __reduce_add_sync(FULL, P.ay2);
                    const int cfar = __reduce_add_sync(FULL, P.c_far), cnear = __reduce_add_sync(FULL, P.c_near);
                    const int c200 = cfar & 0xFFFF, cfr = cfar >> 16, c100 = cnear & 0xFFFF, c15 = cnear >> 16;
                    const int tsum = __any_sync(FULL, P.ties != 0) ? __reduce_add_sync(FULL, P.ties) : 0;
                    if (STATS) st_inter += (unsigned)c200;
                    if (lane == 0) {
                        if (first_batch) {
                            S.acc[f][0] = r0; S.acc[f][1] = r1; S.acc[f][2] = r2; S.acc[f][3] = r3; S.acc[f][4] = r4; S.acc[f][5] = r5;
                            *reinterpret_cast<int4*>(&S.cnt[f][0]) = make_int4(c15, c100, c200, cfr);
                            S.fties[f] = tsum;
                        } else {
                            S.acc[f][0] += r0; S.acc[f][1] += r1; S.acc[f][2] += r2; S.acc[f][3] += r3; S.acc[f][4] += r4; S.acc[f][5] += r5;
                            int4 cc = *reinterpret_cast<int4*>(&S.cnt[f][0]);
                            cc.x += c15; cc.y += c100; cc.z += c200; cc.w += cfr;
                            *reinterpret_cast<int4*>(&S.cnt[f][0]) = cc;
                            S.fties[f] += tsum;
                        }
                    }
                    __syncwarp();
                }
            }
            if (fill > 0) first_batch = false;
            fill = 0;
        }

        // ---- epilogue, one focal fish per lane: social()'s selection (fish_mod.cpp:355-376), averageturn (:379-408),
        //      random walk (:372-374) and sample()'s feeding decision (:543-547)
        __syncwarp();
        if (lane < nf) {
            const int f = lane;
            const int idx = S.fidx[f];
            const int4 k1 = a.cold1[idx];
            const int n15 = S.cnt[f][0], n100 = S.cnt[f][1], n200 = S.cnt[f][2], nfr = S.cnt[f][3];
            const float4 F = S.frec[f];
            finish_focal(a, idx, k1, F.x, F.y, n15, n100, n200, nfr, &S.acc[f][0], S.fties[f]);
        }
        __syncwarp();
    }
    }  // persistent loop
    if (STATS && a.stats) {
        // st_* are warp-uniform except st_rechk (uniform too: from ballots)
        if (lane == 0) {
            atomicAdd(&a.stats[0], st_focal); atomicAdd(&a.stats[1], st_cand);
            atomicAdd(&a.stats[2], st_inter); atomicAdd(&a.stats[3], st_rechk);
        }
    }
}

}  // namespace fishk
