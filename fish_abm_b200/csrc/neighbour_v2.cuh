// K4, second design (default): candidate-parallel neighbour kernel.
//
// One warp owns one focal cell (persistent warps, cells handed out by an atomic counter).
//   staging   the 3x3 stencil is ONE flattened run of candidates (9 cell ranges, prefix-summed once per cell); each
//             lane copies candidates t = lane, lane+32, ... with a branch-free range lookup into the window (SoA x, y,
//             cos, sin in the focal cell's frame).  The window is staged once per cell and reused by every focal
//             group when the stencil fits it (<= 256 candidates).
//   focal     the ACTIVE fish of the cell are compacted (ballot) into groups of up to 16, wherever they sit in the cell.
//   filter    every lane tests its candidates (read from the window, chunks of 32) against one focal fish with a cheap,
//             conservative, sqrt-free predicate (inside the r=200 disc and not clearly behind the 330 deg cone);
//             survivors are compacted into a per-focal queue in shared memory with ballot/popc.
//   process   the queue is consumed 32 entries at a time, one pair per lane, all lanes working for the SAME focal fish:
//             exact classification (with the fp64 re-check of borderline pairs), zone counts and the avoidance /
//             alignment / attraction turn of the pair as fixed-point cos/sin: the shared pair_math().
//   reduce    eight warp REDUX adds fold the lanes' partial sums into the focal fish's 64-bit accumulators.
// Lanes stay full regardless of how many fish of the cell are active, and the expensive per-pair math only runs on
// (almost only) interacting pairs.  Sums are integers, hence independent of candidate order: results are bit-identical
// to the first design (neighbour_kernel) and across any re-sorting or slab split of the fish.
// 64 registers and 6 KB of shared memory per warp: 4 CTAs (32 warps) per SM.  Measured at the benchmark state: 16 warps/SM
// 344 us, 24 warps/SM (candidates cached in 16 registers per lane, 80 registers) 279 us, 32 warps/SM (this form) 266 us.
// (A half-warp-per-focal variant of process was measured 6 % slower at the benchmark density: queues of ~60 entries
// fill two 32-wide rounds as well as four 16-wide ones, and the half-warp reductions cost twice the REDUX issue slots.)
#pragma once
#include "kernels.cuh"

namespace fishk {

constexpr int V2_WARPS = 8;
constexpr int V2_THREADS = V2_WARPS * 32;
constexpr int V2_WIN = 256;           // candidates per window batch
constexpr int V2_FMAX = 16;           // focal fish per group
constexpr float V2_FAR = 1.0e18f;     // padding coordinate: fails the disc test
constexpr float V2_KCONE = 0.93301270189f + 2.5e-4f;   // cos^2(165 deg) + margin
constexpr float V2_R2HI = (CELL_F + 2.5e-4f) * (CELL_F + 2.5e-4f);

struct V2WarpSmem {
    float wx[V2_WIN], wy[V2_WIN], wc[V2_WIN], ws[V2_WIN];  // window candidates in the focal cell's frame
    long long acc[V2_FMAX][6];                             // fixed-point cos/sin sums: avoid x,y, align x,y, attract x,y
    int4 cnt[V2_FMAX];                                     // n15, n100, n200, front
    float4 frec[V2_FMAX];
    float fdir[V2_FMAX];
    int fidx[V2_FMAX];
    int fties[V2_FMAX];                                // tie draws across window batches (diagnostic)
    int run_pre[12];                                       // exclusive prefix of the 9 stencil cells' counts (+ total)
    int run_start[12];                                     // first sorted index of each stencil cell
    unsigned char queue[V2_WIN];
};
constexpr size_t V2_SMEM_BYTES = sizeof(V2WarpSmem) * V2_WARPS;

template <bool STATS>
__global__ void __launch_bounds__(V2_THREADS, 4) neighbour_kernel_v2(const NeighArgs a) {
    extern __shared__ __align__(16) unsigned char v2_smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const unsigned FULL = 0xFFFFFFFFu;
    const unsigned lt_mask = (1u << lane) - 1u;
    V2WarpSmem& S = reinterpret_cast<V2WarpSmem*>(v2_smem_raw)[warp];

    unsigned long long st_cand = 0, st_inter = 0, st_rechk = 0, st_focal = 0;
    if (blockIdx.x == 0 && threadIdx.x == 0) *a.work_counter_next = 0;
    const int n_items = (a.g.own_c1 - a.g.own_c0) * a.split;
    while (true) {  // persistent warps: work items (a cell, or one of `split` slices of a cell's fish) are handed out in order
        int t = 0;
        if (lane == 0) t = atomicAdd(a.work_counter, 1);
        t = __shfl_sync(FULL, t, 0);
        if (t >= n_items) break;
        const int c = a.g.own_c0 + t / a.split;
        int s0 = a.cell_start[c], e0 = a.cell_start[c + 1];
        if (a.split > 1) {  // this item's slice of the cell's fish (the whole stencil is still its candidate set)
            const int len = (e0 - s0 + a.split - 1) / a.split, sub = t % a.split;
            s0 = min(e0, s0 + sub * len);
            e0 = min(e0, s0 + len);
        }
        if (s0 == e0) continue;

        // ---- the stencil as one flattened run: per-cell ranges, exclusive prefix (lanes 0..8), total T
        int nb_s = 0, nb_n = 0;
        if (lane < 9) {
            const int cc = stencil_cell(a.g, c, lane);
            nb_s = a.cell_start[cc];
            nb_n = a.cell_start[cc + 1] - nb_s;
        }
        int pre = nb_n;
#pragma unroll
        for (int o = 1; o < 16; o <<= 1) { const int t = __shfl_up_sync(FULL, pre, o); if (lane >= o) pre += t; }
        const int T = __shfl_sync(FULL, pre, 8);
        pre -= nb_n;
        __syncwarp();
        if (lane < 9) { S.run_pre[lane] = pre; S.run_start[lane] = nb_s; }
        if (lane == 9) S.run_pre[9] = T;
        __syncwarp();
        const int nbatch = (T + V2_WIN - 1) / V2_WIN;
        int staged = -1;  // batch currently in the window

        auto stage = [&](int b) {
            const int base = b * V2_WIN;
            const int fill = min(T - base, V2_WIN);
            const int padded = (fill + 31) & ~31;
            const int p1 = S.run_pre[1], p2 = S.run_pre[2], p3 = S.run_pre[3], p4 = S.run_pre[4];
            const int p5 = S.run_pre[5], p6 = S.run_pre[6], p7 = S.run_pre[7], p8 = S.run_pre[8];
            for (int j = lane; j < padded; j += 32) {
                float x = V2_FAR, y = V2_FAR, cs = 1.f, sn = 0.f;
                if (j < fill) {
                    const int t = base + j;
                    const int k = (t >= p1) + (t >= p2) + (t >= p3) + (t >= p4) + (t >= p5) + (t >= p6) + (t >= p7) + (t >= p8);
                    const float4 r = __ldg(&a.rec[S.run_start[k] + t - S.run_pre[k]]);
                    const int ky = (k * 11) >> 5, kx = k - 3 * ky;  // k / 3, k % 3 for k < 9
                    x = r.x + (float)(kx - 1) * CELL_F; y = r.y + (float)(ky - 1) * CELL_F; cs = r.z; sn = r.w;
                }
                S.wx[j] = x; S.wy[j] = y; S.wc[j] = cs; S.ws[j] = sn;
            }
            __syncwarp();
            staged = b;
            return fill;
        };
        // sorted index and stencil code of window slot j of batch b (rare paths only: fp64 re-check, tie draws)
        auto locate = [&](int b, int j, int& src, int& k) {
            const int t = b * V2_WIN + j;
            k = 0;
#pragma unroll
            for (int m = 1; m < 9; ++m) k += (t >= S.run_pre[m]) ? 1 : 0;
            src = S.run_start[k] + t - S.run_pre[k];
        };

        int pos = s0;
        while (pos < e0) {
            // ---- focal group: up to V2_FMAX active fish of the cell, compacted
            int nf = 0;
            __syncwarp();
            while (pos < e0 && nf < V2_FMAX) {
                const int i = pos + lane;
                const bool valid = i < e0;
                float4 me = make_float4(0.f, 0.f, 1.f, 0.f);
                int lag = 0x7FFFFFFF;
                if (valid) lag = a.cold1[i].x;
                bool active = valid && (a.active_lag_max < 0 || lag <= a.active_lag_max);
                unsigned amask = __ballot_sync(FULL, active);
                int take = 32;  // fish of this 32-chunk consumed by this group
                if (nf + __popc(amask) > V2_FMAX) {
                    take = __fns(amask, 0, V2_FMAX - nf + 1);  // lane of the first active fish that no longer fits
                    amask &= (1u << take) - 1u;
                    active = active && lane < take;
                }
                if (valid && !active && lane < take) a.aux[i] = 0u;
                if (active) {
                    me = a.rec[i];
                    const int slot = nf + __popc(amask & lt_mask);
                    S.frec[slot] = me; S.fdir[slot] = a.cold0[i].x; S.fidx[slot] = i;
                }
                nf += __popc(amask);
                pos += take;
            }
            __syncwarp();
            if (nf == 0) continue;
            if (STATS) st_focal += (unsigned)nf;

            // lane f keeps focal fish f's reduced sums of the current batch in registers; shared-memory accumulators are
            // only touched when the stencil needs more than one window batch
            int m0 = 0, m1 = 0, m2 = 0, m3 = 0, m4 = 0, m5 = 0, mfar = 0, mnear = 0;
            for (int b = 0; b < nbatch; ++b) {
                const int fill = (staged == b) ? min(T - b * V2_WIN, V2_WIN) : stage(b);
                const int nch = (fill + 31) >> 5;
                for (int f = 0; f < nf; ++f) {
                    const float4 F = S.frec[f];  // broadcast
                    // ---- filter (all candidates, cheap, conservative)
                    int qn = 0;
#pragma unroll 2
                    for (int ch = 0; ch < nch; ++ch) {
                        {
                            const float dx = S.wx[ch * 32 + lane] - F.x, dy = S.wy[ch * 32 + lane] - F.y;
                            const float d2 = fmaf(dx, dx, dy * dy);
                            const float dot = fmaf(F.z, dx, F.w * dy);
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
                            live, F.x, F.y, F.z, F.w, fdir, exotic, S.wx[j], S.wy[j], S.wc[j], S.ws[j], a.seed, a.replicate, a.step_move,
                            [&](float& rx, float& ry, int& kcode) {
                                int src;
                                locate(b, j, src, kcode);
                                const float4 raw = __ldg(&a.rec[src]);
                                rx = raw.x; ry = raw.y;
                            },
                            [&](uint32_t& my, uint32_t& nb) {
                                int src, k;
                                locate(b, j, src, k);
                                nb = (uint32_t)a.cold1[src].w;
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
                    if (STATS) st_inter += (unsigned)(cfar & 0xFFFF);
                    if (lane == f) { m0 = r0; m1 = r1; m2 = r2; m3 = r3; m4 = r4; m5 = r5; mfar = cfar; mnear = cnear; }
                }
                if (nbatch > 1 && lane < nf) {
                    const int4 cc = unpack_counts(mfar, mnear);
                    if (b == 0) {
                        S.acc[lane][0] = m0; S.acc[lane][1] = m1; S.acc[lane][2] = m2; S.acc[lane][3] = m3; S.acc[lane][4] = m4; S.acc[lane][5] = m5;
                        S.cnt[lane] = cc;
                        S.fties[lane] = unpack_ties(mfar);
                    } else {
                        S.acc[lane][0] += m0; S.acc[lane][1] += m1; S.acc[lane][2] += m2; S.acc[lane][3] += m3; S.acc[lane][4] += m4; S.acc[lane][5] += m5;
                        int4 o = S.cnt[lane];
                        o.x += cc.x; o.y += cc.y; o.z += cc.z; o.w += cc.w;
                        S.cnt[lane] = o;
                        S.fties[lane] += unpack_ties(mfar);
                    }
                }
            }
            // ---- epilogue, one focal fish per lane
            __syncwarp();
            if (lane < nf) {
                const int idx = S.fidx[lane];
                const int4 k1 = a.cold1[idx];
                const float4 F = S.frec[lane];
                if (nbatch > 1) {
                    const int4 cc = S.cnt[lane];
                    finish_focal(a, idx, k1, F.x, F.y, cc.x, cc.y, cc.z, cc.w, &S.acc[lane][0], S.fties[lane]);
                } else {
                    const long long sums[6] = {m0, m1, m2, m3, m4, m5};
                    const int4 cc = unpack_counts(mfar, mnear);
                    finish_focal(a, idx, k1, F.x, F.y, cc.x, cc.y, cc.z, cc.w, sums, unpack_ties(mfar));
                }
            }
            __syncwarp();
        }
    }
    if (STATS && a.stats && lane == 0) {
        atomicAdd(&a.stats[0], st_focal); atomicAdd(&a.stats[1], st_cand);
        atomicAdd(&a.stats[2], st_inter); atomicAdd(&a.stats[3], st_rechk);
    }
}

}  // namespace fishk
