// K4, third design (default): candidate-parallel neighbour kernel with a packed-half filter and a branch-free pair round.
//
// Same decomposition as the second design (neighbour_v2.cuh: one warp per focal cell, the 3x3 stencil staged once per cell
// as a window of candidates in the focal cell's frame, active fish compacted into focal groups, a cheap conservative
// filter that compacts survivors into a queue, the exact pair math on the survivors 32 at a time, warp REDUX per focal
// fish, one epilogue lane per focal fish) with the instruction count cut where ncu put it (the kernel issues
// on 85 % of its cycles; the filter was ~40 % of the issued instructions, the pair round ~140 instructions):
//   filter   TWO candidates per lane per instruction: the window also holds the candidates' coordinates as packed fp16
//            pairs (scaled by 1/4 around the cell centre: |value| < 75, resolution 0.25 units), and the disc test
//            dx^2 + dy^2 < (201.1)^2 runs as HADD2 / HMUL2 / HFMA2 / HSETP2: about 21 issued instructions per 64
//            candidates instead of per 32.  It is conservative by a proven margin (below); the exact fp32 / fp64
//            classification is the pair round's.  (The 330 deg cone test left the filter: it removed 1 candidate in 12 at
//            the cost of five more instructions per step.)
//   round    pair_round() (kernels.cuh): straight-line code, no per-lane branches; the fp64 re-check, tie draws and exotic
//            headings are one warp-uniform fallback to the generic pair_math.  Fixed-point sums are raw bit patterns
//            (bias removed once per focal fish).
// Results are bit-identical to the first and second designs (integer sums, same per-pair arithmetic):
// tests/test_gpu_edge_cases.py::test_neighbour_designs_agree.
//
// Filter margin.  Scaled units (1 = 4 world units).  Stored coordinate: RN_half((x - 100) / 4), |.| < 75 -> error <= 2^-5.
// Difference of two stored values: |.| <= 150 -> rounding <= 2^-4, so each of dx, dy is off by <= 0.125 and the distance by
// <= 0.177.  A pair with true distance < 200.0002 (50.00005 scaled) therefore has a computed distance < 50.18, square
// < 2518.1; the two fp16 roundings of dy*dy and of the fma add <= 2^-10 relative -> < 2520.6.  The threshold is 2528
// (representable: fp16 spacing is 2 there), i.e. every pair closer than 200.0002 passes; pairs beyond 201.2 are rejected.
#pragma once
#include <cuda_fp16.h>

#include "kernels.cuh"

namespace fishk {

constexpr int V3_WARPS = 8;
constexpr int V3_THREADS = V3_WARPS * 32;
constexpr int V3_SLOTS = 256;         // window slots (queue entries are bytes); the last one is a permanent far-away candidate
constexpr int V3_WIN = V3_SLOTS - 1;  // candidates per window batch: <= 255, so that the 8-bit packed counts cannot overflow
constexpr int V3_FMAX = 12;           // focal fish per group
constexpr float V3_FAR_X = 1.0e18f, V3_FAR_Y = 3.0e17f;   // padding candidate: beyond every radius, +inf in fp16; not on a lattice bearing
constexpr float V3_HSCALE = 0.25f, V3_HCENTRE = 100.f;
constexpr float V3_R2H = 2528.f;

struct V3WarpSmem {
    float4 win[V3_SLOTS];                                  // window candidates in the focal cell's frame: x, y, cos, sin
    uint32_t hx[V3_SLOTS / 2], hy[V3_SLOTS / 2];           // the same coordinates as fp16 pairs: element 32*s + l = candidates (64*s + l, 64*s + 32 + l)
    long long acc[V3_FMAX][6];                             // multi-batch only: exact cos/sin sums (avoid x,y, align x,y, attract x,y)
    int4 cnt[V3_FMAX];                                     // multi-batch only: n15, n100, n200, front
    float4 frec[V3_FMAX];
    float fdir[V3_FMAX];
    int fidx[V3_FMAX];
    uint32_t hfx[V3_FMAX], hfy[V3_FMAX];                   // focal coordinates, fp16 pair (both halves equal)
    int run_pre[12];                                       // exclusive prefix of the 9 stencil cells' counts (+ total)
    int run_start[12];                                     // first sorted index of each stencil cell
    unsigned char queue[V3_SLOTS + 32];                    // survivors of the filter, padded to a whole round with the far slot
};
constexpr size_t V3_SMEM_BYTES = sizeof(V3WarpSmem) * V3_WARPS;

__device__ __forceinline__ uint32_t v3_pack_half2(float lo, float hi) {
    const __half2 h = __floats2half2_rn(lo, hi);
    return *reinterpret_cast<const uint32_t*>(&h);
}
// per-half "a < b" as two predicates
__device__ __forceinline__ void v3_setp_lt2(uint32_t a, uint32_t b, bool& plo, bool& phi) {
    int lo, hi;
    asm("{\n\t.reg .pred p, q;\n\tsetp.lt.f16x2 p|q, %2, %3;\n\tselp.s32 %0, 1, 0, p;\n\tselp.s32 %1, 1, 0, q;\n\t}" : "=r"(lo), "=r"(hi) : "r"(a), "r"(b));
    plo = lo != 0; phi = hi != 0;
}

template <bool STATS>
__global__ void __launch_bounds__(V3_THREADS, 4) neighbour_kernel_v3(const NeighArgs a) {
    extern __shared__ __align__(16) unsigned char v3_smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const unsigned FULL = 0xFFFFFFFFu;
    const unsigned lt_mask = (1u << lane) - 1u;
    V3WarpSmem& S = reinterpret_cast<V3WarpSmem*>(v3_smem_raw)[warp];

    unsigned long long st_cand = 0, st_inter = 0, st_rechk = 0, st_focal = 0;
    if (blockIdx.x == 0 && threadIdx.x == 0) *a.work_counter_next = 0;
    const int n_items = (a.g.own_c1 - a.g.own_c0) * a.split;
    const __half2 r2h = __float2half2_rn(V3_R2H);
    const uint32_t r2h_bits = *reinterpret_cast<const uint32_t*>(&r2h);
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
        for (int o = 1; o < 16; o <<= 1) { const int u = __shfl_up_sync(FULL, pre, o); if (lane >= o) pre += u; }
        const int T = __shfl_sync(FULL, pre, 8);
        pre -= nb_n;
        __syncwarp();
        if (lane < 9) { S.run_pre[lane] = pre; S.run_start[lane] = nb_s; }
        if (lane == 9) S.run_pre[9] = T;
        __syncwarp();
        const int nbatch = (T + V3_WIN - 1) / V3_WIN;
        int staged = -1;  // batch currently in the window

        auto stage = [&](int b) {
            const int base = b * V3_WIN;
            const int fill = min(T - base, V3_WIN);
            const int padded = (fill + 64) & ~63;   // at least one padding slot; slot V3_SLOTS - 1 is always padding
            const int p1 = S.run_pre[1], p2 = S.run_pre[2], p3 = S.run_pre[3], p4 = S.run_pre[4];
            const int p5 = S.run_pre[5], p6 = S.run_pre[6], p7 = S.run_pre[7], p8 = S.run_pre[8];
            for (int j0 = 0; j0 < padded; j0 += 64) {   // two candidates per lane and iteration: j0 + lane, j0 + 32 + lane
                float xs[2], ys[2];
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const int j = j0 + 32 * h + lane;
                    float4 w = make_float4(V3_FAR_X, V3_FAR_Y, 1.f, 0.f);
                    if (j < fill) {
                        const int u = base + j;
                        const int k = (u >= p1) + (u >= p2) + (u >= p3) + (u >= p4) + (u >= p5) + (u >= p6) + (u >= p7) + (u >= p8);
                        const float4 r = __ldg(&a.rec[S.run_start[k] + u - S.run_pre[k]]);
                        const int ky = (k * 11) >> 5, kx = k - 3 * ky;  // k / 3, k % 3 for k < 9
                        w = make_float4(r.x + (float)(kx - 1) * CELL_F, r.y + (float)(ky - 1) * CELL_F, r.z, r.w);
                    }
                    S.win[j] = w;
                    xs[h] = (w.x - V3_HCENTRE) * V3_HSCALE; ys[h] = (w.y - V3_HCENTRE) * V3_HSCALE;
                }
                S.hx[(j0 >> 1) + lane] = v3_pack_half2(xs[0], xs[1]);
                S.hy[(j0 >> 1) + lane] = v3_pack_half2(ys[0], ys[1]);
            }
            if (lane == 0) S.win[V3_SLOTS - 1] = make_float4(V3_FAR_X, V3_FAR_Y, 1.f, 0.f);
            __syncwarp();
            staged = b;
            return fill;
        };
        // sorted index and stencil code of window slot j of batch b (rare paths only: fp64 re-check, tie draws)
        auto locate = [&](int b, int j, int& src, int& k) {
            const int u = b * V3_WIN + j;
            k = 0;
#pragma unroll
            for (int m = 1; m < 9; ++m) k += (u >= S.run_pre[m]) ? 1 : 0;
            src = S.run_start[k] + u - S.run_pre[k];
        };

        int pos = s0;
        while (pos < e0) {
            // ---- focal group: up to V3_FMAX active fish of the cell, compacted
            int nf = 0;
            __syncwarp();
            while (pos < e0 && nf < V3_FMAX) {
                const int i = pos + lane;
                const bool valid = i < e0;
                int lag = 0x7FFFFFFF;
                if (valid) lag = a.cold1[i].x;
                bool active = valid && (a.active_lag_max < 0 || lag <= a.active_lag_max);
                unsigned amask = __ballot_sync(FULL, active);
                int take = 32;  // fish of this 32-chunk consumed by this group
                if (nf + __popc(amask) > V3_FMAX) {
                    take = __fns(amask, 0, V3_FMAX - nf + 1);  // lane of the first active fish that no longer fits
                    amask &= (1u << take) - 1u;
                    active = active && lane < take;
                }
                if (valid && !active && lane < take) a.aux[i] = 0u;
                if (active) {
                    const float4 me = a.rec[i];
                    const int slot = nf + __popc(amask & lt_mask);
                    S.frec[slot] = me; S.fdir[slot] = a.cold0[i].x; S.fidx[slot] = i;
                    S.hfx[slot] = v3_pack_half2((me.x - V3_HCENTRE) * V3_HSCALE, (me.x - V3_HCENTRE) * V3_HSCALE);
                    S.hfy[slot] = v3_pack_half2((me.y - V3_HCENTRE) * V3_HSCALE, (me.y - V3_HCENTRE) * V3_HSCALE);
                }
                nf += __popc(amask);
                pos += take;
            }
            __syncwarp();
            if (nf == 0) continue;
            if (STATS) st_focal += (unsigned)nf;

            // lane f keeps focal fish f's reduced sums of the current batch in registers; shared-memory accumulators are
            // only touched when the stencil needs more than one window batch
            unsigned m0 = 0, m1 = 0, m2 = 0, m3 = 0, m4 = 0, m5 = 0, mcw = 0;
            for (int b = 0; b < nbatch; ++b) {
                const int fill = (staged == b) ? min(T - b * V3_WIN, V3_WIN) : stage(b);
                const int nst = (fill + 63) >> 6;
                for (int f = 0; f < nf; ++f) {
                    // ---- filter: two candidates per lane and step, packed fp16, conservative disc test
                    const uint32_t hfx = S.hfx[f], hfy = S.hfy[f];  // broadcast
                    const __half2 fx2 = *reinterpret_cast<const __half2*>(&hfx), fy2 = *reinterpret_cast<const __half2*>(&hfy);
                    int qn = 0;
                    int slot = lane;
#pragma unroll 2
                    for (int st = 0; st < nst; ++st) {
                        const uint32_t cxb = S.hx[st * 32 + lane], cyb = S.hy[st * 32 + lane];
                        const __half2 dx = __hsub2(*reinterpret_cast<const __half2*>(&cxb), fx2);
                        const __half2 dy = __hsub2(*reinterpret_cast<const __half2*>(&cyb), fy2);
                        const __half2 d2 = __hfma2(dx, dx, __hmul2(dy, dy));
                        bool p0, p1;
                        v3_setp_lt2(*reinterpret_cast<const uint32_t*>(&d2), r2h_bits, p0, p1);
                        const unsigned b0 = __ballot_sync(FULL, p0), b1 = __ballot_sync(FULL, p1);
                        const int n0 = __popc(b0);
                        if (p0) S.queue[qn + __popc(b0 & lt_mask)] = (unsigned char)slot;
                        if (p1) S.queue[qn + n0 + __popc(b1 & lt_mask)] = (unsigned char)(slot | 32);
                        qn += n0 + __popc(b1);
                        slot += 64;
                    }
                    S.queue[qn + lane] = (unsigned char)(V3_SLOTS - 1);   // pad the last round with the far-away slot
                    __syncwarp();
                    if (STATS) st_cand += (unsigned)fill;
                    // ---- pair rounds (compacted pairs, one per lane, same focal fish in every lane)
                    const float4 F = S.frec[f];  // broadcast
                    const float fdir = S.fdir[f];
                    PairAccR P;
                    pair_acc_clear(P);
                    auto raw_of = [&](int j, float& rx, float& ry, int& kcode) {
                        int src;
                        locate(b, j, src, kcode);
                        const float4 raw = __ldg(&a.rec[src]);
                        rx = raw.x; ry = raw.y;
                    };
                    auto ids_of = [&](int j, uint32_t& my, uint32_t& nb) {
                        int src, k;
                        locate(b, j, src, k);
                        nb = (uint32_t)a.cold1[src].w;
                        my = (uint32_t)a.cold1[S.fidx[f]].w;
                    };
                    if (fabsf(fdir) < 540.f) {
                        const unsigned char* qp = S.queue + lane;
                        for (int q0 = 0; q0 < qn; q0 += 32, qp += 32) {
                            const int j = (int)*qp;
                            const float4 Q = S.win[j];
                            const bool unc = pair_round<true>(
                                q0 + lane < qn, F.x, F.y, F.z, F.w, fdir, Q.x, Q.y, Q.z, Q.w, a.seed, a.replicate, a.step_move,
                                [&](float& rx, float& ry, int& kcode) { raw_of(j, rx, ry, kcode); },
                                [&](uint32_t& my, uint32_t& nb) { ids_of(j, my, nb); }, P);
                            if (STATS) st_rechk += __popc(__ballot_sync(FULL, unc));
                        }
                    } else {   // exotic raw heading: the generic code for every round
                        for (int q0 = 0; q0 < qn; q0 += 32) {
                            const int j = (int)S.queue[q0 + lane];
                            const float4 Q = S.win[j];
                            const bool unc = pair_round_generic<true>(
                                q0 + lane < qn, F.x, F.y, F.z, F.w, fdir, true, Q.x, Q.y, Q.z, Q.w, a.seed, a.replicate, a.step_move,
                                [&](float& rx, float& ry, int& kcode) { raw_of(j, rx, ry, kcode); },
                                [&](uint32_t& my, uint32_t& nb) { ids_of(j, my, nb); }, P);
                            if (STATS) st_rechk += __popc(__ballot_sync(FULL, unc));
                        }
                    }
                    // ---- reduce the lanes' partial sums; lane f owns focal fish f's values
                    const unsigned r0 = __reduce_add_sync(FULL, P.ax0), r1 = __reduce_add_sync(FULL, P.ay0);
                    const unsigned r2 = __reduce_add_sync(FULL, P.ax1), r3 = __reduce_add_sync(FULL, P.ay1);
                    const unsigned r4 = __reduce_add_sync(FULL, P.ax2), r5 = __reduce_add_sync(FULL, P.ay2);
                    const unsigned cw = __reduce_add_sync(FULL, P.c_far);
                    if (STATS) st_inter += (unsigned)((cw >> 16) & 0xFFu);
                    if (lane == f) { m0 = r0; m1 = r1; m2 = r2; m3 = r3; m4 = r4; m5 = r5; mcw = cw; }
                }
                if (nbatch > 1 && lane < nf) {
                    const unsigned r[6] = {m0, m1, m2, m3, m4, m5};
                    long long s[6];
                    const int4 cc = unpack_counts_r<true>(mcw, 0u);
                    pair_sums_unbias(r, cc, s);
                    if (b == 0) {
#pragma unroll
                        for (int k = 0; k < 6; ++k) S.acc[lane][k] = s[k];
                        S.cnt[lane] = cc;
                    } else {
#pragma unroll
                        for (int k = 0; k < 6; ++k) S.acc[lane][k] += s[k];
                        int4 o = S.cnt[lane];
                        o.x += cc.x; o.y += cc.y; o.z += cc.z; o.w += cc.w;
                        S.cnt[lane] = o;
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
                    finish_focal(a, idx, k1, F.x, F.y, cc.x, cc.y, cc.z, cc.w, &S.acc[lane][0], 0);
                } else {
                    const unsigned r[6] = {m0, m1, m2, m3, m4, m5};
                    long long sums[6];
                    const int4 cc = unpack_counts_r<true>(mcw, 0u);
                    pair_sums_unbias(r, cc, sums);
                    finish_focal(a, idx, k1, F.x, F.y, cc.x, cc.y, cc.z, cc.w, sums, 0);
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
