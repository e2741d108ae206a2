// K4, fourth design (default): the candidate-parallel kernel of neighbour_v3.cuh with the per-cell fixed costs halved.
//
// ncu on v3 (profiles/r02_ncu_v3_sass_regions.txt) put 24 % of the issued instructions into work that is done once per
// CELL, not per pair: handing out the work item and building the stencil (6 %), staging the 3x3 window (10 %), and the
// per-fish epilogue that runs with 6 of 32 lanes busy (8 %).  Cells are column-major, so the two cells (2p, 2p+1) of a
// column are neighbours in memory and their stencils share 6 of 9 cells.  Here
//   * a work item is a vertical PAIR of cells: ONE 3x4 window (<= 255 candidates) serves the fish of both cells, in one
//     y frame (yframe_shift, kernels.cuh); a fish filters only the three window rows around its own cell.  A pair whose
//     window would not fit is done as two single cells in the same frame (identical arithmetic), as is every cell of a
//     school so small or dense that cells are split over several warps (`split` > 1);
//   * focal groups hold up to 24 fish (16 in v3): both cells' active fish usually share one group and one epilogue;
//   * the round has no fallback branch: pair_fast() accumulates unconditionally and flags borderline pairs and ties; a
//     flagged fish (1 in ~1000) is redone by the generic per-lane code.  Vote, branch and reconvergence leave the round;
//   * per focal fish ONE 16-byte record carries the filter's operands, its window range and the fish's index;
//   * the stencil is built from the item's (column, row) without integer divisions per lane.
// Filter (packed fp16, two candidates per lane and instruction), queue, REDUX reduction and epilogue are v3's.  Pair sums
// are integers and the per-pair arithmetic is shared, so results are bit-identical to v3 / v1 / the ensemble kernel
// (tests/test_gpu_edge_cases.py::test_neighbour_designs_agree).
//
// Filter margin (as in neighbour_v3.cuh, for |y| < 400).  Scaled units (1 = 4 world units).  Stored coordinate:
// RN_half((x - 100) / 4) in [-75, 75), RN_half(y / 4) in [-100, 100): error <= 2^-5.  Difference of two stored values:
// |.| < 256 -> rounding <= 2^-4, so each of dx, dy is off by <= 0.125 and the distance by <= 0.177.  A pair with true distance
// < 200.0002 (50.00005 scaled) has a computed distance < 50.18, square < 2518.1; the two fp16 roundings of dy*dy and of the
// fma add <= 2^-10 relative -> < 2520.6.  The threshold is 2528: every pair closer than 200.0002 passes; pairs beyond
// 201.2 are rejected.  Differences up to 200 scaled square to < 65504 (no overflow before the sum; the sum may become +inf,
// which fails the test as it must).
#pragma once
#include <cuda_fp16.h>

#include "kernels.cuh"

namespace fishk {

constexpr int V4_WARPS = 8;
constexpr int V4_THREADS = V4_WARPS * 32;
constexpr int V4_SLOTS = 256;         // window slots (queue entries are bytes); the last one is a permanent far-away candidate
constexpr int V4_WIN = V4_SLOTS - 1;  // candidates per window batch: <= 255, so that the 8-bit packed counts cannot overflow
constexpr int V4_FMAX = 24;           // focal fish per group (one window batch)
constexpr int V4_FMAX_MB = 12;        // focal fish per group when the stencil needs several window batches (64-bit sums)
constexpr float V4_FAR_X = 1.0e18f, V4_FAR_Y = 3.0e17f;   // padding candidate: beyond every radius, +inf in fp16; not on a lattice bearing
constexpr float V4_HSCALE = 0.25f, V4_HCX = 100.f;
constexpr float V4_R2H = 2528.f;
constexpr uint32_t V4_META_EXOTIC = 0x80000000u;

struct V4WarpSmem {
    float4 win[V4_SLOTS];                                  // window candidates in the item's frame: x, y, cos, sin
    uint32_t hx[V4_SLOTS / 2], hy[V4_SLOTS / 2];           // the same coordinates as fp16 pairs: element 32*s + l = candidates (64*s + l, 64*s + 32 + l)
    union {
        uint32_t res[V4_FMAX][8];                          // one window batch: the reduced raw sums (6) and packed counts of every focal fish
        struct {
            long long acc[V4_FMAX_MB][6];                  // several batches: exact cos/sin sums (avoid x,y, align x,y, attract x,y)
            int4 cnt[V4_FMAX_MB];                          //                  n15, n100, n200, front
        } mb;
    };
    float4 frec[V4_FMAX];                                  // focal fish in the item's frame: x, y, cos, sin
    uint4 fmeta[V4_FMAX];                                  // fp16 x pair, fp16 y pair, flags (exotic heading), sorted index
    int run_pre[16];                                       // exclusive prefix of the window cells' counts (3 per row, + total)
    int run_off[12];                                       // sorted index of a window cell's fish = run_off[cell] + flat slot
    float2 shift[12];                                      // what takes a record of that cell into the item's frame (periodic minimum image)
    unsigned char queue[V4_SLOTS + 32];                    // survivors of the filter, padded to a whole round with the far slot
};
constexpr size_t V4_SMEM_BYTES = sizeof(V4WarpSmem) * V4_WARPS;
static_assert(4 * (V4_SMEM_BYTES + 1024) <= 228 * 1024, "four CTAs per SM: 228 KB of shared memory, 1 KB of it reserved per CTA");

__device__ __forceinline__ uint32_t v4_pack_half2(float lo, float hi) {
    const __half2 h = __floats2half2_rn(lo, hi);
    return *reinterpret_cast<const uint32_t*>(&h);
}
// per-half "a < b" as two predicates
__device__ __forceinline__ void v4_setp_lt2(uint32_t a, uint32_t b, bool& plo, bool& phi) {
    int lo, hi;
    asm("{\n\t.reg .pred p, q;\n\tsetp.lt.f16x2 p|q, %2, %3;\n\tselp.s32 %0, 1, 0, p;\n\tselp.s32 %1, 1, 0, q;\n\t}" : "=r"(lo), "=r"(hi) : "r"(a), "r"(b));
    plo = lo != 0; phi = hi != 0;
}

template <bool STATS>
__global__ void __launch_bounds__(V4_THREADS, 4) neighbour_kernel_v4(const NeighArgs a) {
    extern __shared__ __align__(16) unsigned char v4_smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const unsigned FULL = 0xFFFFFFFFu;
    const unsigned lt_mask = (1u << lane) - 1u;
    V4WarpSmem& S = reinterpret_cast<V4WarpSmem*>(v4_smem_raw)[warp];

    unsigned long long st_cand = 0, st_inter = 0, st_rechk = 0, st_focal = 0, st_t0 = 0;
    if (STATS) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(st_t0));
    if (blockIdx.x == 0 && threadIdx.x == 0) *a.work_counter_next = 0;
    const int ny = a.g.ny;
    const int col0 = a.g.own_c0 / ny, ncols = (a.g.own_c1 - a.g.own_c0) / ny;
    const int npc = (ny + 1) >> 1;                                // cell pairs per column (the last one is a single cell when ny is odd)
    const int pair_items = (ncols - a.single_cols) * npc;        // split == 1: cell pairs first, the last columns cell by cell
    const int n_items = a.split > 1 ? ncols * ny * a.split : pair_items + a.single_cols * ny;
    const __half2 r2h = __float2half2_rn(V4_R2H);
    const uint32_t r2h_bits = *reinterpret_cast<const uint32_t*>(&r2h);
    const int kx_l = lane % 3, ky_l = lane / 3;                  // this lane's window cell (lanes 0..11)

    while (true) {  // persistent warps: work items are handed out in order
        int t = 0;
        if (lane == 0) t = atomicAdd(a.work_counter, 1);
        t = __shfl_sync(FULL, t, 0);
        if (t >= n_items) break;
        // ---- the item: cells (cx, cy0 .. cy0 + left - 1) and, when cells are split over several warps, a slice of the cell's fish
        int cx, cy0, left, sub = 0;
        if (a.split > 1) {
            int cell;
            cell_xy(t, a.split, cell, sub);
            cell_xy(cell, ny, cx, cy0); left = 1;
        } else if (t < pair_items) {
            int p;
            cell_xy(t, npc, cx, p);
            cy0 = 2 * p; left = min(2, ny - cy0);
        } else {
            cell_xy(t - pair_items, ny, cx, cy0); left = 1;
            cx += ncols - a.single_cols;
        }
        cx += col0;
        int nfc = left;                                           // cells whose fish this window serves: 2, or 1
        while (left > 0) {
            const int c = cx * ny + cy0;
            if (a.cell_start[c] == a.cell_start[c + nfc]) { left -= nfc; cy0 += nfc; nfc = left; continue; }   // no fish here
            // ---- the window's cells, row-major (3 per row, nfc + 2 rows): per-cell ranges, exclusive prefix, total T
            // rows cy0 - 1 .. cy0 + nfc; the item's frame has the ODD row of the pair (2p, 2p + 1) at offset 0
            const int ncw = 3 * (nfc + 2);
            int nb_s = 0, nb_n = 0;
            if (lane < ncw) {
                int xx = cx + kx_l - 1, yy = cy0 + ky_l - 1;
                if (a.g.wrap_x) xx = xx < 0 ? xx + a.g.lnx : (xx >= a.g.lnx ? xx - a.g.lnx : xx);
                yy = yy < 0 ? yy + ny : yy;
                while (yy >= ny) yy -= ny;                        // (cy0 + 2 can pass the top twice only when ny == 1)
                const int cc = xx * ny + yy;
                nb_s = a.cell_start[cc];
                nb_n = a.cell_start[cc + 1] - nb_s;
            }
            int pre = nb_n;
#pragma unroll
            for (int o = 1; o < 16; o <<= 1) { const int u = __shfl_up_sync(FULL, pre, o); if (lane >= o) pre += u; }
            const int T = __shfl_sync(FULL, pre, ncw - 1);
            if (nfc == 2 && T > V4_WIN) { nfc = 1; continue; }    // the pair's window would not fit: two single cells
            pre -= nb_n;
            // filter steps (64 candidates each) of the three window rows around each cell of the item (single batch): the first
            // cell sees rows 0..2 = slots [0, pre[9]) (= T for a single cell: lanes >= ncw hold no fish), the second rows 1..3 =
            // slots [pre[3], T).  (Shuffles, not shared-memory reads: the compiler keeps shuffled values in uniform registers.)
            const int pre9 = __shfl_sync(FULL, pre, 9), pre3 = __shfl_sync(FULL, pre, 3);
            const int st_end0 = (pre9 + 63) >> 6;
            const int st_beg1 = pre3 >> 6;
            const int st_end1 = (T + 63) >> 6;
            __syncwarp();
            // y frame: window row ky holds cell row cy0 - 1 + ky; the frame's zero is the odd row of the pair
            const int row0 = (cy0 & 1) ? 1 : 2;                   // window row of the cell row at offset 0
            const float hcy = (cy0 & 1) ? 100.f : 0.f;            // centre of the fp16 y coordinates (|y - hcy| <= 400)
            if (lane < ncw) {
                S.run_pre[lane] = pre; S.run_off[lane] = nb_s - pre;
                S.shift[lane] = make_float2((float)(kx_l - 1) * CELL_F, (float)(ky_l - row0) * CELL_F);
            }
            if (lane == ncw) S.run_pre[ncw] = T;
            __syncwarp();
            int s0 = a.cell_start[c], e0 = a.cell_start[c + nfc];
            const int mid = nfc == 2 ? a.cell_start[c + 1] : e0;  // fish from here on belong to the second cell of the pair
            if (a.split > 1) {  // this item's slice of the cell's fish (the whole stencil is still its candidate set)
                const int len = (e0 - s0 + a.split - 1) / a.split;
                s0 = min(e0, s0 + sub * len);
                e0 = min(e0, s0 + len);
            }
            const int nbatch = (T + V4_WIN - 1) / V4_WIN;
            const int fmax = nbatch > 1 ? V4_FMAX_MB : V4_FMAX;
            int staged = -1;  // batch currently in the window

            auto stage = [&](int b) {
                const int base = b * V4_WIN;
                const int fill = min(T - base, V4_WIN);
                const int padded = (fill + 64) & ~63;   // at least one padding slot; slot V4_SLOTS - 1 is always padding
                const float4 far = make_float4(V4_FAR_X, V4_FAR_Y, 1.f, 0.f);
                // flat copy, 64 slots per pass (two per lane, the pair the filter reads as one fp16x2 word): the window's cells are
                // contiguous runs of sorted records side by side in the slot space; a lane finds its slots' cells by counting
                // the cell boundaries inside the pass's span (one or two, read as broadcasts), loads the records, shifts them
                // into the item's frame and writes them as fp32 (rounds) and as packed fp16 pairs (filter)
                int kc = 0;                              // cell of the first slot of the pass (the same in every lane)
                for (int j0 = 0; j0 < padded; j0 += 64) {
                    const int u = base + j0 + lane, v = u + 32;
                    int ku = kc, kv = kc;
                    for (int m = kc + 1; m < ncw; ++m) {
                        const int pm = S.run_pre[m];
                        if (pm > base + j0 + 63) break;
                        ku += (u >= pm) ? 1 : 0; kv += (v >= pm) ? 1 : 0;
                    }
                    float4 q = far, r = far;
                    if (j0 + lane < fill) {
                        q = __ldg(&a.rec[S.run_off[ku] + u]);
                        const float2 sh = S.shift[ku];
                        q.x += sh.x; q.y += sh.y;
                    }
                    if (j0 + 32 + lane < fill) {
                        r = __ldg(&a.rec[S.run_off[kv] + v]);
                        const float2 sh = S.shift[kv];
                        r.x += sh.x; r.y += sh.y;
                    }
                    S.win[j0 + lane] = q; S.win[j0 + 32 + lane] = r;
                    S.hx[(j0 >> 1) + lane] = v4_pack_half2(fmaf(q.x, V4_HSCALE, -V4_HCX * V4_HSCALE), fmaf(r.x, V4_HSCALE, -V4_HCX * V4_HSCALE));
                    S.hy[(j0 >> 1) + lane] = v4_pack_half2((q.y - hcy) * V4_HSCALE, (r.y - hcy) * V4_HSCALE);
                    kc = __shfl_sync(FULL, kv, 31);
                }
                if (lane == 0) S.win[V4_SLOTS - 1] = far;
                __syncwarp();
                staged = b;
                return fill;
            };
            // sorted index and window cell of slot j of batch b (rare paths only: fp64 re-check, tie draws)
            auto locate = [&](int b, int j, int& src, int& k) {
                const int u = b * V4_WIN + j;
                k = 0;
#pragma unroll
                for (int m = 1; m < 12; ++m) k += (m < ncw && u >= S.run_pre[m]) ? 1 : 0;
                src = S.run_off[k] + u;
            };

            int pos = s0;
            while (pos < e0) {
                // ---- focal group: up to fmax active fish of the item, compacted
                int nf = 0, nf0 = 0;      // fish in the group; those of the item's first cell (they come first: fish are cell-sorted)
                unsigned any_exotic = 0u;
                __syncwarp();
                while (pos < e0 && nf < fmax) {
                    const int i = pos + lane;
                    const bool valid = i < e0;
                    int lag = 0x7FFFFFFF;
                    if (valid) lag = a.cold1[i].x;
                    bool active = valid && (a.active_lag_max < 0 || lag <= a.active_lag_max);
                    unsigned amask = __ballot_sync(FULL, active);
                    int take = 32;  // fish of this 32-chunk consumed by this group
                    if (nf + __popc(amask) > fmax) {
                        take = __fns(amask, 0, fmax - nf + 1);  // lane of the first active fish that no longer fits
                        amask &= (1u << take) - 1u;
                        active = active && lane < take;
                    }
                    if (valid && !active && lane < take) a.aux[i] = 0u;
                    bool exo = false;
                    if (active) {
                        float4 me = a.rec[i];
                        const float fdir = a.cold0[i].x;
                        const int fc = i >= mid ? 1 : 0;
                        me.y += (float)(1 + fc - row0) * CELL_F;   // the cell's window row is 1 + fc
                        const int slot = nf + __popc(amask & lt_mask);
                        S.frec[slot] = me;
                        const float hxs = fmaf(me.x, V4_HSCALE, -V4_HCX * V4_HSCALE), hys = (me.y - hcy) * V4_HSCALE;
                        exo = !(fabsf(fdir) < 540.f);              // exotic raw heading: the generic code for every round
                        S.fmeta[slot] = make_uint4(v4_pack_half2(hxs, hxs), v4_pack_half2(hys, hys), exo ? V4_META_EXOTIC : 0u, (uint32_t)i);
                    }
                    any_exotic |= __ballot_sync(FULL, exo);
                    nf0 += __popc(__ballot_sync(FULL, active && i < mid));
                    nf += __popc(amask);
                    pos += take;
                }
                __syncwarp();
                if (nf == 0) continue;
                if (STATS) st_focal += (unsigned)nf;

                // the reduced sums of every focal fish go to shared memory (one lane stores them): raw for a single window batch,
                // unbiased 64-bit running sums when the stencil needs several batches
                if (nbatch > 1 && lane < nf) {
#pragma unroll
                    for (int k = 0; k < 6; ++k) S.mb.acc[lane][k] = 0;
                    S.mb.cnt[lane] = make_int4(0, 0, 0, 0);
                }
                for (int b = 0; b < nbatch; ++b) {
                    const int fill = (staged == b) ? min(T - b * V4_WIN, V4_WIN) : stage(b);
                    // the group's fish of the item's first cell come first (fish are cell-sorted): two runs of fish, each with its own
                    // range of filter steps (64 candidates each)
                    for (int part = 0; part < 2; ++part) {
                    const int f_lo = part ? nf0 : 0, f_hi = part ? nf : nf0;
                    int st_b = 0, st_e = (fill + 63) >> 6;
                    if (nbatch == 1) { st_b = part ? st_beg1 : 0; st_e = part ? st_end1 : st_end0; }
                    for (int f = f_lo; f < f_hi; ++f) {
                        // ---- filter: two candidates per lane and step, packed fp16, conservative disc test
                        const uint4 M = S.fmeta[f];   // broadcast
                        const __half2 fx2 = *reinterpret_cast<const __half2*>(&M.x), fy2 = *reinterpret_cast<const __half2*>(&M.y);
                        int qn = 0;
#pragma unroll
                        for (int st = 0; st < V4_SLOTS / 64; ++st) {
                            if (st < st_b || st >= st_e) continue;   // (warp-uniform)
                            const uint32_t cxb = S.hx[st * 32 + lane], cyb = S.hy[st * 32 + lane];
                            const __half2 dx = __hsub2(*reinterpret_cast<const __half2*>(&cxb), fx2);
                            const __half2 dy = __hsub2(*reinterpret_cast<const __half2*>(&cyb), fy2);
                            const __half2 d2 = __hfma2(dx, dx, __hmul2(dy, dy));
                            bool p0, p1;
                            v4_setp_lt2(*reinterpret_cast<const uint32_t*>(&d2), r2h_bits, p0, p1);
                            const unsigned b0 = __ballot_sync(FULL, p0), b1 = __ballot_sync(FULL, p1);
                            const int n0 = __popc(b0);
                            if (p0) S.queue[qn + __popc(b0 & lt_mask)] = (unsigned char)(lane + 64 * st);
                            if (p1) S.queue[qn + n0 + __popc(b1 & lt_mask)] = (unsigned char)(lane + 64 * st + 32);
                            qn += n0 + __popc(b1);
                        }
                        S.queue[qn + lane] = (unsigned char)(V4_SLOTS - 1);   // pad the last round with the far-away slot
                        __syncwarp();
                        // candidates in SURVEY's sense: the fish of the 3 x 3 cells around the focal fish's own cell (not the
                        // whole pair window)
                        if (STATS) st_cand += (unsigned)(nbatch > 1 ? fill : (part ? T - pre3 : pre9));
                        // ---- pair rounds (compacted pairs, one per lane, same focal fish in every lane)
                        const float4 F = S.frec[f];  // broadcast
                        PairAccR P;
                        pair_acc_clear(P);
                        const bool exotic = any_exotic != 0u && __any_sync(FULL, (M.z & V4_META_EXOTIC) != 0u);   // (warp-uniform)
                        unsigned dirty = exotic ? 1u : 0u;
                        if (!exotic && qn > 0) {
                            const unsigned char* qp = S.queue + lane;
                            int q0 = 0;
                            do {
                                const float4 Q = S.win[(int)*qp];
                                if (pair_fast<true>(F.x, F.y, F.z, F.w, Q.x, Q.y, Q.z, Q.w, P)) dirty = 1u;
                                q0 += 32; qp += 32;
                            } while (q0 < qn);
                        }
                        if (__any_sync(FULL, dirty != 0u)) {   // rare: redo this fish's rounds with the generic per-lane code
                            pair_acc_clear(P);
                            const int fi = (int)M.w;
                            const float fdir = a.cold0[fi].x;
                            for (int q0 = 0; q0 < qn; q0 += 32) {
                                const int j = (int)S.queue[q0 + lane];
                                const float4 Q = S.win[j];
                                const bool unc = pair_round_generic<true>(
                                    q0 + lane < qn, F.x, F.y, F.z, F.w, fdir, exotic, Q.x, Q.y, Q.z, Q.w, a.seed, a.replicate, a.step_move,
                                    [&](float& rx, float& ry, int& kcode, float&, float& fyr) {
                                        int src, k;
                                        locate(b, j, src, k);
                                        const float4 raw = __ldg(&a.rec[src]);
                                        rx = raw.x; ry = raw.y;
                                        const int ky = (k * 11) >> 5, kx = k - 3 * ky;   // k / 3, k % 3 for k < 12
                                        kcode = kx + 3 * (ky - part);                    // (kx - 1 + 1) + 3 * (ky - (1 + fc) + 1)
                                        fyr = __ldg(&a.rec[fi]).y;
                                    },
                                    [&](uint32_t& my, uint32_t& nb) {
                                        int src, k;
                                        locate(b, j, src, k);
                                        nb = (uint32_t)a.cold1[src].w;
                                        my = (uint32_t)a.cold1[fi].w;
                                    }, P);
                                if (STATS) st_rechk += __popc(__ballot_sync(FULL, unc));
                            }
                        }
                        // ---- reduce the lanes' partial sums; lane 0 files them under focal fish f
                        const unsigned r0 = __reduce_add_sync(FULL, P.ax0), r1 = __reduce_add_sync(FULL, P.ay0);
                        const unsigned r2 = __reduce_add_sync(FULL, P.ax1), r3 = __reduce_add_sync(FULL, P.ay1);
                        const unsigned r4 = __reduce_add_sync(FULL, P.ax2), r5 = __reduce_add_sync(FULL, P.ay2);
                        const unsigned cw = __reduce_add_sync(FULL, P.c_far);
                        if (STATS) st_inter += (unsigned)((cw >> 16) & 0xFFu);
                        if (lane == 0) {
                            if (nbatch == 1) {
                                *reinterpret_cast<uint4*>(&S.res[f][0]) = make_uint4(r0, r1, r2, r3);
                                *reinterpret_cast<uint4*>(&S.res[f][4]) = make_uint4(r4, r5, cw, 0u);
                            } else {
                                const unsigned r[6] = {r0, r1, r2, r3, r4, r5};
                                long long sm[6];
                                const int4 cc = unpack_counts_r<true>(cw, 0u);
                                pair_sums_unbias(r, cc, sm);
#pragma unroll
                                for (int k = 0; k < 6; ++k) S.mb.acc[f][k] += sm[k];
                                int4 o = S.mb.cnt[f];
                                o.x += cc.x; o.y += cc.y; o.z += cc.z; o.w += cc.w;
                                S.mb.cnt[f] = o;
                            }
                        }
                    }
                    }
                }
                // ---- epilogue, one focal fish per lane
                __syncwarp();
                if (lane < nf) {
                    const int idx = (int)S.fmeta[lane].w;
                    const int4 k1 = a.cold1[idx];
                    const float4 F = a.rec[idx];   // (the unshifted offset: sample()'s food cell)
                    if (nbatch > 1) {
                        const int4 cc = S.mb.cnt[lane];
                        finish_focal(a, idx, k1, F.x, F.y, cc.x, cc.y, cc.z, cc.w, &S.mb.acc[lane][0], 0);
                    } else {
                        const uint4 ra = *reinterpret_cast<const uint4*>(&S.res[lane][0]), rb = *reinterpret_cast<const uint4*>(&S.res[lane][4]);
                        const unsigned r[6] = {ra.x, ra.y, ra.z, ra.w, rb.x, rb.y};
                        long long sums[6];
                        const int4 cc = unpack_counts_r<true>(rb.z, 0u);
                        pair_sums_unbias(r, cc, sums);
                        finish_focal(a, idx, k1, F.x, F.y, cc.x, cc.y, cc.z, cc.w, sums, 0);
                    }
                }
                __syncwarp();
            }
            left -= nfc; cy0 += nfc; nfc = left;
        }
    }
    if (STATS && a.stats && lane == 0) {
        atomicAdd(&a.stats[0], st_focal); atomicAdd(&a.stats[1], st_cand);
        atomicAdd(&a.stats[2], st_inter); atomicAdd(&a.stats[3], st_rechk);
        // how evenly the dispenser kept the warps busy (FISHABM_TAIL_DEBUG prints it): last exit, summed busy time, earliest
        // start (as a maximum of the complement), warps
        unsigned long long t1;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
        atomicMax(&a.stats[4], t1); atomicAdd(&a.stats[5], t1 - st_t0); atomicMax(&a.stats[6], ~st_t0); atomicAdd(&a.stats[7], 1ull);
    }
}

}  // namespace fishk
