// K4, third design (default): candidate-parallel neighbour kernel with a packed-half filter and a branch-free pair round.
//
// Same decomposition as the second design (one warp per focal cell, the 3x3 stencil staged once per cell
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

#ifndef FISHABM_V3_MIN_CTAS
#define FISHABM_V3_MIN_CTAS 4        // resident CTAs per SM the register allocation is held to (4: 64 registers, 32 warps per SM)
#endif
constexpr int V3_WARPS = 8;
constexpr int V3_THREADS = V3_WARPS * 32;
constexpr int V3_SLOTS = 256;         // window slots (queue entries are bytes); the last one is a permanent far-away candidate
constexpr int V3_WIN = V3_SLOTS - 1;  // candidates per window batch: <= 255, so that the 8-bit packed counts cannot overflow
constexpr int V3_FMAX = 16;           // focal fish per group
constexpr float V3_FAR_X = 1.0e18f, V3_FAR_Y = 3.0e17f;   // padding candidate: beyond every radius, +inf in fp16; not on a lattice bearing
constexpr float V3_HSCALE = 0.25f, V3_HCENTRE = 100.f;
constexpr float V3_R2H = 2528.f;

struct V3WarpSmem {
    float4 win[V3_SLOTS];                                  // window candidates in the focal cell's frame: x, y, cos, sin
    uint32_t hx[V3_SLOTS / 2], hy[V3_SLOTS / 2];           // the same coordinates as fp16 pairs: element 32*s + l = candidates (64*s + l, 64*s + 32 + l)
    union {
        uint32_t res[V3_FMAX][8];                          // one window batch: the reduced raw sums (6) and packed counts of every focal fish
        long long acc[V3_FMAX][6];                         // several batches: exact cos/sin sums (avoid x,y, align x,y, attract x,y)
    };
    int4 cnt[V3_FMAX];                                     // several batches: n15, n100, n200, front
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

// K5 for the fish [s0, e0) of cell c, one fish per lane.  Deliberately NOT inlined: its registers (Philox, sincospi, the
// feeding-rank loop) must not weigh on the allocation of the pair loops of the kernel that calls it once per cell.
__device__ __noinline__ void v3_update_cell(const UpdateArgs* u, int s0, int e0, int c, int lane) {
    for (int base = s0; base < e0; base += 32) {
        const int i = base + lane;
        if (i < e0) update_fish(*u, i, c);
    }
}

// FUSE: the warp that has finished the neighbour pass of a cell also runs the per-fish update (K5: sample / feed / move) of
// that cell's fish right away -- their pass outputs are still in L1 / L2, the separate update kernel and its re-read of the
// whole state disappear.  The moved records go to a scratch array (u.rec_out): `rec` must stay as it is while other warps
// read it as candidates.  Everything else the update touches belongs to the cell's own fish or its own 10 x 10 food cells.
template <bool STATS, bool FUSE>
__global__ void __launch_bounds__(V3_THREADS, FISHABM_V3_MIN_CTAS) neighbour_kernel_v3(const NeighArgs a, const __grid_constant__ UpdateArgs u) {
    extern __shared__ __align__(16) unsigned char v3_smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const unsigned FULL = 0xFFFFFFFFu;
    const unsigned lt_mask = (1u << lane) - 1u;
    V3WarpSmem& S = reinterpret_cast<V3WarpSmem*>(v3_smem_raw)[warp];

    unsigned long long st_cand = 0, st_inter = 0, st_rechk = 0, st_focal = 0;
    if (blockIdx.x == 0 && threadIdx.x == 0) *a.work_counter_next = 0;
    if (FUSE && u.slab && u.do_move) {   // last step's ghosts are dropped by the coming sort (update_kernel does this when unfused)
        const int total = a.cell_start[a.g.ncell], b1 = a.cell_start[a.g.own_c0], e2 = a.cell_start[a.g.own_c1];
        for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < b1 + (total - e2); t += gridDim.x * blockDim.x)
            u.next_cell[t < b1 ? t : e2 + (t - b1)] = -1;
    }
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
        const float ysh = yframe_shift(c % a.g.ny);   // the pair arithmetic's y frame (kernels.cuh)

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
            const float4 far = make_float4(V3_FAR_X, V3_FAR_Y, 1.f, 0.f);
            // range-major copy: each of the 9 stencil cells is a contiguous run of sorted records, shifted into the focal
            // cell's frame by a per-cell constant (the periodic minimum image).  Three cells (one stencil row) at a time so
            // that three loads are in flight; a cell with more than 32 fish in this batch takes the trailing loop.
#pragma unroll 1
            for (int ky = 0; ky < 3; ++ky) {
                float4 r[3];
                int u[3], hi[3], src[3];
#pragma unroll
                for (int kx = 0; kx < 3; ++kx) {
                    const int k = 3 * ky + kx;
                    const int pk = S.run_pre[k];
                    u[kx] = max(pk, base) + lane; hi[kx] = min(S.run_pre[k + 1], base + fill); src[kx] = S.run_start[k] - pk;
                    r[kx] = far;   // (initialised: an undefined value would be carried, through local memory, across the ky iterations)
                    if (u[kx] < hi[kx]) r[kx] = __ldg(&a.rec[src[kx] + u[kx]]);
                }
                const float sy = (float)(ky - 1) * CELL_F + ysh;
#pragma unroll
                for (int kx = 0; kx < 3; ++kx) {
                    const float sx = (float)(kx - 1) * CELL_F;
                    if (u[kx] < hi[kx]) S.win[u[kx] - base] = make_float4(r[kx].x + sx, r[kx].y + sy, r[kx].z, r[kx].w);
                    for (int v = u[kx] + 32; v < hi[kx]; v += 32) {
                        const float4 q = __ldg(&a.rec[src[kx] + v]);
                        S.win[v - base] = make_float4(q.x + sx, q.y + sy, q.z, q.w);
                    }
                }
            }
            for (int j = fill + lane; j < padded; j += 32) S.win[j] = far;
            if (lane == 0) S.win[V3_SLOTS - 1] = far;
            __syncwarp();
            // the same coordinates as packed fp16 pairs for the filter: element 32*s + l = candidates (64*s + l, 64*s + 32 + l)
            for (int j0 = 0; j0 < padded; j0 += 64) {
                const float2 p = *reinterpret_cast<const float2*>(&S.win[j0 + lane]), q = *reinterpret_cast<const float2*>(&S.win[j0 + 32 + lane]);
                S.hx[(j0 >> 1) + lane] = v3_pack_half2(fmaf(p.x, V3_HSCALE, -V3_HCENTRE * V3_HSCALE), fmaf(q.x, V3_HSCALE, -V3_HCENTRE * V3_HSCALE));
                S.hy[(j0 >> 1) + lane] = v3_pack_half2(fmaf(p.y, V3_HSCALE, -V3_HCENTRE * V3_HSCALE), fmaf(q.y, V3_HSCALE, -V3_HCENTRE * V3_HSCALE));
            }
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
                    float4 me = a.rec[i];
                    me.y += ysh;
                    const int slot = nf + __popc(amask & lt_mask);
                    S.frec[slot] = me; S.fdir[slot] = a.cold0[i].x; S.fidx[slot] = i;
                    const float hxs = fmaf(me.x, V3_HSCALE, -V3_HCENTRE * V3_HSCALE), hys = fmaf(me.y, V3_HSCALE, -V3_HCENTRE * V3_HSCALE);
                    S.hfx[slot] = v3_pack_half2(hxs, hxs);
                    S.hfy[slot] = v3_pack_half2(hys, hys);
                }
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
                for (int k = 0; k < 6; ++k) S.acc[lane][k] = 0;
                S.cnt[lane] = make_int4(0, 0, 0, 0);
            }
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
                    auto raw_of = [&](int j, float& rx, float& ry, int& kcode, float& fyr) {
                        int src;
                        locate(b, j, src, kcode);
                        const float4 raw = __ldg(&a.rec[src]);
                        rx = raw.x; ry = raw.y;
                        fyr = __ldg(&a.rec[S.fidx[f]]).y;
                    };
                    auto ids_of = [&](int j, uint32_t& my, uint32_t& nb) {
                        int src, k;
                        locate(b, j, src, k);
                        nb = (uint32_t)a.cold1[src].w;
                        my = (uint32_t)a.cold1[S.fidx[f]].w;
                    };
                    const bool exotic = !(fabsf(fdir) < 540.f);   // exotic raw heading: the generic code for every round
                    const unsigned char* qp = S.queue + lane;
                    for (int q0 = 0; q0 < qn; q0 += 32, qp += 32) {
                        const int j = (int)*qp;
                        const float4 Q = S.win[j];
                        const bool unc = pair_round<true>(
                            q0 + lane < qn, exotic, F.x, F.y, F.z, F.w, fdir, Q.x, Q.y, Q.z, Q.w, a.seed, a.replicate, a.step_move,
                            [&](float& rx, float& ry, int& kcode, float&, float& fyr) { raw_of(j, rx, ry, kcode, fyr); },
                            [&](uint32_t& my, uint32_t& nb) { ids_of(j, my, nb); }, P);
                        if (STATS) st_rechk += __popc(__ballot_sync(FULL, unc));
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
                            for (int k = 0; k < 6; ++k) S.acc[f][k] += sm[k];
                            int4 o = S.cnt[f];
                            o.x += cc.x; o.y += cc.y; o.z += cc.z; o.w += cc.w;
                            S.cnt[f] = o;
                        }
                    }
                }
            }
            // ---- epilogue, one focal fish per lane
            __syncwarp();
            if (lane < nf) {
                const int idx = S.fidx[lane];
                const int4 k1 = a.cold1[idx];
                const float4 F = a.rec[idx];   // (the unshifted offset: sample()'s food cell)
                if (nbatch > 1) {
                    const int4 cc = S.cnt[lane];
                    finish_focal(a, idx, k1, F.x, F.y, cc.x, cc.y, cc.z, cc.w, &S.acc[lane][0], 0);
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
        if (FUSE) {   // K5 for this cell's fish (the host only fuses with split == 1: the whole cell is this warp's)
            __syncwarp();
            v3_update_cell(&u, s0, e0, c, lane);
        }
    }
    if (STATS && a.stats && lane == 0) {
        atomicAdd(&a.stats[0], st_focal); atomicAdd(&a.stats[1], st_cand);
        atomicAdd(&a.stats[2], st_inter); atomicAdd(&a.stats[3], st_rechk);
    }
}

}  // namespace fishk
