// bellman_tile.cu — D = 3 / 4 stage with the J_{k+1} neighbourhood of a state tile staged in shared
// memory by ONE TMA box (cp.async.bulk.tensor.{4,5}d): the kernel for Solver_pos_att's 4-D channel
// sweep (pos-att/Solver_pos_att.m:244-297; min(..., [], 5) at :272).
//
// Applicability is decided on the host (tile_setup): every dimension's query must be a bounded
// stencil of the state's own index — cell(x'_d) - i_d in [lo_d, hi_d] for every state and control —
// which holds for the reference's dynamics (x' = x + h v, v' = v + h a(u), ...: a fraction of a
// cell per stage).  A CTA owns a tile of T0 x T1 x T2 x T3 states (T0 = 32 = one warp along the
// contiguous dimension) and stages the box  prod_d (T_d + hi_d - lo_d + 1)  once; the 2^D corner
// gathers of every (state, control) pair are then shared-memory loads instead of L2 sectors, and
// the SEARCH locate starts from the state's own cell instead of the bucket table.
//
// Arithmetic: the normative operations of include/bellman.h on the same operands as k_stage_direct
// (same association, dimension 0 lerped first), so results are bit-identical.  -fmad=false.
#include <cuda.h>
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <string>
#include <vector>

#include "bellman_handle.h"
#include "bellman_internal.h"

namespace bellman {

namespace {

constexpr int TNT = 256;        // threads per CTA
constexpr int TT0 = 32;         // tile extent along dimension 0 (one warp)

struct TileParams {
    int T[MAXD];                // tile extents (T[0] = 32; unused dims 1)
    int ntile[MAXD];
    int lo[MAXD];               // box origin relative to the tile origin (<= 0), dimension 0 before even-rounding
    int box[MAXD];              // box extents
    int bstride[MAXD];          // element strides inside the box
    int box_elems;
    int own_stride[MAXD];       // strides of the owned index space (idx_out)
    // Control-dependent dimensions whose query depends only on the state's own index and the control
    // (x'_d = Ta_d[i_d] (+ Tb_d[i_d]) + Tc_d[c]: v' = v + h a(u), w' = w + h alpha(u)) are located once
    // per handle: lt[d][(p * n_d + i) * C + c] = {t, cell as raw bits}; null for the other dimensions
    const double2 *lt[MAXD];
    int pf_dist;                // k_stage_tile_pa: L2 prefetch distance in CTAs (0 = off)
    // k_stage_tile_pa: the control-independent dimensions d = 0, 2 (x'_d = Ta_d[i_d] + Tb_d[i_{d+1}]) are
    // located once per handle too: ft[d/2][(p * n_{d+1} + i_{d+1}) * n_d + i_d] = {t, box byte offset}
    const double2 *ft[2];
    int q_identity;             // q_order == {0, 1, 2, 3}
};

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void tma_load_4d(void *dst, const CUtensorMap *map, uint64_t *bar, int c0, int c1, int c2,
                                            int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::
            "r"(smem_u32(dst)),
        "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
}
__device__ __forceinline__ void tma_load_5d(void *dst, const CUtensorMap *map, uint64_t *bar, int c0, int c1, int c2,
                                            int c3, int c4) {
    asm volatile(
        "cp.async.bulk.tensor.5d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];" ::
            "r"(smem_u32(dst)),
        "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
        : "memory");
}
// TMA prefetch of a box into L2 (no shared memory, no barrier)
__device__ __forceinline__ void tma_prefetch_5d(const CUtensorMap *map, int c0, int c1, int c2, int c3, int c4) {
    asm volatile("cp.async.bulk.prefetch.tensor.5d.L2.global.tile [%0, {%1, %2, %3, %4, %5}];" ::"l"(map), "r"(c0), "r"(c1),
                 "r"(c2), "r"(c3), "r"(c4)
                 : "memory");
}
__device__ __forceinline__ double lds_f64(uint32_t addr) {
    double v;
    asm("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr));
    return v;
}

// locate_d of include/bellman.h.  UNIFORM: x is the fractional cell coordinate.  SEARCH: the exact
// bin rule cell = clamp(#{ s[i] <= x } - 1, 0, n-2), found by walking from `guess` (the state's own
// cell: the answer is a cell or two away) — the same cell the bucket-table search returns.
__device__ __forceinline__ int locate_near(const double *__restrict__ s, const double *__restrict__ rinv, int n, int mode,
                                           int guess, double x, double &t) {
    int cell;
    if (mode == BELLMAN_LOCATE_UNIFORM) {
        cell = min(max(__double2int_rd(x), 0), n - 2);
        t = x - (double)cell;
    } else {
        cell = min(max(guess, 0), n - 2);
        while (cell > 0 && x < __ldg(s + cell)) --cell;
        while (cell < n - 2 && __ldg(s + cell + 1) <= x) ++cell;
        t = (x - __ldg(s + cell)) * __ldg(rinv + cell);
    }
    return cell;
}

template <int D>
__device__ __forceinline__ int pick(const int (&gi)[D], int k) {
    int v = gi[0];
#pragma unroll
    for (int d = 1; d < D; ++d) v = (k == d) ? gi[d] : v;
    return v;
}

template <int D>
__global__ void __launch_bounds__(TNT, 3)
k_stage_tile(const __grid_constant__ StageParams sp, const __grid_constant__ TileParams tp,
             const __grid_constant__ CUtensorMap tmap) {
    extern __shared__ __align__(128) double box[];
    __shared__ __align__(8) uint64_t mbar;

    const int tid = threadIdx.x, lane = tid & 31, wrp = tid >> 5;
    const int prob = blockIdx.z;
    // tile indices: x = t0 + ntile0 * t1, y = t2 + ntile2 * t3
    int ti[MAXD] = {0, 0, 0, 0};
    ti[0] = blockIdx.x % tp.ntile[0];
    ti[1] = blockIdx.x / tp.ntile[0];
    ti[2] = blockIdx.y % tp.ntile[2];
    ti[3] = blockIdx.y / tp.ntile[2];
    int t_lo[D], t_hi[D], org[D];
#pragma unroll
    for (int d = 0; d < D; ++d) {
        t_lo[d] = sp.dim[d].own_lo + ti[d] * tp.T[d];
        t_hi[d] = min(t_lo[d] + tp.T[d], sp.dim[d].own_lo + sp.dim[d].own_n);
        org[d] = t_lo[d] + tp.lo[d];
    }
    org[0] -= (org[0] - sp.dim[0].ext_lo) & 1;        // TMA: 16-byte aligned innermost coordinate

    if (tid == 0) {
        mbar_init(&mbar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        mbar_expect_tx(&mbar, (uint32_t)tp.box_elems * 8u);
        if (D == 4)
            tma_load_5d(box, &tmap, &mbar, org[0] - sp.dim[0].ext_lo, org[1] - sp.dim[1].ext_lo, org[2] - sp.dim[2].ext_lo,
                        org[D - 1] - sp.dim[D - 1].ext_lo, prob);
        else
            tma_load_4d(box, &tmap, &mbar, org[0] - sp.dim[0].ext_lo, org[1] - sp.dim[1].ext_lo, org[2] - sp.dim[2].ext_lo,
                        prob);
    }

    // per-problem tables
    const double *grid[D], *rinv[D], *Ta[D], *Tb[D], *Tc[D], *q[D];
    int mode[D], n[D];
#pragma unroll
    for (int d = 0; d < D; ++d) {
        const DimParams &dp = sp.dim[d];
        n[d] = dp.n;
        grid[d] = dp.grid + (size_t)prob * dp.n;
        rinv[d] = dp.rinv + (size_t)prob * dp.n;
        Ta[d] = dp.Ta + (size_t)prob * dp.n_a;
        Tb[d] = dp.Tb ? dp.Tb + (size_t)prob * dp.n_b : nullptr;
        Tc[d] = dp.Tc ? dp.Tc + (size_t)prob * sp.C : nullptr;
        q[d] = dp.q + (size_t)prob * dp.n;
        mode[d] = __ldg(dp.mode + prob);
    }
    const double *rr = sp.r + (size_t)prob * sp.C;

    // byte offsets of the 2^(D-1) dimension-0 pairs relative to the base corner
    uint32_t ro[1 << (D - 1)];
#pragma unroll
    for (int k = 0; k < (1 << (D - 1)); ++k) {
        int r = 0;
#pragma unroll
        for (int d = 1; d < D; ++d)
            if (k & (1 << (d - 1))) r += tp.bstride[d];
        ro[k] = 8u * (uint32_t)r;
    }

    __syncthreads();          // mbarrier initialised
    mbar_wait(&mbar, 0);
    uint32_t base = smem_u32(box);
    asm volatile("" : "+r"(base)::"memory");   // box loads depend on `base`: none is scheduled above the wait

    const int i0 = t_lo[0] + lane;
    const bool row_ok = i0 < t_hi[0];
    const int g0 = row_ok ? i0 : t_hi[0] - 1;
    const int nsub = (D == 4 ? tp.T[1] * tp.T[2] * tp.T[3] : tp.T[1] * tp.T[2]);

#pragma unroll 1
    for (int s = wrp; s < nsub; s += TNT / 32) {
        int gi[D];
        gi[0] = g0;
        {
            int r = s;
            gi[1] = t_lo[1] + r % tp.T[1];
            r /= tp.T[1];
            if (D == 4) {
                gi[2] = t_lo[2] + r % tp.T[2];
                gi[D - 1] = t_lo[D - 1] + r / tp.T[2];
            } else {
                gi[2] = t_lo[2] + r;
            }
        }
        bool ok = row_ok;
#pragma unroll
        for (int d = 1; d < D; ++d) ok = ok && gi[d] < t_hi[d];
        // warp-uniform for d >= 1: skip sub-columns outside a ragged tile
        if (__all_sync(0xffffffffu, !ok)) continue;
#pragma unroll
        for (int d = 1; d < D; ++d) gi[d] = min(gi[d], t_hi[d] - 1);

        // state terms (include/bellman.h: base_d = Ta + Tb, gs = sum of q in q_order)
        double bq[D];
#pragma unroll
        for (int d = 0; d < D; ++d) {
            double b = __ldg(Ta[d] + pick<D>(gi, sp.dim[d].src_a));
            if (Tb[d]) b = b + __ldg(Tb[d] + pick<D>(gi, sp.dim[d].src_b));
            bq[d] = b;
        }
        double gs;
        {
            const int o0 = sp.q_order[0];
            const double *qp = q[0];
#pragma unroll
            for (int d = 1; d < D; ++d) qp = (o0 == d) ? q[d] : qp;
            gs = __ldg(qp + pick<D>(gi, o0));
#pragma unroll
            for (int m = 1; m < D; ++m) {
                const int om = sp.q_order[m];
                const double *qm = q[0];
#pragma unroll
                for (int d = 1; d < D; ++d) qm = (om == d) ? q[d] : qm;
                gs = gs + __ldg(qm + pick<D>(gi, om));
            }
        }
        // control-independent dimensions: located once per state
        const double2 *lt[D];
#pragma unroll
        for (int d = 0; d < D; ++d)
            lt[d] = tp.lt[d] ? tp.lt[d] + ((size_t)prob * n[d] + gi[d]) * sp.C : nullptr;
        double tf[D];
        uint32_t off_fixed = base;
#pragma unroll
        for (int d = 0; d < D; ++d) {
            tf[d] = 0.0;
            if (!Tc[d]) {
                const int cell = locate_near(grid[d], rinv[d], n[d], mode[d], gi[d], bq[d], tf[d]);
                off_fixed += 8u * (uint32_t)((cell - org[d]) * tp.bstride[d]);
            }
        }

        double best = __longlong_as_double(0x7ff0000000000000LL);
        int arg = 0;
#pragma unroll 2
        for (int c = 0; c < sp.C; ++c) {
            double t[D];
            uint32_t o = off_fixed;
#pragma unroll
            for (int d = 0; d < D; ++d) {
                if (lt[d]) {
                    const double2 e = __ldg(lt[d] + c);
                    t[d] = e.x;
                    o += 8u * (uint32_t)((__double2loint(e.y) - org[d]) * tp.bstride[d]);
                } else if (Tc[d]) {
                    const double xq = bq[d] + __ldg(Tc[d] + c);
                    const int cell = locate_near(grid[d], rinv[d], n[d], mode[d], gi[d], xq, t[d]);
                    o += 8u * (uint32_t)((cell - org[d]) * tp.bstride[d]);
                } else {
                    t[d] = tf[d];
                }
            }
            double v[1 << D];
#pragma unroll
            for (int k = 0; k < (1 << (D - 1)); ++k) {
                v[2 * k] = lds_f64(o + ro[k]);
                v[2 * k + 1] = lds_f64(o + ro[k] + 8);
            }
#pragma unroll
            for (int d = 0; d < D; ++d)               // dimension 0 reduced first
#pragma unroll
                for (int m = 0; m < (1 << (D - 1 - d)); ++m)
                    v[m] = fma(t[d], v[2 * m + 1] - v[2 * m], v[2 * m]);
            const double tot = (gs + __ldg(rr + c)) + v[0];
            if (tot < best) { best = tot; arg = c; }
        }
        if (ok) {
            long long jo = (long long)prob * sp.S_ext, io = (long long)prob * sp.S_own;
#pragma unroll
            for (int d = 0; d < D; ++d) {
                jo += (long long)(gi[d] - sp.dim[d].ext_lo) * sp.dim[d].stride;
                io += (long long)(gi[d] - sp.dim[d].own_lo) * tp.own_stride[d];
            }
            sp.J_out[jo] = best;
            idx_store(sp.idx_out, sp.idx_bytes, io, arg);
            if (sp.n_peers) peer_store<D>(sp, prob, gi, best);
        }
    }
}

// ---------------------------------------------------------------------------------------------
// k_stage_tile_pa: the tile kernel specialised for the structure of Solver_pos_att's channel
// (pos-att/Solver_pos_att.m:299-328): two coupled (position, rate) blocks,
//     x'_0 = Ta_0[i0] + Tb_0[i1]      (control independent)      x'_1 = Ta_1[i1] + Tc_1[c]   (table)
//     x'_2 = Ta_2[i2] + Tb_2[i3]      (control independent)      x'_3 = Ta_3[i3] + Tc_3[c]   (table)
// A warp owns (i2, i3) pairs of the tile and walks i1; lanes run along i0.  What depends on (i2, i3)
// only — the dimension-2 cell/weight, the dimension-3 table row — is hoisted out of the i1 loop;
// per state a thread locates dimension 0 once; per (state, control) it reads two table entries
// {t, box byte offset}, 16 shared-memory corners and does the 15 lerps.  No run-time structure
// checks inside the loops (k_stage_tile: 229 instructions per update, this kernel: see profiles/).
// Same operations on the same operands as k_stage_direct ⇒ bit-identical results.
// ---------------------------------------------------------------------------------------------
template <int NT, int UNROLL>
__global__ void __launch_bounds__(NT, 2)
k_stage_tile_pa(const __grid_constant__ StageParams sp, const __grid_constant__ TileParams tp,
                const __grid_constant__ CUtensorMap tmap) {
    constexpr int D = 4;
    extern __shared__ __align__(128) double box[];
    __shared__ __align__(8) uint64_t mbar;

    const int tid = threadIdx.x, lane = tid & 31, wrp = tid >> 5;
    const uint32_t prob = blockIdx.z;
    int ti[D];
    ti[0] = blockIdx.x % tp.ntile[0];
    ti[1] = blockIdx.x / tp.ntile[0];
    ti[2] = blockIdx.y % tp.ntile[2];
    ti[3] = blockIdx.y / tp.ntile[2];
    int t_lo[D], t_hi[D], org[D];
#pragma unroll
    for (int d = 0; d < D; ++d) {
        t_lo[d] = sp.dim[d].own_lo + ti[d] * tp.T[d];
        t_hi[d] = min(t_lo[d] + tp.T[d], sp.dim[d].own_lo + sp.dim[d].own_n);
        org[d] = t_lo[d] + tp.lo[d];
    }
    org[0] -= (org[0] - sp.dim[0].ext_lo) & 1;

    if (tid == 0) {
        mbar_init(&mbar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        mbar_expect_tx(&mbar, (uint32_t)tp.box_elems * 8u);
        tma_load_5d(box, &tmap, &mbar, org[0] - sp.dim[0].ext_lo, org[1] - sp.dim[1].ext_lo, org[2] - sp.dim[2].ext_lo,
                    org[3] - sp.dim[3].ext_lo, (int)prob);
        // warm L2 with the box of the tile pf_dist CTAs ahead in launch order (about one wave): the CTA
        // that takes this one's place then waits for L2, not DRAM
        if (tp.pf_dist) {
            const unsigned lin = blockIdx.x + gridDim.x * (blockIdx.y + gridDim.y * blockIdx.z) + (unsigned)tp.pf_dist;
            if (lin < gridDim.x * gridDim.y * gridDim.z) {
                const unsigned bx = lin % gridDim.x, rest = lin / gridDim.x;
                const unsigned by = rest % gridDim.y, pz = rest / gridDim.y;
                const int p0 = (int)(bx % tp.ntile[0]), p1 = (int)(bx / tp.ntile[0]);
                const int p2 = (int)(by % tp.ntile[2]), p3 = (int)(by / tp.ntile[2]);
                int o0 = p0 * tp.T[0] + tp.lo[0] + sp.dim[0].own_lo - sp.dim[0].ext_lo;
                o0 -= o0 & 1;
                tma_prefetch_5d(&tmap, o0, p1 * tp.T[1] + tp.lo[1] + sp.dim[1].own_lo - sp.dim[1].ext_lo,
                                p2 * tp.T[2] + tp.lo[2] + sp.dim[2].own_lo - sp.dim[2].ext_lo,
                                p3 * tp.T[3] + tp.lo[3] + sp.dim[3].own_lo - sp.dim[3].ext_lo, (int)pz);
            }
        }
    }

    const int C = sp.C;
    const DimParams &d0 = sp.dim[0], &d1 = sp.dim[1], &d2 = sp.dim[2], &d3 = sp.dim[3];
    const double *rr = sp.r + prob * (uint32_t)C;

    // this lane's row (dimension 0): table entries that depend on i0 only
    const int i0 = min(t_lo[0] + lane, t_hi[0] - 1);
    const bool row_ok = t_lo[0] + lane < t_hi[0];
    const double q0 = __ldg(d0.q + prob * (uint32_t)d0.n + i0);
    const int qo0 = sp.q_order[0], qo1 = sp.q_order[1], qo2 = sp.q_order[2], qo3 = sp.q_order[3];

    // byte offsets of the 8 dimension-0 pairs relative to the base corner
    uint32_t ro[8];
#pragma unroll
    for (int k = 0; k < 8; ++k)
        ro[k] = 8u * (uint32_t)((k & 1 ? tp.bstride[1] : 0) + (k & 2 ? tp.bstride[2] : 0) + (k & 4 ? tp.bstride[3] : 0));

    __syncthreads();
    mbar_wait(&mbar, 0);
    uint32_t base = smem_u32(box);
    asm volatile("" : "+r"(base)::"memory");
    // fold the box origin of every dimension into the base address (table offsets are absolute cells * stride)
    base -= 8u * (uint32_t)(org[0] * tp.bstride[0] + org[1] * tp.bstride[1] + org[2] * tp.bstride[2] + org[3] * tp.bstride[3]);

    const int n1t = t_hi[1] - t_lo[1];
    const int npair = tp.T[2] * tp.T[3];
#pragma unroll 1
    for (int pq = wrp; pq < npair; pq += NT / 32) {
        const int i2 = t_lo[2] + pq % tp.T[2], i3 = t_lo[3] + pq / tp.T[2];
        if (i2 >= t_hi[2] || i3 >= t_hi[3]) continue;          // warp-uniform: ragged tile
        // dimension 2 (control independent): one table entry per (i2, i3) pair
        const double2 e2 = __ldg(tp.ft[1] + (prob * (uint32_t)d3.n + (uint32_t)i3) * (uint32_t)d2.n + (uint32_t)i2);
        const double t2 = e2.x;
        const uint32_t off2 = base + (uint32_t)__double2loint(e2.y);
        const double2 *lt3 = tp.lt[3] + (prob * (uint32_t)d3.n + (uint32_t)i3) * (uint32_t)C;
        const double q2 = __ldg(d2.q + prob * (uint32_t)d2.n + i2), q3 = __ldg(d3.q + prob * (uint32_t)d3.n + i3);
        long long jo = (long long)prob * sp.S_ext + (long long)(i0 - d0.ext_lo) * d0.stride +
                       (long long)(t_lo[1] - d1.ext_lo) * d1.stride + (long long)(i2 - d2.ext_lo) * d2.stride +
                       (long long)(i3 - d3.ext_lo) * d3.stride;
        long long io = (long long)prob * sp.S_own + (long long)(i0 - d0.own_lo) * tp.own_stride[0] +
                       (long long)(t_lo[1] - d1.own_lo) * tp.own_stride[1] + (long long)(i2 - d2.own_lo) * tp.own_stride[2] +
                       (long long)(i3 - d3.own_lo) * tp.own_stride[3];
#pragma unroll 1
        for (int j1 = 0; j1 < n1t; ++j1) {
            const int i1 = t_lo[1] + j1;
            // dimension 0 (control independent, lane specific): one coalesced table entry per state
            const double2 e0 = __ldg(tp.ft[0] + (prob * (uint32_t)d1.n + (uint32_t)i1) * (uint32_t)d0.n + (uint32_t)i0);
            const double t0 = e0.x;
            const uint32_t off02 = off2 + (uint32_t)__double2loint(e0.y);
            const double2 *lt1 = tp.lt[1] + (prob * (uint32_t)d1.n + (uint32_t)i1) * (uint32_t)C;
            // stage cost of the state: q terms summed in q_order
            const double q1 = __ldg(d1.q + prob * (uint32_t)d1.n + i1);
            double gs;
            if (tp.q_identity) {
                gs = ((q0 + q1) + q2) + q3;
            } else {
                auto qsel = [&](int o) { return o == 0 ? q0 : o == 1 ? q1 : o == 2 ? q2 : q3; };
                gs = ((qsel(qo0) + qsel(qo1)) + qsel(qo2)) + qsel(qo3);
            }

            double best = __longlong_as_double(0x7ff0000000000000LL);
            int arg = 0;
#pragma unroll UNROLL
            for (int c = 0; c < C; ++c) {
                const double2 e1 = __ldg(lt1 + c), e3 = __ldg(lt3 + c);
                const uint32_t o = off02 + (uint32_t)__double2loint(e1.y) + (uint32_t)__double2loint(e3.y);
                double v[16];
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    v[2 * k] = lds_f64(o + ro[k]);
                    v[2 * k + 1] = lds_f64(o + ro[k] + 8);
                }
#pragma unroll
                for (int m = 0; m < 8; ++m) v[m] = fma(t0, v[2 * m + 1] - v[2 * m], v[2 * m]);      // dimension 0 first
#pragma unroll
                for (int m = 0; m < 4; ++m) v[m] = fma(e1.x, v[2 * m + 1] - v[2 * m], v[2 * m]);
#pragma unroll
                for (int m = 0; m < 2; ++m) v[m] = fma(t2, v[2 * m + 1] - v[2 * m], v[2 * m]);
                const double val = fma(e3.x, v[1] - v[0], v[0]);
                const double tot = (gs + __ldg(rr + c)) + val;
                if (tot < best) { best = tot; arg = c; }
            }
            if (row_ok) {
                sp.J_out[jo] = best;
                idx_store(sp.idx_out, sp.idx_bytes, io, arg);
                if (sp.n_peers) { const int gi[4] = {i0, i1, i2, i3}; peer_store<4>(sp, (int)prob, gi, best); }
            }
            jo += d1.stride;
            io += tp.own_stride[1];
        }
    }
}

typedef CUresult (*PFN_encodeTiled)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                    const cuuint64_t *, const cuuint32_t *, const cuuint32_t *,
                                    CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                    CUtensorMapFloatOOBfill);
PFN_encodeTiled get_encode() {
    static PFN_encodeTiled fn = nullptr;
    if (fn) return fn;
    void *p = nullptr;
    cudaDriverEntryPointQueryResult qr;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qr) != cudaSuccess ||
        qr != cudaDriverEntryPointSuccess)
        return nullptr;
    fn = reinterpret_cast<PFN_encodeTiled>(p);
    return fn;
}

struct TileState {
    bool pa = false;                 // k_stage_tile_pa applies (table entries hold box byte offsets)
    int nt = 512;                    // threads per CTA of k_stage_tile_pa (512 x 64 registers: 32 warps per SM)
    TileParams tp{};
    std::vector<CUtensorMap> maps;   // one per J slot
    size_t smem = 0;
    void *d_lt[MAXD] = {nullptr, nullptr, nullptr, nullptr};
    void *d_ft[2] = {nullptr, nullptr};
    ~TileState() { for (void *p : d_lt) cudaFree(p); for (void *p : d_ft) cudaFree(p); }
};

}  // namespace

// Exact stencil bounds of dimension d: min / max over every state and control of
// cell(x'_d) - i_d.  Floating-point addition is monotone in each argument, so for a fixed own index
// the extreme queries are (Ta[i] + min Tb) + min Tc and (Ta[i] + max Tb) + max Tc, formed with the
// kernel's own association.  Returns false when x'_d does not depend on the state's own index.
bool stencil_reach(const HostProblem &hp, int d, int &lo, int &hi) {
    const int sa = hp.src_a[d], sb = hp.has_b[d] ? hp.src_b[d] : -1;
    if (sa != d && sb != d) return false;
    lo = std::numeric_limits<int>::max();
    hi = std::numeric_limits<int>::min();
    auto mm = [](const double *v, int cnt, double &mn, double &mx) {
        mn = std::numeric_limits<double>::infinity(); mx = -mn;
        for (int k = 0; k < cnt; ++k) { mn = std::min(mn, v[k]); mx = std::max(mx, v[k]); }
    };
    for (int p = 0; p < hp.P; ++p) {
        const double *ta = hp.Ta[d].data() + (size_t)p * hp.n[sa];
        const double *tb = hp.has_b[d] ? hp.Tb[d].data() + (size_t)p * hp.n[sb] : nullptr;
        double cmn = 0.0, cmx = 0.0, omn = 0.0, omx = 0.0;
        if (hp.has_c[d]) mm(hp.Tc[d].data() + (size_t)p * hp.C, hp.C, cmn, cmx);
        // the table NOT indexed by the own dimension contributes its extremes
        const bool a_own = sa == d, b_own = sb == d;
        if (tb && !b_own) mm(tb, hp.n[sb], omn, omx);
        if (!a_own) mm(ta, hp.n[sa], omn, omx);
        for (int i = 0; i < hp.n[d]; ++i) {
            double qlo, qhi;
            if (a_own && b_own) { qlo = qhi = ta[i] + tb[i]; }
            else if (a_own) { qlo = tb ? ta[i] + omn : ta[i]; qhi = tb ? ta[i] + omx : ta[i]; }
            else { qlo = omn + tb[i]; qhi = omx + tb[i]; }
            if (hp.has_c[d]) { qlo = qlo + cmn; qhi = qhi + cmx; }
            lo = std::min(lo, host_locate(hp, p, d, qlo) - i);
            hi = std::max(hi, host_locate(hp, p, d, qhi) - i);
        }
    }
    return true;
}

}  // namespace bellman

// host-only: the exact stencil bounds the tile kernel sizes its box with (include/bellman.h)
extern "C" int bellman_query_stencil(const bellman_desc *d, int32_t *lo_out, int32_t *hi_out) {
    bellman::HostProblem hp;
    if (!bellman::load_problem(d, hp).empty() || !lo_out || !hi_out) return BELLMAN_ERR_BAD_ARG;
    for (int k = 0; k < hp.D; ++k) {
        int lo = 0, hi = 0;
        if (bellman::stencil_reach(hp, k, lo, hi)) { lo_out[k] = lo; hi_out[k] = hi; }
        else { lo_out[k] = 1; hi_out[k] = -1; }          // lo > hi: x'_k does not depend on the own index
    }
    return BELLMAN_OK;
}

namespace bellman {

void tile_teardown(bellman_handle *h) {
    delete static_cast<TileState *>(h->tstate);
    h->tstate = nullptr;
}

void tile_setup(bellman_handle *h) {
    h->tstate = nullptr;
    const HostProblem &hp = h->hp;
    if (hp.D != 3 && hp.D != 4) return;
    if (std::getenv("BELLMAN_NO_TILE")) return;
    if (h->ld0 % 2) return;                       // TMA global strides: multiples of 16 bytes
    PFN_encodeTiled enc = get_encode();
    if (!enc) return;
    const int D = hp.D;
    int lo[MAXD] = {0, 0, 0, 0}, hi[MAXD] = {0, 0, 0, 0};
    for (int d = 0; d < D; ++d) {
        if (!stencil_reach(hp, d, lo[d], hi[d])) return;
        lo[d] = std::min(lo[d], 0);
        hi[d] = std::max(hi[d], 0);
        if (hi[d] - lo[d] > 24) return;           // not a narrow stencil: the gathers stay in L2 (direct kernel)
    }
    // tile extents: best states per staged element among boxes that leave three CTAs per SM
    int T[MAXD] = {TT0, 1, 1, 1};
    {
        int forced[3] = {0, 0, 0};
        if (const char *e = std::getenv("BELLMAN_TILE")) std::sscanf(e, "%d,%d,%d", &forced[0], &forced[1], &forced[2]);
        double best = -1.0;
        const size_t budget = (size_t)(std::getenv("BELLMAN_TILE_SMEM_KB") ? std::atoi(std::getenv("BELLMAN_TILE_SMEM_KB")) : 110) * 1024;
        for (int a : {1, 2, 4, 8, 16})
            for (int b : {1, 2, 4, 8, 16})
                for (int c : {1, 2, 4, 8, 16}) {
                    if (D == 3 && c != 1) continue;
                    if (forced[0] && (a != forced[0] || b != forced[1] || (D == 4 && c != forced[2]))) continue;
                    if (a * b * c < 8 || a * b * c > 256) continue;          // at least one state per warp
                    const int t[MAXD] = {TT0, a, b, c};
                    size_t el = 1;
                    for (int d = 0; d < D; ++d) {
                        int bx = t[d] + hi[d] - lo[d] + 1;
                        if (d == 0) bx = (bx + 1 + 1) / 2 * 2;               // even origin (+1), even extent
                        el *= (size_t)bx;
                    }
                    if (el * 8 > budget) continue;
                    // states per staged element, discounted when fewer than three CTAs fit an SM
                    const int ctas = (int)std::min<size_t>(3, (size_t)(227 * 1024) / (el * 8 + 1024));
                    const double score = (double)(TT0 * a * b * c) / (double)el * (ctas >= 3 ? 1.0 : ctas == 2 ? 0.8 : 0.5);
                    if (score > best) { best = score; T[1] = a; T[2] = b; T[3] = c; }
                }
        if (best < 0.0) return;
    }
    auto *ts = new TileState();
    TileParams &tp = ts->tp;
    int bs = 1, os = 1;
    tp.box_elems = 1;
    for (int d = 0; d < MAXD; ++d) {
        tp.T[d] = d < D ? T[d] : 1;
        tp.ntile[d] = d < D ? (h->own_n[d] + T[d] - 1) / T[d] : 1;
        tp.lo[d] = d < D ? lo[d] : 0;
        int bx = d < D ? T[d] + hi[d] - lo[d] + 1 : 1;
        if (d == 0) bx = (bx + 1 + 1) / 2 * 2;
        tp.box[d] = bx;
        tp.bstride[d] = bs;
        bs *= bx;
        tp.own_stride[d] = os;
        if (d < D) os *= h->own_n[d];
    }
    tp.box_elems = bs;
    ts->smem = (size_t)bs * 8;
    tp.pf_dist = std::getenv("BELLMAN_TILE_PF") ? std::atoi(std::getenv("BELLMAN_TILE_PF")) : 0;   // measured: +1 % on large grids, -45 % on L2-resident ones
    if ((long long)tp.ntile[0] * tp.ntile[1] > 2147483647LL || (long long)tp.ntile[2] * tp.ntile[3] > 65535 || hp.P > 65535) {
        delete ts;
        return;
    }
    for (int d = 0; d < D; ++d)
        if (tp.box[d] > 256) { delete ts; return; }

    // structure of Solver_pos_att's channel: dims 0 / 2 control independent with Tb from dims 1 / 3
    // (or absent), dims 1 / 3 control dependent on their own index only
    ts->pa = D == 4 && !std::getenv("BELLMAN_TILE_GENERIC");
    for (int d = 0; d < D && ts->pa; ++d) {
        if (hp.src_a[d] != d) ts->pa = false;
        if (d % 2 == 0 && (hp.has_c[d] || (hp.has_b[d] && hp.src_b[d] != d + 1))) ts->pa = false;
        if (d % 2 == 1 && (!hp.has_c[d] || (hp.has_b[d] && hp.src_b[d] != d))) ts->pa = false;
    }
    // locate tables of the (own index, control) dimensions, built with the normative operations
    // (this file is compiled with -ffp-contract=off on the host side: one rounding per operation)
    for (int d = 0; d < MAXD; ++d) tp.lt[d] = nullptr;
    for (int d = 0; d < D; ++d) {
        if (!hp.has_c[d] || hp.src_a[d] != d || (hp.has_b[d] && hp.src_b[d] != d)) continue;
        const int nd = hp.n[d];
        std::vector<double> tab((size_t)hp.P * nd * hp.C * 2);
        for (int p = 0; p < hp.P; ++p) {
            const double *sgrid = hp.grid[d].data() + (size_t)p * nd, *ri = hp.rinv[d].data() + (size_t)p * nd;
            const bool uni = hp.mode[(size_t)p * hp.D + d] == BELLMAN_LOCATE_UNIFORM;
            for (int i = 0; i < nd; ++i) {
                double b = hp.Ta[d][(size_t)p * nd + i];
                if (hp.has_b[d]) b = b + hp.Tb[d][(size_t)p * nd + i];
                for (int c = 0; c < hp.C; ++c) {
                    const double xq = b + hp.Tc[d][(size_t)p * hp.C + c];
                    const int cell = host_locate(hp, p, d, xq);
                    const double t = uni ? xq - (double)cell : (xq - sgrid[cell]) * ri[cell];
                    double cb;
                    // generic kernel: the cell; pos-att kernel: its byte offset inside the box
                    const long long bits = (long long)(unsigned int)(ts->pa ? cell * tp.bstride[d] * 8 : cell);
                    std::memcpy(&cb, &bits, 8);
                    tab[(((size_t)p * nd + i) * hp.C + c) * 2] = t;
                    tab[(((size_t)p * nd + i) * hp.C + c) * 2 + 1] = cb;
                }
            }
        }
        if (cudaMalloc(&ts->d_lt[d], tab.size() * 8) != cudaSuccess ||
            cudaMemcpy(ts->d_lt[d], tab.data(), tab.size() * 8, cudaMemcpyHostToDevice) != cudaSuccess) {
            delete ts;
            return;
        }
        tp.lt[d] = static_cast<const double2 *>(ts->d_lt[d]);
    }
    tp.ft[0] = tp.ft[1] = nullptr;
    tp.q_identity = (hp.q_order[0] == 0 && hp.q_order[1] == 1 && hp.q_order[2] == 2 && hp.q_order[3] == 3) ? 1 : 0;
    for (int k = 0; k < 2 && ts->pa; ++k) {
        const int d = 2 * k, e = d + 1, nd = hp.n[d], ne = hp.n[e];
        std::vector<double> tab((size_t)hp.P * ne * nd * 2);
        for (int p = 0; p < hp.P; ++p) {
            const double *sgrid = hp.grid[d].data() + (size_t)p * nd, *ri = hp.rinv[d].data() + (size_t)p * nd;
            const bool uni = hp.mode[(size_t)p * hp.D + d] == BELLMAN_LOCATE_UNIFORM;
            for (int ie = 0; ie < ne; ++ie)
                for (int i = 0; i < nd; ++i) {
                    double xq = hp.Ta[d][(size_t)p * nd + i];
                    if (hp.has_b[d]) xq = xq + hp.Tb[d][(size_t)p * ne + ie];
                    const int cell = host_locate(hp, p, d, xq);
                    const double t = uni ? xq - (double)cell : (xq - sgrid[cell]) * ri[cell];
                    double cb;
                    const long long bits = (long long)(unsigned int)(cell * tp.bstride[d] * 8);
                    std::memcpy(&cb, &bits, 8);
                    const size_t o = (((size_t)p * ne + ie) * nd + i) * 2;
                    tab[o] = t;
                    tab[o + 1] = cb;
                }
        }
        if (cudaMalloc(&ts->d_ft[k], tab.size() * 8) != cudaSuccess ||
            cudaMemcpy(ts->d_ft[k], tab.data(), tab.size() * 8, cudaMemcpyHostToDevice) != cudaSuccess) {
            delete ts;
            return;
        }
        tp.ft[k] = static_cast<const double2 *>(ts->d_ft[k]);
    }
    if (std::getenv("BELLMAN_TILE_DEBUG"))
        std::fprintf(stderr, "bellman tile%s: T = %d %d %d %d, stencil lo = %d %d %d %d hi = %d %d %d %d, box = %d %d %d %d (%zu KB), tables = %d%d%d%d\n",
                     ts->pa ? " (pos-att kernel)" : "", tp.T[0], tp.T[1], tp.T[2], tp.T[3], lo[0], lo[1], lo[2], lo[3], hi[0], hi[1], hi[2], hi[3], tp.box[0],
                     tp.box[1], tp.box[2], tp.box[3], ts->smem / 1024, tp.lt[0] != nullptr, tp.lt[1] != nullptr,
                     tp.lt[2] != nullptr, tp.lt[3] != nullptr);

    const int nslots = h->store_J_all ? hp.N : 2;
    ts->maps.resize(nslots);
    for (int s = 0; s < nslots; ++s) {
        cuuint64_t gdim[5], gstr[4];
        cuuint32_t box[5], estr[5] = {1, 1, 1, 1, 1};
        for (int d = 0; d < D; ++d) { gdim[d] = (cuuint64_t)h->ext_n[d]; box[d] = (cuuint32_t)tp.box[d]; }
        gdim[D] = (cuuint64_t)hp.P;
        box[D] = 1;
        for (int d = 1; d < D; ++d) gstr[d - 1] = (cuuint64_t)h->stride[d] * 8;
        gstr[D - 1] = (cuuint64_t)h->S_ext * 8;
        void *base = h->d_J + (size_t)s * h->slot_elems_J();
        CUresult r = enc(&ts->maps[s], CU_TENSOR_MAP_DATA_TYPE_FLOAT64, (cuuint32_t)(D + 1), base, gdim, gstr, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { delete ts; return; }
    }
    ts->nt = std::getenv("BELLMAN_TILE_NT") ? std::atoi(std::getenv("BELLMAN_TILE_NT")) : 512;
    const void *fn = ts->pa ? (ts->nt == 512 ? (const void *)k_stage_tile_pa<512, 1> : (const void *)k_stage_tile_pa<256, 3>) : D == 4 ? (const void *)k_stage_tile<4> : (const void *)k_stage_tile<3>;
    if (!raise_smem_limit(fn, ts->smem)) {
        delete ts;
        return;
    }
    h->tstate = ts;
}

bool tile_valid(const bellman_handle *h) { return h->tstate != nullptr; }

cudaError_t tile_launch_for_handle(bellman_handle *h, const StageParams &sp, int slot_next, cudaStream_t st) {
    auto *ts = static_cast<TileState *>(h->tstate);
    if (!ts) return cudaErrorNotSupported;
    const TileParams &tp = ts->tp;
    const dim3 grid((unsigned)(tp.ntile[0] * tp.ntile[1]), (unsigned)(tp.ntile[2] * tp.ntile[3]), (unsigned)sp.P);
    if (ts->pa && ts->nt == 512) k_stage_tile_pa<512, 1><<<grid, 512, ts->smem, st>>>(sp, tp, ts->maps[slot_next]);
    else if (ts->pa) k_stage_tile_pa<256, 3><<<grid, 256, ts->smem, st>>>(sp, tp, ts->maps[slot_next]);
    else if (h->hp.D == 4) k_stage_tile<4><<<grid, TNT, ts->smem, st>>>(sp, tp, ts->maps[slot_next]);
    else k_stage_tile<3><<<grid, TNT, ts->smem, st>>>(sp, tp, ts->maps[slot_next]);
    return cudaGetLastError();
}

}  // namespace bellman
