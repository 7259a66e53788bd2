// bellman_stream.cu — the D = 4 stage of Solver_pos_att's channel sweep (pos-att/Solver_pos_att.m:
// 244-297; `[F_gI.Values, U_Optimal_id] = min(J_current + F_gI(x', v', theta', w'), [], 5)` at :272)
// as a STREAMING, FACTORISED kernel.
//
// Structure of the channel (pos-att/Solver_pos_att.m:299-328), two coupled (position, rate) blocks:
//     x'_0 = Ta_0[i0] + Tb_0[i1]   (control independent)     x'_1 = Ta_1[i1] + Tc_1[c]
//     x'_2 = Ta_2[i2] + Tb_2[i3]   (control independent)     x'_3 = Ta_3[i3] + Tc_3[c]
// The normative 4-linear interpolation (include/bellman.h) reduces dimension 0 first, then 1, 2, 3.
// Its first two levels do not depend on (i2, i3), and the second depends on the control only through
// Tc_1[c], which takes few distinct values (5 for the 9 thruster combinations: net force 0, +-T, +-2T):
//     H[m1][m2][m3]  = lerp0( J[c0, m1, m2, m3], J[c0+1, m1, m2, m3];  t0(i0,i1) )
//     K[f][m2][m3]   = lerp1( H[c1(f)], H[c1(f)+1];  t1(i1, f) )              f = class of Tc_1[c]
// are shared by every state of the (i0, i1) column and every control of class f.  Per (state, control)
// only  lerp2 x 2  and  lerp3  remain.  Same operations on the same operands as k_stage_direct, so the
// results are bit-identical — the factorisation only removes recomputation: 27 + ~12 lerps per state
// instead of 135, 18 + ~15 shared-memory accesses instead of 144 (k_stage_tile_pa).
//
// A CTA owns a patch of T0 x T1 = 32 (i0, i1) columns (one per lane), T2 values of i2 (one per
// consumer warp) and WALKS i3.  Per step
//   * one TMA box (cp.async.bulk.tensor.5d) brings the J_{k+1} slab of ONE new dimension-3 node
//     (B0 x B1 x B2 doubles) into a small ring,
//   * warp j turns slab row m2 = j into K[f][j][node] for every class f (ring of W3 = stencil + 2 nodes
//     in shared memory, laid out [node][f][m2][lane]: conflict-free),
//   * consumer warp w finishes the state (i0, i1, i2 = w, i3): per control it reads the two K values of
//     the UPPER dimension-3 node (the lower node's pair is the previous step's upper pair, kept in
//     registers whenever the cells advance by exactly one — a host-built flag per step says so).
// Ring slot byte offsets of every (i3, control) come from a host table, so the loop has no modular
// arithmetic.  One __syncthreads per step.
#include <cuda.h>
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <type_traits>
#include <vector>

#include "bellman_handle.h"
#include "bellman_internal.h"

namespace bellman {

namespace {


constexpr int SMAXJ = 4;        // slab ring depth (TMA boxes in flight)

struct StreamParams {
    int T0, T0_log2, T1, T2, T3;    // tile: T0 * T1 = 32 columns, T2 consumer warps, T3 steps per CTA
    int ntile[MAXD];
    int lo[MAXD];                   // stencil lower bounds: cell(x'_d) - i_d >= lo[d]
    int B0, B1, B2;                 // slab extents (nodes)
    int NH1;                        // dimension-1 nodes a column needs (2 or 3)
    int NF1;                        // dimension-1 control classes
    int span3;                      // dimension-3 nodes a step needs: hi3 - lo3 + 2
    int W3;                         // K ring slots = span3 + 2
    int NJ;                         // slab ring slots
    int debug_sync;                 // BELLMAN_STREAM_DEBUG_SYNC=1: an extra __syncthreads() per iteration (racecheck runs:
                                    // the tool does not model mbarrier arrive / try_wait pairs between warps)
    int NP;                         // 0: every warp produces (plane j = warp) and warps < T2 also consume;
                                    // > 0: warp specialisation — warps < T2 only consume, the NP warps after them
                                    // share the B2 planes
    int slab_doubles;               // doubles per slab slot (128-byte aligned)
    int ring_doubles;               // W3 * NF1 * B2 * 32
    int own_stride[MAXD];           // strides of the owned index space (idx_out)
    uint32_t allmask;               // (1 << C) - 1
    double rc[4][12];               // control-cost term r[p][c] (constant bank: no registers, no loads)
    const double2 *ft0;             // [(p n1 + i1) n0 + i0] = {t0, cell0}
    const double2 *k1;              // [(p n1 + i1) NF1 + f]  = {t1, cell1 - (i1 + lo1)}
    const double2 *ft2;             // [(p n3 + i3) n2 + i2] = {t2, cell2 | chain flag << 32}
    // one record per (p, i3), blk3_bytes long: { double q3; uint32 flag3 (bit c: cell3(i3, c) ==
    // cell3(i3 - 1, c) + 1); uint32 pad; double t3[C]; uint32 up3[C] (ring BYTE offset of the UPPER node's
    // pair of control c; the lower node is one slot before) } — one cursor walks everything a step needs
    const unsigned char *blk3;
    int blk3_bytes;
};

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// Waits for the phase of `parity` to complete.  try_wait carries a suspend-time hint, so a waiting warp
// sleeps in hardware instead of competing for issue slots.  Bounded: a wait that outlives ~2^22 timed-out
// polls (seconds — a protocol bug or a lost TMA, never a legitimate state) traps, so the launch fails with
// an error instead of hanging the GPU.
__device__ __forceinline__ uint32_t mbar_try(uint32_t addr, uint32_t parity) {
    uint32_t done;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(done)
        : "r"(addr), "r"(parity), "r"(20000u)
        : "memory");
    return done;
}
__device__ __noinline__ void mbar_wait_slow(uint32_t addr, uint32_t parity) {
    for (uint32_t spins = 0; !mbar_try(addr, parity); ++spins)
        if (spins > (1u << 22)) asm volatile("trap;");
}
__device__ __forceinline__ void mbar_wait32(uint32_t addr, uint32_t parity) {
    if (!mbar_try(addr, parity)) mbar_wait_slow(addr, parity);
}
__device__ __forceinline__ void mbar_arrive32(uint32_t addr) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(addr) : "memory");
}
__device__ __forceinline__ void tma_load_5d(void *dst, const CUtensorMap *map, uint64_t *bar, int c0, int c1, int c2,
                                            int c3, int c4) {
    asm volatile(
        "cp.async.bulk.tensor.5d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];" ::
            "r"(smem_u32(dst)),
        "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
        : "memory");
}

// shared-memory accesses by 32-bit address.  volatile + "memory": the K ring is rewritten every step, so
// these must stay ordered with the barriers (and are never merged across iterations)
__device__ __forceinline__ double lds64(uint32_t addr) {
    double v;
    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr) : "memory");
    return v;
}
__device__ __forceinline__ void sts64(uint32_t addr, double v) {
    asm volatile("st.shared.f64 [%0], %1;" ::"r"(addr), "d"(v) : "memory");
}

// C controls, NF classes of dimension 1 (compile time: the per-control and per-class values live in registers)
template <int C, int NF, bool IDX32>
__global__ void __launch_bounds__(384, 1)
k_stage_stream(const __grid_constant__ StageParams sp, const __grid_constant__ StreamParams tp,
               const __grid_constant__ CUtensorMap tmap) {
    extern __shared__ __align__(128) double smem[];
    // slab[s]: TMA completion of slab ring slot s.  full[it & 3]: every producer warp arrives after writing
    // node `it` into the K ring; empty[it & 3]: every consumer warp arrives after finishing the step of
    // iteration `it`.  The ring has one slot more than a step needs, so a warp only ever waits for what the
    // OTHER warps did an iteration ago:
    //   consume(it) waits full of it - 1  (reads nodes it - span3 .. it - 1)
    //   produce(it) waits empty of it - 2 (overwrites node it - W3 = it - span3 - 2, last read by the step of it - 2)
    // Four barriers of each kind although warps are at most two iterations apart: with two, a producer-only
    // warp that is slow to poll empty(it - 2) could find the barrier already completed AGAIN for iteration
    // `it` (the consumers only need its node it - 1 to get that far) and would wait for a parity that never
    // comes; the completion after next on the same barrier (it + 2) needs that warp's own node it + 1.
    __shared__ __align__(8) uint64_t mbar_all[SMAXJ + 8];
    uint64_t *const mbar_slab = mbar_all, *const mbar_full = mbar_all + SMAXJ, *const mbar_empty = mbar_all + SMAXJ + 4;

    const int tid = threadIdx.x, lane = tid & 31, wrp = tid >> 5, NW = blockDim.x >> 5;
    const uint32_t prob = blockIdx.z;
    const DimParams &d0 = sp.dim[0], &d1 = sp.dim[1], &d2 = sp.dim[2], &d3 = sp.dim[3];
    const int tx = blockIdx.x % tp.ntile[0], ty = blockIdx.x / tp.ntile[0];
    const int tz = blockIdx.y % tp.ntile[2], tw = blockIdx.y / tp.ntile[2];
    const int a0 = d0.own_lo + tx * tp.T0, h0 = min(a0 + tp.T0, d0.own_lo + d0.own_n);
    const int a1 = d1.own_lo + ty * tp.T1, h1 = min(a1 + tp.T1, d1.own_lo + d1.own_n);
    const int a2 = d2.own_lo + tz * tp.T2, h2 = min(a2 + tp.T2, d2.own_lo + d2.own_n);
    const int a3 = d3.own_lo + tw * tp.T3, h3 = min(a3 + tp.T3, d3.own_lo + d3.own_n);
    int org0 = a0 + tp.lo[0];
    org0 -= (org0 - d0.ext_lo) & 1;                      // TMA: 16-byte aligned innermost coordinate
    const int org1 = a1 + tp.lo[1], org2 = a2 + tp.lo[2], m3_first = a3 + tp.lo[3];
    const int T3n = h3 - a3;
    const int span3 = tp.span3;
    const int n_prod = T3n + span3 - 1;                  // dimension-3 nodes this CTA turns into K
    const int n_iter = T3n + span3;                      // node `it` is produced in iteration it, step it - span3 consumed
    const int NJ = tp.NJ;
    const int n_cons = min(tp.T2, h2 - a2);              // consumer warps with a row of states

    uint32_t ring32 = smem_u32(smem);
    asm volatile("" : "+r"(ring32));             // opaque: kept in a register, not rebuilt from special registers
    const uint32_t slabs32 = ring32 + 8u * (uint32_t)tp.ring_doubles;
    const uint32_t slab_stride = 8u * (uint32_t)tp.slab_doubles;
    const uint32_t slab_bytes = (uint32_t)(tp.B0 * tp.B1 * tp.B2) * 8u;
    auto issue = [&](int node, int slot) {               // one thread only; slot = node % NJ
        mbar_expect_tx(&mbar_slab[slot], slab_bytes);
        tma_load_5d(smem + tp.ring_doubles + (size_t)slot * tp.slab_doubles, &tmap, &mbar_slab[slot], org0 - d0.ext_lo,
                    org1 - d1.ext_lo, org2 - d2.ext_lo, m3_first + node - d3.ext_lo, (int)prob);
    };
    if (tid == 0) {
        for (int s = 0; s < NJ; ++s) mbar_init(&mbar_slab[s], 1);
        for (int b = 0; b < 4; ++b) {
            mbar_init(&mbar_full[b], (uint32_t)(tp.NP > 0 ? tp.NP : tp.B2));
            mbar_init(&mbar_empty[b], (uint32_t)n_cons);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        for (int s = 0; s < NJ && s < n_prod; ++s) issue(s, s);
    }

    // ---- this lane's (i0, i1) column: everything the producer role needs, for the whole walk ----
    const int l0 = lane & (tp.T0 - 1), l1 = lane >> tp.T0_log2;
    const bool ok01 = (a0 + l0 < h0) && (a1 + l1 < h1);
    const int i0 = min(a0 + l0, h0 - 1), i1 = min(a1 + l1, h1 - 1);
    const double2 e0 = __ldg(tp.ft0 + ((size_t)prob * d1.n + i1) * d0.n + i0);
    const double t0 = e0.x;
    double t1[NF];
    uint32_t selbits = 0;
#pragma unroll
    for (int f = 0; f < NF; ++f) {
        const double2 e = __ldg(tp.k1 + ((size_t)prob * d1.n + i1) * NF + f);
        t1[f] = e.x;
        selbits |= (uint32_t)__double2loint(e.y) << f;
    }
    const uint32_t row_b = 8u * (uint32_t)tp.B0;                  // bytes between slab rows m1
    const uint32_t node2_b = row_b * (uint32_t)tp.B1;             // bytes between slab planes m2
    const uint32_t cls_b = 256u * (uint32_t)tp.B2;                // bytes between classes inside a ring slot
    const uint32_t slot_b = cls_b * (uint32_t)NF;                 // bytes between ring slots
    const uint32_t ring_b = slot_b * (uint32_t)tp.W3;
    const bool nh3 = tp.NH1 == 3;
    // producer: lower-left corner of this column inside slab plane m2 = wrp, and where its K values go
    // producer warps: all of them (plane j = wrp, stride NW), or — specialised — warps T2 .. T2 + NP - 1
    const bool prod = tp.NP > 0 ? wrp >= tp.T2 : wrp < tp.B2;
    const int pj0 = tp.NP > 0 ? wrp - tp.T2 : wrp, pjs = tp.NP > 0 ? tp.NP : NW;
    const uint32_t pA0 = slabs32 + 8u * (uint32_t)((__double2loint(e0.y) - org0) + tp.B0 * (i1 - a1)) + node2_b * (uint32_t)pj0;
    const uint32_t kout0 = ring32 + 8u * (uint32_t)(pj0 * 32 + lane);

    // ---- consumer role: warp w finishes the states (i0, i1, a2 + w, i3) ----
    const int i2 = a2 + wrp;
    const bool cons = wrp < n_cons;
    const int i2c = min(i2, h2 - 1);
    const double *rc = tp.rc[prob];              // constant bank, warp-uniform
    // stage cost: the q terms are summed in q_order; dimension 3's term changes every step, the partial
    // sums that do not involve it are formed once (same association as the normative left-to-right sum)
    int qpos = 0;
    double qa = 0.0, qb = 0.0, qc = 0.0;
    {
        const double qv0 = __ldg(d0.q + (size_t)prob * d0.n + i0), qv1 = __ldg(d1.q + (size_t)prob * d1.n + i1),
                     qv2 = __ldg(d2.q + (size_t)prob * d2.n + i2c);
        int k = 0;
#pragma unroll
        for (int m = 0; m < 4; ++m) {
            const int o = sp.q_order[m];
            if (o == 3) { qpos = m; continue; }
            const double v = o == 0 ? qv0 : o == 1 ? qv1 : qv2;
            if (k == 0) qa = v; else if (k == 1) qb = v; else qc = v;
            ++k;
        }
        if (qpos >= 2) qa = qa + qb;                     // (qa + qb) precedes dimension 3's term
        if (qpos == 3) qa = qa + qc;
    }
    const double2 *ft2p = tp.ft2 + (size_t)prob * d3.n * d2.n + i2c;
    long long jo = (long long)prob * sp.S_ext + (long long)(i0 - d0.ext_lo) * d0.stride +
                   (long long)(i1 - d1.ext_lo) * d1.stride + (long long)(i2c - d2.ext_lo) * d2.stride +
                   (long long)(a3 - d3.ext_lo) * d3.stride;
    long long io = (long long)prob * sp.S_own + (long long)(i0 - d0.own_lo) * tp.own_stride[0] +
                   (long long)(i1 - d1.own_lo) * tp.own_stride[1] + (long long)(i2c - d2.own_lo) * tp.own_stride[2] +
                   (long long)(a3 - d3.own_lo) * tp.own_stride[3];
    // tables of the NEXT step, fetched right after a step's stores: they travel while the warp runs its
    // producer role, so no global-memory latency is left on the step itself
    double2 e2n = make_double2(0.0, 0.0);
    double t3n[C], q3n = 0.0;
    uint32_t up3n[C], f3n = 0;
    // two cursors advance by one step (one i3) per prefetch: no index arithmetic in the loop
    const double2 *c_ft2 = ft2p + (size_t)a3 * d2.n;
    const unsigned char *c_blk = tp.blk3 + ((size_t)prob * d3.n + a3) * (size_t)tp.blk3_bytes;
    const int ft2_step = d2.n, blk_step = tp.blk3_bytes;
    auto prefetch = [&]() {
        e2n = __ldg(c_ft2);
        const double2 hdr = __ldg(reinterpret_cast<const double2 *>(c_blk));
        q3n = hdr.x;
        f3n = (uint32_t)__double2loint(hdr.y);
        const double *t3p = reinterpret_cast<const double *>(c_blk + 16);
        const uint32_t *u3p = reinterpret_cast<const uint32_t *>(c_blk + 16 + 8 * C);
#pragma unroll
        for (int c = 0; c < C; ++c) { t3n[c] = __ldg(t3p + c); up3n[c] = __ldg(u3p + c); }
        c_ft2 += ft2_step;
        c_blk += blk_step;
    };
#pragma unroll
    for (int c = 0; c < C; ++c) { t3n[c] = 0.0; up3n[c] = 0; }
    if (cons) prefetch();

    __syncthreads();          // mbarriers initialised

    uint32_t bar_slab = smem_u32(&mbar_slab[0]);
    asm volatile("" : "+r"(bar_slab));
    const uint32_t bar_full = bar_slab + 8u * SMAXJ, bar_empty = bar_full + 32u;   // mbar_slab / mbar_full / mbar_empty are one array
    uint32_t slotK_b = 0;                     // byte offset of ring slot it % W3
    uint32_t slab_b = 0, slab_i = 0, slab_ph = 0;   // slab ring slot of node `it`: byte offset, index, phase parity
    const uint32_t slab_ring_b = slab_stride * (uint32_t)NJ;
    const int emp_from = span3 + 2;           // first iteration whose producer has a step (it - 2) to wait for

    // ---- produce: slab plane m2 = j of node `it`  ->  K[f][j] of ring slot it % W3 ----
    auto produce = [&](int it) {
        if (prod) {
            if (it >= emp_from) mbar_wait32(bar_empty + 8u * (uint32_t)((it - 2) & 3), (uint32_t)(((it - emp_from) >> 2) & 1));
            mbar_wait32(bar_slab + 8u * slab_i, slab_ph);
            uint32_t pA = pA0 + slab_b, ko = kout0 + slotK_b;
            for (int j = pj0; j < tp.B2; j += pjs) {
                const double lo0 = lds64(pA), hi0 = lds64(pA + 8), lo1 = lds64(pA + row_b), hi1 = lds64(pA + row_b + 8);
                double lo2 = 0.0, hi2 = 0.0;
                if (nh3) { lo2 = lds64(pA + 2 * row_b); hi2 = lds64(pA + 2 * row_b + 8); }
                const double H0 = fma(t0, hi0 - lo0, lo0);
                const double H1 = fma(t0, hi1 - lo1, lo1);
                const double H2 = fma(t0, hi2 - lo2, lo2);
                const double dA = H1 - H0, dB = H2 - H1;
#pragma unroll
                for (int f = 0; f < NF; ++f) {
                    const bool sl = (selbits >> f) & 1u;
                    sts64(ko + (uint32_t)f * cls_b, fma(t1[f], sl ? dB : dA, sl ? H1 : H0));
                }
                pA += node2_b * (uint32_t)pjs;
                ko += 256u * (uint32_t)pjs;
            }
            __syncwarp();
            if (lane == 0) mbar_arrive32(bar_full + 8u * (uint32_t)(it & 3));
        }
        slab_b += slab_stride;
        ++slab_i;
        if (slab_b == slab_ring_b) { slab_b = 0; slab_i = 0; slab_ph ^= 1u; }
    };
    // every producer is done with node it - 1: its K values are visible, its slab slot is free again
    // The slab slot is refilled by the LAST warp (a producer-only warp: it has the time), so no consumer
    // carries the TMA issue on its critical path.
    const bool tma_warp = wrp == NW - 1;
    int iss_slot = 0;                          // slab slot of node it - 1 (the one sync_prev(it) refills)
    auto sync_prev = [&](int it) {
        if (cons || tma_warp) {
            mbar_wait32(bar_full + 8u * (uint32_t)((it - 1) & 3), (uint32_t)(((it - 1) >> 2) & 1));
            if (tma_warp && lane == 0 && it - 1 + NJ < n_prod) issue(it - 1 + NJ, iss_slot);
        }
        if (++iss_slot == NJ) iss_slot = 0;
    };
    // ---- consume: step i3 = a3 + it - span3.  (alo, dlo) hold the pairs the previous step read as its
    // UPPER node, (aup, dup) receive this step's — the caller alternates two register sets, nothing is copied
    double *Jp = sp.J_out + jo;
    long long Io = io;
    auto consume = [&](int it, double (&alo)[C], double (&dlo)[C], double (&aup)[C], double (&dup)[C]) {
        if (cons) {
            const double t2 = e2n.x;
            const bool fast = it > span3 && __double2hiint(e2n.y) != 0 && f3n == tp.allmask;
            const uint32_t kb = ring32 + 8u * (uint32_t)((__double2loint(e2n.y) - org2) * 32 + lane);
            double gs = qa + q3n;                         // q_order: see qpos above
            if (qpos <= 2) gs = gs + (qpos == 2 ? qc : qb);
            if (qpos <= 1) gs = gs + qc;
            double bu[C];
#pragma unroll
            for (int c = 0; c < C; ++c) {                 // the upper node's pair of every control
                const uint32_t p = kb + up3n[c];
                aup[c] = lds64(p);
                bu[c] = lds64(p + 256);
            }
            if (!fast) {                                  // first step, clamped edge: the lower node's pairs too
#pragma unroll
                for (int c = 0; c < C; ++c) {
                    // the lower node sits one ring slot before the upper one (wrapping)
                    const uint32_t p = kb + (up3n[c] >= slot_b ? up3n[c] - slot_b : up3n[c] + (ring_b - slot_b));
                    alo[c] = lds64(p);
                    dlo[c] = lds64(p + 256) - alo[c];
                }
            }
            // two serial first-index-wins chains (controls [0, CH) and [CH, C)), merged at the end: half the
            // dependent compare latency of one chain; the higher-indexed half only wins when strictly smaller
            constexpr int CH = (C + 1) / 2;
            double best = __longlong_as_double(0x7ff0000000000000LL), best2 = best;
            int arg = 0, arg2 = CH;
#pragma unroll
            for (int c = 0; c < C; ++c) {
                dup[c] = bu[c] - aup[c];
                const double vlo = fma(t2, dlo[c], alo[c]), vup = fma(t2, dup[c], aup[c]);   // dimension 2
                const double val = fma(t3n[c], vup - vlo, vlo);                              // dimension 3
                const double tot = (gs + rc[c]) + val;
                if (c < CH) { if (tot < best) { best = tot; arg = c; } }
                else { if (tot < best2) { best2 = tot; arg2 = c; } }
            }
            if (best2 < best) { best = best2; arg = arg2; }
            if (ok01) {
                *Jp = best;
                if (IDX32) sp.idx_out[Io] = arg;          // the default int32 store stays branch-free in the step loop
                else idx_store(sp.idx_out, sp.idx_bytes, Io, arg);
                if (sp.n_peers) { const int gi[4] = {i0, i1, i2, a3 + it - span3}; peer_store<4>(sp, (int)prob, gi, best); }
            }
            Jp += d3.stride;
            Io += tp.own_stride[3];
            __syncwarp();
            if (lane == 0) mbar_arrive32(bar_empty + 8u * (uint32_t)(it & 3));
            if (it + 1 < n_iter) prefetch();
        }
    };
    auto advance = [&]() {
        slotK_b += slot_b;
        if (slotK_b == ring_b) slotK_b = 0;
    };
    double A0[C], D0[C], A1[C], D1[C];
#pragma unroll
    for (int c = 0; c < C; ++c) { A0[c] = 0.0; D0[c] = 0.0; A1[c] = 0.0; D1[c] = 0.0; }
    // prologue: the first span3 nodes (no step can run yet)
    int it = 0;
#pragma unroll 1
    for (; it < span3; ++it) {
        produce(it);
        if (it >= 1) sync_prev(it);
        advance();
        if (tp.debug_sync) __syncthreads();
    }
    // steady state, two iterations per trip so that the register sets alternate without copies
#pragma unroll 1
    for (; it + 1 < n_iter; it += 2) {
        produce(it);                          // it < n_prod always holds here (it <= n_iter - 2)
        sync_prev(it);
        consume(it, A0, D0, A1, D1);
        advance();
        if (tp.debug_sync) __syncthreads();
        if (it + 1 < n_prod) produce(it + 1);
        sync_prev(it + 1);
        consume(it + 1, A1, D1, A0, D0);
        advance();
        if (tp.debug_sync) __syncthreads();
    }
    if (it < n_iter) {                        // odd count: the last step (it == n_prod: nothing left to produce)
        if (it < n_prod) produce(it);
        sync_prev(it);
        consume(it, A0, D0, A1, D1);
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

struct StreamState {
    StreamParams tp{};
    std::vector<CUtensorMap> maps;   // one per J slot
    size_t smem = 0;
    int nthreads = 0, C = 0;
    void *d_tab[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    ~StreamState() { for (void *p : d_tab) cudaFree(p); }
};

double pack_bits(uint32_t lo, uint32_t hi) {
    const uint64_t b = (uint64_t)lo | ((uint64_t)hi << 32);
    double d;
    std::memcpy(&d, &b, 8);
    return d;
}

template <int C, int NF>
bool stream_attr(size_t smem) {
    return raise_smem_limit((const void *)k_stage_stream<C, NF, true>, smem) &&
           raise_smem_limit((const void *)k_stage_stream<C, NF, false>, smem);
}

}  // namespace

void stream_teardown(bellman_handle *h) {
    delete static_cast<StreamState *>(h->sstate);
    h->sstate = nullptr;
}

bool stream_valid(const bellman_handle *h) { return h->sstate != nullptr; }

void stream_setup(bellman_handle *h) {
    h->sstate = nullptr;
    const HostProblem &hp = h->hp;
    if (hp.D != 4 || std::getenv("BELLMAN_NO_STREAM") || std::getenv("BELLMAN_NO_TILE")) return;
    if (hp.C != 6 && hp.C != 9) return;               // instantiated control counts (9 combinations, 6 in failure mode)
    if (hp.P > 4) return;                             // r[p][c] travels in the kernel parameters
    if (h->ld0 % 2) return;                           // TMA global strides: multiples of 16 bytes
    // structure of Solver_pos_att's channel (same test as k_stage_tile_pa)
    for (int d = 0; d < 4; ++d) {
        if (hp.src_a[d] != d) return;
        if (d % 2 == 0 && (hp.has_c[d] || (hp.has_b[d] && hp.src_b[d] != d + 1))) return;
        if (d % 2 == 1 && (!hp.has_c[d] || (hp.has_b[d] && hp.src_b[d] != d))) return;
    }
    PFN_encodeTiled enc = get_encode();
    if (!enc) return;
    int lo[4], hi[4];
    for (int d = 0; d < 4; ++d) {
        if (!stencil_reach(hp, d, lo[d], hi[d])) return;
        lo[d] = std::min(lo[d], 0);
        hi[d] = std::max(hi[d], 0);
    }
    if (hi[0] - lo[0] > 6 || hi[1] - lo[1] > 1 || hi[2] - lo[2] > 4 || hi[3] - lo[3] > 40) return;
    const int C = hp.C, P = hp.P;
    // classes of dimension 1: controls whose Tc_1 entries are bit-identical in every problem
    std::vector<int> cls(C, -1);
    int NF1 = 0;
    for (int c = 0; c < C; ++c) {
        for (int e = 0; e < c && cls[c] < 0; ++e) {
            bool same = true;
            for (int p = 0; p < P && same; ++p)
                same = std::memcmp(&hp.Tc[1][(size_t)p * C + c], &hp.Tc[1][(size_t)p * C + e], 8) == 0;
            if (same) cls[c] = cls[e];
        }
        if (cls[c] < 0) cls[c] = NF1++;
    }
    if (!((C == 9 && NF1 == 5) || (C == 6 && NF1 == 4))) return;   // the instantiated (controls, classes) pairs

    auto *ss = new StreamState();
    StreamParams &tp = ss->tp;
    ss->C = C;
    int fT1 = 0, fT2 = 0, fT3 = 0, fNJ = 0, fNP = -1;
    if (const char *e = std::getenv("BELLMAN_STREAM")) std::sscanf(e, "%d,%d,%d,%d,%d", &fT1, &fT2, &fT3, &fNJ, &fNP);
    tp.T1 = (fT1 == 1 || fT1 == 2 || fT1 == 4) ? fT1 : 2;
    tp.T0 = 32 / tp.T1;
    tp.T0_log2 = tp.T1 == 1 ? 5 : tp.T1 == 2 ? 4 : 3;
    tp.NJ = (fNJ >= 2 && fNJ <= SMAXJ) ? fNJ : 3;
    tp.debug_sync = std::getenv("BELLMAN_STREAM_DEBUG_SYNC") ? 1 : 0;
    for (int d = 0; d < 4; ++d) tp.lo[d] = lo[d];
    tp.B0 = (tp.T0 + hi[0] - lo[0] + 1 + 1 + 1) / 2 * 2;          // +1: even origin; even extent
    tp.B1 = tp.T1 + hi[1] - lo[1] + 1;
    tp.NH1 = hi[1] - lo[1] + 2;
    tp.NF1 = NF1;
    tp.span3 = hi[3] - lo[3] + 2;
    tp.W3 = tp.span3 + 2;                                        // one slot of slack: warps run up to an iteration apart
    const int w2 = hi[2] - lo[2] + 1;
    auto smem_of = [&](int T2) {
        const int B2 = T2 + w2;
        const size_t slab = ((size_t)tp.B0 * tp.B1 * B2 * 8 + 127) / 128 * 128;
        return (size_t)tp.W3 * NF1 * B2 * 256 + (size_t)tp.NJ * slab;
    };
    // Consumer rows per CTA: as many as the shared memory of one SM holds (one CTA per SM), with one
    // dedicated producer warp per two consumer warps (measured on B200, pos-att x4: 8 + 4 warps 6.8 ms;
    // 4 + 2 warps x 2 CTAs 7.9 ms; unspecialised 10 + 2 warps 7.5 ms, 4 + 2 warps x 2 CTAs 8.5 ms)
    int T2 = 0;
    if (fT2 > 0) {
        T2 = fT2;
    } else {
        for (int t : {8, 7, 6, 5, 4, 3, 2})
            if (t + w2 <= 12 && smem_of(t) + 1024 <= 225 * 1024) { T2 = t; break; }
    }
    if (T2 < 1 || T2 + w2 > 12 || smem_of(T2) + 1024 > 225 * 1024) { delete ss; return; }
    tp.T2 = T2;
    tp.B2 = T2 + w2;
    tp.slab_doubles = (int)((((size_t)tp.B0 * tp.B1 * tp.B2 * 8 + 127) / 128 * 128) / 8);
    tp.ring_doubles = tp.W3 * NF1 * tp.B2 * 32;
    ss->smem = smem_of(T2);
    tp.NP = fNP >= 0 ? std::min(fNP, tp.B2) : (T2 + 1) / 2;
    if (tp.NP > 0 && T2 + tp.NP > 12) tp.NP = 12 - T2;
    ss->nthreads = tp.NP > 0 ? 32 * (T2 + tp.NP) : 32 * tp.B2;
    // steps per CTA: the whole owned range unless that leaves the GPU short of CTAs; chunks are
    // multiples of W3 so that a node's ring slot does not depend on the chunk (host table lt3)
    tp.ntile[0] = (h->own_n[0] + tp.T0 - 1) / tp.T0;
    tp.ntile[1] = (h->own_n[1] + tp.T1 - 1) / tp.T1;
    tp.ntile[2] = (h->own_n[2] + tp.T2 - 1) / tp.T2;
    {
        const long long base_ctas = (long long)tp.ntile[0] * tp.ntile[1] * tp.ntile[2] * P;
        int T3 = h->own_n[3];
        if (fT3 > 0) T3 = std::max(tp.W3, fT3 / tp.W3 * tp.W3);
        else
            while (base_ctas * ((h->own_n[3] + T3 - 1) / T3) < 2LL * 148 * 2 && T3 > 4 * tp.W3)
                T3 = std::max(4 * tp.W3, (T3 / 2 + tp.W3 - 1) / tp.W3 * tp.W3);
        tp.T3 = std::min(T3, h->own_n[3]);
        if (tp.T3 < h->own_n[3] && tp.T3 % tp.W3) tp.T3 = h->own_n[3];
    }
    tp.ntile[3] = (h->own_n[3] + tp.T3 - 1) / tp.T3;
    if ((long long)tp.ntile[0] * tp.ntile[1] > 2147483647LL || (long long)tp.ntile[2] * tp.ntile[3] > 65535 || P > 65535) {
        delete ss;
        return;
    }
    {
        int os = 1;
        for (int d = 0; d < 4; ++d) { tp.own_stride[d] = os; os *= h->own_n[d]; }
    }
    tp.allmask = (1u << C) - 1u;
    for (int p = 0; p < P; ++p)
        for (int c = 0; c < C; ++c) tp.rc[p][c] = hp.r[(size_t)p * C + c];

    // ---- host tables, built with the normative operations (one rounding per operation) ----
    auto locate_t = [&](int p, int d, double xq, int &cell) {
        const int nd = hp.n[d];
        const double *sgrid = hp.grid[d].data() + (size_t)p * nd, *ri = hp.rinv[d].data() + (size_t)p * nd;
        cell = host_locate(hp, p, d, xq);
        return hp.mode[(size_t)p * hp.D + d] == BELLMAN_LOCATE_UNIFORM ? xq - (double)cell : (xq - sgrid[cell]) * ri[cell];
    };
    const int n0 = hp.n[0], n1 = hp.n[1], n2 = hp.n[2], n3 = hp.n[3];
    std::vector<double> ft0((size_t)P * n1 * n0 * 2), k1((size_t)P * n1 * NF1 * 2), ft2((size_t)P * n3 * n2 * 2),
        lt3((size_t)P * n3 * C);
    std::vector<uint32_t> up3((size_t)P * n3 * C, 0u);
    std::vector<uint32_t> flag3((size_t)P * n3, 0u);
    tp.blk3_bytes = (16 + 12 * C + 15) / 16 * 16;
    std::vector<int> rep(NF1, 0);                     // a representative control of every class
    for (int c = C - 1; c >= 0; --c) rep[cls[c]] = c;
    const int own_lo3 = h->own_lo[3];
    for (int p = 0; p < P; ++p) {
        for (int i1 = 0; i1 < n1; ++i1) {
            for (int i0 = 0; i0 < n0; ++i0) {
                double xq = hp.Ta[0][(size_t)p * n0 + i0];
                if (hp.has_b[0]) xq = xq + hp.Tb[0][(size_t)p * n1 + i1];
                int cell;
                const double t = locate_t(p, 0, xq, cell);
                const size_t o = (((size_t)p * n1 + i1) * n0 + i0) * 2;
                ft0[o] = t;
                ft0[o + 1] = pack_bits((uint32_t)cell, 0);
            }
            double b = hp.Ta[1][(size_t)p * n1 + i1];
            if (hp.has_b[1]) b = b + hp.Tb[1][(size_t)p * n1 + i1];
            for (int f = 0; f < NF1; ++f) {
                int cell;
                const double t = locate_t(p, 1, b + hp.Tc[1][(size_t)p * C + rep[f]], cell);
                const int sel = cell - (i1 + lo[1]);
                if (sel < 0 || sel > tp.NH1 - 2) { delete ss; return; }      // cannot happen: lo/hi are exact bounds
                const size_t o = (((size_t)p * n1 + i1) * NF1 + f) * 2;
                k1[o] = t;
                k1[o + 1] = pack_bits((uint32_t)sel, 0);
            }
        }
        std::vector<int> prev2(n2, 0), prev3(C, 0);
        for (int i3 = 0; i3 < n3; ++i3) {
            for (int i2 = 0; i2 < n2; ++i2) {
                double xq = hp.Ta[2][(size_t)p * n2 + i2];
                if (hp.has_b[2]) xq = xq + hp.Tb[2][(size_t)p * n3 + i3];
                int cell;
                const double t = locate_t(p, 2, xq, cell);
                const size_t o = (((size_t)p * n3 + i3) * n2 + i2) * 2;
                ft2[o] = t;
                ft2[o + 1] = pack_bits((uint32_t)cell, (i3 > 0 && cell == prev2[i2]) ? 1u : 0u);
                prev2[i2] = cell;
            }
            double b = hp.Ta[3][(size_t)p * n3 + i3];
            if (hp.has_b[3]) b = b + hp.Tb[3][(size_t)p * n3 + i3];
            uint32_t fl = 0;
            for (int c = 0; c < C; ++c) {
                int cell;
                const double t = locate_t(p, 3, b + hp.Tc[3][(size_t)p * C + c], cell);
                // ring slot of a node: (node - first node of the rank's walk) mod W3; chunk starts are
                // multiples of W3 steps, so this is the same in every chunk
                auto slot_off = [&](int node) {
                    int s = (node - (own_lo3 + lo[3])) % tp.W3;
                    if (s < 0) s += tp.W3;
                    return (uint32_t)((s * NF1 + cls[c]) * tp.B2 * 256);   // byte offset inside the K ring
                };
                const size_t o = ((size_t)p * n3 + i3) * C + c;
                lt3[o] = t;
                up3[o] = slot_off(cell + 1);
                if (i3 > 0 && cell == prev3[c] + 1) fl |= 1u << c;
                prev3[c] = cell;
            }
            flag3[(size_t)p * n3 + i3] = fl;
        }
    }
    auto upload = [&](const void *src, size_t bytes, void **dst) {
        return cudaMalloc(dst, bytes) == cudaSuccess && cudaMemcpy(*dst, src, bytes, cudaMemcpyHostToDevice) == cudaSuccess;
    };
    std::vector<unsigned char> blk((size_t)P * n3 * tp.blk3_bytes, 0);
    for (size_t r = 0; r < (size_t)P * n3; ++r) {
        unsigned char *b = blk.data() + r * tp.blk3_bytes;
        std::memcpy(b, &hp.q[3][r], 8);
        std::memcpy(b + 8, &flag3[r], 4);
        std::memcpy(b + 16, &lt3[r * C], 8 * (size_t)C);
        std::memcpy(b + 16 + 8 * C, &up3[r * C], 4 * (size_t)C);
    }
    if (!upload(ft0.data(), ft0.size() * 8, &ss->d_tab[0]) || !upload(k1.data(), k1.size() * 8, &ss->d_tab[1]) ||
        !upload(ft2.data(), ft2.size() * 8, &ss->d_tab[2]) || !upload(blk.data(), blk.size(), &ss->d_tab[3])) {
        cudaGetLastError();
        delete ss;
        return;
    }
    tp.ft0 = static_cast<const double2 *>(ss->d_tab[0]);
    tp.k1 = static_cast<const double2 *>(ss->d_tab[1]);
    tp.ft2 = static_cast<const double2 *>(ss->d_tab[2]);
    tp.blk3 = static_cast<const unsigned char *>(ss->d_tab[3]);

    // one tensor map per J slot: [P][n3][n2][n1][ld0] fp64, box = B0 x B1 x B2 x 1 x 1
    const int nslots = h->store_J_all ? hp.N : 2;
    ss->maps.resize(nslots);
    for (int s = 0; s < nslots; ++s) {
        cuuint64_t gdim[5], gstr[4];
        cuuint32_t box[5] = {(cuuint32_t)tp.B0, (cuuint32_t)tp.B1, (cuuint32_t)tp.B2, 1, 1}, estr[5] = {1, 1, 1, 1, 1};
        for (int d = 0; d < 4; ++d) gdim[d] = (cuuint64_t)h->ext_n[d];
        gdim[4] = (cuuint64_t)P;
        for (int d = 1; d < 4; ++d) gstr[d - 1] = (cuuint64_t)h->stride[d] * 8;
        gstr[3] = (cuuint64_t)h->S_ext * 8;
        void *base = h->d_J + (size_t)s * h->slot_elems_J();
        if (enc(&ss->maps[s], CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 5, base, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) {
            delete ss;
            return;
        }
    }
    if (!(C == 9 ? stream_attr<9, 5>(ss->smem) : stream_attr<6, 4>(ss->smem))) {
        cudaGetLastError();
        delete ss;
        return;
    }
    if (std::getenv("BELLMAN_TILE_DEBUG"))
        std::fprintf(stderr,
                     "bellman stream: T = %d %d %d %d, lo = %d %d %d %d hi = %d %d %d %d, slab = %d %d %d, NF1 = %d, W3 = %d, "
                     "smem = %zu KB, %d threads (NP = %d), grid = %d x %d x %d\n",
                     tp.T0, tp.T1, tp.T2, tp.T3, lo[0], lo[1], lo[2], lo[3], hi[0], hi[1], hi[2], hi[3], tp.B0, tp.B1, tp.B2, NF1,
                     tp.W3, ss->smem / 1024, ss->nthreads, tp.NP, tp.ntile[0] * tp.ntile[1], tp.ntile[2] * tp.ntile[3], P);
    h->sstate = ss;
}

cudaError_t stream_launch_for_handle(bellman_handle *h, const StageParams &sp, int slot_next, cudaStream_t st) {
    auto *ss = static_cast<StreamState *>(h->sstate);
    if (!ss) return cudaErrorNotSupported;
    const StreamParams &tp = ss->tp;
    const dim3 grid((unsigned)(tp.ntile[0] * tp.ntile[1]), (unsigned)(tp.ntile[2] * tp.ntile[3]), (unsigned)sp.P);
    const bool i32 = sp.idx_bytes == 4;
    if (ss->C == 9 && i32) k_stage_stream<9, 5, true><<<grid, ss->nthreads, ss->smem, st>>>(sp, tp, ss->maps[slot_next]);
    else if (ss->C == 9) k_stage_stream<9, 5, false><<<grid, ss->nthreads, ss->smem, st>>>(sp, tp, ss->maps[slot_next]);
    else if (i32) k_stage_stream<6, 4, true><<<grid, ss->nthreads, ss->smem, st>>>(sp, tp, ss->maps[slot_next]);
    else k_stage_stream<6, 4, false><<<grid, ss->nthreads, ss->smem, st>>>(sp, tp, ss->maps[slot_next]);
    return cudaGetLastError();
}

}  // namespace bellman
