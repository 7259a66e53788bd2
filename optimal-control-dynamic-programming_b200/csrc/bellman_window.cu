// bellman_window.cu — the D = 2 stage kernel with the J_{k+1} neighbourhood staged in shared
// memory by TMA (sm_100a).  Same normative arithmetic as the direct kernel (include/bellman.h),
// compiled with -fmad=false; only the data movement differs.
//
// One CTA = one 32 x 64 tile of the state grid (dimension 0, the contiguous one, across the lanes
// of a warp).  The control loop is cut into chunks; for each (tile, chunk) the bounding box of the
// queried cells is a small window of J_{k+1} because the next-state map is affine in the state
// and monotone in the control index.  A single elected thread issues cp.async.bulk.tensor (TMA)
// box loads of that window into a double-buffered shared-memory ring, completion is tracked by
// mbarriers, and all 256 threads then gather their four corners from shared memory
// (conflict-free: the box pitch is a multiple of 16 doubles and lanes walk dimension 0).
// Every thread owns 8 states of one grid row and loops the controls itself, so min/argmin needs
// no cross-thread reduction and the first-index tie rule falls out of the strict compare.
#include <cuda.h>
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <type_traits>
#include <vector>

#include "bellman_handle.h"
#include "bellman_kernels.cuh"

namespace bellman {

namespace {

constexpr int WT0 = 32;        // tile extent along dimension 0 (one warp)
constexpr int WNT = 256;       // threads per CTA
// states per thread R (same row, R consecutive columns): 8 normally (tile 32 x 64), 4 for the
// small-control CHAIN problems (tile 32 x 32, half the registers, twice the CTAs per SM)
constexpr int tile1_of(int R) { return (WNT / 32) * R; }

struct WindowParams {
    int win0, win1;            // window extent (cells): rows (dim 0, pitch) and columns
    int boxes, box1;           // TMA boxes per window along dim 1, columns per box
    int cchunk, nchunks;
    int ntile0, ntile1;
    int tj_fastest;            // 1: consecutive CTAs walk dimension 1 (the one the controls sweep)
    int buf_doubles;           // doubles per ring slot (the window, 128 B aligned)
    const double *cmm;         // [P][nchunks][4]: min/max of Tc_0, min/max of Tc_1 per chunk
    // per-tile-index min/max of the state-indexed tables, computed once on the host:
    // tmm[p * tmm_stride + tmm_off[d][ab] + 2*t + {0: min, 1: max}], ab = 0 for Ta_d, 1 for Tb_d,
    // t = tile index along the dimension that indexes that table
    const double *tmm;
    int tmm_off[2][2];
    int tmm_stride;
    // canonical state tables (built on the host from Ta/Tb/q; IEEE addition is commutative, so
    // base_d = row_d[i] + col_d[j] and gs = qrow[i] + qcol[j] reproduce the normative sums):
    // rowpack[p][i] = {row_0, row_1, qrow, 0},  colpack[p][j] = {col_0, col_1, qcol, 0}
    const double4 *rowpack, *colpack;
    int col0_zero, col1_zero;  // the column part of dimension 0 / 1 is absent (all zeros)
    // k_stage_strip: colq[p][j] = {col_1, qcol} (one 16-byte load per column), strip_r = columns
    // walked by one warp
    const double2 *colq;
    int strip_r;
    int pf_dist;               // k_stage_strip: L2 prefetch distance in CTAs (0 = off)
    int tj_off;                // k_stage_wide: first dimension-1 tile of this launch (ntile1 = tiles launched)
};

// --- PTX wrappers ------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
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
// bounded wait (traps instead of hanging if a phase never completes) and plain arrive
__device__ __forceinline__ uint32_t mbar_try_hint(uint32_t addr, uint32_t parity) {
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
__device__ __noinline__ void mbar_wait_slow_w(uint32_t addr, uint32_t parity) {
    for (uint32_t spins = 0; !mbar_try_hint(addr, parity); ++spins)
        if (spins > (1u << 22)) asm volatile("trap;");
}
__device__ __forceinline__ void mbar_wait_bounded(uint64_t *bar, uint32_t parity) {
    const uint32_t addr = smem_u32(bar);
    if (!mbar_try_hint(addr, parity)) mbar_wait_slow_w(addr, parity);
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tma_load_3d(void *dst, const CUtensorMap *map, uint64_t *bar, int x, int y,
                                            int z) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::
            "r"(smem_u32(dst)),
        "l"(map), "r"(smem_u32(bar)), "r"(x), "r"(y), "r"(z)
        : "memory");
}

// TMA prefetch of a box into L2 (no shared memory, no barrier): used to warm L2 for the tile a later
// CTA on this SM will stage, so that its TMA load sees L2 latency instead of DRAM latency
__device__ __forceinline__ void tma_prefetch_3d(const CUtensorMap *map, int x, int y, int z) {
    asm volatile("cp.async.bulk.prefetch.tensor.3d.L2.global.tile [%0, {%1, %2, %3}];" ::"l"(map), "r"(x), "r"(y), "r"(z)
                 : "memory");
}

// shared-memory load by 32-bit address (keeps the address arithmetic in 32-bit integer registers).
// Not volatile: the caller orders it after the mbarrier wait through a data dependence.
__device__ __forceinline__ double lds_f64(uint32_t addr) {
    double v;
    asm("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr));
    return v;
}

// UNIFORM locate of include/bellman.h: the tables are pre-scaled to cell units on the host, so the
// query g IS the fractional cell coordinate: cell = clamp(floor(g)), t = g - cell.
__device__ __forceinline__ int cell_uniform(double g, int n) {
    return min(max(__double2int_rd(g), 0), n - 2);
}
// (double)cell for 0 <= cell < 2^31 without a conversion instruction: the double with high word
// 0x43300000 and low word cell is exactly 2^52 + cell.  Measured on B200 (scripts/ubench/pipes.cu):
// F2I.F64 / I2F.F64 issue at 16 lanes/clk/SM (2 cycles per warp instruction), a DADD at 64 — and
// the conversion pipe only partly overlaps with the shared-memory loads.
__device__ __forceinline__ double cell_to_double(int cell) {
    return __hiloint2double(0x43300000, cell) - 4503599627370496.0;
}
// XUCVT: take (double)cell from the conversion pipe (I2F.F64) instead of the fp64 pipe.  The hot
// loop uses one of each per update so that neither pipe carries both conversions.
// locate_magic: floor on the fp64 pipe for queries known to be interior (|g| < 2^31, no clamp).
// M = 2^52 + 2^51: g + M rounded toward -inf is floor(g) + M exactly (ulp 1 in [2^52, 2^53)), its low
// word is floor(g) in two's complement, and (g + M) - M is floor(g) as a double, exactly — so
// t = g - floor(g) is the same exact difference the conversion path forms.  Three DADDs (64
// lanes/clk/SM) instead of F2I + I2F (16 lanes/clk/SM each, issued through the same MIO queue as the
// shared-memory loads).
__device__ __forceinline__ int locate_magic(double g, double &t) {
    const double M = 6755399441055744.0;
    const double s = __dadd_rd(g, M);
    t = g - (s - M);
    return __double2loint(s);
}

// The clamped locate without the conversion pipe, for queries known to satisfy |g| < 2^31 (the host checks
// the bound): floor from the low word of g +(rd) M as above, clamp on the integer pipe, (double)cell by the
// bit trick.  Same cell and the same exact difference g - cell as locate_uniform<true> — F2I.F64 / I2F.F64
// occupy a scheduler's conversion pipe for 8 cycles per warp instruction, a DADD the fp64 pipe for 2.
// clamp(x, 0, hi) in one instruction (VIMNMX.RELU)
__device__ __forceinline__ int clamp_relu(int x, int hi) {
    int r;
    asm("min.relu.s32 %0, %1, %2;" : "=r"(r) : "r"(x), "r"(hi));
    return r;
}
template <bool XUCVT = false>
__device__ __forceinline__ int locate_magic_clamp(double g, int n, double &t) {
    const double M = 6755399441055744.0;
    const int cell = clamp_relu(__double2loint(__dadd_rd(g, M)), n - 2);
    t = g - (XUCVT ? (double)cell : cell_to_double(cell));   // XUCVT: one I2F.F64 instead of the bit trick's DADD
    return cell;
}
// one dimension of k_stage_wide's locate: CLAMPED as above, else the interior form (floor(g) as a double from
// `s - M`, or from the conversion pipe when XUCVT)
template <bool CLAMPED, bool XUCVT>
__device__ __forceinline__ int locate_dim(double g, int n, double &t) {
    if (CLAMPED) return locate_magic_clamp<XUCVT>(g, n, t);
    const double M = 6755399441055744.0;
    const double s = __dadd_rd(g, M);
    const int cell = __double2loint(s);
    t = g - (XUCVT ? (double)cell : s - M);
    return cell;
}

template <bool CLAMP = true, bool XUCVT = false>
__device__ __forceinline__ int locate_uniform(double g, int n, double &t) {
    int cell = __double2int_rd(g);
    if (CLAMP) cell = min(max(cell, 0), n - 2);   // skipped when the whole chunk is known to be interior
    t = g - (XUCVT ? (double)cell : cell_to_double(cell));
    return cell;
}

// BATCH: states evaluated together (instruction-level parallelism); OCC: CTAs per SM the register
// allocation is sized for.
// CHAIN (requires !HC1 and a dimension-0 query that does not depend on the dimension-1 index, e.g.
// Solver_attitude: w' = w[i] + D[c], theta' = theta[j] + Wt[i]): the 8 states of a thread share
// cell and weight along dimension 0, and their dimension-1 cells are consecutive, so per control
// the thread interpolates 9 columns once along dimension 0 and every state blends two neighbouring
// columns.  Same operations on the same operands as the generic path (bit-identical results),
// 18 shared-memory loads and ~61 fp64 instructions per 8 updates instead of 32 and ~120.
// W0C: the window pitch (rows per column) as a compile-time constant when it is one of the common
// values (0 = read it from the parameters); turns the second-column offset into an immediate.
// LOC: how the interior (unclamped) control loop locates — 0: dim 0 F2I + bit-trick, dim 1 F2I + I2F;
// 1: dim 0 on the fp64 pipe (locate_magic), dim 1 F2I + I2F; 2: both on the fp64 pipe.
template <bool HC0, bool HC1, bool CHAIN, int BATCH, int OCC, int WR_STATES, int W0C = 0, int LOC = 0>
__global__ void __launch_bounds__(WNT, OCC)
k_stage_window(const __grid_constant__ StageParams sp, const __grid_constant__ WindowParams wp,
               const __grid_constant__ CUtensorMap tmap) {
    constexpr int WT1 = (WNT / 32) * WR_STATES;
    // ring of two window slots, win0 x win1 doubles each (dimension 0 contiguous)
    extern __shared__ __align__(128) double ring[];
    __shared__ __align__(8) uint64_t mbar[2];

    const int tid = threadIdx.x, lane = tid & 31, wrp = tid >> 5;
    const int prob = blockIdx.y;
    // CTAs that run together should sweep the same columns of J_{k+1} at staggered times: that
    // keeps the union of their windows a thin band that lives in L2
    const int ti = wp.tj_fastest ? blockIdx.x / wp.ntile1 : blockIdx.x % wp.ntile0;
    const int tj = wp.tj_fastest ? blockIdx.x % wp.ntile1 : blockIdx.x / wp.ntile0;
    const DimParams &d0 = sp.dim[0], &d1 = sp.dim[1];
    const int n0 = d0.n, n1 = d1.n;
    // tile ranges in global grid indices, clipped to the owned range
    const int i_lo = d0.own_lo + ti * WT0, i_hi = min(i_lo + WT0, d0.own_lo + d0.own_n);
    const int j_lo = d1.own_lo + tj * WT1, j_hi = min(j_lo + WT1, d1.own_lo + d1.own_n);

    const int win_elems = wp.win0 * wp.win1;
    const uint32_t win_bytes = (uint32_t)win_elems * 8u;

    // per-problem tables
    const double *Tc0 = HC0 ? d0.Tc + (size_t)prob * sp.C : nullptr;
    const double *Tc1 = HC1 ? d1.Tc + (size_t)prob * sp.C : nullptr;
    const double *rr = sp.r + (size_t)prob * sp.C;
    const double *cmm = wp.cmm + (size_t)prob * wp.nchunks * 4;

    if (tid == 0) {
        mbar_init(&mbar[0], 1);
        mbar_init(&mbar[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    // extrema of the state-indexed tables over this tile: precomputed on the host per tile index
    // (no in-kernel reduction).  Long control loops keep them in shared memory (registers are the
    // scarce resource there); the small-control CHAIN kernels keep them in registers so that no
    // warp waits on another warp's global load.
    __shared__ double tmm_s[8];
    double tmm_r[8];
    auto tmm_load = [&](int k) -> double {
        const double *tm = wp.tmm + (size_t)prob * wp.tmm_stride;
        const int d = k >> 2, ab = (k >> 1) & 1, mx = k & 1;
        const DimParams &dd = d == 0 ? d0 : d1;
        if (ab == 1 && !dd.Tb) return 0.0;
        const int src = ab == 0 ? dd.src_a : dd.src_b;
        return __ldg(tm + wp.tmm_off[d][ab] + 2 * (src == 0 ? ti : tj) + mx);
    };
    if (CHAIN) {
#pragma unroll
        for (int k = 0; k < 8; ++k) tmm_r[k] = tmm_load(k);
    } else if (tid < 8) {
        tmm_s[tid] = tmm_load(tid);
    }
    __syncthreads();   // mbarrier init (and tmm_s) visible to every thread
    const double *tmm = CHAIN ? tmm_r : tmm_s;

    // window origin of chunk ch: the cell of the smallest query, formed with the kernel's own
    // association so that the bound is exact
    auto origin = [&](int ch, int &r0, int &c0) {
        double lo0 = tmm[0], lo1 = tmm[4];
        if (d0.Tb) lo0 = lo0 + tmm[2];
        if (d1.Tb) lo1 = lo1 + tmm[6];
        if (HC0) lo0 = lo0 + __ldg(cmm + 4 * ch);
        if (HC1) lo1 = lo1 + __ldg(cmm + 4 * ch + 2);
        // TMA needs the byte offset of the innermost box coordinate to be a multiple of 16
        // (measured on B200: an odd fp64 coordinate raises "illegal instruction"), so the window
        // starts on an even row of the local array; the planner adds the extra row.
        r0 = cell_uniform(lo0, n0);
        r0 -= (r0 - d0.ext_lo) & 1;
        c0 = cell_uniform(lo1, n1);
    };
    auto issue = [&](int ch) {   // thread 0 only
        int r0, c0;
        origin(ch, r0, c0);
        uint64_t *bar = &mbar[ch & 1];
        mbar_expect_tx(bar, win_bytes);
        double *dst = ring + (ch & 1) * wp.buf_doubles;
        for (int b = 0; b < wp.boxes; ++b)
            tma_load_3d(dst + (size_t)b * wp.box1 * wp.win0, &tmap, bar, r0 - d0.ext_lo,
                        c0 - d1.ext_lo + b * wp.box1, prob);
    };
    if (tid == 0) {
        issue(0);
        if (wp.nchunks > 1) issue(1);
    }

    // this thread's states: row i, columns j0..j0+7
    const int i = min(i_lo + lane, i_hi - 1);
    const int jbase = j_lo + wrp * WR_STATES;
    double base0[WR_STATES], base1[WR_STATES], gs[WR_STATES], best[WR_STATES];
    int arg[WR_STATES];
    int cellK0[WR_STATES], cellK1[WR_STATES];   // used only for control-independent dimensions
    double tK0[WR_STATES], tK1[WR_STATES];
    // canonical packed tables: one 32-byte load per row, one per column
    const double2 *rpk = reinterpret_cast<const double2 *>(wp.rowpack + (size_t)prob * n0 + i);
    const double2 rp01 = __ldg(rpk), rp23 = __ldg(rpk + 1);
    const double2 *cpk = reinterpret_cast<const double2 *>(wp.colpack + (size_t)prob * n1);
#pragma unroll
    for (int m = 0; m < WR_STATES; ++m) {
        const int j = min(jbase + m, j_hi - 1);
        const double2 cp01 = __ldg(cpk + 2 * j), cp23 = __ldg(cpk + 2 * j + 1);
        const double b0 = wp.col0_zero ? rp01.x : rp01.x + cp01.x;
        const double b1 = wp.col1_zero ? rp01.y : rp01.y + cp01.y;
        base0[m] = b0;
        base1[m] = b1;
        gs[m] = rp23.x + cp23.x;
        best[m] = __longlong_as_double(0x7ff0000000000000LL);
        arg[m] = 0;
        if (!HC0) cellK0[m] = locate_uniform<true>(b0, n0, tK0[m]);
        if (!HC1) cellK1[m] = locate_uniform<true>(b1, n1, tK1[m]);
    }
    // CHAIN fast path is taken by a warp only if, for every lane, the dimension-1 cells of the 8
    // states are consecutive (always true in the interior of a uniform grid; ragged tiles and the
    // clamped edge fall back to the generic path)
    bool chain_ok = CHAIN;
    if (CHAIN) {
#pragma unroll
        for (int m = 1; m < WR_STATES; ++m) chain_ok = chain_ok && (cellK1[m] == cellK1[0] + m);
        chain_ok = __all_sync(0xffffffffu, chain_ok);
    }

    const int W0 = W0C ? W0C : wp.win0;
    // the control loop over one staged chunk; CLAMP = false when every query of the chunk falls in
    // an interior cell (decided from the same exact bounds that place the window)
    auto chunk_loop = [&](auto clamp_tag, int ch, const double *__restrict__ Wb) {
        constexpr bool CLAMP = decltype(clamp_tag)::value;
        const int c_end = min(sp.C, (ch + 1) * wp.cchunk);
        for (int c = ch * wp.cchunk; c < c_end; ++c) {
            const double bu0 = HC0 ? __ldg(Tc0 + c) : 0.0;
            const double bu1 = HC1 ? __ldg(Tc1 + c) : 0.0;
            const double rc = __ldg(rr + c);
            if (CHAIN && chain_ok) {
                double t0c;
                const int cell0 = locate_uniform<CLAMP>(base0[0] + bu0, n0, t0c);
                const double *p = Wb + (cellK1[0] * W0 + cell0);
                double a[WR_STATES + 1];
#pragma unroll
                for (int k = 0; k <= WR_STATES; ++k) {
                    const double lo = p[k * W0], hi = p[k * W0 + 1];
                    a[k] = fma(t0c, hi - lo, lo);
                }
#pragma unroll
                for (int m = 0; m < WR_STATES; ++m) {
                    const double v = fma(tK1[m], a[m + 1] - a[m], a[m]);
                    const double tot = (gs[m] + rc) + v;
                    if (tot < best[m]) { best[m] = tot; arg[m] = c; }
                }
                continue;
            }
            // two batches of four independent states: enough instruction-level parallelism to
            // cover the fp64 / shared-memory latencies with 16 warps per SM, within 128 registers
#pragma unroll
            for (int mb = 0; mb < WR_STATES; mb += BATCH) {
                int off[BATCH];
                double t0[BATCH], t1[BATCH];
#pragma unroll
                for (int u = 0; u < BATCH; ++u) {
                    const int m = mb + u;
                    int cell0, cell1;
                    if (HC0) cell0 = (!CLAMP && LOC >= 1) ? locate_magic(base0[m] + bu0, t0[u])
                                                          : locate_uniform<CLAMP>(base0[m] + bu0, n0, t0[u]);
                    else { cell0 = cellK0[m]; t0[u] = tK0[m]; }
                    if (HC1) cell1 = (!CLAMP && LOC >= 2) ? locate_magic(base1[m] + bu1, t1[u])
                                                          : locate_uniform<CLAMP, true>(base1[m] + bu1, n1, t1[u]);
                    else { cell1 = cellK1[m]; t1[u] = tK1[m]; }
                    off[u] = cell1 * W0 + cell0;      // Wb already carries the window origin
                }
                double v00[BATCH], v10[BATCH], v01[BATCH], v11[BATCH];
#pragma unroll
                for (int u = 0; u < BATCH; ++u) {
                    const double *p = Wb + off[u];
                    v00[u] = p[0]; v10[u] = p[1]; v01[u] = p[W0]; v11[u] = p[W0 + 1];
                }
#pragma unroll
                for (int u = 0; u < BATCH; ++u) {
                    const int m = mb + u;
                    const double a = fma(t0[u], v10[u] - v00[u], v00[u]);
                    const double b = fma(t0[u], v11[u] - v01[u], v01[u]);
                    const double v = fma(t1[u], b - a, a);
                    const double tot = (gs[m] + rc) + v;
                    if (tot < best[m]) { best[m] = tot; arg[m] = c; }
                }
            }
        }
    };

    for (int ch = 0; ch < wp.nchunks; ++ch) {
        // exact bounds of this chunk's queries, formed with the kernel's own association
        double lo0 = tmm[0], hi0 = tmm[1], lo1 = tmm[4], hi1 = tmm[5];
        if (d0.Tb) { lo0 = lo0 + tmm[2]; hi0 = hi0 + tmm[3]; }
        if (d1.Tb) { lo1 = lo1 + tmm[6]; hi1 = hi1 + tmm[7]; }
        if (HC0) { lo0 = lo0 + __ldg(cmm + 4 * ch); hi0 = hi0 + __ldg(cmm + 4 * ch + 1); }
        if (HC1) { lo1 = lo1 + __ldg(cmm + 4 * ch + 2); hi1 = hi1 + __ldg(cmm + 4 * ch + 3); }
        const bool interior = lo0 >= 0.0 && hi0 < (double)(n0 - 1) && lo1 >= 0.0 && hi1 < (double)(n1 - 1);
        int r0, c0;
        origin(ch, r0, c0);
        mbar_wait(&mbar[ch & 1], (ch >> 1) & 1);
        const double *__restrict__ Wb = ring + (ch & 1) * wp.buf_doubles - (c0 * W0 + r0);
        if (interior) chunk_loop(std::false_type{}, ch, Wb);
        else chunk_loop(std::true_type{}, ch, Wb);
        __syncthreads();   // every thread is done with this buffer
        if (tid == 0 && ch + 2 < wp.nchunks) issue(ch + 2);
    }

    if (i_lo + lane < i_hi) {
        double *jo = sp.J_out + (size_t)prob * sp.S_ext + (long long)(i - d0.ext_lo) * d0.stride +
                     (long long)(jbase - d1.ext_lo) * d1.stride;
        long long io = (long long)prob * sp.S_own + (long long)(i - d0.own_lo) +
                      (long long)(jbase - d1.own_lo) * d0.own_n;
#pragma unroll
        for (int m = 0; m < WR_STATES; ++m) {
            if (jbase + m < j_hi) {
                jo[(long long)m * d1.stride] = best[m];
                idx_store(sp.idx_out, sp.idx_bytes, io + (long long)m * d0.own_n, arg[m]);
            }
        }
        if (sp.n_peers) {
#pragma unroll   // keeps best[] in registers (a rolled loop would index it dynamically)
            for (int m = 0; m < WR_STATES; ++m)
                if (jbase + m < j_hi) { const int gi[2] = {i, jbase + m}; peer_store<2>(sp, prob, gi, best[m]); }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// k_stage_wide: the long-control-loop case (both dimensions depend on the control, e.g. Dynamic_Solver's
// x' = A x + B u) re-cut for instruction-level parallelism.  Same tile (32 x 64), same window ring and
// same operations per update as k_stage_window, but
//   * 512 threads own 4 states each instead of 256 threads owning 8: the per-state registers (two query
//     bases, the state cost, best, argmin = 9 per state) drop from 72 to 36 of the 128 a thread may
//     hold, which is what lets the four updates of a control interleave instruction by instruction (in
//     k_stage_window ptxas serialises them to stay under 128: ~7 cycles per issued instruction per warp);
//   * the control tables Tc_0, Tc_1, r come from the constant bank (a 12 KB kernel parameter) instead of
//     three global loads and ten address instructions per control;
//   * shared-memory addresses are 32-bit (two IMADs per update).
// One CTA per SM (16 warps), NS ring slots.
// ---------------------------------------------------------------------------------------------
#ifndef WIDE_UNROLL
#define WIDE_UNROLL 2      // controls per iteration of k_stage_wide's loop (measured: 1 -> +4 %, 4 -> same as 2)
#endif
constexpr int WIDE_UNROLL_N = WIDE_UNROLL;
constexpr int WIDE_NT = 512, WIDE_R = 4, WIDE_MAXC = 512, WIDE_MAXCH = 128;
struct WideTables {
    double tc0[WIDE_MAXC], tc1[WIDE_MAXC], r[WIDE_MAXC];   // [P * C] each
};

// BAR: a CTA barrier ends every chunk (all 16 warps drain and refill their pipelines together).  Otherwise
// the slot is handed back through an `empty` mbarrier (one arrival per warp) and the warp whose turn it is
// (chunk index mod 16) waits for it and issues the refill, so the other warps run on into the next chunk
// and their pipeline bubbles no longer coincide.
// XU: how many of the two floor(g)-as-double values of an interior update come from the conversion pipe
// (I2F.F64 of the cell the low word of g +(rd) M already holds) instead of the fp64 pipe's `s - M`: the
// same exact value, one DADD fewer per dimension on the pipe that bounds the kernel.
template <int NS, int W0C, bool IDX32, bool BAR, int XU = 0>
__global__ void __launch_bounds__(WIDE_NT, 1)
k_stage_wide(const __grid_constant__ StageParams sp, const __grid_constant__ WindowParams wp,
             const __grid_constant__ CUtensorMap tmap, const __grid_constant__ WideTables tb) {
    constexpr int R = WIDE_R, WT1 = (WIDE_NT / 32) * R;
    extern __shared__ __align__(128) double ring[];
    __shared__ __align__(8) uint64_t mbar[NS], empty[NS];
    __shared__ int4 cinfo[WIDE_MAXCH];

    const int tid = threadIdx.x, lane = tid & 31, wrp = tid >> 5;
    const int prob = blockIdx.y;
    const int ti = wp.tj_fastest ? blockIdx.x / wp.ntile1 : blockIdx.x % wp.ntile0;
    const int tj = wp.tj_off + (wp.tj_fastest ? blockIdx.x % wp.ntile1 : blockIdx.x / wp.ntile0);
    const DimParams &d0 = sp.dim[0], &d1 = sp.dim[1];
    const int n0 = d0.n, n1 = d1.n;
    const int i_lo = d0.own_lo + ti * WT0, i_hi = min(i_lo + WT0, d0.own_lo + d0.own_n);
    const int j_lo = d1.own_lo + tj * WT1, j_hi = min(j_lo + WT1, d1.own_lo + d1.own_n);
    const uint32_t win_bytes = (uint32_t)(wp.win0 * wp.win1) * 8u;
    const double *cmm = wp.cmm + (size_t)prob * wp.nchunks * 4;
    const int pc = prob * sp.C;

    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < NS; ++s) { mbar_init(&mbar[s], 1); mbar_init(&empty[s], WIDE_NT / 32); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    // Chunk table, one entry per chunk, built once per CTA by the first nchunks threads: the window origin
    // (TMA coordinates and element offset) and how the chunk's queries may be located.  The bounds are exact:
    // per-tile / per-chunk table extrema summed with the kernel's own association (as k_stage_window).
    //   mode bit 0 / bit 1: dimension 0 / 1 has queries outside the interior cells and is clamped (|g| < 2^30: no
    //   F2I); mode 4: clamped through the conversion pipe (queries beyond 2^30 cells: never in practice)
    if (tid < wp.nchunks) {
        const double *tm = wp.tmm + (size_t)prob * wp.tmm_stride;
        auto tmm_load = [&](int k) -> double {
            const int d = k >> 2, ab = (k >> 1) & 1, mx = k & 1;
            const DimParams &dd = d == 0 ? d0 : d1;
            if (ab == 1 && !dd.Tb) return 0.0;
            const int src = ab == 0 ? dd.src_a : dd.src_b;
            return __ldg(tm + wp.tmm_off[d][ab] + 2 * (src == 0 ? ti : tj) + mx);
        };
        double lo0 = tmm_load(0), hi0 = tmm_load(1), lo1 = tmm_load(4), hi1 = tmm_load(5);
        if (d0.Tb) { lo0 = lo0 + tmm_load(2); hi0 = hi0 + tmm_load(3); }
        if (d1.Tb) { lo1 = lo1 + tmm_load(6); hi1 = hi1 + tmm_load(7); }
        lo0 = lo0 + __ldg(cmm + 4 * tid); hi0 = hi0 + __ldg(cmm + 4 * tid + 1);
        lo1 = lo1 + __ldg(cmm + 4 * tid + 2); hi1 = hi1 + __ldg(cmm + 4 * tid + 3);
        int r0 = cell_uniform(lo0, n0);
        r0 -= (r0 - d0.ext_lo) & 1;                  // TMA: even innermost coordinate
        const int c0 = cell_uniform(lo1, n1);
        const double BIG = 1073741824.0;
        const bool in0 = lo0 >= 0.0 && hi0 < (double)(n0 - 1), in1 = lo1 >= 0.0 && hi1 < (double)(n1 - 1);
        const bool small = lo0 > -BIG && hi0 < BIG && lo1 > -BIG && hi1 < BIG;
        cinfo[tid] = make_int4(r0 - d0.ext_lo, c0 - d1.ext_lo, c0 * (W0C ? W0C : wp.win0) + r0,
                               small ? (in0 ? 0 : 1) | (in1 ? 0 : 2) : 4);
    }
    __syncthreads();

    auto issue = [&](int ch) {   // thread 0 only
        const int4 ci = cinfo[ch];
        const int s = ch % NS;
        mbar_expect_tx(&mbar[s], win_bytes);
        double *dst = ring + s * wp.buf_doubles;
        for (int b = 0; b < wp.boxes; ++b)
            tma_load_3d(dst + (size_t)b * wp.box1 * wp.win0, &tmap, &mbar[s], ci.x, ci.y + b * wp.box1, prob);
    };
    if (tid == 0)
        for (int ch = 0; ch < NS && ch < wp.nchunks; ++ch) issue(ch);

    // this thread's states: row i, columns jbase .. jbase + 3
    const int i = min(i_lo + lane, i_hi - 1);
    const int jbase = j_lo + wrp * R;
    double base0[R], base1[R], gs[R], best[R];
    int arg[R];
    {
        const double2 *rpk = reinterpret_cast<const double2 *>(wp.rowpack + (size_t)prob * n0 + i);
        const double2 rp01 = __ldg(rpk), rp23 = __ldg(rpk + 1);
        const double2 *cpk = reinterpret_cast<const double2 *>(wp.colpack + (size_t)prob * n1);
#pragma unroll
        for (int m = 0; m < R; ++m) {
            const int j = min(jbase + m, j_hi - 1);
            const double2 cp01 = __ldg(cpk + 2 * j), cp23 = __ldg(cpk + 2 * j + 1);
            base0[m] = wp.col0_zero ? rp01.x : rp01.x + cp01.x;
            base1[m] = wp.col1_zero ? rp01.y : rp01.y + cp01.y;
            gs[m] = rp23.x + cp23.x;
            best[m] = __longlong_as_double(0x7ff0000000000000LL);
            arg[m] = 0;
        }
    }

    const uint32_t PB = (uint32_t)(W0C ? W0C : wp.win0) * 8u;      // window pitch in bytes
    const uint32_t ring_u32 = smem_u32(ring);

    // MODE: the chunk-table mode (which dimensions are clamped)
    auto chunk_loop = [&](auto mode_tag, int ch, uint32_t wb) {
        constexpr int MODE = decltype(mode_tag)::value;
        const int c_end = min(sp.C, (ch + 1) * wp.cchunk);
#pragma unroll WIDE_UNROLL_N
        for (int c = ch * wp.cchunk; c < c_end; ++c) {
            const double bu0 = tb.tc0[pc + c], bu1 = tb.tc1[pc + c], rc = tb.r[pc + c];
            uint32_t a[R];
            double t0[R], t1[R];
#pragma unroll
            for (int u = 0; u < R; ++u) {
                int cell0, cell1;
                if (MODE == 4) {
                    cell0 = locate_uniform<true>(base0[u] + bu0, n0, t0[u]);
                    cell1 = locate_uniform<true, true>(base1[u] + bu1, n1, t1[u]);
                } else {
                    cell0 = locate_dim<(MODE & 1) != 0, (XU >= 2)>(base0[u] + bu0, n0, t0[u]);
                    cell1 = locate_dim<(MODE & 2) != 0, (XU >= 1)>(base1[u] + bu1, n1, t1[u]);
                }
                a[u] = (uint32_t)cell1 * PB + ((uint32_t)cell0 * 8u + wb);
            }
            double v00[R], v10[R], v01[R], v11[R];
#pragma unroll
            for (int u = 0; u < R; ++u) {
                v00[u] = lds_f64(a[u]);
                v10[u] = lds_f64(a[u] + 8u);
                v01[u] = lds_f64(a[u] + PB);
                v11[u] = lds_f64(a[u] + PB + 8u);
            }
#pragma unroll
            for (int u = 0; u < R; ++u) {
                const double x = fma(t0[u], v10[u] - v00[u], v00[u]);
                const double y = fma(t0[u], v11[u] - v01[u], v01[u]);
                const double v = fma(t1[u], y - x, x);
                const double tot = (gs[u] + rc) + v;
                if (tot < best[u]) { best[u] = tot; arg[u] = c; }
            }
        }
    };

    for (int ch = 0; ch < wp.nchunks; ++ch) {
        const int4 ci = cinfo[ch];
        const int s = ch % NS;
        mbar_wait_bounded(&mbar[s], (uint32_t)(ch / NS) & 1u);
        // byte address of window element (cell0 = 0, cell1 = 0); passes through a volatile asm placed after
        // the wait so that no window load is scheduled above it
        uint32_t wb = ring_u32 + (uint32_t)(s * wp.buf_doubles) * 8u - (uint32_t)ci.z * 8u;
        asm volatile("" : "+r"(wb)::"memory");
        if (ci.w == 0) chunk_loop(std::integral_constant<int, 0>{}, ch, wb);
        else if (ci.w == 2) chunk_loop(std::integral_constant<int, 2>{}, ch, wb);
        else if (ci.w == 1) chunk_loop(std::integral_constant<int, 1>{}, ch, wb);
        else if (ci.w == 3) chunk_loop(std::integral_constant<int, 3>{}, ch, wb);
        else chunk_loop(std::integral_constant<int, 4>{}, ch, wb);
        if (BAR) {
            __syncthreads();   // every thread is done with this slot
            if (tid == 0 && ch + NS < wp.nchunks) issue(ch + NS);
        } else if (ch + NS < wp.nchunks) {
            __syncwarp();      // every lane's window loads of this chunk have returned (their values were consumed)
            if (lane == 0) {
                mbar_arrive(&empty[s]);
                if (wrp == (ch & (WIDE_NT / 32 - 1))) {
                    mbar_wait_bounded(&empty[s], (uint32_t)(ch / NS) & 1u);
                    issue(ch + NS);
                }
            }
            __syncwarp();
        }
    }

    if (i_lo + lane < i_hi) {
        double *jo = sp.J_out + (size_t)prob * sp.S_ext + (long long)(i - d0.ext_lo) * d0.stride +
                     (long long)(jbase - d1.ext_lo) * d1.stride;
        const long long io = (long long)prob * sp.S_own + (long long)(i - d0.own_lo) +
                             (long long)(jbase - d1.own_lo) * d0.own_n;
#pragma unroll
        for (int m = 0; m < R; ++m) {
            if (jbase + m < j_hi) {
                jo[(long long)m * d1.stride] = best[m];
                if (IDX32) sp.idx_out[io + (long long)m * d0.own_n] = arg[m];
                else idx_store(sp.idx_out, sp.idx_bytes, io + (long long)m * d0.own_n, arg[m]);
            }
        }
        if (sp.n_peers) {
#pragma unroll
            for (int m = 0; m < R; ++m)
                if (jbase + m < j_hi) { const int gi[2] = {i, jbase + m}; peer_store<2>(sp, prob, gi, best[m]); }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// k_stage_chain: the CHAIN structure (see k_stage_window) as a kernel of its own for problems with
// at most MAXC controls (Solver_attitude: 3 torque levels).  Every control's window is in flight
// from the first instruction (one TMA box per control, one mbarrier), there is no chunk loop, no
// ring reuse and a single __syncthreads per tile; window origins are computed once by thread 0 and
// broadcast through shared memory.  Same operations on the same operands as the generic path.
// ---------------------------------------------------------------------------------------------
template <int R, int MAXC>
__global__ void __launch_bounds__(WNT, 4)
k_stage_chain(const __grid_constant__ StageParams sp, const __grid_constant__ WindowParams wp,
              const __grid_constant__ CUtensorMap tmap) {
    constexpr int WT1 = (WNT / 32) * R;
    extern __shared__ __align__(128) double ring[];    // C windows, win0 x win1 doubles each
    __shared__ __align__(8) uint64_t mbar;
    __shared__ int org[MAXC][2];

    const int tid = threadIdx.x, lane = tid & 31, wrp = tid >> 5;
    const int prob = blockIdx.y;
    const int ti = wp.tj_fastest ? blockIdx.x / wp.ntile1 : blockIdx.x % wp.ntile0;
    const int tj = wp.tj_fastest ? blockIdx.x % wp.ntile1 : blockIdx.x / wp.ntile0;
    const DimParams &d0 = sp.dim[0], &d1 = sp.dim[1];
    const int n0 = d0.n, n1 = d1.n, C = sp.C, W0 = wp.win0;
    const int win_elems = wp.win0 * wp.win1;
    const int i_lo = d0.own_lo + ti * WT0, i_hi = min(i_lo + WT0, d0.own_lo + d0.own_n);
    const int j_lo = d1.own_lo + tj * WT1, j_hi = min(j_lo + WT1, d1.own_lo + d1.own_n);
    const double *Tc0 = d0.Tc + (size_t)prob * C;
    const double *rr = sp.r + (size_t)prob * C;

    // lanes 0..C-1 of warp 0 each place and issue one control's window (their table loads overlap)
    if (wrp == 0) {
        if (lane == 0) {
            mbar_init(&mbar, (uint32_t)C);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncwarp();
        if (lane < C) {
            const int c = lane;
            // smallest query of the tile for this control, formed with the kernel's own association
            const double *tm = wp.tmm + (size_t)prob * wp.tmm_stride;
            const double *cm = wp.cmm + (size_t)prob * wp.nchunks * 4;
            double lo0 = __ldg(tm + wp.tmm_off[0][0] + 2 * ti);                  // Ta_0 is indexed by the row
            if (d0.Tb) lo0 = lo0 + __ldg(tm + wp.tmm_off[0][1] + 2 * ti);        // (and so is Tb_0 in CHAIN problems)
            double lo1 = __ldg(tm + wp.tmm_off[1][0] + 2 * (d1.src_a == 0 ? ti : tj));
            if (d1.Tb) lo1 = lo1 + __ldg(tm + wp.tmm_off[1][1] + 2 * (d1.src_b == 0 ? ti : tj));
            const int c0 = cell_uniform(lo1, n1);
            int r0 = cell_uniform(lo0 + __ldg(cm + 4 * c), n0);     // chunk size is 1: cmm[c] = Tc_0[c]
            r0 -= (r0 - d0.ext_lo) & 1;                              // TMA: even innermost coordinate
            org[c][0] = r0;
            org[c][1] = c0;
            mbar_expect_tx(&mbar, (uint32_t)win_elems * 8u);
            for (int b = 0; b < wp.boxes; ++b)
                tma_load_3d(ring + (size_t)c * wp.buf_doubles + (size_t)b * wp.box1 * W0, &tmap, &mbar, r0 - d0.ext_lo,
                            c0 - d1.ext_lo + b * wp.box1, prob);
        }
    }

    // this thread's R states: row i, columns jbase .. jbase+R-1 (loads overlap the TMA latency)
    const int i = min(i_lo + lane, i_hi - 1);
    const int jbase = j_lo + wrp * R;
    const double2 *rpk = reinterpret_cast<const double2 *>(wp.rowpack + (size_t)prob * n0 + i);
    const double2 rp01 = __ldg(rpk), rp23 = __ldg(rpk + 1);
    const double2 *cpk = reinterpret_cast<const double2 *>(wp.colpack + (size_t)prob * n1);
    const double base0 = rp01.x;                     // CHAIN: no column part in dimension 0
    double gs[R], tK1[R], best[R];
    int arg[R], cellK1[R];
    bool chain_ok = true;
    const bool full = jbase + R <= j_hi;             // warp-uniform: all R columns exist
    const double2 *cpb = cpk + 2 * (size_t)min(jbase, j_hi - 1);
#pragma unroll
    for (int m = 0; m < R; ++m) {
        const int jo_ = full ? m : min(jbase + m, j_hi - 1) - min(jbase, j_hi - 1);
        const double2 cp01 = __ldg(cpb + 2 * jo_), cp23 = __ldg(cpb + 2 * jo_ + 1);
        const double b1 = wp.col1_zero ? rp01.y : rp01.y + cp01.y;
        gs[m] = rp23.x + cp23.x;
        cellK1[m] = locate_uniform<true, true>(b1, n1, tK1[m]);
        best[m] = __longlong_as_double(0x7ff0000000000000LL);
        arg[m] = 0;
        if (m) chain_ok = chain_ok && (cellK1[m] == cellK1[0] + m);
    }
    chain_ok = __all_sync(0xffffffffu, chain_ok);

    __syncthreads();          // org[] and the mbarrier are visible
    mbar_wait(&mbar, 0);

    for (int c = 0; c < C; ++c) {
        const double rc = __ldg(rr + c);
        double t0;
        const int cell0 = locate_uniform<true, false>(base0 + __ldg(Tc0 + c), n0, t0);
        const double *__restrict__ Wb = ring + (size_t)c * wp.buf_doubles - (org[c][1] * W0 + org[c][0]);
        if (chain_ok) {
            const double *p = Wb + (cellK1[0] * W0 + cell0);
            double a[R + 1];
#pragma unroll
            for (int k = 0; k <= R; ++k) {
                const double lo = p[k * W0], hi = p[k * W0 + 1];
                a[k] = fma(t0, hi - lo, lo);
            }
#pragma unroll
            for (int m = 0; m < R; ++m) {
                const double v = fma(tK1[m], a[m + 1] - a[m], a[m]);
                const double tot = (gs[m] + rc) + v;
                if (tot < best[m]) { best[m] = tot; arg[m] = c; }
            }
        } else {
#pragma unroll
            for (int m = 0; m < R; ++m) {
                const double *p = Wb + (cellK1[m] * W0 + cell0);
                const double v00 = p[0], v10 = p[1], v01 = p[W0], v11 = p[W0 + 1];
                const double a = fma(t0, v10 - v00, v00);
                const double b = fma(t0, v11 - v01, v01);
                const double v = fma(tK1[m], b - a, a);
                const double tot = (gs[m] + rc) + v;
                if (tot < best[m]) { best[m] = tot; arg[m] = c; }
            }
        }
    }

    if (i_lo + lane < i_hi) {
        double *jo = sp.J_out + (size_t)prob * sp.S_ext + (long long)(i - d0.ext_lo) * d0.stride +
                     (long long)(jbase - d1.ext_lo) * d1.stride;
        long long io = (long long)prob * sp.S_own + (long long)(i - d0.own_lo) +
                      (long long)(jbase - d1.own_lo) * d0.own_n;
        const long long sj = d1.stride;
        const int si = d0.own_n;
#pragma unroll
        for (int m = 0; m < R; ++m) {
            if (full || jbase + m < j_hi) {
                *jo = best[m];
                idx_store(sp.idx_out, sp.idx_bytes, io, arg[m]);
            }
            jo += sj;
            io += si;
        }
        if (sp.n_peers) {
#pragma unroll
            for (int m = 0; m < R; ++m)
                if (jbase + m < j_hi) { const int gi[2] = {i, jbase + m}; peer_store<2>(sp, prob, gi, best[m]); }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// k_stage_strip: the CHAIN problems (dimension-0 query = row part + control, dimension-1 query
// independent of the control: Solver_attitude, 3 torque levels) with the loops turned inside out.
// A warp owns 32 rows x strip_r columns and WALKS the columns: everything that depends on the row
// only — the dimension-0 cell and weight of every control, the window base pointers, r[c] — is
// computed once per thread and kept in registers; per column a thread does one 16-byte table load,
// one locate, and per control two shared-memory loads and the three lerps of the generic path (the
// dimension-0 lerp of the lower column is the upper column's of the previous step whenever the cells
// are consecutive, which they are except at the clamped grid edge).  No per-state arrays, so the
// strip length costs no registers and the per-thread prologue is amortised over strip_r x C updates
// (k_stage_chain: 54 instructions per update, half of them prologue/epilogue — profiles/r01).
// Same operations on the same operands as the generic path ⇒ bit-identical results.
// All C windows of the tile (32+ rows x NW*strip_r+ columns) are in flight from the first instruction.
// ---------------------------------------------------------------------------------------------
// IDX32: the argmin is stored as int32 (the default); false = 1- or 2-byte storage, decided at run time.
// A template parameter because the store sits in the branch-free column loop.
template <int NW, int CC, int OCC, bool PEER, bool IDX32>
__global__ void __launch_bounds__(NW * 32, OCC)
k_stage_strip(const __grid_constant__ StageParams sp, const __grid_constant__ WindowParams wp,
              const __grid_constant__ CUtensorMap tmap) {
    extern __shared__ __align__(128) double ring[];    // CC windows, win0 x win1 doubles each
    __shared__ __align__(8) uint64_t mbar;
    __shared__ int org[CC];

    // grid = (fast tile index, slow tile index, problem): no division to decode the tile; every
    // element offset fits 32 bits (checked by the planner), so addresses cost one IMAD.WIDE each
    const int tid = threadIdx.x, lane = tid & 31, wrp = tid >> 5;
    const uint32_t prob = blockIdx.z;
    const int R = wp.strip_r, WT1 = NW * R;
    const int ti = wp.tj_fastest ? blockIdx.y : blockIdx.x;
    const int tj = wp.tj_fastest ? blockIdx.x : blockIdx.y;
    const DimParams &d0 = sp.dim[0], &d1 = sp.dim[1];
    const int n0 = d0.n, n1 = d1.n, W0 = wp.win0;
    const int i_lo = d0.own_lo + ti * WT0, i_hi = min(i_lo + WT0, d0.own_lo + d0.own_n);
    const int j_lo = d1.own_lo + tj * WT1, j_hi = min(j_lo + WT1, d1.own_lo + d1.own_n);
    const double *Tc0 = d0.Tc + prob * (uint32_t)CC;
    const double *rr = sp.r + prob * (uint32_t)CC;

    // lanes 0..C-1 of warp 0 each place and issue one control's window (their table loads overlap)
    if (wrp == 0) {
        if (lane == 0) {
            mbar_init(&mbar, (uint32_t)CC);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncwarp();
        if (lane < CC) {
            const int c = lane;
            const double *tm = wp.tmm + prob * (uint32_t)wp.tmm_stride;
            double lo0 = __ldg(tm + wp.tmm_off[0][0] + 2 * ti);
            if (d0.Tb) lo0 = lo0 + __ldg(tm + wp.tmm_off[0][1] + 2 * ti);
            double lo1 = __ldg(tm + wp.tmm_off[1][0] + 2 * (d1.src_a == 0 ? ti : tj));
            if (d1.Tb) lo1 = lo1 + __ldg(tm + wp.tmm_off[1][1] + 2 * (d1.src_b == 0 ? ti : tj));
            const int c0 = cell_uniform(lo1, n1);
            int r0 = cell_uniform(lo0 + __ldg(Tc0 + c), n0);
            r0 -= (r0 - d0.ext_lo) & 1;                              // TMA: even innermost coordinate
            org[c] = c0 * W0 + r0;
            mbar_expect_tx(&mbar, (uint32_t)(wp.win0 * wp.win1) * 8u);
            tma_load_3d(ring + (uint32_t)(c * wp.buf_doubles), &tmap, &mbar, r0 - d0.ext_lo, c0 - d1.ext_lo, (int)prob);
        }
    }

    // row constants of this thread
    const int i = min(i_lo + lane, i_hi - 1);
    const double2 *rpk = reinterpret_cast<const double2 *>(wp.rowpack + (prob * (uint32_t)n0 + (uint32_t)i));
    const double2 rp01 = __ldg(rpk);
    const double qrow = __ldg(reinterpret_cast<const double *>(rpk + 1));
    double t0[CC], rc[CC];
    int cell0[CC];
#pragma unroll
    for (int c = 0; c < CC; ++c) {
        rc[c] = __ldg(rr + c);
        cell0[c] = locate_uniform<true, false>(rp01.x + __ldg(Tc0 + c), n0, t0[c]);
    }
    const int jb = j_lo + wrp * R;
    const int jcnt = min(R, j_hi - jb);                  // warp-uniform; <= 0 for a ragged last tile
    const double2 *cq = wp.colq + (prob * (uint32_t)n1 + (uint32_t)jb);
    double2 cd = jcnt > 0 ? __ldg(cq) : make_double2(0.0, 0.0);
    constexpr int PF = OCC == 5 ? 2 : 1;                 // column-table prefetch distance (2 only where registers allow)
    constexpr bool MAGIC = OCC == 5;                     // experiment: dimension-1 locate without conversion instructions
    auto locate1 = [&](double g, double &t) -> int {
        return MAGIC ? locate_magic_clamp(g, n1, t) : locate_uniform<true, true>(g, n1, t);
    };
    double2 cd1 = PF == 2 && jcnt > 0 ? __ldg(cq + 1) : make_double2(0.0, 0.0);
    const bool row_ok = i_lo + lane < i_hi;
    const uint32_t sj = (uint32_t)d1.stride, si = (uint32_t)d0.own_n;
    uint32_t jo = prob * (uint32_t)sp.S_ext + (uint32_t)(i - d0.ext_lo) * (uint32_t)d0.stride + (uint32_t)(jb - d1.ext_lo) * sj;
    // argmin store: byte offset into idx_out (1, 2 or 4 bytes per element, decided once per thread)
    const uint32_t ibytes = (uint32_t)sp.idx_bytes;
    const bool i1 = ibytes == 1;
    char *const ibase = reinterpret_cast<char *>(sp.idx_out);
    uint32_t io = prob * (uint32_t)sp.S_own + (uint32_t)(i - d0.own_lo) + (uint32_t)(jb - d1.own_lo) * si;   // elements

    __syncthreads();          // org[] and the mbarrier are visible
    if (wrp == 0 && lane < CC) {
        const int c = lane;
        // warm L2 with the same control's window of the tile pf_dist CTAs ahead in launch order (about
        // one wave: the CTA that will take this one's place): its TMA then sees L2, not DRAM, latency
        if (wp.pf_dist) {
            const unsigned lin = blockIdx.x + gridDim.x * (blockIdx.y + gridDim.y * blockIdx.z) + (unsigned)wp.pf_dist;
            if (lin < gridDim.x * gridDim.y * gridDim.z) {
                const unsigned bx = lin % gridDim.x, rest = lin / gridDim.x;
                const unsigned by = rest % gridDim.y, pz = rest / gridDim.y;
                const int pti = wp.tj_fastest ? (int)by : (int)bx, ptj = wp.tj_fastest ? (int)bx : (int)by;
                const double *ptm = wp.tmm + pz * (uint32_t)wp.tmm_stride;
                double plo0 = __ldg(ptm + wp.tmm_off[0][0] + 2 * pti);
                if (d0.Tb) plo0 = plo0 + __ldg(ptm + wp.tmm_off[0][1] + 2 * pti);
                double plo1 = __ldg(ptm + wp.tmm_off[1][0] + 2 * (d1.src_a == 0 ? pti : ptj));
                if (d1.Tb) plo1 = plo1 + __ldg(ptm + wp.tmm_off[1][1] + 2 * (d1.src_b == 0 ? pti : ptj));
                const int pc0 = cell_uniform(plo1, n1);
                int pr0 = cell_uniform(plo0 + __ldg(d0.Tc + pz * (uint32_t)CC + c), n0);
                pr0 -= (pr0 - d0.ext_lo) & 1;
                tma_prefetch_3d(&tmap, pr0 - d0.ext_lo, pc0 - d1.ext_lo, (int)pz);
            }
        }
    }
    mbar_wait(&mbar, 0);

    if (jcnt <= 0) return;    // ragged last tile: this warp has no columns (no barrier follows)

    // shared-memory byte addresses (32-bit): window of control c shifted so that cell1 * pitch is
    // this row's lower-left corner.  `base` passes through a volatile asm placed after the wait, so
    // no window load can be scheduled above it.
    uint32_t base = smem_u32(ring);
    asm volatile("" : "+r"(base)::"memory");
    const uint32_t pitch = (uint32_t)W0 * 8u;
    uint32_t pc[CC];
#pragma unroll
    for (int c = 0; c < CC; ++c) pc[c] = base + 8u * (uint32_t)(c * wp.buf_doubles + (cell0[c] - org[c]));


    // dimension-0 lerp of one window column for every control
    auto column = [&](uint32_t o, double (&a)[CC]) {
        double lo[CC], hi[CC];
#pragma unroll
        for (int c = 0; c < CC; ++c) {
            lo[c] = lds_f64(pc[c] + o);
            hi[c] = lds_f64(pc[c] + o + 8);
        }
#pragma unroll
        for (int c = 0; c < CC; ++c) a[c] = fma(t0[c], hi[c] - lo[c], lo[c]);
    };
    // the lower column of the first state; after that the lower column of a state is the upper
    // column of the previous one whenever the cells are consecutive — always, except at the clamped
    // grid edge or when rounding makes a query skip a cell
    double ahi[CC];
    int prev;
    {
        double t1;
        prev = locate1(rp01.y + cd.x, t1) - 1;
        column((uint32_t)(prev + 1) * pitch, ahi);
    }
    auto finish = [&](int mm, double gs, double t1, const double (&alo)[CC], const double (&a)[CC]) {
        double best = __longlong_as_double(0x7ff0000000000000LL);
        int arg = 0;
#pragma unroll
        for (int c = 0; c < CC; ++c) {
            const double v = fma(t1, a[c] - alo[c], alo[c]);
            const double tot = (gs + rc[c]) + v;
            if (tot < best) { best = tot; arg = c; }
        }
        if (row_ok) {
            sp.J_out[jo] = best;
            if (IDX32) {
                sp.idx_out[io] = arg;
            } else {
                char *const ip = ibase + (size_t)io * ibytes;
                if (i1) *reinterpret_cast<unsigned char *>(ip) = (unsigned char)arg;
                else *reinterpret_cast<unsigned short *>(ip) = (unsigned short)arg;
            }
            if (PEER) { const int gi[2] = {i, jb + mm}; peer_store<2>(sp, (int)prob, gi, best); }
        }
        jo += sj;
        io += si;
    };
    // Chained columns.  `lower` holds the dimension-0 lerps of window column cell1 (computed by the
    // previous step), `upper` receives those of column cell1 + 1.  Whether every lane's cell really
    // was the successor of its previous one is only accumulated in `bad` — no branch between the
    // columns, so two of them interleave — and checked once per strip: a strip that fails (clamped
    // grid edge, a query that skipped a cell through rounding) or that is ragged is redone by the
    // generic loop below, which overwrites what the chained pass stored.
    int m = 0;
    if (jcnt == R) {
        bool bad = false;
        int expect = prev + 1;
        auto step = [&](double (&lower)[CC], double (&upper)[CC]) {
            const double2 cdn = __ldg(cq + m + PF);                     // a later column's tables (colq is padded)
            double t1;
            const int cell1 = locate1(rp01.y + cd.x, t1);
            bad = bad || cell1 != expect;
            column((uint32_t)cell1 * pitch + pitch, upper);
            finish(m, qrow + cd.y, t1, lower, upper);
            expect = cell1 + 1;
            if (PF == 2) { cd = cd1; cd1 = cdn; } else cd = cdn;
            ++m;
        };
        double bhi[CC];
#pragma unroll 1
        while (m < R) {             // strip_r is even
            step(ahi, bhi);
            step(bhi, ahi);         // registers ping-pong, no copies
        }
        if (!__any_sync(0xffffffffu, bad)) return;
        jo -= (uint32_t)R * sj;
        io -= (uint32_t)R * si;
        m = 0;
    }
    // generic columns: both window columns are read
#pragma unroll 1
    for (; m < jcnt; ++m) {
        const double2 cdc = __ldg(cq + m);
        double t1, alo[CC];
        const int cell1 = locate1(rp01.y + cdc.x, t1);
        column((uint32_t)cell1 * pitch, alo);
        column((uint32_t)cell1 * pitch + pitch, ahi);
        finish(m, qrow + cdc.y, t1, alo, ahi);
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
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess)
        return nullptr;
    fn = reinterpret_cast<PFN_encodeTiled>(p);
    return fn;
}

struct WindowState {
    WindowParams wp{};
    std::vector<CUtensorMap> maps;   // one per J slot
    size_t smem = 0;
    bool hc0 = false, hc1 = false, chain = false, lean = false, strip = false;
    size_t lean_smem = 0;
    int strip_nw = 4;
    void *d_colq = nullptr;
    int occ = 2, rstates = 8;
    WideTables *wide = nullptr;     // k_stage_wide: host copy of the constant-bank control tables (null = not used)
    int wide_ns = 2;                // ring slots
    bool strip_magic_ok = false;    // dimension-1 queries of the strip kernel stay below 2^30 cells
    int wide_xu = 2;                // BELLMAN_WIDE_XU=0|1|2: weights of that many dimensions through I2F (k_stage_wide's XU)
    bool wide_bar = false;          // BELLMAN_WIDE_BARRIER=1: a CTA barrier per chunk instead of the empty-slot mbarriers
    size_t wide_smem = 0;
    void *d_cmm = nullptr, *d_tmm = nullptr, *d_rowp = nullptr, *d_colp = nullptr;
};

void minmax_range(const double *v, int lo, int hi, double &mn, double &mx) {
    mn = std::numeric_limits<double>::infinity();
    mx = -mn;
    for (int k = lo; k < hi; ++k) { mn = std::min(mn, v[k]); mx = std::max(mx, v[k]); }
}

}  // namespace

// picks the kernel instantiation; with set_attr_only it just raises the dynamic shared memory limit
template <bool HC0, bool HC1, bool CHAIN, int BATCH, int OCC, int R = 8>
static bool window_go(const WindowState *ws, const StageParams *sp, const CUtensorMap *map, dim3 grid,
                      cudaStream_t st, bool set_attr_only) {
    if (HC0 && HC1 && BATCH == 4 && OCC == 2 && R == 8) {   // the long-control-loop kernel: pitch 48 specialisation
        auto fn48 = k_stage_window<HC0, HC1, CHAIN, BATCH, OCC, R, 48, 2>;
        if (set_attr_only) {
            if (!raise_smem_limit((const void *)fn48, ws->smem)) return false;
        } else if (ws->wp.win0 == 48) {
            fn48<<<grid, WNT, ws->smem, st>>>(*sp, ws->wp, *map);
            return true;
        }
    }
    auto fn = k_stage_window<HC0, HC1, CHAIN, BATCH, OCC, R>;
    if (set_attr_only)
        return raise_smem_limit((const void *)fn, ws->smem);
    fn<<<grid, WNT, ws->smem, st>>>(*sp, ws->wp, *map);
    return true;
}
// (round 1 also instantiated BATCH 2 / 8, one CTA per SM, the conversion-pipe locate variants and a 4-states-per-
// thread form of the long-control-loop kernel as experiment knobs; none beat this configuration — DESIGN.md section 4 —
// and they were dropped to keep the build short)
template <bool HC0, bool HC1, bool CHAIN>
static bool window_go_hc(const WindowState *ws, const StageParams *sp, const CUtensorMap *map, dim3 grid,
                         cudaStream_t st, bool sa) {
    if (CHAIN && ws->rstates == 4) return window_go<HC0, HC1, CHAIN, 4, 4, 4>(ws, sp, map, grid, st, sa);
    return window_go<HC0, HC1, CHAIN, 4, 2>(ws, sp, map, grid, st, sa);
}
static bool window_dispatch(const WindowState *ws, const StageParams *sp, const CUtensorMap *map,
                            const void *, dim3 grid, cudaStream_t st, bool sa) {
    if (ws->hc0 && ws->hc1) return window_go_hc<true, true, false>(ws, sp, map, grid, st, sa);
    if (ws->hc0 && ws->chain) return window_go_hc<true, false, true>(ws, sp, map, grid, st, sa);
    if (ws->hc0) return window_go_hc<true, false, false>(ws, sp, map, grid, st, sa);
    return window_go_hc<false, true, false>(ws, sp, map, grid, st, sa);
}

template <int NW, int CC, int OCC>
static bool strip_go(const WindowState *ws, const StageParams *sp, const CUtensorMap *map, dim3 grid, cudaStream_t st,
                     bool set_attr_only) {
    auto fn = k_stage_strip<NW, CC, OCC, false, true>;
    auto fnp = k_stage_strip<NW, CC, OCC, true, true>;      // multi-GPU: halo states also go to the neighbours
    auto fnn = k_stage_strip<NW, CC, OCC, false, false>;    // 1- / 2-byte argmin storage
    auto fnpn = k_stage_strip<NW, CC, OCC, true, false>;
    if (set_attr_only)
        return raise_smem_limit((const void *)fn, ws->lean_smem) && raise_smem_limit((const void *)fnp, ws->lean_smem) &&
               raise_smem_limit((const void *)fnn, ws->lean_smem) && raise_smem_limit((const void *)fnpn, ws->lean_smem);
    if (sp->idx_bytes == 4) {
        if (sp->n_peers) fnp<<<grid, NW * 32, ws->lean_smem, st>>>(*sp, ws->wp, *map);
        else fn<<<grid, NW * 32, ws->lean_smem, st>>>(*sp, ws->wp, *map);
    } else {
        if (sp->n_peers) fnpn<<<grid, NW * 32, ws->lean_smem, st>>>(*sp, ws->wp, *map);
        else fnn<<<grid, NW * 32, ws->lean_smem, st>>>(*sp, ws->wp, *map);
    }
    return true;
}
static bool strip_dispatch(const WindowState *ws, const StageParams *sp, const CUtensorMap *map, dim3 grid,
                           cudaStream_t st, bool sa) {
    const int C = ws->wp.nchunks;   // chunk size 1: one window per control
    if (ws->strip_nw == 8) {
        switch (C) {
            case 1: return strip_go<8, 1, 3>(ws, sp, map, grid, st, sa);
            case 2: return strip_go<8, 2, 3>(ws, sp, map, grid, st, sa);
            case 3: return strip_go<8, 3, 3>(ws, sp, map, grid, st, sa);
            default: return strip_go<8, 4, 3>(ws, sp, map, grid, st, sa);
        }
    }
    if (ws->strip_nw == 2) {   // 64-thread CTAs: half the window per CTA, twice the CTAs per SM
        switch (C) {
            case 1: return strip_go<2, 1, 14>(ws, sp, map, grid, st, sa);
            case 2: return strip_go<2, 2, 14>(ws, sp, map, grid, st, sa);
            case 3: return strip_go<2, 3, 14>(ws, sp, map, grid, st, sa);
            default: return strip_go<2, 4, 14>(ws, sp, map, grid, st, sa);
        }
    }
    if (ws->occ == 3 && ws->strip_magic_ok) {   // BELLMAN_WIN_OCC=3: fewer CTAs per SM, more registers (experiments)
        switch (C) {
            case 3: return strip_go<4, 3, 5>(ws, sp, map, grid, st, sa);
            default: break;
        }
    }
    switch (C) {
        case 1: return strip_go<4, 1, 7>(ws, sp, map, grid, st, sa);
        case 2: return strip_go<4, 2, 7>(ws, sp, map, grid, st, sa);
        case 3: return strip_go<4, 3, 7>(ws, sp, map, grid, st, sa);
        default: return strip_go<4, 4, 7>(ws, sp, map, grid, st, sa);
    }
}
template <bool BAR, int XU = 0>
static bool wide_go(const WindowState *ws, const StageParams *sp, const CUtensorMap *map, dim3 grid, cudaStream_t st,
                    bool set_attr_only, const WindowParams *wpo = nullptr) {
    const WindowParams &wpl = wpo ? *wpo : ws->wp;
    auto f48 = k_stage_wide<2, 48, true, BAR, XU>, f48n = k_stage_wide<2, 48, false, BAR, XU>;
    auto f0 = k_stage_wide<2, 0, true, BAR, XU>, f0n = k_stage_wide<2, 0, false, BAR, XU>;
    if (set_attr_only) {
        for (const void *f : {(const void *)f48, (const void *)f48n, (const void *)f0, (const void *)f0n})
            if (!raise_smem_limit(f, ws->wide_smem)) return false;
        return true;
    }
    const bool i32 = sp->idx_bytes == 4;
    if (wpl.win0 == 48) (i32 ? f48 : f48n)<<<grid, WIDE_NT, ws->wide_smem, st>>>(*sp, wpl, *map, *ws->wide);
    else (i32 ? f0 : f0n)<<<grid, WIDE_NT, ws->wide_smem, st>>>(*sp, wpl, *map, *ws->wide);
    return true;
}
static bool wide_dispatch(const WindowState *ws, const StageParams *sp, const CUtensorMap *map, dim3 grid,
                          cudaStream_t st, bool sa, const WindowParams *wpo = nullptr) {
    if (ws->wide_bar) return wide_go<true>(ws, sp, map, grid, st, sa, wpo);
    if (ws->wide_xu == 1) return wide_go<false, 1>(ws, sp, map, grid, st, sa, wpo);
    if (ws->wide_xu == 2) return wide_go<false, 2>(ws, sp, map, grid, st, sa, wpo);
    return wide_go<false>(ws, sp, map, grid, st, sa, wpo);
}
static void window_teardown_state(WindowState *ws) {
    cudaFree(ws->d_cmm); cudaFree(ws->d_tmm); cudaFree(ws->d_rowp); cudaFree(ws->d_colp); cudaFree(ws->d_colq);
    delete ws->wide;
    delete ws;
}

// Exact worst-case window extents for a given chunk size: replays, on the host, the bound the
// kernel uses ((min Ta + min Tb) + min Tc .. (max Ta + max Tb) + max Tc, located with the same
// rule) for every tile and chunk.
static void window_extents(const bellman_handle *h, int cchunk, int wt1, int &w0, int &w1) {
    const HostProblem &hp = h->hp;
    const int nch = (hp.C + cchunk - 1) / cchunk;
    w0 = w1 = 0;
    for (int p = 0; p < hp.P; ++p) {
        for (int d = 0; d < 2; ++d) {
            const int sa = hp.src_a[d], sb = hp.src_b[d];
            // per tile-index min/max of Ta and Tb along the dimension that indexes them
            auto tile_mm = [&](const std::vector<double> &tab, int src, std::vector<double> &mn,
                               std::vector<double> &mx) {
                const int T = src == 0 ? WT0 : wt1;
                const int lo0 = h->own_lo[src], cnt = h->own_n[src];
                const int nt = (cnt + T - 1) / T;
                mn.resize(nt); mx.resize(nt);
                for (int t = 0; t < nt; ++t)
                    minmax_range(tab.data() + (size_t)p * hp.n[src], lo0 + t * T,
                                 std::min(lo0 + (t + 1) * T, lo0 + cnt), mn[t], mx[t]);
            };
            std::vector<double> amn, amx, bmn{0.0}, bmx{0.0}, cmn(nch, 0.0), cmx(nch, 0.0);
            tile_mm(hp.Ta[d], sa, amn, amx);
            if (hp.has_b[d]) tile_mm(hp.Tb[d], sb, bmn, bmx);
            if (hp.has_c[d])
                for (int c = 0; c < nch; ++c)
                    minmax_range(hp.Tc[d].data() + (size_t)p * hp.C, c * cchunk,
                                 std::min(hp.C, (c + 1) * cchunk), cmn[c], cmx[c]);
            // if Ta and Tb are indexed by the same dimension their tile indices coincide
            const bool same = hp.has_b[d] && sa == sb;
            int ext = 0;
            for (size_t ta = 0; ta < amn.size(); ++ta)
                for (size_t tb = same ? ta : 0; tb < (same ? ta + 1 : bmn.size()); ++tb)
                    for (int c = 0; c < nch; ++c) {
                        double lo = amn[ta], hi = amx[ta];
                        if (hp.has_b[d]) { lo = lo + bmn[tb]; hi = hi + bmx[tb]; }
                        if (hp.has_c[d]) { lo = lo + cmn[c]; hi = hi + cmx[c]; }
                        ext = std::max(ext, host_locate(hp, p, d, hi) + 2 - host_locate(hp, p, d, lo));
                    }
            (d == 0 ? w0 : w1) = std::max(d == 0 ? w0 : w1, ext);
        }
    }
}

// bytes of one ring slot, rounded up to the 128-byte alignment TMA needs for its destination
static size_t slot_bytes(int w0, int w1) { return ((size_t)w0 * w1 * 8 + 127) / 128 * 128; }

void window_setup(bellman_handle *h) {
    h->wcfg.valid = false;
    const HostProblem &hp = h->hp;
    if (hp.D != 2) return;
    for (int32_t m : hp.mode)
        if (m != BELLMAN_LOCATE_UNIFORM) return;
    if (h->ld0 % 2) return;                      // TMA global strides must be multiples of 16 bytes
    if (((size_t)h->S_ext * 8) % 16) return;
    PFN_encodeTiled enc = get_encode();
    if (!enc) return;

    // small-control CHAIN problems run 4 states per thread (see k_stage_window)
    const bool chain_cfg = hp.has_c[0] && !hp.has_c[1] && hp.src_a[0] == 0 && (!hp.has_b[0] || hp.src_b[0] == 0) &&
                           !std::getenv("BELLMAN_WIN_NOCHAIN");
    const int rstates = (chain_cfg && hp.C <= 8 && !std::getenv("BELLMAN_WIN_R8")) ? 4 : 8;
    // k_stage_strip geometry: NW warps per CTA, each walking strip_r columns (tile 32 x NW*strip_r)
    const bool lean_cfg = chain_cfg && rstates == 4 && hp.C <= 4 && !std::getenv("BELLMAN_WIN_NOLEAN");
    const bool strip_cfg = lean_cfg && !std::getenv("BELLMAN_WIN_NOSTRIP");
    int strip_nw = 4, strip_r = 8;
    if (const char *e = std::getenv("BELLMAN_STRIP_NW")) strip_nw = std::atoi(e) == 8 ? 8 : std::atoi(e) == 2 ? 2 : 4;
    if (const char *e = std::getenv("BELLMAN_STRIP_R")) strip_r = std::max(2, std::min(64, std::atoi(e) / 2 * 2));
    else if ((long long)((h->own_n[0] + WT0 - 1) / WT0) * ((h->own_n[1] + strip_nw * 16 - 1) / (strip_nw * 16)) * hp.P >=
             32LL * 148 * 3)
        strip_r = 16;   // many waves even with 32 x 64 tiles: longer strips amortise the per-thread prologue (measured +2 %)
    const int wt1 = strip_cfg ? strip_nw * strip_r : tile1_of(rstates);

    // k_stage_wide (one 512-thread CTA per SM) takes the long-control-loop problems whose control tables fit
    // its constant-bank parameter; it may use the whole shared memory of an SM for its two ring slots
    const bool wide_cfg = hp.has_c[0] && hp.has_c[1] && !chain_cfg && rstates == 8 &&
                          (long long)hp.P * hp.C <= WIDE_MAXC && !std::getenv("BELLMAN_NO_WIDE");
    // (chunks of 1..3 controls are not offered to it: its per-CTA chunk table holds WIDE_MAXCH entries)
    // pick the chunk size: most updates per staged byte among configs that keep two CTAs per SM
    const size_t budget2 = wide_cfg ? 220 * 1024 : 110 * 1024, budget1 = 220 * 1024;
    int best_cc = 0, best_w0 = 0, best_w1 = 0;
    double best_score = -1.0;
    std::vector<int> cands;
    if (lean_cfg) {
        cands.push_back(1);                       // k_stage_chain: one window per control, all in flight
    } else {
        for (int cc : {1, 2, 3, 4, 6, 8, 12, 16, 20, 24, 26, 28, 32})
            if (cc <= hp.C) cands.push_back(cc);
        if (hp.C <= 32 && std::find(cands.begin(), cands.end(), hp.C) == cands.end()) cands.push_back(hp.C);
        if (const char *e = std::getenv("BELLMAN_WIN_CC")) { cands.clear(); cands.push_back(std::max(1, std::min(hp.C, std::atoi(e)))); }
    }
    for (int cc : cands) {
        int w0, w1;
        window_extents(h, cc, wt1, w0, w1);
        // +1: the window origin is rounded down to an even row.  Pitch: a multiple of 16 doubles keeps
        // the generic gathers conflict-free when a warp straddles columns; the CHAIN kernel reads
        // whole columns per warp, so any even pitch is conflict-free and the window can be tight
        w0 = lean_cfg ? (w0 + 1 + 1) / 2 * 2 : (w0 + 1 + 15) / 16 * 16;
        if (w0 > 256) continue;
        const int boxes = (w1 + 255) / 256;
        const int box1 = (w1 + boxes - 1) / boxes;
        w1 = boxes * box1;
        const size_t bytes = 2 * slot_bytes(w0, w1);
        if (bytes > budget1) continue;
        const int nch = (hp.C + cc - 1) / cc;
        double score = (double)std::min(cc, hp.C) / ((double)w0 * w1);
        if (bytes > budget2) score *= 0.5;       // one CTA per SM only
        if (nch == 1) score *= 1.0;
        if (score > best_score) { best_score = score; best_cc = cc; best_w0 = w0; best_w1 = w1; }
    }
    if (best_cc == 0) return;

    auto *ws = new WindowState();
    WindowParams &wp = ws->wp;
    wp.win0 = best_w0;
    wp.boxes = (best_w1 + 255) / 256;
    wp.box1 = best_w1 / wp.boxes;
    wp.win1 = best_w1;
    wp.cchunk = best_cc;
    wp.nchunks = (hp.C + best_cc - 1) / best_cc;
    wp.ntile0 = (h->own_n[0] + WT0 - 1) / WT0;
    wp.ntile1 = (h->own_n[1] + wt1 - 1) / wt1;
    ws->smem = 2 * slot_bytes(wp.win0, wp.win1);
    wp.buf_doubles = (int)(slot_bytes(wp.win0, wp.win1) / 8);
    {   // fastest tile index = the dimension the control grid sweeps furthest (in cells)
        double sweep[2] = {0.0, 0.0};
        for (int d = 0; d < 2; ++d)
            if (hp.has_c[d]) {
                double mn, mx;
                minmax_range(hp.Tc[d].data(), 0, hp.C, mn, mx);
                sweep[d] = mx - mn;   // already in cells
            }
        wp.tj_fastest = sweep[1] >= sweep[0] ? 1 : 0;
    }
    ws->hc0 = hp.has_c[0];
    ws->hc1 = hp.has_c[1];
    // dimension-0 query independent of the dimension-1 index, dimension 1 independent of the control
    ws->chain = chain_cfg;
    ws->rstates = rstates;
    if (hp.q_order[0] != 0 && hp.q_order[0] != 1) { window_teardown_state(ws); return; }

    // per-chunk control min/max and interleaved (grid, rinv) tables
    std::vector<double> cmm((size_t)hp.P * wp.nchunks * 4, 0.0);
    for (int p = 0; p < hp.P; ++p)
        for (int c = 0; c < wp.nchunks; ++c)
            for (int d = 0; d < 2; ++d)
                if (hp.has_c[d])
                    minmax_range(hp.Tc[d].data() + (size_t)p * hp.C, c * wp.cchunk,
                                 std::min(hp.C, (c + 1) * wp.cchunk), cmm[((size_t)p * wp.nchunks + c) * 4 + 2 * d],
                                 cmm[((size_t)p * wp.nchunks + c) * 4 + 2 * d + 1]);
    auto upload = [&](const std::vector<double> &v, void **dptr) {
        return cudaMalloc(dptr, v.size() * sizeof(double)) == cudaSuccess &&
               cudaMemcpy(*dptr, v.data(), v.size() * sizeof(double), cudaMemcpyHostToDevice) == cudaSuccess;
    };
    if (!upload(cmm, &ws->d_cmm)) { window_teardown_state(ws); return; }
    wp.cmm = static_cast<const double *>(ws->d_cmm);

    // canonical packed state tables
    {
        std::vector<double> rowp((size_t)hp.P * hp.n[0] * 4, 0.0), colp((size_t)hp.P * hp.n[1] * 4, 0.0);
        bool colz[2] = {true, true};
        for (int p = 0; p < hp.P; ++p) {
            for (int d = 0; d < 2; ++d) {
                const double *ta = hp.Ta[d].data() + (size_t)p * hp.n[hp.src_a[d]];
                const double *tb = hp.has_b[d] ? hp.Tb[d].data() + (size_t)p * hp.n[hp.src_b[d]] : nullptr;
                const int sa = hp.src_a[d], sb = hp.has_b[d] ? hp.src_b[d] : -1;
                for (int i = 0; i < hp.n[0]; ++i) {
                    double v = 0.0;
                    if (sa == 0 && sb == 0) v = ta[i] + tb[i];
                    else if (sa == 0) v = ta[i];
                    else if (sb == 0) v = tb[i];
                    rowp[((size_t)p * hp.n[0] + i) * 4 + d] = v;
                }
                if (sa == 1 || sb == 1) colz[d] = false;
                for (int j = 0; j < hp.n[1]; ++j) {
                    double v = 0.0;
                    if (sa == 1 && sb == 1) v = ta[j] + tb[j];
                    else if (sa == 1) v = ta[j];
                    else if (sb == 1) v = tb[j];
                    colp[((size_t)p * hp.n[1] + j) * 4 + d] = v;
                }
            }
            for (int i = 0; i < hp.n[0]; ++i) rowp[((size_t)p * hp.n[0] + i) * 4 + 2] = hp.q[0][(size_t)p * hp.n[0] + i];
            for (int j = 0; j < hp.n[1]; ++j) colp[((size_t)p * hp.n[1] + j) * 4 + 2] = hp.q[1][(size_t)p * hp.n[1] + j];
        }
        // a dimension with neither table indexed by the row still needs the row part to be a true
        // zero that is skipped, otherwise -0.0 + 0.0 would flip a sign; rows are always added, so
        // require the row part to exist (true for every reference class)
        for (int d = 0; d < 2; ++d)
            if (hp.src_a[d] != 0 && !(hp.has_b[d] && hp.src_b[d] == 0)) { window_teardown_state(ws); return; }
        if (!upload(rowp, &ws->d_rowp) || !upload(colp, &ws->d_colp)) { window_teardown_state(ws); return; }
        wp.rowpack = static_cast<const double4 *>(ws->d_rowp);
        wp.colpack = static_cast<const double4 *>(ws->d_colp);
        wp.col0_zero = colz[0] ? 1 : 0;
        wp.col1_zero = colz[1] ? 1 : 0;
        std::vector<double> colq((size_t)hp.P * hp.n[1] * 2 + 4, 0.0);   // +2 entries: the strip kernel prefetches up to two ahead
        for (size_t k = 0; k < (size_t)hp.P * hp.n[1]; ++k) { colq[2 * k] = colp[4 * k + 1]; colq[2 * k + 1] = colp[4 * k + 2]; }
        if (!upload(colq, &ws->d_colq)) { window_teardown_state(ws); return; }
        // the strip kernel's conversion-free locate (5-CTA variant) needs dimension-1 queries below 2^30 cells
        double q1 = 0.0, q1r = 0.0;
        for (size_t k = 0; k < (size_t)hp.P * hp.n[1]; ++k) q1 = std::max(q1, std::fabs(colp[4 * k + 1]));
        for (size_t k = 0; k < (size_t)hp.P * hp.n[0]; ++k) q1r = std::max(q1r, std::fabs(rowp[4 * k + 1]));
        ws->strip_magic_ok = q1 + q1r < 1073741824.0;
        wp.colq = static_cast<const double2 *>(ws->d_colq);
        wp.strip_r = strip_r;
        wp.pf_dist = std::getenv("BELLMAN_STRIP_PF") ? std::atoi(std::getenv("BELLMAN_STRIP_PF")) : 148 * 7;   // one wave of 7 CTAs per SM
        // only grids of many waves: on small (L2-resident) grids the prefetch is pure extra TMA traffic
        if (!std::getenv("BELLMAN_STRIP_PF") &&
            (long long)((h->own_n[0] + WT0 - 1) / WT0) * ((h->own_n[1] + wt1 - 1) / wt1) * hp.P < 4LL * 148 * 7)
            wp.pf_dist = 0;
    }

    // per-tile-index extrema of the state-indexed tables (read by the kernel instead of reducing per tile)
    {
        int off = 0;
        for (int d = 0; d < 2; ++d)
            for (int ab = 0; ab < 2; ++ab) {
                wp.tmm_off[d][ab] = off;
                if (ab == 1 && !hp.has_b[d]) continue;
                const int src = ab == 0 ? hp.src_a[d] : hp.src_b[d];
                off += 2 * (src == 0 ? wp.ntile0 : wp.ntile1);
            }
        wp.tmm_stride = off;
        std::vector<double> tmm((size_t)hp.P * off, 0.0);
        for (int p = 0; p < hp.P; ++p)
            for (int d = 0; d < 2; ++d)
                for (int ab = 0; ab < 2; ++ab) {
                    if (ab == 1 && !hp.has_b[d]) continue;
                    const int src = ab == 0 ? hp.src_a[d] : hp.src_b[d];
                    const std::vector<double> &tab = ab == 0 ? hp.Ta[d] : hp.Tb[d];
                    const int T = src == 0 ? WT0 : wt1, nt = src == 0 ? wp.ntile0 : wp.ntile1;
                    const int lo0 = h->own_lo[src], cnt = h->own_n[src];
                    for (int t = 0; t < nt; ++t)
                        minmax_range(tab.data() + (size_t)p * hp.n[src], lo0 + t * T, std::min(lo0 + (t + 1) * T, lo0 + cnt),
                                     tmm[(size_t)p * off + wp.tmm_off[d][ab] + 2 * t],
                                     tmm[(size_t)p * off + wp.tmm_off[d][ab] + 2 * t + 1]);
                }
        if (!upload(tmm, &ws->d_tmm)) { window_teardown_state(ws); return; }
        wp.tmm = static_cast<const double *>(ws->d_tmm);
    }

    // one tensor map per J slot: [P][ext_n1][ext_n0] fp64, box = win0 x box1 x 1
    const int nslots = h->store_J_all ? hp.N : 2;
    ws->maps.resize(nslots);
    for (int s = 0; s < nslots; ++s) {
        cuuint64_t gdim[3] = {(cuuint64_t)h->ext_n[0], (cuuint64_t)h->ext_n[1], (cuuint64_t)hp.P};
        cuuint64_t gstr[2] = {(cuuint64_t)h->ld0 * 8, (cuuint64_t)h->S_ext * 8};
        cuuint32_t box[3] = {(cuuint32_t)wp.win0, (cuuint32_t)wp.box1, 1};
        cuuint32_t estr[3] = {1, 1, 1};
        void *base = h->d_J + (size_t)s * h->slot_elems_J();
        CUresult r = enc(&ws->maps[s], CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, base, gdim, gstr, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                         CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { window_teardown_state(ws); return; }
    }
    {
        const char *eo = std::getenv("BELLMAN_WIN_OCC");      // 3: the strip kernel's 5-CTA register budget
        ws->occ = eo ? std::atoi(eo) : 2;
        if (ws->occ < 1 || ws->occ > 4) ws->occ = 2;
    }
    if (!window_dispatch(ws, nullptr, nullptr, nullptr, dim3(), nullptr, true)) { window_teardown_state(ws); return; }
    ws->strip = strip_cfg && wp.cchunk == 1 && wp.nchunks <= 4 && wp.boxes == 1 && !wp.col1_zero &&
                (uint64_t)hp.P * (uint64_t)h->S_ext < (1ull << 31) && wp.ntile0 <= 65535 && wp.ntile1 <= 65535 &&
                hp.P <= 65535;
    if (ws->strip) {
        ws->strip_nw = strip_nw;
        ws->lean_smem = (size_t)wp.nchunks * slot_bytes(wp.win0, wp.win1);
        if (ws->lean_smem > 200 * 1024 || !strip_dispatch(ws, nullptr, nullptr, dim3(), nullptr, true)) ws->strip = false;
        if (!ws->strip && wt1 != tile1_of(rstates)) { window_teardown_state(ws); return; }
    }
    ws->lean = !ws->strip && lean_cfg && wp.cchunk == 1 && wp.nchunks <= 4 && wp.boxes == 1;
    if (ws->lean) {
        ws->lean_smem = (size_t)wp.nchunks * slot_bytes(wp.win0, wp.win1);
        if (ws->lean_smem > 56 * 1024 ||
            !raise_smem_limit((const void *)k_stage_chain<4, 4>, ws->lean_smem))
            ws->lean = false;
    }
    if (!ws->hc0 && !ws->hc1) { window_teardown_state(ws); return; }   // no control dependence at all: nothing to stage for
    // k_stage_wide takes the long-control-loop problems whose control tables fit its constant-bank parameter
    if (wide_cfg && !ws->strip && !ws->lean && wp.nchunks <= WIDE_MAXCH) {
        ws->wide_bar = std::getenv("BELLMAN_WIDE_BARRIER") != nullptr;
        if (const char *e = std::getenv("BELLMAN_WIDE_XU")) ws->wide_xu = std::max(0, std::min(2, std::atoi(e)));
        ws->wide_smem = (size_t)ws->wide_ns * slot_bytes(wp.win0, wp.win1);
        if (ws->wide_smem <= 225 * 1024) {
            ws->wide = new WideTables();
            std::memset(ws->wide, 0, sizeof(WideTables));
            for (size_t k = 0; k < (size_t)hp.P * hp.C; ++k) {
                ws->wide->tc0[k] = hp.Tc[0][k];
                ws->wide->tc1[k] = hp.Tc[1][k];
                ws->wide->r[k] = hp.r[k];
            }
            if (!wide_dispatch(ws, nullptr, nullptr, dim3(), nullptr, true)) { delete ws->wide; ws->wide = nullptr; }
        }
    }
    h->wstate = ws;
    h->wcfg.tile0 = WT0; h->wcfg.tile1 = wt1; h->wcfg.cchunk = wp.cchunk;
    h->wcfg.win0 = wp.win0; h->wcfg.win1 = wp.win1;
    h->wcfg.valid = true;
}

void window_teardown(bellman_handle *h) {
    auto *ws = static_cast<WindowState *>(h->wstate);
    if (!ws) return;
    window_teardown_state(ws);
    h->wstate = nullptr;
}

const char *window_variant(const bellman_handle *h) {
    auto *ws = static_cast<const WindowState *>(h->wstate);
    if (!ws) return "window";
    return ws->strip ? "window:strip" : ws->lean ? "window:chain" : ws->wide ? "window:wide" :
           ws->chain ? "window:ring-chain" : "window:ring";
}

cudaError_t window_launch_for_handle(bellman_handle *h, const StageParams &sp, int slot_next, cudaStream_t st) {
    auto *ws = static_cast<WindowState *>(h->wstate);
    if (!ws || !h->wcfg.valid) return cudaErrorNotSupported;
    const WindowParams &wp = ws->wp;
    const dim3 grid((unsigned)(wp.ntile0 * wp.ntile1), (unsigned)sp.P);
    const CUtensorMap &map = ws->maps[slot_next];
    if (ws->strip) {
        const dim3 g3(wp.tj_fastest ? wp.ntile1 : wp.ntile0, wp.tj_fastest ? wp.ntile0 : wp.ntile1, (unsigned)sp.P);
        strip_dispatch(ws, &sp, &map, g3, st, false);
    }
    else if (ws->lean) k_stage_chain<4, 4><<<grid, WNT, ws->lean_smem, st>>>(sp, wp, map);
    else if (ws->wide) wide_dispatch(ws, &sp, &map, grid, st, false);
    else window_dispatch(ws, &sp, &map, nullptr, grid, st, false);
    return cudaGetLastError();
}

// k_stage_wide over the dimension-1 tiles [tj0, tj0 + ntj) only (bellman_stage_host runs a stage slab by slab
// while the host copies overlap).  Returns cudaErrorNotSupported when the handle does not use that kernel.
cudaError_t window_launch_tile_range(bellman_handle *h, const StageParams &sp, int slot_next, cudaStream_t st,
                                     int tj0, int ntj) {
    auto *ws = static_cast<WindowState *>(h->wstate);
    if (!ws || !h->wcfg.valid || !ws->wide || ws->strip || ws->lean) return cudaErrorNotSupported;
    WindowParams wpl = ws->wp;
    if (tj0 < 0 || ntj < 1 || tj0 + ntj > wpl.ntile1) return cudaErrorInvalidValue;
    wpl.tj_off = tj0;
    wpl.ntile1 = ntj;
    const dim3 grid((unsigned)(wpl.ntile0 * ntj), (unsigned)sp.P);
    wide_dispatch(ws, &sp, &ws->maps[slot_next], grid, st, false, &wpl);
    return cudaGetLastError();
}
// tile geometry of the wide kernel (0 when the handle does not use it)
int window_wide_tiles(const bellman_handle *h, int *tile1) {
    auto *ws = static_cast<const WindowState *>(h->wstate);
    if (!ws || !h->wcfg.valid || !ws->wide || ws->strip || ws->lean) return 0;
    if (tile1) *tile1 = (WIDE_NT / 32) * WIDE_R;
    return ws->wp.ntile1;
}

}  // namespace bellman
