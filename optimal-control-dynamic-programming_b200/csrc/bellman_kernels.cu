// bellman_kernels.cu — sm_100a kernels of the backward Bellman stage.
//
// Compiled with -fmad=false: every written operation is rounded once, exactly as
// include/bellman.h specifies; the only fused operations are the explicit fma() calls.
// No tensor cores: the stage is a gather-and-reduce, not a contraction.
#include <cstdlib>
#include <cstdint>

#include "bellman_kernels.cuh"

namespace bellman {

namespace {

constexpr int BLOCK = 256;

// --- locate (include/bellman.h, "locate_d") ---------------------------------------------------
// x is in kernel units: the fractional cell coordinate for UNIFORM dimensions (the host pre-scaled
// the tables), the state value for SEARCH dimensions.
__device__ __forceinline__ int locate(const double *__restrict__ s, const double *__restrict__ rinv, int n,
                                      int mode, const int32_t *__restrict__ lut, int lut_n, double lut_invw,
                                      double x, double &t) {
    int cell;
    if (mode == BELLMAN_LOCATE_UNIFORM) {
        cell = min(max(__double2int_rd(x), 0), n - 2);   // floor (saturating), then clamp
        t = x - (double)cell;
    } else {
        // exact bin rule  cell = clamp(#{ s[i] <= x } - 1, 0, n-2): the bucket table gives a cell at
        // or below the answer (one bucket of slack absorbs rounding), a short scan finishes it
        const int b = min(max(__double2int_rd((x - __ldg(s)) * lut_invw) - 1, 0), lut_n);
        cell = __ldg(lut + b);
        while (cell < n - 2 && __ldg(s + cell + 1) <= x) ++cell;
        t = (x - __ldg(s + cell)) * __ldg(rinv + cell);
    }
    return cell;
}

template <int D>
struct Prob {
    const double *grid[D], *rinv[D], *Ta[D], *Tb[D], *Tc[D], *q[D];
    int mode[D], n[D], lut_n[D];
    const int32_t *lut[D];
    double lut_invw[D];
    const double *r;
    __device__ __forceinline__ void load(const StageParams &sp, int prob) {
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
            lut[d] = dp.lut + (size_t)prob * (dp.lut_n + 1);
            lut_n[d] = dp.lut_n;
            lut_invw[d] = __ldg(dp.loc + 2 * prob);
        }
        r = sp.r + (size_t)prob * sp.C;
    }
};

// control-independent dimensions are located once per state; FixedDims carries their cells/weights
template <int D>
struct FixedDims {
    double t[D];
    long long off;      // element offset contributed by the control-independent dimensions
    int ro[1 << (D - 1)];   // element offsets of the 2^(D-1) dimension-0 pairs relative to the base corner
                            // (exact when the J slab has fewer than 2^31 elements: SMALL kernels)
};

template <int D>
__device__ __forceinline__ void locate_fixed(const Prob<D> &pb, const StageParams &sp, const double (&base)[D],
                                             FixedDims<D> &fx) {
    fx.off = 0;
#pragma unroll
    for (int k = 0; k < (1 << (D - 1)); ++k) {
        int r = 0;
#pragma unroll
        for (int d = 1; d < D; ++d)
            if (k & (1 << (d - 1))) r += (int)sp.dim[d].stride;
        fx.ro[k] = r;
    }
#pragma unroll
    for (int d = 0; d < D; ++d) {
        fx.t[d] = 0.0;
        if (!pb.Tc[d]) {
            const int cell = locate(pb.grid[d], pb.rinv[d], pb.n[d], pb.mode[d], pb.lut[d], pb.lut_n[d],
                                    pb.lut_invw[d], base[d], fx.t[d]);
            fx.off += (long long)(cell - sp.dim[d].ext_lo) * sp.dim[d].stride;
        }
    }
}

// one (state, control) evaluation: returns the interpolated J_{k+1}(x')
// COHERENT: read J_{k+1} with ld.global.cg (L2 only).  Needed when J_{k+1} was written earlier in
// the SAME kernel by other SMs (persistent multi-stage kernel); L1 is not coherent across SMs.
template <int D, bool COHERENT = false, bool SMALL = true>
__device__ __forceinline__ double interp_at(const Prob<D> &pb, const StageParams &sp,
                                            const double *__restrict__ Jn, const double (&base)[D],
                                            const FixedDims<D> &fx, int c) {
    double t[D];
    long long o = fx.off;
#pragma unroll
    for (int d = 0; d < D; ++d) {
        if (pb.Tc[d]) {
            const double xq = base[d] + __ldg(pb.Tc[d] + c);
            const int cell = locate(pb.grid[d], pb.rinv[d], pb.n[d], pb.mode[d], pb.lut[d], pb.lut_n[d],
                                    pb.lut_invw[d], xq, t[d]);
            o += (long long)(cell - sp.dim[d].ext_lo) * sp.dim[d].stride;
        } else {
            t[d] = fx.t[d];
        }
    }
    const double *__restrict__ p = Jn + o;
    double v[1 << D];
    if (SMALL) {
        // 32-bit corner offsets precomputed per thread: one address computation per dimension-0 pair
#pragma unroll
        for (int k = 0; k < (1 << (D - 1)); ++k) {
            const double *q = p + fx.ro[k];
            v[2 * k] = COHERENT ? __ldcg(q) : __ldg(q);
            v[2 * k + 1] = COHERENT ? __ldcg(q + 1) : __ldg(q + 1);
        }
    } else {
#pragma unroll
        for (int m = 0; m < (1 << D); ++m) {
            long long oo = 0;
#pragma unroll
            for (int d = 0; d < D; ++d)
                if (m & (1 << d)) oo += sp.dim[d].stride;
            v[m] = COHERENT ? __ldcg(p + oo) : __ldg(p + oo);
        }
    }
#pragma unroll
    for (int d = 0; d < D; ++d)               // dimension 0 reduced first
#pragma unroll
        for (int m = 0; m < (1 << (D - 1 - d)); ++m)
            v[m] = fma(t[d], v[2 * m + 1] - v[2 * m], v[2 * m]);
    return v[0];
}

// decompose an owned-state linear index into global grid indices; returns the element offset of
// the state inside the (extended) J arrays
template <int D>
__device__ __forceinline__ long long decompose(const StageParams &sp, long long s, int (&gi)[D]) {
    long long o = 0;
#pragma unroll
    for (int d = 0; d < D; ++d) {
        const int on = sp.dim[d].own_n;
        const int loc = (d == D - 1) ? (int)s : (int)(s % on);
        s /= on;
        gi[d] = loc + sp.dim[d].own_lo;
        o += (long long)(gi[d] - sp.dim[d].ext_lo) * sp.dim[d].stride;
    }
    return o;
}

// gi[k] for a run-time k without indexing the register array dynamically (that would put it in
// local memory)
template <int D>
__device__ __forceinline__ int pick(const int (&gi)[D], int k) {
    int v = gi[0];
#pragma unroll
    for (int d = 1; d < D; ++d) v = (k == d) ? gi[d] : v;
    return v;
}
template <int D>
__device__ __forceinline__ const double *pickp(const double *const (&q)[D], int k) {
    const double *v = q[0];
#pragma unroll
    for (int d = 1; d < D; ++d) v = (k == d) ? q[d] : v;
    return v;
}

template <int D>
__device__ __forceinline__ double state_terms(const Prob<D> &pb, const StageParams &sp,
                                              const int (&gi)[D], double (&base)[D]) {
#pragma unroll
    for (int d = 0; d < D; ++d) {
        double b = __ldg(pb.Ta[d] + pick<D>(gi, sp.dim[d].src_a));
        if (pb.Tb[d]) b = b + __ldg(pb.Tb[d] + pick<D>(gi, sp.dim[d].src_b));
        base[d] = b;
    }
    double gs = __ldg(pickp<D>(pb.q, sp.q_order[0]) + pick<D>(gi, sp.q_order[0]));
#pragma unroll
    for (int m = 1; m < D; ++m) gs = gs + __ldg(pickp<D>(pb.q, sp.q_order[m]) + pick<D>(gi, sp.q_order[m]));
    return gs;
}

// ---------------------------------------------------------------------------------------------
// K_direct: one thread per state
// ---------------------------------------------------------------------------------------------
template <int D, bool SMALL>
__global__ void __launch_bounds__(BLOCK, 4)
k_stage_direct(const __grid_constant__ StageParams sp) {
    const int prob = blockIdx.y;
    const long long s = (long long)blockIdx.x * BLOCK + threadIdx.x;
    if (s >= sp.S_own) return;
    Prob<D> pb;
    pb.load(sp, prob);
    int gi[D];
    const long long o_self = decompose<D>(sp, s, gi);
    double base[D];
    const double gs = state_terms<D>(pb, sp, gi, base);
    const double *__restrict__ Jn = sp.J_next + (size_t)prob * sp.S_ext;

    FixedDims<D> fx;
    locate_fixed<D>(pb, sp, base, fx);
    double best = __longlong_as_double(0x7ff0000000000000LL);   // +inf
    int arg = 0;
#pragma unroll 2
    for (int c = 0; c < sp.C; ++c) {
        const double v = interp_at<D, false, SMALL>(pb, sp, Jn, base, fx, c);
        const double tot = (gs + __ldg(pb.r + c)) + v;
        if (tot < best) { best = tot; arg = c; }
    }
    sp.J_out[(size_t)prob * sp.S_ext + o_self] = best;
    idx_store(sp.idx_out, sp.idx_bytes, (long long)prob * sp.S_own + s, arg);
    if (sp.n_peers) peer_store<D>(sp, prob, gi, best);
}

// ---------------------------------------------------------------------------------------------
// K_splitc: L lanes share one state and split the control loop; lexicographic (value, index)
// shuffle reduction keeps MATLAB's first-index tie rule.
// ---------------------------------------------------------------------------------------------
template <int D, int L>
__global__ void __launch_bounds__(BLOCK)
k_stage_splitc(const __grid_constant__ StageParams sp) {
    const int prob = blockIdx.y;
    const int lane = threadIdx.x % L;
    long long s = ((long long)blockIdx.x * BLOCK + threadIdx.x) / L;
    const bool live = s < sp.S_own;
    if (!live) s = sp.S_own - 1;            // keep the whole warp in the shuffles
    Prob<D> pb;
    pb.load(sp, prob);
    int gi[D];
    const long long o_self = decompose<D>(sp, s, gi);
    double base[D];
    const double gs = state_terms<D>(pb, sp, gi, base);
    const double *__restrict__ Jn = sp.J_next + (size_t)prob * sp.S_ext;

    FixedDims<D> fx;
    locate_fixed<D>(pb, sp, base, fx);
    double best = __longlong_as_double(0x7ff0000000000000LL);
    int arg = 0x7fffffff;
    for (int c = lane; c < sp.C; c += L) {
        const double v = interp_at<D>(pb, sp, Jn, base, fx, c);
        const double tot = (gs + __ldg(pb.r + c)) + v;
        if (tot < best) { best = tot; arg = c; }
    }
#pragma unroll
    for (int w = L / 2; w >= 1; w >>= 1) {
        const double ob = __shfl_xor_sync(0xffffffffu, best, w, L);
        const int oa = __shfl_xor_sync(0xffffffffu, arg, w, L);
        if (ob < best || (ob == best && oa < arg)) { best = ob; arg = oa; }
    }
    // no control had tot < +inf (every total +inf or NaN): first index, as the serial loop of
    // k_stage_direct / the oracle / MATLAB's min leave it
    if (arg == 0x7fffffff) arg = 0;
    if (live && lane == 0) {
        sp.J_out[(size_t)prob * sp.S_ext + o_self] = best;
        idx_store(sp.idx_out, sp.idx_bytes, (long long)prob * sp.S_own + s, arg);
        if (sp.n_peers) peer_store<D>(sp, prob, gi, best);
    }
}

// ---------------------------------------------------------------------------------------------
// K_persistent: the whole stage loop in ONE cooperative launch, for grids too small for a launch
// per stage to make sense (Solver_position: 3 x 201 x 201 states x 3 controls x 5999 stages).
// Every thread owns one (state, control-lane) pair for all stages; stages are separated by a
// grid-wide barrier (monotone atomic counter; co-residency is guaranteed by the cooperative
// launch).  J ping-pongs in global memory and stays in L2; gathers use ld.global.cg.
// ---------------------------------------------------------------------------------------------
struct PersistParams {
    double *J_base;          // slot 0
    int32_t *idx_base;
    long long J_slot_elems, idx_slot_elems;
    int store_J_all, store_idx_all, N;
    int stage_from;          // stage number of the input J of the first stage
    int n_stages;
    int lanes;               // control lanes per state (power of two <= 32)
    unsigned int *barrier;   // zero-initialised counter
};

__device__ __forceinline__ void grid_barrier(unsigned int *counter, unsigned int target) {
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        atomicAdd(counter, 1u);
        unsigned int v;
        do {
            asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(counter) : "memory");
        } while (v < target);
    }
    __syncthreads();
}

// PB: threads per CTA.  The grid barrier costs one atomic per CTA per stage, and for launch-latency-sized
// problems the barrier IS the stage time, so D = 2 also exists with 1024-thread CTAs (a quarter of the arrivals).
template <int D, int PB = BLOCK>
__global__ void __launch_bounds__(PB, (D == 2 ? 1024 / PB : 2))
k_sweep_persistent(const __grid_constant__ StageParams sp, const __grid_constant__ PersistParams pp) {
    const int L = pp.lanes;
    const long long gtid = (long long)blockIdx.x * PB + threadIdx.x;
    const long long total = sp.S_own * sp.P;              // states over all problems
    long long g = gtid / L;
    const int lane = (int)(gtid % L);
    const bool live = g < total;
    if (!live) g = total - 1;
    const int prob = (int)(g / sp.S_own);
    const long long s = g - (long long)prob * sp.S_own;
    Prob<D> pb;
    pb.load(sp, prob);
    int gi[D];
    const long long o_self = decompose<D>(sp, s, gi);
    double base[D];
    const double gs = state_terms<D>(pb, sp, gi, base);
    FixedDims<D> fx;
    locate_fixed<D>(pb, sp, base, fx);

    for (int it = 0; it < pp.n_stages; ++it) {
        const int from = pp.stage_from - it, to = from - 1;
        const int js_from = pp.store_J_all ? from - 1 : ((pp.N - from) & 1);
        const int js_to = pp.store_J_all ? to - 1 : ((pp.N - to) & 1);
        const double *__restrict__ Jn = pp.J_base + (size_t)js_from * pp.J_slot_elems + (size_t)prob * sp.S_ext;
        double *Jo = pp.J_base + (size_t)js_to * pp.J_slot_elems + (size_t)prob * sp.S_ext;
        const long long Io = (long long)(pp.store_idx_all ? to - 1 : 0) * pp.idx_slot_elems + (long long)prob * sp.S_own;

        double best = __longlong_as_double(0x7ff0000000000000LL);
        int arg = 0x7fffffff;
        for (int c = lane; c < sp.C; c += L) {
            const double v = interp_at<D, true>(pb, sp, Jn, base, fx, c);
            const double tot = (gs + __ldg(pb.r + c)) + v;
            if (tot < best) { best = tot; arg = c; }
        }
        for (int w = L / 2; w >= 1; w >>= 1) {
            const double ob = __shfl_xor_sync(0xffffffffu, best, w);
            const int oa = __shfl_xor_sync(0xffffffffu, arg, w);
            if (ob < best || (ob == best && oa < arg)) { best = ob; arg = oa; }
        }
        if (arg == 0x7fffffff) arg = 0;   // all totals +inf / NaN: first index (see k_stage_splitc)
        if (live && lane == 0) {
            Jo[o_self] = best;
            idx_store(pp.idx_base, sp.idx_bytes, Io + s, arg);
        }
        if (it + 1 < pp.n_stages) grid_barrier(pp.barrier, (unsigned int)(it + 1) * gridDim.x);
    }
}

// ---------------------------------------------------------------------------------------------
// check sums (Solver_pos_att.m:273-285): sum(J) and sum(idx+1) over the owned states, in a fixed
// (deterministic) order: per-thread strided partials -> block tree -> serial sum of block partials
// ---------------------------------------------------------------------------------------------
template <int D>
__global__ void __launch_bounds__(BLOCK)
k_check_partials(const __grid_constant__ StageParams sp, double *__restrict__ partials) {
    __shared__ double sh[2][BLOCK / 32];
    double sj = 0.0, si = 0.0;
    const long long total = sp.S_own * sp.P;
    for (long long g = (long long)blockIdx.x * BLOCK + threadIdx.x; g < total;
         g += (long long)gridDim.x * BLOCK) {
        const int prob = (int)(g / sp.S_own);
        const long long s = g - (long long)prob * sp.S_own;
        int gi[D];
        const long long o = decompose<D>(sp, s, gi);
        sj += sp.J_out[(size_t)prob * sp.S_ext + o];
        si += (double)(idx_load(sp.idx_out, sp.idx_bytes, (long long)prob * sp.S_own + s) + 1);
    }
#pragma unroll
    for (int w = 16; w >= 1; w >>= 1) {
        sj += __shfl_xor_sync(0xffffffffu, sj, w);
        si += __shfl_xor_sync(0xffffffffu, si, w);
    }
    if ((threadIdx.x & 31) == 0) { sh[0][threadIdx.x >> 5] = sj; sh[1][threadIdx.x >> 5] = si; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double a = 0.0, b = 0.0;
        for (int w = 0; w < BLOCK / 32; ++w) { a += sh[0][w]; b += sh[1][w]; }
        partials[2 * blockIdx.x] = a;
        partials[2 * blockIdx.x + 1] = b;
    }
}

__global__ void k_check_final(const double *__restrict__ partials, int n, double *__restrict__ out2) {
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        double a = 0.0, b = 0.0;
        for (int i = 0; i < n; ++i) { a += partials[2 * i]; b += partials[2 * i + 1]; }
        out2[0] = a;
        out2[1] = b;
    }
}

// ---------------------------------------------------------------------------------------------
// rollout (test/Dynamic_Solver.m:108-145,191-194): one thread per initial state
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) k_rollout(const __grid_constant__ RolloutParams rp) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= rp.batch) return;
    double x1 = rp.x0[2 * b], x2 = rp.x0[2 * b + 1];
    double *X = rp.X_out + (size_t)b * 2 * rp.N;
    double *U = rp.U_out + (size_t)b * rp.N;
    const long long S = (long long)rp.n0 * rp.n1;
    X[0] = x1;
    X[1] = x2;
    for (int k = 1; k <= rp.N - 1; ++k) {
        const int st = rp.mode == 1 ? rp.ssu_stage : k;
        const long long ib = (long long)(st - 1) * S;      // element offset of the stage's policy
        double t0, t1;
        // a free state is brought to kernel units first (include/bellman.h, bellman_rollout)
        const int c0 = locate(rp.grid0, rp.rinv0, rp.n0, rp.mode0, rp.lut0, rp.lut_n0, rp.lut_invw0,
                              rp.mode0 == BELLMAN_LOCATE_UNIFORM ? fma(x1, rp.inv_h0, rp.off0) : x1, t0);
        const int c1 = locate(rp.grid1, rp.rinv1, rp.n1, rp.mode1, rp.lut1, rp.lut_n1, rp.lut_invw1,
                              rp.mode1 == BELLMAN_LOCATE_UNIFORM ? fma(x2, rp.inv_h1, rp.off1) : x2, t1);
        const long long o = c0 + (long long)c1 * rp.n0;
        auto uat = [&](long long e) { return rp.u_values[idx_load(rp.idx_all, rp.idx_bytes, ib + e)]; };
        const double v00 = uat(o), v10 = uat(o + 1);
        const double v01 = uat(o + rp.n0), v11 = uat(o + rp.n0 + 1);
        const double a = fma(t0, v10 - v00, v00), bb = fma(t0, v11 - v01, v01);
        const double u = fma(t1, bb - a, a);
        U[k - 1] = u;
        const double nx1 = (rp.A[0] * x1 + rp.A[2] * x2) + rp.B[0] * u;
        const double nx2 = (rp.A[1] * x1 + rp.A[3] * x2) + rp.B[1] * u;
        x1 = nx1;
        x2 = nx2;
        X[2 * k] = x1;
        X[2 * k + 1] = x2;
    }
    U[rp.N - 1] = 0.0;
}

// ---------------------------------------------------------------------------------------------
// nearest-policy lookup and simplified-plant axis rollout
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ int nearest_node(const PolicyParams &pp, int d, double x) {
    double t;
    const double q = pp.mode[d] == BELLMAN_LOCATE_UNIFORM ? fma(x, pp.inv_h[d], pp.off[d]) : x;
    const int cell = locate(pp.grid[d], pp.rinv[d], pp.n[d], pp.mode[d], pp.lut[d], pp.lut_n[d], pp.lut_invw[d], q, t);
    const double lo = pp.grid[d][cell], hi = pp.grid[d][cell + 1];
    return cell + ((x - lo) >= (hi - x) ? 1 : 0);      // exact midpoint -> upper node
}

// `ibase` = element offset of the policy inside pp.idx (0, or the stage offset when time varying).
// On a partitioned handle a state whose nearest node another rank owns yields -1 (the owner answers).
__device__ __forceinline__ int policy_at(const PolicyParams &pp, long long ibase, const double *x) {
    long long o = 0, st = 1;
    for (int d = 0; d < pp.D; ++d) {
        const int node = nearest_node(pp, d, x[d]) - pp.own_lo[d];
        if (node < 0 || node >= pp.own_n[d]) return -1;
        o += (long long)node * st;
        st *= pp.own_n[d];
    }
    return idx_load(pp.idx, pp.idx_bytes, ibase + o);
}

__global__ void __launch_bounds__(128) k_policy_lookup(const __grid_constant__ PolicyParams pp) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= pp.batch) return;
    double x[MAXD];
    for (int d = 0; d < pp.D; ++d) x[d] = pp.x[(size_t)b * pp.D + d];
    pp.idx_out[b] = policy_at(pp, 0, x);
}

__global__ void __launch_bounds__(128) k_rollout_axis(const __grid_constant__ PolicyParams pp) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= pp.batch) return;
    double x[2] = {pp.x[2 * (size_t)b], pp.x[2 * (size_t)b + 1]};
    double *X = pp.X_out + (size_t)b * 2 * (pp.n_steps + 1);
    int32_t *Cc = pp.idx_out + (size_t)b * pp.n_steps;
    const int r = pp.rate_dim, o = 1 - r;
    const double h = pp.h_step;
    X[0] = x[0];
    X[1] = x[1];
    for (int k = 1; k <= pp.n_steps; ++k) {
        const int c = policy_at(pp, pp.time_varying ? (long long)(k - 1) * pp.idx_stage_stride : 0, x);
        Cc[k - 1] = c;
        const double k1 = x[r];
        const double k2 = x[r] + (k1 * h) / 2;
        const double k3 = x[r] + (k2 * h) / 2;
        const double k4 = x[r] + k3 * h;
        const double xo = x[o] + (h * (((k1 + 2 * k2) + 2 * k3) + k4)) / 6;
        const double xr = x[r] + pp.u_inc[c];
        x[r] = xr;
        x[o] = xo;
        X[2 * k] = x[0];
        X[2 * k + 1] = x[1];
    }
}


// ---------------------------------------------------------------------------------------------
// Orbital forward simulation of Solver_position.get_optimal_path (position-control/Solver_position.m:
// 189-224): per stage, the three axis policies U{1,2,3}_Opt are looked up at the nearest grid node and
// the relative-motion equations (:259-309, the target orbit propagated by the universal Kepler
// equation, private/kepler_U.m, f_and_g.m, fDot_and_gDot.m, stumpC.m, stumpS.m) are integrated over
// one stage by Runge-Kutta-Fehlberg 4(5) with adaptive steps (private/rkf45.m:49-118).  One thread
// per initial state.  Operation order as oracle/bellman_oracle.c restates it; cos/sin/cosh/sinh/pow
// are CUDA's, so parity with the oracle is a tolerance, not bit equality.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ double stumpC(double z) {
    if (z > 0) return (1 - cos(sqrt(z))) / z;
    if (z < 0) return (cosh(sqrt(-z)) - 1) / (-z);
    return 0.5;
}
__device__ __forceinline__ double stumpS(double z) {
    if (z > 0) { const double s = sqrt(z); return (s - sin(s)) / pow(s, 3.0); }
    if (z < 0) { const double s = sqrt(-z); return (sinh(s) - s) / pow(s, 3.0); }
    return 1.0 / 6;
}
struct OrbitTarget { double mu, smu, r0, vr0, alpha; double R0[3], V0[3]; };

__device__ double kepler_U(const OrbitTarget &tg, double dt) {
    const double ro = tg.r0, vro = tg.vr0, a = tg.alpha, smu = tg.smu;
    double x = smu * fabs(a) * dt;
    int n = 0;
    double ratio = 1;
    while (fabs(ratio) > 1.e-8 && n <= 1000) {
        n = n + 1;
        const double x2 = x * x;
        const double Cz = stumpC(a * x2);
        const double Sz = stumpS(a * x2);
        const double F = ro * vro / smu * x2 * Cz + (1 - a * ro) * pow(x, 3.0) * Sz + ro * x - smu * dt;
        const double dFdx = ro * vro / smu * x * (1 - a * x2 * Sz) + (1 - a * ro) * x2 * Cz + ro;
        ratio = F / dFdx;
        x = x - ratio;
    }
    return x;
}

// The part of the relative-motion equations that depends on t only (Solver_position.m:288-303 /
// Solver_pos_att.m:713-733): the target propagated to t (update_RV_target) and the five coefficient
// prefixes of the expressions below, in the reference's operation order.
__device__ void target_coef(const OrbitTarget &tg, double t, double (&c)[5]) {
    // update_RV_target (:333-361)
    const double x = kepler_U(tg, t);
    const double z = tg.alpha * (x * x);
    const double f = 1 - x * x / tg.r0 * stumpC(z);
    const double g = t - 1 / tg.smu * pow(x, 3.0) * stumpS(z);
    double R[3], V[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) R[k] = f * tg.R0[k] + g * tg.V0[k];
    const double r2 = sqrt(R[0] * R[0] + R[1] * R[1] + R[2] * R[2]);
    const double fdot = tg.smu / r2 / tg.r0 * (z * stumpS(z) - 1) * x;
    const double gdot = 1 - x * x / r2 * stumpC(z);
#pragma unroll
    for (int k = 0; k < 3; ++k) V[k] = fdot * tg.R0[k] + gdot * tg.V0[k];
    // rates (:288-308)
    const double norm_R = pow((R[0] * R[0] + R[1] * R[1]) + R[2] * R[2], .5);
    const double RdotV = (R[0] * V[0] + R[1] * V[1]) + R[2] * V[2];
    const double c0 = R[1] * V[2] - R[2] * V[1], c1 = R[2] * V[0] - R[0] * V[2], c2 = R[0] * V[1] - R[1] * V[0];
    const double H = pow((c0 * c0 + c1 * c1) + c2 * c2, .5);
    const double mu = tg.mu;
    const double nR2 = norm_R * norm_R, nR3 = pow(norm_R, 3.0), nR4 = pow(norm_R, 4.0), H2 = H * H;
    c[0] = 2 * mu / nR3 + H2 / nR4;
    c[1] = 2 * RdotV / nR4 * H;
    c[2] = 2 * H / nR2;
    c[3] = mu / nR3 - H2 / nR4;
    c[4] = -mu / nR3;
}
__device__ __forceinline__ void relative_motion(const double (&c)[5], const double (&acc)[3], const double (&y)[6], double (&dydt)[6]) {
    dydt[0] = y[3]; dydt[1] = y[4]; dydt[2] = y[5];
    dydt[3] = c[0] * y[0] - c[1] * y[1] + c[2] * y[4] + acc[0];
    dydt[4] = -c[3] * y[1] + c[1] * y[0] - c[2] * y[3] + acc[1];
    dydt[5] = c[4] * y[2] + acc[2];
}
__device__ void orbit_rates(const OrbitTarget &tg, const double (&acc)[3], double t, const double (&y)[6], double (&dydt)[6]) {
    double c[5];
    target_coef(tg, t, c);
    relative_motion(c, acc, y, dydt);
}

__device__ __forceinline__ double eps_of(double t) {   // MATLAB eps(t)
    t = fabs(t);
    if (t < 2.2250738585072014e-308) return 4.9406564584124654e-324;
    int e;
    frexp(t, &e);
    return ldexp(1.0, e - 53);
}

__global__ void __launch_bounds__(32) k_rollout_orbit(const __grid_constant__ OrbitParams op) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= op.batch) return;
    const double ra[6] = {0, 1. / 4, 3. / 8, 12. / 13, 1, 1. / 2};
    const double rb[6][5] = {{0, 0, 0, 0, 0},
                             {1. / 4, 0, 0, 0, 0},
                             {3. / 32, 9. / 32, 0, 0, 0},
                             {1932. / 2197, -7200. / 2197, 7296. / 2197, 0, 0},
                             {439. / 216, -8, 3680. / 513, -845. / 4104, 0},
                             {-8. / 27, 2, -3544. / 2565, 1859. / 4104, -11. / 40}};
    const double c4[6] = {25. / 216, 0, 1408. / 2565, 2197. / 4104, -1. / 5, 0};
    const double c5[6] = {16. / 135, 0, 6656. / 12825, 28561. / 56430, -9. / 50, 2. / 55};
    OrbitTarget tg;
    tg.mu = op.mu;
    tg.smu = sqrt(op.mu);
#pragma unroll
    for (int k = 0; k < 3; ++k) { tg.R0[k] = op.R0[k]; tg.V0[k] = op.V0[k]; }
    tg.r0 = sqrt(tg.R0[0] * tg.R0[0] + tg.R0[1] * tg.R0[1] + tg.R0[2] * tg.R0[2]);
    const double v0 = sqrt(tg.V0[0] * tg.V0[0] + tg.V0[1] * tg.V0[1] + tg.V0[2] * tg.V0[2]);
    tg.vr0 = (tg.R0[0] * tg.V0[0] + tg.R0[1] * tg.V0[1] + tg.R0[2] * tg.V0[2]) / tg.r0;
    tg.alpha = 2 / tg.r0 - v0 * v0 / tg.mu;

    double y[6];
#pragma unroll
    for (int k = 0; k < 6; ++k) y[k] = op.y0[(size_t)b * 6 + k];
    const int n_out = op.n_steps / op.stride_out;
    double *X = op.X_out + (size_t)b * 6 * (n_out + 1);
    int32_t *Cc = op.C_out + (size_t)b * 3 * n_out;
#pragma unroll
    for (int k = 0; k < 6; ++k) X[k] = y[k];
    int warns = 0;
    for (int ks = 1; ks <= op.n_steps; ++ks) {
        double acc[3];
        int ci[3];
#pragma unroll
        for (int p = 0; p < 3; ++p) {       // a_x = U1_Opt(x1, v1) ... (:215-217)
            const double xq[2] = {y[p], y[3 + p]};
            ci[p] = policy_at(op.pol[p], 0, xq);
            acc[p] = op.u_values[ci[p]];
        }
        if ((ks - 1) % op.stride_out == 0) {
            const int o = (ks - 1) / op.stride_out;
#pragma unroll
            for (int p = 0; p < 3; ++p) Cc[(size_t)o * 3 + p] = ci[p];
        }
        // rkf45(@rates, [tspan(k), tspan(k+1)], X(:,k)) (:222), rkf45.m:67-118
        const double t0 = (double)(ks - 1) * op.h, tf = (double)ks * op.h;
        double t = t0, hh = (tf - t0) / 100, f[6][6], yi[6], yin[6];
        int iters = 0;
        while (t < tf && iters++ < op.max_rkf) {
            const double hmin = 16 * eps_of(t);
            const double ti = t;
#pragma unroll
            for (int k = 0; k < 6; ++k) yi[k] = y[k];
#pragma unroll
            for (int i = 0; i < 6; ++i) {
                const double t_inner = ti + ra[i] * hh;
#pragma unroll
                for (int k = 0; k < 6; ++k) yin[k] = yi[k];
#pragma unroll
                for (int j = 0; j < 5; ++j)
                    if (j < i) {
#pragma unroll
                        for (int k = 0; k < 6; ++k) yin[k] = yin[k] + hh * rb[i][j] * f[j][k];
                    }
                orbit_rates(tg, acc, t_inner, yin, f[i]);
            }
            double te_max = 0, ymax = 0;
#pragma unroll
            for (int k = 0; k < 6; ++k) {
                double te = 0;
#pragma unroll
                for (int i = 0; i < 6; ++i) te = te + (hh * f[i][k]) * (c4[i] - c5[i]);
                te_max = fmax(te_max, fabs(te));
                ymax = fmax(ymax, fabs(y[k]));
            }
            const double te_allowed = op.tol * fmax(ymax, 1.0);
            const double delta = pow(te_allowed / (te_max + 2.220446049250313e-16), 1. / 5);
            if (te_max <= te_allowed) {
                hh = fmin(hh, tf - t);
                t = t + hh;
#pragma unroll
                for (int k = 0; k < 6; ++k) {
                    double sacc = 0;
#pragma unroll
                    for (int i = 0; i < 6; ++i) sacc = sacc + (hh * f[i][k]) * c5[i];
                    y[k] = yi[k] + sacc;
                }
            }
            hh = fmin(delta * hh, 4 * hh);
            if (hh < hmin) { ++warns; break; }
        }
        if (ks % op.stride_out == 0) {
            double *Xk = X + (size_t)(ks / op.stride_out) * 6;
#pragma unroll
            for (int k = 0; k < 6; ++k) Xk[k] = y[k];
        }
    }
    if (op.warn_out) op.warn_out[b] = warns;
}

// ---------------------------------------------------------------------------------------------
// Full-plant forward simulations the reference integrates with ode45 (one thread per initial state):
//   kind 0  Solver_pos_att.get_optimal_path, pos-att/Solver_pos_att.m:452-500 — per stage the twelve
//           thruster levels from the three 4-D channel policies (:404-449, frame change :411-415,
//           ECI2body :825-829, RSW2ECI :831-847), moments and RSW accelerations (:805-823), then
//           ode45 over [tspan(k), tspan(k+1)] on the 13-state plant of :696-754;
//   kind 1  Solver_attitude.get_optimal_path_simplified_testode45, attitude-control/
//           Solver_attitude.m:1669-1705 — three 2-D axis policies at (w_k, 2*asin(q_k)), ode45 on the
//           7-state plant of :1803-1849.
// ode45 = Dormand-Prince 5(4) with MATLAB's published step control and default options (RelTol 1e-3,
// AbsTol 1e-6, MaxStep (tf-t0)/10, initial step from y'(t0)); operation order as oracle_ode45_last in
// oracle/bellman_oracle.c states it.  pow / asin / the Kepler transcendentals are CUDA's, so parity
// with the oracle is a tolerance.
// ---------------------------------------------------------------------------------------------
struct Chol3 { double r11, r12, r13, r22, r23, r33; };
__device__ __forceinline__ void chol3_factor(const double *A, Chol3 &c) {       // A column-major, symmetric
    c.r11 = sqrt(A[0]);
    c.r12 = A[3] / c.r11;
    c.r13 = A[6] / c.r11;
    c.r22 = sqrt(A[4] - c.r12 * c.r12);
    c.r23 = (A[7] - c.r12 * c.r13) / c.r22;
    c.r33 = sqrt(A[8] - (c.r13 * c.r13 + c.r23 * c.r23));
}
__device__ __forceinline__ void chol3_solve(const Chol3 &c, const double *b, double *x) {
    const double z1 = b[0] / c.r11;
    const double z2 = (b[1] - c.r12 * z1) / c.r22;
    const double z3 = ((b[2] - c.r13 * z1) - c.r23 * z2) / c.r33;
    x[2] = z3 / c.r33;
    x[1] = (z2 - c.r23 * x[2]) / c.r22;
    x[0] = ((z1 - c.r12 * x[1]) - c.r13 * x[2]) / c.r11;
}
__device__ void lu3_solve(const double (&M)[3][3], const double *b, double *x) {  // partial pivoting
    double a[3][4];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
#pragma unroll
        for (int j = 0; j < 3; ++j) a[i][j] = M[i][j];
        a[i][3] = b[i];
    }
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        int p = k;
#pragma unroll
        for (int i = k + 1; i < 3; ++i) if (fabs(a[i][k]) > fabs(a[p][k])) p = i;
#pragma unroll
        for (int r = k + 1; r < 3; ++r)
            if (p == r) {
#pragma unroll
                for (int j = 0; j < 4; ++j) { const double tmp = a[k][j]; a[k][j] = a[r][j]; a[r][j] = tmp; }
            }
#pragma unroll
        for (int i = k + 1; i < 3; ++i) {
            const double l = a[i][k] / a[k][k];
#pragma unroll
            for (int j = k + 1; j < 4; ++j) a[i][j] = a[i][j] - l * a[k][j];
        }
    }
    x[2] = a[2][3] / a[2][2];
    x[1] = (a[1][3] - a[1][2] * x[2]) / a[1][1];
    x[0] = ((a[0][3] - a[0][1] * x[1]) - a[0][2] * x[2]) / a[0][0];
}
__device__ __forceinline__ void matvec3(const double (&M)[3][3], const double *v, double *o) {
#pragma unroll
    for (int i = 0; i < 3; ++i) o[i] = (M[i][0] * v[0] + M[i][1] * v[1]) + M[i][2] * v[2];
}
__device__ __forceinline__ void eci2body(const double *q, double (&M)[3][3]) {
    M[0][0] = 1 - 2 * (q[1] * q[1] + q[2] * q[2]); M[0][1] = 2 * (q[0] * q[1] + q[2] * q[3]); M[0][2] = 2 * (q[0] * q[2] - q[1] * q[3]);
    M[1][0] = 2 * (q[1] * q[0] - q[2] * q[3]); M[1][1] = 1 - 2 * (q[0] * q[0] + q[2] * q[2]); M[1][2] = 2 * (q[1] * q[2] + q[0] * q[3]);
    M[2][0] = 2 * (q[2] * q[0] + q[1] * q[3]); M[2][1] = 2 * (q[2] * q[1] - q[0] * q[3]); M[2][2] = 1 - 2 * (q[0] * q[0] + q[1] * q[1]);
}
__device__ __forceinline__ void rsw2eci(const double *pos, const double *vel, double (&M)[3][3]) {
    const double np_ = sqrt((pos[0] * pos[0] + pos[1] * pos[1]) + pos[2] * pos[2]);
    const double c[3] = {pos[1] * vel[2] - pos[2] * vel[1], pos[2] * vel[0] - pos[0] * vel[2], pos[0] * vel[1] - pos[1] * vel[0]};
    const double nc = sqrt((c[0] * c[0] + c[1] * c[1]) + c[2] * c[2]);
    const double R[3] = {pos[0] / np_, pos[1] / np_, pos[2] / np_};
    const double W[3] = {c[0] / nc, c[1] / nc, c[2] / nc};
    const double S[3] = {W[1] * R[2] - W[2] * R[1], W[2] * R[0] - W[0] * R[2], W[0] * R[1] - W[1] * R[0]};
#pragma unroll
    for (int i = 0; i < 3; ++i) { M[i][0] = R[i]; M[i][1] = S[i]; M[i][2] = W[i]; }
}
// obj.InertiaM\(U - cross(w, obj.InertiaM*w))
__device__ __forceinline__ void euler_wdot(const double *Im, const Chol3 &ch, const double *U, const double *w, double *wd) {
    double Iw[3], b[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) Iw[i] = (Im[i] * w[0] + Im[3 + i] * w[1]) + Im[6 + i] * w[2];
    b[0] = U[0] - (w[1] * Iw[2] - w[2] * Iw[1]);
    b[1] = U[1] - (w[2] * Iw[0] - w[0] * Iw[2]);
    b[2] = U[2] - (w[0] * Iw[1] - w[1] * Iw[0]);
    chol3_solve(ch, b, wd);
}

// Warp-cooperative target ephemeris.  The Kepler propagation of the target is by far the most expensive
// part of a rate evaluation and depends on t only, and the lanes of a warp march in lock step as long
// as ode45's step limit binds (MaxStep = (tf - t0)/10: always, for this plant).  So before each ode45 call
// the warp predicts the 61 evaluation times of the call (f0, then six per step of ten steps) and its
// lanes compute the coefficient sets of two times each, in parallel, into a shared table keyed by the
// exact bits of t.  A rate evaluation looks its t up (cursor first, then a scan); a lane whose steps
// leave the predicted sequence (a rejected step, another first step) simply misses and computes its own
// coefficients.  Same function of the same t by the same instructions: results are bit-identical to
// evaluating target_coef at every call.
constexpr int EPH_SLOTS = 61;
struct PosAttRhs {                         // Solver_pos_att.m:696-754
    static constexpr int NEQ = 13;
    OrbitTarget tg;
    double acc[3], UM[3], Im[9];
    Chol3 ch;
    const double (*eph)[6];                // [EPH_SLOTS][t, c0..c4] in shared memory, or nullptr
    mutable int cur;
    __device__ void operator()(double t, const double (&X)[13], double (&Xd)[13]) const {
        double c[5], y6[6], d6[6];
        int hit = -1;
        if (eph) {
            if (cur < EPH_SLOTS && eph[cur][0] == t) hit = cur;
            else
                for (int e = 0; e < EPH_SLOTS; ++e)
                    if (eph[e][0] == t) { hit = e; break; }
        }
        if (hit >= 0) {
#pragma unroll
            for (int k = 0; k < 5; ++k) c[k] = eph[hit][1 + k];
            cur = hit + 1;
        } else {
            target_coef(tg, t, c);
        }
#pragma unroll
        for (int k = 0; k < 6; ++k) y6[k] = X[k];
        relative_motion(c, acc, y6, d6);
#pragma unroll
        for (int k = 0; k < 6; ++k) Xd[k] = d6[k];
        const double q1 = X[6], q2 = X[7], q3 = X[8], q4 = X[9], w1 = X[10], w2 = X[11], w3 = X[12];
        Xd[6] = 0.5 * ((w3 * q2 - w2 * q3) + w1 * q4);
        Xd[7] = 0.5 * ((-w3 * q1 + w1 * q3) + w2 * q4);
        Xd[8] = 0.5 * ((w2 * q1 - w1 * q2) + w3 * q4);
        Xd[9] = 0.5 * ((-w1 * q1 - w2 * q2) - w3 * q3);
        euler_wdot(Im, ch, UM, &X[10], &Xd[10]);
    }
};
// evaluation time number e (0 = f0 at t0; 1 + 6 s + k = stage k of step s) of ode45 over [t0, tf] when
// every step is accepted at the step limit; NaN beyond the last step
__device__ double eph_time(double t0, double tf, int e) {
    if (e == 0) return t0;
    const double A[6] = {1. / 5, 3. / 10, 4. / 5, 8. / 9, 1, 1};
    const double hmax = fabs(0.1 * (tf - t0));
    const int s_want = (e - 1) / 6, k = (e - 1) % 6;
    double t = t0;
    for (int s = 0;; ++s) {
        const double hmin = 16 * eps_of(t);
        const double absh = fmin(hmax, fmax(hmin, hmax));
        double h = absh;
        bool done = false;
        if (1.1 * absh >= fabs(tf - t)) { h = tf - t; done = true; }
        double tnew = t + h * A[5];
        if (done) tnew = tf;
        if (s == s_want) return k < 5 ? t + h * A[k] : tnew;
        if (done) return __longlong_as_double(0x7ff8000000000000LL);
        t = tnew;
    }
}
struct AttRhs {                            // Solver_attitude.m:1803-1849
    static constexpr int NEQ = 7;
    double U[3], Im[9];
    Chol3 ch;
    __device__ void operator()(double, const double (&X)[7], double (&Xd)[7]) const {
        const double x1 = X[0], x2 = X[1], x3 = X[2], x4 = X[3], x5 = X[4], x6 = X[5], x7 = X[6];
        euler_wdot(Im, ch, U, &X[0], &Xd[0]);
        Xd[3] = 0.5 * ((x3 * x5 - x2 * x6) + x1 * x7);
        Xd[4] = 0.5 * ((-x3 * x4 + x1 * x6) + x2 * x7);
        Xd[5] = 0.5 * ((x2 * x4 - x1 * x5) + x3 * x7);
        Xd[6] = 0.5 * ((-x1 * x4 - x2 * x5) - x3 * x6);
    }
};

// [~, Y] = ode45(rhs, [t0 tf], y); y = Y(end, :).  Returns 1 when MATLAB would have warned (step size at
// hmin) or the step bound was hit.
template <class Rhs>
__device__ int ode45_last(const Rhs &rhs, double t0, double tf, double (&y)[Rhs::NEQ], double rtol, double atol, int max_steps) {
    constexpr int NEQ = Rhs::NEQ;
    constexpr double A[6] = {1. / 5, 3. / 10, 4. / 5, 8. / 9, 1, 1};
    constexpr double B[7][6] = {{1. / 5, 3. / 40, 44. / 45, 19372. / 6561, 9017. / 3168, 35. / 384},
                                {0, 9. / 40, -56. / 15, -25360. / 2187, -355. / 33, 0},
                                {0, 0, 32. / 9, 64448. / 6561, 46732. / 5247, 500. / 1113},
                                {0, 0, 0, -212. / 729, 49. / 176, 125. / 192},
                                {0, 0, 0, 0, -5103. / 18656, -2187. / 6784},
                                {0, 0, 0, 0, 0, 11. / 84},
                                {0, 0, 0, 0, 0, 0}};
    constexpr double E[7] = {71. / 57600, 0, -71. / 16695, 71. / 1920, -17253. / 339200, 22. / 525, -1. / 40};
    const double pw = 1. / 5;
    double f[7][NEQ], ys[NEQ];
    const double htspan = fabs(tf - t0), hmax = fabs(0.1 * (tf - t0)), threshold = atol / rtol;
    double t = t0;
    int steps = 0;
    bool done = false;
    rhs(t, y, f[0]);
    double hmin = 16 * eps_of(t);
    double absh = fmin(hmax, htspan);
    double rh = 0;
#pragma unroll
    for (int i = 0; i < NEQ; ++i) rh = fmax(rh, fabs(f[0][i] / fmax(fabs(y[i]), threshold)));
    rh = rh / (0.8 * pow(rtol, pw));
    if (absh * rh > 1) absh = 1 / rh;
    absh = fmax(absh, hmin);
    while (!done) {
        if (steps >= max_steps) return 1;
        hmin = 16 * eps_of(t);
        absh = fmin(hmax, fmax(hmin, absh));
        double h = absh;
        if (1.1 * absh >= fabs(tf - t)) {
            h = tf - t;
            absh = fabs(h);
            done = true;
        }
        bool nofailed = true;
        double err, tnew;
        for (;;) {
#pragma unroll
            for (int k = 0; k < 6; ++k) {
#pragma unroll
                for (int i = 0; i < NEQ; ++i) {
                    double s = 0;
#pragma unroll
                    for (int j = 0; j <= k; ++j) s = s + f[j][i] * (h * B[j][k]);
                    ys[i] = y[i] + s;
                }
                if (k < 5) rhs(t + h * A[k], ys, f[k + 1]);
            }
            tnew = t + h * A[5];
            if (done) tnew = tf;
            rhs(tnew, ys, f[6]);
            err = 0;
#pragma unroll
            for (int i = 0; i < NEQ; ++i) {
                double fe = 0;
#pragma unroll
                for (int j = 0; j < 7; ++j) fe = fe + f[j][i] * E[j];
                err = fmax(err, fabs(fe / fmax(fmax(fabs(y[i]), fabs(ys[i])), threshold)));
            }
            err = absh * err;
            if (err > rtol) {
                if (absh <= hmin) return 1;
                if (nofailed) {
                    nofailed = false;
                    absh = fmax(hmin, absh * fmax(0.1, 0.8 * pow(rtol / err, pw)));
                } else {
                    absh = fmax(hmin, 0.5 * absh);
                }
                h = absh;
                done = false;
            } else {
                break;
            }
        }
        ++steps;
        if (!done && nofailed) {
            const double temp = 1.25 * pow(err / rtol, pw);
            if (temp > 0.2) absh = absh / temp;
            else absh = 5.0 * absh;
        }
        t = tnew;
#pragma unroll
        for (int i = 0; i < NEQ; ++i) { y[i] = ys[i]; f[0][i] = f[6][i]; }
    }
    return 0;
}

__global__ void __launch_bounds__(32) k_rollout_pos_att(const __grid_constant__ PlantParams pl) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= pl.batch) return;
    PosAttRhs rhs;
    OrbitTarget &tg = rhs.tg;
    tg.mu = pl.mu;
    tg.smu = sqrt(pl.mu);
#pragma unroll
    for (int k = 0; k < 3; ++k) { tg.R0[k] = pl.R0[k]; tg.V0[k] = pl.V0[k]; }
    tg.r0 = sqrt(tg.R0[0] * tg.R0[0] + tg.R0[1] * tg.R0[1] + tg.R0[2] * tg.R0[2]);
    const double v0 = sqrt(tg.V0[0] * tg.V0[0] + tg.V0[1] * tg.V0[1] + tg.V0[2] * tg.V0[2]);
    tg.vr0 = (tg.R0[0] * tg.V0[0] + tg.R0[1] * tg.V0[1] + tg.R0[2] * tg.V0[2]) / tg.r0;
    tg.alpha = 2 / tg.r0 - v0 * v0 / tg.mu;
#pragma unroll
    for (int k = 0; k < 9; ++k) rhs.Im[k] = pl.Im[k];
    chol3_factor(rhs.Im, rhs.ch);
    double M1[3][3];
    rsw2eci(pl.R0, pl.V0, M1);

    double y[13];
#pragma unroll
    for (int k = 0; k < 13; ++k) y[k] = pl.y0[(size_t)b * 13 + k];
    const int n_out = pl.n_steps / pl.stride_out;
    double *X = pl.X_out + (size_t)b * 13 * (n_out + 1);
#pragma unroll
    for (int k = 0; k < 13; ++k) X[k] = y[k];
    int warns = 0;
    __shared__ double eph[EPH_SLOTS][6];
    const unsigned live = __activemask();                  // lanes of this (one-warp) CTA that own a trajectory
    const int n_live = __popc(live), my_rank = __popc(live & ((1u << (threadIdx.x & 31)) - 1));
    rhs.eph = eph;
    for (int ks = 1; ks <= pl.n_steps; ++ks) {
        {   // the warp's ephemeris table for this stage's ode45 call
            const double t0s = (double)(ks - 1) * pl.h, tfs = (double)ks * pl.h;
            __syncwarp(live);
            for (int e = my_rank; e < EPH_SLOTS; e += n_live) {
                const double te = eph_time(t0s, tfs, e);
                eph[e][0] = te;
                if (te == te) {
                    double c[5];
                    target_coef(rhs.tg, te, c);
#pragma unroll
                    for (int k = 0; k < 5; ++k) eph[e][1 + k] = c[k];
                }
            }
            __syncwarp(live);
            rhs.cur = 0;
        }
        double tq[3], M2[3][3], tmp[3], xb[3], vb[3], f[12];
#pragma unroll
        for (int k = 0; k < 3; ++k) tq[k] = 2 * asin(y[6 + k]);            // :472-474
        eci2body(&y[6], M2);
        matvec3(M1, &y[0], tmp); matvec3(M2, tmp, xb);                      // :414-415
        matvec3(M1, &y[3], tmp); matvec3(M2, tmp, vb);
#pragma unroll
        for (int p = 0; p < 3; ++p) {                                       // :432-447
            const int a = p == 0 ? 1 : (p == 1 ? 2 : 0);                    // channel x: t_y, w_y; y: t_z, w_z; z: t_x, w_x
            const double xq[4] = {xb[p], vb[p], tq[a], y[10 + a]};
            const int ci = policy_at(pl.pol[p], 0, xq);
            const double *fv = pl.fv[p];
            const int C = pl.C[p];
            f[2 * p] = fv[ci];                                              // thrusters (0,1,6,7), (2,3,8,9), (4,5,10,11)
            f[2 * p + 1] = fv[C + ci];
            f[6 + 2 * p] = fv[2 * C + ci];
            f[7 + 2 * p] = fv[3 * C + ci];
        }
        // to_Moments_Forces :805-823
        const double UMy = (((f[0] - f[1]) + f[6]) - f[7]) * pl.t_dist;
        const double UMz = (((f[2] - f[3]) + f[8]) - f[9]) * pl.t_dist;
        const double UMx = (((f[4] - f[5]) + f[10]) - f[11]) * pl.t_dist;
        const double ab[3] = {(((f[0] + f[1]) + f[6]) + f[7]) / pl.mass, (((f[2] + f[3]) + f[8]) + f[9]) / pl.mass,
                              (((f[4] + f[5]) + f[10]) + f[11]) / pl.mass};
        lu3_solve(M2, ab, tmp);
        lu3_solve(M1, tmp, rhs.acc);
        rhs.UM[0] = UMx; rhs.UM[1] = UMy; rhs.UM[2] = UMz;
        if ((ks - 1) % pl.stride_out == 0) {
            const size_t o = (size_t)b * n_out + (ks - 1) / pl.stride_out;
#pragma unroll
            for (int k = 0; k < 12; ++k) pl.F_out[o * 12 + k] = f[k];
            if (pl.FM_out) {
#pragma unroll
                for (int k = 0; k < 3; ++k) { pl.FM_out[o * 6 + k] = rhs.acc[k]; pl.FM_out[o * 6 + 3 + k] = rhs.UM[k]; }
            }
        }
        warns += ode45_last(rhs, (double)(ks - 1) * pl.h, (double)ks * pl.h, y, pl.rtol, pl.atol, pl.max_ode);
        if (ks % pl.stride_out == 0) {
            double *Xk = X + (size_t)(ks / pl.stride_out) * 13;
#pragma unroll
            for (int k = 0; k < 13; ++k) Xk[k] = y[k];
        }
    }
    if (pl.warn_out) pl.warn_out[b] = warns;
}

__global__ void __launch_bounds__(32) k_rollout_attitude(const __grid_constant__ PlantParams pl) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= pl.batch) return;
    AttRhs rhs;
#pragma unroll
    for (int k = 0; k < 9; ++k) rhs.Im[k] = pl.Im[k];
    chol3_factor(rhs.Im, rhs.ch);
    double y[7];
#pragma unroll
    for (int k = 0; k < 7; ++k) y[k] = pl.y0[(size_t)b * 7 + k];
    const int n_out = pl.n_steps / pl.stride_out;
    double *X = pl.X_out + (size_t)b * 7 * (n_out + 1);
    int32_t *Cc = pl.C_out + (size_t)b * 3 * n_out;
#pragma unroll
    for (int k = 0; k < 7; ++k) X[k] = y[k];
    int warns = 0;
    for (int ks = 1; ks <= pl.n_steps; ++ks) {
        int ci[3];
#pragma unroll
        for (int p = 0; p < 3; ++p) {                                       // U1(k) = FU_k(X(k), 2*asin(X(3+k))) :1693-1697
            const double xq[2] = {y[p], 2 * asin(y[3 + p])};
            ci[p] = policy_at(pl.pol[p], 0, xq);
            rhs.U[p] = pl.fv[0][ci[p]];
        }
        if ((ks - 1) % pl.stride_out == 0) {
            const int o = (ks - 1) / pl.stride_out;
#pragma unroll
            for (int p = 0; p < 3; ++p) Cc[(size_t)o * 3 + p] = ci[p];
        }
        warns += ode45_last(rhs, (double)(ks - 1) * pl.h, (double)ks * pl.h, y, pl.rtol, pl.atol, pl.max_ode);
        if (ks % pl.stride_out == 0) {
            double *Xk = X + (size_t)(ks / pl.stride_out) * 7;
#pragma unroll
            for (int k = 0; k < 7; ++k) Xk[k] = y[k];
        }
    }
    if (pl.warn_out) pl.warn_out[b] = warns;
}

}  // namespace

// ---------------------------------------------------------------------------------------------
// launchers
// ---------------------------------------------------------------------------------------------
cudaError_t launch_stage_direct(const StageParams &sp, cudaStream_t st) {
    const dim3 grid((unsigned)((sp.S_own + BLOCK - 1) / BLOCK), (unsigned)sp.P);
    const bool small = sp.S_ext < (1LL << 31);     // 32-bit corner offsets are exact
    switch (sp.D * 2 + (small ? 1 : 0)) {
        case 4: k_stage_direct<2, false><<<grid, BLOCK, 0, st>>>(sp); break;
        case 5: k_stage_direct<2, true><<<grid, BLOCK, 0, st>>>(sp); break;
        case 6: k_stage_direct<3, false><<<grid, BLOCK, 0, st>>>(sp); break;
        case 7: k_stage_direct<3, true><<<grid, BLOCK, 0, st>>>(sp); break;
        case 8: k_stage_direct<4, false><<<grid, BLOCK, 0, st>>>(sp); break;
        case 9: k_stage_direct<4, true><<<grid, BLOCK, 0, st>>>(sp); break;
        default: return cudaErrorInvalidValue;
    }
    return cudaGetLastError();
}

template <int D>
static cudaError_t splitc_D(const StageParams &sp, int L, cudaStream_t st) {
    const long long threads = sp.S_own * L;
    const dim3 grid((unsigned)((threads + BLOCK - 1) / BLOCK), (unsigned)sp.P);
    switch (L) {
        case 2: k_stage_splitc<D, 2><<<grid, BLOCK, 0, st>>>(sp); break;
        case 4: k_stage_splitc<D, 4><<<grid, BLOCK, 0, st>>>(sp); break;
        case 8: k_stage_splitc<D, 8><<<grid, BLOCK, 0, st>>>(sp); break;
        case 16: k_stage_splitc<D, 16><<<grid, BLOCK, 0, st>>>(sp); break;
        case 32: k_stage_splitc<D, 32><<<grid, BLOCK, 0, st>>>(sp); break;
        default: return cudaErrorInvalidValue;
    }
    return cudaGetLastError();
}

cudaError_t launch_stage_splitc(const StageParams &sp, int lanes_per_state, cudaStream_t st) {
    if (sp.S_ext >= (1LL << 31)) return launch_stage_direct(sp, st);   // 32-bit corner offsets only
    switch (sp.D) {
        case 2: return splitc_D<2>(sp, lanes_per_state, st);
        case 3: return splitc_D<3>(sp, lanes_per_state, st);
        case 4: return splitc_D<4>(sp, lanes_per_state, st);
        default: return cudaErrorInvalidValue;
    }
}

cudaError_t launch_check_sums(const StageParams &sp, double *d_partials, int n_partials,
                              double *d_out2, cudaStream_t st) {
    switch (sp.D) {
        case 2: k_check_partials<2><<<n_partials, BLOCK, 0, st>>>(sp, d_partials); break;
        case 3: k_check_partials<3><<<n_partials, BLOCK, 0, st>>>(sp, d_partials); break;
        case 4: k_check_partials<4><<<n_partials, BLOCK, 0, st>>>(sp, d_partials); break;
        default: return cudaErrorInvalidValue;
    }
    k_check_final<<<1, 32, 0, st>>>(d_partials, n_partials, d_out2);
    return cudaGetLastError();
}

int persistent_capacity_threads(int D) {
    int dev = 0, sms = 0, per_sm = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaError_t e = cudaSuccess;
    switch (D) {
        case 2: e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_sweep_persistent<2>, BLOCK, 0); break;   // (the 1024-thread form holds the same number of threads)
        case 3: e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_sweep_persistent<3>, BLOCK, 0); break;
        case 4: e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_sweep_persistent<4>, BLOCK, 0); break;
        default: return 0;
    }
    if (e != cudaSuccess) { cudaGetLastError(); return 0; }
    return sms * per_sm * BLOCK;
}

cudaError_t launch_sweep_persistent(const StageParams &sp, double *J_base, int32_t *idx_base,
                                    long long J_slot_elems, long long idx_slot_elems, int store_J_all,
                                    int store_idx_all, int N, int stage_from, int n_stages, int lanes,
                                    unsigned int *d_barrier, cudaStream_t st) {
    PersistParams pp;
    pp.J_base = J_base; pp.idx_base = idx_base;
    pp.J_slot_elems = J_slot_elems; pp.idx_slot_elems = idx_slot_elems;
    pp.store_J_all = store_J_all; pp.store_idx_all = store_idx_all; pp.N = N;
    pp.stage_from = stage_from; pp.n_stages = n_stages; pp.lanes = lanes; pp.barrier = d_barrier;
    cudaError_t e = cudaMemsetAsync(d_barrier, 0, sizeof(unsigned int), st);
    if (e != cudaSuccess) return e;
    const long long threads = sp.S_own * sp.P * lanes;
    const dim3 grid((unsigned)((threads + BLOCK - 1) / BLOCK));
    StageParams spc = sp;
    void *args[] = {(void *)&spc, (void *)&pp};
    const bool small_block = std::getenv("BELLMAN_PERSIST_BLOCK256") != nullptr;
    if (sp.D == 2 && !small_block) {
        const dim3 grid1k((unsigned)((threads + 1023) / 1024));
        return cudaLaunchCooperativeKernel((const void *)k_sweep_persistent<2, 1024>, grid1k, dim3(1024), args, 0, st);
    }
    switch (sp.D) {
        case 2: return cudaLaunchCooperativeKernel((const void *)k_sweep_persistent<2>, grid, dim3(BLOCK), args, 0, st);
        case 3: return cudaLaunchCooperativeKernel((const void *)k_sweep_persistent<3>, grid, dim3(BLOCK), args, 0, st);
        case 4: return cudaLaunchCooperativeKernel((const void *)k_sweep_persistent<4>, grid, dim3(BLOCK), args, 0, st);
        default: return cudaErrorInvalidValue;
    }
}

cudaError_t launch_rollout(const RolloutParams &rp, cudaStream_t st) {
    k_rollout<<<(rp.batch + 127) / 128, 128, 0, st>>>(rp);
    return cudaGetLastError();
}

cudaError_t launch_policy_lookup(const PolicyParams &pp, cudaStream_t st) {
    k_policy_lookup<<<(pp.batch + 127) / 128, 128, 0, st>>>(pp);
    return cudaGetLastError();
}

cudaError_t launch_rollout_axis(const PolicyParams &pp, cudaStream_t st) {
    k_rollout_axis<<<(pp.batch + 127) / 128, 128, 0, st>>>(pp);
    return cudaGetLastError();
}

cudaError_t launch_rollout_plant(const PlantParams &pl, cudaStream_t st) {
    // one warp per CTA: the trajectories are long serial chains (latency-bound), so a batch of a few
    // thousand initial states must spread over all 148 SMs rather than fill a few of them
    if (pl.kind == 0) k_rollout_pos_att<<<(pl.batch + 31) / 32, 32, 0, st>>>(pl);
    else k_rollout_attitude<<<(pl.batch + 31) / 32, 32, 0, st>>>(pl);
    return cudaGetLastError();
}
cudaError_t launch_rollout_orbit(const OrbitParams &op, cudaStream_t st) {
    k_rollout_orbit<<<(op.batch + 31) / 32, 32, 0, st>>>(op);      // one warp per CTA (see launch_rollout_plant)
    return cudaGetLastError();
}

// window kernel: see bellman_window.cu
}  // namespace bellman
