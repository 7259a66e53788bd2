// bellman_dense6.cu — the coupled 6-D attitude sweep, Solver_attitude.run
// (attitude-control/Solver_attitude.m:521-601, calculate_J_U_opt_state_M :767-823): one fused stage kernel
// that replaces the nine-dimensional arrays the reference builds with repmat (next states of every
// (state, U1, U2, U3), J_current_state_fix + F(...), three nested min calls).
//
// Compiled with -fmad=false: one rounding per written operation; the only fused operations are the
// explicit fma() calls of the lerps (include/bellman.h, "dense stage operator").
//
// Data layout in HBM (dimension 0 = w1 fastest, S = n0*n1*n2*n3*n4*n5, S3 = n0*n1*n2):
//   J_next, J_out [S] fp64 (ping-pong), idx [S] int32
//   a_next[3] [S]          next yaw / pitch / roll of every state (no control dependence)
//   w_next[3] [nu][S3]     next w_d for level u of control d (no angle dependence: read through L1/L2)
//   gs [S]                 state cost
// Algorithmic HBM bytes per state and stage: 8 (J_next) + 24 (a_next) + 8 (gs) + 8 (J) + 4 (idx) = 52; the
// 64-corner gathers and the w_next reads hit L1 / L2.  With nu^3 = 27 controls the stage is bound by the
// fp64 pipe (about 60 lerps of 2 fp64 instructions per control after the sharing below), not by HBM.
//
// One thread per state.  The interpolation reduces dimension 0 first (w1, weight of U1), then 1 (w2, U2),
// then 2 (w3, U3), then the angles.  The loops run U1, U2 outside and U3 inside, so for a fixed (U1, U2)
// the reduction over dimensions 0 and 1 of one w3-node "plane" (8 values, one per angle corner) is shared
// by every U3 whose cell touches that node: the two planes of the current w3 cell are kept in registers
// and recomputed only when U3 moves to another cell.  Same operations in the same order as the
// 64-corner evaluation — bit-identical to oracle_dense6_run.
#include <cstdint>
#include <cstring>
#include <string>
#include <vector>

#include <cuda_runtime.h>

#include "bellman.h"
#include "bellman_internal.h"

namespace bellman {
void set_global_error(const std::string &msg);   // bellman_api.cu: text behind bellman_last_error(NULL)

namespace {

constexpr int D6_BLOCK = 128;
constexpr int D6_MAXU = 8;

struct Dense6Params {
    const double *grid[6], *rinv[6];
    double inv_h[6];
    int n[6];
    long long stride[6];
    long long S, S3;
    int nu, tiles0, tiles1;
    const double *w_next[3], *a_next[3], *gs, *r[3];
    const double *J_next;
    double *J_out;
    int32_t *idx_out;
};

// exact bin rule cell = clamp(#{ s[i] <= x } - 1, 0, n-2): uniform guess, then a short scan either way
__device__ __forceinline__ int locate6(const double *__restrict__ s, const double *__restrict__ rinv, int n, double inv_h,
                                       double x, double &t) {
    int cell = min(max(__double2int_rd((x - __ldg(s)) * inv_h), 0), n - 2);
    while (cell < n - 2 && __ldg(s + cell + 1) <= x) ++cell;
    while (cell > 0 && __ldg(s + cell) > x) --cell;
    t = (x - __ldg(s + cell)) * __ldg(rinv + cell);
    return cell;
}

__device__ __forceinline__ double lerp(double a, double b, double t) { return fma(t, b - a, a); }

// dimensions 0 and 1 reduced at w3 node `node2`: one value per angle corner (bit m: +1 in dims 3, 4, 5)
__device__ __forceinline__ void plane(const Dense6Params &p, const double *__restrict__ J, long long o01, long long oa,
                                      int node2, double t0, double t1, double (&out)[8]) {
    const long long base = o01 + (long long)node2 * p.stride[2] + oa;
#pragma unroll
    for (int m = 0; m < 8; ++m) {
        const long long o = base + ((m & 1) ? p.stride[3] : 0) + ((m & 2) ? p.stride[4] : 0) + ((m & 4) ? p.stride[5] : 0);
        const double v00 = __ldg(J + o), v10 = __ldg(J + o + 1);
        const double v01 = __ldg(J + o + p.stride[1]), v11 = __ldg(J + o + p.stride[1] + 1);
        out[m] = lerp(lerp(v00, v10, t0), lerp(v01, v11, t0), t1);
    }
}

// a[u] without dynamic register indexing for small arrays
template <int MU, class T> __device__ __forceinline__ T pick(const T (&a)[MU], int u) {
    if (MU <= 4) {
        T v = a[0];
#pragma unroll
        for (int k = 1; k < MU; ++k) v = (u == k) ? a[k] : v;
        return v;
    }
    return a[u];
}

// NU > 0: compile-time control count (unrolled U3 loop); NU = 0: run-time p.nu <= D6_MAXU.
// (32-bit element offsets with the eight angle-corner offsets precomputed per state were measured: 5.4 ms
// against 5.1 ms on the 24^3 x 10^3 mesh — no gain, the kernel waits on L1 / L2 latency, not on address arithmetic;
// 5 CTAs per SM at 96 registers (__launch_bounds__(128, 5), 20 instead of 16 warps): 5.6 ms — the few spills cost more.)
template <int NU>
__global__ void __launch_bounds__(D6_BLOCK) k_stage_dense6(const __grid_constant__ Dense6Params p) {
    // a CTA is an 8 x 4 x 4 tile of (w1, w2, w3) at one angle triple (blockIdx.y = yaw + n3*pitch, blockIdx.z =
    // roll): the threads of a tile gather from the same few w-neighbourhoods of 8 angle corners, so most of
    // the 64-corner traffic is served by L1, and consecutive tiles (blockIdx.x) stay on one angle triple
    const int i0 = (blockIdx.x % p.tiles0) * 8 + (threadIdx.x & 7);
    const int i1 = ((blockIdx.x / p.tiles0) % p.tiles1) * 4 + ((threadIdx.x >> 3) & 3);
    const int i2 = (blockIdx.x / (p.tiles0 * p.tiles1)) * 4 + (threadIdx.x >> 5);
    if (i0 >= p.n[0] || i1 >= p.n[1] || i2 >= p.n[2]) return;
    const long long s3 = i0 + (long long)p.n[0] * (i1 + (long long)p.n[1] * i2);
    const long long s = s3 + p.S3 * (blockIdx.y + (long long)p.n[3] * p.n[4] * blockIdx.z);
    const int nu = NU > 0 ? NU : p.nu;
    constexpr int MU = NU > 0 ? NU : D6_MAXU;
    int cw[3][MU], ca[3];
    double tw[3][MU], ta[3];
#pragma unroll
    for (int d = 0; d < 3; ++d) {
#pragma unroll
        for (int u = 0; u < MU; ++u)
            if (u < nu) cw[d][u] = locate6(p.grid[d], p.rinv[d], p.n[d], p.inv_h[d], __ldg(p.w_next[d] + (size_t)u * p.S3 + s3), tw[d][u]);
        ca[d] = locate6(p.grid[3 + d], p.rinv[3 + d], p.n[3 + d], p.inv_h[3 + d], __ldg(p.a_next[d] + s), ta[d]);
    }
    const long long oa = ca[0] * p.stride[3] + ca[1] * p.stride[4] + ca[2] * p.stride[5];
    const double gs = __ldg(p.gs + s);
    const double *__restrict__ J = p.J_next;
    double best = __longlong_as_double(0x7ff0000000000000LL);
    int arg = 0;
    // U1 and U2 loops stay rolled: fully unrolled, the 27 combinations with their plane evaluations are
    // 350 KB of code per kernel, far beyond the instruction caches
#pragma unroll 1
    for (int u1 = 0; u1 < nu; ++u1) {
        const double g1 = gs + __ldg(p.r[0] + u1);
#pragma unroll 1
        for (int u2 = 0; u2 < nu; ++u2) {
            const double g2 = g1 + __ldg(p.r[1] + u2);
            const long long o01 = pick<MU>(cw[0], u1) + pick<MU>(cw[1], u2) * p.stride[1];
            const double t0 = pick<MU>(tw[0], u1), t1 = pick<MU>(tw[1], u2);
            double lo[8], hi[8];
            int have = -2;                                   // w3 cell whose two planes are in lo / hi
#pragma unroll
            for (int u3 = 0; u3 < MU; ++u3) {
                if (u3 >= nu) break;
                const int c2 = cw[2][u3];
                if (c2 != have) {
                    if (c2 == have + 1) {
#pragma unroll
                        for (int m = 0; m < 8; ++m) lo[m] = hi[m];
                    } else if (c2 == have - 1) {
#pragma unroll
                        for (int m = 0; m < 8; ++m) hi[m] = lo[m];
                    }
                    if (c2 != have + 1) plane(p, J, o01, oa, c2, t0, t1, lo);
                    if (c2 != have - 1 || have < 0) plane(p, J, o01, oa, c2 + 1, t0, t1, hi);
                    have = c2;
                }
                const double t2 = tw[2][u3];
                double v[8];
#pragma unroll
                for (int m = 0; m < 8; ++m) v[m] = lerp(lo[m], hi[m], t2);          // dimension 2
#pragma unroll
                for (int m = 0; m < 4; ++m) v[m] = lerp(v[2 * m], v[2 * m + 1], ta[0]);   // yaw
#pragma unroll
                for (int m = 0; m < 2; ++m) v[m] = lerp(v[2 * m], v[2 * m + 1], ta[1]);   // pitch
                const double val = lerp(v[0], v[1], ta[2]);                               // roll
                const double tot = (g2 + __ldg(p.r[2] + u3)) + val;
                if (tot < best) { best = tot; arg = (u1 * nu + u2) * nu + u3; }
            }
        }
    }
    p.J_out[s] = best;
    p.idx_out[s] = arg;
}

// Solver_attitude.get_optimal_path (attitude-control/Solver_attitude.m:1487-1530), the consumer of the 6-D
// policy: one thread per initial state.  Per step quat2angle([X7 X6 X5 X4]) (input normalised), the
// 'nearest' node of (w1 w2 w3 yaw pitch roll), the three torque levels stored there, then
// next_stage_states(.., 'taylor') (:1339-1371).  atan2 / asin are CUDA's: parity with the oracle is a tolerance.
struct Roll6Params {
    const double *grid[6];
    int n[6];
    long long stride[6];
    int nu, n_steps, batch;
    const int32_t *idx;
    const double *u_values, *x0;
    double J1, J2, J3, h;
    double *X_out, *U_out;
};
__device__ __forceinline__ int nearest6(const double *__restrict__ s, int n, double x) {
    int lo = 0, hi = n;                                   // count of s[i] <= x
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (__ldg(s + mid) <= x) lo = mid + 1; else hi = mid;
    }
    const int cell = min(max(lo - 1, 0), n - 2);
    return cell + ((x - __ldg(s + cell)) >= (__ldg(s + cell + 1) - x) ? 1 : 0);    // exact midpoint -> upper node
}
__global__ void __launch_bounds__(64) k_rollout_attitude6(const __grid_constant__ Roll6Params p) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= p.batch) return;
    double X[7];
#pragma unroll
    for (int k = 0; k < 7; ++k) X[k] = p.x0[(size_t)b * 7 + k];
    double *Xo = p.X_out + (size_t)b * 7 * (p.n_steps + 1), *Uo = p.U_out + (size_t)b * 3 * p.n_steps;
#pragma unroll
    for (int k = 0; k < 7; ++k) Xo[k] = X[k];
    const double J1 = p.J1, J2 = p.J2, J3 = p.J3, h = p.h;
    const int nu = p.nu;
    for (int ks = 0; ks < p.n_steps; ++ks) {
        const double qm = sqrt(((X[6] * X[6] + X[5] * X[5]) + X[4] * X[4]) + X[3] * X[3]);
        const double q0 = X[6] / qm, q1 = X[5] / qm, q2 = X[4] / qm, q3 = X[3] / qm;
        const double xq[6] = {X[0], X[1], X[2],
                              atan2(2 * (q1 * q2 + q0 * q3), ((q0 * q0 + q1 * q1) - q2 * q2) - q3 * q3),
                              asin(-2 * (q1 * q3 - q0 * q2)),
                              atan2(2 * (q2 * q3 + q0 * q1), ((q0 * q0 - q1 * q1) - q2 * q2) + q3 * q3)};
        long long o = 0;
#pragma unroll
        for (int k = 0; k < 6; ++k) o += nearest6(p.grid[k], p.n[k], xq[k]) * p.stride[k];
        const int c = __ldg(p.idx + o);
        const double U[3] = {__ldg(p.u_values + c / (nu * nu)), __ldg(p.u_values + (c / nu) % nu), __ldg(p.u_values + c % nu)};
#pragma unroll
        for (int k = 0; k < 3; ++k) Uo[(size_t)ks * 3 + k] = U[k];
        const double x1 = X[0], x2 = X[1], x3 = X[2], x4 = X[3], x5 = X[4], x6 = X[5], x7 = X[6];
        double d[7];
        d[0] = (J2 - J3) / J1 * x2 * x3 + U[0] / J1;
        d[1] = (J3 - J1) / J2 * x3 * x1 + U[1] / J2;
        d[2] = (J1 - J2) / J3 * x1 * x2 + U[2] / J3;
        d[3] = 0.5 * ((x3 * x5 - x2 * x6) + x1 * x7);
        d[4] = 0.5 * ((-x3 * x4 + x1 * x6) + x2 * x7);
        d[5] = 0.5 * ((x2 * x4 - x1 * x5) + x3 * x7);
        d[6] = 0.5 * ((-x1 * x4 - x2 * x5) - x3 * x6);
#pragma unroll
        for (int k = 0; k < 7; ++k) X[k] = X[k] + h * d[k];
        const double qs = sqrt(((X[3] * X[3] + X[4] * X[4]) + X[5] * X[5]) + X[6] * X[6]);
#pragma unroll
        for (int k = 3; k < 7; ++k) X[k] = X[k] / qs;
#pragma unroll
        for (int k = 0; k < 7; ++k) Xo[(size_t)(ks + 1) * 7 + k] = X[k];
    }
}

struct DevBuf {
    std::vector<void *> ptrs;
    ~DevBuf() { for (void *q : ptrs) cudaFree(q); }
    template <class T> cudaError_t alloc(T **out, size_t count) {
        void *q = nullptr;
        const cudaError_t e = cudaMalloc(&q, sizeof(T) * count);
        if (e == cudaSuccess) ptrs.push_back(q);
        *out = static_cast<T *>(q);
        return e;
    }
};

}  // namespace
}  // namespace bellman

using namespace bellman;

extern "C" int bellman_dense6_run(const bellman_dense6_desc *d, int32_t n_stages, const double *J_N, double *J_out,
                                  int32_t *idx_out, float *ms_out) {
    auto fail = [](int code, const std::string &m) { set_global_error(m); return code; };
    if (!d || !J_out || !idx_out) return fail(BELLMAN_ERR_BAD_ARG, "bellman_dense6_run: null argument");
    if (d->struct_size != (int32_t)sizeof(bellman_dense6_desc)) return fail(BELLMAN_ERR_BAD_ARG, "bellman_dense6_desc.struct_size mismatch");
    if (n_stages < 1) return fail(BELLMAN_ERR_BAD_ARG, "n_stages must be >= 1");
    if (d->nu < 1 || d->nu > D6_MAXU) return fail(BELLMAN_ERR_BAD_ARG, "nu must be in 1..8");
    Dense6Params p;
    std::memset(&p, 0, sizeof(p));
    long long S = 1;
    for (int k = 0; k < 6; ++k) {
        if (d->n[k] < 2 || !d->grid[k]) return fail(BELLMAN_ERR_BAD_ARG, "every dimension needs a grid of >= 2 points");
        for (int i = 1; i < d->n[k]; ++i)
            if (!(d->grid[k][i] > d->grid[k][i - 1])) return fail(BELLMAN_ERR_BAD_ARG, "grid must be strictly increasing");
        p.n[k] = d->n[k];
        p.stride[k] = S;
        S *= d->n[k];
    }
    for (int k = 0; k < 3; ++k)
        if (!d->w_next[k] || !d->a_next[k] || !d->r[k]) return fail(BELLMAN_ERR_BAD_ARG, "missing table");
    if (!d->gs) return fail(BELLMAN_ERR_BAD_ARG, "missing table");
    p.S = S;
    p.S3 = (long long)d->n[0] * d->n[1] * d->n[2];
    p.nu = d->nu;
    int dev = d->device;
    cudaError_t e = dev >= 0 ? cudaSetDevice(dev) : cudaGetDevice(&dev);
    if (e != cudaSuccess) return fail(BELLMAN_ERR_CUDA, std::string("no usable CUDA device (there is no CPU fallback): ") + cudaGetErrorString(e));
    DevBuf buf;
    cudaStream_t st = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    auto cleanup = [&]() { if (ev0) cudaEventDestroy(ev0); if (ev1) cudaEventDestroy(ev1); if (st) cudaStreamDestroy(st); };
#define D6(expr) do { cudaError_t _e = (expr); if (_e != cudaSuccess) { cleanup(); \
        return fail(_e == cudaErrorMemoryAllocation ? BELLMAN_ERR_OOM : BELLMAN_ERR_CUDA, cudaGetErrorString(_e)); } } while (0)
    D6(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
    D6(cudaEventCreate(&ev0));
    D6(cudaEventCreate(&ev1));
    for (int k = 0; k < 6; ++k) {
        double *g = nullptr, *ri = nullptr;
        std::vector<double> rinv((size_t)d->n[k] - 1);
        for (int i = 0; i + 1 < d->n[k]; ++i) rinv[i] = 1.0 / (d->grid[k][i + 1] - d->grid[k][i]);
        D6(buf.alloc(&g, (size_t)d->n[k]));
        D6(buf.alloc(&ri, rinv.size()));
        D6(cudaMemcpyAsync(g, d->grid[k], sizeof(double) * d->n[k], cudaMemcpyHostToDevice, st));
        D6(cudaMemcpyAsync(ri, rinv.data(), sizeof(double) * rinv.size(), cudaMemcpyHostToDevice, st));
        D6(cudaStreamSynchronize(st));               // rinv is a local
        p.grid[k] = g; p.rinv[k] = ri;
        p.inv_h[k] = (double)(d->n[k] - 1) / (d->grid[k][d->n[k] - 1] - d->grid[k][0]);
    }
    for (int k = 0; k < 3; ++k) {
        double *w = nullptr, *a = nullptr, *r = nullptr;
        D6(buf.alloc(&w, (size_t)d->nu * p.S3));
        D6(buf.alloc(&a, (size_t)S));
        D6(buf.alloc(&r, (size_t)d->nu));
        D6(cudaMemcpyAsync(w, d->w_next[k], sizeof(double) * d->nu * p.S3, cudaMemcpyHostToDevice, st));
        D6(cudaMemcpyAsync(a, d->a_next[k], sizeof(double) * S, cudaMemcpyHostToDevice, st));
        D6(cudaMemcpyAsync(r, d->r[k], sizeof(double) * d->nu, cudaMemcpyHostToDevice, st));
        p.w_next[k] = w; p.a_next[k] = a; p.r[k] = r;
    }
    double *gs = nullptr, *A = nullptr, *B = nullptr;
    int32_t *idx = nullptr;
    D6(buf.alloc(&gs, (size_t)S));
    D6(buf.alloc(&A, (size_t)S));
    D6(buf.alloc(&B, (size_t)S));
    D6(buf.alloc(&idx, (size_t)S));
    D6(cudaMemcpyAsync(gs, d->gs, sizeof(double) * S, cudaMemcpyHostToDevice, st));
    if (J_N) D6(cudaMemcpyAsync(A, J_N, sizeof(double) * S, cudaMemcpyHostToDevice, st));
    else D6(cudaMemsetAsync(A, 0, sizeof(double) * S, st));
    p.gs = gs; p.idx_out = idx;
    p.tiles0 = (d->n[0] + 7) / 8;
    p.tiles1 = (d->n[1] + 3) / 4;
    if ((long long)d->n[3] * d->n[4] > 65535 || d->n[5] > 65535) { cleanup(); return fail(BELLMAN_ERR_BAD_ARG, "angle mesh too fine for the launch grid (n3*n4 and n5 must be <= 65535)"); }
    const dim3 grid((unsigned)(p.tiles0 * p.tiles1 * ((d->n[2] + 3) / 4)), (unsigned)(d->n[3] * d->n[4]), (unsigned)d->n[5]);
    D6(cudaEventRecord(ev0, st));
    for (int k = 0; k < n_stages; ++k) {
        p.J_next = A; p.J_out = B;
        if (d->nu == 3) k_stage_dense6<3><<<grid, D6_BLOCK, 0, st>>>(p);
        else k_stage_dense6<0><<<grid, D6_BLOCK, 0, st>>>(p);
        D6(cudaGetLastError());
        double *t = A; A = B; B = t;
    }
    D6(cudaEventRecord(ev1, st));
    D6(cudaMemcpyAsync(J_out, A, sizeof(double) * S, cudaMemcpyDeviceToHost, st));
    D6(cudaMemcpyAsync(idx_out, idx, sizeof(int32_t) * S, cudaMemcpyDeviceToHost, st));
    D6(cudaStreamSynchronize(st));
    if (ms_out) { float ms = 0; D6(cudaEventElapsedTime(&ms, ev0, ev1)); *ms_out = ms; }
#undef D6
    cleanup();
    return BELLMAN_OK;
}

extern "C" int bellman_rollout_attitude6(const bellman_dense6_desc *d, const int32_t *idx, const double *u_values,
                                         const double *J123, double h, int32_t n_steps, const double *x0, int32_t batch,
                                         double *X_out, double *U_out) {
    auto fail = [](int code, const std::string &m) { set_global_error(m); return code; };
    if (!d || !idx || !u_values || !J123 || !x0 || !X_out || !U_out || batch < 1 || n_steps < 1)
        return fail(BELLMAN_ERR_BAD_ARG, "bellman_rollout_attitude6: null or non-positive argument");
    if (d->struct_size != (int32_t)sizeof(bellman_dense6_desc)) return fail(BELLMAN_ERR_BAD_ARG, "bellman_dense6_desc.struct_size mismatch");
    if (d->nu < 1 || d->nu > D6_MAXU) return fail(BELLMAN_ERR_BAD_ARG, "nu must be in 1..8");
    if (!(J123[0] > 0 && J123[1] > 0 && J123[2] > 0) || !(h > 0)) return fail(BELLMAN_ERR_BAD_ARG, "J1, J2, J3 and h must be positive");
    Roll6Params p;
    std::memset(&p, 0, sizeof(p));
    long long S = 1;
    for (int k = 0; k < 6; ++k) {
        if (d->n[k] < 2 || !d->grid[k]) return fail(BELLMAN_ERR_BAD_ARG, "every dimension needs a grid of >= 2 points");
        p.n[k] = d->n[k];
        p.stride[k] = S;
        S *= d->n[k];
    }
    const int C = d->nu * d->nu * d->nu;
    for (long long i = 0; i < S; ++i)
        if (idx[i] < 0 || idx[i] >= C) return fail(BELLMAN_ERR_BAD_ARG, "policy index out of range");
    int dev = d->device;
    cudaError_t e = dev >= 0 ? cudaSetDevice(dev) : cudaGetDevice(&dev);
    if (e != cudaSuccess) return fail(BELLMAN_ERR_CUDA, std::string("no usable CUDA device (there is no CPU fallback): ") + cudaGetErrorString(e));
    DevBuf buf;
    cudaStream_t st = nullptr;
    auto cleanup = [&]() { if (st) cudaStreamDestroy(st); };
#define R6(expr) do { cudaError_t _e = (expr); if (_e != cudaSuccess) { cleanup(); \
        return fail(_e == cudaErrorMemoryAllocation ? BELLMAN_ERR_OOM : BELLMAN_ERR_CUDA, cudaGetErrorString(_e)); } } while (0)
    R6(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
    for (int k = 0; k < 6; ++k) {
        double *g = nullptr;
        R6(buf.alloc(&g, (size_t)d->n[k]));
        R6(cudaMemcpyAsync(g, d->grid[k], sizeof(double) * d->n[k], cudaMemcpyHostToDevice, st));
        p.grid[k] = g;
    }
    int32_t *d_idx = nullptr;
    double *d_u = nullptr, *d_x0 = nullptr, *d_X = nullptr, *d_U = nullptr;
    R6(buf.alloc(&d_idx, (size_t)S));
    R6(buf.alloc(&d_u, (size_t)d->nu));
    R6(buf.alloc(&d_x0, 7 * (size_t)batch));
    R6(buf.alloc(&d_X, 7 * (size_t)(n_steps + 1) * batch));
    R6(buf.alloc(&d_U, 3 * (size_t)n_steps * batch));
    R6(cudaMemcpyAsync(d_idx, idx, sizeof(int32_t) * S, cudaMemcpyHostToDevice, st));
    R6(cudaMemcpyAsync(d_u, u_values, sizeof(double) * d->nu, cudaMemcpyHostToDevice, st));
    R6(cudaMemcpyAsync(d_x0, x0, sizeof(double) * 7 * batch, cudaMemcpyHostToDevice, st));
    p.nu = d->nu; p.n_steps = n_steps; p.batch = batch;
    p.idx = d_idx; p.u_values = d_u; p.x0 = d_x0; p.X_out = d_X; p.U_out = d_U;
    p.J1 = J123[0]; p.J2 = J123[1]; p.J3 = J123[2]; p.h = h;
    k_rollout_attitude6<<<(batch + 63) / 64, 64, 0, st>>>(p);
    R6(cudaGetLastError());
    R6(cudaMemcpyAsync(X_out, d_X, sizeof(double) * 7 * (size_t)(n_steps + 1) * batch, cudaMemcpyDeviceToHost, st));
    R6(cudaMemcpyAsync(U_out, d_U, sizeof(double) * 3 * (size_t)n_steps * batch, cudaMemcpyDeviceToHost, st));
    R6(cudaStreamSynchronize(st));
#undef R6
    cleanup();
    return BELLMAN_OK;
}
