// bellman_mex.cpp — thin MEX gateway over the C ABI of libbellman.so (include/bellman.h).
//
//   h      = bellman_mex('create', desc)              desc: struct, see Dynamic_Solver.m (build_desc)
//            bellman_mex('set_J', h, J)               J: S x P double (terminal cost / resume), [] = zeros
//            bellman_mex('run', h, n_stages, opts)    opts: struct with optional fields kernel,
//                                                     check_period, check_tol, use_graph, sync_each_stage
//   k      = bellman_mex('current_stage', h)
//   J      = bellman_mex('get_J', h, stage)           S_own x P double, column-major like F.Values(:)
//   idx    = bellman_mex('get_idx', h, stage)         S_own x P int32, 1-BASED (MATLAB's min index)
//   log    = bellman_mex('check_log', h)              3 x n double: stage; sum(J); sum(idx)
//   s      = bellman_mex('stats', h)                  struct: ms, launches, ms_exchange, kernel
//   [X,U]  = bellman_mex('rollout', h, N, A, B, u_values, X0, mode, ssu_stage)   X0: 2 x batch,
//                                                     X: 2 x (N*batch), U: N x batch
//   pd     = bellman_mex('plan', desc, n)             slab dimension (1-based) with the smallest halo for n slabs
//   r      = bellman_mex('owned_range', h)            [own_lo own_hi) along the slab dimension, 0-based
//            bellman_mex('group_init', hs)            hs: uint64 vector of handles = slabs 1..n of ONE problem
//            bellman_mex('group_run', hs, n_stages, opts)   all slabs together, from this one host thread
//   [X,id,w] = bellman_mex('rollout_orbit', h, stage, o, u_values, Y0)   Solver_position.get_optimal_path:
//                                                     o: struct n_steps, stride_out, mu, R0, V0, h, tol; Y0: 6 x batch;
//                                                     X: 6 x ((n_steps/stride_out+1)*batch), id (1-based): 3 x (n_out*batch)
//   [X,F,FM,w] = bellman_mex('rollout_pos_att', [hx hy hz], stages, o, fx, fy, fz, Y0)   Solver_pos_att.get_optimal_path:
//                                                     o: struct n_steps, stride_out, mu, R0, V0, h, rtol, atol, InertiaM (3x3),
//                                                     Mass, T_dist; f*: C_ch x 4 thruster levels (columns f0 f1 f6 f7 of the channel); Y0: 13 x batch;
//                                                     X: 13 x ((n_out+1)*batch), F: 12 x (n_out*batch), FM: 6 x (n_out*batch)
//   [X,id,w] = bellman_mex('rollout_attitude', h, stage, o, u_values, Y0)   Solver_attitude.get_optimal_path_simplified_testode45:
//                                                     o: struct n_steps, stride_out, h, rtol, atol, InertiaM; Y0: 7 x batch
//   [J,id,ms] = bellman_mex('dense6_run', d6, n_stages, J_N)   Solver_attitude.run (the coupled 6-D sweep):
//                                                     d6: struct n (1x6), nu, device, grid {6}, w_next {3} (S3 x nu each),
//                                                     a_next {3} (S x 1), gs (S x 1), r {3} (nu x 1); J_N [] = zeros;
//                                                     J: S x 1, id (1-based, (u1-1)*nu^2 + (u2-1)*nu + u3): S x 1
//   [X,U]  = bellman_mex('rollout_attitude6', d6, id, u_values, J123, h, n_steps, X0)   Solver_attitude.get_optimal_path
//                                                     under the 6-D policy id (S x 1, 1-based as 'dense6_run' returns it);
//                                                     d6 needs n, nu, grid, device; X0: 7 x batch; X: 7 x ((n_steps+1)*batch),
//                                                     U: 3 x (n_steps*batch)
//            bellman_mex('destroy', h)
//   v      = bellman_mex('version')
//
// Arrays cross as column-major double / int32 with no transposition.  Library error codes become
// mexErrMsgIdAndTxt('bellman:<code>', message).  The MEX file is locked while handles are alive
// and frees them at exit.  There is no CPU fallback: without a B200 'create' raises bellman:CUDA.
//
// Build (on a machine with MATLAB):  mex -R2018a bellman_mex.cpp -I../../include -L.. -lbellman
#include <cstdint>
#include <cstring>
#include <set>
#include <string>
#include <vector>

#include "mex.h"

#include "bellman.h"

static std::set<bellman_handle *> g_live;

static void at_exit() {
    for (bellman_handle *h : g_live) bellman_destroy(h);
    g_live.clear();
}

static const char *code_name(int rc) {
    switch (rc) {
        case BELLMAN_ERR_BAD_ARG: return "bellman:BAD_ARG";
        case BELLMAN_ERR_CUDA: return "bellman:CUDA";
        case BELLMAN_ERR_NCCL: return "bellman:NCCL";
        case BELLMAN_ERR_OOM: return "bellman:OOM";
        case BELLMAN_ERR_NOT_RUN: return "bellman:NOT_RUN";
        case BELLMAN_ERR_STATE: return "bellman:STATE";
        default: return "bellman:UNKNOWN";
    }
}

static void check(int rc, bellman_handle *h) {
    if (rc != BELLMAN_OK) mexErrMsgIdAndTxt(code_name(rc), "%s", bellman_last_error(h));
}

static bellman_handle *get_handle(const mxArray *a) {
    if (!mxIsClass(a, "uint64") || mxGetNumberOfElements(a) != 1)
        mexErrMsgIdAndTxt("bellman:BAD_ARG", "handle must be a uint64 scalar");
    bellman_handle *h = reinterpret_cast<bellman_handle *>(*static_cast<uint64_t *>(mxGetData(a)));
    if (!g_live.count(h)) mexErrMsgIdAndTxt("bellman:BAD_ARG", "stale or foreign handle");
    return h;
}

static double field_scalar(const mxArray *s, const char *name, double dflt) {
    const mxArray *f = mxGetField(s, 0, name);
    return (f && !mxIsEmpty(f)) ? mxGetScalar(f) : dflt;
}

// cell field `name`, entry d -> pointer to doubles (or NULL when the entry is []), size checked
static const double *cell_table(const mxArray *s, const char *name, int d, size_t want, bool optional) {
    const mxArray *c = mxGetField(s, 0, name);
    if (!c || !mxIsCell(c) || mxGetNumberOfElements(c) <= (size_t)d)
        mexErrMsgIdAndTxt("bellman:BAD_ARG", "desc.%s must be a cell with one entry per dimension", name);
    const mxArray *e = mxGetCell(c, d);
    if (!e || mxIsEmpty(e)) {
        if (!optional) mexErrMsgIdAndTxt("bellman:BAD_ARG", "desc.%s{%d} is empty", name, d + 1);
        return nullptr;
    }
    if (!mxIsDouble(e) || mxGetNumberOfElements(e) != want)
        mexErrMsgIdAndTxt("bellman:BAD_ARG", "desc.%s{%d} must be double with %d elements", name, d + 1, (int)want);
    return mxGetPr(e);
}

// desc struct -> bellman_desc (pointers into the MATLAB arrays: valid for the duration of the call)
static void fill_desc(const mxArray *s, bellman_desc &d) {
    std::memset(&d, 0, sizeof(d));
    d.struct_size = (int32_t)sizeof(d);
    const mxArray *nn = mxGetField(s, 0, "n");
    if (!nn || !mxIsDouble(nn)) mexErrMsgIdAndTxt("bellman:BAD_ARG", "desc.n must be a double row vector");
    d.D = (int32_t)mxGetNumberOfElements(nn);
    if (d.D < 2 || d.D > BELLMAN_MAX_DIM) mexErrMsgIdAndTxt("bellman:BAD_ARG", "desc.n must have 2..4 entries");
    for (int k = 0; k < d.D; ++k) d.n[k] = (int32_t)mxGetPr(nn)[k];
    d.C = (int32_t)field_scalar(s, "C", 0);
    d.P = (int32_t)field_scalar(s, "P", 1);
    d.N = (int32_t)field_scalar(s, "N", 0);
    const mxArray *sa = mxGetField(s, 0, "src_a"), *sb = mxGetField(s, 0, "src_b"), *qo = mxGetField(s, 0, "q_order");
    if (!sa || !sb || !qo) mexErrMsgIdAndTxt("bellman:BAD_ARG", "desc needs src_a, src_b, q_order (1-based, 0 = absent)");
    for (int k = 0; k < d.D; ++k) {
        d.src_a[k] = (int32_t)mxGetPr(sa)[k] - 1;          // MATLAB dims are 1-based
        d.src_b[k] = (int32_t)mxGetPr(sb)[k] - 1;          // 0 -> -1 = absent
        d.q_order[k] = (int32_t)mxGetPr(qo)[k] - 1;
    }
    for (int k = 0; k < d.D; ++k) {
        const size_t P = (size_t)d.P;
        d.grid[k] = cell_table(s, "grid", k, P * d.n[k], false);
        d.q[k] = cell_table(s, "q", k, P * d.n[k], false);
        if (d.src_a[k] < 0 || d.src_a[k] >= d.D) mexErrMsgIdAndTxt("bellman:BAD_ARG", "src_a out of range");
        d.Ta[k] = cell_table(s, "Ta", k, P * d.n[d.src_a[k]], false);
        d.Tb[k] = d.src_b[k] >= 0 ? cell_table(s, "Tb", k, P * d.n[d.src_b[k]], false) : nullptr;
        d.Tc[k] = cell_table(s, "Tc", k, P * d.C, true);
    }
    const mxArray *r = mxGetField(s, 0, "r");
    if (!r || !mxIsDouble(r) || mxGetNumberOfElements(r) != (size_t)d.P * d.C)
        mexErrMsgIdAndTxt("bellman:BAD_ARG", "desc.r must be C x P double");
    d.r = mxGetPr(r);
    d.store_J_all = (int32_t)field_scalar(s, "store_J_all", 0);
    d.store_idx_all = (int32_t)field_scalar(s, "store_idx_all", 0);
    d.device = (int32_t)field_scalar(s, "device", -1);
    d.part_dim = (int32_t)field_scalar(s, "part_dim", 0) - 1;
    d.rank = (int32_t)field_scalar(s, "rank", 0);
    d.nranks = (int32_t)field_scalar(s, "nranks", 1);
    d.idx_bytes = (int32_t)field_scalar(s, "idx_bytes", 0);     // device storage of the argmin: 0/4 int32, 2 uint16, 1 uint8
}

static void cmd_create(int nlhs, mxArray *plhs[], int nrhs, const mxArray *prhs[]) {
    if (nrhs != 2 || !mxIsStruct(prhs[1])) mexErrMsgIdAndTxt("bellman:BAD_ARG", "usage: h = bellman_mex('create', desc)");
    bellman_desc d;
    fill_desc(prhs[1], d);
    bellman_handle *h = nullptr;
    check(bellman_create(&d, &h), nullptr);
    if (g_live.empty()) { mexLock(); mexAtExit(at_exit); }
    g_live.insert(h);
    plhs[0] = mxCreateNumericMatrix(1, 1, mxUINT64_CLASS, mxREAL);
    *static_cast<uint64_t *>(mxGetData(plhs[0])) = reinterpret_cast<uint64_t>(h);
    (void)nlhs;
}

// what the gateway needs to size and check every array argument of a handle
struct Shape { size_t S_own, P, S_global; int N, C, D; };

// a real double array with exactly `want` elements (the reference hands `single` arrays around,
// e.g. zeros(...,'single') in Solver_pos_att.m:265 — those must be cast by the caller, not read as
// doubles here)
static const double *need_doubles(const mxArray *a, size_t want, const char *what) {
    if (!a || !mxIsDouble(a) || mxIsComplex(a) || mxGetNumberOfElements(a) != want)
        mexErrMsgIdAndTxt("bellman:BAD_ARG", "%s must be a real double array with %d elements", what, (int)want);
    return mxGetPr(a);
}
static void plant_opts_from(const mxArray *s, bellman_plant_opts &o) {
    std::memset(&o, 0, sizeof(o));
    o.struct_size = (int32_t)sizeof(o);
    o.n_steps = (int32_t)field_scalar(s, "n_steps", 0);
    o.stride_out = (int32_t)field_scalar(s, "stride_out", 1);
    o.mu = field_scalar(s, "mu", 398600.0);
    o.h = field_scalar(s, "h", 0.0);
    o.rtol = field_scalar(s, "rtol", 1e-3);
    o.atol = field_scalar(s, "atol", 1e-6);
    o.mass = field_scalar(s, "Mass", 0.0);
    o.t_dist = field_scalar(s, "T_dist", 0.0);
    const double *Im = need_doubles(mxGetField(s, 0, "InertiaM"), 9, "o.InertiaM");
    for (int k = 0; k < 9; ++k) o.inertia[k] = Im[k];
    const mxArray *r0 = mxGetField(s, 0, "R0"), *v0 = mxGetField(s, 0, "V0");
    if (r0 && v0) {
        const double *R0 = need_doubles(r0, 3, "o.R0"), *V0 = need_doubles(v0, 3, "o.V0");
        for (int k = 0; k < 3; ++k) { o.R0[k] = R0[k]; o.V0[k] = V0[k]; }
    }
    if (o.n_steps < 1 || o.stride_out < 1 || o.n_steps % o.stride_out) mexErrMsgIdAndTxt("bellman:BAD_ARG", "n_steps must be a positive multiple of stride_out");
}
static std::vector<std::pair<bellman_handle *, Shape>> g_shapes;

static Shape shape_of(bellman_handle *h) {
    for (auto &p : g_shapes) if (p.first == h) return p.second;
    mexErrMsgIdAndTxt("bellman:BAD_ARG", "unknown handle");
    return Shape{0, 0};
}

void mexFunction(int nlhs, mxArray *plhs[], int nrhs, const mxArray *prhs[]) {
    if (nrhs < 1 || !mxIsChar(prhs[0])) mexErrMsgIdAndTxt("bellman:BAD_ARG", "first argument must be a command string");
    char *c = mxArrayToString(prhs[0]);
    const std::string cmd(c);
    mxFree(c);

    if (cmd == "version") { plhs[0] = mxCreateDoubleScalar(bellman_version()); return; }
    if (cmd == "plan") {
        // pd = bellman_mex('plan', desc, n): the dimension whose slabs need the smallest halo (host-only)
        if (nrhs != 3 || !mxIsStruct(prhs[1])) mexErrMsgIdAndTxt("bellman:BAD_ARG", "usage: pd = bellman_mex('plan', desc, n)");
        bellman_desc d;
        fill_desc(prhs[1], d);
        const int n = (int)mxGetScalar(prhs[2]);
        if (n < 1 || n > 64) mexErrMsgIdAndTxt("bellman:BAD_ARG", "n must be 1..64");
        std::vector<bellman_slab> sl((size_t)n);
        int best = -1;
        double best_cost = 0.0;
        for (int pd = 0; pd < d.D; ++pd) {
            if (bellman_plan_slabs(&d, pd, n, sl.data()) != BELLMAN_OK) continue;
            double cost = 0.0;
            for (const bellman_slab &m : sl) {
                const double c = (double)((m.ext_hi - m.ext_lo) - (m.own_hi - m.own_lo)) / (double)(m.own_hi - m.own_lo);
                if (c > cost) cost = c;
            }
            if (best < 0 || cost < best_cost - 1e-9) { best = pd; best_cost = cost; }
        }
        if (best < 0) mexErrMsgIdAndTxt("bellman:BAD_ARG", "no dimension can be cut into %d slabs", n);
        plhs[0] = mxCreateDoubleScalar(best + 1);
        return;
    }
    if (cmd == "group_init" || cmd == "group_run") {
        if (nrhs < 2 || !mxIsClass(prhs[1], "uint64")) mexErrMsgIdAndTxt("bellman:BAD_ARG", "hs must be a uint64 vector of handles");
        const size_t n = mxGetNumberOfElements(prhs[1]);
        std::vector<bellman_handle *> hs(n);
        for (size_t k = 0; k < n; ++k) {
            hs[k] = reinterpret_cast<bellman_handle *>(static_cast<uint64_t *>(mxGetData(prhs[1]))[k]);
            if (!g_live.count(hs[k])) mexErrMsgIdAndTxt("bellman:BAD_ARG", "stale or foreign handle");
        }
        if (cmd == "group_init") { check(bellman_group_init(hs.data(), (int32_t)n), hs[0]); return; }
        if (nrhs < 3) mexErrMsgIdAndTxt("bellman:BAD_ARG", "usage: bellman_mex('group_run', hs, n_stages, opts)");
        bellman_run_opts o;
        std::memset(&o, 0, sizeof(o));
        o.struct_size = (int32_t)sizeof(o);
        if (nrhs > 3 && mxIsStruct(prhs[3])) {
            o.kernel = (int32_t)field_scalar(prhs[3], "kernel", 0);
            o.check_period = (int32_t)field_scalar(prhs[3], "check_period", 0);
            o.check_tol = field_scalar(prhs[3], "check_tol", 0.0);
        }
        const int rc = bellman_group_run(hs.data(), (int32_t)n, (int32_t)mxGetScalar(prhs[2]), &o);
        if (rc != BELLMAN_OK)
            for (bellman_handle *h : hs)
                if (bellman_last_error(h)[0]) check(rc, h);
        check(rc, hs[0]);
        return;
    }
    if (cmd == "dense6_run") {
        if (nrhs < 3 || !mxIsStruct(prhs[1])) mexErrMsgIdAndTxt("bellman:BAD_ARG", "usage: [J,id,ms] = bellman_mex('dense6_run', d6, n_stages, J_N)");
        const mxArray *s = prhs[1];
        bellman_dense6_desc d;
        std::memset(&d, 0, sizeof(d));
        d.struct_size = (int32_t)sizeof(d);
        const double *nn = need_doubles(mxGetField(s, 0, "n"), 6, "d6.n");
        size_t S = 1;
        for (int k = 0; k < 6; ++k) { d.n[k] = (int32_t)nn[k]; S *= (size_t)d.n[k]; }
        const size_t S3 = (size_t)d.n[0] * d.n[1] * d.n[2];
        d.nu = (int32_t)field_scalar(s, "nu", 0);
        d.device = (int32_t)field_scalar(s, "device", -1);
        if (d.nu < 1 || d.nu > 8) mexErrMsgIdAndTxt("bellman:BAD_ARG", "d6.nu must be in 1..8");
        for (int k = 0; k < 6; ++k) d.grid[k] = cell_table(s, "grid", k, (size_t)d.n[k], false);
        for (int k = 0; k < 3; ++k) {
            d.w_next[k] = cell_table(s, "w_next", k, S3 * (size_t)d.nu, false);
            d.a_next[k] = cell_table(s, "a_next", k, S, false);
            d.r[k] = cell_table(s, "r", k, (size_t)d.nu, false);
        }
        d.gs = need_doubles(mxGetField(s, 0, "gs"), S, "d6.gs");
        const double *JN = (nrhs > 3 && !mxIsEmpty(prhs[3])) ? need_doubles(prhs[3], S, "J_N") : nullptr;
        plhs[0] = mxCreateDoubleMatrix(S, 1, mxREAL);
        mxArray *id = mxCreateNumericMatrix(S, 1, mxINT32_CLASS, mxREAL);
        int32_t *pi = static_cast<int32_t *>(mxGetData(id));
        float ms = 0;
        const int rc = bellman_dense6_run(&d, (int32_t)mxGetScalar(prhs[2]), JN, mxGetPr(plhs[0]), pi, &ms);
        if (rc != BELLMAN_OK) mexErrMsgIdAndTxt(code_name(rc), "%s", bellman_last_error(nullptr));
        for (size_t k = 0; k < S; ++k) pi[k] += 1;
        if (nlhs > 1) plhs[1] = id; else mxDestroyArray(id);
        if (nlhs > 2) plhs[2] = mxCreateDoubleScalar((double)ms);
        return;
    }
    if (cmd == "rollout_attitude6") {
        if (nrhs < 8 || !mxIsStruct(prhs[1])) mexErrMsgIdAndTxt("bellman:BAD_ARG", "usage: [X,U] = bellman_mex('rollout_attitude6', d6, id, u_values, J123, h, n_steps, X0)");
        const mxArray *s = prhs[1];
        bellman_dense6_desc d;
        std::memset(&d, 0, sizeof(d));
        d.struct_size = (int32_t)sizeof(d);
        const double *nn = need_doubles(mxGetField(s, 0, "n"), 6, "d6.n");
        size_t S = 1;
        for (int k = 0; k < 6; ++k) { d.n[k] = (int32_t)nn[k]; S *= (size_t)d.n[k]; }
        d.nu = (int32_t)field_scalar(s, "nu", 0);
        d.device = (int32_t)field_scalar(s, "device", -1);
        if (d.nu < 1 || d.nu > 8) mexErrMsgIdAndTxt("bellman:BAD_ARG", "d6.nu must be in 1..8");
        for (int k = 0; k < 6; ++k) d.grid[k] = cell_table(s, "grid", k, (size_t)d.n[k], false);
        if (!mxIsClass(prhs[2], "int32") || mxGetNumberOfElements(prhs[2]) != S) mexErrMsgIdAndTxt("bellman:BAD_ARG", "id must be the int32 S-by-1 policy 'dense6_run' returns");
        std::vector<int32_t> idx(S);
        const int32_t *id1 = static_cast<const int32_t *>(mxGetData(prhs[2]));
        for (size_t k = 0; k < S; ++k) idx[k] = id1[k] - 1;
        const double *uv = need_doubles(prhs[3], (size_t)d.nu, "u_values"), *J123 = need_doubles(prhs[4], 3, "J123");
        const int32_t n_steps = (int32_t)mxGetScalar(prhs[6]);
        if (n_steps < 1) mexErrMsgIdAndTxt("bellman:BAD_ARG", "n_steps must be >= 1");
        const size_t batch = mxGetN(prhs[7]);
        need_doubles(prhs[7], 7 * batch, "X0 (7-by-batch)");
        plhs[0] = mxCreateDoubleMatrix(7, (size_t)(n_steps + 1) * batch, mxREAL);
        mxArray *U = mxCreateDoubleMatrix(3, (size_t)n_steps * batch, mxREAL);
        const int rc = bellman_rollout_attitude6(&d, idx.data(), uv, J123, mxGetScalar(prhs[5]), n_steps, mxGetPr(prhs[7]),
                                                 (int32_t)batch, mxGetPr(plhs[0]), mxGetPr(U));
        if (rc != BELLMAN_OK) mexErrMsgIdAndTxt(code_name(rc), "%s", bellman_last_error(nullptr));
        if (nlhs > 1) plhs[1] = U; else mxDestroyArray(U);
        return;
    }
    if (cmd == "rollout_pos_att") {
        if (nrhs < 8 || !mxIsClass(prhs[1], "uint64") || mxGetNumberOfElements(prhs[1]) != 3 || !mxIsStruct(prhs[3]))
            mexErrMsgIdAndTxt("bellman:BAD_ARG", "usage: [X,F,FM,w] = bellman_mex('rollout_pos_att', [hx hy hz], stages, o, fx, fy, fz, Y0)");
        bellman_handle *hs[3];
        int32_t stage[3];
        const double *fv[3];
        const double *st = need_doubles(prhs[2], 3, "stages");
        for (int k = 0; k < 3; ++k) {
            hs[k] = reinterpret_cast<bellman_handle *>(static_cast<uint64_t *>(mxGetData(prhs[1]))[k]);
            if (!g_live.count(hs[k])) mexErrMsgIdAndTxt("bellman:BAD_ARG", "stale or foreign handle");
            stage[k] = (int32_t)st[k];
            fv[k] = need_doubles(prhs[4 + k], 4 * (size_t)shape_of(hs[k]).C, "f_x / f_y / f_z (C-by-4 of the channel)");
        }
        bellman_plant_opts o;
        plant_opts_from(prhs[3], o);
        const size_t batch = mxGetN(prhs[7]);
        need_doubles(prhs[7], 13 * batch, "Y0 (13-by-batch)");
        const size_t n_out = (size_t)(o.n_steps / o.stride_out);
        plhs[0] = mxCreateDoubleMatrix(13, (n_out + 1) * batch, mxREAL);
        mxArray *F = mxCreateDoubleMatrix(12, n_out * batch, mxREAL);
        mxArray *FM = mxCreateDoubleMatrix(6, n_out * batch, mxREAL);
        mxArray *w = mxCreateNumericMatrix(1, batch, mxINT32_CLASS, mxREAL);
        check(bellman_rollout_pos_att(hs[0], hs[1], hs[2], stage, &o, fv[0], fv[1], fv[2], mxGetPr(prhs[7]), (int32_t)batch,
                                      mxGetPr(plhs[0]), mxGetPr(F), mxGetPr(FM), static_cast<int32_t *>(mxGetData(w))), hs[0]);
        if (nlhs > 1) plhs[1] = F; else mxDestroyArray(F);
        if (nlhs > 2) plhs[2] = FM; else mxDestroyArray(FM);
        if (nlhs > 3) plhs[3] = w; else mxDestroyArray(w);
        return;
    }
    if (cmd == "create") {
        cmd_create(nlhs, plhs, nrhs, prhs);
        bellman_handle *h = reinterpret_cast<bellman_handle *>(*static_cast<uint64_t *>(mxGetData(plhs[0])));
        // remember the owned shape for get_J / get_idx
        const mxArray *s = prhs[1];
        const mxArray *nn = mxGetField(s, 0, "n");
        bellman_slab sl;
        bellman_owned_range(h, &sl);
        const int D = (int)mxGetNumberOfElements(nn);
        const int pd = (int)field_scalar(s, "part_dim", 0) - 1;
        size_t S = 1, Sg = 1;
        for (int k = 0; k < D; ++k) {
            S *= (k == pd) ? (size_t)(sl.own_hi - sl.own_lo) : (size_t)mxGetPr(nn)[k];
            Sg *= (size_t)mxGetPr(nn)[k];
        }
        g_shapes.push_back({h, Shape{S, (size_t)field_scalar(s, "P", 1), Sg, (int)field_scalar(s, "N", 0),
                                     (int)field_scalar(s, "C", 0), D}});
        return;
    }
    if (nrhs < 2) mexErrMsgIdAndTxt("bellman:BAD_ARG", "missing handle");
    bellman_handle *h = get_handle(prhs[1]);
    const Shape sh = shape_of(h);

    if (cmd == "destroy") {
        bellman_destroy(h);
        g_live.erase(h);
        for (size_t i = 0; i < g_shapes.size(); ++i) if (g_shapes[i].first == h) { g_shapes.erase(g_shapes.begin() + i); break; }
        if (g_live.empty()) mexUnlock();
    } else if (cmd == "set_J") {
        const double *J = (nrhs > 2 && !mxIsEmpty(prhs[2])) ? need_doubles(prhs[2], sh.S_global * sh.P, "J") : nullptr;
        check(bellman_set_J(h, J), h);
    } else if (cmd == "run") {
        if (nrhs < 3) mexErrMsgIdAndTxt("bellman:BAD_ARG", "usage: bellman_mex('run', h, n_stages, opts)");
        bellman_run_opts o;
        std::memset(&o, 0, sizeof(o));
        o.struct_size = (int32_t)sizeof(o);
        if (nrhs > 3 && mxIsStruct(prhs[3])) {
            o.kernel = (int32_t)field_scalar(prhs[3], "kernel", 0);
            o.check_period = (int32_t)field_scalar(prhs[3], "check_period", 0);
            o.check_tol = field_scalar(prhs[3], "check_tol", 0.0);
            o.use_graph = (int32_t)field_scalar(prhs[3], "use_graph", 0);
            o.sync_each_stage = (int32_t)field_scalar(prhs[3], "sync_each_stage", 0);
        }
        check(bellman_run(h, (int32_t)mxGetScalar(prhs[2]), &o), h);
    } else if (cmd == "stage") {
        check(bellman_stage(h), h);
    } else if (cmd == "stage_host") {
        // [J, idx] = bellman_mex('stage_host', h, J_next)   one stage from / to MATLAB arrays, copies overlapped
        // with the kernel (J_next [] = continue from the device's J); idx is 1-based like min()'s second output
        const double *J = (nrhs > 2 && !mxIsEmpty(prhs[2])) ? need_doubles(prhs[2], sh.S_global * sh.P, "J_next") : nullptr;
        plhs[0] = mxCreateDoubleMatrix(sh.S_own, sh.P, mxREAL);
        mxArray *I = mxCreateNumericMatrix(sh.S_own, sh.P, mxINT32_CLASS, mxREAL);
        int32_t *p = static_cast<int32_t *>(mxGetData(I));
        check(bellman_stage_host(h, J, mxGetPr(plhs[0]), p, nullptr), h);
        for (size_t k = 0; k < sh.S_own * sh.P; ++k) p[k] += 1;
        if (nlhs > 1) plhs[1] = I; else mxDestroyArray(I);
    } else if (cmd == "owned_range") {
        bellman_slab sl;
        check(bellman_owned_range(h, &sl), h);
        plhs[0] = mxCreateDoubleMatrix(1, 2, mxREAL);
        mxGetPr(plhs[0])[0] = sl.own_lo;
        mxGetPr(plhs[0])[1] = sl.own_hi;
    } else if (cmd == "current_stage") {
        plhs[0] = mxCreateDoubleScalar(bellman_current_stage(h));
    } else if (cmd == "get_J") {
        const int32_t stage = nrhs > 2 ? (int32_t)mxGetScalar(prhs[2]) : bellman_current_stage(h);
        plhs[0] = mxCreateDoubleMatrix(sh.S_own, sh.P, mxREAL);
        check(bellman_get_J(h, stage, mxGetPr(plhs[0])), h);
    } else if (cmd == "get_idx") {
        const int32_t stage = nrhs > 2 ? (int32_t)mxGetScalar(prhs[2]) : bellman_current_stage(h);
        plhs[0] = mxCreateNumericMatrix(sh.S_own, sh.P, mxINT32_CLASS, mxREAL);
        int32_t *p = static_cast<int32_t *>(mxGetData(plhs[0]));
        check(bellman_get_idx(h, stage, p), h);
        for (size_t k = 0; k < sh.S_own * sh.P; ++k) p[k] += 1;   // MATLAB's min() index is 1-based
    } else if (cmd == "check_log") {
        const int n = bellman_get_check_log(h, nullptr, 0);
        plhs[0] = mxCreateDoubleMatrix(3, n > 0 ? n : 0, mxREAL);
        if (n > 0) bellman_get_check_log(h, mxGetPr(plhs[0]), n);
    } else if (cmd == "stats") {
        double ms = 0, mx = 0;
        int64_t launches = 0;
        check(bellman_last_run_stats(h, &ms, &launches, &mx), h);
        const char *names[] = {"ms", "launches", "ms_exchange", "kernel"};
        plhs[0] = mxCreateStructMatrix(1, 1, 4, names);
        mxSetField(plhs[0], 0, "ms", mxCreateDoubleScalar(ms));
        mxSetField(plhs[0], 0, "launches", mxCreateDoubleScalar((double)launches));
        mxSetField(plhs[0], 0, "ms_exchange", mxCreateDoubleScalar(mx));
        mxSetField(plhs[0], 0, "kernel", mxCreateString(bellman_last_kernel(h)));
    } else if (cmd == "rollout") {
        // [X,U] = bellman_mex('rollout', h, N, A, B, u_values, X0, mode, ssu_stage)
        if (nrhs < 9) mexErrMsgIdAndTxt("bellman:BAD_ARG", "usage: [X,U] = bellman_mex('rollout', h, N, A, B, u_values, X0, mode, ssu_stage)");
        const int N = (int)mxGetScalar(prhs[2]);
        if (N != sh.N) mexErrMsgIdAndTxt("bellman:BAD_ARG", "rollout: N = %d differs from the handle's horizon %d", N, sh.N);
        const size_t batch = mxGetNumberOfElements(prhs[6]) / 2;
        if (batch < 1 || !mxIsDouble(prhs[6]) || mxGetNumberOfElements(prhs[6]) != 2 * batch)
            mexErrMsgIdAndTxt("bellman:BAD_ARG", "rollout: X0 must be a 2-by-batch double array");
        need_doubles(prhs[3], 4, "A");
        need_doubles(prhs[4], 2, "B");
        need_doubles(prhs[5], (size_t)sh.C, "u_values");
        plhs[0] = mxCreateDoubleMatrix(2, (size_t)N * batch, mxREAL);     // [2][N][batch]
        mxArray *U = mxCreateDoubleMatrix((size_t)N, batch, mxREAL);
        check(bellman_rollout(h, mxGetPr(prhs[3]), mxGetPr(prhs[4]), mxGetPr(prhs[5]), mxGetPr(prhs[6]),
                              (int32_t)batch, (int32_t)mxGetScalar(prhs[7]), (int32_t)mxGetScalar(prhs[8]),
                              mxGetPr(plhs[0]), mxGetPr(U)), h);
        if (nlhs > 1) plhs[1] = U;
    } else if (cmd == "set_stage") {
        // bellman_mex('set_stage', h, stage, J, U_Optimal_id)   (J [] = zeros; U_Optimal_id 1-based, [] = none)
        if (nrhs < 3) mexErrMsgIdAndTxt("bellman:BAD_ARG", "usage: bellman_mex('set_stage', h, stage, J, idx)");
        const double *J = (nrhs > 3 && !mxIsEmpty(prhs[3])) ? need_doubles(prhs[3], sh.S_global * sh.P, "J") : nullptr;
        std::vector<int32_t> idx;
        if (nrhs > 4 && !mxIsEmpty(prhs[4])) {
            const size_t n = mxGetNumberOfElements(prhs[4]);
            if (n != sh.S_own * sh.P || !mxIsDouble(prhs[4])) mexErrMsgIdAndTxt("bellman:BAD_ARG", "idx must be a double array with one entry per state");
            idx.resize(n);
            for (size_t k = 0; k < n; ++k) idx[k] = (int32_t)mxGetPr(prhs[4])[k] - 1;
        }
        check(bellman_set_stage(h, (int32_t)mxGetScalar(prhs[2]), J, idx.empty() ? nullptr : idx.data()), h);
    } else if (cmd == "policy_lookup") {
        // id = bellman_mex('policy_lookup', h, prob, stage, X)   X is D-by-batch; id is 1-based like U_Optimal_id
        if (nrhs < 5) mexErrMsgIdAndTxt("bellman:BAD_ARG", "usage: id = bellman_mex('policy_lookup', h, prob, stage, X)");
        const size_t batch = mxGetN(prhs[4]);
        need_doubles(prhs[4], (size_t)sh.D * batch, "X (D-by-batch)");
        plhs[0] = mxCreateNumericMatrix(1, batch, mxINT32_CLASS, mxREAL);
        int32_t *p = static_cast<int32_t *>(mxGetData(plhs[0]));
        check(bellman_policy_lookup(h, (int32_t)mxGetScalar(prhs[2]) - 1, (int32_t)mxGetScalar(prhs[3]), mxGetPr(prhs[4]),
                                    (int32_t)batch, p), h);
        for (size_t k = 0; k < batch; ++k) p[k] += 1;
    } else if (cmd == "rollout_axis") {
        // [X,id] = bellman_mex('rollout_axis', h, prob, time_varying, stage, rate_dim, h_step, u_inc, X0, n_steps)
        if (nrhs < 10) mexErrMsgIdAndTxt("bellman:BAD_ARG", "usage: [X,id] = bellman_mex('rollout_axis', h, prob, time_varying, stage, rate_dim, h_step, u_inc, X0, n_steps)");
        const size_t batch = mxGetN(prhs[8]);
        const int n_steps = (int)mxGetScalar(prhs[9]);
        if (n_steps < 1) mexErrMsgIdAndTxt("bellman:BAD_ARG", "rollout_axis: n_steps must be positive");
        need_doubles(prhs[7], (size_t)sh.C, "u_inc");
        need_doubles(prhs[8], 2 * batch, "X0 (2-by-batch)");
        plhs[0] = mxCreateDoubleMatrix(2, (size_t)(n_steps + 1) * batch, mxREAL);   // [2][n_steps+1][batch]
        mxArray *id = mxCreateNumericMatrix((size_t)n_steps, batch, mxINT32_CLASS, mxREAL);
        int32_t *p = static_cast<int32_t *>(mxGetData(id));
        check(bellman_rollout_axis(h, (int32_t)mxGetScalar(prhs[2]) - 1, (int32_t)mxGetScalar(prhs[3]),
                                   (int32_t)mxGetScalar(prhs[4]), (int32_t)mxGetScalar(prhs[5]) - 1, mxGetScalar(prhs[6]),
                                   mxGetPr(prhs[7]), mxGetPr(prhs[8]), (int32_t)batch, n_steps, mxGetPr(plhs[0]), p), h);
        for (size_t k = 0; k < (size_t)n_steps * batch; ++k) p[k] += 1;
        if (nlhs > 1) plhs[1] = id; else mxDestroyArray(id);
    } else if (cmd == "rollout_orbit") {
        if (nrhs < 6 || !mxIsStruct(prhs[3])) mexErrMsgIdAndTxt("bellman:BAD_ARG", "usage: [X,id,w] = bellman_mex('rollout_orbit', h, stage, o, u_values, Y0)");
        bellman_orbit_opts o;
        std::memset(&o, 0, sizeof(o));
        o.struct_size = (int32_t)sizeof(o);
        o.n_steps = (int32_t)field_scalar(prhs[3], "n_steps", 0);
        o.stride_out = (int32_t)field_scalar(prhs[3], "stride_out", 1);
        o.mu = field_scalar(prhs[3], "mu", 398600.0);
        o.h = field_scalar(prhs[3], "h", 0.0);
        o.tol = field_scalar(prhs[3], "tol", 1e-8);
        const double *R0 = need_doubles(mxGetField(prhs[3], 0, "R0"), 3, "o.R0"), *V0 = need_doubles(mxGetField(prhs[3], 0, "V0"), 3, "o.V0");
        for (int k = 0; k < 3; ++k) { o.R0[k] = R0[k]; o.V0[k] = V0[k]; }
        if (o.n_steps < 1 || o.stride_out < 1 || o.n_steps % o.stride_out) mexErrMsgIdAndTxt("bellman:BAD_ARG", "n_steps must be a positive multiple of stride_out");
        const size_t batch = mxGetN(prhs[5]);
        need_doubles(prhs[4], (size_t)sh.C, "u_values");
        need_doubles(prhs[5], 6 * batch, "Y0 (6-by-batch)");
        const size_t n_out = (size_t)(o.n_steps / o.stride_out);
        plhs[0] = mxCreateDoubleMatrix(6, (n_out + 1) * batch, mxREAL);
        mxArray *id = mxCreateNumericMatrix(3, n_out * batch, mxINT32_CLASS, mxREAL);
        mxArray *w = mxCreateNumericMatrix(1, batch, mxINT32_CLASS, mxREAL);
        int32_t *pi = static_cast<int32_t *>(mxGetData(id));
        check(bellman_rollout_orbit(h, (int32_t)mxGetScalar(prhs[2]), &o, mxGetPr(prhs[4]), mxGetPr(prhs[5]), (int32_t)batch,
                                    mxGetPr(plhs[0]), pi, static_cast<int32_t *>(mxGetData(w))), h);
        for (size_t k = 0; k < 3 * n_out * batch; ++k) pi[k] += 1;
        if (nlhs > 1) plhs[1] = id; else mxDestroyArray(id);
        if (nlhs > 2) plhs[2] = w; else mxDestroyArray(w);
    } else if (cmd == "rollout_attitude") {
        if (nrhs < 6 || !mxIsStruct(prhs[3])) mexErrMsgIdAndTxt("bellman:BAD_ARG", "usage: [X,id,w] = bellman_mex('rollout_attitude', h, stage, o, u_values, Y0)");
        bellman_plant_opts o;
        plant_opts_from(prhs[3], o);
        const size_t batch = mxGetN(prhs[5]);
        need_doubles(prhs[4], (size_t)sh.C, "u_values");
        need_doubles(prhs[5], 7 * batch, "Y0 (7-by-batch)");
        const size_t n_out = (size_t)(o.n_steps / o.stride_out);
        plhs[0] = mxCreateDoubleMatrix(7, (n_out + 1) * batch, mxREAL);
        mxArray *id = mxCreateNumericMatrix(3, n_out * batch, mxINT32_CLASS, mxREAL);
        mxArray *w = mxCreateNumericMatrix(1, batch, mxINT32_CLASS, mxREAL);
        int32_t *pi = static_cast<int32_t *>(mxGetData(id));
        check(bellman_rollout_attitude(h, (int32_t)mxGetScalar(prhs[2]), &o, mxGetPr(prhs[4]), mxGetPr(prhs[5]), (int32_t)batch,
                                       mxGetPr(plhs[0]), pi, static_cast<int32_t *>(mxGetData(w))), h);
        for (size_t k = 0; k < 3 * n_out * batch; ++k) pi[k] += 1;
        if (nlhs > 1) plhs[1] = id; else mxDestroyArray(id);
        if (nlhs > 2) plhs[2] = w; else mxDestroyArray(w);
    } else {
        mexErrMsgIdAndTxt("bellman:BAD_ARG", "unknown command '%s'", cmd.c_str());
    }
}
