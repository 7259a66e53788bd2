/*
 * bellman_oracle.c — CPU restatement of the reference's backward Bellman stage.  TEST INFRASTRUCTURE.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * build, load or call this file.  The product (libbellman.so and the package above it) never does.
 *
 * What it restates (reference file:line, all MATLAB):
 *   stage operator  [F.Values,idx] = min(J_current + F(x_next..), [], ctrl_dim)
 *       test/Dynamic_Solver.m:207-210      position-control/Solver_position.m:135-137
 *       attitude-control/Solver_attitude.m:239-241      pos-att/Solver_pos_att.m:272
 *   griddedInterpolant(...,'linear') evaluation incl. linear extrapolation from the edge cell
 *       (MATLAB runtime, closed source, un-vendored; version pinned only by test/obj_1.mat's header:
 *        PCWIN64, 24 Mar 2017 => R2016b/R2017a).  Published algorithm: N-linear interpolation on
 *        the cell s[i] <= x < s[i+1], ExtrapolationMethod defaults to Method.
 *   min(X,[],dim): first (lowest) index wins ties.
 *   stage loop      Dynamic_Solver.m:86-102, Solver_position.m:132-141, Solver_attitude.m:236-247,
 *                   Solver_pos_att.m:270-286 (incl. the every-50-stages sum check :273-285)
 *   rollout         Dynamic_Solver.m:108-145,191-194
 *
 * The S x C arrays of the reference are sums of 1-D tables; the tables are evaluated by
 * oracle/matlab_literal.py (array-at-a-time, literally as the .m files do) or by the product's
 * facade, and passed here through the same descriptor the library takes (include/bellman.h),
 * whose header comment is the normative operation-by-operation arithmetic.
 *
 * Parity pin: tests/test_oracle_golden.py checks this file against the reference's only golden
 * vector, test/obj_1.mat (35x35 grid, 100 controls, 129 stages, fp64): u_star exact on all
 * 158 025 entries, J within 1e-12 relative.  Position / attitude / pos-att have no stored
 * outputs in the reference => parity unpinned against MATLAB for those tables (same operator).
 *
 * Build: make -C oracle   (gcc -O2 -fopenmp -ffp-contract=off; fma() calls are explicit)
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif
#include "../include/bellman.h"

#define MAXD BELLMAN_MAX_DIM

/* --- locate-mode rule (same rule as the library; restated, not shared) -------------------- */
static int grid_is_uniform(const double *s, int n)
{
    const double h = (s[n - 1] - s[0]) / (double)(n - 1);
    for (int i = 0; i < n; ++i) {
        const double dev = fabs(s[i] - (s[0] + (double)i * h));
        if (!(dev <= 1e-14 * (s[n - 1] - s[0]))) return 0;
    }
    return 1;
}

int oracle_locate_modes(const bellman_desc *d, int32_t *modes /*[P][D]*/)
{
    for (int p = 0; p < d->P; ++p)
        for (int k = 0; k < d->D; ++k)
            modes[p * d->D + k] = grid_is_uniform(d->grid[k] + (size_t)p * d->n[k], d->n[k])
                                      ? BELLMAN_LOCATE_UNIFORM : BELLMAN_LOCATE_SEARCH;
    return 0;
}

typedef struct {
    const double *s;
    double *rinv;
    int n, mode;
    double inv_h, off;
} dimtab;

/* x is in kernel units: fractional cell coordinate (UNIFORM, tables pre-scaled) or state value (SEARCH) */
static inline int locate(const dimtab *g, double x, double *t)
{
    int cell;
    if (g->mode == BELLMAN_LOCATE_UNIFORM) {
        if (!(x >= 0.0)) cell = 0;                    /* floor() then clamp, without int overflow; NaN -> 0
                                                         like the GPU's saturating cvt.rmi.s32.f64 */
        else if (x >= (double)(g->n - 1)) cell = g->n - 2;
        else cell = (int)x;
        *t = x - (double)cell;
    } else {
        int lo = 0, hi = g->n;                        /* count of s[i] <= x */
        while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            if (g->s[mid] <= x) lo = mid + 1; else hi = mid;
        }
        cell = lo - 1;
        if (cell < 0) cell = 0;
        if (cell > g->n - 2) cell = g->n - 2;
        *t = (x - g->s[cell]) * g->rinv[cell];
    }
    return cell;
}

/* a free state value (rollout): brought to kernel units first */
static inline int locate_state(const dimtab *g, double x, double *t)
{
    return locate(g, g->mode == BELLMAN_LOCATE_UNIFORM ? fma(x, g->inv_h, g->off) : x, t);
}

static void dimtab_init(dimtab *g, const double *s, int n, int mode)
{
    g->s = s; g->n = n; g->mode = mode;
    g->rinv = (double *)malloc(sizeof(double) * (size_t)(n - 1));
    for (int i = 0; i < n - 1; ++i) g->rinv[i] = 1.0 / (s[i + 1] - s[i]);
    g->inv_h = (double)(n - 1) / (s[n - 1] - s[0]);
    g->off = -(s[0] * g->inv_h);
}

/* J_{N-1} of one node when the terminal cost J_N is identically zero (Dynamic_Solver.m:83-84 and the
 * zeros(...) interpolants of the other classes): every corner is 0, every lerp fma(t, 0 - 0, 0) = +0,
 * so the stage reduces to min_c ((gs + r[c]) + 0.0) — evaluated here exactly as eval_state would.
 * Lets a spot check of stage N-2 run at grid sizes whose J arrays do not fit the host. */
static double first_stage_value(const bellman_desc *d, const double *const *q, const double *r, int64_t s)
{
    const int D = d->D;
    int i[MAXD] = {0, 0, 0, 0};
    for (int k = 0; k < D; ++k) { i[k] = (int)(s % d->n[k]); s /= d->n[k]; }
    double gs = q[d->q_order[0]][i[d->q_order[0]]];
    for (int m = 1; m < D; ++m) gs = gs + q[d->q_order[m]][i[d->q_order[m]]];
    /* min_c ((gs + r[c]) + 0.0): rounding is monotone, so the minimum is attained at the smallest r[c] */
    double rmin = r[0];
    for (int c = 1; c < d->C; ++c) if (r[c] < rmin) rmin = r[c];
    return (gs + rmin) + 0.0;
}

/* one state of one problem: the normative operation order of include/bellman.h
 * (Jn == NULL: J_{k+1} is the first stage from a zero terminal cost, evaluated on the fly) */
static void eval_state(const bellman_desc *d, const dimtab *g, const double *const *Ta,
                       const double *const *Tb, const double *const *Tc, const double *const *q,
                       const double *r, const int64_t *stride, const double *Jn, int64_t s,
                       double *J_out, int32_t *idx_out)
{
    const int D = d->D, C = d->C;
    int i[MAXD] = {0, 0, 0, 0};
    int64_t rem = s;
    for (int k = 0; k < D; ++k) { i[k] = (int)(rem % d->n[k]); rem /= d->n[k]; }
    double base[MAXD];
    for (int k = 0; k < D; ++k) {
        base[k] = Ta[k][i[d->src_a[k]]];
        if (Tb[k]) base[k] = base[k] + Tb[k][i[d->src_b[k]]];
    }
    double gs = q[d->q_order[0]][i[d->q_order[0]]];
    for (int m = 1; m < D; ++m) gs = gs + q[d->q_order[m]][i[d->q_order[m]]];

    double best = INFINITY;
    int arg = 0;
    for (int c = 0; c < C; ++c) {
        int cell[MAXD];
        double t[MAXD];
        int64_t o = 0;
        for (int k = 0; k < D; ++k) {
            const double xq = Tc[k] ? base[k] + Tc[k][c] : base[k];
            cell[k] = locate(&g[k], xq, &t[k]);
            o += cell[k] * stride[k];
        }
        double v[1 << MAXD];
        for (int m = 0; m < (1 << D); ++m) {
            int64_t oo = o;
            for (int k = 0; k < D; ++k) if (m & (1 << k)) oo += stride[k];
            v[m] = Jn ? Jn[oo] : first_stage_value(d, q, r, oo);
        }
        for (int k = 0; k < D; ++k)                       /* dimension 0 reduced first */
            for (int m = 0; m < (1 << (D - 1 - k)); ++m)
                v[m] = fma(t[k], v[2 * m + 1] - v[2 * m], v[2 * m]);
        const double tot = (gs + r[c]) + v[0];
        if (tot < best) { best = tot; arg = c; }          /* strict: first index wins ties */
    }
    *J_out = best;
    *idx_out = arg;
}

/* per-problem tables in kernel units: UNIFORM dimensions are pre-scaled to cell units
 * (include/bellman.h): Ta' = fma(Ta, inv_h, off), Tb' = Tb*inv_h, Tc' = Tc*inv_h, one rounding each.
 * The scaled copies are owned by `own` and released by problem_tables_free(). */
typedef struct { double *buf[3 * MAXD]; int n; } owned_tabs;

static void problem_tables(const bellman_desc *d, const int32_t *modes, int p, dimtab *g,
                           const double **Ta, const double **Tb, const double **Tc, const double **q,
                           owned_tabs *own)
{
    own->n = 0;
    for (int k = 0; k < d->D; ++k) {
        dimtab_init(&g[k], d->grid[k] + (size_t)p * d->n[k], d->n[k], modes[p * d->D + k]);
        Ta[k] = d->Ta[k] + (size_t)p * d->n[d->src_a[k]];
        Tb[k] = (d->Tb[k] && d->src_b[k] >= 0) ? d->Tb[k] + (size_t)p * d->n[d->src_b[k]] : NULL;
        Tc[k] = d->Tc[k] ? d->Tc[k] + (size_t)p * d->C : NULL;
        q[k] = d->q[k] + (size_t)p * d->n[k];
        if (g[k].mode == BELLMAN_LOCATE_UNIFORM) {
            const double ih = g[k].inv_h, of = g[k].off;
            const int na = d->n[d->src_a[k]];
            double *a = (double *)malloc(sizeof(double) * (size_t)na);
            for (int i = 0; i < na; ++i) a[i] = fma(Ta[k][i], ih, of);
            Ta[k] = own->buf[own->n++] = a;
            if (Tb[k]) {
                const int nb = d->n[d->src_b[k]];
                double *b = (double *)malloc(sizeof(double) * (size_t)nb);
                for (int i = 0; i < nb; ++i) b[i] = Tb[k][i] * ih;
                Tb[k] = own->buf[own->n++] = b;
            }
            if (Tc[k]) {
                double *c = (double *)malloc(sizeof(double) * (size_t)d->C);
                for (int i = 0; i < d->C; ++i) c[i] = Tc[k][i] * ih;
                Tc[k] = own->buf[own->n++] = c;
            }
        }
    }
}

static void problem_tables_free(const bellman_desc *d, dimtab *g, owned_tabs *own)
{
    for (int k = 0; k < d->D; ++k) free(g[k].rinv);
    for (int k = 0; k < own->n; ++k) free(own->buf[k]);
}

/*
 * Evaluate only the listed states (size-independent spot check at BASELINE's full grid sizes):
 * states[m] is a linear state index of problem `p`; J_next is that problem's [S] array.
 */
int oracle_stage_points(const bellman_desc *d, const int32_t *modes, int p, const double *J_next,
                        const int64_t *states, int64_t n_states, double *J_out, int32_t *idx_out)
{
    const int D = d->D;
    if (D < 1 || D > MAXD) return -1;
    int64_t stride[MAXD], S = 1;
    for (int k = 0; k < D; ++k) { stride[k] = S; S *= d->n[k]; }
    dimtab g[MAXD];
    const double *Ta[MAXD], *Tb[MAXD], *Tc[MAXD], *q[MAXD];
    owned_tabs own;
    problem_tables(d, modes, p, g, Ta, Tb, Tc, q, &own);
    const double *r = d->r + (size_t)p * d->C;
#pragma omp parallel for schedule(static)
    for (int64_t m = 0; m < n_states; ++m)
        eval_state(d, g, Ta, Tb, Tc, q, r, stride, J_next, states[m], &J_out[m], &idx_out[m]);
    problem_tables_free(d, g, &own);
    return 0;
}

/*
 * One backward stage over all P problems.  J_next / J_out are [P][S] column-major (dim 0 fastest),
 * idx_out is [P][S] 0-based.  own_lo/own_hi restrict the states computed along part_dim (used by
 * the multi-rank tests); pass part_dim = -1 for the whole grid.  Outputs are always indexed
 * globally.
 */
int oracle_stage(const bellman_desc *d, const int32_t *modes, const double *J_next, double *J_out,
                 int32_t *idx_out, int part_dim, int own_lo, int own_hi, int nthreads)
{
    const int D = d->D, C = d->C;
    if (D < 1 || D > MAXD) return -1;
    int64_t stride[MAXD], S = 1;
    for (int k = 0; k < D; ++k) { stride[k] = S; S *= d->n[k]; }
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#else
    (void)nthreads;
#endif
    for (int p = 0; p < d->P; ++p) {
        dimtab g[MAXD];
        const double *Ta[MAXD], *Tb[MAXD], *Tc[MAXD], *q[MAXD];
        owned_tabs own;
        problem_tables(d, modes, p, g, Ta, Tb, Tc, q, &own);
        const double *r = d->r + (size_t)p * C;
        const double *Jn = J_next + (size_t)p * S;
        double *Jo = J_out + (size_t)p * S;
        int32_t *Io = idx_out + (size_t)p * S;

#pragma omp parallel for schedule(static)
        for (int64_t s = 0; s < S; ++s) {
            if (part_dim >= 0) {
                const int ip = (int)((s / stride[part_dim]) % d->n[part_dim]);
                if (ip < own_lo || ip >= own_hi) continue;
            }
            eval_state(d, g, Ta, Tb, Tc, q, r, stride, Jn, s, &Jo[s], &Io[s]);
        }
        problem_tables_free(d, g, &own);
    }
    return 0;
}

/*
 * Full sweep from the terminal cost J_N (NULL = zeros) down n_stages stages.
 * J_all / idx_all (may be NULL) receive every stage: layout [N][P][S], slot (stage-1); slot N-1 of
 * J_all holds the terminal cost, slot N-1 of idx_all is left untouched.  J_last/idx_last (may be
 * NULL) receive the final stage only.  check_period/check_tol restate Solver_pos_att.m:273-285
 * (with idsum50_prev defined as 0 — the reference uses it undefined); the return
 * value is the stage number of the last computed J.
 */
int oracle_sweep(const bellman_desc *d, const int32_t *modes, const double *J_N, int n_stages,
                 double *J_all, int32_t *idx_all, double *J_last, int32_t *idx_last,
                 int check_period, double check_tol, int nthreads)
{
    int64_t S = 1;
    for (int k = 0; k < d->D; ++k) S *= d->n[k];
    const size_t PS = (size_t)d->P * (size_t)S;
    double *a = (double *)malloc(sizeof(double) * PS), *b = (double *)malloc(sizeof(double) * PS);
    int32_t *ib = (int32_t *)malloc(sizeof(int32_t) * PS);
    if (J_N) memcpy(a, J_N, sizeof(double) * PS); else memset(a, 0, sizeof(double) * PS);
    int stage = d->N;
    if (J_all) memcpy(J_all + (size_t)(stage - 1) * PS, a, sizeof(double) * PS);
    double fsum_prev = 0.0;
    for (int it = 0; it < n_stages && stage > 1; ++it) {
        oracle_stage(d, modes, a, b, ib, -1, 0, 0, nthreads);
        --stage;
        double *tmp = a; a = b; b = tmp;
        if (J_all) memcpy(J_all + (size_t)(stage - 1) * PS, a, sizeof(double) * PS);
        if (idx_all) memcpy(idx_all + (size_t)(stage - 1) * PS, ib, sizeof(int32_t) * PS);
        if (check_period > 0 && (stage % check_period) == 0) {
            double fsum = 0.0;
            for (size_t k = 0; k < PS; ++k) fsum += a[k];
            const double e = fsum - fsum_prev;
            fsum_prev = fsum;
            if (fabs(e) < check_tol) break;
        }
    }
    if (J_last) memcpy(J_last, a, sizeof(double) * PS);
    if (idx_last) memcpy(idx_last, ib, sizeof(int32_t) * PS);
    free(a); free(b); free(ib);
    return stage;
}

/*
 * Rollout of Dynamic_Solver.get_optimal_path (test/Dynamic_Solver.m:108-145,191-194), fp64:
 *   U(k) = Fu(X(1,k),X(2,k)) with Fu = griddedInterpolant(X1_mesh,X2_mesh,u_star(:,:,k),'linear')
 *   X(:,k+1) = A*[X1;X2] + B*U(k)
 * idx_all is [N][S] (slot stage-1), u_values[C].  The 2x2 product is evaluated as
 * (A(r,1)*x1 + A(r,2)*x2) + B(r)*u with separate roundings (MATLAB's own mtimes rounding for a
 * 2x2 is not documented; the rollout is compared with a tolerance).
 */
int oracle_rollout(const bellman_desc *d, const int32_t *modes, const int32_t *idx_all,
                   const double *A, const double *B, const double *u_values, const double *x0,
                   int batch, int mode, int ssu_stage, double *X_out, double *U_out)
{
    if (d->D != 2 || d->P != 1) return -1;
    const int N = d->N, n0 = d->n[0], n1 = d->n[1];
    const int64_t S = (int64_t)n0 * n1;
    dimtab g[2];
    dimtab_init(&g[0], d->grid[0], n0, modes[0]);
    dimtab_init(&g[1], d->grid[1], n1, modes[1]);
    for (int b = 0; b < batch; ++b) {
        double x1 = x0[2 * b], x2 = x0[2 * b + 1];
        double *X = X_out + (size_t)b * 2 * N, *U = U_out + (size_t)b * N;
        X[0] = x1; X[1] = x2;
        for (int k = 1; k <= N - 1; ++k) {
            const int st = (mode == 1) ? ssu_stage : k;
            const int32_t *id = idx_all + (size_t)(st - 1) * S;
            double t0, t1;
            const int c0 = locate_state(&g[0], x1, &t0), c1 = locate_state(&g[1], x2, &t1);
            const double v00 = u_values[id[c0 + (int64_t)c1 * n0]];
            const double v10 = u_values[id[c0 + 1 + (int64_t)c1 * n0]];
            const double v01 = u_values[id[c0 + (int64_t)(c1 + 1) * n0]];
            const double v11 = u_values[id[c0 + 1 + (int64_t)(c1 + 1) * n0]];
            const double a = fma(t0, v10 - v00, v00), bb = fma(t0, v11 - v01, v01);
            const double u = fma(t1, bb - a, a);
            U[k - 1] = u;
            const double nx1 = (A[0] * x1 + A[2] * x2) + B[0] * u;
            const double nx2 = (A[1] * x1 + A[3] * x2) + B[1] * u;
            x1 = nx1; x2 = nx2;
            X[2 * k] = x1; X[2 * k + 1] = x2;
        }
        U[N - 1] = 0.0;
    }
    free(g[0].rinv); free(g[1].rinv);
    return 0;
}

/* torchrun exports OMP_NUM_THREADS=1; the CPU arm wants every host core */
void oracle_set_threads(int n)
{
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

int oracle_num_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* ---- consumers of the sweep's output ------------------------------------------------------- */

/* nearest grid node of dimension d (clamped; an exact midpoint goes to the upper node) */
static int nearest_node(const dimtab *g, double x)
{
    double t;
    const int cell = locate_state(g, x, &t);
    return cell + ((x - g->s[cell]) >= (g->s[cell + 1] - x) ? 1 : 0);
}

/* griddedInterpolant({s1,..}, U_vector(idx), 'nearest') at `batch` states of problem p
 * (Solver_position.m:144-146, Solver_attitude.m:249-251, Solver_pos_att.m:851-861):
 * x is [D][batch] (batch slowest), idx the [S] policy of one stage, idx_out[batch]. */
int oracle_policy_lookup(const bellman_desc *d, const int32_t *modes, int p, const int32_t *idx,
                         const double *x, int batch, int32_t *idx_out)
{
    const int D = d->D;
    dimtab g[MAXD];
    for (int k = 0; k < D; ++k) dimtab_init(&g[k], d->grid[k] + (size_t)p * d->n[k], d->n[k], modes[p * D + k]);
    for (int b = 0; b < batch; ++b) {
        int64_t o = 0, st = 1;
        for (int k = 0; k < D; ++k) { o += nearest_node(&g[k], x[(size_t)b * D + k]) * st; st *= d->n[k]; }
        idx_out[b] = idx[o];
    }
    for (int k = 0; k < D; ++k) free(g[k].rinv);
    return 0;
}

/* Simplified-plant rollout under the nearest policy, attitude-control/test/test_simplified.m:129-151
 * with next_stage_states / RK4_w / RK4_t of :273-310 (the plant Solver_position discretises too):
 *   c = nearest policy;  x_r' = x_r + u_inc[c];
 *   x_o' = x_o + h*(k1 + 2*k2 + 2*k3 + k4)/6,  k1 = x_r, k2 = x_r + k1*h/2, k3 = x_r + k2*h/2, k4 = x_r + k3*h
 * idx: [S] (fixed policy) or [n_steps][stride] (time varying, slot k-1 = stage k).
 * x0 [2][batch]; X_out [2][n_steps+1][batch]; C_out [n_steps][batch]. */
int oracle_rollout_axis(const bellman_desc *d, const int32_t *modes, int p, const int32_t *idx,
                        int64_t idx_stage_stride, int time_varying, int rate_dim, double h,
                        const double *u_inc, const double *x0, int batch, int n_steps,
                        double *X_out, int32_t *C_out)
{
    if (d->D != 2) return -1;
    dimtab g[2];
    for (int k = 0; k < 2; ++k) dimtab_init(&g[k], d->grid[k] + (size_t)p * d->n[k], d->n[k], modes[p * 2 + k]);
    const int r = rate_dim, o = 1 - rate_dim;
    for (int b = 0; b < batch; ++b) {
        double x[2] = {x0[2 * (size_t)b], x0[2 * (size_t)b + 1]};
        double *X = X_out + (size_t)b * 2 * (n_steps + 1);
        int32_t *Cc = C_out + (size_t)b * n_steps;
        X[0] = x[0]; X[1] = x[1];
        for (int k = 1; k <= n_steps; ++k) {
            const int32_t *id = idx + (time_varying ? (size_t)(k - 1) * idx_stage_stride : 0);
            const int c = id[nearest_node(&g[0], x[0]) + (int64_t)nearest_node(&g[1], x[1]) * d->n[0]];
            Cc[k - 1] = c;
            const double k1 = x[r];
            const double k2 = x[r] + (k1 * h) / 2;
            const double k3 = x[r] + (k2 * h) / 2;
            const double k4 = x[r] + k3 * h;
            const double xo = x[o] + (h * (((k1 + 2 * k2) + 2 * k3) + k4)) / 6;
            const double xr = x[r] + u_inc[c];
            x[r] = xr; x[o] = xo;
            X[2 * k] = x[0]; X[2 * k + 1] = x[1];
        }
    }
    free(g[0].rinv); free(g[1].rinv);
    return 0;
}

/* =============================================================================================
 * Orbital forward simulation of Solver_position.get_optimal_path.
 *
 * Restates, operation by operation (left-to-right association as MATLAB evaluates it):
 *   position-control/private/stumpC.m:11-17, stumpS.m:11-17      Stumpff functions C(z), S(z)
 *   position-control/private/kepler_U.m:25-44                    Newton iteration on the universal Kepler equation
 *   position-control/private/f_and_g.m:21-27, fDot_and_gDot.m:23-29   Lagrange coefficients
 *   position-control/private/sv_from_coe.m:31-66                 state vector from orbital elements
 *   position-control/private/rkf45.m:49-118                      Runge-Kutta-Fehlberg 4(5), adaptive step
 *   position-control/Solver_position.m:189-224                   stage loop: nearest policy, one rkf45 call per stage
 *   position-control/Solver_position.m:259-309 (rates), :313-331 (get_target_R0V0), :333-361 (update_RV_target)
 * The private/*.m files use bare-CR line endings (read them with tr '\r' '\n').
 *
 * PARITY UNPINNED: the reference stores no output of this path and MATLAB cannot run here.  Details
 * MATLAB does not document and this restatement fixes: dot products / matrix-vector products are
 * summed left to right (MATLAB calls BLAS), norm() is sqrt of the left-to-right sum of squares,
 * tspan(k) = (k-1)*h, x^n is pow(x, n) except x^2 = x*x.  cos/sin/cosh/sinh/pow come from the C
 * library here and from CUDA's math library on the GPU, so GPU-vs-oracle parity for this path is a
 * tolerance (tests/test_gpu_orbit.py), not bit equality.
 * ============================================================================================= */
static double stumpC(double z)
{
    if (z > 0) return (1 - cos(sqrt(z))) / z;
    if (z < 0) return (cosh(sqrt(-z)) - 1) / (-z);
    return 0.5;
}

static double stumpS(double z)
{
    if (z > 0) { const double s = sqrt(z); return (s - sin(s)) / pow(s, 3); }
    if (z < 0) { const double s = sqrt(-z); return (sinh(s) - s) / pow(s, 3); }
    return 1.0 / 6;
}

/* kepler_U.m:25-44; *n_iter receives the iteration count */
double oracle_kepler_U(double mu, double dt, double ro, double vro, double a, int *n_iter)
{
    const double error = 1.e-8;
    const int nMax = 1000;
    const double smu = sqrt(mu);
    double x = smu * fabs(a) * dt;
    int n = 0;
    double ratio = 1;
    while (fabs(ratio) > error && n <= nMax) {
        n = n + 1;
        const double x2 = x * x;
        const double C = stumpC(a * x2);
        const double S = stumpS(a * x2);
        const double F = ro * vro / smu * x2 * C + (1 - a * ro) * pow(x, 3) * S + ro * x - smu * dt;
        const double dFdx = ro * vro / smu * x * (1 - a * x2 * S) + (1 - a * ro) * x2 * C + ro;
        ratio = F / dFdx;
        x = x - ratio;
    }
    if (n_iter) *n_iter = n;
    return x;
}

/* Solver_position.m:333-361 (update_RV_target) with f_and_g.m / fDot_and_gDot.m inlined */
void oracle_update_RV_target(double mu, const double *R0, const double *V0, double t, double *R2, double *V2)
{
    const double r0 = sqrt(R0[0] * R0[0] + R0[1] * R0[1] + R0[2] * R0[2]);
    const double v0 = sqrt(V0[0] * V0[0] + V0[1] * V0[1] + V0[2] * V0[2]);
    const double vr0 = (R0[0] * V0[0] + R0[1] * V0[1] + R0[2] * V0[2]) / r0;
    const double alpha = 2 / r0 - v0 * v0 / mu;
    const double x = oracle_kepler_U(mu, t, r0, vr0, alpha, NULL);
    const double z = alpha * (x * x);
    const double f = 1 - x * x / r0 * stumpC(z);
    const double g = t - 1 / sqrt(mu) * pow(x, 3) * stumpS(z);
    for (int k = 0; k < 3; ++k) R2[k] = f * R0[k] + g * V0[k];
    const double r2 = sqrt(R2[0] * R2[0] + R2[1] * R2[1] + R2[2] * R2[2]);
    const double fdot = sqrt(mu) / r2 / r0 * (z * stumpS(z) - 1) * x;
    const double gdot = 1 - x * x / r2 * stumpC(z);
    for (int k = 0; k < 3; ++k) V2[k] = fdot * R0[k] + gdot * V0[k];
}

/* sv_from_coe.m:31-66; coe = [h e RA incl w TA] */
void oracle_sv_from_coe(const double *coe, double mu, double *r, double *v)
{
    const double h = coe[0], e = coe[1], RA = coe[2], incl = coe[3], w = coe[4], TA = coe[5];
    const double krp = (h * h / mu) * (1 / (1 + e * cos(TA)));
    const double rp[3] = {krp * (cos(TA) * 1 + sin(TA) * 0), krp * (cos(TA) * 0 + sin(TA) * 1), krp * (cos(TA) * 0 + sin(TA) * 0)};
    const double kvp = mu / h;
    const double vp[3] = {kvp * (-sin(TA) * 1 + (e + cos(TA)) * 0), kvp * (-sin(TA) * 0 + (e + cos(TA)) * 1),
                          kvp * (-sin(TA) * 0 + (e + cos(TA)) * 0)};
    const double R3W[3][3] = {{cos(RA), sin(RA), 0}, {-sin(RA), cos(RA), 0}, {0, 0, 1}};
    const double R1i[3][3] = {{1, 0, 0}, {0, cos(incl), sin(incl)}, {0, -sin(incl), cos(incl)}};
    const double R3w[3][3] = {{cos(w), sin(w), 0}, {-sin(w), cos(w), 0}, {0, 0, 1}};
    double A[3][3], Q[3][3];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) A[i][j] = (R3w[i][0] * R1i[0][j] + R3w[i][1] * R1i[1][j]) + R3w[i][2] * R1i[2][j];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) Q[i][j] = (A[i][0] * R3W[0][j] + A[i][1] * R3W[1][j]) + A[i][2] * R3W[2][j];
    for (int i = 0; i < 3; ++i) {      /* Q_pX = Q' */
        r[i] = (Q[0][i] * rp[0] + Q[1][i] * rp[1]) + Q[2][i] * rp[2];
        v[i] = (Q[0][i] * vp[0] + Q[1][i] * vp[1]) + Q[2][i] * vp[2];
    }
}

/* Solver_position.m:313-331 */
void oracle_get_target_R0V0(double mu, double *R0, double *V0)
{
    const double RE = 6378;
    const double rp = RE + 300, e = 0.1;
    const double ra = rp * (1 + e) / (1 - e);
    const double h_ = sqrt(2 * mu * rp * ra / (ra + rp));
    const double coe[6] = {h_, e, 0, 0, 0, 0};
    oracle_sv_from_coe(coe, mu, R0, V0);
}

typedef struct { double mu; const double *R0, *V0; double a[3]; } orbit_ctx;

/* Solver_position.m:259-309 (rates) */
static void orbit_rates(const orbit_ctx *c, double t, const double *y, double *dydt)
{
    double R[3], V[3];
    oracle_update_RV_target(c->mu, c->R0, c->V0, t, R, V);
    const double norm_R = pow((R[0] * R[0] + R[1] * R[1]) + R[2] * R[2], .5);
    const double RdotV = (R[0] * V[0] + R[1] * V[1]) + R[2] * V[2];
    const double cr[3] = {R[1] * V[2] - R[2] * V[1], R[2] * V[0] - R[0] * V[2], R[0] * V[1] - R[1] * V[0]};
    const double H = pow((cr[0] * cr[0] + cr[1] * cr[1]) + cr[2] * cr[2], .5);
    const double dx = y[0], dy = y[1], dz = y[2], dvx = y[3], dvy = y[4], dvz = y[5];
    const double mu = c->mu;
    const double nR2 = norm_R * norm_R, nR3 = pow(norm_R, 3), nR4 = pow(norm_R, 4), H2 = H * H;
    const double dax = (2 * mu / nR3 + H2 / nR4) * dx - 2 * RdotV / nR4 * H * dy + 2 * H / nR2 * dvy + c->a[0];
    const double day = -(mu / nR3 - H2 / nR4) * dy + 2 * RdotV / nR4 * H * dx - 2 * H / nR2 * dvx + c->a[1];
    const double daz = -mu / nR3 * dz + c->a[2];
    dydt[0] = dvx; dydt[1] = dvy; dydt[2] = dvz; dydt[3] = dax; dydt[4] = day; dydt[5] = daz;
}

static double eps_of(double t)   /* MATLAB eps(t): spacing of doubles at |t| */
{
    t = fabs(t);
    if (t < 2.2250738585072014e-308) return 4.9406564584124654e-324;
    int e;
    frexp(t, &e);
    return ldexp(1.0, e - 53);
}

/* rkf45.m:49-118 for a 6-state system; returns the last row of yout in y, the number of accepted
 * steps, and 1 in *warned when the step size fell below hmin (rkf45.m:113-117) */
static int orbit_rkf45(const orbit_ctx *c, double t0, double tf, double *y, double tol, int max_steps, int *warned)
{
    static const double a[6] = {0, 1. / 4, 3. / 8, 12. / 13, 1, 1. / 2};
    static const double b[6][5] = {{0, 0, 0, 0, 0},
                                   {1. / 4, 0, 0, 0, 0},
                                   {3. / 32, 9. / 32, 0, 0, 0},
                                   {1932. / 2197, -7200. / 2197, 7296. / 2197, 0, 0},
                                   {439. / 216, -8, 3680. / 513, -845. / 4104, 0},
                                   {-8. / 27, 2, -3544. / 2565, 1859. / 4104, -11. / 40}};
    static const double c4[6] = {25. / 216, 0, 1408. / 2565, 2197. / 4104, -1. / 5, 0};
    static const double c5[6] = {16. / 135, 0, 6656. / 12825, 28561. / 56430, -9. / 50, 2. / 55};
    double t = t0, h = (tf - t0) / 100, f[6][6], yi[6], yin[6];
    int accepted = 0, iters = 0;
    *warned = 0;
    while (t < tf && iters++ < max_steps) {
        const double hmin = 16 * eps_of(t);
        const double ti = t;
        memcpy(yi, y, sizeof(yi));
        for (int i = 0; i < 6; ++i) {
            const double t_inner = ti + a[i] * h;
            memcpy(yin, yi, sizeof(yin));
            for (int j = 0; j < i; ++j)
                for (int k = 0; k < 6; ++k) yin[k] = yin[k] + h * b[i][j] * f[j][k];
            orbit_rates(c, t_inner, yin, f[i]);
        }
        double te_max = 0, ymax = 0;
        for (int k = 0; k < 6; ++k) {
            double te = 0;
            for (int i = 0; i < 6; ++i) te = te + (h * f[i][k]) * (c4[i] - c5[i]);
            te_max = fmax(te_max, fabs(te));
            ymax = fmax(ymax, fabs(y[k]));
        }
        const double te_allowed = tol * fmax(ymax, 1.0);
        const double delta = pow(te_allowed / (te_max + 2.220446049250313e-16), 1. / 5);
        if (te_max <= te_allowed) {
            h = fmin(h, tf - t);
            t = t + h;
            for (int k = 0; k < 6; ++k) {
                double s = 0;
                for (int i = 0; i < 6; ++i) s = s + (h * f[i][k]) * c5[i];
                y[k] = yi[k] + s;
            }
            ++accepted;
        }
        h = fmin(delta * h, 4 * h);
        if (h < hmin) { *warned = 1; break; }
    }
    return accepted;
}

/* Solver_position.m:206-224: the stage loop.  The three axis policies are the nearest-node
 * interpolants U{1,2,3}_Opt = griddedInterpolant({s_x, s_v}, U_vector(U_idx), 'nearest') (:144-146)
 * given as problems 0..2 of `d` with idx [3][S] (0-based control indices) and u_values [C].
 * y0 [6][batch]; X_out [6][n_steps+1][batch]; C_out [3][n_steps][batch]; warn_out [batch] counts the
 * rkf45 calls that stopped on the minimum step size. */
int oracle_rollout_orbit(const bellman_desc *d, const int32_t *modes, const int32_t *idx, const double *u_values,
                         double mu, const double *R0, const double *V0, double h, int n_steps, double tol,
                         const double *y0, int batch, double *X_out, int32_t *C_out, int32_t *warn_out)
{
    if (d->D != 2 || d->P < 3) return -1;
    dimtab g[3][2];
    for (int p = 0; p < 3; ++p)
        for (int k = 0; k < 2; ++k) dimtab_init(&g[p][k], d->grid[k] + (size_t)p * d->n[k], d->n[k], modes[p * 2 + k]);
    const int64_t S = (int64_t)d->n[0] * d->n[1];
#pragma omp parallel for schedule(dynamic)
    for (int b = 0; b < batch; ++b) {
        double y[6];
        for (int k = 0; k < 6; ++k) y[k] = y0[(size_t)b * 6 + k];
        double *X = X_out + (size_t)b * 6 * (n_steps + 1);
        int32_t *Cc = C_out + (size_t)b * 3 * n_steps;
        memcpy(X, y, sizeof(y));
        orbit_ctx c = {mu, R0, V0, {0, 0, 0}};
        int warns = 0;
        for (int ks = 1; ks <= n_steps; ++ks) {
            for (int p = 0; p < 3; ++p) {     /* a_x = U1_Opt(x1, v1) ... :215-217 */
                const int ci = idx[(size_t)p * S + nearest_node(&g[p][0], y[p]) + (int64_t)nearest_node(&g[p][1], y[3 + p]) * d->n[0]];
                Cc[(size_t)(ks - 1) * 3 + p] = ci;
                c.a[p] = u_values[ci];
            }
            int w;
            orbit_rkf45(&c, (double)(ks - 1) * h, (double)ks * h, y, tol, 100000, &w);
            warns += w;
            memcpy(X + (size_t)ks * 6, y, sizeof(y));
        }
        if (warn_out) warn_out[b] = warns;
    }
    for (int p = 0; p < 3; ++p)
        for (int k = 0; k < 2; ++k) free(g[p][k].rinv);
    return 0;
}

/* ==== 13-state / 7-state plants integrated with ode45 (second half) ===========
 * Solver_pos_att.get_optimal_path (pos-att/Solver_pos_att.m:452-500, ode_eq :692-757,
 * get_thruster_on_off_optimal :404-449, to_Moments_Forces :805-823, ECI2body :825-829,
 * RSW2ECI :831-847) and Solver_attitude.get_optimal_path_simplified_testode45
 * (attitude-control/Solver_attitude.m:1669-1705, ode_eq :1795-1851).
 *
 * ode45 itself is MathWorks' Dormand-Prince 5(4) driver.  It is not part of the reference repository;
 * what is restated here is its published algorithm (Dormand & Prince 1980 tableau; step control as
 * described by Shampine & Reichelt, "The MATLAB ODE Suite", SIAM J. Sci. Comput. 18, 1997, and as the
 * readable ode45.m / odearguments.m shipped with MATLAB implement it) with the DEFAULT options the
 * reference's two call sites use: RelTol 1e-3, AbsTol 1e-6, MaxStep 0.1*|tf-t0|, no InitialStep,
 * component-wise error control, tspan = [t0 tf] (only the last row of the output is used).
 * Parity unpinned: the reference stores no output of this path and MATLAB cannot run here. */
typedef void (*ode_rhs)(const void *ctx, double t, const double *y, double *dydt);
#define ODE_MAXN 13
static const double dp_A[6] = {1. / 5, 3. / 10, 4. / 5, 8. / 9, 1, 1};
static const double dp_B[7][6] = {                       /* dp_B[j][k]: slope j in the argument of stage k + 2 */
    {1. / 5, 3. / 40, 44. / 45, 19372. / 6561, 9017. / 3168, 35. / 384},
    {0, 9. / 40, -56. / 15, -25360. / 2187, -355. / 33, 0},
    {0, 0, 32. / 9, 64448. / 6561, 46732. / 5247, 500. / 1113},
    {0, 0, 0, -212. / 729, 49. / 176, 125. / 192},
    {0, 0, 0, 0, -5103. / 18656, -2187. / 6784},
    {0, 0, 0, 0, 0, 11. / 84},
    {0, 0, 0, 0, 0, 0}};
static const double dp_E[7] = {71. / 57600, 0, -71. / 16695, 71. / 1920, -17253. / 339200, 22. / 525, -1. / 40};

/* [~, Y] = ode45(fn, [t0 tf], y); y = Y(end,:).  Returns the number of accepted steps; *nfailed counts
 * rejected attempts; *warned = 1 when the step size reached hmin (MATLAB warns and returns early) or
 * max_steps was hit (a bound the GPU kernel needs; never reached by the reference's plants). */
int oracle_ode45_last(ode_rhs fn, const void *ctx, int neq, double t0, double tf, double *y, double rtol,
                      double atol, int max_steps, int *nfailed, int *warned)
{
    const double pw = 1. / 5;
    double f[7][ODE_MAXN], ys[ODE_MAXN];
    const double htspan = fabs(tf - t0), hmax = fabs(0.1 * (tf - t0)), threshold = atol / rtol;
    double t = t0;
    int steps = 0, done = 0;
    *nfailed = 0;
    *warned = 0;
    if (neq > ODE_MAXN) return -1;
    fn(ctx, t, y, f[0]);
    double hmin = 16 * eps_of(t);
    double absh = fmin(hmax, htspan);                     /* initial step from y'(t0) */
    double rh = 0;
    for (int i = 0; i < neq; ++i) rh = fmax(rh, fabs(f[0][i] / fmax(fabs(y[i]), threshold)));
    rh = rh / (0.8 * pow(rtol, pw));
    if (absh * rh > 1) absh = 1 / rh;
    absh = fmax(absh, hmin);
    while (!done) {
        if (steps >= max_steps) { *warned = 1; break; }
        hmin = 16 * eps_of(t);
        absh = fmin(hmax, fmax(hmin, absh));
        double h = absh;                                  /* tdir = +1 */
        if (1.1 * absh >= fabs(tf - t)) {                 /* stretch the step if within 10% of tf - t */
            h = tf - t;
            absh = fabs(h);
            done = 1;
        }
        int nofailed = 1;
        double err, tnew;
        for (;;) {
            for (int k = 0; k < 6; ++k) {                 /* y + f*hB(:,k), hB = h*B */
                for (int i = 0; i < neq; ++i) {
                    double s = 0;
                    for (int j = 0; j <= k; ++j) s = s + f[j][i] * (h * dp_B[j][k]);
                    ys[i] = y[i] + s;
                }
                if (k < 5) fn(ctx, t + h * dp_A[k], ys, f[k + 1]);
            }
            tnew = t + h * dp_A[5];
            if (done) tnew = tf;                          /* hit the end point exactly */
            fn(ctx, tnew, ys, f[6]);                      /* ys = ynew */
            err = 0;
            for (int i = 0; i < neq; ++i) {
                double fe = 0;
                for (int j = 0; j < 7; ++j) fe = fe + f[j][i] * dp_E[j];
                err = fmax(err, fabs(fe / fmax(fmax(fabs(y[i]), fabs(ys[i])), threshold)));
            }
            err = absh * err;
            if (err > rtol) {                             /* failed step */
                ++*nfailed;
                if (absh <= hmin) { *warned = 1; return steps; }
                if (nofailed) {
                    nofailed = 0;
                    absh = fmax(hmin, absh * fmax(0.1, 0.8 * pow(rtol / err, pw)));
                } else {
                    absh = fmax(hmin, 0.5 * absh);
                }
                h = absh;
                done = 0;
            } else {
                break;
            }
        }
        ++steps;
        if (!done && nofailed) {                          /* no failures: new step size */
            const double temp = 1.25 * pow(err / rtol, pw);
            if (temp > 0.2) absh = absh / temp;
            else absh = 5.0 * absh;
        }
        t = tnew;
        memcpy(y, ys, sizeof(double) * neq);
        memcpy(f[0], f[6], sizeof(double) * neq);         /* first-same-as-last */
    }
    return steps;
}

/* test hook: ode45 on y' = lambda*y (neq components with lambda[i]) */
static void linear_rhs(const void *ctx, double t, const double *y, double *dydt)
{
    const double *lam = (const double *)ctx;
    (void)t;
    for (int i = 0; i < (int)lam[0]; ++i) dydt[i] = lam[1 + i] * y[i];
}
int oracle_ode45_linear(int neq, const double *lambda, double t0, double tf, double *y, double rtol, double atol,
                        int *nfailed, int *warned)
{
    double ctx[1 + ODE_MAXN];
    ctx[0] = neq;
    for (int i = 0; i < neq; ++i) ctx[1 + i] = lambda[i];
    return oracle_ode45_last(linear_rhs, ctx, neq, t0, tf, y, rtol, atol, 1000000, nfailed, warned);
}

/* A\b for the symmetric inertia matrix with positive diagonal: mldivide takes the Cholesky path
 * (A = R'R, R upper).  A is column-major 3x3. */
typedef struct { double r11, r12, r13, r22, r23, r33; } chol3;
static void chol3_factor(const double *A, chol3 *c)
{
    c->r11 = sqrt(A[0]);
    c->r12 = A[3] / c->r11;
    c->r13 = A[6] / c->r11;
    c->r22 = sqrt(A[4] - c->r12 * c->r12);
    c->r23 = (A[7] - c->r12 * c->r13) / c->r22;
    c->r33 = sqrt(A[8] - (c->r13 * c->r13 + c->r23 * c->r23));
}
static void chol3_solve(const chol3 *c, const double *b, double *x)
{
    const double z1 = b[0] / c->r11;
    const double z2 = (b[1] - c->r12 * z1) / c->r22;
    const double z3 = ((b[2] - c->r13 * z1) - c->r23 * z2) / c->r33;
    x[2] = z3 / c->r33;
    x[1] = (z2 - c->r23 * x[2]) / c->r22;
    x[0] = ((z1 - c->r12 * x[1]) - c->r13 * x[2]) / c->r11;
}
/* A\b for a general 3x3 (row-major M[i][j]): LU with partial pivoting */
static void lu3_solve(const double M[3][3], const double *b, double *x)
{
    double a[3][4];
    for (int i = 0; i < 3; ++i) { for (int j = 0; j < 3; ++j) a[i][j] = M[i][j]; a[i][3] = b[i]; }
    for (int k = 0; k < 3; ++k) {
        int p = k;
        for (int i = k + 1; i < 3; ++i) if (fabs(a[i][k]) > fabs(a[p][k])) p = i;
        if (p != k) for (int j = 0; j < 4; ++j) { const double tmp = a[k][j]; a[k][j] = a[p][j]; a[p][j] = tmp; }
        for (int i = k + 1; i < 3; ++i) {
            const double l = a[i][k] / a[k][k];
            for (int j = k + 1; j < 4; ++j) a[i][j] = a[i][j] - l * a[k][j];
        }
    }
    x[2] = a[2][3] / a[2][2];
    x[1] = (a[1][3] - a[1][2] * x[2]) / a[1][1];
    x[0] = ((a[0][3] - a[0][1] * x[1]) - a[0][2] * x[2]) / a[0][0];
}
static void matvec3(const double M[3][3], const double *v, double *o)
{
    for (int i = 0; i < 3; ++i) o[i] = (M[i][0] * v[0] + M[i][1] * v[1]) + M[i][2] * v[2];
}
/* Solver_pos_att.m:825-829 */
static void eci2body(const double *q, double M[3][3])
{
    M[0][0] = 1 - 2 * (q[1] * q[1] + q[2] * q[2]); M[0][1] = 2 * (q[0] * q[1] + q[2] * q[3]); M[0][2] = 2 * (q[0] * q[2] - q[1] * q[3]);
    M[1][0] = 2 * (q[1] * q[0] - q[2] * q[3]); M[1][1] = 1 - 2 * (q[0] * q[0] + q[2] * q[2]); M[1][2] = 2 * (q[1] * q[2] + q[0] * q[3]);
    M[2][0] = 2 * (q[2] * q[0] + q[1] * q[3]); M[2][1] = 2 * (q[2] * q[1] - q[0] * q[3]); M[2][2] = 1 - 2 * (q[0] * q[0] + q[1] * q[1]);
}
/* Solver_pos_att.m:831-847: columns R, S, W */
static void rsw2eci(const double *pos, const double *vel, double M[3][3])
{
    const double np_ = sqrt((pos[0] * pos[0] + pos[1] * pos[1]) + pos[2] * pos[2]);
    const double c[3] = {pos[1] * vel[2] - pos[2] * vel[1], pos[2] * vel[0] - pos[0] * vel[2], pos[0] * vel[1] - pos[1] * vel[0]};
    const double nc = sqrt((c[0] * c[0] + c[1] * c[1]) + c[2] * c[2]);
    const double R[3] = {pos[0] / np_, pos[1] / np_, pos[2] / np_};
    const double W[3] = {c[0] / nc, c[1] / nc, c[2] / nc};
    const double S[3] = {W[1] * R[2] - W[2] * R[1], W[2] * R[0] - W[0] * R[2], W[0] * R[1] - W[1] * R[0]};
    for (int i = 0; i < 3; ++i) { M[i][0] = R[i]; M[i][1] = S[i]; M[i][2] = W[i]; }
}
/* obj.InertiaM\(U - cross(w, obj.InertiaM*w)); Im column-major */
static void euler_wdot(const double *Im, const chol3 *ch, const double *U, const double *w, double *wd)
{
    double Iw[3], b[3];
    for (int i = 0; i < 3; ++i) Iw[i] = (Im[i] * w[0] + Im[3 + i] * w[1]) + Im[6 + i] * w[2];
    b[0] = U[0] - (w[1] * Iw[2] - w[2] * Iw[1]);
    b[1] = U[1] - (w[2] * Iw[0] - w[0] * Iw[2]);
    b[2] = U[2] - (w[0] * Iw[1] - w[1] * Iw[0]);
    chol3_solve(ch, b, wd);
}

typedef struct { orbit_ctx o; double UM[3]; const double *Im; chol3 ch; } posatt_ctx;

/* Solver_pos_att.m:696-754 (system_dynamics) */
static void posatt_rates(const void *vc, double t, const double *X, double *Xd)
{
    const posatt_ctx *c = (const posatt_ctx *)vc;
    orbit_rates(&c->o, t, X, Xd);                         /* X_dot(1:6): same expressions as Solver_position's rates */
    const double q1 = X[6], q2 = X[7], q3 = X[8], q4 = X[9], w1 = X[10], w2 = X[11], w3 = X[12];
    Xd[6] = 0.5 * ((w3 * q2 - w2 * q3) + w1 * q4);
    Xd[7] = 0.5 * ((-w3 * q1 + w1 * q3) + w2 * q4);
    Xd[8] = 0.5 * ((w2 * q1 - w1 * q2) + w3 * q4);
    Xd[9] = 0.5 * ((-w1 * q1 - w2 * q2) - w3 * q3);
    euler_wdot(c->Im, &c->ch, c->UM, X + 10, Xd + 10);
}

/* par = {mu, h, rtol, atol, Mass, T_dist, R0[3], V0[3], InertiaM[9] column-major} (21 doubles).
 * d[ch], modes[ch] ([4]), idx[ch] ([S_ch], 0-based combination index), fv[ch] ([4][C_ch]: the channel's
 * f0/f1/f6/f7_allcomb vectors) for ch = x, y, z.  y0 [13][batch]; X_out [13][n_steps+1][batch];
 * F_out [12][n_steps][batch] (f0..f11, F_Th_Opt of :481); FM_out [6][n_steps][batch]
 * (a_x a_y a_z U_M', Force_Moment_log of :482); warn_out [batch]. */
int oracle_rollout_pos_att(const bellman_desc *dx, const bellman_desc *dy, const bellman_desc *dz,
                           const int32_t *mx, const int32_t *my, const int32_t *mz,
                           const int32_t *ix, const int32_t *iy, const int32_t *iz,
                           const double *fx, const double *fy, const double *fz, const double *par,
                           int n_steps, const double *y0, int batch, double *X_out, double *F_out,
                           double *FM_out, int32_t *warn_out)
{
    const bellman_desc *dd[3] = {dx, dy, dz};
    const int32_t *mm[3] = {mx, my, mz}, *ii[3] = {ix, iy, iz};
    const double *ff[3] = {fx, fy, fz};
    dimtab g[3][4];
    for (int p = 0; p < 3; ++p) {
        if (dd[p]->D != 4) return -1;
        for (int k = 0; k < 4; ++k) dimtab_init(&g[p][k], dd[p]->grid[k], dd[p]->n[k], mm[p][k]);
    }
    const double mu = par[0], h = par[1], rtol = par[2], atol = par[3], Mass = par[4], T_dist = par[5];
    const double *R0 = par + 6, *V0 = par + 9, *Im = par + 12;
    double M1[3][3];
    rsw2eci(R0, V0, M1);
    static const int ang_of[3] = {1, 2, 0};               /* channel x: t_y, w_y; y: t_z, w_z; z: t_x, w_x (:432-447) */
    static const int thr_of[3][4] = {{0, 1, 6, 7}, {2, 3, 8, 9}, {4, 5, 10, 11}};
#pragma omp parallel for schedule(dynamic)
    for (int b = 0; b < batch; ++b) {
        double y[13];
        for (int k = 0; k < 13; ++k) y[k] = y0[(size_t)b * 13 + k];
        double *X = X_out + (size_t)b * 13 * (n_steps + 1);
        memcpy(X, y, sizeof(y));
        posatt_ctx c;
        c.o.mu = mu; c.o.R0 = R0; c.o.V0 = V0;
        c.Im = Im;
        chol3_factor(Im, &c.ch);
        int warns = 0;
        for (int ks = 1; ks <= n_steps; ++ks) {
            double tq[3], M2[3][3], tmp[3], xb[3], vb[3], f[12];
            for (int k = 0; k < 3; ++k) tq[k] = 2 * asin(y[6 + k]);      /* :472-474 */
            eci2body(y + 6, M2);
            matvec3(M1, y, tmp); matvec3(M2, tmp, xb);                   /* :414-415 */
            matvec3(M1, y + 3, tmp); matvec3(M2, tmp, vb);
            for (int p = 0; p < 3; ++p) {
                const bellman_desc *d = dd[p];
                const double xq[4] = {xb[p], vb[p], tq[ang_of[p]], y[10 + ang_of[p]]};
                int64_t o = 0, st = 1;
                for (int k = 0; k < 4; ++k) { o += nearest_node(&g[p][k], xq[k]) * st; st *= d->n[k]; }
                const int ci = ii[p][o];
                for (int m = 0; m < 4; ++m) f[thr_of[p][m]] = ff[p][(size_t)m * d->C + ci];
            }
            /* to_Moments_Forces :805-823 */
            const double UMy = (((f[0] - f[1]) + f[6]) - f[7]) * T_dist;
            const double UMz = (((f[2] - f[3]) + f[8]) - f[9]) * T_dist;
            const double UMx = (((f[4] - f[5]) + f[10]) - f[11]) * T_dist;
            const double ab[3] = {(((f[0] + f[1]) + f[6]) + f[7]) / Mass, (((f[2] + f[3]) + f[8]) + f[9]) / Mass,
                                  (((f[4] + f[5]) + f[10]) + f[11]) / Mass};
            double acc[3];
            lu3_solve(M2, ab, tmp);
            lu3_solve(M1, tmp, acc);
            c.UM[0] = UMx; c.UM[1] = UMy; c.UM[2] = UMz;
            for (int k = 0; k < 3; ++k) c.o.a[k] = acc[k];
            if (F_out) memcpy(F_out + ((size_t)b * n_steps + (ks - 1)) * 12, f, sizeof(f));
            if (FM_out) {
                double *fm = FM_out + ((size_t)b * n_steps + (ks - 1)) * 6;
                fm[0] = acc[0]; fm[1] = acc[1]; fm[2] = acc[2]; fm[3] = UMx; fm[4] = UMy; fm[5] = UMz;
            }
            int nf, w;
            oracle_ode45_last(posatt_rates, &c, 13, (double)(ks - 1) * h, (double)ks * h, y, rtol, atol, 100000, &nf, &w);
            warns += w;
            memcpy(X + (size_t)ks * 13, y, sizeof(y));
        }
        if (warn_out) warn_out[b] = warns;
    }
    for (int p = 0; p < 3; ++p)
        for (int k = 0; k < 4; ++k) free(g[p][k].rinv);
    return 0;
}

typedef struct { double U[3]; const double *Im; chol3 ch; } att_ctx;

/* Solver_attitude.m:1803-1849 (system_dynamics): X = (w1 w2 w3 q1 q2 q3 q4) */
static void att_rates(const void *vc, double t, const double *X, double *Xd)
{
    const att_ctx *c = (const att_ctx *)vc;
    (void)t;
    const double x1 = X[0], x2 = X[1], x3 = X[2], x4 = X[3], x5 = X[4], x6 = X[5], x7 = X[6];
    euler_wdot(c->Im, &c->ch, c->U, X, Xd);
    Xd[3] = 0.5 * ((x3 * x5 - x2 * x6) + x1 * x7);
    Xd[4] = 0.5 * ((-x3 * x4 + x1 * x6) + x2 * x7);
    Xd[5] = 0.5 * ((x2 * x4 - x1 * x5) + x3 * x7);
    Xd[6] = 0.5 * ((-x1 * x4 - x2 * x5) - x3 * x6);
}

/* Solver_attitude.m:1669-1705: U1(k) = FU_k(X(k), 2*asin(X(3+k))), then ode45 over one stage.
 * d: D = 2 (w, theta), problems 0..2 = the three axes; idx [3][S]; u_values [C];
 * par = {h, rtol, atol, InertiaM[9] column-major}.  y0 [7][batch]; X_out [7][n_steps+1][batch];
 * C_out [3][n_steps][batch]. */
int oracle_rollout_attitude(const bellman_desc *d, const int32_t *modes, const int32_t *idx, const double *u_values,
                            const double *par, int n_steps, const double *y0, int batch, double *X_out,
                            int32_t *C_out, int32_t *warn_out)
{
    if (d->D != 2 || d->P < 3) return -1;
    dimtab g[3][2];
    for (int p = 0; p < 3; ++p)
        for (int k = 0; k < 2; ++k) dimtab_init(&g[p][k], d->grid[k] + (size_t)p * d->n[k], d->n[k], modes[p * 2 + k]);
    const int64_t S = (int64_t)d->n[0] * d->n[1];
    const double h = par[0], rtol = par[1], atol = par[2], *Im = par + 3;
#pragma omp parallel for schedule(dynamic)
    for (int b = 0; b < batch; ++b) {
        double y[7];
        for (int k = 0; k < 7; ++k) y[k] = y0[(size_t)b * 7 + k];
        double *X = X_out + (size_t)b * 7 * (n_steps + 1);
        int32_t *Cc = C_out + (size_t)b * 3 * n_steps;
        memcpy(X, y, sizeof(y));
        att_ctx c;
        c.Im = Im;
        chol3_factor(Im, &c.ch);
        int warns = 0;
        for (int ks = 1; ks <= n_steps; ++ks) {
            for (int p = 0; p < 3; ++p) {
                const double th = 2 * asin(y[3 + p]);
                const int ci = idx[(size_t)p * S + nearest_node(&g[p][0], y[p]) + (int64_t)nearest_node(&g[p][1], th) * d->n[0]];
                Cc[(size_t)(ks - 1) * 3 + p] = ci;
                c.U[p] = u_values[ci];
            }
            int nf, w;
            oracle_ode45_last(att_rates, &c, 7, (double)(ks - 1) * h, (double)ks * h, y, rtol, atol, 100000, &nf, &w);
            warns += w;
            memcpy(X + (size_t)ks * 7, y, sizeof(y));
        }
        if (warn_out) warn_out[b] = warns;
    }
    for (int p = 0; p < 3; ++p)
        for (int k = 0; k < 2; ++k) free(g[p][k].rinv);
    return 0;
}

/* ==== the coupled 6-D attitude sweep ==========================================
 * Solver_attitude.run (attitude-control/Solver_attitude.m:521-601) with calculate_J_U_opt_state_M
 * (:767-823): F = griddedInterpolant({sr_1, sr_2, sr_3, s_yaw, s_pitch, s_roll}, ., 'linear') evaluated at
 * the precomputed next states of every (state, U1, U2, U3), added to J_current_state_fix, then
 * min over dim_U3, dim_U2, dim_U1 (nested minima = the first minimiser in the order U1 slowest, U3
 * fastest).  The reference never ran this path (one-argument method called with two, :282 vs :384;
 * 2.7e13-element arrays at the default mesh), so what is restated is the code as written, in fp64.
 * Normative arithmetic of the dense stage operator (include/bellman.h, bellman_dense6_run):
 *   per dimension the exact bin rule  cell = clamp(#{ s[i] <= x } - 1, 0, n-2),  t = (x - s[cell]) * rinv[cell]
 *   v   = 6-linear interpolation, dimension 0 reduced first, lerp(a,b,t) = fma(t, b - a, a)
 *   tot = (((gs + r1[u1]) + r2[u2]) + r3[u3]) + v        strict '<' in the order c = (u1*nu + u2)*nu + u3
 * w_next[d] is [nu][n0*n1*n2], a_next[d] and gs are [S] (dimension 0 fastest), idx_out = c. */
int oracle_dense6_run(const int32_t *n, int nu, const double *const *grid, const double *const *w_next,
                      const double *const *a_next, const double *gs, const double *const *r, int n_stages,
                      const double *J_N, double *J_out, int32_t *idx_out)
{
    int64_t S = 1, S3 = (int64_t)n[0] * n[1] * n[2], stride[6];
    for (int k = 0; k < 6; ++k) { stride[k] = S; S *= n[k]; }
    if (nu < 1 || nu > 8) return -1;
    dimtab g[6];
    for (int k = 0; k < 6; ++k) dimtab_init(&g[k], grid[k], n[k], BELLMAN_LOCATE_SEARCH);
    double *A = (double *)malloc(sizeof(double) * (size_t)S), *B = (double *)malloc(sizeof(double) * (size_t)S);
    if (J_N) memcpy(A, J_N, sizeof(double) * (size_t)S); else memset(A, 0, sizeof(double) * (size_t)S);
    for (int st = 0; st < n_stages; ++st) {
#pragma omp parallel for schedule(static)
        for (int64_t s = 0; s < S; ++s) {
            const int64_t s3 = s % S3;
            int cw[3][8], ca[3];
            double tw[3][8], ta[3];
            for (int d = 0; d < 3; ++d) {
                for (int u = 0; u < nu; ++u) cw[d][u] = locate(&g[d], w_next[d][(size_t)u * S3 + s3], &tw[d][u]);
                ca[d] = locate(&g[3 + d], a_next[d][s], &ta[d]);
            }
            const int64_t oa = ca[0] * stride[3] + ca[1] * stride[4] + ca[2] * stride[5];
            double best = INFINITY;
            int arg = 0;
            for (int u1 = 0; u1 < nu; ++u1)
                for (int u2 = 0; u2 < nu; ++u2)
                    for (int u3 = 0; u3 < nu; ++u3) {
                        const int64_t o = oa + cw[0][u1] + cw[1][u2] * stride[1] + cw[2][u3] * stride[2];
                        const double t[6] = {tw[0][u1], tw[1][u2], tw[2][u3], ta[0], ta[1], ta[2]};
                        double v[64];
                        for (int m = 0; m < 64; ++m) {
                            int64_t oo = o;
                            for (int d = 0; d < 6; ++d) if ((m >> d) & 1) oo += stride[d];
                            v[m] = A[oo];
                        }
                        for (int d = 0, len = 32; d < 6; ++d, len >>= 1)
                            for (int m = 0; m < len; ++m) v[m] = fma(t[d], v[2 * m + 1] - v[2 * m], v[2 * m]);
                        const double tot = (((gs[s] + r[0][u1]) + r[1][u2]) + r[2][u3]) + v[0];
                        if (tot < best) { best = tot; arg = (u1 * nu + u2) * nu + u3; }
                    }
            B[s] = best;
            idx_out[s] = arg;
        }
        double *tmp = A; A = B; B = tmp;
    }
    memcpy(J_out, A, sizeof(double) * (size_t)S);
    free(A); free(B);
    for (int k = 0; k < 6; ++k) free(g[k].rinv);
    return 0;
}

/* Solver_attitude.get_optimal_path (attitude-control/Solver_attitude.m:1487-1530), the consumer of the
 * 6-D policy: per step  [yaw, pitch, roll] = quat2angle([X7 X6 X5 X4])  (Aerospace Toolbox, 'ZYX', the
 * quaternion normalised first),  U_k = FU_k(w1, w2, w3, yaw, pitch, roll)  with 'nearest' interpolants over
 * U{1,2,3}_Opt,  X_next = next_stage_states(X, U, h, 'taylor')  (:1339-1371: X + h*f(X, U) with
 * spacecraft_dynamics_list :1199-1245, then the quaternion renormalised).
 * idx [S] holds c = (u1*nu + u2)*nu + u3 per grid node; x0 [7][batch]; X_out [7][n_steps+1][batch];
 * U_out [3][n_steps][batch]; Jd = {J1, J2, J3}. */
int oracle_rollout_attitude6(const int32_t *n, int nu, const double *const *grid, const int32_t *idx,
                             const double *u_values, const double *Jd, double h, int n_steps, const double *x0,
                             int batch, double *X_out, double *U_out)
{
    dimtab g[6];
    int64_t stride[6], S = 1;
    for (int k = 0; k < 6; ++k) { dimtab_init(&g[k], grid[k], n[k], BELLMAN_LOCATE_SEARCH); stride[k] = S; S *= n[k]; }
    const double J1 = Jd[0], J2 = Jd[1], J3 = Jd[2];
#pragma omp parallel for schedule(dynamic)
    for (int b = 0; b < batch; ++b) {
        double X[7];
        for (int k = 0; k < 7; ++k) X[k] = x0[(size_t)b * 7 + k];
        double *Xo = X_out + (size_t)b * 7 * (n_steps + 1), *Uo = U_out + (size_t)b * 3 * n_steps;
        memcpy(Xo, X, sizeof(X));
        for (int ks = 0; ks < n_steps; ++ks) {
            /* quat2angle([X7 X6 X5 X4]): q0 = X7 (scalar), q1 = X6, q2 = X5, q3 = X4 */
            const double qm = sqrt(((X[6] * X[6] + X[5] * X[5]) + X[4] * X[4]) + X[3] * X[3]);
            const double q0 = X[6] / qm, q1 = X[5] / qm, q2 = X[4] / qm, q3 = X[3] / qm;
            const double yaw = atan2(2 * (q1 * q2 + q0 * q3), ((q0 * q0 + q1 * q1) - q2 * q2) - q3 * q3);
            const double pitch = asin(-2 * (q1 * q3 - q0 * q2));
            const double roll = atan2(2 * (q2 * q3 + q0 * q1), ((q0 * q0 - q1 * q1) - q2 * q2) + q3 * q3);
            const double xq[6] = {X[0], X[1], X[2], yaw, pitch, roll};
            int64_t o = 0;
            for (int k = 0; k < 6; ++k) o += nearest_node(&g[k], xq[k]) * stride[k];
            const int c = idx[o];
            const double U[3] = {u_values[c / (nu * nu)], u_values[(c / nu) % nu], u_values[c % nu]};
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
            for (int k = 0; k < 7; ++k) X[k] = X[k] + h * d[k];
            const double qs = sqrt(((X[3] * X[3] + X[4] * X[4]) + X[5] * X[5]) + X[6] * X[6]);
            for (int k = 3; k < 7; ++k) X[k] = X[k] / qs;
            memcpy(Xo + (size_t)(ks + 1) * 7, X, sizeof(X));
        }
    }
    for (int k = 0; k < 6; ++k) free(g[k].rinv);
    return 0;
}
