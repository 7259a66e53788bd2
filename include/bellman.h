/*
 * bellman.h — C ABI of libbellman.so: the backward Bellman value-iteration sweep on B200.
 *
 * This is the drop-in boundary for ONE path of abdolrezat/Optimal-Control-Dynamic-Programming:
 * the per-stage operator
 *
 *     [F.Values, idx] = min( J_current + F(x_next...), [], ctrl_dim )
 *
 * and the stage loop around it, i.e. (reference file:line)
 *     test/Dynamic_Solver.m:86-102 + 202-211        (run / J_state_M)
 *     position-control/Solver_position.m:132-141    (simplified_run)
 *     attitude-control/Solver_attitude.m:236-247    (simplified_run)
 *     pos-att/Solver_pos_att.m:270-286              (calculate_one_channel_U_Opt)
 * plus the batched forward rollout of test/Dynamic_Solver.m:108-145,191-194 (get_optimal_path).
 *
 * The reference has no FFI layer (it is 100 % MATLAB); these entry points are what a MEX gateway
 * (optimal-control-dynamic-programming_b200/matlab/bellman_mex.cpp) or a ctypes binding
 * (optimal-control-dynamic-programming_b200/_lib.py) binds.  See INTEGRATION.md.
 *
 * Conventions
 *   - plain C types only; every host pointer is owned by the caller and is copied during the call,
 *     never retained.  All device memory, streams and NCCL communicators live behind the handle.
 *   - arrays are COLUMN-MAJOR (dimension 0 fastest), exactly as MATLAB stores them; no transposes.
 *   - control indices are 0-based int32 (the MATLAB facade adds 1).
 *   - every function returns 0 on success or a negative bellman_status; text via
 *     bellman_last_error().  No exceptions or aborts cross the ABI.  There is NO CPU fallback:
 *     without a usable CUDA device bellman_create() fails with BELLMAN_ERR_CUDA.
 *   - a handle is not thread-safe; calls are synchronous on return unless stated otherwise.
 *
 * ---------------------------------------------------------------------------------------------
 * The stage operator (normative arithmetic; oracle/bellman_oracle.c restates it on the CPU)
 * ---------------------------------------------------------------------------------------------
 * All arithmetic is IEEE-754 binary64, round-to-nearest-even, one rounding per written operation,
 * no contraction except where fma() is written.  For every grid state i = (i_0 .. i_{D-1}):
 *
 *   base_d = Ta_d[i_{src_a[d]}]                         if Tb_d is absent
 *          = Ta_d[i_{src_a[d]}] + Tb_d[i_{src_b[d]}]    otherwise
 *   gs     = q_{o0}[i_{o0}] + q_{o1}[i_{o1}] + ...      (left-associated, o = q_order)
 *   for c = 0 .. C-1 (in this order):
 *       xq_d = base_d + Tc_d[c]     (or base_d when Tc_d is absent)
 *       (cell_d, t_d) = locate_d(xq_d)
 *       v    = multilinear interpolation of J_{k+1} at cells/weights, dimension 0 reduced first,
 *              each 1-D step  lerp(a,b,t) = fma(t, b - a, a);  queries outside the grid use the
 *              edge cell with t<0 or t>1 (linear extrapolation, griddedInterpolant's default)
 *       tot  = (gs + r[c]) + v
 *       if tot < best (strictly): best = tot, arg = c        -> first index wins ties (MATLAB min)
 *   J_k[i] = best ; idx_k[i] = arg
 *
 *   locate_d(x) -> (cell, t), by the locate mode of dimension d:
 *     BELLMAN_LOCATE_SEARCH  : cell = clamp( #{ i : s[i] <= x } - 1, 0, n-2 )    (exact bin rule)
 *                              t = (x - s[cell]) * rinv[cell],  rinv[i] = 1/(s[i+1]-s[i]) (IEEE division)
 *     BELLMAN_LOCATE_UNIFORM : the dimension is evaluated in CELL UNITS.  With
 *                                  inv_h = (n-1)/(s[n-1]-s[0]),  off = -(s[0]*inv_h)
 *                              the library rescales the three next-state tables once, on the host
 *                              (one rounding per entry):
 *                                  Ta'[i] = fma(Ta[i], inv_h, off),  Tb'[i] = Tb[i]*inv_h,  Tc'[c] = Tc[c]*inv_h
 *                              and the sums above are formed from the primed tables, so the query
 *                              g = (Ta' + Tb') + Tc' is already the fractional cell coordinate:
 *                                  cell = clamp( (int)floor(g), 0, n-2 ),   t = g - (double)cell
 *                              Cell and weight come from one number, so they can never disagree, and
 *                              no table is read per query.  The rounding error of g (~1e-12 cell at
 *                              n = 8192) is the same size as that of forming x' in state units and
 *                              locating it against rounded grid nodes, as the reference does.
 *   The library picks UNIFORM when every node lies within 1e-14 of the grid's range from the
 *   uniform formula s[0] + i*h (MATLAB linspace grids do, with ~100x margin), else SEARCH;
 *   bellman_query_locate() reports the choice so the oracle uses the same one.
 *   bellman_rollout() locates a free state x with g = fma(x, inv_h, off) (UNIFORM) or the bin rule.
 */
#ifndef BELLMAN_H
#define BELLMAN_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BELLMAN_ABI_VERSION 1
#define BELLMAN_MAX_DIM 4

typedef enum bellman_status {
    BELLMAN_OK = 0,
    BELLMAN_ERR_BAD_ARG = -1,
    BELLMAN_ERR_CUDA = -2,
    BELLMAN_ERR_NCCL = -3,
    BELLMAN_ERR_OOM = -4,
    BELLMAN_ERR_NOT_RUN = -5,   /* requested stage has not been computed / not stored */
    BELLMAN_ERR_STATE = -6      /* call sequence error (e.g. run past stage 1)          */
} bellman_status;

enum { BELLMAN_LOCATE_UNIFORM = 0, BELLMAN_LOCATE_SEARCH = 1 };
enum { BELLMAN_KERNEL_AUTO = 0, BELLMAN_KERNEL_DIRECT = 1, BELLMAN_KERNEL_WINDOW = 2,
       BELLMAN_KERNEL_SPLITC = 3, BELLMAN_KERNEL_TILE = 4 };
/* WINDOW and TILE both ask for the TMA-staged kernel of the problem's dimensionality (D = 2: the
   window kernels; D = 3 / 4: the tile kernel) and fall back to DIRECT when it does not apply. */

typedef struct bellman_handle bellman_handle;

/*
 * Problem descriptor: P independent problems ("axes" / "channels") that share grid SHAPE and
 * control count but have their own tables.  Every table pointer addresses P consecutive rows.
 *
 * Replaces the S x C arrays the reference precomputes and streams every stage:
 *   X_next_M1/M2, J_current_state      test/Dynamic_Solver.m:81-82,184-200
 *   x*_next, v*_next, J_current_*      position-control/Solver_position.m:113-128
 *   w*_next, t*_next, J_current_*      attitude-control/Solver_attitude.m:220-233
 *   x_next..w_next (repmat), J_current pos-att/Solver_pos_att.m:257-263,299-328,784-802
 * All of them are sums of 1-D tables; the facade evaluates the 1-D tables with the reference's
 * own operation order and the kernel forms the sums on the fly.
 */
typedef struct bellman_desc {
    int32_t struct_size;               /* = sizeof(bellman_desc)                               */
    int32_t D;                         /* state dimensions, 2..BELLMAN_MAX_DIM                 */
    int32_t n[BELLMAN_MAX_DIM];        /* grid points per dimension (>= 2)                     */
    int32_t C;                         /* number of discrete controls                          */
    int32_t P;                         /* independent problems sharing the shape (>= 1)        */
    int32_t N;                         /* horizon: stages are numbered N (terminal) .. 1       */
    const double *grid[BELLMAN_MAX_DIM];  /* [P][n[d]] strictly increasing grid vectors        */
    int32_t src_a[BELLMAN_MAX_DIM];    /* state dim that indexes Ta[d]                         */
    int32_t src_b[BELLMAN_MAX_DIM];    /* state dim that indexes Tb[d], -1 if Tb[d] absent     */
    const double *Ta[BELLMAN_MAX_DIM]; /* [P][n[src_a[d]]]                                     */
    const double *Tb[BELLMAN_MAX_DIM]; /* [P][n[src_b[d]]] or NULL                             */
    const double *Tc[BELLMAN_MAX_DIM]; /* [P][C] or NULL (dimension does not depend on control)*/
    int32_t q_order[BELLMAN_MAX_DIM];  /* order in which the state-cost terms are summed       */
    const double *q[BELLMAN_MAX_DIM];  /* [P][n[d]] state-cost term of dimension d             */
    const double *r;                   /* [P][C] control-cost term                             */
    int32_t store_J_all;               /* keep J_k of every stage on the device                */
    int32_t store_idx_all;             /* keep idx_k of every stage on the device              */
    int32_t device;                    /* CUDA ordinal, -1 = current device                    */
    /* slab partition of ONE dimension across ranks (one process per GPU); part_dim = -1: none */
    int32_t part_dim;
    int32_t rank;
    int32_t nranks;
    /* optional explicit slab boundaries along part_dim: rank r owns [part_cuts[r], part_cuts[r+1]),
       part_cuts[0] = 0, part_cuts[nranks] = n[part_dim], strictly increasing; NULL = equal slabs.  Lets
       the host balance slabs whose cost per index differs (e.g. from per-rank times of a trial run). */
    const int32_t *part_cuts;
    /* bytes per stored argmin on the device: 0 or 4 = int32 (default), 2 = uint16 (C <= 65536),
       1 = uint8 (C <= 256).  Only the device storage changes (store_idx_all on the 8192^2 x 512 grid:
       53 GB as int32, 13 GB as uint8; 17 instead of 20 bytes of traffic per state and stage); every
       entry point still takes and returns 0-based int32 indices. */
    int32_t idx_bytes;
} bellman_desc;

typedef struct bellman_run_opts {
    int32_t struct_size;     /* = sizeof(bellman_run_opts)                                      */
    int32_t kernel;          /* BELLMAN_KERNEL_*                                                */
    int32_t check_period;    /* >0: when (stage %% period)==0 form sum(J), sum(idx+1) and stop  */
    double  check_tol;       /*     once |sum(J) - previous sum(J)| < tol (Solver_pos_att.m:273-285) */
    int32_t use_graph;       /* 1: replay stages through a CUDA graph (small grids)             */
    int32_t sync_each_stage; /* 1: host-sync after every stage (per-stage timing prints)        */
} bellman_run_opts;

/* slab plan of one rank, computed on the host from the tables alone (no GPU needed) */
typedef struct bellman_slab {
    int32_t own_lo, own_hi;  /* owned index range [lo,hi) along part_dim                        */
    int32_t ext_lo, ext_hi;  /* range of J_{k+1} this rank reads (owned + halo), from an exact  */
                             /* reach analysis of the next-state tables                         */
} bellman_slab;

int  bellman_version(void);
const char *bellman_last_error(const bellman_handle *h);   /* h may be NULL: create() errors   */

/* host-only helpers (usable without a GPU) */
int  bellman_query_locate(const bellman_desc *d, int32_t mode_out[/*P*D*/]);
int  bellman_plan_slabs(const bellman_desc *d, int32_t part_dim, int32_t nranks,
                        bellman_slab slabs_out[/*nranks*/]);

/* Exact stencil bounds per dimension: over every state and control, cell(x'_d) - i_d lies in
 * [lo_out[d], hi_out[d]] (i_d = the state's own index along d; the upper interpolation corner is one
 * further).  The D = 3 / 4 tile kernel sizes its shared-memory box with these.  lo > hi marks a
 * dimension whose query does not depend on the state's own index (no bounded stencil). */
int  bellman_query_stencil(const bellman_desc *d, int32_t lo_out[/*D*/], int32_t hi_out[/*D*/]);

/* lifecycle */
int  bellman_create(const bellman_desc *d, bellman_handle **out);
void bellman_destroy(bellman_handle *h);

/* multi-GPU bootstrap: rank 0 calls get_unique_id, the host distributes the 128 bytes
   (torch.distributed / MPI / a file), every rank calls comm_init */
int  bellman_get_unique_id(void *id128_out);
int  bellman_comm_init(bellman_handle *h, const void *id128);
/* 1: halo states are stored straight into the neighbours' J buffers by the stage kernel (CUDA IPC
 * peer memory over NVLink); stages are ordered by neighbour-only release/acquire flags (each rank
 * publishes its stage count into its neighbours and waits for theirs), no collective per stage;
 * 0: grouped ncclSend/ncclRecv of the halo ranges after every stage (fallback, or BELLMAN_NO_P2P=1);
 * valid after bellman_comm_init */
int  bellman_halo_mode(const bellman_handle *h);

/* Single-process multi-GPU (one host thread drives every GPU — what a MATLAB interpreter calling
 * run(obj) needs): create n handles with part_dim >= 0, nranks = n, rank = r and device = the GPU of
 * slab r (several slabs may share one GPU), call bellman_group_init once, then run them TOGETHER with
 * bellman_group_run (bellman_run refuses such handles: a slab cannot advance alone).  No NCCL: the
 * stage kernels store halo values straight into the neighbouring slabs' buffers (CUDA peer access)
 * and stages are ordered by neighbour flags.  set_J / get_J / get_idx / get_points work per handle.
 * The check log (opts->check_period) holds the sums over all slabs, in every handle. */
int  bellman_group_init(bellman_handle **handles, int32_t n);
int  bellman_group_run(bellman_handle **handles, int32_t n, int32_t n_stages, const bellman_run_opts *opts);

/* sweep */
int  bellman_set_J(bellman_handle *h, const double *J_host /*[P][S] global, NULL = zeros*/);
/* Resume / load a saved controller: make `stage` (1..N) the current stage with value function
 * J_host ([P][S] global, NULL = zeros) and, optionally, its policy idx_host ([P][S_own], 0-based;
 * must be NULL for the terminal stage N).  This is the inverse of bellman_get_J / bellman_get_idx:
 * what Solver_pos_att.m:291 saves (F_gI.Values, U_Optimal_id) can be put back, the sweep continued
 * with bellman_run, or the policy queried with bellman_policy_lookup (Solver_pos_att.m:849-882). */
int  bellman_set_stage(bellman_handle *h, int32_t stage, const double *J_host, const int32_t *idx_host);
int  bellman_stage(bellman_handle *h);                      /* one backward stage, default opts */
/* One backward stage driven from HOST buffers, with the copies overlapped with the kernel: equivalent to
 *   bellman_set_J(h, J_next_host); bellman_run(h, 1, opts); bellman_get_J(h, N-1, J_out_host);
 *   bellman_get_idx(h, N-1, idx_out_host)
 * (the per-stage host round trip of a MATLAB loop that keeps F.Values on the host, Dynamic_Solver.m:86-102),
 * same results bit for bit.  J_next_host NULL = continue from the J already on the device (plain
 * bellman_run(1) + the two reads); J_out_host / idx_out_host may be NULL.  Where the stage kernel can run a
 * range of tiles (k_stage_wide: D = 2, P = 1, int32 indices; one rank, or slabs along dimension 0 with the
 * peer-memory halo, every rank calling it) the grid is processed in up to 16 slabs of
 * dimension 1: J_{k+1} goes up in column chunks, a slab starts as soon as the highest column it can query
 * has arrived (exact reach analysis), and finished slabs go down while later ones compute.  Every other
 * configuration runs the plain sequence.  Pinned host memory is needed for the overlap (pageable memory
 * works, serialised by the driver). */
int  bellman_stage_host(bellman_handle *h, const double *J_next_host, double *J_out_host, int32_t *idx_out_host,
                        const bellman_run_opts *opts);
int  bellman_run(bellman_handle *h, int32_t n_stages, const bellman_run_opts *opts);
int  bellman_current_stage(const bellman_handle *h);        /* stage number of the current J    */
int  bellman_get_J(bellman_handle *h, int32_t stage, double *J_host_out /*[P][S_own]*/);
int  bellman_get_idx(bellman_handle *h, int32_t stage, int32_t *idx_host_out /*[P][S_own]*/);
/* J and (idx_out != NULL) argmin of `stage` at listed states of problem `prob`: states[m] is a GLOBAL
 * linear index (dimension 0 fastest) and must lie in this rank's owned range.  For spot checks and
 * sub-array reads at grid sizes whose full arrays are impractical to copy to the host. */
int  bellman_get_points(bellman_handle *h, int32_t stage, int32_t prob, const int64_t *states, int64_t n,
                        double *J_out, int32_t *idx_out);
int  bellman_get_check_log(const bellman_handle *h, double *out /*[max][3]: stage,sumJ,sumIdx*/,
                           int32_t max_entries);
int  bellman_owned_range(const bellman_handle *h, bellman_slab *out);

/* timing of the last bellman_run, measured with CUDA events on the library's own stream:
   total ms, number of stage-kernel launches, and ms spent in halo exchange */
int  bellman_last_run_stats(const bellman_handle *h, double *ms_total, int64_t *kernel_launches,
                            double *ms_exchange);
/* name of the stage kernel the last run used: "direct", "splitc", "persistent", "tile", "stream" (the factorised
   D = 4 kernel), or "window:<variant>" (variant = wide | strip | chain | ring-chain | ring, the TMA-staged D = 2
   kernels) */
const char *bellman_last_kernel(const bellman_handle *h);

/* batched forward rollout of test/Dynamic_Solver.m:108-145 (D = 2, P = 1, store_idx_all):
 *   U(k) = interp2(u_values[idx_k], X(:,k));  X(:,k+1) = A*X(:,k) + B*U(k),  k = 1..N-1
 * mode 0: time-varying policy u_star(:,:,k);  mode 1 ('ssu'): fixed stage ssu_stage.
 * A is 2x2 column-major, B 2x1, u_values[C]; x0 is [2][batch] column-major;
 * X_out [2][N][batch] (batch slowest), U_out [N][batch]. */
int  bellman_rollout(bellman_handle *h, const double *A, const double *B, const double *u_values,
                     const double *x0, int32_t batch, int32_t mode, int32_t ssu_stage,
                     double *X_out, double *U_out);

/* ---- consumers of the sweep's output ("next" rows: policy lookup and simplified-plant rollout) ---- */

/* Batched 'nearest' policy lookup — griddedInterpolant({s1,..}, U_vector(idx), 'nearest') of
 * Solver_position.m:144-146, Solver_attitude.m:249-251, Solver_pos_att.m:851-861 evaluated at
 * `batch` query states of problem `prob`: x is [D][batch] column-major (batch slowest),
 * idx_out[batch] receives the 0-based control index stored at the nearest grid node of `stage`
 * (clamped outside the grid; an exact midpoint goes to the upper node — MATLAB's tie side is
 * undocumented).  On a partitioned handle the rank that owns the nearest node answers and every
 * other rank writes -1, so the element-wise maximum over the ranks is the lookup. */
int  bellman_policy_lookup(bellman_handle *h, int32_t prob, int32_t stage, const double *x,
                           int32_t batch, int32_t *idx_out);

/* Batched rollout of the SIMPLIFIED plant of the two-state axis problems under the nearest policy
 * (attitude-control/test/test_simplified.m:129-151 and its next_stage_states :273-310; the same
 * plant Solver_position discretises, position-control/Solver_position.m:152-186):
 *     c      = nearest-policy index at (x_0, x_1)               stage k if time_varying, else `stage`
 *     x_r'   = x_r + u_inc[c]                                    r = rate_dim (the state the control drives)
 *     x_o'   = x_o + h*(k1 + 2*k2 + 2*k3 + k4)/6,  k1 = x_r, k2 = x_r + k1*h/2, k3 = x_r + k2*h/2, k4 = x_r + k3*h
 * with MATLAB's operation order.  D = 2; x0 is [2][batch], X_out [2][n_steps+1][batch] (batch
 * slowest), C_out [n_steps][batch] the applied control indices.  Needs store_idx_all when
 * time_varying.  Single rank only. */
int  bellman_rollout_axis(bellman_handle *h, int32_t prob, int32_t time_varying, int32_t stage,
                          int32_t rate_dim, double h_step, const double *u_inc /*[C]*/,
                          const double *x0, int32_t batch, int32_t n_steps, double *X_out,
                          int32_t *C_out);

/* Orbital forward simulation of Solver_position.get_optimal_path (position-control/Solver_position.m:
 * 189-224): every stage k the three axis policies are evaluated at the nearest grid node,
 *     a_x = U1_Opt(x1, v1), a_y = U2_Opt(x2, v2), a_z = U3_Opt(x3, v3)                    (:215-217)
 * and the relative-motion equations of :259-309 (the target orbit propagated with the universal
 * Kepler equation, position-control/private/kepler_U.m, f_and_g.m, fDot_and_gDot.m, stumpC.m,
 * stumpS.m; update_RV_target :333-361) are integrated from tspan(k) to tspan(k+1) = k*h by the
 * adaptive Runge-Kutta-Fehlberg 4(5) of position-control/private/rkf45.m:49-118 (first step
 * (tf-t0)/100, growth <= 4x, stop below 16*eps(t)).  One GPU thread per initial state.
 * Needs D = 2 and P >= 3 (problems 0..2 = the x, y, z axes), the policy of `stage` available.
 * y0 is [6][batch] (dr, dv of each trajectory, batch slowest); X_out [6][n_steps/stride_out + 1][batch],
 * C_out [3][n_steps/stride_out][batch] (0-based control index of each axis at the stored stages),
 * warn_out [batch] (may be NULL) counts rkf45 calls that stopped on the minimum step size.
 * The transcendental functions are CUDA's: results agree with the CPU restatement to a tolerance. */
typedef struct bellman_orbit_opts {
    int32_t struct_size;     /* = sizeof(bellman_orbit_opts)                                   */
    int32_t n_steps;         /* stages to simulate (reference: ceil(T_final/h) - 1)            */
    int32_t stride_out;      /* store every stride_out-th stage; n_steps %% stride_out == 0    */
    int32_t max_rkf_steps;   /* bound on rkf45 iterations per stage, 0 = 100000               */
    double  mu;              /* gravitational parameter (398600)                               */
    double  R0[3], V0[3];    /* target state vector at t = 0 (get_target_R0V0, :313-331)       */
    double  h;               /* stage length obj.h                                             */
    double  tol;             /* rkf45 tolerance (rkf45.m:62: 1e-8)                             */
} bellman_orbit_opts;
int  bellman_rollout_orbit(bellman_handle *h, int32_t stage, const bellman_orbit_opts *o,
                           const double *u_values /*[C]*/, const double *y0, int32_t batch,
                           double *X_out, int32_t *C_out, int32_t *warn_out);

/* Full-plant forward simulations that the reference integrates with ode45 (one GPU thread per initial
 * state).  ode45 is restated from its published algorithm — Dormand-Prince 5(4), MATLAB's step control
 * with the default options the reference's call sites use (RelTol 1e-3, AbsTol 1e-6, MaxStep
 * (tf - t0)/10, initial step from y'(t0), tspan = [t0 tf], last output row taken) — parity unpinned
 * (MATLAB cannot run here; the reference stores no output of these paths).  Results agree with the CPU
 * restatement (oracle_rollout_pos_att / oracle_rollout_attitude) to a tolerance: pow / asin / cos / sin
 * are CUDA's. */
typedef struct bellman_plant_opts {
    int32_t struct_size;     /* = sizeof(bellman_plant_opts)                                    */
    int32_t n_steps;         /* stages to simulate (reference: N_stage - 1)                     */
    int32_t stride_out;      /* store every stride_out-th stage; n_steps %% stride_out == 0     */
    int32_t max_ode_steps;   /* bound on ode45 steps per stage, 0 = 100000                      */
    double  mu;              /* gravitational parameter (398600); pos-att only                  */
    double  R0[3], V0[3];    /* target state vector at t = 0 (get_target_R0V0); pos-att only    */
    double  h;               /* stage length obj.h                                              */
    double  rtol, atol;      /* ode45 RelTol / AbsTol (defaults 1e-3 / 1e-6)                    */
    double  inertia[9];      /* obj.InertiaM, column-major (symmetric, positive diagonal)       */
    double  mass, t_dist;    /* obj.Mass, obj.T_dist; pos-att only                              */
} bellman_plant_opts;

/* Solver_pos_att.get_optimal_path (pos-att/Solver_pos_att.m:452-500): every stage the twelve thruster
 * levels come from the three 4-D channel policies evaluated at the nearest grid node
 * (get_thruster_on_off_optimal :404-449, x and v first taken from the RSW frame to the body frame,
 * :411-415, ECI2body :825-829, RSW2ECI :831-847), moments and RSW accelerations follow
 * to_Moments_Forces (:805-823), and the 13-state plant of :696-754 — relative motion about the target
 * orbit (universal Kepler propagation), quaternion kinematics, Euler's equations with the full inertia
 * matrix — is integrated over [tspan(k), tspan(k+1)] by ode45 (:484).
 * hx, hy, hz: handles of the x, y, z channel problems (D = 4, problem 0 of each, same device, one
 * rank), the policy of stage[ch] available on each.  f_x / f_y / f_z: [4][C_ch] levels of the
 * channel's four thrusters per control combination (f0/f1/f6/f7_allcomb, set_controller :849-882).
 * y0 [13][batch] = (dr dv q(scalar last) w); X_out [13][n_steps/stride_out + 1][batch];
 * F_out [12][n_out][batch] (f0..f11, F_Th_Opt :481); FM_out [6][n_out][batch] (a_x a_y a_z U_M',
 * Force_Moment_log :482; may be NULL); warn_out [batch] (may be NULL) counts ode45 calls that stopped
 * on the minimum step size. */
int  bellman_rollout_pos_att(bellman_handle *hx, bellman_handle *hy, bellman_handle *hz,
                             const int32_t stage[3], const bellman_plant_opts *o, const double *f_x,
                             const double *f_y, const double *f_z, const double *y0, int32_t batch,
                             double *X_out, double *F_out, double *FM_out, int32_t *warn_out);

/* Solver_attitude.get_optimal_path_simplified_testode45 (attitude-control/Solver_attitude.m:1669-1705):
 * U(k) = FU_k(X(k), 2*asin(X(3+k))) from the three axis policies (D = 2: (w, theta), problems 0..2 of
 * h, nearest node), then ode45 over one stage on the 7-state plant of :1803-1849 (Euler's equations
 * with the full inertia matrix + quaternion kinematics).  y0 [7][batch] = (w1 w2 w3 q1 q2 q3 q4);
 * X_out [7][n_out + 1][batch]; C_out [3][n_out][batch] 0-based control indices; u_values [C]. */
int  bellman_rollout_attitude(bellman_handle *h, int32_t stage, const bellman_plant_opts *o,
                              const double *u_values, const double *y0, int32_t batch, double *X_out,
                              int32_t *C_out, int32_t *warn_out);

/* The coupled 6-D attitude sweep — Solver_attitude.run (attitude-control/Solver_attitude.m:521-601): six
 * state dimensions (w1 w2 w3 yaw pitch roll) and three controls with nu levels each.  The next state is
 * not a sum of 1-D tables (Euler's equations couple the rates, the angle update goes through a
 * quaternion: spacecraft_dynamics_taylor_estimate :825-925), so this path has its own descriptor and its
 * own fused stage kernel instead of bellman_desc.  It replaces the nine-dimensional arrays the reference
 * builds with repmat (:905-921) and calculate_J_U_opt_state_M (:767-823): J_current_state_fix +
 * F(X1_next .. X6_next), then min over dim_U3, dim_U2, dim_U1.  The reference never ran this path (a
 * one-argument method is called with two, :282 vs :384; the default mesh needs 2.7e13-element arrays):
 * the semantics are the code as written with that call fixed, in fp64 — parity unpinned.
 *
 * Dense stage operator (normative arithmetic, restated by oracle_dense6_run).  S = prod n, dimension 0
 * fastest, S3 = n0*n1*n2, s3 = the (w1, w2, w3) part of state s.  For every state s:
 *   xq_d(u) = w_next[d][u*S3 + s3]  (d = 0..2, u = level of control d),   xq_{3+d} = a_next[d][s]
 *   (cell, t) per dimension by the exact bin rule: cell = clamp(#{ i : grid[i] <= x } - 1, 0, n-2),
 *              t = (x - grid[cell]) * rinv[cell],  rinv[i] = 1/(grid[i+1] - grid[i])
 *   v   = 6-linear interpolation of J_{k+1}, dimension 0 reduced first, lerp(a,b,t) = fma(t, b - a, a),
 *         linear extrapolation outside the grid
 *   tot = (((gs[s] + r[0][u1]) + r[1][u2]) + r[2][u3]) + v
 *   strict '<' over c = (u1*nu + u2)*nu + u3 in increasing order: the first minimiser, which is what the
 *   three nested min calls of the reference select (U3 innermost).  J_k[s] = best, idx_k[s] = c. */
typedef struct bellman_dense6_desc {
    int32_t struct_size;          /* = sizeof(bellman_dense6_desc)                                   */
    int32_t n[6];                 /* grid points of w1 w2 w3 yaw pitch roll (>= 2 each)              */
    int32_t nu;                   /* levels of each of the three controls (1..8)                     */
    int32_t device;               /* CUDA ordinal, -1 = current device                               */
    const double *grid[6];        /* strictly increasing grid vectors                                */
    const double *w_next[3];      /* [nu][S3] X{1,2,3}_next (:829-833)                               */
    const double *a_next[3];      /* [S] X{4,5,6}_next: yaw, pitch, roll (:835-897)                  */
    const double *gs;             /* [S] state part of J_current_state_fix (:629-640)                */
    const double *r[3];           /* [nu] R_d * U_d^2                                                */
} bellman_dense6_desc;
/* Runs n_stages backward stages from J_N ([S], NULL = zeros: the reference's terminal cost) entirely on
 * the device and returns the last stage computed: J_out [S], idx_out [S] (c as above, 0-based).  ms_out
 * (may be NULL) receives the device time of the stage loop.  Stateless: every buffer is released before
 * the call returns; errors are reported through bellman_last_error(NULL). */
int  bellman_dense6_run(const bellman_dense6_desc *d, int32_t n_stages, const double *J_N, double *J_out,
                        int32_t *idx_out, float *ms_out);

/* Solver_attitude.get_optimal_path (attitude-control/Solver_attitude.m:1487-1530), the consumer of the 6-D
 * policy, for a batch of initial states (one GPU thread each): per step
 *   [yaw, pitch, roll] = quat2angle([X7 X6 X5 X4])      (Aerospace Toolbox, 'ZYX', input normalised)
 *   U_k = FU_k(w1, w2, w3, yaw, pitch, roll)             'nearest' interpolants over U{1,2,3}_Opt (:1505-1519)
 *   X   = next_stage_states(X, U, h, 'taylor')           X + h*f(X, U), quaternion renormalised (:1339-1371)
 * d supplies n, nu, grid and device (its table pointers are not read); idx [S] is the policy of
 * bellman_dense6_run (c = (u1*nu + u2)*nu + u3), u_values [nu], J123 = {J1, J2, J3} (the diagonal inertia
 * spacecraft_dynamics_list uses, :1199-1245).  x0 [7][batch] = (w1 w2 w3 q1 q2 q3 q4), X_out
 * [7][n_steps+1][batch], U_out [3][n_steps][batch].  An exact midpoint goes to the upper node. */
int  bellman_rollout_attitude6(const bellman_dense6_desc *d, const int32_t *idx, const double *u_values,
                               const double *J123, double h, int32_t n_steps, const double *x0, int32_t batch,
                               double *X_out, double *U_out);

#ifdef __cplusplus
}
#endif
#endif /* BELLMAN_H */
