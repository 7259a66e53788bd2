// bellman_kernels.cuh — launcher prototypes (implemented in bellman_kernels.cu, sm_100a).
#pragma once
#include <cuda_runtime.h>

#include "bellman_internal.h"

namespace bellman {

// generic D = 2..4 stage: one thread per state, all controls looped in registers, J_{k+1}
// gathered through L1/L2.  Always applicable.
cudaError_t launch_stage_direct(const StageParams &sp, cudaStream_t st);

// D = 2 stage with controls split across the lanes of a warp-group and a lexicographic
// (value, index) shuffle reduction: for grids too small to fill the GPU with one thread per state.
cudaError_t launch_stage_splitc(const StageParams &sp, int lanes_per_state, cudaStream_t st);

// whole stage loop in one cooperative launch (small grids): every thread keeps its state for all
// stages, a grid-wide barrier separates stages.  persistent_capacity_threads() = resident threads.
int persistent_capacity_threads(int D);
cudaError_t launch_sweep_persistent(const StageParams &sp, double *J_base, int32_t *idx_base,
                                    long long J_slot_elems, long long idx_slot_elems, int store_J_all,
                                    int store_idx_all, int N, int stage_from, int n_stages, int lanes,
                                    unsigned int *d_barrier, cudaStream_t st);

// D = 2 stage with the J_{k+1} neighbourhood of a state tile staged in shared memory by TMA.
struct WindowLaunch {
    WindowConfig cfg;
    const void *tmap_next;   // device-resident CUtensorMap (128 B) describing J_next
};
cudaError_t launch_stage_window(const StageParams &sp, const WindowLaunch &wl, cudaStream_t st);
size_t window_smem_bytes(const WindowConfig &cfg);

// deterministic sum(J) and sum(idx+1) over the owned states (pos-att early-stop check)
cudaError_t launch_check_sums(const StageParams &sp, double *d_partials, int n_partials,
                              double *d_out2, cudaStream_t st);

// batched rollout, one thread per initial state
struct RolloutParams {
    const double *grid0, *rinv0, *grid1, *rinv1;
    double inv_h0, off0, inv_h1, off1;
    const int32_t *lut0, *lut1;
    int lut_n0, lut_n1;
    double lut_invw0, lut_invw1;
    int mode0, mode1, n0, n1, N, C, batch, mode, ssu_stage;
    const int32_t *idx_all;   // [N][S], idx_bytes per element
    int idx_bytes;
    const double *u_values;   // [C]
    double A[4], B[2];
    const double *x0;         // [batch][2]
    double *X_out;            // [batch][N][2]
    double *U_out;            // [batch][N]
};
cudaError_t launch_rollout(const RolloutParams &rp, cudaStream_t st);

// nearest-policy lookup / simplified-plant axis rollout (consumers of the sweep's output)
struct PolicyParams {
    const double *grid[MAXD], *rinv[MAXD];   // problem `prob` already applied
    const int32_t *lut[MAXD];
    double inv_h[MAXD], off[MAXD], lut_invw[MAXD];
    int mode[MAXD], n[MAXD], lut_n[MAXD];
    int own_lo[MAXD], own_n[MAXD];   // owned index range per dimension (the whole grid unless partitioned)
    int D, batch;
    const int32_t *idx;        // [S] policy of one stage, or [N][S_slot] when time varying (idx_bytes per element)
    int idx_bytes;
    long long idx_stage_stride;
    int time_varying, stage, rate_dim, n_steps;
    double h_step;
    const double *u_inc;       // [C]
    const double *x;           // [D][batch]
    int32_t *idx_out;          // lookup: [batch]; rollout: [n_steps][batch]
    double *X_out;             // rollout: [2][n_steps+1][batch]
};
cudaError_t launch_policy_lookup(const PolicyParams &pp, cudaStream_t st);

// orbital forward simulation of Solver_position.get_optimal_path (one thread per initial state)
struct OrbitParams {
    PolicyParams pol[3];       // nearest policy of axis p (D = 2; .idx = policy of the requested stage)
    double mu, R0[3], V0[3];   // gravitational parameter, target state vector at t = 0
    double h, tol;             // stage length, rkf45 tolerance
    int n_steps, batch, stride_out, max_rkf;
    const double *u_values;    // [C]
    const double *y0;          // [batch][6]
    double *X_out;             // [batch][n_steps / stride_out + 1][6]
    int32_t *C_out;            // [batch][n_steps / stride_out][3]
    int32_t *warn_out;         // [batch] rkf45 calls that stopped on the minimum step size
};
cudaError_t launch_rollout_orbit(const OrbitParams &op, cudaStream_t st);
cudaError_t launch_rollout_axis(const PolicyParams &pp, cudaStream_t st);

// full-plant forward simulations integrated with ode45 (one thread per initial state):
// kind 0 = Solver_pos_att.get_optimal_path (13 states, three 4-D channel policies),
// kind 1 = Solver_attitude.get_optimal_path_simplified_testode45 (7 states, three 2-D axis policies)
struct PlantParams {
    PolicyParams pol[3];       // .idx = policy of the requested stage
    int kind;
    int C[3];                  // controls per channel (kind 0: row length of fv[ch])
    const double *fv[3];       // kind 0: [4][C[ch]] thruster levels f0/f1/f6/f7_allcomb; kind 1: fv[0] = u_values [C]
    double mu, R0[3], V0[3];
    double h, rtol, atol;      // stage length, ode45 RelTol / AbsTol
    double Im[9];              // InertiaM, column-major
    double mass, t_dist;
    int n_steps, batch, stride_out, max_ode;
    const double *y0;          // [batch][NEQ]
    double *X_out;             // [batch][n_steps / stride_out + 1][NEQ]
    double *F_out;             // kind 0: [batch][n_out][12] thruster levels
    double *FM_out;            // kind 0: [batch][n_out][6] (a_x a_y a_z U_M_x U_M_y U_M_z), may be null
    int32_t *C_out;            // kind 1: [batch][n_out][3] control indices
    int32_t *warn_out;         // [batch] ode45 calls that stopped on the minimum step size / the step bound
};
cudaError_t launch_rollout_plant(const PlantParams &pl, cudaStream_t st);

}  // namespace bellman
