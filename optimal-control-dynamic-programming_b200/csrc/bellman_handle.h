// bellman_handle.h — the opaque handle behind the C ABI (shared by bellman_api.cu / bellman_window.cu)
#pragma once
#include <cuda_runtime.h>
#include <nccl.h>   // types only; libnccl.so.2 is dlopen()ed lazily in bellman_comm_init

#include <string>
#include <vector>

#include "bellman_internal.h"

struct bellman_handle {
    bellman::HostProblem hp;
    std::string err;
    int device = 0;
    cudaStream_t stream = nullptr;
    // partition
    int part_dim = -1, rank = 0, nranks = 1;
    std::vector<bellman_slab> slabs;
    int own_n[bellman::MAXD], own_lo[bellman::MAXD], ext_lo[bellman::MAXD], ext_n[bellman::MAXD];
    long long stride[bellman::MAXD];
    // leading dimension of the J arrays (elements): ext_n[0], rounded up to even for D = 2 grids and
    // for dimension-0 slabs so that TMA's 16-byte stride rule holds for any grid / slab width
    int ld0 = 1;
    long long row_elems(int p) const {   // stored elements per outer index of a slab along dim p
        long long inner = 1;
        for (int k = 0; k < p; ++k) inner *= (k == 0 ? ld0 : hp.n[k]);
        return p == 0 ? (long long)ld0 : (long long)ext_n[p] * inner;
    }
    long long S_ext = 0, S_own = 0;
    // device memory
    double *d_tab = nullptr;          // all fp64 tables
    int32_t *d_mode = nullptr;        // [D][P]
    bellman::StageParams sp{};                 // template with table pointers filled in
    bool store_J_all = false, store_idx_all = false;
    double *d_J = nullptr;            // [N or 2][P][S_ext]
    int32_t *d_idx = nullptr;         // [N or 1][P][S_own]
    int cur_stage = 0;
    bool J_set = false;
    // check sums
    double *d_partials = nullptr, *d_sums = nullptr;
    int n_partials = 0;
    std::vector<double> check_log;    // triples
    // nccl
    ncclComm_t comm = nullptr;
    // fused halo mode: neighbours' J allocations mapped through CUDA IPC (index = rank, own = nullptr)
    bool fused_halo = false;
    bool group_mode = false;          // slab of a single-process group (bellman_group_init): peers are plain pointers, no NCCL
    std::vector<double *> peer_J;
    double *d_barrier = nullptr;
    uint32_t *d_flags = nullptr;      // tail of the J allocation: flags[q] = stages rank q has completed (fused halo)
    uint32_t halo_seq = 0;            // stages this rank has completed since bellman_comm_init
    unsigned char *d_comm_scratch = nullptr;   // IPC-handle all-gather buffer (partitioned handles)
    unsigned char *d_roll = nullptr;           // bellman_rollout's device buffers (grow-only, reused across calls)
    size_t d_roll_bytes = 0;
    // window kernel state
    bellman::WindowConfig wcfg;
    void *wstate = nullptr;           // bellman_window.cu: WindowState (tensor maps, chunk tables)
    void *tstate = nullptr;           // bellman_tile.cu: TileState (D = 3 / 4 TMA-staged tile kernel)
    void *sstate = nullptr;           // bellman_stream.cu: StreamState (D = 4 streaming factorised kernel)
    // cached 2-stage CUDA graphs (ping-pong storage has period 2), keyed by the parity of the
    // slot the first stage reads and by the kernel variant
    cudaGraphExec_t graph_exec[2] = {nullptr, nullptr};
    int graph_kernel = -1, graph_lanes = -1;
    // stats
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    double last_ms = 0.0, last_ms_exchange = 0.0;
    int64_t last_launches = 0;
    std::string last_kernel = "none";

    size_t slot_elems_J() const { return (size_t)hp.P * (size_t)S_ext; }
    size_t slot_elems_idx() const { return (size_t)hp.P * (size_t)S_own; }
    int J_slot(int stage) const { return store_J_all ? stage - 1 : ((hp.N - stage) & 1); }
    int idx_slot(int stage) const { return store_idx_all ? stage - 1 : 0; }
    double *J_ptr(int stage) const { return d_J + (size_t)J_slot(stage) * slot_elems_J(); }
    // idx storage is idx_bytes per element (hp.idx_bytes); pointers stay typed int32_t*, offsets are in elements
    int32_t *idx_ptr(int stage, int prob = 0) const {
        return reinterpret_cast<int32_t *>(reinterpret_cast<char *>(d_idx) +
                                           ((size_t)idx_slot(stage) * slot_elems_idx() + (size_t)prob * (size_t)S_own) * (size_t)hp.idx_bytes);
    }
};


namespace bellman {
// bellman_window.cu: plan the TMA-staged D = 2 kernel for this handle (sets h->wcfg, encodes one
// tensor map per J slot).  Leaves wcfg.valid = false when the problem does not qualify.
void window_setup(bellman_handle *h);
void window_teardown(bellman_handle *h);
// bellman_tile.cu: the D = 3 / 4 tile kernel (one TMA box per state tile); tstate stays null when
// the problem is not a narrow stencil
void tile_setup(bellman_handle *h);
void tile_teardown(bellman_handle *h);
bool tile_valid(const bellman_handle *h);
cudaError_t tile_launch_for_handle(bellman_handle *h, const StageParams &sp, int slot_next, cudaStream_t st);
// bellman_stream.cu: the streaming, factorised D = 4 kernel for Solver_pos_att's channel structure;
// sstate stays null when the problem does not have that structure
void stream_setup(bellman_handle *h);
void stream_teardown(bellman_handle *h);
bool stream_valid(const bellman_handle *h);
cudaError_t stream_launch_for_handle(bellman_handle *h, const StageParams &sp, int slot_next, cudaStream_t st);
// "window:strip" / "window:chain" / "window:ring-chain" / "window:ring": which TMA-staged kernel runs
const char *window_variant(const bellman_handle *h);
cudaError_t window_launch_tile_range(bellman_handle *h, const StageParams &sp, int slot_next, cudaStream_t st, int tj0, int ntj);
int window_wide_tiles(const bellman_handle *h, int *tile1);
cudaError_t window_launch_for_handle(bellman_handle *h, const StageParams &sp, int slot_next,
                                     cudaStream_t st);
}  // namespace bellman
