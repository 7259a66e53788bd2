// bellman_api.cu — C ABI of libbellman.so (include/bellman.h): handles, device residency,
// the stage loop, NCCL halo exchange, copy-in/out.  One process drives one GPU.
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nccl.h>   // types only; libnccl.so.2 is dlopen()ed lazily in bellman_comm_init

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "bellman_handle.h"
#include "bellman_kernels.cuh"

using namespace bellman;

static thread_local std::string g_create_error;
namespace bellman { void set_global_error(const std::string &msg) { g_create_error = msg; } }   // bellman_dense6.cu

namespace bellman {
bool raise_smem_limit(const void *fn, size_t bytes) {
    static std::mutex mu;
    static std::map<std::pair<int, const void *>, size_t> limit;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return false;
    std::lock_guard<std::mutex> lock(mu);
    size_t &cur = limit[{dev, fn}];
    if (bytes <= cur) return true;
    if (cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    cur = bytes;
    return true;
}

}  // namespace bellman


#define CUDA_TRY(h, expr)                                                                     \
    do {                                                                                      \
        cudaError_t _e = (expr);                                                              \
        if (_e != cudaSuccess) {                                                              \
            (h)->err = std::string(#expr) + ": " + cudaGetErrorString(_e);                    \
            return _e == cudaErrorMemoryAllocation ? BELLMAN_ERR_OOM : BELLMAN_ERR_CUDA;      \
        }                                                                                     \
    } while (0)

// ---------------------------------------------------------------------------------------------
// NCCL, loaded on demand (single-GPU use never needs it)
// ---------------------------------------------------------------------------------------------
struct NcclApi {
    void *lib = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*Send)(const void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t,
                              cudaStream_t) = nullptr;
    ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
};

static NcclApi *nccl_api(std::string &err) {
    static NcclApi api;
    if (api.lib) return &api;
    const char *override_path = std::getenv("BELLMAN_NCCL_LIB");
    const char *names[] = {override_path, "libnccl.so.2", "libnccl.so"};
    for (const char *nm : names) {
        if (!nm) continue;
        api.lib = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
        if (api.lib) break;
    }
    if (!api.lib) { err = std::string("cannot dlopen libnccl: ") + dlerror(); return nullptr; }
#define LOAD(field, sym)                                                       \
    api.field = reinterpret_cast<decltype(api.field)>(dlsym(api.lib, sym));    \
    if (!api.field) { err = std::string("libnccl lacks ") + sym; api.lib = nullptr; return nullptr; }
    LOAD(GetUniqueId, "ncclGetUniqueId")
    LOAD(CommInitRank, "ncclCommInitRank")
    LOAD(CommDestroy, "ncclCommDestroy")
    LOAD(Send, "ncclSend")
    LOAD(Recv, "ncclRecv")
    LOAD(AllReduce, "ncclAllReduce")
    LOAD(AllGather, "ncclAllGather")
    LOAD(GroupStart, "ncclGroupStart")
    LOAD(GroupEnd, "ncclGroupEnd")
    LOAD(GetErrorString, "ncclGetErrorString")
#undef LOAD
    return &api;
}

static int stage_barrier(bellman_handle *h);

extern "C" int bellman_version(void) { return BELLMAN_ABI_VERSION; }

extern "C" const char *bellman_last_error(const bellman_handle *h) {
    return h ? h->err.c_str() : g_create_error.c_str();
}

// argmin storage narrower than int32: conversion to / from the ABI's int32 happens on the device
__global__ void k_widen_idx(const void *src, int bytes, long long n, int32_t *dst) {
    for (long long k = blockIdx.x * (long long)blockDim.x + threadIdx.x; k < n; k += (long long)gridDim.x * blockDim.x)
        dst[k] = idx_load(src, bytes, k);
}
__global__ void k_narrow_idx(const int32_t *src, int bytes, long long n, void *dst, int C, int *bad) {
    for (long long k = blockIdx.x * (long long)blockDim.x + threadIdx.x; k < n; k += (long long)gridDim.x * blockDim.x) {
        const int v = src[k];
        if (v < 0 || v >= C) *bad = 1;
        idx_store(dst, bytes, k, v);
    }
}

static int upload_tables(bellman_handle *h) {
    HostProblem &hp = h->hp;
    const int D = hp.D, P = hp.P;
    // layout: per dim: grid, rinv, Ta, Tb, Tc, q, loc ; then r
    std::vector<double> buf;
    size_t o_grid[MAXD], o_rinv[MAXD], o_Ta[MAXD], o_Tb[MAXD], o_Tc[MAXD], o_q[MAXD], o_loc[MAXD], o_r;
    auto push = [&](const std::vector<double> &v) {
        const size_t o = buf.size();
        buf.insert(buf.end(), v.begin(), v.end());
        while (buf.size() % 2) buf.push_back(0.0);   // keep 16-byte alignment of every table
        return o;
    };
    for (int d = 0; d < D; ++d) {
        o_grid[d] = push(hp.grid[d]);
        o_rinv[d] = push(hp.rinv[d]);
        o_Ta[d] = push(hp.Ta[d]);
        o_Tb[d] = hp.has_b[d] ? push(hp.Tb[d]) : 0;
        o_Tc[d] = hp.has_c[d] ? push(hp.Tc[d]) : 0;
        o_q[d] = push(hp.q[d]);
        std::vector<double> loc(2 * (size_t)P);
        for (int p = 0; p < P; ++p) {
            const bool uni = hp.mode[(size_t)p * D + d] == BELLMAN_LOCATE_UNIFORM;
            loc[2 * p] = uni ? hp.inv_h[d][p] : hp.lut_invw[d][p];
            loc[2 * p + 1] = uni ? hp.off[d][p] : hp.grid[d][(size_t)p * hp.n[d]];
        }
        o_loc[d] = push(loc);
    }
    o_r = push(hp.r);
    CUDA_TRY(h, cudaMalloc(&h->d_tab, buf.size() * sizeof(double)));
    CUDA_TRY(h, cudaMemcpy(h->d_tab, buf.data(), buf.size() * sizeof(double), cudaMemcpyHostToDevice));
    // int table: modes [D][P], then the bucket tables of every dimension
    std::vector<int32_t> modes((size_t)D * P);
    for (int d = 0; d < D; ++d)
        for (int p = 0; p < P; ++p) modes[(size_t)d * P + p] = hp.mode[(size_t)p * D + d];
    size_t o_lut[MAXD];
    for (int d = 0; d < D; ++d) { o_lut[d] = modes.size(); modes.insert(modes.end(), hp.lut[d].begin(), hp.lut[d].end()); }
    CUDA_TRY(h, cudaMalloc(&h->d_mode, modes.size() * sizeof(int32_t)));
    CUDA_TRY(h, cudaMemcpy(h->d_mode, modes.data(), modes.size() * sizeof(int32_t), cudaMemcpyHostToDevice));

    StageParams &sp = h->sp;
    std::memset(&sp, 0, sizeof(sp));
    sp.D = D; sp.C = hp.C; sp.P = P;
    sp.idx_bytes = hp.idx_bytes;
    sp.S_ext = h->S_ext; sp.S_own = h->S_own;
    for (int d = 0; d < D; ++d) {
        DimParams &dp = sp.dim[d];
        dp.grid = h->d_tab + o_grid[d];
        dp.rinv = h->d_tab + o_rinv[d];
        dp.Ta = h->d_tab + o_Ta[d];
        dp.Tb = hp.has_b[d] ? h->d_tab + o_Tb[d] : nullptr;
        dp.Tc = hp.has_c[d] ? h->d_tab + o_Tc[d] : nullptr;
        dp.q = h->d_tab + o_q[d];
        dp.loc = h->d_tab + o_loc[d];
        dp.mode = h->d_mode + (size_t)d * P;
        dp.lut = h->d_mode + o_lut[d];
        dp.lut_n = hp.lut_n[d];
        dp.n = hp.n[d];
        dp.src_a = hp.src_a[d];
        dp.src_b = hp.has_b[d] ? hp.src_b[d] : 0;
        dp.n_a = hp.n[hp.src_a[d]];
        dp.n_b = hp.has_b[d] ? hp.n[hp.src_b[d]] : 0;
        dp.own_n = h->own_n[d];
        dp.own_lo = h->own_lo[d];
        dp.ext_lo = h->ext_lo[d];
        dp.stride = h->stride[d];
        sp.q_order[d] = hp.q_order[d];
    }
    sp.r = h->d_tab + o_r;
    return BELLMAN_OK;
}

extern "C" int bellman_create(const bellman_desc *d, bellman_handle **out) {
    if (!out) { g_create_error = "out is NULL"; return BELLMAN_ERR_BAD_ARG; }
    *out = nullptr;
    bellman_handle *h = new bellman_handle();
    h->err = load_problem(d, h->hp);
    if (!h->err.empty()) { g_create_error = h->err; delete h; return BELLMAN_ERR_BAD_ARG; }
    HostProblem &hp = h->hp;
    auto fail = [&](int code) { g_create_error = h->err; bellman_destroy(h); return code; };

    // partition
    h->part_dim = d->part_dim;
    h->rank = d->part_dim >= 0 ? d->rank : 0;
    h->nranks = d->part_dim >= 0 ? d->nranks : 1;
    for (int k = 0; k < MAXD; ++k) { h->own_n[k] = 1; h->own_lo[k] = 0; h->ext_lo[k] = 0; h->ext_n[k] = 1; h->stride[k] = 0; }
    for (int k = 0; k < hp.D; ++k) { h->own_n[k] = hp.n[k]; h->ext_n[k] = hp.n[k]; }
    if (h->part_dim >= 0) {
        if (h->nranks < 1 || h->rank < 0 || h->rank >= h->nranks) { h->err = "bad rank/nranks"; return fail(BELLMAN_ERR_BAD_ARG); }
        h->slabs.resize(h->nranks);
        h->err = plan_slabs(hp, h->part_dim, h->nranks, h->slabs.data());
        if (!h->err.empty()) return fail(BELLMAN_ERR_BAD_ARG);
        const bellman_slab &m = h->slabs[h->rank];
        h->own_lo[h->part_dim] = m.own_lo;
        h->own_n[h->part_dim] = m.own_hi - m.own_lo;
        h->ext_lo[h->part_dim] = m.ext_lo;
        h->ext_n[h->part_dim] = m.ext_hi - m.ext_lo;
    } else {
        h->slabs.assign(1, bellman_slab{0, 0, 0, 0});
    }
    h->S_ext = 1; h->S_own = 1;
    // D = 2 grids (the TMA-staged kernels) and dimension-0 slabs keep an even leading dimension
    h->ld0 = (hp.D == 2 || h->part_dim == 0) ? ((h->ext_n[0] + 1) & ~1) : h->ext_n[0];
    for (int k = 0; k < hp.D; ++k) {
        h->stride[k] = h->S_ext;
        h->S_ext *= (k == 0) ? h->ld0 : h->ext_n[k];
        h->S_own *= h->own_n[k];
    }

    // device
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) {
        h->err = std::string("no usable CUDA device (there is no CPU fallback): ") + cudaGetErrorString(e);
        return fail(BELLMAN_ERR_CUDA);
    }
    if (d->device >= 0) {
        h->device = d->device;
        if (cudaSetDevice(h->device) != cudaSuccess) { h->err = "cudaSetDevice failed"; return fail(BELLMAN_ERR_CUDA); }
    } else if (cudaGetDevice(&h->device) != cudaSuccess) { h->err = "cudaGetDevice failed"; return fail(BELLMAN_ERR_CUDA); }
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, h->device) != cudaSuccess) { h->err = "cudaGetDeviceProperties failed"; return fail(BELLMAN_ERR_CUDA); }
    if (prop.major != 10) {
        h->err = "device is sm_" + std::to_string(prop.major) + std::to_string(prop.minor) +
                 "; this library carries sm_100a code only";
        return fail(BELLMAN_ERR_CUDA);
    }
    int rc;
#define TRY_RC(expr) if ((rc = (expr)) != BELLMAN_OK) return fail(rc)
    auto cu = [&](cudaError_t ce, const char *what) {
        if (ce == cudaSuccess) return (int)BELLMAN_OK;
        h->err = std::string(what) + ": " + cudaGetErrorString(ce);
        return (int)(ce == cudaErrorMemoryAllocation ? BELLMAN_ERR_OOM : BELLMAN_ERR_CUDA);
    };
    TRY_RC(cu(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking), "cudaStreamCreate"));
    TRY_RC(cu(cudaEventCreate(&h->ev0), "cudaEventCreate"));
    TRY_RC(cu(cudaEventCreate(&h->ev1), "cudaEventCreate"));
    TRY_RC(upload_tables(h));

    h->store_J_all = d->store_J_all != 0;
    h->store_idx_all = d->store_idx_all != 0;
    const size_t nJ = (h->store_J_all ? (size_t)hp.N : 2) * h->slot_elems_J();
    const size_t nI = (h->store_idx_all ? (size_t)hp.N : 1) * h->slot_elems_idx();
    // + 256 bytes: the halo flag words (one uint32 per rank) live in the tail of the J allocation, so the
    // single CUDA IPC handle of that allocation also maps them into the neighbours
    TRY_RC(cu(cudaMalloc(&h->d_J, nJ * sizeof(double) + 256), "cudaMalloc(J)"));
    TRY_RC(cu(cudaMemsetAsync(h->d_J + nJ, 0, 256, h->stream), "cudaMemset(flags)"));
    h->d_flags = reinterpret_cast<uint32_t *>(h->d_J + nJ);
    TRY_RC(cu(cudaMalloc(&h->d_idx, nI * (size_t)hp.idx_bytes), "cudaMalloc(idx)"));
    TRY_RC(cu(cudaMemsetAsync(h->d_idx, 0, nI * (size_t)hp.idx_bytes, h->stream), "cudaMemset(idx)"));
    h->n_partials = 592;
    TRY_RC(cu(cudaMalloc(&h->d_partials, 2 * h->n_partials * sizeof(double)), "cudaMalloc"));
    TRY_RC(cu(cudaMalloc(&h->d_sums, 2 * sizeof(double)), "cudaMalloc"));
    if (h->nranks > 1) {
        // scratch of bellman_comm_init (IPC handle all-gather) and the stage barrier: allocated here so
        // that an allocation failure is a clean create() error and never skips a collective
        TRY_RC(cu(cudaMalloc(&h->d_comm_scratch, sizeof(cudaIpcMemHandle_t) * (size_t)(h->nranks + 1) + 16), "cudaMalloc"));
        TRY_RC(cu(cudaMalloc(&h->d_barrier, 2 * sizeof(double)), "cudaMalloc"));
    }
#undef TRY_RC
    h->cur_stage = hp.N;
    rc = bellman_set_J(h, nullptr);
    if (rc != BELLMAN_OK) return fail(rc);
    window_setup(h);   // optional fast path; leaves wcfg.valid = false when it does not apply
    tile_setup(h);     // D = 3 / 4 counterpart
    stream_setup(h);   // D = 4, Solver_pos_att's channel structure: streaming factorised kernel
    *out = h;
    return BELLMAN_OK;
}

extern "C" void bellman_destroy(bellman_handle *h) {
    if (!h) return;
    if (h->stream) cudaStreamSynchronize(h->stream);
    if (h->fused_halo && !h->group_mode) {
        // nobody may still be storing into my J when it is freed
        stage_barrier(h);
        cudaStreamSynchronize(h->stream);
        for (double *pj : h->peer_J) if (pj) cudaIpcCloseMemHandle(pj);
    }
    cudaFree(h->d_barrier);
    cudaFree(h->d_comm_scratch);
    cudaFree(h->d_roll);
    if (h->comm) {
        std::string e;
        NcclApi *api = nccl_api(e);
        if (api) api->CommDestroy(h->comm);
    }
    cudaFree(h->d_tab); cudaFree(h->d_mode); cudaFree(h->d_J); cudaFree(h->d_idx);
    cudaFree(h->d_partials); cudaFree(h->d_sums);
    window_teardown(h);
    tile_teardown(h);
    stream_teardown(h);
    for (auto &g : h->graph_exec) if (g) cudaGraphExecDestroy(g);
    if (h->ev0) cudaEventDestroy(h->ev0);
    if (h->ev1) cudaEventDestroy(h->ev1);
    if (h->stream) cudaStreamDestroy(h->stream);
    delete h;
}

extern "C" int bellman_owned_range(const bellman_handle *h, bellman_slab *out) {
    if (!h || !out) return BELLMAN_ERR_BAD_ARG;
    if (h->part_dim >= 0) *out = h->slabs[h->rank];
    else *out = bellman_slab{0, 0, 0, 0};
    return BELLMAN_OK;
}

extern "C" int bellman_current_stage(const bellman_handle *h) { return h ? h->cur_stage : BELLMAN_ERR_BAD_ARG; }

// ---------------------------------------------------------------------------------------------
// copy-in / copy-out.  Along part_dim the device arrays hold [ext_lo, ext_hi); every other
// dimension is complete, so a slab is (outer) rows of (inner*len) contiguous elements.
// ---------------------------------------------------------------------------------------------
static void slab_geometry(const bellman_handle *h, long long &inner, long long &outer) {
    inner = 1; outer = 1;
    const int p = h->part_dim < 0 ? h->hp.D - 1 : h->part_dim;
    for (int k = 0; k < p; ++k) inner *= h->hp.n[k];
    for (int k = p + 1; k < h->hp.D; ++k) outer *= h->hp.n[k];
}

static int upload_J(bellman_handle *h, int stage, const double *J_host) {
    const HostProblem &hp = h->hp;
    double *dst = h->J_ptr(stage);
    if (!J_host) {
        CUDA_TRY(h, cudaMemsetAsync(dst, 0, h->slot_elems_J() * sizeof(double), h->stream));
    } else {
        long long inner, outer;
        slab_geometry(h, inner, outer);
        const int p = h->part_dim < 0 ? hp.D - 1 : h->part_dim;
        const size_t S = (size_t)hp.S();
        for (int pr = 0; pr < hp.P; ++pr) {
            const double *src = J_host + (size_t)pr * S + (size_t)h->ext_lo[p] * inner;
            if (p >= 1 && h->ld0 != hp.n[0]) {
                // padded leading dimension (D = 2, slab along dimension 1): copy row by row
                CUDA_TRY(h, cudaMemcpy2DAsync(dst + (size_t)pr * h->S_ext, (size_t)h->ld0 * 8, src, (size_t)hp.n[0] * 8,
                                              (size_t)hp.n[0] * 8, (size_t)h->ext_n[p], cudaMemcpyHostToDevice,
                                              h->stream));
                continue;
            }
            CUDA_TRY(h, cudaMemcpy2DAsync(dst + (size_t)pr * h->S_ext, (size_t)h->row_elems(p) * 8, src,
                                          (size_t)hp.n[p] * inner * 8, (size_t)h->ext_n[p] * inner * 8,
                                          (size_t)outer, cudaMemcpyHostToDevice, h->stream));
        }
    }
    return BELLMAN_OK;
}

extern "C" int bellman_set_J(bellman_handle *h, const double *J_host) {
    if (!h) return BELLMAN_ERR_BAD_ARG;
    h->cur_stage = h->hp.N;
    h->check_log.clear();
    int rc = upload_J(h, h->hp.N, J_host);
    if (rc != BELLMAN_OK) return rc;
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    h->J_set = true;
    return BELLMAN_OK;
}

extern "C" int bellman_set_stage(bellman_handle *h, int32_t stage, const double *J_host, const int32_t *idx_host) {
    if (!h) return BELLMAN_ERR_BAD_ARG;
    const HostProblem &hp = h->hp;
    if (stage < 1 || stage > hp.N || (stage == hp.N && idx_host)) {
        h->err = "bellman_set_stage: stage must be in 1..N (and the terminal stage N has no policy)";
        return BELLMAN_ERR_BAD_ARG;
    }
    if (idx_host && hp.idx_bytes == 4) {      // (narrower storage checks the range while narrowing, below)
        const size_t ne = h->slot_elems_idx();
        for (size_t k = 0; k < ne; ++k)       // the consumers index u_values / thruster tables with these
            if (idx_host[k] < 0 || idx_host[k] >= hp.C) { h->err = "bellman_set_stage: control index out of range"; return BELLMAN_ERR_BAD_ARG; }
    }
    h->cur_stage = stage;
    h->check_log.clear();
    int rc = upload_J(h, stage, J_host);
    if (rc != BELLMAN_OK) return rc;
    if (idx_host) {
        const size_t ne = h->slot_elems_idx();
        if (hp.idx_bytes == 4) {
            CUDA_TRY(h, cudaMemcpyAsync(h->idx_ptr(stage), idx_host, ne * sizeof(int32_t), cudaMemcpyHostToDevice, h->stream));
        } else {                           // the device stores 1 or 2 bytes per index: narrow there
            int32_t *tmp = nullptr;
            int *d_bad = nullptr, bad = 0;
            CUDA_TRY(h, cudaMalloc(&tmp, ne * sizeof(int32_t) + sizeof(int)));
            d_bad = reinterpret_cast<int *>(tmp + ne);
            cudaMemsetAsync(d_bad, 0, sizeof(int), h->stream);
            cudaMemcpyAsync(tmp, idx_host, ne * sizeof(int32_t), cudaMemcpyHostToDevice, h->stream);
            k_narrow_idx<<<592, 256, 0, h->stream>>>(tmp, hp.idx_bytes, (long long)ne, h->idx_ptr(stage), hp.C, d_bad);
            cudaMemcpyAsync(&bad, d_bad, sizeof(int), cudaMemcpyDeviceToHost, h->stream);
            cudaError_t e = cudaStreamSynchronize(h->stream);
            cudaFree(tmp);
            if (e != cudaSuccess) { h->err = cudaGetErrorString(e); return BELLMAN_ERR_CUDA; }
            if (bad) { h->err = "bellman_set_stage: control index out of range"; return BELLMAN_ERR_BAD_ARG; }
        }
    }
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    h->J_set = true;
    return BELLMAN_OK;
}

static int stage_available(const bellman_handle *h, int stage, bool stored_all, bool is_idx) {
    const int N = h->hp.N;
    const int top = is_idx ? N - 1 : N;
    if (stage < 1 || stage > top) return BELLMAN_ERR_BAD_ARG;
    if (stage < h->cur_stage) return BELLMAN_ERR_NOT_RUN;
    if (!stored_all && stage != h->cur_stage) return BELLMAN_ERR_NOT_RUN;
    return BELLMAN_OK;
}

extern "C" int bellman_get_J(bellman_handle *h, int32_t stage, double *out) {
    if (!h || !out) return BELLMAN_ERR_BAD_ARG;
    int rc = stage_available(h, stage, h->store_J_all, false);
    if (rc != BELLMAN_OK) { h->err = "stage not available"; return rc; }
    const HostProblem &hp = h->hp;
    long long inner, outer;
    slab_geometry(h, inner, outer);
    const int p = h->part_dim < 0 ? hp.D - 1 : h->part_dim;
    const double *src0 = h->J_ptr(stage);
    for (int pr = 0; pr < hp.P; ++pr) {
        if (p >= 1 && h->ld0 != hp.n[0]) {   // padded leading dimension: row by row
            const double *srcp = src0 + (size_t)pr * h->S_ext + (size_t)(h->own_lo[p] - h->ext_lo[p]) * h->ld0;
            CUDA_TRY(h, cudaMemcpy2DAsync(out + (size_t)pr * h->S_own, (size_t)hp.n[0] * 8, srcp, (size_t)h->ld0 * 8,
                                          (size_t)hp.n[0] * 8, (size_t)h->own_n[p], cudaMemcpyDeviceToHost, h->stream));
            continue;
        }
        const double *src = src0 + (size_t)pr * h->S_ext + (size_t)(h->own_lo[p] - h->ext_lo[p]) * inner;
        CUDA_TRY(h, cudaMemcpy2DAsync(out + (size_t)pr * h->S_own, (size_t)h->own_n[p] * inner * 8, src,
                                      (size_t)h->row_elems(p) * 8, (size_t)h->own_n[p] * inner * 8,
                                      (size_t)outer, cudaMemcpyDeviceToHost, h->stream));
    }
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    return BELLMAN_OK;
}

extern "C" int bellman_get_idx(bellman_handle *h, int32_t stage, int32_t *out) {
    if (!h || !out) return BELLMAN_ERR_BAD_ARG;
    int rc = stage_available(h, stage, h->store_idx_all, true);
    if (rc != BELLMAN_OK) { h->err = "stage not available"; return rc; }
    const size_t ne = h->slot_elems_idx();
    const int ib = h->hp.idx_bytes;
    if (ib == 4) {
        CUDA_TRY(h, cudaMemcpyAsync(out, h->idx_ptr(stage), ne * sizeof(int32_t), cudaMemcpyDeviceToHost, h->stream));
        CUDA_TRY(h, cudaStreamSynchronize(h->stream));
        return BELLMAN_OK;
    }
    int32_t *tmp = nullptr;                // widened on the device, then copied as int32
    CUDA_TRY(h, cudaMalloc(&tmp, ne * sizeof(int32_t)));
    k_widen_idx<<<592, 256, 0, h->stream>>>(h->idx_ptr(stage), ib, (long long)ne, tmp);
    cudaMemcpyAsync(out, tmp, ne * sizeof(int32_t), cudaMemcpyDeviceToHost, h->stream);
    cudaError_t e = cudaStreamSynchronize(h->stream);
    cudaFree(tmp);
    if (e != cudaSuccess) { h->err = cudaGetErrorString(e); return BELLMAN_ERR_CUDA; }
    return BELLMAN_OK;
}

extern "C" int bellman_get_check_log(const bellman_handle *h, double *out, int32_t max_entries) {
    if (!h) return BELLMAN_ERR_BAD_ARG;
    const int n = (int)(h->check_log.size() / 3);
    const int m = std::min(n, (int)max_entries);
    if (out && m > 0) std::memcpy(out, h->check_log.data(), sizeof(double) * 3 * (size_t)m);
    return n;
}

// ---------------------------------------------------------------------------------------------
// NCCL bootstrap + halo exchange
// ---------------------------------------------------------------------------------------------
extern "C" int bellman_get_unique_id(void *id128_out) {
    if (!id128_out) return BELLMAN_ERR_BAD_ARG;
    std::string e;
    NcclApi *api = nccl_api(e);
    if (!api) { g_create_error = e; return BELLMAN_ERR_NCCL; }
    ncclUniqueId id;
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
    if (api->GetUniqueId(&id) != ncclSuccess) { g_create_error = "ncclGetUniqueId failed"; return BELLMAN_ERR_NCCL; }
    std::memcpy(id128_out, &id, 128);
    return BELLMAN_OK;
}

// Fused halo mode: map every neighbour's J allocation through CUDA IPC so the stage kernel can
// store halo states straight into peer memory over NVLink.  The 64-byte IPC handles travel through
// NCCL itself (all-gather), so the host language needs no extra plumbing.  Every rank takes the
// same decision (an all-reduce of the per-rank outcome); on any failure the library keeps the
// NCCL send/recv exchange.  BELLMAN_NO_P2P=1 forces the fallback.
static void setup_fused_halo(bellman_handle *h, NcclApi *api) {
    h->fused_halo = false;
    const int n = h->nranks;
    int ok = (n - 1 <= MAX_PEERS) && !std::getenv("BELLMAN_NO_P2P");
    cudaIpcMemHandle_t mine;
    std::memset(&mine, 0, sizeof(mine));
    if (ok && cudaIpcGetMemHandle(&mine, h->d_J) != cudaSuccess) { ok = 0; cudaGetLastError(); }
    const size_t hb = sizeof(cudaIpcMemHandle_t);
    // the scratch buffers were allocated by bellman_create (a failing cudaMalloc is reported there,
    // before any collective), so every rank always reaches both collectives below
    unsigned char *gbuf = h->d_comm_scratch;
    cudaMemcpyAsync(gbuf + hb * n, &mine, hb, cudaMemcpyHostToDevice, h->stream);
    std::vector<cudaIpcMemHandle_t> all(n);
    bool comm_ok = api->AllGather(gbuf + hb * n, gbuf, hb, ncclChar, h->comm, h->stream) == ncclSuccess;
    cudaMemcpyAsync(all.data(), gbuf, hb * n, cudaMemcpyDeviceToHost, h->stream);
    cudaStreamSynchronize(h->stream);
    h->peer_J.assign(n, nullptr);
    if (ok && comm_ok) {
        for (int q = 0; q < n && ok; ++q) {
            if (q == h->rank) continue;
            void *ptr = nullptr;
            if (cudaIpcOpenMemHandle(&ptr, all[q], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { ok = 0; cudaGetLastError(); }
            h->peer_J[q] = static_cast<double *>(ptr);
        }
    } else ok = 0;
    // unanimous decision
    double flag = ok ? 1.0 : 0.0;
    cudaMemcpyAsync(h->d_barrier, &flag, sizeof(double), cudaMemcpyHostToDevice, h->stream);
    if (api->AllReduce(h->d_barrier, h->d_barrier, 1, ncclDouble, ncclMin, h->comm, h->stream) != ncclSuccess) flag = 0.0;
    else {
        cudaMemcpyAsync(&flag, h->d_barrier, sizeof(double), cudaMemcpyDeviceToHost, h->stream);
        cudaStreamSynchronize(h->stream);
    }
    if (flag < 0.5) {
        for (double *&pj : h->peer_J) if (pj) { cudaIpcCloseMemHandle(pj); pj = nullptr; }
        return;
    }
    h->fused_halo = true;
}

// per-problem stride of rank q's J arrays (same padding rule as bellman_create)
static long long peer_S_ext(const bellman_handle *h, int q) {
    const HostProblem &hp = h->hp;
    const bellman_slab &o = h->slabs[q];
    long long st = 1;
    for (int d = 0; d < hp.D; ++d) {
        int ext = (d == h->part_dim) ? (o.ext_hi - o.ext_lo) : hp.n[d];
        if (d == 0 && (hp.D == 2 || h->part_dim == 0)) ext = (ext + 1) & ~1;
        st *= ext;
    }
    return st;
}

// ---------------------------------------------------------------------------------------------
// Fused halo mode, stage-to-stage ordering without a collective: neighbour-only release/acquire flags.
// Rank r owns flags[0..nranks): flags[q] = number of stages rank q has COMPLETED (stage kernel done,
// hence every store it made into r's halo).  After its stage kernel a rank runs k_halo_signal, which
// publishes its stage count into the flag word it owns inside every neighbour (st.release.sys over
// NVLink); before the next stage kernel it runs k_halo_wait, which spins (ld.acquire.sys) until every
// neighbour's count has reached the stage it depends on.  Only ranks whose slabs exchange halo are
// coupled; a rank never waits for the far end of the chain, and there is no NCCL launch per stage.
// Read-after-write: my next stage reads halo values the neighbour stored during its previous stage.
// Write-after-read: the neighbour's next stage overwrites (ping-pong slot) halo values my previous
// stage was still reading — it only starts that stage after seeing MY count, so both directions are
// covered by waiting on the union of "reads from me" and "I read from".
// ---------------------------------------------------------------------------------------------
struct HaloFlagArgs {
    uint32_t *remote[MAX_PEERS];        // signal: my word inside neighbour e
    const uint32_t *local[MAX_PEERS];   // wait: neighbour e's word inside my allocation
    int n;
    uint32_t seq;
    unsigned int *timeout_flag;         // set to 1 when a wait gives up (a neighbour stopped progressing)
    unsigned long long timeout_ns;
};

__global__ void k_halo_signal(const HaloFlagArgs a) {
    const int e = threadIdx.x;
    if (e >= a.n) return;
    __threadfence_system();
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(a.remote[e]), "r"(a.seq) : "memory");
}

__global__ void k_halo_wait(const HaloFlagArgs a) {
    const int e = threadIdx.x;
    if (e >= a.n) return;
    unsigned long long t0;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    for (;;) {
        uint32_t v;
        asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(a.local[e]) : "memory");
        if ((int32_t)(v - a.seq) >= 0) break;
        __nanosleep(200);
        unsigned long long t1;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
        if (t1 - t0 > a.timeout_ns) { atomicExch(a.timeout_flag, 1u); break; }
    }
    __threadfence_system();
}

static void halo_flag_args(const bellman_handle *h, HaloFlagArgs &a, uint32_t seq) {
    a.n = 0;
    a.seq = seq;
    a.timeout_flag = reinterpret_cast<unsigned int *>(h->d_flags + 48);     // a word of the tail no rank owns
    a.timeout_ns = 30ull * 1000000000ull;
    const bellman_slab &me = h->slabs[h->rank];
    const size_t nslots = h->store_J_all ? (size_t)h->hp.N : 2;
    for (int q = 0; q < h->nranks && a.n < MAX_PEERS; ++q) {
        if (q == h->rank) continue;
        const bellman_slab &o = h->slabs[q];
        const bool reads_me = std::max(o.ext_lo, me.own_lo) < std::min(o.ext_hi, me.own_hi);
        const bool i_read = std::max(me.ext_lo, o.own_lo) < std::min(me.ext_hi, o.own_hi);
        if (!reads_me && !i_read) continue;
        // the neighbour's flag words sit right after its J slots
        double *tail = h->peer_J[q] + nslots * (size_t)h->hp.P * (size_t)peer_S_ext(h, q);
        a.remote[a.n] = reinterpret_cast<uint32_t *>(tail) + h->rank;
        a.local[a.n] = h->d_flags + q;
        ++a.n;
    }
}

static int halo_signal(bellman_handle *h) {
    HaloFlagArgs a;
    halo_flag_args(h, a, h->halo_seq);
    if (a.n == 0) return BELLMAN_OK;
    k_halo_signal<<<1, 32, 0, h->stream>>>(a);
    if (cudaGetLastError() != cudaSuccess) { h->err = "halo signal launch failed"; return BELLMAN_ERR_CUDA; }
    return BELLMAN_OK;
}

static int halo_wait(bellman_handle *h) {
    HaloFlagArgs a;
    halo_flag_args(h, a, h->halo_seq);
    if (a.n == 0) return BELLMAN_OK;
    k_halo_wait<<<1, 32, 0, h->stream>>>(a);
    if (cudaGetLastError() != cudaSuccess) { h->err = "halo wait launch failed"; return BELLMAN_ERR_CUDA; }
    return BELLMAN_OK;
}

static void fill_peers(const bellman_handle *h, StageParams &sp, int out_stage) {
    sp.n_peers = 0;
    sp.part_dim = h->part_dim < 0 ? 0 : h->part_dim;
    if (!h->fused_halo) return;
    const HostProblem &hp = h->hp;
    const bellman_slab &me = h->slabs[h->rank];
    for (int q = 0; q < h->nranks; ++q) {
        if (q == h->rank) continue;
        const bellman_slab &o = h->slabs[q];
        const int lo = std::max(o.ext_lo, me.own_lo), hi = std::min(o.ext_hi, me.own_hi);
        if (lo >= hi) continue;
        PeerHalo &ph = sp.peer[sp.n_peers++];
        long long st = 1;
        for (int d = 0; d < MAXD; ++d) {
            ph.stride[d] = st;
            if (d < hp.D) {
                int ext = (d == h->part_dim) ? (o.ext_hi - o.ext_lo) : hp.n[d];
                // the neighbour's padded leading dimension: same rule as bellman_create's ld0
                if (d == 0 && (hp.D == 2 || h->part_dim == 0)) ext = (ext + 1) & ~1;
                st *= ext;
            }
        }
        ph.S_ext = st;
        ph.J = h->peer_J[q] + (size_t)h->J_slot(out_stage) * (size_t)hp.P * (size_t)st;
        ph.lo = lo; ph.hi = hi; ph.ext_lo = o.ext_lo; ph.pad_ = 0;
    }
}

// a 1-element all-reduce on the stream: the all-ranks barrier used once, before the peer mappings are
// closed in bellman_destroy (the per-stage ordering uses the neighbour flags above)
static int stage_barrier(bellman_handle *h) {
    NcclApi *api = nccl_api(h->err);
    if (!api) return BELLMAN_ERR_NCCL;
    if (api->AllReduce(h->d_barrier, h->d_barrier + 1, 1, ncclDouble, ncclSum, h->comm, h->stream) != ncclSuccess) {
        h->err = "stage barrier (ncclAllReduce) failed";
        return BELLMAN_ERR_NCCL;
    }
    return BELLMAN_OK;
}

extern "C" int bellman_halo_mode(const bellman_handle *h) { return h && h->fused_halo ? 1 : 0; }

extern "C" int bellman_comm_init(bellman_handle *h, const void *id128) {
    if (!h || !id128) return BELLMAN_ERR_BAD_ARG;
    if (h->nranks <= 1) return BELLMAN_OK;
    NcclApi *api = nccl_api(h->err);
    if (!api) return BELLMAN_ERR_NCCL;
    ncclUniqueId id;
    std::memcpy(&id, id128, 128);
    CUDA_TRY(h, cudaSetDevice(h->device));
    ncclResult_t r = api->CommInitRank(&h->comm, h->nranks, id, h->rank);
    if (r != ncclSuccess) { h->err = std::string("ncclCommInitRank: ") + api->GetErrorString(r); return BELLMAN_ERR_NCCL; }
    setup_fused_halo(h, api);
    return BELLMAN_OK;
}

// after stage `stage` has been written into its slot: fill the halo part of that slot
static int exchange_halo(bellman_handle *h, int stage) {
    if (h->nranks <= 1) return BELLMAN_OK;
    if (!h->comm) { h->err = "partitioned handle needs bellman_comm_init (or bellman_group_init) before running"; return BELLMAN_ERR_STATE; }
    NcclApi *api = nccl_api(h->err);
    if (!api) return BELLMAN_ERR_NCCL;
    const HostProblem &hp = h->hp;
    long long inner, outer;
    slab_geometry(h, inner, outer);
    const int p = h->part_dim;
    double *J = h->J_ptr(stage);
    const bellman_slab &me = h->slabs[h->rank];
    const long long row = h->row_elems(p);   // stored elements per outer index
    if (p >= 1) inner = inner / hp.n[0] * h->ld0;   // device stride of one index along p (leading dim may be padded)
    ncclResult_t r = api->GroupStart();
    for (int q = 0; q < h->nranks && r == ncclSuccess; ++q) {
        if (q == h->rank) continue;
        const bellman_slab &o = h->slabs[q];
        const int slo = std::max(o.ext_lo, me.own_lo), shi = std::min(o.ext_hi, me.own_hi);
        const int rlo = std::max(me.ext_lo, o.own_lo), rhi = std::min(me.ext_hi, o.own_hi);
        for (int pr = 0; pr < hp.P && r == ncclSuccess; ++pr)
            for (long long ou = 0; ou < outer && r == ncclSuccess; ++ou) {
                double *base = J + (size_t)pr * h->S_ext + (size_t)ou * row;
                if (slo < shi)
                    r = api->Send(base + (size_t)(slo - me.ext_lo) * inner, (size_t)(shi - slo) * inner,
                                  ncclDouble, q, h->comm, h->stream);
                if (r == ncclSuccess && rlo < rhi)
                    r = api->Recv(base + (size_t)(rlo - me.ext_lo) * inner, (size_t)(rhi - rlo) * inner,
                                  ncclDouble, q, h->comm, h->stream);
            }
    }
    ncclResult_t r2 = api->GroupEnd();
    if (r == ncclSuccess) r = r2;
    if (r != ncclSuccess) { h->err = std::string("halo exchange: ") + api->GetErrorString(r); return BELLMAN_ERR_NCCL; }
    return BELLMAN_OK;
}

// ---------------------------------------------------------------------------------------------
// the stage loop
// ---------------------------------------------------------------------------------------------
static int pick_kernel(bellman_handle *h, int requested, int &lanes) {
    const HostProblem &hp = h->hp;
    lanes = 1;
    if (requested == BELLMAN_KERNEL_DIRECT) return BELLMAN_KERNEL_DIRECT;
    const long long states = h->S_own * hp.P;
    auto pick_lanes = [&]() {
        int L = 1;
        while (L < 32 && 2 * L <= hp.C && states * L < 148LL * 2048) L *= 2;
        return L;
    };
    if (requested == BELLMAN_KERNEL_SPLITC) { lanes = std::max(2, pick_lanes()); return BELLMAN_KERNEL_SPLITC; }
    if (requested == BELLMAN_KERNEL_WINDOW || requested == BELLMAN_KERNEL_TILE)   // the TMA-staged kernels
        return h->wcfg.valid ? BELLMAN_KERNEL_WINDOW : (tile_valid(h) || stream_valid(h)) ? BELLMAN_KERNEL_TILE : BELLMAN_KERNEL_DIRECT;
    // AUTO
    if (h->wcfg.valid && states >= 148LL * 2048) return BELLMAN_KERNEL_WINDOW;
    if ((tile_valid(h) || stream_valid(h)) && states >= 148LL * 2048) return BELLMAN_KERNEL_TILE;
    lanes = pick_lanes();
    return lanes > 1 ? BELLMAN_KERNEL_SPLITC : BELLMAN_KERNEL_DIRECT;
}

static int launch_one_stage(bellman_handle *h, int kernel, int lanes) {
    const int from = h->cur_stage, to = from - 1;
    StageParams sp = h->sp;
    sp.J_next = h->J_ptr(from);
    sp.J_out = h->J_ptr(to);
    sp.idx_out = h->idx_ptr(to);
    fill_peers(h, sp, to);
    cudaError_t e;
    if (kernel == BELLMAN_KERNEL_SPLITC) e = launch_stage_splitc(sp, lanes, h->stream);
    else if (kernel == BELLMAN_KERNEL_WINDOW) e = window_launch_for_handle(h, sp, h->J_slot(from), h->stream);
    else if (kernel == BELLMAN_KERNEL_TILE)
        e = stream_valid(h) ? stream_launch_for_handle(h, sp, h->J_slot(from), h->stream)
                            : tile_launch_for_handle(h, sp, h->J_slot(from), h->stream);
    else e = launch_stage_direct(sp, h->stream);
    if (e != cudaSuccess) { h->err = std::string("stage launch: ") + cudaGetErrorString(e); return BELLMAN_ERR_CUDA; }
    h->last_launches += 1;
    return BELLMAN_OK;
}

extern "C" int bellman_run(bellman_handle *h, int32_t n_stages, const bellman_run_opts *opts) {
    if (!h || n_stages < 0) return BELLMAN_ERR_BAD_ARG;
    bellman_run_opts o;
    std::memset(&o, 0, sizeof(o));
    if (opts) {
        if (opts->struct_size != (int32_t)sizeof(bellman_run_opts)) { h->err = "bellman_run_opts.struct_size mismatch"; return BELLMAN_ERR_BAD_ARG; }
        o = *opts;
    }
    if (h->cur_stage - n_stages < 1) { h->err = "run would pass stage 1"; return BELLMAN_ERR_STATE; }
    if (h->group_mode && h->nranks > 1) { h->err = "handles of a single-process group run through bellman_group_run"; return BELLMAN_ERR_STATE; }
    const HostProblem &hp = h->hp;
    CUDA_TRY(h, cudaSetDevice(h->device));
    int lanes = 1;
    const int kernel = pick_kernel(h, o.kernel, lanes);
    h->last_kernel = kernel == BELLMAN_KERNEL_WINDOW ? window_variant(h) : kernel == BELLMAN_KERNEL_TILE ? (stream_valid(h) ? "stream" : "tile")
                     : kernel == BELLMAN_KERNEL_SPLITC ? "splitc" : "direct";
    h->last_launches = 0;
    h->last_ms_exchange = 0.0;
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> xev;
    CUDA_TRY(h, cudaEventRecord(h->ev0, h->stream));

    const bool graphable = o.use_graph && h->nranks == 1 && !o.sync_each_stage && !h->store_J_all &&
                           !h->store_idx_all;
    int done = 0;
    int rc = BELLMAN_OK;
    double fsum_prev = 0.0;
    if (!h->check_log.empty()) fsum_prev = h->check_log[h->check_log.size() - 2];
    bool stop = false;
    while (done < n_stages && !stop) {
        // stages until the next check point (or the end)
        int span = n_stages - done;
        if (o.check_period > 0) {
            int next_check = ((h->cur_stage - 1) / o.check_period) * o.check_period;   // largest multiple < cur
            if (next_check >= 1) span = std::min(span, h->cur_stage - next_check);
        }
        // small single-GPU problems: the whole span in ONE cooperative launch (grid barrier per stage)
        // (only under BELLMAN_KERNEL_AUTO: an explicitly requested kernel is never silently replaced)
        if (o.use_graph && h->nranks == 1 && !o.sync_each_stage && span >= 2 && o.kernel == BELLMAN_KERNEL_AUTO &&
            !std::getenv("BELLMAN_NO_PERSISTENT")) {
            int L = 1;
            const long long states = h->S_own * hp.P;
            // only when a stage is launch-latency sized (< ~2M updates); bigger stages want every
            // resident thread busy and are better served by one launch per stage
            const int cap = states * hp.C <= 2000000 ? persistent_capacity_threads(hp.D) : 0;
            while (L < 32 && 2 * L <= hp.C && states * (2 * L) <= cap) L *= 2;
            if (states * L <= cap) {
                if (!h->d_barrier) CUDA_TRY(h, cudaMalloc(&h->d_barrier, 2 * sizeof(double)));
                StageParams sp = h->sp;
                sp.n_peers = 0; sp.part_dim = 0;
                cudaError_t e = launch_sweep_persistent(sp, h->d_J, h->d_idx, (long long)h->slot_elems_J(),
                                                        (long long)h->slot_elems_idx(), h->store_J_all ? 1 : 0,
                                                        h->store_idx_all ? 1 : 0, hp.N, h->cur_stage, span, L,
                                                        reinterpret_cast<unsigned int *>(h->d_barrier), h->stream);
                if (e != cudaSuccess) { h->err = std::string("persistent sweep launch: ") + cudaGetErrorString(e); return BELLMAN_ERR_CUDA; }
                h->last_kernel = "persistent";
                h->last_launches += 1;
                h->cur_stage -= span;
                done += span;
                span = 0;
            }
        }
        if (graphable && span >= 4) {
            // ping-pong storage has period 2: a captured pair of stages is replayed span/2 times.
            // The instantiated graph is cached on the handle, so only the first run pays for it.
            if (h->graph_kernel != kernel || h->graph_lanes != lanes) {
                for (auto &g : h->graph_exec) if (g) { cudaGraphExecDestroy(g); g = nullptr; }
                h->graph_kernel = kernel; h->graph_lanes = lanes;
            }
            const int par = h->J_slot(h->cur_stage);
            const int pairs = span / 2;
            if (!h->graph_exec[par]) {
                cudaGraph_t graph = nullptr;
                const int stage0 = h->cur_stage;
                const int64_t l0 = h->last_launches;
                CUDA_TRY(h, cudaStreamBeginCapture(h->stream, cudaStreamCaptureModeThreadLocal));
                rc = launch_one_stage(h, kernel, lanes);
                h->cur_stage -= 1;
                if (rc == BELLMAN_OK) rc = launch_one_stage(h, kernel, lanes);
                h->cur_stage = stage0;
                h->last_launches = l0;
                cudaError_t ce = cudaStreamEndCapture(h->stream, &graph);
                if (rc != BELLMAN_OK) { if (graph) cudaGraphDestroy(graph); return rc; }
                CUDA_TRY(h, ce);
                ce = cudaGraphInstantiate(&h->graph_exec[par], graph, 0);
                cudaGraphDestroy(graph);
                CUDA_TRY(h, ce);
            }
            for (int i = 0; i < pairs; ++i) CUDA_TRY(h, cudaGraphLaunch(h->graph_exec[par], h->stream));
            h->last_launches += 2LL * pairs;
            h->cur_stage -= 2 * pairs;
            done += 2 * pairs;
            span -= 2 * pairs;
        }
        for (int i = 0; i < span; ++i) {
            rc = launch_one_stage(h, kernel, lanes);
            if (rc != BELLMAN_OK) return rc;
            h->cur_stage -= 1;
            ++done;
            if (h->nranks > 1) {
                cudaEvent_t a, b;
                CUDA_TRY(h, cudaEventCreate(&a));
                CUDA_TRY(h, cudaEventCreate(&b));
                CUDA_TRY(h, cudaEventRecord(a, h->stream));
                if (h->fused_halo) {
                    // my stage count goes to the neighbours; then wait until theirs has reached it.  The
                    // wait sits right behind the signal (not in front of the next launch) so that
                    // bellman_run returns with every halo of the last stage in place.
                    h->halo_seq += 1;
                    rc = halo_signal(h);
                    if (rc == BELLMAN_OK) rc = halo_wait(h);
                } else {
                    rc = exchange_halo(h, h->cur_stage);
                }
                if (rc != BELLMAN_OK) return rc;
                CUDA_TRY(h, cudaEventRecord(b, h->stream));
                xev.emplace_back(a, b);
            }
            if (o.sync_each_stage) CUDA_TRY(h, cudaStreamSynchronize(h->stream));
        }
        if (o.check_period > 0 && h->cur_stage % o.check_period == 0) {
            StageParams sp = h->sp;
            sp.J_out = h->J_ptr(h->cur_stage);
            sp.idx_out = h->idx_ptr(h->cur_stage);
            cudaError_t e = launch_check_sums(sp, h->d_partials, h->n_partials, h->d_sums, h->stream);
            if (e != cudaSuccess) { h->err = cudaGetErrorString(e); return BELLMAN_ERR_CUDA; }
            if (h->nranks > 1) {
                NcclApi *api = nccl_api(h->err);
                if (!api) return BELLMAN_ERR_NCCL;
                if (api->AllReduce(h->d_sums, h->d_sums, 2, ncclDouble, ncclSum, h->comm, h->stream) != ncclSuccess) {
                    h->err = "ncclAllReduce failed"; return BELLMAN_ERR_NCCL;
                }
            }
            double sums[2];
            CUDA_TRY(h, cudaMemcpyAsync(sums, h->d_sums, sizeof(sums), cudaMemcpyDeviceToHost, h->stream));
            CUDA_TRY(h, cudaStreamSynchronize(h->stream));
            h->check_log.push_back((double)h->cur_stage);
            h->check_log.push_back(sums[0]);
            h->check_log.push_back(sums[1]);
            const double e_f = sums[0] - fsum_prev;
            fsum_prev = sums[0];
            if (std::fabs(e_f) < o.check_tol) stop = true;
        }
    }
    CUDA_TRY(h, cudaEventRecord(h->ev1, h->stream));
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    if (h->fused_halo && done > 0) {
        unsigned int timed_out = 0;
        CUDA_TRY(h, cudaMemcpy(&timed_out, h->d_flags + 48, sizeof(timed_out), cudaMemcpyDeviceToHost));
        if (timed_out) { h->err = "halo flag wait timed out: a neighbour rank stopped progressing"; return BELLMAN_ERR_NCCL; }
    }
    float ms = 0.f;
    CUDA_TRY(h, cudaEventElapsedTime(&ms, h->ev0, h->ev1));
    h->last_ms = ms;
    for (auto &pr : xev) {
        float x = 0.f;
        cudaEventElapsedTime(&x, pr.first, pr.second);
        h->last_ms_exchange += x;
        cudaEventDestroy(pr.first);
        cudaEventDestroy(pr.second);
    }
    return BELLMAN_OK;
}

extern "C" int bellman_stage(bellman_handle *h) { return bellman_run(h, 1, nullptr); }

// ---------------------------------------------------------------------------------------------
// One stage with host buffers, pipelined (the host <-> device copies of bellman_set_J / bellman_get_J /
// bellman_get_idx overlapped with the stage kernel).  The grid is cut into slabs of dimension-1 tiles.
// J_{k+1} goes up in column chunks on a copy stream; slab s is launched as soon as the chunk holding the
// highest column it can query (exact reach analysis) has arrived; its J_k and argmin go down on a second
// copy stream while later slabs compute.  Only k_stage_wide can run a tile range; every other
// configuration takes the plain sequence, with identical results.
// ---------------------------------------------------------------------------------------------
extern "C" int bellman_stage_host(bellman_handle *h, const double *J_next_host, double *J_out_host,
                                  int32_t *idx_out_host, const bellman_run_opts *opts) {
    if (!h) return BELLMAN_ERR_BAD_ARG;
    const HostProblem &hp = h->hp;
    bellman_run_opts o;
    std::memset(&o, 0, sizeof(o));
    if (opts) {
        if (opts->struct_size != (int32_t)sizeof(bellman_run_opts)) { h->err = "bellman_run_opts.struct_size mismatch"; return BELLMAN_ERR_BAD_ARG; }
        o = *opts;
    }
    int lanes = 1, tile1 = 0;
    const int ntile1 = window_wide_tiles(h, &tile1);
    // sharded handles: slabs along dimension 0 only (the tile ranges run along dimension 1), halo through the
    // peer stores + neighbour flags
    const bool shard_ok = h->nranks == 1 || (h->part_dim == 0 && h->fused_halo && !h->group_mode);
    const bool pipelined = ntile1 >= 8 && shard_ok && hp.P == 1 && hp.D == 2 && hp.idx_bytes == 4 &&
                           (o.kernel == BELLMAN_KERNEL_AUTO || o.kernel == BELLMAN_KERNEL_WINDOW) &&
                           pick_kernel(h, o.kernel, lanes) == BELLMAN_KERNEL_WINDOW && o.check_period == 0 &&
                           !std::getenv("BELLMAN_NO_HOST_PIPELINE");
    if (!pipelined) {
        int rc = BELLMAN_OK;
        if (J_next_host) rc = bellman_set_J(h, J_next_host);
        if (rc == BELLMAN_OK) rc = bellman_run(h, 1, opts);
        if (rc == BELLMAN_OK && J_out_host) rc = bellman_get_J(h, h->cur_stage, J_out_host);
        if (rc == BELLMAN_OK && idx_out_host) rc = bellman_get_idx(h, h->cur_stage, idx_out_host);
        return rc;
    }
    CUDA_TRY(h, cudaSetDevice(h->device));
    if (J_next_host) { h->cur_stage = hp.N; h->check_log.clear(); }
    else if (!h->J_set) { h->err = "bellman_stage_host: no J_{k+1} on the device and none given"; return BELLMAN_ERR_STATE; }
    if (h->cur_stage < 2) { h->err = "run would pass stage 1"; return BELLMAN_ERR_STATE; }
    const int from = h->cur_stage, to = from - 1;
    const int n0 = hp.n[0], n1 = hp.n[1];
    const int e0 = h->ext_lo[0], en0 = h->ext_n[0], o0 = h->own_lo[0], on0 = h->own_n[0];   // rows held / owned
    // slabs of dimension-1 tiles: 16 while a launch keeps >= 4 waves of CTAs, else 8, else 2 (measured end to end
    // on cfg 4: N = 1: 8 slabs 50.0 ms, 16 slabs 48.1, 24 slabs 50.0; N = 2: 4 / 8 / 16 slabs 27.0 / 25.6 / 24.8 ms;
    // N = 8 ran with 8); BELLMAN_HOST_SLABS overrides (2..32)
    const long long ctas = (long long)((h->own_n[0] + 31) / 32) * ntile1;
    int NSLAB = ctas >= 16LL * 148 * 4 && ntile1 >= 16 ? 16 : ctas >= 8LL * 148 * 2 ? 8 : 2;
    if (const char *e = std::getenv("BELLMAN_HOST_SLABS")) NSLAB = std::max(2, std::min(32, std::atoi(e)));
    const int tps = (ntile1 + NSLAB - 1) / NSLAB;                  // tiles per slab
    const int nslab = (ntile1 + tps - 1) / tps;
    cudaStream_t s_in = nullptr, s_out = nullptr;
    std::vector<cudaEvent_t> ev_in(nslab, nullptr), ev_k(nslab, nullptr);
    int rc = BELLMAN_OK;
    auto cleanup = [&]() {
        for (auto e : ev_in) if (e) cudaEventDestroy(e);
        for (auto e : ev_k) if (e) cudaEventDestroy(e);
        if (s_in) cudaStreamDestroy(s_in);
        if (s_out) cudaStreamDestroy(s_out);
    };
#define ST(expr) do { cudaError_t _e = (expr); if (_e != cudaSuccess) { h->err = std::string("bellman_stage_host: ") + cudaGetErrorString(_e); cleanup(); return BELLMAN_ERR_CUDA; } } while (0)
    ST(cudaStreamCreateWithFlags(&s_in, cudaStreamNonBlocking));
    ST(cudaStreamCreateWithFlags(&s_out, cudaStreamNonBlocking));
    for (int s = 0; s < nslab; ++s) {
        ST(cudaEventCreateWithFlags(&ev_in[s], cudaEventDisableTiming));
        ST(cudaEventCreateWithFlags(&ev_k[s], cudaEventDisableTiming));
    }
    ST(cudaEventRecord(h->ev0, h->stream));
    // columns [c_lo, c_hi) of slab s (whole tiles)
    auto col_lo = [&](int s) { return std::min(n1, s * tps * tile1); };
    double *dJ_next = h->J_ptr(from), *dJ_out = h->J_ptr(to);
    int32_t *d_idx = h->idx_ptr(to);
    if (J_next_host) {
        for (int s = 0; s < nslab; ++s) {
            const int c0 = col_lo(s), c1 = col_lo(s + 1);
            ST(cudaMemcpy2DAsync(dJ_next + (size_t)c0 * h->ld0, (size_t)h->ld0 * 8, J_next_host + (size_t)c0 * n0 + e0,
                                 (size_t)n0 * 8, (size_t)en0 * 8, (size_t)(c1 - c0), cudaMemcpyHostToDevice, s_in));
            ST(cudaEventRecord(ev_in[s], s_in));
        }
    }
    StageParams sp = h->sp;
    sp.J_next = dJ_next;
    sp.J_out = dJ_out;
    sp.idx_out = d_idx;
    fill_peers(h, sp, to);
    h->last_kernel = window_variant(h);
    h->last_launches = 0;
    h->last_ms_exchange = 0.0;
    int waited = -1;
    for (int s = 0; s < nslab; ++s) {
        const int c0 = col_lo(s), c1 = col_lo(s + 1);
        if (J_next_host) {
            // highest J_{k+1} column any state of this slab can touch, plus the window's fixed extent (the TMA
            // box is loaded whole); the stream waits for the chunk that holds it
            int rlo = 0, rhi = 0;
            reach_range(hp, 1, c0, c1, rlo, rhi);
            const int top = std::min(n1 - 1, rhi + h->wcfg.win1);
            int need = nslab - 1;
            for (int q = 0; q < nslab; ++q) if (top < col_lo(q + 1)) { need = q; break; }
            if (need > waited) { ST(cudaStreamWaitEvent(h->stream, ev_in[need], 0)); waited = need; }
        }
        ST(window_launch_tile_range(h, sp, h->J_slot(from), h->stream, s * tps, std::min(tps, ntile1 - s * tps)));
        h->last_launches += 1;
        ST(cudaEventRecord(ev_k[s], h->stream));
        if (J_out_host || idx_out_host) ST(cudaStreamWaitEvent(s_out, ev_k[s], 0));
        if (J_out_host)
            ST(cudaMemcpy2DAsync(J_out_host + (size_t)c0 * on0, (size_t)on0 * 8, dJ_out + (size_t)c0 * h->ld0 + (o0 - e0),
                                 (size_t)h->ld0 * 8, (size_t)on0 * 8, (size_t)(c1 - c0), cudaMemcpyDeviceToHost, s_out));
        if (idx_out_host)
            ST(cudaMemcpyAsync(idx_out_host + (size_t)c0 * on0, d_idx + (size_t)c0 * on0, (size_t)(c1 - c0) * on0 * sizeof(int32_t),
                               cudaMemcpyDeviceToHost, s_out));
    }
    if (h->nranks > 1) {
        // as bellman_run: my stage count goes to the neighbours, then wait until theirs has reached it (their
        // peer stores of this stage are then in my halo rows)
        h->halo_seq += 1;
        rc = halo_signal(h);
        if (rc == BELLMAN_OK) rc = halo_wait(h);
        if (rc != BELLMAN_OK) { cleanup(); return rc; }
    }
    ST(cudaEventRecord(h->ev1, h->stream));
    ST(cudaStreamSynchronize(s_in));
    ST(cudaStreamSynchronize(h->stream));
    ST(cudaStreamSynchronize(s_out));
    float ms = 0.f;
    ST(cudaEventElapsedTime(&ms, h->ev0, h->ev1));
#undef ST
    h->last_ms = ms;
    h->cur_stage = to;
    h->J_set = true;
    cleanup();
    if (h->fused_halo && h->nranks > 1) {
        unsigned int timed_out = 0;
        CUDA_TRY(h, cudaMemcpy(&timed_out, h->d_flags + 48, sizeof(timed_out), cudaMemcpyDeviceToHost));
        if (timed_out) { h->err = "halo flag wait timed out: a neighbour rank stopped progressing"; return BELLMAN_ERR_NCCL; }
    }
    return rc;
}

extern "C" int bellman_last_run_stats(const bellman_handle *h, double *ms_total, int64_t *kernel_launches,
                                      double *ms_exchange) {
    if (!h) return BELLMAN_ERR_BAD_ARG;
    if (ms_total) *ms_total = h->last_ms;
    if (kernel_launches) *kernel_launches = h->last_launches;
    if (ms_exchange) *ms_exchange = h->last_ms_exchange;
    return BELLMAN_OK;
}

extern "C" const char *bellman_last_kernel(const bellman_handle *h) { return h ? h->last_kernel.c_str() : ""; }

// ---------------------------------------------------------------------------------------------
// rollout
// ---------------------------------------------------------------------------------------------
// One grow-only device buffer per handle for the consumers' inputs and outputs (a simulation loop calls the
// lookup / rollout entry points over and over: cudaMalloc + cudaFree per call cost more than the kernels).
// Returns nullptr (and sets h->err) when the allocation fails.
static unsigned char *consumer_scratch(bellman_handle *h, size_t bytes) {
    if (h->d_roll_bytes < bytes) {
        cudaFree(h->d_roll);
        h->d_roll = nullptr;
        h->d_roll_bytes = 0;
        const cudaError_t e = cudaMalloc(&h->d_roll, bytes);
        if (e != cudaSuccess) { h->err = cudaGetErrorString(e); h->d_roll = nullptr; return nullptr; }
        h->d_roll_bytes = bytes;
    }
    return h->d_roll;
}
static size_t align256(size_t b) { return (b + 255) / 256 * 256; }

extern "C" int bellman_rollout(bellman_handle *h, const double *A, const double *B, const double *u_values,
                               const double *x0, int32_t batch, int32_t mode, int32_t ssu_stage,
                               double *X_out, double *U_out) {
    if (!h || !A || !B || !u_values || !x0 || !X_out || !U_out || batch < 1) return BELLMAN_ERR_BAD_ARG;
    const HostProblem &hp = h->hp;
    if (hp.D != 2 || hp.P != 1 || h->nranks != 1) { h->err = "rollout needs D=2, P=1, one rank"; return BELLMAN_ERR_BAD_ARG; }
    if (!h->store_idx_all) { h->err = "rollout needs store_idx_all"; return BELLMAN_ERR_STATE; }
    if (h->cur_stage != 1) { h->err = "rollout needs a completed sweep (stage 1)"; return BELLMAN_ERR_NOT_RUN; }
    if (mode == 1 && (ssu_stage < 1 || ssu_stage > hp.N - 1)) return BELLMAN_ERR_BAD_ARG;
    CUDA_TRY(h, cudaSetDevice(h->device));
    const int N = hp.N;
    auto cleanup = [&]() {};
#define RT(expr) do { cudaError_t _e = (expr); if (_e != cudaSuccess) { h->err = cudaGetErrorString(_e); cleanup(); return BELLMAN_ERR_CUDA; } } while (0)
    const size_t b_u = align256(sizeof(double) * hp.C), b_x0 = align256(sizeof(double) * 2 * (size_t)batch),
                 b_X = align256(sizeof(double) * 2 * (size_t)N * batch), b_U = align256(sizeof(double) * (size_t)N * batch);
    unsigned char *buf = consumer_scratch(h, b_u + b_x0 + b_X + b_U);      // u_values | x0 | X | U
    if (!buf) return BELLMAN_ERR_CUDA;
    double *d_u = reinterpret_cast<double *>(buf), *d_x0 = reinterpret_cast<double *>(buf + b_u),
           *d_X = reinterpret_cast<double *>(buf + b_u + b_x0), *d_U = reinterpret_cast<double *>(buf + b_u + b_x0 + b_X);
    RT(cudaMemcpyAsync(d_u, u_values, sizeof(double) * hp.C, cudaMemcpyHostToDevice, h->stream));
    RT(cudaMemcpyAsync(d_x0, x0, sizeof(double) * 2 * (size_t)batch, cudaMemcpyHostToDevice, h->stream));
    RolloutParams rp;
    rp.grid0 = h->sp.dim[0].grid; rp.rinv0 = h->sp.dim[0].rinv;
    rp.grid1 = h->sp.dim[1].grid; rp.rinv1 = h->sp.dim[1].rinv;
    rp.inv_h0 = hp.inv_h[0][0]; rp.off0 = hp.off[0][0];
    rp.inv_h1 = hp.inv_h[1][0]; rp.off1 = hp.off[1][0];
    rp.lut0 = h->sp.dim[0].lut; rp.lut1 = h->sp.dim[1].lut;
    rp.lut_n0 = hp.lut_n[0]; rp.lut_n1 = hp.lut_n[1];
    rp.lut_invw0 = hp.lut_invw[0][0]; rp.lut_invw1 = hp.lut_invw[1][0];
    rp.mode0 = hp.mode[0]; rp.mode1 = hp.mode[1];
    rp.n0 = hp.n[0]; rp.n1 = hp.n[1]; rp.N = N; rp.C = hp.C; rp.batch = batch;
    rp.mode = mode; rp.ssu_stage = ssu_stage;
    rp.idx_all = h->d_idx; rp.idx_bytes = hp.idx_bytes; rp.u_values = d_u;
    for (int i = 0; i < 4; ++i) rp.A[i] = A[i];
    rp.B[0] = B[0]; rp.B[1] = B[1];
    rp.x0 = d_x0; rp.X_out = d_X; rp.U_out = d_U;
    RT(launch_rollout(rp, h->stream));
    RT(cudaMemcpyAsync(X_out, d_X, sizeof(double) * 2 * (size_t)N * batch, cudaMemcpyDeviceToHost, h->stream));
    RT(cudaMemcpyAsync(U_out, d_U, sizeof(double) * (size_t)N * batch, cudaMemcpyDeviceToHost, h->stream));
    RT(cudaStreamSynchronize(h->stream));
#undef RT
    cleanup();
    return BELLMAN_OK;
}

// ---------------------------------------------------------------------------------------------
// consumers of the sweep's output: nearest-policy lookup, simplified-plant axis rollout
// ---------------------------------------------------------------------------------------------
static int fill_policy_params(bellman_handle *h, int prob, PolicyParams &pp) {
    const HostProblem &hp = h->hp;
    if (prob < 0 || prob >= hp.P) { h->err = "problem index out of range"; return BELLMAN_ERR_BAD_ARG; }
    std::memset(&pp, 0, sizeof(pp));
    pp.D = hp.D;
    for (int d = 0; d < hp.D; ++d) {
        const DimParams &dp = h->sp.dim[d];
        pp.grid[d] = dp.grid + (size_t)prob * dp.n;
        pp.rinv[d] = dp.rinv + (size_t)prob * dp.n;
        pp.lut[d] = dp.lut + (size_t)prob * (dp.lut_n + 1);
        pp.inv_h[d] = hp.inv_h[d][prob];
        pp.off[d] = hp.off[d][prob];
        pp.lut_invw[d] = hp.lut_invw[d][prob];
        pp.mode[d] = hp.mode[(size_t)prob * hp.D + d];
        pp.n[d] = hp.n[d];
        pp.lut_n[d] = hp.lut_n[d];
        pp.own_lo[d] = h->own_lo[d];
        pp.own_n[d] = h->own_n[d];
    }
    pp.idx_bytes = hp.idx_bytes;
    return BELLMAN_OK;
}

extern "C" int bellman_policy_lookup(bellman_handle *h, int32_t prob, int32_t stage, const double *x,
                                     int32_t batch, int32_t *idx_out) {
    if (!h || !x || !idx_out || batch < 1) return BELLMAN_ERR_BAD_ARG;
    PolicyParams pp;
    int rc = fill_policy_params(h, prob, pp);
    if (rc != BELLMAN_OK) return rc;
    rc = stage_available(h, stage, h->store_idx_all, true);
    if (rc != BELLMAN_OK) { h->err = "stage not available"; return rc; }
    CUDA_TRY(h, cudaSetDevice(h->device));
    const HostProblem &hp = h->hp;
    auto cleanup = [&]() {};
#define PT(expr) do { cudaError_t _e = (expr); if (_e != cudaSuccess) { h->err = cudaGetErrorString(_e); cleanup(); return BELLMAN_ERR_CUDA; } } while (0)
    const size_t b_x = align256(sizeof(double) * (size_t)hp.D * batch);
    unsigned char *buf = consumer_scratch(h, b_x + align256(sizeof(int32_t) * (size_t)batch));      // x | idx
    if (!buf) return BELLMAN_ERR_CUDA;
    double *d_x = reinterpret_cast<double *>(buf);
    int32_t *d_o = reinterpret_cast<int32_t *>(buf + b_x);
    PT(cudaMemcpyAsync(d_x, x, sizeof(double) * (size_t)hp.D * batch, cudaMemcpyHostToDevice, h->stream));
    pp.batch = batch;
    pp.idx = h->idx_ptr(stage, prob);
    pp.x = d_x;
    pp.idx_out = d_o;
    PT(launch_policy_lookup(pp, h->stream));
    PT(cudaMemcpyAsync(idx_out, d_o, sizeof(int32_t) * (size_t)batch, cudaMemcpyDeviceToHost, h->stream));
    PT(cudaStreamSynchronize(h->stream));
    cleanup();
    return BELLMAN_OK;
}

extern "C" int bellman_rollout_axis(bellman_handle *h, int32_t prob, int32_t time_varying, int32_t stage,
                                    int32_t rate_dim, double h_step, const double *u_inc, const double *x0,
                                    int32_t batch, int32_t n_steps, double *X_out, int32_t *C_out) {
    if (!h || !u_inc || !x0 || !X_out || !C_out || batch < 1 || n_steps < 1) return BELLMAN_ERR_BAD_ARG;
    const HostProblem &hp = h->hp;
    if (hp.D != 2 || rate_dim < 0 || rate_dim > 1) { h->err = "axis rollout needs D = 2 and rate_dim in {0,1}"; return BELLMAN_ERR_BAD_ARG; }
    if (h->nranks != 1) { h->err = "rollouts cross slab boundaries: gather the policy into an unsharded handle (bellman_set_stage) first"; return BELLMAN_ERR_BAD_ARG; }
    PolicyParams pp;
    int rc = fill_policy_params(h, prob, pp);
    if (rc != BELLMAN_OK) return rc;
    if (time_varying) {
        if (!h->store_idx_all) { h->err = "time-varying rollout needs store_idx_all"; return BELLMAN_ERR_STATE; }
        if (h->cur_stage != 1) { h->err = "time-varying rollout needs a completed sweep"; return BELLMAN_ERR_NOT_RUN; }
        if (n_steps > hp.N - 1) { h->err = "n_steps exceeds the horizon"; return BELLMAN_ERR_BAD_ARG; }
    } else {
        rc = stage_available(h, stage, h->store_idx_all, true);
        if (rc != BELLMAN_OK) { h->err = "stage not available"; return rc; }
    }
    CUDA_TRY(h, cudaSetDevice(h->device));
    auto cleanup = [&]() {};
    const size_t b_x = align256(sizeof(double) * 2 * (size_t)batch), b_X = align256(sizeof(double) * 2 * (size_t)(n_steps + 1) * batch),
                 b_u = align256(sizeof(double) * (size_t)hp.C), b_c = align256(sizeof(int32_t) * (size_t)n_steps * batch);
    unsigned char *buf = consumer_scratch(h, b_x + b_X + b_u + b_c);       // x0 | X | u_inc | control indices
    if (!buf) return BELLMAN_ERR_CUDA;
    double *d_x = reinterpret_cast<double *>(buf), *d_X = reinterpret_cast<double *>(buf + b_x),
           *d_u = reinterpret_cast<double *>(buf + b_x + b_X);
    int32_t *d_c = reinterpret_cast<int32_t *>(buf + b_x + b_X + b_u);
    PT(cudaMemcpyAsync(d_x, x0, sizeof(double) * 2 * (size_t)batch, cudaMemcpyHostToDevice, h->stream));
    PT(cudaMemcpyAsync(d_u, u_inc, sizeof(double) * (size_t)hp.C, cudaMemcpyHostToDevice, h->stream));
    pp.batch = batch;
    pp.time_varying = time_varying; pp.stage = stage; pp.rate_dim = rate_dim; pp.n_steps = n_steps;
    pp.h_step = h_step;
    pp.idx = time_varying ? h->idx_ptr(1, prob) : h->idx_ptr(stage, prob);    // time varying: slot 0 = stage 1 (store_idx_all)
    pp.idx_stage_stride = (long long)h->slot_elems_idx();
    pp.u_inc = d_u; pp.x = d_x; pp.X_out = d_X; pp.idx_out = d_c;
    PT(launch_rollout_axis(pp, h->stream));
    PT(cudaMemcpyAsync(X_out, d_X, sizeof(double) * 2 * (size_t)(n_steps + 1) * batch, cudaMemcpyDeviceToHost, h->stream));
    PT(cudaMemcpyAsync(C_out, d_c, sizeof(int32_t) * (size_t)n_steps * batch, cudaMemcpyDeviceToHost, h->stream));
    PT(cudaStreamSynchronize(h->stream));
#undef PT
    cleanup();
    return BELLMAN_OK;
}

extern "C" int bellman_rollout_orbit(bellman_handle *h, int32_t stage, const bellman_orbit_opts *o,
                                     const double *u_values, const double *y0, int32_t batch, double *X_out,
                                     int32_t *C_out, int32_t *warn_out) {
    if (!h || !o || !u_values || !y0 || !X_out || !C_out || batch < 1) return BELLMAN_ERR_BAD_ARG;
    if (o->struct_size != (int32_t)sizeof(bellman_orbit_opts)) { h->err = "bellman_orbit_opts.struct_size mismatch"; return BELLMAN_ERR_BAD_ARG; }
    const HostProblem &hp = h->hp;
    if (hp.D != 2 || hp.P < 3) { h->err = "orbit rollout needs D = 2 and P >= 3 (the x, y, z axis problems)"; return BELLMAN_ERR_BAD_ARG; }
    if (h->nranks != 1) { h->err = "rollouts cross slab boundaries: gather the policy into an unsharded handle (bellman_set_stage) first"; return BELLMAN_ERR_BAD_ARG; }
    if (o->n_steps < 1 || o->stride_out < 1 || o->n_steps % o->stride_out) { h->err = "n_steps must be a positive multiple of stride_out"; return BELLMAN_ERR_BAD_ARG; }
    int rc = stage_available(h, stage, h->store_idx_all, true);
    if (rc != BELLMAN_OK) { h->err = "stage not available"; return rc; }
    OrbitParams op;
    std::memset(&op, 0, sizeof(op));
    for (int p = 0; p < 3; ++p) {
        rc = fill_policy_params(h, p, op.pol[p]);
        if (rc != BELLMAN_OK) return rc;
        op.pol[p].idx = h->idx_ptr(stage, p);
    }
    CUDA_TRY(h, cudaSetDevice(h->device));
    const int n_out = o->n_steps / o->stride_out;
    double *d_u = nullptr, *d_y = nullptr, *d_X = nullptr;
    int32_t *d_c = nullptr, *d_w = nullptr;
    auto cleanup = [&]() { cudaFree(d_u); cudaFree(d_y); cudaFree(d_X); cudaFree(d_c); cudaFree(d_w); };
#define OT(expr) do { cudaError_t _e = (expr); if (_e != cudaSuccess) { h->err = cudaGetErrorString(_e); cleanup(); return _e == cudaErrorMemoryAllocation ? BELLMAN_ERR_OOM : BELLMAN_ERR_CUDA; } } while (0)
    OT(cudaMalloc(&d_u, sizeof(double) * (size_t)hp.C));
    OT(cudaMalloc(&d_y, sizeof(double) * 6 * (size_t)batch));
    OT(cudaMalloc(&d_X, sizeof(double) * 6 * (size_t)(n_out + 1) * batch));
    OT(cudaMalloc(&d_c, sizeof(int32_t) * 3 * (size_t)n_out * batch));
    OT(cudaMalloc(&d_w, sizeof(int32_t) * (size_t)batch));
    OT(cudaMemcpyAsync(d_u, u_values, sizeof(double) * (size_t)hp.C, cudaMemcpyHostToDevice, h->stream));
    OT(cudaMemcpyAsync(d_y, y0, sizeof(double) * 6 * (size_t)batch, cudaMemcpyHostToDevice, h->stream));
    op.mu = o->mu;
    for (int k = 0; k < 3; ++k) { op.R0[k] = o->R0[k]; op.V0[k] = o->V0[k]; }
    op.h = o->h; op.tol = o->tol;
    op.n_steps = o->n_steps; op.batch = batch; op.stride_out = o->stride_out;
    op.max_rkf = o->max_rkf_steps > 0 ? o->max_rkf_steps : 100000;
    op.u_values = d_u; op.y0 = d_y; op.X_out = d_X; op.C_out = d_c; op.warn_out = d_w;
    OT(launch_rollout_orbit(op, h->stream));
    OT(cudaMemcpyAsync(X_out, d_X, sizeof(double) * 6 * (size_t)(n_out + 1) * batch, cudaMemcpyDeviceToHost, h->stream));
    OT(cudaMemcpyAsync(C_out, d_c, sizeof(int32_t) * 3 * (size_t)n_out * batch, cudaMemcpyDeviceToHost, h->stream));
    if (warn_out) OT(cudaMemcpyAsync(warn_out, d_w, sizeof(int32_t) * (size_t)batch, cudaMemcpyDeviceToHost, h->stream));
    OT(cudaStreamSynchronize(h->stream));
#undef OT
    cleanup();
    return BELLMAN_OK;
}

// ---------------------------------------------------------------------------------------------
// full-plant forward simulations (ode45): Solver_pos_att.get_optimal_path, Solver_attitude's testode45
// ---------------------------------------------------------------------------------------------
static int plant_common(bellman_handle *h, const bellman_plant_opts *o, PlantParams &pl) {
    if (o->struct_size != (int32_t)sizeof(bellman_plant_opts)) { h->err = "bellman_plant_opts.struct_size mismatch"; return BELLMAN_ERR_BAD_ARG; }
    if (o->n_steps < 1 || o->stride_out < 1 || o->n_steps % o->stride_out) { h->err = "n_steps must be a positive multiple of stride_out"; return BELLMAN_ERR_BAD_ARG; }
    if (!(o->h > 0) || !(o->rtol > 0) || !(o->atol > 0)) { h->err = "h, rtol, atol must be positive"; return BELLMAN_ERR_BAD_ARG; }
    if (!(o->inertia[0] > 0 && o->inertia[4] > 0 && o->inertia[8] > 0) || o->inertia[1] != o->inertia[3] ||
        o->inertia[2] != o->inertia[6] || o->inertia[5] != o->inertia[7]) { h->err = "inertia must be symmetric with a positive diagonal"; return BELLMAN_ERR_BAD_ARG; }
    pl.mu = o->mu;
    for (int k = 0; k < 3; ++k) { pl.R0[k] = o->R0[k]; pl.V0[k] = o->V0[k]; }
    for (int k = 0; k < 9; ++k) pl.Im[k] = o->inertia[k];
    pl.h = o->h; pl.rtol = o->rtol; pl.atol = o->atol; pl.mass = o->mass; pl.t_dist = o->t_dist;
    pl.n_steps = o->n_steps; pl.stride_out = o->stride_out;
    pl.max_ode = o->max_ode_steps > 0 ? o->max_ode_steps : 100000;
    return BELLMAN_OK;
}

extern "C" int bellman_rollout_pos_att(bellman_handle *hx, bellman_handle *hy, bellman_handle *hz, const int32_t stage[3],
                                       const bellman_plant_opts *o, const double *f_x, const double *f_y, const double *f_z,
                                       const double *y0, int32_t batch, double *X_out, double *F_out, double *FM_out,
                                       int32_t *warn_out) {
    if (!hx || !hy || !hz) return BELLMAN_ERR_BAD_ARG;
    bellman_handle *h = hx;
    if (!stage || !o || !f_x || !f_y || !f_z || !y0 || !X_out || !F_out || batch < 1) { h->err = "null argument"; return BELLMAN_ERR_BAD_ARG; }
    bellman_handle *hs[3] = {hx, hy, hz};
    const double *fh[3] = {f_x, f_y, f_z};
    PlantParams pl;
    std::memset(&pl, 0, sizeof(pl));
    int rc = plant_common(h, o, pl);
    if (rc != BELLMAN_OK) return rc;
    if (!(o->mass > 0) || !(o->mu > 0)) { h->err = "mass and mu must be positive"; return BELLMAN_ERR_BAD_ARG; }
    for (int p = 0; p < 3; ++p) {
        if (hs[p]->hp.D != 4) { h->err = "pos-att rollout needs three D = 4 channel handles"; return BELLMAN_ERR_BAD_ARG; }
        if (hs[p]->nranks != 1) { h->err = "rollouts cross slab boundaries: gather the policy into an unsharded handle (bellman_set_stage) first"; return BELLMAN_ERR_BAD_ARG; }
        if (hs[p]->device != h->device) { h->err = "the three channel handles must live on one device"; return BELLMAN_ERR_BAD_ARG; }
        rc = fill_policy_params(hs[p], 0, pl.pol[p]);
        if (rc != BELLMAN_OK) { h->err = hs[p]->err; return rc; }
        rc = stage_available(hs[p], stage[p], hs[p]->store_idx_all, true);
        if (rc != BELLMAN_OK) { h->err = "stage not available"; return rc; }
        pl.pol[p].idx = hs[p]->idx_ptr(stage[p], 0);
        pl.C[p] = hs[p]->hp.C;
    }
    CUDA_TRY(h, cudaSetDevice(h->device));
    CUDA_TRY(h, cudaStreamSynchronize(hy->stream));      // their policies were written on their own streams
    CUDA_TRY(h, cudaStreamSynchronize(hz->stream));
    const int n_out = o->n_steps / o->stride_out;
    const size_t b_f[3] = {align256(sizeof(double) * 4 * (size_t)pl.C[0]), align256(sizeof(double) * 4 * (size_t)pl.C[1]),
                           align256(sizeof(double) * 4 * (size_t)pl.C[2])};
    const size_t b_y = align256(sizeof(double) * 13 * (size_t)batch), b_X = align256(sizeof(double) * 13 * (size_t)(n_out + 1) * batch),
                 b_F = align256(sizeof(double) * 12 * (size_t)n_out * batch), b_M = align256(sizeof(double) * 6 * (size_t)n_out * batch),
                 b_w = align256(sizeof(int32_t) * (size_t)batch);
    unsigned char *buf = consumer_scratch(h, b_f[0] + b_f[1] + b_f[2] + b_y + b_X + b_F + b_M + b_w);
    if (!buf) return BELLMAN_ERR_CUDA;
    auto cleanup = [&]() {};
#define LT(expr) do { cudaError_t _e = (expr); if (_e != cudaSuccess) { h->err = cudaGetErrorString(_e); cleanup(); return BELLMAN_ERR_CUDA; } } while (0)
    unsigned char *cur = buf;
    for (int p = 0; p < 3; ++p) {
        LT(cudaMemcpyAsync(cur, fh[p], sizeof(double) * 4 * (size_t)pl.C[p], cudaMemcpyHostToDevice, h->stream));
        pl.fv[p] = reinterpret_cast<const double *>(cur);
        cur += b_f[p];
    }
    double *d_y = reinterpret_cast<double *>(cur); cur += b_y;
    double *d_X = reinterpret_cast<double *>(cur); cur += b_X;
    double *d_F = reinterpret_cast<double *>(cur); cur += b_F;
    double *d_M = reinterpret_cast<double *>(cur); cur += b_M;
    int32_t *d_w = reinterpret_cast<int32_t *>(cur);
    LT(cudaMemcpyAsync(d_y, y0, sizeof(double) * 13 * (size_t)batch, cudaMemcpyHostToDevice, h->stream));
    pl.kind = 0; pl.batch = batch;
    pl.y0 = d_y; pl.X_out = d_X; pl.F_out = d_F; pl.FM_out = FM_out ? d_M : nullptr; pl.warn_out = d_w;
    LT(launch_rollout_plant(pl, h->stream));
    LT(cudaMemcpyAsync(X_out, d_X, sizeof(double) * 13 * (size_t)(n_out + 1) * batch, cudaMemcpyDeviceToHost, h->stream));
    LT(cudaMemcpyAsync(F_out, d_F, sizeof(double) * 12 * (size_t)n_out * batch, cudaMemcpyDeviceToHost, h->stream));
    if (FM_out) LT(cudaMemcpyAsync(FM_out, d_M, sizeof(double) * 6 * (size_t)n_out * batch, cudaMemcpyDeviceToHost, h->stream));
    if (warn_out) LT(cudaMemcpyAsync(warn_out, d_w, sizeof(int32_t) * (size_t)batch, cudaMemcpyDeviceToHost, h->stream));
    LT(cudaStreamSynchronize(h->stream));
    cleanup();
    return BELLMAN_OK;
}

extern "C" int bellman_rollout_attitude(bellman_handle *h, int32_t stage, const bellman_plant_opts *o, const double *u_values,
                                        const double *y0, int32_t batch, double *X_out, int32_t *C_out, int32_t *warn_out) {
    if (!h) return BELLMAN_ERR_BAD_ARG;
    if (!o || !u_values || !y0 || !X_out || !C_out || batch < 1) { h->err = "null argument"; return BELLMAN_ERR_BAD_ARG; }
    const HostProblem &hp = h->hp;
    if (hp.D != 2 || hp.P < 3) { h->err = "attitude rollout needs D = 2 and P >= 3 (the three axis problems)"; return BELLMAN_ERR_BAD_ARG; }
    if (h->nranks != 1) { h->err = "rollouts cross slab boundaries: gather the policy into an unsharded handle (bellman_set_stage) first"; return BELLMAN_ERR_BAD_ARG; }
    PlantParams pl;
    std::memset(&pl, 0, sizeof(pl));
    int rc = plant_common(h, o, pl);
    if (rc != BELLMAN_OK) return rc;
    rc = stage_available(h, stage, h->store_idx_all, true);
    if (rc != BELLMAN_OK) { h->err = "stage not available"; return rc; }
    for (int p = 0; p < 3; ++p) {
        rc = fill_policy_params(h, p, pl.pol[p]);
        if (rc != BELLMAN_OK) return rc;
        pl.pol[p].idx = h->idx_ptr(stage, p);
        pl.C[p] = hp.C;
    }
    CUDA_TRY(h, cudaSetDevice(h->device));
    const int n_out = o->n_steps / o->stride_out;
    const size_t b_u = align256(sizeof(double) * (size_t)hp.C), b_y = align256(sizeof(double) * 7 * (size_t)batch),
                 b_X = align256(sizeof(double) * 7 * (size_t)(n_out + 1) * batch), b_c = align256(sizeof(int32_t) * 3 * (size_t)n_out * batch),
                 b_w = align256(sizeof(int32_t) * (size_t)batch);
    unsigned char *buf = consumer_scratch(h, b_u + b_y + b_X + b_c + b_w);
    if (!buf) return BELLMAN_ERR_CUDA;
    auto cleanup = [&]() {};
    double *d_u = reinterpret_cast<double *>(buf), *d_y = reinterpret_cast<double *>(buf + b_u),
           *d_X = reinterpret_cast<double *>(buf + b_u + b_y);
    int32_t *d_c = reinterpret_cast<int32_t *>(buf + b_u + b_y + b_X), *d_w = reinterpret_cast<int32_t *>(buf + b_u + b_y + b_X + b_c);
    LT(cudaMemcpyAsync(d_u, u_values, sizeof(double) * (size_t)hp.C, cudaMemcpyHostToDevice, h->stream));
    LT(cudaMemcpyAsync(d_y, y0, sizeof(double) * 7 * (size_t)batch, cudaMemcpyHostToDevice, h->stream));
    pl.kind = 1; pl.batch = batch;
    pl.fv[0] = d_u; pl.y0 = d_y; pl.X_out = d_X; pl.C_out = d_c; pl.warn_out = d_w;
    LT(launch_rollout_plant(pl, h->stream));
    LT(cudaMemcpyAsync(X_out, d_X, sizeof(double) * 7 * (size_t)(n_out + 1) * batch, cudaMemcpyDeviceToHost, h->stream));
    LT(cudaMemcpyAsync(C_out, d_c, sizeof(int32_t) * 3 * (size_t)n_out * batch, cudaMemcpyDeviceToHost, h->stream));
    if (warn_out) LT(cudaMemcpyAsync(warn_out, d_w, sizeof(int32_t) * (size_t)batch, cudaMemcpyDeviceToHost, h->stream));
    LT(cudaStreamSynchronize(h->stream));
#undef LT
    cleanup();
    return BELLMAN_OK;
}

// ---------------------------------------------------------------------------------------------
// point reads: J and argmin of listed states (spot checks at grid sizes whose arrays do not fit the host)
// ---------------------------------------------------------------------------------------------
struct PointArgs {
    const double *J;          // stage slot, problem applied
    const int32_t *idx;       // stage slot, problem applied (nullptr: skip); idx_bytes per element
    int idx_bytes;
    const long long *states;  // global linear indices (dimension 0 fastest)
    long long n;
    int D;
    int ng[MAXD], own_lo[MAXD], own_n[MAXD], ext_lo[MAXD];
    long long stride[MAXD];
    double *J_out;
    int32_t *idx_out;
    int *bad;                 // set when a state is outside the owned range
};

__global__ void k_get_points(const PointArgs a) {
    const long long m = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (m >= a.n) return;
    long long s = a.states[m], jo = 0, io = 0, os = 1;
    bool ok = s >= 0;
    for (int d = 0; d < a.D; ++d) {
        const int g = (int)(s % a.ng[d]);
        s /= a.ng[d];
        ok = ok && g >= a.own_lo[d] && g < a.own_lo[d] + a.own_n[d];
        jo += (long long)(g - a.ext_lo[d]) * a.stride[d];
        io += (long long)(g - a.own_lo[d]) * os;
        os *= a.own_n[d];
    }
    ok = ok && s == 0;
    if (!ok) { *a.bad = 1; return; }
    a.J_out[m] = a.J[jo];
    if (a.idx) a.idx_out[m] = idx_load(a.idx, a.idx_bytes, io);
}

extern "C" int bellman_get_points(bellman_handle *h, int32_t stage, int32_t prob, const int64_t *states, int64_t n,
                                  double *J_out, int32_t *idx_out) {
    if (!h || !states || !J_out || n < 1) return BELLMAN_ERR_BAD_ARG;
    const HostProblem &hp = h->hp;
    if (prob < 0 || prob >= hp.P) { h->err = "problem index out of range"; return BELLMAN_ERR_BAD_ARG; }
    int rc = stage_available(h, stage, h->store_J_all, false);
    if (rc == BELLMAN_OK && idx_out) rc = stage_available(h, stage, h->store_idx_all, true);
    if (rc != BELLMAN_OK) { h->err = "stage not available"; return rc; }
    CUDA_TRY(h, cudaSetDevice(h->device));
    long long *d_s = nullptr;
    double *d_J = nullptr;
    int32_t *d_i = nullptr;
    int *d_bad = nullptr;
    auto cleanup = [&]() { cudaFree(d_s); cudaFree(d_J); cudaFree(d_i); cudaFree(d_bad); };
#define GT(expr) do { cudaError_t _e = (expr); if (_e != cudaSuccess) { h->err = cudaGetErrorString(_e); cleanup(); return BELLMAN_ERR_CUDA; } } while (0)
    GT(cudaMalloc(&d_s, sizeof(long long) * (size_t)n));
    GT(cudaMalloc(&d_J, sizeof(double) * (size_t)n));
    GT(cudaMalloc(&d_i, sizeof(int32_t) * (size_t)n));
    GT(cudaMalloc(&d_bad, sizeof(int)));
    GT(cudaMemsetAsync(d_bad, 0, sizeof(int), h->stream));
    GT(cudaMemcpyAsync(d_s, states, sizeof(long long) * (size_t)n, cudaMemcpyHostToDevice, h->stream));
    PointArgs a;
    std::memset(&a, 0, sizeof(a));
    a.J = h->J_ptr(stage) + (size_t)prob * h->S_ext;
    a.idx = idx_out ? h->idx_ptr(stage, prob) : nullptr;
    a.idx_bytes = hp.idx_bytes;
    a.states = d_s; a.n = n; a.D = hp.D;
    for (int d = 0; d < hp.D; ++d) {
        a.ng[d] = hp.n[d]; a.own_lo[d] = h->own_lo[d]; a.own_n[d] = h->own_n[d]; a.ext_lo[d] = h->ext_lo[d];
        a.stride[d] = h->stride[d];
    }
    a.J_out = d_J; a.idx_out = d_i; a.bad = d_bad;
    k_get_points<<<(unsigned)((n + 255) / 256), 256, 0, h->stream>>>(a);
    GT(cudaGetLastError());
    int bad = 0;
    GT(cudaMemcpyAsync(J_out, d_J, sizeof(double) * (size_t)n, cudaMemcpyDeviceToHost, h->stream));
    if (idx_out) GT(cudaMemcpyAsync(idx_out, d_i, sizeof(int32_t) * (size_t)n, cudaMemcpyDeviceToHost, h->stream));
    GT(cudaMemcpyAsync(&bad, d_bad, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    GT(cudaStreamSynchronize(h->stream));
#undef GT
    cleanup();
    if (bad) { h->err = "bellman_get_points: a state lies outside this rank's owned range"; return BELLMAN_ERR_BAD_ARG; }
    return BELLMAN_OK;
}

// ---------------------------------------------------------------------------------------------
// Single-process multi-GPU: ONE host thread drives every slab (SURVEY 8b "one host thread drives all
// GPUs" — what a MATLAB interpreter calling run(obj) through the MEX gateway needs).  The handles are
// created with part_dim >= 0, nranks = n, rank = position in the array, device = the GPU of that slab
// (several slabs may share a GPU).  No NCCL: the slabs' J buffers are in one address space, so the
// stage kernels store halo values straight into the neighbours' buffers (peer access between devices)
// and stages are ordered by the same neighbour flags as the multi-process fused mode.
// ---------------------------------------------------------------------------------------------
extern "C" int bellman_group_init(bellman_handle **hs, int32_t n) {
    if (!hs || n < 1) return BELLMAN_ERR_BAD_ARG;
    for (int r = 0; r < n; ++r) {
        bellman_handle *h = hs[r];
        if (!h) return BELLMAN_ERR_BAD_ARG;
        if (h->nranks != n || h->rank != r || (n > 1 && h->part_dim != hs[0]->part_dim) || h->hp.S() != hs[0]->hp.S() ||
            h->store_J_all != hs[0]->store_J_all) {
            h->err = "bellman_group_init: handle r must be rank r of n, all on the same problem and partition dimension";
            return BELLMAN_ERR_BAD_ARG;
        }
        if (n - 1 > MAX_PEERS) { h->err = "too many slabs for the fused halo"; return BELLMAN_ERR_BAD_ARG; }
    }
    for (int r = 0; r < n; ++r) {
        bellman_handle *h = hs[r];
        CUDA_TRY(h, cudaSetDevice(h->device));
        h->peer_J.assign(n, nullptr);
        for (int q = 0; q < n; ++q) {
            if (q == r) continue;
            if (hs[q]->device != h->device) {
                int can = 0;
                CUDA_TRY(h, cudaDeviceCanAccessPeer(&can, h->device, hs[q]->device));
                if (!can) { h->err = "no peer access between the devices of the group"; return BELLMAN_ERR_CUDA; }
                cudaError_t e = cudaDeviceEnablePeerAccess(hs[q]->device, 0);
                if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) { h->err = cudaGetErrorString(e); return BELLMAN_ERR_CUDA; }
                cudaGetLastError();
            }
            h->peer_J[q] = hs[q]->d_J;
        }
        h->fused_halo = n > 1;
        h->group_mode = true;
    }
    return BELLMAN_OK;
}

extern "C" int bellman_group_run(bellman_handle **hs, int32_t n, int32_t n_stages, const bellman_run_opts *opts) {
    if (!hs || n < 1 || n_stages < 0) return BELLMAN_ERR_BAD_ARG;
    bellman_run_opts o;
    std::memset(&o, 0, sizeof(o));
    if (opts) {
        if (opts->struct_size != (int32_t)sizeof(bellman_run_opts)) { hs[0]->err = "bellman_run_opts.struct_size mismatch"; return BELLMAN_ERR_BAD_ARG; }
        o = *opts;
    }
    std::vector<int> kernel(n), lanes(n, 1);
    for (int r = 0; r < n; ++r) {
        bellman_handle *h = hs[r];
        if (!h->group_mode) { h->err = "bellman_group_run needs bellman_group_init first"; return BELLMAN_ERR_STATE; }
        if (h->cur_stage != hs[0]->cur_stage) { h->err = "the slabs of a group must be at the same stage"; return BELLMAN_ERR_STATE; }
        if (h->cur_stage - n_stages < 1) { h->err = "run would pass stage 1"; return BELLMAN_ERR_STATE; }
        CUDA_TRY(h, cudaSetDevice(h->device));
        kernel[r] = pick_kernel(h, o.kernel, lanes[r]);
        h->last_kernel = kernel[r] == BELLMAN_KERNEL_WINDOW ? window_variant(h) : kernel[r] == BELLMAN_KERNEL_TILE
                             ? (stream_valid(h) ? "stream" : "tile") : kernel[r] == BELLMAN_KERNEL_SPLITC ? "splitc" : "direct";
        h->last_launches = 0;
        h->last_ms_exchange = 0.0;
        CUDA_TRY(h, cudaEventRecord(h->ev0, h->stream));
    }
    double fsum_prev = hs[0]->check_log.empty() ? 0.0 : hs[0]->check_log[hs[0]->check_log.size() - 2];
    bool stop = false;
    for (int s = 0; s < n_stages && !stop; ++s) {
        // every slab's stage, signal and wait go to its own stream; the host thread never blocks here
        for (int r = 0; r < n; ++r) {
            bellman_handle *h = hs[r];
            CUDA_TRY(h, cudaSetDevice(h->device));
            int rc = launch_one_stage(h, kernel[r], lanes[r]);
            if (rc != BELLMAN_OK) return rc;
            h->cur_stage -= 1;
            if (n > 1) {
                h->halo_seq += 1;
                rc = halo_signal(h);
                if (rc == BELLMAN_OK) rc = halo_wait(h);
                if (rc != BELLMAN_OK) return rc;
            }
        }
        const int cur = hs[0]->cur_stage;
        if (o.check_period > 0 && cur % o.check_period == 0) {
            double tot[2] = {0.0, 0.0};
            for (int r = 0; r < n; ++r) {
                bellman_handle *h = hs[r];
                CUDA_TRY(h, cudaSetDevice(h->device));
                StageParams sp = h->sp;
                sp.J_out = h->J_ptr(cur);
                sp.idx_out = h->idx_ptr(cur);
                cudaError_t e = launch_check_sums(sp, h->d_partials, h->n_partials, h->d_sums, h->stream);
                if (e != cudaSuccess) { h->err = cudaGetErrorString(e); return BELLMAN_ERR_CUDA; }
            }
            for (int r = 0; r < n; ++r) {     // slab sums added in rank order (deterministic)
                bellman_handle *h = hs[r];
                double sums[2];
                CUDA_TRY(h, cudaSetDevice(h->device));
                CUDA_TRY(h, cudaMemcpyAsync(sums, h->d_sums, sizeof(sums), cudaMemcpyDeviceToHost, h->stream));
                CUDA_TRY(h, cudaStreamSynchronize(h->stream));
                tot[0] += sums[0];
                tot[1] += sums[1];
            }
            for (int r = 0; r < n; ++r) {
                hs[r]->check_log.push_back((double)cur);
                hs[r]->check_log.push_back(tot[0]);
                hs[r]->check_log.push_back(tot[1]);
            }
            if (std::fabs(tot[0] - fsum_prev) < o.check_tol) stop = true;
            fsum_prev = tot[0];
        }
    }
    double ms_max = 0.0;
    for (int r = 0; r < n; ++r) {
        bellman_handle *h = hs[r];
        CUDA_TRY(h, cudaSetDevice(h->device));
        CUDA_TRY(h, cudaEventRecord(h->ev1, h->stream));
    }
    for (int r = 0; r < n; ++r) {
        bellman_handle *h = hs[r];
        CUDA_TRY(h, cudaSetDevice(h->device));
        CUDA_TRY(h, cudaStreamSynchronize(h->stream));
        float ms = 0.f;
        CUDA_TRY(h, cudaEventElapsedTime(&ms, h->ev0, h->ev1));
        ms_max = std::max(ms_max, (double)ms);
        if (n > 1) {
            unsigned int timed_out = 0;
            CUDA_TRY(h, cudaMemcpy(&timed_out, h->d_flags + 48, sizeof(timed_out), cudaMemcpyDeviceToHost));
            if (timed_out) { h->err = "halo flag wait timed out: a slab of the group stopped progressing"; return BELLMAN_ERR_NCCL; }
        }
    }
    for (int r = 0; r < n; ++r) hs[r]->last_ms = ms_max;
    return BELLMAN_OK;
}
