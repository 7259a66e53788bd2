// bellman_internal.h — shared between the host-only planner, the kernels and the C-ABI layer.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

#include "../../include/bellman.h"

namespace bellman {

constexpr int MAXD = BELLMAN_MAX_DIM;

// ---------------------------------------------------------------------------------------------
// Host-side, GPU-free image of a validated descriptor (all tables copied).
// ---------------------------------------------------------------------------------------------
struct HostProblem {
    int D = 0, C = 0, P = 0, N = 0;
    int n[MAXD] = {0, 0, 0, 0};
    int src_a[MAXD] = {0, 0, 0, 0}, src_b[MAXD] = {-1, -1, -1, -1};
    int q_order[MAXD] = {0, 1, 2, 3};
    bool has_b[MAXD] = {false, false, false, false}, has_c[MAXD] = {false, false, false, false};
    std::vector<double> grid[MAXD], rinv[MAXD], Ta[MAXD], Tb[MAXD], Tc[MAXD], q[MAXD], r;
    std::vector<double> inv_h[MAXD], off[MAXD];   // [P] per dim
    // SEARCH dimensions: bucket table that starts the exact bin search a cell or two below the
    // answer: lut[d][p*(lut_n[d]+1) + b] = cell of the lower edge of bucket b (4 buckets per cell),
    // bucket(x) = floor((x - s[0]) * lut_invw[d][p])
    std::vector<int32_t> lut[MAXD];
    std::vector<double> lut_invw[MAXD];
    int lut_n[MAXD] = {0, 0, 0, 0};
    std::vector<int32_t> mode;                    // [P][D]
    std::vector<int32_t> part_cuts;               // explicit slab boundaries (empty: equal slabs)
    int idx_bytes = 4;                            // device storage of the argmin
    int64_t S() const { int64_t s = 1; for (int d = 0; d < D; ++d) s *= n[d]; return s; }
};

// returns "" on success, else an error message
std::string load_problem(const bellman_desc *d, HostProblem &hp);
// the locate rule of include/bellman.h evaluated on the host (used by the reach analysis)
int host_locate(const HostProblem &hp, int p, int d, double x);
// exact reach analysis: range of cells [lo, hi] (inclusive, hi = cell+1 node) of dimension `dim`
// touched by states whose index along `dim` lies in [own_lo, own_hi)
void reach_range(const HostProblem &hp, int dim, int own_lo, int own_hi, int &ext_lo, int &ext_hi);
std::string plan_slabs(const HostProblem &hp, int part_dim, int nranks, bellman_slab *out);
// exact stencil bounds of dimension d (bellman_tile.cu): over every state and control,
// cell(x'_d) - i_d in [lo, hi]; false when x'_d does not depend on the state's own index
bool stencil_reach(const HostProblem &hp, int d, int &lo, int &hi);

// ---------------------------------------------------------------------------------------------
// Kernel parameter block (passed by value; device pointers address problem 0, rows are P-strided)
// ---------------------------------------------------------------------------------------------
struct DimParams {
    const double *grid;   // [P][n]
    const double *rinv;   // [P][n]   (last entry unused)
    const double *Ta;     // [P][n_a]
    const double *Tb;     // [P][n_b] or nullptr
    const double *Tc;     // [P][C]   or nullptr
    const double *q;      // [P][n]
    const double *loc;    // [P][2] = {inv_h, off} (UNIFORM) or {lut_invw, s[0]} (SEARCH)
    const int32_t *mode;  // [P]
    const int32_t *lut;   // [P][lut_n + 1] (SEARCH dimensions)
    int lut_n;
    int n, n_a, n_b, src_a, src_b;
    int own_n;            // number of owned indices along this dim (== n unless partitioned)
    int own_lo;           // first owned global index
    int ext_lo;           // global index of local slot 0 of the J arrays
    long long stride;     // element stride of this dim in the (extended) J arrays
};

// A neighbour rank that reads part of my owned slab: in fused (peer-memory) mode the stage kernel
// stores those states straight into the neighbour's J buffer over NVLink, so no separate halo
// exchange (pack / send / recv) exists — only a per-stage barrier.
constexpr int MAX_PEERS = 7;
struct PeerHalo {
    double *J;              // neighbour's J_out slot (problem 0), peer-mapped (cudaIpc)
    long long S_ext;        // neighbour's per-problem stride
    long long stride[MAXD]; // neighbour's element strides
    int lo, hi;             // global index range along part_dim the neighbour reads from me
    int ext_lo;             // neighbour's first stored index along part_dim
    int pad_;
};

struct StageParams {
    DimParams dim[MAXD];
    const double *r;          // [P][C]
    const double *J_next;     // [P][S_ext]
    double *J_out;            // [P][S_ext]
    int32_t *idx_out;         // [P][S_own]
    long long S_ext, S_own;
    int D, C, P;
    int q_order[MAXD];
    int n_peers, part_dim;
    int idx_bytes;            // bytes per stored argmin: 4 (int32), 2 (uint16) or 1 (uint8); idx_out is typed int32_t* regardless
    PeerHalo peer[MAX_PEERS];
};

// argmin storage of 1 / 2 / 4 bytes behind one element-indexed interface (o = element offset)
#ifdef __CUDACC__
#define BELLMAN_HD __host__ __device__ __forceinline__
#else
#define BELLMAN_HD inline
#endif
BELLMAN_HD void idx_store(void *base, int bytes, long long o, int v) {
    if (bytes == 4) static_cast<int32_t *>(base)[o] = v;
    else if (bytes == 1) static_cast<unsigned char *>(base)[o] = (unsigned char)v;
    else static_cast<unsigned short *>(base)[o] = (unsigned short)v;
}
BELLMAN_HD int idx_load(const void *base, int bytes, long long o) {
    if (bytes == 4) return static_cast<const int32_t *>(base)[o];
    if (bytes == 1) return static_cast<const unsigned char *>(base)[o];
    return static_cast<const unsigned short *>(base)[o];
}

#ifdef __CUDACC__
// store `v` (the new J of global state gi) into every neighbour whose halo covers it
template <int D>
__device__ __forceinline__ void peer_store(const StageParams &sp, int prob, const int (&gi)[D], double v) {
    for (int e = 0; e < sp.n_peers; ++e) {
        const PeerHalo &ph = sp.peer[e];
        const int ip = gi[sp.part_dim];
        if (ip >= ph.lo && ip < ph.hi) {
            long long o = (long long)prob * ph.S_ext;
#pragma unroll
            for (int d = 0; d < D; ++d) o += (long long)(gi[d] - (d == sp.part_dim ? ph.ext_lo : 0)) * ph.stride[d];
            ph.J[o] = v;
        }
    }
}
#endif

// cudaFuncAttributeMaxDynamicSharedMemorySize is a property of the FUNCTION (per device), not of a
// handle: two handles that plan different window sizes for the same kernel (uneven slabs of a group,
// two sweeps alive in one process) must not lower each other's limit.  Keeps the maximum ever asked
// for per (device, function) and only raises the attribute.  Returns false when CUDA refuses.
bool raise_smem_limit(const void *fn, size_t bytes);

// window (TMA-staged) kernel configuration for D = 2
struct WindowConfig {
    int tile0 = 0, tile1 = 0;     // states per CTA tile
    int cchunk = 0;               // controls per staged window
    int win0 = 0, win1 = 0;       // window box (rows, cols) in cells
    bool valid = false;
};

}  // namespace bellman
