// probe3: the CUDA programming guide's TMA example shape (libcu++ wrappers), int32 2-D tile
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda/barrier>
#include <cstdio>
using barrier = cuda::barrier<cuda::thread_scope_block>;
namespace cde = cuda::device::experimental;
constexpr int GW = 1024, GH = 1024, SW = 32, SH = 16;
#ifdef V_DTYPE64
typedef double elem_t;
#define DT CU_TENSOR_MAP_DATA_TYPE_FLOAT64
#else
typedef int elem_t;
#define DT CU_TENSOR_MAP_DATA_TYPE_INT32
#endif
#ifdef V_L2
#define L2P CU_TENSOR_MAP_L2_PROMOTION_L2_128B
#else
#define L2P L2P
#endif
__device__ __forceinline__ unsigned s32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__global__ void kernel(const __grid_constant__ CUtensorMap tensor_map, int x, int y, elem_t *out) {
    __shared__ alignas(128) elem_t smem_buffer[SH][SW];
#pragma nv_diag_suppress static_var_with_dynamic_init
    __shared__ barrier bar;
    if (threadIdx.x == 0) { init(&bar, blockDim.x); cde::fence_proxy_async_shared_cta(); }
    __syncthreads();
    barrier::arrival_token token;
    if (threadIdx.x == 0) {
#ifdef V_PTX
        unsigned long long *nb = (unsigned long long *)cuda::device::barrier_native_handle(bar);
#ifdef V_ORDER
        token = cuda::device::barrier_arrive_tx(bar, 1, sizeof(smem_buffer));
#endif
        asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(s32(&smem_buffer)), "l"(&tensor_map), "r"(s32(nb)), "r"(x), "r"(y) : "memory");
#ifndef V_ORDER
        token = cuda::device::barrier_arrive_tx(bar, 1, sizeof(smem_buffer));
#endif
#else
        cde::cp_async_bulk_tensor_2d_global_to_shared(&smem_buffer, &tensor_map, x, y, bar);
        token = cuda::device::barrier_arrive_tx(bar, 1, sizeof(smem_buffer));
#endif
    } else {
        token = bar.arrive();
    }
    bar.wait(std::move(token));
    if (threadIdx.x == 0) { out[0] = smem_buffer[0][0]; out[1] = smem_buffer[1][0]; }
}
typedef CUresult (*PFN)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
int main() {
    elem_t *d, *o; cudaMalloc(&d, GW * GH * sizeof(elem_t)); cudaMalloc(&o, 2 * sizeof(elem_t));
    elem_t *h = (elem_t *)malloc(GW * GH * sizeof(elem_t)); for (int i = 0; i < GW * GH; ++i) h[i] = i; cudaMemcpy(d, h, GW * GH * sizeof(elem_t), cudaMemcpyHostToDevice);
    void *p; cudaDriverEntryPointQueryResult q; cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
    CUtensorMap tm{}; cuuint64_t size[2] = {GW, GH}, stride[1] = {GW * sizeof(elem_t)}; cuuint32_t box[2] = {SW, SH}, es[2] = {1, 1};
    CUresult r = ((PFN)p)(&tm, DT, 2, d, size, stride, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, L2P, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("encode=%d q=%d\n", (int)r, (int)q);
    kernel<<<1, 128>>>(tm, XX, 3, o);
    cudaError_t e = cudaDeviceSynchronize();
    elem_t ho[2] = {0, 0}; if (e == cudaSuccess) cudaMemcpy(ho, o, 2 * sizeof(elem_t), cudaMemcpyDeviceToHost);
    printf("guide sample: %s out=%d,%d expect %d,%d\n", cudaGetErrorString(e), (int)ho[0], (int)ho[1], 3 * GW + 64, 4 * GW + 64);
    return 0;
}
