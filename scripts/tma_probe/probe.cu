// TMA probe: which variant of a fp64 box load works on this B200?  usage: probe <variant>
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
typedef CUresult (*PFN)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
__device__ __forceinline__ uint32_t s32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
template <int RANK>
__global__ void k(const __grid_constant__ CUtensorMap tm, const CUtensorMap *gtm, int use_global, int x, int y, double *out, int n) {
    extern __shared__ __align__(128) unsigned char raw[];
    __shared__ __align__(8) uint64_t bar;
    double *dst = reinterpret_cast<double *>(((uintptr_t)raw + 127) & ~(uintptr_t)127);
    const CUtensorMap *m = use_global ? gtm : &tm;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(&bar)), "r"(n * 8) : "memory");
        if (RANK == 3)
            asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(s32(dst)), "l"(m), "r"(s32(&bar)), "r"(x), "r"(y), "r"(0) : "memory");
        else
            asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(s32(dst)), "l"(m), "r"(s32(&bar)), "r"(x), "r"(y) : "memory");
    }
    asm volatile("{\n.reg .pred p;\nW: mbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n@p bra D;\nbra W;\nD:\n}\n" ::"r"(s32(&bar)) : "memory");
    for (int i = threadIdx.x; i < n; i += blockDim.x) out[i] = dst[i];
}
int main(int argc, char **argv) {
    int variant = argc > 1 ? atoi(argv[1]) : 0;
    int rank = (variant & 1) ? 2 : 3, use_global = (variant & 2) ? 1 : 0;
    int b0 = argc > 2 ? atoi(argv[2]) : 48, b1 = argc > 3 ? atoi(argv[3]) : 97; int dt = argc > 4 ? atoi(argv[4]) : 0; int cx = argc > 5 ? atoi(argv[5]) : 5, cy = argc > 6 ? atoi(argv[6]) : 7;
    const int n0 = 128, n1 = 128;
    double *d, *o; cudaMalloc(&d, n0 * n1 * 8); cudaMalloc(&o, b0 * b1 * 8);
    double *h = (double *)malloc(n0 * n1 * 8); for (int i = 0; i < n0 * n1; ++i) h[i] = i; cudaMemcpy(d, h, n0 * n1 * 8, cudaMemcpyHostToDevice);
    void *p; cudaDriverEntryPointQueryResult q; cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q); PFN enc = (PFN)p;
    CUtensorMap tm; cuuint64_t gd[3] = {n0, n1, 1}, gs[2] = {n0 * 8, (cuuint64_t)n0 * n1 * 8}; cuuint32_t bx[3] = {(cuuint32_t)b0, (cuuint32_t)b1, 1}, es[3] = {1, 1, 1};
    CUtensorMapDataType dtype = CU_TENSOR_MAP_DATA_TYPE_FLOAT64;
    if (dt == 1) { dtype = CU_TENSOR_MAP_DATA_TYPE_FLOAT32; gd[0] *= 2; bx[0] *= 2; }
    if (dt == 2) dtype = CU_TENSOR_MAP_DATA_TYPE_UINT64;
    if (dt == 3) dtype = CU_TENSOR_MAP_DATA_TYPE_INT64;
    if (dt == 4) { dtype = CU_TENSOR_MAP_DATA_TYPE_UINT32; gd[0] *= 2; bx[0] *= 2; }
    CUresult r = enc(&tm, dtype, rank, d, gd, gs, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("variant %d rank %d global %d box %dx%d dt %d encode=%d\n", variant, rank, use_global, b0, b1, dt, (int)r);
    CUtensorMap *gtm; cudaMalloc(&gtm, sizeof(tm)); cudaMemcpy(gtm, &tm, sizeof(tm), cudaMemcpyHostToDevice);
    size_t smem = (size_t)b0 * b1 * 8 + 128;
    cudaFuncSetAttribute(k<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaFuncSetAttribute(k<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (rank == 3) k<3><<<1, 128, smem>>>(tm, gtm, use_global, cx, cy, o, b0 * b1); else k<2><<<1, 128, smem>>>(tm, gtm, use_global, cx, cy, o, b0 * b1);
    cudaError_t e = cudaDeviceSynchronize();
    printf("  sync: %s\n", cudaGetErrorString(e));
    if (e == cudaSuccess) { double *ho = (double *)malloc(b0 * b1 * 8); cudaMemcpy(ho, o, b0 * b1 * 8, cudaMemcpyDeviceToHost); printf("  out[0]=%g (expect %d) out[b0]=%g (expect %d)\n", ho[0], 7 * n0 + 5, ho[b0], 8 * n0 + 5); }
    return 0;
}
