// Mock of the window kernel's inner loop (one warp-update = 32 lanes x 1 control), with parts
// removable by template flags, to see which resource bounds it on B200.
//   bit0: shared-memory gathers   bit1: F2I conversions   bit2: select (argmin)   bit3: lerps
//   bit4: weights via bit-trick DADD (else skip)           bit5: address IMADs
#include <cuda_runtime.h>
#include <cstdio>
#define ITER 512
template <int F>
__global__ void __launch_bounds__(256, 2) k(double *out, const double *in, int *iout, long long *cyc, int W0) {
    extern __shared__ double sm[];
    for (int i = threadIdx.x; i < 6144; i += blockDim.x) sm[i] = in[i & 4095];
    __syncthreads();
    double base0[8], base1[8], gs[8], best[8]; int arg[8];
    for (int u = 0; u < 8; ++u) { base0[u] = 1.0 + (threadIdx.x & 31) * 0.9974 + u * 0.05; base1[u] = 2.0 + u * 1.159 - (threadIdx.x & 31) * 0.1078; gs[u] = in[u]; best[u] = 1e300; arg[u] = 0; }
    long long t0 = clock64();
#pragma unroll 1
    for (int c = 0; c < ITER; ++c) {
        const double bu0 = in[(c & 7)] * 1e-3, bu1 = in[8 + (c & 7)] * 1e-2, rc = in[16 + (c & 7)];
#pragma unroll
        for (int mb = 0; mb < 8; mb += 4) {
            int off[4]; double t0_[4], t1_[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int m = mb + u;
                const double g0 = base0[m] + bu0, g1 = base1[m] + bu1;
                int c0, c1;
                if (F & 2) { c0 = __double2int_rd(g0); c1 = __double2int_rd(g1); }
                else { c0 = __double2hiint(g0) & 31; c1 = __double2loint(g1) & 63; }
                if (F & 16) { t0_[u] = g0 - (__hiloint2double(0x43300000, c0) - 4503599627370496.0); t1_[u] = g1 - (__hiloint2double(0x43300000, c1) - 4503599627370496.0); }
                else { t0_[u] = g0; t1_[u] = g1; }
                if (F & 32) off[u] = (c1 & 63) * W0 + (c0 & 31) + (threadIdx.x & 31) * 0; else off[u] = (threadIdx.x & 31) + u;
            }
            double v00[4], v10[4], v01[4], v11[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                if (F & 1) { const double *p = sm + off[u]; v00[u] = p[0]; v10[u] = p[1]; v01[u] = p[W0]; v11[u] = p[W0 + 1]; }
                else { v00[u] = t0_[u]; v10[u] = t1_[u]; v01[u] = t0_[u]; v11[u] = t1_[u]; }
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int m = mb + u;
                double v;
                if (F & 8) { const double a = fma(t0_[u], v10[u] - v00[u], v00[u]); const double b = fma(t0_[u], v11[u] - v01[u], v01[u]); v = fma(t1_[u], b - a, a); }
                else v = v00[u] + v11[u];
                const double tot = (gs[m] + rc) + v;
                if (F & 4) { if (tot < best[m]) { best[m] = tot; arg[m] = c; } }
                else best[m] = best[m] + tot;
            }
        }
    }
    long long t1 = clock64();
    double s = 0; int qi = 0;
    for (int u = 0; u < 8; ++u) { s += best[u]; qi += arg[u]; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s; iout[blockIdx.x * blockDim.x + threadIdx.x] = qi;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
template <int F> void run(const char *name) {
    double *out, *in; int *iout; long long *cyc;
    cudaMalloc(&out, 296 * 256 * 8); cudaMalloc(&in, 8192 * 8); cudaMalloc(&iout, 296 * 256 * 4); cudaMalloc(&cyc, 296 * 8);
    double h[8192]; for (int i = 0; i < 8192; ++i) h[i] = 1.0 + i * 1e-3; cudaMemcpy(in, h, sizeof(h), cudaMemcpyHostToDevice);
    cudaFuncSetAttribute(k<F>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    for (int rep = 0; rep < 2; ++rep) { k<F><<<296, 256, 100 * 1024>>>(out, in, iout, cyc, 48); cudaDeviceSynchronize(); }
    long long hc[296]; cudaMemcpy(hc, cyc, sizeof(hc), cudaMemcpyDeviceToHost);
    double avg = 0; for (int i = 0; i < 296; ++i) avg += hc[i]; avg /= 296;
    // per SM: 2 CTAs x 8 warps x 8 states x ITER warp-updates
    printf("%-58s flags=%2d  cycles per warp-update per SM = %6.2f   (%s)\n", name, F, avg / (2.0 * 8 * 8 * ITER), cudaGetErrorString(cudaGetLastError()));
    cudaFree(out); cudaFree(in); cudaFree(iout); cudaFree(cyc);
}
int main() {
    run<63>("full: LDS + F2I + select + lerp + bit-trick + addr");
    run<62>("no LDS");
    run<61>("no F2I");
    run<59>("no select");
    run<55>("no lerp");
    run<47>("no bit-trick weights");
    run<31>("no address IMAD");
    run<60>("no LDS, no F2I");
    run<12>("only lerp + select (+x adds, cost)");
    run<8>("only lerp (+x adds, cost)");
    run<0>("only x adds + cost");
    return 0;
}
