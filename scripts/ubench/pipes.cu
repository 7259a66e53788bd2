// Pipe-throughput microbenchmark for the instruction mix of the Bellman stage kernel (B200).
// For each test: 148 CTAs x W warps, every thread runs ITER iterations of 8 independent chains.
// Reports warp-instructions per SM per cycle (and lanes/clk/SM).
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#define ITER 2048
template <int T>
__global__ void k(double *out, const double *in, int *iout, long long *cyc) {
    __shared__ double sm[4096];
    for (int i = threadIdx.x; i < 4096; i += blockDim.x) sm[i] = in[i];
    __syncthreads();
    double a[8]; int q[8];
    for (int u = 0; u < 8; ++u) { a[u] = in[threadIdx.x + u]; q[u] = threadIdx.x * 8 + u; }
    const double c1 = in[9], c2 = in[10];
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITER; ++it) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            if (T == 0) a[u] = a[u] + c1;                                   // DADD
            if (T == 1) a[u] = fma(a[u], c1, c2);                           // DFMA
            if (T == 2) { q[u] = __double2int_rd(a[u]) + q[u]; a[u] = __hiloint2double(q[u], q[u]); }  // F2I.F64 (+int add)
            if (T == 3) { a[u] = (double)q[u]; q[u] = __double2hiint(a[u]) ^ q[u]; }                   // I2F.F64 (+xor)
            if (T == 4) { a[u] = sm[(q[u]) & 4095]; q[u] += __double2loint(a[u]); }                  // LDS.64 (+add)
            if (T == 5) { a[u] = a[u] + c1; q[u] = q[u] * 3 + 1; }          // DADD + IMAD
            if (T == 6) { a[u] = a[u] + c1; q[u] = q[u] * 3 + 1; q[u] ^= (q[u] >> 3); }  // DADD + 3 int
            if (T == 7) { a[u] = fma(a[u], c1, c2); a[u] = a[u] + c1; q[u] = __double2int_rd(a[u]); a[u] = a[u] - (double)q[u]; }  // fma,add,F2I,I2F,add
            if (T == 8) { q[u] = q[u] * 3 + 1; }                            // IMAD only
            if (T == 9) { a[u] = (a[u] < c1) ? c2 : a[u]; }                 // DSETP + 2 FSEL
            if (T == 10) { a[u] = sm[(threadIdx.x & 31) + ((q[u] >> 5) & 127) * 32]; q[u] += __double2loint(a[u]); }   // conflict-free LDS.64 + IADD
            if (T == 11) { a[u] = sm[(threadIdx.x & 31) + ((q[u] >> 5) & 127) * 32]; q[u] += __double2int_rd(a[u]); }   // LDS.64 + F2I
            if (T == 12) { double x = sm[(threadIdx.x & 31) + ((q[u] >> 5) & 127) * 32]; double y = sm[(threadIdx.x & 31) + ((q[u] >> 7) & 127) * 32];
                           int c = __double2int_rd(x + a[u]); a[u] = y - (double)c; q[u] += c; }                        // 2 LDS + F2I + I2F + 2 DADD
            if (T == 13) { a[u] = a[u] + c1; a[u] = a[u] + c2; a[u] = a[u] + c1; q[u] += __double2int_rd(a[u]); }      // 3 DADD + F2I
            if (T == 14) { a[u] = a[u] + c1; a[u] = a[u] + c2; a[u] = a[u] + c1; a[u] += sm[(threadIdx.x & 31) + ((q[u] >> 5) & 127) * 32]; q[u] += 7; }  // 4 DADD + LDS
            if (T == 16) { a[u] = a[u] + c1; a[u] = a[u] + c2; a[u] = a[u] + c1; a[u] = a[u] + c2; a[u] = a[u] + c1; a[u] = a[u] + c2; a[u] = a[u] + c1;
                           const int o = (threadIdx.x & 31) + ((q[u] >> 5) & 63) * 32; a[u] += sm[o]; a[u] += sm[o + 2048]; q[u] += 7; }  // 9 DADD + 2 LDS.64
            if (T == 17) { a[u] = a[u] + c1; a[u] = a[u] + c2; a[u] = a[u] + c1; a[u] = a[u] + c2; a[u] = a[u] + c1; a[u] = a[u] + c2; a[u] = a[u] + c1;
                           const double2 v = reinterpret_cast<const double2 *>(sm)[(threadIdx.x & 31) + ((q[u] >> 5) & 63) * 32]; a[u] += v.x; a[u] += v.y; q[u] += 7; }  // 9 DADD + 1 LDS.128
            if (T == 18) { a[u] = a[u] + c1; a[u] = a[u] + c2; a[u] = a[u] + c1; a[u] = a[u] + c2; a[u] = a[u] + c1; a[u] = a[u] + c2; a[u] = a[u] + c1; a[u] = a[u] + c2; a[u] = a[u] + c1; q[u] += 7; }  // 9 DADD
            if (T == 19) { const double2 v = reinterpret_cast<const double2 *>(sm)[(threadIdx.x & 31) + ((q[u] >> 5) & 63) * 32]; q[u] += __double2loint(v.x) + __double2loint(v.y); }  // LDS.128 alone
            if (T == 15) { int c = __double2int_rd(a[u]); a[u] = a[u] - (double)c; q[u] += c; }                         // F2I + I2F + DADD
        }
    }
    long long t1 = clock64();
    double s = 0; int qi = 0;
    for (int u = 0; u < 8; ++u) { s += a[u]; qi += q[u]; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s; iout[blockIdx.x * blockDim.x + threadIdx.x] = qi;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
template <int T> void run(const char *name, int warps, double instr_per_chain_iter) {
    double *out, *in; int *iout; long long *cyc;
    cudaMalloc(&out, 148 * 1024 * 8); cudaMalloc(&in, 8192 * 8); cudaMalloc(&iout, 148 * 1024 * 4); cudaMalloc(&cyc, 148 * 8);
    double h[8192]; for (int i = 0; i < 8192; ++i) h[i] = 1.0 + i * 1e-3; cudaMemcpy(in, h, sizeof(h), cudaMemcpyHostToDevice);
    k<T><<<148, warps * 32>>>(out, in, iout, cyc); cudaDeviceSynchronize();
    k<T><<<148, warps * 32>>>(out, in, iout, cyc); cudaDeviceSynchronize();
    long long hc[148]; cudaMemcpy(hc, cyc, sizeof(hc), cudaMemcpyDeviceToHost);
    double avg = 0; for (int i = 0; i < 148; ++i) avg += hc[i]; avg /= 148;
    double chains = (double)warps * 8 * ITER;   // warp-level chain-iterations per SM
    printf("%-34s warps=%2d  cycles=%9.0f  cyc per warp-chain-iter per SM = %6.3f  (x%.0f instr)\n", name, warps, avg, avg / chains, instr_per_chain_iter);
    cudaFree(out); cudaFree(in); cudaFree(iout); cudaFree(cyc);
}
int main() {
    for (int w : {8, 16}) {
        run<10>("LDS.64 conflict-free + IADD", w, 2); run<11>("LDS.64 + F2I", w, 2); run<12>("2 LDS + F2I + I2F + 2 DADD", w, 6);
        run<13>("3 DADD + F2I", w, 4); run<14>("4 DADD + LDS", w, 5); run<15>("F2I + I2F + DADD", w, 3);
        run<18>("9 DADD", w, 9); run<16>("9 DADD + 2 LDS.64", w, 11); run<17>("9 DADD + 1 LDS.128", w, 10); run<19>("LDS.128 conflict-free", w, 1);
        printf("\n");
    }
    return 0;
}
