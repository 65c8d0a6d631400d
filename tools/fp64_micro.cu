// FP64 pipe micro-benchmarks: DFMA issue rate per SMSP under different operand patterns.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_micro tools/fp64_micro.cu && ./fp64_micro
#include <cstdio>
#include <cuda_runtime.h>

constexpr int ITERS = 2048;

template <int PAT>
__global__ void __launch_bounds__(256) k(double* out, const double* in, int iters) {
    double a[8], b[8], c[8];
    const double x = in[threadIdx.x & 7], y = in[8 + (threadIdx.x & 7)];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        a[k] = in[k] + threadIdx.x;
        b[k] = in[16 + k] + 1e-7 * threadIdx.x;
        c[k] = in[24 + k] + 1e-9 * threadIdx.x;
    }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            if (PAT == 0) a[k] = fma(a[k], 1.0000001, 1e-9);        // 1 reg + 2 constants
            if (PAT == 1) a[k] = fma(a[k], x, y);                   // 3 regs, two shared by all chains
            if (PAT == 2) a[k] = fma(a[k], b[k], c[k]);             // 3 distinct regs
            if (PAT == 3) a[k] = fma(x, b[k], a[k]);                // shared multiplicand, distinct multiplier, accumulate
            if (PAT == 4) a[k] = fma(b[k], c[k], a[k]);             // distinct, accumulate
            if (PAT == 5) a[k] = fma(a[k], a[k], c[k]);             // 2 distinct
            if (PAT == 6) a[k] = a[k] * b[k];                       // DMUL 2 regs
            if (PAT == 7) a[k] = fma(a[k], b[k], y);                // 2 distinct + shared
        }
    }
    double s = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) s += a[k];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int PAT>
void run(const char* name, double* out, double* in, int sms) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    for (int warps_per_smsp : {1, 2, 4, 8}) {
        const int blocks = sms * (warps_per_smsp * 4 * 32 / 256 > 0 ? warps_per_smsp * 4 * 32 / 256 : 1);
        const int threads = warps_per_smsp * 4 * 32 >= 256 ? 256 : warps_per_smsp * 4 * 32;
        k<PAT><<<blocks, threads>>>(out, in, 16);
        cudaEventRecord(e0);
        k<PAT><<<blocks, threads>>>(out, in, ITERS);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        const double warp_instr_per_smsp = (double)ITERS * 8 * warps_per_smsp;
        int clk_khz = 0;
        cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
        const double cycles = ms * 1e-3 * 1.965e9;
        printf("%-44s warps/SMSP %d : %.2f cycles per warp-DFMA per SMSP  (%.2f T-DFMA/s chip)\n", name, warps_per_smsp,
               cycles / warp_instr_per_smsp, (double)blocks * threads * ITERS * 8 / (ms * 1e-3) / 1e12);
    }
}

int main() {
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    printf("%s, %d SMs\n", p.name, p.multiProcessorCount);
    double *out, *in;
    cudaMalloc(&out, 1 << 24);
    cudaMalloc(&in, 4096);
    double h[64];
    for (int i = 0; i < 64; ++i) h[i] = 1.0 + 1e-6 * i;
    cudaMemcpy(in, h, sizeof h, cudaMemcpyHostToDevice);
    const int sms = p.multiProcessorCount;
    run<0>("P0 fma(a, const, const)", out, in, sms);
    run<1>("P1 fma(a, x, y)  x,y shared regs", out, in, sms);
    run<2>("P2 fma(a_k, b_k, c_k) 3 distinct regs", out, in, sms);
    run<3>("P3 fma(x, b_k, a_k) shared x, accumulate", out, in, sms);
    run<4>("P4 fma(b_k, c_k, a_k) distinct, accumulate", out, in, sms);
    run<5>("P5 fma(a_k, a_k, c_k)", out, in, sms);
    run<6>("P6 a_k * b_k (DMUL)", out, in, sms);
    run<7>("P7 fma(a_k, b_k, y)", out, in, sms);
    return 0;
}
