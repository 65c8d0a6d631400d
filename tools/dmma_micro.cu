// Is the FP64 tensor path (DMMA) a second FP64 pipe on B200, or the same units as DFMA?
// Times DFMA-only, DMMA-only (three shapes) and an interleaved mix with the same per-warp work.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/bin/dmma_micro tools/dmma_micro.cu
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma884(double& d0, double& d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}
__device__ __forceinline__ void dmma1684(double (&d)[4], const double (&a)[2], double b) {
    asm volatile("mma.sync.aligned.m16n8k4.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};"
                 : "+d"(d[0]), "+d"(d[1]), "+d"(d[2]), "+d"(d[3]) : "d"(a[0]), "d"(a[1]), "d"(b));
}
__device__ __forceinline__ void dmma1688(double (&d)[4], const double (&a)[4], const double (&b)[2]) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+d"(d[0]), "+d"(d[1]), "+d"(d[2]), "+d"(d[3])
                 : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(b[0]), "d"(b[1]));
}

// MODE 0: NF DFMA chains.  MODE 1: NM m8n8k4 accumulators.  MODE 2: m16n8k4.  MODE 3: m16n8k8.
// MODE 4: NF DFMA + NM m8n8k4 interleaved.  MODE 5: NF DFMA + NM m16n8k8 interleaved.
template <int MODE, int NF, int NM>
__global__ void __launch_bounds__(128) k(double* out, const double* in, int iters) {
    double f[NF > 0 ? NF : 1];
    double acc[NM > 0 ? NM : 1][4];
    double a4[4], b2[2], a2[2];
    const double x = in[threadIdx.x & 7], y = in[8 + (threadIdx.x & 7)];
    for (int i = 0; i < 4; ++i) a4[i] = in[i] * 1e-3 + 1e-9 * threadIdx.x;
    for (int i = 0; i < 2; ++i) { b2[i] = in[4 + i] * 1e-3; a2[i] = a4[i]; }
#pragma unroll
    for (int i = 0; i < NF; ++i) f[i] = in[i & 7] + threadIdx.x;
#pragma unroll
    for (int i = 0; i < NM; ++i)
        for (int j = 0; j < 4; ++j) acc[i][j] = in[16 + j] + i;
    for (int it = 0; it < iters; ++it) {
        if (MODE == 0 || MODE >= 4) {
#pragma unroll
            for (int i = 0; i < NF; ++i) f[i] = fma(f[i], x, y);
        }
        if (MODE == 1 || MODE == 4) {
#pragma unroll
            for (int i = 0; i < NM; ++i) dmma884(acc[i][0], acc[i][1], a4[0], b2[0]);
        }
        if (MODE == 2) {
#pragma unroll
            for (int i = 0; i < NM; ++i) dmma1684(acc[i], a2, b2[0]);
        }
        if (MODE == 3 || MODE == 5) {
#pragma unroll
            for (int i = 0; i < NM; ++i) dmma1688(acc[i], a4, b2);
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < NF; ++i) s += f[i];
#pragma unroll
    for (int i = 0; i < NM; ++i) s += acc[i][0] + acc[i][1] + acc[i][2] + acc[i][3];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE, int NF, int NM>
float run(const char* name, double* out, double* in, int sms, int warps_per_smsp, double fma_per_mma) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    const int iters = 4096;
    const int blocks = sms * warps_per_smsp, threads = 128;
    k<MODE, NF, NM><<<blocks, threads>>>(out, in, 16);
    float best = 1e30f;
    for (int r = 0; r < 3; ++r) {
        cudaEventRecord(e0);
        k<MODE, NF, NM><<<blocks, threads>>>(out, in, iters);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    const double warps = (double)blocks * threads / 32;
    const double dfma = (MODE == 0 || MODE >= 4) ? warps * 32.0 * NF * iters : 0.0;
    const double mma_fma = (MODE != 0) ? warps * NM * fma_per_mma * iters : 0.0;
    const double cyc = best * 1e-3 * 1.965e9;
    printf("%-34s w/SMSP %d: %8.3f ms  vector %.2f TFMA/s  tensor %.2f TFMA/s  total %.2f TFMA/s  (%.1f FMA/clk/SM at 1965 MHz)\n",
           name, warps_per_smsp, best, dfma / (best * 1e-3) / 1e12, mma_fma / (best * 1e-3) / 1e12,
           (dfma + mma_fma) / (best * 1e-3) / 1e12, (dfma + mma_fma) / cyc / sms);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) printf("  CUDA error: %s\n", cudaGetErrorString(e));
    return best;
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
    for (int w : {2, 4, 8}) {
        run<0, 16, 0>("DFMA x16", out, in, sms, w, 0);
        run<1, 0, 8>("DMMA m8n8k4 x8", out, in, sms, w, 256);
        run<2, 0, 8>("DMMA m16n8k4 x8", out, in, sms, w, 512);
        run<3, 0, 8>("DMMA m16n8k8 x8", out, in, sms, w, 1024);
        run<4, 16, 8>("mix DFMA x16 + m8n8k4 x8", out, in, sms, w, 256);
        run<4, 16, 2>("mix DFMA x16 + m8n8k4 x2", out, in, sms, w, 256);
        run<5, 16, 4>("mix DFMA x16 + m16n8k8 x4", out, in, sms, w, 1024);
        run<5, 16, 1>("mix DFMA x16 + m16n8k8 x1", out, in, sms, w, 1024);
    }
    return 0;
}
