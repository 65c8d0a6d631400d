#include <cstdio>
#include <cuda_runtime.h>
struct __align__(32) Corr { double xa, ya, xb, yb; };
#ifndef NPTS
#define NPTS 512
#endif
#define VFMA(d, a, b, c) asm volatile("fma.rn.f64 %0, %1, %2, %3;" : "=d"(d) : "d"(a), "d"(b), "d"(c))
// ORDER 0: compiler order (fma()); 1: pinned hypothesis-innermost; 2: pinned, two points interleaved by stage
template <int HPT, int G, int ORDER, int MINB>
__global__ void __launch_bounds__(128, MINB) k(const double* __restrict__ E, const Corr* __restrict__ pts, int reps,
                                         unsigned* __restrict__ out) {
    __shared__ __align__(128) Corr tile[1024];
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) tile[i] = pts[i % NPTS];
    __syncthreads();
    double e[HPT][9], kap[HPT];
#pragma unroll
    for (int j = 0; j < HPT; ++j) {
#pragma unroll
        for (int q = 0; q < 9; ++q) e[j][q] = E[((blockIdx.x * HPT + j) * 128 + threadIdx.x) * 9 + q];
        kap[j] = 1e-30 * e[j][0];
    }
    unsigned acc = 0;
    for (int r = 0; r < reps; ++r) {
        for (int p = 0; p < NPTS; p += G) {
            unsigned pm = 0;
            if (ORDER <= 1) {
#pragma unroll
            for (int g = 0; g < G; ++g) {
                const Corr c = tile[(p + g) & 1023];
                double t0[HPT], t1[HPT], t2[HPT], d[HPT];
                if (ORDER == 1) {
#pragma unroll
                    for (int j = 0; j < HPT; ++j) VFMA(t0[j], c.yb, e[j][3], e[j][6]);
#pragma unroll
                    for (int j = 0; j < HPT; ++j) VFMA(t1[j], c.yb, e[j][4], e[j][7]);
#pragma unroll
                    for (int j = 0; j < HPT; ++j) VFMA(t2[j], c.yb, e[j][5], e[j][8]);
#pragma unroll
                    for (int j = 0; j < HPT; ++j) VFMA(t0[j], c.xb, e[j][0], t0[j]);
#pragma unroll
                    for (int j = 0; j < HPT; ++j) VFMA(t1[j], c.xb, e[j][1], t1[j]);
#pragma unroll
                    for (int j = 0; j < HPT; ++j) VFMA(t2[j], c.xb, e[j][2], t2[j]);
#pragma unroll
                    for (int j = 0; j < HPT; ++j) VFMA(t2[j], c.ya, t1[j], t2[j]);
#pragma unroll
                    for (int j = 0; j < HPT; ++j) VFMA(t2[j], c.xa, t0[j], t2[j]);
#pragma unroll
                    for (int j = 0; j < HPT; ++j) VFMA(t1[j], t1[j], t1[j], kap[j]);
#pragma unroll
                    for (int j = 0; j < HPT; ++j) VFMA(t1[j], t0[j], t0[j], t1[j]);
#pragma unroll
                    for (int j = 0; j < HPT; ++j) { double nt = -t1[j]; VFMA(d[j], t2[j], t2[j], nt); }
                } else {
#pragma unroll
                    for (int j = 0; j < HPT; ++j) { t0[j] = fma(c.yb, e[j][3], e[j][6]); t1[j] = fma(c.yb, e[j][4], e[j][7]); t2[j] = fma(c.yb, e[j][5], e[j][8]); }
#pragma unroll
                    for (int j = 0; j < HPT; ++j) { t0[j] = fma(c.xb, e[j][0], t0[j]); t1[j] = fma(c.xb, e[j][1], t1[j]); t2[j] = fma(c.xb, e[j][2], t2[j]); }
#pragma unroll
                    for (int j = 0; j < HPT; ++j) t2[j] = fma(c.ya, t1[j], t2[j]);
#pragma unroll
                    for (int j = 0; j < HPT; ++j) t2[j] = fma(c.xa, t0[j], t2[j]);
#pragma unroll
                    for (int j = 0; j < HPT; ++j) t1[j] = fma(t1[j], t1[j], kap[j]);
#pragma unroll
                    for (int j = 0; j < HPT; ++j) t1[j] = fma(t0[j], t0[j], t1[j]);
#pragma unroll
                    for (int j = 0; j < HPT; ++j) d[j] = fma(t2[j], t2[j], -t1[j]);
                }
#pragma unroll
                for (int j = 0; j < HPT; ++j) pm = __funnelshift_l((unsigned)__double2hiint(d[j]), pm, 1);
            }
            } else {
            // ORDER 2: PP=2 points in lockstep, stage by stage, all pinned
#pragma unroll
            for (int g = 0; g < G; g += 2) {
                Corr c[2];
                c[0] = tile[(p + g) & 1023]; c[1] = tile[(p + g + 1) & 1023];
                double t0[2][HPT], t1[2][HPT], t2[2][HPT], d[2][HPT];
#pragma unroll
                for (int u = 0; u < 2; ++u) {
#pragma unroll
                    for (int j = 0; j < HPT; ++j) VFMA(t0[u][j], c[u].yb, e[j][3], e[j][6]);
#pragma unroll
                    for (int j = 0; j < HPT; ++j) VFMA(t1[u][j], c[u].yb, e[j][4], e[j][7]);
#pragma unroll
                    for (int j = 0; j < HPT; ++j) VFMA(t2[u][j], c[u].yb, e[j][5], e[j][8]);
                }
#pragma unroll
                for (int u = 0; u < 2; ++u) {
#pragma unroll
                    for (int j = 0; j < HPT; ++j) VFMA(t0[u][j], c[u].xb, e[j][0], t0[u][j]);
#pragma unroll
                    for (int j = 0; j < HPT; ++j) VFMA(t1[u][j], c[u].xb, e[j][1], t1[u][j]);
#pragma unroll
                    for (int j = 0; j < HPT; ++j) VFMA(t2[u][j], c[u].xb, e[j][2], t2[u][j]);
                }
#pragma unroll
                for (int u = 0; u < 2; ++u)
#pragma unroll
                    for (int j = 0; j < HPT; ++j) VFMA(t2[u][j], c[u].ya, t1[u][j], t2[u][j]);
#pragma unroll
                for (int u = 0; u < 2; ++u)
#pragma unroll
                    for (int j = 0; j < HPT; ++j) VFMA(t1[u][j], t1[u][j], t1[u][j], kap[j]);
#pragma unroll
                for (int u = 0; u < 2; ++u)
#pragma unroll
                    for (int j = 0; j < HPT; ++j) VFMA(t2[u][j], c[u].xa, t0[u][j], t2[u][j]);
#pragma unroll
                for (int u = 0; u < 2; ++u)
#pragma unroll
                    for (int j = 0; j < HPT; ++j) VFMA(t1[u][j], t0[u][j], t0[u][j], t1[u][j]);
#pragma unroll
                for (int u = 0; u < 2; ++u)
#pragma unroll
                    for (int j = 0; j < HPT; ++j) { double nt = -t1[u][j]; VFMA(d[u][j], t2[u][j], t2[u][j], nt); }
#pragma unroll
                for (int u = 0; u < 2; ++u)
#pragma unroll
                    for (int j = 0; j < HPT; ++j) pm = __funnelshift_l((unsigned)__double2hiint(d[u][j]), pm, 1);
            }
            }
            if (__any_sync(0xffffffffu, pm != 0u)) acc += pm;
        }
    }
    out[blockIdx.x * 128 + threadIdx.x] = acc;
}
template <int HPT, int G, int ORDER, int MINB>
void run(const char* name, const double* E, const Corr* pts, unsigned* out, int sms) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    int occ = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k<HPT, G, ORDER, MINB>, 128, 0);
    cudaFuncAttributes fa; cudaFuncGetAttributes(&fa, k<HPT, G, ORDER, MINB>);
    const int blocks = sms * occ, reps = 64 * 512 / NPTS;
    k<HPT, G, ORDER, MINB><<<blocks, 128>>>(E, pts, 2, out);
    float best = 1e30f;
    for (int r = 0; r < 3; ++r) {
        cudaEventRecord(e0);
        k<HPT, G, ORDER, MINB><<<blocks, 128>>>(E, pts, reps, out);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    const double evals = (double)blocks * 128 * HPT * NPTS * reps;
    const double w = evals * 11 / 32 / (sms * 4);
    printf("%-34s regs %3d occ %d: %7.3f ms  %.3e evals/s  %.2f cycles/warp-DFMA/SMSP\n", name, fa.numRegs, occ, best,
           evals / (best * 1e-3), best * 1e-3 * 1.965e9 / w);
}
int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    const int sms = p.multiProcessorCount;
    printf("%s, %d SMs\n", p.name, sms);
    const size_t nE = (size_t)sms * 8 * 8 * 128 * 9;
    double* E; Corr* pts; unsigned* out;
    cudaMalloc(&E, nE * 8); cudaMalloc(&pts, NPTS * sizeof(Corr)); cudaMalloc(&out, (size_t)sms * 8 * 128 * 4);
    double* hE = new double[nE];
    for (size_t i = 0; i < nE; ++i) hE[i] = 0.1 + 1e-3 * (double)(i % 977);
    cudaMemcpy(E, hE, nE * 8, cudaMemcpyHostToDevice);
    Corr h[NPTS];
    for (int i = 0; i < NPTS; ++i) h[i] = {0.01 * i, 0.3 - 0.002 * i, 0.5 + 0.001 * i, -0.2 + 0.003 * i};
    cudaMemcpy(pts, h, sizeof h, cudaMemcpyHostToDevice);
    run<2, 16, 0, 5>("HPT2 G16 compiler order minb5", E, pts, out, sms);
    run<2, 16, 1, 5>("HPT2 G16 pinned hyp-inner minb5", E, pts, out, sms);
    run<2, 16, 2, 5>("HPT2 G16 pinned PP2 minb5", E, pts, out, sms);
    run<2, 16, 1, 4>("HPT2 G16 pinned hyp-inner minb4", E, pts, out, sms);
    run<2, 16, 2, 4>("HPT2 G16 pinned PP2 minb4", E, pts, out, sms);
    run<4, 8, 0, 3>("HPT4 G8 compiler order minb3", E, pts, out, sms);
    run<4, 8, 1, 3>("HPT4 G8 pinned hyp-inner minb3", E, pts, out, sms);
    run<4, 8, 2, 3>("HPT4 G8 pinned PP2 minb3", E, pts, out, sms);
    run<4, 8, 1, 2>("HPT4 G8 pinned hyp-inner minb2", E, pts, out, sms);
    return 0;
}
