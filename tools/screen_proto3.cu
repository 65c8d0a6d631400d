// Micro-benchmark: the one-sided screen's main loop with the warp-uniform record operands delivered
//   V0  from shared memory into vector registers (what k_score does: LDS.128, every lane the same value)
//   V1  from __constant__ memory (upper bound: constant-bank / uniform-register operands cost no vector-register reads)
//   V2  from shared memory, then moved to uniform registers with REDUX (__reduce_or_sync of the two halves)
//   V3  from shared memory, then __shfl_sync(.., 0)
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o screen_proto3 screen_proto3.cu
#include <cstdio>
#include <cuda_runtime.h>
struct __align__(32) Corr { double xa, ya, xb, yb; };
#define NPTS 512
__constant__ Corr ctile[NPTS];

__device__ __forceinline__ double uni_redux(double x) {
    const unsigned lo = __reduce_or_sync(0xffffffffu, (unsigned)__double2loint(x));
    const unsigned hi = __reduce_or_sync(0xffffffffu, (unsigned)__double2hiint(x));
    return __hiloint2double((int)hi, (int)lo);
}
__device__ __forceinline__ double uni_shfl(double x) { return __shfl_sync(0xffffffffu, x, 0); }

template <int HPT, int G, int V, int MINB>
__global__ void __launch_bounds__(128, MINB) k(const double* __restrict__ E, const Corr* __restrict__ pts, int reps,
                                               unsigned* __restrict__ out) {
    __shared__ __align__(128) Corr tile[NPTS];
    for (int i = threadIdx.x; i < NPTS; i += blockDim.x) tile[i] = pts[i];
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
#pragma unroll
            for (int g = 0; g < G; ++g) {
                Corr c;
                if (V == 1) c = ctile[p + g];
                else c = tile[p + g];
                if (V == 2) { c.xa = uni_redux(c.xa); c.ya = uni_redux(c.ya); c.xb = uni_redux(c.xb); c.yb = uni_redux(c.yb); }
                if (V == 3) { c.xa = uni_shfl(c.xa); c.ya = uni_shfl(c.ya); c.xb = uni_shfl(c.xb); c.yb = uni_shfl(c.yb); }
                double t0[HPT], t1[HPT], t2[HPT], d[HPT];
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
#pragma unroll
                for (int j = 0; j < HPT; ++j) pm = __funnelshift_l((unsigned)__double2hiint(d[j]), pm, 1);
            }
            if (__any_sync(0xffffffffu, pm != 0u)) acc += pm;
        }
    }
    out[blockIdx.x * 128 + threadIdx.x] = acc;
}
template <int HPT, int G, int V, int MINB>
void run(const char* name, const double* E, const Corr* pts, unsigned* out, int sms) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    int occ = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k<HPT, G, V, MINB>, 128, 0);
    cudaFuncAttributes fa; cudaFuncGetAttributes(&fa, k<HPT, G, V, MINB>);
    const int blocks = sms * occ, reps = 64;
    k<HPT, G, V, MINB><<<blocks, 128>>>(E, pts, 2, out);
    float best = 1e30f;
    for (int r = 0; r < 3; ++r) {
        cudaEventRecord(e0);
        k<HPT, G, V, MINB><<<blocks, 128>>>(E, pts, reps, out);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
    }
    const double evals = (double)blocks * 128 * HPT * NPTS * reps;
    const double dfma = evals * 11 / 32;  // warp-DFMAs
    const double cyc = best * 1e-3 * 1.965e9 * sms * 4 / dfma;
    printf("%-34s regs %3d occ %d: %7.3f ms  %.3e evals/s  %.2f cycles/warp-DFMA/SMSP  (%s)\n", name, fa.numRegs, occ, best,
           evals / (best * 1e-3), cyc, cudaGetErrorString(cudaGetLastError()));
}
int main() {
    cudaDeviceProp pr; cudaGetDeviceProperties(&pr, 0);
    const int sms = pr.multiProcessorCount;
    printf("%s, %d SMs\n", pr.name, sms);
    const size_t nE = (size_t)sms * 8 * 4 * 128 * 9;
    double* hE = new double[nE];
    for (size_t i = 0; i < nE; ++i) hE[i] = 0.1 + 1e-3 * (double)(i % 977);
    Corr hp[NPTS];
    for (int i = 0; i < NPTS; ++i) hp[i] = Corr{0.01 * i, 0.02 * i, 0.5 - 0.01 * i, 0.3 + 0.005 * i};
    double* E; Corr* pts; unsigned* out;
    cudaMalloc(&E, nE * 8); cudaMalloc(&pts, sizeof hp); cudaMalloc(&out, (size_t)sms * 8 * 128 * 4);
    cudaMemcpy(E, hE, nE * 8, cudaMemcpyHostToDevice); cudaMemcpy(pts, hp, sizeof hp, cudaMemcpyHostToDevice);
    cudaMemcpyToSymbol(ctile, hp, sizeof hp);
    run<2, 16, 0, 5>("HPT2 G16 V0 shared->vector", E, pts, out, sms);
    run<2, 16, 1, 5>("HPT2 G16 V1 constant bank", E, pts, out, sms);
    run<2, 16, 2, 5>("HPT2 G16 V2 shared->REDUX->uniform", E, pts, out, sms);
    run<2, 16, 3, 5>("HPT2 G16 V3 shared->SHFL lane 0", E, pts, out, sms);
    run<4, 8, 0, 3>("HPT4 G8 V0 shared->vector", E, pts, out, sms);
    run<4, 8, 1, 3>("HPT4 G8 V1 constant bank", E, pts, out, sms);
    run<4, 8, 2, 3>("HPT4 G8 V2 shared->REDUX->uniform", E, pts, out, sms);
    run<1, 32, 1, 8>("HPT1 G32 V1 constant bank", E, pts, out, sms);
    run<1, 32, 2, 8>("HPT1 G32 V2 shared->REDUX->uniform", E, pts, out, sms);
    return 0;
}
