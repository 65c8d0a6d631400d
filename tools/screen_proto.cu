// Prototype of the K2 screening inner loop under different operand-delivery schemes.
// Question: how close to the FP64 pipe rate (2 cycles per warp-DFMA per SMSP) can the
// 11-DFMA test run, and what delivers the warp-uniform correspondence operand best?
//   MODE 0  shared memory (LDS.128 broadcast) + compiler scheduling           (what K2 v5 does)
//   MODE 1  shared memory + order pinned with volatile asm (hypothesis-innermost => .reuse)
//   MODE 2  constant memory (LDCU -> uniform registers) + compiler scheduling
//   MODE 3  constant memory + pinned order
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/bin/screen_proto tools/screen_proto.cu
#include <cstdio>
#include <cuda_runtime.h>

struct __align__(32) Corr { double xa, ya, xb, yb; };
#ifndef NPTS
#define NPTS 512
#endif
#ifndef ROTATE
#define ROTATE 0
#endif
__constant__ Corr c_pts[NPTS];

#define VFMA(d, a, b, c) asm volatile("fma.rn.f64 %0, %1, %2, %3;" : "=d"(d) : "d"(a), "d"(b), "d"(c))

template <int HPT, int G, int MODE, int UNR = G, int XTRA = 0>
__global__ void __launch_bounds__(128) k(const double* __restrict__ E, const Corr* __restrict__ pts, int reps,
                                         unsigned* __restrict__ out) {
    __shared__ __align__(128) Corr tile[MODE < 2 ? (NPTS > 1024 ? 1024 : NPTS) : 1];
    if (MODE < 2) {
        for (int i = threadIdx.x; i < (NPTS > 1024 ? 1024 : NPTS); i += blockDim.x) tile[i] = pts[i];
        __syncthreads();
    }
    double e[HPT][9], kap[HPT];
#pragma unroll
    for (int j = 0; j < HPT; ++j) {
#pragma unroll
        for (int q = 0; q < 9; ++q) e[j][q] = E[((blockIdx.x * HPT + j) * 128 + threadIdx.x) * 9 + q];
        kap[j] = 1e-30 * e[j][0];
    }
    unsigned acc = 0, dummy = 0;
    // ROTATE: every warp starts at its own offset, so the warps of an SM stream different parts of the buffer
    const int rot = ROTATE ? (int)(((blockIdx.x * 4 + (threadIdx.x >> 5)) * 7919u) % (NPTS / G)) * G : 0;
    for (int r = 0; r < reps; ++r) {
        for (int p0 = 0; p0 < NPTS; p0 += G) {
            const int p = __shfl_sync(0xffffffffu, (p0 + rot) % NPTS, 0);
            unsigned pm = 0;
#pragma unroll UNR
            for (int g = 0; g < G; ++g) {
                const Corr c = (MODE < 2) ? tile[(p + g) & 1023] : c_pts[p + g];
                double t0[HPT], t1[HPT], t2[HPT], d[HPT];
                if (MODE & 1) {
#pragma unroll
                    for (int j = 0; j < HPT; ++j) { VFMA(t0[j], c.yb, e[j][3], e[j][6]); VFMA(t1[j], c.yb, e[j][4], e[j][7]); VFMA(t2[j], c.yb, e[j][5], e[j][8]); }
#pragma unroll
                    for (int j = 0; j < HPT; ++j) { VFMA(t0[j], c.xb, e[j][0], t0[j]); VFMA(t1[j], c.xb, e[j][1], t1[j]); VFMA(t2[j], c.xb, e[j][2], t2[j]); }
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
                for (int j = 0; j < HPT; ++j) {
                    if (XTRA < 0) pm |= (unsigned)__double2hiint(d[j]);  // sensitivity test: OR only (LOP3 merges three)
                    else pm = __funnelshift_l((unsigned)__double2hiint(d[j]), pm, 1);
                    if (XTRA > 0) dummy = __funnelshift_l(dummy ^ (unsigned)__double2loint(d[j]), pm, 3);  // one more SHF
                }
            }
            if (XTRA < 0) pm &= 0x80000000u;
            if (__any_sync(0xffffffffu, pm != 0u)) acc += pm;
        }
    }
    out[blockIdx.x * 128 + threadIdx.x] = acc + dummy;
}


// Lockstep variant: PP correspondences x HPT hypotheses advance through the 7 stages together, so every
// stage offers PP*HPT (x3 for the first two) independent DFMAs to the scheduler.
template <int HPT, int G, int PP, int MINB>
__global__ void __launch_bounds__(128, MINB) klock(const double* __restrict__ E, const Corr* __restrict__ pts, int reps,
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
#pragma unroll
            for (int g = 0; g < G; g += PP) {
                Corr c[PP];
#pragma unroll
                for (int u = 0; u < PP; ++u) c[u] = tile[(p + g + u) & 1023];
                double t0[PP][HPT], t1[PP][HPT], t2[PP][HPT], d[PP][HPT];
#pragma unroll
                for (int u = 0; u < PP; ++u)
#pragma unroll
                    for (int j = 0; j < HPT; ++j) { t0[u][j] = fma(c[u].yb, e[j][3], e[j][6]); t1[u][j] = fma(c[u].yb, e[j][4], e[j][7]); t2[u][j] = fma(c[u].yb, e[j][5], e[j][8]); }
#pragma unroll
                for (int u = 0; u < PP; ++u)
#pragma unroll
                    for (int j = 0; j < HPT; ++j) { t0[u][j] = fma(c[u].xb, e[j][0], t0[u][j]); t1[u][j] = fma(c[u].xb, e[j][1], t1[u][j]); t2[u][j] = fma(c[u].xb, e[j][2], t2[u][j]); }
#pragma unroll
                for (int u = 0; u < PP; ++u)
#pragma unroll
                    for (int j = 0; j < HPT; ++j) t2[u][j] = fma(c[u].ya, t1[u][j], t2[u][j]);
#pragma unroll
                for (int u = 0; u < PP; ++u)
#pragma unroll
                    for (int j = 0; j < HPT; ++j) t1[u][j] = fma(t1[u][j], t1[u][j], kap[j]);
#pragma unroll
                for (int u = 0; u < PP; ++u)
#pragma unroll
                    for (int j = 0; j < HPT; ++j) t2[u][j] = fma(c[u].xa, t0[u][j], t2[u][j]);
#pragma unroll
                for (int u = 0; u < PP; ++u)
#pragma unroll
                    for (int j = 0; j < HPT; ++j) t1[u][j] = fma(t0[u][j], t0[u][j], t1[u][j]);
#pragma unroll
                for (int u = 0; u < PP; ++u)
#pragma unroll
                    for (int j = 0; j < HPT; ++j) d[u][j] = fma(t2[u][j], t2[u][j], -t1[u][j]);
#pragma unroll
                for (int u = 0; u < PP; ++u)
#pragma unroll
                    for (int j = 0; j < HPT; ++j) pm = __funnelshift_l((unsigned)__double2hiint(d[u][j]), pm, 1);
            }
            if (__any_sync(0xffffffffu, pm != 0u)) acc += pm;
        }
    }
    out[blockIdx.x * 128 + threadIdx.x] = acc;
}

template <int HPT, int G, int PP, int MINB>
void runlock(const char* name, const double* E, const Corr* pts, unsigned* out, int sms) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    int occ = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, klock<HPT, G, PP, MINB>, 128, 0);
    cudaFuncAttributes fa;
    cudaFuncGetAttributes(&fa, klock<HPT, G, PP, MINB>);
    const int blocks = sms * occ, reps = 64 * 512 / NPTS;
    klock<HPT, G, PP, MINB><<<blocks, 128>>>(E, pts, 2, out);
    float best = 1e30f;
    for (int r = 0; r < 3; ++r) {
        cudaEventRecord(e0);
        klock<HPT, G, PP, MINB><<<blocks, 128>>>(E, pts, reps, out);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    const double evals = (double)blocks * 128 * HPT * NPTS * reps;
    const double warp_dfma_per_smsp = evals * 11 / 32 / (sms * 4);
    printf("%-34s regs %3d occ %d: %7.3f ms  %.3e evals/s  %.2f cycles/warp-DFMA/SMSP\n", name, fa.numRegs, occ, best,
           evals / (best * 1e-3), best * 1e-3 * 1.965e9 / warp_dfma_per_smsp);
}

template <int HPT, int G, int MODE, int UNR = G, int XTRA = 0>
void run(const char* name, const double* E, const Corr* pts, unsigned* out, int sms) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    int occ = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k<HPT, G, MODE, UNR, XTRA>, 128, 0);
    for (int bps : {2, occ}) {
        if (bps > occ) continue;
        const int blocks = sms * bps, reps = 64 * 512 / NPTS;
        k<HPT, G, MODE, UNR, XTRA><<<blocks, 128>>>(E, pts, 2, out);
        float best = 1e30f;
        for (int r = 0; r < 3; ++r) {
            cudaEventRecord(e0);
            k<HPT, G, MODE, UNR, XTRA><<<blocks, 128>>>(E, pts, reps, out);
            cudaEventRecord(e1);
            cudaEventSynchronize(e1);
            float ms;
            cudaEventElapsedTime(&ms, e0, e1);
            if (ms < best) best = ms;
        }
        const double evals = (double)blocks * 128 * HPT * NPTS * reps;
        const double warp_dfma_per_smsp = evals * 11 / 32 / (sms * 4);
        printf("%-40s blocks/SM %d (occ %d): %7.3f ms  %.3e evals/s  %.2f cycles/warp-DFMA/SMSP @1965MHz\n", name, bps, occ,
               best, evals / (best * 1e-3), best * 1e-3 * 1.965e9 / warp_dfma_per_smsp);
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) printf("  CUDA error: %s\n", cudaGetErrorString(e));
    }
}

int main() {
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    const int sms = p.multiProcessorCount;
    printf("%s, %d SMs\n", p.name, sms);
    const size_t nE = (size_t)sms * 8 * 8 * 128 * 9;
    double* E;
    Corr* pts;
    unsigned* out;
    cudaMalloc(&E, nE * 8);
    cudaMalloc(&pts, NPTS * sizeof(Corr));
    cudaMalloc(&out, (size_t)sms * 8 * 128 * 4);
    double* hE = new double[nE];
    for (size_t i = 0; i < nE; ++i) hE[i] = 0.1 + 1e-3 * (double)(i % 977);
    cudaMemcpy(E, hE, nE * 8, cudaMemcpyHostToDevice);
    Corr h[NPTS];
    for (int i = 0; i < NPTS; ++i) h[i] = {0.01 * i, 0.3 - 0.002 * i, 0.5 + 0.001 * i, -0.2 + 0.003 * i};
    cudaMemcpy(pts, h, sizeof h, cudaMemcpyHostToDevice);
    cudaMemcpyToSymbol(c_pts, h, sizeof h);
    run<2, 16, 0>("smem  HPT2 G16 baseline", E, pts, out, sms);
    run<4, 8, 0>("smem  HPT4 G8 baseline", E, pts, out, sms);
    runlock<2, 16, 1, 6>("lock HPT2 PP1 minb6", E, pts, out, sms);
    runlock<2, 16, 2, 6>("lock HPT2 PP2 minb6", E, pts, out, sms);
    runlock<2, 16, 2, 5>("lock HPT2 PP2 minb5", E, pts, out, sms);
    runlock<2, 16, 4, 5>("lock HPT2 PP4 minb5", E, pts, out, sms);
    runlock<2, 16, 2, 4>("lock HPT2 PP2 minb4", E, pts, out, sms);
    runlock<2, 16, 4, 4>("lock HPT2 PP4 minb4", E, pts, out, sms);
    runlock<2, 16, 4, 3>("lock HPT2 PP4 minb3", E, pts, out, sms);
    runlock<2, 16, 8, 3>("lock HPT2 PP8 minb3", E, pts, out, sms);
    runlock<4, 8, 1, 4>("lock HPT4 PP1 minb4", E, pts, out, sms);
    runlock<4, 8, 2, 4>("lock HPT4 PP2 minb4", E, pts, out, sms);
    runlock<4, 8, 2, 3>("lock HPT4 PP2 minb3", E, pts, out, sms);
    runlock<4, 8, 4, 2>("lock HPT4 PP4 minb2", E, pts, out, sms);
    runlock<1, 32, 4, 8>("lock HPT1 PP4 minb8", E, pts, out, sms);
    runlock<1, 32, 8, 6>("lock HPT1 PP8 minb6", E, pts, out, sms);
    runlock<1, 32, 4, 6>("lock HPT1 PP4 minb6", E, pts, out, sms);
    return 0;
}
