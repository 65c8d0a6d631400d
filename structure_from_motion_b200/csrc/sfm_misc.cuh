// Small utility kernels: inlier compaction, winner gathering, FP64 peak probe.
#pragma once
#include "sfm_device.cuh"
#include "sfm_score.cuh"

namespace sfm {

// lib/ransac/ransac.py:76 — the 8 sample points are part of the returned inliers whatever
// their score: force them into the winner's mask.
__global__ void k_mark_samples(const int32_t* __restrict__ row, uint8_t* __restrict__ mask) {
    if (threadIdx.x < 8 && blockIdx.x == 0) mask[row[threadIdx.x]] = 1;
}
// the same with the winner taken from K3's device output (no host round trip between selection and the tail)
__global__ void k_mark_samples_dev(const int32_t* __restrict__ table, const Best* __restrict__ best,
                                   long long idx_offset, uint8_t* __restrict__ mask) {
    const long long local = best->idx - idx_offset;
    if (best->idx < 0 || threadIdx.x >= 8 || blockIdx.x != 0) return;
    mask[table[8 * local + threadIdx.x]] = 1;
}

// Stream compaction of a byte mask into ascending indices: per-1024-element block counts,
// single-block exclusive scan, ordered scatter.  scan[nblocks] receives the total.
__global__ void __launch_bounds__(256) k_compact_count(const uint8_t* __restrict__ mask, long long n,
                                                       long long* __restrict__ scan) {
    __shared__ int s;
    if (threadIdx.x == 0) s = 0;
    __syncthreads();
    const long long base = (long long)blockIdx.x * 1024;
    int c = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const long long i = base + k * 256 + threadIdx.x;
        c += (i < n && mask[i]) ? 1 : 0;
    }
    for (int d = 16; d > 0; d >>= 1) c += __shfl_down_sync(0xffffffffu, c, d);
    if ((threadIdx.x & 31) == 0 && c) atomicAdd(&s, c);
    __syncthreads();
    if (threadIdx.x == 0) scan[blockIdx.x] = s;
}

__global__ void __launch_bounds__(1024) k_compact_scan(long long* __restrict__ scan, int nblocks) {
    __shared__ long long warp_tot[32];
    __shared__ long long carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int base = 0; base < nblocks; base += 1024) {
        const int i = base + threadIdx.x;
        const long long v = (i < nblocks) ? scan[i] : 0;
        long long x = v;
        for (int d = 1; d < 32; d <<= 1) {
            const long long y = __shfl_up_sync(0xffffffffu, x, d);
            if (lane >= d) x += y;
        }
        if (lane == 31) warp_tot[warp] = x;
        __syncthreads();
        if (warp == 0) {
            long long w = warp_tot[lane];
            for (int d = 1; d < 32; d <<= 1) {
                const long long y = __shfl_up_sync(0xffffffffu, w, d);
                if (lane >= d) w += y;
            }
            warp_tot[lane] = w;
        }
        __syncthreads();
        const long long before = carry + (warp ? warp_tot[warp - 1] : 0) + x - v;
        if (i < nblocks) scan[i] = before;
        __syncthreads();
        if (threadIdx.x == 1023) carry = before + v;
        __syncthreads();
    }
    if (threadIdx.x == 0) scan[nblocks] = carry;
}

__global__ void __launch_bounds__(256) k_compact_scatter(const uint8_t* __restrict__ mask, long long n,
                                                         const long long* __restrict__ scan,
                                                         long long* __restrict__ idx) {
    __shared__ int warp_cnt[4][8];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const long long base = (long long)blockIdx.x * 1024;
    bool f[4];
    unsigned b[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const long long i = base + k * 256 + threadIdx.x;
        f[k] = (i < n) && mask[i];
        b[k] = __ballot_sync(0xffffffffu, f[k]);
        if (lane == 0) warp_cnt[k][warp] = __popc(b[k]);
    }
    __syncthreads();
    long long off = scan[blockIdx.x];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        long long o = off;
        for (int w = 0; w < 8; ++w) {
            if (w < warp) o += warp_cnt[k][w];
            off += warp_cnt[k][w];
        }
        if (f[k]) idx[o + __popc(b[k] & ((1u << lane) - 1u))] = base + k * 256 + threadIdx.x;
    }
}

__global__ void k_gather_winners(const Best* __restrict__ best, const double* __restrict__ E, long long h,
                                 int npairs, double* __restrict__ out) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= npairs) return;
    const long long i = best[p].idx;
    for (int k = 0; k < 9; ++k) out[9 * p + k] = (i >= 0) ? E[9 * ((long long)p * h + i) + k] : 0.0;
}

// 16 independent DFMA chains per thread: the FP64 pipe's sustained rate.
__global__ void __launch_bounds__(256) k_fp64_peak(double* __restrict__ out, int iters, double m) {
    double a[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) a[k] = 1.0 + 1e-3 * (threadIdx.x + k);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int k = 0; k < 16; ++k) a[k] = fma(a[k], m, 1e-9);
    }
    double s = 0.0;
#pragma unroll
    for (int k = 0; k < 16; ++k) s += a[k];
    if (s == 123.456) out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

}  // namespace sfm
