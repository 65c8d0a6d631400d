// N2 — the first stage of apps/sfm.py:64-71: Harris corner detection.
//
// Restates lib/harris/harris_detector.py:11-113 and lib/common/correlate.py:4-39.  The reference
// runs Python double loops over the pixels (one np.dot per pixel for the Sobel responses, three
// np.sum + np.linalg.det per pixel for the cornerness, a sequential in-place non-maximum
// suppression).  Here
//   C1 k_cross_correlate   thread per pixel, any odd square kernel, zero "same" border (correlate.py)
//   C2 k_sobel_pair        both Sobel responses in one pass (harris_detector.py:107-112)
//   C3 k_cornerness        block_size x block_size sums of Ix^2, IxIy, Iy^2; det(M) - k trace(M)^2;
//                          negative values clamped to 0 (harris_detector.py:28, 58-86)
//   C4 k_nms_*             harris_detector.py:95-104 suppresses IN PLACE while scanning row-major, so
//                          a pixel is compared with the already-suppressed values of its upper and left
//                          neighbours.  That is a dependency along strictly increasing values only; it is
//                          solved as a fixed point: alive(p) = no later neighbour is larger and every
//                          earlier neighbour is either not larger or not alive.  Iterated in parallel
//                          until nothing changes (level k of the dependency DAG is final after k sweeps).
//   C5 k_corner_* + bitonic  compaction of the surviving non-zero pixels and a bitonic sort by
//                          (cornerness descending, flat index descending) = np.flip(np.argsort(.)) with
//                          ties in descending index order (numpy's order among exact ties is
//                          unspecified); the first num_corners are the features.
// uint8 images make every Sobel response, product and block sum an exact integer in fp64, so the
// cornerness differs from the reference only through its LAPACK/log/exp determinant (1e-13 relative).
#pragma once
#include "sfm_device.cuh"
#include "sfm_match.cuh"

namespace sfm {

// C1.  correlate.py:24-37: np.dot of the flattened window with the flattened kernel, border = 0.
__global__ void k_cross_correlate(const void* __restrict__ img, int dtype, int rows, int cols,
                                  const double* __restrict__ kernel, int ksize, double* __restrict__ out) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= (long long)rows * cols) return;
    const int r = (int)(i / cols), c = (int)(i % cols), h = ksize / 2;
    double acc = 0.0;
    if (r >= h && r < rows - h && c >= h && c < cols - h) {
        for (int dr = 0; dr < ksize; ++dr)
            for (int dc = 0; dc < ksize; ++dc)
                acc = fma(load_pixel(img, dtype, (long long)(r - h + dr) * cols + (c - h + dc)),
                          kernel[dr * ksize + dc], acc);
    }
    out[i] = acc;
}

// C2.  Sobel x = [[-1,0,1],[-2,0,2],[-1,0,1]], Sobel y = its transpose, written as the same
// nine-term dot products (zero terms included: they do not change an fp64 sum).
__global__ void k_sobel_pair(const void* __restrict__ img, int dtype, int rows, int cols, double* __restrict__ gx,
                             double* __restrict__ gy) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= (long long)rows * cols) return;
    const int r = (int)(i / cols), c = (int)(i % cols);
    double sx = 0.0, sy = 0.0;
    if (r >= 1 && r < rows - 1 && c >= 1 && c < cols - 1) {
        double p[3][3];
#pragma unroll
        for (int dr = 0; dr < 3; ++dr)
#pragma unroll
            for (int dc = 0; dc < 3; ++dc) p[dr][dc] = load_pixel(img, dtype, (long long)(r - 1 + dr) * cols + (c - 1 + dc));
        const double kx[3][3] = {{-1, 0, 1}, {-2, 0, 2}, {-1, 0, 1}};
#pragma unroll
        for (int dr = 0; dr < 3; ++dr)
#pragma unroll
            for (int dc = 0; dc < 3; ++dc) {
                sx = fma(p[dr][dc], kx[dr][dc], sx);
                sy = fma(p[dr][dc], kx[dc][dr], sy);
            }
    }
    gx[i] = sx;
    gy[i] = sy;
}

// C3.  Output is out_rows x out_cols (= height/width - int(np.around(block_size / 2)), :66-72); only
// rows < height - block_size and cols < width - block_size are computed (:74-84), the rest stays 0.
__global__ void k_cornerness(const double* __restrict__ gx, const double* __restrict__ gy, int rows, int cols,
                             int block, double k, int out_rows, int out_cols, double* __restrict__ out) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= (long long)out_rows * out_cols) return;
    const int r = (int)(i / out_cols), c = (int)(i % out_cols);
    double v = 0.0;
    if (r < rows - block && c < cols - block) {
        double a = 0.0, b = 0.0, d = 0.0;  // Ix2, IxIy, Iy2
        for (int dr = 0; dr < block; ++dr)
            for (int dc = 0; dc < block; ++dc) {
                const double x = gx[(long long)(r + dr) * cols + c + dc], y = gy[(long long)(r + dr) * cols + c + dc];
                a = __dadd_rn(a, __dmul_rn(x, x));
                b = __dadd_rn(b, __dmul_rn(x, y));
                d = __dadd_rn(d, __dmul_rn(y, y));
            }
        const double det = __dsub_rn(__dmul_rn(a, d), __dmul_rn(b, b));
        const double tr = __dadd_rn(a, d);
        v = __dsub_rn(det, __dmul_rn(k, __dmul_rn(tr, tr)));
        if (v < 0.0) v = 0.0;  // harris_detector.py:28
    }
    out[i] = v;
}

// C4.  alive0: every pixel starts alive; each sweep recomputes alive from the previous sweep.
__global__ void k_nms_sweep(const double* __restrict__ v, int rows, int cols, const uint8_t* __restrict__ alive_in,
                            uint8_t* __restrict__ alive_out, int* __restrict__ changed) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= (long long)rows * cols) return;
    const int r = (int)(i / cols), c = (int)(i % cols);
    const double p = v[i];
    bool alive = true;
    for (int dr = -1; dr <= 1 && alive; ++dr)
        for (int dc = -1; dc <= 1; ++dc) {
            const int rr = r + dr, cc = c + dc;
            if ((dr == 0 && dc == 0) || rr < 0 || rr >= rows || cc < 0 || cc >= cols) continue;
            const long long j = (long long)rr * cols + cc;
            if (!(p < v[j])) continue;
            const bool earlier = dr < 0 || (dr == 0 && dc < 0);  // already visited by the row-major scan
            if (!earlier || alive_in[j]) { alive = false; break; }
        }
    const uint8_t a = alive ? 1 : 0;
    if (a != alive_in[i]) *changed = 1;
    alive_out[i] = a;
}

__global__ void k_nms_apply(double* __restrict__ v, long long n, const uint8_t* __restrict__ alive) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i < n && !alive[i]) v[i] = 0.0;
}

// C5.  Candidates: the non-zero pixels after suppression (harris_detector.py:35-40 prunes zeros).
__global__ void k_corner_compact(const double* __restrict__ v, long long n, unsigned long long* __restrict__ key,
                                 unsigned* __restrict__ idx, unsigned* __restrict__ count) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    const bool take = i < n && v[i] != 0.0;
    const unsigned m = __ballot_sync(0xffffffffu, take);
    if (!m) return;
    const int lane = threadIdx.x & 31;
    unsigned base = 0;
    if (lane == __ffs(m) - 1) base = atomicAdd(count, __popc(m));
    base = __shfl_sync(0xffffffffu, base, __ffs(m) - 1);
    if (take) {
        const unsigned o = base + __popc(m & ((1u << lane) - 1));
        key[o] = ordered_key(v[i]);
        idx[o] = (unsigned)i;
    }
}

// pad [count, padded) with the smallest key so the padding sorts last in descending order
__global__ void k_corner_pad(unsigned long long* __restrict__ key, unsigned* __restrict__ idx,
                             const unsigned* __restrict__ count, unsigned padded) {
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < padded && i >= *count) { key[i] = 0ull; idx[i] = 0u; }
}

// descending by (key, idx)
__device__ __forceinline__ bool corner_before(unsigned long long ka, unsigned ia, unsigned long long kb, unsigned ib) {
    return ka > kb || (ka == kb && ia > ib);
}

constexpr int kSortTile = 1024;  // elements per block in the shared-memory stages (512 threads)

// all stages with k <= kSortTile (mode 0: full local sort), or the steps j < kSortTile of stage k (mode 1)
__global__ void __launch_bounds__(kSortTile / 2)
k_bitonic_local(unsigned long long* __restrict__ key, unsigned* __restrict__ idx, unsigned k_stage, int mode) {
    __shared__ unsigned long long sk[kSortTile];
    __shared__ unsigned si[kSortTile];
    const unsigned base = blockIdx.x * kSortTile;
    for (int t = threadIdx.x; t < kSortTile; t += blockDim.x) { sk[t] = key[base + t]; si[t] = idx[base + t]; }
    __syncthreads();
    const unsigned k_first = mode == 0 ? 2u : k_stage, k_last = mode == 0 ? (unsigned)kSortTile : k_stage;
    for (unsigned k = k_first; k <= k_last; k <<= 1) {
        for (unsigned j = (k > (unsigned)kSortTile ? kSortTile : k) >> 1; j > 0; j >>= 1) {
            const unsigned t = threadIdx.x;
            const unsigned lo = 2 * t - (t & (j - 1));  // index with bit j clear
            const unsigned hi = lo + j;
            const bool desc = (((base + lo) & k) == 0);  // this half-block sorts "before-first"
            const bool swap = desc ? corner_before(sk[hi], si[hi], sk[lo], si[lo]) : corner_before(sk[lo], si[lo], sk[hi], si[hi]);
            if (swap) {
                const unsigned long long tk = sk[lo]; sk[lo] = sk[hi]; sk[hi] = tk;
                const unsigned ti = si[lo]; si[lo] = si[hi]; si[hi] = ti;
            }
            __syncthreads();
        }
    }
    for (int t = threadIdx.x; t < kSortTile; t += blockDim.x) { key[base + t] = sk[t]; idx[base + t] = si[t]; }
}

// one step (k, j) with j >= kSortTile in global memory
__global__ void k_bitonic_global(unsigned long long* __restrict__ key, unsigned* __restrict__ idx, unsigned n,
                                 unsigned k, unsigned j) {
    const unsigned t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n / 2) return;
    const unsigned lo = 2 * t - (t & (j - 1)), hi = lo + j;
    const bool desc = ((lo & k) == 0);
    const unsigned long long kl = key[lo], kh = key[hi];
    const unsigned il = idx[lo], ih = idx[hi];
    const bool swap = desc ? corner_before(kh, ih, kl, il) : corner_before(kl, il, kh, ih);
    if (swap) { key[lo] = kh; key[hi] = kl; idx[lo] = ih; idx[hi] = il; }
}

// first min(count, num_corners) sorted candidates -> (x, y) = (col, row) + block_size / 2 (:47-53)
__global__ void k_corner_emit(const unsigned long long* __restrict__ key, const unsigned* __restrict__ idx,
                              const unsigned* __restrict__ count, long long num, int out_cols, double offset,
                              const double* __restrict__ v, double* __restrict__ xy, double* __restrict__ score) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= num || i >= (long long)*count) return;
    const unsigned p = idx[i];
    xy[2 * i] = (double)(p % (unsigned)out_cols) + offset;
    xy[2 * i + 1] = (double)(p / (unsigned)out_cols) + offset;
    score[i] = v[p];
}

}  // namespace sfm
