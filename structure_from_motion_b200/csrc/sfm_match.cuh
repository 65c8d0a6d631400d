// N1 — the stage right before the hot path: brute-force feature matching with
// NCC / SSD patch scores.
//
// Restates lib/feature_matching/matching.py:36-118 (match_brute_force with the RATIO_TEST and
// CROSSCHECK validations), ncc.py:7-54, ssd.py:7-36 and util.py:8-27.  The reference pushes
// Na*Nb Python objects through heapq; here
//   M1 k_patch_prepare   one thread per feature: bounds test, window gather, mean shift, sum of
//                        squares in numpy's pairwise order  -> W[n][w*w], ss[n], ok[n]
//   M2 k_patch_scores    score matrix S[Na][Nb] (a small GEMM over the window axis, shared-memory
//                        tiled), final operations exactly as ncc.py:47-52 / ssd.py:33-36
//   M3 k_match_select    one warp per feature of image A: heap[0] (first minimum) and heap[1]
//                        of the heap the reference builds, WITHOUT building it — heap[1] is the
//                        root of the left subtree, i.e. the minimum over the pushes whose
//                        position k+1 has binary prefix '10' of max(score_k, running minimum
//                        before k); then the ratio test (IEEE division, NaN fails)
//   M4 k_cross_*         cross-check: per feature of B the smallest (score, a) among the
//                        surviving matches (order-preserving 64-bit keys + atomicMin)
// Everything is fp64; image pixels may be uint8 (what cv.cvtColor hands to the reference,
// including numpy's uint8 wrap-around in ssd.py) or float64.
#pragma once
#include "sfm_device.cuh"

namespace sfm {

constexpr int kMaxWindow = 15;                       // window_size <= 15 (225 values per patch)
enum { SCORE_NCC = 0, SCORE_SSD = 1 };
enum { IMG_U8 = 0, IMG_F64 = 1 };
enum { VALIDATE_RATIO = 1, VALIDATE_CROSSCHECK = 2 };

__device__ __forceinline__ double load_pixel(const void* img, int dtype, long long i) {
    return dtype == IMG_U8 ? (double)reinterpret_cast<const unsigned char*>(img)[i]
                           : reinterpret_cast<const double*>(img)[i];
}

// numpy's pairwise summation of n <= 128 contiguous doubles (the np.sum / np.mean kernel):
// eight running sums over the multiples of 8, a fixed tree, then the tail in order.
template <class F>
__device__ __forceinline__ double numpy_pairwise_sum(int n, F get) {
    if (n < 8) {
        double r = 0.0;
        for (int i = 0; i < n; ++i) r = __dadd_rn(r, get(i));
        return r;
    }
    double r[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) r[j] = get(j);
    int i = 8;
    for (; i < n - (n % 8); i += 8) {
#pragma unroll
        for (int j = 0; j < 8; ++j) r[j] = __dadd_rn(r[j], get(i + j));
    }
    double res = __dadd_rn(__dadd_rn(__dadd_rn(r[0], r[1]), __dadd_rn(r[2], r[3])),
                           __dadd_rn(__dadd_rn(r[4], r[5]), __dadd_rn(r[6], r[7])));
    for (; i < n; ++i) res = __dadd_rn(res, get(i));
    return res;
}

// M1.  util.py:8-27 + ncc.py:33-45.  W holds, per feature, the window values (SSD) or the
// mean-shifted window values (NCC); ss the sum of squares of the shifted values (NCC).
__global__ void k_patch_prepare(const void* __restrict__ img, int dtype, long long rows, long long cols,
                                const double* __restrict__ feats, long long n, int window, int kind,
                                double* __restrict__ W, double* __restrict__ ss, uint8_t* __restrict__ ok) {
    const long long f = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (f >= n) return;
    const double x = feats[2 * f], y = feats[2 * f + 1];
    const int half = window / 2;  // int(window_size / 2)
    // is_within_bounds compares the float coordinates (util.py:12-16)
    const bool inside = ((double)half <= y) && (y < (double)(rows - half)) && ((double)half <= x) &&
                        (x < (double)(cols - half));
    ok[f] = inside ? 1 : 0;
    const int ww = window * window;
    double* w = W + f * (long long)ww;
    if (!inside) {
        for (int i = 0; i < ww; ++i) w[i] = 0.0;
        ss[f] = 0.0;
        return;
    }
    const long long x0 = (long long)x - half, y0 = (long long)y - half;  // int(feature.x): truncation
    for (int r = 0; r < window; ++r)
        for (int c = 0; c < window; ++c) w[r * window + c] = load_pixel(img, dtype, (y0 + r) * cols + (x0 + c));
    if (kind == SCORE_NCC) {
        // np.mean: pixel sums of uint8 windows are exact in any order; float images follow the
        // contiguous pairwise order (the window is a strided view there: documented 1e-13 tolerance)
        const double mu = numpy_pairwise_sum(ww, [&](int i) { return w[i]; }) / (double)ww;
        for (int i = 0; i < ww; ++i) w[i] = __dsub_rn(w[i], mu);
        ss[f] = numpy_pairwise_sum(ww, [&](int i) { return __dmul_rn(w[i], w[i]); });  // np.sum(np.square(.))
    } else {
        ss[f] = 0.0;
    }
}

// M2.  16x16 output tile per block, the window axis streamed through shared memory.
constexpr int kScoreTile = 16;
__global__ void __launch_bounds__(kScoreTile * kScoreTile)
k_patch_scores(const double* __restrict__ Wa, const double* __restrict__ ssa, const uint8_t* __restrict__ oka,
               long long na, const double* __restrict__ Wb, const double* __restrict__ ssb,
               const uint8_t* __restrict__ okb, long long nb, int ww, int kind, int dtype,
               double* __restrict__ S) {
    __shared__ double sa[kScoreTile][kScoreTile + 1], sb[kScoreTile][kScoreTile + 1];
    const int tx = threadIdx.x % kScoreTile, ty = threadIdx.x / kScoreTile;
    const long long a = blockIdx.y * (long long)kScoreTile + ty, b = blockIdx.x * (long long)kScoreTile + tx;
    double acc = 0.0;
    unsigned long long iacc = 0;  // uint8 SSD: numpy squares the wrapped uint8 difference in uint8
    for (int k0 = 0; k0 < ww; k0 += kScoreTile) {
        const long long ra = blockIdx.y * (long long)kScoreTile + ty, rb = blockIdx.x * (long long)kScoreTile + ty;
        sa[ty][tx] = (ra < na && k0 + tx < ww) ? Wa[ra * ww + k0 + tx] : 0.0;
        sb[ty][tx] = (rb < nb && k0 + tx < ww) ? Wb[rb * ww + k0 + tx] : 0.0;
        __syncthreads();
        const int kn = (ww - k0 < kScoreTile) ? ww - k0 : kScoreTile;
        if (kind == SCORE_NCC) {
            for (int k = 0; k < kn; ++k) acc = fma(sa[ty][k], sb[tx][k], acc);  // np.dot (ncc.py:42)
        } else if (dtype == IMG_U8) {
            for (int k = 0; k < kn; ++k) {
                const unsigned d = ((unsigned)sa[ty][k] - (unsigned)sb[tx][k]) & 255u;  // uint8 - uint8 wraps
                iacc += (d * d) & 255u;                                                  // np.square stays uint8
            }
        } else {
            for (int k = 0; k < kn; ++k) {
                const double d = __dsub_rn(sa[ty][k], sb[tx][k]);
                acc = __dadd_rn(acc, __dmul_rn(d, d));
            }
        }
        __syncthreads();
    }
    if (a >= na || b >= nb) return;
    double s;
    const bool inside = oka[a] && okb[b];
    if (kind == SCORE_NCC) {
        if (!inside) s = 2.0;  // ncc.py:28-31
        else {
            const double den = sqrt(__dmul_rn(ssa[a], ssb[b]));  // ncc.py:43-45
            if (den == 0.0) s = 2.0;                             // ncc.py:47-48
            else s = __dadd_rn(__dmul_rn(__ddiv_rn(acc, den), -1.0), 1.0);  // ncc.py:50-52
        }
    } else {
        if (!inside) s = __longlong_as_double(0x7ff0000000000000LL);  // ssd.py:27-30
        else s = __ddiv_rn(dtype == IMG_U8 ? (double)iacc : acc, (double)ww);  // ssd.py:36
    }
    S[a * nb + b] = s;
}

// M3.  One warp per feature of A.  Outputs the first minimum (heap[0]), the score of heap[1] and
// the ratio-test verdict (matching.py:84-97; with no RATIO_TEST every feature is kept).
__global__ void __launch_bounds__(128)
k_match_select(const double* __restrict__ S, long long na, long long nb, int validation, double ratio_thr,
               int32_t* __restrict__ best_b, double* __restrict__ best_s, double* __restrict__ heap1,
               uint8_t* __restrict__ keep) {
    const unsigned full = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const long long a = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
    if (a >= na) return;
    const double inf = __longlong_as_double(0x7ff0000000000000LL);
    const double* s = S + a * nb;
    double run = inf;        // minimum of the scores before the current chunk (+inf before the first push)
    long long arg = -1;      // its first position
    double h1 = inf;
    bool have_h1 = false;
    for (long long k0 = 0; k0 < nb; k0 += 32) {
        const long long k = k0 + lane;
        const bool in = k < nb;
        const double v = in ? s[k] : inf;
        // inclusive prefix minimum inside the chunk, first position on ties (strict <, matching.py:22-24)
        double pmin = v;
        long long parg = k;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const double om = __shfl_up_sync(full, pmin, d);
            const long long oa = __shfl_up_sync(full, parg, d);
            if (lane >= d && !(pmin < om)) { pmin = om; parg = oa; }   // keep the earlier one unless strictly smaller
        }
        // exclusive running minimum in front of position k
        double before = __shfl_up_sync(full, pmin, 1);
        if (lane == 0) before = inf;
        before = (before < run) ? before : run;
        // the loser of the comparison between push k and the root goes down the insertion path
        if (in && k >= 1) {
            const unsigned long long p = (unsigned long long)k + 1ull;
            const int top = 63 - __clzll((long long)p);
            const bool left = ((p >> (top - 1)) == 2ull);   // binary prefix '10'
            if (left) {
                const double loser = (v < before) ? before : v;
                if (!have_h1 || loser < h1) h1 = loser;
                have_h1 = true;
            }
        }
        // fold the chunk into the running minimum (first position wins ties)
        const double cm = __shfl_sync(full, pmin, 31);
        const long long ca = __shfl_sync(full, parg, 31);
        if (arg < 0 || cm < run) { run = cm; arg = ca; }
    }
    // heap[1] over the lanes
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        const double o = __shfl_xor_sync(full, h1, d);
        const bool oh = __shfl_xor_sync(full, have_h1 ? 1 : 0, d) != 0;
        if (oh && (!have_h1 || o < h1)) h1 = o;
        have_h1 = have_h1 || oh;
    }
    if (lane == 0) {
        best_b[a] = (int32_t)arg;
        best_s[a] = run;
        heap1[a] = have_h1 ? h1 : __longlong_as_double(0x7ff8000000000000LL);
        bool k = nb >= 1;
        if ((validation & VALIDATE_RATIO) && nb > 1) k = (__ddiv_rn(run, h1) <= ratio_thr);  // NaN fails
        keep[a] = k ? 1 : 0;
    }
}

// order-preserving map double -> uint64 (handles the slightly negative NCC scores)
__device__ __forceinline__ unsigned long long ordered_key(double v) {
    const unsigned long long b = (unsigned long long)__double_as_longlong(v);
    return (b & 0x8000000000000000ull) ? ~b : (b | 0x8000000000000000ull);
}

// M4.  matching.py:100-118: per feature of B keep the match with the smallest score, the earliest
// feature of A on ties; a match survives iff it is that one.
__global__ void k_cross_min_score(const int32_t* __restrict__ best_b, const double* __restrict__ best_s,
                                  const uint8_t* __restrict__ keep, long long na,
                                  unsigned long long* __restrict__ key_b) {
    const long long a = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (a >= na || !keep[a]) return;
    atomicMin(key_b + best_b[a], ordered_key(best_s[a]));
}
__global__ void k_cross_min_index(const int32_t* __restrict__ best_b, const double* __restrict__ best_s,
                                  const uint8_t* __restrict__ keep, long long na,
                                  const unsigned long long* __restrict__ key_b, int* __restrict__ first_a) {
    const long long a = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (a >= na || !keep[a]) return;
    if (ordered_key(best_s[a]) == key_b[best_b[a]]) atomicMin(first_a + best_b[a], (int)a);
}
__global__ void k_cross_filter(const int32_t* __restrict__ best_b, long long na, const int* __restrict__ first_a,
                               uint8_t* __restrict__ keep) {
    const long long a = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (a >= na || !keep[a]) return;
    if (first_a[best_b[a]] != (int)a) keep[a] = 0;
}

}  // namespace sfm
