// K2 (scoring), K3 (finalise + select) and K4 (inlier mask of one hypothesis).
#pragma once
#include "sfm_device.cuh"

namespace sfm {

// ------------------------------------------------------------------------------------
// mbarrier / bulk-copy (TMA) primitives, raw PTX for sm_100a
// ------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(void* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(void* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(void* bar, uint32_t parity) {
    uint32_t ok;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
    } while (!ok);
}
// 1-D bulk async copy global -> shared, completion signalled on an mbarrier (SASS: UBLKCP).
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, void* bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
            smem_u32(dst_smem)),
        "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}

// ------------------------------------------------------------------------------------
// K2 — score H hypotheses against N correspondences.
//
// Restates the inner loop of fit_with_ransac (lib/ransac/ransac.py:66-82) with
// calculate_sed_inlier_score (lib/epipolar/epipolar_ransac.py:18-25, lib/epipolar/sed.py:7-30)
// as the scorer: for every hypothesis, the number of correspondences with sed <= thr and
// the sums of sed and sed^2 over them.  (The 8 sample points are scored like any other
// point here; K3 applies the "samples are not thresholded but always counted in the
// error" rule of ransac.py:63-64,76.)
//
// Mapping: lane = hypothesis.  Each thread keeps HPT essential matrices in registers and
// streams correspondences from shared-memory tiles (every lane of a warp reads the same
// 32-byte record: a broadcast, conflict-free).  Tiles are filled by 1-D bulk async copies
// (TMA) into a two-stage ring, signalled through mbarriers, so loads overlap the FP64
// work.  Accumulators are private to the owning lane: no atomics, no floating-point
// shuffles, and a fixed summation order (point order) => run-to-run deterministic.
//
// Two-level evaluation: a cheap division-free test (SCREEN: 12 FP64 issue slots using only
// the image-A distance, a necessary condition; FULL: the 21-slot two-sided decision) runs
// for every (hypothesis, correspondence).  Candidates (typically < 1 %) are appended to a
// per-warp queue and processed 32 at a time by all lanes (dense, no divergence): the exact
// reference-order SED is evaluated, compared with thr, and the result is routed back to
// the owning lane in queue order.
// ------------------------------------------------------------------------------------
constexpr int kScoreThreads = 128;
constexpr int kTile = 256;
constexpr unsigned kMaxPoints = 1u << 26;

struct ScoreArgs {
    const Corr* pts;
    long long n;
    const long long* offsets;  // [npairs+1] or null (single pair of n records)
    const double* E;  // [npairs][h][9]
    long long h;
    double thr, thr_pre;
    long long chunk;  // correspondences per split (multiple of kTile)
    int32_t* pcount;  // [npairs][nsplit][h]
    double* ps1;
    double* ps2;
};

template <int HPT, bool SCREEN>
__global__ void __launch_bounds__(kScoreThreads) k_score(const ScoreArgs a) {
    __shared__ __align__(128) Corr tile[2][kTile];
    __shared__ __align__(8) unsigned long long full_bar[2];
    __shared__ unsigned queue[kScoreThreads / 32][64];

    const unsigned full = 0xffffffffu;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned lt_mask = (1u << lane) - 1u;
    const long long hyp_base = (long long)blockIdx.x * (kScoreThreads * HPT) + threadIdx.x;
    // blockIdx.z = image pair, blockIdx.y = split of that pair's correspondences
    const long long pbase = a.offsets ? a.offsets[blockIdx.z] : 0;
    const long long plen = a.offsets ? a.offsets[blockIdx.z + 1] - pbase : a.n;
    long long begin = (long long)blockIdx.y * a.chunk;
    if (begin > plen) begin = plen;
    const long long end = pbase + ((begin + a.chunk < plen) ? begin + a.chunk : plen);
    begin += pbase;
    const int ntiles = (int)((end - begin + kTile - 1) / kTile);
    const double* Ep = a.E + 9 * (long long)blockIdx.z * a.h;

    double e[HPT][9];
#pragma unroll
    for (int j = 0; j < HPT; ++j) {
        const long long hyp = hyp_base + (long long)j * kScoreThreads;
#pragma unroll
        for (int k = 0; k < 9; ++k) e[j][k] = (hyp < a.h) ? Ep[9 * hyp + k] : 0.0;
    }
    int cnt[HPT];
    double s1[HPT], s2[HPT];
#pragma unroll
    for (int j = 0; j < HPT; ++j) { cnt[j] = 0; s1[j] = 0.0; s2[j] = 0.0; }

    auto issue = [&](int t) {
        const int s = t & 1;
        const long long first = begin + (long long)t * kTile;
        const long long rem = end - first;
        const uint32_t bytes = (uint32_t)((rem < kTile ? rem : kTile) * sizeof(Corr));
        mbar_expect_tx(&full_bar[s], bytes);
        bulk_g2s(&tile[s][0], a.pts + first, bytes, &full_bar[s]);
    };
    if (threadIdx.x == 0) {
        mbar_init(&full_bar[0], 1);
        mbar_init(&full_bar[1], 1);
        mbar_fence_init();
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        if (ntiles > 0) issue(0);
        if (ntiles > 1) issue(1);
    }

    unsigned qhead = 0, qn = 0;
    unsigned* q = queue[warp];

    // Process m (<= 32) queued candidates with all 32 lanes.
    auto drain = [&](unsigned m) {
        __syncwarp();
        const unsigned ent = (lane < (int)m) ? q[(qhead + lane) & 63u] : ((unsigned)lane << 27);
        const int owner = (int)(ent >> 27);
        const int slot = (int)((ent >> 26) & 1u);
        const unsigned gi = ent & (kMaxPoints - 1u);
        double eo[9];
#pragma unroll
        for (int k = 0; k < 9; ++k) {
            double v = __shfl_sync(full, e[0][k], owner);
            if (HPT == 2) {
                const double v1 = __shfl_sync(full, e[HPT - 1][k], owner);
                v = slot ? v1 : v;
            }
            eo[k] = v;
        }
        const Corr c = a.pts[gi];
        const double sv = sed_exact(eo, c.xa, c.ya, c.xb, c.yb);
        const bool inl = (lane < (int)m) && (sv <= a.thr);  // ransac.py:73  score <= threshold
        unsigned mask = __ballot_sync(full, inl);
        while (mask) {
            const int src = __ffs(mask) - 1;
            mask &= mask - 1;
            const unsigned oe = __shfl_sync(full, ent, src);
            const double v = __shfl_sync(full, sv, src);
            const bool mine = ((int)(oe >> 27) == lane);
            const int sl = (int)((oe >> 26) & 1u);
#pragma unroll
            for (int j = 0; j < HPT; ++j) {
                if (mine && sl == j) {
                    cnt[j] += 1;
                    s1[j] = __dadd_rn(s1[j], v);
                    s2[j] = __dadd_rn(s2[j], __dmul_rn(v, v));
                }
            }
        }
        qhead = (qhead + m) & 63u;
        qn -= m;
        __syncwarp();
    };

    for (int t = 0; t < ntiles; ++t) {
        const int s = t & 1;
        mbar_wait(&full_bar[s], (uint32_t)((t >> 1) & 1));
        const long long first = begin + (long long)t * kTile;
        const int np = (int)((end - first < kTile) ? (end - first) : kTile);
        const Corr* tp = tile[s];
#pragma unroll 4
        for (int p = 0; p < np; ++p) {
            const Corr c = tp[p];
            bool pass[HPT];
            bool any = false;
#pragma unroll
            for (int j = 0; j < HPT; ++j) {
                const double d = SCREEN ? sed_screen(e[j], c.xa, c.ya, c.xb, c.yb, a.thr_pre)
                                        : sed_full_decision(e[j], c.xa, c.ya, c.xb, c.yb, a.thr_pre);
                pass[j] = __double2hiint(d) < 0;
                any |= pass[j];
            }
            if (__any_sync(full, any)) {
                const unsigned gi = (unsigned)(first + p);
#pragma unroll
                for (int j = 0; j < HPT; ++j) {
                    const unsigned b = __ballot_sync(full, pass[j]);
                    if (b) {
                        if (pass[j])
                            q[(qhead + qn + __popc(b & lt_mask)) & 63u] =
                                ((unsigned)lane << 27) | ((unsigned)j << 26) | gi;
                        qn += __popc(b);
                        if (qn >= 32u) drain(32u);
                    }
                }
            }
        }
        __syncthreads();  // every warp is done reading stage s
        if (threadIdx.x == 0 && t + 2 < ntiles) issue(t + 2);
    }
    while (qn > 0u) drain(qn < 32u ? qn : 32u);

#pragma unroll
    for (int j = 0; j < HPT; ++j) {
        const long long hyp = hyp_base + (long long)j * kScoreThreads;
        if (hyp < a.h) {
            const long long o = ((long long)blockIdx.z * gridDim.y + blockIdx.y) * a.h + hyp;
            a.pcount[o] = cnt[j];
            a.ps1[o] = s1[j];
            a.ps2[o] = s2[j];
        }
    }
}

// ------------------------------------------------------------------------------------
// K3 — finalise + select (lib/ransac/ransac.py:70-86, 96-108).
//
// Per hypothesis: combine the per-split partials in split order; apply the sample rule
// (ransac.py:63-64,76: the 8 sample points are excluded from the threshold count and
// included unconditionally in the error); n = 8 + count_extra; aggregate
// (SUM / SQUARE / MEAN / RMS); candidate iff valid and min_extra <= count_extra
// (ransac.py:75; min_extra may be fractional).  Then argmin of the error with the lowest
// index winning ties (ransac.py:83: strict <, earliest iteration kept).  mode 1 selects by
// maximum inlier count instead (lowest index on ties; non-default).
// ------------------------------------------------------------------------------------
enum { AGG_SUM = 0, AGG_SQUARE = 1, AGG_MEAN = 2, AGG_RMS = 3 };
enum { SELECT_MIN_ERROR = 0, SELECT_MAX_INLIERS = 1 };

struct Best {
    double err;
    long long idx;
    int count;
    int pad;
};

__device__ __forceinline__ bool better(const Best& x, const Best& y, int mode) {
    // is x better than y ?
    if (y.idx < 0) return x.idx >= 0;
    if (x.idx < 0) return false;
    if (mode == SELECT_MAX_INLIERS) {
        if (x.count != y.count) return x.count > y.count;
        return x.idx < y.idx;
    }
    if (x.err != y.err) return x.err < y.err;
    return x.idx < y.idx;
}

__device__ __forceinline__ Best shfl_best(const Best& b, int src_delta) {
    Best r;
    r.err = __shfl_down_sync(0xffffffffu, b.err, src_delta);
    r.idx = __shfl_down_sync(0xffffffffu, b.idx, src_delta);
    r.count = __shfl_down_sync(0xffffffffu, b.count, src_delta);
    r.pad = 0;
    return r;
}

__device__ __forceinline__ Best block_best(Best b, int mode, Best* sm /* 32 */) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        const Best o = shfl_best(b, d);
        if (better(o, b, mode)) b = o;
    }
    if (lane == 0) sm[warp] = b;
    __syncthreads();
    if (warp == 0) {
        const int nw = (blockDim.x + 31) >> 5;
        Best x;
        if (lane < nw) x = sm[lane];
        else { x.err = 0.0; x.idx = -1; x.count = 0; x.pad = 0; }
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) {
            const Best o = shfl_best(x, d);
            if (better(o, x, mode)) x = o;
        }
        b = x;
    }
    return b;  // valid in thread 0
}

struct FinalArgs {
    const Corr* pts;
    const long long* offsets;
    const double* E;
    const uint8_t* valid;
    const int32_t* table;  // may be null: no sample rule
    long long h;
    long long idx_offset;  // global index of hypothesis 0 (hypothesis-sharded runs)
    int nsplit;
    const int32_t* pcount;
    const double* ps1;
    const double* ps2;
    double thr, min_extra;
    int agg, mode;
    int32_t* count_extra;
    double* S1;
    double* S2;
    double* err;
    Best* block_out;
};

__global__ void __launch_bounds__(256) k_finalise(const FinalArgs a) {
    __shared__ Best sm[32];
    const long long li = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    const long long i = (long long)blockIdx.y * a.h + li;  // blockIdx.y = image pair
    const Corr* pts = a.pts + (a.offsets ? a.offsets[blockIdx.y] : 0);
    Best b;
    b.err = 0.0; b.idx = -1; b.count = 0; b.pad = 0;
    if (li < a.h) {
        long long cnt = 0;
        double s1 = 0.0, s2 = 0.0;
        for (int s = 0; s < a.nsplit; ++s) {
            const long long o = ((long long)blockIdx.y * a.nsplit + s) * a.h + li;
            cnt += a.pcount[o];
            s1 = __dadd_rn(s1, a.ps1[o]);
            s2 = __dadd_rn(s2, a.ps2[o]);
        }
        const bool valid = a.valid ? (a.valid[i] != 0) : true;
        if (a.table && valid) {
            double e[9];
#pragma unroll
            for (int k = 0; k < 9; ++k) e[k] = a.E[9 * i + k];
            for (int k = 0; k < 8; ++k) {
                const Corr c = pts[a.table[8 * i + k]];
                const double sv = sed_exact(e, c.xa, c.ya, c.xb, c.yb);
                if (sv <= a.thr) {
                    cnt -= 1;  // was counted by K2, but samples are not "extra" inliers
                } else {
                    s1 = __dadd_rn(s1, sv);  // not counted by K2, but always part of the error
                    s2 = __dadd_rn(s2, __dmul_rn(sv, sv));
                }
            }
        }
        const double n = (double)((a.table ? 8 : 0) + cnt);
        double err;
        switch (a.agg) {
            case AGG_SUM: err = s1; break;
            case AGG_SQUARE: err = s2; break;
            case AGG_MEAN: err = s1 / n; break;
            default: err = sqrt(s2 / n); break;
        }
        const bool cand = valid && (a.min_extra <= (double)cnt) && (err == err);
        a.count_extra[i] = valid ? (int32_t)cnt : -1;
        a.S1[i] = s1;
        a.S2[i] = s2;
        a.err[i] = cand ? err : __longlong_as_double(0x7ff0000000000000LL);
        if (cand) { b.err = err; b.idx = a.idx_offset + li; b.count = (int)cnt; }
    }
    b = block_best(b, a.mode, sm);
    if (threadIdx.x == 0) a.block_out[(long long)blockIdx.y * gridDim.x + blockIdx.x] = b;
}

// Single block: reduce per-block bests; also counts invalid hypotheses (ransac.py:65 has no
// try/except around the fitter, so one degenerate sample aborts the reference run).
__global__ void __launch_bounds__(256)
k_select(const Best* __restrict__ blocks, int nblocks, int mode, const uint8_t* __restrict__ valid,
         long long h, long long idx_offset, Best* __restrict__ out, long long* __restrict__ invalid_out) {
    __shared__ Best sm[32];
    __shared__ long long s_ninv, s_first;
    if (threadIdx.x == 0) { s_ninv = 0; s_first = 0x7fffffffffffffffLL; }
    __syncthreads();
    // blockIdx.x = image pair
    blocks += (long long)blockIdx.x * nblocks;
    if (valid) valid += (long long)blockIdx.x * h;
    out += blockIdx.x;
    invalid_out += 2 * (long long)blockIdx.x;
    Best b;
    b.err = 0.0; b.idx = -1; b.count = 0; b.pad = 0;
    for (int i = threadIdx.x; i < nblocks; i += blockDim.x) {
        const Best o = blocks[i];
        if (better(o, b, mode)) b = o;
    }
    long long ninv = 0, first = 0x7fffffffffffffffLL;
    if (valid) {
        for (long long i = threadIdx.x; i < h; i += blockDim.x) {
            if (!valid[i]) { ++ninv; if (i + idx_offset < first) first = i + idx_offset; }
        }
        if (ninv) {
            atomicAdd((unsigned long long*)&s_ninv, (unsigned long long)ninv);
            atomicMin(&s_first, first);
        }
    }
    b = block_best(b, mode, sm);
    __syncthreads();
    if (threadIdx.x == 0) {
        *out = b;
        invalid_out[0] = s_ninv;
        invalid_out[1] = s_ninv ? s_first : -1;
    }
}

// ------------------------------------------------------------------------------------
// K4 — inlier mask + SED values of one hypothesis over all correspondences (the winner):
// the same exact scorer, one thread per correspondence.
// ------------------------------------------------------------------------------------
__global__ void k_inlier_mask(const Corr* __restrict__ pts, long long n, const double* __restrict__ E,
                              const Best* __restrict__ best, long long idx_offset, long long hyp,
                              double thr, uint8_t* __restrict__ mask, double* __restrict__ sed) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= n) return;
    long long local = hyp;
    if (best) local = best->idx - idx_offset;  // take the winner straight from K3's output
    if (local < 0) {
        mask[i] = 0;
        sed[i] = __longlong_as_double(0x7ff8000000000000LL);
        return;
    }
    double e[9];
#pragma unroll
    for (int k = 0; k < 9; ++k) e[k] = E[9 * local + k];
    const Corr c = pts[i];
    const double sv = sed_exact(e, c.xa, c.ya, c.xb, c.yb);
    sed[i] = sv;
    mask[i] = (sv <= thr) ? 1 : 0;
}

}  // namespace sfm
