// K2 (scoring), K3 (finalise + select) and K4 (inlier mask of one hypothesis).
#pragma once
#include <type_traits>

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
// (TMA) into a kStages-deep ring signalled through mbarriers; a stage is refilled by
// whichever warp finishes it last, so no warp ever waits for a slower one (warps may drift
// kStages-1 tiles apart) and there is no block barrier in the loop.
//
// Persistent blocks: the grid is one wave (SMs x resident blocks); work items
// (pair, correspondence split, hypothesis block) are claimed from an atomic counter, which
// balances the uneven cost of "good" hypotheses.
//
// Two-level evaluation.  G correspondences are evaluated per step with a cheap
// division-free test (SCREEN: 12 FP64 issue slots using only the image-A distance, a
// necessary condition; FULL: the 21-slot two-sided decision).  Candidates (~1 %) are pushed
// to a per-warp ring and processed 32 at a time by all lanes (dense, no divergence): the
// exact reference-order SED is evaluated and compared with thr.
//
// Exact accumulation.  An inlier's sed (and sed^2) is split into three 23-bit chunks of a
// fixed-point number scaled so that thr < 2^0 maps below 2^69; chunks are added with native
// 32-bit shared-memory atomics, folded into 64-bit registers of the owning lane, and finally
// into 64-bit global accumulators.  Integer addition is associative, so the sums are EXACT
// (every double s >= thr*2^-16 is represented without rounding) and independent of warp
// scheduling, split count and GPU count — run-to-run deterministic by construction, and
// closer to the true sum than any floating-point summation order (numpy's included).
// ------------------------------------------------------------------------------------
constexpr int kScoreThreads = 128;
constexpr int kScoreWarps = kScoreThreads / 32;
constexpr int kTile = 128;    // correspondences per stage (4 KB)
constexpr int kStages = 4;
constexpr unsigned kMaxPoints = 1u << 26;
constexpr int kAccWords = 7;  // count, 3 chunks of sum(sed), 3 chunks of sum(sed^2)
constexpr int kFlushEvery = 8;

struct ScoreArgs {
    const Corr* pts;
    long long n;
    const long long* offsets;  // [npairs+1] or null (single pair of n records)
    const double* E;           // [npairs][h][9]
    long long h;
    double thr, thr_pre;
    double scale1, scale2;     // 2^(23-e), 2^(23-2e) with thr < 2^e
    long long chunk;           // correspondences per split (multiple of kTile)
    int hblocks, nsplit;
    long long total_items;
    long long htotal;          // npairs * h
    unsigned* work_counter;
    unsigned long long* acc;   // [kAccWords][npairs*h] exact integer accumulators (pre-zeroed)
};

__device__ __forceinline__ unsigned atom_add_acq_rel_shared(unsigned* p, unsigned v) {
    unsigned old;
    asm volatile("atom.acq_rel.cta.shared::cta.add.u32 %0, [%1], %2;" : "=r"(old) : "r"(smem_u32(p)), "r"(v) : "memory");
    return old;
}

// three 23-bit chunks of floor(x * 2^46), x in [0, 2^23)
__device__ __forceinline__ void chunks23(double x, unsigned& c2, unsigned& c1, unsigned& c0) {
    c2 = __double2uint_rz(x);
    const double r1 = (x - (double)c2) * 8388608.0;
    c1 = __double2uint_rz(r1);
    const double r0 = (r1 - (double)c1) * 8388608.0;
    c0 = __double2uint_rn(r0);
}

template <int HPT, int G, bool SCREEN>
__global__ void __launch_bounds__(kScoreThreads) k_score(const ScoreArgs a) {
    constexpr int RING = (HPT * G >= 8) ? 512 : 64 * HPT * G;  // >= new entries of one push slice + 31 pending
    __shared__ __align__(128) Corr tile[kStages][kTile];
    __shared__ __align__(8) unsigned long long full_bar[kStages];
    __shared__ unsigned done[kStages];
    __shared__ unsigned ring[kScoreWarps][RING];
    __shared__ unsigned ring_tail[kScoreWarps];
    __shared__ unsigned sacc[kScoreWarps][HPT][kAccWords][32];
    __shared__ unsigned s_item;

    const unsigned full = 0xffffffffu;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;

    if (threadIdx.x == 0) {
#pragma unroll
        for (int s = 0; s < kStages; ++s) { mbar_init(&full_bar[s], 1); done[s] = 0; }
        mbar_fence_init();
    }
    if (lane == 0) ring_tail[warp] = 0;
#pragma unroll
    for (int j = 0; j < HPT; ++j)
#pragma unroll
        for (int k = 0; k < kAccWords; ++k) sacc[warp][j][k][lane] = 0;
    unsigned* q = ring[warp];
    unsigned head = 0;        // entries consumed so far (warp-uniform, monotone)
    unsigned gt = 0;          // tiles consumed so far by this block (drives stage + parity)

    for (;;) {
        __syncthreads();  // previous item completely finished; s_item free
        if (threadIdx.x == 0) s_item = atomicAdd(a.work_counter, 1u);
        __syncthreads();
        const unsigned item = s_item;
        if (item >= a.total_items) break;
        const int hb = (int)(item % (unsigned)a.hblocks);
        const unsigned rest = item / (unsigned)a.hblocks;
        const int split = (int)(rest % (unsigned)a.nsplit);
        const int pair = (int)(rest / (unsigned)a.nsplit);

        const long long pbase = a.offsets ? a.offsets[pair] : 0;
        const long long plen = a.offsets ? a.offsets[pair + 1] - pbase : a.n;
        long long begin = (long long)split * a.chunk;
        if (begin > plen) begin = plen;
        const long long end = pbase + ((begin + a.chunk < plen) ? begin + a.chunk : plen);
        begin += pbase;
        const int ntiles = (int)((end - begin + kTile - 1) / kTile);
        const long long hyp_base = (long long)hb * (kScoreThreads * HPT) + threadIdx.x;
        const double* Ep = a.E + 9 * (long long)pair * a.h;

        auto issue = [&](int t) {
            const int s = (int)((gt + (unsigned)t) % kStages);
            const long long first = begin + (long long)t * kTile;
            const long long rem = end - first;
            const uint32_t bytes = (uint32_t)((rem < kTile ? rem : kTile) * sizeof(Corr));
            mbar_expect_tx(&full_bar[s], bytes);
            bulk_g2s(&tile[s][0], a.pts + first, bytes, &full_bar[s]);
        };
        if (threadIdx.x == 0)
            for (int t = 0; t < kStages && t < ntiles; ++t) issue(t);

        double e[HPT][9];
#pragma unroll
        for (int j = 0; j < HPT; ++j) {
            const long long hyp = hyp_base + (long long)j * kScoreThreads;
#pragma unroll
            for (int k = 0; k < 9; ++k) e[j][k] = (hyp < a.h) ? Ep[9 * hyp + k] : 0.0;
        }
        unsigned long long tot[HPT][kAccWords];
#pragma unroll
        for (int j = 0; j < HPT; ++j)
#pragma unroll
            for (int k = 0; k < kAccWords; ++k) tot[j][k] = 0ull;
        int drains = 0;

        // fold the warp's shared 32-bit chunk sums into the owning lanes' 64-bit registers
        auto flush = [&]() {
            __syncwarp();
#pragma unroll
            for (int j = 0; j < HPT; ++j)
#pragma unroll
                for (int k = 0; k < kAccWords; ++k) {
                    tot[j][k] += sacc[warp][j][k][lane];
                    sacc[warp][j][k][lane] = 0;
                }
            __syncwarp();
            drains = 0;
        };

        // exact evaluation of m (<= 32) queued candidates by all 32 lanes
        auto drain = [&](unsigned m) {
            const unsigned ent = (lane < (int)m) ? q[(head + lane) & (RING - 1)] : ((unsigned)lane << 27);
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
            if ((lane < (int)m) && (sv <= a.thr)) {  // ransac.py:73  score <= threshold
                unsigned c2, c1, c0;
                unsigned* dst = &sacc[warp][HPT == 2 ? slot : 0][0][owner];
                atomicAdd(dst, 1u);
                chunks23(sv * a.scale1, c2, c1, c0);
                atomicAdd(dst + 32, c2);
                atomicAdd(dst + 64, c1);
                atomicAdd(dst + 96, c0);
                chunks23(__dmul_rn(sv, sv) * a.scale2, c2, c1, c0);
                atomicAdd(dst + 128, c2);
                atomicAdd(dst + 160, c1);
                atomicAdd(dst + 192, c0);
            }
            head += m;
            if (++drains >= kFlushEvery) flush();
        };

        // one step: G correspondences starting at tile record p (global index gi0)
        auto step = [&](const Corr* tp, int p, unsigned gi0, auto gtag) {
            constexpr int GG = decltype(gtag)::value;
            Corr c[GG];
#pragma unroll
            for (int g = 0; g < GG; ++g) c[g] = tp[p + g];
            unsigned pm = 0;
#pragma unroll
            for (int j = 0; j < HPT; ++j)
#pragma unroll
                for (int g = 0; g < GG; ++g) {
                    const double d = SCREEN ? sed_screen(e[j], c[g].xa, c[g].ya, c[g].xb, c[g].yb, a.thr_pre)
                                            : sed_full_decision(e[j], c[g].xa, c[g].ya, c[g].xb, c[g].yb, a.thr_pre);
                    pm |= ((unsigned)__double2hiint(d) >> 31) << (j * GG + g);
                }
            // push in slices of <= 8 bit positions (<= 256 new entries) so that the ring never overflows
#pragma unroll
            for (int c0 = 0; c0 < HPT * GG; c0 += 8) {
                unsigned mbits = (pm >> c0) & 0xffu;
                if (__any_sync(full, mbits != 0u)) {
                    while (mbits) {  // usually one bit in a few lanes
                        const int k = c0 + __ffs(mbits) - 1;
                        mbits &= mbits - 1;
                        const unsigned pos = atomicAdd(&ring_tail[warp], 1u);
                        q[pos & (RING - 1)] = ((unsigned)lane << 27) | ((unsigned)(k / GG) << 26) | (gi0 + (unsigned)(k % GG));
                    }
                    __syncwarp();
                    const unsigned tail = *(volatile unsigned*)&ring_tail[warp];
                    while (tail - head >= 32u) drain(32u);
                }
            }
        };

        for (int t = 0; t < ntiles; ++t) {
            const unsigned gti = gt + (unsigned)t;
            const int s = (int)(gti % kStages);
            mbar_wait(&full_bar[s], (gti / kStages) & 1u);
            const long long first = begin + (long long)t * kTile;
            const int np = (int)((end - first < kTile) ? (end - first) : kTile);
            const Corr* tp = tile[s];
            int p = 0;
            for (; p + G <= np; p += G) step(tp, p, (unsigned)(first + p), std::integral_constant<int, G>{});
            for (; p < np; ++p) step(tp, p, (unsigned)(first + p), std::integral_constant<int, 1>{});
            // release the stage: the last warp to finish it refills it (nobody waits)
            __syncwarp();
            if (lane == 0) {
                const unsigned old = atom_add_acq_rel_shared(&done[s], 1u);
                if (old == kScoreWarps - 1) {
                    done[s] = 0;
                    if (t + kStages < ntiles) issue(t + kStages);
                }
            }
        }
        gt += (unsigned)ntiles;

        // tail of the queue, then publish this item's exact sums
        {
            __syncwarp();
            const unsigned tail = *(volatile unsigned*)&ring_tail[warp];
            while (tail != head) drain(tail - head < 32u ? tail - head : 32u);
            flush();
        }
#pragma unroll
        for (int j = 0; j < HPT; ++j) {
            const long long hyp = hyp_base + (long long)j * kScoreThreads;
            if (hyp < a.h && tot[j][0]) {
                const long long HT = a.htotal;
                unsigned long long* dst = a.acc + (long long)pair * a.h + hyp;
#pragma unroll
                for (int k = 0; k < kAccWords; ++k)
                    if (tot[j][k]) atomicAdd(dst + (long long)k * HT, tot[j][k]);
            }
        }
    }
}

// ------------------------------------------------------------------------------------
// K2, resident-range form (the default).
//
// Same evaluation, queue and exact accumulation as k_score above, different data movement:
// a block keeps ONE range of correspondences (<= kMaxRange records, one TMA bulk copy)
// resident in shared memory and its four warps independently claim sets of 32*HPT
// hypotheses from a per-range global counter until the range has met every hypothesis.
// No per-tile mbarrier waits, no stage hand-off, no block barrier inside a range, and warps
// that draw expensive ("good") hypotheses do not hold the others back; several blocks can
// share a range (their sets come from the same counter), which is how a single large pair
// fills the machine.  Range "units" and their (pair, first, count) live in a small table
// built by the host; blocks claim (unit, replica) slots from a global counter.
// ------------------------------------------------------------------------------------
constexpr int kMaxRange = 1344;  // 42 KB of Corr: four blocks per SM

struct RangeUnit {
    long long first;  // global record index
    int count;
    int pair;
};

struct ScoreResArgs {
    const Corr* pts;
    const double* E;  // [npairs][h][9]
    long long h;
    long long htotal;
    double thr, thr_pre, scale1, scale2;
    const RangeUnit* units;
    int nunits;
    unsigned total_slots;      // nunits * replicas
    unsigned* slot_counter;    // 1
    unsigned* set_counters;    // [nunits]
    unsigned long long* acc;   // [kAccWords][htotal]
};

template <int HPT, int G, bool SCREEN>
__global__ void __launch_bounds__(kScoreThreads) k_score_res(const ScoreResArgs a) {
    constexpr int RING = 512 * HPT;  // one push slice adds <= 256*HPT entries to <= 31 pending
    extern __shared__ __align__(128) unsigned char res_smem[];
    Corr* range = reinterpret_cast<Corr*>(res_smem);
    __shared__ __align__(8) unsigned long long full_bar;
    __shared__ unsigned ring[kScoreWarps][RING];
    __shared__ unsigned ring_tail[kScoreWarps];
    __shared__ unsigned sacc[kScoreWarps][HPT][kAccWords][32];
    __shared__ unsigned s_slot;

    const unsigned full = 0xffffffffu;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) {
        mbar_init(&full_bar, 1);
        mbar_fence_init();
    }
    if (lane == 0) ring_tail[warp] = 0;
#pragma unroll
    for (int j = 0; j < HPT; ++j)
#pragma unroll
        for (int k = 0; k < kAccWords; ++k) sacc[warp][j][k][lane] = 0;
    unsigned* q = ring[warp];
    unsigned head = 0;
    const unsigned nsets = (unsigned)((a.h + 32 * HPT - 1) / (32 * HPT));

    for (unsigned round = 0;; ++round) {
        __syncthreads();  // every warp is done with the resident range
        if (threadIdx.x == 0) s_slot = atomicAdd(a.slot_counter, 1u);
        __syncthreads();
        const unsigned slot = s_slot;
        if (slot >= a.total_slots) break;
        const int unit = (int)(slot % (unsigned)a.nunits);
        const RangeUnit u = a.units[unit];
        if (threadIdx.x == 0) {
            const uint32_t bytes = (uint32_t)u.count * (uint32_t)sizeof(Corr);
            mbar_expect_tx(&full_bar, bytes);
            bulk_g2s(range, a.pts + u.first, bytes, &full_bar);
        }
        mbar_wait(&full_bar, round & 1u);
        const double* Ep = a.E + 9 * (long long)u.pair * a.h;
        const int np = u.count;

        for (;;) {
            unsigned set = 0;
            if (lane == 0) set = atomicAdd(&a.set_counters[unit], 1u);
            set = __shfl_sync(full, set, 0);
            if (set >= nsets) break;
            const long long hyp0 = (long long)set * (32 * HPT) + lane;

            double e[HPT][9];
#pragma unroll
            for (int j = 0; j < HPT; ++j) {
                const long long hyp = hyp0 + 32 * j;
#pragma unroll
                for (int k = 0; k < 9; ++k) e[j][k] = (hyp < a.h) ? Ep[9 * hyp + k] : 0.0;
            }
            unsigned long long tot[HPT][kAccWords];
#pragma unroll
            for (int j = 0; j < HPT; ++j)
#pragma unroll
                for (int k = 0; k < kAccWords; ++k) tot[j][k] = 0ull;
            int drains = 0;

            // fold the warp's shared 32-bit chunk sums into the owning lanes' 64-bit registers
            auto flush = [&]() {
                __syncwarp();
#pragma unroll
                for (int j = 0; j < HPT; ++j)
#pragma unroll
                    for (int k = 0; k < kAccWords; ++k) {
                        tot[j][k] += sacc[warp][j][k][lane];
                        sacc[warp][j][k][lane] = 0;
                    }
                __syncwarp();
                drains = 0;
            };

            // exact evaluation of m (<= 32) queued candidates by all 32 lanes
            auto drain = [&](unsigned m) {
                const unsigned ent = (lane < (int)m) ? q[(head + lane) & (RING - 1)] : ((unsigned)lane << 27);
                const int owner = (int)(ent >> 27);
                const int slot_j = (int)((ent >> 26) & 1u);
                const unsigned pi = ent & 0xffffu;  // index into the resident range
                double eo[9];
#pragma unroll
                for (int k = 0; k < 9; ++k) {
                    double v = __shfl_sync(full, e[0][k], owner);
                    if (HPT == 2) {
                        const double v1 = __shfl_sync(full, e[HPT - 1][k], owner);
                        v = slot_j ? v1 : v;
                    }
                    eo[k] = v;
                }
                const Corr c = range[pi];
                const double sv = sed_exact(eo, c.xa, c.ya, c.xb, c.yb);
                if ((lane < (int)m) && (sv <= a.thr)) {  // ransac.py:73  score <= threshold
                    unsigned c2, c1, c0;
                    unsigned* dst = &sacc[warp][HPT == 2 ? slot_j : 0][0][owner];
                    atomicAdd(dst, 1u);
                    chunks23(sv * a.scale1, c2, c1, c0);
                    atomicAdd(dst + 32, c2);
                    atomicAdd(dst + 64, c1);
                    atomicAdd(dst + 96, c0);
                    chunks23(__dmul_rn(sv, sv) * a.scale2, c2, c1, c0);
                    atomicAdd(dst + 128, c2);
                    atomicAdd(dst + 160, c1);
                    atomicAdd(dst + 192, c0);
                }
                head += m;
                if (++drains >= kFlushEvery) flush();
            };

            // 32 correspondences per chunk: screen them G at a time into per-lane bit masks,
            // then push the survivors of the whole chunk in one go
            for (int p0 = 0; p0 < np; p0 += 32) {
                const int nv = (np - p0 < 32) ? (np - p0) : 32;
                unsigned pm[HPT];
#pragma unroll
                for (int j = 0; j < HPT; ++j) pm[j] = 0u;
#pragma unroll 1
                for (int g0 = 0; g0 < nv; g0 += G) {
#pragma unroll
                    for (int g = 0; g < G; ++g) {
                        int pi = p0 + g0 + g;
                        pi = (pi < np) ? pi : (np - 1);  // clamped duplicates are masked off below
                        const Corr c = range[pi];
#pragma unroll
                        for (int j = 0; j < HPT; ++j) {
                            const double d = SCREEN ? sed_screen(e[j], c.xa, c.ya, c.xb, c.yb, a.thr_pre)
                                                    : sed_full_decision(e[j], c.xa, c.ya, c.xb, c.yb, a.thr_pre);
                            pm[j] |= ((unsigned)__double2hiint(d) >> 31) << (g0 + g);
                        }
                    }
                }
                const unsigned vmask = (nv == 32) ? 0xffffffffu : ((1u << nv) - 1u);
                bool any = false;
#pragma unroll
                for (int j = 0; j < HPT; ++j) {
                    pm[j] &= vmask;
                    any |= pm[j] != 0u;
                }
                if (__any_sync(full, any)) {
#pragma unroll 1
                    for (int c0 = 0; c0 < 32; c0 += 8) {  // slices of 8 bit positions bound the ring
#pragma unroll
                        for (int j = 0; j < HPT; ++j) {
                            unsigned mbits = (pm[j] >> c0) & 0xffu;
                            while (mbits) {
                                const int k = c0 + __ffs(mbits) - 1;
                                mbits &= mbits - 1;
                                const unsigned pos = atomicAdd(&ring_tail[warp], 1u);
                                q[pos & (RING - 1)] = ((unsigned)lane << 27) | ((unsigned)j << 26) | (unsigned)(p0 + k);
                            }
                        }
                        __syncwarp();
                        const unsigned tail = *(volatile unsigned*)&ring_tail[warp];
                        while (tail - head >= 32u) drain(32u);
                    }
                }
            }
            {
                __syncwarp();
                const unsigned tail = *(volatile unsigned*)&ring_tail[warp];
                while (tail != head) drain(tail - head < 32u ? tail - head : 32u);
                flush();
            }
#pragma unroll
            for (int j = 0; j < HPT; ++j) {
                const long long hyp = hyp0 + 32 * j;
                if (hyp < a.h && tot[j][0]) {
                    unsigned long long* dst = a.acc + (long long)u.pair * a.h + hyp;
#pragma unroll
                    for (int k = 0; k < kAccWords; ++k)
                        if (tot[j][k]) atomicAdd(dst + (long long)k * a.htotal, tot[j][k]);
                }
            }
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

// (c2 * 2^46 + c1 * 2^23 + c0) as a double; the chunk sums are < 2^49 each, the total < 2^96.
__device__ __forceinline__ double fixed69_to_double(unsigned long long c2, unsigned long long c1,
                                                    unsigned long long c0) {
    const unsigned __int128 v = ((unsigned __int128)c2 << 46) + ((unsigned __int128)c1 << 23) + c0;
    const unsigned long long hi = (unsigned long long)(v >> 64), lo = (unsigned long long)v;
    return fma((double)hi, 18446744073709551616.0, (double)lo);
}

struct FinalArgs {
    const Corr* pts;
    const long long* offsets;
    const double* E;
    const uint8_t* valid;
    const int32_t* table;  // may be null: no sample rule
    long long h;
    long long idx_offset;  // global index of hypothesis 0 (hypothesis-sharded runs)
    long long htotal;                 // npairs * h (stride of the accumulator planes)
    const unsigned long long* acc;    // [kAccWords][htotal] exact sums from K2
    double inv_scale1, inv_scale2;    // 2^(e-69), 2^(2e-69)
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
        // exact integer sums -> one rounding each
        long long cnt = (long long)a.acc[i];
        double s1 = fixed69_to_double(a.acc[1 * a.htotal + i], a.acc[2 * a.htotal + i], a.acc[3 * a.htotal + i]) * a.inv_scale1;
        double s2 = fixed69_to_double(a.acc[4 * a.htotal + i], a.acc[5 * a.htotal + i], a.acc[6 * a.htotal + i]) * a.inv_scale2;
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
