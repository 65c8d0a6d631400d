// K5 — essential-matrix decomposition, cheirality vote and DLT triangulation.
#pragma once
#include "sfm_device.cuh"
#include "sfm_fastsvd.cuh"
#include "sfm_linalg.cuh"

namespace sfm {

// The four pose candidates in the reference's enumeration order
// (lib/epipolar/eight_point.py:210-212): (R1,t), (R1,-t), (R2,t), (R2,-t).
struct PoseSet {
    double R[4][9];
    double t[4][3];
    double sv[3];  // singular values of E, descending
    long long counts[4];
    int best;  // argmax of counts (first maximum), -1 until voted
    int pad;
};

// _recover_all_r_t (lib/epipolar/eight_point.py:245-280): SVD of E, proper U and V,
// R1 = U W^T V^T, R2 = U W V^T, t from U Z U^T.  One-sided Jacobi SVD; U and V are made
// proper by construction (third columns are cross products), which is what the
// reference's sign fixes (:263-266) achieve.  The LAPACK sign convention is not
// reproducible, so (R1,R2) may come out swapped and t negated relative to numpy — the
// candidate *set* is identical (the reference's own test accepts exactly this ambiguity,
// lib/epipolar/tests/test_epipolar.py:205-229).
__device__ inline void decompose_essential(const double (&E)[9], PoseSet& out) {
    double u0[3], u1[3], u2[3], v0[3], v1[3], v2[3], s0, s1, s2;
    double Uf[9], Vf[9], svf[3];
    if (svd3_rank2_frames(E, Uf, Vf, svf)) {
        // closed form (sfm_fastsvd.cuh): E is rank 2 after the fit's projection, so its null vectors are cross
        // products and the rest is one 2x2 rotation - the same frames the Jacobi SVD below converges to
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            u0[i] = Uf[3 * i]; u1[i] = Uf[3 * i + 1]; u2[i] = Uf[3 * i + 2];
            v0[i] = Vf[3 * i]; v1[i] = Vf[3 * i + 1]; v2[i] = Vf[3 * i + 2];
        }
        s0 = svf[0]; s1 = svf[1]; s2 = svf[2];
    } else {
        double g[9], v[9];
#pragma unroll
        for (int i = 0; i < 9; ++i) g[i] = E[i];
        jacobi_svd_onesided<3, 3>(g, v, 30);
        double s[3];
#pragma unroll
        for (int j = 0; j < 3; ++j) s[j] = sqrt(fma(g[6 + j], g[6 + j], fma(g[3 + j], g[3 + j], g[j] * g[j])));
        // order columns by descending singular value
        int o0 = 0, o1 = 1, o2 = 2;
        if (s[o0] < s[o1]) { int t = o0; o0 = o1; o1 = t; }
        if (s[o1] < s[o2]) { int t = o1; o1 = o2; o2 = t; }
        if (s[o0] < s[o1]) { int t = o0; o0 = o1; o1 = t; }
        auto col = [&](const double (&m)[9], int j, double (&c)[3]) {
#pragma unroll
            for (int i = 0; i < 3; ++i) c[i] = (j == 0) ? m[i * 3] : ((j == 1) ? m[i * 3 + 1] : m[i * 3 + 2]);
        };
        col(g, o0, u0); col(g, o1, u1); col(v, o0, v0); col(v, o1, v1);
        s0 = (o0 == 0) ? s[0] : ((o0 == 1) ? s[1] : s[2]);
        s1 = (o1 == 0) ? s[0] : ((o1 == 1) ? s[1] : s[2]);
        s2 = (o2 == 0) ? s[0] : ((o2 == 1) ? s[1] : s[2]);
#pragma unroll
        for (int i = 0; i < 3; ++i) { u0[i] /= s0; u1[i] /= s1; }
        // proper U and V by construction: third columns are cross products
        u2[0] = u0[1] * u1[2] - u0[2] * u1[1];
        u2[1] = u0[2] * u1[0] - u0[0] * u1[2];
        u2[2] = u0[0] * u1[1] - u0[1] * u1[0];
        v2[0] = v0[1] * v1[2] - v0[2] * v1[1];
        v2[1] = v0[2] * v1[0] - v0[0] * v1[2];
        v2[2] = v0[0] * v1[1] - v0[1] * v1[0];
    }
    const double U[9] = {u0[0], u1[0], u2[0], u0[1], u1[1], u2[1], u0[2], u1[2], u2[2]};
    const double Vh[9] = {v0[0], v0[1], v0[2], v1[0], v1[1], v1[2], v2[0], v2[1], v2[2]};
    const double W[9] = {0, -1, 0, 1, 0, 0, 0, 0, 1};   // :273
    const double Wt[9] = {0, 1, 0, -1, 0, 0, 0, 0, 1};
    const double Z[9] = {0, 1, 0, -1, 0, 0, 0, 0, 0};   // :274
    double Ut[9], tmp[9], tx[9], R1[9], R2[9];
    mat3_transpose(U, Ut);
    mat3_mul(U, Z, tmp);
    mat3_mul(tmp, Ut, tx);                               // :275
    const double t1[3] = {-tx[5], tx[2], -tx[1]};        // :276
    mat3_mul(U, Wt, tmp);
    mat3_mul(tmp, Vh, R1);                               // :277
    mat3_mul(U, W, tmp);
    mat3_mul(tmp, Vh, R2);                               // :278
#pragma unroll
    for (int i = 0; i < 9; ++i) {
        out.R[0][i] = R1[i]; out.R[1][i] = R1[i]; out.R[2][i] = R2[i]; out.R[3][i] = R2[i];
    }
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        out.t[0][i] = t1[i]; out.t[1][i] = -t1[i]; out.t[2][i] = t1[i]; out.t[3][i] = -t1[i];
    }
    out.sv[0] = s0; out.sv[1] = s1; out.sv[2] = s2;
#pragma unroll
    for (int i = 0; i < 4; ++i) out.counts[i] = 0;
    out.best = -1;
    out.pad = 0;
}

__global__ void k_decompose(const double* __restrict__ E, const Best* __restrict__ best,
                            long long idx_offset, PoseSet* __restrict__ out, int npairs) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= npairs) return;
    const double* src = E + 9 * (long long)i;
    if (best) {  // single-pair pipeline: decompose the RANSAC winner in place on the device
        const long long local = best->idx - idx_offset;
        if (local < 0) { out[i].best = -2; return; }
        src = E + 9 * local;
    }
    double e[9];
#pragma unroll
    for (int k = 0; k < 9; ++k) e[k] = src[k];
    decompose_essential(e, out[i]);
}

// triangulate_point_correspondence (lib/epipolar/triangulation.py:9-39): 4x4 DLT system,
// right singular vector of the smallest singular value, dehomogenise.  P1, P2 are 3x4
// row-major (only rows 0-2 of the reference's 4x4 Tmat are used, :24-31).
// One copy per kernel (noinline): T2 calls it for the cheirality test and again for the short triangulation, and the
// second use finds the code in the instruction cache.
__device__ __noinline__ void dlt_triangulate(double xa, double ya, double xb, double yb,
                                             const double* __restrict__ P1,
                                             const double* __restrict__ P2, double (&X)[3]) {
    double g[16], v[16];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        g[0 + j] = __dsub_rn(__dmul_rn(ya, P1[8 + j]), P1[4 + j]);   // ya * P1[2,:] - P1[1,:]
        g[4 + j] = __dsub_rn(P1[j], __dmul_rn(xa, P1[8 + j]));       // P1[0,:] - xa * P1[2,:]
        g[8 + j] = __dsub_rn(__dmul_rn(yb, P2[8 + j]), P2[4 + j]);
        g[12 + j] = __dsub_rn(P2[j], __dmul_rn(xb, P2[8 + j]));
    }
    double xf[4];
    if (null_vector4_fast(g, xf)) {  // QR + checked inverse iteration (sfm_fastsvd.cuh); inliers always take this route
        X[0] = xf[0] / xf[3];
        X[1] = xf[1] / xf[3];
        X[2] = xf[2] / xf[3];
        return;
    }
    jacobi_svd_onesided<4, 4>(g, v, 20);
    double nrm[4];
#pragma unroll
    for (int j = 0; j < 4; ++j)
        nrm[j] = fma(g[12 + j], g[12 + j], fma(g[8 + j], g[8 + j], fma(g[4 + j], g[4 + j], g[j] * g[j])));
    int m = 0;
#pragma unroll
    for (int j = 1; j < 4; ++j) {
        const double cur = (m == 0) ? nrm[0] : ((m == 1) ? nrm[1] : ((m == 2) ? nrm[2] : nrm[3]));
        if (nrm[j] < cur) m = j;
    }
    double x[4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
        x[i] = (m == 0) ? v[i * 4] : ((m == 1) ? v[i * 4 + 1] : ((m == 2) ? v[i * 4 + 2] : v[i * 4 + 3]));
    X[0] = x[0] / x[3];
    X[1] = x[1] / x[3];
    X[2] = x[2] / x[3];
}

// _cheirality_check for the 4 candidates of one pair (lib/epipolar/eight_point.py:449-488,
// looped as in :210-230).  One thread per (correspondence, candidate pose): the four poses of a
// correspondence sit in four consecutive lanes (the 4x4 Jacobi SVDs are latency-bound, so the pose loop
// is spread over lanes instead of running in sequence).  pass[i] bit p = correspondence i passes pose p.
// counts follow the reference's np.count_nonzero(passing_indices) (:228-230): correspondence 0 never counts.
__global__ void __launch_bounds__(128)
k_cheirality(const Corr* __restrict__ pts, long long m, const long long* __restrict__ gather,
             const long long* __restrict__ m_dev, PoseSet* __restrict__ poses, double dist_thr,
             uint8_t* __restrict__ pass, const int32_t* __restrict__ quirk_row,
             const Best* __restrict__ quirk_best = nullptr, long long quirk_offset = 0) {
    // quirk_best != null: quirk_row is the BASE of the sample table and the winner's row is looked up on the device
    if (quirk_best) quirk_row = quirk_best->idx >= 0 ? quirk_row + 8 * (quirk_best->idx - quirk_offset) : nullptr;
    // gather != null: correspondence i is pts[gather[i]] and the count lives on the device
    const long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    const long long i = t >> 2;
    const int p = (int)(t & 3);
    if (m_dev) m = *m_dev;
    __shared__ int s_cnt[4];
    if (threadIdx.x < 4) s_cnt[threadIdx.x] = 0;
    __syncthreads();
    bool ok = false, counts = false;
    if (i < m) {
        const long long gi = gather ? gather[i] : i;
        const Corr c = pts[gi];
        const double P1[12] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0};  // Transform3D.identity() (:473)
        double P2[12];
#pragma unroll
        for (int r = 0; r < 3; ++r) {
            P2[4 * r + 0] = poses->R[p][3 * r + 0];
            P2[4 * r + 1] = poses->R[p][3 * r + 1];
            P2[4 * r + 2] = poses->R[p][3 * r + 2];
            P2[4 * r + 3] = poses->t[p][r];
        }
        double X1[3];
        dlt_triangulate(c.xa, c.ya, c.xb, c.yb, P1, P2, X1);
        const double z2 = fma(P2[8], X1[0], fma(P2[9], X1[1], fma(P2[10], X1[2], P2[11])));  // :476
        const double nrm = sqrt(fma(X1[2], X1[2], fma(X1[1], X1[1], X1[0] * X1[0])));
        ok = (X1[2] >= -kCheiralityTolerance) && (z2 >= -kCheiralityTolerance) && (nrm <= dist_thr);  // :478-487
        // the count_nonzero-of-indices quirk (:228-230): the correspondence at list position 0
        // never counts.  In the fused pipeline position 0 is the winner's first sample point
        // (lib/ransac/ransac.py:76 returns the samples first).
        const bool has_row = quirk_row && quirk_row[0] >= 0;  // a row of -1 = no sample table
        counts = ok && !(has_row ? (gi == quirk_row[0]) : (i == 0));
    }
    const unsigned lane = threadIdx.x & 31u;
    const unsigned okb = __ballot_sync(0xffffffffu, ok);
    if (i < m && p == 0) pass[i] = (uint8_t)((okb >> (lane & ~3u)) & 0xfu);
    const unsigned cb = __ballot_sync(0xffffffffu, counts);
    if (lane < 4) {
        const int n = __popc(cb & (0x11111111u << lane));  // lanes with pose == lane
        if (n) atomicAdd(&s_cnt[lane], n);
    }
    __syncthreads();
    if (threadIdx.x < 4 && s_cnt[threadIdx.x])
        atomicAdd((unsigned long long*)&poses->counts[threadIdx.x], (unsigned long long)s_cnt[threadIdx.x]);
}

// np.argmax(num_good_correspondences) (:237): first maximum.
__global__ void k_vote(PoseSet* poses) {
    if (threadIdx.x || blockIdx.x) return;
    int b = 0;
    for (int p = 1; p < 4; ++p)
        if (poses->counts[p] > poses->counts[b]) b = p;
    poses->best = b;
}

// triangulate_points (lib/epipolar/triangulation.py:42-62), pixel coordinates.
// When use_vote != 0 the second camera is K [R|t] of the voted pose (built on the device)
// and only correspondences passing that pose are triangulated (others get NaN) — the fused
// tail of the single-pair pipeline (apps/sfm.py:133-186).
__global__ void __launch_bounds__(128)
k_triangulate(const double* __restrict__ xa, const double* __restrict__ ya, const double* __restrict__ xb,
              const double* __restrict__ yb, long long stride, long long m, const long long* __restrict__ m_dev,
              const double* __restrict__ P1g, const double* __restrict__ P2g, const double* __restrict__ Kmat,
              const PoseSet* __restrict__ poses, const uint8_t* __restrict__ pass, int use_vote,
              double* __restrict__ X, const long long* __restrict__ gather = nullptr) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (m_dev) m = *m_dev;
    if (i >= m) return;
    const long long gi = gather ? gather[i] : i;
    double P1[12], P2[12];
    if (use_vote) {
        const int b = poses->best;
        const double* R = poses->R[b];
        const double* t = poses->t[b];
#pragma unroll
        for (int r = 0; r < 3; ++r)
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                // K_ext @ Tmat: K[r,:] . [R|t][:,c]
                const double a0 = (c < 3) ? R[c] : t[0], a1 = (c < 3) ? R[3 + c] : t[1],
                             a2 = (c < 3) ? R[6 + c] : t[2];
                P2[4 * r + c] = fma(Kmat[3 * r + 2], a2, fma(Kmat[3 * r + 1], a1, Kmat[3 * r] * a0));
                P1[4 * r + c] = (c < 3) ? Kmat[3 * r + c] : 0.0;
            }
        if (!((pass[i] >> b) & 1u)) {
            const double nan = __longlong_as_double(0x7ff8000000000000LL);
            X[3 * i] = nan; X[3 * i + 1] = nan; X[3 * i + 2] = nan;
            return;
        }
    } else {
#pragma unroll
        for (int k = 0; k < 12; ++k) { P1[k] = P1g[k]; P2[k] = P2g[k]; }
    }
    double Xo[3];
    dlt_triangulate(xa[gi * stride], ya[gi * stride], xb[gi * stride], yb[gi * stride], P1, P2, Xo);
    X[3 * i] = Xo[0];
    X[3 * i + 1] = Xo[1];
    X[3 * i + 2] = Xo[2];
}

}  // namespace sfm
