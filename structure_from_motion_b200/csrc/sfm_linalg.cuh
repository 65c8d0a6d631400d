// Small dense fp64 linear algebra in registers / shared memory for the fitter, the pose
// decomposition and the DLT.  Jacobi methods throughout: they are branch-light, need no
// pivoting, and deliver eigen/singular vectors to near machine precision — the reference
// uses LAPACK dgeev / dgesdd for the same jobs (lib/epipolar/eight_point.py:253,410,440,
// lib/epipolar/triangulation.py:34).
#pragma once
#include <cuda_runtime.h>

namespace sfm {

// Symmetric Schur rotation annihilating a_pq:  returns (c, s, t) with t = tan(theta).
__device__ __forceinline__ void sym_schur(double app, double aqq, double apq, double& c, double& s,
                                          double& t) {
    const double theta = (aqq - app) / (2.0 * apq);
    const double ath = fabs(theta);
    // t = sgn(theta) / (|theta| + sqrt(theta^2 + 1)); guard the overflow of theta^2
    double tt = (ath > 1e150) ? 0.5 / ath : 1.0 / (ath + sqrt(fma(theta, theta, 1.0)));
    t = (theta < 0.0) ? -tt : tt;
    c = rsqrt(fma(t, t, 1.0));
    s = t * c;
}

// The same rotation with three long-latency operations instead of five (one sqrt, one divide, one rsqrt): with
// d = aqq - app and g = 2 apq,  t = sgn(d g) |g| / (|d| + sqrt(d^2 + g^2)).  For inputs whose squares neither
// overflow nor underflow (the fitter's 3x3 estimate has unit Frobenius norm).
__device__ __forceinline__ void sym_schur_short(double app, double aqq, double apq, double& c, double& s) {
    const double d = aqq - app, g = apq + apq;
    const double den = fabs(d) + sqrt(fma(d, d, g * g));
    const double tt = den > 0.0 ? fabs(g) / den : 1.0;
    const double t = ((d < 0.0) != (g < 0.0)) ? -tt : tt;
    c = rsqrt(fma(t, t, 1.0));
    s = t * c;
}

// Cyclic Jacobi eigen-decomposition of a symmetric NxN matrix held in registers
// (a: row-major full matrix, only the upper triangle is referenced; v receives the
// eigenvectors as columns).  Fully unrolled: indices are compile-time constants.
template <int N>
__device__ __forceinline__ void jacobi_eig_reg(double (&a)[N * N], double (&v)[N * N], int max_sweeps) {
#pragma unroll
    for (int i = 0; i < N; ++i)
#pragma unroll
        for (int j = 0; j < N; ++j) v[i * N + j] = (i == j) ? 1.0 : 0.0;
    for (int sweep = 0; sweep < max_sweeps; ++sweep) {
        double off = 0.0, diag = 0.0;
#pragma unroll
        for (int i = 0; i < N; ++i) {
            diag += fabs(a[i * N + i]);
#pragma unroll
            for (int j = i + 1; j < N; ++j) off += fabs(a[i * N + j]);
        }
        if (off <= 1e-300 || off <= 1e-22 * diag) break;
#pragma unroll
        for (int p = 0; p < N - 1; ++p) {
#pragma unroll
            for (int q = p + 1; q < N; ++q) {
                const double apq = a[p * N + q];
                const double g = 1e3 * fabs(apq);
                // negligible against both diagonal entries: just drop it
                if (fabs(a[p * N + p]) + g == fabs(a[p * N + p]) &&
                    fabs(a[q * N + q]) + g == fabs(a[q * N + q])) {
                    a[p * N + q] = 0.0;
                    continue;
                }
                if (apq == 0.0) continue;
                double c, s, t;
                sym_schur(a[p * N + p], a[q * N + q], apq, c, s, t);
                a[p * N + p] = fma(-t, apq, a[p * N + p]);
                a[q * N + q] = fma(t, apq, a[q * N + q]);
                a[p * N + q] = 0.0;
#pragma unroll
                for (int k = 0; k < N; ++k) {
                    if (k == p || k == q) continue;
                    // upper-triangle addresses of (k,p) and (k,q)
                    const int ikp = (k < p) ? k * N + p : p * N + k;
                    const int ikq = (k < q) ? k * N + q : q * N + k;
                    const double x = a[ikp], y = a[ikq];
                    a[ikp] = fma(c, x, -s * y);
                    a[ikq] = fma(s, x, c * y);
                }
#pragma unroll
                for (int k = 0; k < N; ++k) {
                    const double x = v[k * N + p], y = v[k * N + q];
                    v[k * N + p] = fma(c, x, -s * y);
                    v[k * N + q] = fma(s, x, c * y);
                }
            }
        }
    }
}

// One-sided (Hestenes) Jacobi SVD of an MxN matrix in registers, M >= N not required.
// On exit the columns of g are U*Sigma (mutually orthogonal), v holds the right singular
// vectors as columns.  No A^T A is formed, so small singular directions keep full relative
// accuracy — this is what the DLT null vector (triangulation.py:34-35) needs.
// SHORT: the fitter's variant - sym_schur_short and an orthogonality test without the square root
// (gamma^2 <= 1e-30 alpha beta, i.e. cos(angle between the columns) <= 1e-15).
template <int M, int N, bool SHORT = false>
__device__ __forceinline__ void jacobi_svd_onesided(double (&g)[M * N], double (&v)[N * N],
                                                    int max_sweeps) {
#pragma unroll
    for (int i = 0; i < N; ++i)
#pragma unroll
        for (int j = 0; j < N; ++j) v[i * N + j] = (i == j) ? 1.0 : 0.0;
    for (int sweep = 0; sweep < max_sweeps; ++sweep) {
        bool rotated = false;
#pragma unroll
        for (int p = 0; p < N - 1; ++p) {
#pragma unroll
            for (int q = p + 1; q < N; ++q) {
                double alpha = 0.0, beta = 0.0, gamma = 0.0;
#pragma unroll
                for (int k = 0; k < M; ++k) {
                    alpha = fma(g[k * N + p], g[k * N + p], alpha);
                    beta = fma(g[k * N + q], g[k * N + q], beta);
                    gamma = fma(g[k * N + p], g[k * N + q], gamma);
                }
                // columns already orthogonal to working precision?
                if (SHORT ? (gamma * gamma <= 1e-30 * (alpha * beta))
                          : (gamma == 0.0 || fabs(gamma) <= 1e-16 * sqrt(alpha * beta))) continue;
                rotated = true;
                double c, s, t;
                if (SHORT) sym_schur_short(alpha, beta, gamma, c, s);
                else sym_schur(alpha, beta, gamma, c, s, t);
#pragma unroll
                for (int k = 0; k < M; ++k) {
                    const double x = g[k * N + p], y = g[k * N + q];
                    g[k * N + p] = fma(c, x, -s * y);
                    g[k * N + q] = fma(s, x, c * y);
                }
#pragma unroll
                for (int k = 0; k < N; ++k) {
                    const double x = v[k * N + p], y = v[k * N + q];
                    v[k * N + p] = fma(c, x, -s * y);
                    v[k * N + q] = fma(s, x, c * y);
                }
            }
        }
        if (!rotated) break;
    }
}

__device__ __forceinline__ void mat3_mul(const double (&a)[9], const double (&b)[9], double (&c)[9]) {
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j)
            c[i * 3 + j] = fma(a[i * 3 + 2], b[6 + j], fma(a[i * 3 + 1], b[3 + j], a[i * 3] * b[j]));
}

__device__ __forceinline__ void mat3_transpose(const double (&a)[9], double (&t)[9]) {
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) t[i * 3 + j] = a[j * 3 + i];
}

}  // namespace sfm
