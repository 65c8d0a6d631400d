/* TEST INFRASTRUCTURE ONLY — plain-C restatement of the reference scorer.
 *
 * Restates lib/epipolar/sed.py:7-30 (calculate_symmetric_epipolar_distance) on
 * K-normalised coordinates (lib/epipolar/eight_point.py:127-133 applied beforehand, as
 * lib/epipolar/epipolar_ransac.py:21-22 does), with the exact evaluation order that
 * numpy 2.3 + OpenBLAS 0.3.30 use for the 3x3 products in sed.py:21-25 (SURVEY.md
 * Appendix B): each 3-term dot product is  fma(u1, v1, u0*v0) + u2*1.0.
 *
 * oracle_score_batch restates the candidate test and the error sums of
 * lib/ransac/ransac.py:66-82 for H hypotheses (sample points excluded from the count,
 * included unconditionally in the sums), threaded over hypotheses; it is the CPU
 * baseline that bench.py times.  Compile with -ffp-contract=off so that only the
 * explicit fma() calls fuse.
 *
 * Never linked into, or called by, the product library.
 */
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>

double oracle_sed_exact(const double *e, double xa, double ya, double xb, double yb) {
    /* b.T @ E  and  E.T @ b  (sed.py:21,25) */
    double lb0 = fma(yb, e[3], xb * e[0]) + e[6];
    double lb1 = fma(yb, e[4], xb * e[1]) + e[7];
    double lb2 = fma(yb, e[5], xb * e[2]) + e[8];
    /* E @ a  (sed.py:24) */
    double la0 = fma(e[0], xa, e[1] * ya) + e[2];
    double la1 = fma(e[3], xa, e[4] * ya) + e[5];
    /* (b.T @ E) @ a  (sed.py:21) */
    double r = fma(lb1, ya, lb0 * xa) + lb2;
    /* sed.py:27-29 */
    return (1.0 / (la0 * la0 + la1 * la1) + 1.0 / (lb0 * lb0 + lb1 * lb1)) * (r * r);
}

void oracle_sed_exact_many(const double *e, const double *xa, const double *ya, const double *xb,
                           const double *yb, int64_t n, double *out) {
    for (int64_t i = 0; i < n; ++i) out[i] = oracle_sed_exact(e, xa[i], ya[i], xb[i], yb[i]);
}

typedef struct {
    const double *E, *xa, *ya, *xb, *yb;
    const int32_t *table;
    const uint8_t *valid;
    int64_t n, h0, h1;
    double thr;
    int32_t *count_extra;
    double *s1, *s2;
} job_t;

static void *worker(void *p) {
    job_t *j = (job_t *)p;
    for (int64_t h = j->h0; h < j->h1; ++h) {
        if (j->valid && !j->valid[h]) {
            j->count_extra[h] = -1;
            j->s1[h] = j->s2[h] = 0.0;
            continue;
        }
        const double *e = j->E + 9 * h;
        int64_t cnt = 0;
        double s1 = 0.0, s2 = 0.0;
        for (int64_t i = 0; i < j->n; ++i) {
            double s = oracle_sed_exact(e, j->xa[i], j->ya[i], j->xb[i], j->yb[i]);
            if (s <= j->thr) { /* ransac.py:73  score <= inlier_threshold */
                ++cnt;
                s1 += s;
                s2 += s * s;
            }
        }
        if (j->table) { /* ransac.py:63-64,76: samples are not thresholded but always in the error */
            for (int k = 0; k < 8; ++k) {
                int64_t i = j->table[8 * h + k];
                double s = oracle_sed_exact(e, j->xa[i], j->ya[i], j->xb[i], j->yb[i]);
                if (s <= j->thr) {
                    --cnt;
                } else {
                    s1 += s;
                    s2 += s * s;
                }
            }
        }
        j->count_extra[h] = (int32_t)cnt;
        j->s1[h] = s1;
        j->s2[h] = s2;
    }
    return 0;
}

/* count_extra[h] = #{i not in sample_h : sed <= thr}; s1/s2 = sum of sed / sed^2 over
 * samples U extra inliers (ransac.py:70-82).  valid may be NULL. */
void oracle_score_batch(const double *E, const uint8_t *valid, const double *xa, const double *ya,
                        const double *xb, const double *yb, int64_t n, int64_t h,
                        const int32_t *table, double thr, int32_t *count_extra, double *s1,
                        double *s2, int nthreads) {
    if (nthreads < 1) nthreads = 1;
    if (nthreads > 256) nthreads = 256;
    pthread_t th[256];
    job_t jobs[256];
    int64_t per = (h + nthreads - 1) / nthreads;
    int used = 0;
    for (int t = 0; t < nthreads; ++t) {
        int64_t h0 = t * per, h1 = h0 + per > h ? h : h0 + per;
        if (h0 >= h1) break;
        jobs[t] = (job_t){E, xa, ya, xb, yb, table, valid, n, h0, h1, thr, count_extra, s1, s2};
        pthread_create(&th[t], 0, worker, &jobs[t]);
        ++used;
    }
    for (int t = 0; t < used; ++t) pthread_join(th[t], 0);
}
