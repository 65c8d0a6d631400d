/* sfm_b200 — C ABI of the B200-native two-view geometry hot path.
 *
 * Drop-in acceleration boundary for Bazs/structure_from_motion's
 *   RANSAC essential-matrix estimation -> cheirality vote -> linear triangulation.
 * The reference is pure Python and has no FFI of its own; the boundary it exposes is the
 * set of Python callables in lib/ransac/ransac.py and lib/epipolar/*.py.  Each entry point
 * below names the reference callable (file:line) whose work it replaces; the Python
 * mirror in structure_from_motion_b200/ (re-exported under the reference's import paths
 * in lib/) binds these with ctypes — see INTEGRATION.md.
 *
 * Conventions
 *   - plain C types only; all pointers are HOST memory unless the parameter name ends in _d
 *   - every function returns 0 on success, a negative sfm_status otherwise;
 *     sfm_last_error() gives the message for the calling thread
 *   - a context owns one CUDA device, one stream and grow-only device buffers; calls on
 *     one context are serialised by the caller (the reference is single-threaded)
 *   - "correspondence" = one matched feature pair (xa, ya) <-> (xb, yb) in pixel
 *     coordinates; "hypothesis" = one RANSAC iteration (one minimal sample of 8
 *     correspondences and the essential matrix fitted to it)
 */
#ifndef SFM_B200_H
#define SFM_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct sfm_ctx sfm_ctx;

enum sfm_status {
    SFM_OK = 0,
    SFM_ERR_ARG = -1,      /* bad argument                                  */
    SFM_ERR_CUDA = -2,     /* CUDA runtime error (see sfm_last_error)        */
    SFM_ERR_STATE = -3,    /* call order violated (e.g. score before fit)   */
    SFM_ERR_NO_DEVICE = -4 /* no usable CUDA device                         */
};

/* lib/ransac/ransac.py:12-16  ErrorAggregationMethod */
enum sfm_aggregation { SFM_AGG_SUM = 0, SFM_AGG_SQUARE = 1, SFM_AGG_MEAN = 2, SFM_AGG_RMS = 3 };
/* lib/ransac/ransac.py:83 selects by minimum aggregated error (default).  Extras behind non-default values
 * (SURVEY.md 8(f) N4): max-inliers; MSAC = minimum of sum_i min(sed_i, threshold) over all correspondences
 * (best.err / err[] then hold that cost). */
enum sfm_selection { SFM_SELECT_MIN_ERROR = 0, SFM_SELECT_MAX_INLIERS = 1, SFM_SELECT_MSAC = 2 };
/* scoring kernel variant — all of them give bit-identical results (every inlier decision and
 * every summed value comes from the exact fp64 scorer):
 *   AUTO      (default) a pilot on the device measures the survivor rate of the one-sided screen and the
 *             scoring kernel takes SCREEN below 9 % survivors, FULL above
 *   SCREEN    fp64 one-sided screen (11 FP64 slots per evaluation) + exact re-check of survivors
 *   FULL      fp64 two-sided division-free decision (21 slots) + exact re-check
 *   SCREEN32  the screen evaluated in fp32 as a pre-filter (rigorous guard band, ~2 % more
 *             survivors) + exact fp64 re-check; reported separately from the fp64 headline */
enum sfm_score_variant { SFM_SCORE_SCREEN = 0, SFM_SCORE_FULL = 1, SFM_SCORE_SCREEN32 = 2, SFM_SCORE_AUTO = 3 };

/* ---- context ------------------------------------------------------------------------ */
int sfm_version(void);
const char *sfm_last_error(void);
int sfm_device_count(void);
int sfm_create(int device, sfm_ctx **out);
int sfm_destroy(sfm_ctx *ctx);
/* Use an externally owned cudaStream_t (e.g. torch's current stream); NULL restores the
 * context's own (non-blocking) stream. */
int sfm_set_stream(sfm_ctx *ctx, void *cuda_stream);
/* Use the legacy default stream (cudaStreamLegacy) - what a framework's "default stream" handle 0 means; needed
 * when the context's work has to be ordered with that stream's work (NCCL collectives, timing events). */
int sfm_use_default_stream(sfm_ctx *ctx);
int sfm_synchronize(sfm_ctx *ctx);
/* hyps_per_thread: essential matrices per thread (1, 2, 4; fp32 pre-filter: 2, 4, 8); group:
 * correspondences per vote (hyps_per_thread * group <= 32 tests per lane and vote); 0 keeps the
 * current shape, or selects the variant's default when switching to/from SCREEN32. */
int sfm_set_score_variant(sfm_ctx *ctx, int variant, int hyps_per_thread, int group);
/* Pinned host memory for the caller's buffers (so that H2D/D2H copies are true async DMA). */
int sfm_host_alloc(uint64_t bytes, void **out);
int sfm_host_free(void *p);

/* ---- sampling ----------------------------------------------------------------------- */
/* lib/ransac/ransac.py:62-63 — `random.shuffle(data); data[:8]`, cumulative over iterations,
 * restated for CPython's MT19937 + Fisher-Yates (random.py shuffle/_randbelow_with_getrandbits).
 * state: the 625 words of random.getstate()[1] (624 MT words + position), updated in place so
 * the caller can random.setstate() it back.  table: int32[h][8].  If perm_at >= 0,
 * perm_out[n] receives the full permutation after iteration perm_at (the reference returns
 * the inliers in that order, ransac.py:70-76).  Host-only; needs no context. */
int sfm_mt_shuffle_table(uint32_t *state625, int64_t n, int64_t h, int32_t *table, int64_t perm_at,
                         int32_t *perm_out);
/* The same iterations continued from a caller-held permutation (perm_inout[n], updated in place; table may be NULL):
 * lets the caller keep (state, permutation) snapshots every few iterations and replay only a short stretch to
 * recover the permutation of the winning iteration. */
int sfm_mt_shuffle_resume(uint32_t *state625, int64_t n, int64_t h, int32_t *table, int32_t *perm_inout);
/* sfm_mt_shuffle_table from the identity permutation that also records, before every iteration k * stride, the
 * generator state (snap_states[k][625]) and the permutation (snap_perms[k][n]) - so that the permutation or the state
 * after ANY iteration is recovered by replaying at most `stride` iterations (sfm_mt_shuffle_resume) instead of the
 * whole run: the reference returns the inliers in that iteration's permutation order (ransac.py:70-76) and, when a
 * degenerate sample aborts the run, leaves the generator where that iteration left it. */
int sfm_mt_shuffle_snapshots(uint32_t *state625, int64_t n, int64_t h, int32_t *table, int64_t stride,
                             uint32_t *snap_states, int32_t *snap_perms);
/* Upload a sample table (h x 8 indices into the correspondences). */
int sfm_set_table(sfm_ctx *ctx, const int32_t *table, int64_t h);
/* Device sampler (Philox4x32-10 keyed by seed/stream/global hypothesis index): 8 distinct
 * uniform indices per hypothesis; hyp_offset shifts the global index for hypothesis-sharded
 * runs so that the union over ranks equals the single-GPU table. */
int sfm_sample_device(sfm_ctx *ctx, uint64_t seed, uint64_t stream, int64_t hyp_offset, int64_t h);
/* Read back rows [first, first + h) of the current table. */
int sfm_get_table(sfm_ctx *ctx, int32_t *table, int64_t first, int64_t h);

/* ---- correspondences ---------------------------------------------------------------- */
/* lib/epipolar/eight_point.py:127-133 (to_normalized_image_coords), applied once on the
 * device.  xa/ya/xb/yb: pixel coordinates, element i at ptr[i*stride] (stride 1 = four
 * separate arrays; stride 2 = two interleaved [n][2] arrays with ya = xa + 1).
 * K: row-major 3x3; only fx, fy, cx, cy are used, as in the reference. */
int sfm_upload_pairs(sfm_ctx *ctx, const double *xa, const double *ya, const double *xb,
                     const double *yb, int64_t stride, int64_t n, const double *K);
/* The same without the trailing synchronisation: the copies out of the caller's arrays are only ENQUEUED, so the
 * arrays must stay valid and unchanged until the next synchronising call on this context (sfm_two_view_fetch,
 * sfm_synchronize, ...).  Lets the upload of one estimate overlap the kernels of another context. */
int sfm_upload_pairs_async(sfm_ctx *ctx, const double *xa, const double *ya, const double *xb,
                           const double *yb, int64_t stride, int64_t n, const double *K);
/* Same, inputs already in device memory. */
int sfm_upload_pairs_d(sfm_ctx *ctx, const double *xa_d, const double *ya_d, const double *xb_d,
                       const double *yb_d, int64_t stride, int64_t n, const double *K);
/* Read back the K-normalised records: out[n][4] = (xa, ya, xb, yb). */
int sfm_get_normalised(sfm_ctx *ctx, double *out, int64_t n);

/* ---- model fitting ------------------------------------------------------------------- */
/* lib/epipolar/epipolar_ransac.py:28-42 (eight_point_model_fitter) ->
 * lib/epipolar/eight_point.py:99-170 for every row of the current sample table.
 * E_out: double[h][9] or NULL; valid_out: uint8[h] or NULL (0 = the reference would raise
 * EightPointCalculationError, eight_point.py:414-421); eig_out: double[h][9] or NULL
 * (eigenvalues of Y^T Y, diagnostics). */
int sfm_fit(sfm_ctx *ctx, double *E_out, uint8_t *valid_out, double *eig_out);
/* The fitted (or uploaded) models [first, first+count) back to the host: E_out double[count][9], valid_out uint8[count]
 * (either may be NULL). */
int sfm_get_models(sfm_ctx *ctx, int64_t first, int64_t count, double *E_out, uint8_t *valid_out);
/* Replace the fitted models by caller-supplied ones (scorer-only parity tests). */
int sfm_set_models(sfm_ctx *ctx, const double *E, const uint8_t *valid /* or NULL */, int64_t h);

/* ---- scoring + selection ------------------------------------------------------------- */
/* lib/ransac/ransac.py:66-86 + :96-108 with lib/epipolar/epipolar_ransac.py:18-25 /
 * lib/epipolar/sed.py:7-30 as the scorer.  use_table != 0 applies the sample rule of
 * ransac.py:63-64,76 with the current table.  Outputs (each may be NULL):
 * count_extra int32[h] (-1 for invalid hypotheses), S1/S2 double[h] (sum of sed / sed^2 over
 * samples + extra inliers), err double[h] (+inf when not a candidate). */
int sfm_score(sfm_ctx *ctx, double threshold, double min_extra, int aggregation, int selection,
              int use_table, int64_t idx_offset, int32_t *count_extra, double *S1, double *S2,
              double *err);

typedef struct sfm_best {
    double err;            /* aggregated error of the winner                          */
    int64_t index;         /* global hypothesis index, -1 if no candidate              */
    int32_t count_extra;   /* inliers beyond the 8 samples                             */
    int32_t reserved;
    int64_t num_invalid;   /* hypotheses the reference would have raised on           */
    int64_t first_invalid; /* lowest such index, -1 if none                            */
    double E[9];           /* the winning model (row-major), E[8] == 1                 */
    int32_t sample[8];     /* the winner's minimal sample (ransac.py:63), -1 without a table or a winner */
} sfm_best;
/* Winner of the last sfm_score (ransac.py:83: minimum error, earliest iteration on ties). */
int sfm_get_best(sfm_ctx *ctx, sfm_best *out);
/* Tell the context which hypothesis won (hypothesis-sharded runs: after the cross-rank
 * merge).  local_index < 0 = the winner lives on another rank; E then supplies the model. */
int sfm_set_winner(sfm_ctx *ctx, int64_t local_index, const double *E);
/* SURVEY.md H1 (ransac.py:83,96-108: strict < on errors summed in LIST order): the local indices of the hypotheses
 * of the last score whose error is <= best * (1 + rel_tol), unordered, at most cap of them; *count is their total
 * number.  The Python mirror re-evaluates such near-ties in the reference's own summation order. */
int sfm_near_ties(sfm_ctx *ctx, double rel_tol, int64_t cap, int64_t *idx_out, int64_t *count);
/* How many hypotheses of the last score went through K3's exact double-double rescore (their inliers were more than
 * 2^31 times tighter than the threshold, so the 84-bit fixed-point sums of K2 would have lost precision). */
int sfm_get_rescored(sfm_ctx *ctx, int64_t *count);
/* Inlier mask (sed <= threshold) and SED value of every correspondence under the current
 * winner.  mask uint8[n], sed double[n]; either may be NULL. */
int sfm_inlier_mask(sfm_ctx *ctx, double threshold, uint8_t *mask, double *sed);

/* One call for the whole estimate (lib/epipolar/epipolar_ransac.py:45-70): fit -> score ->
 * select -> mask of the winner, no host round trip in between. */
int sfm_ransac_essential(sfm_ctx *ctx, double threshold, double min_extra, int aggregation,
                         int selection, sfm_best *best, uint8_t *mask, double *sed);

/* ---- pose + triangulation ------------------------------------------------------------ */
typedef struct sfm_poses {
    double R[4][9];      /* (R1,t) (R1,-t) (R2,t) (R2,-t)  — eight_point.py:210-212 order */
    double t[4][3];
    double sv[3];        /* singular values of E, descending                             */
    int64_t counts[4];   /* cheirality votes (index-0 quirk of eight_point.py:228-230)   */
    int32_t best;        /* np.argmax(counts) (eight_point.py:237), -1 before the vote    */
    int32_t reserved;
} sfm_poses;
/* lib/epipolar/eight_point.py:245-280 (_recover_all_r_t) on the device. */
int sfm_decompose_essential(sfm_ctx *ctx, const double *E, sfm_poses *out);
/* lib/epipolar/eight_point.py:181-242 (_recover_r_t): decomposition + 4-pose cheirality vote
 * (eight_point.py:449-488) over m correspondences given in K-NORMALISED coordinates
 * (element i at ptr[i*stride]).  pass4: uint8[m], bit p set = passes pose p. */
int sfm_recover_pose(sfm_ctx *ctx, const double *E, const double *xa, const double *ya,
                     const double *xb, const double *yb, int64_t stride, int64_t m,
                     double distance_threshold, sfm_poses *out, uint8_t *pass4);
/* The same with PIXEL coordinates: to_normalized_image_coords (eight_point.py:127-133) is applied on the device with
 * K (double[9], row-major) first — recover_r_t_from_e (eight_point.py:65-96). */
int sfm_recover_pose_pixels(sfm_ctx *ctx, const double *E, const double *K, const double *xa, const double *ya,
                            const double *xb, const double *yb, int64_t stride, int64_t m,
                            double distance_threshold, sfm_poses *out, uint8_t *pass4);
/* lib/epipolar/triangulation.py:9-62: DLT triangulation of m correspondences (pixel
 * coordinates) with 3x4 row-major camera matrices P1, P2.  X: double[m][3]. */
int sfm_triangulate(sfm_ctx *ctx, const double *P1, const double *P2, const double *xa,
                    const double *ya, const double *xb, const double *yb, int64_t stride, int64_t m,
                    double *X);
/* Fused tail of the pipeline for the current winner (apps/sfm.py:110-186): inliers of the
 * winner -> decomposition -> cheirality vote -> triangulation of the passing inliers with
 * P1 = K[I|0], P2 = K[R|t], all on the device.  inlier_idx int64[cap] receives the indices
 * (ascending) of the winner's inliers, pass uint8[cap] the vote result per inlier, X
 * double[cap][3] the points (NaN rows where pass == 0). */
int sfm_pose_and_triangulate(sfm_ctx *ctx, double threshold, double distance_threshold,
                             sfm_poses *poses, int64_t cap, int64_t *num_inliers,
                             int64_t *inlier_idx, uint8_t *pass, double *X);
/* sfm_ransac_essential + sfm_pose_and_triangulate in one call (apps/sfm.py:110-186): the winner never leaves the
 * device between selection and the tail, the host synchronises once for the fixed-size results.  mask / sed
 * (optional) are "sed <= threshold" / the distances under the winner, as sfm_ransac_essential returns them. */
int sfm_two_view(sfm_ctx *ctx, double threshold, double min_extra, int aggregation, int selection,
                 double distance_threshold, sfm_best *best, sfm_poses *poses, int64_t cap, int64_t *num_inliers,
                 int64_t *inlier_idx, uint8_t *pass, double *X, uint8_t *mask, double *sed);
/* sfm_two_view in two halves: _async enqueues everything (device sampler or uploaded table -> fit -> score -> select ->
 * tail, plus the copies of mask / sed into the caller's arrays when given) and returns at once; _fetch synchronises
 * and returns the results.  With two contexts on one GPU the host buffers of estimate s+1 upload while estimate s is
 * being scored (structure_from_motion_b200.two_view.TwoViewStream). */
int sfm_two_view_async(sfm_ctx *ctx, double threshold, double min_extra, int aggregation, int selection,
                       double distance_threshold, uint8_t *mask, double *sed);
int sfm_two_view_fetch(sfm_ctx *ctx, sfm_best *best, sfm_poses *poses, int64_t cap, int64_t *num_inliers,
                       int64_t *inlier_idx, uint8_t *pass, double *X);

/* ---- batched image pairs (pair-sharded workloads) ------------------------------------ */
/* P independent pairs; pair p owns correspondences [offsets[p], offsets[p+1]) of the
 * concatenated arrays, its own K (Ks[p][9]) and h hypotheses drawn by the device sampler
 * (stream = pair_id0 + p).  Outputs per pair: E[p][9], best_index[p] (-1 = none),
 * best_err[p], count_extra[p], num_invalid[p].  Degenerate samples are skipped. */
int sfm_batch_ransac(sfm_ctx *ctx, const double *xa, const double *ya, const double *xb,
                     const double *yb, int64_t stride, const int64_t *offsets, int64_t npairs,
                     const double *Ks, int64_t h, uint64_t seed, uint64_t pair_id0, double threshold,
                     double min_extra, int aggregation, int selection, double *E, int64_t *best_index,
                     double *best_err, int32_t *count_extra, int64_t *num_invalid);

/* The same batch carried through the rest of the path (apps/sfm.py:118-186 per pair): for every pair's winner the
 * inlier list (sed <= threshold plus the 8 sample points, ransac.py:70-76, ascending pair-relative indices), the four
 * candidate poses with the cheirality vote (lib/epipolar/eight_point.py:65-96, 181-280, 449-488; poses[p].best = the
 * voted candidate, -2 = the pair has no model) and the triangulated points of the inliers that pass the voted pose
 * (lib/epipolar/triangulation.py:42-62, pixel coordinates with the pair's K; NaN for the others).  Per-inlier results
 * are packed densely: pair p owns [inlier_offsets[p], inlier_offsets[p+1]) of inlier_idx int32[cap], pass uint8[cap]
 * (bit q = passes candidate q) and X double[cap][3]; cap >= offsets[npairs] always suffices. */
int sfm_batch_two_view(sfm_ctx *ctx, const double *xa, const double *ya, const double *xb, const double *yb,
                       int64_t stride, const int64_t *offsets, int64_t npairs, const double *Ks, int64_t h,
                       uint64_t seed, uint64_t pair_id0, double threshold, double min_extra, int aggregation,
                       int selection, double distance_threshold, double *E, int64_t *best_index, double *best_err,
                       int32_t *count_extra, int64_t *num_invalid, sfm_poses *poses, int64_t *inlier_offsets,
                       int64_t cap, int32_t *inlier_idx, uint8_t *pass, double *X);

/* ---- hypothesis-sharded estimates without a host round trip ---------- */
/* One rank of a run whose hypotheses are split over `world` GPUs (correspondences replicated):
 *   1. sfm_score_async  : fit + score + select of this rank's hypotheses, nothing synchronised; *record_dev is the
 *                         DEVICE address of this rank's SFM_RECORD_BYTES-byte selection record;
 *   2. the caller all-gathers the records of all ranks into one device buffer on the context's stream
 *      (ncclAllGather / torch.distributed.all_gather_into_tensor - the path's only collective, SFM_RECORD_BYTES per rank);
 *      a record = {err f64, index i64, count i32, pad, num_invalid i64, first_invalid i64, E f64[9], sample i32[8]}:
 *      the winner's model AND its minimal sample travel with it, so every rank finishes with identical results;
 *   3. sfm_sharded_tail : merges the records on the device with the reference's rule (ransac.py:83: smallest error,
 *                         earliest GLOBAL iteration = rank * hyps_per_rank + local index on ties) and enqueues the
 *                         inlier mask, pose vote and triangulation of the global winner;
 *   4. sfm_sharded_fetch: the single synchronisation; best->index is the global index, *owner the rank that fitted
 *                         the winner (-1: no model anywhere). */
#define SFM_RECORD_BYTES 144
int sfm_score_async(sfm_ctx *ctx, double threshold, double min_extra, int aggregation, int selection,
                    void **record_dev);
int sfm_sharded_tail(sfm_ctx *ctx, const void *gathered_records_dev, int world, int rank, int64_t hyps_per_rank,
                     int selection, double threshold, double distance_threshold);
int sfm_sharded_fetch(sfm_ctx *ctx, sfm_best *best, int32_t *owner, sfm_poses *poses, int64_t cap,
                      int64_t *num_inliers, int64_t *inlier_idx, uint8_t *pass, double *X);

/* The same without any host framework: the collective behind the C ABI (SURVEY.md 8(b) sfm_nccl_init /
 * sfm_ransac_E_sharded).  libnccl.so.2 is resolved at run time (the copy already loaded in the process, else the
 * system's).  Rank 0 calls sfm_nccl_unique_id and ships the SFM_NCCL_ID_BYTES bytes to its peers by any means; every
 * rank calls sfm_nccl_init (collective).  sfm_two_view_sharded then enqueues one complete estimate - this rank's
 * hypotheses [rank * hyps_per_rank, (rank+1) * hyps_per_rank) of the device sampler, ONE ncclAllGather of the
 * selection records on the context's stream, merge kernel, tail - and sfm_sharded_fetch returns, on every rank, what
 * one GPU returns for the union of the hypotheses (ransac.py:83 applied across ranks). */
#define SFM_NCCL_ID_BYTES 128
int sfm_nccl_unique_id(void *id_out);
int sfm_nccl_init(sfm_ctx *ctx, int rank, int nranks, const void *unique_id);
int sfm_nccl_destroy(sfm_ctx *ctx);
int sfm_two_view_sharded(sfm_ctx *ctx, uint64_t seed, int64_t hyps_per_rank, double threshold, double min_extra,
                         int aggregation, int selection, double distance_threshold);

/* ---- the stage in front of the hot path: brute-force matcher ------ */
/* lib/feature_matching/matching.py:36-118 match_brute_force with ncc.py:7-54 (score_kind 0, score
 * in [0,2], 2.0 when a window leaves the image) or ssd.py:7-36 (score_kind 1, +inf outside) as the
 * score function.  window <= 15; like util.py:21-27 the patch spans center +- int(window/2), so an even
 * window covers window + 1 pixels.  Images: row-major rows x cols, image_dtype 0 = uint8 (numpy's uint8 arithmetic
 * of ssd.py is reproduced), 1 = float64.  feats_*: double[n][2] = (x, y) (lib/common/feature.py).
 * validation: bit 0 RATIO_TEST (heap[0]/heap[1] <= ratio_threshold, matching.py:84-97 - heap[1] of
 * the reference's heapq, not the second smallest score), bit 1 CROSSCHECK (matching.py:100-118).
 * Outputs per feature of image A: best_b int32[na], best_score double[na], keep uint8[na] (1 = the
 * match survives the validations; the reference returns exactly those, in order); scores (optional)
 * double[na][nb] = the full score matrix. */
int sfm_match_brute_force(sfm_ctx *ctx, const void *image_a, const void *image_b, int image_dtype,
                          int64_t rows, int64_t cols, const double *feats_a, int64_t na,
                          const double *feats_b, int64_t nb, int score_kind, int window, int validation,
                          double ratio_threshold, int32_t *best_b, double *best_score, uint8_t *keep,
                          double *scores);

/* The same selection + validations on a caller-supplied score matrix double[na][nb] (the scores of
 * an arbitrary Python score_function, matching.py:55-65). */
int sfm_match_from_scores(sfm_ctx *ctx, const double *scores, int64_t na, int64_t nb, int validation,
                          double ratio_threshold, int32_t *best_b, double *best_score, uint8_t *keep);

/* ---- the first stage of apps/sfm.py: Harris corners -------------------- */
/* lib/common/correlate.py:4-39 cross_correlate: zero "same" border, odd square kernel double[ksize][ksize],
 * out double[rows][cols].  image_dtype as above. */
int sfm_cross_correlate(sfm_ctx *ctx, const void *image, int image_dtype, int64_t rows, int64_t cols,
                        const double *kernel, int ksize, double *out);
/* Shape of the cornerness image: rows/cols - int(np.around(block_size / 2)) (harris_detector.py:66-72). */
int sfm_harris_output_shape(int64_t rows, int64_t cols, int block_size, int64_t *out_rows, int64_t *out_cols);
/* lib/harris/harris_detector.py:11-55 detect_harris_corners: Sobel responses, block sums, det - k trace^2,
 * negatives clamped to 0, the reference's in-place (scan-order dependent) 3x3 non-maximum suppression, the
 * num_corners highest non-zero values in descending order (exact ties: descending flat index).
 * xy double[num_corners][2] = (x, y) = (col, row) + block_size/2; score double[num_corners];
 * *num_found <= num_corners; cornerness (optional) double[out_rows][out_cols] after suppression;
 * nms_sweeps (optional) = parallel sweeps the suppression needed. */
int sfm_harris_corners(sfm_ctx *ctx, const void *image, int image_dtype, int64_t rows, int64_t cols,
                       int64_t num_corners, int block_size, double k, double *xy, double *score,
                       int64_t *num_found, double *cornerness, int32_t *nms_sweeps);

/* ---- measurement --------------------------------------------------------------------- */
/* Per-stage device times (CUDA events on the context's stream) of the most recent
 * pipeline call: ms[0]=upload+normalise ms[1]=sample ms[2]=fit ms[3]=score ms[4]=finalise+select
 * ms[5]=mask ms[6]=pose ms[7]=triangulate (a stage is reported once, then reads 0 until it runs
 * again).  launches = kernels launched by this library since the context was created. */
int sfm_enable_timing(sfm_ctx *ctx, int on);
int sfm_get_timing(sfm_ctx *ctx, float ms[8], int64_t *launches);
/* FP64 FMA throughput microbenchmark (denominator of the FP64 roofline): returns achieved
 * DFMA/s over the whole device. */
int sfm_measure_fp64_peak(sfm_ctx *ctx, double *dfma_per_s);

#ifdef __cplusplus
}
#endif
#endif /* SFM_B200_H */
