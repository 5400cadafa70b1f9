/*
 * surf_b200.h -- C ABI of the B200-native SURF engine (libsurf_b200.so, sm_100a CUDA kernels).
 *
 * Drop-in boundary for the OpenCV calls made by stropitek/SurfFeatures:
 *   cv::SurfFeatureDetector(minHessian).detect(img, kps, mask)
 *        SurfFeatures/Logic/vtkSlicerSurfFeaturesLogic.cxx:1384-1385
 *   cv::SurfDescriptorExtractor().compute(img, kps, desc)
 *        SurfFeatures/Logic/vtkSlicerSurfFeaturesLogic.cxx:1388-1389
 *   descriptorMatcher->knnMatch(queryDescriptors, matches, k)
 *        SurfFeatures/Logic/vtkSlicerSurfFeaturesLogic.cxx:640   (+ clear/add at :1402-1403)
 * Output layouts are those of cv::KeyPoint (28 B), a continuous CV_32F N x D cv::Mat, and
 * cv::DMatch (16 B): holders at SurfFeatures/Logic/vtkSlicerSurfFeaturesLogic.h:211-237,258-260.
 *
 * Conventions: every function returns a SURF_* status (0 = OK); no C++ exception and no exit()
 * crosses this boundary; there is NO CPU fallback -- without an sm_100 device surf_create fails.
 * An engine is bound to one device and one CUDA stream and is not thread-safe: use one engine per
 * host thread / per GPU.  Plain pointers and sizes only; no torch or OpenCV types.
 */
#ifndef SURF_B200_H
#define SURF_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SURF_OK 0
#define SURF_ERR_BAD_ARG (-1)     /* null pointer, non-positive size, invalid parameter */
#define SURF_ERR_UNSUPPORTED (-2) /* frame or batch larger than the engine was created for, ... */
#define SURF_ERR_CAPACITY (-3)    /* output would not fit; surf_required_capacity() tells the need */
#define SURF_ERR_CUDA (-4)
#define SURF_ERR_NO_DEVICE (-5)   /* no sm_100 device: there is no CPU path */
#define SURF_ERR_NCCL (-6)

#define SURF_MEM_HOST 0
#define SURF_MEM_DEVICE 1

/* cv::KeyPoint field order: Point2f pt; float size, angle, response; int octave, class_id. */
typedef struct surf_keypoint {
    float x, y;
    float size;
    float angle;
    float response;
    int32_t octave;
    int32_t class_id;
} surf_keypoint;

/* cv::DMatch field order. */
typedef struct surf_dmatch {
    int32_t query_idx;
    int32_t train_idx;
    int32_t img_idx;
    float distance;
} surf_dmatch;

/* The five public fields of cv::SURF (nonfree/features2d.hpp). */
typedef struct surf_params {
    double hessian_threshold;
    int32_t n_octaves;
    int32_t n_octave_layers;
    int32_t extended; /* 0: 64-d, 1: 128-d */
    int32_t upright;  /* 1: skip orientation, angle = 270 */
} surf_params;

typedef struct surf_engine surf_engine;

/* Per-stage device times of the batch collected last, milliseconds (CUDA events on the engine's streams).  d2h is
 * the fetch of the per-frame counters; total ends there (result copies run on a separate copy stream). */
typedef struct surf_timings {
    float h2d, integral, hessian, nms, sort, describe, pack, d2h, total;
    int32_t launches; /* kernels launched by the last call */
} surf_timings;

/* ---- lifetime ---- */

/* Creates an engine on `device` with workspace for `max_batch` frames of up to max_width x
 * max_height and `max_keypoints` raw keypoints per frame.  Mirrors `explicit SURF(double
 * hessianThreshold, int nOctaves, int nOctaveLayers, bool extended, bool upright)`.  max_width is limited to 7 224
 * pixels (SURF_ERR_UNSUPPORTED beyond: a band of the integral-image kernel holds whole rows in shared memory). */
int surf_create(const surf_params *params, int device, int max_width, int max_height, int max_batch,
                int max_keypoints, surf_engine **out);
void surf_destroy(surf_engine *e);
int surf_set_params(surf_engine *e, const surf_params *params);
int surf_get_params(const surf_engine *e, surf_params *params);
const char *surf_last_error(const surf_engine *e);
/* After SURF_ERR_CAPACITY: the per-frame (or total) count that would have been needed. */
long surf_required_capacity(const surf_engine *e);
int surf_descriptor_size(const surf_engine *e); /* SURF::descriptorSize(): 64 or 128 */
int surf_get_timings(const surf_engine *e, surf_timings *t);
/* cudaStream_t of the engine, as void* (for callers that time with their own CUDA events). */
void *surf_stream(const surf_engine *e);

/* ---- single frame: the cv::Feature2D calls ---- */

/* FeatureDetector::detect(image, keypoints, mask).  mask may be NULL (non-zero = keep).
 * Host pointers.  Writes up to `capacity` keypoints (oriented, sorted, deleted ones removed). */
int surf_detect(surf_engine *e, const uint8_t *image, int width, int height, size_t pitch, const uint8_t *mask,
                size_t mask_pitch, surf_keypoint *keypoints, int capacity, int *n_out);

/* DescriptorExtractor::compute(image, keypoints, descriptors); keypoints are in-out and may shrink. */
int surf_compute(surf_engine *e, const uint8_t *image, int width, int height, size_t pitch,
                 surf_keypoint *keypoints, int n_in, float *descriptors, int *n_out);

/* SURF::operator()(img, mask, keypoints, descriptors, useProvidedKeypoints=false). */
int surf_detect_and_compute(surf_engine *e, const uint8_t *image, int width, int height, size_t pitch,
                            const uint8_t *mask, size_t mask_pitch, surf_keypoint *keypoints, float *descriptors,
                            int capacity, int *n_out);

/* ---- batches: the frame loop of computeNext (vtkSlicerSurfFeaturesLogic.cxx:1392-1406) ---- */

/* B frames of identical size, frame b at images + b*frame_stride.  Outputs are PACKED in frame
 * order: frame b owns entries [offsets[b], offsets[b+1]) of keypoints / descriptor rows;
 * offsets has batch+1 entries (host memory always).  descriptors may be NULL (detect only).
 * in_mem / out_mem say where images(+mask) and keypoints/descriptors live (SURF_MEM_*).
 * `capacity` is the total number of keypoint slots in the output arrays. */
int surf_detect_and_compute_batch(surf_engine *e, const uint8_t *images, int in_mem, int batch, int width,
                                  int height, size_t pitch, size_t frame_stride, const uint8_t *mask,
                                  size_t mask_pitch, surf_keypoint *keypoints, float *descriptors, int out_mem,
                                  long capacity, int32_t *offsets);

/* Asynchronous halves of the call above: submit enqueues upload + all kernels; collect waits for the batch and
 * copies the packed results out.  TWO batches may be in flight: surf_submit_batch(i+1) may precede
 * surf_collect_batch(i) (collect always takes the oldest submitted batch), so the GPU never idles while the host
 * reads counters and enqueues copies; a third submit without a collect is SURF_ERR_BAD_ARG.
 * Sizing the outputs: a collect whose arrays are too small returns SURF_ERR_CAPACITY with offsets[] filled and leaves
 * the batch collectable; with no batch in flight, a further collect repeats the copy of the batch collected last (a
 * counts-only call with NULL arrays first, then the real one). */
int surf_submit_batch(surf_engine *e, const uint8_t *images, int in_mem, int batch, int width, int height,
                      size_t pitch, size_t frame_stride, const uint8_t *mask, size_t mask_pitch, int want_descriptors);
int surf_collect_batch(surf_engine *e, surf_keypoint *keypoints, float *descriptors, int out_mem, long capacity,
                       int32_t *offsets);
/* Optional: starts the host->device copy of the NEXT batch's frames on its own stream, into a second frame
 * buffer, while the batch in flight computes.  A following surf_submit_batch with the same host pointer and
 * geometry (in_mem = SURF_MEM_HOST) uses those frames instead of uploading again; anything else discards
 * them.  The host frames must stay unchanged until that submit's kernels have started. */
int surf_prefetch_batch(surf_engine *e, const uint8_t *images, int batch, int width, int height, size_t pitch,
                        size_t frame_stride);
/* Optional, stronger than surf_prefetch_batch: uploads the NEXT batch's frames (host input) and runs stages 1-3 on
 * them (integral, Hessian, maxima, sort) on a separate stream and a second set of buffers, while the submitted
 * batch is in its descriptor stage; the two stages stress different units of the SM and overlap well.  A following
 * surf_submit_batch with the same arguments (pointers, geometry, mask) continues with stage 4; anything else discards
 * the look-ahead.  Frames and mask must stay unchanged until that submit's results are collected. */
int surf_lookahead_batch(surf_engine *e, const uint8_t *images, int in_mem, int batch, int width, int height, size_t pitch,
                         size_t frame_stride, const uint8_t *mask, size_t mask_pitch);
/* Device-resident results of the batch collected last (valid until the second submit after that batch's own):
 * packed keypoints, packed descriptors, device int32 offsets[batch+1].  Before any collect: of the batch
 * submitted last (surf_sync first if the host needs them complete). */
int surf_device_results(surf_engine *e, const surf_keypoint **keypoints, const float **descriptors,
                        const int32_t **offsets);
int surf_sync(surf_engine *e);
/* Asynchronous-output mode: surf_collect_batch and surf_match_prev enqueue their device->host copies
 * on a second stream and return without waiting for them, so the copies overlap the next batch's
 * kernels.  The host buffers are valid after surf_wait_outputs(); alternate two sets of buffers. */
int surf_set_async_outputs(surf_engine *e, int enable);
int surf_wait_outputs(surf_engine *e);

/* ---- stage-level exports (parity checks of the per-stage tolerances) ---- */

/* cv::integral(img, sum, CV_32S): sum is (height+1) x (width+1) int32, dense, host. */
int surf_integral(surf_engine *e, const uint8_t *image, int width, int height, size_t pitch, int clamp01,
                  int32_t *sum);
/* det of pyramid layer (octave, layer) for one frame: (height>>octave) x (width>>octave) fp32, host. */
int surf_hessian_layer(surf_engine *e, const uint8_t *image, int width, int height, size_t pitch, int octave,
                       int layer, float *det, int *rows, int *cols);
/* Interpolated, sorted maxima before orientation (angle = -1). */
int surf_raw_keypoints(surf_engine *e, const uint8_t *image, int width, int height, size_t pitch,
                       const uint8_t *mask, size_t mask_pitch, surf_keypoint *keypoints, int capacity, int *n_out);
/* Orientation + descriptor of given keypoints WITHOUT removing deleted ones (size = -1 marks
 * them); patches (optional) receives the 21 x 21 8-bit window of each keypoint. */
int surf_describe_keypoints(surf_engine *e, const uint8_t *image, int width, int height, size_t pitch,
                            surf_keypoint *keypoints, int n, float *descriptors, uint8_t *patches);

/* ---- stage 5: BFMatcher(NORM_L2).knnMatch ---- */

/* k nearest train rows for every query row (k <= 4), ascending distance, ties -> lower train_idx.
 * out has n_query*k records (img_idx = 0).  mem says where query/train/out live. */
int surf_match_knn(surf_engine *e, const float *query, int n_query, const float *train, int n_train, int dim,
                   int k, surf_dmatch *out, int mem);
/* knnMatch(k=2) + ratio test d1 < ratio*d2; good[i] = 1 if query i passes.  out as above (k=2). */
int surf_match_ratio(surf_engine *e, const float *query, int n_query, const float *train, int n_train, int dim,
                     float ratio, surf_dmatch *out, uint8_t *good, int mem);

/* Live-stream matching (BASELINE config 2): every frame of the batch collected last
 * on this engine is matched (k = 2 + ratio test) against its predecessor -- the first frame against
 * the last frame of the previous batch (no matches on the very first call or after
 * surf_stream_reset).  Results are packed like the keypoints: 2 records per keypoint, row
 * 2*(offsets[b]+i) for keypoint i of frame b, query_idx / train_idx local to their frames;
 * good[offsets[b]+i] = ratio test.  out / good may be NULL to leave the results on the device
 * (surf_device_matches).  One batched launch sequence, no host synchronisation inside. */
int surf_match_prev(surf_engine *e, float ratio, surf_dmatch *out, uint8_t *good, int out_mem, long capacity);
int surf_stream_reset(surf_engine *e);
int surf_device_matches(surf_engine *e, const surf_dmatch **matches, const uint8_t **good);
/* 0: tensor-core candidate search + exact rescoring (default); 1: exact CUDA-core kernel only.
 * Both return identical results. */
int surf_set_match_mode(surf_engine *e, int exact_only);
/* Small batches (<= max_batch frames, default 16; 0 keeps the current value): the launch sequence of one submit is
 * captured into a CUDA graph the second time the same buffers and sizes come round and replayed afterwards (the live,
 * one-frame-per-tick use of vtkSlicerSurfFeaturesLogic.cxx:1392-1406).  Same kernels, same arguments: identical results.
 * On by default; surf_graph_replays counts the submits served by a replay. */
int surf_set_graphs(surf_engine *e, int enable, int max_batch);
int surf_graph_replays(const surf_engine *e, long *replays);
/* Number of queries of the last tensor-core match call whose candidate set could not be proven and
 * that were redone by the exact kernel (diagnostic; synchronises the engine stream). */
int surf_match_redo_count(surf_engine *e, int *count);

/* ---- row-sharded keyframe database helpers (multi-GPU, config 5) ---- */

/* Merges per-rank top-k candidate lists: in is n_ranks x n_query x k records (train_idx already
 * global), out is n_query x k.  Device pointers; runs on the engine stream. */
int surf_merge_topk(surf_engine *e, const surf_dmatch *in, int n_ranks, int n_query, int k, surf_dmatch *out);

/* ---- "next" row f2: the train database behind DescriptorMatcher::add / clear / knnMatch ----
 * Reference: computeNext() hands all descriptor Mats of a sweep to the matcher,
 * vtkSlicerSurfFeaturesLogic.cxx:1402-1403 (clear + add(vector<Mat>)), and
 * matchDescriptorsKNN :635-643 queries it; DMatch.imgIdx / trainIdx follow the add() order.
 * The database lives on the device: fp32 rows (exact rescoring), their BF16 hi/lo tensor-core
 * form (prepared once per add), |row|^2, the keypoints of the rows (for the Hough vote below) and
 * the image row offsets.  Exact brute-force L2 like surf_match_knn; ties -> lower image, lower row. */
typedef struct surf_db surf_db;

int surf_db_create(surf_engine *e, int dim, long max_rows, int max_images, surf_db **out);
void surf_db_destroy(surf_db *db);
int surf_db_clear(surf_db *db);                                   /* DescriptorMatcher::clear() */
/* DescriptorMatcher::add(): n_images more train images; image i owns rows
 * [offsets[i], offsets[i+1]) of descriptors (offsets: host array of n_images+1, offsets[0] may be
 * non-zero).  keypoints (optional, same rows) are kept for surf_correspond.  mem tells where
 * descriptors / keypoints live (e.g. SURF_MEM_DEVICE with the pointers of surf_device_results). */
int surf_db_add(surf_db *db, const float *descriptors, const surf_keypoint *keypoints, const int32_t *offsets,
                int n_images, int mem);
int surf_db_size(const surf_db *db, long *rows, int *images);
/* knnMatch(query, k) against the whole database (k <= 4): out[n_query*k], img_idx = train image,
 * train_idx = row within that image (-1 / -1 when the database has fewer than k rows). */
int surf_db_knn(surf_db *db, const float *query, int n_query, int k, surf_dmatch *out, int mem);

/* ---- "next" row f3: inter-slice correspondence, computeNextInterSliceCorrespondence :1434-1571 ----
 * For every query image: knnMatch(k=3) against the train database and knnMatch(k=1) against the
 * bogus database (:1457-1458); drop a query keypoint when d_train / d_bogus > bogus_ratio (0.95,
 * :1463-1480); Hough pose-consistency vote over the surviving matches (houghTransform :241-345 with
 * the poll booth of :86-164, reference point = mask centroid); per-train-image vote count and the
 * best image (first maximum, :1485-1512).
 * matches of query image i start at matches[q_offsets[i]] (results[i].n_matches of them, in query
 * order); entries the Hough vote rejects are kept with query/train/img index -1 exactly as the
 * reference leaves them in afterHoughMatches. */
typedef struct surf_vote_params {
    float bogus_ratio; /* 0.95 */
    int32_t ref_x, ref_y; /* xcentroid, ycentroid of the cropped mask (computeCentroid :839-860) */
} surf_vote_params;

typedef struct surf_vote_result {
    int32_t n_matches;   /* matches that passed the bogus filter (size of afterHoughMatches[i]) */
    int32_t hough_count; /* votes of the winning poll booth (iMatchCount) */
    int32_t best_image;  /* bestMatches[i]: train image with most surviving matches, -1 if no train image */
    int32_t best_count;  /* bestMatchesCount[i] */
} surf_vote_result;

int surf_correspond(surf_db *train, surf_db *bogus, const float *query_descriptors,
                    const surf_keypoint *query_keypoints, const int32_t *q_offsets, int n_query_images,
                    const surf_vote_params *params, surf_dmatch *matches, surf_vote_result *results, int mem);

/* ---- "next" row f1: MHA sequence ingest + crop + mask, readAndComputeFeaturesOnMhaFile :1337-1379 ----
 * Host-side reader of the PLUS MetaImage sequences the reference loads (readImageDimensions_mha :444-476,
 * readImageTransforms_mha :478-541, readImages_mha :543-595, read_mha :598-632) and the driver that pushes
 * a whole file through the engine in batches (computeNext :1392-1406), reading the next batch from disk into
 * pinned memory while the current one is on the GPU.  The crop (cropData :745-752) is applied during the
 * host->device copy: only the rectangle's bytes are uploaded. */
typedef struct surf_mha surf_mha;

typedef struct surf_mha_info {
    int32_t width, height, frames_in_file; /* DimSize */
    int32_t first_frame, last_frame;       /* requested range after the reference's clamping */
    int32_t n_frames;                      /* frames kept: in range and tracking status OK */
    int32_t n_transforms;                  /* transforms / names kept by the same filter */
} surf_mha_info;

/* By default a frame with status INVALID is not read and its bytes are NOT skipped (what readImages_mha
 * does: every later frame then comes from one frame earlier in the file).  This flag skips them. */
#define SURF_MHA_SKIP_INVALID_BYTES 1

int surf_mha_open(const char *path, int first_frame, int last_frame, int flags, surf_mha **out, surf_mha_info *info);
void surf_mha_close(surf_mha *m);
const char *surf_mha_last_error(const surf_mha *m);
/* n_transforms x 12 floats (rows of the 3x4 tracker transform), kept frames only */
int surf_mha_transforms(const surf_mha *m, float *transforms, int capacity);
/* validity flags of frames first_frame..last_frame (what read_mha leaves in transformsValidity) */
int surf_mha_validity(const surf_mha *m, uint8_t *valid, int capacity);
/* "<directory of the file><frame field name>.png" of kept frame i */
int surf_mha_frame_name(const surf_mha *m, int i, char *buf, int capacity);
/* kept frames [start, start+count) -> dst (width*height bytes each, frame_stride apart) */
int surf_mha_read(surf_mha *m, int start, int count, uint8_t *dst, size_t frame_stride);
/* cropData's cv::Rect(x, y, w, h) for crop_ratios = {x0, x1, y0, y1} (fp32 arithmetic, truncation) */
int surf_crop_rect(int width, int height, const float crop_ratios[4], int32_t rect[4]);
/* computeCentroid :839-860 of a (cropped) mask */
int surf_mask_centroid(const uint8_t *mask, int width, int height, size_t pitch, int32_t *x, int32_t *y);
/* Every kept frame of the file: crop (NULL = whole frame), detect with the full-size mask's same rectangle
 * (NULL = no mask), describe.  Results go to the optional host arrays (packed, offsets[n_frames+1]) and /
 * or are appended to db on the device (matcher->add).  The engine's max_batch sets the batch size. */
int surf_sweep_mha(surf_engine *e, surf_mha *m, const float crop_ratios[4], const uint8_t *mask, size_t mask_pitch,
                   surf_db *db, surf_keypoint *keypoints, float *descriptors, int32_t *offsets, long capacity,
                   int32_t *n_frames_out);

/* ---- "next" row f4: plane fit by RANSAC and pose estimate (ransac :2085-2188, updateMatchNodeRansac :1655-1851) ----
 * rand(): the reference draws from the process-global rand(), never seeded (= srand(1)).  surf_rand is that
 * generator (glibc TYPE_3) as an explicit state; create it with seed 1 and pass it to the calls below in the
 * reference's call order to reproduce its draws. */
typedef struct surf_rand surf_rand;
int surf_rand_create(uint32_t seed, surf_rand **out);
void surf_rand_destroy(surf_rand *r);
int surf_rand_next(surf_rand *r); /* rand() */
/* trainPoints / queryPoints :1672-1703: one pair per match whose imgIdx is not -1 (afterHoughMatches keeps rejected
 * entries with -1).  train_transforms: 16 floats per train image (trainImagesTransform), image_to_probe: 16 doubles,
 * both row-major.  Outputs: n_points x 3 doubles each (capacity n_matches). */
int surf_match_points(const surf_dmatch *matches, int n_matches, const surf_keypoint *query_keypoints,
                      const surf_keypoint *train_keypoints, const int32_t *train_offsets, int n_train_images,
                      const float *train_transforms, const double image_to_probe[16], double *train_points,
                      double *query_points, int32_t *n_points);
/* ransac(points, inliersIdx, planePointsIdx) :2085-2188 with threshold = ransacMargin (default 1.0, :1002) and 1000
 * iterations in the reference.  The hypotheses are scored on the device.  inliers: capacity n; order as the reference
 * leaves it (fitted inliers in point order, then the three plane points).  n < 3 is SURF_ERR_BAD_ARG (reference: returns 1). */
int surf_ransac_plane(surf_engine *e, surf_rand *rng, const double *points, int n, float margin, int iterations,
                      int32_t *inliers, int32_t *n_inliers, int32_t plane_idx[3], double *best_error, int32_t *n_outliers);
/* The estimate matrix of updateMatchNodeRansac :1714-1808 (row-major 4x4, IJK to RAS) and the mean match distance
 * :1811-1823.  Host arithmetic (about 2 MFLOP of dependent fp64 with libm calls). */
int surf_estimate_pose(surf_rand *rng, const double *query_points, const double *train_points, int n, const int32_t *inliers,
                       int n_inliers, const int32_t plane_idx[3], double estimate[16], double *mean_match_distance);
/* Mean plane-to-plane distance :1826-1846 over every step-th non-zero mask pixel (step 4 in the reference). */
int surf_mean_plane_distance(const uint8_t *mask, int rows, int cols, size_t pitch, int step, const double ground_truth[16],
                             const double estimate[16], double *out);

/* ---- timing marks: CUDA events on the engine's own stream (what bench.py times with) ---- */
#define SURF_MAX_MARKS 8
int surf_mark(surf_engine *e, int slot);                                  /* records mark `slot` (0..7) now */
int surf_mark_elapsed(surf_engine *e, int from_slot, int to_slot, float *ms); /* waits for `to_slot` */
/* Device time and kernel launches of the last surf_match_prev (waits for it). */
int surf_match_prev_time(surf_engine *e, float *ms, int *launches);

/* ---- multi-GPU: one process per GPU, NCCL on device buffers --------------------
 * Frames are independent (computeNext touches only images[index], vtkSlicerSurfFeaturesLogic.cxx:1392-1406),
 * so a sweep is sharded by contiguous frame blocks with no collective on the compute path; NCCL carries only
 *   - the gather of the variable-length keypoint / descriptor blocks to one rank, in frame order
 *     (surf_sweep_sharded; BASELINE config 4), and
 *   - the all-gather of per-rank top-k records of a row-sharded train database, merged with ties going to
 *     the lower global row as BFMatcher's single scan does (surf_db_knn_sharded; BASELINE config 5; the
 *     matcher calls being distributed are matchDescriptorsKNN :635-643 / :1457-1458).
 * libnccl.so.2 is resolved at run time (the copy already loaded in the process, e.g. torch's, else the system
 * one); a missing library is SURF_ERR_NCCL from surf_comm_*, never a load failure of libsurf_b200.so. */
typedef struct surf_comm surf_comm;
#define SURF_COMM_ID_BYTES 128
/* ncclGetUniqueId: call on one rank and hand the 128 bytes to every rank (any transport). */
int surf_comm_unique_id(void *id);
/* ncclCommInitRank on the engine's device; collective over the n_ranks processes. */
int surf_comm_init(surf_engine *e, const void *id, int rank, int n_ranks, surf_comm **out);
void surf_comm_destroy(surf_comm *c);
int surf_comm_info(const surf_comm *c, int *rank, int *n_ranks, int *nccl_version);
/* Contiguous block [lo, hi) of `rank`: the first n_items % n_ranks ranks get one item more. */
int surf_shard_range(long n_items, int rank, int n_ranks, long *lo, long *hi);

typedef struct surf_sweep_stats {
    float total_ms;        /* device time from the first kernel to the last copy into the output arrays */
    float exchange_ms;     /* device time of the NCCL operations (they overlap the compute of later chunks) */
    float assemble_ms;     /* destination rank: frame-order assembly of the gathered blocks */
    int64_t bytes_sent;    /* payload this rank sent */
    int64_t bytes_received;/* payload this rank received (destination rank) */
    int32_t chunks;        /* batches per rank (every rank runs the same number) */
    int32_t launches;      /* kernels launched by this rank's engine */
} surf_sweep_stats;

/* computeNext over a whole sweep, sharded by frame: this rank holds frames [lo, hi) = surf_shard_range(
 * n_frames_total, rank, n_ranks) at images_local (frame_stride apart); it runs them in batches of the engine's
 * max_batch and sends every batch's packed keypoints / descriptors to dst_rank while the next batch computes
 * (per-frame counts by ncclAllGather, payload by grouped ncclSend / ncclRecv straight from the device results).
 * On dst_rank keypoints / descriptors (out_mem host or device, `capacity` rows) receive ALL frames in frame
 * order and offsets[n_frames_total + 1] (host) the row offsets; other ranks may pass NULL for the three.
 * comm == NULL runs the same code on one GPU.  Collective: every rank of the communicator must call it. */
int surf_sweep_sharded(surf_engine *e, surf_comm *comm, const uint8_t *images_local, int in_mem, long n_frames_total,
                       int width, int height, size_t pitch, size_t frame_stride, const uint8_t *mask, size_t mask_pitch,
                       int want_descriptors, int dst_rank, surf_keypoint *keypoints, float *descriptors, int out_mem,
                       long capacity, int64_t *offsets, surf_sweep_stats *stats);

/* knnMatch(query, k) against a train database whose rows are sharded over the ranks (rank r holds the r-th
 * contiguous block in its own surf_db).  query_root >= 0: that rank's query rows are broadcast first (the other
 * ranks pass a buffer of the same size that is overwritten when mem is device, or ignored when host);
 * query_root < 0: every rank already passes the same rows.  Every rank receives the merged result: train_idx
 * is the GLOBAL row (rank-order concatenation), img_idx 0, ties -> lower global row.  Collective. */
int surf_db_knn_sharded(surf_db *db_shard, surf_comm *comm, const float *query, int n_query, int k, int query_root,
                        surf_dmatch *out, int mem);
/* Device time of the last surf_db_knn_sharded on this rank: whole call, and the NCCL part of it. */
int surf_comm_last_times(const surf_comm *c, float *total_ms, float *nccl_ms);

const char *surf_version(void);

#ifdef __cplusplus
}
#endif
#endif
