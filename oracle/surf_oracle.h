/*
 * surf_oracle.h -- CPU restatement of the OpenCV 2.4.5 SURF path.  TEST INFRASTRUCTURE ONLY.
 *
 * This is the parity oracle for the B200 engine.  Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may load it.  The product library
 * (surffeatures_b200/libsurf_b200.so) never links, loads or calls anything in oracle/.
 *
 * PARITY UNPINNED BY THE REFERENCE: stropitek/SurfFeatures holds no tests, fixtures or golden
 * vectors (SurfFeatures/Testing/Cxx/CMakeLists.txt:4-6,16) and the arithmetic lives in OpenCV 2.4.5
 * `nonfree` (SuperBuild/External_OpenCV.cmake:6-7,56), which is not vendored and cannot be fetched.
 * The restatement below follows the published algorithm of modules/nonfree/src/surf.cpp and its
 * helpers; sub-steps are pinned against cv2 4.13 in tests/ (integral,
 * INTER_AREA resize, fastAtan2, getGaussianKernel, BFMatcher).
 *
 * Call sites in the reference this path serves:
 *   SurfFeatures/Logic/vtkSlicerSurfFeaturesLogic.cxx:1381-1390  (detect + compute)
 *   SurfFeatures/Logic/vtkSlicerSurfFeaturesLogic.cxx:635-643    (knnMatch)
 */
#ifndef SURF_ORACLE_H
#define SURF_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Same field order and size (28 B) as cv::KeyPoint. */
typedef struct {
    float x, y;
    float size;
    float angle;
    float response;
    int32_t octave;
    int32_t class_id;
} orc_keypoint;

/* Same field order and size (16 B) as cv::DMatch. */
typedef struct {
    int32_t query_idx, train_idx, img_idx;
    float distance;
} orc_dmatch;

typedef struct {
    double hessian_threshold;
    int32_t n_octaves;
    int32_t n_octave_layers;
    int32_t extended;
    int32_t upright;
} orc_params;

/* Named reading choices as process-wide run-time switches; all zero / null = the 2.4.5 reading the CUDA
 * path implements.  See surf_oracle.c. */
void orc_set_switches(int angle_pre244, int libm_sincos, const float *desc_gauss20);

/* ---- sub-steps (exported so tests can pin each one) ---- */

/* sum is (H+1) x (W+1) int32, dense.  clamp01 != 0 applies min(pixel,1) first (mask integral). */
void orc_integral(const uint8_t *img, int H, int W, int pitch, int32_t *sum, int clamp01);

/* det / trace: (H/step) x (W/step) float, caller zero-fills.  Returns 0 if the layer is skipped. */
int orc_hessian_layer(const int32_t *sum, int H, int W, int size, int step, float *det, float *trace);

/* n-tap normalised Gaussian as fp32, the way getGaussianKernel(n, sigma, CV_32F) builds it. */
void orc_gaussian_kernel(int n, double sigma, float *out);

float orc_fast_atan2(float y, float x);

/* INTER_AREA down-scale of an n x n 8-bit window to 21 x 21 (n >= 21). */
int orc_resize_area_21(const uint8_t *win, int n, uint8_t *patch);

/* ---- the path ---- */

/* Fast-Hessian maxima after interpolation, sorted, before orientation.  Returns the count found
 * (may exceed cap; only cap entries are written). mask may be NULL. */
int orc_raw_keypoints(const uint8_t *img, int H, int W, int pitch,
                      const uint8_t *mask, int mask_pitch,
                      const orc_params *p, orc_keypoint *kps, int cap);

/* Orientation (+ descriptor when desc != NULL) for n given keypoints; marks deletions by size=-1.
 * patches (optional) receives the 21x21 window of every keypoint (n*441 bytes). */
void orc_describe(const uint8_t *img, int H, int W, int pitch,
                  const orc_params *p, orc_keypoint *kps, int n, float *desc, uint8_t *patches);

/* cv::FeatureDetector::detect: maxima + orientation + removal of deleted keypoints. */
int orc_detect(const uint8_t *img, int H, int W, int pitch,
               const uint8_t *mask, int mask_pitch,
               const orc_params *p, orc_keypoint *kps, int cap);

/* cv::DescriptorExtractor::compute: in-out keypoints (may shrink), descriptors n_out x D. */
int orc_compute(const uint8_t *img, int H, int W, int pitch,
                const orc_params *p, orc_keypoint *kps, int n, float *desc);

/* SURF::operator() with useProvidedKeypoints=false: both in one pass. */
int orc_detect_and_compute(const uint8_t *img, int H, int W, int pitch,
                           const uint8_t *mask, int mask_pitch,
                           const orc_params *p, orc_keypoint *kps, int cap, float *desc);

/* Frame-parallel batch (OpenMP over frames, n_threads<=0 = all). counts[b] receives per-frame N;
 * frame b writes kps + b*cap and desc + b*cap*D.  Returns total keypoints. */
long orc_detect_and_compute_batch(const uint8_t *imgs, int B, int H, int W,
                                  const orc_params *p, orc_keypoint *kps, int cap, float *desc,
                                  int32_t *counts, int n_threads);

/* BFMatcher(NORM_L2).knnMatch: k nearest train rows per query row, ascending, ties -> lower index.
 * out is nq*k records; img_idx is filled with 0. */
void orc_knn_match(const float *q, int nq, const float *t, int nt, int d, int k,
                   orc_dmatch *out, int n_threads);

/* ---- "next" rows f2 / f3 (vtkSlicerSurfFeaturesLogic.cxx:73-345, :1434-1571, :839-860) ---- */
int orc_hough_transform(const orc_keypoint *qk, const orc_keypoint *tk, const int32_t *img_off, orc_dmatch *matches,
                        int n, int refx, int refy, int libm_float);
void orc_correspond(const orc_keypoint *qk, int nq, const orc_dmatch *m_train, const orc_dmatch *m_bogus,
                    const orc_keypoint *tk, const int32_t *img_off, int n_img, float bogus_ratio, int refx, int refy,
                    orc_dmatch *out, int32_t *result, int libm_float);
void orc_mask_centroid(const uint8_t *mask, int H, int W, int pitch, int *x, int *y);

int orc_max_threads(void);

#ifdef __cplusplus
}
#endif
#endif
