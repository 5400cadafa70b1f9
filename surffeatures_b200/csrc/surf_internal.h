// surf_internal.h -- types shared by the host engine and the sm_100a kernels (not part of the ABI).
#pragma once

#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "surf_b200.h"

namespace surfb200 {

constexpr int kMaxLayersPerOctave = 8;  // nOctaveLayers + 2 <= 8
constexpr int kMaxOctaves = 8;
constexpr int kPatch = 21;              // PATCH_SZ + 1 (OpenCV nonfree/surf.cpp)
constexpr int kPatchPx = kPatch * kPatch;
constexpr int kPatchStride = 448;        // a 21 x 21 patch in a descriptor CTA's shared-memory ring (441 bytes, rounded up to 16)
constexpr int kOriSamples = 113;        // integer points with i^2 + j^2 <= 36
constexpr int kOriWindows = 72;         // 0, 5, ..., 355 degrees
constexpr int kMaxWindow = 768;         // largest descriptor window side handled (size 264 -> 739)
constexpr int kWinSmemSide = 128;       // windows up to 128 x 128 are staged in shared memory
constexpr int kSortChunk = 1024;

// One box of a resized Haar pattern: corner coordinates relative to the filter origin in the
// integral image and the signed fp32 weight 1/area  (resizeHaarPattern).
struct Box {
    int x1, y1, x2, y2;
    float w;
};

struct LayerPat {
    int size;    // filter side in pixels
    int m;       // (size/2)/step: first written sample row/column
    int ni, nj;  // written samples: 1 + (H - size)/step, 1 + (W - size)/step
    int skip;    // size > H or size > W: layer left at zero
    Box dxx[3], dyy[3], dxy[4];
    Box mask;    // (0,0,9,9) resized: mean of the mask integral under the filter
};

struct OctavePat {
    int step, rows, cols, lpo;  // sampling step 2^o, sample grid, layers in this octave
    LayerPat L[kMaxLayersPerOctave];
};

// Fused, TMA-staged Hessian + maxima kernel (octaves whose integral tile fits shared memory).
// samples per CTA incl. the 1-sample halo (useful: 30 x 14).  512 threads: two CTAs per SM, so one tile's TMA wait
// and maxima phase overlap the other's box filters (measured, Hessian stage of 64 frames: 64x16 1.90 ms,
// 32x16 1.61 ms, 32x8 1.62 ms, 64x8 1.78 ms, 32x32 1.78 ms)
constexpr int kTileW = 32, kTileH = 16;
constexpr int kTileThreads = kTileW * kTileH;

struct TileLayer {
    int dxx[8];   // word offsets of the 2 x 4 Dxx corners in the tile: row-major (y1: x0..x3, y2: x0..x3)
    int dyy[8];   // 4 x 2 Dyy corners: (y0: x1, x2), (y1: ...), ...
    int dxy[16];  // 4 x 4 Dxy corners, row-major
    float w[10];  // box weights: Dxx[3], Dyy[3], Dxy[4]
    int size, m, ni, nj, skip;
    Box mask;
};

struct TileOctave {
    int step, rows, cols, lpo;
    int m_max;               // largest layer margin (samples)
    int tile_p, tile_h;      // integral tile: pitch in words (multiple of 4) and rows
    int tile_bytes;
    TileLayer L[kMaxLayersPerOctave];
};

struct FrameGeo {
    int W, H;
    int SP;            // integral row pitch in 32-bit words (multiple of 4)
    size_t sum_frame;  // words per integral frame: (H + 1) * SP
};

struct ImageRef {
    const uint8_t *ptr;
    size_t pitch;         // bytes between rows
    size_t frame_stride;  // bytes between frames
};

// Descriptor windows are handled by size class (window side n = (int)(21 * size * 1.2 / 9)).  Classes 0-2
// (n <= 64 / 96 / 128) stage the rotated window's whole bounding box and the window itself in shared memory;
// the last class (n > 128) cuts the window into sub-windows of at most 96 x 96 taps, each with its own
// image tile, and streams their INTER_AREA row sums through shared memory (the window is never materialised).
constexpr int kNumClasses = 4;
constexpr int kClassMaxN0 = 64, kClassMaxN1 = 96, kClassMaxN2 = 128;  // last class: up to kMaxWindow

// d_misc word layout
constexpr int kMiscStatus = 1;
constexpr int kMiscClassCount = 2;   // [kNumClasses]
constexpr int kMiscClassHead = 8;    // [kNumClasses]
constexpr int kMiscWords = 16;

struct DescribeArgs {
    ImageRef img;
    int img_aligned4;     // base, pitch and frame stride are multiples of 4: word copies allowed
    int img_aligned16;    // ... multiples of 16: 16-byte copies allowed
    const uint32_t *sum;
    FrameGeo geo;
    surf_keypoint *kps;   // [B][cap], updated in place (angle, size = -1 on deletion)
    int cap;
    const int *offsets;   // exclusive prefix of the per-frame counts, B + 1 entries
    int B;
    float *desc;          // [B][cap][D] or nullptr (orientation only)
    uint8_t *patches;     // optional [B][cap][441]
    int extended, upright;
    int *misc;            // status word, class counters and queue heads (kMisc*)
    int *ndeleted;        // [B]
    int *deleted_slots;   // [B][cap]
    int *class_list;      // [kNumClasses][list_stride]: keypoint index b * cap + slot
    size_t list_stride;
    float2 *sincos;       // [B][cap]: (-sin, cos) of the dominant angle, written by k_orient
};


// ---- launchers (surf_kernels.cu) ----
int integral_band_rows(int W);
int integral_max_width();
size_t integral_bandsum_words(int W, int H, int B);
size_t integral_flag_words(int H, int B);
void launch_integral(const ImageRef &img, int B, const FrameGeo &geo, uint32_t *sum, uint32_t *bandsum, int *bandflags, int clamp01,
                     cudaStream_t st, int *launches);
void launch_hessian(const uint32_t *sum, const FrameGeo &geo, const OctavePat &pat, float *det, size_t det_frame, int B,
                    cudaStream_t st, int *launches);
void launch_nms(const float *det, size_t det_frame, const uint32_t *sum, const uint32_t *msum, const FrameGeo &geo,
                const OctavePat &pat, int octave, int n_mid, float thr, surf_keypoint *raw, int *counts, int cap, int B,
                cudaStream_t st, int *launches);
void launch_hessian_nms_tma(const CUtensorMap &tmap, const TileOctave &oct, const uint32_t *msum, const FrameGeo &geo,
                            int octave, float thr, surf_keypoint *raw, int *counts, int cap, int B, cudaStream_t st,
                            int *launches);
// Sorts every frame's raw keypoints (response desc, size desc, octave desc, y desc, x asc).
// Returns the buffer (a or b) that holds the sorted result.
surf_keypoint *launch_sort(surf_keypoint *a, surf_keypoint *b, const int *counts, int cap, int B, cudaStream_t st,
                           int *launches);
// offsets[b] = sum_{b' < b} (min(counts[b'], cap) - (ndel ? ndel[b'] : 0)), B + 1 entries.
void launch_offsets(const int *counts, const int *ndel, int cap, int B, int *offsets, cudaStream_t st, int *launches);
// imaps: TMA descriptors of the 8-bit frames, one per staged size class (box = that class's tile)
void launch_describe(const DescribeArgs &a, const CUtensorMap *imaps, int use_tma, cudaStream_t st, int *launches);
void describe_tile_box(int cls, int *pitch_bytes, int *rows);
void launch_pack(const surf_keypoint *kps, const float *desc, int D, int cap, int B, const int *in_offsets,
                 const int *ndel, const int *deleted_slots, const int *out_offsets, surf_keypoint *out_kps,
                 float *out_desc, cudaStream_t st, int *launches);
void launch_zero(int *p, int n_words, cudaStream_t st, int *launches);
void upload_tables();  // orientation sample weights + descriptor Gaussian -> __constant__ (per device)

// ---- tensor-core candidate search (surf_match_tc.cu) ----
constexpr int kCand = 8;  // approximate candidates kept per query and train split

struct TcCand {
    float d[kCand];  // approximate |t|^2 - 2 q.t (unsorted)
    int i[kCand];    // train row within its slot, -1 = empty
    float worst;     // largest kept value once the list is full, else INF
    float pad[3];
};

// Descriptor rows prepared for the tensor cores: BF16 hi / lo parts in the UMMA canonical K-major
// layout, one slot per frame (or per operand), slot_rows rows each (multiple of 256).
struct TcPool {
    uint16_t *hi, *lo;
    float *norm;      // [slots][slot_rows] |row|^2 (INF for padding rows)
    float *max_norm;  // [slots]
    int *count;       // [slots] live rows
    int slot_rows;
};

// Rows [off[0], off[1]) of `src` (or [row0, row0 + n) when off is null) go to `slot`, starting at row
// dst_row0 of the slot (append to a database slot); the slot then holds dst_row0 + n live rows.
struct TcSegment {
    const float *src;
    const int *off;
    int row0, n;
    int slot;
    int dst_row0;
};

// One knnMatch: queries = slot q_slot, train = slot t_slot; fp32 rows for the exact rescoring at
// qsrc + (q_off ? q_off[0] : q_row0) * D (same for train); results at out + (out_off ? ... ) * k.
struct TcProblem {
    int q_slot, t_slot;
    const float *qsrc, *tsrc;
    const int *q_off, *t_off, *out_off;
    int q_row0, t_row0, out_row0;
};

struct TcMatchArgs {
    TcPool pool;            // query slots (and train slots when tpool.hi is null)
    TcPool tpool;           // train slots living in another pool (a descriptor database), else zeroed
    const TcSegment *segs;  // device; rows to prepare into `pool` before matching (n_segs may be 0)
    int n_segs;
    const TcProblem *probs;  // device
    int n_probs;
    int n_splits;
    int max_q_rows;   // upper bound on queries per problem (multiple of 128)
    TcCand *cand;     // [n_probs][n_splits][max_q_rows]: the KC smallest approximate values of each train split
    int *counters;    // [0] work-queue head, [1] redo count
    int *redo_list;   // 2 ints per entry
    int dim, k;
    float ratio;
    surf_dmatch *out;
    uint8_t *good;    // optional
};

size_t tc_candidates_smem(int dim);
void launch_tc_match(const TcMatchArgs &a, cudaStream_t st, int *launches);
// prepares the rows of `segs` (device table) into `pool` without matching (database append)
void launch_tc_prep(const TcPool &pool, const TcSegment *segs, int n_segs, int max_rows, int dim, bool reset_norms,
                    cudaStream_t st, int *launches);
// rewrites global train rows of a multi-image database as (img_idx, row within the image): img_off[n_img + 1]
void launch_db_localize(surf_dmatch *m, long n, const int *img_off, int n_img, cudaStream_t st, int *launches);
// Builds the segment / problem tables of the frame-stream matcher on the device: frame i of the batch goes
// to ring slot (prev_slot + 1 + i) % slots and is matched against the frame before it.
void launch_tc_stream_tables(TcSegment *segs, TcProblem *probs, int B, int slots, int prev_slot, const float *desc,
                             const int *off, const float *prev_desc, const int *prev_off, cudaStream_t st, int *launches);
// copies rows [off[0], off[1]) of src (D floats each) to dst and writes {0, n} to dst_off (device-side count)
void launch_copy_rows(const float *src, const int *off, int dim, float *dst, int max_rows, int *dst_off, cudaStream_t st,
                      int *launches);

// ---- launchers (surf_match.cu) ----
size_t match_scratch_records(int nq, int nt, int k);  // partial top-k records launch_match_knn needs
void launch_match_knn(const float *q, int nq, const float *t, int nt, int dim, int k, surf_dmatch *out, int train_base,
                      surf_dmatch *scratch, cudaStream_t st, int *launches);
void launch_ratio_test(const surf_dmatch *m, int nq, float ratio, uint8_t *good, cudaStream_t st, int *launches);
void launch_merge_topk(const surf_dmatch *in, int n_ranks, int nq, int k, surf_dmatch *out, cudaStream_t st,
                       int *launches);

// ---- inter-slice correspondence votes (surf_vote.cu) ----
struct VoteArgs {
    const surf_dmatch *m_train;   // [rows][3] knnMatch(k=3) against the train database (image-local indices)
    const surf_dmatch *m_bogus;   // [rows][1] knnMatch(k=1) against the bogus database
    const surf_keypoint *qk;      // [rows] query keypoints, all query images packed
    const surf_keypoint *tk;      // train database keypoints
    const int *t_img_off;         // train database image offsets (device)
    const int *q_off;             // [n_query_images + 1] (device)
    int n_query_images, n_train_images, max_rows_per_image;
    float bogus_ratio;
    int ref_x, ref_y;
    float thr_scale, thr_ori;     // (float)log(1.5f), (float)(20 pi / 180)
    surf_dmatch *out;             // [rows] surviving matches of image i from q_off[i]
    float4 *booth;                // [rows]
    float *booth_thr;             // [rows]
    int *n_matches;               // [n_query_images]
    unsigned long long *keys;     // [n_query_images]
    int *votes;                   // [n_query_images][n_train_images]
    surf_vote_result *results;    // [n_query_images]
};
void launch_vote(const VoteArgs &a, cudaStream_t st, int *launches);

}  // namespace surfb200
