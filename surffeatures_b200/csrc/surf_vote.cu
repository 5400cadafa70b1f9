// surf_vote.cu -- kernels of the inter-slice correspondence step ("next" row f3 of SURVEY.md section 8):
// bogus-ratio filter, Hough poll booths, pairwise pose-consistency vote, per-train-image vote count.
// Reference: SurfFeatures/Logic/vtkSlicerSurfFeaturesLogic.cxx:73-345 (HoughCodeSimilarity, angle_radians,
// compatible_poll_booths_line_segment, houghTransform) and :1457-1512 (computeNextInterSliceCorrespondence).
//
// Arithmetic model: fp32 expressions are separate operations
// (this unit is compiled with -fmad=false), the float libm calls of the reference are evaluated in double and
// rounded once.  The reference's angle_radians() maps degrees to 2*pi*a/180 and is applied to the booth
// orientation a second time inside the compatibility test; both quirks are kept.
#include <cfloat>

#include "surf_internal.h"

namespace surfb200 {

namespace {

constexpr double kPi = 3.141592653589793;

__device__ __forceinline__ float angle_radians(float a) { return (float)((2.0 * kPi * (double)a) / 180.0); }

// One Hough poll booth, stored ready for the pair test: col, row, angle_radians(ori), log(scale); thr = (30/43) * scale.
struct Booth {
    float4 v;
    float thr;
};

// HoughCodeSimilarity::SetHoughCode(refy, refx, 0, 10) on the query keypoint, then FindPollBooth on the train keypoint.
__device__ Booth poll_booth(const surf_keypoint &q, const surf_keypoint &t, int refx, int refy)
{
    const float key_ori = angle_radians(q.angle);
    const float diff_row = (float)refy - q.y, diff_col = (float)refx - q.x;
    const float dist_to_booth = sqrtf(diff_col * diff_col + diff_row * diff_row);
    const float angle_to_booth = (float)atan2((double)diff_row, (double)diff_col);
    const float angle_diff = 0.0f - key_ori;
    const float scale_diff = 10.0f / q.size;
    const float f_ori = angle_radians(t.angle);
    const float dist = dist_to_booth * (t.size / q.size);
    const float angle = angle_to_booth + (f_ori - key_ori);
    const float row = (float)sin((double)angle) * dist + t.y;
    const float col = (float)cos((double)angle) * dist + t.x;
    const float ori = f_ori + angle_diff;
    const float scl = scale_diff * t.size;
    Booth b;
    b.v = make_float4(col, row, angle_radians(ori), (float)log((double)scl));
    b.thr = (30.0f / 43.0f) * scl;
    return b;
}

// compatible_poll_booths_line_segment(a as "training", b as "poll") with the default thresholds.
__device__ __forceinline__ bool compatible(const float4 a, const float a_thr, const float4 b, float thr_scale,
                                           float thr_ori)
{
    const float log_diff = fabsf(a.w - b.w);
    const float dc = b.x - a.x, dr = b.y - a.y;
    const float dist = sqrtf(dc * dc + dr * dr);
    float od = b.z > a.z ? b.z - a.z : a.z - b.z;
    // `od > 2*pi - od` in double is true exactly for the floats above pi (the smallest is 0x40490FDB)
    if (od >= __int_as_float(0x40490FDB)) od = (float)(2.0 * kPi - (double)od);
    return dist < a_thr && log_diff < thr_scale && od < thr_ori;
}

// One CTA per query image: bogus-ratio filter with order-preserving compaction, DMatch + poll booth of every
// surviving match.  m_train: k = 3 records per query row (already image-local), m_bogus: k = 1.
__global__ void __launch_bounds__(256) k_vote_filter(const surf_dmatch *__restrict__ m_train,
                                                     const surf_dmatch *__restrict__ m_bogus,
                                                     const surf_keypoint *__restrict__ qk,
                                                     const surf_keypoint *__restrict__ tk,
                                                     const int *__restrict__ t_img_off, const int *__restrict__ q_off,
                                                     float bogus_ratio, int refx, int refy, surf_dmatch *__restrict__ out,
                                                     float4 *__restrict__ booth, float *__restrict__ booth_thr,
                                                     int *__restrict__ n_matches)
{
    __shared__ int s_warp[8];
    __shared__ int s_base;
    const int img = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int r0 = q_off[img], n = q_off[img + 1] - r0;
    if (tid == 0) s_base = 0;
    __syncthreads();
    for (int c = 0; c < n; c += 256) {
        const int j = c + tid;
        bool keep = false;
        surf_dmatch m{};
        if (j < n) {
            m = m_train[(size_t)(r0 + j) * 3];
            const float ratio = m.distance / m_bogus[r0 + j].distance;
            keep = m.train_idx >= 0 && !(ratio > bogus_ratio);
        }
        const unsigned bal = __ballot_sync(0xffffffffu, keep);
        if (lane == 0) s_warp[warp] = __popc(bal);
        __syncthreads();
        int before = s_base;
        for (int w = 0; w < warp; w++) before += s_warp[w];
        if (keep) {
            const int pos = before + __popc(bal & ((1u << lane) - 1u));
            m.query_idx = j;
            out[r0 + pos] = m;
            const Booth b = poll_booth(qk[r0 + j], tk[t_img_off[m.img_idx] + m.train_idx], refx, refy);
            booth[r0 + pos] = b.v;
            booth_thr[r0 + pos] = b.thr;
        }
        __syncthreads();
        if (tid == 0) {
            int tot = 0;
            for (int w = 0; w < 8; w++) tot += s_warp[w];
            s_base += tot;
        }
        __syncthreads();
    }
    if (tid == 0) n_matches[img] = s_base;
}

// Votes of every booth: thread a counts the booths b of its image compatible with a; the winner is the largest
// count with the lowest index (the reference's `if (iCount > iMaxCount)` scan), kept as one 64-bit key.
__global__ void __launch_bounds__(256) k_hough_count(const float4 *__restrict__ booth, const float *__restrict__ booth_thr,
                                                     const int *__restrict__ q_off, const int *__restrict__ n_matches,
                                                     float thr_scale, float thr_ori, unsigned long long *__restrict__ keys)
{
    __shared__ float4 s_b[256];
    const int img = blockIdx.y, tid = threadIdx.x;
    const int r0 = q_off[img], m = n_matches[img];
    if ((int)blockIdx.x * 256 >= m) return;
    const int a = blockIdx.x * 256 + tid;
    const bool live = a < m;
    const float4 va = live ? booth[r0 + a] : make_float4(0.f, 0.f, 0.f, 0.f);
    const float ta = live ? booth_thr[r0 + a] : 0.f;
    int count = 0;
    for (int c = 0; c < m; c += 256) {
        __syncthreads();
        if (c + tid < m) s_b[tid] = booth[r0 + c + tid];
        __syncthreads();
        const int lim = min(256, m - c);
        for (int b = 0; b < lim; b++) count += compatible(va, ta, s_b[b], thr_scale, thr_ori) ? 1 : 0;
    }
    if (live && count > 0)
        atomicMax(keys + img, ((unsigned long long)count << 32) | (unsigned long long)(0xFFFFFFFFu - (unsigned)a));
}

// One CTA per query image: reject the matches incompatible with the winning booth, count the votes per train
// image, pick the first image with the largest count.
__global__ void __launch_bounds__(256) k_hough_apply(const float4 *__restrict__ booth, const float *__restrict__ booth_thr,
                                                     const int *__restrict__ q_off, const int *__restrict__ n_matches,
                                                     const unsigned long long *__restrict__ keys, float thr_scale,
                                                     float thr_ori, surf_dmatch *__restrict__ out, int *__restrict__ votes,
                                                     int n_train_images, surf_vote_result *__restrict__ results)
{
    __shared__ unsigned long long s_best[8];
    const int img = blockIdx.x, tid = threadIdx.x;
    const int r0 = q_off[img], m = n_matches[img];
    const unsigned long long key = keys[img];
    const int hough_count = (int)(key >> 32);
    int *v = votes + (size_t)img * n_train_images;
    if (m > 0 && hough_count > 0) {
        const int best = (int)(0xFFFFFFFFu - (unsigned)(key & 0xFFFFFFFFu));
        const float4 vb = booth[r0 + best];
        const float tb = booth_thr[r0 + best];
        for (int b = tid; b < m; b += 256) {
            if (compatible(vb, tb, booth[r0 + b], thr_scale, thr_ori)) {
                atomicAdd(v + out[r0 + b].img_idx, 1);
            } else {
                surf_dmatch &o = out[r0 + b];
                o.img_idx = -1;
                o.query_idx = -1;
                o.train_idx = -1;
            }
        }
    }
    __syncthreads();
    // first maximum over the train images: largest count, lowest image index
    unsigned long long bk = 0;  // (count + 1) in the high word so that an all-zero histogram still picks image 0
    for (int j = tid; j < n_train_images; j += 256) {
        const unsigned long long k2 = ((unsigned long long)(unsigned)(v[j] + 1) << 32) | (0xFFFFFFFFu - (unsigned)j);
        bk = k2 > bk ? k2 : bk;
    }
    for (int s = 16; s > 0; s >>= 1) {
        const unsigned long long o = __shfl_xor_sync(0xffffffffu, bk, s);
        bk = o > bk ? o : bk;
    }
    if ((tid & 31) == 0) s_best[tid >> 5] = bk;
    __syncthreads();
    if (tid == 0) {
        for (int w = 1; w < 8; w++) bk = s_best[w] > bk ? s_best[w] : bk;
        surf_vote_result r;
        r.n_matches = m;
        r.hough_count = hough_count;
        r.best_image = n_train_images > 0 ? (int)(0xFFFFFFFFu - (unsigned)(bk & 0xFFFFFFFFu)) : -1;
        r.best_count = n_train_images > 0 ? (int)(bk >> 32) - 1 : -1;
        results[img] = r;
    }
}

}  // namespace

void launch_vote(const VoteArgs &a, cudaStream_t st, int *launches)
{
    if (a.n_query_images <= 0) return;
    cudaMemsetAsync(a.keys, 0, sizeof(unsigned long long) * a.n_query_images, st);
    if (a.n_train_images > 0) cudaMemsetAsync(a.votes, 0, sizeof(int) * (size_t)a.n_query_images * a.n_train_images, st);
    k_vote_filter<<<a.n_query_images, 256, 0, st>>>(a.m_train, a.m_bogus, a.qk, a.tk, a.t_img_off, a.q_off, a.bogus_ratio,
                                                    a.ref_x, a.ref_y, a.out, a.booth, a.booth_thr, a.n_matches);
    if (a.max_rows_per_image > 0) {
        dim3 g((a.max_rows_per_image + 255) / 256, a.n_query_images);
        k_hough_count<<<g, 256, 0, st>>>(a.booth, a.booth_thr, a.q_off, a.n_matches, a.thr_scale, a.thr_ori, a.keys);
    }
    k_hough_apply<<<a.n_query_images, 256, 0, st>>>(a.booth, a.booth_thr, a.q_off, a.n_matches, a.keys, a.thr_scale,
                                                    a.thr_ori, a.out, a.votes, a.n_train_images, a.results);
    if (launches) *launches += 3;
}

}  // namespace surfb200
