// surf_match.cu -- stage 5: brute-force L2 k-nearest-neighbour matching (BFMatcher(NORM_L2)::knnMatch),
// the call made at SurfFeatures/Logic/vtkSlicerSurfFeaturesLogic.cxx:640.
//
// Exact path: fp32 sum of squared differences per (query, train) pair, per-query running top-k
// (k <= 4) in ascending order, ties resolved to the lower train index, sqrt at the end.
// The train set is split across gridDim.y CTAs; partial lists are merged by k_merge_topk, which is
// also the merge step for a row-sharded keyframe database after the candidates were all-gathered.
#include <cfloat>

#include "surf_internal.h"

namespace surfb200 {

constexpr int kMaxK = 4;
constexpr int kTile = 64;

struct TopK {
    float d[kMaxK];
    int i[kMaxK];
};

__device__ __forceinline__ void topk_init(TopK &t)
{
#pragma unroll
    for (int r = 0; r < kMaxK; r++) {
        t.d[r] = INFINITY;
        t.i[r] = -1;
    }
}

// Insert keeping ascending (distance, index) order; `k` live entries.
__device__ __forceinline__ void topk_insert(TopK &t, int k, float dist, int idx)
{
    if (!(dist < t.d[k - 1] || (dist == t.d[k - 1] && (unsigned)idx < (unsigned)t.i[k - 1]))) return;
    t.d[k - 1] = dist;
    t.i[k - 1] = idx;
#pragma unroll
    for (int r = kMaxK - 1; r > 0; r--) {
        if (r < k) {
            const bool sw = t.d[r] < t.d[r - 1] || (t.d[r] == t.d[r - 1] && (unsigned)t.i[r] < (unsigned)t.i[r - 1]);
            if (sw) {
                float td = t.d[r]; t.d[r] = t.d[r - 1]; t.d[r - 1] = td;
                int ti = t.i[r]; t.i[r] = t.i[r - 1]; t.i[r - 1] = ti;
            }
        }
    }
}

// Squared distances of a 64-query tile against a range of train rows.  Shared tiles are stored
// dimension-major so that each thread reads its 4 queries / 4 train rows as one float4.
template <int D>
__global__ void __launch_bounds__(256) k_match_exact(const float *__restrict__ q, int nq, const float *__restrict__ t,
                                                     int nt, int k, surf_dmatch *__restrict__ part, int chunk,
                                                     int train_base)
{
    extern __shared__ __align__(16) float sm[];
    float *qs = sm;                     // [D][64]
    float *ts = sm + D * kTile;         // [D][64]
    float *ds = sm + 2 * D * kTile;     // [64][65] squared distances of the current tile
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int q0 = blockIdx.x * kTile;
    const int t_begin = blockIdx.y * chunk, t_end = min(nt, t_begin + chunk);

    for (int idx = tid; idx < kTile * (D / 4); idx += 256) {
        const int row = idx % kTile, d4 = idx / kTile;
        float4 v = make_float4(0, 0, 0, 0);
        if (q0 + row < nq) v = __ldg(reinterpret_cast<const float4 *>(q + (size_t)(q0 + row) * D) + d4);
        qs[(d4 * 4 + 0) * kTile + row] = v.x;
        qs[(d4 * 4 + 1) * kTile + row] = v.y;
        qs[(d4 * 4 + 2) * kTile + row] = v.z;
        qs[(d4 * 4 + 3) * kTile + row] = v.w;
    }
    TopK best;
    topk_init(best);

    for (int t0 = t_begin; t0 < t_end; t0 += kTile) {
        __syncthreads();  // previous tile fully consumed (also orders the q tile stores)
        for (int idx = tid; idx < kTile * (D / 4); idx += 256) {
            const int row = idx % kTile, d4 = idx / kTile;
            float4 v = make_float4(0, 0, 0, 0);
            if (t0 + row < t_end) v = __ldg(reinterpret_cast<const float4 *>(t + (size_t)(t0 + row) * D) + d4);
            ts[(d4 * 4 + 0) * kTile + row] = v.x;
            ts[(d4 * 4 + 1) * kTile + row] = v.y;
            ts[(d4 * 4 + 2) * kTile + row] = v.z;
            ts[(d4 * 4 + 3) * kTile + row] = v.w;
        }
        __syncthreads();
        float acc[4][4];
#pragma unroll
        for (int a = 0; a < 4; a++)
#pragma unroll
            for (int c = 0; c < 4; c++) acc[a][c] = 0.f;
#pragma unroll 8
        for (int d = 0; d < D; d++) {
            const float4 a4 = *reinterpret_cast<const float4 *>(qs + d * kTile + ty * 4);
            const float4 b4 = *reinterpret_cast<const float4 *>(ts + d * kTile + tx * 4);
            const float av[4] = {a4.x, a4.y, a4.z, a4.w}, bv[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
            for (int a = 0; a < 4; a++)
#pragma unroll
                for (int c = 0; c < 4; c++) {
                    const float df = av[a] - bv[c];
                    acc[a][c] = fmaf(df, df, acc[a][c]);
                }
        }
#pragma unroll
        for (int a = 0; a < 4; a++)
#pragma unroll
            for (int c = 0; c < 4; c++) ds[(ty * 4 + a) * (kTile + 1) + tx * 4 + c] = acc[a][c];
        __syncthreads();
        if (tid < kTile) {
            const int lim = min(kTile, t_end - t0);
            for (int c = 0; c < lim; c++) topk_insert(best, k, ds[tid * (kTile + 1) + c], train_base + t0 + c);
        }
    }
    if (tid < kTile && q0 + tid < nq) {
        surf_dmatch *o = part + ((size_t)blockIdx.y * nq + q0 + tid) * k;
        for (int r = 0; r < k; r++) {
            o[r].query_idx = q0 + tid;
            o[r].train_idx = best.i[r];
            o[r].img_idx = best.i[r] < 0 ? -1 : 0;
            o[r].distance = best.d[r];  // squared; the merge takes the root
        }
    }
}

// Merges n_lists per-query candidate lists (each ascending) into the final k; optional sqrt.
__global__ void __launch_bounds__(128) k_merge_topk(const surf_dmatch *__restrict__ in, int n_lists, int nq, int k,
                                                    surf_dmatch *__restrict__ out, int take_sqrt)
{
    const int qi = blockIdx.x * blockDim.x + threadIdx.x;
    if (qi >= nq) return;
    TopK best;
    topk_init(best);
    for (int l = 0; l < n_lists; l++) {
        const surf_dmatch *m = in + ((size_t)l * nq + qi) * k;
        for (int r = 0; r < k; r++)
            if (m[r].train_idx >= 0) topk_insert(best, k, m[r].distance, m[r].train_idx);
    }
    surf_dmatch *o = out + (size_t)qi * k;
    for (int r = 0; r < k; r++) {
        o[r].query_idx = qi;
        o[r].train_idx = best.i[r];
        o[r].img_idx = best.i[r] < 0 ? -1 : 0;
        o[r].distance = take_sqrt ? sqrtf(best.d[r]) : best.d[r];
    }
}

__global__ void __launch_bounds__(256) k_ratio_test(const surf_dmatch *__restrict__ m, int nq, float ratio,
                                                    uint8_t *__restrict__ good)
{
    const int qi = blockIdx.x * blockDim.x + threadIdx.x;
    if (qi >= nq) return;
    const surf_dmatch a = m[2 * qi], b = m[2 * qi + 1];
    good[qi] = (a.train_idx >= 0 && b.train_idx >= 0 && a.distance < ratio * b.distance) ? 1 : 0;
}

// Train-range split used for (nq, nt): enough CTAs to give every SM about two.
static void match_split(int nq, int nt, int *splits_out, int *chunk_out)
{
    if (nt <= 0 || nq <= 0) {
        *splits_out = 1;
        *chunk_out = kTile;
        return;
    }
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int qtiles = (nq + kTile - 1) / kTile;
    int splits = (2 * sms + qtiles - 1) / qtiles;
    const int ttiles = (nt + kTile - 1) / kTile;
    if (splits > ttiles) splits = ttiles;
    if (splits < 1) splits = 1;
    const int chunk = ((ttiles + splits - 1) / splits) * kTile;
    *splits_out = (nt + chunk - 1) / chunk;
    *chunk_out = chunk;
}

size_t match_scratch_records(int nq, int nt, int k)
{
    int splits, chunk;
    match_split(nq, nt, &splits, &chunk);
    return (size_t)splits * nq * k;
}

void launch_match_knn(const float *q, int nq, const float *t, int nt, int dim, int k, surf_dmatch *out, int train_base,
                      surf_dmatch *part, cudaStream_t st, int *launches)
{
    int splits, chunk;
    match_split(nq, nt, &splits, &chunk);
    const int qtiles = (nq + kTile - 1) / kTile;
    dim3 grid(qtiles, splits);
    const size_t smem = (size_t)(2 * dim * kTile + kTile * (kTile + 1)) * sizeof(float);
    if (dim == 64) {
        cudaFuncSetAttribute(k_match_exact<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        k_match_exact<64><<<grid, 256, smem, st>>>(q, nq, t, nt, k, part, chunk, train_base);
    } else {
        cudaFuncSetAttribute(k_match_exact<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        k_match_exact<128><<<grid, 256, smem, st>>>(q, nq, t, nt, k, part, chunk, train_base);
    }
    k_merge_topk<<<(nq + 127) / 128, 128, 0, st>>>(part, splits, nq, k, out, 1);
    *launches += 2;
}

void launch_ratio_test(const surf_dmatch *m, int nq, float ratio, uint8_t *good, cudaStream_t st, int *launches)
{
    k_ratio_test<<<(nq + 255) / 256, 256, 0, st>>>(m, nq, ratio, good);
    *launches += 1;
}

void launch_merge_topk(const surf_dmatch *in, int n_ranks, int nq, int k, surf_dmatch *out, cudaStream_t st,
                       int *launches)
{
    k_merge_topk<<<(nq + 127) / 128, 128, 0, st>>>(in, n_ranks, nq, k, out, 0);
    *launches += 1;
}

}  // namespace surfb200
