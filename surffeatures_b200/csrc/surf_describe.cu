// surf_describe.cu -- stage 4 of the SURF path (SURFInvoker in OpenCV 2.4.5 nonfree/surf.cpp; SURVEY.md A4,
// A5; rows a7/a8), reached from SurfFeatures/Logic/vtkSlicerSurfFeaturesLogic.cxx:1384-1389.
//
//   k_orient      four keypoints per warp: 113 Haar samples from the integral image each, 72 sliding 60-degree
//                 windows (lane = keypoint x window residue), dominant angle; classifies the keypoint by
//                 descriptor-window size and appends it to that class's work list.
//   k_descriptor  one CTA per keypoint, persistent CTAs on an atomic queue per size class.  The rotated
//                 n x n window's bounding box in the 8-bit image is staged in shared memory, every window
//                 pixel is a bilinear tap rounded to uchar, the window is reduced to 21 x 21 with
//                 INTER_AREA, then 2x2 Haar gradients, 4x4 sub-region sums, L2 normalisation.
//
// Compiled with -fmad=false: every fp32 expression rounds as in the SSE-only reference build.
#include <cuda_pipeline.h>

#include <cfloat>
#include <cmath>

#include "surf_internal.h"
#include "surf_ptx.cuh"

namespace surfb200 {

#define FULL_MASK 0xffffffffu

// ------------------------------------------------------------------------------------------------
// constant tables (fixed by the algorithm, not by engine parameters)
// ------------------------------------------------------------------------------------------------
__constant__ float c_ori_x[kOriSamples];  // sample offsets -6..6, kept as floats (an I2F per use sat on the slow conversion pipe)
__constant__ float c_ori_y[kOriSamples];
__constant__ float c_ori_w[kOriSamples];
__constant__ float c_desc_g[20];   // getGaussianKernel(20, 3.3) as floats; the sample weight is the float product g[y] * g[x]
// 1.0f twice, opaque to ptxas (see tap_fixed2)
__constant__ float2 c_one2 = {1.f, 1.f};

static void gaussian_taps(int n, double sigma, float *out)
{
    // getGaussianKernel(n, sigma, CV_32F) as OpenCV 2.4.5 builds it.
    double s2 = -0.5 / (sigma * sigma), acc = 0;
    for (int i = 0; i < n; i++) {
        double x = i - (n - 1) * 0.5;
        out[i] = (float)std::exp(s2 * x * x);
        acc += out[i];
    }
    acc = 1. / acc;
    for (int i = 0; i < n; i++) out[i] = (float)(out[i] * acc);
}

void upload_tables()
{
    float px[kOriSamples], py[kOriSamples];
    float pw[kOriSamples], g[13], gd[20];
    gaussian_taps(13, 2.5, g);
    int n = 0;
    for (int i = -6; i <= 6; i++)
        for (int j = -6; j <= 6; j++)
            if (i * i + j * j <= 36) {
                px[n] = (float)i;  // the outer loop variable is the x offset
                py[n] = (float)j;
                pw[n++] = g[i + 6] * g[j + 6];
            }
    gaussian_taps(20, 3.3, gd);
    cudaMemcpyToSymbol(c_ori_x, px, sizeof(px));
    cudaMemcpyToSymbol(c_ori_y, py, sizeof(py));
    cudaMemcpyToSymbol(c_ori_w, pw, sizeof(pw));
    cudaMemcpyToSymbol(c_desc_g, gd, sizeof(gd));
}

__device__ __forceinline__ float fast_atan2_deg(float y, float x)
{
    // fastAtan2 (core/mathfuncs.cpp): 7th-order odd polynomial, degrees in [0, 360].
    const float k = (float)(180 / 3.1415926535897932384626433832795);
    const float p1 = 0.9997878412794807f * k, p3 = -0.3258083974640975f * k;
    const float p5 = 0.1555786518463281f * k, p7 = -0.04432655554792128f * k;
    const float ax = fabsf(x), ay = fabsf(y);
    float a, c, c2;
    if (ax >= ay) {
        c = ay / (ax + (float)DBL_EPSILON);
        c2 = c * c;
        a = (((p7 * c2 + p5) * c2 + p3) * c2 + p1) * c;
    } else {
        c = ax / (ay + (float)DBL_EPSILON);
        c2 = c * c;
        a = 90.f - (((p7 * c2 + p5) * c2 + p3) * c2 + p1) * c;
    }
    if (x < 0) a = 180.f - a;
    if (y < 0) a = 360.f - a;
    return a;
}

// ---- packed fp32 (Blackwell FADD2 / FMUL2 / FFMA2): two IEEE single operations per instruction ----
// compile-time switches of the three places that use it (A/B builds: tools/build_variant.sh)
#ifndef SURF_TAP_F32X2
#define SURF_TAP_F32X2 1
#endif
__device__ __forceinline__ uint64_t f2_pack(float lo, float hi)
{
    uint64_t r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void f2_unpack(uint64_t v, float &lo, float &hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ uint64_t f2_mul(uint64_t a, uint64_t b)
{
    uint64_t r;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ uint64_t f2_sub(uint64_t a, uint64_t b)
{
    uint64_t r;
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ uint64_t f2_fma(uint64_t a, uint64_t b, uint64_t c)
{
    uint64_t r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
    return r;
}

// uchar -> float without a conversion instruction: 2^23 + v is exact, minus 2^23 gives v.
__device__ __forceinline__ float u8f(unsigned v) { return __uint_as_float(0x4B000000u | v) - 8388608.0f; }

// cvRound of a float in [0, 2^22) -> low byte: adding 1.5 * 2^23 rounds to nearest-even at integer
// granularity in one fp32 addition.
__device__ __forceinline__ unsigned round_u8(float v) { return __float_as_uint(v + 12582912.0f) & 0xffu; }

__device__ __forceinline__ int sat_u8(float v)
{
    int r = __float2int_rn(v);
    return r < 0 ? 0 : (r > 255 ? 255 : r);
}

__device__ __forceinline__ void find_frame(const int *offsets, int B, int w, int &b, int &slot)
{
    int lo = 0, hi = B;
    while (hi - lo > 1) {
        int m = (lo + hi) >> 1;
        if (offsets[m] <= w) lo = m; else hi = m;
    }
    b = lo;
    slot = w - offsets[lo];
}

// ------------------------------------------------------------------------------------------------
// K4a: orientation.  A warp takes FOUR keypoints at a time.  The 113 Haar samples of each are gathered by the whole
// warp (one keypoint after the other, 32 samples per round, compacted in sample order); the 72 sliding windows of the
// four keypoints are then summed together: lane l owns the nine windows (l % 8) + 8 s, s = 0..8, of keypoint l / 8, so
// all 288 (keypoint, window) pairs are busy (one keypoint per warp left 24 of 96 window slots idle and spent 12
// instructions per sample; here 31 per sample serve four keypoints).  Sums run in sample order in fp32, the first strict
// maximum over increasing window index wins: the same arithmetic in the same order as before, bit-identical angles.
// The per-keypoint tail (correctly rounded sin / cos in double, class lists) runs on four lanes at once.
// ------------------------------------------------------------------------------------------------
constexpr int kOriWarps = 4;
constexpr int kOriPerWarp = 4;

struct __align__(16) OriSample {
    float X, Y;          // Gaussian-weighted Haar responses
    unsigned m0, m1;     // membership of the sample in windows 0..31 / 32..63 (bit w = window w); 0 for padding
};

// The 72 sliding windows (0, 5, ..., 355 degrees; a sample with rounded angle a belongs to window w iff
// d = |a - 5w| < 30 or d > 330) that contain angle a form one circular run of 11 or 12 windows:
// [floor((a-30)/5)+1, ceil((a+30)/5)-1] mod 72.  Returned as a 72-bit mask (m0, m1, m2 = windows 64..71).
__device__ __forceinline__ void window_mask(int a, unsigned &m0, unsigned &m1, unsigned &m2)
{
    const int wmin = (a + 330) / 5 - 71, wmax = (a + 34) / 5 - 1;
    const int c = wmax - wmin + 1;
    const int start = wmin < 0 ? wmin + 72 : wmin;
    const unsigned long long base = (1ull << c) - 1;
    unsigned long long lo = start < 64 ? base << start : 0ull;
    unsigned long long hi = start >= 64 ? base << (start - 64) : (start + c > 64 ? base >> (64 - start) : 0ull);
    lo |= hi >> 8;  // windows 72.. wrap around to 0..
    m0 = (unsigned)lo;
    m1 = (unsigned)(lo >> 32);
    m2 = (unsigned)hi & 0xffu;
}

__global__ void __launch_bounds__(kOriWarps * 32) k_orient(const __grid_constant__ DescribeArgs A)
{
    // 117 slots: the four lists of a warp are 468 words apart, i.e. 20 banks: the four 16-byte reads of a warp (one
    // per keypoint, same sample index) fall into four different bank groups
    constexpr int kSlots = kOriSamples + 4;
    __shared__ OriSample s_smp[kOriWarps][kOriPerWarp][kSlots];
    __shared__ unsigned char s_m2[kOriWarps][kOriPerWarp][kSlots + 3];  // membership in windows 64..71
    // window membership of every rounded angle 0..360, built once per CTA (window_mask is ~40 integer instructions
    // with 64-bit shifts; the table makes it one 16-byte shared-memory read per sample)
    __shared__ uint4 s_wmask[361];
    for (int a = threadIdx.x; a <= 360; a += kOriWarps * 32) {
        unsigned m0, m1, m2;
        window_mask(a, m0, m1, m2);
        s_wmask[a] = make_uint4(m0, m1, m2, 0u);
    }
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int grp = lane >> 3, l8 = lane & 7;  // window phase: keypoint of the four, window residue
    const int W = A.geo.W, H = A.geo.H, SP = A.geo.SP;
    const int total = A.offsets[A.B];
    const bool want_desc = (A.desc != nullptr) || (A.patches != nullptr);
    const unsigned bit0 = 1u << l8, bit1 = bit0 << 8, bit2 = bit0 << 16, bit3 = bit0 << 24;
    // a lane always gathers the same four samples (lane, lane + 32, ...): their offsets and weights live in registers (indexed
    // per lane the constant bank served one address at a time: the top stall of the gather)
    float kx[4], ky[4], kw[4];
#pragma unroll
    for (int r = 0; r < 4; r++) {
        const int k = min(r * 32 + lane, kOriSamples - 1);
        kx[r] = c_ori_x[k];
        ky[r] = c_ori_y[k];
        kw[r] = c_ori_w[k];
    }
    int fb = 0;  // frame of the last keypoint: keypoints come in frame order, so the next one is almost always in it too

    for (int w0 = (blockIdx.x * kOriWarps + warp) * kOriPerWarp; w0 < total; w0 += gridDim.x * kOriWarps * kOriPerWarp) {
        // this lane's keypoint of the four (window phase and tail): filled in while the warp gathers the samples
        int my_b = 0, my_slot = 0, my_na = 0;
        float my_s = 0.f;
        bool my_live = false, my_dead = false;
        int na_max = 0;
        for (int j = 0; j < kOriPerWarp; j++) {
            const int w = w0 + j;
            if (w >= total) {  // (uniform) no keypoint: an all-padding list, so the joint walk below reads valid entries
                for (int k = lane; k < kSlots; k += 32) {
                    s_smp[warp][j][k] = OriSample{0.f, 0.f, 0u, 0u};
                    s_m2[warp][j][k] = 0;
                }
                continue;
            }
            int b, slot;
            if (w >= A.offsets[fb] && w < A.offsets[fb + 1]) {
                b = fb;
                slot = w - A.offsets[fb];
            } else {
                find_frame(A.offsets, A.B, w, b, slot);
                fb = b;
            }
            const surf_keypoint *kpp = A.kps + (size_t)b * A.cap + slot;
            const float cx = kpp->x, cy = kpp->y, size = kpp->size;
            const int *S = reinterpret_cast<const int *>(A.sum) + (size_t)b * A.geo.sum_frame;
            const float s = size * 1.2f / 9.0f;
            const int g = 2 * __float2int_rn(2 * s);
            bool dead = (H + 1 < g) || (W + 1 < g);
            int na = 0;
            OriSample *smp = s_smp[warp][j];
            unsigned char *smp2 = s_m2[warp][j];
            if (!dead && !A.upright) {
                // ---- 113 samples, 4 rounds of 32, compacted in sample order ----
                const float hg = (float)(g - 1) / 2;
                const int h = g / 2;
                const float wa = 1 / ((float)h * g), wn = -1 / ((float)h * g);
#pragma unroll
                for (int r = 0; r < 4; r++) {
                    const int k = r * 32 + lane;
                    bool valid = false;
                    float X = 0, Y = 0;
                    if (k < kOriSamples) {
                        const int x = __float2int_rn(cx + kx[r] * s - hg);
                        const int y = __float2int_rn(cy + ky[r] * s - hg);
                        if (!(y < 0 || y >= (H + 1) - g || x < 0 || x >= (W + 1) - g)) {
                            valid = true;
                            const int *o = S + (size_t)y * SP + x;
                            const int *oh = o + (size_t)h * SP, *og = o + (size_t)g * SP;
                            const int p00 = __ldg(o), p0h = __ldg(o + h), p0g = __ldg(o + g);
                            const int ph0 = __ldg(oh), phg = __ldg(oh + g);
                            const int pg0 = __ldg(og), pgh = __ldg(og + h), pgg = __ldg(og + g);
                            // x gradient: right half minus left half; y gradient: top half minus bottom half
                            const int left = p00 + pgh - pg0 - p0h, right = p0h + pgg - pgh - p0g;
                            const int top = p00 + phg - ph0 - p0g, bottom = ph0 + pgg - pg0 - phg;
                            const float vx = (float)((double)((float)left * wn) + (double)((float)right * wa));
                            const float vy = (float)((double)((float)top * wa) + (double)((float)bottom * wn));
                            X = vx * kw[r];
                            Y = vy * kw[r];
                        }
                    }
                    const unsigned bal = __ballot_sync(FULL_MASK, valid);
                    if (valid) {
                        const uint4 wm = s_wmask[__float2int_rn(fast_atan2_deg(Y, X))];
                        const int pos = na + __popc(bal & ((1u << lane) - 1));
                        smp[pos] = OriSample{X, Y, wm.x, wm.y};
                        smp2[pos] = (unsigned char)wm.z;
                    }
                    na += __popc(bal);
                }
                if (na == 0) dead = true;
            }
            // padding up to the end of the list: samples that belong to no window (the four lists are walked together)
            for (int k = na + lane; k < kSlots; k += 32) {
                smp[k] = OriSample{0.f, 0.f, 0u, 0u};
                smp2[k] = 0;
            }
            if (grp == j) {
                my_b = b; my_slot = slot; my_na = na; my_s = s;
                my_live = true;
                my_dead = dead;
            }
            na_max = max(na_max, na);
        }
        __syncwarp();

        float dir = 360.f - 90.f;
        if (!A.upright && na_max > 0) {
            // ---- 72 sliding windows per keypoint, fp32 sums in sample order; lane owns windows l8 + 8 s of keypoint grp ----
            const OriSample *smp = s_smp[warp][grp];
            const unsigned char *smp2 = s_m2[warp][grp];
            float sx[9], sy[9];
#pragma unroll
            for (int q = 0; q < 9; q++) sx[q] = sy[q] = 0.f;
#pragma unroll 2
            for (int k = 0; k < na_max; k++) {
                const OriSample o = smp[k];
                const unsigned m2 = smp2[k];
                if (o.m0 & bit0) { sx[0] += o.X; sy[0] += o.Y; }
                if (o.m0 & bit1) { sx[1] += o.X; sy[1] += o.Y; }
                if (o.m0 & bit2) { sx[2] += o.X; sy[2] += o.Y; }
                if (o.m0 & bit3) { sx[3] += o.X; sy[3] += o.Y; }
                if (o.m1 & bit0) { sx[4] += o.X; sy[4] += o.Y; }
                if (o.m1 & bit1) { sx[5] += o.X; sy[5] += o.Y; }
                if (o.m1 & bit2) { sx[6] += o.X; sy[6] += o.Y; }
                if (o.m1 & bit3) { sx[7] += o.X; sy[7] += o.Y; }
                if (m2 & bit0) { sx[8] += o.X; sy[8] += o.Y; }
            }
            // first strict maximum over increasing window index: inside the lane (s ascending), then over the 8 lanes
            float bmod = 0, bx = 0, by = 0;
            int bi = 1 << 30;
#pragma unroll
            for (int q = 0; q < 9; q++) {
                const float qq = sx[q] * sx[q] + sy[q] * sy[q];
                if (qq > bmod) { bmod = qq; bx = sx[q]; by = sy[q]; bi = l8 + 8 * q; }
            }
#pragma unroll
            for (int d = 4; d > 0; d >>= 1) {
                const float om = __shfl_xor_sync(FULL_MASK, bmod, d);
                const float ox = __shfl_xor_sync(FULL_MASK, bx, d);
                const float oy = __shfl_xor_sync(FULL_MASK, by, d);
                const int oi = __shfl_xor_sync(FULL_MASK, bi, d);
                if (om > bmod || (om == bmod && oi < bi)) {
                    bmod = om;
                    bx = ox;
                    by = oy;
                    bi = oi;
                }
            }
            if (my_na > 0) dir = fast_atan2_deg(-by, bx);
        }
        __syncwarp();  // the next four keypoints overwrite the sample lists

        // ---- tail: lane 0 of each group finishes its keypoint
        if (my_live && l8 == 0) {
            const int b = my_b, slot = my_slot;
            surf_keypoint *kpp = A.kps + (size_t)b * A.cap + slot;
            bool dead = my_dead;
            const int n = (int)(21 * my_s);
            if (!dead && want_desc && (n < 1 || n > kMaxWindow)) {
                // n < 1 (size < 0.36): upstream builds an empty window and cv::resize raises; deleted here, as the oracle
                // does; n > kMaxWindow: unsupported window, reported through the status word.
                if (n > kMaxWindow) atomicOr(A.misc + kMiscStatus, 1);
                dead = true;
            }
            if (dead) {
                kpp->size = -1.f;
                const int pos = atomicAdd(A.ndeleted + b, 1);
                A.deleted_slots[(size_t)b * A.cap + pos] = slot;
            } else {
                kpp->angle = dir;
                if (want_desc && !A.upright) {
                    // -sin / cos of the dominant angle, correctly rounded to fp32,
                    // computed once here instead of by every thread of the descriptor CTA
                    const float th = dir * (float)(3.1415926535897932384626433832795 / 180);
                    A.sincos[(size_t)b * A.cap + slot] = make_float2(-(float)sin((double)th), (float)cos((double)th));
                }
                if (want_desc) {
                    const int cls = (n > kClassMaxN0) + (n > kClassMaxN1) + (n > kClassMaxN2);
                    const int pos = atomicAdd(A.misc + kMiscClassCount + cls, 1);
                    A.class_list[(size_t)cls * A.list_stride + pos] = b * A.cap + slot;
                }
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// K4b: descriptor.
//   k_descriptor<CLS>   windows up to 128 x 128: the whole bounding box of the rotated window is one TMA tile
//   k_descriptor_big    larger windows: the window is cut into sub-windows of at most 96 x 96 taps, each with
//                       its own TMA tile (double-buffered: the next tile lands while the current one is
//                       tapped); the window itself lives in a per-CTA global scratch buffer
// Both share the arithmetic below, statement for statement.
// ------------------------------------------------------------------------------------------------
struct AreaTab {  // one destination cell of the INTER_AREA table (same for x and y: square window)
    int start[kPatch];
    int count[kPatch];
    float a_first[kPatch], a_mid[kPatch], a_last[kPatch];
    unsigned char has_first[kPatch], has_last[kPatch];
};

struct DescShared {  // static shared state of one descriptor CTA
    AreaTab tab;
    int next;     // the following queue item, fetched while the current keypoint is tapped
    struct Geo {  // window geometry of the current keypoint: written by warp 0 (isc / int_scale by the last warp)
        int n, b, ux0, uy0, tx0a, ty0, interior, bad, isc, int_scale;
        float sd, cd;
    } geo;
    int inexact;  // a row start of this window is not a multiple of 2^-32: fp64 tap coordinates
};

// One TMA box load of the 8-bit frame b at (x, y) (x 16-byte aligned) into shared memory.
__device__ __forceinline__ void tma_load_tile(const CUtensorMap *map, uint32_t dst, int x, int y, int b, uint32_t bar, uint32_t bytes)
{
    fence_async_smem();  // earlier generic accesses of the buffer are done
    mbar_expect_tx(bar, bytes);
    tma_load_3d(map, dst, x, y, b, bar);
}

// INTER_AREA table entry of destination cell d for an n-pixel source axis (computeResizeAreaTab).
__device__ __forceinline__ void area_tab_entry(AreaTab &t, int d, int n, double scale)
{
    const double f1 = d * scale, f2 = f1 + scale;
    const double cell = fmin(scale, n - f1);
    int s1 = __double2int_ru(f1), s2 = __double2int_rd(f2);
    s2 = min(s2, n - 1);
    s1 = min(s1, s2);
    const bool hf = (s1 - f1) > 1e-3, hl = (f2 - s2) > 1e-3;
    t.start[d] = hf ? s1 - 1 : s1;
    t.count[d] = (hf ? 1 : 0) + (s2 - s1) + (hl ? 1 : 0);
    t.has_first[d] = hf;
    t.has_last[d] = hl;
    t.a_first[d] = (float)((s1 - f1) / cell);
    t.a_mid[d] = (float)(1.0 / cell);
    t.a_last[d] = (float)(fmin(fmin(f2 - s2, 1.), cell) / cell);
}

// One rotated window pixel: bilinear tap of the frame at (fx, fy), rounded to uchar; nearest with clamping
// outside the interpolable area.  `tile` holds the frame region starting at (x0, y0) with pitch TP bytes.
template <class Src>
__device__ __forceinline__ unsigned rotated_tap(double fx, double fy, int W, int H, int x0, int y0, int TP, Src src)
{
    const double kMagic = 6755399441055744.0;  // 1.5 * 2^52: floor via round-down addition
    const double mx = __dadd_rd(fx, kMagic), my = __dadd_rd(fy, kMagic);
    const int ix = __double2loint(mx), iy = __double2loint(my);
    if ((unsigned)ix < (unsigned)(W - 1) && (unsigned)iy < (unsigned)(H - 1)) {
        const float a = (float)(fx - (mx - kMagic)), bb = (float)(fy - (my - kMagic));
        const int q = (iy - y0) * TP + (ix - x0);
        const float na = 1.f - a, nb = 1.f - bb;
        const float p00 = u8f(src(q)), p01 = u8f(src(q + 1));
        const float p10 = u8f(src(q + TP)), p11 = u8f(src(q + TP + 1));
        return round_u8(p00 * na * nb + p01 * a * nb + p10 * na * bb + p11 * a * bb);
    }
    const int x = min(max(__double2int_rn(fx), 0), W - 1);
    const int y = min(max(__double2int_rn(fy), 0), H - 1);
    return src((y - y0) * TP + (x - x0));
}

// ------------------------------------------------------------------------------------------------
// Fixed-point taps.  The reference walks a window row as `double pixel_x = start_x; pixel_x += cos_dir` with
// float start_x / cos_dir.  When both are multiples of 2^-32 (|v| >= 2^-9 or v == 0: every keypoint except
// the ~0.3 % whose direction is within 0.11 degrees of an axis) that sum is exact in fp64 AND in a 64-bit
// integer with 32 fraction bits.  In the integer form floor() is the high word, and the fraction's
// conversion to float (one round-to-nearest of a 32-bit integer) is the reference's `(float)(pixel_x - ix)`
// scaled by 2^32.  The scale is a power of two, so every product and sum below rounds exactly as the
// reference's; it is removed by the final fused multiply-add with 1.5 * 2^23 (cvRound to the low byte).
// Cost per tap: 4 integer adds + 2 conversions instead of 8 fp64 additions and 2 fp64->fp32 conversions.
// Shared memory is addressed with raw 32-bit shared-window addresses (ld/st.shared through inline PTX):
// with generic pointers into the dynamic allocation the compiler re-derives the window base inside the loop.
// ------------------------------------------------------------------------------------------------
template <int OFF>
__device__ __forceinline__ uint32_t lds_u8(uint32_t addr)
{
    uint32_t v;
    asm volatile("ld.shared.u8 %0, [%1+%2];" : "=r"(v) : "r"(addr), "n"(OFF));
    return v;
}
__device__ __forceinline__ float lds_f32(uint32_t addr)
{
    float v;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ void sts_u8(uint32_t addr, uint32_t v) { asm volatile("st.shared.u8 [%0], %1;" ::"r"(addr), "r"(v)); }
__device__ __forceinline__ void sts_f32(uint32_t addr, float v) { asm volatile("st.shared.f32 [%0], %1;" ::"r"(addr), "f"(v)); }
// keeps a loop-invariant value in a register (the compiler would otherwise rematerialise it per use)
__device__ __forceinline__ uint32_t pin_reg(uint32_t v)
{
    asm volatile("" : "+r"(v));
    return v;
}

// uchar -> float as one I2FP instruction (measured: 6.43 ms per 64 frames against 6.61 ms with the two-instruction
// 2^23 trick of u8f)
__device__ __forceinline__ float tap_u8f(uint32_t v) { return __uint2float_rn(v); }

// float (multiple of 2^-32, |v| < 2^31) -> signed 32.32 fixed point, exact
__device__ __forceinline__ long long fix32(float v) { return __double2ll_rn((double)v * 4294967296.0); }
__device__ __forceinline__ bool fix32_exact(float v) { return v == 0.f || fabsf(v) >= 0.001953125f; }
constexpr float kFix32Min = 0.001953125f;  // 2^-9

// Interior tap: X's high word already carries (tile address - x0), Y's high word (-y0).
template <int TP>
__device__ __forceinline__ uint32_t tap_fixed(long long X, long long Y)
{
    const float a = __uint2float_rn((uint32_t)X), b = __uint2float_rn((uint32_t)Y);  // fractions * 2^32
    const float na = 4294967296.f - a, nb = 4294967296.f - b;
    const uint32_t q = (uint32_t)((int)(Y >> 32) * TP + (int)(X >> 32));
    const float p00 = tap_u8f(lds_u8<0>(q)), p01 = tap_u8f(lds_u8<1>(q));
    const float p10 = tap_u8f(lds_u8<TP>(q)), p11 = tap_u8f(lds_u8<TP + 1>(q));
    const float t = p00 * na * nb + p01 * a * nb + p10 * na * b + p11 * a * b;  // value * 2^64
    return __float_as_uint(__fmaf_rn(t, 5.421010862427522e-20f, 12582912.0f));  // low byte = cvRound(value)
}

// ---- two taps per thread on Blackwell's packed fp32 pipe (FMUL2 / FADD2 / FFMA2: mul / add / fma.rn.f32x2) ----
// Each lane of a packed operation is the IEEE single operation, so the results are those of tap_fixed bit for bit;
// 14 packed instructions replace 28 scalar ones per pair of taps.  One catch, measured on ptxas 12.9: a
// mul.rn.f32x2 feeding an add.rn.f32x2 is contracted into FFMA2 even under --fmad=false (the scalar forms are
// not).  The sums are therefore written as fma(x, 1, y) with the two ones read from constant memory: that is
// round(x + y) exactly, it already is a fused multiply-add (nothing left to contract) and ptxas cannot fold a
// multiplier it does not know.

// Two interior taps at once (same contract as tap_fixed for each of them).  `one` = (1.0f, 1.0f) from c_one2.
template <int TP>
__device__ __forceinline__ void tap_fixed2(long long X0, long long Y0, long long X1, long long Y1, uint64_t one, uint32_t &v0,
                                           uint32_t &v1)
{
    const uint64_t a = f2_pack(__uint2float_rn((uint32_t)X0), __uint2float_rn((uint32_t)X1));  // fractions * 2^32
    const uint64_t b = f2_pack(__uint2float_rn((uint32_t)Y0), __uint2float_rn((uint32_t)Y1));
    const uint64_t c32 = f2_pack(4294967296.f, 4294967296.f);
    const uint64_t na = f2_sub(c32, a), nb = f2_sub(c32, b);
    const uint32_t q0 = (uint32_t)((int)(Y0 >> 32) * TP + (int)(X0 >> 32));
    const uint32_t q1 = (uint32_t)((int)(Y1 >> 32) * TP + (int)(X1 >> 32));
    const uint64_t p00 = f2_pack(tap_u8f(lds_u8<0>(q0)), tap_u8f(lds_u8<0>(q1)));
    const uint64_t p01 = f2_pack(tap_u8f(lds_u8<1>(q0)), tap_u8f(lds_u8<1>(q1)));
    const uint64_t p10 = f2_pack(tap_u8f(lds_u8<TP>(q0)), tap_u8f(lds_u8<TP>(q1)));
    const uint64_t p11 = f2_pack(tap_u8f(lds_u8<TP + 1>(q0)), tap_u8f(lds_u8<TP + 1>(q1)));
    // p00 * na * nb + p01 * a * nb + p10 * na * b + p11 * a * b, left to right, every product and sum rounded
    uint64_t t = f2_mul(f2_mul(p00, na), nb);
    t = f2_fma(t, one, f2_mul(f2_mul(p01, a), nb));
    t = f2_fma(t, one, f2_mul(f2_mul(p10, na), b));
    t = f2_fma(t, one, f2_mul(f2_mul(p11, a), b));                                             // value * 2^64
    t = f2_fma(t, f2_pack(5.421010862427522e-20f, 5.421010862427522e-20f), f2_pack(12582912.0f, 12582912.0f));
    float r0, r1;
    f2_unpack(t, r0, r1);
    v0 = __float_as_uint(r0);  // low byte = cvRound(value)
    v1 = __float_as_uint(r1);
}

// Tap of a window that may leave the interpolable area: X, Y are the true coordinates.
template <int TP>
__device__ __forceinline__ uint32_t tap_fixed_checked(long long X, long long Y, int W1, int H1, int x0, int y0, uint32_t tile)
{
    const int ix = (int)(X >> 32), iy = (int)(Y >> 32);
    const uint32_t xl = (uint32_t)X, yl = (uint32_t)Y;
    if ((unsigned)ix < (unsigned)W1 && (unsigned)iy < (unsigned)H1) {
        const float a = __uint2float_rn(xl), b = __uint2float_rn(yl);
        const float na = 4294967296.f - a, nb = 4294967296.f - b;
        const uint32_t q = tile + (uint32_t)((iy - y0) * TP + (ix - x0));
        const float p00 = tap_u8f(lds_u8<0>(q)), p01 = tap_u8f(lds_u8<1>(q));
        const float p10 = tap_u8f(lds_u8<TP>(q)), p11 = tap_u8f(lds_u8<TP + 1>(q));
        const float t = p00 * na * nb + p01 * a * nb + p10 * na * b + p11 * a * b;
        return __float_as_uint(__fmaf_rn(t, 5.421010862427522e-20f, 12582912.0f));
    }
    // cvRound (nearest, ties to even) of the coordinate, clamped into the frame
    const int rx = ix + (int)(xl > 0x80000000u || (xl == 0x80000000u && (ix & 1)));
    const int ry = iy + (int)(yl > 0x80000000u || (yl == 0x80000000u && (iy & 1)));
    const int x = min(max(rx, 0), W1), y = min(max(ry, 0), H1);
    return lds_u8<0>(tile + (uint32_t)((y - y0) * TP + (x - x0)));
}

// Enlarging INTER_AREA (source axis of n < 21 pixels -> 21): cv::resize has no area mode for enlargement and runs its
// bilinear code with the area-mode coefficients (imgwarp.cpp resize(): sx = floor(d * scale), fx = (d + 1) - (sx + 1) *
// inv_scale, fx <= 0 ? 0 : fx - floor(fx)) in 8-bit fixed point: coefficients * 2048 (HResizeLinear), then
// ((b0 * (S0 >> 4) >> 16) + (b1 * (S1 >> 4) >> 16) + 2) >> 2 (VResizeLinear).  The oracle pins this against cv2.
__device__ __forceinline__ void area_up_coef(int d, int n, double scale, double inv, int &s0, int &c0, int &c1)
{
    int sx = __double2int_rd(d * scale);
    float fx = (float)((d + 1) - (sx + 1) * inv);
    fx = fx <= 0 ? 0.f : fx - floorf(fx);
    if (sx >= n - 1) {  // the last source pixel alone
        fx = 0;
        sx = n - 1;
    }
    s0 = sx;
    c0 = __float2int_rn((1.f - fx) * 2048.f);
    c1 = __float2int_rn(fx * 2048.f);
}

// Rows [i0, i1) x columns [j0, j1) of the rotated window from the shared-memory frame tile at `tile`
// (origin (x0, y0), pitch TP).  A warp covers 4 rows x 8 columns per step.  The window goes to shared memory
// (win_s, GLOBAL = false) or to the per-CTA global scratch (win_g).
template <int TP, int NW, bool CHECK, bool GLOBAL>
__device__ __forceinline__ void window_taps_fixed(int n, int i0, int i1, int j0, int j1, const float *rsx, const float *rsy,
                                                  long long cdq, long long sdq, uint32_t tile, int x0, int y0, int W, int H,
                                                  uint32_t win_s, unsigned char *win_g)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int lj = lane & 7, li = lane >> 3;
    const long long sx8 = 8 * cdq, sy8 = 8 * sdq;
    const long long xbias = CHECK ? 0ll : ((long long)(int)(tile - (uint32_t)x0)) << 32;
    const long long ybias = CHECK ? 0ll : ((long long)(-y0)) << 32;
    const long long lane_x = (long long)(j0 + lj) * cdq + xbias, lane_y = (long long)(j0 + lj) * sdq - ybias;
    const int W1 = W - 1, H1 = H - 1;
#if SURF_TAP_F32X2
    const uint64_t one2 = f2_pack(c_one2.x, c_one2.y);
#endif
    // Column groups of 16 (two taps per lane, 8 columns apart).  The last group may hold 9..15 columns: every lane
    // still has its first tap, only lanes lj < rem - 8 a second one.  That group runs the pair code once for the whole
    // warp (lanes without a second tap repeat their first one and drop it) instead of the pair code for some lanes
    // and then the single-tap code for the others.
    const int full_end = j0 + ((j1 - j0) & ~15);  // columns [j0, full_end) in whole groups
    const int rem = j1 - full_end;                // 0..15 columns left
    const bool ok2 = lj < rem - 8;
    for (int i = i0 + warp * 4 + li; i < i1; i += NW * 4) {
        long long X = fix32(rsx[i]) + lane_x, Y = fix32(rsy[i]) - lane_y;
        int j = j0 + lj;
        const int o = i * n;
        for (; j < full_end; j += 16, X += 2 * sx8, Y -= 2 * sy8) {
            uint32_t v0, v1;
            if (CHECK) {
                v0 = tap_fixed_checked<TP>(X, Y, W1, H1, x0, y0, tile);
                v1 = tap_fixed_checked<TP>(X + sx8, Y - sy8, W1, H1, x0, y0, tile);
            } else {
#if SURF_TAP_F32X2
                tap_fixed2<TP>(X, Y, X + sx8, Y - sy8, one2, v0, v1);
#else
                v0 = tap_fixed<TP>(X, Y);
                v1 = tap_fixed<TP>(X + sx8, Y - sy8);
#endif
            }
            if (GLOBAL) {
                win_g[o + j] = (unsigned char)v0;
                win_g[o + j + 8] = (unsigned char)v1;
            } else {
                sts_u8(win_s + o + j, v0);
                sts_u8(win_s + o + j + 8, v1);
            }
        }
        if (rem > 8) {
            uint32_t v0, v1 = 0;
            if (CHECK) {
                v0 = tap_fixed_checked<TP>(X, Y, W1, H1, x0, y0, tile);
                if (ok2) v1 = tap_fixed_checked<TP>(X + sx8, Y - sy8, W1, H1, x0, y0, tile);
            } else {
#if SURF_TAP_F32X2
                tap_fixed2<TP>(X, Y, ok2 ? X + sx8 : X, ok2 ? Y - sy8 : Y, one2, v0, v1);
#else
                v0 = tap_fixed<TP>(X, Y);
                if (ok2) v1 = tap_fixed<TP>(X + sx8, Y - sy8);
#endif
            }
            if (GLOBAL) {
                win_g[o + j] = (unsigned char)v0;
                if (ok2) win_g[o + j + 8] = (unsigned char)v1;
            } else {
                sts_u8(win_s + o + j, v0);
                if (ok2) sts_u8(win_s + o + j + 8, v1);
            }
        } else if (lj < rem) {
            const uint32_t v0 = CHECK ? tap_fixed_checked<TP>(X, Y, W1, H1, x0, y0, tile) : tap_fixed<TP>(X, Y);
            if (GLOBAL) win_g[o + j] = (unsigned char)v0; else sts_u8(win_s + o + j, v0);
        }
    }
}

// From the n x n window (`win`, shared or global) to the descriptor: INTER_AREA to 21 x 21, 2x2 Haar gradients,
// 4x4 sub-region sums, normalisation.  `buf` holds n x 21 floats (row sums of the fractional resize).
// One weight per source pixel of INTER_AREA cell d: first / middle / last as the table says, 0 past the end.
template <int MAXC>
__device__ __forceinline__ void cell_weights(const AreaTab &t, int d, float (&w)[MAXC])
{
    const int c = t.count[d];
    const bool hf = t.has_first[d], hl = t.has_last[d];
    const float af = t.a_first[d], am = t.a_mid[d], al = t.a_last[d];
#pragma unroll
    for (int k = 0; k < MAXC; k++) w[k] = k >= c ? 0.f : ((k == 0 && hf) ? af : ((k == c - 1 && hl) ? al : am));
}

// The window and the row sums live in shared memory (addresses win_addr / buf_addr) and no cell of the resize table
// touches more than MAXC source pixels.
template <int THREADS, int MAXC>
__device__ __forceinline__ void descriptor_tail(const DescribeArgs &A, int kidx, int n, const unsigned char *win, float *buf,
                                                uint32_t win_addr, uint32_t buf_addr, DescShared &S, bool int_scale, int isc,
                                                unsigned char *patch)
{
    const int tid = threadIdx.x;
    // Fractional scales run the reference's two passes: row sums buf[sy][dx] = sum_x win[sy][x] * alpha (fp32,
    // table order), then the column pass.  Thread = (row group, dx): the dx entry of the table stays in registers.
    static_assert(MAXC > 0, "windows beyond the staged classes go through k_descriptor_big's streaming resize");
    if (!int_scale) {
        // Fixed trip count: weights beyond the cell's pixel count are 0 and `b + p * 0 == b` exactly (the extra
        // reads stay inside the CTA's shared memory; any byte is a finite number).
        constexpr int G = THREADS / kPatch;
        constexpr int M = MAXC;
        if (tid < G * kPatch) {
            const int g = tid / kPatch, dx = tid - g * kPatch;
            if (S.tab.count[dx] > M) atomicOr(A.misc + kMiscStatus, 2);  // table bound violated: reported by the host
            float w[M];
            cell_weights<M>(S.tab, dx, w);
            uint32_t ra = win_addr + g * n + S.tab.start[dx], ba = buf_addr + (g * kPatch + dx) * 4;
            const uint32_t rstep = G * n;
#pragma unroll 2
            for (int sy = g; sy < n; sy += G, ra += rstep, ba += G * kPatch * 4) {
                float b = tap_u8f(lds_u8<0>(ra)) * w[0];
                if (M > 1) b += tap_u8f(lds_u8<1 % M>(ra)) * w[1 % M];
                if (M > 2) b += tap_u8f(lds_u8<2 % M>(ra)) * w[2 % M];
                if (M > 3) b += tap_u8f(lds_u8<3 % M>(ra)) * w[3 % M];
                if (M > 4) b += tap_u8f(lds_u8<4 % M>(ra)) * w[4 % M];
                if (M > 5) b += tap_u8f(lds_u8<5 % M>(ra)) * w[5 % M];
                if (M > 6) b += tap_u8f(lds_u8<6 % M>(ra)) * w[6 % M];
                if (M > 7) b += tap_u8f(lds_u8<7 % M>(ra)) * w[7 % M];
                static_assert(M <= 8, "extend the unrolled row pass");
                sts_f32(ba, b);
            }
        }
        __syncthreads();
    }
    if (!int_scale) {
        // Column pass.  Thread = (dy, every G-th dx): the dy entry of the table -- its weights and the addresses of its
        // source rows -- is worked out once per thread and serves its 21 / G outputs (it used to be re-derived for every
        // output: about half of this pass's instructions).
        constexpr int G = THREADS / kPatch;
        constexpr int M = MAXC;
        if (tid < G * kPatch) {
            const int dy = tid / G, c = tid - dy * G;
            float v[M];
            cell_weights<M>(S.tab, dy, v);
            const int ys = S.tab.start[dy];
            uint32_t ra[M];
#pragma unroll
            for (int e = 0; e < M; e++)  // rows past the window are clamped: their weight is 0, the value finite
                ra[e] = buf_addr + (uint32_t)(min(ys + e, n - 1) * kPatch + c) * 4;
            for (int dx = c; dx < kPatch; dx += G) {
                float acc = 0;
#pragma unroll
                for (int e = 0; e < M; e++) acc += v[e] * lds_f32(ra[e] + (uint32_t)(dx - c) * 4);
                patch[dy * kPatch + dx] = (unsigned char)sat_u8(acc);
            }
        }
    } else {
        for (int p = tid; p < kPatchPx; p += THREADS) {
            const int dy = p / kPatch, dx = p - dy * kPatch;
            int outv;
            if (isc == 2) {
                const unsigned char *r0 = win + (2 * dy) * n + 2 * dx, *r1 = r0 + n;
                outv = (r0[0] + r0[1] + r1[0] + r1[1] + 2) >> 2;
            } else if (isc == 0) {
                // n < 21 (only keypoints handed to compute() with size < 7.9): cv::resize's fixed-point bilinear code
                // with the area-mode coefficients (see area_up_coef)
                const double inv = (double)kPatch / n, scale = 1. / inv;
                int x0, xa, xb, y0, ya, yb;
                area_up_coef(dx, n, scale, inv, x0, xa, xb);
                area_up_coef(dy, n, scale, inv, y0, ya, yb);
                const int x1 = min(x0 + 1, n - 1);
                const unsigned char *r0 = win + y0 * n, *r1 = win + min(y0 + 1, n - 1) * n;
                const int h0 = r0[x0] * xa + r0[x1] * xb, h1 = r1[x0] * xa + r1[x1] * xb;
                outv = (((ya * (h0 >> 4)) >> 16) + ((yb * (h1 >> 4)) >> 16) + 2) >> 2;
            } else {
                int acc = 0;
                for (int v = 0; v < isc; v++) {
                    const unsigned char *r0 = win + (dy * isc + v) * n + dx * isc;
                    for (int u = 0; u < isc; u++) acc += r0[u];
                }
                const float wgt = 1.f / (isc * isc);
                outv = sat_u8(acc * wgt);
            }
            patch[p] = (unsigned char)outv;
        }
    }
    __syncthreads();  // the patch (a slot of the CTA's ring) is complete
    if (A.patches) {
        unsigned char *po = A.patches + (size_t)kidx * kPatchPx;
        for (int p = tid; p < kPatchPx; p += THREADS) po[p] = patch[p];
    }
}

// ------------------------------------------------------------------------------------------------
// K4c: from the 21 x 21 patch to the descriptor, inside the window kernels.  A CTA parks the patches of its last
// kRing keypoints in shared memory; when the ring is full every warp finishes one (or two) of them, one warp per
// keypoint, no block barrier inside: lane = (sub-region, dx-or-dy half).  Every output component is one
// sequential fp32 sum over the sub-region's 5 x 5 samples in (y, x) order, as the reference accumulates it.
//   64-d : half 0 -> sum dx, sum |dx| (components 0, 2 of the region); half 1 -> sum dy, sum |dy| (1, 3)
//   128-d: half 0 -> dx and |dx| split by the sign of dy (0..3);       half 1 -> dy and |dy| split by the sign of dx (4..7)
// (Round 1 wrote the patches to a global workspace and finished them in a second kernel: 122 MB per 64-frame
// batch written and read back, and four more launches.  Finishing them in batches keeps every warp of the CTA busy
// at the same time, which finishing each keypoint's 25-step chains right after its window did not.)
// ------------------------------------------------------------------------------------------------
constexpr int kRing = 4;  // (8 patches + the 400-entry weight table in shared memory cost a CTA per SM in two classes)

template <bool EXT>
__device__ __forceinline__ void patch_descriptor_warp(const DescribeArgs &A, const unsigned char *P, int kidx)
{
    const int lane = threadIdx.x & 31;
    const int region = lane >> 1, half = lane & 1;
    const int y0 = (region >> 2) * 5, x0 = (region & 3) * 5;
    // Gaussian weights of this lane's 5 x 5 samples: the table upstream builds is the float product g[y] * g[x]
    float gx[5];
#pragma unroll
    for (int x = 0; x < 5; x++) gx[x] = c_desc_g[x0 + x];
    float a0 = 0, a1 = 0, a2 = 0, a3 = 0;
    for (int y = y0; y < y0 + 5; y++) {
        const unsigned char *r0 = P + y * kPatch + x0, *r1 = r0 + kPatch;
        const float gy = c_desc_g[y];
        int p00 = r0[0], p10 = r1[0];
#pragma unroll
        for (int x = 0; x < 5; x++) {
            const int p01 = r0[x + 1], p11 = r1[x + 1];
            const float wgt = __fmul_rn(gy, gx[x]);
            // dx = p01 - p00 + p11 - p10, dy = p10 - p00 + p11 - p01 (integers: any order of the terms is the same value)
            const int da = p11 - p00, db = p01 - p10;
            if (EXT) {
                const float tx = (da + db) * wgt, ty = (da - db) * wgt;
                const float u = half ? ty : tx, o = half ? tx : ty;
                if (o >= 0) {
                    a0 += u;
                    a1 += fabsf(u);
                } else {
                    a2 += u;
                    a3 += fabsf(u);
                }
            } else {
                const float u = (half ? da - db : da + db) * wgt;
                a0 += u;
                a1 += fabsf(u);
            }
            p00 = p01;
            p10 = p11;
        }
    }
    double mag = (double)(a0 * a0) + (double)(a1 * a1);
    if (EXT) mag += (double)(a2 * a2) + (double)(a3 * a3);
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) mag += __shfl_xor_sync(FULL_MASK, mag, d);
    const float scale = (float)(1. / (sqrt(mag) + DBL_EPSILON));
    if (EXT) {
        float4 *vo = reinterpret_cast<float4 *>(A.desc + (size_t)kidx * 128 + region * 8 + half * 4);
        *vo = make_float4(a0 * scale, a1 * scale, a2 * scale, a3 * scale);
    } else {
        float *vo = A.desc + (size_t)kidx * 64 + region * 4 + half;
        vo[0] = a0 * scale;
        vo[2] = a1 * scale;
    }
}

template <int R>
struct PatchRingT {
    __align__(16) unsigned char patch[R][kPatchStride];
    int kidx[R];
};

using PatchRing = PatchRingT<kRing>;
// the sub-window kernel parks two: its shared memory is sized to the byte for three CTAs per SM, and its keypoints are
// few and long (a flush every second keypoint costs nothing against > 16 000 taps each)
constexpr int kRingBig = 2;

// Every warp finishes the ring entries warp, warp + NW, ...  The caller guarantees the entries are visible (a block
// barrier lies between the last patch write and this call).
template <int NW, class Ring>
__device__ __forceinline__ void ring_flush(const DescribeArgs &A, const Ring &R, int pending)
{
    if (A.extended) {
        for (int w = threadIdx.x >> 5; w < pending; w += NW) patch_descriptor_warp<true>(A, R.patch[w], R.kidx[w]);
    } else {
        for (int w = threadIdx.x >> 5; w < pending; w += NW) patch_descriptor_warp<false>(A, R.patch[w], R.kidx[w]);
    }
}

// Per-class compile-time geometry.  The image tile of a class has a FIXED pitch and row count (the largest
// bounding box the class can need), so it is fetched by one TMA box load and indexed with constant strides.
template <int CLS>
struct ClassT {
    static constexpr int max_n = CLS == 0 ? kClassMaxN0 : CLS == 1 ? kClassMaxN1 : kClassMaxN2;
    // measured (describe stage, 64 frames): 128/256/256 threads 6.43 ms, 128/128/256 6.28 ms, 256/256/256 7.46 ms
    static constexpr int threads = CLS == 2 ? 256 : 128;
    // resident CTAs per SM the register allocation must allow (class 0 is register-limited: 10 x 128 threads x 48)
    static constexpr int min_ctas = CLS == 0 ? 10 : CLS == 1 ? 6 : 3;
    // bounding box side of a rotated n x n window: n (|cos| + |sin|) + 3 <= 1.4143 n + 3, plus the 1-pixel
    // margin on both sides of the closed-form box; origin aligned down / pitch rounded up to 16 bytes
    static constexpr int side = (int)(1.4143 * max_n) + 4 + 2;
    // Pitch: room for the 0..15 columns of the aligned-down origin, a multiple of 16 bytes (TMA box), and an ODD
    // multiple of 16: the bank of a tile byte is (x / 4 + y * tp / 4) mod 32, so with tp / 4 = 4 (mod 8) eight
    // consecutive rows start in eight different groups of four banks and the ~9 x 9 pixel footprint of a warp's
    // 4 x 8 rotated taps is nearly conflict-free.  Class 1 (and the sub-window kernel) had tp = 160 = 40 words:
    // rows four apart shared their banks, 91 M shared-memory bank conflicts per batch against 33 - 52 M in the
    // classes whose pitch happened to be odd (profiles/r2/describe_stage_v5_ncu_summary.txt).
    static constexpr int tp_min = (side + 15 + 15) & ~15;
    static constexpr int tp = (tp_min & 16) ? tp_min : tp_min + 16;
    static constexpr int tr = side;
    static constexpr int win_bytes = (max_n * max_n + 127) & ~127;
    static constexpr int smem = 128 + max_n * 8 + win_bytes + tp * tr;
    // most source pixels one INTER_AREA cell can touch (table maximum over every n of the class; checked on the device)
    static constexpr int max_cell = CLS == 0 ? 4 : CLS == 1 ? 6 : 8;
};

// Sub-window kernel: tiles sized for sub-windows of at most kSubSide taps per side (the class-1 tile).
constexpr int kSubSide = kClassMaxN1;
struct BigT {
    static constexpr int threads = 256;
    static constexpr int tp = ClassT<1>::tp, tr = ClassT<1>::tr;
    static constexpr int tile_bytes = tp * tr;
    static constexpr int tile_stride = (tile_bytes + 127) & ~127;  // TMA destinations are 128-byte aligned
    // one sub-window of taps (kSubSide x kSubSide bytes) and its INTER_AREA row sums (kSubSide rows x 21 cells)
    static constexpr int blk_bytes = kSubSide * kSubSide;
    static constexpr int rows_bytes = kSubSide * kPatch * 4;
    static constexpr int smem = 128 + 2 * tile_stride + kMaxWindow * 8 + blk_bytes + rows_bytes;
    static constexpr int max_sub = 96;  // sub-windows per keypoint: <= 8 row blocks x <= 11 cell groups
};

// frame-tile fallback when the frames are not 16-byte aligned (no TMA): asynchronous 4-byte copies
template <int NW, int TP>
__device__ __forceinline__ void copy_tile_fallback(const DescribeArgs &A, const uint8_t *img, int pitch, unsigned char *tile,
                                                   int tx0a, int tx1, int ty0, int th_rows)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int W = A.geo.W;
    const int wpr = min(tx1 - tx0a + 1 + 3, TP) >> 2;
    const int wfull = A.img_aligned4 ? min(wpr, (W - tx0a) >> 2) : 0;  // words fully inside the row
    for (int r = warp; r < th_rows; r += NW) {
        const uint8_t *src = img + (size_t)(ty0 + r) * pitch + tx0a;
        uint32_t *dst = reinterpret_cast<uint32_t *>(tile + r * TP);
        for (int c = lane; c < wfull; c += 32) __pipeline_memcpy_async(dst + c, src + 4 * c, 4);
        for (int c = wfull + lane; c < wpr; c += 32) {
            const int x = tx0a + 4 * c;
            uint32_t v = 0;
#pragma unroll
            for (int k = 0; k < 4; k++)
                if (x + k < W) v |= (uint32_t)src[4 * c + k] << (8 * k);
            dst[c] = v;
        }
    }
    __pipeline_commit();
    __pipeline_wait_prior(0);
}

template <int CLS>
__global__ void __launch_bounds__(ClassT<CLS>::threads, ClassT<CLS>::min_ctas) k_descriptor(const __grid_constant__ DescribeArgs A,
                                                                     const __grid_constant__ CUtensorMap imap, int use_tma)
{
    using CT = ClassT<CLS>;
    constexpr int THREADS = CT::threads, NW = THREADS / 32;
    constexpr int max_n = CT::max_n, tp = CT::tp;
    extern __shared__ __align__(128) unsigned char dyn_raw[];
    // dynamic layout (128-byte aligned by hand: TMA destinations need it):
    //   uchar tile[tr][tp] | uchar win[max_n^2] | float rsx[max_n], rsy[max_n]
    unsigned char *dyn = dyn_raw + ((128u - ((uint32_t)__cvta_generic_to_shared(dyn_raw) & 127u)) & 127u);
    unsigned char *s_tile = dyn;
    unsigned char *s_win = dyn + tp * CT::tr;
    float *s_rsx = reinterpret_cast<float *>(s_win + CT::win_bytes);
    float *s_rsy = s_rsx + max_n;
    const uint32_t tile_addr = pin_reg(smem_u32(s_tile)), win_addr = pin_reg(smem_u32(s_win));
    __shared__ __align__(8) uint64_t s_bar;
    __shared__ DescShared S;
    __shared__ PatchRing ring;
    uint32_t tma_phase = 0;
    int pending = 0;  // patches parked in the ring (uniform across the CTA)

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int W = A.geo.W, H = A.geo.H;
    const int pitch = (int)A.img.pitch;
    const int count = A.misc[kMiscClassCount + CLS];
    const int *list = A.class_list + (size_t)CLS * A.list_stride;
    const uint32_t bar = (uint32_t)__cvta_generic_to_shared(&s_bar);
    if (tid == 0) {
        mbar_init(bar, 1);
        mbar_init_fence();
        S.next = atomicAdd(A.misc + kMiscClassHead + CLS, 1);
    }

    for (;;) {
        __syncthreads();  // everyone is done with the previous keypoint's shared state; S.next is visible
        const int work = S.next;
        if (work >= count) break;
        const int kidx = list[work];
        const surf_keypoint *kpp = A.kps + kidx;

        // ---- prologue.  Only warp 0 works out the window geometry (about 250 instructions that every warp used to
        //      repeat), issues the tile load and walks the row-start chains; the last warp builds the resize table;
        //      the others go straight to the barrier and read the result from shared memory.  Without TMA the tile
        //      copy needs every warp, so all of them take the lead role. ----
        const bool lead = warp == 0 || !use_tma;
        if (lead) {
            const float cx = kpp->x, cy = kpp->y, size = kpp->size;
            const float s = size * 1.2f / 9.0f;
            const int n = (int)(21 * s);
            const float off = -(float)(n - 1) / 2;
            const int b = kidx / A.cap;
            float sd = 1.f, cd = 0.f;
            int ux0 = 0, uy0 = 0;  // upright: integer window origin
            int tx0, tx1, ty0, ty1;
            bool interior = false;  // every tap has its four neighbours inside the frame
            if (A.upright) {
                ux0 = __float2int_rn(cx + off);
                uy0 = __float2int_rn(cy - off);
                tx0 = min(max(ux0, 0), W - 1);
                tx1 = min(max(ux0 + n - 1, 0), W - 1);
                ty1 = min(max(uy0, 0), H - 1);
                ty0 = min(max(uy0 - n + 1, 0), H - 1);
            } else {
                const float2 sc = A.sincos[kidx];
                sd = sc.x;
                cd = sc.y;
                // Bounding box of all taps from the closed-form window corners.  The true row starts come from
                // sequential fp32 additions (computed below, overlapped with the tile load) and drift from the
                // closed form by far less than the 1-pixel margin taken here.
                const float sx0 = cx + off * cd + off * sd, sy0 = cy - off * sd + off * cd;
                const float ext = (float)(n - 1);
                const float xa = sx0, xb = sx0 + ext * sd, xc2 = sx0 + ext * cd, xd = xb + ext * cd;
                const float ya = sy0, yb = sy0 + ext * cd, yc2 = sy0 - ext * sd, yd = yb - ext * sd;
                const float minx = fminf(fminf(xa, xb), fminf(xc2, xd)) - 1.f, maxx = fmaxf(fmaxf(xa, xb), fmaxf(xc2, xd)) + 1.f;
                const float miny = fminf(fminf(ya, yb), fminf(yc2, yd)) - 1.f, maxy = fmaxf(fmaxf(ya, yb), fmaxf(yc2, yd)) + 1.f;
                const int bx0 = __float2int_rd(minx), bx1 = __float2int_rd(maxx) + 1;
                const int by0 = __float2int_rd(miny), by1 = __float2int_rd(maxy) + 1;
                interior = bx0 >= 0 && by0 >= 0 && bx1 <= W - 1 && by1 <= H - 1;
                tx0 = min(max(bx0, 0), W - 1);
                tx1 = min(max(bx1, 0), W - 1);
                ty0 = min(max(by0, 0), H - 1);
                ty1 = min(max(by1, 0), H - 1);
            }
            const int tx0a = tx0 & ~15;              // tile origin aligned down to 16 bytes (TMA / vector copies)
            const int th_rows = ty1 - ty0 + 1;
            // the class tile is an upper bound by construction (ClassT); a violation would be a bug, reported
            // through the status word instead of corrupting memory
            const bool bad = tx1 - tx0a + 1 > tp || th_rows > CT::tr;
            if (tid == 0) {
                S.inexact = 0;  // set again (same warp, later) by the row-start chains
                S.geo.n = n; S.geo.b = b; S.geo.ux0 = ux0; S.geo.uy0 = uy0; S.geo.tx0a = tx0a; S.geo.ty0 = ty0;
                S.geo.interior = interior; S.geo.bad = bad; S.geo.sd = sd; S.geo.cd = cd;
                // ---- the window's bounding box -> shared memory: one TMA box load (issued now, awaited below) ----
                if (use_tma && !bad)
                    tma_load_tile(&imap, (uint32_t)__cvta_generic_to_shared(s_tile), tx0a, ty0, b, bar, (uint32_t)(tp * CT::tr));
            }
            if (!use_tma && !bad) {
                const uint8_t *img = A.img.ptr + (size_t)b * A.img.frame_stride;
                copy_tile_fallback<NW, tp>(A, img, pitch, s_tile, tx0a, tx1, ty0, th_rows);
            }
            if (!A.upright && tid < 2 && !bad) {
                // row starts advance by sequential fp32 additions (x chain on lane 0, y chain on lane 1 of the same
                // warp: one instruction stream for both); runs while the tile is in flight
                float v = tid == 0 ? cx + off * cd + off * sd : cy - off * sd + off * cd;
                const float inc = tid == 0 ? sd : cd;
                float *dst = tid == 0 ? s_rsx : s_rsy;
                // every row start must be a multiple of 2^-32 for the fixed-point taps: running minimum of |v| (a row
                // start of exactly 0 would be fine but is sent to the fp64 taps as well: one FMNMX per step)
                float lo = INFINITY;
                for (int i = 0; i < n; i++, v += inc) {
                    dst[i] = v;
                    lo = fminf(lo, fabsf(v));
                }
                if (lo < kFix32Min) S.inexact = 1;
            }
        }
        if (warp == NW - 1) {
            // ---- INTER_AREA table for this window size ----
            const float s = kpp->size * 1.2f / 9.0f;
            const int n = (int)(21 * s);
            const double inv = (double)kPatch / n;
            const double scale = 1. / inv;
            // n < 21: the window is enlarged (descriptor_tail's bilinear path), flagged as "integer scale 0"
            const bool enlarge = n < kPatch;
            const int isc = enlarge ? 0 : __double2int_rn(scale);
            const bool int_scale = enlarge || fabs(scale - isc) < DBL_EPSILON;
            if (lane == 31) {
                S.geo.isc = isc;
                S.geo.int_scale = int_scale;
            }
            if (!int_scale && lane < kPatch) area_tab_entry(S.tab, lane, n, scale);
        }
        __syncthreads();  // geometry, row starts and the resize table are visible
        const int n = S.geo.n, b = S.geo.b, ux0 = S.geo.ux0, uy0 = S.geo.uy0, tx0a = S.geo.tx0a, ty0 = S.geo.ty0;
        const bool interior = S.geo.interior != 0, int_scale = S.geo.int_scale != 0;
        const int isc = S.geo.isc;
        const float sd = S.geo.sd, cd = S.geo.cd;
        (void)b;
        if (S.geo.bad) {
            __syncthreads();  // every thread has read S.next and the geometry
            if (tid == 0) {
                atomicOr(A.misc + kMiscStatus, 2);
                S.next = atomicAdd(A.misc + kMiscClassHead + CLS, 1);
            }
            continue;
        }
        if (use_tma) {
            if (!mbar_wait_bounded(bar, tma_phase) && tid == 0) atomicOr(A.misc + kMiscStatus, 4);  // reported by the host
            tma_phase ^= 1;
        }

        // The next queue item is fetched now (every thread has read S.next: a barrier lies behind us) and published
        // after the taps, so the atomic's round trip hides under them.
        int next_item = 0;
        if (tid == THREADS - 32) next_item = atomicAdd(A.misc + kMiscClassHead + CLS, 1);

        // ---- window pixels ----
        auto SRC = [&](int o) -> unsigned { return (unsigned)s_tile[o]; };
        if (A.upright) {
            for (int i = warp; i < n; i += NW) {
                const int x = min(max(ux0 + i, 0), W - 1) - tx0a;
                for (int j = lane; j < n; j += 32) {
                    const int y = min(max(uy0 - j, 0), H - 1) - ty0;
                    s_win[i * n + j] = (unsigned char)SRC(y * tp + x);
                }
            }
        } else if (!S.inexact && fix32_exact(cd) && fix32_exact(sd)) {
            // 32.32 fixed-point coordinates (exact here); interior windows skip the per-tap frame test
            const long long cdq = fix32(cd), sdq = fix32(sd);
            if (interior)
                window_taps_fixed<tp, NW, false, false>(n, 0, n, 0, n, s_rsx, s_rsy, cdq, sdq, tile_addr, tx0a, ty0, W, H, win_addr, nullptr);
            else
                window_taps_fixed<tp, NW, true, false>(n, 0, n, 0, n, s_rsx, s_rsy, cdq, sdq, tile_addr, tx0a, ty0, W, H, win_addr, nullptr);
        } else {
            // direction within 0.11 degrees of an axis, or a row start inside (-2^-9, 2^-9): fp64 coordinates.
            // A warp covers 4 rows x 8 columns of the window.
            const int lj = lane & 7, li = lane >> 3;
            const double lane_dx = (double)lj * (double)cd, lane_dy = (double)lj * (double)sd;
            const double step_x = 8.0 * (double)cd, step_y = 8.0 * (double)sd;
            for (int i = warp * 4 + li; i < n; i += NW * 4) {
                double fx = (double)s_rsx[i] + lane_dx;  // start + j * step, exact in fp64
                double fy = (double)s_rsy[i] - lane_dy;
                unsigned char *wrow = s_win + i * n;
                for (int j = lj; j < n; j += 8, fx += step_x, fy -= step_y)
                    wrow[j] = (unsigned char)rotated_tap(fx, fy, W, H, tx0a, ty0, tp, SRC);
            }
        }
        if (tid == THREADS - 32) S.next = next_item;
        __syncthreads();
        // the image tile is no longer needed: its space holds the row sums of the resize
        if (tid == 0) ring.kidx[pending] = kidx;
        descriptor_tail<THREADS, CT::max_cell>(A, kidx, n, s_win, reinterpret_cast<float *>(s_tile), win_addr, tile_addr, S, int_scale, isc,
                                               ring.patch[pending]);
        if (A.desc && ++pending == kRing) {  // (the tail ended with a block barrier: the ring is visible)
            ring_flush<NW>(A, ring, pending);
            pending = 0;
        }
    }
    ring_flush<NW>(A, ring, pending);  // the loop left through its barrier
}

// Windows wider than 128 pixels (keypoint sizes above 46).  The window is never materialised: it is cut into
// sub-windows of at most kSubSide x kSubSide taps -- row blocks of equal height, column groups that hold whole cells of
// the 21-cell INTER_AREA table (neighbouring groups share the one boundary pixel two cells share) -- and every sub-window
// goes  TMA tile -> taps (shared memory) -> row sums of its cells (shared memory).  When a row block is complete its row
// sums are folded into the 21 x 21 column accumulators, in source-row order as the reference's vertical pass adds them
// (two cells per thread, in registers across the row blocks).  Round 1 kept the whole window and its row sums in a
// per-CTA global scratch: 52 MB written and 24 MB read back per 64-frame batch, a generic resize pass over global
// memory, and 390 MB of workspace.
__global__ void __launch_bounds__(BigT::threads, 3) k_descriptor_big(const __grid_constant__ DescribeArgs A,
                                                                  const __grid_constant__ CUtensorMap imap, int use_tma)
{
    constexpr int THREADS = BigT::threads, NW = THREADS / 32;
    constexpr int tp = BigT::tp, tr = BigT::tr;
    constexpr int CLS = kNumClasses - 1;
    extern __shared__ __align__(128) unsigned char dyn_raw[];
    unsigned char *dyn = dyn_raw + ((128u - ((uint32_t)__cvta_generic_to_shared(dyn_raw) & 127u)) & 127u);
    unsigned char *s_tile0 = dyn;                       // two tile buffers
    float *s_rsx = reinterpret_cast<float *>(dyn + 2 * BigT::tile_stride);
    float *s_rsy = s_rsx + kMaxWindow;
    unsigned char *s_blk = reinterpret_cast<unsigned char *>(s_rsy + kMaxWindow);   // taps of the current sub-window, pitch kSubSide
    float *s_rows = reinterpret_cast<float *>(s_blk + BigT::blk_bytes);             // row sums of the current row block [row][21]
    const uint32_t tile0_addr = pin_reg(smem_u32(s_tile0)), blk_addr = pin_reg(smem_u32(s_blk));
    __shared__ __align__(8) uint64_t s_bar[2];
    __shared__ DescShared S;
    __shared__ PatchRingT<kRingBig> ring;
    int pending = 0;
    __shared__ int4 s_sub[BigT::max_sub];  // per sub-window: tile origin x, y, flags
    uint32_t phase[2] = {0, 0};

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int W = A.geo.W, H = A.geo.H;
    const int pitch = (int)A.img.pitch;
    const int count = A.misc[kMiscClassCount + CLS];
    const int *list = A.class_list + (size_t)CLS * A.list_stride;
    const uint32_t bar0 = (uint32_t)__cvta_generic_to_shared(&s_bar[0]);
    if (tid == 0) {
        mbar_init(bar0, 1);
        mbar_init(bar0 + 8, 1);
        mbar_init_fence();
        S.next = atomicAdd(A.misc + kMiscClassHead + CLS, 1);
    }

    for (;;) {
        __syncthreads();  // S.next is visible
        const int work = S.next;
        if (work >= count) break;
        if (tid == 0) S.inexact = 0;
        const int kidx = list[work];
        const int b = kidx / A.cap;
        const surf_keypoint *kpp = A.kps + kidx;
        const float cx = kpp->x, cy = kpp->y, size = kpp->size;
        const uint8_t *img = A.img.ptr + (size_t)b * A.img.frame_stride;
        const float s = size * 1.2f / 9.0f;
        const int n = (int)(21 * s);
        const float off = -(float)(n - 1) / 2;

        float sd = 1.f, cd = 0.f;
        int ux0 = 0, uy0 = 0;
        if (A.upright) {
            ux0 = __float2int_rn(cx + off);
            uy0 = __float2int_rn(cy - off);
        } else {
            const float2 sc = A.sincos[kidx];
            sd = sc.x;
            cd = sc.y;
            if (tid < 2) {  // x chain on lane 0, y chain on lane 1
                float v = tid == 0 ? cx + off * cd + off * sd : cy - off * sd + off * cd;
                const float inc = tid == 0 ? sd : cd;
                float *dst = tid == 0 ? s_rsx : s_rsy;
                float lo = INFINITY;
                for (int i = 0; i < n; i++, v += inc) {
                    dst[i] = v;
                    lo = fminf(lo, fabsf(v));
                }
                if (lo < kFix32Min) S.inexact = 1;
            }
        }
        const double inv = (double)kPatch / n;
        const double scale = 1. / inv;
        const int isc = __double2int_rn(scale);
        const bool int_scale = fabs(scale - isc) < DBL_EPSILON;
        if (!int_scale && warp == NW - 1 && lane < kPatch) area_tab_entry(S.tab, lane, n, scale);
        __syncthreads();  // row starts and the resize table are visible
        int next_item = 0;    // fetched under the taps, published after them (see k_descriptor)
        if (tid == THREADS - 32) next_item = atomicAdd(A.misc + kMiscClassHead + CLS, 1);

        // ---- sub-window grid: qy row blocks of sh rows, qx column groups of cpg cells ----
        const int qy = (n + kSubSide - 1) / kSubSide;
        const int sh = (n + qy - 1) / qy;
        // a group of c cells spans at most c * scale + 2 source pixels (integer scales: exactly c * isc)
        const int cpg = int_scale ? max(1, kSubSide / isc) : max(1, (int)((kSubSide - 2) / scale));
        const int qx = (kPatch + cpg - 1) / cpg;
        const int n_sub = qy * qx;
        auto group_cols = [&](int gj, int &d0, int &d1, int &j0, int &j1) {  // cells [d0, d1) = source columns [j0, j1)
            d0 = gj * cpg;
            d1 = min(kPatch, d0 + cpg);
            if (int_scale) {
                j0 = d0 * isc;
                j1 = d1 * isc;
            } else {
                j0 = S.tab.start[d0];
                j1 = S.tab.start[d1 - 1] + S.tab.count[d1 - 1];
            }
        };
        const bool fixed_ok = !A.upright && !S.inexact && fix32_exact(cd) && fix32_exact(sd);
        const long long cdq = fixed_ok ? fix32(cd) : 0, sdq = fixed_ok ? fix32(sd) : 0;
        // every sub-window's tile origin, once per keypoint: thread t works out sub-window t (fp64 extremes of its taps)
        if (tid < n_sub && tid < BigT::max_sub) {
            const int si = tid / qx, gj = tid - si * qx;
            const int i0 = si * sh, i1 = min(n, i0 + sh);
            int d0, d1, j0, j1;
            group_cols(gj, d0, d1, j0, j1);
            int tx0, tx1, ty0, ty1;
            bool inside = false;
            if (A.upright) {
                // window pixel (i, j) = frame (clamp(ux0 + i), clamp(uy0 - j))
                tx0 = min(max(ux0 + i0, 0), W - 1);
                tx1 = min(max(ux0 + i1 - 1, 0), W - 1);
                ty1 = min(max(uy0 - j0, 0), H - 1);
                ty0 = min(max(uy0 - (j1 - 1), 0), H - 1);
            } else {
                const double xa = s_rsx[i0], xb = s_rsx[i1 - 1], ya = s_rsy[i0], yb = s_rsy[i1 - 1];
                const double ex0 = (double)j0 * (double)cd, ex1 = (double)(j1 - 1) * (double)cd;
                const double ey0 = (double)j0 * (double)sd, ey1 = (double)(j1 - 1) * (double)sd;
                const double minx = fmin(fmin(xa + ex0, xa + ex1), fmin(xb + ex0, xb + ex1));
                const double maxx = fmax(fmax(xa + ex0, xa + ex1), fmax(xb + ex0, xb + ex1));
                const double miny = fmin(fmin(ya - ey0, ya - ey1), fmin(yb - ey0, yb - ey1));
                const double maxy = fmax(fmax(ya - ey0, ya - ey1), fmax(yb - ey0, yb - ey1));
                // exact extremes of the taps: floor(min) >= 0 and floor(max) + 1 <= limit means every tap has its
                // four neighbours inside the frame
                const int bx0 = __double2int_rd(minx), bx1 = __double2int_rd(maxx) + 1;
                const int by0 = __double2int_rd(miny), by1 = __double2int_rd(maxy) + 1;
                inside = bx0 >= 0 && by0 >= 0 && bx1 <= W - 1 && by1 <= H - 1;
                tx0 = min(max(bx0, 0), W - 1);
                tx1 = min(max(bx1, 0), W - 1);
                ty0 = min(max(by0, 0), H - 1);
                ty1 = min(max(by1, 0), H - 1);
            }
            const int tx0a = tx0 & ~15;
            const bool fits = (tx1 - tx0a + 1 <= tp) && (ty1 - ty0 + 1 <= tr) && (j1 - j0 <= kSubSide) && (i1 - i0 <= kSubSide);
            s_sub[tid] = make_int4(tx0a, ty0, (fits ? 1 : 0) | (inside ? 2 : 0), 0);
        }
        __syncthreads();
        const bool too_many = n_sub > BigT::max_sub;  // cannot happen for n <= kMaxWindow; reported, nothing loaded
        bool bad = too_many;
        // the two cells of the 21 x 21 output this thread owns: p0 = tid, p1 = tid + THREADS
        float acc0 = 0.f, acc1 = 0.f;
        int iacc0 = 0, iacc1 = 0;
        if (use_tma && tid == 0 && !too_many)  // prologue: first tile
            tma_load_tile(&imap, (uint32_t)__cvta_generic_to_shared(s_tile0), s_sub[0].x, s_sub[0].y, b, bar0, BigT::tile_bytes);
        // (a sub-window that does not fit its tile is skipped but the loop goes on: every tile that was requested is
        // awaited, or the barrier phases of the next keypoint would be off by one)
        for (int sub = 0; sub < (too_many ? 0 : n_sub); sub++) {
            const int cur = sub & 1;
            unsigned char *tile = s_tile0 + cur * BigT::tile_stride;
            const int si = sub / qx, gj = sub - si * qx;
            const int i0 = si * sh, i1 = min(n, i0 + sh);
            int d0, d1, j0, j1;
            group_cols(gj, d0, d1, j0, j1);
            const int4 sb = s_sub[sub];
            const int tx0a = sb.x, ty0 = sb.y;
            const bool fits = (sb.z & 1) != 0, interior = (sb.z & 2) != 0;
            bad = bad || !fits;
            if (use_tma) {
                if (tid == 0 && sub + 1 < n_sub)  // prefetch the next sub-window's tile into the other buffer
                    tma_load_tile(&imap, (uint32_t)__cvta_generic_to_shared(s_tile0 + (cur ^ 1) * BigT::tile_stride), s_sub[sub + 1].x,
                                  s_sub[sub + 1].y, b, bar0 + 8 * (cur ^ 1), BigT::tile_bytes);
                if (!mbar_wait_bounded(bar0 + 8 * cur, phase[cur]) && tid == 0) atomicOr(A.misc + kMiscStatus, 4);
                phase[cur] ^= 1;
            } else {
                int tx1 = min(tx0a + tp - 1, W - 1);
                copy_tile_fallback<NW, tp>(A, img, pitch, tile, tx0a, tx1, ty0, min(tr, H - ty0));
                __syncthreads();
            }
            // ---- taps of rows [i0, i1) x columns [j0, j1) -> s_blk[(i - i0) * kSubSide + (j - j0)] ----
            auto SRC = [&](int o) -> unsigned { return (unsigned)tile[o]; };
            if (fits) {
                if (A.upright) {
                    for (int i = i0 + warp; i < i1; i += NW) {
                        const int x = min(max(ux0 + i, 0), W - 1) - tx0a;
                        for (int j = j0 + lane; j < j1; j += 32) {
                            const int y = min(max(uy0 - j, 0), H - 1) - ty0;
                            s_blk[(i - i0) * kSubSide + (j - j0)] = (unsigned char)SRC(y * tp + x);
                        }
                    }
                } else if (fixed_ok) {
                    const uint32_t tile_addr = tile0_addr + cur * BigT::tile_stride;
                    // the tap routine addresses its output as base + i * pitch + j: bias the base by the block origin
                    const uint32_t out_addr = blk_addr - (uint32_t)(i0 * kSubSide + j0);
                    if (interior)
                        window_taps_fixed<tp, NW, false, false>(kSubSide, i0, i1, j0, j1, s_rsx, s_rsy, cdq, sdq, tile_addr, tx0a, ty0, W, H, out_addr, nullptr);
                    else
                        window_taps_fixed<tp, NW, true, false>(kSubSide, i0, i1, j0, j1, s_rsx, s_rsy, cdq, sdq, tile_addr, tx0a, ty0, W, H, out_addr, nullptr);
                } else {
                    const int lj = lane & 7, li = lane >> 3;  // 4 rows x 8 columns per warp
                    const double lane_dx = (double)(j0 + lj) * (double)cd, lane_dy = (double)(j0 + lj) * (double)sd;
                    const double step_x = 8.0 * (double)cd, step_y = 8.0 * (double)sd;
                    for (int i = i0 + warp * 4 + li; i < i1; i += NW * 4) {
                        double fx = (double)s_rsx[i] + lane_dx;
                        double fy = (double)s_rsy[i] - lane_dy;
                        unsigned char *wrow = s_blk + (i - i0) * kSubSide - j0;
                        for (int j = j0 + lj; j < j1; j += 8, fx += step_x, fy -= step_y)
                            wrow[j] = (unsigned char)rotated_tap(fx, fy, W, H, tx0a, ty0, tp, SRC);
                    }
                }
            }
            __syncthreads();  // the sub-window's taps are complete; this tile buffer may be refilled (two iterations ahead)
            // ---- row sums of the cells [d0, d1) over the block's rows (the reference's horizontal pass) ----
            const int nr = i1 - i0, nc = d1 - d0;
            const unsigned nc_inv = (65536u + nc - 1) / nc;  // t / nc == (t * nc_inv) >> 16 for t <= 96 * 21 (checked exhaustively)
            for (int t = tid; t < nr * nc; t += THREADS) {
                const int r = (int)(((unsigned)t * nc_inv) >> 16), d = d0 + (t - r * nc);
                if (int_scale) {
                    const unsigned char *row = s_blk + r * kSubSide + (d * isc - j0);
                    int a = 0;
                    for (int u = 0; u < isc; u++) a += row[u];
                    reinterpret_cast<int *>(s_rows)[r * kPatch + d] = a;
                } else {
                    const int xs = S.tab.start[d], xc = S.tab.count[d];
                    const bool xhf = S.tab.has_first[d], xhl = S.tab.has_last[d];
                    const float xaf = S.tab.a_first[d], xam = S.tab.a_mid[d], xal = S.tab.a_last[d];
                    const unsigned char *row = s_blk + r * kSubSide + (xs - j0);
                    const int fend = xc - (xhl ? 1 : 0);
                    float bsum = 0;
                    int f = 0;
                    if (xhf) {
                        bsum += tap_u8f(row[0]) * xaf;
                        f = 1;
                    }
#pragma unroll 4
                    for (; f < fend; f++) bsum += tap_u8f(row[f]) * xam;
                    if (xhl) bsum += tap_u8f(row[xc - 1]) * xal;
                    s_rows[r * kPatch + d] = bsum;
                }
            }
            if (gj == qx - 1) {
                // ---- the row block is complete: fold its rows into the column accumulators, rows in increasing order ----
                __syncthreads();
#pragma unroll
                for (int h = 0; h < 2; h++) {
                    const int p = tid + h * THREADS;
                    if (p < kPatchPx) {
                        const int dy = p / kPatch, dx = p - dy * kPatch;
                        if (int_scale) {
                            int a = h ? iacc1 : iacc0;
                            const int lo = max(dy * isc, i0), hi = min(dy * isc + isc, i1);
                            for (int sy = lo; sy < hi; sy++) a += reinterpret_cast<const int *>(s_rows)[(sy - i0) * kPatch + dx];
                            if (h) iacc1 = a; else iacc0 = a;
                        } else {
                            const int ys = S.tab.start[dy], yc = S.tab.count[dy];
                            const bool yhf = S.tab.has_first[dy], yhl = S.tab.has_last[dy];
                            const float yaf = S.tab.a_first[dy], yam = S.tab.a_mid[dy], yal = S.tab.a_last[dy];
                            float a = h ? acc1 : acc0;
                            const int lo = max(ys, i0), hi = min(ys + yc, i1);
                            for (int sy = lo; sy < hi; sy++) {
                                const int e = sy - ys;
                                const float beta = (e == 0 && yhf) ? yaf : ((e == yc - 1 && yhl) ? yal : yam);
                                const float bv = s_rows[(sy - i0) * kPatch + dx];
                                if (e == 0) a = beta * bv; else a += beta * bv;
                            }
                            if (h) acc1 = a; else acc0 = a;
                        }
                    }
                }
            }
            __syncthreads();  // s_blk / s_rows may be overwritten by the next sub-window
        }
        if (bad && tid == 0) atomicOr(A.misc + kMiscStatus, 2);
        if (tid == THREADS - 32) S.next = next_item;
        if (tid == 0) ring.kidx[pending] = kidx;
        // ---- the 21 x 21 patch ----
        unsigned char *patch = ring.patch[pending];
#pragma unroll
        for (int h = 0; h < 2; h++) {
            const int p = tid + h * THREADS;
            if (p < kPatchPx) {
                int outv;
                if (int_scale) {
                    const float wgt = 1.f / (isc * isc);
                    outv = sat_u8((h ? iacc1 : iacc0) * wgt);
                } else {
                    outv = sat_u8(h ? acc1 : acc0);
                }
                patch[p] = (unsigned char)outv;
            }
        }
        __syncthreads();  // the patch (a slot of the CTA's ring) is complete
        if (A.patches) {
            unsigned char *po = A.patches + (size_t)kidx * kPatchPx;
            for (int p = tid; p < kPatchPx; p += THREADS) po[p] = patch[p];
        }
        if (A.desc && ++pending == kRingBig) {
            ring_flush<NW>(A, ring, pending);
            pending = 0;
        }
    }
    ring_flush<NW>(A, ring, pending);
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
namespace {

struct ClassCfg {
    int smem, grid;
};

struct DevCfg {
    bool ready = false;
    int ori_grid = 0;
    ClassCfg cls[kNumClasses];
    cudaStream_t side[kNumClasses - 1];
    cudaEvent_t fork, join[kNumClasses - 1];
};
DevCfg g_cfg[64];

template <int CLS>
void setup_class(ClassCfg &c, int sms)
{
    using CT = ClassT<CLS>;
    c.smem = CT::smem;
    cudaFuncSetAttribute(k_descriptor<CLS>, cudaFuncAttributeMaxDynamicSharedMemorySize, c.smem);
    int per_sm = 1;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_descriptor<CLS>, CT::threads, c.smem);
    if (per_sm < 1) per_sm = 1;
    c.grid = sms * per_sm;  // persistent CTAs: a multiple of the SM count
}

void setup_big(ClassCfg &c, int sms)
{
    c.smem = BigT::smem;
    cudaFuncSetAttribute(k_descriptor_big, cudaFuncAttributeMaxDynamicSharedMemorySize, c.smem);
    int per_sm = 1;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_descriptor_big, BigT::threads, c.smem);
    if (per_sm < 1) per_sm = 1;
    c.grid = sms * per_sm;
}

DevCfg &dev_cfg()
{
    int dev = 0;
    cudaGetDevice(&dev);
    DevCfg &c = g_cfg[dev & 63];
    if (!c.ready) {
        int sms = 148;
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        int per_sm = 4;
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_orient, kOriWarps * 32, 0);
        c.ori_grid = sms * (per_sm < 1 ? 1 : per_sm);
        setup_class<0>(c.cls[0], sms);
        setup_class<1>(c.cls[1], sms);
        setup_class<2>(c.cls[2], sms);
        setup_big(c.cls[3], sms);
        for (int k = 0; k < kNumClasses - 1; k++) {
            cudaStreamCreateWithFlags(&c.side[k], cudaStreamNonBlocking);
            cudaEventCreateWithFlags(&c.join[k], cudaEventDisableTiming);
        }
        cudaEventCreateWithFlags(&c.fork, cudaEventDisableTiming);
        c.ready = true;
    }
    return c;
}

}  // namespace

// TMA box (pitch bytes x rows) of the image tile used by size class `cls`
void describe_tile_box(int cls, int *pitch_bytes, int *rows)
{
    const int tp[kNumClasses] = {ClassT<0>::tp, ClassT<1>::tp, ClassT<2>::tp, BigT::tp};
    const int tr[kNumClasses] = {ClassT<0>::tr, ClassT<1>::tr, ClassT<2>::tr, BigT::tr};
    *pitch_bytes = tp[cls];
    *rows = tr[cls];
}

void launch_describe(const DescribeArgs &a, const CUtensorMap *imaps, int use_tma, cudaStream_t st, int *launches)
{
    DevCfg &c = dev_cfg();
    k_orient<<<c.ori_grid, kOriWarps * 32, 0, st>>>(a);
    *launches += 1;
    if (!a.desc && !a.patches) return;
    // The size classes are independent: fork them onto side streams so their CTAs share the GPU and the
    // tails of the short launches overlap the long ones; join back on `st`.
    cudaEventRecord(c.fork, st);
    for (int k = 0; k < kNumClasses - 1; k++) cudaStreamWaitEvent(c.side[k], c.fork, 0);
    // each class: windows -> 21 x 21 patches -> descriptors, in one kernel
    k_descriptor_big<<<c.cls[3].grid, BigT::threads, c.cls[3].smem, st>>>(a, imaps[3], use_tma);
    k_descriptor<1><<<c.cls[1].grid, ClassT<1>::threads, c.cls[1].smem, c.side[0]>>>(a, imaps[1], use_tma);
    k_descriptor<2><<<c.cls[2].grid, ClassT<2>::threads, c.cls[2].smem, c.side[1]>>>(a, imaps[2], use_tma);
    k_descriptor<0><<<c.cls[0].grid, ClassT<0>::threads, c.cls[0].smem, c.side[2]>>>(a, imaps[0], use_tma);
    for (int k = 0; k < kNumClasses - 1; k++) {
        cudaEventRecord(c.join[k], c.side[k]);
        cudaStreamWaitEvent(st, c.join[k], 0);
    }
    *launches += 4;
}

}  // namespace surfb200
