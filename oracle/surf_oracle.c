/*
 * surf_oracle.c -- CPU restatement of OpenCV 2.4.5 SURF (nonfree) + BFMatcher L2.
 * TEST INFRASTRUCTURE ONLY -- see surf_oracle.h.  PARITY UNPINNED BY THE REFERENCE (no golden
 * vectors exist in stropitek/SurfFeatures); sub-steps are pinned against cv2 4.13 in tests/.
 *
 * Every function names the upstream routine it restates (OpenCV tag 2.4.5, pinned by
 * /root/reference/SuperBuild/External_OpenCV.cmake:6-7) and the SURVEY.md Appendix-A paragraph
 * that specifies it.  Build with -ffp-contract=off: the reference build is SSE-only
 * (External_OpenCV.cmake:40-45), so there is no FMA contraction on its side.
 */
#include "surf_oracle.h"

#include <float.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

enum {
    ORI_RADIUS = 6,
    ORI_WIN = 60,
    ORI_SEARCH_INC = 5,
    PATCH_SZ = 20,
    PATCH_W = PATCH_SZ + 1,
    HAAR_SIZE0 = 9,
    HAAR_SIZE_INC = 6,
    ORI_BOUND = (2 * ORI_RADIUS + 1) * (2 * ORI_RADIUS + 1)
};
static const double ORI_SIGMA = 2.5;
static const double DESC_SIGMA = 3.3;

/* cvRound: round-half-to-even of a double (SSE2 cvtsd2si).  SURVEY A7. */
static inline int rnd(double v) { return (int)lrint(v); }
static inline int ifloor(double v) { return (int)floor(v); }
static inline int iceil(double v) { return (int)ceil(v); }
static inline int imin(int a, int b) { return a < b ? a : b; }
static inline int imax(int a, int b) { return a > b ? a : b; }

int orc_max_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* ------------------------------------------------------------------------------------------
 * Stage 1.  cv::integral(img, sum, CV_32S)  [imgproc/sumpixels.cpp]; SURVEY A2.2, row a4.
 * ------------------------------------------------------------------------------------------ */
void orc_integral(const uint8_t *img, int H, int W, int pitch, int32_t *sum, int clamp01)
{
    const int SW = W + 1;
    memset(sum, 0, sizeof(int32_t) * (size_t)SW);
    for (int y = 0; y < H; y++) {
        const uint8_t *row = img + (size_t)y * pitch;
        const int32_t *above = sum + (size_t)y * SW;
        int32_t *out = sum + (size_t)(y + 1) * SW;
        int32_t run = 0;
        out[0] = 0;
        for (int x = 0; x < W; x++) {
            int v = row[x];
            if (clamp01 && v > 1) v = 1;
            run += v;
            out[x + 1] = above[x + 1] + run;
        }
    }
}

/* ------------------------------------------------------------------------------------------
 * Box patterns.  resizeHaarPattern / calcHaarPattern [nonfree/surf.cpp]; SURVEY A3.2.
 * A box is four offsets into the integral plus an fp32 weight 1/area (signed).
 * ------------------------------------------------------------------------------------------ */
typedef struct {
    int p0, p1, p2, p3; /* (y1,x1) (y2,x1) (y1,x2) (y2,x2) as linear offsets */
    float w;
} box_t;

static void scale_pattern(const int (*src)[5], box_t *dst, int n, int old_size, int new_size, int stride)
{
    float ratio = (float)new_size / old_size;
    for (int k = 0; k < n; k++) {
        int x1 = rnd(ratio * src[k][0]);
        int y1 = rnd(ratio * src[k][1]);
        int x2 = rnd(ratio * src[k][2]);
        int y2 = rnd(ratio * src[k][3]);
        dst[k].p0 = y1 * stride + x1;
        dst[k].p1 = y2 * stride + x1;
        dst[k].p2 = y1 * stride + x2;
        dst[k].p3 = y2 * stride + x2;
        dst[k].w = src[k][4] / ((float)(x2 - x1) * (y2 - y1));
    }
}

/* int corner sum -> fp32 product with the weight -> accumulated in fp64 -> cast to fp32. */
static inline float box_response(const int32_t *origin, const box_t *f, int n)
{
    double d = 0;
    for (int k = 0; k < n; k++)
        d += (origin[f[k].p0] + origin[f[k].p3] - origin[f[k].p1] - origin[f[k].p2]) * f[k].w;
    return (float)d;
}

static const int PAT_DXX[3][5] = {{0, 2, 3, 7, 1}, {3, 2, 6, 7, -2}, {6, 2, 9, 7, 1}};
static const int PAT_DYY[3][5] = {{2, 0, 7, 3, 1}, {2, 3, 7, 6, -2}, {2, 6, 7, 9, 1}};
static const int PAT_DXY[4][5] = {{1, 1, 4, 4, 1}, {5, 1, 8, 4, -1}, {1, 5, 4, 8, -1}, {5, 5, 8, 8, 1}};
static const int PAT_MASK[1][5] = {{0, 0, 9, 9, 1}};
static const int PAT_HX[2][5] = {{0, 0, 2, 4, -1}, {2, 0, 4, 4, 1}};
static const int PAT_HY[2][5] = {{0, 0, 4, 2, 1}, {0, 2, 4, 4, -1}};

/* ------------------------------------------------------------------------------------------
 * Stage 2.  calcLayerDetAndTrace [nonfree/surf.cpp]; SURVEY A3.3, row a5.
 * ------------------------------------------------------------------------------------------ */
int orc_hessian_layer(const int32_t *sum, int H, int W, int size, int step, float *det, float *trace)
{
    if (size > H || size > W) return 0;
    const int SW = W + 1;
    const int cols = W / step;
    box_t dxx[3], dyy[3], dxy[4];
    scale_pattern(PAT_DXX, dxx, 3, 9, size, SW);
    scale_pattern(PAT_DYY, dyy, 3, 9, size, SW);
    scale_pattern(PAT_DXY, dxy, 4, 9, size, SW);
    const int ni = 1 + (H - size) / step;
    const int nj = 1 + (W - size) / step;
    const int m = (size / 2) / step;
    for (int i = 0; i < ni; i++) {
        const int32_t *o = sum + (size_t)(i * step) * SW;
        float *drow = det + (size_t)(i + m) * cols + m;
        float *trow = trace ? trace + (size_t)(i + m) * cols + m : NULL;
        for (int j = 0; j < nj; j++, o += step) {
            float dx = box_response(o, dxx, 3);
            float dy = box_response(o, dyy, 3);
            float dc = box_response(o, dxy, 4);
            drow[j] = dx * dy - 0.81f * dc * dc;
            if (trow) trow[j] = dx + dy;
        }
    }
    return 1;
}

/* ------------------------------------------------------------------------------------------
 * Stage 3.  interpolateKeypoint / findMaximaInLayer / KeypointGreater [nonfree/surf.cpp];
 * Matx33f::solve 3x3 closed form [core/operations.hpp].  SURVEY A3.4, row a6.
 * ------------------------------------------------------------------------------------------ */
static int solve3_cramer(const float A[3][3], const float b[3], float x[3])
{
    float d = A[0][0] * (A[1][1] * A[2][2] - A[2][1] * A[1][2]) -
              A[0][1] * (A[1][0] * A[2][2] - A[2][0] * A[1][2]) +
              A[0][2] * (A[1][0] * A[2][1] - A[2][0] * A[1][1]);
    d = (float)(double)d;
    if (d == 0) {
        x[0] = x[1] = x[2] = 0;
        return 0;
    }
    d = 1 / d;
    x[0] = d * (b[0] * (A[1][1] * A[2][2] - A[1][2] * A[2][1]) -
                A[0][1] * (b[1] * A[2][2] - A[1][2] * b[2]) +
                A[0][2] * (b[1] * A[2][1] - A[1][1] * b[2]));
    x[1] = d * (A[0][0] * (b[1] * A[2][2] - A[1][2] * b[2]) -
                b[0] * (A[1][0] * A[2][2] - A[1][2] * A[2][0]) +
                A[0][2] * (A[1][0] * b[2] - b[1] * A[2][0]));
    x[2] = d * (A[0][0] * (A[1][1] * b[2] - b[1] * A[2][1]) -
                A[0][1] * (A[1][0] * b[2] - b[1] * A[2][0]) +
                b[0] * (A[1][0] * A[2][1] - A[1][1] * A[2][0]));
    return 1;
}

/* N[k][0..8]: 3x3 neighbourhood of layer below (0), current (1), above (2); index 4 = centre. */
static int refine_maximum(float N[3][9], int dxy, int ds, orc_keypoint *kp)
{
    float b[3] = {-(N[1][5] - N[1][3]) / 2, -(N[1][7] - N[1][1]) / 2, -(N[2][4] - N[0][4]) / 2};
    float xy = (N[1][8] - N[1][6] - N[1][2] + N[1][0]) / 4;
    float xs = (N[2][5] - N[2][3] - N[0][5] + N[0][3]) / 4;
    float ys = (N[2][7] - N[2][1] - N[0][7] + N[0][1]) / 4;
    float A[3][3] = {{N[1][3] - 2 * N[1][4] + N[1][5], xy, xs},
                     {xy, N[1][1] - 2 * N[1][4] + N[1][7], ys},
                     {xs, ys, N[0][4] - 2 * N[1][4] + N[2][4]}};
    float x[3];
    solve3_cramer(A, b, x);
    int ok = (x[0] != 0 || x[1] != 0 || x[2] != 0) && fabsf(x[0]) <= 1 && fabsf(x[1]) <= 1 && fabsf(x[2]) <= 1;
    if (ok) {
        kp->x += x[0] * dxy;
        kp->y += x[1] * dxy;
        kp->size = (float)rnd(kp->size + x[2] * ds);
    }
    return ok;
}

typedef struct {
    orc_keypoint *v;
    int n, cap;
} kpvec;

static void kp_push(kpvec *a, const orc_keypoint *k)
{
    if (a->n == a->cap) {
        a->cap = a->cap ? a->cap * 2 : 1024;
        a->v = (orc_keypoint *)realloc(a->v, sizeof(orc_keypoint) * (size_t)a->cap);
    }
    a->v[a->n++] = *k;
}

static void find_maxima(const int32_t *msum, int H, int W, float *const *dets, float *const *traces,
                        const int *sizes, int layer, int octave, int step, float thr, kpvec *out)
{
    const int size = sizes[layer];
    const int rows = H / step, cols = W / step;
    const int margin = (sizes[layer + 1] / 2) / step + 1;
    const int SW = W + 1;
    box_t dm[1];
    if (msum) scale_pattern(PAT_MASK, dm, 1, 9, size, SW);
    const float *d0 = dets[layer - 1], *d1 = dets[layer], *d2 = dets[layer + 1];
    for (int i = margin; i < rows - margin; i++) {
        for (int j = margin; j < cols - margin; j++) {
            size_t c = (size_t)i * cols + j;
            float v = d1[c];
            if (!(v > thr)) continue;
            int si = step * (i - (size / 2) / step);
            int sj = step * (j - (size / 2) / step);
            float N[3][9];
            const float *L[3] = {d0, d1, d2};
            for (int k = 0; k < 3; k++)
                for (int r = 0; r < 3; r++)
                    for (int q = 0; q < 3; q++)
                        N[k][r * 3 + q] = L[k][c + (size_t)((r - 1) * cols) + (q - 1)];
            if (msum) {
                float mv = box_response(msum + (size_t)si * SW + sj, dm, 1);
                if (mv < 0.5) continue;
            }
            int is_max = 1;
            for (int k = 0; k < 3 && is_max; k++)
                for (int q = 0; q < 9; q++)
                    if (!(k == 1 && q == 4) && !(v > N[k][q])) {
                        is_max = 0;
                        break;
                    }
            if (!is_max) continue;
            orc_keypoint kp;
            float ci = si + (size - 1) * 0.5f;
            float cj = sj + (size - 1) * 0.5f;
            kp.x = cj;
            kp.y = ci;
            kp.size = (float)size;
            kp.angle = -1;
            kp.response = v;
            kp.octave = octave;
            float tr = traces[layer][c];
            kp.class_id = (tr > 0) - (tr < 0);
            if (refine_maximum(N, step, size - sizes[layer - 1], &kp)) kp_push(out, &kp);
        }
    }
}

/* response desc, size desc, octave desc, y desc, x asc. */
static int kp_order(const void *pa, const void *pb)
{
    const orc_keypoint *a = (const orc_keypoint *)pa, *b = (const orc_keypoint *)pb;
    if (a->response > b->response) return -1;
    if (a->response < b->response) return 1;
    if (a->size > b->size) return -1;
    if (a->size < b->size) return 1;
    if (a->octave > b->octave) return -1;
    if (a->octave < b->octave) return 1;
    if (a->y < b->y) return 1;
    if (a->y > b->y) return -1;
    if (a->x < b->x) return -1;
    if (a->x > b->x) return 1;
    return 0;
}

/* fastHessianDetector [nonfree/surf.cpp]; SURVEY A3.1. */
static kpvec fast_hessian(const int32_t *sum, const int32_t *msum, int H, int W, const orc_params *p)
{
    const int LPO = p->n_octave_layers + 2;
    const int total = LPO * p->n_octaves;
    float **dets = (float **)calloc((size_t)total, sizeof(float *));
    float **traces = (float **)calloc((size_t)total, sizeof(float *));
    int *sizes = (int *)calloc((size_t)total, sizeof(int));
    int *steps = (int *)calloc((size_t)total, sizeof(int));
    int step = 1, idx = 0;
    for (int o = 0; o < p->n_octaves; o++, step *= 2)
        for (int l = 0; l < LPO; l++, idx++) {
            size_t cnt = (size_t)(H / step) * (size_t)(W / step);
            dets[idx] = (float *)calloc(cnt ? cnt : 1, sizeof(float));
            traces[idx] = (float *)calloc(cnt ? cnt : 1, sizeof(float));
            sizes[idx] = (HAAR_SIZE0 + HAAR_SIZE_INC * l) << o;
            steps[idx] = step;
            orc_hessian_layer(sum, H, W, sizes[idx], step, dets[idx], traces[idx]);
        }
    kpvec out = {NULL, 0, 0};
    const float thr = (float)p->hessian_threshold;
    for (int o = 0; o < p->n_octaves; o++)
        for (int l = 1; l <= p->n_octave_layers; l++) {
            int base = o * LPO;
            find_maxima(msum, H, W, dets + base, traces + base, sizes + base, l, o, steps[base], thr, &out);
        }
    if (out.n > 1) qsort(out.v, (size_t)out.n, sizeof(orc_keypoint), kp_order);
    for (int i = 0; i < total; i++) {
        free(dets[i]);
        free(traces[i]);
    }
    free(dets);
    free(traces);
    free(sizes);
    free(steps);
    return out;
}

/* ------------------------------------------------------------------------------------------
 * Helpers restated from core / imgproc.
 * ------------------------------------------------------------------------------------------ */

/* getGaussianKernel(n, sigma, CV_32F) [imgproc/smooth.cpp]; closed form, SURVEY A5.3 / C-4. */
void orc_gaussian_kernel(int n, double sigma, float *out)
{
    double s2 = -0.5 / (sigma * sigma), acc = 0;
    for (int i = 0; i < n; i++) {
        double x = i - (n - 1) * 0.5;
        out[i] = (float)exp(s2 * x * x);
        acc += out[i];
    }
    acc = 1. / acc;
    for (int i = 0; i < n; i++) out[i] = (float)(out[i] * acc);
}

/* fastAtan2 [core/mathfuncs.cpp]; degrees in [0,360]; SURVEY A4. */
float orc_fast_atan2(float y, float x)
{
    const float k = (float)(180 / 3.1415926535897932384626433832795);
    const float p1 = 0.9997878412794807f * k, p3 = -0.3258083974640975f * k;
    const float p5 = 0.1555786518463281f * k, p7 = -0.04432655554792128f * k;
    float ax = fabsf(x), ay = fabsf(y), a, c, c2;
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

/* One axis of the INTER_AREA table: computeResizeAreaTab [imgproc/imgwarp.cpp]; SURVEY A5.4. */
typedef struct {
    int di, si;
    float alpha;
} area_tap;

static int area_table(int ssize, int dsize, double scale, area_tap *tab)
{
    int k = 0;
    for (int d = 0; d < dsize; d++) {
        double f1 = d * scale, f2 = f1 + scale;
        double cell = fmin(scale, ssize - f1);
        int s1 = iceil(f1), s2 = ifloor(f2);
        s2 = imin(s2, ssize - 1);
        s1 = imin(s1, s2);
        if (s1 - f1 > 1e-3) {
            tab[k].di = d;
            tab[k].si = s1 - 1;
            tab[k++].alpha = (float)((s1 - f1) / cell);
        }
        for (int s = s1; s < s2; s++) {
            tab[k].di = d;
            tab[k].si = s;
            tab[k++].alpha = (float)(1.0 / cell);
        }
        if (f2 - s2 > 1e-3) {
            tab[k].di = d;
            tab[k].si = s2;
            tab[k++].alpha = (float)(fmin(fmin(f2 - s2, 1.), cell) / cell);
        }
    }
    return k;
}

static inline uint8_t sat_u8(float v)
{
    int r = rnd(v);
    return (uint8_t)(r < 0 ? 0 : r > 255 ? 255 : r);
}

/* cv::resize(win, patch, Size(21,21), 0, 0, INTER_AREA) for an n x n 8-bit window, n >= 21.
 * Three code paths of imgwarp.cpp: 2x integer (ResizeAreaFastVec), k x integer (ResizeAreaFast_),
 * fractional table (ResizeArea_).  SURVEY A5.4. */
/* n < 21 (keypoints smaller than 7.9 px handed to compute()): cv::resize does not have a true area mode for
 * enlargement; INTER_AREA then runs the bilinear code with its own coefficients (`area_mode` in imgwarp.cpp
 * resize(): sx = floor(dx * scale), fx = (dx + 1) - (sx + 1) * inv_scale, fx <= 0 ? 0 : fx - floor(fx)), the 8-bit
 * fixed-point kernels HResizeLinear (coefficients * 2048, int rows) and VResizeLinear
 * ((b0 * (S0 >> 4) >> 16) + (b1 * (S1 >> 4) >> 16) + 2) >> 2.  Pinned against cv2 4.13 for every n = 1..20. */
static void area_up_coef(int d, int n, double scale, double inv, int *s0, int *c0, int *c1)
{
    int sx = ifloor(d * scale);
    float fx = (float)((d + 1) - (sx + 1) * inv);
    fx = fx <= 0 ? 0.f : fx - (float)ifloor(fx);
    if (sx >= n - 1) { /* sx + ksize/2 >= ssize: the last source pixel alone */
        fx = 0;
        sx = n - 1;
    }
    *s0 = sx;
    *c0 = rnd((1.f - fx) * 2048);
    *c1 = rnd(fx * 2048);
}

static void resize_area_up_21(const uint8_t *win, int n, uint8_t *patch)
{
    const double inv = (double)PATCH_W / n;
    const double scale = 1. / inv;
    int xs[PATCH_W], xa[PATCH_W], xb[PATCH_W];
    for (int dx = 0; dx < PATCH_W; dx++) area_up_coef(dx, n, scale, inv, &xs[dx], &xa[dx], &xb[dx]);
    for (int dy = 0; dy < PATCH_W; dy++) {
        int sy, b0, b1;
        area_up_coef(dy, n, scale, inv, &sy, &b0, &b1);
        const uint8_t *r0 = win + (size_t)sy * n;
        const uint8_t *r1 = win + (size_t)(sy + 1 < n ? sy + 1 : n - 1) * n;
        for (int dx = 0; dx < PATCH_W; dx++) {
            const int x0 = xs[dx], x1 = x0 + 1 < n ? x0 + 1 : x0; /* weight of x1 is 0 when it is clamped */
            const int h0 = r0[x0] * xa[dx] + r0[x1] * xb[dx];
            const int h1 = r1[x0] * xa[dx] + r1[x1] * xb[dx];
            patch[dy * PATCH_W + dx] = (uint8_t)((((b0 * (h0 >> 4)) >> 16) + ((b1 * (h1 >> 4)) >> 16) + 2) >> 2);
        }
    }
}

int orc_resize_area_21(const uint8_t *win, int n, uint8_t *patch)
{
    if (n < 1) return -1;
    if (n < PATCH_W) {
        resize_area_up_21(win, n, patch);
        return 0;
    }
    const double inv = (double)PATCH_W / n;
    const double scale = 1. / inv;
    const int isc = (int)lrint(scale);
    if (fabs(scale - isc) < DBL_EPSILON) {
        if (isc == 2) {
            for (int y = 0; y < PATCH_W; y++)
                for (int x = 0; x < PATCH_W; x++) {
                    const uint8_t *s = win + (size_t)(2 * y) * n + 2 * x;
                    patch[y * PATCH_W + x] = (uint8_t)((s[0] + s[1] + s[n] + s[n + 1] + 2) >> 2);
                }
        } else {
            const float w = 1.f / (isc * isc);
            for (int y = 0; y < PATCH_W; y++)
                for (int x = 0; x < PATCH_W; x++) {
                    int acc = 0;
                    for (int v = 0; v < isc; v++)
                        for (int u = 0; u < isc; u++) acc += win[(size_t)(y * isc + v) * n + x * isc + u];
                    patch[y * PATCH_W + x] = sat_u8(acc * w);
                }
        }
        return 0;
    }
    area_tap *tab = (area_tap *)malloc(sizeof(area_tap) * (size_t)(2 * n + 2 * PATCH_W));
    int nt = area_table(n, PATCH_W, scale, tab);
    float buf[PATCH_W], acc[PATCH_W];
    int prev = -1;
    for (int j = 0; j < nt; j++) {
        const float beta = tab[j].alpha;
        const int dy = tab[j].di;
        const uint8_t *S = win + (size_t)tab[j].si * n;
        for (int x = 0; x < PATCH_W; x++) buf[x] = 0;
        for (int k = 0; k < nt; k++) buf[tab[k].di] += S[tab[k].si] * tab[k].alpha;
        if (dy != prev) {
            if (prev >= 0)
                for (int x = 0; x < PATCH_W; x++) patch[prev * PATCH_W + x] = sat_u8(acc[x]);
            for (int x = 0; x < PATCH_W; x++) acc[x] = beta * buf[x];
            prev = dy;
        } else {
            for (int x = 0; x < PATCH_W; x++) acc[x] += beta * buf[x];
        }
    }
    if (prev >= 0)
        for (int x = 0; x < PATCH_W; x++) patch[prev * PATCH_W + x] = sat_u8(acc[x]);
    free(tab);
    return 0;
}

/* ------------------------------------------------------------------------------------------
 * Stage 4.  SURFInvoker [nonfree/surf.cpp]; SURVEY A4 (orientation) and A5 (descriptor).
 * sin/cos of the dominant angle are taken correctly rounded to fp32 ((float)sin((double)t)):
 * upstream calls the platform libm's sinf/cosf, whose last bit differs between platforms, so this
 * is the platform-independent reading of the same step (numerics).
 * ------------------------------------------------------------------------------------------ */
typedef struct {
    int npts;
    int px[ORI_BOUND], py[ORI_BOUND];
    float pw[ORI_BOUND];
    float dw[PATCH_SZ * PATCH_SZ];
} surf_tables;

/* Named reading choices of SURVEY Appendix C as run-time switches, so that a genuine OpenCV -- if one with nonfree
 * SURF ever is at hand (tests/test_opencv_probe.py) -- can arbitrate between the readings.  All zero = the 2.4.5
 * reading the CUDA path implements.  Process-wide; test infrastructure only.
 *   angle_pre244   C-2: 0 = 2.4.4+ convention, kp.angle = fastAtan2(-sum_y, sum_x), sin_dir = -sin(dir), upright 270;
 *                       1 = the convention before 2.4.4: kp.angle = fastAtan2(sum_y, sum_x), sin_dir = sin(dir),
 *                       upright 90 (the same physical direction, reported mirrored)
 *   libm_sincos    0 = sin / cos of the dominant angle correctly rounded to fp32; 1 = this platform's sinf / cosf
 *                       (what upstream literally calls)
 *   desc_gauss     C-4: when set, the 20 taps of the descriptor's Gaussian (PATCH_SZ, sigma 3.3) are taken from here
 *                       instead of the closed form (e.g. the installed cv2.getGaussianKernel(20, 3.3, CV_32F), whose two
 *                       centre taps differ by ~1 % since the bit-exact even-length kernel of OpenCV 4.x) */
static struct {
    int angle_pre244, libm_sincos, have_desc_gauss;
    float desc_gauss[PATCH_SZ];
} g_sw;

void orc_set_switches(int angle_pre244, int libm_sincos, const float *desc_gauss20)
{
    g_sw.angle_pre244 = angle_pre244;
    g_sw.libm_sincos = libm_sincos;
    g_sw.have_desc_gauss = desc_gauss20 != 0;
    if (desc_gauss20)
        for (int i = 0; i < PATCH_SZ; i++) g_sw.desc_gauss[i] = desc_gauss20[i];
}

static void build_tables(surf_tables *t)
{
    float g[2 * ORI_RADIUS + 1], gd[PATCH_SZ];
    orc_gaussian_kernel(2 * ORI_RADIUS + 1, ORI_SIGMA, g);
    t->npts = 0;
    for (int i = -ORI_RADIUS; i <= ORI_RADIUS; i++)
        for (int j = -ORI_RADIUS; j <= ORI_RADIUS; j++)
            if (i * i + j * j <= ORI_RADIUS * ORI_RADIUS) {
                t->px[t->npts] = i;
                t->py[t->npts] = j;
                t->pw[t->npts++] = g[i + ORI_RADIUS] * g[j + ORI_RADIUS];
            }
    orc_gaussian_kernel(PATCH_SZ, DESC_SIGMA, gd);
    if (g_sw.have_desc_gauss)
        for (int i = 0; i < PATCH_SZ; i++) gd[i] = g_sw.desc_gauss[i];
    for (int i = 0; i < PATCH_SZ; i++)
        for (int j = 0; j < PATCH_SZ; j++) t->dw[i * PATCH_SZ + j] = gd[i] * gd[j];
}

static void describe_one(const uint8_t *img, int H, int W, int pitch, const int32_t *sum,
                         const surf_tables *T, const orc_params *p, orc_keypoint *kp, float *vec,
                         uint8_t *patch_out, uint8_t **winbuf, size_t *wincap)
{
    const int SH = H + 1, SW = W + 1;
    const float size = kp->size;
    const float cx = kp->x, cy = kp->y;
    const float s = size * 1.2f / 9.0f;
    const int g = 2 * rnd(2 * s);
    if (SH < g || SW < g) {
        kp->size = -1;
        return;
    }
    float dir = g_sw.angle_pre244 ? 90.f : 360.f - 90.f;
    if (!p->upright) {
        box_t hx[2], hy[2];
        float X[ORI_BOUND], Y[ORI_BOUND], ang[ORI_BOUND];
        int na = 0;
        scale_pattern(PAT_HX, hx, 2, 4, g, SW);
        scale_pattern(PAT_HY, hy, 2, 4, g, SW);
        for (int k = 0; k < T->npts; k++) {
            int x = rnd(cx + T->px[k] * s - (float)(g - 1) / 2);
            int y = rnd(cy + T->py[k] * s - (float)(g - 1) / 2);
            if (y < 0 || y >= SH - g || x < 0 || x >= SW - g) continue;
            const int32_t *o = sum + (size_t)y * SW + x;
            float vx = box_response(o, hx, 2);
            float vy = box_response(o, hy, 2);
            X[na] = vx * T->pw[k];
            Y[na] = vy * T->pw[k];
            na++;
        }
        if (na == 0) {
            kp->size = -1;
            return;
        }
        int iang[ORI_BOUND]; /* cvRound(angle[j]) hoisted out of the window loop: same values */
        for (int k = 0; k < na; k++) {
            ang[k] = orc_fast_atan2(Y[k], X[k]);
            iang[k] = rnd(ang[k]);
        }
        float bx = 0, by = 0, best = 0;
        for (int i = 0; i < 360; i += ORI_SEARCH_INC) {
            float sx = 0, sy = 0;
            for (int j = 0; j < na; j++) {
                int d = abs(iang[j] - i);
                if (d < ORI_WIN / 2 || d > 360 - ORI_WIN / 2) {
                    sx += X[j];
                    sy += Y[j];
                }
            }
            float mod = sx * sx + sy * sy;
            if (mod > best) {
                best = mod;
                bx = sx;
                by = sy;
            }
        }
        dir = g_sw.angle_pre244 ? orc_fast_atan2(by, bx) : orc_fast_atan2(-by, bx);
    }
    kp->angle = dir;
    if (!vec && !patch_out) return;

    const int n = (int)((PATCH_SZ + 1) * s);
    if (n < 1) { /* size < 0.36: upstream builds an empty window and cv::resize raises; the keypoint is deleted here */
        kp->size = -1;
        return;
    }
    if ((size_t)n * n > *wincap) {
        *wincap = (size_t)n * n;
        *winbuf = (uint8_t *)realloc(*winbuf, *wincap);
    }
    uint8_t *win = *winbuf;
    if (!p->upright) {
        float th = dir * (float)(3.1415926535897932384626433832795 / 180);
        float sd = g_sw.libm_sincos ? sinf(th) : (float)sin((double)th);
        float cd = g_sw.libm_sincos ? cosf(th) : (float)cos((double)th);
        if (!g_sw.angle_pre244) sd = -sd;
        float off = -(float)(n - 1) / 2;
        float sx0 = cx + off * cd + off * sd;
        float sy0 = cy - off * sd + off * cd;
        const int xlim = W - 1, ylim = H - 1;
        for (int i = 0; i < n; i++, sx0 += sd, sy0 += cd) {
            double fx = sx0, fy = sy0;
            for (int j = 0; j < n; j++, fx += cd, fy -= sd) {
                int ix = ifloor(fx), iy = ifloor(fy);
                if ((unsigned)ix < (unsigned)xlim && (unsigned)iy < (unsigned)ylim) {
                    float a = (float)(fx - ix), b = (float)(fy - iy);
                    const uint8_t *q = img + (size_t)iy * pitch + ix;
                    win[i * n + j] = (uint8_t)rnd(q[0] * (1.f - a) * (1.f - b) + q[1] * a * (1.f - b) +
                                                  q[pitch] * (1.f - a) * b + q[pitch + 1] * a * b);
                } else {
                    int x = imin(imax(rnd(fx), 0), xlim);
                    int y = imin(imax(rnd(fy), 0), ylim);
                    win[i * n + j] = img[(size_t)y * pitch + x];
                }
            }
        }
    } else {
        float off = -(float)(n - 1) / 2;
        int x0 = rnd(cx + off);
        int y0 = rnd(cy - off);
        for (int i = 0; i < n; i++) {
            int x = imin(imax(x0 + i, 0), W - 1);
            for (int j = 0; j < n; j++) {
                int y = imin(imax(y0 - j, 0), H - 1);
                win[i * n + j] = img[(size_t)y * pitch + x];
            }
        }
    }
    uint8_t P[PATCH_W * PATCH_W];
    orc_resize_area_21(win, n, P);
    if (patch_out) memcpy(patch_out, P, sizeof(P));
    if (!vec) return;

    float DX[PATCH_SZ][PATCH_SZ], DY[PATCH_SZ][PATCH_SZ];
    for (int i = 0; i < PATCH_SZ; i++)
        for (int j = 0; j < PATCH_SZ; j++) {
            float w = T->dw[i * PATCH_SZ + j];
            const uint8_t *r0 = P + i * PATCH_W + j, *r1 = r0 + PATCH_W;
            DX[i][j] = (r0[1] - r0[0] + r1[1] - r1[0]) * w;
            DY[i][j] = (r1[0] - r0[0] + r1[1] - r0[1]) * w;
        }
    const int D = p->extended ? 128 : 64;
    for (int k = 0; k < D; k++) vec[k] = 0;
    double mag = 0;
    float *v = vec;
    for (int I = 0; I < 4; I++)
        for (int J = 0; J < 4; J++) {
            for (int y = I * 5; y < I * 5 + 5; y++)
                for (int x = J * 5; x < J * 5 + 5; x++) {
                    float tx = DX[y][x], ty = DY[y][x];
                    if (p->extended) {
                        if (ty >= 0) {
                            v[0] += tx;
                            v[1] += fabsf(tx);
                        } else {
                            v[2] += tx;
                            v[3] += fabsf(tx);
                        }
                        if (tx >= 0) {
                            v[4] += ty;
                            v[5] += fabsf(ty);
                        } else {
                            v[6] += ty;
                            v[7] += fabsf(ty);
                        }
                    } else {
                        v[0] += tx;
                        v[1] += ty;
                        v[2] += fabsf(tx);
                        v[3] += fabsf(ty);
                    }
                }
            const int nb = p->extended ? 8 : 4;
            for (int k = 0; k < nb; k++) mag += v[k] * v[k];
            v += nb;
        }
    float sc = (float)(1. / (sqrt(mag) + DBL_EPSILON));
    for (int k = 0; k < D; k++) vec[k] *= sc;
}

void orc_describe(const uint8_t *img, int H, int W, int pitch, const orc_params *p, orc_keypoint *kps,
                  int n, float *desc, uint8_t *patches)
{
    surf_tables T;
    build_tables(&T);
    int32_t *sum = (int32_t *)malloc(sizeof(int32_t) * (size_t)(H + 1) * (size_t)(W + 1));
    orc_integral(img, H, W, pitch, sum, 0);
    const int D = p->extended ? 128 : 64;
    uint8_t *winbuf = NULL;
    size_t wincap = 0;
    for (int k = 0; k < n; k++)
        describe_one(img, H, W, pitch, sum, &T, p, kps + k, desc ? desc + (size_t)k * D : NULL,
                     patches ? patches + (size_t)k * PATCH_W * PATCH_W : NULL, &winbuf, &wincap);
    free(winbuf);
    free(sum);
}

/* Stable removal of keypoints marked size <= 0 (and their descriptor rows): SURF::operator() tail. */
static int squeeze(orc_keypoint *kps, int n, float *desc, int D)
{
    int j = 0;
    for (int i = 0; i < n; i++)
        if (kps[i].size > 0) {
            if (i > j) {
                kps[j] = kps[i];
                if (desc) memcpy(desc + (size_t)j * D, desc + (size_t)i * D, sizeof(float) * (size_t)D);
            }
            j++;
        }
    return j;
}

int orc_raw_keypoints(const uint8_t *img, int H, int W, int pitch, const uint8_t *mask, int mask_pitch,
                      const orc_params *p, orc_keypoint *kps, int cap)
{
    if (H <= 0 || W <= 0) return 0;
    size_t cnt = (size_t)(H + 1) * (size_t)(W + 1);
    int32_t *sum = (int32_t *)malloc(sizeof(int32_t) * cnt);
    int32_t *msum = NULL;
    orc_integral(img, H, W, pitch, sum, 0);
    if (mask) {
        msum = (int32_t *)malloc(sizeof(int32_t) * cnt);
        orc_integral(mask, H, W, mask_pitch, msum, 1);
    }
    kpvec v = fast_hessian(sum, msum, H, W, p);
    int n = v.n;
    memcpy(kps, v.v, sizeof(orc_keypoint) * (size_t)imin(n, cap));
    free(v.v);
    free(sum);
    free(msum);
    return n;
}

static int run_full(const uint8_t *img, int H, int W, int pitch, const uint8_t *mask, int mask_pitch,
                    const orc_params *p, orc_keypoint *kps, int cap, float *desc)
{
    int n = orc_raw_keypoints(img, H, W, pitch, mask, mask_pitch, p, kps, cap);
    if (n > cap) return -n; /* caller's buffer too small: report the need, write nothing more */
    if (n == 0) return 0;
    orc_describe(img, H, W, pitch, p, kps, n, desc, NULL);
    return squeeze(kps, n, desc, p->extended ? 128 : 64);
}

int orc_detect(const uint8_t *img, int H, int W, int pitch, const uint8_t *mask, int mask_pitch,
               const orc_params *p, orc_keypoint *kps, int cap)
{
    return run_full(img, H, W, pitch, mask, mask_pitch, p, kps, cap, NULL);
}

int orc_detect_and_compute(const uint8_t *img, int H, int W, int pitch, const uint8_t *mask, int mask_pitch,
                           const orc_params *p, orc_keypoint *kps, int cap, float *desc)
{
    return run_full(img, H, W, pitch, mask, mask_pitch, p, kps, cap, desc);
}

/* DescriptorExtractor::compute [features2d/descriptors.cpp]: drop size < FLT_EPSILON first. */
int orc_compute(const uint8_t *img, int H, int W, int pitch, const orc_params *p, orc_keypoint *kps, int n,
                float *desc)
{
    if (H <= 0 || W <= 0 || n <= 0) return 0;
    int m = 0;
    for (int i = 0; i < n; i++)
        if (!(kps[i].size < FLT_EPSILON)) kps[m++] = kps[i];
    if (m == 0) return 0;
    orc_describe(img, H, W, pitch, p, kps, m, desc, NULL);
    return squeeze(kps, m, desc, p->extended ? 128 : 64);
}

long orc_detect_and_compute_batch(const uint8_t *imgs, int B, int H, int W, const orc_params *p,
                                  orc_keypoint *kps, int cap, float *desc, int32_t *counts, int n_threads)
{
    const int D = p->extended ? 128 : 64;
    long total = 0;
#ifdef _OPENMP
    if (n_threads <= 0) n_threads = omp_get_max_threads();
#pragma omp parallel for schedule(dynamic, 1) num_threads(n_threads) reduction(+ : total)
#endif
    for (int b = 0; b < B; b++) {
        int n = orc_detect_and_compute(imgs + (size_t)b * H * W, H, W, W, NULL, 0, p, kps + (size_t)b * cap, cap,
                                       desc ? desc + (size_t)b * cap * D : NULL);
        counts[b] = n;
        if (n > 0) total += n;
    }
    return total;
}

/* ------------------------------------------------------------------------------------------
 * Stage 5.  BFMatcher(NORM_L2)::knnMatch [features2d/matchers.cpp]; SURVEY A6, row a11.
 * ------------------------------------------------------------------------------------------ */
static inline float l2sqr(const float *a, const float *b, int n)
{
    float s = 0;
    int j = 0;
    for (; j <= n - 4; j += 4) {
        float t0 = a[j] - b[j], t1 = a[j + 1] - b[j + 1], t2 = a[j + 2] - b[j + 2], t3 = a[j + 3] - b[j + 3];
        s += t0 * t0 + t1 * t1 + t2 * t2 + t3 * t3;
    }
    for (; j < n; j++) {
        float t = a[j] - b[j];
        s += t * t;
    }
    return s;
}

void orc_knn_match(const float *q, int nq, const float *t, int nt, int d, int k, orc_dmatch *out, int n_threads)
{
#ifdef _OPENMP
    if (n_threads <= 0) n_threads = omp_get_max_threads();
#pragma omp parallel for schedule(static) num_threads(n_threads)
#endif
    for (int i = 0; i < nq; i++) {
        orc_dmatch *o = out + (size_t)i * k;
        int have = 0;
        for (int j = 0; j < nt; j++) {
            float dist = l2sqr(q + (size_t)i * d, t + (size_t)j * d, d);
            if (have == k && !(dist < o[k - 1].distance)) continue;
            int pos = have < k ? have++ : k - 1;
            while (pos > 0 && dist < o[pos - 1].distance) {
                o[pos] = o[pos - 1];
                pos--;
            }
            o[pos].query_idx = i;
            o[pos].train_idx = j;
            o[pos].img_idx = 0;
            o[pos].distance = dist;
        }
        for (int r = 0; r < have; r++) o[r].distance = sqrtf(o[r].distance);
        for (int r = have; r < k; r++) {
            o[r].query_idx = i;
            o[r].train_idx = -1;
            o[r].img_idx = -1;
            o[r].distance = INFINITY;
        }
    }
}

/* =====================================================================================================
 * "Next" rows f2 / f3: train database localisation, bogus filter, Hough pose vote, per-image votes.
 * Restates SurfFeatures/Logic/vtkSlicerSurfFeaturesLogic.cxx:73-345 (HoughCodeSimilarity, angle_radians,
 * compatible_poll_booths_line_segment, houghTransform) and :1434-1571 (computeNextInterSliceCorrespondence).
 *
 * Arithmetic model.  The reference file says `using namespace std;`, so sin / cos / atan2 / log / sqrt on
 * float arguments resolve to the float overloads, and every product-sum is separate fp32 operations (x86-64
 * SSE, no contraction).  Which float libm produced the reference's numbers is platform-dependent (it was built
 * on Windows and macOS); this restatement therefore defines the float functions as CORRECTLY ROUNDED
 * (evaluate in double, round once) -- libm_float != 0 switches to this machine's sinf/cosf/atan2f/logf so a
 * test can measure how often that choice matters.  angle_radians() deliberately keeps the reference's factor
 * 2 (it maps degrees to 2*pi*a/180) and is applied a second time inside the compatibility test, as upstream.
 * ===================================================================================================== */
#define ORC_PI 3.141592653589793

static float h_sin(float x, int lf) { return lf ? sinf(x) : (float)sin((double)x); }
static float h_cos(float x, int lf) { return lf ? cosf(x) : (float)cos((double)x); }
static float h_atan2(float y, float x, int lf) { return lf ? atan2f(y, x) : (float)atan2((double)y, (double)x); }
static float h_log(float x, int lf) { return lf ? logf(x) : (float)log((double)x); }

/* angle_radians (:196-201): double arithmetic, result returned as float */
static float h_angle_radians(float a) { return (float)((2.0f * ORC_PI * a) / 180.0f); }

typedef struct { float row, col, ori, scl; } orc_booth;

/* HoughCodeSimilarity::SetHoughCode (:86-110) + FindPollBooth (:112-141, non-mirrored branch) for one match */
static orc_booth h_poll_booth(const orc_keypoint *q, const orc_keypoint *t, int refx, int refy, int lf)
{
    const float key_row = q->y, key_col = q->x, key_ori = h_angle_radians(q->angle), key_scl = q->size;
    const float poll_row = (float)refy, poll_col = (float)refx;
    const float diff_row = poll_row - key_row, diff_col = poll_col - key_col;
    const float dist_sqr = diff_col * diff_col + diff_row * diff_row;
    const float dist_to_booth = sqrtf(dist_sqr);
    const float angle_to_booth = h_atan2(diff_row, diff_col, lf);
    const float angle_diff = 0.0f - key_ori;
    const float scale_diff = 10.0f / key_scl;
    const float f_row = t->y, f_col = t->x, f_ori = h_angle_radians(t->angle), f_scl = t->size;
    const float ori_diff = f_ori - key_ori;
    const float scl_ratio = f_scl / key_scl;
    const float dist = dist_to_booth * scl_ratio;
    const float angle = angle_to_booth + ori_diff;
    orc_booth b;
    b.row = h_sin(angle, lf) * dist + f_row;
    b.col = h_cos(angle, lf) * dist + f_col;
    b.ori = f_ori + angle_diff;
    b.scl = scale_diff * f_scl;
    return b;
}

/* compatible_poll_booths_line_segment (:207-260) with its default thresholds; a = "training", b = "poll" */
static int h_compatible(const orc_booth *a, const orc_booth *b, int lf)
{
    const float thr_trans = 30.0f / 43.0f, thr_scale = h_log(1.5f, lf), thr_ori = (float)(20.0f * ORC_PI / 180.0F);
    const float a_ori = h_angle_radians(a->ori), b_ori = h_angle_radians(b->ori);
    const float log_diff = fabsf(h_log(a->scl, lf) - h_log(b->scl, lf));
    const float dc = b->col - a->col, dr = b->row - a->row;
    const float dist = sqrtf(dc * dc + dr * dr);
    float od = b_ori > a_ori ? b_ori - a_ori : a_ori - b_ori;
    if (od > (2 * ORC_PI - od)) od = (float)(2 * ORC_PI - od);
    return dist < thr_trans * a->scl && log_diff < thr_scale && od < thr_ori;
}

/* houghTransform(queryKeypoints, trainKeypoints (per image), matches, refx, refy) (:262-345).  train keypoints
 * are one packed array with img_off[n_img + 1].  Entries with negative train / image index are skipped when
 * the booths are built; the rejection loop then indexes `matches` by booth index, as upstream does.  Returns
 * the vote count of the winning booth (0 when there is no valid match). */
int orc_hough_transform(const orc_keypoint *qk, const orc_keypoint *tk, const int32_t *img_off, orc_dmatch *matches,
                        int n, int refx, int refy, int libm_float)
{
    orc_booth *kg = (orc_booth *)malloc(sizeof(orc_booth) * (size_t)(n > 0 ? n : 1));
    int m = 0;
    for (int i = 0; i < n; i++) {
        if (matches[i].train_idx < 0 || matches[i].img_idx < 0) continue;
        kg[m++] = h_poll_booth(qk + matches[i].query_idx, tk + img_off[matches[i].img_idx] + matches[i].train_idx, refx,
                               refy, libm_float);
    }
    int best = -1, best_count = 0;
    for (int i = 0; i < m; i++) {
        int c = 0;
        for (int j = 0; j < m; j++) c += h_compatible(kg + i, kg + j, libm_float);
        if (c > best_count) {
            best = i;
            best_count = c;
        }
    }
    if (best >= 0)
        for (int j = 0; j < m; j++)
            if (!h_compatible(kg + best, kg + j, libm_float)) {
                matches[j].img_idx = -1;
                matches[j].query_idx = -1;
                matches[j].train_idx = -1;
            }
    free(kg);
    return best_count;
}

/* One query image of computeNextInterSliceCorrespondence (:1457-1558): m_train = best train match per query
 * keypoint (first column of knnMatch k=3), m_bogus = best bogus match (k=1).  out receives the surviving
 * matches (afterHoughMatches[i]); result = {n_matches, hough_count, best_image, best_count}. */
void orc_correspond(const orc_keypoint *qk, int nq, const orc_dmatch *m_train, const orc_dmatch *m_bogus,
                    const orc_keypoint *tk, const int32_t *img_off, int n_img, float bogus_ratio, int refx, int refy,
                    orc_dmatch *out, int32_t *result, int libm_float)
{
    int m = 0;
    for (int j = 0; j < nq; j++) {
        const float ratio = m_train[j].distance / m_bogus[j].distance;
        if (ratio > bogus_ratio) continue;
        out[m++] = m_train[j];
    }
    const int hough = orc_hough_transform(qk, tk, img_off, out, m, refx, refy, libm_float);
    int32_t *votes = (int32_t *)calloc((size_t)(n_img > 0 ? n_img : 1), sizeof(int32_t));
    for (int j = 0; j < m; j++)
        if (out[j].img_idx >= 0) votes[out[j].img_idx]++;
    int best = -1, best_count = -1;
    for (int j = 0; j < n_img; j++)
        if (votes[j] > best_count) {
            best_count = votes[j];
            best = j;
        }
    free(votes);
    result[0] = m;
    result[1] = hough;
    result[2] = best;
    result[3] = best_count;
}

/* computeCentroid (:839-860): fp32 running sums of the row / column indices of non-zero mask pixels */
void orc_mask_centroid(const uint8_t *mask, int H, int W, int pitch, int *x, int *y)
{
    float rowc = 0, colc = 0;
    int count = 0;
    for (int i = 0; i < H; i++)
        for (int j = 0; j < W; j++)
            if (mask[(size_t)i * pitch + j] > 0) {
                rowc += i;
                colc += j;
                count += 1;
            }
    rowc /= count;
    colc /= count;
    *y = (int)rowc;
    *x = (int)colc;
}
