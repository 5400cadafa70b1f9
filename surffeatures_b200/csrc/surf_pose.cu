// surf_pose.cu -- the "next" row f4 of SURVEY.md section 8 behind the C ABI: plane fit by RANSAC and the pose
// estimate that follow the inter-slice correspondences.
//   surf_rand_*               rand() as the reference's platform provides it (glibc TYPE_3 generator, never seeded
//                             by the reference = srand(1)); an explicit state object instead of the process-global one
//   surf_match_points         trainPoints / queryPoints of updateMatchNodeRansac, ...Logic.cxx:1672-1703
//   surf_ransac_plane         vtkSlicerSurfFeaturesLogic::ransac, ...Logic.cxx:2085-2188
//   surf_estimate_pose        the arithmetic of updateMatchNodeRansac, ...Logic.cxx:1714-1823
//   surf_mean_plane_distance  ...Logic.cxx:1826-1846
// The random index triples are drawn on the host in the reference's order (the rejection loop consumes a variable
// number of rand() calls); the 1000 hypotheses are then independent: k_ransac_score evaluates one per thread,
// each walking the points in index order, so that its fp64 error sum is the reference's sequential sum.  The
// host picks the winner with the reference's scan (fewest outliers, then smallest error, first wins) and lists
// its inliers.  The pose estimate is ~2 MFLOP of dependent double arithmetic with libm calls: it stays on the host.
// Compiled with -fmad=false: the reference's vnl arithmetic is separate multiplies and adds.
#include <cmath>
#include <cstring>
#include <new>
#include <vector>

#include "surf_engine_priv.h"

struct surf_rand {
    int32_t st[31];
    int f, b;
};

namespace {

struct V3 {
    double v[3];
};
__host__ __device__ inline V3 sub3(const V3 &a, const V3 &b) { return V3{{a.v[0] - b.v[0], a.v[1] - b.v[1], a.v[2] - b.v[2]}}; }
__host__ __device__ inline double dot3(const V3 &a, const V3 &b)
{
    double ip = 0;
    for (int i = 0; i < 3; i++) ip += a.v[i] * b.v[i];
    return ip;
}
__host__ __device__ inline double norm3(const V3 &a) { return sqrt(dot3(a, a)); }
// vnl_c_vector<double>::normalize: scale by 1 / sqrt(squared norm) unless the norm is 0
__host__ __device__ inline V3 normalize3(V3 a)
{
    double t = dot3(a, a);
    if (t != 0) {
        t = 1.0 / sqrt(t);
        for (int i = 0; i < 3; i++) a.v[i] = t * a.v[i];
    }
    return a;
}
__host__ __device__ inline V3 cross3(const V3 &a, const V3 &b)
{
    return V3{{a.v[1] * b.v[2] - a.v[2] * b.v[1], a.v[2] * b.v[0] - a.v[0] * b.v[2], a.v[0] * b.v[1] - a.v[1] * b.v[0]}};
}

// Plane of hypothesis (i0, i1, i2); false where the reference `continue`s (coincident or collinear points).
__host__ __device__ inline bool plane_of(const V3 *P, int i0, int i1, int i2, V3 &N, double &D)
{
    const V3 p1 = P[i0];
    const V3 v1 = normalize3(sub3(p1, P[i1])), v2 = normalize3(sub3(p1, P[i2]));
    if (norm3(v1) < 0.99 || norm3(v2) < 0.99) return false;
    if (fabs(dot3(v1, v2)) > 0.99) return false;
    N = normalize3(cross3(v1, v2));
    D = -dot3(p1, N);
    return true;
}

// One thread per hypothesis.  n_out[h] = -1: skipped.
__global__ void __launch_bounds__(128) k_ransac_score(const V3 *__restrict__ P, int n, const int *__restrict__ triples, int iters,
                                                      float threshold, int *__restrict__ n_out, double *__restrict__ err_out)
{
    const int h = blockIdx.x * blockDim.x + threadIdx.x;
    if (h >= iters) return;
    const int i0 = triples[3 * h], i1 = triples[3 * h + 1], i2 = triples[3 * h + 2];
    V3 N;
    double D;
    if (!plane_of(P, i0, i1, i2, N, D)) {
        n_out[h] = -1;
        err_out[h] = 0;
        return;
    }
    int outliers = 0;
    double err = 0;
    for (int j = 0; j < n; j++) {
        if (j == i0 || j == i1 || j == i2) continue;
        const float dist = (float)fabs(dot3(N, P[j]) + D);
        if (dist > threshold)
            outliers++;
        else
            err += dist;
    }
    n_out[h] = outliers;
    err_out[h] = err;
}

V3 transform_point(const V3 &p, const double M[16])
{
    V3 r;
    for (int i = 0; i < 3; i++) r.v[i] = p.v[0] * M[4 * i] + p.v[1] * M[4 * i + 1] + p.v[2] * M[4 * i + 2] + 1.0 * M[4 * i + 3];
    return r;
}

// getRotationMatrix :768-782 (columns normalised as vnl_matrix::normalize_columns does)
void rotation_matrix(const V3 &axis, double theta, double R[9])
{
    const double *a = axis.v;
    R[0] = cos(theta) + a[0] * a[0] * (1 - cos(theta));
    R[1] = a[0] * a[1] * (1 - cos(theta)) - a[2] * sin(theta);
    R[2] = a[0] * a[2] * (1 - cos(theta)) + a[1] * sin(theta);
    R[3] = a[1] * a[0] * (1 - cos(theta)) + a[2] * sin(theta);
    R[4] = cos(theta) + a[1] * a[1] * (1 - cos(theta));
    R[5] = a[1] * a[2] * (1 - cos(theta)) - a[0] * sin(theta);
    R[6] = a[2] * a[0] * (1 - cos(theta)) - a[1] * sin(theta);
    R[7] = a[2] * a[1] * (1 - cos(theta)) + a[0] * sin(theta);
    R[8] = cos(theta) + a[2] * a[2] * (1 - cos(theta));
    for (int j = 0; j < 3; j++) {
        double nrm = 0;
        for (int i = 0; i < 3; i++) nrm += R[3 * i + j] * R[3 * i + j];
        if (nrm != 0) {
            const double sc = 1.0 / sqrt(nrm);
            for (int i = 0; i < 3; i++) R[3 * i + j] *= sc;
        }
    }
}
V3 mat_vec(const double R[9], const V3 &x)
{
    V3 r;
    for (int i = 0; i < 3; i++) {
        double s = 0;
        for (int k = 0; k < 3; k++) s += R[3 * i + k] * x.v[k];
        r.v[i] = s;
    }
    return r;
}
double get_angle(const V3 &u, const V3 &v) { return acos(dot3(u, v) / (norm3(u) * norm3(v))); }

unsigned long draw(surf_rand *r, unsigned long n) { return (unsigned long)surf_rand_next(r) % n; }

}  // namespace

extern "C" {

int surf_rand_create(uint32_t seed, surf_rand **out)
{
    if (!out) return SURF_ERR_BAD_ARG;
    surf_rand *r = new (std::nothrow) surf_rand();
    if (!r) return SURF_ERR_CUDA;
    // srandom_r, TYPE_3: Lehmer sequence (16807, Schrage's method) into 31 words, then 310 outputs discarded
    if (seed == 0) seed = 1;
    r->st[0] = (int32_t)seed;
    for (int i = 1; i < 31; i++) {
        const long hi = r->st[i - 1] / 127773, lo = r->st[i - 1] % 127773;
        long word = 16807 * lo - 2836 * hi;
        if (word < 0) word += 2147483647;
        r->st[i] = (int32_t)word;
    }
    r->f = 3;
    r->b = 0;
    for (int i = 0; i < 310; i++) surf_rand_next(r);
    *out = r;
    return SURF_OK;
}

void surf_rand_destroy(surf_rand *r) { delete r; }

int surf_rand_next(surf_rand *r)
{
    const uint32_t v = (uint32_t)r->st[r->f] + (uint32_t)r->st[r->b];
    r->st[r->f] = (int32_t)v;
    if (++r->f >= 31) r->f = 0;
    if (++r->b >= 31) r->b = 0;
    return (int)(v >> 1);
}

int surf_match_points(const surf_dmatch *matches, int n_matches, const surf_keypoint *query_keypoints,
                      const surf_keypoint *train_keypoints, const int32_t *train_offsets, int n_train_images,
                      const float *train_transforms, const double image_to_probe[16], double *train_points,
                      double *query_points, int32_t *n_points)
{
    if (!n_points || n_matches < 0 || (n_matches > 0 && (!matches || !query_keypoints || !train_keypoints || !train_offsets ||
                                                          !train_transforms || !image_to_probe || !train_points || !query_points)))
        return SURF_ERR_BAD_ARG;
    int m = 0;
    for (int j = 0; j < n_matches; j++) {
        const int img = matches[j].img_idx;
        if (img == -1) continue;  // rejected by the Hough vote (afterHoughMatches keeps the entry)
        if (img < 0 || img >= n_train_images) return SURF_ERR_BAD_ARG;
        // combinedTransform = trainImagesTransform[img] * ImageToProbeTransform (vtkTransform pre-multiplies)
        double M[16];
        const float *A = train_transforms + 16 * (size_t)img;
        for (int i = 0; i < 4; i++)
            for (int c = 0; c < 4; c++) {
                double s = 0;
                for (int k = 0; k < 4; k++) s += (double)A[4 * i + k] * image_to_probe[4 * k + c];
                M[4 * i + c] = s;
            }
        const surf_keypoint &tk = train_keypoints[(size_t)train_offsets[img] + matches[j].train_idx];
        const V3 r = transform_point(V3{{tk.x, tk.y, 0.0}}, M);
        memcpy(train_points + 3 * (size_t)m, r.v, sizeof(r.v));
        const surf_keypoint &qk = query_keypoints[matches[j].query_idx];
        query_points[3 * (size_t)m] = qk.x;
        query_points[3 * (size_t)m + 1] = qk.y;
        query_points[3 * (size_t)m + 2] = 0.0;
        m++;
    }
    *n_points = m;
    return SURF_OK;
}

int surf_ransac_plane(surf_engine *e, surf_rand *rng, const double *points, int n, float margin, int iterations,
                      int32_t *inliers, int32_t *n_inliers, int32_t plane_idx[3], double *best_error, int32_t *n_outliers)
{
    if (!e) return SURF_ERR_BAD_ARG;
    if (!rng || !inliers || !n_inliers || !plane_idx || n < 0 || iterations < 0 || (n > 0 && !points))
        return fail(e, SURF_ERR_BAD_ARG, "null pointer or negative count");
    *n_inliers = 0;
    plane_idx[0] = plane_idx[1] = plane_idx[2] = -1;
    if (best_error) *best_error = 0;
    if (n_outliers) *n_outliers = 0;
    if (n == 3) {  // :2090-2099
        for (int i = 0; i < 3; i++) plane_idx[i] = inliers[i] = i;
        *n_inliers = 3;
        return SURF_OK;
    }
    if (n < 3) return fail(e, SURF_ERR_BAD_ARG, "a plane needs at least three points (the reference returns 1)");
    int rc = ensure_device(e);
    if (rc) return rc;
    // the index triples, in the reference's draw order
    std::vector<int> triples((size_t)iterations * 3);
    for (int it = 0; it < iterations; it++) {
        int idx[3] = {0, 0, 0};
        while (idx[0] == idx[1] || idx[1] == idx[2] || idx[0] == idx[2]) {
            idx[0] = (int)draw(rng, (unsigned long)n);
            idx[1] = (int)draw(rng, (unsigned long)n);
            idx[2] = (int)draw(rng, (unsigned long)n);
        }
        for (int k = 0; k < 3; k++) triples[3 * (size_t)it + k] = idx[k];
    }
    const size_t pb = sizeof(V3) * (size_t)n, tb = sizeof(int) * triples.size();
    const size_t ob = (sizeof(int) + sizeof(double)) * (size_t)iterations;
    unsigned char *d = nullptr;
    CU(e, cudaMalloc(&d, pb + tb + ob + 64));
    V3 *d_p = reinterpret_cast<V3 *>(d);
    int *d_t = reinterpret_cast<int *>(d + pb);
    double *d_err = reinterpret_cast<double *>(d + ((pb + tb + 7) & ~(size_t)7));
    int *d_no = reinterpret_cast<int *>(reinterpret_cast<unsigned char *>(d_err) + sizeof(double) * (size_t)iterations);
    std::vector<int> h_no((size_t)iterations);
    std::vector<double> h_err((size_t)iterations);
    cudaError_t ce = cudaMemcpyAsync(d_p, points, pb, cudaMemcpyHostToDevice, e->stream);
    if (ce == cudaSuccess) ce = cudaMemcpyAsync(d_t, triples.data(), tb, cudaMemcpyHostToDevice, e->stream);
    if (ce == cudaSuccess && iterations > 0) {
        k_ransac_score<<<(iterations + 127) / 128, 128, 0, e->stream>>>(d_p, n, d_t, iterations, margin, d_no, d_err);
        e->launches = 1;
        ce = cudaGetLastError();
    }
    if (ce == cudaSuccess) ce = cudaMemcpyAsync(h_no.data(), d_no, sizeof(int) * (size_t)iterations, cudaMemcpyDeviceToHost, e->stream);
    if (ce == cudaSuccess) ce = cudaMemcpyAsync(h_err.data(), d_err, sizeof(double) * (size_t)iterations, cudaMemcpyDeviceToHost, e->stream);
    if (ce == cudaSuccess) ce = cudaStreamSynchronize(e->stream);
    cudaFree(d);
    if (ce != cudaSuccess) return fail(e, SURF_ERR_CUDA, "plane scoring failed: %s", cudaGetErrorString(ce));
    // the reference's scan :2163-2173
    double best_n = n, best_e = 1.7976931348623157e308;
    int best = -1;
    for (int it = 0; it < iterations; it++) {
        if (h_no[it] < 0) continue;
        if (h_no[it] < best_n || (h_no[it] == best_n && h_err[it] < best_e)) {
            best_n = h_no[it];
            best_e = h_err[it];
            best = it;
        }
    }
    if (best < 0) return fail(e, SURF_ERR_UNSUPPORTED, "no valid plane among %d hypotheses (the reference would read an empty vector)", iterations);
    // inliers of the winner, in point order, then its three points (:2172, :2175-2177)
    const V3 *P = reinterpret_cast<const V3 *>(points);
    const int *idx = &triples[3 * (size_t)best];
    V3 N;
    double D;
    plane_of(P, idx[0], idx[1], idx[2], N, D);
    int m = 0, outl = 0;
    double err = 0;
    for (int j = 0; j < n; j++) {
        if (j == idx[0] || j == idx[1] || j == idx[2]) continue;
        const float dist = (float)fabs(dot3(N, P[j]) + D);
        if (dist > margin) {
            outl++;
        } else {
            err += dist;
            inliers[m++] = j;
        }
    }
    if (outl != h_no[best] || err != h_err[best])
        return fail(e, SURF_ERR_CUDA, "internal error: host and device disagree on the winning plane (%d/%d outliers)", outl, h_no[best]);
    for (int k = 0; k < 3; k++) inliers[m++] = plane_idx[k] = idx[k];
    *n_inliers = m;
    if (best_error) *best_error = best_e;
    if (n_outliers) *n_outliers = (int32_t)best_n;
    return SURF_OK;
}

int surf_estimate_pose(surf_rand *rng, const double *query_points, const double *train_points, int n, const int32_t *inliers,
                       int n_inliers, const int32_t plane_idx[3], double estimate[16], double *mean_match_distance)
{
    if (!rng || !query_points || !train_points || !inliers || !plane_idx || !estimate || n < 3 || n_inliers < 2) return SURF_ERR_BAD_ARG;
    for (int k = 0; k < 3; k++)
        if (plane_idx[k] < 0 || plane_idx[k] >= n) return SURF_ERR_BAD_ARG;
    for (int i = 0; i < n_inliers; i++)
        if (inliers[i] < 0 || inliers[i] >= n) return SURF_ERR_BAD_ARG;
    const V3 *Q = reinterpret_cast<const V3 *>(query_points), *T = reinterpret_cast<const V3 *>(train_points);
    // getPlaneFromPoints :792-803
    const V3 v1 = normalize3(sub3(T[plane_idx[0]], T[plane_idx[1]])), v2 = normalize3(sub3(T[plane_idx[0]], T[plane_idx[2]]));
    const V3 plane = normalize3(cross3(v1, v2));
    const double d = -dot3(T[plane_idx[0]], plane);
    std::vector<V3> proj((size_t)n);
    for (int i = 0; i < n; i++) {  // projectPoint :784-790
        const V3 nn = normalize3(plane);
        const double dist = dot3(T[i], nn) + d;
        for (int k = 0; k < 3; k++) proj[i].v[k] = T[i].v[k] - dist * nn.v[k];
    }
    const int max_iter = 10000, mult = 20;
    const int iter = n * mult < max_iter ? n * mult : max_iter;
    V3 u{{0, 0, 0}}, v{{0, 0, 0}};
    const V3 ex{{1.0, 0.0, 0.0}}, ey{{0.0, 1.0, 0.0}};
    for (int i = 0; i < iter; i++) {
        int idx1 = 0, idx2 = 0;
        while (idx1 == idx2) {
            idx1 = (int)draw(rng, (unsigned long)n_inliers);
            idx2 = (int)draw(rng, (unsigned long)n_inliers);
        }
        const V3 qd = sub3(Q[inliers[idx1]], Q[inliers[idx2]]);
        const double theta = get_angle(ex, qd), theta_v = get_angle(ey, qd);
        const V3 td = sub3(proj[inliers[idx1]], proj[inliers[idx2]]);
        double R[9], Rv[9];
        rotation_matrix(plane, theta, R);
        rotation_matrix(plane, theta_v, Rv);
        const V3 ui = normalize3(mat_vec(R, td)), vi = normalize3(mat_vec(Rv, td));
        for (int k = 0; k < 3; k++) {
            v.v[k] += vi.v[k];
            u.v[k] += ui.v[k];
        }
    }
    for (int k = 0; k < 3; k++) {
        u.v[k] /= iter;
        v.v[k] /= iter;
    }
    V3 w = cross3(u, v);
    if (dot3(plane, w) > 0)
        for (int k = 0; k < 3; k++) w.v[k] = -plane.v[k];
    else
        w = plane;
    w = normalize3(w);
    u = normalize3(u);
    v = normalize3(v);
    v = normalize3(cross3(u, w));
    for (int k = 0; k < 3; k++) {  // pixel spacing of the reference's probe, hard-coded there (:1772-1774)
        u.v[k] *= 0.10763;
        v.v[k] *= 0.10530;
        w.v[k] *= 0.10646;
    }
    V3 tc{{0, 0, 0}}, qc{{0, 0, 0}};
    for (int i = 0; i < n_inliers; i++)
        for (int k = 0; k < 3; k++) {
            tc.v[k] += proj[inliers[i]].v[k];
            qc.v[k] += Q[inliers[i]].v[k];
        }
    for (int k = 0; k < 3; k++) {  // by the number of matches, not of inliers: kept (:1782-1783)
        tc.v[k] /= n;
        qc.v[k] /= n;
    }
    for (int i = 0; i < 16; i++) estimate[i] = (i % 5 == 0) ? 1.0 : 0.0;
    for (int k = 0; k < 3; k++) {
        estimate[4 * k + 0] = u.v[k];
        estimate[4 * k + 1] = v.v[k];
        estimate[4 * k + 2] = w.v[k];
        estimate[4 * k + 3] = tc.v[k] - (u.v[k] * qc.v[0] + v.v[k] * qc.v[1] + w.v[k] * qc.v[2]);
    }
    if (mean_match_distance) {
        double res = 0;
        for (int i = 0; i < n_inliers; i++) {
            V3 q = Q[inliers[i]];
            q.v[2] = 0.0;
            res += norm3(sub3(T[inliers[i]], transform_point(q, estimate)));
        }
        *mean_match_distance = res / n_inliers;
    }
    return SURF_OK;
}

int surf_mean_plane_distance(const uint8_t *mask, int rows, int cols, size_t pitch, int step, const double ground_truth[16],
                             const double estimate[16], double *out)
{
    if (!mask || !ground_truth || !estimate || !out || rows <= 0 || cols <= 0 || step <= 0 || pitch < (size_t)cols) return SURF_ERR_BAD_ARG;
    double res = 0;
    long cnt = 0;
    for (int i = 0; i < rows; i += step)
        for (int j = 0; j < cols; j += step) {
            if (mask[(size_t)i * pitch + j] == 0) continue;
            const V3 p{{(double)i, (double)j, 0.0}};  // (row, column) passed as (x, y), as the reference does
            res += norm3(sub3(transform_point(p, ground_truth), transform_point(p, estimate)));
            cnt++;
        }
    *out = res / (double)cnt;
    return SURF_OK;
}

}  // extern "C"
