/* pose_oracle.c -- CPU restatement of the reference's plane fit and pose estimate.
 *
 * TEST INFRASTRUCTURE ONLY: the checker, never the thing measured or shipped.
 *
 * Follows, statement by statement,
 *   vtkSlicerSurfFeaturesLogic::ransac                  SurfFeatures/Logic/vtkSlicerSurfFeaturesLogic.cxx:2085-2188
 *   vtkSlicerSurfFeaturesLogic::updateMatchNodeRansac   ...Logic.cxx:1655-1851 (the arithmetic; MRML/VTK display calls are not restated)
 *   helpers getRotationMatrix / projectPoint / getPlaneFromPoints / getAngle / transformPoint / computeMeanDistance  ...Logic.cxx:768-838
 *
 * Third-party pieces the reference calls and that are absent from /root/reference (restated from their published
 * definitions; parity on these is unpinned by the reference, which has no tests):
 *   - rand(): glibc's default TYPE_3 additive-feedback generator (stdlib/random_r.c), never seeded by the reference,
 *     i.e. srand(1).  PINNED here against the real libc of this machine (tests/test_oracle_pose.py).
 *   - vnl (VXL, via ITK): vnl_vector_fixed<double,3>::normalize() multiplies by 1/sqrt(sum of squares) when the
 *     sum is non-zero (vnl_c_vector<T>::normalize); two_norm() = sqrt(sum of squares in index order);
 *     dot_product accumulates a[i]*b[i] from 0 in index order; vnl_matrix::normalize_columns() scales every
 *     column by 1/sqrt(its squared norm); matrix * vector accumulates along k from 0.
 *   - vtkMatrix4x4::MultiplyPoint / Multiply4x4: row-times-column sums in index order, double precision.
 *   - acos / cos / sin: this machine's libm (the GPU-side host code calls the same libm; tolerance is stated in the tests).
 * `abs` in the reference resolves to std::abs(double) (the file says `using namespace std`).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* ---- glibc rand(): r[i] = r[i-3] + r[i-31], output = r[i] >> 1; seeding by the Lehmer generator 16807 ---- */
typedef struct orc_rand {
    int32_t r[34];
    int f, b; /* front / rear indices into r[3..33] as glibc keeps them (fptr = rptr + 3) */
} orc_rand;

void orc_srand(orc_rand *s, uint32_t seed)
{
    int32_t *st = s->r; /* glibc's state[] of degree 31: st[0..30] */
    if (seed == 0) seed = 1;
    st[0] = (int32_t)seed;
    for (int i = 1; i < 31; i++) {
        /* word = 16807 * hi_lo by Schrage's method, modulo 2^31 - 1 */
        long hi = st[i - 1] / 127773, lo = st[i - 1] % 127773;
        long word = 16807 * lo - 2836 * hi;
        if (word < 0) word += 2147483647;
        st[i] = (int32_t)word;
    }
    s->f = 3; /* fptr = &state[rand_sep = 3] */
    s->b = 0; /* rptr = &state[0] */
    for (int i = 0; i < 310; i++) { /* discard 10 * degree outputs */
        st[s->f] = (int32_t)((uint32_t)st[s->f] + (uint32_t)st[s->b]);
        if (++s->f >= 31) s->f = 0;
        if (++s->b >= 31) s->b = 0;
    }
}

int orc_rand_next(orc_rand *s)
{
    int32_t *st = s->r;
    const uint32_t v = (uint32_t)st[s->f] + (uint32_t)st[s->b];
    st[s->f] = (int32_t)v;
    if (++s->f >= 31) s->f = 0;
    if (++s->b >= 31) s->b = 0;
    return (int)(v >> 1);
}

/* ---- vnl_double_3 helpers ---- */
typedef struct { double v[3]; } v3;

static v3 sub3(v3 a, v3 b) { v3 r = {{a.v[0] - b.v[0], a.v[1] - b.v[1], a.v[2] - b.v[2]}}; return r; }
static double dot3(v3 a, v3 b)
{
    double ip = 0;
    for (int i = 0; i < 3; i++) ip += a.v[i] * b.v[i];
    return ip;
}
static double norm3(v3 a)
{
    double s = 0;
    for (int i = 0; i < 3; i++) s += a.v[i] * a.v[i];
    return sqrt(s);
}
static v3 normalize3(v3 a)
{
    double t = 0;
    for (int i = 0; i < 3; i++) t += a.v[i] * a.v[i];
    if (t != 0) {
        t = 1.0 / sqrt(t);
        for (int i = 0; i < 3; i++) a.v[i] = t * a.v[i];
    }
    return a;
}
static v3 cross3(v3 a, v3 b)
{
    v3 r = {{a.v[1] * b.v[2] - a.v[2] * b.v[1], a.v[2] * b.v[0] - a.v[0] * b.v[2], a.v[0] * b.v[1] - a.v[1] * b.v[0]}};
    return r;
}

/* ransac() :2085-2188.  points: n x 3 doubles.  inliers gets the indices in the reference's order (the fitted
 * inliers in j order, then the three plane points).  Returns 0, or 1 when n < 3 (nothing written). */
int orc_ransac_plane(orc_rand *rng, const double *points, int n, float threshold, int iterations, int32_t *inliers,
                     int32_t *n_inliers, int32_t plane_idx[3], double *best_error_out, double *n_outliers_out)
{
    *n_inliers = 0;
    plane_idx[0] = plane_idx[1] = plane_idx[2] = -1;
    if (n == 3) {
        for (int i = 0; i < 3; i++) plane_idx[i] = inliers[i] = i;
        *n_inliers = 3;
        return 0;
    }
    if (n < 3) return 1;
    const v3 *P = (const v3 *)points;
    double num_outliers_best = n, best_error = 1.7976931348623157e308;
    int32_t *tmp = (int32_t *)malloc(sizeof(int32_t) * (size_t)n);
    int have = 0;
    for (int it = 0; it < iterations; it++) {
        int idx[3] = {0, 0, 0};
        while (idx[0] == idx[1] || idx[1] == idx[2] || idx[0] == idx[2]) {
            idx[0] = (int)((unsigned long)orc_rand_next(rng) % (unsigned long)n);
            idx[1] = (int)((unsigned long)orc_rand_next(rng) % (unsigned long)n);
            idx[2] = (int)((unsigned long)orc_rand_next(rng) % (unsigned long)n);
        }
        const v3 p1 = P[idx[0]], p2 = P[idx[1]], p3 = P[idx[2]];
        v3 v1 = normalize3(sub3(p1, p2)), v2 = normalize3(sub3(p1, p3));
        if (norm3(v1) < 0.99 || norm3(v2) < 0.99) continue;
        if (fabs(dot3(v1, v2)) > 0.99) continue;
        const v3 N = normalize3(cross3(v1, v2));
        const double D = -dot3(p1, N);
        double num_outliers = 0, err = 0;
        int nt = 0;
        for (int j = 0; j < n; j++) {
            if (j == idx[0] || j == idx[1] || j == idx[2]) continue;
            const float dist = (float)fabs(dot3(N, P[j]) + D);
            if (dist > threshold) {
                num_outliers++;
            } else {
                err += dist;
                tmp[nt++] = j;
            }
        }
        if (num_outliers < num_outliers_best || (num_outliers == num_outliers_best && err < best_error)) {
            num_outliers_best = num_outliers;
            best_error = err;
            for (int k = 0; k < 3; k++) plane_idx[k] = idx[k];
            memcpy(inliers, tmp, sizeof(int32_t) * (size_t)nt);
            *n_inliers = nt;
            have = 1;
        }
    }
    free(tmp);
    if (!have) return 2; /* the reference would index an empty planePointsIdx here (undefined behaviour) */
    for (int k = 0; k < 3; k++) inliers[(*n_inliers)++] = plane_idx[k];
    if (best_error_out) *best_error_out = best_error;
    if (n_outliers_out) *n_outliers_out = num_outliers_best;
    return 0;
}

/* vtkMatrix4x4::MultiplyPoint of (p, 1), first three components (transformPoint :816-823) */
static v3 transform_point(v3 p, const double M[16])
{
    v3 r;
    for (int i = 0; i < 3; i++) r.v[i] = p.v[0] * M[4 * i] + p.v[1] * M[4 * i + 1] + p.v[2] * M[4 * i + 2] + 1.0 * M[4 * i + 3];
    return r;
}

/* combinedTransform->Concatenate(A); ->Concatenate(B); GetMatrix() = A * B (vtkTransform is PreMultiply by default);
 * A comes from a std::vector<float> (getVtkMatrixFromVector), B is double. */
void orc_concat(const float A[16], const double B[16], double out[16])
{
    for (int i = 0; i < 4; i++)
        for (int j = 0; j < 4; j++) {
            double s = 0;
            for (int k = 0; k < 4; k++) s += (double)A[4 * i + k] * B[4 * k + j];
            out[4 * i + j] = s;
        }
}

/* trainPoints / queryPoints of updateMatchNodeRansac :1672-1703: one pair per match whose imgIdx is not -1.
 * matches: n x {queryIdx, trainIdx, imgIdx, distance}; train keypoints packed with per-image offsets; keypoints
 * are 28-byte cv::KeyPoint records (x, y first).  Returns the number of pairs. */
int orc_match_points(const int32_t *matches, int n_matches, const float *query_kps, const float *train_kps,
                     const int32_t *train_offsets, const float *train_transforms, const double image_to_probe[16],
                     double *train_points, double *query_points)
{
    int m = 0;
    for (int j = 0; j < n_matches; j++) {
        const int q = matches[4 * j], t = matches[4 * j + 1], img = matches[4 * j + 2];
        if (img == -1) continue;
        double M[16];
        orc_concat(train_transforms + 16 * (size_t)img, image_to_probe, M);
        const float *tk = train_kps + 7 * ((size_t)train_offsets[img] + t);
        const v3 tp = {{tk[0], tk[1], 0.0}};
        const v3 r = transform_point(tp, M);
        memcpy(train_points + 3 * (size_t)m, r.v, sizeof(r.v));
        const float *qk = query_kps + 7 * (size_t)q;
        query_points[3 * (size_t)m] = qk[0];
        query_points[3 * (size_t)m + 1] = qk[1];
        query_points[3 * (size_t)m + 2] = 0.0;
        m++;
    }
    return m;
}

static void rotation_matrix(v3 axis, double theta, double R[9])
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
    for (int j = 0; j < 3; j++) { /* normalize_columns */
        double nrm = 0;
        for (int i = 0; i < 3; i++) nrm += R[3 * i + j] * R[3 * i + j];
        if (nrm != 0) {
            const double sc = 1.0 / sqrt(nrm);
            for (int i = 0; i < 3; i++) R[3 * i + j] *= sc;
        }
    }
}
static v3 mat_vec(const double R[9], v3 x)
{
    v3 r;
    for (int i = 0; i < 3; i++) {
        double s = 0;
        for (int k = 0; k < 3; k++) s += R[3 * i + k] * x.v[k];
        r.v[i] = s;
    }
    return r;
}
static double get_angle(v3 u, v3 v) { return acos(dot3(u, v) / (norm3(u) * norm3(v))); }

/* The pose part of updateMatchNodeRansac :1714-1822.  query / train points: n x 3; inliers / plane_idx from the
 * plane fit.  estimate: row-major 4x4.  Returns 0. */
int orc_estimate_pose(orc_rand *rng, const double *query_points, const double *train_points, int n, const int32_t *inliers,
                      int n_inliers, const int32_t plane_idx[3], double estimate[16], double *mean_match_distance)
{
    const v3 *Q = (const v3 *)query_points, *T = (const v3 *)train_points;
    /* getPlaneFromPoints :792-803 */
    v3 v1 = normalize3(sub3(T[plane_idx[0]], T[plane_idx[1]])), v2 = normalize3(sub3(T[plane_idx[0]], T[plane_idx[2]]));
    const v3 plane = normalize3(cross3(v1, v2));
    const double d = -dot3(T[plane_idx[0]], plane);
    v3 *proj = (v3 *)malloc(sizeof(v3) * (size_t)n);
    for (int i = 0; i < n; i++) { /* projectPoint :784-790 */
        const v3 nn = normalize3(plane);
        const double dist = dot3(T[i], nn) + d;
        for (int k = 0; k < 3; k++) proj[i].v[k] = T[i].v[k] - dist * nn.v[k];
    }
    const int max_iter = 10000, mult = 20;
    const int iter = n * mult < max_iter ? n * mult : max_iter;
    v3 u = {{0, 0, 0}}, v = {{0, 0, 0}};
    const v3 ex = {{1.0, 0.0, 0.0}}, ey = {{0.0, 1.0, 0.0}};
    for (int i = 0; i < iter; i++) {
        int idx1 = 0, idx2 = 0;
        while (idx1 == idx2) {
            idx1 = (int)((unsigned long)orc_rand_next(rng) % (unsigned long)n_inliers);
            idx2 = (int)((unsigned long)orc_rand_next(rng) % (unsigned long)n_inliers);
        }
        const v3 qd = sub3(Q[inliers[idx1]], Q[inliers[idx2]]);
        const double theta = get_angle(ex, qd), theta_v = get_angle(ey, qd);
        const v3 td = sub3(proj[inliers[idx1]], proj[inliers[idx2]]);
        double R[9], Rv[9];
        rotation_matrix(plane, theta, R);
        rotation_matrix(plane, theta_v, Rv);
        const v3 ui = normalize3(mat_vec(R, td)), vi = normalize3(mat_vec(Rv, td));
        for (int k = 0; k < 3; k++) {
            v.v[k] += vi.v[k];
            u.v[k] += ui.v[k];
        }
    }
    for (int k = 0; k < 3; k++) {
        u.v[k] /= iter;
        v.v[k] /= iter;
    }
    v3 w = cross3(u, v);
    if (dot3(plane, w) > 0)
        for (int k = 0; k < 3; k++) w.v[k] = -plane.v[k];
    else
        w = plane;
    w = normalize3(w);
    u = normalize3(u);
    v = normalize3(v);
    v = normalize3(cross3(u, w));
    for (int k = 0; k < 3; k++) {
        u.v[k] *= 0.10763;
        v.v[k] *= 0.10530;
        w.v[k] *= 0.10646;
    }
    v3 tc = {{0, 0, 0}}, qc = {{0, 0, 0}};
    for (int i = 0; i < n_inliers; i++)
        for (int k = 0; k < 3; k++) {
            tc.v[k] += proj[inliers[i]].v[k];
            qc.v[k] += Q[inliers[i]].v[k];
        }
    for (int k = 0; k < 3; k++) { /* divided by the number of matches, not of inliers, as the reference does */
        tc.v[k] /= n;
        qc.v[k] /= n;
    }
    double t[3];
    for (int k = 0; k < 3; k++) t[k] = tc.v[k] - (u.v[k] * qc.v[0] + v.v[k] * qc.v[1] + w.v[k] * qc.v[2]);
    for (int i = 0; i < 16; i++) estimate[i] = (i % 5 == 0) ? 1.0 : 0.0;
    for (int k = 0; k < 3; k++) {
        estimate[4 * k + 0] = u.v[k];
        estimate[4 * k + 1] = v.v[k];
        estimate[4 * k + 2] = w.v[k];
        estimate[4 * k + 3] = t[k];
    }
    if (mean_match_distance) { /* computeMeanDistance of the inliers :1811-1823 */
        double res = 0;
        for (int i = 0; i < n_inliers; i++) {
            v3 q = Q[inliers[i]];
            q.v[2] = 0.0;
            res += norm3(sub3(T[inliers[i]], transform_point(q, estimate)));
        }
        *mean_match_distance = res / n_inliers;
    }
    free(proj);
    return 0;
}

/* Mean plane-to-plane distance :1826-1846: every step-th mask pixel (i, j) != 0 mapped through both transforms
 * (the reference passes (row, column) as (x, y); kept). */
double orc_mean_plane_distance(const uint8_t *mask, int rows, int cols, size_t pitch, int step, const double ground_truth[16],
                               const double estimate[16])
{
    double res = 0;
    long cnt = 0;
    for (int i = 0; i < rows; i += step)
        for (int j = 0; j < cols; j += step) {
            if (mask[(size_t)i * pitch + j] == 0) continue;
            const v3 p = {{(double)i, (double)j, 0.0}};
            res += norm3(sub3(transform_point(p, ground_truth), transform_point(p, estimate)));
            cnt++;
        }
    return res / (double)cnt;
}
