// surf_db.cu -- the "next" rows f2 / f3 of SURVEY.md section 8 behind the C ABI:
//   * surf_db_*        the multi-image train database of cv::DescriptorMatcher (add / clear / knnMatch),
//                      reference vtkSlicerSurfFeaturesLogic.cxx:1402-1403 and :635-643
//   * surf_correspond  computeNextInterSliceCorrespondence :1434-1571 (k=3 vs train, k=1 vs bogus, bogus-ratio
//                      filter, Hough pose vote, per-image vote count) for all query images of a sweep at once
// The database is device-resident and prepared for the tensor-core matcher once per add(); queries are
// prepared per call.  Kernels: surf_match_tc.cu (k-NN), surf_vote.cu (filter + Hough + votes).
#include <cmath>
#include <cstring>
#include <new>

#include "surf_engine_priv.h"

namespace {

// Device buffers of a host-memory call: [query rows | out records], reusing the engine's match staging area.
int reserve_io(surf_engine *e, size_t bytes)
{
    if (bytes > e->match_io_cap) {
        CU(e, cudaStreamSynchronize(e->stream));
        cudaFree(e->d_match_io);
        e->d_match_io = nullptr;
        e->match_io_cap = 0;
        CU(e, cudaMalloc(&e->d_match_io, bytes));
        e->match_io_cap = bytes;
    }
    return SURF_OK;
}

}  // namespace

namespace surfb200 {

// k-NN of device-resident query rows against a database, records with (img_idx, local train_idx).
int db_knn_device(surf_db *db, const float *dq, int n_query, int k, surf_dmatch *dout, int *launches, bool localize)
{
    surf_engine *e = db->e;
    const long n_train = db_rows(db);
    const int dim = db->dim;
    int rc;
    if (e->match_exact_only || n_train == 0) {
        if ((rc = reserve_match_part(e, n_query, (int)n_train, k))) return rc;
        launch_match_knn(dq, n_query, db->rows, (int)n_train, dim, k, dout, 0, e->d_match_part, e->stream, launches);
    } else {
        surf_engine::TcWs &w = e->tc[2];
        const int max_q = (n_query + 255) & ~255;
        if ((rc = tc_reserve(e, w, 1, max_q, dim))) return rc;
        int sms = 148;
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, e->device);
        const int n_tiles = (int)((n_train + 255) / 256), mt = max_q / 128;
        int splits = (2 * sms + mt - 1) / mt;  // enough (query tile, train split) items for two CTAs per SM
        if (splits > n_tiles) splits = n_tiles;
        if (splits < 1) splits = 1;
        if ((rc = tc_reserve_work(e, w, 1, splits, max_q, 1))) return rc;
        TcSegment seg{dq, nullptr, 0, n_query, 0, 0};
        TcProblem pr{};
        pr.q_slot = 0; pr.t_slot = 0; pr.qsrc = dq; pr.tsrc = db->rows;
        CU(e, cudaMemcpyAsync(w.segs, &seg, sizeof(seg), cudaMemcpyHostToDevice, e->stream));
        CU(e, cudaMemcpyAsync(w.probs, &pr, sizeof(pr), cudaMemcpyHostToDevice, e->stream));
        TcMatchArgs a{};
        a.pool = w.pool; a.tpool = db->pool; a.segs = w.segs; a.n_segs = 1; a.probs = w.probs; a.n_probs = 1;
        a.n_splits = splits; a.max_q_rows = max_q; a.cand = w.cand; a.counters = w.counters; a.redo_list = w.redo;
        a.dim = dim; a.k = k; a.ratio = 0.f; a.out = dout; a.good = nullptr;
        launch_tc_match(a, e->stream, launches);
        e->last_tc_ws = 2;
    }
    if (localize) launch_db_localize(dout, (long)n_query * k, db->d_img_off, (int)db->img_off.size() - 1, e->stream, launches);
    CU(e, cudaGetLastError());
    return SURF_OK;
}

}  // namespace surfb200

extern "C" {

int surf_db_create(surf_engine *e, int dim, long max_rows, int max_images, surf_db **out)
{
    if (!e || !out) return SURF_ERR_BAD_ARG;
    *out = nullptr;
    if (dim != 64 && dim != 128) return fail(e, SURF_ERR_UNSUPPORTED, "descriptor length must be 64 or 128");
    if (max_rows <= 0 || max_images <= 0 || max_rows > (1L << 30))
        return fail(e, SURF_ERR_BAD_ARG, "database capacity must be positive (rows <= 2^30)");
    int rc = ensure_device(e);
    if (rc) return rc;
    surf_db *db = new (std::nothrow) surf_db();
    if (!db) return fail(e, SURF_ERR_CUDA, "out of host memory");
    db->e = e;
    db->dim = dim;
    db->max_rows = max_rows;
    db->max_images = max_images;
    const size_t slot_rows = ((size_t)max_rows + 255) & ~(size_t)255;
    db->pool.slot_rows = (int)slot_rows;
    cudaError_t r = cudaSuccess;
    auto grab = [&](void **p, size_t bytes) { if (r == cudaSuccess) r = cudaMalloc(p, bytes); };
    grab((void **)&db->pool.hi, slot_rows * dim * sizeof(uint16_t));
    grab((void **)&db->pool.lo, slot_rows * dim * sizeof(uint16_t));
    grab((void **)&db->pool.norm, slot_rows * sizeof(float));
    grab((void **)&db->pool.max_norm, sizeof(float));
    grab((void **)&db->pool.count, sizeof(int));
    grab((void **)&db->rows, (size_t)max_rows * dim * sizeof(float));
    grab((void **)&db->kps, (size_t)max_rows * sizeof(surf_keypoint));
    grab((void **)&db->d_img_off, ((size_t)max_images + 1) * sizeof(int));
    grab((void **)&db->d_seg, sizeof(TcSegment));
    if (r != cudaSuccess) {
        surf_db_destroy(db);
        return fail(e, SURF_ERR_CUDA, "database allocation failed: %s", cudaGetErrorString(r));
    }
    *out = db;
    return surf_db_clear(db);
}

void surf_db_destroy(surf_db *db)
{
    if (!db) return;
    cudaSetDevice(db->e->device);
    cudaStreamSynchronize(db->e->stream);
    cudaFree(db->pool.hi); cudaFree(db->pool.lo); cudaFree(db->pool.norm); cudaFree(db->pool.max_norm);
    cudaFree(db->pool.count); cudaFree(db->rows); cudaFree(db->kps); cudaFree(db->d_img_off); cudaFree(db->d_seg);
    cudaFree(db->vote_ws);
    delete db;
}

int surf_db_clear(surf_db *db)
{
    if (!db) return SURF_ERR_BAD_ARG;
    surf_engine *e = db->e;
    int rc = ensure_device(e);
    if (rc) return rc;
    db->img_off.assign(1, 0);
    db->has_kps = true;
    CU(e, cudaMemsetAsync(db->pool.count, 0, sizeof(int), e->stream));
    CU(e, cudaMemsetAsync(db->pool.max_norm, 0, sizeof(float), e->stream));
    CU(e, cudaMemsetAsync(db->d_img_off, 0, sizeof(int), e->stream));
    return SURF_OK;
}

int surf_db_add(surf_db *db, const float *descriptors, const surf_keypoint *keypoints, const int32_t *offsets,
                int n_images, int mem)
{
    if (!db) return SURF_ERR_BAD_ARG;
    surf_engine *e = db->e;
    if (!offsets || n_images < 0) return fail(e, SURF_ERR_BAD_ARG, "null offsets or negative image count");
    if (n_images == 0) return SURF_OK;
    for (int i = 0; i < n_images; i++)
        if (offsets[i + 1] < offsets[i]) return fail(e, SURF_ERR_BAD_ARG, "offsets must be non-decreasing");
    const long n = (long)offsets[n_images] - offsets[0], cur = db_rows(db);
    if (n > 0 && !descriptors) return fail(e, SURF_ERR_BAD_ARG, "null descriptors");
    if ((long)db->img_off.size() - 1 + n_images > db->max_images || cur + n > db->max_rows) {
        e->required = cur + n;
        return fail(e, SURF_ERR_CAPACITY, "database full: %ld rows / %ld images after this add, capacity %ld / %d",
                    cur + n, (long)db->img_off.size() - 1 + n_images, db->max_rows, db->max_images);
    }
    int rc = ensure_device(e);
    if (rc) return rc;
    const cudaMemcpyKind kind = mem == SURF_MEM_DEVICE ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice;
    const size_t first = db->img_off.size();
    for (int i = 0; i < n_images; i++) db->img_off.push_back((int)(cur + offsets[i + 1] - offsets[0]));
    CU(e, cudaMemcpyAsync(db->d_img_off + first, db->img_off.data() + first, n_images * sizeof(int),
                          cudaMemcpyHostToDevice, e->stream));
    if (n > 0) {
        CU(e, cudaMemcpyAsync(db->rows + (size_t)cur * db->dim, descriptors + (size_t)offsets[0] * db->dim,
                              (size_t)n * db->dim * sizeof(float), kind, e->stream));
        if (keypoints)
            CU(e, cudaMemcpyAsync(db->kps + cur, keypoints + offsets[0], (size_t)n * sizeof(surf_keypoint), kind, e->stream));
        else
            db->has_kps = false;
        // BF16 hi/lo + norms of the new rows, appended to the slot (prepared once, matched many times)
        TcSegment seg{db->rows, nullptr, (int)cur, (int)n, 0, (int)cur};
        CU(e, cudaMemcpyAsync(db->d_seg, &seg, sizeof(seg), cudaMemcpyHostToDevice, e->stream));
        launch_tc_prep(db->pool, db->d_seg, 1, (int)n, db->dim, false, e->stream, &e->launches);
        CU(e, cudaGetLastError());
    }
    // the host arrays (offsets table, pageable sources) may be reused by the caller right away
    CU(e, cudaStreamSynchronize(e->stream));
    return SURF_OK;
}

int surf_db_size(const surf_db *db, long *rows, int *images)
{
    if (!db) return SURF_ERR_BAD_ARG;
    if (rows) *rows = db_rows(db);
    if (images) *images = (int)db->img_off.size() - 1;
    return SURF_OK;
}

int surf_db_knn(surf_db *db, const float *query, int n_query, int k, surf_dmatch *out, int mem)
{
    if (!db) return SURF_ERR_BAD_ARG;
    surf_engine *e = db->e;
    if (n_query < 0) return fail(e, SURF_ERR_BAD_ARG, "negative row count");
    if (k < 1 || k > 4) return fail(e, SURF_ERR_UNSUPPORTED, "k must be in 1..4");
    if (n_query == 0) return SURF_OK;
    if (!query || !out) return fail(e, SURF_ERR_BAD_ARG, "null pointer");
    int rc = ensure_device(e);
    if (rc) return rc;
    e->launches = 0;
    const float *dq = query;
    surf_dmatch *dout = out;
    const size_t bq = ((size_t)n_query * db->dim * sizeof(float) + 255) & ~(size_t)255;
    const size_t bo = (size_t)n_query * k * sizeof(surf_dmatch);
    if (mem == SURF_MEM_HOST) {
        if ((rc = reserve_io(e, bq + bo))) return rc;
        dq = (const float *)e->d_match_io;
        dout = (surf_dmatch *)((char *)e->d_match_io + bq);
        CU(e, cudaMemcpyAsync((void *)dq, query, (size_t)n_query * db->dim * sizeof(float), cudaMemcpyHostToDevice,
                              e->stream));
    }
    if ((rc = db_knn_device(db, dq, n_query, k, dout, &e->launches))) return rc;
    if (mem == SURF_MEM_HOST) {
        CU(e, cudaMemcpyAsync(out, dout, bo, cudaMemcpyDeviceToHost, e->stream));
        CU(e, cudaStreamSynchronize(e->stream));
    }
    return SURF_OK;
}

int surf_correspond(surf_db *train, surf_db *bogus, const float *query_descriptors,
                    const surf_keypoint *query_keypoints, const int32_t *q_offsets, int n_query_images,
                    const surf_vote_params *params, surf_dmatch *matches, surf_vote_result *results, int mem)
{
    if (!train || !bogus) return SURF_ERR_BAD_ARG;
    surf_engine *e = train->e;
    if (bogus->e != e || bogus->dim != train->dim)
        return fail(e, SURF_ERR_BAD_ARG, "train and bogus databases must belong to one engine and share the descriptor length");
    if (!q_offsets || !params || !results || n_query_images < 0) return fail(e, SURF_ERR_BAD_ARG, "null pointer");
    if (n_query_images == 0) return SURF_OK;
    int max_rows = 0;
    for (int i = 0; i < n_query_images; i++) {
        if (q_offsets[i + 1] < q_offsets[i] || q_offsets[0] != 0)
            return fail(e, SURF_ERR_BAD_ARG, "q_offsets must start at 0 and be non-decreasing");
        if (q_offsets[i + 1] - q_offsets[i] > max_rows) max_rows = q_offsets[i + 1] - q_offsets[i];
    }
    const long n = q_offsets[n_query_images];
    if (n > 0 && (!query_descriptors || !query_keypoints || !matches)) return fail(e, SURF_ERR_BAD_ARG, "null pointer");
    // the reference indexes mmatches[j][0] and mmatchesBogus[j][0] unconditionally: an empty database is an error there
    if (db_rows(train) == 0 || db_rows(bogus) == 0)
        return fail(e, SURF_ERR_BAD_ARG, "surf_correspond needs a non-empty train and bogus database");
    if (!train->has_kps) return fail(e, SURF_ERR_BAD_ARG, "the train database was filled without keypoints");
    int rc = ensure_device(e);
    if (rc) return rc;
    e->launches = 0;
    const int D = train->dim, n_img = (int)train->img_off.size() - 1;
    // scratch layout (256-byte aligned pieces)
    auto al = [](size_t b) { return (b + 255) & ~(size_t)255; };
    const size_t b_desc = mem == SURF_MEM_HOST ? al((size_t)n * D * sizeof(float)) : 0;
    const size_t b_kps = mem == SURF_MEM_HOST ? al((size_t)n * sizeof(surf_keypoint)) : 0;
    const size_t b_out = mem == SURF_MEM_HOST ? al((size_t)n * sizeof(surf_dmatch)) : 0;
    const size_t b_res = mem == SURF_MEM_HOST ? al((size_t)n_query_images * sizeof(surf_vote_result)) : 0;
    const size_t b_off = al(((size_t)n_query_images + 1) * sizeof(int));
    const size_t b_m3 = al((size_t)n * 3 * sizeof(surf_dmatch)), b_m1 = al((size_t)n * sizeof(surf_dmatch));
    const size_t b_booth = al((size_t)n * sizeof(float4)), b_thr = al((size_t)n * sizeof(float));
    const size_t b_cnt = al((size_t)n_query_images * sizeof(int)), b_keys = al((size_t)n_query_images * 8);
    const size_t b_votes = al((size_t)n_query_images * (size_t)(n_img > 0 ? n_img : 1) * sizeof(int));
    const size_t need = b_desc + b_kps + b_out + b_res + b_off + b_m3 + b_m1 + b_booth + b_thr + b_cnt + b_keys + b_votes;
    if (need > train->vote_ws_cap) {
        CU(e, cudaStreamSynchronize(e->stream));
        cudaFree(train->vote_ws);
        train->vote_ws = nullptr;
        train->vote_ws_cap = 0;
        CU(e, cudaMalloc(&train->vote_ws, need));
        train->vote_ws_cap = need;
    }
    char *p = (char *)train->vote_ws;
    auto take = [&](size_t b) { char *r = p; p += b; return r; };
    const float *dq = query_descriptors;
    const surf_keypoint *dk = query_keypoints;
    surf_dmatch *dout = matches;
    surf_vote_result *dres = results;
    if (mem == SURF_MEM_HOST) {
        float *q2 = (float *)take(b_desc);
        surf_keypoint *k2 = (surf_keypoint *)take(b_kps);
        dout = (surf_dmatch *)take(b_out);
        dres = (surf_vote_result *)take(b_res);
        if (n) {
            CU(e, cudaMemcpyAsync(q2, query_descriptors, (size_t)n * D * sizeof(float), cudaMemcpyHostToDevice, e->stream));
            CU(e, cudaMemcpyAsync(k2, query_keypoints, (size_t)n * sizeof(surf_keypoint), cudaMemcpyHostToDevice, e->stream));
        }
        dq = q2;
        dk = k2;
    }
    int *d_off = (int *)take(b_off);
    surf_dmatch *m3 = (surf_dmatch *)take(b_m3), *m1 = (surf_dmatch *)take(b_m1);
    float4 *booth = (float4 *)take(b_booth);
    float *booth_thr = (float *)take(b_thr);
    int *n_matches = (int *)take(b_cnt);
    unsigned long long *keys = (unsigned long long *)take(b_keys);
    int *votes = (int *)take(b_votes);
    CU(e, cudaMemcpyAsync(d_off, q_offsets, ((size_t)n_query_images + 1) * sizeof(int), cudaMemcpyHostToDevice, e->stream));
    if (n) {
        // mmatches = knnMatch(query, train, 3); mmatchesBogus = knnMatch(query, bogus, 1)   (:1457-1458)
        if ((rc = db_knn_device(train, dq, (int)n, 3, m3, &e->launches))) return rc;
        if ((rc = db_knn_device(bogus, dq, (int)n, 1, m1, &e->launches))) return rc;
    }
    VoteArgs a{};
    a.m_train = m3; a.m_bogus = m1; a.qk = dk; a.tk = train->kps; a.t_img_off = train->d_img_off; a.q_off = d_off;
    a.n_query_images = n_query_images; a.n_train_images = n_img; a.max_rows_per_image = max_rows;
    a.bogus_ratio = params->bogus_ratio; a.ref_x = params->ref_x; a.ref_y = params->ref_y;
    a.thr_scale = (float)log((double)1.5f);
    a.thr_ori = (float)(20.0f * 3.141592653589793 / 180.0F);
    a.out = dout; a.booth = booth; a.booth_thr = booth_thr; a.n_matches = n_matches; a.keys = keys; a.votes = votes;
    a.results = dres;
    launch_vote(a, e->stream, &e->launches);
    CU(e, cudaGetLastError());
    if (mem == SURF_MEM_HOST) {
        if (n) CU(e, cudaMemcpyAsync(matches, dout, (size_t)n * sizeof(surf_dmatch), cudaMemcpyDeviceToHost, e->stream));
        CU(e, cudaMemcpyAsync(results, dres, (size_t)n_query_images * sizeof(surf_vote_result), cudaMemcpyDeviceToHost,
                              e->stream));
    }
    // q_offsets is a pageable host array: make sure its upload has been consumed before returning
    CU(e, cudaStreamSynchronize(e->stream));
    return SURF_OK;
}

}  // extern "C"
