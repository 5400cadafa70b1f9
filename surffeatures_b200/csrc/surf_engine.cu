// surf_engine.cu -- host side of libsurf_b200.so: the C ABI declared in include/surf_b200.h.
//
// Replaces, for the calls at SurfFeatures/Logic/vtkSlicerSurfFeaturesLogic.cxx:1384-1389 and :640,
// what OpenCV 2.4.5's cv::SURF / BFMatcher do on the CPU.  One engine = one device + one stream +
// one workspace; every stage is a kernel launch on that stream (surf_kernels.cu, surf_match.cu).
// There is no CPU path: without an sm_100 device surf_create fails with SURF_ERR_NO_DEVICE.
#include <cfloat>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <utility>
#include <ctime>
#include <new>
#include <vector>

#include <cuda.h>

#include "surf_engine_priv.h"

using namespace surfb200;

namespace surfb200 {

int fail(surf_engine *e, int code, const char *fmt, ...)
{
    if (e) {
        va_list ap;
        va_start(ap, fmt);
        vsnprintf(e->err, sizeof(e->err), fmt, ap);
        va_end(ap);
    }
    return code;
}

int ensure_device(surf_engine *e)
{
    CU(e, cudaSetDevice(e->device));
    return SURF_OK;
}

// (Re)allocates the prepared-descriptor pool of a tensor-core workspace.
int tc_reserve(surf_engine *e, surf_engine::TcWs &w, int slots, int slot_rows, int dim)
{
    if (w.pool.hi && slots <= w.slots && slot_rows <= w.pool.slot_rows && dim <= w.dim) {
        return SURF_OK;
    }
    CU(e, cudaStreamSynchronize(e->stream));
    CU(e, cudaStreamSynchronize(e->match_stream));
    const int old_slot_rows = w.pool.hi ? w.pool.slot_rows : 0;
    cudaFree(w.pool.hi); cudaFree(w.pool.lo); cudaFree(w.pool.norm); cudaFree(w.pool.max_norm); cudaFree(w.pool.count);
    w.pool = TcPool();
    if (slots < w.slots) slots = w.slots;
    if (slot_rows < old_slot_rows) slot_rows = old_slot_rows;
    if (dim < w.dim) dim = w.dim;
    const size_t rows = (size_t)slots * slot_rows;
    CU(e, cudaMalloc(&w.pool.hi, rows * dim * sizeof(uint16_t)));
    CU(e, cudaMalloc(&w.pool.lo, rows * dim * sizeof(uint16_t)));
    CU(e, cudaMalloc(&w.pool.norm, rows * sizeof(float)));
    CU(e, cudaMalloc(&w.pool.max_norm, slots * sizeof(float)));
    CU(e, cudaMalloc(&w.pool.count, slots * sizeof(int)));
    CU(e, cudaMemset(w.pool.count, 0, slots * sizeof(int)));
    CU(e, cudaMemset(w.pool.max_norm, 0, slots * sizeof(float)));
    w.pool.slot_rows = slot_rows;
    w.slots = slots;
    w.dim = dim;
    return SURF_OK;
}

int tc_reserve_work(surf_engine *e, surf_engine::TcWs &w, int n_probs, int splits, int max_q, int n_segs)
{
    const size_t need = (size_t)n_probs * splits * max_q;  // one merged candidate list per query and train split
    if (need > w.cand_cap) {
        CU(e, cudaStreamSynchronize(e->stream));
        cudaFree(w.cand);
        w.cand = nullptr;
        w.cand_cap = 0;
        CU(e, cudaMalloc(&w.cand, need * sizeof(TcCand)));
        w.cand_cap = need;
    }
    const size_t need_redo = (size_t)n_probs * max_q;
    if (need_redo > w.redo_cap) {
        CU(e, cudaStreamSynchronize(e->stream));
        cudaFree(w.redo);
        w.redo = nullptr;
        w.redo_cap = 0;
        CU(e, cudaMalloc(&w.redo, need_redo * 2 * sizeof(int)));
        w.redo_cap = need_redo;
    }
    const int tab = n_probs > n_segs ? n_probs : n_segs;
    if (tab > w.table_cap) {
        CU(e, cudaStreamSynchronize(e->stream));
        cudaFree(w.segs);
        cudaFree(w.probs);
        cudaFree(w.counters);
        w.segs = nullptr;
        w.probs = nullptr;
        w.counters = nullptr;
        CU(e, cudaMalloc(&w.counters, (2 + (size_t)tab) * sizeof(int)));  // work head, redo total, redo entries per problem
        CU(e, cudaMemset(w.counters, 0, (2 + (size_t)tab) * sizeof(int)));
        CU(e, cudaMalloc(&w.segs, tab * sizeof(TcSegment)));
        CU(e, cudaMalloc(&w.probs, tab * sizeof(TcProblem)));
        w.table_cap = tab;
    }
    return SURF_OK;
}

int reserve_match_part(surf_engine *e, int n_query, int n_train, int k)
{
    const size_t need_part = match_scratch_records(n_query, n_train > 0 ? n_train : 1, k);
    if (need_part > e->match_part_cap) {
        CU(e, cudaStreamSynchronize(e->stream));
        cudaFree(e->d_match_part);
        e->d_match_part = nullptr;
        e->match_part_cap = 0;
        CU(e, cudaMalloc(&e->d_match_part, need_part * sizeof(surf_dmatch)));
        e->match_part_cap = need_part;
    }
    return SURF_OK;
}

}  // namespace surfb200

namespace {

inline int cv_round(double v) { return (int)lrint(v); }

// resizeHaarPattern (nonfree/surf.cpp): coordinates rounded half-to-even, fp32 weight.
void scale_boxes(const int (*src)[5], Box *dst, int n, int old_size, int new_size)
{
    const float ratio = (float)new_size / old_size;
    for (int k = 0; k < n; k++) {
        dst[k].x1 = cv_round(ratio * src[k][0]);
        dst[k].y1 = cv_round(ratio * src[k][1]);
        dst[k].x2 = cv_round(ratio * src[k][2]);
        dst[k].y2 = cv_round(ratio * src[k][3]);
        dst[k].w = src[k][4] / ((float)(dst[k].x2 - dst[k].x1) * (dst[k].y2 - dst[k].y1));
    }
}

const int kDxx[3][5] = {{0, 2, 3, 7, 1}, {3, 2, 6, 7, -2}, {6, 2, 9, 7, 1}};
const int kDyy[3][5] = {{2, 0, 7, 3, 1}, {2, 3, 7, 6, -2}, {2, 6, 7, 9, 1}};
const int kDxy[4][5] = {{1, 1, 4, 4, 1}, {5, 1, 8, 4, -1}, {1, 5, 4, 8, -1}, {5, 5, 8, 8, 1}};
const int kDm[1][5] = {{0, 0, 9, 9, 1}};

OctavePat make_octave(int octave, int lpo, int W, int H)
{
    OctavePat p{};
    p.step = 1 << octave;
    p.rows = H / p.step;
    p.cols = W / p.step;
    p.lpo = lpo;
    for (int l = 0; l < lpo; l++) {
        LayerPat &L = p.L[l];
        L.size = (9 + 6 * l) << octave;
        L.m = (L.size / 2) / p.step;
        L.skip = (L.size > H || L.size > W) ? 1 : 0;
        L.ni = L.skip ? 0 : 1 + (H - L.size) / p.step;
        L.nj = L.skip ? 0 : 1 + (W - L.size) / p.step;
        scale_boxes(kDxx, L.dxx, 3, 9, L.size);
        scale_boxes(kDyy, L.dyy, 3, 9, L.size);
        scale_boxes(kDxy, L.dxy, 4, 9, L.size);
        scale_boxes(kDm, &L.mask, 1, 9, L.size);
    }
    return p;
}

// Shared-memory tile geometry and corner offsets of the fused Hessian + maxima kernel for one octave.
bool make_tile_octave(const OctavePat &pat, TileOctave *out)
{
    TileOctave t{};
    t.step = pat.step;
    t.rows = pat.rows;
    t.cols = pat.cols;
    t.lpo = pat.lpo;
    t.m_max = 0;
    for (int l = 0; l < pat.lpo; l++) t.m_max = pat.L[l].m > t.m_max ? pat.L[l].m : t.m_max;
    int tw = 0, th = 0;
    for (int l = 0; l < pat.lpo; l++) {
        const int ex = (kTileW - 1 + t.m_max - pat.L[l].m) * pat.step + pat.L[l].size + 1;
        const int ey = (kTileH - 1 + t.m_max - pat.L[l].m) * pat.step + pat.L[l].size + 1;
        tw = ex > tw ? ex : tw;
        th = ey > th ? ey : th;
    }
    t.tile_p = (tw + 3 + 3) & ~3;  // + up to 3 columns: the TMA start column is aligned down to 16 bytes
    t.tile_h = th;
    t.tile_bytes = t.tile_p * t.tile_h * 4;
    if (t.tile_p > 256 || t.tile_h > 256) return false;  // TMA box limit
    if ((size_t)t.tile_bytes + (size_t)pat.lpo * kTileW * kTileH * 4 > 190 * 1024) return false;
    for (int l = 0; l < pat.lpo; l++) {
        const LayerPat &P = pat.L[l];
        TileLayer &T = t.L[l];
        T.size = P.size; T.m = P.m; T.ni = P.ni; T.nj = P.nj; T.skip = P.skip; T.mask = P.mask;
        // the three patterns share their box edges (same cvRound of the same products)
        if (P.dxx[1].x1 != P.dxx[0].x2 || P.dxx[2].x1 != P.dxx[1].x2 || P.dyy[1].y1 != P.dyy[0].y2 ||
            P.dyy[2].y1 != P.dyy[1].y2 || P.dxy[2].x1 != P.dxy[0].x1 || P.dxy[3].x1 != P.dxy[1].x1 ||
            P.dxy[1].y1 != P.dxy[0].y1 || P.dxy[3].y1 != P.dxy[2].y1 || P.dxy[2].x2 != P.dxy[0].x2 ||
            P.dxy[3].x2 != P.dxy[1].x2 || P.dxy[1].y2 != P.dxy[0].y2 || P.dxy[3].y2 != P.dxy[2].y2 ||
            P.dxx[1].y1 != P.dxx[0].y1 || P.dxx[2].y2 != P.dxx[0].y2 || P.dyy[1].x1 != P.dyy[0].x1 ||
            P.dyy[2].x2 != P.dyy[0].x2)
            return false;
        const int xx[4] = {P.dxx[0].x1, P.dxx[0].x2, P.dxx[1].x2, P.dxx[2].x2}, xy[2] = {P.dxx[0].y1, P.dxx[0].y2};
        for (int r = 0; r < 2; r++)
            for (int c = 0; c < 4; c++) T.dxx[r * 4 + c] = xy[r] * t.tile_p + xx[c];
        const int yy[4] = {P.dyy[0].y1, P.dyy[0].y2, P.dyy[1].y2, P.dyy[2].y2}, yx[2] = {P.dyy[0].x1, P.dyy[0].x2};
        for (int r = 0; r < 4; r++)
            for (int c = 0; c < 2; c++) T.dyy[r * 2 + c] = yy[r] * t.tile_p + yx[c];
        const int cx[4] = {P.dxy[0].x1, P.dxy[0].x2, P.dxy[1].x1, P.dxy[1].x2};
        const int cy[4] = {P.dxy[0].y1, P.dxy[0].y2, P.dxy[2].y1, P.dxy[2].y2};
        for (int r = 0; r < 4; r++)
            for (int c = 0; c < 4; c++) T.dxy[r * 4 + c] = cy[r] * t.tile_p + cx[c];
        for (int k = 0; k < 3; k++) {
            T.w[k] = P.dxx[k].w;
            T.w[3 + k] = P.dyy[k].w;
        }
        for (int k = 0; k < 4; k++) T.w[6 + k] = P.dxy[k].w;
    }
    *out = t;
    return true;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_tiled_fn()
{
    static EncodeTiledFn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
        cudaGetLastError();
    }
    return fn;
}

// (Re)builds the TMA descriptors of the integral image for octaves 0 and 1 (3-D: x, y, frame).
void ensure_tensor_maps(surf_engine *e, const FrameGeo &geo, int B)
{
    const int layers = e->params.n_octave_layers;
    if (e->tmap_w == geo.W && e->tmap_h == geo.H && e->tmap_b == B && e->tmap_layers == layers) return;
    e->tmap_w = geo.W; e->tmap_h = geo.H; e->tmap_b = B; e->tmap_layers = layers;
    e->tmap_ok[0] = e->tmap_ok[1] = false;
    EncodeTiledFn enc = encode_tiled_fn();
    if (!enc || !e->use_tma) return;
    for (int o = 0; o < 2 && o < e->params.n_octaves; o++) {
        const OctavePat pat = make_octave(o, layers + 2, geo.W, geo.H);
        if (!make_tile_octave(pat, &e->tile_oct[o])) continue;
        const TileOctave &t = e->tile_oct[o];
        cuuint64_t dims[3] = {(cuuint64_t)(geo.W + 1), (cuuint64_t)(geo.H + 1), (cuuint64_t)B};
        cuuint64_t strides[2] = {(cuuint64_t)geo.SP * 4, (cuuint64_t)geo.sum_frame * 4};
        cuuint32_t box[3] = {(cuuint32_t)t.tile_p, (cuuint32_t)t.tile_h, 1};
        cuuint32_t estr[3] = {1, 1, 1};
        CUresult r = enc(&e->tmap[o], CU_TENSOR_MAP_DATA_TYPE_UINT32, 3, e->d_sum, dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        e->tmap_ok[o] = (r == CUDA_SUCCESS);
    }
}

// TMA descriptors of the frames (3-D: x, y, frame; uint8) for the descriptor kernels' image tiles.
void ensure_image_maps(surf_engine *e, const ImageRef &img, const FrameGeo &geo, int B)
{
    if (e->imap_ptr == img.ptr && e->imap_pitch == img.pitch && e->imap_stride == img.frame_stride && e->imap_w == geo.W &&
        e->imap_h == geo.H && e->imap_b == B)
        return;
    e->imap_ptr = img.ptr; e->imap_pitch = img.pitch; e->imap_stride = img.frame_stride;
    e->imap_w = geo.W; e->imap_h = geo.H; e->imap_b = B;
    e->imap_ok = false;
    EncodeTiledFn enc = encode_tiled_fn();
    const bool aligned16 = ((reinterpret_cast<uintptr_t>(img.ptr) | img.pitch | img.frame_stride) & 15) == 0;
    if (!enc || !e->use_tma || !aligned16) return;
    bool ok = true;
    for (int c = 0; c < kNumClasses; c++) {
        int tp = 0, tr = 0;
        describe_tile_box(c, &tp, &tr);
        cuuint64_t dims[3] = {(cuuint64_t)geo.W, (cuuint64_t)geo.H, (cuuint64_t)B};
        cuuint64_t strides[2] = {(cuuint64_t)img.pitch, (cuuint64_t)(B > 1 ? img.frame_stride : img.pitch * geo.H)};
        cuuint32_t box[3] = {(cuuint32_t)tp, (cuuint32_t)tr, 1};
        cuuint32_t estr[3] = {1, 1, 1};
        if ((strides[1] & 15) || tp > 256 || tr > 256) { ok = false; break; }
        CUresult r = enc(&e->imap[c], CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, const_cast<uint8_t *>(img.ptr), dims, strides, box,
                         estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { ok = false; break; }
    }
    e->imap_ok = ok;
}

size_t det_octave_words(const surf_params &p, int octave, int W, int H)
{
    return (size_t)(p.n_octave_layers + 2) * (size_t)((H >> octave)) * (size_t)((W >> octave));
}

FrameGeo make_geo(int W, int H)
{
    FrameGeo g;
    g.W = W;
    g.H = H;
    g.SP = (W + 1 + 3) & ~3;
    g.sum_frame = (size_t)(H + 1) * g.SP;
    return g;
}

int check_params(surf_engine *e, const surf_params *p)
{
    if (!p) return fail(e, SURF_ERR_BAD_ARG, "params is null");
    if (!(p->hessian_threshold >= 0)) return fail(e, SURF_ERR_BAD_ARG, "hessianThreshold must be >= 0");
    if (p->n_octaves <= 0 || p->n_octaves > kMaxOctaves)
        return fail(e, SURF_ERR_BAD_ARG, "nOctaves must be in 1..%d", kMaxOctaves);
    if (p->n_octave_layers <= 0 || p->n_octave_layers + 2 > kMaxLayersPerOctave)
        return fail(e, SURF_ERR_BAD_ARG, "nOctaveLayers must be in 1..%d", kMaxLayersPerOctave - 2);
    return SURF_OK;
}

// The buffers named d_out_kp / d_out_desc / d_off_out / h_pin are those of output set s from here on.
void bind_out(surf_engine *e, int s)
{
    e->out_bound = s;
    e->d_out_kp = e->out[s].d_kp;
    e->d_out_desc = e->out[s].d_desc;
    e->d_off_out = e->out[s].d_off;
    e->h_pin = e->out[s].h_pin;
}

// A stage-event array no batch in flight is using.
cudaEvent_t *next_ev_array(surf_engine *e)
{
    for (int t = 0; t < surf_engine::kEvSets; t++) {
        cudaEvent_t *a = e->evpool[e->ev_next];
        e->ev_next = (e->ev_next + 1) % surf_engine::kEvSets;
        bool used = (e->la.valid && e->la_ev == a);
        for (const OutSet &o : e->out) used = used || ((o.active || o.collected) && o.ev == a);
        if (!used) return a;
    }
    return e->evpool[0];
}

// The single-shot calls (surf_detect, surf_integral, ...) abandon batches that were submitted and not collected.
void drop_batches(surf_engine *e)
{
    for (OutSet &o : e->out) o.active = o.collected = false;
    e->out_collected = -1;
}

void drop_graphs(surf_engine *e);

void free_workspace(surf_engine *e)
{
    drop_graphs(e);  // they hold the addresses of the buffers freed below
    for (auto &p : e->imgbuf) { cudaFree(p); p = nullptr; }
    cudaFree(e->d_mask); cudaFree(e->d_sum); cudaFree(e->d_msum); cudaFree(e->d_bandsum); cudaFree(e->d_bandflags);
    cudaFree(e->d_det); cudaFree(e->d_kp_a); cudaFree(e->d_kp_b); cudaFree(e->d_desc);
    cudaFree(e->d_patches); cudaFree(e->d_counts); cudaFree(e->d_ndel);
    cudaFree(e->d_off_in); cudaFree(e->d_misc); cudaFree(e->d_deleted);
    for (OutSet &o : e->out) {
        cudaFree(o.d_kp); cudaFree(o.d_desc); cudaFree(o.d_off);
        if (o.h_pin) cudaFreeHost(o.h_pin);
        o.d_kp = nullptr; o.d_desc = nullptr; o.d_off = nullptr; o.h_pin = nullptr;
        o.active = o.collected = false;
    }
    cudaFree(e->d_match_part); cudaFree(e->d_match_io); cudaFree(e->d_class_list); cudaFree(e->d_sincos);
    for (auto &w : e->tc) {
        cudaFree(w.pool.hi); cudaFree(w.pool.lo); cudaFree(w.pool.norm); cudaFree(w.pool.max_norm);
        cudaFree(w.pool.count); cudaFree(w.cand); cudaFree(w.redo); cudaFree(w.counters); cudaFree(w.segs);
        cudaFree(w.probs);
        w = surf_engine::TcWs();
    }
    cudaFree(e->d_prev_desc); cudaFree(e->d_prev_off); cudaFree(e->d_stream_matches); cudaFree(e->d_stream_good);
    e->d_prev_desc = nullptr; e->d_prev_off = nullptr; e->d_stream_matches = nullptr; e->d_stream_good = nullptr;
    cudaFree(e->alt.d_sum); cudaFree(e->alt.d_kp_a); cudaFree(e->alt.d_kp_b); cudaFree(e->alt.d_counts);
    e->alt.d_sum = nullptr; e->alt.d_kp_a = e->alt.d_kp_b = e->alt.sorted = nullptr; e->alt.d_counts = nullptr;
    e->alt.tmap_w = 0; e->alt.free_recorded = false;
    e->la.valid = false;
    e->d_mask = nullptr; e->prefetched.valid = false; e->buf_staged = -1; e->d_sum = e->d_msum = e->d_bandsum = nullptr; e->d_bandflags = nullptr; e->d_det = nullptr;
    e->d_kp_a = e->d_kp_b = e->d_out_kp = nullptr; e->d_desc = e->d_out_desc = nullptr; e->d_patches = nullptr;
    e->d_counts = e->d_ndel = e->d_off_in = e->d_off_out = e->d_misc = e->d_deleted = nullptr;
    e->d_match_part = nullptr; e->d_match_io = nullptr; e->h_pin = nullptr; e->d_class_list = nullptr;
    e->d_sincos = nullptr;
    e->match_part_cap = e->match_io_cap = 0;
}

int alloc_det_and_desc(surf_engine *e)
{
    // det pyramid and descriptor buffers depend on the parameters; (re)allocate when they change
    size_t words = 0;
    for (int o = 0; o < e->params.n_octaves; o++) words += det_octave_words(e->params, o, e->max_w, e->max_h);
    words *= (size_t)e->max_b;
    if (words > e->det_words) {
        cudaFree(e->d_det);
        e->d_det = nullptr;
        CU(e, cudaMalloc(&e->d_det, (words ? words : 1) * sizeof(float)));
        e->det_words = words;
    }
    const int D = e->params.extended ? 128 : 64;
    if (D > e->desc_dim) {
        cudaFree(e->d_desc);
        e->d_desc = nullptr;
        const size_t n = (size_t)e->max_b * e->cap * D;
        CU(e, cudaMalloc(&e->d_desc, n * sizeof(float)));
        for (OutSet &o : e->out) {
            cudaFree(o.d_desc);
            o.d_desc = nullptr;
            CU(e, cudaMalloc(&o.d_desc, n * sizeof(float)));
        }
        e->desc_dim = D;
        bind_out(e, e->out_bound);
    }
    return SURF_OK;
}

int alloc_workspace(surf_engine *e)
{
    const int W = e->max_w, H = e->max_h, B = e->max_b, cap = e->cap;
    const FrameGeo g = make_geo(W, H);
    e->img_pitch = (size_t)((W + 15) & ~15);
    CU(e, cudaMalloc(&e->imgbuf[0], e->img_pitch * H * B));  // the other two on first use (prefetch / look-ahead)
    CU(e, cudaMalloc(&e->d_mask, e->img_pitch * H));
    CU(e, cudaMalloc(&e->d_sum, g.sum_frame * sizeof(uint32_t) * B));
    CU(e, cudaMalloc(&e->d_msum, g.sum_frame * sizeof(uint32_t)));
    CU(e, cudaMalloc(&e->d_bandsum, integral_bandsum_words(W, H, B) * sizeof(uint32_t)));
    CU(e, cudaMalloc(&e->d_bandflags, integral_flag_words(H, B) * sizeof(int)));
    CU(e, cudaMemset(e->d_bandflags, 0, integral_flag_words(H, B) * sizeof(int)));
    const size_t nk = (size_t)B * cap;
    CU(e, cudaMalloc(&e->d_kp_a, nk * sizeof(surf_keypoint)));
    CU(e, cudaMalloc(&e->d_kp_b, nk * sizeof(surf_keypoint)));
    for (OutSet &o : e->out) {
        CU(e, cudaMalloc(&o.d_kp, nk * sizeof(surf_keypoint)));
        CU(e, cudaMalloc(&o.d_off, sizeof(int) * (B + 1)));
        CU(e, cudaMallocHost(&o.h_pin, sizeof(int) * (3 * (size_t)B + 8 + kMiscWords)));
    }
    CU(e, cudaMalloc(&e->d_deleted, nk * sizeof(int)));
    CU(e, cudaMalloc(&e->d_counts, sizeof(int) * (B + 1)));  // + status word of the TMA kernel
    CU(e, cudaMalloc(&e->d_ndel, sizeof(int) * B));
    CU(e, cudaMalloc(&e->d_off_in, sizeof(int) * (B + 1)));
    CU(e, cudaMalloc(&e->d_misc, sizeof(int) * kMiscWords));
    CU(e, cudaMalloc(&e->d_class_list, nk * kNumClasses * sizeof(int)));
    CU(e, cudaMalloc(&e->d_sincos, nk * sizeof(float2)));
    int rc = alloc_det_and_desc(e);
    bind_out(e, 0);
    return rc;
}

struct BatchIn {
    const uint8_t *images;
    int in_mem, batch, width, height;
    size_t pitch, frame_stride;
    const uint8_t *mask;
    size_t mask_pitch;
};

int check_frame_args(surf_engine *e, const BatchIn &in)
{
    if (!e) return SURF_ERR_BAD_ARG;
    if (!in.images) return fail(e, SURF_ERR_BAD_ARG, "image pointer is null");
    if (in.width <= 0 || in.height <= 0 || in.batch <= 0) return fail(e, SURF_ERR_BAD_ARG, "non-positive size");
    if (in.pitch < (size_t)in.width) return fail(e, SURF_ERR_BAD_ARG, "pitch smaller than width");
    if (in.batch > 1 && in.frame_stride < in.pitch * (size_t)(in.height - 1) + in.width)
        return fail(e, SURF_ERR_BAD_ARG, "frame stride smaller than a frame");
    if (in.width > e->max_w || in.height > e->max_h || in.batch > e->max_b)
        return fail(e, SURF_ERR_UNSUPPORTED, "%dx%d x%d exceeds the engine's %dx%d x%d workspace", in.width, in.height,
                    in.batch, e->max_w, e->max_h, e->max_b);
    if (in.in_mem != SURF_MEM_HOST && in.in_mem != SURF_MEM_DEVICE) return fail(e, SURF_ERR_BAD_ARG, "bad in_mem");
    if (in.mask && in.mask_pitch < (size_t)in.width) return fail(e, SURF_ERR_BAD_ARG, "mask pitch smaller than width");
    return SURF_OK;
}

// Host frames are staged in a rotation of three device buffers: the batch in its descriptor stage reads one, the batch
// submitted after it (continuing from its look-ahead) the second, and the upload of the batch after that goes to the
// third -- the buffer whose batch ended two submits ago -- so it starts at once instead of waiting for the batch in
// flight.  Returns the next buffer of the rotation (allocated on first use).
int next_frame_buffer(surf_engine *e, int *out)
{
    const int b = (e->buf_cur + 1) % 3;
    if (!e->imgbuf[b]) CU(e, cudaMalloc(&e->imgbuf[b], e->img_pitch * e->max_h * e->max_b));
    *out = b;
    return SURF_OK;
}

// Copies B host frames (any pitch / frame stride) into a dense device stack of pitch e->img_pitch.
int upload_frames(surf_engine *e, const BatchIn &in, uint8_t *dst, cudaStream_t st)
{
    const size_t dp = e->img_pitch;
    if (in.batch == 1 || in.frame_stride == in.pitch * (size_t)in.height) {
        CU(e, cudaMemcpy2DAsync(dst, dp, in.images, in.pitch, in.width, (size_t)in.height * in.batch, cudaMemcpyHostToDevice, st));
    } else if (in.frame_stride % in.pitch == 0) {
        // a rectangle of every frame of a strided stack (e.g. the crop of an MHA sweep): one 3-D copy
        cudaMemcpy3DParms p3{};
        p3.srcPtr = make_cudaPitchedPtr((void *)in.images, in.pitch, in.width, in.frame_stride / in.pitch);
        p3.dstPtr = make_cudaPitchedPtr(dst, dp, in.width, in.height);
        p3.extent = make_cudaExtent(in.width, in.height, in.batch);
        p3.kind = cudaMemcpyHostToDevice;
        CU(e, cudaMemcpy3DAsync(&p3, st));
    } else {
        for (int b = 0; b < in.batch; b++)
            CU(e, cudaMemcpy2DAsync(dst + (size_t)b * dp * in.height, dp, in.images + (size_t)b * in.frame_stride, in.pitch,
                                    in.width, in.height, cudaMemcpyHostToDevice, st));
    }
    return SURF_OK;
}

// Exchanges the current detect set with the alternate one.
void swap_detect_sets(surf_engine *e)
{
    surf_engine::DetectSet &a = e->alt;
    std::swap(e->d_sum, a.d_sum);
    std::swap(e->d_kp_a, a.d_kp_a);
    std::swap(e->d_kp_b, a.d_kp_b);
    std::swap(e->sorted, a.sorted);
    std::swap(e->d_counts, a.d_counts);
    std::swap(e->tmap[0], a.tmap[0]);
    std::swap(e->tmap[1], a.tmap[1]);
    std::swap(e->tile_oct[0], a.tile_oct[0]);
    std::swap(e->tile_oct[1], a.tile_oct[1]);
    std::swap(e->tmap_w, a.tmap_w);
    std::swap(e->tmap_h, a.tmap_h);
    std::swap(e->tmap_b, a.tmap_b);
    std::swap(e->tmap_layers, a.tmap_layers);
    std::swap(e->tmap_ok[0], a.tmap_ok[0]);
    std::swap(e->tmap_ok[1], a.tmap_ok[1]);
    std::swap(e->ev_set_free, a.ev_free);
    std::swap(e->set_free_recorded, a.free_recorded);
}

// A look-ahead that will not be consumed: later work on the main stream must still come after it (the two share the
// determinant pyramid, the band sums and the mask integral).
int discard_lookahead(surf_engine *e)
{
    if (!e->la.valid) return SURF_OK;
    e->la.valid = false;
    CU(e, cudaStreamWaitEvent(e->stream, e->ev_la_done, 0));
    return SURF_OK;
}

// Uploads (or references) the frames; returns the device view.
int stage_images(surf_engine *e, const BatchIn &in, ImageRef *ref)
{
    if (int rc = discard_lookahead(e)) return rc;
    if (in.in_mem == SURF_MEM_DEVICE) {
        ref->ptr = in.images;
        ref->pitch = in.pitch;
        ref->frame_stride = in.frame_stride;
        return SURF_OK;
    }
    surf_engine::Prefetched &pf = e->prefetched;
    if (pf.valid && pf.images == in.images && pf.batch == in.batch && pf.width == in.width && pf.height == in.height &&
        pf.pitch == in.pitch && pf.frame_stride == in.frame_stride) {
        // these frames are already on their way (surf_prefetch_batch): take their buffer, wait on the device only
        e->buf_cur = e->buf_staged;
        CU(e, cudaStreamWaitEvent(e->stream, e->ev_upload_done, 0));
    } else {
        int b = 0;
        int rc = next_frame_buffer(e, &b);
        if (rc) return rc;
        if (e->buf_free_rec[b]) CU(e, cudaStreamWaitEvent(e->stream, e->ev_buf_free[b], 0));
        if ((rc = upload_frames(e, in, e->imgbuf[b], e->stream))) return rc;
        e->buf_cur = b;
    }
    pf.valid = false;
    e->buf_staged = -1;
    ref->ptr = e->imgbuf[e->buf_cur];
    ref->pitch = e->img_pitch;
    ref->frame_stride = e->img_pitch * in.height;
    return SURF_OK;
}

int stage_mask(surf_engine *e, const BatchIn &in, const FrameGeo &geo, const uint32_t **msum)
{
    *msum = nullptr;
    if (!in.mask) return SURF_OK;
    ImageRef m;
    if (in.in_mem == SURF_MEM_DEVICE) {
        m.ptr = in.mask;
        m.pitch = in.mask_pitch;
    } else {
        CU(e, cudaMemcpy2DAsync(e->d_mask, e->img_pitch, in.mask, in.mask_pitch, in.width, in.height,
                                cudaMemcpyHostToDevice, e->stream));
        m.ptr = e->d_mask;
        m.pitch = e->img_pitch;
    }
    m.frame_stride = 0;
    launch_integral(m, 1, geo, e->d_msum, e->d_bandsum, e->d_bandflags, 1, e->stream, &e->launches);
    *msum = e->d_msum;
    return SURF_OK;
}

// e->ev: the batch's event array.  Inside a stream capture the record is an "external" node: the event is really
// recorded when the graph runs, so the stage timings and the cross-stream waits on these events keep working.
void rec_event(surf_engine *e, cudaEvent_t ev)
{
    cudaEventRecordWithFlags(ev, e->stream, e->capturing ? cudaEventRecordExternal : cudaEventRecordDefault);
}
void rec(surf_engine *e, Ev which) { rec_event(e, e->ev[which]); }

// integral -> Hessian pyramid -> maxima -> sort.  Leaves sorted raw keypoints in e->sorted.
int run_detect(surf_engine *e, const ImageRef &img, const uint32_t *msum, const FrameGeo &geo, int B)
{
    NvtxRange nvtx("surf:stages1-3 integral+hessian+maxima+sort");
    const surf_params &p = e->params;
    launch_zero(e->d_counts, B + 1, e->stream, &e->launches);  // per-frame counters + status word
    launch_integral(img, B, geo, e->d_sum, e->d_bandsum, e->d_bandflags, 0, e->stream, &e->launches);
    rec(e, EV_INTEGRAL);
    const float thr = (float)p.hessian_threshold;
    float *det = e->d_det;
    ensure_tensor_maps(e, geo, B);
    for (int o = 0; o < p.n_octaves; o++) {
        const OctavePat pat = make_octave(o, p.n_octave_layers + 2, geo.W, geo.H);
        const size_t det_frame = (size_t)pat.lpo * pat.rows * pat.cols;
        if (o < 2 && e->tmap_ok[o]) {
            // fused: integral tile by TMA -> determinants in shared memory -> maxima; no det pyramid in HBM
            launch_hessian_nms_tma(e->tmap[o], e->tile_oct[o], msum, geo, o, thr, e->d_kp_a, e->d_counts, e->cap, B,
                                   e->stream, &e->launches);
            det += det_frame * B;
            continue;
        }
        launch_hessian(e->d_sum, geo, pat, det, det_frame, B, e->stream, &e->launches);
        launch_nms(det, det_frame, e->d_sum, msum, geo, pat, o, p.n_octave_layers, thr, e->d_kp_a, e->d_counts, e->cap,
                   B, e->stream, &e->launches);
        det += det_frame * B;
    }
    rec(e, EV_HESSIAN_NMS);
    e->sorted = launch_sort(e->d_kp_a, e->d_kp_b, e->d_counts, e->cap, B, e->stream, &e->launches);
    rec(e, EV_SORT);
    rec_event(e, e->ev_detect_done);
    e->detect_done_recorded = true;
    CU(e, cudaGetLastError());
    return SURF_OK;
}

// The output set about to be written: its previous batch's result copies, stream matcher (own stream) and match copies
// may still be reading it.  (Run before a stream capture begins: these are events of work outside the graph.)
int wait_out_set_readers(surf_engine *e)
{
    OutSet &o = e->out[e->out_bound];
    if (o.copies_in_flight) CU(e, cudaStreamWaitEvent(e->stream, o.ev_copy_done, 0));
    if (o.match_in_flight) CU(e, cudaStreamWaitEvent(e->stream, o.ev_match_done, 0));
    o.copies_in_flight = o.match_in_flight = false;
    return SURF_OK;
}

// orientation (+ descriptors) on e->sorted, then removal of deleted keypoints and packing.
int run_describe_pack(surf_engine *e, const ImageRef &img, const FrameGeo &geo, int B, bool want_desc,
                      uint8_t *patches, bool pack)
{
    NvtxRange nvtx("surf:stage4 orientation+descriptor+pack");
    const surf_params &p = e->params;
    rec(e, EV_DESC_START);
    launch_zero(e->d_ndel, B, e->stream, &e->launches);
    launch_zero(e->d_misc, kMiscWords, e->stream, &e->launches);
    launch_offsets(e->d_counts, nullptr, e->cap, B, e->d_off_in, e->stream, &e->launches);
    DescribeArgs a{};
    a.img = img;
    a.sum = e->d_sum;
    a.geo = geo;
    a.kps = e->sorted;
    a.cap = e->cap;
    a.offsets = e->d_off_in;
    a.B = B;
    a.desc = want_desc ? e->d_desc : nullptr;
    a.patches = patches;
    a.extended = p.extended;
    a.upright = p.upright;
    a.img_aligned4 = ((reinterpret_cast<uintptr_t>(img.ptr) | img.pitch | img.frame_stride) & 3) == 0;
    a.img_aligned16 = ((reinterpret_cast<uintptr_t>(img.ptr) | img.pitch | img.frame_stride) & 15) == 0;
    a.misc = e->d_misc;
    a.ndeleted = e->d_ndel;
    a.deleted_slots = e->d_deleted;
    a.class_list = e->d_class_list;
    a.list_stride = (size_t)e->max_b * e->cap;
    a.sincos = e->d_sincos;
    ensure_image_maps(e, img, geo, B);
    launch_describe(a, e->imap, e->imap_ok ? 1 : 0, e->stream, &e->launches);
    rec(e, EV_DESCRIBE);
    if (pack) {
        // the output set being written: its previous batch's result copies, stream matcher (own stream) and match
        // copies may still be reading it
        if (int rcw = wait_out_set_readers(e)) return rcw;
        launch_offsets(e->d_counts, e->d_ndel, e->cap, B, e->d_off_out, e->stream, &e->launches);
        launch_pack(e->sorted, want_desc ? e->d_desc : nullptr, p.extended ? 128 : 64, e->cap, B, e->d_off_in, e->d_ndel,
                    e->d_deleted, e->d_off_out, e->d_out_kp, want_desc ? e->d_out_desc : nullptr, e->stream,
                    &e->launches);
    }
    rec(e, EV_PACK);
    // the detect set (integral images, keypoints, counters) is free for a look-ahead once these kernels end
    rec_event(e, e->ev_set_free);
    e->set_free_recorded = true;
    CU(e, cudaGetLastError());
    return SURF_OK;
}

int fetch_counts(surf_engine *e, int B)
{
    // counts[B] | ndel[B] | off_out[B+1] | misc[4]
    int *h = e->h_pin;
    CU(e, cudaMemcpyAsync(h, e->d_counts, sizeof(int) * B, cudaMemcpyDeviceToHost, e->stream));
    CU(e, cudaMemcpyAsync(h + B, e->d_ndel, sizeof(int) * B, cudaMemcpyDeviceToHost, e->stream));
    CU(e, cudaMemcpyAsync(h + 2 * B, e->d_off_out, sizeof(int) * (B + 1), cudaMemcpyDeviceToHost, e->stream));
    CU(e, cudaMemcpyAsync(h + 3 * B + 1, e->d_misc, sizeof(int) * kMiscWords, cudaMemcpyDeviceToHost, e->stream));
    CU(e, cudaMemcpyAsync(h + 3 * B + 1 + kMiscWords, e->d_counts + B, sizeof(int), cudaMemcpyDeviceToHost, e->stream));
    // these copies read the detect set's counters: a look-ahead may overwrite the set only after them
    rec_event(e, e->ev_set_free);
    e->set_free_recorded = true;
    return SURF_OK;
}

void drop_graphs(surf_engine *e)
{
    for (surf_engine::GraphEntry &g : e->graphs)
        if (g.exec) cudaGraphExecDestroy(g.exec);
    e->graphs.clear();
}

// Stages 1-4 and the count copies of one submitted batch (everything between the upload and the result copies).  For
// small batches the host cost of ~25 launches and ~15 event records is a quarter of the call: the second time the same
// buffers, sizes and events come round the sequence is captured into a CUDA graph and replayed from then on.  The
// graph holds exactly the nodes the eager path enqueues (same kernels, same arguments), so results are byte-identical.
int run_stages(surf_engine *e, const ImageRef &img, const uint32_t *msum, const FrameGeo &geo, int B, bool want_desc)
{
    auto eager = [&]() -> int {
        int rc;
        if ((rc = run_detect(e, img, msum, geo, B))) return rc;
        if ((rc = run_describe_pack(e, img, geo, B, want_desc, nullptr, true))) return rc;
        if ((rc = fetch_counts(e, B))) return rc;
        rec(e, EV_END);
        return SURF_OK;
    };
    if (!e->use_graphs || B > e->graph_max_batch) return eager();
    surf_engine::GraphKey key;
    memset(&key, 0, sizeof(key));
    key.img = img.ptr; key.msum = msum; key.d_sum = e->d_sum; key.d_kp_a = e->d_kp_a; key.d_kp_b = e->d_kp_b;
    key.d_counts = e->d_counts; key.ev = e->ev; key.h_pin = e->h_pin; key.ev_detect_done = e->ev_detect_done;
    key.ev_set_free = e->ev_set_free; key.pitch = img.pitch; key.fstride = img.frame_stride;
    key.W = geo.W; key.H = geo.H; key.B = B; key.want_desc = want_desc; key.out_set = e->out_bound; key.cap = e->cap;
    surf_engine::GraphEntry *g = nullptr;
    for (surf_engine::GraphEntry &c : e->graphs)
        if (!memcmp(&c.key, &key, sizeof(key))) g = &c;
    if (!g) {
        if (e->graphs.size() >= 32) {  // replace the least recently used entry
            size_t lru = 0;
            for (size_t i = 1; i < e->graphs.size(); i++)
                if (e->graphs[i].last_use < e->graphs[lru].last_use) lru = i;
            if (e->graphs[lru].exec) cudaGraphExecDestroy(e->graphs[lru].exec);
            e->graphs.erase(e->graphs.begin() + lru);
        }
        e->graphs.emplace_back();
        g = &e->graphs.back();
        g->key = key;
    }
    g->last_use = ++e->graph_clock;
    g->seen++;
    if (g->failed || g->seen < 2) return eager();  // first sighting: eager (one-time allocations and attributes happen here)
    int rc;
    if ((rc = wait_out_set_readers(e))) return rc;  // events of work outside the graph: ordered before it on the stream
    if (!g->exec) {
        const int launches0 = e->launches;
        cudaGraph_t graph = nullptr;
        if (cudaStreamBeginCapture(e->stream, cudaStreamCaptureModeRelaxed) != cudaSuccess) {
            cudaGetLastError();
            g->failed = true;
            return eager();
        }
        e->capturing = true;
        rc = eager();
        e->capturing = false;
        const cudaError_t ce = cudaStreamEndCapture(e->stream, &graph);
        if (rc || ce != cudaSuccess || !graph || cudaGraphInstantiate(&g->exec, graph, 0) != cudaSuccess) {
            cudaGetLastError();
            if (graph) cudaGraphDestroy(graph);
            g->exec = nullptr;
            g->failed = true;
            e->launches = launches0;
            return eager();  // nothing of the capture ran
        }
        cudaGraphDestroy(graph);
        g->sorted = e->sorted;
        g->launches = e->launches - launches0;
        e->launches = launches0;
    }
    CU(e, cudaGraphLaunch(g->exec, e->stream));
    e->graph_replays++;
    e->sorted = g->sorted;
    e->launches += g->launches;
    e->detect_done_recorded = true;
    e->set_free_recorded = true;
    return SURF_OK;
}

int check_counts(surf_engine *e, int B)
{
    const int *h = e->h_pin;
    int worst = 0;
    for (int b = 0; b < B; b++) worst = h[b] > worst ? h[b] : worst;
    if (worst > e->cap) {
        e->required = worst;
        return fail(e, SURF_ERR_CAPACITY, "a frame produced %d raw keypoints; the engine was created for %d per frame",
                    worst, e->cap);
    }
    if (h[3 * B + 1 + kMiscWords] & 1)
        return fail(e, SURF_ERR_CUDA, "internal error: a TMA tile load of the integral image did not complete");
    if (h[3 * B + 1 + kMiscStatus] & 4)
        return fail(e, SURF_ERR_CUDA, "internal error: a TMA tile load of a frame did not complete");
    if (h[3 * B + 1 + kMiscStatus] & 2)
        return fail(e, SURF_ERR_CUDA, "internal error: descriptor tile capacity exceeded");
    if (h[3 * B + 1 + kMiscStatus] & 1)
        return fail(e, SURF_ERR_UNSUPPORTED, "a keypoint needs a descriptor window wider than %d pixels", kMaxWindow);
    return SURF_OK;
}

void collect_timings(surf_engine *e, cudaEvent_t *ev, int launches)
{
    auto ms = [&](Ev a, Ev b) {
        float t = 0;
        if (cudaEventElapsedTime(&t, ev[a], ev[b]) != cudaSuccess) {
            cudaGetLastError();
            t = 0;
        }
        return t;
    };
    surf_timings &t = e->timings;
    t.h2d = ms(EV_START, EV_H2D);
    t.integral = ms(EV_H2D, EV_INTEGRAL);
    t.hessian = ms(EV_INTEGRAL, EV_HESSIAN_NMS);  // Hessian + maxima of all octaves, interleaved
    t.nms = 0;
    t.sort = ms(EV_HESSIAN_NMS, EV_SORT);
    t.describe = ms(EV_DESC_START, EV_DESCRIBE);
    t.pack = ms(EV_DESCRIBE, EV_PACK);
    t.d2h = ms(EV_PACK, EV_END);
    t.total = ms(EV_START, EV_END);
    t.launches = launches;
}

}  // namespace

extern "C" {

const char *surf_version(void) { return "surf_b200 0.1 (sm_100a)"; }

int surf_create(const surf_params *params, int device, int max_width, int max_height, int max_batch, int max_keypoints,
                surf_engine **out)
{
    if (!out) return SURF_ERR_BAD_ARG;
    *out = nullptr;
    if (max_width <= 0 || max_height <= 0 || max_batch <= 0 || max_keypoints <= 0) return SURF_ERR_BAD_ARG;
    if (max_width > integral_max_width()) return SURF_ERR_UNSUPPORTED;  // a band of the integral kernel holds whole rows
    int rc = check_params(nullptr, params);
    if (rc) return rc;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0 || device < 0 || device >= ndev) {
        cudaGetLastError();
        return SURF_ERR_NO_DEVICE;
    }
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return SURF_ERR_NO_DEVICE;
    if (prop.major != 10) return SURF_ERR_NO_DEVICE;  // kernels are built for sm_100a only
    surf_engine *e = new (std::nothrow) surf_engine();
    if (!e) return SURF_ERR_BAD_ARG;
    e->params = *params;
    e->device = device;
    e->max_w = max_width;
    e->max_h = max_height;
    e->max_b = max_batch;
    e->cap = (max_keypoints + 127) & ~127;
    if (getenv("SURF_B200_NO_TMA")) e->use_tma = 0;
    if (getenv("SURF_B200_NO_GRAPHS")) e->use_graphs = 0;
    if (const char *v = getenv("SURF_B200_EARLY_UPLOAD")) e->early_upload = atoi(v);
    if (cudaSetDevice(device) != cudaSuccess || cudaStreamCreateWithFlags(&e->stream, cudaStreamNonBlocking) != cudaSuccess) {
        delete e;
        return SURF_ERR_CUDA;
    }
    for (auto &arr : e->evpool)
        for (int i = 0; i < EV_COUNT; i++) cudaEventCreate(&arr[i]);
    e->ev = e->evpool[0];
    e->ev_valid = true;
    cudaStreamCreateWithFlags(&e->copy_stream, cudaStreamNonBlocking);
    cudaStreamCreateWithFlags(&e->match_stream, cudaStreamNonBlocking);
    cudaEventCreate(&e->ev_match_t0);
    cudaEventCreate(&e->ev_match_t1);
    for (OutSet &o : e->out) {
        cudaEventCreateWithFlags(&o.ev_ready, cudaEventDisableTiming);
        cudaEventCreateWithFlags(&o.ev_copy_done, cudaEventDisableTiming);
        cudaEventCreateWithFlags(&o.ev_match_done, cudaEventDisableTiming);
    }
    cudaEventCreateWithFlags(&e->ev_match_copy_done, cudaEventDisableTiming);
    cudaEventCreateWithFlags(&e->ev_upload_done, cudaEventDisableTiming);
    for (auto &ev : e->ev_buf_free) cudaEventCreateWithFlags(&ev, cudaEventDisableTiming);
    cudaStreamCreateWithFlags(&e->upload_stream, cudaStreamNonBlocking);
    // Default priority on purpose.  Measured (64 frames, 640x480): with a high-priority look-ahead stream and
    // descriptor CTAs that turn over every 8-128 keypoints the Hessian kernels run at full speed but stretch the
    // descriptor stage by more than they save (8.24k frames/s); queued behind the persistent descriptor CTAs they
    // fill the slots those leave free (8.60k frames/s against 8.24k without look-ahead).
    cudaStreamCreateWithFlags(&e->la_stream, cudaStreamNonBlocking);
    cudaEventCreateWithFlags(&e->ev_detect_done, cudaEventDisableTiming);
    cudaEventCreateWithFlags(&e->ev_la_done, cudaEventDisableTiming);
    cudaEventCreateWithFlags(&e->ev_set_free, cudaEventDisableTiming);
    cudaEventCreateWithFlags(&e->alt.ev_free, cudaEventDisableTiming);
    upload_tables();
    rc = alloc_workspace(e);
    if (rc == SURF_OK && cudaGetLastError() != cudaSuccess) rc = SURF_ERR_CUDA;
    if (rc) {
        fprintf(stderr, "surf_create: %s\n", e->err);
        surf_destroy(e);
        return rc;
    }
    *out = e;
    return SURF_OK;
}

void surf_destroy(surf_engine *e)
{
    if (!e) return;
    cudaSetDevice(e->device);
    // every stream that may still touch the workspace drains before it is freed
    for (cudaStream_t st : {e->la_stream, e->upload_stream, e->match_stream, e->copy_stream, e->stream})
        if (st) cudaStreamSynchronize(st);
    if (e->solo) {
        surf_comm_destroy(e->solo);
        e->solo = nullptr;
    }
    free_workspace(e);
    if (e->ev_valid)
        for (auto &arr : e->evpool)
            for (int i = 0; i < EV_COUNT; i++) cudaEventDestroy(arr[i]);
    if (e->copy_stream) {
        cudaStreamSynchronize(e->copy_stream);
        cudaStreamDestroy(e->copy_stream);
    }
    if (e->match_stream) {
        cudaStreamSynchronize(e->match_stream);
        cudaStreamDestroy(e->match_stream);
    }
    if (e->ev_match_t0) cudaEventDestroy(e->ev_match_t0);
    if (e->ev_match_t1) cudaEventDestroy(e->ev_match_t1);
    for (cudaEvent_t m : e->marks)
        if (m) cudaEventDestroy(m);
    for (OutSet &o : e->out)
        for (cudaEvent_t ev : {o.ev_ready, o.ev_copy_done, o.ev_match_done})
            if (ev) cudaEventDestroy(ev);
    if (e->ev_match_copy_done) cudaEventDestroy(e->ev_match_copy_done);
    if (e->ev_upload_done) cudaEventDestroy(e->ev_upload_done);
    for (cudaEvent_t ev : e->ev_buf_free)
        if (ev) cudaEventDestroy(ev);
    if (e->upload_stream) {
        cudaStreamSynchronize(e->upload_stream);
        cudaStreamDestroy(e->upload_stream);
    }
    if (e->la_stream) {
        cudaStreamSynchronize(e->la_stream);
        cudaStreamDestroy(e->la_stream);
    }
    if (e->ev_detect_done) cudaEventDestroy(e->ev_detect_done);
    if (e->ev_la_done) cudaEventDestroy(e->ev_la_done);
    if (e->ev_set_free) cudaEventDestroy(e->ev_set_free);
    if (e->alt.ev_free) cudaEventDestroy(e->alt.ev_free);
    if (e->stream) cudaStreamDestroy(e->stream);
    delete e;
}

int surf_set_params(surf_engine *e, const surf_params *params)
{
    if (!e) return SURF_ERR_BAD_ARG;
    int rc = check_params(e, params);
    if (rc) return rc;
    if ((rc = ensure_device(e))) return rc;
    // every stream that may still touch buffers about to be re-allocated drains first
    for (cudaStream_t st : {e->la_stream, e->upload_stream, e->match_stream, e->copy_stream, e->stream})
        CU(e, cudaStreamSynchronize(st));
    e->la.valid = false;
    for (OutSet &o : e->out) o.match_in_flight = o.copies_in_flight = false;
    e->match_copies_in_flight = false;
    // pending batches and the stream matcher's predecessor frame were produced with the old parameters
    // (descriptor length, thresholds): neither is carried over
    drop_batches(e);
    drop_graphs(e);  // thresholds, layer tables and descriptor length are baked into the captured launches
    e->stream_primed = false;
    e->params = *params;
    e->tmap_w = 0;  // layer tables depend on the parameters
    e->alt.tmap_w = 0;
    return alloc_det_and_desc(e);
}

int surf_get_params(const surf_engine *e, surf_params *params)
{
    if (!e || !params) return SURF_ERR_BAD_ARG;
    *params = e->params;
    return SURF_OK;
}

const char *surf_last_error(const surf_engine *e) { return e ? e->err : "null engine"; }
long surf_required_capacity(const surf_engine *e) { return e ? e->required : 0; }
int surf_descriptor_size(const surf_engine *e) { return e ? (e->params.extended ? 128 : 64) : 0; }
void *surf_stream(const surf_engine *e) { return e ? (void *)e->stream : nullptr; }

int surf_get_timings(const surf_engine *e, surf_timings *t)
{
    if (!e || !t) return SURF_ERR_BAD_ARG;
    *t = e->timings;
    return SURF_OK;
}

int surf_sync(surf_engine *e)
{
    if (!e) return SURF_ERR_BAD_ARG;
    int rc = ensure_device(e);
    if (rc) return rc;
    CU(e, cudaStreamSynchronize(e->stream));
    CU(e, cudaStreamSynchronize(e->match_stream));
    for (OutSet &o : e->out) o.match_in_flight = false;
    return SURF_OK;
}

int surf_submit_batch(surf_engine *e, const uint8_t *images, int in_mem, int batch, int width, int height, size_t pitch,
                      size_t frame_stride, const uint8_t *mask, size_t mask_pitch, int want_descriptors)
{
    NvtxRange nvtx("surf_submit_batch");
    BatchIn in{images, in_mem, batch, width, height, pitch, frame_stride, mask, mask_pitch};
    int rc = check_frame_args(e, in);
    if (rc) return rc;
    if ((rc = ensure_device(e))) return rc;
    // two batches may be in flight: this one takes the output set the batch before the previous one used
    const int s = e->out_next;
    OutSet &o = e->out[s];
    if (o.active)
        return fail(e, SURF_ERR_BAD_ARG, "two batches are already submitted: collect one before submitting a third");
    if (e->out_collected == s) e->out_collected = -1;   // its buffers are about to be overwritten
    o.collected = false;
    bind_out(e, s);
    e->launches = 0;
    const FrameGeo geo = make_geo(width, height);
    ImageRef img;
    bool adopted = false;
    const surf_engine::LookAhead &la = e->la;
    if (la.valid && la.images == images && la.in_mem == in_mem && la.batch == batch && la.width == width && la.height == height &&
        la.pitch == pitch && la.frame_stride == frame_stride && la.mask == mask && la.mask_pitch == mask_pitch) {
        // stages 1-3 of these frames already ran (or are running) on the look-ahead stream: adopt their detect set,
        // their stage events and, for host frames, the frame buffer they were uploaded into
        adopted = true;
        e->la.valid = false;
        swap_detect_sets(e);
        e->ev = e->la_ev;
        if (in_mem == SURF_MEM_HOST) {
            e->buf_cur = e->buf_staged;
            e->buf_staged = -1;
        }
        img = la.img;
        e->launches = la.launches;
        CU(e, cudaStreamWaitEvent(e->stream, e->ev_la_done, 0));
    } else {
        e->ev = next_ev_array(e);
        rec(e, EV_START);
        if ((rc = stage_images(e, in, &img))) return rc;
        const uint32_t *msum = nullptr;
        if ((rc = stage_mask(e, in, geo, &msum))) return rc;
        rec(e, EV_H2D);
        if ((rc = run_stages(e, img, msum, geo, batch, want_descriptors != 0))) return rc;
    }
    if (adopted) {
        if ((rc = run_describe_pack(e, img, geo, batch, want_descriptors != 0, nullptr, true))) return rc;
        if ((rc = fetch_counts(e, batch))) return rc;
        rec(e, EV_END);
    }
    CU(e, cudaEventRecord(o.ev_ready, e->stream));
    if (in_mem == SURF_MEM_HOST) {  // the staged frames are free again after these kernels
        CU(e, cudaEventRecord(e->ev_buf_free[e->buf_cur], e->stream));
        e->buf_free_rec[e->buf_cur] = true;
    }
    o.ev = e->ev;
    o.active = true;
    o.matched = false;
    o.batch = batch;
    o.want_desc = want_descriptors != 0;
    o.launches = e->launches;
    e->out_next = s ^ 1;
    return SURF_OK;
}

int surf_prefetch_batch(surf_engine *e, const uint8_t *images, int batch, int width, int height, size_t pitch,
                        size_t frame_stride)
{
    BatchIn in{images, SURF_MEM_HOST, batch, width, height, pitch, frame_stride, nullptr, 0};
    int rc = check_frame_args(e, in);
    if (rc) return rc;
    if ((rc = ensure_device(e))) return rc;
    if (e->la.valid) {  // a pending look-ahead owns the staged frame buffer: it is given up
        e->la.valid = false;
        CU(e, cudaStreamWaitEvent(e->upload_stream, e->ev_la_done, 0));
        CU(e, cudaStreamWaitEvent(e->stream, e->ev_la_done, 0));
    }
    int b = 0;
    if ((rc = next_frame_buffer(e, &b))) return rc;
    // the upload may start as soon as the last kernels that read this buffer (three batches ago) have ended; it
    // overlaps the kernels of the batches in flight, which read the other buffers.  The host does not wait.
    if (e->buf_free_rec[b]) CU(e, cudaStreamWaitEvent(e->upload_stream, e->ev_buf_free[b], 0));
    if ((rc = upload_frames(e, in, e->imgbuf[b], e->upload_stream))) return rc;
    CU(e, cudaEventRecord(e->ev_upload_done, e->upload_stream));
    e->buf_staged = b;
    surf_engine::Prefetched &pf = e->prefetched;
    pf.images = images; pf.batch = batch; pf.width = width; pf.height = height; pf.pitch = pitch;
    pf.frame_stride = frame_stride; pf.valid = true;
    return SURF_OK;
}

int surf_lookahead_batch(surf_engine *e, const uint8_t *images, int in_mem, int batch, int width, int height, size_t pitch,
                         size_t frame_stride, const uint8_t *mask, size_t mask_pitch)
{
    NvtxRange nvtx("surf_lookahead_batch");
    BatchIn in{images, in_mem, batch, width, height, pitch, frame_stride, mask, mask_pitch};
    int rc = check_frame_args(e, in);
    if (rc) return rc;
    if ((rc = ensure_device(e))) return rc;
    if (e->la.valid) {  // replaces an unused look-ahead: same stream, so it simply queues behind it
        e->la.valid = false;
    }
    e->prefetched.valid = false;  // the second frame buffer is about to be overwritten
    surf_engine::DetectSet &a = e->alt;
    if (!a.d_sum) {
        const FrameGeo g = make_geo(e->max_w, e->max_h);
        const size_t nk = (size_t)e->max_b * e->cap;
        CU(e, cudaMalloc(&a.d_sum, g.sum_frame * sizeof(uint32_t) * e->max_b));
        CU(e, cudaMalloc(&a.d_kp_a, nk * sizeof(surf_keypoint)));
        CU(e, cudaMalloc(&a.d_kp_b, nk * sizeof(surf_keypoint)));
        CU(e, cudaMalloc(&a.d_counts, sizeof(int) * (e->max_b + 1)));
    }
    int fb = 0;
    cudaStream_t ls = e->la_stream;
    if (in_mem == SURF_MEM_HOST) {
        if ((rc = next_frame_buffer(e, &fb))) return rc;
        if (e->early_upload) {
            // the frames go up on the upload stream right away (third buffer of the rotation: its batch ended two
            // submits ago); only stages 1-3 wait for the detect set
            if (e->la_done_recorded) CU(e, cudaStreamWaitEvent(e->upload_stream, e->ev_la_done, 0));  // a replaced look-ahead
            if (e->buf_free_rec[fb]) CU(e, cudaStreamWaitEvent(e->upload_stream, e->ev_buf_free[fb], 0));
            if ((rc = upload_frames(e, in, e->imgbuf[fb], e->upload_stream))) return rc;
            CU(e, cudaEventRecord(e->ev_upload_done, e->upload_stream));
        }
        e->buf_staged = fb;
    }
    // Ordering: after stages 1-3 of the batch in flight (the determinant pyramid, band sums and mask integral are
    // shared scratch; it is also the overlap wanted: these kernels run beside that batch's descriptor stage), after
    // the stage-4 kernels that last read the alternate set, and after the last readers of the frame buffer.
    if (e->detect_done_recorded) CU(e, cudaStreamWaitEvent(ls, e->ev_detect_done, 0));
    if (a.free_recorded) CU(e, cudaStreamWaitEvent(ls, a.ev_free, 0));
    if (in_mem == SURF_MEM_HOST) {
        if (e->early_upload) {
            CU(e, cudaStreamWaitEvent(ls, e->ev_upload_done, 0));
        } else {
            // the upload leads the look-ahead's own stream: it starts when the batch two submits back has left its
            // descriptor stage, together with that batch's result copies in the other direction
            if (e->buf_free_rec[fb]) CU(e, cudaStreamWaitEvent(ls, e->ev_buf_free[fb], 0));
            if ((rc = upload_frames(e, in, e->imgbuf[fb], ls))) return rc;
        }
    }
    // run stages 1-3 with the alternate set, the look-ahead stream and the look-ahead's own stage events in place
    const int launches0 = e->launches;
    e->launches = 0;
    swap_detect_sets(e);
    std::swap(e->stream, e->la_stream);
    cudaEvent_t *const main_ev = e->ev;
    e->la_ev = next_ev_array(e);   // (la.valid is false here: a replaced look-ahead's array may be taken again)
    e->ev = e->la_ev;
    const FrameGeo geo = make_geo(width, height);
    ImageRef img;
    rec(e, EV_START);
    if (in_mem == SURF_MEM_DEVICE) {
        img.ptr = images;
        img.pitch = pitch;
        img.frame_stride = frame_stride;
    } else {
        img.ptr = e->imgbuf[fb];
        img.pitch = e->img_pitch;
        img.frame_stride = e->img_pitch * height;
    }
    const uint32_t *msum = nullptr;
    if (!rc) rc = stage_mask(e, in, geo, &msum);
    if (!rc) {
        rec(e, EV_H2D);
        rc = run_detect(e, img, msum, geo, batch);
    }
    // run_detect recorded ev_detect_done on the look-ahead stream: that is this look-ahead's completion; the main
    // stream's own marker must be recorded again by its next run_detect
    std::swap(e->ev_detect_done, e->ev_la_done);
    e->detect_done_recorded = false;
    e->la_done_recorded = true;
    e->ev = main_ev;
    std::swap(e->stream, e->la_stream);
    swap_detect_sets(e);
    const int la_launches = e->launches;
    e->launches = launches0;
    if (rc) return rc;
    surf_engine::LookAhead &la = e->la;
    la.images = images; la.mask = mask; la.in_mem = in_mem; la.batch = batch; la.width = width; la.height = height;
    la.pitch = pitch; la.frame_stride = frame_stride; la.mask_pitch = mask_pitch; la.img = img; la.launches = la_launches;
    la.valid = true;
    return SURF_OK;
}

int surf_collect_batch(surf_engine *e, surf_keypoint *keypoints, float *descriptors, int out_mem, long capacity,
                       int32_t *offsets)
{
    if (!e) return SURF_ERR_BAD_ARG;
    NvtxRange nvtx("surf_collect_batch");
    // the OLDEST submitted batch: with two in flight that is the one the next submit would overwrite
    int s = e->out_next;
    if (!e->out[s].active) s ^= 1;
    // nothing in flight: the batch collected last may be collected again (counts-only call first, then a second call
    // into arrays sized from the counts) until the next submit
    if (!e->out[s].active && e->out_collected >= 0 && e->out[e->out_collected].collected) s = e->out_collected;
    OutSet &o = e->out[s];
    if (!o.active && !o.collected) return fail(e, SURF_ERR_BAD_ARG, "no submitted batch to collect");
    if (!offsets) return fail(e, SURF_ERR_BAD_ARG, "offsets is null");
    int rc = ensure_device(e);
    if (rc) return rc;
    const int B = o.batch;
    CU(e, cudaEventSynchronize(o.ev_ready));   // this batch only: a batch submitted after it keeps the GPU busy meanwhile
    const int bound = e->out_bound;
    bind_out(e, s);
    rc = check_counts(e, B);
    const int *off = e->h_pin + 2 * B;
    const long total = off[B];
    if (!rc) {
        for (int b = 0; b <= B; b++) offsets[b] = off[b];
        if (total > capacity) {
            e->required = total;
            rc = fail(e, SURF_ERR_CAPACITY, "%ld keypoints in the batch; the output arrays hold %ld", total, capacity);
        }
    }
    if (rc) {   // the batch stays collectable (e.g. with larger arrays after SURF_ERR_CAPACITY)
        bind_out(e, bound);
        return rc;
    }
    collect_timings(e, o.ev, o.launches);
    const cudaMemcpyKind kind = out_mem == SURF_MEM_DEVICE ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost;
    const int D = e->params.extended ? 128 : 64;
    // The result copies go to the copy stream (the batch's kernels have ended: ev_ready) so that they neither wait
    // for nor delay a batch submitted after this one.  Asynchronous-output mode returns without waiting for them:
    // they overlap the following kernels and surf_wait_outputs() completes them.
    const bool any = total > 0 && (keypoints || (descriptors && o.want_desc));
    if (any) {
        if (keypoints) CU(e, cudaMemcpyAsync(keypoints, o.d_kp, sizeof(surf_keypoint) * total, kind, e->copy_stream));
        if (descriptors && o.want_desc)
            CU(e, cudaMemcpyAsync(descriptors, o.d_desc, sizeof(float) * D * total, kind, e->copy_stream));
        CU(e, cudaEventRecord(o.ev_copy_done, e->copy_stream));
        o.copies_in_flight = true;
        if (!e->async_outputs) {
            CU(e, cudaStreamSynchronize(e->copy_stream));
            o.copies_in_flight = false;
        }
    }
    o.active = false;
    o.collected = true;
    e->out_collected = s;
    bind_out(e, bound);
    return SURF_OK;
}

int surf_set_async_outputs(surf_engine *e, int enable)
{
    if (!e) return SURF_ERR_BAD_ARG;
    int rc = ensure_device(e);
    if (rc) return rc;
    CU(e, cudaStreamSynchronize(e->copy_stream));
    for (OutSet &o : e->out) o.copies_in_flight = false;
    e->match_copies_in_flight = false;
    e->async_outputs = enable ? 1 : 0;
    return SURF_OK;
}

int surf_wait_outputs(surf_engine *e)
{
    if (!e) return SURF_ERR_BAD_ARG;
    int rc = ensure_device(e);
    if (rc) return rc;
    CU(e, cudaStreamSynchronize(e->copy_stream));
    for (OutSet &o : e->out) o.copies_in_flight = false;
    e->match_copies_in_flight = false;
    return SURF_OK;
}

int surf_device_results(surf_engine *e, const surf_keypoint **keypoints, const float **descriptors,
                        const int32_t **offsets)
{
    if (!e) return SURF_ERR_BAD_ARG;
    // the batch collected last; before any collect, the batch submitted last
    int s = e->out_collected;
    if (s < 0) {
        s = e->out_next ^ 1;
        if (!e->out[s].active) return fail(e, SURF_ERR_BAD_ARG, "no submitted batch");
    }
    const OutSet &o = e->out[s];
    if (keypoints) *keypoints = o.d_kp;
    if (descriptors) *descriptors = o.want_desc ? o.d_desc : nullptr;
    if (offsets) *offsets = o.d_off;
    return SURF_OK;
}

int surf_detect_and_compute_batch(surf_engine *e, const uint8_t *images, int in_mem, int batch, int width, int height,
                                  size_t pitch, size_t frame_stride, const uint8_t *mask, size_t mask_pitch,
                                  surf_keypoint *keypoints, float *descriptors, int out_mem, long capacity,
                                  int32_t *offsets)
{
    int rc = surf_submit_batch(e, images, in_mem, batch, width, height, pitch, frame_stride, mask, mask_pitch,
                               descriptors != nullptr);
    if (rc) return rc;
    return surf_collect_batch(e, keypoints, descriptors, out_mem, capacity, offsets);
}

int surf_detect_and_compute(surf_engine *e, const uint8_t *image, int width, int height, size_t pitch,
                            const uint8_t *mask, size_t mask_pitch, surf_keypoint *keypoints, float *descriptors,
                            int capacity, int *n_out)
{
    if (!e) return SURF_ERR_BAD_ARG;
    if (!n_out) return fail(e, SURF_ERR_BAD_ARG, "n_out is null");
    *n_out = 0;
    // cv::FeatureDetector::detect: an empty image yields no keypoints, not an error
    if (!image || width == 0 || height == 0) return image ? SURF_OK : fail(e, SURF_ERR_BAD_ARG, "image pointer is null");
    int32_t off[2] = {0, 0};
    int rc = surf_detect_and_compute_batch(e, image, SURF_MEM_HOST, 1, width, height, pitch, pitch * (size_t)height, mask,
                                           mask_pitch, keypoints, descriptors, SURF_MEM_HOST, capacity, off);
    if (rc) return rc;
    *n_out = off[1];
    return SURF_OK;
}

int surf_detect(surf_engine *e, const uint8_t *image, int width, int height, size_t pitch, const uint8_t *mask,
                size_t mask_pitch, surf_keypoint *keypoints, int capacity, int *n_out)
{
    return surf_detect_and_compute(e, image, width, height, pitch, mask, mask_pitch, keypoints, nullptr, capacity, n_out);
}

static int describe_given(surf_engine *e, const uint8_t *image, int width, int height, size_t pitch,
                          surf_keypoint *keypoints, int n, float *descriptors, uint8_t *patches, bool pack, int *n_out)
{
    BatchIn in{image, SURF_MEM_HOST, 1, width, height, pitch, pitch * (size_t)height, nullptr, 0};
    int rc = check_frame_args(e, in);
    if (rc) return rc;
    if (!keypoints) return fail(e, SURF_ERR_BAD_ARG, "keypoints is null");
    if (n > e->cap) {
        e->required = n;
        return fail(e, SURF_ERR_CAPACITY, "%d keypoints given; the engine was created for %d per frame", n, e->cap);
    }
    if ((rc = ensure_device(e))) return rc;
    drop_batches(e);
    bind_out(e, e->out_next);
    e->ev = next_ev_array(e);
    e->launches = 0;
    const FrameGeo geo = make_geo(width, height);
    const int D = e->params.extended ? 128 : 64;
    rec(e, EV_START);
    ImageRef img;
    if ((rc = stage_images(e, in, &img))) return rc;
    CU(e, cudaMemcpyAsync(e->d_kp_a, keypoints, sizeof(surf_keypoint) * n, cudaMemcpyHostToDevice, e->stream));
    CU(e, cudaMemcpyAsync(e->d_counts, &n, sizeof(int), cudaMemcpyHostToDevice, e->stream));
    e->sorted = e->d_kp_a;
    rec(e, EV_H2D);
    launch_integral(img, 1, geo, e->d_sum, e->d_bandsum, e->d_bandflags, 0, e->stream, &e->launches);
    rec(e, EV_INTEGRAL);
    rec(e, EV_HESSIAN_NMS);
    rec(e, EV_SORT);
    if (patches && !e->d_patches) CU(e, cudaMalloc(&e->d_patches, (size_t)e->cap * kPatchPx));
    if ((rc = run_describe_pack(e, img, geo, 1, descriptors != nullptr, patches ? e->d_patches : nullptr, pack))) return rc;
    if ((rc = fetch_counts(e, 1))) return rc;
    CU(e, cudaStreamSynchronize(e->stream));
    if ((rc = check_counts(e, 1))) return rc;
    if (pack) {
        const int m = e->h_pin[2 + 1];
        if (m > 0) {
            CU(e, cudaMemcpyAsync(keypoints, e->d_out_kp, sizeof(surf_keypoint) * m, cudaMemcpyDeviceToHost, e->stream));
            if (descriptors)
                CU(e, cudaMemcpyAsync(descriptors, e->d_out_desc, sizeof(float) * D * m, cudaMemcpyDeviceToHost, e->stream));
        }
        if (n_out) *n_out = m;
    } else {
        CU(e, cudaMemcpyAsync(keypoints, e->d_kp_a, sizeof(surf_keypoint) * n, cudaMemcpyDeviceToHost, e->stream));
        if (descriptors)
            CU(e, cudaMemcpyAsync(descriptors, e->d_desc, sizeof(float) * D * n, cudaMemcpyDeviceToHost, e->stream));
        if (patches)
            CU(e, cudaMemcpyAsync(patches, e->d_patches, (size_t)n * kPatchPx, cudaMemcpyDeviceToHost, e->stream));
        if (n_out) *n_out = n;
    }
    rec(e, EV_END);
    CU(e, cudaStreamSynchronize(e->stream));
    collect_timings(e, e->ev, e->launches);
    return SURF_OK;
}

int surf_compute(surf_engine *e, const uint8_t *image, int width, int height, size_t pitch, surf_keypoint *keypoints,
                 int n_in, float *descriptors, int *n_out)
{
    if (!e) return SURF_ERR_BAD_ARG;
    if (!n_out) return fail(e, SURF_ERR_BAD_ARG, "n_out is null");
    *n_out = 0;
    // DescriptorExtractor::compute: empty image or no keypoints -> empty result
    if (!image || width == 0 || height == 0 || n_in <= 0) return SURF_OK;
    if (!keypoints || !descriptors) return fail(e, SURF_ERR_BAD_ARG, "null keypoints or descriptors");
    // KeyPointsFilter::runByKeypointSize(keypoints, FLT_EPSILON): stable removal on the host
    int m = 0;
    for (int i = 0; i < n_in; i++)
        if (!(keypoints[i].size < FLT_EPSILON)) keypoints[m++] = keypoints[i];
    if (m == 0) return SURF_OK;
    return describe_given(e, image, width, height, pitch, keypoints, m, descriptors, nullptr, true, n_out);
}

int surf_describe_keypoints(surf_engine *e, const uint8_t *image, int width, int height, size_t pitch,
                            surf_keypoint *keypoints, int n, float *descriptors, uint8_t *patches)
{
    if (!e) return SURF_ERR_BAD_ARG;
    if (n <= 0) return SURF_OK;
    return describe_given(e, image, width, height, pitch, keypoints, n, descriptors, patches, false, nullptr);
}

int surf_integral(surf_engine *e, const uint8_t *image, int width, int height, size_t pitch, int clamp01, int32_t *sum)
{
    BatchIn in{image, SURF_MEM_HOST, 1, width, height, pitch, pitch * (size_t)height, nullptr, 0};
    int rc = check_frame_args(e, in);
    if (rc) return rc;
    if (!sum) return fail(e, SURF_ERR_BAD_ARG, "sum is null");
    if ((rc = ensure_device(e))) return rc;
    drop_batches(e);
    bind_out(e, e->out_next);
    e->ev = next_ev_array(e);
    e->launches = 0;
    const FrameGeo geo = make_geo(width, height);
    ImageRef img;
    if ((rc = stage_images(e, in, &img))) return rc;
    launch_integral(img, 1, geo, e->d_sum, e->d_bandsum, e->d_bandflags, clamp01, e->stream, &e->launches);
    CU(e, cudaGetLastError());
    CU(e, cudaMemcpy2DAsync(sum, sizeof(int32_t) * (width + 1), e->d_sum, sizeof(uint32_t) * geo.SP,
                            sizeof(int32_t) * (width + 1), height + 1, cudaMemcpyDeviceToHost, e->stream));
    CU(e, cudaStreamSynchronize(e->stream));
    return SURF_OK;
}

int surf_hessian_layer(surf_engine *e, const uint8_t *image, int width, int height, size_t pitch, int octave, int layer,
                       float *det, int *rows, int *cols)
{
    BatchIn in{image, SURF_MEM_HOST, 1, width, height, pitch, pitch * (size_t)height, nullptr, 0};
    int rc = check_frame_args(e, in);
    if (rc) return rc;
    const int lpo = e->params.n_octave_layers + 2;
    if (!det || octave < 0 || octave >= e->params.n_octaves || layer < 0 || layer >= lpo)
        return fail(e, SURF_ERR_BAD_ARG, "bad octave/layer or null output");
    if ((rc = ensure_device(e))) return rc;
    drop_batches(e);
    bind_out(e, e->out_next);
    e->ev = next_ev_array(e);
    e->launches = 0;
    const FrameGeo geo = make_geo(width, height);
    ImageRef img;
    if ((rc = stage_images(e, in, &img))) return rc;
    launch_integral(img, 1, geo, e->d_sum, e->d_bandsum, e->d_bandflags, 0, e->stream, &e->launches);
    const OctavePat pat = make_octave(octave, lpo, width, height);
    const size_t det_frame = (size_t)pat.lpo * pat.rows * pat.cols;
    launch_hessian(e->d_sum, geo, pat, e->d_det, det_frame, 1, e->stream, &e->launches);
    CU(e, cudaGetLastError());
    const size_t plane = (size_t)pat.rows * pat.cols;
    if (plane)
        CU(e, cudaMemcpyAsync(det, e->d_det + plane * layer, plane * sizeof(float), cudaMemcpyDeviceToHost, e->stream));
    CU(e, cudaStreamSynchronize(e->stream));
    if (rows) *rows = pat.rows;
    if (cols) *cols = pat.cols;
    return SURF_OK;
}

int surf_raw_keypoints(surf_engine *e, const uint8_t *image, int width, int height, size_t pitch, const uint8_t *mask,
                       size_t mask_pitch, surf_keypoint *keypoints, int capacity, int *n_out)
{
    BatchIn in{image, SURF_MEM_HOST, 1, width, height, pitch, pitch * (size_t)height, mask, mask_pitch};
    int rc = check_frame_args(e, in);
    if (rc) return rc;
    if (!keypoints || !n_out) return fail(e, SURF_ERR_BAD_ARG, "null output");
    if ((rc = ensure_device(e))) return rc;
    drop_batches(e);
    bind_out(e, e->out_next);
    e->ev = next_ev_array(e);
    e->launches = 0;
    const FrameGeo geo = make_geo(width, height);
    rec(e, EV_START);
    ImageRef img;
    if ((rc = stage_images(e, in, &img))) return rc;
    const uint32_t *msum = nullptr;
    if ((rc = stage_mask(e, in, geo, &msum))) return rc;
    rec(e, EV_H2D);
    if ((rc = run_detect(e, img, msum, geo, 1))) return rc;
    CU(e, cudaMemcpyAsync(e->h_pin, e->d_counts, 2 * sizeof(int), cudaMemcpyDeviceToHost, e->stream));
    CU(e, cudaStreamSynchronize(e->stream));
    if (e->h_pin[1] & 1)
        return fail(e, SURF_ERR_CUDA, "internal error: a TMA tile load of the integral image did not complete");
    const int n = e->h_pin[0];
    *n_out = n;
    if (n > e->cap || n > capacity) {
        e->required = n;
        return fail(e, SURF_ERR_CAPACITY, "%d raw keypoints; capacity %d (engine %d)", n, capacity, e->cap);
    }
    if (n > 0) CU(e, cudaMemcpy(keypoints, e->sorted, sizeof(surf_keypoint) * n, cudaMemcpyDeviceToHost));
    return SURF_OK;
}

// ---- stage 5 ----

static int match_common(surf_engine *e, const float *query, int n_query, const float *train, int n_train, int dim, int k,
                        float ratio, surf_dmatch *out, uint8_t *good, int mem)
{
    if (!e) return SURF_ERR_BAD_ARG;
    if (!query || !train || !out) return fail(e, SURF_ERR_BAD_ARG, "null pointer");
    if (n_query < 0 || n_train < 0) return fail(e, SURF_ERR_BAD_ARG, "negative row count");
    if (dim != 64 && dim != 128) return fail(e, SURF_ERR_UNSUPPORTED, "descriptor length must be 64 or 128");
    if (k < 1 || k > 4) return fail(e, SURF_ERR_UNSUPPORTED, "k must be in 1..4");
    if (n_query == 0) return SURF_OK;
    int rc = ensure_device(e);
    if (rc) return rc;
    e->launches = 0;
    const float *dq = query, *dt = train;
    surf_dmatch *dout = out;
    uint8_t *dgood = good;
    if (mem == SURF_MEM_HOST) {
        const size_t bq = sizeof(float) * dim * (size_t)n_query, bt = sizeof(float) * dim * (size_t)n_train;
        const size_t bo = sizeof(surf_dmatch) * (size_t)k * n_query;
        const size_t need = ((bq + 255) & ~(size_t)255) + ((bt + 255) & ~(size_t)255) + ((bo + 255) & ~(size_t)255) + n_query;
        if (need > e->match_io_cap) {
            cudaFree(e->d_match_io);
            e->d_match_io = nullptr;
            e->match_io_cap = 0;
            CU(e, cudaMalloc(&e->d_match_io, need));
            e->match_io_cap = need;
        }
        char *p = (char *)e->d_match_io;
        float *q2 = (float *)p;
        p += (bq + 255) & ~(size_t)255;
        float *t2 = (float *)p;
        p += (bt + 255) & ~(size_t)255;
        dout = (surf_dmatch *)p;
        p += (bo + 255) & ~(size_t)255;
        dgood = (uint8_t *)p;
        CU(e, cudaMemcpyAsync(q2, query, bq, cudaMemcpyHostToDevice, e->stream));
        if (n_train) CU(e, cudaMemcpyAsync(t2, train, bt, cudaMemcpyHostToDevice, e->stream));
        dq = q2;
        dt = t2;
    }
    // (a train set of a few rows: the key lists' quad tournament would send 3 / n_train of the queries to the exact redo
    // anyway, and the exact kernel is as fast there)
    if (e->match_exact_only || n_train < 128) {
        if ((rc = reserve_match_part(e, n_query, n_train, k))) return rc;
        launch_match_knn(dq, n_query, dt, n_train, dim, k, dout, 0, e->d_match_part, e->stream, &e->launches);
        if (good) launch_ratio_test(dout, n_query, ratio, dgood, e->stream, &e->launches);
        e->last_tc_ws = -1;  // no candidate search ran: surf_match_redo_count reports 0
    } else {
        // tensor-core candidate search + exact rescoring: slot 0 = queries, slot 1 = train rows
        const int rows = ((n_query > n_train ? n_query : n_train) + 255) & ~255;
        int rc2 = tc_reserve(e, e->tc[0], 2, rows, dim);
        if (rc2) return rc2;
        surf_engine::TcWs &w = e->tc[0];
        const int max_q = (n_query + 127) & ~127;
        int sms = 148;
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, e->device);
        const int n_tiles = (n_train + 255) / 256, mt = max_q / 128;
        int splits = (2 * sms + mt - 1) / mt;
        if (splits > n_tiles) splits = n_tiles;
        if (splits < 1) splits = 1;
        if ((rc2 = tc_reserve_work(e, w, 1, splits, max_q, 2))) return rc2;
        TcSegment segs[2] = {{dq, nullptr, 0, n_query, 0}, {dt, nullptr, 0, n_train, 1}};
        TcProblem pr{};
        pr.q_slot = 0; pr.t_slot = 1; pr.qsrc = dq; pr.tsrc = dt;
        CU(e, cudaMemcpyAsync(w.segs, segs, sizeof(segs), cudaMemcpyHostToDevice, e->stream));
        CU(e, cudaMemcpyAsync(w.probs, &pr, sizeof(pr), cudaMemcpyHostToDevice, e->stream));
        TcMatchArgs a{};
        a.pool = w.pool; a.segs = w.segs; a.n_segs = 2; a.probs = w.probs; a.n_probs = 1; a.n_splits = splits;
        a.max_q_rows = max_q; a.cand = w.cand; a.counters = w.counters; a.redo_list = w.redo; a.dim = dim; a.k = k;
        a.ratio = ratio; a.out = dout; a.good = good ? dgood : nullptr;
        launch_tc_match(a, e->stream, &e->launches);
        e->last_tc_ws = 0;
    }
    CU(e, cudaGetLastError());
    if (mem == SURF_MEM_HOST) {
        CU(e, cudaMemcpyAsync(out, dout, sizeof(surf_dmatch) * (size_t)k * n_query, cudaMemcpyDeviceToHost, e->stream));
        if (good) CU(e, cudaMemcpyAsync(good, dgood, n_query, cudaMemcpyDeviceToHost, e->stream));
        CU(e, cudaStreamSynchronize(e->stream));
    }
    return SURF_OK;
}

int surf_match_knn(surf_engine *e, const float *query, int n_query, const float *train, int n_train, int dim, int k,
                   surf_dmatch *out, int mem)
{
    return match_common(e, query, n_query, train, n_train, dim, k, 0.f, out, nullptr, mem);
}

int surf_match_ratio(surf_engine *e, const float *query, int n_query, const float *train, int n_train, int dim,
                     float ratio, surf_dmatch *out, uint8_t *good, int mem)
{
    if (e && !good) return fail(e, SURF_ERR_BAD_ARG, "good is null");
    return match_common(e, query, n_query, train, n_train, dim, 2, ratio, out, good, mem);
}

int surf_set_match_mode(surf_engine *e, int exact_only)
{
    if (!e) return SURF_ERR_BAD_ARG;
    e->match_exact_only = exact_only ? 1 : 0;
    return SURF_OK;
}

int surf_set_graphs(surf_engine *e, int enable, int max_batch)
{
    if (!e || max_batch < 0) return SURF_ERR_BAD_ARG;
    e->use_graphs = enable ? 1 : 0;
    if (max_batch > 0) e->graph_max_batch = max_batch;
    if (!enable) {
        cudaStreamSynchronize(e->stream);
        drop_graphs(e);
    }
    return SURF_OK;
}

int surf_graph_replays(const surf_engine *e, long *replays)
{
    if (!e || !replays) return SURF_ERR_BAD_ARG;
    *replays = (long)e->graph_replays;
    return SURF_OK;
}

int surf_match_redo_count(surf_engine *e, int *count)
{
    if (!e || !count) return SURF_ERR_BAD_ARG;
    *count = 0;
    if (e->last_tc_ws < 0 || !e->tc[e->last_tc_ws].counters) return SURF_OK;
    int rc = ensure_device(e);
    if (rc) return rc;
    CU(e, cudaStreamSynchronize(e->stream));
    CU(e, cudaStreamSynchronize(e->match_stream));
    CU(e, cudaMemcpy(count, e->tc[e->last_tc_ws].counters + 1, sizeof(int), cudaMemcpyDeviceToHost));
    return SURF_OK;
}

int surf_stream_reset(surf_engine *e)
{
    if (!e) return SURF_ERR_BAD_ARG;
    int rc = ensure_device(e);
    if (rc) return rc;
    e->stream_primed = false;
    return SURF_OK;
}

int surf_device_matches(surf_engine *e, const surf_dmatch **matches, const uint8_t **good)
{
    if (!e) return SURF_ERR_BAD_ARG;
    if (matches) *matches = e->d_stream_matches;
    if (good) *good = e->d_stream_good;
    return SURF_OK;
}

int surf_match_prev(surf_engine *e, float ratio, surf_dmatch *out, uint8_t *good, int out_mem, long capacity)
{
    if (!e) return SURF_ERR_BAD_ARG;
    NvtxRange nvtx("surf_match_prev (stage 5)");
    const int s = e->out_collected;
    if (s < 0 || !e->out[s].collected || !e->out[s].want_desc)
        return fail(e, SURF_ERR_BAD_ARG, "surf_match_prev needs a submitted and collected batch with descriptors");
    OutSet &o = e->out[s];
    if (o.matched) return fail(e, SURF_ERR_BAD_ARG, "surf_match_prev was already called for this batch");
    int rc = ensure_device(e);
    if (rc) return rc;
    const int B = o.batch, D = e->params.extended ? 128 : 64;
    const long total = o.h_pin[2 * B + B];  // offsets[B] fetched by surf_collect_batch
    if ((out || good) && total > capacity) {  // nothing launched, stream state untouched: the caller may retry
        e->required = total;
        return fail(e, SURF_ERR_CAPACITY, "%ld keypoints in the batch; the match arrays hold %ld", total, capacity);
    }
    const int slots = e->max_b + 1, rows = (e->cap + 255) & ~255;
    surf_engine::TcWs &w = e->tc[1];
    if ((rc = tc_reserve(e, w, slots, rows, D))) return rc;
    if ((rc = tc_reserve_work(e, w, e->max_b, 1, rows, e->max_b))) return rc;
    const size_t nk = (size_t)e->max_b * e->cap;
    if (!e->d_prev_desc) {
        CU(e, cudaMalloc(&e->d_prev_desc, (size_t)e->cap * 128 * sizeof(float)));
        CU(e, cudaMalloc(&e->d_prev_off, 2 * sizeof(int)));
        CU(e, cudaMalloc(&e->d_stream_matches, nk * 2 * sizeof(surf_dmatch)));
        CU(e, cudaMalloc(&e->d_stream_good, nk));
    }
    o.matched = true;
    // The matcher runs on its own stream: it only depends on the batch just collected (whose kernels have ended:
    // the collect waited for ev_ready), so batches submitted meanwhile overlap the (latency-bound) match kernels.
    cudaStream_t ms = e->match_stream;
    CU(e, cudaStreamWaitEvent(ms, o.ev_ready, 0));
    if (!e->stream_primed) {  // forget the previous frame: an empty predecessor slot
        launch_zero(w.pool.count + e->tc_prev_slot, 1, ms, nullptr);
        launch_zero(e->d_prev_off, 2, ms, nullptr);
    }
    int launches = 0;
    CU(e, cudaEventRecord(e->ev_match_t0, ms));
    launch_tc_stream_tables(w.segs, w.probs, B, slots, e->tc_prev_slot, o.d_desc, o.d_off, e->d_prev_desc, e->d_prev_off, ms,
                            &launches);
    const int prev = (e->tc_prev_slot + B) % slots;
    // the previous batch's match results (d_stream_matches is one buffer) may still be on their way to the host
    if (e->match_copies_in_flight) CU(e, cudaStreamWaitEvent(ms, e->ev_match_copy_done, 0));
    TcMatchArgs a{};
    a.pool = w.pool; a.segs = w.segs; a.n_segs = B; a.probs = w.probs; a.n_probs = B; a.n_splits = 1;
    a.max_q_rows = rows; a.cand = w.cand; a.counters = w.counters; a.redo_list = w.redo; a.dim = D; a.k = 2;
    a.ratio = ratio; a.out = e->d_stream_matches; a.good = e->d_stream_good;
    launch_tc_match(a, ms, &launches);
    e->last_tc_ws = 1;
    // keep the last frame (fp32 rows + its slot) as the next batch's predecessor
    launch_copy_rows(o.d_desc, o.d_off + B - 1, D, e->d_prev_desc, e->cap, e->d_prev_off, ms, &launches);
    CU(e, cudaEventRecord(e->ev_match_t1, ms));
    CU(e, cudaEventRecord(o.ev_match_done, ms));
    o.match_in_flight = true;
    e->match_launches = launches;
    e->launches = launches;
    e->timings.launches = launches;
    CU(e, cudaGetLastError());
    e->tc_prev_slot = prev;
    e->stream_primed = true;
    if (out || good) {
        const cudaMemcpyKind kind = out_mem == SURF_MEM_DEVICE ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost;
        cudaStream_t cs = ms;
        if (e->async_outputs) {  // copy stream waits for the match kernels, the host does not
            CU(e, cudaStreamWaitEvent(e->copy_stream, o.ev_match_done, 0));
            cs = e->copy_stream;
        }
        if (out && total) CU(e, cudaMemcpyAsync(out, e->d_stream_matches, sizeof(surf_dmatch) * 2 * total, kind, cs));
        if (good && total) CU(e, cudaMemcpyAsync(good, e->d_stream_good, total, kind, cs));
        if (e->async_outputs) {
            CU(e, cudaEventRecord(e->ev_match_copy_done, e->copy_stream));
            e->match_copies_in_flight = true;
        } else {
            CU(e, cudaStreamSynchronize(ms));
        }
    }
    return SURF_OK;
}

int surf_match_prev_time(surf_engine *e, float *ms, int *launches)
{
    if (!e || !ms) return SURF_ERR_BAD_ARG;
    *ms = 0.f;
    if (launches) *launches = e->match_launches;
    if (!e->stream_primed) return SURF_OK;  // no surf_match_prev since the last reset
    int rc = ensure_device(e);
    if (rc) return rc;
    CU(e, cudaEventSynchronize(e->ev_match_t1));
    CU(e, cudaEventElapsedTime(ms, e->ev_match_t0, e->ev_match_t1));
    return SURF_OK;
}

int surf_mark(surf_engine *e, int slot)
{
    if (!e || slot < 0 || slot >= SURF_MAX_MARKS) return SURF_ERR_BAD_ARG;
    int rc = ensure_device(e);
    if (rc) return rc;
    if (!e->marks[slot]) CU(e, cudaEventCreate(&e->marks[slot]));
    CU(e, cudaEventRecord(e->marks[slot], e->stream));
    return SURF_OK;
}

int surf_mark_elapsed(surf_engine *e, int from_slot, int to_slot, float *ms)
{
    if (!e || !ms || from_slot < 0 || from_slot >= SURF_MAX_MARKS || to_slot < 0 || to_slot >= SURF_MAX_MARKS)
        return SURF_ERR_BAD_ARG;
    if (!e->marks[from_slot] || !e->marks[to_slot]) return fail(e, SURF_ERR_BAD_ARG, "mark not recorded");
    int rc = ensure_device(e);
    if (rc) return rc;
    CU(e, cudaEventSynchronize(e->marks[to_slot]));
    CU(e, cudaEventElapsedTime(ms, e->marks[from_slot], e->marks[to_slot]));
    return SURF_OK;
}

int surf_merge_topk(surf_engine *e, const surf_dmatch *in, int n_ranks, int n_query, int k, surf_dmatch *out)
{
    if (!e) return SURF_ERR_BAD_ARG;
    if (!in || !out || n_ranks <= 0 || n_query < 0 || k < 1 || k > 4) return fail(e, SURF_ERR_BAD_ARG, "bad argument");
    if (n_query == 0) return SURF_OK;
    int rc = ensure_device(e);
    if (rc) return rc;
    launch_merge_topk(in, n_ranks, n_query, k, out, e->stream, &e->launches);
    CU(e, cudaGetLastError());
    return SURF_OK;
}

}  // extern "C"
