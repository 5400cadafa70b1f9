// surf_engine_priv.h -- the engine object and the helpers shared by the host-side translation units
// (surf_engine.cu: stages 1-5 behind the C ABI; surf_db.cu: descriptor database, correspondence votes).
// Private to libsurf_b200.so; the public boundary is include/surf_b200.h.
#pragma once

#include <cstdarg>
#include <cstdio>
#include <vector>

#include <cuda.h>
#include <nvtx3/nvToolsExt.h>   // header-only NVTX v3: ranges show up in Nsight Systems / ncu --nvtx, cost nothing otherwise

#include "surf_internal.h"

// Host-side NVTX range around a phase of the C ABI (tracing).
struct NvtxRange {
    explicit NvtxRange(const char *name) { nvtxRangePushA(name); }
    ~NvtxRange() { nvtxRangePop(); }
    NvtxRange(const NvtxRange &) = delete;
    NvtxRange &operator=(const NvtxRange &) = delete;
};

using namespace surfb200;  // private header of the library's own translation units

enum Ev { EV_START, EV_H2D, EV_INTEGRAL, EV_HESSIAN_NMS, EV_SORT, EV_DESC_START, EV_DESCRIBE, EV_PACK, EV_END, EV_COUNT };

// Results of one submitted batch.  There are two sets, used alternately, so that surf_submit_batch of batch i+1 may
// precede surf_collect_batch of batch i: the host then never keeps the GPU waiting for its own wake-up.
struct OutSet {
    surf_keypoint *d_kp = nullptr;   // packed keypoints
    float *d_desc = nullptr;         // packed descriptors [rows][D]
    int *d_off = nullptr;            // offsets[B + 1]
    int *h_pin = nullptr;            // pinned: counts[B], ndel[B], off_out[B+1], misc, status
    cudaEvent_t ev_ready = nullptr;  // this batch's kernels and counter copies have ended (main stream)
    cudaEvent_t ev_copy_done = nullptr, ev_match_done = nullptr;   // readers of this set: result copies, stream matcher
    bool copies_in_flight = false, match_in_flight = false;
    cudaEvent_t *ev = nullptr;       // stage events of the batch (one array of the engine's pool)
    bool active = false;             // submitted, not yet collected
    bool collected = false;          // collected: surf_match_prev / surf_device_results refer to it
    bool matched = false;            // surf_match_prev already ran on this batch
    int batch = 0;
    bool want_desc = false;
    int launches = 0;
};

struct surf_engine {
    surf_params params{};
    int device = 0;
    int max_w = 0, max_h = 0, max_b = 0, cap = 0;
    cudaStream_t stream = nullptr;
    cudaStream_t copy_stream = nullptr;  // device->host result copies in asynchronous-output mode
    cudaStream_t match_stream = nullptr; // surf_match_prev runs here, concurrently with the next batch's kernels
    cudaEvent_t ev_match_t0 = nullptr, ev_match_t1 = nullptr;  // timed: the last surf_match_prev on match_stream
    int match_launches = 0;
    cudaEvent_t marks[SURF_MAX_MARKS]{};                       // surf_mark
    int async_outputs = 0;
    cudaEvent_t ev_match_copy_done = nullptr;   // the match results (one device buffer) are on their way to the host
    bool match_copies_in_flight = false;
    // stage events: a pool of arrays, one per batch in flight (two submitted + one look-ahead at most); `ev` is the
    // array the launch sequence being enqueued records into
    static constexpr int kEvSets = 4;
    cudaEvent_t evpool[kEvSets][EV_COUNT]{};
    int ev_next = 0;
    cudaEvent_t *ev = nullptr;
    bool ev_valid = false;
    OutSet out[2];
    int out_next = 0;        // set the next surf_submit_batch writes
    int out_bound = 0;       // set whose buffers d_out_kp / d_out_desc / d_off_out / h_pin currently alias
    int out_collected = -1;  // set of the batch collected last

    // workspace
    // Host frames are staged in a rotation of three device buffers (see next_frame_buffer): buf_cur holds the frames of
    // the batch submitted last, buf_staged those of a prefetch / look-ahead not yet adopted (-1: none).
    uint8_t *imgbuf[3] = {nullptr, nullptr, nullptr};   size_t img_pitch = 0;
    int buf_cur = 0, buf_staged = -1;
    cudaEvent_t ev_buf_free[3] = {nullptr, nullptr, nullptr};  // end of the last kernels that read the buffer
    bool buf_free_rec[3] = {false, false, false};
    cudaStream_t upload_stream = nullptr;   // uploads of surf_prefetch_batch / surf_lookahead_batch
    cudaEvent_t ev_upload_done = nullptr;
    bool la_done_recorded = false;
    int early_upload = 0;   // look-ahead uploads on the upload stream at once (1) or at the head of the look-ahead stream (0)
    struct Prefetched {
        const uint8_t *images = nullptr;
        int batch = 0, width = 0, height = 0;
        size_t pitch = 0, frame_stride = 0;
        bool valid = false;
    } prefetched;
    // surf_lookahead_batch: stages 1-3 of the NEXT batch run on their own stream, into the second detect set, while
    // the submitted batch is in its descriptor stage.  A detect set is what stages 1-3 write and stage 4 reads:
    // integral images, the two keypoint buffers, the per-frame counters, and the TMA maps of the integral images.
    // The members below (d_sum, d_kp_a, d_kp_b, d_counts, sorted, tmap*) are the CURRENT set; `alt` holds the other one.
    struct DetectSet {
        uint32_t *d_sum = nullptr;
        surf_keypoint *d_kp_a = nullptr, *d_kp_b = nullptr, *sorted = nullptr;
        int *d_counts = nullptr;
        CUtensorMap tmap[2];
        TileOctave tile_oct[2];
        int tmap_w = 0, tmap_h = 0, tmap_b = 0, tmap_layers = 0;
        bool tmap_ok[2] = {false, false};
        cudaEvent_t ev_free = nullptr;   // end of the stage-4 kernels that last read this set
        bool free_recorded = false;
    } alt;
    cudaEvent_t ev_set_free = nullptr;   // the current set's ev_free
    bool set_free_recorded = false;
    cudaStream_t la_stream = nullptr;
    cudaEvent_t *la_ev = nullptr;        // stage-event array of the pending look-ahead
    cudaEvent_t ev_detect_done = nullptr, ev_la_done = nullptr;
    bool detect_done_recorded = false;
    struct LookAhead {
        bool valid = false;
        const uint8_t *images = nullptr, *mask = nullptr;
        int in_mem = 0, batch = 0, width = 0, height = 0;
        size_t pitch = 0, frame_stride = 0, mask_pitch = 0;
        ImageRef img{};                  // device view of the frames the look-ahead ran on
        int launches = 0;
    } la;
    uint8_t *d_mask = nullptr;
    uint32_t *d_sum = nullptr, *d_msum = nullptr, *d_bandsum = nullptr;
    int *d_bandflags = nullptr;  // ticket / exit counters and per-band flags of the integral kernel
    float *d_det = nullptr;     size_t det_words = 0;
    surf_keypoint *d_kp_a = nullptr, *d_kp_b = nullptr, *d_out_kp = nullptr;
    float *d_desc = nullptr, *d_out_desc = nullptr;  int desc_dim = 0;
    uint8_t *d_patches = nullptr;
    int *d_counts = nullptr, *d_ndel = nullptr, *d_off_in = nullptr, *d_off_out = nullptr, *d_misc = nullptr;
    int *d_deleted = nullptr, *d_class_list = nullptr;
    float2 *d_sincos = nullptr;
    int *h_pin = nullptr;  // alias of out[out_bound].h_pin (as are d_out_kp, d_out_desc, d_off_out)
    surf_dmatch *d_match_part = nullptr;  size_t match_part_cap = 0;
    void *d_match_io = nullptr;           size_t match_io_cap = 0;
    int match_exact_only = 0;
    // CUDA graphs of the small-batch path: the launch sequence of stages 1-4 + the count copies of one submit, keyed on
    // every pointer and size it bakes in; captured the second time a key is seen, replayed afterwards
    struct GraphKey {
        const void *img, *msum, *d_sum, *d_kp_a, *d_kp_b, *d_counts, *ev, *h_pin;
        const void *ev_detect_done, *ev_set_free;  // handles the look-ahead logic swaps around: baked into the graph, too
        size_t pitch, fstride;
        int W, H, B, want_desc, out_set, cap;
    };
    struct GraphEntry {
        GraphKey key;
        cudaGraphExec_t exec = nullptr;
        surf_keypoint *sorted = nullptr;
        int launches = 0;
        int seen = 0;
        bool failed = false;
        unsigned long last_use = 0;
    };
    std::vector<GraphEntry> graphs;
    int use_graphs = 1;
    int graph_max_batch = 16;
    bool capturing = false;
    unsigned long graph_clock = 0, graph_replays = 0;
    int use_tma = 1;                 // fused TMA-staged Hessian + maxima kernel for octaves 0-1
    CUtensorMap tmap[2];             // integral-image tile maps of octaves 0 and 1
    TileOctave tile_oct[2];
    int tmap_w = 0, tmap_h = 0, tmap_b = 0, tmap_layers = 0;
    bool tmap_ok[2] = {false, false};
    CUtensorMap imap[kNumClasses];   // 8-bit frame tiles for the descriptor kernels, one box size per class
    bool imap_ok = false;
    const void *imap_ptr = nullptr;  size_t imap_pitch = 0, imap_stride = 0;  int imap_w = 0, imap_h = 0, imap_b = 0;
    int last_tc_ws = -1;  // workspace used by the last tensor-core match (for surf_match_redo_count)

    // tensor-core matcher workspaces: [0] generic two-operand calls, [1] the frame stream ring,
    // [2] queries against a descriptor database (surf_db.cu)
    struct TcWs {
        TcPool pool{};
        int slots = 0, dim = 0;
        TcCand *cand = nullptr;      size_t cand_cap = 0;
        int *redo = nullptr;         size_t redo_cap = 0;
        int *counters = nullptr;
        TcSegment *segs = nullptr;   TcProblem *probs = nullptr;  int table_cap = 0;
    } tc[3];
    float *d_prev_desc = nullptr;    // last frame of the previous batch (fp32 rows) for exact rescoring
    int *d_prev_off = nullptr;       // {0, n_prev}
    int tc_prev_slot = 0;
    bool stream_primed = false;
    surf_dmatch *d_stream_matches = nullptr;
    uint8_t *d_stream_good = nullptr;

    struct surf_comm *solo = nullptr;  // single-GPU context of surf_sweep_sharded (surf_comm.cu), destroyed with the engine
    surf_keypoint *sorted = nullptr;  // which of d_kp_a / d_kp_b holds the current keypoints
    surf_timings timings{};
    int launches = 0;
    long required = 0;
    char err[512] = "";
};


struct surf_db {
    surf_engine *e = nullptr;
    int dim = 64;
    long max_rows = 0;
    int max_images = 0;
    TcPool pool{};                  // one slot of slot_rows >= max_rows rows
    float *rows = nullptr;          // fp32 [max_rows][dim]
    surf_keypoint *kps = nullptr;   // [max_rows]
    int *d_img_off = nullptr;       // [max_images + 1]
    TcSegment *d_seg = nullptr;     // one table entry for the append being prepared
    std::vector<int> img_off;       // host copy, img_off[0] = 0
    bool has_kps = true;            // false once an add() came without keypoints
    void *vote_ws = nullptr;        // scratch of surf_correspond (owned by the train database)
    size_t vote_ws_cap = 0;
};

namespace surfb200 {

inline long db_rows(const surf_db *db) { return db->img_off.empty() ? 0 : db->img_off.back(); }
// k-NN of device-resident query rows against a database on the engine stream; records carry (img_idx, row within
// the image) when `localize`, else (0, row within the database)
int db_knn_device(surf_db *db, const float *dq, int n_query, int k, surf_dmatch *dout, int *launches, bool localize = true);

int fail(surf_engine *e, int code, const char *fmt, ...);
int ensure_device(surf_engine *e);
// (re)allocate the prepared-descriptor pool / the candidate, redo and table buffers of a matcher workspace
int tc_reserve(surf_engine *e, surf_engine::TcWs &w, int slots, int slot_rows, int dim);
int tc_reserve_work(surf_engine *e, surf_engine::TcWs &w, int n_probs, int splits, int max_q, int n_segs);
// scratch of the exact CUDA-core matcher (partial top-k records)
int reserve_match_part(surf_engine *e, int n_query, int n_train, int k);

}  // namespace surfb200

#define CU(e, call)                                                                                          \
    do {                                                                                                     \
        cudaError_t _r = (call);                                                                             \
        if (_r != cudaSuccess)                                                                               \
            return surfb200::fail((e), SURF_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(_r), \
                                  __FILE__, __LINE__);                                                       \
    } while (0)
