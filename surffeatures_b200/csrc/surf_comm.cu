// surf_comm.cu -- the multi-GPU exchange steps of SURVEY.md section 8(e) behind the C ABI, NCCL on device buffers:
//   * surf_sweep_sharded   computeNext (vtkSlicerSurfFeaturesLogic.cxx:1392-1406) over a frame-sharded sweep; every
//                          batch's packed keypoints / descriptors go to the destination rank by grouped
//                          ncclSend / ncclRecv while the next batch computes (BASELINE config 4)
//   * surf_db_knn_sharded  matchDescriptorsKNN (:635-643) against a row-sharded train database: per-rank top-k,
//                          ncclAllGather of 16 B x k x Nq records, k_merge_topk (BASELINE config 5)
// One process per GPU.  libnccl.so.2 is resolved with dlopen: libsurf_b200.so itself has no NCCL dependency, and
// inside a torch process the copy torch already loaded is the one used.
#include <dlfcn.h>

#include <cstring>
#include <new>
#include <vector>

#include "surf_engine_priv.h"

namespace {

// The slice of the NCCL ABI used here (nccl.h 2.x: stable since 2.7 for send / recv).
typedef struct { char internal[128]; } NcclUniqueId;
typedef void *NcclComm;
enum { kNcclSuccess = 0, kNcclInt8 = 0, kNcclInt32 = 2, kNcclFloat32 = 7 };

struct NcclApi {
    void *lib = nullptr;
    int (*GetVersion)(int *) = nullptr;
    int (*GetUniqueId)(NcclUniqueId *) = nullptr;
    int (*CommInitRank)(NcclComm *, int, NcclUniqueId, int) = nullptr;
    int (*CommDestroy)(NcclComm) = nullptr;
    int (*AllGather)(const void *, void *, size_t, int, NcclComm, cudaStream_t) = nullptr;
    int (*Broadcast)(const void *, void *, size_t, int, int, NcclComm, cudaStream_t) = nullptr;
    int (*Send)(const void *, size_t, int, int, NcclComm, cudaStream_t) = nullptr;
    int (*Recv)(void *, size_t, int, int, NcclComm, cudaStream_t) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    const char *(*GetErrorString)(int) = nullptr;
    char err[256] = "";
};

NcclApi *nccl_api()
{
    static NcclApi api;
    static bool tried = false;
    if (tried) return api.lib ? &api : nullptr;
    tried = true;
    // the copy already mapped into the process (torch ships its own libnccl.so.2) wins over the system one
    void *h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h) {
        snprintf(api.err, sizeof(api.err), "libnccl.so.2 not found: %s", dlerror());
        return nullptr;
    }
    bool ok = true;
    auto sym = [&](const char *name) {
        void *p = dlsym(h, name);
        if (!p) {
            ok = false;
            snprintf(api.err, sizeof(api.err), "libnccl.so.2 lacks %s", name);
        }
        return p;
    };
    api.GetVersion = (int (*)(int *))sym("ncclGetVersion");
    api.GetUniqueId = (int (*)(NcclUniqueId *))sym("ncclGetUniqueId");
    api.CommInitRank = (int (*)(NcclComm *, int, NcclUniqueId, int))sym("ncclCommInitRank");
    api.CommDestroy = (int (*)(NcclComm))sym("ncclCommDestroy");
    api.AllGather = (int (*)(const void *, void *, size_t, int, NcclComm, cudaStream_t))sym("ncclAllGather");
    api.Broadcast = (int (*)(const void *, void *, size_t, int, int, NcclComm, cudaStream_t))sym("ncclBroadcast");
    api.Send = (int (*)(const void *, size_t, int, int, NcclComm, cudaStream_t))sym("ncclSend");
    api.Recv = (int (*)(void *, size_t, int, int, NcclComm, cudaStream_t))sym("ncclRecv");
    api.GroupStart = (int (*)())sym("ncclGroupStart");
    api.GroupEnd = (int (*)())sym("ncclGroupEnd");
    api.GetErrorString = (const char *(*)(int))sym("ncclGetErrorString");
    if (!ok) return nullptr;
    api.lib = h;
    return &api;
}

const char *nccl_load_error()
{
    static NcclApi probe;  // nccl_api() failed: say why in one line
    void *h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);
    snprintf(probe.err, sizeof(probe.err), "%s", h ? "libnccl.so.2 lacks a required symbol" : "libnccl.so.2 could not be loaded");
    return probe.err;
}

}  // namespace

struct surf_comm {
    surf_engine *e = nullptr;
    NcclApi *api = nullptr;
    NcclComm comm = nullptr;
    int rank = 0, n_ranks = 1, version = 0;
    cudaStream_t stream = nullptr;            // NCCL operations of the sweep run here, beside the engine's kernels
    // sweep: per-frame counts of one chunk, double-buffered
    int *d_cnt_local[2] = {nullptr, nullptr}, *d_cnt_all[2] = {nullptr, nullptr};
    int *h_cnt_local[2] = {nullptr, nullptr}, *h_cnt_all[2] = {nullptr, nullptr};
    int cnt_cap = 0;                          // ints per local buffer
    cudaEvent_t ev_cnt[2] = {nullptr, nullptr};
    std::vector<cudaEvent_t> ev_x;            // pairs of timing events around every payload exchange
    cudaEvent_t ev_t[6] = {};                 // timing of the sweep / of surf_db_knn_sharded
    // gathered blocks: region r holds rank r's rows (only this rank's own region exists off the destination)
    surf_keypoint *stage_kp = nullptr;  size_t stage_kp_rows = 0;
    float *stage_desc = nullptr;        size_t stage_desc_floats = 0;
    // sharded database match
    int *d_rows = nullptr;                    // [1 + n_ranks]: own row count, then everybody's
    int *h_rows = nullptr;                    // pinned
    void *d_knn = nullptr;  size_t knn_cap = 0;
    float last_total_ms = 0.f, last_nccl_ms = 0.f;
};

namespace {

#define NC(e, c, call)                                                                                          \
    do {                                                                                                        \
        int _r = (call);                                                                                        \
        if (_r != kNcclSuccess)                                                                                 \
            return surfb200::fail((e), SURF_ERR_NCCL, "%s failed: %s (%s:%d)", #call,                          \
                                  (c)->api->GetErrorString ? (c)->api->GetErrorString(_r) : "?", __FILE__, __LINE__); \
    } while (0)

int reserve_counts(surf_comm *c, int B)
{
    surf_engine *e = c->e;
    if (B <= c->cnt_cap) return SURF_OK;
    for (int i = 0; i < 2; i++) {
        cudaFree(c->d_cnt_local[i]); cudaFree(c->d_cnt_all[i]);
        if (c->h_cnt_local[i]) cudaFreeHost(c->h_cnt_local[i]);
        if (c->h_cnt_all[i]) cudaFreeHost(c->h_cnt_all[i]);
        c->d_cnt_local[i] = c->d_cnt_all[i] = c->h_cnt_local[i] = c->h_cnt_all[i] = nullptr;
    }
    c->cnt_cap = 0;
    for (int i = 0; i < 2; i++) {
        CU(e, cudaMalloc(&c->d_cnt_local[i], sizeof(int) * B));
        CU(e, cudaMalloc(&c->d_cnt_all[i], sizeof(int) * B * c->n_ranks));
        CU(e, cudaMallocHost(&c->h_cnt_local[i], sizeof(int) * B));
        CU(e, cudaMallocHost(&c->h_cnt_all[i], sizeof(int) * B * c->n_ranks));
    }
    c->cnt_cap = B;
    return SURF_OK;
}

int reserve_stage(surf_comm *c, size_t rows, int D, bool want_desc)
{
    surf_engine *e = c->e;
    if (rows > c->stage_kp_rows) {
        cudaFree(c->stage_kp);
        c->stage_kp = nullptr;
        c->stage_kp_rows = 0;
        CU(e, cudaMalloc(&c->stage_kp, (rows ? rows : 1) * sizeof(surf_keypoint)));
        c->stage_kp_rows = rows;
    }
    const size_t fl = want_desc ? rows * (size_t)D : 0;
    if (fl > c->stage_desc_floats) {
        cudaFree(c->stage_desc);
        c->stage_desc = nullptr;
        c->stage_desc_floats = 0;
        CU(e, cudaMalloc(&c->stage_desc, fl * sizeof(float)));
        c->stage_desc_floats = fl;
    }
    return SURF_OK;
}

// A communicator for the single-GPU case (comm == NULL at the ABI): same code path, no NCCL call.  Owned by the engine.
surf_comm *solo_comm(surf_engine *e)
{
    if (e->solo) return e->solo;
    surf_comm *c = new (std::nothrow) surf_comm();
    if (!c) return nullptr;
    c->e = e;
    if (cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) != cudaSuccess) {
        delete c;
        return nullptr;
    }
    for (auto &ev : c->ev_cnt) cudaEventCreateWithFlags(&ev, cudaEventDisableTiming);
    for (auto &ev : c->ev_t) cudaEventCreate(&ev);
    e->solo = c;
    return c;
}

__global__ void k_add_row_base(surf_dmatch *m, long n, const int *rows_all, int rank)
{
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int base = 0;
    for (int r = 0; r < rank; r++) base += rows_all[r];
    surf_dmatch v = m[i];
    if (v.train_idx >= 0) {
        v.train_idx += base;
        v.img_idx = 0;
        m[i] = v;
    }
}

}  // namespace

extern "C" {

int surf_shard_range(long n_items, int rank, int n_ranks, long *lo, long *hi)
{
    if (n_items < 0 || n_ranks <= 0 || rank < 0 || rank >= n_ranks || !lo || !hi) return SURF_ERR_BAD_ARG;
    const long base = n_items / n_ranks, extra = n_items % n_ranks;
    *lo = rank * base + (rank < extra ? rank : extra);
    *hi = *lo + base + (rank < extra ? 1 : 0);
    return SURF_OK;
}

int surf_comm_unique_id(void *id)
{
    if (!id) return SURF_ERR_BAD_ARG;
    NcclApi *api = nccl_api();
    if (!api) {
        fprintf(stderr, "surf_comm_unique_id: %s\n", nccl_load_error());
        return SURF_ERR_NCCL;
    }
    NcclUniqueId u;
    if (api->GetUniqueId(&u) != kNcclSuccess) return SURF_ERR_NCCL;
    memcpy(id, u.internal, SURF_COMM_ID_BYTES);
    return SURF_OK;
}

int surf_comm_init(surf_engine *e, const void *id, int rank, int n_ranks, surf_comm **out)
{
    if (!e || !out) return SURF_ERR_BAD_ARG;
    *out = nullptr;
    if (!id || n_ranks <= 0 || rank < 0 || rank >= n_ranks) return fail(e, SURF_ERR_BAD_ARG, "bad rank / n_ranks / id");
    NcclApi *api = nccl_api();
    if (!api) return fail(e, SURF_ERR_NCCL, "%s", nccl_load_error());
    int rc = ensure_device(e);
    if (rc) return rc;
    surf_comm *c = new (std::nothrow) surf_comm();
    if (!c) return fail(e, SURF_ERR_CUDA, "out of host memory");
    c->e = e;
    c->api = api;
    c->rank = rank;
    c->n_ranks = n_ranks;
    api->GetVersion(&c->version);
    NcclUniqueId u;
    memcpy(u.internal, id, SURF_COMM_ID_BYTES);
    int r = api->CommInitRank(&c->comm, n_ranks, u, rank);
    if (r != kNcclSuccess) {
        rc = fail(e, SURF_ERR_NCCL, "ncclCommInitRank failed: %s", api->GetErrorString(r));
        delete c;
        return rc;
    }
    cudaError_t cr = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking);
    for (auto &ev : c->ev_cnt)
        if (cr == cudaSuccess) cr = cudaEventCreateWithFlags(&ev, cudaEventDisableTiming);
    for (auto &ev : c->ev_t)
        if (cr == cudaSuccess) cr = cudaEventCreate(&ev);
    if (cr == cudaSuccess) cr = cudaMalloc(&c->d_rows, sizeof(int) * (1 + n_ranks));
    if (cr == cudaSuccess) cr = cudaMallocHost(&c->h_rows, sizeof(int) * (1 + n_ranks));
    if (cr != cudaSuccess) {
        rc = fail(e, SURF_ERR_CUDA, "communicator resources: %s", cudaGetErrorString(cr));
        surf_comm_destroy(c);
        return rc;
    }
    *out = c;
    return SURF_OK;
}

void surf_comm_destroy(surf_comm *c)
{
    if (!c) return;
    cudaSetDevice(c->e->device);
    if (c->stream) cudaStreamSynchronize(c->stream);
    cudaStreamSynchronize(c->e->stream);
    if (c->comm && c->api) c->api->CommDestroy(c->comm);
    for (int i = 0; i < 2; i++) {
        cudaFree(c->d_cnt_local[i]); cudaFree(c->d_cnt_all[i]);
        if (c->h_cnt_local[i]) cudaFreeHost(c->h_cnt_local[i]);
        if (c->h_cnt_all[i]) cudaFreeHost(c->h_cnt_all[i]);
        if (c->ev_cnt[i]) cudaEventDestroy(c->ev_cnt[i]);
    }
    for (auto ev : c->ev_x) cudaEventDestroy(ev);
    for (auto ev : c->ev_t)
        if (ev) cudaEventDestroy(ev);
    cudaFree(c->stage_kp); cudaFree(c->stage_desc); cudaFree(c->d_rows); cudaFree(c->d_knn);
    if (c->h_rows) cudaFreeHost(c->h_rows);
    if (c->stream) cudaStreamDestroy(c->stream);
    delete c;
}

int surf_comm_info(const surf_comm *c, int *rank, int *n_ranks, int *nccl_version)
{
    if (!c) return SURF_ERR_BAD_ARG;
    if (rank) *rank = c->rank;
    if (n_ranks) *n_ranks = c->n_ranks;
    if (nccl_version) *nccl_version = c->version;
    return SURF_OK;
}

int surf_comm_last_times(const surf_comm *c, float *total_ms, float *nccl_ms)
{
    if (!c) return SURF_ERR_BAD_ARG;
    if (total_ms) *total_ms = c->last_total_ms;
    if (nccl_ms) *nccl_ms = c->last_nccl_ms;
    return SURF_OK;
}

int surf_sweep_sharded(surf_engine *e, surf_comm *comm, const uint8_t *images_local, int in_mem, long n_frames_total,
                       int width, int height, size_t pitch, size_t frame_stride, const uint8_t *mask, size_t mask_pitch,
                       int want_descriptors, int dst_rank, surf_keypoint *keypoints, float *descriptors, int out_mem,
                       long capacity, int64_t *offsets, surf_sweep_stats *stats)
{
    if (!e) return SURF_ERR_BAD_ARG;
    NvtxRange nvtx("surf_sweep_sharded");
    if (comm && comm->e != e) return fail(e, SURF_ERR_BAD_ARG, "the communicator belongs to another engine");
    int rc = ensure_device(e);
    if (rc) return rc;
    surf_comm *c = comm ? comm : solo_comm(e);
    if (!c) return fail(e, SURF_ERR_CUDA, "could not create the single-GPU sweep context");
    const int world = c->n_ranks, rank = c->rank;
    if (n_frames_total < 0 || dst_rank < 0 || dst_rank >= world) return fail(e, SURF_ERR_BAD_ARG, "bad frame count or dst_rank");
    const bool is_dst = rank == dst_rank;
    if (is_dst && (!offsets || (n_frames_total > 0 && !keypoints) || (want_descriptors && n_frames_total > 0 && !descriptors)))
        return fail(e, SURF_ERR_BAD_ARG, "the destination rank needs keypoints, descriptors and offsets");
    if (out_mem != SURF_MEM_HOST && out_mem != SURF_MEM_DEVICE) return fail(e, SURF_ERR_BAD_ARG, "bad out_mem");
    long lo = 0, hi = 0;
    surf_shard_range(n_frames_total, rank, world, &lo, &hi);
    const long n_local = hi - lo;
    if (n_local > 0 && !images_local) return fail(e, SURF_ERR_BAD_ARG, "image pointer is null");
    const int B = e->max_b, D = e->params.extended ? 128 : 64;
    const long max_local = (n_frames_total + world - 1) / world;
    const int n_chunks = (int)((max_local + B - 1) / B);
    if ((rc = reserve_counts(c, B))) return rc;
    // regions: this rank's own rows; on the destination one more per source rank
    std::vector<size_t> region_rows(world, 0), region_base(world + 1, 0);
    for (int r = 0; r < world; r++) {
        long rlo, rhi;
        surf_shard_range(n_frames_total, r, world, &rlo, &rhi);
        if (r == rank || is_dst) region_rows[r] = (size_t)(rhi - rlo) * e->cap;
        region_base[r + 1] = region_base[r] + region_rows[r];
    }
    if ((rc = reserve_stage(c, region_base[world], D, want_descriptors != 0))) return rc;
    while ((int)c->ev_x.size() < 2 * (n_chunks + 1)) {
        cudaEvent_t ev;
        CU(e, cudaEventCreate(&ev));
        c->ev_x.push_back(ev);
    }
    const int was_async = e->async_outputs;
    if (was_async && (rc = surf_set_async_outputs(e, 0))) return rc;

    std::vector<int64_t> frame_counts((size_t)n_frames_total, 0);
    std::vector<size_t> got_rows(world, 0);        // rows of region r filled so far
    std::vector<size_t> chunk_row0(n_chunks + 1, 0);  // this rank's first row of every chunk
    std::vector<long> rank_lo(world);
    for (int r = 0; r < world; r++) {
        long rhi;
        surf_shard_range(n_frames_total, r, world, &rank_lo[r], &rhi);
    }
    surf_keypoint *own_kp = c->stage_kp + region_base[rank];
    float *own_desc = want_descriptors ? c->stage_desc + region_base[rank] * D : nullptr;
    int first_error = SURF_OK;
    char first_msg[sizeof(e->err)] = "";
    int launches = 0;
    int64_t sent = 0, received = 0;
    int n_exchanges = 0;
    std::vector<int32_t> off32(B + 1, 0);

    // payload of chunk x: counts known (ev_cnt), rows sit in the regions
    auto exchange = [&](int x) -> int {
        const int s = x & 1;
        CU(e, cudaEventSynchronize(c->ev_cnt[s]));
        const int *all = world > 1 ? c->h_cnt_all[s] : c->h_cnt_local[s];
        std::vector<size_t> n_r(world, 0);
        for (int r = 0; r < world; r++) {
            long rlo, rhi;
            surf_shard_range(n_frames_total, r, world, &rlo, &rhi);
            for (int i = 0; i < B; i++) {
                const long f = rlo + (long)x * B + i;
                const int cnt = all[r * B + i];
                if (f < rhi) frame_counts[f] = cnt;
                n_r[r] += cnt;
            }
        }
        if (world > 1) {
            CU(e, cudaEventRecord(c->ev_x[2 * n_exchanges], c->stream));
            NC(e, c, c->api->GroupStart());
            if (!is_dst && n_r[rank] > 0) {
                NC(e, c, c->api->Send(own_kp + chunk_row0[x], n_r[rank] * sizeof(surf_keypoint), kNcclInt8, dst_rank, c->comm, c->stream));
                sent += (int64_t)(n_r[rank] * sizeof(surf_keypoint));
                if (want_descriptors) {
                    NC(e, c, c->api->Send(own_desc + chunk_row0[x] * D, n_r[rank] * D, kNcclFloat32, dst_rank, c->comm, c->stream));
                    sent += (int64_t)(n_r[rank] * D * sizeof(float));
                }
            }
            if (is_dst) {
                for (int r = 0; r < world; r++) {
                    if (r == rank || n_r[r] == 0) continue;
                    if (got_rows[r] + n_r[r] > region_rows[r])
                        return fail(e, SURF_ERR_CAPACITY, "rank %d sent more rows than its region holds", r);
                    NC(e, c, c->api->Recv(c->stage_kp + region_base[r] + got_rows[r], n_r[r] * sizeof(surf_keypoint), kNcclInt8, r,
                                          c->comm, c->stream));
                    received += (int64_t)(n_r[r] * sizeof(surf_keypoint));
                    if (want_descriptors) {
                        NC(e, c, c->api->Recv(c->stage_desc + (region_base[r] + got_rows[r]) * D, n_r[r] * D, kNcclFloat32, r, c->comm,
                                              c->stream));
                        received += (int64_t)(n_r[r] * D * sizeof(float));
                    }
                }
            }
            NC(e, c, c->api->GroupEnd());
            CU(e, cudaEventRecord(c->ev_x[2 * n_exchanges + 1], c->stream));
            n_exchanges++;
        }
        for (int r = 0; r < world; r++) got_rows[r] += n_r[r];
        return SURF_OK;
    };

    CU(e, cudaEventRecord(c->ev_t[0], e->stream));
    size_t local_rows = 0;
    auto chunk_frames = [&](int x) {
        const long f0 = (long)x * B;
        return (int)(n_local - f0 < 0 ? 0 : (n_local - f0 > B ? B : n_local - f0));
    };
    auto chunk_ptr = [&](int x) { return images_local + (size_t)x * B * frame_stride; };
    auto note_error = [&](int r) {
        if (r && first_error == SURF_OK) {  // keep the collective protocol going with empty chunks; report at the end
            first_error = r;
            snprintf(first_msg, sizeof(first_msg), "%s", e->err);
        }
    };
    // Two batches in flight: chunk x+1 is submitted (continuing from its look-ahead) and chunk x+2's stages 1-3 are
    // started before the host waits for chunk x, so the GPU always has the next descriptor stage queued.
    bool submitted_next = false;
    if (chunk_frames(0) > 0) {
        note_error(surf_submit_batch(e, chunk_ptr(0), in_mem, chunk_frames(0), width, height, pitch, frame_stride, mask, mask_pitch,
                                     want_descriptors));
        launches += e->launches;
        if (first_error == SURF_OK && chunk_frames(1) > 0)
            note_error(surf_lookahead_batch(e, chunk_ptr(1), in_mem, chunk_frames(1), width, height, pitch, frame_stride, mask,
                                            mask_pitch));
    }
    for (int x = 0; x < n_chunks; x++) {
        const int s = x & 1;
        const int nb = chunk_frames(x);
        chunk_row0[x] = local_rows;
        memset(c->h_cnt_local[s], 0, sizeof(int) * B);
        if (nb > 0 && first_error == SURF_OK) {
            submitted_next = false;
            if (chunk_frames(x + 1) > 0) {
                rc = surf_submit_batch(e, chunk_ptr(x + 1), in_mem, chunk_frames(x + 1), width, height, pitch, frame_stride, mask,
                                       mask_pitch, want_descriptors);
                launches += e->launches;
                submitted_next = !rc;
                if (!rc && chunk_frames(x + 2) > 0)
                    rc = surf_lookahead_batch(e, chunk_ptr(x + 2), in_mem, chunk_frames(x + 2), width, height, pitch, frame_stride,
                                              mask, mask_pitch);
                note_error(rc);
            }
            rc = surf_collect_batch(e, own_kp + local_rows, own_desc ? own_desc + local_rows * D : nullptr, SURF_MEM_DEVICE,
                                    (long)(region_rows[rank] - local_rows), off32.data());
            note_error(rc);
            if (!rc) {
                for (int i = 0; i < nb; i++) c->h_cnt_local[s][i] = off32[i + 1] - off32[i];
                local_rows += (size_t)off32[nb];
            }
        }
        if (world > 1) {
            CU(e, cudaMemcpyAsync(c->d_cnt_local[s], c->h_cnt_local[s], sizeof(int) * B, cudaMemcpyHostToDevice, c->stream));
            NC(e, c, c->api->AllGather(c->d_cnt_local[s], c->d_cnt_all[s], B, kNcclInt32, c->comm, c->stream));
            CU(e, cudaMemcpyAsync(c->h_cnt_all[s], c->d_cnt_all[s], sizeof(int) * B * world, cudaMemcpyDeviceToHost, c->stream));
        }
        CU(e, cudaEventRecord(c->ev_cnt[s], c->stream));
        if (x > 0 && (rc = exchange(x - 1))) return rc;
    }
    (void)submitted_next;
    chunk_row0[n_chunks] = local_rows;
    if (n_chunks > 0 && (rc = exchange(n_chunks - 1))) return rc;
    CU(e, cudaStreamSynchronize(c->stream));

    // destination: offsets, capacity, frame-order assembly (regions are already in rank order = frame order)
    float assemble_ms = 0.f;
    if (is_dst && first_error == SURF_OK) {
        offsets[0] = 0;
        for (long f = 0; f < n_frames_total; f++) offsets[f + 1] = offsets[f] + frame_counts[f];
        const int64_t total = offsets[n_frames_total];
        if (total > capacity) {
            e->required = (long)total;
            first_error = fail(e, SURF_ERR_CAPACITY, "%lld keypoints in the sweep; the output arrays hold %ld", (long long)total, capacity);
            snprintf(first_msg, sizeof(first_msg), "%s", e->err);
        } else {
            const cudaMemcpyKind kind = out_mem == SURF_MEM_DEVICE ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost;
            CU(e, cudaEventRecord(c->ev_t[2], e->stream));
            size_t row = 0;
            for (int r = 0; r < world; r++) {
                if (!got_rows[r]) continue;
                CU(e, cudaMemcpyAsync(keypoints + row, c->stage_kp + region_base[r], got_rows[r] * sizeof(surf_keypoint), kind, e->stream));
                if (want_descriptors)
                    CU(e, cudaMemcpyAsync(descriptors + row * D, c->stage_desc + region_base[r] * D, got_rows[r] * D * sizeof(float),
                                          kind, e->stream));
                row += got_rows[r];
            }
            CU(e, cudaEventRecord(c->ev_t[3], e->stream));
        }
    }
    CU(e, cudaEventRecord(c->ev_t[1], e->stream));
    CU(e, cudaStreamSynchronize(e->stream));
    if (is_dst && first_error == SURF_OK && n_frames_total > 0) cudaEventElapsedTime(&assemble_ms, c->ev_t[2], c->ev_t[3]);
    if (was_async) surf_set_async_outputs(e, 1);
    if (stats) {
        memset(stats, 0, sizeof(*stats));
        cudaEventElapsedTime(&stats->total_ms, c->ev_t[0], c->ev_t[1]);
        for (int i = 0; i < n_exchanges; i++) {
            float t = 0.f;
            cudaEventElapsedTime(&t, c->ev_x[2 * i], c->ev_x[2 * i + 1]);
            stats->exchange_ms += t;
        }
        stats->assemble_ms = assemble_ms;
        stats->bytes_sent = sent;
        stats->bytes_received = received;
        stats->chunks = n_chunks;
        stats->launches = launches;
    }
    cudaGetLastError();
    if (first_error != SURF_OK) {
        // leave no batch behind: whatever was submitted and not collected is collected into nowhere
        while (e->out[0].active || e->out[1].active)
            if (surf_collect_batch(e, nullptr, nullptr, SURF_MEM_DEVICE, 1L << 40, off32.data()) != SURF_OK) break;
        snprintf(e->err, sizeof(e->err), "%s", first_msg);
        return first_error;
    }
    return SURF_OK;
}

int surf_db_knn_sharded(surf_db *db, surf_comm *c, const float *query, int n_query, int k, int query_root, surf_dmatch *out,
                        int mem)
{
    if (!db) return SURF_ERR_BAD_ARG;
    NvtxRange nvtx("surf_db_knn_sharded");
    surf_engine *e = db->e;
    if (!c) return surf_db_knn(db, query, n_query, k, out, mem);  // one GPU: the whole database is local
    if (c->e != e) return fail(e, SURF_ERR_BAD_ARG, "database and communicator belong to different engines");
    if (n_query < 0) return fail(e, SURF_ERR_BAD_ARG, "negative row count");
    if (k < 1 || k > 4) return fail(e, SURF_ERR_UNSUPPORTED, "k must be in 1..4");
    if (query_root >= c->n_ranks) return fail(e, SURF_ERR_BAD_ARG, "query_root out of range");
    if (n_query == 0) return SURF_OK;
    if (!out || (!query && (query_root < 0 || query_root == c->rank || mem == SURF_MEM_DEVICE)))
        return fail(e, SURF_ERR_BAD_ARG, "null pointer");
    int rc = ensure_device(e);
    if (rc) return rc;
    e->launches = 0;
    const int world = c->n_ranks, D = db->dim;
    const size_t bq = ((size_t)n_query * D * sizeof(float) + 255) & ~(size_t)255;
    const size_t brec = ((size_t)n_query * k * sizeof(surf_dmatch) + 255) & ~(size_t)255;
    // scratch: [query rows (host calls) | local records | all ranks' records | merged records (host calls)]
    const size_t need = bq + brec + brec * world + brec;
    if (need > c->knn_cap) {
        CU(e, cudaStreamSynchronize(e->stream));
        cudaFree(c->d_knn);
        c->d_knn = nullptr;
        c->knn_cap = 0;
        CU(e, cudaMalloc(&c->d_knn, need));
        c->knn_cap = need;
    }
    char *p = (char *)c->d_knn;
    float *dq = (float *)p;
    surf_dmatch *d_local = (surf_dmatch *)(p + bq);
    surf_dmatch *d_all = (surf_dmatch *)(p + bq + brec);
    surf_dmatch *d_out = (surf_dmatch *)(p + bq + brec + brec * world);
    cudaStream_t st = e->stream;
    CU(e, cudaEventRecord(c->ev_t[0], st));
    if (mem == SURF_MEM_HOST) {
        if (query_root < 0 || query_root == c->rank)
            CU(e, cudaMemcpyAsync(dq, query, (size_t)n_query * D * sizeof(float), cudaMemcpyHostToDevice, st));
    } else {
        dq = const_cast<float *>(query);
        d_out = out;
    }
    CU(e, cudaEventRecord(c->ev_t[2], st));
    if (query_root >= 0) NC(e, c, c->api->Broadcast(dq, dq, (size_t)n_query * D, kNcclFloat32, query_root, c->comm, st));
    CU(e, cudaEventRecord(c->ev_t[3], st));
    // per-rank exact top-k against the local rows (tensor-core candidates + exact rescoring, as on one GPU)
    if ((rc = db_knn_device(db, dq, n_query, k, d_local, &e->launches, false))) return rc;
    c->h_rows[0] = (int)db_rows(db);
    CU(e, cudaMemcpyAsync(c->d_rows, c->h_rows, sizeof(int), cudaMemcpyHostToDevice, st));
    CU(e, cudaEventRecord(c->ev_t[4], st));
    NC(e, c, c->api->AllGather(c->d_rows, c->d_rows + 1, 1, kNcclInt32, c->comm, st));
    const long nrec = (long)n_query * k;
    k_add_row_base<<<(unsigned)((nrec + 255) / 256), 256, 0, st>>>(d_local, nrec, c->d_rows + 1, c->rank);
    e->launches++;
    // the whole exchange step: 16 B x k x Nq per rank, lists contiguous in rank order
    NC(e, c, c->api->AllGather(d_local, d_all, (size_t)nrec * sizeof(surf_dmatch), kNcclInt8, c->comm, st));
    CU(e, cudaEventRecord(c->ev_t[5], st));
    launch_merge_topk(d_all, world, n_query, k, d_out, st, &e->launches);
    CU(e, cudaGetLastError());
    if (mem == SURF_MEM_HOST) CU(e, cudaMemcpyAsync(out, d_out, (size_t)nrec * sizeof(surf_dmatch), cudaMemcpyDeviceToHost, st));
    CU(e, cudaEventRecord(c->ev_t[1], st));
    CU(e, cudaStreamSynchronize(st));
    float a = 0.f, b = 0.f;
    cudaEventElapsedTime(&c->last_total_ms, c->ev_t[0], c->ev_t[1]);
    cudaEventElapsedTime(&a, c->ev_t[2], c->ev_t[3]);
    cudaEventElapsedTime(&b, c->ev_t[4], c->ev_t[5]);
    c->last_nccl_ms = a + b;
    return SURF_OK;
}

}  // extern "C"
