// surf_mha.cu -- "next" row f1 of SURVEY.md section 8: the on-disk format on the input side of the hot path.
// Host code only (no kernels): MetaImage (.mha) sequence header + raw 8-bit frames as written by PLUS, the
// crop rectangle, the mask centroid, and the sweep driver that feeds whole files to the engine in batches.
//
// Reference (SurfFeatures/Logic/vtkSlicerSurfFeaturesLogic.cxx):
//   readImageDimensions_mha :444-476, readImageTransforms_mha :478-541, readImages_mha :543-595, read_mha :598-632,
//   cropData(cv::Mat&) :745-752, computeCentroid :839-860, readAndComputeFeaturesOnMhaFile :1337-1379,
//   computeNext :1392-1406.
// Behaviours kept on purpose (each is what the reference does on the same file):
//   * header parsing stops at the first empty line or at "ElementDataFile = LOCAL";
//   * every "...ProbeToTrackerTransform =" AND "...UltrasoundToTrackerTransform =" line yields a transform;
//   * first / last frame are clamped as in readImages_mha;
//   * a frame whose tracking status is INVALID is not read AND its bytes are not skipped, so every later frame
//     comes from one frame earlier in the file (readImages_mha has no seek on that branch).  Flag
//     SURF_MHA_SKIP_INVALID_BYTES selects the evidently intended behaviour instead.
#include <cerrno>
#include <unistd.h>
#include <cstdlib>
#include <cstring>
#include <future>
#include <new>
#include <string>

#include "surf_engine_priv.h"

struct surf_mha {
    FILE *fp = nullptr;
    std::string path, dir;
    int width = 0, height = 0, count = 0;
    int first = 0, last = 0;
    long data_offset = 0;
    int flags = 0;
    std::vector<float> transforms;       // kept transforms, 12 floats each
    std::vector<std::string> names;      // kept "<dir><frame field name>.png"
    std::vector<uint8_t> validity;       // flags of frames first..last
    std::vector<long> frame_offset;      // file offset of every kept frame
    uint8_t *stage[3] = {nullptr, nullptr, nullptr};  // pinned staging buffers of the sweep driver
    size_t stage_bytes = 0;
    char err[256] = "";
};

namespace {

int mha_fail(surf_mha *m, int code, const char *msg)
{
    if (m) snprintf(m->err, sizeof(m->err), "%s", msg);
    return code;
}

// getDir (:410-422): everything up to and including the last '/'
std::string dir_of(const std::string &f)
{
    const size_t pos = f.rfind('/');
    return pos == std::string::npos ? std::string() : f.substr(0, pos) + '/';
}

// one header line without its '\n' (std::getline semantics); false at end of file
bool next_line(FILE *fp, std::string &line)
{
    line.clear();
    int c;
    bool any = false;
    while ((c = fgetc(fp)) != EOF) {
        any = true;
        if (c == '\n') return true;
        line.push_back((char)c);
    }
    return any;
}

}  // namespace

extern "C" {

int surf_mha_open(const char *path, int first_frame, int last_frame, int flags, surf_mha **out, surf_mha_info *info)
{
    if (!path || !out) return SURF_ERR_BAD_ARG;
    *out = nullptr;
    surf_mha *m = new (std::nothrow) surf_mha();
    if (!m) return SURF_ERR_CUDA;
    m->path = path;
    m->dir = dir_of(m->path);
    m->flags = flags;
    m->fp = fopen(path, "rb");
    if (!m->fp) {
        delete m;
        return SURF_ERR_BAD_ARG;
    }
    // ---- header: dimensions, transforms, validity (one pass; the reference makes three with the same rules) ----
    std::vector<float> all_transforms;
    std::vector<std::string> all_names;
    std::vector<uint8_t> all_valid;
    bool have_dims = false, have_data = false;
    std::string line;
    while (next_line(m->fp, line)) {
        if (line.empty()) break;
        const char *pch = line.c_str();
        if (!have_dims && strstr(pch, "DimSize =")) {
            if (sscanf(pch, "DimSize = %d %d %d", &m->width, &m->height, &m->count) != 3) break;
            have_dims = true;
        }
        if (strstr(pch, "ProbeToTrackerTransform =") || strstr(pch, "UltrasoundToTrackerTransform =")) {
            // "<name> = f f f ..." : the name ends one character before '=', then 12 numbers each found after a space
            const char *eq = strchr(pch, '=');
            std::string name(pch, eq > pch ? (size_t)(eq - 1 - pch) : 0);
            float v[12];
            const char *q = eq;
            bool ok = true;
            for (int j = 0; j < 12; j++) {
                q = strchr(q + 1, ' ');
                if (!q) {
                    ok = false;
                    break;
                }
                v[j] = (float)atof(q);
                q++;
            }
            if (!ok) break;  // the reference returns from the parser here, keeping what it has
            all_transforms.insert(all_transforms.end(), v, v + 12);
            all_names.push_back(m->dir + name + ".png");
        } else if (strstr(pch, "UltrasoundToTrackerTransformStatus") || strstr(pch, "ProbeToTrackerTransformStatus")) {
            if (strstr(pch, "OK")) all_valid.push_back(1);
            else if (strstr(pch, "INVALID")) all_valid.push_back(0);
        }
        if (strstr(pch, "ElementDataFile = LOCAL")) {
            have_data = true;
            m->data_offset = ftell(m->fp);
            break;
        }
    }
    if (!have_dims || m->width <= 0 || m->height <= 0 || m->count <= 0) {
        surf_mha_close(m);
        return SURF_ERR_BAD_ARG;
    }
    if (!have_data) {
        // readImages_mha keeps scanning (fgets) for the marker even past an empty line
        while (next_line(m->fp, line))
            if (strstr(line.c_str(), "ElementDataFile = LOCAL")) {
                have_data = true;
                m->data_offset = ftell(m->fp);
                break;
            }
        if (!have_data) {
            surf_mha_close(m);
            return SURF_ERR_BAD_ARG;
        }
    }
    // ---- frame range as readImages_mha clamps it (:562-569) ----
    int first = first_frame, last = last_frame;
    if (first < 0) first = 0;
    if (first >= m->count) first = m->count - 1;
    if (last < 0 || last >= m->count) last = m->count - 1;
    if (first > last) first = last;
    m->first = first;
    m->last = last;
    if ((int)all_valid.size() < last + 1) {  // the reference indexes transformsValidity[i] for every frame in range
        surf_mha_close(m);
        return SURF_ERR_BAD_ARG;
    }
    const long frame_bytes = (long)m->width * m->height;
    long pos = m->data_offset + (long)first * frame_bytes;  // frames before `first` are skipped with fseek
    for (int i = first; i <= last; i++) {
        if (all_valid[i]) {
            m->frame_offset.push_back((flags & SURF_MHA_SKIP_INVALID_BYTES) ? m->data_offset + (long)i * frame_bytes : pos);
            pos += frame_bytes;
        }
    }
    // ---- read_mha (:613-624): transforms / names of the kept frames, validity sliced to [first, last] ----
    for (size_t i = 0; i < all_valid.size(); i++) {
        if ((int)i < first || (int)i > last || !all_valid[i]) continue;
        if (i < all_names.size()) {
            m->transforms.insert(m->transforms.end(), all_transforms.begin() + 12 * i, all_transforms.begin() + 12 * i + 12);
            m->names.push_back(all_names[i]);
        }
    }
    m->validity.assign(all_valid.begin() + first, all_valid.begin() + last + 1);
    if (info) {
        info->width = m->width;
        info->height = m->height;
        info->frames_in_file = m->count;
        info->first_frame = first;
        info->last_frame = last;
        info->n_frames = (int)m->frame_offset.size();
        info->n_transforms = (int)m->names.size();
    }
    *out = m;
    return SURF_OK;
}

void surf_mha_close(surf_mha *m)
{
    if (!m) return;
    if (m->fp) fclose(m->fp);
    for (auto &s : m->stage)
        if (s) cudaFreeHost(s);
    delete m;
}

const char *surf_mha_last_error(const surf_mha *m) { return m ? m->err : "null sequence"; }

int surf_mha_transforms(const surf_mha *m, float *transforms, int capacity)
{
    if (!m || !transforms) return SURF_ERR_BAD_ARG;
    if ((size_t)capacity * 12 < m->transforms.size()) return SURF_ERR_CAPACITY;
    if (!m->transforms.empty()) memcpy(transforms, m->transforms.data(), m->transforms.size() * sizeof(float));
    return SURF_OK;
}

int surf_mha_validity(const surf_mha *m, uint8_t *valid, int capacity)
{
    if (!m || !valid) return SURF_ERR_BAD_ARG;
    if ((size_t)capacity < m->validity.size()) return SURF_ERR_CAPACITY;
    if (!m->validity.empty()) memcpy(valid, m->validity.data(), m->validity.size());
    return SURF_OK;
}

int surf_mha_frame_name(const surf_mha *m, int i, char *buf, int capacity)
{
    if (!m || !buf || i < 0 || i >= (int)m->names.size()) return SURF_ERR_BAD_ARG;
    if ((int)m->names[i].size() + 1 > capacity) return SURF_ERR_CAPACITY;
    memcpy(buf, m->names[i].c_str(), m->names[i].size() + 1);
    return SURF_OK;
}

int surf_mha_read(surf_mha *m, int start, int count, uint8_t *dst, size_t frame_stride)
{
    if (!m || !dst || start < 0 || count < 0 || start + count > (int)m->frame_offset.size()) return SURF_ERR_BAD_ARG;
    const size_t frame_bytes = (size_t)m->width * m->height;
    if (frame_stride < frame_bytes) return mha_fail(m, SURF_ERR_BAD_ARG, "frame stride smaller than a frame");
    for (int i = 0; i < count; i++) {
        if (fseek(m->fp, m->frame_offset[start + i], SEEK_SET) != 0) return mha_fail(m, SURF_ERR_BAD_ARG, "seek failed");
        const size_t got = fread(dst + (size_t)i * frame_stride, 1, frame_bytes, m->fp);
        // a short file leaves the tail of the frame as it was, like the reference's unchecked fread
        if (got < frame_bytes) memset(dst + (size_t)i * frame_stride + got, 0, frame_bytes - got);
    }
    return SURF_OK;
}

int surf_crop_rect(int width, int height, const float crop_ratios[4], int32_t rect[4])
{
    if (!crop_ratios || !rect || width <= 0 || height <= 0) return SURF_ERR_BAD_ARG;
    // cropRatiosValid (:2198-2201) -- note the reference's `||`
    if (!(crop_ratios[0] < crop_ratios[1] || crop_ratios[2] < crop_ratios[3])) return SURF_ERR_BAD_ARG;
    // cv::Rect(int) from float expressions: int * float in fp32, truncated toward zero (:747-749)
    const volatile float x = (float)width * crop_ratios[0], y = (float)height * crop_ratios[2];
    const volatile float dw = crop_ratios[1] - crop_ratios[0], dh = crop_ratios[3] - crop_ratios[2];
    const volatile float w = (float)width * dw, h = (float)height * dh;
    rect[0] = (int32_t)x;
    rect[1] = (int32_t)y;
    rect[2] = (int32_t)w;
    rect[3] = (int32_t)h;
    // img(roi) asserts the rectangle lies inside the image
    if (rect[0] < 0 || rect[1] < 0 || rect[2] <= 0 || rect[3] <= 0 || rect[0] + rect[2] > width || rect[1] + rect[3] > height)
        return SURF_ERR_BAD_ARG;
    return SURF_OK;
}

int surf_mask_centroid(const uint8_t *mask, int width, int height, size_t pitch, int32_t *x, int32_t *y)
{
    if (!mask || !x || !y || width <= 0 || height <= 0 || pitch < (size_t)width) return SURF_ERR_BAD_ARG;
    // computeCentroid (:839-860): fp32 running sums in row-major order
    volatile float rowc = 0, colc = 0;
    int count = 0;
    for (int i = 0; i < height; i++)
        for (int j = 0; j < width; j++)
            if (mask[(size_t)i * pitch + j] > 0) {
                rowc = rowc + (float)i;
                colc = colc + (float)j;
                count += 1;
            }
    rowc = rowc / (float)count;
    colc = colc / (float)count;
    *y = (int32_t)rowc;
    *x = (int32_t)colc;
    return SURF_OK;
}

int surf_sweep_mha(surf_engine *e, surf_mha *m, const float crop_ratios[4], const uint8_t *mask, size_t mask_pitch,
                   surf_db *db, surf_keypoint *keypoints, float *descriptors, int32_t *offsets, long capacity,
                   int32_t *n_frames_out)
{
    if (!e) return SURF_ERR_BAD_ARG;
    if (!m || !offsets) return fail(e, SURF_ERR_BAD_ARG, "null sequence or offsets");
    int32_t rect[4] = {0, 0, m->width, m->height};
    if (crop_ratios && surf_crop_rect(m->width, m->height, crop_ratios, rect) != SURF_OK)
        return fail(e, SURF_ERR_BAD_ARG, "crop ratios do not give a rectangle inside the %dx%d frame", m->width, m->height);
    if (mask && mask_pitch < (size_t)m->width) return fail(e, SURF_ERR_BAD_ARG, "mask must be as large as the frames");
    const int n = (int)m->frame_offset.size(), B = e->max_b, D = surf_descriptor_size(e);
    const size_t frame_bytes = (size_t)m->width * m->height;
    // only the rows of the crop rectangle are read from the file (full width: one contiguous run per frame)
    const size_t band_bytes = (size_t)m->width * rect[3];
    const long band_off = (long)rect[1] * m->width;
    int rc = ensure_device(e);
    if (rc) return rc;
    if (m->stage_bytes < frame_bytes * B) {
        for (auto &s : m->stage) {
            if (s) cudaFreeHost(s);
            s = nullptr;
            CU(e, cudaHostAlloc((void **)&s, frame_bytes * B, cudaHostAllocDefault));
        }
        m->stage_bytes = frame_bytes * B;
    }
    // croppedMask = mask(roi) (:1368-1371): the same rectangle of the full-size mask
    const uint8_t *cmask = mask ? mask + (size_t)rect[1] * mask_pitch + rect[0] : nullptr;
    offsets[0] = 0;
    if (n_frames_out) *n_frames_out = n;
    long total = 0;
    // kReaders threads pread() interleaved frames of a chunk (page-cache reads are memcpy-bound per thread)
    constexpr int kReaders = 4;
    const int fd = fileno(m->fp);
    auto read_part = [m, fd, band_bytes, band_off](int start, int count, int part, uint8_t *dst) {
        for (int i = part; i < count; i += kReaders) {
            uint8_t *d = dst + (size_t)i * band_bytes;
            size_t got = 0;
            while (got < band_bytes) {
                const ssize_t r = pread(fd, d + got, band_bytes - got, m->frame_offset[start + i] + band_off + (long)got);
                if (r < 0 && errno == EINTR) continue;
                if (r <= 0) break;
                got += (size_t)r;
            }
            // a short file leaves the rest of the frame zero (the reference's fread is unchecked)
            if (got < band_bytes) memset(d + got, 0, band_bytes - got);
        }
        return 0;
    };
    std::vector<std::future<int> > pending;
    auto start_chunk = [&](int start, int count, uint8_t *dst) {
        pending.clear();
        for (int t = 0; t < kReaders; t++) pending.push_back(std::async(std::launch::async, read_part, start, count, t, dst));
    };
    auto wait_chunk = [&]() {
        for (auto &f : pending)
            if (f.valid()) f.get();
        pending.clear();
    };
    // Three pinned buffers: while chunk c is in its descriptor stage, chunk c+1 (already read) is uploaded and run
    // through stages 1-3 by surf_lookahead_batch, and chunk c+2 is read from the file.
    auto chunk_count = [&](int c) {
        const long s0 = (long)c * B;
        return s0 >= n ? 0 : (int)(n - s0 < B ? n - s0 : B);
    };
    auto chunk_view = [&](int c) { return m->stage[c % 3] + rect[0]; };  // cropData: the roi of every frame (rows already selected)
    if (n > 0) {
        start_chunk(0, chunk_count(0), m->stage[0]);
        wait_chunk();
        if (chunk_count(1)) start_chunk(B, chunk_count(1), m->stage[1]);
    }
    std::vector<int32_t> off((size_t)B + 1);
    for (int start = 0, c = 0; start < n; start += B, c++) {
        const int cnt = chunk_count(c);
        const uint8_t *view = chunk_view(c);
        rc = surf_submit_batch(e, view, SURF_MEM_HOST, cnt, rect[2], rect[3], (size_t)m->width, band_bytes, cmask, mask_pitch, 1);
        if (!rc && chunk_count(c + 1)) {
            wait_chunk();  // chunk c+1 is in its buffer
            rc = surf_lookahead_batch(e, chunk_view(c + 1), SURF_MEM_HOST, chunk_count(c + 1), rect[2], rect[3], (size_t)m->width,
                                      band_bytes, cmask, mask_pitch);
            // chunk c-1's buffer is free: its upload ended before its results were collected
            if (!rc && chunk_count(c + 2)) start_chunk((c + 2) * B, chunk_count(c + 2), m->stage[(c + 2) % 3]);
        }
        if (!rc) {
            if (keypoints || descriptors)
                rc = surf_collect_batch(e, keypoints ? keypoints + total : nullptr, descriptors ? descriptors + (size_t)total * D : nullptr,
                                        SURF_MEM_HOST, capacity - total, off.data());
            else
                rc = surf_collect_batch(e, nullptr, nullptr, SURF_MEM_DEVICE, 1L << 40, off.data());
        }
        if (!rc && db) {  // computeNext's matcher->add(descriptors) (:1402-1403), straight from the device results
            const surf_keypoint *dk = nullptr;
            const float *dd = nullptr;
            rc = surf_device_results(e, &dk, &dd, nullptr);
            if (!rc) rc = surf_db_add(db, dd, dk, off.data(), cnt, SURF_MEM_DEVICE);
        }
        if (rc) {
            wait_chunk();
            return rc;
        }
        for (int b = 0; b < cnt; b++) offsets[start + b + 1] = (int32_t)(total + off[b + 1]);
        total += off[cnt];
    }
    return SURF_OK;
}

}  // extern "C"
