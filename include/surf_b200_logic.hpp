// surf_b200_logic.hpp -- header-only C++ mirror of the part of vtkSlicerSurfFeaturesLogic that drives the hot
// path from files: the three sweeps (bogus / train / query) and the inter-slice correspondences, with the
// reference's own method names so that its headless driver's computeAll() (SurfFeatures/Logic/surfcmd.cxx:3-41)
// compiles against this class unchanged.
//
//   reference (SurfFeatures/Logic/vtkSlicerSurfFeaturesLogic.{h,cxx})        here
//   setBogusFile / setTrainFile / setQueryFile            .h:314-316          same
//   set/get{Bogus,Train,Query}{Start,Stop}Frame           .h:321-326          same
//   setMinHessian / getMinHessian                         .h:307-308          same
//   set{Left,Right,Top,Bottom}Crop                        .h:337-340          same
//   mask = cv::imread(hard-coded path)                    .cxx:1039-1044      setMask(data, width, height, step)
//   computeBogus / computeTrain / computeQuery            .cxx:1313-1335      open the file, reset the sweep
//   is{Bogus,Train,Query}Loading, isLoading               .cxx:1289-1311      same
//   computeNext{Bogus,Train,Query}                        .cxx:1408-1421      one BATCH of frames per call
//   startInterSliceCorrespondence                         .cxx:1423-1432      same
//   isCorrespondenceComputing                                                  same
//   computeNextInterSliceCorrespondence                   .cxx:1434-1571      all query frames in one call
//   bestMatches / bestMatchesCount / afterHoughMatches / matchesWithBestTrainImage   .h:258-262   getters
//
// Differences that are deliberate: a computeNext*() call advances by one engine batch instead of one frame (the
// progress getters still run 0..100); the descriptors stay on the device (train / bogus databases) unless asked
// for; simulateAll and the overlay drawing are not part of this library (ransac() and the pose estimate, row f4, are:
// see below).  The start / stop frame defaults are the reference constructor's (.cxx:1021-1026).
#ifndef SURF_B200_LOGIC_HPP
#define SURF_B200_LOGIC_HPP

#include <string>
#include <vector>

#include "surf_b200.hpp"

namespace surfb200 {

class SurfFeaturesLogic {
public:
    // max_width / max_height bound the CROPPED frame size; batch = frames per computeNext*() call
    explicit SurfFeaturesLogic(int device = 0, int max_width = 1024, int max_height = 1024, int batch = 32,
                               int max_keypoints = 16384, long db_rows = 1L << 21, int db_images = 8192)
        : minHessian(400), xcentroid(0), ycentroid(0), correspondenceProgress(0), ransacMargin(1.0f), e_(0), trainDb_(0), bogusDb_(0), rand_(0),
          maskData_(0), maskW_(0), maskH_(0), maskStep_(0), batch_(batch), dbRows_(db_rows), dbImages_(db_images)
    {
        cropRatios[0] = 1 / 5.; cropRatios[1] = 3. / 3.9; cropRatios[2] = 1 / 6.; cropRatios[3] = 3. / 4.;  // .cxx:1008-1009
        bogusStartFrame = trainStartFrame = 0;                                                             // .cxx:1021-1024
        bogusStopFrame = trainStopFrame = 2;
        queryStartFrame = 3; queryStopFrame = 5;                                                           // .cxx:1025-1026
        // the reference's effective parameters: detector (minHessian, 4, 2), default-constructed extractor
        surf_params p = {(double)minHessian, 4, 2, 1, 0};
        const int rc = surf_create(&p, device, max_width, max_height, batch, max_keypoints, &e_);
        if (rc != SURF_OK) throw Error(rc, "surf_create failed (an sm_100 GPU is required; there is no CPU path)");
        int rc2 = surf_db_create(e_, 128, db_rows, db_images, &trainDb_);
        if (rc2 == SURF_OK) rc2 = surf_db_create(e_, 128, db_rows, db_images, &bogusDb_);
        if (rc2 != SURF_OK) {  // a constructor that throws runs no destructor
            const std::string msg = surf_last_error(e_);
            surf_db_destroy(trainDb_);
            surf_destroy(e_);
            throw Error(rc2, msg);
        }
    }
    ~SurfFeaturesLogic()
    {
        for (int s = 0; s < 3; s++)
            if (sweep_[s].seq) surf_mha_close(sweep_[s].seq);
        surf_db_destroy(trainDb_);
        surf_db_destroy(bogusDb_);
        surf_rand_destroy(rand_);
        surf_destroy(e_);
    }

    void setBogusFile(const std::string &f) { sweep_[BOGUS].file = f; }
    void setTrainFile(const std::string &f) { sweep_[TRAIN].file = f; }
    void setQueryFile(const std::string &f) { sweep_[QUERY].file = f; }
    void setMinHessian(int h) { minHessian = h; }
    int getMinHessian() const { return minHessian; }
    void setLeftCrop(double v) { cropRatios[0] = (float)v; }
    void setRightCrop(double v) { cropRatios[1] = (float)v; }
    void setTopCrop(double v) { cropRatios[2] = (float)v; }
    void setBottomCrop(double v) { cropRatios[3] = (float)v; }
#define SURF_B200_GETSET(type, var, Name) \
    void set##Name(type v) { var = v; }   \
    type get##Name() const { return var; }
    SURF_B200_GETSET(int, trainStartFrame, TrainStartFrame)
    SURF_B200_GETSET(int, trainStopFrame, TrainStopFrame)
    SURF_B200_GETSET(int, bogusStartFrame, BogusStartFrame)
    SURF_B200_GETSET(int, bogusStopFrame, BogusStopFrame)
    SURF_B200_GETSET(int, queryStartFrame, QueryStartFrame)
    SURF_B200_GETSET(int, queryStopFrame, QueryStopFrame)
#undef SURF_B200_GETSET
    // full-size mask (same size as the uncropped frames); the pointer must stay valid
    void setMask(const unsigned char *data, int width, int height, size_t step)
    {
        maskData_ = data; maskW_ = width; maskH_ = height; maskStep_ = step ? step : (size_t)width;
    }

    void computeBogus() { begin(BOGUS, bogusStartFrame, bogusStopFrame); }
    void computeTrain() { begin(TRAIN, trainStartFrame, trainStopFrame); }
    void computeQuery() { begin(QUERY, queryStartFrame, queryStopFrame); }
    bool isBogusLoading() const { return loading(BOGUS); }
    bool isTrainLoading() const { return loading(TRAIN); }
    bool isQueryLoading() const { return loading(QUERY); }
    bool isLoading() const { return isQueryLoading() || isBogusLoading() || isTrainLoading(); }
    void computeNextBogus() { next(BOGUS); }
    void computeNextTrain() { next(TRAIN); }
    void computeNextQuery() { next(QUERY); }
    int getBogusProgress() const { return progress(BOGUS); }
    int getTrainProgress() const { return progress(TRAIN); }
    int getQueryProgress() const { return progress(QUERY); }

    void startInterSliceCorrespondence()
    {
        bestMatches.clear();
        bestMatchesCount.clear();
        matchesWithBestTrainImage.clear();
        afterHoughMatches.clear();
        houghCounts.clear();
        correspondenceProgress = 0;
    }
    bool isCorrespondenceComputing() const { return correspondenceProgress != 100 && sweep_[QUERY].done > 0; }
    void computeNextInterSliceCorrespondence()
    {
        const Sweep &q = sweep_[QUERY];
        if (q.done == 0 || sweep_[BOGUS].done == 0 || sweep_[TRAIN].done == 0) return;   // .cxx:1436-1437
        if (isLoading()) return;                                                          // .cxx:1439-1440
        if (afterHoughMatches.size() == (size_t)q.done) return;
        const int nq = q.done;
        std::vector<DMatch> m(q.offsets[nq] > 0 ? q.offsets[nq] : 1);
        std::vector<surf_vote_result> r(nq);
        surf_vote_params vp = {0.95f, xcentroid, ycentroid};                              // fRatioThreshold .cxx:1463
        check(surf_correspond(trainDb_, bogusDb_, q.desc.data(), q.kps.data(), q.offsets.data(), nq, &vp, m.data(), r.data(),
                              SURF_MEM_HOST));
        for (int i = 0; i < nq; i++) {
            const DMatch *b = m.data() + q.offsets[i];
            afterHoughMatches.push_back(std::vector<DMatch>(b, b + r[i].n_matches));
            std::vector<DMatch> with;                                                     // .cxx:1551-1557
            for (int k = 0; k < r[i].n_matches; k++)
                if (b[k].img_idx == r[i].best_image) with.push_back(b[k]);
            matchesWithBestTrainImage.push_back(with);
            bestMatches.push_back(r[i].best_image);
            bestMatchesCount.push_back(r[i].best_count);
            houghCounts.push_back(r[i].hough_count);
        }
        correspondenceProgress = 100;
    }

    // ---- plane fit of the matched train keypoints (row f4) ----
    struct Point3 { double v[3]; };   // vnl_double_3
    void setRansacMargin(float m) { ransacMargin = m; }                                   // .h setter, default 1.0 (.cxx:1002)
    float getRansacMargin() const { return ransacMargin; }
    // ransac(points, inliersIdx, planePointsIdx) .cxx:2085-2188: 0 on success, 1 with fewer than three points.
    // rand() is the logic's own stream: glibc's generator, never seeded, as in the reference's process.
    int ransac(const std::vector<Point3> &points, std::vector<int> &inliersIdx, std::vector<int> &planePointsIdx)
    {
        inliersIdx.clear();
        planePointsIdx.clear();
        if (points.size() < 3) return 1;
        if (!rand_) check(surf_rand_create(1, &rand_));
        std::vector<int32_t> inl(points.size() + 3);
        int32_t n = 0, plane[3];
        check(surf_ransac_plane(e_, rand_, &points[0].v[0], (int)points.size(), ransacMargin, 1000, inl.data(), &n, plane, 0, 0));
        inliersIdx.assign(inl.begin(), inl.begin() + n);
        planePointsIdx.assign(plane, plane + 3);
        return 0;
    }

    // results (public members in the reference, .h:258-262)
    std::vector<int> bestMatches, bestMatchesCount, houghCounts;
    std::vector<std::vector<DMatch> > afterHoughMatches, matchesWithBestTrainImage;
    const std::vector<KeyPoint> &queryKeypoints() const { return sweep_[QUERY].kps; }
    const std::vector<int32_t> &queryOffsets() const { return sweep_[QUERY].offsets; }
    long trainRows() const { long r = 0; surf_db_size(trainDb_, &r, 0); return r; }
    surf_engine *engine() const { return e_; }

    int minHessian;
    float cropRatios[4];
    int trainStartFrame, trainStopFrame, bogusStartFrame, bogusStopFrame, queryStartFrame, queryStopFrame;
    int xcentroid, ycentroid;
    int correspondenceProgress;
    float ransacMargin;

private:
    enum { BOGUS = 0, TRAIN = 1, QUERY = 2 };
    struct Sweep {
        std::string file;
        surf_mha *seq;
        surf_mha_info info;
        int done;                      // frames processed so far
        std::vector<KeyPoint> kps;     // query sweep only: host copies for surf_correspond
        std::vector<float> desc;
        std::vector<int32_t> offsets;
        std::vector<unsigned char> stage;
        int32_t rect[4];               // cropData's rectangle for THIS file's frame size (cv::Rect x, y, w, h)
        Sweep() : seq(0), info(), done(0) { rect[0] = rect[1] = rect[2] = rect[3] = 0; }
    };

    void check(int rc) const
    {
        if (rc != SURF_OK) throw Error(rc, e_ ? surf_last_error(e_) : "surf_b200 call failed");
    }
    bool loading(int s) const { return sweep_[s].seq && sweep_[s].done < sweep_[s].info.n_frames; }
    int progress(int s) const
    {
        return sweep_[s].seq && sweep_[s].info.n_frames ? (sweep_[s].done * 100) / sweep_[s].info.n_frames : 0;
    }
    // readAndComputeFeaturesOnMhaFile (.cxx:1337-1379): read the header, crop the mask, compute its centroid
    void begin(int s, int &first, int &last)
    {
        Sweep &w = sweep_[s];
        correspondenceProgress = 0;
        if (w.seq) surf_mha_close(w.seq);
        w.seq = 0;
        w.done = 0;
        w.kps.clear(); w.desc.clear(); w.offsets.assign(1, 0);
        if (!(cropRatios[0] < cropRatios[1] || cropRatios[2] < cropRatios[3])) return;       // cropRatiosValid
        if (w.file.find(".mha") == std::string::npos) return;
        if (surf_mha_open(w.file.c_str(), first, last, 0, &w.seq, &w.info) != SURF_OK) {
            w.seq = 0;
            return;
        }
        first = w.info.first_frame;                                                          // "may have been modified"
        last = w.info.last_frame;
        surf_params p = {(double)minHessian, 4, 2, 1, 0};
        check(surf_set_params(e_, &p));
        check(surf_crop_rect(w.info.width, w.info.height, cropRatios, w.rect));   // every file is cropped by its own size
        if (maskData_) {
            if (maskW_ != w.info.width || maskH_ != w.info.height) throw Error(SURF_ERR_BAD_ARG, "mask size differs from the frames");
            check(surf_mask_centroid(maskData_ + (size_t)w.rect[1] * maskStep_ + w.rect[0], w.rect[2], w.rect[3], maskStep_, &xcentroid,
                                     &ycentroid));
        }
        if (s == TRAIN) check(surf_db_clear(trainDb_));
        if (s == BOGUS) check(surf_db_clear(bogusDb_));
        w.stage.resize((size_t)batch_ * w.info.width * w.info.height);
    }
    // computeNext (.cxx:1392-1406) for the next batch of frames; the matcher is filled as the frames complete
    void next(int s)
    {
        Sweep &w = sweep_[s];
        if (!loading(s)) return;
        const int cnt = w.info.n_frames - w.done < batch_ ? w.info.n_frames - w.done : batch_;
        const size_t fb = (size_t)w.info.width * w.info.height;
        check(surf_mha_read(w.seq, w.done, cnt, w.stage.data(), fb));
        const int32_t *rect = w.rect;
        const unsigned char *view = w.stage.data() + (size_t)rect[1] * w.info.width + rect[0];
        const unsigned char *cmask = maskData_ ? maskData_ + (size_t)rect[1] * maskStep_ + rect[0] : 0;
        check(surf_submit_batch(e_, view, SURF_MEM_HOST, cnt, rect[2], rect[3], (size_t)w.info.width, fb, cmask, maskStep_, 1));
        std::vector<int32_t> off(cnt + 1);
        if (s == QUERY) {
            // counts first (a collect without outputs), then a second collect into the grown host arrays
            check(surf_collect_batch(e_, 0, 0, SURF_MEM_DEVICE, 1L << 40, off.data()));
            const size_t base = w.kps.size(), D = (size_t)surf_descriptor_size(e_);
            w.kps.resize(base + off[cnt]);
            w.desc.resize((base + off[cnt]) * D);
            if (off[cnt])
                check(surf_collect_batch(e_, w.kps.data() + base, w.desc.data() + base * D, SURF_MEM_HOST, off[cnt], off.data()));
            for (int b = 0; b < cnt; b++) w.offsets.push_back((int32_t)(base + off[b + 1]));
        } else {
            check(surf_collect_batch(e_, 0, 0, SURF_MEM_DEVICE, 1L << 40, off.data()));
            const surf_keypoint *dk = 0;
            const float *dd = 0;
            check(surf_device_results(e_, &dk, &dd, 0));
            check(surf_db_add(s == TRAIN ? trainDb_ : bogusDb_, dd, dk, off.data(), cnt, SURF_MEM_DEVICE));
        }
        w.done += cnt;
    }

    surf_engine *e_;
    surf_db *trainDb_, *bogusDb_;
    surf_rand *rand_;
    const unsigned char *maskData_;
    int maskW_, maskH_;
    size_t maskStep_;
    int batch_;
    long dbRows_;
    int dbImages_;
    Sweep sweep_[3];
};

}  // namespace surfb200

#endif  // SURF_B200_LOGIC_HPP
