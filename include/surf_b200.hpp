// surf_b200.hpp -- header-only C++ host mirror of the OpenCV interface the reference calls, over the
// C ABI in surf_b200.h.
//
//   surfb200::SURF         mirrors cv::SURF (nonfree/features2d.hpp, OpenCV 2.4.5): ctor
//                          (hessianThreshold, nOctaves, nOctaveLayers, extended, upright), detect,
//                          compute, operator(), descriptorSize/Type, the five public fields.
//                          Used where the reference builds cv::SurfFeatureDetector /
//                          cv::SurfDescriptorExtractor: vtkSlicerSurfFeaturesLogic.cxx:1384-1389.
//   surfb200::BFMatcherL2  mirrors DescriptorMatcher add/clear/knnMatch for exact L2:
//                          vtkSlicerSurfFeaturesLogic.cxx:640, :1402-1403.
//   surfb200::CvSURF       (only with -DSURF_B200_WITH_OPENCV) the same engine as a
//                          cv::Feature2D subclass.  OpenCV headers do not exist in the build
//                          container, so this part has never been compiled here; it is the binding
//                          INTEGRATION.md describes.
//
// Errors: OpenCV throws cv::Exception from CV_Assert; these wrappers throw surfb200::Error carrying
// the SURF_* status and surf_last_error().  The C ABI underneath never throws.
#ifndef SURF_B200_HPP
#define SURF_B200_HPP

#include <cfloat>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

#include "surf_b200.h"

namespace surfb200 {

struct Error : std::runtime_error {
    int code;
    Error(int c, const std::string &m) : std::runtime_error(m), code(c) {}
};

typedef surf_keypoint KeyPoint;  // same 28-byte layout as cv::KeyPoint
typedef surf_dmatch DMatch;      // same 16-byte layout as cv::DMatch

// Minimal view of an 8-bit single-channel image (cv::Mat data/cols/rows/step).
struct ImageView {
    const unsigned char *data;
    int cols, rows;
    size_t step;
    ImageView() : data(0), cols(0), rows(0), step(0) {}
    ImageView(const unsigned char *d, int c, int r, size_t s) : data(d), cols(c), rows(r), step(s ? s : (size_t)c) {}
    bool empty() const { return !data || cols <= 0 || rows <= 0; }
};

class SURF {
public:
    double hessianThreshold;
    int nOctaves, nOctaveLayers;
    bool extended, upright;

    explicit SURF(double hessianThreshold_ = 100, int nOctaves_ = 4, int nOctaveLayers_ = 2, bool extended_ = true,
                  bool upright_ = false, int device = 0, int max_width = 1024, int max_height = 1024,
                  int max_batch = 1, int max_keypoints = 32768)
        : hessianThreshold(hessianThreshold_), nOctaves(nOctaves_), nOctaveLayers(nOctaveLayers_),
          extended(extended_), upright(upright_), e_(0), cap_(max_keypoints)
    {
        surf_params p = params();
        int rc = surf_create(&p, device, max_width, max_height, max_batch, max_keypoints, &e_);
        if (rc != SURF_OK) throw Error(rc, "surf_create failed (an sm_100 GPU is required; there is no CPU path)");
    }
    ~SURF() { surf_destroy(e_); }
    SURF(const SURF &) = delete;
    SURF &operator=(const SURF &) = delete;

    int descriptorSize() const { return extended ? 128 : 64; }
    int descriptorType() const { return 5; /* CV_32F */ }
    surf_engine *engine() const { return e_; }

    // FeatureDetector::detect(image, keypoints, mask)
    void detect(const ImageView &image, std::vector<KeyPoint> &keypoints, const ImageView &mask = ImageView())
    {
        keypoints.clear();
        if (image.empty()) return;
        sync_params();
        keypoints.resize(cap_);
        int n = 0;
        check(surf_detect(e_, image.data, image.cols, image.rows, image.step, mask.data, mask.step, keypoints.data(),
                          cap_, &n));
        keypoints.resize(n);
    }

    // DescriptorExtractor::compute(image, keypoints /*in-out*/, descriptors): N x descriptorSize() floats
    void compute(const ImageView &image, std::vector<KeyPoint> &keypoints, std::vector<float> &descriptors)
    {
        descriptors.clear();
        if (image.empty() || keypoints.empty()) {
            keypoints.clear();
            return;
        }
        sync_params();
        descriptors.resize(keypoints.size() * (size_t)descriptorSize());
        int n = 0;
        check(surf_compute(e_, image.data, image.cols, image.rows, image.step, keypoints.data(), (int)keypoints.size(),
                           descriptors.data(), &n));
        keypoints.resize(n);
        descriptors.resize((size_t)n * descriptorSize());
    }

    // SURF::operator()(img, mask, keypoints, descriptors, useProvidedKeypoints)
    void operator()(const ImageView &image, const ImageView &mask, std::vector<KeyPoint> &keypoints,
                    std::vector<float> &descriptors, bool useProvidedKeypoints = false)
    {
        if (useProvidedKeypoints) {
            compute(image, keypoints, descriptors);
            return;
        }
        keypoints.clear();
        descriptors.clear();
        if (image.empty()) return;
        sync_params();
        keypoints.resize(cap_);
        descriptors.resize((size_t)cap_ * descriptorSize());
        int n = 0;
        check(surf_detect_and_compute(e_, image.data, image.cols, image.rows, image.step, mask.data, mask.step,
                                      keypoints.data(), descriptors.data(), cap_, &n));
        keypoints.resize(n);
        descriptors.resize((size_t)n * descriptorSize());
    }

    void check(int rc) const
    {
        if (rc != SURF_OK) throw Error(rc, surf_last_error(e_));
    }

private:
    surf_params params() const
    {
        surf_params p;
        p.hessian_threshold = hessianThreshold;
        p.n_octaves = nOctaves;
        p.n_octave_layers = nOctaveLayers;
        p.extended = extended ? 1 : 0;
        p.upright = upright ? 1 : 0;
        return p;
    }
    void sync_params()
    {
        // the five fields are public and mutable in cv::SURF; push them before every call
        surf_params want = params(), have;
        surf_get_params(e_, &have);
        if (std::memcmp(&want, &have, sizeof(want)) != 0) check(surf_set_params(e_, &want));
    }
    surf_engine *e_;
    int cap_;
};

// Exact brute-force L2 matcher with the DescriptorMatcher bookkeeping (add / clear / imgIdx).
class BFMatcherL2 {
public:
    explicit BFMatcherL2(surf_engine *engine, int dim = 64) : e_(engine), dim_(dim) {}
    void clear()
    {
        train_.clear();
        starts_.assign(1, 0);
    }
    void add(const float *descriptors, int rows)  // one image's descriptors (a cv::Mat of the vector<Mat>)
    {
        if (starts_.empty()) starts_.push_back(0);
        train_.insert(train_.end(), descriptors, descriptors + (size_t)rows * dim_);
        starts_.push_back(starts_.back() + rows);
    }
    // matches[i] holds up to k DMatch records for query row i, ascending distance
    void knnMatch(const float *query, int rows, std::vector<std::vector<DMatch> > &matches, int k)
    {
        matches.assign(rows, std::vector<DMatch>());
        const int nt = starts_.empty() ? 0 : starts_.back();
        if (rows == 0 || nt == 0) return;
        std::vector<DMatch> flat((size_t)rows * k);
        int rc = surf_match_knn(e_, query, rows, train_.data(), nt, dim_, k, flat.data(), SURF_MEM_HOST);
        if (rc != SURF_OK) throw Error(rc, surf_last_error(e_));
        for (int i = 0; i < rows; i++)
            for (int r = 0; r < k; r++) {
                DMatch m = flat[(size_t)i * k + r];
                if (m.train_idx < 0) continue;
                size_t img = 0;
                while (img + 1 < starts_.size() && starts_[img + 1] <= m.train_idx) img++;
                m.img_idx = (int)img;
                m.train_idx -= starts_[img];
                matches[i].push_back(m);
            }
    }

private:
    surf_engine *e_;
    int dim_;
    std::vector<float> train_;
    std::vector<int> starts_;
};

}  // namespace surfb200

#ifdef SURF_B200_WITH_OPENCV
// Not compilable in the build container (no OpenCV headers): OpenCV 2.4-style Feature2D adapter.
#include <opencv2/features2d/features2d.hpp>
namespace surfb200 {
class CvSURF : public cv::Feature2D {
public:
    double hessianThreshold;
    CvSURF(double thr, int nOctaves = 4, int nOctaveLayers = 2, bool extended = true, bool upright = false,
           int device = 0, int max_width = 1024, int max_height = 1024)
        : hessianThreshold(thr), impl_(thr, nOctaves, nOctaveLayers, extended, upright, device, max_width, max_height)
    {
        static_assert(sizeof(cv::KeyPoint) == sizeof(surf_keypoint), "cv::KeyPoint layout");
        static_assert(sizeof(cv::DMatch) == sizeof(surf_dmatch), "cv::DMatch layout");
    }
    int descriptorSize() const { return impl_.descriptorSize(); }
    int descriptorType() const { return CV_32F; }
    surf_engine *engine() const { return impl_.engine(); }
    void operator()(cv::InputArray img, cv::InputArray mask, std::vector<cv::KeyPoint> &kps, cv::OutputArray desc,
                    bool useProvided = false) const
    {
        cv::Mat im = img.getMat(), mk = mask.getMat();
        CV_Assert(!im.empty() && im.type() == CV_8UC1);
        CV_Assert(mk.empty() || (mk.type() == CV_8UC1 && mk.size() == im.size()));
        impl_.hessianThreshold = hessianThreshold;
        std::vector<KeyPoint> k(kps.size());
        if (!kps.empty()) std::memcpy(k.data(), kps.data(), kps.size() * sizeof(KeyPoint));
        std::vector<float> d;
        if (desc.needed())
            impl_(ImageView(im.data, im.cols, im.rows, im.step), ImageView(mk.data, mk.cols, mk.rows, mk.step), k, d, useProvided);
        else
            impl_.detect(ImageView(im.data, im.cols, im.rows, im.step), k, ImageView(mk.data, mk.cols, mk.rows, mk.step));
        kps.resize(k.size());
        if (!k.empty()) std::memcpy(kps.data(), k.data(), k.size() * sizeof(KeyPoint));
        if (desc.needed()) {
            if (k.empty()) { desc.release(); return; }
            desc.create((int)k.size(), impl_.descriptorSize(), CV_32F);
            std::memcpy(desc.getMat().data, d.data(), d.size() * sizeof(float));
        }
    }
protected:
    void detectImpl(const cv::Mat &image, std::vector<cv::KeyPoint> &keypoints, const cv::Mat &mask) const
    {
        (*this)(image, mask, keypoints, cv::noArray(), false);
    }
    void computeImpl(const cv::Mat &image, std::vector<cv::KeyPoint> &keypoints, cv::Mat &descriptors) const
    {
        (*this)(image, cv::Mat(), keypoints, descriptors, true);
    }
private:
    mutable SURF impl_;
};
}  // namespace surfb200
#endif  // SURF_B200_WITH_OPENCV

#endif  // SURF_B200_HPP
