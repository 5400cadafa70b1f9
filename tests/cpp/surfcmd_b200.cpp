// surfcmd_b200.cpp -- C++ host program over the C ABI, shaped like the reference's headless driver
// (SurfFeatures/Logic/surfcmd.cxx:3-41: computeAll = for every frame, detect + describe; then
// match every query frame against its predecessor).  Reads raw 8-bit frames from a file, prints
// one line per frame.  Built and run by tests/test_cpp_host.py (compile check on CPU, run on GPU).
//   surfcmd_b200 <raw-u8-file> <width> <height> <n_frames> [minHessian]
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "surf_b200.hpp"

int main(int argc, char **argv)
{
    if (argc < 5) {
        std::fprintf(stderr, "usage: %s raw W H N [minHessian]\n", argv[0]);
        return 2;
    }
    const int W = std::atoi(argv[2]), H = std::atoi(argv[3]), N = std::atoi(argv[4]);
    const double thr = argc > 5 ? std::atof(argv[5]) : 400.0;
    std::vector<unsigned char> buf((size_t)W * H * N);
    FILE *f = std::fopen(argv[1], "rb");
    if (!f || std::fread(buf.data(), 1, buf.size(), f) != buf.size()) {
        std::fprintf(stderr, "cannot read %s\n", argv[1]);
        return 2;
    }
    std::fclose(f);
    try {
        // the reference's parameters: detector (minHessian, 4, 2), default-constructed extractor
        surfb200::SURF detector(thr, 4, 2, false, false, 0, W, H);
        surfb200::SURF extractor(100, 4, 2, false, false, 0, W, H);
        surfb200::BFMatcherL2 matcher(detector.engine(), 64);
        std::vector<float> prev;
        int prev_n = 0;
        for (int i = 0; i < N; i++) {
            surfb200::ImageView img(buf.data() + (size_t)i * W * H, W, H, W);
            std::vector<surfb200::KeyPoint> kps;
            std::vector<float> desc;
            detector.detect(img, kps);
            extractor.compute(img, kps, desc);
            int good = 0;
            if (prev_n > 0 && !kps.empty()) {
                matcher.clear();
                matcher.add(prev.data(), prev_n);
                std::vector<std::vector<surfb200::DMatch> > m;
                matcher.knnMatch(desc.data(), (int)kps.size(), m, 2);
                for (size_t q = 0; q < m.size(); q++)
                    if (m[q].size() == 2 && m[q][0].distance < 0.8f * m[q][1].distance) good++;
            }
            double cs = 0;
            for (size_t k = 0; k < desc.size(); k++) cs += desc[k];
            std::printf("frame %d keypoints %zu desc_sum %.6f ratio_matches %d\n", i, kps.size(), cs, good);
            prev.swap(desc);
            prev_n = (int)kps.size();
        }
    } catch (const surfb200::Error &e) {
        std::fprintf(stderr, "surf_b200 error %d: %s\n", e.code, e.what());
        return e.code == SURF_ERR_NO_DEVICE ? 3 : 1;
    }
    return 0;
}
