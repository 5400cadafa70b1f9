// surfcmd_mha_b200.cpp -- the reference's headless driver flow (SurfFeatures/Logic/surfcmd.cxx) over the
// surfb200::SurfFeaturesLogic mirror: computeAll() below is the reference's function body, unchanged
// (surfcmd.cxx:3-41); main() sets the files / frame ranges / minHessian the way surfcmd.cxx:170-184 does and
// prints one line per query frame.  Built and run by tests/test_cpp_host.py.
//   surfcmd_mha_b200 bogus.mha train.mha query.mha mask.raw|- minHessian
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "surf_b200_logic.hpp"

typedef surfb200::SurfFeaturesLogic vtkSlicerSurfFeaturesLogic;

void computeAll(vtkSlicerSurfFeaturesLogic* surf)
{
  surf->computeBogus();
  surf->computeTrain();
  surf->computeQuery();

  // Compute bogus
  while(true)
  {
    if(!surf->isBogusLoading())
      break;
    surf->computeNextBogus();
  }

  // Compute Train
  while(true)
  {
    if(!surf->isTrainLoading())
      break;
    surf->computeNextTrain();
  }

  // Compute query
  while(true)
  {
    if(!surf->isQueryLoading())
      break;
    surf->computeNextQuery();
  }

  // Compute correspondences
  surf->startInterSliceCorrespondence();
  while(true)
  {
    if(!surf->isCorrespondenceComputing())
      break;
    surf->computeNextInterSliceCorrespondence();
  }
}

int main(int argc, char **argv)
{
    if (argc < 6) {
        std::fprintf(stderr, "usage: %s bogus.mha train.mha query.mha mask.raw|- minHessian\n", argv[0]);
        return 2;
    }
    try {
        vtkSlicerSurfFeaturesLogic surf(0, 640, 480, 4, 8192, 1L << 18, 512);
        std::vector<unsigned char> mask;
        surf_mha *probe = 0;
        surf_mha_info info;
        if (surf_mha_open(argv[2], -1, -1, 0, &probe, &info) != SURF_OK) {
            std::fprintf(stderr, "cannot read %s\n", argv[2]);
            return 2;
        }
        surf_mha_close(probe);
        if (std::strcmp(argv[4], "-") != 0) {
            mask.resize((size_t)info.width * info.height);
            FILE *f = std::fopen(argv[4], "rb");
            if (!f || std::fread(mask.data(), 1, mask.size(), f) != mask.size()) {
                std::fprintf(stderr, "cannot read mask %s\n", argv[4]);
                return 2;
            }
            std::fclose(f);
            surf.setMask(mask.data(), info.width, info.height, info.width);
        }
        surf.setBogusFile(argv[1]);
        surf.setTrainFile(argv[2]);
        surf.setQueryFile(argv[3]);
        surf.setBogusStartFrame(0);
        surf.setTrainStartFrame(0);
        surf.setQueryStartFrame(0);
        surf.setBogusStopFrame(-1);
        surf.setTrainStopFrame(-1);
        surf.setQueryStopFrame(-1);
        surf.setMinHessian(std::atoi(argv[5]));
        computeAll(&surf);
        std::printf("centroid %d %d train_rows %ld\n", surf.xcentroid, surf.ycentroid, surf.trainRows());
        for (size_t i = 0; i < surf.bestMatches.size(); i++) {
            int alive = 0;
            for (size_t k = 0; k < surf.afterHoughMatches[i].size(); k++) alive += surf.afterHoughMatches[i][k].train_idx >= 0;
            std::printf("query %zu keypoints %d filtered %zu hough %d alive %d best %d count %d with_best %zu\n", i,
                        surf.queryOffsets()[i + 1] - surf.queryOffsets()[i], surf.afterHoughMatches[i].size(), surf.houghCounts[i],
                        alive, surf.bestMatches[i], surf.bestMatchesCount[i], surf.matchesWithBestTrainImage[i].size());
        }
        // row f4: the logic's ransac() on a synthetic tilted plane through query frame 0's keypoints
        {
            const int nk = surf.queryOffsets()[1] < 200 ? surf.queryOffsets()[1] : 200;
            std::vector<surfb200::SurfFeaturesLogic::Point3> pts(nk);
            for (int j = 0; j < nk; j++) {
                const surfb200::KeyPoint &k = surf.queryKeypoints()[j];
                pts[j].v[0] = k.x;
                pts[j].v[1] = k.y;
                pts[j].v[2] = 0.05 * k.x - 0.02 * k.y + (j % 5 == 0 ? 30.0 : 0.0);
            }
            std::vector<int> inl, plane;
            const int rc = surf.ransac(pts, inl, plane);
            std::printf("ransac rc %d points %d inliers %zu plane %d %d %d\n", rc, nk, inl.size(), plane.size() == 3 ? plane[0] : -1,
                        plane.size() == 3 ? plane[1] : -1, plane.size() == 3 ? plane[2] : -1);
        }
    } catch (const surfb200::Error &e) {
        std::fprintf(stderr, "surf_b200 error %d: %s\n", e.code, e.what());
        return e.code == SURF_ERR_NO_DEVICE ? 3 : 1;
    }
    return 0;
}
