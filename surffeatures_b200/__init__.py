"""surffeatures_b200 -- B200-native SURF detect / describe / L2-match engine.

Drop-in for the OpenCV calls stropitek/SurfFeatures makes from its logic class
(SurfFeatures/Logic/vtkSlicerSurfFeaturesLogic.cxx:1381-1390, :635-643): a C-ABI library of
hand-written sm_100a CUDA kernels (csrc/, include/surf_b200.h) plus this thin host-side mirror of
the cv::Feature2D / DescriptorMatcher interface.  No CPU fallback.
"""
from .engine import (  # noqa: F401
    ABI_SYMBOLS, BFMatcher, Comm, DescriptorDB, DMATCH_DTYPE, KEYPOINT_DTYPE, LIB_PATH, MEM_DEVICE, MEM_HOST, SURF,
    SurfError, load_library, shard_range, to_cv2_keypoints,
)

__all__ = ["SURF", "BFMatcher", "SurfError", "KEYPOINT_DTYPE", "DMATCH_DTYPE", "load_library", "LIB_PATH",
           "ABI_SYMBOLS", "MEM_HOST", "MEM_DEVICE", "to_cv2_keypoints", "Comm", "DescriptorDB", "shard_range"]
