"""B200-native Conditional Regression Forest inference path (Dantone et al., CVPR'12).

Drop-in for the hot path of MatrixPlayer/face_alignment_cvpr_2012: FaceForest::analyzeFace and
everything below it (channels, head-pose forest, pose-conditioned facial-feature forests, votes,
MeanShift) as hand-written sm_100a kernels behind the C ABI of include/crf_b200.h.
"""
from .capi import FACE_DTYPE, CrfError, Options, Rect, build, lib  # noqa: F401
from .face_forest import (CascadeClassifier, Context, Face, FaceDetectionOption, FaceForest, FaceForestOptions, ForestParam, HeadPoseEstimatorOption, enlarge_detections, intersect,  # noqa: F401
                          MeanShift, MeanShiftOption, Model, MultiContext, MultiPartEstimatorOption, loadConfigFile, _options)

__all__ = ["FaceForest", "FaceForestOptions", "Face", "ForestParam", "Model", "Context", "MultiContext", "MeanShift", "CrfError", "build", "lib"]
